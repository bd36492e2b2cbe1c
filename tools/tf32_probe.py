import os, sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import test_gpu_parity as T
for tf in ("0", "1", "3"):
    os.environ["TTX_TF32_PROJ"] = tf
    for (B, Tt, U, V, D, H, al, ll, seed) in [(3, 47, 10, 1100, 64, 512, [47, 31, 6], [10, 8, 1], 21), (2, 60, 12, 2000, 512, 512, [60, 44], [12, 7], 5), (2, 40, 8, 500, 512, 256, [40, 31], [8, 5], 9)]:
        case = T._espnet_case(B, Tt, U, V, D, H, al, ll, seed=seed)
        errs, _ = T._run_pair(*case)
        print("tf32=%s D=%d H=%d" % (tf, D, H), {k: "%.1e" % v for k, v in errs.items()})
