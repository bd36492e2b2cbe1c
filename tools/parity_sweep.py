"""Randomised shape sweep of the joint + loss path against the CPU oracle: joint widths on every route (fused widths,
streamed-product widths, the library-GEMM fallback width 768), vocabularies from 2 up, ragged lengths incl. T = 1 / U = 0,
bf16 inputs, per-utterance grad_output spread over four orders of magnitude with both signs.
    python tools/parity_sweep.py SEED N          small shapes
    python tools/parity_sweep.py SEED N big      hundreds of 128-row tiles, P' kept whole / in chunks / not at all
Run under gpurun.  Round 2: 114 shapes, 111 within the tests' tolerances (loss 1e-4, gradients 1e-3 relative L2 per tensor
and per utterance); the other three are at 1.4e-3 - 2.3e-3 on one tensor, identical on all three routes, i.e. the 16-bit
operand rounding on an ill-conditioned instance (the fp32 oracle itself is 2e-4 from the float64 one there)."""
import os, sys, random, json, traceback
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_parity as t

rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
bad = 0
big = len(sys.argv) > 3 and sys.argv[3] == "big"
for case in range(n):
    H = rng.choice([64, 128, 256, 384, 512, 512, 512, 1024, 1024, 1536, 768])
    V = rng.choice([2, 3, 17, 63, 64, 65, 127, 129, 255, 256, 257, 300, 511, 513, 700, 1000, 1025])
    B = rng.randint(1, 5)
    T = rng.randint(1, 70)
    U = rng.randint(0, 12)
    D = rng.choice([32, 64, 512])
    if big:                      # many more 128-row tiles than CTA pairs, P' kept whole or in chunks
        H = rng.choice([512, 512, 1024, 256])
        V = rng.choice([129, 300, 513, 600])
        B, T, U, D = rng.randint(2, 4), rng.randint(150, 400), rng.randint(20, 45), 64
        os.environ["TTX_KEEP_GB"] = rng.choice(["32", "0.004", "0.02", "1e-9"])
    al = [rng.randint(1, T) for _ in range(B)]; al[rng.randrange(B)] = T
    ll = [rng.randint(0, U) for _ in range(B)]; ll[rng.randrange(B)] = U
    dtype = torch.bfloat16 if rng.random() < 0.25 else torch.float32
    wts = torch.tensor([10.0 ** rng.uniform(-2, 2) * rng.choice([1, 1, -1]) for _ in range(B)])
    if os.environ.get("SWEEP_UNIT_WTS") == "1":
        wts = torch.ones(B)
    elif os.environ.get("SWEEP_UNIT_WTS") == "abs":
        wts = wts.abs()
    desc = dict(H=H, V=V, B=B, T=T, U=U, D=D, al=al, ll=ll, dtype=str(dtype), keep=os.environ.get("TTX_KEEP_GB"))
    try:
        args = t._espnet_case(B, T, U, V, D, H, al, ll, seed=case)
        errs, _ = t._run_pair(*args, weights=wts, dtype=dtype)
        lt, gt = (t.LOSS_TOL, t.GRAD_TOL) if dtype == torch.float32 else (2e-2, 3e-2)
        worst = max(v for k, v in errs.items() if k != "loss")
        ok = errs["loss"] < lt and worst < gt and all(v == v for v in errs.values())
        desc["wts"] = [float("%.3g" % x) for x in wts.tolist()]
        print(("ok  " if ok else "FAIL"), json.dumps(desc), "loss %.2e worst grad %.2e" % (errs["loss"], worst), flush=True)
        if not ok:
            bad += 1
            print("     ", {k: float("%.3g" % v) for k, v in errs.items() if v != v or v > (gt if k != "loss" else lt)})
        if big:
            print("     ", {k: float("%.3g" % v) for k, v in errs.items() if k.startswith("d_pred") or k.startswith("d_enc")})
    except Exception as e:
        bad += 1
        print("EXC ", json.dumps(desc), repr(e)[:300], flush=True)
        traceback.print_exc(limit=3)
print("bad", bad, "of", n)
