"""Greedy-search timing on cuda:0: the reference's per-frame loop (tt/model.py:70-90) vs the rebound decode
(transformer_transducer_b200/decode.py) on the same model and encoder states.  Needs baseline/_ref (staged reference).
Prints one JSON line; run under gpurun."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import transformer_transducer_b200 as ttb  # noqa: E402
import test_gpu_callers as t  # noqa: E402

t.ref_import.prepare(stub_train_deps=True)
tt_model = t.ref_import.tt_model()
V, T, B = 4232, 200, 4                                       # aishell.yaml dims (configs[0]): joint 1024 -> 1024 -> 4232
cfg = t._tt_config(1024, V)
t._seed(0)
model = tt_model.Transducer(cfg.model).cuda().eval()
t._boost_blank(model.joint.project_layer, 1.3)      # ~1 label per 7 frames, like configs[0]'s U / T
inputs = torch.randn(B, T, 512, device="cuda")
lengths = [T] * B


def run():
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        out = model.recognize(inputs, lengths)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


run()
ref_s, want = min((run() for _ in range(3)), key=lambda r: r[0])
ttb.install(patch_espnet=False)
run()
our_s, got = min((run() for _ in range(3)), key=lambda r: r[0])
ttb.uninstall()
print(json.dumps({"workload": "greedy recognize B=%d T=%d V=%d joint 1024 (1-layer encoder / decoder)" % (B, T, V),
                  "labels_emitted": [len(x) for x in want], "identical": got == want,
                  "reference_ms_per_utt": 1e3 * ref_s / B, "ours_ms_per_utt": 1e3 * our_s / B, "speedup": ref_s / our_s}))
