"""Greedy- and beam-search timing on cuda:0: the reference's per-frame loops (tt/model.py:70-90, :110-179) vs the rebound ones
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


def run(beam=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        out = model.recognize_beam_search(inputs, lengths) if beam else model.recognize(inputs, lengths)
    torch.cuda.synchronize()
    return time.perf_counter() - t0, out


for beam in (False, True):                            # tt/model.py:92-108 ; :181-198 (beam width 5)
    run(beam)
    ref_s, want = min((run(beam) for _ in range(3)), key=lambda r: r[0])
    ttb.install(patch_espnet=False)
    run(beam)
    our_s, got = min((run(beam) for _ in range(3)), key=lambda r: r[0])
    ttb.uninstall()
    print(json.dumps({"workload": "%s B=%d T=%d V=%d joint 1024 (1-layer encoder / decoder)"
                      % ("recognize_beam_search (width 5)" if beam else "greedy recognize", B, T, V),
                      "labels_emitted": [len(x) for x in want], "identical": got == want,
                      "reference_ms_per_utt": 1e3 * ref_s / B, "ours_ms_per_utt": 1e3 * our_s / B,
                      "speedup": ref_s / our_s}))
