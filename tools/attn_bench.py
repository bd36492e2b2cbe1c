"""One streaming encoder attention block (tt/transformer.py:106-177) forward + backward on cuda:0 under the context mask
(left 10 / right 2): the reference's dense T x T path vs the rebound banded forward, joint_streaming.yaml's encoder dims
(d_model 512, 8 heads x 64, max_len 410) at the global batch of configs[2].  Needs baseline/_ref.  One JSON line."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import transformer_transducer_b200 as ttb  # noqa: E402
import test_gpu_callers as t  # noqa: E402

t.ref_import.prepare(stub_train_deps=True)
import tt.transformer as ttr  # noqa: E402
import tt.utils as tu  # noqa: E402

T, B, n_head, d_head, d_model, max_len = 410, 64, 8, 64, 512, 410
torch.manual_seed(0)
attn = ttr.RelLearnableMultiHeadAttn(n_head, d_model, d_head, dropout=0.0).cuda()
r_emb = torch.randn(max_len, n_head, d_head, device="cuda", requires_grad=True)
r_w_bias = torch.randn(n_head, d_head, device="cuda", requires_grad=True)
r_bias = torch.randn(max_len, n_head, device="cuda", requires_grad=True)
w = torch.randn(T, B, d_model, device="cuda", requires_grad=True)
g = torch.randn(T, B, d_model, device="cuda")
mask = tu.context_mask(torch.empty(1, T, 1, device="cuda"))[:, :, None]


def step():
    out = attn(w, r_emb, r_w_bias, r_bias, attn_mask=mask)
    out.backward(g)
    return out


def timed(n=10):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, torch.cuda.max_memory_allocated() / 2**20, out.detach().clone()


ref_ms, ref_mb, want = timed()
ttb.install(patch_tt=False, patch_espnet=False, patch_decode=False, patch_data=False)
our_ms, our_mb, got = timed()
ttb.uninstall()
err = float((got - want).norm() / want.norm())
print(json.dumps({"workload": "RelLearnableMultiHeadAttn fwd+bwd T=%d B=%d heads=%dx%d context=(10,2)" % (T, B, n_head, d_head),
                  "reference_ms": ref_ms, "ours_ms": our_ms, "speedup": ref_ms / our_ms, "reference_peak_MiB": ref_mb,
                  "ours_peak_MiB": our_mb, "rel_l2_output": err}))

# ---- espnet side: RelPositionMultiHeadedAttention (attention.py:212-308) under padding + context mask, espnet_aishell.yaml's
# encoder dims (512, 8 heads), same T / B
from espnet.nets.pytorch_backend.nets_utils import make_attention_mask, make_pad_mask  # noqa: E402
from espnet.nets.pytorch_backend.transformer import attention as eatt  # noqa: E402
from espnet.nets.pytorch_backend.transformer.embedding import RelPositionalEncoding  # noqa: E402

torch.manual_seed(1)
eattn = eatt.RelPositionMultiHeadedAttention(n_head, d_model, 0.0).cuda()
x = torch.randn(B, T, d_model, device="cuda", requires_grad=True)
ge = torch.randn(B, T, d_model, device="cuda")
_, pos_emb = RelPositionalEncoding(d_model, 0.0).cuda()(x.detach())
ilens = torch.randint(T // 2, T + 1, (B,), device="cuda")
ilens[0] = T
emask = (~make_pad_mask(ilens)[:, None, :]).cuda() & ~make_attention_mask(x, 10, 2)[None, :, :]


def step():  # noqa: F811
    out = eattn(x, x, x, pos_emb, emask)
    out.backward(ge)
    return out


ref_ms, ref_mb, want = timed()
ttb.install(patch_tt=False, patch_espnet=False, patch_decode=False, patch_data=False)
our_ms, our_mb, got = timed()
ttb.uninstall()
err = float((got - want).norm() / want.norm())
print(json.dumps({"workload": "espnet RelPositionMultiHeadedAttention fwd+bwd T=%d B=%d heads=%dx%d padding + context=(10,2)"
                  % (T, B, n_head, d_model // n_head),
                  "reference_ms": ref_ms, "ours_ms": our_ms, "speedup": ref_ms / our_ms, "reference_peak_MiB": ref_mb,
                  "ours_peak_MiB": our_mb, "rel_l2_output": err}))
