"""Model-level training step on cuda:0 ("model-level bench"): the reference's own `Transducer`
(tt/model.py:40-68: encoder + label encoder + joint) built from config/aishell.yaml (configs[0]) and
config/joint_streaming.yaml (configs[2]) with the streaming mask passed explicitly (tt/utils.py:242-251, left 10 /
right 2), criterion as in train.py:53, backward, SGD step.  Dropout as in the YAML (training mode).  Arms:

  dense   the reference's classes untouched (dense B x T x U1 x V logits, dense T x T attention) with
          torchaudio.functional.rnnt_loss standing in for the uninstalled warprnnt_pytorch -- what the reference's
          algorithm costs on this GPU; only where its logits fit comfortably (cfg1; cfg3 at the per-GPU batch of 8)
  joint   install(patch_attention=False): this repo's JointNet + RNNTLoss, dense attention
  ours    install(): joint, loss and the banded attention core

Needs baseline/_ref (staged reference).  One JSON line per (workload, arm); run under gpurun.
"""
import json
import os
import sys

import torch
import torch.nn.functional as F
import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import transformer_transducer_b200 as ttb  # noqa: E402
from oracle import ref_import  # noqa: E402

ref_import.prepare(stub_train_deps=True)
tt_model = ref_import.tt_model()
import tt.utils as tu  # noqa: E402

DEV = "cuda"


def build(cfg_file, vocab):
    cfg = tu.AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", cfg_file))))
    cfg.model.vocab_size = vocab
    torch.manual_seed(0)
    return tt_model.Transducer(cfg.model).to(DEV).train()


def run(name, cfg_file, B, T, U, V, arm, steps=8, warmup=3):
    import tt.model as tm
    if arm == "dense":
        import torchaudio
        model = build(cfg_file, V)

        def criterion(logits, targets, in_len, tgt_len):
            return torchaudio.functional.rnnt_loss(logits, targets, in_len, tgt_len, blank=0, reduction="mean")
    else:
        ttb.install(patch_espnet=False, patch_attention=(arm == "ours"))
        try:
            model = build(cfg_file, V)
        finally:
            pass
        assert isinstance(model.joint, ttb.JointNet)
        criterion = ttb.RNNTLoss()
    opt = torch.optim.SGD(model.parameters(), lr=1e-4)
    torch.manual_seed(1)
    inputs = torch.randn(B, T, 512, device=DEV)
    targets = torch.randint(1, V, (B, U), device=DEV)
    in_len = torch.full((B,), T, dtype=torch.int32, device=DEV)
    tgt_len = torch.full((B,), U, dtype=torch.int32, device=DEV)

    def step():
        opt.zero_grad(set_to_none=True)
        padded = F.pad(targets, pad=[1, 0, 0, 0], value=0)                                   # tt/model.py:59
        enc = model.encoder(inputs, tu.context_mask(inputs)[:, :, None])                     # :60 (streaming) ,:63
        dec = model.decoder(padded, tu.look_ahead_mask(padded)[:, :, None])                  # :62,:64
        logits = model.joint(enc, dec)                                                       # :66
        loss = criterion(logits, targets.int(), in_len, tgt_len)                             # train.py:53
        loss.backward()
        opt.step()
        return loss

    try:
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out = {"workload": name, "arm": arm, "B": B, "T": T, "U": U, "V": V, "ms_per_step": ms, "utt_per_s": 1e3 * B / ms,
               "peak_GiB": torch.cuda.max_memory_allocated() / 2**30, "loss": float(loss.detach().float().mean())}
    except torch.cuda.OutOfMemoryError:
        out = {"workload": name, "arm": arm, "B": B, "out_of_memory": True}
    finally:
        if arm != "dense":
            ttb.uninstall()
        del model, opt
        torch.cuda.empty_cache()
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    run("configs[0] aishell.yaml model step", "aishell.yaml", 4, 200, 30, 4232, "dense")
    run("configs[0] aishell.yaml model step", "aishell.yaml", 4, 200, 30, 4232, "joint")
    run("configs[0] aishell.yaml model step", "aishell.yaml", 4, 200, 30, 4232, "ours")
    run("configs[2] joint_streaming.yaml model step, per-GPU batch at 8 GPUs", "joint_streaming.yaml", 8, 410, 42, 6485, "dense")
    run("configs[2] joint_streaming.yaml model step, per-GPU batch at 8 GPUs", "joint_streaming.yaml", 8, 410, 42, 6485, "joint")
    run("configs[2] joint_streaming.yaml model step, per-GPU batch at 8 GPUs", "joint_streaming.yaml", 8, 410, 42, 6485, "ours")
    run("configs[2] joint_streaming.yaml model step, global batch", "joint_streaming.yaml", 64, 410, 42, 6485, "joint", steps=4)
    run("configs[2] joint_streaming.yaml model step, global batch", "joint_streaming.yaml", 64, 410, 42, 6485, "ours", steps=4)
