"""Generates the committed fixtures in tests/golden/ (run in the BUILD container, where
/root/reference and torchaudio exist):  python tests/golden/make_golden.py

  warp_transducer_kat.json  upstream warp-transducer known-answer vector,
                            re-verified here with torchaudio.functional.rnnt_loss.
  tt_joint.npz              UNMODIFIED /root/reference/tt/model.py::JointNet on seeded inputs
                            (3-D training branch and 1-D decode branch).
  espnet_joint.npz          UNMODIFIED reference espnet JointNetwork on seeded inputs.
  ragged_loss.npz           ragged batch (lens [9,6,1]/[4,2,0], -1 label padding, repeated labels):
                            costs/grads from torchaudio's CPU rnnt_loss (independent implementation).
  espnet_transloss.npz      reference TransLoss("warp-transducer") driven through the oracle RNNTLoss
                            installed as `warprnnt_pytorch` -- pins argument order / reduction.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import, rnnt_oracle  # noqa: E402


def kat():
    import torchaudio
    acts = [[[[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.6, 0.1, 0.1], [0.1, 0.1, 0.2, 0.8, 0.1]],
             [[0.1, 0.6, 0.1, 0.1, 0.1], [0.1, 0.1, 0.2, 0.1, 0.1], [0.7, 0.1, 0.2, 0.1, 0.1]]]]
    a = torch.tensor(acts, requires_grad=True)
    labels = torch.tensor([[1, 2]], dtype=torch.int32)
    al = torch.tensor([2], dtype=torch.int32)
    ll = torch.tensor([2], dtype=torch.int32)
    c = torchaudio.functional.rnnt_loss(a, labels, al, ll, blank=0, reduction="none", fused_log_softmax=True)
    c.sum().backward()
    out = {"acts": acts, "labels": [[1, 2]], "act_lens": [2], "label_lens": [2], "blank": 0,
           "cost_upstream": 4.495666,
           "cost_torchaudio": float(c[0]), "grads_torchaudio": a.grad.tolist(),
           "grads_upstream": [[[[-0.13116688, -0.3999269, 0.17703125, 0.17703125, 0.17703125],
                                [-0.18572757, 0.12247056, -0.18168412, 0.12247056, 0.12247056],
                                [-0.32091254, 0.06269141, 0.06928472, 0.12624499, 0.06269141]],
                               [[0.05456069, -0.21824276, 0.05456069, 0.05456069, 0.05456069],
                                [0.12073959, 0.12073959, -0.48295835, 0.12073959, 0.12073959],
                                [-0.6925882, 0.16871116, 0.18645467, 0.16871116, 0.16871116]]]]}
    json.dump(out, open(os.path.join(HERE, "warp_transducer_kat.json"), "w"), indent=1)


def sd_np(m):
    return {"sd_" + k: v.detach().numpy() for k, v in m.state_dict().items()}


def tt_joint():
    torch.manual_seed(11)
    JointNet = ref_import.tt_model().JointNet
    m = JointNet(input_size=48, inner_dim=40, vocab_size=23)
    enc = torch.randn(3, 7, 24)
    dec = torch.randn(3, 5, 24)
    out3 = m(enc, dec)
    out1 = m(enc[1, 2].view(-1), dec[1, 3].view(-1))
    np.savez(os.path.join(HERE, "tt_joint.npz"), enc=enc.numpy(), dec=dec.numpy(),
             logits=out3.detach().numpy(), logits_1d=out1.detach().numpy(), **sd_np(m))


def espnet_joint():
    torch.manual_seed(12)
    JointNetwork = ref_import.espnet_joint_module().JointNetwork
    m = JointNetwork(vocab_size=19, encoder_output_size=24, decoder_output_size=20, joint_space_size=32,
                     joint_activation_type="tanh")
    h_enc = torch.randn(2, 6, 1, 24)
    h_dec = torch.randn(2, 1, 4, 20)
    z = m(h_enc, h_dec)
    np.savez(os.path.join(HERE, "espnet_joint.npz"), h_enc=h_enc.numpy(), h_dec=h_dec.numpy(),
             logits=z.detach().numpy(), **sd_np(m))


def ragged():
    import torchaudio
    torch.manual_seed(13)
    B, T, U, V = 3, 9, 4, 11
    logits = (torch.randn(B, T, U + 1, V) * 2).requires_grad_()
    labels = torch.tensor([[3, 3, 7, 1], [5, 5, -1, -1], [-1, -1, -1, -1]], dtype=torch.int32)
    al = torch.tensor([9, 6, 1], dtype=torch.int32)
    ll = torch.tensor([4, 2, 0], dtype=torch.int32)
    c = torchaudio.functional.rnnt_loss(logits, labels.clamp(min=0), al, ll, blank=0, reduction="none",
                                        fused_log_softmax=True)
    c.sum().backward()
    np.savez(os.path.join(HERE, "ragged_loss.npz"), logits=logits.detach().numpy(), labels=labels.numpy(),
             act_lens=al.numpy(), label_lens=ll.numpy(), costs=c.detach().numpy(), grads=logits.grad.numpy())


def espnet_transloss():
    torch.manual_seed(14)
    shim = types.ModuleType("warprnnt_pytorch")
    shim.RNNTLoss = rnnt_oracle.RNNTLoss
    saved = sys.modules.get("warprnnt_pytorch")
    sys.modules["warprnnt_pytorch"] = shim
    try:
        TransLoss = ref_import.espnet_loss_module().TransLoss
        crit = TransLoss("warp-transducer", 0)
        B, T, U, V = 2, 8, 3, 13
        pred = torch.randn(B, T, U + 1, V, requires_grad=True)
        target = torch.tensor([[4, 2, 9], [1, 1, -1]], dtype=torch.int32)
        pl = torch.tensor([8, 5], dtype=torch.int32)
        tl = torch.tensor([3, 2], dtype=torch.int32)
        loss = crit(pred, target, pl, tl)
        loss.backward()
        np.savez(os.path.join(HERE, "espnet_transloss.npz"), pred=pred.detach().numpy(), target=target.numpy(),
                 pred_len=pl.numpy(), target_len=tl.numpy(), loss=loss.detach().numpy(), grad=pred.grad.numpy())
    finally:
        if saved is None:
            del sys.modules["warprnnt_pytorch"]
        else:
            sys.modules["warprnnt_pytorch"] = saved


if __name__ == "__main__":
    kat()
    tt_joint()
    espnet_joint()
    ragged()
    espnet_transloss()
    print("fixtures written to", HERE)
