"""Pins the oracle (test infrastructure) before anything is checked against it:
known-answer vector, fixtures produced by the UNMODIFIED reference modules, torchaudio's independent
CPU implementation, the float64 restatement and a finite-difference check."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import joint_ref, ref_import, rnnt_oracle


def _i32(x):
    return torch.as_tensor(x, dtype=torch.int32)


def test_known_answer_vector(golden_dir):
    k = json.load(open(os.path.join(golden_dir, "warp_transducer_kat.json")))
    acts = torch.tensor(k["acts"], requires_grad=True)
    c = rnnt_oracle.rnnt_loss(acts, _i32(k["labels"]), _i32(k["act_lens"]), _i32(k["label_lens"]), k["blank"], "none")
    c.sum().backward()
    assert abs(float(c[0]) - k["cost_upstream"]) < 2e-6
    assert abs(k["cost_torchaudio"] - k["cost_upstream"]) < 2e-6
    np.testing.assert_allclose(acts.grad.numpy(), np.array(k["grads_upstream"], dtype=np.float32), atol=1e-6)
    np.testing.assert_allclose(acts.grad.numpy(), np.array(k["grads_torchaudio"], dtype=np.float32), atol=1e-6)
    a64 = torch.tensor(k["acts"], dtype=torch.float64, requires_grad=True)
    c64 = rnnt_oracle.rnnt_loss_fp64(a64, _i32(k["labels"]), _i32(k["act_lens"]), _i32(k["label_lens"]))
    c64.sum().backward()
    assert abs(float(c64[0]) - k["cost_upstream"]) < 1e-6
    np.testing.assert_allclose(a64.grad.numpy(), np.array(k["grads_upstream"]), atol=1e-6)


def test_reductions_and_shapes(golden_dir):
    g = np.load(os.path.join(golden_dir, "ragged_loss.npz"))
    args = (torch.tensor(g["logits"]), _i32(g["labels"]), _i32(g["act_lens"]), _i32(g["label_lens"]))
    none = rnnt_oracle.rnnt_loss(*args, reduction="none")
    mean = rnnt_oracle.RNNTLoss()(*args)
    total = rnnt_oracle.RNNTLoss(reduction="sum")(*args)
    assert none.shape == (3,) and mean.shape == (1,) and total.shape == (1,)
    assert torch.allclose(total, none.sum().view(1)) and torch.allclose(mean, none.sum().view(1) / 3)
    with pytest.raises(ValueError):
        rnnt_oracle.rnnt_loss(*args, reduction="avg")


def test_ragged_against_torchaudio_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "ragged_loss.npz"))
    logits = torch.tensor(g["logits"], requires_grad=True)
    c = rnnt_oracle.rnnt_loss(logits, _i32(g["labels"]), _i32(g["act_lens"]), _i32(g["label_lens"]), 0, "none")
    c.sum().backward()
    np.testing.assert_allclose(c.detach().numpy(), g["costs"], rtol=2e-6)
    np.testing.assert_allclose(logits.grad.numpy(), g["grads"], atol=5e-6)
    # exact zeros outside the ragged region (t >= T_b or u > U_b)
    gr = logits.grad
    assert gr[1, 6:].abs().max() == 0 and gr[1, :, 3:].abs().max() == 0
    assert gr[2, 1:].abs().max() == 0 and gr[2, :, 1:].abs().max() == 0


def test_fp64_restatement_agrees_and_gradchecks(golden_dir):
    g = np.load(os.path.join(golden_dir, "ragged_loss.npz"))
    labels, al, ll = _i32(g["labels"]), _i32(g["act_lens"]), _i32(g["label_lens"])
    l64 = torch.tensor(g["logits"], dtype=torch.float64, requires_grad=True)
    c64 = rnnt_oracle.rnnt_loss_fp64(l64, labels, al, ll)
    c64.sum().backward()
    np.testing.assert_allclose(c64.detach().numpy(), g["costs"], rtol=2e-6)
    np.testing.assert_allclose(l64.grad.numpy(), g["grads"], atol=5e-6)
    small = torch.randn(2, 4, 3, 5, dtype=torch.float64, requires_grad=True)
    lab = _i32([[1, 2], [3, -1]])
    assert torch.autograd.gradcheck(lambda x: rnnt_oracle.rnnt_loss_fp64(x, lab, _i32([4, 3]), _i32([2, 1])), (small,))


def test_lattice_f64_matches_restatement():
    torch.manual_seed(3)
    B, T, U1, V = 2, 6, 4, 7
    logits = torch.randn(B, T, U1, V, dtype=torch.float64)
    labels, al, ll = _i32([[2, 2, 5], [1, 6, -1]]), _i32([6, 4]), _i32([3, 2])
    lp = torch.log_softmax(logits, -1)
    lpb = lp[..., 0]
    lab_ext = torch.cat([labels.clamp(min=0).long(), torch.zeros(B, 1, dtype=torch.long)], 1)
    lpl = lp.gather(3, lab_ext.view(B, 1, U1, 1).expand(B, T, U1, 1)).squeeze(3)
    alpha, beta, costs = rnnt_oracle.lattice_fp64(lpb, lpl, al, ll)
    ref = rnnt_oracle.rnnt_loss_fp64(logits, labels, al, ll)
    np.testing.assert_allclose(costs.numpy(), ref.numpy(), rtol=1e-12)
    np.testing.assert_allclose((-beta[:, 0, 0]).numpy(), ref.numpy(), rtol=1e-12)


def test_certify_inputs():
    acts = torch.zeros(2, 3, 2, 5)
    lab, al, ll = _i32([[1], [1]]), _i32([3, 2]), _i32([1, 1])
    rnnt_oracle.certify_inputs(acts, lab, al, ll)
    with pytest.raises(TypeError):
        rnnt_oracle.certify_inputs(acts, lab.long(), al, ll)
    with pytest.raises(ValueError):
        rnnt_oracle.certify_inputs(acts, lab, _i32([2, 2]), ll)
    with pytest.raises(ValueError):
        rnnt_oracle.certify_inputs(acts, lab, al, _i32([0, 0]))
    with pytest.raises(ValueError):
        rnnt_oracle.certify_inputs(acts, lab, _i32([3]), ll)


def _load_sd(module, g):
    module.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd_")})


def test_tt_joint_restatement_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "tt_joint.npz"))
    m = joint_ref.TTJointNet(48, 40, 23)
    _load_sd(m, g)
    out = m(torch.tensor(g["enc"]), torch.tensor(g["dec"]))
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    out1 = m(torch.tensor(g["enc"])[1, 2].view(-1), torch.tensor(g["dec"])[1, 3].view(-1))
    np.testing.assert_allclose(out1.detach().numpy(), g["logits_1d"], rtol=1e-5, atol=1e-6)


def test_espnet_joint_restatement_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "espnet_joint.npz"))
    m = joint_ref.EspnetJointNetwork(19, 24, 20, 32, "tanh")
    _load_sd(m, g)
    out = m(torch.tensor(g["h_enc"]), torch.tensor(g["h_dec"]))
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)


def test_espnet_transloss_fixture(golden_dir):
    """Fixture came from the reference TransLoss wrapper driving the oracle; re-derive it directly."""
    g = np.load(os.path.join(golden_dir, "espnet_transloss.npz"))
    pred = torch.tensor(g["pred"], requires_grad=True)
    loss = rnnt_oracle.RNNTLoss(blank=0)(pred, _i32(g["target"]), _i32(g["pred_len"]), _i32(g["target_len"]))
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g["loss"], rtol=1e-6)
    np.testing.assert_allclose(pred.grad.numpy(), g["grad"], atol=1e-7)
    c64 = rnnt_oracle.rnnt_loss_fp64(pred.detach(), _i32(g["target"]), _i32(g["pred_len"]), _i32(g["target_len"]))
    assert abs(float(c64.sum() / 2) - float(g["loss"][0])) < 1e-5


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_restatement_against_live_reference_modules():
    torch.manual_seed(5)
    ref = ref_import.tt_model().JointNet(64, 48, 29)
    mine = joint_ref.TTJointNet(64, 48, 29)
    mine.load_state_dict(ref.state_dict())
    enc, dec = torch.randn(2, 9, 32), torch.randn(2, 4, 32)
    assert torch.allclose(ref(enc, dec), mine(enc, dec), atol=1e-6)
    ref2 = ref_import.espnet_joint_module().JointNetwork(31, 32, 32, 40, "tanh")
    mine2 = joint_ref.EspnetJointNetwork(31, 32, 32, 40, "tanh")
    mine2.load_state_dict(ref2.state_dict())
    assert torch.allclose(ref2(enc[:, :, None], dec[:, None]), mine2(enc[:, :, None], dec[:, None]), atol=1e-6)
