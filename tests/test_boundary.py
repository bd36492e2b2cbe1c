"""CPU-side tests of the drop-in boundary and host logic (no GPU compute)."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch

import transformer_transducer_b200 as ttb
import warprnnt_pytorch
from oracle import ref_import, rnnt_oracle
from transformer_transducer_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _i32(x):
    return torch.as_tensor(x, dtype=torch.int32)


def test_library_exports_every_declared_symbol():
    ttb.build()
    header = open(os.path.join(ROOT, "include", "ttx.h")).read()
    declared = set(re.findall(r"\b(ttx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    lib = _lib.get()
    assert lib.ttx_version() == 1
    assert [h for h in range(32, 1100, 32) if lib.ttx_supported_h(h)] == [64, 128, 192, 256, 384, 512]
    assert lib.ttx_tiles_upper_bound(32, 400, 41) == 32 * 129
    assert lib.ttx_meta_ints(32, 32 * 129) == 4 + 2 * 33 + 32 * 129
    assert lib.ttx_lattice_elems_upper_bound(32, 400, 41) == 32 * 440 * 44


def test_c_abi_argument_errors_without_gpu():
    lib = _lib.get()
    rc = lib.ttx_joint_lse_fwd(None, None, None, None, None, None, 1, 512, 10, 0, 0, None, None, None, 0, None)
    assert rc == 1 and b"null pointer" in lib.ttx_last_error()
    rc = lib.ttx_prepare(None, None, 1, 1, 1, 1, None, 0, None)
    assert rc == 1


def test_kept_matrix_chunking_and_entry_points_reject_bad_arguments():
    """Host logic of the streamed products: the 16-bit softmax-numerator matrix covers the batch when it fits the budget
    (one range, kept for the backward), otherwise even-aligned tile ranges of at least one tile pair; the C entry points
    check their arguments before any launch."""
    from transformer_transducer_b200 import functional as F
    assert F._chunk_ranges(4100, 4352, 32 * 2**30) == [(0, 4100)]                   # cfg2: 4.57 GB
    r = F._chunk_ranges(4100, 4352, 1 * 2**30)
    assert r[0] == (0, 962) and all(t0 % 2 == 0 for t0, _ in r) and sum(n for _, n in r) == 4100
    assert F._chunk_ranges(5, 4352, 0.0) == [(0, 2), (2, 2), (4, 1)]
    lib = _lib.get()
    null = ctypes.c_void_p(0)
    rc = lib.ttx_joint_fwd_grad(*([null] * 7), 1, 512, 10, 0, 0, *([null] * 5), 0, 0, null)
    assert rc == 1 and b"null pointer" in lib.ttx_last_error()
    rc = lib.ttx_wide_sp(*([null] * 6), 4, 0, 4, 512, 10, 0, 0, *([null] * 6), 512, null, 0, null)
    assert rc == 1 and b"null pointer" in lib.ttx_last_error()
    x = ctypes.c_void_p(256)                                                          # never dereferenced: checks come first
    rc = lib.ttx_wide_pw(x, 512, x, x, x, x, 4, 1, 4, 512, 10, 0, x, 0, null)           # odd tile_lo
    assert rc == 1 and b"tile_lo must be even" in lib.ttx_last_error()
    rc = lib.ttx_wide_dw(x, 256, x, x, x, 4, 0, 4, 512, 10, 0, x, x, 0, null)           # matrix smaller than the range
    assert rc == 1 and b"bad tile range" in lib.ttx_last_error()
    rc = lib.ttx_wide_dw(x, 512, x, x, x, 4, 0, 4, 768, 10, 0, x, x, 0, null)
    assert rc == 1 and b"not supported" in lib.ttx_last_error()


def test_warprnnt_pytorch_surface():
    assert warprnnt_pytorch.RNNTLoss is ttb.RNNTLoss and warprnnt_pytorch.rnnt_loss is ttb.rnnt_loss
    crit = warprnnt_pytorch.RNNTLoss()
    assert crit.blank == 0 and crit.reduction == "mean"
    with pytest.raises(ValueError):
        warprnnt_pytorch.RNNTLoss(reduction="avg")
    acts = torch.zeros(2, 3, 2, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(acts, _i32([[1], [1]]), _i32([3, 2]), _i32([1, 1]))
    with pytest.raises(NotImplementedError):
        ttb.rnnt_loss(acts, _i32([[1], [1]]), _i32([3, 2]), _i32([1, 1]), fastemit_lambda=0.1)


def test_certify_inputs_matches_upstream_conventions():
    acts = torch.zeros(2, 3, 2, 5)
    lab, al, ll = _i32([[1], [1]]), _i32([3, 2]), _i32([1, 1])
    ttb.certify_inputs(acts, lab, al, ll)
    rnnt_oracle.certify_inputs(acts, lab, al, ll)
    for bad, exc in (((acts, lab.long(), al, ll), TypeError), ((acts, lab, al.long(), ll), TypeError),
                     ((acts[0], lab, al, ll), ValueError), ((acts, lab[0], al, ll), ValueError),
                     ((acts, lab, _i32([3]), ll), ValueError), ((acts, lab, _i32([2, 2]), ll), ValueError),
                     ((acts, lab, al, _i32([0, 0])), ValueError)):
        with pytest.raises(exc):
            ttb.certify_inputs(*bad)
        with pytest.raises(exc):
            rnnt_oracle.certify_inputs(*bad)


def test_certify_inputs_cache_follows_in_place_edits_and_label_range():
    """The length check is skipped only for the very same tensors at the same version: an in-place edit of the lengths or
    labels is seen by the next call.  Labels inside label_lens must index the vocabulary; the padding beyond may hold
    anything (tt/dataset.py:46-48 pads with -1)."""
    acts = torch.zeros(2, 3, 3, 5)
    lab, al, ll = _i32([[1, 4], [2, -1]]), _i32([3, 2]), _i32([2, 1])
    sizes = ttb.certify_inputs(acts, lab, al, ll)
    assert sizes == ttb.certify_inputs(acts, lab, al, ll) and sizes[0] == 2          # one 128-row tile per utterance
    al[0] = 2                                                                       # in place: T no longer max(act_lens)
    with pytest.raises(ValueError, match="Input length mismatch"):
        ttb.certify_inputs(acts, lab, al, ll)
    al[0] = 3
    ttb.certify_inputs(acts, lab, al, ll)
    lab[0, 1] = 5                                                                   # == V, inside label_lens[0] = 2
    with pytest.raises(ValueError, match="labels must lie"):
        ttb.certify_inputs(acts, lab, al, ll)
    lab[0, 1] = -1
    with pytest.raises(ValueError, match="labels must lie"):
        ttb.certify_inputs(acts, lab, al, ll)
    lab[0, 1] = 4
    lab[1, 1] = 77                                                                  # beyond label_lens[1] = 1: never read
    ttb.certify_inputs(acts, lab, al, ll)


def _load_sd(module, g):
    module.load_state_dict({k[3:]: torch.tensor(g[k]) for k in g.files if k.startswith("sd_")})


def test_jointnet_dense_fallback_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "tt_joint.npz"))
    m = ttb.JointNet(48, 40, 23)
    assert list(m.state_dict()) == ["forward_layer.weight", "forward_layer.bias", "project_layer.weight",
                                    "project_layer.bias"]
    _load_sd(m, g)
    out = m(torch.tensor(g["enc"]), torch.tensor(g["dec"]))
    assert type(out) is torch.Tensor
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    out1 = m(torch.tensor(g["enc"])[1, 2].view(-1), torch.tensor(g["dec"])[1, 3].view(-1))  # decode call, tt/model.py:77
    np.testing.assert_allclose(out1.detach().numpy(), g["logits_1d"], rtol=1e-5, atol=1e-6)


def test_jointnetwork_dense_fallback_matches_reference_fixture(golden_dir):
    g = np.load(os.path.join(golden_dir, "espnet_joint.npz"))
    m = ttb.JointNetwork(19, 24, 20, 32, "tanh")
    assert list(m.state_dict()) == ["lin_enc.weight", "lin_enc.bias", "lin_dec.weight", "lin_out.weight",
                                    "lin_out.bias"]
    _load_sd(m, g)
    out = m(torch.tensor(g["h_enc"]), torch.tensor(g["h_dec"]))
    np.testing.assert_allclose(out.detach().numpy(), g["logits"], rtol=1e-5, atol=1e-6)
    for act in ("relu", "hardtanh", "selu", "swish"):
        ttb.JointNetwork(19, 24, 20, 32, act)(torch.tensor(g["h_enc"]), torch.tensor(g["h_dec"]))


def test_lazy_handle_looks_like_the_logits():
    torch.manual_seed(0)
    ep, pp = torch.randn(2, 5, 16, requires_grad=True), torch.randn(2, 3, 16, requires_grad=True)
    lin = torch.nn.Linear(16, 7)
    h = ttb.LazyJointLogits(ep, pp, lin.weight, lin.bias)
    assert tuple(h.shape) == (2, 5, 3, 7) and h.dim() == 4 and h.size(3) == 7 and h.dtype == torch.float32
    assert h.is_contiguous() and h.contiguous() is h and not h.is_cuda
    assert h.to(dtype=torch.float32) is h and h.float() is h
    hb = h.to(dtype=torch.bfloat16)                       # transducer/loss.py:57-60 style cast keeps it lazy
    assert isinstance(hb, ttb.LazyJointLogits) and hb.dtype == torch.bfloat16 and hb.parts[0] is ep
    dense = torch.nn.functional.linear(torch.tanh(ep[:, :, None] + pp[:, None]), lin.weight, lin.bias)
    assert torch.allclose(h.materialize(), dense)
    assert torch.allclose(h + 1, dense + 1) and torch.equal(h.argmax(-1), dense.argmax(-1))
    os.environ["TTX_MATERIALIZE_LIMIT_GB"] = "0.0000001"
    try:
        with pytest.raises(RuntimeError, match="refusing to materialise"):
            h.materialize()
    finally:
        del os.environ["TTX_MATERIALIZE_LIMIT_GB"]
    h.materialize().sum().backward()                      # materialisation is differentiable w.r.t. the parts
    assert ep.grad is not None and lin.weight.grad is not None
    # transducer/loss.py:61-62 ("warp-rnnt" branch): log_softmax over the vocabulary keeps the handle lazy
    for lp in (torch.log_softmax(h, dim=-1), torch.nn.functional.log_softmax(h, -1), h.log_softmax(3)):
        assert isinstance(lp, ttb.LazyJointLogits) and lp.normalised and lp.parts[0] is ep
        assert torch.allclose(lp.materialize(), torch.log_softmax(dense, -1), atol=1e-6)
    assert lp.detach().normalised and lp.to(dtype=torch.bfloat16).normalised and not h.normalised
    other = torch.log_softmax(h, dim=1)                   # any other dimension is an ordinary (dense) tensor operation
    assert not isinstance(other, ttb.LazyJointLogits) and torch.allclose(other, torch.log_softmax(dense, 1), atol=1e-6)


def test_warp_rnnt_surface():
    """espnet's TransLoss("warp-rnnt") imports `from warp_rnnt import rnnt_loss` (transducer/loss.py:29) and calls it with
    reduction="mean", blank=..., gather=True (:64-72)."""
    import inspect
    import warp_rnnt
    params = list(inspect.signature(warp_rnnt.rnnt_loss).parameters)
    assert params[:4] == ["log_probs", "labels", "frames_lengths", "labels_lengths"]
    assert {"average_frames", "reduction", "blank", "gather", "fastemit_lambda"} <= set(params)
    with pytest.raises(ValueError):
        warp_rnnt.rnnt_loss(torch.zeros(1, 1, 1, 2), torch.zeros(1, 0, dtype=torch.int32), torch.ones(1, dtype=torch.int32),
                            torch.zeros(1, dtype=torch.int32), reduction="average")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        warp_rnnt.rnnt_loss(torch.zeros(1, 1, 1, 2), torch.zeros(1, 0, dtype=torch.int32), torch.ones(1, dtype=torch.int32),
                            torch.zeros(1, dtype=torch.int32))


def test_algorithm_model_with_rounding_meets_tolerances_against_oracle():
    """The kernel decomposition + fp16 operand rounding, modelled on CPU, vs the oracle: <=1e-4 loss, <=1e-3 grads."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import algo_model
    torch.manual_seed(0)
    B, T, U, V, H = 3, 14, 5, 300, 128
    E = (torch.randn(B, T, H) * 0.6).requires_grad_()
    P = (torch.randn(B, U + 1, H) * 0.6).requires_grad_()
    lo = torch.nn.Linear(H, V)
    W, b = lo.weight.detach().clone().requires_grad_(), lo.bias.detach().clone().requires_grad_()
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    al, ll = _i32([T, T - 6, 1]), _i32([U, U - 2, 0])
    labels[1, U - 2:] = -1
    labels[2, :] = -1
    gc = torch.tensor([0.5, 1.0, 0.25])
    z = torch.tanh(E[:, :, None] + P[:, None]) @ W.T + b
    c = rnnt_oracle.rnnt_loss(z, labels, al, ll, 0, "none")
    (c * gc).sum().backward()
    rel = lambda a, r: float((a - r.double()).norm() / r.double().norm())  # noqa: E731
    for emulate, tl, tg in ((False, 1e-6, 5e-5), (True, 1e-4, 1e-3)):
        m = algo_model.forward_backward(E.detach(), P.detach(), W.detach(), b.detach(), labels, al, ll, gc,
                                        emulate=emulate)
        assert float(((m["costs"] - c.detach()) / c.detach()).abs().max()) < tl
        for k, ref in (("dEproj", E.grad), ("dPproj", P.grad), ("dW", W.grad), ("db", b.grad)):
            assert rel(m[k], ref) < tg, (emulate, k, rel(m[k], ref))


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_install_rebinds_reference_classes_and_model_builds_unchanged():
    tt_model = ref_import.tt_model()
    ref_import.espnet_joint_module()
    orig_tt, orig_es = tt_model.JointNet, sys.modules["espnet.nets.pytorch_backend.transducer.joint_network"].JointNetwork
    try:
        done = ttb.install()
        assert "tt.model.JointNet" in done and tt_model.JointNet is ttb.JointNet
        import yaml
        from tt.utils import AttrDict
        cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "aishell.yaml"))))
        cfg.model.enc.n_layer = 1
        cfg.model.dec.n_layer = 1
        cfg.model.vocab_size = 37
        model = tt_model.Transducer(cfg.model)            # unmodified reference assembly, our joint inside
        assert isinstance(model.joint, ttb.JointNet)
        logits = model(torch.randn(2, 12, 512), torch.randint(1, 37, (2, 4)))
        assert logits.shape == (2, 12, 5, 37)
        import tt_espnet.model as tem
        assert tem.JointNetwork is orig_es or tem.JointNetwork is ttb.JointNetwork
    finally:
        ttb.uninstall()
        tt_model.JointNet = orig_tt
        sys.modules["espnet.nets.pytorch_backend.transducer.joint_network"].JointNetwork = orig_es


@pytest.mark.skipif(not ref_import.available(), reason="reference tree only exists in the build container")
def test_unmodified_train_loop_runs_with_the_drop_ins():
    """`import train` resolves `from warprnnt_pytorch import RNNTLoss` (train.py:13) to this repo's package, and the
    UNMODIFIED train.train() (train.py:22-91) steps a Transducer built with our JointNet.  No GPU here, so the
    criterion handed to train() is the oracle's CPU RNNTLoss (the product loss is CUDA-only and is checked
    against the same oracle in tests/test_gpu_parity.py)."""
    import logging
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    orig = tt_model.JointNet
    try:
        ttb.install(patch_espnet=False)
        import train as ref_train
        assert ref_train.RNNTLoss is ttb.RNNTLoss
        import yaml
        from tt.optim import Optimizer
        from tt.utils import AttrDict
        cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "aishell.yaml"))))
        cfg.model.enc.n_layer = 1
        cfg.model.dec.n_layer = 1
        cfg.model.vocab_size = 31
        cfg.training.num_gpu = 0
        cfg.training.show_interval = 1
        torch.manual_seed(0)
        model = tt_model.Transducer(cfg.model)
        assert isinstance(model.joint, ttb.JointNet)
        opt = Optimizer(model.parameters(), cfg.optim)
        data = [(torch.randn(2, 20, 512), torch.tensor([20, 16]), torch.randint(1, 31, (2, 5)), torch.tensor([5, 3]))
                for _ in range(3)]
        log = logging.getLogger("tt-test")
        records = []
        handler = logging.Handler()
        handler.emit = lambda r: records.append(r.getMessage())
        log.addHandler(handler)
        log.setLevel(logging.INFO)
        ref_train.train(0, cfg, model, data, opt, rnnt_oracle.RNNTLoss(), log)
        losses = [float(m.split(", Loss:")[1].split(",")[0]) for m in records if "Global Step" in m]
        assert len(losses) == 3 and losses[-1] < losses[0]
    finally:
        ttb.uninstall()
        tt_model.JointNet = orig


def test_length_bucket_sampler_covers_every_index_once_and_balances_ranks():
    rng = np.random.RandomState(0)
    T = rng.randint(50, 1000, size=1003)
    U = np.clip(T // 5 + rng.randint(-8, 9, size=1003), 5, 200)          # label count follows duration, as in speech
    cells = T * (U + 1)
    world, bs = 4, 8
    per_rank = []
    for rank in range(world):
        s = ttb.LengthBucketSampler(cells, bs, world=world, rank=rank, bucket=4, seed=3)
        s.set_epoch(2)
        batches = list(s)
        assert len(batches) == len(s) == -(-1003 // (world * bs)) and all(len(b) == bs for b in batches)
        per_rank.append(batches)
    seen = np.concatenate([np.asarray(b) for r in per_rank for b in r])
    assert set(seen.tolist()) == set(range(1003)) and len(seen) == len(per_rank[0]) * world * bs      # + 21 repeats of the shortest
    # the ranks of one step work on lattices of similar size: the padded size B * max T * max (U + 1) of the slowest rank,
    # summed over the steps, against random batching of the same corpus
    def padded(idx):
        return len(idx) * T[idx].max() * (U[idx].max() + 1)
    bucketed = sum(max(padded(np.asarray(per_rank[r][i])) for r in range(world)) for i in range(len(per_rank[0])))
    perm = rng.permutation(1003)[: len(per_rank[0]) * world * bs - 21]
    perm = np.concatenate([perm, perm[:21]]).reshape(len(per_rank[0]), world, bs)
    random_batches = sum(max(padded(perm[i, r]) for r in range(world)) for i in range(len(per_rank[0])))
    assert bucketed < 0.6 * random_batches
    # another epoch, another order; drop_last drops the tail instead of padding it
    s0 = ttb.LengthBucketSampler(cells, bs, world=world, rank=0, bucket=4, seed=3)
    assert list(s0) != per_rank[0]
    sd = ttb.LengthBucketSampler(cells, bs, world=world, rank=1, drop_last=True)
    assert len(list(sd)) == len(sd) == 1003 // (world * bs)


def test_mask_augment_draws_like_the_reference_and_crop_matches_train_loop():
    """Same random-number consumption and the same masked bins as tt/utils.py:297-329 (CPU tensors take the slice
    assignments; the CUDA launch is compared in tests/test_gpu_callers.py); crop_to_max = train.py:32-35."""
    import random
    x = torch.randn(3, 50, 40)

    def ref_time(inputs, max_mask_time=5, mask_num=10):          # restated from tt/utils.py:297-312 for the CPU-only check
        for _ in range(mask_num):
            t = int(np.random.uniform(low=0.0, high=max_mask_time))
            t0 = random.randint(0, inputs.shape[1] - t)
            inputs[:, t0:t0 + t, :] = 0
        return inputs

    def ref_freq(inputs, max_mask_frequency=5, mask_num=10):
        for _ in range(mask_num):
            f = int(np.random.uniform(low=0.0, high=max_mask_frequency))
            f0 = random.randint(0, inputs.shape[2] - f)
            inputs[:, :, f0:f0 + f] = 0
        return inputs

    np.random.seed(5); random.seed(6)
    want = ref_time(ref_freq(x.clone(), 5, 10), 5, 10)
    state = (np.random.get_state()[1][:4].tolist(), random.random())
    np.random.seed(5); random.seed(6)
    got = ttb.time_mask_augment(ttb.frequency_mask_augment(x.clone(), max_mask_frequency=5, mask_num=10), max_mask_time=5,
                                mask_num=10)
    assert torch.equal(got, want) and (np.random.get_state()[1][:4].tolist(), random.random()) == state
    np.random.seed(5); random.seed(6)
    assert torch.equal(ttb.mask_augment(x.clone()), want)
    assert 0 < int((want == 0).sum()) < want.numel()
    inputs, il = torch.randn(2, 30, 8), torch.tensor([22, 17])
    targets, tl = torch.randint(1, 9, (2, 12)), torch.tensor([5, 9])
    a, _, b, _ = ttb.crop_to_max(inputs, il, targets, tl)
    assert a.shape == (2, 22, 8) and b.shape == (2, 9)
    lib = _lib.get()
    bad = (ctypes.c_int32 * 3)(1, 48, 5)                                       # time mask running past T = 50
    rc = lib.ttx_spec_mask(ctypes.c_void_p(256), 3, 50, 40, 2000, 40, bad, 1, 0, None)
    assert rc == 1 and b"outside" in lib.ttx_last_error()
