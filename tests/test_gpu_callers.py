"""GPU gate (-m gpu): the reference's OWN callers drive the product on CUDA.

  * tt path     -- unmodified `train.train()` (train.py:22-91) steps an unmodified `tt.model.Transducer`
                   (tt/model.py:42-68) whose joint is our drop-in (`install()`), with `warprnnt_pytorch.RNNTLoss`
                   (this repository's package, train.py:13) as the criterion, on cuda:0.
  * espnet path -- unmodified `tt_espnet.model.TransformerTransducer.forward` (tt_espnet/model.py:35-81) with the
                   reference's own `TransLoss` wrapper (transducer/loss.py:43-77: fp32 up-cast of the joint output,
                   loss cast back) around our joint + loss, fp32 and bf16.

The arbiter is the SAME reference model with the reference's own joint, run on the CPU with the oracle loss
(fp32), or -- bf16, where CPU and GPU encoders round differently -- run on the GPU with the reference's dense
joint and the dense-logits entry of the product loss (itself pinned to the oracle in test_gpu_parity.py).

The reference sources come from baseline/_ref/ (staged by __graft_entry__.build(), git-ignored, travels with the
snapshot).  Nothing here reads /root/reference on the GPU box.
"""
import copy
import logging
import os
import random
import sys

import numpy as np
import pytest
import torch

import transformer_transducer_b200 as ttb
from oracle import ref_import, rnnt_oracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref (staged reference) is missing")]
DEV = "cuda:0"
LOSS_TOL, GRAD_TOL = 1e-4, 1e-3


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _assert_grads_close(model, ref_model, tol):
    """Relative L2 per parameter; a gradient that is mathematically zero (the key bias of softmax attention: pure
    rounding noise on both sides, ~1e-11 here) is compared against the largest gradient of the model instead."""
    pairs = [(n, a.grad, b.grad) for (n, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters())]
    gmax = max(float(b.detach().double().norm()) for _, _, b in pairs if b is not None)
    for n, a, b in pairs:
        if b is None:
            assert a is None or float(a.abs().max()) == 0, n
            continue
        assert a is not None, n
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        assert float((a - b).norm()) <= tol * float(b.norm()) + 1e-7 * gmax, (n, rel(a, b))


def _seed(s):
    torch.manual_seed(s)
    np.random.seed(s)
    random.seed(s)


def _tt_config(inner, vocab):
    import yaml
    from tt.utils import AttrDict
    cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "aishell.yaml"))))
    cfg.model.enc.n_layer = 1
    cfg.model.dec.n_layer = 1
    cfg.model.vocab_size = vocab
    cfg.model.joint.inner_size = inner
    cfg.model.dropout = 0.0                      # CPU and CUDA dropout streams differ
    cfg.training.show_interval = 1
    return cfg


def _capture_logger(name):
    log = logging.getLogger(name)
    records = []
    handler = logging.Handler()
    handler.emit = lambda r: records.append(r.getMessage())
    log.handlers = [handler]
    log.setLevel(logging.INFO)
    return log, records


def _losses(records):
    return [float(m.split(", Loss:")[1].split(",")[0]) for m in records if "Global Step" in m]


@pytest.mark.parametrize("inner", [512, 1024])          # fused tcgen05 width / aishell.yaml's own joint width
def test_unmodified_train_loop_drives_product_on_cuda(inner):
    """train.py:22-91 unchanged, model on cuda:0 with our JointNet + our RNNTLoss, vs the reference joint + oracle
    loss on the CPU: same per-step losses (SGD steps in between, so later losses also check the gradients) and the
    same parameters after three optimizer steps; a separate single step compares every parameter gradient."""
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    import train as ref_train
    from tt.optim import Optimizer
    assert ref_train.RNNTLoss is ttb.RNNTLoss          # `from warprnnt_pytorch import RNNTLoss` resolved to the product
    V = 211
    cfg = _tt_config(inner, V)
    _seed(0)
    data = [(torch.randn(3, 36, 512), torch.tensor([36, 30, 17]), torch.randint(1, V, (3, 7)), torch.tensor([7, 5, 2]))
            for _ in range(3)]
    orig = tt_model.JointNet
    try:
        _seed(1)
        ref_model = tt_model.Transducer(cfg.model)       # reference joint (dense logits)
        ttb.install(patch_espnet=False)
        model = tt_model.Transducer(cfg.model)           # same assembly, our joint inside
        assert isinstance(model.joint, ttb.JointNet) and not isinstance(ref_model.joint, ttb.JointNet)
        model.load_state_dict(ref_model.state_dict())    # state-dict keys are the reference's
        model = model.to(DEV)
        init = {n: p_.detach().clone() for n, p_ in ref_model.named_parameters()}

        # (1) one step by hand: every parameter gradient
        inputs, ilen, targets, tlen = data[0]
        ref_model.train()
        model.train()
        want = rnnt_oracle.RNNTLoss()(ref_model(inputs, targets), targets.int(), ilen.int(), tlen.int())
        want.backward()
        logits = model(inputs.to(DEV), targets.to(DEV))
        assert isinstance(logits, ttb.LazyJointLogits)
        got = ref_train.RNNTLoss()(logits, targets.int().to(DEV), ilen.int().to(DEV), tlen.int().to(DEV))
        got.backward()
        assert abs(float(got.detach()) - float(want.detach())) / abs(float(want.detach())) < LOSS_TOL
        for (n, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
            assert b.grad is not None and a.grad is not None, n
            assert rel(a.grad, b.grad) < GRAD_TOL, (n, rel(a.grad, b.grad))
        ref_model.zero_grad()
        model.zero_grad()

        # (2) the unmodified loop
        cfg.training.num_gpu = 0
        log_r, rec_r = _capture_logger("tt-ref")
        _seed(2)                                          # time / frequency masks draw from numpy + random
        ref_train.train(0, cfg, ref_model, copy.deepcopy(data), Optimizer(ref_model.parameters(), cfg.optim),
                        rnnt_oracle.RNNTLoss(), log_r)
        cfg.training.num_gpu = 1
        log_g, rec_g = _capture_logger("tt-gpu")
        _seed(2)
        ref_train.train(0, cfg, model, copy.deepcopy(data), Optimizer(model.parameters(), cfg.optim),
                        ref_train.RNNTLoss(), log_g)
        lr_, lg_ = _losses(rec_r), _losses(rec_g)
        assert len(lr_) == 3 and len(lg_) == 3
        assert abs(lg_[0] - lr_[0]) / abs(lr_[0]) < LOSS_TOL, (lg_, lr_)
        for a, b in zip(lg_[1:], lr_[1:]):                   # after SGD steps taken with the compared gradients
            assert abs(a - b) / abs(b) < 1e-3, (lg_, lr_)
        for (n, a), (_, b) in zip(model.named_parameters(), ref_model.named_parameters()):
            # the three (momentum, norm-clipped) SGD updates themselves, not the parameters they are small against; a
            # clipped update is a few float32 ulps of an O(1) parameter, so this is a coarse check -- the gradients were
            # compared to 1e-3 above and the loss trajectory to 1e-3 here
            assert rel(a.detach().cpu() - init[n], b.detach() - init[n]) < 5e-2, n
    finally:
        ttb.uninstall()
        tt_model.JointNet = orig


def _espnet_config(vocab):
    import yaml
    from tt.utils import AttrDict
    cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "espnet_aishell.yaml"))))
    m = cfg.model
    for part in (m.enc, m.dec):                  # item assignment: the model is built from **config.enc (dict items)
        part["dropout_rate"] = 0.0               # CPU and CUDA dropout streams differ
        part["positional_dropout_rate"] = 0.0
        part["attention_dropout_rate"] = 0.0
    m.dec["input_size"] = vocab
    m.joint["vocab_size"] = vocab
    return m


def _espnet_models(vocab):
    ref_import.prepare(stub_train_deps=True)
    jn = ref_import.espnet_joint_module()
    import tt_espnet.model as tem
    orig = (jn.JointNetwork, tem.JointNetwork)
    cfg = _espnet_config(vocab)
    _seed(3)
    ref_model = tem.TransformerTransducer(cfg)           # reference JointNetwork inside
    ttb.install(patch_tt=False)
    try:
        model = tem.TransformerTransducer(cfg)
    finally:
        ttb.uninstall()
        jn.JointNetwork, tem.JointNetwork = orig
    assert isinstance(model.joint, ttb.JointNetwork) and not isinstance(ref_model.joint, ttb.JointNetwork)
    model.load_state_dict(ref_model.state_dict())
    return ref_model, model


def _espnet_batch(vocab):
    _seed(4)
    speech = torch.randn(3, 40, 512)
    slen = torch.tensor([40, 33, 12])
    text = torch.randint(1, vocab - 1, (3, 8))
    tlen = torch.tensor([8, 6, 1])
    for i, n in enumerate(tlen):
        text[i, int(n):] = -1                             # tt/dataset.py:46-48 pads with ignore_id = -1
    return speech, slen, text, tlen


def test_espnet_transformer_transducer_forward_backward_fp32():
    """tt_espnet/model.py:35-81 + transducer/loss.py:43-77 unchanged around our joint and loss (fp32), vs the same
    model with the reference joint and the oracle loss on the CPU."""
    V = 333
    ref_model, model = _espnet_models(V)
    speech, slen, text, tlen = _espnet_batch(V)
    ref_model.loss.trans_loss = rnnt_oracle.RNNTLoss(blank=0)      # the CPU arbiter behind the reference's wrapper
    ref_model.train()
    want = ref_model(speech, slen, text, tlen)
    want.backward()
    model = model.to(DEV).train()
    assert isinstance(model.loss.trans_loss, ttb.RNNTLoss)         # TransLoss found `warprnnt_pytorch` = the product
    got = model(speech.to(DEV), slen.to(DEV), text.to(DEV), tlen.to(DEV))
    assert got.dtype == torch.float32 and got.shape == (1,)
    got.backward()
    assert abs(float(got.detach()) - float(want.detach())) / abs(float(want.detach())) < LOSS_TOL
    _assert_grads_close(model, ref_model, GRAD_TOL)


def test_espnet_transformer_transducer_forward_backward_bf16_joint():
    """configs[4]'s "bf16 joint inputs via tt_espnet/model.py": the reference's encoder cannot run in bf16 at all
    (attention.py:81 asks numpy for finfo of the score dtype), so the model stays fp32 and only the joint is bf16 --
    a forward pre-hook casts its inputs, everything else (TransformerTransducer.forward, TransLoss with its fp32
    up-cast of the joint output and the cast of the loss back to bf16, transducer/loss.py:57-60,75) is unmodified.
    Fused path vs the SAME model on the GPU with the reference's dense joint + the dense-logits entry of the product
    loss.  Tolerances are the bf16-variant ones (the reference rounds its logits to bf16, the fused path does not):
    loss 1e-2, gradients 5e-2 relative L2."""
    V = 333
    ref_model, model = _espnet_models(V)
    speech, slen, text, tlen = _espnet_batch(V)

    def to_bf16(_module, args, kwargs):
        return args, {k: v.bfloat16() for k, v in kwargs.items()}

    outs = []
    for m in (ref_model, model):
        m = m.to(DEV).train()
        m.joint.bfloat16()
        m.joint.register_forward_pre_hook(to_bf16, with_kwargs=True)
        loss = m(speech.to(DEV), slen.to(DEV), text.to(DEV), tlen.to(DEV))
        assert loss.dtype == torch.bfloat16 and loss.shape == (1,)     # transducer/loss.py:75 casts the loss back
        loss.backward()
        outs.append(loss)
    want, got = outs
    assert abs(float(got.detach()) - float(want.detach())) / abs(float(want.detach())) < 1e-2
    _assert_grads_close(model, ref_model, 5e-2)


# ----------------------------------------------------------------------------- greedy search (decode-time joint)
def test_espnet_transloss_warp_rnnt_branch_stays_on_the_fused_path():
    """espnet/nets/pytorch_backend/transducer/loss.py:27-31,61-72: TransLoss("warp-rnnt") takes log_softmax of the joint's
    output and calls warp_rnnt.rnnt_loss(log_probs, ..., reduction="mean", gather=True).  With our JointNetwork the handle
    stays lazy through the log_softmax; loss and gradients equal the oracle's; a dense tensor of log-probabilities works
    too (costs, reductions, average_frames)."""
    import warp_rnnt
    ref_import.prepare(stub_train_deps=True)
    from espnet.nets.pytorch_backend.transducer.loss import TransLoss
    from oracle import joint_ref
    torch.manual_seed(2)
    B, T, U, V, D, H = 3, 21, 5, 97, 32, 512
    ref = joint_ref.EspnetJointNetwork(V, D, D, H, "tanh")
    mine = ttb.JointNetwork(V, D, D, H, "tanh")
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    enc, pred = torch.randn(B, T, 1, D), torch.randn(B, 1, U + 1, D)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    al, ll = torch.tensor([T, 15, 4], dtype=torch.int32), torch.tensor([U, 2, 0], dtype=torch.int32)
    e0, p0 = enc.clone().requires_grad_(), pred.clone().requires_grad_()
    want_costs = rnnt_oracle.rnnt_loss(ref(e0, p0), labels, al, ll, 0, "none")
    want_costs.mean().backward()
    crit = TransLoss("warp-rnnt", 0)
    assert crit.trans_loss is warp_rnnt.rnnt_loss
    e1, p1 = enc.to(DEV).requires_grad_(), pred.to(DEV).requires_grad_()
    seen = []
    orig = warp_rnnt.rnnt_loss
    crit.trans_loss = lambda lp, *a, **k: (seen.append(type(lp)), orig(lp, *a, **k))[1]
    got = crit(mine(e1, p1), labels.to(DEV), al.to(DEV), ll.to(DEV))
    got.backward()
    assert seen == [ttb.LazyJointLogits]
    assert abs(float(got.detach()) - float(want_costs.mean())) / float(want_costs.mean()) < LOSS_TOL
    assert rel(e1.grad, e0.grad) < GRAD_TOL and rel(p1.grad, p0.grad) < GRAD_TOL
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel(a.grad, b.grad) < GRAD_TOL, n
    # dense log-probabilities, every reduction, average_frames
    lp = torch.log_softmax(ref(enc, pred).detach(), -1).to(DEV)
    args = (labels.to(DEV), al.to(DEV), ll.to(DEV))
    costs = warp_rnnt.rnnt_loss(lp, *args)
    assert costs.shape == (B,) and float(((costs.cpu() - want_costs.detach()) / want_costs.detach()).abs().max()) < 1e-5
    assert abs(float(warp_rnnt.rnnt_loss(lp, *args, reduction="sum")) - float(want_costs.sum())) < 1e-4 * float(want_costs.sum())
    avg = warp_rnnt.rnnt_loss(lp, *args, average_frames=True, reduction="mean")
    assert avg.dim() == 0 and abs(float(avg) - float((want_costs.detach() / al).mean())) < 1e-5 * float(want_costs.mean())


def test_decode_scan_kernel_matches_torch_argmax():
    """ttx_decode_scan through the C ABI: per-frame argmax of tanh(eproj + pvec) . W^T + b, first non-blank frame and
    its label, against float64 torch (frames whose two best logits are closer than 1e-4 are not compared: fp32
    summation order decides those, in the reference too)."""
    import ctypes
    from transformer_transducer_b200 import _lib
    lib = _lib.get()
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    for seed, (n, H, V, blank_boost) in enumerate([(64, 512, 4232, 6.0), (17, 1024, 211, 3.0), (1, 200, 70, 0.0),
                                                   (64, 2048, 6485, 50.0)]):
        g = torch.Generator().manual_seed(seed)
        eproj, pvec = torch.randn(n, H, generator=g), torch.randn(H, generator=g)
        w, b = torch.randn(V, H, generator=g) / H ** 0.5, torch.randn(V, generator=g)
        b[0] += blank_boost
        z = torch.tanh(eproj.double() + pvec.double()) @ w.double().T + b.double()
        top = z.topk(2, dim=1)
        want = top.indices[:, 0]
        clear = (top.values[:, 0] - top.values[:, 1]) > 1e-4
        out = torch.full((2 + 64,), -7, dtype=torch.int32, device=DEV)
        scratch = torch.empty(64, dtype=torch.int64, device=DEV)
        args = [t.to(DEV) for t in (eproj, pvec, w, b)]
        _lib.check(lib.ttx_decode_scan(p(args[0]), H, p(args[1]), p(args[2]), p(args[3]), n, H, V, 0, p(scratch), p(out), 0,
                                       None), "decode_scan")
        got = out.cpu()
        assert torch.equal(got[2:2 + n][clear].long(), want[clear])
        if bool(clear.all()):
            nz = (want != 0).nonzero()
            first = int(nz[0]) if len(nz) else n
            assert int(got[0]) == first and int(got[1]) == (int(want[first]) if first < n else 0)


def _boost_blank(out_layer, amount):
    with torch.no_grad():
        out_layer.bias[0] += amount


@pytest.mark.parametrize("inner", [512, 1024])
def test_tt_greedy_decode_equals_reference_decode(inner):
    """tt/model.py:70-90: the reference's per-frame loop (joint -> softmax -> argmax -> .item()) on cuda:0 vs the
    rebound `Transducer.decode` (install()): identical label sequences, integer-exact, through `recognize()`."""
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    V = 211
    cfg = _tt_config(inner, V)
    _seed(21)
    model = tt_model.Transducer(cfg.model).to(DEV).eval()
    _boost_blank(model.joint.project_layer, 0.7)          # most frames predict the blank, like a trained model
    inputs = torch.randn(3, 90, 512, device=DEV)
    lengths = [90, 71, 33]
    with torch.no_grad():
        want = model.recognize(inputs, lengths)
        try:
            done = ttb.install(patch_espnet=False)
            assert "tt.model.Transducer.decode" in done and "tt.model.Transducer.recognize" in done
            got = model.recognize(inputs, lengths)        # tt/model.py:92-108 rebound: batched search over the utterances
            enc = model.encoder(inputs, None)
            got_single = [model.decode(enc[b], lengths[b]) for b in range(3)]      # tt/model.py:70-90 rebound
        finally:
            ttb.uninstall()
    assert got == want and got_single == want
    assert all(0 < len(w) < n for w, n in zip(want, lengths))    # labels were emitted, and blanks in between
    with pytest.raises(IndexError):                              # like enc_state[t] past the end in the reference's loop
        ttb.greedy_search(model.joint, enc[0], 91, lambda toks: model.decoder(torch.tensor([toks], device=DEV))[:, -1, :])


@pytest.mark.parametrize("beam", [5, 3])
def test_tt_beam_search_equals_reference_beam_search(beam):
    """tt/model.py:110-179: the reference's beam search (per-frame joint, top-k expansion of every hypothesis, its own
    child-table bookkeeping) on cuda:0 vs the rebound `Transducer.beam_search`: identical label sequences, directly and
    through `recognize_beam_search()` (tt/model.py:181-198, beam width 5)."""
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    V = 173
    cfg = _tt_config(512, V)
    _seed(33)
    model = tt_model.Transducer(cfg.model).to(DEV).eval()
    _boost_blank(model.joint.project_layer, 0.35)
    inputs = torch.randn(2, 70, 512, device=DEV)
    lengths = [70, 38]
    with torch.no_grad():
        enc = model.encoder(inputs, None)
        want = [model.beam_search(enc[b], lengths[b], beam_width=beam) for b in range(2)]
        want_rec = model.recognize_beam_search(inputs, lengths) if beam == 5 else None
        try:
            done = ttb.install(patch_espnet=False)
            assert "tt.model.Transducer.beam_search" in done
            got = [model.beam_search(enc[b], lengths[b], beam_width=beam) for b in range(2)]
            got_rec = model.recognize_beam_search(inputs, lengths) if beam == 5 else None
        finally:
            ttb.uninstall()
    assert got == want and got_rec == want_rec
    assert any(len(w) > 1 for w in want), want


def test_streaming_window_search_equals_demo_loop():
    """audio/streamRec_unlimit_dynamic_window.py:113-115,186-211 (the demo's per-frame loop, restated here around the
    reference's own joint / decoder because it lives inside a tkinter class) vs decode.StreamingGreedy fed the same
    windows: same labels, same blank-frame counts at every label, state carried across windows, 40-label history."""
    from transformer_transducer_b200.decode import StreamingGreedy
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    cfg = _tt_config(1024, 211)
    _seed(9)
    model = tt_model.Transducer(cfg.model).to(DEV).eval()
    _boost_blank(model.joint.project_layer, 0.25)
    windows = [torch.randn(n, 512, device=DEV) for n in (37, 1, 64, 90, 5, 130)]
    with torch.no_grad():
        enc = [model.encoder(w[None], None)[0] for w in windows]
        dec_state = model.decoder(torch.tensor([[0]], device=DEV))
        result, blank_frame, want = [], 0, []
        for e in enc:
            events = []
            for t in range(e.shape[0]):
                pred = int(torch.argmax(torch.softmax(model.joint(e[t].view(-1), dec_state.view(-1)), dim=0), dim=0).item())
                if pred != 0:
                    events.append((pred, blank_frame))
                    result.append(pred)
                    dec_state = model.decoder(torch.tensor([result[-40:]], device=DEV))[:, -1, :]
                    blank_frame = 0
                elif result:
                    blank_frame += 1
            want.append(events)
        search = StreamingGreedy(model.joint, model.decoder)
        got = [search.feed(e) for e in enc]
    assert got == want and search.result == result and search.blank_frame == blank_frame
    assert len(result) > 40                          # the history window was exercised


def test_espnet_greedy_decode_equals_reference_decode():
    """tt_espnet/model.py:83-121: same check for TransformerTransducer.decode / recognize."""
    V = 333
    ref_model, _ = _espnet_models(V)
    import tt_espnet.model as tem
    model = ref_model.to(DEV).eval()
    _boost_blank(model.joint.lin_out, 1.5)
    _seed(5)
    speech = torch.randn(2, 60, 512, device=DEV)
    slen = torch.tensor([60, 41], device=DEV)
    want = model.recognize(speech, slen)
    try:
        ttb.install(patch_tt=False)
        assert tem.TransformerTransducer.decode is not tem.TransformerTransducer._ttb_reference_decode
        assert tem.TransformerTransducer.recognize is not tem.TransformerTransducer._ttb_reference_recognize
        got = model.recognize(speech, slen)               # batched search
        with torch.no_grad():
            enc, _, _ = model.encoder(speech, slen, left_mask=model.encoder_left_mask, right_mask=model.encoder_right_mask)
            got_single = [model.decode(enc[b], slen[b]) for b in range(2)]
    finally:
        ttb.uninstall()
    assert got == want and got_single == want
    assert all(len(w) > 0 for w in want)


# ----------------------------------------------------------------------------- input pipeline
def test_mask_augment_launch_is_bit_identical_to_reference_slices():
    """tt/utils.py:297-329 (the reference's own functions, on a CUDA tensor) vs the rebound single-launch versions under the
    same seeds, contiguous and cropped (strided) batches."""
    ref_import.prepare(stub_train_deps=True)
    import tt.utils as tu
    ref_time, ref_freq = tu.time_mask_augment, tu.frequency_mask_augment
    for seed, crop in ((1, 410), (2, 333), (3, 57)):
        _seed(seed)
        full = torch.randn(5, 410, 512, device=DEV)
        _seed(100 + seed)
        want = ref_time(ref_freq(full.clone()[:, :crop, :], max_mask_frequency=5, mask_num=10), max_mask_time=5, mask_num=10)
        try:
            done = ttb.install(patch_tt=False, patch_espnet=False)
            assert any("mask_augment" in d for d in done) and tu.time_mask_augment is ttb.time_mask_augment
            _seed(100 + seed)
            got = tu.time_mask_augment(tu.frequency_mask_augment(full.clone()[:, :crop, :], max_mask_frequency=5, mask_num=10),
                                       max_mask_time=5, mask_num=10)
            _seed(100 + seed)
            fused = ttb.mask_augment(full.clone()[:, :crop, :])
        finally:
            ttb.uninstall()
        assert tu.time_mask_augment is ref_time
        assert torch.equal(got, want) and torch.equal(fused, want)
        assert 0 < int((want == 0).sum()) < want.numel()


# ----------------------------------------------------------------------------- banded streaming attention
@pytest.mark.parametrize("T,max_len,ctx", [(57, 410, (10, 2)), (410, 410, (10, 2)), (45, 30, (10, 2)), (33, 64, (3, 4)),
                                          (5, 16, (10, 2))])
def test_banded_attention_equals_reference_under_context_mask(T, max_len, ctx):
    """tt/transformer.py:106-177 (the reference's own module, dense T x T scores, on cuda:0) vs the rebound forward whose
    attention core runs on the band, under tt/utils.py:242-251's context mask as tt/model.py:60 would pass it ((T, T, 1)):
    output and every gradient (input, qkv_net, o_net, LayerNorm, r_emb, r_w_bias, r_bias), including the
    keys right of the diagonal whose position term _rel_shift wraps around, T > max_len (padded tables), T shorter than the
    context, another context width.  Tolerance 2e-5 relative L2 (fp32 both sides, different summation order)."""
    ref_import.prepare(stub_train_deps=True)
    import tt.transformer as ttr
    import tt.utils as tu
    from transformer_transducer_b200 import attention as att
    _seed(T)
    B, n_head, d_head, d_model = 3, 4, 64, 256
    attn = ttr.RelLearnableMultiHeadAttn(n_head, d_model, d_head, dropout=0.0).to(DEV)
    r_emb = torch.randn(max_len, n_head, d_head, device=DEV).mul_(0.3).requires_grad_()
    r_w_bias = torch.randn(n_head, d_head, device=DEV).mul_(0.3).requires_grad_()
    r_bias = torch.randn(max_len, n_head, device=DEV).mul_(0.3).requires_grad_()
    w = torch.randn(T, B, d_model, device=DEV, requires_grad=True)
    g = torch.randn(T, B, d_model, device=DEV)
    mask = tu.context_mask(torch.empty(1, T, 1, device=DEV), left_context=ctx[0], right_context=ctx[1])[:, :, None]
    leaves = [w, r_emb, r_w_bias, r_bias] + list(attn.parameters())

    def run():
        for t_ in leaves:
            t_.grad = None
        out = attn(w, r_emb, r_w_bias, r_bias, attn_mask=mask)
        out.backward(g)
        return out.detach().clone(), [t_.grad.detach().clone() for t_ in leaves]

    want, want_g = run()
    calls = []
    orig_apply = att.BandAttnCore.apply
    try:
        done = ttb.install(patch_tt=False, patch_espnet=False, patch_decode=False, patch_data=False, streaming_context=ctx)
        assert "tt.transformer.RelLearnableMultiHeadAttn.forward" in done
        att.BandAttnCore.apply = staticmethod(lambda *a: (calls.append(1), orig_apply(*a))[1])
        got, got_g = run()
        # a mask that is not the context band goes to the reference's forward
        other = mask.clone()
        other[0, min(T - 1, 1), 0] = 1 - other[0, min(T - 1, 1), 0]
        n_calls = len(calls)
        attn(w, r_emb, r_w_bias, r_bias, attn_mask=other)
        assert len(calls) == n_calls
    finally:
        att.BandAttnCore.apply = orig_apply
        ttb.uninstall()
    assert calls, "the band kernel did not run"
    assert ttr.RelLearnableMultiHeadAttn.forward is not att.banded_forward
    assert rel(got, want) < 2e-5
    names = ["w", "r_emb", "r_w_bias", "r_bias"] + [n for n, _ in attn.named_parameters()]
    for n, a, b in zip(names, got_g, want_g):
        assert rel(a, b) < 2e-5, (n, rel(a, b))


@pytest.mark.parametrize("T,lens,ctx", [(57, [57, 40, 9], (10, 2)), (130, [130, 130], (10, 2)), (9, [9, 4, 1], (2, 0)),
                                        (6, [6, 2], (10, 2)), (40, [40, 31], (3, 4))])
def test_espnet_banded_attention_equals_reference_under_padding_and_context_mask(T, lens, ctx):
    """espnet/nets/pytorch_backend/transformer/attention.py:212-308 (RelPositionMultiHeadedAttention, the reference's own
    module with dense T x T scores, on cuda:0) vs the rebound forward on the band kernels (mode 1), under the mask
    tt_espnet's encoders build (espnet2/asr/encoder/transformer_encoder.py:205-210: padding mask AND
    ~make_attention_mask, nets_utils.py:268-281): the encoder's (10, 2) band, the label encoder's (2, 0), T shorter than
    the band, queries beyond a short utterance's length (no key left: zeros), another width.  Output and every gradient
    (input, the four projections, linear_pos, pos_bias_u / _v).  Tolerance 2e-5 relative L2."""
    ref_import.prepare(stub_train_deps=True)
    from espnet.nets.pytorch_backend.nets_utils import make_attention_mask, make_pad_mask
    from espnet.nets.pytorch_backend.transformer import attention as eatt
    from espnet.nets.pytorch_backend.transformer.embedding import RelPositionalEncoding
    from transformer_transducer_b200 import attention as att
    _seed(T + len(lens))
    B, n_head, d_model = len(lens), 4, 256
    attn = eatt.RelPositionMultiHeadedAttention(n_head, d_model, 0.0).to(DEV)
    posenc = RelPositionalEncoding(d_model, 0.0).to(DEV)
    x = torch.randn(B, T, d_model, device=DEV, requires_grad=True)
    g = torch.randn(B, T, d_model, device=DEV)
    _, pos_emb = posenc(x.detach())
    ilens = torch.tensor(lens, device=DEV)
    mask = (~make_pad_mask(ilens)[:, None, :]).to(DEV) & ~make_attention_mask(x, ctx[0], ctx[1])[None, :, :]
    leaves = [x] + list(attn.parameters())

    def run(m):
        for t_ in leaves:
            t_.grad = None
        out = attn(x, x, x, pos_emb, m)
        out.backward(g)
        return out.detach().clone(), [t_.grad.detach().clone() for t_ in leaves]

    want, want_g = run(mask)
    calls = []
    orig_apply = att.BandAttnCore.apply
    try:
        done = ttb.install(patch_tt=False, patch_espnet=False, patch_decode=False, patch_data=False)
        assert "espnet...attention.RelPositionMultiHeadedAttention.forward" in done
        att.BandAttnCore.apply = staticmethod(lambda *a: (calls.append(1), orig_apply(*a))[1])
        got, got_g = run(mask)
        other = mask.clone()                                  # not a band: the reference's own forward
        other[0, 0, 0] = ~other[0, 0, 0]
        n_calls = len(calls)
        attn(x, x, x, pos_emb, other)
        assert len(calls) == n_calls
    finally:
        att.BandAttnCore.apply = orig_apply
        ttb.uninstall()
    assert calls, "the band kernel did not run"
    assert eatt.RelPositionMultiHeadedAttention.forward is not att.espnet_banded_forward
    assert rel(got, want) < 2e-5
    names = ["x"] + [n for n, _ in attn.named_parameters()]
    for n, a, b in zip(names, got_g, want_g):
        if n == "linear_k.bias":          # shifts every score of a query alike: the gradient is zero, both sides hold rounding
            assert float(a.abs().max()) < 1e-5 and float(b.abs().max()) < 1e-5
            continue
        assert rel(a, b) < 2e-5, (n, rel(a, b))
