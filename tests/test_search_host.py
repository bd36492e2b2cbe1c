"""Host-side bookkeeping of decode.beam_search against the reference's own beam search (tt/model.py:110-179) on the CPU.

The frame scanner (the CUDA part) is replaced by a plain-torch stand-in with the same interface, so what is compared is
the search logic: leader selection, skipped blank frames, top-k expansion with the blank removed, the reference's
child table (appended to at every label frame, filled column-wise the first time) and the survivor selection.  The GPU
run of the same comparison, with the real scanner, is tests/test_gpu_callers.py.
"""
import contextlib
import os

import pytest
import torch
import yaml

import transformer_transducer_b200  # noqa: F401  (registers the package under its importable name)
from oracle import ref_import
from transformer_transducer_b200 import decode as D

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref (staged reference) is missing")


class _TorchScanner:
    def __init__(self, joint, enc_state, length):
        self.joint, self.enc, self.length = joint, enc_state, length

    def decoder_halves(self, dec_outs):
        return dec_outs

    def posteriors(self, t, pvecs):
        return torch.stack([torch.softmax(self.joint(self.enc[t].view(-1), pv), dim=0) for pv in pvecs])

    def next_label(self, t, pvec, blank):
        while t < self.length:
            label = int(self.posteriors(t, pvec[None])[0].argmax())
            if label != blank:
                return t, label
            t += 1
        return self.length, blank


@pytest.mark.parametrize("seed,boost", [(2, 0.5), (1, 0.9), (3, 1.3), (4, 0.0)])
def test_beam_search_bookkeeping_equals_reference(monkeypatch, seed, boost):
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    from tt.utils import AttrDict
    cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "aishell.yaml"))))
    cfg.model.enc.n_layer = 1
    cfg.model.dec.n_layer = 1
    cfg.model.vocab_size = 97
    cfg.model.joint.inner_size = 128
    cfg.model.dropout = 0.0
    torch.manual_seed(seed)
    model = tt_model.Transducer(cfg.model).eval()
    monkeypatch.setattr(D, "_FrameScanner", _TorchScanner)
    monkeypatch.setattr(torch.cuda, "device", lambda d: contextlib.nullcontext())
    with torch.no_grad():
        model.joint.project_layer.bias[0] += boost
        enc = model.encoder(torch.randn(1, 40, 512), None)[0]
        step = lambda hyps: model.decoder(torch.tensor(hyps))[:, -1, :]  # noqa: E731
        for width in (5, 3, 2):
            want = model.beam_search(enc, 40, beam_width=width)
            got = D.beam_search(model.joint, enc, 40, step, beam_width=width)
            assert got == want


def test_install_is_idempotent_uninstall_restores_and_cpu_calls_take_the_reference_path():
    """install() twice then uninstall(): every rebound name is the reference's again; while installed, CPU tensors go
    through the reference's own decode / recognize / beam_search / attention forward (the product has no CPU kernels)."""
    import transformer_transducer_b200 as ttb
    ref_import.prepare(stub_train_deps=True)
    tt_model = ref_import.tt_model()
    import tt.transformer as ttr
    import tt.utils as tu
    from espnet.nets.pytorch_backend.transformer import attention as eatt
    from tt.utils import AttrDict
    names = [(tt_model, "JointNet"), (tt_model.Transducer, "decode"), (tt_model.Transducer, "recognize"),
             (tt_model.Transducer, "beam_search"), (ttr.RelLearnableMultiHeadAttn, "forward"),
             (eatt.RelPositionMultiHeadedAttention, "forward"), (tu, "time_mask_augment"), (tu, "frequency_mask_augment")]
    before = [getattr(o, a) for o, a in names]
    cfg = AttrDict(yaml.safe_load(open(os.path.join(ref_import.REF_ROOT, "config", "aishell.yaml"))))
    cfg.model.enc.n_layer = 1
    cfg.model.dec.n_layer = 1
    cfg.model.vocab_size = 53
    cfg.model.joint.inner_size = 64
    cfg.model.dropout = 0.0
    torch.manual_seed(11)
    model = tt_model.Transducer(cfg.model).eval()
    x = torch.randn(2, 30, 512)
    with torch.no_grad():
        model.joint.project_layer.bias[0] += 0.4
        want = model.recognize(x, [30, 17])
        want_beam = model.beam_search(model.encoder(x, None)[0], 30, beam_width=3)
        want_enc = model.encoder(x, tu.context_mask(x)[:, :, None])
    try:
        first = ttb.install()
        second = ttb.install()
        assert "tt.model.Transducer.recognize" in first and not [n for n in second if "Transducer" in n]
        assert all(getattr(o, a) is not b for (o, a), b in zip(names, before))
        with torch.no_grad():
            assert model.recognize(x, [30, 17]) == want
            assert model.beam_search(model.encoder(x, None)[0], 30, beam_width=3) == want_beam
            assert torch.equal(model.encoder(x, tu.context_mask(x)[:, :, None]), want_enc)
    finally:
        ttb.uninstall()
    assert all(getattr(o, a) is b for (o, a), b in zip(names, before))
