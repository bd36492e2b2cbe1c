"""Rebinds the reference's joint classes to the B200 drop-ins without touching reference files.

``Transducer.__init__`` looks up the module global ``JointNet`` at call time (/root/reference/tt/model.py:48)
and ``tt_espnet/model.py:14`` copies ``JointNetwork`` at import, so patching the defining modules (and
``tt_espnet.model`` if it is already imported) before the model is constructed is enough.  The loss
needs no patch: ``from warprnnt_pytorch import RNNTLoss`` (train.py:13) finds the ``warprnnt_pytorch``
package of this repository once the repository root is on ``sys.path``.  The greedy-search methods
(``Transducer.decode``, ``TransformerTransducer.decode``) are rebound to decode.py's versions, the mask augmentation
(``tt.utils.time_mask_augment`` / ``frequency_mask_augment``) to data.py's single-launch versions, and
``tt.transformer.RelLearnableMultiHeadAttn.forward`` to attention.py's banded version (it only takes calls whose mask
is the streaming context mask; everything else runs the reference's forward).
"""
import importlib
import sys

from . import attention as _attention
from . import data as _data
from . import decode as _decode
from .joint import JointNet, JointNetwork


_ORIGINALS = []          # (object, attribute, original value) in patch order, for uninstall()


def _set(obj, attr, value):
    _ORIGINALS.append((obj, attr, getattr(obj, attr)))
    setattr(obj, attr, value)


def uninstall():
    """Undo every rebinding install() made (tests; A/B runs against the reference's own classes)."""
    while _ORIGINALS:
        obj, attr, value = _ORIGINALS.pop()
        setattr(obj, attr, value)


def _patch_decode(cls, fn, name, done, attr="decode"):
    """Greedy / beam search: same signature, the per-frame joint loop replaced by the GPU scan (decode.py); the
    reference's method stays reachable (CPU tensors fall back to it)."""
    if getattr(cls, attr) is fn:
        return
    setattr(cls, "_ttb_reference_" + attr, getattr(cls, attr))
    _set(cls, attr, fn)
    done.append(name)


def install(patch_tt=True, patch_espnet=True, patch_decode=True, patch_data=True, patch_attention=True,
            streaming_context=(10, 2)):
    done = []
    if patch_attention:
        # tt/transformer.py:106-177: the attention core on a band when the mask is tt.utils.context_mask(left, right)
        try:
            m = importlib.import_module("tt.transformer")
            cls = m.RelLearnableMultiHeadAttn
            if cls.forward is not _attention.banded_forward:
                cls._ttb_reference_forward = cls.forward
                _set(cls, "forward", _attention.banded_forward)
                done.append("tt.transformer.RelLearnableMultiHeadAttn.forward")
            _set(_attention, "CONTEXT", (int(streaming_context[0]), int(streaming_context[1])))
        except ImportError:
            pass
        # espnet/nets/pytorch_backend/transformer/attention.py:264-308, same band (tt_espnet's encoder: 10 / 2)
        try:
            m = importlib.import_module("espnet.nets.pytorch_backend.transformer.attention")
            cls = m.RelPositionMultiHeadedAttention
            if cls.forward is not _attention.espnet_banded_forward:
                cls._ttb_reference_forward = cls.forward
                _set(cls, "forward", _attention.espnet_banded_forward)
                done.append("espnet...attention.RelPositionMultiHeadedAttention.forward")
            _set(_attention, "CONTEXT", (int(streaming_context[0]), int(streaming_context[1])))
        except ImportError:
            pass
    if patch_data:
        # tt/utils.py:297-329; train.py:18 copies the names at import (`from tt.utils import ...`)
        for modname in ("tt.utils", "train"):
            m = sys.modules.get(modname)
            if m is None and modname == "tt.utils":
                try:
                    m = importlib.import_module(modname)
                except ImportError:
                    m = None
            if m is not None and hasattr(m, "time_mask_augment"):
                _set(m, "time_mask_augment", _data.time_mask_augment)
                _set(m, "frequency_mask_augment", _data.frequency_mask_augment)
                done.append(modname + ".{time,frequency}_mask_augment")
    if patch_tt:
        try:
            m = importlib.import_module("tt.model")
            _set(m, "JointNet", JointNet)
            done.append("tt.model.JointNet")
            if patch_decode:
                _patch_decode(m.Transducer, _decode.tt_decode, "tt.model.Transducer.decode", done)
                _patch_decode(m.Transducer, _decode.tt_beam_search, "tt.model.Transducer.beam_search", done,
                              attr="beam_search")
                _patch_decode(m.Transducer, _decode.tt_recognize, "tt.model.Transducer.recognize", done,
                              attr="recognize")
        except ImportError:
            pass
    if patch_espnet:
        try:
            m = importlib.import_module("espnet.nets.pytorch_backend.transducer.joint_network")
            _set(m, "JointNetwork", JointNetwork)
            done.append("espnet...joint_network.JointNetwork")
        except ImportError:
            pass
        # modules that copied the name at import (`from ...joint_network import JointNetwork`): tt_espnet/model.py:14 and the
        # upstream-style caller espnet/nets/pytorch_backend/e2e_asr_transducer.py:21
        other = "espnet.nets.pytorch_backend.e2e_asr_transducer"
        if other in sys.modules and hasattr(sys.modules[other], "JointNetwork"):
            _set(sys.modules[other], "JointNetwork", JointNetwork)
            done.append("espnet...e2e_asr_transducer.JointNetwork")
        if "tt_espnet.model" in sys.modules:
            _set(sys.modules["tt_espnet.model"], "JointNetwork", JointNetwork)
            done.append("tt_espnet.model.JointNetwork")
            if patch_decode:
                _patch_decode(sys.modules["tt_espnet.model"].TransformerTransducer, _decode.espnet_decode,
                              "tt_espnet.model.TransformerTransducer.decode", done)
                _patch_decode(sys.modules["tt_espnet.model"].TransformerTransducer, _decode.espnet_recognize,
                              "tt_espnet.model.TransformerTransducer.recognize", done, attr="recognize")
    return done
