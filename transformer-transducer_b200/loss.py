"""Drop-in for the module the reference imports as ``warprnnt_pytorch`` (HawkAaron/warp-transducer's
PyTorch binding; /root/reference/train.py:13,53,231, espnet/nets/pytorch_backend/transducer/loss.py:22-25,74).

Same names, argument order, reductions and error behaviour as upstream: ``'mean'`` divides by the batch
size, ``'mean'``/``'sum'`` return shape ``(1,)``, labels / lengths must be int32, ``T == max(act_lens)``
and ``U + 1 == max(label_lens) + 1`` are checked on the host.  CUDA only -- the CPU path is the oracle,
which is test infrastructure and not part of the product.
"""
import os

import torch

from . import _lib
from . import functional as F
from .lazy import LazyJointLogits


_checked = {}


def certify_inputs(acts, labels, act_lens, label_lens):
    """Upstream's argument checks.  Returns the batch's real sizes (128-row lattice tiles sum_b ceil(T_b (U_b + 1) / 128),
    elements of the diagonal-major lattice arrays), taken from the same single host synchronisation as the length
    checks, or None when TTX_SKIP_LENGTH_CHECKS=1 skips that synchronisation (buffers are then sized by the dense
    upper bounds)."""
    for name, t in (("labels", labels), ("act_lens", act_lens), ("label_lens", label_lens)):
        if t.dtype != torch.int32:
            raise TypeError("%s must be int32" % name)
    if acts.dim() != 4:
        raise ValueError("acts must have 4 dimensions")
    if labels.dim() != 2:
        raise ValueError("labels must have 2 dimensions")
    if act_lens.dim() != 1:
        raise ValueError("act_lens must have 1 dimension")
    if label_lens.dim() != 1:
        raise ValueError("label_lens must have 1 dimension")
    B = acts.shape[0]
    if act_lens.shape[0] != B or label_lens.shape[0] != B or labels.shape[0] != B:
        raise ValueError("must have a length per example.")
    if os.environ.get("TTX_SKIP_LENGTH_CHECKS", "0") == "1":
        return None
    # The checks below cost one device -> host synchronisation, which also drains everything queued on the stream.  When
    # the very same tensors (storage, version counter, shape) were checked by the previous call with the same logits
    # shape -- a loop that keeps its batch resident -- the result cannot have changed and the synchronisation is skipped.
    key = tuple((t.data_ptr(), t._version, tuple(t.shape), t.device) for t in (labels, act_lens, label_lens)) + \
        (tuple(acts.shape),)
    if _checked.get("key") == key:
        return _checked["sizes"]
    if acts.is_cuda:
        # one launch + one 56-byte read (the values upstream's checks need, the batch's real lattice size, the label range)
        dev = acts.device
        lab = labels.to(dev).contiguous()
        al, ll = act_lens.to(dev).contiguous(), label_lens.to(dev).contiguous()
        out = torch.empty(7, dtype=torch.int64, device=dev)
        with F._guard(dev):
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            p = lambda t: t.data_ptr() if t.numel() else None  # noqa: E731
            _lib.check(_lib.get().ttx_check_inputs(p(lab), lab.shape[1], p(al), p(ll), B, acts.shape[3], p(out), idx,
                                                   torch._C._cuda_getCurrentRawStream(idx)), "ttx_check_inputs")
        m = out.tolist()                                         # the synchronisation
        mx = [m[0], m[1], m[2], m[3], m[4], m[5], m[6]]
    else:
        al, ll = act_lens.long(), label_lens.long().to(act_lens.device)
        tiles = ((al * (ll + 1) + 127) // 128).sum()
        lat = ((al + ll) * ((ll + 4) // 4 * 4)).sum()            # (T + U1 - 1) * pitch(U1), pitch = U1 rounded up to 4
        # labels inside the valid region must index the vocabulary (the kernels gather W_out rows / scatter into them)
        if labels.shape[1] > 0:
            lab = labels.to(act_lens.device)
            valid = torch.arange(labels.shape[1], device=lab.device)[None, :] < ll[:, None]
            bad = (valid & ((lab < 0) | (lab >= acts.shape[3]))).any().long()
        else:
            bad = tiles.new_zeros(())
        mx = torch.stack((al.max(), ll.max(), al.min(), ll.min(), tiles, bad, lat)).tolist()  # one sync
    if mx[0] != acts.shape[1]:
        raise ValueError("Input length mismatch")
    if mx[1] + 1 != acts.shape[2]:
        raise ValueError("Output length mismatch")
    if mx[2] < 1 or mx[3] < 0:
        raise ValueError("lengths must be positive")
    if labels.shape[1] < mx[1]:
        raise ValueError("labels is shorter than max(label_lens)")
    if mx[5]:
        raise ValueError("labels must lie in [0, %d) inside each utterance's label_lens" % acts.shape[3])
    # (the tensors are kept alive with the key: a freed and re-used allocation must not be mistaken for them)
    _checked.update(key=key, sizes=(int(mx[4]), int(mx[6])), keep=(labels, act_lens, label_lens))
    return _checked["sizes"]


def rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction="mean", fastemit_lambda=0.0):
    """Transducer loss.  ``acts``: (B,T,U+1,V) fp32 CUDA logits or the lazy handle of our joint modules."""
    if reduction not in ("none", "mean", "sum"):
        raise ValueError("reduction must be one of none, mean, sum")
    if fastemit_lambda != 0.0:
        raise NotImplementedError("fastemit_lambda is not part of the reference's call and is not supported")
    if not acts.is_cuda:
        raise RuntimeError("warprnnt_pytorch (B200): acts must be a CUDA tensor -- there is no CPU fallback")
    sizes = certify_inputs(acts, labels, act_lens, label_lens)
    if isinstance(acts, LazyJointLogits):
        ep, pp, w, b = acts.parts
        bf16 = ep.dtype == torch.bfloat16
        costs = F.fused_joint_rnnt(ep, pp, w, b, labels, act_lens, label_lens, blank, bf16, sizes=sizes,
                                   pre=getattr(acts, "pre", None))
    else:
        costs = F.dense_rnnt(acts, labels, act_lens, label_lens, blank, sizes=sizes)
    if reduction in ("sum", "mean"):
        costs = costs.sum().unsqueeze(-1)
        if reduction == "mean":
            costs = costs / acts.shape[0]
    return costs


class RNNTLoss(torch.nn.Module):
    """``RNNTLoss(blank=0, reduction='mean')(acts, labels, act_lens, label_lens)`` as in train.py:231,53."""

    def __init__(self, blank=0, reduction="mean", fastemit_lambda=0.0):
        super().__init__()
        if reduction not in ("none", "mean", "sum"):
            raise ValueError("reduction must be one of none, mean, sum")
        self.blank = blank
        self.reduction = reduction
        self.fastemit_lambda = fastemit_lambda

    def forward(self, acts, labels, act_lens, label_lens):
        return rnnt_loss(acts, labels, act_lens, label_lens, self.blank, self.reduction, self.fastemit_lambda)
