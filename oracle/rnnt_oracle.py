"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU transducer loss with the semantics of ``warprnnt_pytorch.RNNTLoss`` as the reference uses it
(/root/reference/train.py:53,231; /root/reference/espnet/nets/pytorch_backend/transducer/loss.py:22-25,74).
``warprnnt_pytorch`` (HawkAaron/warp-transducer, un-pinned, requirements.txt:6) is absent from the
reference tree, so the arithmetic is restated in oracle/rnnt_cpu.c; this file restates the Python
binding around it: CPU inputs get ``log_softmax`` outside the kernel, gradients are computed during
forward and scaled by ``grad_output`` in backward, ``'mean'`` divides by the batch size and
``'mean'``/``'sum'`` return shape ``(1,)``.

Also here: ``rnnt_loss_fp64`` -- an independent float64 pure-torch restatement (autograd does the
gradient) used as the arbiter, and ``certify_inputs`` restating upstream's argument checks.
"""
import ctypes

import torch

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build())
        _lib.oracle_rnnt_f32.restype = ctypes.c_int
        _lib.oracle_rnnt_f64.restype = ctypes.c_int
        _lib.oracle_lattice_f64.restype = ctypes.c_int
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def certify_inputs(acts, labels, act_lens, label_lens):
    """Upstream ``certify_inputs`` (SURVEY.md section 8(b) 'Error conventions')."""
    for name, t in (("labels", labels), ("act_lens", act_lens), ("label_lens", label_lens)):
        if t.dtype != torch.int32:
            raise TypeError("%s must be int32" % name)
    if acts.dim() != 4:
        raise ValueError("acts must have 4 dimensions")
    if labels.dim() != 2:
        raise ValueError("labels must have 2 dimensions")
    if act_lens.dim() != 1 or label_lens.dim() != 1:
        raise ValueError("lengths must have 1 dimension")
    if act_lens.shape[0] != acts.shape[0] or label_lens.shape[0] != acts.shape[0]:
        raise ValueError("must have a length per example.")
    if int(act_lens.max()) != acts.shape[1]:
        raise ValueError("Input length mismatch")
    if int(label_lens.max()) + 1 != acts.shape[2]:
        raise ValueError("Output length mismatch")


class _RNNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, log_probs, labels, act_lens, label_lens, blank, reduction):
        certify_inputs(log_probs, labels, act_lens, label_lens)
        f64 = log_probs.dtype == torch.float64          # float64 in -> float64 arbiter; otherwise upstream's float32
        lp = log_probs.detach().contiguous().to(torch.float64 if f64 else torch.float32)
        B, T, U1, V = lp.shape
        labels_c = labels.contiguous()
        costs = torch.zeros(B, dtype=lp.dtype)
        grads = torch.empty_like(lp)
        fn = lib().oracle_rnnt_f64 if f64 else lib().oracle_rnnt_f32
        rc = fn(_p(lp), _p(labels_c), _p(act_lens.contiguous()), _p(label_lens.contiguous()),
                B, T, U1, V, int(blank), _p(costs), _p(grads))
        if rc:
            raise ValueError("bad lengths for utterance %d" % (rc - 1))
        if reduction in ("sum", "mean"):
            costs = costs.sum().unsqueeze(-1)
            if reduction == "mean":
                costs = costs / B
                grads = grads / B
        ctx.grads = grads
        return costs

    @staticmethod
    def backward(ctx, grad_output):
        g = grad_output.view(-1, 1, 1, 1).to(ctx.grads)
        return ctx.grads * g, None, None, None, None, None


def rnnt_loss(acts, labels, act_lens, label_lens, blank=0, reduction="mean"):
    if reduction not in ("none", "mean", "sum"):
        raise ValueError("reduction must be none, mean or sum")
    log_probs = torch.nn.functional.log_softmax(acts, -1)  # CPU binding does this outside the kernel
    return _RNNT.apply(log_probs, labels, act_lens, label_lens, blank, reduction)


class RNNTLoss(torch.nn.Module):
    def __init__(self, blank=0, reduction="mean"):
        super().__init__()
        self.blank = blank
        self.reduction = reduction

    def forward(self, acts, labels, act_lens, label_lens):
        return rnnt_loss(acts, labels, act_lens, label_lens, self.blank, self.reduction)


def rnnt_loss_fp64(logits, labels, act_lens, label_lens, blank=0):
    """Independent arbiter: float64, pure torch, per-utterance costs (B,), differentiable by autograd.

    Anti-diagonal-free formulation: row-by-row over t with a cumulative scan over u done in Python,
    so keep it to small cases (seconds for T*U up to a few thousand).
    """
    lp = torch.log_softmax(logits.double(), -1)
    B = lp.shape[0]
    costs = []
    for b in range(B):
        T, U = int(act_lens[b]), int(label_lens[b])
        lpb = lp[b, :T, : U + 1, blank]
        if U > 0:
            idx = labels[b, :U].long().view(1, U, 1).expand(T, U, 1)
            lpl = lp[b, :T, :U].gather(2, idx).squeeze(2)
        row = [lpb.new_zeros(())]
        for u in range(1, U + 1):
            row.append(row[u - 1] + lpl[0, u - 1])
        for t in range(1, T):
            new = [row[0] + lpb[t - 1, 0]]
            for u in range(1, U + 1):
                new.append(torch.logaddexp(row[u] + lpb[t - 1, u], new[u - 1] + lpl[t, u - 1]))
            row = new
        costs.append(-(row[U] + lpb[T - 1, U]))
    return torch.stack(costs)


def lattice_fp64(lpb, lpl, act_lens, label_lens):
    """float64 alpha/beta/costs on gathered log-probs (B,T,U1) via oracle/rnnt_cpu.c."""
    lpb = lpb.detach().double().contiguous()
    lpl = lpl.detach().double().contiguous()
    B, T, U1 = lpb.shape
    alpha = torch.full_like(lpb, float("nan"))
    beta = torch.full_like(lpb, float("nan"))
    costs = torch.zeros(B, dtype=torch.float64)
    rc = lib().oracle_lattice_f64(_p(lpb), _p(lpl), _p(act_lens.int().contiguous()), _p(label_lens.int().contiguous()),
                                  B, T, U1, _p(alpha), _p(beta), _p(costs))
    if rc:
        raise ValueError("bad lengths for utterance %d" % (rc - 1))
    return alpha, beta, costs
