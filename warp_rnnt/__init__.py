"""``warp_rnnt`` as espnet's ``TransLoss(trans_type="warp-rnnt")`` imports it
(/root/reference/espnet/nets/pytorch_backend/transducer/loss.py:27-31,61-72): ``rnnt_loss`` over LOG-PROBABILITIES with
1ytic/warp-rnnt's argument names, on the B200 kernels.

``log_probs`` is either the lazy handle of this repo's joint modules after ``torch.log_softmax(handle, dim=-1)`` (which
keeps the handle lazy, so this branch runs the same fused path as ``warprnnt_pytorch``) or a dense (B, T, U+1, V) CUDA
tensor of log-probabilities.  The transducer loss normalises over the vocabulary itself and log_softmax is idempotent,
so the costs equal upstream's; the gradient with respect to the LOGITS behind the log_softmax is upstream's too (for a
dense input the gradient of ``log_probs`` itself additionally carries the softmax term that upstream leaves to
autograd's log_softmax backward, where it cancels).  ``gather`` / ``compact`` are upstream's memory optimisations and
change nothing here; ``fastemit_lambda`` must be 0.
"""
from transformer_transducer_b200.loss import rnnt_loss as _rnnt_loss

__all__ = ["rnnt_loss"]


def rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, average_frames=False, reduction=None, blank=0,
              gather=False, fastemit_lambda=0.0, compact=False):
    if reduction not in (None, "none", "sum", "mean"):
        raise ValueError("Unknown reduction method: {}".format(reduction))
    if compact:
        raise NotImplementedError("warp_rnnt: the compact memory layout is not part of the reference's call")
    costs = _rnnt_loss(log_probs, labels, frames_lengths, labels_lengths, blank=blank, reduction="none",
                       fastemit_lambda=fastemit_lambda)
    if average_frames:
        costs = costs / frames_lengths.to(costs)
    if reduction == "sum":
        return costs.sum()
    if reduction == "mean":
        return costs.mean()
    return costs
