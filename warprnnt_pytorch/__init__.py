"""``warprnnt_pytorch`` as the reference imports it (/root/reference/train.py:13,
espnet/nets/pytorch_backend/transducer/loss.py:23): the B200 implementation behind the upstream names."""
from transformer_transducer_b200.loss import RNNTLoss, rnnt_loss  # noqa: F401

__all__ = ["RNNTLoss", "rnnt_loss"]
