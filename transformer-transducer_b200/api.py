"""Public surface of the package."""
from . import _lib
from .data import LengthBucketSampler, crop_to_max, frequency_mask_augment, mask_augment, time_mask_augment
from .decode import StreamingGreedy, beam_search, greedy_search
from .functional import dense_rnnt, fused_joint_rnnt, supported_width
from .install import install, uninstall
from .joint import JointNet, JointNetwork
from .lazy import LazyJointLogits
from .loss import RNNTLoss, certify_inputs, rnnt_loss

build = _lib.build
TTXError = _lib.TTXError

__all__ = ["JointNet", "JointNetwork", "LazyJointLogits", "RNNTLoss", "rnnt_loss", "certify_inputs",
           "fused_joint_rnnt", "dense_rnnt", "supported_width", "greedy_search", "beam_search", "StreamingGreedy", "install",
           "uninstall", "LengthBucketSampler", "crop_to_max", "mask_augment", "time_mask_augment", "frequency_mask_augment",
           "build", "TTXError"]
