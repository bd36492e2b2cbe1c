"""Rebinds the reference's joint classes to the B200 drop-ins without touching reference files.

``Transducer.__init__`` looks up the module global ``JointNet`` at call time (/root/reference/tt/model.py:48)
and ``tt_espnet/model.py:14`` copies ``JointNetwork`` at import, so patching the defining modules (and
``tt_espnet.model`` if it is already imported) before the model is constructed is enough.  The loss
needs no patch: ``from warprnnt_pytorch import RNNTLoss`` (train.py:13) finds the ``warprnnt_pytorch``
package of this repository once the repository root is on ``sys.path``.
"""
import importlib
import sys

from .joint import JointNet, JointNetwork


def install(patch_tt=True, patch_espnet=True):
    done = []
    if patch_tt:
        try:
            m = importlib.import_module("tt.model")
            m.JointNet = JointNet
            done.append("tt.model.JointNet")
        except ImportError:
            pass
    if patch_espnet:
        try:
            m = importlib.import_module("espnet.nets.pytorch_backend.transducer.joint_network")
            m.JointNetwork = JointNetwork
            done.append("espnet...joint_network.JointNetwork")
        except ImportError:
            pass
        if "tt_espnet.model" in sys.modules:
            sys.modules["tt_espnet.model"].JointNetwork = JointNetwork
            done.append("tt_espnet.model.JointNetwork")
    return done
