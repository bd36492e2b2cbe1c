"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the UNMODIFIED reference modules from /root/reference (build container only; the GPU box
has no copy -- callers must check ``available()``).  The reference's module-top imports of packages
that are not installed and not used on the joint/loss path (tt/utils.py:5-8, train.py:12,
augment/speed_augment.py:8) are satisfied with empty stub modules.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("TT_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "tt", "model.py"))


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    return sys.modules[name]


def prepare(stub_train_deps=False):
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    for n in ("librosa", "editdistance"):
        try:
            __import__(n)
        except ImportError:
            _stub(n)
    try:
        import matplotlib.pyplot  # noqa: F401
    except ImportError:
        mpl = _stub("matplotlib")
        mpl.pyplot = _stub("matplotlib.pyplot")
    if stub_train_deps:
        try:
            import tensorboardX  # noqa: F401
        except ImportError:
            _stub("tensorboardX", SummaryWriter=object)
        try:
            import pydub  # noqa: F401
        except ImportError:
            _stub("pydub", AudioSegment=object)
    if REF_ROOT not in sys.path:
        sys.path.append(REF_ROOT)


def tt_model():
    prepare()
    import tt.model
    return tt.model


def espnet_joint_module():
    prepare()
    import espnet.nets.pytorch_backend.transducer.joint_network as m
    return m


def espnet_loss_module():
    prepare()
    import espnet.nets.pytorch_backend.transducer.loss as m
    return m
