"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Imports the UNMODIFIED reference modules from /root/reference (build container) or from the git-ignored
copy ``baseline/_ref/`` that ``stage()`` makes of the reference's Python / YAML files (it travels to the
GPU box with the repository snapshot, like built .so files) -- callers must check ``available()``.  The reference's module-top imports of packages
that are not installed and not used on the joint/loss path (tt/utils.py:5-8, train.py:12,
augment/speed_augment.py:8) are satisfied with empty stub modules.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(_REPO, "baseline", "_ref")


def _find_root():
    for cand in (os.environ.get("TT_REFERENCE_ROOT"), "/root/reference", STAGED):
        if cand and os.path.exists(os.path.join(cand, "tt", "model.py")):
            return cand
    return os.environ.get("TT_REFERENCE_ROOT", "/root/reference")


REF_ROOT = _find_root()


def available():
    return os.path.exists(os.path.join(REF_ROOT, "tt", "model.py"))


def stage(src="/root/reference"):
    """Copy the reference's own Python sources and YAML configs (unmodified) into baseline/_ref/ so that the
    GPU box, which has no /root/reference, can run the reference's callers and its CPU joint.  The directory is
    git-ignored: reference sources never enter this repository's history."""
    import shutil
    if not os.path.exists(os.path.join(src, "tt", "model.py")):
        return None
    keep = ("tt", "tt_espnet", "espnet", "espnet2", "config", "augment")
    for top in keep:
        for dirpath, _, files in os.walk(os.path.join(src, top)):
            for f in files:
                if f.endswith((".py", ".yaml", ".txt")):      # espnet/version.txt is read at import
                    rel = os.path.relpath(os.path.join(dirpath, f), src)
                    dst = os.path.join(STAGED, rel)
                    os.makedirs(os.path.dirname(dst), exist_ok=True)
                    shutil.copyfile(os.path.join(dirpath, f), dst)
    for f in ("train.py", "train_esptt.py"):
        shutil.copyfile(os.path.join(src, f), os.path.join(STAGED, f))
    return STAGED


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
