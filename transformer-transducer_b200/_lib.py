"""Loads (and, on request, builds) libttx.so -- the C-ABI CUDA library declared in include/ttx.h.

No fallback: if the library is missing or a call fails the caller gets an exception.
"""
import ctypes
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libttx.so")
SOURCES = ["ttx_api.cu", "ttx_small.cu", "ttx_joint_mma.cu", "ttx_wide.cu", "ttx_proj.cu", "ttx_decode.cu", "ttx_attn.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None

c_i32, c_i64, c_p = ctypes.c_int, ctypes.c_int64, ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES)
_PROTOS = {
    "ttx_version": [],
    "ttx_last_error": [],
    "ttx_supported_h": [c_i32],
    "ttx_tiles_upper_bound": [c_i32, c_i32, c_i32],
    "ttx_meta_ints": [c_i32, c_i64],
    "ttx_prepare": [c_p, c_p, c_i32, c_i32, c_i32, c_i64, c_p, c_i32, c_p],
    "ttx_cast_weight": [c_p, c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_i32, c_p],
    "ttx_joint_act": [c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64, c_i32, c_p, c_p, c_p,
                      c_i32, c_p],
    "ttx_joint_lse_fwd": [c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_i32, c_p],
    "ttx_lattice_elems_upper_bound": [c_i32, c_i32, c_i32],
    "ttx_lattice_fwd_bwd": [c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i64, c_i64, c_p, c_p, c_p, c_p, c_p, c_i32, c_p],
    "ttx_grad_coeffs": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i64, c_p, c_p,
                        c_i32, c_p],
    "ttx_transpose16": [c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p],
    "ttx_fwd_grad_supported_h": [c_i32],
    "ttx_joint_fwd_grad": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p,
                           c_p, c_i64, c_i32, c_p],
    "ttx_joint_workspace_bytes": [c_i32, c_i64, c_i32, c_i32, c_i32],
    "ttx_reduce_act_grad_ew": [c_p, c_p, c_p, c_p, c_p, c_i32, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32,
                               c_p, c_p, c_i32, c_p],
    "ttx_wide_supported_h": [c_i32],
    "ttx_wide_sp": [c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_p, c_p,
                    c_p, c_i64, c_p, c_i32, c_p],
    "ttx_wide_pw": [c_p, c_i64, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p],
    "ttx_wide_dw": [c_p, c_i64, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_i32, c_p],
    "ttx_kept_prepare": [c_p] * 10 + [c_i32, c_i32, c_i32, c_i64, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_i32, c_i32, c_p],
    "ttx_rows_lse": [c_p, c_i32, c_i32, c_i32, c_p, c_p, c_p, c_i32, c_p, c_p, c_p, c_i32, c_p],
    "ttx_rows_grad": [c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p],
    "ttx_joint_grad": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i64, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_p,
                       c_i32, c_p, c_i64, c_i32, c_p],
    "ttx_reduce_act_grad": [c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_i32, c_p],
    "ttx_proj_fwd": [c_p, c_i32, c_p, c_i32, c_p, c_i32, c_i32, c_i32, c_p, c_i32, c_i32, c_p],
    "ttx_proj_bwd_x": [c_p, c_i32, c_p, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_i32, c_p],
    "ttx_proj_bwd_w": [c_p, c_i32, c_p, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p, c_i32, c_p],
    "ttx_decode_scan": [c_p, c_i32, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_p, c_p, c_i32, c_p],
    "ttx_spec_mask": [c_p, c_i32, c_i32, c_i32, c_i64, c_i64, c_p, c_i32, c_i32, c_p],
    "ttx_band_attn_fwd": [c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_float, c_i32, c_p,
                          c_p, c_p, c_i32, c_p],
    "ttx_band_attn_bwd": [c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, ctypes.c_float, c_i32, c_p,
                          c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_p],
    "ttx_check_inputs": [c_p, c_i32, c_p, c_p, c_i32, c_i32, c_p, c_i32, c_p],
    "ttx_dense_lse": [c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i64, c_p, c_p, c_p, c_p,
                      c_i32, c_p],
    "ttx_dense_grad": [c_p, c_p, c_p, c_p, c_p, c_p, c_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_p, c_i32, c_p],
}
_RESTYPES = {"ttx_last_error": ctypes.c_char_p, "ttx_tiles_upper_bound": c_i64, "ttx_meta_ints": c_i64,
             "ttx_lattice_elems_upper_bound": c_i64, "ttx_joint_workspace_bytes": c_i64}
EXPORTS = tuple(_PROTOS)


class TTXError(RuntimeError):
    pass


def build(force=False, verbose=False):
    """Compile csrc/*.cu for sm_100a into lib/libttx.so (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, "ttx_common.cuh"), os.path.join(ROOT, "include", "ttx.h")]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + ".tmp%d" % os.getpid()
    flags = NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else [])
    objs = [os.path.join(LIB_DIR, os.path.basename(s)[:-3] + ".o") for s in srcs]
    procs = [subprocess.Popen(["nvcc"] + flags + ["-c", "-o", o, s], cwd=CSRC) for s, o in zip(srcs, objs)]   # in parallel
    failed = [s for s, p in zip(srcs, procs) if p.wait() != 0]
    if failed:
        raise subprocess.CalledProcessError(1, "nvcc -c " + " ".join(failed))
    subprocess.check_call(["nvcc"] + flags + ["-shared", "-o", tmp] + objs, cwd=CSRC)
    for o in objs:
        os.remove(o)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


def get():
    """The loaded library (ctypes.CDLL with prototypes set).  Raises TTXError if it was never built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise TTXError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                                   "(there is no CPU / library fallback)" % LIB_PATH)
                lib = ctypes.CDLL(LIB_PATH)
                for name, argtypes in _PROTOS.items():
                    fn = getattr(lib, name)
                    fn.argtypes = argtypes
                    fn.restype = _RESTYPES.get(name, c_i32)
                _lib = lib
    return _lib


def check(rc, what):
    if rc != 0:
        msg = get().ttx_last_error()
        raise TTXError("%s failed (status %d): %s" % (what, rc, msg.decode() if msg else "?"))
