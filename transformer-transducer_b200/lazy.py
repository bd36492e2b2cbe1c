"""``LazyJointLogits`` -- what the drop-in joint modules return instead of the (B,T,U+1,V) logits.

The reference hands the joint's output across user code before it reaches the loss
(/root/reference/tt/model.py:66-68 -> train.py:51-53; tt_espnet/model.py:73-80), so the handle has to
look like the logits tensor (shape, dtype, device, ``.to(dtype=...)`` as used by
espnet/nets/pytorch_backend/transducer/loss.py:57-60) while carrying only the small pre-projected
operands.  ``RNNTLoss`` recognises it and runs the fused kernels; any other tensor operation
materialises the logits explicitly (guarded by a size limit) -- nothing allocates B*T*U*V silently.
"""
import os

import torch

_ALLOWED_NOOP = ("aten.detach.default", "aten.alias.default")
# torch-function names that only look at / re-type the handle: they keep it lazy (everything else materialises)
_META = ("__get__", "size", "dim", "ndimension", "is_contiguous", "contiguous", "to", "float", "type", "__repr__",
         "__str__", "__format__", "numel", "nelement", "is_floating_point", "get_device", "stride", "__len__",
         "__bool__", "__dir__", "__hash__", "requires_grad_", "data_ptr", "__reduce_ex__", "element_size",
         "storage_offset", "is_complex", "is_pinned")


def _limit_bytes():
    return float(os.environ.get("TTX_MATERIALIZE_LIMIT_GB", "8")) * (1 << 30)


class LazyJointLogits(torch.Tensor):
    @staticmethod
    def __new__(cls, eproj, pproj, w_out, b_out, dtype=None, normalised=False):
        B, T, _ = eproj.shape
        U1 = pproj.shape[1]
        V = w_out.shape[0]
        r = torch.Tensor._make_wrapper_subclass(cls, (B, T, U1, V), dtype=dtype or eproj.dtype, device=eproj.device,
                                                requires_grad=False)
        r.eproj, r.pproj, r.w_out, r.b_out = eproj, pproj, w_out, b_out
        r.pre = None        # (enc, w_enc, b_enc, dec, w_dec) when the pre-projections ran on our kernels (joint._handle)
        # True: the handle stands for log_softmax(logits, -1) (espnet's "warp-rnnt" branch, transducer/loss.py:61-62).  The
        # loss normalises its input anyway, so the kernels see no difference; materialize() applies the log_softmax.
        r.normalised = bool(normalised)
        return r

    def __repr__(self):
        return "LazyJointLogits(shape=%s, dtype=%s, device=%s%s)" % (tuple(self.shape), self.dtype, self.device,
                                                                   ", log_softmax" if self.normalised else "")

    def _like(self, parts=None, dtype=None, normalised=None):
        out = LazyJointLogits(*(parts or self.parts), dtype=dtype or self.dtype,
                              normalised=self.normalised if normalised is None else normalised)
        out.pre = self.pre if parts is None else None
        return out

    @property
    def parts(self):
        return self.eproj, self.pproj, self.w_out, self.b_out

    def materialize(self):
        """Dense logits by plain torch ops (differentiable w.r.t. the parts); refuses above the size limit."""
        nbytes = self.numel() * 4
        if nbytes > _limit_bytes():
            raise RuntimeError("refusing to materialise %.1f GB of joint logits (set TTX_MATERIALIZE_LIMIT_GB to "
                               "override); pass the handle to RNNTLoss instead" % (nbytes / (1 << 30)))
        h = torch.tanh(self.eproj.unsqueeze(2) + self.pproj.unsqueeze(1))
        z = torch.nn.functional.linear(h, self.w_out, self.b_out)
        if self.normalised:
            z = torch.log_softmax(z.float(), dim=-1)
        return z.to(self.dtype)

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        """Anything but the loss and the metadata calls above gets the dense logits -- materialised WITH autograd, so
        slicing the handle, a log_softmax for another loss, or mixing it into an auxiliary loss keeps gradients to the
        joint's inputs and parameters (guarded by the size limit)."""
        kwargs = kwargs or {}
        name = getattr(func, "__name__", "")
        if name in _META:
            return super().__torch_function__(func, types, args, kwargs)
        if name == "detach":
            src = args[0]
            return src._like(parts=[t.detach() for t in src.parts])
        if name == "log_softmax" and args and isinstance(args[0], LazyJointLogits):
            # log_softmax over the vocabulary keeps the handle lazy (transducer/loss.py:61-62 in front of warp_rnnt)
            src = args[0]
            dim = kwargs.get("dim", args[1] if len(args) > 1 else None)
            if dim in (-1, src.dim() - 1) and len(args) <= 2:
                return src._like(dtype=kwargs.get("dtype", None), normalised=True)

        def unwrap(x):
            return x.materialize() if isinstance(x, LazyJointLogits) else x

        with torch._C.DisableTorchFunctionSubclass():
            return func(*torch.utils._pytree.tree_map(unwrap, args), **torch.utils._pytree.tree_map(unwrap, kwargs))

    @classmethod
    def __torch_dispatch__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = str(func)
        if name in _ALLOWED_NOOP:
            return args[0]
        if name == "aten._to_copy.default":
            src = args[0]
            dev = kwargs.get("device", None)
            if dev is None or torch.device(dev) == src.device:
                return src._like(dtype=kwargs.get("dtype", None) or src.dtype)

        def unwrap(x):
            return x.materialize().detach() if isinstance(x, LazyJointLogits) else x

        return func(*torch.utils._pytree.tree_map(unwrap, args), **torch.utils._pytree.tree_map(unwrap, kwargs))
