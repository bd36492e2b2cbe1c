"""Drop-in joint networks with the reference's constructor signatures, sub-module names and state-dict keys.

  JointNet      <-> /root/reference/tt/model.py:12-39      (forward_layer / tanh / project_layer)
  JointNetwork  <-> /root/reference/espnet/nets/pytorch_backend/transducer/joint_network.py:8-51
                    (lin_enc / lin_dec(no bias) / lin_out / joint_activation)

For batched CUDA inputs with a tanh joint of a supported width they return a ``LazyJointLogits`` handle
(the first Linear is split algebraically: cat(e,d) W^T = e W[:, :De]^T + d W[:, De:]^T, so it runs on
B*T + B*U rows instead of the reference's B*T*U).  Everything else -- 1-D decode inputs
(tt/model.py:77), CPU tensors, other activations, widths that are not a multiple of 64 -- is the reference's
dense math.
"""
import os

import torch

from . import functional as F
from .lazy import LazyJointLogits


class _tf32_matmul:
    """cuBLAS TF32 (fp32 accumulate) for the GEMMs issued inside the block, whatever the process-wide setting."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev
        return False


def _split_tf32(x):
    """x = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared) and lo = x - hi exact in fp32."""
    hi = (x.contiguous().view(torch.int32) & -8192).view(torch.float32)
    return hi, x - hi


def _mm3(a, b_t, passes):
    """a @ b_t^T on the tensor cores (cuBLAS TF32, fp32 accumulate).  passes = 3: error-compensated split
    a_hi b_hi + a_lo b_hi + a_hi b_lo (the dropped a_lo b_lo term is 2^-22 relative), i.e. fp32-grade results at three
    small tensor-core GEMMs; passes = 1: plain TF32 operands (2^-11 relative rounding)."""
    with _tf32_matmul():
        if passes == 1:
            return torch.matmul(a, b_t.t())
        a_hi, a_lo = _split_tf32(a)
        b_hi, b_lo = _split_tf32(b_t)
        a2 = a_hi.reshape(-1, a.shape[-1])
        y = torch.mm(a2, b_hi.t())
        y.addmm_(a_lo.reshape(-1, a.shape[-1]), b_hi.t())
        y.addmm_(a2, b_lo.t())
        return y.view(*a.shape[:-1], b_t.shape[0])


class _ProjTC(torch.autograd.Function):
    """y = x W^T (+ b) for the two small pre-projections of the joint (float32).  torch runs float32 GEMMs as fp32 SIMT
    kernels by default (0.58 ms per cfg2 step for the six of them).  `TTX_TF32_PROJ` selects where cuBLAS TF32 tensor-core
    GEMMs (fp32 accumulate) are used instead:
      0  nowhere (torch.nn.functional.linear and its autograd);
      2  (default) the four backward GEMMs only -- the forward stays exact because its rounding would reach the loss and
         every gradient; d_enc / d_pred / first-layer weight gradients move from ~1e-4 to ~3e-4 relative error against
         the oracle (tolerance 1e-3; the output-layer gradients are at 1-3e-4 anyway), ~0.3 ms per step;
      1  forward too (another 0.15 ms; every gradient at 3-6e-4);
      3  error-compensated 3 x TF32 everywhere (fp32-grade results, but as separate library launches no faster than
         the SIMT GEMMs -- one fused kernel is the next step)."""

    @staticmethod
    def forward(ctx, x, w, b, passes):
        ctx.save_for_backward(x, w)
        ctx.has_bias, ctx.passes = b is not None, passes
        if passes == 2:      # exact forward (its rounding would reach every gradient), TF32 only in the backward GEMMs
            return torch.nn.functional.linear(x, w, b)
        y = _mm3(x, w, passes)
        return y + b if b is not None else y

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
        bp = 1 if ctx.passes == 2 else ctx.passes
        dx = _mm3(dy, w.t(), bp) if ctx.needs_input_grad[0] else None
        dw = _mm3(dy2.t(), x2.t(), bp) if ctx.needs_input_grad[1] else None
        db = dy2.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, dw, db, None


def _proj(x, w, b=None):
    passes = int(os.environ.get("TTX_TF32_PROJ", "2"))
    if x.dtype == torch.float32 and w.dtype == torch.float32 and passes in (1, 2, 3):
        return _ProjTC.apply(x, w, b, passes)
    return torch.nn.functional.linear(x, w, b)


def _fusable(x, width):
    # widths covered by the fused tcgen05 kernels, or any other multiple of 64 (chunked path, see functional.py)
    return x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and width % 64 == 0


class JointNet(torch.nn.Module):
    def __init__(self, input_size, inner_dim, vocab_size):
        super().__init__()
        self.forward_layer = torch.nn.Linear(input_size, inner_dim, bias=True)
        self.tanh = torch.nn.Tanh()
        self.project_layer = torch.nn.Linear(inner_dim, vocab_size, bias=True)
        self.fused = True

    def forward(self, enc_state, dec_state):
        if (self.fused and enc_state.dim() == 3 and dec_state.dim() == 3 and
                enc_state.size(-1) + dec_state.size(-1) == self.forward_layer.in_features and
                _fusable(enc_state, self.forward_layer.out_features)):
            de = enc_state.size(-1)
            w = self.forward_layer.weight
            eproj = _proj(enc_state, w[:, :de], self.forward_layer.bias)
            pproj = _proj(dec_state, w[:, de:])
            return LazyJointLogits(eproj, pproj, self.project_layer.weight, self.project_layer.bias)
        if enc_state.dim() == 3 and dec_state.dim() == 3:  # tt/model.py:21-29
            t, u = enc_state.size(1), dec_state.size(1)
            enc_state = enc_state.unsqueeze(2).expand(-1, -1, u, -1)
            dec_state = dec_state.unsqueeze(1).expand(-1, t, -1, -1)
        else:
            assert enc_state.dim() == dec_state.dim()
        x = torch.cat((enc_state, dec_state), dim=-1)
        return self.project_layer(self.tanh(self.forward_layer(x)))


_ACTIVATIONS = {"hardtanh": torch.nn.Hardtanh, "tanh": torch.nn.Tanh, "relu": torch.nn.ReLU, "selu": torch.nn.SELU}


class _Swish(torch.nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


class JointNetwork(torch.nn.Module):
    def __init__(self, vocab_size, encoder_output_size, decoder_output_size, joint_space_size,
                 joint_activation_type="tanh"):
        super().__init__()
        self.lin_enc = torch.nn.Linear(encoder_output_size, joint_space_size)
        self.lin_dec = torch.nn.Linear(decoder_output_size, joint_space_size, bias=False)
        self.lin_out = torch.nn.Linear(joint_space_size, vocab_size)
        if joint_activation_type == "swish":
            self.joint_activation = _Swish()
        else:
            self.joint_activation = _ACTIVATIONS[joint_activation_type]()  # nets_utils.py:501-514
        self.joint_activation_type = joint_activation_type
        self.fused = True

    def forward(self, h_enc, h_dec):
        if (self.fused and self.joint_activation_type == "tanh" and h_enc.dim() == 4 and h_dec.dim() == 4 and
                h_enc.size(2) == 1 and h_dec.size(1) == 1 and h_enc.size(0) == h_dec.size(0) and
                _fusable(h_enc, self.lin_enc.out_features)):
            eproj = _proj(h_enc.squeeze(2), self.lin_enc.weight, self.lin_enc.bias)
            pproj = _proj(h_dec.squeeze(1), self.lin_dec.weight)
            return LazyJointLogits(eproj, pproj, self.lin_out.weight, self.lin_out.bias)
        z = self.joint_activation(self.lin_enc(h_enc) + self.lin_dec(h_dec))  # joint_network.py:48-49
        return self.lin_out(z)
