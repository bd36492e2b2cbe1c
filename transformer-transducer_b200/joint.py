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

import torch

from . import _lib
from . import functional as F
from .lazy import LazyJointLogits


def _p(t):
    return t.data_ptr() if t is not None else None


class _Proj(torch.autograd.Function):
    """y = x W^T (+ b) for the two small pre-projections of the joint (tt/model.py:35, joint_network.py:28-31,48), float32,
    forward and backward on our tcgen05 kernels (csrc/ttx_proj.cu: TF32 tensor-core products with on-chip error
    compensation, fp32-grade like the reference's SGEMM).  W may be a column slice of a wider matrix (the split
    forward_layer): only its row stride is used."""

    @staticmethod
    def forward(ctx, x, w, b):
        lib = _lib.get()
        dev = x.device
        N, K = w.shape
        x2 = x.detach().reshape(-1, K).contiguous()
        M = x2.shape[0]
        y = torch.empty(M, N, dtype=torch.float32, device=dev)
        with F._guard(dev):
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            st = torch._C._cuda_getCurrentRawStream(idx)
            bb = b.detach().contiguous() if b is not None else None
            F._call("ttx_proj_fwd", dev, _p(x2), K, _p(w), w.stride(0), _p(bb), M, N, K, _p(y), N, idx, st)
        ctx.save_for_backward(x2, w)
        ctx.has_bias, ctx.x_shape = b is not None, x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        dev = x2.device
        N, K = w.shape
        M = x2.shape[0]
        dy2 = dy.reshape(M, N).contiguous()
        dx = dw = db = None
        with F._guard(dev):
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            st = torch._C._cuda_getCurrentRawStream(idx)
            if ctx.needs_input_grad[0]:
                dx = torch.empty(M, K, dtype=torch.float32, device=dev)
                F._call("ttx_proj_bwd_x", dev, _p(dy2), N, _p(w), w.stride(0), M, N, K, _p(dx), K, idx, st)
                dx = dx.view(ctx.x_shape)
            if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
                dw = torch.zeros(N, K, dtype=torch.float32, device=dev)
                db = torch.zeros(N, dtype=torch.float32, device=dev) if ctx.has_bias else None
                F._call("ttx_proj_bwd_w", dev, _p(dy2), N, _p(x2), K, M, N, K, _p(dw), K, _p(db), idx, st,
                        n_kernels=2 if db is not None else 1)
        return dx, dw, db


def _on_kernels(x, w, b=None):
    return (x.is_cuda and x.dtype == torch.float32 and w.dtype == torch.float32 and w.stride(1) == 1 and
            w.shape[0] % 4 == 0 and w.shape[1] % 4 == 0 and w.stride(0) % 4 == 0 and w.data_ptr() % 16 == 0 and
            (b is None or b.dtype == torch.float32) and x.numel() > 0)


def _proj(x, w, b=None):
    """The projection on our kernels when the operands allow it (float32 CUDA, 16-byte aligned rows), else torch's."""
    if _on_kernels(x, w, b):
        return _Proj.apply(x, w, b)
    return torch.nn.functional.linear(x, w, b)


def _handle(enc, w_enc, b_enc, dec, w_dec, w_out, b_out):
    """The lazy logits handle of a batched call.  When both pre-projections run on our kernels the handle also carries
    their inputs, so that the loss node can own their backward (functional.WideJointRNNT)."""
    h = LazyJointLogits(_proj(enc, w_enc, b_enc), _proj(dec, w_dec), w_out, b_out)
    if _on_kernels(enc, w_enc, b_enc) and _on_kernels(dec, w_dec):
        h.pre = (enc, w_enc, b_enc, dec, w_dec)
    return h


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
            return _handle(enc_state, w[:, :de], self.forward_layer.bias, dec_state, w[:, de:],
                           self.project_layer.weight, self.project_layer.bias)
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
            return _handle(h_enc.squeeze(2), self.lin_enc.weight, self.lin_enc.bias, h_dec.squeeze(1), self.lin_dec.weight,
                           self.lin_out.weight, self.lin_out.bias)
        z = self.joint_activation(self.lin_enc(h_enc) + self.lin_dec(h_dec))  # joint_network.py:48-49
        return self.lin_out(z)
