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

from . import functional as F
from .lazy import LazyJointLogits


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
            eproj = torch.nn.functional.linear(enc_state, w[:, :de], self.forward_layer.bias)
            pproj = torch.nn.functional.linear(dec_state, w[:, de:])
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
            eproj = self.lin_enc(h_enc.squeeze(2))
            pproj = self.lin_dec(h_dec.squeeze(1))
            return LazyJointLogits(eproj, pproj, self.lin_out.weight, self.lin_out.bias)
        z = self.joint_activation(self.lin_enc(h_enc) + self.lin_dec(h_dec))  # joint_network.py:48-49
        return self.lin_out(z)
