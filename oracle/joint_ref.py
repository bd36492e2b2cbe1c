"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restatement of the reference's two joint networks as plain functions / minimal modules with the
reference's parameter names, used (a) as the checker for the CUDA path and (b) as the joint of the
timed CPU baseline on the GPU box, where /root/reference does not exist.

  * ``TTJointNet``        follows /root/reference/tt/model.py:12-39
  * ``EspnetJointNetwork`` follows /root/reference/espnet/nets/pytorch_backend/transducer/joint_network.py:8-51

tests/test_oracle.py pins both against the real reference modules (imported from /root/reference in
the build container) through the committed fixtures in tests/golden/.
"""
import torch


class TTJointNet(torch.nn.Module):
    def __init__(self, input_size, inner_dim, vocab_size):
        super().__init__()
        self.forward_layer = torch.nn.Linear(input_size, inner_dim, bias=True)
        self.tanh = torch.nn.Tanh()
        self.project_layer = torch.nn.Linear(inner_dim, vocab_size, bias=True)

    def forward(self, enc_state, dec_state):
        # tt/model.py:21-29 -- 3-D inputs are broadcast against each other (the reference does it
        # with .repeat; expand gives the same values)
        if enc_state.dim() == 3 and dec_state.dim() == 3:
            t, u = enc_state.size(1), dec_state.size(1)
            enc_state = enc_state.unsqueeze(2).expand(-1, -1, u, -1)
            dec_state = dec_state.unsqueeze(1).expand(-1, t, -1, -1)
        else:
            assert enc_state.dim() == dec_state.dim()  # tt/model.py:31
        x = torch.cat((enc_state, dec_state), dim=-1)  # tt/model.py:33
        return self.project_layer(self.tanh(self.forward_layer(x)))  # tt/model.py:35-37


class EspnetJointNetwork(torch.nn.Module):
    def __init__(self, vocab_size, encoder_output_size, decoder_output_size, joint_space_size,
                 joint_activation_type="tanh"):
        super().__init__()
        self.lin_enc = torch.nn.Linear(encoder_output_size, joint_space_size)
        self.lin_dec = torch.nn.Linear(decoder_output_size, joint_space_size, bias=False)
        self.lin_out = torch.nn.Linear(joint_space_size, vocab_size)
        acts = {"hardtanh": torch.nn.Hardtanh, "tanh": torch.nn.Tanh, "relu": torch.nn.ReLU,
                "selu": torch.nn.SELU}  # nets_utils.py:501-514 (swish omitted: needs the espnet class)
        self.joint_activation = acts[joint_activation_type]()

    def forward(self, h_enc, h_dec):
        # joint_network.py:48-49: h_enc (B,T,1,De), h_dec (B,1,U,Dd)
        return self.lin_out(self.joint_activation(self.lin_enc(h_enc) + self.lin_dec(h_dec)))
