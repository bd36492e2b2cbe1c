"""Greedy transducer search with the decode-time joint on the GPU (csrc/ttx_decode.cu).

Drop-in for the reference's per-utterance greedy loops
  tt.model.Transducer.decode                      /root/reference/tt/model.py:70-90
  tt_espnet.model.TransformerTransducer.decode    /root/reference/tt_espnet/model.py:83-106
(same arguments, same return value: the label sequence without the start symbol).  The reference evaluates
``joint(enc_state[t].view(-1), dec_state.view(-1)) -> softmax -> argmax -> .item()`` once per frame; the decoder state
only changes when a label is emitted, so here every frame up to the next non-blank prediction is scored against the
current decoder state in one launch group (64 frames at a time) and the host reads two integers per emitted label.
The decoder (prediction network) is the model's own module, called exactly as the reference calls it.

Arithmetic: the first joint layer is split algebraically like in training (encoder half once per utterance, decoder
half once per emitted label, both fp32 ``torch.nn.functional.linear``), tanh, output layer and argmax are fp32 in the
kernel -- decoding compares logits, so none of the 16-bit tensor-core operands of the training path are used here.
Equal to the reference up to fp32 summation order; ``install()`` rebinds the two methods.
"""
import ctypes

import torch

from . import _lib
from .joint import JointNet, JointNetwork

SCAN_FRAMES = 64        # frames per launch group at most (the kernel's tile)
SCAN_FIRST = 8          # frames scored right after a label; doubled while only blanks come back


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


class _JointParts:
    """The three linear maps of either joint flavour, as (W_enc, b_enc, W_dec, W_out, b_out) float32 views."""

    def __init__(self, joint, enc_dim):
        if isinstance(joint, JointNetwork) or hasattr(joint, "lin_enc"):
            if getattr(joint, "joint_activation_type", "tanh") != "tanh":
                raise ValueError("decode-time joint kernel: tanh joints only")
            self.w_enc, self.b_enc = joint.lin_enc.weight, joint.lin_enc.bias
            self.w_dec = joint.lin_dec.weight
            out = joint.lin_out
        elif isinstance(joint, JointNet) or hasattr(joint, "forward_layer"):
            w = joint.forward_layer.weight
            self.w_enc, self.b_enc = w[:, :enc_dim], joint.forward_layer.bias
            self.w_dec = w[:, enc_dim:]
            out = joint.project_layer
        else:
            raise TypeError("unknown joint module %r" % type(joint))
        f32 = lambda t: None if t is None else t.detach().float().contiguous()  # noqa: E731  (once per utterance)
        self.w_enc, self.b_enc, self.w_dec = f32(self.w_enc), f32(self.b_enc), f32(self.w_dec)
        self.w_out, self.b_out = f32(out.weight), f32(out.bias)


@torch.no_grad()
def greedy_search(joint, enc_state, length, step_decoder, start_token=0, blank=0):
    """enc_state (T, D_enc) CUDA tensor of one utterance, length = frames to decode, step_decoder(token_list) -> the
    decoder's last output (D_dec) for that label history.  Returns the emitted labels (start symbol excluded)."""
    if not enc_state.is_cuda:
        raise RuntimeError("greedy_search needs CUDA tensors (there is no CPU fallback)")
    lib = _lib.get()
    dev = enc_state.device
    length = int(length)
    parts = _JointParts(joint, enc_state.size(-1))
    H, V = parts.w_out.shape[1], parts.w_out.shape[0]
    tokens = [int(start_token)]
    if length <= 0:
        return tokens[1:]
    with torch.cuda.device(dev):
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        eproj = torch.nn.functional.linear(enc_state[:length].float(), parts.w_enc, parts.b_enc).contiguous()
        scratch = torch.empty(SCAN_FRAMES, dtype=torch.int64, device=dev)
        out = torch.empty(2 + SCAN_FRAMES, dtype=torch.int32, device=dev)
        pvec = None
        t = 0
        look = SCAN_FIRST
        while t < length:
            if pvec is None:
                dec = step_decoder(tokens).reshape(-1).float()
                pvec = torch.nn.functional.linear(dec, parts.w_dec)
            n = min(look, length - t)
            st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            _lib.check(lib.ttx_decode_scan(_p(eproj[t]), H, _p(pvec), _p(parts.w_out), _p(parts.b_out), n, H, V, int(blank),
                                           _p(scratch), _p(out), idx, st), "ttx_decode_scan")
            first, label = out[:2].tolist()                     # the one host read per emitted label / 64 blank frames
            if first >= n:
                t += n
                look = min(SCAN_FRAMES, 2 * look)               # a run of blanks: look further ahead next time
                continue
            tokens.append(int(label))
            pvec = None
            t += first + 1
            look = SCAN_FIRST
    return tokens[1:]


def tt_decode(self, enc_state, lengths):
    """tt.model.Transducer.decode (tt/model.py:70-90) with the joint scan on the GPU."""
    if not enc_state.is_cuda:
        return type(self)._ttb_reference_decode(self, enc_state, lengths)
    dev = enc_state.device

    def step(tokens):
        token = torch.tensor([tokens], dtype=torch.long, device=dev)
        return self.decoder(token)[:, -1, :]                     # tt/model.py:75,88: full history, last output

    return greedy_search(self.joint, enc_state, lengths, step, start_token=0, blank=0)


def espnet_decode(self, enc_state, lengths):
    """tt_espnet.model.TransformerTransducer.decode (tt_espnet/model.py:83-106) with the joint scan on the GPU."""
    if not enc_state.is_cuda:
        return type(self)._ttb_reference_decode(self, enc_state, lengths)
    dev = enc_state.device
    first = [True]

    def step(tokens):
        token = torch.tensor([tokens], dtype=torch.long, device=dev)
        if first[0]:                                             # model.py:89-90 passes the left mask on the first call only
            first[0] = False
            out, _, _ = self.decoder.forward_one_step(token, self.decoder_left_mask)
        else:
            out, _, _ = self.decoder.forward_one_step(token)
        return out[:, -1, :]

    return greedy_search(self.joint, enc_state, lengths, step, start_token=self.sos, blank=0)
