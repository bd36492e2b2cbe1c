"""Greedy transducer search with the decode-time joint on the GPU (csrc/ttx_decode.cu).

Drop-in for the reference's per-utterance search loops
  tt.model.Transducer.decode                      /root/reference/tt/model.py:70-90
  tt.model.Transducer.beam_search                 /root/reference/tt/model.py:110-179
  tt_espnet.model.TransformerTransducer.decode    /root/reference/tt_espnet/model.py:83-106
(same arguments, same return value: the label sequence without the start symbol), the two `recognize` methods
(tt/model.py:92-108, tt_espnet/model.py:108-121) as one batched search, and `StreamingGreedy` for the
streaming demo's window-by-window loop (/root/reference/audio/streamRec_unlimit_dynamic_window.py:186-211).  The reference evaluates
``joint(enc_state[t].view(-1), dec_state.view(-1)) -> softmax -> argmax -> .item()`` once per frame; the decoder state
only changes when a label is emitted, so here every frame up to the next non-blank prediction is scored against the
current decoder state in one launch group (64 frames at a time) and the host reads two integers per emitted label.
The decoder (prediction network) is the model's own module, called exactly as the reference calls it.

Arithmetic: the first joint layer is split algebraically like in training (encoder half once per utterance, decoder
half once per emitted label, both fp32 ``torch.nn.functional.linear``), tanh, output layer and argmax are fp32 in the
kernel -- decoding compares logits, so none of the 16-bit tensor-core operands of the training path are used here.
Equal to the reference up to fp32 summation order; ``install()`` rebinds the two methods.
"""

import torch

from . import _lib
from .joint import JointNet, JointNetwork

SCAN_FRAMES = 64        # frames per launch group at most (the kernel's tile)
SCAN_FIRST = 8          # frames scored right after a label; doubled while only blanks come back


def _p(t):
    return t.data_ptr()


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


class _FrameScanner:
    """One utterance's decode-time joint: encoder half of the first layer for all frames (once), decoder half per label
    history, and the launch group that scores a run of frames against one decoder state."""

    def __init__(self, joint, enc_state, length, parts=None, eproj=None, out=None):
        if not enc_state.is_cuda:
            raise RuntimeError("the decode-time joint kernel needs CUDA tensors (there is no CPU fallback)")
        if length > enc_state.shape[0]:           # the reference's loop would index enc_state[t] past its end
            raise IndexError("decode length %d exceeds the %d encoder frames" % (length, enc_state.shape[0]))
        self.lib = _lib.get()
        self.dev = dev = enc_state.device
        self.length = length
        self.parts = parts = parts if parts is not None else _JointParts(joint, enc_state.size(-1))
        self.H, self.V = parts.w_out.shape[1], parts.w_out.shape[0]
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        if eproj is None:
            eproj = torch.nn.functional.linear(enc_state[:length].float(), parts.w_enc, parts.b_enc).contiguous()
        self.eproj = eproj                                  # (>= length, H), rows contiguous
        self.scratch = torch.empty(SCAN_FRAMES, dtype=torch.int64, device=dev)
        self.out = out if out is not None else torch.empty(2 + SCAN_FRAMES, dtype=torch.int32, device=dev)

    def decoder_half(self, dec_out):
        return torch.nn.functional.linear(dec_out.reshape(-1).float(), self.parts.w_dec)

    def launch(self, t, n, pvec, blank):
        """Frames t .. t + n - 1 against pvec, stream-ordered: out[0] = offset of the first frame whose argmax is not
        blank (n if none), out[1] = its label."""
        st = torch._C._cuda_getCurrentRawStream(self.idx)
        _lib.check(self.lib.ttx_decode_scan(_p(self.eproj[t]), self.eproj.stride(0), _p(pvec), _p(self.parts.w_out),
                                            _p(self.parts.b_out), n, self.H, self.V, int(blank), _p(self.scratch),
                                            _p(self.out), self.idx, st), "ttx_decode_scan")

    def scan(self, t, n, pvec, blank):
        """launch() + the one host read per emitted label / run of blank frames: (offset, label)."""
        self.launch(t, n, pvec, blank)
        first, label = self.out[:2].tolist()
        return first, label

    def next_label(self, t, pvec, blank):
        """First frame >= t whose prediction against pvec is a label, with doubling look-ahead: (frame, label), or
        (length, blank) when only blanks are left."""
        look = SCAN_FIRST
        while t < self.length:
            n = min(look, self.length - t)
            first, label = self.scan(t, n, pvec, blank)
            if first < n:
                return t + first, int(label)
            t += n
            look = min(SCAN_FRAMES, 2 * look)           # a run of blanks: look further ahead next time
        return self.length, int(blank)

    def decoder_halves(self, dec_outs):
        return torch.nn.functional.linear(dec_outs.float(), self.parts.w_dec).contiguous()        # (n, H)

    def posteriors(self, t, pvecs):
        """softmax over the vocabulary at frame t (fp32) for n decoder states (n, H) -> (n, V), for the searches that
        rank more than the best label."""
        hidden = torch.tanh(self.eproj[t] + pvecs)
        return torch.softmax(torch.nn.functional.linear(hidden, self.parts.w_out, self.parts.b_out), dim=-1)


@torch.no_grad()
def greedy_search(joint, enc_state, length, step_decoder, start_token=0, blank=0):
    """enc_state (T, D_enc) CUDA tensor of one utterance, length = frames to decode, step_decoder(token_list) -> the
    decoder's last output (D_dec) for that label history.  Returns the emitted labels (start symbol excluded)."""
    if not enc_state.is_cuda:
        raise RuntimeError("greedy_search needs CUDA tensors (there is no CPU fallback)")
    length = int(length)
    tokens = [int(start_token)]
    if length <= 0:
        return tokens[1:]
    with torch.cuda.device(enc_state.device):
        scanner = _FrameScanner(joint, enc_state, length)
        t = 0
        while t < length:
            pvec = scanner.decoder_half(step_decoder(tokens))
            t, label = scanner.next_label(t, pvec, blank)
            if t >= length:
                break
            tokens.append(label)
            t += 1
    return tokens[1:]


@torch.no_grad()
def greedy_search_batch(joint, enc_states, lengths, step_decoder, start_token=0, blank=0):
    """Greedy search over a whole batch: enc_states (B, T, D_enc) CUDA, lengths[b] frames each, step_decoder(list of
    EQUAL-LENGTH label histories) -> (n, D_dec) last decoder outputs.  Returns B label lists, the ones the reference's
    per-utterance loop (`recognize`: tt/model.py:92-108, tt_espnet/model.py:108-121) produces.

    Round r: every utterance still running has emitted exactly r labels, so their r + 1 long histories go through the
    decoder as ONE batch; then each utterance's frames are scanned from its own position to its next label (launches
    only; one host read per round for all of them, plus one per extra look-ahead group on long runs of blanks)."""
    if not enc_states.is_cuda:
        raise RuntimeError("greedy_search_batch needs CUDA tensors (there is no CPU fallback)")
    B = int(enc_states.shape[0])
    lengths = [int(x) for x in lengths]
    tokens = [[int(start_token)] for _ in range(B)]
    active = [b for b in range(B) if lengths[b] > 0]
    if not active:
        return [t[1:] for t in tokens]
    dev = enc_states.device
    with torch.cuda.device(dev):
        parts = _JointParts(joint, enc_states.size(-1))
        eproj = torch.nn.functional.linear(enc_states.float(), parts.w_enc, parts.b_enc).contiguous()      # (B, T, H)
        outs = torch.zeros(B, 2 + SCAN_FRAMES, dtype=torch.int32, device=dev)
        scanners = [_FrameScanner(joint, enc_states[b], lengths[b], parts=parts, eproj=eproj[b], out=outs[b])
                    for b in range(B)]
        pos = [0] * B
        while active:
            pvecs = scanners[0].decoder_halves(step_decoder([tokens[b] for b in active]))
            row = {b: i for i, b in enumerate(active)}
            look = {b: SCAN_FIRST for b in active}
            pending = list(active)
            while pending:
                count = {}
                for b in pending:
                    count[b] = min(look[b], lengths[b] - pos[b])
                    scanners[b].launch(pos[b], count[b], pvecs[row[b]], blank)
                found = outs[:, :2].tolist()                                  # the round's host read
                still = []
                for b in pending:
                    first, label = found[b]
                    if first < count[b]:
                        tokens[b].append(int(label))
                        pos[b] += first + 1
                        continue
                    pos[b] += count[b]
                    look[b] = min(SCAN_FRAMES, 2 * look[b])
                    if pos[b] < lengths[b]:
                        still.append(b)
                pending = still
            active = [b for b in active if pos[b] < lengths[b]]
    return [t[1:] for t in tokens]


@torch.no_grad()
def beam_search(joint, enc_state, length, step_decoder, beam_width=5, start_token=0, blank=0):
    """The reference's beam search (tt/model.py:110-179), same bookkeeping and therefore the same result: the hypothesis
    with the best score leads; frames on which the leader predicts the blank change nothing, so they are skipped by the
    frame scan; on a label frame every hypothesis is expanded by its `beam_width` best non-blank labels (scores = log
    posteriors, accumulated per parent) and the best `beam_width` children survive.  The child lists are the
    reference's: one (beam_width x beam_width) table that is appended to at every label frame and never reset, the
    first label frame filling it column-wise.

    step_decoder(list of label histories) -> (n, D_dec) last decoder outputs.  All hypotheses have the same length (the
    child lists grow together), so the reference's 1 + beam_width decoder runs and beam_width joint / top-k / host reads
    per label frame are ONE batched decoder run, one posterior GEMM and one host read here."""
    import heapq
    import numpy as np
    length = int(length)
    W = int(beam_width)
    hyps = [[int(start_token)] for _ in range(W)]
    score = np.zeros((W,), dtype=float)
    child = [[[int(start_token)] for _ in range(W)] for _ in range(W)]
    child_score = np.zeros((W, W), dtype=float)
    if length <= 0:
        return hyps[0][1:]
    with torch.cuda.device(enc_state.device):
        scanner = _FrameScanner(joint, enc_state, length)
        fresh = True
        t = 0
        while t < length:
            pvecs = scanner.decoder_halves(step_decoder(hyps))               # (W, H): valid until the next label frame
            t, _ = scanner.next_label(t, pvecs[int(score.argmax())], blank)
            if t >= length:
                break
            top = torch.topk(scanner.posteriors(t, pvecs), k=W + 1, dim=-1)
            both = torch.cat([top.values, top.indices.to(top.values.dtype)], dim=1).tolist()     # labels < 2^24: exact
            for k in range(W):
                values, labels = both[k][:W + 1], [int(x) for x in both[k][W + 1:]]
                drop = labels.index(blank) if blank in labels else W        # the blank if it made the list, else the last
                del values[drop], labels[drop]
                logs = np.log(values)
                if fresh:
                    for i, lab in enumerate(labels):
                        child[i][k].append(lab)
                    child_score[:, k] = logs
                else:
                    for i, lab in enumerate(labels):
                        child[k][i].append(lab)
                    child_score[k] = score[k] + logs
            if fresh:
                fresh = False
                for i in range(W):
                    hyps[i] = list(child[i][0])
                    score[i] = child_score[i, 0]
            else:
                for i, flat in enumerate(heapq.nlargest(W, range(W * W), child_score.take)):
                    score[i] = child_score[flat // W, flat % W]
                    hyps[i] = list(child[flat // W][flat % W])
            t += 1
    return hyps[int(score.argmax())][1:]


class StreamingGreedy:
    """The streaming demo's recognition loop (audio/streamRec_unlimit_dynamic_window.py:113-115,186-211) without its
    per-frame host read: encoder states arrive window by window, the decoder state, the label history (the decoder sees
    the last `history` labels, :201-207) and the count of blank frames since the last label (the demo starts a new line
    at >= 15, :193) carry over from one window to the next."""

    def __init__(self, joint, decoder, history=40, start_token=0, blank=0):
        self.joint, self.decoder = joint, decoder
        self.history, self.blank, self.start_token = int(history), int(blank), int(start_token)
        self.result = []
        self.blank_frame = 0
        self.dec_state = None
        self._parts = None

    @torch.no_grad()
    def feed(self, enc_states):
        """enc_states (n, D_enc) CUDA: the effective frames of one window.  Returns [(label, blank frames since the
        previous label)] for the labels emitted in it; `self.result` holds everything emitted so far."""
        n = int(enc_states.shape[0])
        events = []
        if n == 0:
            return events
        dev = enc_states.device
        with torch.cuda.device(dev):
            if self.dec_state is None:
                self.dec_state = self.decoder(torch.tensor([[self.start_token]], dtype=torch.long, device=dev))   # :112-113
            if self._parts is None:
                self._parts = _JointParts(self.joint, enc_states.size(-1))        # (weights are read once per stream)
            scanner = _FrameScanner(self.joint, enc_states, n, parts=self._parts)
            t = 0
            while t < n:
                hit, label = scanner.next_label(t, scanner.decoder_half(self.dec_state), self.blank)
                if self.result:
                    self.blank_frame += hit - t                                    # :210-211
                if hit >= n:
                    break
                events.append((label, self.blank_frame))
                self.result.append(label)
                token = torch.tensor([self.result[-self.history:]], dtype=torch.long, device=dev)
                self.dec_state = self.decoder(token)[:, -1, :]                     # :201-207
                self.blank_frame = 0
                t = hit + 1
        return events


def tt_decode(self, enc_state, lengths):
    """tt.model.Transducer.decode (tt/model.py:70-90) with the joint scan on the GPU."""
    if not enc_state.is_cuda:
        return type(self)._ttb_reference_decode(self, enc_state, lengths)
    dev = enc_state.device

    def step(tokens):
        token = torch.tensor([tokens], dtype=torch.long, device=dev)
        return self.decoder(token)[:, -1, :]                     # tt/model.py:75,88: full history, last output

    return greedy_search(self.joint, enc_state, lengths, step, start_token=0, blank=0)


def tt_recognize(self, inputs, inputs_length=None, audio_mask=None):
    """tt.model.Transducer.recognize (tt/model.py:92-108): same encoder call, the per-utterance decode loop replaced by
    the batched search."""
    enc_states = self.encoder(inputs, audio_mask)
    if not enc_states.is_cuda:
        return [self.decode(enc_states[b], inputs_length[b]) for b in range(inputs.size(0))]
    dev = enc_states.device

    def step(histories):
        return self.decoder(torch.tensor(histories, dtype=torch.long, device=dev))[:, -1, :]

    return greedy_search_batch(self.joint, enc_states, [int(x) for x in inputs_length], step, start_token=0, blank=0)


def tt_beam_search(self, enc_state, lengths, beam_width=5):
    """tt.model.Transducer.beam_search (tt/model.py:110-179) with the leader's frame scan on the GPU."""
    if not enc_state.is_cuda:
        return type(self)._ttb_reference_beam_search(self, enc_state, lengths, beam_width)
    dev = enc_state.device

    def step(histories):
        tokens = torch.tensor(histories, dtype=torch.long, device=dev)
        return self.decoder(tokens)[:, -1, :]                    # tt/model.py:135-136,141, all hypotheses in one batch

    return beam_search(self.joint, enc_state, lengths, step, beam_width=beam_width, start_token=0, blank=0)


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


@torch.no_grad()
def espnet_recognize(self, speech, speech_lengths):
    """tt_espnet.model.TransformerTransducer.recognize (tt_espnet/model.py:108-121), batched search."""
    encoder_out, _, _ = self.encoder(speech, speech_lengths, left_mask=self.encoder_left_mask,
                                     right_mask=self.encoder_right_mask)
    if not encoder_out.is_cuda:
        return [self.decode(encoder_out[b], speech_lengths[b]) for b in range(speech.size(0))]
    dev = encoder_out.device
    first = [True]

    def step(histories):
        token = torch.tensor(histories, dtype=torch.long, device=dev)
        if first[0]:                                             # model.py:89-90: the left mask on the first call only
            first[0] = False
            out, _, _ = self.decoder.forward_one_step(token, self.decoder_left_mask)
        else:
            out, _, _ = self.decoder.forward_one_step(token)
        return out[:, -1, :]

    return greedy_search_batch(self.joint, encoder_out, [int(x) for x in speech_lengths], step, start_token=self.sos,
                               blank=0)
