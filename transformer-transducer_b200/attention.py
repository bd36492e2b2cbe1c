"""Banded (streaming) relative-position self-attention for the tt and espnet encoders.

Drop-in for ``RelLearnableMultiHeadAttn.forward`` (/root/reference/tt/transformer.py:106-177) when ``attn_mask`` is the
streaming context mask of ``tt.utils.context_mask`` (tt/utils.py:242-251; left 10 / right 2 by default): every query
attends to ``left + right + 1`` keys, so the attention core runs on a (T, left + right + 1) band (csrc/ttx_attn.cu)
instead of four dense (T, T, B, n_head) tensors.  Everything around the core -- ``qkv_net``, ``o_net``, dropout, the
residual LayerNorm -- is the module's own code path; the arithmetic of the core is the reference's, including what
``_rel_shift`` does right of the diagonal.  ``install(patch_attention=True, streaming_context=(10, 2))`` rebinds the
method; any call it cannot take (no mask, another mask, CPU tensors, attention dropout in training, other dtypes) goes
to the reference's own forward.

The espnet side -- ``RelPositionMultiHeadedAttention.forward`` (espnet/nets/pytorch_backend/transformer/attention.py:
264-308) as ``tt_espnet``'s encoder calls it (espnet2/asr/encoder/transformer_encoder.py:205-210: padding mask AND the
context mask of nets_utils.py:268-281) -- runs on the same kernels (``mode 1``: 2T - 1 position rows, no wrap in its
rel_shift, keys beyond the utterance's length masked): ``espnet_banded_forward``.
"""
import ctypes
import weakref

import torch

from . import _lib

CONTEXT = (10, 2)            # (left, right) of the masks the band kernel accepts; set by install(streaming_context=...)
_verified = {}               # id(mask) -> (weakref to the mask, bool): one check per mask object (= per forward pass)


def _p(t):
    return t.data_ptr() if t is not None else None


def _is_context_band(attn_mask, T, left, right):
    """True when attn_mask -- (T, T, 1), nonzero = masked: tt/utils.py:242-251's mask as tt/model.py:60 passes it (a 2-D
    mask means (key, batch) padding in transformer.py:156-157) -- masks exactly the keys outside [i - left, i + right].
    One device comparison + host read per mask OBJECT: the encoder hands the same tensor to every layer
    (tt/encoder.py:48-49)."""
    key = id(attn_mask)
    hit = _verified.get(key)
    if hit is not None and hit[0]() is attn_mask:
        return hit[1]
    ok = False
    if attn_mask.dim() == 3 and attn_mask.size(0) == T and attn_mask.size(1) == T and attn_mask.size(2) == 1:
        m = attn_mask.reshape(T, T).bool()
        idx = torch.arange(T, device=m.device)
        delta = idx[None, :] - idx[:, None]                      # j - i
        ok = bool(torch.equal(m, (delta > right) | (delta < -left)))
    if len(_verified) > 64:
        _verified.clear()
    _verified[key] = (weakref.ref(attn_mask), ok)
    return ok


class BandAttnCore(torch.autograd.Function):
    """attn_vec (T, B, n_head * d_head) from w_heads (T, B, 3 * n_head * d_head) and the position tables."""

    @staticmethod
    def forward(ctx, w_heads, r_emb, r_w_bias, r_bias, n_head, d_head, left, right, scale, mode=0, key_lens=None):
        lib = _lib.get()
        dev = w_heads.device
        T, B = w_heads.shape[0], w_heads.shape[1]
        wh = w_heads.detach().contiguous()
        re, rw, rb = r_emb.detach().contiguous(), r_w_bias.detach().contiguous(), r_bias.detach().contiguous()
        S = left + right + 1
        prob = torch.empty(T, B, n_head, S, dtype=torch.float32, device=dev)
        out = torch.empty(T, B, n_head * d_head, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            st = torch._C._cuda_getCurrentRawStream(idx)
            _lib.check(lib.ttx_band_attn_fwd(_p(wh), _p(re), _p(rw), _p(rb), T, B, n_head, d_head, re.shape[0], left, right,
                                             ctypes.c_float(scale), mode, _p(key_lens), _p(prob), _p(out), idx, st),
                       "ttx_band_attn_fwd")
        ctx.save_for_backward(wh, re, rw, prob)
        ctx.cfg = (n_head, d_head, left, right, scale, mode, key_lens)
        return out

    @staticmethod
    def backward(ctx, d_out):
        wh, re, rw, prob = ctx.saved_tensors
        n_head, d_head, left, right, scale, mode, key_lens = ctx.cfg
        lib = _lib.get()
        dev = wh.device
        T, B = wh.shape[0], wh.shape[1]
        d_out = d_out.contiguous()
        ds = torch.empty_like(prob)
        dq_ac = torch.empty(T, B, n_head * d_head, dtype=torch.float32, device=dev)
        d_wh = torch.empty_like(wh)
        d_re, d_rw = torch.zeros_like(re), torch.zeros_like(rw)
        d_rb = torch.zeros(re.shape[0], n_head, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            st = torch._C._cuda_getCurrentRawStream(idx)
            _lib.check(lib.ttx_band_attn_bwd(_p(wh), _p(re), _p(rw), _p(prob), _p(d_out), T, B, n_head, d_head, re.shape[0],
                                             left, right, ctypes.c_float(scale), mode, _p(key_lens), _p(ds), _p(dq_ac),
                                             _p(d_wh), _p(d_re), _p(d_rw), _p(d_rb), idx, st), "ttx_band_attn_bwd")
        return d_wh, d_re, d_rw, d_rb, None, None, None, None, None, None, None


def banded_forward(self, w, r_emb, r_w_bias, r_bias, attn_mask=None):
    """RelLearnableMultiHeadAttn.forward with the attention core on the band kernel (see the module docstring)."""
    left, right = CONTEXT
    reference = type(self)._ttb_reference_forward
    if (attn_mask is None or not w.is_cuda or w.dtype != torch.float32 or r_emb.dtype != torch.float32 or
            self.d_head % 32 != 0 or self.d_head > 128 or left + right + 1 > 32 or
            (self.training and self.dropatt.p > 0) or r_bias.shape[0] != r_emb.shape[0] or
            not _is_context_band(attn_mask, w.size(0), left, right)):
        return reference(self, w, r_emb, r_w_bias, r_bias, attn_mask)
    w_heads = self.qkv_net(w)                                              # transformer.py:115
    attn_vec = BandAttnCore.apply(w_heads, r_emb, r_w_bias, r_bias, self.n_head, self.d_head, left, right,
                                  float(self.scale))
    attn_out = self.drop(self.o_net(attn_vec))                             # transformer.py:170-171
    return self.layer_norm(w + attn_out)                                   # transformer.py:173


# ------------------------------------------------------------------------------------------------ espnet side
def _espnet_band(mask, T):
    """mask (1 | B, T, T), nonzero = attend (transformer_encoder.py:205-210: padding mask & ~make_attention_mask).  When
    it is exactly `key within [i - left, i + right] AND key < length[b]` returns (left, right, lengths int32 on the
    device), else None.  left / right are read off the longest entry's first and last rows (the encoder's band is
    (10, 2), the label encoder's (2, 0), tt_espnet/model.py:56-66); two small host reads + one device comparison per mask
    OBJECT -- every layer gets the same tensor."""
    key = id(mask)
    hit = _verified.get(key)
    if hit is not None and hit[0]() is mask:
        return hit[1]
    found = None
    if mask.dim() == 3 and mask.size(1) == T and mask.size(2) == T:
        m = mask.bool()
        n_keys = m.any(dim=1).sum(dim=1)                                     # (1 | B): key j < length is seen by query j
        longest = int(torch.argmax(n_keys))
        n_max, first_row = (int(x) for x in torch.stack([n_keys[longest], m[longest, 0].sum()]).tolist())
        if n_max >= 1:
            last_row = int(m[longest, n_max - 1].sum())
            right, left = first_row - 1, last_row - 1
            idx = torch.arange(T, device=m.device)
            delta = idx[None, :] - idx[:, None]                              # j - i
            want = ((delta <= right) & (delta >= -left))[None] & (idx[None, None, :] < n_keys[:, None, None])
            if right >= 0 and left >= 0 and bool(torch.equal(m, want)):
                found = (left, right, n_keys.to(torch.int32).contiguous())
    if len(_verified) > 64:
        _verified.clear()
    _verified[key] = (weakref.ref(mask), found)
    return found


def espnet_banded_forward(self, query, key, value, pos_emb, mask):
    """RelPositionMultiHeadedAttention.forward (attention.py:264-308) with the attention core on the band kernels."""
    reference = type(self)._ttb_reference_forward
    T = query.size(1)
    if (mask is None or not query.is_cuda or query.dtype != torch.float32 or key is not query or value is not query or
            self.d_k % 32 != 0 or self.d_k > 128 or self.zero_triu or (self.training and self.dropout.p > 0) or
            pos_emb.size(0) != 1 or pos_emb.size(1) != 2 * T - 1):
        return reference(self, query, key, value, pos_emb, mask)
    band = _espnet_band(mask, T)
    if band is None or band[0] + band[1] + 1 > 32:
        return reference(self, query, key, value, pos_emb, mask)
    left, right, lens = band
    B = query.size(0)
    if lens.numel() == 1 and B > 1:
        lens = lens.expand(B).contiguous()
    # [q | k | v] in the kernels' (T, B, 3 * h * d_k) layout (attention.py:45-62 without the head transposes)
    w_heads = torch.cat([self.linear_q(query), self.linear_k(key), self.linear_v(value)], dim=-1).transpose(0, 1)
    p = self.linear_pos(pos_emb).view(2 * T - 1, self.h, self.d_k)           # attention.py:283-284
    v_dot_p = torch.einsum("rhd,hd->rh", p, self.pos_bias_v)                 # (q + v) . p = q . p + v . p   (:289,:299)
    attn_vec = BandAttnCore.apply(w_heads, p, self.pos_bias_u, v_dot_p, self.h, self.d_k, left, right,
                                  1.0 / float(self.d_k) ** 0.5, 1, lens)
    self.attn = None                                                         # (the dense probabilities are never formed)
    return self.linear_out(attn_vec.transpose(0, 1))                         # attention.py:96
