// Banded (streaming) self-attention with learnable relative positions: the attention core of
// RelLearnableMultiHeadAttn.forward (/root/reference/tt/transformer.py:106-177) when the mask is the streaming context
// mask of tt/utils.py:242-251 (left 10 / right 2 frames in config/joint_streaming.yaml's use): a query attends to
// S = left + right + 1 keys, so the dense T x T score tensors (AC, B_, the _rel_shift copy, the masked softmax, T x T x B x n
// each) collapse to T x S.  Same arithmetic as the reference, fp32, INCLUDING what _rel_shift (transformer.py:82-95) does
// to the keys right of the diagonal, which a causal model never looks at but a right context of 2 does:
//   key j <= i     BD = q_i . r_emb'[T-1+j-i] + r_bias'[T-1+j-i]
//   key j = i + 1  BD = 0                                      (the zero column the shift pads with)
//   key j >= i + 2 BD = q_{i+1} . r_emb'[j-i-2] + r_bias'[j-i-2]   (the NEXT query's row, wrapped around)
// with r_emb' = the last T rows of r_emb, or r_emb padded in front with copies of its row 0 when T > max_len
// (transformer.py:130-137).  score = (AC + BD) * scale, AC = (q_i + r_w_bias) . k_j; softmax over the band; out = P . V.
//
// Layouts are the module's own: w_heads (T, B, 3 * n_head * d_head) = [q | k | v] from qkv_net, output (T, B, n_head * d_head).
// One warp per (i, b, head): lanes over d_head, a slot's score stays on lane s.  The backward is three gather passes (no
// atomics on activations): ds from the stored probabilities; dq / dk / dv per position; one accumulation pass for the
// position tables and biases (atomics on max_len x n_head x d_head entries only).
//
// mode 1 = the espnet side, RelPositionMultiHeadedAttention (/root/reference/espnet/nets/pytorch_backend/transformer/
// attention.py:212-308) under the context mask of nets_utils.py:268-281 AND the padding mask (encoder: espnet2/asr/encoder/
// transformer_encoder.py:205-210): the position table has 2T - 1 rows (relative positions T-1 ... -(T-1)), its rel_shift
// has no wrap -- key j uses row T - 1 + j - i with the query's own vector --, and keys at or beyond key_lens[b] are masked
// (a query with no key left gets zeros, like softmax(...).masked_fill(mask, 0) there).  The caller folds pos_bias_v into the
// bias table: (q + v) . p = q . p + (v . p).
#include "ttx_common.cuh"

namespace ttx {

constexpr int kAttnMaxSlots = 32;
constexpr int kAttnMaxDL = 4;          // d_head <= 128

struct AttnParams {
    int T, B, NH, D, L, R, S, maxlen;
    int mode;              // 0: tt (_rel_shift with its wrap), 1: espnet (2T - 1 rows, no wrap)
    const int* klen;       // (B) keys per batch entry, or null: all T
    float scale;
    const float* wh;       // (T, B, 3 * NH * D)
    const float* remb;     // (maxlen, NH, D)
    const float* rwb;      // (NH, D)
    const float* rbias;    // (maxlen, NH)
    float* prob;           // (T, B, NH, S)
    float* out;            // (T, B, NH * D)
    const float* dout;     // (T, B, NH * D)
    float* ds;             // (T, B, NH, S)
    float* dq_ac;          // (T, B, NH, D): the content part of dq (its sum over T, B is d r_w_bias)
    float* dwh;            // (T, B, 3 * NH * D)
    float* d_remb;
    float* d_rwb;
    float* d_rbias;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// row of the (possibly front-padded) position table that slot offset delta = j - i uses, or -1 when the slot has no
// position term (delta = 1); `next` = the term is taken with the NEXT query's vector
__device__ __forceinline__ int rel_row(const AttnParams& p, int delta, bool& next) {
    if (p.mode == 1) {
        next = false;
        return (p.maxlen - 1) / 2 + delta;
    }
    next = delta >= 2;
    if (delta == 1) return -1;
    const int x = delta <= 0 ? p.T - 1 + delta : delta - 2;
    return max(x + p.maxlen - p.T, 0);
}

__global__ void __launch_bounds__(128) band_attn_fwd_kernel(const AttnParams p) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= p.T * p.B * p.NH) return;
    const int n = w % p.NH, b = (w / p.NH) % p.B, i = w / (p.NH * p.B);
    const int HD = p.NH * p.D, ld = 3 * HD, DL = p.D >> 5;
    const float* qrow = p.wh + ((size_t)i * p.B + b) * ld + n * p.D;
    float qi[kAttnMaxDL], qn[kAttnMaxDL], rw[kAttnMaxDL];
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c) {
        const int e = lane + 32 * c;
        qi[c] = qn[c] = rw[c] = 0.f;
        if (c < DL) {
            qi[c] = qrow[e];
            if (i + 1 < p.T) qn[c] = qrow[(size_t)p.B * ld + e];
            rw[c] = p.rwb[n * p.D + e];
        }
    }
    const int kend = p.klen ? min(p.T, p.klen[b]) : p.T;    // keys [0, kend)
    float my = -INFINITY;                                   // lane s keeps the score of slot s
    for (int s = 0; s < p.S; ++s) {
        const int j = i + s - p.L;
        if (j < 0 || j >= kend) continue;                    // (whole warp alike)
        const float* krow = p.wh + ((size_t)j * p.B + b) * ld + HD + n * p.D;
        bool next;
        const int g = rel_row(p, s - p.L, next);
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < kAttnMaxDL; ++c)
            if (c < DL) {
                const int e = lane + 32 * c;
                part = fmaf(qi[c] + rw[c], krow[e], part);
                if (g >= 0) part = fmaf(next ? qn[c] : qi[c], p.remb[((size_t)g * p.NH + n) * p.D + e], part);
            }
        float sc = warp_sum(part);
        if (g >= 0) sc += p.rbias[(size_t)g * p.NH + n];
        if (lane == s) my = sc * p.scale;
    }
    const float mx = warp_max(my);
    const float ex = (my == -INFINITY) ? 0.f : __expf(my - mx);
    const float den = warp_sum(ex);
    const float pr = den > 0.f ? ex / den : 0.f;            // (no key left: zeros)
    if (lane < p.S) p.prob[(size_t)w * p.S + lane] = pr;
    float acc[kAttnMaxDL];
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c) acc[c] = 0.f;
    for (int s = 0; s < p.S; ++s) {
        const int j = i + s - p.L;
        const float ps = __shfl_sync(0xffffffffu, pr, s);
        if (j < 0 || j >= kend) continue;
        const float* vrow = p.wh + ((size_t)j * p.B + b) * ld + 2 * HD + n * p.D;
#pragma unroll
        for (int c = 0; c < kAttnMaxDL; ++c)
            if (c < DL) acc[c] = fmaf(ps, vrow[lane + 32 * c], acc[c]);
    }
    float* orow = p.out + ((size_t)i * p.B + b) * HD + n * p.D;
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c)
        if (c < DL) orow[lane + 32 * c] = acc[c];
}

// ds[i, s] = p (dp - sum_s p dp) * scale, dp[s] = dout_i . v_j
__global__ void __launch_bounds__(128) band_attn_ds_kernel(const AttnParams p) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= p.T * p.B * p.NH) return;
    const int n = w % p.NH, b = (w / p.NH) % p.B, i = w / (p.NH * p.B);
    const int HD = p.NH * p.D, ld = 3 * HD, DL = p.D >> 5;
    const float* drow = p.dout + ((size_t)i * p.B + b) * HD + n * p.D;
    float d[kAttnMaxDL];
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c) d[c] = (c < DL) ? drow[lane + 32 * c] : 0.f;
    float dp = 0.f;
    for (int s = 0; s < p.S; ++s) {
        const int j = i + s - p.L;
        if (j < 0 || j >= p.T) continue;
        const float* vrow = p.wh + ((size_t)j * p.B + b) * ld + 2 * HD + n * p.D;
        float part = 0.f;
#pragma unroll
        for (int c = 0; c < kAttnMaxDL; ++c)
            if (c < DL) part = fmaf(d[c], vrow[lane + 32 * c], part);
        const float v = warp_sum(part);
        if (lane == s) dp = v;
    }
    const float pr = lane < p.S ? p.prob[(size_t)w * p.S + lane] : 0.f;
    const float tot = warp_sum(pr * dp);
    if (lane < p.S) p.ds[(size_t)w * p.S + lane] = pr * (dp - tot) * p.scale;       // 0 outside the sequence (pr = 0)
}

// dq, dk, dv of position i (gathers), written into the [q | k | v] layout of w_heads' gradient
__global__ void __launch_bounds__(128) band_attn_dqkv_kernel(const AttnParams p) {
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (w >= p.T * p.B * p.NH) return;
    const int n = w % p.NH, b = (w / p.NH) % p.B, i = w / (p.NH * p.B);
    const int HD = p.NH * p.D, ld = 3 * HD, DL = p.D >> 5;
    const size_t wstep = (size_t)p.B * p.NH;                 // warps (rows of prob / ds) per time step
    float ac[kAttnMaxDL], bd[kAttnMaxDL], dk[kAttnMaxDL], dv[kAttnMaxDL], rw[kAttnMaxDL];
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c) {
        ac[c] = bd[c] = dk[c] = dv[c] = 0.f;
        rw[c] = (c < DL) ? p.rwb[n * p.D + lane + 32 * c] : 0.f;
    }
    // ---- dq_i: own row of ds against the keys and the position rows of the keys left of / on the diagonal ...
    const float dsv = lane < p.S ? p.ds[(size_t)w * p.S + lane] : 0.f;
    for (int s = 0; s < p.S; ++s) {
        const int j = i + s - p.L;
        const float x = __shfl_sync(0xffffffffu, dsv, s);
        if (j < 0 || j >= p.T) continue;
        const float* krow = p.wh + ((size_t)j * p.B + b) * ld + HD + n * p.D;
        bool next;
        const int g = rel_row(p, s - p.L, next);
#pragma unroll
        for (int c = 0; c < kAttnMaxDL; ++c)
            if (c < DL) {
                const int e = lane + 32 * c;
                ac[c] = fmaf(x, krow[e], ac[c]);
                if (g >= 0 && !next) bd[c] = fmaf(x, p.remb[((size_t)g * p.NH + n) * p.D + e], bd[c]);
            }
    }
    // ... and the previous query's slots right of the diagonal, whose position term was taken with q_i
    if (p.mode == 0 && i >= 1) {
        const float dsp = lane < p.S ? p.ds[((size_t)w - wstep) * p.S + lane] : 0.f;
        for (int delta = 2; delta <= p.R; ++delta) {
            const float x = __shfl_sync(0xffffffffu, dsp, p.L + delta);
            if (i - 1 + delta >= p.T) continue;
            bool next;
            const int g = rel_row(p, delta, next);
#pragma unroll
            for (int c = 0; c < kAttnMaxDL; ++c)
                if (c < DL) bd[c] = fmaf(x, p.remb[((size_t)g * p.NH + n) * p.D + lane + 32 * c], bd[c]);
        }
    }
    // ---- dk_i, dv_i: every query i2 that has position i in its band
    for (int i2 = max(0, i - p.R); i2 <= min(p.T - 1, i + p.L); ++i2) {
        const int s2 = i - i2 + p.L;
        const size_t w2 = ((size_t)i2 * p.B + b) * p.NH + n;
        const float x = p.ds[w2 * p.S + s2], pv = p.prob[w2 * p.S + s2];
        const float* qrow = p.wh + ((size_t)i2 * p.B + b) * ld + n * p.D;
        const float* drow = p.dout + ((size_t)i2 * p.B + b) * HD + n * p.D;
#pragma unroll
        for (int c = 0; c < kAttnMaxDL; ++c)
            if (c < DL) {
                const int e = lane + 32 * c;
                dk[c] = fmaf(x, qrow[e] + rw[c], dk[c]);
                dv[c] = fmaf(pv, drow[e], dv[c]);
            }
    }
    float* grow = p.dwh + ((size_t)i * p.B + b) * ld + n * p.D;
    float* arow = p.dq_ac + (size_t)w * p.D;
#pragma unroll
    for (int c = 0; c < kAttnMaxDL; ++c)
        if (c < DL) {
            const int e = lane + 32 * c;
            grow[e] = ac[c] + bd[c];
            grow[HD + e] = dk[c];
            grow[2 * HD + e] = dv[c];
            arow[e] = ac[c];
        }
}

// d r_emb, d r_bias, d r_w_bias: grid = (NH, row chunks), block = D threads; thread e accumulates its column over the
// chunk's (i, b) rows, then one atomic per (slot, e)
__global__ void band_attn_dparam_kernel(const AttnParams p, int rows_per_block) {
    const int n = blockIdx.x, e = threadIdx.x;
    const int HD = p.NH * p.D, ld = 3 * HD;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(p.T * p.B, r0 + rows_per_block);
    float acc[kAttnMaxSlots], bacc[kAttnMaxSlots];
#pragma unroll
    for (int s = 0; s < kAttnMaxSlots; ++s) acc[s] = bacc[s] = 0.f;
    float racc = 0.f;
    for (int r = r0; r < r1; ++r) {
        const int i = r / p.B;
        const float* qrow = p.wh + (size_t)r * ld + n * p.D;
        const float qi = qrow[e];
        const float qn = (i + 1 < p.T) ? qrow[(size_t)p.B * ld + e] : 0.f;
        const float* dsr = p.ds + ((size_t)r * p.NH + n) * p.S;
        racc += p.dq_ac[((size_t)r * p.NH + n) * p.D + e];
#pragma unroll
        for (int s = 0; s < kAttnMaxSlots; ++s)
            if (s < p.S) {
                const float x = dsr[s];                      // 0 for slots outside the sequence
                acc[s] = fmaf(x, (p.mode == 0 && s - p.L >= 2) ? qn : qi, acc[s]);
                bacc[s] += x;
            }
    }
    atomicAdd(p.d_rwb + n * p.D + e, racc);
#pragma unroll
    for (int s = 0; s < kAttnMaxSlots; ++s)
        if (s < p.S) {
            bool next;
            const int g = rel_row(p, s - p.L, next);
            if (g >= 0) {
                atomicAdd(p.d_remb + ((size_t)g * p.NH + n) * p.D + e, acc[s]);
                if (e == 0) atomicAdd(p.d_rbias + (size_t)g * p.NH + n, bacc[s]);
            }
        }
}

static AttnParams attn_params(const float* wh, const float* remb, const float* rwb, const float* rbias, int T, int B, int NH,
                              int D, int maxlen, int L, int R, float scale, int mode, const int* klen) {
    AttnParams p{};
    p.T = T; p.B = B; p.NH = NH; p.D = D; p.L = L; p.R = R; p.S = L + R + 1; p.maxlen = maxlen;
    p.mode = mode;
    p.klen = klen;
    p.scale = scale;
    p.wh = wh; p.remb = remb; p.rwb = rwb; p.rbias = rbias;
    return p;
}

int launch_band_attn_fwd(const float* wh, const float* remb, const float* rwb, const float* rbias, int T, int B, int NH, int D,
                         int maxlen, int L, int R, float scale, int mode, const int* klen, float* prob, float* out,
                         cudaStream_t s) {
    AttnParams p = attn_params(wh, remb, rwb, rbias, T, B, NH, D, maxlen, L, R, scale, mode, klen);
    p.prob = prob;
    p.out = out;
    const long long warps = (long long)T * B * NH;
    band_attn_fwd_kernel<<<(unsigned)((warps + 3) / 4), 128, 0, s>>>(p);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_band_attn_bwd(const float* wh, const float* remb, const float* rwb, const float* prob, const float* dout, int T,
                         int B, int NH, int D, int maxlen, int L, int R, float scale, int mode, const int* klen, float* ds,
                         float* dq_ac, float* dwh, float* d_remb, float* d_rwb, float* d_rbias, cudaStream_t s) {
    AttnParams p = attn_params(wh, remb, rwb, nullptr, T, B, NH, D, maxlen, L, R, scale, mode, klen);
    p.prob = const_cast<float*>(prob);
    p.dout = dout;
    p.ds = ds;
    p.dq_ac = dq_ac;
    p.dwh = dwh;
    p.d_remb = d_remb;
    p.d_rwb = d_rwb;
    p.d_rbias = d_rbias;
    const long long warps = (long long)T * B * NH;
    const unsigned grid = (unsigned)((warps + 3) / 4);
    band_attn_ds_kernel<<<grid, 128, 0, s>>>(p);
    band_attn_dqkv_kernel<<<grid, 128, 0, s>>>(p);
    const int rows = T * B, rpb = max(32, (rows + 255) / 256);
    band_attn_dparam_kernel<<<dim3(NH, (rows + rpb - 1) / rpb), D, 0, s>>>(p, rpb);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ttx
