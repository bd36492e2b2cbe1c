// Memory- and latency-bound kernels around the tensor-core projection: tile table, operand casts,
// the broadcast-add + tanh producer, the alpha/beta lattice wavefront, gradient coefficients and the
// reductions back to (B,T,H) / (B,U1,H).  Lattice cells live in a COMPACT row space: utterance b owns
// tiles [tile0[b], tile0[b+1]) of 128 rows; row r of the utterance is cell (t, u) = (r / U1_b, r % U1_b)
// with U1_b = label_len[b] + 1, so ragged batches cost no work on padding.
#include "ttx_common.cuh"
#include <stdlib.h>

namespace ttx {

// ------------------------------------------------------------------------------------------- tile table
__global__ void prep_kernel(const int* __restrict__ act_lens, const int* __restrict__ label_lens, int B, int T,
                            int U1, int n_tiles_ub, int* __restrict__ meta) {
    __shared__ int part[1024];
    __shared__ int lpart[1024];
    __shared__ int bad;
    __shared__ long long cells_sh;
    const int tid = threadIdx.x;
    const int per = (B + blockDim.x - 1) / blockDim.x;
    const int b_lo = min(B, tid * per), b_hi = min(B, b_lo + per);
    if (tid == 0) {
        bad = 0x7fffffff;
        cells_sh = 0;
    }
    __syncthreads();
    int local = 0, llocal = 0;
    long long cells = 0;
    for (int b = b_lo; b < b_hi; ++b) {
        const int t = act_lens[b], u1 = label_lens[b] + 1;
        if (t < 1 || t > T || u1 < 1 || u1 > U1) {
            atomicMin(&bad, b + 1);  // report the first bad utterance
            continue;
        }
        local += (t * u1 + kTile - 1) / kTile;
        llocal += lat_elems(t, u1);
        cells += (long long)t * u1;
    }
    part[tid] = local;
    lpart[tid] = llocal;
    atomicAdd(reinterpret_cast<unsigned long long*>(&cells_sh), (unsigned long long)cells);
    __syncthreads();
    // inclusive Hillis-Steele scans over the per-thread partial tile counts / lattice sizes
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
        const int v = (tid >= off) ? part[tid - off] : 0;
        const int lv = (tid >= off) ? lpart[tid - off] : 0;
        __syncthreads();
        part[tid] += v;
        lpart[tid] += lv;
        __syncthreads();
    }
    int tile = part[tid] - local;  // exclusive prefix
    int lat = lpart[tid] - llocal;
    int* tile0 = meta + kMetaHdr;
    int* tile_b = meta + kMetaHdr + B + 1;
    int* lat0 = meta + kMetaHdr + B + 1 + n_tiles_ub;
    const int total = part[blockDim.x - 1];
    const bool fits = total <= n_tiles_ub;       // the caller's buffers hold n_tiles_ub tiles
    for (int b = b_lo; b < b_hi; ++b) {
        const int t = act_lens[b], u1 = label_lens[b] + 1;
        const bool ok = !(t < 1 || t > T || u1 < 1 || u1 > U1);
        const int nt = ok ? (t * u1 + kTile - 1) / kTile : 0;
        tile0[b] = tile;
        lat0[b] = lat;
        if (fits)
            for (int i = 0; i < nt; ++i) tile_b[tile + i] = b;
        tile += nt;
        lat += ok ? lat_elems(t, u1) : 0;
    }
    for (int i = (fits ? total : 0) + tid; i < n_tiles_ub; i += blockDim.x) tile_b[i] = -1;
    if (tid == 0) {
        const int first_bad = (bad == 0x7fffffff) ? (fits ? 0 : B + 1) : bad;
        tile0[B] = total;
        lat0[B] = lpart[blockDim.x - 1];
        meta[0] = first_bad ? 0 : total;
        meta[1] = first_bad;
        meta[2] = (int)min(cells_sh, (long long)0x7fffffff);
        meta[3] = n_tiles_ub;
    }
}

// ------------------------------------------------------------------------------------------- W_out -> 16 bit
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned int* __restrict__ out_bits) {
    float m = 0.f;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        m = fmaxf(m, fabsf(w[i]));
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));  // non-negative floats order as uints
}

// scal[0] = w_scale (power of two putting max|W| into [256, 512) for fp16, 1 for bf16), scal[1] = 1/w_scale
// bias2[v] = b_out[v] * log2(e) for v < V, -inf for the padding rows (so padded columns drop out of every softmax)
template <bool BF16>
__global__ void cast_w_kernel(const float* __restrict__ w, const float* __restrict__ b_out, int V, int Vpad, int H,
                              const unsigned int* __restrict__ absmax_bits, float* __restrict__ scal,
                              uint16_t* __restrict__ w16, float* __restrict__ bias2, uint16_t* __restrict__ w16t) {
    float ws = 1.f;
    if (!BF16) {
        const float m = __uint_as_float(*absmax_bits);
        if (m > 0.f && m < INFINITY) {
            int e;
            frexpf(m, &e);
            ws = ldexpf(1.f, 9 - e);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        scal[0] = ws;
        scal[1] = 1.f / ws;
    }
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < Vpad; v += gridDim.x * blockDim.x)
        bias2[v] = (v < V) ? b_out[v] * kLog2e : -INFINITY;
    const size_t n = (size_t)Vpad * H / 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t e0 = 2 * i;
        const int row = (int)(e0 / H);
        float a = 0.f, b = 0.f;
        if (row < V) {
            const float2 v = *reinterpret_cast<const float2*>(w + e0);
            a = v.x * ws;
            b = v.y * ws;
        }
        const uint32_t pk = pack16<BF16>(a, b);
        reinterpret_cast<uint32_t*>(w16)[i] = pk;
        if (w16t) {   // W16^T (H, Vpad): 4 MB, written uncoalesced once per step (the matrix lives in L2)
            const int col = (int)(e0 - (size_t)row * H);
            w16t[(size_t)col * Vpad + row] = (uint16_t)(pk & 0xffff);
            w16t[(size_t)(col + 1) * Vpad + row] = (uint16_t)(pk >> 16);
        }
    }
}

// ------------------------------------------------------------------------------------------- A16 = tanh(E + P)
// The joint's broadcast-add + tanh (/root/reference/tt/model.py:22-36 after splitting forward_layer,
// espnet joint_network.py:48) written once per lattice cell as the 16-bit tensor-core operand.
// tanh from one exponential and one reciprocal: (1 - e) / (1 + e), e = exp(-2|x|).  Absolute error ~1e-7 (the result
// is rounded to 16 bits, 2.4e-4 relative, right after); tanhf() costs three times the instructions and this kernel
// is issue-bound.
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = __expf(-2.f * fabsf(x));
    return copysignf(__fdividef(1.f - e, 1.f + e), x);
}

template <bool BF16>
__global__ void joint_act_kernel(const float* __restrict__ eproj, const float* __restrict__ pproj,
                                 const int* __restrict__ labels, const int* __restrict__ act_lens,
                                 const int* __restrict__ label_lens, const int* __restrict__ meta, int B, int T,
                                 int U1, int H, int label_stride, int V, uint16_t* __restrict__ a16,
                                 int* __restrict__ row_label, uint16_t* __restrict__ a16t, size_t rows_total) {
    __shared__ uint32_t tr[kTile][33];           // one 64-column slab of the tile (row-major, 33-word rows), for the transposed copy
    const int tile = blockIdx.x;
    if (tile >= meta[0]) {
        // CTA pairs work on tile pairs: with an odd tile count the partner of the last tile must read zeros
        if (tile == meta[0] && (tile & 1)) {
            uint4* dst = reinterpret_cast<uint4*>(a16 + (size_t)tile * kTile * H);
            for (int idx = threadIdx.x; idx < kTile * H / 8; idx += blockDim.x) dst[idx] = make_uint4(0, 0, 0, 0);
            for (int lr = threadIdx.x; lr < kTile; lr += blockDim.x) row_label[(size_t)tile * kTile + lr] = -1;
            if (a16t) {      // blocked layout [rows / 64][H][64]: the tile's two blocks are one contiguous range
                uint4* dt = reinterpret_cast<uint4*>(a16t + (size_t)tile * kTile * H);
                for (int idx = threadIdx.x; idx < kTile * H / 8; idx += blockDim.x) dt[idx] = make_uint4(0, 0, 0, 0);
            }
        }
        return;
    }
    const int b = meta[kMetaHdr + B + 1 + tile];
    const int r0 = (tile - meta[kMetaHdr + b]) * kTile;
    const int Tb = act_lens[b], U1b = label_lens[b] + 1;
    const int nrows = Tb * U1b;
    const float* eb = eproj + (size_t)b * T * H;
    const float* pb = pproj + (size_t)b * U1 * H;
    // 64-column slabs: 8 vectors of 8 columns per row; the slab is also staged in shared memory and written out
    // transposed (A16^T[h][row], 256 contiguous bytes per joint column) for the gradient pass's K-major operand.
    // 256 threads: thread (row group tid >> 3, vector tid & 7) handles the same four rows lr = (tid >> 3) + 32 k in every
    // slab, so their (t, u) -- an integer division each -- are worked out once (the kernel is issue-bound)
    int eoff[4], poff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = r0 + (threadIdx.x >> 3) + 32 * k;
        const int t = r / U1b, u = r - t * U1b;
        eoff[k] = (r < nrows) ? t * H : -1;
        poff[k] = u * H;
    }
    for (int h0 = 0; h0 < H; h0 += 64) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = threadIdx.x + 256 * k;
            const int lr = idx >> 3, vc = (h0 >> 3) + (idx & 7);
            uint4 out = make_uint4(0, 0, 0, 0);
            if (eoff[k] >= 0) {
                const float4* e4 = reinterpret_cast<const float4*>(eb + eoff[k] + vc * 8);
                const float4* p4 = reinterpret_cast<const float4*>(pb + poff[k] + vc * 8);
                const float4 e0 = __ldg(e4), e1 = __ldg(e4 + 1), p0 = __ldg(p4), p1 = __ldg(p4 + 1);
                out.x = pack16<BF16>(tanh_fast(e0.x + p0.x), tanh_fast(e0.y + p0.y));
                out.y = pack16<BF16>(tanh_fast(e0.z + p0.z), tanh_fast(e0.w + p0.w));
                out.z = pack16<BF16>(tanh_fast(e1.x + p1.x), tanh_fast(e1.y + p1.y));
                out.w = pack16<BF16>(tanh_fast(e1.z + p1.z), tanh_fast(e1.w + p1.w));
            }
            *reinterpret_cast<uint4*>(a16 + ((size_t)tile * kTile + lr) * H + vc * 8) = out;
            if (a16t) {
                // word stores, conflict-free: bank = (row + 4 * (idx & 7) + j) mod 32 over a warp's 4 rows x 8 vectors
                uint32_t* dst = &tr[lr][(idx & 7) * 4];
                dst[0] = out.x; dst[1] = out.y; dst[2] = out.z; dst[3] = out.w;
            }
        }
        if (a16t) {
            __syncthreads();
            for (int idx = threadIdx.x; idx < 64 * (kTile / 2); idx += blockDim.x) {
                const int c = idx / (kTile / 2), r2 = idx % (kTile / 2);
                // column c of rows 2 r2 and 2 r2 + 1: the low or the high halves of two words, one byte permute
                const uint32_t v = __byte_perm(tr[2 * r2][c >> 1], tr[2 * r2 + 1][c >> 1], (c & 1) ? 0x7632 : 0x5410);
                // A16^T is stored in blocks of 64 lattice rows, [rows / 64][H][64]: a tile writes two contiguous 64 KiB ranges
                *reinterpret_cast<uint32_t*>(a16t + (((size_t)tile * 2 + (r2 >> 5)) * H + h0 + c) * 64 + ((2 * r2) & 63)) = v;
            }
            __syncthreads();
        }
    }
    for (int lr = threadIdx.x; lr < kTile; lr += blockDim.x) {
        const int r = r0 + lr;
        int lab = -1;
        if (r < nrows) {
            const int u = r % U1b;
            if (u < U1b - 1) lab = labels[(size_t)b * label_stride + u];
            // a label outside the vocabulary is stored as "none": row_label is what every later kernel gathers W_out rows
            // and scatters gradient rows through (the host-side check raises for such input; this keeps a caller that
            // skipped it inside its buffers)
            if (V > 0 && lab >= V) lab = -1;
        }
        row_label[(size_t)tile * kTile + lr] = lab;
    }
}

// out[c][r] = in[r][c] for a row-major [R x C] matrix of 16-bit values (C % 64 == 0, R % 64 == 0): the K-major
// operand copies (W16^T, A16^T) that the backward pair kernel streams for its G pass.
__global__ void transpose16_kernel(const uint16_t* __restrict__ in, uint16_t* __restrict__ out, int R, int C,
                                   const int* __restrict__ meta_rows) {
    __shared__ uint16_t tile[64][66];
    const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
    if (meta_rows && r0 >= ((meta_rows[0] + 1) & ~1) * kTile) return;   // lattice rows beyond the tile pairs in use
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int r = i / 32, c2 = i % 32;
        const uint32_t v = *reinterpret_cast<const uint32_t*>(in + (size_t)(r0 + r) * C + c0 + 2 * c2);
        tile[r][2 * c2] = (uint16_t)(v & 0xffff);
        tile[r][2 * c2 + 1] = (uint16_t)(v >> 16);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 64 * 32; i += blockDim.x) {
        const int c = i / 32, r2 = i % 32;
        const uint32_t v = (uint32_t)tile[2 * r2][c] | ((uint32_t)tile[2 * r2 + 1][c] << 16);
        if (meta_rows)   // A16^T: blocks of 64 lattice rows, [R / 64][C][64] (r0 is a multiple of 64)
            *reinterpret_cast<uint32_t*>(out + ((size_t)(r0 >> 6) * C + c0 + c) * 64 + 2 * r2) = v;
        else
            *reinterpret_cast<uint32_t*>(out + (size_t)(c0 + c) * R + r0 + 2 * r2) = v;
    }
}

int launch_transpose16(const void* in, void* out, int R, int C, const int* meta_rows, cudaStream_t s) {
    transpose16_kernel<<<dim3(C / 64, R / 64), 256, 0, s>>>((const uint16_t*)in, (uint16_t*)out, R, C, meta_rows);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------- lattice
// alpha / beta wavefront of the transducer loss (warp-transducer semantics; called
// train.py:53).  Everything the wavefront touches is stored DIAGONAL-MAJOR: cell (t, u) of utterance b lives at
// lat0[b] + (t + u) * P_b + u with P_b = U1_b rounded up to a multiple of 4 (ttx_common.cuh: lat_pitch / lat_elems), so
// one anti-diagonal is one contiguous, 16-byte aligned run of memory.  beta is computed on the MIRRORED lattice
// (t', u') = (T-1-t, U1-1-u), on which it is the same recursion as alpha,
//     x(t, u) = logaddexp(x(t-1, u) + inB(t, u), x(t, u-1) + inL(t, u)),   x(0, 0) = init,
// with inB / inL the log-probabilities of the blank / label arc ARRIVING in the cell -- so one code path serves both:
//   * lattice_skew_kernel re-orders the two log-probs per cell that the projection kernels wrote in row order
//     (lp_blank, lp_label -- nothing else of the V-wide distribution is ever read) into the four arc arrays;
//   * lattice_wave_kernel: one CTA per (utterance, direction), lane l owns K <= 2 adjacent columns and carries them in
//     registers; the neighbour column's value crosses lanes with a warp shuffle (and warps, for lattices wider than
//     64 columns, through shared memory behind the one block barrier per diagonal).  The operand diagonals are staged
//     PD - 1 steps ahead in a shared-memory ring by 16-byte cp.async copies (whole diagonals, coalesced), read back as
//     one vector per lane, and alpha / beta leave as contiguous runs of doubles.
// The recursion needs more than float32: alpha/beta reach |ll| ~ (T+U) * log V (thousands), where float32 has only
// ~5e-4 of absolute resolution and exp(alpha + beta - ll) would lose 3 digits (the float32 reference does).  It is
// carried as an unevaluated sum of two floats (hi + lo, ~48 bits -- 2e-11 at 6000; a compensated float32 carrier) and
// stored as float64.  (Measured: the same time per diagonal as a float64 carrier, 0.25 us = ~470 cycles at configs[1]; ncu
// shows ~210 instructions per step spread evenly over a single warp -- the step is issue-bound, not latency-bound.  Staging
// the operand diagonals in groups of 16 instead of one per step did not change it either: 0.153 ms, and 0.62 vs 0.54 ms at
// configs[3].)  The
// increment log(1 + exp(-d)) lies in (0, ln 2] and is evaluated in float32 with ex2 / lg2 (absolute error ~1e-7 per
// cell, a random walk of ~3e-6 over a 1200-step lattice).  "log 0" is the finite sentinel kLatNeg.
constexpr float kLatNeg = -1.0e30f;
struct df32 {
    float hi, lo;
};
__device__ __forceinline__ df32 df_add(df32 a, float b) {            // (hi + lo) + b, error-free two-sum + renormalisation
    const float s = __fadd_rn(a.hi, b);
    const float bb = __fsub_rn(s, a.hi);
    float e = __fadd_rn(__fsub_rn(a.hi, __fsub_rn(s, bb)), __fsub_rn(b, bb));
    e = __fadd_rn(e, a.lo);
    const float hi = __fadd_rn(s, e);
    return {hi, __fsub_rn(e, __fsub_rn(hi, s))};
}
__device__ __forceinline__ df32 log_add(df32 a, df32 b) {
    const bool a_ge = a.hi > b.hi || (a.hi == b.hi && a.lo >= b.lo);
    const df32 mx = a_ge ? a : b, mn = a_ge ? b : a;
    // the difference of the high parts is exact whenever it matters (|d| < 30 between values of equal magnitude)
    const float d = __fadd_rn(__fsub_rn(mn.hi, mx.hi), __fsub_rn(mn.lo, mx.lo));
    const float e = ex2f(d * kLog2e);                          // both "log 0": e = 1, the sum stays at the sentinel
    return df_add(mx, lg2f(1.f + e) * kLn2);
}
__device__ __forceinline__ double df_value(df32 a) { return (double)a.hi + (double)a.lo; }

// lat_ws = four arrays of lat_elems floats: [0] inB, [1] inL of the alpha lattice, [2] inB, [3] inL of the mirrored
// (beta) lattice.  Entries that no arc arrives in (first row / first column) stay unwritten and are never used.
__global__ void lattice_skew_kernel(const float* __restrict__ lpb, const float* __restrict__ lpl,
                                    const int* __restrict__ act_lens, const int* __restrict__ label_lens,
                                    const int* __restrict__ meta, int B, size_t lat_elems, float* __restrict__ ws) {
    const int tile = blockIdx.x;
    if (tile >= meta[0]) return;
    const int b = meta[kMetaHdr + B + 1 + tile];
    const int r = (tile - meta[kMetaHdr + b]) * kTile + threadIdx.x;
    const int T = act_lens[b], U1 = label_lens[b] + 1;
    if (r >= T * U1) return;
    const int t = r / U1, u = r - t * U1;
    const int P = lat_pitch(U1);
    const size_t base = (size_t)meta[kMetaHdr + B + 1 + meta[3] + b];
    const size_t g = (size_t)tile * kTile + threadIdx.x;
    const float vb = lpb[g], vl = lpl[g];
    const size_t nxt = base + (size_t)(t + u + 1) * P + u;      // next diagonal, same column: cell (t+1, u)
    if (t + 1 < T) ws[nxt] = vb;                               // blank arc (t, u) -> (t+1, u)
    if (u + 1 < U1) ws[lat_elems + nxt + 1] = vl;              // label arc (t, u) -> (t, u+1)
    const int tm = T - 1 - t, um = U1 - 1 - u;                 // mirrored lattice: both arcs of (t, u) arrive in (tm, um)
    const size_t m = base + (size_t)(tm + um) * P + um;
    ws[2 * lat_elems + m] = vb;
    ws[3 * lat_elems + m] = vl;
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// grid = 2B: block b computes alpha of utterance b, block B + b its beta (on the mirrored lattice, stored mirrored:
// beta(t, u) is element (T-1-t + U1-1-u) * P + U1-1-u of the utterance's run).  W warps x 32 lanes x K columns.
template <int K, int PD, int W>
__global__ void __launch_bounds__(32 * W) lattice_wave_kernel(
        const float* __restrict__ ws, size_t lat_elems, const int* __restrict__ act_lens,
        const int* __restrict__ label_lens, const int* __restrict__ meta, int B, double* __restrict__ alpha_d,
        double* __restrict__ beta_d, float* __restrict__ costs, double* __restrict__ ll_beta) {
    static_assert(PD >= 2 && (K == 1 || K == 2), "ring of at least two diagonals, one or two columns per lane");
    constexpr int PMAX = 32 * K * W;                       // floats per staged diagonal
    __shared__ __align__(16) float stage[PD][2][PMAX];
    __shared__ float2 bnd[2][W + 1];
    if (meta[1] != 0) return;
    const bool is_beta = blockIdx.x >= (unsigned)B;
    const int b = is_beta ? blockIdx.x - B : blockIdx.x;
    const int T = act_lens[b], U1 = label_lens[b] + 1;
    const int P = lat_pitch(U1);
    const size_t base = (size_t)meta[kMetaHdr + B + 1 + meta[3] + b];
    const float* inB = ws + (is_beta ? 2 * lat_elems : 0) + base;
    const float* inL = inB + lat_elems;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u0 = threadIdx.x * K;                        // first column of this lane
    const int nd = T + U1 - 1;
    const int nchunk = P >> 2;                             // 16-byte chunks per diagonal
    auto stage_step = [&](int s, int slot) {               // operands of diagonal s into ring slot `slot`
        if (s < nd) {
            const float* gb = inB + (size_t)s * P;
            const float* gl = inL + (size_t)s * P;
            for (int c = threadIdx.x; c < nchunk; c += 32 * W) {
                cp_async16(smem_u32(&stage[slot][0][4 * c]), gb + 4 * c);
                cp_async16(smem_u32(&stage[slot][1][4 * c]), gl + 4 * c);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int s = 0; s < PD - 1; ++s) stage_step(s, s);
    const df32 neg{kLatNeg, 0.f};
    df32 v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = neg;
    double* out = (is_beta ? beta_d : alpha_d) + base + u0;
    for (int s0 = 0; s0 < nd; s0 += PD) {
#pragma unroll
        for (int i = 0; i < PD; ++i) {
            const int s = s0 + i;                          // diagonal of this step; ring slot i
            if (s >= nd) break;
            if (W > 1 && lane == 31) bnd[i & 1][warp + 1] = make_float2(v[K - 1].hi, v[K - 1].lo);   // boundary column, previous step
            cp_async_wait<PD - 2>();                        // this step's diagonal has landed (own copies) ...
            if (W > 1) __syncthreads();                     // ... everybody's, and the slot refilled below has been read
            else __syncwarp();
            stage_step(s + PD - 1, (i + PD - 1) % PD);
            float sb[K], sl[K];
            if (K == 2) {
                const float2 b2 = *reinterpret_cast<const float2*>(&stage[i][0][u0]);
                const float2 l2 = *reinterpret_cast<const float2*>(&stage[i][1][u0]);
                sb[0] = b2.x; sb[K - 1] = b2.y; sl[0] = l2.x; sl[K - 1] = l2.y;
            } else {
                sb[0] = stage[i][0][u0];
                sl[0] = stage[i][1][u0];
            }
            // left neighbour column's value of the previous step
            df32 left;
            left.hi = __shfl_up_sync(0xffffffffu, v[K - 1].hi, 1);
            left.lo = __shfl_up_sync(0xffffffffu, v[K - 1].lo, 1);
            if (lane == 0) left = neg;
            if (W > 1 && lane == 0 && warp > 0) {
                const float2 bv = bnd[i & 1][warp];
                left = {bv.x, bv.y};
            }
            df32 nv[K];
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int u = u0 + j, t = s - u;
                const bool on = u < U1 && t >= 0 && t < T;
                const df32 from_t = df_add(v[j], t > 0 ? sb[j] : 0.f);               // v = sentinel when there is no cell above
                const df32 from_u = df_add(j > 0 ? v[j > 0 ? j - 1 : 0] : left, u > 0 ? sl[j] : 0.f);
                df32 val = log_add(from_t, from_u);
                if (s == 0) val = {is_beta ? sb[j] : 0.f, 0.f};                      // x(0,0): beta starts from lp_blank(T-1,U)
                val = on ? val : neg;
                if (on) out[(size_t)s * P + j] = df_value(val);
                nv[j] = val;
            }
#pragma unroll
            for (int j = 0; j < K; ++j) v[j] = nv[j];
        }
    }
    cp_async_wait<0>();
    // both lattices end in their last cell (T-1, U1-1), alone on diagonal nd - 1
    const int j_last = U1 - 1 - u0;
    if (j_last >= 0 && j_last < K) {
        const double x = df_value((K == 2 && j_last == 1) ? v[K - 1] : v[0]);
        if (is_beta) ll_beta[b] = x;                                                   // beta(0, 0)
        else costs[b] = (float)-(x + (double)__ldg(ws + 2 * lat_elems + base));       // + lp_blank(T-1, U1-1)
    }
}

// ------------------------------------------------------------------------------------------- gradient coefficients
__global__ void gmax_kernel(const float* __restrict__ grad_costs, int B, float* __restrict__ scal) {
    __shared__ float sm[32];
    __shared__ int neg;
    if (threadIdx.x == 0) neg = 0;
    __syncthreads();
    float m = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        m = fmaxf(m, fabsf(grad_costs[i]));
        if (grad_costs[i] < 0.f) neg = 1;
    }
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)(blockDim.x >> 5); ++i) m = fmaxf(m, sm[i]);
        scal[2] = (m > 0.f && m < INFINITY) ? m : 1.f;
        scal[3] = neg ? 1.f : 0.f;
    }
}

// rowmeta[row] = {lse, p_blank - rb, p_label - rl, w = gamma * g_b / gmax}: rb / rl are the posteriors of leaving
// the cell by a blank / label arc given that the cell is visited, and
//   dL/dz(row, v) = gmax * w * (softmax_v - rb [v == blank] - rl [v == label]).
// .y / .z are the complete bracket at the blank / label column (p from the forward pass: exp(lp)), so the
// gradient kernels write those two entries exactly and keep the dense loop free of per-element compares.
// The sparse part of dL/db_out (-w * rb at blank, -w * rl at the label) is accumulated here as well.
__global__ void grad_prep_kernel(const float* __restrict__ lse, const float* __restrict__ lpb,
                                 const float* __restrict__ lpl, const double* __restrict__ alpha,
                                 const double* __restrict__ beta, const double* __restrict__ ll_beta,
                                 const float* __restrict__ grad_costs, const float* __restrict__ scal,
                                 const int* __restrict__ row_label, const int* __restrict__ act_lens,
                                 const int* __restrict__ label_lens, const int* __restrict__ meta, int B, int blank,
                                 float4* __restrict__ rowmeta, float* __restrict__ d_b_out) {
    __shared__ float blank_sum[kTile / 32];
    const int tile = blockIdx.x;
    if (tile >= meta[0]) return;
    const int b = meta[kMetaHdr + B + 1 + tile];
    const int r = (tile - meta[kMetaHdr + b]) * kTile + threadIdx.x;
    const int T = act_lens[b], U1 = label_lens[b] + 1;
    const size_t base = (size_t)meta[kMetaHdr + b] * kTile;
    float4 out = make_float4(INFINITY, 0.f, 0.f, 0.f);
    float db_blank = 0.f;
    if (r < T * U1) {
        const int t = r / U1, u = r - t * U1;
        const double ll = ll_beta[b];
        // alpha / beta are diagonal-major (see the lattice kernels): (t+1, u) and (t, u+1) sit on the next diagonal
        // beta is stored on the mirrored lattice: (t+1, u) and (t, u+1) sit on its previous diagonal
        const int P = lat_pitch(U1);
        const size_t l0 = (size_t)meta[kMetaHdr + B + 1 + meta[3] + b];
        const size_t lb = l0 + (size_t)(T - 1 - t + U1 - 1 - u) * P + (U1 - 1 - u);
        const double be = beta[lb];
        const float gam = expf((float)(alpha[l0 + (size_t)(t + u) * P + u] + be - ll));
        float rb, rl = 0.f;
        if (t < T - 1) rb = expf((float)((double)lpb[base + r] + beta[lb - P] - be));
        else rb = (u == U1 - 1) ? 1.f : 0.f;
        if (u < U1 - 1) rl = expf((float)((double)lpl[base + r] + beta[lb - P - 1] - be));
        const int lab = row_label[base + r];
        const float w = gam * grad_costs[b] / scal[2];
        float fb = expf(lpb[base + r]) - rb;
        float fl = (lab >= 0) ? expf(lpl[base + r]) - rl : 0.f;
        if (lab == blank) fb = fl = fb - rl;
        out = make_float4(lse[base + r], fb, fl, w);
        if (d_b_out) {
            db_blank = -w * scal[2] * rb;
            if (lab >= 0 && rl != 0.f) atomicAdd(d_b_out + lab, -w * scal[2] * rl);
        }
    }
    rowmeta[(size_t)tile * kTile + threadIdx.x] = out;
    if (d_b_out) {
        for (int o = 16; o; o >>= 1) db_blank += __shfl_xor_sync(0xffffffffu, db_blank, o);
        if ((threadIdx.x & 31) == 0) blank_sum[threadIdx.x >> 5] = db_blank;
        __syncthreads();
        if (threadIdx.x == 0) {
            float sacc = 0.f;
            for (int i = 0; i < kTile / 32; ++i) sacc += blank_sum[i];
            if (sacc != 0.f) atomicAdd(d_b_out + blank, sacc);
        }
    }
}

// ------------------------------------------------------------------------------------------- dA -> dEproj, dPproj
__device__ __forceinline__ float sech2(float x) {
    const float e = __expf(-2.f * fabsf(x));
    const float d = 1.f + e;
    return __fdividef(4.f * e, d * d);      // d in [1, 2]: the fast division is good to 2 ulp here
}

// Source of dL/dA for the two reductions below.  EW = false: `src` already holds dL/dA.  EW = true: `src` holds
// EW = sum_v p_v W_out[v] without the blank / label columns (forward pass, MODE_FG) and
//   dL/dA[row] = gmax * w * (EW[row] + (p_b - rb) W_out[blank] + (p_l - rl) W_out[label])
// with the bracketed coefficients from rowmeta (.y, .z; identical when label == blank) and fp32 W_out rows.
struct ActGradSrc {
    const float* src;
    const float4* rowmeta;
    const int* row_label;
    const float* w_out;
    const float* scal;
    int blank;
};

// Both reductions in ONE pass over the source.  grid = (ceil(T / 8), B, ceil(H / 512)); a block owns 8 frames of one
// utterance and a 512-column slab of h, keeps the frames' Eproj values and dEproj accumulators in registers, walks u, and
// for every u adds the 8 products to both sums: dEproj[b, t, :] is written once at the end, dPproj[b, u, :] gets one
// 16-byte reduction per (block, u).  The source rows (2 KiB each, 8 per u) do not go through registers on their way in:
// a producer warp streams them into a four-stage shared-memory ring with bulk asynchronous copies
// (cp.async.bulk, completion counted on an mbarrier), so ~48 KiB per block are in flight whatever the consumers are
// doing -- with plain loads the kernel ran at the latency of two outstanding loads per thread (2.1 TB/s).
constexpr int kReduceFrames = 8;
constexpr int kReduceStages = 4;
constexpr int kReduceSlab = 512;
constexpr int kReduceThreads = 128 + 32;            // four consumer warps + the producer warp

__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

template <bool EW>
__global__ void __launch_bounds__(kReduceThreads, 4) reduce_both_kernel(const ActGradSrc a, const float* __restrict__ eproj,
                                                                     const float* __restrict__ pproj,
                                                                     const int* __restrict__ act_lens,
                                                                     const int* __restrict__ label_lens,
                                                                     const int* __restrict__ meta, int T, int U1, int H,
                                                                     float* __restrict__ d_eproj, float* __restrict__ d_pproj) {
    constexpr int TT = kReduceFrames;
    const int b = blockIdx.y, t0 = blockIdx.x * TT, h0 = blockIdx.z * kReduceSlab;
    if (meta[1] != 0) return;
    const int slab = min(kReduceSlab, H - h0);
    const int Tb = act_lens[b], U1b = label_lens[b] + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hl = threadIdx.x * 4;                                 // consumer thread's first column inside the slab
    const bool active = warp < 4 && hl < slab;
    const int h = h0 + hl;
    const int nv = min(TT, Tb - t0);                                // frames of the block inside the utterance
    if (nv <= 0) {                                                  // beyond the utterance: exact zeros
        if (active)
            for (int i = 0; i < TT && t0 + i < T; ++i)
                *reinterpret_cast<float4*>(d_eproj + ((size_t)b * T + t0 + i) * H + h) = make_float4(0.f, 0.f, 0.f, 0.f);
        return;
    }
    extern __shared__ __align__(128) uint8_t rsm[];
    const uint32_t row_bytes = slab * 4;
    float* sSrc = reinterpret_cast<float*>(rsm);                                                    // [stage][frame][slab]
    float4* sRm = reinterpret_cast<float4*>(rsm + kReduceStages * TT * kReduceSlab * 4);             // [stage][frame]
    const uint32_t sBar = smem_u32(rsm) + kReduceStages * TT * (kReduceSlab * 4 + 16);
    auto bar_full = [&](int s) { return sBar + 8 * s; };
    auto bar_empty = [&](int s) { return sBar + 8 * (kReduceStages + s); };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kReduceStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 4);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const size_t base = (size_t)meta[kMetaHdr + b] * kTile;
    if (warp == 4) {
        // ----------------------------------------------------------- producer: 8 source rows (+ their coefficients) per u,
        // lane i issues frame i's copies (a single thread issuing all sixteen took as long as the consumers' arithmetic)
        Ring r;
        const char* src_row = reinterpret_cast<const char*>(a.src + (base + (size_t)(t0 + lane) * U1b) * H + h0);
        const float4* rm_row = EW ? a.rowmeta + base + (size_t)(t0 + lane) * U1b : nullptr;
        const size_t src_step = (size_t)H * 4;
        for (int u = 0; u < U1b; ++u) {
            if (lane == 0) {
                mbar_wait(bar_empty(r.stage), r.phase ^ 1);
                mbar_arrive_expect_tx(bar_full(r.stage), nv * (row_bytes + (EW ? 16u : 0u)));
            }
            __syncwarp();
            if (lane < nv) {
                bulk_copy_g2s(smem_u32(sSrc + (r.stage * TT + lane) * kReduceSlab), src_row + u * src_step, row_bytes,
                              bar_full(r.stage));
                if (EW) bulk_copy_g2s(smem_u32(sRm + r.stage * TT + lane), rm_row + u, 16, bar_full(r.stage));
            }
            r.advance(kReduceStages);
        }
        return;
    }
    // --------------------------------------------------------------- consumers: thread = one float4 of h
    // The arithmetic is packed (fma / add / mul .f32x2, two lanes per issue slot): with the loads out of the way the kernel
    // is issue-bound, and what remains per element is the two MUFU operations of sech^2 = 4 e / (1 + e)^2, e = exp(-2|x|).
    const float gscale4 = 4.f * (EW ? a.scal[2] : 1.f);              // (the 4 of sech^2 rides on the row coefficient)
    const float kNeg2Log2e = -2.f * kLog2e;
    uint64_t e01[TT], e23[TT], a01[TT], a23[TT];
    const uint64_t zero2 = pk2(0.f, 0.f), one2 = pk2(1.f, 1.f);
#pragma unroll
    for (int i = 0; i < TT; ++i) {
        a01[i] = a23[i] = zero2;
        float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
        if (active && i < nv) ev = __ldg(reinterpret_cast<const float4*>(eproj + ((size_t)b * T + t0 + i) * H + h));
        e01[i] = pk2(ev.x, ev.y);
        e23[i] = pk2(ev.z, ev.w);
    }
    uint64_t wb01 = zero2, wb23 = zero2;
    if (EW && active) {
        const float4 wb = __ldg(reinterpret_cast<const float4*>(a.w_out + (size_t)a.blank * H + h));
        wb01 = pk2(wb.x, wb.y);
        wb23 = pk2(wb.z, wb.w);
    }
    Ring r;
    for (int u = 0; u < U1b; ++u) {
        uint64_t pp01 = zero2, pp23 = zero2, wl01 = zero2, wl23 = zero2;      // (no label: its term multiplies zeros)
        if (active) {
            const float4 pp = __ldg(reinterpret_cast<const float4*>(pproj + ((size_t)b * U1 + u) * H + h));
            pp01 = pk2(pp.x, pp.y);
            pp23 = pk2(pp.z, pp.w);
            if (EW) {
                const int lab = __ldg(a.row_label + base + (size_t)t0 * U1b + u);       // depends on (b, u) only
                if (lab >= 0 && lab != a.blank) {
                    const float4 wl = __ldg(reinterpret_cast<const float4*>(a.w_out + (size_t)lab * H + h));
                    wl01 = pk2(wl.x, wl.y);
                    wl23 = pk2(wl.z, wl.w);
                }
            }
        }
        mbar_wait(bar_full(r.stage), r.phase);
        uint64_t ap01 = zero2, ap23 = zero2;
        if (active) {
#pragma unroll
            for (int i = 0; i < TT; ++i) {
                if (i < nv) {
                    const uint4 g = *reinterpret_cast<const uint4*>(sSrc + (r.stage * TT + i) * kReduceSlab + hl);
                    uint64_t v01 = pk2u(g.x, g.y), v23 = pk2u(g.z, g.w);
                    float coef = gscale4;
                    if (EW) {
                        const float4 rm = sRm[r.stage * TT + i];
                        coef *= rm.w;
                        const uint64_t y2 = pk2(rm.y, rm.y), z2 = pk2(rm.z, rm.z);
                        v01 = fma2(y2, wb01, v01);
                        v23 = fma2(y2, wb23, v23);
                        v01 = fma2(z2, wl01, v01);
                        v23 = fma2(z2, wl23, v23);
                    }
                    float x0, x1, x2, x3;
                    unpk2(add2(e01[i], pp01), x0, x1);
                    unpk2(add2(e23[i], pp23), x2, x3);
                    const float q0 = ex2f(fabsf(x0) * kNeg2Log2e), q1 = ex2f(fabsf(x1) * kNeg2Log2e);
                    const float q2 = ex2f(fabsf(x2) * kNeg2Log2e), q3 = ex2f(fabsf(x3) * kNeg2Log2e);
                    const uint64_t q01 = pk2(q0, q1), q23 = pk2(q2, q3);
                    float d0, d1, d2, d3;
                    unpk2(add2(q01, one2), d0, d1);
                    unpk2(add2(q23, one2), d2, d3);
                    const uint64_t r01 = pk2(rcp_approx(d0), rcp_approx(d1));       // d in [1, 2]
                    const uint64_t r23 = pk2(rcp_approx(d2), rcp_approx(d3));
                    const uint64_t c2 = pk2(coef, coef);
                    const uint64_t s01 = mul2(mul2(q01, r01), mul2(r01, c2));      // coef * 4 e / (1 + e)^2
                    const uint64_t s23 = mul2(mul2(q23, r23), mul2(r23, c2));
                    v01 = mul2(v01, s01);
                    v23 = mul2(v23, s23);
                    a01[i] = add2(a01[i], v01);
                    a23[i] = add2(a23[i], v23);
                    ap01 = add2(ap01, v01);
                    ap23 = add2(ap23, v23);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(r.stage));              // the stage may be refilled
        if (active) {
            float p0, p1, p2, p3;
            unpk2(ap01, p0, p1);
            unpk2(ap23, p2, p3);
            red_add_v4(d_pproj + ((size_t)b * U1 + u) * H + h, p0, p1, p2, p3);
        }
        r.advance(kReduceStages);
    }
    if (active) {
#pragma unroll
        for (int i = 0; i < TT; ++i)
            if (t0 + i < T) {
                float4 o;
                unpk2(a01[i], o.x, o.y);
                unpk2(a23[i], o.z, o.w);
                *reinterpret_cast<float4*>(d_eproj + ((size_t)b * T + t0 + i) * H + h) = o;
            }
    }
}

constexpr int kBlankSlots = 64;                 // partial rows for the blank row's accumulation (dw_sparse_kernel)

// ------------------------------------------------------------------------------------------- dense-logits entry
// rnnt_loss called on a materialised (B,T,U1,V) fp32 logits tensor (someone else's joint): one warp per
// lattice cell streams the row once (online log-sum-exp) and drops the 3 floats into the compact row space.
__global__ void dense_lse_kernel(const float* __restrict__ acts, const int* __restrict__ labels,
                                 const int* __restrict__ act_lens, const int* __restrict__ label_lens,
                                 const int* __restrict__ meta, int B, int T, int U1, int V, int label_stride,
                                 int blank, float* __restrict__ lse, float* __restrict__ lpb,
                                 float* __restrict__ lpl, int* __restrict__ row_label) {
    const int tile = blockIdx.x;
    if (tile >= meta[0]) return;
    const int b = meta[kMetaHdr + B + 1 + tile];
    const int r0 = (tile - meta[kMetaHdr + b]) * kTile;
    const int Tb = act_lens[b], U1b = label_lens[b] + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    for (int lr = warp; lr < kTile; lr += nwarps) {
        const int r = r0 + lr;
        const size_t grow = (size_t)tile * kTile + lr;
        if (r >= Tb * U1b) {
            if (lane == 0) {
                lse[grow] = 0.f;
                lpb[grow] = 0.f;
                lpl[grow] = 0.f;
                row_label[grow] = -1;
            }
            continue;
        }
        const int t = r / U1b, u = r - t * U1b;
        const float* row = acts + (((size_t)b * T + t) * U1 + u) * V;
        float m = -INFINITY, s = 0.f;
        for (int v = lane; v < V; v += 32) {
            const float z = __ldg(row + v);
            const float mn = fmaxf(m, z);
            if (mn > -INFINITY) {
                s = s * __expf(m - mn) + __expf(z - mn);
                m = mn;
            }
        }
        for (int o = 16; o; o >>= 1) {
            const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
            const float mn = fmaxf(m, m2);
            s = ((m > -INFINITY) ? s * __expf(m - mn) : 0.f) + ((m2 > -INFINITY) ? s2 * __expf(m2 - mn) : 0.f);
            m = mn;
        }
        if (lane == 0) {
            const float l = m + __logf(s);
            int lab = (u < U1b - 1) ? labels[(size_t)b * label_stride + u] : -1;
            if (lab >= V) lab = -1;                        // (as joint_act_kernel: nothing gathers or scatters through it)
            lse[grow] = l;
            lpb[grow] = row[blank] - l;
            lpl[grow] = (lab >= 0) ? row[lab] - l : 0.f;
            row_label[grow] = lab;
        }
    }
}

// Dense gradient w.r.t. the logits (what upstream's compute_grad_kernel writes): one warp per (b,t,u) row of the
// PADDED tensor; rows outside the utterance's lattice get exact zeros.
__global__ void dense_grad_kernel(const float* __restrict__ acts, const float4* __restrict__ rowmeta,
                                  const int* __restrict__ row_label, const float* __restrict__ scal,
                                  const int* __restrict__ act_lens, const int* __restrict__ label_lens,
                                  const int* __restrict__ meta, int B, int T, int U1, int V, int blank,
                                  float* __restrict__ grads) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const size_t cell = (size_t)blockIdx.x * nwarps + warp;
    if (cell >= (size_t)B * T * U1) return;
    const int u = (int)(cell % U1);
    const int t = (int)((cell / U1) % T);
    const int b = (int)(cell / ((size_t)U1 * T));
    float* g = grads + cell * V;
    const int Tb = act_lens[b], U1b = label_lens[b] + 1;
    if (meta[1] != 0 || t >= Tb || u >= U1b) {
        for (int v = lane; v < V; v += 32) g[v] = 0.f;
        return;
    }
    const size_t grow = (size_t)meta[kMetaHdr + b] * kTile + (size_t)t * U1b + u;
    const float4 rm = rowmeta[grow];
    const int lab = row_label[grow];
    const float coef = rm.w * scal[2];
    const float* row = acts + cell * V;
    for (int v = lane; v < V; v += 32) {
        float pr = __expf(__ldg(row + v) - rm.x);
        if (v == blank) pr = rm.y;
        if (v == lab) pr = rm.z;
        g[v] = coef * pr;
    }
}

// ------------------------------------------------------------------------------------------- chunked wide-joint path
// Joint widths outside the fused tensor-core kernels (H = 1024, 2048 in the reference's configs): the host loops over
// chunks of lattice rows, the projection of a chunk runs as a library GEMM on the 16-bit operands (fp32 out), and
// these two kernels do everything else on the chunk, so memory stays bounded by the chunk, never B*T*U*V.
//   z: (R, Vpad) fp32 = A16[chunk] . W16^T  (un-biased, still scaled by w_scale);  one warp per row.
__global__ void rows_lse_kernel(const float* __restrict__ z, int R, int Vpad, int V, const float* __restrict__ bias2,
                                const float* __restrict__ scal, const int* __restrict__ row_label, int blank,
                                float* __restrict__ lse, float* __restrict__ lpb, float* __restrict__ lpl) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= R) return;
    const float c1 = scal[1] * kLog2e;
    const float* row = z + (size_t)r * Vpad;
    float m = -INFINITY, s = 0.f;
    for (int v = lane; v < V; v += 32) {
        const float y = fmaf(__ldg(row + v), c1, __ldg(bias2 + v));
        const float mn = fmaxf(m, y);
        s = s * ex2f(m - mn) + ex2f(y - mn);
        m = mn;
    }
    for (int o = 16; o; o >>= 1) {
        const float m2 = __shfl_xor_sync(0xffffffffu, m, o), s2 = __shfl_xor_sync(0xffffffffu, s, o);
        const float mn = fmaxf(m, m2);
        s = ((m > -INFINITY) ? s * ex2f(m - mn) : 0.f) + ((m2 > -INFINITY) ? s2 * ex2f(m2 - mn) : 0.f);
        m = mn;
    }
    if (lane == 0) {
        const float l2 = m + lg2f(s);
        const int lab = row_label[r];
        lse[r] = l2 * kLn2;
        lpb[r] = (fmaf(row[blank], c1, bias2[blank]) - l2) * kLn2;
        lpl[r] = (lab >= 0) ? (fmaf(row[lab], c1, bias2[lab]) - l2) * kLn2 : 0.f;
    }
}

// q[r, v] = 16-bit(scale * w_r * (softmax(r, v) - rb [v == blank] - rl [v == label])) for v < V, 0 on the padding
// columns: the operand of both gradient GEMMs (dA = q . W16, dW = q^T . A16).  Also db[v] += w_r * gmax * (...).
template <bool BF16>
__global__ void rows_grad_kernel(const float* __restrict__ z, const float4* __restrict__ rowmeta,
                                 const int* __restrict__ row_label, const float* __restrict__ bias2,
                                 const float* __restrict__ scal, int R, int Vpad, int V, int blank,
                                 uint16_t* __restrict__ q) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + warp;
    if (r >= R) return;
    const float c1 = scal[1] * kLog2e;
    const float pscale = BF16 ? 1.f : kPScale;
    const float4 rm = rowmeta[r];
    const int lab = row_label[r];
    const float k = fmaf(rm.x, -kLog2e, 0.f);       // -lse2 (+inf lse on padding rows -> -inf -> zeros)
    const float wq = rm.w * pscale;
    const float* row = z + (size_t)r * Vpad;
    uint16_t* out = q + (size_t)r * Vpad;
    for (int v = lane * 2; v < Vpad; v += 64) {
        float a = 0.f, b = 0.f;
        if (v < V) {
            a = ex2f(fmaf(__ldg(row + v), c1, __ldg(bias2 + v) + k));
            if (v == blank) a = rm.y;
            if (v == lab) a = rm.z;
            a *= wq;
        }
        if (v + 1 < V) {
            b = ex2f(fmaf(__ldg(row + v + 1), c1, __ldg(bias2 + v + 1) + k));
            if (v + 1 == blank) b = rm.y;
            if (v + 1 == lab) b = rm.z;
            b *= wq;
        }
        *reinterpret_cast<uint32_t*>(out + v) = pack16<BF16>(a, b);
    }
}

int launch_rows_lse(const float* z, int R, int Vpad, int V, const float* bias2, const float* scal,
                    const int* row_label, int blank, float* lse, float* lpb, float* lpl, cudaStream_t s) {
    const int wpb = 8;
    rows_lse_kernel<<<(R + wpb - 1) / wpb, wpb * 32, 0, s>>>(z, R, Vpad, V, bias2, scal, row_label, blank, lse, lpb, lpl);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_rows_grad(const float* z, const float4* rowmeta, const int* row_label, const float* bias2,
                     const float* scal, int R, int Vpad, int V, int blank, bool bf16, void* q, cudaStream_t s) {
    const int wpb = 8;
    if (bf16)
        rows_grad_kernel<true><<<(R + wpb - 1) / wpb, wpb * 32, 0, s>>>(z, rowmeta, row_label, bias2, scal, R, Vpad, V,
                                                                        blank, (uint16_t*)q);
    else
        rows_grad_kernel<false><<<(R + wpb - 1) / wpb, wpb * 32, 0, s>>>(z, rowmeta, row_label, bias2, scal, R, Vpad, V,
                                                                         blank, (uint16_t*)q);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------- launchers
int launch_prep(const int* act_lens, const int* label_lens, int B, int T, int U1, int n_tiles_ub, int* meta,
                cudaStream_t s) {
    prep_kernel<<<1, 1024, 0, s>>>(act_lens, label_lens, B, T, U1, n_tiles_ub, meta);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_cast_w(const float* w, const float* b_out, int V, int Vpad, int H, bool bf16, float* scal, void* w16,
                  float* bias2, void* w16t, cudaStream_t s) {
    unsigned int* bits = reinterpret_cast<unsigned int*>(scal + 4);
    TTX_CUDA_OK(cudaMemsetAsync(bits, 0, sizeof(unsigned int), s));
    const size_t n = (size_t)V * H;
    if (!bf16) absmax_kernel<<<296, 256, 0, s>>>(w, n, bits);
    if (bf16)
        cast_w_kernel<true><<<592, 256, 0, s>>>(w, b_out, V, Vpad, H, bits, scal, (uint16_t*)w16, bias2, (uint16_t*)w16t);
    else
        cast_w_kernel<false><<<592, 256, 0, s>>>(w, b_out, V, Vpad, H, bits, scal, (uint16_t*)w16, bias2, (uint16_t*)w16t);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_joint_act(const float* eproj, const float* pproj, const int* labels, const int* act_lens,
                     const int* label_lens, const int* meta, int B, int T, int U1, int H, int label_stride, int V,
                     int n_tiles_ub, bool bf16, void* a16, int* row_label, void* a16t, cudaStream_t s) {
    const size_t rows_total = (size_t)n_tiles_ub * kTile;
    if (bf16)
        joint_act_kernel<true><<<n_tiles_ub, 256, 0, s>>>(eproj, pproj, labels, act_lens, label_lens, meta, B, T, U1,
                                                         H, label_stride, V, (uint16_t*)a16, row_label, (uint16_t*)a16t,
                                                         rows_total);
    else
        joint_act_kernel<false><<<n_tiles_ub, 256, 0, s>>>(eproj, pproj, labels, act_lens, label_lens, meta, B, T,
                                                          U1, H, label_stride, V, (uint16_t*)a16, row_label,
                                                          (uint16_t*)a16t, rows_total);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K, int PD, int W>
static void lattice_launch(int B, cudaStream_t s, const float* ws, size_t lat_elems, const int* act_lens,
                           const int* label_lens, const int* meta, double* alpha, double* beta, float* costs,
                           double* ll_beta) {
    lattice_wave_kernel<K, PD, W><<<2 * B, 32 * W, 0, s>>>(ws, lat_elems, act_lens, label_lens, meta, B, alpha, beta, costs,
                                                          ll_beta);
}

// lat_ws: 4 * lat_elems floats (the arc log-probs in diagonal-major order); alpha / beta: lat_elems doubles each
int launch_lattice(const float* lpb, const float* lpl, const int* act_lens, const int* label_lens, const int* meta,
                   int B, int U1, int n_tiles_ub, size_t lat_elems, float* lat_ws, double* alpha, double* beta,
                   float* costs, double* ll_beta, cudaStream_t s) {
    if (U1 > 1024) {
        set_error("lattice kernel supports at most 1023 labels per utterance (got U+1 = %d)", U1);
        return 1;
    }
    lattice_skew_kernel<<<n_tiles_ub, kTile, 0, s>>>(lpb, lpl, act_lens, label_lens, meta, B, lat_elems, lat_ws);
#define TTX_LAT(K, PD, W) lattice_launch<K, PD, W>(B, s, lat_ws, lat_elems, act_lens, label_lens, meta, alpha, beta, costs, ll_beta)
    // (one column per lane with twice the warps measured the same: 0.140 vs 0.136 ms at configs[1], 0.45 vs 0.47 at configs[3])
    if (U1 <= 32) TTX_LAT(1, 8, 1);
    else if (U1 <= 64) TTX_LAT(2, 8, 1);
    else if (U1 <= 128) TTX_LAT(2, 8, 2);
    else if (U1 <= 256) TTX_LAT(2, 8, 4);
    else if (U1 <= 512) TTX_LAT(2, 8, 8);
    else TTX_LAT(2, 4, 16);
#undef TTX_LAT
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_grad_prep(const float* lse, const float* lpb, const float* lpl, const double* alpha, const double* beta,
                     const double* ll_beta, const float* grad_costs, float* scal, const int* row_label,
                     const int* act_lens, const int* label_lens, const int* meta, int B, int blank, int n_tiles_ub,
                     float4* rowmeta, float* d_b_out, cudaStream_t s) {
    gmax_kernel<<<1, 256, 0, s>>>(grad_costs, B, scal);
    grad_prep_kernel<<<n_tiles_ub, kTile, 0, s>>>(lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label,
                                                  act_lens, label_lens, meta, B, blank, rowmeta, d_b_out);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_reduce(const float* src, const float4* rowmeta, const int* row_label, const float* w_out,
                  const float* scal, int blank, const float* eproj, const float* pproj, const int* act_lens,
                  const int* label_lens, const int* meta, int B, int T, int U1, int H, float* d_eproj,
                  float* d_pproj, cudaStream_t s) {
    const ActGradSrc a{src, rowmeta, row_label, w_out, scal, blank};
    TTX_CUDA_OK(cudaMemsetAsync(d_pproj, 0, (size_t)B * U1 * H * sizeof(float), s));
    const dim3 grid((T + kReduceFrames - 1) / kReduceFrames, B, (H + kReduceSlab - 1) / kReduceSlab);
    const size_t smem = (size_t)kReduceStages * kReduceFrames * (kReduceSlab * 4 + 16) + 2 * kReduceStages * 8;
    if (rowmeta != nullptr) {
        TTX_CUDA_OK(cudaFuncSetAttribute(reduce_both_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        reduce_both_kernel<true><<<grid, kReduceThreads, smem, s>>>(a, eproj, pproj, act_lens, label_lens, meta, T, U1, H, d_eproj, d_pproj);
    } else {
        TTX_CUDA_OK(cudaFuncSetAttribute(reduce_both_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        reduce_both_kernel<false><<<grid, kReduceThreads, smem, s>>>(a, eproj, pproj, act_lens, label_lens, meta, T, U1, H, d_eproj, d_pproj);
    }
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}


// ------------------------------------------------------------------------------------------- weight gradient from kept P'
// The S pass keeps its softmax numerators P' (rows x Vpad, 16 bit; softmax = P' * pfac[row]).  The dense part of the weight
// gradient is then one product with no projection pass and no exponentials,
//   dW[v, h] = gmax * 2^-shift * sum_m P'[m, v] * As[m, h],   As[m, h] = 2^shift * w_m * pfac_m * A16[m, h].
// The scale is per lattice row -- the contraction index -- so it cannot ride on the product's epilogue, and P' is
// streamed straight into the tensor core: As is a scaled copy of A16, written here, ROW-MAJOR like A16 (the product reads
// it MN-major, so no transposed copy of the activations exists anywhere on this route).  svec[m] = 2^shift * w_m * pfac_m
// goes out next to it: its product with P' is the dense part of db.  shift (16 for fp16, 0 for bf16) keeps As out of the
// subnormals.  Rows beyond the tiles in use, and padding rows (w = 0), give exact zeros.
template <bool BF16>
__global__ void scale_rows_kernel(const uint16_t* __restrict__ a16, const float4* __restrict__ rowmeta,
                                  const float* __restrict__ pfac, const int* __restrict__ meta, int H, size_t rows_total,
                                  float up, uint16_t* __restrict__ as, uint16_t* __restrict__ svec) {
    // thread = 8 joint columns of one lattice row; a warp covers 256 columns of a row (or several rows for H < 256)
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int per_row = H / 8;
    const size_t m = idx / per_row;
    if (m >= rows_total) return;
    const int h = (int)(idx - m * per_row) * 8;
    float sc = 0.f;
    if (m < (size_t)meta[0] * kTile) {
        const float w = __ldg(&rowmeta[m].w);
        if (w != 0.f) sc = w * __ldg(pfac + m) * up;
    }
    uint4 o = make_uint4(0, 0, 0, 0);
    if (sc != 0.f) {
        const uint4 in = __ldg(reinterpret_cast<const uint4*>(a16 + m * H + h));
        const uint32_t w4[4] = {in.x, in.y, in.z, in.w};
        uint32_t r4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float x0, x1;
            unpk16<BF16>(w4[e], x0, x1);
            r4[e] = pack16<BF16>(x0 * sc, x1 * sc);
        }
        o = make_uint4(r4[0], r4[1], r4[2], r4[3]);
    }
    *reinterpret_cast<uint4*>(as + m * H + h) = o;
    if (h == 0) svec[m] = (uint16_t)(pack16<BF16>(sc, 0.f) & 0xffffu);
}

// The blank and label entries are absent from the kept P' (the forward+gradient launch zeroes them): their exact terms
//   dW[blank] += gmax * sum_m w (p_b - rb) A[m],  dW[label_m] += gmax * w (p_l - rl) A[m]   (rowmeta .y / .z),
//   db[blank] += gmax * sum_m w p_b,              db[label_m] += gmax * w p_l               (dense part they left out)
// are added here.  grid = (U1, B, t-chunks) like reduce_pred_kernel: the label depends on (b, u) only.
template <bool BF16>
__global__ void __launch_bounds__(128, 5) dw_sparse_kernel(const uint16_t* __restrict__ a16, const float4* __restrict__ rowmeta,
                                 const int* __restrict__ row_label, const float* __restrict__ lpb,
                                 const float* __restrict__ lpl, const float* __restrict__ scal,
                                 const int* __restrict__ act_lens, const int* __restrict__ label_lens,
                                 const int* __restrict__ meta, int U1, int H, int blank,
                                 int t_chunk, float* __restrict__ d_w, float* __restrict__ d_b,
                                 float* __restrict__ blank_slots, const float* __restrict__ pfac, float up,
                                 uint16_t* __restrict__ as, uint16_t* __restrict__ svec) {
    // as != null: the pass also writes the scaled operand copy As = (up * w * pfac) * A16 of the rows it visits (and their
    // scales) -- it has the row and its weight in registers anyway, so A16 is read once for both (scale_rows_kernel is
    // the stand-alone version; rows outside the utterances are zeroed by pad_rows_kernel)
    if (meta[1] != 0) return;
    const int u = blockIdx.x, b = blockIdx.y;
    const int Tb = act_lens[b], U1b = label_lens[b] + 1;
    if (u >= U1b) return;
    const int t0 = blockIdx.z * t_chunk, t1 = min(Tb, t0 + t_chunk);
    if (t0 >= t1) return;
    const size_t base = (size_t)meta[kMetaHdr + b] * kTile;
    const int lab = __ldg(row_label + base + (size_t)t0 * U1b + u);
    const bool has_label = lab >= 0 && lab != blank;
    const float gmax = scal[2];
    // thread = 8 joint columns (one 16-byte load per row); the block's two 64-thread groups take alternate rounds of
    // eight frames, all loads of a round issued before the first use (128 bytes in flight per thread)
    const int grp = threadIdx.x >> 6, ngrp = blockDim.x >> 6;
    for (int h = (threadIdx.x & 63) * 8; h < H; h += 64 * 8) {
        float ab[8], al[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) ab[k] = al[k] = 0.f;
        float dbb = 0.f, dbl = 0.f;
        for (int tb = t0 + grp * 8; tb < t1; tb += ngrp * 8) {
            float ry[8], rz[8];
            uint4 av[8];
            float eb[8], el[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const bool in = tb + e < t1;
                const size_t m = base + (size_t)(in ? tb + e : t0) * U1b + u;
                const float4 rm = __ldg(rowmeta + m);
                const float w = in ? rm.w : 0.f;
                ry[e] = w * rm.y;
                rz[e] = w * rm.z;
                av[e] = __ldg(reinterpret_cast<const uint4*>(a16 + m * H + h));
                eb[e] = (h == 0) ? w * __expf(__ldg(lpb + m)) : 0.f;
                el[e] = (h == 0 && has_label) ? w * __expf(__ldg(lpl + m)) : 0.f;
                if (as != nullptr && in) {
                    const float sc = (w != 0.f) ? w * __ldg(pfac + m) * up : 0.f;
                    uint4 o = make_uint4(0, 0, 0, 0);
                    if (sc != 0.f) {
                        const uint32_t i4[4] = {av[e].x, av[e].y, av[e].z, av[e].w};
                        uint32_t r4[4];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float x0, x1;
                            unpk16<BF16>(i4[k], x0, x1);
                            r4[k] = pack16<BF16>(x0 * sc, x1 * sc);
                        }
                        o = make_uint4(r4[0], r4[1], r4[2], r4[3]);
                    }
                    *reinterpret_cast<uint4*>(as + m * H + h) = o;
                    if (h == 0) svec[m] = (uint16_t)(pack16<BF16>(sc, 0.f) & 0xffffu);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t w4[4] = {av[e].x, av[e].y, av[e].z, av[e].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float a0, a1;
                    unpk16<BF16>(w4[k], a0, a1);
                    ab[2 * k] = fmaf(ry[e], a0, ab[2 * k]);
                    ab[2 * k + 1] = fmaf(ry[e], a1, ab[2 * k + 1]);
                    if (has_label) {
                        al[2 * k] = fmaf(rz[e], a0, al[2 * k]);
                        al[2 * k + 1] = fmaf(rz[e], a1, al[2 * k + 1]);
                    }
                }
                dbb += eb[e];
                dbl += el[e];
            }
        }
        // every block adds to the blank row: spread over kBlankSlots partial rows (same-address atomics serialise in L2),
        // folded into d_w[blank] / d_b[blank] by blank_fold_kernel
        float* db_ = blank_slots + (size_t)((blockIdx.x + blockIdx.y * gridDim.x + blockIdx.z * 7) % kBlankSlots) * (H + 4) + h;
        red_add_v4(db_, ab[0] * gmax, ab[1] * gmax, ab[2] * gmax, ab[3] * gmax);
        red_add_v4(db_ + 4, ab[4] * gmax, ab[5] * gmax, ab[6] * gmax, ab[7] * gmax);
        if (has_label) {
            float* dl = d_w + (size_t)lab * H + h;
            red_add_v4(dl, al[0] * gmax, al[1] * gmax, al[2] * gmax, al[3] * gmax);
            red_add_v4(dl + 4, al[4] * gmax, al[5] * gmax, al[6] * gmax, al[7] * gmax);
        }
        if (h == 0) {
            atomicAdd(db_ + H, dbb * gmax);
            if (has_label) atomicAdd(d_b + lab, dbl * gmax);
        }
    }
}

// rows of As / svec that belong to no lattice cell -- the tail of every utterance's last tile, and the pad tile of an odd
// tile count -- must be exact zeros (their P' rows are finite).  grid = B + 1.
__global__ void pad_rows_kernel(const int* __restrict__ act_lens, const int* __restrict__ label_lens,
                                const int* __restrict__ meta, int B, int H, uint16_t* __restrict__ as,
                                uint16_t* __restrict__ svec) {
    if (meta[1] != 0) return;
    size_t r0, r1;
    if ((int)blockIdx.x < B) {
        const int b = blockIdx.x;
        r0 = (size_t)meta[kMetaHdr + b] * kTile + (size_t)act_lens[b] * (label_lens[b] + 1);
        r1 = (size_t)meta[kMetaHdr + b + 1] * kTile;
    } else {
        r0 = (size_t)meta[0] * kTile;
        r1 = (size_t)((meta[0] + 1) & ~1) * kTile;
    }
    const size_t n16 = (r1 - r0) * (size_t)(H / 8);                  // 16-byte pieces
    uint4* dst = reinterpret_cast<uint4*>(as + r0 * H);
    for (size_t i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = make_uint4(0, 0, 0, 0);
    for (size_t r = r0 + threadIdx.x; r < r1; r += blockDim.x) svec[r] = 0;
}

__global__ void blank_fold_kernel(const float* __restrict__ blank_slots, int H, int blank,
                                  float* __restrict__ d_w, float* __restrict__ d_b) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h > H) return;
    float acc = 0.f;
    for (int k = 0; k < kBlankSlots; ++k) acc += blank_slots[(size_t)k * (H + 4) + h];
    if (h < H) atomicAdd(d_w + (size_t)blank * H + h, acc);
    else atomicAdd(d_b + blank, acc);
}

int launch_kept_prepare(const void* a16, const float4* rowmeta, const int* row_label,
                        const float* lpb, const float* lpl, const float* pfac, const float* scal, const int* act_lens,
                        const int* label_lens, const int* meta, int B, int T, int U1, int H, int blank,
                        bool bf16, size_t rows_total, void* a16st, float* d_w, float* d_b, int parts, cudaStream_t s) {
    const float up = bf16 ? 1.f : kKeptUp;
    uint16_t* as = static_cast<uint16_t*>(a16st);
    uint16_t* svec = as + rows_total * (size_t)H;              // [rows]: behind the matrix (see ttx.h for the layout)
    if (parts == 1) {                            // the scaled operand copy alone
        const unsigned g1 = (unsigned)((rows_total * (size_t)(H / 8) + 255) / 256);
        if (bf16)
            scale_rows_kernel<true><<<g1, 256, 0, s>>>((const uint16_t*)a16, rowmeta, pfac, meta, H, rows_total, up, as, svec);
        else
            scale_rows_kernel<false><<<g1, 256, 0, s>>>((const uint16_t*)a16, rowmeta, pfac, meta, H, rows_total, up, as, svec);
        TTX_CUDA_OK(cudaGetLastError());
        return 0;
    }
    // the blank / label terms, and (parts == 3) the operand copy in the same pass over A16
    const bool both = parts == 3;
    if (both) pad_rows_kernel<<<B + 1, 256, 0, s>>>(act_lens, label_lens, meta, B, H, as, svec);
    const int t_chunk = 64;                      // (longer chunks = fewer atomics were slower: 0.31 -> 0.35 ms at 512)
    const dim3 g2(U1, B, (T + t_chunk - 1) / t_chunk);
    const int threads = 128;                     // two groups of 64 threads x 8 joint columns
    // the partial rows live behind the (H + 16) x rows_total matrix in the caller's a16st buffer
    float* slots = reinterpret_cast<float*>(static_cast<uint8_t*>(a16st) + (size_t)(H + 16) * rows_total * 2);
    const size_t slot_bytes = (size_t)kBlankSlots * (H + 4) * sizeof(float);
    TTX_CUDA_OK(cudaMemsetAsync(slots, 0, slot_bytes, s));
    if (bf16)
        dw_sparse_kernel<true><<<g2, threads, 0, s>>>((const uint16_t*)a16, rowmeta, row_label, lpb, lpl, scal, act_lens,
                                                      label_lens, meta, U1, H, blank, t_chunk, d_w, d_b, slots, pfac, up,
                                                      both ? as : nullptr, svec);
    else
        dw_sparse_kernel<false><<<g2, threads, 0, s>>>((const uint16_t*)a16, rowmeta, row_label, lpb, lpl, scal, act_lens,
                                                       label_lens, meta, U1, H, blank, t_chunk, d_w, d_b, slots, pfac, up,
                                                       both ? as : nullptr, svec);
    blank_fold_kernel<<<(H + 1 + 127) / 128, 128, 0, s>>>(slots, H, blank, d_w, d_b);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_dense_lse(const float* acts, const int* labels, const int* act_lens, const int* label_lens,
                     const int* meta, int B, int T, int U1, int V, int label_stride, int blank, int n_tiles_ub,
                     float* lse, float* lpb, float* lpl, int* row_label, cudaStream_t s) {
    dense_lse_kernel<<<n_tiles_ub, 256, 0, s>>>(acts, labels, act_lens, label_lens, meta, B, T, U1, V, label_stride,
                                               blank, lse, lpb, lpl, row_label);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_dense_grad(const float* acts, const float4* rowmeta, const int* row_label, const float* scal,
                      const int* act_lens, const int* label_lens, const int* meta, int B, int T, int U1, int V,
                      int blank, float* grads, cudaStream_t s) {
    const size_t cells = (size_t)B * T * U1;
    const int wpb = 8;
    dense_grad_kernel<<<(unsigned)((cells + wpb - 1) / wpb), wpb * 32, 0, s>>>(acts, rowmeta, row_label, scal,
                                                                              act_lens, label_lens, meta, B, T, U1, V,
                                                                              blank, grads);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ttx
