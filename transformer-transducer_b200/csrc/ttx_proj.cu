// Pre-projections of the joint network, forward and backward, as tcgen05 GEMMs on fp32 data:
//   tt JointNet        forward_layer split algebraically into its encoder / decoder halves  (/root/reference/tt/model.py:35)
//   espnet JointNetwork lin_enc / lin_dec                  (espnet/nets/pytorch_backend/transducer/joint_network.py:28-31,48)
// The reference runs them as fp32 SGEMMs, and their rounding reaches the loss and every gradient, so a plain TF32
// product (2^-11 operand rounding) is not good enough.  Error-compensated 3 x TF32 instead: every operand is split on
// the fly into x = hi + lo, hi = x with the low 13 mantissa bits cleared (exactly representable in TF32), lo = x - hi
// (exact in fp32, 13 significant bits), and the accumulator in TMEM collects hi.hi + lo.hi + hi.lo in fp32 -- the
// dropped lo.lo term and the rounding of lo are 2^-22 relative; what remains is the tensor core's truncating fp32
// accumulation (measured 3e-6 .. 8e-6 relative L2 against float64; plain TF32: 3e-4, SGEMM: 3e-7).
//
// One kernel, three uses (output [Mo x No], contraction over Kc):
//   forward   y  = x . w^T + b      A = x  [Mo x Kc] row-major (K-major),   B = w [No x Kc] row-major (K-major)
//   backward  dx = dy . w           A = dy [Mo x Kc] row-major (K-major),   B from w  stored [Kc x No]  -> transposed on chip
//   backward  dw = dy^T . x         A from dy stored [Kc x Mo], B from x stored [Kc x No]  -> both transposed on chip,
//                                   contraction (B*T rows) split over gridDim.z, red.add into dw
// MN-major TF32 operands need the 32-byte-atom swizzle mode; rather than depend on it, the warps that split hi / lo --
// they touch every element anyway -- write the transposed tile in the K-major 128-byte-swizzle layout themselves.
//
// CTA = 128 x 128 output tile, contraction chunks of 32 floats (one 128-byte swizzle row).  Warp 0: TMA producer,
// warp 1: MMA issuer (cta_group::1, M = 128, N = 128, K = 8 per instruction, 12 per chunk), warps 2-9: split /
// transpose every landed chunk, then read the accumulator out of TMEM (bias add, 16-byte stores or reductions).
// Shared-memory bandwidth is what bounds it: per chunk the MMAs read 6 x 16 KiB and the split moves another 96 KiB.
#include "ttx_common.cuh"

namespace ttx {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode();

constexpr int kPT = 128;                       // output tile (rows and columns)
constexpr int kPK = 32;                        // contraction chunk: 32 floats = 128 bytes
constexpr int kPBytes = kPT * kPK * 4;         // one operand chunk: 16 KiB
constexpr int kPWorkWarps = 8;
constexpr int kPThreads = (2 + kPWorkWarps) * 32;

struct ProjParams {
    int Mo, No, Kc;          // output rows / columns, contraction length
    int chunks_per_split;    // contraction chunks per blockIdx.z
    int NS;                  // pipeline stages
    int ldo;                 // leading dimension of out (floats)
    int accumulate;          // 1: red.add into out (split contraction), 0: plain stores
    const float* bias;       // (No) or null
    float* out;
};

__device__ __forceinline__ void umma_tf32_ss_lo(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
        : "memory");
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
    lo = x - hi;
}

// TA / TB: the operand's source is stored [Kc x MN] (contraction along rows) and is transposed on chip.
template <bool TA, bool TB>
__global__ void __launch_bounds__(kPThreads, 1)
proj_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ProjParams p) {
    constexpr int STAGE = 4 * kPBytes + (TA ? kPBytes : 0) + (TB ? kPBytes : 0);
    const int m0 = blockIdx.y * kPT, n0 = blockIdx.x * kPT;
    const int total_chunks = (p.Kc + kPK - 1) / kPK;
    const int c_lo = blockIdx.z * p.chunks_per_split;
    const int n_chunks = min(total_chunks, c_lo + p.chunks_per_split) - c_lo;
    if (n_chunks <= 0) return;                                   // (whole CTA alike)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) {
        if (threadIdx.x == 0) printf("ttx: dynamic shared memory is not 1024-byte aligned (0x%x)\n", smem_base);
        __trap();
    }
    // stage layout: hiA | loA | hiB | loB | rawA (TA) | rawB (TB); a K-major source lands in its hi buffer and is split in place
    auto s_hiA = [&](int s) { return smem_base + s * STAGE; };
    auto s_loA = [&](int s) { return s_hiA(s) + kPBytes; };
    auto s_hiB = [&](int s) { return s_hiA(s) + 2 * kPBytes; };
    auto s_loB = [&](int s) { return s_hiA(s) + 3 * kPBytes; };
    auto s_rawA = [&](int s) { return s_hiA(s) + 4 * kPBytes; };
    auto s_rawB = [&](int s) { return s_hiA(s) + 4 * kPBytes + (TA ? kPBytes : 0); };
    const uint32_t sBar = smem_base + p.NS * STAGE;
    auto bar_full = [&](int s) { return sBar + 8 * s; };          // TMA landed
    auto bar_ready = [&](int s) { return sBar + 8 * (4 + s); };   // hi / lo tiles written
    auto bar_empty = [&](int s) { return sBar + 8 * (8 + s); };   // MMAs have read the stage
    const uint32_t bar_acc = sBar + 8 * 12;
    const uint32_t sTmemPtr = sBar + 8 * 13;
    float* sBias = reinterpret_cast<float*>(smem_raw + (sBar + 128 - smem_base));      // this tile's 128 bias values
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
        for (int s = 0; s < p.NS; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_ready(s), kPWorkWarps);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(sTmemPtr, kPT);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_raw + (sTmemPtr - smem_base));

    if (warp == 0) {
        if (lane == 0) {
            // ======================================================= TMA producer
            Ring r;
            for (int c = 0; c < n_chunks; ++c) {
                const int k0 = (c_lo + c) * kPK;
                mbar_wait(bar_empty(r.stage), r.phase ^ 1);
                mbar_arrive_expect_tx(bar_full(r.stage), 2 * kPBytes);
                if (TA) tma_load_2d(s_rawA(r.stage), &mapA, bar_full(r.stage), m0, k0);      // [32 k rows x 128 m], plain
                else tma_load_2d(s_hiA(r.stage), &mapA, bar_full(r.stage), k0, m0);          // [128 m rows x 32 k], swizzled
                if (TB) tma_load_2d(s_rawB(r.stage), &mapB, bar_full(r.stage), n0, k0);
                else tma_load_2d(s_hiB(r.stage), &mapB, bar_full(r.stage), k0, n0);
                r.advance(p.NS);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ======================================================= MMA issuer
            const uint32_t idesc = make_idesc(2, 0, 0, kPT, kPT);          // TF32 operands, fp32 accumulate, K-major A and B
            Ring r;
            for (int c = 0; c < n_chunks; ++c) {
                mbar_wait(bar_ready(r.stage), r.phase);
                tc_fence_after();
                const uint32_t ah = desc_lo(s_hiA(r.stage)), al = desc_lo(s_loA(r.stage));
                const uint32_t bh = desc_lo(s_hiB(r.stage)), bl = desc_lo(s_loB(r.stage));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_ss_lo(tmem_base, al + 2 * k, bh + 2 * k, idesc, (c | k) != 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_ss_lo(tmem_base, ah + 2 * k, bl + 2 * k, idesc, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32_ss_lo(tmem_base, ah + 2 * k, bh + 2 * k, idesc, 1);
                umma_commit(bar_empty(r.stage));
                r.advance(p.NS);
            }
            umma_commit(bar_acc);
        }
    } else {
        // =========================================================== split (+ transpose) warps, then the epilogue
        const int wt = threadIdx.x - 64;                            // 0 .. 255
        if (wt < kPT) sBias[wt] = (p.bias && n0 + wt < p.No) ? __ldg(p.bias + n0 + wt) : 0.f;
        Ring r;
        for (int c = 0; c < n_chunks; ++c) {
            mbar_wait(bar_full(r.stage), r.phase);
            uint8_t* base = smem_raw + (s_hiA(r.stage) - smem_base);
#pragma unroll
            for (int op = 0; op < 2; ++op) {
                uint8_t* hi = base + op * 2 * kPBytes;
                uint8_t* lo = hi + kPBytes;
                const bool trans = op == 0 ? TA : TB;
                if (!trans) {
                    // K-major source, landed swizzled in `hi`: split in place, position by position
#pragma unroll
                    for (int it = 0; it < kPBytes / 16 / (kPWorkWarps * 32); ++it) {
                        const int off = (it * kPWorkWarps * 32 + wt) * 16;
                        const float4 v = *reinterpret_cast<const float4*>(hi + off);
                        float4 h, l;
                        split_tf32(v.x, h.x, l.x);
                        split_tf32(v.y, h.y, l.y);
                        split_tf32(v.z, h.z, l.z);
                        split_tf32(v.w, h.w, l.w);
                        *reinterpret_cast<float4*>(hi + off) = h;
                        *reinterpret_cast<float4*>(lo + off) = l;
                    }
                } else {
                    // source chunk [32 k rows x 128 mn] (plain rows of 512 bytes): this thread takes column mn and half of the
                    // k rows, and writes row mn of the K-major tiles: 16-byte chunk k4 of the 128-byte row goes to position
                    // k4 ^ (mn & 7)
                    const float* raw = reinterpret_cast<const float*>(base + 4 * kPBytes + ((op == 1 && TA) ? kPBytes : 0));
                    const int mn = wt & (kPT - 1), kh = (wt >> 7) * 4;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k4 = kh + kk;
                        float4 h, l;
                        split_tf32(raw[(k4 * 4 + 0) * kPT + mn], h.x, l.x);
                        split_tf32(raw[(k4 * 4 + 1) * kPT + mn], h.y, l.y);
                        split_tf32(raw[(k4 * 4 + 2) * kPT + mn], h.z, l.z);
                        split_tf32(raw[(k4 * 4 + 3) * kPT + mn], h.w, l.w);
                        const int off = mn * 128 + ((k4 ^ (mn & 7)) << 4);
                        *reinterpret_cast<float4*>(hi + off) = h;
                        *reinterpret_cast<float4*>(lo + off) = l;
                    }
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_ready(r.stage));
            r.advance(p.NS);
        }
        // ---- epilogue: warp w reads TMEM lanes 32 * (w % 4) ..; thread = one output row, 32 columns per load; the two
        // warps of a lane quarter take alternate column groups
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        asm volatile("bar.sync 1, %0;" ::"n"(kPWorkWarps * 32) : "memory");        // sBias is complete
        mbar_wait(bar_acc, 0);
        tc_fence_after();
        uint32_t acc[32];
#pragma unroll 1
        for (int cc = (warp - 2) >> 2; cc < kPT / 32; cc += 2) {
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + cc * 32, acc);
            tmem_ld_wait();
            const int col0 = n0 + cc * 32;
            if (row < p.Mo) {
                float* dst = p.out + (size_t)row * p.ldo + col0;
#pragma unroll
                for (int e = 0; e < 32; e += 4) {
                    if (col0 + e < p.No) {                          // No % 4 == 0
                        float4 v = make_float4(__uint_as_float(acc[e]), __uint_as_float(acc[e + 1]),
                                               __uint_as_float(acc[e + 2]), __uint_as_float(acc[e + 3]));
                        const float4 b = *reinterpret_cast<const float4*>(sBias + cc * 32 + e);
                        v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
                        if (p.accumulate) red_add_v4(dst + e, v.x, v.y, v.z, v.w);
                        else *reinterpret_cast<float4*>(dst + e) = v;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kPT);
    }
}

// column sums of a row-major [rows x cols] matrix (bias gradient): grid (ceil(cols / 128), row splits), red.add
__global__ void colsum_kernel(const float* __restrict__ x, int rows, int cols, int ld, int rows_per_block,
                              float* __restrict__ out) {
    const int c = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < cols)
        for (int r = r0 + (threadIdx.x >> 5); r < r1; r += blockDim.x >> 5) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + (size_t)r * ld + c));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    __shared__ float4 part[8][32];
    part[threadIdx.x >> 5][threadIdx.x & 31] = acc;
    __syncthreads();
    if (threadIdx.x < 32 && c < cols) {
        float4 s = part[0][threadIdx.x];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            const float4 v = part[w][threadIdx.x];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        red_add_v4(out + c, s.x, s.y, s.z, s.w);
    }
}

// ------------------------------------------------------------------------------------------- host side
// fp32 row-major [rows x cols] with leading dimension ld; box = box_cols x box_rows
static int make_f32_map(CUtensorMap* map, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, int box_cols,
                        int box_rows, bool swizzle) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (fp32) failed with CUresult %d (rows=%llu cols=%llu ld=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld);
        return 2;
    }
    return 0;
}

template <bool TA, bool TB>
static int launch_proj_t(const CUtensorMap& ma, const CUtensorMap& mb, ProjParams p, dim3 grid, cudaStream_t stream) {
    constexpr int STAGE = 4 * kPBytes + (TA ? kPBytes : 0) + (TB ? kPBytes : 0);
    p.NS = min(4, (232448 - 1024) / STAGE);
    const size_t smem = (size_t)p.NS * STAGE + 1024;
    auto kern = proj_kernel<TA, TB>;
    TTX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    kern<<<grid, kPThreads, smem, stream>>>(ma, mb, p);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

static int proj_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// y (M x N, ld ldy) = x (M x K, ld ldx) . w (N x K, ld ldw)^T + bias
int launch_proj_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, int M, int N, int K, float* y,
                    int ldy, cudaStream_t stream) {
    CUtensorMap ma, mb;
    if (int rc = make_f32_map(&ma, x, M, K, ldx, kPK, kPT, true)) return rc;
    if (int rc = make_f32_map(&mb, w, N, K, ldw, kPK, kPT, true)) return rc;
    ProjParams p{};
    p.Mo = M; p.No = N; p.Kc = K;
    p.chunks_per_split = (K + kPK - 1) / kPK;
    p.ldo = ldy;
    p.bias = bias;
    p.out = y;
    return launch_proj_t<false, false>(ma, mb, p, dim3((N + kPT - 1) / kPT, (M + kPT - 1) / kPT, 1), stream);
}

// dx (M x K, ld lddx) = dy (M x N, ld lddy) . w (N x K, ld ldw)
int launch_proj_bwd_x(const float* dy, int lddy, const float* w, int ldw, int M, int N, int K, float* dx, int lddx,
                      cudaStream_t stream) {
    CUtensorMap ma, mb;
    if (int rc = make_f32_map(&ma, dy, M, N, lddy, kPK, kPT, true)) return rc;
    if (int rc = make_f32_map(&mb, w, N, K, ldw, kPT, kPK, false)) return rc;       // [32 n rows x 128 k]
    ProjParams p{};
    p.Mo = M; p.No = K; p.Kc = N;
    p.chunks_per_split = (N + kPK - 1) / kPK;
    p.ldo = lddx;
    p.out = dx;
    return launch_proj_t<false, true>(ma, mb, p, dim3((K + kPT - 1) / kPT, (M + kPT - 1) / kPT, 1), stream);
}

// dw (N x K, ld lddw) += dy (M x N)^T . x (M x K); db (N) += column sums of dy.  dw / db zero-filled by the caller.
int launch_proj_bwd_w(const float* dy, int lddy, const float* x, int ldx, int M, int N, int K, float* dw, int lddw,
                      float* db, cudaStream_t stream) {
    CUtensorMap ma, mb;
    if (int rc = make_f32_map(&ma, dy, M, N, lddy, kPT, kPK, false)) return rc;      // [32 m rows x 128 n]
    if (int rc = make_f32_map(&mb, x, M, K, ldx, kPT, kPK, false)) return rc;        // [32 m rows x 128 k]
    ProjParams p{};
    p.Mo = N; p.No = K; p.Kc = M;
    const int tiles = ((N + kPT - 1) / kPT) * ((K + kPT - 1) / kPT);
    const int chunks = (M + kPK - 1) / kPK;
    // split the contraction so that the grid fills the device about twice, with at least 8 chunks per CTA
    int splits = max(1, min((2 * proj_sm_count() + tiles - 1) / tiles, (chunks + 7) / 8));
    p.chunks_per_split = (chunks + splits - 1) / splits;
    splits = (chunks + p.chunks_per_split - 1) / p.chunks_per_split;
    p.ldo = lddw;
    p.accumulate = 1;
    p.out = dw;
    if (int rc = launch_proj_t<true, true>(ma, mb, p, dim3((K + kPT - 1) / kPT, (N + kPT - 1) / kPT, splits), stream))
        return rc;
    if (db) {
        const int rpb = max(64, (M + 63) / 64);
        colsum_kernel<<<dim3((N + 127) / 128, (M + rpb - 1) / rpb), 256, 0, stream>>>(dy, M, N, lddy, rpb, db);
        TTX_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

}  // namespace ttx
