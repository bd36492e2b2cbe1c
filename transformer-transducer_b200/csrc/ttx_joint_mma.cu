// The one dense contraction of the path -- the joint's H -> V projection -- on tcgen05 tensor cores.
//
// One kernel template, three modes, all "stationary tile x streamed tiles" with the 16-bit operands
// staged by TMA into 128B-swizzled shared memory and fp32 accumulators in TMEM:
//
//   FWD  X = A16 tile (128 lattice rows x H), Y = W16 streamed over V.
//        S = X . Y^T  -> epilogue: + bias, online log-sum-exp over V, pick blank / label logits.
//        Writes 3 floats per lattice cell; the (B,T,U1,V) logits never leave the SM.
//        (replaces the Linear -> log_softmax hand-off, /root/reference/tt/model.py:37 + train.py:53)
//   DA   X = A16 tile, Y = W16.  Recomputes S, turns it into P' = softmax - transition posteriors
//        (16-bit, shared memory) and accumulates G = P' . W16[:, half] in TMEM -> dL/dA tile.
//   DW   X = W16 tile (128 vocab rows), Y = A16 streamed over lattice rows.  Recomputes S^T, builds
//        Q = gamma * P' and accumulates G = Q . A16[:, half] in TMEM -> dL/dW_out tile (+ dL/db_out).
//        (DA + DW replace autograd through log_softmax + project_layer, train.py:58)
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + single-thread MMA issuer,
// warps 2..9 = epilogue (two per TMEM lane quarter, 64 accumulator columns each).
#include "ttx_common.cuh"

namespace ttx {

enum { MODE_FWD = 0, MODE_DA = 1, MODE_DW = 2, MODE_FG = 3 };

struct MmaParams {
    int H, NKC;            // joint width, H / 64
    int V;                 // vocabulary size (valid rows of W16)
    int n_halves, HH;      // backward: H is processed in n_halves column slabs of HH
    int NGCL;              // backward: 64-column G chunks this CTA loads per stream tile (HH / 64 / CG)
    int GCH;               // backward: chunks per G-pass MMA group (CG = 2: all of them)
    int NS;                // ring stages
    int blank;
    int splits;            // DW: lattice-row splits
    const int* meta;       // tile table
    const float* bias2;    // (Vpad) b_out * log2(e), -inf for v >= V
    const float* scal;     // [0] w_scale, [1] 1 / w_scale, [2] gmax, [3] != 0 if some grad_costs[b] < 0
    const int* row_label;  // (rows) label emitted from the cell's u, -1 if none / padding
    float* lse;            // FWD out (rows)
    float* lpb;            // FWD out (rows) log p(blank)
    float* lpl;            // FWD out (rows) log p(label)
    const float4* rowmeta; // BWD in (rows): {lse, p_blank - rb, p_label - rl, gamma * g_b / gmax}
    float* dA;             // DA out (rows x H)
    float* dW;             // DW out (V x H), accumulated with red.add
    float* db;             // DW out (V)
    uint8_t* scratch;      // pair kernel: flags (64 KiB) of the P' scratch matrix (replay), or null
    int scr_rows;          // rows R of the P' matrix; it is stored in 64-column blocks, [cols / 64][R][64]
};

constexpr int kEpiWarps = 8;                       // two per TMEM lane quarter
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 64 + kEpiThreads;         // + TMA producer warp + MMA issuer warp
constexpr int kNumBars = 52;
constexpr int kPB = 2;                             // pair kernel: P' sub-tile buffers (16 KiB each)
constexpr int kEpiBarrier = 1;                     // named barrier id for the epilogue warps
constexpr int kMaxStages = 16;

// Position in the pair kernel's ring of P' sub-tile buffers: sub-pass n uses buffer n % kPB; `phase` = parity of its
// use count n / kPB.  The MMA issuer and every epilogue thread step through the sub-passes in the same order.
struct PRing {
    int buf = 0;
    uint32_t phase = 0;
    __device__ __forceinline__ void advance() {
        if (++buf == kPB) {
            buf = 0;
            phase ^= 1;
        }
    }
};

constexpr int kPairEpiWarps = 8;                   // pair kernel: two epilogue warps per TMEM lane quarter
constexpr int kPairEpiThreads = kPairEpiWarps * 32;
constexpr int kPairThreads = kPairEpiThreads + 128;   // + one control warpgroup: TMA producer, MMA issuer, two idle warps
constexpr int kPairCtrlRegs = 72;                  // setmaxnreg: the control warpgroup gives its registers ...
constexpr int kPairEpiRegs = 216;                  // ... to the epilogue warps (12 warps x 168 = 4 x 72 + 8 x 216)
// The two single-thread control warps get the HIGHEST warp ids: the issue arbiter prefers high warp ids, and a TMA
// producer / MMA issuer starved by spinning epilogue warps stalls the whole pipeline (measured: 1.5x slower with ids 0, 1).
constexpr int kPairProducerWarp = kPairEpiWarps;
constexpr int kPairMmaWarp = kPairEpiWarps + 1;
constexpr int kPairWatchWarp = kPairEpiWarps + 2;
constexpr int kQuarterBarrier = 2;                 // named barriers 2..5: the four epilogue warps of lane quarter q

__device__ __forceinline__ void pair_epi_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(kPairEpiThreads) : "memory");
}
__device__ __forceinline__ void quarter_sync(int q) {
    asm volatile("bar.sync %0, %1;" ::"r"(kQuarterBarrier + q), "n"(kPairEpiThreads / 4) : "memory");
}
// Barrier over the epilogue threads of lane quarter q that also ORs a predicate across them.
__device__ __forceinline__ bool quarter_any(int q, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred pin, pout;\n\t"
        "setp.ne.b32 pin, %2, 0;\n\t"
        "bar.red.or.pred pout, %1, %3, pin;\n\t"
        "selp.u32 %0, 1, 0, pout;\n\t}"
        : "=r"(r)
        : "r"(kQuarterBarrier + q), "r"((uint32_t)pred), "n"(kPairEpiThreads / 4)
        : "memory");
    return r != 0;
}

__device__ __forceinline__ void epi_sync() {
    asm volatile("bar.sync %0, %1;" ::"n"(kEpiBarrier), "n"(kEpiThreads) : "memory");
}

// byte offset of 16-bit element (row, col) inside the 128 x 128 K-major / MN-major tile (2 blocks of 64 columns)
__device__ __forceinline__ uint32_t ptile_off(int row, int col) {
    return (col >> 6) * kChunkBytes + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
}

template <bool BF16>
__device__ __forceinline__ uint16_t to16(float x) {
    if (BF16) return __bfloat16_as_ushort(__float2bfloat16_rn(x));
    return __half_as_ushort(__float2half_rn(x));
}

template <int CG>
__device__ __forceinline__ void bwait(uint32_t bar, uint32_t parity) {
    mbar_wait(bar, parity);
}

// CG = 1: one CTA per 128-row stationary tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2): each CTA
// keeps its own stationary tile (MMA M = 256), the streamed operand is split between the two CTAs' shared memory
// (half the TMA bytes and half the operand reads per SM), the leader CTA issues every MMA and its commits are
// multicast to both CTAs' barriers.
template <int MODE, bool BF16, int CG>
__global__ void __launch_bounds__(kThreads, 1)
joint_mma_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                 const MmaParams p) {
    constexpr bool BWD = (MODE != MODE_FWD);
    constexpr int NT = (CG == 2 && !BWD) ? 256 : 128;   // columns of one S accumulator = stream rows per step
    constexpr int SR = NT / CG;                          // stream rows this CTA loads per S chunk
    constexpr int STAGE = SR * 128;                      // bytes of one ring stage
    constexpr int GST = kChunkBytes / STAGE;             // ring stages per 128-row G chunk
    constexpr int NG = NT / 64;                          // 32-column groups per epilogue thread
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    const bool leader = (rank == 0);
    const int n_tiles = p.meta[0];
    // ---- work assignment (uniform per CTA pair; pairs without work leave before touching TMEM / the cluster barrier)
    int j0, j1;
    bool valid_x = true;
    if (MODE == MODE_DW) {
        const int per = (n_tiles + p.splits - 1) / p.splits;
        j0 = blockIdx.z * per;
        j1 = min(n_tiles, j0 + per);
    } else {
        const int first = (CG == 2) ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x;
        if (first >= n_tiles) return;
        valid_x = (int)blockIdx.x < n_tiles;
        j0 = 0;
        j1 = (p.V + NT - 1) / NT;
    }
    if (j0 >= j1) return;
    const int x_row0 = blockIdx.x * kTile;
    const int half = BWD ? blockIdx.y : 0;
    const int n_iter = j1 - j0;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) {   // 128B-swizzle atoms need 1024-byte alignment
        if (threadIdx.x == 0) printf("ttx: dynamic shared memory is not 1024-byte aligned (0x%x)\n", smem_base);
        __trap();
    }
    const uint32_t sX = smem_base;                                  // NKC chunks
    const uint32_t sP = sX + p.NKC * kChunkBytes;                   // BWD: 2 chunks (128 x 128 16-bit)
    const uint32_t sRing = sP + (BWD ? 2 * kChunkBytes : 0);        // NS stages
    const uint32_t sBar = sRing + p.NS * STAGE;                     // barriers
    const uint32_t sTmemPtr = sBar + kNumBars * 8;
    const uint32_t sKbuf = sTmemPtr + 16;                           // DW: 2 x 2 x 128 floats of per-column constants
    uint8_t* smem_gen = smem_raw;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (sTmemPtr - smem_base));

    // barrier map (same offsets in both CTAs of a pair)
    const uint32_t bar_xfull = sBar;
    auto bar_full = [&](int s) { return sBar + 8 * (1 + s); };
    auto bar_empty = [&](int s) { return sBar + 8 * (17 + s); };
    auto bar_sfull = [&](int b) { return sBar + 8 * (33 + b); };
    auto bar_sempty = [&](int b) { return sBar + 8 * (35 + b); };
    const uint32_t bar_pfull = sBar + 8 * 37;
    const uint32_t bar_pempty = sBar + 8 * 38;
    const uint32_t bar_gfull = sBar + 8 * 39;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = (BWD || NT == 256) ? 512 : 256;
    constexpr uint32_t kEpiArrivals = (CG == 2) ? 2 * kEpiWarps : kEpiThreads;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapY);
        mbar_init(bar_xfull, 1);
        for (int s = 0; s < p.NS; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_sfull(b), 1);
            mbar_init(bar_sempty(b), kEpiArrivals);
        }
        mbar_init(bar_pfull, kEpiArrivals);
        mbar_init(bar_pempty, 1);
        mbar_init(bar_gfull, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(sTmemPtr, kTmemCols);
        else tmem_alloc(sTmemPtr, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // peer barriers are initialised before any remote arrive / TMA signal
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    const uint32_t tmem_G = tmem_base + 256;

    // epilogue -> MMA-issuer arrival (the issuer lives in the leader CTA)
    auto epi_arrive = [&](uint32_t bar) {
        if (CG == 2) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar, 0);
        } else {
            mbar_arrive(bar);
        }
    };

    if (warp == 0) {
        // =========================================================== TMA producer
        if (lane == 0) {
            if (CG == 2) {
                if (leader) mbar_arrive_expect_tx(bar_xfull, 2 * p.NKC * kChunkBytes);
                for (int c = 0; c < p.NKC; ++c)
                    tma_load_2d_pair(sX + c * kChunkBytes, &mapX, bar_xfull, c * kKC, x_row0);
            } else {
                mbar_arrive_expect_tx(bar_xfull, p.NKC * kChunkBytes);
                for (int c = 0; c < p.NKC; ++c) tma_load_2d(sX + c * kChunkBytes, &mapX, bar_xfull, c * kKC, x_row0);
            }
            Ring r;
            auto load_stage = [&](int col, int row) {      // one [SR rows x 64 cols] box into the next ring stage
                bwait<CG>(bar_empty(r.stage), r.phase ^ 1);
                if (CG == 2) {
                    if (leader) mbar_arrive_expect_tx(bar_full(r.stage), 2 * STAGE);
                    tma_load_2d_pair(sRing + r.stage * STAGE, &mapY, bar_full(r.stage), col, row);
                } else {
                    mbar_arrive_expect_tx(bar_full(r.stage), STAGE);
                    tma_load_2d(sRing + r.stage * STAGE, &mapY, bar_full(r.stage), col, row);
                }
                r.advance(p.NS);
            };
            auto load_S = [&](int j) {
                for (int c = 0; c < p.NKC; ++c) load_stage(c * kKC, j * NT + (int)rank * SR);
            };
            auto load_G = [&](int j) {                     // this CTA's share of the h columns, all 128 stream rows
                const int col0 = half * p.HH + (int)rank * (p.HH / CG);
                for (int c = 0; c < p.NGCL; ++c)
                    for (int g = 0; g < GST; ++g) load_stage(col0 + c * kKC, j * kTile + g * SR);
            };
            if (!BWD) {
                for (int j = j0; j < j1; ++j) load_S(j);
            } else {
                load_S(j0);
                for (int i = 0; i < n_iter; ++i) {
                    if (i + 1 < n_iter) load_S(j0 + i + 1);
                    load_G(j0 + i);
                }
            }
        }
    } else if (warp == 1) {
        // =========================================================== MMA issuer (one thread of the leader CTA)
        if (lane == 0 && leader) {
            constexpr int fmt = BF16 ? 1 : 0;
            const uint32_t idescS = make_idesc(fmt, 0, 0, 128 * CG, NT);
            const uint32_t idescG = make_idesc(fmt, 0, 1, 128 * CG, 64 * p.GCH * CG);
            const int gstages = p.GCH * GST;
            auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
                if (CG == 2) umma_f16_ss_pair(d, da, db, idesc, acc);
                else umma_f16_ss(d, da, db, idesc, acc);
            };
            auto commit = [&](uint32_t bar) {
                if (CG == 2) umma_commit_pair(bar);
                else umma_commit(bar);
            };
            Ring r;
            bwait<CG>(bar_xfull, 0);
            auto issue_S = [&](int idx) {
                const int buf = idx & 1;
                bwait<CG>(bar_sempty(buf), ((idx >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + buf * NT;
                for (int c = 0; c < p.NKC; ++c) {
                    bwait<CG>(bar_full(r.stage), r.phase);
                    tc_fence_after();
                    const uint32_t a = sX + c * kChunkBytes;
                    const uint32_t b = sRing + r.stage * STAGE;
#pragma unroll
                    for (int k = 0; k < 4; ++k) mma(d, desc_kmajor(a, k), desc_kmajor(b, k), idescS, (c | k) != 0);
                    commit(bar_empty(r.stage));
                    r.advance(p.NS);
                }
                commit(bar_sfull(buf));
            };
            auto issue_G = [&](int idx) {
                bwait<CG>(bar_pfull, idx & 1);
                tc_fence_after();
                for (int g = 0; g < p.NGCL; g += p.GCH) {
                    for (int s = 0; s < gstages; ++s) bwait<CG>(bar_full(r.stage + s), r.phase);
                    tc_fence_after();
                    const uint32_t d = tmem_G + g * 64;
                    const uint32_t b = sRing + r.stage * STAGE;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        mma(d, desc_kmajor(sP + (k >> 2) * kChunkBytes, k & 3), desc_mnmajor(b, k, kChunkBytes), idescG,
                            (idx | k) != 0);
                    for (int s = 0; s < gstages; ++s) commit(bar_empty(r.stage + s));
                    r.advance(p.NS, gstages);
                }
                commit(bar_pempty);
            };
            if (!BWD) {
                for (int i = 0; i < n_iter; ++i) issue_S(i);
            } else {
                issue_S(0);
                for (int i = 0; i < n_iter; ++i) {
                    if (i + 1 < n_iter) issue_S(i + 1);
                    issue_G(i);
                }
                commit(bar_gfull);
            }
        }
    } else {
        // =========================================================== epilogue warps (256 threads)
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int ch = (warp - 2) >> 2;               // which half of the accumulator columns
        const int row = q * 32 + lane;                // accumulator row handled by this thread
        const int et = threadIdx.x - 64;              // 0..255 among the epilogue threads
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const float inv_ws = p.scal[1];
        const float c1 = inv_ws * kLog2e;
        uint32_t acc[32];

        if (MODE == MODE_FWD) {
            const int grow = x_row0 + row;
            const int label = valid_x ? p.row_label[grow] : -1;
            // log2-domain running reference mref and sum of 2^(y - mref); mref only moves when a logit exceeds it
            // by more than 2^40, so the usual step is one ex2(fma) + add per element.
            float mref = -INFINITY, ssum = 0.f, zb = 0.f, zl = 0.f;
            for (int i = 0; i < n_iter; ++i) {
                const int buf = i & 1;
                const int v0 = (j0 + i) * NT + ch * (NT / 2);
                bwait<CG>(bar_sfull(buf), (i >> 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int g = 0; g < NG; ++g) {
                    const int vb = v0 + g * 32;
                    tmem_ld32(tmem_base + lane_addr + buf * NT + ch * (NT / 2) + g * 32, acc);
                    float y[32];
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias2 + vb);
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float4 bv = __ldg(b4 + e);
                        y[4 * e + 0] = bv.x; y[4 * e + 1] = bv.y; y[4 * e + 2] = bv.z; y[4 * e + 3] = bv.w;
                    }
                    tmem_ld_wait();
                    float gmax = -INFINITY;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        y[e] = fmaf(__uint_as_float(acc[e]), c1, y[e]);
                        gmax = fmaxf(gmax, y[e]);
                    }
                    if (gmax > mref + 40.f) {          // rare: move the reference (first group, or a big outlier)
                        ssum *= ex2f(mref - gmax);     // ex2(-inf) = 0 on the first group
                        mref = gmax;
                    }
                    if (mref > -INFINITY) {            // false only while every column seen so far is padding
                        float part = 0.f;
#pragma unroll
                        for (int e = 0; e < 32; ++e) part += ex2f(y[e] - mref);
                        ssum += part;
                    }
                    if (p.blank >= vb && p.blank < vb + 32) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) zb = (vb + e == p.blank) ? y[e] : zb;
                    }
                    if (label >= vb && label < vb + 32) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) zl = (vb + e == label) ? y[e] : zl;
                    }
                }
                tc_fence_before();
                epi_arrive(bar_sempty(buf));
            }
            // combine the two column halves of each row through (now idle) ring memory
            float4* xch = reinterpret_cast<float4*>(smem_gen + (sRing - smem_base));
            if (ch == 1) xch[row] = make_float4(mref, ssum, zb, zl);
            epi_sync();
            if (ch == 0 && valid_x) {
                const float4 o = xch[row];
                const float mn = fmaxf(mref, o.x);
                const float tot = ssum * ex2f(mref - mn) + ((o.x > -INFINITY) ? o.y * ex2f(o.x - mn) : 0.f);
                const float lse2 = mn + lg2f(tot);
                zb = ((p.blank % NT) < NT / 2) ? zb : o.z;          // blank / label columns live in exactly one half
                if (label >= 0) zl = ((label % NT) < NT / 2) ? zl : o.w;
                p.lse[grow] = lse2 * kLn2;
                p.lpb[grow] = (zb - lse2) * kLn2;
                p.lpl[grow] = (label >= 0) ? (zl - lse2) * kLn2 : 0.f;
            }
        } else {
            // ---- backward: S -> 16-bit gradient operand in shared memory (128B swizzle), 64 columns per thread
            uint8_t* sP_gen = smem_gen + (sP - smem_base);
            float* kbuf = reinterpret_cast<float*>(smem_gen + (sKbuf - smem_base));
            const float pscale = BF16 ? 1.0f : kPScale;
            const float lg_scale = BF16 ? 0.0f : 12.0f;          // log2(kPScale)
            const bool any_neg = p.scal[3] != 0.f;
            float4 rm = make_float4(INFINITY, 0.f, 0.f, 0.f);
            int label = -1;
            float krow = 0.f, db_acc = 0.f;
            int vrow = 0;
            if (MODE == MODE_DA) {
                if (valid_x) {
                    rm = p.rowmeta[x_row0 + row];
                    label = p.row_label[x_row0 + row];
                }
                krow = fmaf(rm.x, -kLog2e, lg_scale);              // -lse2 + log2(scale); -inf for padding rows
            } else {
                vrow = x_row0 + row;
                krow = __ldg(p.bias2 + vrow);                      // -inf for vocabulary padding rows
            }
            for (int i = 0; i < n_iter; ++i) {
                const int buf = i & 1;
                const int c0 = (j0 + i) * kTile;   // first vocab id (DA) / lattice row (DW) of this stream tile
                float4 cm = make_float4(0.f, 0.f, 0.f, 0.f);
                int clabel = -1;
                if (MODE == MODE_DW) {
                    // per-column constant k_m = -lse2_m + log2(|w_m| * scale): Q = +-2^(acc*c1 + bias2_v + k_m)
                    if (et < kTile) {
                        cm = __ldg(p.rowmeta + c0 + et);
                        clabel = __ldg(p.row_label + c0 + et);
                        const float k = fmaf(cm.x, -kLog2e, lg2f(fabsf(cm.w)) + lg_scale);
                        kbuf[buf * 2 * kTile + et] = k;
                        kbuf[buf * 2 * kTile + kTile + et] = (cm.w < 0.f) ? -1.f : 1.f;
                    }
                    epi_sync();
                }
                bwait<CG>(bar_sfull(buf), (i >> 1) & 1);
                tc_fence_after();
                uint32_t packed[32];
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int cb = ch * 64 + g * 32;           // first accumulator column of this group
                    tmem_ld32(tmem_base + lane_addr + buf * 128 + cb, acc);
                    float kc[32];
                    if (MODE == MODE_DA) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.bias2 + c0 + cb);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 bv = __ldg(b4 + e);
                            kc[4 * e + 0] = bv.x; kc[4 * e + 1] = bv.y; kc[4 * e + 2] = bv.z; kc[4 * e + 3] = bv.w;
                        }
                    } else {
                        const float4* k4 = reinterpret_cast<const float4*>(kbuf + buf * 2 * kTile + cb);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 kv = k4[e];
                            kc[4 * e + 0] = kv.x; kc[4 * e + 1] = kv.y; kc[4 * e + 2] = kv.z; kc[4 * e + 3] = kv.w;
                        }
                    }
                    tmem_ld_wait();
                    float val[32];
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        val[e] = ex2f(fmaf(__uint_as_float(acc[e]), c1, kc[e] + krow));
                    if (MODE == MODE_DW) {
                        if (any_neg) {
                            const float* sg = kbuf + buf * 2 * kTile + kTile + cb;
#pragma unroll
                            for (int e = 0; e < 32; ++e) val[e] *= sg[e];
                        }
#pragma unroll
                        for (int e = 0; e < 32; ++e) db_acc += val[e];
                    }
#pragma unroll
                    for (int e = 0; e < 16; ++e) packed[g * 16 + e] = pack16<BF16>(val[2 * e], val[2 * e + 1]);
                }
                // S buffer is free again as soon as it sits in registers
                tc_fence_before();
                epi_arrive(bar_sempty(buf));
                // wait until the previous G pass has finished reading the P tile, then overwrite it
                bwait<CG>(bar_pempty, (i & 1) ^ 1);
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {       // 8 chunks of 8 values (16 B) = this thread's 64 columns
                    uint4 v4 = make_uint4(packed[cc * 4 + 0], packed[cc * 4 + 1], packed[cc * 4 + 2], packed[cc * 4 + 3]);
                    *reinterpret_cast<uint4*>(sP_gen + ch * kChunkBytes + row * 128 + ((cc ^ (row & 7)) << 4)) = v4;
                }
                // sparse corrections: the blank and label entries are p - rb / p - rl, with p = exp(lp) from the
                // forward pass (rowmeta .y / .z), written exactly instead of being carried through the dense loop
                if (MODE == MODE_DA) {
                    const int cbl = p.blank - c0, clb = label - c0;
                    if (cbl >= ch * 64 && cbl < ch * 64 + 64)
                        *reinterpret_cast<uint16_t*>(sP_gen + ptile_off(row, cbl)) = to16<BF16>(rm.y * pscale);
                    if (clb >= ch * 64 && clb < ch * 64 + 64)
                        *reinterpret_cast<uint16_t*>(sP_gen + ptile_off(row, clb)) = to16<BF16>(rm.z * pscale);
                } else {
                    epi_sync();                        // column owners patch rows written by other threads
                    if (et < kTile) {
                        const int rbl = p.blank - x_row0, rlb = clabel - x_row0;
                        if (rbl >= 0 && rbl < kTile)
                            *reinterpret_cast<uint16_t*>(sP_gen + ptile_off(rbl, et)) = to16<BF16>(cm.y * cm.w * pscale);
                        if (rlb >= 0 && rlb < kTile)
                            *reinterpret_cast<uint16_t*>(sP_gen + ptile_off(rlb, et)) = to16<BF16>(cm.z * cm.w * pscale);
                    }
                }
                fence_proxy_async_smem();
                epi_arrive(bar_pfull);
            }
            // ---- final: G (128 x HH fp32 in TMEM) -> global; the two column halves split the HH columns
            bwait<CG>(bar_gfull, 0);
            tc_fence_after();
            const float gmax = p.scal[2];
            const int ngrp = p.HH / 32;
            if (MODE == MODE_DA) {
                const float f = rm.w * gmax * inv_ws / pscale;
                float* dst = p.dA + (size_t)(x_row0 + row) * p.H + half * p.HH;
                for (int cc = ch; cc < ngrp; cc += 2) {
                    tmem_ld32(tmem_G + lane_addr + cc * 32, acc);
                    tmem_ld_wait();
                    if (valid_x) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            float4 o = make_float4(__uint_as_float(acc[e]) * f, __uint_as_float(acc[e + 1]) * f,
                                                   __uint_as_float(acc[e + 2]) * f, __uint_as_float(acc[e + 3]) * f);
                            *reinterpret_cast<float4*>(dst + cc * 32 + e) = o;
                        }
                    }
                }
            } else {
                const float f = gmax / pscale;
                const bool ok = vrow < p.V;
                float* dst = p.dW + (size_t)vrow * p.H + half * p.HH;
                for (int cc = ch; cc < ngrp; cc += 2) {
                    tmem_ld32(tmem_G + lane_addr + cc * 32, acc);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) atomicAdd(dst + cc * 32 + e, __uint_as_float(acc[e]) * f);
                    }
                }
                // dense part of db: sum_m w_m * softmax(m, v); the sparse -rb / -rl terms are added by
                // grad_prep_kernel.  (the patched entries above do not enter db_acc.)
                if (ok && half == 0) atomicAdd(p.db + vrow, db_acc * gmax / pscale);
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();     // the leader's MMAs read the peer's shared memory / write its TMEM
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ===========================================================================================================
// Backward pair kernel (v3).  Measured on B200: in SS mode every tcgen05.mma (K = 16) costs ~140 cycles however
// small N is, so the backward modes above (S pass with N = 128, G pass with an MN-major B operand at ~250
// cycles / instruction) are instruction-issue bound at ~45 % of the tensor rate.  This kernel does the same math
// with fewer, larger instructions: the pair streams 256-row tiles (S pass: M = 256, N = 256), keeps ONE 256-column
// S accumulator whose read-out overlaps the previous tile's G pass, feeds the G pass (M = 256, N = HH) from the
// 128-column P' tile in two sub-passes, and takes the G-pass B operand K-major from a transposed copy of the
// streamed matrix (W16^T for dA, A16^T for dW).  48 instructions per 256 stream rows instead of 80.
template <int MODE, bool BF16>
__global__ void __launch_bounds__(kPairThreads, 1)
joint_bwd_pair_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
                      const __grid_constant__ CUtensorMap mapYT, const __grid_constant__ CUtensorMap mapScr,
                      const MmaParams p) {
    constexpr int NT = 256;                 // stream rows per step (pair-wide) = S accumulator columns
    constexpr int SR = 128;                 // stream rows this CTA loads per S chunk
    constexpr int STAGE = kChunkBytes;      // 16 KiB ring stages
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int n_tiles = p.meta[0];
    // Persistent: the grid is one CTA pair per SM pair and each pair walks work units one after the other, all slabs of
    // a unit in turn, so the next item's stationary tile loads behind the last G sub-passes, its first S passes run
    // behind the read-out of G, and TMEM / barriers are set up once.  FG / DA: unit = tile pair, streams the whole
    // vocabulary.  DW: unit = (vocabulary tile pair, lattice-row split), consecutive units share the split.
    constexpr bool PERSIST = true;
    int j0 = 0, n_iter = 0;                           // the unit's stream chunks [j0, j0 + n_iter)
    int n_units, n_vq = 1, per = 0, n_st = 0;
    const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
    if (MODE == MODE_DW) {
        n_st = (n_tiles + 1) / 2;
        per = (n_st + p.splits - 1) / p.splits;
        n_vq = ((p.V + kTile - 1) / kTile + 1) / 2;
        n_units = n_vq * p.splits;
    } else {
        n_units = (n_tiles + 1) >> 1;
        n_iter = (p.V + NT - 1) / NT;
    }
    if (unit0 >= n_units) return;
    // every role calls this at the top of its unit loop; false = no work left (empty trailing splits)
    auto begin_unit = [&](int unit) {
        if (MODE == MODE_DW) {
            j0 = (unit / n_vq) * per;
            n_iter = min(n_st, j0 + per) - j0;
        }
        return n_iter > 0;
    };
    const int n_hl = p.n_halves;                      // items per unit: one per slab
    auto unit_tile = [&](int unit) { return (MODE == MODE_DW ? (unit % n_vq) * 2 : unit * 2) + (int)rank; };
    auto slab_of = [&](int hh) { return hh; };
    // Forward+gradient with a scratch area (REPLAY): the first slab of a unit also sends every P' sub-tile to a scratch
    // matrix in global memory (TMA store from the shared-memory buffer the G sub-pass reads); the second slab then
    // needs neither S passes nor exponentials -- it streams P' back as the A operand next to the W16^T chunks.  If the
    // running reference moved after the unit's first tile (rare), the stored sub-tiles carry mixed scales: the unit is
    // flagged and its second slab recomputes everything as without the scratch area.
    const bool rp = (MODE == MODE_FG || MODE == MODE_DW) && p.scratch != nullptr && p.n_halves == 2 && p.NS <= 4 &&
                    p.NKC + kPB <= 12;
    // The replay streams two operands and touches neither the X tile nor the P' buffers: their shared memory (contiguous,
    // 10 x 16 KiB) is its ring, with its own barriers (slots kRB.. of the full / empty arrays) and its own position.
    constexpr int kRB = 4;                            // first barrier slot of the replay ring (the S / G ring uses < 4)
    const int nrs = p.NKC + kPB;                      // replay ring stages
    // Flags (first 64 KiB of the scratch area, zeroed by the launcher): one word per CTA pair, stamped with the pair's
    // unit counter + 1 when the unit's stored sub-tiles carry mixed scales.
    volatile int* const flags = reinterpret_cast<volatile int*>(p.scratch);
    auto unit_clean = [&](int unit, int uidx) { return flags[blockIdx.x >> 1] != uidx + 1; };
    auto scr_row = [&](int x_row0) { return (int)blockIdx.x * kTile; };   // the CTA's rows of the scratch matrix
    const int hh2 = p.HH / 2;               // G columns (N rows of the K-major B operand) held by this CTA

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) {
        if (threadIdx.x == 0) printf("ttx: dynamic shared memory is not 1024-byte aligned (0x%x)\n", smem_base);
        __trap();
    }
    const uint32_t sX = smem_base;
    const uint32_t sP = sX + p.NKC * kChunkBytes;                   // kPB x [128 x 64] 16-bit, K-major: P' sub-tile ring
    const uint32_t sRing = sP + kPB * kChunkBytes;
    const uint32_t sBar = sRing + p.NS * STAGE;
    const uint32_t sTmemPtr = sBar + kNumBars * 8;
    const uint32_t sKbuf = sTmemPtr + 16;                           // DW: 256 exponent offsets + 256 signs
    const uint32_t sWatch = sTmemPtr + 8;                           // barrier watcher's event counter
    uint8_t* smem_gen = smem_raw;
    volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (sTmemPtr - smem_base));

    auto bar_xfull = [&](int k) { return sBar + 8 * (kNumBars - 8 + k); };   // one per 64-column chunk of the X tile
    auto bar_full = [&](int s) { return sBar + 8 * (1 + s); };
    auto bar_empty = [&](int s) { return sBar + 8 * (17 + s); };
    const uint32_t bar_sfull = sBar + 8 * 33;
    const uint32_t bar_sempty = sBar + 8 * 34;
    const uint32_t bar_gfull = sBar + 8 * 35;
    const uint32_t bar_xempty = sBar;                               // the item's last S pass has read the X tile
    const uint32_t bar_gempty = sBar + 8 * 43;                      // the epilogue has read the item's G out of TMEM
    auto bar_pfull = [&](int b) { return sBar + 8 * (36 + b); };     // one barrier pair per P' sub-tile buffer
    auto bar_pempty = [&](int b) { return sBar + 8 * (36 + kPB + b); };
    static_assert(kPB == 2, "barrier slots 40..42 are used below");
    auto bar_pwritten = [&](int b) { return sBar + 8 * (40 + b); };  // REPLAY: this CTA's epilogue warps wrote buffer b
    const uint32_t bar_unit = sBar + 8 * 42;                         // REPLAY: first slab of the unit is complete

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    constexpr uint32_t kTmemCols = 512;

    if (warp == kPairProducerWarp && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapYT);
        tma_prefetch_desc(&mapScr);
        for (int c = 0; c < 8; ++c) mbar_init(bar_xfull(c), 1);
        *reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base)) = 0;
        for (int s = 0; s < 16; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_sfull, 1);
        mbar_init(bar_sempty, 2 * kPairEpiWarps);       // every epilogue warp of both CTAs
        for (int b = 0; b < kPB; ++b) {
            mbar_init(bar_pfull(b), 2 * kPairEpiWarps);  // every epilogue warp of both CTAs, once per sub-pass
            mbar_init(bar_pempty(b), rp ? 2 : 1);        // the G sub-pass (and the store to the scratch area) have read it
            mbar_init(bar_pwritten(b), kPairEpiWarps);
        }
        mbar_init(bar_unit, 2 * kPairEpiWarps + 1);      // every epilogue warp of both CTAs + this CTA's storer
        mbar_init(bar_gfull, 1);
        mbar_init(bar_xempty, 1);
        mbar_init(bar_gempty, 2 * kPairEpiWarps);
        fence_barrier_init();
    }
    if (warp == kPairMmaWarp) tmem_alloc_pair(sTmemPtr, kTmemCols);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_gen;
    const uint32_t tmem_G = tmem_base + 256;

    auto epi_arrive = [&](uint32_t bar) {
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar, 0);
    };

    if (warp >= kPairEpiWarps) {
        // =========================================================== control warpgroup
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPairCtrlRegs));
      if (warp == kPairProducerWarp) {
        // =========================================================== TMA producer
        if (lane == 0) {
            Ring r;
            auto load_stage = [&](const CUtensorMap* map, int col, int row, int bytes) {
                mbar_wait(bar_empty(r.stage), r.phase ^ 1);
                if (leader) mbar_arrive_expect_tx(bar_full(r.stage), 2 * bytes);
                tma_load_2d_pair(sRing + r.stage * STAGE, map, bar_full(r.stage), col, row);
                r.advance(p.NS);
            };
            auto load_S = [&](int j) {
                for (int c = 0; c < p.NKC; ++c) load_stage(&mapY, c * kKC, j * NT + (int)rank * SR, STAGE);
            };
            auto load_G = [&](int j, int half) {       // K-major B chunks: [hh2 joint columns x 64 stream rows]
                const int h0 = half * p.HH + (int)rank * hh2;
                for (int c = 0; c < 4; ++c) {
                    const int k0 = j * NT + c * kKC;
                    // W16^T (FG / DA) is row-major [H][Vpad]; A16^T (DW) is stored in blocks of 64 lattice rows, [rows / 64][H][64]
                    if (MODE == MODE_DW) load_stage(&mapYT, 0, (k0 / kKC) * p.H + h0, hh2 * 128);
                    else load_stage(&mapYT, k0, h0, hh2 * 128);
                }
            };
            int xt = 0, uidx = 0, it = 0;                       // X tiles loaded, units started, items started
            Ring rr;                                            // replay ring position
            bool replayed = false;                              // the previous item was a replay (its ring is our X tile)
            for (int unit = unit0; unit < n_units; unit += unit_step, ++uidx) {
                if (!begin_unit(unit)) break;
                for (int hh = 0; hh < n_hl; ++hh, ++it) {
                    const int x_row0 = unit_tile(unit) * kTile, half = slab_of(hh);
                    if (rp && hh == 1) {
                        mbar_wait(bar_unit, uidx & 1);
                        if (unit_clean(unit, uidx)) {
                            // replay: per sub-pass one stage of P' (A operand, from the scratch matrix) and one of W16^T;
                            // the ring is the X tile + P' buffers, free once the first slab's last S pass / G sub-pass
                            // have read them (xempty; the epilogue's unit arrival came after its last pempty wait)
                            const int h0 = half * p.HH + (int)rank * hh2;
                            mbar_wait(bar_xempty, (xt - 1) & 1);
                            auto load_rstage = [&](const CUtensorMap* map, int col, int row, int bytes) {
                                mbar_wait(bar_empty(kRB + rr.stage), rr.phase ^ 1);
                                if (leader) mbar_arrive_expect_tx(bar_full(kRB + rr.stage), 2 * bytes);
                                tma_load_2d_pair(sX + rr.stage * STAGE, map, bar_full(kRB + rr.stage), col, row);
                                rr.advance(nrs);
                            };
                            for (int i = 0; i < n_iter; ++i)
                                for (int c = 0; c < 4; ++c) {
                                    load_rstage(&mapScr, 0, (i * 4 + c) * p.scr_rows + scr_row(x_row0), STAGE);
                                    if (MODE == MODE_DW) load_rstage(&mapYT, 0, (((j0 + i) * NT + c * kKC) / kKC) * p.H + h0, hh2 * 128);
                                    else load_rstage(&mapYT, (j0 + i) * NT + c * kKC, h0, hh2 * 128);
                                }
                            replayed = true;
                            continue;
                        }
                    }
                    // the stationary tile arrives chunk by chunk (own barrier each): the first S pass starts after 16 KiB
                    if (xt > 0) mbar_wait(bar_xempty, (xt - 1) & 1);
                    if (replayed) {                             // ... and the replay's MMAs have read the ring stages there
                        mbar_wait(bar_gfull, (it - 1) & 1);
                        replayed = false;
                    }
                    ++xt;
                    for (int c = 0; c < p.NKC; ++c) {
                        if (leader) mbar_arrive_expect_tx(bar_xfull(c), 2 * kChunkBytes);
                        tma_load_2d_pair(sX + c * kChunkBytes, &mapX, bar_xfull(c), c * kKC, x_row0);
                    }
                    // same order as the MMA issuer: S(i+1), then the four G sub-passes of tile i
                    load_S(j0);
                    for (int i = 0; i < n_iter; ++i) {
                        if (i + 1 < n_iter) load_S(j0 + i + 1);
                        load_G(j0 + i, half);
                    }
                }
            }
        }
      } else if (warp == kPairWatchWarp + 1) {
        // =========================================================== REPLAY: storer (each CTA): P' sub-tile -> scratch
        if (lane == 0 && rp) {
            PRing sr;
            int uidx = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step, ++uidx) {
                if (!begin_unit(unit)) break;
                for (int hh = 0; hh < n_hl; ++hh) {
                    if (hh == 1) {
                        mbar_wait(bar_unit, uidx & 1);
                        if (unit_clean(unit, uidx)) continue;
                    }
                    for (int i = 0; i < n_iter; ++i)
                        for (int c = 0; c < 4; ++c) {
                            mbar_wait(bar_pwritten(sr.buf), sr.phase);
                            if (hh == 0) {
                                asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                             ::"l"(reinterpret_cast<uint64_t>(&mapScr)), "r"(sP + sr.buf * kChunkBytes),
                                               "r"(0), "r"((i * 4 + c) * p.scr_rows + scr_row(unit_tile(unit) * kTile))
                                             : "memory");
                                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                            }
                            mbar_arrive(bar_pempty(sr.buf));
                            sr.advance();
                        }
                    if (hh == 0) {
                        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // the stores are complete: replay may load
                        asm volatile("fence.proxy.async;" ::: "memory");
                        mbar_arrive(bar_unit);
                    }
                }
            }
        }
      } else if (warp == kPairWatchWarp) {
        // =========================================================== barrier watcher (leader CTA)
        // A tcgen05.mma is accepted only when the previous one has left the issue stage, so whatever the issuing
        // thread does between two MMAs must fit into the ~128 cycles one of them takes; an mbarrier try_wait alone
        // costs ~90 (measured: 128 -> 142..172 cycles per MMA with one wait per four).  This thread does the waiting
        // instead: it walks the issuer's barriers in the issuer's order and publishes how many have completed; the
        // issuer compares a register copy of that counter and touches shared memory only when it has run out.
        if (lane == 0 && leader) {
            volatile int* ready = reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base));
            int done = 0;
            Ring r;
            PRing pr;
            auto publish = [&]() { *ready = ++done; };
            int it = 0, gs = 0, xt = 0;                         // items done, S passes watched, X tiles loaded
            Ring rr;                                            // replay ring position
            auto watch_S = [&](int idx) {
                mbar_wait(bar_sempty, (gs & 1) ^ 1);
                publish();
                for (int c = 0; c < p.NKC; ++c) {
                    if (idx == 0) mbar_wait(bar_xfull(c), xt & 1);
                    mbar_wait(bar_full(r.stage), r.phase);
                    publish();
                    r.advance(p.NS);
                }
                ++gs;
            };
            auto watch_G = [&](int i) {
                for (int sp = 0; sp < 4; ++sp) {
                    if (i == 0 && sp == 0 && it > 0) mbar_wait(bar_gempty, (it - 1) & 1);   // previous item's G was read out
                    mbar_wait(bar_pfull(pr.buf), pr.phase);
                    mbar_wait(bar_full(r.stage), r.phase);
                    publish();
                    r.advance(p.NS);
                    pr.advance();
                }
            };
            int uidx = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step, ++uidx) {
                if (!begin_unit(unit)) break;
                for (int hh = 0; hh < n_hl; ++hh, ++it) {
                    if (rp && hh == 1) {
                        mbar_wait(bar_unit, uidx & 1);
                        if (unit_clean(unit, uidx)) {
                            for (int i = 0; i < n_iter; ++i)
                                for (int c = 0; c < 4; ++c) {
                                    if (i == 0 && c == 0) mbar_wait(bar_gempty, (it - 1) & 1);
                                    mbar_wait(bar_full(kRB + rr.stage), rr.phase);
                                    rr.advance(nrs);
                                    mbar_wait(bar_full(kRB + rr.stage), rr.phase);
                                    rr.advance(nrs);
                                    publish();
                                }
                            continue;
                        }
                    }
                    watch_S(0);
                    for (int i = 0; i < n_iter; ++i) {
                        if (i + 1 < n_iter) watch_S(i + 1);
                        watch_G(i);
                    }
                    ++xt;
                }
            }
        }
      } else if (warp == kPairMmaWarp) {
        // =========================================================== MMA issuer (leader CTA)
        if (lane == 0 && leader) {
            constexpr int fmt = BF16 ? 1 : 0;
            const uint32_t idescS = make_idesc(fmt, 0, 0, 256, NT);
            const uint32_t idescG = make_idesc(fmt, 0, 0, 256, p.HH);
            // low descriptor words (start address >> 4 | LBO); a 16 KiB chunk / stage is +1024, a 16-element k slice +2
            const uint32_t xlo = desc_lo(sX), plo = desc_lo(sP), rlo = desc_lo(sRing);
            volatile int* ready = reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base));
            int need = 0, have = 0;
            auto wait_event = [&]() {
                ++need;
                if (have < need) {
                    uint32_t spins = 0;
                    while ((have = *ready) < need) {
                        if (++spins > (1u << 26)) {
                            printf("ttx: MMA issuer timed out waiting for event %d (block %d,%d,%d)\n", need, blockIdx.x,
                                   blockIdx.y, blockIdx.z);
                            __trap();
                        }
                    }
                }
            };
            int stage = 0;
            // Issue order S(i+1), G(i,0..3): the epilogue's read-out of S(i) and its exponentials run behind the next S
            // pass; the four G sub-passes of tile i (one 64-column P' sub-tile each) follow as their operands arrive.
            auto issue_S = [&](int idx) {
                wait_event();                                   // the epilogue has read the previous S tile out of TMEM
                tc_fence_after();
                for (int c = 0; c < p.NKC; ++c) {
                    wait_event();                               // ring stage (and, first tile, X chunk c) has landed
                    tc_fence_after();
                    const uint32_t a = xlo + c * 1024, b = rlo + stage * 1024;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss_pair_lo(tmem_base, a + 2 * k, b + 2 * k, idescS, (c | k) != 0);
                    umma_commit_pair(bar_empty(stage));
                    if (++stage == p.NS) stage = 0;
                }
                if (idx == n_iter - 1) umma_commit_pair(bar_xempty);    // the X tile may be replaced by the next item's
                umma_commit_pair(bar_sfull);
            };
            int pb = 0;                                         // P' sub-tile ring: sub-pass n uses buffer n % kPB
            auto issue_G = [&](int idx) {
                for (int sp = 0; sp < 4; ++sp) {
                    wait_event();                               // P' sub-tile stored and ring stage landed
                    tc_fence_after();
                    const uint32_t a = plo + pb * 1024, b = rlo + stage * 1024;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss_pair_lo(tmem_G, a + 2 * k, b + 2 * k, idescG, (idx | sp | k) != 0);
                    umma_commit_pair(bar_empty(stage));
                    umma_commit_pair(bar_pempty(pb));
                    if (++stage == p.NS) stage = 0;
                    if (++pb == kPB) pb = 0;
                }
            };
            int itt = 0, uidx = 0, rstage = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step, ++uidx) {
                if (!begin_unit(unit)) break;
                for (int hh = 0; hh < n_hl; ++hh, ++itt) {
                    bool replay = false;
                    if (rp && hh == 1) {
                        mbar_wait(bar_unit, uidx & 1);
                        replay = unit_clean(unit, uidx);
                    }
                    if (replay) {
                        // G(slab 1) += P'(i, c) . W16^T chunk, both operands from consecutive ring stages
                        for (int i = 0; i < n_iter; ++i)
                            for (int c = 0; c < 4; ++c) {
                                wait_event();
                                tc_fence_after();
                                const int s2 = (rstage + 1 == nrs) ? 0 : rstage + 1;
                                const uint32_t a = xlo + rstage * 1024, b = xlo + s2 * 1024;
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16_ss_pair_lo(tmem_G, a + 2 * k, b + 2 * k, idescG, (i | c | k) != 0);
                                umma_commit_pair(bar_empty(kRB + rstage));
                                umma_commit_pair(bar_empty(kRB + s2));
                                rstage = (s2 + 1 == nrs) ? 0 : s2 + 1;
                            }
                    } else {
                        issue_S(0);
                        for (int i = 0; i < n_iter; ++i) {
                            if (i + 1 < n_iter) issue_S(i + 1);
                            issue_G(i);
                        }
                    }
                    umma_commit_pair(bar_gfull);
                }
            }
        }
      }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPairEpiRegs));
        // =========================================================== epilogue warps: thread = (row, column half ch)
        // Eight warps, two per TMEM lane quarter: thread (row, ch) owns columns ch*32 .. ch*32+31 of each of the tile's
        // four 64-column sub-tiles.  A whole 256-column S tile is processed in registers in ONE pass per warp: one
        // TMEM read-out (the S accumulator is handed back immediately), one reference vote (forward+gradient mode),
        // and two store rounds (sub-tiles 0..2 into buffers that are already free, sub-tile 3 once the first G
        // sub-pass of this tile has released a buffer) -- every synchronisation point costs a few hundred cycles of
        // latency per warp, so there are as few of them per tile as the three P' buffers allow.
        const int q = warp & 3;
        const int ch = warp >> 2;
        const int row = q * 32 + lane;
        const int et = threadIdx.x;                   // 0..255
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const float inv_ws = p.scal[1];
        const float c1 = inv_ws * kLog2e;
        uint8_t* sP_gen = smem_gen + (sP - smem_base);
        float* kbuf = reinterpret_cast<float*>(smem_gen + (sKbuf - smem_base));
        const float pscale = BF16 ? 1.0f : kPScale;
        const float lg_scale = BF16 ? 0.0f : 12.0f;
        const bool any_neg = p.scal[3] != 0.f;
        const int n_valid_rows = n_tiles * kTile;
        PRing pr;                                     // buffer of the tile's sub-tile 0 (sub-pass n = 4 i)
        // this thread's 32 columns of sub-tile g: four 16-byte chunks at swizzled positions (ch*4 + c) ^ (row & 7)
        auto store_p = [&](uint8_t* dstP, const uint32_t* packed) {
            uint8_t* r0 = dstP + row * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(r0 + (((ch * 4 + c) ^ (row & 7)) << 4)) =
                    make_uint4(packed[4 * c + 0], packed[4 * c + 1], packed[4 * c + 2], packed[4 * c + 3]);
        };
        // Store rounds.  `patch(g, dstP)` fixes up the blank / label entries of sub-tile g after the dense stores.
        auto store_rounds = [&](const uint32_t (&packed)[4][16], auto&& patch, const int i) {
            PRing r = pr;
            uint8_t* dst[4];
            int bufs[4];
            // round A: the first kPB sub-tiles go into buffers released by the previous tile's G sub-passes
#pragma unroll
            for (int g = 0; g < kPB; ++g) {
                mbar_wait(bar_pempty(r.buf), r.phase ^ 1);    // the G sub-pass of this buffer's previous use is done
                bufs[g] = r.buf;
                dst[g] = sP_gen + r.buf * kChunkBytes;
                store_p(dst[g], packed[g]);
                r.advance();
            }
            patch(0, kPB, dst);
            fence_proxy_async_smem();
#pragma unroll
            for (int g = 0; g < kPB; ++g) {
                epi_arrive(bar_pfull(bufs[g]));
                if (rp && lane == 0) mbar_arrive(bar_pwritten(bufs[g]));
            }
            // later rounds: one sub-tile each, as this tile's own G sub-passes release the buffers
#pragma unroll
            for (int g = kPB; g < 4; ++g) {
                mbar_wait(bar_pempty(r.buf), r.phase ^ 1);
                bufs[g] = r.buf;
                dst[g] = sP_gen + r.buf * kChunkBytes;
                store_p(dst[g], packed[g]);
                r.advance();
                patch(g, g + 1, dst);
                fence_proxy_async_smem();
                epi_arrive(bar_pfull(bufs[g]));
                if (rp && lane == 0) mbar_arrive(bar_pwritten(bufs[g]));
            }
            pr = r;
        };
        int it = 0, gs0 = 0, uidx = 0;                // items, S passes (accumulator barrier parity), units so far
        float f_keep = 0.f;                           // REPLAY: the row's output scale, from the unit's first slab
        for (int unit = unit0; unit < n_units; unit += unit_step, ++uidx) {
        if (!begin_unit(unit)) break;
        for (int hh = 0; hh < n_hl; ++hh, ++it) {
        const int x_row0 = unit_tile(unit) * kTile, half = slab_of(hh);
        const bool valid_x = MODE == MODE_DW || unit_tile(unit) < n_tiles;
        bool replay = false;
        if (rp && hh == 1) {
            mbar_wait(bar_unit, uidx & 1);
            replay = unit_clean(unit, uidx);
        }
        if (MODE == MODE_FG) {
            // ---- forward + expected-output-row mode (flash-attention style): besides the log-softmax statistics the
            // pair accumulates G = sum_v 2^(y_v - mref) * W16[v, slab] in TMEM against a per-row running reference
            // mref (log2 units).  The reference is fixed by the first tile and only moves when a later tile would
            // overflow the 16-bit operand; then the rows' accumulators are rescaled in TMEM (rare).
            // The blank and label columns are left out of G: their exact contribution (p - rb) W_blank + (p - rl) W_label
            // is added after the lattice by the reduction kernels.  EW = G * 2^(mref - lse2) / w_scale = sum_v p_v W_v.
            const int grow = x_row0 + row;
            const int label = valid_x ? p.row_label[grow] : -1;
            float mref = 0.f, ssum = 0.f, zb = 0.f, zl = 0.f;     // mref: finite start, fixed by the first tile
            // Range plan of the 16-bit operand P' = 2^(y - mref) (y = logit in log2 units + lg_scale).  fp16: the first
            // tile's row maximum sits at 2^2, later tiles may reach 2^15 before the reference has to move -- 13 binades
            // (9 nats) of headroom for logits above anything in the first 256 columns (a trained model's label logit
            // against a first tile that holds the blank), 16 normal + 10 subnormal binades below.  bf16 has the fp32
            // exponent range: the reference practically never moves.
            const float ref_exp = BF16 ? -2.f : 2.f;
            const float ref_limit = BF16 ? 100.f : 15.f;
            float* xg = kbuf;                                      // [2 column halves][128] row maxima (rare path)
            const int ngrp = p.HH / 32;
            for (int i = 0; i < (replay ? 0 : n_iter); ++i) {
                const int t0 = (j0 + i) * NT;
                const float* bias_t = p.bias2 + t0 + ch * 32;
                float4 bpre[8];                               // bias of sub-tile 0, fetched while waiting for the S tile
#pragma unroll
                for (int e = 0; e < 8; ++e) bpre[e] = __ldg(reinterpret_cast<const float4*>(bias_t) + e);
                mbar_wait(bar_sfull, (gs0 + i) & 1);
                tc_fence_after();
                uint32_t acc[4][32];
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld32(tmem_base + lane_addr + g * 64 + ch * 32, acc[g]);
                tmem_ld_wait();
                tc_fence_before();
                epi_arrive(bar_sempty);
                float lmax = -INFINITY;
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
                if (i == 0) {
                    // First tile: exact two-step (row maximum first, then the exponentials against the new reference).
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 bv = (g == 0) ? bpre[e] : __ldg(reinterpret_cast<const float4*>(bias_t + g * 64) + e);
                            const float y0 = fmaf(__uint_as_float(acc[g][4 * e + 0]), c1, bv.x + lg_scale);
                            const float y1 = fmaf(__uint_as_float(acc[g][4 * e + 1]), c1, bv.y + lg_scale);
                            const float y2 = fmaf(__uint_as_float(acc[g][4 * e + 2]), c1, bv.z + lg_scale);
                            const float y3 = fmaf(__uint_as_float(acc[g][4 * e + 3]), c1, bv.w + lg_scale);
                            lmax = fmaxf(lmax, fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)));
                            acc[g][4 * e + 0] = __float_as_uint(y0); acc[g][4 * e + 1] = __float_as_uint(y1);
                            acc[g][4 * e + 2] = __float_as_uint(y2); acc[g][4 * e + 3] = __float_as_uint(y3);
                        }
                    }
                    xg[ch * kTile + row] = lmax;
                    quarter_sync(q);
                    const float rmax = fmaxf(lmax, xg[(ch ^ 1) * kTile + row]);
                    // reference: the first tile's row maximum becomes 2^kRefExp in the 16-bit operand (-inf only on
                    // all-padding columns)
                    mref = (rmax > -INFINITY) ? (rmax - ref_exp) : 0.f;
                    quarter_sync(q);                              // xg may be rewritten
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            const float e0 = ex2f(__uint_as_float(acc[g][e]) - mref), e1 = ex2f(__uint_as_float(acc[g][e + 1]) - mref);
                            const float e2 = ex2f(__uint_as_float(acc[g][e + 2]) - mref), e3 = ex2f(__uint_as_float(acc[g][e + 3]) - mref);
                            p0 += e0; p1 += e1; p2 += e2; p3 += e3;
                            acc[g][e] = __float_as_uint(e0); acc[g][e + 1] = __float_as_uint(e1);
                            acc[g][e + 2] = __float_as_uint(e2); acc[g][e + 3] = __float_as_uint(e3);
                        }
                    }
                    ssum = (p0 + p1) + (p2 + p3);
                } else {
                    // Later tiles, optimistic single pass: exponentials against the CURRENT reference, fused with the
                    // logits so that FMA-pipe and MUFU work interleave; the partner warps then vote on the row maxima
                    // and only if a value came near the 16-bit range limit (rare: the reference is the first tile's
                    // maximum) the reference moves and everything computed so far is rescaled by a power of two.
                    const float krow = lg_scale - mref;
                    const uint64_t krow2 = pk2(krow, krow), c2 = pk2(c1, c1);
                    uint64_t s01 = pk2(0.f, 0.f), s23 = s01;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 bv = (g == 0) ? bpre[e] : __ldg(reinterpret_cast<const float4*>(bias_t + g * 64) + e);
                            // packed pairs: y = acc * c1 + (bias + krow)
                            const uint64_t y01 = fma2(pk2u(acc[g][4 * e + 0], acc[g][4 * e + 1]), c2, add2(pk2(bv.x, bv.y), krow2));
                            const uint64_t y23 = fma2(pk2u(acc[g][4 * e + 2], acc[g][4 * e + 3]), c2, add2(pk2(bv.z, bv.w), krow2));
                            float y0, y1, y2, y3;
                            unpk2(y01, y0, y1);
                            unpk2(y23, y2, y3);
                            lmax = fmaxf(lmax, fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)));
                            const float e0 = ex2f(y0), e1 = ex2f(y1), e2 = ex2f(y2), e3 = ex2f(y3);
                            s01 = add2(s01, pk2(e0, e1));
                            s23 = add2(s23, pk2(e2, e3));
                            acc[g][4 * e + 0] = __float_as_uint(e0); acc[g][4 * e + 1] = __float_as_uint(e1);
                            acc[g][4 * e + 2] = __float_as_uint(e2); acc[g][4 * e + 3] = __float_as_uint(e3);
                        }
                    }
                    {
                        const uint64_t st = add2(s01, s23);
                        unpk2(st, p0, p1);
                    }
                    float part = p0 + p1;
                    if (quarter_any(q, lmax > ref_limit)) {
                        if (rp && hh == 0) {                       // stored sub-tiles now carry mixed scales: no replay
                            flags[blockIdx.x >> 1] = uidx + 1;
                        }
                        xg[ch * kTile + row] = lmax;
                        quarter_sync(q);
                        const float rmax = fmaxf(lmax, xg[(ch ^ 1) * kTile + row]);
                        // new reference: the row maximum becomes 2^kRefExp again; values, sums and accumulators scale by 2^-delta
                        const float delta = (rmax > ref_limit) ? (rmax - ref_exp) : 0.f;
                        const float fsc = ex2f(-delta);
                        ssum *= fsc;
                        part *= fsc;
                        mref += delta;
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int e = 0; e < 32; ++e) acc[g][e] = __float_as_uint(__uint_as_float(acc[g][e]) * fsc);
                        // every G sub-pass issued so far (up to the previous tile's last one, which used the buffer
                        // before pr.buf) must have completed before the accumulators are scaled
                        mbar_wait(bar_pempty(pr.buf == 0 ? kPB - 1 : pr.buf - 1), pr.buf == 0 ? pr.phase ^ 1 : pr.phase);
                        tc_fence_after();
                        uint32_t gacc[16];
                        for (int cc = ch; cc < 2 * ngrp; cc += 2) {
                            tmem_ld16(tmem_G + lane_addr + cc * 16, gacc);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) gacc[e] = __float_as_uint(__uint_as_float(gacc[e]) * fsc);
                            tmem_st16(tmem_G + lane_addr + cc * 16, gacc);
                        }
                        tmem_st_wait();
                        tc_fence_before();
                        quarter_sync(q);                          // xg may be rewritten
                    }
                    ssum += part;
                }
                {
                    // blank / label logits (log2 units) of this row, recovered from the exponentials: once per row
                    const int cbl = p.blank - t0, clb = label - t0;   // column inside this tile, if any
                    if (cbl >= 0 && cbl < NT && ((cbl >> 5) & 1) == ch) {
                        float v = 0.f;
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int e = 0; e < 32; ++e) v = (g * 64 + ch * 32 + e == cbl) ? __uint_as_float(acc[g][e]) : v;
                        zb = lg2f(v) + mref - lg_scale;
                    }
                    if (clb >= 0 && clb < NT && ((clb >> 5) & 1) == ch) {
                        float v = 0.f;
#pragma unroll
                        for (int g = 0; g < 4; ++g)
#pragma unroll
                            for (int e = 0; e < 32; ++e) v = (g * 64 + ch * 32 + e == clb) ? __uint_as_float(acc[g][e]) : v;
                        zl = lg2f(v) + mref - lg_scale;
                    }
                }
                uint32_t packed[4][16];
#pragma unroll
                for (int g = 0; g < 4; ++g)
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        packed[g][e] = pack16<BF16>(__uint_as_float(acc[g][2 * e]), __uint_as_float(acc[g][2 * e + 1]));
                store_rounds(packed, [&](const int g0, const int g1, uint8_t* const* dst) {
                    const int cbl = p.blank - t0, clb = label - t0;
#pragma unroll
                    for (int g = g0; g < g1; ++g) {
                        const int lo = g * 64 + ch * 32;
                        if (cbl >= lo && cbl < lo + 32) *reinterpret_cast<uint16_t*>(dst[g] + ptile_off(row, cbl - g * 64)) = 0;
                        if (clb >= lo && clb < lo + 32) *reinterpret_cast<uint16_t*>(dst[g] + ptile_off(row, clb - g * 64)) = 0;
                    }
                }, i);
            }
            if (!replay) gs0 += n_iter;
            mbar_wait(bar_gfull, it & 1);
            tc_fence_after();
            if (!replay) {
                // combine the two column halves of each row through the (now idle) P' buffers
                float4* xch = reinterpret_cast<float4*>(sP_gen);
                pair_epi_sync();
                xch[ch * kTile + row] = make_float4(ssum, zb, zl, 0.f);
                pair_epi_sync();
                const float4 o = xch[(ch ^ 1) * kTile + row];
                pair_epi_sync();                          // (the next item's P' sub-tiles go into the same memory)
                const float lse2 = mref - lg_scale + lg2f(ssum + o.x);
                if (ch == 0 && valid_x && half == 0) {
                    zb = ((p.blank & 63) < 32) ? zb : o.y;        // which column half owns the blank / label column
                    if (label >= 0) zl = ((label & 63) < 32) ? zl : o.z;
                    p.lse[grow] = lse2 * kLn2;
                    p.lpb[grow] = (zb - lse2) * kLn2;
                    p.lpl[grow] = (label >= 0) ? (zl - lse2) * kLn2 : 0.f;
                }
                const float pf = ex2f(mref - lg_scale - lse2);    // softmax(row, v) = P'(row, v) * pf
                f_keep = pf * inv_ws;
            }
            const float f = f_keep;
            float* dst = p.dA + (size_t)grow * p.H + half * p.HH;
            uint32_t gacc[32];
            for (int cc = ch; cc < ngrp; cc += 2) {
                tmem_ld32(tmem_G + lane_addr + cc * 32, gacc);
                tmem_ld_wait();
                if (valid_x) {
#pragma unroll
                    for (int e = 0; e < 32; e += 4) {
                        float4 o4 = make_float4(__uint_as_float(gacc[e]) * f, __uint_as_float(gacc[e + 1]) * f,
                                                __uint_as_float(gacc[e + 2]) * f, __uint_as_float(gacc[e + 3]) * f);
                        *reinterpret_cast<float4*>(dst + cc * 32 + e) = o4;
                    }
                }
            }
        } else {
            float4 rm = make_float4(INFINITY, 0.f, 0.f, 0.f);
            int label = -1;
            float krow = 0.f, db_acc = 0.f;
            int vrow = 0;
            if (MODE == MODE_DA) {
                if (valid_x) {
                    rm = p.rowmeta[x_row0 + row];
                    label = p.row_label[x_row0 + row];
                }
                krow = fmaf(rm.x, -kLog2e, lg_scale);
            } else {
                vrow = x_row0 + row;
                krow = __ldg(p.bias2 + vrow);
            }
            for (int i = 0; i < (replay ? 0 : n_iter); ++i) {
                const int t0 = (j0 + i) * NT;               // first vocab id (DA) / lattice row (DW) of this stream tile
                float4 cm = make_float4(INFINITY, 0.f, 0.f, 0.f);
                int clabel = -1;
                if (MODE == MODE_DW) {
                    const int col = t0 + et;                // this thread owns column et of the 256-column tile
                    if (col < n_valid_rows) {
                        cm = __ldg(p.rowmeta + col);
                        clabel = __ldg(p.row_label + col);
                    }
                    kbuf[et] = fmaf(cm.x, -kLog2e, lg2f(fabsf(cm.w)) + lg_scale);
                    kbuf[NT + et] = (cm.w < 0.f) ? -1.f : 1.f;
                    pair_epi_sync();
                }
                mbar_wait(bar_sfull, (gs0 + i) & 1);
                tc_fence_after();
                // Pull this thread's share of the S tile (4 sub-tiles x 32 columns) into registers and hand the single S
                // accumulator back at once: the next tile's S pass then overlaps the exponentials below.
                uint32_t acc[4][32];
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld32(tmem_base + lane_addr + g * 64 + ch * 32, acc[g]);
                tmem_ld_wait();
                tc_fence_before();
                epi_arrive(bar_sempty);
                uint32_t packed[4][16];
                {
                    uint64_t d01 = pk2(0.f, 0.f), d23 = d01;
                    const uint64_t krow2 = pk2(krow, krow), c2 = pk2(c1, c1);
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int cb = g * 64 + ch * 32;    // this thread's first column of sub-tile g inside the tile
                        const float4* k4 = (MODE == MODE_DA) ? reinterpret_cast<const float4*>(p.bias2 + t0 + cb)
                                                             : reinterpret_cast<const float4*>(kbuf + cb);
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 kv = (MODE == MODE_DA) ? __ldg(k4 + e) : k4[e];
                            const uint64_t y01 = fma2(pk2u(acc[g][4 * e + 0], acc[g][4 * e + 1]), c2, add2(pk2(kv.x, kv.y), krow2));
                            const uint64_t y23 = fma2(pk2u(acc[g][4 * e + 2], acc[g][4 * e + 3]), c2, add2(pk2(kv.z, kv.w), krow2));
                            float y0, y1, y2, y3;
                            unpk2(y01, y0, y1);
                            unpk2(y23, y2, y3);
                            float v0 = ex2f(y0), v1 = ex2f(y1), v2 = ex2f(y2), v3 = ex2f(y3);
                            if (MODE == MODE_DW) {
                                if (any_neg) {
                                    const float4 sg = *reinterpret_cast<const float4*>(kbuf + NT + cb + 4 * e);
                                    v0 *= sg.x; v1 *= sg.y; v2 *= sg.z; v3 *= sg.w;
                                }
                                d01 = add2(d01, pk2(v0, v1));
                                d23 = add2(d23, pk2(v2, v3));
                            }
                            packed[g][2 * e] = pack16<BF16>(v0, v1);
                            packed[g][2 * e + 1] = pack16<BF16>(v2, v3);
                        }
                    }
                    if (MODE == MODE_DW) {
                        float d0, d1;
                        unpk2(add2(d01, d23), d0, d1);
                        db_acc += d0 + d1;
                    }
                }
                // sparse corrections: the blank and label entries are p - rb / p - rl, with p = exp(lp) from the
                // forward pass (rowmeta .y / .z), written exactly instead of being carried through the dense loop
                store_rounds(packed, [&](const int g0, const int g1, uint8_t* const* dst) {
                    if (MODE == MODE_DA) {
                        const int cbl = p.blank - t0, clb = label - t0;
#pragma unroll
                        for (int g = g0; g < g1; ++g) {
                            const int lo = g * 64 + ch * 32;
                            if (cbl >= lo && cbl < lo + 32)
                                *reinterpret_cast<uint16_t*>(dst[g] + ptile_off(row, cbl - g * 64)) = to16<BF16>(rm.y * pscale);
                            if (clb >= lo && clb < lo + 32)
                                *reinterpret_cast<uint16_t*>(dst[g] + ptile_off(row, clb - g * 64)) = to16<BF16>(rm.z * pscale);
                        }
                    } else {
                        pair_epi_sync();                    // column owners patch rows written by other threads
                        const int g = et >> 6;
                        if (g >= g0 && g < g1) {
                            uint8_t* d = (g == 0) ? dst[0] : (g == 1) ? dst[1] : (g == 2) ? dst[2] : dst[3];
                            const int rbl = p.blank - x_row0, rlb = clabel - x_row0;
                            if (rbl >= 0 && rbl < kTile)
                                *reinterpret_cast<uint16_t*>(d + ptile_off(rbl, et & 63)) = to16<BF16>(cm.y * cm.w * pscale);
                            if (rlb >= 0 && rlb < kTile)
                                *reinterpret_cast<uint16_t*>(d + ptile_off(rlb, et & 63)) = to16<BF16>(cm.z * cm.w * pscale);
                        }
                    }
                }, i);
            }
            if (!replay) gs0 += n_iter;
            // ---- final: G (128 x HH fp32 in TMEM) -> global
            mbar_wait(bar_gfull, it & 1);
            tc_fence_after();
            const float gmax = p.scal[2];
            const int ngrp = p.HH / 32;
            uint32_t gacc[32];
            // G columns [0, hh2) came from the leader's B rows, [hh2, HH) from the peer's: column c <-> joint column c
            if (MODE == MODE_DA) {
                const float f = rm.w * gmax * inv_ws / pscale;
                float* dst = p.dA + (size_t)(x_row0 + row) * p.H + half * p.HH;
                for (int cc = ch; cc < ngrp; cc += 2) {
                    tmem_ld32(tmem_G + lane_addr + cc * 32, gacc);
                    tmem_ld_wait();
                    if (valid_x) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4) {
                            float4 o = make_float4(__uint_as_float(gacc[e]) * f, __uint_as_float(gacc[e + 1]) * f,
                                                   __uint_as_float(gacc[e + 2]) * f, __uint_as_float(gacc[e + 3]) * f);
                            *reinterpret_cast<float4*>(dst + cc * 32 + e) = o;
                        }
                    }
                }
            } else {
                const float f = gmax / pscale;
                const bool ok = vrow < p.V;
                float* dst = p.dW + (size_t)vrow * p.H + half * p.HH;
                for (int cc = ch; cc < ngrp; cc += 2) {
                    tmem_ld32(tmem_G + lane_addr + cc * 32, gacc);
                    tmem_ld_wait();
                    if (ok) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            red_add_v4(dst + cc * 32 + e, __uint_as_float(gacc[e]) * f, __uint_as_float(gacc[e + 1]) * f,
                                       __uint_as_float(gacc[e + 2]) * f, __uint_as_float(gacc[e + 3]) * f);
                    }
                }
                // dense part of db: sum_m w_m * softmax(m, v); the sparse -rb / -rl terms are added by grad_prep_kernel
                if (ok && half == 0 && !replay) atomicAdd(p.db + vrow, db_acc * gmax / pscale);
            }
        }
        if (PERSIST) {                                // G has left TMEM: the next item may overwrite it
            tc_fence_before();
            epi_arrive(bar_gempty);
        }
        if (rp && hh == 0) {                          // first slab done in this warp (flag writes included)
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(bar_unit);
                mbar_arrive_cluster(bar_unit, rank ^ 1);
            }
        }
        }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kPairMmaWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// 2-D row-major [rows x H] 16-bit matrix, box = 64 columns x box_rows rows, 128-byte swizzle.
int make_tile_map(CUtensorMap* map, const void* base, uint64_t rows, int H, bool bf16, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)H, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)H * 2};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu H=%d)", (int)r,
                  (unsigned long long)rows, H);
        return 2;
    }
    return 0;
}

// 2-D row-major [rows x cols] 16-bit matrix (transposed operand copies), box = 64 columns x box_rows rows.
int make_matrix_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, bool bf16, int box_rows) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return 2;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)kKC, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu)", (int)r,
                  (unsigned long long)rows, (unsigned long long)cols);
        return 2;
    }
    return 0;
}

bool mma_supported_h(int H) {
    return H > 0 && H <= 512 && H % 64 == 0 && (H <= 256 || H == 384 || H == 512);
}

struct Plan {
    MmaParams p;
    int cg, stage_bytes;
    size_t smem;
};

// Shape-derived launch plan of joint_mma_kernel: forward = CTA pairs, generic backward = single CTAs.
static Plan plan(int H, int V, bool bwd) {
    Plan pl{};
    MmaParams& p = pl.p;
    p.H = H;
    p.NKC = H / 64;
    p.V = V;
    p.n_halves = (bwd && H > 256) ? 2 : 1;
    p.HH = H / p.n_halves;
    const int cg = bwd ? 1 : 2;          // (the pair kernels below cover the backward of H = 128, 256, 512)
    pl.cg = cg;
    pl.stage_bytes = (bwd && cg == 2) ? kChunkBytes / 2 : kChunkBytes;
    const int gst = kChunkBytes / pl.stage_bytes;
    p.NGCL = p.HH / 64 / (bwd ? cg : 1);
    const size_t fixed = (size_t)(p.NKC + (bwd ? 2 : 0)) * kChunkBytes + kNumBars * 8 + 16 + 4 * kTile * sizeof(float);
    const size_t limit = 232448;
    int ns = kMaxStages;
    while (ns > 2 && fixed + (size_t)ns * pl.stage_bytes > limit) --ns;
    p.GCH = 1;
    if (bwd) {
        if (cg == 2) {
            p.GCH = p.NGCL;
        } else if (p.NKC % 4 == 0 && p.NGCL % 4 == 0 && ns >= 4) {
            p.GCH = 4;
        } else if (p.NKC % 2 == 0 && p.NGCL % 2 == 0) {
            p.GCH = 2;
        }
        const int gstages = p.GCH * gst;
        ns = (ns / gstages) * gstages;
    }
    p.NS = ns;
    pl.smem = fixed + (size_t)ns * pl.stage_bytes;
    return pl;
}

template <int MODE, bool BF16, int CG>
static int launch(const CUtensorMap& mx, const CUtensorMap& my, const MmaParams& p, dim3 grid, size_t smem,
                  cudaStream_t stream) {
    auto kern = joint_mma_kernel<MODE, BF16, CG>;
    TTX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TTX_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mx, my, p));
    return 0;
}

int launch_joint_fwd(const void* a16, const void* w16, uint64_t rows_ub, int n_tiles_ub, int H, int V, int Vpad,
                     bool bf16, const int* meta, const float* bias2, const float* scal, const int* row_label,
                     int blank, float* lse, float* lpb, float* lpl, cudaStream_t stream) {
    Plan pl = plan(H, V, false);
    MmaParams& p = pl.p;
    p.blank = blank;
    p.splits = 1;
    p.meta = meta;
    p.bias2 = bias2;
    p.scal = scal;
    p.row_label = row_label;
    p.lse = lse;
    p.lpb = lpb;
    p.lpl = lpl;
    CUtensorMap mx, my;
    if (int rc = make_tile_map(&mx, a16, rows_ub, H, bf16, kTile)) return rc;
    if (int rc = make_tile_map(&my, w16, (uint64_t)Vpad, H, bf16, pl.stage_bytes / 128)) return rc;
    const dim3 grid((n_tiles_ub + 1) & ~1u, 1, 1);
    return bf16 ? launch<MODE_FWD, true, 2>(mx, my, p, grid, pl.smem, stream)
                : launch<MODE_FWD, false, 2>(mx, my, p, grid, pl.smem, stream);
}

template <int MODE, bool BF16>
static int launch_v3(const CUtensorMap& mx, const CUtensorMap& my, const CUtensorMap& myt, const MmaParams& p,
                     dim3 grid, size_t smem, cudaStream_t stream, const CUtensorMap* mscr = nullptr) {
    auto kern = joint_bwd_pair_kernel<MODE, BF16>;
    TTX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    cudaLaunchConfig_t cfg{};
    grid.x = (grid.x + 1) & ~1u;
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kPairThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TTX_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, mx, my, myt, mscr ? *mscr : mx, p));
    return 0;
}

// Split count of the persistent weight-gradient launch: units = vocabulary tile pairs x lattice-row splits should fill
// whole waves of the device's CTA pairs, with long units preferred (every item ends with a G read-out and a red.add
// of its tile: ~1.5 stream chunks' worth of time) up to `cap` stream chunks per unit.
static int dw_splits(int n_st, int n_vq, int pairs, int cap) {
    const int sp0 = (n_st + cap - 1) / cap;
    double best_score = -1.0;
    int best = sp0;
    for (int sp = sp0; sp <= min(n_st, sp0 + 96); ++sp) {
        const int len = (n_st + sp - 1) / sp;
        if ((n_st + len - 1) / len != sp) continue;         // trailing splits would be empty
        const int units = n_vq * sp;
        const int waves = (units + pairs - 1) / pairs;
        const double score = (double)units / ((double)pairs * waves) * len / (len + 1.5);
        if (score > best_score) {
            best_score = score;
            best = sp;
        }
    }
    return best;
}

// The pair kernels apply when both transposed copies exist and every CTA's share of a G slab is a whole number of
// 64-row chunk rows that fits one 16 KiB stage: HH / 2 in {64, 128}.
static bool v3_applicable(int H, const void* w16t, const void* a16t) {
    return w16t && a16t && (H == 128 || H == 256 || H == 512);
}

bool fwd_grad_supported_h(int H) { return H == 128 || H == 256 || H == 512; }

static int sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

// Persistent pair kernels: one CTA pair per SM pair, or fewer when there are fewer work items.
static unsigned persistent_grid(int n_tiles_ub, int n_halves) {
    const int items = ((n_tiles_ub + 1) / 2) * n_halves;
    return 2u * (unsigned)max(1, min(items, sm_count() / 2));
}

// P' replay scratch (pair kernels, H = 512): 64 KiB of per-pair flags + a 16-bit matrix of 128 rows per CTA, in the
// caller's workspace (ttx_joint_workspace_bytes).  Without a workspace the second slab of a unit recomputes its S pass.
static size_t fg_scratch_bytes(int n_tiles_ub, int V) {
    const unsigned grid = persistent_grid(n_tiles_ub, 2);
    return 65536 + (size_t)grid * kTile * ((V + 255) / 256) * 256 * 2;
}

struct DwShape {
    int splits, per;
    unsigned grid;
};
static DwShape dw_shape(int n_tiles_ub, int V) {
    // Persistent: units = vocabulary tile pairs x lattice-row splits of at most 160 stream chunks -- a unit's P' must fit
    // its CTA's share of the scratch matrix (10 MiB per CTA, 1.5 GB on a B200).
    const int n_vtiles = (V + kTile - 1) / kTile;
    const int n_st = (n_tiles_ub + 1) / 2, n_vq = (n_vtiles + 1) / 2, pairs = max(1, sm_count() / 2);
    DwShape d;
    d.splits = dw_splits(n_st, n_vq, pairs, 160);
    d.per = (n_st + d.splits - 1) / d.splits;
    d.grid = 2u * (unsigned)max(1, min(n_vq * d.splits, pairs));
    return d;
}
static size_t dw_scratch_bytes(int n_tiles_ub, int V) {
    const DwShape d = dw_shape(n_tiles_ub, V);
    return 65536 + (size_t)d.grid * kTile * d.per * 256 * 2;
}

size_t joint_workspace_bytes(int which, int n_tiles_ub, int H, int V) {
    if (H != 512) return 0;
    return which == 0 ? fg_scratch_bytes(n_tiles_ub, V) : dw_scratch_bytes(n_tiles_ub, V);
}

// Forward statistics + EW = sum_v p_v W_v (blank / label columns excluded) in one pass (MODE_FG of the pair kernel).
int launch_joint_fwd_grad(const void* a16, const void* w16, const void* w16t, uint64_t rows_ub, int n_tiles_ub, int H,
                          int V, int Vpad, bool bf16, const int* meta, const float* bias2, const float* scal,
                          const int* row_label, int blank, float* lse, float* lpb, float* lpl, float* ew,
                          void* ws, size_t ws_bytes, cudaStream_t stream) {
    MmaParams p{};
    p.H = H;
    p.NKC = H / 64;
    p.V = V;
    p.n_halves = (H > 256) ? 2 : 1;
    p.HH = H / p.n_halves;
    const size_t fixed = (size_t)(p.NKC + kPB) * kChunkBytes + kNumBars * 8 + 16 + 4 * kTile * sizeof(float);
    int ns = 8;
    while (ns > 2 && fixed + (size_t)ns * kChunkBytes > 232448) --ns;
    p.NS = ns;
    const size_t smem = fixed + (size_t)ns * kChunkBytes;
    p.blank = blank;
    p.splits = 1;
    p.meta = meta;
    p.bias2 = bias2;
    p.scal = scal;
    p.row_label = row_label;
    p.lse = lse;
    p.lpb = lpb;
    p.lpl = lpl;
    p.dA = ew;
    CUtensorMap mx, my, myt;
    if (int rc = make_tile_map(&mx, a16, rows_ub, H, bf16, kTile)) return rc;
    if (int rc = make_tile_map(&my, w16, (uint64_t)Vpad, H, bf16, kTile)) return rc;
    if (int rc = make_matrix_map(&myt, w16t, (uint64_t)H, (uint64_t)Vpad, bf16, p.HH / 2)) return rc;
    dim3 grid(persistent_grid(n_tiles_ub, p.n_halves), 1, 1);
    // P' replay (H = 512): the scratch matrix holds the P' of the unit a CTA pair is working on: 128 rows per CTA x Vpad
    // columns (165 MB on a B200, whatever the problem size).
    CUtensorMap mscr;
    const int n_chunks = (V + 255) / 256;
    bool have_map = false;
    if (p.n_halves == 2 && ws != nullptr) {
        if (ws_bytes < fg_scratch_bytes(n_tiles_ub, V)) {
            set_error("ttx_joint_fwd_grad: workspace of %zu bytes, ttx_joint_workspace_bytes asks for %zu", ws_bytes,
                      fg_scratch_bytes(n_tiles_ub, V));
            return 1;
        }
        TTX_CUDA_OK(cudaMemsetAsync(ws, 0, 65536, stream));
        if (int rc = make_matrix_map(&mscr, static_cast<uint8_t*>(ws) + 65536,
                                     (uint64_t)grid.x * kTile * (uint64_t)(n_chunks * 4), kKC, bf16, kTile))
            return rc;
        p.scr_rows = (int)(grid.x * kTile);
        p.scratch = static_cast<uint8_t*>(ws);
        have_map = true;
    }
    int rc = bf16 ? launch_v3<MODE_FG, true>(mx, my, myt, p, grid, smem, stream, have_map ? &mscr : nullptr)
                  : launch_v3<MODE_FG, false>(mx, my, myt, p, grid, smem, stream, have_map ? &mscr : nullptr);
    return rc;
}

int launch_joint_bwd(const void* a16, const void* w16, const void* a16t, const void* w16t, uint64_t rows_ub,
                     int n_tiles_ub, int H, int V, int Vpad, bool bf16, const int* meta, const float* bias2,
                     const float* scal, const int* row_label, int blank, const float4* rowmeta, float* dA, float* dW,
                     float* db, int splits, void* ws, size_t ws_bytes, cudaStream_t stream) {
    if (v3_applicable(H, w16t, a16t)) {
        MmaParams p{};
        p.H = H;
        p.NKC = H / 64;
        p.V = V;
        p.n_halves = (H > 256) ? 2 : 1;
        p.HH = H / p.n_halves;
                const size_t fixed = (size_t)(p.NKC + kPB) * kChunkBytes + kNumBars * 8 + 16 + 4 * kTile * sizeof(float);
        int ns = 8;
            while (ns > 2 && fixed + (size_t)ns * kChunkBytes > 232448) --ns;
        p.NS = ns;
        const size_t smem = fixed + (size_t)ns * kChunkBytes;
        p.blank = blank;
        p.meta = meta;
        p.bias2 = bias2;
        p.scal = scal;
        p.row_label = row_label;
        p.rowmeta = rowmeta;
        p.dA = dA;
        p.dW = dW;
        p.db = db;
        if (dA) {
            CUtensorMap mx, my, myt;
            if (int rc = make_tile_map(&mx, a16, rows_ub, H, bf16, kTile)) return rc;
            if (int rc = make_tile_map(&my, w16, (uint64_t)Vpad, H, bf16, kTile)) return rc;
            if (int rc = make_matrix_map(&myt, w16t, (uint64_t)H, (uint64_t)Vpad, bf16, p.HH / 2)) return rc;
            p.splits = 1;
            dim3 grid(persistent_grid(n_tiles_ub, p.n_halves), 1, 1);
            int rc = bf16 ? launch_v3<MODE_DA, true>(mx, my, myt, p, grid, smem, stream)
                          : launch_v3<MODE_DA, false>(mx, my, myt, p, grid, smem, stream);
            if (rc) return rc;
        }
        if (dW) {
            CUtensorMap mx, my, myt;
            if (int rc = make_tile_map(&mx, w16, (uint64_t)Vpad, H, bf16, kTile)) return rc;
                    if (int rc = make_tile_map(&my, a16, rows_ub, H, bf16, kTile)) return rc;
            if (int rc = make_matrix_map(&myt, a16t, (uint64_t)H * (rows_ub / kKC), kKC, bf16, p.HH / 2)) return rc;
            const DwShape d = dw_shape(n_tiles_ub, V);
            p.splits = d.splits;
            (void)splits;
            dim3 grid(d.grid, 1, 1);
            CUtensorMap mscr;
            bool have_map = false;
            if (p.n_halves == 2 && ws != nullptr) {
                if (ws_bytes < dw_scratch_bytes(n_tiles_ub, V)) {
                    set_error("ttx_joint_grad: workspace of %zu bytes, ttx_joint_workspace_bytes asks for %zu", ws_bytes,
                              dw_scratch_bytes(n_tiles_ub, V));
                    return 1;
                }
                TTX_CUDA_OK(cudaMemsetAsync(ws, 0, 65536, stream));
                if (int rc = make_matrix_map(&mscr, static_cast<uint8_t*>(ws) + 65536,
                                             (uint64_t)grid.x * kTile * (uint64_t)(d.per * 4), kKC, bf16, kTile))
                    return rc;
                p.scr_rows = (int)(grid.x * kTile);
                p.scratch = static_cast<uint8_t*>(ws);
                have_map = true;
            }
            int rc = bf16 ? launch_v3<MODE_DW, true>(mx, my, myt, p, grid, smem, stream, have_map ? &mscr : nullptr)
                          : launch_v3<MODE_DW, false>(mx, my, myt, p, grid, smem, stream, have_map ? &mscr : nullptr);
            if (rc) return rc;
        }
        return 0;
    }
    Plan pl = plan(H, V, true);
    MmaParams& p = pl.p;
    p.blank = blank;
    p.meta = meta;
    p.bias2 = bias2;
    p.scal = scal;
    p.row_label = row_label;
    p.rowmeta = rowmeta;
    p.dA = dA;
    p.dW = dW;
    p.db = db;
    const int box = pl.stage_bytes / 128;
    const int n_vtiles = (V + kTile - 1) / kTile;
    if (dA) {
        CUtensorMap mx, my;
        if (int rc = make_tile_map(&mx, a16, rows_ub, H, bf16, kTile)) return rc;
        if (int rc = make_tile_map(&my, w16, (uint64_t)Vpad, H, bf16, box)) return rc;
        p.splits = 1;
        const dim3 grid(n_tiles_ub, p.n_halves, 1);
        if (int rc = bf16 ? launch<MODE_DA, true, 1>(mx, my, p, grid, pl.smem, stream)
                          : launch<MODE_DA, false, 1>(mx, my, p, grid, pl.smem, stream))
            return rc;
    }
    if (dW) {
        CUtensorMap mx, my;
        if (int rc = make_tile_map(&mx, w16, (uint64_t)Vpad, H, bf16, kTile)) return rc;
        if (int rc = make_tile_map(&my, a16, rows_ub, H, bf16, box)) return rc;
        p.splits = splits;
        const dim3 grid(n_vtiles, p.n_halves, splits);
        if (int rc = bf16 ? launch<MODE_DW, true, 1>(mx, my, p, grid, pl.smem, stream)
                          : launch<MODE_DW, false, 1>(mx, my, p, grid, pl.smem, stream))
            return rc;
    }
    return 0;
}

}  // namespace ttx
