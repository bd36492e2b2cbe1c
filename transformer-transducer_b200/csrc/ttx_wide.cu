// Joint widths beyond one shared-memory tile (aishell.yaml's inner_size 1024, joint_streaming.yaml's 2048;
// /root/reference/tt/model.py:35-37 with config/aishell.yaml:42-45, config/joint_streaming.yaml:42-45): the same three
// contractions as ttx_joint_mma.cu, on tcgen05 tensor cores, organised as three plain streamed products around the
// 16-bit softmax numerators P' = 2^(y - mref) (y = logit in log2 units):
//
//   sp_kernel      S = A16 . W16^T, BOTH operands streamed in 64-column K chunks (no stationary tile: 128 rows x H no longer
//                  fit shared memory), two 256-column S accumulators in TMEM so that the exponential epilogue of chunk j
//                  overlaps the MMAs of chunk j + 1.  Epilogue: online log-sum-exp statistics (lse, log p(blank), log p(label))
//                  + P' staged in shared memory and moved by TMA stores to the blocked P' matrix, [Vpad / 64][rows][64].
//   kp_kernel<PW>  EW = P' . W16 for a 512-column block of H (two 256-column slabs = all of TMEM), K = vocabulary.
//   kp_kernel<DW>  dW_out += P'^T . As for a 512-column block of H, K = lattice rows; As = row-scaled A16 (ttx_small.cu).
//
// P' rows are stored against a per-row reference mref fixed by the first vocabulary chunk; a row whose later logits
// would overflow the 16-bit range moves its reference (rare, see the range plan in ttx_joint_mma.cu) and flags its tile
// pair, and a second launch of sp_kernel (redo mode, same grid, skips every unflagged pair on the device) recomputes
// the flagged pairs against the FINAL reference -- so the matrix the two products read is always consistent and there is
// no whole-batch fallback.
#include "ttx_common.cuh"

namespace ttx {

int make_tile_map(CUtensorMap* map, const void* base, uint64_t rows, int H, bool bf16, int box_rows);
int make_matrix_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, bool bf16, int box_rows);

struct WideParams {
    int H, NKC;              // joint width, H / 64
    int V, n_vchunks;        // vocabulary size, 256-column chunks of it
    int n_k64;               // ceil(V / 64): K groups of the EW product
    int blank;
    int NS;                  // sp: ring stages (32 KiB each)
    int tile_lo, tile_cnt;   // lattice tiles [tile_lo, tile_lo + tile_cnt) (clipped to the tiles in use); tile_lo even
    int store_rows;          // rows of the P' matrix; its row 0 is lattice row tile_lo * 128
    int redo;                // sp: only tile pairs whose flag is set, against the stored final reference
    int n_hb;                // kp: 512-column blocks of H
    int splits;              // kp<DW>: lattice-row splits
    const int* meta;
    const float* bias2;
    const float* scal;
    const int* row_label;
    float* lse;
    float* lpb;
    float* lpl;
    float* pfac;             // softmax(row, v) = P'(row, v) * pfac[row]
    float* mref;             // the row's final reference (log2 units)
    uint16_t* pstore;
    int* flags;              // one word per tile pair of the whole batch
    float* ew;               // kp<PW> out (rows x H)
    float* dW;               // kp<DW> out (V x H), red.add
    float* db;               // kp<DW> out (V)
};

constexpr int kWEpiWarps = 8;
constexpr int kWEpiThreads = kWEpiWarps * 32;
constexpr int kWThreads = kWEpiThreads + 128;          // + control warpgroup: producer, MMA issuer, watcher, idle
constexpr int kWProducerWarp = kWEpiWarps, kWMmaWarp = kWEpiWarps + 1, kWWatchWarp = kWEpiWarps + 2;
// 128 registers per thread: three quarters of the register file, so that a block of a bandwidth-bound kernel launched on
// another stream (the activation-gradient reduction, the lattice) fits next to it -- these products leave the SM's issue
// slots and load/store path idle.
constexpr int kWRegs = 128;
constexpr int kSpStage = 2 * kChunkBytes;              // A chunk + W chunk
// The S pass runs SIXTEEN epilogue warps (four per TMEM lane quarter): its epilogue -- one exponential per logit -- is
// latency-bound with two warps per scheduler (ncu: the schedulers issue 30 % of the cycles, profiles/r2_sp_epilogue.md).
constexpr int kSpEpiWarps = 16;
constexpr int kSpEpiThreads = kSpEpiWarps * 32;
constexpr int kSpThreads = kSpEpiThreads + 128;
constexpr int kSpProducerWarp = kSpEpiWarps, kSpMmaWarp = kSpEpiWarps + 1, kSpWatchWarp = kSpEpiWarps + 2;
// (640 threads: the compiler's budget is 96 registers per thread for every role, so there is nothing for setmaxnreg to move)

__device__ __forceinline__ void w_epi_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kWEpiThreads) : "memory");
}
__device__ __forceinline__ void sp_epi_sync() {
    asm volatile("bar.sync 1, %0;" ::"n"(kSpEpiThreads) : "memory");
}
__device__ __forceinline__ void sp_quarter_sync(int q) {
    asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(kSpEpiThreads / 4) : "memory");
}
__device__ __forceinline__ bool sp_quarter_any(int q, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred pin, pout;\n\t"
        "setp.ne.b32 pin, %2, 0;\n\t"
        "bar.red.or.pred pout, %1, %3, pin;\n\t"
        "selp.u32 %0, 1, 0, pout;\n\t}"
        : "=r"(r)
        : "r"(2 + q), "r"((uint32_t)pred), "n"(kSpEpiThreads / 4)
        : "memory");
    return r != 0;
}
__device__ __forceinline__ void w_quarter_sync(int q) {
    asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "n"(kWEpiThreads / 4) : "memory");
}
__device__ __forceinline__ bool w_quarter_any(int q, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred pin, pout;\n\t"
        "setp.ne.b32 pin, %2, 0;\n\t"
        "bar.red.or.pred pout, %1, %3, pin;\n\t"
        "selp.u32 %0, 1, 0, pout;\n\t}"
        : "=r"(r)
        : "r"(2 + q), "r"((uint32_t)pred), "n"(kWEpiThreads / 4)
        : "memory");
    return r != 0;
}

// byte offset of 16-bit element (row, col < 64) inside a staged [128 x 64] sub-tile (128-byte rows, 128B swizzle)
__device__ __forceinline__ uint32_t stile_off(int row, int col) {
    return row * 128 + (((col >> 3) ^ (row & 7)) << 4) + (col & 7) * 2;
}

// The issuing thread's view of the barrier watcher's event counter (see ttx_joint_mma.cu: an mbarrier try_wait costs
// ~90 cycles, a tcgen05.mma must be issued every ~128).
struct EventWait {
    volatile int* ready;
    int need = 0, have = 0;
    __device__ __forceinline__ void wait() {
        ++need;
        if (have < need) {
            uint32_t spins = 0;
            while ((have = *ready) < need) {
                if (++spins > (1u << 26)) {
                    printf("ttx: MMA issuer timed out waiting for event %d (block %d)\n", need, blockIdx.x);
                    __trap();
                }
            }
        }
    }
};

// =========================================================================================================== S pass
// With both operands streamed the S pass moves 256 KiB per 128 x 256 tile through TMA in 6100 - 7300 cycles, against 4096
// cycles of MMA time (ncu: tensor pipe 66 %).  Timing decomposition at configs[1] (profiles/r2_summary.md): 1.55 ms with the
// epilogue switched off -- the operand stream at the SM's TMA receive limit (~40 B/cycle/SM; full-rate MMAs want 64) --
// + ~0.25 ms for the P' store (same port) + ~0.25 ms for the epilogue's math (issue slots / shared memory next to the
// single-thread roles) = 2.05 - 2.2 ms.  Variants that cut the TMA bytes, all measured and removed:
//   A16 tiles stationary in shared memory (H <= 512: 128 KiB per CTA, only W16 streams): the tile takes the room of either
//     the ring or the P' staging buffers -- staging buffers + three 16 KiB ring stages 2.23 ms (latency-bound ring); five
//     stages + P' stored straight from registers 3.32 ms with 16-byte stores (half-sector writes), 2.87 ms with 32-byte
//     stores; HALF of the tile stationary (192 KiB per tile, seven 16 KiB stages) 2.42 ms against 2.21 ms;
//   A16 tiles stationary in TENSOR memory (TS-mode tcgen05.mma, tile copied ring -> tcgen05.st, eleven 16 KiB W16 stages):
//     256 columns of A leave ONE accumulator, and its hand-over chain (commit -> 16 warps wake -> tcgen05.ld -> 32 remote
//     arrivals -> issuer, ~2300 cycles per tile) costs what the halved stream saves: 2.25 - 2.31 ms against 2.05 - 2.20 ms,
//     parity-green (lane = row, a 32-bit column = two consecutive K elements, low half first).
// At full chip the board's power cap is the next limit anyway: on 148 SMs the clock sits at ~1515 of 1965 MHz, on 74 SMs
// the kernel loses only 1.5x (profiles/r2_summary.md).
template <bool BF16>
__global__ void __launch_bounds__(kSpThreads, 1)
sp_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapY,
          const __grid_constant__ CUtensorMap mapP, const WideParams p) {
    constexpr int STG = kSpStage;                          // bytes of one ring stage (A chunk + W chunk)
    constexpr int NSB = 2;                                 // staging buffers: one 64-column P' sub-tile each, used in turn
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int n_tiles = p.meta[0];
    const int n_range = min(n_tiles, p.tile_lo + p.tile_cnt) - p.tile_lo;
    const int n_units = (n_range + 1) >> 1;
    const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
    if (unit0 >= n_units) return;
    auto skip_unit = [&](int unit) { return p.redo && p.flags[(p.tile_lo >> 1) + unit] == 0; };

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) {
        if (threadIdx.x == 0) printf("ttx: dynamic shared memory is not 1024-byte aligned (0x%x)\n", smem_base);
        __trap();
    }
    const uint32_t sRing = smem_base;
    const uint32_t sStage = sRing + p.NS * STG;            // NSB [128 x 64] 16-bit P' sub-tiles (128B swizzle) for the TMA store
    const uint32_t sBar = sStage + NSB * kChunkBytes;
    const uint32_t sTmemPtr = sBar + 40 * 8;
    const uint32_t sWatch = sTmemPtr + 8;
    const uint32_t sXg = sTmemPtr + 16;                    // [4][128] floats: row maxima of the four column groups
    const uint32_t sXch = sXg + 4 * kTile * 4;             // [4][128] float4: per-row statistics of the four column groups
    const uint32_t sBias = sXch + 4 * kTile * 16;          // [4 lane quarters][2][256] floats: bias2 of this / the next chunk
    uint8_t* smem_gen = smem_raw;
    auto bar_full = [&](int s) { return sBar + 8 * s; };
    auto bar_empty = [&](int s) { return sBar + 8 * (8 + s); };
    auto bar_sfull = [&](int b) { return sBar + 8 * (16 + b); };
    auto bar_sempty = [&](int b) { return sBar + 8 * (18 + b); };
    auto bar_pwritten = [&](int b) { return sBar + 8 * (20 + b); };   // this CTA's epilogue warps have staged a sub-tile in buffer b
    auto bar_pfree = [&](int b) { return sBar + 8 * (22 + b); };      // ... and the TMA store has read it out of shared memory
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kSpProducerWarp && lane == 0) {
        tma_prefetch_desc(&mapX);
        tma_prefetch_desc(&mapY);
        tma_prefetch_desc(&mapP);
        for (int b = 0; b < NSB; ++b) {
            mbar_init(bar_pwritten(b), kSpEpiWarps / 2);     // the eight warps that stage into buffer b
            mbar_init(bar_pfree(b), 1);
        }
        for (int s = 0; s < p.NS; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_empty(s), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_sfull(b), 1);
            mbar_init(bar_sempty(b), 2 * kSpEpiWarps);     // every epilogue warp of both CTAs
        }
        *reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base)) = 0;
        fence_barrier_init();
    }
    if (warp == kSpMmaWarp) tmem_alloc_pair(sTmemPtr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (sTmemPtr - smem_base));

    if (warp >= kSpEpiWarps) {
        if (warp == kSpProducerWarp && lane == 0) {
            // =================================================== TMA producer (each CTA: its rows of A16, its half of W16)
            Ring r;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (skip_unit(unit)) continue;
                const int x_row0 = (p.tile_lo + unit * 2 + (int)rank) * kTile;
                for (int j = 0; j < p.n_vchunks; ++j)
                    for (int c = 0; c < p.NKC; ++c) {
                        mbar_wait(bar_empty(r.stage), r.phase ^ 1);
                        if (leader) mbar_arrive_expect_tx(bar_full(r.stage), 2 * STG);
                        const uint32_t dst = sRing + r.stage * STG;
                        tma_load_2d_pair(dst, &mapX, bar_full(r.stage), c * kKC, x_row0);
                        tma_load_2d_pair(dst + kChunkBytes, &mapY, bar_full(r.stage), c * kKC, j * 256 + (int)rank * kTile);
                        r.advance(p.NS);
                    }
            }
        } else if (warp == kSpWatchWarp && lane == 0 && leader) {
            // =================================================== barrier watcher: the issuer's barriers, in its order
            volatile int* ready = reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base));
            int done = 0, g = 0;
            Ring r;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (skip_unit(unit)) continue;
                for (int j = 0; j < p.n_vchunks; ++j, ++g) {
                    mbar_wait(bar_sempty(g & 1), ((g >> 1) & 1) ^ 1);
                    *ready = ++done;
                    for (int c = 0; c < p.NKC; ++c) {
                        mbar_wait(bar_full(r.stage), r.phase);
                        *ready = ++done;
                        r.advance(p.NS);
                    }
                }
            }
        } else if (warp == kSpWatchWarp + 1 && lane == 0) {
            // =================================================== storer (each CTA): staged P' tile -> the blocked matrix
            int n = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (skip_unit(unit)) continue;
                const int srow0 = (unit * 2 + (int)rank) * kTile;          // row of the P' matrix (relative to tile_lo)
                for (int j = 0; j < p.n_vchunks; ++j)
                    for (int gg = 0; gg < 4; ++gg, ++n) {           // sub-tile n goes through buffer n % NSB
                        const int b = n % NSB;
                            mbar_wait(bar_pwritten(b), (n / NSB) & 1);
                        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                                     ::"l"(reinterpret_cast<uint64_t>(&mapP)), "r"(sStage + b * kChunkBytes), "r"(0),
                                       "r"((j * 4 + gg) * p.store_rows + srow0)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                        // release the buffer as soon as THIS store has read it (a few hundred cycles; this thread has
                        // nothing else to do).  Releasing a buffer only when the next sub-tile's store had been issued made
                        // the warps of buffer 1 wait for the staging of buffer 0 of the same round (13 % of the kernel's
                        // stall samples); the global writes stay in flight either way.
                        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                        mbar_arrive(bar_pfree(b));
                    }
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        } else if (warp == kSpMmaWarp && lane == 0 && leader) {
            // =================================================== MMA issuer
            const uint32_t idescS = make_idesc(BF16 ? 1 : 0, 0, 0, 256, 256);
            const uint32_t rlo = desc_lo(sRing);
            EventWait ev{reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base))};
            int stage = 0, g = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (skip_unit(unit)) continue;
                for (int j = 0; j < p.n_vchunks; ++j, ++g) {
                    ev.wait();                              // the epilogue has read this accumulator's previous tile
                    tc_fence_after();
                    const uint32_t d = tmem_base + (g & 1) * 256;
                    for (int c = 0; c < p.NKC; ++c) {
                        ev.wait();                          // ring stage has landed
                        tc_fence_after();
                        const uint32_t a = rlo + stage * (STG >> 4), b = a + (kChunkBytes >> 4);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_f16_ss_pair_lo(d, a + 2 * k, b + 2 * k, idescS, (c | k) != 0);
                        umma_commit_pair(bar_empty(stage));
                        if (++stage == p.NS) stage = 0;
                    }
                    umma_commit_pair(bar_sfull(g & 1));
                }
            }
        }
    } else {
        // ======================================================= epilogue: thread = (row, column group cq)
        // Sixteen warps, four per TMEM lane quarter.  A 256-column S tile is done in two ROUNDS of two 64-column
        // sub-tiles; in a round thread (row, cq) owns 32 columns: sub-tile rd * 2 + (cq >> 1), half cq & 1 -- read-out,
        // exponentials, pack, and the sub-tile is staged (128-byte swizzled rows) for the storer thread's TMA store,
        // buffer = cq >> 1.  Later chunks are OPTIMISTIC: values go out against the current reference and the row's four
        // threads vote afterwards; if one left the 16-bit range the reference moves and the tile pair is flagged -- the
        // redo launch rewrites all of its P' against the final reference.
        const int q = warp & 3, cq = warp >> 2;
        const int sub = cq >> 1, hf = cq & 1;
        const int row = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const float inv_ws = p.scal[1];
        const float c1 = inv_ws * kLog2e;
        const float lg_scale = BF16 ? 0.0f : 12.0f;
        const float ref_exp = BF16 ? -2.f : 2.f;           // range plan of P': see ttx_joint_mma.cu (MODE_FG)
        const float ref_limit = BF16 ? 100.f : 15.f;
        float* xg = reinterpret_cast<float*>(smem_gen + (sXg - smem_base));
        float4* xch = reinterpret_cast<float4*>(smem_gen + (sXch - smem_base));
        uint8_t* dstP = smem_gen + (sStage - smem_base) + sub * kChunkBytes;
        uint8_t* r0 = dstP + row * 128;
        auto epi_arrive = [&](uint32_t bar) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar, 0);
        };
        auto row_max4 = [&](float v) {                       // maximum over the row's four threads (-> all four)
            xg[cq * kTile + row] = v;
            sp_quarter_sync(q);
            const float m = fmaxf(fmaxf(xg[row], xg[kTile + row]), fmaxf(xg[2 * kTile + row], xg[3 * kTile + row]));
            sp_quarter_sync(q);                              // xg may be rewritten
            return m;
        };
        // bias2 of a chunk comes from shared memory: with 227 KB of it configured there is no L1 left, and a global load
        // per use costs an L2 round trip each time (ncu: the largest stall of the epilogue).  Each lane quarter keeps its
        // own double-buffered copy, filled one chunk ahead by its 128 threads (two floats each) and published by the
        // quarter barrier that ends every tile.
        float* qbias = reinterpret_cast<float*>(smem_gen + (sBias - smem_base)) + q * 512;
        const int bt = cq * 32 + lane;                     // this thread's two floats of a chunk: bt, bt + 128
        qbias[bt] = __ldg(p.bias2 + bt);
        qbias[bt + 128] = __ldg(p.bias2 + bt + 128);
        sp_quarter_sync(q);
        int g = 0;                                         // S tiles so far (accumulator + staging + bias parities)
        for (int unit = unit0; unit < n_units; unit += unit_step) {
            if (skip_unit(unit)) continue;
            const int tile = p.tile_lo + unit * 2 + (int)rank;
            const bool valid_x = tile < n_tiles;
            const int grow = tile * kTile + row;
            const int label = valid_x ? p.row_label[grow] : -1;
            // (redo: rows of a pad tile have no stored reference -- an unreachable one makes their P' exact zeros)
            float mref = p.redo ? (valid_x ? p.mref[grow] : 1.0e4f) : 0.f;
            float ssum = 0.f, zb = 0.f, zl = 0.f;
            for (int j = 0; j < p.n_vchunks; ++j, ++g) {
                const int t0 = j * 256;
                const uint32_t tcol = tmem_base + lane_addr + (g & 1) * 256 + sub * 64 + hf * 32;
                const float* bias_t = qbias + (g & 1) * 256 + sub * 64 + hf * 32;
                const bool first = (j == 0 && !p.redo);
                // next tile's chunk (the next unit starts at chunk 0 again): fetched now, stored at the end of this tile
                const int jn = (j + 1 == p.n_vchunks) ? 0 : j + 1;
                const float nb0 = __ldg(p.bias2 + jn * 256 + bt), nb1 = __ldg(p.bias2 + jn * 256 + bt + 128);
                mbar_wait(bar_sfull(g & 1), (g >> 1) & 1);
                tc_fence_after();
                uint32_t acc[32];
                if (first) {
                    // First chunk: exact two-step -- the row maximum fixes the reference, then the rounds below re-read
                    // the accumulator and take their exponentials against it.
                    float lm = -INFINITY;
#pragma unroll
                    for (int rd = 0; rd < 2; ++rd) {
                        tmem_ld32(tcol + rd * 128, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float4 bv = *(reinterpret_cast<const float4*>(bias_t + rd * 128) + e);
                            lm = fmaxf(lm, fmaxf(fmaxf(fmaf(__uint_as_float(acc[4 * e + 0]), c1, bv.x), fmaf(__uint_as_float(acc[4 * e + 1]), c1, bv.y)),
                                                 fmaxf(fmaf(__uint_as_float(acc[4 * e + 2]), c1, bv.z), fmaf(__uint_as_float(acc[4 * e + 3]), c1, bv.w))));
                        }
                    }
                    const float rmax = row_max4(lm);
                    mref = (rmax > -INFINITY) ? (rmax + lg_scale - ref_exp) : 0.f;
                }
                const float krow = lg_scale - mref;
                const uint64_t krow2 = pk2(krow, krow), c2 = pk2(c1, c1);
                uint64_t s01 = pk2(0.f, 0.f), s23 = s01;
                float lmax = -INFINITY;
#pragma unroll
                for (int rd = 0; rd < 2; ++rd) {
                    tmem_ld32(tcol + rd * 128, acc);
                    tmem_ld_wait();
                    if (rd == 1) {                            // the accumulator is free again as soon as it sits in registers
                        tc_fence_before();
                        epi_arrive(bar_sempty(g & 1));
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        float y0, y1, y2, y3;
                        const float4 bv = *(reinterpret_cast<const float4*>(bias_t + rd * 128) + e);
                        unpk2(fma2(pk2u(acc[4 * e + 0], acc[4 * e + 1]), c2, add2(pk2(bv.x, bv.y), krow2)), y0, y1);
                        unpk2(fma2(pk2u(acc[4 * e + 2], acc[4 * e + 3]), c2, add2(pk2(bv.z, bv.w), krow2)), y2, y3);
                        lmax = fmaxf(lmax, fmaxf(fmaxf(y0, y1), fmaxf(y2, y3)));
                        const float e0 = ex2f(y0), e1 = ex2f(y1), e2 = ex2f(y2), e3 = ex2f(y3);
                        s01 = add2(s01, pk2(e0, e1));
                        s23 = add2(s23, pk2(e2, e3));
                        acc[4 * e + 0] = __float_as_uint(y0); acc[4 * e + 1] = __float_as_uint(y1);
                        acc[4 * e + 2] = __float_as_uint(y2); acc[4 * e + 3] = __float_as_uint(y3);
                        pk[2 * e] = pack16<BF16>(e0, e1);
                        pk[2 * e + 1] = pack16<BF16>(e2, e3);
                    }
                    // blank / label logits of this row (log2 units): acc holds the exponents y = logit + lg_scale - mref
                    const int c0 = t0 + rd * 128 + sub * 64 + hf * 32;      // first vocabulary id of this round's 32 columns
                    const int cbl = p.blank - c0, clb = label - c0;
                    if (cbl >= 0 && cbl < 32) {
                        float v = 0.f;
#pragma unroll
                        for (int e = 0; e < 32; ++e) v = (e == cbl) ? __uint_as_float(acc[e]) : v;
                        zb = v - krow;
                    }
                    if (clb >= 0 && clb < 32) {
                        float v = 0.f;
#pragma unroll
                        for (int e = 0; e < 32; ++e) v = (e == clb) ? __uint_as_float(acc[e]) : v;
                        zl = v - krow;
                    }
                    // stage this round's sub-tile (the store of the previous round's has read the buffer); the blank and
                    // label columns are left out of P' (their exact terms are added in fp32 later)
                    mbar_wait(bar_pfree(sub), (rd & 1) ^ 1);
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(r0 + (((hf * 4 + c) ^ (row & 7)) << 4)) =
                            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                    if (cbl >= 0 && cbl < 32) *reinterpret_cast<uint16_t*>(dstP + stile_off(row, hf * 32 + cbl)) = 0;
                    if (clb >= 0 && clb < 32) *reinterpret_cast<uint16_t*>(dstP + stile_off(row, hf * 32 + clb)) = 0;
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_pwritten(sub));
                }
                float part;
                {
                    float p0, p1;
                    unpk2(add2(s01, s23), p0, p1);
                    part = p0 + p1;
                }
                qbias[((g + 1) & 1) * 256 + bt] = nb0;
                qbias[((g + 1) & 1) * 256 + bt + 128] = nb1;
                if (first) sp_quarter_sync(q);              // (publishes the next chunk's bias; later chunks: the vote below)
                if (!first && sp_quarter_any(q, lmax > ref_limit)) {
                    if (!p.redo) p.flags[(p.tile_lo >> 1) + unit] = 1;         // this pair's P' now carries mixed scales
                    const float rmax = row_max4(lmax);
                    const float delta = (rmax > ref_limit) ? (rmax - ref_exp) : 0.f;
                    const float fsc = ex2f(-delta);
                    ssum *= fsc;
                    part *= fsc;
                    mref += delta;
                }
                ssum += part;
            }
            // combine the row's four column groups
            sp_epi_sync();
            xch[cq * kTile + row] = make_float4(ssum, zb, zl, 0.f);
            sp_epi_sync();
            if (cq == 0 && valid_x) {
                const float4 o1 = xch[kTile + row], o2 = xch[2 * kTile + row], o3 = xch[3 * kTile + row];
                const float lse2 = mref - lg_scale + lg2f((ssum + o1.x) + (o2.x + o3.x));
                // which group owns the blank / label column: sub-tile parity (bit 6) and half (bit 5) of the column
                const int ob = ((p.blank >> 6) & 1) * 2 + ((p.blank >> 5) & 1);
                zb = ob == 0 ? zb : ob == 1 ? o1.y : ob == 2 ? o2.y : o3.y;
                if (label >= 0) {
                    const int ol = ((label >> 6) & 1) * 2 + ((label >> 5) & 1);
                    zl = ol == 0 ? zl : ol == 1 ? o1.z : ol == 2 ? o2.z : o3.z;
                }
                p.lse[grow] = lse2 * kLn2;
                p.lpb[grow] = (zb - lse2) * kLn2;
                p.lpl[grow] = (label >= 0) ? (zl - lse2) * kLn2 : 0.f;
                p.pfac[grow] = ex2f(mref - lg_scale - lse2);
                p.mref[grow] = mref;
            }
            sp_epi_sync();                                   // (the next unit rewrites xch)
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kSpMmaWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// =========================================================================================================== kept products
// One product with all 512 TMEM columns as accumulator (two 256-column slabs of the H block), no S pass, no exponentials.
// Per 64 elements of K the ring takes a GROUP of three 16 KiB stages: the A operand and the two slabs' B operands; eight
// MMAs (four k slices x two slabs) per group.
//   PW: rows = lattice rows (pair: two 128-row tiles), K = vocabulary.  A = P' box [128 rows x 64 v] (K-major),
//       B = W16^T boxes [128 h x 64 v].  Out: EW = G * pfac / w_scale.
//   DW: rows = vocabulary (pair: two 128-row vocabulary tiles), K = lattice rows.  A = P'^T, read MN-major straight from
//       the row-major P' blocks (two [64 m x 64 v] boxes), B = As^T, read MN-major from the row-major scaled copy of A16
//       (two [64 m x 64 h] boxes per slab) -- no transposed copy of either exists; + the 64 row scales of the stage
//       (H block 0 only), from which the idle epilogue warps form the dense part of db out of the P' stage in shared
//       memory.  Out: red.add into dW.
enum { KP_PW = 0, KP_DW = 1 };
constexpr int kKG = 3;                                  // stages per group = groups in flight

template <int MODE, bool BF16>
__global__ void __maxnreg__(kWRegs)
kp_kernel(const __grid_constant__ CUtensorMap mapP, const __grid_constant__ CUtensorMap mapB,
          const __grid_constant__ CUtensorMap mapS, const WideParams p) {
    constexpr int STAGE = kChunkBytes;
    constexpr int NRING = kKG * kKG;
    const uint32_t rank = cluster_ctarank();
    const bool leader = (rank == 0);
    const int n_tiles = p.meta[0];
    const int n_range = min(n_tiles, p.tile_lo + p.tile_cnt) - p.tile_lo;
    const int n_tp = (n_range + 1) >> 1;                 // lattice tile pairs in range
    const int n_vq = ((p.V + kTile - 1) / kTile + 1) / 2;
    const int per = (MODE == KP_DW) ? (n_tp + p.splits - 1) / p.splits : 0;
    const int n_units = (MODE == KP_DW) ? n_vq * p.n_hb * p.splits : n_tp * p.n_hb;
    const int unit0 = blockIdx.x >> 1, unit_step = gridDim.x >> 1;
    if (unit0 >= n_units || n_range <= 0) return;
    // unit -> (row tile pair of the output, H block, K range in 64-element groups)
    int hb = 0, rt = 0, k_lo = 0, k_n = 0;
    auto begin_unit = [&](int unit) {
        hb = unit % p.n_hb;
        const int rest = unit / p.n_hb;
        if (MODE == KP_DW) {
            rt = rest % n_vq;
            const int j0 = (rest / n_vq) * per;
            k_lo = j0 * 4;                               // 64-row groups: four per tile pair
            k_n = (min(n_tp, j0 + per) - j0) * 4;
        } else {
            rt = rest;
            k_lo = 0;
            k_n = p.n_k64;
        }
        return k_n > 0;
    };

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = smem_u32(smem_raw);
    if (smem_base & 1023u) {
        if (threadIdx.x == 0) printf("ttx: dynamic shared memory is not 1024-byte aligned (0x%x)\n", smem_base);
        __trap();
    }
    const uint32_t sX = smem_base;                                   // NRING stages
    const uint32_t sScale = sX + NRING * STAGE;                      // DW: kKG x [8 x 64] 16-bit scale rows
    const uint32_t sBar = sScale + kKG * 1024;
    const uint32_t sTmemPtr = sBar + 32 * 8;
    const uint32_t sWatch = sTmemPtr + 8;
    const uint32_t sPart = sTmemPtr + 16;                            // DW: [16][128] floats, partial sums of db
    uint8_t* smem_gen = smem_raw;
    auto bar_full = [&](int s) { return sBar + 8 * s; };
    auto bar_empty = [&](int s) { return sBar + 8 * (NRING + s); };
    auto bar_cons = [&](int g) { return sBar + 8 * (2 * NRING + g); };
    const uint32_t bar_gfull = sBar + 8 * (2 * NRING + kKG);
    const uint32_t bar_gempty = bar_gfull + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == kWProducerWarp && lane == 0) {
        tma_prefetch_desc(&mapP);
        tma_prefetch_desc(&mapB);
        tma_prefetch_desc(&mapS);
        for (int s = 0; s < NRING; ++s) {
            mbar_init(bar_full(s), 1);
            // DW: the P' stage and the first As^T stage of a group are released by this CTA's epilogue warps (which read
            // them after the MMAs, see bar_cons); everything else by the MMAs' commit
            const bool epi_released = MODE == KP_DW && s % kKG != kKG - 1;
            mbar_init(bar_empty(s), epi_released ? kWEpiWarps : 1);
        }
        for (int g = 0; g < kKG; ++g) mbar_init(bar_cons(g), 1);
        mbar_init(bar_gfull, 1);
        mbar_init(bar_gempty, 2 * kWEpiWarps);
        *reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base)) = 0;
        fence_barrier_init();
    }
    if (warp == kWMmaWarp) tmem_alloc_pair(sTmemPtr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem_gen + (sTmemPtr - smem_base));
    const int row_off = p.tile_lo * kTile;                           // lattice row of the P' matrix's row 0

    if (warp >= kWEpiWarps) {
        if (warp == kWProducerWarp && lane == 0) {
            // =================================================== TMA producer
            Ring r;
            auto stage_in = [&](uint32_t& full, uint32_t& dst, uint32_t bytes) {
                mbar_wait(bar_empty(r.stage), r.phase ^ 1);
                full = bar_full(r.stage);
                dst = sX + r.stage * STAGE;
                if (leader) mbar_arrive_expect_tx(full, 2 * bytes);
            };
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (!begin_unit(unit)) continue;
                for (int i = 0; i < k_n; ++i) {
                    uint32_t full, dst;
                    if (MODE == KP_PW) {
                        const int srow0 = (p.tile_lo + rt * 2 + (int)rank) * kTile - row_off;
                        stage_in(full, dst, STAGE);
                        tma_load_2d_pair(dst, &mapP, full, 0, (k_lo + i) * p.store_rows + srow0);
                        r.advance(NRING);
                        for (int s = 0; s < 2; ++s) {
                            stage_in(full, dst, STAGE);
                            tma_load_2d_pair(dst, &mapB, full, (k_lo + i) * kKC, hb * 512 + s * 256 + (int)rank * kTile);
                            r.advance(NRING);
                        }
                    } else {
                        const int v0 = (rt * 2 + (int)rank) * kTile;            // this CTA's vocabulary rows = P' columns
                        const int m0 = (k_lo + i) * kKC;                        // first lattice row (of the range) of the group
                        stage_in(full, dst, STAGE);
                        tma_load_2d_pair(dst, &mapP, full, 0, (v0 / kKC) * p.store_rows + m0);
                        tma_load_2d_pair(dst + STAGE / 2, &mapP, full, 0, (v0 / kKC + 1) * p.store_rows + m0);
                        const int grp = r.stage / kKG;
                        r.advance(NRING);
                        // As: row-major [rows x H] like A16; the CTA's 128 joint columns of a slab = two [64 m x 64 h] boxes
                        // (the B operand is read MN-major, exactly like P'^T on the A side)
                        for (int s = 0; s < 2; ++s) {
                            const bool with_scale = (s == 0 && hb == 0);
                            stage_in(full, dst, STAGE + (with_scale ? 128 : 0));
                            const int hcol = hb * 512 + s * 256 + (int)rank * kTile;
                            tma_load_2d_pair(dst, &mapB, full, hcol, row_off + m0);
                            tma_load_2d_pair(dst + STAGE / 2, &mapB, full, hcol + kKC, row_off + m0);
                            if (with_scale) tma_load_2d_pair(sScale + grp * 1024, &mapS, full, 0, (row_off + m0) / kKC);
                            r.advance(NRING);
                        }
                    }
                }
            }
        } else if (warp == kWWatchWarp && lane == 0 && leader) {
            // =================================================== barrier watcher
            volatile int* ready = reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base));
            int done = 0, it = 0;
            Ring r;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (!begin_unit(unit)) continue;
                for (int i = 0; i < k_n; ++i) {
                    if (i == 0 && it > 0) mbar_wait(bar_gempty, (it - 1) & 1);      // previous unit's G has left TMEM
                    for (int s = 0; s < kKG; ++s) {
                        mbar_wait(bar_full(r.stage), r.phase);
                        r.advance(NRING);
                    }
                    *ready = ++done;
                }
                ++it;
            }
        } else if (warp == kWMmaWarp && lane == 0 && leader) {
            // =================================================== MMA issuer
            constexpr int fmt = BF16 ? 1 : 0;
            const uint32_t idesc = make_idesc(fmt, MODE == KP_DW ? 1 : 0, MODE == KP_DW ? 1 : 0, 256, 256);
            const uint32_t xlo = desc_lo(sX);
            // MN-major A (DW): 64-wide blocks 8 KiB apart (LBO), a 16-row k slice = +2 KiB
            const uint32_t amn = (xlo & 0xFFFFu) | (512u << 16);
            EventWait ev{reinterpret_cast<volatile int*>(smem_gen + (sWatch - smem_base))};
            int rstage = 0;
            for (int unit = unit0; unit < n_units; unit += unit_step) {
                if (!begin_unit(unit)) continue;
                for (int i = 0; i < k_n; ++i) {
                    ev.wait();
                    tc_fence_after();
                    const uint32_t b0 = xlo + (rstage + 1) * 1024, b1 = b0 + 1024;
                    if (MODE == KP_DW) {
                        const uint32_t a = amn + rstage * 1024, bm0 = a + 1024, bm1 = a + 2048;     // all three MN-major
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16_ss_pair_lo(tmem_base, a + 128 * k, bm0 + 128 * k, idesc, (i | k) != 0);
                            umma_f16_ss_pair_lo(tmem_base + 256, a + 128 * k, bm1 + 128 * k, idesc, (i | k) != 0);
                        }
                        umma_commit_pair(bar_cons(rstage / kKG));       // the epilogue warps release stages r and r + 1
                        umma_commit_pair(bar_empty(rstage + 2));
                    } else {
                        const uint32_t a = xlo + rstage * 1024;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            umma_f16_ss_pair_lo(tmem_base, a + 2 * k, b0 + 2 * k, idesc, (i | k) != 0);
                            umma_f16_ss_pair_lo(tmem_base + 256, a + 2 * k, b1 + 2 * k, idesc, (i | k) != 0);
                        }
                        umma_commit_pair(bar_empty(rstage));
                        umma_commit_pair(bar_empty(rstage + 1));
                        umma_commit_pair(bar_empty(rstage + 2));
                    }
                    rstage = (rstage + kKG == NRING) ? 0 : rstage + kKG;
                }
                umma_commit_pair(bar_gfull);
            }
        }
    } else {
        // ======================================================= epilogue warps
        const int q = warp & 3, ch = warp >> 2;
        const int row = q * 32 + lane;
        const int et = threadIdx.x;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        auto epi_arrive = [&](uint32_t bar) {
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(bar, 0);
        };
        Ring kr;
        int it = 0;
        for (int unit = unit0; unit < n_units; unit += unit_step) {
            if (!begin_unit(unit)) continue;
            uint32_t gacc[32];
            if (MODE == KP_DW) {
                // dense part of db[v] = sum_m s_m P'[m, v] from the P' stage and the row scales (row 0 of the 8-row scale
                // tile) once the MMAs have read the group; then the warp releases both stages.  (H block 0 only.)  Thread =
                // (box, 16-byte chunk of 8 v, group of 4 lattice rows): four 16-byte loads per stage -- scalar 2-byte loads
                // cost the shared-memory port a whole wavefront each, and the port is what bounds this kernel (TMA writes,
                // MMA reads and these reads together).  The 16 row groups' partial sums meet in shared memory per unit.
                const int vc = et & 7, bx = (et >> 3) & 1, mg = et >> 4;
                const int x_row0 = (rt * 2 + (int)rank) * kTile;
                const uint8_t* sX_gen = smem_gen + (sX - smem_base);
                const uint8_t* sS_gen = smem_gen + (sScale - smem_base);
                float dacc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) dacc[e] = 0.f;
                for (int i = 0; i < k_n; ++i) {
                    const int g = kr.stage / kKG;
                    mbar_wait(bar_cons(g), kr.phase);
                    if (hb == 0) {
                        const uint8_t* pst = sX_gen + kr.stage * STAGE + bx * (STAGE / 2);
                        const uint2 s4 = *reinterpret_cast<const uint2*>(sS_gen + g * 1024 + mg * 8);   // scales of 4 rows
                        float sf[4];
                        unpk16<BF16>(s4.x, sf[0], sf[1]);
                        unpk16<BF16>(s4.y, sf[2], sf[3]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int m = mg * 4 + j;
                            const uint4 pv = *reinterpret_cast<const uint4*>(pst + m * 128 + ((vc ^ (m & 7)) << 4));
                            const uint32_t w4[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                float p0, p1;
                                unpk16<BF16>(w4[e], p0, p1);
                                dacc[2 * e] = fmaf(p0, sf[j], dacc[2 * e]);
                                dacc[2 * e + 1] = fmaf(p1, sf[j], dacc[2 * e + 1]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(bar_empty(kr.stage));
                        mbar_arrive(bar_empty(kr.stage + 1));
                    }
                    kr.advance(NRING, kKG);
                }
                const float f = p.scal[2] * (BF16 ? 1.0f : 1.0f / kKeptUp);
                if (hb == 0) {
                    float* part = reinterpret_cast<float*>(smem_gen + (sPart - smem_base));      // [16 row groups][128 v]
#pragma unroll
                    for (int e = 0; e < 8; ++e) part[mg * kTile + bx * 64 + vc * 8 + e] = dacc[e];
                    w_epi_sync();
                    if (et < kTile) {
                        float sum = 0.f;
#pragma unroll
                        for (int gq = 0; gq < 16; ++gq) sum += part[gq * kTile + et];
                        if (x_row0 + et < p.V) atomicAdd(p.db + x_row0 + et, sum * f);
                    }
                    w_epi_sync();
                }
                mbar_wait(bar_gfull, it & 1);
                tc_fence_after();
                const int vrow = x_row0 + row;
                const bool ok = vrow < p.V;
                for (int cc = ch; cc < 16; cc += 2) {               // TMEM column cc * 32 <-> joint column hb * 512 + cc * 32
                    tmem_ld32(tmem_base + lane_addr + cc * 32, gacc);
                    tmem_ld_wait();
                    if (ok) {
                        float* dst = p.dW + (size_t)vrow * p.H + hb * 512 + cc * 32;
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            red_add_v4(dst + e, __uint_as_float(gacc[e]) * f, __uint_as_float(gacc[e + 1]) * f,
                                       __uint_as_float(gacc[e + 2]) * f, __uint_as_float(gacc[e + 3]) * f);
                    }
                }
            } else {
                const int tile = p.tile_lo + rt * 2 + (int)rank;
                const bool valid_x = tile < n_tiles;
                const int grow = tile * kTile + row;
                const float f = valid_x ? p.pfac[grow] * p.scal[1] : 0.f;
                mbar_wait(bar_gfull, it & 1);
                tc_fence_after();
                float* dst = p.ew + (size_t)grow * p.H + hb * 512;
                for (int cc = ch; cc < 16; cc += 2) {
                    tmem_ld32(tmem_base + lane_addr + cc * 32, gacc);
                    tmem_ld_wait();
                    if (valid_x) {
#pragma unroll
                        for (int e = 0; e < 32; e += 4)
                            *reinterpret_cast<float4*>(dst + cc * 32 + e) =
                                make_float4(__uint_as_float(gacc[e]) * f, __uint_as_float(gacc[e + 1]) * f,
                                            __uint_as_float(gacc[e + 2]) * f, __uint_as_float(gacc[e + 3]) * f);
                    }
                }
            }
            tc_fence_before();
            epi_arrive(bar_gempty);
            ++it;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kWMmaWarp) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------- host side
bool wide_supported_h(int H) { return H >= 512 && H <= 4096 && H % 512 == 0; }

static int wide_sm_count() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    return sms;
}

template <typename K, typename... Args>
static int launch_pair(K kern, int threads, unsigned grid_x, size_t smem, cudaStream_t stream, Args... args) {
    TTX_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((grid_x + 1) & ~1u, 1, 1);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TTX_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, args...));
    return 0;
}

static WideParams wide_params(int H, int V, int Vpad, int tile_lo, int tile_cnt, uint64_t store_rows, const int* meta,
                              const float* scal) {
    WideParams p{};
    p.H = H;
    p.NKC = H / kKC;
    p.V = V;
    p.n_vchunks = Vpad / 256;
    p.n_k64 = (V + kKC - 1) / kKC;          // (64-column groups that hold vocabulary: the padding beyond has P' = 0)
    p.tile_lo = tile_lo;
    p.tile_cnt = tile_cnt;
    p.store_rows = (int)store_rows;
    p.n_hb = H / 512;
    p.meta = meta;
    p.scal = scal;
    return p;
}

// S pass over lattice tiles [tile_lo, tile_lo + tile_cnt): statistics + P' (pass over every tile pair, then the redo
// pass over the flagged ones).  store_rows >= 128 * (tile_cnt rounded up to even).
int launch_wide_sp(const void* a16, const void* w16, uint64_t rows_ub, int tile_lo, int tile_cnt, int H, int V, int Vpad,
                   bool bf16, const int* meta, const float* bias2, const float* scal, const int* row_label, int blank,
                   float* lse, float* lpb, float* lpl, float* pfac, float* mref, void* pstore, uint64_t store_rows,
                   int* flags, cudaStream_t stream) {
    WideParams p = wide_params(H, V, Vpad, tile_lo, tile_cnt, store_rows, meta, scal);
    p.blank = blank;
    p.bias2 = bias2;
    p.row_label = row_label;
    p.lse = lse;
    p.lpb = lpb;
    p.lpl = lpl;
    p.pfac = pfac;
    p.mref = mref;
    p.pstore = static_cast<uint16_t*>(pstore);
    p.flags = flags;
    const size_t fixed = 2 * kChunkBytes + 40 * 8 + 16 + 4 * kTile * 4 + 4 * kTile * 16 + 4 * 512 * 4;
    p.NS = 8;
    while (p.NS > 2 && (size_t)p.NS * kSpStage + fixed > 232448) --p.NS;
    const size_t smem = (size_t)p.NS * kSpStage + fixed;
    CUtensorMap mx, my, mp;
    if (int rc = make_tile_map(&mx, a16, rows_ub, H, bf16, kTile)) return rc;
    if (int rc = make_tile_map(&my, w16, (uint64_t)Vpad, H, bf16, kTile)) return rc;
    if (int rc = make_matrix_map(&mp, pstore, store_rows * (uint64_t)(Vpad / kKC), kKC, bf16, kTile)) return rc;
    const unsigned grid = 2u * (unsigned)max(1, min((tile_cnt + 1) / 2, wide_sm_count() / 2));
    for (int redo = 0; redo < 2; ++redo) {
        p.redo = redo;
        int rc = bf16 ? launch_pair(sp_kernel<true>, kSpThreads, grid, smem, stream, mx, my, mp, p)
                      : launch_pair(sp_kernel<false>, kSpThreads, grid, smem, stream, mx, my, mp, p);
        if (rc) return rc;
    }
    return 0;
}

// EW rows of tiles [tile_lo, tile_lo + tile_cnt) = P' . W16 * pfac / w_scale
int launch_wide_pw(const void* pstore, uint64_t store_rows, const void* w16t, int tile_lo, int tile_cnt, int H, int V,
                   int Vpad, bool bf16, const int* meta, const float* scal, const float* pfac, float* ew,
                   cudaStream_t stream) {
    WideParams p = wide_params(H, V, Vpad, tile_lo, tile_cnt, store_rows, meta, scal);
    p.pfac = const_cast<float*>(pfac);
    p.ew = ew;
    const size_t smem = (size_t)kKG * kKG * kChunkBytes + kKG * 1024 + 32 * 8 + 16;
    CUtensorMap mp, mb;
    if (int rc = make_matrix_map(&mp, pstore, store_rows * (uint64_t)(Vpad / kKC), kKC, bf16, kTile)) return rc;
    if (int rc = make_matrix_map(&mb, w16t, (uint64_t)H, (uint64_t)Vpad, bf16, kTile)) return rc;
    const int units = ((tile_cnt + 1) / 2) * p.n_hb;
    const unsigned grid = 2u * (unsigned)max(1, min(units, wide_sm_count() / 2));
    return bf16 ? launch_pair(kp_kernel<KP_PW, true>, kWThreads, grid, smem, stream, mp, mb, mb, p)
                : launch_pair(kp_kernel<KP_PW, false>, kWThreads, grid, smem, stream, mp, mb, mb, p);
}

// dW += gmax / kKeptUp * P'^T . As over the lattice rows of tiles [tile_lo, tile_lo + tile_cnt), db likewise (dense part)
int launch_wide_dw(const void* pstore, uint64_t store_rows, const void* a16st, uint64_t rows_ub, int tile_lo, int tile_cnt,
                   int H, int V, int Vpad, bool bf16, const int* meta, const float* scal, float* dW, float* db,
                   cudaStream_t stream) {
    WideParams p = wide_params(H, V, Vpad, tile_lo, tile_cnt, store_rows, meta, scal);
    p.dW = dW;
    p.db = db;
    const size_t smem = (size_t)kKG * kKG * kChunkBytes + kKG * 1024 + 32 * 8 + 16 + 16 * kTile * 4;
    const int n_tp = (tile_cnt + 1) / 2, n_vq = ((V + kTile - 1) / kTile + 1) / 2, pairs = max(1, wide_sm_count() / 2);
    // lattice-row splits: units = vocabulary tile pairs x H blocks x splits should fill whole waves of the CTA pairs,
    // long units preferred (every unit ends with a read-out + red.add of its 256 x 512 tile: ~1.5 tile pairs' worth)
    int best = 1;
    double best_score = -1.0;
    for (int sp = 1; sp <= min(n_tp, 128); ++sp) {
        const int len = (n_tp + sp - 1) / sp;
        if ((n_tp + len - 1) / len != sp) continue;
        const int units = n_vq * p.n_hb * sp;
        const int waves = (units + pairs - 1) / pairs;
        const double score = (double)units / ((double)pairs * waves) * len / (len + 1.5);
        if (score > best_score) {
            best_score = score;
            best = sp;
        }
    }
    p.splits = best;
    CUtensorMap mp, mb, ms;
    if (int rc = make_matrix_map(&mp, pstore, store_rows * (uint64_t)(Vpad / kKC), kKC, bf16, 64)) return rc;
    if (int rc = make_tile_map(&mb, a16st, rows_ub, H, bf16, kKC)) return rc;                          // As, boxes [64 m x 64 h]
    const uint16_t* svec = static_cast<const uint16_t*>(a16st) + rows_ub * (uint64_t)H;                  // [rows / 64][64]
    if (int rc = make_matrix_map(&ms, svec, rows_ub / kKC, kKC, bf16, 1)) return rc;
    const unsigned grid = 2u * (unsigned)max(1, min(n_vq * p.n_hb * p.splits, pairs));
    return bf16 ? launch_pair(kp_kernel<KP_DW, true>, kWThreads, grid, smem, stream, mp, mb, ms, p)
                : launch_pair(kp_kernel<KP_DW, false>, kWThreads, grid, smem, stream, mp, mb, ms, p);
}

}  // namespace ttx
