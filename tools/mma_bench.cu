// Microbenchmark (design evidence, not product): tcgen05.mma issue rate by shape / CTA-group / operand source, and
// tcgen05.ld read-out rate, on all 148 SMs.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_bench.bin tools/mma_bench.cu -lcuda
#include "../transformer-transducer_b200/csrc/ttx_common.cuh"
#include <stdarg.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

namespace ttx {
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
    fputc('\n', stderr);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts_pair(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
}  // namespace ttx
using namespace ttx;

// mode: 0 = SS, 1 = TS (A from TMEM columns 256..), 2 = SS + 8 warps reading TMEM concurrently, 3 = TMEM read only,
// 4 = SS with an MN-major A operand, 5 = SS with an MN-major B operand, 6 = both MN-major, 7 = SS with A bf16 / B f16
template <int CG>
__global__ void __launch_bounds__(320, 1) bench(int M, int N, int mode, int n_mma, int ld_warps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t sA = base;                    // 4 chunks of 16 KiB
    const uint32_t sB = base + 4 * 16384;        // 4 stages of 32 KiB
    const uint32_t sBar = sB + 4 * 32768;
    const uint32_t sTp = sBar + 64;
    volatile uint32_t* tp = reinterpret_cast<volatile uint32_t*>(smem_raw + (sTp - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;
    // deterministic finite operands
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;   // fp16 1.0 pairs
    if (threadIdx.x == 0) {
        mbar_init(sBar, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) tmem_alloc_pair(sTp, 512);
        else tmem_alloc(sTp, 512);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tp;
    long long t0 = 0, t1 = 0;
    if (warp == 1) {
        if (lane == 0 && rank == 0 && mode != 3) {
            uint32_t idesc = make_idesc(0, mode == 4 || mode == 6, mode == 5 || mode == 6, M, N);
            if (mode == 7) idesc |= 1u << 7;               // A format bf16, B stays f16
            // descriptors = base + (chunk, k slice) steps of the start-address field (16-byte units), fixed per mode so
            // that the issue loop stays within one MMA's time
            const bool amn = (mode == 4 || mode == 6), bmn = (mode == 5 || mode == 6);
            const uint64_t da0 = amn ? desc_mnmajor(sA, 0, 8192) : desc_kmajor(sA, 0);
            const uint64_t db0 = bmn ? desc_mnmajor(sB, 0, 8192) : desc_kmajor(sB, 0);
            const uint32_t aks = amn ? 128 : 2, bks = bmn ? 128 : 2;
            t0 = clock64();
            for (int i = 0; i < n_mma; ++i) {
                const int c = (i >> 2) & 3, k = i & 3;
                const uint64_t db = db0 + (uint64_t)(c * 2048 + k * bks);
                if (mode == 1) {
                    if (CG == 2) umma_f16_ts_pair(tmem, tmem + 256 + (i & 7) * 8, db, idesc, i != 0);
                    else umma_f16_ts(tmem, tmem + 256 + (i & 7) * 8, db, idesc, i != 0);
                } else {
                    const uint64_t da = da0 + (uint64_t)(c * 1024 + k * aks);
                    if (CG == 2) umma_f16_ss_pair(tmem, da, db, idesc, i != 0);
                    else umma_f16_ss(tmem, da, db, idesc, i != 0);
                }
            }
            if (CG == 2) umma_commit_pair(sBar);
            else umma_commit(sBar);
            mbar_wait(sBar, 0);
            t1 = clock64();
            out[blockIdx.x] = t1 - t0;
        }
    } else if (warp >= 2 && (mode == 2 || mode == 3) && warp - 2 < ld_warps) {
        const int q = warp & 3;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        uint32_t acc[32];
        uint32_t sink = 0;
        const int reps = (mode == 3) ? n_mma : n_mma / 4;
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            tmem_ld32(tmem + lane_addr + 256 + ((i * 32) & 255), acc);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) sink ^= acc[e];
        }
        t1 = clock64();
        if (sink == 0x12345678u) printf("x");
        if (mode == 3 && lane == 0 && warp == 2) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        if (CG == 2) tmem_dealloc_pair(tmem, 512);
        else tmem_dealloc(tmem, 512);
    }
}

template <int CG>
static void run(const char* name, int M, int N, int mode, int n_mma, int ld_warps, long long* d_out) {
    auto kern = bench<CG>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = 4 * 16384 + 4 * 32768 + 128;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    std::vector<long long> h(148);
    for (int rep = 0; rep < 3; ++rep) {
        cudaMemset(d_out, 0, 148 * sizeof(long long));
        cudaEventRecord(e0);
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, M, N, mode, n_mma, ld_warps, d_out);
        cudaEventRecord(e1);
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e != cudaSuccess || e2 != cudaSuccess) {
            printf("%-34s FAILED: %s / %s\n", name, cudaGetErrorString(e), cudaGetErrorString(e2));
            exit(1);
        }
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    cudaMemcpy(h.data(), d_out, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::vector<long long> v;
    for (auto x : h) if (x > 0) v.push_back(x);
    std::sort(v.begin(), v.end());
    const double med = v.empty() ? 0 : (double)v[v.size() / 2];
    if (mode == 3) {
        printf("%-34s ld_warps=%d  cycles/ld(32x32 cols, per warp)=%.1f  -> %.1f B/cycle/SM   (%.3f ms)\n", name, ld_warps,
               med / n_mma, 4096.0 * ld_warps * n_mma / med, best);
    } else {
        const double flops = 2.0 * M * N * 16 * (double)n_mma * (148 / CG);
        const double ideal = (double)std::max(M / CG, 128) * N / 256.0 / 1.0;   // cycles per instruction per SM at 4096 MAC/clk
        printf("%-34s M=%3d N=%3d cycles/mma=%.1f (ideal %.0f)  %.0f TFLOP/s  (%.3f ms)\n", name, M, N, med / n_mma,
               (double)(M / CG) * N * 16 / 4096.0, flops / (best * 1e-3) / 1e12, best);
        (void)ideal;
    }
}


// ---- DSMEM throughput: CTA 0 of a pair sends `reps` x 64 KiB to CTA 1's shared memory.
// mode 0: st.shared::cluster.v4 from 256 threads, fence.acq_rel.cluster once per 64 KiB
// mode 1: st.async.v4 with complete_tx on a barrier in the destination CTA
// mode 2: cp.async.bulk.shared::cluster from local shared memory, 16 KiB per copy, complete_tx at the destination
__global__ void __launch_bounds__(256, 1) dsmem_bench(int mode, int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t sBuf = base;                 // 64 KiB
    const uint32_t sBar = base + 65536;         // barrier (in the destination: counts bytes)
    const uint32_t sBack = sBar + 8;            // barrier (in the source: destination says "received")
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        mbar_init(sBar, 1);
        mbar_init(sBack, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = i;
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync_all();
    uint32_t rBuf, rBar, rBack;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rBuf) : "r"(sBuf), "r"(1));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rBar) : "r"(sBar), "r"(1));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rBack) : "r"(sBack), "r"(0));
    long long t0 = clock64();
    if (rank == 0) {
        for (int r = 0; r < reps; ++r) {
            if (mode == 0) {
                for (int k = 0; k < 16; ++k) {
                    const uint32_t off = (k * 256 + threadIdx.x) * 16;
                    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rBuf + off), "r"(r), "r"(k), "r"(off), "r"(1) : "memory");
                }
                asm volatile("fence.acq_rel.cluster;" ::: "memory");
                __syncthreads();
                if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rBar) : "memory");
            } else if (mode == 1) {
                for (int k = 0; k < 16; ++k) {
                    const uint32_t off = (k * 256 + threadIdx.x) * 16;
                    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                                 ::"r"(rBuf + off), "r"(r), "r"(k), "r"(off), "r"(1), "r"(rBar) : "memory");
                }
            } else {
                if (threadIdx.x == 0) {
                    for (int k = 0; k < 4; ++k)
                        asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                     ::"r"(rBuf + k * 16384), "r"(sBuf + k * 16384), "r"(16384), "r"(rBar) : "memory");
                }
            }
            // wait until the destination has everything (so at most one 64 KiB block is in flight, as in the kernel)
            mbar_wait(sBack, r & 1);
        }
    } else {
        for (int r = 0; r < reps; ++r) {
            if (threadIdx.x == 0) {
                if (mode != 0) mbar_arrive_expect_tx(sBar, 65536);
                mbar_wait(sBar, r & 1);
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(rBack) : "memory");
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && rank == 0) out[blockIdx.x / 2] = t1 - t0;
    __syncthreads();
    cluster_sync_all();
}

static void run_dsmem(int mode, long long* d_out) {
    cudaFuncSetAttribute(dsmem_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = 65536 + 64;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int reps = 64;
    cudaMemset(d_out, 0, 148 * sizeof(long long));
    cudaError_t e = cudaLaunchKernelEx(&cfg, dsmem_bench, mode, reps, d_out);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        printf("dsmem mode %d FAILED: %s / %s\n", mode, cudaGetErrorString(e), cudaGetErrorString(e2));
        return;
    }
    std::vector<long long> h(74);
    cudaMemcpy(h.data(), d_out, 74 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    printf("dsmem mode %d: %.0f cycles per 64 KiB (median over pairs) -> %.1f B/cycle  [min %.0f max %.0f]\n", mode,
           (double)h[37] / reps, 65536.0 * reps / h[37], (double)h[0] / reps, (double)h[73] / reps);
}


// ---- issue-pattern test (cg2, M = 256, N = 256): groups of 4 MMAs with a commit after each group.
// pat 0: commits only.  pat 1: ring of `ns` stages -- the issuer waits full[s] before a group and commits empty[s]
// after it; a producer thread waits empty[s] and re-arms full[s] at once (no data movement).
__global__ void __launch_bounds__(128, 1) ring_bench(int pat, int ns, int groups, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t sA = base, sB = base + 4 * 16384;
    const uint32_t sBar = sB + 4 * 32768;
    const uint32_t sTp = sBar + 256;
    volatile uint32_t* tp = reinterpret_cast<volatile uint32_t*>(smem_raw + (sTp - base));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    auto full = [&](int i) { return sBar + 8 * i; };
    auto empty = [&](int i) { return sBar + 8 * (8 + i); };
    const uint32_t done = sBar + 8 * 16;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), 1); }
        mbar_init(done, 1);
        *reinterpret_cast<volatile int*>(smem_raw + (sBar + 8 * 20 - base)) = 0;
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < (4 * 16384 + 4 * 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3c003c00u;
    if (warp == 1) tmem_alloc_pair(sTp, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tp;
    if (warp == 2 && lane == 0 && rank == 0 && pat == 3) {
        // helper: waits the full barriers in order and publishes how many groups are ready
        int st = 0; uint32_t ph = 0;
        for (int g = 0; g < groups; ++g) {
            mbar_wait(full(st), ph);
            *reinterpret_cast<volatile int*>(smem_raw + (sBar + 8 * 20 - base)) = g + 1;
            if (++st == ns) { st = 0; ph ^= 1; }
        }
    }
    if (warp == 0 && lane == 0 && rank == 0 && pat != 0) {
        // producer: re-arm full[s] as soon as the stage was released
        int st = 0; uint32_t ph = 0;
        for (int g = 0; g < groups; ++g) {
            mbar_wait(empty(st), ph ^ 1);
            mbar_arrive(full(st));
            if (++st == ns) { st = 0; ph ^= 1; }
        }
    } else if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = make_idesc(0, 0, 0, 256, 256);
        int st = 0; uint32_t ph = 0;
        int ready = 0;
        long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            if (pat == 1) { mbar_wait(full(st), ph); tc_fence_after(); }
            if (pat == 2) { mbar_wait(full(st), ph); }
            if (pat == 4) {
                uint32_t ok = 0;
                while (!ok) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(full(st)), "r"(ph) : "memory");
                }
                tc_fence_after();
            }
            if (pat == 3) {
                // readiness counter published by a helper warp: touch shared memory only when the cached value is used up
                while (ready <= g) ready = *reinterpret_cast<volatile int*>(smem_raw + (sBar + 8 * 20 - base));
                tc_fence_after();
            }
            const int c = g & 3;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_f16_ss_pair(tmem, desc_kmajor(sA + c * 16384, k), desc_kmajor(sB + c * 32768, k), idesc, (g | k) != 0);
            umma_commit_pair(empty(st));
            if (++st == ns) { st = 0; ph ^= 1; }
        }
        umma_commit_pair(done);
        mbar_wait(done, 0);
        out[blockIdx.x / 2] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) { tc_fence_after(); tmem_dealloc_pair(tmem, 512); }
}

static void run_ring(int pat, int ns, long long* d_out) {
    cudaFuncSetAttribute(ring_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(148);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 4 * 16384 + 4 * 32768 + 512;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int groups = 1024;
    cudaMemset(d_out, 0, 148 * sizeof(long long));
    cudaError_t e = cudaLaunchKernelEx(&cfg, ring_bench, pat, ns, groups, d_out);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("ring pat %d FAILED: %s / %s\n", pat, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    std::vector<long long> h(74);
    cudaMemcpy(h.data(), d_out, 74 * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    printf("ring pat %d ns %d: %.1f cycles per MMA (4 per group, commit per group)\n", pat, ns, (double)h[37] / groups / 4);
}


// ---- TMA ingest: every CTA streams [128 rows x 64 cols] 16-bit boxes (16 KiB) of an L2-resident matrix into a ring of
// shared-memory stages; nothing consumes the data except a thread that recycles the stage at once.
//   mc = 0: unicast, each CTA fetches its own 16 KiB per stage
//   mc = 1: clusters of 2: each CTA fetches 8 KiB (64 rows) and multicasts it to both CTAs (each still receives 16 KiB)
//   mc = 2: clusters of 4, pairs (0,2) and (1,3) share: each CTA fetches 8 KiB and multicasts to its partner and itself
__global__ void __launch_bounds__(128, 1) tma_bench(const __grid_constant__ CUtensorMap map16, const __grid_constant__ CUtensorMap map8,
                                                    int mc, int ns, int stages_total, int rows_total, long long* out, uint8_t* scratch, int st_every) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    const uint32_t sBar = base + 12 * 16384;
    auto full = [&](int i) { return sBar + 8 * i; };
    auto empty = [&](int i) { return sBar + 8 * (16 + i); };
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) {
        for (int i = 0; i < ns; ++i) { mbar_init(full(i), 1); mbar_init(empty(i), mc ? 2 : 1); }
        fence_barrier_init();
    }
    __syncthreads();
    if (mc) cluster_sync_all();
    const int cta = blockIdx.x;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        // producer
        int st = 0; uint32_t ph = 0;
        const uint32_t partner = (mc == 1) ? (rank ^ 1) : (rank ^ 2);
        const uint16_t mask = (uint16_t)((1u << rank) | (1u << partner));
        const int half = (mc == 1) ? (int)(rank & 1) : (int)((rank >> 1) & 1);
        const int group = (mc == 0) ? cta : (mc == 1 ? cta / 2 : (cta / 4) * 2 + (int)(rank & 1));
        for (int g = 0; g < stages_total; ++g) {
            mbar_wait(empty(st), ph ^ 1);
            mbar_arrive_expect_tx(full(st), 16384);
            const int col = (g & 7) * 64;
            const int row = ((group * 37 + (g >> 3)) * 128) % rows_total;
            if (mc == 0) {
                tma_load_2d(base + st * 16384, &map16, full(st), col, row);
            } else {
                asm volatile(
                    "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                    ::"r"(base + st * 16384 + half * 8192), "l"(reinterpret_cast<uint64_t>(&map8)), "r"(full(st)), "r"(col), "r"(row + half * 64), "h"(mask)
                    : "memory");
            }
            if (++st == ns) { st = 0; ph ^= 1; }
        }
    } else if (threadIdx.x == 32) {
        // consumer: release each stage as soon as it is full (to both producers that write into it)
        int st = 0; uint32_t ph = 0;
        const uint32_t partner = (mc == 1) ? (rank ^ 1) : (rank ^ 2);
        for (int g = 0; g < stages_total; ++g) {
            mbar_wait(full(st), ph);
            if (st_every && (g % st_every) == 0) {
                // send the stage back out to an L2-resident scratch (bulk store), wait until the source may be reused
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(scratch + ((size_t)cta * 4 + (g & 3)) * 16384), "r"(base + st * 16384), "r"(16384) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (mc) {
                mbar_arrive(empty(st));
                mbar_arrive_cluster(empty(st), partner);
            } else {
                mbar_arrive(empty(st));
            }
            if (++st == ns) { st = 0; ph ^= 1; }
        }
        out[cta] = clock64() - t0;
    }
    __syncthreads();
    if (mc) cluster_sync_all();
}

static uint8_t* g_scratch = nullptr;
static void run_tma(int mc, int ns, long long* d_out, int st_every = 0) {
    typedef CUresult (*PFN)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                            const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                            CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static PFN enc = nullptr;
    static void* buf = nullptr;
    const int rows = 8192, cols = 512;      // 8 MiB of 16-bit values: stays in L2
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q);
        enc = (PFN)ptr;
        cudaMalloc(&buf, (size_t)rows * cols * 2);
        cudaMemset(buf, 0, (size_t)rows * cols * 2);
        cudaMalloc(&g_scratch, (size_t)148 * 4 * 16384);
    }
    CUtensorMap m16, m8;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t estr[2] = {1, 1};
    cuuint32_t box16[2] = {64, 128}, box8[2] = {64, 64};
    enc(&m16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box16, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    enc(&m8, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, buf, dims, strides, box8, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaLaunchConfig_t cfg{};
    const int csz = mc == 0 ? 1 : (mc == 1 ? 2 : 4);
    const int grid = mc == 2 ? 132 : 148;
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = 12 * 16384 + 512;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = csz;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const int stages_total = 2048;
    cudaMemset(d_out, 0, 148 * sizeof(long long));
    cudaError_t e = cudaLaunchKernelEx(&cfg, tma_bench, m16, m8, mc, ns, stages_total, rows, d_out, g_scratch, st_every);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) { printf("tma mc %d FAILED: %s / %s\n", mc, cudaGetErrorString(e), cudaGetErrorString(e2)); return; }
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    const double cyc = (double)h[grid / 2] / stages_total;
    printf("tma ingest mc=%d ring=%2d stages, bulk store of every %d-th stage: %.0f cycles per 16 KiB stage -> %.1f B/cycle/SM received (%d SMs)\n", mc, ns, st_every, cyc, 16384.0 / cyc, grid);
}

__global__ void __cluster_dims__(1, 1, 1) dummy_kernel(int* x) { if (x) *x = 1; }
template <int CS>
static void occupancy() {
    auto kern = bench<2>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CS * 64);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = 226 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    printf("cluster size %d: max active clusters %d (%d SMs) [%s]\n", CS, n, n * CS, cudaGetErrorString(e));
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'm') {            // operand-major / mixed-format variants only
        long long* d_out;
        cudaMalloc(&d_out, 148 * sizeof(long long));
        const int n = 4096;
        run<2>("cg2 SS K-major (warm-up)", 256, 256, 0, 4 * n, 0, d_out);
        run<2>("cg2 SS K-major (reference)", 256, 256, 0, n, 0, d_out);
        run<2>("cg2 SS A MN-major", 256, 256, 4, n, 0, d_out);
        run<2>("cg2 SS B MN-major", 256, 256, 5, n, 0, d_out);
        run<2>("cg2 SS A and B MN-major", 256, 256, 6, n, 0, d_out);
        run<1>("cg1 SS A MN-major", 128, 256, 4, n, 0, d_out);
        run<1>("cg1 SS B MN-major", 128, 256, 5, n, 0, d_out);
        run<2>("cg2 SS K-major (reference)", 256, 256, 0, n, 0, d_out);
        if (argv[1][1] == 'x') run<2>("cg2 SS A bf16, B f16", 256, 256, 7, n, 0, d_out);   // illegal instruction on sm_100a
        return 0;
    }
    { long long* d; cudaMalloc(&d, 148 * sizeof(long long)); run_dsmem(0, d); run_dsmem(1, d); run_dsmem(2, d); cudaFree(d); }
    { long long* d; cudaMalloc(&d, 148 * sizeof(long long)); run_ring(0, 4, d); run_ring(1, 4, d); run_ring(2, 4, d); run_ring(3, 4, d); run_ring(4, 4, d); cudaFree(d); }
    { long long* d; cudaMalloc(&d, 148 * sizeof(long long)); run_tma(0, 8, d); run_tma(0, 8, d, 2); run_tma(0, 8, d, 1); run_tma(1, 8, d); cudaFree(d); }
    occupancy<2>();
    occupancy<4>();
    occupancy<8>();
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const int n = 4096;
    run<1>("cg1 SS", 128, 256, 0, n, 0, d_out);
    run<1>("cg1 SS", 128, 128, 0, n, 0, d_out);
    run<1>("cg1 SS", 128, 64, 0, n, 0, d_out);
    run<1>("cg1 SS", 64, 256, 0, n, 0, d_out);
    run<2>("cg2 SS", 256, 256, 0, n, 0, d_out);
    run<2>("cg2 SS", 256, 128, 0, n, 0, d_out);
    run<2>("cg2 SS", 256, 64, 0, n, 0, d_out);
    run<2>("cg2 SS", 128, 256, 0, n, 0, d_out);
    run<2>("cg2 SS", 128, 128, 0, n, 0, d_out);
    run<1>("cg1 TS", 128, 256, 1, n, 0, d_out);
    run<1>("cg1 TS", 128, 128, 1, n, 0, d_out);
    run<2>("cg2 TS", 256, 256, 1, n, 0, d_out);
    run<2>("cg2 TS", 256, 128, 1, n, 0, d_out);
    run<2>("cg2 TS", 128, 256, 1, n, 0, d_out);
    run<2>("cg2 SS + 8 warps tcgen05.ld", 256, 256, 2, n, 8, d_out);
    run<2>("cg2 SS + 4 warps tcgen05.ld", 256, 256, 2, n, 4, d_out);
    run<1>("tcgen05.ld only", 128, 256, 3, n, 1, d_out);
    run<1>("tcgen05.ld only", 128, 256, 3, n, 4, d_out);
    run<1>("tcgen05.ld only", 128, 256, 3, n, 8, d_out);
    return 0;
}
