// Decode-time joint: the reference's greedy search evaluates the joint on ONE frame at a time,
//   logits = project_layer(tanh(forward_layer(cat(enc[t], dec))));  pred = argmax(softmax(logits)).item()
// (/root/reference/tt/model.py:70-90, tt_espnet/model.py:83-106), i.e. ~10 kernel launches and a host synchronisation per
// frame.  Between two emitted labels the decoder state -- hence the predictor half of the first layer -- does not change,
// so the frames up to the next non-blank prediction can all be scored against it at once:
//   scan_kernel   for n <= 64 consecutive frames: A[f] = tanh(eproj[f] + pvec), z[f] = A[f] . W_out^T + b_out in plain fp32
//                 FMAs (decoding compares logits, so no 16-bit operands here), per-frame argmax merged across the
//                 vocabulary tiles with one 64-bit atomicMax (value bits, then the LOWEST index on ties, like torch.argmax)
//   pick_kernel   first frame whose argmax is not the blank + that label -> 2 ints the host reads with ONE synchronisation
// A register-tiled SGEMM: block = 64 vocabulary rows x 64 frames, 256 threads x (4 frames x 4 rows), K chunks of 32.
#include <limits.h>

#include "ttx_common.cuh"

namespace ttx {

constexpr int kDF = 64;          // frames per call
constexpr int kDV = 64;          // vocabulary rows per block
constexpr int kDK = 32;          // K chunk

// monotone map float -> uint32 (larger float <-> larger key; NaN excluded by the caller's finite inputs)
__device__ __forceinline__ uint32_t float_key(float x) {
    const uint32_t b = __float_as_uint(x);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256)
scan_kernel(const float* __restrict__ eproj, int ld_e, const float* __restrict__ pvec, const float* __restrict__ w_out,
            const float* __restrict__ b_out, int n, int H, int V, unsigned long long* __restrict__ best) {
    __shared__ float As[kDK][kDF + 4];          // As[k][frame]
    __shared__ float Ws[kDK][kDV + 4];          // Ws[k][vocabulary row]
    const int v0 = blockIdx.x * kDV;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;      // 4 vocabulary rows tx * 4 .., 4 frames ty * 4 ..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < H; k0 += kDK) {
        // stage the chunk: 64 x 32 activations (tanh on the fly) and 64 x 32 weights, 8 of each per thread
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int idx = it * 256 + threadIdx.x;
            const int r = idx >> 5, k = idx & 31;                 // row (frame / vocabulary row), k inside the chunk
            float a = 0.f, w = 0.f;
            if (k0 + k < H) {
                if (r < n) a = tanhf(eproj[(size_t)r * ld_e + k0 + k] + pvec[k0 + k]);
                if (v0 + r < V) w = w_out[(size_t)(v0 + r) * H + k0 + k];
            }
            As[k][r] = a;
            Ws[k][r] = w;
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < kDK; ++k) {
            const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w}, w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
    // per frame: best (logit, lowest index) over this thread's 4 rows, then over the 16 threads of the frame group
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int f = ty * 4 + i;
        unsigned long long key = 0ull;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = v0 + tx * 4 + j;
            if (v < V) {
                const float z = acc[i][j] + b_out[v];
                const unsigned long long kk = ((unsigned long long)float_key(z) << 32) | (0xFFFFFFFFu - (uint32_t)v);
                key = kk > key ? kk : key;
            }
        }
#pragma unroll
        for (int o = 8; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
            key = other > key ? other : key;
        }
        if (tx == 0 && f < n) atomicMax(best + f, key);
    }
}

__global__ void pick_kernel(const unsigned long long* __restrict__ best, int n, int blank, int* __restrict__ out) {
    // out[0] = first frame (0 .. n-1) whose argmax is not the blank, or n; out[1] = that label (or the blank);
    // out[2 + f] = argmax of frame f
    __shared__ int first;
    if (threadIdx.x == 0) first = n;
    __syncthreads();
    for (int f = threadIdx.x; f < n; f += blockDim.x) {
        const int v = (int)(0xFFFFFFFFu - (uint32_t)(best[f] & 0xFFFFFFFFull));
        out[2 + f] = v;
        if (v != blank) atomicMin(&first, f);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out[0] = first;
        out[1] = first < n ? out[2 + first] : blank;
    }
}

int launch_decode_scan(const float* eproj, int ld_e, const float* pvec, const float* w_out, const float* b_out, int n, int H,
                       int V, int blank, unsigned long long* scratch, int* out, cudaStream_t s) {
    TTX_CUDA_OK(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * kDF, s));
    scan_kernel<<<(V + kDV - 1) / kDV, 256, 0, s>>>(eproj, ld_e, pvec, w_out, b_out, n, H, V, scratch);
    pick_kernel<<<1, 64, 0, s>>>(scratch, n, blank, out);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ttx

// ------------------------------------------------------------------------------------------- input pipeline
// SpecAugment-style masking of a (B, T, F) feature batch, in place: the reference applies ten frequency masks and ten time
// masks as twenty slice assignments `inputs[:, :, f0:f0+f] = 0` / `inputs[:, t0:t0+t, :] = 0`
// (/root/reference/tt/utils.py:297-329, called train.py:41-44), twenty launches over the whole batch.  Here one launch
// zeroes exactly the union of those slices; the mask positions come from the caller (drawn on the host with the
// reference's own random-number calls, so the result is bit-identical).  masks: int32 [n_masks][3] = {axis (1 = time,
// 2 = frequency), start, width}.
namespace ttx {

constexpr int kMaxMasks = 64;
struct MaskList {
    int n;
    int axis[kMaxMasks], start[kMaxMasks], width[kMaxMasks];
};

__global__ void spec_mask_kernel(float* __restrict__ x, int B, int T, int F, long long ld_b, long long ld_t, const MaskList m) {
    // grid = (T, B); a frame that a time mask covers is zeroed whole, otherwise only its masked frequency bins
    const int t = blockIdx.x, b = blockIdx.y;
    float* row = x + b * ld_b + t * ld_t;
    bool whole = false;
    for (int i = 0; i < m.n; ++i) whole |= (m.axis[i] == 1 && t >= m.start[i] && t < m.start[i] + m.width[i]);
    if (whole) {
        for (int f = threadIdx.x; f < F; f += blockDim.x) row[f] = 0.f;
        return;
    }
    for (int i = 0; i < m.n; ++i)
        if (m.axis[i] == 2)
            for (int f = m.start[i] + threadIdx.x; f < min(F, m.start[i] + m.width[i]); f += blockDim.x) row[f] = 0.f;
}

int launch_spec_mask(float* x, int B, int T, int F, long long ld_b, long long ld_t, const int* masks_host, int n_masks,
                     cudaStream_t s) {
    MaskList m{};
    m.n = n_masks;
    for (int i = 0; i < n_masks; ++i) {
        m.axis[i] = masks_host[3 * i];
        m.start[i] = masks_host[3 * i + 1];
        m.width[i] = masks_host[3 * i + 2];
    }
    spec_mask_kernel<<<dim3(T, B), 128, 0, s>>>(x, B, T, F, ld_b, ld_t, m);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ttx

// ------------------------------------------------------------------------------------------- argument checks
// What warprnnt_pytorch's certify_inputs needs from the device (upstream reads max(act_lens), max(label_lens) on the
// host), plus the batch's real lattice size, in ONE launch and one 56-byte read instead of a dozen small reductions:
// out[0..6] = max T, max U, min T, min U, 128-row lattice tiles, labels outside [0, V) inside their utterance (count),
// elements of the diagonal-major lattice arrays.
namespace ttx {

__global__ void check_inputs_kernel(const int* __restrict__ labels, int label_stride, const int* __restrict__ act_lens,
                                    const int* __restrict__ label_lens, int B, int V, long long* __restrict__ out) {
    __shared__ long long red[7][8];
    long long v[7] = {LLONG_MIN, LLONG_MIN, LLONG_MAX, LLONG_MAX, 0, 0, 0};
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const long long t = act_lens[b], u = label_lens[b];
        v[0] = max(v[0], t); v[1] = max(v[1], u); v[2] = min(v[2], t); v[3] = min(v[3], u);
        v[4] += (t * (u + 1) + 127) / 128;
        v[6] += (t + u) * ((u + 4) / 4 * 4);                  // (T + U1 - 1) * pitch(U1)
    }
    if (labels != nullptr && label_stride > 0)
        for (long long i = threadIdx.x; i < (long long)B * label_stride; i += blockDim.x) {
            const int b = (int)(i / label_stride), u = (int)(i - (long long)b * label_stride);
            if (u < label_lens[b]) {
                const int l = labels[i];
                v[5] += (l < 0 || l >= V) ? 1 : 0;
            }
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        long long x = v[k];
        for (int o = 16; o; o >>= 1) {
            const long long y = __shfl_xor_sync(0xffffffffu, x, o);
            x = k < 2 ? max(x, y) : k < 4 ? min(x, y) : x + y;
        }
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < 7) {
        const int k = threadIdx.x;
        long long x = red[k][0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) x = k < 2 ? max(x, red[k][w]) : k < 4 ? min(x, red[k][w]) : x + red[k][w];
        out[k] = x;
    }
}

int launch_check_inputs(const int* labels, int label_stride, const int* act_lens, const int* label_lens, int B, int V,
                        long long* out, cudaStream_t s) {
    check_inputs_kernel<<<1, 256, 0, s>>>(labels, label_stride, act_lens, label_lens, B, V, out);
    TTX_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace ttx
