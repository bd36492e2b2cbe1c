// extern "C" boundary of libttx.so (declared in include/ttx.h): argument checks + kernel launches.
#include <stdarg.h>
#include <stdio.h>

#include "../../include/ttx.h"
#include "ttx_common.cuh"

namespace ttx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// launchers (ttx_small.cu / ttx_joint_mma.cu)
int launch_prep(const int*, const int*, int, int, int, int, int*, cudaStream_t);
int launch_cast_w(const float*, const float*, int, int, int, bool, float*, void*, float*, void*, cudaStream_t);
int launch_joint_act(const float*, const float*, const int*, const int*, const int*, const int*, int, int, int, int,
                     int, int, int, bool, void*, int*, void*, cudaStream_t);
int launch_lattice(const float*, const float*, const int*, const int*, const int*, int, int, int, size_t, float*, double*,
                   double*, float*, double*, cudaStream_t);
int launch_grad_prep(const float*, const float*, const float*, const double*, const double*, const double*,
                     const float*, float*, const int*, const int*, const int*, const int*, int, int, int, float4*,
                     float*, cudaStream_t);
int launch_reduce(const float*, const float4*, const int*, const float*, const float*, int, const float*, const float*,
                  const int*, const int*, const int*, int, int, int, int, float*, float*, cudaStream_t);
bool fwd_grad_supported_h(int H);
int launch_joint_fwd_grad(const void*, const void*, const void*, uint64_t, int, int, int, int, bool, const int*,
                          const float*, const float*, const int*, int, float*, float*, float*, float*, void*, size_t,
                          cudaStream_t);
size_t joint_workspace_bytes(int which, int n_tiles_ub, int H, int V);
int launch_kept_prepare(const void*, const float4*, const int*, const float*, const float*, const float*,
                        const float*, const int*, const int*, const int*, int, int, int, int, int, bool, size_t, void*,
                        float*, float*, int, cudaStream_t);
int launch_dense_lse(const float*, const int*, const int*, const int*, const int*, int, int, int, int, int, int, int,
                     float*, float*, float*, int*, cudaStream_t);
int launch_dense_grad(const float*, const float4*, const int*, const float*, const int*, const int*, const int*, int,
                      int, int, int, int, float*, cudaStream_t);
int launch_transpose16(const void*, void*, int, int, const int*, cudaStream_t);
int launch_rows_lse(const float*, int, int, int, const float*, const float*, const int*, int, float*, float*, float*,
                    cudaStream_t);
int launch_rows_grad(const float*, const float4*, const int*, const float*, const float*, int, int, int, int, bool,
                     void*, cudaStream_t);
bool mma_supported_h(int H);
int launch_joint_fwd(const void*, const void*, uint64_t, int, int, int, int, bool, const int*, const float*,
                     const float*, const int*, int, float*, float*, float*, cudaStream_t);
int launch_joint_bwd(const void*, const void*, const void*, const void*, uint64_t, int, int, int, int, bool,
                     const int*, const float*, const float*, const int*, int, const float4*, float*, float*, float*,
                     int, void*, size_t, cudaStream_t);

bool wide_supported_h(int H);
int launch_wide_sp(const void*, const void*, uint64_t, int, int, int, int, int, bool, const int*, const float*, const float*,
                   const int*, int, float*, float*, float*, float*, float*, void*, uint64_t, int*, cudaStream_t);
int launch_wide_pw(const void*, uint64_t, const void*, int, int, int, int, int, bool, const int*, const float*, const float*,
                   float*, cudaStream_t);
int launch_wide_dw(const void*, uint64_t, const void*, uint64_t, int, int, int, int, int, bool, const int*, const float*,
                   float*, float*, cudaStream_t);

int launch_proj_fwd(const float*, int, const float*, int, const float*, int, int, int, float*, int, cudaStream_t);
int launch_proj_bwd_x(const float*, int, const float*, int, int, int, int, float*, int, cudaStream_t);
int launch_proj_bwd_w(const float*, int, const float*, int, int, int, int, float*, int, float*, cudaStream_t);

int launch_decode_scan(const float*, int, const float*, const float*, const float*, int, int, int, int, unsigned long long*,
                       int*, cudaStream_t);

int launch_spec_mask(float*, int, int, int, long long, long long, const int*, int, cudaStream_t);

int launch_band_attn_fwd(const float*, const float*, const float*, const float*, int, int, int, int, int, int, int, float, int,
                         const int*, float*, float*, cudaStream_t);
int launch_band_attn_bwd(const float*, const float*, const float*, const float*, const float*, int, int, int, int, int, int, int,
                         float, int, const int*, float*, float*, float*, float*, float*, float*, cudaStream_t);

int launch_check_inputs(const int*, int, const int*, const int*, int, int, long long*, cudaStream_t);

static int enter(int device) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        set_error("cudaSetDevice(%d) failed: %s", device, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

}  // namespace ttx

using namespace ttx;

#define TTX_REQUIRE(cond, ...)     \
    do {                           \
        if (!(cond)) {             \
            set_error(__VA_ARGS__); \
            return 1;              \
        }                          \
    } while (0)
#define TTX_ENTER(device)                 \
    do {                                  \
        if (int _rc = enter(device)) return _rc; \
    } while (0)

extern "C" {

int ttx_version(void) { return 1; }

const char* ttx_last_error(void) { return g_err; }

int ttx_supported_h(int H) { return mma_supported_h(H) ? 1 : 0; }

int64_t ttx_tiles_upper_bound(int B, int T, int U1) {
    if (B <= 0 || T <= 0 || U1 <= 0) return 0;
    return (int64_t)B * (((int64_t)T * U1 + kTile - 1) / kTile);
}

int64_t ttx_meta_ints(int B, int64_t n_tiles_ub) { return kMetaHdr + 2 * ((int64_t)B + 1) + n_tiles_ub; }

int64_t ttx_lattice_elems_upper_bound(int B, int T, int U1) {
    if (B <= 0 || T <= 0 || U1 <= 0) return 0;
    return (int64_t)B * lat_elems(T, U1);
}

int ttx_prepare(const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int64_t n_tiles_ub,
                int32_t* meta, int device, void* stream) {
    TTX_REQUIRE(act_lens && label_lens && meta, "ttx_prepare: null pointer");
    TTX_REQUIRE(B > 0 && T > 0 && U1 > 0, "ttx_prepare: bad shape B=%d T=%d U1=%d", B, T, U1);
    // any bound is accepted (callers that know the batch's real tile count size their buffers by it); a bound smaller
    // than the count found on the device is flagged in meta[1] like bad lengths and no kernel does any work
    TTX_REQUIRE(n_tiles_ub >= 1 && n_tiles_ub < (1 << 24), "ttx_prepare: tile bound %lld out of range",
                (long long)n_tiles_ub);
    TTX_ENTER(device);
    return launch_prep(act_lens, label_lens, B, T, U1, (int)n_tiles_ub, meta, (cudaStream_t)stream);
}

int ttx_cast_weight(const float* w_out, const float* b_out, int V, int H, int bf16, float* scal, void* w16,
                    float* bias2, void* w16t, int device, void* stream) {
    TTX_REQUIRE(w_out && b_out && scal && w16 && bias2, "ttx_cast_weight: null pointer");
    TTX_REQUIRE(V > 0 && H > 0 && H % 8 == 0, "ttx_cast_weight: bad shape V=%d H=%d", V, H);
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_cast_w(w_out, b_out, V, Vpad, H, bf16 != 0, scal, w16, bias2, w16t, (cudaStream_t)stream);
}

int ttx_joint_act(const float* eproj, const float* pproj, const int32_t* labels, const int32_t* act_lens,
                  const int32_t* label_lens, const int32_t* meta, int B, int T, int U1, int H, int label_stride, int V,
                  int64_t n_tiles_ub, int bf16, void* a16, int32_t* row_label, void* a16t, int device, void* stream) {
    TTX_REQUIRE(eproj && pproj && act_lens && label_lens && meta && a16 && row_label, "ttx_joint_act: null pointer");
    TTX_REQUIRE(labels || U1 == 1, "ttx_joint_act: labels is null");
    TTX_REQUIRE(H > 0 && H % 64 == 0, "ttx_joint_act: H=%d must be a positive multiple of 64", H);
    TTX_ENTER(device);
    TTX_REQUIRE(V >= 0, "ttx_joint_act: V=%d", V);
    return launch_joint_act(eproj, pproj, labels, act_lens, label_lens, meta, B, T, U1, H, label_stride, V,
                            (int)n_tiles_ub, bf16 != 0, a16, row_label, a16t, (cudaStream_t)stream);
}

int ttx_joint_lse_fwd(const void* a16, const void* w16, const float* bias2, const float* scal,
                      const int32_t* row_label, const int32_t* meta, int64_t n_tiles_ub, int H, int V, int blank,
                      int bf16, float* lse, float* lp_blank, float* lp_label, int device, void* stream) {
    TTX_REQUIRE(a16 && w16 && bias2 && scal && row_label && meta && lse && lp_blank && lp_label,
                "ttx_joint_lse_fwd: null pointer");
    TTX_REQUIRE(mma_supported_h(H), "ttx_joint_lse_fwd: joint width H=%d is not supported by the tensor-core path", H);
    TTX_REQUIRE(V > 0 && blank >= 0 && blank < V, "ttx_joint_lse_fwd: bad V=%d / blank=%d", V, blank);
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_joint_fwd(a16, w16, (uint64_t)n_tiles_ub * kTile, (int)n_tiles_ub, H, V, Vpad, bf16 != 0, meta,
                            bias2, scal, row_label, blank, lse, lp_blank, lp_label, (cudaStream_t)stream);
}

int ttx_lattice_fwd_bwd(const float* lp_blank, const float* lp_label, const int32_t* act_lens,
                        const int32_t* label_lens, const int32_t* meta, int B, int U1, int64_t n_tiles_ub,
                        int64_t lat_elems, float* lat_ws, double* alpha, double* beta, float* costs, double* ll_beta,
                        int device, void* stream) {
    TTX_REQUIRE(lp_blank && lp_label && act_lens && label_lens && meta && lat_ws && alpha && beta && costs && ll_beta,
                "ttx_lattice_fwd_bwd: null pointer");
    TTX_REQUIRE(B > 0 && U1 > 0 && n_tiles_ub >= 1 && lat_elems >= 4 && lat_elems % 4 == 0 && lat_elems < (1ll << 31),
                "ttx_lattice_fwd_bwd: bad shape B=%d U1=%d tiles=%lld lattice elements=%lld", B, U1,
                (long long)n_tiles_ub, (long long)lat_elems);
    TTX_ENTER(device);
    return launch_lattice(lp_blank, lp_label, act_lens, label_lens, meta, B, U1, (int)n_tiles_ub, (size_t)lat_elems,
                          lat_ws, alpha, beta, costs, ll_beta, (cudaStream_t)stream);
}

int ttx_grad_coeffs(const float* lse, const float* lp_blank, const float* lp_label, const double* alpha,
                    const double* beta, const double* ll_beta, const float* grad_costs, float* scal,
                    const int32_t* row_label, const int32_t* act_lens, const int32_t* label_lens,
                    const int32_t* meta, int B, int blank, int64_t n_tiles_ub, void* rowmeta, float* d_b_out,
                    int device, void* stream) {
    TTX_REQUIRE(lse && lp_blank && lp_label && alpha && beta && ll_beta && grad_costs && scal && row_label && rowmeta,
                "ttx_grad_coeffs: null pointer");
    TTX_ENTER(device);
    return launch_grad_prep(lse, lp_blank, lp_label, alpha, beta, ll_beta, grad_costs, scal, row_label, act_lens,
                            label_lens, meta, B, blank, (int)n_tiles_ub, (float4*)rowmeta, d_b_out,
                            (cudaStream_t)stream);
}

int ttx_rows_lse(const float* z, int rows, int Vpad, int V, const float* bias2, const float* scal,
                 const int32_t* row_label, int blank, float* lse, float* lp_blank, float* lp_label, int device,
                 void* stream) {
    TTX_REQUIRE(z && bias2 && scal && row_label && lse && lp_blank && lp_label, "ttx_rows_lse: null pointer");
    TTX_REQUIRE(rows > 0 && V > 0 && Vpad >= V && blank >= 0 && blank < V, "ttx_rows_lse: bad shape");
    TTX_ENTER(device);
    return launch_rows_lse(z, rows, Vpad, V, bias2, scal, row_label, blank, lse, lp_blank, lp_label, (cudaStream_t)stream);
}

int ttx_rows_grad(const float* z, const void* rowmeta, const int32_t* row_label, const float* bias2,
                  const float* scal, int rows, int Vpad, int V, int blank, int bf16, void* q, int device, void* stream) {
    TTX_REQUIRE(z && rowmeta && row_label && bias2 && scal && q, "ttx_rows_grad: null pointer");
    TTX_REQUIRE(rows > 0 && V > 0 && Vpad >= V && Vpad % 2 == 0, "ttx_rows_grad: bad shape");
    TTX_ENTER(device);
    return launch_rows_grad(z, (const float4*)rowmeta, row_label, bias2, scal, rows, Vpad, V, blank, bf16 != 0, q,
                            (cudaStream_t)stream);
}

int ttx_transpose16(const void* in, void* out, int rows, int cols, const int32_t* meta, int device, void* stream) {
    TTX_REQUIRE(in && out, "ttx_transpose16: null pointer");
    TTX_REQUIRE(rows > 0 && cols > 0 && rows % 64 == 0 && cols % 64 == 0, "ttx_transpose16: rows=%d cols=%d must be multiples of 64", rows, cols);
    TTX_ENTER(device);
    return launch_transpose16(in, out, rows, cols, meta, (cudaStream_t)stream);
}

int ttx_joint_grad(const void* a16, const void* w16, const void* a16t, const void* w16t, const float* bias2,
                   const float* scal,
                   const int32_t* row_label, const int32_t* meta, const void* rowmeta, int64_t n_tiles_ub, int H,
                   int V, int blank, int bf16, float* d_act, float* d_w_out, float* d_b_out, int splits,
                   void* workspace, int64_t workspace_bytes, int device, void* stream) {
    TTX_REQUIRE(a16 && w16 && bias2 && scal && row_label && meta && rowmeta, "ttx_joint_grad: null pointer");
    TTX_REQUIRE((d_w_out == nullptr) == (d_b_out == nullptr), "ttx_joint_grad: d_w_out and d_b_out go together");
    TTX_REQUIRE(mma_supported_h(H), "ttx_joint_grad: joint width H=%d is not supported by the tensor-core path", H);
    TTX_REQUIRE(V > 0 && blank >= 0 && blank < V, "ttx_joint_grad: bad V=%d / blank=%d", V, blank);
    TTX_REQUIRE(splits >= 1 && splits <= 65535, "ttx_joint_grad: bad splits=%d", splits);
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_joint_bwd(a16, w16, a16t, w16t, (uint64_t)n_tiles_ub * kTile, (int)n_tiles_ub, H, V, Vpad, bf16 != 0, meta,
                            bias2, scal, row_label, blank, (const float4*)rowmeta, d_act, d_w_out, d_b_out, splits,
                            workspace, workspace ? (size_t)workspace_bytes : 0, (cudaStream_t)stream);
}

int ttx_reduce_act_grad(const float* d_act, const float* eproj, const float* pproj, const int32_t* act_lens,
                        const int32_t* label_lens, const int32_t* meta, int B, int T, int U1, int H,
                        float* d_eproj, float* d_pproj, int device, void* stream) {
    TTX_REQUIRE(d_act && eproj && pproj && act_lens && label_lens && meta && d_eproj && d_pproj,
                "ttx_reduce_act_grad: null pointer");
    TTX_REQUIRE(H % 4 == 0 && B <= 65535 && T > 0 && U1 > 0, "ttx_reduce_act_grad: bad shape");
    TTX_ENTER(device);
    return launch_reduce(d_act, nullptr, nullptr, nullptr, nullptr, 0, eproj, pproj, act_lens, label_lens, meta, B, T,
                         U1, H, d_eproj, d_pproj, (cudaStream_t)stream);
}

int ttx_fwd_grad_supported_h(int H) { return fwd_grad_supported_h(H) ? 1 : 0; }

int ttx_joint_fwd_grad(const void* a16, const void* w16, const void* w16t, const float* bias2, const float* scal,
                       const int32_t* row_label, const int32_t* meta, int64_t n_tiles_ub, int H, int V, int blank,
                       int bf16, float* lse, float* lp_blank, float* lp_label, float* ew, void* workspace,
                       int64_t workspace_bytes, int device, void* stream) {
    TTX_REQUIRE(a16 && w16 && w16t && bias2 && scal && row_label && meta && lse && lp_blank && lp_label && ew,
                "ttx_joint_fwd_grad: null pointer");
    TTX_REQUIRE(fwd_grad_supported_h(H), "ttx_joint_fwd_grad: joint width H=%d is not supported (128, 256, 512)", H);
    TTX_REQUIRE(V > 0 && blank >= 0 && blank < V, "ttx_joint_fwd_grad: bad V=%d / blank=%d", V, blank);
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_joint_fwd_grad(a16, w16, w16t, (uint64_t)n_tiles_ub * kTile, (int)n_tiles_ub, H, V, Vpad, bf16 != 0,
                                 meta, bias2, scal, row_label, blank, lse, lp_blank, lp_label, ew, workspace,
                                 workspace ? (size_t)workspace_bytes : 0, (cudaStream_t)stream);
}

int64_t ttx_joint_workspace_bytes(int which, int64_t n_tiles_ub, int H, int V, int device) {
    if (which < 0 || which > 1 || n_tiles_ub < 1 || n_tiles_ub >= (1 << 24) || V <= 0 || enter(device)) return 0;
    return (int64_t)joint_workspace_bytes(which, (int)n_tiles_ub, H, V);
}

int ttx_wide_supported_h(int H) { return wide_supported_h(H) ? 1 : 0; }

static int wide_range_ok(const char* who, int64_t n_tiles_ub, int tile_lo, int tile_cnt, int64_t store_rows) {
    if (tile_lo < 0 || (tile_lo & 1) || tile_cnt < 1 || tile_lo + (int64_t)tile_cnt > n_tiles_ub + 1 ||
        store_rows < (int64_t)kTile * ((tile_cnt + 1) & ~1) || store_rows % kTile != 0) {
        set_error("%s: bad tile range [%d, +%d) of %lld tiles / P' matrix of %lld rows (tile_lo must be even, the matrix "
                  "must hold the range rounded up to a tile pair)", who, tile_lo, tile_cnt, (long long)n_tiles_ub,
                  (long long)store_rows);
        return 1;
    }
    return 0;
}

int ttx_wide_sp(const void* a16, const void* w16, const float* bias2, const float* scal, const int32_t* row_label,
                const int32_t* meta, int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int blank, int bf16,
                float* lse, float* lp_blank, float* lp_label, float* pfac, float* mref, void* pstore, int64_t store_rows,
                int32_t* flags, int device, void* stream) {
    TTX_REQUIRE(a16 && w16 && bias2 && scal && row_label && meta && lse && lp_blank && lp_label && pfac && mref && pstore &&
                    flags, "ttx_wide_sp: null pointer");
    TTX_REQUIRE(wide_supported_h(H), "ttx_wide_sp: joint width H=%d is not supported (multiples of 512 up to 4096)", H);
    TTX_REQUIRE(V > 0 && blank >= 0 && blank < V, "ttx_wide_sp: bad V=%d / blank=%d", V, blank);
    if (int rc = wide_range_ok("ttx_wide_sp", n_tiles_ub, tile_lo, tile_cnt, store_rows)) return rc;
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_wide_sp(a16, w16, (uint64_t)n_tiles_ub * kTile, tile_lo, tile_cnt, H, V, Vpad, bf16 != 0, meta, bias2, scal,
                          row_label, blank, lse, lp_blank, lp_label, pfac, mref, pstore, (uint64_t)store_rows, flags,
                          (cudaStream_t)stream);
}

int ttx_wide_pw(const void* pstore, int64_t store_rows, const void* w16t, const float* pfac, const float* scal,
                const int32_t* meta, int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int bf16, float* ew,
                int device, void* stream) {
    TTX_REQUIRE(pstore && w16t && pfac && scal && meta && ew, "ttx_wide_pw: null pointer");
    TTX_REQUIRE(wide_supported_h(H), "ttx_wide_pw: joint width H=%d is not supported (multiples of 512 up to 4096)", H);
    TTX_REQUIRE(V > 0, "ttx_wide_pw: bad V=%d", V);
    if (int rc = wide_range_ok("ttx_wide_pw", n_tiles_ub, tile_lo, tile_cnt, store_rows)) return rc;
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_wide_pw(pstore, (uint64_t)store_rows, w16t, tile_lo, tile_cnt, H, V, Vpad, bf16 != 0, meta, scal, pfac, ew,
                          (cudaStream_t)stream);
}

int ttx_wide_dw(const void* pstore, int64_t store_rows, const void* a16st, const float* scal, const int32_t* meta,
                int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int bf16, float* d_w_out, float* d_b_out,
                int device, void* stream) {
    TTX_REQUIRE(pstore && a16st && scal && meta && d_w_out && d_b_out, "ttx_wide_dw: null pointer");
    TTX_REQUIRE(wide_supported_h(H), "ttx_wide_dw: joint width H=%d is not supported (multiples of 512 up to 4096)", H);
    TTX_REQUIRE(V > 0, "ttx_wide_dw: bad V=%d", V);
    if (int rc = wide_range_ok("ttx_wide_dw", n_tiles_ub, tile_lo, tile_cnt, store_rows)) return rc;
    TTX_ENTER(device);
    const int Vpad = ((V + 2 * kTile - 1) / (2 * kTile)) * (2 * kTile);
    return launch_wide_dw(pstore, (uint64_t)store_rows, a16st, (uint64_t)n_tiles_ub * kTile, tile_lo, tile_cnt, H, V, Vpad,
                          bf16 != 0, meta, scal, d_w_out, d_b_out, (cudaStream_t)stream);
}

int ttx_kept_prepare(const void* a16, const void* rowmeta, const int32_t* row_label,
                     const float* lp_blank, const float* lp_label, const float* pfac, const float* scal,
                     const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                     int64_t n_tiles_ub, int H, int blank, int bf16, void* a16st, float* d_w_out, float* d_b_out,
                     int parts, int device, void* stream) {
    TTX_REQUIRE(a16 && rowmeta && row_label && lp_blank && lp_label && pfac && scal && act_lens && label_lens &&
                    meta && a16st && d_w_out && d_b_out, "ttx_kept_prepare: null pointer");
    TTX_REQUIRE(H > 0 && H % 64 == 0 && B > 0 && B <= 65535 && T > 0 && U1 > 0, "ttx_kept_prepare: bad shape");
    TTX_REQUIRE(parts >= 1 && parts <= 3, "ttx_kept_prepare: parts = %d (1 operand copy, 2 blank / label terms, 3 both)", parts);
    TTX_ENTER(device);
    return launch_kept_prepare(a16, (const float4*)rowmeta, row_label, lp_blank, lp_label, pfac, scal, act_lens,
                               label_lens, meta, B, T, U1, H, blank, bf16 != 0, (size_t)n_tiles_ub * kTile, a16st, d_w_out,
                               d_b_out, parts, (cudaStream_t)stream);
}

int ttx_reduce_act_grad_ew(const float* ew, const void* rowmeta, const int32_t* row_label, const float* w_out,
                           const float* scal, int blank, const float* eproj, const float* pproj,
                           const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                           int H, float* d_eproj, float* d_pproj, int device, void* stream) {
    TTX_REQUIRE(ew && rowmeta && row_label && w_out && scal && eproj && pproj && act_lens && label_lens && meta &&
                    d_eproj && d_pproj, "ttx_reduce_act_grad_ew: null pointer");
    TTX_REQUIRE(H % 4 == 0 && B <= 65535 && T > 0 && U1 > 0, "ttx_reduce_act_grad_ew: bad shape");
    TTX_ENTER(device);
    return launch_reduce(ew, (const float4*)rowmeta, row_label, w_out, scal, blank, eproj, pproj, act_lens, label_lens,
                         meta, B, T, U1, H, d_eproj, d_pproj, (cudaStream_t)stream);
}

static int proj_args_ok(const char* who, const void* a, const void* b, const void* c, int M, int N, int K, int lda, int ldb,
                        int ldc, int mina, int minb, int minc) {
    if (!a || !b || !c) {
        set_error("%s: null pointer", who);
        return 1;
    }
    if (M <= 0 || N <= 0 || K <= 0 || (N & 3) || (K & 3) || lda < mina || ldb < minb || ldc < minc || ((lda | ldb | ldc) & 3) ||
        (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15)) {
        set_error("%s: M=%d N=%d K=%d with leading dimensions %d, %d, %d -- N, K and the leading dimensions must be multiples "
                  "of 4 floats, the pointers 16-byte aligned", who, M, N, K, lda, ldb, ldc);
        return 1;
    }
    return 0;
}

int ttx_proj_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, int M, int N, int K, float* y, int ldy,
                 int device, void* stream) {
    if (int rc = proj_args_ok("ttx_proj_fwd", x, w, y, M, N, K, ldx, ldw, ldy, K, K, N)) return rc;
    TTX_REQUIRE(!bias || ((uintptr_t)bias & 15) == 0, "ttx_proj_fwd: bias must be 16-byte aligned");
    TTX_ENTER(device);
    return launch_proj_fwd(x, ldx, w, ldw, bias, M, N, K, y, ldy, (cudaStream_t)stream);
}

int ttx_proj_bwd_x(const float* dy, int lddy, const float* w, int ldw, int M, int N, int K, float* dx, int lddx, int device,
                   void* stream) {
    if (int rc = proj_args_ok("ttx_proj_bwd_x", dy, w, dx, M, N, K, lddy, ldw, lddx, N, K, K)) return rc;
    TTX_ENTER(device);
    return launch_proj_bwd_x(dy, lddy, w, ldw, M, N, K, dx, lddx, (cudaStream_t)stream);
}

int ttx_proj_bwd_w(const float* dy, int lddy, const float* x, int ldx, int M, int N, int K, float* dw, int lddw, float* db,
                   int device, void* stream) {
    if (int rc = proj_args_ok("ttx_proj_bwd_w", dy, x, dw, M, N, K, lddy, ldx, lddw, N, K, K)) return rc;
    TTX_REQUIRE(!db || ((uintptr_t)db & 15) == 0, "ttx_proj_bwd_w: db must be 16-byte aligned");
    TTX_ENTER(device);
    return launch_proj_bwd_w(dy, lddy, x, ldx, M, N, K, dw, lddw, db, (cudaStream_t)stream);
}

int ttx_decode_scan(const float* eproj, int ld_e, const float* pvec, const float* w_out, const float* b_out, int n, int H, int V,
                    int blank, void* scratch, int32_t* out, int device, void* stream) {
    TTX_REQUIRE(eproj && pvec && w_out && b_out && scratch && out, "ttx_decode_scan: null pointer");
    TTX_REQUIRE(n >= 1 && n <= 64 && H > 0 && V > 0 && ld_e >= H && blank >= 0 && blank < V,
                "ttx_decode_scan: bad shape n=%d (1..64) H=%d V=%d ld=%d blank=%d", n, H, V, ld_e, blank);
    TTX_ENTER(device);
    return launch_decode_scan(eproj, ld_e, pvec, w_out, b_out, n, H, V, blank, (unsigned long long*)scratch, out,
                              (cudaStream_t)stream);
}

int ttx_spec_mask(float* x, int B, int T, int F, int64_t stride_b, int64_t stride_t, const int32_t* masks_host, int n_masks,
                  int device, void* stream) {
    TTX_REQUIRE(x && (masks_host || n_masks == 0), "ttx_spec_mask: null pointer");
    TTX_REQUIRE(B > 0 && B <= 65535 && T > 0 && F > 0 && n_masks >= 0 && n_masks <= 64,
                "ttx_spec_mask: bad shape B=%d T=%d F=%d masks=%d (at most 64)", B, T, F, n_masks);
    for (int i = 0; i < n_masks; ++i) {
        const int axis = masks_host[3 * i], start = masks_host[3 * i + 1], width = masks_host[3 * i + 2];
        TTX_REQUIRE((axis == 1 || axis == 2) && start >= 0 && width >= 0 && start + width <= (axis == 1 ? T : F),
                    "ttx_spec_mask: mask %d = {axis %d, start %d, width %d} outside the (%d, %d) frame", i, axis, start,
                    width, T, F);
    }
    if (n_masks == 0) return 0;
    TTX_ENTER(device);
    return launch_spec_mask(x, B, T, F, stride_b, stride_t, masks_host, n_masks, (cudaStream_t)stream);
}

static int band_attn_shape_ok(const char* who, int T, int B, int n_head, int d_head, int max_len, int left, int right,
                              int mode) {
    if (mode != 0 && mode != 1) {
        set_error("%s: mode %d (0 = tt, 1 = espnet)", who, mode);
        return 1;
    }
    if (mode == 1 && max_len != 2 * T - 1) {
        set_error("%s: mode 1 takes a position table of 2T - 1 = %d rows, got %d", who, 2 * T - 1, max_len);
        return 1;
    }
    if (T < 1 || B < 1 || n_head < 1 || d_head < 32 || d_head > 128 || d_head % 32 || max_len < 1 || left < 0 || right < 0 ||
        left + right + 1 > 32 || (long long)T * B * n_head > (1ll << 30)) {
        set_error("%s: bad shape T=%d B=%d heads=%d d_head=%d (multiple of 32 up to 128) max_len=%d context=(%d, %d) (at most "
                  "32 keys per query)", who, T, B, n_head, d_head, max_len, left, right);
        return 1;
    }
    return 0;
}

int ttx_band_attn_fwd(const float* w_heads, const float* r_emb, const float* r_w_bias, const float* r_bias, int T, int B,
                      int n_head, int d_head, int max_len, int left, int right, float scale, int mode, const int32_t* key_lens,
                      float* prob, float* out, int device, void* stream) {
    TTX_REQUIRE(w_heads && r_emb && r_w_bias && r_bias && prob && out, "ttx_band_attn_fwd: null pointer");
    if (int rc = band_attn_shape_ok("ttx_band_attn_fwd", T, B, n_head, d_head, max_len, left, right, mode)) return rc;
    TTX_ENTER(device);
    return launch_band_attn_fwd(w_heads, r_emb, r_w_bias, r_bias, T, B, n_head, d_head, max_len, left, right, scale, mode,
                                key_lens, prob, out, (cudaStream_t)stream);
}

int ttx_band_attn_bwd(const float* w_heads, const float* r_emb, const float* r_w_bias, const float* prob, const float* d_out,
                      int T, int B, int n_head, int d_head, int max_len, int left, int right, float scale, int mode,
                      const int32_t* key_lens, float* ds, float* dq_content, float* d_w_heads, float* d_r_emb,
                      float* d_r_w_bias, float* d_r_bias, int device, void* stream) {
    TTX_REQUIRE(w_heads && r_emb && r_w_bias && prob && d_out && ds && dq_content && d_w_heads && d_r_emb && d_r_w_bias &&
                    d_r_bias, "ttx_band_attn_bwd: null pointer");
    if (int rc = band_attn_shape_ok("ttx_band_attn_bwd", T, B, n_head, d_head, max_len, left, right, mode)) return rc;
    TTX_ENTER(device);
    return launch_band_attn_bwd(w_heads, r_emb, r_w_bias, prob, d_out, T, B, n_head, d_head, max_len, left, right, scale, mode,
                                key_lens, ds, dq_content, d_w_heads, d_r_emb, d_r_w_bias, d_r_bias, (cudaStream_t)stream);
}

int ttx_check_inputs(const int32_t* labels, int label_stride, const int32_t* act_lens, const int32_t* label_lens, int B, int V,
                     int64_t* out, int device, void* stream) {
    TTX_REQUIRE(act_lens && label_lens && out && (labels || label_stride == 0), "ttx_check_inputs: null pointer");
    TTX_REQUIRE(B > 0 && V > 0 && label_stride >= 0, "ttx_check_inputs: bad shape B=%d V=%d", B, V);
    TTX_ENTER(device);
    return launch_check_inputs(labels, label_stride, act_lens, label_lens, B, V, (long long*)out, (cudaStream_t)stream);
}

int ttx_dense_lse(const float* acts, const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens,
                  const int32_t* meta, int B, int T, int U1, int V, int label_stride, int blank,
                  int64_t n_tiles_ub, float* lse, float* lp_blank, float* lp_label, int32_t* row_label, int device,
                  void* stream) {
    TTX_REQUIRE(acts && act_lens && label_lens && meta && lse && lp_blank && lp_label && row_label,
                "ttx_dense_lse: null pointer");
    TTX_REQUIRE(labels || U1 == 1, "ttx_dense_lse: labels is null");
    TTX_REQUIRE(V > 0 && blank >= 0 && blank < V, "ttx_dense_lse: bad V=%d / blank=%d", V, blank);
    TTX_ENTER(device);
    return launch_dense_lse(acts, labels, act_lens, label_lens, meta, B, T, U1, V, label_stride, blank,
                            (int)n_tiles_ub, lse, lp_blank, lp_label, row_label, (cudaStream_t)stream);
}

int ttx_dense_grad(const float* acts, const void* rowmeta, const int32_t* row_label, const float* scal,
                   const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                   int V, int blank, float* grads, int device, void* stream) {
    TTX_REQUIRE(acts && rowmeta && row_label && scal && act_lens && label_lens && meta && grads,
                "ttx_dense_grad: null pointer");
    TTX_ENTER(device);
    return launch_dense_grad(acts, (const float4*)rowmeta, row_label, scal, act_lens, label_lens, meta, B, T, U1, V,
                             blank, grads, (cudaStream_t)stream);
}

}  // extern "C"
