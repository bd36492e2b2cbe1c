/*
 * ttx.h -- C ABI of the B200 joint + transducer-loss path (libttx.so).
 *
 * Plain pointers and sizes only; every pointer is DEVICE memory owned by the caller unless stated
 * otherwise, `stream` is a cudaStream_t passed as void*, all work is stream-ordered with no hidden
 * synchronisation and no global state, so the library is re-entrant per stream.  Every function returns
 * 0 on success, 1 for an argument / shape error, 2 for a CUDA error; ttx_last_error() gives the message of
 * the calling thread's last failure.
 *
 * What each entry point replaces in the reference (zzpDapeng/Transformer-Transducer):
 *   ttx_joint_act + ttx_joint_lse_fwd   JointNet.forward tanh + project_layer, /root/reference/tt/model.py:33-37,
 *                                       JointNetwork.forward, espnet/nets/pytorch_backend/transducer/joint_network.py:48-49,
 *                                       and the log_softmax of warprnnt_pytorch.RNNTLoss (called train.py:53)
 *   ttx_lattice_fwd_bwd                 the alpha/beta recursion of warprnnt_pytorch (train.py:53; transducer/loss.py:74)
 *   ttx_grad_coeffs + ttx_joint_grad + ttx_reduce_act_grad
 *                                       loss.backward() through RNNTLoss, log_softmax, project_layer, tanh and the
 *                                       broadcast add (train.py:58)
 *   ttx_dense_lse / ttx_dense_grad      RNNTLoss on an already materialised (B,T,U+1,V) logits tensor (train.py:53)
 *
 * Lattice cells are addressed in a compact row space described by the tile table ("meta"): utterance b owns
 * 128-row tiles [meta[4+b], meta[4+b+1]); row r of the utterance is cell (t,u) = (r / (label_len[b]+1),
 * r % (label_len[b]+1)).  All per-row arrays below have ttx_tiles_upper_bound(...) * 128 entries.
 */
#ifndef TTX_H_
#define TTX_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int ttx_version(void);
const char* ttx_last_error(void);

/* 1 when the fused tensor-core kernels support joint width H (multiple of 64 up to 256, 384 or 512). */
int ttx_supported_h(int H);
/* Upper bound of 128-row lattice tiles for a (B, T, U1 = U + 1) batch; sizes every per-row buffer. */
int64_t ttx_tiles_upper_bound(int B, int T, int U1);
/* Number of int32 entries the tile table needs. */
int64_t ttx_meta_ints(int B, int64_t n_tiles_ub);

/* Builds the tile table from the (device) length vectors.  meta[0] = tiles in use, meta[1] = 0 or
 * 1 + index of the first utterance whose lengths are out of range (then no kernel does any work). */
int ttx_prepare(const int32_t* act_lens, const int32_t* label_lens, int B, int T, int U1, int64_t n_tiles_ub,
                int32_t* meta, int device, void* stream);

/* W_out (V,H) fp32 -> 16-bit tensor-core operand (Vpad = 256 * ceil(V/256) rows, zero padded), and
 * bias2 (Vpad) = b_out * log2(e), -inf on the padding rows.
 * bf16 = 0: fp16 scaled by a power of two w_scale; bf16 = 1: bfloat16.  scal: 8 floats
 * {w_scale, 1/w_scale, gmax, any-negative-grad flag (both set by ttx_grad_coeffs), scratch...}. */
int ttx_cast_weight(const float* w_out, const float* b_out, int V, int H, int bf16, float* scal, void* w16,
                    float* bias2, void* w16t /* NULL or (H, Vpad): transposed copy for the gradient pass */,
                    int device, void* stream);

/* A16[row,:] = 16-bit(tanh(eproj[b,t,:] + pproj[b,u,:])) for every lattice cell; row_label[row] = label
 * emitted from the cell's u (labels[b,u]) or -1.  eproj (B,T,H), pproj (B,U1,H) fp32 contiguous;
 * labels (B, label_stride) int32, entries at u >= label_lens[b] are never read.  V = vocabulary size: a label
 * >= V is stored as -1 (row_label is what the later calls gather W_out rows and scatter gradient rows through, so
 * they stay inside their buffers whatever the labels hold; the Python front end raises ValueError for such input,
 * as for negative labels inside label_lens); 0 = labels are trusted. */
int ttx_joint_act(const float* eproj, const float* pproj, const int32_t* labels, const int32_t* act_lens,
                  const int32_t* label_lens, const int32_t* meta, int B, int T, int U1, int H, int label_stride,
                  int V, int64_t n_tiles_ub, int bf16, void* a16, int32_t* row_label,
                  void* a16t /* NULL or the transposed copy for the gradient pass, H x rows values stored in blocks of 64 rows: [rows / 64][H][64] */, int device, void* stream);

/* tcgen05 projection A16 . W16^T + b_out fused with log-softmax statistics: per row lse, log p(blank),
 * log p(label).  The (B,T,U1,V) logits are never written. */
int ttx_joint_lse_fwd(const void* a16, const void* w16, const float* bias2, const float* scal,
                      const int32_t* row_label, const int32_t* meta, int64_t n_tiles_ub, int H, int V, int blank,
                      int bf16, float* lse, float* lp_blank, float* lp_label, int device, void* stream);

/* Elements of the diagonal-major lattice arrays (upper bound for a (B, T, U1) batch; the exact figure is
 * sum_b (T_b + U_b) * pitch(U_b + 1), pitch(n) = n rounded up to a multiple of 4). */
int64_t ttx_lattice_elems_upper_bound(int B, int T, int U1);

/* Anti-diagonal wavefront alpha and beta over every utterance's (T_b, U_b+1) lattice, carried as float + float (48 bits)
 * and stored as float64.  One warp per
 * (utterance, direction), warp-shuffle hand-off between columns, operand diagonals staged in shared memory by
 * 16-byte asynchronous copies.  alpha / beta (lat_elems doubles each) are DIAGONAL-MAJOR: cell (t, u) of utterance b
 * is element meta[4 + B+1 + n_tiles_ub + b] + (t + u) * pitch(U_b + 1) + u of alpha; beta lives on the mirrored
 * lattice, beta(t, u) at the position of cell (T_b-1-t, U_b-u).  lat_ws: 4 * lat_elems floats of scratch (the two
 * log-probs per cell re-ordered the same way, as arriving arcs of the lattice and of its mirror image).  lat_elems must cover the batch (ttx_prepare computes the
 * offsets; use the upper bound or the exact sum).
 * costs[b] = -(alpha(T_b-1,U_b) + lp_blank(T_b-1,U_b)); ll_beta[b] = beta(0,0). */
int ttx_lattice_fwd_bwd(const float* lp_blank, const float* lp_label, const int32_t* act_lens,
                        const int32_t* label_lens, const int32_t* meta, int B, int U1, int64_t n_tiles_ub,
                        int64_t lat_elems, float* lat_ws, double* alpha, double* beta, float* costs, double* ll_beta,
                        int device, void* stream);

/* (alpha / beta: the diagonal-major arrays of ttx_lattice_fwd_bwd.)
 * Per-row gradient coefficients rowmeta[row] = {lse, p_blank - rb, p_label - rl, gamma * grad_costs[b] / gmax}
 * (float4; rb / rl = posteriors of the blank / label arc out of the cell), gmax = max_b |grad_costs[b]| -> scal[2].
 * If d_b_out != NULL the sparse (blank / label) part of dL/db_out is accumulated into it. */
int ttx_grad_coeffs(const float* lse, const float* lp_blank, const float* lp_label, const double* alpha,
                    const double* beta, const double* ll_beta, const float* grad_costs, float* scal,
                    const int32_t* row_label, const int32_t* act_lens, const int32_t* label_lens,
                    const int32_t* meta, int B, int blank, int64_t n_tiles_ub, void* rowmeta, float* d_b_out,
                    int device, void* stream);

/* Fused gradient: recomputes the joint tile on the tensor cores and accumulates
 *   d_act (rows,H) fp32 = dL/dA (before the tanh derivative)        if d_act  != NULL
 *   d_w_out (V,H), d_b_out (V) fp32 += dL/dW_out, dense part of dL/db_out   if d_w_out != NULL (caller zero-fills
 *                                      before ttx_grad_coeffs, which adds the sparse part of dL/db_out)
 * `splits` = lattice-row splits of the weight-gradient grid (>= 1).  a16t / w16t (transposed operand copies from
 * ttx_transpose16) may be NULL: the kernels then read the gradient pass's B operand MN-major from a16 / w16. */
/* Wide-joint (H outside the fused kernels) chunked path.  z (rows, Vpad) fp32 = A16[chunk] . W16^T from a library GEMM
 * (un-biased, scaled by w_scale).  ttx_rows_lse: per-row lse / log p(blank) / log p(label).  ttx_rows_grad: q (rows, Vpad)
 * 16-bit = scale * w * (softmax - rb [blank] - rl [label]), the operand of both gradient GEMMs.  Replaces the same
 * reference lines as ttx_joint_lse_fwd / ttx_joint_grad for those widths. */
int ttx_rows_lse(const float* z, int rows, int Vpad, int V, const float* bias2, const float* scal,
                 const int32_t* row_label, int blank, float* lse, float* lp_blank, float* lp_label, int device,
                 void* stream);
int ttx_rows_grad(const float* z, const void* rowmeta, const int32_t* row_label, const float* bias2,
                  const float* scal, int rows, int Vpad, int V, int blank, int bf16, void* q, int device, void* stream);

/* out (cols, rows) = transpose of the row-major 16-bit matrix in (rows, cols); rows, cols multiples of 64.
 * meta != NULL: `in` is the A16 operand and row blocks beyond the tiles in use (meta[0]) are skipped.  Produces the
 * K-major operand copies W16^T (H, Vpad) and A16^T (with meta: blocks of 64 lattice rows, [rows / 64][H][64], the layout
 * ttx_joint_act writes) streamed by the backward pair kernel. */
int ttx_transpose16(const void* in, void* out, int rows, int cols, const int32_t* meta, int device, void* stream);

int ttx_joint_grad(const void* a16, const void* w16, const void* a16t, const void* w16t, const float* bias2,
                   const float* scal,
                   const int32_t* row_label, const int32_t* meta, const void* rowmeta, int64_t n_tiles_ub, int H,
                   int V, int blank, int bf16, float* d_act, float* d_w_out, float* d_b_out, int splits,
                   void* workspace /* NULL or ttx_joint_workspace_bytes(1, ...) bytes, see ttx_joint_fwd_grad */,
                   int64_t workspace_bytes, int device, void* stream);

/* d_eproj[b,t,:] = sum_u d_act * (1 - tanh^2), d_pproj[b,u,:] = sum_t d_act * (1 - tanh^2); both fully
 * overwritten (zeros outside the ragged region). */
int ttx_reduce_act_grad(const float* d_act, const float* eproj, const float* pproj, const int32_t* act_lens,
                        const int32_t* label_lens, const int32_t* meta, int B, int T, int U1, int H,
                        float* d_eproj, float* d_pproj, int device, void* stream);

/* Forward with the activation-gradient contraction fused in (flash-attention style): same outputs as
 * ttx_joint_lse_fwd plus ew (rows, H) fp32 = sum_v softmax_v * W_out[v, :] over all v except the blank and the
 * cell's label.  With it the backward needs no activation-gradient tensor-core pass: ttx_reduce_act_grad_ew forms
 * dL/dA = gmax * w * (ew + (p_b - rb) W_out[blank] + (p_l - rl) W_out[label]) from the lattice coefficients
 * (rowmeta of ttx_grad_coeffs) on the fly.  H in {128, 256, 512} (ttx_fwd_grad_supported_h); needs w16t. */
int ttx_fwd_grad_supported_h(int H);
/* workspace: NULL, or ttx_joint_workspace_bytes(0, ...) bytes of caller-owned scratch (H = 512: the launch stores the
 * 16-bit softmax numerators of the tile pair a CTA pair is working on there and replays them for the second half of the
 * joint columns instead of recomputing the projection; bounded by the device's CTA count, independent of B, T, U). */
int ttx_joint_fwd_grad(const void* a16, const void* w16, const void* w16t, const float* bias2, const float* scal,
                       const int32_t* row_label, const int32_t* meta, int64_t n_tiles_ub, int H, int V, int blank,
                       int bf16, float* lse, float* lp_blank, float* lp_label, float* ew, void* workspace,
                       int64_t workspace_bytes, int device, void* stream);
/* Scratch the fused launches can use: which = 0 ttx_joint_fwd_grad, 1 = ttx_joint_grad with d_w_out.  0 = none. */
int64_t ttx_joint_workspace_bytes(int which, int64_t n_tiles_ub, int H, int V, int device);
int ttx_reduce_act_grad_ew(const float* ew, const void* rowmeta, const int32_t* row_label, const float* w_out,
                           const float* scal, int blank, const float* eproj, const float* pproj,
                           const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                           int H, float* d_eproj, float* d_pproj, int device, void* stream);

/* ---- Wide joints (H a multiple of 512 up to 4096: aishell.yaml's 1024, joint_streaming.yaml's 2048; tt/model.py:35-37).
 * The same three contractions as three streamed tcgen05 products around the 16-bit softmax numerators P' (store_rows x
 * Vpad values, Vpad = V rounded up to 256, in 64-column blocks [Vpad / 64][store_rows][64] so that every tile the kernels
 * move is contiguous; blank / label entries zero).  Every call works on lattice tiles [tile_lo, tile_lo + tile_cnt) (tile_lo
 * even; clipped on the device to the tiles in use) and addresses the P' matrix relative to tile_lo: a caller that can
 * hold P' for the whole batch passes (0, n_tiles_ub) and keeps it for the backward; otherwise it walks the batch in
 * chunks with one chunk-sized matrix and calls ttx_wide_sp again in the backward.
 *   ttx_wide_sp   S = A16 . W16^T with both operands streamed; lse / log p(blank) / log p(label) per row, P', pfac
 *                 (softmax = P' * pfac) and mref (the row's reference, scratch).  flags: one int32 per tile pair of the
 *                 batch, zeroed by the caller; rows whose reference had to move are recomputed by a second launch, so
 *                 P' is always consistent on return.
 *   ttx_wide_pw   ew (rows, H) fp32 = P' . W16 * pfac / w_scale  (as ttx_joint_fwd_grad's ew)
 *   ttx_wide_dw   d_w_out += P'^T . As, d_b_out += dense part, As = a16st from ttx_kept_prepare.
 *   ttx_kept_prepare  a16st (caller scratch of (H + 16) x rows_ub 16-bit values + 64 x (H + 4) floats) = As, the copy of
 *                 A16 with every lattice row scaled by its gradient weight (rows_ub x H, row-major like A16 -- the
 *                 product reads it MN-major, there is no transposed copy), then the rows_ub scales themselves, then
 *                 (at byte (H + 16) * rows_ub * 2) the partial rows of the blank term; and the exact blank / label terms
 *                 of d_w_out / d_b_out. */
int ttx_wide_supported_h(int H);
int ttx_wide_sp(const void* a16, const void* w16, const float* bias2, const float* scal, const int32_t* row_label,
                const int32_t* meta, int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int blank, int bf16,
                float* lse, float* lp_blank, float* lp_label, float* pfac, float* mref, void* pstore, int64_t store_rows,
                int32_t* flags, int device, void* stream);
int ttx_wide_pw(const void* pstore, int64_t store_rows, const void* w16t, const float* pfac, const float* scal,
                const int32_t* meta, int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int bf16, float* ew,
                int device, void* stream);
int ttx_wide_dw(const void* pstore, int64_t store_rows, const void* a16st, const float* scal, const int32_t* meta,
                int64_t n_tiles_ub, int tile_lo, int tile_cnt, int H, int V, int bf16, float* d_w_out, float* d_b_out,
                int device, void* stream);
int ttx_kept_prepare(const void* a16, const void* rowmeta, const int32_t* row_label,
                     const float* lp_blank, const float* lp_label, const float* pfac, const float* scal,
                     const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                     int64_t n_tiles_ub, int H, int blank, int bf16, void* a16st, float* d_w_out, float* d_b_out,
                     int parts /* 1: the operand copy, 2: the blank / label terms, 3: both -- the two are independent and may
                                  run on different streams (the terms use the 64 x (H + 4) floats at the end of a16st) */,
                     int device, void* stream);

/* ---- Pre-projections of the joint (tt/model.py:35 forward_layer, split into its encoder / decoder halves;
 * joint_network.py:28-31,48 lin_enc / lin_dec), fp32 row-major with leading dimensions in floats.  tcgen05 kind::tf32 with
 * on-chip error compensation (x = hi + lo; hi.hi + lo.hi + hi.lo accumulated in fp32): fp32-grade results like the
 * reference's SGEMM.  N, K and the leading dimensions must be multiples of 4, pointers 16-byte aligned.
 *   ttx_proj_fwd     y (M,N)  = x (M,K) . w (N,K)^T + bias (N, or NULL)
 *   ttx_proj_bwd_x   dx (M,K) = dy (M,N) . w (N,K)
 *   ttx_proj_bwd_w   dw (N,K) += dy (M,N)^T . x (M,K);  db (N) += column sums of dy (db may be NULL).  dw and db must be
 *                    zero-filled (or hold a running sum): the contraction is split over CTAs that reduce into them. */
int ttx_proj_fwd(const float* x, int ldx, const float* w, int ldw, const float* bias, int M, int N, int K, float* y, int ldy,
                 int device, void* stream);
int ttx_proj_bwd_x(const float* dy, int lddy, const float* w, int ldw, int M, int N, int K, float* dx, int lddx, int device,
                   void* stream);
int ttx_proj_bwd_w(const float* dy, int lddy, const float* x, int ldx, int M, int N, int K, float* dw, int lddw, float* db,
                   int device, void* stream);

/* ---- Decode-time joint (greedy search, tt/model.py:70-90, tt_espnet/model.py:83-106).  Scores n <= 64 consecutive frames
 * against ONE decoder state: z[f] = tanh(eproj[f] + pvec) . w_out^T + b_out in fp32, per-frame argmax (lowest index on
 * ties).  eproj: rows of the pre-projected encoder states (leading dimension ld_e floats), pvec (H): pre-projected decoder
 * state, w_out (V,H), b_out (V).  scratch: 64 x 8 bytes.  out (2 + n int32): out[0] = first frame whose argmax is not
 * `blank` (n if none), out[1] = that label, out[2 + f] = argmax of frame f.  Replaces the per-frame
 * joint -> softmax -> argmax -> .item() loop by one launch group and one host read per emitted label. */
int ttx_decode_scan(const float* eproj, int ld_e, const float* pvec, const float* w_out, const float* b_out, int n, int H, int V,
                    int blank, void* scratch, int32_t* out, int device, void* stream);

/* ---- Input pipeline: the reference's time / frequency mask augmentation (tt/utils.py:297-329, train.py:41-44: twenty
 * slice assignments over the batch) as one in-place launch.  x (B,T,F) fp32 on the device with element strides stride_b,
 * stride_t (frequency contiguous); masks_host: HOST array of n_masks x {axis (1 = time, 2 = frequency), start, width},
 * at most 64 -- the positions are drawn by the caller with the reference's own random-number calls. */
int ttx_spec_mask(float* x, int B, int T, int F, int64_t stride_b, int64_t stride_t, const int32_t* masks_host, int n_masks,
                  int device, void* stream);

/* ---- Banded (streaming) relative-position attention core of the tt encoder: RelLearnableMultiHeadAttn.forward
 * (tt/transformer.py:121-167) under the context mask of tt/utils.py:242-251 (a query sees `left` frames back and `right`
 * frames ahead).  fp32.  w_heads (T, B, 3 * n_head * d_head) = qkv_net's output [q | k | v]; r_emb (max_len, n_head, d_head),
 * r_w_bias (n_head, d_head), r_bias (max_len, n_head); scale = 1 / sqrt(d_head).  prob (T, B, n_head, left + right + 1):
 * the band's softmax, kept for the backward; out (T, B, n_head * d_head) = attn_vec (transformer.py:165-167).
 * Backward: ds (like prob) and dq_content (T, B, n_head * d_head) are scratch; d_w_heads is fully written; d_r_emb,
 * d_r_w_bias, d_r_bias must be zero-filled (they are accumulated).
 * mode 0 = the tt module above (its _rel_shift, transformer.py:82-95, wraps right of the diagonal; key_lens = NULL).
 * mode 1 = the espnet side: RelPositionMultiHeadedAttention.forward (espnet/nets/pytorch_backend/transformer/
 * attention.py:264-308) under make_attention_mask (nets_utils.py:268-281) and the padding mask (espnet2/asr/encoder/
 * transformer_encoder.py:205-210): r_emb = linear_pos(pos_emb) with max_len = 2T - 1 rows, r_w_bias = pos_bias_u,
 * r_bias[r, h] = pos_bias_v[h] . r_emb[r, h] (folded by the caller), key j of batch entry b is masked when
 * j >= key_lens[b] (device, int32, B entries; NULL = no padding); a query without any key gets zeros. */
int ttx_band_attn_fwd(const float* w_heads, const float* r_emb, const float* r_w_bias, const float* r_bias, int T, int B,
                      int n_head, int d_head, int max_len, int left, int right, float scale, int mode,
                      const int32_t* key_lens, float* prob, float* out, int device, void* stream);
int ttx_band_attn_bwd(const float* w_heads, const float* r_emb, const float* r_w_bias, const float* prob, const float* d_out,
                      int T, int B, int n_head, int d_head, int max_len, int left, int right, float scale, int mode,
                      const int32_t* key_lens, float* ds, float* dq_content, float* d_w_heads, float* d_r_emb,
                      float* d_r_w_bias, float* d_r_bias, int device, void* stream);

/* The device side of warprnnt_pytorch's argument checks (certify_inputs) in one launch: out (7 x int64, device) = max T,
 * max U, min T, min U over the batch, the batch's 128-row lattice tiles, the number of labels outside [0, V) inside their
 * utterance's label_lens, the elements of the diagonal-major lattice arrays.  labels (B, label_stride) int32. */
int ttx_check_inputs(const int32_t* labels, int label_stride, const int32_t* act_lens, const int32_t* label_lens, int B, int V,
                     int64_t* out, int device, void* stream);

/* Dense-logits entry: acts (B,T,U1,V) fp32 contiguous. */
int ttx_dense_lse(const float* acts, const int32_t* labels, const int32_t* act_lens, const int32_t* label_lens,
                  const int32_t* meta, int B, int T, int U1, int V, int label_stride, int blank,
                  int64_t n_tiles_ub, float* lse, float* lp_blank, float* lp_label, int32_t* row_label, int device,
                  void* stream);
int ttx_dense_grad(const float* acts, const void* rowmeta, const int32_t* row_label, const float* scal,
                   const int32_t* act_lens, const int32_t* label_lens, const int32_t* meta, int B, int T, int U1,
                   int V, int blank, float* grads, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TTX_H_ */
