"""Autograd functions that drive the CUDA kernels through the C ABI (include/ttx.h).

``fused_joint_rnnt``  pre-projected encoder / predictor activations + output layer -> per-utterance
                      costs, never materialising the (B,T,U+1,V) logits.  Replaces
                      tanh + project_layer + log_softmax + warprnnt_pytorch.RNNTLoss
                      (/root/reference/tt/model.py:36-37, train.py:53,58).
``dense_rnnt``        the same loss on an already materialised logits tensor (train.py:53 when the
                      joint is not ours).
"""
import contextlib
import os

import torch

from . import _lib


def _p(t):
    """Device address for a c_void_p argument (a plain int: ctypes converts it; None = NULL)."""
    return t.data_ptr() if t is not None else None


_NO_SWITCH = contextlib.nullcontext()


def _guard(dev):
    """`with torch.cuda.device(dev)` only when dev is not already the current device (the usual one-process-per-GPU case
    enters nothing: the context manager costs more than a kernel launch)."""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _NO_SWITCH
    return torch.cuda.device(dev)


def _stream(dev):
    """The current stream's raw handle on dev (an int; 0 = the legacy default stream).  The private accessor skips
    torch.cuda.current_stream()'s Stream object: this runs once per kernel launch."""
    return torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device())


# name -> number of kernels one call launches (bench.py's gpu_launches / per-kernel timing)
KERNELS_PER_CALL = {"ttx_joint_fwd_grad": 1, "ttx_reduce_act_grad_ew": 1, "ttx_rows_lse": 1, "ttx_rows_grad": 1,
                    "ttx_wide_sp": 2, "ttx_wide_pw": 1, "ttx_wide_dw": 1, "ttx_kept_prepare": 3, "ttx_transpose16": 1,
                    "ttx_prepare": 1, "ttx_cast_weight": 2, "ttx_joint_act": 1, "ttx_joint_lse_fwd": 1,
                    "ttx_lattice_fwd_bwd": 2, "ttx_grad_coeffs": 2, "ttx_joint_grad": 2, "ttx_reduce_act_grad": 1,
                    "ttx_dense_lse": 1, "ttx_dense_grad": 1, "ttx_proj_fwd": 1, "ttx_proj_bwd_x": 1, "ttx_proj_bwd_w": 2}
PROFILE = None  # set to a list by bench.py: receives (name, start_event, end_event, n_kernels)


def _call(name, dev, *args, n_kernels=None, label=None):
    fn = getattr(_lib.get(), name)
    prof = PROFILE
    if prof is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(dev))
        rc = fn(*args)
        e1.record(torch.cuda.current_stream(dev))
        prof.append((label or name, e0, e1, KERNELS_PER_CALL[name] if n_kernels is None else n_kernels))
    else:
        rc = fn(*args)
    _lib.check(rc, name)


def _i32_cuda(t, dev, name):
    if t.dtype != torch.int32:
        raise TypeError("%s must be int32" % name)
    return t.to(device=dev).contiguous()


class _Plan:
    """Shape bookkeeping + the buffers shared by both entry paths."""

    def __init__(self, B, T, U1, dev, act_lens, label_lens, sizes=None):
        n_tiles, n_lat = sizes if sizes is not None else (None, None)
        lib = _lib.get()
        self.lib, self.dev, self.B, self.T, self.U1 = lib, dev, B, T, U1
        self.idx = dev.index if dev.index is not None else torch.cuda.current_device()
        # Every per-row buffer is sized by `ntub` tiles.  With the exact tile count from the host-side length check
        # (loss.certify_inputs; ragged batches: ~1/4 of the dense bound at configs[4]) rounded up to a whole tile pair,
        # otherwise by the dense upper bound B * ceil(T * U1 / 128).  ttx_prepare flags a count that is too small.
        ub = int(lib.ttx_tiles_upper_bound(B, T, U1))
        self.ntub = ub if n_tiles is None else max(2, min(ub + 1, (int(n_tiles) + 1) & ~1))
        self.rows = self.ntub * 128
        lat_ub = int(lib.ttx_lattice_elems_upper_bound(B, T, U1))
        self.lat = lat_ub if n_lat is None else max(4, min(lat_ub, int(n_lat)))     # diagonal-major lattice elements
        self.meta = torch.empty(int(lib.ttx_meta_ints(B, self.ntub)), dtype=torch.int32, device=dev)
        self.act_lens, self.label_lens = act_lens, label_lens
        _call("ttx_prepare", dev, _p(act_lens), _p(label_lens), B, T, U1, self.ntub, _p(self.meta), self.idx,
                                   _stream(dev))

    def rowf(self, n=1):
        return torch.empty(self.rows * n, dtype=torch.float32, device=self.dev)

    def lattice(self, lse, lpb, lpl):
        alpha = torch.empty(self.lat, dtype=torch.float64, device=self.dev)
        beta = torch.empty(self.lat, dtype=torch.float64, device=self.dev)
        lat_ws = torch.empty(4 * self.lat, dtype=torch.float32, device=self.dev)
        costs = torch.full((self.B,), float("nan"), dtype=torch.float32, device=self.dev)   # stays NaN on bad lengths
        ll_beta = torch.empty(self.B, dtype=torch.float64, device=self.dev)
        _call("ttx_lattice_fwd_bwd", self.dev, _p(lpb), _p(lpl), _p(self.act_lens), _p(self.label_lens),
                                               _p(self.meta), self.B, self.U1, self.ntub, self.lat, _p(lat_ws),
                                               _p(alpha), _p(beta), _p(costs), _p(ll_beta), self.idx,
                                               _stream(self.dev))
        return alpha, beta, costs, ll_beta

    def grad_coeffs(self, lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label, blank, d_b=None):
        rowmeta = self.rowf(4)
        g = grad_costs.detach().to(torch.float32).contiguous()
        _call("ttx_grad_coeffs", self.dev, _p(lse), _p(lpb), _p(lpl), _p(alpha), _p(beta), _p(ll_beta), _p(g),
              _p(scal), _p(row_label), _p(self.act_lens), _p(self.label_lens), _p(self.meta), self.B, int(blank),
              self.ntub, _p(rowmeta), _p(d_b), self.idx, _stream(self.dev))
        return rowmeta


def _dw_splits(sms, V, H, ntub):
    """Lattice-row splits of the weight-gradient grid: CTA pairs = ceil(V/256) * slabs * splits should fill whole
    waves of sms/2 pair slots (the kernel is one pair per SM pair), with enough stream tiles left per CTA."""
    pairs0 = ((V + 255) // 256) * (2 if H > 256 else 1)
    slots = max(1, sms // 2)
    best, best_eff = 1, 0.0
    for s in range(1, 65):
        if s > 1 and (ntub + 1) // 2 < 16 * s:
            break
        pairs = pairs0 * s
        eff = pairs / (slots * -(-pairs // slots))
        if eff > best_eff + 0.02:
            best, best_eff = s, eff
    return best


def _workspace(lib, which, plan, H, V, dev):
    """Scratch of the fused launches' P' replay (H = 512; bounded by the CTA count, ~165 MB / <= 1.5 GB on a B200)."""
    n = int(lib.ttx_joint_workspace_bytes(which, plan.ntub, H, V, plan.idx))
    return (torch.empty(n, dtype=torch.uint8, device=dev), n) if n > 0 else (None, 0)


def supported_width(H):
    return bool(_lib.get().ttx_supported_h(int(H)))


class FusedJointRNNT(torch.autograd.Function):
    """Joint widths up to 512 without any V-wide tensor in HBM: forward statistics + EW in one tensor-core launch, weight
    gradient by a second launch that recomputes the projection.  The default for H < 512, the memory-bounded alternative
    at 512 (WideJointRNNT keeps the 16-bit softmax numerators instead and is faster when they fit)."""

    SEPARATE_ACT_GRAD = False   # True: plain forward + a separate activation-gradient launch (tests A/B the two routes)
    REPLAY = True               # False: no workspace -- the second half of the joint columns recomputes its projection
    PAIR_WIDTHS = (128, 256, 512)   # widths whose backward runs the CTA-pair kernels (they need transposed operand copies)

    @staticmethod
    def forward(ctx, eproj, pproj, w_out, b_out, labels, act_lens, label_lens, blank, bf16, sizes=None):
        need_grad = any(ctx.needs_input_grad[:4])      # (grad mode is off inside Function.forward)
        if not eproj.is_cuda:
            raise RuntimeError("fused_joint_rnnt needs CUDA tensors (there is no CPU fallback)")
        dev = eproj.device
        B, T, H = eproj.shape
        U1 = pproj.shape[1]
        V = w_out.shape[0]
        if pproj.shape[0] != B or pproj.shape[2] != H or w_out.shape[1] != H or b_out.shape[0] != V:
            raise ValueError("inconsistent joint shapes")
        lib = _lib.get()
        if not lib.ttx_supported_h(H):
            raise ValueError("joint width %d is not supported by the fused tensor-core path" % H)
        ep = eproj.detach().float().contiguous()
        pp = pproj.detach().float().contiguous()
        w = w_out.detach().float().contiguous()
        b = b_out.detach().float().contiguous()
        labels = labels.contiguous()
        with _guard(dev):
            plan = _Plan(B, T, U1, dev, act_lens, label_lens, sizes)
            st = _stream(dev)
            Vpad = (V + 255) // 256 * 256
            scal = torch.zeros(8, dtype=torch.float32, device=dev)
            w16 = torch.empty(Vpad * H, dtype=torch.int16, device=dev)
            bias2 = torch.empty(Vpad, dtype=torch.float32, device=dev)
            # K-major (transposed) copies of both operands feed the gradient pass of the backward pair kernel
            with_t = need_grad and H in FusedJointRNNT.PAIR_WIDTHS
            w16t = torch.empty(H * Vpad, dtype=torch.int16, device=dev) if with_t else None
            a16t = torch.empty(H * plan.rows, dtype=torch.int16, device=dev) if with_t else None
            _call("ttx_cast_weight", dev, _p(w), _p(b), V, H, int(bf16), _p(scal), _p(w16), _p(bias2), _p(w16t),
                  plan.idx, st)
            a16 = torch.empty(plan.rows * H, dtype=torch.int16, device=dev)
            row_label = torch.empty(plan.rows, dtype=torch.int32, device=dev)
            lstride = labels.shape[1] if labels.dim() == 2 else 0
            _call("ttx_joint_act", dev, _p(ep), _p(pp), _p(labels) if labels.numel() else None, _p(act_lens),
                                         _p(label_lens), _p(plan.meta), B, T, U1, H, lstride, V, plan.ntub, int(bf16),
                                         _p(a16), _p(row_label), _p(a16t), plan.idx, st)
            lse, lpb, lpl = plan.rowf(), plan.rowf(), plan.rowf()
            # When the activations need gradients the forward also accumulates EW = sum_v p_v W_out[v] (minus the
            # blank / label columns) on the tensor cores, so the backward has no activation-gradient MMA pass.
            ew = None
            if (with_t and (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]) and lib.ttx_fwd_grad_supported_h(H)
                    and not FusedJointRNNT.SEPARATE_ACT_GRAD):
                ew = plan.rowf(H)
                ws, ws_n = _workspace(lib, 0, plan, H, V, dev) if FusedJointRNNT.REPLAY else (None, 0)
                _call("ttx_joint_fwd_grad", dev, _p(a16), _p(w16), _p(w16t), _p(bias2), _p(scal), _p(row_label),
                      _p(plan.meta), plan.ntub, H, V, int(blank), int(bf16), _p(lse), _p(lpb), _p(lpl), _p(ew), _p(ws),
                      ws_n, plan.idx, st)
            else:
                _call("ttx_joint_lse_fwd", dev, _p(a16), _p(w16), _p(bias2), _p(scal), _p(row_label), _p(plan.meta),
                      plan.ntub, H, V, int(blank), int(bf16), _p(lse), _p(lpb), _p(lpl), plan.idx, st)
            alpha, beta, costs, ll_beta = plan.lattice(lse, lpb, lpl)
        ctx.plan, ctx.blank, ctx.bf16, ctx.dims = plan, int(blank), int(bf16), (B, T, U1, H, V)
        ctx.in_dtypes = (eproj.dtype, pproj.dtype, w_out.dtype, b_out.dtype)
        ctx.transposed = (a16t, w16t)
        ctx.ew, ctx.w32 = ew, (w if ew is not None else None)
        ctx.save_for_backward(ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta)
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta = ctx.saved_tensors
        plan, (B, T, U1, H, V) = ctx.plan, ctx.dims
        lib, dev = plan.lib, plan.dev
        need_act = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        d_ep = d_pp = d_w = d_b = None
        with _guard(dev):
            st = _stream(dev)
            scal = scal.clone()
            if need_w:
                d_w = torch.zeros(V, H, dtype=torch.float32, device=dev)
                d_b = torch.zeros(V, dtype=torch.float32, device=dev)
            rowmeta = plan.grad_coeffs(lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label, ctx.blank, d_b)
            ew = ctx.ew
            d_act = plan.rowf(H) if (need_act and ew is None) else None
            if need_act or need_w:
                a16t, w16t = ctx.transposed
                splits = _dw_splits(torch.cuda.get_device_properties(dev).multi_processor_count, V, H, plan.ntub)
                # two launches (activation gradient, weight gradient) so each shows up separately in profiles
                if need_act and ew is None:
                    _call("ttx_joint_grad", dev, _p(a16), _p(w16), _p(a16t), _p(w16t), _p(bias2), _p(scal), _p(row_label), _p(plan.meta),
                          _p(rowmeta), plan.ntub, H, V, ctx.blank, ctx.bf16, _p(d_act), None, None, 1, None, 0, plan.idx, st,
                          n_kernels=1, label="ttx_joint_grad[dA]")
                if need_w:
                    ws, ws_n = _workspace(lib, 1, plan, H, V, dev) if (FusedJointRNNT.REPLAY and a16t is not None) else (None, 0)
                    _call("ttx_joint_grad", dev, _p(a16), _p(w16), _p(a16t), _p(w16t), _p(bias2), _p(scal), _p(row_label), _p(plan.meta),
                          _p(rowmeta), plan.ntub, H, V, ctx.blank, ctx.bf16, None, _p(d_w), _p(d_b), splits, _p(ws), ws_n,
                          plan.idx, st, n_kernels=1, label="ttx_joint_grad[dW]")
            if need_act:
                d_ep = torch.zeros(B, T, H, dtype=torch.float32, device=dev)   # (stay zero if the lengths were flagged bad)
                d_pp = torch.zeros(B, U1, H, dtype=torch.float32, device=dev)
                if ew is not None:
                    _call("ttx_reduce_act_grad_ew", dev, _p(ew), _p(rowmeta), _p(row_label), _p(ctx.w32), _p(scal),
                          ctx.blank, _p(ep), _p(pp), _p(plan.act_lens), _p(plan.label_lens), _p(plan.meta), B, T, U1, H,
                          _p(d_ep), _p(d_pp), plan.idx, st)
                else:
                    _call("ttx_reduce_act_grad", dev, _p(d_act), _p(ep), _p(pp), _p(plan.act_lens),
                          _p(plan.label_lens), _p(plan.meta), B, T, U1, H, _p(d_ep), _p(d_pp), plan.idx, st)
        dt = ctx.in_dtypes
        cast = lambda g, d, need: g.to(d) if (g is not None and need) else None  # noqa: E731
        return (cast(d_ep, dt[0], ctx.needs_input_grad[0]), cast(d_pp, dt[1], ctx.needs_input_grad[1]),
                cast(d_w, dt[2], ctx.needs_input_grad[2]), cast(d_b, dt[3], ctx.needs_input_grad[3]),
                None, None, None, None, None, None)


def wide_width(H):
    return bool(_lib.get().ttx_wide_supported_h(int(H)))


def _chunk_ranges(ntub, Vpad, budget_bytes):
    """Tile ranges (tile_lo even) whose 16-bit P' matrix (128 * tiles rows x Vpad) fits the budget; at least one tile
    pair per range."""
    tiles = max(2, int(budget_bytes // (128 * Vpad * 2)) & ~1)
    return [(t0, min(tiles, ntub - t0)) for t0 in range(0, ntub, tiles)]


def _wide_pstore(plan, Vpad, dev):
    """The P' matrix and the tile ranges it is used for.  One range = the whole batch: P' is kept from forward to
    backward; several: one chunk-sized matrix, and the backward recomputes each chunk's P'.  Budget: TTX_KEEP_GB (default
    32); if the device cannot give that much next to the rest of the model, halve until it can (an allocation attempt is
    the only cheap probe: cudaMemGetInfo costs milliseconds while kernels are running)."""
    budget = max(float(os.environ.get("TTX_KEEP_GB", "32")), 0.0) * 2**30
    while True:
        chunks = _chunk_ranges(plan.ntub, Vpad, budget)
        store_rows = 128 * ((max(c[1] for c in chunks) + 1) & ~1)
        try:
            return chunks, store_rows, torch.empty(store_rows * Vpad, dtype=torch.int16, device=dev)
        except torch.cuda.OutOfMemoryError:
            if store_rows <= 256:
                raise
            budget = store_rows * Vpad * 2 / 2


_HI_STREAMS = {}


def _hi_stream(dev):
    """A high-priority stream per device for the streamed products: they are bound by the SM's TMA ingest and leave its
    issue slots and load/store path idle (and a quarter of its registers and shared memory free), so the bandwidth-bound
    kernel that is independent of them -- the lattice next to EW = P' . W, the activation-gradient reduction next to
    dW = P'^T . As -- runs on the launching stream at the same time, its blocks filling the space the product's CTAs
    leave.  The priority makes the block scheduler place the product's CTAs first."""
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device())
    s = _HI_STREAMS.get(key)
    if s is None:
        s = _HI_STREAMS[key] = torch.cuda.Stream(dev, priority=-1)
    return s


class WideJointRNNT(torch.autograd.Function):
    """Same contract as FusedJointRNNT for joint widths that are multiples of 512 (aishell.yaml's 1024,
    joint_streaming.yaml's 2048; /root/reference/tt/model.py:35-37): three streamed tcgen05 products around the 16-bit
    softmax numerators P' (csrc/ttx_wide.cu) -- S pass with both operands streamed, EW = P' . W16, dW = P'^T . As --
    with everything else (operand casts, lattice, coefficients, reductions) shared with the fused path.  No library GEMM."""

    OVERLAP = True      # False: everything on the launching stream (A/B runs)

    @staticmethod
    def forward(ctx, eproj, pproj, w_out, b_out, labels, act_lens, label_lens, blank, bf16, sizes=None,
                enc=None, w_enc=None, b_enc=None, dec=None, w_dec=None):
        """enc .. w_dec (optional, float32): the inputs of the two pre-projections eproj = enc w_enc^T + b_enc,
        pproj = dec w_dec^T.  When given, eproj / pproj come in detached and THIS node returns the gradients of
        enc, w_enc, b_enc, dec, w_dec: its backward issues the projection-backward kernels right behind the
        activation-gradient reduction, i.e. while the weight-gradient product is still running on the other stream
        (as separate autograd nodes they could only start after the streams have joined)."""
        pre = enc is not None
        need_act = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (pre and any(ctx.needs_input_grad[10:15]))
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        if not eproj.is_cuda:
            raise RuntimeError("fused_joint_rnnt needs CUDA tensors (there is no CPU fallback)")
        dev = eproj.device
        B, T, H = eproj.shape
        U1 = pproj.shape[1]
        V = w_out.shape[0]
        if pproj.shape[0] != B or pproj.shape[2] != H or w_out.shape[1] != H or b_out.shape[0] != V:
            raise ValueError("inconsistent joint shapes")
        ep, pp = eproj.detach().float().contiguous(), pproj.detach().float().contiguous()
        w, b = w_out.detach().float().contiguous(), b_out.detach().float().contiguous()
        labels = labels.contiguous()
        ctx.pre = None
        if pre:
            ctx.pre = (enc.detach().reshape(-1, enc.shape[-1]).contiguous(), w_enc.detach(), b_enc is not None,
                       dec.detach().reshape(-1, dec.shape[-1]).contiguous(), w_dec.detach(), enc.shape, dec.shape)
        with _guard(dev):
            plan = _Plan(B, T, U1, dev, act_lens, label_lens, sizes)
            st = _stream(dev)
            Vpad = (V + 255) // 256 * 256
            scal = torch.zeros(8, dtype=torch.float32, device=dev)
            w16 = torch.empty(Vpad * H, dtype=torch.int16, device=dev)
            bias2 = torch.empty(Vpad, dtype=torch.float32, device=dev)
            w16t = torch.empty(H * Vpad, dtype=torch.int16, device=dev) if need_act else None
            a16 = torch.empty(plan.rows * H, dtype=torch.int16, device=dev)
            row_label = torch.empty(plan.rows, dtype=torch.int32, device=dev)
            main = torch.cuda.current_stream(dev)
            side = _hi_stream(dev) if WideJointRNNT.OVERLAP else None
            if side is not None:               # the weight cast and the activation kernel do not depend on each other
                side.wait_stream(main)
            with torch.cuda.stream(side if side is not None else main):
                _call("ttx_cast_weight", dev, _p(w), _p(b), V, H, int(bf16), _p(scal), _p(w16), _p(bias2), _p(w16t),
                      plan.idx, _stream(dev))
            lstride = labels.shape[1] if labels.dim() == 2 else 0
            _call("ttx_joint_act", dev, _p(ep), _p(pp), _p(labels) if labels.numel() else None, _p(act_lens),
                  _p(label_lens), _p(plan.meta), B, T, U1, H, lstride, V, plan.ntub, int(bf16), _p(a16), _p(row_label),
                  None, plan.idx, st)          # (no transposed copy: the weight gradient reads its operands MN-major)
            if side is not None:
                main.wait_stream(side)
            lse, lpb, lpl, pfac, mref = (plan.rowf() for _ in range(5))
            ew = plan.rowf(H) if need_act else None
            chunks, store_rows, pstore = _wide_pstore(plan, Vpad, dev)
            flags = torch.zeros((plan.ntub + 1) // 2 + 1, dtype=torch.int32, device=dev)
            # one chunk (the usual case): EW = P' . W runs on the high-priority stream while the lattice runs here
            main, hi = torch.cuda.current_stream(dev), None
            if need_act and len(chunks) == 1 and WideJointRNNT.OVERLAP:
                hi = _hi_stream(dev)
            for t0, cnt in chunks:
                WideJointRNNT._sp(dev, plan, st, a16, w16, bias2, scal, row_label, t0, cnt, H, V, blank, bf16, lse, lpb, lpl,
                                  pfac, mref, pstore, store_rows, flags)
                if need_act and hi is None:
                    _call("ttx_wide_pw", dev, _p(pstore), store_rows, _p(w16t), _p(pfac), _p(scal), _p(plan.meta), plan.ntub,
                          t0, cnt, H, V, int(bf16), _p(ew), plan.idx, st)
            if hi is not None:
                hi.wait_stream(main)
                with torch.cuda.stream(hi):
                    _call("ttx_wide_pw", dev, _p(pstore), store_rows, _p(w16t), _p(pfac), _p(scal), _p(plan.meta), plan.ntub,
                          0, plan.ntub, H, V, int(bf16), _p(ew), plan.idx, _stream(dev))
            alpha, beta, costs, ll_beta = plan.lattice(lse, lpb, lpl)
            if hi is not None:
                main.wait_stream(hi)
        ctx.plan, ctx.blank, ctx.bf16, ctx.dims = plan, int(blank), int(bf16), (B, T, U1, H, V)
        ctx.in_dtypes = (eproj.dtype, pproj.dtype, w_out.dtype, b_out.dtype)
        ctx.chunks, ctx.store_rows = chunks, store_rows
        ctx.ew, ctx.w32 = ew, (w if need_act else None)
        ctx.pstore = pstore if (need_w and len(chunks) == 1) else None      # kept from forward to backward
        ctx.save_for_backward(ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta, pfac, mref)
        return costs

    @staticmethod
    def _sp(dev, plan, st, a16, w16, bias2, scal, row_label, t0, cnt, H, V, blank, bf16, lse, lpb, lpl, pfac, mref, pstore,
            store_rows, flags):
        _call("ttx_wide_sp", dev, _p(a16), _p(w16), _p(bias2), _p(scal), _p(row_label), _p(plan.meta), plan.ntub, t0, cnt,
              H, V, int(blank), int(bf16), _p(lse), _p(lpb), _p(lpl), _p(pfac), _p(mref), _p(pstore), store_rows,
              _p(flags), plan.idx, st)

    @staticmethod
    def backward(ctx, grad_costs):
        (ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta, pfac,
         mref) = ctx.saved_tensors
        plan, (B, T, U1, H, V) = ctx.plan, ctx.dims
        dev = plan.dev
        pre = ctx.pre
        need_act = ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or (pre is not None and any(ctx.needs_input_grad[10:15]))
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        d_ep = d_pp = d_w = d_b = None
        pre_grads = [None] * 5
        with _guard(dev):
            st = _stream(dev)
            scal = scal.clone()
            Vpad = w16.numel() // H
            if need_w:
                d_w = torch.zeros(V, H, dtype=torch.float32, device=dev)
                d_b = torch.zeros(V, dtype=torch.float32, device=dev)
            rowmeta = plan.grad_coeffs(lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label, ctx.blank, d_b)
            if need_w:
                a16st = torch.empty((H + 16) * plan.rows + 64 * (H + 4) * 2, dtype=torch.int16, device=dev)
                pstore = ctx.pstore
                recompute = pstore is None
                if recompute:
                    pstore = torch.empty(ctx.store_rows * Vpad, dtype=torch.int16, device=dev)
                    flags = torch.zeros((plan.ntub + 1) // 2 + 1, dtype=torch.int32, device=dev)
                    sink = [plan.rowf() for _ in range(5)]          # the statistics come out again; only P' is wanted
                # the weight-gradient chain (scaled operand copy, exact blank / label terms, dW = P'^T . As) on the
                # high-priority stream while the activation-gradient reduction runs here
                main, hi = torch.cuda.current_stream(dev), None
                if need_act and WideJointRNNT.OVERLAP:
                    hi = _hi_stream(dev)
                    hi.wait_stream(main)

                def prepare(parts, stream_ptr):
                    _call("ttx_kept_prepare", dev, _p(a16), _p(rowmeta), _p(row_label), _p(lpb), _p(lpl),
                          _p(pfac), _p(scal), _p(plan.act_lens), _p(plan.label_lens), _p(plan.meta), B, T, U1, plan.ntub, H,
                          ctx.blank, ctx.bf16, _p(a16st), _p(d_w), _p(d_b), parts, plan.idx, stream_ptr,
                          n_kernels=1 if parts == 1 else 2 if parts == 2 else 3,
                          label="ttx_kept_prepare" + {1: "[As]", 2: "[blank/label]", 3: ""}[parts])

                # (measured: with the blank / label terms on the launching stream next to the reduction that chain becomes
                # the longer one -- a co-resident block per SM is all the reduction gets while dW runs: 7.89 vs 7.77 ms)
                with torch.cuda.stream(hi if hi is not None else main):
                    st_w = _stream(dev)
                    prepare(3, st_w)
                    for t0, cnt in ctx.chunks:
                        if recompute:
                            WideJointRNNT._sp(dev, plan, st_w, a16, w16, bias2, scal, row_label, t0, cnt, H, V, ctx.blank,
                                              ctx.bf16, sink[0], sink[1], sink[2], sink[3], sink[4], pstore, ctx.store_rows,
                                              flags)
                        _call("ttx_wide_dw", dev, _p(pstore), ctx.store_rows, _p(a16st), _p(scal), _p(plan.meta), plan.ntub,
                              t0, cnt, H, V, ctx.bf16, _p(d_w), _p(d_b), plan.idx, st_w)
            if need_act:
                d_ep = torch.zeros(B, T, H, dtype=torch.float32, device=dev)
                d_pp = torch.zeros(B, U1, H, dtype=torch.float32, device=dev)
                _call("ttx_reduce_act_grad_ew", dev, _p(ctx.ew), _p(rowmeta), _p(row_label), _p(ctx.w32), _p(scal),
                      ctx.blank, _p(ep), _p(pp), _p(plan.act_lens), _p(plan.label_lens), _p(plan.meta), B, T, U1, H,
                      _p(d_ep), _p(d_pp), plan.idx, st)
                if pre is not None:
                    # the pre-projections' backward, here, behind the reduction and next to the weight-gradient product
                    x_e, w_e, has_b, x_d, w_d, enc_shape, dec_shape = pre
                    ge = _proj_backward(dev, d_ep.view(-1, H), x_e, w_e, has_b, ctx.needs_input_grad[10],
                                        ctx.needs_input_grad[11] or ctx.needs_input_grad[12], plan.idx, st)
                    gd = _proj_backward(dev, d_pp.view(-1, H), x_d, w_d, False, ctx.needs_input_grad[13],
                                        ctx.needs_input_grad[14], plan.idx, st)
                    pre_grads = [ge[0].view(enc_shape) if ge[0] is not None else None, ge[1], ge[2],
                                 gd[0].view(dec_shape) if gd[0] is not None else None, gd[1]]
            if need_w:
                if hi is not None:
                    main.wait_stream(hi)
                ctx.pstore = None
        dt = ctx.in_dtypes
        cast = lambda g, d, need: g.to(d) if (g is not None and need) else None  # noqa: E731
        return (cast(d_ep, dt[0], ctx.needs_input_grad[0]), cast(d_pp, dt[1], ctx.needs_input_grad[1]),
                cast(d_w, dt[2], ctx.needs_input_grad[2]), cast(d_b, dt[3], ctx.needs_input_grad[3]),
                None, None, None, None, None, None) + tuple(
                    g if need else None for g, need in zip(pre_grads, ctx.needs_input_grad[10:15]))


def _proj_backward(dev, dy, x, w, has_bias, need_x, need_w, idx, st):
    """dx, dw, db of y = x w^T (+ b) on the projection kernels (csrc/ttx_proj.cu); dy (M, N) contiguous float32."""
    M, N = dy.shape
    K = x.shape[1]
    dx = dw = db = None
    if need_x:
        dx = torch.empty(M, K, dtype=torch.float32, device=dev)
        _call("ttx_proj_bwd_x", dev, _p(dy), N, _p(w), w.stride(0), M, N, K, _p(dx), K, idx, st)
    if need_w:
        dw = torch.zeros(N, K, dtype=torch.float32, device=dev)
        db = torch.zeros(N, dtype=torch.float32, device=dev) if has_bias else None
        _call("ttx_proj_bwd_w", dev, _p(dy), N, _p(x), K, M, N, K, _p(dw), K, _p(db), idx, st,
              n_kernels=2 if db is not None else 1)
    return dx, dw, db


class ChunkedJointRNNT(torch.autograd.Function):
    """Same contract as FusedJointRNNT for joint widths the fused tensor-core kernels do not cover (any multiple of
    64, e.g. aishell.yaml's 1024 and joint_streaming.yaml's 2048): lattice rows are processed in chunks, the
    projection of a chunk is a library GEMM on the same 16-bit operands (fp32 accumulate), our row kernels do the
    log-softmax statistics / the gradient operand, and memory stays bounded by the chunk."""

    CHUNK_BYTES = int(os.environ.get("TTX_CHUNK_MB", "512")) << 20

    @staticmethod
    def _chunks(plan, Vpad):
        tiles = max(1, ChunkedJointRNNT.CHUNK_BYTES // (128 * Vpad * 4))
        return [(t0, min(plan.ntub, t0 + tiles)) for t0 in range(0, plan.ntub, tiles)]

    @staticmethod
    def forward(ctx, eproj, pproj, w_out, b_out, labels, act_lens, label_lens, blank, bf16, sizes=None):
        if not eproj.is_cuda:
            raise RuntimeError("fused_joint_rnnt needs CUDA tensors (there is no CPU fallback)")
        dev = eproj.device
        B, T, H = eproj.shape
        U1 = pproj.shape[1]
        V = w_out.shape[0]
        if H % 64 != 0:
            raise ValueError("joint width %d must be a multiple of 64" % H)
        ep, pp = eproj.detach().float().contiguous(), pproj.detach().float().contiguous()
        w, b = w_out.detach().float().contiguous(), b_out.detach().float().contiguous()
        labels = labels.contiguous()
        dt16 = torch.bfloat16 if bf16 else torch.float16
        with _guard(dev):
            plan = _Plan(B, T, U1, dev, act_lens, label_lens, sizes)
            st = _stream(dev)
            Vpad = (V + 255) // 256 * 256
            scal = torch.zeros(8, dtype=torch.float32, device=dev)
            w16 = torch.empty(Vpad * H, dtype=torch.int16, device=dev)
            bias2 = torch.empty(Vpad, dtype=torch.float32, device=dev)
            _call("ttx_cast_weight", dev, _p(w), _p(b), V, H, int(bf16), _p(scal), _p(w16), _p(bias2), None, plan.idx, st)
            a16 = torch.zeros(plan.rows * H, dtype=torch.int16, device=dev)   # unused tiles must be zeros, not NaNs
            row_label = torch.full((plan.rows,), -1, dtype=torch.int32, device=dev)
            lstride = labels.shape[1] if labels.dim() == 2 else 0
            _call("ttx_joint_act", dev, _p(ep), _p(pp), _p(labels) if labels.numel() else None, _p(act_lens),
                  _p(label_lens), _p(plan.meta), B, T, U1, H, lstride, V, plan.ntub, int(bf16), _p(a16), _p(row_label),
                  None, plan.idx, st)
            lse, lpb, lpl = (torch.zeros(plan.rows, dtype=torch.float32, device=dev) for _ in range(3))
            a2, w2 = a16.view(dt16).view(plan.rows, H), w16.view(dt16).view(Vpad, H)
            for t0, t1 in ChunkedJointRNNT._chunks(plan, Vpad):
                r0, r1 = t0 * 128, t1 * 128
                z = torch.mm(a2[r0:r1], w2.t(), out_dtype=torch.float32)
                _call("ttx_rows_lse", dev, _p(z), r1 - r0, Vpad, V, _p(bias2), _p(scal), _p(row_label[r0:]), int(blank),
                      _p(lse[r0:]), _p(lpb[r0:]), _p(lpl[r0:]), plan.idx, st)
                del z
            alpha, beta, costs, ll_beta = plan.lattice(lse, lpb, lpl)
        ctx.plan, ctx.blank, ctx.bf16, ctx.dims = plan, int(blank), int(bf16), (B, T, U1, H, V)
        ctx.in_dtypes = (eproj.dtype, pproj.dtype, w_out.dtype, b_out.dtype)
        ctx.save_for_backward(ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta)
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        ep, pp, bias2, a16, w16, scal, row_label, lse, lpb, lpl, alpha, beta, ll_beta = ctx.saved_tensors
        plan, (B, T, U1, H, V) = ctx.plan, ctx.dims
        dev = plan.dev
        dt16 = torch.bfloat16 if ctx.bf16 else torch.float16
        need_act = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        need_w = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        d_ep = d_pp = d_w = d_b = None
        with _guard(dev):
            st = _stream(dev)
            scal = scal.clone()
            Vpad = w16.numel() // H
            rowmeta = torch.zeros(plan.rows * 4, dtype=torch.float32, device=dev)
            rm_new = plan.grad_coeffs(lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label, ctx.blank, None)
            # rows of tiles not in use keep w = 0 (grad_coeffs only writes the tiles in use)
            n_rows_used = plan.meta[0].to(torch.int64) * 128
            used = (torch.arange(plan.rows, device=dev) < n_rows_used).view(-1, 1)
            rowmeta = torch.where(used, rm_new.view(-1, 4), rowmeta.view(-1, 4)).contiguous().view(-1)
            pscale = 1.0 if ctx.bf16 else 4096.0
            a2, w2 = a16.view(dt16).view(plan.rows, H), w16.view(dt16).view(Vpad, H)
            d_act = torch.empty(plan.rows, H, dtype=torch.float32, device=dev) if need_act else None
            if need_w:
                d_w_pad = torch.zeros(Vpad, H, dtype=torch.float32, device=dev)
                d_b_pad = torch.zeros(Vpad, dtype=torch.float32, device=dev)
            for t0, t1 in ChunkedJointRNNT._chunks(plan, Vpad):
                r0, r1 = t0 * 128, t1 * 128
                z = torch.mm(a2[r0:r1], w2.t(), out_dtype=torch.float32)
                q = torch.empty(r1 - r0, Vpad, dtype=dt16, device=dev)
                _call("ttx_rows_grad", dev, _p(z), _p(rowmeta[4 * r0:]), _p(row_label[r0:]), _p(bias2), _p(scal),
                      r1 - r0, Vpad, V, ctx.blank, ctx.bf16, _p(q), plan.idx, st)
                del z
                if need_act:
                    torch.mm(q, w2, out_dtype=torch.float32, out=d_act[r0:r1])
                if need_w:
                    d_w_pad += torch.mm(q.t(), a2[r0:r1], out_dtype=torch.float32)
                    d_b_pad += q.sum(0, dtype=torch.float32)
                del q
            if need_act:
                d_act.mul_(scal[2] * scal[1] / pscale)
                d_ep = torch.empty(B, T, H, dtype=torch.float32, device=dev)
                d_pp = torch.empty(B, U1, H, dtype=torch.float32, device=dev)
                _call("ttx_reduce_act_grad", dev, _p(d_act), _p(ep), _p(pp), _p(plan.act_lens), _p(plan.label_lens),
                      _p(plan.meta), B, T, U1, H, _p(d_ep), _p(d_pp), plan.idx, st)
            if need_w:
                d_w = d_w_pad[:V] * (scal[2] / pscale)
                d_b = d_b_pad[:V] * (scal[2] / pscale)
        dt = ctx.in_dtypes
        cast = lambda g, d, need: g.to(d) if (g is not None and need) else None  # noqa: E731
        return (cast(d_ep, dt[0], ctx.needs_input_grad[0]), cast(d_pp, dt[1], ctx.needs_input_grad[1]),
                cast(d_w, dt[2], ctx.needs_input_grad[2]), cast(d_b, dt[3], ctx.needs_input_grad[3]),
                None, None, None, None, None, None)


# None: the default routing below.  "fused": the recomputing kernels also at H = 512 (nothing V-wide in HBM).  "chunked":
# the library-GEMM fallback for every width.  (Tests and bench.py A/B the routes through this.)
ROUTE = None
MERGE_PROJ_BACKWARD = True      # False: the pre-projections keep their own autograd nodes (A/B runs)


def fused_joint_rnnt(eproj, pproj, w_out, b_out, labels, act_lens, label_lens, blank=0, bf16=False, sizes=None, pre=None):
    """costs (B,) fp32 of the transducer loss of logits = tanh(eproj[:, :, None] + pproj[:, None]) @ w_out.T + b_out.

    H a multiple of 512: three streamed tcgen05 products around the kept 16-bit softmax numerators (WideJointRNNT);
    smaller multiples of 64 (<= 256, 384): the fused recomputing kernels (FusedJointRNNT); other multiples of 64: the
    chunked library-GEMM fallback."""
    dev = eproj.device
    H = eproj.shape[-1]
    if ROUTE == "chunked":
        fn = ChunkedJointRNNT
    elif wide_width(H) and ROUTE != "fused":                            # 512, 1024, 1536, ...
        fn = WideJointRNNT
    else:
        fn = FusedJointRNNT if supported_width(H) else ChunkedJointRNNT
    args = (_i32_cuda(labels, dev, "labels"), _i32_cuda(act_lens, dev, "act_lens"), _i32_cuda(label_lens, dev, "label_lens"),
            blank, bf16, sizes)
    if fn is WideJointRNNT and pre is not None and MERGE_PROJ_BACKWARD:
        # pre = (enc, w_enc, b_enc, dec, w_dec): this node owns the pre-projections' backward (see WideJointRNNT.forward);
        # eproj / pproj go in detached, so their own autograd nodes are never run
        return fn.apply(eproj.detach(), pproj.detach(), w_out, b_out, *args, *pre)
    return fn.apply(eproj, pproj, w_out, b_out, *args)


class DenseRNNT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, acts, labels, act_lens, label_lens, blank, sizes=None):
        if not acts.is_cuda:
            raise RuntimeError("rnnt_loss needs CUDA tensors (there is no CPU fallback)")
        dev = acts.device
        B, T, U1, V = acts.shape
        a = acts.detach().float().contiguous()
        labels = labels.contiguous()
        lib = _lib.get()
        with _guard(dev):
            plan = _Plan(B, T, U1, dev, act_lens, label_lens, sizes)
            lse, lpb, lpl = plan.rowf(), plan.rowf(), plan.rowf()
            row_label = torch.empty(plan.rows, dtype=torch.int32, device=dev)
            lstride = labels.shape[1] if labels.dim() == 2 else 0
            _call("ttx_dense_lse", dev, _p(a), _p(labels) if labels.numel() else None, _p(act_lens), _p(label_lens),
                                         _p(plan.meta), B, T, U1, V, lstride, int(blank), plan.ntub, _p(lse), _p(lpb),
                                         _p(lpl), _p(row_label), plan.idx, _stream(dev))
            alpha, beta, costs, ll_beta = plan.lattice(lse, lpb, lpl)
        ctx.plan, ctx.blank, ctx.in_dtype = plan, int(blank), acts.dtype
        ctx.save_for_backward(a, row_label, lse, lpb, lpl, alpha, beta, ll_beta)
        return costs

    @staticmethod
    def backward(ctx, grad_costs):
        a, row_label, lse, lpb, lpl, alpha, beta, ll_beta = ctx.saved_tensors
        plan = ctx.plan
        B, T, U1, V = a.shape
        with _guard(plan.dev):
            scal = torch.zeros(8, dtype=torch.float32, device=plan.dev)
            rowmeta = plan.grad_coeffs(lse, lpb, lpl, alpha, beta, ll_beta, grad_costs, scal, row_label, ctx.blank)
            grads = torch.empty_like(a)
            _call("ttx_dense_grad", plan.dev, _p(a), _p(rowmeta), _p(row_label), _p(scal), _p(plan.act_lens),
                                               _p(plan.label_lens), _p(plan.meta), B, T, U1, V, ctx.blank, _p(grads),
                                               plan.idx, _stream(plan.dev))
        return grads.to(ctx.in_dtype), None, None, None, None, None


def dense_rnnt(acts, labels, act_lens, label_lens, blank=0, sizes=None):
    dev = acts.device
    return DenseRNNT.apply(acts, _i32_cuda(labels, dev, "labels"), _i32_cuda(act_lens, dev, "act_lens"),
                           _i32_cuda(label_lens, dev, "label_lens"), blank, sizes)
