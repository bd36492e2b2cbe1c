"""GPU parity gate (-m gpu): the CUDA path through the public drop-in API vs the CPU oracle and the committed
golden fixtures.  Tolerances (BASELINE.json north_star): fp32 variant <= 1e-4 relative on per-utterance loss,
<= 1e-3 relative (L2 per tensor) on gradients; bf16-input variant stated separately below; exact zeros outside
the ragged region.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest
import torch

import transformer_transducer_b200 as ttb
from oracle import joint_ref, rnnt_oracle
from transformer_transducer_b200 import _lib

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOSS_TOL, GRAD_TOL = 1e-4, 1e-3
BF16_LOSS_TOL, BF16_GRAD_TOL = 2e-3, 1.5e-2     # bf16-input variant, vs the oracle fed the same bf16-rounded inputs


def _i32(x, dev="cpu"):
    return torch.as_tensor(x, dtype=torch.int32, device=dev)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_native_library_is_loaded():
    lib = _lib.get()
    assert lib.ttx_version() == 1
    assert any("libttx.so" in l for l in open("/proc/self/maps"))


# ----------------------------------------------------------------------------- dense-logits entry (RNNTLoss on acts)
def test_known_answer_vector_on_gpu(golden_dir):
    k = json.load(open(os.path.join(golden_dir, "warp_transducer_kat.json")))
    acts = torch.tensor(k["acts"], device=DEV, requires_grad=True)
    c = ttb.rnnt_loss(acts, _i32(k["labels"], DEV), _i32(k["act_lens"], DEV), _i32(k["label_lens"], DEV), 0, "none")
    c.sum().backward()
    assert abs(float(c[0]) - k["cost_upstream"]) < 5e-6
    np.testing.assert_allclose(acts.grad.cpu().numpy(), np.array(k["grads_upstream"], dtype=np.float32), atol=2e-6)


def test_ragged_fixture_on_gpu_exact_zero_outside(golden_dir):
    g = np.load(os.path.join(golden_dir, "ragged_loss.npz"))
    logits = torch.tensor(g["logits"], device=DEV, requires_grad=True)
    c = ttb.rnnt_loss(logits, _i32(g["labels"], DEV), _i32(g["act_lens"], DEV), _i32(g["label_lens"], DEV), 0, "none")
    c.sum().backward()
    np.testing.assert_allclose(c.detach().cpu().numpy(), g["costs"], rtol=LOSS_TOL)
    assert rel(logits.grad, torch.tensor(g["grads"])) < GRAD_TOL
    gr = logits.grad
    assert gr[1, 6:].abs().max() == 0 and gr[1, :, 3:].abs().max() == 0      # t >= T_b, u > U_b
    assert gr[2, 1:].abs().max() == 0 and gr[2, :, 1:].abs().max() == 0      # T_b = 1, U_b = 0


def test_espnet_transloss_fixture_reductions_and_grad_output(golden_dir):
    g = np.load(os.path.join(golden_dir, "espnet_transloss.npz"))
    args = (_i32(g["target"], DEV), _i32(g["pred_len"], DEV), _i32(g["target_len"], DEV))
    pred = torch.tensor(g["pred"], device=DEV, requires_grad=True)
    loss = ttb.RNNTLoss(blank=0)(pred, *args)
    assert loss.shape == (1,)
    loss.backward()
    np.testing.assert_allclose(loss.detach().cpu().numpy(), g["loss"], rtol=LOSS_TOL)
    assert rel(pred.grad, torch.tensor(g["grad"])) < GRAD_TOL
    none = ttb.rnnt_loss(pred, *args, reduction="none")
    total = ttb.RNNTLoss(reduction="sum")(pred, *args)
    assert none.shape == (2,) and total.shape == (1,)
    assert torch.allclose(total, none.sum().view(1), rtol=1e-6) and torch.allclose(loss, total / 2, rtol=1e-6)
    p2 = torch.tensor(g["pred"], device=DEV, requires_grad=True)
    (ttb.RNNTLoss(blank=0)(p2, *args) * 3.0).sum().backward()           # backward scales by grad_output
    assert rel(p2.grad, 3.0 * torch.tensor(g["grad"])) < GRAD_TOL


def test_error_conventions_on_gpu():
    acts = torch.zeros(2, 3, 2, 5, device=DEV)
    lab, al, ll = _i32([[1], [1]], DEV), _i32([3, 2], DEV), _i32([1, 1], DEV)
    ttb.rnnt_loss(acts, lab, al, ll)
    with pytest.raises(TypeError):
        ttb.rnnt_loss(acts, lab.long(), al, ll)
    with pytest.raises(ValueError, match="Input length mismatch"):
        ttb.rnnt_loss(acts, lab, _i32([2, 2], DEV), ll)
    with pytest.raises(ValueError, match="Output length mismatch"):
        ttb.rnnt_loss(acts, lab, al, _i32([0, 0], DEV))
    with pytest.raises(ValueError):
        ttb.rnnt_loss(acts, lab, al, ll, reduction="avg")


# ----------------------------------------------------------------------------- lattice kernel through the C ABI
@pytest.mark.parametrize("T,U", [(37, 0), (1, 5), (60, 31), (50, 40), (45, 100), (40, 200), (30, 300), (21, 700)])
def test_lattice_wavefront_matches_float64_oracle(T, U):
    """ttx_prepare + ttx_lattice_fwd_bwd alone: every alpha / beta cell, the costs and beta(0,0) against the float64
    restatement in oracle/rnnt_cpu.c, on a ragged batch.  Column counts cover all kernel shapes: 1 / 2 / 4 / 8 columns
    per lane in a single warp (U+1 <= 32 ... 256) and the multi-warp hand-off (2 and 4 warps)."""
    import ctypes
    lib = _lib.get()
    g = torch.Generator().manual_seed(T * 1000 + U)
    B, U1 = 4, U + 1
    al = _i32([T, max(1, T - 3), 1, max(1, T // 2)])
    ll = _i32([U, max(0, U - 5), U // 2, 0])
    lpb = -torch.rand(B, T, U1, generator=g) * 3 - 0.05
    lpl = -torch.rand(B, T, U1, generator=g) * 9 - 0.5
    alpha_w, beta_w, costs_w = rnnt_oracle.lattice_fp64(lpb, lpl, al, ll)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    ntub = int(lib.ttx_tiles_upper_bound(B, T, U1))
    meta = torch.empty(int(lib.ttx_meta_ints(B, ntub)), dtype=torch.int32, device=DEV)
    ald, lld = al.to(DEV), ll.to(DEV)
    _lib.check(lib.ttx_prepare(p(ald), p(lld), B, T, U1, ntub, p(meta), 0, None), "prepare")
    mh = meta.cpu()
    rows = ntub * 128
    lpb_r, lpl_r = torch.zeros(rows), torch.zeros(rows)           # compact row space, as the projection kernels write it
    for b in range(B):
        Tb, U1b, base = int(al[b]), int(ll[b]) + 1, int(mh[4 + b]) * 128
        lpb_r[base: base + Tb * U1b] = lpb[b, :Tb, :U1b].reshape(-1)
        lpl_r[base: base + Tb * U1b] = lpl[b, :Tb, :U1b].reshape(-1)
    lat = int(lib.ttx_lattice_elems_upper_bound(B, T, U1))
    alpha = torch.full((lat,), float("nan"), dtype=torch.float64, device=DEV)
    beta = torch.full((lat,), float("nan"), dtype=torch.float64, device=DEV)
    ws = torch.empty(4 * lat, dtype=torch.float32, device=DEV)
    costs = torch.empty(B, dtype=torch.float32, device=DEV)
    llb = torch.empty(B, dtype=torch.float64, device=DEV)
    lpb_g, lpl_g = lpb_r.to(DEV), lpl_r.to(DEV)
    _lib.check(lib.ttx_lattice_fwd_bwd(p(lpb_g), p(lpl_g), p(ald), p(lld), p(meta), B, U1, ntub, lat, p(ws), p(alpha),
                                       p(beta), p(costs), p(llb), 0, None), "lattice")
    torch.cuda.synchronize()
    alpha, beta = alpha.cpu(), beta.cpu()
    for b in range(B):
        Tb, U1b = int(al[b]), int(ll[b]) + 1
        P, o = (U1b + 3) // 4 * 4, int(mh[4 + B + 1 + ntub + b])
        t, u = torch.arange(Tb).view(-1, 1), torch.arange(U1b).view(1, -1)
        idx = (o + (t + u) * P + u).reshape(-1)
        idx_m = (o + (Tb - 1 - t + U1b - 1 - u) * P + (U1b - 1 - u)).reshape(-1)      # beta: mirrored lattice
        assert (alpha[idx] - alpha_w[b, :Tb, :U1b].reshape(-1)).abs().max() < 2e-5
        assert (beta[idx_m] - beta_w[b, :Tb, :U1b].reshape(-1)).abs().max() < 2e-5
    assert ((costs.double().cpu() - costs_w) / costs_w).abs().max() < 1e-6
    assert ((-llb.cpu() - costs_w) / costs_w).abs().max() < 1e-6


# ----------------------------------------------------------------------------- pre-projection kernels through the C ABI
@pytest.mark.parametrize("M,N,K,ldw", [(12800, 512, 512, 512), (1312, 512, 512, 1024), (37, 128, 100, 100), (300, 1024, 320, 640),
                                       (129, 260, 36, 36)])
def test_projection_kernels_are_fp32_grade(M, N, K, ldw):
    """ttx_proj_fwd / _bwd_x / _bwd_w (tcgen05 TF32 products with on-chip hi/lo error compensation) against float64, on
    shapes with row / column / contraction tails and a weight that is a column slice of a wider matrix.  Tolerance 2e-5
    relative L2: the compensated product itself is good to 2^-22, what remains is the tensor core's truncating fp32
    accumulation (measured 3e-6 .. 8e-6); a plain TF32 product sits at 3e-4, the reference's SGEMM (tt/model.py:35,
    joint_network.py:28-31) at 3e-7, and the path's gradient tolerance is 1e-3."""
    import ctypes
    lib = _lib.get()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g)
    wfull = torch.randn(N, ldw, generator=g) / K ** 0.5
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    w = wfull[:, ldw - K:]
    xd, wd, bd, dyd = x.to(DEV), wfull.to(DEV)[:, ldw - K:], b.to(DEV), dy.to(DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    y = torch.full((M, N), float("nan"), device=DEV)
    dx = torch.full((M, K), float("nan"), device=DEV)
    dw, db = torch.zeros(N, K, device=DEV), torch.zeros(N, device=DEV)
    _lib.check(lib.ttx_proj_fwd(p(xd), K, p(wd), ldw, p(bd), M, N, K, p(y), N, 0, None), "proj_fwd")
    _lib.check(lib.ttx_proj_bwd_x(p(dyd), N, p(wd), ldw, M, N, K, p(dx), K, 0, None), "proj_bwd_x")
    _lib.check(lib.ttx_proj_bwd_w(p(dyd), N, p(xd), K, M, N, K, p(dw), K, p(db), 0, None), "proj_bwd_w")
    torch.cuda.synchronize()
    x64, w64, dy64 = x.double(), w.double(), dy.double()
    for got, want, ref32 in ((y, x64 @ w64.T + b.double(), x @ w.T + b), (dx, dy64 @ w64, dy @ w),
                             (dw, dy64.T @ x64, dy.T @ x), (db, dy64.sum(0), dy.sum(0))):
        err, err32 = rel(got, want), rel(ref32, want)
        assert err < max(2e-5, 3 * err32), (err, err32)


# ----------------------------------------------------------------------------- fused path vs oracle
def _espnet_case(B, T, U, V, D, H, act_lens, label_lens, seed=0, dtype=torch.float32):
    torch.manual_seed(seed)
    ref = joint_ref.EspnetJointNetwork(V, D, D, H, "tanh")
    mine = ttb.JointNetwork(V, D, D, H, "tanh")
    mine.load_state_dict(ref.state_dict())
    enc, pred = torch.randn(B, T, D), torch.randn(B, U + 1, D)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    for i, n in enumerate(label_lens):
        labels[i, n:] = -1                      # tt/dataset.py:46-48 pads with ignore_id = -1
    return ref, mine, enc, pred, labels, _i32(act_lens), _i32(label_lens)


def _run_pair(ref, mine, enc, pred, labels, al, ll, weights=None, dtype=torch.float32, arbiter64=False):
    mine = mine.to(DEV).to(dtype)
    if dtype != torch.float32:                  # oracle sees the same rounded parameters / inputs
        ref.load_state_dict({k: v.float().cpu() for k, v in mine.state_dict().items()})
        enc, pred = enc.to(dtype).float(), pred.to(dtype).float()
    if arbiter64:                               # float64 oracle (oracle_rnnt_f64): see test_fused_long_utterance_lattice
        ref = ref.double()
    e0 = enc.clone().to(torch.float64 if arbiter64 else torch.float32).requires_grad_()
    p0 = pred.clone().to(e0.dtype).requires_grad_()
    want = rnnt_oracle.rnnt_loss(ref(e0[:, :, None], p0[:, None]), labels, al, ll, 0, "none")
    wts = torch.ones(len(al)) if weights is None else weights
    (want * wts).sum().backward()
    e1 = enc.to(DEV).to(dtype).requires_grad_()
    p1 = pred.to(DEV).to(dtype).requires_grad_()
    z = mine(e1[:, :, None], p1[:, None])
    assert isinstance(z, ttb.LazyJointLogits)
    got = ttb.rnnt_loss(z, labels.to(DEV), al.to(DEV), ll.to(DEV), 0, "none")
    (got.float() * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    errs = {"loss": float(((got.detach().double().cpu() - want.detach().double()) / want.detach().double()).abs().max()),
            "d_enc": rel(e1.grad, e0.grad), "d_pred": rel(p1.grad, p0.grad)}
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        errs["d_" + n] = rel(a.grad, b.grad)
    for b in range(len(al)):                    # per utterance too: a large grad_output must not hide a small one
        errs["d_enc[%d]" % b] = rel(e1.grad[b], e0.grad[b])
        errs["d_pred[%d]" % b] = rel(p1.grad[b], p0.grad[b])
    return errs, (e1, p1, got)


def _check(errs, lt=LOSS_TOL, gt=GRAD_TOL):
    assert errs["loss"] < lt, errs
    assert all(v < gt for k, v in errs.items() if k != "loss"), errs


@pytest.mark.parametrize("H", [64, 128, 192, 256, 384, 512])
def test_fused_espnet_joint_matches_oracle_all_widths(H):
    case = _espnet_case(3, 33, 7, 300, 64, H, [33, 29, 9], [7, 4, 0], seed=H)
    errs, _ = _run_pair(*case, weights=torch.tensor([1.0, 0.5, 2.0]))
    _check(errs)


@pytest.mark.parametrize("H", [128, 512])
def test_fused_forward_grad_rescale_path_with_outlier_logits(H):
    """Peaked distributions whose largest logits sit in LATE vocabulary tiles: the forward's running reference has to
    move after the first tile, which rescales the TMEM accumulator of the fused forward+gradient mode."""
    ref, mine, enc, pred, labels, al, ll = _espnet_case(2, 30, 6, 1500, 64, H, [30, 22], [6, 3], seed=77)
    with torch.no_grad():
        ref.lin_out.bias[700] += 9.0            # tile 2 of 6 (256-wide tiles)
        ref.lin_out.bias[1490] += 14.0          # last tile
        ref.lin_out.bias[0] += 5.0              # blank, first tile
        ref.lin_out.weight[1301] *= 6.0         # row-dependent outliers in tile 5
    labels[0, 2] = 1490
    labels[1, 0] = 700
    mine.load_state_dict(ref.state_dict())
    errs, _ = _run_pair(ref, mine, enc, pred, labels, al, ll, weights=torch.tensor([1.0, 3.0]))
    _check(errs)


@pytest.mark.parametrize("H,V", [(128, 150), (512, 150), (512, 97), (1024, 300)])     # fused kernels / streamed products
def test_fused_edge_lengths_t1_u0_and_zero_grads_outside(H, V):
    """T = 1 and U = 0 utterances, a vocabulary smaller than one 256-column chunk, exact zeros outside the ragged region."""
    case = _espnet_case(4, 20, 5, V, 32, H, [20, 1, 1, 7], [5, 0, 3, 0], seed=3)
    errs, (e1, p1, _) = _run_pair(*case)
    _check(errs)
    assert e1.grad[1, 1:].abs().max() == 0 and p1.grad[1, 1:].abs().max() == 0      # T_b = 1, U_b = 0
    assert e1.grad[3, 7:].abs().max() == 0 and p1.grad[2, 4:].abs().max() == 0


def test_fused_microbench_dims_small_batch():
    """configs[1] dims (D=H=512, V=4232) at a batch the CPU oracle finishes in seconds."""
    case = _espnet_case(2, 60, 12, 4232, 512, 512, [60, 47], [12, 9], seed=5)
    errs, _ = _run_pair(*case)
    _check(errs)


def test_fused_long_utterance_lattice():
    """configs[3]-like lattice extents (T=1000, U=200) with a small vocabulary so the oracle stays fast.

    At this lattice size alpha/beta reach ~6e3 and the reference's float32 recursion itself is only good to
    ~1.1e-3 on gradients (float32 vs float64 run of the same oracle, checked below), so the arbiter here is the
    float64 oracle; the CUDA lattice carries alpha/beta in float64 for the same reason."""
    case = _espnet_case(1, 1000, 200, 130, 32, 64, [1000], [200], seed=6)
    errs, _ = _run_pair(*case, arbiter64=True)
    _check(errs)
    ref, _, enc, pred, labels, al, ll = _espnet_case(1, 1000, 200, 130, 32, 64, [1000], [200], seed=6)
    z = ref(enc[:, :, None], pred[:, None]).detach()
    z32, z64 = z.clone().requires_grad_(), z.double().requires_grad_()
    rnnt_oracle.rnnt_loss(z32, labels, al, ll, 0, "none").sum().backward()
    rnnt_oracle.rnnt_loss(z64, labels, al, ll, 0, "none").sum().backward()
    assert 2e-4 < rel(z32.grad, z64.grad) < 5e-3        # the float32 reference's own rounding at this size


def test_fused_tt_jointnet_split_first_layer_matches_oracle():
    torch.manual_seed(7)
    B, T, U, V, D, H = 2, 25, 6, 211, 48, 256
    ref = joint_ref.TTJointNet(2 * D, H, V)
    mine = ttb.JointNet(2 * D, H, V)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    enc, dec = torch.randn(B, T, D), torch.randn(B, U + 1, D)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    al, ll = _i32([T, T - 3]), _i32([U, U - 1])
    e0, d0 = enc.clone().requires_grad_(), dec.clone().requires_grad_()
    want = rnnt_oracle.RNNTLoss()(ref(e0, d0), labels, al, ll)          # train.py:51-53
    want.backward()
    e1, d1 = enc.to(DEV).requires_grad_(), dec.to(DEV).requires_grad_()
    z = mine(e1, d1)
    assert isinstance(z, ttb.LazyJointLogits) and tuple(z.shape) == (B, T, U + 1, V)
    got = ttb.RNNTLoss()(z, labels.to(DEV), al.to(DEV), ll.to(DEV))
    got.backward()
    assert abs(float(got) - float(want)) / float(want) < LOSS_TOL
    assert rel(e1.grad, e0.grad) < GRAD_TOL and rel(d1.grad, d0.grad) < GRAD_TOL
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel(a.grad, b.grad) < GRAD_TOL, n


def _tt_wide_case(H, V, B=3, T=40, U=9, D=64):
    torch.manual_seed(H)
    ref = joint_ref.TTJointNet(2 * D, H, V)
    mine = ttb.JointNet(2 * D, H, V)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    enc, dec = torch.randn(B, T, D), torch.randn(B, U + 1, D)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    al, ll = _i32([T, T - 9, 7]), _i32([U, U - 4, 0])
    labels[1, U - 4:] = -1
    labels[2, :] = -1
    wts = torch.tensor([1.0, 2.0, 0.5])
    e0, d0 = enc.clone().requires_grad_(), dec.clone().requires_grad_()
    want = rnnt_oracle.rnnt_loss(ref(e0, d0), labels, al, ll, 0, "none")
    (want * wts).sum().backward()
    e1, d1 = enc.to(DEV).requires_grad_(), dec.to(DEV).requires_grad_()
    z = mine(e1, d1)
    assert isinstance(z, ttb.LazyJointLogits)
    got = ttb.rnnt_loss(z, labels.to(DEV), al.to(DEV), ll.to(DEV), 0, "none")
    (got * wts.to(DEV)).sum().backward()
    assert float(((got.cpu() - want.detach()) / want.detach()).abs().max()) < LOSS_TOL
    assert rel(e1.grad, e0.grad) < GRAD_TOL and rel(d1.grad, d0.grad) < GRAD_TOL
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel(a.grad, b.grad) < GRAD_TOL, (n, rel(a.grad, b.grad))
    assert e1.grad[2, 7:].abs().max() == 0 and d1.grad[2, 1:].abs().max() == 0


@pytest.mark.parametrize("H,V,keep", [(1024, 4232, "32"), (2048, 700, "32"), (2048, 6485, "32"), (1024, 4232, "1e-9"),
                                      (2048, 6485, "1e-9"), (1536, 900, "32")])
def test_wide_joint_tcgen05_path_matches_oracle(H, V, keep, monkeypatch):
    """aishell.yaml (H=1024, V=4232) / joint_streaming.yaml (H=2048, V=6485) joint widths through the tt JointNet: lazy
    handle + the three streamed tcgen05 products of csrc/ttx_wide.cu, ragged lengths.  keep = "32": P' for the whole
    batch, kept for the backward; keep = "1e-9": a budget of one tile pair -- several chunks, and the backward recomputes
    each chunk's P'."""
    from transformer_transducer_b200 import functional as Fn
    monkeypatch.setenv("TTX_KEEP_GB", keep)
    calls = []
    orig = Fn.WideJointRNNT.apply
    monkeypatch.setattr(Fn.WideJointRNNT, "apply", lambda *a: (calls.append(1), orig(*a))[1])
    _tt_wide_case(H, V)
    assert calls, "the wide tcgen05 path did not run"


def test_wide_path_is_the_default_at_h512_and_matches_oracle(monkeypatch):
    """H = 512 takes the streamed products around the kept P' by default (the fused recomputing kernels are the
    memory-bounded alternative, see test_kernel_variants_agree_with_oracle)."""
    from transformer_transducer_b200 import functional as Fn
    calls = []
    orig = Fn.WideJointRNNT.apply
    monkeypatch.setattr(Fn.WideJointRNNT, "apply", lambda *a: (calls.append(1), orig(*a))[1])
    case = _espnet_case(3, 47, 10, 1100, 64, 512, [47, 31, 6], [10, 8, 1], seed=21)
    errs, _ = _run_pair(*case, weights=torch.tensor([1.0, -0.5, 2.0]))
    _check(errs)
    assert calls, "the wide tcgen05 path did not run"


def test_odd_wide_width_takes_library_gemm_fallback_and_matches_oracle(monkeypatch):
    """A multiple of 64 that is neither a fused width nor a multiple of 512 (768): the chunked fallback (library GEMM on
    the 16-bit operands + our row kernels), several chunks."""
    from transformer_transducer_b200 import functional as Fn
    monkeypatch.setattr(Fn.ChunkedJointRNNT, "CHUNK_BYTES", 3 * 128 * 1024 * 4)   # 3 tiles / chunk
    _tt_wide_case(768, 900)


def test_unsupported_width_uses_dense_entry_and_matches_oracle():
    """A joint width that is not a multiple of 64 is outside every fused path: the module returns dense logits and
    the loss runs through the dense-logits CUDA entry."""
    torch.manual_seed(8)
    B, T, U, V, D, H = 2, 12, 4, 97, 32, 1000
    ref = joint_ref.TTJointNet(2 * D, H, V)
    mine = ttb.JointNet(2 * D, H, V)
    mine.load_state_dict(ref.state_dict())
    mine = mine.to(DEV)
    enc, dec = torch.randn(B, T, D), torch.randn(B, U + 1, D)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32)
    al, ll = _i32([T, T - 2]), _i32([U, U - 2])
    want = rnnt_oracle.RNNTLoss()(ref(enc, dec), labels, al, ll)
    want.backward()
    z = mine(enc.to(DEV), dec.to(DEV))
    assert type(z) is torch.Tensor
    got = ttb.RNNTLoss()(z, labels.to(DEV), al.to(DEV), ll.to(DEV))
    got.backward()
    assert abs(float(got) - float(want)) / float(want) < LOSS_TOL
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        assert rel(a.grad, b.grad) < GRAD_TOL, n


@pytest.mark.parametrize("H,route", [(256, None), (512, None), (512, "fused"), (1024, None)])   # fused kernels / kept P' / fused at 512 / streamed products beyond one tile
def test_bf16_input_variant_stated_tolerance(H, route, monkeypatch):
    from transformer_transducer_b200 import functional as Fn
    monkeypatch.setattr(Fn, "ROUTE", route)
    case = _espnet_case(2, 40, 8, 500, 64, H, [40, 31], [8, 5], seed=9)
    errs, (_, _, got) = _run_pair(*case, dtype=torch.bfloat16)
    _check(errs, BF16_LOSS_TOL, BF16_GRAD_TOL)


def test_c_abi_rejects_unsupported_width_with_message():
    lib = _lib.get()
    x = torch.zeros(16, device=DEV)
    p = lambda t: t.data_ptr()  # noqa: E731
    rc = lib.ttx_joint_lse_fwd(p(x), p(x), p(x), p(x), p(x), p(x), 1, 1024, 10, 0, 0, p(x), p(x), p(x), 0, None)
    assert rc == 1 and b"not supported" in lib.ttx_last_error()


@pytest.mark.parametrize("variant", ["wide", "wide-chunked", "fused", "fused-no-replay", "fused-separate-dA", "generic",
                                     "chunked"])
def test_kernel_variants_agree_with_oracle(variant, monkeypatch):
    """Every route of H = 512 meets the tolerance on a batch with an odd number of lattice tiles (the pad tile of the last
    pair) and a negative grad_output: the default (streamed products around the kept P'); the same with a budget of one
    tile pair (several chunks, P' recomputed in the backward); the fused recomputing kernels with their bounded P' replay
    workspace, without it, and with the separate activation-gradient launch; the generic single-CTA backward kernels
    (no transposed operand copies: what H = 64 / 192 / 384 run); the library-GEMM chunked fallback."""
    from transformer_transducer_b200 import functional as Fn
    if variant == "wide-chunked":
        monkeypatch.setenv("TTX_KEEP_GB", "1e-9")
    elif variant == "chunked":
        monkeypatch.setattr(Fn, "ROUTE", "chunked")
    elif variant != "wide":
        monkeypatch.setattr(Fn, "ROUTE", "fused")
        if variant == "fused-no-replay":
            monkeypatch.setattr(Fn.FusedJointRNNT, "REPLAY", False)
        if variant in ("fused-separate-dA", "generic"):
            monkeypatch.setattr(Fn.FusedJointRNNT, "SEPARATE_ACT_GRAD", True)
        if variant == "generic":
            monkeypatch.setattr(Fn.FusedJointRNNT, "PAIR_WIDTHS", ())
    case = _espnet_case(3, 47, 10, 1100, 64, 512, [47, 31, 6], [10, 8, 1], seed=21)     # 5 + 3 + 1 = 9 tiles (odd)
    errs, _ = _run_pair(*case, weights=torch.tensor([1.0, -0.5, 2.0]))
    _check(errs)


# ----------------------------------------------------------------------------- BASELINE.json configs at their real dims
@pytest.mark.parametrize("route", [None, "fused"])      # kept softmax numerators / bounded scratch + recompute
def test_cfg2_full_lattice_vs_oracle(route, monkeypatch):
    """configs[1] with its full lattice (T=400, U=40, V=4232, D=H=512), B=2 so the CPU oracle finishes in seconds:
    129 + 99 = 228 lattice tiles = 114 tile pairs, more than the 74 CTA pairs of a B200, so the persistent kernels'
    multi-unit loop (next unit's operands behind the previous read-out, barrier parities across units) is checked
    against the oracle, with and without the kept P' matrix."""
    from transformer_transducer_b200 import functional as Fn
    monkeypatch.setattr(Fn, "ROUTE", route)
    case = _espnet_case(2, 400, 40, 4232, 512, 512, [400, 371], [40, 33], seed=11)
    errs, (e1, p1, _) = _run_pair(*case, weights=torch.tensor([1.0, 0.5]))
    _check(errs)
    assert e1.grad[1, 371:].abs().max() == 0 and p1.grad[1, 34:].abs().max() == 0


def test_cfg4_long_utterance_real_dims_vs_float64_oracle():
    """configs[3] dims with the real joint (H=512, V=4232) on a T=1000, U=100 lattice (101 000 cells, 790 tiles), B=1
    (the float64 arbiter needs ~14 GB of host memory for it).  See test_fused_long_utterance_lattice for why the
    arbiter is the float64 oracle at this lattice size."""
    case = _espnet_case(1, 1000, 100, 4232, 512, 512, [1000], [100], seed=12)
    errs, _ = _run_pair(*case, arbiter64=True)
    _check(errs)


def test_cfg5_ragged_bf16_real_lengths_vs_oracle():
    """configs[4]: ragged batch with lengths from the config's ranges (T in [50,1000], U in [5,200]), bf16 joint inputs and
    parameters, V=4233, -1 label padding.  The oracle (fed the same bf16-rounded inputs, fp32 arithmetic) runs one
    utterance at a time on that utterance's own (T_b, U_b) crop -- the padded dense logits would be 13.6 GB."""
    T_l, U_l = [1000, 430, 50, 640], [200, 77, 5, 120]
    ref, mine, enc, pred, labels, al, ll = _espnet_case(4, 1000, 200, 4233, 512, 512, T_l, U_l, seed=13)
    wts = torch.tensor([1.0, 2.0, 0.5, 1.0])
    mine = mine.to(DEV).bfloat16()
    ref.load_state_dict({k: v.float().cpu() for k, v in mine.state_dict().items()})
    enc, pred = enc.bfloat16().float(), pred.bfloat16().float()
    want = []
    g_enc, g_pred = torch.zeros_like(enc), torch.zeros_like(pred)
    for b, (t, u) in enumerate(zip(T_l, U_l)):
        e0 = enc[b:b + 1, :t].clone().requires_grad_()
        p0 = pred[b:b + 1, :u + 1].clone().requires_grad_()
        c = rnnt_oracle.rnnt_loss(ref(e0[:, :, None], p0[:, None]), labels[b:b + 1, :u].contiguous(), _i32([t]), _i32([u]),
                                  0, "none")
        (c * wts[b]).sum().backward()
        want.append(float(c))
        g_enc[b, :t], g_pred[b, :u + 1] = e0.grad[0], p0.grad[0]
    e1 = enc.to(DEV).bfloat16().requires_grad_()
    p1 = pred.to(DEV).bfloat16().requires_grad_()
    z = mine(e1[:, :, None], p1[:, None])
    assert isinstance(z, ttb.LazyJointLogits) and z.dtype == torch.bfloat16
    got = ttb.rnnt_loss(z.to(dtype=torch.float32), labels.to(DEV), al.to(DEV), ll.to(DEV), 0, "none")   # loss.py:57-60
    (got.float() * wts.to(DEV)).sum().backward()
    torch.cuda.synchronize()
    want = torch.tensor(want, dtype=torch.float64)
    errs = {"loss": float(((got.double().cpu() - want) / want).abs().max()),
            "d_enc": rel(e1.grad, g_enc), "d_pred": rel(p1.grad, g_pred)}
    for (n, a), (_, b) in zip(mine.named_parameters(), ref.named_parameters()):
        errs["d_" + n] = rel(a.grad, b.grad)
    _check(errs, BF16_LOSS_TOL, BF16_GRAD_TOL)
    assert e1.grad[2, 50:].abs().max() == 0 and p1.grad[2, 6:].abs().max() == 0


@pytest.mark.parametrize("route", [None, "fused"])
def test_trained_like_distribution(route, monkeypatch):
    """Beyond random initialisation: heavy-tailed output weights (a few |w| far above the median), posteriors peaked
    in EVERY vocabulary tile (the row maximum sits in a late tile for most rows, so the forward's running reference
    moves after its first tile), label logits boosted like a trained model's, and grad_output spanning six orders of
    magnitude between utterances (the gmax normalisation of the 16-bit gradient operand)."""
    from transformer_transducer_b200 import functional as Fn
    monkeypatch.setattr(Fn, "ROUTE", route)
    ref, mine, enc, pred, labels, al, ll = _espnet_case(3, 48, 9, 2100, 64, 512, [48, 40, 21], [9, 7, 3], seed=14)
    g = torch.Generator().manual_seed(15)
    with torch.no_grad():
        w = ref.lin_out.weight
        w.mul_(3.0)                                               # logits with a standard deviation of several nats
        tail = torch.rand(w.shape, generator=g) < 2e-3
        w[tail] *= 12.0                                           # weight outliers
        for t0 in range(0, 2100, 256):                            # one strongly preferred unit per 256-wide tile
            ref.lin_out.bias[min(2099, t0 + int(torch.randint(0, 200, (1,), generator=g)))] += 6.0
        ref.lin_out.bias[0] += 4.0                                # blank
        ref.lin_out.bias[labels[labels > 0].long().unique()] += 5.0
    mine.load_state_dict(ref.state_dict())
    errs, _ = _run_pair(ref, mine, enc, pred, labels, al, ll, weights=torch.tensor([1e-3, 1.0, 1e3]))
    _check(errs)


# ----------------------------------------------------------------------------- full-size properties (configs[1])
def test_full_size_properties_cfg2():
    """B=32, T=400, U=40, V=4232, D=H=512 -- too big for the CPU oracle, so size-independent properties:
    fused == dense-entry on a sub-batch, duplicate utterances give identical costs, gradients are linear in
    grad_output, alpha- and beta-side log-likelihoods agree, padding gets exact zeros."""
    torch.manual_seed(10)
    B, T, U, V, D, H = 32, 400, 40, 4232, 512, 512
    joint = ttb.JointNetwork(V, D, D, H, "tanh").to(DEV)
    enc, pred = torch.randn(B, T, D, device=DEV), torch.randn(B, U + 1, D, device=DEV)
    enc[1], pred[1] = enc[0], pred[0]
    labels = torch.randint(1, V, (B, U), dtype=torch.int32, device=DEV)
    labels[1] = labels[0]
    al = torch.full((B,), T, dtype=torch.int32, device=DEV)
    ll = torch.full((B,), U, dtype=torch.int32, device=DEV)
    al[2], ll[2] = 123, 17

    def run(scale):
        e, p_ = enc.clone().requires_grad_(), pred.clone().requires_grad_()
        joint.zero_grad()
        c = ttb.rnnt_loss(joint(e[:, :, None], p_[:, None]), labels, al, ll, 0, "none")
        (c * scale).sum().backward()
        return c.detach(), e.grad, p_.grad, joint.lin_out.weight.grad.clone()

    c1, ge1, gp1, gw1 = run(1.0)
    c2, ge2, gp2, gw2 = run(0.25)
    assert torch.isfinite(c1).all() and float(c1.min()) > 0
    assert float(c1[0]) == float(c1[1])
    assert torch.equal(c1, c2)
    assert rel(ge2, 0.25 * ge1) < 1e-5 and rel(gw2, 0.25 * gw1) < 1e-4
    assert ge1[2, 123:].abs().max() == 0 and gp1[2, 18:].abs().max() == 0
    # sub-batch through the dense entry (logits materialised by torch on the GPU: 4 x 400 x 41 x 4232 fp32 = 1.1 GB)
    sub = slice(0, 4)
    e, p_ = enc[sub].clone().requires_grad_(), pred[sub].clone().requires_grad_()
    joint.zero_grad()
    z = joint.lin_out(torch.tanh(joint.lin_enc(e)[:, :, None] + joint.lin_dec(p_)[:, None]))
    cd = ttb.rnnt_loss(z, labels[sub], al[sub].clamp(max=T), ll[sub], 0, "none")
    cd.sum().backward()
    assert float(((cd - c1[sub]) / cd).abs().max()) < LOSS_TOL
    e3, p3 = enc[sub].clone().requires_grad_(), pred[sub].clone().requires_grad_()
    gw_dense = joint.lin_out.weight.grad.clone()
    joint.zero_grad()
    ttb.rnnt_loss(joint(e3[:, :, None], p3[:, None]), labels[sub], al[sub], ll[sub], 0, "none").sum().backward()
    assert rel(e3.grad, e.grad) < GRAD_TOL and rel(p3.grad, p_.grad) < GRAD_TOL
    assert rel(joint.lin_out.weight.grad, gw_dense) < GRAD_TOL


def test_random_shape_sweep_matches_oracle():
    """Sixteen random shapes (tools/parity_sweep.py, seed 0): every route's widths, V from 2 to 1025, ragged lengths down to
    T = 1 / U = 0, bf16 and fp32 inputs, grad_output weights spread over 1e-2 .. 1e2 with both signs."""
    import random
    rng = random.Random(0)
    for case in range(16):
        H = rng.choice([64, 128, 256, 384, 512, 512, 512, 1024, 1024, 1536, 768])
        V = rng.choice([2, 3, 17, 63, 64, 65, 127, 129, 255, 256, 257, 300, 511, 513, 700, 1000, 1025])
        B = rng.randint(1, 5)
        T = rng.randint(1, 70)
        U = rng.randint(0, 12)
        D = rng.choice([32, 64, 512])
        al = [rng.randint(1, T) for _ in range(B)]
        al[rng.randrange(B)] = T
        ll = [rng.randint(0, U) for _ in range(B)]
        ll[rng.randrange(B)] = U
        dtype = torch.bfloat16 if rng.random() < 0.25 else torch.float32
        wts = torch.tensor([10.0 ** rng.uniform(-2, 2) * rng.choice([1, 1, -1]) for _ in range(B)])
        errs, _ = _run_pair(*_espnet_case(B, T, U, V, D, H, al, ll, seed=case), weights=wts, dtype=dtype)
        if dtype == torch.float32:
            _check(errs)
        else:
            _check(errs, BF16_LOSS_TOL, BF16_GRAD_TOL)


def test_dense_logits_entry_random_sweep_matches_oracle():
    """warprnnt_pytorch.rnnt_loss on DENSE (B, T, U1, V) logits (the upstream calling convention, train.py:53 without the
    lazy handle) over twenty random ragged shapes, V from 2 to 4233, logit scales 0.1 .. 5, all three reductions, random
    grad_output: fp32 arithmetic -- loss 1e-5, gradient 1e-4, exact zeros outside each utterance's lattice."""
    import random
    rng = random.Random(5)
    for case in range(20):
        B, T, U = rng.randint(1, 6), rng.randint(1, 90), rng.randint(0, 15)
        V = rng.choice([2, 3, 5, 31, 32, 33, 100, 255, 256, 257, 1000, 4233])
        al = [rng.randint(1, T) for _ in range(B)]
        al[rng.randrange(B)] = T
        ll = [rng.randint(0, U) for _ in range(B)]
        ll[rng.randrange(B)] = U
        torch.manual_seed(case)
        acts = torch.randn(B, T, U + 1, V) * rng.choice([0.1, 1.0, 5.0])
        labels = torch.randint(1, V, (B, U), dtype=torch.int32)
        for i, n in enumerate(ll):
            labels[i, n:] = -1
        red = rng.choice(["none", "mean", "sum"])
        a0 = acts.clone().requires_grad_()
        want = rnnt_oracle.rnnt_loss(a0, labels, _i32(al), _i32(ll), 0, red)
        wts = torch.randn_like(want)
        (want * wts).sum().backward()
        a1 = acts.to(DEV).requires_grad_()
        got = ttb.rnnt_loss(a1, labels.to(DEV), _i32(al).to(DEV), _i32(ll).to(DEV), 0, red)
        (got * wts.to(DEV)).sum().backward()
        assert got.shape == want.shape
        err = (got.detach().cpu().double() - want.detach().double()).abs() / want.detach().double().abs().clamp_min(1e-6)
        assert float(err.max()) < 1e-5, (case, float(err.max()))
        assert rel(a1.grad, a0.grad) < 1e-4, (case, rel(a1.grad, a0.grad))
        for b in range(B):
            assert al[b] == T or float(a1.grad[b, al[b]:].abs().max()) == 0
            assert ll[b] == U or float(a1.grad[b, :, ll[b] + 1:].abs().max()) == 0


def test_autograd_usage_patterns():
    """What a training script may do around the drop-in: backward twice with retain_graph, one handle through the loss
    twice, gradient accumulation over steps, a combined objective, a non-default stream, materialising the handle."""
    torch.manual_seed(0)
    B, T, U, V, D, H = 3, 30, 6, 211, 64, 512
    joint = ttb.JointNetwork(V, D, D, H, "tanh").to(DEV)
    crit = ttb.RNNTLoss()
    enc = torch.randn(B, T, 1, D, device=DEV, requires_grad=True)
    pred = torch.randn(B, 1, U + 1, D, device=DEV, requires_grad=True)
    labels = torch.randint(1, V, (B, U), dtype=torch.int32, device=DEV)
    al, ll = _i32([T, T - 4, 9]).to(DEV), _i32([U, 3, 0]).to(DEV)

    def grads():
        g = [enc.grad.clone(), pred.grad.clone()] + [p.grad.clone() for p in joint.parameters()]
        enc.grad = pred.grad = None
        for p in joint.parameters():
            p.grad = None
        return g

    def close(a, b, tol=1e-5):
        return all(float((x - y).abs().max()) <= tol * (1 + float(y.abs().max())) for x, y in zip(a, b))

    loss = crit(joint(enc, pred), labels, al, ll)
    loss.backward()
    g0, l0 = grads(), float(loss.detach())
    loss = crit(joint(enc, pred), labels, al, ll)
    loss.backward(retain_graph=True)
    loss.backward()
    assert close(grads(), [2 * x for x in g0])
    z = joint(enc, pred)
    l1, l2 = crit(z, labels, al, ll), crit(z, labels, al, ll)
    (l1 + l2).backward()
    assert abs(float(l1.detach()) - l0) < 1e-6 * abs(l0) and close(grads(), [2 * x for x in g0])
    for _ in range(2):
        crit(joint(enc, pred), labels, al, ll).backward()
    assert close(grads(), [2 * x for x in g0])
    (0.3 * crit(joint(enc, pred), labels, al, ll) + 1e-3 * enc.pow(2).sum()).backward()
    gf = grads()
    assert close([gf[1]], [0.3 * g0[1]]) and close([gf[0]], [0.3 * g0[0] + 2e-3 * enc.detach()])
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        loss = crit(joint(enc, pred), labels, al, ll)
        loss.backward()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert abs(float(loss.detach()) - l0) < 1e-6 * abs(l0) and close(grads(), g0)
    dense = joint(enc, pred).materialize()
    assert tuple(dense.shape) == (B, T, U + 1, V) and dense.requires_grad


def test_out_of_vocabulary_label_stays_inside_the_buffers_when_the_host_check_is_skipped(monkeypatch):
    """ttx_joint_act stores a label >= V as 'none' (include/ttx.h), so a caller that bypasses certify_inputs
    (TTX_SKIP_LENGTH_CHECKS=1, or the C ABI directly) gets finite numbers and no out-of-bounds gather / scatter; with the
    check in place the same input raises ValueError."""
    case = _espnet_case(2, 20, 5, 150, 64, 512, [20, 13], [5, 2], seed=11)
    _, mine, enc, pred, labels, al, ll = case
    labels = labels.clone()
    labels[0, 2] = 150 + 77
    mine = mine.to(DEV)
    e1, p1 = enc.to(DEV).requires_grad_(), pred.to(DEV).requires_grad_()
    with pytest.raises(ValueError):
        ttb.rnnt_loss(mine(e1[:, :, None], p1[:, None]), labels.to(DEV), al.to(DEV), ll.to(DEV))
    monkeypatch.setenv("TTX_SKIP_LENGTH_CHECKS", "1")
    guard = torch.zeros(4096, device=DEV)                   # (something for a stray scatter to hit)
    costs = ttb.rnnt_loss(mine(e1[:, :, None], p1[:, None]), labels.to(DEV), al.to(DEV), ll.to(DEV), 0, "none")
    costs.sum().backward()
    torch.cuda.synchronize()
    assert torch.isfinite(costs).all() and torch.isfinite(mine.lin_out.weight.grad).all()
    assert torch.isfinite(e1.grad).all() and float(guard.abs().max()) == 0
