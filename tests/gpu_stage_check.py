"""Stage-by-stage check of the C-ABI kernels on a GPU against tests/algo_model.py (float64 CPU model).

    python tests/gpu_stage_check.py <stage> [B T U V H]      stage in: small fwd fg bwd

Run by hand / from gpurun while bringing kernels up; the pytest suite (tests/test_gpu_*.py) is the gate.
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import algo_model  # noqa: E402
from transformer_transducer_b200 import _lib  # noqa: E402
from transformer_transducer_b200.functional import _p, _stream  # noqa: E402


def make_case(B, T, U, V, H, seed=0, ragged=True):
    g = torch.Generator().manual_seed(seed)
    E = torch.randn(B, T, H, generator=g) * 0.6
    P = torch.randn(B, U + 1, H, generator=g) * 0.6
    W = (torch.rand(V, H, generator=g) * 2 - 1) / H ** 0.5
    b = (torch.rand(V, generator=g) * 2 - 1) / H ** 0.5
    labels = torch.randint(1, V, (B, U), generator=g, dtype=torch.int32)
    act_lens = torch.full((B,), T, dtype=torch.int32)
    label_lens = torch.full((B,), U, dtype=torch.int32)
    if ragged and B > 1:
        for i in range(1, B):
            act_lens[i] = max(1, T - 3 * i)
            label_lens[i] = max(0, U - 2 * i)
            labels[i, int(label_lens[i]):] = -1
    gc = torch.linspace(0.25, 1.0, B)
    return E, P, W, b, labels, act_lens, label_lens, gc


def compact(meta, model_t, act_lens, label_lens, rows):
    """(B,T,U1) float64 model tensor -> compact row vector (nan where unused)."""
    out = torch.full((rows,), float("nan"), dtype=torch.float64)
    B = model_t.shape[0]
    for i in range(B):
        Tb, U1b = int(act_lens[i]), int(label_lens[i]) + 1
        base = int(meta[4 + i]) * 128
        out[base: base + Tb * U1b] = model_t[i, :Tb, :U1b].reshape(-1)
    return out


def report(name, got, want, tol):
    mask = ~torch.isnan(want)
    g, w = got.double().cpu()[mask], want[mask]
    err = (g - w).abs().max().item() if w.numel() else 0.0
    rel = ((g - w).norm() / (w.norm() + 1e-30)).item() if w.numel() else 0.0
    bad = not (rel <= tol) or torch.isnan(g).any().item()
    print("%-10s max|d|=%.3e relL2=%.3e (tol %.1e) %s" % (name, err, rel, tol, "FAIL" if bad else "ok"), flush=True)
    return bad


def run(stage, B, T, U, V, H):
    dev = torch.device("cuda:0")
    lib = _lib.get()
    E, P, W, b, labels, al, ll, gc = make_case(B, T, U, V, H)
    U1 = U + 1
    m = algo_model.forward_backward(E, P, W, b, labels, al, ll, gc, emulate=True)
    st = _stream(dev)
    ntub = int(lib.ttx_tiles_upper_bound(B, T, U1))
    rows = ntub * 128
    meta = torch.empty(int(lib.ttx_meta_ints(B, ntub)), dtype=torch.int32, device=dev)
    ald, lld, labd = al.to(dev), ll.to(dev), labels.to(dev).contiguous()
    _lib.check(lib.ttx_prepare(_p(ald), _p(lld), B, T, U1, ntub, _p(meta), 0, st), "prepare")
    torch.cuda.synchronize()
    mh = meta.cpu()
    print("tiles in use %d of %d, status %d, cells %d" % (mh[0], ntub, mh[1], mh[2]), flush=True)
    fail = False
    f32 = lambda n: torch.empty(n, dtype=torch.float32, device=dev)  # noqa: E731
    Ed, Pd, Wd, bd = E.to(dev), P.to(dev), W.to(dev).contiguous(), b.to(dev)
    Vpad = (V + 255) // 256 * 256
    scal = torch.zeros(8, dtype=torch.float32, device=dev)
    w16 = torch.empty(Vpad * H, dtype=torch.int16, device=dev)
    bias2 = torch.empty(Vpad, dtype=torch.float32, device=dev)
    use_t = H in (128, 256, 512) and os.environ.get("TTX_STAGE_TRANSPOSE", "fused") == "fused"
    w16t = torch.empty(H * Vpad, dtype=torch.int16, device=dev) if use_t else None
    _lib.check(lib.ttx_cast_weight(_p(Wd), _p(bd), V, H, 0, _p(scal), _p(w16), _p(bias2), _p(w16t), 0, st), "cast")
    a16 = torch.empty(rows * H, dtype=torch.int16, device=dev)
    row_label = torch.empty(rows, dtype=torch.int32, device=dev)
    a16t = torch.empty(H * rows, dtype=torch.int16, device=dev) if use_t else None
    _lib.check(lib.ttx_joint_act(_p(Ed), _p(Pd), _p(labd), _p(ald), _p(lld), _p(meta), B, T, U1, H, U, V, ntub, 0,
                                 _p(a16), _p(row_label), _p(a16t), 0, st), "act")
    torch.cuda.synchronize()
    ws = float(scal[0])
    print("w_scale %g (model %g)" % (ws, m["w_scale"]), flush=True)
    w16f = w16.view(torch.float16).view(Vpad, H)[:V].double().cpu() / ws
    fail |= report("W16", w16f.reshape(-1), (algo_model._r16(W.double() * ws, True) / ws).reshape(-1), 1e-7)
    a16f = a16.view(torch.float16).view(rows, H).double().cpu()
    A_model = torch.full((rows, H), float("nan"), dtype=torch.float64)
    for i in range(B):
        Tb, U1b = int(al[i]), int(ll[i]) + 1
        base = int(mh[4 + i]) * 128
        A_model[base: base + Tb * U1b] = m["A16"][i, :Tb, :U1b].reshape(-1, H)
    fail |= report("A16", a16f.reshape(-1), A_model.reshape(-1), 3e-4)

    lse, lpb, lpl = f32(rows), f32(rows), f32(rows)
    if stage == "small":
        # dense path: materialise logits with torch on the GPU, run our dense kernels
        z = (torch.tanh(Ed[:, :, None] + Pd[:, None]) @ Wd.T + bd).contiguous()
        rl2 = torch.empty(rows, dtype=torch.int32, device=dev)
        _lib.check(lib.ttx_dense_lse(_p(z), _p(labd), _p(ald), _p(lld), _p(meta), B, T, U1, V, U, 0, ntub, _p(lse),
                                     _p(lpb), _p(lpl), _p(rl2), 0, st), "dense_lse")
        torch.cuda.synchronize()
        mm = algo_model.forward_backward(E, P, W, b, labels, al, ll, gc, emulate=False)
        tol = 2e-5
    elif stage == "fg":
        ew = f32(rows * H)
        ws_n = int(lib.ttx_joint_workspace_bytes(0, ntub, H, V, 0))
        ws = torch.empty(max(ws_n, 1), dtype=torch.uint8, device=dev)
        _lib.check(lib.ttx_joint_fwd_grad(_p(a16), _p(w16), _p(w16t), _p(bias2), _p(scal), _p(row_label), _p(meta), ntub,
                                          H, V, 0, 0, _p(lse), _p(lpb), _p(lpl), _p(ew), _p(ws) if ws_n else None, ws_n, 0,
                                          st), "fwd_grad")
        torch.cuda.synchronize()
        mm = m
        tol = 2e-5
        # EW model: sum_v p_v W16_v / w_scale without the blank and label columns
        soft = torch.exp(m["z"] - m["lse"][..., None])
        soft[..., 0] = 0.0
        lab_full = torch.zeros(B, U1, dtype=torch.long)
        for i in range(B):
            lab_full[i, : int(ll[i])] = labels[i, : int(ll[i])].long()
        mask = torch.zeros(B, U1, dtype=torch.bool)
        for i in range(B):
            mask[i, : int(ll[i])] = True
        idx = lab_full.view(B, 1, U1, 1).expand(B, T, U1, 1)
        keep = soft.gather(3, idx)
        soft.scatter_(3, idx, torch.where(mask.view(B, 1, U1, 1), torch.zeros_like(keep), keep))
        soft[..., 0] = 0.0
        W16m = algo_model._r16(W.double() * m["w_scale"], True) / m["w_scale"]
        EW_model = torch.full((rows, H), float("nan"), dtype=torch.float64)
        ewm = soft @ W16m
        for i in range(B):
            Tb, U1b = int(al[i]), int(ll[i]) + 1
            base = int(mh[4 + i]) * 128
            EW_model[base: base + Tb * U1b] = ewm[i, :Tb, :U1b].reshape(-1, H)
        fail |= report("EW", ew.view(rows, H).cpu().reshape(-1), EW_model.reshape(-1), 5e-4)
    else:
        _lib.check(lib.ttx_joint_lse_fwd(_p(a16), _p(w16), _p(bias2), _p(scal), _p(row_label), _p(meta), ntub, H, V, 0, 0,
                                         _p(lse), _p(lpb), _p(lpl), 0, st), "fwd")
        torch.cuda.synchronize()
        mm = m
        tol = 2e-5
    valid_lab = torch.zeros(B, T, U1, dtype=torch.bool)
    for i in range(B):
        valid_lab[i, :, : int(ll[i])] = True
    fail |= report("lse", lse, compact(mh, mm["lse"], al, ll, rows), tol)
    fail |= report("lp_blank", lpb, compact(mh, mm["lpb"], al, ll, rows), tol)
    lpl_want = torch.where(valid_lab, mm["lpl"], torch.full_like(mm["lpl"], float("nan")))
    fail |= report("lp_label", lpl, compact(mh, lpl_want, al, ll, rows), tol)

    lat = int(lib.ttx_lattice_elems_upper_bound(B, T, U1))
    alpha_d = torch.full((lat,), float("nan"), dtype=torch.float64, device=dev)
    beta_d = torch.full((lat,), float("nan"), dtype=torch.float64, device=dev)
    lat_ws = f32(4 * lat)
    costs, llb = f32(B), torch.empty(B, dtype=torch.float64, device=dev)
    _lib.check(lib.ttx_lattice_fwd_bwd(_p(lpb), _p(lpl), _p(ald), _p(lld), _p(meta), B, U1, ntub, lat, _p(lat_ws),
                                       _p(alpha_d), _p(beta_d), _p(costs), _p(llb), 0, st), "lattice")
    torch.cuda.synchronize()

    def from_diag(x, mirrored):          # diagonal-major lattice array -> the compact row space the other stages use
        out = torch.full((rows,), float("nan"), dtype=torch.float64)
        mh2, xc = meta.cpu(), x.cpu()
        for i in range(B):
            Tb, U1b = int(al[i]), int(ll[i]) + 1
            P, o, base = (U1b + 3) // 4 * 4, int(mh2[4 + B + 1 + ntub + i]), int(mh2[4 + i]) * 128
            t, u = torch.arange(Tb).view(-1, 1), torch.arange(U1b).view(1, -1)
            if mirrored:
                t, u = Tb - 1 - t, U1b - 1 - u
            out[base: base + Tb * U1b] = xc[(o + (t + u) * P + u).reshape(-1)]
        return out
    alpha, beta = from_diag(alpha_d, False), from_diag(beta_d, True)
    inf2nan = lambda x: torch.where(torch.isinf(x), torch.full_like(x, float("nan")), x)  # noqa: E731
    fail |= report("alpha", alpha, compact(mh, inf2nan(mm["alpha"]), al, ll, rows), 1e-5)
    fail |= report("beta", beta, compact(mh, inf2nan(mm["beta"]), al, ll, rows), 1e-5)
    fail |= report("costs", costs, mm["costs"], 1e-5)
    fail |= report("ll_beta", -llb, mm["costs"], 1e-5)

    gcd = gc.to(dev)
    rowmeta = f32(rows * 4)
    db = torch.zeros(V, dtype=torch.float32, device=dev)
    which = os.environ.get("TTX_BWD", "both")
    rlab = rl2 if stage == "small" else row_label
    _lib.check(lib.ttx_grad_coeffs(_p(lse), _p(lpb), _p(lpl), _p(alpha_d), _p(beta_d), _p(llb), _p(gcd), _p(scal), _p(rlab),
                                   _p(ald), _p(lld), _p(meta), B, 0, ntub, _p(rowmeta),
                                   _p(db) if (stage == "bwd" and which in ("both", "dw")) else None, 0, st), "coeffs")
    torch.cuda.synchronize()
    rm = rowmeta.view(rows, 4).cpu()
    gmax = float(scal[2])
    fail |= report("pb-rb", rm[:, 1], compact(mh, torch.exp(mm["lpb"]) - mm["rb"], al, ll, rows), 1e-4)
    fail |= report("pl-rl", rm[:, 2], compact(mh, torch.where(valid_lab, torch.exp(mm["lpl"]) - mm["rl"],
                                                              torch.full_like(mm["rl"], float("nan"))), al, ll, rows), 1e-4)
    fail |= report("gamma*g", rm[:, 3] * gmax, compact(mh, mm["gamma"] * gc.double().view(B, 1, 1), al, ll, rows), 1e-4)

    if stage == "small":
        grads = torch.empty_like(z)
        _lib.check(lib.ttx_dense_grad(_p(z), _p(rowmeta), _p(rl2), _p(scal), _p(ald), _p(lld), _p(meta), B, T, U1, V, 0,
                                      _p(grads), 0, st), "dense_grad")
        torch.cuda.synchronize()
        from oracle import rnnt_oracle
        zc = z.detach().cpu().requires_grad_()
        c = rnnt_oracle.rnnt_loss(zc, labels, al, ll, 0, "none")
        (c * gc).sum().backward()
        fail |= report("dense cost", costs, c.detach().double(), 1e-5)
        fail |= report("dense grad", grads.reshape(-1), zc.grad.double().reshape(-1), 1e-4)
        return fail
    if stage == "fwd":
        return fail
    if stage == "fg":
        dE, dP = f32(B * T * H), f32(B * U1 * H)
        _lib.check(lib.ttx_reduce_act_grad_ew(_p(ew), _p(rowmeta), _p(row_label), _p(Wd), _p(scal), 0, _p(Ed), _p(Pd),
                                              _p(ald), _p(lld), _p(meta), B, T, U1, H, _p(dE), _p(dP), 0, st), "reduce_ew")
        torch.cuda.synchronize()
        fail |= report("dEproj", dE, m["dEproj"].reshape(-1), 3e-4)
        fail |= report("dPproj", dP, m["dPproj"].reshape(-1), 3e-4)
        return fail

    d_act = f32(rows * H)
    dW = torch.zeros(V, H, dtype=torch.float32, device=dev)
    splits = int(os.environ.get("TTX_SPLITS", "2"))
    if H in (128, 256, 512) and not use_t:      # separate-transpose path (ttx_transpose16)
        w16t = torch.empty(H * Vpad, dtype=torch.int16, device=dev)
        a16t = torch.empty(H * rows, dtype=torch.int16, device=dev)
        _lib.check(lib.ttx_transpose16(_p(w16), _p(w16t), Vpad, H, None, 0, st), "transpose w")
        _lib.check(lib.ttx_transpose16(_p(a16), _p(a16t), rows, H, _p(meta), 0, st), "transpose a")
    _lib.check(lib.ttx_joint_grad(_p(a16), _p(w16), _p(a16t), _p(w16t), _p(bias2), _p(scal), _p(row_label), _p(meta), _p(rowmeta), ntub, H, V,
                                  0, 0, _p(d_act) if which in ("both", "da") else None,
                                  _p(dW) if which in ("both", "dw") else None,
                                  _p(db) if which in ("both", "dw") else None, splits, None, 0, 0, st), "joint_grad")
    torch.cuda.synchronize()
    if which in ("both", "da"):
        dA_model = torch.full((rows, H), float("nan"), dtype=torch.float64)
        for i in range(B):
            Tb, U1b = int(al[i]), int(ll[i]) + 1
            base = int(mh[4 + i]) * 128
            dA_model[base: base + Tb * U1b] = m["dA"][i, :Tb, :U1b].reshape(-1, H)
        fail |= report("dA", d_act.view(rows, H).cpu().reshape(-1), dA_model.reshape(-1), 2e-4)
        dE, dP = f32(B * T * H), f32(B * U1 * H)
        _lib.check(lib.ttx_reduce_act_grad(_p(d_act), _p(Ed), _p(Pd), _p(ald), _p(lld), _p(meta), B, T, U1, H, _p(dE),
                                           _p(dP), 0, st), "reduce")
        torch.cuda.synchronize()
        fail |= report("dEproj", dE, m["dEproj"].reshape(-1), 2e-4)
        fail |= report("dPproj", dP, m["dPproj"].reshape(-1), 2e-4)
    if which in ("both", "dw"):
        fail |= report("dW", dW.reshape(-1), m["dW"].reshape(-1), 2e-4)
        fail |= report("db", db, m["db"], 2e-4)
    return fail


if __name__ == "__main__":
    stage = sys.argv[1] if len(sys.argv) > 1 else "all"
    dims = [int(x) for x in sys.argv[2:7]] if len(sys.argv) >= 7 else [2, 40, 6, 300, 128]
    bad = run(stage, *dims)
    print("STAGE %s %s: %s" % (stage, dims, "FAILED" if bad else "PASSED"), flush=True)
    sys.exit(1 if bad else 0)
