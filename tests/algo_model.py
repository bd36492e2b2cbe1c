"""Pure-torch model of the arithmetic the CUDA kernels perform (test helper, CPU, float64 carrier).

It mirrors the kernel decomposition documented in DESIGN.md so each kernel stage can be compared
against an intermediate here, and so the rounding design (fp16 tensor-core operands, fp32
accumulation, power-of-two scales) can be checked against the oracle on CPU:

  A16   = fp16(tanh(Eproj[b,t] + Pproj[b,u]))                       joint_act
  W16   = fp16(W_out * w_scale)                                      cast_w
  z     = (A16 . W16^T) / w_scale + b_out ; lse, lp_blank, lp_label  joint_fwd
  alpha, beta, ll                                                    lattice
  gamma = exp(alpha+beta-ll), rb / rl = blank / label transition posteriors
  P'    = fp16(SCALE * (softmax - rb[v==blank] - rl[v==label]))      joint_bwd (dA pass)
  dA    = g_b*gamma/SCALE * (P' . W16)/w_scale
  Q     = fp16(SCALE * g_b*gamma/gmax * (softmax - ...))             joint_bwd (dW pass)
  dW    = gmax/SCALE * Q^T . A16 ; db = sum_m g_b*gamma*(softmax - ...)
"""
import math

import torch

SCALE = 4096.0


def _r16(x, enable, kind="fp16"):
    if not enable:
        return x
    if kind == "bf16":
        return x.float().bfloat16().to(x.dtype)
    return x.float().half().to(x.dtype)


def w_scale_for(W):
    wmax = float(W.abs().max())
    if wmax == 0.0 or not math.isfinite(wmax):
        return 1.0
    return 2.0 ** (9 - math.frexp(wmax)[1])  # wmax * scale in [256, 512)


def forward_backward(Eproj, Pproj, W, b, labels, act_lens, label_lens, grad_costs, blank=0, emulate=True,
                     kind="fp16"):
    """All inputs torch CPU; returns dict of float64 intermediates and gradients."""
    dt = torch.float64
    E, P, W, b = Eproj.to(dt), Pproj.to(dt), W.to(dt), b.to(dt)
    B, T, H = E.shape
    U1 = P.shape[1]
    V = W.shape[0]
    scale = SCALE if (emulate and kind == "fp16") else 1.0
    ws = w_scale_for(W) if (emulate and kind == "fp16") else 1.0
    A = torch.tanh(E[:, :, None, :] + P[:, None, :, :])
    A16 = _r16(A, emulate, kind)
    W16 = _r16(W * ws, emulate, kind)
    z = (A16 @ W16.T) / ws + b
    lse = torch.logsumexp(z, -1)
    lab = torch.full((B, U1), blank, dtype=torch.long)
    valid_lab = torch.zeros(B, U1, dtype=torch.bool)
    for i in range(B):
        n = int(label_lens[i])
        lab[i, :n] = labels[i, :n].long()
        valid_lab[i, :n] = True
    lpb = z[..., blank] - lse
    lpl = z.gather(3, lab.view(B, 1, U1, 1).expand(B, T, U1, 1)).squeeze(3) - lse
    ninf = float("-inf")
    alpha = torch.full((B, T, U1), ninf, dtype=dt)
    beta = torch.full((B, T, U1), ninf, dtype=dt)
    ll = torch.zeros(B, dtype=dt)
    for i in range(B):
        Tb, Ub = int(act_lens[i]), int(label_lens[i])
        for t in range(Tb):
            for u in range(Ub + 1):
                if t == 0 and u == 0:
                    v = 0.0
                elif u == 0:
                    v = alpha[i, t - 1, 0] + lpb[i, t - 1, 0]
                elif t == 0:
                    v = alpha[i, 0, u - 1] + lpl[i, 0, u - 1]
                else:
                    v = torch.logaddexp(alpha[i, t - 1, u] + lpb[i, t - 1, u], alpha[i, t, u - 1] + lpl[i, t, u - 1])
                alpha[i, t, u] = v
        for t in range(Tb - 1, -1, -1):
            for u in range(Ub, -1, -1):
                if t == Tb - 1 and u == Ub:
                    v = lpb[i, t, u]
                elif u == Ub:
                    v = beta[i, t + 1, u] + lpb[i, t, u]
                elif t == Tb - 1:
                    v = beta[i, t, u + 1] + lpl[i, t, u]
                else:
                    v = torch.logaddexp(beta[i, t + 1, u] + lpb[i, t, u], beta[i, t, u + 1] + lpl[i, t, u])
                beta[i, t, u] = v
        ll[i] = alpha[i, Tb - 1, Ub] + lpb[i, Tb - 1, Ub]
    costs = -ll
    gamma = torch.zeros(B, T, U1, dtype=dt)
    rb = torch.zeros(B, T, U1, dtype=dt)
    rl = torch.zeros(B, T, U1, dtype=dt)
    for i in range(B):
        Tb, Ub = int(act_lens[i]), int(label_lens[i])
        a, be = alpha[i, :Tb, : Ub + 1], beta[i, :Tb, : Ub + 1]
        gamma[i, :Tb, : Ub + 1] = torch.exp(a + be - ll[i])
        rb[i, : Tb - 1, : Ub + 1] = torch.exp(lpb[i, : Tb - 1, : Ub + 1] + be[1:] - be[:-1])
        rb[i, Tb - 1, Ub] = 1.0
        if Ub > 0:
            rl[i, :Tb, :Ub] = torch.exp(lpl[i, :Tb, :Ub] + be[:, 1:] - be[:, :-1])
    g = grad_costs.to(dt)
    gmax = float(g.abs().max()) or 1.0
    soft = torch.exp(z - lse[..., None])
    Pp = soft.clone()
    Pp[..., blank] -= rb
    Pp.scatter_add_(3, lab.view(B, 1, U1, 1).expand(B, T, U1, 1), -(rl * valid_lab[:, None, :])[..., None])
    gam_g = gamma * g.view(B, 1, 1)
    P16 = _r16(Pp * scale, emulate, kind)
    dA = gam_g[..., None] * (P16 @ W16) / (scale * ws)
    Q16 = _r16(Pp * (gam_g / gmax)[..., None] * scale, emulate, kind)
    dW = (Q16.reshape(-1, V).T @ A16.reshape(-1, H)) * (gmax / scale)
    db = (Pp * gam_g[..., None]).reshape(-1, V).sum(0)
    dpre = dA * (1.0 - torch.tanh(E[:, :, None, :] + P[:, None, :, :]) ** 2)
    return dict(A16=A16, z=z, lse=lse, lpb=lpb, lpl=lpl, alpha=alpha, beta=beta, costs=costs, gamma=gamma, rb=rb,
                rl=rl, dA=dA, dEproj=dpre.sum(2), dPproj=dpre.sum(1), dW=dW, db=db, w_scale=ws)
