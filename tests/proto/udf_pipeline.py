"""Layer-by-layer prototype of the hand-derived UDF-network forward / reverse / tangent / backward chains,
i.e. the exact sequence of GEMMs and epilogues the CUDA path (neuraludf_b200/csrc/udf_net.cu) executes.
Pure torch; checked against autograd of the oracle in tests/test_derivation.py.  Test infrastructure."""
import math

import torch
import torch.nn.functional as F

SQ = 1.0 / math.sqrt(2.0)


def pe(x, L):
    out = [x]
    for k in range(L):
        out += [torch.sin(x * 2.0 ** k), torch.cos(x * 2.0 ** k)]
    return torch.cat(out, -1)


def pe_jvp(x, v, L):
    """J_e(x) v : [P,3]->[P,E]"""
    out = [v]
    for k in range(L):
        f = 2.0 ** k
        out += [f * torch.cos(x * f) * v, -f * torch.sin(x * f) * v]
    return torch.cat(out, -1)


def pe_vjp(x, ge, L):
    """J_e(x)^T ge : [P,E]->[P,3]"""
    d = x.shape[1]
    g = ge[:, :d].clone()
    for k in range(L):
        f = 2.0 ** k
        g = g + f * (torch.cos(x * f) * ge[:, d * (1 + 2 * k): d * (2 + 2 * k)]
                     - torch.sin(x * f) * ge[:, d * (2 + 2 * k): d * (3 + 2 * k)])
    return g


def fold(g, v):
    return v * (g / v.norm(dim=1, keepdim=True))


def fold_backward(g, v, dW):
    """dW -> (dg, dv) for W = g v/||v||."""
    n = v.norm(dim=1, keepdim=True)
    vh = v / n
    dg = (dW * vh).sum(1, keepdim=True)
    dv = (g / n) * (dW - dg * vh)
    return dg, dv


def forward(p, cfg, x):
    """returns saved dict with out [P,d_out], grad [P,3] and everything backward needs."""
    n_lin = len(cfg["layers"])
    ls = cfg["skip_in"][0] if len(cfg["skip_in"]) else -1
    L, scale = cfg["multires"], cfg["scale"]
    W = [fold(p["lin%d.weight_g" % l], p["lin%d.weight_v" % l]) for l in range(n_lin)]
    b = [p["lin%d.bias" % l] for l in range(n_lin)]
    xs = x * scale
    E0 = pe(xs, L)
    A = [E0]
    for l in range(n_lin):
        Z = A[l] @ W[l].t() + b[l]
        if l < n_lin - 1:
            a = F.softplus(Z, beta=100.0)
            if l + 1 == ls:
                A.append(torch.cat([a, E0], 1) * SQ)
            else:
                A.append(a)
        else:
            Y = Z
    sgn = torch.sign(Y[:, :1])
    out = torch.cat([Y[:, :1].abs() / scale, Y[:, 1:]], 1)

    def S_of(l):  # sigma(100 z_l) from the stored activation a_l = A[l+1] (first out_l columns)
        out_l = cfg["layers"][l][1]
        a = A[l + 1][:, :out_l]
        if l + 1 == ls:
            a = a / SQ
        # torch softplus(beta=100, threshold=20): identity (derivative exactly 1) where 100 z > 20
        return torch.where(100.0 * a > 20.0, torch.ones_like(a), -torch.expm1(-100.0 * a))

    # reverse sweep
    G = (sgn / scale) * W[n_lin - 1][0:1, :]
    Gpe = torch.zeros_like(E0)
    D = [None] * (n_lin - 1)
    for l in range(n_lin - 2, -1, -1):
        out_l = cfg["layers"][l][1]
        if l + 1 == ls:
            Gpe = G[:, out_l:] * SQ
            G = G[:, :out_l] * SQ
        D[l] = G * S_of(l)
        G = D[l] @ W[l]
    Ge = G + Gpe
    grad = scale * pe_vjp(xs, Ge, L)
    return dict(out=out, grad=grad, W=W, A=A, D=D, sgn=sgn, xs=xs, S_of=S_of)


def backward(p, cfg, sv, out_bar, grad_bar):
    """param grads (dict like p) given d loss/d out [P,d_out] and d loss/d grad [P,3]."""
    n_lin = len(cfg["layers"])
    ls = cfg["skip_in"][0] if len(cfg["skip_in"]) else -1
    L, scale = cfg["multires"], cfg["scale"]
    W, A, D, sgn, xs, S_of = sv["W"], sv["A"], sv["D"], sv["sgn"], sv["xs"], sv["S_of"]
    dW = [torch.zeros_like(w) for w in W]
    db = [torch.zeros(w.shape[0], dtype=w.dtype) for w in W]
    # tangent chain
    Edot = scale * pe_jvp(xs, grad_bar, L)
    Adot = Edot
    Q = [None] * (n_lin - 1)
    for l in range(n_lin - 1):
        Zdot = Adot @ W[l].t()
        dW[l] += D[l].t() @ Adot
        S = S_of(l)
        Q[l] = Zdot * D[l] * (100.0 * (1.0 - S))
        Adot = S * Zdot
        if l + 1 == ls:
            Adot = torch.cat([Adot, Edot], 1) * SQ
    dW[n_lin - 1][0] += ((sgn / scale) * Adot).sum(0)
    # backward chain
    Zbar = torch.cat([sgn * out_bar[:, :1] / scale, out_bar[:, 1:]], 1)
    dW[n_lin - 1] += Zbar.t() @ A[n_lin - 1]
    db[n_lin - 1] += Zbar.sum(0)
    Abar = Zbar @ W[n_lin - 1]
    for l in range(n_lin - 2, -1, -1):
        out_l = cfg["layers"][l][1]
        if l + 1 == ls:
            Abar = Abar[:, :out_l] * SQ
        Zbar = Abar * S_of(l) + Q[l]
        dW[l] += Zbar.t() @ A[l]
        db[l] += Zbar.sum(0)
        if l > 0:
            Abar = Zbar @ W[l]
    grads = {}
    for l in range(n_lin):
        dg, dv = fold_backward(p["lin%d.weight_g" % l], p["lin%d.weight_v" % l], dW[l])
        grads["lin%d.weight_g" % l] = dg
        grads["lin%d.weight_v" % l] = dv
        grads["lin%d.bias" % l] = db[l]
    return grads
