"""NumPy statement of the BLOCKED LINEAR algorithm the `lin` kernel family (chain_control_b200/csrc/cc_lin_*.cuh)
runs for depth-0 linear models (f, g single Linear layers, identity final activation, time-invariant):
TEST INFRASTRUCTURE, checked against the oracle in tests/test_linblock_algebra.py.

A linear right-hand side f(z, u~) = A z + B u~ + c makes one integrator step (integrate.py:1-28) an affine map
    x+ = Phi x + Gam (B u~ + c),   RK4: Phi = sum_{n<=4} (hA)^n / n!,  Gam = h sum_{n<=3} (hA)^n / (n+1)!
(Euler: Phi = I + hA, Gam = hI; no-integrate: Phi = A, Gam = I).  m steps collapse into ONE matrix product
    [p' ; yfree_0 .. yfree_{m-1}] = Wf [p ; u~_prev(0..m-1) ; e]
with p the block-start state without the previous block's input response (x_k = p + C u~_prev + c_m e), so the only
per-step sequential work left in closed loop is scalar: y_j = yfree_j + const_j + sum_i H_{j+q-1-i} u~_i (Markov
parameters H_n = G Phi^n Gam_u) and the controller.  The reverse pass is the transposed recursion, and the gradient of a
TRAINABLE linear model follows from Q = sum over blocks and trajectories of Lam (x) X  (Lam = [mu_E ; w_0..w_{m-1}],
X = [x_k ; u~_0..u~_{m-1} ; 1]) by a fixed post-processing (matrix-polynomial derivative).
"""
import numpy as np

from oracle import np_oracle as O


def lin_parts(spec, theta):
    """A[S,S], B[S,I], c[S], G[O,S] (out_scale folded), d[O] in fp64 from the flat theta."""
    th = np.asarray(theta, np.float64)
    (fl, gl) = O.split_params(spec, th)
    assert len(fl) == 1 and len(gl) == 1 and spec.f.act_final == O.ACT_IDENTITY and spec.g.act_final == O.ACT_IDENTITY
    assert not spec.f_time_variant and not spec.g_time_variant and not spec.g_feedthrough_u
    S, I = spec.state_dim, spec.input_dim
    W, b = fl[0]
    A, B = W[:, :S], W[:, S:S + I]
    c = np.zeros(S) if b is None else b
    Gw, gb = gl[0]
    os_ = float(spec.out_scale)
    G = os_ * Gw
    d = os_ * (np.zeros(spec.output_dim) if gb is None else gb)
    return A, B, c, G, d


def phi_gam(spec, A):
    """(Phi, Gam, poly) with poly = coefficients phi_n of Phi = sum phi_n (hA)^n (None for no-integrate)."""
    S = A.shape[0]
    h = float(np.float32(spec.dt))
    Id = np.eye(S)
    if spec.integrator == O.INT_RK4:
        hA = h * A
        P = [Id, hA, hA @ hA, hA @ hA @ hA, hA @ hA @ hA @ hA]
        Phi = P[0] + P[1] + P[2] / 2 + P[3] / 6 + P[4] / 24
        Gam = h * (P[0] + P[1] / 2 + P[2] / 6 + P[3] / 24)
        return Phi, Gam, [1.0, 1.0, 0.5, 1 / 6, 1 / 24]
    if spec.integrator == O.INT_EE:
        return Id + h * A, h * Id, [1.0, 1.0]
    return A.copy(), Id, None


class LinBlock:
    """all block matrices for block length m."""

    def __init__(self, spec, theta, m):
        self.spec, self.m = spec, m
        S, I, Od = spec.state_dim, spec.input_dim, spec.output_dim
        self.S, self.I, self.O = S, I, Od
        self.q = q = 0 if spec.readout_pre_step else 1
        A, B, c, G, d = lin_parts(spec, theta)
        self.A, self.B, self.c, self.G, self.d = A, B, c, G, d
        Phi, Gam, self.poly = phi_gam(spec, A)
        self.Phi, self.Gam = Phi, Gam
        Gu, gam = Gam @ B, Gam @ c
        self.Gu, self.gam = Gu, gam
        pw = [np.eye(S)]
        for _ in range(m + 1):
            pw.append(pw[-1] @ Phi)
        self.pw = pw
        # Markov parameters H_n = G Phi^n Gam_u  [O, I], n = 0..m
        self.H = [G @ pw[n] @ Gu for n in range(m + 1)]
        # C = [Phi^{m-1} Gu, ..., Phi^0 Gu]  (S x mI), c_m = sum_{i<m} Phi^i gam
        self.C = np.concatenate([pw[m - 1 - i] @ Gu for i in range(m)], 1)
        self.cm = sum(pw[i] @ gam for i in range(m))
        # readout constants const_j = d + G sum_{i<j+q} Phi^i gam
        self.const = [d + G @ sum((pw[i] @ gam for i in range(j + q)), np.zeros(S)) for j in range(m)]
        rows = np.concatenate([pw[m]] + [G @ pw[j + q] for j in range(m)], 0)  # (S + mO) x S
        self.Wf = rows @ np.concatenate([np.eye(S), self.C, self.cm[:, None]], 1)  # (S+mO) x (S+mI+1)
        # reverse: Chat columns (i*O+o) = (Phi^T)^{i+q} G^T[:, o]
        self.Chat = np.concatenate([pw[i + q].T @ G.T for i in range(m)], 1)  # S x mO
        rrows = np.concatenate([pw[m].T] + [Gu.T @ pw[m - 1 - j].T for j in range(m)], 0)  # (S + mI) x S
        self.Wb = rrows @ np.concatenate([np.eye(S), self.Chat], 1)  # (S+mI) x (S+mO)

    # ---- in-block maps used by the trainable-model gradient: x_{k+j} = R_j X, mu_{k+j+1} = L_j Lam ----
    def R(self, j):
        S, I, m = self.S, self.I, self.m
        R = np.zeros((S, S + m * I + 1))
        R[:, :S] = self.pw[j]
        for i in range(j):
            R[:, S + i * I:S + (i + 1) * I] = self.pw[j - 1 - i] @ self.Gu
            R[:, -1] += self.pw[i] @ self.gam
        return R

    def L(self, j):
        S, Od, m, q = self.S, self.O, self.m, self.q
        L = np.zeros((S, S + m * Od))
        L[:, :S] = self.pw[m - j - 1].T
        for i in range(j + 1 - q, m):
            L[:, S + i * Od:S + (i + 1) * Od] = self.pw[i - j - 1 + q].T @ self.G.T
        return L


def closedloop_fwd(cspec, mspec, theta_c, theta_m, refs, m, noise=None):
    """blocked closed-loop forward; the controller is evaluated by the oracle's node_step (generic)."""
    lb = LinBlock(mspec, theta_m, m)
    S, I, Od, q = lb.S, lb.I, lb.O, lb.q
    refs = np.asarray(refs, np.float64)
    B, Tr, R = refs.shape
    cfl, cgl = O.split_params(cspec, np.asarray(theta_c, np.float64))
    xc = np.zeros((B, cspec.state_dim))
    tc = 0.0
    yprev = np.broadcast_to(lb.d, (B, Od)).copy()  # y0 = g(x0 = 0)
    Avec = np.zeros((B, S + m * I + 1))  # [p ; u_prev ; e]: first block p = x0 = 0, e = 0
    yhat = np.empty((B, Tr, Od))
    us = np.empty((B, Tr, I))
    tape = []
    for k0 in range(0, Tr, m):
        D = Avec @ lb.Wf.T
        acc = [D[:, S + j * Od:S + (j + 1) * Od] + lb.const[j] for j in range(m)]
        ublk = np.zeros((B, m * I))
        for j in range(min(m, Tr - k0)):
            k = k0 + j
            v = np.concatenate([refs[:, k], yprev], 1)
            tape.append((xc.copy(), yprev.copy()))
            xc, a, _ = O.node_step(cspec, cfl, cgl, xc, tc, v)
            ut = O._ut(mspec.u_transform, mspec.u_scale, a)
            ublk[:, j * I:(j + 1) * I] = ut
            for jj in range(j + 1 - q, m):  # y_jj += H_{jj+q-1-j} u~_j
                acc[jj] = acc[jj] + ut @ lb.H[jj + q - 1 - j].T
            y = acc[j]
            if noise is not None:
                y = y + noise[:, k]
            yhat[:, k], us[:, k] = y, a
            yprev = y
            if cspec.has_time:
                tc = tc + float(np.float32(cspec.dt))
        Avec = np.concatenate([D[:, :S], ublk, np.ones((B, 1))], 1)
    return yhat, us, tape


def closedloop_bwd(cspec, mspec, theta_c, theta_m, refs, ybar, m, tape):
    """blocked reverse pass of the closed loop (frozen linear model): theta_c_bar."""
    lb = LinBlock(mspec, theta_m, m)
    S, I, Od, q = lb.S, lb.I, lb.O, lb.q
    refs = np.asarray(refs, np.float64)
    ybar = np.asarray(ybar, np.float64)
    B, Tr, R = refs.shape
    cfl, cgl = O.split_params(cspec, np.asarray(theta_c, np.float64))
    fg, gg = O._zero_grads(cfl), O._zero_grads(cgl)
    lamc = np.zeros((B, cspec.state_dim))
    yfb = np.zeros((B, Od))
    Avec = np.zeros((B, S + m * Od))  # [r ; w_later]
    nblk = (Tr + m - 1) // m
    for b in range(nblk - 1, -1, -1):
        k0 = b * m
        D = Avec @ lb.Wb.T
        ufree = [D[:, S + j * I:S + (j + 1) * I] for j in range(m)]
        w = np.zeros((B, m * Od))
        for j in range(min(m, Tr - k0) - 1, -1, -1):
            k = k0 + j
            wk = ybar[:, k] + yfb
            w[:, j * Od:(j + 1) * Od] = wk
            utb = ufree[j].copy()
            for i in range(j + 1 - q, m):
                utb = utb + w[:, i * Od:(i + 1) * Od] @ lb.H[i - j - 1 + q]
            xc, yprev = tape[k]
            tc = float(np.float32(cspec.dt)) * k if cspec.has_time else 0.0
            v = np.concatenate([refs[:, k], yprev], 1)
            _, a, cc_ = O.node_step(cspec, cfl, cgl, xc, tc, v)
            abar = utb * O._ut_grad(mspec.u_transform, mspec.u_scale, a)
            lamc, vbar = O.node_step_bwd(cspec, cfl, cgl, cc_, abar, lamc, fg, gg)
            yfb = vbar[:, R:]
        Avec = np.concatenate([D[:, :S], w], 1)
    return O._flatten_grads(fg, gg)


def rollout_fwd(spec, theta, u, m, x0=None):
    """blocked open-loop forward: y[B,T+1,O] and the block-start tape X_b = [x_k ; u~ block ; 1]."""
    lb = LinBlock(spec, theta, m)
    S, I, Od, q = lb.S, lb.I, lb.O, lb.q
    u = np.asarray(u, np.float64)
    B, T, _ = u.shape
    x = np.zeros((B, S)) if x0 is None else np.asarray(x0, np.float64)
    y = np.empty((B, T + 1, Od))
    y[:, 0] = x @ lb.G.T + lb.d
    Avec = np.concatenate([x, np.zeros((B, m * I + 1))], 1)
    ut_all = O._ut(spec.u_transform, spec.u_scale, u)
    xstart = []
    for k0 in range(0, T, m):
        D = Avec @ lb.Wf.T
        n = min(m, T - k0)
        ublk = np.zeros((B, m * I))
        ublk[:, :n * I] = ut_all[:, k0:k0 + n].reshape(B, n * I)
        # block-start state (the tape of the trainable reverse pass)
        xstart.append(Avec[:, :S] + Avec[:, S:S + m * I] @ lb.C.T + Avec[:, -1:] * lb.cm)
        for j in range(n):
            yj = D[:, S + j * Od:S + (j + 1) * Od] + lb.const[j]
            for i in range(j + q):
                yj = yj + ublk[:, i * I:(i + 1) * I] @ lb.H[j + q - 1 - i].T
            y[:, k0 + j + 1] = yj
        Avec = np.concatenate([D[:, :S], ublk, np.ones((B, 1))], 1)
    return y, xstart


def matpoly_grad(lb, Mx, Mu, M1):
    """(Abar, Bbar, cbar) from M = sum mu+ (x) [x ; u~ ; 1]: derivative of x+ = Phi(A) x + Gam(A)(B u~ + c)."""
    S = lb.S
    h = float(np.float32(lb.spec.dt))
    A, B, c = lb.A, lb.B, lb.c
    if lb.poly is None:  # no-integrate: x+ = A x + B u + c
        return Mx, Mu, M1
    At = A.T
    pw = [np.eye(S)]
    for _ in range(len(lb.poly)):
        pw.append(pw[-1] @ At)
    Mbeta = Mu @ B.T + np.outer(M1, c)
    Abar = np.zeros((S, S))
    for n in range(1, len(lb.poly)):
        cn = lb.poly[n] * h ** n
        for a in range(n):
            Abar += cn * pw[a] @ Mx @ pw[n - 1 - a]
    # Gam = h sum_{n>=0} poly[n+1] (hA)^n
    for n in range(1, len(lb.poly) - 1):
        cn = h * lb.poly[n + 1] * h ** n
        for a in range(n):
            Abar += cn * pw[a] @ Mbeta @ pw[n - 1 - a]
    return Abar, lb.Gam.T @ Mu, lb.Gam.T @ M1


def rollout_bwd(spec, theta, u, ybar, m, x0=None):
    """blocked reverse pass of the open loop with a TRAINABLE linear model: theta_bar via Q = sum Lam (x) X, and ubar."""
    lb = LinBlock(spec, theta, m)
    S, I, Od, q = lb.S, lb.I, lb.O, lb.q
    u = np.asarray(u, np.float64)
    ybar = np.asarray(ybar, np.float64)
    B, T, _ = u.shape
    _, xstart = rollout_fwd(spec, theta, u, m, x0)
    ut_all = O._ut(spec.u_transform, spec.u_scale, u)
    Kl, Kx = S + m * Od, S + m * I + 1
    Q = np.zeros((Kl, Kx))
    Avec = np.zeros((B, Kl))
    ubar = np.zeros_like(u)
    nblk = (T + m - 1) // m
    for b in range(nblk - 1, -1, -1):
        k0 = b * m
        n = min(m, T - k0)
        D = Avec @ lb.Wb.T
        w = np.zeros((B, m * Od))
        for j in range(n):  # y_{k+1} is the output of step k (row k+1 of ybar)
            w[:, j * Od:(j + 1) * Od] = ybar[:, k0 + j + 1]
        muE = Avec[:, :S] + Avec[:, S:] @ lb.Chat.T
        Lam = np.concatenate([muE, w], 1)
        ublk = np.zeros((B, m * I))
        ublk[:, :n * I] = ut_all[:, k0:k0 + n].reshape(B, n * I)
        X = np.concatenate([xstart[b], ublk, np.ones((B, 1))], 1)
        Q += Lam.T @ X
        for j in range(n):
            utb = D[:, S + j * I:S + (j + 1) * I].copy()
            for i in range(j + 1 - q, m):
                utb = utb + w[:, i * Od:(i + 1) * Od] @ lb.H[i - j - 1 + q]
            ubar[:, k0 + j] = utb * O._ut_grad(spec.u_transform, spec.u_scale, u[:, k0 + j])
        Avec = np.concatenate([D[:, :S], w], 1)
    # ---- post-processing: M = sum_j L_j Q R'_j^T ; Gbar = sum_j E_j Q R_{j+q}^T ; dbar = sum_j E_j Q e_1 ----
    Mx, Mu, M1 = np.zeros((S, S)), np.zeros((S, I)), np.zeros(S)
    Gbar, dbar = np.zeros((Od, S)), np.zeros(Od)
    for j in range(m):
        LQ = lb.L(j) @ Q  # S x Kx
        Mx += LQ @ lb.R(j).T
        Mu += LQ[:, S + j * I:S + (j + 1) * I]
        M1 += LQ[:, -1]
        Ej = Q[S + j * Od:S + (j + 1) * Od]  # O x Kx
        # x_{k+j+q}: R_{j+q} (R_m for the post-step readout of the block's last step)
        Rq = lb.R(j + q) if j + q < m else _R_full(lb)
        Gbar += Ej @ Rq.T
        dbar += Ej[:, -1]
    # y0 = g(x0) (only g sees it)
    x0v = np.zeros((B, S)) if x0 is None else np.asarray(x0, np.float64)
    Gbar += ybar[:, 0].T @ x0v
    dbar += ybar[:, 0].sum(0)
    Abar, Bbar, cbar = matpoly_grad(lb, Mx, Mu, M1)
    os_ = float(spec.out_scale)
    fW = np.concatenate([Abar, Bbar], 1)
    out = [fW.ravel()]
    if spec.f.use_bias:
        out.append(cbar)
    out.append((os_ * Gbar).ravel())
    if spec.g.use_bias:
        out.append(os_ * dbar)
    return np.concatenate(out), ubar, Q


def _R_full(lb):
    """R_m: x_{k+m} = Phi^m x_k + C u~ + c_m."""
    return np.concatenate([lb.pw[lb.m], lb.C, lb.cm[:, None]], 1)
