"""Literal restatement of the reference's iLEQG (StanfordMSL/RATiLQR.jl, src/ileqg.jl) -- TEST INFRASTRUCTURE.

Purpose: an implementation that shares NOTHING with oracle/oracle.cpp or the CUDA kernels -- no header, no
Woodbury step, no Cholesky-of-M shortcut, no analytic Jacobians -- so that a misreading of ileqg.jl common to the
oracle and the kernels cannot pass the suite.  Every statement below follows the Julia source operation by
operation, in the order Julia evaluates it (`a*b*c` is `(a*b)*c`; `-H\\G` is `(-H)\\G`; `θ.*S/M` is `(θ.*S)/M`):

    simulate_dynamics (open loop)       ileqg.jl:18-38
    simulate_dynamics (closed loop)     ileqg.jl:62-87
    initialize!                         ileqg.jl:214-236
    approximate_model                   ileqg.jl:258-322   (derivatives by forward-mode AD, like ForwardDiff)
    solve_approximate_dp!               ileqg.jl:341-406
    solve_approximate_dp                ileqg.jl:412-465
    increase_μ_and_Δ!                   ileqg.jl:471-474
    line_search!                        ileqg.jl:494-592
    step! / solve!                      ileqg.jl:598-613 / 635-659

Two arithmetics, selected by the `LA` object handed to every routine:

    F64  -- numpy float64; inv / general solve / logdet through LAPACK (numpy.linalg: getri, gesv, getrf),
            isposdef through potrf.  This is what Julia's stdlib does, except that Julia solves with the
            Symmetric M by Bunch-Kaufman (sytrf) where numpy uses LU (getrf): both are backward stable, the
            results agree to cond(M)*eps.
    MP   -- mpmath, 60 significant digits; the same statements with a textbook partially pivoted LU written
            here.  At this precision the choice of factorisation is immaterial: MP results are the TRUE value
            of the reference's formulas and are what tests/golden/literal_*.npz store (rounded to float64).

The registered models (power law, unicycle, quadrotor, ...) are restated here from DESIGN.md section 1 as plain
scalar code over a generic number type; their derivatives come from the second-order forward-mode jets below,
never from hand-written Jacobians.
"""
import math

import mpmath
import numpy as np

mpmath.mp.dps = 60


class NotPosDef(Exception):
    """@assert isposdef(M) (ileqg.jl:366,440) / PosDefException"""


class DomainError(Exception):
    """Julia DomainError: negative base of a real power, logdet of a matrix with negative determinant"""


# ----------------------------------------------------------------------------------------------------------------
# the two arithmetics
# ----------------------------------------------------------------------------------------------------------------
class F64:
    name = "f64"
    dtype = np.float64

    @staticmethod
    def num(v):
        return float(v)

    @staticmethod
    def arr(a):
        return np.array(a, dtype=np.float64)

    @staticmethod
    def zeros(*shape):
        return np.zeros(shape)

    @staticmethod
    def eye(n):
        return np.eye(n)

    @staticmethod
    def inv(A):  # inv(W): LU + getri
        return np.linalg.inv(A)

    @staticmethod
    def solve(A, B):  # A \ B for a general square A: LU with partial pivoting (LinearAlgebra generic `\`)
        return np.linalg.solve(A, B)

    @staticmethod
    def isposdef(A):  # isposdef = Cholesky succeeds (reads the UPPER triangle, like cholesky(Symmetric(A)))
        U = np.triu(A)
        As = U + np.triu(A, 1).T
        if not np.all(np.isfinite(As)):
            return False
        try:
            np.linalg.cholesky(As)
            return True
        except np.linalg.LinAlgError:
            return False

    @staticmethod
    def logdet(A):  # logdet(A): LU; negative determinant -> DomainError
        sign, ld = np.linalg.slogdet(A)
        if sign < 0:
            raise DomainError("logdet of a matrix with negative determinant")
        return float(ld) if sign > 0 else -math.inf

    sqrt = staticmethod(math.sqrt)
    sin = staticmethod(math.sin)
    cos = staticmethod(math.cos)

    @staticmethod
    def isfinite(v):
        return math.isfinite(v)

    @staticmethod
    def to_f64(a):
        return np.asarray(a, dtype=np.float64)


class MP:
    name = "mp"
    dtype = object

    @staticmethod
    def num(v):
        return mpmath.mpf(v)

    @staticmethod
    def arr(a):
        a = np.asarray(a)
        out = np.empty(a.shape, dtype=object)
        for idx in np.ndindex(a.shape):
            out[idx] = mpmath.mpf(a[idx]) if not isinstance(a[idx], mpmath.mpf) else a[idx]
        return out

    @staticmethod
    def zeros(*shape):
        out = np.empty(shape, dtype=object)
        out.fill(mpmath.mpf(0))
        return out

    @staticmethod
    def eye(n):
        out = MP.zeros(n, n)
        for i in range(n):
            out[i, i] = mpmath.mpf(1)
        return out

    @staticmethod
    def _lu(A):
        """P A = L U by Gaussian elimination with partial pivoting; returns (LU packed, perm, sign)"""
        A = A.copy()
        n = A.shape[0]
        perm = list(range(n))
        sign = 1
        for k in range(n):
            p = max(range(k, n), key=lambda i: abs(A[i, k]))
            if A[p, k] == 0:
                raise ZeroDivisionError("singular matrix")
            if p != k:
                A[[k, p], :] = A[[p, k], :]
                perm[k], perm[p] = perm[p], perm[k]
                sign = -sign
            for i in range(k + 1, n):
                A[i, k] = A[i, k] / A[k, k]
                for j in range(k + 1, n):
                    A[i, j] = A[i, j] - A[i, k] * A[k, j]
        return A, perm, sign

    @staticmethod
    def solve(A, B):
        LU, perm, _ = MP._lu(A)
        n = A.shape[0]
        vec = B.ndim == 1
        Bm = B.reshape(n, -1)
        X = MP.zeros(n, Bm.shape[1])
        for c in range(Bm.shape[1]):
            y = [Bm[perm[i], c] for i in range(n)]
            for i in range(n):
                for k in range(i):
                    y[i] = y[i] - LU[i, k] * y[k]
            for i in reversed(range(n)):
                for k in range(i + 1, n):
                    y[i] = y[i] - LU[i, k] * y[k]
                y[i] = y[i] / LU[i, i]
            for i in range(n):
                X[i, c] = y[i]
        return X[:, 0] if vec else X

    @staticmethod
    def inv(A):
        return MP.solve(A, MP.eye(A.shape[0]))

    @staticmethod
    def isposdef(A):  # Cholesky on the upper triangle; fails on a non-positive pivot
        n = A.shape[0]
        U = MP.zeros(n, n)
        for j in range(n):
            d = A[j, j]
            for k in range(j):
                d = d - U[k, j] * U[k, j]
            if not (d > 0):
                return False
            U[j, j] = mpmath.sqrt(d)
            for i in range(j + 1, n):
                a = A[j, i]
                for k in range(j):
                    a = a - U[k, j] * U[k, i]
                U[j, i] = a / U[j, j]
        return True

    @staticmethod
    def logdet(A):
        LU, _, sign = MP._lu(A)
        det = mpmath.mpf(sign)
        for i in range(A.shape[0]):
            det = det * LU[i, i]
        if det < 0:
            raise DomainError("logdet of a matrix with negative determinant")
        return mpmath.log(det)

    sqrt = staticmethod(mpmath.sqrt)
    sin = staticmethod(mpmath.sin)
    cos = staticmethod(mpmath.cos)

    @staticmethod
    def isfinite(v):
        return mpmath.isfinite(v)

    @staticmethod
    def to_f64(a):
        a = np.asarray(a, dtype=object)
        out = np.empty(a.shape, dtype=np.float64)
        for idx in np.ndindex(a.shape):
            out[idx] = float(a[idx])
        return out


# ----------------------------------------------------------------------------------------------------------------
# forward-mode AD: second-order jets over nv seeded variables (what ForwardDiff's nested duals compute,
# ileqg.jl:265-273: fx, fu, cx, cu, cux, cxx, cuu, hx, hxx)
# ----------------------------------------------------------------------------------------------------------------
class Jet:
    __slots__ = ("v", "g", "H")

    def __init__(self, v, g, H):
        self.v, self.g, self.H = v, g, H  # H is None for first-order jets

    @staticmethod
    def seed(LA, values, second_order):
        nv = len(values)
        out = []
        for i, v in enumerate(values):
            g = LA.zeros(nv)
            g[i] = LA.num(1)
            out.append(Jet(v, g, LA.zeros(nv, nv) if second_order else None))
        return out

    def _chain(self, f0, f1, f2):
        H = None
        if self.H is not None:
            H = f1 * self.H + f2 * np.outer(self.g, self.g)
        return Jet(f0, f1 * self.g, H)

    def __add__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v + o.v, self.g + o.g, None if self.H is None else self.H + o.H)
        return Jet(self.v + o, self.g, self.H)

    __radd__ = __add__

    def __neg__(self):
        return Jet(-self.v, -self.g, None if self.H is None else -self.H)

    def __sub__(self, o):
        if isinstance(o, Jet):
            return Jet(self.v - o.v, self.g - o.g, None if self.H is None else self.H - o.H)
        return Jet(self.v - o, self.g, self.H)

    def __rsub__(self, o):
        return (-self) + o

    def __mul__(self, o):
        if isinstance(o, Jet):
            H = None
            if self.H is not None:
                gg = np.outer(self.g, o.g)
                H = self.H * o.v + self.v * o.H + gg + gg.T
            return Jet(self.v * o.v, self.g * o.v + self.v * o.g, H)
        return Jet(self.v * o, self.g * o, None if self.H is None else self.H * o)

    __rmul__ = __mul__

    def recip(self):
        r = 1 / self.v
        return self._chain(r, -r * r, 2 * r * r * r)

    def __truediv__(self, o):
        if isinstance(o, Jet):
            return self * o.recip()
        return self * (1 / o)

    def __rtruediv__(self, o):
        return self.recip() * o


def _pow(x, p):
    """x^p for a real exponent p (Julia: negative base -> DomainError)"""
    v = x.v if isinstance(x, Jet) else x
    if v < 0:
        raise DomainError("negative base of a real power")
    if not isinstance(x, Jet):
        return v ** p
    f0 = v ** p
    f1 = p * v ** (p - 1)
    f2 = p * (p - 1) * v ** (p - 2) if x.H is not None else 0
    return x._chain(f0, f1, f2)


def _sin(LA, x):
    if isinstance(x, Jet):
        s, c = LA.sin(x.v), LA.cos(x.v)
        return x._chain(s, c, -s)
    return LA.sin(x)


def _cos(LA, x):
    if isinstance(x, Jet):
        s, c = LA.sin(x.v), LA.cos(x.v)
        return x._chain(c, -s, -c)
    return LA.cos(x)


# ----------------------------------------------------------------------------------------------------------------
# problems: f(x, u), c(k, x, u), h(x), W(k), N  (optimal_control_problems.jl:67-73); models restated from DESIGN.md
# ----------------------------------------------------------------------------------------------------------------
def make_dynamics(LA, model, p):
    """model: name; p: parameter list (same numbers the C ABI takes).  Returns f(x, u) over generic scalars."""
    p = [LA.num(v) for v in p]

    if model == "single_integrator":  # x + dt u
        def f(x, u):
            return [x[0] + p[0] * u[0], x[1] + p[0] * u[1]]
    elif model == "power_law":  # x.^a + u.^b   (test/ileqg_test.jl:151)
        def f(x, u):
            return [_pow(x[0], p[0]) + _pow(u[0], p[1]), _pow(x[1], p[0]) + _pow(u[1], p[1])]
    elif model == "double_integrator":
        def f(x, u):
            dt = p[0]
            return [x[0] + dt * x[2], x[1] + dt * x[3], x[2] + dt * u[0], x[3] + dt * u[1]]
    elif model == "pendulum":
        def f(x, u):
            dt, g, ln, mass, damp = p
            inertia = mass * ln * ln
            alpha = (u[0] - damp * x[1] - (mass * g * ln) * _sin(LA, x[0])) / inertia
            return [x[0] + dt * x[1], x[1] + dt * alpha]
    elif model == "cartpole":
        def f(x, u):
            dt, mc, mpole, ln, g = p
            s, c = _sin(LA, x[1]), _cos(LA, x[1])
            den = mc + mpole * (s * s)
            thd2 = x[3] * x[3]
            acc = (u[0] + mpole * s * (ln * thd2 + g * c)) / den
            thacc = (-(u[0] * c) - (mpole * ln) * thd2 * c * s - ((mc + mpole) * g) * s) / (ln * den)
            return [x[0] + dt * x[2], x[1] + dt * x[3], x[2] + dt * acc, x[3] + dt * thacc]
    elif model == "unicycle":  # (px, py, psi, v ; a, omega)
        def f(x, u):
            dt = p[0]
            s, c = _sin(LA, x[2]), _cos(LA, x[2])
            return [x[0] + dt * (x[3] * c), x[1] + dt * (x[3] * s), x[2] + dt * u[1], x[3] + dt * u[0]]
    elif model == "quadrotor":
        def f(x, u):
            dt, mass, g, Ix, Iy, Iz = p
            sph, cph = _sin(LA, x[3]), _cos(LA, x[3])
            sth, cth = _sin(LA, x[4]), _cos(LA, x[4])
            sps, cps = _sin(LA, x[5]), _cos(LA, x[5])
            tth = sth / cth
            wp, wq, wr = x[9], x[10], x[11]
            qr = wq * sph + wr * cph
            dphi = wp + qr * tth
            dth = wq * cph - wr * sph
            dpsi = qr / cth
            tm = u[0] / mass
            ax = tm * (cph * sth * cps + sph * sps)
            ay = tm * (cph * sth * sps - sph * cps)
            az = tm * (cph * cth) - g
            dwp = (u[1] + (Iy - Iz) * (wq * wr)) / Ix
            dwq = (u[2] + (Iz - Ix) * (wp * wr)) / Iy
            dwr = (u[3] + (Ix - Iy) * (wp * wq)) / Iz
            d = [x[6], x[7], x[8], dphi, dth, dpsi, ax, ay, az, dwp, dwq, dwr]
            return [x[i] + dt * d[i] for i in range(12)]
    else:
        raise ValueError(model)
    return f


def make_quadratic_cost(LA, n, m, cp):
    """c(k,x,u) = (ws0 + ws1 k)(1/2 dx'Q dx + 1/2 u'R u + dx'Pc u) + c0 + c1 k ;  h(x) = 1/2 dx'Qf dx + h0
    cp = [ws0, ws1, c0, c1, h0, xg(n), Q(n*n), R(m*m), Pc(n*m), Qf(n*n)], matrices column-major (include/ratilqr.h)"""
    cp = [LA.num(v) for v in cp]
    ws0, ws1, c0, c1, h0 = cp[:5]
    o = 5
    xg = cp[o:o + n]; o += n
    Q = [[cp[o + i + j * n] for j in range(n)] for i in range(n)]; o += n * n
    R = [[cp[o + i + j * m] for j in range(m)] for i in range(m)]; o += m * m
    Pc = [[cp[o + i + j * n] for j in range(m)] for i in range(n)]; o += n * m
    Qf = [[cp[o + i + j * n] for j in range(n)] for i in range(n)]

    def quad(M, a, b):  # a' M b, skipping exact zeros of M (a zero coefficient contributes an exact zero)
        acc = 0
        for i in range(len(a)):
            for j in range(len(b)):
                if M[i][j] != 0:
                    acc = acc + M[i][j] * (a[i] * b[j])
        return acc

    def c(k, x, u):
        dx = [x[i] - xg[i] for i in range(n)]
        w = ws0 + ws1 * k
        return w * (0.5 * quad(Q, dx, dx) + 0.5 * quad(R, u, u) + quad(Pc, dx, u)) + c0 + c1 * k

    def h(x):
        dx = [x[i] - xg[i] for i in range(n)]
        return 0.5 * quad(Qf, dx, dx) + h0

    return c, h


def make_power_law_cost(LA, cp):
    """c = sum(x.^p) + sum(u.^p), h = h0 (test/ileqg_test.jl:152-153); cp = [p, h0]"""
    p, h0 = LA.num(cp[0]), LA.num(cp[1])

    def c(k, x, u):
        acc = 0
        for i in range(len(x)):
            acc = acc + (_pow(x[i], p) + _pow(u[i], p))
        return acc

    def h(x):
        return h0

    return c, h


class Problem:
    def __init__(self, f, c, h, W, N, n, m):
        self.f, self.c, self.h, self.W, self.N, self.n, self.m = f, c, h, W, N, n, m


# ----------------------------------------------------------------------------------------------------------------
# ileqg.jl
# ----------------------------------------------------------------------------------------------------------------
def simulate_dynamics_open(LA, problem, x_0, u_array):  # :18-38
    x_array = [LA.arr(x_0)]
    for ii in range(problem.N):
        x_array.append(LA.arr(problem.f(list(x_array[ii]), list(u_array[ii]))))
    return x_array


def simulate_dynamics_closed(LA, problem, x_array, l_array, L_array):  # :62-87
    x_new = [x_array[0].copy()]
    u_new = []
    for ii in range(problem.N):
        u = l_array[ii] + L_array[ii].dot(x_new[ii] - x_array[ii])  # l + L*(x - x̄)
        u_new.append(u)
        x_new.append(LA.arr(problem.f(list(x_new[ii]), list(u))))
    return x_new, u_new


class Approx:
    pass


def approximate_model(LA, problem, u_array, x_array):  # :258-322
    n, m, N = problem.n, problem.m, problem.N
    ap = Approx()
    ap.q, ap.q_vec, ap.Q, ap.r, ap.R, ap.P, ap.A, ap.B, ap.W = [], [], [], [], [], [], [], [], []
    for ii in range(N):
        x, u = x_array[ii], u_array[ii]
        k = ii  # Julia passes ii - 1 with 1-based ii: the 0-based stage index
        jets = Jet.seed(LA, list(x) + list(u), True)
        cj = problem.c(k, jets[:n], jets[n:])
        if not isinstance(cj, Jet):  # constant cost
            cj = Jet(cj, LA.zeros(n + m), LA.zeros(n + m, n + m))
        ap.q.append(cj.v)                                   # c(k, x, u)
        ap.q_vec.append(cj.g[:n].copy())                    # cx
        Qxx = cj.H[:n, :n]
        ap.Q.append(np.triu(Qxx) + np.triu(Qxx, 1).T)       # Symmetric(hessian): upper triangle mirrored
        ap.r.append(cj.g[n:].copy())                        # cu
        Ruu = cj.H[n:, n:]
        ap.R.append(np.triu(Ruu) + np.triu(Ruu, 1).T)       # Symmetric(cuu)
        ap.P.append(cj.H[n:, :n].copy())                    # cux = d(grad_u c)/dx   (m x n)
        j1 = Jet.seed(LA, list(x) + list(u), False)
        fj = problem.f(j1[:n], j1[n:])
        A = LA.zeros(n, n)
        B = LA.zeros(n, m)
        for i in range(n):
            if isinstance(fj[i], Jet):
                A[i, :] = fj[i].g[:n]
                B[i, :] = fj[i].g[n:]
        ap.A.append(A)                                      # fx
        ap.B.append(B)                                      # fu
        ap.W.append(problem.W(k))
    jets = Jet.seed(LA, list(x_array[N]), True)
    hj = problem.h(jets)
    if not isinstance(hj, Jet):
        hj = Jet(hj, LA.zeros(n), LA.zeros(n, n))
    ap.q.append(hj.v)
    ap.q_vec.append(hj.g.copy())
    ap.Q.append(np.triu(hj.H) + np.triu(hj.H, 1).T)
    return ap


def _sym_upper(M):  # Symmetric(M): the upper triangle, mirrored
    return np.triu(M) + np.triu(M, 1).T


class DPResult:
    pass


def _stage(LA, ap, ii, S_next, s_vec_next, s_next, theta, mu, L_given, dl_given, optimise):
    """one pass through the loop body of :360-395 (optimise) / :434-461 (evaluate); returns None when H is not PD"""
    q, q_vec, Q = ap.q[ii], ap.q_vec[ii], ap.Q[ii]
    r, R, P = ap.r[ii], ap.R[ii], ap.P[ii]
    A, B = ap.A[ii], ap.B[ii]
    W = ap.W[ii]
    n, m = A.shape[0], B.shape[1]
    M = _sym_upper(LA.inv(W) - theta * S_next)                              # :365 / :439
    if not LA.isposdef(M):                                                   # :366 / :440
        raise NotPosDef(f"(inv(W) - θ*S) is not PSD at ii = {ii + 1}")
    # D = I + (θ.*S)/M ; X/M = (M' \ X')'                                     # :367 / :441
    D = LA.eye(n) + LA.solve(M.T, (theta * S_next).T).T
    g = r + (B.T.dot(D)).dot(s_vec_next)                                     # :368 / :442
    DS = D.dot(S_next)
    G = P + (B.T.dot(DS)).dot(A)                                             # :369 / :443
    H = R + (B.T.dot(DS)).dot(B) + mu * LA.eye(m)                            # :370 / :444
    H = _sym_upper(H)                                                        # :371 / :445
    if optimise:
        if not LA.isposdef(H):                                               # :372
            return None
        L = LA.solve(-H, G)                                                  # :379  (-H)\G
        dl = LA.solve(-H, g)                                                 # :381
    else:
        L = L_given                                                          # :446
        dl = dl_given if dl_given is not None else LA.zeros(m)               # :447-451
    s = q + s_next + ((0.5 * dl).dot(H)).dot(dl) + dl.dot(g)                 # :383 / :452
    if theta == 0:                                                           # :384 / :453
        s = s + 0.5 * np.trace(W.dot(S_next))
    else:
        # θ/2*s_vec'/M*s_vec - 1/(2θ)*logdet(W*M)                            # :387 / :456
        row = LA.solve(M.T, (theta / 2) * s_vec_next)  # (θ/2 s_vec') / M
        s = s + row.dot(s_vec_next) - 1 / (2 * theta) * LA.logdet(W.dot(M))
    s_vec = q_vec + (A.T.dot(D)).dot(s_vec_next) + (L.T.dot(H)).dot(dl) + L.T.dot(g) + G.T.dot(dl)          # :389 / :458
    S = Q + ((A.T.dot(D)).dot(S_next)).dot(A) + (L.T.dot(H)).dot(L) + L.T.dot(G) + G.T.dot(L)               # :390 / :459
    S = _sym_upper(S)                                                        # :391 / :460
    return s, s_vec, S, g, G, H, L, dl


def solve_approximate_dp_opt(LA, sol, ap, theta):  # solve_approximate_dp! :341-406
    N = len(ap.W)
    res = DPResult()
    res.s, res.s_vec, res.S = [None] * (N + 1), [None] * (N + 1), [None] * (N + 1)
    res.g, res.G, res.H = [None] * N, [None] * N, [None] * N
    res.s[N], res.s_vec[N], res.S[N] = ap.q[N], ap.q_vec[N], _sym_upper(ap.Q[N])
    dl_new = [None] * N
    all_psd = False
    while not all_psd:
        for ii in reversed(range(N)):
            out = _stage(LA, ap, ii, res.S[ii + 1], res.s_vec[ii + 1], res.s[ii + 1], theta, sol.mu, None, None, True)
            if out is None:
                increase_mu_and_delta(sol)                                   # :373
                break
            res.s[ii], res.s_vec[ii], res.S[ii], res.g[ii], res.G[ii], res.H[ii], L, dl = out
            sol.L_array[ii] = L                                              # :380
            dl_new[ii] = dl
            if ii == 0:
                all_psd = True
    return res, dl_new


def solve_approximate_dp(LA, ap, L_array, dl_array, theta, mu):  # :412-465
    N = len(ap.W)
    res = DPResult()
    res.s, res.s_vec, res.S = [None] * (N + 1), [None] * (N + 1), [None] * (N + 1)
    res.g, res.G, res.H = [None] * N, [None] * N, [None] * N
    res.s[N], res.s_vec[N], res.S[N] = ap.q[N], ap.q_vec[N], _sym_upper(ap.Q[N])
    for ii in reversed(range(N)):
        out = _stage(LA, ap, ii, res.S[ii + 1], res.s_vec[ii + 1], res.s[ii + 1], theta, mu, L_array[ii],
                     None if dl_array is None else dl_array[ii], False)
        res.s[ii], res.s_vec[ii], res.S[ii], res.g[ii], res.G[ii], res.H[ii], _, _ = out
    return res


class Solver:  # ILEQGSolver :164-208 (defaults :191-194)
    def __init__(self, LA, mu_min=1e-6, delta_0=2.0, lam=0.5, d=1e-2, iter_max=100, eps_init=1.0, eps_min=1e-6,
                 adaptive_eps_init=False):
        self.LA = LA
        self.mu_min, self.delta_0, self.lam, self.d = LA.num(mu_min), LA.num(delta_0), LA.num(lam), LA.num(d)
        self.iter_max, self.eps_init_init, self.eps_min = iter_max, LA.num(eps_init), LA.num(eps_min)
        self.eps_init_auto = adaptive_eps_init
        self.restarts = 0


def increase_mu_and_delta(sol):  # :471-474
    sol.delta = max(sol.delta_0, sol.delta * sol.delta_0)
    sol.mu = max(sol.mu_min, sol.mu * sol.delta)
    sol.restarts += 1


def initialize(LA, sol, problem, x_0, u_array, theta):  # :214-236
    sol.mu, sol.delta = LA.num(0), sol.delta_0
    sol.d_current = math.inf
    sol.iter_current = 0
    sol.eps_init = sol.eps_init_init
    sol.eps_history = []
    sol.x_array = simulate_dynamics_open(LA, problem, x_0, u_array)
    sol.l_array = [LA.arr(u) for u in u_array]
    sol.L_array = [LA.zeros(problem.m, problem.n) for _ in range(problem.N)]
    ap = approximate_model(LA, problem, sol.l_array, sol.x_array)
    dp = solve_approximate_dp(LA, ap, sol.L_array, None, theta, sol.mu)
    sol.value_current = dp.s[0]


def _isapprox(a, b):  # Base.isapprox: rtol = sqrt(eps(Float64)), atol = 0
    if a == b:
        return True
    return abs(a - b) <= 1.4901161193847656e-8 * max(abs(a), abs(b))


def _norm(LA, v):
    acc = 0
    for e in v:
        acc = acc + e * e
    return LA.sqrt(acc)


def line_search(LA, sol, problem, dl_array_new, theta):  # :494-592
    cur = sol.value_current
    eps = sol.eps_init
    count = 0
    while True:
        count += 1
        if count > 4000:
            raise RuntimeError("line search does not terminate (the reference has no exit here)")
        l_new = [sol.l_array[i] + eps * dl_array_new[i] for i in range(problem.N)]           # :509
        x_new, u_new = simulate_dynamics_closed(LA, problem, sol.x_array, l_new, sol.L_array)  # :517-519
        ap_new = approximate_model(LA, problem, u_new, x_new)                                 # :520
        try:
            dp_new = solve_approximate_dp(LA, ap_new, sol.L_array, None, theta, sol.mu)       # :523-525
        except (NotPosDef, DomainError, ZeroDivisionError, np.linalg.LinAlgError):
            eps = eps * sol.lam                                                                # :530
            continue
        new = dp_new.s[0]
        sol.eps_history.append((eps, new - cur))                                              # :537
        accept = _isapprox(new, cur) or new < cur                                             # :538
        if not accept:
            eps = eps * sol.lam                                                                # :557
            if eps < sol.eps_min:                                                              # :558
                accept = True
        if accept:
            sol.d_current = max(_norm(LA, sol.l_array[i] - u_new[i]) for i in range(problem.N))  # :539 / :559
            sol.value_current = new
            sol.x_array = x_new
            sol.l_array = u_new
            break
    if sol.eps_init_auto:                                                                      # :582-591
        if count == 1:
            sol.eps_init = min(sol.eps_init_init, eps / sol.lam)
        else:
            while eps < sol.eps_min:
                eps = eps / sol.lam
            sol.eps_init = eps


def step(LA, sol, problem, theta):  # :598-613
    sol.iter_current += 1
    ap = approximate_model(LA, problem, sol.l_array, sol.x_array)
    _, dl_array = solve_approximate_dp_opt(LA, sol, ap, theta)
    line_search(LA, sol, problem, dl_array, theta)


def solve(LA, sol, problem, x_0, u_array, theta):  # :635-659
    theta = LA.num(theta)
    initialize(LA, sol, problem, LA.arr(x_0), [LA.arr(u) for u in u_array], theta)
    while True:
        step(LA, sol, problem, theta)
        if sol.d > sol.d_current and sol.mu <= sol.mu_min:       # :642
            break
        elif sol.iter_current == sol.iter_max:                   # :648
            break
    return sol.x_array, sol.l_array, sol.L_array, sol.value_current, sol.eps_history
