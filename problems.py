"""Synthetic workloads of BASELINE.json's configs.  Nothing in the reference defines them (it ships no
benchmark): the constructions are frozen here with their seeds and shared by bench.py and the parity tests."""
import numpy as np
import scipy.sparse as sp


def random_qp_c2(n, m, density, seed):
    """Config 2: random sparse QP, A = sprandn(m, n, density); P = S + S' + diag (diagonally dominant)
    with S the strict upper triangle of sprandn(n, n, density), i.e. P holds `density` of the full n x n off the
    diagonal: n = 50k, density 1e-3 gives nnz(A) = 5.0e6 and nnz(P_full) = 2.55e6 (triu 1.3e6), the figures of
    SURVEY 8.  (Round 1 drew S at density / 2 -- half the P of the specification.)  Two-sided feasible rows."""
    rng = np.random.default_rng(seed)
    rvs = rng.standard_normal
    A = sp.random(m, n, density=density, random_state=rng, data_rvs=rvs, format="csc")
    S = sp.triu(sp.random(n, n, density=density, random_state=rng, data_rvs=rvs, format="csc"), k=1)
    S = (S + S.T).tocsc()
    d = np.asarray(abs(S).sum(axis=1)).ravel() + rng.uniform(0.1, 1.0, n)
    P = (S + sp.diags(d)).tocsc()
    q = rng.standard_normal(n)
    x0 = rng.standard_normal(n)
    Ax0 = A @ x0
    l = Ax0 - rng.uniform(0, 1, m)
    u = Ax0 + rng.uniform(0, 1, m)
    return dict(P=P, q=q, A=A, l=l, u=u)


def lasso_c3(n_feat, n_samp, density, seed):
    """Config 3: Lasso as a QP (OSQP-paper form).  Variables [x (n_feat); y (n_samp); t (n_feat)];
    min y'y + lambda 1't  s.t.  y = Ad x - b,  -t <= x <= t.  Returns the problem for lambda = 1 together with
    lambda_max and a function giving q(lambda) -- the lambda sweep only changes q (osqp_update_lin_cost)."""
    rng = np.random.default_rng(seed)
    Ad = sp.random(n_samp, n_feat, density=density, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    xh = rng.standard_normal(n_feat) * (rng.random(n_feat) < 0.5) / np.sqrt(n_feat)
    b = Ad @ xh + 0.1 * rng.standard_normal(n_samp)
    n = 2 * n_feat + n_samp
    In, Im = sp.eye(n_feat, format="csc"), sp.eye(n_samp, format="csc")
    P = sp.block_diag([sp.csc_matrix((n_feat, n_feat)), 2.0 * Im, sp.csc_matrix((n_feat, n_feat))], format="csc")
    A = sp.vstack([sp.hstack([Ad, -Im, sp.csc_matrix((n_samp, n_feat))]),
                   sp.hstack([In, sp.csc_matrix((n_feat, n_samp)), -In]),
                   sp.hstack([In, sp.csc_matrix((n_feat, n_samp)), In])], format="csc")
    l = np.concatenate([b, -np.inf * np.ones(n_feat), np.zeros(n_feat)])
    u = np.concatenate([b, np.zeros(n_feat), np.inf * np.ones(n_feat)])
    lam_max = float(np.max(np.abs(Ad.T @ b)))  # x = 0 is optimal for lambda >= 2 lam_max (the objective is y'y)

    def q_of(lam):
        return np.concatenate([np.zeros(n_feat + n_samp), lam * np.ones(n_feat)])

    return dict(P=P, q=q_of(1.0), A=A, l=l, u=u), lam_max, q_of, n


def portfolio_c4(n_assets, k_factors, seed, gamma=1.0):
    """Config 4: factor-model portfolio (OSQP-paper form).  Variables [x (assets); y (factors)];
    min gamma (x'Dx + y'y) - mu'x  s.t.  y = F'x, 1'x = 1, 0 <= x <= 1."""
    rng = np.random.default_rng(seed)
    F = sp.random(n_assets, k_factors, density=0.5, random_state=rng, data_rvs=rng.standard_normal, format="csc")
    D = sp.diags(rng.random(n_assets) * np.sqrt(k_factors), format="csc")
    mu = rng.standard_normal(n_assets)
    P = sp.block_diag([2 * gamma * D, 2 * gamma * sp.eye(k_factors)], format="csc")
    q = np.concatenate([-mu, np.zeros(k_factors)])
    A = sp.vstack([sp.hstack([F.T, -sp.eye(k_factors)]),
                   sp.hstack([sp.csc_matrix(np.ones((1, n_assets))), sp.csc_matrix((1, k_factors))]),
                   sp.hstack([sp.eye(n_assets), sp.csc_matrix((n_assets, k_factors))])], format="csc")
    l = np.concatenate([np.zeros(k_factors), [1.0], np.zeros(n_assets)])
    u = np.concatenate([np.zeros(k_factors), [1.0], np.ones(n_assets)])
    return dict(P=P, q=q, A=A, l=l, u=u)


def mpc_batch_c5(count, seed, T=10, nx=2, nu=1):
    """Config 5: `count` MPC QPs with one sparsity pattern.  Variables [x_1..x_T; u_0..u_{T-1}] (x_0 eliminated):
    n = T (nx + nu) = 30; rows = T nx dynamics equalities + n variable bounds + T nu input-rate bounds = 60.
    (Ad, Bd) random stable and x_0 random per instance, Q = I, R = 0.1.
    Returns (P_pattern, A_pattern, Px [count, nnzP], Ax [count, nnzA], q, l, u)."""
    n = T * (nx + nu)
    nxT = T * nx
    # pattern: every structurally possible entry set to 1
    rows, cols = [], []

    def add(r, c):
        rows.append(r); cols.append(c)

    for k in range(T):  # x_{k+1} - Ad x_k - Bd u_k = [k == 0] Ad x_0
        for i in range(nx):
            r = k * nx + i
            add(r, k * nx + i)
            if k > 0:
                for j in range(nx):
                    add(r, (k - 1) * nx + j)
            for j in range(nu):
                add(r, nxT + k * nu + j)
    for v in range(n):
        add(nxT + v, v)
    for k in range(T):
        for j in range(nu):
            r = nxT + n + k * nu + j
            add(r, nxT + k * nu + j)
            if k > 0:
                add(r, nxT + (k - 1) * nu + j)
    m = nxT + n + T * nu
    pat = sp.csc_matrix((np.ones(len(rows)), (rows, cols)), shape=(m, n))
    pat.sort_indices()
    Ppat = sp.eye(n, format="csc")
    # per-instance values, written through a dense scratch matrix in the pattern's CSC order
    rng = np.random.default_rng(seed)
    ci = np.repeat(np.arange(n), np.diff(pat.indptr))
    ri = pat.indices
    Ax = np.empty((count, pat.nnz))
    Px = np.tile(np.concatenate([np.ones(nxT), 0.1 * np.ones(T * nu)]), (count, 1))
    q = np.zeros((count, n))
    l = np.empty((count, m))
    u = np.empty((count, m))
    xmax, umax, dumax = 10.0, 2.0, 1.0
    for b in range(count):
        M = rng.standard_normal((nx, nx))
        Ad = 0.95 * M / max(1e-9, np.max(np.abs(np.linalg.eigvals(M))))
        Bd = rng.standard_normal((nx, nu))
        x0 = rng.uniform(-1, 1, nx)
        Dm = np.zeros((m, n))
        for k in range(T):
            Dm[k * nx:(k + 1) * nx, k * nx:(k + 1) * nx] = np.eye(nx)
            if k > 0:
                Dm[k * nx:(k + 1) * nx, (k - 1) * nx:k * nx] = -Ad
            Dm[k * nx:(k + 1) * nx, nxT + k * nu:nxT + (k + 1) * nu] = -Bd
        Dm[nxT:nxT + n, :] = np.eye(n)
        for k in range(T):
            for j in range(nu):
                r = nxT + n + k * nu + j
                Dm[r, nxT + k * nu + j] = 1.0
                if k > 0:
                    Dm[r, nxT + (k - 1) * nu + j] = -1.0
        Ax[b] = Dm[ri, ci]
        rhs = np.zeros(nxT)
        rhs[:nx] = Ad @ x0
        l[b] = np.concatenate([rhs, -xmax * np.ones(nxT), -umax * np.ones(T * nu), -dumax * np.ones(T * nu)])
        u[b] = np.concatenate([rhs, xmax * np.ones(nxT), umax * np.ones(T * nu), dumax * np.ones(T * nu)])
    return Ppat, pat, Px, Ax, q, l, u


def batch_instance(Ppat, Apat, Px, Ax, q, l, u, k):
    """The k-th QP of a batch as a standalone problem (for the per-QP oracle check)."""
    P = sp.csc_matrix((Px[k], Ppat.indices, Ppat.indptr), shape=Ppat.shape)
    A = sp.csc_matrix((Ax[k], Apat.indices, Apat.indptr), shape=Apat.shape)
    return dict(P=P, q=q[k], A=A, l=l[k], u=u[k])
