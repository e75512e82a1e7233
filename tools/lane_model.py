"""Lane-level numpy model of the warp tile algebra used by mx_sweep2.cuh (design validation, CPU only).

A warp is 32 lanes; lane L has r = L >> 2, q = L & 3.  An 8x8 tile T lives in "C layout":
lane L holds t0 = T[r][2q], t1 = T[r][2q+1] (the accumulator layout of mma.m8n8k4.f64).
Everything below is written with whole-warp arrays of shape [32] so that every shuffle is an
explicit gather, exactly like the CUDA code.  Run:  python tools/lane_model.py
"""
import numpy as np

L = np.arange(32)
R = L >> 2
Qn = L & 3


def shfl(x, src):
    return x[src]


def to_tile(t0, t1):
    T = np.zeros((8, 8))
    T[R, 2 * Qn] = t0
    T[R, 2 * Qn + 1] = t1
    return T


def from_tile(T):
    return T[R, 2 * Qn].copy(), T[R, 2 * Qn + 1].copy()


def dmma(c, a, b):
    """D(8x8) += A(8x4) B(4x8); lane gives a = A[r][q], b = B[q][r]; c = (C[r][2q], C[r][2q+1])."""
    A = np.zeros((8, 4)); B = np.zeros((4, 8))
    A[R, Qn] = a
    B[Qn, R] = b
    C = to_tile(*c) + A @ B
    return from_tile(C)


def mma_nt(c, x, y):
    """c += X Y^T for two C-layout tiles (k permuted: k-step e uses columns 2q+e)."""
    c = dmma(c, x[0], y[0])
    c = dmma(c, x[1], y[1])
    return c


def panel(tiles, E):
    """Right-looking Cholesky of the 8-column panel [tiles[0] (diag); tiles[1:]] plus the identity tile E.
    Returns (ok, logdet).  tiles / E are lists [t0, t1] modified in place.  E becomes U = L_d^{-T}."""
    ok = True
    logdet = 0.0
    P0 = tiles[0]
    for j in range(8):
        jq, je = j >> 1, j & 1
        ajj = shfl(P0[je], np.full(32, 4 * j + jq))
        if not np.all(ajj > 0):
            ok = False
        logdet += np.log(ajj[0])
        rinv = 1.0 / np.sqrt(ajj)
        # scale column j of every tile
        for T in tiles + [E]:
            T[je] = np.where(Qn == jq, T[je] * rinv, T[je])
        lk0 = shfl(P0[je], 4 * (2 * Qn) + jq)
        lk1 = shfl(P0[je], 4 * (2 * Qn + 1) + jq)
        for T in tiles + [E]:
            lij = shfl(T[je], 4 * R + jq)
            T[0] = np.where(2 * Qn > j, T[0] - lij * lk0, T[0])
            T[1] = np.where(2 * Qn + 1 > j, T[1] - lij * lk1, T[1])
    P0[0] = np.where(R >= 2 * Qn, P0[0], 0.0)
    P0[1] = np.where(R >= 2 * Qn + 1, P0[1], 0.0)
    return ok, logdet


def panel_unnormalized(P0, E):
    """Column steps as in chol_panel of mx_sweep2.cuh: the trailing update uses the UNSCALED column and 1/a_jj,
    so that the reciprocal square root (needed only for the final L values) is off the critical path."""
    ok = True
    logdet = 0.0
    for j in range(8):
        jq, je = j >> 1, j & 1
        ajj = shfl(P0[je], np.full(32, 4 * j + jq))
        if not np.all(ajj > 0):
            ok = False
        logdet += np.log(ajj[0])
        lk0 = shfl(P0[je], 4 * (2 * Qn) + jq)
        lk1 = shfl(P0[je], 4 * (2 * Qn + 1) + jq)
        lp = shfl(P0[je], 4 * R + jq)
        le = shfl(E[je], 4 * R + jq)
        inv = 1.0 / ajj
        rinv = 1.0 / np.sqrt(ajj)
        P0[0] = np.where(2 * Qn > j, P0[0] - (lp * lk0) * inv, P0[0])
        P0[1] = np.where(2 * Qn + 1 > j, P0[1] - (lp * lk1) * inv, P0[1])
        E[0] = np.where(2 * Qn > j, E[0] - (le * lk0) * inv, E[0])
        E[1] = np.where(2 * Qn + 1 > j, E[1] - (le * lk1) * inv, E[1])
        P0[je] = np.where(Qn == jq, P0[je] * rinv, P0[je])
        E[je] = np.where(Qn == jq, E[je] * rinv, E[je])
    P0[0] = np.where(R >= 2 * Qn, P0[0], 0.0)
    P0[1] = np.where(R >= 2 * Qn + 1, P0[1], 0.0)
    return ok, logdet


def panel_w(tiles, E):
    """The variant used by mx_sweep2.cuh: column steps on the diagonal tile and the identity tile only, then
    L[I][JB] = A[I][JB] W^T through two MMAs per tile with W = U^T (in-register transpose)."""
    P0 = tiles[0]
    ok, logdet = panel_unnormalized(P0, E)
    s0 = 8 * Qn + (R >> 1)
    s1 = s0 + 4
    a0, b0 = shfl(E[0], s0), shfl(E[1], s0)
    a1, b1 = shfl(E[0], s1), shfl(E[1], s1)
    W0 = np.where(R & 1, b0, a0)
    W1 = np.where(R & 1, b1, a1)
    for T in tiles[1:]:
        c = mma_nt((np.zeros(32), np.zeros(32)), (T[0], T[1]), (W0, W1))
        T[0], T[1] = c[0], c[1]
    return ok, logdet


def cholesky(A, NT):
    """A: dict (I,J)->[t0,t1] lower tiles.  In place -> L; returns (ok, logdet, U list)."""
    U = []
    ok = True
    logdet = 0.0
    for jb in range(NT):
        E = list(from_tile(np.eye(8)))
        col = [A[(I, jb)] for I in range(jb, NT)]
        o, ld = panel_w(col, E)
        ok &= o
        logdet += ld
        U.append(E)
        for I in range(jb + 1, NT):
            for J in range(jb + 1, I + 1):
                neg = [-A[(I, jb)][0], -A[(I, jb)][1]]
                A[(I, J)] = list(mma_nt(tuple(A[(I, J)]), neg, A[(J, jb)]))
    return ok, logdet, U


def quadreduce(x):
    x = x + shfl(x, L ^ 1)
    x = x + shfl(x, L ^ 2)
    return x


def colreduce(x):
    x = x + shfl(x, L ^ 4)
    x = x + shfl(x, L ^ 8)
    x = x + shfl(x, L ^ 16)
    return x


def solve(Lt, U, f, NT):
    """L L^T x = f.  f: flat vector (8 NT).  Returns x (flat)."""
    fr = [f[8 * I + R] for I in range(NT)]                 # row replicated
    racc = [np.zeros(32) for _ in range(NT)]
    zc = [None] * NT
    for jb in range(NT):
        rr = fr[jb] - (quadreduce(racc[jb]) if jb > 0 else 0.0)
        zc0 = colreduce(U[jb][0] * rr)
        zc1 = colreduce(U[jb][1] * rr)
        zc[jb] = (zc0, zc1)
        for I in range(jb + 1, NT):
            racc[I] = racc[I] + Lt[(I, jb)][0] * zc0 + Lt[(I, jb)][1] * zc1
    cacc = [[np.zeros(32), np.zeros(32)] for _ in range(NT)]
    x = np.zeros(8 * NT)
    for jb in range(NT - 1, -1, -1):
        c0 = zc[jb][0] - (colreduce(cacc[jb][0]) if jb < NT - 1 else 0.0)
        c1 = zc[jb][1] - (colreduce(cacc[jb][1]) if jb < NT - 1 else 0.0)
        xr = quadreduce(U[jb][0] * c0 + U[jb][1] * c1)
        x[8 * jb + R] = xr
        for J in range(jb):
            cacc[J][0] = cacc[J][0] + Lt[(jb, J)][0] * xr
            cacc[J][1] = cacc[J][1] + Lt[(jb, J)][1] * xr
    return x


def main():
    rng = np.random.default_rng(0)
    NT = 4
    n = 8 * NT
    M = rng.standard_normal((n, n + 5))
    A = M @ M.T + 0.1 * np.eye(n)
    tiles = {(I, J): list(from_tile(A[8 * I:8 * I + 8, 8 * J:8 * J + 8])) for I in range(NT) for J in range(I + 1)}
    ok, logdet, U = cholesky(tiles, NT)
    Lfull = np.zeros((n, n))
    for (I, J), t in tiles.items():
        Lfull[8 * I:8 * I + 8, 8 * J:8 * J + 8] = to_tile(*t)
    Lref = np.linalg.cholesky(A)
    print("ok", ok, "max |L - Lref|", np.max(np.abs(Lfull - Lref)), "logdet err", logdet - np.linalg.slogdet(A)[1])
    for jb in range(NT):
        Uref = np.linalg.inv(Lref[8 * jb:8 * jb + 8, 8 * jb:8 * jb + 8]).T
        print("  U", jb, np.max(np.abs(to_tile(*U[jb]) - Uref)))
    f = rng.standard_normal(n)
    x = solve(tiles, U, f, NT)
    print("solve err", np.max(np.abs(x - np.linalg.solve(A, f))))
    # tile product / symmetric product check:  J = Z Lam Z
    Zm = M[:, :n] @ M[:, :n].T
    lam = rng.random(n)
    Jref = Zm @ np.diag(lam) @ Zm
    Zt = {(I, K): from_tile(Zm[8 * I:8 * I + 8, 8 * K:8 * K + 8]) for I in range(NT) for K in range(NT)}
    err = 0.0
    for I in range(NT):
        for Jt in range(I + 1):
            c = (np.zeros(32), np.zeros(32))
            for K in range(NT):
                x0 = Zt[(I, K)][0] * lam[8 * K + 2 * Qn]
                x1 = Zt[(I, K)][1] * lam[8 * K + 2 * Qn + 1]
                c = mma_nt(c, (x0, x1), Zt[(Jt, K)])
            err = max(err, np.max(np.abs(to_tile(*c) - Jref[8 * I:8 * I + 8, 8 * Jt:8 * Jt + 8])))
    print("J = Z Lam Z tile err", err / np.max(np.abs(Jref)))


if __name__ == "__main__":
    main()
