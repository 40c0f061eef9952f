"""Lane choreography of the banded LDL^T solver (csrc/mccba_dense.cuh: band_ldlt_solve_warp), restated in numpy and run on
the CPU: lane = row mod 32, a row's window shifts by one register per pivot, rows are fetched one step early and one
slot to the right, the pivot row retires with its final right-hand side, the backward sweep takes two rows per step.
The device code is tested against numpy in tests/test_dense_gpu.py; this file pins the index arithmetic (entry,
retirement, wrap-around of the 32 lanes, ragged sizes) without a GPU, for every band width the kernel is built for."""
import numpy as np
import pytest


def band_solve(S, g, NW):
    n = len(g); w = NW - 1; L = 32
    band = np.zeros((n + 2, NW))
    for r in range(n):
        for k in range(NW):
            c = r - w + k
            if c >= 0:
                band[r, k] = S[r, c]
    rhs = np.concatenate([g.astype(float), np.zeros(2)])
    a = np.zeros((L, NW + 1)); rr = np.zeros(L); row = np.arange(L)
    for l in range(L):                                   # rows 0..w active, row w+1 waits one slot to the right
        inn = row[l] <= w and row[l] < n
        nx = row[l] == w + 1 and row[l] < n
        for k in range(NW + 1):
            x = 0.0
            if inn and k <= row[l]:
                x = band[row[l], (w - row[l]) + k]
            if nx and k >= 1:
                x = band[row[l], k - 1]
            a[l, k] = x
        rr[l] = rhs[row[l]] if (inn or nx) else 0.0
        if (not inn) and (not nx) and row[l] <= w + 1:
            row[l] += 32
    bufs = np.zeros((2, 128))

    def fetch(buf, pj):
        return buf[pj], [buf[pj + 1 + k] for k in range(w)], buf[64 + pj]

    for l in range(L):
        bufs[0, l] = a[l, 0]; bufs[0, l + 32] = a[l, 0]; bufs[0, 64 + l] = rr[l]
    d, c, rp = fetch(bufs[0], 0)
    t = np.array([(-a[l, 0] / d) if (0 < row[l] <= w and row[l] < n) else 0.0 for l in range(L)])
    for j in range(n):                                   # forward: software-pipelined pivot steps
        par = (j + 1) & 1
        a0 = a[:, 0].copy(); a0n = a[:, 1] + t * (c[0] if w else 0.0); rn = rr + t * rp
        for l in range(L):
            bufs[par, l] = a0n[l]; bufs[par, l + 32] = a0n[l]; bufs[par, 64 + l] = rn[l]
        dn, cn, rpn = fetch(bufs[par], (j + 1) & 31)
        tn = np.zeros(L)
        for l in range(L):
            if j <= row[l] <= j + w and row[l] < n:
                band[row[l], j - (row[l] - w)] = a0[l]   # unscaled factor entry; d_j for the pivot row
            piv = row[l] == j
            if piv:
                rhs[j] = rn[l]
            for k in range(1, w):
                a[l, k] = a[l, k + 1] + t[l] * c[k]
            a[l, 0] = a0n[l]; a[l, w] = a[l, w + 1]; a[l, w + 1] = 0.0
            rr[l] = rn[l]
            if piv:
                row[l] += 32
            re = j + 2 + w
            ent = (l == (re & 31)) and re < n
            if ent:
                row[l] = re
            below = j + 1 < row[l] <= j + 1 + w and row[l] < n
            tn[l] = (-a0n[l] / dn) if below else 0.0
            if ent:
                a[l, 1:NW + 1] = band[re, 0:NW]; rr[l] = rhs[re]
        d, c, rp, t = dn, cn, rpn, tn
    ne = n + (n & 1)                                     # backward: pairs of rows, identity padding
    if ne != n:
        band[n, :] = 0.0; band[n, w] = 1.0; rhs[n] = 0.0
    inv = np.array([1.0 / band[q, w] if q < n else band[q, w] for q in range(ne)])
    x = np.zeros(ne); acc = rhs[:ne].copy()
    for j in range(ne - 2, -1, -2):
        q00, q01, q11 = inv[j], -band[j + 1, w - 1] * inv[j] * inv[j + 1], inv[j + 1]
        x0 = q00 * acc[j] + q01 * acc[j + 1]; x1 = q11 * acc[j + 1]
        x[j], x[j + 1] = x0, x1
        for r in range(max(0, j - w), j):
            u0 = band[j, r - (j - w)] if r >= j - w else 0.0
            u1 = band[j + 1, r - (j + 1 - w)] if r >= j + 1 - w else 0.0
            acc[r] -= u0 * x0 + u1 * x1
    return x[:n]


def _banded(n, w, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n)); S = M @ M.T
    i, j = np.indices((n, n)); S[np.abs(i - j) > w] = 0
    S += (np.abs(S).sum(axis=1).max() + 1) * np.eye(n)
    return S, rng.standard_normal(n)


@pytest.mark.parametrize("n,w,NW", [(1, 0, 6), (2, 1, 6), (5, 4, 6), (6, 5, 6), (7, 5, 6), (12, 11, 12), (13, 11, 12),
                                    (33, 11, 12), (64, 17, 18), (65, 23, 24), (100, 5, 6), (97, 29, 30), (378, 11, 12)])
def test_band_choreography_matches_numpy(n, w, NW):
    S, g = _banded(n, w, 10 + n + w)
    x = band_solve(S, g, NW)
    ref = np.linalg.solve(S, g)
    assert np.abs(x - ref).max() <= 1e-12 * max(np.abs(ref).max(), 1e-300) * n
