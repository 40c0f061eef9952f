"""Reduced-system solvers in isolation (mccba_debug_solve_dense): every size class of the tile DAG -- single tile, ragged
last block column, g row in its own tile row (n % 32 == 0), the config #5 size -- and the block cyclic reduction of banded
systems (mode 3) against numpy.  (The per-column and panel / update solvers of round 1 are gone: two solvers ship.)  Replaces the Eigen CG of
src/multicalib.cpp:565-592 on the Schur-reduced system."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SIZES = [1, 5, 6, 12, 31, 32, 33, 63, 64, 65, 96, 127, 128, 200, 378, 384]


@pytest.fixture(scope="module")
def solver():
    import multi_camera_calibration_b200 as m
    s = m.Solver(device=0)
    yield s
    s.close()


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T + n * np.eye(n), rng.standard_normal(n)


@pytest.mark.parametrize("mode", [2])
@pytest.mark.parametrize("n", SIZES)
def test_dense_solve_matches_numpy(solver, n, mode):
    S, g = _spd(n, 100 + n)
    x, _ = solver.debug_solve_dense(S, g, mode)
    ref = np.linalg.solve(S, g)
    assert np.abs(x - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300) * n      # fp64, cond ~ 10


@pytest.mark.parametrize("mode", [2])
def test_dense_solve_ill_conditioned(solver, mode):
    """Reduced camera systems of real rigs have condition numbers ~1e9: the factorisation must stay backward stable."""
    n = 190
    rng = np.random.default_rng(7)
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    S = (Q * np.logspace(0, 9, n)) @ Q.T
    S = 0.5 * (S + S.T)
    g = rng.standard_normal(n)
    x, _ = solver.debug_solve_dense(S, g, mode)
    assert np.linalg.norm(S @ x - g) <= 1e-9 * (np.linalg.norm(S, 2) * np.linalg.norm(x) + np.linalg.norm(g))


@pytest.mark.parametrize("mode", [2])
@pytest.mark.parametrize("n", [7, 40, 100])
def test_dense_solve_rejects_indefinite(solver, n, mode):
    import multi_camera_calibration_b200 as m
    S, g = _spd(n, 5)
    S[n // 2, n // 2] = -1.0
    with pytest.raises(m.MccbaError):
        solver.debug_solve_dense(S, g, mode)
    # the handle stays usable
    S2, g2 = _spd(n, 6)
    x, _ = solver.debug_solve_dense(S2, g2, mode)
    assert np.allclose(x, np.linalg.solve(S2, g2), rtol=1e-9, atol=1e-12)


def _banded_spd(n, w, seed):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    S = M @ M.T
    i, j = np.indices((n, n))
    S[np.abs(i - j) > w] = 0.0
    S += (np.abs(S).sum(axis=1).max() + 1.0) * np.eye(n)      # diagonally dominant: SPD
    return S, rng.standard_normal(n)


def _block_banded_spd(n_cam, m, seed):
    """The structure of a reduced camera system: 6 x 6 blocks, block (A, B) non-zero iff |A - B| <= m."""
    rng = np.random.default_rng(seed)
    n = 6 * n_cam
    M = rng.standard_normal((n, n))
    S = M @ M.T
    blk = np.arange(n) // 6
    S[np.abs(blk[:, None] - blk[None, :]) > m] = 0.0
    S += (np.abs(S).sum(axis=1).max() + 1.0) * np.eye(n)
    return S, rng.standard_normal(n)


@pytest.mark.parametrize("n_cam,m", [(1, 0), (2, 1), (3, 1), (8, 1), (63, 1), (64, 1), (65, 1), (63, 2), (63, 3), (63, 4), (40, 4),
                                     (127, 1), (200, 1), (21, 3), (5, 4)])
def test_block_banded_solve_matches_numpy(solver, n_cam, m):
    """mode 3: block cyclic reduction (mccba_bcr.cuh) on block-banded camera systems -- every super-block size
    (B = 6, 12, 18, 24), block counts that are / are not powers of two, ragged last super-block."""
    S, g = _block_banded_spd(n_cam, m, 900 + 10 * n_cam + m)
    x, _ = solver.debug_solve_dense(S, g, 3)
    ref = np.linalg.solve(S, g)
    assert np.abs(x - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300) * S.shape[0]


def test_block_banded_solve_ill_conditioned(solver):
    """cond ~1e9 (undamped Gauss-Newton systems of camera chains): elimination without pivoting must stay backward stable."""
    n_cam, m = 63, 1
    S, g = _block_banded_spd(n_cam, m, 5)
    n = S.shape[0]
    d = np.logspace(0, 4.5, n)
    S = (S * d[:, None]) * d[None, :]                    # congruence: SPD, same sparsity, cond ~ 1e9
    x, _ = solver.debug_solve_dense(S, g, 3)
    assert np.linalg.norm(S @ x - g) <= 1e-9 * (np.linalg.norm(S, 2) * np.linalg.norm(x) + np.linalg.norm(g))


@pytest.mark.parametrize("n,w", [(1, 0), (5, 4), (6, 5), (12, 11), (12, 5), (37, 11), (64, 17), (200, 23), (378, 11), (378, 23), (400, 5)])
def test_banded_solve_matches_numpy(solver, n, w):
    """mode 3 on generic banded matrices: half bandwidth w <= 23 maps to a block bandwidth <= 4."""
    S, g = _banded_spd(n, w, 300 + n + w)
    x, _ = solver.debug_solve_dense(S, g, 3)
    ref = np.linalg.solve(S, g)
    assert np.abs(x - ref).max() <= 1e-11 * max(np.abs(ref).max(), 1e-300) * n


def test_banded_solve_rejects_indefinite_and_wide(solver):
    import multi_camera_calibration_b200 as m
    S, g = _banded_spd(50, 11, 1)
    S[20, 20] = -3.0
    with pytest.raises(m.MccbaError):
        solver.debug_solve_dense(S, g, 3)
    S, g = _spd(64, 2)            # dense: half bandwidth 63 does not fit
    with pytest.raises(m.MccbaError):
        solver.debug_solve_dense(S, g, 3)
