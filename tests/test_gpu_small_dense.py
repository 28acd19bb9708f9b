"""GPU tests of the b x b dense kernels inside stage 3 (tp_pca): cluster Cholesky + triangular inverse
(csrc/cholinv.cu) and the one-sided Jacobi eigensolver (csrc/osj.cu), through the C-ABI test hooks,
against numpy.linalg (float64).  Tolerances are relative to the matrix norm and written in each check."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spd(b, cond, seed):
    rng = np.random.default_rng(seed)
    q, _ = np.linalg.qr(rng.standard_normal((b, b)))
    w = np.logspace(0, -np.log10(cond), b)
    return (q * w) @ q.T


@pytest.mark.parametrize("b", [1, 7, 32, 33, 96, 198, 224, 256, 300])
@pytest.mark.parametrize("cond", [1e2, 1e8])
def test_cholesky_and_inverse(ctx, b, cond):
    g = _spd(b, cond, seed=b)
    g = 0.5 * (g + g.T)
    l, li, bad = ctx.test_cholinv(g)
    assert bad == 0
    ref = np.linalg.cholesky(g)
    # backward error of the factorisation: ||L L^T - G|| <= 1e-14 ||G||
    assert np.linalg.norm(l @ l.T - g) <= 1e-14 * np.linalg.norm(g) * max(1, b ** 0.5)
    # forward error of a backward-stable Cholesky: ~ cond(G) * eps
    np.testing.assert_allclose(l, ref, rtol=0, atol=1e-14 * cond * np.abs(ref).max())
    # the inverse is lower triangular and inverts L: ||Linv L - I|| <= 1e-13 cond(L)
    assert np.abs(np.triu(li, 1)).max() == 0.0
    assert np.linalg.norm(li @ l - np.eye(b)) <= 1e-13 * cond ** 0.5 * b


def test_cholesky_flags_indefinite(ctx):
    g = _spd(128, 1e2, seed=3)
    g[70, 70] = -1.0
    _, _, bad = ctx.test_cholinv(g)
    assert bad == 1


@pytest.mark.parametrize("b", [40, 198, 256])
def test_cholesky_factor_only_semidefinite(ctx, b):
    rng = np.random.default_rng(b)
    a = rng.standard_normal((b, b - 5))
    g = a @ a.T                                    # rank b - 5: pivots are clamped, no failure
    l, _, _ = ctx.test_cholinv(g, factor_only=True)
    assert np.isfinite(l).all()
    assert np.linalg.norm(l @ l.T - g) <= 1e-10 * np.linalg.norm(g)


@pytest.mark.parametrize("b", [2, 31, 198, 256, 300])
def test_eigensolver(ctx, b):
    t = _spd(b, 1e6, seed=100 + b)
    t = 0.5 * (t + t.T)
    w, v, sweeps = ctx.test_eig(t, tol=1e-14)
    wref = np.linalg.eigvalsh(t)[::-1]
    np.testing.assert_allclose(w, wref, rtol=1e-11, atol=1e-14 * wref[0])
    assert np.linalg.norm(v.T @ v - np.eye(b)) <= 1e-12 * b
    assert np.linalg.norm(t @ v - v * w) <= 1e-12 * wref[0] * b
    assert 1 <= sweeps <= 30


@pytest.mark.parametrize("b,scale", [(256, 1.0), (256, 1e-30), (224, 1e12), (96, 1.0)])
def test_eigensolver_fp32_mode(ctx, b, scale):
    """Loose tolerances (>= 1e-4: the start Rayleigh-Ritz step, whose Ritz values only set the filter bounds) are solved by
    the FP32 variant of the cluster Jacobi kernel, whatever the scale of the matrix (the factor is scaled by a power of two
    on load).  Eigenvalues to ~1e-5 of the largest, eigenvectors orthogonal to ~1e-5."""
    t = _spd(b, 1e3, seed=7 + b) * scale
    t = 0.5 * (t + t.T)
    w, v, sweeps = ctx.test_eig(t, tol=1e-2)
    wref = np.linalg.eigvalsh(t)[::-1]
    np.testing.assert_allclose(w, wref, rtol=0, atol=2e-4 * wref[0])
    assert np.abs(v.T @ v - np.eye(b)).max() <= 5e-2
    assert np.linalg.norm(t @ v - v * w, axis=0).max() <= 5e-2 * wref[0]
    assert 1 <= sweeps <= 30
