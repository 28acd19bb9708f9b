"""Ad-hoc (not a test): phase trace of the one-sided Jacobi cluster kernel.  TADPOLE_OSJ_TRACE=1 python tests/osj_trace.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
ctx = Context(0)
rng = np.random.default_rng(0)
b = 256
q, _ = np.linalg.qr(rng.standard_normal((b, b)))
lam = 100.0 * (1.0 + np.arange(b)) ** -0.9
dense = (q * lam) @ q.T                                   # dense T with a slowly decaying spectrum (the start solve)
e = rng.standard_normal((b, b)) * 1e-5
near = np.diag(lam) + (e + e.T) * np.sqrt(np.outer(lam, lam))          # nearly diagonal T (later Rayleigh-Ritz steps)
for name, t, tol in (("dense 1e-3", dense, 1e-3), ("dense 1e-9", dense, 1.6e-9), ("near-diagonal 1e-13", near, 1.2e-13)):
    for rep in range(2):
        t0 = time.perf_counter()
        w, v, sw = ctx.test_eig(t, tol=tol)
        dt = time.perf_counter() - t0
    err = np.abs(np.sort(w) - np.sort(np.linalg.eigvalsh(t))).max() / lam[0]
    print(f"{name}: sweeps {sw} wall {dt * 1e3:.2f} ms eig err {err:.1e} orth {np.abs(v.T @ v - np.eye(b)).max():.1e}", file=sys.stderr, flush=True)
