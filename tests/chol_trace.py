"""Ad-hoc (not a test): phase trace of the cluster Cholesky kernel.  TADPOLE_CHOL_TRACE=1 python tests/chol_trace.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tadpole_b200 import Context
ctx = Context(0)
rng = np.random.default_rng(0)
a = rng.standard_normal((2000, 256))
g = a.T @ a
for rep in range(2):
    print("rep", rep, file=sys.stderr)
    l, li, bad = ctx.test_cholinv(g)
print(bad, np.linalg.norm(l @ l.T - g) / np.linalg.norm(g))
