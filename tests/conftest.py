import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


@pytest.fixture(scope="session")
def ctx():
    """One GPU context for the session.  No CPU fallback: a missing library or GPU is an error."""
    from tadpole_b200 import Context
    c = Context(0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def synth_cache():
    """Synthetic matrices + oracle intermediates, computed once per session."""
    from oracle import tadpole_oracle as O
    from tadpole_b200.synth import synth_hic
    cache = {}

    def get(n, seed=1, **kw):
        key = (n, seed, tuple(sorted(kw.items())))
        if key not in cache:
            m = synth_hic(n, seed=seed, **kw)
            lm = O.load_mat_numeric(m)
            cor = O.sparse_cor(lm.mat)
            k = min(200, lm.mat.shape[0])
            pcs = O.prcomp_scores(cor, k)
            cache[key] = dict(mat=m, lm=lm, cor=cor, pcs=pcs, k=k)
        return cache[key]

    return get
