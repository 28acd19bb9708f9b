"""The bridge to a real R installation (none here: parity of stages 1-5 stays "unpinned").

    python tests/golden/r_golden.py            # if Rscript + rioja + fpc + jsonlite are found: writes r_pipeline_*.json

For every case below the matrix (integer-only generator, intgen.py: the same bits on any machine) is written as the
header-less TSV the reference reads, the reference's own TADpole() is run on it through make_r_golden.R, and the result is
stored next to this file.  tests/test_oracle.py::test_oracle_against_r_golden compares the oracle with every
r_pipeline_*.json it finds (and tries to make them first when R is there); until such a file exists that test is skipped.
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from intgen import golden_int_matrix  # noqa: E402

CASES = [dict(name="n200", n=200, seed=5, max_pcs=200, centromere=False),
         dict(name="n600", n=600, seed=6, max_pcs=200, centromere=False)]


def r_available():
    rs = shutil.which("Rscript")
    if not rs:
        return None
    ok = subprocess.run([rs, "-e", "for (p in c('rioja','fpc','jsonlite','bigmemory','Matrix','foreach','doParallel')) "
                                   "stopifnot(requireNamespace(p, quietly = TRUE))"], capture_output=True).returncode == 0
    return rs if ok else None


def make(case, rs, out_dir=HERE):
    m = golden_int_matrix(case["n"], case["seed"])
    with tempfile.TemporaryDirectory() as tmp:
        tsv = os.path.join(tmp, "m.tsv")
        np.savetxt(tsv, m.astype(np.int64), fmt="%d", delimiter="\t")
        out = os.path.join(out_dir, f"r_pipeline_{case['name']}.json")
        subprocess.check_call([rs, os.path.join(HERE, "make_r_golden.R"), tsv, out, str(case["max_pcs"]),
                               "TRUE" if case["centromere"] else "FALSE"])
    with open(out) as fh:
        res = json.load(fh)
    res["case"] = case
    with open(out, "w") as fh:
        json.dump(res, fh)
    return out


if __name__ == "__main__":
    rs = r_available()
    if not rs:
        sys.exit("Rscript with rioja, fpc, jsonlite, bigmemory, Matrix, foreach, doParallel not found")
    for c in CASES:
        print(make(c, rs))
