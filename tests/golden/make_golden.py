"""Regenerates tests/golden/*.  Run in the build container, where
/root/reference is mounted:  python tests/golden/make_golden.py

  reader.json  -- for every matrix under tests/data: n, nnz, base and sha256 of
                  offs/cols/vals as produced by the REFERENCE's own
                  lsbench_matrix_read (oracle/_ref/libref_lsbench.so, compiled
                  from /root/reference/src/lsbench-csr.c).
  direct.npz   -- x = A^-1 b for the operator CHOLMOD factorises
                  (src/cholmod-impl.h:5-21, upper triangle mirrored), b[i] = i
                  (src/lsbench.c:159-160), solved with scipy SuperLU -- an
                  implementation independent of both the oracle and the GPU path.
  I1 analytic  -- diag(1..5) x = [0,1,2,3,4]  =>  x = [0, 1/2, 2/3, 3/4, 4/5].
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    import scipy.sparse.linalg as spl
    orc.build_oracle()
    reader, direct = {}, {}
    for name in orc.TOY + orc.NEK:
        R = orc.ref_matrix_read(orc.matrix_path(name))
        assert R is not None, "oracle/_ref missing: run in the build container"
        reader[name] = dict(n=R.nrows, nnz=R.nnz, base=R.base, offs=sha(R.offs),
                            cols=sha(R.cols), vals=sha(R.vals))
        if name.startswith("A"):
            continue  # indefinite 2x2: reader/base fixture only
        M = orc.op_upper_mirror(R)
        S = M.scipy().tocsc()
        x = spl.splu(S).solve(orc.rhs(M.n))
        x = x + spl.splu(S).solve(orc.rhs(M.n) - S @ x)  # one refinement step
        direct[name] = x
        reader[name]["op_nnz"] = M.nnz
        reader[name]["op_vals"] = sha(M.vals)
        reader[name]["op_cols"] = sha(M.cols)
    with open(os.path.join(HERE, "reader.json"), "w") as f:
        json.dump(reader, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "direct.npz"), **direct)
    print("wrote", len(reader), "reader entries,", len(direct), "direct solves")


if __name__ == "__main__":
    main()
