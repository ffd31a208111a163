"""SURVEY 8(f) row 1, measured: the harness reader on a large COO text file, host
path (tokeniser + merge sort + fold on one core, as the reference does with
fscanf + qsort) against LSBENCH_B200_INGEST=1 (lines, numbers, sort, fold and
row compression on the GPU).  The file is the body of a Nek matrix repeated
`reps` times, i.e. unsorted with every entry duplicated `reps` times -- so the
sort and the fold both have work.  Checks that the two CSRs are identical.
   python tools/ingest_bench.py [reps] [matrix]
"""
import ctypes as C
import hashlib
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import orc
from lsbench_b200 import build_host

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
name = sys.argv[2] if len(sys.argv) > 2 else "tj7a_A_12"
raw = open(orc.matrix_path(name), "rb").read()
nl = raw.index(b"\n")
nnz, base = (int(t) for t in raw[:nl].split())
path = os.path.join(tempfile.gettempdir(), "ingest_bench_%s_x%d.txt" % (name, reps))
with open(path, "wb") as f:
    f.write(b"%d %d\n" % (nnz * reps, base))
    f.write(raw[nl + 1:] * reps)


class Csr(C.Structure):
    _fields_ = [("nrows", C.c_uint), ("base", C.c_uint), ("offs", C.POINTER(C.c_uint)),
                ("cols", C.POINTER(C.c_uint)), ("vals", C.POINTER(C.c_double))]


L = C.CDLL(build_host.LIB)
L.lsbench_matrix_read.restype = C.POINTER(Csr)
L.lsbench_matrix_read.argtypes = [C.c_char_p]
L.lsbench_matrix_free.argtypes = [C.POINTER(Csr)]


def read(ingest):
    os.environ["LSBENCH_B200_INGEST"] = "1" if ingest else "0"
    t0 = time.perf_counter()
    p = L.lsbench_matrix_read(path.encode())
    dt = time.perf_counter() - t0
    a = p.contents
    m = a.offs[a.nrows]
    h = hashlib.sha256()
    for arr, n, ty in ((a.offs, a.nrows + 1, np.uint32), (a.cols, m, np.uint32), (a.vals, m, np.float64)):
        h.update(np.ctypeslib.as_array(arr, (n,)).astype(ty, copy=False).tobytes())
    out = (a.nrows, int(m), h.hexdigest())
    L.lsbench_matrix_free(p)
    return dt, out


read(True)                       # context creation, module load
t_dev = min(read(True)[0] for _ in range(3))
dev = read(True)[1]
t_host, host = read(False)
assert dev == host, (dev, host)
print(json.dumps({"file_MB": os.path.getsize(path) / 1e6, "records": nnz * reps, "rows": dev[0],
                  "nnz_after_fold": dev[1], "host_reader_s": t_host, "device_ingest_s": t_dev,
                  "speedup": t_host / t_dev, "identical_csr_sha256": dev[2][:16]}))
os.unlink(path)
