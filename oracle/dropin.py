"""TEST INFRASTRUCTURE: the drop-in, done for real.

Takes the reference tree where it lies (REF, default /root/reference), makes a
scratch copy OUTSIDE this repository, applies the five registration edits of
INTEGRATION.md section 2 to that copy, drops this repository's
lsbench_b200/host/b200.c and include/b200.h beside src/cusparse.c, and
compiles the reference's own sources + b200.c with -DLSBENCH_B200 against
libb200.so:

    oracle/_ref/libref_lsbench_b200.so     the reference's liblsbench + the b200 backend
    oracle/_ref/driver_ref_b200            the reference's bin/driver.c, unmodified

i.e. `--solver b200` inside a stock lsbench.  Nothing of the reference is
copied into the repository (only the two binaries land in the git-ignored
oracle/_ref/, which travels to the GPU box).  The edits are made by anchor, not
by a stored patch, so no reference text lives here either.

    python oracle/dropin.py [--ref /root/reference]
"""
import argparse
import os
import re
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(HERE, "_ref")
CC = os.environ.get("HOSTCC") or ("/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc")
CXX = os.environ.get("HOSTCXX") or ("/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++")


def edit(path, fn):
    text = open(path).read()
    new = fn(text)
    assert new != text, "registration edit did not apply to " + path
    open(path, "w").write(new)


def sub_once(pattern, repl, text, flags=0):
    new, n = re.subn(pattern, repl, text, count=1, flags=flags)
    assert n == 1, pattern
    return new


def register(src):
    # 1. enum value after the last solver (src/lsbench.h:15)
    edit(os.path.join(src, "lsbench.h"), lambda t: sub_once(
        r"(LSBENCH_SOLVER_GINKGO = 5)\n", r"\1,\n  LSBENCH_SOLVER_B200 = 6\n", t))
    # 2. prototypes beside the other backends' (src/lsbench-impl.h:42-45)
    edit(os.path.join(src, "lsbench-impl.h"), lambda t: sub_once(
        r"(int ginkgo_bench\()",
        "int b200_init();\nint b200_finalize();\n"
        "int b200_bench(double *x, struct csr *A, const double *r,\n"
        "               const struct lsbench *cb);\n\n\\1", t))

    def lsbench_c(t):
        # 3. the solver name (str_to_solver, src/lsbench.c:19-34)
        t = sub_once(r'(\} else if \(strcmp\(up, "GINKGO"\) == 0\) \{\n\s*return LSBENCH_SOLVER_GINKGO;\n)',
                     '\\1  } else if (strcmp(up, "B200") == 0) {\n    return LSBENCH_SOLVER_B200;\n', t)
        # 4. init / finalize beside the others (:143-147, :190-194)
        t = sub_once(r"(\n  paralmond_init\(\);\n)", r"\1  b200_init();\n", t)
        t = sub_once(r"(\n  paralmond_finalize\(\);\n)", r"\1  b200_finalize();\n", t)
        # 5. dispatch (:162-184)
        t = sub_once(r"(\n  case LSBENCH_SOLVER_GINKGO:\n\s*ginkgo_bench\(x, A, r, cb\);\n\s*break;\n)",
                     "\\1  case LSBENCH_SOLVER_B200:\n    b200_bench(x, A, r, cb);\n    break;\n", t)
        return t
    edit(os.path.join(src, "lsbench.c"), lsbench_c)


def build(ref="/root/reference", verbose=False):
    if not os.path.isdir(os.path.join(ref, "src")):
        print("no reference tree at %s: keeping prebuilt oracle/_ref/" % ref)
        return None
    lib_dir = os.path.join(ROOT, "lsbench_b200")
    if not os.path.exists(os.path.join(lib_dir, "libb200.so")):
        raise RuntimeError("build libb200.so first (python -m lsbench_b200.build)")
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory(prefix="lsbench_dropin_") as tmp:
        src = os.path.join(tmp, "src")
        shutil.copytree(os.path.join(ref, "src"), src)
        shutil.copytree(os.path.join(ref, "bin"), os.path.join(tmp, "bin"))
        for root, _, files in os.walk(tmp):
            for f in files:
                os.chmod(os.path.join(root, f), 0o644)
        register(src)
        shutil.copy(os.path.join(ROOT, "lsbench_b200", "host", "b200.c"), os.path.join(src, "b200.c"))
        shutil.copy(os.path.join(ROOT, "include", "b200.h"), os.path.join(src, "b200.h"))
        objs = []

        def run(cmd):
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)
        for f in sorted(os.listdir(src)):
            o = os.path.join(tmp, f + ".o")
            if f == "b200.c":   # our file: the warnings this repository builds with
                run([CC, "-O2", "-fPIC", "-std=gnu11", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
                     "-DLSBENCH_B200",
                     "-I", src, "-c", os.path.join(src, f), "-o", o])
            elif f.endswith(".c"):   # the reference's files: its own flags, minus -Werror
                run([CC, "-O2", "-fPIC", "-std=gnu11", "-w", "-D_GNU_SOURCE", "-DLSBENCH_B200", "-I", src,
                     "-c", os.path.join(src, f), "-o", o])
            elif f.endswith(".cpp"):
                run([CXX, "-O2", "-fPIC", "-std=c++17", "-w", "-I", src, "-c", os.path.join(src, f), "-o", o])
            else:
                continue
            objs.append(o)
        lib = os.path.join(OUT, "libref_lsbench_b200.so")
        rpath_lib = "-Wl,-rpath,$ORIGIN/../../lsbench_b200"
        run([CXX, "-shared", "-o", lib] + objs + ["-L", lib_dir, "-lb200", "-lpthread", rpath_lib])
        drv = os.path.join(OUT, "driver_ref_b200")
        run([CC, "-O2", "-w", "-I", src, os.path.join(tmp, "bin", "driver.c"), "-o", drv, "-L", OUT,
             "-lref_lsbench_b200", "-Wl,-rpath,$ORIGIN", rpath_lib])
    return drv


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    a = ap.parse_args()
    print(build(a.ref, verbose=True))
