"""In-tree build of the C host shell: liblsbench.so (harness + b200 backend)
and the `driver` executable, linked against libb200.so.

    python -m lsbench_b200.build_host

Plain gcc; the CMake route (CMakeLists.txt, -DENABLE_B200=ON) builds the same
targets.  When the reference tree is mounted, its unmodified bin/driver.c is
also compiled against this library (-> host/_build/driver_ref), which is the
"driver.c works unchanged" check.
"""
import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
HOST = os.path.join(PKG, "host")
OUT = os.path.join(HOST, "_build")
LIB = os.path.join(OUT, "liblsbench.so")
DRIVER = os.path.join(OUT, "driver")
DRIVER_REF = os.path.join(OUT, "driver_ref")
REF_DRIVER_SRC = "/root/reference/bin/driver.c"


def _cc():
    for cand in (os.environ.get("HOSTCC"), "/usr/bin/gcc", shutil.which("gcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("gcc not found")


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("%s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def build(enable_b200=True):
    cc = _cc()
    os.makedirs(OUT, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", HOST]
    cflags = ["-O2", "-g", "-std=c11", "-fPIC", "-Wall", "-Wextra", "-Werror"] + inc
    if enable_b200:
        cflags.append("-DLSBENCH_B200")
    srcs = [os.path.join(HOST, f) for f in ("lsbench.c", "lsbench-csr.c", "b200.c")]
    link = ["-L", PKG, "-lb200", "-Wl,-rpath," + PKG, "-lpthread"] if enable_b200 else []
    _run([cc] + cflags + ["-shared", "-o", LIB] + srcs + link)
    rpath = ["-L", OUT, "-llsbench", "-Wl,-rpath," + OUT, "-Wl,-rpath," + PKG]
    _run([cc, "-O2", "-std=c11", "-Wall", "-Wextra", "-Werror"] + inc +
         [os.path.join(HOST, "main.c"), "-o", DRIVER] + rpath)
    if os.path.exists(REF_DRIVER_SRC):
        # the reference's driver, byte for byte, against our lsbench.h + library
        _run([cc, "-O2", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
              REF_DRIVER_SRC, "-o", DRIVER_REF] + rpath)
    return LIB


if __name__ == "__main__":
    print(build())
