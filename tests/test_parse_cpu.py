"""The record parser the device uses (lsbench_b200/csrc/parse.cuh) is plain
host/device C++: compile it for the host and hold it against libc.  Whenever
it accepts a line, row / col / value must equal (unsigned)strtoul / strtod bit
for bit; what it does not accept it must hand back, never guess
(src/lsbench-csr.c:49-53 semantics)."""
import os
import re
import subprocess

import pytest

import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("parse") / "parse_check")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-x", "c++", "-I", os.path.join(ROOT, "lsbench_b200", "csrc"),
                    os.path.join(ROOT, "tests", "parse_check.cpp"), "-o", exe], check=True)
    return exe


def run(exe, *args):
    r = subprocess.run([exe] + list(args), capture_output=True, text=True, timeout=600)
    m = re.search(r"accepted (\d+) of (\d+), mismatches (\d+)", r.stdout)
    assert m, r.stdout + r.stderr
    return r.returncode, int(m.group(1)), int(m.group(2)), int(m.group(3))


def test_every_reference_record(checker):
    rc, acc, tot, bad = run(checker, *[orc.matrix_path(n) for n in orc.TOY + orc.NEK])
    assert rc == 0 and bad == 0
    assert tot == 4 + 4 + 5 + 138756 + 110153 + 91095 + 145538 + 119656 + 94283 + 76591
    assert acc >= 0.9999 * tot        # the Nek files are on the exact fast path


def test_random_records_in_many_formats(checker):
    rc, acc, tot, bad = run(checker, "--random", "1000000")
    assert rc == 0 and bad == 0
    assert 0.5 * tot < acc < tot      # %.17g and large exponents must be handed back
