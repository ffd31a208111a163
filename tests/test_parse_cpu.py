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


def test_values_next_to_rounding_boundaries(checker, tmp_path):
    """17-19 digit decimals at and one decimal ulp either side of the midpoint of
    two adjacent doubles: the exact integer path (product / long division with a
    sticky bit, round half to even) must agree with strtod on every accepted one."""
    import random
    from decimal import Decimal, getcontext
    getcontext().prec = 60
    random.seed(7)
    lines = []
    for _ in range(40000):
        e = random.randint(-60, 100)
        m = random.getrandbits(52) | (1 << 52)
        mid = (Decimal(m) + Decimal(m + 1)) / 2 * (Decimal(2) ** (e - 52))
        nd = random.choice([17, 18, 19])
        s = format(mid, ".%de" % (nd - 1))
        d = Decimal(s)
        step = Decimal(1).scaleb(d.adjusted() - (nd - 1))
        for v in (d, d + step, d - step):
            lines.append("1 2 %s\n" % format(v, ".%de" % (nd - 1)))
    f = tmp_path / "halfway.txt"
    f.write_text("%d 1\n" % len(lines) + "".join(lines))
    rc, acc, tot, bad = run(checker, str(f))
    assert rc == 0 and bad == 0 and tot == len(lines)
    assert acc > 0.7 * tot            # exponents within +-27 after scaling: most are taken
