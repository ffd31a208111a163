"""CPU-side checks of the drop-in boundary: libb200.so loads, exports every
symbol include/b200.h declares, and fails loudly (no fallback) without a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def abi():
    from lsbench_b200 import abi as m
    m.load()
    return m


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree(abi):
    assert declared_symbols() == sorted(abi.SYMBOLS)


def test_library_exports_every_declared_symbol(abi):
    L = ctypes.CDLL(abi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert abi.load().b200_abi_version() == 5


def test_struct_sizes_match_header(abi, tmp_path):
    """sizes and the offsets of the newest members, as the C compiler sees
    include/b200.h, against the ctypes mirror"""
    import subprocess
    src = tmp_path / "sizes.c"
    src.write_text(
        '#include "b200.h"\n#include <stdio.h>\n#include <stddef.h>\n'
        'int main(void) { printf("%zu %zu %zu %zu %zu\\n", sizeof(b200_mat_info), sizeof(b200_pcg_opts),'
        ' sizeof(b200_pcg_result), offsetof(b200_mat_info, values_f32), offsetof(b200_pcg_result, outer_iters));'
        ' return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.run(["/usr/bin/gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [ctypes.sizeof(abi.MatInfo), ctypes.sizeof(abi.PcgOpts), ctypes.sizeof(abi.PcgResult),
                   abi.MatInfo.values_f32.offset, abi.PcgResult.outer_iters.offset]
    # b200_mat_info: 16 u64 + 24 u64 hist + u64 + 2 u32 + 3 u64 + 2 u32
    assert got[:3] == [8 * (16 + 24 + 1) + 8 + 8 * 3 + 8, 24, 72]


def test_product_does_not_reach_into_the_oracle():
    pkg = os.path.join(ROOT, "lsbench_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".cmake", ".txt")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle/" not in text and "liboracle" not in text and "import orc" not in text, f


def test_no_gpu_means_an_error_not_a_fallback(abi):
    try:
        n = abi.device_count()
    except abi.B200Error as e:
        assert e.code == 2  # B200_ECUDA
        with pytest.raises(abi.B200Error):
            abi.Context(0)
        return
    if n == 0:
        with pytest.raises(abi.B200Error):
            abi.Context(0)
