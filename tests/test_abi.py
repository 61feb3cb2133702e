"""CPU-side checks of the C-ABI boundary: the library builds for sm_100a, loads
without a GPU or driver, and exports every symbol include/flash_attn_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fa():
    import flash_attention_metal_b200 as fa

    if not os.path.exists(fa.LIB_PATH):
        fa.build()
    return fa


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "flash_attn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b([a-z_0-9]+)\s*\(", text)) - {"defined"})


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "flash_attn_b200.h"\nint main(void){return FA_OK;}\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_library_exports_every_declared_symbol(fa):
    lib = ctypes.CDLL(fa.LIB_PATH)
    names = _declared_symbols()
    assert "flash_attention_v4_half" in names and "flash_attention_backward" in names
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(fa.EXPORTS) == names


def test_library_has_no_driver_or_torch_link_dependency(fa):
    out = subprocess.check_output(["ldd", fa.LIB_PATH]).decode()
    assert "libcuda.so" not in out and "torch" not in out and "not found" not in out


def test_version_and_error_paths_without_gpu(fa):
    assert fa.version() >= 100
    # argument validation happens before any CUDA call
    with pytest.raises(fa.FlashAttnError, match="D must be 64 or 128"):
        fa.naive_attention(16, 16, 16, 16, 128, 48, 0.125)
    with pytest.raises(fa.FlashAttnError, match="null tensor"):
        fa.flash_attention_v4_half(0, 16, 16, 16, 128, 64, 0.125, 8192, 8192, None, False)
    with pytest.raises(fa.FlashAttnError, match="16-byte aligned"):
        fa.flash_attention_v2(8, 16, 16, 16, 128, 64, 0.125)
    assert fa.workspace_bytes_backward(128, 64, 2, 3) >= 2 * 3 * 128 * 4


def test_product_package_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flash_attention_metal_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), (dirpath, f)


def test_ctypes_signatures_match_the_header(fa):
    """Every prototype in the header has as many parameters as the ctypes binding declares."""
    text = open(os.path.join(ROOT, "include", "flash_attn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    lib = fa.lib()
    checked = 0
    for m in re.finditer(r"\b([a-z_0-9]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        if not hasattr(lib, name):
            continue
        n = 0 if params in ("", "void") else params.count(",") + 1
        argtypes = getattr(lib, name).argtypes
        if argtypes is not None:
            assert len(argtypes) == n, (name, len(argtypes), n)
            checked += 1
    assert checked >= 15
