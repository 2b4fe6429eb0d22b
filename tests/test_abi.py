"""The C-ABI library: loads, exports every declared symbol, validates arguments.  CPU only
(no compute call succeeds without a GPU -- and that failure must be loud)."""
import ctypes
import os
import re

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "gsb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for must in ("gsb_summate", "gsb_summate_incompr", "gsb_summate_structured",
                 "gsb_summate_incompr_structured", "gsb_last_error", "gsb_version"):
        assert must in syms


def test_library_exports_every_declared_symbol(gsb):
    lib = ctypes.CDLL(gsb._lib.lib_path())
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/gsb200.h but not exported"


def test_binding_table_covers_header(gsb):
    assert sorted(gsb._lib.SIGNATURES) == declared_symbols()


def test_version_and_device_count(gsb):
    assert gsb._lib.load().gsb_version() >= 100
    assert gsb.device_count() >= 0


def test_argument_errors_need_no_gpu(gsb):
    lib = gsb._lib.load()
    a = np.zeros(8)
    p = a.ctypes.data
    # dim out of range
    assert lib.gsb_summate(p, p, p, p, 1, 0, 1, 1, p, 0, 0, None) == 1
    assert b"dim" in lib.gsb_last_error()
    # negative sizes
    assert lib.gsb_summate(p, p, p, p, 1, 2, -1, 1, p, 0, 0, None) == 1
    assert lib.gsb_summate(p, p, p, p, 1, 2, 1, -1, p, 0, 0, None) == 1
    # NULL buffers
    assert lib.gsb_summate(None, p, p, p, 1, 2, 1, 1, p, 0, 0, None) == 1
    # leading dimension too small
    assert lib.gsb_summate(p, p, p, p, 1, 2, 1, 4, p, 0, 0, None) == 1
    # vector field dims (generator.py:514-517)
    assert lib.gsb_summate_incompr(p, p, p, p, 1, 1, 1, 1, p, 1, 0, 0, None) == 1
    # bad mem kind
    assert lib.gsb_summate(p, p, p, p, 1, 2, 1, 1, p, 7, 0, None) == 1
    # unknown option
    assert lib.gsb_set_option(b"nope", 1) == 1
    assert lib.gsb_get_counter(b"nope") == -1


def test_empty_inputs_return_without_device(gsb):
    out = gsb.summate(np.zeros((2, 4)), np.zeros(4), np.zeros(4), np.zeros((2, 0)))
    assert out.shape == (0,)
    out = gsb.summate_incompr(np.zeros((3, 4)), np.zeros(4), np.zeros(4), np.zeros((3, 0)))
    assert out.shape == (3, 0)
    out = gsb.summate_structured(np.zeros((2, 4)), np.zeros(4), np.zeros(4),
                                 [np.zeros(0), np.zeros(3)])
    assert out.shape == (0, 3)


def test_no_cpu_fallback_is_loud(gsb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gsb.GSB200Error, match="no CPU fallback"):
        gsb.summate(np.zeros((2, 4)), np.zeros(4), np.zeros(4), np.zeros((2, 3)))
    with pytest.raises(gsb.GSB200Error):
        gsb.summate_structured(np.zeros((2, 4)), np.zeros(4), np.zeros(4),
                               [np.arange(3.0), np.arange(3.0)])


def test_product_package_never_imports_oracle():
    pkg = os.path.join(REPO, "gstools_b200")
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "liboracle" not in text, f


def test_krige_argument_errors_need_no_gpu(gsb):
    lib = gsb._lib.load()
    a = np.zeros(16)
    p = a.ctypes.data
    assert lib.gsb_calc_field_krige_and_variance(p, p, 4, p, -1, 4, p, p, 0, 0, None) == 1
    assert lib.gsb_calc_field_krige_and_variance(p, p, 4, p, 2, -4, p, p, 0, 0, None) == 1
    assert lib.gsb_calc_field_krige_and_variance(p, p, 2, p, 2, 4, p, p, 0, 0, None) == 1   # ld < n
    assert lib.gsb_calc_field_krige_and_variance(p, p, 4, p, 2, 4, p, None, 0, 0, None) == 1
    assert lib.gsb_calc_field_krige(None, p, 4, p, 2, 4, p, 0, 0, None) == 1
    assert lib.gsb_calc_field_krige(p, p, 4, p, 2, 4, p, 9, 0, None) == 1
    # n == 0 returns without touching a device
    f, e = gsb.calc_field_krige_and_variance(np.zeros((3, 3)), np.zeros((3, 0)), np.zeros(3))
    assert f.shape == (0,) and e.shape == (0,)


def _build_c_consumer(tmp_path, gsb):
    import shutil
    import subprocess

    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no C compiler")
    lib = gsb._lib.lib_path()
    exe = str(tmp_path / "abi_smoke")
    cmd = [cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(REPO, "include"),
           os.path.join(REPO, "tests", "c_abi", "abi_smoke.c"), "-o", exe, lib, "-lm",
           "-Wl,-rpath," + os.path.dirname(lib)]
    subprocess.check_call(cmd)
    return exe


def test_header_is_plain_c_and_library_links_from_c(gsb, tmp_path):
    """include/gsb200.h compiles as strict C99 and a C program can call the library (argument errors
    need no device)."""
    import subprocess

    exe = _build_c_consumer(tmp_path, gsb)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "abi ok" in out.stdout, out.stderr


def test_plan_share_is_the_sharding_rule_of_dist(gsb):
    """gsb_plan_share (single-process multi-GPU plan) cuts exactly like dist.shard_range (one process per GPU)."""
    from gstools_b200 import dist

    for n in (0, 1, 7, 512, 1000003):
        for parts in (1, 2, 3, 8):
            got = [gsb._lib.plan_share(n, parts, g) for g in range(parts)]
            assert got == [dist.shard_range(n, g, parts) for g in range(parts)]
            assert got[0][0] == 0 and got[-1][1] == n
    lib = gsb._lib.load()
    lo, hi = ctypes.c_int64(), ctypes.c_int64()
    assert lib.gsb_plan_share(10, 0, 0, ctypes.byref(lo), ctypes.byref(hi)) == 1
    assert lib.gsb_plan_share(10, 2, 2, ctypes.byref(lo), ctypes.byref(hi)) == 1


def test_plan_needs_a_gpu(gsb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    with pytest.raises(gsb.GSB200Error, match="no CPU fallback"):
        gsb.Plan()
    lib = gsb._lib.load()
    assert lib.gsb_plan_destroy(None) == 0
    assert lib.gsb_plan_summate(None, None, None, None, None, 0, 1, 0, 0, None, 0, 0, None, None, 0, 0, None) == 1
