"""The C-ABI library builds for sm_100a, loads without a GPU, exports every symbol include/*.h
declares, and refuses to run without a device (no CPU fallback)."""
import ctypes
import os
import re

import pytest

import mods_b200 as mb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def library():
    mb.build()
    return ctypes.CDLL(mb.LIB_PATH)


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mods_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mb2_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(library):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(library, s), "include/mods_b200.h declares %s but libmods_b200.so does not export it" % s
    assert set(mb.EXPORTS) <= set(syms)


def test_host_mirror_symbols_exported(library):
    """libmods_host.so (the C++ mirror of the reference's plugin surface) exports every extern "C" entry mods_host.hpp declares."""
    text = open(os.path.join(ROOT, "mods_b200", "host", "mods_host.hpp")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    block = text[text.index('extern "C" {'):]
    syms = sorted(set(re.findall(r"\b(mb2_[a-z0-9_]+)\s*\(", block)))
    assert len(syms) >= 10 and "mb2_views_sharded_pairs" in syms and "mb2_mods_pairs" in syms
    host = ctypes.CDLL(mb.HOST_LIB_PATH)
    for s in syms:
        assert hasattr(host, s) or hasattr(library, s), "mods_host.hpp declares %s but neither library exports it" % s


def test_no_cpu_fallback_without_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(mb.Mb2Error):
        mb.Context(0)


def test_sass_is_blackwell_native():
    """tcgen05 / TMA must be in the shipped binary (SASS names from B200_PROFILING.md)."""
    import shutil
    import subprocess
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    mb.build()
    sass = subprocess.run(["cuobjdump", "-sass", mb.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert mnemonic in sass, mnemonic
