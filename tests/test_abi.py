"""CPU tests: the C-ABI library builds for sm_100a, loads, exports every symbol include/mepb.h declares, and fails
loudly (no CPU fallback) when no CUDA device is present.  No compute calls here."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "mepb.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mepb_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_documented_entry_points():
    names = declared_symbols()
    for must in ("mepb_create", "mepb_destroy", "mepb_run_mesh", "mepb_accumulate", "mepb_finalize", "mepb_interp",
                 "mepb_fourier", "mepb_get_partials", "mepb_set_partials", "mepb_partials_device_ptr", "mepb_strerror"):
        assert must in names


def test_library_exports_every_declared_symbol(libmepb):
    L = C.CDLL(libmepb.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(L, name), f"{name} declared in include/mepb.h but not exported by libmepb.so"
    assert sorted(libmepb.EXPORTS) == [n for n in declared_symbols()]
    L.mepb_version.restype = C.c_char_p
    assert b"sm_100a" in L.mepb_version()


def test_library_is_sm100a_only_and_does_not_link_the_oracle(libmepb):
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", libmepb.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    ldd = subprocess.run(["ldd", libmepb.LIB_PATH], capture_output=True, text=True).stdout
    assert "cpu_ref" not in ldd and "openblas" not in ldd.lower()
    nm = subprocess.run(["nm", "-D", libmepb.LIB_PATH], capture_output=True, text=True).stdout
    assert "cpu_ref" not in nm and "zheevd" not in nm


def test_struct_layout_matches_header(libmepb):
    """ctypes mirrors of mepb_model / mepb_params / mepb_result have the C sizes (checked with gcc)."""
    prog = r'''
#include <stdio.h>
#include "mepb.h"
int main(void){ printf("%zu %zu %zu\n", sizeof(mepb_model), sizeof(mepb_params), sizeof(mepb_result)); return 0; }
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(prog)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "s.c"), "-o", os.path.join(d, "s")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(d, "s")]).split()]
    from mepinterp_b200 import api
    assert sizes == [C.sizeof(api.CModel), C.sizeof(api.CParams), C.sizeof(api.CResult)]
    from oracle import cpu_ref
    assert sizes == [C.sizeof(cpu_ref.CModel), C.sizeof(cpu_ref.CParams), C.sizeof(cpu_ref.CResult)]


def test_no_gpu_fails_loudly(libmepb):
    """Without a device mepb_create must return MEPB_ERR_CUDA — there is no CPU path to fall back to."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    from tests.cases import synthetic_model
    from oracle import mep_oracle as mo
    m = synthetic_model(4, 1)
    with pytest.raises(libmepb.MepbError) as ei:
        libmepb.Context(m, mo.Params(mp_grid=(2, 2, 2), do_ahc=True))
    assert ei.value.code == libmepb.api.ERR_CUDA
    assert "no CPU path" in str(ei.value) or "CUDA" in str(ei.value)


def test_argument_errors_do_not_need_a_gpu(libmepb):
    L = libmepb.lib()
    h = C.c_void_p()
    assert L.mepb_create(None, None, 0, C.byref(h)) == libmepb.api.ERR_ARG
    assert L.mepb_strerror(libmepb.api.ERR_ARG) == b"invalid argument"
    assert L.mepb_partials_size(None) == 0
    assert L.mepb_finalize(None, None) == libmepb.api.ERR_ARG
    L.mepb_destroy(None)     # no-op
