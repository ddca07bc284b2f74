"""CPU-only checks of the drop-in boundary: the shared object loads without a GPU, exports every
symbol include/versalign_cuda.h declares plus the reference's four plug-in symbols, and fails
loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from versalignlib_b200 import build, capi
from versalignlib_b200.host import PluginHost, PluginError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _no_gpu():
    try:
        import torch
        return not torch.cuda.is_available()
    except Exception:
        return True


def test_header_and_binding_agree():
    text = open(os.path.join(ROOT, "include", "versalign_cuda.h")).read() + open(os.path.join(ROOT, "include", "versalign_fasta.h")).read()
    declared = sorted(set(re.findall(r"\b(va_(?:cuda|fasta)_[a-z_0-9]+)\s*\(", text)))
    assert declared == sorted(capi.C_ABI_SYMBOLS)


def test_library_exports_every_declared_symbol():
    path = build.build_cuda()
    out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.strip()}
    for sym in capi.C_ABI_SYMBOLS + capi.PLUGIN_SYMBOLS:
        assert sym in exported, sym
    # the interface headers declare these two globals with C++ linkage
    assert "_parameters" in exported and "_logger" in exported


def test_library_loads_and_reports_version():
    L = capi.lib()
    assert L.va_cuda_abi_version() == 2
    for sym in capi.C_ABI_SYMBOLS + capi.PLUGIN_SYMBOLS:
        assert getattr(L, sym) is not None


def test_sass_is_sm100a_with_dpx():
    """The cubin inside the .so targets sm_100a."""
    path = build.build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", path], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback_without_device():
    with pytest.raises(capi.CudaError):
        capi.CudaContext()
    # through the plug-in boundary the constructor must throw like the reference's do
    with pytest.raises(PluginError) as ei:
        PluginHost(capi.library_path(), 10, 10, verbosity=0)
    assert "Cannot instantiate Kernel" in str(ei.value)


def test_plugin_ctor_requires_the_six_keys():
    """DefaultKernel.h:68-81: a missing key -> throw "Cannot instantiate Kernel. Lacking parameters"."""
    for missing in ["score_match", "score_mismatch", "score_gap_read", "score_gap_ref", "read_length", "ref_length"]:
        with pytest.raises(PluginError) as ei:
            PluginHost(capi.library_path(), 10, 10, verbosity=0, omit=(missing,))
        assert "Lacking parameters" in str(ei.value), missing
