"""The C-ABI shared library loads and exports every symbol include/pmc_b200.h declares (no compute calls here:
they need a GPU), and fails loudly -- not silently on a CPU path -- when no device is present."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "pmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(pmc_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_and_exports_header_symbols():
    from parelagmc_b200 import capi
    capi.build()
    lib = capi.load()
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pmc_b200.h but not exported"
    assert sorted(capi.SYMBOLS) == syms, "capi.SYMBOLS out of sync with include/pmc_b200.h"


def test_library_is_sm100a_cuda_not_a_cpu_stub():
    """The .so must carry sm_100a device code for the hot kernels."""
    import subprocess
    from parelagmc_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "k_run_program" in sass and "k_rng" in sass
    assert "LDG.E.128" in sass, "128-bit vector loads expected in the batched kernels"


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from parelagmc_b200 import capi
    with pytest.raises(capi.PmcError) as e:
        capi.Context(3)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "parelagmc_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dp, f)).read()
                for pat in ("import oracle", "from oracle", "libpmc_oracle", "pmc_oracle.h", "po_create", "oracle.binding"):
                    assert pat not in txt, f"{f} reaches into the oracle ({pat})"
