"""The C-ABI library loads without a GPU and exports exactly the symbols include/b200dit.h declares;
compute entry points fail loudly (no CPU fallback)."""
import os
import re
import subprocess

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    if not os.path.isfile(os.path.join(ROOT, "omnihuman-1-hack_b200", "libb200dit.so")):
        g.build()
    import b200dit
    return b200dit


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b200dit.h")).read()
    return set(re.findall(r"B200_API[^;(]*?\b(b200[a-z0-9_]*)\s*\(", src))


def test_header_symbols_exported(built):
    from b200dit import _lib
    syms = _header_symbols()
    assert len(syms) >= 19
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert syms <= exported, syms - exported
    assert {s for s in exported if s.startswith("b200")} == syms          # nothing undeclared leaks out
    assert set(_lib.SIGNATURES) == syms                                   # the ctypes table covers the header


def test_library_has_no_torch_or_libcuda_dependency(built):
    from b200dit import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    assert "torch" not in out and "libcuda.so" not in out


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.B200Error):
        built.DitEngine(dim=128, ffn_dim=256, num_heads=1, num_layers=1, text_dim=32)
    from b200dit import _lib
    lib = _lib.lib()
    assert lib.b200_version().startswith(b"b200dit")
    assert lib.b200_kernel_launches() == 0


def test_hard_limits_fail_loudly_on_the_host(built):
    """Limits of the engine that the reference does not have are refused with a message, not mis-computed:
    head_dim must be 128 (attention.py:54 allows up to 256), text_len a multiple of 8, freq_dim a multiple of 4.
    b200dit_weight_names runs the same configuration check as b200dit_create and needs no GPU."""
    for bad, needle in ((dict(dim=1536, num_heads=16), "head_dim"), (dict(dim=1536, num_heads=11), "divisible"),
                        (dict(text_len=500), "text_len"), (dict(freq_dim=254), "freq_dim")):
        with pytest.raises(built.B200Error) as e:
            built.DitEngine.expected_weight_names(**bad)
        assert needle in str(e.value), (bad, str(e.value))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "omnihuman-1-hack_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                assert "oracle" not in open(os.path.join(dp, f)).read().replace("checked against the oracle", ""), f
