"""The C-ABI library builds, loads, and exports every symbol include/deephall_b200.h declares
(no compute call is made: there is no GPU on the CPU test box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "deephall_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dh_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = header_functions()
    for need in ("dh_logpsi", "dh_local_energy", "dh_mcmc_sweep", "dh_logpsi_vjp", "dh_slogdet", "dh_potential",
                 "dh_plan_create", "dh_param_layout", "dh_workspace_bytes"):
        assert need in names


def test_library_exports_every_declared_symbol():
    from deephall_b200 import _native

    if not os.path.exists(_native.lib_path()):
        import __graft_entry__

        __graft_entry__.build()
    lib = _native.load()
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _native.SIGNATURES, f"{name} has no ctypes signature"
    assert b"sm_100a" in lib.dh_version()


def test_no_cpu_fallback():
    import torch

    from deephall_b200 import _native

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(_native.NativeError):
        _native.Plan(nspins=(3, 0), flux=2)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "deephall_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
