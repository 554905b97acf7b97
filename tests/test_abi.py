"""The product C-ABI library loads without a GPU and exports every symbol include/macjd.h
declares; the ctypes structures match the library's struct sizes.  No compute calls."""
import ctypes
import os
import re

import pytest

from tests.helpers import ROOT


def declared_symbols():
    with open(os.path.join(ROOT, "include", "macjd.h")) as f:
        text = f.read()
    return re.findall(r"MACJD_API\s+[\w\s\*]+?\b(macjd_\w+)\s*\(", text)


def test_header_declares_entry_points():
    syms = declared_symbols()
    assert {"macjd_env_step", "macjd_env_reset", "macjd_agent_forward"} <= set(syms)
    assert len(syms) == len(set(syms))


def test_product_library_exports_all_symbols():
    import __graft_entry__ as ge
    path = ge.build()
    lib = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/macjd.h but not exported"


def test_binding_matches_library_struct_sizes():
    import __graft_entry__ as ge
    from macjd_b200 import _native as N
    lib = N.NativeLib(ge.build())           # raises on any size / version mismatch
    assert set(N.NativeLib.SIGNATURES) <= set(declared_symbols())
    assert lib.lib.macjd_status_string(-2).decode() == "unsupported dimensions"


def test_missing_library_fails_loudly(tmp_path):
    from macjd_b200 import _native as N
    with pytest.raises(N.MacjdError, match="no CPU fallback"):
        N.NativeLib(str(tmp_path / "libmacjd_b200.so"))


def test_product_classes_refuse_cpu():
    """Without the host-emulation test library the classes never run on the CPU."""
    import types
    from macjd_b200 import _native as N
    from macjd_b200.simulation.environment import ElectromagneticEnvironment
    from macjd_b200.simulation.scenario import default_spec
    with pytest.raises(N.MacjdError, match="CUDA only"):
        ElectromagneticEnvironment(types.SimpleNamespace(), spec=default_spec(2), device="cpu")


def test_product_library_exports_nothing_undeclared():
    """Development aids (phase stamps, tensor-pipe micro-benchmarks) live in the tooling build only."""
    import subprocess
    import __graft_entry__ as ge
    out = subprocess.run(["nm", "-D", "--defined-only", ge.build()], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("macjd_")}
    assert exported == set(declared_symbols()), exported ^ set(declared_symbols())


def test_build_records_what_it_compiled():
    import json
    import __graft_entry__ as ge
    lib = ge.build()
    with open(lib + ".build_info.json") as f:
        info = json.load(f)
    assert info["compiled"] is True and "arch=compute_100a,code=sm_100a" in info["flags"]
