"""The C-ABI library loads and exports every symbol include/flechasdb_b200.h declares (no GPU)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, "include", "flechasdb_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fdb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from flechasdb_b200 import _capi as capi
    L = capi.lib()
    names = header_functions()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), "libflechasdb_b200.so does not export %s" % n
    # and the ctypes table covers exactly the header
    assert sorted(capi.SIGNATURES) == names


def test_rust_bindings_are_generated_from_the_header():
    """rust_shim/src/ffi.rs declares every function of the header, with the same number of arguments: the committed
    file is exactly what tools/gen_ffi.py prints for the current header."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ffi", os.path.join(ROOT, "tools", "gen_ffi.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    text, funcs = gen.generate()
    assert sorted(f[0] for f in funcs) == header_functions()
    committed = open(os.path.join(ROOT, "rust_shim", "src", "ffi.rs")).read()
    assert committed == text, "rust_shim/src/ffi.rs is stale: run python tools/gen_ffi.py --write"
    from flechasdb_b200 import _capi as capi
    for name, ret, params in funcs:
        assert len(params) == len(capi.SIGNATURES[name][1]), name


def test_error_codes_match_header():
    from flechasdb_b200 import _capi as capi
    text = open(os.path.join(ROOT, "include", "flechasdb_b200.h")).read()
    for name, val in re.findall(r"#define (FDB_ERR_[A-Z_]+) \((-\d+)\)", text):
        assert getattr(capi, name[4:]) == int(val)


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the engine must fail loudly, not compute on the CPU."""
    from flechasdb_b200 import _capi as capi
    from flechasdb_b200.engine import Context
    if capi.lib().fdb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.FdbError) as e:
        Context(0)
    assert e.value.code == capi.ERR_CUDA


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under flechasdb_b200/ may reference it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "flechasdb_b200")):
        if os.path.basename(dirpath) in ("build", "__pycache__"):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                t = open(os.path.join(dirpath, f), errors="replace").read()
                if re.search(r"oracle|fo_[a-z_]+\(", t):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
