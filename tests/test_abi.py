"""The C-ABI shared library loads (no GPU needed) and exports every symbol include/b200exec.h declares."""
import os
import re

import ballista_b200 as bb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "b200exec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = bb.engine.load_library()
    fns = header_functions()
    assert len(fns) >= 20
    missing = [f for f in fns if not hasattr(L, f)]
    assert not missing, missing


def test_python_binding_lists_the_same_symbols():
    assert sorted(bb.engine.EXPORTED_SYMBOLS) == header_functions()


def test_version_string():
    assert b"sm_100a" in bb.engine.load_library().b200_version()


def test_struct_layouts_match_header():
    import ctypes as C
    assert C.sizeof(bb.engine.ShuffleWritePartition) == 48   # 4*u64 + i64 + 2*i32
    assert C.sizeof(bb.engine.OperatorMetrics) == 48 + 6 * 8
    assert C.sizeof(bb.engine.DeviceBuffer) == 16


def test_engine_requires_a_gpu_no_cpu_fallback():
    """Without a CUDA device the product must fail loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        bb.GpuExecutionEngine(0)
    except bb.B200Error as e:
        assert e.code == -4
    else:
        raise AssertionError("engine creation must fail without a GPU")
