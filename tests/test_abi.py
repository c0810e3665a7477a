"""The C-ABI library builds, loads and exports every symbol include/mss.h declares (no compute without a GPU)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mss.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mss_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from ms_slam_b200 import engine
    assert declared_symbols() == sorted(engine.SYMBOLS)


def test_library_exports_every_declared_symbol(build_native):
    lib = build_native
    for name in declared_symbols():
        assert hasattr(lib, name), f"libmss.so does not export {name}"
    assert lib.mss_version() == 210


def test_struct_sizes_match_header(build_native):
    # sizes the C compiler sees for the ABI structs (guards against ctypes/struct drift)
    import subprocess, tempfile
    from ms_slam_b200 import engine
    src = '#include "mss.h"\n#include <stdio.h>\nint main(){printf("%zu %zu %zu %zu\\n", sizeof(mss_config), ' \
          'sizeof(mss_window_view), sizeof(mss_result), sizeof(mss_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "t"), os.path.join(d, "t.c")],
                       check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    sizes = [int(x) for x in out]
    assert sizes == [C.sizeof(engine.mss_config), C.sizeof(engine.mss_window_view), C.sizeof(engine.mss_result),
                     C.sizeof(engine.mss_stats)]


def test_no_cpu_fallback(build_native):
    """Without a CUDA device the product path fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ms_slam_b200.engine import Engine, MssError, MSS_E_CUDA
    with pytest.raises(MssError) as ei:
        Engine()
    assert ei.value.status == MSS_E_CUDA


def test_bad_arguments_rejected_without_gpu(build_native):
    lib = build_native
    from ms_slam_b200 import engine
    assert lib.mss_create(None, None) == engine.MSS_E_BADARG
    assert lib.mss_solve(None, None, None) == engine.MSS_E_BADARG
    assert lib.mss_last_error(None) == b"null handle"


def test_product_code_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "ms_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, os.path.join(dirpath, f)


def test_host_mirror_library_exports(build_native):
    """libmss_host.so (C++ MapSparsification mirror + harness) builds, loads next to libmss.so and exports its harness."""
    from ms_slam_b200 import host_mirror
    lib = host_mirror.load_library()
    for name in host_mirror.SYMBOLS:
        assert hasattr(lib, name), f"libmss_host.so does not export {name}"
