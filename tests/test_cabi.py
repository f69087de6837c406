"""The C-ABI shared library loads and exports every symbol include/enzymm_b200.h declares.
No compute calls here: without a GPU every entry point must fail loudly, never fall back."""
import ctypes
import re

import numpy as np
import pytest

from conftest import ROOT
from enzymm_b200.engine import Engine, EngineError, HIT_DTYPE, library_path, load_cdll
from enzymm_b200.library import CompiledLibrary


def declared_symbols():
    text = (ROOT / "include" / "enzymm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(emm_[a-z_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    lib = ctypes.CDLL(str(library_path()))
    names = declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_abi_constants():
    lib = load_cdll()
    assert lib.emm_abi_version() == 2
    assert lib.emm_hit_size() == HIT_DTYPE.itemsize == 280


def test_product_path_never_touches_the_oracle():
    """No import, include, link or dlopen of anything under oracle/ from the product package."""
    pkg = ROOT / "enzymm_b200"
    for path in list(pkg.glob("*.py")) + list((pkg / "csrc").glob("*")):
        text = path.read_text() if path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h") or path.name == "Makefile" else ""
        assert not re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M), path
        assert not re.search(r"#include\s*[\"<][^\">]*oracle", text), path
        assert "libjess_oracle" not in text, path


@pytest.mark.skipif(load_cdll().emm_device_count() > 0, reason="only meaningful without a GPU")
def test_no_cpu_fallback(active_templates):
    with pytest.raises(EngineError) as info:
        Engine(CompiledLibrary(active_templates[:3], 2.0, 1.5, 1.5))
    assert info.value.status == -3


@pytest.mark.skipif(load_cdll().emm_device_count() > 0, reason="only meaningful without a GPU")
def test_stream_entry_point_without_gpu():
    lib = load_cdll()
    stream = ctypes.c_void_p()
    assert lib.emm_stream_create(ctypes.c_int(0), ctypes.byref(stream)) == -3
    assert not stream.value
    assert lib.emm_stream_destroy(ctypes.c_int(0), ctypes.c_void_p()) == 0


def test_library_is_sm100a_with_tma_staging():
    """The shipped shared library holds sm_100a SASS and the search kernel stages blobs with the
    TMA bulk copy (UBLKCP + mbarrier transaction waits), not with per-thread loads."""
    import shutil
    import subprocess
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not shutil.which(tool):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([tool, "-sass", str(library_path())], capture_output=True, text=True, timeout=300).stdout
    assert "arch = sm_100a" in sass and "arch = sm_90" not in sass
    assert "emm_search_kernel" in sass and "emm_prepare_kernel" in sass
    assert "UBLKCP.S.G" in sass and "SYNCS.ARRIVE.TRANS64" in sass and "SYNCS.PHASECHK.TRANS64" in sass


def test_c99_client_links_and_runs(tmp_path):
    """The header is plain C (no C++ or torch types in any signature): a C99 translation unit
    includes it with -pedantic, links the shared library and gets sane constants, the loud no-device
    error (on a box without a GPU) and the host-only PDB reader's atom count for the 1AMY fixture."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    exe = tmp_path / "abi_client"
    libdir = library_path().parent
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", f"-I{ROOT / 'include'}", "-o", str(exe),
                    str(ROOT / "tests" / "c" / "abi_client.c"), f"-L{libdir}", "-lenzymm_b200", f"-Wl,-rpath,{libdir}"],
                   check=True, capture_output=True, text=True, timeout=120)
    run = subprocess.run([str(exe), str(ROOT / "tests" / "golden" / "1AMY.pdb")], capture_output=True, text=True, timeout=120)
    assert run.returncode == 0, run.stderr
    assert "abi=2 hit=280" in run.stdout and "atoms=3339" in run.stdout
