"""The C ABI exercised by a caller that is not Python: tests/c_abi_smoke.c is compiled with gcc against
include/climaseaice_b200.h and binds the library with dlopen/dlsym, as a Julia ccall / cgo / JNI host would."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "climaseaice.jl_b200" / "libclimaseaice_b200.so"
FIXTURE = ROOT / "tests" / "golden" / "c_abi_periodic_40x24.bin"


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    out = tmp_path_factory.mktemp("cabi") / "c_abi_smoke"
    subprocess.run(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "tests" / "c_abi_smoke.c"), "-ldl", "-o", str(out)],
                   check=True)
    return out


def test_header_compiles_as_c_and_every_symbol_resolves(exe):
    """The public header is valid C11 (-Wall -Wextra -Werror) and the library exports everything it declares."""
    assert LIB.exists(), "build the library first (__graft_entry__.build())"
    r = subprocess.run([str(exe), str(LIB), "--symbols"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "C_ABI_SYMBOLS_OK" in r.stdout


def test_symbol_list_of_the_c_caller_matches_the_header():
    import re
    hdr = (ROOT / "include" / "climaseaice_b200.h").read_text()
    declared = set(re.findall(r"\b(csi_[a-z0-9_]+)\s*\(", hdr)) - {"csi_array", "csi_config", "csi_fields"}
    src = (ROOT / "tests" / "c_abi_smoke.c").read_text()
    listed = set(re.findall(r'"(csi_[a-z0-9_]+)"', src))
    assert declared == listed, (declared ^ listed)


@pytest.mark.gpu
def test_c_caller_reproduces_the_fixture(exe):
    """csi_create -> csi_evp_substeps_host (host buffers) -> csi_destroy from C, bit-identical to the oracle's fixture."""
    r = subprocess.run([str(exe), str(LIB), str(FIXTURE)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "C_ABI_SMOKE_OK" in r.stdout
