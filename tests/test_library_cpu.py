"""CPU tests of the product library: it loads, exports every symbol of include/climaseaice_b200.h,
its host-side numerics helpers are exact, and it fails loudly without a GPU (no fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import climaseaice_b200 as csi
from climaseaice_b200 import _lib as L
from oracle import oracle as O

ROOT = Path(__file__).resolve().parent.parent


def test_every_declared_symbol_is_exported():
    header = (ROOT / "include" / "climaseaice_b200.h").read_text()
    declared = set(re.findall(r"\b(csi_[a-z0-9_]+)\s*\(", header))
    lib = csi.lib()
    assert declared == set(L.EXPORTS), declared ^ set(L.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym
    assert lib.csi_version() == L.ABI_VERSION


def test_struct_layout_matches_header():
    # sizes computed by hand from include/climaseaice_b200.h (natural alignment)
    assert C.sizeof(L.csi_array) == 24
    assert C.sizeof(L.csi_fields) == 24 * len(L.FIELD_NAMES)
    assert C.sizeof(L.csi_config) == 8 + 24 + 16 + 8 + 56 + 8 + 24 + 8 + 24 + 8 + 40 + 8 + 8 + 8 + 16 + 16 + 8 + 96


def test_correctly_rounded_exp_matches_binary128():
    rng = np.random.default_rng(3)
    xs = np.concatenate([-20 * rng.uniform(0, 1, 20000), rng.uniform(-700, 700, 5000), [0.0, -0.0, 1.0, -1.0, -20.0, 709.0, -744.0]])
    lib = csi.lib()
    for x in xs:
        assert lib.csi_host_exp(float(x)) == O.exp_cr(float(x)), x
    assert np.isnan(lib.csi_host_exp(float("nan")))
    assert lib.csi_host_exp(1000.0) == np.inf and lib.csi_host_exp(-1000.0) == 0.0


def test_constant_division_is_bit_exact():
    """The Markstein quotient used for reused divisors equals the IEEE quotient."""
    rng = np.random.default_rng(4)
    lib = csi.lib()
    divisors = [4000.0, 16e6, 1250.0, 3.0, 0.3 * 900, 300.0, 50.0, 173.2, 2e-9, 1 - 2 ** -53, 2 - 2 ** -52]
    divisors += list(np.ldexp(0.5 + rng.uniform(0, 0.5, 300), rng.integers(-30, 30, 300)))
    for d in divisors:
        xs = np.concatenate([rng.normal(0, 1, 200) * 10.0 ** rng.integers(-12, 12, 200), [0.0, -0.0, d, -d, 1e-300, 1e300]])
        for x in xs:
            got, want = lib.csi_host_div_by_const(float(x), float(d)), float(x) / float(d)
            assert got == want and np.signbit(got) == np.signbit(want), (x, d)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    cfg = L.csi_config()
    cfg.abi_version, cfg.Nx, cfg.Ny, cfg.Hx, cfg.Hy, cfg.dx, cfg.dy, cfg.substeps = 1, 8, 8, 4, 4, 1.0, 1.0, 1
    h = C.c_void_p()
    rc = csi.lib().csi_create(C.byref(cfg), C.byref(h))
    assert rc == -4 and not h.value
    assert b"no CUDA device" in csi.lib().csi_last_error(None)


def test_create_rejects_bad_arguments():
    lib = csi.lib()
    h = C.c_void_p()
    assert lib.csi_create(None, C.byref(h)) == -1
    cfg = L.csi_config()
    cfg.abi_version = 99
    assert lib.csi_create(C.byref(cfg), C.byref(h)) == -1 and b"abi_version" in lib.csi_last_error(None)
    cfg.abi_version, cfg.Nx, cfg.Ny, cfg.Hx, cfg.Hy, cfg.dx, cfg.dy, cfg.substeps = 1, 8, 8, 2, 2, 1.0, 1.0, 1
    assert lib.csi_create(C.byref(cfg), C.byref(h)) == -1 and b"halos" in lib.csi_last_error(None)
    cfg.Hx = cfg.Hy = 4
    cfg.advection_order = 4
    assert lib.csi_create(C.byref(cfg), C.byref(h)) == -1 and b"advection_order" in lib.csi_last_error(None)
    cfg.advection_order = 7
    cfg.Hx = cfg.Hy = 3
    assert lib.csi_create(C.byref(cfg), C.byref(h)) == -1 and b"stencil" in lib.csi_last_error(None)
    cfg.Hx = cfg.Hy = 7
    cfg.nranks, cfg.exchange_every = 2, 4
    assert lib.csi_create(C.byref(cfg), C.byref(h)) == -1 and b"2*exchange_every" in lib.csi_last_error(None)
    assert lib.csi_evp_substeps(None, None, 1.0, 1, None) == -1


def test_synthetic_cases_are_deterministic_and_periodic():
    from climaseaice_b200.synthetic import anticyclone_case, periodic_case, slab_of
    a, b = periodic_case(32), periodic_case(32)
    for k in a.fields:
        assert np.array_equal(a.fields[k], b.fields[k])
    H, N = a.Hx, a.Nx
    assert np.array_equal(a.fields["h"][:, :H], a.fields["h"][:, N:N + H])
    assert (a.fields["a"] == 0).any() and ((a.fields["a"] > 0) & (a.fields["a"] < 1e-3)).any()
    c = anticyclone_case(16)
    assert c.fields["u"].shape == (16 + 14, 16 + 15) and c.fields["v"].shape == (16 + 15, 16 + 14)
    s = slab_of(periodic_case(16, Ny=32), 1, 2, 9)
    assert s.fields["h"].shape == (16 + 18, 16 + 14)
