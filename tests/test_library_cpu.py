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


def test_struct_layout_matches_header(tmp_path):
    """The ctypes mirrors agree with what a C compiler makes of include/climaseaice_b200.h: total sizes and the
    offset of every member of csi_config (members of csi_fields are all csi_array, in FIELD_NAMES order)."""
    import subprocess
    from pathlib import Path
    inc = Path(__file__).resolve().parents[1] / "include"
    members = [n for n, _ in L.csi_config._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "climaseaice_b200.h"', 'int main(void) {',
             'printf("%zu %zu %zu\\n", sizeof(csi_array), sizeof(csi_fields), sizeof(csi_config));']
    lines += [f'printf("%zu\\n", offsetof(csi_config, {m}));' for m in members]
    lines += ['return 0; }']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(inc), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out[:3]] == [C.sizeof(L.csi_array), C.sizeof(L.csi_fields), C.sizeof(L.csi_config)]
    assert [int(x) for x in out[3:]] == [getattr(L.csi_config, m).offset for m in members]
    assert C.sizeof(L.csi_array) == 24 and C.sizeof(L.csi_fields) == 24 * len(L.FIELD_NAMES) == 24 * 29


def test_julia_shim_structs_match_the_ctypes_mirror():
    """Julia cannot run in this image, so the shim's C-struct mirrors are desk-checked mechanically: `CsiConfig` in
    julia/ClimaSeaIceB200.jl must list the members of csi_config in the header's order with the matching Julia type, and
    FIELD_ORDER the 29 members of csi_fields (the ctypes mirror is itself checked against the C compiler above)."""
    import re
    from pathlib import Path
    src = (Path(__file__).resolve().parents[1] / "julia" / "ClimaSeaIceB200.jl").read_text()
    body = src[src.index("Base.@kwdef struct CsiConfig"):]
    body = body[:body.index("\nend")]
    jl = re.findall(r"^\s*([A-Za-z_][A-Za-z0-9_]*)\s*::\s*([A-Za-z0-9{}, ]+?)(?=\s*(?:=|;|\n|$))", body.replace(";", "\n"), re.M)
    jl = [(n, t.strip()) for n, t in jl]
    want = []
    for n, ct in L.csi_config._fields_:
        if ct is C.c_int32:
            t = "Int32"
        elif ct is C.c_double:
            t = "Float64"
        elif n == "immersed_mask":
            t = "Ptr{UInt8}"
        elif n == "metrics":
            t = "NTuple{12, Ptr{Float64}}"
        elif n in ("fold_target", "fold_source"):
            t = "NTuple{4, Ptr{Int32}}"
        elif n == "fold_count":
            t = "NTuple{4, Int32}"
        else:
            t = "Ptr{Float64}"
        want.append((n, t))
    assert jl == want, [(a, b) for a, b in zip(jl, want) if a != b][:3]
    order = re.search(r"const FIELD_ORDER = \((.*?)\)", src, re.S).group(1)
    assert tuple(re.findall(r":([A-Za-z0-9_]+)", order)) == tuple(L.FIELD_NAMES)
    # the drop-in methods have the reference's arities (3-argument time_step_momentum!, no extra handle argument)
    assert re.search(r"function time_step_momentum!\(model, dynamics::B200Dynamics, Δt\)", src)
    assert "h::Handle)" not in src


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


def test_create_validates_fold_copy_lists():
    """CSI_FOLDED: the copy lists are checked before anything touches the device -- all four locations present, indices inside the
    parent of their location, signs +-1, every target written once and nothing both read and written (the copies run concurrently)."""
    from climaseaice_b200.synthetic import example_fold_maps
    lib = csi.lib()
    h = C.c_void_p()
    Nx = Ny = 16
    H = 4

    def attempt(maps, sv=-1.0, se=1.0, partition_x=1):
        cfg = L.csi_config()
        cfg.abi_version, cfg.Nx, cfg.Ny, cfg.Hx, cfg.Hy, cfg.dx, cfg.dy, cfg.substeps = 1, Nx, Ny, H, H, 1.0, 1.0, 1
        cfg.advection_order, cfg.topo_x, cfg.topo_y = 3, L.PERIODIC, L.FOLDED
        cfg.fold_sign_velocity, cfg.fold_sign_external = sv, se
        if partition_x > 1:
            cfg.nranks, cfg.partition_x, cfg.exchange_every = partition_x, partition_x, 1
        keep = []
        for (lx, ly), (tg, sr) in maps.items():
            k = lx + 2 * ly
            tg, sr = np.ascontiguousarray(tg, dtype=np.int32), np.ascontiguousarray(sr, dtype=np.int32)
            keep += [tg, sr]
            cfg.fold_target[k] = tg.ctypes.data_as(C.POINTER(C.c_int32))
            cfg.fold_source[k] = sr.ctypes.data_as(C.POINTER(C.c_int32))
            cfg.fold_count[k] = tg.size
        rc = lib.csi_create(C.byref(cfg), C.byref(h))
        return rc, lib.csi_last_error(None)

    good = example_fold_maps(Nx, Ny, H, H)
    rc, msg = attempt({k: v for k, v in good.items() if k != (1, 1)})
    assert rc == -1 and b"each of the four locations" in msg
    bad = dict(good); bad[(0, 0)] = (good[(0, 0)][0], good[(0, 0)][1] + 10 ** 6)
    rc, msg = attempt(bad)
    assert rc == -1 and b"outside the parent" in msg
    rc, msg = attempt(good, sv=0.5)
    assert rc == -1 and b"signs" in msg
    bad = dict(good); bad[(1, 0)] = (np.r_[good[(1, 0)][0], good[(1, 0)][0][:1]], np.r_[good[(1, 0)][1], good[(1, 0)][1][:1]])
    rc, msg = attempt(bad)
    assert rc == -1 and b"same target twice" in msg
    bad = dict(good); bad[(0, 1)] = (good[(0, 1)][0], np.r_[good[(0, 1)][0][1:2], good[(0, 1)][1][1:]])
    rc, msg = attempt(bad)
    assert rc == -1 and b"also writes" in msg
    rc, msg = attempt(good, partition_x=2)
    assert rc != 0 and b"partition along x" in msg
    if not torch.cuda.is_available():   # a valid description gets as far as the device
        rc, msg = attempt(good)
        assert rc == -4 and b"no CUDA device" in msg


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


def test_bench_block_generator_is_consistent_across_ranks():
    """bench.py builds each rank's block of the global doubly periodic case without the global arrays: the analytic fields
    of neighbouring blocks must agree where a block's halo overlaps its neighbour's interior (2 x 2 partition, rank = ry Rx + rx)."""
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("bench_mod", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import __graft_entry__ as entry
    entry.load_package()
    nx, ny, H, Rx, Ry = 24, 16, 5, 2, 2
    blocks = {(rx, ry): bench.periodic_slab_case(nx, ny, ry, Ry, H, rx=rx, Rx=Rx) for rx in range(Rx) for ry in range(Ry)}
    for name in ("u", "v", "ue", "ve", "top_x", "top_y"):
        for (rx, ry), c in blocks.items():
            a = c.fields[name]
            assert a.shape == (ny + 2 * H, nx + 2 * H)
            east = blocks[((rx + 1) % Rx, ry)].fields[name]
            north = blocks[(rx, (ry + 1) % Ry)].fields[name]
            # my east halo columns = the first H interior columns of the block to the east (periodic wrap: the fields are periodic)
            assert np.allclose(a[H:H + ny, H + nx:], east[H:H + ny, H:2 * H], rtol=0, atol=1e-15)
            assert np.allclose(a[H + ny:, H:H + nx], north[H:2 * H, H:H + nx], rtol=0, atol=1e-15)
