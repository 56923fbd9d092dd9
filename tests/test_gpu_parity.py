"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs.  Tolerance: north_star asks for max relative error <= 1e-12 after one full time step;
because the substep loop amplifies round-off exponentially the library is built
to reproduce the oracle's operation sequence exactly, so these tests also require BITWISE equality
on the interior."""
import numpy as np
import pytest
import torch

from climaseaice_b200.driver import HostStepper, model_from_case
from climaseaice_b200.synthetic import anticyclone_case, periodic_case
from tests.helpers import NAME_MAP, compare_model, interior_of, oracle_from_case, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12
IMPLS = ("unfused", "auto")


def _assert_parity(res, bitwise=True):
    for n, (err, same) in res.items():
        assert err <= TOL, (n, err)
        if bitwise:
            assert same, (n, err)


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("aice", ["ones", "mixed"])
@pytest.mark.parametrize("nsub", [1, 2, 5, 21])
def test_momentum_substeps_periodic(impl, aice, nsub):
    """time_step_momentum! alone: stresses, u, v, alpha after nsub substeps."""
    case = periodic_case(40, Ny=56, substeps=nsub, aice=aice, timestepper="ForwardEuler")
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    m.update_state(); o.update_state()
    m.time_step_momentum(case.dt); o.time_step_momentum(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "s11", "s22", "s12", "alpha", "P")))
    m.close()


@pytest.mark.parametrize("impl", IMPLS)
def test_full_rk3_step_periodic(impl):
    """One full time_step! (3 stages x {WENO7 tendencies, 150 substeps, h/aice update}) -- the north_star gate."""
    case = periodic_case(64, Ny=48, substeps=150, aice="mixed")
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    # halos of the prognostic fields are periodic images after update_state!, those of the stresses after
    # finalize_rheology! (evp:275-280)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "s11", "s22", "s12"), interior_only=False))
    m.close()


@pytest.mark.parametrize("impl", IMPLS)
def test_two_steps_forward_euler_weno5(impl):
    case = periodic_case(32, substeps=30, aice="mixed", advection_order=5, timestepper="ForwardEuler")
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    m.close()


@pytest.mark.parametrize("order", [1, 3, 5, 7])
def test_advection_tendencies_and_update(order):
    case = periodic_case(48, Ny=40, substeps=1, aice="mixed", advection_order=order)
    m = model_from_case(case)
    o = oracle_from_case(case)
    m.update_state(); o.update_state()
    m.cache_current_fields()
    for n in ("h", "a"):
        o.arr[n + "m"][:] = o.arr[n]
    m.compute_tracer_tendencies(); o.compute_tracer_tendencies()
    _assert_parity(compare_model(m, o, case, names=("Gh", "Ga")))
    m.dynamic_time_step(40.0); o.dynamic_time_step(40.0)
    _assert_parity(compare_model(m, o, case, names=("h", "a")))
    m.close()


@pytest.mark.parametrize("impl", IMPLS)
def test_anticyclone_bounded_domain(impl):
    """BASELINE config 1 geometry (Bounded x Bounded, value BCs, FPlane, wind + ocean drag), reduced size."""
    case = anticyclone_case(48, substeps=40)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "fused"))
def test_five_time_steps_stay_bitwise(impl):
    """Round-off does not creep in over several steps: five full time_step! calls (2250 substeps, 15 advection stages) of the
    Bounded anticyclone case, every prognostic and stress field bit for bit equal to the oracle's after each step."""
    case = anticyclone_case(56, substeps=150)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for step in range(5):
        m.time_step(case.dt); o.time_step(case.dt)
        res = compare_model(m, o, case)
        assert all(same for _, same in res.values()), (step, res)
    m.close()


def test_async_halo_entry_points_on_one_rank():
    """csi_exchange_halos_async / csi_wait_halos (fill_halo_regions!(...; async = true) / synchronize_communication!, evp:204-206,
    275-280) are no-ops without a partition; the partitioned behaviour is what the multi-GPU tests compare with one GPU."""
    import ctypes as C
    from climaseaice_b200 import _lib as L
    case = periodic_case(24, substeps=2)
    m = model_from_case(case)
    arr = (L.csi_array * 1)(m.all_fields()["s11"].as_csi())
    lib = L.lib()
    lib.csi_exchange_halos_async.argtypes = [C.c_void_p, C.POINTER(L.csi_array), C.c_int32, C.c_int32, C.c_void_p]
    lib.csi_wait_halos.argtypes = [C.c_void_p, C.c_void_p]
    assert lib.csi_exchange_halos_async(m._handle, arr, 1, 3, m._stream()) == 0
    assert lib.csi_wait_halos(m._handle, m._stream()) == 0
    assert lib.csi_exchange_halos_async(m._handle, None, 1, 3, m._stream()) == -1   # CSI_ERR_ARG
    m.close()


def test_anticyclone_config1_as_shipped():
    """examples/ice_advected_by_anticyclone.jl as shipped: 128^2, H = 7, substeps = 150, RK3, WENO7."""
    case = anticyclone_case(128, noise=0.0)
    m = model_from_case(case)
    o = oracle_from_case(case)
    m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    m.close()


def test_host_buffer_entry_point_matches_device_entry_point():
    """csi_time_step_host (host buffers, copies inside) == csi_time_step on device-resident fields."""
    case = periodic_case(40, substeps=12, aice="mixed")
    m = model_from_case(case)
    hs = HostStepper(case)
    m.time_step(case.dt)
    hs.time_step(case.dt)
    torch.cuda.synchronize()
    for n in ("u", "v", "h", "a", "s11", "s22", "s12", "alpha"):
        assert np.array_equal(m.all_fields()[n].numpy(), hs.host[n].numpy()), n
    m.close(); hs.model.close()


@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_host_buffer_momentum_entry_point_matches_device_entry_point(timestepper):
    """csi_evp_substeps_host == csi_evp_substeps.  Under RK3 the host call does not upload u, v (reset_velocities! overwrites
    them with Psi^-.u, .v, se:87-93) nor alpha (written, never read; only the rewritten window comes back): poison those host
    arrays to prove they are not read."""
    case = anticyclone_case(40, substeps=12, timestepper=timestepper)
    m = model_from_case(case)
    hs = HostStepper(case)
    m.update_state(); hs.model.update_state()
    if timestepper == "SplitRungeKutta3":
        m.cache_current_fields()
        for n in ("u", "v"):
            hs.host[n + "m"].copy_(hs.model.all_fields()[n].parent)    # Psi^- = the state update_state! left on the device
    for n in ("u", "v", "h", "a"):                                      # the halos update_state! filled, as a host caller holds them
        hs.host[n].copy_(hs.model.all_fields()[n].parent)
    if timestepper == "SplitRungeKutta3":
        hs.host["u"].fill_(float("nan")); hs.host["v"].fill_(float("nan"))
    hs.host["alpha"].fill_(float("nan"))
    m.time_step_momentum(case.dt)
    hs.evp_substeps(case.dt, case.substeps)
    torch.cuda.synchronize()
    for n in ("u", "v", "s11", "s22", "s12", "alpha"):
        g, w = interior_of(m.all_fields()[n].numpy(), case), interior_of(hs.host[n].numpy(), case)
        assert np.isfinite(w).all() and np.array_equal(g, w), n
    h2d, d2h = hs.last_transfer_bytes()
    assert h2d == sum(hs.host[k].numel() * 8 for k in (("h", "a", "s11", "s22", "s12", "top_x", "top_y", "ue", "ve") + (("um", "vm") if timestepper == "SplitRungeKutta3" else ("u", "v")))), h2d
    m.close(); hs.model.close()


def test_diagnostics_and_cfl():
    case = periodic_case(64, substeps=10, aice="mixed")
    m = model_from_case(case)
    o = oracle_from_case(case)
    m.time_step(case.dt); o.time_step(case.dt)
    assert m.cell_advection_timescale() == pytest.approx(o.cell_advection_timescale(), rel=1e-15)
    d = m.diagnostics()
    az = case.dx * case.dy
    assert d["sum_h_Az"] == pytest.approx(o.interior("h").sum() * az, rel=1e-13)
    assert d["max_abs_u"] == np.abs(o.interior("u")).max()
    # deterministic: a second evaluation returns the same bits
    assert m.diagnostics() == d
    m.close()


def test_full_size_properties_4096():
    """BASELINE config 2 size: oracle too slow, so check size-independent properties -- the fused and
    unfused formulations agree bit for bit, sum(h Az) is conserved, everything stays finite."""
    case = periodic_case(4096, substeps=6, aice="ones")
    case.fields["a"] *= 0.6
    a = model_from_case(case, solver_impl="unfused")
    b = model_from_case(case, solver_impl="auto")
    before = a.diagnostics()
    a.time_step(case.dt); b.time_step(case.dt)
    for n in ("u", "v", "h", "a", "s11", "s22", "s12"):
        fa, fb = a.all_fields()[n].parent, b.all_fields()[n].parent
        assert torch.isfinite(fa).all(), n
        assert torch.equal(fa, fb), n
    after = a.diagnostics()
    assert abs(after["sum_h_Az"] - before["sum_h_Az"]) / before["sum_h_Az"] < 1e-12
    a.close(); b.close()


def _assert_same_bits(a, b, case, names=("u", "v", "s11", "s22", "s12", "alpha", "zeta_c", "zeta_f", "delta")):
    """Interior plus the first halo ring (everything an interior stencil can read).  Deeper halo cells of the stresses on a
    Bounded axis are scratch: the general kernels run the stress update over the extended window like the reference's launch
    (evp:145-167), the tile kernel stores the ring the velocity stencils read; no boundary condition ever fills them."""
    for n in names:
        fa, fb = a.all_fields()[n].parent, b.all_fields()[n].parent
        assert torch.isfinite(fa).all(), n
        if n in ("u", "v"):
            assert torch.equal(fa, fb), n
        else:
            r = 1 if n in ("s11", "s22", "s12") else 0   # alpha, zeta, Delta are diagnostics: no halo fill, interior only
            sl = (slice(case.Hy - r, case.Hy + case.Ny + r), slice(case.Hx - r, case.Hx + case.Nx + r))   # j, i in [1 - r, N + r]
            assert torch.equal(fa[sl], fb[sl]), n


def test_bench_configuration_2_fused_equals_general_kernels():
    """The configuration bench.py times at N = 1 -- BASELINE config 2, the Bounded anticyclone case at 4096^2, one
    time_step_momentum! of 150 substeps -- is too large for the oracle; the general kernels (one thread per node, the
    reference's launch structure, themselves compared with the oracle at small sizes) and the fused tile kernel must agree
    bit for bit on every output, halos included."""
    case = anticyclone_case(4096)
    a = model_from_case(case, solver_impl="unfused")
    b = model_from_case(case, solver_impl="fused")
    for m in (a, b):
        m.update_state()
        m.time_step_momentum(120.0, 150)
    _assert_same_bits(a, b, case)
    inv, redone, tiles = b.fused_stats()
    assert inv == 0 and tiles > 0 and redone < 0.01 * 150 * tiles   # the FAST pass did the work
    assert float(b.all_fields()["u"].parent.abs().max()) > 1e-3
    a.close(); b.close()


def test_bench_configuration_3_block_fused_equals_general_kernels():
    """The per-GPU block of BASELINE config 3 that bench.py times at N > 1 (16384 x 2048, doubly periodic, 150 substeps),
    here as one periodic domain on one GPU: fused tile kernel == general kernels, bit for bit."""
    case = periodic_case(16384, Ny=2048, substeps=150, aice="mixed")
    a = model_from_case(case, solver_impl="unfused")
    b = model_from_case(case, solver_impl="fused")
    for m in (a, b):
        m.update_state()
        m.time_step_momentum(120.0, 150)
    _assert_same_bits(a, b, case)
    assert b.fused_stats()[0] == 0 and b.fused_stats()[2] > 0
    a.close(); b.close()


@pytest.mark.parametrize("span", [30, 140])
def test_fast_arithmetic_equals_ieee_operators(span):
    """The kernels' branch-free rcp / div / sqrt / constant-quotient return the IEEE results bit for bit
    (2e9 random operand sets per span, incl. perfect squares, near-equal operands and signed zeros)."""
    import ctypes as C
    from climaseaice_b200 import lib
    out = (C.c_uint64 * 5)()
    rc = lib().csi_selftest_math(2_000_000_000, 20260417 + span, span, out)  # divisors are drawn positive, as in every kernel use
    assert rc == 0
    assert list(out)[:4] == [0, 0, 0, 0], list(out)
    if span <= 140:
        assert out[4] < 0.03 * 2e9   # these exponent ranges stay inside the fast windows (bar the all-ones guard)


def _variant_case(kind):
    """Configurations off the benchmark's common path: every run-time switch of the kernels."""
    from climaseaice_b200.synthetic import Case
    if kind == "no_top_no_coriolis":
        c = periodic_case(37, Ny=23, substeps=7, aice="mixed")
        c.coriolis_f = None
        del c.fields["top_x"], c.fields["top_y"]
    elif kind == "const_ocean":
        c = periodic_case(33, Ny=41, substeps=6, aice="ones")
        del c.fields["ue"], c.fields["ve"]
    elif kind == "periodic_x_bounded_y":
        c = periodic_case(40, Ny=36, substeps=9, aice="mixed")
        c.topology = ("Periodic", "Bounded")
        c.u_bc_value = 0.0
        for k in ("v", "top_y", "ve"):   # Face-y fields carry Ny+1 rows on a Bounded y axis
            c.fields[k] = np.ascontiguousarray(np.vstack([c.fields[k], c.fields[k][-1:]]))
    elif kind == "bounded_x_periodic_y":
        c = periodic_case(36, Ny=40, substeps=9, aice="mixed")
        c.topology = ("Bounded", "Periodic")
        c.v_bc_value = 0.0
        for k in ("u", "top_x", "ue"):
            c.fields[k] = np.ascontiguousarray(np.hstack([c.fields[k], c.fields[k][:, -1:]]))
    elif kind == "tiny_grid":
        c = periodic_case(9, Ny=8, substeps=5, aice="ones")
    elif kind == "halo3":
        c = periodic_case(48, Ny=32, H=4, substeps=8, aice="mixed", advection_order=5)
    else:
        raise KeyError(kind)
    return c


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("kind", ["no_top_no_coriolis", "const_ocean", "periodic_x_bounded_y", "bounded_x_periodic_y", "tiny_grid", "halo3"])
def test_configuration_switches(impl, kind):
    case = _variant_case(kind)
    over = {}
    if kind == "const_ocean":
        over = dict(ue_c=0.03, ve_c=-0.02)
    m = model_from_case(case, solver_impl=impl)
    if kind == "const_ocean":
        m.close()
        from climaseaice_b200 import SemiImplicitStress
        from climaseaice_b200.driver import grid_from_case
        import climaseaice_b200 as csi
        grid = grid_from_case(case)
        F = case.fields
        dyn = csi.SeaIceMomentumEquation(grid, coriolis=csi.FPlane(case.coriolis_f),
                                         top_momentum_stress=dict(u=csi.Field((1, 0), grid, F["top_x"]), v=csi.Field((0, 1), grid, F["top_y"])),
                                         bottom_momentum_stress=SemiImplicitStress(ue=0.03, ve=-0.02), solver=csi.SplitExplicitSolver(substeps=case.substeps))
        m = csi.SeaIceModel(grid, dynamics=dyn, advection=csi.WENO(7), solver_impl=impl)
        m.set(h=F["h"], a=F["a"], u=F["u"], v=F["v"])
    o = oracle_from_case(case, **over)
    m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    m.close()


@pytest.mark.parametrize("impl", IMPLS)
def test_ice_strength_pressure_and_forward_euler(impl):
    case = periodic_case(40, Ny=30, substeps=11, aice="mixed", timestepper="ForwardEuler")
    import climaseaice_b200 as csi
    from climaseaice_b200.driver import grid_from_case
    grid = grid_from_case(case)
    F = case.fields
    dyn = csi.SeaIceMomentumEquation(grid, coriolis=csi.FPlane(case.coriolis_f),
                                     rheology=csi.ElastoViscoPlasticRheology(pressure_formulation="IceStrength", min_relaxation_parameter=30.0),
                                     top_momentum_stress=dict(u=0.05, v=-0.02),
                                     bottom_momentum_stress=csi.SemiImplicitStress(ue=csi.Field((1, 0), grid, F["ue"]), ve=csi.Field((0, 1), grid, F["ve"])),
                                     solver=csi.SplitExplicitSolver(substeps=case.substeps))
    m = csi.SeaIceModel(grid, dynamics=dyn, advection=csi.WENO(7), timestepper="ForwardEuler", solver_impl=impl)
    m.set(h=F["h"], a=F["a"], u=F["u"], v=F["v"])
    del case.fields["top_x"], case.fields["top_y"]
    from oracle import oracle as O
    o = oracle_from_case(case, pressure_formulation=1, alpha_min=30.0, top_kind=O.STRESS_CONST, top_tx=0.05, top_ty=-0.02)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    m.close()


def test_shape_errors_are_reported():
    """A field whose parent does not match the grid is refused with CSI_ERR_SHAPE, not silently accepted."""
    import climaseaice_b200 as csi
    case = periodic_case(24, substeps=2)
    m = model_from_case(case)
    other = csi.RectilinearGrid(size=(25, 24), x=(0, 1e5), y=(0, 1e5), halo=(7, 7))
    m.ice_thickness = csi.Field((0, 0), other)
    with pytest.raises(csi.CsiError) as e:
        m.time_step(case.dt)
    assert e.value.code == -2 and "field 'h'" in str(e.value)
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "auto", "fused"))
def test_immersed_mask(impl):
    """Immersed-boundary land mask (BASELINE config 4 ingredients: masked stresses isd:16-24, peripheral-node
    velocity mask se:226,261, zero flux through immersed faces, mask_immersed_field_xy! in update_state!): in the
    general kernels and inside the fused tile kernel (node flags in shared memory), which 'auto' must pick."""
    case = periodic_case(40, Ny=36, substeps=9, aice="mixed")
    X, Y = case.nodes((0, 0))
    sy, sx = case.parent_shape((0, 0))
    mask = (((X / case.Lx - 0.5) ** 2 + (Y / case.Ly - 0.45) ** 2) < 0.03).astype(np.uint8)   # an island
    mask[:, :case.Hx] = mask[:, case.Nx:case.Nx + case.Hx]; mask[:, case.Nx + case.Hx:] = mask[:, case.Hx:2 * case.Hx]
    mask[:case.Hy, :] = mask[case.Ny:case.Ny + case.Hy, :]; mask[case.Ny + case.Hy:, :] = mask[case.Hy:2 * case.Hy, :]
    import climaseaice_b200 as csi
    from climaseaice_b200.driver import grid_from_case
    grid = grid_from_case(case)
    F = case.fields
    dyn = csi.SeaIceMomentumEquation(grid, coriolis=csi.FPlane(case.coriolis_f),
                                     top_momentum_stress=dict(u=csi.Field((1, 0), grid, F["top_x"]), v=csi.Field((0, 1), grid, F["top_y"])),
                                     bottom_momentum_stress=csi.SemiImplicitStress(ue=csi.Field((1, 0), grid, F["ue"]), ve=csi.Field((0, 1), grid, F["ve"])),
                                     solver=csi.SplitExplicitSolver(substeps=case.substeps))
    m = csi.SeaIceModel(grid, dynamics=dyn, advection=csi.WENO(7), solver_impl=impl, immersed_mask=mask)
    m.set(h=F["h"], a=F["a"], u=F["u"], v=F["v"])
    from oracle import oracle as O
    from tests.helpers import oracle_params
    o = O.OracleModel(case.Nx, case.Ny, case.Hx, case.Hy, dx=case.dx, dy=case.dy, params=oracle_params(case),
                      fields={k: v.copy() for k, v in case.fields.items()}, mask=mask)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    inside = mask[case.Hy:-case.Hy, case.Hx:-case.Hx].astype(bool)
    assert np.all(interior_of(m.all_fields()["h"].numpy(), case)[inside] == 0)   # land stays ice free
    assert (m.fused_stats()[2] > 0) == (impl != "unfused")   # the tile kernel really ran (tiles per substep)
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "auto", "fused"))
def test_coastline_config4_reduced(impl):
    """BASELINE config 4 (examples/ice_advected_on_coastline.jl) at reduced size: Periodic x Bounded, immersed
    triangular coast, uniform wind, ocean at rest, wall BCs and the linear immersed drag flux BC."""
    from climaseaice_b200.synthetic import coastline_case
    case = coastline_case(Ny=48, substeps=30)
    assert case.mask.any() and not case.mask.all()
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    u = interior_of(m.all_fields()["u"].numpy(), case)
    assert np.abs(u).max() > 1e-4          # the wind moved the ice
    assert (m.fused_stats()[2] > 0) == (impl != "unfused")
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "fused"))
@pytest.mark.parametrize("topology", [("Bounded", "Bounded"), ("Periodic", "Bounded")])
@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_latitude_longitude_grid(impl, topology, timestepper):
    """j-dependent metrics (LatitudeLongitudeGrid, the grid of test/test_rheology_energy_budget.jl:18-24): every
    dx/dy/Az in the strain rates, the SBP stress divergence, the relaxation factors, the flux divergence, the value
    BCs and the CFL reduction is the row's own -- in the general kernels and in the fused tile kernel, which reads a
    per-row table of metrics and correctly rounded reciprocals."""
    from climaseaice_b200.synthetic import latlon_case
    case = latlon_case(48, substeps=20, topology=topology, timestepper=timestepper)
    met = case.metrics()
    assert met["dxcc"].max() / met["dxcc"].min() > 2          # the metrics really vary
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "s11", "s22", "s12", "alpha", "zeta_c", "delta")))
    assert np.abs(interior_of(m.all_fields()["u"].numpy(), case)).max() > 1e-4
    assert m.cell_advection_timescale() == pytest.approx(o.cell_advection_timescale(), rel=1e-15)
    d = m.diagnostics()
    az = met["azcc"][case.Hy:case.Hy + case.Ny][:, None]
    assert d["sum_h_Az"] == pytest.approx((o.interior("h") * az).sum(), rel=1e-13)
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "fused", "auto"))
@pytest.mark.parametrize("topology", [("Bounded", "Bounded"), ("Periodic", "Bounded"), ("Periodic", "Periodic")])
@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_two_dimensional_metrics(impl, topology, timestepper):
    """Metrics that depend on i and j (CSI_METRIC_IJ: orthogonal curvilinear grids, the metric layout of Oceananigans'
    OrthogonalSphericalShellGrid family): every dx/dy/Az of the strain rates, the stress divergence, the relaxation
    factors, the flux divergence, the value BCs and the reductions is taken at the
    (i, j) the reference's operators name -- in the general kernels and in the fused tile kernel, which reads each node's
    metrics and their correctly rounded reciprocals from planes in its own layout."""
    from climaseaice_b200.synthetic import curvilinear_case
    case = curvilinear_case(72, 56, substeps=20, topology=topology, timestepper=timestepper)
    met = case.metrics()
    assert met["dxcc"].max() / met["dxcc"].min() > 1.5 and np.ptp(met["dxcc"], axis=1).max() > 100.0   # they vary along i too
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "s11", "s22", "s12", "alpha", "zeta_c", "delta")))
    assert np.abs(interior_of(m.all_fields()["u"].numpy(), case)).max() > 1e-4
    st = m.fused_stats()
    assert (st[2] > 0) == (impl != "unfused") and st[0] == 0   # the tile kernel ran (tiles per substep) on validated inputs
    assert st[1] <= st[2] // 4                                  # ... and its FAST pass carried the tiles, not the IEEE re-pass
    assert m.cell_advection_timescale() == pytest.approx(o.cell_advection_timescale(), rel=1e-15)
    d = m.diagnostics()
    az = met["azcc"][case.Hy:case.Hy + case.Ny, case.Hx:case.Hx + case.Nx]
    assert d["sum_h_Az"] == pytest.approx((o.interior("h") * az).sum(), rel=1e-13)
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "fused"))
def test_two_dimensional_metrics_with_a_coast(impl):
    """An island and the linear immersed drag on a curvilinear mesh (Periodic x Bounded): the immersed stress divergence takes its
    face lengths and areas at the node's own (i, j) too."""
    from climaseaice_b200.synthetic import LOC, curvilinear_case
    case = curvilinear_case(64, 48, H=5, substeps=16, topology=("Periodic", "Bounded"))
    X, Y = case.nodes(LOC["h"])
    land = (((X / case.Lx - 0.4) ** 2 + (Y / case.Ly - 0.55) ** 2) < 0.012) | ((X / case.Lx > 0.8) & (Y / case.Ly > 0.85))
    mask = land.astype(np.uint8)
    mask[:, :case.Hx] = mask[:, case.Nx:case.Nx + case.Hx]
    mask[:, case.Nx + case.Hx:] = mask[:, case.Hx:2 * case.Hx]
    case.mask = np.ascontiguousarray(mask)
    case.immersed_drag = (2e-3, 1e-3)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "s11", "s22", "s12", "alpha", "zeta_c", "delta")))
    assert (m.fused_stats()[2] > 0) == (impl == "fused")
    m.close()


@pytest.mark.parametrize("topology", [("Periodic", "Bounded"), ("Periodic", "Periodic")])
def test_two_dimensional_metrics_many_tiles_fused_equals_general(topology):
    """11 x 18 tiles, most of them interior tiles (the instantiation without edge logic): the fused kernel on per-node metric
    planes equals the general kernels bit for bit on every evolving field and on the auxiliaries of the last substep."""
    from climaseaice_b200.synthetic import curvilinear_case
    case = curvilinear_case(320, 240, H=5, substeps=12, topology=topology)
    a, b = model_from_case(case, solver_impl="fused"), model_from_case(case, solver_impl="unfused")
    for _ in range(2):
        a.time_step(case.dt); b.time_step(case.dt)
    Fa, Fb = a.all_fields(), b.all_fields()
    for n in ("u", "v", "h", "a", "s11", "s22", "s12", "alpha", "zeta_c", "zeta_f", "delta"):
        assert np.array_equal(interior_of(Fa[n].numpy(), case), interior_of(Fb[n].numpy(), case)), n
    st = a.fused_stats()
    assert st[2] >= 11 * 18 and st[0] == 0 and st[1] <= st[2] // 10
    a.close(); b.close()


def test_latitude_longitude_grid_rejects_bad_metrics():
    from climaseaice_b200.synthetic import latlon_case
    bad = latlon_case(32, substeps=4)
    met = bad.metrics()
    met["azff"][bad.Hy + 3] = 0.0
    bad.metrics = lambda: met
    with pytest.raises(RuntimeError, match="metrics must be positive"):
        model_from_case(bad)


@pytest.mark.parametrize("impl", ["unfused", "fused"])
@pytest.mark.parametrize("variant", ["bottom_drag", "top_drag", "fields", "both_drag", "const_top_drag"])
@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_free_drift_top_drag_and_snow(variant, timestepper, impl):
    """free-drift velocities of marginal ice (StressBalanceFreeDrift closed forms for either side, and
    (u=, v=) arrays: stress_balance_free_drift.jl:61-129), SemiImplicitStress as the top stress and prescribed
    bottom stresses (ext.jl:8-40,176-210), and snow-thickness advection (tracer_tendency:47-52, fe.jl:84-94)."""
    from climaseaice_b200.synthetic import marginal_ice_case
    case = marginal_ice_case(48, substeps=12, variant=variant, timestepper=timestepper)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "hs", "s11", "s22", "s12", "Ghs")))
    u = interior_of(m.all_fields()["u"].numpy(), case)
    assert np.isfinite(u).all() and np.abs(u).max() > 1e-3
    if impl == "fused":   # the tile kernel ran, and on its FAST pass: free drift, top drag and prescribed bottom stress are
        inv, redone, tiles = m.fused_stats()   # stage constants / GEN branches of it, not a reason to fall back
        assert tiles > 0 and inv == 0 and redone <= 0.05 * case.substeps * tiles, (inv, redone, tiles)
    m.close()


def test_free_drift_on_a_bounded_domain_and_host_entry():
    """Free drift + snow on a Bounded x Bounded domain, through the device and the host-buffer entry points."""
    from climaseaice_b200.synthetic import marginal_ice_case
    case = marginal_ice_case(40, substeps=8, variant="bottom_drag", topology=("Bounded", "Bounded"))
    m = model_from_case(case)
    o = oracle_from_case(case)
    hs = HostStepper(case)
    m.time_step(case.dt); o.time_step(case.dt); hs.time_step(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "h", "a", "hs", "s12")))
    for n in ("u", "v", "h", "a", "hs"):
        assert np.array_equal(interior_of(hs.host[n].numpy(), case), interior_of(m.all_fields()[n].numpy(), case)), n
    m.close(); hs.model.close()


def test_stress_balance_free_drift_argument_errors():
    from climaseaice_b200 import SeaIceMomentumEquation, SemiImplicitStress, StressBalanceFreeDrift
    from climaseaice_b200.driver import grid_from_case
    grid = grid_from_case(periodic_case(16))
    sis = SemiImplicitStress()
    with pytest.raises(ValueError, match="not both"):
        SeaIceMomentumEquation(grid, top_momentum_stress=sis, bottom_momentum_stress=sis, free_drift=StressBalanceFreeDrift())
    with pytest.raises(ValueError, match="requires using a `SemiImplicitStress`"):
        SeaIceMomentumEquation(grid, top_momentum_stress=dict(u=0.1, v=0.0), free_drift=StressBalanceFreeDrift())
    # the library checks the same thing for callers that fill csi_config themselves
    import ctypes as C
    from climaseaice_b200 import _lib as L
    case = periodic_case(16, substeps=2)
    m = model_from_case(case)
    cfg = m._config()
    cfg.free_drift_kind = L.FD_STRESS_BALANCE
    cfg.top_stress_kind = L.STRESS_SEMI_IMPLICIT
    h = C.c_void_p()
    assert L.lib().csi_create(C.byref(cfg), C.byref(h)) == -1 and b"not both" in L.lib().csi_last_error(None)
    m.close()
    # "fused" refuses what only the general kernels implement: a folded north boundary without two-dimensional metrics
    mc = _regular_fold_case()
    mf = model_from_case(mc, solver_impl="fused")
    with pytest.raises(RuntimeError, match="folded"):
        mf.time_step(mc.dt)
    mf.close()


@pytest.mark.parametrize("impl", ("unfused", "fused"))
def test_hydrostatic_spherical_coriolis_on_latlon_grid(impl):
    """HydrostaticSphericalCoriolis (EnstrophyConserving, f^ff = 2 Omega sin(phi_f) per row) on the lat-lon basin."""
    from climaseaice_b200.synthetic import latlon_case
    case = latlon_case(48, substeps=20, topology=("Periodic", "Bounded"))
    m0 = model_from_case(case)                      # FPlane(1e-4), for contrast
    case.rotation_rate = 7.292115e-5
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    assert o.prm["coriolis_kind"] == 2
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt); m0.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    u, u0 = interior_of(m.all_fields()["u"].numpy(), case), interior_of(m0.all_fields()["u"].numpy(), case)
    assert np.abs(u - u0).max() > 1e-6               # f varies with latitude: differs from the f-plane run
    m.close(); m0.close()


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("nsub", [1, 4])
def test_extreme_magnitudes_stay_bitwise(impl, nsub):
    """Subnormal and near-underflow values in every input (u, v, ocean velocity, wind stress, h, aice, sigma): the fused
    kernel's shortcut arithmetic (branch-free division / sqrt, power-of-two-scaled expression tree) is only valid inside
    its exponent windows, so such tiles must be detected and redone with the IEEE operators -- same bits as the oracle."""
    case = periodic_case(64, Ny=48, substeps=nsub, aice="mixed", timestepper="ForwardEuler")
    F, H = case.fields, case.Hx

    def patch(arr, j0, j1, i0, i1, val):
        arr[H + j0:H + j1, H + i0:H + i1] = val

    patch(F["u"], 5, 9, 5, 9, 4.9e-324)
    patch(F["v"], 5, 9, 12, 16, -1e-310)
    patch(F["u"], 12, 15, 5, 9, 3e-200)
    patch(F["v"], 12, 15, 12, 16, 7e-160)
    patch(F["ue"], 20, 24, 5, 9, 1e-309)
    patch(F["ve"], 20, 24, 12, 16, -2e-308)
    patch(F["top_x"], 10, 14, 30, 34, 1e-312)
    patch(F["top_y"], 10, 14, 36, 40, -3e-250)
    patch(F["h"], 30, 34, 5, 9, 1e-180)
    patch(F["a"], 30, 34, 12, 16, 1e-200)
    patch(F["h"], 36, 40, 20, 24, 1e-300)
    patch(F["a"], 36, 40, 28, 32, 5e-324)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for name, (j0, i0, val) in dict(s11=(20, 40, 1e-320), s22=(24, 44, -3e-310), s12=(28, 48, 2e-200)).items():
        patch(m.all_fields()[name].parent, j0, j0 + 3, i0, i0 + 3, val)
        patch(o.arr[name], j0, j0 + 3, i0, i0 + 3, val)
    m.update_state(); o.update_state()
    m.time_step_momentum(case.dt); o.time_step_momentum(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "s11", "s22", "s12", "alpha", "P")))
    if impl == "auto":
        assert m.fused_stats()[0] == 1   # subnormal inputs fail the validation: the whole stage takes the IEEE pass
    m.close()


@pytest.mark.parametrize("variant", ["bottom_drag", "top_drag", "fields", "both_drag", "const_top_drag"])
def test_ieee_pass_of_the_less_common_configurations(variant):
    """Free drift, a SemiImplicitStress on top and prescribed bottom stresses inside the fused kernel's IEEE re-pass (the
    reference's expression tree with plain operators): a subnormal thickness fails the stage's input validation, every
    tile takes that pass, and the result is still the oracle's bit for bit."""
    from climaseaice_b200.synthetic import marginal_ice_case
    case = marginal_ice_case(48, substeps=6, variant=variant, timestepper="ForwardEuler", snow=False)
    case.fields["h"][case.Hy + 3, case.Hx + 3] = 1e-310
    m = model_from_case(case, solver_impl="fused")
    o = oracle_from_case(case)
    m.update_state(); o.update_state()
    m.time_step_momentum(case.dt); o.time_step_momentum(case.dt)
    _assert_parity(compare_model(m, o, case, names=("u", "v", "s11", "s22", "s12", "alpha")))
    inv, redone, tiles = m.fused_stats()
    assert inv == 1 and redone >= tiles
    m.close()


@pytest.mark.parametrize("impl", IMPLS)
def test_out_of_window_intermediates_fall_back_per_tile(impl):
    """Inputs inside the validated range [2^-300, 2^300) whose *intermediates* leave the kernel's windows (squares of
    1e-89 strain rates underflow the radicand window, 1e81 strain rates overflow it): only the tiles concerned are
    redone with the IEEE operators; everything stays bitwise equal to the oracle."""
    case = periodic_case(96, Ny=64, substeps=3, aice="ones", timestepper="ForwardEuler")
    F, H = case.fields, case.Hx
    rng = np.random.default_rng(7)
    F["u"][H + 6:H + 14, H + 6:H + 14] = 1e-85 * rng.uniform(0.5, 1.0, (8, 8))
    F["v"][H + 6:H + 14, H + 20:H + 28] = -3e-88 * rng.uniform(0.5, 1.0, (8, 8))
    F["u"][H + 40:H + 46, H + 60:H + 66] = 1e85 * rng.uniform(0.5, 1.0, (6, 6))
    F["v"][H + 40:H + 46, H + 70:H + 76] = -2e84 * rng.uniform(0.5, 1.0, (6, 6))
    F["ue"][H + 24:H + 30, H + 40:H + 46] = 1e-80
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    m.update_state(); o.update_state()
    m.time_step_momentum(case.dt); o.time_step_momentum(case.dt)
    res = compare_model(m, o, case, names=("u", "v", "s11", "s22", "s12", "alpha"))
    for n, (err, same) in res.items():
        assert same, (n, err)
    if impl == "auto":
        invalid, redone, tiles = m.fused_stats()
        assert invalid == 0 and 0 < redone < 3 * tiles, (invalid, redone, tiles)   # the tiles concerned, not the whole stage
    m.close()


@pytest.mark.parametrize("impl", ("unfused", "auto", "fused"))
@pytest.mark.parametrize("mask", (False, True))
@pytest.mark.parametrize("timestepper", ["SplitRungeKutta3", "ForwardEuler"])
def test_folded_north_boundary(impl, mask, timestepper):
    """topo_y = CSI_FOLDED (the north fold of Oceananigans' TripolarGrid, handed over as copy lists the host reads off its own
    fill_halo_regions!): velocities change sign across the fold, thickness and concentration do not, the south is a wall.
    Parity with the oracle on the whole interior AND on the folded halo elements -- on the general kernels, and on the fused tile
    kernel, whose tile rows next to the fold run the substep as two launches with the fold fill of the first velocity between."""
    from climaseaice_b200.synthetic import folded_case
    case = folded_case(substeps=12, mask=mask, timestepper=timestepper)
    m = model_from_case(case, solver_impl=impl)
    o = oracle_from_case(case)
    for _ in range(2):
        m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    F = m.all_fields()
    for n, loc, sign in (("u", (1, 0), -1.0), ("v", (0, 1), -1.0), ("h", (0, 0), 1.0), ("a", (0, 0), 1.0)):
        tg, sr = case.fold["maps"][loc]
        g = F[n].numpy().reshape(-1)
        assert np.array_equal(g[tg], sign * g[sr]), n
        assert np.array_equal(g[tg], o.arr[NAME_MAP[n]].reshape(-1)[tg]), n
    st = m.fused_stats()
    assert (st[2] > 0) == (impl != "unfused") and st[0] == 0
    m.close()


def test_folded_north_boundary_larger_than_one_tile_row():
    """Several tile rows: the bulk runs the one-launch substep, the rows next to the fold the split one; fused == general kernels
    bit for bit on every evolving field (interior and the fold's targets), and the FAST pass carries the tiles."""
    from climaseaice_b200.synthetic import folded_case
    case = folded_case(96, 80, H=7, substeps=20)
    a, b = model_from_case(case, solver_impl="fused"), model_from_case(case, solver_impl="unfused")
    for _ in range(2):
        a.time_step(case.dt); b.time_step(case.dt)
    Fa, Fb = a.all_fields(), b.all_fields()
    for n in ("u", "v", "h", "a", "s11", "s22", "s12"):
        assert np.array_equal(interior_of(Fa[n].numpy(), case), interior_of(Fb[n].numpy(), case)), n
    for n, loc in (("u", (1, 0)), ("v", (0, 1))):
        tg, _ = case.fold["maps"][loc]
        assert np.array_equal(Fa[n].numpy().reshape(-1)[tg], Fb[n].numpy().reshape(-1)[tg]), n
    st = a.fused_stats()
    assert st[2] > 0 and st[0] == 0 and st[1] <= st[2] // 2
    a.close(); b.close()


def _regular_fold_case():
    """A fold on a regular grid (no metric arrays): only the general kernels take it."""
    from climaseaice_b200.synthetic import example_fold_maps
    case = periodic_case(48, Ny=40, substeps=4, aice="mixed")
    case.topology = ("Periodic", "Folded")
    case.u_bc_value = 0.0
    for k in ("v", "top_y", "ve"):
        case.fields[k] = np.ascontiguousarray(np.vstack([case.fields[k], case.fields[k][-1:]]))
    case.fold = dict(maps=example_fold_maps(case.Nx, case.Ny, case.Hx, case.Hy), sign_velocity=-1.0, sign_external=1.0)
    return case


def test_fused_solver_refuses_a_fold_without_two_dimensional_metrics():
    case = _regular_fold_case()
    m = model_from_case(case, solver_impl="fused")
    with pytest.raises(RuntimeError, match="folded"):
        m.time_step(case.dt)
    m.close()
    m, o = model_from_case(case, solver_impl="auto"), oracle_from_case(case)   # auto: the general kernels, equal to the oracle
    m.time_step(case.dt); o.time_step(case.dt)
    _assert_parity(compare_model(m, o, case))
    assert m.fused_stats()[2] == 0
    m.close()


@pytest.mark.parametrize("kind", ("anticyclone", "periodic_mixed", "coastline"))
def test_small_grid_cooperative_launch_equals_launch_per_substep(kind, monkeypatch):
    """Grids whose tiles are all resident at once run the substeps of a stage, but the last, as ONE cooperative launch with a
    grid-wide barrier between substeps (k_evp_substeps_persistent); CSI_PERSISTENT=0 keeps the launch per substep.  Same tile
    passes, same stores: every field equal bit for bit, and both equal to the oracle; the launch count shows which path ran."""
    from climaseaice_b200.synthetic import coastline_case
    if kind == "anticyclone":
        case = anticyclone_case(128, noise=0.0, substeps=40)      # BASELINE config 1 as shipped, fewer substeps
    elif kind == "periodic_mixed":
        case = periodic_case(96, Ny=64, substeps=25, aice="mixed")
    else:
        case = coastline_case(Ny=48, substeps=30)
    runs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("CSI_PERSISTENT", mode)
        m = model_from_case(case, solver_impl="fused")
        m.time_step(case.dt)
        l0 = m.launch_count
        m.time_step(case.dt)
        F = m.all_fields()
        runs[mode] = ({n: F[n].numpy().copy() for n in ("u", "v", "h", "a", "s11", "s22", "s12", "alpha", "zeta_c", "zeta_f", "delta")}, m.launch_count - l0, m.fused_stats())
        if mode == "1":
            o = oracle_from_case(case)
            for _ in range(2):
                o.time_step(case.dt)
            _assert_parity(compare_model(m, o, case))
        m.close()
    for n, a in runs["0"][0].items():
        assert np.array_equal(interior_of(a, case), interior_of(runs["1"][0][n], case)), n
    assert runs["0"][2] == runs["1"][2]
    stages = 3 if case.timestepper == "SplitRungeKutta3" else 1
    saved = runs["0"][1] - runs["1"][1]
    if saved == 0:   # (seen under ncu: the device reports no cooperative launch and the library keeps one launch per substep)
        pytest.skip("cooperative launch not available in this environment: both runs took one launch per substep")
    assert saved == stages * (case.substeps - 2)     # substeps 1 .. n - 1 became one launch per stage
