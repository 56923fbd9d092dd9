"""CPU tests: the oracle against the reference's own property tests.

The reference ships no golden vectors for this path; these are the pins it does have."""
import numpy as np
import pytest

from climaseaice_b200.synthetic import anticyclone_case, periodic_case, slab_of
from oracle import oracle as O
from tests.helpers import oracle_from_case


def latlon_metrics(N, H, lam=(0, 60), phi=(20, 70)):
    R = 6371e3
    dl = np.deg2rad((lam[1] - lam[0]) / N)
    dp = (phi[1] - phi[0]) / N
    j = np.arange(1 - H, N + H + 2)
    phif, phic = np.deg2rad(phi[0] + (j - 1) * dp), np.deg2rad(phi[0] + (j - 0.5) * dp)
    phif_n, phic_s = np.deg2rad(phi[0] + j * dp), np.deg2rad(phi[0] + (j - 1.5) * dp)
    dxc, dxf = R * np.cos(phic) * dl, R * np.cos(phif) * dl
    dy = np.full_like(dxc, R * np.deg2rad(dp))
    azc, azf = R * R * dl * (np.sin(phif_n) - np.sin(phif)), R * R * dl * (np.sin(phic) - np.sin(phic_s))
    return dict(dxcc=dxc, dxfc=dxc, dxcf=dxf, dxff=dxf, dycc=dy, dyfc=dy, dycf=dy, dyff=dy, azcc=azc, azfc=azc, azcf=azf, azff=azf)


@pytest.mark.parametrize("N", [40, 80])

def test_energy_budget_adjoint_identity(N):
    """test/test_rheology_energy_budget.jl:50-124 on the same lat-lon grid and trig fields."""
    H = 4
    m = O.OracleModel(N, N, H, H, topo=(O.BOUNDED, O.BOUNDED), metrics=latlon_metrics(N, H))
    dl, dp = 60 / N, 50 / N
    lh = lambda l: l / 60 * 2 * np.pi
    ph = lambda p: (p - 20) / 50 * 2 * np.pi

    def setf(name, fn, xf, yf):
        a = m.arr[name]
        a[:] = 0
        for i in range(3, N - 1):          # 1+margin : N-margin, margin = 2
            for j in range(3, N - 1):
                l = (i - 1) * dl if xf else (i - 0.5) * dl
                p = 20 + ((j - 1) * dp if yf else (j - 0.5) * dp)
                a[j - 1 + H, i - 1 + H] = fn(l, p)

    setf("u", lambda l, p: np.sin(2 * lh(l)) * np.cos(3 * ph(p)), 1, 0)
    setf("v", lambda l, p: np.cos(3 * lh(l)) * np.sin(2 * ph(p)), 0, 1)
    setf("s11", lambda l, p: np.sin(lh(l)) * np.sin(2 * ph(p)), 0, 0)
    setf("s22", lambda l, p: np.cos(2 * lh(l)) * np.cos(ph(p)), 0, 0)
    setf("s12", lambda l, p: np.sin(3 * lh(l)) * np.cos(2 * ph(p)), 1, 1)
    Wn, Wo, D = m.stress_power_budget()
    imb = lambda W: abs(W + D) / max(abs(W), abs(D))
    assert imb(Wn) < 1e-10            # :117
    assert imb(Wo) > 1e-3             # :120
    assert imb(Wn) < 1e-6 * imb(Wo)   # :123


def test_energy_budget_adjoint_identity_two_dimensional_metrics():
    """The same summation-by-parts identity (test/test_rheology_energy_budget.jl:50-124) with metrics that depend on i and j
    (orthogonal curvilinear grid): it holds only if the strain rates and the stress divergence take every metric at
    matching (i, j) -- the pairing of evp.jl:360-375 with ice_stress_divergence.jl:39-51."""
    from climaseaice_b200.synthetic import curvilinear_case
    N, Ny, H = 36, 28, 4
    c = curvilinear_case(N, Ny, H=H)
    m = O.OracleModel(N, Ny, H, H, topo=(O.BOUNDED, O.BOUNDED), metrics=c.metrics())
    rng = np.random.default_rng(3)
    for name in ("u", "v", "s11", "s22", "s12"):
        a = m.arr[name]
        a[:] = 0
        a[H + 2:H + Ny - 2, H + 2:H + N - 2] = rng.uniform(-1, 1, (Ny - 4, N - 4))   # compact support: no boundary terms
    Wn, Wo, D = m.stress_power_budget()
    imb = lambda W: abs(W + D) / max(abs(W), abs(D))
    assert imb(Wn) < 1e-12
    assert imb(Wo) > 1e-3


def test_semi_implicit_ocean_drag_bounds():
    """test/test_time_stepping.jl:56-80: 8x8 periodic, substeps=10, 20 steps of 60 s from rest."""
    N, H, uo = 8, 4, 0.1
    m = O.OracleModel(N, N, H, H, dx=10000 / 8, dy=10000 / 8,
                      params=dict(substeps=10, bot_kind=O.STRESS_SEMI_IMPLICIT, ue_c=uo, ve_c=0.0, advection_order=0))
    m.arr["h"][:] = 1
    m.arr["a"][:] = 1
    for _ in range(20):
        m.time_step(60.0)
    u = m.interior("u")
    assert np.isfinite(u).all()
    assert 0 < u.max() <= uo


def test_constant_state_is_preserved():
    """Uniform ice at rest with no forcing stays at rest; h, aice unchanged."""
    N, H = 16, 7
    m = O.OracleModel(N, N, H, H, dx=4000.0, dy=4000.0, params=dict(substeps=10, advection_order=7))
    m.arr["h"][:] = 0.5
    m.arr["a"][:] = 0.8
    m.time_step(120.0)
    assert np.all(m.interior("u") == 0) and np.all(m.interior("v") == 0)
    assert np.all(m.interior("h") == 0.5) and np.all(m.interior("a") == 0.8)


def test_tracer_mass_conserved_on_periodic_domain():
    """Flux-form advection: sum(h Az) is invariant to round-off absent clipping."""
    case = periodic_case(32, substeps=6, aice="ones")
    case.fields["a"] *= 0.6   # keep aice < 1 so the ridging branch (which trades aice for h) never fires
    o = oracle_from_case(case)
    before_h, before_a = o.interior("h").sum(), o.interior("a").sum()
    for _ in range(2):
        o.time_step(case.dt)
    assert abs(o.interior("h").sum() - before_h) / before_h < 1e-13
    assert abs(o.interior("a").sum() - before_a) / before_a < 1e-13
    assert not np.array_equal(o.interior("h"), case.fields["h"][case.Hy:-case.Hy, case.Hx:-case.Hx])  # it did move


def test_transpose_symmetry_of_u_and_v_kernels():
    """Swapping x<->y (and u<->v) of every input must swap the outputs: the u and v kernels and both
    stress-divergence components are mirror images (se.jl:197-264, isd.jl:39-51)."""
    case = periodic_case(24, Ny=24, substeps=4, aice="ones")
    case.coriolis_f = None  # Coriolis breaks the mirror symmetry by its sign
    a = oracle_from_case(case)
    F = case.fields
    swapped = dict(h=F["h"].T.copy(), a=F["a"].T.copy(), u=F["v"].T.copy(), v=F["u"].T.copy(), ue=F["ve"].T.copy(),
                   ve=F["ue"].T.copy(), top_x=F["top_y"].T.copy(), top_y=F["top_x"].T.copy())
    from tests.helpers import oracle_params
    b = O.OracleModel(case.Ny, case.Nx, case.Hy, case.Hx, dx=case.dy, dy=case.dx, params=oracle_params(case), fields=swapped)
    # one stress update + the u/v pair of an *even* substep on A corresponds to the pair of an *odd* substep on B
    a.initialize_rheology(); b.initialize_rheology()
    a.compute_stresses(case.dt); b.compute_stresses(case.dt)
    from tests.helpers import rel_err
    assert np.array_equal(a.arr["s11"], b.arr["s22"].T)
    # the 4-point averages are y-of-x (not symmetric in association), so sigma12 mirrors to round-off only
    assert rel_err(a.arr["s12"], b.arr["s12"].T) < 1e-12
    a.u_velocity_step(case.dt); b.v_velocity_step(case.dt)
    assert rel_err(a.interior("u"), b.interior("v").T) < 1e-11


def test_decomposition_invariance_of_the_stencils():
    """test/distributed_tests_utils.jl:40-88 analogue on the oracle: two y-slabs with halo 2K+3 stepped
    K substeps reproduce the single-domain interior bit for bit."""
    K = 3
    case = periodic_case(24, Ny=64, substeps=K, aice="mixed")
    whole = oracle_from_case(case)
    whole.p.timestepper = O.FE   # no reset_velocities!: keep the synthetic initial u, v
    whole.time_step_momentum(case.dt, K)
    Hy = 2 * K + 3
    for rank in range(2):
        sl = slab_of(case, rank, 2, Hy)
        o = oracle_from_case(sl)
        # a slab is not periodic in y: emulate `only_local_halos` by stepping the kernels over the widened range
        # (the oracle's velocity kernels cover 1:Ny only, so compare the rows whose stencils never left the slab)
        o.g.topo_y = O.BOUNDED  # no periodic y-fill inside the loop; walls only touch rows within 2K+2 of the edge
        o.p.timestepper = O.FE
        o.time_step_momentum(sl.dt, K)
        ny = sl.Ny
        lo, hi = 2 * K + 2, ny - (2 * K + 2)
        ref = whole.interior("u")[rank * ny + lo: rank * ny + hi]
        assert hi - lo >= 8
        assert np.array_equal(o.interior("u")[lo:hi][:, :case.Nx], ref)


def test_halo_width_formula():
    """test/distributed_tests_utils.jl:16-38: Hx = max(2*substeps + 3, H)."""
    import ctypes
    from climaseaice_b200 import lib
    assert lib().csi_host_halo_width(8) == 2 * 8 + 3


def test_anticyclone_runs_and_stays_bounded():
    """Config 1 (examples/ice_advected_by_anticyclone.jl) at reduced substeps: finite, |u| sane, walls impenetrable."""
    case = anticyclone_case(32, substeps=20)
    o = oracle_from_case(case)
    for _ in range(2):
        o.time_step(case.dt)
    for n in ("u", "v", "h", "a", "s11", "s22", "s12"):
        assert np.isfinite(o.interior(n)).all(), n
    assert np.abs(o.interior("u")).max() < 1.0
    assert np.all(o.interior("u")[:, 0] == 0) and np.all(o.interior("u")[:, -1] == 0)
    assert np.all(o.interior("v")[0, :] == 0) and np.all(o.interior("v")[-1, :] == 0)
    assert o.cell_advection_timescale() > 0


def test_weno_reconstruction_is_exact_for_polynomials():
    """Every candidate stencil reproduces the face value of polynomials up to its degree; smooth data gives the
    optimal-weight value (SURVEY Appendix A check 2)."""
    N, H = 16, 7
    m = O.OracleModel(N, N, H, H, dx=1.0, dy=1.0)
    sy, sx = m.arr["h"].shape
    xc = (np.arange(sx) - H + 0.5)
    for order, deg in ((3, 1), (5, 2), (7, 3)):
        coef = np.array([0.3, -0.2, 0.05, 0.01])[:deg + 1]
        prim = np.polyint(coef[::-1])                       # cell averages of p(x) on unit cells
        avg = np.polyval(prim, xc + 0.5) - np.polyval(prim, xc - 0.5)
        m.arr["h"][:] = avg[None, :]
        for bias in (0, 1):
            i = 8
            got = m.reconstruct(0, "h", order, bias, i, 5)
            want = np.polyval(coef[::-1], float(i - 1))     # face i sits at x = i-1
            assert abs(got - want) < 1e-12, (order, bias, got, want)


def test_immersed_drag_slows_the_ice_along_the_coast():
    """The linear immersed drag BC (isd.jl:57-123 with the coastline example's -C*u flux) only acts on faces next to
    land and opposes the motion."""
    from climaseaice_b200.synthetic import coastline_case
    case = coastline_case(Ny=32, substeps=20)
    a = oracle_from_case(case)
    case.immersed_drag = (0.0, 0.0)
    b = oracle_from_case(case)
    for _ in range(2):
        a.time_step(case.dt); b.time_step(case.dt)
    ua, ub = a.interior("u"), b.interior("u")
    assert np.isfinite(ua).all() and not np.array_equal(ua, ub)
    assert np.abs(ua).sum() < np.abs(ub).sum()


def test_latlon_case_steps_and_conserves_volume():
    """The lat-lon synthetic case (j-dependent metrics) runs a full step on the oracle, stays finite and -- closed
    basin, flux-form advection -- conserves sum(h * aice-independent volume) = sum(h Az) to round-off."""
    from climaseaice_b200.synthetic import latlon_case
    from tests.helpers import oracle_from_case
    case = latlon_case(32, substeps=10)
    met = case.metrics()
    az = met["azcc"][case.Hy:case.Hy + case.Ny][:, None]
    o = oracle_from_case(case)
    o.arr["a"][:] = np.minimum(o.arr["a"], 0.6)        # keep ridging out of the way (it trades aice for h)
    v0 = (o.interior("h") * az).sum()
    o.time_step(case.dt)
    for n in ("u", "v", "h", "a", "s11", "s12"):
        assert np.isfinite(o.arr[n]).all(), n
    assert np.abs(o.interior("u")).max() > 1e-5
    assert abs((o.interior("h") * az).sum() - v0) <= 1e-12 * abs(v0)


@pytest.mark.parametrize("variant", ["bottom_drag", "top_drag"])
def test_stress_balance_free_drift_balances_the_stresses(variant):
    """stress_balance_free_drift.jl:61-109: marginal cells take U_d - tau_o / sqrt(C_d |tau_o|), i.e. the velocity at
    which the velocity-dependent stress equals the prescribed one: C_d |U_d - u| (U_d - u) = tau_o, componentwise."""
    from climaseaice_b200.synthetic import marginal_ice_case
    from tests.helpers import oracle_from_case
    case = marginal_ice_case(48, substeps=3, variant=variant, snow=False)
    o = oracle_from_case(case)
    o.update_state()
    o.time_step_momentum(case.dt / 3, 3)
    H, N = case.Hy, case.Ny
    a = o.arr["a"]
    ai = 0.5 * (a[:, 1:] + a[:, :-1])[H:H + N, H - 1:H - 1 + N]                 # aice at the u points 1..N
    m = o.arr["h"] * 900.0 * a
    mi = 0.5 * (m[:, 1:] + m[:, :-1])[H:H + N, H - 1:H - 1 + N]
    marginal = ((mi < 1.0) | (ai < 1e-3)) & (mi > 2.3e-16) & (ai > 2.3e-16)
    assert marginal.sum() > 100
    u = o.interior("u")
    d, oth = (("ue", "ve"), ("top_x", "top_y")) if variant == "bottom_drag" else (("top_x", "top_y"), ("ue", "ve"))
    rho, Cd = (case.rho_e, case.Cd) if variant == "bottom_drag" else case.top_rho_Cd
    Ud = o.arr[d[0]][H:H + N, H:H + N]
    tx = o.arr[oth[0]][H:H + N, H:H + N]
    ty4 = o.arr[oth[1]]
    ty = 0.25 * (ty4[H:H + N, H - 1:H - 1 + N] + ty4[H:H + N, H:H + N] + ty4[H + 1:H + 1 + N, H - 1:H - 1 + N] + ty4[H + 1:H + 1 + N, H:H + N])
    t = np.sqrt(tx ** 2 + ty ** 2)
    drag_x = rho * Cd * np.sqrt(t / (rho * Cd)) * (Ud - u)                      # |U_d - u_F| = sqrt(|tau| / C)
    sel = marginal & (t > 0)
    assert np.allclose(drag_x[sel], tx[sel], rtol=1e-12, atol=1e-15)
    calm = marginal & (t == 0)
    assert np.array_equal(u[calm], Ud[calm])                                    # no stress: drift with the other medium


def test_snow_is_advected_and_clipped_with_the_ice():
    """tracer_tendency:47-52, fe.jl:84-94: hs moves with the same fluxes as h; it is zeroed where aice <= 0; its volume
    sum(hs Az) is conserved by the flux form as long as no cell is clipped."""
    from climaseaice_b200.synthetic import periodic_case
    from tests.helpers import oracle_from_case
    case = periodic_case(48, substeps=4, aice="ones", timestepper="ForwardEuler", advection_order=5)
    case.fields["hs"] = 0.1 + 0.02 * np.sin(2 * np.pi * case.nodes((0, 0))[0] / case.Lx)
    o = oracle_from_case(case)
    o.update_state()
    v0 = o.interior("hs").sum()
    o.time_step(case.dt)
    assert np.abs(o.interior("Ghs")).max() > 0
    assert abs(o.interior("hs").sum() - v0) <= 1e-12 * v0
    case2 = periodic_case(48, substeps=4, aice="mixed", timestepper="ForwardEuler")
    case2.fields["hs"] = np.full_like(case2.fields["h"], 0.1)
    o2 = oracle_from_case(case2)
    o2.time_step(case2.dt)
    assert np.all(o2.interior("hs")[o2.interior("a") <= 0] == 0.0)
    assert (o2.interior("a") <= 0).any()


def test_one_ulp_of_input_exceeds_the_tolerance_after_one_step():
    """The intrinsic-sensitivity yardstick of SURVEY 8(d): one ulp added to ONE thickness value changes u, v, sigma by
    more than north_star's 1e-12 after a single time_step! (3 stages x 150 substeps) -- by 6e-12 on the smooth periodic
    case and by 5e-3 on the anticyclone case, whose plastic regime amplifies round-off exponentially.  Hence the library
    reproduces the oracle's operation sequence bit for bit instead of aiming at a tolerance."""
    for make, floor in ((lambda: periodic_case(32, substeps=150, aice="ones"), 1e-12), (lambda: anticyclone_case(32, substeps=150), 1e-6)):
        case = make()
        a, b = oracle_from_case(case), oracle_from_case(case)
        H = case.Hx
        b.arr["h"][H + 10, H + 12] = np.nextafter(b.arr["h"][H + 10, H + 12], 10.0)
        a.time_step(case.dt)
        b.time_step(case.dt)
        rel = lambda n: np.abs(a.arr[n] - b.arr[n]).max() / np.abs(a.arr[n]).max()
        assert rel("u") > floor and rel("s11") > floor, (case.name, rel("u"), rel("s11"))
        assert rel("h") < 1e-13          # the perturbation itself stays one ulp in h


def _masked_model(order=7):
    """A periodic 24 x 20 oracle with an immersed island and a smooth positive tracer."""
    N, Ny, H = 24, 20, 7
    sy, sx = Ny + 2 * H, N + 2 * H
    mask = np.zeros((sy, sx), dtype=np.uint8)
    mask[H + 8:H + 12, H + 10:H + 14] = 1          # cells i = 11..14, j = 9..12
    m = O.OracleModel(N, Ny, H, H, dx=1.0, dy=1.0, params=dict(advection_order=order), mask=mask)
    return m, mask, N, Ny, H


def test_weno_order_is_reduced_next_to_immersed_cells():
    """Oceananigans' ImmersedBoundaryGrid reconstruction (call site src/sea_ice_advection.jl:51-58, grid of
    examples/ice_advected_on_coastline.jl:54-55): no stencil may read an immersed cell.  Poison the masked cells with a huge
    value: every face value that is used (faces that are not immersed-peripheral) must stay within the range of the active data,
    for both biases and both directions, and far from the island the full-order value is unchanged."""
    for order in (3, 5, 7):
        m, mask, N, Ny, H = _masked_model(order)
        X = np.arange(N + 2 * H)[None, :] - H + 0.5
        Y = np.arange(Ny + 2 * H)[:, None] - H + 0.5
        smooth = 1.0 + 0.1 * np.sin(2 * np.pi * X / N) * np.cos(2 * np.pi * Y / Ny)
        m.arr["h"][:] = smooth
        clean = {(d, b, i, j): m.reconstruct(d, "h", order, b, i, j) for d in (0, 1) for b in (0, 1) for i in range(1, N + 1) for j in range(1, Ny + 1)}
        m.arr["h"][mask.astype(bool)] = 1e30
        lo, hi = smooth.min() - 0.05, smooth.max() + 0.05
        changed = 0
        for (d, b, i, j), ref in clean.items():
            # a face between an active and an immersed cell (or two immersed cells) carries no flux: skip it
            c0 = mask[H + j - 1 - (1 if d == 1 else 0), H + i - 1 - (1 if d == 0 else 0)]
            c1 = mask[H + j - 1, H + i - 1]
            if c0 or c1:
                continue
            got = m.reconstruct(d, "h", order, b, i, j)
            assert lo <= got <= hi, (order, d, b, i, j, got)
            changed += got != ref
        assert changed == 0    # poisoning cells no stencil may read changes nothing


def test_constant_state_is_preserved_along_a_coast():
    """With order reduction no reconstruction mixes the masked zeros into the ice next to the island: a uniform tracer advected by
    a uniform flow keeps zero tendency away from the coast faces, and one oracle step keeps h uniform on cells whose four faces
    are all active."""
    m, mask, N, Ny, H = _masked_model(7)
    m.arr["h"][:] = 1.0
    m.arr["a"][:] = 1.0
    m.arr["h"][mask.astype(bool)] = 0.0
    m.arr["a"][mask.astype(bool)] = 0.0
    m.arr["u"][:] = 0.3
    m.arr["v"][:] = -0.2
    m.compute_tracer_tendencies()
    G = m.interior("Gh")
    mk = mask[H:-H, H:-H].astype(bool)
    near = np.zeros_like(mk)
    for dj, di in ((0, 0), (0, 1), (0, -1), (1, 0), (-1, 0)):
        near |= np.roll(np.roll(mk, dj, axis=0), di, axis=1)
    # uniform flow, uniform tracer, every stencil on active cells: no divergence beyond the round-off of the WENO weights
    # (a stencil reading the island's zeros would give O(0.1))
    assert np.abs(G[~near]).max() < 1e-14
    assert np.isfinite(G).all()

def test_fold_fill_applies_the_copy_lists_with_their_signs():
    """CSI_FOLDED (the north fold of a tripolar grid, given as copy lists): after every step the velocities obey their lists with
    sign -1, thickness and concentration with sign +1; the south stays a wall; the lists of the example never read what they write."""
    from climaseaice_b200.synthetic import folded_case
    case = folded_case(substeps=8)
    for (tg, sr) in case.fold["maps"].values():
        assert np.unique(tg).size == tg.size and not np.intersect1d(tg, sr).size
    o = oracle_from_case(case)
    for _ in range(2):
        o.time_step(case.dt)
    flat = lambda n: o.arr[n].reshape(-1)
    for n, loc, sign in (("u", (1, 0), -1.0), ("v", (0, 1), -1.0), ("h", (0, 0), 1.0), ("a", (0, 0), 1.0)):
        tg, sr = case.fold["maps"][loc]
        assert np.array_equal(flat(n)[tg], sign * flat(n)[sr]), n
        assert np.isfinite(o.arr[n]).all()
    assert np.abs(o.arr["u"]).max() > 1e-3 and np.abs(o.arr["v"][-case.Hy - 3:-case.Hy]).max() > 1e-4   # ice moves at the fold
    assert np.all(o.arr["v"][case.Hy, case.Hx:-case.Hx] == 0)   # the southern wall
    # the fold is not a wall: the flow carries ice across it (thickness changes in the last row)
    h0 = case.fields["h"][case.Hy + case.Ny - 1, case.Hx:-case.Hx]
    assert np.abs(o.arr["h"][case.Hy + case.Ny - 1, case.Hx:-case.Hx] - h0).max() > 1e-6
