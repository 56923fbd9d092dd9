# ClimaSeaIceB200.jl -- thin `ccall` shim over libclimaseaice_b200.so (include/climaseaice_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is the binding a ClimaSeaIce.jl maintainer
# would add.  The seam is multiple dispatch: `enable_b200(model)` returns the same model whose split-explicit solver
# carries a `B200` tag in its `kernel_parameters` slot (`SplitExplicitSolver{I, K}`,
# src/SeaIceDynamics/split_explicit_momentum_equations.jl:18-21), and the methods below -- same names, same arity as the
# ones the reference's drivers call -- are more specific for that tag, so `time_step!(model, Δt)` runs unchanged:
#   time_step_momentum!(model, dynamics, Δt)   called at src/sea_ice_rk_substep.jl:87, src/sea_ice_fe_step.jl:22
#   compute_tendencies!(model, Δt)             called at src/sea_ice_rk_substep.jl:84, src/sea_ice_fe_step.jl:19
#   dynamic_time_step!(model, Δt)              src/sea_ice_rk_substep.jl:134-152, src/sea_ice_fe_step.jl:36-50
#   cache_current_fields!(model)               src/sea_ice_rk_substep.jl:29-42
#   update_state!(model, callbacks)            src/sea_ice_model.jl:379-394
# Each forwards to the C ABI on the model's own CuArrays (the parents of the Oceananigans fields): nothing is copied.
# A configuration the library has no encoding for makes `enable_b200` throw, and the stock Julia methods stay in place.
module ClimaSeaIceB200

using CUDA
using Oceananigans
using Oceananigans.Architectures: architecture
using Oceananigans.BoundaryConditions: BoundaryCondition, Value, fill_halo_regions!
using Oceananigans.Fields: instantiated_location
using Oceananigans.Coriolis: FPlane, HydrostaticSphericalCoriolis, fᶠᶠᵃ
using Oceananigans.DistributedComputations: Distributed
using Oceananigans.Grids: halo_size, topology, inactive_cell, Periodic, Bounded, LeftConnected, RightConnected, FullyConnected
using Oceananigans.ImmersedBoundaries: ImmersedBoundaryGrid
using Oceananigans.Models: update_model_field_time_series!
using Oceananigans.Operators
using Oceananigans.TimeSteppers: SplitRungeKuttaTimeStepper
using Oceananigans.Advection: WENO, UpwindBiased
using ClimaSeaIce
using ClimaSeaIce: SeaIceModel, ForwardEulerTimeStepper
using ClimaSeaIce.Rheologies: ElastoViscoPlasticRheology, IceStrength
using ClimaSeaIce.SeaIceDynamics: SeaIceMomentumEquation, SplitExplicitSolver, SemiImplicitStress, StressBalanceFreeDrift
import ClimaSeaIce.SeaIceDynamics: time_step_momentum!
import ClimaSeaIce: compute_tendencies!, dynamic_time_step!
import Oceananigans.TimeSteppers: cache_current_fields!, update_state!

const LIB = get(ENV, "CLIMASEAICE_B200_LIB", "libclimaseaice_b200.so")

# ---- C structs, field for field as in include/climaseaice_b200.h ------------------------------------------------
struct CsiArray            # csi_array
    ptr    :: CuPtr{Float64}
    nx_tot :: Int32
    ny_tot :: Int32
    off_x  :: Int32
    off_y  :: Int32
end
CsiArray() = CsiArray(CU_NULL, 0, 0, 0, 0)
CsiArray(::Nothing) = CsiArray()
function CsiArray(f::Field)
    p = parent(f)                       # (Nx+2Hx[+1]) x (Ny+2Hy[+1]) x 1, column-major: i fastest
    Hx, Hy, _ = halo_size(f.grid)
    CsiArray(pointer(p), size(p, 1), size(p, 2), Hx, Hy)
end

Base.@kwdef struct CsiConfig
    abi_version :: Int32 = 1;  device :: Int32 = 0
    Nx :: Int32; Ny :: Int32; Hx :: Int32; Hy :: Int32
    topo_x :: Int32; topo_y :: Int32
    dx :: Float64 = 0.0; dy :: Float64 = 0.0
    immersed_mask :: Ptr{UInt8} = C_NULL
    ice_compressive_strength :: Float64; ice_compaction_hardening :: Float64; yield_curve_eccentricity :: Float64
    minimum_plastic_stress :: Float64; min_relaxation_parameter :: Float64; max_relaxation_parameter :: Float64
    relaxation_strength :: Float64
    pressure_formulation :: Int32 = 0;  substeps :: Int32
    minimum_mass :: Float64; minimum_concentration :: Float64; ice_density :: Float64
    coriolis_kind :: Int32 = 0; top_stress_kind :: Int32 = 0
    coriolis_f :: Float64 = 0.0; top_tau_x :: Float64 = 0.0; top_tau_y :: Float64 = 0.0
    bottom_stress_kind :: Int32 = 0; u_south_north_bc :: Int32 = 0
    rho_e :: Float64 = 1026.0; Cd :: Float64 = 5.5e-3; ue_const :: Float64 = 0.0; ve_const :: Float64 = 0.0
    u_south_north_value :: Float64 = 0.0
    v_west_east_bc :: Int32 = 0; advection_order :: Int32 = 0
    v_west_east_value :: Float64 = 0.0
    timestepper :: Int32 = 0; solver_impl :: Int32 = 0
    rank :: Int32 = 0; nranks :: Int32 = 1; exchange_every :: Int32 = 0; partition_x :: Int32 = 0
    immersed_drag_u :: Float64 = 0.0; immersed_drag_v :: Float64 = 0.0
    metric_kind :: Int32 = 0; serial_exchange :: Int32 = 0
    metrics :: NTuple{12, Ptr{Float64}} = ntuple(_ -> Ptr{Float64}(C_NULL), 12)
    free_drift_kind :: Int32 = 0; reserved3_ :: Int32 = 0
    top_rho_e :: Float64 = 1.3; top_Cd :: Float64 = 1.2e-3
    coriolis_f_ff :: Ptr{Float64} = C_NULL
    fold_target :: NTuple{4, Ptr{Int32}} = ntuple(_ -> Ptr{Int32}(C_NULL), 4)
    fold_source :: NTuple{4, Ptr{Int32}} = ntuple(_ -> Ptr{Int32}(C_NULL), 4)
    fold_count :: NTuple{4, Int32} = ntuple(_ -> Int32(0), 4)
    fold_sign_velocity :: Float64 = -1.0; fold_sign_external :: Float64 = 1.0
end

# csi_fields: 29 csi_array in header order
const FIELD_ORDER = (:u, :v, :h, :a, :s11, :s22, :s12, :zeta_f, :zeta_c, :delta, :alpha, :un, :vn, :P,
                     :top_x, :top_y, :ue, :ve, :Gh, :Ga, :hm, :am, :um, :vm, :hs, :Ghs, :hsm, :fd_u, :fd_v)
const CsiFields = NTuple{29, CsiArray}
@assert length(FIELD_ORDER) == 29

mutable struct Handle
    ptr  :: Ptr{Cvoid}
    keep :: Any            # host arrays the library copied at csi_create are kept until then; the device mask afterwards
end

check(rc, h = C_NULL) = rc == 0 ? nothing :
    error("libclimaseaice_b200 ($rc): ", unsafe_string(ccall((:csi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))
unsupported(what) = throw(ArgumentError("ClimaSeaIceB200: no csi_config encoding for $what; the stock ClimaSeaIce methods remain in use"))

# ---- the tag ---------------------------------------------------------------------------------------------------
"""`B200(kernel_parameters, handle)` replaces `solver.kernel_parameters`; the stock value is kept for `disable_b200`."""
struct B200{K}
    kernel_parameters :: K
    handle :: Handle
end
const B200Solver   = SplitExplicitSolver{<:Any, <:B200}
const B200Dynamics = SeaIceMomentumEquation{<:B200Solver}
const B200Model    = SeaIceModel{<:Any, <:Any, <:Any, <:B200Dynamics}                                # D is the 4th parameter
const B200RKModel  = SeaIceModel{<:Any, <:Any, <:Any, <:B200Dynamics, <:SplitRungeKuttaTimeStepper}   # TS the 5th
const B200FEModel  = SeaIceModel{<:Any, <:Any, <:Any, <:B200Dynamics, <:ForwardEulerTimeStepper}
handle(model::B200Model) = model.dynamics.solver.kernel_parameters.handle

# rebuild an immutable struct with some fields replaced (positional default constructor)
rebuild(x; kw...) = typeof(x).name.wrapper((haskey(kw, f) ? kw[f] : getfield(x, f) for f in fieldnames(typeof(x)))...)

"""
    enable_b200(model; solver_impl = 0, exchange_every = 4, immersed_drag = nothing)

Returns `model` with its split-explicit momentum solver tagged for libclimaseaice_b200 (the fields are shared, not copied).
"""
function enable_b200(model::SeaIceModel; solver_impl = 0, exchange_every = 4, immersed_drag = nothing)
    dyn = model.dynamics
    dyn isa SeaIceMomentumEquation{<:SplitExplicitSolver} || unsupported("dynamics = $(summary(dyn)) (only SeaIceMomentumEquation with a SplitExplicitSolver)")
    h = create(model; solver_impl, exchange_every, immersed_drag)
    solver = SplitExplicitSolver(dyn.solver.substeps, B200(dyn.solver.kernel_parameters, h))
    return rebuild(model; dynamics = rebuild(dyn; solver))
end
disable_b200(model::B200Model) =
    rebuild(model; dynamics = rebuild(model.dynamics; solver = SplitExplicitSolver(model.dynamics.solver.substeps, model.dynamics.solver.kernel_parameters.kernel_parameters)))

# ---- csi_config from the model -----------------------------------------------------------------------------------
topo_code(::Type{Periodic}) = Int32(0)
topo_code(::Type{Bounded})  = Int32(1)
# a partitioned axis reports LeftConnected / RightConnected / FullyConnected locally; the library wants the GLOBAL topology
# (its rank index tells it which sides are rank boundaries): FullyConnected everywhere <=> Periodic, else Bounded
global_topo_code(T, periodic_globally) = T in (LeftConnected, RightConnected, FullyConnected) ? Int32(periodic_globally ? 0 : 1) : topo_code(T)

# ---- the north fold of a TripolarGrid (topo_y = CSI_FOLDED = 2) -------------------------------------------------
# The library builds in NO index convention of the fold: it applies, after its own periodic / wall fills, a list of copies
# parent[target] = sign * parent[source] per location.  The lists are read off Oceananigans' own fill_halo_regions!: a scratch
# field with the boundary conditions of the real one is filled with (its own 0-based linear parent index + 1), its halos are
# filled, and every element at or beyond row Ny whose value changed names its source (|value| - 1) and its sign.
is_folded(ugrid) = ugrid isa OrthogonalSphericalShellGrid && nameof(typeof(ugrid.conformal_mapping)) === :Tripolar

function fold_probe(field)
    cpu = on_architecture(CPU(), field.grid)
    loc = instantiated_location(field)
    f = Field(loc, cpu; boundary_conditions = field.boundary_conditions)
    p = parent(f)
    sx, sy = size(p, 1), size(p, 2)
    p[:, :, 1] .= reshape(Float64.(1:sx*sy), sx, sy)
    fill_halo_regions!(f)
    _, Ny, _ = size(cpu); _, Hy, _ = halo_size(cpu)
    target, source, signs = Int32[], Int32[], Float64[]
    for pj in Ny+Hy:sy, pi in 1:sx              # the pivot row (parent row Ny + Hy, 1-based) and everything north of it
        own, val = (pj - 1) * sx + pi, p[pi, pj, 1]
        abs(val) == own && val > 0 && continue
        push!(target, own - 1); push!(source, Int32(abs(val)) - 1); push!(signs, sign(val))
    end
    allequal(signs) || unsupported("a fold whose sign varies within one field")
    return target, source, isempty(signs) ? 1.0 : first(signs)
end

function fold_maps(model)
    u, v, hh = model.velocities.u, model.velocities.v, model.ice_thickness
    grid = u.grid
    ff = Field((Face(), Face(), Center()), grid)      # the stress nodes: default boundary conditions of the location
    tu, su, gu = fold_probe(u); tv, sv, gv = fold_probe(v); tc, sc, gc = fold_probe(hh); tf, sf, _ = fold_probe(ff)
    gu == gv || unsupported("different fold signs for u and v")
    gc == 1.0 || unsupported("a sign-reversing fold for thickness")
    bot = model.dynamics.external_momentum_stresses.bottom
    ge = (bot isa SemiImplicitStress && bot.uₑ isa Field) ? fold_probe(bot.uₑ)[3] : 1.0
    # order of csi_config.fold_*: (Center, Center), (Face, Center), (Center, Face), (Face, Face)
    return (targets = (tc, tu, tv, tf), sources = (sc, su, sv, sf), sign_velocity = gu, sign_external = ge)
end

const METRIC_OPS = (Δxᶜᶜᶜ, Δxᶠᶜᶜ, Δxᶜᶠᶜ, Δxᶠᶠᶜ, Δyᶜᶜᶜ, Δyᶠᶜᶜ, Δyᶜᶠᶜ, Δyᶠᶠᶜ, Azᶜᶜᶜ, Azᶠᶜᶜ, Azᶜᶠᶜ, Azᶠᶠᶜ)
# LatitudeLongitudeGrid: twelve j-indexed vectors (metric_kind = 1), index j at [j + Hy] (1-based), evaluated with
# Oceananigans' own operators so the library divides by exactly the numbers the reference kernels would
function metric_vectors(grid)
    _, Ny, _ = size(grid); _, Hy, _ = halo_size(grid)
    cpu = on_architecture(CPU(), grid)
    return [Float64[op(1, j, 1, cpu) for j in 1-Hy:Ny+Hy+1] for op in METRIC_OPS]
end
# orthogonal curvilinear grids: the same metrics as (Nx + 2Hx + 1) x (Ny + 2Hy + 1) arrays, i fastest (metric_kind = 2)
function metric_arrays(grid)
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid)
    cpu = on_architecture(CPU(), grid)
    return [Float64[op(i, j, 1, cpu) for i in 1-Hx:Nx+Hx+1, j in 1-Hy:Ny+Hy+1] for op in METRIC_OPS]
end

underlying(grid) = grid isa ImmersedBoundaryGrid ? grid.underlying_grid : grid

stress_kind(::Nothing) = (Int32(0), 0.0, 0.0)                                                     # CSI_STRESS_NONE
stress_kind(τ::NamedTuple) = τ.u isa Number && τ.v isa Number ? (Int32(1), Float64(τ.u), Float64(τ.v)) :   # CSI_STRESS_CONST
                             τ.u isa Field  && τ.v isa Field  ? (Int32(2), 0.0, 0.0) :                     # CSI_STRESS_FIELD
                             unsupported("a momentum stress mixing numbers and fields")
stress_kind(τ::SemiImplicitStress) = τ.uₑ isa Field ? (Int32(3), 0.0, 0.0) : (Int32(3), Float64(τ.uₑ), Float64(τ.vₑ))
stress_kind(τ) = unsupported("momentum stress of type $(typeof(τ))")

value_bc(bc) = bc isa BoundaryCondition && bc.classification isa Value && bc.condition isa Number ? Float64(bc.condition) : nothing

advection_order(::Nothing) = Int32(0)
advection_order(::UpwindBiased{1}) = Int32(1)
advection_order(::WENO{N}) where N = Int32(2N - 1)                    # buffer N: WENO(order = 2N - 1)
advection_order(a) = unsupported("advection scheme $(summary(a)) (WENO(order = 3, 5, 7), UpwindBiased(order = 1) or nothing)")

"""Build the handle once (the construction point of the reference is `SeaIceModel(grid; …)`, src/sea_ice_model.jl:140-297)."""
function create(model::SeaIceModel; solver_impl = 0, exchange_every = 4, immersed_drag = nothing)
    grid = model.velocities.u.grid                    # the (possibly halo-extended) velocity grid, src/sea_ice_model.jl:180-200
    ugrid = underlying(grid)
    dyn, r = model.dynamics, model.dynamics.rheology
    r isa ElastoViscoPlasticRheology || unsupported("rheology $(summary(r))")
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid); TX, TY, _ = topology(grid)
    keep = Any[]

    # grid metrics
    regular, latlon = ugrid isa RectilinearGrid, ugrid isa LatitudeLongitudeGrid
    metric_kind, metrics = Int32(0), ntuple(_ -> Ptr{Float64}(C_NULL), 12)
    if regular
        (ugrid.Δxᶜᵃᵃ isa Number && ugrid.Δyᵃᶜᵃ isa Number) || unsupported("a stretched RectilinearGrid")
    else
        arrs = latlon ? metric_vectors(ugrid) : metric_arrays(ugrid)
        push!(keep, arrs)
        metric_kind, metrics = Int32(latlon ? 1 : 2), ntuple(k -> pointer(arrs[k]), 12)
    end

    # a tripolar grid: copy lists of the fold, taken from Oceananigans' own halo fill (see fold_probe)
    folded = is_folded(ugrid)
    fold = nothing
    if folded
        fold = fold_maps(model)
        push!(keep, fold)
    end

    # partition: Distributed(arch; partition = Partition(Rx, Ry)), rank = ry * Rx + rx (test/distributed_tests_utils.jl:60-62)
    arch = architecture(grid)
    rank, nranks, Rx = Int32(0), Int32(1), Int32(0)
    xper = yper = true
    if arch isa Distributed
        Rx, Ry = Int32(arch.ranks[1]), Int32(arch.ranks[2])
        rx, ry = arch.local_index[1] - 1, arch.local_index[2] - 1
        rank, nranks = Int32(ry * Rx + rx), Int32(Rx * Ry)
        xper = arch.connectivity.west !== nothing && arch.connectivity.east !== nothing && (Rx == 1 ? TX == Periodic : true)
        yper = arch.connectivity.south !== nothing && arch.connectivity.north !== nothing && (Ry == 1 ? TY == Periodic : true)
    end

    # immersed boundary: centre mask over the parent index range, 1 = inactive; linear drag -C u of the coastline example
    mask = Ptr{UInt8}(C_NULL)
    drag_u = drag_v = 0.0
    if grid isa ImmersedBoundaryGrid
        cpu = on_architecture(CPU(), grid)
        # (cells outside a Bounded domain are not "immersed": the library derives those from the topology)
        outside(i, j) = (TX == Bounded && (i < 1 || i > Nx)) || (TY == Bounded && (j < 1 || j > Ny))
        m = UInt8[(inactive_cell(i, j, 1, cpu) && !outside(i, j)) ? 1 : 0 for i in 1-Hx:Nx+Hx, j in 1-Hy:Ny+Hy]
        push!(keep, m); mask = pointer(m)
        # An immersed velocity boundary condition is a Julia closure (examples/ice_advected_on_coastline.jl:91-98) and cannot
        # cross a C ABI; its only form in the reference's examples, the linear drag flux -C u / -C v, is passed as a number
        ibu, ibv = model.velocities.u.boundary_conditions.immersed, model.velocities.v.boundary_conditions.immersed
        if !(isnothing(ibu) && isnothing(ibv))
            isnothing(immersed_drag) && unsupported("an immersed velocity boundary condition (pass the linear drag coefficients: enable_b200(model; immersed_drag = (Cu, Cv)))")
            drag_u, drag_v = Float64.(immersed_drag)
        end
    end

    # Coriolis
    cor = dyn.coriolis
    coriolis_kind, coriolis_f, f_ff = Int32(0), 0.0, Ptr{Float64}(C_NULL)
    if cor isa FPlane
        coriolis_kind, coriolis_f = Int32(1), Float64(cor.f)
    elseif cor isa HydrostaticSphericalCoriolis
        latlon || unsupported("HydrostaticSphericalCoriolis on a grid that is not a LatitudeLongitudeGrid")
        cpu = on_architecture(CPU(), ugrid)
        fv = Float64[fᶠᶠᵃ(1, j, 1, cpu, cor) for j in 1-Hy:Ny+Hy+1]
        push!(keep, fv); coriolis_kind, f_ff = Int32(2), pointer(fv)
    elseif !isnothing(cor)
        unsupported("coriolis = $(summary(cor))")
    end

    # stresses, free drift
    top, bot = dyn.external_momentum_stresses.top, dyn.external_momentum_stresses.bottom
    tk, ttx, tty = stress_kind(top)
    bk, bcx, bcy = stress_kind(bot)
    fd = dyn.free_drift
    fd_kind = isnothing(fd) ? Int32(0) : fd isa NamedTuple ? Int32(1) : fd isa StressBalanceFreeDrift ? Int32(2) : unsupported("free_drift = $(summary(fd))")

    # velocity boundary conditions on Bounded axes: ValueBoundaryCondition(number) or the default (examples/ice_advected_by_anticyclone.jl)
    ubc, vbc = model.velocities.u.boundary_conditions, model.velocities.v.boundary_conditions
    us, un_ = value_bc(ubc.south), value_bc(ubc.north)
    vw, ve_ = value_bc(vbc.west), value_bc(vbc.east)
    (us == un_ && vw == ve_) || unsupported("different Value boundary conditions on opposite walls")

    ρ = model.sea_ice_density
    ρ_host = Array(interior(ρ))
    all(==(first(ρ_host)), ρ_host) || unsupported("a non-uniform sea_ice_density field")

    ts = model.timestepper
    timestepper = ts isa SplitRungeKuttaTimeStepper ? Int32(0) : ts isa ForwardEulerTimeStepper ? Int32(1) : unsupported("timestepper $(typeof(ts))")

    cfg = CsiConfig(; Nx, Ny, Hx, Hy, device = CUDA.deviceid(),
                    topo_x = global_topo_code(TX, xper), topo_y = folded ? Int32(2) : global_topo_code(TY, yper),
                    fold_target = folded ? ntuple(k -> pointer(fold.targets[k]), 4) : ntuple(_ -> Ptr{Int32}(C_NULL), 4),
                    fold_source = folded ? ntuple(k -> pointer(fold.sources[k]), 4) : ntuple(_ -> Ptr{Int32}(C_NULL), 4),
                    fold_count = folded ? ntuple(k -> Int32(length(fold.targets[k])), 4) : ntuple(_ -> Int32(0), 4),
                    fold_sign_velocity = folded ? fold.sign_velocity : -1.0, fold_sign_external = folded ? fold.sign_external : 1.0,
                    dx = regular ? Float64(ugrid.Δxᶜᵃᵃ) : 0.0, dy = regular ? Float64(ugrid.Δyᵃᶜᵃ) : 0.0,
                    immersed_mask = mask, metric_kind, metrics,
                    ice_compressive_strength = r.ice_compressive_strength, ice_compaction_hardening = r.ice_compaction_hardening,
                    yield_curve_eccentricity = r.yield_curve_eccentricity, minimum_plastic_stress = r.minimum_plastic_stress,
                    min_relaxation_parameter = r.min_relaxation_parameter, max_relaxation_parameter = r.max_relaxation_parameter,
                    relaxation_strength = r.relaxation_strength,
                    pressure_formulation = r.pressure_formulation isa IceStrength ? 1 : 0,
                    substeps = dyn.solver.substeps,
                    minimum_mass = dyn.minimum_mass, minimum_concentration = dyn.minimum_concentration, ice_density = first(ρ_host),
                    coriolis_kind, coriolis_f, coriolis_f_ff = f_ff,
                    top_stress_kind = tk, top_tau_x = ttx, top_tau_y = tty,
                    bottom_stress_kind = bk, ue_const = bcx, ve_const = bcy,
                    rho_e = bot isa SemiImplicitStress ? bot.ρₑ : 1026.0, Cd = bot isa SemiImplicitStress ? bot.Cᴰ : 5.5e-3,
                    top_rho_e = top isa SemiImplicitStress ? top.ρₑ : 1.3, top_Cd = top isa SemiImplicitStress ? top.Cᴰ : 1.2e-3,
                    free_drift_kind = fd_kind,
                    u_south_north_bc = isnothing(us) ? 0 : 1, u_south_north_value = something(us, 0.0),
                    v_west_east_bc = isnothing(vw) ? 0 : 1, v_west_east_value = something(vw, 0.0),
                    advection_order = advection_order(model.advection), timestepper, solver_impl,
                    rank, nranks, partition_x = Rx, exchange_every = nranks > 1 ? exchange_every : 0,
                    immersed_drag_u = drag_u, immersed_drag_v = drag_v)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep check(ccall((:csi_create, LIB), Cint, (Ref{CsiConfig}, Ref{Ptr{Cvoid}}), cfg, out))
    h = Handle(out[], nothing)           # the library copied the metrics, f and the mask
    finalizer(x -> ccall((:csi_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

"""After `enable_b200` on a partitioned model: broadcast the NCCL id from rank 0 (e.g. `MPI.Bcast!`) and call this on every rank."""
nccl_unique_id() = (id = zeros(UInt8, 128); check(ccall((:csi_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), id)); id)
comm_init!(model::B200Model, id::Vector{UInt8}, rank, nranks) =
    check(ccall((:csi_comm_init, LIB), Cint, (Ptr{Cvoid}, Ptr{UInt8}, Int32, Int32), handle(model).ptr, id, rank, nranks), handle(model).ptr)

# ---- csi_fields from the model ------------------------------------------------------------------------------------
function csi_fields(model)
    aux = model.dynamics.auxiliaries.fields
    top, bot = model.dynamics.external_momentum_stresses.top, model.dynamics.external_momentum_stresses.bottom
    G, ts = model.timestepper.Gⁿ, model.timestepper
    d = Dict{Symbol, CsiArray}(
        :u => CsiArray(model.velocities.u), :v => CsiArray(model.velocities.v),
        :h => CsiArray(model.ice_thickness), :a => CsiArray(model.ice_concentration),
        :s11 => CsiArray(aux.σ₁₁), :s22 => CsiArray(aux.σ₂₂), :s12 => CsiArray(aux.σ₁₂),
        :zeta_f => CsiArray(aux.ζᶠᶠᶜ), :zeta_c => CsiArray(aux.ζᶜᶜᶜ), :delta => CsiArray(aux.Δ), :alpha => CsiArray(aux.α),
        :un => CsiArray(aux.uⁿ), :vn => CsiArray(aux.vⁿ), :P => CsiArray(aux.P),
        :Gh => CsiArray(G.h), :Ga => CsiArray(G.ℵ))
    if ts isa SplitRungeKuttaTimeStepper       # Ψ⁻: the state cached by cache_current_fields!
        Ψ = ts.Ψ⁻
        d[:hm] = CsiArray(Ψ.h); d[:am] = CsiArray(Ψ.ℵ); d[:um] = CsiArray(Ψ.u); d[:vm] = CsiArray(Ψ.v)
    end
    if top isa NamedTuple && top.u isa Field
        d[:top_x] = CsiArray(top.u); d[:top_y] = CsiArray(top.v)
    elseif top isa SemiImplicitStress && top.uₑ isa Field
        d[:top_x] = CsiArray(top.uₑ); d[:top_y] = CsiArray(top.vₑ)
    end
    if bot isa NamedTuple && bot.u isa Field
        d[:ue] = CsiArray(bot.u); d[:ve] = CsiArray(bot.v)
    elseif bot isa SemiImplicitStress && bot.uₑ isa Field
        d[:ue] = CsiArray(bot.uₑ); d[:ve] = CsiArray(bot.vₑ)
    end
    hs = model.snow_thickness
    if !isnothing(hs)
        d[:hs] = CsiArray(hs); d[:Ghs] = CsiArray(G.hs)
        ts isa SplitRungeKuttaTimeStepper && (d[:hsm] = CsiArray(ts.Ψ⁻.hs))
    end
    fd = model.dynamics.free_drift
    fd isa NamedTuple && (d[:fd_u] = CsiArray(fd.u); d[:fd_v] = CsiArray(fd.v))
    return ntuple(k -> get(d, FIELD_ORDER[k], CsiArray()), 29)
end

stream() = CUDA.stream().handle

# ---- the drop-in methods: the reference's own names and arities ------------------------------------------------------
function time_step_momentum!(model, dynamics::B200Dynamics, Δt)
    h = dynamics.solver.kernel_parameters.handle
    GC.@preserve model check(ccall((:csi_evp_substeps, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Cdouble, Int32, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), Δt, dynamics.solver.substeps, stream()), h.ptr)
    return nothing
end

# compute_tendencies! = tracer tendencies + compute_momentum_tendencies!, which is a no-op for split-explicit dynamics
function compute_tendencies!(model::B200Model, Δt)
    h = handle(model)
    GC.@preserve model check(ccall((:csi_compute_tracer_tendencies, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), stream()), h.ptr)
    return nothing
end

function b200_dynamic_time_step!(model, Δt)
    h = handle(model)
    GC.@preserve model check(ccall((:csi_dynamic_time_step, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Cdouble, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), Δt, stream()), h.ptr)
    return nothing
end
dynamic_time_step!(model::B200RKModel, Δt) = b200_dynamic_time_step!(model, Δt)     # (one method per stock method: no ambiguity)
dynamic_time_step!(model::B200FEModel, Δt) = b200_dynamic_time_step!(model, Δt)

function cache_current_fields!(model::B200RKModel)
    isempty(model.tracers) || return invoke(cache_current_fields!, Tuple{ClimaSeaIce.RKSeaIceModel}, model)   # extra tracers: stock path
    h = handle(model)
    GC.@preserve model check(ccall((:csi_cache_current_fields, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), stream()), h.ptr)
    return nothing
end

function update_state!(model::B200Model, callbacks = [])
    h = handle(model)
    GC.@preserve model check(ccall((:csi_update_state, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), stream()), h.ptr)
    # (the three mass-flux diagnostics are masked by the library only when its thermodynamics are attached; the stock
    #  thermodynamic_time_step! keeps running in Julia otherwise, and masks nothing the library does not own)
    update_model_field_time_series!(model, model.clock)
    return nothing
end

end # module
