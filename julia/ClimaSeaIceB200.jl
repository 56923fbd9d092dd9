# ClimaSeaIceB200.jl -- thin `ccall` shim over libclimaseaice_b200.so (include/climaseaice_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It shows the binding a
# ClimaSeaIce.jl maintainer would add: a `B200` solver tag whose methods replace the hot-path methods
#   time_step_momentum!        src/SeaIceDynamics/split_explicit_momentum_equations.jl:103-195
#   compute_tracer_tendencies! src/tracer_tendency_kernel_functions.jl:9-25
#   dynamic_time_step!         src/sea_ice_rk_substep.jl:134-152, src/sea_ice_fe_step.jl:36-50
#   cache_current_fields!      src/sea_ice_rk_substep.jl:29-42
#   update_state! halo fills   src/sea_ice_model.jl:379-394
# and forward to the C ABI on the model's own CuArrays (parents of the Oceananigans fields).
module ClimaSeaIceB200

using CUDA
using Oceananigans
using Oceananigans.Grids: halo_size, topology, Periodic, Bounded
using ClimaSeaIce
using ClimaSeaIce.SeaIceDynamics: SeaIceMomentumEquation, SplitExplicitSolver, SemiImplicitStress
import ClimaSeaIce.SeaIceDynamics: time_step_momentum!
import ClimaSeaIce: compute_tracer_tendencies!, dynamic_time_step!

const LIB = get(ENV, "CLIMASEAICE_B200_LIB", "libclimaseaice_b200.so")

struct CsiArray            # csi_array
    ptr    :: CuPtr{Float64}
    nx_tot :: Int32
    ny_tot :: Int32
    off_x  :: Int32
    off_y  :: Int32
end
CsiArray() = CsiArray(CU_NULL, 0, 0, 0, 0)
function CsiArray(f::Field)
    p = parent(f)                       # (Nx+2Hx[+1]) x (Ny+2Hy[+1]) x 1, column-major
    Hx, Hy, _ = halo_size(f.grid)
    CsiArray(pointer(p), size(p, 1), size(p, 2), Hx, Hy)
end

# csi_config: field order and types exactly as in include/climaseaice_b200.h
Base.@kwdef struct CsiConfig
    abi_version :: Int32 = 1;  device :: Int32 = 0
    Nx :: Int32; Ny :: Int32; Hx :: Int32; Hy :: Int32
    topo_x :: Int32; topo_y :: Int32
    dx :: Float64; dy :: Float64
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
    v_west_east_bc :: Int32 = 0; advection_order :: Int32 = 7
    v_west_east_value :: Float64 = 0.0
    timestepper :: Int32 = 0; solver_impl :: Int32 = 0
    rank :: Int32 = 0; nranks :: Int32 = 1; exchange_every :: Int32 = 0; partition_x :: Int32 = 0
    immersed_drag_u :: Float64 = 0.0; immersed_drag_v :: Float64 = 0.0
    metric_kind :: Int32 = 0; serial_exchange :: Int32 = 0
    metrics :: NTuple{12, Ptr{Float64}} = ntuple(_ -> Ptr{Float64}(C_NULL), 12)
    free_drift_kind :: Int32 = 0; reserved3_ :: Int32 = 0     # 0 nothing, 1 (u=, v=) arrays, 2 StressBalanceFreeDrift
    top_rho_e :: Float64 = 1.3; top_Cd :: Float64 = 1.2e-3    # SemiImplicitStress as the top stress
    coriolis_f_ff :: Ptr{Float64} = C_NULL                     # HydrostaticSphericalCoriolis: fᶠᶠᵃ per row (coriolis_kind = 2)
end

# LatitudeLongitudeGrid: the twelve j-indexed metric vectors csi_config.metrics takes (metric_kind = 1), each
# Ny + 2Hy + 1 long with index j at [j - 1 + Hy] (1-based: [j + Hy]).  Evaluated with Oceananigans' own operators so
# the library sees exactly the numbers the reference kernels would; keep `vecs` alive until csi_create returns.
function metric_vectors(grid)
    _, Ny, _ = size(grid); _, Hy, _ = halo_size(grid)
    ops = (Δxᶜᶜᶜ, Δxᶠᶜᶜ, Δxᶜᶠᶜ, Δxᶠᶠᶜ, Δyᶜᶜᶜ, Δyᶠᶜᶜ, Δyᶜᶠᶜ, Δyᶠᶠᶜ, Azᶜᶜᶜ, Azᶠᶜᶜ, Azᶜᶠᶜ, Azᶠᶠᶜ)
    cpu = on_architecture(CPU(), grid)
    vecs = [Float64[op(1, j, 1, cpu) for j in 1-Hy:Ny+Hy+1] for op in ops]
    return vecs, ntuple(k -> pointer(vecs[k]), 12)
end

# Orthogonal curvilinear grids (OrthogonalSphericalShellGrid without a fold): the same twelve metrics as two-dimensional
# arrays (metric_kind = 2), (Nx + 2Hx + 1) x (Ny + 2Hy + 1) column-major = i fastest, index (i, j) at [i + Hx, j + Hy].
function metric_arrays(grid)
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid)
    ops = (Δxᶜᶜᶜ, Δxᶠᶜᶜ, Δxᶜᶠᶜ, Δxᶠᶠᶜ, Δyᶜᶜᶜ, Δyᶠᶜᶜ, Δyᶜᶠᶜ, Δyᶠᶠᶜ, Azᶜᶜᶜ, Azᶠᶜᶜ, Azᶜᶠᶜ, Azᶠᶠᶜ)
    cpu = on_architecture(CPU(), grid)
    arrs = [Float64[op(i, j, 1, cpu) for i in 1-Hx:Nx+Hx+1, j in 1-Hy:Ny+Hy+1] for op in ops]
    return arrs, ntuple(k -> pointer(arrs[k]), 12)
end

# csi_fields: 29 csi_array in header order
const FIELD_ORDER = (:u, :v, :h, :a, :s11, :s22, :s12, :zeta_f, :zeta_c, :delta, :alpha, :un, :vn, :P,
                     :top_x, :top_y, :ue, :ve, :Gh, :Ga, :hm, :am, :um, :vm, :hs, :Ghs, :hsm, :fd_u, :fd_v)
const CsiFields = NTuple{29, CsiArray}

mutable struct Handle
    ptr :: Ptr{Cvoid}
end

check(rc, h = C_NULL) = rc == 0 ? nothing :
    error("libclimaseaice_b200 ($rc): ", unsafe_string(ccall((:csi_last_error, LIB), Cstring, (Ptr{Cvoid},), h)))

topo_code(::Type{Periodic}) = Int32(0)
topo_code(::Type{Bounded})  = Int32(1)

"""Build the handle once, at `SeaIceModel` construction time (src/sea_ice_model.jl:140-297)."""
function create(model::SeaIceModel)
    grid = model.velocities.u.grid
    dyn, r = model.dynamics, model.dynamics.rheology
    Nx, Ny, _ = size(grid); Hx, Hy, _ = halo_size(grid); TX, TY, _ = topology(grid)
    bottom = dyn.external_momentum_stresses.bottom
    regular = grid isa RectilinearGrid      # else: LatitudeLongitudeGrid (metrics depend on j) or an orthogonal curvilinear grid
    latlon = grid isa LatitudeLongitudeGrid
    vecs, metrics = regular ? (nothing, ntuple(_ -> Ptr{Float64}(C_NULL), 12)) : (latlon ? metric_vectors(grid) : metric_arrays(grid))
    cfg = CsiConfig(; Nx, Ny, Hx, Hy, topo_x = topo_code(TX), topo_y = topo_code(TY),
                    dx = regular ? grid.Δxᶜᵃᵃ : 0.0, dy = regular ? grid.Δyᵃᶜᵃ : 0.0, device = CUDA.deviceid(),
                    metric_kind = regular ? 0 : (latlon ? 1 : 2), metrics,
                    ice_compressive_strength = r.ice_compressive_strength, ice_compaction_hardening = r.ice_compaction_hardening,
                    yield_curve_eccentricity = r.yield_curve_eccentricity, minimum_plastic_stress = r.minimum_plastic_stress,
                    min_relaxation_parameter = r.min_relaxation_parameter, max_relaxation_parameter = r.max_relaxation_parameter,
                    relaxation_strength = r.relaxation_strength, substeps = dyn.solver.substeps,
                    minimum_mass = dyn.minimum_mass, minimum_concentration = dyn.minimum_concentration,
                    ice_density = model.sea_ice_density[1, 1, 1],
                    coriolis_kind = isnothing(dyn.coriolis) ? 0 : 1, coriolis_f = isnothing(dyn.coriolis) ? 0.0 : dyn.coriolis.f,
                    top_stress_kind = dyn.external_momentum_stresses.top isa NamedTuple ? 2 : 0,
                    bottom_stress_kind = bottom isa SemiImplicitStress ? 3 : 0,
                    rho_e = bottom isa SemiImplicitStress ? bottom.ρₑ : 1026.0, Cd = bottom isa SemiImplicitStress ? bottom.Cᴰ : 5.5e-3)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve vecs check(ccall((:csi_create, LIB), Cint, (Ref{CsiConfig}, Ref{Ptr{Cvoid}}), cfg, out))
    h = Handle(out[])
    finalizer(x -> ccall((:csi_destroy, LIB), Cint, (Ptr{Cvoid},), x.ptr), h)
    return h
end

function csi_fields(model)
    aux = model.dynamics.auxiliaries.fields
    top, bot = model.dynamics.external_momentum_stresses
    G, Ψ = model.timestepper.Gⁿ, model.timestepper.Ψ⁻
    d = Dict{Symbol, CsiArray}(
        :u => CsiArray(model.velocities.u), :v => CsiArray(model.velocities.v),
        :h => CsiArray(model.ice_thickness), :a => CsiArray(model.ice_concentration),
        :s11 => CsiArray(aux.σ₁₁), :s22 => CsiArray(aux.σ₂₂), :s12 => CsiArray(aux.σ₁₂),
        :zeta_f => CsiArray(aux.ζᶠᶠᶜ), :zeta_c => CsiArray(aux.ζᶜᶜᶜ), :delta => CsiArray(aux.Δ), :alpha => CsiArray(aux.α),
        :un => CsiArray(aux.uⁿ), :vn => CsiArray(aux.vⁿ), :P => CsiArray(aux.P),
        :Gh => CsiArray(G.h), :Ga => CsiArray(G.ℵ),
        :hm => CsiArray(Ψ.h), :am => CsiArray(Ψ.ℵ), :um => CsiArray(Ψ.u), :vm => CsiArray(Ψ.v))
    top isa NamedTuple && (d[:top_x] = CsiArray(top.u); d[:top_y] = CsiArray(top.v))
    bot isa SemiImplicitStress && bot.uₑ isa Field && (d[:ue] = CsiArray(bot.uₑ); d[:ve] = CsiArray(bot.vₑ))
    return ntuple(k -> get(d, FIELD_ORDER[k], CsiArray()), 24)
end

stream() = CUDA.stream().handle

# The drop-in methods.  `model.b200` is the Handle stored next to the model by the host package.
function time_step_momentum!(model, dynamics::SeaIceMomentumEquation{<:SplitExplicitSolver}, Δt, h::Handle)
    f = csi_fields(model)
    GC.@preserve model check(ccall((:csi_evp_substeps, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Cdouble, Int32, Ptr{Cvoid}),
                                   h.ptr, f, Δt, dynamics.solver.substeps, stream()), h.ptr)
end
compute_tracer_tendencies!(model, h::Handle) =
    GC.@preserve model check(ccall((:csi_compute_tracer_tendencies, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Ptr{Cvoid}), h.ptr, csi_fields(model), stream()), h.ptr)
dynamic_time_step!(model, Δt, h::Handle) =
    GC.@preserve model check(ccall((:csi_dynamic_time_step, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Cdouble, Ptr{Cvoid}), h.ptr, csi_fields(model), Δt, stream()), h.ptr)
time_step_b200!(model, Δt, h::Handle) =   # the whole time_step! on the device
    GC.@preserve model check(ccall((:csi_time_step, LIB), Cint, (Ptr{Cvoid}, Ref{CsiFields}, Cdouble, Int32, Ptr{Cvoid}),
                                   h.ptr, csi_fields(model), Δt, model.clock.iteration == 0, stream()), h.ptr)

end # module
