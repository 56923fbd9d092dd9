/*
 * climaseaice_b200.h -- C ABI of libclimaseaice_b200.so
 *
 * A B200-native (sm_100a) drop-in for ONE hot path of CliMA/ClimaSeaIce.jl v0.5.8: the
 * split-explicit EVP momentum substep loop plus the h / aice advection update.  The reference
 * has no FFI: its seam is Julia multiple dispatch.  Each entry point below replaces the method
 * named beside it (paths relative to the reference tree); a Julia host overrides those methods
 * for a `B200()` architecture tag and `ccall`s this library (INTEGRATION.md shows the shim).
 *
 * Data contract
 *   - Every field is an Oceananigans `Field.data` parent: dense, column-major, i fastest,
 *     Float64, extents (Nx + 2Hx [+1 if Face on a Bounded x]) x (Ny + 2Hy [+1 ...]) x 1.
 *     Element (i, j) (1-based Julia index) lives at ptr[(i-1+off_x) + (j-1+off_y)*nx_tot].
 *   - The caller owns every field buffer; the library updates them in place and owns only
 *     scratch inside the handle (checkpointing / output on the Julia side keep working).
 *   - Device entry points take DEVICE pointers and are asynchronous on the given CUDA stream.
 *     `*_host` entry points take HOST pointers and include the host<->device copies.
 *   - All functions return 0 on success, <0 for argument/shape errors, >0 for CUDA/NCCL errors;
 *     `csi_last_error` returns the message.  Nothing throws or exits across the ABI.  There is no
 *     CPU fallback: without a CUDA device every compute entry point fails with CSI_ERR_NO_DEVICE.
 */
#ifndef CLIMASEAICE_B200_H
#define CLIMASEAICE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSI_ABI_VERSION 1

typedef struct csi_handle csi_handle;
typedef void *csi_stream; /* cudaStream_t */

typedef struct {
    double *ptr;            /* parent array base (device or host, see entry point) */
    int32_t nx_tot, ny_tot; /* parent extents */
    int32_t off_x, off_y;   /* halo offsets (Hx, Hy of the grid the field lives on) */
} csi_array;

/* CSI_FOLDED (topo_y only): a wall in the south, a fold in the north -- the Zipper boundary condition of Oceananigans' TripolarGrid
 * (src/sea_ice_model.jl:56-64 gives u and v the sign -1).  The fold itself is passed as a copy list, see csi_config.fold_target */
enum { CSI_PERIODIC = 0, CSI_BOUNDED = 1, CSI_FOLDED = 2 };
enum { CSI_STRESS_NONE = 0, CSI_STRESS_CONST = 1, CSI_STRESS_FIELD = 2, CSI_STRESS_SEMI_IMPLICIT = 3 };
enum { CSI_REPLACEMENT_PRESSURE = 0, CSI_ICE_STRENGTH = 1 };
/* SPHERICAL: Oceananigans' HydrostaticSphericalCoriolis (EnstrophyConserving scheme) on a LatitudeLongitudeGrid;
 * f at (Face, Face) is passed per row in csi_config.coriolis_f_ff */
enum { CSI_CORIOLIS_NONE = 0, CSI_CORIOLIS_FPLANE = 1, CSI_CORIOLIS_SPHERICAL = 2 };
enum { CSI_BC_DEFAULT = 0, CSI_BC_VALUE = 1 };
/* free_drift of SeaIceMomentumEquation (src/SeaIceDynamics/stress_balance_free_drift.jl:61-129): nothing, a (u=, v=)
 * pair of arrays, or StressBalanceFreeDrift (closed form from the model's own stresses, exactly one of which must be
 * a SemiImplicitStress) */
enum { CSI_FD_NONE = 0, CSI_FD_FIELDS = 1, CSI_FD_STRESS_BALANCE = 2 };
enum { CSI_RK3 = 0, CSI_FE = 1 };
enum { CSI_SOLVER_AUTO = 0, CSI_SOLVER_UNFUSED = 1, CSI_SOLVER_FUSED = 2 };

enum {
    CSI_OK = 0,
    CSI_ERR_ARG = -1,         /* null pointer, bad enum, non-positive size */
    CSI_ERR_SHAPE = -2,       /* a csi_array does not match the grid */
    CSI_ERR_UNSUPPORTED = -3, /* valid in the reference, not implemented here yet */
    CSI_ERR_NO_DEVICE = -4,   /* no CUDA device / wrong architecture */
    CSI_ERR_NCCL_MISSING = -5
};

/* Mirrors the keyword constructors of the reference:
 *   RectilinearGrid(size, x, y, halo, topology)                       (Oceananigans)
 *   ElastoViscoPlasticRheology(...)      src/Rheologies/elasto_visco_plastic_rheology.jl:119-137
 *   SplitExplicitSolver(grid; substeps)  src/SeaIceDynamics/split_explicit_momentum_equations.jl:18-46
 *   SeaIceMomentumEquation(grid; ...)    src/SeaIceDynamics/sea_ice_momentum_equations.jl:67-94
 *   SemiImplicitStress(; ue, ve, rho_e, Cd)  src/SeaIceDynamics/sea_ice_external_stress.jl:84-130
 *   SeaIceModel(grid; advection, timestepper, boundary_conditions)  src/sea_ice_model.jl:140-158 */
typedef struct {
    int32_t abi_version; /* CSI_ABI_VERSION */
    int32_t device;      /* CUDA device ordinal */
    /* grid (the local block when partitioned) */
    int32_t Nx, Ny, Hx, Hy;
    int32_t topo_x, topo_y;
    double dx, dy;
    const uint8_t *immersed_mask; /* optional, centres, (Nx+2Hx) x (Ny+2Hy), 1 = immersed; NULL = none */
    /* ElastoViscoPlasticRheology */
    double ice_compressive_strength; /* P*   = 27500 */
    double ice_compaction_hardening; /* C    = 20    */
    double yield_curve_eccentricity; /* e    = 2     */
    double minimum_plastic_stress;   /* Dmin = 2e-9  */
    double min_relaxation_parameter; /* 50  */
    double max_relaxation_parameter; /* 300 */
    double relaxation_strength;      /* pi^2 */
    int32_t pressure_formulation;
    /* SplitExplicitSolver */
    int32_t substeps;
    /* SeaIceMomentumEquation */
    double minimum_mass, minimum_concentration, ice_density;
    int32_t coriolis_kind;
    int32_t top_stress_kind;    /* NONE / CONST / FIELD (tau arrays in csi_fields.top_x/top_y) */
    double coriolis_f;
    double top_tau_x, top_tau_y;
    int32_t bottom_stress_kind; /* NONE / SEMI_IMPLICIT (ue/ve arrays, or constants if ptr NULL) */
    int32_t u_south_north_bc;   /* tangential BC of u on Bounded y: DEFAULT (no-flux) or VALUE */
    double rho_e, Cd, ue_const, ve_const;
    double u_south_north_value;
    int32_t v_west_east_bc;
    int32_t advection_order;    /* 0 none, 1 upwind, 3/5/7 WENO(order) */
    double v_west_east_value;
    int32_t timestepper;        /* CSI_RK3 (":SplitRungeKutta3", default) or CSI_FE */
    int32_t solver_impl;        /* CSI_SOLVER_* : kernel formulation used by csi_evp_substeps */
    /* partition (rank-local block; halos of connected sides are exchanged): slabs along y, or Rx x Ry blocks */
    int32_t rank, nranks;
    int32_t exchange_every;     /* K: substeps between halo exchanges (needs Hy >= 2K+3) */
    int32_t partition_x;        /* Rx of a 2-D partition Rx x (nranks / Rx), rank = ry * Rx + rx; 0 or 1 = slabs along y
                                   (needs Hx >= 2K+3; Face fields then carry no extra column on a Bounded x axis) */
    /* ImmersedBoundaryCondition of examples/ice_advected_on_coastline.jl:91-98: discrete-form flux -C*u on the
     * south/north immersed faces of u and -C*v on the west/east ones of v (isd.jl:57-123); 0 = none */
    double immersed_drag_u, immersed_drag_v;
    /* LatitudeLongitudeGrid (or any grid whose metrics depend on j only): metric_kind = CSI_METRIC_J and
     * metrics[k] -> host array of Ny + 2*Hy + 1 doubles, the value at index j stored at [j - 1 + Hy], in the
     * order dx{cc,fc,cf,ff}, dy{cc,fc,cf,ff}, Az{cc,fc,cf,ff} (Oceananigans' Delta-x/Delta-y/Az at the four
     * horizontal locations).  The arrays are copied at csi_create.  CSI_METRIC_REGULAR uses dx, dy above.
     * CSI_METRIC_IJ (orthogonal curvilinear grids: OrthogonalSphericalShellGrid, rotated or stretched meshes): the same
     * twelve metrics as two-dimensional host arrays of (Ny + 2*Hy + 1) rows x (Nx + 2*Hx + 1) columns, i fastest, the value
     * at (i, j) stored at [(j - 1 + Hy) * (Nx + 2*Hx + 1) + (i - 1 + Hx)]; one rank or y-slabs (a partition along x: general kernels refuse it). */
    int32_t metric_kind;
    int32_t serial_exchange;    /* slabs + fused solver: 0 = the halo exchange between blocks of K substeps runs on its own stream
                                   while the next substep's interior tiles compute (boundary tiles wait for it); 1 = on the compute stream */
    const double *metrics[12];
    /* free drift velocity of marginal ice (mass or concentration under the thresholds but above eps): CSI_FD_* */
    int32_t free_drift_kind;
    int32_t reserved3_;
    /* SemiImplicitStress as the TOP stress (top_stress_kind = CSI_STRESS_SEMI_IMPLICIT): rho_e, Cd of the atmosphere;
     * its u_e, v_e are csi_fields.top_x/top_y, or the constants top_tau_x/top_tau_y when those are NULL.  Likewise the
     * BOTTOM stress may be CSI_STRESS_CONST (ue_const, ve_const hold tau) or CSI_STRESS_FIELD (ue, ve hold tau). */
    double top_rho_e, top_Cd;
    /* CSI_CORIOLIS_SPHERICAL: f^ffa = 2 Omega sin(phi^f) per row, Ny + 2*Hy + 1 doubles, row j at [j - 1 + Hy] (host
     * array, copied at csi_create); evaluated by the host with Oceananigans' own f^ffa */
    const double *coriolis_f_ff;
    /* topo_y = CSI_FOLDED: what fill_halo_regions! does at the north boundary of the host's grid, as one copy list per location
     * (c,c), (f,c), (c,f), (f,f): after the local fills of the other sides, parent[target[k]] = sign * parent[source[k]], with
     * linear parent indices (i fastest) into an array of that location.  The host derives the lists from its own grid -- a
     * Julia host by filling a scratch field with its linear indices and calling fill_halo_regions! once per location
     * (julia/ClimaSeaIceB200.jl: fold_maps) -- so no index convention of the fold is built into the library.  Targets may be
     * halo cells of any side and interior cells (the duplicated row of a centre-pivot fold).  sign: fold_sign_velocity for u, v
     * (-1 in the reference), fold_sign_external for the external stress / velocity arrays top_x, top_y, ue, ve, +1 for every
     * other field.  Checked at csi_create: indices inside the parent, every target once, nothing both read and written (the
     * copies of one list run concurrently).  The lists should cover the x halo columns of the rows they write (the fold is
     * applied after the periodic fill in x).  Host arrays, copied at csi_create.  One rank or y-slabs (the last slab holds the
     * fold; the other ranks ignore these members); the fused solver takes a fold on grids with two-dimensional metrics. */
    const int32_t *fold_target[4];
    const int32_t *fold_source[4];
    int32_t fold_count[4];
    double fold_sign_velocity, fold_sign_external;
} csi_config;

/* The arrays the hot path touches.  Unused ones may have ptr == NULL. */
typedef struct {
    csi_array u, v;                 /* model.velocities      (f,c) (c,f) */
    csi_array h, a;                 /* ice_thickness, ice_concentration (c,c) */
    csi_array s11, s22, s12;        /* auxiliaries.fields sigma11, sigma22 (c,c), sigma12 (f,f) */
    csi_array zeta_f, zeta_c, delta, alpha, un, vn, P; /* remaining EVP auxiliaries (evp.jl:147-169) */
    csi_array top_x, top_y;         /* top stress tau_x (f,c), tau_y (c,f) when FIELD */
    csi_array ue, ve;               /* SemiImplicitStress external velocities (f,c) (c,f) */
    csi_array Gh, Ga;               /* timestepper.G^n.h, .aice */
    csi_array hm, am, um, vm;       /* timestepper.Psi^- (RK3) */
    csi_array hs, Ghs, hsm;         /* optional snow_thickness (c,c), G^n.hs, Psi^-.hs: advected with h and aice
                                       (src/tracer_tendency_kernel_functions.jl:47-52, src/sea_ice_fe_step.jl:84-94) */
    csi_array fd_u, fd_v;           /* free-drift velocities (f,c) (c,f) when free_drift_kind = CSI_FD_FIELDS */
} csi_fields;

enum { CSI_METRIC_REGULAR = 0, CSI_METRIC_J = 1, CSI_METRIC_IJ = 2 };

int csi_version(void);
const char *csi_last_error(const csi_handle *h); /* h may be NULL: last error of csi_create */

/* SeaIceModel(...) construction point: src/sea_ice_model.jl:140-297 */
int csi_create(const csi_config *cfg, csi_handle **out);
int csi_destroy(csi_handle *h);

/* time_step_momentum!(model, ::SplitExplicitMomentumEquation, dt)
 *   src/SeaIceDynamics/split_explicit_momentum_equations.jl:103-195
 * = reset_velocities! (:87-93) + initialize_rheology! (evp.jl:192-216) + update_external_stress!
 *   (sea_ice_external_stress.jl:72-78) + nsubsteps x {compute_stresses! (evp.jl:222-354);
 *   alternating _u/_v_velocity_step! (:197-264) with local halo fills} + finalize_rheology!
 *   (evp.jl:275-280). */
int csi_evp_substeps(csi_handle *h, const csi_fields *f, double dt_stage, int32_t nsubsteps, csi_stream stream);

/* compute_tracer_tendencies!(model)  src/tracer_tendency_kernel_functions.jl:9-45 */
int csi_compute_tracer_tendencies(csi_handle *h, const csi_fields *f, csi_stream stream);
/* dynamic_time_step!(model, dt)      src/sea_ice_rk_substep.jl:134-152, src/sea_ice_fe_step.jl:36-82 */
int csi_dynamic_time_step(csi_handle *h, const csi_fields *f, double dt_stage, csi_stream stream);
/* cache_current_fields!(model)       src/sea_ice_rk_substep.jl:29-42 */
int csi_cache_current_fields(csi_handle *h, const csi_fields *f, csi_stream stream);
/* update_state!(model)               src/sea_ice_model.jl:379-394 (mask + halo fill of h, aice, u, v) */
int csi_update_state(csi_handle *h, const csi_fields *f, csi_stream stream);
/* fill_halo_regions!(field): loc_x/loc_y 0 = Center, 1 = Face; which 0 = default BCs, 1 = u, 2 = v */
int csi_fill_halos(csi_handle *h, const csi_array *a, int32_t loc_x, int32_t loc_y, int32_t which, csi_stream stream);
/* time_step!(model, dt): src/sea_ice_fe_step.jl:13-34 / src/sea_ice_rk_substep.jl:81-94 driven by
 * Oceananigans' SplitRungeKuttaTimeStepper; first != 0 mirrors `clock.iteration == 0`. */
int csi_time_step(csi_handle *h, const csi_fields *f, double dt, int32_t first, csi_stream stream);
/* cell_advection_timescale(model)    src/ClimaSeaIce.jl:66-69; synchronises the stream */
int csi_cell_advection_timescale(csi_handle *h, const csi_fields *f, double *out_host, csi_stream stream);
/* diagnostics (deterministic two-pass reductions; synchronises the stream):
 * out[0] = sum h*Az, out[1] = sum aice*Az, out[2] = sum h*aice*Az, out[3] = max|u|, out[4] = max|v| */
int csi_diagnostics(csi_handle *h, const csi_fields *f, double *out_host5, csi_stream stream);

/* Same as csi_time_step but with HOST buffers: uploads every non-NULL array, runs `nsteps` model
 * steps, downloads u, v, h, a, s11, s22, s12, alpha.  Blocking.  Used for the end-to-end metric. */
int csi_time_step_host(csi_handle *h, const csi_fields *host_fields, double dt, int32_t nsteps, int32_t first);
/* Same for the momentum solve alone (time_step_momentum! with host buffers). */
int csi_evp_substeps_host(csi_handle *h, const csi_fields *host_fields, double dt_stage, int32_t nsubsteps);

/* Bytes the last *_host call copied host->device and device->host (inputs only go up, results only come down). */
int csi_last_transfer_bytes(const csi_handle *h, uint64_t *h2d, uint64_t *d2h);

/* Multi-GPU (one process per GPU, slabs along y).  The 128-byte NCCL unique id is produced on
 * rank 0 and distributed by the host application (torch.distributed / MPI.jl broadcast). */
int csi_nccl_unique_id(uint8_t out128[128]);
int csi_comm_init(csi_handle *h, const uint8_t id128[128], int32_t rank, int32_t nranks);
/* distributed fill_halo_regions! of one field across slab neighbours (Oceananigans
 * DistributedComputations, used at src/sea_ice_model.jl:381-384, evp.jl:275-280) */
int csi_exchange_halos(csi_handle *h, const csi_array *arrays, int32_t narrays, int32_t width, csi_stream stream);
/* fill_halo_regions!(fields; async = true) and synchronize_communication!(field)
 * (src/Rheologies/elasto_visco_plastic_rheology.jl:275-280 and :204-206): csi_exchange_halos_async orders the exchange behind
 * the work already queued on `stream` and runs it on the handle's communication stream -- `stream` itself does not wait;
 * csi_wait_halos makes `stream` wait for the exchange in flight (stream-ordered, the host does not block).  One exchange
 * may be in flight per handle.  csi_time_step uses this pair itself for the stress halos of finalize_rheology!: the exchange
 * overlaps the stage's h / aice update, thermodynamics and update_state!, and is synchronised where the reference does it,
 * at the next initialize_rheology! (and before csi_time_step returns control of the arrays to the caller's stream). */
int csi_exchange_halos_async(csi_handle *h, const csi_array *arrays, int32_t narrays, int32_t width, csi_stream stream);
int csi_wait_halos(csi_handle *h, csi_stream stream);

/* ---- slab thermodynamics -----------------------------------------------
 * thermodynamic_time_step!(model, ice_thermodynamics, snow_thermodynamics, dt)
 *     src/SeaIceThermodynamics/thermodynamic_time_step.jl:6-59: one pointwise kernel per call, the bare-ice
 *     `_ice_thermodynamic_time_step!` (:76-118) or the layered snow + ice `_layered_thermodynamic_time_step!`
 *     (:132-291), with thermodynamic_tendency / ice_melt_freeze_tendency (slab_thermodynamics_tendencies.jl:30-135),
 *     ice_volume_update, snow_ice_formation, ProportionalEvolution (:297-369), ConductiveFlux /
 *     IceSnowConductiveFlux (slab_heat_and_tracer_fluxes.jl), PhaseTransitions / LinearLiquidus
 *     (SeaIceThermodynamics.jl:22-167) and the heat boundary conditions (HeatBoundaryConditions/ *.jl).
 * Julia closures cannot cross a C ABI, so the external top flux is a menu: up to two summed terms, each a constant,
 * an array, RadiativeEmission, a linear bulk flux, or the conductive flux (the model's default for PrescribedTemperature,
 * src/sea_ice_model.jl:244-252).  The surface-temperature solve restates RootSolvers.jl's SecantMethod. */
enum { CSI_TOP_MELTING_CONSTRAINED_FLUX_BALANCE = 0, CSI_TOP_PRESCRIBED_TEMPERATURE = 1 };
enum { CSI_BOTTOM_ICE_WATER_EQUILIBRIUM = 0, CSI_BOTTOM_PRESCRIBED_TEMPERATURE = 1 };
/* CSI_FLUX_LINEAR: coefficient * (T_top - temperature), optionally times the concentration -- the bulk sensible-heat
 * FluxFunction the reference's own tests use (test/test_energy_conservation.jl:8-13,103-110) */
enum { CSI_FLUX_CONST = 0, CSI_FLUX_ARRAY = 1, CSI_FLUX_RADIATIVE_EMISSION = 2, CSI_FLUX_CONDUCTIVE = 3, CSI_FLUX_LINEAR = 4 };

typedef struct {
    /* PhaseTransitions(density=917, heat_capacity=2000, liquid_density=999.8, liquid_heat_capacity=4186,
     * reference_latent_heat=334e3, reference_temperature=0, liquidus=LinearLiquidus(slope=0.054, T0=0)) */
    double density, heat_capacity, liquid_density, liquid_heat_capacity, reference_latent_heat, reference_temperature;
    double liquidus_freshwater_melting_temperature, liquidus_slope;
    /* SlabThermodynamics(top_heat_boundary_condition, bottom_heat_boundary_condition, internal_heat_flux) of the ice
     * slab and, when layered != 0, of the snow slab (snow_slab_thermodynamics: conductivity 0.31) */
    int32_t top_heat_bc, snow_top_heat_bc; /* CSI_TOP_* */
    int32_t bottom_heat_bc;                /* CSI_BOTTOM_* */
    int32_t layered;                       /* 0: bare ice, 1: snow + ice */
    double ice_conductivity, snow_conductivity;
    double bottom_salinity, bottom_temperature; /* constants, used when csi_thermo_fields.Sb / .Tb are NULL */
    /* model.external_heat_fluxes: top = term[0] (+ term[1]); bottom = Qbot array or the constant */
    int32_t n_top_terms, top_term_kind[2], reserved_;
    double top_flux_const, emissivity, stefan_boltzmann_constant, emission_reference_temperature;
    double bottom_flux_const;
    /* SeaIceModel keywords, each used when its array is NULL: snowfall = 0, snow_density = 330,
     * ice_consolidation_thickness = 0.05, ice_salinity = 0 (src/sea_ice_model.jl:66-83) */
    double snowfall, snow_density, ice_consolidation_thickness, ice_salinity;
    /* RootSolvers.find_zero defaults: SolutionTolerance(1e-3), maxiters = 10 000 */
    double secant_tolerance;
    int32_t secant_maxiters, reserved2_;
    double linear_coefficient, linear_temperature; /* CSI_FLUX_LINEAR */
    int32_t linear_times_concentration, reserved3_;
} csi_thermo_config;

typedef struct {
    csi_array h, a, hs;           /* ice_thickness, ice_concentration, snow_thickness (layered) -- all (c,c) */
    csi_array Tu, Tus;            /* top_surface_temperature of the ice slab / of the snow slab */
    csi_array S, hc;              /* ice salinity, ice_consolidation_thickness */
    csi_array Qtop, Qbot;         /* external heat flux arrays */
    csi_array Sb, Tb;             /* IceWaterThermalEquilibrium salinity / PrescribedTemperature at the bottom */
    csi_array snowfall, rho_s;    /* snowfall (kg m^-2 s^-1), snow_density */
    csi_array mf_ice, mf_snow, mf_snowfall; /* mass_fluxes.thermodynamics.ice / .snow, .intercepted_snowfall (outputs) */
} csi_thermo_fields;

int csi_thermodynamic_time_step(csi_handle *h, const csi_thermo_config *cfg, const csi_thermo_fields *f, double dt, csi_stream stream);
/* Make the thermodynamics part of the model: csi_time_step then runs thermodynamic_time_step! after
 * dynamic_time_step! in every stage (src/sea_ice_fe_step.jl:27-30, src/sea_ice_rk_substep.jl:89-91) and
 * csi_update_state masks the three mass-flux diagnostics on immersed cells (src/sea_ice_model.jl:386-389).
 * h, a (and hs) must be the arrays of the csi_fields passed to csi_time_step.  cfg == NULL detaches. */
int csi_attach_thermodynamics(csi_handle *h, const csi_thermo_config *cfg, const csi_thermo_fields *f);

/* Instrumentation: kernels launched by this handle so far; elapsed ms of the last device call
 * measured with CUDA events on its stream. */
int64_t csi_launch_count(const csi_handle *h);
/* Diagnostics of the fused solver's last momentum solve (synchronises the device): out[0] != 0 if an input failed the
 * range validation (whole stage on the IEEE pass), out[1] = tile passes redone with the IEEE operators, out[2] = tiles
 * per substep.  All zero when the general kernels ran. */
int csi_fused_stats(const csi_handle *h, int64_t out3[3]);
double csi_last_elapsed_ms(const csi_handle *h);

/* Launches the dominant kernel of csi_evp_substeps (the fused substep kernel, or the stress kernel of
 * the unfused formulation) `reps` times between two CUDA events on `stream` and returns the average
 * duration per launch, its name and its algorithmic bytes per cell.  Advances the state
 * by `reps` stress updates; for bench.py's roofline only.  Synchronises the stream. */
int csi_time_dominant_kernel(csi_handle *h, const csi_fields *f, double dt_stage, int32_t reps, double *out_ms_per_launch,
                             char *name64, int32_t *bytes_per_cell, csi_stream stream);

/* Device self test: compares the kernels' branch-free division / reciprocal / square root / constant
 * quotient with the hardware IEEE operators on `samples` random operand sets whose exponents spread
 * over +-2^exponent_span.  out5 = {rcp, div, sqrt, constant-quotient mismatches, samples outside the
 * fast windows}.  A correct build returns zeros in out5[0..3]. */
int csi_selftest_math(int64_t samples, uint64_t seed, int32_t exponent_span, uint64_t *out5);

/* Device microbenchmark behind bench.py's FP64 roofline: thread-level FP64 FMA instructions per second of `device`
 * (eight independent chains per thread, full occupancy), as a burst (best short launch) and sustained over the second half
 * of `seconds` of back-to-back launches (power-capped clock).  Instrumentation only. */
int csi_measure_fp64_rate(int32_t device, double seconds, double *burst_fma_per_s, double *sustained_fma_per_s);

/* Host-side helpers exported for CPU tests (no GPU needed). */
double csi_host_exp(double x);                       /* the correctly rounded exp used for ice_strength */
double csi_host_div_by_const(double x, double c);    /* the Markstein constant-division used in the kernels */
int csi_host_halo_width(int32_t substeps_between_exchanges); /* 2K+3, se.jl:55-56 */

#ifdef __cplusplus
}
#endif
#endif
