/*
 * csi_oracle.h -- CPU ORACLE for the ClimaSeaIce.jl EVP-substep + h/aice advection hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product library
 * (libclimaseaice_b200.so) never includes, links or calls anything under oracle/.
 *
 * It is a plain-C restatement of the reference's KernelAbstractions-CPU path, one function per
 * reference function, recomputing every quantity at every neighbour exactly as the reference's
 * inlined operators do.  Every function cites the reference file:line it follows (paths relative
 * to /root/reference).  Compile with -ffp-contract=off: Julia does not contract a*b+c into FMA.
 *
 * PARITY UNPINNED: the reference is Julia + Oceananigans.jl; neither is runnable in this image and
 * the reference ships no golden vectors for this path.  The oracle is pinned
 * only by the reference's own *property* tests restated in tests/ (SBP adjoint identity of
 * test/test_rheology_energy_budget.jl, drag bound of test/test_time_stepping.jl:56-80, decomposition
 * invariance of test/distributed_tests_utils.jl:40-88).  Oceananigans primitives are restated from
 * SURVEY.md Appendix A ("[OCN-recall]") and are isolated in small functions below.
 *
 * Index convention: (i, j) are the reference's 1-based Julia indices; face i lies between
 * centres i-1 and i.
 */
#ifndef CSI_ORACLE_H
#define CSI_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Oceananigans Field.data: OffsetArray over a dense column-major parent. */
typedef struct {
    double *p;      /* parent array, i fastest */
    int32_t sx, sy; /* parent extents */
    int32_t ox, oy; /* halo offsets: element (i,j) lives at p[(i-1+ox) + (j-1+oy)*sx] */
} csio_field;

/* CSIO_FOLDED (y axis only): a wall in the south, a fold (Oceananigans' Zipper boundary condition of the tripolar grid) in the north */
enum { CSIO_PERIODIC = 0, CSIO_BOUNDED = 1, CSIO_FOLDED = 2 };
enum { CSIO_REGULAR = 0, CSIO_JMETRIC = 1, CSIO_IJMETRIC = 2 };

typedef struct {
    int32_t Nx, Ny, Hx, Hy;
    int32_t topo_x, topo_y;  /* CSIO_PERIODIC / CSIO_BOUNDED */
    int32_t metric_kind;     /* CSIO_REGULAR: dx,dy scalars; CSIO_JMETRIC: j-indexed arrays (lat-lon); CSIO_IJMETRIC: 2-D arrays */
    int32_t metW;            /* CSIO_IJMETRIC: columns per row of the metric arrays (Nx + 2 Hx + 1) */
    double dx, dy;
    /* j-indexed metrics, entry for index j at arr[j-1+Hy], length Ny+2Hy+1 (CSIO_JMETRIC only) */
    const double *dxcc, *dxfc, *dxcf, *dxff;
    const double *dycc, *dyfc, *dycf, *dyff;
    const double *azcc, *azfc, *azcf, *azff;
    /* optional immersed mask at cell centres, shape (Nx+2Hx) x (Ny+2Hy), 1 = immersed (inactive) */
    const uint8_t *mask;
    /* CSIO_FOLDED: the north fold as a copy list per location (c,c), (f,c), (c,f), (f,f): after the local fills of the other
     * sides, parent[target[k]] = sign * parent[source[k]] (linear parent indices, i fastest).  The list is whatever the host's
     * fill_halo_regions! does on that grid (src/sea_ice_model.jl:56-64 only sets the sign of u and v to -1); sign: u, v take
     * fold_sign_velocity, external stress / velocity arrays fold_sign_external, everything else +1 */
    const int32_t *fold_target[4], *fold_source[4];
    int32_t fold_count[4];
    double fold_sign_velocity, fold_sign_external;
} csio_grid;

/* external stress kinds: sea_ice_external_stress.jl:8-27,176-202 */
enum { CSIO_STRESS_NONE = 0, CSIO_STRESS_CONST = 1, CSIO_STRESS_FIELD = 2, CSIO_STRESS_SEMI_IMPLICIT = 3 };
enum { CSIO_REPLACEMENT_PRESSURE = 0, CSIO_ICE_STRENGTH = 1 };
enum { CSIO_CORIOLIS_NONE = 0, CSIO_CORIOLIS_FPLANE = 1, CSIO_CORIOLIS_SPHERICAL = 2 };
enum { CSIO_BC_DEFAULT = 0, CSIO_BC_VALUE = 1 };
enum { CSIO_RK3 = 0, CSIO_FE = 1 };
enum { CSIO_FD_NONE = 0, CSIO_FD_FIELDS = 1, CSIO_FD_STRESS_BALANCE = 2 };

typedef struct {
    /* ElastoViscoPlasticRheology: elasto_visco_plastic_rheology.jl:14-25,119-137 */
    double Pstar, C, e, Dmin, alpha_min, alpha_max, c_alpha;
    int32_t pressure_formulation;
    int32_t substeps;               /* SplitExplicitSolver.substeps */
    /* SeaIceMomentumEquation: sea_ice_momentum_equations.jl:67-94 */
    double min_mass, min_conc, rho_ice;
    int32_t coriolis_kind, pad0_;
    double f;
    /* top (atmosphere) stress: NONE / CONST / FIELD */
    int32_t top_kind, pad1_;
    double top_tx, top_ty;
    csio_field top_x, top_y;
    /* bottom (ocean) stress: NONE / SEMI_IMPLICIT (ue,ve arrays, or constants when ue.p == NULL) */
    int32_t bot_kind, pad2_;
    double rho_e, Cd, ue_c, ve_c;
    csio_field ue, ve;
    /* tangential velocity BCs on Bounded axes: DEFAULT (no-flux) or VALUE */
    int32_t u_sn_bc, v_we_bc;
    double u_sn_val, v_we_val;
    /* advection: 0 = nothing, 1 = first-order upwind, 3/5/7 = WENO(order) */
    int32_t advection_order;
    int32_t timestepper;            /* CSIO_RK3 / CSIO_FE */
    /* immersed boundary condition of examples/ice_advected_on_coastline.jl:91-98: discrete-form
     * FluxBoundaryCondition -C*u on the south/north immersed faces of u, -C*v on west/east of v; 0 = none */
    double imm_drag_u, imm_drag_v;
    /* free drift for marginal ice (stress_balance_free_drift.jl:61-129): NONE -> 0, FIELDS -> fd_u/fd_v arrays,
     * STRESS_BALANCE -> closed form from the model's own stresses (exactly one side SEMI_IMPLICIT) */
    int32_t free_drift_kind, pad3_;
    csio_field fd_u, fd_v;
    /* top SemiImplicitStress (top_kind == SEMI_IMPLICIT): u_e, v_e = top_x/top_y arrays or top_tx/top_ty constants.
     * Likewise bot_kind may be CONST / FIELD, the stress then being ue/ve (arrays) or ue_c/ve_c (constants). */
    double top_rho, top_Cd;
    /* HydrostaticSphericalCoriolis (coriolis_kind == SPHERICAL): f at (Face, Face) = 2 Omega sin(phi_f), j-indexed like the
     * grid metrics (entry for row j at [j-1+Hy], length Ny+2Hy+1) */
    const double *f_ff;
} csio_params;

typedef struct {
    csio_field u, v, h, a;                               /* prognostic: velocities, thickness, concentration */
    csio_field s11, s22, s12, zf, zc, delta, alpha, un, vn, P;  /* EVP Auxiliaries (evp.jl:147-169) */
    csio_field Gh, Ga;                                   /* timestepper.G^n */
    csio_field hm, am, um, vm;                           /* timestepper.Psi^- (RK3 only) */
    csio_field hs, Ghs, hsm;                             /* optional snow thickness, its tendency, Psi^-.hs (p == NULL: no snow) */
} csio_state;

/* ---- entry points (all return 0 on success) ---- */
double csio_exp(double x);   /* correctly rounded exp via binary128 */

int csio_fill_halo(const csio_grid *g, const csio_params *p, csio_field *f, int lx, int ly, int which);
int csio_initialize_rheology(const csio_grid *g, const csio_params *p, csio_state *s);
int csio_compute_stresses(const csio_grid *g, const csio_params *p, csio_state *s, double dt);
int csio_u_velocity_step(const csio_grid *g, const csio_params *p, csio_state *s, double dt);
int csio_v_velocity_step(const csio_grid *g, const csio_params *p, csio_state *s, double dt);
int csio_time_step_momentum(const csio_grid *g, const csio_params *p, csio_state *s, double dt, int nsub);
int csio_compute_tracer_tendencies(const csio_grid *g, const csio_params *p, csio_state *s);
int csio_dynamic_time_step(const csio_grid *g, const csio_params *p, csio_state *s, double dt);
int csio_update_state(const csio_grid *g, const csio_params *p, csio_state *s);
int csio_time_step(const csio_grid *g, const csio_params *p, csio_state *s, double dt, int first);
double csio_cell_advection_timescale(const csio_grid *g, const csio_state *s);
/* out[0]=W_new, out[1]=W_old, out[2]=D of test/test_rheology_energy_budget.jl:50-91 */
int csio_stress_power_budget(const csio_grid *g, const csio_state *s, double *out);
/* WENO face reconstruction exposed for unit tests: bias 0 = left (u>0), 1 = right */
double csio_reconstruct_x(const csio_grid *g, int order, int bias, const csio_field *c, int i, int j);
double csio_reconstruct_y(const csio_grid *g, int order, int bias, const csio_field *c, int i, int j);
int csio_set_threads(int n);

/* ---- slab thermodynamics (csi_oracle_thermo.c) ---- */
enum { CSIO_TOP_MELTING_CONSTRAINED_FLUX_BALANCE = 0, CSIO_TOP_PRESCRIBED_TEMPERATURE = 1 };
enum { CSIO_BOTTOM_ICE_WATER_EQUILIBRIUM = 0, CSIO_BOTTOM_PRESCRIBED_TEMPERATURE = 1 };
enum { CSIO_FLUX_CONST = 0, CSIO_FLUX_ARRAY = 1, CSIO_FLUX_RADIATIVE_EMISSION = 2, CSIO_FLUX_CONDUCTIVE = 3, CSIO_FLUX_LINEAR = 4 };

typedef struct {
    /* PhaseTransitions + LinearLiquidus: SeaIceThermodynamics.jl:22-127 */
    double density, heat_capacity, liquid_density, liquid_heat_capacity, reference_latent_heat, reference_temperature;
    double liquidus_T0, liquidus_slope;
    /* SlabThermodynamics: slab_sea_ice_thermodynamics.jl:18-108 */
    int32_t top_bc, snow_top_bc;          /* CSIO_TOP_* of the ice slab / of the snow slab */
    int32_t bottom_bc;                    /* CSIO_BOTTOM_* */
    int32_t layered;                      /* 1: snow + ice kernel, 0: bare ice */
    double ice_conductivity, snow_conductivity;
    double bottom_salinity, bottom_temperature;  /* constants, used when Sb / Tb arrays are NULL */
    /* external fluxes: top = term[0] (+ term[1]), bottom = array or constant */
    int32_t n_top_terms, top_term_kind[2], pad_;
    double top_flux_const, emissivity, stefan_boltzmann, emission_reference_temperature;
    double bottom_flux_const;
    /* model-level constants, each used when its array is NULL: sea_ice_model.jl:66-83 */
    double snowfall, snow_density, consolidation_thickness, ice_salinity;
    /* RootSolvers SecantMethod defaults [RS-recall] */
    double secant_tol;
    int32_t secant_maxiters, pad2_;
    /* CSIO_FLUX_LINEAR: coefficient * (T - temperature) [* aice], the sensible-heat FluxFunction of
     * test/test_energy_conservation.jl:8-13,103-110 */
    double linear_coefficient, linear_temperature;
    int32_t linear_times_concentration, pad3_;
} csio_thermo_params;

typedef struct {
    csio_field h, a, hs;          /* ice thickness, concentration, snow thickness (layered) */
    csio_field Tu, Tus;           /* top surface temperature of the ice slab / of the snow slab */
    csio_field S, hc;             /* ice salinity, consolidation thickness */
    csio_field Qtop, Qbot;        /* external heat flux arrays */
    csio_field Sb, Tb;            /* bottom salinity / prescribed bottom temperature */
    csio_field snowfall, rho_s;   /* snowfall, snow density */
    csio_field mf_ice, mf_snow, mf_snowfall; /* mass_fluxes.thermodynamics.ice/.snow, .intercepted_snowfall */
} csio_thermo_state;

int csio_thermodynamic_time_step(const csio_grid *g, const csio_thermo_params *p, csio_thermo_state *s, double rho_ice, double dt);
double csio_pow4(double x);

#ifdef __cplusplus
}
#endif
#endif
