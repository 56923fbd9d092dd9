/* c_abi_smoke.c -- a caller of libclimaseaice_b200.so that is not Python: plain C, compiled against the public header,
 * binding the library at run time the way a Julia `ccall` / cgo / JNI host would (dlopen + dlsym).
 *
 *   c_abi_smoke <libclimaseaice_b200.so> --symbols          every symbol include/climaseaice_b200.h declares resolves (no GPU)
 *   c_abi_smoke <libclimaseaice_b200.so> <fixture.bin>      csi_create -> csi_evp_substeps_host -> csi_destroy on host buffers,
 *                                                           results compared bit for bit with the committed fixture
 *                                                           (tests/golden/make_c_abi_fixture.py; needs a B200)
 * Built and run by tests/test_c_abi.py:  gcc -std=c11 -Iinclude tests/c_abi_smoke.c -ldl -o c_abi_smoke */
#include <dlfcn.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "climaseaice_b200.h"

static const char *SYMBOLS[] = {
    "csi_version", "csi_last_error", "csi_create", "csi_destroy", "csi_evp_substeps", "csi_compute_tracer_tendencies",
    "csi_dynamic_time_step", "csi_cache_current_fields", "csi_update_state", "csi_fill_halos", "csi_time_step",
    "csi_cell_advection_timescale", "csi_diagnostics", "csi_time_step_host", "csi_evp_substeps_host", "csi_last_transfer_bytes",
    "csi_nccl_unique_id", "csi_comm_init", "csi_exchange_halos", "csi_exchange_halos_async", "csi_wait_halos", "csi_launch_count",
    "csi_fused_stats", "csi_last_elapsed_ms", "csi_time_dominant_kernel", "csi_thermodynamic_time_step", "csi_attach_thermodynamics",
    "csi_selftest_math", "csi_measure_fp64_rate", "csi_host_exp", "csi_host_div_by_const", "csi_host_halo_width"};

typedef int (*create_fn)(const csi_config *, csi_handle **);
typedef int (*destroy_fn)(csi_handle *);
typedef int (*substeps_host_fn)(csi_handle *, const csi_fields *, double, int32_t);
typedef const char *(*last_error_fn)(const csi_handle *);
typedef int (*version_fn)(void);
typedef int (*stats_fn)(const csi_handle *, int64_t[3]);

static double *read_array(FILE *f, size_t n)
{
    double *a = (double *)malloc(n * sizeof(double));
    if (!a || fread(a, sizeof(double), n, f) != n) {
        fprintf(stderr, "fixture truncated\n");
        exit(2);
    }
    return a;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "usage: %s lib.so --symbols | fixture.bin\n", argv[0]);
        return 2;
    }
    void *lib = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        fprintf(stderr, "dlopen: %s\n", dlerror());
        return 2;
    }
    for (size_t k = 0; k < sizeof SYMBOLS / sizeof *SYMBOLS; k++)
        if (!dlsym(lib, SYMBOLS[k])) {
            fprintf(stderr, "missing symbol %s\n", SYMBOLS[k]);
            return 1;
        }
    version_fn version = (version_fn)dlsym(lib, "csi_version");
    if (version() != CSI_ABI_VERSION) {
        fprintf(stderr, "ABI version %d, header says %d\n", version(), CSI_ABI_VERSION);
        return 1;
    }
    if (strcmp(argv[2], "--symbols") == 0) {
        printf("C_ABI_SYMBOLS_OK %zu symbols, ABI version %d\n", sizeof SYMBOLS / sizeof *SYMBOLS, version());
        return 0;
    }

    FILE *f = fopen(argv[2], "rb");
    if (!f) {
        perror(argv[2]);
        return 2;
    }
    int32_t hdr[4];
    double dt;
    if (fread(hdr, sizeof(int32_t), 4, f) != 4 || fread(&dt, sizeof(double), 1, f) != 1) return 2;
    const int Nx = hdr[0], Ny = hdr[1], H = hdr[2], nsub = hdr[3];
    const int sx = Nx + 2 * H, sy = Ny + 2 * H;
    const size_t n = (size_t)sx * sy;
    double *in[11], *want[5];
    for (int k = 0; k < 11; k++) in[k] = read_array(f, n);
    for (int k = 0; k < 5; k++) want[k] = read_array(f, n);
    fclose(f);

    /* the model of tests/golden/make_c_abi_fixture.py: doubly periodic, dx = dy = 4 km, default EVP rheology, FPlane,
     * wind-stress arrays, SemiImplicitStress ocean drag with velocity arrays, ForwardEuler (no Psi^- copies) */
    csi_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = CSI_ABI_VERSION;
    cfg.Nx = Nx; cfg.Ny = Ny; cfg.Hx = H; cfg.Hy = H;
    cfg.topo_x = CSI_PERIODIC; cfg.topo_y = CSI_PERIODIC;
    cfg.dx = 4000.0; cfg.dy = 4000.0;
    cfg.ice_compressive_strength = 27500.0; cfg.ice_compaction_hardening = 20.0; cfg.yield_curve_eccentricity = 2.0;
    cfg.minimum_plastic_stress = 2e-9; cfg.min_relaxation_parameter = 50.0; cfg.max_relaxation_parameter = 300.0;
    cfg.relaxation_strength = 3.141592653589793 * 3.141592653589793;
    cfg.pressure_formulation = CSI_REPLACEMENT_PRESSURE;
    cfg.substeps = nsub;
    cfg.minimum_mass = 1.0; cfg.minimum_concentration = 1e-3; cfg.ice_density = 900.0;
    cfg.coriolis_kind = CSI_CORIOLIS_FPLANE; cfg.coriolis_f = 1e-4;
    cfg.top_stress_kind = CSI_STRESS_FIELD;
    cfg.bottom_stress_kind = CSI_STRESS_SEMI_IMPLICIT; cfg.rho_e = 1026.0; cfg.Cd = 5.5e-3;
    cfg.advection_order = 7;
    cfg.timestepper = CSI_FE;
    cfg.solver_impl = CSI_SOLVER_AUTO;
    cfg.nranks = 1;
    cfg.top_rho_e = 1.3; cfg.top_Cd = 1.2e-3;

    create_fn create = (create_fn)dlsym(lib, "csi_create");
    destroy_fn destroy = (destroy_fn)dlsym(lib, "csi_destroy");
    substeps_host_fn substeps_host = (substeps_host_fn)dlsym(lib, "csi_evp_substeps_host");
    last_error_fn last_error = (last_error_fn)dlsym(lib, "csi_last_error");
    stats_fn fused_stats = (stats_fn)dlsym(lib, "csi_fused_stats");

    csi_handle *h = NULL;
    int rc = create(&cfg, &h);
    if (rc) {
        fprintf(stderr, "csi_create: %d %s\n", rc, last_error(NULL));
        return rc == CSI_ERR_NO_DEVICE ? 77 : 1;
    }
    csi_fields fl;
    memset(&fl, 0, sizeof fl);
    csi_array *slot[11] = {&fl.u, &fl.v, &fl.h, &fl.a, &fl.s11, &fl.s22, &fl.s12, &fl.top_x, &fl.top_y, &fl.ue, &fl.ve};
    for (int k = 0; k < 11; k++) {
        slot[k]->ptr = in[k];
        slot[k]->nx_tot = sx; slot[k]->ny_tot = sy; slot[k]->off_x = H; slot[k]->off_y = H;
    }
    /* outputs the call writes: the EVP auxiliaries */
    csi_array *aux[7] = {&fl.zeta_f, &fl.zeta_c, &fl.delta, &fl.alpha, &fl.un, &fl.vn, &fl.P};
    for (int k = 0; k < 7; k++) {
        aux[k]->ptr = (double *)calloc(n, sizeof(double));
        aux[k]->nx_tot = sx; aux[k]->ny_tot = sy; aux[k]->off_x = H; aux[k]->off_y = H;
    }
    rc = substeps_host(h, &fl, dt, nsub);
    if (rc) {
        fprintf(stderr, "csi_evp_substeps_host: %d %s\n", rc, last_error(h));
        return 1;
    }
    int64_t st[3];
    fused_stats(h, st);
    double *got[5] = {in[0], in[1], in[4], in[5], in[6]};
    const char *names[5] = {"u", "v", "s11", "s22", "s12"};
    int bad = 0;
    for (int k = 0; k < 5; k++)
        for (int j = H; j < H + Ny; j++)   /* interior */
            if (memcmp(got[k] + (size_t)j * sx + H, want[k] + (size_t)j * sx + H, Nx * sizeof(double)) != 0) {
                fprintf(stderr, "field %s differs from the fixture in row %d\n", names[k], j - H + 1);
                bad = 1;
                break;
            }
    rc = destroy(h);
    if (rc || bad) return 1;
    printf("C_ABI_SMOKE_OK %dx%d, %d substeps, fused tiles per substep %lld, bit-identical to the fixture\n", Nx, Ny, nsub, (long long)st[2]);
    return 0;
}
