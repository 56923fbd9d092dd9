// csi_internal.h -- launchers shared between the translation units of libclimaseaice_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/climaseaice_b200.h"
#include "csi_types.cuh"

namespace csi {

struct LaunchCtx {
    cudaStream_t stream;
    int64_t *launches;  // incremented once per kernel launch
};

// ---- unfused path (csi_unfused.cu) ------------------------------------------------------------
void launch_initialize_rheology(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f);
void launch_evp_stress(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt);
void launch_u_step(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, Range2 r);
void launch_v_step(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, Range2 r);

// ---- halo fills (csi_halo.cu): which 0 = default BCs, 1 = u, 2 = v -----------------------------
void launch_fill_halo(const LaunchCtx &c, const DGrid &g, const DParams &p, const DArr &a, int lx, int ly, int which);
void launch_mask_immersed(const LaunchCtx &c, const DGrid &g, const DArr &a, int lx, int ly);

// ---- advection (csi_advection.cu) ------------------------------------------------------------
void launch_tracer_tendencies(const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f);
void launch_thermodynamics(const LaunchCtx &c, const DGrid &g, const csi_thermo_config &p, const DThermoFields &f, double rho_i, double dt);
void launch_dynamic_step(const LaunchCtx &c, const DGrid &g, const DFields &f, const DArr &hn, const DArr &an, const DArr &hsn, double dt);

// ---- reductions (csi_reduce.cu); results land in `scratch` (device), final value in out_dev ------
void launch_cfl(const LaunchCtx &c, const DGrid &g, const DFields &f, double *scratch, int nscratch, double *out_dev);
void launch_diagnostics(const LaunchCtx &c, const DGrid &g, const DFields &f, double *scratch, int nscratch, double *out_dev5);

// ---- fused path (csi_fused.cu) ---------------------------------------------------------------
struct FusedPlan;  // opaque; owns ping-pong buffers and tensor maps
int fused_supported(const DGrid &g, const DParams &p, const DFields &f, char *why, int nwhy);
void launch_fold_list(const LaunchCtx &c, const DArr &a, const int32_t *target, const int32_t *source, int n, double sign);
const char *fused_metrics_check(const DGrid &g);   // NULL, or why the grid's metric arrays rule the fused kernel out (csi_create)
FusedPlan *fused_create(const DGrid &g, const DParams &p, char *err, int nerr);
void fused_destroy(FusedPlan *);
// runs `nsub` substeps starting at substep index `first_sub` (1-based parity as in se.jl:173-189)
int fused_run(FusedPlan *, const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, int first_sub,
              int nsub, char *err, int nerr);

int fused_begin(FusedPlan *, const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, char *err, int nerr);
int fused_steps(FusedPlan *, const LaunchCtx &c, int first_sub, int nsub, bool aux_last, char *err, int nerr, cudaEvent_t halo_ready);
void fused_views(const FusedPlan *, DArr out[5]);
void fused_stats(const FusedPlan *, long long out[3]);
int fused_end(FusedPlan *, const LaunchCtx &c, const DFields &f, char *err, int nerr);

namespace fz {
int selftest_math(long long samples, unsigned long long seed, int span, unsigned long long *out5);
}

}  // namespace csi
