// csi_api.cu -- the C ABI of libclimaseaice_b200.so (include/climaseaice_b200.h): handle,
// argument checking, and the host-side orchestration of the hot path, i.e. the bodies of
//   time_step_momentum!   src/SeaIceDynamics/split_explicit_momentum_equations.jl:103-195
//   rk_substep!/time_step! src/sea_ice_rk_substep.jl:29-94, src/sea_ice_fe_step.jl:13-50
//   update_state!         src/sea_ice_model.jl:379-394
// as sequences of asynchronous kernel launches on the caller's stream.  No CPU fallback exists:
// every compute entry point needs a CUDA device of compute capability 10.x.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "csi_internal.h"
#include "csi_math.cuh"

using namespace csi;

namespace {

thread_local std::string g_create_error;

// ---- NCCL through dlopen: single-GPU users never need the library ------------------------------
typedef struct { char internal[128]; } nccl_uid;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(nccl_uid *) = nullptr;
    int (*CommInitRank)(void **, int, nccl_uid, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};
NcclApi &nccl()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (api.lib) {
#define CSI_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, name))
            CSI_SYM(GetUniqueId, "ncclGetUniqueId");
            CSI_SYM(CommInitRank, "ncclCommInitRank");
            CSI_SYM(CommDestroy, "ncclCommDestroy");
            CSI_SYM(Send, "ncclSend");
            CSI_SYM(Recv, "ncclRecv");
            CSI_SYM(AllReduce, "ncclAllReduce");
            CSI_SYM(GroupStart, "ncclGroupStart");
            CSI_SYM(GroupEnd, "ncclGroupEnd");
            CSI_SYM(GetErrorString, "ncclGetErrorString");
#undef CSI_SYM
            api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.Send && api.Recv && api.GroupStart && api.GroupEnd;
        }
    }
    return api;
}
const int NCCL_FLOAT64 = 8, NCCL_INT32 = 2, NCCL_MIN = 3;

}  // namespace

struct csi_handle {
    csi_config cfg;
    DGrid g;
    DParams p;
    int64_t launches = 0;
    double last_ms = 0.0;
    std::string err;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    double *scratch = nullptr;
    int nscratch = 0;
    double *out_dev = nullptr;
    uint8_t *mask_dev = nullptr;
    double *met_dev = nullptr;
    double *fff_dev = nullptr;
    std::vector<int32_t *> fold_dev;
    std::vector<int32_t> fold_host[8];   // host copies of the fold lists: target / source of the four locations
    std::vector<uint8_t> mask_host;
    std::vector<double> met_host, fff_host;
    FusedPlan *fused = nullptr;
    bool fused_failed = false;
    int solver_agreed = 0;  // partitions: 0 = not yet agreed, 1 = every rank runs the fused solver, 2 = every rank the general kernels
    int *agree_dev = nullptr;
    // device mirrors for the *_host entry points, in csi_fields member order
    std::vector<double *> mirror;
    std::vector<size_t> mirror_n;
    cudaStream_t own_stream = nullptr;
    size_t last_h2d = 0, last_d2h = 0;  // bytes moved by the last *_host call
    // attached thermodynamics (csi_attach_thermodynamics)
    bool thermo_on = false;
    csi_thermo_config thermo_cfg;
    DThermoFields thermo_f;
    // slab partition
    void *comm = nullptr;
    int rank = 0, nranks = 1;
    int Rx = 1, Ry = 1, rx = 0, ry = 0;  // 2-D partition: rank = ry * Rx + rx
    double *xbuf = nullptr;              // packed west/east strips: [send_w | send_e | recv_w | recv_e]
    size_t xbuf_each = 0;
    cudaStream_t comm_stream = nullptr;  // halo exchanges overlapped with interior tiles (slabs, fused solver)
    cudaEvent_t ev_block = nullptr, ev_halo = nullptr;
    // asynchronous exchange (fill_halo_regions!(...; async = true) + synchronize_communication!, evp:204-206,275-280)
    cudaEvent_t ev_async_in = nullptr, ev_async_done = nullptr;
    bool halo_pending = false;
};

namespace {

int fail(csi_handle *h, int code, const std::string &msg)
{
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}
int cuda_fail(csi_handle *h, cudaError_t e, const char *where)
{
    return fail(h, (int)e > 0 ? (int)e : 1, std::string(where) + ": " + cudaGetErrorString(e));
}
#define CSI_CUDA(h, call)                                      \
    do {                                                       \
        cudaError_t e__ = (call);                              \
        if (e__ != cudaSuccess) return cuda_fail(h, e__, #call); \
    } while (0)

constexpr int NFIELDS = sizeof(csi_fields) / sizeof(csi_array);
struct FieldInfo {
    const char *name;
    int lx, ly;
};
// locations in csi_fields member order
const FieldInfo FIELD_INFO[NFIELDS] = {
    {"u", 1, 0},     {"v", 0, 1},     {"h", 0, 0},      {"a", 0, 0},      {"s11", 0, 0},   {"s22", 0, 0},
    {"s12", 1, 1},   {"zeta_f", 1, 1}, {"zeta_c", 0, 0}, {"delta", 0, 0},  {"alpha", 0, 0}, {"un", 1, 0},
    {"vn", 0, 1},    {"P", 0, 0},     {"top_x", 1, 0},  {"top_y", 0, 1},  {"ue", 1, 0},    {"ve", 0, 1},
    {"Gh", 0, 0},    {"Ga", 0, 0},    {"hm", 0, 0},     {"am", 0, 0},     {"um", 1, 0},    {"vm", 0, 1},
    {"hs", 0, 0},    {"Ghs", 0, 0},   {"hsm", 0, 0},    {"fd_u", 1, 0},   {"fd_v", 0, 1}};
enum { K_UE = 16, K_VE = 17, K_HS = 24, K_GHS = 25, K_HSM = 26, K_FDU = 27, K_FDV = 28 };

const csi_array &field_at(const csi_fields &f, int k) { return reinterpret_cast<const csi_array *>(&f)[k]; }
csi_array &field_at(csi_fields &f, int k) { return reinterpret_cast<csi_array *>(&f)[k]; }

int check_array(csi_handle *h, const csi_array &a, const FieldInfo &fi, bool required)
{
    if (!a.ptr) return required ? fail(h, CSI_ERR_ARG, std::string("field '") + fi.name + "' is required but NULL") : CSI_OK;
    const csi_config &c = h->cfg;
    // Face fields carry N+1 points along a Bounded axis, except along a partitioned axis (the wall point lives in the halo)
    const int ex = c.Nx + 2 * c.Hx + ((fi.lx && c.topo_x == CSI_BOUNDED && h->Rx == 1) ? 1 : 0);
    const int ey = c.Ny + 2 * c.Hy + ((fi.ly && (c.topo_y == CSI_BOUNDED || c.topo_y == CSI_FOLDED) && h->nranks == 1) ? 1 : 0);
    if (a.nx_tot != ex || a.ny_tot != ey || a.off_x != c.Hx || a.off_y != c.Hy) {
        char buf[256];
        snprintf(buf, sizeof buf, "field '%s': parent %dx%d offsets (%d,%d), expected %dx%d offsets (%d,%d)", fi.name, a.nx_tot,
                 a.ny_tot, a.off_x, a.off_y, ex, ey, c.Hx, c.Hy);
        return fail(h, CSI_ERR_SHAPE, buf);
    }
    return CSI_OK;
}

DArr to_darr(const csi_array &a)
{
    DArr d;
    d.p = a.ptr;
    d.sx = a.nx_tot;
    d.sy = a.ny_tot;
    d.ox = a.off_x;
    d.oy = a.off_y;
    return d;
}

// which fields each entry point needs
enum Need { NEED_MOMENTUM = 1, NEED_TRACERS = 2, NEED_RK = 4 };

int convert_fields(csi_handle *h, const csi_fields *f, int need, DFields *out)
{
    if (!h) return CSI_ERR_ARG;
    if (!f) return fail(h, CSI_ERR_ARG, "csi_fields pointer is NULL");
    const csi_config &c = h->cfg;
    for (int k = 0; k < NFIELDS; k++) {
        bool req = false;
        const std::string n = FIELD_INFO[k].name;
        if (need & NEED_MOMENTUM) {
            req |= (n == "u" || n == "v" || n == "h" || n == "a" || n == "s11" || n == "s22" || n == "s12" || n == "zeta_f" ||
                    n == "zeta_c" || n == "delta" || n == "alpha" || n == "un" || n == "vn" || n == "P");
            if (c.top_stress_kind == CSI_STRESS_FIELD) req |= (n == "top_x" || n == "top_y");
            if (c.bottom_stress_kind == CSI_STRESS_FIELD) req |= (n == "ue" || n == "ve");
            if (c.free_drift_kind == CSI_FD_FIELDS) req |= (n == "fd_u" || n == "fd_v");
        }
        const bool snow = field_at(*f, K_HS).ptr != nullptr;
        if (need & NEED_TRACERS) req |= (n == "u" || n == "v" || n == "h" || n == "a" || n == "Gh" || n == "Ga" || (snow && n == "Ghs"));
        if ((need & NEED_RK) && c.timestepper == CSI_RK3) req |= (n == "hm" || n == "am" || n == "um" || n == "vm" || (snow && n == "hsm"));
        int rc = check_array(h, field_at(*f, k), FIELD_INFO[k], req);
        if (rc) return rc;
        reinterpret_cast<DArr *>(out)[k] = to_darr(field_at(*f, k));
    }
    if ((field_at(*f, K_UE).ptr == nullptr) != (field_at(*f, K_VE).ptr == nullptr))
        return fail(h, CSI_ERR_ARG, "ue and ve must both be arrays or both be constants");
    if ((field_at(*f, 14).ptr == nullptr) != (field_at(*f, 15).ptr == nullptr))
        return fail(h, CSI_ERR_ARG, "top_x and top_y must both be arrays or both be constants");
    return CSI_OK;
}

constexpr int NTHERMO = sizeof(csi_thermo_fields) / sizeof(csi_array);
const char *THERMO_NAMES[NTHERMO] = {"h", "a", "hs", "Tu", "Tus", "S", "hc", "Qtop", "Qbot", "Sb", "Tb", "snowfall", "rho_s", "mf_ice", "mf_snow", "mf_snowfall"};

int convert_thermo(csi_handle *h, const csi_thermo_config *cfg, const csi_thermo_fields *f, DThermoFields *out)
{
    if (!h) return CSI_ERR_ARG;
    if (!cfg || !f) return fail(h, CSI_ERR_ARG, "csi_thermo_config / csi_thermo_fields pointer is NULL");
    if (cfg->n_top_terms < 1 || cfg->n_top_terms > 2) return fail(h, CSI_ERR_ARG, "thermodynamics: n_top_terms must be 1 or 2");
    bool need_qtop = false;
    for (int t = 0; t < cfg->n_top_terms; t++) {
        const int k = cfg->top_term_kind[t];
        if (k < CSI_FLUX_CONST || k > CSI_FLUX_LINEAR) return fail(h, CSI_ERR_ARG, "thermodynamics: bad top flux term kind");
        need_qtop |= k == CSI_FLUX_ARRAY;
    }
    for (int bc : {cfg->top_heat_bc, cfg->snow_top_heat_bc})
        if (bc != CSI_TOP_MELTING_CONSTRAINED_FLUX_BALANCE && bc != CSI_TOP_PRESCRIBED_TEMPERATURE) return fail(h, CSI_ERR_ARG, "thermodynamics: bad top heat boundary condition");
    if (cfg->bottom_heat_bc != CSI_BOTTOM_ICE_WATER_EQUILIBRIUM && cfg->bottom_heat_bc != CSI_BOTTOM_PRESCRIBED_TEMPERATURE)
        return fail(h, CSI_ERR_ARG, "thermodynamics: bad bottom heat boundary condition");
    if (!(cfg->secant_tolerance > 0) || cfg->secant_maxiters < 1) return fail(h, CSI_ERR_ARG, "thermodynamics: secant_tolerance and secant_maxiters must be positive");
    for (int k = 0; k < NTHERMO; k++) {
        const std::string n = THERMO_NAMES[k];
        bool req = (n == "h" || n == "a" || n == "Tu");
        if (cfg->layered) req |= (n == "hs" || n == "Tus");
        if (need_qtop) req |= n == "Qtop";
        FieldInfo fi{THERMO_NAMES[k], 0, 0};
        const csi_array &a = reinterpret_cast<const csi_array *>(f)[k];
        int rc = check_array(h, a, fi, req);
        if (rc) return rc;
        reinterpret_cast<DArr *>(out)[k] = to_darr(a);
    }
    return CSI_OK;
}

struct Timed {
    csi_handle *h;
    cudaStream_t s;
    Timed(csi_handle *h_, cudaStream_t s_) : h(h_), s(s_) { cudaEventRecord(h->ev0, s); }
    ~Timed()
    {
        cudaEventRecord(h->ev1, s);
        h->timed = true;
    }
};

int copy_parent(csi_handle *h, const DArr &dst, const DArr &src, cudaStream_t s)
{
    CSI_CUDA(h, cudaMemcpyAsync(dst.p, src.p, sizeof(double) * (size_t)src.sx * src.sy, cudaMemcpyDeviceToDevice, s));
    return CSI_OK;
}

Range2 velocity_range(const DGrid &g)
{
    // `:xy` on serial grids (se.jl:31); widened into the halo on connected sides (se.jl:40-46)
    Range2 r{1, g.Nx, 1, g.Ny};
    if (g.conn_s) r.j0 = -g.Hy + 2;
    if (g.conn_n) r.j1 = g.Ny + g.Hy - 1;
    if (g.conn_w) r.i0 = -g.Hx + 2;
    if (g.conn_e) r.i1 = g.Nx + g.Hx - 1;
    return r;
}

int exchange_slab_halos(csi_handle *h, const DArr *arrs, int n, int width, cudaStream_t s);

int ensure_comm_stream(csi_handle *h)
{
    if (!h->comm_stream) {
        CSI_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
        CSI_CUDA(h, cudaEventCreateWithFlags(&h->ev_block, cudaEventDisableTiming));
        CSI_CUDA(h, cudaEventCreateWithFlags(&h->ev_halo, cudaEventDisableTiming));
        CSI_CUDA(h, cudaEventCreateWithFlags(&h->ev_async_in, cudaEventDisableTiming));
        CSI_CUDA(h, cudaEventCreateWithFlags(&h->ev_async_done, cudaEventDisableTiming));
    }
    return CSI_OK;
}
// fill_halo_regions!(fields; async = true): the exchange is ordered behind everything already on `s` and runs on the
// handle's communication stream; `s` does not wait for it.  wait_halos = synchronize_communication! (stream-ordered).
int exchange_halos_async(csi_handle *h, const DArr *arrs, int n, int width, cudaStream_t s)
{
    if (h->nranks <= 1) return CSI_OK;
    int rc = ensure_comm_stream(h);
    if (rc) return rc;
    CSI_CUDA(h, cudaEventRecord(h->ev_async_in, s));
    CSI_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_async_in, 0));
    if ((rc = exchange_slab_halos(h, arrs, n, width, h->comm_stream))) return rc;
    CSI_CUDA(h, cudaEventRecord(h->ev_async_done, h->comm_stream));
    h->halo_pending = true;
    return CSI_OK;
}
int wait_halos(csi_handle *h, cudaStream_t s)
{
    if (!h->halo_pending) return CSI_OK;
    CSI_CUDA(h, cudaStreamWaitEvent(s, h->ev_async_done, 0));
    h->halo_pending = false;
    return CSI_OK;
}

// time_step_momentum!  (se.jl:103-195).  defer_sigma: the stress halo exchange of finalize_rheology! is left in flight
// (evp:275-280, async = true); the next momentum step -- or the caller, through wait_halos -- synchronises it (evp:204-206)
int momentum_impl(csi_handle *h, const DFields &f, double dt, int nsub, cudaStream_t s, bool defer_sigma = false)
{
    LaunchCtx c{s, &h->launches};
    const DGrid &g = h->g;
    const DParams &p = h->p;
    {
        int rc = wait_halos(h, s);  // synchronize_communication!(fields.sigma*)  evp:204-206
        if (rc) return rc;
    }
    if (h->cfg.timestepper == CSI_RK3 && f.um.p && f.vm.p) {  // reset_velocities!  se.jl:87-93
        int rc = copy_parent(h, f.u, f.um, s);
        if (rc) return rc;
        rc = copy_parent(h, f.v, f.vm, s);
        if (rc) return rc;
    }
    launch_initialize_rheology(c, g, p, f);  // evp.jl:192-216
    // update_external_stress!  ext.jl:72-78,148-152
    if (p.top_kind == CSI_STRESS_FIELD || p.top_kind == CSI_STRESS_SEMI_IMPLICIT) {
        launch_fill_halo(c, g, p, f.top_x, 1, 0, 3);
        launch_fill_halo(c, g, p, f.top_y, 0, 1, 3);
    }
    if (p.bot_kind == CSI_STRESS_FIELD || p.bot_kind == CSI_STRESS_SEMI_IMPLICIT) {
        launch_fill_halo(c, g, p, f.ue, 1, 0, 3);
        launch_fill_halo(c, g, p, f.ve, 0, 1, 3);
    }
    launch_fill_halo(c, g, p, f.u, 1, 0, 1);  // se.jl:170-171
    launch_fill_halo(c, g, p, f.v, 0, 1, 2);
    if (h->nranks > 1) {
        // the reference fills the external fields' halos with a full (communicating) fill_halo_regions!
        // (ext.jl:57-61,72-78); P, un, vn are computed over the whole parent from exchanged h, aice, u, v
        const DArr ext[4] = {f.top_x, f.top_y, f.ue, f.ve};
        int rc = exchange_slab_halos(h, ext, 4, g.Hy, s);
        if (rc) return rc;
    }

    bool use_fused = false;
    std::string why_not;
    if (h->cfg.solver_impl != CSI_SOLVER_UNFUSED && !h->fused_failed) {
        char why[256] = {0};
        use_fused = fused_supported(g, p, f, why, sizeof why) != 0;
        if (!use_fused) why_not = std::string("fused solver: ") + why;
        if (!use_fused && h->cfg.solver_impl == CSI_SOLVER_FUSED && h->nranks == 1) return fail(h, CSI_ERR_UNSUPPORTED, why_not);
    }
    if (use_fused && !h->fused) {
        char err[256] = {0};
        h->fused = fused_create(g, p, err, sizeof err);
        if (!h->fused) {
            h->fused_failed = true;
            use_fused = false;
            why_not = std::string("fused_create: ") + err;
            if (h->nranks == 1) return fail(h, 1, why_not);
        }
    }
    if (h->nranks > 1 && h->cfg.solver_impl != CSI_SOLVER_UNFUSED) {
        // Every rank of a partition must take the same path: the two solvers exchange different buffers (internal planes vs
        // the caller's parents), and a rank that fell back or failed alone would leave its neighbours blocked in NCCL.  The
        // choice depends on rank-local data (metric rows, device memory), so it is agreed once: min over the ranks.
        if (h->solver_agreed == 0) {
            if (!h->comm) return fail(h, CSI_ERR_ARG, "csi_comm_init has not been called on this handle");
            NcclApi &api = nccl();
            if (!api.AllReduce) return fail(h, CSI_ERR_NCCL_MISSING, "ncclAllReduce missing");
            if (!h->agree_dev) CSI_CUDA(h, cudaMalloc(&h->agree_dev, sizeof(int)));
            int mine = use_fused ? 1 : 0;
            CSI_CUDA(h, cudaMemcpyAsync(h->agree_dev, &mine, sizeof(int), cudaMemcpyHostToDevice, s));
            int nrc = api.AllReduce(h->agree_dev, h->agree_dev, 1, NCCL_INT32, NCCL_MIN, h->comm, s);
            if (nrc != 0) return fail(h, 1000 + nrc, "ncclAllReduce (solver agreement) failed");
            int all = 0;
            CSI_CUDA(h, cudaMemcpyAsync(&all, h->agree_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
            CSI_CUDA(h, cudaStreamSynchronize(s));
            h->solver_agreed = all ? 1 : 2;
            if (!all && use_fused) why_not = "fused solver: another rank of the partition cannot run it";
        }
        use_fused = h->solver_agreed == 1;
        // (collectively: every rank returns here when any rank cannot run an explicitly requested fused solver)
        if (!use_fused && h->cfg.solver_impl == CSI_SOLVER_FUSED) return fail(h, CSI_ERR_UNSUPPORTED, why_not.empty() ? "fused solver unavailable on this partition" : why_not);
    }
    if (use_fused) {
        char err[256] = {0};
        // pack -> blocks of K substeps with a slab halo exchange of the internal (double-buffered) fields
        // in between -> unpack.  K = exchange_every (halo Hy >= 2K+3, se.jl:55-56); one block when serial.
        int rc = fused_begin(h->fused, c, g, p, f, dt, err, sizeof err);
        if (rc) return fail(h, rc, std::string("fused_begin: ") + err);
        const int K = (h->nranks > 1 && h->cfg.exchange_every > 0) ? h->cfg.exchange_every : nsub;
        for (int sub = 1; sub <= nsub; sub += K) {
            const int n = std::min(K, nsub - sub + 1);
            cudaEvent_t halo_ready = nullptr;
            if (h->nranks > 1 && sub > 1) {
                DArr views[5];
                fused_views(h->fused, views);
                // slabs: the exchange runs on its own stream behind the block just launched, and the next substep's
                // interior tile rows do not wait for it (north_star: halo exchange overlapped with interior compute)
                const bool overlap = h->Rx == 1 && !h->cfg.serial_exchange;
                if (overlap) {
                    if ((rc = ensure_comm_stream(h))) return rc;
                    CSI_CUDA(h, cudaEventRecord(h->ev_block, s));
                    CSI_CUDA(h, cudaStreamWaitEvent(h->comm_stream, h->ev_block, 0));
                    if ((rc = exchange_slab_halos(h, views, 5, g.Hy, h->comm_stream))) return rc;
                    CSI_CUDA(h, cudaEventRecord(h->ev_halo, h->comm_stream));
                    halo_ready = h->ev_halo;
                } else if ((rc = exchange_slab_halos(h, views, 5, g.Hy, s))) return rc;
            }
            rc = fused_steps(h->fused, c, sub, n, sub + n - 1 == nsub, err, sizeof err, halo_ready);
            if (rc) return fail(h, rc, std::string("fused_steps: ") + err);
        }
        if (nsub > 0 && (rc = fused_end(h->fused, c, f, err, sizeof err))) return fail(h, rc, std::string("fused_end: ") + err);
        // the in-loop fills of se.jl:180-187 happen inside the fused kernel on its internal layout;
        // refresh the caller-visible halos once
        launch_fill_halo(c, g, p, f.u, 1, 0, 1);
        launch_fill_halo(c, g, p, f.v, 0, 1, 2);
    } else {
        const Range2 r = velocity_range(g);
        const int K = h->cfg.exchange_every > 0 ? h->cfg.exchange_every : nsub;
        for (int sub = 1; sub <= nsub; sub++) {  // se.jl:173-189
            if (h->nranks > 1 && sub > 1 && (sub - 1) % K == 0) {
                const DArr arrs[5] = {f.u, f.v, f.s11, f.s22, f.s12};
                int rc = exchange_slab_halos(h, arrs, 5, g.Hy, s);
                if (rc) return rc;
            }
            launch_evp_stress(c, g, p, f, dt);
            if (sub % 2 == 0) {
                launch_u_step(c, g, p, f, dt, r);
                launch_fill_halo(c, g, p, f.u, 1, 0, 1);
                launch_v_step(c, g, p, f, dt, r);
                launch_fill_halo(c, g, p, f.v, 0, 1, 2);
            } else {
                launch_v_step(c, g, p, f, dt, r);
                launch_fill_halo(c, g, p, f.v, 0, 1, 2);
                launch_u_step(c, g, p, f, dt, r);
                launch_fill_halo(c, g, p, f.u, 1, 0, 1);
            }
        }
    }
    // finalize_rheology!  evp.jl:275-280
    launch_fill_halo(c, g, p, f.s11, 0, 0, 0);
    launch_fill_halo(c, g, p, f.s12, 1, 1, 0);
    launch_fill_halo(c, g, p, f.s22, 0, 0, 0);
    if (h->nranks > 1) {
        const DArr arrs[3] = {f.s11, f.s12, f.s22};
        // (slabs only: the rows go zero-copy; the packed strips of a 2-D partition share one staging buffer with the
        // exchanges that follow on the compute stream)
        int rc = (defer_sigma && h->Rx == 1) ? exchange_halos_async(h, arrs, 3, g.Hy, s) : exchange_slab_halos(h, arrs, 3, g.Hy, s);
        if (rc) return rc;
    }
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

// update_state!  (sea_ice_model.jl:379-394): prognostic order h, aice, u, v
int update_state_impl(csi_handle *h, const DFields &f, cudaStream_t s)
{
    LaunchCtx c{s, &h->launches};
    const DGrid &g = h->g;
    const DParams &p = h->p;
    launch_mask_immersed(c, g, f.h, 0, 0);
    launch_fill_halo(c, g, p, f.h, 0, 0, 0);
    launch_mask_immersed(c, g, f.a, 0, 0);
    launch_fill_halo(c, g, p, f.a, 0, 0, 0);
    if (f.hs.p) {  // snow_fields(model.snow_thickness) follows h, aice among the prognostic fields
        launch_mask_immersed(c, g, f.hs, 0, 0);
        launch_fill_halo(c, g, p, f.hs, 0, 0, 0);
    }
    launch_mask_immersed(c, g, f.u, 1, 0);
    launch_fill_halo(c, g, p, f.u, 1, 0, 1);
    launch_mask_immersed(c, g, f.v, 0, 1);
    launch_fill_halo(c, g, p, f.v, 0, 1, 2);
    if (h->thermo_on) {  // sea_ice_model.jl:386-389
        launch_mask_immersed(c, g, h->thermo_f.mf_ice, 0, 0);
        launch_mask_immersed(c, g, h->thermo_f.mf_snow, 0, 0);
        launch_mask_immersed(c, g, h->thermo_f.mf_snowfall, 0, 0);
    }
    if (h->nranks > 1) {
        const DArr arrs[5] = {f.h, f.a, f.u, f.v, f.hs};
        int rc = exchange_slab_halos(h, arrs, 5, g.Hy, s);
        if (rc) return rc;
    }
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

int time_step_impl(csi_handle *h, const DFields &f, double dt, int first, cudaStream_t s)
{
    LaunchCtx c{s, &h->launches};
    const DGrid &g = h->g;
    const DParams &p = h->p;
    int rc;
    if (first && (rc = update_state_impl(h, f, s))) return rc;
    const int nsub = h->cfg.substeps;
    if (h->cfg.timestepper == CSI_FE) {  // fe.jl:13-34
        launch_tracer_tendencies(c, g, p, f);
        if ((rc = momentum_impl(h, f, dt, nsub, s, true))) return rc;
        launch_dynamic_step(c, g, f, f.h, f.a, f.hs, dt);
        if (h->thermo_on) launch_thermodynamics(c, g, h->thermo_cfg, h->thermo_f, h->cfg.ice_density, dt);  // fe.jl:30
        if ((rc = update_state_impl(h, f, s))) return rc;
        return wait_halos(h, s);  // the caller sees exchanged stress halos
    }
    // cache_current_fields!  rk.jl:29-42
    if ((rc = copy_parent(h, f.hm, f.h, s))) return rc;
    if ((rc = copy_parent(h, f.am, f.a, s))) return rc;
    if (f.hs.p && (rc = copy_parent(h, f.hsm, f.hs, s))) return rc;
    if ((rc = copy_parent(h, f.um, f.u, s))) return rc;
    if ((rc = copy_parent(h, f.vm, f.v, s))) return rc;
    for (int beta = 3; beta >= 1; beta--) {  // SplitRungeKutta3: dtau = dt / beta
        const double dtau = dt / beta;
        launch_tracer_tendencies(c, g, p, f);                      // rk.jl:84
        if ((rc = momentum_impl(h, f, dtau, nsub, s, true))) return rc;  // rk.jl:87 (stress halos exchanged asynchronously, evp:275-280)
        launch_dynamic_step(c, g, f, f.hm, f.am, f.hsm, dtau);     // rk.jl:89
        if (h->thermo_on) launch_thermodynamics(c, g, h->thermo_cfg, h->thermo_f, h->cfg.ice_density, dtau);  // rk.jl:91
        if ((rc = update_state_impl(h, f, s))) return rc;
    }
    return wait_halos(h, s);  // the caller sees exchanged stress halos
}

// Packed west/east strips of a 2-D partition: `width` columns x every parent row of each array, array after array.
// All rows, not 1..Ny: a y axis that is not partitioned has its halo rows filled locally BEFORE the exchange, and the
// corners (x halo columns of those rows) must come from the neighbour's already filled rows; with a partitioned y axis
// the row exchange that follows overwrites the corners anyway.
__global__ void k_pack_x_strips(const DArr a, double *send_w, double *send_e, int Nx, int width)
{
    const int pj = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;  // parent row, column within the strip
    if (pj >= a.sy) return;
    const int j = pj + 1 - a.oy;
    send_w[(size_t)pj * width + k] = at(a, 1 + k, j);               // interior columns 1..width
    send_e[(size_t)pj * width + k] = at(a, Nx - width + 1 + k, j);  // interior columns Nx-width+1..Nx
}
__global__ void k_unpack_x_strips(DArr a, const double *recv_w, const double *recv_e, int Nx, int width, int do_w, int do_e)
{
    const int pj = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;
    if (pj >= a.sy) return;
    const int j = pj + 1 - a.oy;
    if (do_w) at(a, 1 - width + k, j) = recv_w[(size_t)pj * width + k];  // halo columns 1-width..0
    if (do_e) at(a, Nx + 1 + k, j) = recv_e[(size_t)pj * width + k];      // halo columns Nx+1..Nx+width
}

// Distributed fill_halo_regions! of a partition Rx x Ry (rank = ry Rx + rx).  West/east first: strips of `width` columns
// are packed, sent and unpacked; then south/north: `width` whole parent rows -- x halos included, which carries the
// corners -- go zero-copy (rows are contiguous in the i-fastest layout).
int exchange_slab_halos(csi_handle *h, const DArr *arrs, int n, int width, cudaStream_t s)
{
    if (h->nranks <= 1) return CSI_OK;
    if (!h->comm) return fail(h, CSI_ERR_ARG, "csi_comm_init has not been called on this handle");
    NcclApi &api = nccl();
    const DGrid &g = h->g;
    // conn_* already encode the topology: a Bounded axis has no wrap-around neighbour
    const int south = g.conn_s ? ((h->ry + h->Ry - 1) % h->Ry) * h->Rx + h->rx : -1;
    const int north = (g.conn_n && !g.fold) ? ((h->ry + 1) % h->Ry) * h->Rx + h->rx : -1;  // (a fold is filled locally)
    const int west = g.conn_w ? h->ry * h->Rx + (h->rx + h->Rx - 1) % h->Rx : -1;
    const int east = g.conn_e ? h->ry * h->Rx + (h->rx + 1) % h->Rx : -1;
    if (west >= 0 || east >= 0) {
        const int wx = width > g.Hx ? g.Hx : width;
        size_t total = 0;
        for (int k = 0; k < n; k++)
            if (arrs[k].p) total += (size_t)wx * arrs[k].sy;
        if (h->xbuf_each < total) {
            if (h->xbuf) cudaFree(h->xbuf);
            h->xbuf = nullptr;
            h->xbuf_each = 0;
            CSI_CUDA(h, cudaMalloc(&h->xbuf, 4 * total * sizeof(double)));
            h->xbuf_each = total;
        }
        double *send_w = h->xbuf, *send_e = h->xbuf + h->xbuf_each, *recv_w = h->xbuf + 2 * h->xbuf_each, *recv_e = h->xbuf + 3 * h->xbuf_each;
        size_t off = 0;
        for (int k = 0; k < n; k++) {
            if (!arrs[k].p) continue;
            k_pack_x_strips<<<dim3((arrs[k].sy + 127) / 128, wx), 128, 0, s>>>(arrs[k], send_w + off, send_e + off, g.Nx, wx);
            ++h->launches;
            off += (size_t)wx * arrs[k].sy;
        }
        api.GroupStart();
        // same pairing rule as below when both neighbours are the same rank
        if (west >= 0) api.Send(send_w, total, NCCL_FLOAT64, west, h->comm, s);
        if (east >= 0) api.Recv(recv_e, total, NCCL_FLOAT64, east, h->comm, s);
        if (east >= 0) api.Send(send_e, total, NCCL_FLOAT64, east, h->comm, s);
        if (west >= 0) api.Recv(recv_w, total, NCCL_FLOAT64, west, h->comm, s);
        int rc = api.GroupEnd();
        if (rc != 0) return fail(h, 1000 + rc, std::string("ncclGroupEnd: ") + (api.GetErrorString ? api.GetErrorString(rc) : "error"));
        off = 0;
        for (int k = 0; k < n; k++) {
            if (!arrs[k].p) continue;
            k_unpack_x_strips<<<dim3((arrs[k].sy + 127) / 128, wx), 128, 0, s>>>(arrs[k], recv_w + off, recv_e + off, g.Nx, wx, west >= 0, east >= 0);
            ++h->launches;
            off += (size_t)wx * arrs[k].sy;
        }
        CSI_CUDA(h, cudaGetLastError());
    }
    if (south < 0 && north < 0) return CSI_OK;
    if (width > g.Hy) width = g.Hy;
    api.GroupStart();
    for (int k = 0; k < n; k++) {
        const DArr &a = arrs[k];
        if (!a.p) continue;
        const size_t row = (size_t)a.sx, cnt = row * width;
        // rows are contiguous in the i-fastest layout: zero-copy sends
        double *send_s = a.p + (size_t)a.oy * row;                       // interior rows 1..width
        double *recv_s = a.p + (size_t)(a.oy - width) * row;             // halo rows 1-width..0
        double *send_n = a.p + (size_t)(a.oy + g.Ny - width) * row;      // interior rows Ny-width+1..Ny
        double *recv_n = a.p + (size_t)(a.oy + g.Ny) * row;              // halo rows Ny+1..Ny+width
        // order matters when both neighbours are the same rank (2 slabs, periodic): my southward send
        // must pair with the peer's receive from its north, and vice versa
        if (south >= 0) api.Send(send_s, cnt, NCCL_FLOAT64, south, h->comm, s);
        if (north >= 0) api.Recv(recv_n, cnt, NCCL_FLOAT64, north, h->comm, s);
        if (north >= 0) api.Send(send_n, cnt, NCCL_FLOAT64, north, h->comm, s);
        if (south >= 0) api.Recv(recv_s, cnt, NCCL_FLOAT64, south, h->comm, s);
    }
    int rc = api.GroupEnd();
    if (rc != 0) return fail(h, 1000 + rc, std::string("ncclGroupEnd: ") + (api.GetErrorString ? api.GetErrorString(rc) : "error"));
    return CSI_OK;
}

}  // namespace

// ================================================================================================
extern "C" {

int csi_version(void) { return CSI_ABI_VERSION; }

const char *csi_last_error(const csi_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int csi_create(const csi_config *cfg, csi_handle **out)
{
    if (!cfg || !out) return fail(nullptr, CSI_ERR_ARG, "csi_create: NULL argument");
    *out = nullptr;
    if (cfg->abi_version != CSI_ABI_VERSION) return fail(nullptr, CSI_ERR_ARG, "csi_create: abi_version mismatch");
    if (cfg->Nx < 1 || cfg->Ny < 1 || cfg->Hx < 1 || cfg->Hy < 1) return fail(nullptr, CSI_ERR_ARG, "csi_create: sizes and halos must be positive");
    if (cfg->Hx < 3 || cfg->Hy < 3) return fail(nullptr, CSI_ERR_ARG, "csi_create: halos must be at least 3 (the EVP stencils reach 2 cells)");
    const int B = cfg->advection_order <= 1 ? 1 : (cfg->advection_order + 1) / 2;
    if (cfg->advection_order != 0 && cfg->advection_order != 1 && cfg->advection_order != 3 && cfg->advection_order != 5 && cfg->advection_order != 7)
        return fail(nullptr, CSI_ERR_ARG, "csi_create: advection_order must be 0, 1, 3, 5 or 7");
    if (cfg->Hx < B || cfg->Hy < B) return fail(nullptr, CSI_ERR_ARG, "csi_create: halo smaller than the advection stencil");
    if ((cfg->topo_x != CSI_PERIODIC && cfg->topo_x != CSI_BOUNDED) || (cfg->topo_y != CSI_PERIODIC && cfg->topo_y != CSI_BOUNDED && cfg->topo_y != CSI_FOLDED))
        return fail(nullptr, CSI_ERR_ARG, "csi_create: topology must be CSI_PERIODIC or CSI_BOUNDED (topo_y: or CSI_FOLDED)");
    if (cfg->topo_y == CSI_FOLDED) {
        // the north fold of a tripolar grid: copy lists for the four locations, indices inside a parent of that location
        if (cfg->partition_x > 1) return fail(nullptr, CSI_ERR_UNSUPPORTED, "csi_create: a folded grid with a partition along x (the fold mirrors columns across ranks)");
        const int part = cfg->nranks > 1;
        const bool holder = !part || cfg->rank == cfg->nranks - 1;   // y-slabs: the last one holds the fold, the others ignore the lists
        for (int loc = 0; holder && loc < 4; loc++) {
            const int lx = loc & 1, ly = loc >> 1;
            const long ex = cfg->Nx + 2 * cfg->Hx + ((lx && cfg->topo_x == CSI_BOUNDED) ? 1 : 0), ey = cfg->Ny + 2 * cfg->Hy + ((ly && !part) ? 1 : 0);
            if (cfg->fold_count[loc] < 0 || (cfg->fold_count[loc] > 0 && (!cfg->fold_target[loc] || !cfg->fold_source[loc])))
                return fail(nullptr, CSI_ERR_ARG, "csi_create: CSI_FOLDED needs fold_target / fold_source for every location with fold_count > 0");
            for (int k = 0; k < cfg->fold_count[loc]; k++)
                if (cfg->fold_target[loc][k] < 0 || cfg->fold_target[loc][k] >= ex * ey || cfg->fold_source[loc][k] < 0 || cfg->fold_source[loc][k] >= ex * ey)
                    return fail(nullptr, CSI_ERR_ARG, "csi_create: a fold index lies outside the parent array of its location");
            // the copies of one list run concurrently: every target is written once, and nothing that is read is also written
            std::vector<char> role((size_t)(ex * ey), 0);
            for (int k = 0; k < cfg->fold_count[loc]; k++) {
                if (role[cfg->fold_target[loc][k]]) return fail(nullptr, CSI_ERR_ARG, "csi_create: a fold copy list names the same target twice");
                role[cfg->fold_target[loc][k]] = 1;
            }
            for (int k = 0; k < cfg->fold_count[loc]; k++)
                if (role[cfg->fold_source[loc][k]]) return fail(nullptr, CSI_ERR_ARG, "csi_create: a fold copy list reads an element that it also writes (order-dependent)");
        }
        if (holder && (cfg->fold_count[0] == 0 || cfg->fold_count[1] == 0 || cfg->fold_count[2] == 0 || cfg->fold_count[3] == 0))
            return fail(nullptr, CSI_ERR_ARG, "csi_create: CSI_FOLDED needs a copy list for each of the four locations");
        if (holder && (!(cfg->fold_sign_velocity == 1.0 || cfg->fold_sign_velocity == -1.0) || !(cfg->fold_sign_external == 1.0 || cfg->fold_sign_external == -1.0)))
            return fail(nullptr, CSI_ERR_ARG, "csi_create: fold signs must be +1 or -1");
    }
    if (cfg->metric_kind != CSI_METRIC_REGULAR && cfg->metric_kind != CSI_METRIC_J && cfg->metric_kind != CSI_METRIC_IJ)
        return fail(nullptr, CSI_ERR_ARG, "csi_create: metric_kind must be CSI_METRIC_REGULAR, CSI_METRIC_J or CSI_METRIC_IJ");
    if (cfg->metric_kind == CSI_METRIC_IJ) {
        const int L = cfg->Ny + 2 * cfg->Hy + 1, Wd = cfg->Nx + 2 * cfg->Hx + 1;
        if (cfg->coriolis_kind == CSI_CORIOLIS_SPHERICAL) return fail(nullptr, CSI_ERR_UNSUPPORTED, "csi_create: HydrostaticSphericalCoriolis with two-dimensional metrics");
        if (cfg->nranks > 1 && cfg->partition_x > 1) return fail(nullptr, CSI_ERR_UNSUPPORTED, "csi_create: partitions along x with two-dimensional metrics");
        for (int k = 0; k < 12; k++) {
            if (!cfg->metrics[k]) return fail(nullptr, CSI_ERR_ARG, "csi_create: CSI_METRIC_IJ needs all 12 metric arrays");
            // cells the stencils touch: i = 0 .. Nx+2, j = 0 .. Ny+2 (every halo row on a partition: widened windows, se:40-46)
            for (int j = (cfg->nranks > 1 ? 1 - cfg->Hy : 0); j <= (cfg->nranks > 1 ? cfg->Ny + cfg->Hy + 1 : cfg->Ny + 2); j++)
                for (int i = 0; i <= cfg->Nx + 2; i++)
                    if (j - 1 + cfg->Hy < L && i - 1 + cfg->Hx < Wd && !(cfg->metrics[k][(size_t)(j - 1 + cfg->Hy) * Wd + (i - 1 + cfg->Hx)] > 0))
                        return fail(nullptr, CSI_ERR_ARG, "csi_create: grid metrics must be positive");
        }
    }
    if (cfg->metric_kind == CSI_METRIC_REGULAR && (!(cfg->dx > 0) || !(cfg->dy > 0))) return fail(nullptr, CSI_ERR_ARG, "csi_create: dx, dy must be positive");
    if (cfg->metric_kind == CSI_METRIC_J) {
        const int L = cfg->Ny + 2 * cfg->Hy + 1;
        for (int k = 0; k < 12; k++) {
            if (!cfg->metrics[k]) return fail(nullptr, CSI_ERR_ARG, "csi_create: CSI_METRIC_J needs all 12 metric arrays");
            // rows the stencils touch: j = 0 .. Ny+2 (stress ring + the j+1 metrics it reads); on a partition the kernels run
            // over the widened range of se:40-46 and divide by the metrics of every halo row of a connected side
            const int Ry_ = cfg->nranks > 1 ? cfg->nranks / (cfg->partition_x > 1 ? cfg->partition_x : 1) : 1;
            const int ry_ = cfg->nranks > 1 ? cfg->rank / (cfg->partition_x > 1 ? cfg->partition_x : 1) : 0;
            const bool cs = Ry_ > 1 && (cfg->topo_y == CSI_PERIODIC || ry_ > 0), cn = (Ry_ > 1 && (cfg->topo_y == CSI_PERIODIC || ry_ < Ry_ - 1)) || cfg->topo_y == CSI_FOLDED;
            for (int j = cs ? 1 - cfg->Hy : 0; j <= (cn ? cfg->Ny + cfg->Hy + 1 : cfg->Ny + 2); j++)
                if (j - 1 + cfg->Hy >= 0 && j - 1 + cfg->Hy < L && !(cfg->metrics[k][j - 1 + cfg->Hy] > 0)) return fail(nullptr, CSI_ERR_ARG, "csi_create: grid metrics must be positive on every row the kernels touch (halo rows of connected sides included)");
        }
    }
    if (cfg->substeps < 1) return fail(nullptr, CSI_ERR_ARG, "csi_create: substeps must be >= 1");
    if (cfg->coriolis_kind < CSI_CORIOLIS_NONE || cfg->coriolis_kind > CSI_CORIOLIS_SPHERICAL) return fail(nullptr, CSI_ERR_ARG, "csi_create: bad coriolis_kind");
    if (cfg->coriolis_kind == CSI_CORIOLIS_SPHERICAL && !cfg->coriolis_f_ff) return fail(nullptr, CSI_ERR_ARG, "csi_create: CSI_CORIOLIS_SPHERICAL needs coriolis_f_ff");
    for (int kind : {cfg->top_stress_kind, cfg->bottom_stress_kind})
        if (kind < CSI_STRESS_NONE || kind > CSI_STRESS_SEMI_IMPLICIT) return fail(nullptr, CSI_ERR_ARG, "csi_create: bad stress kind");
    if (cfg->free_drift_kind < CSI_FD_NONE || cfg->free_drift_kind > CSI_FD_STRESS_BALANCE) return fail(nullptr, CSI_ERR_ARG, "csi_create: bad free_drift_kind");
    if (cfg->free_drift_kind == CSI_FD_STRESS_BALANCE) {
        // stress_balance_free_drift.jl:21-35,111-116
        const bool ts = cfg->top_stress_kind == CSI_STRESS_SEMI_IMPLICIT, bs = cfg->bottom_stress_kind == CSI_STRESS_SEMI_IMPLICIT;
        if (ts && bs) return fail(nullptr, CSI_ERR_ARG, "csi_create: StressBalanceFreeDrift supports a SemiImplicitStress only for the top or the bottom stress, not both");
        if (!ts && !bs) return fail(nullptr, CSI_ERR_ARG, "csi_create: StressBalanceFreeDrift requires a SemiImplicitStress for either the top or the bottom stress");
    }
    const int Rx = cfg->partition_x > 1 ? cfg->partition_x : 1;
    if (cfg->nranks > 1) {
        const int K = cfg->exchange_every > 0 ? cfg->exchange_every : cfg->substeps;
        if (cfg->nranks % Rx != 0) return fail(nullptr, CSI_ERR_ARG, "csi_create: nranks must be a multiple of partition_x");
        if (cfg->nranks / Rx > 1 && cfg->Hy < 2 * K + 3) return fail(nullptr, CSI_ERR_ARG, "csi_create: partitions along y need Hy >= 2*exchange_every + 3 (se.jl:55-56)");
        if (Rx > 1 && cfg->Hx < 2 * K + 3) return fail(nullptr, CSI_ERR_ARG, "csi_create: partitions along x need Hx >= 2*exchange_every + 3 (se.jl:55-56)");
    } else if (Rx > 1) return fail(nullptr, CSI_ERR_ARG, "csi_create: partition_x > 1 needs nranks > 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, CSI_ERR_NO_DEVICE, "csi_create: no CUDA device (libclimaseaice_b200 has no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, CSI_ERR_ARG, "csi_create: bad device ordinal");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, CSI_ERR_NO_DEVICE, "csi_create: cannot query device");
    if (prop.major != 10) return fail(nullptr, CSI_ERR_NO_DEVICE, "csi_create: kernels are built for sm_100a only");

    csi_handle *h = new csi_handle();
    h->cfg = *cfg;
    h->rank = cfg->nranks > 1 ? cfg->rank : 0;
    h->nranks = cfg->nranks > 1 ? cfg->nranks : 1;
    cudaSetDevice(cfg->device);
    DGrid &g = h->g;
    g.Nx = cfg->Nx; g.Ny = cfg->Ny; g.Hx = cfg->Hx; g.Hy = cfg->Hy;
    g.topo_x = cfg->topo_x;
    g.topo_y = cfg->topo_y == CSI_FOLDED ? CSI_BOUNDED : cfg->topo_y;  // a folded axis: a wall in the south, no wall in the north
    // partition Rx x Ry, rank = ry * Rx + rx: a side is "connected" when another rank owns the cells beyond it
    h->Rx = h->nranks > 1 ? Rx : 1;
    h->Ry = h->nranks / h->Rx;
    h->rx = h->rank % h->Rx;
    h->ry = h->rank / h->Rx;
    g.conn_s = h->Ry > 1 && (cfg->topo_y == CSI_PERIODIC || h->ry > 0);
    g.conn_n = h->Ry > 1 && (cfg->topo_y == CSI_PERIODIC || h->ry < h->Ry - 1);
    g.fold = 0;
    for (int loc = 0; loc < 4; loc++) { g.fold_t[loc] = g.fold_s[loc] = g.fold_t_host[loc] = g.fold_s_host[loc] = nullptr; g.fold_n[loc] = 0; }
    g.fold_sv = g.fold_se = 1.0;
    if (cfg->topo_y == CSI_FOLDED && !g.conn_n) {
        // the rank that holds the fold: its north side is connected -- to itself, through the copy lists -- not a wall.  The general
        // kernels then run over the widened window of a connected side (se:40-46); what they compute in the north halo rows
        // is overwritten by the fold fill that follows every velocity update, exactly as a distributed fill would overwrite it
        g.conn_n = 1;
        g.fold = 1;
        g.fold_sv = cfg->fold_sign_velocity;
        g.fold_se = cfg->fold_sign_external;
    }
    g.conn_w = h->Rx > 1 && (cfg->topo_x == CSI_PERIODIC || h->rx > 0);
    g.conn_e = h->Rx > 1 && (cfg->topo_x == CSI_PERIODIC || h->rx < h->Rx - 1);
    g.dx = cfg->dx; g.dy = cfg->dy; g.az = cfg->dx * cfg->dy;
    g.mask = nullptr;
    g.mask_host = nullptr;
    g.met = nullptr;
    g.metL = 0;
    g.metW = 0;
    g.met_host = nullptr;
    g.met_fused_why = nullptr;
    g.fff_host = nullptr;
    DParams &p = h->p;
    p.Pstar = cfg->ice_compressive_strength;
    p.C = cfg->ice_compaction_hardening;
    p.em2 = (1 / cfg->yield_curve_eccentricity) * (1 / cfg->yield_curve_eccentricity);  // e^(-2) = inv(e)^2
    p.Dmin = cfg->minimum_plastic_stress;
    p.amin = cfg->min_relaxation_parameter;
    p.amax = cfg->max_relaxation_parameter;
    p.ca = cfg->relaxation_strength;
    p.pform = cfg->pressure_formulation;
    p.cor = cfg->coriolis_kind;
    p.min_mass = cfg->minimum_mass;
    p.min_conc = cfg->minimum_concentration;
    p.rho_i = cfg->ice_density;
    p.f = cfg->coriolis_f;
    p.top_kind = cfg->top_stress_kind;
    p.bot_kind = cfg->bottom_stress_kind;
    p.ttx = cfg->top_tau_x; p.tty = cfg->top_tau_y;
    p.rho_e = cfg->rho_e; p.Cd = cfg->Cd; p.ue_c = cfg->ue_const; p.ve_c = cfg->ve_const;
    p.u_sn_bc = cfg->u_south_north_bc; p.v_we_bc = cfg->v_west_east_bc;
    p.u_sn_val = cfg->u_south_north_value; p.v_we_val = cfg->v_west_east_value;
    p.adv_order = cfg->advection_order;
    p.pad_ = 0;
    p.imm_u = cfg->immersed_drag_u;
    p.imm_v = cfg->immersed_drag_v;
    p.fd_kind = cfg->free_drift_kind;
    p.pad2_ = 0;
    p.top_rho = cfg->top_rho_e;
    p.top_Cd = cfg->top_Cd;
    p.fff = nullptr;
    cudaError_t e;
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess) {
        delete h;
        return cuda_fail(nullptr, e, "cudaEventCreate");
    }
    h->nscratch = 592 * 8;
    if ((e = cudaMalloc(&h->scratch, sizeof(double) * h->nscratch)) != cudaSuccess || (e = cudaMalloc(&h->out_dev, sizeof(double) * 8)) != cudaSuccess) {
        delete h;
        return cuda_fail(nullptr, e, "cudaMalloc(scratch)");
    }
    if (cfg->immersed_mask) {
        const size_t n = (size_t)(cfg->Nx + 2 * cfg->Hx) * (cfg->Ny + 2 * cfg->Hy);
        if ((e = cudaMalloc(&h->mask_dev, n)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaMalloc(mask)"); }
        cudaMemcpy(h->mask_dev, cfg->immersed_mask, n, cudaMemcpyDefault);
        g.mask = h->mask_dev;
        h->mask_host.assign(cfg->immersed_mask, cfg->immersed_mask + n);
        g.mask_host = h->mask_host.data();
    }
    if (cfg->metric_kind == CSI_METRIC_J) {
        const int L = cfg->Ny + 2 * cfg->Hy + 1;
        if ((e = cudaMalloc(&h->met_dev, sizeof(double) * 12 * L)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaMalloc(metrics)"); }
        for (int k = 0; k < 12; k++) cudaMemcpy(h->met_dev + (size_t)k * L, cfg->metrics[k], sizeof(double) * L, cudaMemcpyDefault);
        g.met = h->met_dev;
        g.metL = L;
        h->met_host.resize((size_t)12 * L);
        for (int k = 0; k < 12; k++) std::copy(cfg->metrics[k], cfg->metrics[k] + L, h->met_host.begin() + (size_t)k * L);
        g.met_host = h->met_host.data();
        for (int k = 0; k < 12; k++) h->cfg.metrics[k] = nullptr;  // the caller's arrays are not retained
    }
    if (cfg->metric_kind == CSI_METRIC_IJ) {  // orthogonal curvilinear grid: twelve (Ny+2Hy+1) x (Nx+2Hx+1) arrays, i fastest
        const int L = cfg->Ny + 2 * cfg->Hy + 1, Wd = cfg->Nx + 2 * cfg->Hx + 1;
        const size_t n = (size_t)L * Wd;
        if ((e = cudaMalloc(&h->met_dev, sizeof(double) * 12 * n)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaMalloc(metrics)"); }
        for (int k = 0; k < 12; k++) cudaMemcpy(h->met_dev + (size_t)k * n, cfg->metrics[k], sizeof(double) * n, cudaMemcpyDefault);
        g.met = h->met_dev;
        g.metL = L;
        g.metW = Wd;
        h->met_host.resize((size_t)12 * n);
        for (int k = 0; k < 12; k++) std::copy(cfg->metrics[k], cfg->metrics[k] + n, h->met_host.begin() + (size_t)k * n);
        g.met_host = h->met_host.data();
        for (int k = 0; k < 12; k++) h->cfg.metrics[k] = nullptr;
    }
    g.met_fused_why = fused_metrics_check(g);
    if (cfg->coriolis_kind == CSI_CORIOLIS_SPHERICAL) {
        const int L = cfg->Ny + 2 * cfg->Hy + 1;
        if ((e = cudaMalloc(&h->fff_dev, sizeof(double) * L)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaMalloc(coriolis)"); }
        cudaMemcpy(h->fff_dev, cfg->coriolis_f_ff, sizeof(double) * L, cudaMemcpyDefault);
        p.fff = h->fff_dev;
        h->fff_host.assign(cfg->coriolis_f_ff, cfg->coriolis_f_ff + L);
        g.fff_host = h->fff_host.data();
        h->cfg.coriolis_f_ff = nullptr;
    }
    if (g.fold) {
        for (int loc = 0; loc < 4; loc++) {
            const size_t nb = sizeof(int32_t) * (size_t)cfg->fold_count[loc];
            int32_t *t = nullptr, *s = nullptr;
            if ((e = cudaMalloc(&t, nb)) != cudaSuccess || (e = cudaMalloc(&s, nb)) != cudaSuccess) { delete h; return cuda_fail(nullptr, e, "cudaMalloc(fold lists)"); }
            cudaMemcpy(t, cfg->fold_target[loc], nb, cudaMemcpyDefault);
            cudaMemcpy(s, cfg->fold_source[loc], nb, cudaMemcpyDefault);
            h->fold_dev.push_back(t);
            h->fold_dev.push_back(s);
            h->fold_host[2 * loc].assign(cfg->fold_target[loc], cfg->fold_target[loc] + cfg->fold_count[loc]);
            h->fold_host[2 * loc + 1].assign(cfg->fold_source[loc], cfg->fold_source[loc] + cfg->fold_count[loc]);
            g.fold_t_host[loc] = h->fold_host[2 * loc].data();
            g.fold_s_host[loc] = h->fold_host[2 * loc + 1].data();
            g.fold_t[loc] = t;
            g.fold_s[loc] = s;
            g.fold_n[loc] = cfg->fold_count[loc];
        }
    }
    for (int loc = 0; loc < 4; loc++) h->cfg.fold_target[loc] = h->cfg.fold_source[loc] = nullptr;  // the caller's arrays are not retained
    h->mirror.assign(NFIELDS, nullptr);
    h->mirror_n.assign(NFIELDS, 0);
    *out = h;
    return CSI_OK;
}

int csi_destroy(csi_handle *h)
{
    if (!h) return CSI_ERR_ARG;
    cudaSetDevice(h->cfg.device);
    if (h->fused) fused_destroy(h->fused);
    if (h->comm && nccl().ok) nccl().CommDestroy(h->comm);
    for (double *m : h->mirror) if (m) cudaFree(m);
    for (int32_t *q : h->fold_dev) if (q) cudaFree(q);
    if (h->xbuf) cudaFree(h->xbuf);
    if (h->agree_dev) cudaFree(h->agree_dev);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    if (h->ev_block) cudaEventDestroy(h->ev_block);
    if (h->ev_halo) cudaEventDestroy(h->ev_halo);
    if (h->ev_async_in) cudaEventDestroy(h->ev_async_in);
    if (h->ev_async_done) cudaEventDestroy(h->ev_async_done);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->scratch) cudaFree(h->scratch);
    if (h->out_dev) cudaFree(h->out_dev);
    if (h->mask_dev) cudaFree(h->mask_dev);
    if (h->met_dev) cudaFree(h->met_dev);
    if (h->fff_dev) cudaFree(h->fff_dev);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    delete h;
    return CSI_OK;
}

int csi_evp_substeps(csi_handle *h, const csi_fields *f, double dt_stage, int32_t nsubsteps, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_MOMENTUM, &df);
    if (rc) return rc;
    if (nsubsteps < 0) return fail(h, CSI_ERR_ARG, "nsubsteps must be >= 0");
    cudaStream_t s = (cudaStream_t)stream;
    Timed t(h, s);
    return momentum_impl(h, df, dt_stage, nsubsteps, s);
}

int csi_compute_tracer_tendencies(csi_handle *h, const csi_fields *f, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_TRACERS, &df);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Timed t(h, s);
    LaunchCtx c{s, &h->launches};
    launch_tracer_tendencies(c, h->g, h->p, df);
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

int csi_dynamic_time_step(csi_handle *h, const csi_fields *f, double dt_stage, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_TRACERS | NEED_RK, &df);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Timed t(h, s);
    LaunchCtx c{s, &h->launches};
    const bool rk = h->cfg.timestepper == CSI_RK3;
    launch_dynamic_step(c, h->g, df, rk ? df.hm : df.h, rk ? df.am : df.a, rk ? df.hsm : df.hs, dt_stage);
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

int csi_cache_current_fields(csi_handle *h, const csi_fields *f, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_TRACERS | NEED_RK, &df);
    if (rc) return rc;
    if (h->cfg.timestepper != CSI_RK3) return CSI_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if ((rc = copy_parent(h, df.hm, df.h, s))) return rc;
    if ((rc = copy_parent(h, df.am, df.a, s))) return rc;
    if (df.hs.p && (rc = copy_parent(h, df.hsm, df.hs, s))) return rc;
    if ((rc = copy_parent(h, df.um, df.u, s))) return rc;
    return copy_parent(h, df.vm, df.v, s);
}

int csi_update_state(csi_handle *h, const csi_fields *f, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_TRACERS & ~0, &df);
    if (rc) return rc;
    return update_state_impl(h, df, (cudaStream_t)stream);
}

int csi_thermodynamic_time_step(csi_handle *h, const csi_thermo_config *cfg, const csi_thermo_fields *f, double dt, csi_stream stream)
{
    DThermoFields tf;
    int rc = convert_thermo(h, cfg, f, &tf);
    if (rc) return rc;
    if (!(dt > 0)) return fail(h, CSI_ERR_ARG, "csi_thermodynamic_time_step: dt must be positive");
    cudaStream_t s = (cudaStream_t)stream;
    Timed t(h, s);
    LaunchCtx c{s, &h->launches};
    launch_thermodynamics(c, h->g, *cfg, tf, h->cfg.ice_density, dt);
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

int csi_attach_thermodynamics(csi_handle *h, const csi_thermo_config *cfg, const csi_thermo_fields *f)
{
    if (!h) return CSI_ERR_ARG;
    if (!cfg) {
        h->thermo_on = false;
        return CSI_OK;
    }
    DThermoFields tf;
    int rc = convert_thermo(h, cfg, f, &tf);
    if (rc) return rc;
    h->thermo_cfg = *cfg;
    h->thermo_f = tf;
    h->thermo_on = true;
    return CSI_OK;
}

int csi_fill_halos(csi_handle *h, const csi_array *a, int32_t loc_x, int32_t loc_y, int32_t which, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    if (!a || !a->ptr) return fail(h, CSI_ERR_ARG, "csi_fill_halos: NULL array");
    FieldInfo fi{"array", loc_x ? 1 : 0, loc_y ? 1 : 0};
    int rc = check_array(h, *a, fi, true);
    if (rc) return rc;
    LaunchCtx c{(cudaStream_t)stream, &h->launches};
    launch_fill_halo(c, h->g, h->p, to_darr(*a), fi.lx, fi.ly, which);
    CSI_CUDA(h, cudaGetLastError());
    return CSI_OK;
}

int csi_time_step(csi_handle *h, const csi_fields *f, double dt, int32_t first, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_MOMENTUM | NEED_TRACERS | NEED_RK, &df);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    Timed t(h, s);
    return time_step_impl(h, df, dt, first, s);
}

static int reduce_to_host(csi_handle *h, const csi_fields *f, double *host6, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, 0, &df);
    if (rc) return rc;
    if (!df.u.p || !df.v.p) return fail(h, CSI_ERR_ARG, "u and v are required");
    cudaStream_t s = (cudaStream_t)stream;
    LaunchCtx c{s, &h->launches};
    launch_diagnostics(c, h->g, df, h->scratch, h->nscratch, h->out_dev);
    CSI_CUDA(h, cudaMemcpyAsync(host6, h->out_dev, sizeof(double) * 6, cudaMemcpyDeviceToHost, s));
    CSI_CUDA(h, cudaStreamSynchronize(s));
    return CSI_OK;
}

int csi_cell_advection_timescale(csi_handle *h, const csi_fields *f, double *out_host, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    if (!out_host) return fail(h, CSI_ERR_ARG, "out_host is NULL");
    double tmp[6];
    int rc = reduce_to_host(h, f, tmp, stream);
    if (rc) return rc;
    *out_host = tmp[5];
    return CSI_OK;
}

int csi_diagnostics(csi_handle *h, const csi_fields *f, double *out_host5, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    if (!out_host5) return fail(h, CSI_ERR_ARG, "out_host5 is NULL");
    double tmp[6];
    int rc = reduce_to_host(h, f, tmp, stream);
    if (rc) return rc;
    for (int k = 0; k < 5; k++) out_host5[k] = tmp[k];
    return CSI_OK;
}

// ---- host-buffer entry points ------------------------------------------------------------------
// Mirrors every non-NULL host array on the device; copies only the ones in `inputs` (a bit mask over
// csi_fields member indices): arrays the call merely writes (P, un, vn, zeta, Delta, alpha, G^n ...) are not uploaded.
static int upload_all(csi_handle *h, const csi_fields *hf, csi_fields *dev, cudaStream_t s, uint32_t inputs, size_t *h2d_bytes)
{
    for (int k = 0; k < NFIELDS; k++) {
        const csi_array &a = field_at(*hf, k);
        csi_array &d = field_at(*dev, k);
        d = a;
        if (!a.ptr) continue;
        const size_t n = (size_t)a.nx_tot * a.ny_tot;
        if (h->mirror_n[k] != n) {
            if (h->mirror[k]) cudaFree(h->mirror[k]);
            h->mirror[k] = nullptr;
            CSI_CUDA(h, cudaMalloc(&h->mirror[k], n * sizeof(double)));
            h->mirror_n[k] = n;
        }
        if (inputs & (1u << k)) {
            CSI_CUDA(h, cudaMemcpyAsync(h->mirror[k], a.ptr, n * sizeof(double), cudaMemcpyHostToDevice, s));
            if (h2d_bytes) *h2d_bytes += n * sizeof(double);
        }
        d.ptr = h->mirror[k];
    }
    return CSI_OK;
}
static int download(csi_handle *h, const csi_fields *hf, const int *which, int n, cudaStream_t s)
{
    for (int q = 0; q < n; q++) {
        const int k = which[q];
        const csi_array &a = field_at(*hf, k);
        if (!a.ptr) continue;
        CSI_CUDA(h, cudaMemcpyAsync(a.ptr, h->mirror[k], h->mirror_n[k] * sizeof(double), cudaMemcpyDeviceToHost, s));
        h->last_d2h += h->mirror_n[k] * sizeof(double);
    }
    return CSI_OK;
}
// An array the call rewrites only on the interior and its first ring (alpha: a diagnostic without halo fill): that window is
// copied back and the rest of the host array is left as it is, so the array need not be uploaded to preserve its halo.
static int download_window(csi_handle *h, const csi_fields *hf, int k, cudaStream_t s)
{
    const csi_array &a = field_at(*hf, k);
    if (!a.ptr) return CSI_OK;
    const int i0 = std::max(a.off_x - 1, 0), i1 = std::min(a.off_x + h->cfg.Nx + 1, a.nx_tot);  // columns [i0, i1)
    const int j0 = std::max(a.off_y - 1, 0), j1 = std::min(a.off_y + h->cfg.Ny + 1, a.ny_tot);
    const size_t pitch = (size_t)a.nx_tot * sizeof(double), off = (size_t)j0 * a.nx_tot + i0;
    CSI_CUDA(h, cudaMemcpy2DAsync(a.ptr + off, pitch, h->mirror[k] + off, pitch, (size_t)(i1 - i0) * sizeof(double), (size_t)(j1 - j0), cudaMemcpyDeviceToHost, s));
    h->last_d2h += (size_t)(i1 - i0) * (j1 - j0) * sizeof(double);
    return CSI_OK;
}

int csi_time_step_host(csi_handle *h, const csi_fields *hf, double dt, int32_t nsteps, int32_t first)
{
    if (!h) return CSI_ERR_ARG;
    if (!hf) return fail(h, CSI_ERR_ARG, "csi_fields pointer is NULL");
    cudaSetDevice(h->cfg.device);
    if (!h->own_stream) CSI_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->own_stream;
    csi_fields dev;
    // inputs of time_step!: u v h a s11 s22 s12 (0-6), top_x top_y ue ve (14-17); Psi^- is written before it is read
    // + alpha (10): only its interior is rewritten, the halo must survive the round trip; + hs (24), fd_u, fd_v (27, 28) when present
    const uint32_t in_mask = 0x7fu | (1u << 10) | (0xfu << 14) | (1u << K_HS) | (3u << K_FDU);
    h->last_h2d = h->last_d2h = 0;
    int rc = upload_all(h, hf, &dev, s, in_mask, &h->last_h2d);
    if (rc) return rc;
    DFields df;
    if ((rc = convert_fields(h, &dev, NEED_MOMENTUM | NEED_TRACERS | NEED_RK, &df))) return rc;
    {
        Timed t(h, s);
        for (int k = 0; k < nsteps; k++)
            if ((rc = time_step_impl(h, df, dt, first && k == 0, s))) return rc;
    }
    const int outs[] = {0, 1, 2, 3, 4, 5, 6, 10, K_HS};  // u v h a s11 s22 s12 alpha (hs)
    if ((rc = download(h, hf, outs, 9, s))) return rc;
    CSI_CUDA(h, cudaStreamSynchronize(s));
    return CSI_OK;
}

int csi_evp_substeps_host(csi_handle *h, const csi_fields *hf, double dt_stage, int32_t nsubsteps)
{
    if (!h) return CSI_ERR_ARG;
    if (!hf) return fail(h, CSI_ERR_ARG, "csi_fields pointer is NULL");
    cudaSetDevice(h->cfg.device);
    if (!h->own_stream) CSI_CUDA(h, cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->own_stream;
    csi_fields dev;
    // inputs of time_step_momentum!: u v h a s11 s22 s12 (0-6), top_x top_y ue ve (14-17), the free-drift arrays, and Psi^-.u, .v
    // (22, 23) under RK3 -- where reset_velocities! (se:87-93) overwrites u, v with them before anything reads u, v, so those two
    // are not uploaded.  alpha is written, not read: only the window the call rewrites comes back (download_window).
    const bool reset = h->cfg.timestepper == CSI_RK3 && field_at(*hf, 22).ptr && field_at(*hf, 23).ptr;
    const uint32_t in_mask = (reset ? 0x7cu : 0x7fu) | (0xfu << 14) | (3u << K_FDU) | (reset ? (0x3u << 22) : 0u);
    h->last_h2d = h->last_d2h = 0;
    int rc = upload_all(h, hf, &dev, s, in_mask, &h->last_h2d);
    if (rc) return rc;
    DFields df;
    if ((rc = convert_fields(h, &dev, NEED_MOMENTUM, &df))) return rc;
    {
        Timed t(h, s);
        if ((rc = momentum_impl(h, df, dt_stage, nsubsteps, s))) return rc;
    }
    const int outs[] = {0, 1, 4, 5, 6};  // u v s11 s22 s12; alpha: the window the call rewrites
    if ((rc = download(h, hf, outs, 5, s))) return rc;
    if ((rc = download_window(h, hf, 10, s))) return rc;
    CSI_CUDA(h, cudaStreamSynchronize(s));
    return CSI_OK;
}

// ---- multi-GPU ------------------------------------------------------------------------------
int csi_nccl_unique_id(uint8_t out128[128])
{
    if (!out128) return CSI_ERR_ARG;
    NcclApi &api = nccl();
    if (!api.ok) return fail(nullptr, CSI_ERR_NCCL_MISSING, "libnccl.so.2 could not be loaded");
    nccl_uid id;
    int rc = api.GetUniqueId(&id);
    if (rc) return fail(nullptr, 1000 + rc, "ncclGetUniqueId failed");
    memcpy(out128, id.internal, 128);
    return CSI_OK;
}

int csi_comm_init(csi_handle *h, const uint8_t id128[128], int32_t rank, int32_t nranks)
{
    if (!h) return CSI_ERR_ARG;
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, CSI_ERR_ARG, "csi_comm_init: bad arguments");
    if (nranks != h->nranks || rank != h->rank) return fail(h, CSI_ERR_ARG, "csi_comm_init: rank/nranks differ from csi_config");
    NcclApi &api = nccl();
    if (!api.ok) return fail(h, CSI_ERR_NCCL_MISSING, "libnccl.so.2 could not be loaded");
    cudaSetDevice(h->cfg.device);
    nccl_uid id;
    memcpy(id.internal, id128, 128);
    int rc = api.CommInitRank(&h->comm, nranks, id, rank);
    if (rc) return fail(h, 1000 + rc, std::string("ncclCommInitRank: ") + (api.GetErrorString ? api.GetErrorString(rc) : "error"));
    return CSI_OK;
}

int csi_exchange_halos(csi_handle *h, const csi_array *arrays, int32_t narrays, int32_t width, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    if (!arrays || narrays < 1 || narrays > 16) return fail(h, CSI_ERR_ARG, "csi_exchange_halos: bad arguments");
    DArr d[16];
    for (int k = 0; k < narrays; k++) d[k] = to_darr(arrays[k]);
    return exchange_slab_halos(h, d, narrays, width, (cudaStream_t)stream);
}

int csi_exchange_halos_async(csi_handle *h, const csi_array *arrays, int32_t narrays, int32_t width, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    if (!arrays || narrays < 1 || narrays > 16) return fail(h, CSI_ERR_ARG, "csi_exchange_halos_async: bad arguments");
    if (h->halo_pending) return fail(h, CSI_ERR_ARG, "csi_exchange_halos_async: an asynchronous exchange is already in flight (csi_wait_halos first)");
    DArr d[16];
    for (int k = 0; k < narrays; k++) d[k] = to_darr(arrays[k]);
    return exchange_halos_async(h, d, narrays, width, (cudaStream_t)stream);
}

int csi_wait_halos(csi_handle *h, csi_stream stream)
{
    if (!h) return CSI_ERR_ARG;
    return wait_halos(h, (cudaStream_t)stream);
}

int64_t csi_launch_count(const csi_handle *h) { return h ? h->launches : 0; }

int csi_fused_stats(const csi_handle *h, int64_t out3[3])
{
    if (!h || !out3) return CSI_ERR_ARG;
    out3[0] = out3[1] = out3[2] = 0;
    if (!h->fused) return CSI_OK;
    long long v[3];
    cudaSetDevice(h->cfg.device);
    fused_stats(h->fused, v);
    for (int k = 0; k < 3; k++) out3[k] = v[k];
    return CSI_OK;
}

double csi_last_elapsed_ms(const csi_handle *h)
{
    if (!h || !h->timed) return 0.0;
    float ms = 0.f;
    cudaEventSynchronize(h->ev1);
    if (cudaEventElapsedTime(&ms, h->ev0, h->ev1) != cudaSuccess) return 0.0;
    return (double)ms;
}

int csi_time_dominant_kernel(csi_handle *h, const csi_fields *f, double dt_stage, int32_t reps, double *out_ms, char *name64,
                             int32_t *bytes_per_cell, csi_stream stream)
{
    DFields df;
    int rc = convert_fields(h, f, NEED_MOMENTUM, &df);
    if (rc) return rc;
    if (!out_ms || !name64 || !bytes_per_cell || reps < 1) return fail(h, CSI_ERR_ARG, "csi_time_dominant_kernel: bad arguments");
    cudaStream_t s = (cudaStream_t)stream;
    LaunchCtx c{s, &h->launches};
    bool use_fused = false;
    if (h->cfg.solver_impl != CSI_SOLVER_UNFUSED && !h->fused_failed) {
        char why[256];
        use_fused = fused_supported(h->g, h->p, df, why, sizeof why) != 0;
    }
    char err[256] = {0};
    if (use_fused && !h->fused) {
        h->fused = fused_create(h->g, h->p, err, sizeof err);
        if (!h->fused) return fail(h, 1, std::string("fused_create: ") + err);
    }
    // one untimed launch first (module load, tensor maps); the timed region holds kernel launches only
    if (use_fused) {
        rc = fused_begin(h->fused, c, h->g, h->p, df, dt_stage, err, sizeof err);
        if (!rc) rc = fused_steps(h->fused, c, 1, 2, false, err, sizeof err, nullptr);
    } else {
        launch_evp_stress(c, h->g, h->p, df, dt_stage);
    }
    if (rc) return fail(h, rc, err);
    CSI_CUDA(h, cudaEventRecord(h->ev0, s));
    if (use_fused) rc = fused_steps(h->fused, c, 1, reps, false, err, sizeof err, nullptr);
    else for (int k = 0; k < reps; k++) launch_evp_stress(c, h->g, h->p, df, dt_stage);
    if (rc) return fail(h, rc, err);
    CSI_CUDA(h, cudaEventRecord(h->ev1, s));
    CSI_CUDA(h, cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    CSI_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (use_fused && (rc = fused_end(h->fused, c, df, err, sizeof err))) return fail(h, rc, err);
    *out_ms = (double)ms / reps;
    snprintf(name64, 64, "%s", use_fused ? "k_evp_substep_fused" : "k_evp_stress");
    *bytes_per_cell = use_fused ? 144 : 120;
    return CSI_OK;
}

int csi_selftest_math(int64_t samples, uint64_t seed, int32_t exponent_span, uint64_t *out5)
{
    if (!out5 || samples < 1 || exponent_span < 0 || exponent_span > 1000) return CSI_ERR_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, CSI_ERR_NO_DEVICE, "csi_selftest_math: no CUDA device");
    }
    unsigned long long tmp[5];
    int rc = csi::fz::selftest_math(samples, seed, exponent_span, tmp);
    for (int k = 0; k < 5; k++) out5[k] = tmp[k];
    return rc;
}

int csi_last_transfer_bytes(const csi_handle *h, uint64_t *h2d, uint64_t *d2h)
{
    if (!h || !h2d || !d2h) return CSI_ERR_ARG;
    *h2d = h->last_h2d;
    *d2h = h->last_d2h;
    return CSI_OK;
}

double csi_host_exp(double x) { return csi::exp_cr(x); }
double csi_host_div_by_const(double x, double c) { return csi::div_by(x, csi::make_recip(c)); }
int csi_host_halo_width(int32_t k) { return 2 * k + 3; }

}  // extern "C"
