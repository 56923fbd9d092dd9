// csi_fused.cu -- one kernel launch per EVP substep (sm_100a).
//
// Replaces, per substep, the reference's four kernels + two local halo fills
// (compute_stresses! evp:222-354, _u/_v_velocity_step! se:197-264, fill_halo_regions! se:170-187)
// with a single streaming kernel:
//
//   * Fields live in an internal planar layout owned by the plan: one allocation
//     [field][row][pitch], pitch a multiple of 16 doubles, interior column 1 at a 128-byte
//     boundary, a halo ring of W cells.  One 3-D TMA tensor map describes all of it.
//   * A CTA owns a strip of columns and marches along y.  Warp 0 is the TMA producer: each step it
//     issues one `cp.async.bulk.tensor` row load (128 columns) per input field into a shared-memory
//     stage, completion on a `full` mbarrier; consumer warps release stages through `empty`
//     mbarriers.  There is no block-wide barrier anywhere in the loop.
//   * Each consumer warp owns 32*NC columns (NC columns per lane) and is autonomous: row history
//     (strain rates, P, m, alpha, new stresses, first velocity of the previous rows) lives in
//     register shift-registers, x-neighbours come from warp shuffles.  Per step a warp computes
//     strain rates (row t, t+1), the stress update (row t), the first velocity (row t or t-1) and the
//     second velocity (row t-1).  Warps overlap by 4 columns (2 per side) and recompute them.
//   * The five evolving fields are double buffered in HBM (read set A, write set B), so there is
//     no hazard between CTAs; the owner of a cell also stores its periodic images / wall values,
//     which replaces the two halo-fill launches per substep.
//   * Arithmetic keeps the reference's Float64 expression trees (compiled with -fmad=false).
//     Divisions by the constant metrics and by divisors used several times (m_i, alpha_bar,
//     gamma) go through the Markstein quotient of csi_math.cuh, which returns the IEEE quotient.
//
// Algorithmic HBM traffic: 14 loads + 5 stores per cell-update (u, v, s11, s22, s12 r/w; h, aice,
// P, un, vn, tau_x, tau_y, ue, ve read) = 152 B, 144 B by the SURVEY convention (P recomputable).
#include <cuda.h>
#include <stdio.h>
#include <string.h>

#include "csi_cell.cuh"
#include "csi_internal.h"

namespace csi {

namespace fz {

constexpr int BOX = 128;         // columns per TMA row box = columns per CTA strip (incl. overlap)
constexpr int W = 3;             // halo ring kept valid in the internal layout
constexpr int OX = 16;           // internal column of i = 1 (128-byte aligned)
constexpr int NSTAGE = 3;        // TMA stages in flight per CTA
constexpr int NIN = 14;          // input rows per stage

// internal field indices
enum { F_U0 = 0, F_V0, F_S11_0, F_S22_0, F_S12_0, F_U1, F_V1, F_S11_1, F_S22_1, F_S12_1,
       F_H, F_A, F_P, F_UN, F_VN, F_TX, F_TY, F_UE, F_VE, F_ALPHA, F_ZC, F_ZF, F_DELTA, NF };
// stage slots
enum { I_U = 0, I_V, I_H, I_A, I_P, I_S11, I_S22, I_S12, I_UN, I_VN, I_TX, I_TY, I_UE, I_VE };

constexpr size_t SMEM_BYTES = (size_t)NSTAGE * NIN * BOX * sizeof(double) + 2 * NSTAGE * sizeof(uint64_t) + 128;

// geometry of one kernel variant: NC columns per lane
template <int NC> struct Geo {
    static constexpr int WCOLS = 32 * NC;              // columns per consumer warp
    static constexpr int WOUT = WCOLS - 4;             // output columns per warp (2 overlap per side)
    static constexpr int NCW = (BOX - 4) / WOUT;       // consumer warps per CTA
    static constexpr int OUTX = NCW * WOUT;            // output columns per CTA
    static constexpr int THREADS = (NCW + 1) * 32;     // + the producer warp
};

struct Params {
    int Nx, Ny;          // interior size
    int pitch, rows;     // internal layout
    int oy;              // internal row of j = 1 is (oy)
    int px, py;          // periodic images along x / y
    int bounded_x, bounded_y;
    // store windows (reference indices, inclusive)
    int sx0, sx1, sy0, sy1;  // stresses
    int vx0, vx1, vy0, vy1;  // velocities
    int cx0, cx1, cy0, cy1;  // cells whose velocity is evolved (periodic images included); others keep their value
    int LY;                  // output rows per CTA
    int a0;                  // first column of strip 0 (chosen so every TMA box starts 16-byte aligned)
    int use_top, use_ue;     // field arrays present
    int u_sn_bc, v_we_bc;
    double u_sn_val, v_we_val;
    double dt;
    double dx, dy, az, dx2, dy2, rdx, rdy, raz;
    double em2, Dmin, amin, amax, amax2, ca, rho_i, rhoCd, f, min_mass, min_conc;
    double ttx, tty, ue_c, ve_c;
    int pform, cor, sis;
    int in_set, out_set;  // 0 / 1: which copy of the evolving fields is read / written
    double *base;         // internal allocation
};


// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_row(double *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- bit-exact arithmetic policies -------------------------------------------------------------
// FAST: branch-free, correctly rounded division / reciprocal / square root built from FMAs:
//   rcp:  r0 = rcp.approx(y); two Newton steps (faithful); one Markstein step  -> RN(1/y)
//   x/y:  q0 = RN(x r), t = y q0 - x (exact, FMA), q = RN(q0 - t r)           -> RN(x/y)
//         (Markstein; exact for every x when r = RN(1/y) and y's significand is not all ones,
//          Brisebarre, Muller, Raina 2004; written so that +-0 / y keeps its sign)
//   sqrt: y0 = rsqrt.approx(x); g = x y0, h = y0/2; two coupled Newton steps; g + h (x - g g)  -> RN(sqrt x)
// These hold only while nothing over/underflows, so every divisor, quotient and radicand is folded
// into integer min/max accumulators of its exponent field (4-5 integer ops, no branch); after the
// step the warp checks the windows once and, if any lane left them (zero ice mass, NaN, denormals,
// a significand of all ones ...), recomputes the step with the SLOW policy: plain IEEE operators.
// Both policies therefore return the IEEE results; tests/ compares them bit for bit on the GPU.
struct NodeRecip {
    double d, r;
};

struct MathFast {
    // exponent field << 21 of: quotients / radicands (zero allowed) and divisors (zero not allowed)
    uint32_t qmn = 0xffffffffu, qmx = 0u, dmn = 0xffffffffu, dmx = 0u, lo1 = 0xffffffffu, neg = 0u;
    static constexpr uint32_t QLO = 0x200u << 21, QHI = (0x600u << 21) - 1u;  // |q| in [2^-511, 2^513)
    static constexpr uint32_t DLO = 0x300u << 21, DHI = (0x500u << 21) - 1u;  // |d| in [2^-255, 2^257)
    __device__ __forceinline__ void chkq(double q)
    {
        const uint32_t g = (uint32_t)__double2hiint(q) << 1;
        qmx = max(qmx, g);
        qmn = min(qmn, g - 1u);  // g == 0 (a zero) wraps to 0xffffffff and is ignored
    }
    __device__ __forceinline__ void chkd(double d)
    {
        const uint32_t g = (uint32_t)__double2hiint(d) << 1;
        dmx = max(dmx, g);
        dmn = min(dmn, g);
        lo1 = min(lo1, (uint32_t)__double2loint(d) + 1u);  // 0 if the low word is all ones (superset of "significand all ones")
    }
    __device__ __forceinline__ bool bad() const { return (qmn < QLO - 1u) | (qmx > QHI) | (dmn < DLO) | (dmx > DHI) | (lo1 == 0u) | ((neg >> 31) != 0u); }

    __device__ __forceinline__ double rcp(double y)
    {
        chkd(y);
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
        double e = __fma_rn(-y, r, 1.0);
        r = __fma_rn(r, e, r);
        e = __fma_rn(-y, r, 1.0);
        r = __fma_rn(r, e, r);
        e = __fma_rn(-y, r, 1.0);
        return __fma_rn(r, e, r);
    }
    __device__ __forceinline__ double quot(double x, double d, double r)
    {
        const double q0 = x * r;
        chkq(q0);
        const double t = __fma_rn(d, q0, -x);
        return __fma_rn(-t, r, q0);
    }
    __device__ __forceinline__ double divc(double x, double d, double r) { return quot(x, d, r); }
    __device__ __forceinline__ NodeRecip recip(double d)
    {
        NodeRecip R;
        R.d = d;
        R.r = rcp(d);
        return R;
    }
    __device__ __forceinline__ double divn(double x, const NodeRecip &R) { return quot(x, R.d, R.r); }
    __device__ __forceinline__ double div(double x, double y) { return quot(x, y, rcp(y)); }
    __device__ __forceinline__ double sqrt_(double x)
    {
        chkq(x);
        neg |= (uint32_t)__double2hiint(x);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        double g = x * y, h = 0.5 * y;
        double e = __fma_rn(-h, g, 0.5);
        g = __fma_rn(g, e, g);
        h = __fma_rn(h, e, h);
        e = __fma_rn(-h, g, 0.5);
        g = __fma_rn(g, e, g);
        h = __fma_rn(h, e, h);
        const double d = __fma_rn(-g, g, x);
        g = __fma_rn(d, h, g);
        return x > 0.0 ? g : x;  // sqrt(+-0) = +-0
    }
};
struct MathSlow {
    __device__ __forceinline__ bool bad() const { return false; }
    __device__ __forceinline__ double divc(double x, double d, double) { return x / d; }
    __device__ __forceinline__ NodeRecip recip(double d)
    {
        NodeRecip R;
        R.d = d;
        R.r = 0.0;
        return R;
    }
    __device__ __forceinline__ double divn(double x, const NodeRecip &R) { return x / R.d; }
    __device__ __forceinline__ double div(double x, double y) { return x / y; }
    __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
};

// ---- per-lane state ---------------------------------------------------------------------------
// rows of this step taken from the TMA stage
template <int NC> struct In {
    double u_n[NC], v_n[NC], h0[NC], a0[NC], P0[NC], o11[NC], o22[NC], o12[NC], un_[NC], vn_[NC], tx_[NC], ty_[NC], ue0[NC], ve0[NC];
};
// row history kept in registers (own columns)
template <int NC> struct Hist {
    double u_c[NC], u_p[NC], v_c[NC], v_p[NC], udx_c[NC], vdx_c[NC];
    double e11_p[NC], e22_p[NC], e12_c[NC];
    double P_p[NC], m_p[NC], m_pp[NC], a_p[NC], a_pp[NC], al_p[NC], al_pp[NC];
    double s11_p[NC], s22_p[NC], s12_p[NC], s11_pp[NC], s22_pp[NC];
    double w_p[NC], ue_p[NC], ue_pp[NC], ve_p[NC];
};
// everything one step produces
template <int NC> struct Out {
    double e11_c[NC], e22_c[NC], e12_n[NC], m_c[NC], undx[NC], vndx[NC];
    double n11[NC], n22[NC], n12[NC], gc[NC], zc[NC], zf[NC], Dc[NC];
    double w1[NC], w2[NC];  // first / second velocity of the step
};

template <int NC> __device__ __forceinline__ void lefts(const double (&x)[NC], double (&o)[NC])
{
    const double s = __shfl_up_sync(0xffffffffu, x[NC - 1], 1);
#pragma unroll
    for (int k = 0; k < NC; k++) o[k] = k > 0 ? x[k - 1] : s;
}
template <int NC> __device__ __forceinline__ void rights(const double (&x)[NC], double (&o)[NC])
{
    const double s = __shfl_down_sync(0xffffffffu, x[0], 1);
#pragma unroll
    for (int k = 0; k < NC; k++) o[k] = k < NC - 1 ? x[k + 1] : s;
}

// u at (i, r): se:197-229, mt:11-41, ext:176-196, isd:39-44, evp:384,391-395.  *0 = column i-1.
template <class M>
__device__ __forceinline__ double u_node(M &mm, const Params &p, bool active, double m1, double m0, double a1, double a0_, double al1, double al0,
                                         double uold, double vbar, double ue, double vebar, double ttop, double un, double sD1, double sD0,
                                         double sT1, double sT0, double s12hi, double s12lo)
{
    const double mi = (m1 + m0) / 2, ai = (a1 + a0_) / 2, abar = (al1 + al0) / 2;
    const NodeRecip Ra = mm.recip(abar), Rm = mm.recip(mi);
    const double dtau = mm.divn(p.dt, Ra);
    double coef = 0.0, tbot = 0.0;
    if (p.sis) {
        const double du = ue - uold, dv = vebar - vbar;
        coef = p.rhoCd * mm.sqrt_(du * du + dv * dv);
        tbot = coef * ue;
    }
    const double xcross = p.cor == CSI_CORIOLIS_NONE ? 0.0 : -p.f * vbar;
    const double rheo = mm.divn(mm.div(un - uold, dtau), Ra);
    const double d = p.dy * (sD1 - sD0) / 2;
    const double tt = mm.divc(p.dy2 * sT1 - p.dy2 * sT0, p.dy, p.rdy) / 2;
    const double SS = mm.divc(p.dx2 * s12hi - p.dx2 * s12lo, p.dx, p.rdx);
    const double dsig = mm.divc(d + tt + SS, p.az, p.raz);
    double G = -xcross - mm.divn(ttop, Rm) * ai + mm.divn(tbot, Rm) * ai + mm.divn(dsig, Rm) + 0.0 + (0.0 + rheo);
    G = mi <= 0 ? 0.0 : G;
    double tau = mm.divn(coef - 0.0, Rm) * ai;
    tau = mi <= 0 ? 0.0 : tau;
    const double uD = mm.div(uold + dtau * G, 1 + dtau * tau);
    const bool active_ice = (mi >= p.min_mass) & (ai >= p.min_conc);
    return jl_mul_bool(active_ice ? uD : 0.0, active);  // free_drift = nothing: marginal ice -> 0
}
// v at (i, r): se:231-264, mt:44-74, ext:183-202, isd:46-51, evp:385,397-401.  *0 = row r-1.
template <class M>
__device__ __forceinline__ double v_node(M &mm, const Params &p, bool active, double m1, double m0, double a1, double a0_, double al1, double al0,
                                         double vold, double ubar, double ve, double uebar, double ttop, double vn, double sD1, double sD0,
                                         double sT1, double sT0, double s12hi, double s12lo)
{
    const double mi = (m1 + m0) / 2, ai = (a1 + a0_) / 2, abar = (al1 + al0) / 2;
    const NodeRecip Ra = mm.recip(abar), Rm = mm.recip(mi);
    const double dtau = mm.divn(p.dt, Ra);
    double coef = 0.0, tbot = 0.0;
    if (p.sis) {
        const double dv = ve - vold, du = uebar - ubar;
        coef = p.rhoCd * mm.sqrt_(du * du + dv * dv);
        tbot = coef * ve;
    }
    const double ycross = p.cor == CSI_CORIOLIS_NONE ? 0.0 : p.f * ubar;
    const double rheo = mm.divn(mm.div(vn - vold, dtau), Ra);
    const double d = p.dx * (sD1 - sD0) / 2;
    const double tt = mm.divc(-(p.dx2 * sT1 - p.dx2 * sT0), p.dx, p.rdx) / 2;
    const double SS = mm.divc(p.dy2 * s12hi - p.dy2 * s12lo, p.dy, p.rdy);
    const double dsig = mm.divc(d + tt + SS, p.az, p.raz);
    double G = -ycross - mm.divn(ttop, Rm) * ai + mm.divn(tbot, Rm) * ai + mm.divn(dsig, Rm) + 0.0 + (0.0 + rheo);
    G = mi <= 0 ? 0.0 : G;
    double tau = mm.divn(coef - 0.0, Rm) * ai;
    tau = mi <= 0 ? 0.0 : tau;
    const double vD = mm.div(vold + dtau * G, 1 + dtau * tau);
    const bool active_ice = (mi >= p.min_mass) & (ai >= p.min_conc);
    return jl_mul_bool(active_ice ? vD : 0.0, active);
}

// One marching step of one lane: strain rates (rows t, t+1), stress update (row t), first and
// second velocity.  Pure function of (history, this step's rows); all lanes of the warp call it
// together (shuffles inside).  updC / updD: whether the first / second velocity cell is evolved.
template <int NC, bool VFIRST, class M>
__device__ __forceinline__ void compute_step(M &mm, const Params &p, const Hist<NC> &h, const In<NC> &in, Out<NC> &o, const bool (&actu)[NC],
                                             bool actv1, bool actv2, const bool (&updC)[NC], const bool (&updD)[NC])
{
    // ---------------- phase A: strain rates (evp:360-375), ice mass (ClimaSeaIce.jl:42) --------
    {
        double u_r[NC], v_l[NC];
        rights<NC>(h.u_c, u_r);
        lefts<NC>(in.v_n, v_l);
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const double D = mm.divc((p.dy * u_r[k] - p.dy * h.u_c[k]) + (p.dx * in.v_n[k] - p.dx * h.v_c[k]), p.az, p.raz);
            o.vndx[k] = mm.divc(in.v_n[k], p.dx, p.rdx);
            const double T = mm.divc(p.dy2 * (mm.divc(u_r[k], p.dy, p.rdy) - mm.divc(h.u_c[k], p.dy, p.rdy)) - p.dx2 * (o.vndx[k] - h.vdx_c[k]), p.az, p.raz);
            o.undx[k] = mm.divc(in.u_n[k], p.dx, p.rdx);
            const double S = mm.divc(p.dx2 * (o.undx[k] - h.udx_c[k]) + p.dy2 * (mm.divc(in.v_n[k], p.dy, p.rdy) - mm.divc(v_l[k], p.dy, p.rdy)), p.az, p.raz);
            o.e11_c[k] = (D + T) / 2;
            o.e22_c[k] = (D - T) / 2;
            o.e12_n[k] = S / 2;
            o.m_c[k] = in.h0[k] * p.rho_i * in.a0[k];
        }
    }
    // ---------------- phase B: viscosities + stress update (evp:236-354) at row t ---------------
    {
        double e12c_r[NC], e12n_r[NC], e11p_l[NC], e11c_l[NC], e22p_l[NC], e22c_l[NC], Pp_l[NC], P0_l[NC], mp_l[NC], mc_l[NC];
        rights<NC>(h.e12_c, e12c_r);
        rights<NC>(o.e12_n, e12n_r);
        lefts<NC>(h.e11_p, e11p_l);
        lefts<NC>(o.e11_c, e11c_l);
        lefts<NC>(h.e22_p, e22p_l);
        lefts<NC>(o.e22_c, e22c_l);
        lefts<NC>(h.P_p, Pp_l);
        lefts<NC>(in.P0, P0_l);
        lefts<NC>(h.m_p, mp_l);
        lefts<NC>(o.m_c, mc_l);
#pragma unroll
        for (int k = 0; k < NC; k++) {
            const double e11c = o.e11_c[k], e22c = o.e22_c[k], e12f = h.e12_c[k];
            const double e12c = ((h.e12_c[k] + e12c_r[k]) / 2 + (o.e12_n[k] + e12n_r[k]) / 2) / 2;
            const double e11f = ((e11p_l[k] + h.e11_p[k]) / 2 + (e11c_l[k] + o.e11_c[k]) / 2) / 2;
            const double e22f = ((e22p_l[k] + h.e22_p[k]) / 2 + (e22c_l[k] + o.e22_c[k]) / 2) / 2;
            const double dc = e11c + e22c, df = e11f + e22f;
            const double sc = mm.sqrt_((e11c - e22c) * (e11c - e22c) + 4 * (e12c * e12c));
            const double sf = mm.sqrt_((e11f - e22f) * (e11f - e22f) + 4 * (e12f * e12f));
            const double Dc = jl_max(mm.sqrt_(dc * dc + sc * sc * p.em2), p.Dmin);
            const double Df = jl_max(mm.sqrt_(df * df + sf * sf * p.em2), p.Dmin);
            const double Pc = in.P0[k];
            const double Pf = ((Pp_l[k] + h.P_p[k]) / 2 + (P0_l[k] + in.P0[k]) / 2) / 2;
            const double zf = mm.div(Pf, 2 * Df), zc = mm.div(Pc, 2 * Dc);
            const double Pr = p.pform == CSI_ICE_STRENGTH ? Pc : mm.div(Pc * Dc, Dc + p.Dmin);
            const double ec = zc * p.em2, ef = zf * p.em2;
            const double s11n = 2 * ec * e11c + ((zc - ec) * (e11c + e22c) - Pr / 2);
            const double s22n = 2 * ec * e22c + ((zc - ec) * (e11c + e22c) - Pr / 2);
            const double s12n = 2 * ef * e12f;
            const double mc = o.m_c[k];
            const double mf = ((mp_l[k] + h.m_p[k]) / 2 + (mc_l[k] + o.m_c[k]) / 2) / 2;
            double g2c = mm.divc(mm.div(zc * p.ca * p.dt, mc), p.az, p.raz);
            g2c = (g2c != g2c) ? p.amax2 : g2c;
            const double gc = jl_clamp(mm.sqrt_(g2c), p.amin, p.amax);
            double g2f = mm.divc(mm.div(zf * p.ca * p.dt, mf), p.az, p.raz);
            g2f = (g2f != g2f) ? p.amax2 : g2f;
            const double gf = jl_clamp(mm.sqrt_(g2f), p.amin, p.amax);
            const NodeRecip Rg = mm.recip(gc);
            const double d11 = mm.divn(s11n - in.o11[k], Rg), d22 = mm.divn(s22n - in.o22[k], Rg), d12 = mm.div(s12n - in.o12[k], gf);
            o.n11[k] = in.o11[k] + (mc > 0 ? d11 : 0.0);
            o.n22[k] = in.o22[k] + (mc > 0 ? d22 : 0.0);
            o.n12[k] = in.o12[k] + (mf > 0 ? d12 : 0.0);
            o.gc[k] = gc;
            o.zc[k] = zc;
            o.zf[k] = zf;
            o.Dc[k] = Dc;
        }
    }
    // ---------------- velocity updates -------------------------------------------------------------
    if (VFIRST) {
        {  // C: v at row t (reads old u rows t-1, t)
            double up_r[NC], uc_r[NC], uep_r[NC], ue0_r[NC], n12_r[NC];
            rights<NC>(h.u_p, up_r);
            rights<NC>(h.u_c, uc_r);
            rights<NC>(h.ue_p, uep_r);
            rights<NC>(in.ue0, ue0_r);
            rights<NC>(o.n12, n12_r);
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const double ubar = ((h.u_p[k] + up_r[k]) / 2 + (h.u_c[k] + uc_r[k]) / 2) / 2;
                const double uebar = ((h.ue_p[k] + uep_r[k]) / 2 + (in.ue0[k] + ue0_r[k]) / 2) / 2;
                const double val = v_node(mm, p, actv1, o.m_c[k], h.m_p[k], in.a0[k], h.a_p[k], o.gc[k], h.al_p[k], h.v_c[k], ubar, in.ve0[k], uebar,
                                          in.ty_[k], in.vn_[k], o.n11[k] + o.n22[k], h.s11_p[k] + h.s22_p[k], o.n11[k] - o.n22[k],
                                          h.s11_p[k] - h.s22_p[k], n12_r[k], o.n12[k]);
                o.w1[k] = updC[k] ? val : h.v_c[k];
            }
        }
        {  // D: u at row t-1 (reads new v rows t-1, t)
            double mp_l[NC], ap_l[NC], alp_l[NC], wp_l[NC], wn_l[NC], vep_l[NC], ve0_l[NC], s11p_l[NC], s22p_l[NC];
            lefts<NC>(h.m_p, mp_l);
            lefts<NC>(h.a_p, ap_l);
            lefts<NC>(h.al_p, alp_l);
            lefts<NC>(h.w_p, wp_l);
            lefts<NC>(o.w1, wn_l);
            lefts<NC>(h.ve_p, vep_l);
            lefts<NC>(in.ve0, ve0_l);
            lefts<NC>(h.s11_p, s11p_l);
            lefts<NC>(h.s22_p, s22p_l);
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const double vbar = ((wp_l[k] + h.w_p[k]) / 2 + (wn_l[k] + o.w1[k]) / 2) / 2;
                const double vebar = ((vep_l[k] + h.ve_p[k]) / 2 + (ve0_l[k] + in.ve0[k]) / 2) / 2;
                const double val = u_node(mm, p, actu[k], h.m_p[k], mp_l[k], h.a_p[k], ap_l[k], h.al_p[k], alp_l[k], h.u_p[k], vbar, h.ue_p[k], vebar,
                                          in.tx_[k], in.un_[k], h.s11_p[k] + h.s22_p[k], s11p_l[k] + s22p_l[k], h.s11_p[k] - h.s22_p[k],
                                          s11p_l[k] - s22p_l[k], o.n12[k], h.s12_p[k]);
                o.w2[k] = updD[k] ? val : h.u_p[k];
            }
        }
    } else {
        {  // C: u at row t-1 (reads old v rows t-1, t)
            double mp_l[NC], ap_l[NC], alp_l[NC], vp_l[NC], vc_l[NC], vep_l[NC], ve0_l[NC], s11p_l[NC], s22p_l[NC];
            lefts<NC>(h.m_p, mp_l);
            lefts<NC>(h.a_p, ap_l);
            lefts<NC>(h.al_p, alp_l);
            lefts<NC>(h.v_p, vp_l);
            lefts<NC>(h.v_c, vc_l);
            lefts<NC>(h.ve_p, vep_l);
            lefts<NC>(in.ve0, ve0_l);
            lefts<NC>(h.s11_p, s11p_l);
            lefts<NC>(h.s22_p, s22p_l);
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const double vbar = ((vp_l[k] + h.v_p[k]) / 2 + (vc_l[k] + h.v_c[k]) / 2) / 2;
                const double vebar = ((vep_l[k] + h.ve_p[k]) / 2 + (ve0_l[k] + in.ve0[k]) / 2) / 2;
                const double val = u_node(mm, p, actu[k], h.m_p[k], mp_l[k], h.a_p[k], ap_l[k], h.al_p[k], alp_l[k], h.u_p[k], vbar, h.ue_p[k], vebar,
                                          in.tx_[k], in.un_[k], h.s11_p[k] + h.s22_p[k], s11p_l[k] + s22p_l[k], h.s11_p[k] - h.s22_p[k],
                                          s11p_l[k] - s22p_l[k], o.n12[k], h.s12_p[k]);
                o.w1[k] = updC[k] ? val : h.u_p[k];
            }
        }
        {  // D: v at row t-1 (reads new u rows t-2, t-1)
            double wp_r[NC], wn_r[NC], uepp_r[NC], uep_r[NC], s12p_r[NC];
            rights<NC>(h.w_p, wp_r);
            rights<NC>(o.w1, wn_r);
            rights<NC>(h.ue_pp, uepp_r);
            rights<NC>(h.ue_p, uep_r);
            rights<NC>(h.s12_p, s12p_r);
#pragma unroll
            for (int k = 0; k < NC; k++) {
                const double ubar = ((h.w_p[k] + wp_r[k]) / 2 + (o.w1[k] + wn_r[k]) / 2) / 2;
                const double uebar = ((h.ue_pp[k] + uepp_r[k]) / 2 + (h.ue_p[k] + uep_r[k]) / 2) / 2;
                const double val = v_node(mm, p, actv2, h.m_p[k], h.m_pp[k], h.a_p[k], h.a_pp[k], h.al_p[k], h.al_pp[k], h.v_p[k], ubar, h.ve_p[k], uebar,
                                          in.ty_[k], in.vn_[k], h.s11_p[k] + h.s22_p[k], h.s11_pp[k] + h.s22_pp[k], h.s11_p[k] - h.s22_p[k],
                                          h.s11_pp[k] - h.s22_pp[k], s12p_r[k], h.s12_p[k]);
                o.w2[k] = updD[k] ? val : h.v_p[k];
            }
        }
    }
}

// ---- the kernel ----------------------------------------------------------------------------
// NC: columns per lane.  VFIRST: odd substep (v then u, se.jl:183-187) or even (u then v, :178-182).
// AUX: also write alpha, zeta_c, zeta_f, Delta (last substep of a stage).
template <int NC, bool VFIRST, bool AUX>
__global__ void __launch_bounds__(Geo<NC>::THREADS, CSI_FUSED_MINB) k_evp_substep_fused(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params p)
{
    using G = Geo<NC>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stages = reinterpret_cast<double *>(smem_raw);
    uint64_t *full = reinterpret_cast<uint64_t *>(stages + (size_t)NSTAGE * NIN * BOX);
    uint64_t *empty = full + NSTAGE;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ia = p.a0 + blockIdx.x * G::OUTX;  // first output column of this strip
    const int y0 = p.sy0 < p.vy0 ? p.sy0 : p.vy0, y1 = p.sy1 > p.vy1 ? p.sy1 : p.vy1;
    const int ja = y0 + blockIdx.y * p.LY;
    const int jb = min(ja + p.LY - 1, y1);
    const int fin = p.in_set ? F_U1 : F_U0, fout = p.out_set ? F_U1 : F_U0;
    const int t_begin = ja - 3, t_end = jb + 1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], G::NCW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            const int xcoord = ia - 2 - 1 + OX;  // internal column of the strip's first (overlap) column
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_begin; t <= t_end; t++) {
                if (t - t_begin >= NSTAGE) mbar_wait(&empty[s], ph ^ 1);
                double *st = stages + (size_t)s * NIN * BOX;
                const int ru = t - 1, rv = VFIRST ? t : t - 1;  // rows of the u / v updates of this step
                int n = 10;
                auto ld = [&](int slot, int field, int r) { tma_load_row(st + slot * BOX, &tmap, &full[s], xcoord, r - 1 + p.oy, field); };
                ld(I_U, fin + 0, t + 1);
                ld(I_V, fin + 1, t + 1);
                ld(I_H, F_H, t);
                ld(I_A, F_A, t);
                ld(I_P, F_P, t);
                ld(I_S11, fin + 2, t);
                ld(I_S22, fin + 3, t);
                ld(I_S12, fin + 4, t);
                ld(I_UN, F_UN, ru);
                ld(I_VN, F_VN, rv);
                if (p.use_top) {
                    ld(I_TX, F_TX, ru);
                    ld(I_TY, F_TY, rv);
                    n += 2;
                }
                if (p.use_ue) {
                    ld(I_UE, F_UE, t);
                    ld(I_VE, F_VE, t);
                    n += 2;
                }
                mbar_expect_tx(&full[s], (uint32_t)n * BOX * sizeof(double));
                if (++s == NSTAGE) {
                    s = 0;
                    ph ^= 1;
                }
            }
        }
        return;
    }

    // ================= consumer warps =================
    const int cw = warp - 1;                  // consumer warp index
    const int q0 = cw * G::WOUT + NC * lane;  // strip-local index of this lane's first column
    const int i0 = ia - 2 + q0;               // reference column index of it

    // per-column constants of the march: output predicates, image offsets, wall masks
    bool st_s[NC], st_v[NC], actu[NC], evx[NC];
    int ix[NC];
#pragma unroll
    for (int k = 0; k < NC; k++) {
        const int wq = NC * lane + k, i = i0 + k;
        const bool outc = wq >= 2 && wq <= G::WCOLS - 3;
        st_s[k] = outc && i >= p.sx0 && i <= p.sx1;
        st_v[k] = outc && i >= p.vx0 && i <= p.vx1;
        actu[k] = !(p.bounded_x && (i <= 1 || i > p.Nx));  // !peripheral_node(f,c,c)
        evx[k] = i >= p.cx0 && i <= p.cx1;
        ix[k] = p.px ? (i <= W ? p.Nx : (i > p.Nx - W ? -p.Nx : 0)) : 0;
    }
    const size_t plane = (size_t)p.pitch * p.rows;
    double *const bS11 = p.base + (size_t)(fout + 2) * plane + (size_t)(i0 - 1 + OX);
    double *const bS22 = p.base + (size_t)(fout + 3) * plane + (size_t)(i0 - 1 + OX);
    double *const bS12 = p.base + (size_t)(fout + 4) * plane + (size_t)(i0 - 1 + OX);
    double *const bU = p.base + (size_t)(fout + 0) * plane + (size_t)(i0 - 1 + OX);
    double *const bV = p.base + (size_t)(fout + 1) * plane + (size_t)(i0 - 1 + OX);

    Hist<NC> h;
    {
        double *z = reinterpret_cast<double *>(&h);
#pragma unroll
        for (int k = 0; k < (int)(sizeof(Hist<NC>) / sizeof(double)); k++) z[k] = 0.0;
    }

    // store val at (column k, row r) of the array starting at b, plus its periodic images
    auto put = [&](double *b, int k, int r, int iy, double val) {
        double *q = b + (size_t)(r - 1 + p.oy) * p.pitch + k;
        q[0] = val;
        if (ix[k]) q[ix[k]] = val;
        if (iy) {
            q += (ptrdiff_t)iy * p.pitch;
            q[0] = val;
            if (ix[k]) q[ix[k]] = val;
        }
    };

    int s = 0;
    uint32_t ph = 0;
    for (int t = t_begin; t <= t_end; t++) {
        // ---- take this step's rows out of the stage, then hand the stage back ----
        mbar_wait(&full[s], ph);
        const double *st = stages + (size_t)s * NIN * BOX + q0;
        In<NC> in;
#pragma unroll
        for (int k = 0; k < NC; k++) {
            in.u_n[k] = st[I_U * BOX + k];
            in.v_n[k] = st[I_V * BOX + k];
            in.h0[k] = st[I_H * BOX + k];
            in.a0[k] = st[I_A * BOX + k];
            in.P0[k] = st[I_P * BOX + k];
            in.o11[k] = st[I_S11 * BOX + k];
            in.o22[k] = st[I_S22 * BOX + k];
            in.o12[k] = st[I_S12 * BOX + k];
            in.un_[k] = st[I_UN * BOX + k];
            in.vn_[k] = st[I_VN * BOX + k];
            in.tx_[k] = p.use_top ? st[I_TX * BOX + k] : p.ttx;
            in.ty_[k] = p.use_top ? st[I_TY * BOX + k] : p.tty;
            in.ue0[k] = p.use_ue ? st[I_UE * BOX + k] : p.ue_c;
            in.ve0[k] = p.use_ue ? st[I_VE * BOX + k] : p.ve_c;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == NSTAGE) {
            s = 0;
            ph ^= 1;
        }

        // rows of the two velocity updates of this step and whether those cells evolve
        const int rC = VFIRST ? t : t - 1, rD = t - 1;
        const bool evC = rC >= p.cy0 && rC <= p.cy1, evD = rD >= p.cy0 && rD <= p.cy1;
        bool updC[NC], updD[NC];
#pragma unroll
        for (int k = 0; k < NC; k++) {
            updC[k] = evC && evx[k];
            updD[k] = evD && evx[k];
        }
        const bool actv1 = !(p.bounded_y && (t <= 1 || t > p.Ny));          // v at row t (VFIRST, phase C)
        const bool actv2 = !(p.bounded_y && (t - 1 <= 1 || t - 1 > p.Ny));  // v at row t-1 (phase D)

        Out<NC> o;
        {
            MathFast mf;
            compute_step<NC, VFIRST>(mf, p, h, in, o, actu, actv1, actv2, updC, updD);
            if (__any_sync(0xffffffffu, mf.bad())) {
                // an operand left the exponent window of the shortcut quotients: redo this step with IEEE divisions
                MathSlow ms;
                compute_step<NC, VFIRST>(ms, p, h, in, o, actu, actv1, actv2, updC, updD);
            }
        }

        // ---- stores (home cell + periodic images + wall cells) ----
        {
            const bool row_s = t >= ja && t <= jb && t >= p.sy0 && t <= p.sy1;
            if (row_s) {
                const int iy = p.py ? (t <= W ? p.Ny : (t > p.Ny - W ? -p.Ny : 0)) : 0;
#pragma unroll
                for (int k = 0; k < NC; k++)
                    if (st_s[k]) {
                        put(bS11, k, t, iy, o.n11[k]);
                        put(bS22, k, t, iy, o.n22[k]);
                        put(bS12, k, t, iy, o.n12[k]);
                        if (AUX) {
                            const ptrdiff_t d = (ptrdiff_t)plane;
                            put(bS11 + ((ptrdiff_t)F_ALPHA - (fout + 2)) * d, k, t, iy, o.gc[k]);
                            put(bS11 + ((ptrdiff_t)F_ZC - (fout + 2)) * d, k, t, iy, o.zc[k]);
                            put(bS11 + ((ptrdiff_t)F_ZF - (fout + 2)) * d, k, t, iy, o.zf[k]);
                            put(bS11 + ((ptrdiff_t)F_DELTA - (fout + 2)) * d, k, t, iy, o.Dc[k]);
                        }
                    }
            }
            auto store_vel = [&](double *b, int r, const double (&val)[NC], bool is_u) {
                if (!(r >= ja && r <= jb && r >= p.vy0 && r <= p.vy1)) return;
                const int iy = p.py ? (r <= W ? p.Ny : (r > p.Ny - W ? -p.Ny : 0)) : 0;
#pragma unroll
                for (int k = 0; k < NC; k++)
                    if (st_v[k]) {
                        put(b, k, r, iy, val[k]);
                        // walls: one tangential halo cell (value / no-flux BC), as fill_halo_regions! does
                        double *q = b + (size_t)(r - 1 + p.oy) * p.pitch + k;
                        const int i = i0 + k;
                        if (is_u && p.bounded_y) {
                            if (r == 1) q[-p.pitch] = p.u_sn_bc == CSI_BC_VALUE ? val[k] + ((val[k] - p.u_sn_val) / (p.dy / 2)) * (-p.dy) : val[k];
                            if (r == p.Ny) q[p.pitch] = p.u_sn_bc == CSI_BC_VALUE ? val[k] + ((p.u_sn_val - val[k]) / (p.dy / 2)) * p.dy : val[k];
                        }
                        if (!is_u && p.bounded_x) {
                            if (i == 1) q[-1] = p.v_we_bc == CSI_BC_VALUE ? val[k] + ((val[k] - p.v_we_val) / (p.dx / 2)) * (-p.dx) : val[k];
                            if (i == p.Nx) q[1] = p.v_we_bc == CSI_BC_VALUE ? val[k] + ((p.v_we_val - val[k]) / (p.dx / 2)) * p.dx : val[k];
                        }
                    }
            };
            if (VFIRST) {
                store_vel(bV, rC, o.w1, false);
                store_vel(bU, rD, o.w2, true);
            } else {
                store_vel(bU, rC, o.w1, true);
                store_vel(bV, rD, o.w2, false);
            }
        }

        // ---- shift the row history ----
#pragma unroll
        for (int k = 0; k < NC; k++) {
            h.u_p[k] = h.u_c[k];
            h.u_c[k] = in.u_n[k];
            h.v_p[k] = h.v_c[k];
            h.v_c[k] = in.v_n[k];
            h.udx_c[k] = o.undx[k];
            h.vdx_c[k] = o.vndx[k];
            h.e11_p[k] = o.e11_c[k];
            h.e22_p[k] = o.e22_c[k];
            h.e12_c[k] = o.e12_n[k];
            h.P_p[k] = in.P0[k];
            h.m_pp[k] = h.m_p[k];
            h.m_p[k] = o.m_c[k];
            h.a_pp[k] = h.a_p[k];
            h.a_p[k] = in.a0[k];
            h.al_pp[k] = h.al_p[k];
            h.al_p[k] = o.gc[k];
            h.s11_pp[k] = h.s11_p[k];
            h.s22_pp[k] = h.s22_p[k];
            h.s11_p[k] = o.n11[k];
            h.s22_p[k] = o.n22[k];
            h.s12_p[k] = o.n12[k];
            h.w_p[k] = o.w1[k];
            h.ue_pp[k] = h.ue_p[k];
            h.ue_p[k] = in.ue0[k];
            h.ve_p[k] = in.ve0[k];
        }
    }
}

// ---- self test of the FAST arithmetic against the IEEE operators ---------------------------------
// Each thread draws pseudo-random operands (splitmix64; random significands, exponents spread over
// +-2^span) and counts results of MathFast that differ in any bit from the hardware's IEEE result
// while MathFast itself reported the operands inside its windows.  out[0..3] = mismatches of
// rcp, division, sqrt, constant-quotient; out[4] = samples the windows rejected.
__device__ __forceinline__ uint64_t splitmix(uint64_t &s)
{
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double rnd_double(uint64_t &s, int span, bool positive)
{
    const uint64_t r = splitmix(s);
    const uint64_t mant = r & 0x000fffffffffffffull;
    const int e = 1023 + (int)((r >> 52) % (uint64_t)(2 * span + 1)) - span;
    const uint64_t sign = positive ? 0ull : (splitmix(s) & 1ull) << 63;
    return __longlong_as_double((long long)(sign | ((uint64_t)e << 52) | mant));
}
__global__ void k_selftest_math(unsigned long long *out, uint64_t seed, int iters, int span)
{
    uint64_t s = seed + (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * 0x632be59bd9b4e019ull;
    unsigned long long bad[5] = {0, 0, 0, 0, 0};
    for (int it = 0; it < iters; it++) {
        double x = rnd_double(s, span, false), y = rnd_double(s, span, false), z = rnd_double(s, span, true);
        if ((it & 15) == 0) {  // near-special operands: perfect squares, powers of two, x close to y
            const double t = rnd_double(s, 20, true);
            z = t * t;
            if (it & 16) y = x * (1.0 + 1.1102230246251565e-16 * (double)(it & 7));
        }
        if ((it & 63) == 1) x = (it & 64) ? 0.0 : -0.0;
        MathFast m;
        const double r = m.rcp(y);
        if (!m.bad() && __double_as_longlong(r) != __double_as_longlong(1.0 / y)) bad[0]++;
        const double q = m.div(x, y);
        if (!m.bad() && __double_as_longlong(q) != __double_as_longlong(x / y)) bad[1]++;
        const double g = m.sqrt_(z);
        if (!m.bad() && __double_as_longlong(g) != __double_as_longlong(sqrt(z))) bad[2]++;
        MathFast m2;
        const double c = 4000.0 * (1.0 + (double)(it & 1023));
        const double qc = m2.divc(x, c, 1.0 / c);
        if (!m2.bad() && __double_as_longlong(qc) != __double_as_longlong(x / c)) bad[3]++;
        if (m.bad() || m2.bad()) bad[4]++;
    }
    for (int k = 0; k < 5; k++)
        if (bad[k]) atomicAdd(&out[k], bad[k]);
}
int selftest_math(long long samples, unsigned long long seed, int span, unsigned long long *out5)
{
    unsigned long long *d;
    if (cudaMalloc(&d, 5 * sizeof(unsigned long long)) != cudaSuccess) return 1;
    cudaMemset(d, 0, 5 * sizeof(unsigned long long));
    const int threads = 256, blocks = 148 * 8;
    const int iters = (int)((samples + (long long)threads * blocks - 1) / ((long long)threads * blocks));
    k_selftest_math<<<blocks, threads>>>(d, seed, iters, span);
    cudaError_t e = cudaMemcpy(out5, d, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : (int)e;
}
}  // namespace fz
namespace fz {

// ---- pack / unpack between the caller's Oceananigans parents and the internal layout ------------
struct PackItem {
    DArr a;
    int field;   // internal field index (first copy)
    int dup;     // also write field + 5 (second copy of an evolving field)
    int lx, ly;  // location, for the extent of the copy window on Bounded axes
};
__global__ void k_pack(PackItem it, Params p, int w)
{
    // window: i in [1-w, Nx+w(+1)], j in [1-w', Ny+w'(+1)] clipped to the parent
    const int i = 1 - w + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = 1 - p.oy + blockIdx.y;
    if (i > p.Nx + w + 1 || j > p.Ny + p.oy) return;
    const int pi = i - 1 + it.a.ox, pj = j - 1 + it.a.oy;
    double val = 0.0;
    if (it.a.p && pi >= 0 && pi < it.a.sx && pj >= 0 && pj < it.a.sy) val = it.a.p[(size_t)pj * it.a.sx + pi];
    const size_t plane = (size_t)p.pitch * p.rows;
    const size_t off = (size_t)(j - 1 + p.oy) * p.pitch + (size_t)(i - 1 + OX);
    p.base[(size_t)it.field * plane + off] = val;
    if (it.dup) p.base[(size_t)(it.field + 5) * plane + off] = val;
}
__global__ void k_unpack(PackItem it, Params p, int i0, int i1, int j0, int j1)
{
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = j0 + blockIdx.y;
    if (i > i1 || j > j1) return;
    const size_t plane = (size_t)p.pitch * p.rows;
    at(it.a, i, j) = p.base[(size_t)it.field * plane + (size_t)(j - 1 + p.oy) * p.pitch + (size_t)(i - 1 + OX)];
}

}  // namespace fz

// ---- host side ----------------------------------------------------------------------------------
struct FusedPlan {
    double *base = nullptr;
    int pitch = 0, rows = 0, oy = 0;
    CUtensorMap tmap;
    int Nx = 0, Ny = 0;
    bool attr_set = false;
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

int fused_supported(const DGrid &g, const DParams &p, const DFields &f, char *why, int nwhy)
{
    if (g.mask) { snprintf(why, nwhy, "immersed masks run on the unfused path"); return 0; }
    if (!recip_is_safe(g.dx) || !recip_is_safe(g.dy) || !recip_is_safe(g.az)) { snprintf(why, nwhy, "grid metric not eligible for the constant-division shortcut"); return 0; }
    if (g.topo_y == CSI_BOUNDED && (g.conn_s || g.conn_n)) { snprintf(why, nwhy, "Bounded y with slabs"); return 0; }
    if (g.Nx < 8 || g.Ny < 8) { snprintf(why, nwhy, "grid too small"); return 0; }
    if ((f.ue.p == nullptr) != (f.ve.p == nullptr)) { snprintf(why, nwhy, "ue/ve kinds differ"); return 0; }
    (void)p;
    return 1;
}

FusedPlan *fused_create(const DGrid &g, const DParams &, char *err, int nerr)
{
    using namespace fz;
    FusedPlan *pl = new FusedPlan();
    pl->Nx = g.Nx;
    pl->Ny = g.Ny;
    pl->oy = (g.conn_s || g.conn_n) ? g.Hy : W + 1;
    pl->pitch = ((OX + g.Nx + 1 + W + 1 + 15) / 16) * 16;
    pl->rows = g.Ny + 2 * pl->oy + 1;
    const size_t bytes = (size_t)NF * pl->pitch * pl->rows * sizeof(double);
    cudaError_t e = cudaMalloc(&pl->base, bytes);
    if (e != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); delete pl; return nullptr; }
    cudaMemset(pl->base, 0, bytes);
    cudaDeviceSynchronize();  // the plan may be used next from a non-blocking stream
    EncodeTiledFn enc = get_encode();
    if (!enc) { snprintf(err, nerr, "cuTensorMapEncodeTiled unavailable"); cudaFree(pl->base); delete pl; return nullptr; }
    cuuint64_t dims[3] = {(cuuint64_t)pl->pitch, (cuuint64_t)pl->rows, (cuuint64_t)NF};
    cuuint64_t strides[2] = {(cuuint64_t)pl->pitch * 8, (cuuint64_t)pl->pitch * pl->rows * 8};
    cuuint32_t box[3] = {(cuuint32_t)BOX, 1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, pl->base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(err, nerr, "cuTensorMapEncodeTiled failed (%d)", (int)r); cudaFree(pl->base); delete pl; return nullptr; }
    return pl;
}

void fused_destroy(FusedPlan *pl)
{
    if (!pl) return;
    if (pl->base) cudaFree(pl->base);
    delete pl;
}

constexpr int FUSED_NC = CSI_FUSED_NC;  // columns per lane of the production variant

template <bool VFIRST, bool AUX> static cudaError_t launch_one(const FusedPlan *pl, const fz::Params &P, dim3 grid, cudaStream_t s)
{
    using namespace fz;
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(k_evp_substep_fused<FUSED_NC, VFIRST, AUX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    k_evp_substep_fused<FUSED_NC, VFIRST, AUX><<<grid, Geo<FUSED_NC>::THREADS, SMEM_BYTES, s>>>(pl->tmap, P);
    return cudaGetLastError();
}

int fused_run(FusedPlan *pl, const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, int first_sub, int nsub,
              char *err, int nerr)
{
    using namespace fz;
    Params P;
    memset(&P, 0, sizeof P);
    P.Nx = g.Nx; P.Ny = g.Ny; P.pitch = pl->pitch; P.rows = pl->rows; P.oy = pl->oy;
    P.px = g.topo_x == CSI_PERIODIC;
    P.py = g.topo_y == CSI_PERIODIC && !g.conn_s && !g.conn_n;
    P.bounded_x = g.topo_x == CSI_BOUNDED;
    P.bounded_y = g.topo_y == CSI_BOUNDED;
    // stresses: interior for periodic axes, one extra ring on Bounded axes (boundary nodes of sigma12 and
    // the first halo cell, which the reference also evolves, evp.jl:145); slabs: the widened range
    P.sx0 = P.bounded_x ? 0 : 1; P.sx1 = P.bounded_x ? g.Nx + 1 : g.Nx;
    P.sy0 = P.bounded_y ? 0 : 1; P.sy1 = P.bounded_y ? g.Ny + 1 : g.Ny;
    P.vx0 = 1; P.vx1 = g.Nx; P.vy0 = 1; P.vy1 = g.Ny;
    if (g.conn_s) { P.sy0 = -g.Hy + 2; P.vy0 = -g.Hy + 2; }
    if (g.conn_n) { P.sy1 = g.Ny + g.Hy - 1; P.vy1 = g.Ny + g.Hy - 1; }
    const int BIG = 1 << 29;
    P.cx0 = P.px ? -BIG : 1; P.cx1 = P.px ? BIG : g.Nx;
    P.cy0 = P.py ? -BIG : P.vy0; P.cy1 = P.py ? BIG : P.vy1;
    P.use_top = p.top_kind == CSI_STRESS_FIELD;
    P.use_ue = f.ue.p != nullptr && p.bot_kind == CSI_STRESS_SEMI_IMPLICIT;
    P.u_sn_bc = p.u_sn_bc; P.v_we_bc = p.v_we_bc; P.u_sn_val = p.u_sn_val; P.v_we_val = p.v_we_val;
    P.dt = dt;
    P.dx = g.dx; P.dy = g.dy; P.az = g.az; P.dx2 = g.dx * g.dx; P.dy2 = g.dy * g.dy;
    P.rdx = 1.0 / g.dx; P.rdy = 1.0 / g.dy; P.raz = 1.0 / g.az;
    P.em2 = p.em2; P.Dmin = p.Dmin; P.amin = p.amin; P.amax = p.amax; P.amax2 = p.amax * p.amax; P.ca = p.ca;
    P.rho_i = p.rho_i; P.rhoCd = p.rho_e * p.Cd; P.f = p.f; P.min_mass = p.min_mass; P.min_conc = p.min_conc;
    P.ttx = p.top_kind == CSI_STRESS_CONST ? p.ttx : 0.0;
    P.tty = p.top_kind == CSI_STRESS_CONST ? p.tty : 0.0;
    P.ue_c = p.ue_c; P.ve_c = p.ve_c;
    P.pform = p.pform; P.cor = p.cor; P.sis = p.bot_kind == CSI_STRESS_SEMI_IMPLICIT;
    P.base = pl->base;

    // the TMA box of strip k starts at internal column a0 - 3 + OX + OUTX k: keep it even (16-byte aligned)
    P.a0 = P.sx0 < P.vx0 ? P.sx0 : P.vx0;
    if ((P.a0 - 3 + OX) & 1) P.a0 -= 1;
    const int ncols = (P.sx1 > P.vx1 ? P.sx1 : P.vx1) - P.a0 + 1;
    const int nrows = (P.sy1 > P.vy1 ? P.sy1 : P.vy1) - (P.sy0 < P.vy0 ? P.sy0 : P.vy0) + 1;
    const int OUTX = Geo<FUSED_NC>::OUTX;
    const int strips = (ncols + OUTX - 1) / OUTX;
    // rows per CTA: aim for a whole number of waves of 2 CTAs per SM over 148 SMs
    int LY = 128;
    {
        static int per_sm = 0;
        if (!per_sm) {
            int a = 1, b = 1, nsm = 148;
            cudaFuncSetAttribute(k_evp_substep_fused<FUSED_NC, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
            cudaFuncSetAttribute(k_evp_substep_fused<FUSED_NC, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_evp_substep_fused<FUSED_NC, true, false>, Geo<FUSED_NC>::THREADS, SMEM_BYTES);
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_evp_substep_fused<FUSED_NC, false, false>, Geo<FUSED_NC>::THREADS, SMEM_BYTES);
            int dev = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
            per_sm = (a < b ? a : b) * nsm;
            if (per_sm < 1) per_sm = 148;
        }
        const int slots = per_sm;
        int best = 128;
        double best_eff = 0.0;
        for (int ly = 48; ly <= 512; ly += 8) {
            const int ctas = strips * ((nrows + ly - 1) / ly);
            const int waves = (ctas + slots - 1) / slots;
            const double eff = (double)ctas / (waves * slots) * ((double)ly / (ly + 5));
            if (eff > best_eff) { best_eff = eff; best = ly; }
        }
        LY = best;
    }
    P.LY = LY;
    dim3 grid(strips, (nrows + LY - 1) / LY);

    // pack: caller parents -> internal layout (window includes W halo cells; evolving fields into both copies)
    const int w = W;
    auto pack = [&](const DArr &a, int field, int dup, int lx, int ly) {
        PackItem it{a, field, dup, lx, ly};
        dim3 pg((g.Nx + 2 * w + 2 + 127) / 128, g.Ny + 2 * pl->oy);
        k_pack<<<pg, 128, 0, c.stream>>>(it, P, w);
        ++*c.launches;
    };
    pack(f.u, F_U0, 1, 1, 0); pack(f.v, F_V0, 1, 0, 1);
    pack(f.s11, F_S11_0, 1, 0, 0); pack(f.s22, F_S22_0, 1, 0, 0); pack(f.s12, F_S12_0, 1, 1, 1);
    pack(f.h, F_H, 0, 0, 0); pack(f.a, F_A, 0, 0, 0); pack(f.P, F_P, 0, 0, 0);
    pack(f.un, F_UN, 0, 1, 0); pack(f.vn, F_VN, 0, 0, 1);
    if (P.use_top) { pack(f.top_x, F_TX, 0, 1, 0); pack(f.top_y, F_TY, 0, 0, 1); }
    if (P.use_ue) { pack(f.ue, F_UE, 0, 1, 0); pack(f.ve, F_VE, 0, 0, 1); }

    int in_set = 0;
    for (int k = 0; k < nsub; k++) {
        const int sub = first_sub + k;
        P.in_set = in_set;
        P.out_set = in_set ^ 1;
        const bool vfirst = (sub % 2) != 0;  // se.jl:178-187: odd substeps update v first
        const bool aux = k == nsub - 1;
        cudaError_t e;
        if (vfirst) e = aux ? launch_one<true, true>(pl, P, grid, c.stream) : launch_one<true, false>(pl, P, grid, c.stream);
        else e = aux ? launch_one<false, true>(pl, P, grid, c.stream) : launch_one<false, false>(pl, P, grid, c.stream);
        if (e != cudaSuccess) { snprintf(err, nerr, "launch: %s", cudaGetErrorString(e)); return (int)e; }
        ++*c.launches;
        in_set ^= 1;
    }
    // unpack the final copy into the caller's arrays (interior / stress window); halos are refilled by the caller
    auto unpack = [&](const DArr &a, int field, int i0, int i1, int j0, int j1) {
        if (!a.p) return;
        PackItem it{a, field, 0, 0, 0};
        dim3 ug((i1 - i0 + 1 + 127) / 128, j1 - j0 + 1);
        k_unpack<<<ug, 128, 0, c.stream>>>(it, P, i0, i1, j0, j1);
        ++*c.launches;
    };
    const int fo = in_set ? F_U1 : F_U0;
    if (nsub > 0) {
        unpack(f.u, fo + 0, P.vx0, P.vx1, P.vy0, P.vy1);
        unpack(f.v, fo + 1, P.vx0, P.vx1, P.vy0, P.vy1);
        unpack(f.s11, fo + 2, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.s22, fo + 3, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.s12, fo + 4, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.alpha, F_ALPHA, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.zc, F_ZC, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.zf, F_ZF, P.sx0, P.sx1, P.sy0, P.sy1);
        unpack(f.delta, F_DELTA, P.sx0, P.sx1, P.sy0, P.sy1);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, nerr, "%s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

}  // namespace csi
