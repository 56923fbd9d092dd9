// csi_fused.cu -- one kernel launch per EVP substep (sm_100a).
//
// Replaces, per substep, the reference's four kernels + two local halo fills
// (compute_stresses! evp:222-354, _u/_v_velocity_step! se:197-264, fill_halo_regions! se:170-187)
// with a single tiled kernel:
//
//   * Fields live in an internal planar layout owned by the plan: one allocation
//     [field][row][pitch], pitch a multiple of 16 doubles, interior column 1 at a 128-byte
//     boundary, a halo ring of W cells.  One 3-D TMA tensor map describes all of it.
//   * A CTA (256 threads) owns a tile: stresses on BX x BY = 32 x 16 nodes, velocities on the
//     30 x 14 cells inside.  One elected thread issues one `cp.async.bulk.tensor` per stencil field
//     (u, v, h, aice, P, s11, s22, s12, ue, ve): a 34 x 18 box = tile + halo, landing in shared
//     memory, completion on an mbarrier.  Then four phases, separated by block barriers:
//       A  strain rates e11, e22, e12 and ice mass on the haloed tile      (evp:360-375); on square grids u/dx, v/dx first
//       B  viscosities, replacement pressure, stress relaxation, alpha     (evp:236-354)
//       C  first velocity component on the tile + 1 ring                   (se:197-264)
//       D  second velocity component on the output cells
//     Strain rates, zeta, Delta, alpha, the new stresses and the first velocity never touch HBM
//     between phases.  About 73 KB of shared memory per CTA: three CTAs (24 warps) per SM, which is
//     what hides the 12-cycle FP64 latency (the earlier warp-marching variant kept all history in
//     registers, ran 8-12 warps per SM and was latency-bound; see git history and DESIGN.md).
//   * The five evolving fields are double buffered in HBM (read set A, write set B), so there is
//     no hazard between CTAs; the owner of a cell also stores its periodic images / wall values,
//     which replaces the two halo-fill launches per substep.
//   * Arithmetic is bit-exact (compiled with -fmad=false).  The FAST pass uses branch-free, correctly rounded
//     division / reciprocal / square root built from FMAs and a power-of-two-scaled form of the reference's expression
//     trees (see MathFast and the comment above vel_node_s); its premises are enforced by range validation of all inputs
//     once per stage (k_pack) and by window tests on the quotients and radicands that carry the state forward.  A tile
//     whose operands leave the windows is recomputed with plain IEEE operators and the reference's own trees (MathSlow);
//     csi_fused_stats reports how often that happened.
//
// Algorithmic HBM traffic: 14 loads + 5 stores per cell-update (u, v, s11, s22, s12 r/w; h, aice,
// P, un, vn, tau_x, tau_y, ue, ve read) = 152 B, 144 B by the SURVEY convention (P recomputable).
#include <cuda.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "csi_cell.cuh"
#include "csi_internal.h"

namespace csi {

namespace fz {

#ifndef CSI_FUSED_MINB_MET2
#define CSI_FUSED_MINB_MET2 2   // CTAs per SM of the two-dimensional-metric instantiation (more registers: its per-node loads in flight)
#endif
#ifndef CSI_FUSED_MINB_MET1
#define CSI_FUSED_MINB_MET1 2   // per-row metrics (lat-lon grids) with the run-time switches: 80 registers spill 0.6-0.8 KB per thread,
                                // measured +16 % at 128.  (The switch-free lat-lon variant -- BASELINE config 5 -- spills under 0.1 KB
                                // and loses 15 % at two CTAs per SM: it stays at CSI_FUSED_MINB.)
#endif
#ifndef CSI_FUSED_MINB_GEN
#define CSI_FUSED_MINB_GEN CSI_FUSED_MINB
#endif
#ifndef CSI_TILE_BY
#define CSI_TILE_BY 16
#endif
constexpr int BX = 32, BY = CSI_TILE_BY;  // stress nodes per tile (a warp owns two rows: BY = 2 x warps per CTA)
constexpr int OUTX = BX - 2, OUTY = BY - 2;  // velocity cells per tile
constexpr int SXD = BX + 2, SYD = BY + 2;    // shared-memory tile = TMA box: tile + 1 halo ring
#ifndef CSI_UNROLL_B
#define CSI_UNROLL_B 2
#endif
#ifndef CSI_UNROLL_CD
#define CSI_UNROLL_CD 2
#endif
constexpr int UNROLL_B = CSI_UNROLL_B, UNROLL_CD = CSI_UNROLL_CD;
constexpr int NT = 32 * (BY / 2);              // threads per CTA
constexpr int NIT = (SXD * SYD + NT - 1) / NT;  // sweeps of the CTA over the haloed tile
constexpr int ASTRIDE = ((SXD * SYD * 8 + 127) / 128) * 128 / 8;  // doubles between shared arrays (TMA destinations are 128-byte aligned)
constexpr int W = 3;             // halo ring kept valid in the internal layout
constexpr int OX = 16;           // internal column of i = 1 (128-byte aligned)

// internal field indices
enum { F_U0 = 0, F_V0, F_S11_0, F_S22_0, F_S12_0, F_U1, F_V1, F_S11_1, F_S22_1, F_S12_1,
       F_H, F_A, F_P, F_UN, F_VN, F_TX, F_TY, F_UE, F_VE, F_ALPHA, F_ZC, F_ZF, F_DELTA,
       // stage constants written once per stage by k_prep (see there): ice mass, and the top-stress term of the velocity
       // tendencies (tau_top / m_i * aice_i at u and v nodes)
       F_M, F_T1X, F_T1Y, NF_COMMON,
       // stage constants of the less common configurations (k_prep): the value of nodes that are not dynamically active
       // (marginal ice ? free-drift velocity : 0), the bottom term tau_bot / m_i * aice_i of a prescribed bottom stress, the
       // 4-point sums of the atmosphere velocity of a top SemiImplicitStress, and the packed free-drift arrays
       F_UFD = NF_COMMON, F_VFD, F_TB1X, F_TB1Y, F_SVA, F_SUA, F_FDU, F_FDV, NF_GEN,
       // optional stage constants (compile-time switches CSI_PRE_*, measured and rejected: DESIGN.md section 5.1): reciprocals of
       // the face-mass sums at u / v nodes (with the marginal-ice decision folded in), of the centre mass and of the corner mass
       // sum, the corner sum of P, and the 4-point sums of the ocean velocity at u / v nodes
       F_RM2U = NF_GEN, F_RM2V, F_RMC, F_RMF, F_PF4, F_SVE, F_SUE, NF };
#ifndef CSI_PRE_RM2
#define CSI_PRE_RM2 0
#endif
#ifndef CSI_PRE_RMC
#define CSI_PRE_RMC 0
#endif
#ifndef CSI_PRE_PF4
#define CSI_PRE_PF4 0
#endif
#ifndef CSI_PRE_SVE
#define CSI_PRE_SVE 0
#endif
// shared-memory arrays (each SXD x SYD doubles)
// (order matters in one place: the branch-free strain-rate sweep of phase A reads one row above / one column left of the tile in
// the u / dx and v / dx arrays, which live in the slots of A_AL and A_W -- i.e. the tail of whatever array precedes them; A_AL
// therefore follows an array nothing writes during that sweep, and A_W only ever reaches the unused padding of its predecessor)
enum { A_U = 0, A_V, A_H, A_A, A_P, A_S11, A_S22, A_S12, A_UE, A_VE, A_AL, A_E11, A_E22, A_E12, A_W, NARR };

constexpr size_t SMEM_BYTES = (size_t)NARR * ASTRIDE * sizeof(double) + 64 + ((SXD * SYD + 63) / 64) * 64;  // + two mbarriers + node flags

// planes of the pointwise inputs of one velocity phase (the component it updates)
struct PhasePtrs {
    const double *n;          // previous-stage velocity (u^n / v^n)
    const double *t1, *tt;    // top stress: FAST precomputed term (or the atmosphere velocity of a top SemiImplicitStress) / IEEE raw array
    const double *sa, *oth;   // top SemiImplicitStress: 4-point sum of the other atmosphere component (FAST) / its raw plane (IEEE)
    const double *tb1, *tbr;  // prescribed bottom stress: FAST precomputed term / IEEE raw array
    const double *fd;         // value of nodes that are not dynamically active
    const double *rm, *ue, *sv;  // experiments CSI_PRE_RM2 / CSI_PRE_SVE
};
struct Params {
    int Nx, Ny;          // interior size
    int pitch, rows;     // internal layout
    int oy;              // internal row of j = 1 is (oy)
    int px, py;          // periodic images along x / y
    int bounded_x, bounded_y;
    int wall_s, wall_n;  // physical walls of a Bounded y axis on this rank (a partition's connected side is not a wall)
    int wall_w, wall_e;  // likewise along x
    // store windows (reference indices, inclusive)
    int sx0, sx1, sy0, sy1;  // stresses
    int vx0, vx1, vy0, vy1;  // velocities
    int cx0, cx1, cy0, cy1;  // cells whose velocity is evolved (periodic images included); others keep their value
    int a0;                  // first column of strip 0 (chosen so every TMA box starts 16-byte aligned)
    int use_top, use_ue;     // field arrays present
    int u_sn_bc, v_we_bc;
    double u_sn_val, v_we_val;
    double dt;
    double dx, dy, az, dx2, dy2, rdx, rdy, raz;
    double em2, Dmin, amin, amax, amax2, ca, rho_i, rhoCd, f, min_mass, min_conc;
    double ttx, tty, ue_c, ve_c;
    double imm_u, imm_v;  // immersed linear-drag flux BC (0 = none)
    // power-of-two multiples used by the scaled expression tree of the FAST pass (all exact)
    double dt2, dt4, f4, Dmin2, Dmin8, min_mass2, min_conc2, dx2d, dy2d;
    double gnan;          // clamp(sqrt(alpha+^2), alpha-, alpha+): the reference's gamma where its gamma^2 is NaN (0 / 0 in open water)
    int sq;               // regular grid with dx == dy: phase A divides every u, v once (u/dx serves both operators)
    int pform, cor, sis;
    int in_set, out_set;  // 0 / 1: which copy of the evolving fields is read / written
    int ty0;              // first tile row of this launch (a substep may be launched in row bands, see fused_steps)
    // planes of this launch (set per substep by fused_steps): the pointwise inputs of the first (c) / second (d) velocity
    // phase -- previous-stage velocity, precomputed top-stress term (FAST pass), raw top stress (IEEE pass) -- and the outputs
    PhasePtrs pc, pd;
    const double *g_rmc, *g_rmf, *g_pf4;  // reciprocal centre mass / corner mass sum, corner sum of P (CSI_PRE_RMC, CSI_PRE_PF4)
    double *o_c, *o_d, *o_s11, *o_s22, *o_s12;  // first / second velocity component, stresses
    int use_t1;           // a top stress exists (field or constant): the FAST pass reads its precomputed term
    // less common configurations (GEN instantiation only)
    int fd_on, fd_kind;   // free drift: marginal-ice nodes take a stage-constant velocity (k_prep) instead of 0
    int top_sis, top_arr; // SemiImplicitStress on top; its u_a, v_a are arrays (else the constants ta_x, ta_y)
    int bot_expl, bot_arr, bot_kind, top_kind;  // prescribed bottom stress (numbers or arrays; constants tb_x, tb_y)
    double top_rhoCd, ta_x, ta_y, tb_x, tb_y;
    // tile columns / rows (inclusive) whose cells all lie inside every store window, have no periodic image and no wall
    // neighbour, and whose velocity nodes are all evolved: the vast majority; they skip the per-node edge tests
    int it_x0, it_x1, it_y0, it_y1;
    double *base;         // internal allocation
    const uint8_t *flags; // immersed-boundary node flags in the internal layout (rows x pitch bytes), or NULL
    const double *met;    // j-dependent metrics (lat-lon grids), or NULL on a regular grid: one record of MC_N doubles per row,
                          // the pointer pre-offset so that the record of reference row j starts at met[j * MC_N]; MET_PAD padding
                          // records on either side, so that no tile needs a clamp
    const double *met2;   // two-dimensional metrics (orthogonal curvilinear grids), or NULL: MC2_N planes in the internal layout
                          // (entry of node (i, j) at the node's in-plane offset), MET2_PAD padding rows on either side of each
                          // plane so that the neighbours an edge tile names need no clamp (they hold valid metrics of other nodes)
    long long met2_stride; // doubles between two planes
    int *invalid;         // device flag raised by k_pack when an input is neither zero nor in [2^-300, 2^300): the whole
                          // stage then runs the IEEE pass (the FAST pass relies on validated inputs, see MathFast)
};

// columns of the per-row metric table (internal row = j - 1 + oy): the twelve metrics, the squares the SBP operators use,
// the correctly rounded reciprocals of every metric that appears as a divisor, and f at (Face, Face)
enum { MC_DXCC = 0, MC_DXFC, MC_DXCF, MC_DXFF, MC_DYCC, MC_DYFC, MC_DYCF, MC_DYFF, MC_AZCC, MC_AZFC, MC_AZCF, MC_AZFF,
       MC_DXCC2, MC_DYCC2, MC_DXFF2, MC_DYFF2, MC_RDXFC, MC_RDXCF, MC_RDYFC, MC_RDYCF, MC_RAZCC, MC_RAZFC, MC_RAZCF, MC_RAZFF,
       MC_FFF, MC_N };
constexpr int MET_PAD = 40;  // padding records of the metric table (>= tile height + halos beyond either end)

// Metric<0>: the regular grid's constants (kernel parameters); Metric<1>: the row's own values, read from the records of the
// tile's rows staged in shared memory (warp-uniform, conflict-free broadcasts); Metric<2>: the node's own values, read through
// the read-only path from planes in the internal layout (coalesced along a row, neighbours from L1 / L2).
// r is the reference row index j, o the node's in-plane offset (only Metric<2> looks at it).
constexpr int MET_ROWS = BY + 5;  // rows J0 - 3 .. J0 + BY + 1: every row a tile's stencils and wall cells name
static_assert(MET_ROWS * MC_N <= SXD * SYD, "the staged metric records must fit the shared-memory array they borrow");
constexpr int MET2_PAD = BY + 8;  // padding rows of the two-dimensional metric planes (a tile's halo and its neighbours beyond either end)
constexpr int MC2_N = 20;         // planes: the twelve metrics and the eight reciprocals
__host__ __device__ constexpr int met2_plane_of(int col) { return col < 12 ? col : col - (MC_RDXFC - 12); }
template <int MET>
struct Metric {
    const Params &p;
    const double *smt;  // staged records (MET == 1): row r at smt[(r - rbase) * MC_N]
    int rbase;
    __device__ __forceinline__ double ld(int col, int o, int r) const
    {
        if (MET == 2) return __ldg(p.met2 + met2_plane_of(col) * p.met2_stride + o);
        return smt[(r - rbase) * MC_N + col];
    }
#define CSI_MET(name, col, regular) \
    __device__ __forceinline__ double name(int o, int r) const { return MET ? ld(col, o, r) : (regular); }
    CSI_MET(dxcc, MC_DXCC, p.dx) CSI_MET(dxfc, MC_DXFC, p.dx) CSI_MET(dxcf, MC_DXCF, p.dx) CSI_MET(dxff, MC_DXFF, p.dx)
    CSI_MET(dycc, MC_DYCC, p.dy) CSI_MET(dyfc, MC_DYFC, p.dy) CSI_MET(dycf, MC_DYCF, p.dy) CSI_MET(dyff, MC_DYFF, p.dy)
    CSI_MET(azcc, MC_AZCC, p.az) CSI_MET(azfc, MC_AZFC, p.az) CSI_MET(azcf, MC_AZCF, p.az) CSI_MET(azff, MC_AZFF, p.az)
    CSI_MET(rdxfc, MC_RDXFC, p.rdx) CSI_MET(rdxcf, MC_RDXCF, p.rdx) CSI_MET(rdyfc, MC_RDYFC, p.rdy) CSI_MET(rdycf, MC_RDYCF, p.rdy)
    CSI_MET(razcc, MC_RAZCC, p.raz) CSI_MET(razfc, MC_RAZFC, p.raz) CSI_MET(razcf, MC_RAZCF, p.raz) CSI_MET(razff, MC_RAZFF, p.raz)
#undef CSI_MET
    // the squares the SBP operators use (the reference writes dx^2 = dx * dx): staged per row, or formed from the node's metric
#define CSI_MET_SQ(name, col, base, regular)                                                           \
    __device__ __forceinline__ double name(int o, int r) const                                         \
    {                                                                                                   \
        if (MET == 2) { const double v = ld(base, o, r); return v * v; }                               \
        return MET ? ld(col, o, r) : (regular);                                                        \
    }
    CSI_MET_SQ(dxcc2, MC_DXCC2, MC_DXCC, p.dx2) CSI_MET_SQ(dycc2, MC_DYCC2, MC_DYCC, p.dy2)
    CSI_MET_SQ(dxff2, MC_DXFF2, MC_DXFF, p.dx2) CSI_MET_SQ(dyff2, MC_DYFF2, MC_DYFF, p.dy2)
#undef CSI_MET_SQ
    __device__ __forceinline__ double fff(int r) const { return MET == 1 ? smt[(r - rbase) * MC_N + MC_FFF] : p.f; }
    __device__ __forceinline__ double dxff2d(int o, int r) const { return MET ? 2 * dxff2(o, r) : p.dx2d; }  // 2 dx^2 (exact)
    __device__ __forceinline__ double dyff2d(int o, int r) const { return MET ? 2 * dyff2(o, r) : p.dy2d; }
};

// the metric factors of one velocity node's stress divergence (isd:39-51): d = a (sD1 - sD0) / 2,
// tt = (t2hi sT1 - t2lo sT0) / td / 2, SS = (s2hi s12hi - s2lo s12lo) / sd, all over az
struct NodeMetric {
    double a, t2hi, t2lo, td, rtd, s2hi, s2lo, sd, rsd, az, raz;
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
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680)  // suspend-time hint: sleep in the barrier unit instead of spinning through issue slots
        : "memory");
}
__device__ __forceinline__ void tma_load_row(double *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cta.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}

// L2 prefetch of one box (the pointwise inputs of the velocity phases: one instruction of the elected thread per field)
__device__ __forceinline__ void tma_prefetch_box(const CUtensorMap *map, int x, int y, int z)
{
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}
// pulls a contiguous run of global memory (16-byte aligned, a multiple of 16 bytes) towards L2
__device__ __forceinline__ void bulk_prefetch_l2(const void *g, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
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
    // Window accumulators on the high word of each checked double (sign, 11 exponent bits, 20 significand bits):
    //  * quotients (any sign, zero allowed) are tested two at a time on the top halves of their high words (sign, exponent,
    //    4 significand bits), packed into one register with the signs masked off: per 16-bit lane qmx = max g,
    //    qmn = min (g - 1) (VIMNMX.U16x2 / VIADDMNMX.U16x2)  -> |q| in [2^-300, 2^300) or g == 0.  A quotient waits in
    //    `pend` for its partner; straight-line code resolves `have` at compile time.  g == 0 is a zero -- or a magnitude
    //    below 2^-1026, which cannot occur: the pack kernel admits only inputs that are zero or in [2^-300, 2^300)
    //    (anything else sends the whole stage to the IEEE pass), checked quotients and radicands stay in that window,
    //    sums and differences of such values are zero or at least 2^-352, and no expression of the step multiplies more
    //    than two of them with constants from [1e-30, 1e30] (>= 2^-804).
    //  * divisors  (positive, non-zero): dacc = umax(hi - DLO)  -> d in [2^-255, 2^257); zeros, negatives, NaN wrap high;
    //    lo1 catches a low word of all ones (superset of "significand all ones", the exception of Markstein's theorem)
    //  * radicands: checked like quotients (a zero radicand -- ice exactly at rest -- is legitimate), plus a sign accumulator
    // With these windows the numerators x = q d stay in [2^-555, 2^557), far from where the exact
    // residual d q0 - x could underflow (2^-969) or anything could overflow.
    static constexpr bool SCALED = true;  // tile_pass evaluates the power-of-two-scaled expression tree (see there)
    uint32_t qmn = 0xffffffffu, qmx = 0u, lo1 = 0xffffffffu, dacc = 0u, neg = 0u, pend = 0u;
    bool have = false;
    static constexpr uint32_t QLO = 0x2d30u, QHI = 0x52afu;  // per lane: exponent field in [1023 - 300, 1023 + 300)
    static constexpr uint32_t DLO = 0x300u << 20, DSPAN = (0x200u << 20) - 1u;  // (public: k_prep applies the same test)
    __device__ __forceinline__ void chk2(uint32_t hi_a, uint32_t hi_b)
    {
        const uint32_t g = __byte_perm(hi_a, hi_b, 0x7632) & 0x7fff7fffu;
        qmx = __vmaxu2(qmx, g);
        qmn = __viaddmin_u16x2(g, 0xffffffffu, qmn);  // g == 0 wraps to 0xffff and is ignored
    }
    __device__ __forceinline__ void chkq(double q)
    {
#ifdef CSI_EXPERIMENT_NOCHECK
        return;
#endif
        const uint32_t hi = (uint32_t)__double2hiint(q);
        if (have) chk2(pend, hi);
        else pend = hi;
        have = !have;
    }
    __device__ __forceinline__ void chkd(double d)
    {
        dacc = max(dacc, (uint32_t)__double2hiint(d) - DLO);
        chklo(d);
    }
    __device__ __forceinline__ void chklo(double d) { lo1 = min(lo1, (uint32_t)__double2loint(d) + 1u); }
    __device__ __forceinline__ bool bad()
    {
        if (have) chk2(pend, 0u);
        have = false;
        const bool q_out = ((qmx & 0xffffu) > QHI) | ((qmx >> 16) > QHI) | ((qmn & 0xffffu) < QLO - 1u) | ((qmn >> 16) < QLO - 1u);
        return q_out | (dacc > DSPAN) | ((neg >> 31) != 0u) | (lo1 == 0u);
    }

    // B = true: the divisor's range is known from earlier checks (see the call sites); only its low word is tested
    template <bool B = false> __device__ __forceinline__ double rcp(double y)
    {
        if (B) chklo(y);
        else chkd(y);
        double r;
        asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(y));
        // one cubic step r (1 + e + e^2) takes the 2^-20 seed to below one ulp; Markstein's step then returns RN(1/y)
        double e = __fma_rn(-y, r, 1.0);
        e = __fma_rn(e, e, e);
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
    // *_nc: quotients whose operands are bounded by checks already made (see the call sites): no window test
    __device__ __forceinline__ double quot_nc(double x, double d, double r)
    {
        const double q0 = x * r;
        const double t = __fma_rn(d, q0, -x);
        return __fma_rn(-t, r, q0);
    }
    __device__ __forceinline__ double divc(double x, double d, double r) { return quot(x, d, r); }
    __device__ __forceinline__ double divc_nc(double x, double d, double r) { return quot_nc(x, d, r); }
    template <bool B = false> __device__ __forceinline__ NodeRecip recip(double d)
    {
        NodeRecip R;
        R.d = d;
        R.r = rcp<B>(d);
        return R;
    }
    __device__ __forceinline__ double divn(double x, const NodeRecip &R) { return quot(x, R.d, R.r); }
    __device__ __forceinline__ double divn_nc(double x, const NodeRecip &R) { return quot_nc(x, R.d, R.r); }
    template <bool B = false> __device__ __forceinline__ double div(double x, double y) { return quot(x, y, rcp<B>(y)); }
    template <bool B = false> __device__ __forceinline__ double div_nc(double x, double y) { return quot_nc(x, y, rcp<B>(y)); }
    // CHK = false: the radicand is a checked quotient times a bounded constant; only its sign is still tested
    // SGN = false: the radicand is a sum of squares (never negative, never -0)
    template <bool CHK = true, bool SGN = true> __device__ __forceinline__ double sqrt_(double x)
    {
        if (CHK) chkq(x);
        if (SGN) neg |= (uint32_t)__double2hiint(x);
        double y;
        asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
        // x = +0 gives y = +inf; clamped to 2^1023 the iteration below returns +0 without a select (-0, negative and NaN
        // radicands are rejected by the windows)
        y = __hiloint2double(min(__double2hiint(y), 0x7fe00000), __double2loint(y));
        double g = x * y, h = 0.5 * y;
        double e = __fma_rn(-h, g, 0.5);
        g = __fma_rn(g, e, g);
        h = __fma_rn(h, e, h);
        e = __fma_rn(-h, g, 0.5);
        g = __fma_rn(g, e, g);  // (h keeps its 2^-44 accuracy: enough for the correction term)
        const double d = __fma_rn(-g, g, x);
        return __fma_rn(d, h, g);
    }
    // max(s, c) for a square root s >= +0 (never NaN in a clean pass) and a constant c > 0
    __device__ __forceinline__ double max_pos(double s, double c) { return s > c ? s : c; }
};
struct MathSlow {
    static constexpr bool SCALED = false;  // the reference's expression tree, operator for operator
    __device__ __forceinline__ void chkq(double) {}
    __device__ __forceinline__ bool bad() { return false; }
    __device__ __forceinline__ double divc(double x, double d, double) { return x / d; }
    __device__ __forceinline__ double divc_nc(double x, double d, double) { return x / d; }
    __device__ __forceinline__ double divn_nc(double x, const NodeRecip &R) { return x / R.d; }
    __device__ __forceinline__ double max_pos(double s, double c) { return jl_max(s, c); }
    template <bool B = false> __device__ __forceinline__ NodeRecip recip(double d)
    {
        NodeRecip R;
        R.d = d;
        R.r = 0.0;
        return R;
    }
    __device__ __forceinline__ double divn(double x, const NodeRecip &R) { return x / R.d; }
    template <bool B = false> __device__ __forceinline__ double div(double x, double y) { return x / y; }
    template <bool B = false> __device__ __forceinline__ double div_nc(double x, double y) { return x / y; }
    template <bool CHK = true, bool SGN = true> __device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
};

// External-stress inputs of one velocity node (ext:8-40,84-146,176-210; stress_balance_free_drift.jl:61-129).  "own" is the
// component being updated, "other" the transverse one.  FAST pass: 4-point SUMS and precomputed stage-constant terms (k_prep);
// IEEE pass: the reference's means and the raw stresses.
struct Ext {
    double ue, oe;   // bottom SemiImplicitStress: u_e at the node; 4 x mean (FAST) / mean (IEEE) of the other component of u_e
    double ua, oa;   // top SemiImplicitStress likewise (atmosphere velocity)
    double t1;       // top stress that is nothing / numbers / arrays: FAST tau_top / m_i * aice_i (k_prep); IEEE raw tau_top
    double tb;       // bottom stress that is numbers / arrays:        FAST tau_bot / m_i * aice_i (k_prep); IEEE raw tau_bot
    double fd;       // value of a node that is not dynamically active: marginal ice ? free-drift velocity : 0 (k_prep)
};

// One velocity node, reference tree: u at (i, r) se:197-229, mt:11-41, ext:176-196, isd:39-44, evp:384,391-395 (*0 = column
// i-1), or v at (i, r) se:231-264, mt:44-74, ext:183-202, isd:46-51, evp:385,397-401 (*0 = row r-1; NEG_TT: the sign of the
// tension term).  old / obar: the node's own component and the 4-point mean of the other one.
template <bool GEN, bool NEG_TT, class M>
__device__ __forceinline__ double vel_node(M &mm, const Params &p, const NodeMetric &nm, bool active, double m1, double m0, double a1, double a0_, double al1, double al0,
                                           double old, double obar, double cross, const Ext &x, double vn, double sD1, double sD0,
                                           double sT1, double sT0, double s12hi, double s12lo, bool has_imm, double imm)
{
    const double mi = (m1 + m0) / 2, ai = (a1 + a0_) / 2, abar = (al1 + al0) / 2;
    const NodeRecip Ra = mm.recip(abar), Rm = mm.recip(mi);
    const double dtau = mm.divn(p.dt, Ra);
    double cbot = 0.0, tbot = 0.0, ctop = 0.0, ttop = x.t1;
    if (GEN ? p.sis != 0 : true) {   // implicit_tau_coefficient / explicit_tau of a SemiImplicitStress (ext:176-202)
        const double d_own = x.ue - old, d_oth = x.oe - obar;
        cbot = p.rhoCd * mm.sqrt_(d_own * d_own + d_oth * d_oth);
        tbot = cbot * x.ue;
    } else if (GEN && p.bot_expl) tbot = x.tb;
    if (GEN && p.top_sis) {
        const double d_own = x.ua - old, d_oth = x.oa - obar;
        ctop = p.top_rhoCd * mm.sqrt_(d_own * d_own + d_oth * d_oth);
        ttop = ctop * x.ua;
    }
    const double rheo = mm.divn(mm.div(vn - old, dtau), Ra);
    const double d = nm.a * (sD1 - sD0) / 2;
    const double tn = nm.t2hi * sT1 - nm.t2lo * sT0;
    const double tt = mm.divc(NEG_TT ? -tn : tn, nm.td, nm.rtd) / 2;
    const double SS = mm.divc(nm.s2hi * s12hi - nm.s2lo * s12lo, nm.sd, nm.rsd);
    const double dsig = mm.divc(d + tt + SS, nm.az, nm.raz);
    double G = -cross - mm.divn(ttop, Rm) * ai + mm.divn(tbot, Rm) * ai + mm.divn(dsig, Rm) + (has_imm ? mm.divn(imm, Rm) : 0.0) + (0.0 + rheo);
    G = mi <= 0 ? 0.0 : G;
    double tau = mm.divn(cbot - ctop, Rm) * ai;
    tau = mi <= 0 ? 0.0 : tau;
    const double D = mm.div(old + dtau * G, 1 + dtau * tau);
    const bool active_ice = (mi >= p.min_mass) & (ai >= p.min_conc);
    return jl_mul_bool(active_ice ? D : ((GEN && p.fd_on) ? x.fd : 0.0), active);  // free_drift = nothing: marginal ice -> 0
}

// ---- the scaled expression tree (FAST pass) -----------------------------------------------------
// Multiplying or dividing by a power of two is exact and commutes with every IEEE rounding as long as nothing
// leaves the normal range, so the FAST pass carries the reference's intermediate values times a known power of
// two and never spends an FP64 issue slot on "/ 2", "2 *" or "4 *":
//     strain rates   stored as 2 e11, 2 e22 (centres), 2 e12 (corners)
//     4-point means  carried as sums: 8 e11f, 8 e22f, 4 Pf, 4 mf, 4 vbar, 4 ve_bar;  4 e12c = (sum of 2 e12) / 2
//     2 Delta_c, 8 Delta_f, face sums 2 m_i, 2 aice_i, 2 alpha_bar, and 2 x the stress divergence
// with the factors cancelling inside quotients (P_f / 2 Delta_f = 4 P_f / 8 Delta_f, tau / m_i * aice_i = tau / 2 m_i * 2 aice_i ...)
// or absorbed by constants (2 dt, 4 dt, f / 4, 2 Delta_min, 8 Delta_min, 2 dx^2) and by explicit FMAs
// (a + X / 2 = fma(X, 0.5, a): one rounding of the same real number).  Every result handed on is bit-identical to the
// reference tree's.  The normal-range premise rests on an induction over the substeps: the pack kernels admit only inputs
// that are zero or in [2^-300, 2^300) (else the whole stage takes the IEEE pass); the quantities that carry the state
// forward -- the stress increments, the new velocities -- are window-tested quotients, so u, v, sigma stay zero or in
// [2^-352, 2^311); sums and differences of such values are zero or normal and halving them is exact; a quotient whose
// operands are bounded this way (by a cell area, by alpha, by the face mass after its range test, of a validated input)
// cannot leave the normal range and is not tested again; products that could underflow (squares of strain rates) only
// feed tested radicands, where an addend below 2^-1022 cannot move a sum of at least 2^-300.  A tile that leaves the
// windows is redone with the reference tree (MathSlow).
template <bool GEN, bool NEG_TT, class M>
__device__ __forceinline__ double vel_node_s(M &mm, const Params &p, const NodeMetric &nm, bool active, double m1, double m0, double a1, double a0_, double al1,
                                             double al0, double old, double osum, double cross, const Ext &x, double vn, double sD1,
                                             double sD0, double sT1, double sT0, double s12hi, double s12lo, bool has_imm, double imm2, double rm)
{
    // nm.s2hi, nm.s2lo hold 2 x the squared metrics; osum, x.oe, x.oa = 4 x the means of the other component; imm2 = 2 x the immersed term
    const double m2 = m1 + m0, a2 = a1 + a0_, ab2 = al1 + al0;
    // marginal ice and open water (face mass or concentration under the thresholds) get their stage-constant value (0, or the
    // free-drift velocity) whatever G is: a harmless mass keeps those nodes from failing the tile's divisor test
#if CSI_PRE_RM2
    // rm: the reciprocal of the face-mass sum from k_prep -- -1 for marginal ice / open water, +inf where the divisor
    // would have failed the range test (the poisoned quotients then fail the window test of the velocity quotient)
    const bool active_ice = rm > 0.0;
    NodeRecip Rm;
    Rm.d = active_ice ? m2 : 1.0;
    Rm.r = active_ice ? rm : 1.0;
    const NodeRecip Ra = mm.template recip<true>(ab2);
#else
    const bool active_ice = (m2 >= p.min_mass2) & (a2 >= p.min_conc2);
    const NodeRecip Ra = mm.template recip<true>(ab2), Rm = mm.recip(active_ice ? m2 : 1.0);
#endif
    const double dtau = mm.divn_nc(p.dt2, Ra);  // dt / alpha_bar; alpha in [alpha-, alpha+]: nothing to check
    // d_own, d_oth4: differences of validated inputs and checked quotients (zero or >= 2^-352, < 2^303): the squares stay normal
    double cbot = 0.0, ctop = 0.0;
    // the top term tau_top / m_i * aice_i and, for a prescribed bottom stress, the bottom term: stage constants computed once
    // per stage by k_prep with these very operations; a SemiImplicitStress side is evaluated here (ext:176-202)
    double topT = x.t1, botT = 0.0;
    if (GEN ? p.sis != 0 : true) {
        const double d_own = x.ue - old, d_oth4 = x.oe - osum;
        cbot = p.rhoCd * mm.template sqrt_<false, false>(__fma_rn(d_oth4 * d_oth4, 0.0625, d_own * d_own));
        // quotients over the face mass (its range is tested): tau_bottom and coef are a checked square root times a validated
        // input and a constant -- they cannot leave the normal range
        botT = mm.divn_nc(cbot * x.ue, Rm) * a2;
    } else if (GEN && p.bot_expl) botT = x.tb;
    else botT = mm.divn_nc(0.0, Rm) * a2;   // (bottom stress nothing: the reference's 0 / m_i * aice_i)
    if (GEN && p.top_sis) {
        const double d_own = x.ua - old, d_oth4 = x.oa - osum;
        ctop = p.top_rhoCd * mm.template sqrt_<false, false>(__fma_rn(d_oth4 * d_oth4, 0.0625, d_own * d_own));
        topT = mm.divn_nc(ctop * x.ua, Rm) * a2;
    }
    const double rheo2 = mm.divn_nc(mm.template div_nc<true>(vn - old, dtau), Ra);  // rheo / 2 (a checked quotient over alpha)
    const double d2 = nm.a * (sD1 - sD0);
    // sigma = validated input + checked increments (zero or >= 2^-352, < 2^310): the stress divergence needs no test;
    // whatever it adds up to, the velocity quotient below is tested
    const double tn = nm.t2hi * sT1 - nm.t2lo * sT0;
    const double tt2 = mm.divc_nc(NEG_TT ? -tn : tn, nm.td, nm.rtd);
    const double SS2 = mm.divc_nc(nm.s2hi * s12hi - nm.s2lo * s12lo, nm.sd, nm.rsd);
    const double dsig2 = mm.divc_nc(d2 + tt2 + SS2, nm.az, nm.raz);
    const double G = -cross - topT + botT + mm.divn_nc(dsig2, Rm) + (has_imm ? mm.divn(imm2, Rm) : 0.0) + __fma_rn(rheo2, 2.0, 0.0);
    const double tau = mm.divn_nc(cbot - ctop, Rm) * a2;  // (m2 <= 0 cannot pass the divisor window: no selects)
    const double D = mm.div(old + dtau * G, 1 + dtau * tau);
    return jl_mul_bool(active_ice ? D : ((GEN && p.fd_on) ? x.fd : 0.0), active);
}

// ---- the kernel ----------------------------------------------------------------------------
// Tile-local coordinates: (sx, sy) in [0,BX) x [0,BY) are the stress nodes; node (sx, sy) is the
// reference index (I0-1+sx, J0-1+sy), so the velocity cells of the tile are sx in [1,30], sy in [1,14].
// Shared arrays hold [-1, BX] x [-1, BY].  S(a, sx, sy) addresses them; SB(b, a, dx, dy) does the
// same relative to a node pointer b = &S(0, sx, sy), with compile-time offsets.
#define S(a, sx, sy) (sm[(a) * ASTRIDE + ((sy) + 1) * SXD + ((sx) + 1)])
#define SB(b, a, dx, dy) ((b)[(a) * ASTRIDE + (dy) * SXD + (dx)])

// pointwise inputs of one velocity node: previous-stage velocity, top-stress term, and the optional stage constants
struct Pt {
    double n, t, sa, tb, fd, rm, ue, sv;
    const double *oth;   // IEEE pass, top SemiImplicitStress: the node in the raw plane of the other atmosphere component
};
struct TileCtx {
    int I0, J0;  // reference index of the first velocity cell of the tile
    int fin, fout;
};

// All phases of one tile under one arithmetic policy.  Returns whether any thread left the policy's
// windows; global stores happen only when the pass is clean (the SLOW pass always is).
// GEN = false compiles the common configuration (SemiImplicitStress with velocity arrays, wind-stress
// arrays, FPlane, ReplacementPressure) without its run-time switches.
// MET: j-dependent metrics (lat-lon grid) read from the per-row table instead of the regular grid's constants.
// INTERIOR: the tile lies inside every store window, has no periodic image, wall neighbour or unevolved node (Params::it_x0):
// its instantiation carries none of the per-node edge logic.
// PH: 0 = the whole substep; 1 = phases A-C alone (stores the stresses, alpha and the first velocity), 2 = phase D alone (reads
// them back).  Meshes with a fold run their northernmost tile rows as 1, fold fill of the first velocity, 2: see fused_steps.
template <bool VFIRST, bool AUX, bool GEN, int MET, int PH, class M, bool INTERIOR>
__device__ __forceinline__ bool tile_pass(double *sm, uint64_t *bar, uint32_t parity, const CUtensorMap *tmap, const Params &p, const TileCtx &tc, int inv)
{
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    // lat-lon variant: the first-velocity array W shares the slot of e11 (dead after phase B) and W's own slot holds the
    // metric records of the tile's rows, copied once per pass from the per-row table
    constexpr int AW = MET == 1 ? A_E11 : A_W;
    const Metric<MET> mt{p, sm + A_W * ASTRIDE, tc.J0 - 3};
    if (MET == 1) {
        const double *src = p.met + (tc.J0 - 3) * MC_N;
        for (int n = tid; n < MET_ROWS * MC_N; n += NT) sm[A_W * ASTRIDE + n] = __ldg(src + n);
        __syncthreads();
    }
    if (MET == 2 && M::SCALED) {
        // two-dimensional metrics are read node by node through the read-only path in every phase: start the tile's rows of all
        // planes on their way from HBM to L2 now, behind the TMA loads below, so that those reads find them there
        const int xb = tc.I0 - 3 + OX, yb = tc.J0 - 3 + p.oy;   // the TMA box of the tile (even column: 16-byte aligned rows)
        for (int n = tid; n < MC2_N * SYD; n += NT) {
            const int k = n / SYD, row = n - k * SYD;
            bulk_prefetch_l2(p.met2 + k * p.met2_stride + (long long)(yb + row) * p.pitch + xb, SXD * (uint32_t)sizeof(double));
        }
    }
    const size_t plane = (size_t)p.pitch * p.rows;
    const bool use_ue = GEN ? p.use_ue != 0 : true, use_top = GEN ? p.use_top != 0 : true;
    // ---- TMA: tile + halo of every stencil field; u, v first (phase A starts on them) ----
    if (tid == 0) {
        const int x = tc.I0 - 2 - 1 + OX, y = tc.J0 - 2 - 1 + p.oy;
#ifdef CSI_EXPERIMENT_CONST_L2
        // timing experiment only (wrong results): every tile reads its stage constants from a 4 x 4-tile patch that stays in L2,
        // i.e. the kernel as it would run if those 72 B per cell-update never came from HBM
        const int xk = OX - 2 + OUTX * (blockIdx.x % 4), yk = p.oy - 3 + OUTY * (blockIdx.y % 4);   // (even column: 16-byte aligned box)
#else
        const int xk = x, yk = y;
#endif
        auto ld = [&](int arr, int field, uint64_t *b) {
            const bool cst = field == F_M || field == F_A || field == F_P || field == F_UE || field == F_VE;
            tma_load_row(sm + arr * ASTRIDE, tmap, b, cst ? xk : x, cst ? yk : y, field);
        };
        if (PH == 2) {
            // the second velocity phase alone: its own component of the previous substep, and what the A-C launch stored -- the
            // first velocity (with the fold applied since), the new stresses and alpha
            ld(VFIRST ? A_U : A_V, tc.fin + (VFIRST ? 0 : 1), &bar[0]);
            ld(AW, tc.fout + (VFIRST ? 1 : 0), &bar[0]);
        } else {
            ld(A_U, tc.fin + 0, &bar[0]);
            ld(A_V, tc.fin + 1, &bar[0]);
        }
        mbar_expect_tx(&bar[0], 2u * SXD * SYD * sizeof(double));
        int n = 6;
        ld(A_H, F_M, &bar[1]);  // ice mass (k_prep): the array is called A_H for historical reasons
        ld(A_A, F_A, &bar[1]);
        if (PH == 2) ld(A_AL, F_ALPHA, &bar[1]);
        else ld(A_P, F_P, &bar[1]);
        ld(A_S11, (PH == 2 ? tc.fout : tc.fin) + 2, &bar[1]);
        ld(A_S22, (PH == 2 ? tc.fout : tc.fin) + 3, &bar[1]);
        ld(A_S12, (PH == 2 ? tc.fout : tc.fin) + 4, &bar[1]);
        if (use_ue && !(M::SCALED && CSI_PRE_SVE)) {
            ld(A_UE, F_UE, &bar[1]);
            ld(A_VE, F_VE, &bar[1]);
            n += 2;
        }
        mbar_expect_tx(&bar[1], (uint32_t)n * SXD * SYD * sizeof(double));
#ifndef CSI_EXPERIMENT_NO_PREFETCH
        // pull the pointwise inputs of phases C / D towards L2 while the tile lands and A, B run
        tma_prefetch_box(tmap, xk, yk, F_UN);
        tma_prefetch_box(tmap, xk, yk, F_VN);
        if (M::SCALED ? (GEN ? p.use_t1 != 0 : true) : use_top) {
            tma_prefetch_box(tmap, xk, yk, M::SCALED ? F_T1X : F_TX);
            tma_prefetch_box(tmap, xk, yk, M::SCALED ? F_T1Y : F_TY);
        }
        if (M::SCALED && CSI_PRE_RMC) { tma_prefetch_box(tmap, x, y, F_RMC); tma_prefetch_box(tmap, x, y, F_RMF); }
        if (M::SCALED && CSI_PRE_PF4) tma_prefetch_box(tmap, x, y, F_PF4);
        if (M::SCALED && CSI_PRE_RM2) { tma_prefetch_box(tmap, x, y, F_RM2U); tma_prefetch_box(tmap, x, y, F_RM2V); }
        if (M::SCALED && CSI_PRE_SVE && use_ue) {
            tma_prefetch_box(tmap, x, y, F_UE); tma_prefetch_box(tmap, x, y, F_VE);
            tma_prefetch_box(tmap, x, y, F_SUE); tma_prefetch_box(tmap, x, y, F_SVE);
        }
#endif
    }
    // global offsets of the nodes this thread updates in phases C and D (32-wide rows: lane = column)
    // A warp owns two adjacent rows (q = 0, 1), so the pair sums of the 4-point means and the loads they need are shared.
    // C: odd  -> v at sx = lane (0..30), sy = 1 + 2 wrp + q (<= 15);  even -> u at sx = lane + 1 (1..31), sy = 2 wrp + q (<= 14)
    // D: sx = lane + 1 (1..30), sy = 1 + 2 wrp + q (<= 14)
    const int c_sx = VFIRST ? lane : lane + 1, c_sy0 = VFIRST ? 1 + 2 * wrp : 2 * wrp;
    const bool c_on[2] = {lane < OUTX + 1, lane < OUTX + 1 && c_sy0 + 1 <= (VFIRST ? OUTY + 1 : OUTY)};
    const int d_sx = lane + 1, d_sy0 = 1 + 2 * wrp;
    const bool d_on[2] = {lane < OUTX && d_sy0 <= OUTY, lane < OUTX && d_sy0 + 1 <= OUTY};
    // offsets inside one plane fit 32 bits (pitch x rows < 2^31 doubles is checked when the plan is built)
    const int o00 = (tc.J0 - 2 + p.oy) * p.pitch + (tc.I0 - 2 + OX);  // node (sx, sy) = (0, 0)
#ifdef CSI_EXPERIMENT_CONST_L2
    const int o00k = (p.oy - 2 + OUTY * (int)(blockIdx.y % 4)) * p.pitch + (OX - 1 + OUTX * (int)(blockIdx.x % 4));
#else
    const int o00k = o00;
#endif
    auto goff = [&](int sx, int sy) -> int {
        // clamped into the plane: edge tiles reach past the allocation (those nodes are never stored)
        const int row = min(max(tc.J0 - 1 + sy - 1 + p.oy, 0), p.rows - 1), col = min(max(tc.I0 - 1 + sx - 1 + OX, 0), p.pitch - 1);
        return row * p.pitch + col;
    };
    // previous-stage velocity and top-stress input of the two velocity phases: the FAST pass reads the precomputed term
    // tau_top / m_i * aice_i (k_prep), the IEEE pass the raw stress
    const bool use_t = M::SCALED ? (GEN ? p.use_t1 != 0 : true) : use_top;
    uint32_t c_g[2], d_g[2];
    if (INTERIOR) {
        c_g[0] = o00k + c_sy0 * p.pitch + c_sx;
        d_g[0] = o00k + d_sy0 * p.pitch + d_sx;
        c_g[1] = c_g[0] + p.pitch;
        d_g[1] = d_g[0] + p.pitch;
    } else {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            c_g[q] = goff(c_sx, c_sy0 + q);
            d_g[q] = goff(d_sx, d_sy0 + q);
        }
    }

    // immersed-boundary node flags of the tile (bit 0: centre masked, 1: corner masked, 2: u face peripheral, 3: v face peripheral)
    uint8_t *smf = reinterpret_cast<uint8_t *>(bar + 8);
    const bool has_mask = GEN && p.flags != nullptr;
    if (has_mask)
        for (int n = tid; n < SXD * SYD; n += NT) {
            const int row = min(max(tc.J0 - 2 + n / SXD - 1 + p.oy, 0), p.rows - 1), col = min(max(tc.I0 - 2 + n % SXD - 1 + OX, 0), p.pitch - 1);
            smf[n] = p.flags[(size_t)row * p.pitch + col];
        }
#define FLG(sx, sy) (smf[((sy) + 1) * SXD + ((sx) + 1)])

    M mm;
    // ---------------- phase A: strain rates (evp:360-375), ice mass (ClimaSeaIce.jl:42) ----------------
    // e11, e22 on [-1, BX-1] x [-1, BY-1]; e12 on [0, BX] x [0, BY]; m everywhere (in place over h)
    mbar_wait(&bar[0], parity);
    if (PH == 2) mbar_wait(&bar[1], parity);
    double aux_zc[2], aux_zf[2], aux_Dc[2];
    if (PH != 2) {
    if (M::SCALED && !MET && p.sq) {
        // regular grid with dx == dy: u / dx and v / dx serve both the tension and the shear operator; divide every
        // u, v of the tile once (into the arrays phases B and C fill later) instead of eight times per node
        double *qu = sm + A_AL * ASTRIDE, *qv = sm + A_W * ASTRIDE;
        // (the element loops of this phase are unrolled: NIT - 1 full sweeps of the CTA and a partial one)
#pragma unroll
        for (int k = 0; k < NIT; k++) {
            const int n = tid + k * NT;
            if ((k + 1) * NT > SXD * SYD && n >= SXD * SYD) break;
            // (u, v are validated inputs or checked quotients of the previous substep: zero or in [2^-300, 2^300))
            qu[n] = mm.divc_nc(sm[A_U * ASTRIDE + n], p.dx, p.rdx);
            qv[n] = mm.divc_nc(sm[A_V * ASTRIDE + n], p.dx, p.rdx);
        }
        __syncthreads();
        // Every element of the haloed tile gets all three strain rates, without edge tests: e11, e22 of the last column / row
        // and e12 of the first ones are built from neighbours outside the tile (whatever the adjacent shared memory holds),
        // are never read by phase B, and these unchecked operators raise no window flag
#pragma unroll
        for (int k = 0; k < NIT; k++) {
            const int n = tid + k * NT;
            if ((k + 1) * NT > SXD * SYD && n >= SXD * SYD) break;
            double *b = sm + n;
            const double u00 = SB(b, A_U, 0, 0), v00 = SB(b, A_V, 0, 0), qu00 = SB(b, A_AL, 0, 0), qv00 = SB(b, A_W, 0, 0);
            // (the numerators are built from checked quotients and bounded metrics: no second window test)
            const double D = mm.divc_nc((p.dy * SB(b, A_U, 1, 0) - p.dy * u00) + (p.dx * SB(b, A_V, 0, 1) - p.dx * v00), p.az, p.raz);
            const double T = mm.divc_nc(p.dy2 * (SB(b, A_AL, 1, 0) - qu00) - p.dx2 * (SB(b, A_W, 0, 1) - qv00), p.az, p.raz);
            SB(b, A_E11, 0, 0) = D + T;  // 2 e11
            SB(b, A_E22, 0, 0) = D - T;  // 2 e22
            SB(b, A_E12, 0, 0) = mm.divc_nc(p.dx2 * (qu00 - SB(b, A_AL, 0, -1)) + p.dy2 * (qv00 - SB(b, A_W, -1, 0)), p.az, p.raz);  // 2 e12
        }
    } else {
    int sx = tid % SXD - 1, sy = tid / SXD - 1;
#pragma unroll
    for (int k = 0; k < NIT; k++, sx += NT % SXD, sy += NT / SXD) {
        const int n = tid + k * NT;
        if ((k + 1) * NT > SXD * SYD && n >= SXD * SYD) break;
        if (sx >= SXD - 1) { sx -= SXD; sy++; }
        const int r = tc.J0 - 1 + sy;  // reference row of the node
        const int o = o00 + sy * p.pitch + sx, op = o + p.pitch, om = o - p.pitch;  // in-plane offsets: the node, its north / south neighbour
        double *b = sm + n;  // = &S(0, sx, sy)
        const double u00 = SB(b, A_U, 0, 0), v00 = SB(b, A_V, 0, 0);
        if (sx < BX && sy < BY) {
            const double u10 = SB(b, A_U, 1, 0), v01 = SB(b, A_V, 0, 1);
            // (on a lat-lon grid the two dy^fc are the same number)
            const double dyf0 = mt.dyfc(o, r), rdyf0 = mt.rdyfc(o, r), dyf1 = mt.dyfc(o + 1, r), rdyf1 = mt.rdyfc(o + 1, r);
            const double dxf0 = mt.dxcf(o, r), dxf1 = mt.dxcf(op, r + 1), az = mt.azcc(o, r), raz = mt.razcc(o, r);
            const double D = mm.divc_nc((dyf1 * u10 - dyf0 * u00) + (dxf1 * v01 - dxf0 * v00), az, raz);
            // (u, v are validated inputs or tested quotients of the previous substep: no window test on u / metric)
            const double T = mm.divc_nc(mt.dycc2(o, r) * (mm.divc_nc(u10, dyf1, rdyf1) - mm.divc_nc(u00, dyf0, rdyf0)) -
                                         mt.dxcc2(o, r) * (mm.divc_nc(v01, dxf1, mt.rdxcf(op, r + 1)) - mm.divc_nc(v00, dxf0, mt.rdxcf(o, r))),
                                     az, raz);
            SB(b, A_E11, 0, 0) = M::SCALED ? D + T : (D + T) / 2;
            SB(b, A_E22, 0, 0) = M::SCALED ? D - T : (D - T) / 2;
        }
        if (sx >= 0 && sy >= 0) {
            const double u0m = SB(b, A_U, 0, -1), vm0 = SB(b, A_V, -1, 0);
            const double dyc0 = mt.dycf(o, r), rdyc0 = mt.rdycf(o, r), dycm = mt.dycf(o - 1, r), rdycm = mt.rdycf(o - 1, r);
            const double Sh = mm.divc_nc(mt.dxff2(o, r) * (mm.divc_nc(u00, mt.dxfc(o, r), mt.rdxfc(o, r)) - mm.divc_nc(u0m, mt.dxfc(om, r - 1), mt.rdxfc(om, r - 1))) +
                                          mt.dyff2(o, r) * (mm.divc_nc(v00, dyc0, rdyc0) - mm.divc_nc(vm0, dycm, rdycm)),
                                      mt.azff(o, r), mt.razff(o, r));
            SB(b, A_E12, 0, 0) = M::SCALED ? Sh : Sh / 2;
        }
    }
    }
    __syncthreads();
    mbar_wait(&bar[1], parity);  // (the ice mass m = h rho aice arrives precomputed: k_prep)

    // ---------------- phase B: viscosities + stress update (evp:236-354), nodes [0,BX) x [0,BY) --------
    constexpr bool PRE_RMC = M::SCALED && CSI_PRE_RMC, PRE_PF4 = M::SCALED && CSI_PRE_PF4;
    double b_rmc[2] = {0.0, 0.0}, b_rmf[2] = {0.0, 0.0}, b_pf4[2] = {0.0, 0.0};
    if (PRE_RMC || PRE_PF4) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const uint32_t g = INTERIOR ? (uint32_t)(o00 + (2 * wrp + q) * p.pitch + lane) : (uint32_t)goff(lane, 2 * wrp + q);
            if (PRE_RMC) { b_rmc[q] = __ldg(p.g_rmc + g); b_rmf[q] = __ldg(p.g_rmf + g); }
            if (PRE_PF4) b_pf4[q] = __ldg(p.g_pf4 + g);
        }
    }
#pragma unroll UNROLL_B
    for (int q = 0; q < 2; q++) {
        const int sx = lane, sy = 2 * wrp + q;
        const int rB = tc.J0 - 1 + sy, oB = o00 + sy * p.pitch + sx;
        double *b = &S(0, sx, sy);
        double zc, zf, Dc, s11n, s22n, s12n, mc, mf, g2c, g2f;
        bool mc0 = false, mf0 = false;
        if (M::SCALED) {
            // a, bb = 2 e11c, 2 e22c; Sh = 2 e12f; c4 = 4 e12c; af, bf = 8 e11f, 8 e22f
            const double a = SB(b, A_E11, 0, 0), bb = SB(b, A_E22, 0, 0), Sh = SB(b, A_E12, 0, 0);
            const double c4 = ((Sh + SB(b, A_E12, 1, 0)) + (SB(b, A_E12, 0, 1) + SB(b, A_E12, 1, 1))) * 0.5;
            const double af = (SB(b, A_E11, -1, -1) + SB(b, A_E11, 0, -1)) + (SB(b, A_E11, -1, 0) + a);
            const double bf = (SB(b, A_E22, -1, -1) + SB(b, A_E22, 0, -1)) + (SB(b, A_E22, -1, 0) + bb);
            const double dc2 = a + bb, df8 = af + bf, Sh8 = 8 * Sh;
            const double sc2 = mm.sqrt_((a - bb) * (a - bb) + c4 * c4);        // 2 sc
            const double sf8 = mm.sqrt_((af - bf) * (af - bf) + Sh8 * Sh8);    // 8 sf
            const double Dc2 = mm.max_pos(mm.sqrt_(dc2 * dc2 + sc2 * sc2 * p.em2), p.Dmin2);  // 2 Delta_c
            const double Df8 = mm.max_pos(mm.sqrt_(df8 * df8 + sf8 * sf8 * p.em2), p.Dmin8);  // 8 Delta_f
            const double Pc = SB(b, A_P, 0, 0);
            const double Pf4 = PRE_PF4 ? b_pf4[q] : (SB(b, A_P, -1, -1) + SB(b, A_P, 0, -1)) + (SB(b, A_P, -1, 0) + Pc);
            // (Delta in [Delta_min, 2^257) follows from its checked radicand: divisor range tests only where unknown)
            // P is a validated input (zero or in [2^-300, 2^300)): these quotients cannot leave the normal range
            zf = mm.template div_nc<true>(Pf4, Df8);
            zc = mm.template div_nc<true>(Pc, Dc2);
            const double Pr = (GEN && p.pform == CSI_ICE_STRENGTH) ? Pc : mm.template div_nc<true>(Pc * Dc2, Dc2 + p.Dmin2);
            const double ec = zc * p.em2, ef = zf * p.em2;
            const double X = (zc - ec) * dc2 - Pr;  // 2 ((zeta - eta)(e11 + e22) - Pr / 2)
            s11n = __fma_rn(X, 0.5, ec * a);
            s22n = __fma_rn(X, 0.5, ec * bb);
            s12n = ef * Sh;
            mc = SB(b, A_H, 0, 0);
            mf = (SB(b, A_H, -1, -1) + SB(b, A_H, 0, -1)) + (SB(b, A_H, -1, 0) + mc);  // 4 mf
            if (!INTERIOR) {
                // nodes beyond a wall's ring of stress nodes (the caller's deeper halo holds no ice there) are computed but
                // never stored nor read by a stored cell: give them a harmless mass instead of failing the tile's divisor test
                const int i = tc.I0 - 1 + sx;
                if ((p.wall_w && i < p.sx0) || (p.wall_e && i > p.sx1) || (p.wall_s && rB < p.sy0) || (p.wall_n && rB > p.sy1)) {
                    mc = 1.0; mf = 4.0;
                    b_rmc[q] = 1.0; b_rmf[q] = 0.25;
                }
            }
            // open water (mass exactly 0, P >= 0): the reference divides by zero -- zeta / 0 = +inf, or NaN when zeta = 0 too, which
            // it replaces by alpha+^2 -- and clamps the result to alpha+ (p.gnan when it came from the NaN branch), and it
            // leaves sigma alone.  Same values here without dividing by zero, so open water does not send the tile to the IEEE pass.
            mc0 = mc == 0.0 && Pc >= 0.0;
            mf0 = mf == 0.0 && Pf4 >= 0.0;
            if (PRE_RMC) {
                // reciprocals from k_prep: 1 resp. 1/4 in open water, +inf where the divisor would have failed its range test
                // (the quotient then fails its window test)
                g2c = mm.divc_nc(mm.divc(zc * p.ca * p.dt, mc0 ? 1.0 : mc, b_rmc[q]), mt.azcc(oB, rB), mt.razcc(oB, rB));
                g2f = mm.divc_nc(mm.divc(zf * p.ca * p.dt4, mf0 ? 4.0 : mf, b_rmf[q]), mt.azff(oB, rB), mt.razff(oB, rB));
            } else {
                g2c = mm.divc_nc(mm.div(zc * p.ca * p.dt, mc0 ? 1.0 : mc), mt.azcc(oB, rB), mt.razcc(oB, rB));
                g2f = mm.divc_nc(mm.div(zf * p.ca * p.dt4, mf0 ? 4.0 : mf), mt.azff(oB, rB), mt.razff(oB, rB));
            }
            Dc = AUX ? Dc2 * 0.5 : 0.0;
        } else {
        const double e11c = SB(b, A_E11, 0, 0), e22c = SB(b, A_E22, 0, 0), e12f = SB(b, A_E12, 0, 0);
        const double e12c = ((e12f + SB(b, A_E12, 1, 0)) / 2 + (SB(b, A_E12, 0, 1) + SB(b, A_E12, 1, 1)) / 2) / 2;
        const double e11f = ((SB(b, A_E11, -1, -1) + SB(b, A_E11, 0, -1)) / 2 + (SB(b, A_E11, -1, 0) + e11c) / 2) / 2;
        const double e22f = ((SB(b, A_E22, -1, -1) + SB(b, A_E22, 0, -1)) / 2 + (SB(b, A_E22, -1, 0) + e22c) / 2) / 2;
        const double dc = e11c + e22c, df = e11f + e22f;
        const double sc = mm.sqrt_((e11c - e22c) * (e11c - e22c) + 4 * (e12c * e12c));
        const double sf = mm.sqrt_((e11f - e22f) * (e11f - e22f) + 4 * (e12f * e12f));
        Dc = jl_max(mm.sqrt_(dc * dc + sc * sc * p.em2), p.Dmin);
        const double Df = jl_max(mm.sqrt_(df * df + sf * sf * p.em2), p.Dmin);
        const double Pc = SB(b, A_P, 0, 0);
        const double Pf = ((SB(b, A_P, -1, -1) + SB(b, A_P, 0, -1)) / 2 + (SB(b, A_P, -1, 0) + Pc) / 2) / 2;
        zf = mm.div(Pf, 2 * Df);
        zc = mm.div(Pc, 2 * Dc);
        const double Pr = (GEN && p.pform == CSI_ICE_STRENGTH) ? Pc : mm.div(Pc * Dc, Dc + p.Dmin);
        const double ec = zc * p.em2, ef = zf * p.em2;
        s11n = 2 * ec * e11c + ((zc - ec) * (e11c + e22c) - Pr / 2);
        s22n = 2 * ec * e22c + ((zc - ec) * (e11c + e22c) - Pr / 2);
        s12n = 2 * ef * e12f;
        mc = SB(b, A_H, 0, 0);
        mf = ((SB(b, A_H, -1, -1) + SB(b, A_H, 0, -1)) / 2 + (SB(b, A_H, -1, 0) + mc) / 2) / 2;
        g2c = mm.divc(mm.div(zc * p.ca * p.dt, mc), mt.azcc(oB, rB), mt.razcc(oB, rB));
        g2f = mm.divc(mm.div(zf * p.ca * p.dt, mf), mt.azff(oB, rB), mt.razff(oB, rB));
        }
        // (a clean FAST pass has no NaN quotient and no mass <= 0: both would have left the windows)
        if (!M::SCALED) g2c = (g2c != g2c) ? p.amax2 : g2c;
        double gc = jl_clamp(mm.template sqrt_<!M::SCALED>(g2c), p.amin, p.amax);
        if (!M::SCALED) g2f = (g2f != g2f) ? p.amax2 : g2f;
        double gf = jl_clamp(mm.template sqrt_<!M::SCALED>(g2f), p.amin, p.amax);
        if (M::SCALED) {
            gc = mc0 ? (zc == 0.0 ? p.gnan : p.amax) : gc;
            gf = mf0 ? (zf == 0.0 ? p.gnan : p.amax) : gf;
        }
        const NodeRecip Rg = mm.template recip<M::SCALED>(gc);  // gamma in [alpha-, alpha+]
        const double o11 = SB(b, A_S11, 0, 0), o22 = SB(b, A_S22, 0, 0), o12 = SB(b, A_S12, 0, 0);
        const double d11 = mm.divn(s11n - o11, Rg), d22 = mm.divn(s22n - o22, Rg), d12 = mm.template div<M::SCALED>(s12n - o12, gf);
        // in place: each thread owns its node of the sigma arrays
        SB(b, A_S11, 0, 0) = o11 + ((M::SCALED ? !mc0 : mc > 0) ? d11 : 0.0);
        SB(b, A_S22, 0, 0) = o22 + ((M::SCALED ? !mc0 : mc > 0) ? d22 : 0.0);
        SB(b, A_S12, 0, 0) = o12 + ((M::SCALED ? !mf0 : mf > 0) ? d12 : 0.0);
        SB(b, A_AL, 0, 0) = gc;
        aux_zc[q] = zc;
        aux_zf[q] = zf;
        aux_Dc[q] = Dc;
    }
    }  // (PH != 2)
    // the pointwise inputs of phase C are requested before the barrier, so their L2 latency overlaps the wait
    constexpr bool PRE_SVE = M::SCALED && CSI_PRE_SVE, PRE_RM2 = M::SCALED && CSI_PRE_RM2;
    auto load_pt = [&](bool on, uint32_t g, const PhasePtrs &pp, double tconst, double bconst) {
        Pt pt;  // (members a configuration does not use stay unset and cost nothing)
        pt.n = on ? __ldg(pp.n + g) : 0.0;
        const double *gt = M::SCALED ? pp.t1 : pp.tt;
        const bool top_sis = GEN && p.top_sis;
        pt.t = (on && (top_sis ? p.top_arr != 0 : use_t)) ? __ldg(gt + g) : (M::SCALED ? 0.0 : tconst);
        if (GEN && p.top_sis) {
            if (M::SCALED) pt.sa = (on && p.top_arr) ? __ldg(pp.sa + g) : 0.0;
            else pt.oth = pp.oth + g;
        }
        if (GEN && p.bot_expl) pt.tb = M::SCALED ? (on ? __ldg(pp.tb1 + g) : 0.0) : ((on && p.bot_arr) ? __ldg(pp.tbr + g) : bconst);
        if (GEN && p.fd_on) pt.fd = on ? __ldg(pp.fd + g) : 0.0;
        if (PRE_RM2) pt.rm = on ? __ldg(pp.rm + g) : 1.0;
        if (PRE_SVE) pt.ue = (on && use_ue) ? __ldg(pp.ue + g) : 0.0;
        if (PRE_SVE) pt.sv = (on && use_ue) ? __ldg(pp.sv + g) : 0.0;
        return pt;
    };
    // the metric factors of a u / v node's stress divergence (isd:39-51).  (Requesting the two-dimensional ones ahead of the barrier
    // that precedes a velocity phase, like the pointwise inputs, was measured: 8.27 -> 7.53e9 cell-updates/s -- the 22 values held
    // across the barrier spill.)
    auto u_metric = [&](int sx, int sy) -> NodeMetric {
        const int r = tc.J0 - 1 + sy, o = o00 + sy * p.pitch + sx, op = o + p.pitch;
        const double dyc2 = mt.dycc2(o, r), dyc2w = mt.dycc2(o - 1, r), dyf = mt.dyfc(o, r);  // (dyc2w == dyc2 unless the metrics depend on i)
        return NodeMetric{dyf, dyc2, dyc2w, dyf, mt.rdyfc(o, r), M::SCALED ? mt.dxff2d(op, r + 1) : mt.dxff2(op, r + 1), M::SCALED ? mt.dxff2d(o, r) : mt.dxff2(o, r),
                          mt.dxfc(o, r), mt.rdxfc(o, r), mt.azfc(o, r), mt.razfc(o, r)};
    };
    auto v_metric = [&](int sx, int sy) -> NodeMetric {
        const int r = tc.J0 - 1 + sy, o = o00 + sy * p.pitch + sx, om = o - p.pitch;
        const double dyf2 = M::SCALED ? mt.dyff2d(o, r) : mt.dyff2(o, r), dyf2e = M::SCALED ? mt.dyff2d(o + 1, r) : mt.dyff2(o + 1, r), dxf = mt.dxcf(o, r);
        return NodeMetric{dxf, mt.dxcc2(o, r), mt.dxcc2(om, r - 1), dxf, mt.rdxcf(o, r), dyf2e, dyf2, mt.dycf(o, r), mt.rdycf(o, r), mt.azcf(o, r), mt.razcf(o, r)};
    };
    Pt cpt[2];
    if (PH != 2) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            cpt[q] = load_pt(c_on[q], c_g[q], p.pc, VFIRST ? p.tty : p.ttx, VFIRST ? p.tb_y : p.tb_x);
        }
    }
    __syncthreads();

    // alpha at a neighbour of a velocity node.  Phase D alone reads it back from its plane: nodes of the box beyond the plane
    // arrive as zeros (TMA), where the whole substep would hold a computed alpha -- none of them feeds a stored cell, but a zero
    // divisor must not send the tile to the IEEE pass
    auto AL = [&](const double *b, int off) -> double {
        const double v = b[A_AL * ASTRIDE + off];
        return (PH == 2 && v == 0.0) ? p.amin : v;
    };
    // u at node (sx, sy); VS = array holding the v it reads (old v, or the first-velocity array)
    auto u_at = [&](int sx, int sy, int VS, const Pt &pt, const NodeMetric &nm) -> double {
        const double un = pt.n, ttop = pt.t;
        const int i = tc.I0 - 1 + sx, r = tc.J0 - 1 + sy;
        const int o = o00 + sy * p.pitch + sx, op = o + p.pitch;
        bool upd = true, wall_active = true;
        if (!INTERIOR) {
            upd = r >= p.cy0 && r <= p.cy1 && i >= p.cx0 && i <= p.cx1;
            wall_active = !((p.wall_w && i <= 1) || (p.wall_e && i > p.Nx));
        }
        const double *b = &S(0, sx, sy);
        // reference tree: vbar; scaled tree: the plain sum 4 vbar (and f / 4)
        const double vsum = (SB(b, VS, -1, 0) + SB(b, VS, 0, 0)) + (SB(b, VS, -1, 1) + SB(b, VS, 0, 1));
        const double vbar = M::SCALED ? vsum : ((SB(b, VS, -1, 0) + SB(b, VS, 0, 0)) / 2 + (SB(b, VS, -1, 1) + SB(b, VS, 0, 1)) / 2) / 2;
        // (the switch-free variants: FPlane on regular grids, HydrostaticSphericalCoriolis on lat-lon grids)
        double xcross = ((GEN && p.cor == CSI_CORIOLIS_NONE) || (MET && !GEN)) ? 0.0 : (M::SCALED ? -p.f4 * vsum : -p.f * vbar);
        if (MET == 1 && (!GEN || p.cor == CSI_CORIOLIS_SPHERICAL)) {  // -Iy(f^ff) * Ix(Iy(dx^cf v)) / dx^fc (lat-lon grids only)
            const double dx0 = mt.dxcf(o, r), dx1 = mt.dxcf(op, r + 1);
            const double gm = (dx0 * SB(b, VS, -1, 0) + dx1 * SB(b, VS, -1, 1)) / 2, g0 = (dx0 * SB(b, VS, 0, 0) + dx1 * SB(b, VS, 0, 1)) / 2;
            const double fbar = (mt.fff(r) + mt.fff(r + 1)) / 2;
            xcross = mm.divc(-fbar * ((gm + g0) / 2), mt.dxfc(o, r), mt.rdxfc(o, r));
        }
        double ue = p.ue_c, vebar = M::SCALED ? (p.ve_c + p.ve_c) + (p.ve_c + p.ve_c) : ((p.ve_c + p.ve_c) / 2 + (p.ve_c + p.ve_c) / 2) / 2;
        if (PRE_SVE && use_ue) {
            ue = pt.ue;
            vebar = pt.sv;
        } else if (use_ue) {
            ue = SB(b, A_UE, 0, 0);
            vebar = M::SCALED ? (SB(b, A_VE, -1, 0) + SB(b, A_VE, 0, 0)) + (SB(b, A_VE, -1, 1) + SB(b, A_VE, 0, 1))
                              : ((SB(b, A_VE, -1, 0) + SB(b, A_VE, 0, 0)) / 2 + (SB(b, A_VE, -1, 1) + SB(b, A_VE, 0, 1)) / 2) / 2;
        }
        const double uold = SB(b, A_U, 0, 0);
        double a1 = SB(b, A_S11, 0, 0), b1 = SB(b, A_S22, 0, 0), a0 = SB(b, A_S11, -1, 0), b0 = SB(b, A_S22, -1, 0);
        double s12hi = SB(b, A_S12, 0, 1), s12lo = SB(b, A_S12, 0, 0);
        bool active = wall_active;
        if (has_mask) {  // isd:21-24: stresses read as 0 on immersed-peripheral nodes; se:226: peripheral u faces stay at rest
            if (FLG(sx, sy) & 1) a1 = b1 = 0.0;
            if (FLG(sx - 1, sy) & 1) a0 = b0 = 0.0;
            if (FLG(sx, sy + 1) & 2) s12hi = 0.0;
            if (FLG(sx, sy) & 2) s12lo = 0.0;
            active = active && !(FLG(sx, sy) & 4);
        }
        // immersed stress divergence (isd:65-82) for the linear-drag flux BC -C*u on south/north immersed faces
        const bool has_imm = has_mask && p.imm_u != 0.0;
        double imm = 0.0;
        if (has_imm) {
            const double bc = (-p.imm_u) * uold;
            const double qW = 0.0 * (mt.dycc(o - 1, r) * 1.0), qE = 0.0 * (mt.dycc(o, r) * 1.0);
            const double qS = ((FLG(sx, sy) & 2) ? -bc : 0.0) * (mt.dxff(o, r) * 1.0);
            const double qN = ((FLG(sx, sy + 1) & 2) ? bc : 0.0) * (mt.dxff(op, r + 1) * 1.0);
            imm = mm.divc(qE - qW + qN - qS, mt.azfc(o, r) * 1.0, mt.razfc(o, r));
        }
        Ext x;
        x.ue = ue; x.oe = vebar; x.t1 = ttop;
        if (GEN && p.top_sis) {  // (u_a, v_a): arrays read pointwise, the 4-point sum of v_a from k_prep / its mean from the raw plane
            x.t1 = 0.0;
            x.ua = p.top_arr ? pt.t : p.ta_x;
            if (M::SCALED) x.oa = p.top_arr ? pt.sa : (p.ta_y + p.ta_y) + (p.ta_y + p.ta_y);
            else if (p.top_arr) x.oa = ((__ldg(pt.oth - 1) + __ldg(pt.oth)) / 2 + (__ldg(pt.oth + p.pitch - 1) + __ldg(pt.oth + p.pitch)) / 2) / 2;
            else x.oa = ((p.ta_y + p.ta_y) / 2 + (p.ta_y + p.ta_y) / 2) / 2;
        }
        if (GEN && p.bot_expl) x.tb = pt.tb;
        if (GEN && p.fd_on) x.fd = pt.fd;
        double val;
        if (M::SCALED) {
            // (a node that is not evolved returns its old value whatever is computed: harmless masses keep it from failing the tile)
            const double m1 = upd ? SB(b, A_H, 0, 0) : 1.0, m0 = upd ? SB(b, A_H, -1, 0) : 1.0;
            val = vel_node_s<GEN, false>(mm, p, nm, active, m1, m0, SB(b, A_A, 0, 0), SB(b, A_A, -1, 0), AL(b, 0), AL(b, -1),
                                         uold, vbar, xcross, x, un, a1 + b1, a0 + b0, a1 - b1, a0 - b0, s12hi, s12lo, has_imm, 2 * imm, upd ? pt.rm : 0.5);
        } else {
            val = vel_node<GEN, false>(mm, p, nm, active, SB(b, A_H, 0, 0), SB(b, A_H, -1, 0), SB(b, A_A, 0, 0), SB(b, A_A, -1, 0), AL(b, 0), AL(b, -1),
                                       uold, vbar, xcross, x, un, a1 + b1, a0 + b0, a1 - b1, a0 - b0, s12hi, s12lo, has_imm, imm);
        }
        return upd ? val : uold;
    };
    auto v_at = [&](int sx, int sy, int US, const Pt &pt, const NodeMetric &nm) -> double {
        const double vn = pt.n, ttop = pt.t;
        const int i = tc.I0 - 1 + sx, r = tc.J0 - 1 + sy;
        const int o = o00 + sy * p.pitch + sx, om = o - p.pitch;
        bool upd = true, wall_active = true;
        if (!INTERIOR) {
            upd = r >= p.cy0 && r <= p.cy1 && i >= p.cx0 && i <= p.cx1;
            wall_active = !((p.wall_s && r <= 1) || (p.wall_n && r > p.Ny));
        }
        const double *b = &S(0, sx, sy);
        const double usum = (SB(b, US, 0, -1) + SB(b, US, 1, -1)) + (SB(b, US, 0, 0) + SB(b, US, 1, 0));
        const double ubar = M::SCALED ? usum : ((SB(b, US, 0, -1) + SB(b, US, 1, -1)) / 2 + (SB(b, US, 0, 0) + SB(b, US, 1, 0)) / 2) / 2;
        double ycross = ((GEN && p.cor == CSI_CORIOLIS_NONE) || (MET && !GEN)) ? 0.0 : (M::SCALED ? p.f4 * usum : p.f * ubar);
        if (MET == 1 && (!GEN || p.cor == CSI_CORIOLIS_SPHERICAL)) {  // +Ix(f^ff) * Iy(Ix(dy^fc u)) / dy^cf (lat-lon grids only)
            const double dy0 = mt.dyfc(om, r - 1), dy1 = mt.dyfc(o, r);
            const double gm = (dy0 * SB(b, US, 0, -1) + dy0 * SB(b, US, 1, -1)) / 2, g0 = (dy1 * SB(b, US, 0, 0) + dy1 * SB(b, US, 1, 0)) / 2;
            const double fj = mt.fff(r), fbar = (fj + fj) / 2;
            ycross = mm.divc(fbar * ((gm + g0) / 2), mt.dycf(o, r), mt.rdycf(o, r));
        }
        double ve = p.ve_c, uebar = M::SCALED ? (p.ue_c + p.ue_c) + (p.ue_c + p.ue_c) : ((p.ue_c + p.ue_c) / 2 + (p.ue_c + p.ue_c) / 2) / 2;
        if (PRE_SVE && use_ue) {
            ve = pt.ue;
            uebar = pt.sv;
        } else if (use_ue) {
            ve = SB(b, A_VE, 0, 0);
            uebar = M::SCALED ? (SB(b, A_UE, 0, -1) + SB(b, A_UE, 1, -1)) + (SB(b, A_UE, 0, 0) + SB(b, A_UE, 1, 0))
                              : ((SB(b, A_UE, 0, -1) + SB(b, A_UE, 1, -1)) / 2 + (SB(b, A_UE, 0, 0) + SB(b, A_UE, 1, 0)) / 2) / 2;
        }
        const double vold = SB(b, A_V, 0, 0);
        double a1 = SB(b, A_S11, 0, 0), b1 = SB(b, A_S22, 0, 0), a0 = SB(b, A_S11, 0, -1), b0 = SB(b, A_S22, 0, -1);
        double s12hi = SB(b, A_S12, 1, 0), s12lo = SB(b, A_S12, 0, 0);
        bool active = wall_active;
        if (has_mask) {
            if (FLG(sx, sy) & 1) a1 = b1 = 0.0;
            if (FLG(sx, sy - 1) & 1) a0 = b0 = 0.0;
            if (FLG(sx + 1, sy) & 2) s12hi = 0.0;
            if (FLG(sx, sy) & 2) s12lo = 0.0;
            active = active && !(FLG(sx, sy) & 8);
        }
        const bool has_imm = has_mask && p.imm_v != 0.0;
        double imm = 0.0;
        if (has_imm) {  // isd:84-101, -C*v on west/east immersed faces
            const double bc = (-p.imm_v) * vold;
            const double qW = ((FLG(sx, sy) & 2) ? -bc : 0.0) * (mt.dyff(o, r) * 1.0);
            const double qE = ((FLG(sx + 1, sy) & 2) ? bc : 0.0) * (mt.dyff(o + 1, r) * 1.0);
            const double qS = 0.0 * (mt.dxcc(om, r - 1) * 1.0), qN = 0.0 * (mt.dxcc(o, r) * 1.0);
            imm = mm.divc(qE - qW + qN - qS, mt.azcf(o, r) * 1.0, mt.razcf(o, r));
        }
        Ext x;
        x.ue = ve; x.oe = uebar; x.t1 = ttop;
        if (GEN && p.top_sis) {
            x.t1 = 0.0;
            x.ua = p.top_arr ? pt.t : p.ta_y;
            if (M::SCALED) x.oa = p.top_arr ? pt.sa : (p.ta_x + p.ta_x) + (p.ta_x + p.ta_x);
            else if (p.top_arr) x.oa = ((__ldg(pt.oth - p.pitch) + __ldg(pt.oth - p.pitch + 1)) / 2 + (__ldg(pt.oth) + __ldg(pt.oth + 1)) / 2) / 2;
            else x.oa = ((p.ta_x + p.ta_x) / 2 + (p.ta_x + p.ta_x) / 2) / 2;
        }
        if (GEN && p.bot_expl) x.tb = pt.tb;
        if (GEN && p.fd_on) x.fd = pt.fd;
        const double val = M::SCALED ? vel_node_s<GEN, true>(mm, p, nm, active, upd ? SB(b, A_H, 0, 0) : 1.0, upd ? SB(b, A_H, 0, -1) : 1.0, SB(b, A_A, 0, 0), SB(b, A_A, 0, -1), AL(b, 0),
                                                             AL(b, -SXD), vold, ubar, ycross, x, vn, a1 + b1, a0 + b0, a1 - b1, a0 - b0, s12hi, s12lo, has_imm, 2 * imm, upd ? pt.rm : 0.5)
                                     : vel_node<GEN, true>(mm, p, nm, active, SB(b, A_H, 0, 0), SB(b, A_H, 0, -1), SB(b, A_A, 0, 0), SB(b, A_A, 0, -1), AL(b, 0),
                                                           AL(b, -SXD), vold, ubar, ycross, x, vn, a1 + b1, a0 + b0, a1 - b1, a0 - b0, s12hi, s12lo, has_imm, imm);
        return upd ? val : vold;
    };

    // ---------------- phase C: first velocity on the cells the second one reads -------------------------
    if (PH != 2) {
#pragma unroll UNROLL_CD
        for (int q = 0; q < 2; q++)
            if (c_on[q]) {
                const int sy = c_sy0 + q;
                const NodeMetric nm = VFIRST ? v_metric(c_sx, sy) : u_metric(c_sx, sy);
                S(AW, c_sx, sy) = VFIRST ? v_at(c_sx, sy, A_U, cpt[q], nm) : u_at(c_sx, sy, A_V, cpt[q], nm);
            }
    }
    Pt dpt[2];  // likewise for phase D
    if (PH != 1) {
#pragma unroll
        for (int q = 0; q < 2; q++) {
            dpt[q] = load_pt(d_on[q], d_g[q], p.pd, VFIRST ? p.ttx : p.tty, VFIRST ? p.tb_x : p.tb_y);
        }
    }
    __syncthreads();

    // ---------------- phase D: second velocity on the output cells [1,30] x [1,14] ----------------------
    double w2[2] = {0.0, 0.0};
    if (PH != 1) {
#pragma unroll UNROLL_CD
        for (int q = 0; q < 2; q++)
            if (d_on[q]) {
                const int sy = d_sy0 + q;
                const NodeMetric nm = VFIRST ? u_metric(d_sx, sy) : v_metric(d_sx, sy);
                w2[q] = VFIRST ? u_at(d_sx, sy, AW, dpt[q], nm) : v_at(d_sx, sy, AW, dpt[q], nm);
            }
    }

    const bool bad = __syncthreads_or(mm.bad() | (inv != 0));
    if (bad) return true;

    // ---------------- stores: home cell + periodic images + wall cells ----------------------------------
    auto put = [&](int field, int i, int r, double val) {
        double *q = p.base + (size_t)field * plane + (size_t)(r - 1 + p.oy) * p.pitch + (size_t)(i - 1 + OX);
        const int ix = p.px ? (i <= W ? p.Nx : (i > p.Nx - W ? -p.Nx : 0)) : 0;
        const int iy = p.py ? (r <= W ? p.Ny : (r > p.Ny - W ? -p.Ny : 0)) : 0;
        q[0] = val;
        if (ix) q[ix] = val;
        if (iy) {
            q += (ptrdiff_t)iy * p.pitch;
            q[0] = val;
            if (ix) q[ix] = val;
        }
    };
    auto put_vel = [&](int field, int i, int r, double val, bool is_u) {
        if (!(i >= p.vx0 && i <= p.vx1 && r >= p.vy0 && r <= p.vy1)) return;
        put(field, i, r, val);
        double *q = p.base + (size_t)field * plane + (size_t)(r - 1 + p.oy) * p.pitch + (size_t)(i - 1 + OX);
        // walls: one tangential halo cell (value / no-flux BC), as fill_halo_regions! does -- and, on a mixed
        // topology, its periodic image along the other axis (the reference fills the periodic axis last,
        // over the full parent extent, so corners hold images of the wall cells)
        if (is_u && ((p.wall_s && r == 1) || (p.wall_n && r == p.Ny))) {
            const int ix = p.px ? (i <= W ? p.Nx : (i > p.Nx - W ? -p.Nx : 0)) : 0;
            const int ow = (p.oy + (r == 1 ? 0 : p.Ny)) * p.pitch + (i - 1 + OX);
            const double Dw = r == 1 ? mt.dyff(ow, 1) : mt.dyff(ow, p.Ny + 1);  // Delta y at (Face, Face) on the wall
            const double wv = r == 1 ? (p.u_sn_bc == CSI_BC_VALUE ? val + ((val - p.u_sn_val) / (Dw / 2)) * (-Dw) : val)
                                     : (p.u_sn_bc == CSI_BC_VALUE ? val + ((p.u_sn_val - val) / (Dw / 2)) * Dw : val);
            double *w = r == 1 ? q - p.pitch : q + p.pitch;
            w[0] = wv;
            if (ix) w[ix] = wv;
        }
        if (!is_u && ((p.wall_w && i == 1) || (p.wall_e && i == p.Nx))) {
            const int iy = p.py ? (r <= W ? p.Ny : (r > p.Ny - W ? -p.Ny : 0)) : 0;
            const double Dw = mt.dxff((r - 1 + p.oy) * p.pitch + (i == 1 ? 0 : p.Nx) + OX, r);  // Delta x at (Face, Face) on the wall
            const double wv = i == 1 ? (p.v_we_bc == CSI_BC_VALUE ? val + ((val - p.v_we_val) / (Dw / 2)) * (-Dw) : val)
                                     : (p.v_we_bc == CSI_BC_VALUE ? val + ((p.v_we_val - val) / (Dw / 2)) * Dw : val);
            double *w = i == 1 ? q - 1 : q + 1;
            w[0] = wv;
            if (iy) w[(ptrdiff_t)iy * p.pitch] = wv;
        }
    };
    // interior tiles (the vast majority): every output cell is inside all store windows and has no periodic
    // image or wall neighbour -> plain stores at one precomputed offset
    {
        if (INTERIOR) {
            // one 32-bit in-plane offset per node and the plane pointers of this launch (kernel parameters): one address
            // instruction per store
            uint32_t so = (uint32_t)(o00 + 2 * wrp * p.pitch + lane);  // this thread's node (lane, 2 wrp)
#pragma unroll
            for (int q = 0; q < 2; q++, so += (uint32_t)p.pitch) {
                const int sx = lane, sy = 2 * wrp + q;
                if (PH != 2 && sx >= 1 && sx <= OUTX && sy >= 1 && sy <= OUTY) {
                    p.o_s11[so] = S(A_S11, sx, sy);
                    p.o_s22[so] = S(A_S22, sx, sy);
                    p.o_s12[so] = S(A_S12, sx, sy);
                    if (AUX || PH == 1) (p.base + so)[(size_t)F_ALPHA * plane] = S(A_AL, sx, sy);   // (phase D alone reads alpha back)
                    if (AUX) {
                        double *g = p.base + so;
                        g[(size_t)F_ZC * plane] = aux_zc[q];
                        g[(size_t)F_ZF * plane] = aux_zf[q];
                        g[(size_t)F_DELTA * plane] = aux_Dc[q];
                    }
                }
                if (d_on[q]) {  // node (lane + 1, 2 wrp + 1 + q): the offset the pointwise inputs of phase D were read at
                    const uint32_t dg = d_g[q] + (uint32_t)(o00 - o00k);   // (o00k == o00 outside the timing experiment)
                    if (PH != 1) p.o_d[dg] = w2[q];
                    if (PH != 2) p.o_c[dg] = S(AW, d_sx, d_sy0 + q);
                }
            }
            return false;
        }
    }
    if (!INTERIOR)
#pragma unroll
    for (int q = 0; q < 2; q++) {
        // stresses (and aux) of the node this thread updated in phase B
        const int sx = lane, sy = 2 * wrp + q;
        const int i = tc.I0 - 1 + sx, r = tc.J0 - 1 + sy;
        if (PH != 2 && sx >= 1 && sx <= OUTX && sy >= 1 && sy <= OUTY && i >= p.sx0 && i <= p.sx1 && r >= p.sy0 && r <= p.sy1) {
            put(tc.fout + 2, i, r, S(A_S11, sx, sy));
            put(tc.fout + 3, i, r, S(A_S22, sx, sy));
            put(tc.fout + 4, i, r, S(A_S12, sx, sy));
            if (AUX || PH == 1) put(F_ALPHA, i, r, S(A_AL, sx, sy));
            if (AUX) {
                put(F_ZC, i, r, aux_zc[q]);
                put(F_ZF, i, r, aux_zf[q]);
                put(F_DELTA, i, r, aux_Dc[q]);
            }
        }
        // velocities of the output cell this thread updated in phase D
        if (d_on[q]) {
            const int dsy = d_sy0 + q;
            const int ui = tc.I0 - 1 + d_sx, ur = tc.J0 - 1 + dsy;
            if (VFIRST) {
                if (PH != 1) put_vel(tc.fout + 0, ui, ur, w2[q], true);
                if (PH != 2) put_vel(tc.fout + 1, ui, ur, S(AW, d_sx, dsy), false);
            } else {
                if (PH != 1) put_vel(tc.fout + 1, ui, ur, w2[q], false);
                if (PH != 2) put_vel(tc.fout + 0, ui, ur, S(AW, d_sx, dsy), true);
            }
        }
    }
    return false;
}

// VFIRST: odd substep (v then u, se.jl:183-187) or even (u then v, :178-182).  AUX: also write
// alpha, zeta_c, zeta_f, Delta (last substep of a stage).  GEN: keep the run-time configuration switches.
template <bool VFIRST, bool AUX, bool GEN, int MET, int PH = 0>
__global__ void __launch_bounds__(NT, MET == 2 ? CSI_FUSED_MINB_MET2 : (MET == 1 && GEN) ? CSI_FUSED_MINB_MET1 : GEN ? CSI_FUSED_MINB_GEN : CSI_FUSED_MINB) k_evp_substep_fused(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params p)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + NARR * ASTRIDE);
    const int y0 = p.sy0 < p.vy0 ? p.sy0 : p.vy0;
    TileCtx tc;
    tc.I0 = p.a0 + blockIdx.x * OUTX;
    tc.J0 = y0 + (blockIdx.y + p.ty0) * OUTY;
    tc.fin = p.in_set ? F_U1 : F_U0;
    tc.fout = p.out_set ? F_U1 : F_U0;
    const int by = blockIdx.y + p.ty0;
    const bool interior = (int)blockIdx.x >= p.it_x0 && (int)blockIdx.x <= p.it_x1 && by >= p.it_y0 && by <= p.it_y1;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // an input outside the validated range (flag raised by k_pack / k_prep) sends every tile of the stage to the IEEE pass.
    // The flag is requested here and consumed where the FAST pass decides whether it may store (no stall on its latency;
    // a pass over unvalidated inputs computes garbage and stores nothing)
    const int inv = *p.invalid;
    const bool redo = interior ? tile_pass<VFIRST, AUX, GEN, MET, PH, MathFast, true>(sm, bar, 0, &tmap, p, tc, inv)
                               : tile_pass<VFIRST, AUX, GEN, MET, PH, MathFast, false>(sm, bar, 0, &tmap, p, tc, inv);
    // an operand left the windows of the shortcut arithmetic (zero ice mass, NaN, a quotient out of range ...):
    // reload the tile and redo it with plain IEEE operators and the reference's expression tree
    if (redo) {
        __syncthreads();
        if (threadIdx.x == 0) atomicAdd(p.invalid + 1, 1);  // diagnostics: tiles that took the IEEE pass
#ifdef CSI_DEBUG_REDO
        if (threadIdx.x == 0) printf("IEEE re-pass: tile (%d, %d) of (%d, %d), I0 = %d, J0 = %d, phases %d, vfirst %d\n", (int)blockIdx.x, by, (int)gridDim.x, (int)gridDim.y + p.ty0, tc.I0, tc.J0, PH, (int)VFIRST);
#endif
        tile_pass<VFIRST, AUX, GEN, MET, PH, MathSlow, false>(sm, bar, 1, &tmap, p, tc, 0);
    }
}
// ---- small grids: a block of substeps in ONE cooperative launch ------------------------------------
// When every tile of the grid is resident at once (tiles <= SMs x CTAs per SM), a substep costs one tile pass plus the gap
// between two launches, and the gap is the larger part (BASELINE config 1 as shipped: 50 tiles, 11 us per substep).  This
// kernel keeps its CTAs over `nsub` substeps: tile pass, grid-wide barrier, tile pass ... -- the same tile_pass
// instantiations, the same stores, with a barrier where the stream order between two launches was.  Two parameter blocks:
// the planes read / written and the order of the velocity phases alternate with the parity of the substep.
__device__ __forceinline__ void grid_barrier(unsigned int *ctr, unsigned int target)
{
    // this CTA's global stores (generic proxy) must be visible to the other CTAs' TMA loads (async proxy) of the next substep
    asm volatile("fence.proxy.async;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
        } while (seen < target);
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
}

template <bool GEN>
__global__ void __launch_bounds__(NT, GEN ? CSI_FUSED_MINB_GEN : CSI_FUSED_MINB)
    k_evp_substeps_persistent(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ Params p_even, const __grid_constant__ Params p_odd, int first_sub, int nsub,
                              unsigned int *sync)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *sm = reinterpret_cast<double *>(smem_raw);
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm + NARR * ASTRIDE);
    const int y0 = p_even.sy0 < p_even.vy0 ? p_even.sy0 : p_even.vy0;
    TileCtx tc;
    tc.I0 = p_even.a0 + blockIdx.x * OUTX;
    tc.J0 = y0 + blockIdx.y * OUTY;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int inv = *p_even.invalid;
    const unsigned int nblocks = gridDim.x * gridDim.y;
    uint32_t parity = 0;   // both mbarriers complete one phase per tile pass
    for (int k = 0; k < nsub; k++) {
        bool redo;
        if ((first_sub + k) & 1) {   // odd substep: v first (se.jl:183-187)
            tc.fin = p_odd.in_set ? F_U1 : F_U0;
            tc.fout = p_odd.out_set ? F_U1 : F_U0;
            redo = tile_pass<true, false, GEN, 0, 0, MathFast, false>(sm, bar, parity, &tmap, p_odd, tc, inv);
            parity ^= 1u;
            if (redo) {
                __syncthreads();
                if (threadIdx.x == 0) atomicAdd(p_odd.invalid + 1, 1);
                tile_pass<true, false, GEN, 0, 0, MathSlow, false>(sm, bar, parity, &tmap, p_odd, tc, 0);
                parity ^= 1u;
            }
        } else {
            tc.fin = p_even.in_set ? F_U1 : F_U0;
            tc.fout = p_even.out_set ? F_U1 : F_U0;
            redo = tile_pass<false, false, GEN, 0, 0, MathFast, false>(sm, bar, parity, &tmap, p_even, tc, inv);
            parity ^= 1u;
            if (redo) {
                __syncthreads();
                if (threadIdx.x == 0) atomicAdd(p_even.invalid + 1, 1);
                tile_pass<false, false, GEN, 0, 0, MathSlow, false>(sm, bar, parity, &tmap, p_even, tc, 0);
                parity ^= 1u;
            }
        }
        grid_barrier(sync, nblocks * (unsigned int)(k + 1));
    }
}
#undef S
#undef SB
#undef FLG

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
        // divisors are positive in every kernel use (metrics, masses, alpha, gamma, Delta, 1 + dtau*tau)
        double x = rnd_double(s, span, false), y = rnd_double(s, span, true), z = rnd_double(s, span, true);
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
// every field of a stage in ONE launch (blockIdx.z selects the item): on small grids the per-field launches cost more than
// the copies
struct PackList {
    PackItem it[16];
    int n;
};
struct UnpackItem {
    DArr a;
    int field, i0, i1, j0, j1;
};
struct UnpackList {
    UnpackItem it[9];
    int n;
};
__global__ void k_pack(const __grid_constant__ PackList L, const __grid_constant__ Params p, int w)
{
    const PackItem &it = L.it[blockIdx.z];
    // window: i in [1-w, Nx+w(+1)] (w = Hx on a partitioned x axis), j in [1-w', Ny+w'(+1)], clipped to the parent
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
    // input validation for the FAST pass: zero, or magnitude in [2^-300, 2^300) (NaN, Inf, subnormals and extremes fail)
    const uint32_t hi = (uint32_t)__double2hiint(val) & 0x7fffffffu, e = hi >> 20;
    const bool zero = (hi | (uint32_t)__double2loint(val)) == 0u;
    if (!zero && (e < 1023u - 300u || e >= 1023u + 300u)) atomicOr(p.invalid, 1);
    // thickness and concentration must not carry a sign bit (not even -0): the open-water shortcut of phase B reads a mass
    // of exactly +0 as "the reference divides by +0 here"
    if ((it.field == F_H || it.field == F_A) && __double2hiint(val) < 0) atomicOr(p.invalid, 1);
}
// Stage constants of the substep loop, written once per time_step_momentum! (h, aice, tau_top do not change inside it):
//   F_M    m = h rho aice at every node (ClimaSeaIce.jl:42; the reference recomputes it at every use)
//   F_T1X  the top-stress term of the u tendency, tau_x / m_i * aice_i (mt:33, ext:176-181) in the FAST pass's scaled form
//          (tau_x / 2 m_i') * 2 aice_i with m_i' the harmless mass of marginal nodes -- the very operations the node function used to
//          repeat every substep, so the bits are the same; F_T1Y likewise for v.
// The IEEE pass keeps reading the raw stress.
__global__ void k_prep(Params p, int top_const)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= p.pitch) return;
    const size_t plane = (size_t)p.pitch * p.rows, o = (size_t)r * p.pitch + c;
    const double *H = p.base + (size_t)F_H * plane, *A = p.base + (size_t)F_A * plane;
    const double a11 = A[o];
    const double m11 = H[o] * p.rho_i * a11;
    p.base[(size_t)F_M * plane + o] = m11;
#if CSI_PRE_RM2 || CSI_PRE_RMC
    // a divisor the FAST pass may use with a stored reciprocal: inside the range window of MathFast::chkd, low word not all ones
    auto safe = [](double d) {
        return (uint32_t)__double2hiint(d) - MathFast::DLO <= MathFast::DSPAN && (uint32_t)__double2loint(d) != 0xffffffffu;
    };
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
#endif
    const bool west = c >= 1, south = r >= 1;
    const double m01 = west ? H[o - 1] * p.rho_i * A[o - 1] : 0.0, m10 = south ? H[o - p.pitch] * p.rho_i * A[o - p.pitch] : 0.0;
    double t1x = 0.0, t1y = 0.0;
#if CSI_PRE_RM2
    double rmu = -1.0, rmv = -1.0;
#endif
    if (west) {
        const double m2 = m11 + m01, a2 = a11 + A[o - 1];
        const bool active_ice = (m2 >= p.min_mass2) & (a2 >= p.min_conc2);
        const double tt = !p.use_t1 ? 0.0 : (top_const ? p.ttx : p.base[(size_t)F_TX * plane + o]);
        t1x = (tt / (active_ice ? m2 : 1.0)) * a2;
#if CSI_PRE_RM2
        rmu = active_ice ? (safe(m2) ? 1.0 / m2 : inf) : -1.0;
#endif
    }
    if (south) {
        const double m2 = m11 + m10, a2 = a11 + A[o - p.pitch];
        const bool active_ice = (m2 >= p.min_mass2) & (a2 >= p.min_conc2);
        const double tt = !p.use_t1 ? 0.0 : (top_const ? p.tty : p.base[(size_t)F_TY * plane + o]);
        t1y = (tt / (active_ice ? m2 : 1.0)) * a2;
#if CSI_PRE_RM2
        rmv = active_ice ? (safe(m2) ? 1.0 / m2 : inf) : -1.0;
#endif
    }
    if (p.use_t1) {
        p.base[(size_t)F_T1X * plane + o] = t1x;
        p.base[(size_t)F_T1Y * plane + o] = t1y;
    }
    // ---- less common configurations (GEN instantiation of the kernel) ----
    if (p.bot_expl) {  // prescribed bottom stress: + tau_bot / m_i * aice_i (mt:33,66), same form as the top term
        double tbx = 0.0, tby = 0.0;
        if (west) {
            const double m2 = m11 + m01, a2 = a11 + A[o - 1];
            const bool active_ice = (m2 >= p.min_mass2) & (a2 >= p.min_conc2);
            tbx = ((p.bot_arr ? p.base[(size_t)F_UE * plane + o] : p.tb_x) / (active_ice ? m2 : 1.0)) * a2;
        }
        if (south) {
            const double m2 = m11 + m10, a2 = a11 + A[o - p.pitch];
            const bool active_ice = (m2 >= p.min_mass2) & (a2 >= p.min_conc2);
            tby = ((p.bot_arr ? p.base[(size_t)F_VE * plane + o] : p.tb_y) / (active_ice ? m2 : 1.0)) * a2;
        }
        p.base[(size_t)F_TB1X * plane + o] = tbx;
        p.base[(size_t)F_TB1Y * plane + o] = tby;
    }
    const bool north = r + 1 < p.rows, east = c + 1 < p.pitch;
    if (p.top_sis && p.top_arr) {  // 4-point sums of the atmosphere velocity: 4 v_a-bar at u nodes, 4 u_a-bar at v nodes (ext:176-202)
        const double *UA = p.base + (size_t)F_TX * plane, *VA = p.base + (size_t)F_TY * plane;
        p.base[(size_t)F_SVA * plane + o] = (west && north) ? (VA[o - 1] + VA[o]) + (VA[o + p.pitch - 1] + VA[o + p.pitch]) : 0.0;
        p.base[(size_t)F_SUA * plane + o] = (south && east) ? (UA[o - p.pitch] + UA[o - p.pitch + 1]) + (UA[o] + UA[o + 1]) : 0.0;
    }
    if (p.fd_on) {
        // The value of a node that is not dynamically active: marginal ice (face mass and concentration above eps) takes the
        // free-drift velocity, anything else 0 (se:224-228, 259-263).  Free drift depends on the external stresses and velocities
        // only (stress_balance_free_drift.jl:61-129), so it is a constant of the stage; reference expression trees throughout.
        const double eps = 2.220446049250313e-16;
        const double *U0 = p.base + (size_t)F_U0 * plane, *V0 = p.base + (size_t)F_V0 * plane;
        const bool d_bot = p.sis != 0;                      // the SemiImplicitStress side; o = the other side
        const int o_kind = d_bot ? p.top_kind : p.bot_kind;
        const bool o_arr = d_bot ? p.top_arr != 0 : p.bot_arr != 0, d_arr = d_bot ? p.use_ue != 0 : p.top_arr != 0;
        const double *OX_ = p.base + (size_t)(d_bot ? F_TX : F_UE) * plane, *OY_ = p.base + (size_t)(d_bot ? F_TY : F_VE) * plane;
        const double *DX_ = p.base + (size_t)(d_bot ? F_UE : F_TX) * plane, *DY_ = p.base + (size_t)(d_bot ? F_VE : F_TY) * plane;
        const double ocx = d_bot ? p.ta_x : p.tb_x, ocy = d_bot ? p.ta_y : p.tb_y, dcx = d_bot ? p.ue_c : p.ta_x, dcy = d_bot ? p.ve_c : p.ta_y;
        const double Cdrag = d_bot ? p.rhoCd : p.top_rhoCd;
        // x/y_momentum_stress of the side that is not a SemiImplicitStress: explicit stress - zero * velocity (ext:34-38).  The
        // velocity is the stage's first; the product is +-0 and only decides the sign of a stress that is exactly -0 where the
        // drag side's velocity is exactly 0 too -- the one place where the reference's per-substep evaluation could differ
        auto xms = [&](size_t q) { return (o_kind == CSI_STRESS_NONE ? 0.0 : (o_arr ? OX_[q] : ocx)) - 0.0 * U0[q]; };
        auto yms = [&](size_t q) { return (o_kind == CSI_STRESS_NONE ? 0.0 : (o_arr ? OY_[q] : ocy)) - 0.0 * V0[q]; };
        double ufd = 0.0, vfd = 0.0;
        if (west) {
            const double mi = (m11 + m01) / 2, ai = (a11 + A[o - 1]) / 2;
            double uF = 0.0;
            if (p.fd_kind == CSI_FD_FIELDS) uF = p.base[(size_t)F_FDU * plane + o];
            else if (north) {
                const double tx = xms(o);
                const double ty = ((yms(o - 1) + yms(o)) / 2 + (yms(o + p.pitch - 1) + yms(o + p.pitch)) / 2) / 2;
                const double t = sqrt(tx * tx + ty * ty);
                const double Ud = d_arr ? DX_[o] : dcx;
                uF = Ud - (t == 0 ? t : tx / sqrt(Cdrag * t));
            }
            ufd = ((mi > eps) & (ai > eps)) ? uF : 0.0;
        }
        if (south) {
            const double mi = (m11 + m10) / 2, ai = (a11 + A[o - p.pitch]) / 2;
            double vF = 0.0;
            if (p.fd_kind == CSI_FD_FIELDS) vF = p.base[(size_t)F_FDV * plane + o];
            else if (east) {
                const double tx = ((xms(o - p.pitch) + xms(o - p.pitch + 1)) / 2 + (xms(o) + xms(o + 1)) / 2) / 2;
                const double ty = yms(o);
                const double t = sqrt(tx * tx + ty * ty);
                const double Ud = d_arr ? DY_[o] : dcy;
                vF = Ud - (t == 0 ? t : ty / sqrt(Cdrag * t));
            }
            vfd = ((mi > eps) & (ai > eps)) ? vF : 0.0;
        }
        p.base[(size_t)F_UFD * plane + o] = ufd;
        p.base[(size_t)F_VFD * plane + o] = vfd;
        // these values become velocities: the FAST pass's induction needs them zero or in [2^-300, 2^300) like every input
        for (double val : {ufd, vfd}) {
            const uint32_t hi = (uint32_t)__double2hiint(val) & 0x7fffffffu, e = hi >> 20;
            if ((hi | (uint32_t)__double2loint(val)) != 0u && (e < 1023u - 300u || e >= 1023u + 300u)) atomicOr(p.invalid, 1);
        }
    }
#if CSI_PRE_RM2
    p.base[(size_t)F_RM2U * plane + o] = rmu;
    p.base[(size_t)F_RM2V * plane + o] = rmv;
#endif
#if CSI_PRE_RMC || CSI_PRE_PF4
    {
        const double *Pp = p.base + (size_t)F_P * plane;
        const double Pc = Pp[o];
        double mf4 = 0.0, Pf4 = 0.0;
        if (west && south) {
            const double m00 = H[o - p.pitch - 1] * p.rho_i * A[o - p.pitch - 1];
            mf4 = (m00 + m10) + (m01 + m11);
            Pf4 = (Pp[o - p.pitch - 1] + Pp[o - p.pitch]) + (Pp[o - 1] + Pc);
        }
        // open water (mass exactly 0, P >= 0): the kernel divides by the harmless 1 resp. 4 (see phase B)
        p.base[(size_t)F_RMC * plane + o] = (m11 == 0.0 && Pc >= 0.0) ? 1.0 : (safe(m11) ? 1.0 / m11 : inf);
        p.base[(size_t)F_RMF * plane + o] = (mf4 == 0.0 && Pf4 >= 0.0) ? 0.25 : (safe(mf4) ? 1.0 / mf4 : inf);
        p.base[(size_t)F_PF4 * plane + o] = Pf4;
    }
#endif
#if CSI_PRE_SVE
    if (p.use_ue) {
        const double *UE = p.base + (size_t)F_UE * plane, *VE = p.base + (size_t)F_VE * plane;
        // 4 v_e-bar at the u node (i, j): (i-1, j) + (i, j) + (i-1, j+1) + (i, j+1); 4 u_e-bar at the v node: (i, j-1) + (i+1, j-1) + (i, j) + (i+1, j)
        p.base[(size_t)F_SVE * plane + o] = (west && north) ? (VE[o - 1] + VE[o]) + (VE[o + p.pitch - 1] + VE[o + p.pitch]) : 0.0;
        p.base[(size_t)F_SUE * plane + o] = (south && east) ? (UE[o - p.pitch] + UE[o - p.pitch + 1]) + (UE[o] + UE[o + 1]) : 0.0;
    }
#endif
}
__global__ void k_unpack(const __grid_constant__ UnpackList L, const __grid_constant__ Params p)
{
    const UnpackItem &it = L.it[blockIdx.z];
    const int i = it.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    const int j = it.j0 + blockIdx.y;
    if (i > it.i1 || j > it.j1) return;
    const size_t plane = (size_t)p.pitch * p.rows;
    at(it.a, i, j) = p.base[(size_t)it.field * plane + (size_t)(j - 1 + p.oy) * p.pitch + (size_t)(i - 1 + OX)];
}

}  // namespace fz

// ---- host side ----------------------------------------------------------------------------------
struct FusedPlan {
    uint8_t *flags = nullptr;
    double *met = nullptr;  // per-row metric table (lat-lon grids)
    // a folded north boundary: the copy lists of u and v in the internal layout (in-plane offsets), their sign, and the first
    // tile row whose second velocity phase can read a first-velocity value the fold replaces (fused_steps)
    int32_t *fold_t[2] = {nullptr, nullptr}, *fold_s[2] = {nullptr, nullptr};
    int fold_n[2] = {0, 0};
    double fold_sign = 1.0;
    int fold_jmin = 0;      // southernmost reference row a list writes
    int fold_jsrc_min = 0;  // southernmost reference row a list reads
    // the split band of a substep runs on a stream of its own, beside the bulk of the mesh (fused_steps)
    cudaStream_t side = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool partitioned = false;  // a side of the block is connected to another rank (halo exchanges between blocks of substeps)
    double *met2 = nullptr; // two-dimensional metric planes (orthogonal curvilinear grids), MC2_N x (rows + 2 MET2_PAD) x pitch
    long long met2_stride = 0;
    int *invalid = nullptr; // device flag: an input of the current stage is outside the validated range
    int persistent = -1;    // small grids: -1 not decided, 0 no, 1 the block of substeps runs as one cooperative launch (GEN = false), 2 (GEN = true)
    long long persistent_launches = 0;
    fz::Params P;
    dim3 grid;
    int cur_set = 0;
    double *base = nullptr;
    int pitch = 0, rows = 0, oy = 0;
    CUtensorMap tmap;
    int nf = 0;             // planes allocated: the common set, + those of the less common configurations, + the experiments'
    int Nx = 0, Ny = 0;
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

// The grid's metric arrays against the premises of the fused kernel: checked once, when the handle is created (the arrays
// never change afterwards, and a sweep over twelve two-dimensional arrays per momentum solve would cost more than the solve).
const char *fused_metrics_check(const DGrid &g)
{
    if (!g.met) return nullptr;
    const int cols = g.metW ? g.metW : 1;
    const size_t metn = (size_t)g.metL * cols;  // entries per metric array
    // every metric that appears as a divisor must qualify for the constant-division shortcut, on every node a tile can touch
    if (!g.met_host) return "host copy of the grid metrics missing";
    const int divisors[8] = {M_DXFC, M_DXCF, M_DYFC, M_DYCF, M_AZCC, M_AZFC, M_AZCF, M_AZFF};
    for (int k : divisors)
        for (size_t q = 0; q < metn; q++)
            if (!recip_is_safe(g.met_host[(size_t)k * metn + q])) return "a grid metric is not eligible for the constant-division shortcut";
    // premises of the scaled expression tree: grid spacings in metres (areas: the square) far inside the normal range, on the rows
    // whose results are kept: the interior and the wall ring; a slab's connected side uses its whole halo
    for (int k : divisors)
        for (int q = (g.conn_s ? 0 : g.Hy - 1); q < (g.conn_n ? g.metL : std::min(g.metL, g.Hy + g.Ny + 2)); q++)
            for (int c = 0; c < cols; c++) {
                const double v = g.met_host[(size_t)k * metn + (size_t)q * cols + c];
                if (!(k >= M_AZCC ? (v >= 1e-12 && v <= 1e24) : (v >= 1e-6 && v <= 1e12))) return "a grid metric is outside the range the fused kernel's exact scalings assume";
            }
    return nullptr;
}

int fused_supported(const DGrid &g, const DParams &p, const DFields &f, char *why, int nwhy)
{
    if (g.fold && !(g.met && g.metW)) { snprintf(why, nwhy, "a folded (tripolar) north boundary on a grid without two-dimensional metrics (general kernels only)"); return 0; }
    if (g.fold && (!g.fold_t_host[1] || !g.fold_t_host[2])) { snprintf(why, nwhy, "host copies of the fold lists missing"); return 0; }
    if (g.met && g.metW && (g.conn_w || g.conn_e)) { snprintf(why, nwhy, "two-dimensional metrics with a partition along x"); return 0; }
    if (g.met && g.met_fused_why) { snprintf(why, nwhy, "%s", g.met_fused_why); return 0; }
    if (p.cor == CSI_CORIOLIS_SPHERICAL && (!g.met || g.metW)) { snprintf(why, nwhy, "HydrostaticSphericalCoriolis needs a lat-lon grid"); return 0; }
    if (p.fd_kind == CSI_FD_FIELDS && (!f.fd_u.p || !f.fd_v.p)) { snprintf(why, nwhy, "free-drift arrays missing"); return 0; }
    if ((f.top_x.p == nullptr) != (f.top_y.p == nullptr)) { snprintf(why, nwhy, "top_x/top_y kinds differ"); return 0; }
    if (!g.met && (!recip_is_safe(g.dx) || !recip_is_safe(g.dy) || !recip_is_safe(g.az))) { snprintf(why, nwhy, "grid metric not eligible for the constant-division shortcut"); return 0; }
    if (g.Nx < 8 || g.Ny < 8) { snprintf(why, nwhy, "grid too small"); return 0; }
    // premises of the scaled expression tree (exact power-of-two scalings): thresholds and constants far inside the normal range
    // (and the quotients left unchecked in the kernel -- by cell areas, by alpha -- stay far from over/underflow)
    auto sane = [](double x) { return x == 0.0 || (fabs(x) >= 1e-30 && fabs(x) <= 1e30); };
    auto sane_len = [](double x) { return x >= 1e-6 && x <= 1e12; };  // grid spacings in metres (areas: the square)
    if (p.top_kind == CSI_STRESS_SEMI_IMPLICIT && !sane(p.top_rho * p.top_Cd)) { snprintf(why, nwhy, "top drag coefficient outside the range the exact scalings assume"); return 0; }
    bool ok = p.min_conc >= 1e-30 && p.min_mass >= 1e-30 && p.amin >= 1e-30 && p.amax <= 1e30 && p.amin <= p.amax && sane(p.f) && p.Dmin >= 1e-30 && p.Dmin <= 1e30 && sane(p.em2) &&
              sane(p.ca) && sane(p.rho_e * p.Cd) && sane(p.rho_i) && p.rho_i > 0;
    if (!g.met) ok = ok && sane_len(g.dx) && sane_len(g.dy);   // (metric arrays: fused_metrics_check, once per handle)
    if (!ok) { snprintf(why, nwhy, "a threshold or constant is outside the range the fused kernel's exact scalings assume"); return 0; }
    if ((f.ue.p == nullptr) != (f.ve.p == nullptr)) { snprintf(why, nwhy, "ue/ve kinds differ"); return 0; }
    (void)p;
    return 1;
}

FusedPlan *fused_create(const DGrid &g, const DParams &prm, char *err, int nerr)
{
    using namespace fz;
    FusedPlan *pl = new FusedPlan();
    const bool gen_planes = prm.fd_kind != CSI_FD_NONE || prm.top_kind == CSI_STRESS_SEMI_IMPLICIT || prm.bot_kind == CSI_STRESS_CONST || prm.bot_kind == CSI_STRESS_FIELD;
    pl->nf = (CSI_PRE_RM2 || CSI_PRE_RMC || CSI_PRE_PF4 || CSI_PRE_SVE) ? (int)NF : (gen_planes ? (int)NF_GEN : (int)NF_COMMON);
    pl->Nx = g.Nx;
    pl->Ny = g.Ny;
    pl->partitioned = g.conn_s || g.conn_n || g.conn_w || g.conn_e;
    pl->oy = (g.conn_s || g.conn_n) ? g.Hy : W + 1;
    const int wx = (g.conn_w || g.conn_e) ? g.Hx : W;  // halo columns kept in the internal layout
    if (wx > OX - 3) { snprintf(err, nerr, "fused solver: Hx = %d exceeds the internal x halo (%d)", g.Hx, OX - 3); delete pl; return nullptr; }
    pl->pitch = ((OX + g.Nx + 1 + wx + 1 + 15) / 16) * 16;
    pl->rows = g.Ny + 2 * pl->oy + 1;
    if ((double)pl->pitch * (double)pl->rows >= 2147483647.0) { snprintf(err, nerr, "fused solver: block too large for 32-bit in-plane offsets"); delete pl; return nullptr; }
    const size_t bytes = (size_t)pl->nf * pl->pitch * pl->rows * sizeof(double);
    cudaError_t e = cudaMalloc(&pl->base, bytes);
    if (e != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); delete pl; return nullptr; }
    cudaMemset(pl->base, 0, bytes);
    if (cudaMalloc(&pl->invalid, 4 * sizeof(int)) != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(flag)"); cudaFree(pl->base); delete pl; return nullptr; }
    cudaMemset(pl->invalid, 0, 4 * sizeof(int));   // [0] inputs failed validation, [1] tile passes redone, [2] grid barrier of the persistent kernel
    cudaDeviceSynchronize();  // the plan may be used next from a non-blocking stream
    if (g.mask_host) {
        // node flags from the centre mask, with the reference's inactive_cell / immersed_peripheral_node logic
        const int msx = g.Nx + 2 * g.Hx, msy = g.Ny + 2 * g.Hy;
        auto outside = [&](int i, int j) {
            return (g.topo_x == CSI_BOUNDED && ((i < 1 && !g.conn_w) || (i > g.Nx && !g.conn_e))) || (g.topo_y == CSI_BOUNDED && ((j < 1 && !g.conn_s) || (j > g.Ny && !g.conn_n)));
        };
        auto immersed = [&](int i, int j) {
            const int pi = std::min(std::max(i - 1 + g.Hx, 0), msx - 1), pj = std::min(std::max(j - 1 + g.Hy, 0), msy - 1);
            return g.mask_host[(size_t)pi + (size_t)pj * msx] != 0;
        };
        auto inactive = [&](int i, int j) { return outside(i, j) || immersed(i, j); };
        std::vector<uint8_t> fl((size_t)pl->pitch * pl->rows, 0);
        for (int r = 0; r < pl->rows; r++)
            for (int c = 0; c < pl->pitch; c++) {
                const int i = c + 1 - OX, j = r + 1 - pl->oy;
                uint8_t f = 0;
                if (inactive(i, j) && !outside(i, j)) f |= 1;
                const bool per = inactive(i - 1, j - 1) || inactive(i, j - 1) || inactive(i - 1, j) || inactive(i, j);
                const bool und = outside(i - 1, j - 1) || outside(i, j - 1) || outside(i - 1, j) || outside(i, j);
                if (per && !und) f |= 2;
                if (immersed(i - 1, j) || immersed(i, j)) f |= 4;
                if (immersed(i, j - 1) || immersed(i, j)) f |= 8;
                fl[(size_t)r * pl->pitch + c] = f;
            }
        if (cudaMalloc(&pl->flags, fl.size()) != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(flags)"); cudaFree(pl->base); delete pl; return nullptr; }
        cudaMemcpy(pl->flags, fl.data(), fl.size(), cudaMemcpyHostToDevice);
    }
    if (g.fold) {
        // the copy lists of u (Face, Center) and v (Center, Face), re-indexed from the caller's parents to the internal layout;
        // targets in halo columns the internal layout does not keep are dropped (nothing reads them), sources must exist
        for (int w = 0; w < 2; w++) {
            const int loc = w + 1, lx = loc & 1;
            const int sxp = g.Nx + 2 * g.Hx + ((lx && g.topo_x == CSI_BOUNDED && !g.conn_w && !g.conn_e) ? 1 : 0);
            std::vector<int32_t> tg, sr;
            int jmin = 1 << 30, jsrc = 1 << 30;
            for (int k = 0; k < g.fold_n[loc]; k++) {
                const int t = g.fold_t_host[loc][k], q = g.fold_s_host[loc][k];
                const int ti = t % sxp + 1 - g.Hx, tj = t / sxp + 1 - g.Hy, qi = q % sxp + 1 - g.Hx, qj = q / sxp + 1 - g.Hy;
                const int tc = ti - 1 + OX, tr = tj - 1 + pl->oy, qc = qi - 1 + OX, qr = qj - 1 + pl->oy;
                if (tc < 0 || tc >= pl->pitch || tr < 0 || tr >= pl->rows) continue;
                if (qc < 0 || qc >= pl->pitch || qr < 0 || qr >= pl->rows) { snprintf(err, nerr, "fused solver: a fold source lies outside the internal layout"); cudaFree(pl->base); delete pl; return nullptr; }
                tg.push_back(tr * pl->pitch + tc);
                sr.push_back(qr * pl->pitch + qc);
                jmin = std::min(jmin, tj);
                jsrc = std::min(jsrc, qj);
            }
            pl->fold_n[w] = (int)tg.size();
            if (w == 0 || jmin < pl->fold_jmin) pl->fold_jmin = jmin;
            if (w == 0 || jsrc < pl->fold_jsrc_min) pl->fold_jsrc_min = jsrc;
            if (tg.empty()) continue;
            if (cudaMalloc(&pl->fold_t[w], tg.size() * sizeof(int32_t)) != cudaSuccess || cudaMalloc(&pl->fold_s[w], sr.size() * sizeof(int32_t)) != cudaSuccess) {
                snprintf(err, nerr, "cudaMalloc(fold lists)"); cudaFree(pl->base); delete pl; return nullptr;
            }
            cudaMemcpy(pl->fold_t[w], tg.data(), tg.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
            cudaMemcpy(pl->fold_s[w], sr.data(), sr.size() * sizeof(int32_t), cudaMemcpyHostToDevice);
        }
        pl->fold_sign = g.fold_sv;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // (hi is the greatest priority: the band's small launches go first when slots free up)
        if (cudaStreamCreateWithPriority(&pl->side, cudaStreamNonBlocking, hi) != cudaSuccess || cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            pl->side = nullptr;   // (the band then runs on the caller's stream, after the bulk)
        }
        // phase D alone reads alpha back from its plane, also at nodes no launch stores: a value of alpha's range there
        std::vector<double> ones((size_t)pl->pitch * pl->rows, prm.amin);
        cudaMemcpy(pl->base + (size_t)F_ALPHA * pl->pitch * pl->rows, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    if (g.met_host && g.metW) {
        // two-dimensional metrics: MC2_N planes in the internal layout, node (i, j) at its in-plane offset; beyond the host arrays
        // (padding rows, slack columns) the nearest host value, so that whatever an edge tile names is a valid metric
        const size_t prow = (size_t)pl->rows + 2 * MET2_PAD, pn = prow * pl->pitch;
        const size_t hn = (size_t)g.metL * g.metW;
        std::vector<double> tb((size_t)MC2_N * pn);
        const int rcol[8] = {M_DXFC, M_DXCF, M_DYFC, M_DYCF, M_AZCC, M_AZFC, M_AZCF, M_AZFF};  // order of MC_RDXFC .. MC_RAZFF
        for (size_t rr = 0; rr < prow; rr++)
            for (int c = 0; c < pl->pitch; c++) {
                const int i = c + 1 - OX, j = (int)rr - MET2_PAD + 1 - pl->oy;
                const size_t pi = (size_t)std::min(std::max(i - 1 + g.Hx, 0), g.metW - 1), pj = (size_t)std::min(std::max(j - 1 + g.Hy, 0), g.metL - 1);
                const size_t src = pj * g.metW + pi, dst = rr * pl->pitch + c;
                for (int k = 0; k < 12; k++) tb[(size_t)k * pn + dst] = g.met_host[(size_t)k * hn + src];
                for (int k = 0; k < 8; k++) tb[(size_t)(12 + k) * pn + dst] = 1.0 / g.met_host[(size_t)rcol[k] * hn + src];
            }
        if (cudaMalloc(&pl->met2, tb.size() * sizeof(double)) != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(metric planes)"); cudaFree(pl->base); delete pl; return nullptr; }
        cudaMemcpy(pl->met2, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice);
        pl->met2_stride = (long long)pn;
    } else if (g.met_host) {
        // per-row metric table in internal row coordinates: row rho holds reference row j = rho + 1 - oy
        std::vector<double> tb((size_t)MC_N * (pl->rows + 2 * MET_PAD), 1.0);
        auto src = [&](int which, int rho) {
            const int q = rho - pl->oy + g.Hy;  // index of row j in the host arrays (j - 1 + Hy)
            return (q >= 0 && q < g.metL) ? g.met_host[(size_t)which * g.metL + q] : 1.0;
        };
        for (int rho = 0; rho < pl->rows; rho++) {
            auto put = [&](int col, double v) { tb[(size_t)(rho + MET_PAD) * MC_N + col] = v; };
            for (int k = 0; k < 12; k++) put(MC_DXCC + k, src(k, rho));
            put(MC_DXCC2, src(M_DXCC, rho) * src(M_DXCC, rho));
            put(MC_DYCC2, src(M_DYCC, rho) * src(M_DYCC, rho));
            put(MC_DXFF2, src(M_DXFF, rho) * src(M_DXFF, rho));
            put(MC_DYFF2, src(M_DYFF, rho) * src(M_DYFF, rho));
            put(MC_RDXFC, 1.0 / src(M_DXFC, rho));
            put(MC_RDXCF, 1.0 / src(M_DXCF, rho));
            put(MC_RDYFC, 1.0 / src(M_DYFC, rho));
            put(MC_RDYCF, 1.0 / src(M_DYCF, rho));
            put(MC_RAZCC, 1.0 / src(M_AZCC, rho));
            put(MC_RAZFC, 1.0 / src(M_AZFC, rho));
            put(MC_RAZCF, 1.0 / src(M_AZCF, rho));
            put(MC_RAZFF, 1.0 / src(M_AZFF, rho));
            const int q = rho - pl->oy + g.Hy;
            put(MC_FFF, (g.fff_host && q >= 0 && q < g.metL) ? g.fff_host[q] : 0.0);
        }
        if (cudaMalloc(&pl->met, tb.size() * sizeof(double)) != cudaSuccess) { snprintf(err, nerr, "cudaMalloc(metric table)"); cudaFree(pl->base); delete pl; return nullptr; }
        cudaMemcpy(pl->met, tb.data(), tb.size() * sizeof(double), cudaMemcpyHostToDevice);
    }
    EncodeTiledFn enc = get_encode();
    if (!enc) { snprintf(err, nerr, "cuTensorMapEncodeTiled unavailable"); cudaFree(pl->base); delete pl; return nullptr; }
    cuuint64_t dims[3] = {(cuuint64_t)pl->pitch, (cuuint64_t)pl->rows, (cuuint64_t)pl->nf};
    cuuint64_t strides[2] = {(cuuint64_t)pl->pitch * 8, (cuuint64_t)pl->pitch * pl->rows * 8};
    cuuint32_t box[3] = {(cuuint32_t)SXD, (cuuint32_t)SYD, 1};
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
    if (pl->flags) cudaFree(pl->flags);
    if (pl->met) cudaFree(pl->met);
    if (pl->met2) cudaFree(pl->met2);
    for (int w = 0; w < 2; w++) { if (pl->fold_t[w]) cudaFree(pl->fold_t[w]); if (pl->fold_s[w]) cudaFree(pl->fold_s[w]); }
    if (pl->side) cudaStreamDestroy(pl->side);
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    if (pl->ev_join) cudaEventDestroy(pl->ev_join);
    if (pl->invalid) cudaFree(pl->invalid);
    delete pl;
}

template <bool VFIRST, bool AUX, bool GEN, int MET, int PH = 0> static cudaError_t launch_one(const FusedPlan *pl, const fz::Params &P, dim3 grid, cudaStream_t s)
{
    using namespace fz;
    // the attribute is per device (a process may hold handles on several devices through csi_config.device)
    static bool attr[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(k_evp_substep_fused<VFIRST, AUX, GEN, MET, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev >= 0 && dev < 64) attr[dev] = true;
    }
    k_evp_substep_fused<VFIRST, AUX, GEN, MET, PH><<<grid, NT, SMEM_BYTES, s>>>(pl->tmap, P);
    return cudaGetLastError();
}
// the split substep of a mesh with a fold (two-dimensional metrics): phases A-C / phase D alone
template <int PH> static cudaError_t launch_split(const FusedPlan *pl, const fz::Params &P, dim3 grid, cudaStream_t s, bool vfirst, bool aux)
{
    if (PH == 2) return vfirst ? launch_one<true, false, true, 2, 2>(pl, P, grid, s) : launch_one<false, false, true, 2, 2>(pl, P, grid, s);
    if (vfirst) return aux ? launch_one<true, true, true, 2, PH>(pl, P, grid, s) : launch_one<true, false, true, 2, PH>(pl, P, grid, s);
    return aux ? launch_one<false, true, true, 2, PH>(pl, P, grid, s) : launch_one<false, false, true, 2, PH>(pl, P, grid, s);
}
// GEN x MET: the regular grid has a switch-free variant; lat-lon grids (MET) always keep the run-time switches
template <bool GEN, int MET> static cudaError_t launch_sub(const FusedPlan *pl, const fz::Params &P, dim3 grid, cudaStream_t s, bool vfirst, bool aux)
{
    if (vfirst) return aux ? launch_one<true, true, GEN, MET>(pl, P, grid, s) : launch_one<true, false, GEN, MET>(pl, P, grid, s);
    return aux ? launch_one<false, true, GEN, MET>(pl, P, grid, s) : launch_one<false, false, GEN, MET>(pl, P, grid, s);
}

int fused_begin(FusedPlan *pl, const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, char *err, int nerr)
{
    using namespace fz;
    Params &P = pl->P;

    if (!(dt >= 1e-30 && dt <= 1e30)) { snprintf(err, nerr, "fused solver: time step outside [1e-30, 1e30]"); return CSI_ERR_ARG; }
    memset(&P, 0, sizeof P);
    P.Nx = g.Nx; P.Ny = g.Ny; P.pitch = pl->pitch; P.rows = pl->rows; P.oy = pl->oy;
    P.px = g.topo_x == CSI_PERIODIC && !g.conn_w && !g.conn_e;
    P.py = g.topo_y == CSI_PERIODIC && !g.conn_s && !g.conn_n;
    P.bounded_x = g.topo_x == CSI_BOUNDED;
    P.bounded_y = g.topo_y == CSI_BOUNDED;
    P.wall_s = P.bounded_y && !g.conn_s;
    P.wall_n = P.bounded_y && !g.conn_n;
    P.wall_w = P.bounded_x && !g.conn_w;
    P.wall_e = P.bounded_x && !g.conn_e;
    // stresses: interior for periodic axes, one extra ring on Bounded axes (boundary nodes of sigma12 and
    // the first halo cell, which the reference also evolves, evp.jl:145); slabs: the widened range
    P.sx0 = P.wall_w ? 0 : 1; P.sx1 = P.wall_e ? g.Nx + 1 : g.Nx;
    P.sy0 = P.wall_s ? 0 : 1; P.sy1 = P.wall_n ? g.Ny + 1 : g.Ny;
    P.vx0 = 1; P.vx1 = g.Nx; P.vy0 = 1; P.vy1 = g.Ny;
    if (g.conn_s) { P.sy0 = -g.Hy + 2; P.vy0 = -g.Hy + 2; }
    if (g.conn_n) { P.sy1 = g.Ny + g.Hy - 1; P.vy1 = g.Ny + g.Hy - 1; }
    if (g.conn_w) { P.sx0 = -g.Hx + 2; P.vx0 = -g.Hx + 2; }
    if (g.conn_e) { P.sx1 = g.Nx + g.Hx - 1; P.vx1 = g.Nx + g.Hx - 1; }
    const int BIG = 1 << 29;
    P.cx0 = P.px ? -BIG : P.vx0; P.cx1 = P.px ? BIG : P.vx1;
    P.cy0 = P.py ? -BIG : P.vy0; P.cy1 = P.py ? BIG : P.vy1;
    P.use_top = p.top_kind == CSI_STRESS_FIELD;
    P.use_t1 = p.top_kind == CSI_STRESS_FIELD || p.top_kind == CSI_STRESS_CONST;
    P.use_ue = f.ue.p != nullptr && p.bot_kind == CSI_STRESS_SEMI_IMPLICIT;
    P.fd_on = p.fd_kind != CSI_FD_NONE; P.fd_kind = p.fd_kind;
    P.top_sis = p.top_kind == CSI_STRESS_SEMI_IMPLICIT; P.top_arr = f.top_x.p != nullptr;
    P.bot_expl = p.bot_kind == CSI_STRESS_CONST || p.bot_kind == CSI_STRESS_FIELD; P.bot_arr = f.ue.p != nullptr; P.bot_kind = p.bot_kind;
    P.top_kind = p.top_kind;
    P.top_rhoCd = p.top_rho * p.top_Cd;
    P.ta_x = p.ttx; P.ta_y = p.tty;   // constants of a top SemiImplicitStress without arrays
    P.tb_x = p.ue_c; P.tb_y = p.ve_c; // constants of a prescribed bottom stress
    P.u_sn_bc = p.u_sn_bc; P.v_we_bc = p.v_we_bc; P.u_sn_val = p.u_sn_val; P.v_we_val = p.v_we_val;
    P.dt = dt;
    P.dx = g.dx; P.dy = g.dy; P.az = g.az; P.dx2 = g.dx * g.dx; P.dy2 = g.dy * g.dy;
    P.rdx = 1.0 / g.dx; P.rdy = 1.0 / g.dy; P.raz = 1.0 / g.az;
    P.em2 = p.em2; P.Dmin = p.Dmin; P.amin = p.amin; P.amax = p.amax; P.amax2 = p.amax * p.amax; P.ca = p.ca;
    P.rho_i = p.rho_i; P.rhoCd = p.rho_e * p.Cd; P.f = p.f; P.min_mass = p.min_mass; P.min_conc = p.min_conc;
    P.ttx = p.top_kind == CSI_STRESS_CONST ? p.ttx : 0.0;
    P.tty = p.top_kind == CSI_STRESS_CONST ? p.tty : 0.0;
    P.ue_c = p.ue_c; P.ve_c = p.ve_c;
    P.imm_u = p.imm_u; P.imm_v = p.imm_v;
    P.pform = p.pform; P.cor = p.cor; P.sis = p.bot_kind == CSI_STRESS_SEMI_IMPLICIT;
    P.dt2 = 2 * dt; P.dt4 = 4 * dt; P.f4 = p.f * 0.25; P.Dmin2 = 2 * p.Dmin; P.Dmin8 = 8 * p.Dmin;
    P.min_mass2 = 2 * p.min_mass; P.min_conc2 = 2 * p.min_conc; P.dx2d = 2 * P.dx2; P.dy2d = 2 * P.dy2;
    P.gnan = jl_clamp(sqrt(P.amax2), p.amin, p.amax);
    P.sq = !pl->met && !pl->met2 && g.dx == g.dy;
    P.base = pl->base;
    P.flags = pl->flags;
    P.met = pl->met ? pl->met + (size_t)(MET_PAD - 1 + pl->oy) * MC_N : nullptr;  // record of row j at met[j * MC_N]
    P.met2 = pl->met2 ? pl->met2 + (size_t)MET2_PAD * pl->pitch : nullptr;  // entry of the node at in-plane offset o at met2[o]
    P.met2_stride = pl->met2_stride;
    P.invalid = pl->invalid;
    cudaMemsetAsync(pl->invalid, 0, 2 * sizeof(int), c.stream);  // re-validated by the pack kernels below; tile counter reset
    if (pl->met || pl->met2) { P.dx = P.dy = P.az = P.dx2 = P.dy2 = P.rdx = P.rdy = P.raz = 1.0; }

    // the TMA box of tile column k starts at internal column a0 - 3 + OX + OUTX k: keep it even (16-byte aligned)
    P.a0 = P.sx0 < P.vx0 ? P.sx0 : P.vx0;
    if ((P.a0 - 3 + OX) & 1) P.a0 -= 1;
    const int ncols = (P.sx1 > P.vx1 ? P.sx1 : P.vx1) - P.a0 + 1;
    const int nrows = (P.sy1 > P.vy1 ? P.sy1 : P.vy1) - (P.sy0 < P.vy0 ? P.sy0 : P.vy0) + 1;
    dim3 grid((ncols + OUTX - 1) / OUTX, (nrows + OUTY - 1) / OUTY);

    // pack: caller parents -> internal layout (window includes W halo cells; evolving fields into both copies)
    const int w = (g.conn_w || g.conn_e) ? g.Hx : W;
    PackList PL;
    PL.n = 0;
    auto pack = [&](const DArr &a, int field, int dup, int lx, int ly) { PL.it[PL.n++] = PackItem{a, field, dup, lx, ly}; };
    pack(f.u, F_U0, 1, 1, 0); pack(f.v, F_V0, 1, 0, 1);
    pack(f.s11, F_S11_0, 1, 0, 0); pack(f.s22, F_S22_0, 1, 0, 0); pack(f.s12, F_S12_0, 1, 1, 1);
    pack(f.h, F_H, 0, 0, 0); pack(f.a, F_A, 0, 0, 0); pack(f.P, F_P, 0, 0, 0);
    pack(f.un, F_UN, 0, 1, 0); pack(f.vn, F_VN, 0, 0, 1);
    if (P.use_top || (P.top_sis && P.top_arr)) { pack(f.top_x, F_TX, 0, 1, 0); pack(f.top_y, F_TY, 0, 0, 1); }
    if (P.use_ue || (P.bot_expl && P.bot_arr)) { pack(f.ue, F_UE, 0, 1, 0); pack(f.ve, F_VE, 0, 0, 1); }
    if (p.fd_kind == CSI_FD_FIELDS) { pack(f.fd_u, F_FDU, 0, 1, 0); pack(f.fd_v, F_FDV, 0, 0, 1); }
    k_pack<<<dim3((g.Nx + 2 * w + 2 + 127) / 128, g.Ny + 2 * pl->oy, PL.n), 128, 0, c.stream>>>(PL, P, w);
    ++*c.launches;
    // stage constants of the substep loop (ice mass, top-stress terms)
    k_prep<<<dim3((pl->pitch + 127) / 128, pl->rows), 128, 0, c.stream>>>(P, p.top_kind == CSI_STRESS_CONST);
    ++*c.launches;

    // interior tile columns / rows: every test the edge code makes is trivially true (see Params::it_x0)
    auto interior_range = [&](int ntiles, int first, int out, int s0, int s1, int v0, int v1, int c0, int c1, bool per, bool wall_lo, bool wall_hi, int N,
                              int &lo, int &hi) {
        lo = 1 << 30; hi = -1;
        for (int k = 0; k < ntiles; k++) {
            const int a = first + k * out, b = a + out - 1;  // output cells a..b; velocity nodes a-1..b+1 are computed
            bool ok = a >= std::max(s0, v0) && b <= std::min(s1, v1);
            ok = ok && (!per || (a > W && b <= N - W));
            ok = ok && (!wall_lo || a - 1 > 1) && (!wall_hi || b + 1 < N);
            ok = ok && a - 1 >= c0 && b + 1 <= c1;
            if (ok) { lo = std::min(lo, k); hi = std::max(hi, k); }
            else if (hi >= 0) break;  // keep the range contiguous
        }
    };
    const int y0 = P.sy0 < P.vy0 ? P.sy0 : P.vy0;
    interior_range((int)grid.x, P.a0, OUTX, P.sx0, P.sx1, P.vx0, P.vx1, P.cx0, P.cx1, P.px != 0, P.wall_w != 0, P.wall_e != 0, g.Nx, P.it_x0, P.it_x1);
    interior_range((int)grid.y, y0, OUTY, P.sy0, P.sy1, P.vy0, P.vy1, P.cy0, P.cy1, P.py != 0, P.wall_s != 0, P.wall_n != 0, g.Ny, P.it_y0, P.it_y1);

    pl->grid = grid;
    pl->cur_set = 0;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, nerr, "pack: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}

// `nsub` substeps starting at substep index `first_sub` (odd substeps update v first, se.jl:178-187);
// `aux_last`: the last of them also writes alpha, zeta, Delta.
// `halo_ready` (slabs): an event recorded on another stream after the halo rows of the current copy have been received.
// The first substep is then launched in row bands: the tile rows whose boxes stay inside rows 1..Ny go first and overlap
// the exchange, the tile rows that read halo rows wait for the event.
// (Replaying a block of substeps as a CUDA graph was measured on BASELINE config 1 as shipped -- 128 x 128, 50 tiles -- and
// changed nothing: 5.029 vs 5.030 ms per time_step!.  Such grids are bound by the latency of one tile pass, about 10 us with
// two warps per scheduler, not by the gaps between launches; the graph path was removed again.)
int fused_steps(FusedPlan *pl, const LaunchCtx &c, int first_sub, int nsub, bool aux_last, char *err, int nerr, cudaEvent_t halo_ready)
{
    using namespace fz;
    Params &P = pl->P;
    const dim3 grid = pl->grid;
    const int y0 = P.sy0 < P.vy0 ? P.sy0 : P.vy0;
    // tile row t covers rows J0 = y0 + OUTY t ... and reads J0 - 2 .. J0 + OUTY + 1
    int t_lo = 0, t_hi = (int)grid.y;  // interior band [t_lo, t_hi)
    while (t_lo < (int)grid.y && y0 + OUTY * t_lo - 2 < 1) t_lo++;
    while (t_hi > t_lo && y0 + OUTY * (t_hi - 1) + OUTY + 1 > P.Ny) t_hi--;
    // the planes one substep reads and writes: they alternate with the parity of the substep
    auto point = [&](Params &P, int sub, int cur_set) {
        P.in_set = cur_set;
        P.out_set = cur_set ^ 1;
        const bool vfirst = (sub % 2) != 0;
        {
            const size_t plane = (size_t)P.pitch * P.rows;
            auto at = [&](int f) { return P.base + (size_t)f * plane; };
            const int fo = P.out_set ? F_U1 : F_U0;
            // x: the planes of a u phase, y: of a v phase
            PhasePtrs X, Y;
            X.n = at(F_UN); Y.n = at(F_VN);
            X.tt = at(F_TX); Y.tt = at(F_TY);
            X.t1 = P.top_sis ? X.tt : at(F_T1X); Y.t1 = P.top_sis ? Y.tt : at(F_T1Y);   // (a top SemiImplicitStress reads u_a, v_a themselves)
            X.sa = at(F_SVA); Y.sa = at(F_SUA); X.oth = Y.tt; Y.oth = X.tt;
            X.tb1 = at(F_TB1X); Y.tb1 = at(F_TB1Y); X.tbr = at(F_UE); Y.tbr = at(F_VE);
            X.fd = at(F_UFD); Y.fd = at(F_VFD);
            X.rm = at(F_RM2U); Y.rm = at(F_RM2V); X.ue = at(F_UE); Y.ue = at(F_VE); X.sv = at(F_SVE); Y.sv = at(F_SUE);
            P.pc = vfirst ? Y : X;
            P.pd = vfirst ? X : Y;
            P.o_c = at(fo + (vfirst ? 1 : 0)); P.o_d = at(fo + (vfirst ? 0 : 1));
            P.o_s11 = at(fo + 2); P.o_s22 = at(fo + 3); P.o_s12 = at(fo + 4);
            P.g_rmc = at(F_RMC); P.g_rmf = at(F_RMF); P.g_pf4 = at(F_PF4);
        }
    };
    // Small grids on one rank: all substeps but the last (which also writes alpha, zeta, Delta) in one cooperative launch
    int k0 = 0;
    {
        const bool common0 = !P.fd_on && P.sis && P.use_ue && P.use_top && P.cor == CSI_CORIOLIS_FPLANE && P.pform == CSI_REPLACEMENT_PRESSURE && !P.flags;
        const int npers = aux_last ? nsub - 1 : nsub;
        if (pl->persistent < 0) {
            pl->persistent = 0;
            const char *env = getenv("CSI_PERSISTENT");   // 0 disables (A/B measurements)
            int dev = 0, sms = 0, per_sm = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            int coop = 0;
            cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
            const bool eligible = coop && !P.met && !P.met2 && !(env && env[0] == '0') && pl->fold_n[0] == 0 && pl->fold_n[1] == 0 &&
                                  !pl->partitioned;
            if (eligible) {
                cudaError_t e1 = common0 ? cudaFuncSetAttribute(k_evp_substeps_persistent<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES)
                                         : cudaFuncSetAttribute(k_evp_substeps_persistent<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
                if (e1 == cudaSuccess)
                    e1 = common0 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_evp_substeps_persistent<false>, NT, SMEM_BYTES)
                                 : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_evp_substeps_persistent<true>, NT, SMEM_BYTES);
                if (e1 == cudaSuccess && (long long)grid.x * grid.y <= (long long)sms * per_sm) pl->persistent = common0 ? 1 : 2;
                cudaGetLastError();
            }
        }
        if (pl->persistent > 0 && (pl->persistent == 1) == common0 && npers >= 2 && !halo_ready) {
            Params Pe = P, Po = P;
            Pe.ty0 = Po.ty0 = 0;
            // substep `first_sub + k` reads copy (cur_set + k) & 1
            const int par0 = first_sub & 1;
            point(par0 ? Po : Pe, first_sub, pl->cur_set);
            point(par0 ? Pe : Po, first_sub + 1, pl->cur_set ^ 1);
            unsigned int *sync = reinterpret_cast<unsigned int *>(pl->invalid + 2);
            cudaMemsetAsync(sync, 0, sizeof(unsigned int), c.stream);
            int fs = first_sub, ns = npers;
            void *args[] = {(void *)&pl->tmap, (void *)&Pe, (void *)&Po, (void *)&fs, (void *)&ns, (void *)&sync};
            cudaError_t e = pl->persistent == 1 ? cudaLaunchCooperativeKernel((void *)k_evp_substeps_persistent<false>, grid, dim3(NT), args, SMEM_BYTES, c.stream)
                                                : cudaLaunchCooperativeKernel((void *)k_evp_substeps_persistent<true>, grid, dim3(NT), args, SMEM_BYTES, c.stream);
            if (e == cudaSuccess) {
                ++*c.launches;
                ++pl->persistent_launches;
                pl->cur_set ^= (npers & 1);
                k0 = npers;
            } else {
                // (e.g. fewer SMs available to this context than the occupancy query assumed: nothing was launched; the
                // substeps below run one launch each, now and from here on)
                cudaGetLastError();
                pl->persistent = 0;
            }
        }
    }
    for (int k = k0; k < nsub; k++) {
        const int sub = first_sub + k;
        point(P, sub, pl->cur_set);
        const bool vfirst = (sub % 2) != 0;
        const bool aux = aux_last && k == nsub - 1;
        // the common configuration runs the variant compiled without run-time switches
        const bool common = !P.fd_on && P.sis && P.use_ue && P.use_top && P.cor == CSI_CORIOLIS_FPLANE && P.pform == CSI_REPLACEMENT_PRESSURE && !P.flags && !P.met && !P.met2;
        const bool common_met = !P.fd_on && P.sis && P.use_ue && P.use_top && P.cor == CSI_CORIOLIS_SPHERICAL && P.pform == CSI_REPLACEMENT_PRESSURE && !P.flags && P.met;
        auto band_on = [&](cudaStream_t st, int t0, int t1, bool aux) -> cudaError_t {
            if (t1 <= t0) return cudaSuccess;
            P.ty0 = t0;
            const dim3 gb(grid.x, t1 - t0);
            ++*c.launches;
            if (P.met2) return launch_sub<true, 2>(pl, P, gb, st, vfirst, aux);   // (orthogonal curvilinear grids keep the run-time switches)
            if (P.met) return common_met ? launch_sub<false, 1>(pl, P, gb, st, vfirst, aux) : launch_sub<true, 1>(pl, P, gb, st, vfirst, aux);
            return common ? launch_sub<false, 0>(pl, P, gb, st, vfirst, aux) : launch_sub<true, 0>(pl, P, gb, st, vfirst, aux);
        };
        auto band = [&](int t0, int t1, bool aux) -> cudaError_t { return band_on(c.stream, t0, t1, aux); };
        cudaError_t e;
        if (pl->fold_n[0] > 0 || pl->fold_n[1] > 0) {
            // A mesh with a fold.  The reference fills halos -- the fold included -- between the two velocity updates of a substep
            // (se:178-187), and the second one reads, next to the fold, first-velocity values the fold has just replaced by those of
            // mirrored columns: another tile's.  Tile rows that can see such a value run the substep in two launches with the
            // fold fill between them; the rows below them (the bulk) run the one-launch substep, the last of them also storing
            // alpha, which phase D alone reads back one row below its own tiles.
            if (k == 0 && halo_ready && (e = cudaStreamWaitEvent(c.stream, halo_ready, 0)) != cudaSuccess) { snprintf(err, nerr, "wait: %s", cudaGetErrorString(e)); return (int)e; }
            int tf = 0;   // first tile row of the split band: its tile, halo ring included, reaches row fold_jmin - 1 or beyond
            while (tf < (int)grid.y && y0 + OUTY * tf + OUTY + 1 < pl->fold_jmin - 1) tf++;
            const int wf = vfirst ? 1 : 0, ws = vfirst ? 0 : 1;   // list (0: u, 1: v) of the first / second velocity of this substep
            const size_t plane = (size_t)P.pitch * P.rows;
            auto fold_fill = [&](const LaunchCtx &cc, int w) {
                if (pl->fold_n[w] <= 0) return;
                DArr a;
                a.p = P.base + (size_t)((P.out_set ? F_U1 : F_U0) + w) * plane;
                a.sx = P.pitch; a.sy = P.rows; a.ox = OX; a.oy = P.oy;
                launch_fold_list(cc, a, pl->fold_t[w], pl->fold_s[w], pl->fold_n[w], pl->fold_sign);
            };
            // Everything the band reads that this substep writes -- the sources of the fold fills, and alpha / sigma / the first
            // velocity one row below its tiles -- comes from the band's own launches or from the tile row right below it (the one
            // that also stores alpha); both read the previous substep's planes only, like the bulk.  So tile row tf - 1 and the
            // band run as a chain on a stream of their own, beside the bulk launch, and join it before the next substep.
            const bool beside = pl->side && tf >= 2 && tf < (int)grid.y && y0 + OUTY * (tf - 1) <= pl->fold_jsrc_min;
            const cudaStream_t bs = beside ? pl->side : c.stream;
            const LaunchCtx cb{bs, c.launches};
            if (beside) {
                e = cudaEventRecord(pl->ev_fork, c.stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->side, pl->ev_fork, 0);
            } else e = band(0, tf - 1, aux);
            if (e == cudaSuccess) e = band_on(bs, std::max(tf - 1, 0), tf, true);
            if (e == cudaSuccess && tf < (int)grid.y) {
                P.ty0 = tf;
                const dim3 gb(grid.x, grid.y - tf);
                ++*c.launches;
                e = launch_split<1>(pl, P, gb, bs, vfirst, aux);
                fold_fill(cb, wf);
                ++*c.launches;
                if (e == cudaSuccess) e = launch_split<2>(pl, P, gb, bs, vfirst, aux);
                fold_fill(cb, ws);
            } else if (e == cudaSuccess) {
                fold_fill(cb, wf);
                fold_fill(cb, ws);
            }
            if (beside) {
                if (e == cudaSuccess) e = cudaEventRecord(pl->ev_join, pl->side);
                if (e == cudaSuccess) e = band(0, tf - 1, aux);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, pl->ev_join, 0);
            }
        } else if (k == 0 && halo_ready && t_hi > t_lo) {
            e = band(t_lo, t_hi, aux);
            if (e == cudaSuccess) e = cudaStreamWaitEvent(c.stream, halo_ready, 0);
            if (e == cudaSuccess) e = band(0, t_lo, aux);
            if (e == cudaSuccess) e = band(t_hi, (int)grid.y, aux);
        } else {
            if (k == 0 && halo_ready && (e = cudaStreamWaitEvent(c.stream, halo_ready, 0)) != cudaSuccess) { snprintf(err, nerr, "wait: %s", cudaGetErrorString(e)); return (int)e; }
            e = band(0, (int)grid.y, aux);
        }
        if (e != cudaSuccess) { snprintf(err, nerr, "launch: %s", cudaGetErrorString(e)); return (int)e; }
        pl->cur_set ^= 1;
    }
    return 0;
}

// diagnostics of the last stage: out[0] = inputs failed validation (whole stage on the IEEE pass), out[1] = tile passes
// redone with the IEEE operators, out[2] = tiles per substep.  Synchronises the device.
void fused_stats(const FusedPlan *pl, long long out[3])
{
    int v[2] = {0, 0};
    cudaDeviceSynchronize();
    cudaMemcpy(v, pl->invalid, sizeof v, cudaMemcpyDeviceToHost);
    out[0] = v[0];
    out[1] = v[1];
    out[2] = (long long)pl->grid.x * pl->grid.y;
}

// views of the current copy of the evolving fields (u, v, s11, s22, s12) in the internal layout,
// for the slab halo exchange between blocks of substeps
void fused_views(const FusedPlan *pl, DArr out[5])
{
    using namespace fz;
    const size_t plane = (size_t)pl->pitch * pl->rows;
    for (int k = 0; k < 5; k++) {
        out[k].p = pl->base + (size_t)((pl->cur_set ? F_U1 : F_U0) + k) * plane;
        out[k].sx = pl->pitch;
        out[k].sy = pl->rows;
        out[k].ox = OX;
        out[k].oy = pl->oy;
    }
}

int fused_end(FusedPlan *pl, const LaunchCtx &c, const DFields &f, char *err, int nerr)
{
    using namespace fz;
    Params &P = pl->P;
    const int in_set = pl->cur_set;
    const int nsub = 1;
    // unpack the final copy into the caller's arrays (interior / stress window); halos are refilled by the caller
    UnpackList UL;
    UL.n = 0;
    int wmax = 0, hmax = 0;
    auto unpack = [&](const DArr &a, int field, int i0, int i1, int j0, int j1) {
        if (!a.p) return;
        UL.it[UL.n++] = UnpackItem{a, field, i0, i1, j0, j1};
        wmax = std::max(wmax, i1 - i0 + 1);
        hmax = std::max(hmax, j1 - j0 + 1);
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
        k_unpack<<<dim3((wmax + 127) / 128, hmax, UL.n), 128, 0, c.stream>>>(UL, P);
        ++*c.launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(err, nerr, "%s", cudaGetErrorString(e)); return (int)e; }
    return 0;
}


int fused_run(FusedPlan *pl, const LaunchCtx &c, const DGrid &g, const DParams &p, const DFields &f, double dt, int first_sub, int nsub,
              char *err, int nerr)
{
    int rc = fused_begin(pl, c, g, p, f, dt, err, nerr);
    if (rc) return rc;
    if ((rc = fused_steps(pl, c, first_sub, nsub, true, err, nerr, nullptr))) return rc;
    return nsub > 0 ? fused_end(pl, c, f, err, nerr) : 0;
}

}  // namespace csi
