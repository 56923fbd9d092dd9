/*
 * csi_oracle.c -- ORACLE (test infrastructure only; see the header of csi_oracle.h).
 *
 * Plain-C restatement of ClimaSeaIce.jl v0.5.8's split-explicit EVP momentum substep loop and
 * h/aice advection update.  "ref:" comments give the file:line under /root/reference that each
 * function follows; "[OCN-recall]" marks Oceananigans primitives restated from SURVEY.md
 * Appendix A (their source is not in the image).  PARITY UNPINNED (no Julia, no golden vectors).
 *
 * Deliberately literal: every operator is re-evaluated at every neighbour as the reference's
 * inlined Julia does, every expression keeps Julia's left-to-right association, and both branches
 * of every `ifelse` are evaluated.  Build: gcc -O2 -ffp-contract=off -fopenmp (see Makefile).
 */
#include "csi_oracle.h"
#include <math.h>
#include <quadmath.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define F(f, i, j) ((f).p[(size_t)((i)-1 + (f).ox) + (size_t)((j)-1 + (f).oy) * (size_t)(f).sx])
typedef const csio_grid *G;
typedef const csio_params *P;

int csio_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * Julia scalar semantics (last rows)
 * ------------------------------------------------------------------------------------------- */
static inline double jl_max(double a, double b)
{ /* Base.max: NaN-propagating, max(-0.0, +0.0) = +0.0 */
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}
static inline double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
/* x * active::Bool -> ifelse(active, x, copysign(0, x)) */
static inline double jl_mul_bool(double x, int b) { return b ? x : copysign(0.0, x); }

/* exp: Julia's Base.exp is not available; both oracle and product use a correctly rounded exp so
 * that any two correct implementations agree bit for bit.  Here: binary128 expq, rounded once. */
double csio_exp(double x) { return (double)expq((__float128)x); }

/* ---------------------------------------------------------------------------------------------
 * Grid metrics [OCN-recall]: regular RectilinearGrid (all spacings scalar, Az = dx*dy) or j-indexed
 * arrays (LatitudeLongitudeGrid).
 * ------------------------------------------------------------------------------------------- */
/* CSIO_JMETRIC: a[j-1+Hy]; CSIO_IJMETRIC: two-dimensional arrays, metW columns per row, entry (i, j) at [(j-1+Hy) metW + (i-1+Hx)] */
#define MJ(a) (g->metric_kind == CSIO_IJMETRIC ? g->a[(size_t)(j - 1 + g->Hy) * g->metW + (size_t)(i - 1 + g->Hx < 0 ? 0 : (i - 1 + g->Hx >= g->metW ? g->metW - 1 : i - 1 + g->Hx))] \
                                               : g->a[j - 1 + g->Hy])
static inline double dxcc(G g, int i, int j) { return g->metric_kind ? MJ(dxcc) : g->dx; }
static inline double dxfc(G g, int i, int j) { return g->metric_kind ? MJ(dxfc) : g->dx; }
static inline double dxcf(G g, int i, int j) { return g->metric_kind ? MJ(dxcf) : g->dx; }
static inline double dxff(G g, int i, int j) { return g->metric_kind ? MJ(dxff) : g->dx; }
static inline double dycc(G g, int i, int j) { return g->metric_kind ? MJ(dycc) : g->dy; }
static inline double dyfc(G g, int i, int j) { return g->metric_kind ? MJ(dyfc) : g->dy; }
static inline double dycf(G g, int i, int j) { return g->metric_kind ? MJ(dycf) : g->dy; }
static inline double dyff(G g, int i, int j) { return g->metric_kind ? MJ(dyff) : g->dy; }
static inline double azcc(G g, int i, int j) { return g->metric_kind ? MJ(azcc) : g->dx * g->dy; }
static inline double azfc(G g, int i, int j) { return g->metric_kind ? MJ(azfc) : g->dx * g->dy; }
static inline double azcf(G g, int i, int j) { return g->metric_kind ? MJ(azcf) : g->dx * g->dy; }
static inline double azff(G g, int i, int j) { return g->metric_kind ? MJ(azff) : g->dx * g->dy; }
/* Flat z: dz = 1, so Ax = dy*1, Ay = dx*1, V = Az*1 */
static inline double axfcc(G g, int i, int j) { return dyfc(g, i, j) * 1.0; }
static inline double aycfc(G g, int i, int j) { return dxcf(g, i, j) * 1.0; }
static inline double vccc(G g, int i, int j) { return azcc(g, i, j) * 1.0; }

/* ---------------------------------------------------------------------------------------------
 * Node activity [OCN-recall]: inactive_cell / peripheral_node / immersed_peripheral_node
 * ------------------------------------------------------------------------------------------- */
static inline int outside_domain(G g, int i, int j)
{
    return (g->topo_x == CSIO_BOUNDED && (i < 1 || i > g->Nx)) || (g->topo_y == CSIO_BOUNDED && (j < 1 || j > g->Ny)) ||
           (g->topo_y == CSIO_FOLDED && j < 1); /* nothing is outside beyond a fold */
}
static inline int immersed_cell(G g, int i, int j)
{
    if (!g->mask) return 0;
    int sx = g->Nx + 2 * g->Hx, sy = g->Ny + 2 * g->Hy;
    int pi = i - 1 + g->Hx, pj = j - 1 + g->Hy;
    if (pi < 0) pi = 0;
    if (pj < 0) pj = 0;
    if (pi >= sx) pi = sx - 1;
    if (pj >= sy) pj = sy - 1;
    return g->mask[(size_t)pi + (size_t)pj * (size_t)sx] != 0;
}
static inline int inactive_cell(G g, int i, int j) { return outside_domain(g, i, j) || immersed_cell(g, i, j); }
static inline int peripheral_fc(G g, int i, int j) { return inactive_cell(g, i - 1, j) | inactive_cell(g, i, j); }
static inline int peripheral_cf(G g, int i, int j) { return inactive_cell(g, i, j - 1) | inactive_cell(g, i, j); }
/* immersed_peripheral_node = peripheral on the immersed grid & not peripheral on the underlying grid */
static inline int imm_peripheral_cc(G g, int i, int j) { return inactive_cell(g, i, j) & !outside_domain(g, i, j); }
static inline int imm_peripheral_ff(G g, int i, int j)
{
    int per = inactive_cell(g, i - 1, j - 1) | inactive_cell(g, i, j - 1) | inactive_cell(g, i - 1, j) | inactive_cell(g, i, j);
    int und = outside_domain(g, i - 1, j - 1) | outside_domain(g, i, j - 1) | outside_domain(g, i - 1, j) | outside_domain(g, i, j);
    return per & !und;
}
static inline int imm_peripheral_fc(G g, int i, int j)
{
    return peripheral_fc(g, i, j) & !(outside_domain(g, i - 1, j) | outside_domain(g, i, j));
}
static inline int imm_peripheral_cf(G g, int i, int j)
{
    return peripheral_cf(g, i, j) & !(outside_domain(g, i, j - 1) | outside_domain(g, i, j));
}

/* ---------------------------------------------------------------------------------------------
 * ice mass.  ref: src/ClimaSeaIce.jl:42   ice_mass = h * rho * aice  (left to right)
 * ------------------------------------------------------------------------------------------- */
static inline double ice_mass(P p, const csio_state *s, int i, int j) { return F(s->h, i, j) * p->rho_ice * F(s->a, i, j); }

/* ---------------------------------------------------------------------------------------------
 * Strain rates.  ref: src/Rheologies/elasto_visco_plastic_rheology.jl:360-375
 * [OCN-recall] delta_x^c(q)_i = q_{i+1}-q_i ; delta_x^f(q)_i = q_i - q_{i-1}
 * ------------------------------------------------------------------------------------------- */
static inline double eps_D(G g, csio_field u, csio_field v, int i, int j)
{ /* evp.jl:365 */
    return ((dyfc(g, i + 1, j) * F(u, i + 1, j) - dyfc(g, i, j) * F(u, i, j)) +
            (dxcf(g, i, j + 1) * F(v, i, j + 1) - dxcf(g, i, j) * F(v, i, j))) /
           azcc(g, i, j);
}
static inline double eps_T(G g, csio_field u, csio_field v, int i, int j)
{ /* evp.jl:367-368 */
    double dy = dycc(g, i, j), dx = dxcc(g, i, j);
    return (dy * dy * (F(u, i + 1, j) / dyfc(g, i + 1, j) - F(u, i, j) / dyfc(g, i, j)) -
            dx * dx * (F(v, i, j + 1) / dxcf(g, i, j + 1) - F(v, i, j) / dxcf(g, i, j))) /
           azcc(g, i, j);
}
static inline double eps_S(G g, csio_field u, csio_field v, int i, int j)
{ /* evp.jl:370-371 */
    double dx = dxff(g, i, j), dy = dyff(g, i, j);
    return (dx * dx * (F(u, i, j) / dxfc(g, i, j) - F(u, i, j - 1) / dxfc(g, i, j - 1)) +
            dy * dy * (F(v, i, j) / dycf(g, i, j) - F(v, i - 1, j) / dycf(g, i - 1, j))) /
           azff(g, i, j);
}
static inline double strain_rate_xx(G g, csio_field u, csio_field v, int i, int j) { return (eps_D(g, u, v, i, j) + eps_T(g, u, v, i, j)) / 2; }
static inline double strain_rate_yy(G g, csio_field u, csio_field v, int i, int j) { return (eps_D(g, u, v, i, j) - eps_T(g, u, v, i, j)) / 2; }
static inline double strain_rate_xy(G g, csio_field u, csio_field v, int i, int j) { return eps_S(g, u, v, i, j) / 2; }

/* [OCN-recall] 4-point averages: y-average of x-averages */
#define IXY_FF(fn, i, j) (((fn((i)-1, (j)-1) + fn((i), (j)-1)) / 2 + (fn((i)-1, (j)) + fn((i), (j))) / 2) / 2)
#define IXY_CC(fn, i, j) (((fn((i), (j)) + fn((i) + 1, (j))) / 2 + (fn((i), (j) + 1) + fn((i) + 1, (j) + 1)) / 2) / 2)
#define IXY_FC(fn, i, j) (((fn((i)-1, (j)) + fn((i), (j))) / 2 + (fn((i)-1, (j) + 1) + fn((i), (j) + 1)) / 2) / 2)
#define IXY_CF(fn, i, j) (((fn((i), (j)-1) + fn((i) + 1, (j)-1)) / 2 + (fn((i), (j)) + fn((i) + 1, (j))) / 2) / 2)

/* ---------------------------------------------------------------------------------------------
 * initialize_rheology!.  ref: evp.jl:192-219.  Whole parent of P (evp.jl:166-167).
 * ------------------------------------------------------------------------------------------- */
int csio_initialize_rheology(G g, P p, csio_state *s)
{
    int i0 = 1 - g->Hx, i1 = g->Nx + g->Hx, j0 = 1 - g->Hy, j1 = g->Ny + g->Hy;
#pragma omp parallel for schedule(static)
    for (int j = j0; j <= j1; j++)
        for (int i = i0; i <= i1; i++) {
            /* ice_strength: P* h exp(-C (1 - aice))   evp.jl:219 */
            F(s->P, i, j) = p->Pstar * F(s->h, i, j) * csio_exp(-p->C * (1 - F(s->a, i, j)));
            F(s->un, i, j) = F(s->u, i, j);
            F(s->vn, i, j) = F(s->v, i, j);
        }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * _compute_evp_viscosities!.  ref: evp.jl:236-273.  Range -H+2 : N+H-1 (evp.jl:145).
 * ------------------------------------------------------------------------------------------- */
static void compute_evp_viscosities(G g, P p, csio_state *s)
{
    const double em2 = (1 / p->e) * (1 / p->e); /* e^(-2) = inv(e)^2 */
    const double Dm = p->Dmin;
    csio_field u = s->u, v = s->v;
#pragma omp parallel for schedule(static)
    for (int j = -g->Hy + 2; j <= g->Ny + g->Hy - 1; j++)
        for (int i = -g->Hx + 2; i <= g->Nx + g->Hx - 1; i++) {
#define EXX(I_, J_) strain_rate_xx(g, u, v, (I_), (J_))
#define EYY(I_, J_) strain_rate_yy(g, u, v, (I_), (J_))
#define EXY(I_, J_) strain_rate_xy(g, u, v, (I_), (J_))
#define PP(I_, J_) F(s->P, (I_), (J_))
            double e11c = EXX(i, j), e22c = EYY(i, j), e12f = EXY(i, j);
            double e11f = IXY_FF(EXX, i, j), e22f = IXY_FF(EYY, i, j), e12c = IXY_CC(EXY, i, j);
            double dc = e11c + e22c, df = e11f + e22f;
            double sc = sqrt((e11c - e22c) * (e11c - e22c) + 4 * (e12c * e12c));
            double sf = sqrt((e11f - e22f) * (e11f - e22f) + 4 * (e12f * e12f));
            double Dc = jl_max(sqrt(dc * dc + sc * sc * em2), Dm);
            double Df = jl_max(sqrt(df * df + sf * sf * em2), Dm);
            double Pc = PP(i, j), Pf = IXY_FF(PP, i, j);
            F(s->zf, i, j) = Pf / (2 * Df);
            F(s->zc, i, j) = Pc / (2 * Dc);
            F(s->delta, i, j) = Dc;
        }
}

/* ice_pressure.  ref: evp.jl:282-289 */
static inline double ice_pressure(P p, const csio_state *s, int i, int j)
{
    double Pc = F(s->P, i, j);
    if (p->pressure_formulation == CSIO_ICE_STRENGTH) return Pc;
    double Dc = F(s->delta, i, j);
    return Pc * Dc / (Dc + p->Dmin);
}

/* ---------------------------------------------------------------------------------------------
 * _compute_evp_stresses!.  ref: evp.jl:294-354.  Same range.
 * ------------------------------------------------------------------------------------------- */
static void compute_evp_stresses(G g, P p, csio_state *s, double dt)
{
    const double em2 = (1 / p->e) * (1 / p->e);
    const double ap = p->alpha_max, am = p->alpha_min, ca = p->c_alpha;
    csio_field u = s->u, v = s->v;
#pragma omp parallel for schedule(static)
    for (int j = -g->Hy + 2; j <= g->Ny + g->Hy - 1; j++)
        for (int i = -g->Hx + 2; i <= g->Nx + g->Hx - 1; i++) {
            double e11 = strain_rate_xx(g, u, v, i, j), e22 = strain_rate_yy(g, u, v, i, j), e12 = strain_rate_xy(g, u, v, i, j);
            double zc = F(s->zc, i, j), zf = F(s->zf, i, j);
            double Pr = ice_pressure(p, s, i, j);
            double ec = zc * em2, ef = zf * em2;
            double s11n = 2 * ec * e11 + ((zc - ec) * (e11 + e22) - Pr / 2);
            double s22n = 2 * ec * e22 + ((zc - ec) * (e11 + e22) - Pr / 2);
            double s12n = 2 * ef * e12;
#define MM(I_, J_) ice_mass(p, s, (I_), (J_))
            double mc = MM(i, j), mf = IXY_FF(MM, i, j);
            double g2c = zc * ca * dt / mc / azcc(g, i, j);
            g2c = (g2c != g2c) ? ap * ap : g2c;
            double gc = jl_clamp(sqrt(g2c), am, ap);
            double g2f = zf * ca * dt / mf / azff(g, i, j);
            g2f = (g2f != g2f) ? ap * ap : g2f;
            double gf = jl_clamp(sqrt(g2f), am, ap);
            double d11 = (s11n - F(s->s11, i, j)) / gc;
            double d22 = (s22n - F(s->s22, i, j)) / gc;
            double d12 = (s12n - F(s->s12, i, j)) / gf;
            F(s->s11, i, j) += (mc > 0 ? d11 : 0.0);
            F(s->s22, i, j) += (mc > 0 ? d22 : 0.0);
            F(s->s12, i, j) += (mf > 0 ? d12 : 0.0);
            F(s->alpha, i, j) = gc;
        }
}

/* compute_stresses!.  ref: evp.jl:222-234 */
int csio_compute_stresses(G g, P p, csio_state *s, double dt)
{
    compute_evp_viscosities(g, p, s);
    compute_evp_stresses(g, p, s, dt);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Stress divergence.  ref: src/Rheologies/ice_stress_divergence.jl:16-51
 * On immersed grids the stresses are masked to 0 on immersed-peripheral nodes (isd.jl:21-24).
 * ------------------------------------------------------------------------------------------- */
static inline double stress_cc(G g, csio_field f, int i, int j) { return (g->mask && imm_peripheral_cc(g, i, j)) ? 0.0 : F(f, i, j); }
static inline double stress_ff(G g, csio_field f, int i, int j) { return (g->mask && imm_peripheral_ff(g, i, j)) ? 0.0 : F(f, i, j); }
static inline double sigD(G g, const csio_state *s, int i, int j) { return stress_cc(g, s->s11, i, j) + stress_cc(g, s->s22, i, j); }
static inline double sigT(G g, const csio_state *s, int i, int j) { return stress_cc(g, s->s11, i, j) - stress_cc(g, s->s22, i, j); }

static inline double div_sigma_1j(G g, const csio_state *s, int i, int j)
{ /* isd.jl:39-44 */
    double d = dyfc(g, i, j) * (sigD(g, s, i, j) - sigD(g, s, i - 1, j)) / 2;
    double t = (dycc(g, i, j) * dycc(g, i, j) * sigT(g, s, i, j) - dycc(g, i - 1, j) * dycc(g, i - 1, j) * sigT(g, s, i - 1, j)) / dyfc(g, i, j) / 2;
    double S = (dxff(g, i, j + 1) * dxff(g, i, j + 1) * stress_ff(g, s->s12, i, j + 1) - dxff(g, i, j) * dxff(g, i, j) * stress_ff(g, s->s12, i, j)) / dxfc(g, i, j);
    return (d + t + S) / azfc(g, i, j);
}
static inline double div_sigma_2j(G g, const csio_state *s, int i, int j)
{ /* isd.jl:46-51 */
    double d = dxcf(g, i, j) * (sigD(g, s, i, j) - sigD(g, s, i, j - 1)) / 2;
    double t = -(dxcc(g, i, j) * dxcc(g, i, j) * sigT(g, s, i, j) - dxcc(g, i, j - 1) * dxcc(g, i, j - 1) * sigT(g, s, i, j - 1)) / dxcf(g, i, j) / 2;
    double S = (dyff(g, i + 1, j) * dyff(g, i + 1, j) * stress_ff(g, s->s12, i + 1, j) - dyff(g, i, j) * dyff(g, i, j) * stress_ff(g, s->s12, i, j)) / dycf(g, i, j);
    return (d + t + S) / azcf(g, i, j);
}

/* ---------------------------------------------------------------------------------------------
 * External stresses.  ref: src/SeaIceDynamics/sea_ice_external_stress.jl:8-40,176-210
 * Either side (w = TOP: atmosphere, BOT: ocean) is nothing / a Number pair / a pair of arrays / a
 * SemiImplicitStress.  ext_x/ext_y is the side's array-or-constant: the stress itself for CONST and
 * FIELD, the external velocity u_e, v_e for a SemiImplicitStress.
 * ------------------------------------------------------------------------------------------- */
enum { TOP = 0, BOT = 1 };
static inline int stress_kind(P p, int w) { return w == TOP ? p->top_kind : p->bot_kind; }
static inline double ext_x(P p, int w, int i, int j)
{
    if (w == TOP) return p->top_x.p ? F(p->top_x, i, j) : p->top_tx;
    return p->ue.p ? F(p->ue, i, j) : p->ue_c;
}
static inline double ext_y(P p, int w, int i, int j)
{
    if (w == TOP) return p->top_y.p ? F(p->top_y, i, j) : p->top_ty;
    return p->ve.p ? F(p->ve, i, j) : p->ve_c;
}
static inline double ext_rho(P p, int w) { return w == TOP ? p->top_rho : p->rho_e; }
static inline double ext_Cd(P p, int w) { return w == TOP ? p->top_Cd : p->Cd; }
#define VV(I_, J_) F(s->v, (I_), (J_))
#define UU(I_, J_) F(s->u, (I_), (J_))

static inline double sis_speed_x(P p, const csio_state *s, int w, int i, int j)
{ /* sqrt(du^2+dv^2) at (f,c): ext.jl:177-180 */
#define EY(I_, J_) ext_y(p, w, (I_), (J_))
    double du = ext_x(p, w, i, j) - F(s->u, i, j);
    double dv = IXY_FC(EY, i, j) - IXY_FC(VV, i, j);
    return sqrt(du * du + dv * dv);
}
static inline double sis_speed_y(P p, const csio_state *s, int w, int i, int j)
{ /* at (c,f): ext.jl:183-187 */
#define EX(I_, J_) ext_x(p, w, (I_), (J_))
    double dv = ext_y(p, w, i, j) - F(s->v, i, j);
    double du = IXY_CF(EX, i, j) - IXY_CF(UU, i, j);
    return sqrt(du * du + dv * dv);
}
/* explicit_tau_x/y: ext.jl:13-20 (nothing -> 0, Number, array) and :176-188 (rho_e * Cd * sqrt(..) * u_e) */
static inline double explicit_tx(P p, const csio_state *s, int w, int i, int j)
{
    int k = stress_kind(p, w);
    if (k == CSIO_STRESS_NONE) return 0.0;
    if (k == CSIO_STRESS_SEMI_IMPLICIT) return ext_rho(p, w) * ext_Cd(p, w) * sis_speed_x(p, s, w, i, j) * ext_x(p, w, i, j);
    return ext_x(p, w, i, j);
}
static inline double explicit_ty(P p, const csio_state *s, int w, int i, int j)
{
    int k = stress_kind(p, w);
    if (k == CSIO_STRESS_NONE) return 0.0;
    if (k == CSIO_STRESS_SEMI_IMPLICIT) return ext_rho(p, w) * ext_Cd(p, w) * sis_speed_y(p, s, w, i, j) * ext_y(p, w, i, j);
    return ext_y(p, w, i, j);
}
/* implicit_tau_x/y_coefficient: ext.jl:8-9 (zero for everything but a SemiImplicitStress) and :192-202 */
static inline double implicit_tx(P p, const csio_state *s, int w, int i, int j)
{
    return stress_kind(p, w) == CSIO_STRESS_SEMI_IMPLICIT ? ext_rho(p, w) * ext_Cd(p, w) * sis_speed_x(p, s, w, i, j) : 0.0;
}
static inline double implicit_ty(P p, const csio_state *s, int w, int i, int j)
{
    return stress_kind(p, w) == CSIO_STRESS_SEMI_IMPLICIT ? ext_rho(p, w) * ext_Cd(p, w) * sis_speed_y(p, s, w, i, j) : 0.0;
}
/* x/y_momentum_stress of a stress that is NOT a SemiImplicitStress (ext.jl:34-38): explicit - zero(grid) * u */
static inline double x_momentum_stress(P p, const csio_state *s, int w, int i, int j)
{
    return explicit_tx(p, s, w, i, j) - 0.0 * F(s->u, i, j);
}
static inline double y_momentum_stress(P p, const csio_state *s, int w, int i, int j)
{
    return explicit_ty(p, s, w, i, j) - 0.0 * F(s->v, i, j);
}

/* ---------------------------------------------------------------------------------------------
 * Free drift.  ref: src/SeaIceDynamics/stress_balance_free_drift.jl:61-129
 *   nothing -> zero(grid) (:128-129);  (u=, v=) arrays -> the array value (:124-125);
 *   StressBalanceFreeDrift, repointed at the model's own stresses (:43-45, sime.jl:80): with `d` the
 *   side that is a SemiImplicitStress and `o` the other,  U_d - tau_o / sqrt(C_d * |tau_o|)  (:61-109).
 * ------------------------------------------------------------------------------------------- */
static inline double free_drift_u(P p, const csio_state *s, int i, int j)
{
    if (p->free_drift_kind == CSIO_FD_NONE) return 0.0;
    if (p->free_drift_kind == CSIO_FD_FIELDS) return F(p->fd_u, i, j);
    int d = p->bot_kind == CSIO_STRESS_SEMI_IMPLICIT ? BOT : TOP, o = 1 - d;
#define YMS(I_, J_) y_momentum_stress(p, s, o, (I_), (J_))
    double tx = x_momentum_stress(p, s, o, i, j);
    double ty = IXY_FC(YMS, i, j);
    double t = sqrt(tx * tx + ty * ty);
    double Ud = ext_x(p, d, i, j);
    double Cdrag = ext_rho(p, d) * ext_Cd(p, d);
    return Ud - (t == 0 ? t : tx / sqrt(Cdrag * t));
}
static inline double free_drift_v(P p, const csio_state *s, int i, int j)
{
    if (p->free_drift_kind == CSIO_FD_NONE) return 0.0;
    if (p->free_drift_kind == CSIO_FD_FIELDS) return F(p->fd_v, i, j);
    int d = p->bot_kind == CSIO_STRESS_SEMI_IMPLICIT ? BOT : TOP, o = 1 - d;
#define XMS(I_, J_) x_momentum_stress(p, s, o, (I_), (J_))
    double tx = IXY_CF(XMS, i, j);
    double ty = y_momentum_stress(p, s, o, i, j);
    double t = sqrt(tx * tx + ty * ty);
    double Ud = ext_y(p, d, i, j);
    double Cdrag = ext_rho(p, d) * ext_Cd(p, d);
    return Ud - (t == 0 ? t : ty / sqrt(Cdrag * t));
}

/* Coriolis [OCN-recall].
 *   FPlane:  x_f_cross_U = -f * Ixy^fc(v);  y_f_cross_U = +f * Ixy^cf(u)
 *   HydrostaticSphericalCoriolis, EnstrophyConserving scheme (f^ffa = 2 Omega sin(phi^f), supplied per row):
 *     x_f_cross_U = -Iy^c(f^ff) * Ix^f(Iy^c(dx^cf * v)) / dx^fc
 *     y_f_cross_U = +Ix^c(f^ff) * Iy^f(Ix^c(dy^fc * u)) / dy^cf
 *   (the active-cell-weighted variant differs from it only next to immersed cells and is not restated) */
#define FFF(J_) (p->f_ff[(J_)-1 + g->Hy])
static inline double x_f_cross_U(G g, P p, const csio_state *s, int i, int j)
{
    if (p->coriolis_kind == CSIO_CORIOLIS_NONE) return 0.0;
    if (p->coriolis_kind == CSIO_CORIOLIS_SPHERICAL) {
#define DXV(I_, J_) ((dxcf(g, (I_), (J_)) * VV((I_), (J_)) + dxcf(g, (I_), (J_) + 1) * VV((I_), (J_) + 1)) / 2)
        double fbar = (FFF(j) + FFF(j + 1)) / 2;
        return -fbar * ((DXV(i - 1, j) + DXV(i, j)) / 2) / dxfc(g, i, j);
    }
    return -p->f * IXY_FC(VV, i, j);
}
static inline double y_f_cross_U(G g, P p, const csio_state *s, int i, int j)
{
    if (p->coriolis_kind == CSIO_CORIOLIS_NONE) return 0.0;
    if (p->coriolis_kind == CSIO_CORIOLIS_SPHERICAL) {
#define DYU(I_, J_) ((dyfc(g, (I_), (J_)) * UU((I_), (J_)) + dyfc(g, (I_) + 1, (J_)) * UU((I_) + 1, (J_))) / 2)
        double fbar = (FFF(j) + FFF(j)) / 2;
        return fbar * ((DYU(i, j - 1) + DYU(i, j)) / 2) / dycf(g, i, j);
    }
    return p->f * IXY_CF(UU, i, j);
}

/* ---------------------------------------------------------------------------------------------
 * Immersed stress divergence.  ref: src/Rheologies/ice_stress_divergence.jl:57-123.  Without an
 * ImmersedBoundaryCondition it is zero(grid).  With the discrete-form FluxBoundaryCondition -C*u of
 * examples/ice_advected_on_coastline.jl:91-98 on south/north (u) and west/east (v), the other sides
 * `nothing`:  ib_*_south/west = -getbc, ib_*_north/east = +getbc (isd.jl:116-123), getbc = (-C)*u[i,j].
 * [OCN-recall] index_left(i, Center) = i, index_right(i, Center) = i+1; conditional_flux_ffc picks the
 * boundary flux on immersed-peripheral (f,f) nodes.
 * ------------------------------------------------------------------------------------------- */
static inline double immersed_div_sigma_1j(G g, P p, const csio_state *s, int i, int j)
{
    if (!g->mask || p->imm_drag_u == 0.0) return 0.0;
    double bc = (-p->imm_drag_u) * F(s->u, i, j);
    double qW = 0.0 * (dycc(g, i - 1, j) * 1.0), qE = 0.0 * (dycc(g, i, j) * 1.0);
    double qS = (imm_peripheral_ff(g, i, j) ? -bc : 0.0) * (dxff(g, i, j) * 1.0);
    double qN = (imm_peripheral_ff(g, i, j + 1) ? bc : 0.0) * (dxff(g, i, j + 1) * 1.0);
    return (qE - qW + qN - qS) / (azfc(g, i, j) * 1.0);
}
static inline double immersed_div_sigma_2j(G g, P p, const csio_state *s, int i, int j)
{
    if (!g->mask || p->imm_drag_v == 0.0) return 0.0;
    double bc = (-p->imm_drag_v) * F(s->v, i, j);
    double qW = (imm_peripheral_ff(g, i, j) ? -bc : 0.0) * (dyff(g, i, j) * 1.0);
    double qE = (imm_peripheral_ff(g, i + 1, j) ? bc : 0.0) * (dyff(g, i + 1, j) * 1.0);
    double qS = 0.0 * (dxcc(g, i, j - 1) * 1.0), qN = 0.0 * (dxcc(g, i, j) * 1.0);
    return (qE - qW + qN - qS) / (azcf(g, i, j) * 1.0);
}

/* ---------------------------------------------------------------------------------------------
 * u / v tendencies.  ref: src/SeaIceDynamics/momentum_tendencies_kernel_functions.jl:11-74
 * `dtau` is what the reference passes as "dt" (the local substep, se.jl:207-211), so the EVP
 * relaxation forcing (evp.jl:391-401) is evaluated literally as (un-u)/dtau/alpha_bar.
 * ------------------------------------------------------------------------------------------- */
static inline double u_velocity_tendency(G g, P p, const csio_state *s, int i, int j, double dtau)
{
#define AA(I_, J_) F(s->a, (I_), (J_))
#define AL(I_, J_) F(s->alpha, (I_), (J_))
    double ai = (AA(i, j) + AA(i - 1, j)) / 2;
    double mi = (MM(i, j) + MM(i - 1, j)) / 2;
    double user_forcing = 0.0;
    double rheology_forcing = (F(s->un, i, j) - F(s->u, i, j)) / dtau / ((AL(i, j) + AL(i - 1, j)) / 2);
    double Gu = -x_f_cross_U(g, p, s, i, j) - explicit_tx(p, s, TOP, i, j) / mi * ai + explicit_tx(p, s, BOT, i, j) / mi * ai +
                div_sigma_1j(g, s, i, j) / mi + immersed_div_sigma_1j(g, p, s, i, j) / mi + (user_forcing + rheology_forcing);
    return mi <= 0 ? 0.0 : Gu;
}
static inline double v_velocity_tendency(G g, P p, const csio_state *s, int i, int j, double dtau)
{
    double ai = (AA(i, j) + AA(i, j - 1)) / 2;
    double mi = (MM(i, j) + MM(i, j - 1)) / 2;
    double user_forcing = 0.0;
    double rheology_forcing = (F(s->vn, i, j) - F(s->v, i, j)) / dtau / ((AL(i, j) + AL(i, j - 1)) / 2);
    double Gv = -y_f_cross_U(g, p, s, i, j) - explicit_ty(p, s, TOP, i, j) / mi * ai + explicit_ty(p, s, BOT, i, j) / mi * ai +
                div_sigma_2j(g, s, i, j) / mi + immersed_div_sigma_2j(g, p, s, i, j) / mi + (user_forcing + rheology_forcing);
    return mi <= 0 ? 0.0 : Gv;
}

/* kernel index ranges: `:xy` = 1:N on serial grids (se.jl:31); widened on connected axes is the
 * caller's business (multi-rank tests pass their own state). */
#define DBL_EPS 2.220446049250313e-16

/* _u_velocity_step!.  ref: src/SeaIceDynamics/split_explicit_momentum_equations.jl:197-229 */
int csio_u_velocity_step(G g, P p, csio_state *s, double dt)
{
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            double mi = (MM(i, j) + MM(i - 1, j)) / 2;
            double ai = (AA(i, j) + AA(i - 1, j)) / 2;
            double dtau = dt / ((AL(i, j) + AL(i - 1, j)) / 2); /* evp.jl:384 */
            double Gu = u_velocity_tendency(g, p, s, i, j, dtau);
            double tau = (implicit_tx(p, s, BOT, i, j) - implicit_tx(p, s, TOP, i, j)) / mi * ai; /* se.jl:214-215 */
            tau = mi <= 0 ? 0.0 : tau;
            double uD = (F(s->u, i, j) + dtau * Gu) / (1 + dtau * tau);
            double uF = free_drift_u(p, s, i, j);
            int marginal = (mi > DBL_EPS) & (ai > DBL_EPS);
            int active_ice = (mi >= p->min_mass) & (ai >= p->min_conc);
            int active = !peripheral_fc(g, i, j);
            F(s->u, i, j) = jl_mul_bool(active_ice ? uD : (marginal ? uF : 0.0), active);
        }
    return 0;
}

/* _v_velocity_step!.  ref: split_explicit_momentum_equations.jl:231-264 */
int csio_v_velocity_step(G g, P p, csio_state *s, double dt)
{
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            double mi = (MM(i, j) + MM(i, j - 1)) / 2;
            double ai = (AA(i, j) + AA(i, j - 1)) / 2;
            double dtau = dt / ((AL(i, j) + AL(i, j - 1)) / 2); /* evp.jl:385 */
            double Gv = v_velocity_tendency(g, p, s, i, j, dtau);
            double tau = (implicit_ty(p, s, BOT, i, j) - implicit_ty(p, s, TOP, i, j)) / mi * ai;
            tau = mi <= 0 ? 0.0 : tau;
            double vD = (F(s->v, i, j) + dtau * Gv) / (1 + dtau * tau);
            double vF = free_drift_v(p, s, i, j);
            int marginal = (mi > DBL_EPS) & (ai > DBL_EPS);
            int active_ice = (mi >= p->min_mass) & (ai >= p->min_conc);
            int active = !peripheral_cf(g, i, j);
            F(s->v, i, j) = jl_mul_bool(active_ice ? vD : (marginal ? vF : 0.0), active);
        }
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * fill_halo_regions! [OCN-recall].  lx/ly: 0 = Center, 1 = Face.  which: 0 = default BCs
 * (no-flux on Center, nothing on Face along Bounded axes), 1 = u-velocity BCs, 2 = v-velocity BCs.
 * Non-periodic sides first over 1:N of the other axis, periodic sides last over the full parent
 * extent of the other axis (so corners hold periodic images).
 *   no-flux:  c[0] = c[1], c[N+1] = c[N]                                  (one halo cell)
 *   value:    c[0] = c[1] + ((c[1]-val)/(D/2))*(-D) ; c[N+1] = c[N] + ((val-c[N])/(D/2))*D
 *   impenetrable normal velocity: c[1] = 0, c[N+1] = 0
 * ------------------------------------------------------------------------------------------- */
static void fill_x_bounded(G g, P p, csio_field *f, int lx, int ly, int which)
{
    (void)ly;
    int Nx = g->Nx;
    for (int j = 1; j <= g->Ny; j++) {
        if (lx == 0) {
            if (which == 2 && p->v_we_bc == CSIO_BC_VALUE) {
                double val = p->v_we_val;
                double Dw = dxff(g, 1, j), De = dxff(g, Nx + 1, j);
                F(*f, 0, j) = F(*f, 1, j) + ((F(*f, 1, j) - val) / (Dw / 2)) * (-Dw);
                F(*f, Nx + 1, j) = F(*f, Nx, j) + ((val - F(*f, Nx, j)) / (De / 2)) * De;
            } else {
                F(*f, 0, j) = F(*f, 1, j);
                F(*f, Nx + 1, j) = F(*f, Nx, j);
            }
        } else if (which == 1) {
            F(*f, 1, j) = 0.0;
            F(*f, Nx + 1, j) = 0.0;
        }
    }
}
static void fill_y_bounded(G g, P p, csio_field *f, int lx, int ly, int which, int north)
{
    (void)lx;
    int Ny = g->Ny;
    for (int i = 1; i <= g->Nx; i++) {
        if (ly == 0) {
            if (which == 1 && p->u_sn_bc == CSIO_BC_VALUE) {
                double val = p->u_sn_val;
                double Ds = dyff(g, i, 1), Dn = dyff(g, i, Ny + 1);
                F(*f, i, 0) = F(*f, i, 1) + ((F(*f, i, 1) - val) / (Ds / 2)) * (-Ds);
                if (north) F(*f, i, Ny + 1) = F(*f, i, Ny) + ((val - F(*f, i, Ny)) / (Dn / 2)) * Dn;
            } else {
                F(*f, i, 0) = F(*f, i, 1);
                if (north) F(*f, i, Ny + 1) = F(*f, i, Ny);
            }
        } else if (which == 2) {
            F(*f, i, 1) = 0.0;
            if (north) F(*f, i, Ny + 1) = 0.0;
        }
    }
}
/* the north fold: a copy list with a sign (see csio_grid); `ext`: the field is an external stress / velocity array */
static void fill_fold(G g, csio_field *f, int lx, int ly, int which, int ext)
{
    const int loc = (lx ? 1 : 0) + (ly ? 2 : 0);
    const double sign = (which == 1 || which == 2) ? g->fold_sign_velocity : (ext ? g->fold_sign_external : 1.0);
    for (int k = 0; k < g->fold_count[loc]; k++) f->p[g->fold_target[loc][k]] = sign * f->p[g->fold_source[loc][k]];
}
static void fill_x_periodic(G g, csio_field *f)
{
    int Nx = g->Nx, Hx = g->Hx;
    for (int pj = 0; pj < f->sy; pj++) {
        int j = pj + 1 - f->oy;
        for (int k = 1; k <= Hx; k++) {
            F(*f, 1 - k, j) = F(*f, Nx + 1 - k, j);
            F(*f, Nx + k, j) = F(*f, k, j);
        }
    }
}
static void fill_y_periodic(G g, csio_field *f)
{
    int Ny = g->Ny, Hy = g->Hy;
    for (int k = 1; k <= Hy; k++)
        for (int pi = 0; pi < f->sx; pi++) {
            int i = pi + 1 - f->ox;
            F(*f, i, 1 - k) = F(*f, i, Ny + 1 - k);
            F(*f, i, Ny + k) = F(*f, i, k);
        }
}
int csio_fill_halo(G g, P p, csio_field *f, int lx, int ly, int which)
{
    if (!f->p) return 0;
    if (g->topo_x == CSIO_BOUNDED) fill_x_bounded(g, p, f, lx, ly, which);
    if (g->topo_y == CSIO_BOUNDED || g->topo_y == CSIO_FOLDED) fill_y_bounded(g, p, f, lx, ly, which, g->topo_y == CSIO_BOUNDED);
    if (g->topo_x == CSIO_PERIODIC) fill_x_periodic(g, f);
    if (g->topo_y == CSIO_PERIODIC) fill_y_periodic(g, f);
    if (g->topo_y == CSIO_FOLDED) fill_fold(g, f, lx, ly, which, which == 3);
    return 0;
}

static void copy_parent(csio_field dst, csio_field src)
{
    memcpy(dst.p, src.p, sizeof(double) * (size_t)src.sx * (size_t)src.sy);
}

/* ---------------------------------------------------------------------------------------------
 * time_step_momentum!.  ref: split_explicit_momentum_equations.jl:103-195
 * ------------------------------------------------------------------------------------------- */
int csio_time_step_momentum(G g, P p, csio_state *s, double dt, int nsub)
{
    csio_params *pm = (csio_params *)p;
    /* reset_velocities! (se.jl:87-93): RK only */
    if (p->timestepper == CSIO_RK3 && s->um.p) {
        copy_parent(s->u, s->um);
        copy_parent(s->v, s->vm);
    }
    csio_initialize_rheology(g, p, s); /* se.jl:130 */
    /* update_external_stress! (ext.jl:72-78,148-152): halo refresh of the stress inputs */
    if (p->top_kind == CSIO_STRESS_FIELD || p->top_kind == CSIO_STRESS_SEMI_IMPLICIT) {
        if (pm->top_x.p) csio_fill_halo(g, p, &pm->top_x, 1, 0, 3);
        if (pm->top_y.p) csio_fill_halo(g, p, &pm->top_y, 0, 1, 3);
    }
    if (p->bot_kind == CSIO_STRESS_FIELD || p->bot_kind == CSIO_STRESS_SEMI_IMPLICIT) {
        if (pm->ue.p) csio_fill_halo(g, p, &pm->ue, 1, 0, 3);
        if (pm->ve.p) csio_fill_halo(g, p, &pm->ve, 0, 1, 3);
    }
    csio_fill_halo(g, p, &s->u, 1, 0, 1); /* se.jl:170-171 */
    csio_fill_halo(g, p, &s->v, 0, 1, 2);
    for (int sub = 1; sub <= nsub; sub++) { /* se.jl:173-189 */
        csio_compute_stresses(g, p, s, dt);
        if (sub % 2 == 0) {
            csio_u_velocity_step(g, p, s, dt);
            csio_fill_halo(g, p, &s->u, 1, 0, 1);
            csio_v_velocity_step(g, p, s, dt);
            csio_fill_halo(g, p, &s->v, 0, 1, 2);
        } else {
            csio_v_velocity_step(g, p, s, dt);
            csio_fill_halo(g, p, &s->v, 0, 1, 2);
            csio_u_velocity_step(g, p, s, dt);
            csio_fill_halo(g, p, &s->u, 1, 0, 1);
        }
    }
    /* finalize_rheology! (evp.jl:275-280) */
    csio_fill_halo(g, p, &s->s11, 0, 0, 0);
    csio_fill_halo(g, p, &s->s12, 1, 1, 0);
    csio_fill_halo(g, p, &s->s22, 0, 0, 0);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Advection.  ref: src/sea_ice_advection.jl:51-58, src/tracer_tendency_kernel_functions.jl:9-45
 * [OCN-recall] advective_tracer_flux_x = Ax^fcc * U * c~, bias left if U > 0 else right; the flux
 * is zeroed on immersed-peripheral faces.
 * ------------------------------------------------------------------------------------------- */
static inline double flux_x(G g, P p, const csio_state *s, const csio_field *c, int i, int j)
{
    double U = F(s->u, i, j);
    double ct = csio_reconstruct_x(g, p->advection_order, U > 0 ? 0 : 1, c, i, j);
    double fl = axfcc(g, i, j) * U * ct;
    return (g->mask && imm_peripheral_fc(g, i, j)) ? 0.0 : fl;
}
static inline double flux_y(G g, P p, const csio_state *s, const csio_field *c, int i, int j)
{
    double V = F(s->v, i, j);
    double ct = csio_reconstruct_y(g, p->advection_order, V > 0 ? 0 : 1, c, i, j);
    double fl = aycfc(g, i, j) * V * ct;
    return (g->mask && imm_peripheral_cf(g, i, j)) ? 0.0 : fl;
}
static inline double horizontal_div_Uc(G g, P p, const csio_state *s, const csio_field *c, int i, int j)
{
    if (p->advection_order == 0) return 0.0;
    return 1 / vccc(g, i, j) * ((flux_x(g, p, s, c, i + 1, j) - flux_x(g, p, s, c, i, j)) + (flux_y(g, p, s, c, i, j + 1) - flux_y(g, p, s, c, i, j)));
}
int csio_compute_tracer_tendencies(G g, P p, csio_state *s)
{
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            F(s->Gh, i, j) = -horizontal_div_Uc(g, p, s, &s->h, i, j);
            F(s->Ga, i, j) = -horizontal_div_Uc(g, p, s, &s->a, i, j);
            if (s->hs.p) F(s->Ghs, i, j) = -horizontal_div_Uc(g, p, s, &s->hs, i, j); /* tracer_tendency:47-52 */
        }
    return 0;
}

/* _dynamic_step_tracers!.  ref: src/sea_ice_fe_step.jl:56-82; drivers fe:36-50, rk:134-152 */
int csio_dynamic_time_step(G g, P p, csio_state *s, double dt)
{
    csio_field hn = (p->timestepper == CSIO_RK3) ? s->hm : s->h;
    csio_field an = (p->timestepper == CSIO_RK3) ? s->am : s->a;
    csio_field hsn = (p->timestepper == CSIO_RK3) ? s->hsm : s->hs;
#pragma omp parallel for schedule(static)
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            double hp = F(hn, i, j) + dt * F(s->Gh, i, j);
            double ap = F(an, i, j) + dt * F(s->Ga, i, j);
            ap = jl_max(0.0, ap);
            hp = jl_max(0.0, hp);
            ap = (hp == 0) ? 0.0 : ap;
            hp = (ap == 0) ? 0.0 : hp;
            double Vp = hp * ap;
            F(s->a, i, j) = ap > 1 ? 1.0 : ap;
            F(s->h, i, j) = ap > 1 ? Vp : hp;
            if (s->hs.p) { /* dynamic_step_snow!: fe.jl:84-94 (reads the concentration just written) */
                double sp = F(hsn, i, j) + dt * F(s->Ghs, i, j);
                sp = jl_max(0.0, sp);
                sp = (F(s->a, i, j) <= 0) ? 0.0 : sp;
                F(s->hs, i, j) = sp;
            }
        }
    return 0;
}

/* mask_immersed_field_xy! [OCN-recall]: zero a field on peripheral nodes of its own location. */
static void mask_immersed(G g, csio_field *f, int lx, int ly)
{
    if (!g->mask) return;
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            int per = lx ? (ly ? 0 : peripheral_fc(g, i, j)) : (ly ? peripheral_cf(g, i, j) : inactive_cell(g, i, j));
            if (per) F(*f, i, j) = 0.0;
        }
}

/* update_state!.  ref: src/sea_ice_model.jl:379-394 (prognostic order: h, aice, u, v) */
int csio_update_state(G g, P p, csio_state *s)
{
    mask_immersed(g, &s->h, 0, 0);
    csio_fill_halo(g, p, &s->h, 0, 0, 0);
    mask_immersed(g, &s->a, 0, 0);
    csio_fill_halo(g, p, &s->a, 0, 0, 0);
    if (s->hs.p) {
        mask_immersed(g, &s->hs, 0, 0);
        csio_fill_halo(g, p, &s->hs, 0, 0, 0);
    }
    mask_immersed(g, &s->u, 1, 0);
    csio_fill_halo(g, p, &s->u, 1, 0, 1);
    mask_immersed(g, &s->v, 0, 1);
    csio_fill_halo(g, p, &s->v, 0, 1, 2);
    return 0;
}

/* time_step!.  FE: src/sea_ice_fe_step.jl:13-34.  RK3: src/sea_ice_rk_substep.jl:29-94 driven by
 * Oceananigans' SplitRungeKuttaTimeStepper [OCN-recall: beta = (3, 2, 1), dtau = dt/beta,
 * update_state! after every stage]. */
int csio_time_step(G g, P p, csio_state *s, double dt, int first)
{
    if (first) csio_update_state(g, p, s);
    if (p->timestepper == CSIO_FE) {
        csio_compute_tracer_tendencies(g, p, s);
        csio_time_step_momentum(g, p, s, dt, p->substeps);
        csio_dynamic_time_step(g, p, s, dt);
        csio_update_state(g, p, s);
        return 0;
    }
    /* cache_current_fields! (rk.jl:29-42) */
    copy_parent(s->hm, s->h);
    copy_parent(s->am, s->a);
    if (s->hs.p) copy_parent(s->hsm, s->hs);
    copy_parent(s->um, s->u);
    copy_parent(s->vm, s->v);
    for (int beta = 3; beta >= 1; beta--) {
        double dtau = dt / beta;
        csio_compute_tracer_tendencies(g, p, s);              /* rk.jl:84 */
        csio_time_step_momentum(g, p, s, dtau, p->substeps);  /* rk.jl:87 */
        csio_dynamic_time_step(g, p, s, dtau);                /* rk.jl:89 */
        csio_update_state(g, p, s);
    }
    return 0;
}

/* cell_advection_timescale.  ref: src/ClimaSeaIce.jl:66-69 [OCN-recall]: min 1/(|u|/dx + |v|/dy) */
double csio_cell_advection_timescale(G g, const csio_state *s)
{
    double tmin = INFINITY;
    for (int j = 1; j <= g->Ny; j++)
        for (int i = 1; i <= g->Nx; i++) {
            double t = 1 / (fabs(F(s->u, i, j)) / dxfc(g, i, j) + fabs(F(s->v, i, j)) / dycf(g, i, j));
            if (t < tmin) tmin = t;
        }
    return tmin;
}

/* stress_power_budget.  ref: test/test_rheology_energy_budget.jl:25-35,50-91 */
int csio_stress_power_budget(G g, const csio_state *s, double *out)
{
    double Wn = 0, Wo = 0, D = 0;
    csio_field u = s->u, v = s->v;
    for (int i = 1; i <= g->Nx; i++)
        for (int j = 1; j <= g->Ny; j++) {
            Wn += F(u, i, j) * div_sigma_1j(g, s, i, j) * azfc(g, i, j);
            Wn += F(v, i, j) * div_sigma_2j(g, s, i, j) * azcf(g, i, j);
            double o1 = ((dycc(g, i, j) * F(s->s11, i, j) - dycc(g, i - 1, j) * F(s->s11, i - 1, j)) +
                         (dxff(g, i, j + 1) * F(s->s12, i, j + 1) - dxff(g, i, j) * F(s->s12, i, j))) / azfc(g, i, j);
            double o2 = ((dyff(g, i + 1, j) * F(s->s12, i + 1, j) - dyff(g, i, j) * F(s->s12, i, j)) +
                         (dxcc(g, i, j) * F(s->s22, i, j) - dxcc(g, i, j - 1) * F(s->s22, i, j - 1))) / azcf(g, i, j);
            Wo += F(u, i, j) * o1 * azfc(g, i, j);
            Wo += F(v, i, j) * o2 * azcf(g, i, j);
            D += F(s->s11, i, j) * strain_rate_xx(g, u, v, i, j) * azcc(g, i, j);
            D += F(s->s22, i, j) * strain_rate_yy(g, u, v, i, j) * azcc(g, i, j);
            D += 2 * F(s->s12, i, j) * strain_rate_xy(g, u, v, i, j) * azff(g, i, j);
        }
    out[0] = Wn;
    out[1] = Wo;
    out[2] = D;
    return 0;
}
