/*
 * csi_oracle_weno.c -- ORACLE (test infrastructure only; see csi_oracle.h).
 *
 * Upwind-biased WENO-Z face reconstruction used by Oceananigans' `_advective_tracer_flux_x/y`
 * (call sites: /root/reference/src/sea_ice_advection.jl:51-58).  Oceananigans' source is not in
 * the image: the tables below are the standard Jiang-Shu / Balsara-Shu ones restated from
 * SURVEY.md Appendix A ("[OCN-recall]", confidence M/L) and are kept in this one file so they can
 * be swapped without touching anything else.  PARITY UNPINNED against the Julia original.
 *
 * Conventions (Appendix A): face i lies between cells i-1 and i.  Left bias (U > 0): candidate
 * stencil k holds cells (i-1-k .. i-1-k+B-1) in increasing order.  Right bias (U <= 0): mirror
 * about the face, stencil k holds cells (i+k .. i+k-B+1) in decreasing order, same coefficients.
 * Weights are WENO-Z: alpha_k = C_k * (1 + (tau/(beta_k+eps))^2), eps = 1e-8.
 * Every sum is evaluated left to right, as Julia's n-ary + does.
 */
#include "csi_oracle.h"
#include <math.h>
#include <stddef.h>

#define F(f, i, j) ((f).p[(size_t)((i)-1 + (f).ox) + (size_t)((j)-1 + (f).oy) * (size_t)(f).sx])

static const double WENO_EPS = 1e-8;

/* reconstruction coefficients, [stencil][cell] */
static const double R2[2][2] = {{1.0 / 2.0, 1.0 / 2.0}, {-1.0 / 2.0, 3.0 / 2.0}};
static const double R3[3][3] = {{1.0 / 3.0, 5.0 / 6.0, -1.0 / 6.0},
                                {-1.0 / 6.0, 5.0 / 6.0, 1.0 / 3.0},
                                {1.0 / 3.0, -7.0 / 6.0, 11.0 / 6.0}};
static const double R4[4][4] = {{1.0 / 4.0, 13.0 / 12.0, -5.0 / 12.0, 1.0 / 12.0},
                                {-1.0 / 12.0, 7.0 / 12.0, 7.0 / 12.0, -1.0 / 12.0},
                                {1.0 / 12.0, -5.0 / 12.0, 13.0 / 12.0, 1.0 / 4.0},
                                {-1.0 / 4.0, 13.0 / 12.0, -23.0 / 12.0, 25.0 / 12.0}};
/* optimal weights */
static const double C2[2] = {2.0 / 3.0, 1.0 / 3.0};
static const double C3[3] = {3.0 / 10.0, 3.0 / 5.0, 1.0 / 10.0};
static const double C4[4] = {4.0 / 35.0, 18.0 / 35.0, 12.0 / 35.0, 1.0 / 35.0};
/* smoothness coefficients */
static const double S2[2][3] = {{1, -2, 1}, {1, -2, 1}};
static const double S3[3][6] = {{10, -31, 11, 25, -19, 4}, {4, -13, 5, 13, -13, 4}, {4, -19, 11, 25, -31, 10}};
static const double S4[4][10] = {
    {2.107, -9.402, 7.042, -1.854, 11.003, -17.246, 4.642, 7.043, -3.882, 0.547},
    {0.547, -2.522, 1.922, -0.494, 3.443, -5.966, 1.602, 2.843, -1.642, 0.267},
    {0.267, -1.642, 1.602, -0.494, 2.843, -5.966, 1.922, 3.443, -2.522, 0.547},
    {0.547, -3.882, 4.642, -1.854, 7.043, -17.246, 7.042, 11.003, -9.402, 2.107}};

static double zweno(double beta, double tau, double C)
{
    double r = tau / (beta + WENO_EPS);
    return C * (1 + r * r);
}

/* psi holds the 2B-1 cells of the biased stencil such that stencil k, slot m is q[(B-1-k)+m]:
 * left bias q[n] = c[i-B+n], right bias q[n] = c[i+B-1-n]  (n = 0..2B-2). */
static double weno_from_cells(int B, const double *q)
{
    if (B == 1) return q[0];
    double beta[4], p[4], al[4], tau, sum, out;
    for (int k = 0; k < B; k++) {
        const double *s = q + (B - 1 - k);
        if (B == 2) {
            beta[k] = s[0] * (S2[k][0] * s[0] + S2[k][1] * s[1]) + S2[k][2] * (s[1] * s[1]);
            p[k] = R2[k][0] * s[0] + R2[k][1] * s[1];
        } else if (B == 3) {
            const double *c = S3[k];
            beta[k] = s[0] * (c[0] * s[0] + c[1] * s[1] + c[2] * s[2]) + s[1] * (c[3] * s[1] + c[4] * s[2]) +
                      c[5] * (s[2] * s[2]);
            p[k] = R3[k][0] * s[0] + R3[k][1] * s[1] + R3[k][2] * s[2];
        } else {
            const double *c = S4[k];
            beta[k] = s[0] * (c[0] * s[0] + c[1] * s[1] + c[2] * s[2] + c[3] * s[3]) +
                      s[1] * (c[4] * s[1] + c[5] * s[2] + c[6] * s[3]) + s[2] * (c[7] * s[2] + c[8] * s[3]) +
                      c[9] * (s[3] * s[3]);
            p[k] = R4[k][0] * s[0] + R4[k][1] * s[1] + R4[k][2] * s[2] + R4[k][3] * s[3];
        }
    }
    if (B == 2) {
        tau = fabs(beta[0] - beta[1]);
        for (int k = 0; k < 2; k++) al[k] = zweno(beta[k], tau, C2[k]);
        sum = al[0] + al[1];
        out = (al[0] / sum) * p[0] + (al[1] / sum) * p[1];
    } else if (B == 3) {
        tau = fabs(beta[0] - beta[2]);
        for (int k = 0; k < 3; k++) al[k] = zweno(beta[k], tau, C3[k]);
        sum = al[0] + al[1] + al[2];
        out = (al[0] / sum) * p[0] + (al[1] / sum) * p[1] + (al[2] / sum) * p[2];
    } else {
        tau = fabs(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]);
        for (int k = 0; k < 4; k++) al[k] = zweno(beta[k], tau, C4[k]);
        sum = al[0] + al[1] + al[2] + al[3];
        out = (al[0] / sum) * p[0] + (al[1] / sum) * p[1] + (al[2] / sum) * p[2] + (al[3] / sum) * p[3];
    }
    return out;
}

/* Order reduction ([OCN-recall], confidence M).  Bounded walls: the buffer at face i is min(B, i-1, N+1-i), at least 1,
 * so no stencil leaves the interior cells (Oceananigans `outside_biased_halo`: the full-order stencil is used where
 * i >= B+1 and i <= N+1-B, else the scheme's `buffer_scheme`, recursively: 7 -> 5 -> 3 -> first-order upwind).
 * ImmersedBoundaryGrid (g->mask): the same chain, driven by `near_x_immersed_boundary_biased` -- a scheme of buffer b is
 * used at face i only if none of the 2b cells i-b .. i+b-1 (the union of the left- and the right-biased stencil) is
 * inactive (immersed, or outside a Bounded domain), else its buffer_scheme is tried; the first-order scheme (b = 1) is
 * never tested.  Call site: /root/reference/src/sea_ice_advection.jl:51-58 (`_advective_tracer_flux_x/y` on the grid of
 * /root/reference/examples/ice_advected_on_coastline.jl:54-55). */
static int cell_inactive(const csio_grid *g, int i, int j)
{
    if ((g->topo_x == CSIO_BOUNDED && (i < 1 || i > g->Nx)) || (g->topo_y == CSIO_BOUNDED && (j < 1 || j > g->Ny)) || (g->topo_y == CSIO_FOLDED && j < 1)) return 1;
    if (!g->mask) return 0;
    int sx = g->Nx + 2 * g->Hx, sy = g->Ny + 2 * g->Hy;
    int pi = i - 1 + g->Hx, pj = j - 1 + g->Hy;
    if (pi < 0) pi = 0;
    if (pj < 0) pj = 0;
    if (pi >= sx) pi = sx - 1;
    if (pj >= sy) pj = sy - 1;
    return g->mask[(size_t)pi + (size_t)pj * (size_t)sx] != 0;
}
static int buffer_at(int B, int topo, int N, int i)
{
    if (topo != CSIO_BOUNDED && topo != CSIO_FOLDED) return B;
    int b = B;
    if (i - 1 < b) b = i - 1;
    if (topo == CSIO_BOUNDED && N + 1 - i < b) b = N + 1 - i;   /* (a fold is not a wall: full order next to it) */
    return b < 1 ? 1 : b;
}
/* dir = 0: x face (i, j), stencil along i; dir = 1: y face, stencil along j */
static int buffer_immersed(const csio_grid *g, int B, int dir, int i, int j)
{
    for (int b = B; b >= 2; b--) {
        int bad = 0;
        for (int c = -b; c <= b - 1 && !bad; c++) bad = dir == 0 ? cell_inactive(g, i + c, j) : cell_inactive(g, i, j + c);
        if (!bad) return b;
    }
    return 1;
}

static int order_to_buffer(int order) { return order <= 1 ? 1 : (order + 1) / 2; }

double csio_reconstruct_x(const csio_grid *g, int order, int bias, const csio_field *c, int i, int j)
{
    int B = buffer_at(order_to_buffer(order), g->topo_x, g->Nx, i);
    if (g->mask) B = buffer_immersed(g, B, 0, i, j);
    double q[7];
    for (int n = 0; n < 2 * B - 1; n++) q[n] = bias == 0 ? F(*c, i - B + n, j) : F(*c, i + B - 1 - n, j);
    return weno_from_cells(B, q);
}

double csio_reconstruct_y(const csio_grid *g, int order, int bias, const csio_field *c, int i, int j)
{
    int B = buffer_at(order_to_buffer(order), g->topo_y, g->Ny, j);
    if (g->mask) B = buffer_immersed(g, B, 1, i, j);
    double q[7];
    for (int n = 0; n < 2 * B - 1; n++) q[n] = bias == 0 ? F(*c, i, j - B + n) : F(*c, i, j + B - 1 - n);
    return weno_from_cells(B, q);
}
