// csi_math.cuh -- scalar building blocks shared by every kernel of libclimaseaice_b200.
//
// The library must reproduce the reference's Float64 operation sequence exactly (SURVEY.md
// section 7: the EVP substep loop amplifies a 1-ulp difference to 1e-7 after 150 substeps), so:
//   * everything is compiled with -fmad=false; FMA appears only where written explicitly,
//     and only in constructs that return the correctly rounded IEEE result;
//   * Julia's max / clamp / Bool-multiply semantics are restated here;
//   * ice_strength's exp is a correctly rounded exp (double-double), so it agrees with any other
//     correctly rounded implementation.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define CSI_HD __host__ __device__ __forceinline__
#else
#define CSI_HD inline
#endif

namespace csi {

CSI_HD double fma_rn(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// ---- Julia scalar semantics ------------------------------------------------------------------
// Base.max: NaN-propagating, max(-0.0, +0.0) = +0.0
CSI_HD double jl_max(double a, double b)
{
    if (a != a) return a;
    if (b != b) return b;
    if (a == b) return signbit(a) ? b : a;
    return a > b ? a : b;
}
// clamp(x, lo, hi) = ifelse(x > hi, hi, ifelse(x < lo, lo, x))
CSI_HD double jl_clamp(double x, double lo, double hi) { return x > hi ? hi : (x < lo ? lo : x); }
// x * b::Bool = ifelse(b, x, copysign(0, x))
CSI_HD double jl_mul_bool(double x, bool b) { return b ? x : copysign(0.0, x); }

// ---- correctly rounded division by a reused divisor --------------------------------------------
// With r = RN(1/d):  q0 = RN(x*r);  e = x - d*q0 (exact, one FMA);  q = RN(q0 + e*r).
// Markstein's theorem: q = RN(x/d) whenever q0 is a faithful rounding of x/d and no intermediate
// over/underflows.  `safe_divisor` screens the divisors for which the theorem's premise can fail
// (significand all ones) and the exponent range; callers fall back to `/` otherwise.
struct Recip {
    double d, r;
    bool fast;
};
CSI_HD bool recip_is_safe(double d)
{
    union { double f; uint64_t u; } c;
    c.f = d;
    const uint64_t mant = c.u & 0x000fffffffffffffull;
    const int ex = (int)((c.u >> 52) & 0x7ff);
    // normal, comfortably inside the exponent range, significand not all ones
    return ex > 200 && ex < 1800 && mant != 0x000fffffffffffffull;
}
CSI_HD Recip make_recip(double d)
{
    Recip R;
    R.d = d;
    R.r = 1.0 / d;
    R.fast = recip_is_safe(d);
    return R;
}
CSI_HD double div_markstein(double x, double d, double r)
{
    const double q0 = x * r;
    const double e = fma_rn(-d, q0, x);
    return fma_rn(e, r, q0);
}
// x / R.d, bit-identical to the IEEE quotient.  The quick path requires |x/d| to stay in the
// normal range with margin (checked on q0); everything else takes the plain division.
CSI_HD double div_by(double x, const Recip &R)
{
    const double q0 = x * R.r;
    const double a = fabs(q0);
    if (R.fast && a > 1e-280 && a < 1e280) {
        const double e = fma_rn(-R.d, q0, x);
        return fma_rn(e, R.r, q0);
    }
    return x / R.d;
}

// ---- correctly rounded exp ---------------------------------------------------------------------
struct dd {
    double hi, lo;
};
CSI_HD dd quick_two_sum(double a, double b)
{
    dd r;
    r.hi = a + b;
    r.lo = b - (r.hi - a);
    return r;
}
CSI_HD dd two_sum(double a, double b)
{
    dd r;
    r.hi = a + b;
    const double bb = r.hi - a;
    r.lo = (a - (r.hi - bb)) + (b - bb);
    return r;
}
CSI_HD dd two_prod(double a, double b)
{
    dd r;
    r.hi = a * b;
    r.lo = fma_rn(a, b, -r.hi);
    return r;
}
CSI_HD dd dd_add(dd a, dd b)
{
    dd s = two_sum(a.hi, b.hi);
    dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
CSI_HD dd dd_mul(dd a, dd b)
{
    dd p = two_prod(a.hi, b.hi);
    p.lo += a.hi * b.lo + a.lo * b.hi;
    return quick_two_sum(p.hi, p.lo);
}

// 1/n! as double-double, n = 0..27
#define CSI_INVFACT_INIT { \
    {0x1.0000000000000p+0, 0x0.0p+0}, \
    {0x1.0000000000000p+0, 0x0.0p+0}, \
    {0x1.0000000000000p-1, 0x0.0p+0}, \
    {0x1.5555555555555p-3, 0x1.5555555555555p-57}, \
    {0x1.5555555555555p-5, 0x1.5555555555555p-59}, \
    {0x1.1111111111111p-7, 0x1.1111111111111p-63}, \
    {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65}, \
    {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73}, \
    {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, \
    {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, \
    {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76}, \
    {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80}, \
    {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, \
    {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87}, \
    {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92}, \
    {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97}, \
    {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101}, \
    {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103}, \
    {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107}, \
    {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112}, \
    {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120}, \
    {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120}, \
    {0x1.0ce396db7f853p-70, -0x1.aebcdbd20331cp-124}, \
    {0x1.761b41316381ap-75, -0x1.3423c7d91404fp-130}, \
    {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135}, \
    {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139}, \
    {0x1.88e85fc6a4e5ap-89, -0x1.71c37ebd16540p-143}, \
    {0x1.d1ab1c2dccea3p-94, 0x1.054d0c78aea14p-149}, \
}
namespace tables {
static const double INVFACT_H[28][2] = CSI_INVFACT_INIT;
#if defined(__CUDACC__)
static __constant__ double INVFACT_D[28][2] = CSI_INVFACT_INIT;
#endif
}  // namespace tables
CSI_HD double invfact(int n, int part)
{
#if defined(__CUDA_ARCH__)
    return tables::INVFACT_D[n][part];
#else
    return tables::INVFACT_H[n][part];
#endif
}

// exp(x) rounded to nearest.  Cody-Waite reduction x = k ln2 + r with ln2 split in three parts
// (k*ln2_hi exact), Taylor series of exp(r), |r| <= 0.347, in double-double (about 2^-100
// relative), final rounding by the double-double normalisation, exact scaling by 2^k.
CSI_HD double exp_cr(double x)
{
    if (x != x) return x;
    if (x > 709.782712893384) return INFINITY;
    if (x < -745.1332191019412) return 0.0;
    const double LN2_HI = 0x1.62e42fee00000p-1, LN2_MID = 0x1.a39ef35793c76p-33, LN2_LO = 0x1.cc01f97b57a08p-87;
    const double INV_LN2 = 0x1.71547652b82fep+0;
    const double kf = rint(x * INV_LN2);
    const int k = (int)kf;
    dd r;
    r.hi = x - kf * LN2_HI;  // exact
    r.lo = 0.0;
    dd m = two_prod(kf, LN2_MID);
    m.hi = -m.hi;
    m.lo = -m.lo;
    r = dd_add(r, m);
    dd l;
    l.hi = -(kf * LN2_LO);
    l.lo = 0.0;
    r = dd_add(r, l);
    dd p;
    p.hi = invfact(27, 0);
    p.lo = invfact(27, 1);
#pragma unroll 1
    for (int n = 26; n >= 0; --n) {
        dd c;
        c.hi = invfact(n, 0);
        c.lo = invfact(n, 1);
        p = dd_add(dd_mul(p, r), c);
    }
    // p.hi = RN(p.hi + p.lo): the correctly rounded exp(r) in [0.70, 1.42]
    if (k > -1021 && k < 1023) {
        union { double f; uint64_t u; } s;
        s.u = (uint64_t)(k + 1023) << 52;
        return p.hi * s.f;
    }
    return ldexp(p.hi, k);  // results in / next to the subnormal range: one extra rounding possible
}

}  // namespace csi
