// dc_point.h -- point functions of the Jacobson-2005 Ch.7 sigma-coordinate scheme.
// Each keeps the evaluation order of the reference expression it replaces (cited), so a
// build without FMA contraction (-fmad=false) is bit-identical to the reference's CPU
// path for everything except pow/log (device libm vs glibc).
#pragma once
#include <math.h>
#include <string.h>

#include "dc_geom.h"

namespace dc {

// ---------------------------------------------------------------------------------------
// Arithmetic mode.  Default (strict): every division of the reference is an IEEE division,
// no FMA contraction (-fmad=false): bit-identical to the reference's CPU path up to pow/log.
// DC_FAST_MATH: a division by a per-level / per-row / per-column quantity becomes a
// multiplication by its reciprocal, computed once (<= 1 ulp per operation; a correctly
// rounded fp64 division costs ~20 instructions incl. its slow-path check), and the build
// allows FMA contraction.  Results stay within the stated parity tolerances
// (tests/helpers.py: TOL) but are no longer bit-identical to the oracle.
// ---------------------------------------------------------------------------------------
#ifdef DC_FAST_MATH
#define DC_FAST 1
#else
#define DC_FAST 0
#endif
struct Div {
    double d;  // the divisor
    double r;  // 1 / d (only meaningful when DC_FAST)
};
DC_HD Div mkdiv(double d) { return Div{d, DC_FAST ? 1. / d : 0.}; }
DC_HD Div mkdiv(double d, double r) { return Div{d, r}; }
DC_HD double operator/(double x, const Div &v) { return DC_FAST ? x * v.r : x / v.d; }

// dyn_continuity.py:40-47
DC_HD double calc_UFLX(double UWIND, double COLP, double COLP_im1, double dyis)
{
    return (COLP_im1 + COLP) / 2. * UWIND * dyis;
}
DC_HD double calc_VFLX(double VWIND, double COLP, double COLP_jm1, double dxjs)
{
    return (COLP_jm1 + COLP) / 2. * VWIND * dxjs;
}
DC_HD double calc_FLXDIV(double UFLX, double UFLX_ip1, double VFLX, double VFLX_jp1,
                         double dsigma, Div A)
{
    return (+UFLX_ip1 - UFLX + VFLX_jp1 - VFLX) * dsigma / A;
}

// dyn_functions.py:211-270 (interior interfaces 0 < k < nz).  P_* = COLP_NEW*A*WWIND at
// the six columns around the velocity point: d = along the wind, p = perpendicular.
// wall = -1: rigid wall on the pm1 side (p_ind == nb); +1: on the pp1 side (p_ind == np)
DC_HD double colpa_wwind(double P, double P_dm1, double P_pm1, double P_pp1, double P_pm1_dm1,
                         double P_pp1_dm1, int wall)
{
    // the reference forms the centre-row terms as (2*COLP_NEW)*A*WWIND; scaling by 2 is
    // exact in binary floating point, so 2*(COLP_NEW*A*WWIND) is the same number
    if (wall < 0) return 0.25 * (P_pp1_dm1 + P_pp1 + P_dm1 + P);
    if (wall > 0) return 0.25 * (P_dm1 + P + P_pm1_dm1 + P_pm1);
    return 0.125 * (P_pp1_dm1 + P_pp1 + 2. * P_dm1 + 2. * P + P_pm1_dm1 + P_pm1);
}
// dss = dsigma + dsigma_km1
DC_HD double interp_ks(double DWIND, double DWIND_km1, double dsigma, double dsigma_km1, Div dss)
{
    return ((dsigma * DWIND_km1 + dsigma_km1 * DWIND) / dss);
}

// dyn_functions.py:429-536 (auxiliary momentum fluxes); u = UFLX, v = VFLX
DC_HD double calc_CFLX(double v_im1_jm1, double v_jm1, double v_im1, double v, double v_im1_jp1,
                       double v_jp1)
{
    return 1. / 12. * (v_im1_jm1 + v_jm1 + 2. * (v_im1 + v) + v_im1_jp1 + v_jp1);
}
DC_HD double calc_QFLX(double u_im1_jm1, double u_im1, double u_jm1, double u, double u_ip1_jm1,
                       double u_ip1)
{
    return 1. / 12. * (u_im1_jm1 + u_im1 + 2. * (u_jm1 + u) + u_ip1_jm1 + u_ip1);
}
DC_HD double calc_DFLX(double v_jm1, double v, double v_jp1, double u_jm1, double u,
                       double u_ip1_jm1, double u_ip1)
{
    return 1. / 24. * (v_jm1 + 2. * v + v_jp1 + u_jm1 + u + u_ip1_jm1 + u_ip1);
}
DC_HD double calc_EFLX(double v_jm1, double v, double v_jp1, double u_jm1, double u,
                       double u_ip1_jm1, double u_ip1)
{
    return 1. / 24. * (v_jm1 + 2. * v + v_jp1 - u_jm1 - u - u_ip1_jm1 - u_ip1);
}
DC_HD double calc_SFLX(double v_im1, double v_im1_jp1, double v, double v_jp1, double u_im1,
                       double u, double u_ip1)
{
    return 1. / 24. * (v_im1 + v_im1_jp1 + v + v_jp1 + u_im1 + 2. * u + u_ip1);
}
DC_HD double calc_TFLX(double v_im1, double v_im1_jp1, double v, double v_jp1, double u_im1,
                       double u, double u_ip1)
{
    return 1. / 24. * (v_im1 + v_im1_jp1 + v + v_jp1 - u_im1 - 2. * u - u_ip1);
}
DC_HD double calc_BFLX(double u_jm1, double u_ip1_jm1, double u, double u_ip1, double u_jp1,
                       double u_ip1_jp1)
{
    return 1. / 12. * (u_jm1 + u_ip1_jm1 + 2. * (u + u_ip1) + u_jp1 + u_ip1_jp1);
}
DC_HD double calc_RFLX(double v_im1, double v_im1_jp1, double v, double v_jp1, double v_ip1,
                       double v_ip1_jp1)
{
    return 1. / 12. * (v_im1 + v_im1_jp1 + 2. * (v + v_jp1) + v_ip1 + v_ip1_jp1);
}

// dyn_functions.py:541-568
DC_HD double UVFLX_hor_adv(double DWIND, double DWIND_dm1, double DWIND_dp1, double DWIND_pm1,
                           double DWIND_pp1, double DWIND_dm1_pm1, double DWIND_dm1_pp1,
                           double DWIND_dp1_pm1, double DWIND_dp1_pp1, double BRFLX,
                           double BRFLX_dm1, double CQFLX, double CQFLX_pp1, double DSFLX_dm1,
                           double DSFLX_pp1, double ETFLX, double ETFLX_dm1_pp1,
                           double sign_ETFLX_term)
{
    return (+BRFLX_dm1 * (DWIND_dm1 + DWIND) / 2. - BRFLX * (DWIND + DWIND_dp1) / 2.
            + CQFLX * (DWIND_pm1 + DWIND) / 2. - CQFLX_pp1 * (DWIND + DWIND_pp1) / 2.
            + DSFLX_dm1 * (DWIND_dm1_pm1 + DWIND) / 2. - DSFLX_pp1 * (DWIND + DWIND_dp1_pp1) / 2.
            + sign_ETFLX_term * (+ETFLX * (DWIND_dp1_pm1 + DWIND) / 2. -
                                 ETFLX_dm1_pp1 * (DWIND + DWIND_dm1_pp1) / 2.));
}

// dyn_functions.py:177-207
DC_HD double pre_grad(double PHI, double PHI_dm1, double COLP, double COLP_dm1, double POTT,
                      double POTT_dm1, double PVTF, double PVTF_dm1, double PVTFVB,
                      double PVTFVB_dm1, double PVTFVB_dm1_kp1, double PVTFVB_kp1, Div dsigma,
                      double sigma_vb, double sigma_vb_kp1, double dgrid)
{
    return (-dgrid *
            ((PHI - PHI_dm1) * (COLP + COLP_dm1) / 2. +
             (COLP - COLP_dm1) * con_cp / 2. *
                 (+POTT_dm1 / dsigma *
                      (sigma_vb_kp1 * (PVTFVB_dm1_kp1 - PVTF_dm1) +
                       sigma_vb * (PVTF_dm1 - PVTFVB_dm1)) +
                  POTT / dsigma *
                      (sigma_vb_kp1 * (PVTFVB_kp1 - PVTF) + sigma_vb * (PVTF - PVTFVB)))));
}

// dyn_functions.py:158-170
DC_HD double num_dif(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1, double VAR_jp1,
                     double VAR_dif_coef)
{
    return VAR_dif_coef * (+VAR_im1 + VAR_ip1 + VAR_jm1 + VAR_jp1 - 4. * VAR);
}

// Coriolis + spherical-metric terms.  The latitude-only factors of the reference expressions
// are formed once per row, with the reference's own operation order, by these two helpers:
//   cor_fcos  = corf * con_rE * cos(lat)          (dyn_UFLX.py:56-57, dyn_VFLX.py:53-54)
//   cor_scale = con_rE * dlon_rad * dlat_rad / 2  (dyn_UFLX.py:52, dyn_VFLX.py:50)
DC_HD double cor_fcos(double corf, double cos_lat) { return corf * con_rE * cos_lat; }
DC_HD double cor_scale(double dlon_rad, double dlat_rad) { return con_rE * dlon_rad * dlat_rad / 2.; }

// dyn_UFLX.py:39-66
DC_HD double coriolis_UWIND(double COLP, double COLP_im1, double VWIND, double VWIND_im1,
                            double VWIND_jp1, double VWIND_im1_jp1, double UWIND,
                            double UWIND_im1, double UWIND_ip1, double fcos_is, double sin_lat_is,
                            double scale)
{
    return (scale *
            (COLP_im1 * (VWIND_im1 + VWIND_im1_jp1) / 2. *
                 (fcos_is + (UWIND_im1 + UWIND) / 2. * sin_lat_is) +
             COLP * (VWIND + VWIND_jp1) / 2. * (fcos_is + (UWIND + UWIND_ip1) / 2. * sin_lat_is)));
}

// dyn_VFLX.py:39-64
DC_HD double coriolis_VWIND(double COLP, double COLP_jm1, double UWIND, double UWIND_jm1,
                            double UWIND_ip1, double UWIND_ip1_jm1, double fcos, double sin_lat,
                            double fcos_jm1, double sin_lat_jm1, double scale)
{
    return (-scale *
            (COLP_jm1 * (UWIND_jm1 + UWIND_ip1_jm1) / 2. *
                 (fcos_jm1 + (UWIND_jm1 + UWIND_ip1_jm1) / 2. * sin_lat_jm1) +
             COLP * (UWIND + UWIND_ip1) / 2. * (fcos + (UWIND + UWIND_ip1) / 2. * sin_lat)));
}

// dyn_functions.py:105-114
DC_HD double hor_adv(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1, double VAR_jp1,
                     double UFLX, double UFLX_ip1, double VFLX, double VFLX_jp1, Div A)
{
    return ((+UFLX * (VAR_im1 + VAR) / 2. - UFLX_ip1 * (VAR + VAR_ip1) / 2.
             + VFLX * (VAR_jm1 + VAR) / 2. - VFLX_jp1 * (VAR + VAR_jp1) / 2.) / A);
}

// dyn_functions.py:118-136 (k == nz branch unreachable)
DC_HD double vert_adv(double VARVB, double VARVB_kp1, double WWIND, double WWIND_kp1,
                      double COLP_NEW, Div dsigma, int k)
{
    if (k == 0) return COLP_NEW * (-WWIND_kp1 * VARVB_kp1) / dsigma;
    return COLP_NEW * (+WWIND * VARVB - WWIND_kp1 * VARVB_kp1) / dsigma;
}

// dyn_functions.py:142-155
DC_HD double num_dif_pw(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1,
                        double VAR_jp1, double COLP, double COLP_im1, double COLP_ip1,
                        double COLP_jm1, double COLP_jp1, double VAR_dif_coef)
{
    return VAR_dif_coef * (+COLP_im1 * VAR_im1 + COLP_ip1 * VAR_ip1 + COLP_jm1 * VAR_jm1 +
                           COLP_jp1 * VAR_jp1 - 4. * COLP * VAR);
}

// dyn_functions.py:70-95
DC_HD double comp_VARVB_log(double VAR, double VAR_km1)
{
    const double min_val = 0.0000001;
    VAR = fmax(VAR, min_val);
    VAR_km1 = fmax(VAR_km1, min_val);
    if (VAR_km1 == VAR) return VAR;
    return ((log(VAR_km1) - log(VAR)) / (1. / VAR - 1. / VAR_km1));
}

// ---------------------------------------------------------------------------------------
// x^kappa for the Exner function of the PRODUCTION build: exp(kappa * log(x)) with the
// classic fdlibm log kernel (atanh series, 7 coefficients, < 1 ulp) and a degree-13 Taylor
// exp after Cody-Waite reduction; valid for finite x > 0 in the normal range (pressures),
// measured <= 3 ulp against glibc pow over [5e-4, 1.3].  The coefficients travel in the kernel
// parameter block, so that they are constant-bank operands of the DFMAs: the CUDA libm
// versions re-materialise each 64-bit constant with two moves per use, which made the
// diagnostics sweep instruction-issue bound (ncu: 78 % issue slots, 269 instructions per cell).
// ---------------------------------------------------------------------------------------
// Table-driven x^kappa (pow_kappa_tab below): x = 2^e * m, m in [1, 2); the top POW_JBITS
// mantissa bits pick the interval centre m_j; with r_j = fl(1 / m_j) and t = fma(m, r_j, -1)
// (|t| <= 2^-(POW_JBITS+1), one rounding),   x^kappa = (2^e / r_j)^kappa * (1 + t)^kappa
// = T[e][j] * (1 + t*(b1 + t*(b2 + ... b8 t^7))),  b_n = binomial(kappa, n).  T is tabulated
// from the ROUNDED r_j in long double, so the identity is exact up to the polynomial
// (truncation < 1e-19) and ~1 ulp of rounding: 18 instructions, 10 of them FP64, against ~55 / 34
// of exp(kappa*log(x)) -- the diagnostics sweep was FP64- and issue-bound on that chain.
constexpr int POW_JBITS = 6, POW_NJ = 1 << POW_JBITS;
constexpr int POW_EMIN = -5, POW_NE = 7;   // 2^-5 <= x < 2^2: pressures 3.1 kPa .. 400 kPa
struct PowCoef {
    double Lg1, Lg2, Lg3, Lg4, Lg5, Lg6, Lg7, ln2_hi, ln2_lo, inv_ln2, big, kappa;
    double e[14];   // 1 / n!
    double b[9];    // binomial(kappa, n), n = 0..8
    const double *tab;   // [POW_NE][POW_NJ][2] = {r_j, T[e][j]} (device memory), or NULL
};
// host: fill the table (2 * POW_NE * POW_NJ doubles)
inline void make_pow_table(double kappa, double *tab)
{
    for (int e = 0; e < POW_NE; e++)
        for (int j = 0; j < POW_NJ; j++) {
            const double mj = 1. + (j + 0.5) / POW_NJ;
            const double r = 1. / mj;
            const long double base = ldexpl(1.0L, e + POW_EMIN) / (long double)r;
            tab[2 * (e * POW_NJ + j)] = r;
            tab[2 * (e * POW_NJ + j) + 1] = (double)powl(base, (long double)kappa);
        }
}
inline PowCoef make_pow_coef(double kappa, const double *tab = nullptr)
{
    PowCoef c;
    c.tab = tab;
    {
        long double bn = 1.0L;
        c.b[0] = 1.;
        for (int n = 1; n < 9; n++) {
            bn = bn * ((long double)kappa - (n - 1)) / n;
            c.b[n] = (double)bn;
        }
    }
    c.Lg1 = 6.666666666666735130e-01; c.Lg2 = 3.999999999940941908e-01;
    c.Lg3 = 2.857142874366239149e-01; c.Lg4 = 2.222219843214978396e-01;
    c.Lg5 = 1.818357216161805012e-01; c.Lg6 = 1.531383769920937332e-01;
    c.Lg7 = 1.479819860511658591e-01;
    c.ln2_hi = 6.93147180369123816490e-01; c.ln2_lo = 1.90821492927058770002e-10;
    c.inv_ln2 = 1.44269504088896338700e+00;
    c.big = 6755399441055744.0;   // 1.5 * 2^52: adding and subtracting rounds to an integer
    c.kappa = kappa;
    c.e[0] = 1.;
    for (int n = 1; n < 14; n++) c.e[n] = c.e[n - 1] / n;
    return c;
}
DC_HD int dc_hi_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2hiint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b >> 32);
#endif
}
DC_HD double dc_with_hi_word(double x, int hi)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, __double2loint(x));
#else
    long long b;
    memcpy(&b, &x, 8);
    b = ((long long)hi << 32) | (b & 0xffffffffLL);
    memcpy(&x, &b, 8);
    return x;
#endif
}
DC_HD double dc_fma(double a, double b, double c)
{
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
// 1 / x for finite x of ordinary magnitude (pressures, Exner differences): the hardware seed
// (rcp.approx.ftz.f64, ~20 bits) and two Newton steps, no special-case branches -- 5 instructions
// against ~15 of __drcp_rn with its slow-path call; <= 1 ulp
DC_HD double dc_rcp(double x)
{
#if defined(__CUDA_ARCH__)
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = __fma_rn(-x, r, 1.);
    r = __fma_rn(r, e, r);
    e = __fma_rn(-x, r, 1.);
    return __fma_rn(r, e, r);
#else
    return 1. / x;
#endif
}
DC_HD double pow_kappa(double x, const PowCoef &c)
{
    // log(x): x = 2^k * m, m in [sqrt(1/2), sqrt(2))
    int hi = dc_hi_word(x);
    int k = (hi >> 20) - 1023;
    hi &= 0x000fffff;
    const int i = (hi + 0x95f64) & 0x100000;
    k += i >> 20;
    const double m = dc_with_hi_word(x, hi | (i ^ 0x3ff00000));
    const double f = m - 1., s = f * dc_rcp(2. + f), dk = (double)k, z = s * s, w = z * z;
    const double t1 = w * (c.Lg2 + w * (c.Lg4 + w * c.Lg6));
    const double t2 = z * (c.Lg1 + w * (c.Lg3 + w * (c.Lg5 + w * c.Lg7)));
    const double hfsq = 0.5 * f * f;
    const double lg = dk * c.ln2_hi - ((hfsq - (s * (hfsq + (t2 + t1)) + dk * c.ln2_lo)) - f);
    // exp(kappa * log x)
    const double y = c.kappa * lg;
    const double kf = (y * c.inv_ln2 + c.big) - c.big;
    const double r = dc_fma(-kf, c.ln2_lo, dc_fma(-kf, c.ln2_hi, y));
    double p = c.e[13];
    for (int n = 12; n >= 0; n--) p = dc_fma(p, r, c.e[n]);
    return dc_with_hi_word(p, dc_hi_word(p) + ((int)kf << 20));
}

DC_HD int dc_lo_word(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2loint(x);
#else
    long long b;
    memcpy(&b, &x, 8);
    return (int)(b & 0xffffffffLL);
#endif
}
DC_HD double dc_from_words(int hi, int lo)
{
#if defined(__CUDA_ARCH__)
    return __hiloint2double(hi, lo);
#else
    long long b = ((long long)hi << 32) | ((long long)lo & 0xffffffffLL);
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}
// cold path of pow_kappa_tab, kept out of line so that the callers' loops stay tight
#if defined(__CUDACC__)
__host__ __device__ __noinline__
#else
inline
#endif
double pow_kappa_cold(double x, double kappa) { return pow(x, kappa); }
// x^kappa through the table of make_pow_table (see PowCoef); arguments outside the tabulated
// binades (or a handle without a table) take the library pow
// `tab`: c.tab (global memory, read through the read-only path) or a copy of it in shared
// memory (`shared` = true: plain loads)
DC_HD double pow_kappa_tab(double x, const PowCoef &c, const double *tab, bool shared)
{
    const int hi = dc_hi_word(x);
    const int te = (hi >> 20) - (1023 + POW_EMIN);
    if (tab == nullptr || (unsigned)te >= (unsigned)POW_NE) return pow_kappa_cold(x, c.kappa);
    const int j = (hi >> (20 - POW_JBITS)) & (POW_NJ - 1);
    const double m = dc_from_words((hi & 0x000fffff) | 0x3ff00000, dc_lo_word(x));
    const double *rt = tab + 2 * (te * POW_NJ + j);
    double r, T;
#if defined(__CUDA_ARCH__)
    if (shared) {
        const double2 v = *reinterpret_cast<const double2 *>(rt);
        r = v.x; T = v.y;
    } else {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(rt));
        r = v.x; T = v.y;
    }
#else
    (void)shared;
    r = rt[0]; T = rt[1];
#endif
    const double tt = dc_fma(m, r, -1.);
    double q = c.b[8];
    for (int n = 7; n >= 1; n--) q = dc_fma(q, tt, c.b[n]);
    return dc_fma(T, q * tt, T);
}
DC_HD double pow_kappa_tab(double x, const PowCoef &c) { return pow_kappa_tab(x, c, c.tab, false); }

// Table-driven natural logarithm for the production build (moisture interface values,
// comp_VARVB_log): x = 2^e * m, t = fma(m, r_j, -1) as in pow_kappa_tab,
//   log x = e ln2 + log(1 / r_j) + log1p(t),   log1p(t) = t - t^2/2 + ... - t^8/8  (|t| <= 2^-7)
// with ln2 split so that e * ln2_hi is exact.  ~1 ulp away from x = 1 (tracer mixing ratios are
// < 0.1).  The table {r_j, log(1/r_j)} travels in the kernel parameter block (2 KB).
struct LogCoef {
    double ln2_hi, ln2_lo;
    double c[9];                   // (-1)^(n+1) / n, n = 1..8 (c[0] unused)
    double tab[POW_NJ][2];         // r_j, log(1 / r_j) from the rounded r_j
};
inline LogCoef make_log_coef()
{
    LogCoef L;
    L.ln2_hi = 6.93147180369123816490e-01;
    L.ln2_lo = 1.90821492927058770002e-10;
    L.c[0] = 0.;
    for (int n = 1; n < 9; n++) L.c[n] = ((n & 1) ? 1. : -1.) / n;
    for (int j = 0; j < POW_NJ; j++) {
        const double r = 1. / (1. + (j + 0.5) / POW_NJ);
        L.tab[j][0] = r;
        L.tab[j][1] = (double)logl(1.0L / (long double)r);
    }
    return L;
}
// `tab` = L.tab or a copy of it in shared memory: a constant-bank read with a per-thread index is
// replayed once per distinct address (ncu: 45 % of the moisture kernel's stalls), a shared-memory
// read is not
DC_HD double log_tab(double x, const LogCoef &L, const double (*tab)[2])
{
    const int hi = dc_hi_word(x);
    const int e = (hi >> 20) - 1023;
    const int j = (hi >> (20 - POW_JBITS)) & (POW_NJ - 1);
    const double m = dc_from_words((hi & 0x000fffff) | 0x3ff00000, dc_lo_word(x));
    const double r = tab[j][0], lm = tab[j][1];
    const double t = dc_fma(m, r, -1.);
    double p = L.c[8];
    for (int n = 7; n >= 1; n--) p = dc_fma(p, t, L.c[n]);
    const double de = (double)e;
    return dc_fma(de, L.ln2_hi, lm + dc_fma(de, L.ln2_lo, p * t));
}

// ---------------------------------------------------------------------------------------
// physics coupling terms (vertical turbulent transport, surface fluxes)
// ---------------------------------------------------------------------------------------
constexpr double con_Lh = 2264E3;   // io_constants.py:25

// dyn_functions.py:276-377 (interior interfaces; 0 at k = 0 and nz).  K = KMOM, R = RHOVB,
// P = PHI[k], Q = PHI[k-1], C = COLP, A = cell area; d = along the wind, p = perpendicular
struct Six {
    double c, dm1, pm1, pp1, pm1_dm1, pp1_dm1;
};
// division by the gravity constant: an IEEE division in the strict build, a multiplication by
// the compile-time reciprocal in the production build
DC_HD Div div_g() { return mkdiv(con_g, 1. / con_g); }
// the six-point average of dyn_functions.py:276-422 around a u- or v-point (rigid walls in
// the perpendicular direction at p_ind == 1 and p_ind == np when rigid_wall)
DC_HD double six_avg(const Six &V, bool rigid_wall, int p_ind, int np)
{
    if (rigid_wall && p_ind == 1) return 0.25 * (V.pp1_dm1 + V.pp1 + V.dm1 + V.c);
    if (rigid_wall && p_ind == np) return 0.25 * (V.dm1 + V.c + V.pm1_dm1 + V.pm1);
    return 0.125 * (V.pp1_dm1 + V.pp1 + 2. * V.dm1 + 2. * V.c + V.pm1_dm1 + V.pm1);
}
// dyn_functions.py:383-422.  The division by con_g belongs to the reference function whatever
// VAR is (it is also applied to RHO and the surface momentum fluxes, and a second time to
// PHIVB by the callers): reproduced as is.
DC_HD double interp_VAR_ds(const Six &V, bool rigid_wall, int p_ind, int np)
{
    return six_avg(V, rigid_wall, p_ind, np) / div_g();
}
// dyn_functions.py:276-377 (interior interfaces; 0 at k = 0 and nz), split in its three
// factors so that a column march can carry the altitude of level k to interface k+1:
//   colpakmom_ds   six-point average of RHOVB*COLP*A*KMOM (R, C, A, K)
//   interp_VAR_ds  of PHI[k-1] and PHI[k] = ALT_ds_km1, ALT_ds (same expression as the
//                  reference's inline averages / con_g)
//   kmom_dwinddz   COLPAKMOM_ds_ks * (DWIND_km1 - DWIND) / (ALT_ds_km1 - ALT_ds)
DC_HD double colpakmom_ds(const Six &K, const Six &R, const Six &C, const Six &A, bool rigid_wall,
                          int p_ind, int np)
{
    if (rigid_wall && p_ind == 1)
        return 0.25 * (R.pp1_dm1 * C.pp1_dm1 * A.pp1_dm1 * K.pp1_dm1 +
                       R.pp1 * C.pp1 * A.pp1 * K.pp1 + R.dm1 * C.dm1 * A.dm1 * K.dm1 +
                       R.c * C.c * A.c * K.c);
    if (rigid_wall && p_ind == np)
        return 0.25 * (R.dm1 * C.dm1 * A.dm1 * K.dm1 + R.c * C.c * A.c * K.c +
                       R.pm1_dm1 * C.pm1_dm1 * A.pm1_dm1 * K.pm1_dm1 +
                       R.pm1 * C.pm1 * A.pm1 * K.pm1);
    return 0.125 * (R.pp1_dm1 * C.pp1_dm1 * A.pp1_dm1 * K.pp1_dm1 + R.pp1 * C.pp1 * A.pp1 * K.pp1 +
                    2. * R.dm1 * C.dm1 * A.dm1 * K.dm1 + 2. * R.c * C.c * A.c * K.c +
                    R.pm1_dm1 * C.pm1_dm1 * A.pm1_dm1 * K.pm1_dm1 + R.pm1 * C.pm1 * A.pm1 * K.pm1);
}
DC_HD double kmom_dwinddz(double COLPAKMOM_ds_ks, double DWIND, double DWIND_km1,
                          double ALT_ds_km1, double ALT_ds)
{
    const double dDWINDdz_ks = ((DWIND_km1 - DWIND) / (ALT_ds_km1 - ALT_ds));
    return COLPAKMOM_ds_ks * dDWINDdz_ks;
}
// dyn_UFLX.py:136-170, dyn_VFLX.py:134-166: vertical turbulent transport of momentum
DC_HD double turb_momentum(double Kd, double Kd_kp1, double SMOMFLX_s, double ALTVB_s,
                           double ALTVB_kp1_s, double RHO_s, int k, int nz)
{
    if (k == 0) return ((0. - Kd_kp1) / ((ALTVB_s - ALTVB_kp1_s) * RHO_s));
    if (k == nz - 1) return ((Kd + SMOMFLX_s) / ((ALTVB_s - ALTVB_kp1_s) * RHO_s));
    return ((Kd - Kd_kp1) / ((ALTVB_s - ALTVB_kp1_s) * RHO_s));
}
// dyn_functions.py:26-67 (turb_flux_tendency_py), split for a column march:
//   turb_iface_flux(k)  = (VAR[k-1] - VAR[k]) / (ALT[k-1] - ALT[k]) * RHOVB[k] * KVAR[k]
// is the first term at level k and, evaluated at k+1, the second term at level k -- the same
// expression on the same operands, so a march computes it once per interface.
DC_HD double turb_iface_flux(double VAR_km1, double VAR, double ALT_km1, double ALT, double RHOVB,
                             double KVAR)
{
    return ((VAR_km1 - VAR) / (ALT_km1 - ALT) * RHOVB * KVAR);
}
// flux_k / flux_kp1: turb_iface_flux at the interfaces k and k+1 (ignored at the model top /
// at the surface, where the reference puts +0 / the surface flux)
DC_HD double turb_flux_tendency(double flux_k, double flux_kp1, double ALTVB, double ALTVB_kp1,
                                double RHO, double COLP, double surf_flux_VAR, int k, int nz)
{
    if (k == 0) return COLP * ((+0. - flux_kp1) / ((ALTVB - ALTVB_kp1) * RHO));
    if (k == nz - 1) return COLP * ((+flux_k + surf_flux_VAR) / ((ALTVB - ALTVB_kp1) * RHO));
    return COLP * ((+flux_k - flux_kp1) / ((ALTVB - ALTVB_kp1) * RHO));
}

// what a column march carries for turb_flux_tendency: altitude of the level and of its upper
// interface, turbulent flux through that interface (unused at the model top)
struct TurbMarch {
    double alt, altvb, flux;
    DC_HD static TurbMarch top(double PHI_0, double PHIVB_0)
    {
        return TurbMarch{PHI_0 / div_g(), PHIVB_0 / div_g(), 0.};
    }
    // tendency of level k; VAR = VAR[k]; the *_kp1 arguments are only read for k < nz-1 except
    // PHIVB_kp1.  Advances the state to level k+1.
    DC_HD double step(double VAR, double VAR_kp1, double PHI_kp1, double PHIVB_kp1, double RHOVB_kp1,
                      double KVAR_kp1, double RHO, double COLP, double surf_flux_VAR, int k, int nz)
    {
        const double altvb_kp1 = PHIVB_kp1 / div_g();
        double alt_kp1 = alt, flux_kp1 = 0.;
        if (k < nz - 1) {
            alt_kp1 = PHI_kp1 / div_g();
            flux_kp1 = turb_iface_flux(VAR, VAR_kp1, alt, alt_kp1, RHOVB_kp1, KVAR_kp1);
        }
        const double t =
            turb_flux_tendency(flux, flux_kp1, altvb, altvb_kp1, RHO, COLP, surf_flux_VAR, k, nz);
        alt = alt_kp1;
        altvb = altvb_kp1;
        flux = flux_kp1;
        return t;
    }
};

// dyn_timestep.py:34-38
DC_HD double euler_forward_pw(double VAR, double dVARdt, Div COLP, double COLP_OLD, double dt)
{
    return VAR * COLP_OLD / COLP + dt * dVARdt / COLP;
}
// dyn_timestep.py:40-52
DC_HD double interp_COLPA_js(double COLP, double COLP_jm1, double COLP_im1, double COLP_ip1,
                             double COLP_jm1_ip1, double COLP_jm1_im1, double A, double A_jm1)
{
    // A depends on the row only: A_im1 = A_ip1 = A, A_jm1_ip1 = A_jm1_im1 = A_jm1
    return 1. / 8. * (COLP_jm1_ip1 * A_jm1 + COLP_ip1 * A + 2. * COLP_jm1 * A_jm1 +
                      2. * COLP * A + COLP_jm1_im1 * A_jm1 + COLP_im1 * A);
}
// dyn_timestep.py:54-79
DC_HD double interp_COLPA_is(double COLP, double COLP_im1, double COLP_jm1, double COLP_jp1,
                             double COLP_im1_jp1, double COLP_im1_jm1, double A, double A_jm1,
                             double A_jp1, int j, int ny)
{
    if (j == 1)
        return 1. / 4. * (COLP_im1_jp1 * A_jp1 + COLP_jp1 * A_jp1 + COLP_im1 * A + COLP * A);
    else if (j == ny)
        return 1. / 4. * (COLP_im1_jm1 * A_jm1 + COLP_jm1 * A_jm1 + COLP_im1 * A + COLP * A);
    return 1. / 8. * (COLP_im1_jp1 * A_jp1 + COLP_jp1 * A_jp1 + 2. * COLP_im1 * A +
                      2. * COLP * A + COLP_im1_jm1 * A_jm1 + COLP_jm1 * A_jm1);
}

// POTT ring of the diagnostics sweep (see PrimaryDiagBody::march): levels in flight and slots
#ifndef DC_DIAG_PF
#define DC_DIAG_PF 3
#endif
constexpr int DIAG_PF = DC_DIAG_PF > 0 ? DC_DIAG_PF : 1;
constexpr int DIAG_SLOTS = DIAG_PF < 2 ? 2 : DIAG_PF < 4 ? 4 : DIAG_PF < 8 ? 8 : 16;
static_assert(DIAG_PF < DIAG_SLOTS, "ring: one slot more than levels in flight");
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void dc_cp_async8(double *s, const double *gp)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gp) : "memory");
}
__device__ __forceinline__ void dc_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void dc_cp_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
#endif

}  // namespace dc
