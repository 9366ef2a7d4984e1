/*
 * dyncore_oracle.c -- CPU restatement of the reference's numba-CPU dynamical core.
 * TEST INFRASTRUCTURE ONLY (see dyncore_oracle.h).  Build: oracle/Makefile
 * (gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp): no FMA contraction, the
 * reference's evaluation order, libm pow/log/sin/cos as numba lowers them.
 *
 * Every function cites the reference file:line it restates
 * (paths relative to the reference checkout).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dyncore_oracle.h"

/* io_constants.py:16-21 */
static const double con_g = 9.81;
static const double con_rE = 6371000.;
static const double con_Rd = 287.058;
static const double con_cp = 1005.;
static const double con_Lh = 2264E3; /* io_constants.py:25 */
#define CON_KAPPA (con_Rd / con_cp)

#define I3(i, j, k, FNY, FNZ) \
    (((size_t)(i) * (size_t)(FNY) + (size_t)(j)) * (size_t)(FNZ) + (size_t)(k))
/* index helpers; need nx, ny, nz in scope */
#define M(i, j, k) I3(i, j, k, ny + 2, nz)       /* mass / x-staggered, full levels   */
#define MS(i, j, k) I3(i, j, k, ny + 2, nz + 1)  /* mass / x-staggered, interfaces    */
#define Y(i, j, k) I3(i, j, k, ny + 3, nz)       /* y- / xy-staggered, full levels    */
#define YS(i, j, k) I3(i, j, k, ny + 3, nz + 1)  /* y-staggered, interfaces           */
#define M2(i, j) ((size_t)(i) * (size_t)(ny + 2) + (size_t)(j))
#define Y2(i, j) ((size_t)(i) * (size_t)(ny + 3) + (size_t)(j))

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* misc_boundaries.py:22-42 (exchange_BC_cpu)                                 */
/* ------------------------------------------------------------------------- */
void orc_exchange_BC(const orc_grid *g, double *F, int fnx, int fny, int fnz)
{
    const int nx = g->nx, ny = g->ny, nxs = nx + 1, nys = ny + 1;
    const size_t row = (size_t)fny * (size_t)fnz; /* one i-slab */
    /* zonal boundaries */
    if (fnx == nxs + 2) { /* staggered in x */
        memcpy(F + 0 * row, F + (size_t)(nxs - 1) * row, row * sizeof(double));
        memcpy(F + (size_t)nxs * row, F + 1 * row, row * sizeof(double));
        memcpy(F + (size_t)(nxs + 1) * row, F + 2 * row, row * sizeof(double));
    } else {
        memcpy(F + 0 * row, F + (size_t)nx * row, row * sizeof(double));
        memcpy(F + (size_t)(nx + 1) * row, F + 1 * row, row * sizeof(double));
    }
    /* meridional boundaries */
    if (fny == nys + 2) { /* staggered in y */
        const int js[4] = {0, 1, nys, nys + 1};
        for (int i = 0; i < fnx; i++)
            for (int m = 0; m < 4; m++)
                for (int k = 0; k < fnz; k++)
                    F[I3(i, js[m], k, fny, fnz)] = 0.;
    } else {
        for (int i = 0; i < fnx; i++)
            for (int k = 0; k < fnz; k++) {
                F[I3(i, 0, k, fny, fnz)] = F[I3(i, 1, k, fny, fnz)];
                F[I3(i, ny + 1, k, fny, fnz)] = F[I3(i, ny, k, fny, fnz)];
            }
    }
}

/* ------------------------------------------------------------------------- */
/* dyn_continuity.py:40-47 (point functions), :170-228 (launch_numba_cpu)     */
/* ------------------------------------------------------------------------- */
static inline double calc_UFLX(double UWIND, double COLP, double COLP_im1, double dyis)
{
    return (COLP_im1 + COLP) / 2. * UWIND * dyis;
}
static inline double calc_VFLX(double VWIND, double COLP, double COLP_jm1, double dxjs)
{
    return (COLP_jm1 + COLP) / 2. * VWIND * dxjs;
}
static inline double calc_FLXDIV(double UFLX, double UFLX_ip1, double VFLX, double VFLX_jp1,
                                 double dsigma, double A)
{
    return (+UFLX_ip1 - UFLX + VFLX_jp1 - VFLX) * dsigma / A;
}

void orc_continuity(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *COLP = f->COLP, *U = f->UWIND, *V = f->VWIND;
    const double dt = g->dt;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
            for (int k = 0; k < nz; k++) {
                double UFLX_i = calc_UFLX(U[M(i, j, k)], COLP[M2(i, j)], COLP[M2(i - 1, j)],
                                          g->dyis[M2(i, j)]);
                double UFLX_ip1 = calc_UFLX(U[M(i + 1, j, k)], COLP[M2(i + 1, j)],
                                            COLP[M2(i, j)], g->dyis[M2(i + 1, j)]);
                double VFLX_j = calc_VFLX(V[Y(i, j, k)], COLP[M2(i, j)], COLP[M2(i, j - 1)],
                                          g->dxjs[Y2(i, j)]);
                double VFLX_jp1 = calc_VFLX(V[Y(i, j + 1, k)], COLP[M2(i, j + 1)],
                                            COLP[M2(i, j)], g->dxjs[Y2(i, j + 1)]);
                f->UFLX[M(i, j, k)] = UFLX_i;
                f->VFLX[Y(i, j, k)] = VFLX_j;
                f->FLXDIV[M(i, j, k)] = calc_FLXDIV(UFLX_i, UFLX_ip1, VFLX_j, VFLX_jp1,
                                                    g->dsigma[k], g->A[M2(i, j)]);
            }
            /* dCOLPdt = -FLXDIV.sum(axis=2): numba's axis sum is a sequential ascending
             * k loop starting from 0.0 */
            double s = 0.;
            for (int k = 0; k < nz; k++) s += f->FLXDIV[M(i, j, k)];
            f->dCOLPdt[M2(i, j)] = -s;
            f->COLP_NEW[M2(i, j)] = f->COLP_OLD[M2(i, j)] + dt * f->dCOLPdt[M2(i, j)];
            /* vertical wind, dyn_continuity.py:222-228 */
            double flxdivsum = f->FLXDIV[M(i, j, 0)];
            for (int k = 1; k < nz; k++) {
                f->WWIND[MS(i, j, k)] = (-flxdivsum / f->COLP_NEW[M2(i, j)] -
                                         g->sigma_vb[k] * f->dCOLPdt[M2(i, j)] /
                                             f->COLP_NEW[M2(i, j)]);
                flxdivsum += f->FLXDIV[M(i, j, k)];
            }
        }
    /* dyn_org_discretizations.py:114-117 */
    orc_exchange_BC(g, f->UFLX, nx + 3, ny + 2, nz);
    orc_exchange_BC(g, f->VFLX, nx + 2, ny + 3, nz);
    orc_exchange_BC(g, f->WWIND, nx + 2, ny + 2, nz + 1);
    orc_exchange_BC(g, f->COLP_NEW, nx + 2, ny + 2, 1);
}

/* ------------------------------------------------------------------------- */
/* dyn_functions.py:211-270 (interp_WWIND_UVWIND_py), interior interfaces only */
/* ------------------------------------------------------------------------- */
static inline double interp_WWIND_UVWIND(
    double DWIND, double DWIND_km1, double WWIND, double WWIND_dm1, double WWIND_pm1,
    double WWIND_pp1, double WWIND_pm1_dm1, double WWIND_pp1_dm1, double COLP_NEW,
    double COLP_NEW_dm1, double COLP_NEW_pm1, double COLP_NEW_pp1, double COLP_NEW_pm1_dm1,
    double COLP_NEW_pp1_dm1, double A, double A_dm1, double A_pm1, double A_pp1,
    double A_pm1_dm1, double A_pp1_dm1, double dsigma, double dsigma_km1, int rigid_wall,
    int p_ind, int np)
{
    double COLPAWWIND_ds_ks;
    if (rigid_wall && p_ind == 1) {
        COLPAWWIND_ds_ks = 0.25 * (COLP_NEW_pp1_dm1 * A_pp1_dm1 * WWIND_pp1_dm1 +
                                   COLP_NEW_pp1 * A_pp1 * WWIND_pp1 +
                                   COLP_NEW_dm1 * A_dm1 * WWIND_dm1 + COLP_NEW * A * WWIND);
    } else if (rigid_wall && p_ind == np) {
        COLPAWWIND_ds_ks = 0.25 * (COLP_NEW_dm1 * A_dm1 * WWIND_dm1 + COLP_NEW * A * WWIND +
                                   COLP_NEW_pm1_dm1 * A_pm1_dm1 * WWIND_pm1_dm1 +
                                   COLP_NEW_pm1 * A_pm1 * WWIND_pm1);
    } else {
        COLPAWWIND_ds_ks = 0.125 * (COLP_NEW_pp1_dm1 * A_pp1_dm1 * WWIND_pp1_dm1 +
                                    COLP_NEW_pp1 * A_pp1 * WWIND_pp1 +
                                    2. * COLP_NEW_dm1 * A_dm1 * WWIND_dm1 +
                                    2. * COLP_NEW * A * WWIND +
                                    COLP_NEW_pm1_dm1 * A_pm1_dm1 * WWIND_pm1_dm1 +
                                    COLP_NEW_pm1 * A_pm1 * WWIND_pm1);
    }
    double DWIND_ks = ((dsigma * DWIND_km1 + dsigma_km1 * DWIND) / (dsigma + dsigma_km1));
    return COLPAWWIND_ds_ks * DWIND_ks;
}

/* dyn_functions.py:276-377 (interp_KMOM_dUVWINDdz_py), interior interfaces only.
 * K = KMOM, R = RHOVB, P = PHI[k], Q = PHI[k-1], C = the COLP argument, A = cell area;
 * suffixes as in the reference: d = along the wind, p = perpendicular. */
static inline double interp_KMOM_dUVWINDdz(
    double DWIND, double DWIND_km1, double K, double K_dm1, double K_pm1, double K_pp1,
    double K_pm1_dm1, double K_pp1_dm1, double R, double R_dm1, double R_pm1, double R_pp1,
    double R_pm1_dm1, double R_pp1_dm1, double P, double P_dm1, double P_pm1, double P_pp1,
    double P_pm1_dm1, double P_pp1_dm1, double Q, double Q_dm1, double Q_pm1, double Q_pp1,
    double Q_pm1_dm1, double Q_pp1_dm1, double C, double C_dm1, double C_pm1, double C_pp1,
    double C_pm1_dm1, double C_pp1_dm1, double A, double A_dm1, double A_pm1, double A_pp1,
    double A_pm1_dm1, double A_pp1_dm1, int rigid_wall, int p_ind, int np)
{
    double COLPAKMOM_ds_ks, ALT_ds_km1, ALT_ds;
    if (rigid_wall && p_ind == 1) {
        COLPAKMOM_ds_ks = 0.25 * (R_pp1_dm1 * C_pp1_dm1 * A_pp1_dm1 * K_pp1_dm1 +
                                  R_pp1 * C_pp1 * A_pp1 * K_pp1 + R_dm1 * C_dm1 * A_dm1 * K_dm1 +
                                  R * C * A * K);
        ALT_ds_km1 = 0.25 * (Q_pp1_dm1 + Q_pp1 + Q_dm1 + Q) / con_g;
        ALT_ds = 0.25 * (P_pp1_dm1 + P_pp1 + P_dm1 + P) / con_g;
    } else if (rigid_wall && p_ind == np) {
        COLPAKMOM_ds_ks = 0.25 * (R_dm1 * C_dm1 * A_dm1 * K_dm1 + R * C * A * K +
                                  R_pm1_dm1 * C_pm1_dm1 * A_pm1_dm1 * K_pm1_dm1 +
                                  R_pm1 * C_pm1 * A_pm1 * K_pm1);
        ALT_ds_km1 = 0.25 * (Q_dm1 + Q + Q_pm1_dm1 + Q_pm1) / con_g;
        ALT_ds = 0.25 * (P_dm1 + P + P_pm1_dm1 + P_pm1) / con_g;
    } else {
        COLPAKMOM_ds_ks =
            0.125 * (R_pp1_dm1 * C_pp1_dm1 * A_pp1_dm1 * K_pp1_dm1 + R_pp1 * C_pp1 * A_pp1 * K_pp1 +
                     2. * R_dm1 * C_dm1 * A_dm1 * K_dm1 + 2. * R * C * A * K +
                     R_pm1_dm1 * C_pm1_dm1 * A_pm1_dm1 * K_pm1_dm1 + R_pm1 * C_pm1 * A_pm1 * K_pm1);
        ALT_ds_km1 =
            0.125 * (Q_pp1_dm1 + Q_pp1 + 2. * Q_dm1 + 2. * Q + Q_pm1_dm1 + Q_pm1) / con_g;
        ALT_ds = 0.125 * (P_pp1_dm1 + P_pp1 + 2. * P_dm1 + 2. * P + P_pm1_dm1 + P_pm1) / con_g;
    }
    const double dDWINDdz_ks = ((DWIND_km1 - DWIND) / (ALT_ds_km1 - ALT_ds));
    return COLPAKMOM_ds_ks * dDWINDdz_ks;
}

/* dyn_functions.py:383-422 (interp_VAR_ds_py).  The division by con_g is part of the
 * reference function whatever VAR is (it is also applied to RHO and the surface momentum
 * fluxes, and a second time to PHIVB by the callers): reproduced as is. */
static inline double interp_VAR_ds(double VAR, double VAR_dm1, double VAR_pm1, double VAR_pp1,
                                   double VAR_pm1_dm1, double VAR_pp1_dm1, int rigid_wall,
                                   int p_ind, int np)
{
    if (rigid_wall && p_ind == 1)
        return 0.25 * (VAR_pp1_dm1 + VAR_pp1 + VAR_dm1 + VAR) / con_g;
    if (rigid_wall && p_ind == np)
        return 0.25 * (VAR_dm1 + VAR + VAR_pm1_dm1 + VAR_pm1) / con_g;
    return 0.125 * (VAR_pp1_dm1 + VAR_pp1 + 2. * VAR_dm1 + 2. * VAR + VAR_pm1_dm1 + VAR_pm1) /
           con_g;
}

/* dyn_functions.py:26-67 (turb_flux_tendency_py).  The neighbour levels that a branch does
 * not use are never dereferenced here (the reference reads one past the column at k = nz-1) */
static inline double turb_flux_tendency(double PHI, double PHI_kp1, double PHI_km1, double PHIVB,
                                        double PHIVB_kp1, double VAR, double VAR_kp1,
                                        double VAR_km1, double KVAR, double KVAR_kp1, double RHO,
                                        double RHOVB, double RHOVB_kp1, double COLP,
                                        double surf_flux_VAR, int k, int nz)
{
    const double ALT = PHI / con_g, ALT_kp1 = PHI_kp1 / con_g, ALT_km1 = PHI_km1 / con_g;
    const double ALTVB = PHIVB / con_g, ALTVB_kp1 = PHIVB_kp1 / con_g;
    if (k == 0)
        return COLP * ((+0. - ((VAR - VAR_kp1) / (ALT - ALT_kp1) * RHOVB_kp1 * KVAR_kp1)) /
                       ((ALTVB - ALTVB_kp1) * RHO));
    if (k == nz - 1)
        return COLP * ((+((VAR_km1 - VAR) / (ALT_km1 - ALT) * RHOVB * KVAR) + surf_flux_VAR) /
                       ((ALTVB - ALTVB_kp1) * RHO));
    return COLP * ((+((VAR_km1 - VAR) / (ALT_km1 - ALT) * RHOVB * KVAR) -
                    ((VAR - VAR_kp1) / (ALT - ALT_kp1) * RHOVB_kp1 * KVAR_kp1)) /
                   ((ALTVB - ALTVB_kp1) * RHO));
}

/* dyn_UVFLX_prepare.py:249-439 (launch_numba_cpu_prep_adv) */
static void orc_UVFLX_prep_adv(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz, nxs = nx + 1, nys = ny + 1;
    const double *U = f->UWIND, *V = f->VWIND, *W = f->WWIND, *CN = f->COLP_NEW, *A = g->A;
    const double *ds = g->dsigma;

    /* dyn_UVFLX_prepare.py:262-278 */
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nxs; i++)
        for (int j = 1; j <= ny; j++) {
            f->WWIND_UWIND[MS(i, j, 0)] = 0.;
            f->WWIND_UWIND[MS(i, j, nz)] = 0.;
            for (int k = 1; k < nz; k++)
                f->WWIND_UWIND[MS(i, j, k)] = interp_WWIND_UVWIND(
                    U[M(i, j, k)], U[M(i, j, k - 1)], W[MS(i, j, k)], W[MS(i - 1, j, k)],
                    W[MS(i, j - 1, k)], W[MS(i, j + 1, k)], W[MS(i - 1, j - 1, k)],
                    W[MS(i - 1, j + 1, k)], CN[M2(i, j)], CN[M2(i - 1, j)], CN[M2(i, j - 1)],
                    CN[M2(i, j + 1)], CN[M2(i - 1, j - 1)], CN[M2(i - 1, j + 1)], A[M2(i, j)],
                    A[M2(i - 1, j)], A[M2(i, j - 1)], A[M2(i, j + 1)], A[M2(i - 1, j - 1)],
                    A[M2(i - 1, j + 1)], ds[k], ds[k - 1], 1, j, ny);
        }
    /* dyn_UVFLX_prepare.py:280-296 */
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= nys; j++) {
            f->WWIND_VWIND[YS(i, j, 0)] = 0.;
            f->WWIND_VWIND[YS(i, j, nz)] = 0.;
            for (int k = 1; k < nz; k++)
                f->WWIND_VWIND[YS(i, j, k)] = interp_WWIND_UVWIND(
                    V[Y(i, j, k)], V[Y(i, j, k - 1)], W[MS(i, j, k)], W[MS(i, j - 1, k)],
                    W[MS(i - 1, j, k)], W[MS(i + 1, j, k)], W[MS(i - 1, j - 1, k)],
                    W[MS(i + 1, j - 1, k)], CN[M2(i, j)], CN[M2(i, j - 1)], CN[M2(i - 1, j)],
                    CN[M2(i + 1, j)], CN[M2(i - 1, j - 1)], CN[M2(i + 1, j - 1)], A[M2(i, j)],
                    A[M2(i, j - 1)], A[M2(i - 1, j)], A[M2(i + 1, j)], A[M2(i - 1, j - 1)],
                    A[M2(i + 1, j - 1)], ds[k], ds[k - 1], 0, i, nx);
        }

    /* dyn_UVFLX_prepare.py:298-343: K * d(wind)/dz on the interfaces (0 at k = 0 and nz) */
    if (g->i_coupling) {
        const double *K = f->KMOM, *R = f->RHOVB, *P = f->PHI, *C = f->COLP;
#pragma omp parallel for schedule(static)
        for (int i = 1; i <= nxs; i++)
            for (int j = 1; j <= ny; j++) {
                f->KMOM_dUWINDdz[MS(i, j, 0)] = 0.;
                f->KMOM_dUWINDdz[MS(i, j, nz)] = 0.;
                for (int k = 1; k < nz; k++)
                    f->KMOM_dUWINDdz[MS(i, j, k)] = interp_KMOM_dUVWINDdz(
                        U[M(i, j, k)], U[M(i, j, k - 1)], K[MS(i, j, k)], K[MS(i - 1, j, k)],
                        K[MS(i, j - 1, k)], K[MS(i, j + 1, k)], K[MS(i - 1, j - 1, k)],
                        K[MS(i - 1, j + 1, k)], R[MS(i, j, k)], R[MS(i - 1, j, k)],
                        R[MS(i, j - 1, k)], R[MS(i, j + 1, k)], R[MS(i - 1, j - 1, k)],
                        R[MS(i - 1, j + 1, k)], P[M(i, j, k)], P[M(i - 1, j, k)],
                        P[M(i, j - 1, k)], P[M(i, j + 1, k)], P[M(i - 1, j - 1, k)],
                        P[M(i - 1, j + 1, k)], P[M(i, j, k - 1)], P[M(i - 1, j, k - 1)],
                        P[M(i, j - 1, k - 1)], P[M(i, j + 1, k - 1)], P[M(i - 1, j - 1, k - 1)],
                        P[M(i - 1, j + 1, k - 1)], C[M2(i, j)], C[M2(i - 1, j)], C[M2(i, j - 1)],
                        C[M2(i, j + 1)], C[M2(i - 1, j - 1)], C[M2(i - 1, j + 1)], A[M2(i, j)],
                        A[M2(i - 1, j)], A[M2(i, j - 1)], A[M2(i, j + 1)], A[M2(i - 1, j - 1)],
                        A[M2(i - 1, j + 1)], 1, j, ny);
            }
#pragma omp parallel for schedule(static)
        for (int i = 1; i <= nx; i++)
            for (int j = 1; j <= nys; j++) {
                f->KMOM_dVWINDdz[YS(i, j, 0)] = 0.;
                f->KMOM_dVWINDdz[YS(i, j, nz)] = 0.;
                for (int k = 1; k < nz; k++)
                    f->KMOM_dVWINDdz[YS(i, j, k)] = interp_KMOM_dUVWINDdz(
                        V[Y(i, j, k)], V[Y(i, j, k - 1)], K[MS(i, j, k)], K[MS(i, j - 1, k)],
                        K[MS(i - 1, j, k)], K[MS(i + 1, j, k)], K[MS(i - 1, j - 1, k)],
                        K[MS(i + 1, j - 1, k)], R[MS(i, j, k)], R[MS(i, j - 1, k)],
                        R[MS(i - 1, j, k)], R[MS(i + 1, j, k)], R[MS(i - 1, j - 1, k)],
                        R[MS(i + 1, j - 1, k)], P[M(i, j, k)], P[M(i, j - 1, k)],
                        P[M(i - 1, j, k)], P[M(i + 1, j, k)], P[M(i - 1, j - 1, k)],
                        P[M(i + 1, j - 1, k)], P[M(i, j, k - 1)], P[M(i, j - 1, k - 1)],
                        P[M(i - 1, j, k - 1)], P[M(i + 1, j, k - 1)], P[M(i - 1, j - 1, k - 1)],
                        P[M(i + 1, j - 1, k - 1)], C[M2(i, j)], C[M2(i, j - 1)], C[M2(i - 1, j)],
                        C[M2(i + 1, j)], C[M2(i - 1, j - 1)], C[M2(i + 1, j - 1)], A[M2(i, j)],
                        A[M2(i, j - 1)], A[M2(i - 1, j)], A[M2(i + 1, j)], A[M2(i - 1, j - 1)],
                        A[M2(i + 1, j - 1)], 0, i, nx);
            }
    }

    /* dyn_UVFLX_prepare.py:345-436 with dyn_functions.py:429-536; only the neighbours
     * each flux uses are loaded (the reference loads a full 3x3 block, part of it out of
     * bounds and unused) */
    const double *u = f->UFLX, *v = f->VFLX;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nxs; i++)
        for (int j = 1; j <= nys; j++)
            for (int k = 0; k < nz; k++) {
                /* (is, js): calc_momentum_fluxes_isjs_py */
                f->CFLX[Y(i, j, k)] =
                    1. / 12. * (v[Y(i - 1, j - 1, k)] + v[Y(i, j - 1, k)] +
                                2. * (v[Y(i - 1, j, k)] + v[Y(i, j, k)]) +
                                v[Y(i - 1, j + 1, k)] + v[Y(i, j + 1, k)]);
                f->QFLX[Y(i, j, k)] =
                    1. / 12. * (u[M(i - 1, j - 1, k)] + u[M(i - 1, j, k)] +
                                2. * (u[M(i, j - 1, k)] + u[M(i, j, k)]) +
                                u[M(i + 1, j - 1, k)] + u[M(i + 1, j, k)]);
                if (i <= nx) { /* (i, js): calc_momentum_fluxes_ijs_py */
                    f->DFLX[Y(i, j, k)] =
                        1. / 24. * (v[Y(i, j - 1, k)] + 2. * v[Y(i, j, k)] + v[Y(i, j + 1, k)] +
                                    u[M(i, j - 1, k)] + u[M(i, j, k)] + u[M(i + 1, j - 1, k)] +
                                    u[M(i + 1, j, k)]);
                    f->EFLX[Y(i, j, k)] =
                        1. / 24. * (v[Y(i, j - 1, k)] + 2. * v[Y(i, j, k)] + v[Y(i, j + 1, k)] -
                                    u[M(i, j - 1, k)] - u[M(i, j, k)] - u[M(i + 1, j - 1, k)] -
                                    u[M(i + 1, j, k)]);
                }
                if (j <= ny) { /* (is, j): calc_momentum_fluxes_isj_py */
                    f->SFLX[M(i, j, k)] =
                        1. / 24. * (v[Y(i - 1, j, k)] + v[Y(i - 1, j + 1, k)] + v[Y(i, j, k)] +
                                    v[Y(i, j + 1, k)] + u[M(i - 1, j, k)] + 2. * u[M(i, j, k)] +
                                    u[M(i + 1, j, k)]);
                    f->TFLX[M(i, j, k)] =
                        1. / 24. * (v[Y(i - 1, j, k)] + v[Y(i - 1, j + 1, k)] + v[Y(i, j, k)] +
                                    v[Y(i, j + 1, k)] - u[M(i - 1, j, k)] - 2. * u[M(i, j, k)] -
                                    u[M(i + 1, j, k)]);
                }
                if (i <= nx && j <= ny) { /* (i, j): calc_momentum_fluxes_ij_py */
                    f->BFLX[M(i, j, k)] =
                        1. / 12. * (u[M(i, j - 1, k)] + u[M(i + 1, j - 1, k)] +
                                    2. * (u[M(i, j, k)] + u[M(i + 1, j, k)]) +
                                    u[M(i, j + 1, k)] + u[M(i + 1, j + 1, k)]);
                    f->RFLX[M(i, j, k)] =
                        1. / 12. * (v[Y(i - 1, j, k)] + v[Y(i - 1, j + 1, k)] +
                                    2. * (v[Y(i, j, k)] + v[Y(i, j + 1, k)]) +
                                    v[Y(i + 1, j, k)] + v[Y(i + 1, j + 1, k)]);
                }
            }
}

/* dyn_functions.py:541-568 (UVFLX_hor_adv_py) */
static inline double UVFLX_hor_adv(double DWIND, double DWIND_dm1, double DWIND_dp1,
                                   double DWIND_pm1, double DWIND_pp1, double DWIND_dm1_pm1,
                                   double DWIND_dm1_pp1, double DWIND_dp1_pm1,
                                   double DWIND_dp1_pp1, double BRFLX, double BRFLX_dm1,
                                   double CQFLX, double CQFLX_pp1, double DSFLX_dm1,
                                   double DSFLX_pp1, double ETFLX, double ETFLX_dm1_pp1,
                                   double sign_ETFLX_term)
{
    return (+BRFLX_dm1 * (DWIND_dm1 + DWIND) / 2. - BRFLX * (DWIND + DWIND_dp1) / 2.
            + CQFLX * (DWIND_pm1 + DWIND) / 2. - CQFLX_pp1 * (DWIND + DWIND_pp1) / 2.
            + DSFLX_dm1 * (DWIND_dm1_pm1 + DWIND) / 2. - DSFLX_pp1 * (DWIND + DWIND_dp1_pp1) / 2.
            + sign_ETFLX_term * (+ETFLX * (DWIND_dp1_pm1 + DWIND) / 2. -
                                 ETFLX_dm1_pp1 * (DWIND + DWIND_dm1_pp1) / 2.));
}

/* dyn_functions.py:177-207 (pre_grad_py) */
static inline double pre_grad(double PHI, double PHI_dm1, double COLP, double COLP_dm1,
                              double POTT, double POTT_dm1, double PVTF, double PVTF_dm1,
                              double PVTFVB, double PVTFVB_dm1, double PVTFVB_dm1_kp1,
                              double PVTFVB_kp1, double dsigma, double sigma_vb,
                              double sigma_vb_kp1, double dgrid)
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

/* dyn_functions.py:158-170 (num_dif_py) */
static inline double num_dif(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1,
                             double VAR_jp1, double VAR_dif_coef)
{
    return VAR_dif_coef * (+VAR_im1 + VAR_ip1 + VAR_jm1 + VAR_jp1 - 4. * VAR);
}

/* dyn_UFLX.py:39-66 */
static inline double coriolis_and_spherical_UWIND(double COLP, double COLP_im1, double VWIND,
                                                  double VWIND_im1, double VWIND_jp1,
                                                  double VWIND_im1_jp1, double UWIND,
                                                  double UWIND_im1, double UWIND_ip1,
                                                  double corf_is, double lat_is_rad,
                                                  double dlon_rad, double dlat_rad)
{
    return (con_rE * dlon_rad * dlat_rad / 2. *
            (COLP_im1 * (VWIND_im1 + VWIND_im1_jp1) / 2. *
                 (corf_is * con_rE * cos(lat_is_rad) +
                  (UWIND_im1 + UWIND) / 2. * sin(lat_is_rad)) +
             COLP * (VWIND + VWIND_jp1) / 2. *
                 (corf_is * con_rE * cos(lat_is_rad) +
                  (UWIND + UWIND_ip1) / 2. * sin(lat_is_rad))));
}

/* dyn_VFLX.py:39-64 */
static inline double coriolis_and_spherical_VWIND(double COLP, double COLP_jm1, double UWIND,
                                                  double UWIND_jm1, double UWIND_ip1,
                                                  double UWIND_ip1_jm1, double corf,
                                                  double corf_jm1, double lat_rad,
                                                  double lat_rad_jm1, double dlon_rad,
                                                  double dlat_rad)
{
    return (-con_rE * dlon_rad * dlat_rad / 2. *
            (COLP_jm1 * (UWIND_jm1 + UWIND_ip1_jm1) / 2. *
                 (corf_jm1 * con_rE * cos(lat_rad_jm1) +
                  (UWIND_jm1 + UWIND_ip1_jm1) / 2. * sin(lat_rad_jm1)) +
             COLP * (UWIND + UWIND_ip1) / 2. *
                 (corf * con_rE * cos(lat_rad) + (UWIND + UWIND_ip1) / 2. * sin(lat_rad))));
}

/* dyn_UFLX.py:339-434 (launcher) + :69-199 (add_up_tendencies_py).
 * Column i = nxs is not computed: the reference evaluates it from never-written (NaN)
 * halo entries of B/D/EFLX and the Euler step's result there is overwritten by the
 * periodic BC (UWIND[nxs] <- UWIND[1]). */
static void orc_UFLX_tendency(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *U = f->UWIND, *V = f->VWIND, *UFLX = f->UFLX;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++)
            for (int k = 0; k < nz; k++) {
                double BFLX = f->BFLX[M(i, j, k)];
                double CFLX = f->CFLX[Y(i, j, k)];
                double EFLX = f->EFLX[Y(i, j, k)];
                double DFLX_jp1 = f->DFLX[Y(i, j + 1, k)];
                double CFLX_jp1 = f->CFLX[Y(i, j + 1, k)];
                const int im1 = (i == 1) ? nx : i - 1; /* BCx, dyn_UFLX.py:367-374 */
                double BFLX_im1 = f->BFLX[M(im1, j, k)];
                double DFLX_im1 = f->DFLX[Y(im1, j, k)];
                double EFLX_im1_jp1 = f->EFLX[Y(im1, j + 1, k)];
                if (j == 1) { /* BCy, dyn_UFLX.py:377-384 */
                    DFLX_im1 = 0.;
                    CFLX = 0.;
                    EFLX = 0.;
                }
                if (j == ny) {
                    DFLX_jp1 = 0.;
                    CFLX_jp1 = 0.;
                    EFLX_im1_jp1 = 0.;
                }
                double d = 0.;
                d = d + UVFLX_hor_adv(U[M(i, j, k)], U[M(i - 1, j, k)], U[M(i + 1, j, k)],
                                      U[M(i, j - 1, k)], U[M(i, j + 1, k)],
                                      U[M(i - 1, j - 1, k)], U[M(i - 1, j + 1, k)],
                                      U[M(i + 1, j - 1, k)], U[M(i + 1, j + 1, k)], BFLX,
                                      BFLX_im1, CFLX, CFLX_jp1, DFLX_im1, DFLX_jp1, EFLX,
                                      EFLX_im1_jp1, 1.);
                d = d + ((f->WWIND_UWIND[MS(i, j, k)] - f->WWIND_UWIND[MS(i, j, k + 1)]) /
                         g->dsigma[k]);
                if (g->i_coupling) { /* dyn_UFLX.py:136-170 */
                    const double *PB = f->PHIVB, *RH = f->RHO, *SF = f->SMOMXFLX;
                    const double ALTVB_is =
                        interp_VAR_ds(PB[MS(i, j, k)], PB[MS(i - 1, j, k)], PB[MS(i, j - 1, k)],
                                      PB[MS(i, j + 1, k)], PB[MS(i - 1, j - 1, k)],
                                      PB[MS(i - 1, j + 1, k)], 1, j, ny) / con_g;
                    const double ALTVB_kp1_is =
                        interp_VAR_ds(PB[MS(i, j, k + 1)], PB[MS(i - 1, j, k + 1)],
                                      PB[MS(i, j - 1, k + 1)], PB[MS(i, j + 1, k + 1)],
                                      PB[MS(i - 1, j - 1, k + 1)], PB[MS(i - 1, j + 1, k + 1)], 1, j,
                                      ny) / con_g;
                    const double RHO_is =
                        interp_VAR_ds(RH[M(i, j, k)], RH[M(i - 1, j, k)], RH[M(i, j - 1, k)],
                                      RH[M(i, j + 1, k)], RH[M(i - 1, j - 1, k)],
                                      RH[M(i - 1, j + 1, k)], 1, j, ny);
                    const double SMOMXFLX_is =
                        interp_VAR_ds(SF[M2(i, j)], SF[M2(i - 1, j)], SF[M2(i, j - 1)],
                                      SF[M2(i, j + 1)], SF[M2(i - 1, j - 1)], SF[M2(i - 1, j + 1)],
                                      1, j, ny);
                    const double Kd = f->KMOM_dUWINDdz[MS(i, j, k)],
                                 Kd_kp1 = f->KMOM_dUWINDdz[MS(i, j, k + 1)];
                    double t;
                    if (k == 0)
                        t = ((0. - Kd_kp1) / ((ALTVB_is - ALTVB_kp1_is) * RHO_is));
                    else if (k == nz - 1)
                        t = ((Kd + SMOMXFLX_is) / ((ALTVB_is - ALTVB_kp1_is) * RHO_is));
                    else
                        t = ((Kd - Kd_kp1) / ((ALTVB_is - ALTVB_kp1_is) * RHO_is));
                    f->dUFLXdt_TURB[M(i, j, k)] = t;
                    d = d + t;
                }
                d = d + coriolis_and_spherical_UWIND(
                            f->COLP[M2(i, j)], f->COLP[M2(i - 1, j)], V[Y(i, j, k)],
                            V[Y(i - 1, j, k)], V[Y(i, j + 1, k)], V[Y(i - 1, j + 1, k)],
                            U[M(i, j, k)], U[M(i - 1, j, k)], U[M(i + 1, j, k)],
                            g->corf_is[M2(i, j)], g->lat_is_rad[M2(i, j)], g->dlon_rad[Y2(i, j)],
                            g->dlat_rad[M2(i, j)]);
                d = d + pre_grad(f->PHI[M(i, j, k)], f->PHI[M(i - 1, j, k)], f->COLP[M2(i, j)],
                                 f->COLP[M2(i - 1, j)], f->POTT[M(i, j, k)],
                                 f->POTT[M(i - 1, j, k)], f->PVTF[M(i, j, k)],
                                 f->PVTF[M(i - 1, j, k)], f->PVTFVB[MS(i, j, k)],
                                 f->PVTFVB[MS(i - 1, j, k)], f->PVTFVB[MS(i - 1, j, k + 1)],
                                 f->PVTFVB[MS(i, j, k + 1)], g->dsigma[k], g->sigma_vb[k],
                                 g->sigma_vb[k + 1], g->dyis[M2(i, j)]);
                if (g->UVFLX_dif_coef[k] > 0.)
                    d = d + num_dif(UFLX[M(i, j, k)], UFLX[M(i - 1, j, k)], UFLX[M(i + 1, j, k)],
                                    UFLX[M(i, j - 1, k)], UFLX[M(i, j + 1, k)],
                                    g->UVFLX_dif_coef[k]);
                f->dUFLXdt[M(i, j, k)] = d;
            }
}

/* dyn_VFLX.py:322-408 (launcher) + :67-198 (add_up_tendencies_py).
 * Wall rows j = 1 and j = nys are not computed: the reference evaluates them to NaN
 * (NaN lat_rad/corf halos, never-written RFLX/SFLX rows) and the BC after the Euler step
 * resets VWIND there to 0. */
static void orc_VFLX_tendency(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *U = f->UWIND, *V = f->VWIND, *VFLX = f->VFLX;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 2; j <= ny; j++)
            for (int k = 0; k < nz; k++) {
                double RFLX = f->RFLX[M(i, j, k)];
                double QFLX = f->QFLX[Y(i, j, k)];
                double TFLX = f->TFLX[M(i, j, k)];
                double RFLX_jm1 = f->RFLX[M(i, j - 1, k)];
                double SFLX_jm1 = f->SFLX[M(i, j - 1, k)];
                const int ip1 = (i == nx) ? 1 : i + 1; /* BCx, dyn_VFLX.py:349-356 */
                double QFLX_ip1 = f->QFLX[Y(ip1, j, k)];
                double TFLX_ip1_jm1 = f->TFLX[M(ip1, j - 1, k)];
                double SFLX_ip1 = f->SFLX[M(ip1, j, k)];
                double d = 0.;
                d = d + UVFLX_hor_adv(V[Y(i, j, k)], V[Y(i, j - 1, k)], V[Y(i, j + 1, k)],
                                      V[Y(i - 1, j, k)], V[Y(i + 1, j, k)],
                                      V[Y(i - 1, j - 1, k)], V[Y(i + 1, j - 1, k)],
                                      V[Y(i - 1, j + 1, k)], V[Y(i + 1, j + 1, k)], RFLX,
                                      RFLX_jm1, QFLX, QFLX_ip1, SFLX_jm1, SFLX_ip1, TFLX,
                                      TFLX_ip1_jm1, -1.);
                d = d + ((f->WWIND_VWIND[YS(i, j, k)] - f->WWIND_VWIND[YS(i, j, k + 1)]) /
                         g->dsigma[k]);
                if (g->i_coupling) { /* dyn_VFLX.py:134-166 */
                    const double *PB = f->PHIVB, *RH = f->RHO, *SF = f->SMOMYFLX;
                    const double ALTVB_js =
                        interp_VAR_ds(PB[MS(i, j, k)], PB[MS(i, j - 1, k)], PB[MS(i - 1, j, k)],
                                      PB[MS(i + 1, j, k)], PB[MS(i - 1, j - 1, k)],
                                      PB[MS(i + 1, j - 1, k)], 0, i, nx) / con_g;
                    const double ALTVB_kp1_js =
                        interp_VAR_ds(PB[MS(i, j, k + 1)], PB[MS(i, j - 1, k + 1)],
                                      PB[MS(i - 1, j, k + 1)], PB[MS(i + 1, j, k + 1)],
                                      PB[MS(i - 1, j - 1, k + 1)], PB[MS(i + 1, j - 1, k + 1)], 0, i,
                                      nx) / con_g;
                    const double RHO_js =
                        interp_VAR_ds(RH[M(i, j, k)], RH[M(i, j - 1, k)], RH[M(i - 1, j, k)],
                                      RH[M(i + 1, j, k)], RH[M(i - 1, j - 1, k)],
                                      RH[M(i + 1, j - 1, k)], 0, i, nx);
                    const double SMOMYFLX_js =
                        interp_VAR_ds(SF[M2(i, j)], SF[M2(i, j - 1)], SF[M2(i - 1, j)],
                                      SF[M2(i + 1, j)], SF[M2(i - 1, j - 1)], SF[M2(i + 1, j - 1)],
                                      0, i, nx);
                    const double Kd = f->KMOM_dVWINDdz[YS(i, j, k)],
                                 Kd_kp1 = f->KMOM_dVWINDdz[YS(i, j, k + 1)];
                    double t;
                    if (k == 0)
                        t = ((0. - Kd_kp1) / ((ALTVB_js - ALTVB_kp1_js) * RHO_js));
                    else if (k == nz - 1)
                        t = ((Kd + SMOMYFLX_js) / ((ALTVB_js - ALTVB_kp1_js) * RHO_js));
                    else
                        t = ((Kd - Kd_kp1) / ((ALTVB_js - ALTVB_kp1_js) * RHO_js));
                    f->dVFLXdt_TURB[Y(i, j, k)] = t;
                    d = d + t;
                }
                d = d + coriolis_and_spherical_VWIND(
                            f->COLP[M2(i, j)], f->COLP[M2(i, j - 1)], U[M(i, j, k)],
                            U[M(i, j - 1, k)], U[M(i + 1, j, k)], U[M(i + 1, j - 1, k)],
                            g->corf[M2(i, j)], g->corf[M2(i, j - 1)], g->lat_rad[M2(i, j)],
                            g->lat_rad[M2(i, j - 1)], g->dlon_rad[Y2(i, j)],
                            g->dlat_rad[M2(i, j)]);
                d = d + pre_grad(f->PHI[M(i, j, k)], f->PHI[M(i, j - 1, k)], f->COLP[M2(i, j)],
                                 f->COLP[M2(i, j - 1)], f->POTT[M(i, j, k)],
                                 f->POTT[M(i, j - 1, k)], f->PVTF[M(i, j, k)],
                                 f->PVTF[M(i, j - 1, k)], f->PVTFVB[MS(i, j, k)],
                                 f->PVTFVB[MS(i, j - 1, k)], f->PVTFVB[MS(i, j - 1, k + 1)],
                                 f->PVTFVB[MS(i, j, k + 1)], g->dsigma[k], g->sigma_vb[k],
                                 g->sigma_vb[k + 1], g->dxjs[Y2(i, j)]);
                if (g->UVFLX_dif_coef[k] > 0.)
                    d = d + num_dif(VFLX[Y(i, j, k)], VFLX[Y(i - 1, j, k)], VFLX[Y(i + 1, j, k)],
                                    VFLX[Y(i, j - 1, k)], VFLX[Y(i, j + 1, k)],
                                    g->UVFLX_dif_coef[k]);
                f->dVFLXdt[Y(i, j, k)] = d;
            }
}

/* dyn_org_discretizations.py:121-249 (CPU branch; the KMOM/SMOM BCs act on zero fields) */
void orc_momentum(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    if (g->i_coupling) orc_exchange_BC(g, f->KMOM, nx + 2, ny + 2, nz + 1);
    orc_UVFLX_prep_adv(g, f);
    if (g->i_coupling) {
        orc_exchange_BC(g, f->KMOM_dUWINDdz, nx + 3, ny + 2, nz + 1);
        orc_exchange_BC(g, f->KMOM_dVWINDdz, nx + 2, ny + 3, nz + 1);
        orc_exchange_BC(g, f->SMOMXFLX, nx + 2, ny + 2, 1);
        orc_exchange_BC(g, f->SMOMYFLX, nx + 2, ny + 2, 1);
    }
    orc_UFLX_tendency(g, f);
    orc_VFLX_tendency(g, f);
}

/* dyn_functions.py:105-114 (hor_adv_py) */
static inline double hor_adv(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1,
                             double VAR_jp1, double UFLX, double UFLX_ip1, double VFLX,
                             double VFLX_jp1, double A)
{
    return ((+UFLX * (VAR_im1 + VAR) / 2. - UFLX_ip1 * (VAR + VAR_ip1) / 2.
             + VFLX * (VAR_jm1 + VAR) / 2. - VFLX_jp1 * (VAR + VAR_jp1) / 2.) / A);
}

/* dyn_functions.py:118-136 (vert_adv_py); the k == nz branch is unreachable */
static inline double vert_adv(double VARVB, double VARVB_kp1, double WWIND, double WWIND_kp1,
                              double COLP_NEW, double dsigma, int k)
{
    if (k == 0) return COLP_NEW * (-WWIND_kp1 * VARVB_kp1) / dsigma;
    return COLP_NEW * (+WWIND * VARVB - WWIND_kp1 * VARVB_kp1) / dsigma;
}

/* dyn_functions.py:142-155 (num_dif_pw_py) */
static inline double num_dif_pw(double VAR, double VAR_im1, double VAR_ip1, double VAR_jm1,
                                double VAR_jp1, double COLP, double COLP_im1, double COLP_ip1,
                                double COLP_jm1, double COLP_jp1, double VAR_dif_coef)
{
    return VAR_dif_coef * (+COLP_im1 * VAR_im1 + COLP_ip1 * VAR_ip1 + COLP_jm1 * VAR_jm1 +
                           COLP_jp1 * VAR_jp1 - 4. * COLP * VAR);
}

/* dyn_POTT.py:180-216 (launcher) + :55-110 (add_up_tendencies_py) */
void orc_temperature(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *P = f->POTT, *C = f->COLP;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++)
            for (int k = 0; k < nz; k++) {
                double d = 0.;
                d = d + hor_adv(P[M(i, j, k)], P[M(i - 1, j, k)], P[M(i + 1, j, k)],
                                P[M(i, j - 1, k)], P[M(i, j + 1, k)], f->UFLX[M(i, j, k)],
                                f->UFLX[M(i + 1, j, k)], f->VFLX[Y(i, j, k)],
                                f->VFLX[Y(i, j + 1, k)], g->A[M2(i, j)]);
                d = d + vert_adv(f->POTTVB[MS(i, j, k)], f->POTTVB[MS(i, j, k + 1)],
                                 f->WWIND[MS(i, j, k)], f->WWIND[MS(i, j, k + 1)],
                                 f->COLP_NEW[M2(i, j)], g->dsigma[k], k);
                if (g->i_coupling) { /* dyn_POTT.py:87-96 */
                    const int km = k > 0 ? k - 1 : k, kp = k < nz - 1 ? k + 1 : k;
                    const double t = turb_flux_tendency(
                        f->PHI[M(i, j, k)], f->PHI[M(i, j, kp)], f->PHI[M(i, j, km)],
                        f->PHIVB[MS(i, j, k)], f->PHIVB[MS(i, j, k + 1)], P[M(i, j, k)],
                        P[M(i, j, kp)], P[M(i, j, km)], f->KHEAT[MS(i, j, k)],
                        f->KHEAT[MS(i, j, k + 1)], f->RHO[M(i, j, k)], f->RHOVB[MS(i, j, k)],
                        f->RHOVB[MS(i, j, k + 1)], C[M2(i, j)], f->SSHFLX[M2(i, j)] / con_cp, k,
                        nz);
                    d = d + t;
                    f->dPOTTdt_TURB[M(i, j, k)] = t / C[M2(i, j)] * 3600.;
                }
                if (g->POTT_dif_coef[k] > 0.)
                    d = d + num_dif_pw(P[M(i, j, k)], P[M(i - 1, j, k)], P[M(i + 1, j, k)],
                                       P[M(i, j - 1, k)], P[M(i, j + 1, k)], C[M2(i, j)],
                                       C[M2(i - 1, j)], C[M2(i + 1, j)], C[M2(i, j - 1)],
                                       C[M2(i, j + 1)], g->POTT_dif_coef[k]);
                if (g->i_coupling) /* dyn_POTT.py:40-41, :107-108 (radiation_py) */
                    d = d + (f->dPOTTdt_RAD[M(i, j, k)] * C[M2(i, j)]);
                f->dPOTTdt[M(i, j, k)] = d;
            }
}

/* dyn_functions.py:70-95 (comp_VARVB_log_py) */
static inline double comp_VARVB_log(double VAR, double VAR_km1)
{
    const double min_val = 0.0000001;
    VAR = fmax(VAR, min_val);
    VAR_km1 = fmax(VAR_km1, min_val);
    if (VAR_km1 == VAR) return VAR;
    return ((log(VAR_km1) - log(VAR)) / (1. / VAR - 1. / VAR_km1));
}

/* dyn_moist.py:201-243 (launcher) + :49-129 (add_up_tendencies_py).
 * The reference reads QV[k-1] at k = 0 (numba wraps to level nz-1) and QV[k+1] at
 * k = nz-1 (one past the column); both only enter products that vert_adv drops (k = 0)
 * or multiplies by WWIND[nz] = 0, so the column's own value is passed instead. */
void orc_moisture(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *C = f->COLP;
    if (!g->i_moist) return;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++)
            for (int k = 0; k < nz; k++) {
                for (int t = 0; t < 2; t++) {
                    const double *Q = t ? f->QC : f->QV;
                    double *dQ = t ? f->dQCdt : f->dQVdt;
                    double Qk = Q[M(i, j, k)];
                    double Q_km1 = (k > 0) ? Q[M(i, j, k - 1)] : Qk;
                    double Q_kp1 = (k < nz - 1) ? Q[M(i, j, k + 1)] : Qk;
                    double d = 0.;
                    d = d + hor_adv(Qk, Q[M(i - 1, j, k)], Q[M(i + 1, j, k)], Q[M(i, j - 1, k)],
                                    Q[M(i, j + 1, k)], f->UFLX[M(i, j, k)],
                                    f->UFLX[M(i + 1, j, k)], f->VFLX[Y(i, j, k)],
                                    f->VFLX[Y(i, j + 1, k)], g->A[M2(i, j)]);
                    double QVB = comp_VARVB_log(Qk, Q_km1);
                    double QVB_kp1 = comp_VARVB_log(Q_kp1, Qk);
                    d = d + vert_adv(QVB, QVB_kp1, f->WWIND[MS(i, j, k)],
                                     f->WWIND[MS(i, j, k + 1)], f->COLP_NEW[M2(i, j)],
                                     g->dsigma[k], k);
                    if (g->i_coupling) { /* dyn_moist.py:100-112 */
                        const int km = k > 0 ? k - 1 : k, kp = k < nz - 1 ? k + 1 : k;
                        const double tq = turb_flux_tendency(
                            f->PHI[M(i, j, k)], f->PHI[M(i, j, kp)], f->PHI[M(i, j, km)],
                            f->PHIVB[MS(i, j, k)], f->PHIVB[MS(i, j, k + 1)], Qk, Q[M(i, j, kp)],
                            Q[M(i, j, km)], f->KHEAT[MS(i, j, k)], f->KHEAT[MS(i, j, k + 1)],
                            f->RHO[M(i, j, k)], f->RHOVB[MS(i, j, k)], f->RHOVB[MS(i, j, k + 1)],
                            C[M2(i, j)], t ? 0. : f->SLHFLX[M2(i, j)] / con_Lh, k, nz);
                        d = d + tq;
                        if (!t) f->dQVdt_TURB[M(i, j, k)] = tq;
                    }
                    if (g->moist_dif_coef[k] > 0.)
                        d = d + num_dif_pw(Qk, Q[M(i - 1, j, k)], Q[M(i + 1, j, k)],
                                           Q[M(i, j - 1, k)], Q[M(i, j + 1, k)], C[M2(i, j)],
                                           C[M2(i - 1, j)], C[M2(i + 1, j)], C[M2(i, j - 1)],
                                           C[M2(i, j + 1)], g->moist_dif_coef[k]);
                    dQ[M(i, j, k)] = d;
                }
            }
}

/* dyn_tendencies.py:25-72 */
void orc_compute_tendencies(const orc_grid *g, orc_fields *f)
{
    orc_continuity(g, f);
    orc_momentum(g, f);
    orc_temperature(g, f);
    orc_moisture(g, f);
}

/* dyn_timestep.py:34-79 */
static inline double euler_forward_pw(double VAR, double dVARdt, double COLP, double COLP_OLD,
                                      double dt)
{
    return VAR * COLP_OLD / COLP + dt * dVARdt / COLP;
}
static inline double interp_COLPA_js(double COLP, double COLP_jm1, double COLP_im1,
                                     double COLP_ip1, double COLP_jm1_ip1, double COLP_jm1_im1,
                                     double A, double A_jm1, double A_im1, double A_ip1,
                                     double A_jm1_ip1, double A_jm1_im1)
{
    return 1. / 8. * (COLP_jm1_ip1 * A_jm1_ip1 + COLP_ip1 * A_ip1 + 2. * COLP_jm1 * A_jm1 +
                      2. * COLP * A + COLP_jm1_im1 * A_jm1_im1 + COLP_im1 * A_im1);
}
static inline double interp_COLPA_is(double COLP, double COLP_im1, double COLP_jm1,
                                     double COLP_jp1, double COLP_im1_jp1, double COLP_im1_jm1,
                                     double A, double A_im1, double A_jm1, double A_jp1,
                                     double A_im1_jp1, double A_im1_jm1, int j, int ny)
{
    if (j == 1)
        return 1. / 4. * (COLP_im1_jp1 * A_im1_jp1 + COLP_jp1 * A_jp1 + COLP_im1 * A_im1 +
                          COLP * A);
    else if (j == ny)
        return 1. / 4. * (COLP_im1_jm1 * A_im1_jm1 + COLP_jm1 * A_jm1 + COLP_im1 * A_im1 +
                          COLP * A);
    return 1. / 8. * (COLP_im1_jp1 * A_im1_jp1 + COLP_jp1 * A_jp1 + 2. * COLP_im1 * A_im1 +
                      2. * COLP * A + COLP_im1_jm1 * A_im1_jm1 + COLP_jm1 * A_jm1);
}

/* dyn_timestep.py:212-296 (make_timestep_cpu) + BCs dyn_org_discretizations.py:388-393.
 * The reference loops over i in [1,nxs], j in [1,nys] for every variable; the entries
 * outside each variable's own interior are halo cells that the BCs below overwrite, so
 * only the interiors are computed here (UWIND: i<=nx; VWIND: 2<=j<=ny; mass: i<=nx,j<=ny). */
void orc_euler_forward(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double *C = f->COLP, *CO = f->COLP_OLD, *A = g->A;
    const double dt = g->dt;
#pragma omp parallel for schedule(static)
    for (int i = 1; i <= nx; i++)
        for (int j = 1; j <= ny; j++) {
            double COLPA_is = interp_COLPA_is(
                C[M2(i, j)], C[M2(i - 1, j)], C[M2(i, j - 1)], C[M2(i, j + 1)],
                C[M2(i - 1, j + 1)], C[M2(i - 1, j - 1)], A[M2(i, j)], A[M2(i - 1, j)],
                A[M2(i, j - 1)], A[M2(i, j + 1)], A[M2(i - 1, j + 1)], A[M2(i - 1, j - 1)], j, ny);
            double COLPA_OLD_is = interp_COLPA_is(
                CO[M2(i, j)], CO[M2(i - 1, j)], CO[M2(i, j - 1)], CO[M2(i, j + 1)],
                CO[M2(i - 1, j + 1)], CO[M2(i - 1, j - 1)], A[M2(i, j)], A[M2(i - 1, j)],
                A[M2(i, j - 1)], A[M2(i, j + 1)], A[M2(i - 1, j + 1)], A[M2(i - 1, j - 1)], j, ny);
            double COLPA_js = interp_COLPA_js(
                C[M2(i, j)], C[M2(i, j - 1)], C[M2(i - 1, j)], C[M2(i + 1, j)],
                C[M2(i + 1, j - 1)], C[M2(i - 1, j - 1)], A[M2(i, j)], A[M2(i, j - 1)],
                A[M2(i - 1, j)], A[M2(i + 1, j)], A[M2(i + 1, j - 1)], A[M2(i - 1, j - 1)]);
            double COLPA_OLD_js = interp_COLPA_js(
                CO[M2(i, j)], CO[M2(i, j - 1)], CO[M2(i - 1, j)], CO[M2(i + 1, j)],
                CO[M2(i + 1, j - 1)], CO[M2(i - 1, j - 1)], A[M2(i, j)], A[M2(i, j - 1)],
                A[M2(i - 1, j)], A[M2(i + 1, j)], A[M2(i + 1, j - 1)], A[M2(i - 1, j - 1)]);
            for (int k = 0; k < nz; k++) {
                f->UWIND[M(i, j, k)] = euler_forward_pw(f->UWIND_OLD[M(i, j, k)],
                                                        f->dUFLXdt[M(i, j, k)], COLPA_is,
                                                        COLPA_OLD_is, dt);
                if (j >= 2)
                    f->VWIND[Y(i, j, k)] = euler_forward_pw(f->VWIND_OLD[Y(i, j, k)],
                                                            f->dVFLXdt[Y(i, j, k)], COLPA_js,
                                                            COLPA_OLD_js, dt);
                f->POTT[M(i, j, k)] = euler_forward_pw(f->POTT_OLD[M(i, j, k)],
                                                       f->dPOTTdt[M(i, j, k)], C[M2(i, j)],
                                                       CO[M2(i, j)], dt);
                if (g->i_moist) {
                    f->QV[M(i, j, k)] = euler_forward_pw(f->QV_OLD[M(i, j, k)],
                                                         f->dQVdt[M(i, j, k)], C[M2(i, j)],
                                                         CO[M2(i, j)], dt);
                    f->QC[M(i, j, k)] = euler_forward_pw(f->QC_OLD[M(i, j, k)],
                                                         f->dQCdt[M(i, j, k)], C[M2(i, j)],
                                                         CO[M2(i, j)], dt);
                }
            }
        }
    orc_exchange_BC(g, f->POTT, nx + 2, ny + 2, nz);
    orc_exchange_BC(g, f->VWIND, nx + 2, ny + 3, nz);
    orc_exchange_BC(g, f->UWIND, nx + 3, ny + 2, nz);
    if (g->i_moist) {
        orc_exchange_BC(g, f->QV, nx + 2, ny + 2, nz);
        orc_exchange_BC(g, f->QC, nx + 2, ny + 2, nz);
    }
}

/* dyn_diagnostics.py:139-195 (diag_PVTF_cpu, diag_PHI_cpu, diag_POTTVB_cpu):
 * all columns including the halos */
void orc_primary_diag(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double con_kappa = CON_KAPPA;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; i++)
        for (int j = 0; j < ny + 2; j++) {
            const double COLP = f->COLP[M2(i, j)];
            for (int k = 0; k < nz; k++) {
                double pairvb_km12 = g->pair_top + g->sigma_vb[k] * COLP;
                double pairvb_kp12 = g->pair_top + g->sigma_vb[k + 1] * COLP;
                f->PVTF[M(i, j, k)] =
                    1. / (1. + con_kappa) *
                    (pow(pairvb_kp12 / 100000., con_kappa) * pairvb_kp12 -
                     pow(pairvb_km12 / 100000., con_kappa) * pairvb_km12) /
                    (pairvb_kp12 - pairvb_km12);
                f->PVTFVB[MS(i, j, k)] = pow(pairvb_km12 / 100000., con_kappa);
                if (k == nz - 1)
                    f->PVTFVB[MS(i, j, k + 1)] = pow(pairvb_kp12 / 100000., con_kappa);
            }
            /* diag_PHI_cpu */
            f->PHIVB[MS(i, j, nz)] = f->HSURF[M2(i, j)] * con_g;
            for (int k = nz - 1; k >= 0; k--) {
                f->PHI[M(i, j, k)] =
                    f->PHIVB[MS(i, j, k + 1)] -
                    con_cp * (f->POTT[M(i, j, k)] *
                              (f->PVTF[M(i, j, k)] - f->PVTFVB[MS(i, j, k + 1)]));
                f->PHIVB[MS(i, j, k)] =
                    f->PHI[M(i, j, k)] -
                    con_cp * (f->POTT[M(i, j, k)] *
                              (f->PVTFVB[MS(i, j, k)] - f->PVTF[M(i, j, k)]));
            }
            /* diag_POTTVB_cpu */
            for (int k = 1; k < nz; k++) {
                f->POTTVB[MS(i, j, k)] =
                    (+(f->PVTFVB[MS(i, j, k)] - f->PVTF[M(i, j, k - 1)]) * f->POTT[M(i, j, k - 1)] +
                     (f->PVTF[M(i, j, k)] - f->PVTFVB[MS(i, j, k)]) * f->POTT[M(i, j, k)]) /
                    (f->PVTF[M(i, j, k)] - f->PVTF[M(i, j, k - 1)]);
                if (k == 1)
                    f->POTTVB[MS(i, j, k - 1)] =
                        f->POTT[M(i, j, k - 1)] -
                        (f->POTTVB[MS(i, j, k)] - f->POTT[M(i, j, k - 1)]);
                else if (k == nz - 1)
                    f->POTTVB[MS(i, j, k + 1)] =
                        f->POTT[M(i, j, k)] - (f->POTTVB[MS(i, j, k)] - f->POTT[M(i, j, k)]);
            }
        }
}

/* dyn_diagnostics.py:199-222 (diag_secondary_cpu): all columns including the halos */
void orc_secondary_diag(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double con_kappa = CON_KAPPA;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; i++)
        for (int j = 0; j < ny + 2; j++) {
            for (int k = 0; k <= nz; k++) {
                f->PAIRVB[MS(i, j, k)] = 100000. * pow(f->PVTFVB[MS(i, j, k)], 1. / con_kappa);
                f->TAIRVB[MS(i, j, k)] = f->POTTVB[MS(i, j, k)] * f->PVTFVB[MS(i, j, k)];
                f->RHOVB[MS(i, j, k)] = f->PAIRVB[MS(i, j, k)] / (con_Rd * f->TAIRVB[MS(i, j, k)]);
            }
            for (int k = 0; k < nz; k++) {
                f->TAIR[M(i, j, k)] = f->POTT[M(i, j, k)] * f->PVTF[M(i, j, k)];
                f->PAIR[M(i, j, k)] = 100000. * pow(f->PVTF[M(i, j, k)], 1. / con_kappa);
                f->RHO[M(i, j, k)] = f->PAIR[M(i, j, k)] / (con_Rd * f->TAIR[M(i, j, k)]);
                double wx = (f->UWIND[M(i, j, k)] + f->UWIND[M(i + 1, j, k)]) / 2.;
                double wy = (f->VWIND[Y(i, j, k)] + f->VWIND[Y(i, j + 1, k)]) / 2.;
                f->WINDX[M(i, j, k)] = wx;
                f->WINDY[M(i, j, k)] = wy;
                f->WIND[M(i, j, k)] = sqrt(wx * wx + wy * wy); /* x**2., x**(1/2): exact pow cases */
            }
        }
}

/* turb_compute.py:190-204 (launch_numba_cpu) + :53-145 (bulk_richardson_py, compute_K_coefs_py,
 * run_all_py) + misc_meteo_utilities.py:36-49 (calc_virtual_temperature_py): the reference's
 * turbulence module, KMOM / KHEAT on the interior interfaces of every column incl. the halo.
 * Mind the launcher's argument mapping: "PHI_k" = PHIVB[k], "POTT_k" = POTTVB[k], km05 = the
 * full level above the interface (k-1), kp05 = the full level below it (k). */
void orc_compute_turbulence(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const double Ri_c = 1.0, free_mix_len = 200., con_k = 0.35, con_Pr = 0.72;
    const double min_wind_diff = 0.0001, min_KMOM = 0.000001, max_KMOM = 0.01;
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nx + 2; i++)
        for (int j = 0; j < ny + 2; j++)
            for (int k = 1; k < nz; k++) {
                double WINDX_km05 = f->WINDX[M(i, j, k - 1)], WINDX_kp05 = f->WINDX[M(i, j, k)];
                double WINDY_km05 = f->WINDY[M(i, j, k - 1)], WINDY_kp05 = f->WINDY[M(i, j, k)];
                if (WINDX_km05 == WINDX_kp05) WINDX_km05 += min_wind_diff;
                if (WINDY_km05 == WINDY_kp05) WINDY_km05 += min_wind_diff;
                const double ALT_k = f->PHIVB[MS(i, j, k)] / con_g;
                const double ALT_km05 = f->PHI[M(i, j, k - 1)] / con_g;
                const double ALT_kp05 = f->PHI[M(i, j, k)] / con_g;
                const double HGT_k = ALT_k - f->HSURF[M2(i, j)];
                const double mix_len = con_k * HGT_k / (1. + con_k * HGT_k / free_mix_len);
                const double QV_k = comp_VARVB_log(f->QV[M(i, j, k)], f->QV[M(i, j, k - 1)]);
                const double POTT_v_k = f->POTTVB[MS(i, j, k)] * (1. + QV_k / 0.622) / (1. + QV_k);
                const double dx = WINDX_km05 - WINDX_kp05, dy = WINDY_km05 - WINDY_kp05;
                const double dalt = ALT_km05 - ALT_kp05;
                const double Ri_b_k =
                    ((con_g / POTT_v_k * (f->POTT[M(i, j, k - 1)] - f->POTT[M(i, j, k)]) * dalt) /
                     (dx * dx + dy * dy));
                const double sx = dx / dalt, sy = dy / dalt;
                const double shear_term = sqrt(sx * sx + sy * sy);
                double KMOM_k = mix_len * mix_len * shear_term * (Ri_c - Ri_b_k) / Ri_c;
                if (KMOM_k < min_KMOM) KMOM_k = min_KMOM;
                if (KMOM_k > max_KMOM) KMOM_k = max_KMOM;
                f->KMOM[MS(i, j, k)] = KMOM_k;
                f->KHEAT[MS(i, j, k)] = KMOM_k / con_Pr;
            }
}

/* dyn_matsuno.py:28-129 (step_matsuno, i_comp_mode == 1) */
void orc_step_matsuno(const orc_grid *g, orc_fields *f)
{
    const int nx = g->nx, ny = g->ny, nz = g->nz;
    const size_t n2 = (size_t)(nx + 2) * (ny + 2) * sizeof(double);
    memcpy(f->COLP_OLD, f->COLP, n2);
    memcpy(f->UWIND_OLD, f->UWIND, (size_t)(nx + 3) * (ny + 2) * nz * sizeof(double));
    memcpy(f->VWIND_OLD, f->VWIND, (size_t)(nx + 2) * (ny + 3) * nz * sizeof(double));
    memcpy(f->POTT_OLD, f->POTT, n2 * nz);
    if (g->i_moist) {
        memcpy(f->QV_OLD, f->QV, n2 * nz);
        memcpy(f->QC_OLD, f->QC, n2 * nz);
    }
    for (int stage = 0; stage < 2; stage++) { /* estimate, final */
        orc_compute_tendencies(g, f);
        memcpy(f->COLP, f->COLP_NEW, n2);
        orc_euler_forward(g, f);
        orc_primary_diag(g, f);
    }
}
