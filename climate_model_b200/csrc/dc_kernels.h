// dc_kernels.h -- per-column kernel bodies of the dynamical core ("v1": one body per
// reference kernel, boundary exchange fused into the producer as image stores).
//
// A body is a functor called once per (i, j) column (global reference indices); it marches
// the sigma column in registers.  With longitude fastest in memory, consecutive threads
// (consecutive i) touch consecutive addresses at every level: every load/store coalesces.
// dyncore.cu wraps each body in a __global__ kernel; tests/emu/ runs the same bodies on
// the host to check the index arithmetic without a GPU (test infrastructure only).
#pragma once
#include "dc_geom.h"
#include "dc_point.h"

namespace dc {

// ---------------------------------------------------------------------------------------
// image stores: write v at (i, j, k) and at every halo cell misc_boundaries.py:22-42
// (exchange_BC_cpu) would copy it to.  Rows are images only on the GLOBAL walls.
// ---------------------------------------------------------------------------------------
DC_HD void put_rows(const Geom &g, double *F, int i, int j, int k, double v)
{
    F[g.idx(i, j, k)] = v;
    if (j == 1) F[g.idx(i, 0, k)] = v;            // FIELD[:,0,:] = FIELD[:,1,:]
    if (j == g.ny) F[g.idx(i, g.ny + 1, k)] = v;  // FIELD[:,ny+1,:] = FIELD[:,ny,:]
}
// unstaggered in x, unstaggered in y  (nx+2, ny+2)
DC_HD void put_mass(const Geom &g, double *F, int i, int j, int k, double v)
{
    put_rows(g, F, i, j, k, v);
    if (i == g.nx) put_rows(g, F, 0, j, k, v);     // FIELD[0] = FIELD[nx]
    if (i == 1) put_rows(g, F, g.nx + 1, j, k, v);  // FIELD[nx+1] = FIELD[1]
}
// staggered in x  (nx+3, ny+2); the body computes i in [1, nx]
DC_HD void put_xstag(const Geom &g, double *F, int i, int j, int k, double v)
{
    put_rows(g, F, i, j, k, v);
    if (i == g.nx) put_rows(g, F, 0, j, k, v);      // FIELD[0] = FIELD[nxs-1]
    if (i == 1) put_rows(g, F, g.nx + 1, j, k, v);  // FIELD[nxs] = FIELD[1]
    if (i == 2) put_rows(g, F, g.nx + 2, j, k, v);  // FIELD[nxs+1] = FIELD[2]
}
// staggered in y  (nx+2, ny+3): rows 0, 1, nys, nys+1 are forced to 0.
// Call for j in [1, nys]; the value is ignored on the wall rows.
DC_HD void put_ystag(const Geom &g, double *F, int i, int j, int k, double v)
{
    const int nys = g.ny + 1;
    const bool wall = (j == 1) || (j == nys);
    if (wall) v = 0.;
    const int jx = (j == 1) ? 0 : nys + 1;  // second zero row beside a wall row
    F[g.idx(i, j, k)] = v;
    if (wall) F[g.idx(i, jx, k)] = v;
    if (i == g.nx) {
        F[g.idx(0, j, k)] = v;
        if (wall) F[g.idx(0, jx, k)] = v;
    }
    if (i == 1) {
        F[g.idx(g.nx + 1, j, k)] = v;
        if (wall) F[g.idx(g.nx + 1, jx, k)] = v;
    }
}

// ---------------------------------------------------------------------------------------
// generic exchange_BC on one field (misc_boundaries.py:22-42), one thread per (i, j) of the
// SOURCE cells; used by the dc_exchange_bc entry (initial conditions, external callers).
// kind: 0 mass, 1 x-staggered, 2 y-staggered, 3 xy-staggered
// ---------------------------------------------------------------------------------------
struct ExchangeBCBody {
    Geom g;
    double *F;
    int kind, nk;
    DC_HD void operator()(int i, int j) const
    {
        // threads cover i in [1, nx(+1)], j in [1, ny(+1)]
        const bool xs = kind & 1, ys = kind & 2;
        // only the cells next to a boundary have images (or are wall rows): every other
        // thread leaves without touching memory
        if (i > 2 && i != g.nx && j != 1 && j < g.ny) return;
        for (int k = 0; k < nk; k++) {
            double v = F[g.idx(i, j, k)];
            if (ys) {
                if (xs) {  // CFLX/QFLX-like: x images of put_xstag, rows of put_ystag
                    const int nys = g.ny + 1;
                    const bool wall = (j == 1) || (j == nys);
                    if (wall) v = 0.;
                    const int jx = (j == 1) ? 0 : nys + 1;
                    int ii[2] = {i, -1};
                    if (i == g.nx) ii[1] = 0;
                    if (i == 1) ii[1] = g.nx + 1;
                    if (i == 2) ii[1] = g.nx + 2;
                    for (int a = 0; a < 2; a++) {
                        if (ii[a] < 0) continue;
                        F[g.idx(ii[a], j, k)] = v;
                        if (wall) F[g.idx(ii[a], jx, k)] = v;
                    }
                } else {
                    put_ystag(g, F, i, j, k, v);
                }
            } else if (xs) {
                put_xstag(g, F, i, j, k, v);
            } else {
                put_mass(g, F, i, j, k, v);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// latitude-band halo rows <-> contiguous message buffers [field][k][2][NI], both directions
// and all fields in one launch
// ---------------------------------------------------------------------------------------
struct HaloPackBody {
    static constexpr int MAXF = 6;
    Geom g;
    double *F[MAXF];   // fields (nk[m] planes each), message segments in this order
    int nk[MAXF];
    int nf;
    double *south, *north;   // message buffers (NULL: that side is a domain wall)
    int j_south, j_north;    // first of the two rows on each side
    int to_buf;
    // threads: i in [0, NI-1], jj in [0, 2*max(nk)-1] with jj = 2*k + r
    DC_HD void operator()(int i, int jj) const
    {
        const int r = jj & 1, k = jj >> 1;
        size_t off = 0;
        for (int m = 0; m < nf; m++) {
            if (k < nk[m]) {
                const size_t b = off + ((size_t)k * 2 + r) * (size_t)g.NI + i;
                if (south) {
                    const size_t a = g.idx(i, j_south + r, k);
                    if (to_buf) south[b] = F[m][a]; else F[m][a] = south[b];
                }
                if (north) {
                    const size_t a = g.idx(i, j_north + r, k);
                    if (to_buf) north[b] = F[m][a]; else F[m][a] = north[b];
                }
            }
            off += (size_t)nk[m] * 2 * g.NI;
        }
    }
};

// ---------------------------------------------------------------------------------------
// continuity: dyn_continuity.py:170-228 + BCs dyn_org_discretizations.py:114-117
// threads: i in [1, nx], j in [1, ny].  U and V are read ONCE: the running flux-divergence
// prefix is parked in WWIND[k] during the first sweep and finalised in the second (which
// re-reads only this thread's own WWIND column).
// ---------------------------------------------------------------------------------------
// MODE bit 0: store FLXDIV; bit 1: store UFLX / VFLX (with their boundary images)
template <int MODE>
struct ContinuityBody {
    Geom g;
    const double *UWIND, *VWIND, *COLP, *COLP_OLD;
    double *UFLX, *VFLX, *FLXDIV, *WWIND, *COLP_NEW, *dCOLPdt;
    DC_HD void operator()(int i, int j) const
    {
        const int nz = g.nz;
        const double c = COLP[g.idx2(i, j)];
        const double c_im1 = COLP[g.idx2(i - 1, j)], c_ip1 = COLP[g.idx2(i + 1, j)];
        const double c_jm1 = COLP[g.idx2(i, j - 1)], c_jp1 = COLP[g.idx2(i, j + 1)];
        const double dxjs = g.dxjs[g.row(j)], dxjs_jp1 = g.dxjs[g.row(j + 1)];
        const Div A = mkdiv(g.A[g.row(j)], g.r_A[g.row(j)]);
        double s = 0.;  // sequential ascending sum (numba's FLXDIV.sum(axis=2))
        for (int k = 0; k < nz; k++) {
            const double uf = calc_UFLX(UWIND[g.idx(i, j, k)], c, c_im1, g.dyis);
            const double uf_ip1 = calc_UFLX(UWIND[g.idx(i + 1, j, k)], c_ip1, c, g.dyis);
            const double vf = calc_VFLX(VWIND[g.idx(i, j, k)], c, c_jm1, dxjs);
            const double vf_jp1 = calc_VFLX(VWIND[g.idx(i, j + 1, k)], c_jp1, c, dxjs_jp1);
            const double fd = calc_FLXDIV(uf, uf_ip1, vf, vf_jp1, g.dsigma[k], A);
            if (MODE & 2) {
                put_xstag(g, UFLX, i, j, k, uf);
                put_ystag(g, VFLX, i, j, k, vf);
                if (j == g.ny) put_ystag(g, VFLX, i, g.ny + 1, k, 0.);
            }
            if (MODE & 1) FLXDIV[g.idx(i, j, k)] = fd;
            s += fd;
            if (k + 1 < nz) WWIND[g.idx(i, j, k + 1)] = s;  // prefix, finalised below
        }
        const double dcdt = -s;
        const double cnew = COLP_OLD[g.idx2(i, j)] + g.dt * dcdt;
        dCOLPdt[g.idx2(i, j)] = dcdt;
        put_mass(g, COLP_NEW, i, j, 0, cnew);
        const Div cn = mkdiv(cnew);
        for (int k = 1; k < nz; k++) {
            const double flxdivsum = WWIND[g.idx(i, j, k)];
            put_mass(g, WWIND, i, j, k, (-flxdivsum / cn - g.sigma_vb[k] * dcdt / cn));
        }
    }
};

// ---------------------------------------------------------------------------------------
// momentum-flux preparation: dyn_UVFLX_prepare.py:249-439 (turbulence part: zero fields)
// threads: i in [1, nxs], j in [1, nys]
// ---------------------------------------------------------------------------------------
struct PrepBody {
    Geom g;
    const double *UWIND, *VWIND, *WWIND, *UFLX, *VFLX, *COLP_NEW;
    double *WWIND_UWIND, *WWIND_VWIND, *BFLX, *CFLX, *DFLX, *EFLX, *RFLX, *QFLX, *SFLX, *TFLX;
    // physics coupling (g.i_coupling): dyn_UVFLX_prepare.py:298-343
    const double *KMOM, *RHOVB, *PHI, *COLP;
    double *KMOM_dUWINDdz, *KMOM_dVWINDdz;
    // the six points around a u-point (rigid = true: d = i, p = j) or a v-point
    DC_HD Six six3(const double *F, int i, int j, int k, bool u) const
    {
        if (u)
            return Six{F[g.idx(i, j, k)],         F[g.idx(i - 1, j, k)],     F[g.idx(i, j - 1, k)],
                       F[g.idx(i, j + 1, k)],     F[g.idx(i - 1, j - 1, k)], F[g.idx(i - 1, j + 1, k)]};
        return Six{F[g.idx(i, j, k)],         F[g.idx(i, j - 1, k)],     F[g.idx(i - 1, j, k)],
                   F[g.idx(i + 1, j, k)],     F[g.idx(i - 1, j - 1, k)], F[g.idx(i + 1, j - 1, k)]};
    }
    DC_HD Six sixA(int j, bool u) const
    {
        const double a = g.A[g.row(j)], am = g.A[g.row(j - 1)], ap = g.A[g.row(j + 1)];
        return u ? Six{a, a, am, ap, am, ap} : Six{a, am, a, a, am, am};
    }
    DC_HD double P(int i, int j, int k) const
    {
        return COLP_NEW[g.idx2(i, j)] * g.A[g.row(j)] * WWIND[g.idx(i, j, k)];
    }
    DC_HD void operator()(int i, int j) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        if (j <= ny) {  // dyn_UVFLX_prepare.py:262-278
            const int wall = (j == 1) ? -1 : ((j == ny) ? 1 : 0);
            WWIND_UWIND[g.idx(i, j, 0)] = 0.;
            WWIND_UWIND[g.idx(i, j, nz)] = 0.;
            for (int k = 1; k < nz; k++)
                WWIND_UWIND[g.idx(i, j, k)] =
                    colpa_wwind(P(i, j, k), P(i - 1, j, k), P(i, j - 1, k), P(i, j + 1, k),
                                P(i - 1, j - 1, k), P(i - 1, j + 1, k), wall) *
                    interp_ks(UWIND[g.idx(i, j, k)], UWIND[g.idx(i, j, k - 1)], g.dsigma[k],
                              g.dsigma[k - 1], mkdiv(g.dsigma[k] + g.dsigma[k - 1], g.r_dss[k]));
        }
        if (i <= nx) {  // dyn_UVFLX_prepare.py:280-296
            WWIND_VWIND[g.idx(i, j, 0)] = 0.;
            WWIND_VWIND[g.idx(i, j, nz)] = 0.;
            for (int k = 1; k < nz; k++)
                WWIND_VWIND[g.idx(i, j, k)] =
                    colpa_wwind(P(i, j, k), P(i, j - 1, k), P(i - 1, j, k), P(i + 1, j, k),
                                P(i - 1, j - 1, k), P(i + 1, j - 1, k), 0) *
                    interp_ks(VWIND[g.idx(i, j, k)], VWIND[g.idx(i, j, k - 1)], g.dsigma[k],
                              g.dsigma[k - 1], mkdiv(g.dsigma[k] + g.dsigma[k - 1], g.r_dss[k]));
        }
        if (g.i_coupling) coupling(i, j);
        prep_rest(i, j);
    }
    // K dU/dz, K dV/dz on the interfaces (dyn_UVFLX_prepare.py:298-343); also the whole job of
    // TurbPrepBody.  threads as operator()
    DC_HD void coupling(int i, int j) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        {
            // the altitude and the wind of level k are carried to interface k+1
            if (j <= ny) {
                KMOM_dUWINDdz[g.idx(i, j, 0)] = 0.;
                KMOM_dUWINDdz[g.idx(i, j, nz)] = 0.;
                const Six C = six3(COLP, i, j, 0, true), A = sixA(j, true);
                double alt_km1 = interp_VAR_ds(six3(PHI, i, j, 0, true), true, j, ny);
                double w_km1 = UWIND[g.idx(i, j, 0)];
                for (int k = 1; k < nz; k++) {
                    const double alt = interp_VAR_ds(six3(PHI, i, j, k, true), true, j, ny);
                    const double w = UWIND[g.idx(i, j, k)];
                    KMOM_dUWINDdz[g.idx(i, j, k)] = kmom_dwinddz(
                        colpakmom_ds(six3(KMOM, i, j, k, true), six3(RHOVB, i, j, k, true), C, A,
                                     true, j, ny),
                        w, w_km1, alt_km1, alt);
                    alt_km1 = alt;
                    w_km1 = w;
                }
            }
            if (i <= nx) {
                KMOM_dVWINDdz[g.idx(i, j, 0)] = 0.;
                KMOM_dVWINDdz[g.idx(i, j, nz)] = 0.;
                const Six C = six3(COLP, i, j, 0, false), A = sixA(j, false);
                double alt_km1 = interp_VAR_ds(six3(PHI, i, j, 0, false), false, i, nx);
                double w_km1 = VWIND[g.idx(i, j, 0)];
                for (int k = 1; k < nz; k++) {
                    const double alt = interp_VAR_ds(six3(PHI, i, j, k, false), false, i, nx);
                    const double w = VWIND[g.idx(i, j, k)];
                    KMOM_dVWINDdz[g.idx(i, j, k)] = kmom_dwinddz(
                        colpakmom_ds(six3(KMOM, i, j, k, false), six3(RHOVB, i, j, k, false), C,
                                     A, false, i, nx),
                        w, w_km1, alt_km1, alt);
                    alt_km1 = alt;
                    w_km1 = w;
                }
            }
        }
    }
    DC_HD void prep_rest(int i, int j) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const double *u = UFLX, *v = VFLX;
        for (int k = 0; k < nz; k++) {  // dyn_UVFLX_prepare.py:345-436
            CFLX[g.idx(i, j, k)] =
                calc_CFLX(v[g.idx(i - 1, j - 1, k)], v[g.idx(i, j - 1, k)], v[g.idx(i - 1, j, k)],
                          v[g.idx(i, j, k)], v[g.idx(i - 1, j + 1, k)], v[g.idx(i, j + 1, k)]);
            QFLX[g.idx(i, j, k)] =
                calc_QFLX(u[g.idx(i - 1, j - 1, k)], u[g.idx(i - 1, j, k)], u[g.idx(i, j - 1, k)],
                          u[g.idx(i, j, k)], u[g.idx(i + 1, j - 1, k)], u[g.idx(i + 1, j, k)]);
            if (i <= nx) {
                DFLX[g.idx(i, j, k)] = calc_DFLX(
                    v[g.idx(i, j - 1, k)], v[g.idx(i, j, k)], v[g.idx(i, j + 1, k)],
                    u[g.idx(i, j - 1, k)], u[g.idx(i, j, k)], u[g.idx(i + 1, j - 1, k)],
                    u[g.idx(i + 1, j, k)]);
                EFLX[g.idx(i, j, k)] = calc_EFLX(
                    v[g.idx(i, j - 1, k)], v[g.idx(i, j, k)], v[g.idx(i, j + 1, k)],
                    u[g.idx(i, j - 1, k)], u[g.idx(i, j, k)], u[g.idx(i + 1, j - 1, k)],
                    u[g.idx(i + 1, j, k)]);
            }
            if (j <= ny) {
                SFLX[g.idx(i, j, k)] = calc_SFLX(
                    v[g.idx(i - 1, j, k)], v[g.idx(i - 1, j + 1, k)], v[g.idx(i, j, k)],
                    v[g.idx(i, j + 1, k)], u[g.idx(i - 1, j, k)], u[g.idx(i, j, k)],
                    u[g.idx(i + 1, j, k)]);
                TFLX[g.idx(i, j, k)] = calc_TFLX(
                    v[g.idx(i - 1, j, k)], v[g.idx(i - 1, j + 1, k)], v[g.idx(i, j, k)],
                    v[g.idx(i, j + 1, k)], u[g.idx(i - 1, j, k)], u[g.idx(i, j, k)],
                    u[g.idx(i + 1, j, k)]);
            }
            if (i <= nx && j <= ny) {
                BFLX[g.idx(i, j, k)] = calc_BFLX(
                    u[g.idx(i, j - 1, k)], u[g.idx(i + 1, j - 1, k)], u[g.idx(i, j, k)],
                    u[g.idx(i + 1, j, k)], u[g.idx(i, j + 1, k)], u[g.idx(i + 1, j + 1, k)]);
                RFLX[g.idx(i, j, k)] = calc_RFLX(
                    v[g.idx(i - 1, j, k)], v[g.idx(i - 1, j + 1, k)], v[g.idx(i, j, k)],
                    v[g.idx(i, j + 1, k)], v[g.idx(i + 1, j, k)], v[g.idx(i + 1, j + 1, k)]);
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// dUFLXdt: dyn_UFLX.py:339-434 + :69-199.  threads: i in [1, nx], j in [1, ny]
// (column nxs of the reference is garbage that the periodic BC overwrites)
// ---------------------------------------------------------------------------------------
struct UFLXTendencyBody {
    Geom g;
    const double *UFLX, *UWIND, *VWIND, *BFLX, *CFLX, *DFLX, *EFLX, *PHI, *COLP, *POTT, *PVTF,
        *PVTFVB, *WWIND_UWIND;
    double *dUFLXdt;
    // physics coupling (g.i_coupling): dyn_UFLX.py:136-170
    const double *PHIVB, *RHO, *SMOMXFLX, *KMOM_dUWINDdz;
    double *dUFLXdt_TURB;
    DC_HD Six six(const double *F, int i, int j, int k) const
    {
        return Six{F[g.idx(i, j, k)],     F[g.idx(i - 1, j, k)],     F[g.idx(i, j - 1, k)],
                   F[g.idx(i, j + 1, k)], F[g.idx(i - 1, j - 1, k)], F[g.idx(i - 1, j + 1, k)]};
    }
    DC_HD void operator()(int i, int j) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const int im1 = (i == 1) ? nx : i - 1;  // BCx, dyn_UFLX.py:367-374
        const double c = COLP[g.idx2(i, j)], c_im1 = COLP[g.idx2(i - 1, j)];
        const double fcos_is = cor_fcos(g.corf_is[g.row(j)], g.cos_lat_is[g.row(j)]);
        const double sinl = g.sin_lat_is[g.row(j)];
        const double scale = cor_scale(g.dlon_rad, g.dlat_rad);
        const double *U = UWIND, *V = VWIND;
        // coupled terms: surface momentum flux, and the interface altitude / K dU/dz of the
        // level's upper interface, carried down the column
        double smomflx_s = 0., altvb = 0., kd = 0.;
        if (g.i_coupling) {
            smomflx_s = interp_VAR_ds(six(SMOMXFLX, i, j, 0), true, j, ny);
            altvb = interp_VAR_ds(six(PHIVB, i, j, 0), true, j, ny) / div_g();
            kd = KMOM_dUWINDdz[g.idx(i, j, 0)];
        }
        for (int k = 0; k < nz; k++) {
            double bflx = BFLX[g.idx(i, j, k)], cflx = CFLX[g.idx(i, j, k)];
            double eflx = EFLX[g.idx(i, j, k)], dflx_jp1 = DFLX[g.idx(i, j + 1, k)];
            double cflx_jp1 = CFLX[g.idx(i, j + 1, k)];
            double bflx_im1 = BFLX[g.idx(im1, j, k)], dflx_im1 = DFLX[g.idx(im1, j, k)];
            double eflx_im1_jp1 = EFLX[g.idx(im1, j + 1, k)];
            if (j == 1) {  // BCy, dyn_UFLX.py:377-384
                dflx_im1 = 0.;
                cflx = 0.;
                eflx = 0.;
            }
            if (j == ny) {
                dflx_jp1 = 0.;
                cflx_jp1 = 0.;
                eflx_im1_jp1 = 0.;
            }
            const double u = U[g.idx(i, j, k)], u_im1 = U[g.idx(i - 1, j, k)],
                         u_ip1 = U[g.idx(i + 1, j, k)];
            double d = 0.;
            d = d + UVFLX_hor_adv(u, u_im1, u_ip1, U[g.idx(i, j - 1, k)], U[g.idx(i, j + 1, k)],
                                  U[g.idx(i - 1, j - 1, k)], U[g.idx(i - 1, j + 1, k)],
                                  U[g.idx(i + 1, j - 1, k)], U[g.idx(i + 1, j + 1, k)], bflx,
                                  bflx_im1, cflx, cflx_jp1, dflx_im1, dflx_jp1, eflx,
                                  eflx_im1_jp1, 1.);
            const Div ds = mkdiv(g.dsigma[k], g.r_dsigma[k]);
            d = d + ((WWIND_UWIND[g.idx(i, j, k)] - WWIND_UWIND[g.idx(i, j, k + 1)]) / ds);
            if (g.i_coupling) {
                const double altvb_kp1 =
                    interp_VAR_ds(six(PHIVB, i, j, k + 1), true, j, ny) / div_g();
                const double kd_kp1 = KMOM_dUWINDdz[g.idx(i, j, k + 1)];
                const double t =
                    turb_momentum(kd, kd_kp1, smomflx_s, altvb, altvb_kp1,
                                  interp_VAR_ds(six(RHO, i, j, k), true, j, ny), k, nz);
                dUFLXdt_TURB[g.idx(i, j, k)] = t;
                d = d + t;
                altvb = altvb_kp1;      // carried down the column
                kd = kd_kp1;
            }
            d = d + coriolis_UWIND(c, c_im1, V[g.idx(i, j, k)], V[g.idx(i - 1, j, k)],
                                   V[g.idx(i, j + 1, k)], V[g.idx(i - 1, j + 1, k)], u, u_im1,
                                   u_ip1, fcos_is, sinl, scale);
            d = d + pre_grad(PHI[g.idx(i, j, k)], PHI[g.idx(i - 1, j, k)], c, c_im1,
                             POTT[g.idx(i, j, k)], POTT[g.idx(i - 1, j, k)], PVTF[g.idx(i, j, k)],
                             PVTF[g.idx(i - 1, j, k)], PVTFVB[g.idx(i, j, k)],
                             PVTFVB[g.idx(i - 1, j, k)], PVTFVB[g.idx(i - 1, j, k + 1)],
                             PVTFVB[g.idx(i, j, k + 1)], ds, g.sigma_vb[k],
                             g.sigma_vb[k + 1], g.dyis);
            const double coef = g.UVFLX_dif_coef[k];
            if (coef > 0.)
                d = d + num_dif(UFLX[g.idx(i, j, k)], UFLX[g.idx(i - 1, j, k)],
                                UFLX[g.idx(i + 1, j, k)], UFLX[g.idx(i, j - 1, k)],
                                UFLX[g.idx(i, j + 1, k)], coef);
            dUFLXdt[g.idx(i, j, k)] = d;
        }
    }
};

// ---------------------------------------------------------------------------------------
// dVFLXdt: dyn_VFLX.py:322-408 + :67-198.  threads: i in [1, nx], j in [2, ny]
// (wall rows 1 and nys are NaN in the reference and zeroed by the BC after the Euler step)
// ---------------------------------------------------------------------------------------
struct VFLXTendencyBody {
    Geom g;
    const double *VFLX, *UWIND, *VWIND, *RFLX, *SFLX, *TFLX, *QFLX, *PHI, *COLP, *POTT, *PVTF,
        *PVTFVB, *WWIND_VWIND;
    double *dVFLXdt;
    // physics coupling (g.i_coupling): dyn_VFLX.py:134-166
    const double *PHIVB, *RHO, *SMOMYFLX, *KMOM_dVWINDdz;
    double *dVFLXdt_TURB;
    DC_HD Six six(const double *F, int i, int j, int k) const
    {
        return Six{F[g.idx(i, j, k)],     F[g.idx(i, j - 1, k)],     F[g.idx(i - 1, j, k)],
                   F[g.idx(i + 1, j, k)], F[g.idx(i - 1, j - 1, k)], F[g.idx(i + 1, j - 1, k)]};
    }
    DC_HD void operator()(int i, int j) const
    {
        const int nx = g.nx, nz = g.nz;
        const int ip1 = (i == nx) ? 1 : i + 1;  // BCx, dyn_VFLX.py:349-356
        const double c = COLP[g.idx2(i, j)], c_jm1 = COLP[g.idx2(i, j - 1)];
        const double fcos = cor_fcos(g.corf[g.row(j)], g.cos_lat[g.row(j)]);
        const double fcos_jm1 = cor_fcos(g.corf[g.row(j - 1)], g.cos_lat[g.row(j - 1)]);
        const double sinl = g.sin_lat[g.row(j)], sinl_jm1 = g.sin_lat[g.row(j - 1)];
        const double scale = cor_scale(g.dlon_rad, g.dlat_rad);
        const double dxjs = g.dxjs[g.row(j)];
        const double *U = UWIND, *V = VWIND;
        double smomflx_s = 0., altvb = 0., kd = 0.;   // coupled terms, as in UFLXTendencyBody
        if (g.i_coupling) {
            smomflx_s = interp_VAR_ds(six(SMOMYFLX, i, j, 0), false, i, nx);
            altvb = interp_VAR_ds(six(PHIVB, i, j, 0), false, i, nx) / div_g();
            kd = KMOM_dVWINDdz[g.idx(i, j, 0)];
        }
        for (int k = 0; k < nz; k++) {
            const double rflx = RFLX[g.idx(i, j, k)], qflx = QFLX[g.idx(i, j, k)];
            const double tflx = TFLX[g.idx(i, j, k)], rflx_jm1 = RFLX[g.idx(i, j - 1, k)];
            const double sflx_jm1 = SFLX[g.idx(i, j - 1, k)];
            const double qflx_ip1 = QFLX[g.idx(ip1, j, k)];
            const double tflx_ip1_jm1 = TFLX[g.idx(ip1, j - 1, k)];
            const double sflx_ip1 = SFLX[g.idx(ip1, j, k)];
            const double v = V[g.idx(i, j, k)];
            double d = 0.;
            d = d + UVFLX_hor_adv(v, V[g.idx(i, j - 1, k)], V[g.idx(i, j + 1, k)],
                                  V[g.idx(i - 1, j, k)], V[g.idx(i + 1, j, k)],
                                  V[g.idx(i - 1, j - 1, k)], V[g.idx(i + 1, j - 1, k)],
                                  V[g.idx(i - 1, j + 1, k)], V[g.idx(i + 1, j + 1, k)], rflx,
                                  rflx_jm1, qflx, qflx_ip1, sflx_jm1, sflx_ip1, tflx,
                                  tflx_ip1_jm1, -1.);
            const Div ds = mkdiv(g.dsigma[k], g.r_dsigma[k]);
            d = d + ((WWIND_VWIND[g.idx(i, j, k)] - WWIND_VWIND[g.idx(i, j, k + 1)]) / ds);
            if (g.i_coupling) {
                const double altvb_kp1 =
                    interp_VAR_ds(six(PHIVB, i, j, k + 1), false, i, nx) / div_g();
                const double kd_kp1 = KMOM_dVWINDdz[g.idx(i, j, k + 1)];
                const double t =
                    turb_momentum(kd, kd_kp1, smomflx_s, altvb, altvb_kp1,
                                  interp_VAR_ds(six(RHO, i, j, k), false, i, nx), k, g.nz);
                dVFLXdt_TURB[g.idx(i, j, k)] = t;
                d = d + t;
                altvb = altvb_kp1;      // carried down the column
                kd = kd_kp1;
            }
            d = d + coriolis_VWIND(c, c_jm1, U[g.idx(i, j, k)], U[g.idx(i, j - 1, k)],
                                   U[g.idx(i + 1, j, k)], U[g.idx(i + 1, j - 1, k)], fcos, sinl,
                                   fcos_jm1, sinl_jm1, scale);
            d = d + pre_grad(PHI[g.idx(i, j, k)], PHI[g.idx(i, j - 1, k)], c, c_jm1,
                             POTT[g.idx(i, j, k)], POTT[g.idx(i, j - 1, k)], PVTF[g.idx(i, j, k)],
                             PVTF[g.idx(i, j - 1, k)], PVTFVB[g.idx(i, j, k)],
                             PVTFVB[g.idx(i, j - 1, k)], PVTFVB[g.idx(i, j - 1, k + 1)],
                             PVTFVB[g.idx(i, j, k + 1)], ds, g.sigma_vb[k],
                             g.sigma_vb[k + 1], dxjs);
            const double coef = g.UVFLX_dif_coef[k];
            if (coef > 0.)
                d = d + num_dif(VFLX[g.idx(i, j, k)], VFLX[g.idx(i - 1, j, k)],
                                VFLX[g.idx(i + 1, j, k)], VFLX[g.idx(i, j - 1, k)],
                                VFLX[g.idx(i, j + 1, k)], coef);
            dVFLXdt[g.idx(i, j, k)] = d;
        }
    }
};

// ---------------------------------------------------------------------------------------
// dPOTTdt: dyn_POTT.py:180-216 + :55-110.  threads: i in [1, nx], j in [1, ny]
// ---------------------------------------------------------------------------------------
struct POTTTendencyBody {
    Geom g;
    const double *POTT, *UFLX, *VFLX, *COLP, *POTTVB, *WWIND, *COLP_NEW;
    double *dPOTTdt;
    // physics coupling (g.i_coupling): dyn_POTT.py:87-96
    const double *PHI, *PHIVB, *KHEAT, *RHO, *RHOVB, *SSHFLX;
    double *dPOTTdt_TURB;
    const double *dPOTTdt_RAD;   // radiative heating rate, dyn_POTT.py:40-41, :107-108
    DC_HD void operator()(int i, int j) const
    {
        const int nz = g.nz;
        const double c = COLP[g.idx2(i, j)];
        const double c_im1 = COLP[g.idx2(i - 1, j)], c_ip1 = COLP[g.idx2(i + 1, j)];
        const double c_jm1 = COLP[g.idx2(i, j - 1)], c_jp1 = COLP[g.idx2(i, j + 1)];
        const double cnew = COLP_NEW[g.idx2(i, j)];
        const Div A = mkdiv(g.A[g.row(j)], g.r_A[g.row(j)]);
        TurbMarch tm{0., 0., 0.};
        double surf = 0.;
        if (g.i_coupling) {
            tm = TurbMarch::top(PHI[g.idx(i, j, 0)], PHIVB[g.idx(i, j, 0)]);
            surf = SSHFLX[g.idx2(i, j)] / con_cp;
        }
        for (int k = 0; k < nz; k++) {
            const Div ds = mkdiv(g.dsigma[k], g.r_dsigma[k]);
            const double p = POTT[g.idx(i, j, k)];
            const double p_im1 = POTT[g.idx(i - 1, j, k)], p_ip1 = POTT[g.idx(i + 1, j, k)];
            const double p_jm1 = POTT[g.idx(i, j - 1, k)], p_jp1 = POTT[g.idx(i, j + 1, k)];
            double d = 0.;
            d = d + hor_adv(p, p_im1, p_ip1, p_jm1, p_jp1, UFLX[g.idx(i, j, k)],
                            UFLX[g.idx(i + 1, j, k)], VFLX[g.idx(i, j, k)],
                            VFLX[g.idx(i, j + 1, k)], A);
            d = d + vert_adv(POTTVB[g.idx(i, j, k)], POTTVB[g.idx(i, j, k + 1)],
                             WWIND[g.idx(i, j, k)], WWIND[g.idx(i, j, k + 1)], cnew, ds, k);
            if (g.i_coupling) {
                const int kp = k < nz - 1 ? k + 1 : k;
                const double t =
                    tm.step(p, POTT[g.idx(i, j, kp)], PHI[g.idx(i, j, kp)],
                            PHIVB[g.idx(i, j, k + 1)], RHOVB[g.idx(i, j, k + 1)],
                            KHEAT[g.idx(i, j, k + 1)], RHO[g.idx(i, j, k)], c, surf, k, nz);
                d = d + t;
                dPOTTdt_TURB[g.idx(i, j, k)] = t / c * 3600.;   // [K hr-1], dyn_POTT.py:97
            }
            const double coef = g.POTT_dif_coef[k];
            if (coef > 0.)
                d = d + num_dif_pw(p, p_im1, p_ip1, p_jm1, p_jp1, c, c_im1, c_ip1, c_jm1, c_jp1,
                                   coef);
            if (g.i_coupling) d = d + (dPOTTdt_RAD[g.idx(i, j, k)] * c);
            dPOTTdt[g.idx(i, j, k)] = d;
        }
    }
};

// ---------------------------------------------------------------------------------------
// dQVdt, dQCdt: dyn_moist.py:201-243 + :49-129.  threads: i in [1, nx], j in [1, ny]
// The reference reads Q[k-1] at k = 0 and Q[k+1] at k = nz-1 out of the column; both only
// enter products that vert_adv drops (k = 0) or multiplies by WWIND[nz] = 0.
// ---------------------------------------------------------------------------------------
struct MoistTendencyBody {
    Geom g;
    const double *QV, *QC, *UFLX, *VFLX, *COLP, *WWIND, *COLP_NEW;
    double *dQVdt, *dQCdt;
    // physics coupling (g.i_coupling): dyn_moist.py:100-112
    const double *PHI, *PHIVB, *KHEAT, *RHO, *RHOVB, *SLHFLX;
    double *dQVdt_TURB;
    DC_HD void one(const double *Q, double *dQ, int i, int j) const
    {
        const int nz = g.nz;
        const double c = COLP[g.idx2(i, j)];
        const double c_im1 = COLP[g.idx2(i - 1, j)], c_ip1 = COLP[g.idx2(i + 1, j)];
        const double c_jm1 = COLP[g.idx2(i, j - 1)], c_jp1 = COLP[g.idx2(i, j + 1)];
        const double cnew = COLP_NEW[g.idx2(i, j)];
        const Div A = mkdiv(g.A[g.row(j)], g.r_A[g.row(j)]);
        double q = Q[g.idx(i, j, 0)];
        double qvb = q;  // unused at k = 0
        const bool vap = (Q == QV);
        TurbMarch tm{0., 0., 0.};
        double surf = 0.;
        if (g.i_coupling) {
            tm = TurbMarch::top(PHI[g.idx(i, j, 0)], PHIVB[g.idx(i, j, 0)]);
            surf = vap ? SLHFLX[g.idx2(i, j)] / con_Lh : 0.;
        }
        for (int k = 0; k < nz; k++) {
            const double q_kp1 = (k + 1 < nz) ? Q[g.idx(i, j, k + 1)] : q;
            const double q_im1 = Q[g.idx(i - 1, j, k)], q_ip1 = Q[g.idx(i + 1, j, k)];
            const double q_jm1 = Q[g.idx(i, j - 1, k)], q_jp1 = Q[g.idx(i, j + 1, k)];
            double d = 0.;
            d = d + hor_adv(q, q_im1, q_ip1, q_jm1, q_jp1, UFLX[g.idx(i, j, k)],
                            UFLX[g.idx(i + 1, j, k)], VFLX[g.idx(i, j, k)],
                            VFLX[g.idx(i, j + 1, k)], A);
            const double qvb_kp1 = comp_VARVB_log(q_kp1, q);
            d = d + vert_adv(qvb, qvb_kp1, WWIND[g.idx(i, j, k)], WWIND[g.idx(i, j, k + 1)], cnew,
                             mkdiv(g.dsigma[k], g.r_dsigma[k]), k);
            if (g.i_coupling) {
                const int kp = k < nz - 1 ? k + 1 : k;
                const double t =
                    tm.step(q, q_kp1, PHI[g.idx(i, j, kp)], PHIVB[g.idx(i, j, k + 1)],
                            RHOVB[g.idx(i, j, k + 1)], KHEAT[g.idx(i, j, k + 1)],
                            RHO[g.idx(i, j, k)], c, surf, k, nz);
                d = d + t;
                if (vap) dQVdt_TURB[g.idx(i, j, k)] = t;
            }
            const double coef = g.moist_dif_coef[k];
            if (coef > 0.)
                d = d + num_dif_pw(q, q_im1, q_ip1, q_jm1, q_jp1, c, c_im1, c_ip1, c_jm1, c_jp1,
                                   coef);
            dQ[g.idx(i, j, k)] = d;
            q = q_kp1;
            qvb = qvb_kp1;  // comp_VARVB_log(Q[k+1], Q[k]) is next level's (VAR, VAR_km1)
        }
    }
    DC_HD void operator()(int i, int j) const
    {
        one(QV, dQVdt, i, j);
        one(QC, dQCdt, i, j);
    }
};

// ---------------------------------------------------------------------------------------
// Euler forward (pressure weighted): dyn_timestep.py:212-296 + BCs
// dyn_org_discretizations.py:388-393.  threads: i in [1, nx], j in [1, ny]
// ---------------------------------------------------------------------------------------
struct TimestepBody {
    Geom g;
    const double *COLP, *COLP_OLD, *UWIND_OLD, *dUFLXdt, *VWIND_OLD, *dVFLXdt, *POTT_OLD,
        *dPOTTdt, *QV_OLD, *dQVdt, *QC_OLD, *dQCdt;
    double *UWIND, *VWIND, *POTT, *QV, *QC;
    DC_HD void operator()(int i, int j) const
    {
        const int ny = g.ny, nz = g.nz;
        const double A = g.A[g.row(j)], A_jm1 = g.A[g.row(j - 1)], A_jp1 = g.A[g.row(j + 1)];
        const double *C = COLP, *CO = COLP_OLD;
        const double c = C[g.idx2(i, j)], co = CO[g.idx2(i, j)];
        const double colpa_is = interp_COLPA_is(
            c, C[g.idx2(i - 1, j)], C[g.idx2(i, j - 1)], C[g.idx2(i, j + 1)],
            C[g.idx2(i - 1, j + 1)], C[g.idx2(i - 1, j - 1)], A, A_jm1, A_jp1, j, ny);
        const double colpa_old_is = interp_COLPA_is(
            co, CO[g.idx2(i - 1, j)], CO[g.idx2(i, j - 1)], CO[g.idx2(i, j + 1)],
            CO[g.idx2(i - 1, j + 1)], CO[g.idx2(i - 1, j - 1)], A, A_jm1, A_jp1, j, ny);
        const double colpa_js =
            interp_COLPA_js(c, C[g.idx2(i, j - 1)], C[g.idx2(i - 1, j)], C[g.idx2(i + 1, j)],
                            C[g.idx2(i + 1, j - 1)], C[g.idx2(i - 1, j - 1)], A, A_jm1);
        const double colpa_old_js =
            interp_COLPA_js(co, CO[g.idx2(i, j - 1)], CO[g.idx2(i - 1, j)], CO[g.idx2(i + 1, j)],
                            CO[g.idx2(i + 1, j - 1)], CO[g.idx2(i - 1, j - 1)], A, A_jm1);
        const Div d_is = mkdiv(colpa_is), d_js = mkdiv(colpa_js), d_c = mkdiv(c);
        for (int k = 0; k < nz; k++) {
            put_xstag(g, UWIND, i, j, k,
                      euler_forward_pw(UWIND_OLD[g.idx(i, j, k)], dUFLXdt[g.idx(i, j, k)],
                                       d_is, colpa_old_is, g.dt));
            if (j >= 2)
                put_ystag(g, VWIND, i, j, k,
                          euler_forward_pw(VWIND_OLD[g.idx(i, j, k)], dVFLXdt[g.idx(i, j, k)],
                                           d_js, colpa_old_js, g.dt));
            else
                put_ystag(g, VWIND, i, 1, k, 0.);
            if (j == ny) put_ystag(g, VWIND, i, ny + 1, k, 0.);
            put_mass(g, POTT, i, j, k,
                     euler_forward_pw(POTT_OLD[g.idx(i, j, k)], dPOTTdt[g.idx(i, j, k)], d_c, co,
                                      g.dt));
            if (g.i_moist) {
                put_mass(g, QV, i, j, k,
                         euler_forward_pw(QV_OLD[g.idx(i, j, k)], dQVdt[g.idx(i, j, k)], d_c, co,
                                          g.dt));
                put_mass(g, QC, i, j, k,
                         euler_forward_pw(QC_OLD[g.idx(i, j, k)], dQCdt[g.idx(i, j, k)], d_c, co,
                                          g.dt));
            }
        }
    }
};

// ---------------------------------------------------------------------------------------
// fused moisture stage: dQVdt / dQCdt (dyn_moist.py:49-129) followed immediately by the
// pressure-weighted Euler step (dyn_timestep.py:292-296) and the boundary images; the
// tendencies are not materialised.  threads: i in [1, nx], j in the band.
// Q_OLD may alias Q_out (stage 2 overwrites the step-start state cell by cell).
// ---------------------------------------------------------------------------------------
struct MoistStageBody {
    Geom g;
    // UFLX / VFLX are formed on the fly from the winds and COLP (calc_UFLX / calc_VFLX, the very
    // expressions of the continuity kernel: same values as the stored fields, which the fused
    // path therefore does not write).  One tracer at a time: marching both in one loop needs
    // 153 registers and was twice as slow (one resident block per SM).
    const double *QV, *QC, *UWIND, *VWIND, *COLP, *WWIND, *COLP_NEW, *COLP_OLD, *QV_OLD, *QC_OLD;
    double *QV_out, *QC_out;
    DC_HD void one(const double *Q, const double *Q_OLD, double *Q_out, int i, int j) const
    {
        const int nz = g.nz;
        const size_t plane = g.plane;
        const double c = COLP[g.idx2(i, j)];
        const double c_im1 = COLP[g.idx2(i - 1, j)], c_ip1 = COLP[g.idx2(i + 1, j)];
        const double c_jm1 = COLP[g.idx2(i, j - 1)], c_jp1 = COLP[g.idx2(i, j + 1)];
        const double cnew = COLP_NEW[g.idx2(i, j)], cold = COLP_OLD[g.idx2(i, j)];
        const double dxjs = g.dxjs[g.row(j)], dxjs_jp1 = g.dxjs[g.row(j + 1)];
        const Div cn = mkdiv(cnew);
        const Div A = mkdiv(g.A[g.row(j)], g.r_A[g.row(j)]);
        const bool edge = (i == 1) || (i == g.nx) || (j == 1) || (j == g.ny);
        // comp_VARVB_log (dyn_functions.py:70-95) needs log and reciprocal of the clamped value
        // of both levels around an interface: computed once per level and carried (the
        // reference evaluates them twice per cell), same values, same result
        const double min_val = 0.0000001;
        const size_t o0 = g.idx(i, j, 0), o_jp1 = g.idx(i, j + 1, 0), o_jm1 = g.idx(i, j - 1, 0);
        double q = Q[o0];
        double qc = fmax(q, min_val), lq = log(qc), rq = 1. / qc;
        double qvb = q;  // unused at k = 0
        double w_k = WWIND[o0];
        for (int k = 0; k < nz; k++) {
            const size_t ko = (size_t)k * plane;
            const double uf = calc_UFLX(UWIND[o0 + ko], c, c_im1, g.dyis);
            const double uf_ip1 = calc_UFLX(UWIND[o0 + ko + 1], c_ip1, c, g.dyis);
            const double vf = calc_VFLX(VWIND[o0 + ko], c, c_jm1, dxjs);
            const double vf_jp1 = calc_VFLX(VWIND[o_jp1 + ko], c_jp1, c, dxjs_jp1);
            const double w_kp1 = WWIND[o0 + ko + plane];
            const double q_kp1 = (k + 1 < nz) ? Q[o0 + ko + plane] : q;
            const double q_im1 = Q[o0 + ko - 1], q_ip1 = Q[o0 + ko + 1];
            const double q_jm1 = Q[o_jm1 + ko], q_jp1 = Q[o_jp1 + ko];
            double d = 0.;
            d = d + hor_adv(q, q_im1, q_ip1, q_jm1, q_jp1, uf, uf_ip1, vf, vf_jp1, A);
            // QVVB_kp1 = comp_VARVB_log(VAR = Q[k+1], VAR_km1 = Q[k])
            const double qc_kp1 = fmax(q_kp1, min_val);
            double lq_kp1 = lq, rq_kp1 = rq, qvb_kp1 = qc_kp1;
            if (qc_kp1 != qc) {
                lq_kp1 = log(qc_kp1);
                rq_kp1 = 1. / qc_kp1;
                qvb_kp1 = ((lq - lq_kp1) / (rq_kp1 - rq));
            }
            d = d + vert_adv(qvb, qvb_kp1, w_k, w_kp1, cnew, mkdiv(g.dsigma[k], g.r_dsigma[k]), k);
            const double coef = g.moist_dif_coef[k];
            if (coef > 0.)
                d = d + num_dif_pw(q, q_im1, q_ip1, q_jm1, q_jp1, c, c_im1, c_ip1, c_jm1, c_jp1,
                                   coef);
            const double qn = euler_forward_pw(Q_OLD[o0 + ko], d, cn, cold, g.dt);
            if (edge)
                put_mass(g, Q_out, i, j, k, qn);
            else
                Q_out[o0 + ko] = qn;
            q = q_kp1;
            qc = qc_kp1;
            lq = lq_kp1;
            rq = rq_kp1;
            qvb = qvb_kp1;  // comp_VARVB_log(Q[k+1], Q[k]) is next level's (VAR, VAR_km1)
            w_k = w_kp1;
        }
    }
    DC_HD void operator()(int i, int j) const
    {
        one(QV, QV_OLD, QV_out, i, j);
        one(QC, QC_OLD, QC_out, i, j);
    }
};

// ---------------------------------------------------------------------------------------
// Euler forward of the moisture tracers only (fused path: U, V, POTT are stepped inside the
// stage kernel).  threads: i in [1, nx], j in the band
// ---------------------------------------------------------------------------------------
struct MoistEulerBody {
    Geom g;
    const double *COLP_NEW, *COLP_OLD, *QV_OLD, *dQVdt, *QC_OLD, *dQCdt;
    double *QV, *QC;
    DC_HD void operator()(int i, int j) const
    {
        const double co = COLP_OLD[g.idx2(i, j)];
        const Div c = mkdiv(COLP_NEW[g.idx2(i, j)]);
        for (int k = 0; k < g.nz; k++) {
            put_mass(g, QV, i, j, k,
                     euler_forward_pw(QV_OLD[g.idx(i, j, k)], dQVdt[g.idx(i, j, k)], c, co, g.dt));
            put_mass(g, QC, i, j, k,
                     euler_forward_pw(QC_OLD[g.idx(i, j, k)], dQCdt[g.idx(i, j, k)], c, co, g.dt));
        }
    }
};

// ---------------------------------------------------------------------------------------
// primary diagnostics: dyn_diagnostics.py:139-195 (PVTF, PVTFVB; PHI, PHIVB bottom-up;
// POTTVB).  threads: ALL columns i in [0, nx+1], j in [0, ny+1] (halos are computed, not
// exchanged).  nz+1 pow per column (the reference evaluates 2 per cell).
// ---------------------------------------------------------------------------------------
// MODE 0: every output (the factory entry); 1: what the third-generation stage kernel reads
// (PHI, POTTVB, PGCOL); 2: what the second-generation stage kernel reads (no PHIVB, no PGCOL)
template <int MODE>
struct PrimaryDiagBody {
    Geom g;
    const double *COLP, *POTT, *HSURF;
    double *PVTF, *PVTFVB, *PHI, *PHIVB, *POTTVB, *PGCOL;
    // production build: a / b as a * (correctly rounded reciprocal of b), <= 1 ulp off
    DC_HD static double fdiv(double a, double b) { return DC_FAST ? a * dc_rcp(b) : a / b; }
    // production build: pow_kappa (dc_point.h) instead of pow(x, kappa)
    DC_HD double exner(double p, const double *tab, bool shared) const
    {
        return DC_FAST ? pow_kappa_tab(p * 1e-5, pc, tab, shared) : pow(p / 100000., con_kappa);
    }
    // One BOTTOM-UP sweep: the hydrostatic integral needs that direction, everything else is
    // level-local or couples two neighbouring levels, so POTT is read once and nothing the
    // thread wrote is read back.  Same operands and operations as the reference's three
    // kernels (diag_PVTF / diag_PHI / diag_POTTVB).
    // PGCOL[k] = POTT/dsigma * (sigma_vb[k+1]*(PVTFVB[k+1]-PVTF) + sigma_vb[k]*(PVTF-PVTFVB[k]))
    // is the per-column sub-expression of the pressure-gradient term (dyn_functions.py:177-207),
    // formed here once per cell instead of four times per cell in the momentum kernels.
    // A thread marches NC columns, (i, j_lo + NC*jj + c), in lockstep.  Measured on B200
    // (0.25 deg x 64 levels): NC = 2 halves the resident warps (86 registers) and is 40 %
    // SLOWER than NC = 1 -- the sweep is bound by the latency of its exp/log/division chains,
    // not by instruction issue -- so NC stays 1.
    static constexpr int NC = 1;
    int j_lo, j_hi;
    PowCoef pc;
    // rows from j_split on are shifted by j_skip: [j_lo, j_split) and [j_split + j_skip, j_hi]
    int j_split = 1 << 30, j_skip = 0;
    DC_HD void operator()(int i, int jj) const { march(i, jj, pc.tab, false, nullptr, nullptr, 0); }
    // `tab` = the power table: pc.tab (global memory) or a copy in shared memory; `lev`:
    // sigma_vb | dsigma | r_dsigma, (nz+1) entries each, in shared memory (or NULL: the geometry
    // vectors in global memory).  The shipped kernel passes pc.tab and NULL: staging both in
    // shared memory measured no gain (the L1 serves them)
    // `ring`: this thread's slot of a shared-memory ring (DIAG_SLOTS slots, `rstride` doubles
    // apart) that POTT is copied into DIAG_PF levels ahead with cp.async (k_diag in dyncore.cu);
    // NULL: POTT through a register, one level ahead.  Why not registers: ptxas put the per-level
    // r_dsigma load and the POTT load of the NEXT level on the same scoreboard (decoded from the
    // control bits of the shipped kernel), so the first use of r_dsigma waited for the POTT
    // request issued 60 instructions earlier -- every level paid the full DRAM latency (ncu: 60 %
    // of all warp samples on that one DMUL) whatever the prefetch distance in the source was.
    // An asynchronous copy has no scoreboard: its only wait is the explicit wait_group.
    DC_HD void march(int i, int jj, const double *tab, bool shared, const double *lev,
                     double *ring, int rstride) const
    {
        const double *sig = lev ? lev : g.sigma_vb;
        const double *dsg = lev ? lev + (g.nz + 1) : g.dsigma;
        const double *rds = lev ? lev + 2 * (g.nz + 1) : g.r_dsigma;
        constexpr bool PV = MODE != 1, PHB = MODE == 0, PG = MODE != 2;
        const int nz = g.nz;
        const size_t plane = g.plane;
        int j0 = j_lo + NC * jj;
        if (j0 >= j_split) j0 += j_skip;   // two row ranges in one launch (halo rows of a band)
        const int nc = (j0 + NC - 1 <= j_hi) ? NC : j_hi - j0 + 1;   // ragged last thread row
        double colp[NC], p_kp12[NC], pw_kp12[NC], phivb[NC], pvtf_kp1[NC], pott_kp1[NC];
        size_t o[NC];   // one running offset per column serves every field
        for (int c = 0; c < NC; c++) {
            const int j = c < nc ? j0 + c : j0;   // a masked second column shadows the first
            colp[c] = COLP[g.idx2(i, j)];
            p_kp12[c] = g.pair_top + sig[nz] * colp[c];
            pw_kp12[c] = exner(p_kp12[c], tab, shared);
            o[c] = g.idx(i, j, nz);
            phivb[c] = HSURF[g.idx2(i, j)] * con_g;
            pvtf_kp1[c] = 0.;
            pott_kp1[c] = 0.;
            if (c < nc) {
                if (PV) PVTFVB[o[c]] = pw_kp12[c];
                if (PHB) PHIVB[o[c]] = phivb[c];
            }
        }
        double svb_kp1 = sig[nz];
        // without a ring (host emulation, DC_DIAG_PF=0): POTT of the next level through a
        // register, requested one iteration ahead
        constexpr int PF = 1;
        double pott_q[NC][PF];
#if defined(__CUDA_ARCH__)
        int slot_rd = 0, slot_wr = DIAG_PF & (DIAG_SLOTS - 1);
        if (ring) {
#pragma unroll
            for (int n = 0; n < DIAG_PF; n++) {
                if (nz - 1 - n >= 0)
                    dc_cp_async8(ring + n * rstride, POTT + (o[0] - (size_t)(n + 1) * plane));
                dc_cp_commit();
            }
        } else
#endif
        for (int c = 0; c < NC; c++)
            for (int n = 0; n < PF; n++)
                pott_q[c][n] = nz - 1 - n >= 0 ? POTT[o[c] - (size_t)(n + 1) * plane] : 0.;
        for (int k = nz - 1; k >= 0; k--) {
            const double svb = sig[k];
            const Div ds = mkdiv(dsg[k], rds[k]);
            for (int c = 0; c < NC; c++) {
                o[c] -= plane;
                double pott;
#if defined(__CUDA_ARCH__)
                if (ring) {
                    // the slot written now was read one level ago and its value consumed since
                    dc_cp_wait<DIAG_PF - 1>();
                    pott = ring[slot_rd * rstride];
                    if (k - DIAG_PF >= 0)
                        dc_cp_async8(ring + slot_wr * rstride, POTT + (o[c] - (size_t)DIAG_PF * plane));
                    dc_cp_commit();
                    slot_rd = (slot_rd + 1) & (DIAG_SLOTS - 1);
                    slot_wr = (slot_wr + 1) & (DIAG_SLOTS - 1);
                } else
#endif
                {
                pott = pott_q[c][0];
#pragma unroll
                for (int n = 0; n + 1 < PF; n++) pott_q[c][n] = pott_q[c][n + 1];
                if (k - PF >= 0) pott_q[c][PF - 1] = POTT[o[c] - (size_t)PF * plane];
                }
                const double p_km12 = g.pair_top + svb * colp[c];
                const double pw_km12 = exner(p_km12, tab, shared);
                const double pvtf = fdiv(1. / (1. + con_kappa) *
                                             (pw_kp12[c] * p_kp12[c] - pw_km12 * p_km12),
                                         p_kp12[c] - p_km12);
                // diag_PHI_cpu
                const double phi = phivb[c] - con_cp * (pott * (pvtf - pw_kp12[c]));
                phivb[c] = phi - con_cp * (pott * (pw_km12 - pvtf));
                // diag_POTTVB_cpu: interface k+1 between level k (above) and k+1 (below)
                double pottvb = 0.;
                if (k + 1 <= nz - 1)
                    pottvb = fdiv(+(pw_kp12[c] - pvtf) * pott + (pvtf_kp1[c] - pw_kp12[c]) * pott_kp1[c],
                                  pvtf_kp1[c] - pvtf);
                if (c < nc) {
                    if (PV) {
                        PVTF[o[c]] = pvtf;
                        PVTFVB[o[c]] = pw_km12;
                    }
                    if (PG)
                        PGCOL[o[c]] = pott / ds * (svb_kp1 * (pw_kp12[c] - pvtf) + svb * (pvtf - pw_km12));
                    PHI[o[c]] = phi;
                    if (PHB) PHIVB[o[c]] = phivb[c];
                    if (k + 1 <= nz - 1) {
                        POTTVB[o[c] + plane] = pottvb;
                        if (k + 1 == nz - 1)  // extrapolate model bottom (dyn_diagnostics.py:188-191)
                            POTTVB[o[c] + 2 * plane] = pott_kp1[c] - (pottvb - pott_kp1[c]);
                        if (k == 0)           // extrapolate model top (dyn_diagnostics.py:184-187)
                            POTTVB[o[c]] = pott - (pottvb - pott);
                    }
                }
                p_kp12[c] = p_km12;
                pw_kp12[c] = pw_km12;
                pvtf_kp1[c] = pvtf;
                pott_kp1[c] = pott;
            }
            svb_kp1 = svb;
        }
    }
};

// ---------------------------------------------------------------------------------------
// secondary diagnostics: dyn_diagnostics.py:199-222.  threads: all columns incl. halos
// ---------------------------------------------------------------------------------------
struct SecondaryDiagBody {
    Geom g;
    const double *POTTVB, *PVTFVB, *POTT, *PVTF, *UWIND, *VWIND;
    double *TAIRVB, *PAIRVB, *RHOVB, *TAIR, *PAIR, *RHO, *WINDX, *WINDY, *WIND;
    DC_HD void operator()(int i, int j) const
    {
        const int nz = g.nz;
        for (int k = 0; k <= nz; k++) {
            const double pvb = PVTFVB[g.idx(i, j, k)];
            const double pairvb = 100000. * pow(pvb, 1. / con_kappa);
            const double tairvb = POTTVB[g.idx(i, j, k)] * pvb;
            PAIRVB[g.idx(i, j, k)] = pairvb;
            TAIRVB[g.idx(i, j, k)] = tairvb;
            RHOVB[g.idx(i, j, k)] = pairvb / (con_Rd * tairvb);
        }
        for (int k = 0; k < nz; k++) {
            const double pv = PVTF[g.idx(i, j, k)];
            const double tair = POTT[g.idx(i, j, k)] * pv;
            const double pair = 100000. * pow(pv, 1. / con_kappa);
            TAIR[g.idx(i, j, k)] = tair;
            PAIR[g.idx(i, j, k)] = pair;
            RHO[g.idx(i, j, k)] = pair / (con_Rd * tair);
            const double wx = (UWIND[g.idx(i, j, k)] + UWIND[g.idx(i + 1, j, k)]) / 2.;
            const double wy = (VWIND[g.idx(i, j, k)] + VWIND[g.idx(i, j + 1, k)]) / 2.;
            WINDX[g.idx(i, j, k)] = wx;
            WINDY[g.idx(i, j, k)] = wy;
            WIND[g.idx(i, j, k)] = sqrt(wx * wx + wy * wy);
        }
    }
};

// ---------------------------------------------------------------------------------------
// coupled terms BESIDE the fused dry stage kernel (dc_handle::coupled_impl == 2; experimental,
// off by default).  The stage kernel advances U, V, POTT with the dry tendencies; these two
// kernels add what the physics coupling fields contribute:
//   TurbPrepBody   K dU/dz, K dV/dz on the interfaces from the stage's INPUT state (= the
//                  coupling part of PrepBody).  threads: i in [1, nx+1], j in [1, ny+1]
//   TurbApplyBody  X_out += dt * dX_turb / C_new for U, V, POTT (QV, QC) and the boundary
//                  images of the updated cells; dX_turb as in dyn_UFLX.py:136-170,
//                  dyn_VFLX.py:134-166, dyn_POTT.py:87-108, dyn_moist.py:100-112, C_new as in the
//                  Euler step (dyn_timestep.py:34-79).  threads: i in [1, nx], j in [1, ny]
// The reference adds these terms INSIDE the tendency sum; adding them after the Euler step is
// the same arithmetic in another order (differences at rounding level, tests: tolerances).
// ---------------------------------------------------------------------------------------
struct TurbPrepBody {
    PrepBody p;
    DC_HD void operator()(int i, int j) const { p.coupling(i, j); }
};

struct TurbApplyBody {
    Geom g;
    const double *POTT_in, *QV_in, *QC_in;     // the stage's input state (vertical gradients)
    const double *COLP, *COLP_NEW;             // column pressure of the input state / new
    const double *PHI, *PHIVB, *KMOM_dUWINDdz, *KMOM_dVWINDdz, *KHEAT, *RHO, *RHOVB, *SMOMXFLX,
        *SMOMYFLX, *SSHFLX, *SLHFLX, *dPOTTdt_RAD;
    double *U_out, *V_out, *T_out, *QV_out, *QC_out;
    double *dUFLXdt_TURB, *dVFLXdt_TURB, *dPOTTdt_TURB, *dQVdt_TURB;
    DC_HD Six six_u(const double *F, int i, int j, int k) const
    {
        return Six{F[g.idx(i, j, k)],     F[g.idx(i - 1, j, k)],     F[g.idx(i, j - 1, k)],
                   F[g.idx(i, j + 1, k)], F[g.idx(i - 1, j - 1, k)], F[g.idx(i - 1, j + 1, k)]};
    }
    DC_HD Six six_v(const double *F, int i, int j, int k) const
    {
        return Six{F[g.idx(i, j, k)],     F[g.idx(i, j - 1, k)],     F[g.idx(i - 1, j, k)],
                   F[g.idx(i + 1, j, k)], F[g.idx(i - 1, j - 1, k)], F[g.idx(i + 1, j - 1, k)]};
    }
    DC_HD void tracer(const double *Q_in, double *Q_out, double *TURB, double surf, double c,
                      Div d_c, int i, int j) const
    {
        const int nz = g.nz;
        TurbMarch tm = TurbMarch::top(PHI[g.idx(i, j, 0)], PHIVB[g.idx(i, j, 0)]);
        double q = Q_in[g.idx(i, j, 0)];
        for (int k = 0; k < nz; k++) {
            const int kp = k < nz - 1 ? k + 1 : k;
            const double q_kp1 = Q_in[g.idx(i, j, kp)];
            const double t = tm.step(q, q_kp1, PHI[g.idx(i, j, kp)], PHIVB[g.idx(i, j, k + 1)],
                                     RHOVB[g.idx(i, j, k + 1)], KHEAT[g.idx(i, j, k + 1)],
                                     RHO[g.idx(i, j, k)], c, surf, k, nz);
            if (TURB) TURB[g.idx(i, j, k)] = t;
            put_mass(g, Q_out, i, j, k, Q_out[g.idx(i, j, k)] + g.dt * t / d_c);
            q = q_kp1;
        }
    }
    DC_HD void operator()(int i, int j) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const double A = g.A[g.row(j)], A_jm1 = g.A[g.row(j - 1)], A_jp1 = g.A[g.row(j + 1)];
        const double *CN = COLP_NEW;
        const double c = COLP[g.idx2(i, j)], cn = CN[g.idx2(i, j)];
        const Div d_c = mkdiv(cn);
        {   // U
            const Div d_is = mkdiv(interp_COLPA_is(
                cn, CN[g.idx2(i - 1, j)], CN[g.idx2(i, j - 1)], CN[g.idx2(i, j + 1)],
                CN[g.idx2(i - 1, j + 1)], CN[g.idx2(i - 1, j - 1)], A, A_jm1, A_jp1, j, ny));
            const double smomflx_s = interp_VAR_ds(six_u(SMOMXFLX, i, j, 0), true, j, ny);
            double altvb = interp_VAR_ds(six_u(PHIVB, i, j, 0), true, j, ny) / div_g();
            double kd = KMOM_dUWINDdz[g.idx(i, j, 0)];
            for (int k = 0; k < nz; k++) {
                const double altvb_kp1 =
                    interp_VAR_ds(six_u(PHIVB, i, j, k + 1), true, j, ny) / div_g();
                const double kd_kp1 = KMOM_dUWINDdz[g.idx(i, j, k + 1)];
                const double t =
                    turb_momentum(kd, kd_kp1, smomflx_s, altvb, altvb_kp1,
                                  interp_VAR_ds(six_u(RHO, i, j, k), true, j, ny), k, nz);
                dUFLXdt_TURB[g.idx(i, j, k)] = t;
                put_xstag(g, U_out, i, j, k, U_out[g.idx(i, j, k)] + g.dt * t / d_is);
                altvb = altvb_kp1;
                kd = kd_kp1;
            }
        }
        if (j >= 2) {   // V (rows 1 and ny+1 are walls: 0)
            const Div d_js = mkdiv(interp_COLPA_js(
                cn, CN[g.idx2(i, j - 1)], CN[g.idx2(i - 1, j)], CN[g.idx2(i + 1, j)],
                CN[g.idx2(i + 1, j - 1)], CN[g.idx2(i - 1, j - 1)], A, A_jm1));
            const double smomflx_s = interp_VAR_ds(six_v(SMOMYFLX, i, j, 0), false, i, nx);
            double altvb = interp_VAR_ds(six_v(PHIVB, i, j, 0), false, i, nx) / div_g();
            double kd = KMOM_dVWINDdz[g.idx(i, j, 0)];
            for (int k = 0; k < nz; k++) {
                const double altvb_kp1 =
                    interp_VAR_ds(six_v(PHIVB, i, j, k + 1), false, i, nx) / div_g();
                const double kd_kp1 = KMOM_dVWINDdz[g.idx(i, j, k + 1)];
                const double t =
                    turb_momentum(kd, kd_kp1, smomflx_s, altvb, altvb_kp1,
                                  interp_VAR_ds(six_v(RHO, i, j, k), false, i, nx), k, nz);
                dVFLXdt_TURB[g.idx(i, j, k)] = t;
                put_ystag(g, V_out, i, j, k, V_out[g.idx(i, j, k)] + g.dt * t / d_js);
                altvb = altvb_kp1;
                kd = kd_kp1;
            }
        }
        {   // POTT: turbulent transport + surface sensible heat flux + radiative heating
            TurbMarch tm = TurbMarch::top(PHI[g.idx(i, j, 0)], PHIVB[g.idx(i, j, 0)]);
            const double surf = SSHFLX[g.idx2(i, j)] / con_cp;
            double p = POTT_in[g.idx(i, j, 0)];
            for (int k = 0; k < nz; k++) {
                const int kp = k < nz - 1 ? k + 1 : k;
                const double p_kp1 = POTT_in[g.idx(i, j, kp)];
                const double t = tm.step(p, p_kp1, PHI[g.idx(i, j, kp)], PHIVB[g.idx(i, j, k + 1)],
                                         RHOVB[g.idx(i, j, k + 1)], KHEAT[g.idx(i, j, k + 1)],
                                         RHO[g.idx(i, j, k)], c, surf, k, nz);
                dPOTTdt_TURB[g.idx(i, j, k)] = t / c * 3600.;   // [K hr-1], dyn_POTT.py:97
                const double d = t + (dPOTTdt_RAD[g.idx(i, j, k)] * c);
                put_mass(g, T_out, i, j, k, T_out[g.idx(i, j, k)] + g.dt * d / d_c);
                p = p_kp1;
            }
        }
        if (g.i_moist) {
            tracer(QV_in, QV_out, dQVdt_TURB, SLHFLX[g.idx2(i, j)] / con_Lh, c, d_c, i, j);
            tracer(QC_in, QC_out, nullptr, 0., c, d_c, i, j);
        }
    }
};

// ---------------------------------------------------------------------------------------
// turbulence module: turb_compute.py:190-204 (launcher) + :53-145 (bulk_richardson_py,
// compute_K_coefs_py, run_all_py) + misc_meteo_utilities.py:36-49.  KMOM / KHEAT from the
// bulk Richardson number and a Blackadar mixing length on the interior interfaces
// k in [1, nz-1] of every column incl. the halo.  threads: i in [0, nx+1], j in [0, ny+1]
// Mind the launcher's argument mapping: "PHI_k" = PHIVB[k], "POTT_k" = POTTVB[k], km05 = the
// full level above the interface (k-1), kp05 = the full level below it (k).
// 10 field accesses per interface; the full-level values of level k are kept for k+1.
// ---------------------------------------------------------------------------------------
struct TurbulenceBody {
    Geom g;
    const double *PHIVB, *HSURF, *PHI, *QV, *WINDX, *WINDY, *POTTVB, *POTT;
    double *KMOM, *KHEAT;
    DC_HD void operator()(int i, int j) const
    {
        const double Ri_c = 1.0, free_mix_len = 200., con_k = 0.35, con_Pr = 0.72;
        const double min_wind_diff = 0.0001, min_KMOM = 0.000001, max_KMOM = 0.01;
        const int nz = g.nz;
        const double hsurf = HSURF[g.idx2(i, j)];
        double phi_m = PHI[g.idx(i, j, 0)], qv_m = QV[g.idx(i, j, 0)], wx_m = WINDX[g.idx(i, j, 0)],
               wy_m = WINDY[g.idx(i, j, 0)], pott_m = POTT[g.idx(i, j, 0)];
        for (int k = 1; k < nz; k++) {
            const double phi_p = PHI[g.idx(i, j, k)], qv_p = QV[g.idx(i, j, k)],
                         wx_p = WINDX[g.idx(i, j, k)], wy_p = WINDY[g.idx(i, j, k)],
                         pott_p = POTT[g.idx(i, j, k)];
            double WINDX_km05 = wx_m, WINDY_km05 = wy_m;
            if (WINDX_km05 == wx_p) WINDX_km05 += min_wind_diff;
            if (WINDY_km05 == wy_p) WINDY_km05 += min_wind_diff;
            const double ALT_k = PHIVB[g.idx(i, j, k)] / con_g;
            const double ALT_km05 = phi_m / con_g, ALT_kp05 = phi_p / con_g;
            const double HGT_k = ALT_k - hsurf;
            const double mix_len = con_k * HGT_k / (1. + con_k * HGT_k / free_mix_len);
            const double QV_k = comp_VARVB_log(qv_p, qv_m);
            const double POTT_v_k = POTTVB[g.idx(i, j, k)] * (1. + QV_k / 0.622) / (1. + QV_k);
            const double dx = WINDX_km05 - wx_p, dy = WINDY_km05 - wy_p;
            const double dalt = ALT_km05 - ALT_kp05;
            const double Ri_b_k =
                ((con_g / POTT_v_k * (pott_m - pott_p) * dalt) / (dx * dx + dy * dy));
            const double sx = dx / dalt, sy = dy / dalt;
            const double shear_term = sqrt(sx * sx + sy * sy);
            double KMOM_k = mix_len * mix_len * shear_term * (Ri_c - Ri_b_k) / Ri_c;
            if (KMOM_k < min_KMOM) KMOM_k = min_KMOM;
            if (KMOM_k > max_KMOM) KMOM_k = max_KMOM;
            KMOM[g.idx(i, j, k)] = KMOM_k;
            KHEAT[g.idx(i, j, k)] = KMOM_k / con_Pr;
            phi_m = phi_p; qv_m = qv_p; wx_m = wx_p; wy_m = wy_p; pott_m = pott_p;
        }
    }
};

// ---------------------------------------------------------------------------------------
// run-time diagnostics on the device (io_functions.py:70-114: diagnose_print_diag_fields and
// the crash check of print_ts_info): vmax, mass-weighted mean wind and potential temperature,
// area-weighted mean column pressure, NaN / over-speed check of UWIND.  The reference copies
// WIND, COLP and POTT (three full fields) to the host every nth_ts_print_diag steps; here
// the fields are reduced where they are and 7 numbers per row travel.  Two deterministic
// passes, no atomics: per-column partials (threads: i in [1, nx+1], j in the band), then one
// thread per row adds its columns in ascending i; the host adds the rows.
//   col planes [NJ][NI]: 0 sum_k WIND, 1 sum_k POTT, 2 max_k WIND, 3 max_k UWIND, 4 #NaN(UWIND)
//   row vectors [NJ]:    0 sum WIND*COLP*A, 1 sum POTT*COLP*A, 2 sum COLP*A, 3 sum A,
//                        4 max WIND, 5 max UWIND, 6 #NaN(UWIND)
// WIND = sqrt(WINDX^2 + WINDY^2) as in diag_secondary (dyn_diagnostics.py:199-222).
// ---------------------------------------------------------------------------------------
struct RunDiagColumnBody {
    Geom g;
    const double *UWIND, *VWIND, *POTT;
    double *col;
    DC_HD void operator()(int i, int j) const
    {
        const size_t o2 = g.idx2(i, j);
        double sw = 0., sp = 0., mw = 0., mu = -1e300, nan = 0.;
        for (int k = 0; k < g.nz; k++) {
            const double u = UWIND[g.idx(i, j, k)];
            if (u != u) nan += 1.;
            mu = fmax(mu, u);
            if (i <= g.nx) {
                const double wx = (u + UWIND[g.idx(i + 1, j, k)]) / 2.;
                const double wy = (VWIND[g.idx(i, j, k)] + VWIND[g.idx(i, j + 1, k)]) / 2.;
                const double w = sqrt(wx * wx + wy * wy);
                sw += w;
                mw = fmax(mw, w);
                sp += POTT[g.idx(i, j, k)];
            }
        }
        col[0 * g.plane + o2] = sw;
        col[1 * g.plane + o2] = sp;
        col[2 * g.plane + o2] = mw;
        col[3 * g.plane + o2] = mu;
        col[4 * g.plane + o2] = nan;
    }
};
struct RunDiagRowBody {
    Geom g;
    const double *col, *COLP;
    double *rows;
    DC_HD void operator()(int, int j) const
    {
        const double A = g.A[g.row(j)];
        double s_w = 0., s_p = 0., s_ca = 0., s_a = 0., m_w = 0., m_u = -1e300, nan = 0.;
        for (int i = 1; i <= g.nx + 1; i++) {
            const size_t o2 = g.idx2(i, j);
            m_u = fmax(m_u, col[3 * g.plane + o2]);
            nan += col[4 * g.plane + o2];
            if (i <= g.nx) {
                const double ca = COLP[o2] * A;
                s_w += col[0 * g.plane + o2] * ca;
                s_p += col[1 * g.plane + o2] * ca;
                s_ca += ca;
                s_a += A;
                m_w = fmax(m_w, col[2 * g.plane + o2]);
            }
        }
        const int jd = g.row(j);
        rows[0 * g.NJ + jd] = s_w;
        rows[1 * g.NJ + jd] = s_p;
        rows[2 * g.NJ + jd] = s_ca;
        rows[3 * g.NJ + jd] = s_a;
        rows[4 * g.NJ + jd] = m_w;
        rows[5 * g.NJ + jd] = m_u;
        rows[6 * g.NJ + jd] = nan;
    }
};

// ---------------------------------------------------------------------------------------
// continuity, single pass (dyn_continuity.py:170-228 + BCs dyn_org_discretizations.py:114-117)
//
// ContinuityBody above parks the running flux-divergence prefix in WWIND and re-reads it
// once the column total (dCOLPdt) is known: 5 field accesses per cell.  Here a block owns 32
// consecutive longitudes of ONE row and splits the column over its warps: warp w keeps the
// flux divergences of levels [w*CL, (w+1)*CL) in registers, the warps then chain their
// sequential prefix sums in level order (same additions in the same order as numba's
// FLXDIV.sum(axis=2): bit-identical), and every thread finalises WWIND for its own levels
// from registers.  U and V are read once, WWIND is written once: 3 accesses per cell.
// Written against a small SPMD layer (CT_*) so that tests/emu runs the same body.
// ---------------------------------------------------------------------------------------
#ifndef DC_CT_TX
#define DC_CT_TX 32
#endif
constexpr int CT_TX = DC_CT_TX;   // longitudes per block
#ifndef DC_CT_L
#define DC_CT_L 16
#endif
constexpr int CT_L = DC_CT_L;     // levels per thread
// (measured and dropped, round 2: per-warp running sums joined after ONE barrier instead of the
// chain through nw barriers -- 1.07 against 0.71 ms per step: two more live doubles per thread
// spill at the 128-register budget the 64 loads in flight need)
constexpr int CT_MAXW = 128 / CT_L;   // nz <= 128

#if defined(__CUDA_ARCH__)
#define CT_PRIV(type, name) type name
#define CT_PRIVN(type, name, n) type name[n]
#define CT_P(name) name
#define CT_PHASE(nt) {                 \
        const int tid = threadIdx.x;   \
        if (tid < (nt)) {
#define CT_PHASE_END \
    }                \
    }                \
    __syncthreads();
#else
#define CT_PRIV(type, name) type name[CT_TX * CT_MAXW]
#define CT_PRIVN(type, name, n) type name[CT_TX * CT_MAXW][n]
#define CT_P(name) name[tid]
#define CT_PHASE(nt) for (int tid = 0; tid < (nt); tid++) { {
#define CT_PHASE_END } }
#endif

struct ContinuitySmem {
    double tot[CT_MAXW][CT_TX];   // running sum at the end of each warp's level range
};

template <int MODE>
struct ContinuityTileBody {
    Geom g;
    const double *UWIND, *VWIND, *COLP, *COLP_OLD;
    double *UFLX, *VFLX, *FLXDIV, *WWIND, *COLP_NEW, *dCOLPdt;
    int j_lo;   // first row of the launch; block (bx, by) owns columns 1+32*bx.., row j_lo+by
    // rows from j_split on are shifted by j_skip (two row ranges in one launch: the band-edge
    // rows of the pipelined band step)
    int j_split = 1 << 30, j_skip = 0;

    DC_HD void run_block(int bx, int by, ContinuitySmem &s) const
    {
        const int nz = g.nz, nw = (nz + CT_L - 1) / CT_L, nt = nw * CT_TX;
        const int j = (j_lo + by >= j_split) ? j_lo + by + j_skip : j_lo + by;
        CT_PRIVN(double, fd, CT_L);   // flux divergence, then running sum, of the own levels
        // COLP_OLD of the column, requested with the winds: its first use comes after the
        // prefix sums, where a fresh request cost 8 % of the kernel's warp samples (ncu)
        CT_PRIV(double, c_old);
        // ---- flux divergences of the own levels -------------------------------------------
        CT_PHASE(nt)
            const int w = tid / CT_TX, i = 1 + bx * CT_TX + tid % CT_TX;
            if (i <= g.nx) {
                const double c = COLP[g.idx2(i, j)];
                CT_P(c_old) = COLP_OLD[g.idx2(i, j)];
                const double c_im1 = COLP[g.idx2(i - 1, j)], c_ip1 = COLP[g.idx2(i + 1, j)];
                const double c_jm1 = COLP[g.idx2(i, j - 1)], c_jp1 = COLP[g.idx2(i, j + 1)];
                const double dxjs = g.dxjs[g.row(j)], dxjs_jp1 = g.dxjs[g.row(j + 1)];
                const Div A = mkdiv(g.A[g.row(j)], g.r_A[g.row(j)]);
                // All 4*CT_L loads of the thread are issued before the first use (levels beyond
                // nz re-read level nz-1 and are never used): the kernel is bound by memory
                // latency, so the bytes in flight per warp are what counts
                const size_t o_u = g.idx(i, j, 0), o_v1 = g.idx(i, j + 1, 0);
                double ru[CT_L], ru1[CT_L], rv[CT_L], rv1[CT_L];
#pragma unroll
                for (int l = 0; l < CT_L; l++) {
                    const int k = w * CT_L + l;
                    const size_t ko = (size_t)(k < nz ? k : nz - 1) * g.plane;
                    ru[l] = UWIND[o_u + ko];
                    ru1[l] = UWIND[o_u + ko + 1];
                    rv[l] = VWIND[o_u + ko];
                    rv1[l] = VWIND[o_v1 + ko];
                }
                // keep the compiler from sinking the loads back into the arithmetic below
                asm volatile("" ::: "memory");
#pragma unroll
                for (int l = 0; l < CT_L; l++) {
                    const int k = w * CT_L + l;
                    const double uf = calc_UFLX(ru[l], c, c_im1, g.dyis);
                    const double uf_ip1 = calc_UFLX(ru1[l], c_ip1, c, g.dyis);
                    const double vf = calc_VFLX(rv[l], c, c_jm1, dxjs);
                    const double vf_jp1 = calc_VFLX(rv1[l], c_jp1, c, dxjs_jp1);
                    const double f =
                        calc_FLXDIV(uf, uf_ip1, vf, vf_jp1, g.dsigma[k < nz ? k : nz - 1], A);
                    if ((MODE & 2) && k < nz) {
                        put_xstag(g, UFLX, i, j, k, uf);
                        put_ystag(g, VFLX, i, j, k, vf);
                        if (j == g.ny) put_ystag(g, VFLX, i, g.ny + 1, k, 0.);
                    }
                    if ((MODE & 1) && k < nz) FLXDIV[g.idx(i, j, k)] = f;
                    CT_P(fd)[l] = f;
                }
            }
        CT_PHASE_END
        // ---- sequential ascending sum, chained through the warps in level order -----------
        for (int ww = 0; ww < nw; ww++) {
            CT_PHASE(nt)
                const int w = tid / CT_TX, lane = tid % CT_TX;
                if (w == ww && 1 + bx * CT_TX + lane <= g.nx) {
                    double sum = ww == 0 ? 0. : s.tot[ww - 1][lane];
#pragma unroll
                    for (int l = 0; l < CT_L; l++)
                        if (w * CT_L + l < nz) {
                            sum += CT_P(fd)[l];
                            CT_P(fd)[l] = sum;
                        }
                    s.tot[ww][lane] = sum;
                }
            CT_PHASE_END
        }
        // ---- column pressure tendency and vertical wind -----------------------------------
        CT_PHASE(nt)
            const int w = tid / CT_TX, lane = tid % CT_TX, i = 1 + bx * CT_TX + lane;
            if (i <= g.nx) {
                const double dcdt = -s.tot[nw - 1][lane];
                const double cnew = CT_P(c_old) + g.dt * dcdt;
                if (w == 0) {
                    dCOLPdt[g.idx2(i, j)] = dcdt;
                    put_mass(g, COLP_NEW, i, j, 0, cnew);
                }
                const Div cn = mkdiv(cnew);
                const bool edge = (i == 1) || (i == g.nx) || (j == 1) || (j == g.ny);
                const size_t o_w = g.idx(i, j, 0);
#pragma unroll
                for (int l = 0; l < CT_L; l++) {
                    const int k = w * CT_L + l + 1;   // interface below level k-1
                    if (k < nz) {
                        const double ww = (-CT_P(fd)[l] / cn - g.sigma_vb[k] * dcdt / cn);
                        if (edge)
                            put_mass(g, WWIND, i, j, k, ww);
                        else
                            WWIND[o_w + (size_t)k * g.plane] = ww;
                    }
                }
            }
        CT_PHASE_END
    }
};

}  // namespace dc
