// dc_fused.h -- the fused stage kernel of the Matsuno step.
//
// One kernel advances U, V and POTT by one Matsuno stage: it replaces the reference's
// UVFLX_prep_adv + UFLX_tendency + VFLX_tendency + POTT_tendency + make_timestep launches
// (dyn_org_discretizations.py:121-273, :359-393) and never materialises UFLX/VFLX, the eight
// auxiliary momentum fluxes, WWIND_UWIND/WWIND_VWIND or the tendencies in HBM.
//
//   thread block = a (TX x TY) tile of (lon, lat) columns, marched through the sigma levels.
//   The (tile + halo) planes of U, V, WWIND, PHI, POTT, PVTF, PVTFVB of level k+1 are copied
//   global -> shared ASYNCHRONOUSLY (cp.async, double buffered) while level k is computed:
//     A) UFLX, VFLX and COLP_NEW*A*WWIND(k+1) of the staged region -> shared memory
//     B) the eight auxiliary fluxes B..T, once per cell, in shared memory
//     C) every thread adds up dUFLXdt, dVFLXdt, dPOTTdt of its cell (vertical momentum flux
//        of interface k carried in registers), applies the pressure-weighted Euler step and
//        stores the new U, V, POTT with their boundary images.
//
// Every expression keeps the reference's evaluation order (dc_point.h), so the result is
// bit-identical to the one-kernel-per-reference-kernel mode.  The block body is written
// against a tiny SPMD macro layer so that tests/emu can execute the very same code on the
// host (a "phase" is a loop over the block's threads there, thread-private state lives in
// arrays, an asynchronous copy is a plain copy).
#pragma once
#include "dc_geom.h"
#include "dc_kernels.h"
#include "dc_point.h"

#if defined(__CUDA_ARCH__)
#include <cuda_pipeline.h>
#endif

namespace dc {

#ifndef DC_TY
#define DC_TY 8
#endif
#ifndef DC_MINBLOCKS
#define DC_MINBLOCKS 2
#endif
constexpr int TX = 32, TY = DC_TY, NT = TX * TY;
constexpr int SW = TX + 3, SH = TY + 3;  // staged region: ri in [-1, TX+1], rj in [-1, TY+1]
constexpr int SN = SW * SH;              // cells of a y-staggered plane (V, VFLX)
constexpr int SNM = SW * (SH - 1);       // every other plane: rows rj in [-1, TY]
constexpr int NQ = (SN + NT - 1) / NT;   // staged cells per thread
#ifndef DC_NBUF
#define DC_NBUF 2
#endif
constexpr int NBUF = DC_NBUF;            // depth of the cp.async plane pipeline
constexpr int NZMAX = 128;               // fused path: nz <= NZMAX (checked by the launcher)

struct StageSmem {
    // raw planes, NBUF-fold buffered (filled by cp.async NBUF-1 levels ahead)
    double rU[NBUF][SNM], rV[NBUF][SN], rW[NBUF][SNM], rPHI[NBUF][SNM], rT[NBUF][SNM],
        rPV[NBUF][SNM], rPB[NBUF][SNM];
    // own-column scalars of a level, one slot per thread, fetched with the level's planes:
    // U[k+1], V[k+1], POTTVB[k+1] and the step-start U, V, POTT of level k
    double oU1[NBUF][NT], oV1[NBUF][NT], oTB1[NBUF][NT], oUo[NBUF][NT], oVo[NBUF][NT],
        oTo[NBUF][NT];
    // derived planes of the current level
    double UF[SNM], VF[SN], P[SNM];
    double B[SNM], C[SNM], D[SNM], E[SNM], R[SNM], Q[SNM], S[SNM], T[SNM];
    // per-level constants (broadcast reads): dsigma, 1/dsigma, sigma_vb, UVFLX / POTT diffusion
    // coefficients, 1/(dsigma[k]+dsigma[k-1])
    double lev[6][NZMAX + 1];
    // per-row constants of rows J0-1 .. J0+TY-1: cor_fcos(corf_is, cos_is), sin_is,
    // cor_fcos(corf, cos), sin, dxjs, A, 1/A
    double row[7][TY + 1];
};

#if defined(__CUDA_ARCH__)
#define DC_PRIV(type, name) type name
#define DC_PRIVN(type, name, n) type name[n]
#define DC_P(name) name
#define DC_PHASE {                                           \
        const int tid = threadIdx.y * TX + threadIdx.x;
#define DC_PHASE_END \
    }                \
    __syncthreads();
#define DC_PHASE_END_NOSYNC }
#define DC_ASYNC_COPY8(dst, src) __pipeline_memcpy_async((dst), (src), 8)
#define DC_ASYNC_COMMIT() __pipeline_commit()
#define DC_ASYNC_WAIT_AND_SYNC(n) \
    __pipeline_wait_prior(n);     \
    __syncthreads();
#else
#define DC_PRIV(type, name) type name[NT]
#define DC_PRIVN(type, name, n) type name[NT][n]
#define DC_P(name) name[tid]
#define DC_PHASE for (int tid = 0; tid < NT; tid++) {
#define DC_PHASE_END }
#define DC_PHASE_END_NOSYNC }
#define DC_ASYNC_COPY8(dst, src) (*(dst) = *(src))
#define DC_ASYNC_COMMIT()
#define DC_ASYNC_WAIT_AND_SYNC(n)
#endif

struct StageBody {
    Geom g;
    // state the tendencies are evaluated at, and its diagnostics
    const double *UWIND, *VWIND, *POTT, *PHI, *PVTF, *PVTFVB, *POTTVB, *WWIND;
    const double *COLP, *COLP_NEW, *COLP_OLD;
    // state at the beginning of the step (dyn_matsuno.py:34-49)
    const double *UWIND_OLD, *VWIND_OLD, *POTT_OLD;
    // result of the stage
    double *UWIND_out, *VWIND_out, *POTT_out;
    int j_lo, j_hi;  // global mass rows to advance

    DC_HD int wrap_i(int i) const { return i < 1 ? i + g.nx : (i > g.nx ? i - g.nx : i); }
    DC_HD static int sidx(int ri, int rj) { return (rj + 1) * SW + (ri + 1); }

    // Tiles that touch no domain edge (periodic seam columns 1, 2, nx; wall rows 1, ny; rows
    // beyond the launch range) run a specialisation without the boundary-image, wall and
    // validity branches.
    DC_HD void run_block(int bx, int by, StageSmem &s) const
    {
        const int I0 = 1 + bx * TX, J0 = j_lo + by * TY;
        const int j_top = (j_hi < g.ny - 1) ? j_hi : g.ny - 1;
        const bool interior = (I0 >= 3) && (I0 + TX - 1 <= g.nx - 1) && (J0 >= 2) &&
                              (J0 + TY - 1 <= j_top);
        if (interior)
            run<false>(bx, by, s);
        else
            run<true>(bx, by, s);
    }

    template <bool EDGE>
    DC_HD void run(int bx, int by, StageSmem &s) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz, NI = g.NI;
        const int I0 = 1 + bx * TX, J0 = j_lo + by * TY;
        const size_t plane = g.plane;
        const double dyis = g.dyis, dt = g.dt;
        const double scale = cor_scale(g.dlon_rad, g.dlat_rad);
        // rows this rank holds (global): halo rows included
        const int j_min = (g.j0 - HJ < 0) ? 0 : g.j0 - HJ;
        const int j_max_m = (g.j1 + HJ > ny + 1) ? ny + 1 : g.j1 + HJ;          // mass rows
        const int j_max_y = (g.j1 + HJ + 1 > ny + 2) ? ny + 2 : g.j1 + HJ + 1;  // y-staggered

        // ---- thread-private state --------------------------------------------------------
        DC_PRIVN(int, offM, NQ);     // plane offsets of this thread's staged cells: mass /
        DC_PRIVN(int, offV, NQ);     //   x-staggered fields, and y-staggered fields
        DC_PRIVN(double, cu, NQ);    // (COLP[i-1,j] + COLP[i,j]) / 2
        DC_PRIVN(double, cv, NQ);    // (COLP[i,j-1] + COLP[i,j]) / 2
        DC_PRIVN(double, dxv, NQ);   // dxjs[j]
        DC_PRIVN(double, cp, NQ);    // COLP_NEW[i,j] * A[j]
        DC_PRIV(int, off0);          // plane offset of the own cell
        DC_PRIV(int, flags);         // bit 0: cell is advanced; bit 1: cell has boundary images
        DC_PRIV(double, c);
        DC_PRIV(double, c_im1);
        DC_PRIV(double, c_ip1);
        DC_PRIV(double, c_jm1);
        DC_PRIV(double, c_jp1);
        DC_PRIV(double, cnew);
        DC_PRIV(double, cold);
        DC_PRIV(double, colpa_is);
        DC_PRIV(double, colpa_old_is);
        DC_PRIV(double, colpa_js);
        DC_PRIV(double, colpa_old_js);
        DC_PRIV(double, r_colpa_is);  // reciprocals: DC_FAST_MATH only (dead otherwise)
        DC_PRIV(double, r_colpa_js);
        DC_PRIV(double, r_cnew);
        DC_PRIV(double, wwu_k);      // WWIND_UWIND at interface k (carried)
        DC_PRIV(double, wwv_k);
        DC_PRIV(double, w_k);        // WWIND[k]
        DC_PRIV(double, pottvb_k);
        DC_PRIV(double, pvb);        // PVTFVB[k] at (i,j), (i-1,j), (i,j-1)
        DC_PRIV(double, pvb_im1);
        DC_PRIV(double, pvb_jm1);

        // ---- set-up -----------------------------------------------------------------------
        DC_PHASE
            // 1) plane offsets of the staged cells, then start the copy pipeline right away so
            //    that the first planes fly while the column constants below are gathered
            for (int q = 0; q < NQ; q++) {
                const int idx = tid + q * NT;
                const int ri = idx % SW - 1, rj = idx / SW - 1;
                int i = I0 + ri, j = J0 + rj;
                if (i > nx + 2) i = nx + 2;   // columns beyond the domain: masked threads only
                if (j < j_min) j = j_min;
                const int iw = wrap_i(i);
                const int jm = j > j_max_m ? j_max_m : j;   // row in a mass / x-staggered field
                const int jy = j > j_max_y ? j_max_y : j;   // row in a y-staggered field
                DC_P(offM)[q] = (int)g.idx2(iw, jm);
                DC_P(offV)[q] = (int)g.idx2(iw, jy);
            }
            {
                const int i = I0 + tid % TX, j = J0 + tid / TX;
                DC_P(off0) = (int)g.idx2(i <= nx ? i : nx, j <= j_hi ? j : j_hi);
            }
            // prologue of the copy pipeline: levels 0 .. NBUF-2 -> buffers 0 .. NBUF-2
            for (int kk = 0; kk < NBUF - 1; kk++) {
                if (kk < nz) {
                    const size_t kn = (size_t)kk * plane;
                    for (int q = 0; q < NQ; q++) {
                        const int idx = tid + q * NT;
                        if (idx < SN) DC_ASYNC_COPY8(&s.rV[kk][idx], VWIND + kn + DC_P(offV)[q]);
                        if (idx < SNM) {
                            const size_t om = kn + DC_P(offM)[q];
                            DC_ASYNC_COPY8(&s.rU[kk][idx], UWIND + om);
                            DC_ASYNC_COPY8(&s.rW[kk][idx], WWIND + plane + om);
                            DC_ASYNC_COPY8(&s.rPHI[kk][idx], PHI + om);
                            DC_ASYNC_COPY8(&s.rT[kk][idx], POTT + om);
                            DC_ASYNC_COPY8(&s.rPV[kk][idx], PVTF + om);
                            DC_ASYNC_COPY8(&s.rPB[kk][idx], PVTFVB + plane + om);
                        }
                    }
                    const size_t o = kn + DC_P(off0);
                    if (kk + 1 < nz) {
                        DC_ASYNC_COPY8(&s.oU1[kk][tid], UWIND + o + plane);
                        DC_ASYNC_COPY8(&s.oV1[kk][tid], VWIND + o + plane);
                    }
                    DC_ASYNC_COPY8(&s.oTB1[kk][tid], POTTVB + o + plane);
                    DC_ASYNC_COPY8(&s.oUo[kk][tid], UWIND_OLD + o);
                    DC_ASYNC_COPY8(&s.oVo[kk][tid], VWIND_OLD + o);
                    DC_ASYNC_COPY8(&s.oTo[kk][tid], POTT_OLD + o);
                }
                DC_ASYNC_COMMIT();
            }
            // 2) flux coefficients of the staged cells
            for (int q = 0; q < NQ; q++) {
                const int idx = tid + q * NT;
                const int ri = idx % SW - 1, rj = idx / SW - 1;
                int i = I0 + ri, j = J0 + rj;
                if (i > nx + 2) i = nx + 2;   // columns beyond the domain: masked threads only
                if (j < j_min) j = j_min;
                // Columns are wrapped into the interior [1, nx]: the x halo cells of the inputs
                // hold exactly these periodic images (misc_boundaries.py:26-32), so the staged
                // values are the ones the reference reads, and never-initialised halo cells
                // (UWIND[nxs+1] of the python set-up BC, main_grid.py:340-343) are not touched.
                const int iw = wrap_i(i), iwm = wrap_i(i - 1);
                const int jm = j > j_max_m ? j_max_m : j;   // row in a mass / x-staggered field
                const int jy = j > j_max_y ? j_max_y : j;   // row in a y-staggered field
                DC_P(offM)[q] = (int)g.idx2(iw, jm);
                DC_P(offV)[q] = (int)g.idx2(iw, jy);
                // UFLX = (C[i-1] + C[i])/2 * U * dyis     (dyn_continuity.py:40-41)
                DC_P(cu)[q] = (COLP[g.idx2(iwm, jm)] + COLP[g.idx2(iw, jm)]) / 2.;
                // VFLX = (C[j-1] + C[j])/2 * V * dxjs     (dyn_continuity.py:43-44)
                const int jc = jy > j_max_m ? j_max_m : jy;
                const int jcm = (jy - 1 < j_min) ? j_min : (jy - 1 > j_max_m ? j_max_m : jy - 1);
                DC_P(cv)[q] = (COLP[g.idx2(iw, jcm)] + COLP[g.idx2(iw, jc)]) / 2.;
                DC_P(dxv)[q] = g.dxjs[g.row(jy)];
                // COLP_NEW * A * WWIND                    (dyn_functions.py:254-260)
                DC_P(cp)[q] = COLP_NEW[g.idx2(iw, jm)] * g.A[g.row(jm)];
            }
            {
                const int tx = tid % TX, ty = tid / TX;
                const int i = I0 + tx, j = J0 + ty;
                const int valid = (i <= nx) && (j <= j_hi);
                const int ii = i <= nx ? i : nx, jj = j <= j_hi ? j : j_hi;
                const int edge = (ii <= 2) || (ii == nx) || (jj == 1) || (jj == ny);
                DC_P(flags) = valid | (edge << 1);
                DC_P(off0) = (int)g.idx2(ii, jj);
                const double *C = COLP, *CN = COLP_NEW, *CO = COLP_OLD;
                DC_P(c) = C[g.idx2(ii, jj)];
                DC_P(c_im1) = C[g.idx2(ii - 1, jj)];
                DC_P(c_ip1) = C[g.idx2(ii + 1, jj)];
                DC_P(c_jm1) = C[g.idx2(ii, jj - 1)];
                DC_P(c_jp1) = C[g.idx2(ii, jj + 1)];
                const double A = g.A[g.row(jj)], A_jm1 = g.A[g.row(jj - 1)],
                             A_jp1 = g.A[g.row(jj + 1)];
                // the Euler step runs after COLP <- COLP_NEW (dyn_matsuno.py:64-67)
                DC_P(cnew) = CN[g.idx2(ii, jj)];
                DC_P(cold) = CO[g.idx2(ii, jj)];
                DC_P(colpa_is) = interp_COLPA_is(
                    CN[g.idx2(ii, jj)], CN[g.idx2(ii - 1, jj)], CN[g.idx2(ii, jj - 1)],
                    CN[g.idx2(ii, jj + 1)], CN[g.idx2(ii - 1, jj + 1)], CN[g.idx2(ii - 1, jj - 1)],
                    A, A_jm1, A_jp1, jj, ny);
                DC_P(colpa_old_is) = interp_COLPA_is(
                    CO[g.idx2(ii, jj)], CO[g.idx2(ii - 1, jj)], CO[g.idx2(ii, jj - 1)],
                    CO[g.idx2(ii, jj + 1)], CO[g.idx2(ii - 1, jj + 1)], CO[g.idx2(ii - 1, jj - 1)],
                    A, A_jm1, A_jp1, jj, ny);
                DC_P(colpa_js) = interp_COLPA_js(
                    CN[g.idx2(ii, jj)], CN[g.idx2(ii, jj - 1)], CN[g.idx2(ii - 1, jj)],
                    CN[g.idx2(ii + 1, jj)], CN[g.idx2(ii + 1, jj - 1)], CN[g.idx2(ii - 1, jj - 1)],
                    A, A_jm1);
                DC_P(colpa_old_js) = interp_COLPA_js(
                    CO[g.idx2(ii, jj)], CO[g.idx2(ii, jj - 1)], CO[g.idx2(ii - 1, jj)],
                    CO[g.idx2(ii + 1, jj)], CO[g.idx2(ii + 1, jj - 1)], CO[g.idx2(ii - 1, jj - 1)],
                    A, A_jm1);
                DC_P(r_colpa_is) = DC_FAST ? 1. / DC_P(colpa_is) : 0.;
                DC_P(r_colpa_js) = DC_FAST ? 1. / DC_P(colpa_js) : 0.;
                DC_P(r_cnew) = DC_FAST ? 1. / DC_P(cnew) : 0.;
                DC_P(wwu_k) = 0.;  // WWIND_UWIND[0] = 0 (dyn_functions.py:236-237)
                DC_P(wwv_k) = 0.;
                DC_P(w_k) = WWIND[DC_P(off0)];
                DC_P(pottvb_k) = POTTVB[DC_P(off0)];
                DC_P(pvb) = PVTFVB[DC_P(off0)];
                DC_P(pvb_im1) = PVTFVB[DC_P(off0) - 1];
                DC_P(pvb_jm1) = PVTFVB[DC_P(off0) - NI];
                // Touch every directly loaded value once HERE ("+ 0." is an exact no-op for
                // these finite, non-negative-zero quantities): the level loop then depends on
                // arithmetic results only.  Otherwise the first use inside the loop keeps a
                // static wait on the load's scoreboard, which the cp.async copies share --
                // every level would drain the prefetch of the next level (ncu: one DMUL with
                // 15 % of all stall samples).
                DC_P(c) += 0.; DC_P(c_im1) += 0.; DC_P(c_ip1) += 0.; DC_P(c_jm1) += 0.;
                DC_P(c_jp1) += 0.; DC_P(cnew) += 0.; DC_P(cold) += 0.; DC_P(w_k) += 0.;
                DC_P(pottvb_k) += 0.; DC_P(pvb) += 0.; DC_P(pvb_im1) += 0.; DC_P(pvb_jm1) += 0.;
                for (int q = 0; q < NQ; q++) DC_P(dxv)[q] += 0.;
            }
            // constant tables
            for (int k = tid; k <= nz; k += NT) {
                s.lev[0][k] = k < nz ? g.dsigma[k] : 0.;
                s.lev[1][k] = k < nz ? g.r_dsigma[k] : 0.;
                s.lev[2][k] = g.sigma_vb[k];
                s.lev[3][k] = k < nz ? g.UVFLX_dif_coef[k] : 0.;
                s.lev[4][k] = k < nz ? g.POTT_dif_coef[k] : 0.;
                s.lev[5][k] = k < nz ? g.r_dss[k] : 0.;
            }
            if (tid <= TY) {
                int j = J0 - 1 + tid;
                if (j > j_max_m) j = j_max_m;
                const int r = g.row(j);
                s.row[0][tid] = cor_fcos(g.corf_is[r], g.cos_lat_is[r]);
                s.row[1][tid] = g.sin_lat_is[r];
                s.row[2][tid] = cor_fcos(g.corf[r], g.cos_lat[r]);
                s.row[3][tid] = g.sin_lat[r];
                s.row[4][tid] = g.dxjs[r];
                s.row[5][tid] = g.A[r];
                s.row[6][tid] = g.r_A[r];
            }
        DC_PHASE_END_NOSYNC

        for (int k = 0; k < nz; k++) {
            const int b = k % NBUF;
            const size_t ko = (size_t)k * plane;
            const bool last = (k + 1 == nz);
            // ---- prefetch: planes of level k+1 -> buffer 1-b; own-column scalars of level k
            DC_PHASE
                {   // planes of level k+NBUF-1 -> the buffer level k-1 has just released; one
                    // (possibly empty) group per level keeps the group count uniform
                    const int kp = k + NBUF - 1, bp = kp % NBUF;
                    if (kp < nz) {
                        const size_t kn = (size_t)kp * plane;
                        for (int q = 0; q < NQ; q++) {
                            const int idx = tid + q * NT;
                            if (idx < SN)
                                DC_ASYNC_COPY8(&s.rV[bp][idx], VWIND + kn + DC_P(offV)[q]);
                            if (idx < SNM) {
                                const size_t om = kn + DC_P(offM)[q];
                                DC_ASYNC_COPY8(&s.rU[bp][idx], UWIND + om);
                                DC_ASYNC_COPY8(&s.rW[bp][idx], WWIND + plane + om);
                                DC_ASYNC_COPY8(&s.rPHI[bp][idx], PHI + om);
                                DC_ASYNC_COPY8(&s.rT[bp][idx], POTT + om);
                                DC_ASYNC_COPY8(&s.rPV[bp][idx], PVTF + om);
                                DC_ASYNC_COPY8(&s.rPB[bp][idx], PVTFVB + plane + om);
                            }
                        }
                        const size_t o = kn + DC_P(off0);
                        if (kp + 1 < nz) {
                            DC_ASYNC_COPY8(&s.oU1[bp][tid], UWIND + o + plane);
                            DC_ASYNC_COPY8(&s.oV1[bp][tid], VWIND + o + plane);
                        }
                        DC_ASYNC_COPY8(&s.oTB1[bp][tid], POTTVB + o + plane);
                        DC_ASYNC_COPY8(&s.oUo[bp][tid], UWIND_OLD + o);
                        DC_ASYNC_COPY8(&s.oVo[bp][tid], VWIND_OLD + o);
                        DC_ASYNC_COPY8(&s.oTo[bp][tid], POTT_OLD + o);
                    }
                    DC_ASYNC_COMMIT();
                }
            DC_PHASE_END_NOSYNC
            DC_ASYNC_WAIT_AND_SYNC(NBUF - 1)   // level k has landed (NBUF-1 younger groups may fly)
            // ---- A: UFLX, VFLX of level k and COLP_NEW*A*WWIND of interface k+1 ----------
            DC_PHASE
                for (int q = 0; q < NQ; q++) {
                    const int idx = tid + q * NT;
                    if (idx < SN)
                        s.VF[idx] = DC_P(cv)[q] * s.rV[b][idx] * DC_P(dxv)[q];  // calc_VFLX
                    if (idx < SNM) {
                        s.UF[idx] = DC_P(cu)[q] * s.rU[b][idx] * dyis;          // calc_UFLX
                        s.P[idx] = DC_P(cp)[q] * s.rW[b][idx];
                    }
                }
            DC_PHASE_END
            // ---- B: auxiliary momentum fluxes (dyn_functions.py:429-536), once per cell --
            DC_PHASE
                const double *u = s.UF, *v = s.VF;
                {   // own cell: every flux is inside the staged region, no guards
                    const int c0 = sidx(tid % TX, tid / TX);
                    s.B[c0] = calc_BFLX(u[c0 - SW], u[c0 - SW + 1], u[c0], u[c0 + 1], u[c0 + SW],
                                        u[c0 + SW + 1]);
                    s.C[c0] = calc_CFLX(v[c0 - SW - 1], v[c0 - SW], v[c0 - 1], v[c0],
                                        v[c0 + SW - 1], v[c0 + SW]);
                    s.D[c0] = calc_DFLX(v[c0 - SW], v[c0], v[c0 + SW], u[c0 - SW], u[c0],
                                        u[c0 - SW + 1], u[c0 + 1]);
                    s.E[c0] = calc_EFLX(v[c0 - SW], v[c0], v[c0 + SW], u[c0 - SW], u[c0],
                                        u[c0 - SW + 1], u[c0 + 1]);
                    s.R[c0] = calc_RFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW], v[c0 + 1],
                                        v[c0 + SW + 1]);
                    s.Q[c0] = calc_QFLX(u[c0 - SW - 1], u[c0 - 1], u[c0 - SW], u[c0],
                                        u[c0 - SW + 1], u[c0 + 1]);
                    s.S[c0] = calc_SFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW], u[c0 - 1],
                                        u[c0], u[c0 + 1]);
                    s.T[c0] = calc_TFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW], u[c0 - 1],
                                        u[c0], u[c0 + 1]);
                }
                // frame around the tile: column ri = -1 / TX and row rj = -1 / TY
                if (tid < 2 * (TX + 2) + 2 * TY) {
                    int ri, rj;
                    if (tid < TX + 2) {
                        ri = tid - 1;
                        rj = -1;
                    } else if (tid < 2 * (TX + 2)) {
                        ri = tid - (TX + 2) - 1;
                        rj = TY;
                    } else if (tid < 2 * (TX + 2) + TY) {
                        ri = -1;
                        rj = tid - 2 * (TX + 2);
                    } else {
                        ri = TX;
                        rj = tid - 2 * (TX + 2) - TY;
                    }
                    const int c0 = sidx(ri, rj);
                    if (ri <= TX - 1 && rj >= 0 && rj <= TY - 1)
                        s.B[c0] = calc_BFLX(u[c0 - SW], u[c0 - SW + 1], u[c0], u[c0 + 1],
                                            u[c0 + SW], u[c0 + SW + 1]);
                    if (ri >= 0 && ri <= TX - 1 && rj >= 0)
                        s.C[c0] = calc_CFLX(v[c0 - SW - 1], v[c0 - SW], v[c0 - 1], v[c0],
                                            v[c0 + SW - 1], v[c0 + SW]);
                    if (ri <= TX - 1 && rj >= 0) {
                        s.D[c0] = calc_DFLX(v[c0 - SW], v[c0], v[c0 + SW], u[c0 - SW], u[c0],
                                            u[c0 - SW + 1], u[c0 + 1]);
                        s.E[c0] = calc_EFLX(v[c0 - SW], v[c0], v[c0 + SW], u[c0 - SW], u[c0],
                                            u[c0 - SW + 1], u[c0 + 1]);
                    }
                    if (ri >= 0 && ri <= TX - 1 && rj <= TY - 1)
                        s.R[c0] = calc_RFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW],
                                            v[c0 + 1], v[c0 + SW + 1]);
                    if (ri >= 0 && rj >= 0 && rj <= TY - 1)
                        s.Q[c0] = calc_QFLX(u[c0 - SW - 1], u[c0 - 1], u[c0 - SW], u[c0],
                                            u[c0 - SW + 1], u[c0 + 1]);
                    if (ri >= 0 && rj <= TY - 1) {
                        s.S[c0] = calc_SFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW],
                                            u[c0 - 1], u[c0], u[c0 + 1]);
                        s.T[c0] = calc_TFLX(v[c0 - 1], v[c0 + SW - 1], v[c0], v[c0 + SW],
                                            u[c0 - 1], u[c0], u[c0 + 1]);
                    }
                }
            DC_PHASE_END
            // ---- C: tendencies of the own cell, Euler step, store with boundary images ----
            DC_PHASE
                const int tx = tid % TX, ty = tid / TX;
                const int i = I0 + tx, j = J0 + ty;
                const int c0 = sidx(tx, ty);
                const size_t o = ko + DC_P(off0);
                const double ds = s.lev[0][k];
                const Div ds_d = mkdiv(ds, s.lev[1][k]);
                const double w_kp1 = s.rW[b][c0];
                const double pottvb_kp1 = s.oTB1[b][tid];
                const double pvb_kp1 = s.rPB[b][c0];
                const double pvb_im1_kp1 = s.rPB[b][c0 - 1];
                const double pvb_jm1_kp1 = s.rPB[b][c0 - SW];
                if (!EDGE || (DC_P(flags) & 1)) {
                    const double *U = s.rU[b], *V = s.rV[b], *T = s.rT[b];
                    const double u = U[c0], v = V[c0];
                    const bool edge = EDGE && (DC_P(flags) & 2);
                    const bool wall_s = EDGE && (j == 1), wall_n = EDGE && (j == ny);
                    // vertical momentum fluxes through interface k+1
                    // (dyn_functions.py:211-270; 0 at the model bottom)
                    double wwu_kp1 = 0., wwv_kp1 = 0.;
                    if (!last) {
                        const double *P = s.P;
                        const double ds_kp1 = s.lev[0][k + 1];
                        const Div dss_d = mkdiv(ds_kp1 + ds, s.lev[5][k + 1]);
                        const int wall = wall_s ? -1 : (wall_n ? 1 : 0);
                        wwu_kp1 = colpa_wwind(P[c0], P[c0 - 1], P[c0 - SW], P[c0 + SW],
                                              P[c0 - SW - 1], P[c0 + SW - 1], wall) *
                                  interp_ks(s.oU1[b][tid], u, ds_kp1, ds, dss_d);
                        wwv_kp1 = colpa_wwind(P[c0], P[c0 - SW], P[c0 - 1], P[c0 + 1],
                                              P[c0 - SW - 1], P[c0 - SW + 1], 0) *
                                  interp_ks(s.oV1[b][tid], v, ds_kp1, ds, dss_d);
                    }
                    const double phi = s.rPHI[b][c0], pott = T[c0], pvtf = s.rPV[b][c0];
                    // ---------------- dUFLXdt (dyn_UFLX.py:69-199) ----------------
                    {
                        double bflx = s.B[c0], bflx_im1 = s.B[c0 - 1];
                        double cflx = s.C[c0], cflx_jp1 = s.C[c0 + SW];
                        double dflx_im1 = s.D[c0 - 1], dflx_jp1 = s.D[c0 + SW];
                        double eflx = s.E[c0], eflx_im1_jp1 = s.E[c0 + SW - 1];
                        if (wall_s) {
                            dflx_im1 = 0.;
                            cflx = 0.;
                            eflx = 0.;
                        }
                        if (wall_n) {
                            dflx_jp1 = 0.;
                            cflx_jp1 = 0.;
                            eflx_im1_jp1 = 0.;
                        }
                        double d = 0.;
                        d = d + UVFLX_hor_adv(u, U[c0 - 1], U[c0 + 1], U[c0 - SW], U[c0 + SW],
                                              U[c0 - SW - 1], U[c0 + SW - 1], U[c0 - SW + 1],
                                              U[c0 + SW + 1], bflx, bflx_im1, cflx, cflx_jp1,
                                              dflx_im1, dflx_jp1, eflx, eflx_im1_jp1, 1.);
                        d = d + ((DC_P(wwu_k) - wwu_kp1) / ds_d);
                        d = d + coriolis_UWIND(DC_P(c), DC_P(c_im1), v, V[c0 - 1], V[c0 + SW],
                                               V[c0 + SW - 1], u, U[c0 - 1], U[c0 + 1],
                                               s.row[0][ty + 1], s.row[1][ty + 1], scale);
                        d = d + pre_grad(phi, s.rPHI[b][c0 - 1], DC_P(c), DC_P(c_im1), pott,
                                         T[c0 - 1], pvtf, s.rPV[b][c0 - 1], DC_P(pvb),
                                         DC_P(pvb_im1), pvb_im1_kp1, pvb_kp1, ds_d, s.lev[2][k],
                                         s.lev[2][k + 1], dyis);
                        const double coef = s.lev[3][k];
                        if (coef > 0.)
                            d = d + num_dif(s.UF[c0], s.UF[c0 - 1], s.UF[c0 + 1], s.UF[c0 - SW],
                                            s.UF[c0 + SW], coef);
                        const double un = euler_forward_pw(
                            s.oUo[b][tid], d, mkdiv(DC_P(colpa_is), DC_P(r_colpa_is)),
                            DC_P(colpa_old_is), dt);
                        if (edge)
                            put_xstag(g, UWIND_out, i, j, k, un);
                        else
                            UWIND_out[o] = un;
                    }
                    // ---------------- dVFLXdt (dyn_VFLX.py:67-198) ----------------
                    if (!EDGE || j >= 2) {
                        double d = 0.;
                        d = d + UVFLX_hor_adv(v, V[c0 - SW], V[c0 + SW], V[c0 - 1], V[c0 + 1],
                                              V[c0 - SW - 1], V[c0 - SW + 1], V[c0 + SW - 1],
                                              V[c0 + SW + 1], s.R[c0], s.R[c0 - SW], s.Q[c0],
                                              s.Q[c0 + 1], s.S[c0 - SW], s.S[c0 + 1], s.T[c0],
                                              s.T[c0 - SW + 1], -1.);
                        d = d + ((DC_P(wwv_k) - wwv_kp1) / ds_d);
                        d = d + coriolis_VWIND(DC_P(c), DC_P(c_jm1), u, U[c0 - SW], U[c0 + 1],
                                               U[c0 - SW + 1], s.row[2][ty + 1], s.row[3][ty + 1],
                                               s.row[2][ty], s.row[3][ty], scale);
                        d = d + pre_grad(phi, s.rPHI[b][c0 - SW], DC_P(c), DC_P(c_jm1), pott,
                                         T[c0 - SW], pvtf, s.rPV[b][c0 - SW], DC_P(pvb),
                                         DC_P(pvb_jm1), pvb_jm1_kp1, pvb_kp1, ds_d, s.lev[2][k],
                                         s.lev[2][k + 1], s.row[4][ty + 1]);
                        const double coef = s.lev[3][k];
                        if (coef > 0.)
                            d = d + num_dif(s.VF[c0], s.VF[c0 - 1], s.VF[c0 + 1], s.VF[c0 - SW],
                                            s.VF[c0 + SW], coef);
                        const double vn = euler_forward_pw(
                            s.oVo[b][tid], d, mkdiv(DC_P(colpa_js), DC_P(r_colpa_js)),
                            DC_P(colpa_old_js), dt);
                        if (edge)
                            put_ystag(g, VWIND_out, i, j, k, vn);
                        else
                            VWIND_out[o] = vn;
                    } else {
                        put_ystag(g, VWIND_out, i, 1, k, 0.);
                    }
                    if (wall_n) put_ystag(g, VWIND_out, i, ny + 1, k, 0.);
                    // ---------------- dPOTTdt (dyn_POTT.py:55-110) ----------------
                    {
                        const double p_im1 = T[c0 - 1], p_ip1 = T[c0 + 1];
                        const double p_jm1 = T[c0 - SW], p_jp1 = T[c0 + SW];
                        double d = 0.;
                        d = d + hor_adv(pott, p_im1, p_ip1, p_jm1, p_jp1, s.UF[c0], s.UF[c0 + 1],
                                        s.VF[c0], s.VF[c0 + SW],
                                        mkdiv(s.row[5][ty + 1], s.row[6][ty + 1]));
                        d = d + vert_adv(DC_P(pottvb_k), pottvb_kp1, DC_P(w_k), w_kp1,
                                         DC_P(cnew), ds_d, k);
                        const double coef = s.lev[4][k];
                        if (coef > 0.)
                            d = d + num_dif_pw(pott, p_im1, p_ip1, p_jm1, p_jp1, DC_P(c),
                                               DC_P(c_im1), DC_P(c_ip1), DC_P(c_jm1), DC_P(c_jp1),
                                               coef);
                        const double tn = euler_forward_pw(
                            s.oTo[b][tid], d, mkdiv(DC_P(cnew), DC_P(r_cnew)), DC_P(cold), dt);
                        if (edge)
                            put_mass(g, POTT_out, i, j, k, tn);
                        else
                            POTT_out[o] = tn;
                    }
                    DC_P(wwu_k) = wwu_kp1;
                    DC_P(wwv_k) = wwv_kp1;
                }
                DC_P(w_k) = w_kp1;
                DC_P(pottvb_k) = pottvb_kp1;
                DC_P(pvb) = pvb_kp1;
                DC_P(pvb_im1) = pvb_im1_kp1;
                DC_P(pvb_jm1) = pvb_jm1_kp1;
            DC_PHASE_END
        }
    }
};

}  // namespace dc
