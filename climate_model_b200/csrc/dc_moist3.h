// dc_moist3.h -- fused moisture stage as a TMA-staged tile kernel.
//
// One launch advances BOTH tracers (QV, QC) by one Matsuno stage: dQVdt / dQCdt
// (dyn_moist.py:49-129: horizontal flux-form advection with UFLX / VFLX, vertical advection
// with the logarithmic interface value comp_VARVB_log dyn_functions.py:70-95, pressure-weighted
// diffusion), the pressure-weighted Euler step (dyn_timestep.py:292-296) and the boundary
// images (misc_boundaries.py:22-42).  It replaces MoistStageBody (dc_kernels.h: a thread per
// column, global loads, one tracer per pass, 2.2 ms per stage at 0.25 deg x 64 levels) with
// the design of the dry stage kernel (dc_stage3.h), whose machinery it reuses:
//   * 32 x 8 column tiles, a thread owns two longitude-adjacent columns, the sigma column is
//     marched top-down with the interface state (log, reciprocal and interface value of the
//     level above; WWIND) carried in registers;
//   * the (tile + halo) planes of U, V, QV, QC and the own-column boxes of WWIND(k+1) and of the
//     step-start QV, QC arrive by TMA into a 3-deep ring (same boxes and descriptors as the dry
//     kernel); UFLX / VFLX are formed from the raw winds and two coefficient planes, once for
//     both tracers;
//   * 68 KB of shared memory and <= 168 registers: three blocks per SM.
// Every expression keeps the reference's evaluation order, so the strict build is bit-identical
// to the kernel decomposition; the production build only swaps log / division for log_tab /
// reciprocal multiplications (dc_point.h).
//
// Written against the SPMD layer of dc_stage3.h so that tests/emu runs the same body.
#pragma once
#include "dc_stage3.h"

namespace dc {

struct alignas(128) Moist3Smem {
    double rU[S3_NBUF][S3_PL], rV[S3_NBUF][S3_PL], rQ[2][S3_NBUF][S3_PL];   // staged planes
    double oW[S3_NBUF][S3_OWN], oQo[2][S3_NBUF][S3_OWN];                    // own-column boxes
    double CU[S3_PL], CV[S3_PL];    // (C[i-1]+C[i])/2, (C[j-1]+C[j])/2 (production: x dyis, dxjs)
    double lev[3][NZMAX + 1];       // dsigma, 1/dsigma, moist_dif_coef
    double dxr[S3_SH + 1];          // dxjs of the staged rows
    double rowA[2][S3_TY + 1];      // A, 1/A of the tile rows
    double ltab[POW_NJ][2];         // log_tab's table (dc_point.h)
    unsigned long long full[S3_NBUF];
};

struct Moist3Body {
    Geom g;
    TmaMap mU, mV, mQ[2];            // boxes S3_SW x S3_SH x 1
    TmaMap mW, mQo[2];               // boxes S3_OW x S3_TY x 1
    const double *COLP, *COLP_NEW, *COLP_OLD, *WWIND;
    const double *Q_in[2];           // the stage's input tracers (set-up reads of level ks - 1)
    double *Q_out[2];
    int j_lo, j_hi;                  // global mass rows to advance ...
    int jt;                          // ... in tiles of the GLOBAL tiling starting at row jt
                                     // (see Stage3Body: bitwise identical across decompositions)
    int have_old;
    int nkc;                         // sigma-column chunks (see Stage3Body::nkc)
    LogCoef lc;

    DC_HD int wrap_i(int i) const { return i < 1 ? i + g.nx : (i > g.nx ? i - g.nx : i); }

    DC_HD void run_block(int bx, int by, int bz, Moist3Smem &s) const
    {
        const int I0 = 1 + bx * S3_TX, J0 = jt + by * S3_TY;
        const bool interior = (I0 >= 3) && (I0 + S3_TX - 1 <= g.nx - 1) && (J0 >= 2) &&
                              (J0 + S3_TY - 1 <= g.ny - 1);
        if (interior)
            run<false>(bx, J0, bz, s);
        else
            run<true>(bx, J0, bz, s);
    }

    DC_HD void issue(Moist3Smem &s, int kp, int ks, int x0, int y0) const
    {
        const int bp = (kp - ks) % S3_NBUF;
        unsigned long long *bar = &s.full[bp];
        const unsigned bytes = 4u * S3_SN * 8u + (have_old ? 3u : 1u) * (unsigned)S3_OWN * 8u;
        s3_mbar_expect(bar, bytes);
        s3_tma_load(s.rU[bp], &mU, x0, y0, kp, bar);
        s3_tma_load(s.rV[bp], &mV, x0, y0, kp, bar);
        s3_tma_load(s.rQ[0][bp], &mQ[0], x0, y0, kp, bar);
        s3_tma_load(s.rQ[1][bp], &mQ[1], x0, y0, kp, bar);
        s3_tma_load(s.oW[bp], &mW, x0, y0 + 1, kp + 1, bar);
        if (have_old) {
            s3_tma_load(s.oQo[0][bp], &mQo[0], x0, y0 + 1, kp, bar);
            s3_tma_load(s.oQo[1][bp], &mQo[1], x0, y0 + 1, kp, bar);
        }
    }

    // clamped value, its logarithm and reciprocal (comp_VARVB_log, dyn_functions.py:70-95)
    DC_HD void clr(const Moist3Smem &s, double q, double *qc, double *lq, double *rq) const
    {
        const double min_val = 0.0000001;
        *qc = fmax(q, min_val);
        *lq = DC_FAST ? log_tab(*qc, lc, s.ltab) : log(*qc);
        *rq = DC_FAST ? dc_rcp(*qc) : 1. / *qc;
    }

    template <bool EDGE>
    DC_HD void run(int bx, int J0, int bz, Moist3Smem &s) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        const int kcl = (nz + nkc - 1) / nkc;
        const int k0 = bz * kcl, k1 = (k0 + kcl < nz) ? k0 + kcl : nz;
        const int ks = k0 > 0 ? k0 - 1 : 0, ke = k1 < nz ? k1 : nz - 1;
        const int I0 = 1 + bx * S3_TX;
        const size_t plane = g.plane;
        const double dyis = g.dyis, dt = g.dt;
        const int x0 = I0 - 1, y0 = g.row(J0 - 1);
        const int j_min = (g.j0 - HJ < 0) ? 0 : g.j0 - HJ;
        const int j_max_m = (g.j1 + HJ > ny + 1) ? ny + 1 : g.j1 + HJ;
        const int j_max_y = (g.j1 + HJ + 1 > ny + 2) ? ny + 2 : g.j1 + HJ + 1;

        S3_PRIV(int, off0);
        S3_PRIV(int, flags);
        S3_PRIV(double, c_m1);
        S3_PRIV(double, c_a);
        S3_PRIV(double, c_b);
        S3_PRIV(double, c_p1);
        S3_PRIVN(double, c_jm1, 2);
        S3_PRIVN(double, c_jp1, 2);
        S3_PRIVN(double, cnew, 2);
        S3_PRIVN(double, cold, 2);
        S3_PRIVN(double, r_cnew, 2);
        S3_PRIVN(double, w_k, 2);
        // interface state of the level above, per tracer and cell: clamped value, log,
        // reciprocal and the interface value comp_VARVB_log(Q[k], Q[k-1])
        S3_PRIVNN(double, qc, 2, 2);
        S3_PRIVNN(double, lq, 2, 2);
        S3_PRIVNN(double, rq, 2, 2);
        S3_PRIVNN(double, qvb, 2, 2);

        S3_PHASE
            if (tid == 0) {
                for (int n = 0; n < S3_NBUF; n++) s3_mbar_init(&s.full[n], 1);
                s3_mbar_init_fence();
            }
            for (int n = tid; n < POW_NJ; n += S3_NT) {   // the set-up below already takes logs
                s.ltab[n][0] = lc.tab[n][0];
                s.ltab[n][1] = lc.tab[n][1];
            }
        S3_PHASE_END
        S3_PHASE
            if (tid == 0) {
                for (int n = 0; n < S3_PF; n++)
                    if (ks + n <= ke) issue(s, ks + n, ks, x0, y0);
            }
            for (int n = tid; n < S3_PL; n += S3_NT) {
                const int r = n / S3_SW, cw = n % S3_SW;
                int i = I0 + cw - 1, j = J0 + r - 1;
                if (i > nx + 2) i = nx + 2;
                if (j < j_min) j = j_min;
                const int jm = j > j_max_m ? j_max_m : j;
                const int jy = j > j_max_y ? j_max_y : j;
                const int jc = jy > j_max_m ? j_max_m : jy;
                const int jcm = (jy - 1 < j_min) ? j_min : (jy - 1 > j_max_m ? j_max_m : jy - 1);
                const int iw = wrap_i(i), iwm = wrap_i(i - 1);
                s.CU[n] = (COLP[g.idx2(iwm, jm)] + COLP[g.idx2(iw, jm)]) / 2. * (DC_FAST ? dyis : 1.);
                s.CV[n] = (COLP[g.idx2(iw, jcm)] + COLP[g.idx2(iw, jc)]) / 2. *
                          (DC_FAST ? g.dxjs[g.row(jy)] : 1.);
            }
            if (tid <= S3_SH) {
                int j = J0 - 1 + tid;
                if (j < j_min) j = j_min;
                if (j > j_max_y) j = j_max_y;
                s.dxr[tid] = g.dxjs[g.row(j)];
            }
            {
                const int tx = tid % S3_NTX, ty = tid / S3_NTX;
                const int ia = I0 + 2 * tx, j = J0 + ty;
                const int vj = (j >= j_lo) && (j <= j_hi);
                const int va = (ia <= nx) && vj, vb = (ia + 1 <= nx) && vj;
                const int ii = ia > nx ? nx - 1 : ia, jj = j < j_lo ? j_lo : (j <= j_hi ? j : j_hi);
                const int ea = (ia <= 2) || (ia == nx) || (jj == 1) || (jj == ny);
                const int eb = (ia + 1 <= 2) || (ia + 1 == nx) || (jj == 1) || (jj == ny);
                S3_P(flags) = va | (vb << 1) | (ea << 2) | (eb << 3);
                S3_P(off0) = (int)g.idx2(ii, jj);
                S3_P(c_m1) = COLP[g.idx2(ii - 1, jj)];
                S3_P(c_a) = COLP[g.idx2(ii, jj)];
                S3_P(c_b) = COLP[g.idx2(ii + 1, jj)];
                S3_P(c_p1) = COLP[g.idx2(ii + 2, jj)];
                for (int e = 0; e < 2; e++) {
                    S3_P(c_jm1)[e] = COLP[g.idx2(ii + e, jj - 1)];
                    S3_P(c_jp1)[e] = COLP[g.idx2(ii + e, jj + 1)];
                    S3_P(cnew)[e] = COLP_NEW[g.idx2(ii + e, jj)];
                    S3_P(cold)[e] = COLP_OLD[g.idx2(ii + e, jj)];
                    S3_P(r_cnew)[e] = DC_FAST ? 1. / S3_P(cnew)[e] : 0.;
                    S3_P(w_k)[e] = WWIND[(size_t)ks * plane + S3_P(off0) + e];
                    for (int t = 0; t < 2; t++) {
                        // state of interface ks: formed from levels ks - 1 and ks by the warm-up
                        // level of a chunk; at the model top the interface value is never used
                        const double q = Q_in[t][(size_t)ks * plane + S3_P(off0) + e];
                        clr(s, q, &S3_P(qc)[t][e], &S3_P(lq)[t][e], &S3_P(rq)[t][e]);
                        S3_P(qvb)[t][e] = q;
                    }
                }
            }
            for (int k = tid; k <= nz; k += S3_NT) {
                s.lev[0][k] = k < nz ? g.dsigma[k] : 0.;
                s.lev[1][k] = k < nz ? g.r_dsigma[k] : 0.;
                s.lev[2][k] = k < nz ? g.moist_dif_coef[k] : 0.;
            }
            if (tid <= S3_TY) {
                int j = J0 - 1 + tid;
                if (j < j_min) j = j_min;
                if (j > j_max_m) j = j_max_m;
                s.rowA[0][tid] = g.A[g.row(j)];
                s.rowA[1][tid] = g.r_A[g.row(j)];
            }
            s3_mbar_wait(&s.full[0], 0);
        S3_PHASE_END

        for (int k = ks; k < k1; k++) {
            const int b = (k - ks) % S3_NBUF, b1 = (k + 1 - ks) % S3_NBUF;
            const bool warm = k < k0;
            const size_t ko = (size_t)k * plane;
            const bool last = (k + 1 == nz);
            S3_PHASE
                if (tid == (S3_ROTATE ? (k % (S3_NT / 32)) * 32 : 0) && k + S3_PF <= ke)
                    issue(s, k + S3_PF, ks, x0, y0);
                if (!last) s3_mbar_wait(&s.full[b1], ((k + 1 - ks) / S3_NBUF) & 1);
            S3_PHASE_END_NOSYNC
            S3_PHASE
                const int tx = tid % S3_NTX, ty = tid / S3_NTX;
                const int ia = I0 + 2 * tx, j = J0 + ty;
                const int b0 = (ty + 1) * S3_SW + 2 * tx;
                const int o0 = ty * S3_OW + 2 * tx + 1;
                const int fl = S3_P(flags);
                const double w_kp1[2] = {s.oW[b][o0], s.oW[b][o0 + 1]};
                if (fl & 3) {
                    const Div ds_d = mkdiv(s.lev[0][k], s.lev[1][k]);
                    const Div A_d = mkdiv(s.rowA[0][ty + 1], s.rowA[1][ty + 1]);
                    const double coef = s.lev[2][k];
                    // UFLX at ia, ia+1, ia+2 and VFLX at rows j, j+1 of the own columns
                    // (calc_UFLX / calc_VFLX, dyn_continuity.py:40-47), once for both tracers
                    const R4 U_0 = ld4(&s.rU[b][b0]), cu = ld4(&s.CU[b0]);
                    const double vr[2][2] = {{s.rV[b][b0 + 1], s.rV[b][b0 + 2]},
                                             {s.rV[b][b0 + S3_SW + 1], s.rV[b][b0 + S3_SW + 2]}};
                    const double cv[2][2] = {{s.CV[b0 + 1], s.CV[b0 + 2]},
                                             {s.CV[b0 + S3_SW + 1], s.CV[b0 + S3_SW + 2]}};
                    double uf[3], vf[2][2];
                    if (DC_FAST) {
                        uf[0] = cu.a * U_0.a; uf[1] = cu.b * U_0.b; uf[2] = cu.p1 * U_0.p1;
                        for (int r = 0; r < 2; r++)
                            for (int e = 0; e < 2; e++) vf[r][e] = cv[r][e] * vr[r][e];
                    } else {
                        uf[0] = cu.a * U_0.a * dyis; uf[1] = cu.b * U_0.b * dyis;
                        uf[2] = cu.p1 * U_0.p1 * dyis;
                        for (int r = 0; r < 2; r++)
                            for (int e = 0; e < 2; e++)
                                vf[r][e] = cv[r][e] * vr[r][e] * s.dxr[ty + 1 + r];
                    }
                    const double cc[2] = {S3_P(c_a), S3_P(c_b)};
                    const double cw[2] = {S3_P(c_m1), S3_P(c_a)};
                    const double ce[2] = {S3_P(c_b), S3_P(c_p1)};
                    for (int t = 0; t < 2; t++) {
                        const R4 Q_0 = ld4(&s.rQ[t][b][b0]);
                        const double q[2] = {Q_0.a, Q_0.b};
                        const double qw[2] = {Q_0.m1, Q_0.a}, qe[2] = {Q_0.b, Q_0.p1};
                        const double q_jm1[2] = {s.rQ[t][b][b0 - S3_SW + 1], s.rQ[t][b][b0 - S3_SW + 2]};
                        const double q_jp1[2] = {s.rQ[t][b][b0 + S3_SW + 1], s.rQ[t][b][b0 + S3_SW + 2]};
                        double q_kp1[2] = {q[0], q[1]};
                        if (!last) {
                            q_kp1[0] = s.rQ[t][b1][b0 + 1]; q_kp1[1] = s.rQ[t][b1][b0 + 2];
                        }
                        double qo[2] = {q[0], q[1]};
                        if (have_old) {
                            qo[0] = s.oQo[t][b][o0]; qo[1] = s.oQo[t][b][o0 + 1];
                        }
                        // interface k+1: comp_VARVB_log(VAR = Q[k+1], VAR_km1 = Q[k]) for both
                        // cells at once: the two log / reciprocal / division chains are
                        // independent and interleave (a per-cell branch kept them serial: ncu
                        // showed 1.8 + 1.4 stall cycles per instruction on dependencies).  A cell
                        // whose clamped values are equal takes the value itself, as the reference
                        // does; log and reciprocal of an equal value are the same numbers, so the
                        // carried state is what the branch would have left
                        double qc1v[2], lq1v[2], rq1v[2], qvb1v[2];
                        {
                            const double min_val = 0.0000001;
                            qc1v[0] = fmax(q_kp1[0], min_val); qc1v[1] = fmax(q_kp1[1], min_val);
                            const bool n0 = qc1v[0] != S3_P(qc)[t][0], n1 = qc1v[1] != S3_P(qc)[t][1];
                            for (int e = 0; e < 2; e++) {
                                lq1v[e] = S3_P(lq)[t][e]; rq1v[e] = S3_P(rq)[t][e]; qvb1v[e] = qc1v[e];
                            }
                            if (n0 | n1) {
                                for (int e = 0; e < 2; e++) clr(s, q_kp1[e], &qc1v[e], &lq1v[e], &rq1v[e]);
                                for (int e = 0; e < 2; e++) {
                                    const double num = S3_P(lq)[t][e] - lq1v[e], den = rq1v[e] - S3_P(rq)[t][e];
                                    const double v = DC_FAST ? num * dc_rcp(den) : (num / den);
                                    qvb1v[e] = (e == 0 ? n0 : n1) ? v : qc1v[e];
                                }
                            }
                        }
                        for (int e = 0; e < 2; e++) {
                            double d = 0.;
                            d = d + hor_adv(q[e], qw[e], qe[e], q_jm1[e], q_jp1[e], uf[e], uf[e + 1],
                                            vf[0][e], vf[1][e], A_d);
                            const double qc1 = qc1v[e], lq1 = lq1v[e], rq1 = rq1v[e], qvb1 = qvb1v[e];
                            d = d + vert_adv(S3_P(qvb)[t][e], qvb1, S3_P(w_k)[e], w_kp1[e],
                                             S3_P(cnew)[e], ds_d, k);
                            if (coef > 0.)
                                d = d + num_dif_pw(q[e], qw[e], qe[e], q_jm1[e], q_jp1[e], cc[e], cw[e],
                                                   ce[e], S3_P(c_jm1)[e], S3_P(c_jp1)[e], coef);
                            const double qn = euler_forward_pw(
                                qo[e], d, mkdiv(S3_P(cnew)[e], S3_P(r_cnew)[e]), S3_P(cold)[e], dt);
                            S3_P(qc)[t][e] = qc1; S3_P(lq)[t][e] = lq1; S3_P(rq)[t][e] = rq1;
                            S3_P(qvb)[t][e] = qvb1;
                            if (warm) continue;
                            if (EDGE) {
                                if (fl & (1 << e)) {
                                    if (fl & (4 << e))
                                        put_mass(g, Q_out[t], ia + e, j, k, qn);
                                    else
                                        Q_out[t][ko + S3_P(off0) + e] = qn;
                                }
                            } else if (fl & (1 << e)) {
                                Q_out[t][ko + S3_P(off0) + e] = qn;
                            }
                        }
                    }
                }
                for (int e = 0; e < 2; e++) S3_P(w_k)[e] = w_kp1[e];
            S3_PHASE_END
        }
    }
};

}  // namespace dc
