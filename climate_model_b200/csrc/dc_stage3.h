// dc_stage3.h -- the fused stage kernel of the Matsuno step, third generation.
//
// One launch advances U, V and POTT by one Matsuno stage (momentum-flux
// preparation dyn_UVFLX_prepare.py:249-439, dUFLXdt dyn_UFLX.py:69-199, dVFLXdt
// dyn_VFLX.py:67-198, dPOTTdt dyn_POTT.py:55-110, pressure-weighted Euler step
// dyn_timestep.py:212-296, boundary images misc_boundaries.py:22-42), re-designed around the
// resource ncu showed to bound the second generation: SHARED-MEMORY BANDWIDTH (128 B/clk/SM;
// 80 % busy, profiles/r1_variants.md).
//
//   * a thread owns TWO longitude-adjacent columns and reads every stencil row with 16-byte
//     shared loads: a 3-wide stencil row of two cells is 4 words instead of 6;
//   * the (tile + halo) planes of U, V, WWIND, PHI, POTT, PVTF, PVTFVB arrive by TMA
//     (cp.async.bulk.tensor, one elected thread, mbarrier completion, 3-deep ring): no
//     per-thread copy instructions, no address registers, no LSU wavefronts for the fill;
//   * the eight auxiliary momentum fluxes (B..T) are NOT exchanged through shared memory: each
//     thread evaluates the 30 flux values its two cells need from the UFLX / VFLX planes
//     (15 per cell instead of 8 + frame, but the FP64 pipe has room and the kernel loses
//     eight planes, one phase and one block barrier per level);
//   * the pressure-gradient term reads ONE plane, PGCOL
//       = POTT/dsigma * (sigma_vb[k+1]*(PVTFVB[k+1]-PVTF) + sigma_vb[k]*(PVTF-PVTFVB[k]))
//     (the reference's per-column sub-expression, dyn_functions.py:177-207), which the
//     diagnostics kernel writes instead of PVTF and PVTFVB;
//   * UFLX, VFLX and COLP_NEW*A*WWIND are formed in registers from the raw planes and three
//     coefficient planes built once per block: no derived planes, no second phase, ONE block
//     barrier per level (the ring-slot release).
//
// Per level: TMA wait -> everything in registers -> barrier (the ring-slot release; a
// barrier-free hand-over through `empty` mbarriers exists behind DC_S3_SYNC=0 and measured
// slower).  The refill of a slot is issued by one lane of warp k mod 4, so the ~130
// instructions of the ten descriptors rotate over the warps.  Every expression keeps the
// reference's evaluation order, so the strict build stays bit-identical to the
// one-kernel-per-reference-kernel mode.  The periodic longitude images are read from the x
// halo cells of the inputs (valid by construction: every producer stores its boundary
// images; the x-staggered cell nx+2 of the initial UWIND is refreshed by k_xhalo_fix).
//
// Written against the SPMD macro layer so that tests/emu runs the same body on the host
// (TMA = a box copy with zero fill, mbarrier = no-op).
#pragma once
#include <stdlib.h>

#include "dc_geom.h"
#include "dc_kernels.h"
#include "dc_point.h"

namespace dc {

constexpr int NZMAX = 128;   // the fused path holds per-level tables in shared memory: nz <= NZMAX

#ifndef DC_S3_TY
#define DC_S3_TY 8
#endif
constexpr int S3_TX = 32, S3_TY = DC_S3_TY;          // tile: 32 x TY columns
constexpr int S3_NTX = S3_TX / 2;                    // threads along longitude (2 cells each)
constexpr int S3_NT = S3_NTX * S3_TY;                // threads per block
constexpr int S3_SW = S3_TX + 4;                     // staged columns ri in [-1, TX+2]
constexpr int S3_SH = S3_TY + 3;                     // staged rows    rj in [-1, TY+1]
constexpr int S3_SN = S3_SW * S3_SH;
constexpr int S3_PL = (S3_SN + 15) / 16 * 16;        // plane pitch: a multiple of 128 B
constexpr int S3_NP = S3_SN / 2;                     // staged pairs (SW is even)
constexpr int S3_NQ = (S3_NP + S3_NT - 1) / S3_NT;   // staged pairs per thread
// own-column boxes start one column left of the tile: the innermost TMA coordinate must be
// 16-byte aligned (an odd fp64 column raises "illegal instruction"), and I0 - 1 is even
constexpr int S3_OW = S3_TX + 2;
constexpr int S3_OWN = S3_OW * S3_TY;
// TMA ring depth.  3: level k computes, k+1 has landed (its own U, V close the interface
// below level k), k+2 is in flight.  2: level k computes, k+1 is in flight; the own U, V of
// level k+1 are read from global memory (L2 hits: the ring is fetching the same lines), which
// shrinks the block to ~68 KB of shared memory = 3 resident blocks per SM.
#ifndef DC_S3_NBUF
#define DC_S3_NBUF 3
#endif
constexpr int S3_NBUF = DC_S3_NBUF;
constexpr int S3_PF = S3_NBUF - 1;                   // prefetch distance in levels
// 1: the own U, V of level k+1 are read from the ring (level k+1 must have landed when level k
// starts); 0: they are read from global memory (L2) and the ring runs one level further ahead
#ifndef DC_S3_NEED
#define DC_S3_NEED (DC_S3_NBUF >= 3 ? 1 : 0)
#endif
constexpr int S3_NEED = DC_S3_NEED;
// where in level k the ring slot of level k-1 is refilled: 0 = at the top of the level (longest
// lead for the copy), 1 = between the U and the V part (more slack for straggling warps)
#ifndef DC_S3_REFILL_MID
#define DC_S3_REFILL_MID 0
#endif
constexpr int S3_REFILL_MID = DC_S3_REFILL_MID;
// How a ring slot is handed back.  1 (default): one block barrier per level, the refill follows
// it at the top of the next level.  0: no block barrier -- every warp arrives on an `empty`
// mbarrier and the refilling lane waits for it.  Measured on B200 (0.25 deg x 64 levels,
// profiles/r2_stage_variants.md): the barrier-free ring is SLOWER (3.19-3.99 ms/step against
// 3.06): the warps of a block drift apart, and the ring's lead time, not the barrier, is what
// the level loop is sensitive to.
#ifndef DC_S3_SYNC
#define DC_S3_SYNC 1
#endif
constexpr int S3_SYNC = DC_S3_SYNC;
// which warp issues the refill of level k: 1 = warp k mod 4 (the ~130 instructions of the ten
// descriptors rotate over the warps), 0 = always warp 0
#ifndef DC_S3_ROTATE
#define DC_S3_ROTATE 1
#endif
constexpr int S3_ROTATE = DC_S3_ROTATE;

// ---- TMA descriptor -------------------------------------------------------------------
#if defined(__CUDACC__)
typedef CUtensorMap TmaMap;
#else
struct alignas(64) TmaMap {   // host emulation: what cuTensorMapEncodeTiled would encode
    const double *base;
    int dim[3];               // NI, NJ, nk (contiguous)
    int box[3];
    char pad[128 - sizeof(const double *) - 6 * sizeof(int)];
};
#endif

struct alignas(16) D2 {
    double x, y;
};
DC_HD D2 ld2(const double *p) { return *reinterpret_cast<const D2 *>(p); }
DC_HD void st2(double *p, double x, double y) { *reinterpret_cast<D2 *>(p) = D2{x, y}; }
// stencil rows of a pair: columns i_a-1, i_a, i_b, i_b+1 (, i_b+2, i_b+3)
struct R4 {
    double m1, a, b, p1;
};
struct R6 {
    double m1, a, b, p1, p2, p3;
};
DC_HD R4 ld4(const double *p)
{
    const D2 x = ld2(p), y = ld2(p + 2);
    return R4{x.x, x.y, y.x, y.y};
}
DC_HD R6 ld6(const double *p)
{
    const D2 x = ld2(p), y = ld2(p + 2), z = ld2(p + 4);
    return R6{x.x, x.y, y.x, y.y, z.x, z.y};
}

struct alignas(128) Stage3Smem {
    // TMA destinations (128-byte aligned): raw planes of a level, 3-deep ring
    double rU[S3_NBUF][S3_PL], rV[S3_NBUF][S3_PL], rW[S3_NBUF][S3_PL], rPHI[S3_NBUF][S3_PL],
        rT[S3_NBUF][S3_PL], rG[S3_NBUF][S3_PL];
    // own-column boxes: POTTVB[k+1] and the step-start U, V, POTT of level k
    double oTB[S3_NBUF][S3_OWN], oUo[S3_NBUF][S3_OWN], oVo[S3_NBUF][S3_OWN], oTo[S3_NBUF][S3_OWN];
    // flux coefficients of the staged cells (level independent):
    //   CU = (COLP[i-1,j] + COLP[i,j]) / 2, CV = (COLP[i,j-1] + COLP[i,j]) / 2, CP = COLP_NEW*A
    double CU[S3_PL], CV[S3_PL], CP[S3_PL];
    double lev[7][NZMAX + 1];   // per-level tables (see the set-up phase)
    double row[7][S3_TY + 1];   // as StageSmem::row
    double dxr[S3_SH + 1];      // dxjs of the staged rows rj = -1 .. TY+1
    unsigned long long full[S3_NBUF];   // mbarriers: "level has landed" (TMA transaction bytes)
    unsigned long long empty[S3_NBUF];  // mbarriers: "every warp is done with the slot"
};

// ---- SPMD layer for this kernel (1-D block of S3_NT threads) ----------------------------
#if defined(__CUDA_ARCH__)
#define S3_PRIV(type, name) type name
#define S3_PRIVN(type, name, n) type name[n]
#define S3_PRIVNN(type, name, n, m) type name[n][m]
#define S3_P(name) name
#define S3_PHASE {                       \
        const int tid = threadIdx.x;
#define S3_PHASE_END \
    }                \
    __syncthreads();
#define S3_PHASE_END_NOSYNC }
#define S3_PHASE_END_MAYBE_SYNC \
    }                           \
    if (S3_SYNC) __syncthreads();
#define S3_SYNCWARP() __syncwarp()
#else
#define S3_SYNCWARP()
#define S3_PRIV(type, name) type name[S3_NT]
#define S3_PRIVN(type, name, n) type name[S3_NT][n]
#define S3_PRIVNN(type, name, n, m) type name[S3_NT][n][m]
#define S3_P(name) name[tid]
#define S3_PHASE for (int tid = 0; tid < S3_NT; tid++) {
#define S3_PHASE_END }
#define S3_PHASE_END_NOSYNC }
#define S3_PHASE_END_MAYBE_SYNC }
#endif

#if defined(__CUDACC__)
// nvcc: device code issues the PTX; the host pass only needs the symbols
DC_HD unsigned s3_smem_u32(const void *p)
{
#if defined(__CUDA_ARCH__)
    return (unsigned)__cvta_generic_to_shared(p);
#else
    (void)p;
    return 0;
#endif
}
DC_HD void s3_mbar_init(unsigned long long *bar, int count)
{
#if defined(__CUDA_ARCH__)
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(count)
                 : "memory");
#else
    (void)bar; (void)count;
#endif
}
DC_HD void s3_mbar_init_fence()
{
#if defined(__CUDA_ARCH__)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#endif
}
DC_HD void s3_mbar_arrive(unsigned long long *bar)
{
#if defined(__CUDA_ARCH__)
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s3_smem_u32(bar)) : "memory");
#else
    (void)bar;
#endif
}
DC_HD void s3_mbar_expect(unsigned long long *bar, unsigned bytes)
{
#if defined(__CUDA_ARCH__)
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
#else
    (void)bar; (void)bytes;
#endif
}
// Bounded wait: a wrong expect_tx byte count or a failed TMA copy must not hang the GPU.  Each
// try_wait blocks for a hardware-defined time slice; 2^24 unsuccessful slices (seconds) trap, which
// surfaces as a launch failure through dc_last_error.
DC_HD void s3_mbar_wait(unsigned long long *bar, unsigned parity)
{
#if defined(__CUDA_ARCH__)
    const unsigned addr = s3_smem_u32(bar);
#pragma unroll 1
    for (int spin = 0; spin < (1 << 24); spin++) {
        unsigned done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
#else
    (void)bar; (void)parity;
#endif
}
DC_HD void s3_tma_load(void *dst, const TmaMap *map, int x, int y, int z, unsigned long long *bar)
{
#if defined(__CUDA_ARCH__)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s3_smem_u32(dst)),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(s3_smem_u32(bar)), "r"(x), "r"(y),
        "r"(z)
        : "memory");
#else
    (void)dst; (void)map; (void)x; (void)y; (void)z; (void)bar;
#endif
}
#else
inline void s3_mbar_init(unsigned long long *, int) {}
inline void s3_mbar_init_fence() {}
inline void s3_mbar_arrive(unsigned long long *) {}
inline void s3_mbar_expect(unsigned long long *, unsigned) {}
inline void s3_mbar_wait(unsigned long long *, unsigned) {}
inline void s3_tma_load(void *dst, const TmaMap *m, int x, int y, int z, unsigned long long *)
{
    double *d = static_cast<double *>(dst);
    // hardware rule (measured on B200): the innermost box coordinate must be 16-byte aligned
    if ((x & 1) || (reinterpret_cast<size_t>(dst) & 127)) abort();
    for (int b = 0; b < m->box[1]; b++)
        for (int a = 0; a < m->box[0]; a++) {
            const int i = x + a, j = y + b;
            const bool in = i >= 0 && i < m->dim[0] && j >= 0 && j < m->dim[1] && z >= 0 &&
                            z < m->dim[2];
            d[b * m->box[0] + a] =
                in ? m->base[((size_t)z * m->dim[1] + j) * (size_t)m->dim[0] + i] : 0.;
        }
}
#endif

// pressure-gradient term with the per-column part G precomputed (dyn_functions.py:177-207):
// csum = COLP + COLP_dm1, cdif = (COLP - COLP_dm1) * con_cp / 2.
DC_HD double pre_grad_g(double PHI, double PHI_dm1, double csum, double cdif, double G,
                        double G_dm1, double dgrid)
{
    return (-dgrid * ((PHI - PHI_dm1) * csum / 2. + cdif * (+G_dm1 + G)));
}

// momentum advection (dyn_functions.py:541-568).  Production build: the auxiliary fluxes arrive
// pre-multiplied by 1/2 (folded into their 1/12, 1/24 factors), one multiplication less per term
DC_HD double hor_adv_uv(double DWIND, double DWIND_dm1, double DWIND_dp1, double DWIND_pm1,
                        double DWIND_pp1, double DWIND_dm1_pm1, double DWIND_dm1_pp1,
                        double DWIND_dp1_pm1, double DWIND_dp1_pp1, double BRFLX, double BRFLX_dm1,
                        double CQFLX, double CQFLX_pp1, double DSFLX_dm1, double DSFLX_pp1,
                        double ETFLX, double ETFLX_dm1_pp1, double sign_ETFLX_term)
{
    if (DC_FAST)
        return (+BRFLX_dm1 * (DWIND_dm1 + DWIND) - BRFLX * (DWIND + DWIND_dp1)
                + CQFLX * (DWIND_pm1 + DWIND) - CQFLX_pp1 * (DWIND + DWIND_pp1)
                + DSFLX_dm1 * (DWIND_dm1_pm1 + DWIND) - DSFLX_pp1 * (DWIND + DWIND_dp1_pp1)
                + sign_ETFLX_term * (+ETFLX * (DWIND_dp1_pm1 + DWIND) -
                                     ETFLX_dm1_pp1 * (DWIND + DWIND_dm1_pp1)));
    return UVFLX_hor_adv(DWIND, DWIND_dm1, DWIND_dp1, DWIND_pm1, DWIND_pp1, DWIND_dm1_pm1,
                         DWIND_dm1_pp1, DWIND_dp1_pm1, DWIND_dp1_pp1, BRFLX, BRFLX_dm1, CQFLX,
                         CQFLX_pp1, DSFLX_dm1, DSFLX_pp1, ETFLX, ETFLX_dm1_pp1, sign_ETFLX_term);
}

struct Stage3Body {
    Geom g;
    TmaMap mU, mV, mW, mPHI, mT, mG;         // boxes S3_SW x S3_SH x 1
    TmaMap mTB, mUo, mVo, mTo;               // boxes S3_OW x S3_TY x 1
    const double *COLP, *COLP_NEW, *COLP_OLD;
    const double *WWIND, *POTTVB;            // set-up reads of interface 0
    const double *U_in, *V_in;               // the stage's input winds (own columns, S3_NBUF == 2)
    double *UWIND_out, *VWIND_out, *POTT_out;
    // Tiles are cut from the GLOBAL tiling (tile rows start at global rows 1 + m * S3_TY)
    // whatever rows a launch advances: a cell is computed in the same tile footprint, by the same
    // template instantiation and hence with the same FMA contraction, on one GPU and on any
    // number of latitude bands -- the production build stays BITWISE identical across
    // decompositions.  Tile rows [0, nby0) start at global row jt and advance the rows
    // j_lo .. j_hi that fall into them; tile rows from nby0 on start at jt2 and advance
    // j_lo2 .. j_hi2 (the two boundary tile rows of a band in ONE launch).
    int j_lo, j_hi, nby0, j_lo2, j_hi2, jt, jt2;
    int have_old;     // 0: the step-start state is the state the tendencies are evaluated at
    // Small grids (a latitude band at N = 8) do not fill the 2 x 148 block slots for long: the
    // sigma column is then cut into nkc chunks, one block each (blockIdx.z).  A chunk that does
    // not start at the model top first runs the level above it with the stores suppressed --
    // that recreates the vertical momentum flux through its top interface, the only state the
    // march carries -- so the result is bit-identical to the unchunked march.
    int nkc;

    DC_HD int wrap_i(int i) const { return i < 1 ? i + g.nx : (i > g.nx ? i - g.nx : i); }

    DC_HD void run_block(int bx, int by, int bz, Stage3Smem &s) const
    {
        const bool second = by >= nby0;
        const int jl = second ? j_lo2 : j_lo, jh = second ? j_hi2 : j_hi;
        const int jo = second ? jt2 : jt;
        if (second) by -= nby0;
        const int I0 = 1 + bx * S3_TX, J0 = jo + by * S3_TY;
        // the footprint decides, not the rows advanced (see above)
        const bool interior = (I0 >= 3) && (I0 + S3_TX - 1 <= g.nx - 1) && (J0 >= 2) &&
                              (J0 + S3_TY - 1 <= g.ny - 1);
        if (interior)
            run<false>(bx, J0, bz, jl, jh, s);
        else
            run<true>(bx, J0, bz, jl, jh, s);
    }

    // issue the TMA copies of level kp into ring slot (kp - ks) % NBUF (one thread)
    DC_HD void issue(Stage3Smem &s, int kp, int ks, int x0, int y0) const
    {
        const int bp = (kp - ks) % S3_NBUF;
        unsigned long long *bar = &s.full[bp];
        const unsigned bytes =
            6u * S3_SN * 8u + (have_old ? 4u : 1u) * (unsigned)S3_OWN * 8u;
        s3_mbar_expect(bar, bytes);
        s3_tma_load(s.rU[bp], &mU, x0, y0, kp, bar);
        s3_tma_load(s.rV[bp], &mV, x0, y0, kp, bar);
        s3_tma_load(s.rW[bp], &mW, x0, y0, kp + 1, bar);
        s3_tma_load(s.rPHI[bp], &mPHI, x0, y0, kp, bar);
        s3_tma_load(s.rT[bp], &mT, x0, y0, kp, bar);
        s3_tma_load(s.rG[bp], &mG, x0, y0, kp, bar);
        s3_tma_load(s.oTB[bp], &mTB, x0, y0 + 1, kp + 1, bar);
        if (have_old) {
            s3_tma_load(s.oUo[bp], &mUo, x0, y0 + 1, kp, bar);
            s3_tma_load(s.oVo[bp], &mVo, x0, y0 + 1, kp, bar);
            s3_tma_load(s.oTo[bp], &mTo, x0, y0 + 1, kp, bar);
        }
    }

    // Refill of the ring during level k: level k + PF goes into the slot level k-1 used, as
    // soon as every warp has reported that slot free.  One lane of warp k mod (warps per block).
    DC_HD void refill(Stage3Smem &s, int tid, int k, int ks, int ke, int x0, int y0) const
    {
        if (k + S3_PF > ke) return;
        if (tid != (S3_ROTATE ? (k % (S3_NT / 32)) * 32 : 0)) return;
        if (!S3_SYNC && k > ks)
            s3_mbar_wait(&s.empty[(k - 1 - ks) % S3_NBUF], ((k - 1 - ks) / S3_NBUF) & 1);
        issue(s, k + S3_PF, ks, x0, y0);
    }

    template <bool EDGE>
    DC_HD void run(int bx, int J0, int bz, int j_lo, int j_hi, Stage3Smem &s) const
    {
        const int nx = g.nx, ny = g.ny, nz = g.nz;
        // levels of this block: k0 .. k1-1, marched from ks (= k0 - 1 for a warm-up level);
        // planes are needed up to level ke (the own U, V of level k1 close the last interface)
        const int kcl = (nz + nkc - 1) / nkc;
        const int k0 = bz * kcl, k1 = (k0 + kcl < nz) ? k0 + kcl : nz;
        const int ks = k0 > 0 ? k0 - 1 : 0, ke = k1 < nz ? k1 : nz - 1;
        const int I0 = 1 + bx * S3_TX;
        const size_t plane = g.plane;
        const double dyis = g.dyis, dt = g.dt;
        const double scale = cor_scale(g.dlon_rad, g.dlat_rad);
        const int x0 = I0 - 1, y0 = g.row(J0 - 1);   // TMA box origin (columns, device rows)
        // rows this rank holds (global): halo rows included
        const int j_min = (g.j0 - HJ < 0) ? 0 : g.j0 - HJ;
        const int j_max_m = (g.j1 + HJ > ny + 1) ? ny + 1 : g.j1 + HJ;          // mass rows
        const int j_max_y = (g.j1 + HJ + 1 > ny + 2) ? ny + 2 : g.j1 + HJ + 1;  // y-staggered

        // ---- thread-private state --------------------------------------------------------
        // the two own columns a = (ia, j), b = (ia + 1, j)
        S3_PRIV(int, off0);                 // plane offset of cell a
        S3_PRIV(int, flags);                // bit 0/1: a/b is advanced; bit 2/3: a/b has images
        S3_PRIV(double, c_m1);              // COLP at ia-1, ia, ia+1, ia+2 of row j
        S3_PRIV(double, c_a);
        S3_PRIV(double, c_b);
        S3_PRIV(double, c_p1);
        S3_PRIVN(double, c_jm1, 2);
        S3_PRIVN(double, c_jp1, 2);
        S3_PRIVN(double, csx, 2);           // COLP + COLP_im1
        S3_PRIVN(double, cdx, 2);           // (COLP - COLP_im1) * con_cp / 2
        S3_PRIVN(double, csy, 2);
        S3_PRIVN(double, cdy, 2);
        S3_PRIVN(double, cnew, 2);
        S3_PRIVN(double, cold, 2);
        S3_PRIVN(double, colpa_is, 2);
        S3_PRIVN(double, colpa_old_is, 2);
        S3_PRIVN(double, colpa_js, 2);
        S3_PRIVN(double, colpa_old_js, 2);
        S3_PRIVN(double, r_colpa_is, 2);    // reciprocals: DC_FAST_MATH only (dead otherwise)
        S3_PRIVN(double, r_colpa_js, 2);
        S3_PRIVN(double, r_cnew, 2);
        S3_PRIVN(double, wwu_k, 2);         // WWIND_UWIND at interface k (carried)
        S3_PRIVN(double, wwv_k, 2);
        S3_PRIVN(double, w_k, 2);           // WWIND[k]
        S3_PRIVN(double, pottvb_k, 2);
        // production arithmetic (dc_stage3_fast.inc): per-column factors with everything that
        // does not depend on the level folded in (dead in the strict build)
        S3_PRIVN(double, q_cs, 4);          // scale/2 * COLP at ia-1, ia, ia+1, ia+2 of row j
        S3_PRIVN(double, q_csm, 2);         // ... at row j-1, j+1 of the own columns
        S3_PRIVN(double, q_csp, 2);
        S3_PRIVN(double, q_ru, 2);          // Euler step: x_old * q_r* + tendency * q_d*
        S3_PRIVN(double, q_du, 2);
        S3_PRIVN(double, q_rv, 2);
        S3_PRIVN(double, q_dv, 2);
        S3_PRIVN(double, q_rt, 2);
        S3_PRIVN(double, q_dtt, 2);
        S3_PRIVN(double, q_px1, 2);         // pressure gradient along x / y
        S3_PRIVN(double, q_px2, 2);
        S3_PRIVN(double, q_py1, 2);
        S3_PRIVN(double, q_py2, 2);
        S3_PRIVN(double, q_tf, 2);          // WWIND * POTTVB at interface k (carried)

        // ---- set-up -----------------------------------------------------------------------
        S3_PHASE
            if (tid == 0) {
                for (int n = 0; n < S3_NBUF; n++) {
                    s3_mbar_init(&s.full[n], 1);
                    s3_mbar_init(&s.empty[n], S3_NT / 32);
                }
                s3_mbar_init_fence();
            }
        S3_PHASE_END
        S3_PHASE
            if (tid == 0) {   // the first levels fly while the column constants are gathered
                for (int n = 0; n < S3_PF; n++)
                    if (ks + n <= ke) issue(s, ks + n, ks, x0, y0);
            }
            // 1) coefficient planes of the staged region
            for (int n = tid; n < S3_PL; n += S3_NT) {
                const int r = n / S3_SW, cw = n % S3_SW;
                int i = I0 + cw - 1, j = J0 + r - 1;
                if (i > nx + 2) i = nx + 2;   // columns beyond the domain: masked cells only
                if (j < j_min) j = j_min;
                const int jm = j > j_max_m ? j_max_m : j;   // row in a mass / x-staggered field
                const int jy = j > j_max_y ? j_max_y : j;   // row in a y-staggered field
                const int jc = jy > j_max_m ? j_max_m : jy;
                const int jcm = (jy - 1 < j_min) ? j_min : (jy - 1 > j_max_m ? j_max_m : jy - 1);
                // the periodic images are formed from the interior columns [1, nx], as
                // exchange_BC does (misc_boundaries.py:26-32)
                const int iw = wrap_i(i), iwm = wrap_i(i - 1);
                // UFLX = (C[i-1] + C[i])/2 * U * dyis     (dyn_continuity.py:40-41)
                // (production build: dyis / dxjs are folded into the planes, one multiplication
                // less per flux value in the level loop)
                s.CU[n] = (COLP[g.idx2(iwm, jm)] + COLP[g.idx2(iw, jm)]) / 2. *
                          (DC_FAST ? dyis * (1. / 48.) : 1.);
                // VFLX = (C[j-1] + C[j])/2 * V * dxjs     (dyn_continuity.py:43-44)
                s.CV[n] = (COLP[g.idx2(iw, jcm)] + COLP[g.idx2(iw, jc)]) / 2. *
                          (DC_FAST ? g.dxjs[g.row(jy)] * (1. / 48.) : 1.);
                // COLP_NEW * A * WWIND                    (dyn_functions.py:254-260)
                s.CP[n] = COLP_NEW[g.idx2(iw, jm)] * g.A[g.row(jm)] * (DC_FAST ? 0.125 : 1.);
            }
            if (tid <= S3_SH) {
                int j = J0 - 1 + tid;
                if (j < j_min) j = j_min;
                if (j > j_max_y) j = j_max_y;
                s.dxr[tid] = g.dxjs[g.row(j)];
            }
            // 2) constants of the two own columns
            {
                const int tx = tid % S3_NTX, ty = tid / S3_NTX;
                const int ia = I0 + 2 * tx, j = J0 + ty;
                const int vj = (j >= j_lo) && (j <= j_hi);
                const int va = (ia <= nx) && vj, vb = (ia + 1 <= nx) && vj;
                // masked pairs read a safe in-domain neighbourhood
                const int ii = ia > nx ? nx - 1 : ia, jj = j < j_lo ? j_lo : (j <= j_hi ? j : j_hi);
                const int ea = (ia <= 2) || (ia == nx) || (jj == 1) || (jj == ny);
                const int eb = (ia + 1 <= 2) || (ia + 1 == nx) || (jj == 1) || (jj == ny);
                S3_P(flags) = va | (vb << 1) | (ea << 2) | (eb << 3);
                S3_P(off0) = (int)g.idx2(ii, jj);
                const double *C = COLP, *CN = COLP_NEW, *CO = COLP_OLD;
                S3_P(c_m1) = C[g.idx2(ii - 1, jj)];
                S3_P(c_a) = C[g.idx2(ii, jj)];
                S3_P(c_b) = C[g.idx2(ii + 1, jj)];
                S3_P(c_p1) = C[g.idx2(ii + 2, jj)];
                const double A = g.A[g.row(jj)], A_jm1 = g.A[g.row(jj - 1)],
                             A_jp1 = g.A[g.row(jj + 1)];
                for (int e = 0; e < 2; e++) {
                    const int ie = ii + e;
                    const double c = C[g.idx2(ie, jj)], cm = C[g.idx2(ie - 1, jj)];
                    S3_P(c_jm1)[e] = C[g.idx2(ie, jj - 1)];
                    S3_P(c_jp1)[e] = C[g.idx2(ie, jj + 1)];
                    S3_P(csx)[e] = c + cm;
                    S3_P(cdx)[e] = (c - cm) * con_cp / 2.;
                    S3_P(csy)[e] = c + S3_P(c_jm1)[e];
                    S3_P(cdy)[e] = (c - S3_P(c_jm1)[e]) * con_cp / 2.;
                    // the Euler step runs after COLP <- COLP_NEW (dyn_matsuno.py:64-67)
                    S3_P(cnew)[e] = CN[g.idx2(ie, jj)];
                    S3_P(cold)[e] = CO[g.idx2(ie, jj)];
                    S3_P(colpa_is)[e] = interp_COLPA_is(
                        CN[g.idx2(ie, jj)], CN[g.idx2(ie - 1, jj)], CN[g.idx2(ie, jj - 1)],
                        CN[g.idx2(ie, jj + 1)], CN[g.idx2(ie - 1, jj + 1)],
                        CN[g.idx2(ie - 1, jj - 1)], A, A_jm1, A_jp1, jj, ny);
                    S3_P(colpa_old_is)[e] = interp_COLPA_is(
                        CO[g.idx2(ie, jj)], CO[g.idx2(ie - 1, jj)], CO[g.idx2(ie, jj - 1)],
                        CO[g.idx2(ie, jj + 1)], CO[g.idx2(ie - 1, jj + 1)],
                        CO[g.idx2(ie - 1, jj - 1)], A, A_jm1, A_jp1, jj, ny);
                    S3_P(colpa_js)[e] = interp_COLPA_js(
                        CN[g.idx2(ie, jj)], CN[g.idx2(ie, jj - 1)], CN[g.idx2(ie - 1, jj)],
                        CN[g.idx2(ie + 1, jj)], CN[g.idx2(ie + 1, jj - 1)],
                        CN[g.idx2(ie - 1, jj - 1)], A, A_jm1);
                    S3_P(colpa_old_js)[e] = interp_COLPA_js(
                        CO[g.idx2(ie, jj)], CO[g.idx2(ie, jj - 1)], CO[g.idx2(ie - 1, jj)],
                        CO[g.idx2(ie + 1, jj)], CO[g.idx2(ie + 1, jj - 1)],
                        CO[g.idx2(ie - 1, jj - 1)], A, A_jm1);
                    S3_P(r_colpa_is)[e] = DC_FAST ? 1. / S3_P(colpa_is)[e] : 0.;
                    S3_P(r_colpa_js)[e] = DC_FAST ? 1. / S3_P(colpa_js)[e] : 0.;
                    S3_P(r_cnew)[e] = DC_FAST ? 1. / S3_P(cnew)[e] : 0.;
                    S3_P(wwu_k)[e] = 0.;   // WWIND_UWIND[0] = 0 (dyn_functions.py:236-237)
                    S3_P(wwv_k)[e] = 0.;
                    S3_P(w_k)[e] = WWIND[(size_t)ks * plane + S3_P(off0) + e];
                    S3_P(pottvb_k)[e] = POTTVB[(size_t)ks * plane + S3_P(off0) + e];
                }
                if (DC_FAST) {
                    const double sc2 = scale / 2., dxj = g.dxjs[g.row(jj)];
                    S3_P(q_cs)[0] = sc2 * S3_P(c_m1); S3_P(q_cs)[1] = sc2 * S3_P(c_a);
                    S3_P(q_cs)[2] = sc2 * S3_P(c_b);  S3_P(q_cs)[3] = sc2 * S3_P(c_p1);
                    for (int e = 0; e < 2; e++) {
                        S3_P(q_csm)[e] = sc2 * S3_P(c_jm1)[e];
                        S3_P(q_csp)[e] = sc2 * S3_P(c_jp1)[e];
                        S3_P(q_ru)[e] = S3_P(colpa_old_is)[e] / S3_P(colpa_is)[e];
                        S3_P(q_du)[e] = dt / S3_P(colpa_is)[e];
                        S3_P(q_rv)[e] = S3_P(colpa_old_js)[e] / S3_P(colpa_js)[e];
                        S3_P(q_dv)[e] = dt / S3_P(colpa_js)[e];
                        S3_P(q_rt)[e] = S3_P(cold)[e] / S3_P(cnew)[e];
                        S3_P(q_dtt)[e] = dt / S3_P(cnew)[e];
                        S3_P(q_px1)[e] = -dyis * S3_P(csx)[e] / 2.;
                        S3_P(q_px2)[e] = -dyis * S3_P(cdx)[e];
                        S3_P(q_py1)[e] = -dxj * S3_P(csy)[e] / 2.;
                        S3_P(q_py2)[e] = -dxj * S3_P(cdy)[e];
                        // WWIND[0] = 0: the flux through the model top is exactly 0
                        S3_P(q_tf)[e] = ks == 0 ? 0. : S3_P(w_k)[e] * S3_P(pottvb_k)[e];
                    }
                }
            }
            // constant tables
            for (int k = tid; k <= nz; k += S3_NT) {
                s.lev[0][k] = k < nz ? g.dsigma[k] : 0.;
                s.lev[1][k] = k < nz ? g.r_dsigma[k] : 0.;
                s.lev[2][k] = g.sigma_vb[k];
                if (DC_FAST) {
                    // UFLX, VFLX travel scaled by 1/48, COLP by scale/2; interp_ks as two weights
                    s.lev[3][k] = k < nz ? g.UVFLX_dif_coef[k] * 48. : 0.;
                    s.lev[4][k] = k < nz ? g.POTT_dif_coef[k] / (scale / 2.) : 0.;
                    s.lev[5][k] = (k >= 1 && k < nz) ? g.dsigma[k] * g.r_dss[k] : 0.;
                    s.lev[6][k] = (k >= 1 && k < nz) ? g.dsigma[k - 1] * g.r_dss[k] : 0.;
                } else {
                    s.lev[3][k] = k < nz ? g.UVFLX_dif_coef[k] : 0.;
                    s.lev[4][k] = k < nz ? g.POTT_dif_coef[k] : 0.;
                    s.lev[5][k] = k < nz ? g.r_dss[k] : 0.;
                    s.lev[6][k] = 0.;
                }
            }
            if (tid <= S3_TY) {
                int j = J0 - 1 + tid;
                if (j < j_min) j = j_min;       // tile clipped by the band: rows nobody advances
                if (j > j_max_m) j = j_max_m;
                const int r = g.row(j);
                s.row[0][tid] = cor_fcos(g.corf_is[r], g.cos_lat_is[r]);
                s.row[1][tid] = g.sin_lat_is[r] * (DC_FAST ? 0.5 : 1.);
                s.row[2][tid] = cor_fcos(g.corf[r], g.cos_lat[r]);
                s.row[3][tid] = g.sin_lat[r] * (DC_FAST ? 0.5 : 1.);
                s.row[4][tid] = g.dxjs[r];
                s.row[5][tid] = g.A[r];
                s.row[6][tid] = g.r_A[r] * (DC_FAST ? 24. : 1.);
            }
            s3_mbar_wait(&s.full[0], 0);   // the first level has landed
        S3_PHASE_END

        for (int k = ks; k < k1; k++) {
            const int b = (k - ks) % S3_NBUF, b1 = (k + 1 - ks) % S3_NBUF;
            const bool warm = k < k0;   // warm-up level of a chunk: nothing is stored
            const size_t ko = (size_t)k * plane;
            const bool last = (k + 1 == nz);
            // ---- ring: make sure the levels this iteration reads have landed ----------------
            S3_PHASE
                if (EDGE || !S3_REFILL_MID) refill(s, tid, k, ks, ke, x0, y0);
                if (S3_NEED) {
                    if (!last) s3_mbar_wait(&s.full[b1], ((k + 1 - ks) / S3_NBUF) & 1);
                } else {
                    s3_mbar_wait(&s.full[b], ((k - ks) / S3_NBUF) & 1);
                }
            S3_PHASE_END_NOSYNC
            // ---- C: fluxes, tendencies, Euler step, stores ------------------------------
            S3_PHASE
                const int tx = tid % S3_NTX, ty = tid / S3_NTX;
                const int ia = I0 + 2 * tx, j = J0 + ty;
                const int b0 = (ty + 1) * S3_SW + 2 * tx;   // staged word of (ia - 1, j)
                const int o0 = ty * S3_OW + 2 * tx + 1;     // own-box word of cell a
                const double ds = s.lev[0][k];
                const Div ds_d = mkdiv(ds, s.lev[1][k]);
                const double pottvb_kp1[2] = {s.oTB[b][o0], s.oTB[b][o0 + 1]};
                const R4 W_0 = ld4(&s.rW[b][b0]);
                const double w_kp1[2] = {W_0.a, W_0.b};
                const int fl = S3_P(flags);
                if (DC_FAST && (fl & 3)) {
#include "dc_stage3_fast.inc"
                }
                if (!DC_FAST && (fl & 3)) {
                    const bool wall_s = EDGE && (j == 1), wall_n = EDGE && (j == ny);
                    // raw winds around the pair and, from them, UFLX / VFLX
                    // (calc_UFLX, calc_VFLX: dyn_continuity.py:40-47)
                    const R6 U_m = ld6(&s.rU[b][b0 - S3_SW]), U_0 = ld6(&s.rU[b][b0]);
                    const R4 U_p = ld4(&s.rU[b][b0 + S3_SW]);
                    const R4 V_m = ld4(&s.rV[b][b0 - S3_SW]), V_0 = ld4(&s.rV[b][b0]),
                             V_p = ld4(&s.rV[b][b0 + S3_SW]), V_pp = ld4(&s.rV[b][b0 + 2 * S3_SW]);
                    R6 u_m, u_0;
                    R4 u_p, v_m, v_0, v_p, v_pp;
                    {
                        const R6 c_m = ld6(&s.CU[b0 - S3_SW]), c_0 = ld6(&s.CU[b0]);
                        const R4 c_p = ld4(&s.CU[b0 + S3_SW]);
                        if (DC_FAST) {
                            u_m = R6{c_m.m1 * U_m.m1, c_m.a * U_m.a, c_m.b * U_m.b, c_m.p1 * U_m.p1,
                                     c_m.p2 * U_m.p2, 0.};
                            u_0 = R6{c_0.m1 * U_0.m1, c_0.a * U_0.a, c_0.b * U_0.b, c_0.p1 * U_0.p1,
                                     c_0.p2 * U_0.p2, 0.};
                            u_p = R4{c_p.m1 * U_p.m1, c_p.a * U_p.a, c_p.b * U_p.b, c_p.p1 * U_p.p1};
                        } else {
                            u_m = R6{c_m.m1 * U_m.m1 * dyis, c_m.a * U_m.a * dyis,
                                     c_m.b * U_m.b * dyis,   c_m.p1 * U_m.p1 * dyis,
                                     c_m.p2 * U_m.p2 * dyis, 0.};
                            u_0 = R6{c_0.m1 * U_0.m1 * dyis, c_0.a * U_0.a * dyis,
                                     c_0.b * U_0.b * dyis,   c_0.p1 * U_0.p1 * dyis,
                                     c_0.p2 * U_0.p2 * dyis, 0.};
                            u_p = R4{c_p.m1 * U_p.m1 * dyis, c_p.a * U_p.a * dyis,
                                     c_p.b * U_p.b * dyis, c_p.p1 * U_p.p1 * dyis};
                        }
                    }
                    {
                        const R4 c_m = ld4(&s.CV[b0 - S3_SW]), c_0 = ld4(&s.CV[b0]),
                                 c_p = ld4(&s.CV[b0 + S3_SW]), c_pp = ld4(&s.CV[b0 + 2 * S3_SW]);
                        if (DC_FAST) {
                            v_m = R4{c_m.m1 * V_m.m1, c_m.a * V_m.a, c_m.b * V_m.b, c_m.p1 * V_m.p1};
                            v_0 = R4{c_0.m1 * V_0.m1, c_0.a * V_0.a, c_0.b * V_0.b, c_0.p1 * V_0.p1};
                            v_p = R4{c_p.m1 * V_p.m1, c_p.a * V_p.a, c_p.b * V_p.b, c_p.p1 * V_p.p1};
                            v_pp = R4{c_pp.m1 * V_pp.m1, c_pp.a * V_pp.a, c_pp.b * V_pp.b, 0.};
                        } else {
                            const double dx_m = s.dxr[ty], dx_0 = s.dxr[ty + 1],
                                         dx_p = s.dxr[ty + 2], dx_pp = s.dxr[ty + 3];
                            v_m = R4{c_m.m1 * V_m.m1 * dx_m, c_m.a * V_m.a * dx_m,
                                     c_m.b * V_m.b * dx_m, c_m.p1 * V_m.p1 * dx_m};
                            v_0 = R4{c_0.m1 * V_0.m1 * dx_0, c_0.a * V_0.a * dx_0,
                                     c_0.b * V_0.b * dx_0, c_0.p1 * V_0.p1 * dx_0};
                            v_p = R4{c_p.m1 * V_p.m1 * dx_p, c_p.a * V_p.a * dx_p,
                                     c_p.b * V_p.b * dx_p, c_p.p1 * V_p.p1 * dx_p};
                            v_pp = R4{c_pp.m1 * V_pp.m1 * dx_pp, c_pp.a * V_pp.a * dx_pp,
                                      c_pp.b * V_pp.b * dx_pp, 0.};
                        }
                    }
                    const double u[2] = {U_0.a, U_0.b}, v[2] = {V_0.a, V_0.b};
                    // vertical momentum fluxes through interface k+1
                    // (dyn_functions.py:211-270; 0 at the model bottom)
                    double wwu_kp1[2] = {0., 0.}, wwv_kp1[2] = {0., 0.};
                    if (!last) {
                        // COLP_NEW * A * WWIND(k+1) around the pair
                        R4 P_m, P_0, P_p;
                        {
                            const R4 W_m = ld4(&s.rW[b][b0 - S3_SW]), W_p = ld4(&s.rW[b][b0 + S3_SW]);
                            const R4 c_m = ld4(&s.CP[b0 - S3_SW]), c_0 = ld4(&s.CP[b0]),
                                     c_p = ld4(&s.CP[b0 + S3_SW]);
                            P_m = R4{c_m.m1 * W_m.m1, c_m.a * W_m.a, c_m.b * W_m.b, c_m.p1 * W_m.p1};
                            P_0 = R4{c_0.m1 * W_0.m1, c_0.a * W_0.a, c_0.b * W_0.b, c_0.p1 * W_0.p1};
                            P_p = R4{c_p.m1 * W_p.m1, c_p.a * W_p.a, c_p.b * W_p.b, 0.};
                        }
                        const double ds_kp1 = s.lev[0][k + 1];
                        const Div dss_d = mkdiv(ds_kp1 + ds, s.lev[5][k + 1]);
                        const int wall = wall_s ? -1 : (wall_n ? 1 : 0);
                        double u1[2], v1[2];
                        if (S3_NEED) {
                            u1[0] = s.rU[b1][b0 + 1]; u1[1] = s.rU[b1][b0 + 2];
                            v1[0] = s.rV[b1][b0 + 1]; v1[1] = s.rV[b1][b0 + 2];
                        } else {
                            const size_t o1 = ko + plane + S3_P(off0);
                            u1[0] = U_in[o1]; u1[1] = U_in[o1 + 1];
                            v1[0] = V_in[o1]; v1[1] = V_in[o1 + 1];
                        }
                        wwu_kp1[0] = colpa_wwind(P_0.a, P_0.m1, P_m.a, P_p.a, P_m.m1, P_p.m1, wall) *
                                     interp_ks(u1[0], u[0], ds_kp1, ds, dss_d);
                        wwu_kp1[1] = colpa_wwind(P_0.b, P_0.a, P_m.b, P_p.b, P_m.a, P_p.a, wall) *
                                     interp_ks(u1[1], u[1], ds_kp1, ds, dss_d);
                        wwv_kp1[0] = colpa_wwind(P_0.a, P_m.a, P_0.m1, P_0.b, P_m.m1, P_m.b, 0) *
                                     interp_ks(v1[0], v[0], ds_kp1, ds, dss_d);
                        wwv_kp1[1] = colpa_wwind(P_0.b, P_m.b, P_0.a, P_0.p1, P_m.a, P_m.p1, 0) *
                                     interp_ks(v1[1], v[1], ds_kp1, ds, dss_d);
                    }
                    const R4 PHI_0 = ld4(&s.rPHI[b][b0]), G_0 = ld4(&s.rG[b][b0]);
                    const double coef_uv = s.lev[3][k];
                    // step-start values of the own cells
                    double uo[2] = {u[0], u[1]}, vo[2] = {v[0], v[1]};
                    if (have_old) {
                        uo[0] = s.oUo[b][o0]; uo[1] = s.oUo[b][o0 + 1];
                        vo[0] = s.oVo[b][o0]; vo[1] = s.oVo[b][o0 + 1];
                    }
                    // ---------------- dUFLXdt (dyn_UFLX.py:69-199) ----------------
                    {
                        // auxiliary fluxes (dyn_functions.py:429-536) around the pair
                        double B_m1, B_a, B_b, C_a, C_b, C_a_jp1, C_b_jp1, D_m1, D_a, D_a_jp1,
                            D_b_jp1, E_a, E_b, E_m1_jp1, E_a_jp1;
                        if (DC_FAST) {
                            // shared partial sums; every flux carries the 1/2 of its advection term
                            const double h12 = 1. / 24., h24 = 1. / 48.;
                            const double uhm_m1 = u_m.m1 + u_m.a, uhm_a = u_m.a + u_m.b,
                                         uhm_b = u_m.b + u_m.p1;
                            const double uh0_m1 = u_0.m1 + u_0.a, uh0_a = u_0.a + u_0.b,
                                         uh0_b = u_0.b + u_0.p1;
                            const double uhp_m1 = u_p.m1 + u_p.a, uhp_a = u_p.a + u_p.b,
                                         uhp_b = u_p.b + u_p.p1;
                            B_m1 = h12 * (uhm_m1 + 2. * uh0_m1 + uhp_m1);
                            B_a = h12 * (uhm_a + 2. * uh0_a + uhp_a);
                            B_b = h12 * (uhm_b + 2. * uh0_b + uhp_b);
                            const double vs0_m1 = v_m.m1 + 2. * v_0.m1 + v_p.m1,
                                         vs0_a = v_m.a + 2. * v_0.a + v_p.a,
                                         vs0_b = v_m.b + 2. * v_0.b + v_p.b;
                            const double vs1_m1 = v_0.m1 + 2. * v_p.m1 + v_pp.m1,
                                         vs1_a = v_0.a + 2. * v_p.a + v_pp.a,
                                         vs1_b = v_0.b + 2. * v_p.b + v_pp.b;
                            C_a = h12 * (vs0_m1 + vs0_a);
                            C_b = h12 * (vs0_a + vs0_b);
                            C_a_jp1 = h12 * (vs1_m1 + vs1_a);
                            C_b_jp1 = h12 * (vs1_a + vs1_b);
                            const double us0_m1 = uhm_m1 + uh0_m1, us0_a = uhm_a + uh0_a,
                                         us0_b = uhm_b + uh0_b;
                            const double us1_m1 = uh0_m1 + uhp_m1, us1_a = uh0_a + uhp_a,
                                         us1_b = uh0_b + uhp_b;
                            D_m1 = h24 * (vs0_m1 + us0_m1);
                            D_a = h24 * (vs0_a + us0_a);
                            D_a_jp1 = h24 * (vs1_a + us1_a);
                            D_b_jp1 = h24 * (vs1_b + us1_b);
                            E_a = h24 * (vs0_a - us0_a);
                            E_b = h24 * (vs0_b - us0_b);
                            E_m1_jp1 = h24 * (vs1_m1 - us1_m1);
                            E_a_jp1 = h24 * (vs1_a - us1_a);
                        } else {
                            B_m1 = calc_BFLX(u_m.m1, u_m.a, u_0.m1, u_0.a, u_p.m1, u_p.a);
                            B_a = calc_BFLX(u_m.a, u_m.b, u_0.a, u_0.b, u_p.a, u_p.b);
                            B_b = calc_BFLX(u_m.b, u_m.p1, u_0.b, u_0.p1, u_p.b, u_p.p1);
                            C_a = calc_CFLX(v_m.m1, v_m.a, v_0.m1, v_0.a, v_p.m1, v_p.a);
                            C_b = calc_CFLX(v_m.a, v_m.b, v_0.a, v_0.b, v_p.a, v_p.b);
                            C_a_jp1 = calc_CFLX(v_0.m1, v_0.a, v_p.m1, v_p.a, v_pp.m1, v_pp.a);
                            C_b_jp1 = calc_CFLX(v_0.a, v_0.b, v_p.a, v_p.b, v_pp.a, v_pp.b);
                            D_m1 = calc_DFLX(v_m.m1, v_0.m1, v_p.m1, u_m.m1, u_0.m1, u_m.a, u_0.a);
                            D_a = calc_DFLX(v_m.a, v_0.a, v_p.a, u_m.a, u_0.a, u_m.b, u_0.b);
                            D_a_jp1 = calc_DFLX(v_0.a, v_p.a, v_pp.a, u_0.a, u_p.a, u_0.b, u_p.b);
                            D_b_jp1 = calc_DFLX(v_0.b, v_p.b, v_pp.b, u_0.b, u_p.b, u_0.p1, u_p.p1);
                            E_a = calc_EFLX(v_m.a, v_0.a, v_p.a, u_m.a, u_0.a, u_m.b, u_0.b);
                            E_b = calc_EFLX(v_m.b, v_0.b, v_p.b, u_m.b, u_0.b, u_m.p1, u_0.p1);
                            E_m1_jp1 =
                                calc_EFLX(v_0.m1, v_p.m1, v_pp.m1, u_0.m1, u_p.m1, u_0.a, u_p.a);
                            E_a_jp1 = calc_EFLX(v_0.a, v_p.a, v_pp.a, u_0.a, u_p.a, u_0.b, u_p.b);
                        }
                        if (wall_s) {   // BCy, dyn_UFLX.py:377-384
                            D_m1 = 0.; D_a = 0.; C_a = 0.; C_b = 0.; E_a = 0.; E_b = 0.;
                        }
                        if (wall_n) {
                            D_a_jp1 = 0.; D_b_jp1 = 0.; C_a_jp1 = 0.; C_b_jp1 = 0.;
                            E_m1_jp1 = 0.; E_a_jp1 = 0.;
                        }
                        double d[2] = {0., 0.};
                        d[0] = d[0] + hor_adv_uv(U_0.a, U_0.m1, U_0.b, U_m.a, U_p.a, U_m.m1, U_p.m1,
                                                    U_m.b, U_p.b, B_a, B_m1, C_a, C_a_jp1, D_m1,
                                                    D_a_jp1, E_a, E_m1_jp1, 1.);
                        d[1] = d[1] + hor_adv_uv(U_0.b, U_0.a, U_0.p1, U_m.b, U_p.b, U_m.a, U_p.a,
                                                    U_m.p1, U_p.p1, B_b, B_a, C_b, C_b_jp1, D_a,
                                                    D_b_jp1, E_b, E_a_jp1, 1.);
                        d[0] = d[0] + ((S3_P(wwu_k)[0] - wwu_kp1[0]) / ds_d);
                        d[1] = d[1] + ((S3_P(wwu_k)[1] - wwu_kp1[1]) / ds_d);
                        const double fcos_is = s.row[0][ty + 1], sin_is = s.row[1][ty + 1];
                        d[0] = d[0] + coriolis_UWIND(S3_P(c_a), S3_P(c_m1), V_0.a, V_0.m1, V_p.a,
                                                     V_p.m1, U_0.a, U_0.m1, U_0.b, fcos_is, sin_is,
                                                     scale);
                        d[1] = d[1] + coriolis_UWIND(S3_P(c_b), S3_P(c_a), V_0.b, V_0.a, V_p.b, V_p.a,
                                                     U_0.b, U_0.a, U_0.p1, fcos_is, sin_is, scale);
                        d[0] = d[0] + pre_grad_g(PHI_0.a, PHI_0.m1, S3_P(csx)[0], S3_P(cdx)[0], G_0.a,
                                                 G_0.m1, dyis);
                        d[1] = d[1] + pre_grad_g(PHI_0.b, PHI_0.a, S3_P(csx)[1], S3_P(cdx)[1], G_0.b,
                                                 G_0.a, dyis);
                        if (coef_uv > 0.) {
                            d[0] = d[0] + num_dif(u_0.a, u_0.m1, u_0.b, u_m.a, u_p.a, coef_uv);
                            d[1] = d[1] + num_dif(u_0.b, u_0.a, u_0.p1, u_m.b, u_p.b, coef_uv);
                        }
                        for (int e = 0; e < 2; e++) {
                            const double un = euler_forward_pw(
                                uo[e], d[e], mkdiv(S3_P(colpa_is)[e], S3_P(r_colpa_is)[e]),
                                S3_P(colpa_old_is)[e], dt);
                            if (warm) continue;
                            if (EDGE) {
                                if (fl & (1 << e)) {
                                    if (fl & (4 << e))
                                        put_xstag(g, UWIND_out, ia + e, j, k, un);
                                    else
                                        UWIND_out[ko + S3_P(off0) + e] = un;
                                }
                            } else if (fl & (1 << e)) {   // rows outside the launch's range are masked
                                UWIND_out[ko + S3_P(off0) + e] = un;
                            }
                        }
                    }
                    if (!EDGE && S3_REFILL_MID) refill(s, tid, k, ks, ke, x0, y0);
                    // ---------------- dVFLXdt (dyn_VFLX.py:67-198) ----------------
                    if (!EDGE || j >= 2) {
                        double R_a, R_b, R_a_jm1, R_b_jm1, Q_a, Q_b, Q_p1, S_a_jm1, S_b_jm1, S_b,
                            S_p1, T_a, T_b, T_b_jm1, T_p1_jm1;
                        if (DC_FAST) {
                            const double h12 = 1. / 24., h24 = 1. / 48.;
                            const double vvm_m1 = v_m.m1 + v_0.m1, vvm_a = v_m.a + v_0.a,
                                         vvm_b = v_m.b + v_0.b, vvm_p1 = v_m.p1 + v_0.p1;
                            const double vv0_m1 = v_0.m1 + v_p.m1, vv0_a = v_0.a + v_p.a,
                                         vv0_b = v_0.b + v_p.b, vv0_p1 = v_0.p1 + v_p.p1;
                            R_a = h12 * (vv0_m1 + 2. * vv0_a + vv0_b);
                            R_b = h12 * (vv0_a + 2. * vv0_b + vv0_p1);
                            R_a_jm1 = h12 * (vvm_m1 + 2. * vvm_a + vvm_b);
                            R_b_jm1 = h12 * (vvm_a + 2. * vvm_b + vvm_p1);
                            const double uv_m1 = u_m.m1 + u_0.m1, uv_a = u_m.a + u_0.a,
                                         uv_b = u_m.b + u_0.b, uv_p1 = u_m.p1 + u_0.p1,
                                         uv_p2 = u_m.p2 + u_0.p2;
                            Q_a = h12 * (uv_m1 + 2. * uv_a + uv_b);
                            Q_b = h12 * (uv_a + 2. * uv_b + uv_p1);
                            Q_p1 = h12 * (uv_b + 2. * uv_p1 + uv_p2);
                            const double uu0_a = u_0.m1 + 2. * u_0.a + u_0.b,
                                         uu0_b = u_0.a + 2. * u_0.b + u_0.p1,
                                         uu0_p1 = u_0.b + 2. * u_0.p1 + u_0.p2;
                            const double uum_a = u_m.m1 + 2. * u_m.a + u_m.b,
                                         uum_b = u_m.a + 2. * u_m.b + u_m.p1,
                                         uum_p1 = u_m.b + 2. * u_m.p1 + u_m.p2;
                            const double wm_a = vvm_m1 + vvm_a, wm_b = vvm_a + vvm_b,
                                         wm_p1 = vvm_b + vvm_p1;
                            const double w0_a = vv0_m1 + vv0_a, w0_b = vv0_a + vv0_b,
                                         w0_p1 = vv0_b + vv0_p1;
                            S_a_jm1 = h24 * (wm_a + uum_a);
                            S_b_jm1 = h24 * (wm_b + uum_b);
                            S_b = h24 * (w0_b + uu0_b);
                            S_p1 = h24 * (w0_p1 + uu0_p1);
                            T_a = h24 * (w0_a - uu0_a);
                            T_b = h24 * (w0_b - uu0_b);
                            T_b_jm1 = h24 * (wm_b - uum_b);
                            T_p1_jm1 = h24 * (wm_p1 - uum_p1);
                        } else {
                            R_a = calc_RFLX(v_0.m1, v_p.m1, v_0.a, v_p.a, v_0.b, v_p.b);
                            R_b = calc_RFLX(v_0.a, v_p.a, v_0.b, v_p.b, v_0.p1, v_p.p1);
                            R_a_jm1 = calc_RFLX(v_m.m1, v_0.m1, v_m.a, v_0.a, v_m.b, v_0.b);
                            R_b_jm1 = calc_RFLX(v_m.a, v_0.a, v_m.b, v_0.b, v_m.p1, v_0.p1);
                            Q_a = calc_QFLX(u_m.m1, u_0.m1, u_m.a, u_0.a, u_m.b, u_0.b);
                            Q_b = calc_QFLX(u_m.a, u_0.a, u_m.b, u_0.b, u_m.p1, u_0.p1);
                            Q_p1 = calc_QFLX(u_m.b, u_0.b, u_m.p1, u_0.p1, u_m.p2, u_0.p2);
                            S_a_jm1 = calc_SFLX(v_m.m1, v_0.m1, v_m.a, v_0.a, u_m.m1, u_m.a, u_m.b);
                            S_b_jm1 = calc_SFLX(v_m.a, v_0.a, v_m.b, v_0.b, u_m.a, u_m.b, u_m.p1);
                            S_b = calc_SFLX(v_0.a, v_p.a, v_0.b, v_p.b, u_0.a, u_0.b, u_0.p1);
                            S_p1 = calc_SFLX(v_0.b, v_p.b, v_0.p1, v_p.p1, u_0.b, u_0.p1, u_0.p2);
                            T_a = calc_TFLX(v_0.m1, v_p.m1, v_0.a, v_p.a, u_0.m1, u_0.a, u_0.b);
                            T_b = calc_TFLX(v_0.a, v_p.a, v_0.b, v_p.b, u_0.a, u_0.b, u_0.p1);
                            T_b_jm1 = calc_TFLX(v_m.a, v_0.a, v_m.b, v_0.b, u_m.a, u_m.b, u_m.p1);
                            T_p1_jm1 =
                                calc_TFLX(v_m.b, v_0.b, v_m.p1, v_0.p1, u_m.b, u_m.p1, u_m.p2);
                        }
                        double d[2] = {0., 0.};
                        d[0] = d[0] + hor_adv_uv(V_0.a, V_m.a, V_p.a, V_0.m1, V_0.b, V_m.m1, V_m.b,
                                                    V_p.m1, V_p.b, R_a, R_a_jm1, Q_a, Q_b, S_a_jm1,
                                                    S_b, T_a, T_b_jm1, -1.);
                        d[1] = d[1] + hor_adv_uv(V_0.b, V_m.b, V_p.b, V_0.a, V_0.p1, V_m.a, V_m.p1,
                                                    V_p.a, V_p.p1, R_b, R_b_jm1, Q_b, Q_p1, S_b_jm1,
                                                    S_p1, T_b, T_p1_jm1, -1.);
                        d[0] = d[0] + ((S3_P(wwv_k)[0] - wwv_kp1[0]) / ds_d);
                        d[1] = d[1] + ((S3_P(wwv_k)[1] - wwv_kp1[1]) / ds_d);
                        const double fcos = s.row[2][ty + 1], sinl = s.row[3][ty + 1],
                                     fcos_jm1 = s.row[2][ty], sinl_jm1 = s.row[3][ty];
                        d[0] = d[0] + coriolis_VWIND(S3_P(c_a), S3_P(c_jm1)[0], U_0.a, U_m.a, U_0.b,
                                                     U_m.b, fcos, sinl, fcos_jm1, sinl_jm1, scale);
                        d[1] = d[1] + coriolis_VWIND(S3_P(c_b), S3_P(c_jm1)[1], U_0.b, U_m.b, U_0.p1,
                                                     U_m.p1, fcos, sinl, fcos_jm1, sinl_jm1, scale);
                        const double dxjs = s.row[4][ty + 1];
                        d[0] = d[0] + pre_grad_g(PHI_0.a, s.rPHI[b][b0 - S3_SW + 1], S3_P(csy)[0],
                                                 S3_P(cdy)[0], G_0.a, s.rG[b][b0 - S3_SW + 1], dxjs);
                        d[1] = d[1] + pre_grad_g(PHI_0.b, s.rPHI[b][b0 - S3_SW + 2], S3_P(csy)[1],
                                                 S3_P(cdy)[1], G_0.b, s.rG[b][b0 - S3_SW + 2], dxjs);
                        if (coef_uv > 0.) {
                            d[0] = d[0] + num_dif(v_0.a, v_0.m1, v_0.b, v_m.a, v_p.a, coef_uv);
                            d[1] = d[1] + num_dif(v_0.b, v_0.a, v_0.p1, v_m.b, v_p.b, coef_uv);
                        }
                        for (int e = 0; e < 2; e++) {
                            const double vn = euler_forward_pw(
                                vo[e], d[e], mkdiv(S3_P(colpa_js)[e], S3_P(r_colpa_js)[e]),
                                S3_P(colpa_old_js)[e], dt);
                            if (warm) continue;
                            if (EDGE) {
                                if (fl & (1 << e)) {
                                    if (fl & (4 << e))
                                        put_ystag(g, VWIND_out, ia + e, j, k, vn);
                                    else
                                        VWIND_out[ko + S3_P(off0) + e] = vn;
                                }
                            } else if (fl & (1 << e)) {   // rows outside the launch's range are masked
                                VWIND_out[ko + S3_P(off0) + e] = vn;
                            }
                        }
                    } else {
                        for (int e = 0; e < 2; e++)
                            if (!warm && (fl & (1 << e))) put_ystag(g, VWIND_out, ia + e, 1, k, 0.);
                    }
                    if (wall_n)
                        for (int e = 0; e < 2; e++)
                            if (!warm && (fl & (1 << e)))
                                put_ystag(g, VWIND_out, ia + e, ny + 1, k, 0.);
                    // ---------------- dPOTTdt (dyn_POTT.py:55-110) ----------------
                    {
                        const R4 T_0 = ld4(&s.rT[b][b0]);
                        const double p_jm1[2] = {s.rT[b][b0 - S3_SW + 1], s.rT[b][b0 - S3_SW + 2]};
                        const double p_jp1[2] = {s.rT[b][b0 + S3_SW + 1], s.rT[b][b0 + S3_SW + 2]};
                        const Div A_d = mkdiv(s.row[5][ty + 1], s.row[6][ty + 1]);
                        const double coef = s.lev[4][k];
                        double to[2] = {T_0.a, T_0.b};
                        if (have_old) {
                            to[0] = s.oTo[b][o0]; to[1] = s.oTo[b][o0 + 1];
                        }
                        double d[2] = {0., 0.};
                        d[0] = d[0] + hor_adv(T_0.a, T_0.m1, T_0.b, p_jm1[0], p_jp1[0], u_0.a, u_0.b,
                                              v_0.a, v_p.a, A_d);
                        d[1] = d[1] + hor_adv(T_0.b, T_0.a, T_0.p1, p_jm1[1], p_jp1[1], u_0.b, u_0.p1,
                                              v_0.b, v_p.b, A_d);
                        for (int e = 0; e < 2; e++)
                            d[e] = d[e] + vert_adv(S3_P(pottvb_k)[e], pottvb_kp1[e], S3_P(w_k)[e],
                                                   w_kp1[e], S3_P(cnew)[e], ds_d, k);
                        if (coef > 0.) {
                            d[0] = d[0] + num_dif_pw(T_0.a, T_0.m1, T_0.b, p_jm1[0], p_jp1[0],
                                                     S3_P(c_a), S3_P(c_m1), S3_P(c_b),
                                                     S3_P(c_jm1)[0], S3_P(c_jp1)[0], coef);
                            d[1] = d[1] + num_dif_pw(T_0.b, T_0.a, T_0.p1, p_jm1[1], p_jp1[1],
                                                     S3_P(c_b), S3_P(c_a), S3_P(c_p1),
                                                     S3_P(c_jm1)[1], S3_P(c_jp1)[1], coef);
                        }
                        for (int e = 0; e < 2; e++) {
                            const double tn = euler_forward_pw(
                                to[e], d[e], mkdiv(S3_P(cnew)[e], S3_P(r_cnew)[e]), S3_P(cold)[e],
                                dt);
                            if (warm) continue;
                            if (EDGE) {
                                if (fl & (1 << e)) {
                                    if (fl & (4 << e))
                                        put_mass(g, POTT_out, ia + e, j, k, tn);
                                    else
                                        POTT_out[ko + S3_P(off0) + e] = tn;
                                }
                            } else if (fl & (1 << e)) {   // rows outside the launch's range are masked
                                POTT_out[ko + S3_P(off0) + e] = tn;
                            }
                        }
                    }
                    for (int e = 0; e < 2; e++) {
                        S3_P(wwu_k)[e] = wwu_kp1[e];
                        S3_P(wwv_k)[e] = wwv_kp1[e];
                    }
                }
                for (int e = 0; e < 2; e++) {
                    S3_P(w_k)[e] = w_kp1[e];
                    S3_P(pottvb_k)[e] = pottvb_kp1[e];
                }
                // this warp is done with the slot of level k (its reads are complete: their
                // values were consumed above); the refill of level k+1 waits for all warps
                if (!S3_SYNC && k + 1 + S3_PF <= ke) {
                    S3_SYNCWARP();
                    if ((tid & 31) == 0) s3_mbar_arrive(&s.empty[b]);
                }
            S3_PHASE_END_MAYBE_SYNC
        }
    }
};

// the x-staggered halo column nx+2 of an imported UWIND is never initialised by the
// reference's set-up boundary condition (main_grid.py:340-343); the TMA-staged kernel reads
// the periodic image there.  threads: i = level, j = held rows
struct XHaloFixBody {
    Geom g;
    double *U;
    DC_HD void operator()(int k, int j) const { U[g.idx(g.nx + 2, j, k)] = U[g.idx(2, j, k)]; }
};

}  // namespace dc
