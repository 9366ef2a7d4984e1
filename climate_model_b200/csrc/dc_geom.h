// dc_geom.h -- geometry / field-table structs shared by the CUDA kernels (dyncore.cu)
// and the host emulation harness used by the CPU tests (tests/emu/).
//
// Device data layout (chosen for B200, NOT the reference's):
//   every field is stored level-major with LONGITUDE FASTEST:  F[k][jd][i]
//   - i  = reference lon index incl. its one halo cell (0 .. nx+1, x-staggered .. nx+2),
//          row pitch NI (multiple of 16 doubles = 128 B) shared by all fields
//   - jd = device row = global reference row j + jshift; every plane has NJ rows.  A rank
//          owns the global mass rows j0..j1 and keeps HJ = 2 halo rows on each side, so
//          jshift = HJ - j0 (single GPU: j0 = 1, jshift = 1, global halo row 0 -> jd 1)
//   - k  = level (nz) or interface (nz+1); plane stride NI*NJ
//   2-D fields are one plane; 1-D per-row geometry is indexed by jd, per-level by k.
// The reference layout is (i, j, k) with k fastest (main_fields.py:477-485); the Python
// field API converts (climate_model_b200/main_fields.py).
#pragma once
#include <stddef.h>

#if defined(__CUDACC__)
#define DC_HD __host__ __device__ __forceinline__
#else
#define DC_HD inline
#endif

namespace dc {

// io_constants.py:16-21
constexpr double con_g = 9.81;
constexpr double con_rE = 6371000.;
constexpr double con_Rd = 287.058;
constexpr double con_cp = 1005.;
constexpr double con_kappa = con_Rd / con_cp;

constexpr int HJ = 2;  // halo rows per side kept by a latitude band

struct Geom {
    int nx, ny, nz;    // GLOBAL interior sizes (reference nx, ny, nz); nb == 1
    int NI, NJ;        // row pitch (doubles) and rows per plane of the local band
    int jshift;        // device row = global j + jshift
    int j0, j1;        // global mass rows owned by this rank (1 <= j0 <= j1 <= ny)
    int i_moist;       // namelist.i_moist_main_switch
    int i_coupling;    // physics coupling terms (KMOM, KHEAT, surface fluxes) are evaluated
    size_t plane;      // NI * NJ
    double dt;         // GR.dt
    double pair_top;   // namelist.pair_top
    double dyis;       // GRF['dyis']  (constant field, main_grid.py:214)
    double dlon_rad;   // GRF['dlon_rad'] (constant field)
    double dlat_rad;   // GRF['dlat_rad'] (constant field)
    // per-row geometry, length NJ, indexed by device row (values depend on latitude only)
    const double *A;           // GRF['A'][.,j]
    const double *dxjs;        // GRF['dxjs'][.,j]   (y-staggered rows; 0 on the walls)
    const double *corf;        // GRF['corf'][.,j]
    const double *corf_is;     // GRF['corf_is'][.,j]
    const double *cos_lat;     // cos(GRF['lat_rad'][.,j])     host libm
    const double *sin_lat;     // sin(GRF['lat_rad'][.,j])
    const double *cos_lat_is;  // cos(GRF['lat_is_rad'][.,j])
    const double *sin_lat_is;  // sin(GRF['lat_is_rad'][.,j])
    // per-level, length nz (+1 for sigma_vb)
    const double *sigma_vb, *dsigma, *UVFLX_dif_coef, *POTT_dif_coef, *moist_dif_coef;
    // reciprocals used by the DC_FAST_MATH build only (dc_point.h: struct Div)
    const double *r_A;        // 1 / A[row]
    const double *r_dsigma;   // 1 / dsigma[k]
    const double *r_dss;      // 1 / (dsigma[k] + dsigma[k-1]),  k >= 1
    const double *powtab;     // table of pow_kappa_tab (dc_point.h), DC_FAST_MATH only

    DC_HD size_t idx(int i, int j, int k) const
    {
        return ((size_t)k * (size_t)NJ + (size_t)(j + jshift)) * (size_t)NI + (size_t)i;
    }
    DC_HD size_t idx2(int i, int j) const { return (size_t)(j + jshift) * (size_t)NI + (size_t)i; }
    DC_HD int row(int j) const { return j + jshift; }
};

// Field table: device pointers bound by the caller (torch owns the memory).
// Order = enum FieldId in include/dyncore.h (DC_FIELD_LIST).
#define DC_FIELD_LIST(X)                                                                     \
    X(COLP) X(COLP_OLD) X(COLP_NEW) X(dCOLPdt) X(HSURF)                                      \
    X(UWIND) X(UWIND_OLD) X(VWIND) X(VWIND_OLD) X(WWIND)                                     \
    X(POTT) X(POTT_OLD) X(QV) X(QV_OLD) X(QC) X(QC_OLD)                                      \
    X(UFLX) X(VFLX) X(FLXDIV)                                                                \
    X(BFLX) X(CFLX) X(DFLX) X(EFLX) X(RFLX) X(QFLX) X(SFLX) X(TFLX)                          \
    X(WWIND_UWIND) X(WWIND_VWIND)                                                            \
    X(dUFLXdt) X(dVFLXdt) X(dPOTTdt) X(dQVdt) X(dQCdt)                                       \
    X(PHI) X(PHIVB) X(PVTF) X(PVTFVB) X(POTTVB)                                              \
    X(TAIR) X(TAIRVB) X(PAIR) X(PAIRVB) X(RHO) X(RHOVB) X(WINDX) X(WINDY) X(WIND)         \
    X(PGCOL)                                                                                 \
    X(KMOM) X(KHEAT) X(SMOMXFLX) X(SMOMYFLX) X(SSHFLX) X(SLHFLX)                             \
    X(KMOM_dUWINDdz) X(KMOM_dVWINDdz)                                                        \
    X(dUFLXdt_TURB) X(dVFLXdt_TURB) X(dPOTTdt_TURB) X(dQVdt_TURB) X(dPOTTdt_RAD)

struct Fields {
#define X(n) double *n;
    DC_FIELD_LIST(X)
#undef X
};

enum FieldId {
#define X(n) F_##n,
    DC_FIELD_LIST(X)
#undef X
        F_COUNT
};

}  // namespace dc
