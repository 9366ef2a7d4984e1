/*
 * dyncore_oracle.h -- CPU restatement (plain C, fp64) of the reference's numba-CPU
 * dynamical-core path (Potopoles/Climate_Model: dyn_matsuno.py and the kernels it drives).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (climate_model_b200/,
 * include/, csrc/) may call, link or import this; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker and as
 * the CPU baseline.
 *
 * Parity pinned: this restatement is checked against outputs of the real reference
 * (numba CPU path, run in the build container by oracle/run_reference.py) that are
 * committed under tests/golden/ (tests/test_oracle_golden.py): stage-1 intermediates of
 * every kernel plus the prognostic state after 1, 2 and 10 Matsuno steps.
 *
 * Scope: every namelist dyn switch = 1.  orc_grid.i_coupling == 0 is the "dry" configuration
 * of SURVEY.md section 0.4: physics modules off, coupling fields (KMOM, KHEAT, SMOM[XY]FLX,
 * SSHFLX, SLHFLX, dPOTTdt_RAD) exactly 0; the turbulence / radiation terms of the reference
 * evaluate to exactly +-0.0 then and are skipped.  With i_coupling != 0 they are evaluated on
 * the given fields (pinned on tests/golden/ref_10deg_coupled.npz), and orc_compute_turbulence
 * is the reference's turbulence module (pinned on ref_10deg_turb.npz).
 *
 * All arrays are in the REFERENCE layout: C-contiguous (i=lon, j=lat, k=level), k fastest,
 * one halo cell (nb=1) each side in i and j (main_fields.py:477-485):
 *   mass points      (nx+2, ny+2, nz)      x-staggered  (nx+3, ny+2, nz)
 *   y-staggered      (nx+2, ny+3, nz)      xy-staggered (nx+3, ny+3, nz)
 *   interface fields (..., nz+1)           2-D fields   (..., 1)
 */
#ifndef DYNCORE_ORACLE_H
#define DYNCORE_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int nx, ny, nz;             /* nb == 1 (io_read_namelist.py:30-31) */
    int i_moist;                /* namelist.i_moist_main_switch (QV/QC advected when != 0) */
    double dt;                  /* GR.dt (python int in the reference, main_grid.py:263) */
    double pair_top;            /* namelist.py:33 */
    /* GRF grid fields (main_grid.py:304-309) */
    const double *A;            /* (nx+2, ny+2) */
    const double *dxjs;         /* (nx+2, ny+3) */
    const double *dyis;         /* (nx+3, ny+2) */
    const double *corf;         /* (nx+2, ny+2) */
    const double *corf_is;      /* (nx+3, ny+2) */
    const double *lat_rad;      /* (nx+2, ny+2) */
    const double *lat_is_rad;   /* (nx+3, ny+2) */
    const double *dlon_rad;     /* (nx+2, ny+3) constant field */
    const double *dlat_rad;     /* (nx+3, ny+2) constant field */
    const double *sigma_vb;     /* (nz+1) */
    const double *dsigma;       /* (nz) */
    const double *UVFLX_dif_coef, *POTT_dif_coef, *moist_dif_coef; /* (nz) */
    /* != 0: the physics coupling fields (KMOM, KHEAT, surface fluxes) enter the tendencies
     * (the reference always evaluates these terms; with zero fields they are exactly 0,
     * which is what i_coupling == 0 assumes) */
    int i_coupling;
} orc_grid;

/* every model field the dry dyn core touches (main_fields.py:233-330) */
typedef struct {
    double *COLP, *COLP_OLD, *COLP_NEW, *dCOLPdt, *HSURF;
    double *UWIND, *UWIND_OLD, *VWIND, *VWIND_OLD, *WWIND;
    double *POTT, *POTT_OLD, *QV, *QV_OLD, *QC, *QC_OLD;
    double *UFLX, *VFLX, *FLXDIV;
    double *BFLX, *CFLX, *DFLX, *EFLX, *RFLX, *QFLX, *SFLX, *TFLX;
    double *WWIND_UWIND, *WWIND_VWIND;
    double *dUFLXdt, *dVFLXdt, *dPOTTdt, *dQVdt, *dQCdt;
    double *PHI, *PHIVB, *PVTF, *PVTFVB, *POTTVB;
    /* secondary diagnostics (dyn_diagnostics.py:199-222) */
    double *TAIR, *TAIRVB, *PAIR, *PAIRVB, *RHO, *RHOVB, *WINDX, *WINDY, *WIND;
    /* physics coupling (inputs: main_fields.py KMOM, KHEAT (nzs), surface fluxes (2-D);
     * outputs of the dyn core: KMOM_dUWINDdz / KMOM_dVWINDdz (nzs), *_TURB tendencies) */
    double *KMOM, *KHEAT, *SMOMXFLX, *SMOMYFLX, *SSHFLX, *SLHFLX;
    double *KMOM_dUWINDdz, *KMOM_dVWINDdz;
    double *dUFLXdt_TURB, *dVFLXdt_TURB, *dPOTTdt_TURB, *dQVdt_TURB;
    /* radiative heating rate [K s-1], input from the radiation module (dyn_POTT.py:107-108) */
    double *dPOTTdt_RAD;
} orc_fields;

/* misc_boundaries.py:22-42 ; (fnx,fny,fnz) = shape of FIELD */
void orc_exchange_BC(const orc_grid *g, double *FIELD, int fnx, int fny, int fnz);

/* dyn_continuity.py:170-228 + BCs of dyn_org_discretizations.py:114-117 */
void orc_continuity(const orc_grid *g, orc_fields *f);
/* dyn_org_discretizations.py:121-249 (prep + UFLX + VFLX tendencies) */
void orc_momentum(const orc_grid *g, orc_fields *f);
/* dyn_POTT.py:180-216 */
void orc_temperature(const orc_grid *g, orc_fields *f);
/* dyn_moist.py:201-243 */
void orc_moisture(const orc_grid *g, orc_fields *f);
/* dyn_tendencies.py:25-72 */
void orc_compute_tendencies(const orc_grid *g, orc_fields *f);
/* dyn_timestep.py:212-296 + BCs of dyn_org_discretizations.py:388-393 */
void orc_euler_forward(const orc_grid *g, orc_fields *f);
/* dyn_diagnostics.py:139-195 */
void orc_primary_diag(const orc_grid *g, orc_fields *f);
/* dyn_diagnostics.py:199-222 */
void orc_secondary_diag(const orc_grid *g, orc_fields *f);
/* turb_main.py:38-50 / turb_compute.py:190-204: KMOM, KHEAT (the reference's turbulence module) */
void orc_compute_turbulence(const orc_grid *g, orc_fields *f);
/* dyn_matsuno.py:28-129 */
void orc_step_matsuno(const orc_grid *g, orc_fields *f);

/* number of OpenMP threads the library will use (1 if built without OpenMP) */
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
