/*
 * dyncore.h -- C ABI of libdyncore.so, the B200 (sm_100a) dynamical core.
 *
 * This is the drop-in boundary for the reference's dynamical-core path: it replaces the
 * numba kernels behind the `target`-keyed factories of dyn_org_discretizations.py
 * (TendencyFactory / DiagnosticsFactory / PrognosticsFactory, :75-393) and the Matsuno
 * stepper dyn_matsuno.py:28-129.  The reference has no C FFI for this path (it is pure
 * Python + numba); a ctypes binding of these entry points is what a maintainer adds --
 * see INTEGRATION.md and climate_model_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers and sizes only, no torch / C++ types.
 *   - every call returns 0 on success, a NEGATIVE dc_status on an argument / state error,
 *     a POSITIVE cudaError_t / ncclResult_t (+1000) on a runtime error; dc_last_error()
 *     gives the message of the last failure on the calling thread.
 *   - compute calls enqueue work on the given cudaStream_t (passed as void*, 0 = default
 *     stream) and return without synchronising; one handle per GPU / rank; a handle is
 *     not thread-safe.
 *   - device buffers are OWNED BY THE CALLER (torch tensors on the Python side); the
 *     library never allocates or frees field memory, it only keeps the bound pointers.
 *
 * Device field layout (dc_get_layout): F[k][jd][i], longitude fastest, row pitch NI,
 * NJ rows per plane, device row jd = global reference row j + jshift; 3-D fields have nz
 * (or nz+1) planes, 2-D fields one.  The reference layout is (i, j, k) with k fastest
 * (main_fields.py:477-485); climate_model_b200/main_fields.py converts.
 */
#ifndef DYNCORE_H
#define DYNCORE_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dc_handle dc_handle;

typedef enum {
    DC_OK = 0,
    DC_ERR_ARG = -1,       /* NULL / out-of-range argument                     */
    DC_ERR_SHAPE = -2,     /* buffer too small / grid field not i-invariant    */
    DC_ERR_UNBOUND = -3,   /* a field the entry needs has not been bound       */
    DC_ERR_STATE = -4,     /* call not valid in this configuration             */
    DC_ERR_NO_DEVICE = -5  /* no CUDA device / wrong architecture              */
} dc_status;

/*
 * Grid description = what the reference kernels receive as GR.GRF[target] + GR.dt
 * (main_grid.py:298-315).  All pointers are HOST arrays in the reference layout:
 * 2-D fields (fnx, fny, 1) C-contiguous, 1-D fields (1, 1, nz[+1]).
 */
typedef struct {
    int nx, ny, nz;          /* main_grid.py:45-52 ; nb == 1                             */
    int j0, j1;              /* global mass rows owned by this rank, 1 <= j0 <= j1 <= ny;
                                single GPU: 1, ny                                        */
    int i_moist;             /* namelist.i_moist_main_switch                             */
    double dt;               /* GR.dt                                                    */
    double pair_top;         /* namelist.pair_top                                        */
    const double *A;         /* (nx+2, ny+2)                                             */
    const double *dxjs;      /* (nx+2, ny+3)                                             */
    const double *dyis;      /* (nx+3, ny+2)                                             */
    const double *corf;      /* (nx+2, ny+2)                                             */
    const double *corf_is;   /* (nx+3, ny+2)                                             */
    const double *lat_rad;   /* (nx+2, ny+2)                                             */
    const double *lat_is_rad;/* (nx+3, ny+2)                                             */
    const double *dlon_rad;  /* (nx+2, ny+3)                                             */
    const double *dlat_rad;  /* (nx+3, ny+2)                                             */
    const double *sigma_vb;  /* (nz+1)                                                   */
    const double *dsigma;    /* (nz)                                                     */
    const double *UVFLX_dif_coef, *POTT_dif_coef, *moist_dif_coef; /* (nz)               */
    int i_coupling;          /* != 0: the physics coupling fields enter the tendencies:
                                vertical turbulent transport of momentum / heat / moisture
                                with KMOM, KHEAT (dyn_functions.py:26-67, :276-422,
                                dyn_UFLX.py:136-170, dyn_VFLX.py:134-166) and the surface
                                fluxes SMOMXFLX, SMOMYFLX, SSHFLX, SLHFLX.  The reference
                                always evaluates these terms; with zero fields they are
                                exactly 0, which is what i_coupling == 0 assumes.  The
                                coupled terms run in the kernel decomposition
                                (DC_MODE_KERNELS) on one device; RHO / RHOVB are the
                                caller's (dc_secondary_diag once per time step, as
                                solver.py:99-101).                                       */
} dc_grid_desc;

/* kind of vertical extent of a field (dc_field_info) */
enum { DC_NK_2D = 0, DC_NK_NZ = 1, DC_NK_NZS = 2 };

const char *dc_last_error(void);
/* 1 when this library runs the kernels on a CUDA device (the product), 0 for the host
 * emulation harness that the CPU tests build from the same kernel bodies. */
int dc_is_cuda(void);

/* Grid() + GRF upload (main_grid.py:298-315): builds the per-row / per-level geometry
 * on the device.  The horizontal grid fields must depend on latitude only. */
int dc_create(const dc_grid_desc *desc, dc_handle **out);
int dc_destroy(dc_handle *h);

/* device layout of every field of this handle */
int dc_get_layout(const dc_handle *h, int *NI, int *NJ, int *jshift);

/* field registry = the dyn-core subset of main_fields.py:233-470 (fdict) */
int dc_num_fields(void);
const char *dc_field_name(int field_id);             /* NULL if out of range           */
int dc_field_id(const char *name);                   /* -1 if unknown                  */
int dc_field_info(int field_id, int *stgx, int *stgy, int *nk_kind);

/* F.device[name] = buffer (main_fields.py:204-208): nbytes must be >= planes*NJ*NI*8 */
int dc_bind_field(dc_handle *h, int field_id, void *devptr, size_t nbytes);

/* ---- fine-grained entries: one per factory method of dyn_org_discretizations.py ---- */
/* TendencyFactory.continuity  (:89-117)  incl. the four exchange_BC calls               */
int dc_continuity(dc_handle *h, void *stream);
/* TendencyFactory.momentum    (:121-249) prep + UFLX + VFLX tendencies                  */
int dc_momentum(dc_handle *h, void *stream);
/* TendencyFactory.temperature (:253-273)                                                */
int dc_temperature(dc_handle *h, void *stream);
/* TendencyFactory.moisture    (:277-294)                                                */
int dc_moisture(dc_handle *h, void *stream);
/* dyn_tendencies.compute_tendencies (dyn_tendencies.py:25-72)                           */
int dc_compute_tendencies(dc_handle *h, void *stream);
/* PrognosticsFactory.euler_forward (:359-393) incl. the exchange_BC calls               */
int dc_euler_forward(dc_handle *h, void *stream);
/* DiagnosticsFactory.primary_diag (:310-326)                                            */
int dc_primary_diag(dc_handle *h, void *stream);
/* DiagnosticsFactory.secondary_diag (:329-346)                                          */
int dc_secondary_diag(dc_handle *h, void *stream);
/* Turbulence.compute_turbulence (turb_main.py:38-50, turb_compute.py:53-204): KMOM, KHEAT
 * on the interior interfaces from PHIVB, HSURF, PHI, QV, WINDX, WINDY, POTTVB, POTT (call
 * after dc_secondary_diag, as solver.py:99-112); feeds the i_coupling terms               */
int dc_compute_turbulence(dc_handle *h, void *stream);
/* misc_boundaries.exchange_BC (misc_boundaries.py:22-42) on one bound field             */
int dc_exchange_bc(dc_handle *h, int field_id, void *stream);

/* ---- run-time diagnostics on the device (io_functions.py:70-114, diagnose_print_diag_fields
 *      + the crash check of print_ts_info).  `scratch` is a caller-owned device buffer of
 *      dc_run_diag_bytes() bytes; on return (stream order) its LAST 7*NJ doubles hold, per
 *      device row jd of this rank's band (other rows untouched), the row sums
 *        [0] sum WIND*COLP*A  [1] sum POTT*COLP*A  [2] sum COLP*A  [3] sum A
 *        [4] max WIND         [5] max UWIND        [6] number of NaNs in UWIND
 *      over the interior longitudes and all levels (WIND as in diag_secondary); the caller
 *      adds the rows (and the ranks).  64 bytes per row instead of three full fields. ---- */
int dc_run_diag_bytes(const dc_handle *h, size_t *nbytes);
int dc_run_diag(dc_handle *h, void *scratch, size_t nbytes, void *stream);

/* ---- coarse entry: dyn_matsuno.step_matsuno (dyn_matsuno.py:28-129), nsteps times ----
 * DC_MODE_FUSED (default): per stage one continuity kernel, one fused stage kernel
 *   (momentum-flux preparation + U/V/POTT tendencies + Euler step + boundary images) and
 *   one diagnostics kernel.  The *_OLD 3-D fields are used as the second state buffer:
 *   after the call they hold the stage-1 estimate, not the state before the step; the
 *   intermediate fields (UFLX, BFLX.., dUFLXdt..) are not written, and of the diagnostics
 *   only PHI, POTTVB and the work field PGCOL are kept current between stages: PVTF,
 *   PVTFVB and PHIVB are refreshed by dc_primary_diag (and implicitly by dc_momentum,
 *   dc_secondary_diag and kernel-mode stepping) before they are read.
 * DC_MODE_KERNELS: the reference's decomposition, one kernel per reference kernel, every
 *   intermediate field written as the reference does (same result, bit for bit). */
enum { DC_MODE_FUSED = 0, DC_MODE_KERNELS = 1 };
int dc_set_mode(dc_handle *h, int mode);
/* Development switches, read from the environment by dc_create (defaults = the measured best):
 *   DC_STAGE_KCHUNKS=n  sigma-column chunks of the stage kernel (default: by launch size)
 *   DC_CONT_IMPL=1      two-sweep column continuity kernel instead of the single-pass tile kernel
 *   DC_COUPLED_IMPL=1|2 i_coupling: 1 = kernel decomposition (reference summation order; default
 *                       of the strict build), 2 = coupled terms as increments beside the fused
 *                       dry stage kernel (default of the production build) */
int dc_step_matsuno(dc_handle *h, int nsteps, void *stream);

/* ---- latitude-band decomposition (one handle per rank, dc_grid_desc.j0 / j1) ----------
 * A Matsuno stage on a band = dc_stage_compute (continuity on rows j0-1..j1+1, fused stage
 * kernel on rows j0..j1, COLP <- COLP_NEW), then the exchange of the two boundary rows of the
 * new U, V, POTT (QV, QC) and COLP with the neighbouring ranks (dc_halo_pack -> NCCL
 * send/recv by the caller -> dc_halo_unpack), then dc_stage_diag on all held rows.
 * dc_step_begin once per step.  Messages are contiguous buffers of dc_halo_bytes() bytes per
 * direction; pass NULL for a direction that ends at a domain wall.  The result is bitwise
 * identical to the single-device run (no cross-rank reductions).
 * Overlap: a stage can be issued in pieces -- DC_PART_CONT (continuity + moisture stage),
 * DC_PART_BOUNDARY (stage kernel on the first / last tile row: what the neighbours wait for),
 * DC_PART_INTERIOR (the other tile rows), DC_PART_COLP (COLP <- COLP_NEW).  BOUNDARY and
 * INTERIOR only depend on CONT and may run concurrently on two streams; the caller packs and
 * sends after BOUNDARY while INTERIOR runs.  COLP must follow both (they read COLP), and
 * dc_halo_unpack must follow COLP (it writes the halo rows of COLP). */
enum { DC_PART_ALL = 0, DC_PART_CONT = 1, DC_PART_BOUNDARY = 2, DC_PART_INTERIOR = 3,
       DC_PART_COLP = 4 };
int dc_step_begin(dc_handle *h, void *stream);
int dc_stage_compute(dc_handle *h, int stage, int part, void *stream);
int dc_stage_diag(dc_handle *h, int stage, void *stream);
int dc_halo_bytes(const dc_handle *h, size_t *nbytes);
int dc_halo_pack(dc_handle *h, int stage, void *send_south, void *send_north, void *stream);
int dc_halo_unpack(dc_handle *h, int stage, const void *recv_south, const void *recv_north,
                   void *stream);

/* ---- in-library halo exchange (dc_set_comm / dc_halo_exchange) -----------
 * The library owns an NCCL communicator over the ranks of the band decomposition (libnccl.so.2
 * is loaded on first use; a single-GPU process never needs it).  Rank 0 obtains a unique id
 * with dc_comm_unique_id, the caller distributes its DC_COMM_ID_BYTES bytes to every rank by
 * any means (torch.distributed.broadcast in parallel_bands.py) and every rank calls
 * dc_set_comm -- a collective call.  With a communicator attached
 *   - dc_halo_exchange(h, stage, stream) = pack -> grouped ncclSend/ncclRecv with the south
 *     and north neighbours -> unpack, all enqueued on `stream`;
 *   - dc_step_matsuno works on a band: per stage, continuity -> the boundary tile rows, the
 *     packing and the NCCL exchange on an internal high-priority stream while the interior
 *     tile rows run on `stream` -> COLP <- COLP_NEW -> diagnostics of the rows that do not
 *     depend on the neighbours WHILE the halo is in flight -> unpack -> diagnostics of the
 *     halo rows.  No host round trip, no Python between the stages.  After the first step
 *     the sequence of one step is replayed from a CUDA graph (DC_BAND_GRAPH=0 disables).
 * The result is bitwise identical to the single-device run. */
#define DC_COMM_ID_BYTES 128
int dc_comm_unique_id(void *id, size_t nbytes);
int dc_set_comm(dc_handle *h, const void *id, size_t nbytes, int rank, int nranks);
int dc_has_comm(const dc_handle *h);
int dc_halo_exchange(dc_handle *h, int stage, void *stream);
/* Peer-memory exchange (optional, after dc_set_comm): the boundary rows travel as ONE copy-engine
 * transfer per neighbour straight into the neighbour's receive buffer over NVLink (CUDA IPC
 * mapping), announced by a stream memory operation on a flag in the neighbour's memory; the
 * receiver waits on its own flag in stream order.  No NCCL kernel has to find a free SM beside
 * the stage kernel (measured on 8 B200: 170 us per NCCL group under load, the longest link of
 * the band step's critical chain).  Every rank publishes dc_comm_p2p_handles (DC_P2P_HANDLE_BYTES
 * bytes), the caller distributes them, every rank connects to its neighbours' (NULL where the
 * band ends at a wall) and, once all ranks report success, switches it on (dc_comm_p2p_enable).  Receive buffers and flags are double-buffered by stage parity, so the
 * constant flag values 1 / 0 suffice and the step can be replayed from a CUDA graph. */
#define DC_P2P_HANDLE_BYTES 256
int dc_comm_p2p_handles(dc_handle *h, void *out, size_t nbytes);
int dc_comm_p2p_connect(dc_handle *h, const void *south, const void *north, size_t nbytes);
/* switch the peer-memory exchange on (after EVERY rank has connected: the caller agrees on that
 * collectively, a rank that could not map its neighbours must make all ranks stay on NCCL) or off */
int dc_comm_p2p_enable(dc_handle *h, int on);

/* ---- layout conversion on the device (F.copy_host_to_device / copy_device_to_host,
 *      main_fields.py:204-215): `ref` is a DEVICE buffer holding the field in the
 *      reference layout (fnx, fny, nk) with k fastest, i.e. a raw byte copy of the host
 *      array; import transposes it into the bound field (rows this rank holds), export
 *      does the reverse.  nbytes = size of `ref`. ---- */
int dc_import_field(dc_handle *h, int field_id, const void *ref, size_t nbytes, void *stream);
int dc_export_field(dc_handle *h, int field_id, void *ref, size_t nbytes, void *stream);
/* The same for a ROW WINDOW of the reference layout: `ref` holds rows ja..jb only, shape
 * (fnx, jb-ja+1, nk) with k fastest.  A rank of a latitude-band run keeps just the rows it
 * holds on the host (ModelFields(band_local=True)): nothing whole-grid sized is ever
 * allocated, on the host or on the device.  Rows of the window this rank does not hold are
 * left untouched. */
int dc_import_rows(dc_handle *h, int field_id, const void *ref, size_t nbytes, int ja, int jb,
                   void *stream);
int dc_export_rows(dc_handle *h, int field_id, void *ref, size_t nbytes, int ja, int jb,
                   void *stream);

/* number of kernel launches this handle has enqueued so far (bench.py: gpu_launches) */
long long dc_launch_count(const dc_handle *h);

/* ---- per-kernel device timing (bench.py roofline): when enabled, every kernel launch is
 *      bracketed by CUDA events on its stream.  dc_profile_read synchronises the device,
 *      adds up the elapsed times per kernel and resets the event list.
 *      names[i] / ms[i] / launches[i] for i < return value (<= max_entries).
 *      dc_profile_enable(h, 2) = timeline mode for the banded dc_step_matsuno: instead of the
 *      brackets, a timing event is recorded at every hand-over of the step (after the
 *      continuity, the boundary tile rows, the pack, the NCCL group, ...) on the stream it
 *      happens on; dc_profile_read then returns the marks in enqueue order with ms[i] = time
 *      since the first mark (launches[i] = i). ---- */
int dc_profile_enable(dc_handle *h, int on);
int dc_profile_read(dc_handle *h, int max_entries, const char **names, double *ms,
                    long long *launches);

#ifdef __cplusplus
}
#endif
#endif
