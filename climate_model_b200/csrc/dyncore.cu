// dyncore.cu -- libdyncore.so: the B200 (sm_100a) dynamical core behind include/dyncore.h.
// CUDA backend of dc_api_impl.h: every kernel body of dc_kernels.h is wrapped in a
// __global__ kernel with one thread per (lon, lat) column, longitude along threadIdx.x
// so that each level's loads and stores coalesce.
//
// Build (see __graft_entry__.build):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared ...
// -fmad=false keeps the reference's evaluation order (numba emits no FMA contraction).
#include <cuda.h>
#include <cuda_runtime.h>

#include "dc_geom.h"
#include "dc_point.h"

#define DC_BACKEND_IS_CUDA 1

static int dcb_malloc(void **p, size_t n) { return (int)cudaMalloc(p, n); }
static int dcb_free(void *p) { return (int)cudaFree(p); }
static int dcb_h2d(void *dst, const void *src, size_t n)
{
    return (int)cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
}
static int dcb_d2d_async(void *dst, const void *src, size_t n, void *stream)
{
    return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
}
static int g_stage3_error = 0;   // descriptor failures of the stage-3 launcher (sticky)
static int dcb_last_error()
{
    if (g_stage3_error) {
        const int e = g_stage3_error;
        g_stage3_error = 0;
        return e;
    }
    return (int)cudaGetLastError();
}
static const char *dcb_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

namespace dc {
constexpr int BX = 64, BY = 4;  // 256 threads: 64 consecutive longitudes x 4 rows

template <class Body>
__global__ void __launch_bounds__(BX *BY) k_columns(const Body b, int i0, int i1, int j0, int j1)
{
    const int i = i0 + blockIdx.x * BX + threadIdx.x;
    const int j = j0 + blockIdx.y * BY + threadIdx.y;
    if (i <= i1 && j <= j1) b(i, j);
}
}  // namespace dc

template <class Body>
static void dcb_launch(const Body &b, int i0, int i1, int j0, int j1, void *stream)
{
    dim3 block(dc::BX, dc::BY);
    dim3 grid((i1 - i0 + dc::BX) / dc::BX, (j1 - j0 + dc::BY) / dc::BY);
    dc::k_columns<Body><<<grid, block, 0, (cudaStream_t)stream>>>(b, i0, i1, j0, j1);
}

// primary diagnostics: thread per column like k_columns, with its own block shape and register
// budget (DC_DIAG_BY rows of 64 longitudes per block, DC_DIAG_MINB blocks per SM): the sweep is
// bound by latency, so resident warps are what the build switches trade against spills.
// A dedicated kernel that staged the Exner power table (7 KB) and the per-level vectors in
// shared memory was measured and dropped: 0.978-0.995 against 0.967 ms per step -- the sweep's
// `long_scoreboard` stalls are not those look-ups.
#ifndef DC_DIAG_BY
#define DC_DIAG_BY 4
#endif
#ifndef DC_DIAG_MINB
#define DC_DIAG_MINB 4
#endif
namespace dc {
template <class Body>
__global__ void __launch_bounds__(BX *DC_DIAG_BY, DC_DIAG_MINB)
    k_diag(const Body b, int i0, int i1, int j0, int j1)
{
    const int i = i0 + blockIdx.x * BX + threadIdx.x;
    const int j = j0 + blockIdx.y * DC_DIAG_BY + threadIdx.y;
#if DC_DIAG_PF > 0
    __shared__ double ring[DIAG_SLOTS][BX * DC_DIAG_BY];
    if (i <= i1 && j <= j1)
        b.march(i, j, b.pc.tab, false, nullptr, &ring[0][threadIdx.y * BX + threadIdx.x],
                BX * DC_DIAG_BY);
#else
    if (i <= i1 && j <= j1) b(i, j);
#endif
}
}  // namespace dc
template <class Body>
static void dcb_launch_diag(const Body &b, int i0, int i1, int j0, int j1, void *stream)
{
    dim3 block(dc::BX, DC_DIAG_BY);
    dim3 grid((i1 - i0 + dc::BX) / dc::BX, (j1 - j0 + DC_DIAG_BY) / DC_DIAG_BY);
    dc::k_diag<Body><<<grid, block, 0, (cudaStream_t)stream>>>(b, i0, i1, j0, j1);
}

namespace dc {
#ifndef DC_CT_MINB
#define DC_CT_MINB 2
#endif
template <class Body, class Smem>
__global__ void __launch_bounds__(256, DC_CT_MINB) k_blocks(const Body b)
{
    __shared__ Smem s;
    b.run_block(blockIdx.x, blockIdx.y, s);
}
}  // namespace dc
template <class Body, class Smem>
static void dcb_launch_blocks(const Body &b, int nbx, int nby, int nthreads, void *stream)
{
    dc::k_blocks<Body, Smem><<<dim3(nbx, nby), dim3(nthreads), 0, (cudaStream_t)stream>>>(b);
}

// cudaFuncSetAttribute is per device: remember which devices a kernel has been configured on
static bool first_use_on_device(unsigned long long *mask)
{
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long bit = 1ull << (dev & 63);
    if (*mask & bit) return false;
    *mask |= bit;
    return true;
}

// ---------------------------------------------------------------------------------------
// third-generation stage kernel (dc_stage3.h): TMA descriptors + launch
// ---------------------------------------------------------------------------------------
#include <map>
#include <utility>
#include "dc_stage3.h"
struct dc_handle;
namespace dc {
struct Stage3Ptrs;
#ifndef DC_S3_MINBLOCKS
#define DC_S3_MINBLOCKS 2
#endif
__global__ void __launch_bounds__(S3_NT, DC_S3_MINBLOCKS) k_stage3(const __grid_constant__ Stage3Body b)
{
    extern __shared__ unsigned char stage3_smem[];
    // the TMA destinations need 128-byte alignment
    const unsigned a = (unsigned)__cvta_generic_to_shared(stage3_smem);
    Stage3Smem &s = *reinterpret_cast<Stage3Smem *>(stage3_smem + ((128u - (a & 127u)) & 127u));
    b.run_block(blockIdx.x, blockIdx.y, blockIdx.z, s);
}
}  // namespace dc
#include "dc_moist3.h"
namespace dc {
struct Moist3Ptrs;
__global__ void __launch_bounds__(S3_NT, 3) k_moist3(const __grid_constant__ Moist3Body b)
{
    extern __shared__ unsigned char moist3_smem[];
    const unsigned a = (unsigned)__cvta_generic_to_shared(moist3_smem);
    Moist3Smem &s = *reinterpret_cast<Moist3Smem *>(moist3_smem + ((128u - (a & 127u)) & 127u));
    b.run_block(blockIdx.x, blockIdx.y, blockIdx.z, s);
}
struct TmaState {
    std::map<std::pair<const void *, int>, CUtensorMap> maps;   // (field base, own box?)
};
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}
}  // namespace dc
static void dcb_launch_stage3(dc_handle *h, dc::Stage3Body &b, const dc::Stage3Ptrs &p, int nbx,
                              int nby, void *stream);
static void dcb_tma_release(dc_handle *h);
static void dcb_launch_moist3(dc_handle *h, dc::Moist3Body &b, const dc::Moist3Ptrs &p, int nbx,
                              int nby, void *stream);

// ---------------------------------------------------------------------------------------
// layout conversion: reference (i, j, k) k-fastest  <->  device F[k][jd][i] i-fastest.
// For every row j a tiled (i, k) transpose through shared memory so that both the read
// and the write side are coalesced.
// ---------------------------------------------------------------------------------------
namespace dc {
constexpr int TT = 32, TR = 8;

template <bool TO_DEVICE>
__global__ void __launch_bounds__(TT *TR)
    k_transpose(const Geom g, double *__restrict__ ref, double *__restrict__ dev, int fnx,
                int fny, int nk, int j_lo)
{
    __shared__ double tile[TT][TT + 1];
    const int j = j_lo + blockIdx.z;
    const int i0 = blockIdx.x * TT, k0 = blockIdx.y * TT;
    const int tx = threadIdx.x, ty = threadIdx.y;
    if (TO_DEVICE) {
        for (int r = ty; r < TT; r += TR) {  // read ref: k along threadIdx.x
            const int i = i0 + r, k = k0 + tx;
            if (i < fnx && k < nk) tile[r][tx] = ref[((size_t)i * fny + j) * nk + k];
        }
        __syncthreads();
        for (int r = ty; r < TT; r += TR) {  // write dev: i along threadIdx.x
            const int i = i0 + tx, k = k0 + r;
            if (i < fnx && k < nk) dev[g.idx(i, j, k)] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < TT; r += TR) {
            const int i = i0 + tx, k = k0 + r;
            if (i < fnx && k < nk) tile[r][tx] = dev[g.idx(i, j, k)];
        }
        __syncthreads();
        for (int r = ty; r < TT; r += TR) {
            const int i = i0 + r, k = k0 + tx;
            if (i < fnx && k < nk) ref[((size_t)i * fny + j) * nk + k] = tile[tx][r];
        }
    }
}
}  // namespace dc

static void dcb_transpose(const dc::Geom &g, double *ref, double *dev, int fnx, int fny, int nk,
                          int j_lo, int j_hi, int to_device, void *stream)
{
    if (j_hi < j_lo) return;
    dim3 block(dc::TT, dc::TR);
    dim3 grid((fnx + dc::TT - 1) / dc::TT, (nk + dc::TT - 1) / dc::TT, j_hi - j_lo + 1);
    if (to_device)
        dc::k_transpose<true><<<grid, block, 0, (cudaStream_t)stream>>>(g, ref, dev, fnx, fny, nk,
                                                                        j_lo);
    else
        dc::k_transpose<false><<<grid, block, 0, (cudaStream_t)stream>>>(g, ref, dev, fnx, fny,
                                                                         nk, j_lo);
}

// ---------------------------------------------------------------------------------------
// per-kernel timing with CUDA events on the launching stream
// ---------------------------------------------------------------------------------------
#include <map>
#include <string>
#include <vector>
struct dc_handle;
namespace dc {
struct ProfileRec {
    const char *name;
    cudaEvent_t a, b;
};
struct ProfileState {
    std::vector<ProfileRec> recs;
    std::vector<cudaEvent_t> pool;
    std::vector<std::string> names;  // storage for dc_profile_read's returned strings
    std::vector<std::pair<const char *, cudaEvent_t>> marks;   // timeline (dc_profile_enable(h, 2))
    cudaEvent_t get()
    {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};
}  // namespace dc
static void dcb_profile_begin(dc_handle *h, const char *name, void *stream);
static void dcb_profile_end(dc_handle *h, void *stream);
static void dcb_mark(dc_handle *h, const char *name, void *stream);
static int dcb_profile_read(dc_handle *h, int max, const char **names, double *ms, long long *n);

// in-library halo exchange: NCCL communicator, side stream, events, CUDA graph (defined below)
static int dcb_comm_unique_id(void *id128);
static int dcb_comm_init(dc_handle *h, const void *id128, int rank, int nranks, size_t halo_elems);
static void dcb_comm_release(dc_handle *h);
static double *dcb_comm_buffer(dc_handle *h, int which, int stage);
static int dcb_comm_sendrecv(dc_handle *h, int stage, void *stream);
static void dcb_comm_consumed(dc_handle *h, int stage, void *stream);
static int dcb_comm_p2p_handles(dc_handle *h, void *out);
static int dcb_comm_p2p_connect(dc_handle *h, const void *south, const void *north);
static int dcb_comm_p2p_enable(dc_handle *h, int on);
static void *dcb_side_stream(dc_handle *h, int which = 0);
static void dcb_event_record(dc_handle *h, int ev, void *stream);
static void dcb_stream_wait(dc_handle *h, int ev, void *stream);
static int dcb_graph_steps(dc_handle *h, int nsteps, void *stream,
                           void (*prologue)(dc_handle *, int, void *),
                           void (*step_tail)(dc_handle *, void *),
                           void (*step_last)(dc_handle *, void *));
static const char *dcb_comm_error();

#include "dc_api_impl.h"

// ---------------------------------------------------------------------------------------
// NCCL, loaded on first use: a single-GPU process never touches it, and a process that has
// torch.distributed loaded gets the very same libnccl.so.2 (same SONAME).  Only the eight entry
// points below are used; their prototypes follow nccl.h (2.x ABI).
// ---------------------------------------------------------------------------------------
#include <dlfcn.h>
namespace dc {
struct NcclUniqueId { char internal[128]; };
typedef void *NcclComm;
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
constexpr int NCCL_FLOAT64 = 8;   // ncclDataType_t::ncclFloat64
static thread_local std::string g_comm_error;
static NcclApi *nccl_api()
{
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            g_comm_error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return nullptr;
        }
        api.lib = lib;
#define DC_NCCL_SYM(field, name) \
        *reinterpret_cast<void **>(&api.field) = dlsym(lib, name); \
        if (!api.field) { g_comm_error = std::string("libnccl: missing symbol ") + name; api.lib = nullptr; }
        DC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
        DC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
        DC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
        DC_NCCL_SYM(Send, "ncclSend")
        DC_NCCL_SYM(Recv, "ncclRecv")
        DC_NCCL_SYM(GroupStart, "ncclGroupStart")
        DC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
        DC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef DC_NCCL_SYM
    }
    return api.lib ? &api : nullptr;
}
struct CommState {
    NcclComm comm = nullptr;
    int rank = 0, nranks = 1;
    size_t nelem = 0;
    double *buf[4] = {nullptr, nullptr, nullptr, nullptr};   // send_s, recv_s, send_n, recv_n
                                                             // (recv: two slots, by stage parity)
    // peer-memory exchange (CUDA IPC): the neighbours' receive buffers and flags mapped here
    bool p2p = false, p2p_connected = false;
    unsigned *flags = nullptr;                 // [from south, from north][slot]: 1 = landed
    double *peer_recv[2] = {nullptr, nullptr}; // south neighbour's recv_n, north neighbour's recv_s
    unsigned *peer_flags[2] = {nullptr, nullptr};
    void *peer_base[4] = {nullptr, nullptr, nullptr, nullptr};   // cudaIpcOpenMemHandle results
    cudaStream_t side = nullptr, side2 = nullptr;
    cudaStream_t main = nullptr;   // stands in for the caller's stream when that is the legacy
                                   // default stream, which cannot be captured
    cudaEvent_t ev[16] = {};
    cudaGraphExec_t graph = nullptr, graph_last = nullptr;
    long long graph_version = -1;
    long long launches_per_step = 0, launches_last_step = 0;
    bool warmed = false;   // one plain step has run (NCCL connections, kernel attributes)
    int error = 0;
};
}  // namespace dc
static const char *dcb_comm_error() { return dc::g_comm_error.c_str(); }
static int nccl_fail(const char *what, int code)
{
    dc::NcclApi *a = dc::nccl_api();
    dc::g_comm_error = std::string(what) + ": " +
                       (a && a->GetErrorString ? a->GetErrorString(code) : "NCCL error");
    return code > 0 ? code : 1;
}
static int dcb_comm_unique_id(void *id128)
{
    dc::g_comm_error.clear();
    dc::NcclApi *a = dc::nccl_api();
    if (!a) return DC_ERR_NO_DEVICE;
    dc::NcclUniqueId id;
    const int e = a->GetUniqueId(&id);
    if (e) return nccl_fail("ncclGetUniqueId", e);
    memcpy(id128, &id, sizeof id);
    return 0;
}
static int dcb_comm_init(dc_handle *h, const void *id128, int rank, int nranks, size_t halo_elems)
{
    dc::g_comm_error.clear();
    dc::NcclApi *a = nranks > 1 ? dc::nccl_api() : nullptr;
    if (nranks > 1 && !a) return DC_ERR_NO_DEVICE;
    dc::CommState *c = new dc::CommState();
    c->rank = rank;
    c->nranks = nranks;
    c->nelem = halo_elems;
    if (nranks > 1) {
        dc::NcclUniqueId id;
        memcpy(&id, id128, sizeof id);
        const int e = a->CommInitRank(&c->comm, nranks, id, rank);
        if (e) {
            delete c;
            return nccl_fail("ncclCommInitRank", e);
        }
    }
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = highest priority
    cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithPriority(&c->side2, cudaStreamNonBlocking, hi);
    cudaStreamCreateWithFlags(&c->main, cudaStreamNonBlocking);
    for (auto &ev : c->ev) cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    for (int n = 0; n < 4; n++) {
        const bool present = (n < 2) ? rank > 0 : rank < nranks - 1;
        const size_t slots = (n & 1) ? 2 : 1;
        if (present && cudaMalloc(&c->buf[n], slots * halo_elems * sizeof(double)) != cudaSuccess) {
            dc::g_comm_error = "cudaMalloc of the halo buffers failed";
            h->comm_state = c;
            dcb_comm_release(h);
            return (int)cudaErrorMemoryAllocation;
        }
    }
    if (cudaMalloc(&c->flags, 4 * sizeof(unsigned)) == cudaSuccess)
        cudaMemset(c->flags, 0, 4 * sizeof(unsigned));
    h->comm_state = c;
    return 0;
}
static void dcb_comm_release(dc_handle *h)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    if (!c) return;
    cudaDeviceSynchronize();
    if (c->graph) cudaGraphExecDestroy(c->graph);
    if (c->graph_last) cudaGraphExecDestroy(c->graph_last);
    for (void *p : c->peer_base)
        if (p) cudaIpcCloseMemHandle(p);
    if (c->flags) cudaFree(c->flags);
    for (double *b : c->buf)
        if (b) cudaFree(b);
    for (auto &ev : c->ev)
        if (ev) cudaEventDestroy(ev);
    if (c->side) cudaStreamDestroy(c->side);
    if (c->side2) cudaStreamDestroy(c->side2);
    if (c->main) cudaStreamDestroy(c->main);
    if (c->comm) {
        dc::NcclApi *a = dc::nccl_api();
        if (a) a->CommDestroy(c->comm);
    }
    delete c;
    h->comm_state = nullptr;
}
static double *dcb_comm_buffer(dc_handle *h, int which, int stage)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    // receive buffers: slot = stage parity with the peer-memory exchange, slot 0 with NCCL
    return c->buf[which] + (((which & 1) && c->p2p) ? (size_t)(stage & 1) * c->nelem : 0);
}
// stream memory operations of the driver API (no kernel, no SM): write / wait on a 32-bit word
typedef CUresult (*StreamMemOpFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
static StreamMemOpFn stream_memop(const char *name)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    return reinterpret_cast<StreamMemOpFn>(p);
}
static StreamMemOpFn write_value32()
{
    static StreamMemOpFn fn = stream_memop("cuStreamWriteValue32");
    return fn;
}
static StreamMemOpFn wait_value32()
{
    static StreamMemOpFn fn = stream_memop("cuStreamWaitValue32");
    return fn;
}
// fallback signal (DC_BAND_P2P_SIGNAL=kernel): a one-thread kernel stores the flag
__global__ void k_signal(volatile unsigned *flag, unsigned v)
{
    __threadfence_system();
    *flag = v;
}
// wait for a flag with a one-thread spin kernel: a stream wait-value operation took ~120 us to
// notice the flag on this system (timeline r2_timeline_proxy5: copies done at 0.179 ms, wait over
// at 0.300 ms on both ranks), a polling thread notices it within a microsecond once it runs.
// Bounded: ~2 s of polling, then the kernel traps (surfaces as a launch failure).
__global__ void k_wait_flag(const volatile unsigned *flag, unsigned v)
{
    const long long t0 = clock64();
    while (*flag < v) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
    __threadfence_system();
}
static bool wait_by_memop()
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("DC_BAND_P2P_WAIT");
        mode = (e && e[0] == 'm') ? 1 : 0;
    }
    return mode == 1;
}
static bool signal_by_kernel()
{
    static int mode = -1;
    if (mode < 0) {
        const char *e = getenv("DC_BAND_P2P_SIGNAL");
        mode = (e && e[0] == 'k') ? 1 : 0;
    }
    return mode == 1;
}
static void memop_check(dc::CommState *c, CUresult r, const char *what)
{
    if (r != CUDA_SUCCESS && !c->error) {
        c->error = (int)r;
        dc::g_comm_error = std::string(what) + " failed (CUresult " + std::to_string((int)r) + ")";
    }
}
static int dcb_comm_p2p_handles(dc_handle *h, void *out)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    dc::g_comm_error.clear();
    // [0..63] recv_s, [64..127] recv_n, [128..191] flags (zeros where a buffer does not exist)
    unsigned char *o = static_cast<unsigned char *>(out);
    memset(o, 0, DC_P2P_HANDLE_BYTES);
    void *ptrs[3] = {c->buf[1], c->buf[3], c->flags};
    for (int n = 0; n < 3; n++) {
        if (!ptrs[n]) continue;
        cudaIpcMemHandle_t hd;
        if (cudaIpcGetMemHandle(&hd, ptrs[n]) != cudaSuccess) {
            dc::g_comm_error = "cudaIpcGetMemHandle failed";
            return (int)cudaGetLastError();
        }
        memcpy(o + 64 * n, &hd, sizeof hd);
    }
    return 0;
}
static int dcb_comm_p2p_connect(dc_handle *h, const void *south, const void *north)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    dc::g_comm_error.clear();
    if (!write_value32() || !wait_value32()) {
        dc::g_comm_error = "cuStreamWriteValue32 / cuStreamWaitValue32 are not available";
        return DC_ERR_NO_DEVICE;
    }
    auto open = [&](const void *blob, int which, void **out) {
        cudaIpcMemHandle_t hd;
        memcpy(&hd, static_cast<const unsigned char *>(blob) + 64 * which, sizeof hd);
        return cudaIpcOpenMemHandle(out, hd, cudaIpcMemLazyEnablePeerAccess);
    };
    // my message to the SOUTH neighbour lands in ITS "from north" buffer (recv_n) and flags
    // [from north]; my message to the NORTH neighbour in its recv_s / flags[from south]
    if (south) {
        if (open(south, 1, &c->peer_base[0]) != cudaSuccess ||
            open(south, 2, &c->peer_base[1]) != cudaSuccess) {
            dc::g_comm_error = "cudaIpcOpenMemHandle (south neighbour) failed";
            return (int)cudaGetLastError();
        }
        c->peer_recv[0] = static_cast<double *>(c->peer_base[0]);
        c->peer_flags[0] = static_cast<unsigned *>(c->peer_base[1]) + 2;   // [from north][slot]
    }
    if (north) {
        if (open(north, 0, &c->peer_base[2]) != cudaSuccess ||
            open(north, 2, &c->peer_base[3]) != cudaSuccess) {
            dc::g_comm_error = "cudaIpcOpenMemHandle (north neighbour) failed";
            return (int)cudaGetLastError();
        }
        c->peer_recv[1] = static_cast<double *>(c->peer_base[2]);
        c->peer_flags[1] = static_cast<unsigned *>(c->peer_base[3]) + 0;   // [from south][slot]
    }
    c->p2p_connected = true;   // switched on by dcb_comm_p2p_enable once EVERY rank is connected
    return 0;
}
static int dcb_comm_p2p_enable(dc_handle *h, int on)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    if (on && !c->p2p_connected) {
        dc::g_comm_error = "the neighbours' buffers are not mapped (dc_comm_p2p_connect)";
        return DC_ERR_STATE;
    }
    c->p2p = on != 0;
    c->graph_version = -1;     // captured graphs hold the other exchange
    return 0;
}
static int dcb_comm_sendrecv(dc_handle *h, int stage, void *stream)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    cudaStream_t st = (cudaStream_t)stream;
    if (c->p2p) {
        const int slot = stage & 1;
        const size_t bytes = c->nelem * sizeof(double);
        for (int d = 0; d < 2; d++) {     // 0: to the south neighbour, 1: to the north
            if (!c->peer_recv[d]) continue;
            cudaMemcpyAsync(c->peer_recv[d] + (size_t)slot * c->nelem, c->buf[d == 0 ? 0 : 2], bytes,
                            cudaMemcpyDefault, st);
            if (signal_by_kernel())
                k_signal<<<1, 1, 0, st>>>(c->peer_flags[d] + slot, 1u);
            else
                memop_check(c, write_value32()((CUstream)st, (CUdeviceptr)(c->peer_flags[d] + slot), 1, 0),
                            "cuStreamWriteValue32 (peer flag)");
        }
        if (h->profiling == 2) dcb_mark(h, "S p2p copies + flags done", stream);
        for (int d = 0; d < 2; d++) {     // 0: from the south neighbour, 1: from the north
            if (!c->buf[d == 0 ? 1 : 3]) continue;
            if (wait_by_memop())
                memop_check(c, wait_value32()((CUstream)st, (CUdeviceptr)(c->flags + 2 * d + slot), 1,
                                              CU_STREAM_WAIT_VALUE_GEQ), "cuStreamWaitValue32");
            else
                k_wait_flag<<<1, 1, 0, st>>>(c->flags + 2 * d + slot, 1u);
        }
        h->launches++;
        return c->error ? 1 : 0;
    }
    dc::NcclApi *a = dc::nccl_api();
    int e = a->GroupStart();
    if (!e && c->rank > 0) {
        e = a->Send(c->buf[0], c->nelem, dc::NCCL_FLOAT64, c->rank - 1, c->comm, st);
        if (!e) e = a->Recv(c->buf[1], c->nelem, dc::NCCL_FLOAT64, c->rank - 1, c->comm, st);
    }
    if (!e && c->rank < c->nranks - 1) {
        e = a->Send(c->buf[2], c->nelem, dc::NCCL_FLOAT64, c->rank + 1, c->comm, st);
        if (!e) e = a->Recv(c->buf[3], c->nelem, dc::NCCL_FLOAT64, c->rank + 1, c->comm, st);
    }
    const int e2 = a->GroupEnd();
    if (e || e2) {
        c->error = e ? e : e2;
        return nccl_fail("ncclSend/ncclRecv", c->error);
    }
    h->launches++;
    return 0;
}
// the receive slot of `stage` has been unpacked: hand it back (the neighbour writes it again two
// stages later, after it has received this rank's next message -- see include/dyncore.h)
static void dcb_comm_consumed(dc_handle *h, int stage, void *stream)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    if (!c->p2p) return;
    const int slot = stage & 1;
    for (int d = 0; d < 2; d++)
        if (c->buf[d == 0 ? 1 : 3])
            memop_check(c, write_value32()((CUstream)(cudaStream_t)stream,
                                           (CUdeviceptr)(c->flags + 2 * d + slot), 0, 0),
                        "cuStreamWriteValue32 (own flag)");
}
static void *dcb_side_stream(dc_handle *h, int which)
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    return which ? c->side2 : c->side;
}
static void dcb_event_record(dc_handle *h, int ev, void *stream)
{
    cudaEventRecord(static_cast<dc::CommState *>(h->comm_state)->ev[ev], (cudaStream_t)stream);
}
static void dcb_stream_wait(dc_handle *h, int ev, void *stream)
{
    cudaStreamWaitEvent((cudaStream_t)stream, static_cast<dc::CommState *>(h->comm_state)->ev[ev], 0);
}
// The banded step captured into CUDA graphs (stream capture of the enqueue functions, which fork
// to the side stream and join again through events) and replayed: one cudaGraphLaunch per step
// instead of ~16 kernel launches, ~20 event operations and 2 NCCL groups.  Two graphs: a step
// that ends with the continuity of the next step, and the last step of a call.
static int capture_graph(dc_handle *h, dc::CommState *c, cudaStream_t st,
                         void (*enqueue)(dc_handle *, void *), cudaGraphExec_t *out,
                         long long *launches)
{
    cudaGraph_t g = nullptr;
    const long long launches0 = h->launches;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    enqueue(h, st);
    const cudaError_t e = cudaStreamEndCapture(st, &g);
    *launches = h->launches - launches0;
    h->launches = launches0;
    if (e != cudaSuccess || !g) {
        cudaGetLastError();
        return 1;
    }
    if (cudaGraphInstantiate(out, g, 0) != cudaSuccess) {
        cudaGraphDestroy(g);
        cudaGetLastError();
        *out = nullptr;
        return 1;
    }
    cudaGraphDestroy(g);
    (void)c;
    return 0;
}
static int dcb_graph_steps(dc_handle *h, int nsteps, void *stream,
                           void (*prologue)(dc_handle *, int, void *),
                           void (*step_tail)(dc_handle *, void *),
                           void (*step_last)(dc_handle *, void *))
{
    dc::CommState *c = static_cast<dc::CommState *>(h->comm_state);
    cudaStream_t caller = (cudaStream_t)stream;
    if (!c->warmed) {   // lazy NCCL connection set-up and cudaFuncSetAttribute must not be captured
        c->warmed = true;
        return 1;
    }
    // the legacy default stream (torch's default current stream) cannot be captured: run on the
    // handle's own stream between two events on the caller's
    const bool legacy = caller == nullptr || caller == cudaStreamLegacy ||
                        caller == cudaStreamPerThread;
    cudaStream_t st = legacy ? c->main : caller;
    if (!c->graph || !c->graph_last || c->graph_version != h->bind_version) {
        if (c->graph) cudaGraphExecDestroy(c->graph);
        if (c->graph_last) cudaGraphExecDestroy(c->graph_last);
        c->graph = c->graph_last = nullptr;
        if (capture_graph(h, c, st, step_tail, &c->graph, &c->launches_per_step) ||
            capture_graph(h, c, st, step_last, &c->graph_last, &c->launches_last_step))
            return 1;
        c->graph_version = h->bind_version;
        c->error = 0;
    }
    if (legacy) {
        cudaEventRecord(c->ev[14], caller);
        cudaStreamWaitEvent(c->main, c->ev[14], 0);
    }
    const long long l0 = h->launches;
    prologue(h, 0, st);
    (void)l0;
    int rc = 0;
    for (int s = 0; s < nsteps && !rc; s++) {
        const bool last = s + 1 == nsteps;
        if (cudaGraphLaunch(last ? c->graph_last : c->graph, st) != cudaSuccess) rc = 1;
        h->launches += last ? c->launches_last_step : c->launches_per_step;
    }
    if (legacy) {
        cudaEventRecord(c->ev[15], c->main);
        cudaStreamWaitEvent(caller, c->ev[15], 0);
    }
    return rc ? 2 : 0;   // 2: a launch failed after steps were enqueued (reported by the caller)
}

static dc::ProfileState *pstate(dc_handle *h)
{
    if (!h->profile_state) h->profile_state = new dc::ProfileState();
    return static_cast<dc::ProfileState *>(h->profile_state);
}
static void dcb_profile_begin(dc_handle *h, const char *name, void *stream)
{
    dc::ProfileState *p = pstate(h);
    dc::ProfileRec r{name, p->get(), p->get()};
    cudaEventRecord(r.a, (cudaStream_t)stream);
    p->recs.push_back(r);
}
static void dcb_profile_end(dc_handle *h, void *stream)
{
    cudaEventRecord(pstate(h)->recs.back().b, (cudaStream_t)stream);
}
static void dcb_mark(dc_handle *h, const char *name, void *stream)
{
    dc::ProfileState *p = pstate(h);
    cudaEvent_t e = p->get();
    cudaEventRecord(e, (cudaStream_t)stream);
    p->marks.push_back({name, e});
}
static int dcb_profile_read(dc_handle *h, int max, const char **names, double *ms, long long *n)
{
    dc::ProfileState *p = pstate(h);
    cudaDeviceSynchronize();
    if (!p->marks.empty()) {
        // timeline mode: one entry per mark, in enqueue order; ms = time since the first mark,
        // launches = the index of the mark
        p->names.clear();
        for (auto &m : p->marks) p->names.push_back(m.first);
        int i = 0;
        for (auto &m : p->marks) {
            float t = 0.f;
            cudaEventElapsedTime(&t, p->marks[0].second, m.second);
            if (i < max) {
                names[i] = p->names[i].c_str();
                ms[i] = t;
                n[i] = i;
                i++;
            }
            p->pool.push_back(m.second);
        }
        p->marks.clear();
        for (const dc::ProfileRec &r : p->recs) {
            p->pool.push_back(r.a);
            p->pool.push_back(r.b);
        }
        p->recs.clear();
        return i;
    }
    std::map<std::string, std::pair<double, long long>> acc;
    for (const dc::ProfileRec &r : p->recs) {
        float t = 0.f;
        cudaEventElapsedTime(&t, r.a, r.b);
        auto &e = acc[r.name];
        e.first += t;
        e.second += 1;
        p->pool.push_back(r.a);
        p->pool.push_back(r.b);
    }
    p->recs.clear();
    p->names.clear();
    for (auto &kv : acc) p->names.push_back(kv.first);
    int i = 0;
    for (auto &kv : acc) {
        if (i >= max) break;
        names[i] = p->names[i].c_str();
        ms[i] = kv.second.first;
        n[i] = kv.second.second;
        i++;
    }
    return i;
}

// ---------------------------------------------------------------------------------------
// third-generation stage kernel: descriptor cache and launch
// ---------------------------------------------------------------------------------------
static const CUtensorMap *stage3_map(dc_handle *h, const double *base, int nk, bool own)
{
    using namespace dc;
    if (!h->tma_state) h->tma_state = new TmaState();
    TmaState *st = static_cast<TmaState *>(h->tma_state);
    const auto key = std::make_pair(static_cast<const void *>(base), (own ? 1 : 0) | (nk << 1));
    auto it = st->maps.find(key);
    if (it != st->maps.end()) return &it->second;
    EncodeTiledFn enc = encode_tiled();
    const Geom &g = h->g;
    CUtensorMap m;
    const cuuint64_t dims[3] = {(cuuint64_t)g.NI, (cuuint64_t)g.NJ, (cuuint64_t)nk};
    const cuuint64_t strides[2] = {(cuuint64_t)g.NI * 8, (cuuint64_t)g.plane * 8};
    const cuuint32_t box[3] = {(cuuint32_t)(own ? S3_OW : S3_SW), (cuuint32_t)(own ? S3_TY : S3_SH),
                               1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (!enc || enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double *>(base), dims,
                    strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) {
        g_stage3_error = (int)cudaErrorInvalidValue;
        return nullptr;
    }
    return &(st->maps[key] = m);
}
static void dcb_tma_release(dc_handle *h)
{
    delete static_cast<dc::TmaState *>(h->tma_state);
    h->tma_state = nullptr;
}
static void dcb_launch_stage3(dc_handle *h, dc::Stage3Body &b, const dc::Stage3Ptrs &p, int nbx,
                              int nby, void *stream)
{
    using namespace dc;
    static unsigned long long done = 0;
    const bool configured = !first_use_on_device(&done);
    const int smem = (int)sizeof(Stage3Smem) + 128;
    if (!configured) {
        cudaFuncSetAttribute(k_stage3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_stage3, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    const int nz = h->g.nz;
    struct { CUtensorMap *dst; const double *base; int nk; bool own; } want[] = {
        {&b.mU, p.U, nz, false},     {&b.mV, p.V, nz, false},   {&b.mW, p.W, nz + 1, false},
        {&b.mPHI, p.PHI, nz, false}, {&b.mT, p.T, nz, false},   {&b.mG, p.G, nz, false},
        {&b.mTB, p.TB, nz + 1, true}, {&b.mUo, p.Uo, nz, true},
        {&b.mVo, p.Vo, nz, true},    {&b.mTo, p.To, nz, true}};
    for (auto &w : want) {
        const CUtensorMap *m = stage3_map(h, w.base, w.nk, w.own);
        if (!m) return;
        *w.dst = *m;
    }
    k_stage3<<<dim3(nbx, nby, b.nkc), dim3(S3_NT), smem, (cudaStream_t)stream>>>(b);
}

static void dcb_launch_moist3(dc_handle *h, dc::Moist3Body &b, const dc::Moist3Ptrs &p, int nbx,
                              int nby, void *stream)
{
    using namespace dc;
    static unsigned long long done = 0;
    const bool configured = !first_use_on_device(&done);
    const int smem = (int)sizeof(Moist3Smem) + 128;
    if (!configured) {
        cudaFuncSetAttribute(k_moist3, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(k_moist3, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    }
    const int nz = h->g.nz;
    struct { CUtensorMap *dst; const double *base; int nk; bool own; } want[] = {
        {&b.mU, p.U, nz, false},      {&b.mV, p.V, nz, false},
        {&b.mQ[0], p.Q[0], nz, false}, {&b.mQ[1], p.Q[1], nz, false},
        {&b.mW, p.W, nz + 1, true},   {&b.mQo[0], p.Qo[0], nz, true},
        {&b.mQo[1], p.Qo[1], nz, true}};
    for (auto &w : want) {
        const CUtensorMap *m = stage3_map(h, w.base, w.nk, w.own);
        if (!m) return;
        *w.dst = *m;
    }
    k_moist3<<<dim3(nbx, nby, b.nkc), dim3(S3_NT), smem, (cudaStream_t)stream>>>(b);
}
