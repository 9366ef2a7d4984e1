// emu_dyncore.cpp -- TEST INFRASTRUCTURE ONLY.
// Host emulation of libdyncore.so: the SAME C ABI implementation (dc_api_impl.h) and the
// SAME kernel bodies (dc_kernels.h) as the CUDA library, with "device" memory = host
// memory and a kernel launch = a loop over the thread box.  It lets the CPU test-suite
// (-m "not gpu") check the index arithmetic, boundary images and host logic of the product
// without a GPU.  It is never loaded by the product package: climate_model_b200/_lib.py only
// loads the CUDA library and fails loudly when that is missing.
#include <stdlib.h>
#include <string.h>

#define DC_BACKEND_IS_CUDA 0

static int dcb_malloc(void **p, size_t n) { *p = malloc(n); return *p ? 0 : 2; }
static int dcb_free(void *p) { free(p); return 0; }
static int dcb_h2d(void *dst, const void *src, size_t n) { memcpy(dst, src, n); return 0; }
static int dcb_d2d_async(void *dst, const void *src, size_t n, void *) { memcpy(dst, src, n); return 0; }
static int dcb_last_error() { return 0; }
static const char *dcb_error_string(int) { return "host emulation"; }

template <class Body>
static void dcb_launch(const Body &b, int i0, int i1, int j0, int j1, void *)
{
    // thread order is irrelevant for a race-free kernel; run rows in DESCENDING order so
    // that a body that wrongly depended on its neighbours' output would not get the
    // "natural" sequential answer by accident
    for (int j = j1; j >= j0; j--)
        for (int i = i1; i >= i0; i--) b(i, j);
}

template <class Body, class Smem>
static void dcb_launch_blocks(const Body &b, int nbx, int nby, int, void *)
{
    static Smem s;
    for (int by = nby - 1; by >= 0; by--)
        for (int bx = nbx - 1; bx >= 0; bx--) {
            for (size_t n = 0; n < sizeof(s) / sizeof(double); n++)
                reinterpret_cast<double *>(&s)[n] = 0.0 / 0.0;   // stale smem must not be read
            b.run_block(bx, by, s);
        }
}

#include "../../climate_model_b200/csrc/dc_stage3.h"
#include "../../climate_model_b200/csrc/dc_moist3.h"
struct dc_handle;
namespace dc { struct Stage3Ptrs; }
static void dcb_launch_stage3(dc_handle *h, dc::Stage3Body &b, const dc::Stage3Ptrs &p, int nbx,
                              int nby, void *);
template <class Body>
static void dcb_launch_diag(const Body &b, int i0, int i1, int j0, int j1, void *stream)
{
    dcb_launch(b, i0, i1, j0, j1, stream);
}
static void dcb_tma_release(dc_handle *) {}
namespace dc { struct Moist3Ptrs; }
static void dcb_launch_moist3(dc_handle *h, dc::Moist3Body &b, const dc::Moist3Ptrs &p, int nbx,
                              int nby, void *);
static void dcb_transpose(const dc::Geom &g, double *ref, double *dev, int fnx, int fny, int nk,
                          int j_lo, int j_hi, int to_device, void *)
{
    for (int i = 0; i < fnx; i++)
        for (int j = j_lo; j <= j_hi; j++)
            for (int k = 0; k < nk; k++) {
                double &r = ref[((size_t)i * fny + j) * nk + k];
                double &d = dev[g.idx(i, j, k)];
                if (to_device) d = r; else r = d;
            }
}
struct dc_handle;
static void dcb_profile_begin(dc_handle *, const char *, void *) {}
static void dcb_profile_end(dc_handle *, void *) {}
static void dcb_mark(dc_handle *, const char *, void *) {}
static int dcb_profile_read(dc_handle *, int, const char **, double *, long long *) { return 0; }

// the host emulation has no in-library communicator: the CPU tests drive the exchange through
// the piecewise band entries over gloo (parallel_bands.py)
#include "../../include/dyncore.h"
static const char *dcb_comm_error() { return ""; }
static int dcb_comm_unique_id(void *) { return DC_ERR_NO_DEVICE; }
static int dcb_comm_init(dc_handle *, const void *, int, int, size_t) { return DC_ERR_NO_DEVICE; }
static void dcb_comm_release(dc_handle *) {}
static double *dcb_comm_buffer(dc_handle *, int, int) { return nullptr; }
static int dcb_comm_sendrecv(dc_handle *, int, void *) { return DC_ERR_NO_DEVICE; }
static void dcb_comm_consumed(dc_handle *, int, void *) {}
static int dcb_comm_p2p_handles(dc_handle *, void *) { return DC_ERR_NO_DEVICE; }
static int dcb_comm_p2p_connect(dc_handle *, const void *, const void *) { return DC_ERR_NO_DEVICE; }
static int dcb_comm_p2p_enable(dc_handle *, int) { return DC_ERR_NO_DEVICE; }
static void *dcb_side_stream(dc_handle *, int = 0) { return nullptr; }
static void dcb_event_record(dc_handle *, int, void *) {}
static void dcb_stream_wait(dc_handle *, int, void *) {}
static int dcb_graph_steps(dc_handle *, int, void *, void (*)(dc_handle *, int, void *),
                           void (*)(dc_handle *, void *), void (*)(dc_handle *, void *))
{
    return 1;
}

#include "../../climate_model_b200/csrc/dc_api_impl.h"

// third-generation stage kernel: a TMA descriptor = (base, dims, box); a block = one call
static void dcb_launch_stage3(dc_handle *, dc::Stage3Body &b, const dc::Stage3Ptrs &p, int nbx,
                              int nby, void *)
{
    using namespace dc;
    const Geom &g = b.g;
    auto mk = [&](TmaMap &m, const double *base, int nk, bool own) {
        m.base = base;
        m.dim[0] = g.NI; m.dim[1] = g.NJ; m.dim[2] = nk;
        m.box[0] = own ? S3_OW : S3_SW; m.box[1] = own ? S3_TY : S3_SH; m.box[2] = 1;
    };
    mk(b.mU, p.U, g.nz, false); mk(b.mV, p.V, g.nz, false); mk(b.mW, p.W, g.nz + 1, false);
    mk(b.mPHI, p.PHI, g.nz, false); mk(b.mT, p.T, g.nz, false); mk(b.mG, p.G, g.nz, false);
    mk(b.mTB, p.TB, g.nz + 1, true); mk(b.mUo, p.Uo, g.nz, true); mk(b.mVo, p.Vo, g.nz, true);
    mk(b.mTo, p.To, g.nz, true);
    static Stage3Smem s;   // one "block" at a time
    for (int bz = 0; bz < b.nkc; bz++)
        for (int by = nby - 1; by >= 0; by--)
            for (int bx = nbx - 1; bx >= 0; bx--) {
                for (size_t n = 0; n < sizeof(s) / sizeof(double); n++)
                    reinterpret_cast<double *>(&s)[n] = 0.0 / 0.0;   // stale smem must not be read
                b.run_block(bx, by, bz, s);
            }
}

static void dcb_launch_moist3(dc_handle *, dc::Moist3Body &b, const dc::Moist3Ptrs &p, int nbx,
                              int nby, void *)
{
    using namespace dc;
    const Geom &g = b.g;
    auto mk = [&](TmaMap &m, const double *base, int nk, bool own) {
        m.base = base;
        m.dim[0] = g.NI; m.dim[1] = g.NJ; m.dim[2] = nk;
        m.box[0] = own ? S3_OW : S3_SW; m.box[1] = own ? S3_TY : S3_SH; m.box[2] = 1;
    };
    mk(b.mU, p.U, g.nz, false); mk(b.mV, p.V, g.nz, false);
    mk(b.mQ[0], p.Q[0], g.nz, false); mk(b.mQ[1], p.Q[1], g.nz, false);
    mk(b.mW, p.W, g.nz + 1, true); mk(b.mQo[0], p.Qo[0], g.nz, true);
    mk(b.mQo[1], p.Qo[1], g.nz, true);
    static Moist3Smem s;
    for (int bz = 0; bz < b.nkc; bz++)
        for (int by = nby - 1; by >= 0; by--)
            for (int bx = nbx - 1; bx >= 0; bx--) {
                for (size_t n = 0; n < sizeof(s) / sizeof(double); n++)
                    reinterpret_cast<double *>(&s)[n] = 0.0 / 0.0;
                b.run_block(bx, by, bz, s);
            }
}

// ---- test hooks (host emulation only): the table-driven power / logarithm of the production
//      build, evaluated on the host with the same code the kernels inline ----
extern "C" double emu_pow_kappa_tab(double x)
{
    static std::vector<double> tab;
    static dc::PowCoef c;
    if (tab.empty()) {
        tab.resize(2 * dc::POW_NE * dc::POW_NJ);
        dc::make_pow_table(dc::con_kappa, tab.data());
        c = dc::make_pow_coef(dc::con_kappa, tab.data());
    }
    return dc::pow_kappa_tab(x, c);
}
extern "C" double emu_log_tab(double x)
{
    static const dc::LogCoef L = dc::make_log_coef();
    return dc::log_tab(x, L, L.tab);
}
