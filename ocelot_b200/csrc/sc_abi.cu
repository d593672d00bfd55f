// sc_abi.cu -- handle management and the extern "C" entry points of include/ocelot_sc.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "../../include/ocelot_sc.h"
#include "sc_kernels.h"

// NVTX range per entry point (tracing).  Header-only NVTX v3: without a profiler attached a
// push/pop pair is two calls through a no-op function pointer.  ncu / nsys show the kick as
// ocl_sc_kick_device > momentum / extent / deposit / rho reduce / solve / kick; under graph replay the stage ranges
// appear once, around the capture.
namespace {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
}  // namespace
#define OCL_RANGE(name) NvtxRange nvtx_range_(name)

using namespace ocl;

namespace {

thread_local std::string g_create_error;

enum TimerSlot { T_BEGIN = 0, T_MOM, T_EXT, T_DEP, T_SOLVE, T_FIELD, T_KICK, T_COUNT };

}  // namespace

struct ocl_sc {
    int device = 0;
    MeshDims md{};
    ReduceState rs{};
    // constants (ocelot/common/globals.py:13-24)
    double m_e_eV = 0, m_e_GeV = 0;
    // grids
    double* rho = nullptr;    // n^3
    double* gtab = nullptr;   // (n+1)^3 antiderivative table
    double* k1 = nullptr;     // n^3 (tap only, lazily allocated)
    // solver 0 (default): hand-written pruned/symmetric Hockney convolution (sc_fft.cu)
    // solver 1: cuFFT D2Z/Z2D on the full padded box (cross-check; also used when M > 1024)
    int solver = 0;
    FftWork fw{};
    double2* tw[3] = {nullptr, nullptr, nullptr};
    double* h3 = nullptr;                     // mesh steps of the current kick
    double* moments = nullptr;                // 18 doubles: beam moments (ocl_sc_beam_moments)
    int* cut_counts = nullptr;                // aperture compaction: per-tile survivor counts / offsets
    long long cut_tiles = 0;
    long long* cut_n = nullptr;               // mapped host word: survivors of the last cut
    double* real_buf = nullptr;               // M^3 real: K, then padded rho, then the convolution
    cufftDoubleComplex* k_hat = nullptr;      // M*M*(M/2+1)
    cufftDoubleComplex* rho_hat = nullptr;    // M*M*(M/2+1)
    double* phi = nullptr;    // n^3
    EQuad* equad = nullptr;   // 3*n^3 (+1 pad) quads: the field table the gather reads
    // 0: z-fastest table, independent gathers | 1: x-fastest table, lane-pair gathers, rows staged through shared
    // memory (cp.async) | 2: same table, rows prefetched into registers.  Default (-1 -> resolved at create): 2 while
    // the table stays near the L2 (the gather is then bound by the L1TEX data stage, which the staging also loads:
    // 12.5 M / 127^3 368 -> 361 us, 1 M / 63^3 37.2 -> 34.8 us on B200), 1 for larger tables (DRAM-bound: the
    // deeper shared-memory pipeline wins, 50 M / 255^3 2.70 vs 2.86 ms).  OCL_SC_GATHER overrides.
    int layout = -1;
    cufftHandle plan_fwd = 0, plan_inv = 0;
    bool plans = false;
    // slab mode (multi-GPU solve)
    int slab_world = 0, slab_rank = 0, sx = 0, fs = 0, nx_pad = 0;
    size_t rho_count = 0;                     // doubles in rho / phi (n^3, or nx_pad*ny*nz in slab mode)
    double* rho_slab = nullptr;
    double* phi_slab = nullptr;
    double2* xchg_a = nullptr;
    double2* xchg_b = nullptr;
    double2* own_xchg_a = nullptr;            // the cudaMalloc'ed buffers (kept for freeing when xchg_* are symmetric)
    double2* own_xchg_b = nullptr;
    PeerXchg peer_xchg{};                     // world > 0: the passes store straight into the peers' exchange buffers
    // peer-memory mailbox (multi-GPU scalar exchanges)
    Mailbox mb{};
    int* mb_err = nullptr;
    PeerRho peer_rho{};                       // world > 0: rho lives in caller-owned symmetric memory
    double* own_rho = nullptr;                // the cudaMalloc'ed grid (kept for freeing)
    double* mc_rho = nullptr;                 // multicast mapping of every rank's rho (NVLS reduction)
    double* mc_phi = nullptr;                 // slab mode: multicast mapping of every rank's phi (the inverse z pass
    double* own_phi = nullptr;                // broadcasts its slab through the switch); own_phi: the cudaMalloc'ed grid
    // ordered deposit (ocl_sc_set_deterministic / OCL_SC_DETERMINISTIC=1): np.bincount's summation order, bit-identical
    // rho from run to run; runs the stages directly (no kick graph), scratch sized for the largest bunch seen
    bool ordered = false;
    void* ordered_buf = nullptr;
    size_t ordered_bytes = 0;
    // ... and the momentum sum follows numpy's pairwise tree over exactly rounded momenta (plan per bunch size)
    long long pw_n = -1;
    int pw_leaves = 0, pw_levels = 0;
    uint2* pw_leaf_dev = nullptr;
    uint2* pw_node_dev = nullptr;
    int* pw_level_dev = nullptr;
    double* pw_values = nullptr;
    int debug_skip = 0;                       // timing experiments only (OCL_SC_DEBUG_SKIP): bit 0/1/2 = leave out the
                                              // momentum exchange / extent exchange / rho reduction of a sharded kick
    // host arrays page-locked in place on first use (numpy buffers persist across kicks)
    bool pin_host = false;
    void* pinned[2] = {nullptr, nullptr};
    size_t pinned_bytes[2] = {0, 0};
    // host-mode staging
    double* stage_r = nullptr;
    double* stage_q = nullptr;
    long long stage_cap = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t side_stream = nullptr;       // K_hat chain runs here, concurrent with the deposit
    cudaEvent_t ev_fork = nullptr, ev_khat = nullptr, ev_handoff = nullptr;
    bool last_stream_valid = false;
    bool khat_pending = false;
    cudaStream_t last_stream = nullptr;
    // whole-kick CUDA graph (single-GPU ocl_sc_kick_device): captured once per (r, ld, q, n),
    // per kick only the parameter node (E, dz, mesh draws) is refreshed
    KickParams* kp_dev = nullptr;
    bool use_graph = true;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraphNode_t param_node = nullptr;
    const void* g_r = nullptr; const void* g_q = nullptr; long long g_ld = 0, g_n = 0;
    long long graph_launches_per_kick = 0;
    const KickParams* cur_pp = nullptr;       // non-null while stages are being captured
    bool publish_params = false;              // capture of the library's own graph: k_momentum is the parameter node
    cudaKernelNodeParams param_np = {};       // its launch shape, as captured
    Mailbox g_mb{};                           // arguments the parameter node was captured with
    ReduceState g_rs{};
    // longitudinal space charge (sc_lsc.cu): 1-D work space, grown on demand
    LscWork lw{};
    int lsc_cap = 0;                          // grid points the buffers hold
    long long lsc_spread_cap = 0;             // 64-bit words of the spread histogram
    int lsc_tw_nb = 0;                        // nb the twiddle table was built for
    int lsc_nb = 0;                           // nb of the last deposit / solve
    int* lsc_err_host = nullptr;              // mapped host flag the asynchronous form raises (grid too large ...)
    bool lsc_async_pending = false;           // last kick was asynchronous: nb lives on the device
    // timers
    bool timers = false;
    cudaEvent_t ev[T_COUNT] = {};
    bool ev_valid = false;
    long long launches = 0;
    std::string err;
};

namespace {

int fail(ocl_sc* h, const char* what, const char* detail) {
    std::string m = std::string(what) + ": " + detail;
    if (h) h->err = m; else g_create_error = m;
    return 1;
}

#define CU(h, call)                                                              \
    do {                                                                         \
        cudaError_t e_ = (call);                                                 \
        if (e_ != cudaSuccess) return fail((h), #call, cudaGetErrorString(e_));  \
    } while (0)

#define FFT(h, call)                                                             \
    do {                                                                         \
        cufftResult r_ = (call);                                                 \
        if (r_ != CUFFT_SUCCESS) {                                               \
            char b_[64];                                                         \
            snprintf(b_, sizeof b_, "cufft error %d", (int)r_);                  \
            return fail((h), #call, b_);                                         \
        }                                                                        \
    } while (0)

void constants(double& m_e_eV, double& m_e_GeV, double& eps0, double& pi, double& c) {
    pi = 3.141592653589793;
    c = 299792458.0;
    const double q_e = 1.6021766208e-19, m_e_kg = 9.10938215e-31;
    m_e_eV = m_e_kg * (c * c) / q_e;
    m_e_GeV = m_e_eV / 1e+9;
    const double mu0 = 4 * pi * 1e-7;
    eps0 = 1 / mu0 / (c * c);
}

RefParams ref_params(const ocl_sc* h, double E_GeV) {
    RefParams rp;
    rp.m_e_eV = h->m_e_eV;
    rp.inv_m2 = 1.0 / (h->m_e_eV * h->m_e_eV);
    rp.gamref = E_GeV / h->m_e_GeV;                                  // sc.py:214
    rp.betaref = std::sqrt(1 - std::pow(rp.gamref, -2.0));           // sc.py:215-216
    rp.inv_betaref = 1.0 / rp.betaref;
    rp.inv_gamref = 1.0 / rp.gamref;
    rp.gb_ref = rp.gamref * rp.betaref;
    rp.inv_gb2 = 1.0 / (rp.gb_ref * rp.gb_ref);
    rp.pc = rp.gb_ref * h->m_e_eV;
    rp.inv_pref = 1.0 / (h->m_e_eV * std::sqrt(rp.gamref * rp.gamref - 1));   // coord_transform.py:19
    return rp;
}

Draws draws_of(const double* mesh_draws) {
    Draws d;
    if (mesh_draws) { d.scale = mesh_draws[0]; d.shift = mesh_draws[1]; }
    else { d.scale = 0.0; d.shift = 0.0; }
    return d;
}

KickParams kick_params(const ocl_sc* h, double E_GeV, double dz, const double* mesh_draws) {
    KickParams k;
    k.rp = ref_params(h, E_GeV);
    k.dr = draws_of(mesh_draws);
    k.cdT = dz / k.rp.betaref;                                       // sc.py:244
    return k;
}

// by value for direct launches, through the device block while a graph is captured / replayed
KP kp_of(const ocl_sc* h, double E_GeV, double dz, const double* mesh_draws) {
    KP k;
    k.v = kick_params(h, E_GeV, dz, mesh_draws);
    k.p = h->cur_pp;
    return k;
}

int check_launch(ocl_sc* h, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(h, what, cudaGetErrorString(e));
    return 0;
}

void mark(ocl_sc* h, int slot, cudaStream_t st) {
    if (h->timers) { cudaEventRecord(h->ev[slot], st); if (slot == T_KICK) h->ev_valid = true; }
}

// Every entry point runs on the handle's device and leaves the calling thread's current device as it
// found it (callers such as torch keep their own notion of the current device).
struct DeviceScope {
    int prev = -1;
    bool switched = false, bad = false;
    explicit DeviceScope(ocl_sc* h) : DeviceScope(h, h->device) {}
    DeviceScope(ocl_sc* h, int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) {
            cudaError_t e = cudaSetDevice(device);
            if (e != cudaSuccess) { fail(h, "cudaSetDevice", cudaGetErrorString(e)); bad = true; }
            else switched = true;
        }
    }
    ~DeviceScope() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define ENTER_DEVICE(h) DeviceScope device_scope_(h); if (device_scope_.bad) return 1

// The handle's scratch (reduction buffers, grids, FFT work space) is shared by consecutive kicks.
// When a call arrives on a different stream than the previous one, order it after the work
// already queued there, so callers may alternate streams (e.g. a device-resident kick on the
// caller's stream followed by a host-array kick on the handle's own stream).
int adopt_stream(ocl_sc* h, cudaStream_t st) {
    if (h->last_stream_valid && h->last_stream != st) {
        cudaStreamCaptureStatus a = cudaStreamCaptureStatusNone, b = cudaStreamCaptureStatusNone;
        cudaStreamIsCapturing(h->last_stream, &a);
        cudaStreamIsCapturing(st, &b);
        cudaGetLastError();
        if (a == cudaStreamCaptureStatusNone && b == cudaStreamCaptureStatusNone) {
            CU(h, cudaEventRecord(h->ev_handoff, h->last_stream));
            CU(h, cudaStreamWaitEvent(st, h->ev_handoff, 0));
        }
    }
    h->last_stream = st;
    h->last_stream_valid = true;
    return 0;
}

int ensure_plans(ocl_sc* h) {
    if (h->plans) return 0;
    const size_t m3 = (size_t)h->md.mx * h->md.my * h->md.mz;
    const size_t c3 = (size_t)h->md.mx * h->md.my * (h->md.mz / 2 + 1);
    CU(h, cudaMalloc(&h->real_buf, sizeof(double) * m3));
    CU(h, cudaMalloc(&h->k_hat, sizeof(cufftDoubleComplex) * c3));
    CU(h, cudaMalloc(&h->rho_hat, sizeof(cufftDoubleComplex) * c3));
    FFT(h, cufftPlan3d(&h->plan_fwd, h->md.mx, h->md.my, h->md.mz, CUFFT_D2Z));
    FFT(h, cufftPlan3d(&h->plan_inv, h->md.mx, h->md.my, h->md.mz, CUFFT_Z2D));
    h->plans = true;
    return 0;
}

// exp(-2 pi i m / M), m = 0..M-1, rounded from long double
int make_twiddles(ocl_sc* h, int M, double2** out) {
    std::vector<double2> t((size_t)M);
    const long double two_pi = 6.283185307179586476925286766559L;
    for (int m = 0; m < M; ++m) {
        long double a = two_pi * (long double)m / (long double)M;
        t[m].x = (double)cosl(a);
        t[m].y = (double)(-sinl(a));
    }
    // exact values on the axes
    t[0] = make_double2(1.0, 0.0);
    if (M % 2 == 0) t[M / 2] = make_double2(-1.0, 0.0);
    if (M % 4 == 0) { t[M / 4] = make_double2(0.0, -1.0); t[3 * M / 4] = make_double2(0.0, 1.0); }
    CU(h, cudaMalloc(out, sizeof(double2) * M));
    CU(h, cudaMemcpy(*out, t.data(), sizeof(double2) * M, cudaMemcpyHostToDevice));
    return 0;
}

// hand-written path: K_hat (3 real-even passes), rho passes with the multiply fused in x; writes phi.
// If the K_hat chain was forked onto the side stream (khat_pending), join it before the x pass.
int solve_fused(ocl_sc* h, cudaStream_t st) {
    if (!h->khat_pending) launch_khat(h->gtab, h->md, h->fw, st);
    launch_convolve_pre(h->rho, h->peer_rho, h->md, h->fw, st);
    if (h->khat_pending) {
        CU(h, cudaStreamWaitEvent(st, h->ev_khat, 0));
        h->khat_pending = false;
    }
    launch_convolve_post(h->md, h->fw, h->h3, four_pi_eps0_value(), h->phi, st);
    h->launches += 8;
    return check_launch(h, "solve_fused");
}

// Green's function table + K_hat on the side stream, ordered after everything already in st
int fork_khat(ocl_sc* h, cudaStream_t st) {
    CU(h, cudaEventRecord(h->ev_fork, st));
    CU(h, cudaStreamWaitEvent(h->side_stream, h->ev_fork, 0));
    launch_green_table(h->rs, h->md, h->gtab, h->h3, h->side_stream);
    launch_khat(h->gtab, h->md, h->fw, h->side_stream);
    CU(h, cudaEventRecord(h->ev_khat, h->side_stream));
    h->khat_pending = true;
    h->launches += 1;
    return check_launch(h, "fork_khat");
}

// IGF -> K_hat, rho -> rho_hat, multiply, inverse: real_buf holds the convolution afterwards
int convolve(ocl_sc* h, cudaStream_t st) {
    if (ensure_plans(h)) return 1;
    FFT(h, cufftSetStream(h->plan_fwd, st));
    FFT(h, cufftSetStream(h->plan_inv, st));
    launch_green_mirror(h->gtab, h->md, h->real_buf, st);
    FFT(h, cufftExecD2Z(h->plan_fwd, h->real_buf, h->k_hat));
    launch_pad_rho(h->rho, h->md, h->real_buf, st);
    FFT(h, cufftExecD2Z(h->plan_fwd, h->real_buf, h->rho_hat));
    launch_multiply(h->rho_hat, h->k_hat, h->md, st);
    FFT(h, cufftExecZ2D(h->plan_inv, h->rho_hat, h->real_buf));
    h->launches += 7;
    return check_launch(h, "convolve");
}

}  // namespace

extern "C" {

static void drop_graph(ocl_sc* h);

int ocl_sc_abi_version(void) { return 1; }

void ocl_sc_get_constants(double out[5]) {
    constants(out[0], out[1], out[2], out[3], out[4]);
}

int ocl_sc_fft_size(int n) {
    int m = 1;
    while (m < 2 * n - 1) m *= 2;
    return m;
}

const char* ocl_sc_last_error(const ocl_sc_t* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int ocl_sc_create(int device, int nx, int ny, int nz, long long max_particles, ocl_sc_t** out) {
    if (!out) return fail(nullptr, "ocl_sc_create", "out is NULL");
    *out = nullptr;
    if (nx < 4 || ny < 4 || nz < 4) return fail(nullptr, "ocl_sc_create", "mesh needs at least 4 points per axis");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, "ocl_sc_create", "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(nullptr, "ocl_sc_create", "bad device index");
    ocl_sc* h = new ocl_sc();
    h->device = device;
    double eps0, pi, c;
    constants(h->m_e_eV, h->m_e_GeV, eps0, pi, c);
    h->md.nx = nx; h->md.ny = ny; h->md.nz = nz;
    h->md.mx = ocl_sc_fft_size(nx); h->md.my = ocl_sc_fft_size(ny); h->md.mz = ocl_sc_fft_size(nz);
    const size_t n3 = (size_t)nx * ny * nz;
    const size_t g3 = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
    const size_t m3 = (size_t)h->md.mx * h->md.my * h->md.mz;
    const size_t c3 = (size_t)h->md.mx * h->md.my * (h->md.mz / 2 + 1);
    h->rs.max_blocks = 148 * 4;
#define TRY(call)                                                             \
    do {                                                                      \
        cudaError_t e2_ = (call);                                             \
        if (e2_ != cudaSuccess) {                                             \
            fail(nullptr, #call, cudaGetErrorString(e2_));                    \
            ocl_sc_destroy(h);                                                \
            return 1;                                                         \
        }                                                                     \
    } while (0)
    DeviceScope device_scope_(nullptr, device);
    if (device_scope_.bad) { ocl_sc_destroy(h); return 1; }
    TRY(cudaMalloc(&h->rs.part, sizeof(double) * 16 * h->rs.max_blocks));
    TRY(cudaMalloc(&h->rs.ticket, sizeof(unsigned int) * 8));
    TRY(cudaMemset(h->rs.ticket, 0, sizeof(unsigned int) * 8));
    TRY(cudaMalloc(&h->rs.geo, sizeof(Geo)));
    TRY(cudaMemset(h->rs.geo, 0, sizeof(Geo)));
    h->rs.defer = 0;
    TRY(cudaMalloc(&h->rs.sums, sizeof(double) * 40));
    TRY(cudaMemset(h->rs.sums, 0, sizeof(double) * 40));
    h->rs.emax = h->rs.sums + 4;
    h->rs.esum = h->rs.sums + 10;
    h->rs.geom = h->rs.sums + 16;
    TRY(cudaMalloc(&h->rho, sizeof(double) * n3));
    TRY(cudaMalloc(&h->gtab, sizeof(double) * g3));
    (void)m3; (void)c3;
    {
        const char* env = getenv("OCL_SC_SOLVER");
        h->solver = (env && strcmp(env, "cufft") == 0) ? 1 : 0;
        if (h->md.mx > fft_max_length() || h->md.my > fft_max_length() || h->md.mz > fft_max_length()) h->solver = 1;
    }
    TRY(cudaMalloc(&h->h3, sizeof(double) * 4));
    TRY(cudaMalloc(&h->moments, sizeof(double) * 18));
    TRY(cudaMalloc(&h->kp_dev, sizeof(KickParams)));
    {
        const char* env = getenv("OCL_SC_GRAPH");
        h->use_graph = !(env && strcmp(env, "0") == 0);
        // Page-locking the caller's arrays from inside the library is opt-in (OCL_SC_PIN=1): the library cannot
        // know when the caller frees them.  The Python binding registers numpy buffers itself and ties the
        // registration to the owning array's lifetime (ocl_sc_host_register / ocl_sc_host_unregister).
        const char* pin = getenv("OCL_SC_PIN");
        h->pin_host = pin && strcmp(pin, "1") == 0;
    }
    if (h->solver == 0) {
        const size_t hx1 = h->md.mx / 2 + 1, hy1 = h->md.my / 2 + 1, hz1 = h->md.mz / 2 + 1;
        fft_init_kernels();
        const int ms[3] = {h->md.mx, h->md.my, h->md.mz};
        for (int a = 0; a < 3; ++a)
            if (make_twiddles(nullptr, ms[a], &h->tw[a])) { ocl_sc_destroy(h); return 1; }
        h->fw.tw_x = h->tw[0]; h->fw.tw_y = h->tw[1]; h->fw.tw_z = h->tw[2];
        TRY(cudaMalloc(&h->fw.P, sizeof(double) * nx * ny * hz1));
        TRY(cudaMalloc(&h->fw.Q, sizeof(double) * nx * hy1 * hz1));
        TRY(cudaMalloc(&h->fw.khat, sizeof(double) * hx1 * hy1 * hz1));
        TRY(cudaMalloc(&h->fw.A, sizeof(double2) * nx * ny * hz1));
        TRY(cudaMalloc(&h->fw.B, sizeof(double2) * nx * h->md.my * hz1));
    }
    TRY(cudaMalloc(&h->phi, sizeof(double) * n3));
    TRY(cudaMalloc(&h->equad, sizeof(EQuad) * (n3 * 3 + 1)));
    TRY(cudaMemset(h->equad + n3 * 3, 0, sizeof(EQuad)));          // pad record behind the x-fastest table
    {
        if (const char* dbg = getenv("OCL_SC_DEBUG_SKIP")) h->debug_skip = atoi(dbg);
        if (const char* det = getenv("OCL_SC_DETERMINISTIC")) h->ordered = atoi(det) != 0;
        const char* env = getenv("OCL_SC_GATHER");
        if (env) h->layout = atoi(env);
        if (h->layout < 0 || h->layout > 2) h->layout = (sizeof(EQuad) * n3 * 3 <= (size_t)400 << 20) ? 2 : 1;
        field_init_kernels();
    }
    TRY(cudaMemset(h->rho, 0, sizeof(double) * n3));
    TRY(cudaMemset(h->phi, 0, sizeof(double) * n3));
    h->rho_count = n3;
    TRY(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    TRY(cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking));
    TRY(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&h->ev_khat, cudaEventDisableTiming));
    TRY(cudaEventCreateWithFlags(&h->ev_handoff, cudaEventDisableTiming));
    for (int i = 0; i < T_COUNT; ++i) TRY(cudaEventCreate(&h->ev[i]));
    if (max_particles > 0) {
        TRY(cudaMalloc(&h->stage_r, sizeof(double) * 6 * max_particles));
        TRY(cudaMalloc(&h->stage_q, sizeof(double) * max_particles));
        h->stage_cap = max_particles;
    }
#undef TRY
    *out = h;
    return 0;
}

void ocl_sc_destroy(ocl_sc_t* h) {
    if (!h) return;
    DeviceScope device_scope_(nullptr, h->device);
    drop_graph(h);
    cudaFree(h->mb.epoch); cudaFree(h->mb_err);
    for (int i = 0; i < 2; ++i) if (h->pinned[i]) cudaHostUnregister(h->pinned[i]);
    cudaGetLastError();
    cudaFree(h->kp_dev);
    if (h->plans) { cufftDestroy(h->plan_fwd); cufftDestroy(h->plan_inv); }
    cudaFree(h->rs.part); cudaFree(h->rs.ticket); cudaFree(h->rs.sums); cudaFree(h->rs.geo);
    cudaFree(h->own_rho ? h->own_rho : h->rho); cudaFree(h->gtab); cudaFree(h->k1); cudaFree(h->real_buf);
    cudaFree(h->k_hat); cudaFree(h->rho_hat); cudaFree(h->own_phi ? h->own_phi : h->phi); cudaFree(h->equad);
    cudaFree(h->fw.P); cudaFree(h->fw.Q); cudaFree(h->fw.khat); cudaFree(h->fw.A); cudaFree(h->fw.B);
    cudaFree(h->tw[0]); cudaFree(h->tw[1]); cudaFree(h->tw[2]); cudaFree(h->h3); cudaFree(h->moments);
    cudaFree(h->stage_r); cudaFree(h->stage_q); cudaFree(h->cut_counts); cudaFree(h->ordered_buf);
    cudaFree(h->pw_leaf_dev); cudaFree(h->pw_node_dev); cudaFree(h->pw_level_dev); cudaFree(h->pw_values);
    if (h->cut_n) cudaFreeHost(h->cut_n);
    cudaFree(h->lw.ticket); cudaFree(h->lw.stats); cudaFree(h->lw.bins); cudaFree(h->lw.cnt); cudaFree(h->lw.Z);
    cudaFree(h->lw.spread);
    cudaFree(h->lw.tw);
    cudaFree(h->rho_slab); cudaFree(h->phi_slab);
    cudaFree(h->own_xchg_a ? h->own_xchg_a : h->xchg_a); cudaFree(h->own_xchg_b ? h->own_xchg_b : h->xchg_b);
    for (int i = 0; i < T_COUNT; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_khat) cudaEventDestroy(h->ev_khat);
    if (h->ev_handoff) cudaEventDestroy(h->ev_handoff);
    delete h;
}

int ocl_sc_collective_buffer(ocl_sc_t* h, int which, double** d_ptr, long long* count) {
    if (!h || !d_ptr || !count) return 1;
    switch (which) {
        case OCL_SC_BUF_MOMENTUM: *d_ptr = h->rs.sums; *count = 4; return 0;
        case OCL_SC_BUF_EXTENT_MAX: *d_ptr = h->rs.emax; *count = 6; return 0;
        case OCL_SC_BUF_EXTENT_SUM: *d_ptr = h->rs.esum; *count = 4; return 0;
        case OCL_SC_BUF_RHO: *d_ptr = h->rho; *count = (long long)h->rho_count; return 0;
        case OCL_SC_BUF_RHO_SLAB: if (!h->slab_world) break; *d_ptr = h->rho_slab; *count = (long long)h->sx * h->md.ny * h->md.nz; return 0;
        case OCL_SC_BUF_PHI_SLAB: if (!h->slab_world) break; *d_ptr = h->phi_slab; *count = (long long)h->sx * h->md.ny * h->md.nz; return 0;
        case OCL_SC_BUF_XCHG_A: if (!h->slab_world) break; *d_ptr = (double*)h->xchg_a; *count = 2LL * h->nx_pad * h->fs; return 0;
        case OCL_SC_BUF_XCHG_B: if (!h->slab_world) break; *d_ptr = (double*)h->xchg_b; *count = 2LL * h->nx_pad * h->fs; return 0;
        case OCL_SC_BUF_PHI: if (!h->slab_world) break; *d_ptr = h->phi; *count = (long long)h->rho_count; return 0;
        case OCL_SC_BUF_EXTENT: *d_ptr = h->rs.emax; *count = 10; return 0;
        case OCL_SC_BUF_LSC_BINS: if (!h->lsc_nb) break; *d_ptr = (double*)h->lw.bins; *count = h->lsc_nb; return 0;
        case OCL_SC_BUF_LSC_SLICE_MAX: if (!h->lsc_cap) break; *d_ptr = h->lw.slice; *count = 4; return 0;
        case OCL_SC_BUF_LSC_SLICE_SUM: if (!h->lsc_cap) break; *d_ptr = h->lw.slice + 4; *count = 5; return 0;
    }
    return fail(h, "ocl_sc_collective_buffer", "unknown buffer id");
}

int ocl_sc_combine_extents(ocl_sc_t* h, const double* d_all, int world, void* stream) {
    if (!h || !d_all || world < 1) return 1;
    ENTER_DEVICE(h);
    launch_combine_extents(d_all, world, h->rs, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_combine_extents");
}

int ocl_sc_mailbox_init(ocl_sc_t* h, int rank, int world, void* const* peer_ptrs) {
    if (!h || !peer_ptrs) return 1;
    if (world < 1 || world > 8 || rank < 0 || rank >= world)
        return fail(h, "ocl_sc_mailbox_init", "need 0 <= rank < world <= 8");
    ENTER_DEVICE(h);
    for (int w = 0; w < 8; ++w) h->mb.peer[w] = (w < world) ? (double*)peer_ptrs[w] : nullptr;
    h->mb.rank = rank; h->mb.world = world;
    if (!h->mb.epoch) {
        CU(h, cudaMalloc(&h->mb.epoch, sizeof(unsigned long long) * 4));
        CU(h, cudaMalloc(&h->mb_err, sizeof(int)));
    }
    CU(h, cudaMemset(h->mb.epoch, 0, sizeof(unsigned long long) * 4));
    CU(h, cudaMemset(h->mb_err, 0, sizeof(int)));
    return 0;
}

int ocl_sc_mailbox_exchange(ocl_sc_t* h, int which, void* stream) {
    if (!h) return 1;
    if (!h->mb.world) return fail(h, "ocl_sc_mailbox_exchange", "call ocl_sc_mailbox_init first");
    if (which < 0 || which > 2) return fail(h, "ocl_sc_mailbox_exchange", "which must be 0, 1 or 2");
    ENTER_DEVICE(h);
    launch_mailbox_exchange(h->mb, which, h->rs, h->mb_err, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_mailbox_exchange");
}

// Failure detection of the fused exchanges: a rank that waits ~4 s for a peer's flag gives up, raises a device
// flag and carries on with whatever the mailbox holds (the kernels must terminate: a hung peer would otherwise hang
// every GPU of the job).  *status = 0: every exchange so far completed; 1 / 2 / 3: a momentum / extent /
// barrier-type exchange (rho reduction, slab transposes, phi broadcast) timed out -- the results since then are
// invalid.  synchronise != 0 waits for the stream last used first.
int ocl_sc_mailbox_status(ocl_sc_t* h, int synchronise, int* status) {
    if (!h || !status) return 1;
    *status = 0;
    if (!h->mb_err) return 0;                                  // no mailbox: nothing can have timed out
    ENTER_DEVICE(h);
    if (synchronise && h->last_stream_valid) CU(h, cudaStreamSynchronize(h->last_stream));
    CU(h, cudaMemcpy(status, h->mb_err, sizeof(int), cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_set_peer_rho(ocl_sc_t* h, int rank, int world, void* const* peer_rho) {
    if (!h || !peer_rho) return 1;
    if (world < 1 || world > 8 || rank < 0 || rank >= world)
        return fail(h, "ocl_sc_set_peer_rho", "need 0 <= rank < world <= 8");
    ENTER_DEVICE(h);
    drop_graph(h);
    if (!h->own_rho) h->own_rho = h->rho;
    for (int w = 0; w < 8; ++w) h->peer_rho.p[w] = (w < world) ? (const double*)peer_rho[w] : nullptr;
    h->peer_rho.world = world;
    h->rho = (double*)peer_rho[rank];
    CU(h, cudaMemset(h->rho, 0, sizeof(double) * h->rho_count));
    return 0;
}

int ocl_sc_set_multicast_rho(ocl_sc_t* h, void* local_rho, void* multicast_rho) {
    if (!h || !local_rho || !multicast_rho) return 1;
    if (!h->mb.world) return fail(h, "ocl_sc_set_multicast_rho", "call ocl_sc_mailbox_init first");
    ENTER_DEVICE(h);
    drop_graph(h);
    if (!h->own_rho) h->own_rho = h->rho;
    h->peer_rho.world = 0;                     // the solve reads the local (already reduced) grid
    h->rho = (double*)local_rho;
    h->mc_rho = (double*)multicast_rho;
    CU(h, cudaMemset(h->rho, 0, sizeof(double) * h->rho_count));
    return 0;
}

int ocl_sc_set_multicast_phi(ocl_sc_t* h, void* local_phi, void* multicast_phi) {
    if (!h || !local_phi || !multicast_phi) return 1;
    if (!h->slab_world) return fail(h, "ocl_sc_set_multicast_phi", "call ocl_sc_slab_init first");
    if (!h->mb.world) return fail(h, "ocl_sc_set_multicast_phi", "call ocl_sc_mailbox_init first");
    ENTER_DEVICE(h);
    drop_graph(h);
    if (!h->own_phi) h->own_phi = h->phi;
    h->phi = (double*)local_phi;
    h->mc_phi = (double*)multicast_phi;
    CU(h, cudaMemset(h->phi, 0, sizeof(double) * h->rho_count));
    return 0;
}

int ocl_sc_nvls_reduce_rho(ocl_sc_t* h, void* stream) {
    OCL_RANGE("ocl_sc_nvls_reduce_rho");
    if (!h) return 1;
    if (!h->mc_rho) return fail(h, "ocl_sc_nvls_reduce_rho", "call ocl_sc_set_multicast_rho first");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (h->debug_skip & 4) return 0;
    const int world = h->mb.world, rank = h->mb.rank;
    // barrier ("every rank's deposit is complete"), in-switch reduction, barrier ("every slice final")
    if (h->slab_world) {                                               // reduce-scatter into this rank's x-slab
        const long long plane = (long long)h->md.ny * h->md.nz;
        const long long lo = (long long)h->slab_rank * h->sx * plane;
        launch_nvls_reduce(h->mc_rho, lo, lo + (long long)h->sx * plane, h->rho_slab, h->mb, h->rs, h->rs.ticket + 4, h->mb_err, st);
    } else {                                                           // all-reduce in place
        const long long n3 = (long long)h->rho_count;
        const long long chunk = (n3 + world - 1) / world;
        const long long lo = std::min(n3, (long long)rank * chunk), hi = std::min(n3, lo + chunk);
        launch_nvls_reduce(h->mc_rho, lo, hi, nullptr, h->mb, h->rs, h->rs.ticket + 4, h->mb_err, st);
    }
    h->launches += 3;
    return check_launch(h, "k_nvls_reduce");
}

int ocl_sc_use_device_params(ocl_sc_t* h, int on) {
    if (!h) return 1;
    h->cur_pp = on ? h->kp_dev : nullptr;
    return 0;
}

// Ordered deposit on / off (SURVEY section 8e "Determinism"): every cell's charges are added in ascending particle
// order, as np.bincount does (sc.py:193), instead of by L2 atomics in arrival order.  For debugging cell-flip /
// reproducibility questions: the whole kick takes 2.5-3x as long (measured, tools/r2_ordered_cost.py) and runs stage by
// stage instead of as one graph.  Also switches sweep 1 to numpy's pairwise tree over exactly rounded momenta.
int ocl_sc_set_deterministic(ocl_sc_t* h, int on) {
    if (!h) return 1;
    h->ordered = on != 0;
    return 0;
}

int ocl_sc_defer_finish(ocl_sc_t* h, int on) {
    if (!h) return 1;
    h->rs.defer = on ? 1 : 0;
    return 0;
}

int ocl_sc_stage_finish(ocl_sc_t* h, int which, double E_GeV, const double* mesh_draws, void* stream) {
    if (!h) return 1;
    if (which != 0 && which != 1) return fail(h, "ocl_sc_stage_finish", "which must be 0 (momentum) or 1 (extent)");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    launch_finish(which, kp_of(h, E_GeV, 0.0, mesh_draws), h->rs, h->md, st);
    h->launches += 1;
    return check_launch(h, "k_finish");
}

int ocl_sc_set_kick_params(ocl_sc_t* h, double E_GeV, double dz, const double* mesh_draws, void* stream) {
    if (!h) return 1;
    if (!(E_GeV > 0.0)) return fail(h, "ocl_sc_set_kick_params", "beam energy must be positive");
    ENTER_DEVICE(h);
    launch_set_params(kick_params(h, E_GeV, dz, mesh_draws), h->kp_dev, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_set_params");
}

// ---- slab-decomposed solve ---------------------------------------------------
int ocl_sc_slab_init(ocl_sc_t* h, int rank, int world) {
    if (!h) return 1;
    if (world < 1 || rank < 0 || rank >= world) return fail(h, "ocl_sc_slab_init", "bad rank/world");
    if (h->solver != 0) return fail(h, "ocl_sc_slab_init", "slab mode needs the hand-written solver");
    ENTER_DEVICE(h);
    drop_graph(h);
    const int hz1 = h->md.mz / 2 + 1;
    const int F = h->md.my * hz1;
    h->slab_world = world; h->slab_rank = rank;
    h->sx = (h->md.nx + world - 1) / world;
    h->nx_pad = h->sx * world;
    h->fs = (F + world - 1) / world;
    const size_t plane = (size_t)h->md.ny * h->md.nz;
    cudaFree(h->own_rho ? h->own_rho : h->rho); cudaFree(h->own_phi ? h->own_phi : h->phi);
    h->own_rho = nullptr; h->peer_rho.world = 0; h->own_phi = nullptr; h->mc_phi = nullptr;
    cudaFree(h->rho_slab); cudaFree(h->phi_slab);
    cudaFree(h->own_xchg_a ? h->own_xchg_a : h->xchg_a); cudaFree(h->own_xchg_b ? h->own_xchg_b : h->xchg_b);
    h->own_xchg_a = h->own_xchg_b = nullptr; h->peer_xchg = PeerXchg{};
    h->rho = h->phi = h->rho_slab = h->phi_slab = nullptr; h->xchg_a = h->xchg_b = nullptr;
    h->rho_count = (size_t)h->nx_pad * plane;
    CU(h, cudaMalloc(&h->rho, sizeof(double) * h->rho_count));
    CU(h, cudaMalloc(&h->phi, sizeof(double) * h->rho_count));
    CU(h, cudaMalloc(&h->rho_slab, sizeof(double) * h->sx * plane));
    CU(h, cudaMalloc(&h->phi_slab, sizeof(double) * h->sx * plane));
    CU(h, cudaMalloc(&h->xchg_a, sizeof(double2) * (size_t)h->nx_pad * h->fs));
    CU(h, cudaMalloc(&h->xchg_b, sizeof(double2) * (size_t)h->nx_pad * h->fs));
    CU(h, cudaMemset(h->rho, 0, sizeof(double) * h->rho_count));
    CU(h, cudaMemset(h->phi, 0, sizeof(double) * h->rho_count));
    CU(h, cudaMemset(h->xchg_a, 0, sizeof(double2) * (size_t)h->nx_pad * h->fs));
    CU(h, cudaMemset(h->xchg_b, 0, sizeof(double2) * (size_t)h->nx_pad * h->fs));
    return 0;
}

int ocl_sc_set_peer_xchg(ocl_sc_t* h, int rank, int world, void* const* peer_a, void* const* peer_b) {
    if (!h || !peer_a || !peer_b) return 1;
    if (!h->slab_world) return fail(h, "ocl_sc_set_peer_xchg", "call ocl_sc_slab_init first");
    if (!h->mb.world) return fail(h, "ocl_sc_set_peer_xchg", "call ocl_sc_mailbox_init first");
    if (world != h->slab_world || rank != h->slab_rank || world > 8)
        return fail(h, "ocl_sc_set_peer_xchg", "rank / world differ from ocl_sc_slab_init (world <= 8)");
    ENTER_DEVICE(h);
    drop_graph(h);
    if (!h->own_xchg_a) { h->own_xchg_a = h->xchg_a; h->own_xchg_b = h->xchg_b; }
    for (int w = 0; w < 8; ++w) {
        h->peer_xchg.a[w] = w < world ? (double2*)peer_a[w] : nullptr;
        h->peer_xchg.b[w] = w < world ? (double2*)peer_b[w] : nullptr;
    }
    h->peer_xchg.rank = rank; h->peer_xchg.world = world;
    h->xchg_a = (double2*)peer_a[rank];
    h->xchg_b = (double2*)peer_b[rank];
    const size_t bytes = sizeof(double2) * (size_t)h->nx_pad * h->fs;
    CU(h, cudaMemset(h->xchg_a, 0, bytes));          // padding planes / lines are never written: they must read as zero
    CU(h, cudaMemset(h->xchg_b, 0, bytes));
    return 0;
}

int ocl_sc_slab_forward(ocl_sc_t* h, void* stream) {
    OCL_RANGE("ocl_sc_slab_forward");
    if (!h || !h->slab_world) return h ? fail(h, "ocl_sc_slab_forward", "call ocl_sc_slab_init first") : 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    launch_slab_forward(h->rho_slab, h->peer_rho, (long long)h->slab_rank * h->sx * h->md.ny, h->md, h->sx, h->fs, h->fw,
                        h->xchg_a, h->peer_xchg, st);
    h->launches += 2;
    if (h->peer_xchg.world > 0) {            // every rank's y output has landed in this rank's x-pass input
        launch_mailbox_exchange(h->mb, 2, h->rs, h->mb_err, st);
        h->launches += 1;
    }
    return check_launch(h, "slab_forward");
}

int ocl_sc_slab_xpass(ocl_sc_t* h, void* stream) {
    OCL_RANGE("ocl_sc_slab_xpass");
    if (!h || !h->slab_world) return h ? fail(h, "ocl_sc_slab_xpass", "call ocl_sc_slab_init first") : 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (h->khat_pending) {                      // join the Green's-function chain forked by stage_deposit
        CU(h, cudaStreamWaitEvent(st, h->ev_khat, 0));
        h->khat_pending = false;
    } else {
        return fail(h, "ocl_sc_slab_xpass", "K_hat not available: ocl_sc_stage_deposit must precede the solve");
    }
    launch_slab_xpass(h->xchg_b, h->md, h->sx, h->fs, h->slab_rank * h->fs, h->fw, h->peer_xchg, st);
    h->launches += 1;
    if (h->peer_xchg.world > 0) {            // every rank's x output has landed in this rank's inverse-y input
        launch_mailbox_exchange(h->mb, 2, h->rs, h->mb_err, st);
        h->launches += 1;
    }
    return check_launch(h, "slab_xpass");
}

int ocl_sc_slab_inverse(ocl_sc_t* h, void* stream) {
    OCL_RANGE("ocl_sc_slab_inverse");
    if (!h || !h->slab_world) return h ? fail(h, "ocl_sc_slab_inverse", "call ocl_sc_slab_init first") : 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (h->mc_phi) {     // phi slab broadcast through the NVSwitch by the inverse z pass itself, then "every slab has landed"
        const size_t off = (size_t)h->slab_rank * h->sx * h->md.ny * h->md.nz;
        launch_slab_inverse(h->xchg_a, h->md, h->sx, h->fs, h->fw, h->h3, four_pi_eps0_value(), h->mc_phi + off, 1, st);
        launch_mailbox_exchange(h->mb, 2, h->rs, h->mb_err, st);
        h->launches += 1;
    } else {
        launch_slab_inverse(h->xchg_a, h->md, h->sx, h->fs, h->fw, h->h3, four_pi_eps0_value(), h->phi_slab, 0, st);
    }
    h->launches += 2;
    mark(h, T_SOLVE, st);
    return check_launch(h, "slab_inverse");
}

int ocl_sc_slab_finish(ocl_sc_t* h, const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_slab_finish");
    if (!h || !h->slab_world) return h ? fail(h, "ocl_sc_slab_finish", "call ocl_sc_slab_init first") : 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    (void)mesh_draws;
    launch_field(h->phi, h->rs, h->md, h->equad, h->layout ? 1 : 0, st);
    h->launches += 1;
    mark(h, T_FIELD, st);
    return check_launch(h, "slab_finish");
}

// ---- stages ---------------------------------------------------------------
// numpy's pairwise summation of a row of n elements as a tree (numpy/_core/src/umath/loops_utils.h.src,
// pairwise_sum: n > 128 splits at n/2 rounded down to a multiple of 8; shorter ranges are leaves).  Leaves are
// (offset, length); internal nodes are (left, right) value indices, sorted by height so that one block can combine
// them level by level; value index = leaf index, or leaf count + position in the sorted node list.
static int ensure_pairwise_plan(ocl_sc* h, long long n, cudaStream_t st) {
    if (h->pw_n == n) return 0;
    struct Node { unsigned l, r; int height; };
    std::vector<uint2> leaves;
    std::vector<Node> nodes;                                   // post order: children before parents
    struct Ref { unsigned id; int height; bool leaf; };
    // explicit recursion (depth <= ~24)
    struct Builder {
        std::vector<uint2>& leaves; std::vector<Node>& nodes;
        Ref build(long long off, long long len) {
            if (len <= 128) {
                leaves.push_back(make_uint2((unsigned)off, (unsigned)len));
                return Ref{(unsigned)(leaves.size() - 1), 0, true};
            }
            long long n2 = len / 2;
            n2 -= n2 % 8;
            const Ref a = build(off, n2), b = build(off + n2, len - n2);
            // leaf ids are final; internal ids are post-order positions tagged with the top bit until the sort below
            nodes.push_back(Node{a.leaf ? a.id : (a.id | 0x80000000u), b.leaf ? b.id : (b.id | 0x80000000u),
                                 std::max(a.height, b.height) + 1});
            return Ref{(unsigned)(nodes.size() - 1), nodes.back().height, false};
        }
    } builder{leaves, nodes};
    builder.build(0, n);
    const int nleaves = (int)leaves.size(), nnodes = (int)nodes.size();
    int levels = 0;
    for (const Node& nd : nodes) levels = std::max(levels, nd.height);
    // stable counting sort by height; remap the tagged post-order ids to sorted positions
    std::vector<int> level_start(levels + 1, 0), pos(nnodes);
    for (const Node& nd : nodes) level_start[nd.height]++;                  // counts at [height], heights are 1-based
    { int acc = 0; for (int l = 1; l <= levels; ++l) { const int c = level_start[l]; level_start[l - 1] = acc; acc += c; }
      level_start[levels] = acc; }
    { std::vector<int> cur(level_start.begin(), level_start.end());
      for (int k = 0; k < nnodes; ++k) pos[k] = cur[nodes[k].height - 1]++; }
    std::vector<uint2> sorted(nnodes);
    for (int k = 0; k < nnodes; ++k) {
        auto remap = [&](unsigned id) { return (id & 0x80000000u) ? (unsigned)(nleaves + pos[id & 0x7fffffffu]) : id; };
        sorted[pos[k]] = make_uint2(remap(nodes[k].l), remap(nodes[k].r));
    }
    CU(h, cudaStreamSynchronize(st));                                       // an earlier kick may still read the old plan
    cudaFree(h->pw_leaf_dev); cudaFree(h->pw_node_dev); cudaFree(h->pw_level_dev); cudaFree(h->pw_values);
    h->pw_leaf_dev = h->pw_node_dev = nullptr; h->pw_level_dev = nullptr; h->pw_values = nullptr; h->pw_n = -1;
    CU(h, cudaMalloc(&h->pw_leaf_dev, sizeof(uint2) * nleaves));
    CU(h, cudaMalloc(&h->pw_node_dev, sizeof(uint2) * std::max(nnodes, 1)));
    CU(h, cudaMalloc(&h->pw_level_dev, sizeof(int) * (levels + 1)));
    CU(h, cudaMalloc(&h->pw_values, sizeof(double) * 3 * (size_t)(nleaves + nnodes)));
    CU(h, cudaMemcpy(h->pw_leaf_dev, leaves.data(), sizeof(uint2) * nleaves, cudaMemcpyHostToDevice));
    if (nnodes) CU(h, cudaMemcpy(h->pw_node_dev, sorted.data(), sizeof(uint2) * nnodes, cudaMemcpyHostToDevice));
    CU(h, cudaMemcpy(h->pw_level_dev, level_start.data(), sizeof(int) * (levels + 1), cudaMemcpyHostToDevice));
    h->pw_n = n; h->pw_leaves = nleaves; h->pw_levels = levels;
    return 0;
}

int ocl_sc_stage_momentum(ocl_sc_t* h, const double* d_r, long long ld, long long n, double E_GeV, void* stream) {
    OCL_RANGE("ocl_sc_stage_momentum");
    if (!h) return 1;
    if (n < 0 || ld < n) return fail(h, "ocl_sc_stage_momentum", "need 0 <= n <= ld");
    if (n == 0 && h->mb.world <= 1 && !h->rs.defer) return fail(h, "ocl_sc_stage_momentum", "empty bunch");
    if (n >= 2147483647LL - 2 * 148 * 4 * 256) return fail(h, "ocl_sc_stage_momentum", "more than 2^31 particles per GPU");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    mark(h, T_BEGIN, st);
    Mailbox mb = h->mb;
    if (h->debug_skip & 1) mb.world = 1;
    KP kp = kp_of(h, E_GeV, 0.0, nullptr);
    KickParams* publish = nullptr;
    if (h->publish_params) {                  // the graph's parameter node: scalars by value, published for the rest
        kp.p = nullptr;
        publish = h->kp_dev;
        h->g_mb = mb; h->g_rs = h->rs;
    }
    if (h->ordered && mb.world <= 1 && n > 0 && !publish) {
        // ordered mode: np.mean's own summation tree over momenta rounded exactly as the reference rounds them
        if (ensure_pairwise_plan(h, n, st)) return 1;
        launch_momentum_exact(d_r, ld, n, kp, h->pw_leaf_dev, h->pw_leaves, h->pw_node_dev, h->pw_level_dev,
                              h->pw_levels, h->pw_values, h->rs.sums, st);
        h->launches += 2;
        if (!h->rs.defer) {
            launch_finish(0, kp, h->rs, h->md, st);
            h->launches += 1;
        }
        mark(h, T_MOM, st);
        return check_launch(h, "k_momentum_exact");
    }
    launch_momentum(d_r, ld, n, kp, h->rs, mb, h->mb_err, publish, st);
    h->launches += 1;
    mark(h, T_MOM, st);
    return check_launch(h, "k_momentum");
}

int ocl_sc_stage_extent(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n, double E_GeV,
                        const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_stage_extent");
    if (!h) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    Mailbox mb = h->mb;
    if (h->debug_skip & 2) mb.world = 1;
    launch_extent(d_r, ld, d_q, n, kp_of(h, E_GeV, 0.0, mesh_draws), h->rs, h->md, mb, h->mb_err, st);
    h->launches += 1;
    mark(h, T_EXT, st);
    return check_launch(h, "k_extent");
}

int ocl_sc_stage_deposit(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n, double E_GeV,
                         const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_stage_deposit");
    if (!h) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    // the mesh steps are final once the extents are reduced: start the Green's-function / K_hat
    // chain now, concurrently with the deposit and the first two rho passes
    if (h->solver == 0 && fork_khat(h, st)) return 1;
    CU(h, cudaMemsetAsync(h->rho, 0, sizeof(double) * h->rho_count, st));
    if (h->ordered && n > 0) {
        const size_t need = deposit_ordered_scratch_bytes(n, h->md);
        if (need > h->ordered_bytes) {
            CU(h, cudaStreamSynchronize(st));                       // an earlier kick may still be using the old scratch
            cudaFree(h->ordered_buf); h->ordered_buf = nullptr; h->ordered_bytes = 0;
            CU(h, cudaMalloc(&h->ordered_buf, need + need / 8));
            h->ordered_bytes = need + need / 8;
        }
        if (launch_deposit_ordered(d_r, ld, d_q, n, kp_of(h, E_GeV, 0.0, mesh_draws), h->rs, h->md, h->rho,
                                   h->ordered_buf, h->ordered_bytes, st))
            return fail(h, "ocl_sc_stage_deposit", "ordered deposit: radix sort failed");
        h->launches += 3;                                           // + the sort's own kernels (library code, not counted)
    } else {
        launch_deposit(d_r, ld, d_q, n, kp_of(h, E_GeV, 0.0, mesh_draws), h->rs, h->md, h->rho, st);
        h->launches += 2;
    }
    mark(h, T_DEP, st);
    return check_launch(h, "k_deposit");
}

int ocl_sc_stage_solve(ocl_sc_t* h, const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_stage_solve");
    if (!h) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    (void)mesh_draws;                         // the mesh (incl. random_mesh draws) was fixed by the extent stage
    if (!h->khat_pending) {
        launch_green_table(h->rs, h->md, h->gtab, h->h3, st);
        h->launches += 1;
    }
    if (h->solver == 0) {
        if (solve_fused(h, st)) return 1;
        mark(h, T_SOLVE, st);
    } else {
        if (convolve(h, st)) return 1;
        mark(h, T_SOLVE, st);
        launch_crop_phi(h->real_buf, h->rs, h->md, h->phi, st);
        h->launches += 1;
    }
    launch_field(h->phi, h->rs, h->md, h->equad, h->layout ? 1 : 0, st);
    h->launches += 1;
    mark(h, T_FIELD, st);
    return check_launch(h, "stage_solve");
}

int ocl_sc_stage_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, double E_GeV, double dz,
                      const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_stage_kick");
    if (!h) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    launch_gather_kick(d_r, ld, n, kp_of(h, E_GeV, dz, mesh_draws), h->rs, h->md, h->equad, nullptr, 1, h->layout, st);
    h->launches += 1;
    mark(h, T_KICK, st);
    return check_launch(h, "k_gather_kick");
}

static int run_stages(ocl_sc_t* h, double* d_r, long long ld, const double* d_q, long long n, double E_GeV,
                      double dz, const double* mesh_draws, void* stream) {
    if (ocl_sc_stage_momentum(h, d_r, ld, n, E_GeV, stream)) return 1;
    if (ocl_sc_stage_extent(h, d_r, ld, d_q, n, E_GeV, mesh_draws, stream)) return 1;
    if (ocl_sc_stage_deposit(h, d_r, ld, d_q, n, E_GeV, mesh_draws, stream)) return 1;
    // particle-sharded handle whose exchanges are all the library's own kernels (peer-memory mailbox for the two
    // scalar exchanges, multicast mapping for rho, redundant solve): the WHOLE sharded kick is this one call, so it
    // is captured into the library's own graph with its parameter node (no caller-side capture, no separate
    // parameter kernel per kick)
    if (h->mc_rho && h->mb.world > 1 && !h->slab_world && ocl_sc_nvls_reduce_rho(h, stream)) return 1;
    if (ocl_sc_stage_solve(h, mesh_draws, stream)) return 1;
    return ocl_sc_stage_kick(h, d_r, ld, n, E_GeV, dz, mesh_draws, stream);
}

static void drop_graph(ocl_sc* h) {
    if (h->graph_exec) cudaGraphExecDestroy(h->graph_exec);
    if (h->graph) cudaGraphDestroy(h->graph);
    h->graph_exec = nullptr; h->graph = nullptr; h->param_node = nullptr;
}

// Capture the whole kick (15 kernels, the first of which doubles as the parameter node, + memset + the
// side-stream fork/join) once.
static int capture_kick(ocl_sc* h, double* d_r, long long ld, const double* d_q, long long n, double E_GeV,
                        double dz, const double* mesh_draws) {
    drop_graph(h);
    cudaStream_t cs = h->own_stream;
    CU(h, cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    h->cur_pp = h->kp_dev;
    h->publish_params = true;                 // k_momentum carries the scalars by value and stores them in kp_dev
    const long long before = h->launches;
    int rc = run_stages(h, d_r, ld, d_q, n, E_GeV, dz, mesh_draws, cs);
    h->cur_pp = nullptr;
    h->publish_params = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(cs, &g);
    h->graph_launches_per_kick = h->launches - before;
    h->launches = before;
    if (rc || e != cudaSuccess || !g) {
        if (g) cudaGraphDestroy(g);
        h->khat_pending = false;
        return rc ? rc : fail(h, "cudaStreamEndCapture", cudaGetErrorString(e));
    }
    h->graph = g;
    CU(h, cudaGraphInstantiate(&h->graph_exec, g, 0));
    size_t nn = 0;
    CU(h, cudaGraphGetNodes(g, nullptr, &nn));
    std::vector<cudaGraphNode_t> nodes(nn);
    CU(h, cudaGraphGetNodes(g, nodes.data(), &nn));
    for (size_t i = 0; i < nn; ++i) {
        cudaGraphNodeType t;
        if (cudaGraphNodeGetType(nodes[i], &t) != cudaSuccess || t != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) == cudaSuccess &&
            (kp.func == momentum_kernel(0) || kp.func == momentum_kernel(1))) {
            h->param_node = nodes[i];
            h->param_np = kp;
            break;
        }
    }
    if (!h->param_node) { drop_graph(h); return fail(h, "capture_kick", "parameter node not found"); }
    h->g_r = d_r; h->g_q = d_q; h->g_ld = ld; h->g_n = n;
    return 0;
}

int ocl_sc_kick_device(ocl_sc_t* h, double* d_r, long long ld, const double* d_q, long long n, double E_GeV,
                       double dz, const double* mesh_draws, void* stream) {
    OCL_RANGE("ocl_sc_kick_device");
    if (!h) return 1;
    if (dz == 0.0) return 0;   // sc.py:210-212
    if (!(E_GeV > 0.0)) return fail(h, "ocl_sc_kick_device", "beam energy must be positive");
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_kick_device", "need 0 < n <= ld");
    cudaStream_t st = (cudaStream_t)stream;
    bool graph_ok = h->use_graph && h->solver == 0 && !h->timers && !h->ordered;
    if (graph_ok) {
        ENTER_DEVICE(h);
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            graph_ok = false;       // the caller is capturing us into a graph of its own
        }
    }
    if (!graph_ok) return run_stages(h, d_r, ld, d_q, n, E_GeV, dz, mesh_draws, stream);
    if (!h->graph_exec || h->g_r != d_r || h->g_q != d_q || h->g_ld != ld || h->g_n != n) {
        const cudaStream_t prev = h->last_stream;          // the capture pass runs the stages on own_stream
        const bool prev_valid = h->last_stream_valid;      // without executing anything: keep the real history
        const int rc = capture_kick(h, d_r, ld, d_q, n, E_GeV, dz, mesh_draws);
        h->last_stream = prev; h->last_stream_valid = prev_valid;
        if (rc) return 1;
    }
    if (adopt_stream(h, st)) return 1;
    // refresh the parameter node (k_momentum): same launch shape and pointers as captured, this kick's scalars
    KP kp;
    kp.v = kick_params(h, E_GeV, dz, mesh_draws);
    kp.p = nullptr;
    const double* a_r = d_r;
    long long a_ld = ld, a_n = n;
    void* args[8] = {&a_r, &a_ld, &a_n, &kp, &h->g_rs, &h->g_mb, &h->mb_err, &h->kp_dev};
    cudaKernelNodeParams np = h->param_np;
    np.kernelParams = args; np.extra = nullptr;
    CU(h, cudaGraphExecKernelNodeSetParams(h->graph_exec, h->param_node, &np));
    CU(h, cudaGraphLaunch(h->graph_exec, st));
    h->launches += h->graph_launches_per_kick;
    return 0;
}

// Page-lock a host range so the copies of ocl_sc_kick_host run at PCIe rate (pageable: ~13 GB/s, pinned:
// ~50 GB/s measured).  The caller owns the lifetime: it must unregister BEFORE the memory is freed (freeing
// registered memory is undefined, and a new allocation at the same address would inherit a stale mapping).
// Returns 0 (registered), 2 (already page-locked / managed: nothing to do), 1 (failed; the copies still work).
int ocl_sc_host_register(void* ptr, long long bytes) {
    if (!ptr || bytes <= 0) return 1;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) {
        cudaGetLastError();
        return 2;
    }
    cudaGetLastError();
    if (cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();
        return 1;
    }
    return 0;
}

int ocl_sc_host_unregister(void* ptr) {
    if (!ptr) return 1;
    const cudaError_t e = cudaHostUnregister(ptr);
    cudaGetLastError();
    return e == cudaSuccess ? 0 : 1;
}

// Legacy opt-in (OCL_SC_PIN=1): the handle registers the buffers it is given and re-validates on every call.
static void pin_in_place(ocl_sc* h, int slot, const void* ptr, size_t bytes) {
    if (!h->pin_host) return;
    if (h->pinned[slot]) { cudaHostUnregister(h->pinned[slot]); h->pinned[slot] = nullptr; cudaGetLastError(); }
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, ptr) == cudaSuccess && at.type != cudaMemoryTypeUnregistered) {
        cudaGetLastError();
        return;                                   // already page-locked or managed
    }
    cudaGetLastError();
    if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterDefault) == cudaSuccess) {
        h->pinned[slot] = const_cast<void*>(ptr);
        h->pinned_bytes[slot] = bytes;
    } else {
        cudaGetLastError();
    }
}

static int ensure_stage(ocl_sc* h, long long n) {
    if (n <= h->stage_cap) return 0;
    cudaFree(h->stage_r); cudaFree(h->stage_q);
    h->stage_r = h->stage_q = nullptr; h->stage_cap = 0;
    long long cap = n + n / 8 + 1024;
    cap = (cap + 31) / 32 * 32;
    CU(h, cudaMalloc(&h->stage_r, sizeof(double) * 6 * cap));
    CU(h, cudaMalloc(&h->stage_q, sizeof(double) * cap));
    h->stage_cap = cap;
    return 0;
}

int ocl_sc_kick_host(ocl_sc_t* h, double* h_r, long long ld, const double* h_q, long long n, double E_GeV, double dz,
                     const double* mesh_draws) {
    OCL_RANGE("ocl_sc_kick_host");
    if (!h) return 1;
    if (dz == 0.0) return 0;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_kick_host", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    if (ensure_stage(h, n)) return 1;
    pin_in_place(h, 0, h_r, sizeof(double) * (size_t)(5 * ld + n));
    pin_in_place(h, 1, h_q, sizeof(double) * (size_t)n);
    cudaStream_t st = h->own_stream;
    // contiguous host rows (numpy's rparticles): one flat copy each way, device pitch = n;
    // otherwise a pitched copy into rows of the staging capacity
    const bool flat = (ld == n);
    const long long pitch = flat ? n : h->stage_cap;
    if (flat) {
        CU(h, cudaMemcpyAsync(h->stage_r, h_r, sizeof(double) * 6 * n, cudaMemcpyHostToDevice, st));
    } else {
        CU(h, cudaMemcpy2DAsync(h->stage_r, sizeof(double) * pitch, h_r, sizeof(double) * ld, sizeof(double) * n, 6,
                                cudaMemcpyHostToDevice, st));
    }
    CU(h, cudaMemcpyAsync(h->stage_q, h_q, sizeof(double) * n, cudaMemcpyHostToDevice, st));
    if (ocl_sc_kick_device(h, h->stage_r, pitch, h->stage_q, n, E_GeV, dz, mesh_draws, st)) return 1;
    if (flat) {
        CU(h, cudaMemcpyAsync(h_r, h->stage_r, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, st));
    } else {
        CU(h, cudaMemcpy2DAsync(h_r, sizeof(double) * ld, h->stage_r, sizeof(double) * pitch, sizeof(double) * n, 6,
                                cudaMemcpyDeviceToHost, st));
    }
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// ---- taps -----------------------------------------------------------------
static int sync_last(ocl_sc* h) {
    ENTER_DEVICE(h);
    if (h->last_stream_valid) CU(h, cudaStreamSynchronize(h->last_stream));
    return 0;
}

int ocl_sc_get_geometry(ocl_sc_t* h, double out[24]) {
    if (!h) return 1;
    if (sync_last(h)) return 1;
    CU(h, cudaMemcpy(out, h->rs.geom, sizeof(double) * 24, cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_get_rho(ocl_sc_t* h, double* h_out) {
    if (!h) return 1;
    if (sync_last(h)) return 1;
    CU(h, cudaMemcpy(h_out, h->rho, sizeof(double) * (size_t)h->md.nx * h->md.ny * h->md.nz, cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_get_phi(ocl_sc_t* h, double* h_out) {
    if (!h) return 1;
    if (sync_last(h)) return 1;
    CU(h, cudaMemcpy(h_out, h->phi, sizeof(double) * (size_t)h->md.nx * h->md.ny * h->md.nz, cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_get_green(ocl_sc_t* h, double* h_out) {
    if (!h) return 1;
    if (sync_last(h)) return 1;
    const size_t n3 = (size_t)h->md.nx * h->md.ny * h->md.nz;
    if (!h->k1) CU(h, cudaMalloc(&h->k1, sizeof(double) * n3));
    launch_green_compact(h->gtab, h->md, h->k1, h->last_stream);
    if (check_launch(h, "k_green_compact")) return 1;
    CU(h, cudaStreamSynchronize(h->last_stream));
    CU(h, cudaMemcpy(h_out, h->k1, sizeof(double) * n3, cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_field_at_particles(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n,
                              double E_GeV, const double* mesh_draws, double* d_exyz, void* stream) {
    if (!h) return 1;
    if (ocl_sc_stage_momentum(h, d_r, ld, n, E_GeV, stream)) return 1;
    if (ocl_sc_stage_extent(h, d_r, ld, d_q, n, E_GeV, mesh_draws, stream)) return 1;
    if (ocl_sc_stage_deposit(h, d_r, ld, d_q, n, E_GeV, mesh_draws, stream)) return 1;
    if (ocl_sc_stage_solve(h, mesh_draws, stream)) return 1;
    launch_gather_kick(const_cast<double*>(d_r), ld, n, kp_of(h, E_GeV, 0.0, mesh_draws), h->rs, h->md, h->equad,
                       d_exyz, 0, h->layout, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_gather");
}

int ocl_sc_mad_to_cartesian(ocl_sc_t* h, const double* d_r, long long ld, long long n, double E_GeV, double* d_xp,
                            long long ld_xp, void* stream) {
    if (!h) return 1;
    ENTER_DEVICE(h);
    if (adopt_stream(h, (cudaStream_t)stream)) return 1;
    launch_mad_to_cart(d_r, ld, n, ref_params(h, E_GeV), d_xp, ld_xp, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_mad_to_cart");
}

int ocl_sc_cartesian_to_mad(ocl_sc_t* h, const double* d_xp, long long ld_xp, long long n, double E_GeV, double* d_r,
                            long long ld, void* stream) {
    if (!h) return 1;
    ENTER_DEVICE(h);
    if (adopt_stream(h, (cudaStream_t)stream)) return 1;
    launch_cart_to_mad(d_xp, ld_xp, n, ref_params(h, E_GeV), d_r, ld, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_cart_to_mad");
}

int ocl_sc_potential_host(ocl_sc_t* h, const double* h_rho, const double steps[3], double* h_phi) {
    if (!h) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = h->own_stream;
    if (adopt_stream(h, st)) return 1;
    const size_t n3 = (size_t)h->md.nx * h->md.ny * h->md.nz;
    CU(h, cudaMemcpyAsync(h->rho, h_rho, sizeof(double) * n3, cudaMemcpyHostToDevice, st));
    launch_green_table_steps(steps, h->md, h->gtab, h->h3, st);
    h->launches += 1;
    if (h->solver == 0) {
        if (solve_fused(h, st)) return 1;
    } else {
        if (convolve(h, st)) return 1;
        launch_crop_phi_steps(h->real_buf, steps, h->md, h->phi, st);
        h->launches += 1;
    }
    if (check_launch(h, "potential")) return 1;
    CU(h, cudaMemcpyAsync(h_phi, h->phi, sizeof(double) * n3, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

int ocl_sc_map_apply(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* R, const double* B,
                     const double* T, void* stream) {
    OCL_RANGE("ocl_sc_map_apply");
    if (!h || !R) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_map_apply", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    MapCoef mc;
    for (int i = 0; i < 36; ++i) mc.R[i] = R[i];
    for (int i = 0; i < 6; ++i) mc.B[i] = B ? B[i] : 0.0;
    mc.nt = 0;
    mc.cav = 0;
    if (T)
        for (int c = 0; c < 216; ++c)
            if (T[c] != 0.0) { mc.tval[mc.nt] = T[c]; mc.tidx[mc.nt] = (unsigned char)c; ++mc.nt; }
    if (adopt_stream(h, (cudaStream_t)stream)) return 1;
    launch_map_apply(d_r, ld, n, mc, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_map_apply");
}

int ocl_sc_cavity_apply(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* R, const double* B,
                        const double* c, int mode, void* stream) {
    OCL_RANGE("ocl_sc_cavity_apply");
    if (!h || !R || !c) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_cavity_apply", "need 0 < n <= ld");
    if (mode != 1 && mode != 2) return fail(h, "ocl_sc_cavity_apply", "mode must be 1 or 2");
    ENTER_DEVICE(h);
    MapCoef mc;
    for (int i = 0; i < 36; ++i) mc.R[i] = R[i];
    for (int i = 0; i < 6; ++i) mc.B[i] = B ? B[i] : 0.0;
    mc.nt = 0;
    mc.cav = mode;
    mc.c1 = c[0]; mc.c2 = c[1]; mc.kb = c[2]; mc.phi = c[3]; mc.cosphi = std::cos(c[3]);
    mc.t566 = c[4]; mc.t556 = c[5]; mc.t555 = c[6];
    if (adopt_stream(h, (cudaStream_t)stream)) return 1;
    launch_map_apply(d_r, ld, n, mc, (cudaStream_t)stream);
    h->launches += 1;
    return check_launch(h, "k_map_apply(cavity)");
}

/* Scalars of the RF-cavity body map.  Specification: CavityTM.map4cav (transformations/cavity.py:29-128).
 * Written in terms of the normalised momenta eta = beta*gamma before and after the cavity:
 *   c1 = eta0/eta1, c2 = dgamma*beta0/eta1 (energy-deviation rescaling and RF curvature, cavity.py:81-84),
 *   kb = beta0*k, and the path-length terms T566, T556, T555 (cavity.py:63, :93-123).
 * delta_length < 0 or NaN means "whole cavity" (the reference's delta_length=None). */
int ocl_sc_cavity_coefficients(double v, double phi_deg, double freq, double E_GeV, double delta_length, double length,
                               double coef[7], int* mode, double* delta_e) {
    if (!coef || !mode || !delta_e) return 1;
    double m_e_eV, m_e_GeV, eps0, pi, c;
    constants(m_e_eV, m_e_GeV, eps0, pi, c);
    const bool partial = delta_length == delta_length && delta_length >= 0.0;
    const double z = partial ? delta_length : length;
    const double V = partial ? (length != 0.0 ? v * delta_length / length : v) : v;
    const double phi = phi_deg * pi / 180.0;
    const double sn = std::sin(phi), cs = std::cos(phi);
    // entrance: a beam "at rest" (E == 0) is treated as ultra-relativistic by the reference (beta0 = 1, gamma0 = 1e10)
    const double g0 = E_GeV != 0.0 ? E_GeV / m_e_GeV : 1e10;
    const double inv_g0sq = E_GeV != 0.0 ? 1.0 / (g0 * g0) : 0.0;
    const double b0 = E_GeV != 0.0 ? std::sqrt(1.0 - inv_g0sq) : 1.0;
    const double b0cube = b0 * b0 * b0;
    *delta_e = V * cs;
    const double E1 = E_GeV + *delta_e;
    for (int i = 0; i < 7; ++i) coef[i] = 0.0;
    coef[3] = phi;
    coef[4] = 1.5 * z * inv_g0sq / b0cube;                        // drift-like T566
    if (E1 <= 0.0) { *mode = 2; return 0; }                       // non-physical final energy: drift only
    *mode = 1;
    const double k = 2.0 * pi * freq / c;
    const double g1 = E1 / m_e_GeV;
    const double b1 = std::sqrt(1.0 - 1.0 / (g1 * g1));
    const double eta0 = b0 * g0, eta1 = b1 * g1;
    const double eta0cube = b0cube * g0 * g0 * g0, eta1cube = b1 * b1 * b1 * g1 * g1 * g1;
    const double dgam = V / m_e_GeV;                              // gamma gained at crest
    coef[0] = E_GeV * b0 / (E1 * b1);
    coef[1] = V * b0 / (E1 * b1);
    coef[2] = b0 * k;
    const double gap = g0 - g1;
    if (std::fabs(g1 - g0) < 1e-8 * std::fabs(g0)) {              // zero crossing: the general forms are 0/0
        if (std::fabs(cs) < 1e-3) {
            coef[5] = 1.5 * z * k * dgam / eta0cube;
            coef[6] = 0.5 * z * k * k * dgam * dgam / (eta0cube * g0);
        }
        return 0;
    }
    coef[4] = z * (eta0cube - eta1cube) / (2.0 * eta0 * eta1cube * gap);
    coef[5] = b0 * k * z * dgam * g0 * (eta1cube + b0 * (g0 - g1 * g1 * g1)) * sn / (eta1cube * gap * gap);
    const double curv = dgam * (2.0 * g0 * g1 * g1 * g1 * (b0 * b1 * b1 * b1 - 1.0) + g0 * g0 + 3.0 * g1 * g1 - 2.0)
                        / (eta1cube * gap * gap * gap) * sn * sn;
    const double lin = (g1 * g0 * (b1 * b0 - 1.0) + 1.0) / (eta1 * gap * gap) * cs;
    coef[6] = b0 * b0 * k * k * z * dgam / 2.0 * (curv - lin);
    return 0;
}

int ocl_sc_aperture_cut(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, const long long* d_ids,
                        long long n, int kind, int row, const double* params, double* d_r_out, long long ld_out,
                        double* d_q_out, long long* d_ids_out, long long* d_lost_out, long long* n_out, void* stream) {
    OCL_RANGE("ocl_sc_aperture_cut");
    if (!h || !params || !n_out || !d_r_out || !d_q_out) return 1;
    if (n < 0 || ld < n || ld_out < n) return fail(h, "ocl_sc_aperture_cut", "need 0 <= n <= ld, ld_out");
    if (kind != 0 && kind != 1) return fail(h, "ocl_sc_aperture_cut", "kind must be 0 (row against [lo, hi]) or 1 (ellipse)");
    if (kind == 0 && (row < 0 || row > 5)) return fail(h, "ocl_sc_aperture_cut", "row must be 0..5");
    *n_out = 0;
    if (n == 0) return 0;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    const long long tiles = (n + 1023) / 1024;
    if (tiles + 1 > h->cut_tiles) {
        cudaFree(h->cut_counts);
        h->cut_counts = nullptr; h->cut_tiles = 0;
        CU(h, cudaMalloc(&h->cut_counts, sizeof(int) * (size_t)(tiles + 1 + tiles / 4)));
        h->cut_tiles = tiles + 1 + tiles / 4;
    }
    if (!h->cut_n) CU(h, cudaHostAlloc(&h->cut_n, sizeof(long long), cudaHostAllocMapped));
    long long* d_n = nullptr;
    CU(h, cudaHostGetDevicePointer((void**)&d_n, h->cut_n, 0));
    CutSpec c;
    c.kind = kind; c.row = row; c.a = params[0]; c.b = params[1]; c.c = params[2]; c.d = params[3];
    launch_cut(d_r, ld, d_q, d_ids, n, c, h->cut_counts, d_n, d_r_out, ld_out, d_q_out, d_ids_out, d_lost_out, st);
    h->launches += 3;
    if (check_launch(h, "k_cut")) return 1;
    CU(h, cudaStreamSynchronize(st));                    // the new particle count shapes everything that follows
    *n_out = *h->cut_n;
    return 0;
}

int ocl_sc_beam_moments(ocl_sc_t* h, const double* d_r, long long ld, long long n, double* h_out, void* stream) {
    OCL_RANGE("ocl_sc_beam_moments");
    if (!h || !h_out) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_beam_moments", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    launch_moments(d_r, ld, n, nullptr, h->rs, h->moments, st);
    h->launches += 2;
    if (check_launch(h, "k_moments")) return 1;
    CU(h, cudaMemcpyAsync(h_out, h->moments, sizeof(double) * 18, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    return 0;
}

// The same two passes with the result left in DEVICE memory and no host synchronisation: a resident tracking loop
// records the moments of every step into its own device buffer and reads them back once at the end (track.py:482
// computes them per step but only returns them when tracking is over).  d_out[0..17] as ocl_sc_beam_moments;
// d_q != NULL adds d_out[18] = sum q.
int ocl_sc_beam_moments_device(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* d_q,
                               double* d_out, void* stream) {
    OCL_RANGE("ocl_sc_beam_moments_device");
    if (!h || !d_out) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_beam_moments_device", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    launch_moments(d_r, ld, n, d_q, h->rs, d_out, st);
    h->launches += 2;
    return check_launch(h, "k_moments");
}

/* ---- longitudinal space charge (LSC, sc.py:261-599) ---- */
static int lsc_async_error(ocl_sc* h, const char* who);
static int ensure_lsc(ocl_sc* h, int nb) {
    if (nb <= h->lsc_cap) return 0;
    int cap = 1024;
    while (cap < nb) cap *= 2;
    cudaFree(h->lw.bins); cudaFree(h->lw.cnt); cudaFree(h->lw.Z); cudaFree(h->lw.tw); cudaFree(h->lw.spread);
    h->lw.bins = nullptr; h->lw.cnt = nullptr; h->lw.Z = nullptr; h->lw.tw = nullptr; h->lw.spread = nullptr;
    h->lsc_cap = 0; h->lsc_tw_nb = 0;
    if (!h->lw.ticket) {
        CU(h, cudaMalloc(&h->lw.ticket, sizeof(unsigned int) * 2));
        CU(h, cudaMemset(h->lw.ticket, 0, sizeof(unsigned int) * 2));
        CU(h, cudaMalloc(&h->lw.stats, sizeof(double) * 32));
        CU(h, cudaMemset(h->lw.stats, 0, sizeof(double) * 32));
        h->lw.slice = h->lw.stats + 16;
        h->lw.sigma = h->lw.stats + 26;
        CU(h, cudaMalloc(&h->lw.dparams, sizeof(LscParams)));
        CU(h, cudaMalloc(&h->lw.dpack, sizeof(LscPack)));
        CU(h, cudaMemset(h->lw.dparams, 0, sizeof(LscParams)));
        CU(h, cudaHostAlloc(&h->lsc_err_host, sizeof(int), cudaHostAllocMapped));
        *h->lsc_err_host = 0;
        CU(h, cudaHostGetDevicePointer((void**)&h->lw.err, h->lsc_err_host, 0));
    }
    h->lw.part = h->rs.part;
    h->lw.max_blocks = h->rs.max_blocks;
    CU(h, cudaMalloc(&h->lw.bins, sizeof(unsigned long long) * cap));
    CU(h, cudaMemset(h->lw.bins, 0, sizeof(unsigned long long) * cap));
    {   // the spread histogram holds any nb <= cap (replicas shrink as nb grows)
        long long words = std::max<long long>(1 << 20, cap);    // nb * replicas <= 2^20 by construction
        CU(h, cudaMalloc(&h->lw.spread, sizeof(unsigned long long) * words));
        CU(h, cudaMemset(h->lw.spread, 0, sizeof(unsigned long long) * words));
        h->lsc_spread_cap = words;
    }
    CU(h, cudaMalloc(&h->lw.cnt, sizeof(double) * 5 * cap));
    h->lw.prof = h->lw.cnt + cap; h->lw.cur = h->lw.cnt + 2 * (size_t)cap; h->lw.W = h->lw.cnt + 3 * (size_t)cap;
    h->lw.A = h->lw.cnt + 4 * (size_t)cap;
    CU(h, cudaMalloc(&h->lw.Z, sizeof(double2) * cap));
    CU(h, cudaMalloc(&h->lw.tw, sizeof(double2) * 2 * cap));
    h->lsc_cap = cap;
    return 0;
}

static int lsc_params(ocl_sc* h, const double* p, LscParams& lp) {
    if (!p) return fail(h, "lsc", "params is NULL");
    lp.slice_min = p[0]; lp.slice_max = p[1]; lp.x_shift = p[2]; lp.y_shift = p[3];
    lp.a = p[4]; lp.ds = p[5]; lp.nb = (int)p[6]; lp.sigma_s = p[7]; lp.K = (int)p[8];
    lp.q = p[9]; lp.v = p[10]; lp.gamma = p[11]; lp.dz = p[12]; lp.und = p[13]; lp.pc_ref = p[14];
    lp.step_profile = p[15] != 0.0;
    double ntot = p[16] < 2.0 ? 2.0 : p[16];
    int bits = 0;
    while ((double)(1ull << bits) < ntot && bits < 62) ++bits;      // ceil(log2 n_total)
    lp.fx_shift = 62 - bits > 52 ? 52 : 62 - bits;
    if (!(lp.ds > 0.0) || lp.nb < 2 || lp.nb > (1 << 20)) return fail(h, "lsc", "bad grid (need ds > 0, 2 <= nb <= 2^20)");
    if (lp.K >= 0 && !(lp.sigma_s > 0.0)) return fail(h, "lsc", "smoothing taps need sigma_s > 0");
    return 0;
}

int ocl_sc_lsc_stats(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* d_q, double h_out[8],
                     void* stream) {
    OCL_RANGE("ocl_sc_lsc_stats");
    if (!h || !h_out) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_lsc_stats", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (ensure_lsc(h, 1)) return 1;
    if (lsc_async_error(h, "ocl_sc_lsc_stats")) return 1;
    launch_lsc_stats(d_r, ld, d_q, n, h->lw, st);
    h->launches += 1;
    if (check_launch(h, "k_lsc_stats")) return 1;
    double raw[9];
    CU(h, cudaMemcpyAsync(raw, h->lw.stats, sizeof raw, cudaMemcpyDeviceToHost, st));
    CU(h, cudaStreamSynchronize(st));
    const double cnt = raw[2], s1 = raw[3], s2 = raw[4], t0 = raw[8];
    h_out[0] = cnt;
    h_out[1] = t0 + s1 / cnt;                 // mean tau
    h_out[2] = s2 - s1 * s1 / cnt;            // sum (tau - mean)^2
    if (h_out[2] < 0.0) h_out[2] = 0.0;
    h_out[3] = -raw[1];                       // min tau
    h_out[4] = raw[0];                        // max tau
    h_out[5] = raw[5];                        // sum q
    h_out[6] = raw[6];                        // sum x
    h_out[7] = raw[7];                        // sum y
    return 0;
}

int ocl_sc_lsc_deposit(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* params, void* stream) {
    OCL_RANGE("ocl_sc_lsc_deposit");
    if (!h) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_lsc_deposit", "need 0 < n <= ld");
    LscParams lp;
    if (lsc_params(h, params, lp)) return 1;
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (ensure_lsc(h, lp.nb)) return 1;
    h->lsc_nb = lp.nb;
    h->lsc_async_pending = false;
    if (launch_lsc_deposit(d_r, ld, n, lp, h->lw, st))
        return fail(h, "ocl_sc_lsc_deposit", "grid too fine for this particle count (fewer than 24 fractional bits left)");
    h->launches += 2;
    return check_launch(h, "k_lsc_deposit");
}

int ocl_sc_lsc_solve_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* params, void* stream) {
    OCL_RANGE("ocl_sc_lsc_solve_kick");
    if (!h) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_lsc_solve_kick", "need 0 < n <= ld");
    LscParams lp;
    if (lsc_params(h, params, lp)) return 1;
    if (lp.nb != h->lsc_nb) return fail(h, "ocl_sc_lsc_solve_kick", "grid differs from the deposited one");
    ENTER_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    if (h->lsc_tw_nb != lp.nb) {
        launch_lsc_twiddles(lp.nb, h->lw, st);
        h->lsc_tw_nb = lp.nb;
        h->launches += 1;
    }
    if (launch_lsc_solve(lp, h->lw, st)) return fail(h, "ocl_sc_lsc_solve_kick", "too many smoothing taps (K > 2559)");
    launch_lsc_kick(d_r, ld, n, lp, h->lw, st);
    h->launches += 5;
    return check_launch(h, "k_lsc_kick");
}

int ocl_sc_lsc_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* params, void* stream) {
    OCL_RANGE("ocl_sc_lsc_kick");
    if (ocl_sc_lsc_deposit(h, d_r, ld, n, params, stream)) return 1;
    return ocl_sc_lsc_solve_kick(h, d_r, ld, n, params, stream);
}

static int lsc_async_error(ocl_sc* h, const char* who) {
    if (h->lsc_err_host && *h->lsc_err_host) {
        const int e = *h->lsc_err_host;
        *h->lsc_err_host = 0;
        return fail(h, who, e == 2 ? "an earlier asynchronous LSC kick was skipped: grid too fine for the particle count"
                                   : "an earlier asynchronous LSC kick was skipped: its grid exceeds the buffer "
                                     "capacity (use the synchronous form, ocl_sc_lsc_stats + ocl_sc_lsc_kick)");
    }
    return 0;
}

int ocl_sc_lsc_kick_async(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* d_q, const double* hostp,
                          void* stream) {
    OCL_RANGE("ocl_sc_lsc_kick_async");
    if (!h || !hostp) return 1;
    if (n <= 0 || ld < n) return fail(h, "ocl_sc_lsc_kick_async", "need 0 < n <= ld");
    ENTER_DEVICE(h);
    if (ensure_lsc(h, kLscAsyncCap)) return 1;
    if (lsc_async_error(h, "ocl_sc_lsc_kick_async")) return 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (adopt_stream(h, st)) return 1;
    LscHost hp;
    hp.gamma = hostp[0]; hp.v = hostp[1]; hp.pc_ref = hostp[2]; hp.dz = hostp[3]; hp.und = hostp[4];
    hp.bound_lo = hostp[5]; hp.bound_hi = hostp[6]; hp.smooth_param = hostp[7];
    hp.step_profile = hostp[8] != 0.0;
    double ntot = hostp[9] < 2.0 ? (n < 2 ? 2.0 : (double)n) : hostp[9];
    int bits = 0;
    while ((double)(1ull << bits) < ntot && bits < 62) ++bits;
    hp.fx_shift = 62 - bits > 52 ? 52 : 62 - bits;
    hp.cap_nb = kLscAsyncCap;
    hp.warps = hp.iters = 0;
    launch_lsc_kick_async(d_r, ld, d_q, n, hp, h->lw, st);
    h->launches += 10;
    h->lsc_async_pending = true;
    h->lsc_tw_nb = 0;                          // the twiddle table now belongs to a device-defined grid
    return check_launch(h, "ocl_sc_lsc_kick_async");
}

int ocl_sc_lsc_async_status(ocl_sc_t* h, int synchronise, int* status) {
    if (!h || !status) return 1;
    *status = 0;
    if (!h->lsc_err_host) return 0;                  // no asynchronous kick was ever issued
    ENTER_DEVICE(h);
    if (synchronise && sync_last(h)) return 1;
    *status = *h->lsc_err_host;
    *h->lsc_err_host = 0;
    return 0;
}

int ocl_sc_lsc_last_params(ocl_sc_t* h, double out[17]) {
    if (!h || !out) return 1;
    if (!h->lsc_async_pending) return fail(h, "ocl_sc_lsc_last_params", "no asynchronous LSC kick recorded");
    ENTER_DEVICE(h);
    if (sync_last(h)) return 1;
    if (lsc_async_error(h, "ocl_sc_lsc_last_params")) return 1;
    LscParams lp;
    CU(h, cudaMemcpy(&lp, h->lw.dparams, sizeof lp, cudaMemcpyDeviceToHost));
    const double v[17] = {lp.slice_min, lp.slice_max, lp.x_shift, lp.y_shift, lp.a, lp.ds, (double)lp.nb, lp.sigma_s,
                          (double)lp.K, lp.q, lp.v, lp.gamma, lp.dz, lp.und, lp.pc_ref, (double)lp.step_profile, 0.0};
    for (int i = 0; i < 17; ++i) out[i] = v[i];
    h->lsc_nb = lp.nb;
    return 0;
}

int ocl_sc_lsc_get_profile(ocl_sc_t* h, int nb, double* h_current, double* h_wake, double* h_sigma) {
    if (!h) return 1;
    if (nb <= 0 || nb != h->lsc_nb) return fail(h, "ocl_sc_lsc_get_profile", "nb differs from the last LSC kick");
    ENTER_DEVICE(h);
    if (sync_last(h)) return 1;
    if (h_current) CU(h, cudaMemcpy(h_current, h->lw.cur, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_wake) CU(h, cudaMemcpy(h_wake, h->lw.W, sizeof(double) * nb, cudaMemcpyDeviceToHost));
    if (h_sigma) CU(h, cudaMemcpy(h_sigma, h->lw.sigma, sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int ocl_sc_enable_timers(ocl_sc_t* h, int enable) {
    if (!h) return 1;
    h->timers = enable != 0;
    h->ev_valid = false;
    return 0;
}

int ocl_sc_get_timers(ocl_sc_t* h, double out[8]) {
    if (!h) return 1;
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    if (!h->timers || !h->ev_valid) return fail(h, "ocl_sc_get_timers", "no timed kick recorded");
    ENTER_DEVICE(h);
    CU(h, cudaEventSynchronize(h->ev[T_KICK]));
    float ms;
    for (int s = T_MOM; s <= T_KICK; ++s) {
        CU(h, cudaEventElapsedTime(&ms, h->ev[s - 1], h->ev[s]));
        out[s - 1] = ms;
    }
    CU(h, cudaEventElapsedTime(&ms, h->ev[T_BEGIN], h->ev[T_KICK]));
    out[6] = ms;
    return 0;
}

long long ocl_sc_launch_count(const ocl_sc_t* h) { return h ? h->launches : 0; }

}  // extern "C"
