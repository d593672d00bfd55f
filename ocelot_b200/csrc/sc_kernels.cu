// sc_kernels.cu -- sm_100a kernels of the 3D space-charge kick.
//
// One SpaceCharge.apply (ocelot/cpbd/sc.py:208-251) is four sweeps over the
// particles separated by the three global dependencies of the algorithm
// (mean momentum -> frame; extents/centroid -> mesh; rho -> potential), plus
// the grid work in between:
//
//   k_momentum     sweep 1  rows x',y',delta            -> sums[4]; tail: exchange, frame -> Geo
//   k_extent       sweep 2  6 rows + q                  -> emax[6], esum[4]; tail: exchange, mesh -> Geo
//   k_deposit      sweep 3  6 rows + q                  -> rho (NGP, fp64 RED)
//   k_green_table / k_green_mirror                      -> K on the padded grid
//   cuFFT D2Z x2, k_multiply, cuFFT Z2D                 -> convolution
//   k_crop_phi, k_field                                 -> phi, Ex/Ey/Ez
//   k_gather_kick  sweep 4  6 rows in, 6 rows out       -> kicked particles
//
// No kernel needs a host round trip: the scalar state (frame, mesh) is
// re-derived on the device by every block from the tiny reduced buffers, which
// is also what lets a multi-GPU caller all-reduce those buffers in between.
#include <cstdint>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "sc_kernels.h"

namespace ocl {

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("OCL_SC_PDL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on != 0;
}

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

int particle_grid(long long n, int max_blocks) {
    long long b = (n + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

constexpr int kPipeDepth = 3;

// ---------------------------------------------------------------------------
// Fused exchange over NVLink peer memory (replaces an NCCL all-reduce / all-gather of a handful of
// doubles): lane w stores this rank's values into rank w's mailbox, fences, raises its epoch flag
// there; then waits for rank w's flag in the local mailbox; lane 0 folds the W slots in rank order
// (bit-identical result on every rank) into the handle's reduced buffers.
//   which 0: momentum {sum px, py, pz, count}        -> SUM
//   which 1: extents  {max x6} MAX, {sum q x3, q} SUM
//   which 2: barrier only
// Slots are single-buffered: a rank can only push exchange e+1 after it has consumed everyone's
// exchange e of the *other* kind, which in turn requires everyone to have consumed this kind's e.
// Runs in ONE warp (all 32 lanes): either the warp that finishes a sweep's grid reduction (no extra
// launch on the critical path) or the stand-alone k_mailbox_exchange.
// ---------------------------------------------------------------------------
// Flags cross NVLink with release / acquire accesses instead of plain accesses bracketed by full system fences:
// a membar.sys per waiting warp is what made a many-block wait expensive (2 x B200: 17 us per kick when every block
// of the reduction kernel fenced after its poll).
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// spin with relaxed loads, then one acquire load of the flag that was seen raised
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long epoch, bool backoff) {
    const long long t0 = clock64();
    while (ld_relaxed_sys(p) < epoch) {
        if (clock64() - t0 > 8000000000LL) return false;            // ~4 s: a peer is gone
        if (backoff) __nanosleep(64);      // many blocks may poll the same line: leave the L2 slice room for the remote write
    }
    return ld_acquire_sys(p) >= epoch;
}

// vio (optional): lane 0 holds this rank's values in registers on entry (else they are read from the handle's
// buffers) and receives the folded values on exit, so the caller need not read them back from memory.
__device__ __forceinline__ void mailbox_exchange_warp(const Mailbox& mb, int which, const ReduceState& rs,
                                                      int* __restrict__ err_flag, double* vio = nullptr) {
    const int lane = threadIdx.x & 31;
    const int nv = which == 0 ? 4 : (which == 1 ? 10 : 0);
    const int vbase = which == 0 ? 0 : 32;
    const int fbase = 128 + 8 * which;
    const double* local_vals = which == 0 ? rs.sums : rs.emax;          // emax[6] and esum[4] are contiguous
    const unsigned long long epoch = mb.epoch[which] + 1;
    __syncwarp();
    double mine_v[10];
    for (int k = 0; k < nv; ++k) {
        const double x = vio ? vio[k] : 0.0;                              // meaningful in lane 0 only
        mine_v[k] = vio ? __shfl_sync(0xffffffffu, x, 0) : __ldcg(local_vals + k);
    }
    if (lane < mb.world) {
        double* dst = mb.peer[lane] + vbase + mb.rank * nv;
        for (int k = 0; k < nv; ++k) dst[k] = mine_v[k];
        // the same thread wrote the values: a release store of the flag orders them before it
        st_release_sys(reinterpret_cast<unsigned long long*>(mb.peer[lane] + fbase) + mb.rank, epoch);
        // wait for rank `lane` to have delivered its values here
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(mb.peer[mb.rank] + fbase) + lane;
        if (!wait_flag(mine, epoch, false)) atomicExch(err_flag, 1 + which);
    }
    __syncwarp();
    if (lane == 0) {
        if (which < 2) {
            const volatile double* box = mb.peer[mb.rank] + vbase;
            double v[10];
            for (int k = 0; k < nv; ++k) v[k] = box[k];
            for (int w = 1; w < mb.world; ++w)
                for (int k = 0; k < nv; ++k) {
                    const double x = box[w * nv + k];
                    v[k] = (which == 1 && k < 6) ? fmax(v[k], x) : v[k] + x;
                }
            if (which == 0) {
                for (int k = 0; k < 4; ++k) rs.sums[k] = v[k];
            } else {
                for (int k = 0; k < 6; ++k) rs.emax[k] = v[k];
                for (int k = 0; k < 4; ++k) rs.esum[k] = v[6 + k];
            }
            if (vio) for (int k = 0; k < nv; ++k) vio[k] = v[k];
        }
        mb.epoch[which] = epoch;
        __threadfence();
    }
    __syncwarp();
}

// frame of the kick from the (globally) reduced momentum sums -> rs.geo->f        (one thread)
// sums_reg: the four reduced values in registers, or nullptr = read the handle's buffer
__device__ __forceinline__ void finish_momentum(const ReduceState& rs, const RefParams& rp, const double* sums_reg = nullptr) {
    Frame f;
    double sums[4];
    for (int k = 0; k < 4; ++k) sums[k] = sums_reg ? sums_reg[k] : __ldcg(rs.sums + k);
    derive_frame(sums, rp.m_e_eV, f);
    rs.geo->f = f;
}

// mesh of the kick from the (globally) reduced extents -> rs.geo->m, geometry tap.  Called by a whole warp:
// lanes 0..2 each derive one axis (the IEEE divisions of an axis are the long pole of this tail; same operations as
// derive_mesh), lane 0 holds the ten reduced values in e_reg (or nullptr = read the handle's buffer).
__device__ __forceinline__ void finish_extent_warp(const ReduceState& rs, const MeshDims& md, const Draws& dr, const Frame& f,
                                                   const double* e_reg = nullptr) {
    const int lane = threadIdx.x & 31;
    double e[10];
    for (int k = 0; k < 10; ++k) {
        const double x = e_reg ? e_reg[k] : 0.0;
        e[k] = e_reg ? __shfl_sync(0xffffffffu, x, 0) : __ldcg(rs.emax + k);      // emax[6] and esum[4] are contiguous
    }
    double* g = rs.geom;
    if (lane < 3) {
        const int c = lane;
        const int n = c == 0 ? md.nx : (c == 1 ? md.ny : md.nz);
        const double lo = -e[3 + c];
        double extent = e[c] - lo;                                     // sc.py:173
        if (dr.scale > 0.0) extent = extent * dr.scale;                // :175
        const double h = extent / (double)(n - 3);                     // :179
        const double xmin = lo / h;                                    // :181 (min commutes with /h)
        const double xmid = (e[6 + c] / h) / e[9];                     // :182
        double off = floor(xmin - xmid) + xmid;                        // :183
        if (dr.scale > 0.0) off = off + dr.shift;                      // :185
        Mesh* m = &rs.geo->m;
        m->steps[c] = h; m->inv_steps[c] = 1.0 / h; m->xoff[c] = off; m->n[c] = n;
        g[12 + c] = h; g[15 + c] = off;
        for (int j = 0; j < 3; ++j) g[c * 3 + j] = f.T[c][j];
    } else if (lane == 3) {
        rs.geo->m.sumq = e[9];
        g[9] = f.pav; g[10] = f.gamma0; g[11] = f.beta0;
        g[18] = e[9]; g[19] = __ldcg(rs.sums + 3);
    }
}

// ---------------------------------------------------------------------------
// sweep 1: mean momentum (sc.py:221,224).  The block that finishes the reduction also runs the
// cross-rank exchange (sharded kick) and derives the frame, unless the caller defers that
// (NCCL fallback: all-reduce of rs.sums, then k_finish).
// ---------------------------------------------------------------------------
// publish != nullptr (whole-kick graph): this kernel is the graph's parameter node -- it receives the kick's
// scalars by value (refreshed per launch with cudaGraphExecKernelNodeSetParams) and stores them for the kernels
// downstream, which read the device block; no separate one-block parameter kernel in front of the kick.
template <bool TMA>
__global__ void __launch_bounds__(kThreads, 4) k_momentum(const double* __restrict__ r, long long ld, long long n,
                                                         KP kp, ReduceState rs, Mailbox mb, int* mb_err,
                                                         KickParams* publish) {
    pdl_enter();
    if (publish && blockIdx.x == 0 && threadIdx.x == 0) *publish = kp.v;
    const RefParams rp = kp_ref(kp);
    __shared__ double sh[3 * kWarps];
    __shared__ __align__(128) double pipe[kPipeDepth * 3 * kThreads];
    __shared__ unsigned long long bars[kPipeDepth];
    double v[3] = {0.0, 0.0, 0.0};
    const double* const base[3] = {r + ld, r + 3 * ld, r + 5 * ld};
    auto body = [&](int, const double (&w)[3]) {
        double gam;
        const double pzr = mad_pz_rel(rp, w[0], w[1], w[2], gam);
        v[0] += w[0]; v[1] += w[1]; v[2] += pzr;       // p = (x', y', pz_rel) * pc: scaled once at the end
    };
    if constexpr (TMA) bulk_sweep<3, kPipeDepth>(base, (int)n, pipe, bars, body);
    else pipelined_sweep<3, kPipeDepth>(base, (int)n, pipe, body);
    pdl_trigger();
    if (!grid_reduce<3, 0>(v, rs.part, rs.ticket + 0, sh)) return;
    if (threadIdx.x >= 32) return;                      // the finishing block's first warp carries on
    if (threadIdx.x == 0) {
        rs.sums[0] = v[0] * rp.pc; rs.sums[1] = v[1] * rp.pc; rs.sums[2] = v[2] * rp.pc;
        rs.sums[3] = (double)n;
        __threadfence();
    }
    if (rs.defer) return;
    double sums[4] = {v[0] * rp.pc, v[1] * rp.pc, v[2] * rp.pc, (double)n};     // valid in thread 0
    if (mb.world > 1) mailbox_exchange_warp(mb, 0, rs, mb_err, sums);
    if (threadIdx.x == 0) finish_momentum(rs, rp, sums);
}

// ---------------------------------------------------------------------------
// sweep 2: extents and charge centroid in the bunch frame (sc.py:172-173,181-182).  Every thread
// remembers which particle gave each of its six extrema; the finishing block re-evaluates those
// (at most six) particles in the reference's exact operation order before the mesh is derived.
// ---------------------------------------------------------------------------
template <bool TMA>
__global__ void __launch_bounds__(kThreads, 3) k_extent(const double* __restrict__ r, long long ld,
                                                       const double* __restrict__ q, long long n, KP kp,
                                                       ReduceState rs, MeshDims md, Mailbox mb, int* mb_err) {
    pdl_enter();
    const RefParams rp = kp_ref(kp);
    __shared__ double sh[10 * kWarps];
    __shared__ __align__(128) double pipe[kPipeDepth * 7 * kThreads];
    __shared__ unsigned long long bars[kPipeDepth];
    __shared__ Frame sf;
    __shared__ double shv[6];
    __shared__ int shi[6];
    {
        constexpr int W = (int)(sizeof(Frame) / sizeof(double));
        if (threadIdx.x < W) reinterpret_cast<double*>(&sf)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(&rs.geo->f) + threadIdx.x);
        __syncthreads();
    }
    const Frame f = sf;
    double v[10] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0, 0.0, 0.0, 0.0};
    int ix[6] = {-1, -1, -1, -1, -1, -1};
    const double* const base[7] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld, q};
    auto body = [&](int i, const double (&w)[7]) {
        const Cart c = mad_to_cart(rp, w[0], w[1], w[2], w[3], w[4], w[5]);
        double a, b, g;
        rotate_stretch(f, c.x, c.y, c.z, a, b, g);
        const double qi = w[6];
#ifndef OCL_EXT_NOIDX
        if (a > v[0]) { v[0] = a; ix[0] = i; }
        if (b > v[1]) { v[1] = b; ix[1] = i; }
        if (g > v[2]) { v[2] = g; ix[2] = i; }
        if (-a > v[3]) { v[3] = -a; ix[3] = i; }
        if (-b > v[4]) { v[4] = -b; ix[4] = i; }
        if (-g > v[5]) { v[5] = -g; ix[5] = i; }
#else
        v[0] = fmax(v[0], a); v[1] = fmax(v[1], b); v[2] = fmax(v[2], g);
        v[3] = fmax(v[3], -a); v[4] = fmax(v[4], -b); v[5] = fmax(v[5], -g);
#endif
        v[6] += qi * a; v[7] += qi * b; v[8] += qi * g; v[9] += qi;
    };
    if constexpr (TMA) bulk_sweep<7, kPipeDepth>(base, (int)n, pipe, bars, body);
    else pipelined_sweep<7, kPipeDepth>(base, (int)n, pipe, body);
    pdl_trigger();
    if (!grid_reduce_extent(v, ix, rs.part, rs.ticket + 1, sh, shv, shi)) return;
    if (threadIdx.x >= 32) return;
    // lanes 0..5: the particle behind extremum k, once more, exactly as the reference computes it
    const int lane = threadIdx.x;
    double exact = -INFINITY;
#ifndef OCL_EXT_NOEXACT
    int who = -1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const int w = __shfl_sync(0xffffffffu, ix[k], 0);      // ix[] is valid in thread 0
        if (lane == k) who = w;
    }
    if (lane < 6 && who >= 0) {                                // six lanes, six particles, in parallel
        double p[3];
        exact_frame_position(rp, f, r[who], r[ld + who], r[2 * ld + who], r[3 * ld + who], r[4 * ld + who],
                             r[5 * ld + who], p[0], p[1], p[2]);
        exact = lane == 0 ? p[0] : lane == 1 ? p[1] : lane == 2 ? p[2] : lane == 3 ? -p[0] : lane == 4 ? -p[1] : -p[2];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double e = __shfl_sync(0xffffffffu, exact, k);
        if (lane == 0 && e > -INFINITY) v[k] = e;
    }
#endif
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) rs.emax[k] = v[k];
#pragma unroll
        for (int k = 0; k < 4; ++k) rs.esum[k] = v[6 + k];
        __threadfence();
    }
    if (rs.defer) return;
    if (mb.world > 1) mailbox_exchange_warp(mb, 1, rs, mb_err, v);
    finish_extent_warp(rs, md, kp_draws(kp), f, v);
}

// deferred tails (NCCL fallback of a sharded kick: the host all-reduces rs.sums / rs.emax,esum in between)
__global__ void k_finish(int which, KP kp, ReduceState rs, MeshDims md) {
    pdl_enter();
    if (blockIdx.x != 0 || threadIdx.x >= 32) return;
    if (which == 0) {
        if (threadIdx.x == 0) finish_momentum(rs, kp_ref(kp));
    } else {
        Frame f;
        constexpr int W = (int)(sizeof(Frame) / sizeof(double));
        for (int k = 0; k < W; ++k) reinterpret_cast<double*>(&f)[k] = __ldcg(reinterpret_cast<const double*>(&rs.geo->f) + k);
        finish_extent_warp(rs, md, kp_draws(kp), f);
    }
}

__device__ __forceinline__ void mailbox_signal(const Mailbox& mb, int fbase, unsigned long long epoch) {
    const int lane = threadIdx.x & 31;
    if (lane < mb.world) st_release_sys(reinterpret_cast<unsigned long long*>(mb.peer[lane] + fbase) + mb.rank, epoch);
}
__device__ __forceinline__ void mailbox_wait(const Mailbox& mb, int fbase, unsigned long long epoch, int* err_flag) {
    const int lane = threadIdx.x & 31;
    if (lane < mb.world) {
        const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(mb.peer[mb.rank] + fbase) + lane;
        if (!wait_flag(mine, epoch, true)) atomicExch(err_flag, 3);
    }
    __syncwarp();
}
constexpr int kFlagRhoReady = 128 + 8 * 2;      // doubles [144,152): "deposit complete" epochs
constexpr int kFlagRhoDone = 128 + 8 * 3;       // doubles [152,160): "slice reduced" epochs
// ---------------------------------------------------------------------------
// sweep 3: nearest-grid-point deposit (sc.py:186-193)
// ---------------------------------------------------------------------------
template <bool TMA>
__global__ void __launch_bounds__(kThreads, 3) k_deposit(const double* __restrict__ r, long long ld,
                                                        const double* __restrict__ q, long long n, KP kp,
                                                        ReduceState rs, MeshDims md, double* __restrict__ rho) {
    pdl_enter();
    const RefParams rp = kp_ref(kp);
    __shared__ __align__(128) double pipe[kPipeDepth * 7 * kThreads];
    __shared__ unsigned long long bars[kPipeDepth];
    __shared__ Geo sg;
    load_geo(rs.geo, &sg);
    const Frame f = sg.f;
    const Mesh m = sg.m;
    const double* const base[7] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld, q};
    auto body = [&](int, const double (&w)[7]) {
        const Cart c = mad_to_cart(rp, w[0], w[1], w[2], w[3], w[4], w[5]);
        double a, b, g, g0, g1, g2;
        rotate_stretch(f, c.x, c.y, c.z, a, b, g);
        to_grid(m, a, b, g, g0, g1, g2);
        const int c0 = (int)floor(g0) + 1, c1 = (int)floor(g1) + 1, c2 = (int)floor(g2) + 1;   // sc.py:191
        if ((unsigned)c0 < (unsigned)md.nx && (unsigned)c1 < (unsigned)md.ny && (unsigned)c2 < (unsigned)md.nz)
            atomicAdd(rho + ((size_t)c0 * md.ny + c1) * md.nz + c2, w[6]);                     // sc.py:192-193
    };
    if constexpr (TMA) bulk_sweep<7, kPipeDepth>(base, (int)n, pipe, bars, body);
    else pipelined_sweep<7, kPipeDepth>(base, (int)n, pipe, body);
}

// ---------------------------------------------------------------------------
// Ordered momentum sum (debugging mode, with the ordered deposit): np.mean(xp[3:6], axis=1) (sc.py:224) bit for
// bit.  numpy adds a contiguous row with its pairwise scheme: ranges longer than 128 are split at n/2 rounded down
// to a multiple of 8, ranges of 8..128 elements are summed with eight interleaved accumulators r[j] += a[i + j],
// combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)), then the last n % 8 elements are added one by one
// (checked against np.add.reduce bit for bit for n = 1 .. 10^6, round 2).  The host lists the leaves and the
// combination tree for a given n once (PairwisePlan, sc_abi.cu); here
//   k_momentum_exact_leaves  one warp per leaf: the momenta of its <= 128 particles exactly as the reference rounds
//                            them (exact_momentum), the leaf's three sums in numpy's order
//   k_pairwise_combine       one block walks the tree level by level -> rs.sums
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_momentum_exact_leaves(const double* __restrict__ r, long long ld, KP kp,
                                                              const uint2* __restrict__ leaves, int nleaves,
                                                              double* __restrict__ V) {
    pdl_enter();
    const RefParams rp = kp_ref(kp);
    __shared__ double sh[8][3][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int leaf = blockIdx.x * 8 + warp;
    if (leaf >= nleaves) return;
    const uint2 lf = leaves[leaf];
    const long long off = lf.x;
    const int len = (int)lf.y;
    for (int i = lane; i < len; i += 32) {
        const long long p = off + i;
        double px, py, pz;
        exact_momentum(rp, r[ld + p], r[3 * ld + p], r[5 * ld + p], px, py, pz);
        sh[warp][0][i] = px; sh[warp][1][i] = py; sh[warp][2][i] = pz;
    }
    __syncwarp();
    if (lane < 24) {                                    // lane = component * 8 + accumulator
        const unsigned mask = 0x00ffffffu;
        const int comp = lane >> 3, j = lane & 7;
        const double* a = sh[warp][comp];
        double res = 0.0;
        if (len < 8) {                                  // only a bunch of fewer than 8 particles: 0 + a0 + a1 + ...
            if (j == 0) for (int i = 0; i < len; ++i) res = __dadd_rn(res, a[i]);
        } else {
            double rj = a[j];
            const int nb = len - (len % 8);
            for (int i = 8; i < nb; i += 8) rj = __dadd_rn(rj, a[i + j]);
            double t = __dadd_rn(rj, __shfl_down_sync(mask, rj, 1, 8));     // j even: r_j + r_j+1
            t = __dadd_rn(t, __shfl_down_sync(mask, t, 2, 8));              // j % 4 == 0: (r_j + r_j+1) + (r_j+2 + r_j+3)
            t = __dadd_rn(t, __shfl_down_sync(mask, t, 4, 8));              // j == 0: the eight
            res = t;
            if (j == 0) for (int i = nb; i < len; ++i) res = __dadd_rn(res, a[i]);
        }
        if (j == 0) V[(size_t)leaf * 3 + comp] = res;
    }
}

// nodes[k] = (left, right) value indices of internal node nleaves + k; nodes are sorted by height and
// level_start[l] .. level_start[l + 1] are the nodes of height l + 1.  The root is the last value.
__global__ void __launch_bounds__(1024) k_pairwise_combine(double* __restrict__ V, const uint2* __restrict__ nodes,
                                                          const int* __restrict__ level_start, int nlevels, int nleaves,
                                                          long long n, double* __restrict__ sums) {
    pdl_enter();
    for (int l = 0; l < nlevels; ++l) {
        const int lo = level_start[l], hi = level_start[l + 1];
        for (int k = lo + (int)threadIdx.x; k < hi; k += (int)blockDim.x) {
            const uint2 c = nodes[k];
            double* out = V + (size_t)(nleaves + k) * 3;
#pragma unroll
            for (int m = 0; m < 3; ++m) out[m] = __dadd_rn(V[(size_t)c.x * 3 + m], V[(size_t)c.y * 3 + m]);
        }
        __syncthreads();
    }
    if (threadIdx.x < 3) {
        const size_t root = (size_t)nleaves + (size_t)level_start[nlevels] - 1;
        sums[threadIdx.x] = V[root * 3 + threadIdx.x];
    }
    if (threadIdx.x == 3) sums[3] = (double)n;
}

// ---------------------------------------------------------------------------
// Ordered deposit (debugging mode, SURVEY section 8e "Determinism"): the same cells, but every cell's charges are
// added one after the other in ascending particle order starting from 0.0 -- the order np.bincount uses
// (sc.py:193) -- so rho is bit-identical from run to run and, for identical cell indices, to the reference's grid.
//   k_cell_index   sweep 3': 6 rows -> (cell, particle) pairs; particles outside the mesh get the key 0xffffffff
//   stable radix sort of the pairs by cell (cub::DeviceRadixSort: equal keys keep their particle order)
//   k_ordered_sum  the thread at the head of each run of equal cells adds that run sequentially
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 3) k_cell_index(const double* __restrict__ r, long long ld, long long n,
                                                           KP kp, ReduceState rs, MeshDims md,
                                                           unsigned* __restrict__ key, unsigned* __restrict__ val) {
    pdl_enter();
    const RefParams rp = kp_ref(kp);
    __shared__ __align__(128) double pipe[kPipeDepth * 6 * kThreads];
    __shared__ Geo sg;
    load_geo(rs.geo, &sg);
    const Frame f = sg.f;
    const Mesh m = sg.m;
    const double* const base[6] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld};
    auto body = [&](int i, const double (&w)[6]) {
        const Cart c = mad_to_cart(rp, w[0], w[1], w[2], w[3], w[4], w[5]);
        double a, b, g, g0, g1, g2;
        rotate_stretch(f, c.x, c.y, c.z, a, b, g);
        to_grid(m, a, b, g, g0, g1, g2);
        const int c0 = (int)floor(g0) + 1, c1 = (int)floor(g1) + 1, c2 = (int)floor(g2) + 1;   // sc.py:191
        const bool in = (unsigned)c0 < (unsigned)md.nx && (unsigned)c1 < (unsigned)md.ny && (unsigned)c2 < (unsigned)md.nz;
        key[i] = in ? (unsigned)(((size_t)c0 * md.ny + c1) * md.nz + c2) : 0xffffffffu;
        val[i] = (unsigned)i;
    };
    pipelined_sweep<6, kPipeDepth>(base, (int)n, pipe, body);
}

__global__ void __launch_bounds__(kThreads) k_ordered_sum(const unsigned* __restrict__ key, const unsigned* __restrict__ val,
                                                         const double* __restrict__ q, long long n, unsigned cells,
                                                         double* __restrict__ rho) {
    pdl_enter();
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const unsigned c = key[j];
        if (c >= cells || (j > 0 && key[j - 1] == c)) continue;          // not the head of a run
        double s = 0.0;
        for (long long k = j; k < n && key[k] == c; ++k) s = __dadd_rn(s, q[val[k]]);
        rho[c] = s;
    }
}

// ---------------------------------------------------------------------------
// integrated Green's function (sc.py:109-133)
// ---------------------------------------------------------------------------
// every product and sum is an explicitly rounded operation, in the order of
// sc.py:124-126, so the compiler cannot contract them: the 8-corner difference
// below amplifies any rounding difference by up to (r/h)^3.
__device__ __forceinline__ double igf_antiderivative(double x, double y, double z) {
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), zz = __dmul_rn(z, z);
    const double rr = sqrt(__dadd_rn(__dadd_rn(xx, yy), zz));
    const double yz = __dmul_rn(y, z), zx = __dmul_rn(z, x), xy = __dmul_rn(x, y);
    double g = __dmul_rn(__dmul_rn(-xx, 0.5), atan(yz / __dmul_rn(x, rr)));
    g = __dadd_rn(g, __dmul_rn(yz, log(__dadd_rn(x, rr))));
    g = __dsub_rn(g, __dmul_rn(__dmul_rn(yy, 0.5), atan(zx / __dmul_rn(y, rr))));
    g = __dadd_rn(g, __dmul_rn(zx, log(__dadd_rn(y, rr))));
    g = __dsub_rn(g, __dmul_rn(__dmul_rn(zz, 0.5), atan(xy / __dmul_rn(z, rr))));
    g = __dadd_rn(g, __dmul_rn(xy, log(__dadd_rn(z, rr))));
    return g;
}

struct StepSrc {
    int given;       // 1: use h[]; 0: derive from the reduced extents
    double h[3];
};

__device__ __forceinline__ void resolve_steps(const StepSrc& src, const ReduceState& rs, double* h /* shared [3] */) {
    if (threadIdx.x < 3) h[threadIdx.x] = src.given ? src.h[threadIdx.x] : __ldcg(rs.geo->m.steps + threadIdx.x);
    __syncthreads();
}

// antiderivative on the (n+1)^3 half-offset points (sc.py:116-126)
__global__ void __launch_bounds__(kThreads) k_green_table(StepSrc src, ReduceState rs, MeshDims md,
                                                         double* __restrict__ gtab, double* __restrict__ h3) {
    pdl_enter();
    __shared__ double h[3];
    resolve_steps(src, rs, h);
    if (blockIdx.x == 0 && threadIdx.x == 0) { h3[0] = h[0]; h3[1] = h[1]; h3[2] = h[2]; }   // for the solver
    const int gx = md.nx + 1, gy = md.ny + 1, gz = md.nz + 1;
    const long long total = (long long)gx * gy * gz;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        int k = (int)(t % gz);
        long long u = t / gz;
        int j = (int)(u % gy);
        int i = (int)(u / gy);
        double x = __dsub_rn(__dmul_rn(h[0], (double)i), h[0] / 2);
        double y = __dsub_rn(__dmul_rn(h[1], (double)j), h[1] / 2);
        double z = __dsub_rn(__dmul_rn(h[2], (double)k), h[2] / 2);
        gtab[t] = igf_antiderivative(x, y, z);
    }
}

// 8-corner difference (sc.py:128-131)
__device__ __forceinline__ double green_entry(const double* __restrict__ G, int gy, int gz, int i, int j, int k) {
    const size_t sx = (size_t)gy * gz, sy = gz;
    const double* lo = G + (size_t)i * sx + (size_t)j * sy + k;
    double v = __ldg(lo + sx + sy + 1) - __ldg(lo + sy + 1);
    v = v - __ldg(lo + sx + 1);
    v = v + __ldg(lo + 1);
    v = v - __ldg(lo + sx + sy);
    v = v + __ldg(lo + sy);
    v = v + __ldg(lo + sx);
    v = v - __ldg(lo);
    return v;
}

// K on the padded periodic grid: K[d mod M] = K1[|d|] (sc.py:145-149).  One
// thread per K1 entry writes its (up to 8) mirror images; the gap planes
// n <= d <= M-n were zeroed by a memset.
__global__ void __launch_bounds__(kThreads) k_green_mirror(const double* __restrict__ gtab, MeshDims md,
                                                          double* __restrict__ kpad) {
    const long long total = (long long)md.nx * md.ny * md.nz;
    const int gy = md.ny + 1, gz = md.nz + 1;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        int c = (int)(t % md.nz);
        long long u = t / md.nz;
        int b = (int)(u % md.ny);
        int a = (int)(u / md.ny);
        const double v = green_entry(gtab, gy, gz, a, b, c);
        const int na = a ? 2 : 1, nb = b ? 2 : 1, nc = c ? 2 : 1;
        for (int ia = 0; ia < na; ++ia) {
            const size_t pa = (size_t)(ia ? md.mx - a : a) * md.my;
            for (int ib = 0; ib < nb; ++ib) {
                const size_t pb = (pa + (ib ? md.my - b : b)) * md.mz;
                for (int ic = 0; ic < nc; ++ic) kpad[pb + (ic ? md.mz - c : c)] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) k_green_compact(const double* __restrict__ gtab, MeshDims md,
                                                           double* __restrict__ k1) {
    const long long total = (long long)md.nx * md.ny * md.nz;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        int c = (int)(t % md.nz);
        long long u = t / md.nz;
        int b = (int)(u % md.ny);
        int a = (int)(u / md.ny);
        k1[t] = green_entry(gtab, md.ny + 1, md.nz + 1, a, b, c);
    }
}

// zero-padded copy of rho (sc.py:142-143)
__global__ void __launch_bounds__(kThreads) k_pad_rho(const double* __restrict__ rho, MeshDims md,
                                                     double* __restrict__ pad) {
    const long long total = (long long)md.mx * md.my * md.mz;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        int c = (int)(t % md.mz);
        long long u = t / md.mz;
        int b = (int)(u % md.my);
        int a = (int)(u / md.my);
        double v = 0.0;
        if (a < md.nx && b < md.ny && c < md.nz) v = __ldg(rho + ((size_t)a * md.ny + b) * md.nz + c);
        pad[t] = v;
    }
}

// rho_hat *= K_hat, then the inverse transform's 1/M^3 (sc.py:164)
__global__ void __launch_bounds__(kThreads) k_multiply(cufftDoubleComplex* __restrict__ rho_hat,
                                                      const cufftDoubleComplex* __restrict__ k_hat, long long total,
                                                      double inv_m3) {
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        cufftDoubleComplex a = rho_hat[t], b = k_hat[t];
        cufftDoubleComplex o;
        o.x = (a.x * b.x - a.y * b.y) * inv_m3;
        o.y = (a.x * b.y + a.y * b.x) * inv_m3;
        rho_hat[t] = o;
    }
}

// phi = conv[:n,:n,:n] / (4 pi eps0 hx hy hz)  (sc.py:167-168)
__global__ void __launch_bounds__(kThreads) k_crop_phi(const double* __restrict__ conv, StepSrc src, ReduceState rs,
                                                      MeshDims md, double four_pi_eps0,
                                                      double* __restrict__ phi) {
    __shared__ double h[3];
    resolve_steps(src, rs, h);
    const double denom = four_pi_eps0 * h[0] * h[1] * h[2];
    const long long total = (long long)md.nx * md.ny * md.nz;
    for (long long t = (long long)blockIdx.x * kThreads + threadIdx.x; t < total; t += (long long)gridDim.x * kThreads) {
        int c = (int)(t % md.nz);
        long long u = t / md.nz;
        int b = (int)(u % md.ny);
        int a = (int)(u / md.ny);
        phi[t] = __ldg(conv + ((size_t)a * md.my + b) * md.mz + c) / denom;
    }
}

// staggered backward differences, last plane zero (sc.py:195-200), stored as
// the quad table the gather reads: equad[comp][i][j][k] = E_comp at
// (i,j,k), (i,j,k+1), (i,j+1,k), (i,j+1,k+1) with the upper indices clamped.
__device__ __forceinline__ double field_value(const double* __restrict__ phi, const MeshDims& md, const double* ih,
                                              int comp, int i, int j, int k) {
    const size_t sy = md.nz, sx = (size_t)md.ny * md.nz;
    const size_t t = (size_t)i * sx + (size_t)j * sy + k;
    // (phi - phi_next) * (1/h): within one ulp of the reference's (phi - phi_next)/h
    if (comp == 0) return (i < md.nx - 1) ? (__ldg(phi + t) - __ldg(phi + t + sx)) * ih[0] : 0.0;
    if (comp == 1) return (j < md.ny - 1) ? (__ldg(phi + t) - __ldg(phi + t + sy)) * ih[1] : 0.0;
    return (k < md.nz - 1) ? (__ldg(phi + t) - __ldg(phi + t + 1)) * ih[2] : 0.0;
}

// grid = (ceil(nz*ny / threads), nx, 3): no integer division by runtime strides per thread
__global__ void __launch_bounds__(kThreads) k_field(const double* __restrict__ phi, StepSrc src, ReduceState rs,
                                                   MeshDims md, EQuad* __restrict__ equad) {
    pdl_enter();
    __shared__ double h[3];
    __shared__ double ih[3];
    resolve_steps(src, rs, h);
    if (threadIdx.x < 3) ih[threadIdx.x] = 1.0 / h[threadIdx.x];
    __syncthreads();
    const int comp = blockIdx.z, i = blockIdx.y;
    const int jk = blockIdx.x * kThreads + threadIdx.x;
    if (jk >= md.ny * md.nz) return;
    const int j = jk / md.nz, k = jk - j * md.nz;
    const int j1 = min(j + 1, md.ny - 1), k1 = min(k + 1, md.nz - 1);
    EQuad e;
    e.v00 = field_value(phi, md, ih, comp, i, j, k);
    e.v01 = field_value(phi, md, ih, comp, i, j, k1);
    e.v10 = field_value(phi, md, ih, comp, i, j1, k);
    e.v11 = field_value(phi, md, ih, comp, i, j1, k1);
    equad[((size_t)comp * md.nx + i) * md.ny * md.nz + jk] = e;
}

// The same table stored x-fastest for the lane-pair gather (gather_pair): rec(comp, i, j, k) at
// equad[comp*cells + (j*nz + k)*nx + i].  One block = one j and kFieldKT consecutive k: the phi values
// it needs ((j..j+2) x (k..k+KT+1) for every i) are staged in shared memory with k-contiguous reads,
// and the records leave with i fastest, i.e. fully coalesced 32-byte stores.
#ifndef OCL_FIELD_KT
#define OCL_FIELD_KT 8
#endif
constexpr int kFieldKT = OCL_FIELD_KT;
__global__ void __launch_bounds__(kThreads) k_field_xf(const double* __restrict__ phi, StepSrc src, ReduceState rs,
                                                      MeshDims md, EQuad* __restrict__ equad) {
    pdl_enter();
    extern __shared__ double tile[];                      // [nx][3][KT + 2]
    __shared__ double h[3];
    __shared__ double ih[3];
    resolve_steps(src, rs, h);
    if (threadIdx.x < 3) ih[threadIdx.x] = 1.0 / h[threadIdx.x];
    constexpr int KW = kFieldKT + 2;
    const int j = blockIdx.y, k0 = blockIdx.x * kFieldKT;
    const int nx = md.nx, ny = md.ny, nz = md.nz;
    for (int t = threadIdx.x; t < nx * 3 * KW; t += kThreads) {
        const int kk = t % KW, u = t / KW, jj = u % 3, i = u / 3;
        const int js = min(j + jj, ny - 1), ks = min(k0 + kk, nz - 1);            // clamped: never used beyond the edge
        tile[t] = __ldg(phi + ((size_t)i * ny + js) * nz + ks);
    }
    __syncthreads();
    const size_t cells = (size_t)nx * ny * nz;
    auto P = [&](int i, int jj, int kk) { return tile[(i * 3 + jj) * KW + kk]; };
    // E at (i, j + dj, k0 + kk + dk) with the upper indices clamped like the z-fastest table
    auto E = [&](int comp, int i, int jj, int kk) -> double {
        const int jg = min(j + jj, ny - 1), kg = min(k0 + kk, nz - 1);
        const int jl = jg - j, kl = kg - k0;                                       // local (clamped) tile coordinates
        if (comp == 0) return (i < nx - 1) ? (P(i, jl, kl) - P(i + 1, jl, kl)) * ih[0] : 0.0;
        if (comp == 1) return (jg < ny - 1) ? (P(i, jl, kl) - P(i, jl + 1, kl)) * ih[1] : 0.0;
        return (kg < nz - 1) ? (P(i, jl, kl) - P(i, jl, kl + 1)) * ih[2] : 0.0;
    };
    const int kt = min(kFieldKT, nz - k0);
    for (int t = threadIdx.x; t < 3 * kt * nx; t += kThreads) {
        const int i = t % nx, u = t / nx, kk = u % kt, comp = u / kt;
        EQuad e;
        e.v00 = E(comp, i, 0, kk);
        e.v01 = E(comp, i, 0, kk + 1);
        e.v10 = E(comp, i, 1, kk);
        e.v11 = E(comp, i, 1, kk + 1);
        equad[comp * cells + ((size_t)j * nz + (k0 + kk)) * nx + i] = e;
    }
}

// ---------------------------------------------------------------------------
// sweep 4: gather, kick, back-transform (sc.py:201-204, :244-251)
//   LAYOUT 0: z-fastest quad table, six independent 256-bit gathers per particle
//   LAYOUT 1: x-fastest quad table fetched by lane pairs (gather_pair), rows staged through shared memory
//   LAYOUT 2: the same with the rows prefetched into registers (prefetch_sweep)
// ---------------------------------------------------------------------------
template <bool KICK, bool TAP, int LAYOUT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB) k_gather_kick(double* __restrict__ r, long long ld, long long n,
                                                            KP kp, ReduceState rs, MeshDims md,
                                                            const EQuad* __restrict__ equad,
                                                            double* __restrict__ exyz_out) {
    pdl_enter();
    const RefParams rp = kp_ref(kp);
    const double cdT = kp_cdT(kp);
    constexpr bool STAGED = LAYOUT < 2;                  // LAYOUT 2: pair gather with register prefetch of the rows
    __shared__ double pipe[STAGED ? kPipeDepth * 6 * kThreads : 1];
    __shared__ Geo sg;
    load_geo(rs.geo, &sg);
    const Frame f = sg.f;
    const Mesh m = sg.m;
    const double kt = cdT * (1.0 - f.beta0 * f.beta0);   // sc.py:246-247
    const size_t cells = (size_t)md.nx * md.ny * md.nz;
    const EQuad* __restrict__ ex = equad;
    const EQuad* __restrict__ ey = equad + cells;
    const EQuad* __restrict__ ez = equad + 2 * cells;
    const double* const base[6] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld};
    auto body = [&](int i, const double (&w)[6], bool valid) {
        Cart c = mad_to_cart(rp, w[0], w[1], w[2], w[3], w[4], w[5]);
        double a, b, g, g0, g1, g2;
        rotate_stretch(f, c.x, c.y, c.z, a, b, g);
        to_grid(m, a, b, g, g0, g1, g2);
        double e0, e1, e2;
        if constexpr (LAYOUT >= 1) {
            gather_pair(ex, ey, ez, md.nx, md.ny, md.nz, g0, g1, g2, e0, e1, e2);                // :202-204
            e0 *= f.gamma0; e1 *= f.gamma0;
        } else {
            e0 = trilinear(ex, md.nx, md.ny, md.nz, g0, g1 + 0.5, g2 + 0.5) * f.gamma0;
            e1 = trilinear(ey, md.nx, md.ny, md.nz, g0 + 0.5, g1, g2 + 0.5) * f.gamma0;
            e2 = trilinear(ez, md.nx, md.ny, md.nz, g0 + 0.5, g1 + 0.5, g2);
        }
        if (!valid) return;
        if (TAP) {
            exyz_out[3 * (size_t)i + 0] = e0; exyz_out[3 * (size_t)i + 1] = e1; exyz_out[3 * (size_t)i + 2] = e2;
        }
        if (KICK) {
            // momenta into the bunch frame (sc.py:234)
            double p0 = c.px * f.T[0][0] + c.py * f.T[1][0] + c.pz * f.T[2][0];
            double p1 = c.px * f.T[0][1] + c.py * f.T[1][1] + c.pz * f.T[2][1];
            double p2 = c.px * f.T[0][2] + c.py * f.T[1][2] + c.pz * f.T[2][2];
            p0 = p0 + kt * e0;                                                             // :246
            p1 = p1 + kt * e1;                                                             // :247
            p2 = p2 + cdT * e2;                                                            // :248
            // back to the lab axes (sc.py:249-250)
            c.px = p0 * f.T[0][0] + p1 * f.T[0][1] + p2 * f.T[0][2];
            c.py = p0 * f.T[1][0] + p1 * f.T[1][1] + p2 * f.T[1][2];
            c.pz = p0 * f.T[2][0] + p1 * f.T[2][1] + p2 * f.T[2][2];
            double x, xs, y, ys, tau, delta;
            cart_to_mad(rp, c, x, xs, y, ys, tau, delta);                                  // :251
            r[i] = x; r[ld + i] = xs; r[2 * ld + i] = y; r[3 * ld + i] = ys; r[4 * ld + i] = tau;
            r[5 * ld + i] = delta;
        }
    };
    if constexpr (LAYOUT == 2) {
        prefetch_sweep<6>(base, (int)n, body);
    } else if constexpr (LAYOUT == 1) {
        pipelined_sweep<6, kPipeDepth, true>(base, (int)n, pipe, body);
    } else {
        pipelined_sweep<6, kPipeDepth, false>(base, (int)n, pipe,
                                              [&](int i, const double (&w)[6]) { body(i, w, true); });
    }
}

// stand-alone transforms (known-answer tests)
__global__ void __launch_bounds__(kThreads) k_mad_to_cart(const double* __restrict__ r, long long ld, long long n,
                                                         RefParams rp, double* __restrict__ xp, long long lx) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        Cart c = mad_to_cart(rp, r[i], r[ld + i], r[2 * ld + i], r[3 * ld + i], r[4 * ld + i], r[5 * ld + i]);
        xp[i] = c.x; xp[lx + i] = c.y; xp[2 * lx + i] = c.z;
        xp[3 * lx + i] = c.px; xp[4 * lx + i] = c.py; xp[5 * lx + i] = c.pz;
    }
}

__global__ void __launch_bounds__(kThreads) k_cart_to_mad(const double* __restrict__ xp, long long lx, long long n,
                                                         RefParams rp, double* __restrict__ r, long long ld) {
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads) {
        Cart c;
        c.x = xp[i]; c.y = xp[lx + i]; c.z = xp[2 * lx + i];
        c.px = xp[3 * lx + i]; c.py = xp[4 * lx + i]; c.pz = xp[5 * lx + i];
        double x, xs, y, ys, tau, delta;
        cart_to_mad(rp, c, x, xs, y, ys, tau, delta);
        r[i] = x; r[ld + i] = xs; r[2 * ld + i] = y; r[3 * ld + i] = ys; r[4 * ld + i] = tau; r[5 * ld + i] = delta;
    }
}

// ---------------------------------------------------------------------------
// launch wrappers
// ---------------------------------------------------------------------------
static inline int grid_for(long long total, int cap) {
    long long b = (total + kThreads - 1) / kThreads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}
constexpr int kSweepCap = 148 * 4;    // persistent grid-stride sweeps: 4 resident blocks per SM
constexpr int kGridCap = 148 * 16;    // grid kernels

__global__ void k_set_params(KickParams v, KickParams* dst) {
    pdl_enter();
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = v;
}
// stand-alone exchange / barrier (which = 2), one warp
__global__ void k_mailbox_exchange(Mailbox mb, int which, ReduceState rs, int* __restrict__ err_flag) {
    pdl_enter();
    mailbox_exchange_warp(mb, which, rs, err_flag);
}
void launch_mailbox_exchange(Mailbox mb, int which, ReduceState rs, int* err_flag, cudaStream_t st) {
    launch_k(k_mailbox_exchange, dim3(1), dim3(32), 0, st, mb, which, rs, err_flag);
}

// ---------------------------------------------------------------------------
// charge-grid reduction inside the NVSwitch (NVLS): every rank's rho lives at the same offset of a
// symmetric allocation that is also mapped as ONE multicast address range.  A multimem.ld_reduce on
// that range makes the switch fetch the element from all ranks and return the sum (LDGMC.E.ADD.F64);
// a multimem.st broadcasts the result back into every rank's copy.  Each rank reduces 1/world of the
// grid, so a full all-reduce moves 2 x n^3 x 8 / world bytes per link; every element is summed once,
// by the switch, so all ranks end up with bit-identical grids.
//   out == nullptr : all-reduce in place (redundant solve on every rank)
//   out != nullptr : reduce-scatter: elements [lo, hi) of the sum go to out[0 .. hi-lo) (slab solve)
// Default form: a one-warp barrier kernel ("every rank's deposit is complete"), the reduction, a one-warp barrier
// kernel ("every slice is final"), all three launched programmatically (PDL).  A single-kernel form with both
// barriers inside exists (k_nvls_reduce, OCL_SC_NVLS_FUSED=1) and measured slower, see there.
// ---------------------------------------------------------------------------
// elements [lo, hi) of the multicast range: in-switch sum, then broadcast (out == nullptr) or local store.
// kNvlsUnroll reductions are in flight per thread before the first store: the kernel runs on a SMALL grid (it
// may have to wait for the peers with all its blocks resident, and the K_hat chain shares the GPU with it:
// 592 waiting blocks cost 11 us per kick on 2 x B200), so each thread carries several elements.
constexpr int kNvlsUnroll = 4;
__device__ __forceinline__ void nvls_reduce_range(double* __restrict__ mc, long long lo, long long hi,
                                                  double* __restrict__ out) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * kNvlsUnroll) {
        double v[kNvlsUnroll];
#pragma unroll
        for (int u = 0; u < kNvlsUnroll; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(v[u]) : "l"(mc + i) : "memory");
        }
#pragma unroll
        for (int u = 0; u < kNvlsUnroll; ++u) {
            const long long i = i0 + u * stride;
            if (i < hi) {
                if (out) out[i - lo] = v[u];
                else asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i), "d"(v[u]) : "memory");
            }
        }
    }
}

// Single-kernel form (OCL_SC_NVLS_FUSED=1; NOT the default): entry barrier, reduction and exit barrier in one
// launch.  Measured on 2 x B200 (1 M / 63^3 per GPU, tools/r2_comm_probe.py): 176.9 us per kick against 173.9 us
// for the three launches below -- waiting for the peers inside a wide grid costs more than the two extra
// launches of one-warp barrier kernels save (and 193.7 us while every waiting block still issued a membar.sys).
// mode bits (timing experiments): 1 = leave out the entry barrier, 2 = leave out the exit barrier (both unsafe),
// 4 = one system fence per block instead of one per thread
__global__ void __launch_bounds__(256) k_nvls_reduce(double* __restrict__ mc, long long lo, long long hi,
                                                    double* __restrict__ out, Mailbox mb, unsigned int* ticket,
                                                    int* __restrict__ err_flag, int mode) {
    pdl_enter();
    const unsigned long long epoch = mb.epoch[2] + 1;       // advanced by the last block, after everyone has read it
    if (!(mode & 1)) {
        if (threadIdx.x < 32) {
            if (blockIdx.x == 0) mailbox_signal(mb, kFlagRhoReady, epoch);
            mailbox_wait(mb, kFlagRhoReady, epoch, err_flag);
        }
        __syncthreads();
    }
    nvls_reduce_range(mc, lo, hi, out);
    if (mode & 4) {
        __syncthreads();
        if (threadIdx.x == 0) __threadfence_system();
    } else {
        __threadfence_system();
        __syncthreads();
    }
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last || threadIdx.x >= 32) return;
    // every block of this rank has issued (and fenced) its multimem stores
    if (!(mode & 2)) {
        mailbox_signal(mb, kFlagRhoDone, epoch);
        mailbox_wait(mb, kFlagRhoDone, epoch, err_flag);
    }
    if (threadIdx.x == 0) { mb.epoch[2] = epoch; *ticket = 0; __threadfence(); }
}
// the reduction proper; launch_nvls_reduce brackets it with two one-warp barrier kernels (the default form)
__global__ void __launch_bounds__(256) k_nvls_reduce_plain(double* __restrict__ mc, long long lo, long long hi,
                                                          double* __restrict__ out) {
    pdl_enter();
    nvls_reduce_range(mc, lo, hi, out);
}
static int nvls_mode() {      // -1: three launches (default); >= 0: the single-kernel form with these mode bits
    static int m = -2;
    if (m == -2) {
        const char* f = getenv("OCL_SC_NVLS_FUSED");
        const char* e = getenv("OCL_SC_NVLS_MODE");
        m = (f && atoi(f)) ? (e ? atoi(e) : 0) : -1;
    }
    return m;
}
void launch_nvls_reduce(double* mc, long long lo, long long hi, double* out, Mailbox mb, ReduceState rs, unsigned int* ticket,
                        int* err_flag, cudaStream_t st) {
    const int mode = nvls_mode();
    // one block per kNvlsUnroll * 256 elements: 123 blocks for the 63^3 grid on two ranks, at most 4 per SM
    long long blocks = (hi - lo + 256 * kNvlsUnroll - 1) / (256 * kNvlsUnroll);
    if (blocks < 1) blocks = 1;
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (mode < 0) {
        launch_mailbox_exchange(mb, 2, rs, err_flag, st);
        launch_k(k_nvls_reduce_plain, dim3((int)blocks), dim3(256), 0, st, mc, lo, hi, out);
        launch_mailbox_exchange(mb, 2, rs, err_flag, st);
        return;
    }
    launch_k(k_nvls_reduce, dim3((int)blocks), dim3(256), 0, st, mc, lo, hi, out, mb, ticket, err_flag, mode);
}

// fold all-gathered extents: max over ranks of the first 6 doubles, sum of the last 4
__global__ void k_combine_extents(const double* __restrict__ all, int world, double* __restrict__ emax,
                                  double* __restrict__ esum) {
    pdl_enter();
    const int k = threadIdx.x;
    if (k >= 10) return;
    double v = all[k];
    for (int w = 1; w < world; ++w) {
        const double x = all[(size_t)w * 10 + k];
        v = (k < 6) ? fmax(v, x) : v + x;
    }
    if (k < 6) emax[k] = v; else esum[k - 6] = v;
}
void launch_combine_extents(const double* all, int world, ReduceState rs, cudaStream_t st) {
    launch_k(k_combine_extents, dim3(1), dim3(32), 0, st, all, world, rs.emax, rs.esum);
}

void launch_set_params(KickParams v, KickParams* dst, cudaStream_t st) { launch_k(k_set_params, dim3(1), dim3(32), 0, st, v, dst); }
const void* set_params_kernel() { return (const void*)k_set_params; }

// Rows go through the bulk-copy engine (bulk_sweep: cp.async.bulk + mbarrier) when every row base is 16-byte aligned
// and the sweep is long enough for it to pay.  Measured on B200, whole kick: 12.5 M / 127^3 996.6 -> 992.1 us, 50 M /
// 255^3 6626 -> 6565 us (k_momentum 205 -> 189 us, k_extent 445 -> 425 us), but 1 M / 63^3 133.7 -> 136.6 us (the
// per-tile block barrier costs more than the per-thread cp.async pipeline when a block sees only 9 tiles).
// OCL_SC_TMA=0 / 1 forces one path.
static bool tma_rows(const double* r, long long ld, const double* q, long long n) {
    static int mode = -2;
    if (mode == -2) { const char* e = getenv("OCL_SC_TMA"); mode = e ? atoi(e) : -1; }
    if (mode == 0) return false;
    if (mode < 0 && n < 4000000) return false;
    return ((uintptr_t)r % 16 == 0) && (ld % 2 == 0) && ((uintptr_t)q % 16 == 0);
}
void launch_momentum(const double* r, long long ld, long long n, KP kp, ReduceState rs, Mailbox mb, int* mb_err,
                     KickParams* publish, cudaStream_t st) {
    const dim3 grid(particle_grid(n, rs.max_blocks));
    if (tma_rows(r, ld, nullptr, n))
        launch_k(k_momentum<true>, grid, dim3(kThreads), 0, st, r, ld, n, kp, rs, mb, mb_err, publish);
    else
        launch_k(k_momentum<false>, grid, dim3(kThreads), 0, st, r, ld, n, kp, rs, mb, mb_err, publish);
}
const void* momentum_kernel(int tma) { return tma ? (const void*)k_momentum<true> : (const void*)k_momentum<false>; }
void launch_extent(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs, MeshDims md,
                   Mailbox mb, int* mb_err, cudaStream_t st) {
    if (tma_rows(r, ld, q, n))
        launch_k(k_extent<true>, dim3(particle_grid(n, 148 * 3)), dim3(kThreads), 0, st, r, ld, q, n, kp, rs, md, mb, mb_err);
    else
        launch_k(k_extent<false>, dim3(particle_grid(n, 148 * 3)), dim3(kThreads), 0, st, r, ld, q, n, kp, rs, md, mb, mb_err);
}
void launch_finish(int which, KP kp, ReduceState rs, MeshDims md, cudaStream_t st) {
    launch_k(k_finish, dim3(1), dim3(32), 0, st, which, kp, rs, md);
}
void launch_deposit(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs,
                    MeshDims md, double* rho, cudaStream_t st) {
    if (tma_rows(r, ld, q, n))
        launch_k(k_deposit<true>, dim3(grid_for(n, 148 * 3)), dim3(kThreads), 0, st, r, ld, q, n, kp, rs, md, rho);
    else
        launch_k(k_deposit<false>, dim3(grid_for(n, 148 * 3)), dim3(kThreads), 0, st, r, ld, q, n, kp, rs, md, rho);
}
void launch_momentum_exact(const double* r, long long ld, long long n, KP kp, const uint2* leaves, int nleaves,
                           const uint2* nodes, const int* level_start, int nlevels, double* V, double* sums,
                           cudaStream_t st) {
    launch_k(k_momentum_exact_leaves, dim3((nleaves + 7) / 8), dim3(256), 0, st, r, ld, kp, leaves, nleaves, V);
    launch_k(k_pairwise_combine, dim3(1), dim3(1024), 0, st, V, nodes, level_start, nlevels, nleaves, n, sums);
}
// scratch layout: key_in | val_in | key_out | val_out (n unsigned each, 256-byte aligned) | cub temporary storage
static size_t ordered_pad(long long n) { return ((size_t)n * sizeof(unsigned) + 255) / 256 * 256; }
size_t deposit_ordered_scratch_bytes(long long n, MeshDims md) {
    size_t tmp = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned*)nullptr, (unsigned*)nullptr, (const unsigned*)nullptr,
                                    (unsigned*)nullptr, (int)n, 0, 32);
    (void)md;
    return 4 * ordered_pad(n) + tmp + 256;
}
int launch_deposit_ordered(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs,
                           MeshDims md, double* rho, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    const size_t pad = ordered_pad(n);
    if (scratch_bytes < 4 * pad) return 1;
    char* base = static_cast<char*>(scratch);
    unsigned* key_in = reinterpret_cast<unsigned*>(base);
    unsigned* val_in = reinterpret_cast<unsigned*>(base + pad);
    unsigned* key_out = reinterpret_cast<unsigned*>(base + 2 * pad);
    unsigned* val_out = reinterpret_cast<unsigned*>(base + 3 * pad);
    void* tmp = base + 4 * pad;
    size_t tmp_bytes = scratch_bytes - 4 * pad;
    launch_k(k_cell_index, dim3(grid_for(n, 148 * 3)), dim3(kThreads), 0, st, r, ld, n, kp, rs, md, key_in, val_in);
    // all 32 key bits: particles outside the mesh carry the key 0xffffffff and must sort behind every cell
    if (cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, key_in, key_out, val_in, val_out, (int)n, 0, 32, st) != cudaSuccess)
        return 1;
    const unsigned cells = (unsigned)((size_t)md.nx * md.ny * md.nz);
    launch_k(k_ordered_sum, dim3(grid_for(n, 148 * 8)), dim3(kThreads), 0, st, (const unsigned*)key_out,
             (const unsigned*)val_out, q, n, cells, rho);
    return 0;
}
void launch_green_table(ReduceState rs, MeshDims md, double* gtab, double* h3, cudaStream_t st) {
    StepSrc src;
    src.given = 0; src.h[0] = src.h[1] = src.h[2] = 0.0;
    long long total = (long long)(md.nx + 1) * (md.ny + 1) * (md.nz + 1);
    launch_k(k_green_table, dim3(grid_for(total, kGridCap)), dim3(kThreads), 0, st, src, rs, md, gtab, h3);
}
void launch_green_table_steps(const double steps[3], MeshDims md, double* gtab, double* h3, cudaStream_t st) {
    StepSrc src;
    src.given = 1; src.h[0] = steps[0]; src.h[1] = steps[1]; src.h[2] = steps[2];
    ReduceState rs = {};
    long long total = (long long)(md.nx + 1) * (md.ny + 1) * (md.nz + 1);
    launch_k(k_green_table, dim3(grid_for(total, kGridCap)), dim3(kThreads), 0, st, src, rs, md, gtab, h3);
}
void launch_green_mirror(const double* gtab, MeshDims md, double* kpad, cudaStream_t st) {
    cudaMemsetAsync(kpad, 0, sizeof(double) * (size_t)md.mx * md.my * md.mz, st);
    long long total = (long long)md.nx * md.ny * md.nz;
    k_green_mirror<<<grid_for(total, kGridCap), kThreads, 0, st>>>(gtab, md, kpad);
}
void launch_green_compact(const double* gtab, MeshDims md, double* k1, cudaStream_t st) {
    long long total = (long long)md.nx * md.ny * md.nz;
    k_green_compact<<<grid_for(total, kGridCap), kThreads, 0, st>>>(gtab, md, k1);
}
void launch_pad_rho(const double* rho, MeshDims md, double* pad, cudaStream_t st) {
    long long total = (long long)md.mx * md.my * md.mz;
    k_pad_rho<<<grid_for(total, kGridCap), kThreads, 0, st>>>(rho, md, pad);
}
void launch_multiply(cufftDoubleComplex* rho_hat, const cufftDoubleComplex* k_hat, MeshDims md, cudaStream_t st) {
    long long total = (long long)md.mx * md.my * (md.mz / 2 + 1);
    double inv = 1.0 / ((double)md.mx * (double)md.my * (double)md.mz);
    k_multiply<<<grid_for(total, kGridCap), kThreads, 0, st>>>(rho_hat, k_hat, total, inv);
}
double four_pi_eps0_value();
static double four_pi_eps0() { return four_pi_eps0_value(); }
double four_pi_eps0_value() {
    const double pi = 3.141592653589793, c = 299792458.0;   // ocelot/common/globals.py:13-24
    const double mu0 = 4 * pi * 1e-7;
    const double eps0 = 1 / mu0 / (c * c);
    return 4 * pi * eps0;
}
void launch_crop_phi(const double* conv, ReduceState rs, MeshDims md, double* phi, cudaStream_t st) {
    StepSrc src;
    src.given = 0; src.h[0] = src.h[1] = src.h[2] = 0.0;
    long long total = (long long)md.nx * md.ny * md.nz;
    k_crop_phi<<<grid_for(total, kGridCap), kThreads, 0, st>>>(conv, src, rs, md, four_pi_eps0(), phi);
}
void launch_crop_phi_steps(const double* conv, const double steps[3], MeshDims md, double* phi, cudaStream_t st) {
    StepSrc src;
    src.given = 1; src.h[0] = steps[0]; src.h[1] = steps[1]; src.h[2] = steps[2];
    ReduceState rs = {};
    long long total = (long long)md.nx * md.ny * md.nz;
    k_crop_phi<<<grid_for(total, kGridCap), kThreads, 0, st>>>(conv, src, rs, md, four_pi_eps0(), phi);
}
size_t field_tile_bytes(MeshDims md) { return sizeof(double) * (size_t)md.nx * 3 * (kFieldKT + 2); }
void field_init_kernels() {
    cudaFuncSetAttribute(k_field_xf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}
void launch_field(const double* phi, ReduceState rs, MeshDims md, EQuad* equad, int layout, cudaStream_t st) {
    StepSrc src;
    src.given = 0; src.h[0] = src.h[1] = src.h[2] = 0.0;
    if (layout == 1) {
        dim3 grid((md.nz + kFieldKT - 1) / kFieldKT, md.ny, 1);
        launch_k(k_field_xf, dim3(grid), dim3(kThreads), field_tile_bytes(md), st, phi, src, rs, md, equad);
    } else {
        dim3 grid((md.ny * md.nz + kThreads - 1) / kThreads, md.nx, 3);
        launch_k(k_field, dim3(grid), dim3(kThreads), 0, st, phi, src, rs, md, equad);
    }
}
// Resident blocks per SM of the gather (= its register budget): 2 blocks x 128 registers, or 3 x 85.  The
// lane-pair gather is bound by the L1TEX data stage; with its smaller wavefront count a third block pays once
// the kernel runs long enough to reach steady state (measured on B200, 12.5 M / 127^3: 428 -> 403 us), while at
// 1 M particles two blocks are faster (39 vs 41 us).  OCL_SC_GK_BLOCKS overrides.
static int gather_blocks(long long n, int layout) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("OCL_SC_GK_BLOCKS");
        forced = e ? atoi(e) : 0;
    }
    if (forced == 2 || forced == 3 || (forced == 4 && layout == 2)) return forced;
    (void)n; (void)layout;
    return 2;
}
template <int LAYOUT, int MINB>
static void launch_gather_kick_l(double* r, long long ld, long long n, KP kp, ReduceState rs, MeshDims md,
                                 const EQuad* equad, double* exyz_out, int do_kick, cudaStream_t st) {
    int grid = grid_for(n, 148 * MINB);
    if (do_kick && exyz_out)
        launch_k(k_gather_kick<true, true, LAYOUT, MINB>, dim3(grid), dim3(kThreads), 0, st, r, ld, n, kp, rs, md, equad, exyz_out);
    else if (do_kick)
        launch_k(k_gather_kick<true, false, LAYOUT, MINB>, dim3(grid), dim3(kThreads), 0, st, r, ld, n, kp, rs, md, equad, nullptr);
    else
        launch_k(k_gather_kick<false, true, LAYOUT, MINB>, dim3(grid), dim3(kThreads), 0, st, r, ld, n, kp, rs, md, equad, exyz_out);
}
void launch_gather_kick(double* r, long long ld, long long n, KP kp, ReduceState rs, MeshDims md,
                        const EQuad* equad, double* exyz_out, int do_kick, int layout, cudaStream_t st) {
    const int mb = gather_blocks(n, layout);
    if (layout == 2 && mb == 4) launch_gather_kick_l<2, 4>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else if (layout == 2 && mb == 3) launch_gather_kick_l<2, 3>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else if (layout == 2) launch_gather_kick_l<2, 2>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else if (layout == 1 && mb == 3) launch_gather_kick_l<1, 3>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else if (layout == 1) launch_gather_kick_l<1, 2>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else if (mb == 3) launch_gather_kick_l<0, 3>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
    else launch_gather_kick_l<0, 2>(r, ld, n, kp, rs, md, equad, exyz_out, do_kick, st);
}
void launch_mad_to_cart(const double* r, long long ld, long long n, RefParams rp, double* xp, long long ld_xp,
                        cudaStream_t st) {
    k_mad_to_cart<<<grid_for(n, kSweepCap), kThreads, 0, st>>>(r, ld, n, rp, xp, ld_xp);
}
void launch_cart_to_mad(const double* xp, long long ld_xp, long long n, RefParams rp, double* r, long long ld,
                        cudaStream_t st) {
    k_cart_to_mad<<<grid_for(n, kSweepCap), kThreads, 0, st>>>(xp, ld_xp, n, rp, r, ld);
}

}  // namespace ocl
