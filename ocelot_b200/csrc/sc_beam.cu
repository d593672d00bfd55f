// sc_beam.cu -- the two per-particle steps either side of the space-charge kick in the
// reference's tracking loop, so that a bunch can stay
// resident in HBM between kicks:
//
//   k_map_apply    X <- R X + T:XX + B     TransferMap.mul_p_array (transformations/transfer_map.py:42-53)
//                                          SecondTM.t_apply        (transformations/second_order.py:31-39,
//                                                                   tm_utils.py:54-55)
//   k_moments_1/2  first moments, then centred second moments: get_envelope default path
//                                          (beam/analysis.py:72-76, :121-123, :125-166)
//
// Both stream the six coordinate rows once (96 B resp. 48 B per particle) through the same
// cp.async row pipeline as the kick sweeps.
#include "sc_kernels.h"

namespace ocl {

constexpr int kBThreads = kSweepThreads;
constexpr int kBWarps = kBThreads / 32;
constexpr int kBDepth = 3;

// ---- reductions (sum only), fixed order ------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = warp_reduce(v[k], OpSum());
        if (lane == 0) sh[k * kBWarps + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = (lane < kBWarps) ? sh[k * kBWarps + lane] : 0.0;
            v[k] = warp_reduce(x, OpSum());
        }
    }
    __syncthreads();
}

template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], double* part, unsigned int* ticket, double* sh) {
    __shared__ bool last;
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[(size_t)blockIdx.x * NV + k] = v[k];
        __threadfence();
        last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kBThreads) {
#pragma unroll
        for (int k = 0; k < NV; ++k) v[k] += __ldcg(part + (size_t)b * NV + k);
    }
    block_sum<NV>(v, sh);
    if (threadIdx.x == 0) *ticket = 0;
    return true;
}

// ---- f1: transfer map ----------------------------------------------------------------------
__global__ void __launch_bounds__(kBThreads, 3) k_map_apply(double* __restrict__ r, long long ld, long long n,
                                                           MapCoef mc) {
    __shared__ double pipe[kBDepth * 6 * kBThreads];
    const double* const base[6] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld};
    pipelined_sweep<6, kBDepth>(base, (int)n, pipe, [&](int i, const double (&x)[6]) {
        double y[6];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
            double s = mc.R[a * 6] * x[0];                       // R X (transfer_map.py:51)
#pragma unroll
            for (int b = 1; b < 6; ++b) s += mc.R[a * 6 + b] * x[b];
            y[a] = s;
        }
        if (mc.nt > 0) {                                         // + T : X X, non-zero terms only (tm_utils.py:55)
            double t[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int m = 0; m < mc.nt; ++m) {
                const int c = mc.tidx[m];
                const int a = c / 36, j = (c / 6) % 6, k = c % 6;
                const double term = mc.tval[m] * x[j] * x[k];
#pragma unroll
                for (int q = 0; q < 6; ++q) t[q] += (q == a) ? term : 0.0;
            }
#pragma unroll
            for (int a = 0; a < 6; ++a) y[a] += t[a];
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) y[a] += mc.B[a];                   // + B (transfer_map.py:51, second_order.py:37)
        if (mc.cav) {                                                  // cavity.py:56-126, x[4], x[5] are the saved X4, X5
            if (mc.cav == 1) y[5] = x[5] * mc.c1 + mc.c2 * (cos(-x[4] * mc.kb + mc.phi) - mc.cosphi);   // :81-84
            y[4] += mc.t566 * x[5] * x[5] + mc.t556 * x[4] * x[5] + mc.t555 * x[4] * x[4];              // :126
        }
#pragma unroll
        for (int a = 0; a < 6; ++a) r[a * ld + i] = y[a];
    });
}

// ---- f2: beam moments ----------------------------------------------------------------------
// pass 1: sums of x, px*f, y, py*f, tau, p with f = 1 - p - p^2/2 + px^2/2 + py^2/2 (analysis.py:121-123)
// NR = 7: the charges are a seventh row and out[18] = sum q (the bunch charge that get_envelope reports, analysis.py:81)
template <int NR>
__global__ void __launch_bounds__(kBThreads, 4) k_moments_1(const double* __restrict__ r, long long ld, long long n,
                                                           const double* __restrict__ q, ReduceState rs,
                                                           double* __restrict__ out) {
    __shared__ double sh[NR * kBWarps];
    __shared__ double pipe[kBDepth * NR * kBThreads];
    double v[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) v[k] = 0.0;
    const double* base[NR];
#pragma unroll
    for (int k = 0; k < 6; ++k) base[k] = r + k * ld;
    if constexpr (NR == 7) base[6] = q;
    pipelined_sweep<NR, kBDepth>(base, (int)n, pipe, [&](int, const double (&x)[NR]) {
        const double p = x[5];
        const double f = 1. - p - 0.5 * p * p + 0.5 * x[1] * x[1] + 0.5 * x[3] * x[3];
        v[0] += x[0]; v[1] += x[1] * f; v[2] += x[2]; v[3] += x[3] * f; v[4] += x[4]; v[5] += p;
        if constexpr (NR == 7) v[6] += x[6];
    });
    if (grid_sum<NR>(v, rs.part, rs.ticket + 2, sh) && threadIdx.x == 0) {
        const double inv = 1.0 / (double)n;
#pragma unroll
        for (int k = 0; k < 6; ++k) out[k] = v[k] * inv;        // np.mean
        if constexpr (NR == 7) out[18] = v[6];
    }
}

// pass 2: centred second moments (analysis.py:153-166), means from pass 1 in out[0..5]
// out[6..17] = xx, xpx, pxpx, yy, ypy, pypy, tautau, pp, xy, pxpy, xpy, ypx
__global__ void __launch_bounds__(kBThreads, 3) k_moments_2(const double* __restrict__ r, long long ld, long long n,
                                                           ReduceState rs, double* __restrict__ out) {
    __shared__ double sh[12 * kBWarps];
    __shared__ double pipe[kBDepth * 6 * kBThreads];
    const double mx = out[0], mpx = out[1], my = out[2], mpy = out[3], mt = out[4], mp = out[5];
    double v[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) v[k] = 0.0;
    const double* const base[6] = {r, r + ld, r + 2 * ld, r + 3 * ld, r + 4 * ld, r + 5 * ld};
    pipelined_sweep<6, kBDepth>(base, (int)n, pipe, [&](int, const double (&x)[6]) {
        const double p = x[5];
        const double f = 1. - p - 0.5 * p * p + 0.5 * x[1] * x[1] + 0.5 * x[3] * x[3];
        const double dx = x[0] - mx, dpx = x[1] * f - mpx, dy = x[2] - my, dpy = x[3] * f - mpy;
        const double dt = x[4] - mt, dp = p - mp;
        v[0] += dx * dx; v[1] += dx * dpx; v[2] += dpx * dpx;
        v[3] += dy * dy; v[4] += dy * dpy; v[5] += dpy * dpy;
        v[6] += dt * dt; v[7] += dp * dp;
        v[8] += dx * dy; v[9] += dpx * dpy; v[10] += dx * dpy; v[11] += dy * dpx;
    });
    if (grid_sum<12>(v, rs.part, rs.ticket + 3, sh) && threadIdx.x == 0) {
        const double inv = 1.0 / (double)n;
#pragma unroll
        for (int k = 0; k < 12; ++k) out[6 + k] = v[k] * inv;
    }
}

static int sweep_grid(long long n, int cap) {
    long long b = (n + kBThreads - 1) / kBThreads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

void launch_map_apply(double* r, long long ld, long long n, const MapCoef& mc, cudaStream_t st) {
    k_map_apply<<<sweep_grid(n, 148 * 3), kBThreads, 0, st>>>(r, ld, n, mc);
}

// out: 18 doubles (q == nullptr) or 19 (q given: out[18] = sum q), device memory
void launch_moments(const double* r, long long ld, long long n, const double* q, ReduceState rs, double* out,
                    cudaStream_t st) {
    if (q) k_moments_1<7><<<sweep_grid(n, rs.max_blocks), kBThreads, 0, st>>>(r, ld, n, q, rs, out);
    else k_moments_1<6><<<sweep_grid(n, rs.max_blocks), kBThreads, 0, st>>>(r, ld, n, q, rs, out);
    k_moments_2<<<sweep_grid(n, 148 * 3), kBThreads, 0, st>>>(r, ld, n, rs, out);
}

// ---- f4: aperture cut + stream compaction ---------------------------------------------------
// RectAperture / EllipticalAperture (physics_proc.py:341-390) on a resident bunch: particles outside the aperture
// are removed and the survivors keep their order (ParticleArray.delete_particles, beam/particle.py:323-333).
// Three small kernels: per-tile survivor counts, one-block exclusive scan, ordered scatter of the six rows, the
// charges and the particle ids into a second buffer (in-place compaction would let one tile overwrite rows another
// tile has not read yet); the ids of the lost particles are written in order as well (lost-particle recorder).
//   kind 0: lost if v < lo || v > hi with v = row `row`            (RectAperture, one plane per call)
//   kind 1: lost if (x-dx)^2/ax^2 + (y-dy)^2/ay^2 > 1              (EllipticalAperture)
constexpr int kCutTile = 1024;
__device__ __forceinline__ bool cut_lost(const double* __restrict__ r, long long ld, long long i, CutSpec c) {
    if (c.kind == 0) {
        const double v = r[c.row * ld + i];
        return v < c.a || v > c.b;
    }
    const double x = r[i] - c.c, y = r[2 * ld + i] - c.d;
    return (x * x) / (c.a * c.a) + (y * y) / (c.b * c.b) > 1.0;                  // physics_proc.py:388
}
__global__ void __launch_bounds__(256) k_cut_count(const double* __restrict__ r, long long ld, long long n, CutSpec c,
                                                  int* __restrict__ counts) {
    __shared__ int sh[8];
    const long long base = (long long)blockIdx.x * kCutTile;
    int keep = 0;
#pragma unroll
    for (int u = 0; u < kCutTile / 256; ++u) {
        const long long i = base + u * 256 + threadIdx.x;
        if (i < n && !cut_lost(r, ld, i, c)) ++keep;
    }
    for (int o = 16; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = keep;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < 8; ++w) t += sh[w];
        counts[blockIdx.x] = t;
    }
}
// exclusive scan of the tile counts in place; counts[tiles] = total survivors
__global__ void __launch_bounds__(1024) k_cut_scan(int* __restrict__ counts, int tiles, long long* __restrict__ n_out) {
    __shared__ int sh[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int b0 = 0; b0 < tiles; b0 += 1024) {
        const int b = b0 + threadIdx.x;
        const int v = b < tiles ? counts[b] : 0;
        int x = v;
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) sh[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int w = sh[threadIdx.x];
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (threadIdx.x >= o) w += y; }
            sh[threadIdx.x] = w;
        }
        __syncthreads();
        const int warp_off = (threadIdx.x >> 5) ? sh[(threadIdx.x >> 5) - 1] : 0;
        if (b < tiles) counts[b] = carry + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += warp_off + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[tiles] = carry; *n_out = carry; }
}
__global__ void __launch_bounds__(256) k_cut_scatter(const double* __restrict__ r, long long ld, const double* __restrict__ q,
                                                    const long long* __restrict__ ids, long long n, CutSpec c,
                                                    const int* __restrict__ offsets, double* __restrict__ r_out,
                                                    long long ld_out, double* __restrict__ q_out,
                                                    long long* __restrict__ ids_out, long long* __restrict__ lost_out) {
    __shared__ int sh[8];
    const long long base = (long long)blockIdx.x * kCutTile;
    long long keep_pos = offsets[blockIdx.x];
    long long lost_pos = base - keep_pos;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < kCutTile / 256; ++u) {
        const long long i = base + u * 256 + threadIdx.x;
        const bool in = i < n;
        const bool keep = in && !cut_lost(r, ld, i, c);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) sh[warp] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < warp) before += sh[w]; total += sh[w]; }
        const int rank_in = before + __popc(m & ((1u << lane) - 1));       // survivors ahead of this one in the chunk
        const int idx_in = u * 256 + threadIdx.x;                          // position in the tile
        if (keep) {
            const long long o = keep_pos + rank_in;
#pragma unroll
            for (int k = 0; k < 6; ++k) r_out[k * ld_out + o] = r[k * ld + i];
            q_out[o] = q[i];
            if (ids_out) ids_out[o] = ids[i];
        } else if (in && lost_out) {
            lost_out[lost_pos + (idx_in - u * 256) - rank_in] = ids ? ids[i] : i;
        }
        const long long chunk = (n - (base + u * 256)) < 256 ? (n - (base + u * 256)) : 256;
        keep_pos += total;
        lost_pos += (chunk > 0 ? chunk : 0) - total;
        __syncthreads();
    }
}
void launch_cut(const double* r, long long ld, const double* q, const long long* ids, long long n, CutSpec c, int* counts,
                long long* n_out, double* r_out, long long ld_out, double* q_out, long long* ids_out, long long* lost_out,
                cudaStream_t st) {
    const int tiles = (int)((n + kCutTile - 1) / kCutTile);
    k_cut_count<<<tiles, 256, 0, st>>>(r, ld, n, c, counts);
    k_cut_scan<<<1, 1024, 0, st>>>(counts, tiles, n_out);
    k_cut_scatter<<<tiles, 256, 0, st>>>(r, ld, q, ids, n, c, counts, r_out, ld_out, q_out, ids_out, lost_out);
}

}  // namespace ocl
