// sc_lsc.cu -- longitudinal space charge (LSC), the 1-D sibling of the 3-D kick
// (reference: class LSC, ocelot/cpbd/sc.py:261-599).
//
// One LSC.apply is three sweeps over the particles and a handful of O(nb^2) kernels on the
// 1-D current profile (nb ~ 400 grid points for the default smoothing):
//
//   k_lsc_stats     sweep A  tau, x, y, q   -> n, sum/sum^2 of tau, min/max tau, sum q, sum x, sum y
//                            (np.mean/np.std/np.sum of sc.py:576-577,590; grid extent analysis.py:296-297)
//       -- the only host synchronisation: the host derives the grid (a, ds, nb) exactly like
//          s_to_cur (analysis.py:293-321) from these scalars --
//   k_lsc_deposit   sweep B  tau, x, y      -> CIC counts on the grid (s2cur_auxil, analysis.py:254-260)
//                                              + transverse size of the central slice (sc.py:579-589)
//   k_lsc_profile            Gaussian smoothing + normalisation (analysis.py:329-339)
//   k_lsc_spectrum           bunch spectrum x impedance (sc.py:453-469; imp_lsc :299-340,
//                            imp_step_lsc :342-369)
//   k_lsc_wake               inverse real transform -> wake on the grid (sc.py:401-416, :471-473)
//   k_lsc_kick      sweep C  tau in, delta in/out: np.interp of the wake, energy kick (sc.py:594-599)
//
// Deposit arithmetic: one native 64-bit integer reduction at the L2 (RED.E.ADD.64) per particle.
// A particle in cell i with fractional position f adds 1 - f to C[i] and f to C[i+1]; instead of two
// weights it adds ONE packed word (1 << cshift) + round(f 2^fbits) to counter i: the high field counts
// the particles of the cell (N_i), the low field accumulates their fractions (F_i), and
// C[i] = N_i - F_i + F_{i-1} is formed when the counters are folded (k_lsc_compact).  Measured on the
// way here (10^6 particles, ~400 bins): shared-memory histograms need 64-bit shared atomics, which
// are CAS spin loops on sm_100a (ATOMS.CAST.SPIN.64): 45 us; two L2 reductions per particle on 16
// replicas of the histogram: 60 us (operations on one address serialise at the L2, ~20 cycles each);
// on 2048 replicas: 9e10 atomics/s.  Hence R ~ the number of warps in flight (replica = global warp
// index mod R) and one reduction per particle.
// Integer addition is associative, so the counts are bit-reproducible from run to run and across any
// particle sharding, and sum(C) == n exactly.  Field widths adapt to the particles per replica:
// fbits >= 28 (400 M particles per GPU), 36 at 12.5 M, capped at the output scale 2^-s,
// s = 62 - ceil(log2 n_total) <= 52.
//
// The transforms are direct DFTs with an exact twiddle table (sincospi): the grid length nb is
// data dependent and not a power of two, and at nb ~ 400 the whole 1-D solve is ~10 us.
#include <cstdlib>

#include "sc_kernels.h"
#include "sc_special.h"

namespace ocl {

constexpr int kLscThreads = kSweepThreads;
constexpr int kLscDepth = 3;
constexpr int kLscSmemBins = 3072;        // wake table in shared memory up to 24 KB (static + dynamic < 48 KB)

// ---------------------------------------------------------------------------
// sweep A
// ---------------------------------------------------------------------------
// out[0..7] = max tau, max -tau, n, sum(tau - t0), sum((tau - t0)^2), sum q, sum x, sum y ; out[8] = t0
__global__ void __launch_bounds__(kLscThreads, 4) k_lsc_stats(const double* __restrict__ r, long long ld,
                                                             const double* __restrict__ q, long long n,
                                                             double* part, unsigned int* ticket,
                                                             double* __restrict__ out) {
    __shared__ double sh[8 * kSweepWarps];
    __shared__ double pipe[kLscDepth * 4 * kLscThreads];
    const double t0 = __ldg(r + 4 * ld);                         // any particle of the bunch: |tau - t0| ~ sigma
    double v[8] = {-INFINITY, -INFINITY, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double* const base[4] = {r + 4 * ld, r, r + 2 * ld, q};
    pipelined_sweep<4, kLscDepth>(base, (int)n, pipe, [&](int, const double (&w)[4]) {
        const double d = w[0] - t0;
        v[0] = fmax(v[0], w[0]); v[1] = fmax(v[1], -w[0]);
        v[2] += 1.0; v[3] += d; v[4] += d * d; v[5] += w[3]; v[6] += w[1]; v[7] += w[2];
    });
    if (grid_reduce<8, 2>(v, part, ticket, sh) && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) out[k] = v[k];
        out[8] = t0;
    }
}

// ---------------------------------------------------------------------------
// sweep B
// ---------------------------------------------------------------------------
// slice[0..3] = max x, max -x, max y, max -y ; slice[4..8] = count, sum dx, sum dx^2, sum dy, sum dy^2
__global__ void __launch_bounds__(kLscThreads, 3) k_lsc_deposit(const double* __restrict__ r, long long ld,
                                                               long long n, LP lpk, PKP pkk,
                                                               double* part, unsigned int* ticket,
                                                               unsigned long long* __restrict__ spread,
                                                               double* __restrict__ slice) {
    __shared__ double sh[9 * kSweepWarps];
    __shared__ double pipe[kLscDepth * 3 * kLscThreads];
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const LscPack pk = pkk.p ? *pkk.p : pkk.v;
    // replica-major layout: counter (replica, j) at word replica * nb + j; one replica per warp in flight
    unsigned long long* const dst =
        spread + (size_t)((blockIdx.x * kSweepWarps + (threadIdx.x >> 5)) & (pk.replicas - 1)) * lp.nb;
    const unsigned long long one = 1ull << pk.cshift, fmax_ = 1ull << pk.fbits;
    const double scale = (double)fmax_;
    double v[9] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double* const base[3] = {r + 4 * ld, r, r + 2 * ld};
    pipelined_sweep<3, kLscDepth>(base, (int)n, pipe, [&](int, const double (&w)[3]) {
        const double tau = w[0];
        const double cA = (tau - lp.a) / lp.ds;                      // analysis.py:324
        const double fl = floor(cA);
        const double xi = (1.0 + fl) - cA;                           // :326
        long long i = (long long)fl;
        if (i > lp.nb - 1) i = lp.nb - 1;                            // :257-258
        if (i >= 0) {
            unsigned long long f = (unsigned long long)((1.0 - xi) * scale + 0.5);
            if (f > fmax_) f = fmax_;
            // C[i] += xi, C[i+1] += 1 - xi (:259-260) as one packed reduction, see the file header
            asm volatile("red.relaxed.gpu.global.add.u64 [%0], %1;" ::"l"(dst + i), "l"(one + f) : "memory");
        }
        if (tau >= lp.slice_min && tau < lp.slice_max) {             // sc.py:582
            const double dx = w[1] - lp.x_shift, dy = w[2] - lp.y_shift;
            v[0] = fmax(v[0], w[1]); v[1] = fmax(v[1], -w[1]); v[2] = fmax(v[2], w[2]); v[3] = fmax(v[3], -w[2]);
            v[4] += 1.0; v[5] += dx; v[6] += dx * dx; v[7] += dy; v[8] += dy * dy;
        }
    });
    if (grid_reduce<9, 4>(v, part, ticket, sh) && threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 9; ++k) slice[k] = v[k];
    }
}

// fold the replicas: bins[j] += sum over replicas of (N_j << s) - (F_j << d) + (F_{j-1} << d), d = s - fbits
// (bins zeroed by the caller).  Tiles of 32 bins x 64 replicas, block (32, 8): coalesced rows of 32 bins;
// the blocks stride over the tiles, so one launch shape serves any (nb, replicas).
__global__ void __launch_bounds__(256) k_lsc_compact(const unsigned long long* __restrict__ spread, LP lpk, PKP pkk,
                                                    unsigned long long* __restrict__ bins) {
    __shared__ unsigned long long sm[8][32];
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const LscPack pk = pkk.p ? *pkk.p : pkk.v;
    const int nb = lp.nb, s_out = lp.fx_shift;
    const unsigned long long fmask = (1ull << pk.cshift) - 1ull;
    const int d = s_out - pk.fbits;
    const int tiles_x = (nb + 31) / 32, tiles_y = (pk.replicas + 63) / 64;
    for (int tile = blockIdx.x; tile < tiles_x * tiles_y; tile += gridDim.x) {
        const int j = (tile % tiles_x) * 32 + threadIdx.x;
        const int rep0 = (tile / tiles_x) * 64;
        unsigned long long c = 0ull;                                // modulo 2^64; the total is non-negative
        if (j < nb) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int rep = rep0 + k * 8 + threadIdx.y;
                if (rep < pk.replicas) {
                    const unsigned long long* p = spread + (size_t)rep * nb + j;
                    const unsigned long long v = p[0];
                    const unsigned long long vm = (j > 0) ? p[-1] : 0ull;
                    c += ((v >> pk.cshift) << s_out) - ((v & fmask) << d) + ((vm & fmask) << d);
                }
            }
        }
        __syncthreads();
        sm[threadIdx.y][threadIdx.x] = c;
        __syncthreads();
        if (threadIdx.y == 0 && j < nb) {
#pragma unroll
            for (int k = 1; k < 8; ++k) c += sm[k][threadIdx.x];
            if (c) atomicAdd(bins + j, c);
        }
    }
}

// ---------------------------------------------------------------------------
// 1-D profile: counts -> smoothed, normalised line density
// ---------------------------------------------------------------------------
// prof[j] = bunch[j] * c  (the array the reference feeds to signal_to_spectrum, sc.py:461-463)
// cur[j]  = I(s_j) [A]    (B[:, 1] of s_to_cur; tap)
// Also derives the transverse size from the slice sums.
__global__ void __launch_bounds__(1024, 1) k_lsc_profile(const unsigned long long* __restrict__ bins, LP lpk,
                                                        const double* __restrict__ slice, double* __restrict__ cnt,
                                                        double* __restrict__ prof, double* __restrict__ cur,
                                                        double* __restrict__ sigma_out) {
    __shared__ double red[32];
    __shared__ double total;
    extern __shared__ double G[];                                   // 2K+1 taps
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const int nb = lp.nb, K = lp.K;
    const double inv_unit = 1.0 / (double)(1ull << lp.fx_shift);
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        cnt[j] = (double)bins[j] * inv_unit;
    }
    auto block_sum = [&](double x) -> double {
        x = warp_reduce(x, OpSum());
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            double y = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
            y = warp_reduce(y, OpSum());
            if (threadIdx.x == 0) total = y;
        }
        __syncthreads();
        return total;
    };
    if (K >= 0) {                                                   // analysis.py:330-333
        double g = 0.0;
        for (int u = threadIdx.x; u <= 2 * K; u += blockDim.x) {
            const double t = (double)(u - K) * lp.ds / lp.sigma_s;
            G[u] = exp(-0.5 * (t * t));
            g += G[u];
        }
        const double gs = block_sum(g);
        for (int u = threadIdx.x; u <= 2 * K; u += blockDim.x) G[u] = G[u] / gs;
    }
    __syncthreads();
    double s = 0.0;
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        double b;
        if (K >= 0) {                                               // convmode(C, G, 1): beam_utils.py:63-67
            b = 0.0;
            const int u0 = max(0, j + K - (nb - 1)), u1 = min(2 * K, j + K);
            for (int u = u0; u <= u1; ++u) b += G[u] * cnt[j + K - u];
        } else {
            b = cnt[j];
        }
        prof[j] = b;
        s += b;
    }
    const double sumB = block_sum(s);
    const double koef = lp.q * lp.v / (lp.ds * sumB);               // analysis.py:338
    for (int j = threadIdx.x; j < nb; j += blockDim.x) {
        const double I = koef * prof[j];
        cur[j] = I;
        prof[j] = I / (lp.q * kSpeedOfLight) * kSpeedOfLight;       // sc.py:591, :463
    }
    if (threadIdx.x == 0) {
        double sigma;
        if (lp.step_profile) {                                      // sc.py:584-587
            sigma = fmin(slice[0] + slice[1], slice[2] + slice[3]) / 2;
        } else {                                                    // sc.py:588-589 (np.std: population)
            const double c = slice[4];
            const double mx = slice[5] / c, my = slice[7] / c;
            double vx = slice[6] / c - mx * mx, vy = slice[8] / c - my * my;
            if (vx < 0.0) vx = 0.0;                                 // (keeps the NaN of an empty slice, like np.std)
            if (vy < 0.0) vy = 0.0;
            sigma = (sqrt(vx) + sqrt(vy)) / 2.;
        }
        *sigma_out = sigma;
    }
}

// twiddle table tw[m] = exp(+2 pi i m / n), m = 0..n-1
__global__ void k_lsc_twiddles(int n_host, const LscParams* __restrict__ dp, double2* __restrict__ tw) {
    const int n = dp ? 2 * dp->nb : n_host;
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= n) return;
    double s, c;
    sincospi(2.0 * (double)m / (double)n, &s, &c);
    tw[m] = make_double2(c, s);
}

// Za(w_k) = i A[k]: impedance of the step at the grid frequencies (imp_lsc sc.py:299-340,
// imp_step_lsc :342-369, undulator factor :460-463).  One warp per frequency: the lanes share the
// quadrature of K1.
__global__ void __launch_bounds__(256) k_lsc_impedance(LP lpk, const double* __restrict__ sigma_p,
                                                      double* __restrict__ A) {
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= lp.nb) return;
    const int n = 2 * lp.nb;
    const double dt = lp.ds / kSpeedOfLight;
    const double f = 1 / dt * (double)k / (double)n;                // sc.py:457
    double w = f * 2 * kPi;
    const double sigma = *sigma_p;
    const double z0 = 1. / (kSpeedOfLight * kEpsilon0);             // globals.py:36
    double a;
    if (lp.step_profile) {                                          // sc.py:363-368
        const bool low = w < 1e-7;
        if (low) w = 1e-7;
        const double x = w * sigma / (kSpeedOfLight * lp.gamma);
        double k1;
        if (!(x > 0.0)) k1 = INFINITY;
        else if (x > 745.0) k1 = 0.0;
        else k1 = bessel_k1_step(x) * exp(-x) * warp_reduce(bessel_k1_partial(x, lane, 32), OpSum());
        a = z0 * kSpeedOfLight / (4 * w * sigma * sigma) * lp.dz * (1 - x * k1);
        if (low) a = 0.0;
    } else {                                                        // sc.py:320-339
        const double alpha = w * sigma / (lp.gamma * kSpeedOfLight);
        const double a2 = alpha * alpha;
        double T = 0.0;
        if (a2 > 40.0) {
            double fact = 1.0, p = a2, sgn = 1.0;
            for (int i = 0; i < 10; ++i) {                          // sum (-1)^i i! / x^(i+1)
                if (i > 0) fact *= (double)i;
                T += sgn * fact / p;
                p *= a2; sgn = -sgn;
            }
        } else if (a2 >= 1e-16) {
            T = exp_e1(a2);
        }
        a = z0 / (4 * kPi * kSpeedOfLight * lp.gamma * lp.gamma) * w * T * lp.dz;
    }
    if (lane == 0) A[k] = a * lp.und;
}

// Z[k] = Za(w_k) * Zb[k], k = 0..nb-1, with Zb = dt * fft(prof, n)  (sc.py:453-469, analysis.py:376-383)
// One warp per frequency, lanes stride over the grid points.
__global__ void __launch_bounds__(256) k_lsc_spectrum(const double* __restrict__ prof, LP lpk,
                                                     const double* __restrict__ A,
                                                     const double2* __restrict__ tw, double2* __restrict__ Z) {
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const int lane = threadIdx.x & 31;
    const int nb = lp.nb, n = 2 * nb;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (k >= nb) return;
    double re = 0.0, im = 0.0;
    int m = (int)(((long long)lane * k) % n);                       // (j * k) mod n
    const int step = (int)((32LL * k) % n);
    for (int j = lane; j < nb; j += 32) {
        const double2 t = __ldg(tw + m);
        const double p = __ldg(prof + j);
        re += p * t.x;                                              // exp(-2 pi i j k / n)
        im -= p * t.y;
        m += step;
        if (m >= n) m -= n;
    }
    re = warp_reduce(re, OpSum());
    im = warp_reduce(im, OpSum());
    if (lane) return;
    const double dt = lp.ds / kSpeedOfLight;
    const double a = A[k];
    Z[k] = make_double2(-a * (im * dt), a * (re * dt));             // i A (re + i im)
}

// W[j] = q / dt * irfft(Z, n)[j], j = 0..nb-1, with Z[nb] = conj(Z[nb-1]) (sc.py:466-473, :401-416, :592)
// One warp per grid point, lanes stride over the frequencies.
__global__ void __launch_bounds__(256) k_lsc_wake(const double2* __restrict__ Z, LP lpk,
                                                 const double2* __restrict__ tw, double* __restrict__ W) {
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    const int lane = threadIdx.x & 31;
    const int nb = lp.nb, n = 2 * nb;
    const int j = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (j >= nb) return;
    double acc = 0.0;
    int m = (int)(((long long)lane * j) % n);
    const int step = (int)((32LL * j) % n);
    for (int k = lane; k < nb; k += 32) {
        const double2 t = __ldg(tw + m);
        const double2 z = __ldg(Z + k);
        const double term = z.x * t.x - z.y * t.y;                  // Re(Z_k exp(+2 pi i j k / n))
        acc += (k == 0) ? 0.5 * term : term;
        m += step;
        if (m >= n) m -= n;
    }
    acc = warp_reduce(acc, OpSum());
    if (lane) return;
    const double nyq = Z[nb - 1].x;                                 // Re conj(Z[nb-1])
    acc = 2.0 * acc + ((j & 1) ? -nyq : nyq);
    const double dt = lp.ds / kSpeedOfLight;
    W[j] = acc / (double)n / dt * lp.q;
}

// ---------------------------------------------------------------------------
// sweep C: delta += interp(tau, x, W) * 1e-9 / pc_ref     (sc.py:594-599)
// ---------------------------------------------------------------------------
template <bool SMEM>
__global__ void __launch_bounds__(kLscThreads, 4) k_lsc_kick(double* __restrict__ r, long long ld, long long n,
                                                            LP lpk, const double* __restrict__ Wg) {
    __shared__ double pipe[kLscDepth * 2 * kLscThreads];
    extern __shared__ double Ws[];
    const LscParams lp = lpk.p ? *lpk.p : lpk.v;
    if (lp.nb < 2) return;                                           // rejected grid (asynchronous form): no kick
    if (SMEM) {
        for (int j = threadIdx.x; j < lp.nb; j += kLscThreads) Ws[j] = Wg[j];
        __syncthreads();
    }
    const double* const W = SMEM ? Ws : Wg;
    const int nb = lp.nb;
    auto xg = [&](int j) { return __dadd_rn(__dmul_rn((double)j, lp.ds), lp.a); };   // np.arange(...)+a, analysis.py:321
    const double x_last = xg(nb - 1);
    double* const delta = r + 5 * ld;
    const double* const base[2] = {r + 4 * ld, r + 5 * ld};
    pipelined_sweep<2, kLscDepth>(base, (int)n, pipe, [&](int i, const double (&w)[2]) {
        const double tau = w[0];
        double dE;
        if (tau >= x_last) dE = W[nb - 1];                           // np.interp right clamp / last node
        else if (tau < lp.a) dE = W[0];                              // left clamp (x_0 = a)
        else {
            int j = (int)floor((tau - lp.a) / lp.ds);
            j = max(0, min(j, nb - 2));
            double xj = xg(j);
            if (tau < xj) { --j; xj = xg(j); }                       // rounding of j*ds + a
            else if (tau >= xg(j + 1)) { ++j; xj = xg(j); }
            const double slope = (W[j + 1] - W[j]) / (xg(j + 1) - xj);
            dE = (tau == xj) ? W[j] : __dadd_rn(__dmul_rn(slope, tau - xj), W[j]);
        }
        delta[i] = w[1] + dE * 1e-9 / lp.pc_ref;
    });
}

// ---------------------------------------------------------------------------
// asynchronous form: the grid definition on the device
// ---------------------------------------------------------------------------
// What ocelot_b200/lsc.py::kick_parameters does on the host after ocl_sc_lsc_stats, with the same
// IEEE operations in the same order (explicitly rounded, so nothing is contracted into an FMA):
// mean / sigma of tau (sc.py:576-580), the grid of s_to_cur (analysis.py:293-321), its smoothing taps
// (:330-331), the packed-word layout of the deposit.  One thread.
__global__ void k_lsc_params(const double* __restrict__ raw, LscHost hp, LscParams* __restrict__ out,
                             LscPack* __restrict__ pk, int* __restrict__ err) {
    if (threadIdx.x || blockIdx.x) return;
    const double cnt = raw[2], s1 = raw[3], s2 = raw[4], t0 = raw[8];
    const double mean = __dadd_rn(t0, __ddiv_rn(s1, cnt));
    double m2 = __dsub_rn(s2, __ddiv_rn(__dmul_rn(s1, s1), cnt));
    if (m2 < 0.0) m2 = 0.0;
    const double sigma_tau = __dsqrt_rn(__ddiv_rn(m2, cnt));
    const double tmin = -raw[1], tmax = raw[0];
    const double sigma = __dmul_rn(sigma_tau, hp.smooth_param);
    const double a = __dsub_rn(tmin, __dmul_rn(3.0, sigma));
    const double b = __dadd_rn(tmax, __dmul_rn(3.0, sigma));
    double ds = sigma > 0.0 ? __dmul_rn(0.25, sigma) : __ddiv_rn(__dsub_rn(b, a), 1000.0);
    const double span = __dsub_rn(b, a);
    const double Nf = ceil(__ddiv_rn(span, ds));
    LscParams lp;
    int bad = 0;
    if (!(Nf >= 1.0) || Nf + 1.0 > (double)hp.cap_nb) { bad = 1; lp.nb = 0; ds = 1.0; }
    else { lp.nb = (int)Nf + 1; ds = __ddiv_rn(span, Nf); }
    lp.a = a; lp.ds = ds; lp.sigma_s = sigma;
    lp.K = sigma > 0.0 ? (int)floor(__dadd_rn(__ddiv_rn(__dmul_rn(3.0, sigma), ds), 0.5)) : -1;
    if (lp.K > 63) { bad = 1; lp.nb = 0; lp.K = -1; }
    lp.slice_min = __dadd_rn(mean, __dmul_rn(sigma_tau, hp.bound_lo));
    lp.slice_max = __dadd_rn(mean, __dmul_rn(sigma_tau, hp.bound_hi));
    lp.x_shift = __ddiv_rn(raw[6], cnt); lp.y_shift = __ddiv_rn(raw[7], cnt);
    lp.q = raw[5]; lp.v = hp.v; lp.gamma = hp.gamma; lp.dz = hp.dz; lp.und = hp.und; lp.pc_ref = hp.pc_ref;
    lp.step_profile = hp.step_profile; lp.fx_shift = hp.fx_shift;
    // packed deposit word (see launch_lsc_deposit)
    LscPack p;
    int rep = 2048;
    while (rep > 1 && (long long)lp.nb * rep > (1 << 20)) rep >>= 1;
    p.replicas = rep;
    const long long share = (hp.warps + rep - 1) / rep;
    const long long per_replica = share * 32 * hp.iters;
    int cbits = 1;
    while ((1ll << cbits) <= per_replica) ++cbits;
    p.cshift = 64 - cbits;
    p.fbits = p.cshift - cbits;
    if (p.fbits > lp.fx_shift) p.fbits = lp.fx_shift;
    if (p.fbits < 24) { bad = 2; lp.nb = 0; p.fbits = 24; }
    *out = lp;
    *pk = p;
    if (bad) *err = bad;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
static int lsc_grid(long long n, int cap) {
    long long b = (n + kLscThreads - 1) / kLscThreads;
    if (b < 1) b = 1;
    if (b > cap) b = cap;
    return (int)b;
}

void launch_lsc_stats(const double* r, long long ld, const double* q, long long n, LscWork w, cudaStream_t st) {
    k_lsc_stats<<<lsc_grid(n, w.max_blocks), kLscThreads, 0, st>>>(r, ld, q, n, w.part, w.ticket, w.stats);
}

void launch_lsc_twiddles(int nb, LscWork w, cudaStream_t st) {
    const int n = 2 * nb;
    k_lsc_twiddles<<<(n + 255) / 256, 256, 0, st>>>(n, nullptr, w.tw);
}

int lsc_replicas(int nb) {
    static int forced = -1;
    if (forced < 0) {
        const char* e = getenv("OCL_LSC_REPLICAS");               // A/B experiments
        forced = e ? atoi(e) : 0;
    }
    int r = forced > 0 ? forced : 2048;                            // ~ warps in flight (148 SMs x 2-3 blocks x 8 warps)
    while (r > 1 && (long long)nb * r > (1 << 20)) r >>= 1;       // <= 8 MB of counters
    return r;
}

long long lsc_spread_words(int nb) { return (long long)nb * lsc_replicas(nb); }

// returns 1 (nothing launched) if the packed word cannot hold 24 fractional bits: a grid so fine that the
// histogram has few replicas, combined with so many particles per replica that the count field crowds
// out the fraction (e.g. nb = 2^20 with 4e8 particles)
int launch_lsc_deposit(const double* r, long long ld, long long n, const LscParams& lp, LscWork w,
                       cudaStream_t st) {
    const int grid = lsc_grid(n, w.max_blocks);
    LscPack pk;
    pk.replicas = lsc_replicas(lp.nb);
    // particles one replica can receive: warps sharing it x iterations per thread x 32 lanes
    const long long warps = (long long)grid * kSweepWarps;
    const long long share = (warps + pk.replicas - 1) / pk.replicas;
    const long long per_replica = share * 32 * ((n + (long long)grid * kLscThreads - 1) / ((long long)grid * kLscThreads));
    int cbits = 1;
    while ((1ll << cbits) <= per_replica) ++cbits;
    pk.cshift = 64 - cbits;
    pk.fbits = pk.cshift - cbits;                                    // F_i < 2^cshift for any fill
    if (pk.fbits > lp.fx_shift) pk.fbits = lp.fx_shift;
    if (pk.fbits < 24) return 1;
    cudaMemsetAsync(w.spread, 0, sizeof(unsigned long long) * (size_t)lp.nb * pk.replicas, st);
    cudaMemsetAsync(w.bins, 0, sizeof(unsigned long long) * lp.nb, st);
    const LP lpk = {lp, nullptr};
    const PKP pkk = {pk, nullptr};
    k_lsc_deposit<<<grid, kLscThreads, 0, st>>>(r, ld, n, lpk, pkk, w.part, w.ticket + 1, w.spread, w.slice);
    const int tiles = ((lp.nb + 31) / 32) * ((pk.replicas + 63) / 64);
    k_lsc_compact<<<tiles, dim3(32, 8), 0, st>>>(w.spread, lpk, pkk, w.bins);
    return 0;
}

int launch_lsc_solve(const LscParams& lp, LscWork w, cudaStream_t st) {
    const size_t taps = lp.K >= 0 ? sizeof(double) * (2 * (size_t)lp.K + 1) : 0;
    if (taps > 40 * 1024) return 1;                                  // K <= 2559
    const LP lpk = {lp, nullptr};
    k_lsc_profile<<<1, 1024, taps, st>>>(w.bins, lpk, w.slice, w.cnt, w.prof, w.cur, w.sigma);
    const int blocks = (lp.nb + 7) / 8;
    k_lsc_impedance<<<blocks, 256, 0, st>>>(lpk, w.sigma, w.A);
    k_lsc_spectrum<<<blocks, 256, 0, st>>>(w.prof, lpk, w.A, w.tw, w.Z);
    k_lsc_wake<<<blocks, 256, 0, st>>>(w.Z, lpk, w.tw, w.W);
    return 0;
}

void launch_lsc_kick(double* r, long long ld, long long n, const LscParams& lp, LscWork w, cudaStream_t st) {
    const int grid = lsc_grid(n, 148 * 4);
    if (lp.nb <= kLscSmemBins)
        k_lsc_kick<true><<<grid, kLscThreads, sizeof(double) * lp.nb, st>>>(r, ld, n, LP{lp, nullptr}, w.W);
    else
        k_lsc_kick<false><<<grid, kLscThreads, 0, st>>>(r, ld, n, LP{lp, nullptr}, w.W);
}

// One LSC.apply without a host round trip: sweep A, k_lsc_params, then the same kernels reading the
// device-resident scalars.  Launch shapes are those of the largest grid the buffers hold
// (hp.cap_nb <= kLscAsyncCap); blocks beyond the actual grid exit at once.
void launch_lsc_kick_async(double* r, long long ld, const double* q, long long n, LscHost hp, LscWork w,
                           cudaStream_t st) {
    const int grid = lsc_grid(n, w.max_blocks);
    hp.warps = (long long)grid * kSweepWarps;
    hp.iters = (n + (long long)grid * kLscThreads - 1) / ((long long)grid * kLscThreads);
    const int cap = hp.cap_nb;
    k_lsc_stats<<<grid, kLscThreads, 0, st>>>(r, ld, q, n, w.part, w.ticket, w.stats);
    k_lsc_params<<<1, 32, 0, st>>>(w.stats, hp, w.dparams, w.dpack, w.err);
    const LP lpk = {LscParams{}, w.dparams};
    const PKP pkk = {LscPack{}, w.dpack};
    cudaMemsetAsync(w.spread, 0, sizeof(unsigned long long) * (size_t)(1 << 20), st);
    cudaMemsetAsync(w.bins, 0, sizeof(unsigned long long) * cap, st);
    k_lsc_deposit<<<grid, kLscThreads, 0, st>>>(r, ld, n, lpk, pkk, w.part, w.ticket + 1, w.spread, w.slice);
    k_lsc_compact<<<512, dim3(32, 8), 0, st>>>(w.spread, lpk, pkk, w.bins);
    k_lsc_twiddles<<<(2 * cap + 255) / 256, 256, 0, st>>>(0, w.dparams, w.tw);
    k_lsc_profile<<<1, 1024, sizeof(double) * 127, st>>>(w.bins, lpk, w.slice, w.cnt, w.prof, w.cur, w.sigma);
    const int blocks = (cap + 7) / 8;
    k_lsc_impedance<<<blocks, 256, 0, st>>>(lpk, w.sigma, w.A);
    k_lsc_spectrum<<<blocks, 256, 0, st>>>(w.prof, lpk, w.A, w.tw, w.Z);
    k_lsc_wake<<<blocks, 256, 0, st>>>(w.Z, lpk, w.tw, w.W);
    k_lsc_kick<false><<<lsc_grid(n, 148 * 4), kLscThreads, 0, st>>>(r, ld, n, lpk, w.W);
}

}  // namespace ocl
