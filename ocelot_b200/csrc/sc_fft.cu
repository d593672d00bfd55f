// sc_fft.cu -- hand-written Hockney convolution for the space-charge potential
// (replaces SpaceCharge.potential, ocelot/cpbd/sc.py:135-168).
//
// The reference zero-pads rho to (2n-1)^3, mirrors the integrated Green's
// function K into the same box and multiplies two full complex FFTs.  The same
// linear convolution is computed here on the power-of-two box M^3 with every
// piece of structure exploited:
//
//   * rho is non-zero only in [0,n)^3 and phi is needed only there, so each 1-D
//     pass touches only the lines that can be non-zero / are needed
//     (z: n*n lines, y: n*(M/2+1), x: M*(M/2+1));
//   * K is real and even in all three axes, so K_hat is real and even: it is
//     built by three real-even 1-D passes on the n^3 octant and stored only for
//     0 <= k <= M/2 per axis ((M/2+1)^3 doubles instead of M^2(M/2+1) complex);
//   * two real lines ride through one complex FFT (real/imaginary packing);
//   * the x pass does forward FFT, multiply by K_hat, inverse FFT without
//     leaving shared memory, so rho_hat is never written to HBM;
//   * the 1/(Mx My Mz) of the inverse transform and the 1/(4 pi eps0 hx hy hz) of
//     sc.py:167 are applied in the last store.
//
// All 1-D transforms are Stockham radix-4/2 FFTs in shared memory (fp64,
// natural order in and out), LB lines per 256-thread block.
#include "sc_kernels.h"

namespace ocl {

constexpr int kFftThreads = 256;

// Shared-memory layout: complex point o of line l lives at x[o * NLP + l] with
// NLP = NL + 1.  Lines run across lanes, so every butterfly stage reads and writes
// whole rows (NL consecutive double2) whatever its stride -- no bank conflicts --
// and the odd pitch keeps the transposed accesses of the z passes (lanes along
// o) conflict-free per quarter-warp as well.
struct FftGeom {
    int M;     // transform length (power of two, 8..512)
    int NL;    // lines resident per block
    int NLP;   // row pitch (NL + 1)
};
__host__ __device__ inline FftGeom fft_geom(int M) {
    FftGeom g;
    g.M = M;
    g.NL = (M >= 256) ? 8 : 2048 / M;
    if (g.NL > 64) g.NL = 64;
    g.NLP = g.NL + 1;
    return g;
}
__host__ __device__ inline size_t fft_buf_elems(const FftGeom& g) { return (size_t)g.M * g.NLP; }

__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
__device__ __forceinline__ double2 rot90(double2 d) {
    return INV ? make_double2(-d.y, d.x) : make_double2(d.y, -d.x);
}

template <bool INV>
__device__ __forceinline__ void dft4(double2& v0, double2& v1, double2& v2, double2& v3) {
    double2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = rot90<INV>(csub(v1, v3));
    v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}

template <bool INV>
__device__ __forceinline__ void dft8(double2 (&v)[8]) {
    // decimation in time: evens and odds through dft4, then the w8^k twiddles
    dft4<INV>(v[0], v[2], v[4], v[6]);
    dft4<INV>(v[1], v[3], v[5], v[7]);
    const double h = 0.70710678118654752440;
    // w8^1 = (1 -+ i)/sqrt2, w8^2 = -+i, w8^3 = (-1 -+ i)/sqrt2   (upper sign: forward)
    double2 o1 = INV ? make_double2(h * (v[3].x - v[3].y), h * (v[3].x + v[3].y))
                     : make_double2(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
    double2 o2 = rot90<INV>(v[5]);
    double2 o3 = INV ? make_double2(-h * (v[7].x + v[7].y), h * (v[7].x - v[7].y))
                     : make_double2(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
    double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6], o0 = v[1];
    v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
    v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
    v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
    v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
}

struct FftSmem {
    double2* a;
    double2* b;
    double2* tw;
};

__device__ __forceinline__ FftSmem fft_smem(const FftGeom& g, const double2* __restrict__ tw_g) {
    extern __shared__ __align__(16) unsigned char raw[];
    FftSmem s;
    s.a = reinterpret_cast<double2*>(raw);
    s.b = s.a + fft_buf_elems(g);
    s.tw = s.b + fft_buf_elems(g);
    for (int t = threadIdx.x; t < g.M; t += kFftThreads) s.tw[t] = tw_g[t];
    return s;
}

// One Stockham stage of radix R on all NL lines: work item t -> (butterfly j, line l), l fastest.
template <bool INV, int R>
__device__ __forceinline__ void fft_stage(const double2* __restrict__ x, double2* __restrict__ y,
                                          const double2* __restrict__ tw, const FftGeom& g, int Ns) {
    const int nb = g.M / R;
    const int step = g.M / (Ns * R);
    const int total = nb * g.NL;
    for (int t = threadIdx.x; t < total; t += kFftThreads) {
        const int j = t / g.NL, l = t - j * g.NL;
        const int k = j & (Ns - 1);
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = x[(j + r * nb) * g.NLP + l];
        if (k) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double2 w = tw[r * k * step];
                if (INV) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        if (R == 8) {
            double2 u[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) u[r] = v[r % R];
            dft8<INV>(u);
#pragma unroll
            for (int r = 0; r < 8; ++r) v[r % R] = u[r];
        } else if (R == 4) {
            dft4<INV>(v[0], v[1 % R], v[2 % R], v[3 % R]);
        } else {
            double2 a = v[0], b = v[1 % R];
            v[0] = cadd(a, b); v[1 % R] = csub(a, b);
        }
        double2* dst = y + ((j - k) * R + k) * g.NLP + l;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * Ns * g.NLP] = v[r];
    }
}

// All NL lines of length M in x.  Returns the buffer holding the result (natural order).
// tw[m] = exp(-2 pi i m / M); the inverse transform conjugates it (unnormalised).
template <bool INV>
__device__ double2* block_fft(double2* x, double2* y, const double2* tw, const FftGeom& g) {
    __syncthreads();
    for (int Ns = 1; Ns < g.M;) {
        const int rem = g.M / Ns;
        int R;
        if ((rem & 7) == 0) { fft_stage<INV, 8>(x, y, tw, g, Ns); R = 8; }
        else if ((rem & 3) == 0) { fft_stage<INV, 4>(x, y, tw, g, Ns); R = 4; }
        else { fft_stage<INV, 2>(x, y, tw, g, Ns); R = 2; }
        __syncthreads();
        double2* tmp = x; x = y; y = tmp;
        Ns *= R;
    }
    return x;
}

__device__ __forceinline__ double green_entry_dev(const double* __restrict__ G, int gy, int gz, int i, int j, int k) {
    const size_t sx = (size_t)gy * gz, sy = gz;
    const double* lo = G + (size_t)i * sx + (size_t)j * sy + k;
    double v = __dsub_rn(__ldg(lo + sx + sy + 1), __ldg(lo + sy + 1));   // order of sc.py:128-131
    v = __dsub_rn(v, __ldg(lo + sx + 1));
    v = __dadd_rn(v, __ldg(lo + 1));
    v = __dsub_rn(v, __ldg(lo + sx + sy));
    v = __dadd_rn(v, __ldg(lo + sy));
    v = __dadd_rn(v, __ldg(lo + sx));
    v = __dsub_rn(v, __ldg(lo));
    return v;
}

// ---------------------------------------------------------------------------
// K_hat, pass z: lines (a,b) of K1 (8-corner difference of the antiderivative
// table, sc.py:128-131), even-extended to Mz, two lines per complex FFT.
// out P[a][b][kz], kz <= Mz/2 (real).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) k_khat_z(const double* __restrict__ gtab, MeshDims md,
                                                       const double2* __restrict__ tw_g, double* __restrict__ P) {
    const FftGeom g = fft_geom(md.mz);
    const int M = g.M, n = md.nz, H = M / 2;
    FftSmem s = fft_smem(g, tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * g.NL);            // 2 real lines per complex line
    for (int t = threadIdx.x; t < M * g.NLP; t += kFftThreads) s.a[t] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int t = threadIdx.x; t < g.NL * n; t += kFftThreads) {
        const int p = t / n, c = t - p * n;               // c fastest: contiguous table reads
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        double e1 = green_entry_dev(gtab, md.ny + 1, md.nz + 1, l1 / md.ny, l1 % md.ny, c);
        double e2 = (l2 < nlines) ? green_entry_dev(gtab, md.ny + 1, md.nz + 1, l2 / md.ny, l2 % md.ny, c) : 0.0;
        s.a[c * g.NLP + p] = make_double2(e1, e2);
        if (c) s.a[(M - c) * g.NLP + p] = make_double2(e1, e2);
    }
    double2* X = block_fft<false>(s.a, s.b, s.tw, g);
    for (int t = threadIdx.x; t < g.NL * (H + 1); t += kFftThreads) {
        const int p = t / (H + 1), kz = t - p * (H + 1);
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 v = X[kz * g.NLP + p];
        P[(size_t)l1 * (H + 1) + kz] = v.x;
        if (l2 < nlines) P[(size_t)l2 * (H + 1) + kz] = v.y;
    }
}

// ---------------------------------------------------------------------------
// K_hat, passes y and x: real-even transform along the OUTER axis of
// in[batch][n][inner] -> out[batch][M/2+1][inner]; two adjacent inner indices
// per complex FFT.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) k_real_even_outer(const double* __restrict__ in,
                                                                double* __restrict__ out, int n, int M, int inner,
                                                                const double2* __restrict__ tw_g) {
    const FftGeom g = fft_geom(M);
    const int H = M / 2;
    FftSmem s = fft_smem(g, tw_g);
    const int pairs_total = (inner + 1) / 2;
    const int blocks_per_batch = (pairs_total + g.NL - 1) / g.NL;
    const int batch = blockIdx.x / blocks_per_batch;
    const int pair0 = (blockIdx.x - batch * blocks_per_batch) * g.NL;
    const int pairs = min(g.NL, pairs_total - pair0);
    const double* src = in + (size_t)batch * n * inner;
    double* dst = out + (size_t)batch * (H + 1) * inner;
    for (int t = threadIdx.x; t < M * g.NLP; t += kFftThreads) s.a[t] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int t = threadIdx.x; t < g.NL * n; t += kFftThreads) {
        const int o = t / g.NL, p = t - o * g.NL;         // p fastest: adjacent inner indices
        if (p >= pairs) continue;
        const int f = 2 * (pair0 + p);
        const double e1 = __ldg(src + (size_t)o * inner + f);
        const double e2 = (f + 1 < inner) ? __ldg(src + (size_t)o * inner + f + 1) : 0.0;
        s.a[o * g.NLP + p] = make_double2(e1, e2);
        if (o) s.a[(M - o) * g.NLP + p] = make_double2(e1, e2);
    }
    double2* X = block_fft<false>(s.a, s.b, s.tw, g);
    for (int t = threadIdx.x; t < g.NL * (H + 1); t += kFftThreads) {
        const int ko = t / g.NL, p = t - ko * g.NL;
        if (p >= pairs) continue;
        const int f = 2 * (pair0 + p);
        const double2 v = X[ko * g.NLP + p];
        dst[(size_t)ko * inner + f] = v.x;
        if (f + 1 < inner) dst[(size_t)ko * inner + f + 1] = v.y;
    }
}

// ---------------------------------------------------------------------------
// rho, pass z: real lines rho[l][k<nz] zero-padded to Mz, two per complex FFT,
// separated by Hermitian symmetry.  out A[l][kz], kz <= Mz/2 (complex).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) k_rho_z(const double* __restrict__ rho, MeshDims md,
                                                      const double2* __restrict__ tw_g, double2* __restrict__ A) {
    const FftGeom g = fft_geom(md.mz);
    const int M = g.M, n = md.nz, H = M / 2;
    FftSmem s = fft_smem(g, tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * g.NL);
    for (int t = threadIdx.x; t < g.NL * M; t += kFftThreads) {
        const int p = t / M, k = t - p * M;               // k fastest: contiguous reads of rho
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        double a = 0.0, b = 0.0;
        if (k < n && l1 < nlines) {
            a = __ldg(rho + (size_t)l1 * n + k);
            if (l2 < nlines) b = __ldg(rho + (size_t)l2 * n + k);
        }
        s.a[k * g.NLP + p] = make_double2(a, b);
    }
    double2* Z = block_fft<false>(s.a, s.b, s.tw, g);
    for (int t = threadIdx.x; t < g.NL * (H + 1); t += kFftThreads) {
        const int p = t / (H + 1), kz = t - p * (H + 1);
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 z = Z[kz * g.NLP + p];
        const double2 w = Z[((M - kz) & (M - 1)) * g.NLP + p];   // Z[M-k], Z[M] == Z[0]
        // F1 = (Z[k] + conj(Z[M-k]))/2 ; F2 = (Z[k] - conj(Z[M-k]))/(2i)
        A[(size_t)l1 * (H + 1) + kz] = make_double2(0.5 * (z.x + w.x), 0.5 * (z.y - w.y));
        if (l2 < nlines) A[(size_t)l2 * (H + 1) + kz] = make_double2(0.5 * (z.y + w.y), 0.5 * (w.x - z.x));
    }
}

// ---------------------------------------------------------------------------
// complex transform along the OUTER axis of in[batch][n_in][inner] (zero-padded
// to M) -> out[batch][n_out][inner] (first n_out outputs kept).
//   MODE 0: forward            (rho pass y)
//   MODE 1: inverse            (inverse pass y)
//   MODE 2: forward, multiply by the real even K_hat, inverse  (pass x, in place)
// For MODE 2 inner = My*(Mz/2+1) and khat is [Mx/2+1][My/2+1][Mz/2+1].
// ---------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kFftThreads) k_cplx_outer(const double2* in, double2* out,   // may alias (MODE 2)
                                                           int n_in, int n_out, int M, int inner,
                                                           const double2* __restrict__ tw_g,
                                                           const double* __restrict__ khat, MeshDims md) {
    const FftGeom g = fft_geom(M);
    FftSmem s = fft_smem(g, tw_g);
    const int blocks_per_batch = (inner + g.NL - 1) / g.NL;
    const int batch = blockIdx.x / blocks_per_batch;
    const int f0 = (blockIdx.x - batch * blocks_per_batch) * g.NL;
    const int nl = min(g.NL, inner - f0);
    const double2* src = in + (size_t)batch * n_in * inner;
    double2* dst = out + (size_t)batch * n_out * inner;
    for (int t = threadIdx.x; t < g.NL * M; t += kFftThreads) {
        const int o = t / g.NL, l = t - o * g.NL;          // l fastest: adjacent inner indices
        double2 v = make_double2(0.0, 0.0);
        if (o < n_in && l < nl) v = src[(size_t)o * inner + f0 + l];
        s.a[o * g.NLP + l] = v;
    }
    double2* X = (MODE == 1) ? block_fft<true>(s.a, s.b, s.tw, g) : block_fft<false>(s.a, s.b, s.tw, g);
    if (MODE == 2) {
        const int hz1 = md.mz / 2 + 1, hy1 = md.my / 2 + 1;
        for (int t = threadIdx.x; t < g.NL * M; t += kFftThreads) {
            const int kx = t / g.NL, l = t - kx * g.NL;
            if (l >= nl) continue;
            const int f = f0 + l;
            const int ky = f / hz1, kz = f - ky * hz1;
            const int sx = min(kx, M - kx), sy = min(ky, md.my - ky);
            const double gk = __ldg(khat + ((size_t)sx * hy1 + sy) * hz1 + kz);
            double2 v = X[kx * g.NLP + l];
            X[kx * g.NLP + l] = make_double2(v.x * gk, v.y * gk);
        }
        double2* Y = (X == s.a) ? s.b : s.a;
        X = block_fft<true>(X, Y, s.tw, g);
    }
    for (int t = threadIdx.x; t < g.NL * n_out; t += kFftThreads) {
        const int o = t / g.NL, l = t - o * g.NL;
        if (l < nl) dst[(size_t)o * inner + f0 + l] = X[o * g.NLP + l];
    }
}

// ---------------------------------------------------------------------------
// inverse pass z: Hermitian lines D[l][kz<=Mz/2] -> real, two per complex FFT;
// phi[l][k<nz] = value / (Mx My Mz) / (4 pi eps0 hx hy hz)   (sc.py:164,167)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kFftThreads) k_inv_z(const double2* __restrict__ D, MeshDims md,
                                                      const double2* __restrict__ tw_g, const double* __restrict__ hsrc,
                                                      double four_pi_eps0, double* __restrict__ phi) {
    const FftGeom g = fft_geom(md.mz);
    const int M = g.M, n = md.nz, H = M / 2;
    FftSmem s = fft_smem(g, tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * g.NL);
    for (int t = threadIdx.x; t < g.NL * (H + 1); t += kFftThreads) {
        const int p = t / (H + 1), k = t - p * (H + 1);   // k fastest: contiguous reads of D
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        double2 d1 = make_double2(0.0, 0.0), d2 = make_double2(0.0, 0.0);
        if (l1 < nlines) d1 = D[(size_t)l1 * (H + 1) + k];
        if (l2 < nlines) d2 = D[(size_t)l2 * (H + 1) + k];
        s.a[k * g.NLP + p] = make_double2(d1.x - d2.y, d1.y + d2.x);                    // d1 + i d2
        if (k > 0 && k < H)                                                             // Hermitian extension
            s.a[(M - k) * g.NLP + p] = make_double2(d1.x + d2.y, -d1.y + d2.x);         // conj(d1) + i conj(d2)
    }
    double2* X = block_fft<true>(s.a, s.b, s.tw, g);
    const double inv_m3 = 1.0 / ((double)md.mx * (double)md.my * (double)md.mz);
    const double denom = four_pi_eps0 * hsrc[0] * hsrc[1] * hsrc[2];
    for (int t = threadIdx.x; t < g.NL * n; t += kFftThreads) {
        const int p = t / n, k = t - p * n;
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 v = X[k * g.NLP + p];
        phi[(size_t)l1 * n + k] = (v.x * inv_m3) / denom;
        if (l2 < nlines) phi[(size_t)l2 * n + k] = (v.y * inv_m3) / denom;
    }
}

static size_t fft_smem_bytes(int M) {
    const FftGeom g = fft_geom(M);
    return sizeof(double2) * (2 * fft_buf_elems(g) + (size_t)M);
}

template <typename K>
static void opt_in(K kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fft_smem_bytes(512));
}

void fft_init_kernels() {
    opt_in(k_khat_z);
    opt_in(k_real_even_outer);
    opt_in(k_rho_z);
    opt_in(k_cplx_outer<0>);
    opt_in(k_cplx_outer<1>);
    opt_in(k_cplx_outer<2>);
    opt_in(k_inv_z);
}

int fft_max_length() { return 512; }

// K_hat from the antiderivative table: three real-even passes
void launch_khat(const double* gtab, MeshDims md, FftWork w, cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1, hy1 = md.my / 2 + 1;
    {   // z: P[nx][ny][hz1]
        const int lb = 2 * fft_geom(md.mz).NL;
        const int blocks = (md.nx * md.ny + lb - 1) / lb;
        k_khat_z<<<blocks, kFftThreads, fft_smem_bytes(md.mz), st>>>(gtab, md, w.tw_z, w.P);
    }
    {   // y: per a, in [ny][hz1] -> Q[a][hy1][hz1]
        const int pb = fft_geom(md.my).NL;
        const int blocks_per_batch = ((hz1 + 1) / 2 + pb - 1) / pb;
        k_real_even_outer<<<blocks_per_batch * md.nx, kFftThreads, fft_smem_bytes(md.my), st>>>(w.P, w.Q, md.ny, md.my,
                                                                                              hz1, w.tw_y);
    }
    {   // x: in [nx][hy1*hz1] -> khat[hx1][hy1*hz1]
        const int inner = hy1 * hz1;
        const int pb = fft_geom(md.mx).NL;
        const int blocks = ((inner + 1) / 2 + pb - 1) / pb;
        k_real_even_outer<<<blocks, kFftThreads, fft_smem_bytes(md.mx), st>>>(w.Q, w.khat, md.nx, md.mx, inner, w.tw_x);
    }
}

// phi = (rho (*) K) / (4 pi eps0 hx hy hz) on [0,n)^3
void launch_convolve(const double* rho, MeshDims md, FftWork w, const double* h3, double four_pi_eps0, double* phi,
                     cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1;
    const int lbz = 2 * fft_geom(md.mz).NL;
    const int zblocks = (md.nx * md.ny + lbz - 1) / lbz;
    k_rho_z<<<zblocks, kFftThreads, fft_smem_bytes(md.mz), st>>>(rho, md, w.tw_z, w.A);
    {   // y forward: per i, [ny][hz1] -> [My][hz1]
        const int lb = fft_geom(md.my).NL;
        const int bpb = (hz1 + lb - 1) / lb;
        k_cplx_outer<0><<<bpb * md.nx, kFftThreads, fft_smem_bytes(md.my), st>>>(w.A, w.B, md.ny, md.my, md.my, hz1,
                                                                                w.tw_y, nullptr, md);
    }
    {   // x: forward, * K_hat, inverse, keep i < nx (in place)
        const int inner = md.my * hz1;
        const int lb = fft_geom(md.mx).NL;
        const int blocks = (inner + lb - 1) / lb;
        k_cplx_outer<2><<<blocks, kFftThreads, fft_smem_bytes(md.mx), st>>>(w.B, w.B, md.nx, md.nx, md.mx, inner,
                                                                           w.tw_x, w.khat, md);
    }
    {   // y inverse: per i, [My][hz1] -> [ny][hz1]
        const int lb = fft_geom(md.my).NL;
        const int bpb = (hz1 + lb - 1) / lb;
        k_cplx_outer<1><<<bpb * md.nx, kFftThreads, fft_smem_bytes(md.my), st>>>(w.B, w.A, md.my, md.ny, md.my, hz1,
                                                                                w.tw_y, nullptr, md);
    }
    k_inv_z<<<zblocks, kFftThreads, fft_smem_bytes(md.mz), st>>>(w.A, md, w.tw_z, h3, four_pi_eps0, phi);
}

}  // namespace ocl
