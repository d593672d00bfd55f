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
// All 1-D transforms are Stockham FFTs in shared memory (fp64, natural order in
// and out) with radix-16/8/4/2 butterflies held in registers; the transform
// length is a template parameter, so every stride is an immediate.
#include "sc_kernels.h"

namespace ocl {

// threads per block: 128, except M = 512: 256 threads, in-place stages (74 KB of shared memory), two
// resident blocks per SM.  Measured at 255^3 on B200: 3.06 -> 1.90 ms per solve against the two-buffer
// version with one resident block; in-place stages at M <= 256 change nothing (three blocks of the
// two-buffer version are already resident).

// Shared-memory layout: complex point o of line l lives at x[o * NLP + l] with
// NLP = NL + 1.  Lines run across lanes, so every butterfly stage reads and
// writes whole rows (NL consecutive double2) whatever its stride -- no bank
// conflicts -- and the odd pitch keeps the transposed accesses of the z passes
// (lanes along o) conflict-free per quarter-warp as well.
template <int M>
struct Geom {
#ifndef OCL_FFT512_NL
#define OCL_FFT512_NL 8
#define OCL_FFT512_T 256
#endif
    // M <= 128: the grids are small (63^3: 3969 z lines), so the lines per block decide how many blocks there
    // are to spread over 148 SMs.  Round 1 used 2048/M lines (125 blocks of 128 threads at M = 128: 6 % of the
    // warp slots busy, 8-16 us per pass); OCL_FFT_LINES_SMALL lines per block with one thread per 8 points
    // gives more blocks, each with fewer loads in front of its first butterfly.  Measured on B200, whole solve
    // at 63^3 (M = 128): 16 lines 50.8 us, 8 lines 42.2 us, 4 lines 54.6 us; no difference at 31^3 (M = 64).
#ifndef OCL_FFT_LINES_SMALL
#define OCL_FFT_LINES_SMALL 8
#endif
    static constexpr int NL_SMALL = (OCL_FFT_LINES_SMALL * M > 2048) ? (2048 / M) : OCL_FFT_LINES_SMALL;
    static constexpr int T_SMALL = (NL_SMALL * M / 8 > 128) ? 128 : ((NL_SMALL * M / 8 < 32) ? 32 : NL_SMALL * M / 8);
    static constexpr int T = (M >= 512) ? OCL_FFT512_T : (M >= 256 ? 128 : T_SMALL);    // threads per block
    static constexpr int NL = (M >= 512) ? OCL_FFT512_NL : (M >= 256) ? 8 : NL_SMALL;   // lines per block
    static constexpr int NLP = NL + 1;
    static constexpr int ELEMS = M * NLP;
    // In-place stages (one shared buffer instead of Stockham's two): every thread reads all the
    // butterflies it owns into registers, the block synchronises, and the results go back into the same
    // buffer.  Halves the shared memory per block (M = 512: 147 -> 74 KB, 1 -> 3 resident blocks per SM),
    // which is what lets the load, transform and store phases of different blocks overlap.
#ifndef OCL_FFT_INPLACE_MIN
#define OCL_FFT_INPLACE_MIN 512
#endif
    static constexpr bool INPLACE = M >= OCL_FFT_INPLACE_MIN;
    static constexpr size_t SMEM = sizeof(double2) * ((INPLACE ? 1 : 2) * (size_t)ELEMS);
    // direct-I/O transforms of at most two stages (every M <= 256) stage through ONE buffer: forward / inverse
    // outer passes and the real-even passes ask for this much, which lets a fourth block share an SM at M = 256
    static constexpr size_t SMEM_ONE = sizeof(double2) * (size_t)ELEMS;
#ifndef OCL_FFT512_MINB
#define OCL_FFT512_MINB 2
#endif
    static constexpr int MINB = (M >= 512 && INPLACE) ? OCL_FFT512_MINB : 1;             // resident blocks asked of ptxas
};

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
// multiply by the constant forward twiddle (cr, -si) (inverse: its conjugate)
template <bool INV>
__device__ __forceinline__ double2 mulw(double2 v, double cr, double si) {
    return INV ? make_double2(v.x * cr - v.y * si, v.y * cr + v.x * si)
               : make_double2(v.x * cr + v.y * si, v.y * cr - v.x * si);
}

template <bool INV>
__device__ __forceinline__ void dft4(double2& v0, double2& v1, double2& v2, double2& v3) {
    double2 a0 = cadd(v0, v2), a1 = csub(v0, v2), a2 = cadd(v1, v3), a3 = rot90<INV>(csub(v1, v3));
    v0 = cadd(a0, a2); v1 = cadd(a1, a3); v2 = csub(a0, a2); v3 = csub(a1, a3);
}

constexpr double kH = 0.70710678118654752440;    // cos(pi/4)
constexpr double kC1 = 0.92387953251128673848;   // cos(pi/8)
constexpr double kS1 = 0.38268343236508978178;   // sin(pi/8)

// natural-order DFTs of register arrays: v[m] <- sum_r v[r] W_R^{r m}
template <bool INV, int R>
__device__ __forceinline__ void dft(double2 (&v)[R]) {
    static_assert(R == 2 || R == 4 || R == 8 || R == 16, "radix");
    if constexpr (R == 2) {
        double2 a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        dft4<INV>(v[0], v[1], v[2], v[3]);
    } else if constexpr (R == 8) {
        dft4<INV>(v[0], v[2], v[4], v[6]);          // evens -> v[0,2,4,6] = E[0..3]
        dft4<INV>(v[1], v[3], v[5], v[7]);          // odds  -> v[1,3,5,7] = O[0..3]
        double2 o0 = v[1], o1 = mulw<INV>(v[3], kH, kH), o2 = rot90<INV>(v[5]), o3 = mulw<INV>(v[7], -kH, kH);
        double2 e0 = v[0], e1 = v[2], e2 = v[4], e3 = v[6];
        v[0] = cadd(e0, o0); v[4] = csub(e0, o0);
        v[1] = cadd(e1, o1); v[5] = csub(e1, o1);
        v[2] = cadd(e2, o2); v[6] = csub(e2, o2);
        v[3] = cadd(e3, o3); v[7] = csub(e3, o3);
    } else {   // R == 16: input r = 4 n1 + a, output m = b + 4 c
#pragma unroll
        for (int a = 0; a < 4; ++a) dft4<INV>(v[a], v[a + 4], v[a + 8], v[a + 12]);   // u_a[b] at v[a + 4 b]
        // t_a[b] = u_a[b] W16^{a b}
        v[1 + 4] = mulw<INV>(v[1 + 4], kC1, kS1);       // a=1,b=1: W^1
        v[1 + 8] = mulw<INV>(v[1 + 8], kH, kH);         // a=1,b=2: W^2
        v[1 + 12] = mulw<INV>(v[1 + 12], kS1, kC1);     // a=1,b=3: W^3
        v[2 + 4] = mulw<INV>(v[2 + 4], kH, kH);         // a=2,b=1: W^2
        v[2 + 8] = rot90<INV>(v[2 + 8]);                // a=2,b=2: W^4
        v[2 + 12] = mulw<INV>(v[2 + 12], -kH, kH);      // a=2,b=3: W^6
        v[3 + 4] = mulw<INV>(v[3 + 4], kS1, kC1);       // a=3,b=1: W^3
        v[3 + 8] = mulw<INV>(v[3 + 8], -kH, kH);        // a=3,b=2: W^6
        v[3 + 12] = mulw<INV>(v[3 + 12], -kC1, -kS1);   // a=3,b=3: W^9
        // X[b + 4 c] = sum_a t_a[b] W4^{a c}: dft4 over a for each b, result c lands at position a
#pragma unroll
        for (int b = 0; b < 4; ++b) dft4<INV>(v[4 * b], v[4 * b + 1], v[4 * b + 2], v[4 * b + 3]);
        // now v[4 b + c] = X[b + 4 c]: transpose the 4x4 index
        double2 t;
#define OCL_SWAP(i, j) t = v[i]; v[i] = v[j]; v[j] = t;
        OCL_SWAP(1, 4) OCL_SWAP(2, 8) OCL_SWAP(3, 12) OCL_SWAP(6, 9) OCL_SWAP(7, 13) OCL_SWAP(11, 14)
#undef OCL_SWAP
    }
}

template <int M>
struct FftSmem {
    double2* a;
    double2* b;
    const double2* tw;
};

template <int M>
__device__ __forceinline__ FftSmem<M> fft_smem(const double2* __restrict__ tw_g) {
    extern __shared__ __align__(16) unsigned char raw[];
    FftSmem<M> s;
    s.a = reinterpret_cast<double2*>(raw);
    s.b = s.a + Geom<M>::ELEMS;
    s.tw = tw_g;          // 16*M bytes, read through L1 with __ldg: keeps three blocks resident per SM
    return s;
}

__host__ __device__ constexpr int radix_for(int rem) { return (rem % 16 == 0) ? 16 : (rem % 8 == 0) ? 8 : (rem % 4 == 0) ? 4 : 2; }
// in-place stages keep ITEMS butterflies per thread in registers: radix 8 (512 = 8 * 8 * 8) bounds that
// at 2 x 8 complex values per thread for 256 threads and 8 lines
#ifndef OCL_FFT_INPLACE_RADIX
#define OCL_FFT_INPLACE_RADIX 8
#endif
__host__ __device__ constexpr int radix_inplace(int rem) {
    return (OCL_FFT_INPLACE_RADIX >= 16 && rem % 16 == 0) ? 16 : (rem % 8 == 0) ? 8 : (rem % 4 == 0) ? 4 : 2;
}

// One Stockham stage of radix R on all NL lines: work item t -> (butterfly j, line l), l fastest.
template <bool INV, int M, int Ns, int R>
__device__ __forceinline__ void fft_stage(const double2* __restrict__ x, double2* __restrict__ y,
                                          const double2* __restrict__ tw) {
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP;
    constexpr int nb = M / R;
    constexpr int step = M / (Ns * R);
    constexpr int total = nb * NL;
    for (int t = threadIdx.x; t < total; t += Geom<M>::T) {
        const int j = t / NL, l = t % NL;       // NL is a power of two: shift / mask
        const int k = j & (Ns - 1);
        double2 v[R];
        const double2* src = x + j * NLP + l;
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src[r * nb * NLP];
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double2 w = __ldg(tw + r * k * step);
                if (INV) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        dft<INV, R>(v);
        double2* dst = y + ((j - k) * R + k) * NLP + l;
#pragma unroll
        for (int r = 0; r < R; ++r) dst[r * Ns * NLP] = v[r];
    }
}

// The same stage with source == destination: all reads, a barrier, all writes.
template <bool INV, int M, int Ns, int R>
__device__ __forceinline__ void fft_stage_inplace(double2* __restrict__ x, const double2* __restrict__ tw) {
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP, T = Geom<M>::T;
    constexpr int nb = M / R;
    constexpr int step = M / (Ns * R);
    constexpr int total = nb * NL;
    constexpr int ITEMS = (total + T - 1) / T;
    double2 v[ITEMS][R];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int t = threadIdx.x + it * T;
        if (total % T == 0 || t < total) {
            const int j = t / NL, l = t % NL;
            const int k = j & (Ns - 1);
            const double2* src = x + j * NLP + l;
#pragma unroll
            for (int r = 0; r < R; ++r) v[it][r] = src[r * nb * NLP];
            if (Ns > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    double2 w = __ldg(tw + r * k * step);
                    if (INV) w.y = -w.y;
                    v[it][r] = cmul(v[it][r], w);
                }
            }
            dft<INV, R>(v[it]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int t = threadIdx.x + it * T;
        if (total % T == 0 || t < total) {
            const int j = t / NL, l = t % NL;
            const int k = j & (Ns - 1);
            double2* dst = x + ((j - k) * R + k) * NLP + l;
#pragma unroll
            for (int r = 0; r < R; ++r) dst[r * Ns * NLP] = v[it][r];
        }
    }
}

template <bool INV, int M, int Ns>
__device__ __forceinline__ void fft_stages(double2*& x, double2*& y, const double2* tw) {
    if constexpr (Ns < M) {
        if constexpr (Geom<M>::INPLACE) {
            constexpr int R = radix_inplace(M / Ns);
            fft_stage_inplace<INV, M, Ns, R>(x, tw);
            __syncthreads();
            fft_stages<INV, M, Ns * R>(x, y, tw);
        } else {
            constexpr int R = radix_for(M / Ns);
            fft_stage<INV, M, Ns, R>(x, y, tw);
            __syncthreads();
            double2* tmp = x; x = y; y = tmp;
            fft_stages<INV, M, Ns * R>(x, y, tw);
        }
    }
}

// All NL lines of length M in x.  Returns the buffer holding the result (natural order).
// tw[m] = exp(-2 pi i m / M); the inverse transform conjugates it (unnormalised).
template <bool INV, int M>
__device__ __forceinline__ double2* block_fft(double2* x, double2* y, const double2* tw) {
    __syncthreads();
    fft_stages<INV, M, 1>(x, y, tw);
    return x;
}

// ---------------------------------------------------------------------------
// Stages with pluggable source and sink.  The passes along an OUTER axis read and write global memory with the
// line index fastest, which is exactly the (butterfly j, line l) work distribution of a stage: the first stage can
// take its R inputs straight from global memory and the last stage can hand its R outputs straight to the
// consumer (global store, or the K_hat multiply of the x pass), so a two-stage transform touches shared memory
// once (one write, one read) instead of three times.  The FFT kernels are bound by the shared-memory / L1TEX data
// stage (profiles/r1_all_kernels_c4_summary.csv: k_cplx_outer<256,2> 67 % L1TEX, 27 % fp64), so this is where
// their time goes.  src(e, l) returns element e of line l; dst(e, l, v) consumes it.
// ---------------------------------------------------------------------------
template <int M>
struct SmemIO {
    double2* p;
    __device__ __forceinline__ double2 operator()(int e, int l) const { return p[e * Geom<M>::NLP + l]; }
    __device__ __forceinline__ void operator()(int e, int l, double2 v) const { p[e * Geom<M>::NLP + l] = v; }
};

template <bool INV, int M, int Ns, int R, typename Src, typename Dst>
__device__ __forceinline__ void fft_stage_io(const Src& src, const Dst& dst, const double2* __restrict__ tw) {
    constexpr int NL = Geom<M>::NL;
    constexpr int nb = M / R;
    constexpr int step = M / (Ns * R);
    constexpr int total = nb * NL;
    for (int t = threadIdx.x; t < total; t += Geom<M>::T) {
        const int j = t / NL, l = t % NL;
        const int k = j & (Ns - 1);
        double2 v[R];
#pragma unroll
        for (int r = 0; r < R; ++r) v[r] = src(j + r * nb, l);
        if (Ns > 1) {
#pragma unroll
            for (int r = 1; r < R; ++r) {
                double2 w = __ldg(tw + r * k * step);
                if (INV) w.y = -w.y;
                v[r] = cmul(v[r], w);
            }
        }
        dft<INV, R>(v);
        const int base = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) dst(base + r * Ns, l, v[r]);
    }
}

// whole transform: first stage from src, last stage into dst, shared memory (buffers a, b) in between.
// Leaves a barrier-free exit: the caller synchronises before reusing the buffers.
template <bool INV, int M, int Ns, typename Src, typename Dst>
__device__ __forceinline__ void fft_io(const Src& src, const Dst& dst, double2* a, double2* b, const double2* tw) {
    constexpr int R = radix_for(M / Ns);
    constexpr bool last = Ns * R == M;
    if constexpr (Ns == 1 && last) {
        fft_stage_io<INV, M, Ns, R>(src, dst, tw);
    } else if constexpr (Ns == 1) {
        fft_stage_io<INV, M, Ns, R>(src, SmemIO<M>{a}, tw);
        __syncthreads();
        fft_io<INV, M, Ns * R>(src, dst, a, b, tw);
    } else if constexpr (last) {
        fft_stage_io<INV, M, Ns, R>(SmemIO<M>{a}, dst, tw);
    } else {
        fft_stage_io<INV, M, Ns, R>(SmemIO<M>{a}, SmemIO<M>{b}, tw);
        __syncthreads();
        fft_io<INV, M, Ns * R>(src, dst, b, a, tw);
    }
}

// The in-place geometry (M = 512: one shared buffer, every thread holds all its butterflies in registers across the
// barrier) with the same pluggable source / sink.  A stage whose source AND sink are the shared buffer needs the barrier
// between its reads and its writes; a stage that reads global memory or writes to the consumer does not.
template <bool INV, int M, int Ns, int R, bool SYNC, typename Src, typename Dst>
__device__ __forceinline__ void fft_stage_inplace_io(const Src& src, const Dst& dst, const double2* __restrict__ tw) {
    constexpr int NL = Geom<M>::NL, T = Geom<M>::T;
    constexpr int nb = M / R;
    constexpr int step = M / (Ns * R);
    constexpr int total = nb * NL;
    constexpr int ITEMS = (total + T - 1) / T;
    double2 v[ITEMS][R];
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int t = threadIdx.x + it * T;
        if (total % T == 0 || t < total) {
            const int j = t / NL, l = t % NL;
            const int k = j & (Ns - 1);
#pragma unroll
            for (int r = 0; r < R; ++r) v[it][r] = src(j + r * nb, l);
            if (Ns > 1) {
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    double2 w = __ldg(tw + r * k * step);
                    if (INV) w.y = -w.y;
                    v[it][r] = cmul(v[it][r], w);
                }
            }
            dft<INV, R>(v[it]);
        }
    }
    if (SYNC) __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
        const int t = threadIdx.x + it * T;
        if (total % T == 0 || t < total) {
            const int j = t / NL, l = t % NL;
            const int k = j & (Ns - 1);
            const int base = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; ++r) dst(base + r * Ns, l, v[it][r]);
        }
    }
}

// SRC_BUF / DST_BUF: the caller's source / sink is the shared buffer `buf` itself
template <bool INV, int M, int Ns, bool SRC_BUF, bool DST_BUF, typename Src, typename Dst>
__device__ __forceinline__ void fft_io_inplace(const Src& src, const Dst& dst, double2* buf, const double2* tw) {
    constexpr int R = radix_inplace(M / Ns);
    constexpr bool first = Ns == 1, last = Ns * R == M;
    constexpr bool src_buf = first ? SRC_BUF : true, dst_buf = last ? DST_BUF : true;
    constexpr bool SYNC = src_buf && dst_buf;
    if constexpr (first && last) {
        fft_stage_inplace_io<INV, M, Ns, R, SYNC>(src, dst, tw);
    } else if constexpr (first) {
        fft_stage_inplace_io<INV, M, Ns, R, SYNC>(src, SmemIO<M>{buf}, tw);
        __syncthreads();
        fft_io_inplace<INV, M, Ns * R, SRC_BUF, DST_BUF>(src, dst, buf, tw);
    } else if constexpr (last) {
        fft_stage_inplace_io<INV, M, Ns, R, SYNC>(SmemIO<M>{buf}, dst, tw);
    } else {
        fft_stage_inplace_io<INV, M, Ns, R, SYNC>(SmemIO<M>{buf}, SmemIO<M>{buf}, tw);
        __syncthreads();
        fft_io_inplace<INV, M, Ns * R, SRC_BUF, DST_BUF>(src, dst, buf, tw);
    }
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
template <int M>
__global__ void __launch_bounds__(Geom<M>::T, Geom<M>::MINB) k_khat_z(const double* __restrict__ gtab, MeshDims md,
                                                       const double2* __restrict__ tw_g, double* __restrict__ P) {
    pdl_enter();
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP, H = M / 2;
    const int n = md.nz;
    FftSmem<M> s = fft_smem<M>(tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * NL);              // 2 real lines per complex line
    for (int t = threadIdx.x; t < M * NLP; t += Geom<M>::T) s.a[t] = make_double2(0.0, 0.0);
    __syncthreads();
    for (int t = threadIdx.x; t < NL * n; t += Geom<M>::T) {
        const int p = t / n, c = t - p * n;               // c fastest: contiguous table reads
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        double e1 = green_entry_dev(gtab, md.ny + 1, md.nz + 1, l1 / md.ny, l1 % md.ny, c);
        double e2 = (l2 < nlines) ? green_entry_dev(gtab, md.ny + 1, md.nz + 1, l2 / md.ny, l2 % md.ny, c) : 0.0;
        s.a[c * NLP + p] = make_double2(e1, e2);
        if (c) s.a[(M - c) * NLP + p] = make_double2(e1, e2);
    }
    double2* X = block_fft<false, M>(s.a, s.b, s.tw);
    pdl_trigger();
    for (int t = threadIdx.x; t < NL * (H + 1); t += Geom<M>::T) {
        const int p = t / (H + 1), kz = t - p * (H + 1);
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 v = X[kz * NLP + p];
        P[(size_t)l1 * (H + 1) + kz] = v.x;
        if (l2 < nlines) P[(size_t)l2 * (H + 1) + kz] = v.y;
    }
}

// ---------------------------------------------------------------------------
// K_hat, passes y and x: real-even transform along the OUTER axis of
// in[batch][n][inner] -> out[batch][M/2+1][inner]; two adjacent inner indices
// per complex FFT.
// ---------------------------------------------------------------------------
template <int M>
__global__ void __launch_bounds__(Geom<M>::T, Geom<M>::MINB) k_real_even_outer(const double* __restrict__ in,
                                                                double* __restrict__ out, int n, int inner,
                                                                const double2* __restrict__ tw_g) {
    pdl_enter();
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP, H = M / 2;
    FftSmem<M> s = fft_smem<M>(tw_g);
    const int pairs_total = (inner + 1) / 2;
    const int blocks_per_batch = (pairs_total + NL - 1) / NL;
    const int batch = blockIdx.x / blocks_per_batch;
    const int pair0 = (blockIdx.x - batch * blocks_per_batch) * NL;
    const int pairs = min(NL, pairs_total - pair0);
    const double* src = in + (size_t)batch * n * inner;
    double* dst = out + (size_t)batch * (H + 1) * inner;
    // direct I/O: the even extension is resolved in the source functor (element o and M - o are the same sample),
    // the first H + 1 outputs go straight to global memory
    auto gsrc = [&](int o, int p) -> double2 {
        const int oo = o < n ? o : ((M - o) < n ? M - o : -1);
        if (oo < 0 || p >= pairs) return make_double2(0.0, 0.0);
        const int f = 2 * (pair0 + p);
        const double e1 = __ldg(src + (size_t)oo * inner + f);
        const double e2 = (f + 1 < inner) ? __ldg(src + (size_t)oo * inner + f + 1) : 0.0;
        return make_double2(e1, e2);
    };
    auto gdst = [&](int ko, int p, double2 v) {
        if (ko > H || p >= pairs) return;
        const int f = 2 * (pair0 + p);
        dst[(size_t)ko * inner + f] = v.x;
        if (f + 1 < inner) dst[(size_t)ko * inner + f + 1] = v.y;
    };
    if constexpr (Geom<M>::INPLACE) fft_io_inplace<false, M, 1, false, false>(gsrc, gdst, s.a, s.tw);
    else fft_io<false, M, 1>(gsrc, gdst, s.a, s.b, s.tw);
    pdl_trigger();
}

// ---------------------------------------------------------------------------
// rho, pass z: real lines rho[l][k<nz] zero-padded to Mz, two per complex FFT,
// separated by Hermitian symmetry.  out A[l][kz], kz <= Mz/2 (complex).
// ---------------------------------------------------------------------------
// With pr.world > 0 the load is the fused all-reduce of the charge grid: every value is the sum, in
// rank order, of the W ranks' partial grids read through NVLink peer mappings (line_offset selects this
// rank's x-slab in slab mode).
template <int M>
__global__ void __launch_bounds__(Geom<M>::T, Geom<M>::MINB) k_rho_z(const double* __restrict__ rho, PeerRho pr, long long line_offset,
                                                      MeshDims md, const double2* __restrict__ tw_g,
                                                      double2* __restrict__ A) {
    pdl_enter();
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP, H = M / 2;
    const int n = md.nz;
    FftSmem<M> s = fft_smem<M>(tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * NL);
    {   // k fastest: contiguous reads of rho; U independent loads in flight per thread
        constexpr int PER = NL * M / Geom<M>::T, U = PER < 8 ? PER : 8;
        for (int c = 0; c < PER; c += U) {
            double a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = threadIdx.x + (c + u) * Geom<M>::T;
                const int p = t / M, k = t % M;
                const int l1 = line0 + 2 * p, l2 = l1 + 1;
                a[u] = 0.0; b[u] = 0.0;
                if (pr.world == 0) {
                    if (k < n && l1 < nlines) a[u] = __ldg(rho + (size_t)l1 * n + k);
                    if (k < n && l2 < nlines) b[u] = __ldg(rho + (size_t)l2 * n + k);
                } else if (k < n) {
                    for (int w = 0; w < pr.world; ++w) {
                        if (l1 < nlines) a[u] += __ldcg(pr.p[w] + (size_t)(line_offset + l1) * n + k);
                        if (l2 < nlines) b[u] += __ldcg(pr.p[w] + (size_t)(line_offset + l2) * n + k);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = threadIdx.x + (c + u) * Geom<M>::T;
                s.a[(t % M) * NLP + t / M] = make_double2(a[u], b[u]);
            }
        }
    }
    double2* Z = block_fft<false, M>(s.a, s.b, s.tw);
    pdl_trigger();
    for (int t = threadIdx.x; t < NL * (H + 1); t += Geom<M>::T) {
        const int p = t / (H + 1), kz = t - p * (H + 1);
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 z = Z[kz * NLP + p];
        const double2 w = Z[((M - kz) & (M - 1)) * NLP + p];   // Z[M-k], Z[M] == Z[0]
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
template <int M, int MODE>
__global__ void __launch_bounds__(Geom<M>::T, Geom<M>::MINB) k_cplx_outer(const double2* in, double2* out,   // may alias (MODE 2)
                                                           int n_in, int n_out, int inner,
                                                           const double2* __restrict__ tw_g,
                                                           const double* __restrict__ khat, MeshDims md, SlabMap sm) {
    pdl_enter();
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP;
    FftSmem<M> s = fft_smem<M>(tw_g);
    const int blocks_per_batch = (inner + NL - 1) / NL;
    const int batch = blockIdx.x / blocks_per_batch;
    const int f0 = (blockIdx.x - batch * blocks_per_batch) * NL;
    int nl = min(NL, inner - f0);
    if (sm.mode == 3) nl = max(0, min(nl, sm.f_total - sm.f_base - f0));   // padding lines of the last chunk
    const double2* src = in + (size_t)batch * n_in * inner + f0;
    double2* dst = out + (size_t)batch * n_out * inner + f0;
    // chunk-layout address of element (x plane = batch, line f)
    auto chunk = [&](int f) -> size_t { return ((size_t)(f / sm.fs) * sm.sx + batch) * sm.fs + f % sm.fs; };
    // direct I/O: inputs from global memory into the first stage, outputs of the last stage straight to global memory
    // (MODE 0/1) or through the K_hat multiply into the inverse transform (MODE 2)
    auto gsrc = [&](int o, int l) -> double2 {
        if (o >= n_in || l >= nl) return make_double2(0.0, 0.0);
        if (sm.mode == 2) return in[chunk(o * inner + f0 + l)];
        return src[(size_t)o * inner + l];
    };
    auto gdst = [&](int o, int l, double2 v) {
        if (o >= n_out || l >= nl) return;
        if (sm.mode == 1) {
            const int f = o * inner + f0 + l;
            if (sm.peer[0]) sm.peer[f / sm.fs][((size_t)sm.rank * sm.sx + batch) * sm.fs + f % sm.fs] = v;   // NVLink store
            else out[chunk(f)] = v;
        } else if (sm.mode == 3 && sm.peer[0]) {
            sm.peer[o / sm.sx][((size_t)sm.rank * sm.sx + o % sm.sx) * sm.fs + f0 + l] = v;                    // NVLink store
        } else {
            dst[(size_t)o * inner + l] = v;
        }
    };
    constexpr bool IP = Geom<M>::INPLACE;
    __syncthreads();
    if constexpr (MODE == 0) {
        if constexpr (IP) fft_io_inplace<false, M, 1, false, false>(gsrc, gdst, s.a, s.tw);
        else fft_io<false, M, 1>(gsrc, gdst, s.a, s.b, s.tw);
    } else if constexpr (MODE == 1) {
        if constexpr (IP) fft_io_inplace<true, M, 1, false, false>(gsrc, gdst, s.a, s.tw);
        else fft_io<true, M, 1>(gsrc, gdst, s.a, s.b, s.tw);
    } else {
        const int hz1 = md.mz / 2 + 1, hy1 = md.my / 2 + 1;
        const size_t kplane = (size_t)hy1 * hz1;
        // every work item of this thread belongs to the same line (T is a multiple of NL): its (ky, kz) column
        // of K_hat is located once
        const int lt = threadIdx.x % Geom<M>::NL;
        const double* kcol = nullptr;
        if (lt < nl) {
            const int f = f0 + lt + (sm.mode == 3 ? sm.f_base : 0);
            const int ky = f / hz1, kz = f - ky * hz1;
            kcol = khat + (size_t)min(ky, md.my - ky) * hz1 + kz;
        }
        // spectrum element kx of line l, times the real even K_hat, is the input of the inverse transform: it lands
        // in buffer b (two-buffer geometry) or back in the one buffer (in-place geometry)
        double2* spec = IP ? s.a : s.b;
        auto mult = [&](int kx, int l, double2 v) {
            const double g = kcol ? __ldg(kcol + (size_t)min(kx, M - kx) * kplane) : 0.0;
            spec[kx * Geom<M>::NLP + l] = make_double2(v.x * g, v.y * g);
        };
        if constexpr (IP) {
            fft_io_inplace<false, M, 1, false, true>(gsrc, mult, s.a, s.tw);
            __syncthreads();
            fft_io_inplace<true, M, 1, true, false>(SmemIO<M>{s.a}, gdst, s.a, s.tw);
        } else {
            fft_io<false, M, 1>(gsrc, mult, s.a, s.b, s.tw);
            __syncthreads();
            // inverse: source = buffer b; the intermediate stage uses a (free again after the barrier above)
            fft_io<true, M, 1>(SmemIO<M>{s.b}, gdst, s.a, s.b, s.tw);
        }
    }
    pdl_trigger();
}

// ---------------------------------------------------------------------------
// inverse pass z: Hermitian lines D[l][kz<=Mz/2] -> real, two per complex FFT;
// phi[l][k<nz] = value / (Mx My Mz) / (4 pi eps0 hx hy hz)   (sc.py:164,167)
// ---------------------------------------------------------------------------
template <int M>
// mc != 0: phi is a multicast address (every rank's potential grid mapped as one range); the stores go through the
// NVSwitch, which replicates them into all ranks' grids (multimem.st): the all-gather of phi is this kernel's epilogue.
__global__ void __launch_bounds__(Geom<M>::T, Geom<M>::MINB) k_inv_z(const double2* __restrict__ D, MeshDims md,
                                                      const double2* __restrict__ tw_g, const double* __restrict__ hsrc,
                                                      double four_pi_eps0, double* __restrict__ phi, int mc) {
    pdl_enter();
    constexpr int NL = Geom<M>::NL, NLP = Geom<M>::NLP, H = M / 2;
    const int n = md.nz;
    FftSmem<M> s = fft_smem<M>(tw_g);
    const int nlines = md.nx * md.ny;
    const int line0 = blockIdx.x * (2 * NL);
    {   // k fastest: contiguous reads of D; U independent load pairs in flight per thread
        constexpr int TOTAL = NL * (H + 1), U = 4;
        for (int t0 = threadIdx.x; t0 < TOTAL; t0 += U * Geom<M>::T) {
            double2 d1[U], d2[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u * Geom<M>::T;
                const int p = t / (H + 1), k = t - p * (H + 1);
                const int l1 = line0 + 2 * p, l2 = l1 + 1;
                d1[u] = (t < TOTAL && l1 < nlines) ? D[(size_t)l1 * (H + 1) + k] : make_double2(0.0, 0.0);
                d2[u] = (t < TOTAL && l2 < nlines) ? D[(size_t)l2 * (H + 1) + k] : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int t = t0 + u * Geom<M>::T;
                if (t >= TOTAL) continue;
                const int p = t / (H + 1), k = t - p * (H + 1);
                s.a[k * NLP + p] = make_double2(d1[u].x - d2[u].y, d1[u].y + d2[u].x);          // d1 + i d2
                if (k > 0 && k < H)                                                             // Hermitian extension
                    s.a[(M - k) * NLP + p] = make_double2(d1[u].x + d2[u].y, -d1[u].y + d2[u].x);   // conj(d1) + i conj(d2)
            }
        }
    }
    double2* X = block_fft<true, M>(s.a, s.b, s.tw);
    pdl_trigger();
    const double inv_m3 = 1.0 / ((double)md.mx * (double)md.my * (double)md.mz);
    const double denom = four_pi_eps0 * hsrc[0] * hsrc[1] * hsrc[2];
    for (int t = threadIdx.x; t < NL * n; t += Geom<M>::T) {
        const int p = t / n, k = t - p * n;
        const int l1 = line0 + 2 * p, l2 = l1 + 1;
        if (l1 >= nlines) continue;
        const double2 v = X[k * NLP + p];
        const double p1 = (v.x * inv_m3) / denom, p2 = (v.y * inv_m3) / denom;
        if (mc) {
            asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(phi + (size_t)l1 * n + k), "d"(p1) : "memory");
            if (l2 < nlines)
                asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(phi + (size_t)l2 * n + k), "d"(p2) : "memory");
        } else {
            phi[(size_t)l1 * n + k] = p1;
            if (l2 < nlines) phi[(size_t)l2 * n + k] = p2;
        }
    }
}

// ---------------------------------------------------------------------------
// host side: dispatch on the transform length
// ---------------------------------------------------------------------------
#define OCL_FFT_DISPATCH(M_, ...)                                  \
    switch (M_) {                                                  \
        case 8: { constexpr int MM = 8; __VA_ARGS__ } break;       \
        case 16: { constexpr int MM = 16; __VA_ARGS__ } break;     \
        case 32: { constexpr int MM = 32; __VA_ARGS__ } break;     \
        case 64: { constexpr int MM = 64; __VA_ARGS__ } break;     \
        case 128: { constexpr int MM = 128; __VA_ARGS__ } break;   \
        case 256: { constexpr int MM = 256; __VA_ARGS__ } break;   \
        case 512: { constexpr int MM = 512; __VA_ARGS__ } break;   \
        default: break;                                            \
    }

template <int M>
static void opt_in_all() {
    const int bytes = (int)Geom<M>::SMEM;
    cudaFuncSetAttribute(k_khat_z<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_real_even_outer<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_rho_z<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_cplx_outer<M, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_cplx_outer<M, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_cplx_outer<M, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    cudaFuncSetAttribute(k_inv_z<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

void fft_init_kernels() {
    opt_in_all<8>(); opt_in_all<16>(); opt_in_all<32>(); opt_in_all<64>();
    opt_in_all<128>(); opt_in_all<256>(); opt_in_all<512>();
}

int fft_max_length() { return 512; }

// K_hat from the antiderivative table: three real-even passes
void launch_khat(const double* gtab, MeshDims md, FftWork w, cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1, hy1 = md.my / 2 + 1;
    OCL_FFT_DISPATCH(md.mz,   // z: P[nx][ny][hz1]
        const int lb = 2 * Geom<MM>::NL;
        const int blocks = (md.nx * md.ny + lb - 1) / lb;
        launch_k(k_khat_z<MM>, dim3(blocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, gtab, md, w.tw_z, w.P);
    )
    OCL_FFT_DISPATCH(md.my,   // y: per a, in [ny][hz1] -> Q[a][hy1][hz1]
        const int pb = Geom<MM>::NL;
        const int blocks_per_batch = ((hz1 + 1) / 2 + pb - 1) / pb;
        launch_k(k_real_even_outer<MM>, dim3(blocks_per_batch * md.nx), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, w.P, w.Q, md.ny, hz1,
                                                                                           w.tw_y);
    )
    OCL_FFT_DISPATCH(md.mx,   // x: in [nx][hy1*hz1] -> khat[hx1][hy1*hz1]
        const int inner = hy1 * hz1;
        const int pb = Geom<MM>::NL;
        const int blocks = ((inner + 1) / 2 + pb - 1) / pb;
        launch_k(k_real_even_outer<MM>, dim3(blocks), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, w.Q, w.khat, md.nx, inner, w.tw_x);
    )
}

// phi = (rho (*) K) / (4 pi eps0 hx hy hz) on [0,n)^3, in two halves: the forward z and y
// passes of rho do not need K_hat, so the caller can overlap them with the K_hat chain.
void launch_convolve_pre(const double* rho, PeerRho pr, MeshDims md, FftWork w, cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1;
    OCL_FFT_DISPATCH(md.mz,
        const int lbz = 2 * Geom<MM>::NL;
        const int zblocks = (md.nx * md.ny + lbz - 1) / lbz;
        launch_k(k_rho_z<MM>, dim3(zblocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, rho, pr, 0, md, w.tw_z, w.A);
    )
    OCL_FFT_DISPATCH(md.my,   // y forward: per i, [ny][hz1] -> [My][hz1]
        const int lb = Geom<MM>::NL;
        const int bpb = (hz1 + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 0>, dim3(bpb * md.nx), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, w.A, w.B, md.ny, md.my, hz1, w.tw_y,
                                                                             nullptr, md, SlabMap{});
    )
}

void launch_convolve_post(MeshDims md, FftWork w, const double* h3, double four_pi_eps0, double* phi,
                          cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1;
    OCL_FFT_DISPATCH(md.mx,   // x: forward, * K_hat, inverse, keep i < nx (in place)
        const int inner = md.my * hz1;
        const int lb = Geom<MM>::NL;
        const int blocks = (inner + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 2>, dim3(blocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, w.B, w.B, md.nx, md.nx, inner, w.tw_x, w.khat,
                                                                        md, SlabMap{});
    )
    OCL_FFT_DISPATCH(md.my,   // y inverse: per i, [My][hz1] -> [ny][hz1]
        const int lb = Geom<MM>::NL;
        const int bpb = (hz1 + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 1>, dim3(bpb * md.nx), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, w.B, w.A, md.my, md.ny, hz1, w.tw_y,
                                                                             nullptr, md, SlabMap{});
    )
    OCL_FFT_DISPATCH(md.mz,
        const int lbz = 2 * Geom<MM>::NL;
        const int zblocks = (md.nx * md.ny + lbz - 1) / lbz;
        launch_k(k_inv_z<MM>, dim3(zblocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, w.A, md, w.tw_z, h3, four_pi_eps0, phi, 0);
    )
}

// ---------------------------------------------------------------------------
// slab-decomposed solve: this rank owns sx x-planes of rho / phi and one chunk of fs (ky,kz)
// lines of the x pass; the caller exchanges `xchg` between the calls (all-to-all).
// ---------------------------------------------------------------------------
void launch_slab_forward(const double* rho_slab, PeerRho pr, long long line_offset, MeshDims md, int sx, int fs,
                         FftWork w, double2* xchg, PeerXchg px, cudaStream_t st) {
    MeshDims ms = md;
    ms.nx = sx;
    const int hz1 = md.mz / 2 + 1;
    OCL_FFT_DISPATCH(md.mz,
        const int lbz = 2 * Geom<MM>::NL;
        const int zblocks = (sx * md.ny + lbz - 1) / lbz;
        launch_k(k_rho_z<MM>, dim3(zblocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, rho_slab, pr, line_offset, ms, w.tw_z, w.A);
    )
    SlabMap sm{1, fs, sx, 0, 0, {}, px.rank};
    for (int w_ = 0; w_ < 8; ++w_) sm.peer[w_] = px.world > 0 ? px.b[w_] : nullptr;     // y output = the peers' x-pass input
    OCL_FFT_DISPATCH(md.my,   // y forward, stored in chunk layout for the all-to-all
        const int lb = Geom<MM>::NL;
        const int bpb = (hz1 + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 0>, dim3(bpb * sx), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, w.A, xchg, md.ny, md.my, hz1, w.tw_y,
                                                                          nullptr, md, sm);
    )
}

void launch_slab_xpass(double2* xchg, MeshDims md, int sx, int fs, int f_base, FftWork w, PeerXchg px, cudaStream_t st) {
    const int hz1 = md.mz / 2 + 1;
    SlabMap sm{3, fs, sx, f_base, md.my * hz1, {}, px.rank};
    for (int w_ = 0; w_ < 8; ++w_) sm.peer[w_] = px.world > 0 ? px.a[w_] : nullptr;     // x output = the peers' inverse-y input
    OCL_FFT_DISPATCH(md.mx,   // x: forward, * K_hat, inverse on this rank's chunk of lines, [nx_pad][fs] in place
        const int lb = Geom<MM>::NL;
        const int blocks = (fs + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 2>, dim3(blocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, xchg, xchg, md.nx, md.nx, fs, w.tw_x, w.khat,
                                                                        md, sm);
    )
}

void launch_slab_inverse(const double2* xchg, MeshDims md, int sx, int fs, FftWork w, const double* h3,
                         double four_pi_eps0, double* phi_slab, int multicast, cudaStream_t st) {
    MeshDims ms = md;
    ms.nx = sx;
    const int hz1 = md.mz / 2 + 1;
    SlabMap sm{2, fs, sx, 0, 0, {}, 0};
    OCL_FFT_DISPATCH(md.my,   // y inverse, read from chunk layout
        const int lb = Geom<MM>::NL;
        const int bpb = (hz1 + lb - 1) / lb;
        launch_k(k_cplx_outer<MM, 1>, dim3(bpb * sx), dim3(Geom<MM>::T), Geom<MM>::SMEM_ONE, st, xchg, w.A, md.my, md.ny, hz1, w.tw_y,
                                                                          nullptr, md, sm);
    )
    OCL_FFT_DISPATCH(md.mz,
        const int lbz = 2 * Geom<MM>::NL;
        const int zblocks = (sx * md.ny + lbz - 1) / lbz;
        launch_k(k_inv_z<MM>, dim3(zblocks), dim3(Geom<MM>::T), Geom<MM>::SMEM, st, w.A, ms, w.tw_z, h3, four_pi_eps0, phi_slab,
                 multicast);
    )
}

}  // namespace ocl
