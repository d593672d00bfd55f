// sc_kernels.h -- host-callable launch wrappers for the space-charge kernels.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include "sc_device.cuh"

namespace ocl {

// device-resident reduction state shared by the particle sweeps
struct ReduceState {
    double* part;          // [max_blocks][16] per-block partials
    unsigned int* ticket;  // [4] last-block tickets
    double* sums;          // [4]  OCL_SC_BUF_MOMENTUM
    double* emax;          // [6]  OCL_SC_BUF_EXTENT_MAX
    double* esum;          // [4]  OCL_SC_BUF_EXTENT_SUM
    double* geom;          // [24] geometry tap
    Geo* geo;              // frame + mesh of the current kick, derived once on the device
    int max_blocks;
    int defer;             // 1: the sweeps only reduce; the caller all-reduces and calls launch_finish (NCCL fallback)
};

struct MeshDims {
    int nx, ny, nz;   // SpaceCharge.nmesh_xyz
    int mx, my, mz;   // padded FFT lengths
};

// Peer-memory mailbox for the two scalar exchanges of a particle-sharded kick: every rank owns one
// small symmetric buffer that all ranks have mapped (NVLink/NVSwitch peer access).
//   doubles [0,32)    momentum slots   (4 per source rank)
//   doubles [32,112)  extent slots     (10 per source rank)
//   doubles [128,136) momentum flags   (one 64-bit epoch per source rank)
//   doubles [136,144) extent flags
struct Mailbox {
    double* peer[8];                 // this mailbox as mapped on every rank (peer[rank] is the local one)
    unsigned long long* epoch;       // local device counters {momentum, extent}
    int rank, world;
};
//   doubles [144,152) "rho ready" flags (barrier before the fused rho reduction)
//   doubles [152,160) "rho slice reduced" flags (exit barrier of k_nvls_reduce)
constexpr int kMailboxDoubles = 256;
// the charge grids of all ranks, as mapped in this rank's address space (world == 0: local rho only)
struct PeerRho {
    const double* p[8];
    int world;
};
void launch_mailbox_exchange(Mailbox mb, int which, ReduceState rs, int* err_flag, cudaStream_t st);
void launch_nvls_reduce(double* mc, long long lo, long long hi, double* out, Mailbox mb, ReduceState rs, unsigned int* ticket,
                        int* err_flag, cudaStream_t st);

int particle_grid(long long n, int max_blocks);

void launch_set_params(KickParams v, KickParams* dst, cudaStream_t st);
void launch_combine_extents(const double* all, int world, ReduceState rs, cudaStream_t st);
const void* set_params_kernel();
void launch_momentum(const double* r, long long ld, long long n, KP kp, ReduceState rs, Mailbox mb, int* mb_err,
                     KickParams* publish, cudaStream_t st);
const void* momentum_kernel(int tma);
void launch_extent(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs, MeshDims md,
                   Mailbox mb, int* mb_err, cudaStream_t st);
void launch_finish(int which, KP kp, ReduceState rs, MeshDims md, cudaStream_t st);
void launch_deposit(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs,
                    MeshDims md, double* rho, cudaStream_t st);
// ordered momentum sum: numpy's pairwise tree over exactly rounded momenta (plan built by the host, sc_abi.cu)
void launch_momentum_exact(const double* r, long long ld, long long n, KP kp, const uint2* leaves, int nleaves,
                           const uint2* nodes, const int* level_start, int nlevels, double* V, double* sums,
                           cudaStream_t st);
// ordered (run-to-run bit-identical, np.bincount order) deposit; scratch from deposit_ordered_scratch_bytes
size_t deposit_ordered_scratch_bytes(long long n, MeshDims md);
int launch_deposit_ordered(const double* r, long long ld, const double* q, long long n, KP kp, ReduceState rs,
                           MeshDims md, double* rho, void* scratch, size_t scratch_bytes, cudaStream_t st);
void launch_green_table(ReduceState rs, MeshDims md, double* gtab, double* h3, cudaStream_t st);
void launch_green_mirror(const double* gtab, MeshDims md, double* kpad, cudaStream_t st);
void launch_green_compact(const double* gtab, MeshDims md, double* k1, cudaStream_t st);
void launch_pad_rho(const double* rho, MeshDims md, double* pad, cudaStream_t st);
void launch_multiply(cufftDoubleComplex* rho_hat, const cufftDoubleComplex* k_hat, MeshDims md, cudaStream_t st);
void launch_crop_phi(const double* conv, ReduceState rs, MeshDims md, double* phi, cudaStream_t st);
// layout 0: z-fastest quad table | 1: x-fastest quad table fetched by lane pairs
void field_init_kernels();
void launch_field(const double* phi, ReduceState rs, MeshDims md, EQuad* equad, int layout, cudaStream_t st);
void launch_gather_kick(double* r, long long ld, long long n, KP kp, ReduceState rs, MeshDims md,
                        const EQuad* equad, double* exyz_out, int do_kick, int layout, cudaStream_t st);
void launch_mad_to_cart(const double* r, long long ld, long long n, RefParams rp, double* xp, long long ld_xp,
                        cudaStream_t st);
void launch_cart_to_mad(const double* xp, long long ld_xp, long long n, RefParams rp, double* r, long long ld,
                        cudaStream_t st);
// ---- transfer maps and beam moments (sc_beam.cu) ----
struct MapCoef {
    double R[36];        // first-order matrix, row-major
    double B[6];         // constant term
    double tval[216];    // non-zero second-order coefficients ...
    unsigned char tidx[216];   // ... and their flat index a*36 + j*6 + k
    int nt;              // number of non-zero T terms (0: first-order map)
    // RF cavity body (CavityTM.map4cav, transformations/cavity.py:29-128), applied after R X + B:
    //   delta <- delta0*c1 + c2*(cos(phi - kb*tau0) - cos(phi));  tau += t566 d0^2 + t556 tau0 d0 + t555 tau0^2
    int cav;             // 0: no cavity step | 1: full | 2: drift-like (final energy non-physical, cavity.py:67-69)
    double c1, c2, kb, phi, cosphi, t566, t556, t555;
};
void launch_map_apply(double* r, long long ld, long long n, const MapCoef& mc, cudaStream_t st);
void launch_moments(const double* r, long long ld, long long n, const double* q, ReduceState rs, double* out,
                    cudaStream_t st);
// aperture cut + ordered stream compaction (sc_beam.cu); counts holds ceil(n/1024) + 1 ints
struct CutSpec {
    int kind;          // 0: one coordinate row against [a, b] | 1: ellipse with semi-axes (a, b) centred at (c, d)
    int row;
    double a, b, c, d;
};
void launch_cut(const double* r, long long ld, const double* q, const long long* ids, long long n, CutSpec c, int* counts,
                long long* n_out, double* r_out, long long ld_out, double* q_out, long long* ids_out, long long* lost_out,
                cudaStream_t st);

// ---- longitudinal space charge (sc_lsc.cu) ----
// physical constants of ocelot/common/globals.py:13-36 (same expressions as the host code)
constexpr double kPi = 3.141592653589793;
constexpr double kSpeedOfLight = 299792458.0;
constexpr double kEpsilon0 = 1 / (4 * kPi * 1e-7) / (kSpeedOfLight * kSpeedOfLight);
// per-kick scalars the host derives from the sweep-A statistics exactly like the reference
// (LSC.apply sc.py:575-592, s_to_cur analysis.py:293-333)
struct LscParams {
    double slice_min, slice_max;   // central slice, sc.py:579-582
    double x_shift, y_shift;       // shifts of the slice sums (bunch centroid)
    double a, ds;                  // grid x_j = j*ds + a
    double sigma_s;                // smoothing width sigma_tau * smooth_param
    double q, v, gamma, dz;        // sum(q_array), mean velocity, E/m_e, step length
    double und;                    // 1 + K_max^2 fill_factor / 2 (undulator factor)
    double pc_ref;                 // sqrt(E^2/m_e^2 - 1) m_e [GeV]
    int nb;                        // grid points (N + 1)
    int K;                         // smoothing taps -K..K; < 0: no smoothing
    int step_profile;              // 0: round Gaussian beam | 1: uniform beam of radius rb
    int fx_shift;                  // fixed-point scale of the CIC counts: one particle = 2^fx_shift
};
// packed deposit word: (count << cshift) | fraction sum with fbits fractional bits
struct LscPack {
    int replicas, cshift, fbits;
};
// Kernels receive the per-kick scalars either by value (host-derived grid, after the one host
// synchronisation of ocl_sc_lsc_stats) or through a device-resident copy that k_lsc_params derives from
// the sweep-A statistics (ocl_sc_lsc_kick_async: no host synchronisation).
struct LP {
    LscParams v;
    const LscParams* p;    // nullptr: use v
};
struct PKP {
    LscPack v;
    const LscPack* p;
};
// what the host still supplies in the asynchronous form (everything that does not depend on the particles)
struct LscHost {
    double gamma, v, pc_ref, dz, und;     // from p_array.E, dz and the undulator profile
    double bound_lo, bound_hi;            // LSC.bounds
    double smooth_param;
    int step_profile;
    int fx_shift;                         // 62 - ceil(log2 n_total), <= 52
    int cap_nb;                           // grid points the buffers hold
    long long warps, iters;               // deposit launch shape: warps in the grid, particles per thread
};
struct LscWork {
    double* part;                  // per-block partials (shared with the kick sweeps)
    unsigned int* ticket;          // [2]
    double* stats;                 // [16] sweep A results
    double* slice;                 // [9]  slice maxima (4) and sums (5)
    double* sigma;                 // [1]  transverse size used by the impedance
    unsigned long long* spread;    // [replica][nb] packed deposit counters
    unsigned long long* bins;      // [cap] the replicas folded: C[j] * 2^fx_shift
    double* A;                     // [cap] impedance Za = i A
    double* cnt;                   // [cap] counts as doubles
    double* prof;                  // [cap] bunch * c
    double* cur;                   // [cap] I(s) [A]
    double* W;                     // [cap] wake * q [eV... V]
    double2* Z;                    // [cap]
    double2* tw;                   // [2 cap] exp(2 pi i m / n)
    LscParams* dparams;            // device-derived scalars of the asynchronous form
    LscPack* dpack;
    int* err;                      // device flag: 1 = grid larger than the buffers, 2 = packed word too narrow
    int max_blocks;
};
void launch_lsc_stats(const double* r, long long ld, const double* q, long long n, LscWork w, cudaStream_t st);
void launch_lsc_twiddles(int nb, LscWork w, cudaStream_t st);
int launch_lsc_deposit(const double* r, long long ld, long long n, const LscParams& lp, LscWork w, cudaStream_t st);
int launch_lsc_solve(const LscParams& lp, LscWork w, cudaStream_t st);
long long lsc_spread_words(int nb);
void launch_lsc_kick(double* r, long long ld, long long n, const LscParams& lp, LscWork w, cudaStream_t st);
// asynchronous form: stats, device-side grid definition, deposit, solve, kick -- no host synchronisation
void launch_lsc_kick_async(double* r, long long ld, const double* q, long long n, LscHost hp, LscWork w,
                           cudaStream_t st);
constexpr int kLscAsyncCap = 8192;        // grid points the asynchronous form is sized for

// ---- hand-written Hockney convolution (sc_fft.cu) ----
struct FftWork {
    const double2* tw_x;   // exp(-2 pi i m / M) tables, one per axis
    const double2* tw_y;
    const double2* tw_z;
    double* P;             // [nx][ny][Mz/2+1]            K_hat after pass z
    double* Q;             // [nx][My/2+1][Mz/2+1]        K_hat after pass y
    double* khat;          // [Mx/2+1][My/2+1][Mz/2+1]    real, even
    double2* A;            // [nx][ny][Mz/2+1]            rho after pass z / before inverse z
    double2* B;            // [nx][My][Mz/2+1]            rho after pass y; pass x in place
};
// slab (multi-GPU) addressing of the y/x passes: the [i][f] array (f = ky*H + kz) is exchanged
// between ranks in chunks of fs lines; chunk layout index = ((f / fs) * sx + i) * fs + f % fs.
struct SlabMap {
    int mode;      // 0 none | 1 chunked output (y forward) | 2 chunked input (y inverse) | 3 x pass on one f-chunk
    int fs;        // lines per chunk
    int sx;        // x planes per rank
    int f_base;    // mode 3: global index of this rank's first line
    int f_total;   // mode 3: number of real lines (My * H); beyond it the chunk is padding
    // Fused transpose (peer[0] != nullptr): the pass stores its output straight into the exchange buffer of the rank
    // that consumes it, over NVLink peer mappings, instead of into a local buffer that an all-to-all then moves.
    //   mode 1: element (plane i of this rank, line f) -> peer[f / fs] [(rank*sx + i)*fs + f % fs]      (the x-pass input)
    //   mode 3: element (plane o, local line fl)       -> peer[o / sx] [(rank*sx + o % sx)*fs + fl]     (the inverse-y input)
    double2* peer[8];
    int rank;
};
// the exchange buffers of every rank as mapped here (world == 0: no peer mappings, NCCL all-to-all)
struct PeerXchg {
    double2* a[8];
    double2* b[8];
    int rank, world;
};
void fft_init_kernels();
int fft_max_length();
void launch_khat(const double* gtab, MeshDims md, FftWork w, cudaStream_t st);
void launch_convolve_pre(const double* rho, PeerRho pr, MeshDims md, FftWork w, cudaStream_t st);
void launch_convolve_post(MeshDims md, FftWork w, const double* h3, double four_pi_eps0, double* phi,
                          cudaStream_t st);
// slab-decomposed variants (this rank owns sx x-planes and one chunk of fs (ky,kz) lines)
void launch_slab_forward(const double* rho_slab, PeerRho pr, long long line_offset, MeshDims md, int sx, int fs,
                         FftWork w, double2* xchg, PeerXchg px, cudaStream_t st);
void launch_slab_xpass(double2* xchg, MeshDims md, int sx, int fs, int f_base, FftWork w, PeerXchg px, cudaStream_t st);
void launch_slab_inverse(const double2* xchg, MeshDims md, int sx, int fs, FftWork w, const double* h3,
                         double four_pi_eps0, double* phi_slab, int multicast, cudaStream_t st);
double four_pi_eps0_value();

// potential KAT helpers: steps given explicitly instead of derived from particles
void launch_green_table_steps(const double steps[3], MeshDims md, double* gtab, double* h3, cudaStream_t st);
void launch_crop_phi_steps(const double* conv, const double steps[3], MeshDims md, double* phi, cudaStream_t st);

}  // namespace ocl
