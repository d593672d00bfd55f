// sc_device.cuh -- per-particle and per-kick device math shared by all kernels.
//
// Everything here restates the arithmetic of the reference (ocelot/cpbd/sc.py,
// ocelot/cpbd/coord_transform.py).  Where rounding is amplified downstream --
// the mesh geometry (cell assignment) and the integrated Green's function
// (8-corner cancellation) -- the reference's operation order is kept exactly,
// with explicit round-to-nearest intrinsics so the compiler cannot contract
// them into FMAs.  The per-particle MAD<->Cartesian transforms are reduced
// algebraically (derivations at the functions); they agree with the reference
// to a few ulp, far inside the 1e-10 parity bound, at a fifth of the fp64 work.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocl {

// scalars derived on the host from p_array.E (sc.py:214-216, coord_transform.py:19-20,:61)
struct RefParams {
    double gamref;     // E / m_e_GeV
    double betaref;    // sqrt(1 - gamref^-2)
    double gb_ref;     // gamref * betaref
    double inv_gb2;    // 1 / gb_ref^2
    double pc;         // gb_ref * m_e_eV : momentum per unit slope (= pref up to rounding)
    double inv_pref;   // 1 / (m_e_eV * sqrt(gamref^2 - 1))      coord_transform.py:19
    double m_e_eV;
    double inv_m2;     // 1 / m_e_eV^2
    double inv_betaref;
    double inv_gamref;
};

// per-kick scalars.  Kernels receive them either by value (direct launches) or through a
// device-resident copy (CUDA-graph launches, where one parameter node is refreshed per kick).
struct Draws {
    double scale;  // <= 0: random_mesh off   (sc.py:175)
    double shift;  //                          (sc.py:185)
};
struct KickParams {
    RefParams rp;
    Draws dr;
    double cdT;    // dz / betaref             (sc.py:244)
};
struct KP {
    KickParams v;
    const KickParams* p;   // nullptr: use v
};
__device__ __forceinline__ RefParams kp_ref(const KP& k) { return k.p ? k.p->rp : k.v.rp; }
__device__ __forceinline__ Draws kp_draws(const KP& k) { return k.p ? k.p->dr : k.v.dr; }
__device__ __forceinline__ double kp_cdT(const KP& k) { return k.p ? k.p->cdT : k.v.cdT; }

// bunch frame (sc.py:224-239); T columns are t1, t2, t3
struct Frame {
    double T[3][3];
    double pav, gamma0, beta0;
};

// mesh geometry (sc.py:173-186)
struct Mesh {
    double steps[3];
    double inv_steps[3];   // 1/steps (IEEE), so the per-particle X/steps is one multiply
    double xoff[3];
    double sumq;
    int n[3];
};

struct Cart {
    double x, y, z, px, py, pz;
};

// Frame and mesh of the current kick, derived ONCE per kick on the device (by the block that finishes the
// momentum / extent reduction, after the cross-rank exchange in a sharded kick) and read by every
// later kernel: no kernel repeats the serial fp64 divisions / square roots of derive_frame / the mesh derivation
// in its prologue, and all of them see bit-identical geometry.
struct Geo {
    Frame f;
    Mesh m;
};

// Branch-free fp64 reciprocal and square root for normal, positive-range operands
// (every use below is on gamma, momenta or mesh steps): hardware seed (MUFU.RCP64H /
// MUFU.RSQ64H, ~2^-20) plus Newton / Goldschmidt steps in FMA arithmetic.  Results
// are within 1 ulp of the IEEE value; the IEEE library routines cost ~4x the
// instructions because of their special-case paths.
__device__ __forceinline__ double fast_rcp(double a) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(a));
    double e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    r = fma(r, e, r);
    e = fma(-a, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ double fast_sqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double g = a * y, h = 0.5 * y;
    double r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    r = fma(-h, g, 0.5);
    g = fma(g, r, g); h = fma(h, r, h);
    const double d = fma(-g, g, a);      // residual correction: result within 1 ulp
    return fma(d, h, g);
}

// ---- coord_transform.py:57-96, algebraically reduced ---------------------------
// The reference builds gamma, beta, pz/p_ref, normalises the direction u and
// multiplies back (:68-95).  In exact arithmetic |u_unnormalised| * pz/p_ref ==
// gamma*beta/(gamref*betaref) =: ratio, so
//     u = (x', y', pz_rel) / ratio,   u*beta = (x', y', pz_rel) * gb_ref/gamma,
//     p = u*gamma*beta*m_e = (x', y', pz_rel) * gb_ref*m_e,
//     pz_rel^2 = (gamma^2 - 1)/gb_ref^2 - x'^2 - y'^2.
// Same values to a few ulp, with 1 division + 1 square root instead of 10.
__device__ __forceinline__ double mad_pz_rel(const RefParams& rp, double xs, double ys, double delta, double& gam) {
    gam = (rp.betaref * delta + 1.0) * rp.gamref;                 // :68
    return fast_sqrt((gam * gam - 1.0) * rp.inv_gb2 - xs * xs - ys * ys);   // :69-70
}

__device__ __forceinline__ Cart mad_to_cart(const RefParams& rp, double x, double xs, double y, double ys,
                                            double tau, double delta) {
    double gam;
    double pzr = mad_pz_rel(rp, xs, ys, delta, gam);
    double kt = (rp.gb_ref * fast_rcp(gam)) * tau;                // beta/ratio * tau
    Cart c;
    c.x = x - xs * kt;                                            // :90
    c.y = y - ys * kt;                                            // :91
    c.z = -pzr * kt;                                              // :92
    c.px = xs * rp.pc;                                            // :93
    c.py = ys * rp.pc;                                            // :94
    c.pz = pzr * rp.pc;                                           // :95
    return c;
}

// ---- coord_transform.py:16-54, algebraically reduced ---------------------------
// With p0 = |p| = gamma*beta*m_e:  beta*u = p/(gamma*m_e), so
//     cdt = -z/(beta*u2) = -z*gamma*m_e/pz,   x + beta*u0*cdt = x - z*px/pz.
__device__ __forceinline__ void cart_to_mad(const RefParams& rp, const Cart& c, double& x, double& xs, double& y,
                                            double& ys, double& tau, double& delta) {
    double s = c.px * c.px + c.py * c.py + c.pz * c.pz;
    double gam = fast_sqrt(1.0 + s * rp.inv_m2);                  // :27
    double zp = c.z * fast_rcp(c.pz);
    tau = -zp * (gam * rp.m_e_eV);                                // :47, :51
    x = c.x - zp * c.px;                                          // :48
    y = c.y - zp * c.py;                                          // :49
    delta = (gam * rp.inv_gamref - 1.0) * rp.inv_betaref;         // :50
    xs = c.px * rp.inv_pref;                                      // :52
    ys = c.py * rp.inv_pref;                                      // :53
}

// 1 / g^2 rounded to nearest: numpy evaluates `gamma ** -2` (coord_transform.py:69) with libm / SVML pow,
// which is (nearly always) the correctly rounded value; a plain 1/(g*g) carries two roundings.
__device__ __forceinline__ double inv_square_rn(double g) {
    const double g2 = __dmul_rn(g, g);
    const double e = __fma_rn(g, g, -g2);                 // g*g = g2 + e exactly
    const double r0 = __ddiv_rn(1.0, g2);
    const double res = __fma_rn(-g2, r0, 1.0) - e * r0;   // 1 - (g2 + e) r0
    return __fma_rn(res, r0, r0);
}

// ---- sc.py:224-239 -----------------------------------------------------------
// sums = {sum px, sum py, sum pz, particle count}
__device__ __forceinline__ void derive_frame(const double* sums, double m_e_eV, Frame& f) {
    // Every operation is rounded on its own, as numpy does -- except the two norms, which numpy takes as
    // sqrt(dot(x, x)) with the BLAS dot kernel's k-sequential FMAs (checked against numpy bit for bit on 3000 random
    // frames, round 2): an FMA-contracted gamma0 or a separately rounded c*c would move pav / gamma0 by one ulp in
    // ~0.03 % / ~14 % of kicks, and one ulp of gamma0 is one ulp of h_z.
    const double cnt = sums[3];
    double a = __ddiv_rn(sums[0], cnt), b = __ddiv_rn(sums[1], cnt), c = __ddiv_rn(sums[2], cnt);   // np.mean(xp[3:6], axis=1)
    const double pav = __dsqrt_rn(__fma_rn(c, c, __fma_rn(b, b, __dmul_rn(a, a))));                // np.linalg.norm(t3)
    a = __ddiv_rn(a, pav); b = __ddiv_rn(b, pav); c = __ddiv_rn(c, pav);
    // t1 = cross(ey, t3) = (c, 0, -a), normalised
    const double n1 = __dsqrt_rn(__fma_rn(a, a, __dmul_rn(c, c)));
    const double t1x = __ddiv_rn(c, n1), t1y = __ddiv_rn(0.0, n1), t1z = __ddiv_rn(-a, n1);
    // t2 = cross(t3, t1): np.cross rounds both products, then subtracts
    const double t2x = __dsub_rn(__dmul_rn(b, t1z), __dmul_rn(c, t1y));
    const double t2y = __dsub_rn(__dmul_rn(c, t1x), __dmul_rn(a, t1z));
    const double t2z = __dsub_rn(__dmul_rn(a, t1y), __dmul_rn(b, t1x));
    f.T[0][0] = t1x; f.T[0][1] = t2x; f.T[0][2] = a;
    f.T[1][0] = t1y; f.T[1][1] = t2y; f.T[1][2] = b;
    f.T[2][0] = t1z; f.T[2][1] = t2z; f.T[2][2] = c;
    f.pav = pav;
    const double g = __ddiv_rn(pav, m_e_eV);
    f.gamma0 = __dsqrt_rn(__dadd_rn(__dmul_rn(g, g), 1.0));                                         // :237
    f.beta0 = __dsqrt_rn(__dsub_rn(1.0, inv_square_rn(f.gamma0)));                                  // :238-239
}

// sc.py:173-186 (mesh steps and origin from the reduced extents) lives in finish_extent_warp (sc_kernels.cu): one lane
// per axis, emax = {max x, max y, max z, -min x, -min y, -min z}, esum = {sum q*x, q*y, q*z, sum q}.

// block-wide copy of the kick geometry into shared memory (ends with a barrier)
__device__ __forceinline__ void load_geo(const Geo* __restrict__ g, Geo* s) {
    constexpr int W = (int)(sizeof(Geo) / sizeof(double));
    static_assert(sizeof(Geo) % sizeof(double) == 0, "Geo is copied as 8-byte words");
    if (threadIdx.x < W) reinterpret_cast<double*>(s)[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(g) + threadIdx.x);
    __syncthreads();
}

// ---- coord_transform.py:68-92 + sc.py:233,172 in the REFERENCE's operation order ------------------------
// Position of one particle in the bunch frame (z stretched by gamma0), every operation rounded separately
// as numpy does; the 3x3 rotation accumulates like the BLAS kernel behind np.dot (k-sequential FMAs).
// Used only for the <= 6 particles that define the mesh extents: the mesh step h = extent / (n - 3) feeds
// the integrated Green's function, whose 8-corner cancellation amplifies a one-ulp change of h to ~1e-10
// of the field, so the extremal coordinates are re-evaluated exactly like the reference does
// (DESIGN.md section 5).  All other per-particle work uses the reduced forms above.
struct ExactDir {
    double gam, bet, u0, u1, u2;
};
__device__ __forceinline__ ExactDir exact_direction(const RefParams& rp, double xs, double ys, double delta) {
    ExactDir d;
    d.gam = __dmul_rn(__dadd_rn(__dmul_rn(rp.betaref, delta), 1.0), rp.gamref);                       // :68
    d.bet = __dsqrt_rn(__dsub_rn(1.0, inv_square_rn(d.gam)));                                          // :69
    const double t = __ddiv_rn(__dmul_rn(d.gam, d.bet), rp.gb_ref);
    const double pz = __dsqrt_rn(__dsub_rn(__dsub_rn(__dmul_rn(t, t), __dmul_rn(xs, xs)), __dmul_rn(ys, ys)));   // :70
    const double d0 = __ddiv_rn(xs, pz), d1 = __ddiv_rn(ys, pz);                                       // :72
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(d0, d0), __dmul_rn(d1, d1)), 1.0));    // :74
    d.u0 = __ddiv_rn(d0, nrm); d.u1 = __ddiv_rn(d1, nrm); d.u2 = __ddiv_rn(1.0, nrm);                  // :77
    return d;
}
__device__ __forceinline__ void exact_frame_position(const RefParams& rp, const Frame& f, double x, double xs, double y,
                                                     double ys, double tau, double delta, double& a, double& b,
                                                     double& c) {
    const ExactDir d = exact_direction(rp, xs, ys, delta);
    const double X = __dsub_rn(x, __dmul_rn(__dmul_rn(d.u0, d.bet), tau));                             // :90
    const double Y = __dsub_rn(y, __dmul_rn(__dmul_rn(d.u1, d.bet), tau));                             // :91
    const double Z = __dmul_rn(__dmul_rn(-d.u2, d.bet), tau);                                          // :92
    a = __fma_rn(Z, f.T[2][0], __fma_rn(Y, f.T[1][0], __dmul_rn(X, f.T[0][0])));                       // sc.py:233
    b = __fma_rn(Z, f.T[2][1], __fma_rn(Y, f.T[1][1], __dmul_rn(X, f.T[0][1])));
    c = __fma_rn(Z, f.T[2][2], __fma_rn(Y, f.T[1][2], __dmul_rn(X, f.T[0][2])));
    c = __dmul_rn(c, f.gamma0);                                                                        // sc.py:172
}
// Cartesian momentum of one particle exactly as the reference rounds it (coord_transform.py:93-95:
// u * gamma * beta * m_e_eV, left to right); used by the ordered mode's pairwise momentum sum.
__device__ __forceinline__ void exact_momentum(const RefParams& rp, double xs, double ys, double delta, double& px,
                                               double& py, double& pz) {
    const ExactDir d = exact_direction(rp, xs, ys, delta);
    px = __dmul_rn(__dmul_rn(__dmul_rn(d.u0, d.gam), d.bet), rp.m_e_eV);
    py = __dmul_rn(__dmul_rn(__dmul_rn(d.u1, d.gam), d.bet), rp.m_e_eV);
    pz = __dmul_rn(__dmul_rn(__dmul_rn(d.u2, d.gam), d.bet), rp.m_e_eV);
}

// rotate into the bunch frame and stretch z (sc.py:233, :172)
__device__ __forceinline__ void rotate_stretch(const Frame& f, double x, double y, double z, double& a, double& b,
                                               double& c) {
    a = x * f.T[0][0] + y * f.T[1][0] + z * f.T[2][0];
    b = x * f.T[0][1] + y * f.T[1][1] + z * f.T[2][1];
    c = x * f.T[0][2] + y * f.T[1][2] + z * f.T[2][2];
    c = c * f.gamma0;
}

// position in cell units relative to the mesh origin (sc.py:180, :186)
__device__ __forceinline__ void to_grid(const Mesh& m, double a, double b, double c, double& g0, double& g1,
                                        double& g2) {
    g0 = __dsub_rn(__dmul_rn(a, m.inv_steps[0]), m.xoff[0]);   // two roundings, like X/steps - X_off
    g1 = __dsub_rn(__dmul_rn(b, m.inv_steps[1]), m.xoff[1]);
    g2 = __dsub_rn(__dmul_rn(c, m.inv_steps[2]), m.xoff[2]);
}

// Field table entry: the four (y,z) neighbours of one component at one cell,
// {E[i][j][k], E[i][j][k+1], E[i][j+1][k], E[i][j+1][k+1]} (indices clamped at
// the upper edge).  32 bytes = one L2 sector = one 256-bit load.
struct __align__(32) EQuad {
    double v00, v01, v10, v11;
};

// order-1 interpolation with zero outside [0, n-1] (scipy.ndimage.map_coordinates
// order=1, mode='constant', cval=0 as used at sc.py:202-204): each corner value
// is multiplied by its x, y, z weights in that order; corners accumulate with z
// fastest.  Two 256-bit loads fetch the eight corners.
__device__ __forceinline__ double trilinear(const EQuad* __restrict__ F, int nx, int ny, int nz, double c0,
                                            double c1, double c2) {
    // map_coordinates(mode='constant', cval=0): any coordinate outside [0, n-1] (or NaN) gives 0.
    // Branch-free on purpose: with an early return each of the three components of the gather sits in
    // its own reconvergence region, the compiler cannot hoist the next component's loads above the
    // previous component's arithmetic, and a particle pays three exposed L2 latencies in a row (60 % of
    // the kernel's stall samples, profiles/r1_gather_kick_c4_stalls.csv).  Out-of-range coordinates are
    // replaced by 0 so the loads stay inside the table, and the result is discarded by the select.
    const bool in = c0 >= 0.0 && c0 <= (double)(nx - 1) && c1 >= 0.0 && c1 <= (double)(ny - 1) && c2 >= 0.0 &&
                    c2 <= (double)(nz - 1);
    c0 = in ? c0 : 0.0; c1 = in ? c1 : 0.0; c2 = in ? c2 : 0.0;
    double f0 = floor(c0), f1 = floor(c1), f2 = floor(c2);
    int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
    double t0 = c0 - f0, t1 = c1 - f1, t2 = c2 - f2;
    int j0 = min(i0 + 1, nx - 1);
    const size_t row = (size_t)i1 * nz + i2, plane = (size_t)ny * nz;
    const EQuad A = F[(size_t)i0 * plane + row];
    const EQuad B = F[(size_t)j0 * plane + row];
    // three nested linear interpolations (z, then y, then x): 14 fp64 operations in a dependent chain of 6,
    // against 32 in a chain of 10 for the sum of eight weighted corners; the two agree to a few ulp
    const double a0 = 1.0 - t0, b0 = 1.0 - t1, d0 = 1.0 - t2;
    const double za0 = fma(A.v01, t2, A.v00 * d0), za1 = fma(A.v11, t2, A.v10 * d0);
    const double zb0 = fma(B.v01, t2, B.v00 * d0), zb1 = fma(B.v11, t2, B.v10 * d0);
    const double ya = fma(za1, t1, za0 * b0), yb = fma(zb1, t1, zb0 * b0);
    const double acc = fma(yb, t0, ya * a0);
    return in ? acc : 0.0;
}

// ---- x-fastest field table, fetched by lane pairs ----------------------------------------------------------
// The gather is bound by L1TEX wavefronts (one per distinct 128-byte line per warp instruction), not by bytes
// (profiles/r1_gather_kick_c4_summary.csv): with the z-fastest table above every particle touches six lines.
// Here the quad records of one component are stored x-fastest, rec(i, j, k) at F[(j*nz + k)*nx + i], so the two
// records a particle needs (x = i0 and i0 + 1) are adjacent, 64 contiguous bytes that share a 128-byte line three
// times out of four.  A lane pair fetches them with ONE request per particle: in the first load instruction the
// even lane reads record i0 and the odd lane record i0 + 1 of the EVEN lane's particle, in the second both read
// the ODD lane's particle; one exchange (8 x SHFL.32 per component) then hands every lane the two records of its
// own particle.  Wavefronts per particle: 3 x 1.25 instead of 6.  Must be called by all 32 lanes.
__device__ __forceinline__ double shfl_xor1(double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }

// cell, fractions and in-range flag of one interpolation point (map_coordinates, mode='constant', cval=0)
struct Interp {
    int rec;            // record index of the lower x plane: (i1*nz + i2)*nx + i0
    double t0, t1, t2;
    bool in;
};
__device__ __forceinline__ Interp interp_point(int nx, int ny, int nz, double c0, double c1, double c2) {
    Interp p;
    p.in = c0 >= 0.0 && c0 <= (double)(nx - 1) && c1 >= 0.0 && c1 <= (double)(ny - 1) && c2 >= 0.0 &&
           c2 <= (double)(nz - 1);
    c0 = p.in ? c0 : 0.0; c1 = p.in ? c1 : 0.0; c2 = p.in ? c2 : 0.0;     // out of range: a valid address, result discarded
    const double f0 = floor(c0), f1 = floor(c1), f2 = floor(c2);
    p.t0 = c0 - f0; p.t1 = c1 - f1; p.t2 = c2 - f2;
    // record i0 + 1 of the last plane (i0 == nx - 1, t0 == 0) is the next row's first record (or the zeroed
    // pad record behind the table): finite, and its weight is exactly zero
    p.rec = ((int)f1 * nz + (int)f2) * nx + (int)f0;
    return p;
}
// (y, z) part of the interpolation inside one quad record
__device__ __forceinline__ double bilinear(const EQuad& q, double t1, double t2) {
    const double d0 = 1.0 - t2;
    const double z0 = fma(q.v01, t2, q.v00 * d0), z1 = fma(q.v11, t2, q.v10 * d0);
    return fma(z1, t1, z0 * (1.0 - t1));
}

// All three field components of this lane's particle at grid position (g0, g1, g2) (sc.py:202-204; Ex, Ey NOT yet
// multiplied by gamma0).  The lane pair (2k, 2k+1) handles its two particles together: for each of them the even
// lane fetches the x = i0 records and the odd lane the x = i0 + 1 records (one request per particle and component
// when the two share a 128-byte line), each lane reduces its record with that particle's (y, z) fractions, and
// only the three reduced values per particle cross lanes.  Per particle: the partner's position (3 doubles) and
// the three partial results (3 doubles) are exchanged = 12 SHFL.32, against 27 for exchanging whole records.
__device__ __forceinline__ void gather_pair(const EQuad* __restrict__ ex, const EQuad* __restrict__ ey,
                                            const EQuad* __restrict__ ez, int nx, int ny, int nz, double g0, double g1,
                                            double g2, double& e0, double& e1, double& e2) {
    const int odd = (int)(threadIdx.x & 1);
    const double h0 = shfl_xor1(g0), h1 = shfl_xor1(g1), h2 = shfl_xor1(g2);       // the partner's particle
    // P: the even lane's particle, Q: the odd lane's particle (one of them is this lane's own)
    const double p0 = odd ? h0 : g0, p1 = odd ? h1 : g1, p2 = odd ? h2 : g2;
    const double q0 = odd ? g0 : h0, q1 = odd ? g1 : h1, q2 = odd ? g2 : h2;
    const Interp px = interp_point(nx, ny, nz, p0, p1 + 0.5, p2 + 0.5), qx = interp_point(nx, ny, nz, q0, q1 + 0.5, q2 + 0.5);
    const Interp py = interp_point(nx, ny, nz, p0 + 0.5, p1, p2 + 0.5), qy = interp_point(nx, ny, nz, q0 + 0.5, q1, q2 + 0.5);
    const Interp pz = interp_point(nx, ny, nz, p0 + 0.5, p1 + 0.5, p2), qz = interp_point(nx, ny, nz, q0 + 0.5, q1 + 0.5, q2);
    // six independent 256-bit loads; in each, the two lanes of a pair read adjacent records
    const EQuad rpx = ex[px.rec + odd], rqx = ex[qx.rec + odd];
    const EQuad rpy = ey[py.rec + odd], rqy = ey[qy.rec + odd];
    const EQuad rpz = ez[pz.rec + odd], rqz = ez[qz.rec + odd];
    const double bpx = bilinear(rpx, px.t1, px.t2), bqx = bilinear(rqx, qx.t1, qx.t2);
    const double bpy = bilinear(rpy, py.t1, py.t2), bqy = bilinear(rqy, qy.t1, qy.t2);
    const double bpz = bilinear(rpz, pz.t1, pz.t2), bqz = bilinear(rqz, qz.t1, qz.t2);
    // the even lane keeps P's lower-plane parts and ships Q's; the odd lane keeps Q's upper-plane parts and ships P's
    const double rx = shfl_xor1(odd ? bpx : bqx), ry = shfl_xor1(odd ? bpy : bqy), rz = shfl_xor1(odd ? bpz : bqz);
    const double tx = odd ? qx.t0 : px.t0, ty = odd ? qy.t0 : py.t0, tz = odd ? qz.t0 : pz.t0;
    const bool inx = odd ? qx.in : px.in, iny = odd ? qy.in : py.in, inz = odd ? qz.in : pz.in;
    const double lox = odd ? rx : bpx, hix = odd ? bqx : rx;           // lower / upper x plane of the own particle
    const double loy = odd ? ry : bpy, hiy = odd ? bqy : ry;
    const double loz = odd ? rz : bpz, hiz = odd ? bqz : rz;
    const double vx = fma(hix, tx, lox * (1.0 - tx));
    const double vy = fma(hiy, ty, loy * (1.0 - ty));
    const double vz = fma(hiz, tz, loz * (1.0 - tz));
    e0 = inx ? vx : 0.0;
    e1 = iny ? vy : 0.0;
    e2 = inz ? vz : 0.0;
}

// ---- asynchronous row pipeline -------------------------------------------------
// Every particle sweep streams NR rows (6 coordinate rows and/or q) exactly once.
// Each thread prefetches its own next D-1 trips into shared memory with
// cp.async (LDGSTS) while it does the fp64 math of the current one: the bytes in
// flight live in shared memory instead of registers, so the kernels keep 4
// blocks/SM resident and neither the HBM latency nor the long fp64 div/sqrt
// dependency chains are exposed.  A thread only ever reads the slots it filled
// itself, so no block barrier is needed.
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---- programmatic dependent launch (PDL) ----
// Every kernel of the kick starts with pdl_enter(): wait until the preceding kernel of the stream has
// completed and its writes are visible (griddepcontrol.wait).  Launched with the programmatic-
// serialization attribute, a kernel is set up while its predecessor is still draining, so the launch
// latency of a 15-kernel chain of 5-50 us kernels no longer adds up.  pdl_trigger() (griddepcontrol.
// launch_dependents) lets the next kernel's blocks become resident before this one exits; it is placed
// AFTER the main loop of a kernel (before its reduction / store tail), never at the top: measured on
// B200, triggering at the top makes the dependents' waiting blocks hold registers and shared memory
// for the whole kernel and costs +14 % at 1 M / 63^3 and +60 % at 12.5 M / 127^3.
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef OCL_PDL_TRIGGER
#define OCL_PDL_TRIGGER 1
#endif
__device__ __forceinline__ void pdl_trigger() {
#if OCL_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;");
#endif
}

#ifdef __CUDACC__
bool pdl_enabled();      // OCL_SC_PDL=0 turns the launch attribute off (sc_kernels.cu)
template <typename... Exp, typename... Act>
inline void launch_k(void (*kernel)(Exp...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Act&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<Act&&>(args)...);
}
#endif

constexpr int kSweepThreads = 256;

// body(i, v) is called once per particle i with v[k] = base[k][i].  n < 2^31.
// UNIFORM: the trip count is uniform across each warp (lanes beyond n run the body with v = 0 and
// valid = false), so the body may use warp shuffles: body(i, v, valid).
template <int NR, int D, bool UNIFORM = false, typename F>
__device__ __forceinline__ void pipelined_sweep(const double* const (&base)[NR], int n, double* sm, F&& body) {
    const int stride = (int)gridDim.x * kSweepThreads;
    int i = (int)blockIdx.x * kSweepThreads + (int)threadIdx.x;
    int ip = i;
    const int lane = UNIFORM ? (int)(threadIdx.x & 31) : 0;      // i - lane: index of the warp's first particle
    double* slot = sm + threadIdx.x;
#pragma unroll
    for (int d = 0; d < D - 1; ++d) {
        if (ip < n) {
#pragma unroll
            for (int k = 0; k < NR; ++k) cp_async8(slot + (d * NR + k) * kSweepThreads, base[k] + ip);
        }
        cp_async_commit();
        ip += stride;
    }
    while (i - lane < n) {
#pragma unroll
        for (int st = 0; st < D; ++st) {           // static stage index: all shared-memory offsets are immediates
            if (i - lane < n) {
                const int sp = (st + D - 1) % D;
                if (ip < n) {
#pragma unroll
                    for (int k = 0; k < NR; ++k) cp_async8(slot + (sp * NR + k) * kSweepThreads, base[k] + ip);
                }
                cp_async_commit();
                ip += stride;
                cp_async_wait<D - 1>();
                double v[NR];
                if constexpr (UNIFORM) {
                    const bool valid = i < n;
#pragma unroll
                    for (int k = 0; k < NR; ++k) v[k] = valid ? slot[(st * NR + k) * kSweepThreads] : 0.0;
                    body(i, v, valid);
                } else {
#pragma unroll
                    for (int k = 0; k < NR; ++k) v[k] = slot[(st * NR + k) * kSweepThreads];
                    body(i, v);
                }
                i += stride;
            }
        }
    }
    cp_async_wait<0>();
}

// ---- TMA row pipeline ------------------------------------------------------------------------------------------
// The same sweep with the rows moved by the bulk-copy engine (cp.async.bulk = 1-D TMA, UBLKCP in the SASS) instead of
// one 8-byte LDGSTS per thread and row: one thread per block posts NR copies of 2 KB (256 particles of one row) per
// stage against an mbarrier, the block consumes the stage when the barrier's transaction count completes.  Tiles are
// block-strided; D stages are in flight per block.  Requirements (checked by the launchers): every row base and the
// shared buffer are 16-byte aligned.  The last, partial tile (n % 256 particles) is read with plain loads -- a bulk
// copy may not run past the end of the caller's arrays.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// sm: [D][NR][kSweepThreads] doubles, 16-byte aligned; bars: [D].  body(i, v) once per particle.
template <int NR, int D, typename F>
__device__ __forceinline__ void bulk_sweep(const double* const (&base)[NR], int n, double* sm, unsigned long long* bars,
                                           F&& body) {
    constexpr int T = kSweepThreads;
    constexpr unsigned kRowBytes = T * sizeof(double);
    const int tiles = (n + T - 1) / T, full = n / T;           // tiles [0, full) are complete
    const int grid = (int)gridDim.x;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int d = 0; d < D; ++d) mbar_init(bars + d, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto post = [&](int it) {                                   // tile of iteration `it` into stage it % D
        const int ti = (int)blockIdx.x + it * grid;
        if (threadIdx.x == 0 && ti < full) {
            unsigned long long* bar = bars + it % D;
            mbar_expect_tx(bar, NR * kRowBytes);
            double* dst = sm + (size_t)(it % D) * NR * T;
#pragma unroll
            for (int k = 0; k < NR; ++k) bulk_load(dst + k * T, base[k] + (size_t)ti * T, kRowBytes, bar);
        }
    };
#pragma unroll
    for (int d = 0; d < D; ++d) post(d);
    for (int it = 0;; ++it) {
        const int ti = (int)blockIdx.x + it * grid;
        if (ti >= tiles) break;
        const int i = ti * T + (int)threadIdx.x;
        double v[NR];
        bool valid = true;
        if (ti < full) {
            mbar_wait(bars + it % D, (unsigned)((it / D) & 1));
            const double* src = sm + (size_t)(it % D) * NR * T + threadIdx.x;
#pragma unroll
            for (int k = 0; k < NR; ++k) v[k] = src[k * T];
        } else {                                                // the ragged tail, at most one tile per sweep
            valid = i < n;
#pragma unroll
            for (int k = 0; k < NR; ++k) v[k] = valid ? __ldcs(base[k] + i) : 0.0;
        }
        // The refill is written by the async proxy (TMA engine), which is not ordered behind this thread's LDS queue:
        // order the generic-proxy reads of the stage before it, then meet the other threads.  (Without the proxy
        // fence ~1 in 10^4 particles of k_deposit saw the NEXT tile's values: measured on B200.)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();                                        // every thread has taken its values out of this stage
        post(it + D);
        if (valid) body(i, v);
    }
}

// The same sweep without shared-memory staging: the next trip's rows are loaded straight into registers
// (plain LDG) before the current particle's arithmetic, so they are in flight during it.  Costs NR more
// registers per thread than the cp.async pipeline and saves its LDGSTS + LDS traffic through the L1TEX data
// stage -- the unit that bounds the gather kernel.  Warp-uniform trip count; body(i, v, valid).
template <int NR, typename F>
__device__ __forceinline__ void prefetch_sweep(const double* const (&base)[NR], int n, F&& body) {
    const int stride = (int)gridDim.x * kSweepThreads;
    const int lane = (int)(threadIdx.x & 31);
    int i = (int)blockIdx.x * kSweepThreads + (int)threadIdx.x;
    double cur[NR], nxt[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) cur[k] = i < n ? __ldcs(base[k] + i) : 0.0;
    while (i - lane < n) {
        const int in = i + stride;
#pragma unroll
        for (int k = 0; k < NR; ++k) nxt[k] = in < n ? __ldcs(base[k] + in) : 0.0;
        body(i, cur, i < n);
#pragma unroll
        for (int k = 0; k < NR; ++k) cur[k] = nxt[k];
        i = in;
    }
}

// ---- block reductions (fixed order => deterministic for a fixed launch shape) ----
template <typename Op>
__device__ __forceinline__ double warp_reduce(double v, Op op) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
struct OpSum { __device__ __forceinline__ double operator()(double a, double b) const { return a + b; } };
struct OpMax { __device__ __forceinline__ double operator()(double a, double b) const { return fmax(a, b); } };

constexpr int kSweepWarps = kSweepThreads / 32;

// ---------------------------------------------------------------------------
// reductions: first NMAX slots reduce with max, the rest with +
// ---------------------------------------------------------------------------
template <int NV, int NMAX>
__device__ __forceinline__ void block_reduce(double (&v)[NV], double* sh /* [NV*kSweepWarps] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        v[k] = (k < NMAX) ? warp_reduce(v[k], OpMax()) : warp_reduce(v[k], OpSum());
        if (lane == 0) sh[k * kSweepWarps + warp] = v[k];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double ident = (k < NMAX) ? -INFINITY : 0.0;
            double x = (lane < kSweepWarps) ? sh[k * kSweepWarps + lane] : ident;
            v[k] = (k < NMAX) ? warp_reduce(x, OpMax()) : warp_reduce(x, OpSum());
        }
    }
    __syncthreads();
}

// every block stores its partial; the block that draws the last ticket folds
// all partials in a fixed order and writes the result (deterministic for a
// fixed grid).  Returns true in the finishing block (v valid in thread 0).
template <int NV, int NMAX>
__device__ __forceinline__ bool grid_reduce(double (&v)[NV], double* part, unsigned int* ticket, double* sh) {
    __shared__ bool last;
    block_reduce<NV, NMAX>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) part[(size_t)blockIdx.x * NV + k] = v[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = (k < NMAX) ? -INFINITY : 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kSweepThreads) {
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            double x = __ldcg(part + (size_t)b * NV + k);
            v[k] = (k < NMAX) ? fmax(v[k], x) : v[k] + x;
        }
    }
    block_reduce<NV, NMAX>(v, sh);
    if (threadIdx.x == 0) *ticket = 0;
    return true;
}

// The extent reduction with the owner of every extremum: v[0..5] reduce with max and carry the index of the
// particle that attains them (ix), v[6..9] reduce with +.  Same fixed-order two-level scheme as grid_reduce;
// per-block partials are 16 doubles {10 values, 6 indices}.  Returns true in the finishing block, where
// thread 0 holds the results.
constexpr int kExtentPartial = 16;
__device__ __forceinline__ void block_reduce_arg(double (&v)[10], int (&ix)[6], double* sh /* [10*kSweepWarps] */,
                                                 double* shv /* [6] */, int* shi /* [6] */) {
    double mine[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) mine[k] = v[k];
    block_reduce<10, 6>(v, sh);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < 6; ++k) { shv[k] = v[k]; shi[k] = -1; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k)
        if (mine[k] == shv[k] && ix[k] >= 0) shi[k] = ix[k];          // ties: any owner will do
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 6; ++k) ix[k] = shi[k];
    __syncthreads();
}

__device__ __forceinline__ bool grid_reduce_extent(double (&v)[10], int (&ix)[6], double* part, unsigned int* ticket,
                                                   double* sh, double* shv, int* shi) {
    __shared__ bool last;
    block_reduce_arg(v, ix, sh, shv, shi);
    if (threadIdx.x == 0) {
        double* dst = part + (size_t)blockIdx.x * kExtentPartial;
#pragma unroll
        for (int k = 0; k < 10; ++k) dst[k] = v[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) dst[10 + k] = (double)ix[k];
        __threadfence();
        unsigned int t = atomicAdd(ticket, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
#pragma unroll
    for (int k = 0; k < 10; ++k) v[k] = (k < 6) ? -INFINITY : 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k) ix[k] = -1;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kSweepThreads) {
        const double* src = part + (size_t)b * kExtentPartial;
#pragma unroll
        for (int k = 0; k < 10; ++k) {
            const double x = __ldcg(src + k);
            if (k < 6) {
                if (x > v[k]) { v[k] = x; ix[k] = (int)__ldcg(src + 10 + k); }
            } else {
                v[k] += x;
            }
        }
    }
    block_reduce_arg(v, ix, sh, shv, shi);
    if (threadIdx.x == 0) *ticket = 0;
    return true;
}

}  // namespace ocl
