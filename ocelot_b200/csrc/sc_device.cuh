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
};

// bunch frame (sc.py:224-239); T columns are t1, t2, t3
struct Frame {
    double T[3][3];
    double pav, gamma0, beta0;
};

// mesh geometry (sc.py:173-186)
struct Mesh {
    double steps[3];
    double xoff[3];
    double sumq;
    int n[3];
};

struct Cart {
    double x, y, z, px, py, pz;
};

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
    return sqrt((gam * gam - 1.0) * rp.inv_gb2 - xs * xs - ys * ys);   // :69-70
}

__device__ __forceinline__ Cart mad_to_cart(const RefParams& rp, double x, double xs, double y, double ys,
                                            double tau, double delta) {
    double gam;
    double pzr = mad_pz_rel(rp, xs, ys, delta, gam);
    double kt = (rp.gb_ref / gam) * tau;                          // beta/ratio * tau
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
    double gam = sqrt(1.0 + s * rp.inv_m2);                       // :27
    double zp = c.z / c.pz;
    tau = -zp * (gam * rp.m_e_eV);                                // :47, :51
    x = c.x - zp * c.px;                                          // :48
    y = c.y - zp * c.py;                                          // :49
    delta = (gam / rp.gamref - 1.0) * rp.inv_betaref;             // :50
    xs = c.px * rp.inv_pref;                                      // :52
    ys = c.py * rp.inv_pref;                                      // :53
}

// ---- sc.py:224-239 -----------------------------------------------------------
// sums = {sum px, sum py, sum pz, particle count}
__device__ __forceinline__ void derive_frame(const double* sums, double m_e_eV, Frame& f) {
    double cnt = sums[3];
    double a = sums[0] / cnt, b = sums[1] / cnt, c = sums[2] / cnt;   // np.mean(xp[3:6], axis=1)
    double pav = sqrt(a * a + b * b + c * c);                          // np.linalg.norm(t3)
    a = a / pav; b = b / pav; c = c / pav;
    // t1 = cross(ey, t3) = (c, 0, -a), normalised
    double n1 = sqrt(c * c + 0.0 + a * a);
    double t1x = c / n1, t1y = 0.0 / n1, t1z = -a / n1;
    // t2 = cross(t3, t1)
    double t2x = b * t1z - c * t1y;
    double t2y = c * t1x - a * t1z;
    double t2z = a * t1y - b * t1x;
    f.T[0][0] = t1x; f.T[0][1] = t2x; f.T[0][2] = a;
    f.T[1][0] = t1y; f.T[1][1] = t2y; f.T[1][2] = b;
    f.T[2][0] = t1z; f.T[2][1] = t2z; f.T[2][2] = c;
    f.pav = pav;
    double g = pav / m_e_eV;
    f.gamma0 = sqrt(g * g + 1.0);                                      // :237
    f.beta0 = sqrt(1.0 - 1.0 / (f.gamma0 * f.gamma0));                 // :238-239
}

// ---- sc.py:173-186 -----------------------------------------------------------
// emax = {max x, max y, max z, -min x, -min y, -min z}; esum = {sum q*x, q*y, q*z, sum q}
// draws = {scale, shift} of random_mesh, or scale <= 0 for "off"
__device__ __forceinline__ void derive_mesh(const double* emax, const double* esum, int nx, int ny, int nz,
                                            double scale, double shift, Mesh& m) {
    m.n[0] = nx; m.n[1] = ny; m.n[2] = nz;
    m.sumq = esum[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double lo = -emax[3 + c];
        double extent = emax[c] - lo;                                  // :173
        if (scale > 0.0) extent = extent * scale;                      // :175
        double h = extent / (double)(m.n[c] - 3);                      // :179
        double xmin = lo / h;                                          // :181 (min commutes with /h)
        double xmid = (esum[c] / h) / esum[3];                         // :182
        double off = floor(xmin - xmid) + xmid;                        // :183
        if (scale > 0.0) off = off + shift;                            // :185
        m.steps[c] = h;
        m.xoff[c] = off;
    }
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
    g0 = a / m.steps[0] - m.xoff[0];
    g1 = b / m.steps[1] - m.xoff[1];
    g2 = c / m.steps[2] - m.xoff[2];
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
    if (!(c0 >= 0.0 && c0 <= (double)(nx - 1) && c1 >= 0.0 && c1 <= (double)(ny - 1) && c2 >= 0.0 &&
          c2 <= (double)(nz - 1)))
        return 0.0;
    double f0 = floor(c0), f1 = floor(c1), f2 = floor(c2);
    int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
    double t0 = c0 - f0, t1 = c1 - f1, t2 = c2 - f2;
    int j0 = min(i0 + 1, nx - 1);
    const size_t row = (size_t)i1 * nz + i2, plane = (size_t)ny * nz;
    const EQuad A = F[(size_t)i0 * plane + row];
    const EQuad B = F[(size_t)j0 * plane + row];
    const double a0 = 1.0 - t0, b0 = 1.0 - t1, d0 = 1.0 - t2;
    double acc = A.v00 * a0 * b0 * d0;
    acc = acc + A.v01 * a0 * b0 * t2;
    acc = acc + A.v10 * a0 * t1 * d0;
    acc = acc + A.v11 * a0 * t1 * t2;
    acc = acc + B.v00 * t0 * b0 * d0;
    acc = acc + B.v01 * t0 * b0 * t2;
    acc = acc + B.v10 * t0 * t1 * d0;
    acc = acc + B.v11 * t0 * t1 * t2;
    return acc;
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

}  // namespace ocl
