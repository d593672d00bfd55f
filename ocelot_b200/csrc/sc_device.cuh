// sc_device.cuh -- per-particle and per-kick device math shared by all kernels.
//
// Everything here restates, operation by operation, the arithmetic of the
// reference (ocelot/cpbd/sc.py, ocelot/cpbd/coord_transform.py).  The library
// is compiled with -fmad=false so that +,-,*,/ and sqrt round exactly like the
// numpy expressions they follow; the only places that cannot be bit-identical
// are sums over particles (order), libm calls (pow/atan/log) and BLAS-backed
// 3x3 products.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ocl {

// scalars derived on the host from p_array.E (sc.py:214-216, coord_transform.py:19-20,:61)
struct RefParams {
    double gamref;     // E / m_e_GeV
    double betaref;    // sqrt(1 - gamref^-2)
    double gb_ref;     // gamref * betaref
    double pref;       // m_e_eV * sqrt(gamref^2 - 1)
    double m_e_eV;
    double m_e_eV2;    // m_e_eV ** 2
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

// ---- coord_transform.py:57-96 (numpy branch) ---------------------------------
__device__ __forceinline__ void mad_momentum(const RefParams& rp, double xs, double ys, double delta,
                                             double& u0, double& u1, double& u2, double& gam, double& bet) {
    gam = (rp.betaref * delta + 1.0) * rp.gamref;                 // :68
    bet = sqrt(1.0 - 1.0 / (gam * gam));                          // :69  (gamma ** -2)
    double ratio = (gam * bet) / rp.gb_ref;
    double pz_rel = sqrt(ratio * ratio - xs * xs - ys * ys);      // :70
    double a = xs / pz_rel, b = ys / pz_rel;                      // :72
    double nrm = sqrt(a * a + b * b + 1.0);                       // :74 (2-norm along axis 1)
    u0 = a / nrm;                                                 // :77
    u1 = b / nrm;
    u2 = 1.0 / nrm;
}

__device__ __forceinline__ void mad_to_cart_momentum(const RefParams& rp, double xs, double ys, double delta,
                                                     double& px, double& py, double& pz) {
    double u0, u1, u2, gam, bet;
    mad_momentum(rp, xs, ys, delta, u0, u1, u2, gam, bet);
    px = u0 * gam * bet * rp.m_e_eV;                              // :93
    py = u1 * gam * bet * rp.m_e_eV;                              // :94
    pz = u2 * gam * bet * rp.m_e_eV;                              // :95
}

__device__ __forceinline__ Cart mad_to_cart(const RefParams& rp, double x, double xs, double y, double ys,
                                            double tau, double delta) {
    double u0, u1, u2, gam, bet;
    mad_momentum(rp, xs, ys, delta, u0, u1, u2, gam, bet);
    Cart c;
    c.x = x - u0 * bet * tau;                                     // :90
    c.y = y - u1 * bet * tau;                                     // :91
    c.z = -u2 * bet * tau;                                        // :92
    c.px = u0 * gam * bet * rp.m_e_eV;
    c.py = u1 * gam * bet * rp.m_e_eV;
    c.pz = u2 * gam * bet * rp.m_e_eV;
    return c;
}

// ---- coord_transform.py:16-54 (numpy branch) ---------------------------------
__device__ __forceinline__ void cart_to_mad(const RefParams& rp, const Cart& c, double& x, double& xs, double& y,
                                            double& ys, double& tau, double& delta) {
    double s = c.px * c.px + c.py * c.py + c.pz * c.pz;           // sum(u*u, 1)
    double gam = sqrt(1.0 + s / rp.m_e_eV2);                      // :27
    double bet = sqrt(1.0 - 1.0 / (gam * gam));                   // :28
    double p0 = sqrt(s);                                          // :31
    double u0 = c.px / p0, u1 = c.py / p0, u2 = c.pz / p0;        // :34
    double cdt = -c.z / (bet * u2);                               // :47
    x = c.x + bet * u0 * cdt;                                     // :48
    y = c.y + bet * u1 * cdt;                                     // :49
    delta = (gam / rp.gamref - 1.0) / rp.betaref;                 // :50
    tau = cdt;                                                    // :51
    xs = c.px / rp.pref;                                          // :52
    ys = c.py / rp.pref;                                          // :53
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

// order-1 interpolation with zero outside [0, n-1] (scipy.ndimage.map_coordinates
// order=1, mode='constant', cval=0 as used at sc.py:202-204): each corner value
// is multiplied by its x, y, z weights in that order; corners accumulate with z
// fastest.
__device__ __forceinline__ double trilinear(const double* __restrict__ F, int nx, int ny, int nz, double c0,
                                            double c1, double c2) {
    if (!(c0 >= 0.0 && c0 <= (double)(nx - 1) && c1 >= 0.0 && c1 <= (double)(ny - 1) && c2 >= 0.0 &&
          c2 <= (double)(nz - 1)))
        return 0.0;
    double f0 = floor(c0), f1 = floor(c1), f2 = floor(c2);
    int i0 = (int)f0, i1 = (int)f1, i2 = (int)f2;
    double t0 = c0 - f0, t1 = c1 - f1, t2 = c2 - f2;
    int j0 = min(i0 + 1, nx - 1), j1 = min(i1 + 1, ny - 1), j2 = min(i2 + 1, nz - 1);
    double w0[2] = {1.0 - t0, t0}, w1[2] = {1.0 - t1, t1}, w2[2] = {1.0 - t2, t2};
    int a[2] = {i0, j0}, b[2] = {i1, j1}, c[2] = {i2, j2};
    double acc = 0.0;
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const double* row = F + ((size_t)a[p] * ny + b[q]) * nz;
#pragma unroll
            for (int s = 0; s < 2; ++s) acc = acc + __ldg(row + c[s]) * w0[p] * w1[q] * w2[s];
        }
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
