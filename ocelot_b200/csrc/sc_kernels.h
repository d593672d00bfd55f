// sc_kernels.h -- host-callable launch wrappers for the space-charge kernels.
#pragma once
#include <cuda_runtime.h>
#include <cufft.h>
#include "sc_device.cuh"

namespace ocl {

// device-resident reduction state shared by the particle sweeps
struct ReduceState {
    double* part;          // [max_blocks][10] per-block partials
    unsigned int* ticket;  // [4] last-block tickets
    double* sums;          // [4]  OCL_SC_BUF_MOMENTUM
    double* emax;          // [6]  OCL_SC_BUF_EXTENT_MAX
    double* esum;          // [4]  OCL_SC_BUF_EXTENT_SUM
    double* geom;          // [24] geometry tap
    int max_blocks;
};

struct MeshDims {
    int nx, ny, nz;   // SpaceCharge.nmesh_xyz
    int mx, my, mz;   // padded FFT lengths
};

struct Draws {
    double scale;  // <= 0: random_mesh off
    double shift;
};

int particle_grid(long long n, int max_blocks);

void launch_momentum(const double* r, long long ld, long long n, RefParams rp, ReduceState rs, cudaStream_t st);
void launch_extent(const double* r, long long ld, const double* q, long long n, RefParams rp, ReduceState rs,
                   cudaStream_t st);
void launch_deposit(const double* r, long long ld, const double* q, long long n, RefParams rp, ReduceState rs,
                    MeshDims md, Draws dr, double* rho, cudaStream_t st);
void launch_green_table(ReduceState rs, MeshDims md, Draws dr, double* gtab, cudaStream_t st);
void launch_green_mirror(const double* gtab, MeshDims md, double* kpad, cudaStream_t st);
void launch_green_compact(const double* gtab, MeshDims md, double* k1, cudaStream_t st);
void launch_pad_rho(const double* rho, MeshDims md, double* pad, cudaStream_t st);
void launch_multiply(cufftDoubleComplex* rho_hat, const cufftDoubleComplex* k_hat, MeshDims md, cudaStream_t st);
void launch_crop_phi(const double* conv, ReduceState rs, MeshDims md, Draws dr, double* phi, cudaStream_t st);
void launch_field(const double* phi, ReduceState rs, MeshDims md, Draws dr, double* ex, double* ey, double* ez,
                  cudaStream_t st);
void launch_gather_kick(double* r, long long ld, long long n, RefParams rp, ReduceState rs, MeshDims md, Draws dr,
                        const double* ex, const double* ey, const double* ez, double dz, double* exyz_out,
                        int do_kick, cudaStream_t st);
void launch_mad_to_cart(const double* r, long long ld, long long n, RefParams rp, double* xp, long long ld_xp,
                        cudaStream_t st);
void launch_cart_to_mad(const double* xp, long long ld_xp, long long n, RefParams rp, double* r, long long ld,
                        cudaStream_t st);
// potential KAT helpers: steps given explicitly instead of derived from particles
void launch_green_table_steps(const double steps[3], MeshDims md, double* gtab, cudaStream_t st);
void launch_crop_phi_steps(const double* conv, const double steps[3], MeshDims md, double* phi, cudaStream_t st);

}  // namespace ocl
