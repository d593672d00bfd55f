/*
 * ocelot_sc.h -- C ABI of the B200-native 3D space-charge kick.
 *
 * This is the drop-in boundary for ONE hot path of Ocelot: SpaceCharge.apply()
 * (reference: ocelot/cpbd/sc.py:208-251 and everything it calls).  The
 * reference has no FFI for this path (it is numpy/scipy); these are the entry
 * points a reference-side binding (ctypes) binds instead of
 * calling sc.py's methods.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returning int returns 0 on success, non-zero on error;
 *     ocl_sc_last_error(h) (or ocl_sc_last_error(NULL) for create failures)
 *     returns a human-readable message.  There is NO CPU fallback.
 *   - "d_" pointers are device pointers on the handle's device, "h_" pointers
 *     are host pointers.  `stream` is a cudaStream_t passed as void* (NULL =
 *     the legacy default stream).  Calls are asynchronous w.r.t. the host
 *     unless stated; one handle = one device = one stream at a time; a handle
 *     is not thread-safe, different handles are independent.
 *   - particle layout = ParticleArray.rparticles (beam/particle.py:79-84):
 *     6 rows [x, x', y, y', tau, delta], row r at d_r + r*ld, fp64; charges
 *     q[n] fp64 (ParticleArray.q_array).
 *   - mesh_draws: NULL, or two doubles {scale in [1,1.1), shift in [-0.5,0.5)}
 *     = the two numpy draws of random_mesh mode (sc.py:175, :185), made by the
 *     host so the numpy global RNG stream is consumed exactly as the reference
 *     does.
 */
#ifndef OCELOT_SC_H
#define OCELOT_SC_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ocl_sc ocl_sc_t;

/* ABI version of this header (bumped on any signature change). */
int ocl_sc_abi_version(void);

/* Physical constants exactly as ocelot/common/globals.py:13-24 builds them:
 * out = {m_e_eV, m_e_GeV, epsilon_0, pi, speed_of_light}.  Host only. */
void ocl_sc_get_constants(double out[5]);

/* Padded FFT length used for a mesh of n points along one axis
 * (smallest power of two >= 2n-1; replaces the (2n-1) grid of sc.py:142). */
int ocl_sc_fft_size(int n);

/* Create / destroy a solver for mesh (nx,ny,nz) = SpaceCharge.nmesh_xyz
 * (sc.py:95) on CUDA device `device`.  max_particles sizes the internal
 * staging used by ocl_sc_kick_host only (0 = grow on demand). */
int ocl_sc_create(int device, int nx, int ny, int nz, long long max_particles, ocl_sc_t** out);
void ocl_sc_destroy(ocl_sc_t* h);
const char* ocl_sc_last_error(const ocl_sc_t* h);

/* One SpaceCharge.apply (sc.py:208-251) on device-resident particles, in
 * place.  dz == 0 returns immediately (sc.py:210-212).  No host
 * synchronisation; safe to capture in a CUDA graph. */
int ocl_sc_kick_device(ocl_sc_t* h, double* d_r, long long ld, const double* d_q, long long n,
                       double E_GeV, double dz, const double* mesh_draws, void* stream);

/* Same, for host arrays (numpy rparticles / q_array): H2D, kick, D2H in place,
 * synchronous.  This is the call a literal drop-in under track() makes. */
int ocl_sc_kick_host(ocl_sc_t* h, double* h_r, long long ld, const double* h_q, long long n,
                     double E_GeV, double dz, const double* mesh_draws);

/* Page-lock a caller-owned host range (numpy rparticles / q_array) so that ocl_sc_kick_host copies at PCIe
 * rate.  Not tied to a handle.  The CALLER owns the lifetime and must unregister before freeing the memory
 * (the Python binding does so from a weakref finaliser of the owning array).  ocl_sc_host_register returns
 * 0 = registered, 2 = already page-locked (nothing to undo), 1 = failed (copies still work, pageable). */
int ocl_sc_host_register(void* ptr, long long bytes);
int ocl_sc_host_unregister(void* ptr);

/* ---- staged form of the same kick, for particle-sharded multi-GPU runs ----
 * With one rank the five stages in order are exactly ocl_sc_kick_device.  Across ranks the two scalar
 * reductions happen either
 *   (a) INSIDE stage_momentum / stage_extent: after ocl_sc_mailbox_init the block that finishes the
 *       sweep's reduction pushes the rank's partial sums into every peer's mailbox over NVLink, waits
 *       for the peers' and folds them in rank order -- no collective call, no extra kernel launch; or
 *   (b) by the caller (e.g. NCCL without peer mappings): ocl_sc_defer_finish(h, 1), all-reduce the
 *       handle's small device buffers (ocl_sc_collective_buffer) after each of the two stages, then
 *       ocl_sc_stage_finish(h, which, ...) derives the frame (which = 0) / the mesh (which = 1).
 * The frame and the mesh are derived once per kick on the device and read by all later stages.
 *   stage_momentum : lab->Cartesian momenta (coord_transform.py:57-96), sums
 *                    -> buffer MOMENTUM {sum px, sum py, sum pz, count}   [SUM]
 *   stage_extent   : bunch frame (sc.py:224-239), rotate, gamma-stretch
 *                    (sc.py:172), extents and charge centroid (sc.py:173,
 *                    :181-182) -> buffer EXTENT_MAX {max xyz, -min xyz}   [MAX]
 *                    and buffer EXTENT_SUM {sum q*x, q*y, q*z, sum q}     [SUM]
 *   stage_deposit  : mesh geometry (sc.py:179-186), NGP deposit (sc.py:191-193)
 *                    -> buffer RHO (nx*ny*nz doubles)                     [SUM]
 *   stage_solve    : IGF + Hockney convolution (sc.py:109-168), staggered
 *                    field (sc.py:195-200); local to each rank
 *   stage_kick     : trilinear gather (sc.py:201-204), kick (sc.py:244-250),
 *                    Cartesian->lab (coord_transform.py:16-54), in place
 */
enum {
    OCL_SC_BUF_MOMENTUM = 0,   /* 4 doubles, reduce SUM */
    OCL_SC_BUF_EXTENT_MAX = 1, /* 6 doubles, reduce MAX */
    OCL_SC_BUF_EXTENT_SUM = 2, /* 4 doubles, reduce SUM */
    OCL_SC_BUF_RHO = 3,        /* nx*ny*nz doubles, reduce SUM */
    OCL_SC_BUF_EXTENT = 4,     /* the 10 doubles of EXTENT_MAX + EXTENT_SUM, contiguous (for all-gather) */
    /* slab mode only (after ocl_sc_slab_init); RHO and PHI then have nx_pad = sx*world planes */
    OCL_SC_BUF_RHO_SLAB = 5,   /* sx*ny*nz doubles: this rank's x-planes of rho (reduce-scatter output) */
    OCL_SC_BUF_XCHG_A = 6,     /* 2*nx_pad*fs doubles (complex): y-pass output / inverse-y input, chunk layout */
    OCL_SC_BUF_XCHG_B = 7,     /* 2*nx_pad*fs doubles (complex): x-pass lines of this rank, [nx_pad][fs] */
    OCL_SC_BUF_PHI_SLAB = 8,   /* sx*ny*nz doubles: this rank's x-planes of phi (all-gather input) */
    OCL_SC_BUF_PHI = 9,        /* nx_pad*ny*nz doubles: the gathered potential */
    /* longitudinal space charge (after ocl_sc_lsc_deposit) */
    OCL_SC_BUF_LSC_BINS = 10,      /* nb 64-bit integers (fixed-point CIC counts), reduce SUM as int64 */
    OCL_SC_BUF_LSC_SLICE_MAX = 11, /* 4 doubles, reduce MAX */
    OCL_SC_BUF_LSC_SLICE_SUM = 12  /* 5 doubles, reduce SUM */
};
int ocl_sc_collective_buffer(ocl_sc_t* h, int which, double** d_ptr, long long* count);

/* One-collective variant of the extent exchange: all-gather buffer EXTENT (10 doubles per rank)
 * into d_all[world][10], then fold it (max over the first 6, sum over the last 4) back into the
 * handle's EXTENT_MAX / EXTENT_SUM buffers. */
int ocl_sc_combine_extents(ocl_sc_t* h, const double* d_all, int world, void* stream);

/* Fused scalar exchanges over NVLink peer memory, replacing the MOMENTUM all-reduce (which = 0) and
 * the EXTENT all-gather + fold (which = 1) by one 32-thread kernel each: every rank passes the
 * device pointers at which all `world` (<= 8) ranks' mailboxes (OCL_SC_MAILBOX_DOUBLES zeroed
 * doubles of symmetric / IPC-mapped memory each) are mapped in ITS address space.  All ranks must
 * issue the same sequence of exchanges. */
#define OCL_SC_MAILBOX_DOUBLES 256
int ocl_sc_mailbox_init(ocl_sc_t* h, int rank, int world, void* const* peer_ptrs);
int ocl_sc_mailbox_exchange(ocl_sc_t* h, int which, void* stream);
/* Failure detection: an exchange whose peer does not answer within ~4 s gives up (the kernels must terminate) and
 * raises a device flag; *status = 0 ok, 1 / 2 / 3 = a momentum / extent / barrier-type exchange timed out and the
 * results since then are invalid.  synchronise != 0 waits for the stream last used first. */
int ocl_sc_mailbox_status(ocl_sc_t* h, int synchronise, int* status);
/* Fused reduction of the charge grid: the handle deposits into peer_rho[rank] (caller-owned symmetric
 * memory of at least the RHO buffer's size, mapped on every rank) and, after
 * ocl_sc_mailbox_exchange(h, 2, stream) -- a barrier: every rank's deposit is complete -- the first FFT
 * pass of ocl_sc_stage_solve / ocl_sc_slab_forward sums the `world` grids in rank order while loading
 * them over NVLink.  Replaces the all-reduce (or reduce-scatter) of RHO; no separate reduction kernel. */
int ocl_sc_set_peer_rho(ocl_sc_t* h, int rank, int world, void* const* peer_rho);
/* NVLS variant of the charge-grid reduction: `local_rho` is this rank's part of a symmetric allocation
 * of nx_pad*ny*nz doubles that is also mapped as one multicast range `multicast_rho` (e.g. torch
 * symmetric memory: buffer_ptrs[rank], multicast_ptr).  ocl_sc_nvls_reduce_rho then replaces the NCCL
 * all-reduce (redundant solve) or reduce-scatter (slab solve) of OCL_SC_BUF_RHO: barrier over the
 * mailbox, one kernel of multimem.ld_reduce / multimem.st (the NVSwitch sums each element once, so
 * every rank ends up with bit-identical sums), barrier.  Needs ocl_sc_mailbox_init. */
/* Once the mailbox and the multicast mapping are set and the handle is not in slab mode, ocl_sc_kick_device on every
 * rank's shard IS the complete sharded kick (both scalar exchanges inside the sweeps, rho summed in the switch, redundant
 * solve), captured in the library's own CUDA graph. */
int ocl_sc_set_multicast_rho(ocl_sc_t* h, void* local_rho, void* multicast_rho);
int ocl_sc_nvls_reduce_rho(ocl_sc_t* h, void* stream);

/* CUDA-graph support for callers that capture the staged kick themselves (e.g. together with
 * their NCCL collectives): with device params on, the stage kernels read E, dz and mesh draws
 * from a device block that ocl_sc_set_kick_params refreshes (a 1-block kernel, launched outside
 * the captured graph), so one captured graph serves every kick. */
int ocl_sc_use_device_params(ocl_sc_t* h, int on);
int ocl_sc_set_kick_params(ocl_sc_t* h, double E_GeV, double dz, const double* mesh_draws, void* stream);

/* ---- slab-decomposed Poisson solve for large meshes ----
 * Rank r of `world` owns sx = ceil(nx/world) x-planes of rho/phi and a chunk of
 * fs = ceil(My*(Mz/2+1)/world) (ky,kz) lines of the x pass.  Per kick, instead of
 * all-reduce(RHO) + ocl_sc_stage_solve:
 *     reduce-scatter RHO -> RHO_SLAB ; ocl_sc_slab_forward   (z, y passes of the slab -> XCHG_A)
 *     all-to-all XCHG_A -> XCHG_B     ; ocl_sc_slab_xpass     (x: forward FFT, * K_hat, inverse FFT)
 *     all-to-all XCHG_B -> XCHG_A     ; ocl_sc_slab_inverse   (inverse y, z -> PHI_SLAB)
 *     all-gather PHI_SLAB -> PHI      ; ocl_sc_slab_finish    (staggered field table)
 * replaces sc.py:135-168 + :195-200 exactly as the single-GPU solve does. */
int ocl_sc_slab_init(ocl_sc_t* h, int rank, int world);
/* Fused transposes: peer_a[w] / peer_b[w] = rank w's XCHG_A / XCHG_B buffers (2*nx_pad*fs doubles each, symmetric /
 * IPC-mapped memory) as mapped in THIS rank's address space.  ocl_sc_slab_forward then stores the y pass's output
 * straight into the peers' XCHG_B over NVLink and ocl_sc_slab_xpass its output into the peers' XCHG_A, each followed by
 * a mailbox barrier: the two all-to-alls of the sequence above disappear (the transfer overlaps the transform tile by
 * tile).  Needs ocl_sc_slab_init and ocl_sc_mailbox_init; world <= 8. */
int ocl_sc_set_peer_xchg(ocl_sc_t* h, int rank, int world, void* const* peer_a, void* const* peer_b);
/* Slab mode, fused all-gather: `local_phi` is this rank's part of a symmetric allocation of nx_pad*ny*nz doubles that is
 * also mapped as one multicast range `multicast_phi`.  ocl_sc_slab_inverse then writes its x-slab of the potential
 * through the multicast mapping (multimem.st: the NVSwitch replicates every store into all ranks' grids) and ends with
 * a mailbox barrier: the all-gather of PHI disappears. */
int ocl_sc_set_multicast_phi(ocl_sc_t* h, void* local_phi, void* multicast_phi);
int ocl_sc_slab_forward(ocl_sc_t* h, void* stream);
int ocl_sc_slab_xpass(ocl_sc_t* h, void* stream);
int ocl_sc_slab_inverse(ocl_sc_t* h, void* stream);
int ocl_sc_slab_finish(ocl_sc_t* h, const double* mesh_draws, void* stream);

int ocl_sc_stage_momentum(ocl_sc_t* h, const double* d_r, long long ld, long long n, double E_GeV, void* stream);
int ocl_sc_stage_extent(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n,
                        double E_GeV, const double* mesh_draws, void* stream);
int ocl_sc_defer_finish(ocl_sc_t* h, int on);
int ocl_sc_stage_finish(ocl_sc_t* h, int which, double E_GeV, const double* mesh_draws, void* stream);
int ocl_sc_stage_deposit(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n,
                         double E_GeV, const double* mesh_draws, void* stream);
int ocl_sc_stage_solve(ocl_sc_t* h, const double* mesh_draws, void* stream);
int ocl_sc_stage_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, double E_GeV, double dz,
                      const double* mesh_draws, void* stream);

/* Ordered deposit on / off (also OCL_SC_DETERMINISTIC=1 at create).  The default deposit adds each particle's
 * charge to its cell with an fp64 L2 atomic, i.e. in arrival order: rho jitters by ~1e-16 relative from run to
 * run.  With this switch the charges of a cell are added one after the other in ascending particle order
 * starting from 0.0 -- the order of np.bincount (sc.py:193) -- so rho is bit-identical from run to run and,
 * whenever every particle lands in the reference's cell, bit-identical to the reference's grid.  A debugging
 * aid (no reference equivalent; the reference is sequential): measured on B200: the whole kick takes 2.5x (1 M / 63^3: 0.42 ms) to 3x (12.5 M / 127^3: 3.1 ms) as long --
 * the sort-based deposit 9-12x, the exactly rounded momentum sweep 3-5x -- and is not graph-captured.  The same switch
 * makes the momentum sum follow np.mean's pairwise tree over exactly rounded momenta (frame and mesh steps become the
 * reference's bits). */
int ocl_sc_set_deterministic(ocl_sc_t* h, int on);

/* ---- stage taps (tests / diagnostics); each synchronises the stream last used ----
 * geometry out[24] = {T row-major [9], pav, gamma0, beta0, steps[3], X_off[3],
 *                     sum q, count, reserved[4]}  (sc.py:224-239, :179-185) */
int ocl_sc_get_geometry(ocl_sc_t* h, double out[24]);
int ocl_sc_get_rho(ocl_sc_t* h, double* h_out);          /* nx*ny*nz, C order, sc.py:193 */
int ocl_sc_get_phi(ocl_sc_t* h, double* h_out);          /* nx*ny*nz, sc.py:167-168 */
int ocl_sc_get_green(ocl_sc_t* h, double* h_out);        /* nx*ny*nz, sym_kernel sc.py:109-133 */
/* rest-frame field at the particles WITHOUT kicking them: Exyz (n,3) row-major
 * as SpaceCharge.el_field returns it (sc.py:201-205).  Runs all stages but the kick. */
int ocl_sc_field_at_particles(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, long long n,
                              double E_GeV, const double* mesh_draws, double* d_exyz, void* stream);

/* Stand-alone stages for known-answer tests (device pointers). */
int ocl_sc_mad_to_cartesian(ocl_sc_t* h, const double* d_r, long long ld, long long n, double E_GeV,
                            double* d_xp, long long ld_xp, void* stream);   /* coord_transform.py:57-96 */
int ocl_sc_cartesian_to_mad(ocl_sc_t* h, const double* d_xp, long long ld_xp, long long n, double E_GeV,
                            double* d_r, long long ld, void* stream);       /* coord_transform.py:16-54 */
/* potential(q, steps) of sc.py:135-168 for a host rho[nx*ny*nz] and steps[3]; result to h_phi. */
int ocl_sc_potential_host(ocl_sc_t* h, const double* h_rho, const double steps[3], double* h_phi);

/* ---- the steps either side of the kick in the reference's tracking loop (track.py:470-482),
 * so that a bunch can stay resident between kicks ---- */
/* X <- R X + T:XX + B in place (TransferMap.mul_p_array, transformations/transfer_map.py:42-53;
 * SecondTM.t_apply, transformations/second_order.py:31-39 with tm_utils.py:54-55).
 * R[36] row-major; B[6] or NULL; T[216] (index a*36 + j*6 + k) or NULL for a first-order map. */
int ocl_sc_map_apply(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* R, const double* B,
                     const double* T, void* stream);
/* The seven scalars ocl_sc_cavity_apply needs, from the cavity's voltage v [GV], phase phi [deg], frequency [Hz],
 * the beam energy E [GeV] and the slice (delta_length of length; delta_length < 0 or NaN: the whole cavity):
 * coef = {c1, c2, beta0*k, phi [rad], T566, T556, T555}, *mode = 1 (full map) or 2 (drift-like: non-physical final
 * energy), *delta_e = V cos(phi).  Specification: CavityTM.map4cav, transformations/cavity.py:29-128. */
int ocl_sc_cavity_coefficients(double v, double phi_deg, double freq, double E_GeV, double delta_length, double length,
                               double coef[7], int* mode, double* delta_e);
/* RF cavity body, CavityTM.map4cav (transformations/cavity.py:29-128): X <- R X + B, then
 *   delta <- delta0*c[0] + c[1]*(cos(c[3] - c[2]*tau0) - cos(c[3]))        (cavity.py:81-84)
 *   tau   += c[4]*delta0^2 + c[5]*tau0*delta0 + c[6]*tau0^2                 (cavity.py:126)
 * with c[7] = {E0 b0/(E1 b1), V b0/(E1 b1), b0 k, phi, T566, T556, T555} computed by the host from the
 * reference's scalar formulas; mode 1 = full, mode 2 = drift-like branch (cavity.py:67-69). */
int ocl_sc_cavity_apply(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* R, const double* B,
                        const double* c, int mode, void* stream);
/* Aperture cut on a resident bunch with ordered stream compaction (RectAperture / EllipticalAperture,
 * physics_proc.py:341-390; ParticleArray.delete_particles, beam/particle.py:323-333).  kind 0: a particle is lost if
 * row `row` is outside [params[0], params[1]]; kind 1: if ((x-params[2])/params[0])^2 + ((y-params[3])/params[1])^2 > 1.
 * Survivors (six rows, charge, id) are written in order to the *_out buffers (distinct from the inputs), the ids of
 * the lost particles in order to d_lost_out (n entries; may be NULL; d_ids NULL = ids are the indices).  Synchronises
 * `stream` and returns the survivor count in *n_out. */
int ocl_sc_aperture_cut(ocl_sc_t* h, const double* d_r, long long ld, const double* d_q, const long long* d_ids,
                        long long n, int kind, int row, const double* params, double* d_r_out, long long ld_out,
                        double* d_q_out, long long* d_ids_out, long long* d_lost_out, long long* n_out, void* stream);

/* First and centred second moments of get_envelope's default path (beam/analysis.py:72-76,
 * :121-166): h_out[18] = {x, px, y, py, tau, p, xx, xpx, pxpx, yy, ypy, pypy, tautau, pp, xy, pxpy,
 * xpy, ypx} with the reference's px, py correction factor applied.  Synchronous. */
int ocl_sc_beam_moments(ocl_sc_t* h, const double* d_r, long long ld, long long n, double* h_out, void* stream);
/* The same with the result left in DEVICE memory d_out and no host synchronisation (a resident tracking loop records
 * every step's moments on the device and reads them back once at the end; the reference's track() also only returns
 * them when tracking is over, track.py:482,504).  d_out[0..17] as above; d_q != NULL adds d_out[18] = sum q
 * (Twiss.q, analysis.py:81). */
int ocl_sc_beam_moments_device(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* d_q,
                               double* d_out, void* stream);

/* ---- longitudinal space charge: class LSC (ocelot/cpbd/sc.py:261-599), the 1-D sibling of the
 * kick.  One LSC.apply = ocl_sc_lsc_stats (sweep A, synchronous) -> the host derives the 1-D grid like
 * s_to_cur (beam/analysis.py:293-333) -> ocl_sc_lsc_kick (deposit, smoothing, impedance, wake, energy
 * kick; asynchronous).  Only row 5 (delta) of the particles is modified (sc.py:599). ---- */
/* h_out[8] = {n, mean(tau), sum (tau-mean)^2, min(tau), max(tau), sum(q), sum(x), sum(y)}
 * (np.mean / np.std / np.sum of sc.py:576-577, :590; np.min / np.max of analysis.py:296-297). */
int ocl_sc_lsc_stats(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* d_q, double h_out[8],
                     void* stream);
/* params[17] = {slice_min, slice_max (sc.py:579-580), x_shift, y_shift (centroid, shifts of the slice
 * sums), a, ds, nb (grid x_j = j*ds + a, analysis.py:318-322), sigma_s, K (Gaussian taps -K..K,
 * analysis.py:330-333; K < 0: no smoothing), q = sum(q_array), v, gamma, dz, 1 + K_max^2 fill/2
 * (sc.py:460-463), pc_ref [GeV] (sc.py:597), step_profile (0/1), n_total (all ranks)}. */
int ocl_sc_lsc_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* params, void* stream);
/* The two halves of ocl_sc_lsc_kick for a particle-sharded bunch: after the deposit the caller
 * all-reduces OCL_SC_BUF_LSC_BINS (int64 SUM: exact), _SLICE_MAX (MAX) and _SLICE_SUM (SUM). */
int ocl_sc_lsc_deposit(ocl_sc_t* h, const double* d_r, long long ld, long long n, const double* params, void* stream);
int ocl_sc_lsc_solve_kick(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* params, void* stream);
/* The same kick with no host synchronisation (CUDA-graph capturable): sweep A, then a one-thread kernel
 * derives the grid on the device with the arithmetic of s_to_cur, then deposit, solve and kick read it
 * from device memory.  hostp[10] = {gamma, v, pc_ref [GeV], dz, 1 + K_max^2 fill/2, bounds[0], bounds[1],
 * smooth_param, step_profile (0/1), n_total (0: n)}.  Sized for grids of up to 8192 points and 127
 * smoothing taps; a kick whose grid does not fit is skipped and reported by the next ocl_sc_lsc_* call
 * on the handle (the synchronous form has no such limit). */
int ocl_sc_lsc_kick_async(ocl_sc_t* h, double* d_r, long long ld, long long n, const double* d_q, const double* hostp,
                          void* stream);
/* scalars the device derived for the last asynchronous kick, in the layout of ocl_sc_lsc_kick's params
 * (n_total not filled).  Synchronises; makes ocl_sc_lsc_get_profile usable with nb = out[6]. */
/* Outcome of the asynchronous kicks issued so far: *status = 0 (all applied), 1 (a kick was SKIPPED because its
 * grid exceeds the asynchronous form's capacity or is degenerate), 2 (skipped: packed deposit word too narrow).
 * A skipped kick leaves the particles untouched, so the caller can redo it with the synchronous form
 * (ocl_sc_lsc_stats + ocl_sc_lsc_kick).  synchronise != 0 waits for the stream last used first; the flag is
 * cleared by the call. */
int ocl_sc_lsc_async_status(ocl_sc_t* h, int synchronise, int* status);
int ocl_sc_lsc_last_params(ocl_sc_t* h, double out[17]);
/* taps of the last LSC kick (any pointer may be NULL): current profile I(s_j) [A] (s_to_cur's B[:,1]),
 * wake W(s_j)*q [V] (sc.py:592), transverse size sigma or rb used by the impedance (sc.py:584-589). */
int ocl_sc_lsc_get_profile(ocl_sc_t* h, int nb, double* h_current, double* h_wake, double* h_sigma);

/* Per-stage device timers.  enable=1 records CUDA events around each stage of
 * every following kick; get returns the last kick's milliseconds:
 * out[8] = {momentum, extent, deposit, green+fft, field, kick, total, reserved}. */
int ocl_sc_enable_timers(ocl_sc_t* h, int enable);
int ocl_sc_get_timers(ocl_sc_t* h, double out[8]);

/* Number of kernel launches (own kernels + cuFFT execs) issued by the handle so far. */
long long ocl_sc_launch_count(const ocl_sc_t* h);

#ifdef __cplusplus
}
#endif
#endif /* OCELOT_SC_H */
