/*
 * jitr_b200 -- C ABI of the B200-native calculable R-matrix hot path.
 *
 * The reference (beykyle/jitr) is pure Python: its "FFI" for this path is the Python call
 * surface Python -> numba @njit -> LAPACK.  Each entry point below states the reference
 * interface it replaces (file:line relative to /root/reference/src/jitr).  The host-side mirror
 * of the reference API (jitr_b200.Solver, jitr_b200.xs.elastic.*Workspace, ...) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - complex128 values are interleaved (re, im) doubles, C-contiguous, last index fastest;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - return value: 0 on success, a JITR_B200_ERR_* code otherwise (nothing was launched, or
 *     the launch failed; jitr_b200_last_error() gives a message);
 *   - per-system numerical failures (non-finite input, zero pivot) never abort the batch: they
 *     are reported in the `status` array (JITR_B200_STATUS_*), like numba's LinAlgError per call.
 */
#ifndef JITR_B200_H
#define JITR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define JITR_B200_OK 0
#define JITR_B200_ERR_BAD_ARG 1      /* ValueError / AssertionError of the reference (shape checks) */
#define JITR_B200_ERR_UNSUPPORTED 2  /* size outside the compiled kernel set */
#define JITR_B200_ERR_CUDA 3         /* CUDA runtime error (no device, launch failure, ...) */

#define JITR_B200_STATUS_OK 0
#define JITR_B200_STATUS_NONFINITE 1   /* NaN/Inf met in the system matrix or result */
#define JITR_B200_STATUS_SINGULAR 2    /* exact zero pivot */

/* built-in optical-potential parameterisations evaluated inside the fused kernel
 * (optical_potentials/{omp,kduq,chuq}.py) */
#define JITR_B200_MODEL_SHAPE 0  /* 17 canonical shape numbers, see jitr_b200_elastic_sweep_params */
#define JITR_B200_MODEL_KDUQ 1   /* 40 global Koning-Delaroche numbers (neutron rows padded with 0) */
#define JITR_B200_MODEL_CHUQ 2   /* 22 global Chapel-Hill numbers */
#define JITR_B200_MODEL_LOCAL 3  /* 15 LocalOpticalPotential numbers */
#define JITR_B200_MODEL_WLH 4    /* 46 global Whitehead-Lim-Holt numbers (optical_potentials/wlh.py) */
#define JITR_B200_MODEL_DOM 6     /* 19 global parameters of the local dispersive optical model (optical_potentials/dom.py), mapped on the
                                    device incl. the analytic surface-dispersion integral; needs JITR_B200_ET_EF in the energy rows */
#define JITR_B200_MODEL_SHAPE_PAIR 5 /* 18 numbers (17 canonical + Vw, the real depth at the imaginary-volume geometry) PER
                                        (sample, energy): params is [n_samples][n_energies][n_params]; for models whose
                                        energy-dependent parameter map is evaluated on the host (optical_potentials/dom.py) */

/* layout of one row of the per-energy table (doubles) */
#define JITR_B200_ET_A 0        /* dimensionless channel radius a = k * R_ch */
#define JITR_B200_ET_ECM 1      /* E0: centre-of-mass energy (MeV) */
#define JITR_B200_ET_K 2        /* k0 (1/fm) */
#define JITR_B200_ET_ELAB 3     /* laboratory energy (MeV), enters the KDUQ/CHUQ energy dependence */
#define JITR_B200_ET_RCH 4      /* channel radius in fm  (r_n = x_n * R_ch) */
#define JITR_B200_ET_AT 5       /* target mass number A */
#define JITR_B200_ET_ZT 6       /* target charge Z */
#define JITR_B200_ET_ZP 7       /* projectile charge (0 or 1 for KDUQ/CHUQ) */
#define JITR_B200_ET_AP 8       /* projectile mass number */
#define JITR_B200_ET_RSCALE 9   /* LocalOpticalPotential radius factor (omp.py:106-113) */
#define JITR_B200_ET_A13 10     /* A^(1/3), A^(-1/3), A^(-2/3), A^(-5/3): host pow() so the global */
#define JITR_B200_ET_AM13 11    /* parameter maps see bit-identical powers (kduq.py:455-507)       */
#define JITR_B200_ET_AM23 12
#define JITR_B200_ET_AM53 13
#define JITR_B200_ET_EF 14      /* effective Fermi energy of the (projectile, target) pair (MeV), DOM only (reactions/reaction.py:528-538) */
#define JITR_B200_ET_STRIDE 16

const char* jitr_b200_last_error(void);
int jitr_b200_version(void);
/* 1 if an sm_100 device is present and usable, 0 otherwise (never raises) */
int jitr_b200_device_ok(void);

/*
 * Spin-1/2 (x) spin-0 elastic S-matrix sweep with BUILT-IN potentials, fully fused:
 * per (sample, energy) the potential forms are evaluated on the mesh from the parameter row,
 * then for l = 0..lmax and j = l +- 1/2 the system  A = T(l) - 1 + (V_c + V_C + (l.s) V_so)/E0  is
 * assembled, LU-factorised (complex128, partial pivoting), R = b^T A^-1 b / a^2 and
 * S = (H- - a R H-')/(H+ - a R H+') are formed.  The loop over l stops after the first l >= 1
 * with |1-S+| < tol and |1-S-| < tol (that l is kept); entries beyond it are left 0.
 *
 * Replaces, batched over samples x energies:
 *   xs/elastic.py:104-183   IntegralWorkspace.smatrix        (loop, early break :178-181)
 *   rmatrix/rmatrix.py:126-246 Solver.interaction_matrix/.solve
 *   rmatrix/core.py:7-60    rmatrix_with_inverse, solve_smatrix_with_inverse
 *   optical_potentials/potential_forms.py:40-153, kduq.py:100-197,370-559, chuq.py:92-227,
 *   wlh.py:61-111,251-383, omp.py:54-149   evaluation of (central, spin_orbit, coulomb) on the mesh
 *
 *   nbasis      N, number of Lagrange-Legendre functions, 1..64 (N <= 40 takes the assembly-free fast path)
 *   That        [N][N] real: a-independent kinetic+Bloch matrix (quadrature.py:197-218 with a=1,l=0)
 *   mesh_x      [N] abscissae on (0,1);  bvec [N] basis functions at the boundary (rmatrix.py:36-42)
 *   etab        [n_energies][JITR_B200_ET_STRIDE]
 *   asym        [n_energies][lmax+1][4] complex: H+, H-, H+', H-' at a   (reactions/system.py:182-213)
 *   params      [n_samples][n_params] rows of `model` parameters
 *   tol         smatrix_abs_tol (xs/elastic.py:49); tol <= 0 disables the early break
 *   S_out       [n_samples][n_energies][2][lmax+1] complex  (S+ then S-)
 *   nl_out      [n_samples][n_energies] int32: number of partial waves kept (last_l + 1)
 *   status      [n_samples][n_energies] int32 JITR_B200_STATUS_*
 */
int jitr_b200_elastic_sweep_params(int nbasis, int lmax, int model, long long n_samples, int n_energies,
                                   const double* That, const double* mesh_x, const double* bvec,
                                   const double* etab, const double* asym, const double* params,
                                   int n_params, double tol, double* S_out, int* nl_out, int* status,
                                   void* stream);

/*
 * Same sweep with USER potentials given as mesh values (MeV on r_n = x_n R_ch):
 *   central, spin_orbit [n_samples][n_energies][N] complex; coulomb [n_samples][n_energies][N]
 *   complex or NULL; spin_orbit may be NULL (xs/elastic.py:96-102: omitted == zeros).
 * Replaces xs/elastic.py:104-183 called once per sample by the user loop
 * (examples/notebooks/kduq_uq_demo.ipynb cell 16).
 */
int jitr_b200_elastic_sweep_tensors(int nbasis, int lmax, long long n_samples, int n_energies,
                                    const double* That, const double* mesh_x, const double* bvec,
                                    const double* etab, const double* asym, const double* central,
                                    const double* spin_orbit, const double* coulomb, double tol,
                                    double* S_out, int* nl_out, int* status, void* stream);

/*
 * Built-in potential forms on the mesh, stand-alone (the same device code the fused sweep inlines):
 * SingleChannelOpticalModel.evaluate, optical_potentials/omp.py:26-43, kduq.py:541-559, chuq.py:92-137.
 *   central, spin_orbit [n_samples][n_energies][N] complex;  coulomb [n_samples][n_energies][N] real
 */
int jitr_b200_eval_potentials(int nbasis, int model, long long n_samples, int n_energies, const double* mesh_x,
                              const double* etab, const double* params, int n_params, double* central,
                              double* spin_orbit, double* coulomb, void* stream);

/*
 * Elastic observables from the truncated S-matrices:  xs/elastic.py:332-381
 * (differential_elastic_xs) and :310-329 (integral_elastic_xs), batched.
 *   S, nl       as produced by the sweeps
 *   P, P1       [n_energies][lmax+1][n_angles] Legendre tables P_l, P^1_l (xs/elastic.py:245-246)
 *   phase       [n_energies][lmax+1] complex  exp(2 i sigma_l)/(2 i k)
 *   f_c         [n_energies][n_angles] complex Coulomb amplitude (0 for neutral)
 *   kvec        [n_energies] wavenumbers
 *   dsdo, Ay, Q [n_samples][n_energies][n_angles] (any of Ay, Q may be NULL)
 *   tot_rxn     [n_samples][n_energies][2]  (sigma_tot, sigma_rxn) in mb
 */
int jitr_b200_elastic_observables(int lmax, int n_angles, long long n_samples, int n_energies,
                                  const double* S, const int* nl, const double* P, const double* P1,
                                  const double* phase, const double* f_c, const double* kvec,
                                  double* dsdo, double* Ay, double* Q, double* tot_rxn, void* stream);

/*
 * General batched R-matrix solve (coupled channels, local and/or nonlocal interaction):
 *   A_b = F + V_b,   X_b = A_b^-1 [b (x) e_j],   R, S, u'_ext, optional coefficients.
 * Replaces rmatrix/rmatrix.py:180-246 Solver.solve (one call per system in the reference),
 * rmatrix/core.py:7-82, quadrature/kernel.py:180-214 (local / nonlocal weighting).
 *
 *   free        [n][n] complex (free_stride = 0, shared) or [batch][n][n] (free_stride = n*n),
 *               n = nch * nbasis                                 (rmatrix.py:106-124)
 *   local       [batch][nch][nch][N] complex mesh values or NULL      (rmatrix.py:141-158)
 *   nonlocal    [batch][nch][nch][N][N] complex mesh values or NULL   (rmatrix.py:160-176)
 *   interaction [batch][n][n] complex precomputed interaction matrix or NULL (used instead of
 *               local/nonlocal when given; rmatrix.py:206-214)
 *   wsqrt       [N] sqrt of the quadrature weights (kernel.py:201)
 *   sys         [batch_or_1][4] doubles: a, E0, k0, unused  (sys_stride = 0 or 4)
 *   bvec        [N];  asym [batch_or_1][4][nch] complex (asym_stride = 0 or 4*nch complex)
 *   weights     [nch] incoming-channel weights (rmatrix.py:203-205)
 *   R, S        [batch][nch][nch] complex;  uext [batch][nch] complex
 *   coeffs      [batch][nch][N] complex or NULL              (core.py:63-82)
 *   workspace   device scratch of jitr_b200_solve_workspace_bytes(nbasis, nch, batch, coeffs != NULL) bytes
 *               (NULL when that is 0); only read and written inside the call
 */
int jitr_b200_solve_batched(int nbasis, int nch, long long batch, const double* free_matrix,
                            long long free_stride, const double* local, const double* nonlocal_,
                            const double* interaction, const double* wsqrt, const double* sys,
                            int sys_stride, const double* bvec, const double* asym, long long asym_stride,
                            const double* weights, double* R, double* S, double* uext, double* coeffs,
                            double* workspace, int* status, void* stream);

/*
 * The same solve for a batch that mixes partial waves (or energies): `free_matrices` holds n_free free matrices [n_free][n][n] and
 * system b uses number free_index[b] (device int32 array of `batch` entries) -- e.g. all (sample, l) systems of a nonlocal sweep in ONE
 * call, free_index = l, with per-system asymptotics (asym_stride = 4 nch).  Everything else as jitr_b200_solve_batched.
 */
int jitr_b200_solve_batched_indexed(int nbasis, int nch, long long batch, const double* free_matrices, int n_free,
                                    const int* free_index, const double* local, const double* nonlocal_,
                                    const double* interaction, const double* wsqrt, const double* sys, int sys_stride,
                                    const double* bvec, const double* asym, long long asym_stride, const double* weights,
                                    double* R, double* S, double* uext, double* coeffs, double* workspace, int* status,
                                    void* stream);

/*
 * Weighted misfit of the angular distributions against data, reduced on the device (the batched likelihood of a
 * calibration loop, e.g. the log-likelihood of the reference's quickstart notebook): same inputs as
 * jitr_b200_elastic_observables, plus
 *   y, w, scale [n_energies][n_angles]   data, weights (1/sigma^2 for chi^2) and an optional factor applied to
 *                                         dsigma/dOmega before the comparison (1/Rutherford for ratio data; NULL = 1)
 *   misfit      [n_samples][n_energies]  sum_theta w (scale * dsdo - y)^2
 *   dsdo        optional output as in jitr_b200_elastic_observables (NULL: nothing but misfit and tot_rxn leaves
 *               the kernel -- 24 bytes per (sample, energy) instead of 8 n_angles)
 */
int jitr_b200_elastic_misfit(int lmax, int n_angles, long long n_samples, int n_energies, const double* S,
                             const int* nl, const double* P, const double* P1, const double* phase,
                             const double* f_c, const double* kvec, const double* y, const double* w,
                             const double* scale, double* misfit, double* tot_rxn, double* dsdo, void* stream);

/*
 * Percentiles over the sample axis, on the device: numpy.percentile(values, q, axis=0) with the default linear
 * interpolation (the uncertainty bands of the reference's kduq_uq_demo notebook, cell 18).
 *   values [n_samples][n_columns] f64 device (e.g. dsdo viewed as [S][E * n_angles]);  q: HOST array of nq <= 8
 *   percentages in [0, 100];  out [nq][n_columns] f64 device.  Exact order statistics (radix selection).
 */
int jitr_b200_percentiles(const double* values, long long n_samples, long long n_columns, const double* q_host, int nq,
                          double* out, void* stream);

/*
 * Quasi-elastic (p,n) DWBA, the steps after the distorted-wave solves (xs/quasielastic_pn.py).
 *
 * jitr_b200_qe_overlap: T_lj = scale * sum_n x_p[n] (U1c[n] + (l.s)_j U1so[n]) x_n[n]   (:317-322)
 *   xp, xn      [nj][batch][nbasis] complex solution coefficients of the entrance / exit channel
 *               (the coeffs output of jitr_b200_solve_batched, systems ordered j-major)
 *   U1c, U1so   [batch][nbasis] complex isovector transition potentials (U1so may be NULL)
 *   lds0, lds1  (l.s) of j = l+1/2 and l-1/2;  scale = 1 / (R_ch k_p k_n)
 *   out         complex, element (b, j) at  b * out_sample_stride + j * out_j_stride  (complex units)
 *
 * jitr_b200_qe_xs: dsigma/dOmega = factor * sum_{m,m'} |sum_{l < lmax, j} G[m][m'][l][j][theta] T[b][l][j]|^2  (:375-392)
 *   T [batch][lmax+1][2] complex,  G [2][2][lmax+1][2][n_angles] complex (:198-217),
 *   factor = 10 * xs_factor,  out [batch][n_angles] in mb/sr
 */
int jitr_b200_qe_overlap(int nbasis, long long batch, int nj, const double* xp, const double* xn, const double* U1c,
                         const double* U1so, double lds0, double lds1, double scale, double* out,
                         long long out_sample_stride, long long out_j_stride, void* stream);
int jitr_b200_qe_xs(int lmax, int n_angles, long long batch, const double* T, const double* G, double factor,
                    double* out, void* stream);

/*
 * Perey-Buck-type nonlocal kernels generated on the device: the overflow-free form of perey_buck_nonlocal,
 * optical_potentials/potential_forms.py:14-19  [ exp(-(r^2+r'^2)/beta^2) 2 i^l z j_l(-i z) / (beta sqrt(pi)),  2 i^l z j_l(-i z) = 2 z i_l(z) ],
 * times a Woods-Saxon depth at the mean radius, for every l = 0..lmax of every sample at once:
 *   U_l(r_n, r_m) = -(V + iW) f((r_n + r_m)/2; R, a) * 2 z e^{-z} i_l(z) exp(z - (r_n^2 + r_m^2)/beta^2) / (beta sqrt(pi)),  z = zfac r_n r_m / beta^2
 *   params [n_samples][5]: V, W, beta, R, a  (a <= 0: no Woods-Saxon factor -> V = -1, W = 0 gives perey_buck_nonlocal itself)
 *   zfac   2 for the Gaussian nonlocality (exponent -(r - r')^2/beta^2);  2 pi reproduces the reference's literal z
 *   rgrid  [N] physical mesh (fm);   out: element (sample s, l, n, m) at  s * sample_stride + l * l_stride + n * N + m  (complex
 *   units; 0, 0 = the dense [n_samples][lmax+1][N][N] layout; sample_stride = N*N, l_stride = n_samples*N*N gives [lmax+1][n_samples][N][N],
 *   one contiguous `nonlocal` batch of jitr_b200_solve_batched per partial wave).   lmax <= 64.
 */
int jitr_b200_perey_buck_kernels(int nbasis, int lmax, long long n_samples, const double* rgrid, const double* params,
                                 double zfac, double* out, long long sample_stride, long long l_stride, void* stream);

/* bytes of device scratch jitr_b200_solve_batched needs for this problem (0: none; -1: bad arguments) */
long long jitr_b200_solve_workspace_bytes(int nbasis, int nch, long long batch, int coeffs);

/*
 * FP64 pipe probes (roofline denominators; MEASURED_PEAKS.json has no FP64 entry).
 *   which = 0: dependent-chain-free DFMA throughput (TFLOP/s written to *tflops_out)
 *           1: mma.sync m8n8k4 f64 (DMMA) throughput
 *           2: mixed, `warps_dmma` of every 8 warps issue DMMA, the rest DFMA
 *           3: latencies (cycles per 1024 dependent ops: DFMA, REDUX, SHFL, DDIV) -> lat_out_host[4]
 */
int jitr_b200_fp64_probe(int which, int iters, int warps_dmma, double* tflops_out_host,
                         long long* lat_out_host);

/*
 * Library options.  JITR_B200_OPT_SYMMETRIC (default 1): the fused sweep first eliminates each system
 * with threshold pivoting on the diagonal (|a_kk| >= max_i |a_ik| / 64), which preserves the complex
 * symmetry of  That + diag(d)  (half the trailing update, no triangular solve, no row interchanges), and
 * re-solves a system with partial pivoting the first time a diagonal entry fails the test.  0 = always
 * use partial pivoting.  Both give the reference's S within its tolerance (tests/test_gpu_parity.py).
 */
#define JITR_B200_OPT_SYMMETRIC 0
/* JITR_B200_OPT_ASSUME_SYMMETRIC_KERNEL (default 0): 1 = the caller vouches that the nonlocal kernels / interaction matrices
 * handed to jitr_b200_solve_batched are symmetric in (r, r'), as every physical nonlocal potential is; the packed solver then
 * reads their lower triangle only (half the HBM traffic).  0 = the upper triangle is read too and compared element by element;
 * a system that is not symmetric is solved by the partial-pivoting kernel instead. */
#define JITR_B200_OPT_ASSUME_SYMMETRIC_KERNEL 1
/* JITR_B200_OPT_PACKED_FIRST_PANEL (default 0): 1 = single-channel systems with 33..64 basis functions and no coefficients eliminate
 * their first eight columns in registers (panel columns of V by bulk-async copies) and accumulate the Schur complement straight from
 * global memory, keeping only its lower triangle in shared memory (seven systems per SM at N = 60 instead of six).  Measured on B200 it
 * is SLOWER than staging the whole lower triangle by bulk-async copies first (1.33e7 vs 1.70e7 solves/s on BASELINE config 4: the
 * accumulator loads expose the HBM latency once per block row), so it is off by default and kept as a measured alternative. */
#define JITR_B200_OPT_PACKED_FIRST_PANEL 2
/* JITR_B200_OPT_OBSERVABLES_DMMA (default 1): the Legendre sums of the elastic observables as a GEMM on the FP64 tensor path (four
 * samples per DMMA row block); 0 = the round-1 kernel (one warp per sample, scalar FMAs), which also serves lmax > 63. */
#define JITR_B200_OPT_OBSERVABLES_DMMA 3
#define JITR_B200_OPT_COUNT 4
int jitr_b200_set_option(int key, int value);
/* systems the symmetric path handed to the partial-pivoting path since load (-1: no device) */
long long jitr_b200_sym_fallback_count(void);

/*
 * Diagnostic for the parity gate: when a device array mask[n_samples * n_energies] of 64-bit words is registered
 * (NULL = off, the default), the fused sweep ORs bit (2 l + (j == l - 1/2)) into mask[pair] for every system the
 * symmetric path handed to partial pivoting, so a test can restrict its error statistics to those systems.
 */
int jitr_b200_debug_set_fallback_mask(unsigned long long* mask);

/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long jitr_b200_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* JITR_B200_H */
