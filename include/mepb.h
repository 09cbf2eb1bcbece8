/*
 * mepb.h — C ABI of the B200-native per-k-point Wannier-interpolation path (libmepb.so).
 *
 * The reference (cMax347/MEPinterp, Fortran 2003 + MPI) has no plugin / FFI layer; the seam this library
 * replaces is the Fortran call site in core_worker:
 *      src/core.f90:165-270   the k-loop body: get_wann_interp (src/wann_interp.f90:56-108) followed by the
 *                             per-k response functions (src/kubo.f90, kubo_mep.f90, mep_niu.f90, photo.f90)
 *                             whose results are += accumulated
 *      src/core.f90:276-292   fermi_selector + integrate_over_k_space (reduce, normalise)
 *      src/band_calc.f90:67-75 get_wann_interp on an explicit k list (band mode)
 * A Fortran host binds these entry points with an `interface ... bind(C)` block (INTEGRATION.md shows it);
 * everything is plain pointers and sizes: column-major, caller-owned, FP64, complex = interleaved (re,im), so
 * Fortran arrays H_tb(n,m,R), R_vect(3,R), wf_centers(3,n), a_latt(3,3) pass unchanged.
 *
 * No CPU fallback exists: every compute entry point runs hand-written sm_100a kernels and returns
 * MEPB_ERR_CUDA when no usable device is present.
 */
#ifndef MEPB_H
#define MEPB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MEPB_OK                 0
#define MEPB_ERR_ARG          (-1)   /* NULL / out-of-range argument                                         */
#define MEPB_ERR_CUDA         (-2)   /* no device, or a CUDA call failed (mepb_last_error has the text)      */
#define MEPB_ERR_UNSUPPORTED  (-4)   /* nw/nR outside kernel limits, or a model update that would change buffer shapes */
#define MEPB_ERR_EIG          (-5)   /* eigensolver did not converge for some k (reference: zheevd info/=0 -> stop,
                                        src/matrix_math.f90:224)                                              */
#define MEPB_ERR_STATE        (-6)   /* call order (finalize before accumulate, ...)                         */
#define MEPB_ERR_NCCL         (-7)   /* libnccl.so.2 not loadable, or an NCCL call failed (reference: MPI ierr /= 0 -> stop,
                                        src/mpi_comm.f90:124)                                                 */

/* == what read_tb_basis (src/file_io.f90:800) + init_parameters (src/input_paras.f90:79) hand to core_worker */
typedef struct {
    int32_t nw;               /* num_wann                                                                     */
    int32_t nR;               /* number of R points                                                           */
    const double *R;          /* (3,nR)   R_vect as stored: fractional                       file_io.f90:724   */
    const double *H;          /* (nw,nw,nR) complex, Hartree                                 file_io.f90:732-741 */
    const double *r;          /* (3,nw,nw,nR) complex position operator <0n|r|Rm> in Bohr (file_io.f90:782-790, 644-652), or NULL.
                                 Non-NULL switches on the Berry connection / curvature branch: A_ka, Om_kab Fourier sums
                                 (wann_interp.f90:288-375), their gauge transformation (:465-658), connection-corrected
                                 velocities (:406-444), mep_*_CS, kubo_opt_tens                                   */
    const double *wf_centers; /* (3,nw) fractional or NULL: NULL selects the Wannier FT convention, non-NULL the
                                 tight-binding convention                                    wann_interp.f90:192 */
    double a_latt[9];         /* Fortran a_latt(3,3) column-major; a_latt(i,:) = a_i*a0      input_paras.f90:165-168 */
} mepb_model;

/* == the module globals of src/input_paras.f90:49-68 that the hot path reads (already in atomic units) */
typedef struct {
    int32_t mp_grid[3];
    int32_t do_gauge_trafo, use_cart_velo;
    int32_t do_mep, do_kubo, do_ahc, do_opt, do_write_mep_bands;
    int32_t valence_bands, n_hw, n_eF;
    double  hw_min, hw_max, eF_min, eF_max;     /* Hartree  (driver converts from eV, input_paras.f90:171-178) */
    double  T_kelvin, eta /* Hartree */, kubo_tol;
    int32_t do_gyro;          /* jobs%do_gyro: gyrotropic tensors C, D, D(w) (src/gyro.f90:74-268)                      */
    int32_t posop_mode;       /* position-operator branch only.  0 = PRB 74 195118 eqs. 21-31 (default): A <- U^dagger A U,
                                 curvature and velocity corrections built from that rotated connection.  1 = the deterministic
                                 part of the reference's coded behaviour: curvature (wann_interp.f90:647-652) and velocity
                                 (:427-436) corrections use A AFTER A += iD.  The coded connection rotation itself
                                 ((U^dagger H_a U) U, :530-533) depends on the eigenvector phases LAPACK happens to return and
                                 the coded V diagonal is stale memory (:427-436): neither is reproducible by any implementation,
                                 both are replaced by the eq. 21 rotation and by Hbar(n,n) in either mode   */
} mepb_params;

/* Outputs of integrate_over_k_space (src/core.f90:419-561), laid out as the arrays save_npy writes
 * (src/file_io.f90:266-398): column-major, first index fastest.  Any pointer may be NULL (skipped). */
typedef struct {
    int64_t n_k;              /* k-points summed                                                              */
    int32_t eF_idx;           /* 0-based index chosen by fermi_selector                       core.f90:317-415 */
    int64_t n_degenerate;     /* (k, n, m) band pairs with |E_n - E_m| == 0 met by an unguarded formula (Q6)   */
    double *Ne;               /* (n_eF)      electrons per k-point for each trial E_F          core.f90:329     */
    double *mep_ic, *mep_lc, *mep_cs;          /* (3,3)                                                         */
    double *mep_bands;                         /* (3,3,valence_bands)  only if do_write_mep_bands               */
    double *kubo_mep_ic, *kubo_mep_lc, *kubo_mep_cs;   /* (3,3)                                                 */
    double *ahc;                               /* (3,3)                       ahc_tens.npy                      */
    double *ahc_velo, *ohc_velo;               /* (3,3,n_hw) complex          ahcVELO.npy, ohcVELO.npy          */
    double *opt_s, *opt_a;                     /* (3,3,n_hw) complex          opt_Ssymm.npy, opt_Asymm.npy (zeros without position operator, kubo.f90:225) */
    double *photo2;                            /* (3,3,3,n_hw,n_eF) complex   photoC_2nd.npy                    */
    double *gyro_c, *gyro_d;                   /* (3,3,n_eF) complex          gyro_C.npy, gyro_D.npy (all E_F, core.f90:523-530) */
    double *gyro_dw;                           /* (3,3,n_hw,n_eF) complex     gyro_Dw.npy                       */
} mepb_result;

typedef struct mepb_ctx mepb_ctx;

/* Upload the model, precompute the Fourier right-hand side, allocate accumulators on CUDA device `device`.
 * Replaces: read_tb_basis hand-over + kspace_allocator + allo_core_loc_arrays (core.f90:127-153). */
int mepb_create(const mepb_model *model, const mepb_params *params, int device, mepb_ctx **out);
void mepb_destroy(mepb_ctx *ctx);

/* Whole Monkhorst-Pack mesh on this context's device: accumulate + finalize.  Replaces core_worker's k-loop
 * and integrate_over_k_space for a 1-rank run (core.f90:165-292). */
int mepb_run_mesh(mepb_ctx *ctx, mepb_result *out);

/* The MPI-host / multi-GPU variant: each rank accumulates its share of the 0-based linear mesh index
 * lin = (kix-1) + N1*((kiy-1) + N2*(kiz-1)) (loop order of core.f90:165-167) and holds its un-normalised partial sums as one
 * packed FP64 buffer; the buffers are summed across ranks — the reference's mpi_reduce_sum calls, core.f90:467-526, fused
 * into one — either by the library itself (mepb_allreduce_partials: NCCL, below) or by the host (mepb_get_partials /
 * MPI_Allreduce / mepb_set_partials); every rank then finalizes (fermi_selector + normalize_k_int) redundantly. */
int     mepb_reset(mepb_ctx *ctx);
int     mepb_accumulate(mepb_ctx *ctx, int64_t lin_begin, int64_t lin_end);     /* contiguous slab                    */
/* k ownership by round-robin BLOCKS: rank takes the blocks j = rank, rank + nranks, ... of `block` consecutive k-points of
 * [lin_begin, lin_end).  The reference deals single k-points round-robin (mpi_ki_selector, mpi_comm.f90:80-99), which is
 * load-balanced by construction; block = mepb_balanced_block(ctx, nk, nranks) keeps that property with waves as large as
 * the context schedules them.  block <= 0: the whole range as one block. */
int     mepb_accumulate_strided(mepb_ctx *ctx, int64_t lin_begin, int64_t lin_end, int64_t block, int rank, int nranks);
int64_t mepb_chunk_size(const mepb_ctx *ctx);                       /* k-points per launch wave                 */
int64_t mepb_balanced_block(const mepb_ctx *ctx, int64_t nk, int nranks);
int64_t mepb_partials_size(const mepb_ctx *ctx);                    /* number of doubles                        */
int     mepb_get_partials(mepb_ctx *ctx, double *buf);              /* host buffer                              */
int     mepb_set_partials(mepb_ctx *ctx, const double *buf);        /* host buffer                              */
/* same, but the packed buffer lives in device memory of this context's device (for a caller-side all-reduce in place) */
int     mepb_partials_device_ptr(mepb_ctx *ctx, double **dptr);
int     mepb_finalize(mepb_ctx *ctx, mepb_result *out);

/* The exchange step inside the library (replaces src/mpi_comm.f90:118-367): one ncclAllReduce(sum, double) of the packed
 * buffer over NVLink, enqueued on the context's launch stream behind the packing kernel.  libnccl.so.2 is bound at run time.
 * One process per GPU: rank 0 calls mepb_comm_unique_id (128 bytes), the host broadcasts them (MPI_Bcast, a file, a
 * torch.distributed store), every rank calls mepb_comm_init_rank; then per run
 *     mepb_reset, mepb_accumulate_strided(..., rank, nranks), mepb_allreduce_partials, mepb_finalize. */
int     mepb_comm_unique_id(void *id128, int nbytes);
int     mepb_comm_init_rank(mepb_ctx *ctx, int nranks, int rank, const void *id128);
int     mepb_comm_size(const mepb_ctx *ctx);
int     mepb_comm_rank(const mepb_ctx *ctx);
int     mepb_allreduce_partials(mepb_ctx *ctx);
/* device time of the stages of the last run on this rank since mepb_reset (ms): k-loop, all-reduce, finalize (host) */
int     mepb_get_timing(const mepb_ctx *ctx, double *accumulate_ms, double *allreduce_ms, double *finalize_ms);

/* One process, several B200s (what `mpirun -np N ./mepInterp` is to the reference): a context per device, ncclCommInitAll,
 * round-robin blocks of the mesh per device on one host thread each, one all-reduce, finalize. */
typedef struct mepb_multi mepb_multi;
int       mepb_create_multi(const mepb_model *model, const mepb_params *params, int n_devices, const int *devices /* NULL: 0..n-1 */,
                            mepb_multi **out);
int       mepb_multi_run_mesh(mepb_multi *mm, mepb_result *out);
int       mepb_multi_update_model(mepb_multi *mm, const mepb_model *model);
int       mepb_multi_size(const mepb_multi *mm);
mepb_ctx *mepb_multi_ctx(mepb_multi *mm, int i);
void      mepb_multi_destroy(mepb_multi *mm);

/* Batched get_wann_interp (wann_interp.f90:56-108) on an explicit list of fractional k-points: band mode
 * (band_calc.f90:67-75) and the parity harness.  kpt (3,nk); e_k (nw,nk); V_ka (3,nw,nw,nk) complex or NULL.
 * V_ka is gauge dependent (eigenvector phases); compare only gauge-invariant combinations. */
int mepb_interp(mepb_ctx *ctx, int64_t nk, const double *kpt, double *e_k, double *V_ka);

/* get_wann_interp on a k list with the outputs of the position-operator branch (wann_interp.f90:56-108, 465-658): e_k (nw,nk);
 * V_ka (3,nw,nw,nk) velocities including the connection term; A_ka (3,nw,nw,nk) Berry connection in the Hamiltonian gauge;
 * Om_vec (3,nw,nk) the curvature pseudo-vector sum_ab eps_abc Om_ab(n,n) of every band (convert_tens_to_vect of the only part of
 * Om_kab the response functions read, matrix_math.f90:511-529).  Any pointer may be NULL.  Needs model.r (or do_gyro). */
int mepb_interp_pos(mepb_ctx *ctx, int64_t nk, const double *kpt, double *e_k, double *V_ka, double *A_ka, double *Om_vec);

/* Fourier sums only (FT_R_to_k, wann_interp.f90:151-281): H_k (nw,nw,nk), H_ka (3,nw,nw,nk) complex, W gauge.
 * Gauge independent, so it is compared element-wise against the oracle. */
int mepb_fourier(mepb_ctx *ctx, int64_t nk, const double *kpt, double *H_k, double *H_ka);

/* Same sums on mesh points: the k-points with 0-based linear mesh index [lin_begin, lin_begin + nk) of the context's
 * mp_grid, through the code path mepb_accumulate uses for them (general path: the mesh-separable transform).  Parity
 * harness for that path; MEPB_ERR_UNSUPPORTED on the small-num_wann path, whose kernels never materialise H(k). */
int mepb_fourier_mesh(mepb_ctx *ctx, int64_t lin_begin, int64_t nk, double *H_k, double *H_ka);

/* Band sums on caller-supplied eigenpairs: e_k (nw,nk) and the H-gauge velocities V_ka (3,nw,nw,nk) complex, as
 * get_wann_interp returns them (wann_interp.f90:56-108).  Runs the band-sum sections of the same kernels mepb_accumulate
 * uses (occupations, kubo_ahc_tens, mep_niu_*, kubo_mep_*, pair / photo items) and the frequency sweeps — the per-k
 * response functions of core.f90:187-253 without interpolation, eigensolve and rotation — and ADDS the un-normalised
 * sums to the context's accumulators exactly like mepb_accumulate (mepb_reset first; mepb_get_partials / mepb_finalize
 * follow).  Parity harness: fed the oracle's own (e, V) it removes the eigensolver from the comparison, so the response
 * formulas are checked at rounding level even where the reference's unguarded 1/dE^2, 1/dE^3 factors amplify
 * eigenvector rounding (kubo.f90:65, kubo_mep.f90:73,126).  Also the consumer of the position-operator kernels. */
int mepb_band_sums(mepb_ctx *ctx, int64_t nk, const double *e_k, const double *V_ka);

/* Re-upload H(R) from a host buffer into an existing context (same nw, nR, R set, FT convention): the per-call input
 * of a parameter sweep — the reference re-reads w90files/<seed>_hr.dat in a fresh process per parameter value
 * (scripts/mep_worker.py:133-150, scripts/probe_phi.py). */
int mepb_update_model(mepb_ctx *ctx, const mepb_model *model);
/* Same for the scalars of a frequency / smearing / Fermi-level / temperature sweep (scripts/probe_hw.py:167-208 starts a fresh
 * process per value): hw_min, hw_max, eF_min, eF_max, T_kelvin, eta, kubo_tol (and valence_bands unless do_mep) may change;
 * mp_grid, the job flags, n_hw and n_eF are fixed per context (MEPB_ERR_ARG otherwise). */
int mepb_update_params(mepb_ctx *ctx, const mepb_params *params);
/* What mepb_create decided for the model: herm_closed = 1 when H(-R) = H(R)^dagger holds bit for bit (compact pair items,
 * upper-triangle Fourier sums; a 6-decimal _hr.dat whose (n,m,R) and (m,n,-R) entries differ in the last digit gets 0 and the
 * general, slower formulation — results are the reference's either way), sparse = 1 when only non-zero H(R) entries are summed,
 * path = 0 / 1 / 2 (thread-per-k kernels for num_wann 4, 8 / warp-per-k kernel up to 16 / general path up to 64).  `why`
 * (may be NULL) receives the reason the closure test failed.  mepb_update_model keeps the flag fixed: a model that flips it is
 * refused with MEPB_ERR_UNSUPPORTED. */
int mepb_model_info(const mepb_ctx *ctx, int *herm_closed, int *sparse, int *path, char *why, int why_len);

/* CUDA-event stopwatch on the context's launch stream (brackets any sequence of mepb_* calls; ms on the device). */
int mepb_timer_start(mepb_ctx *ctx);
int mepb_timer_stop(mepb_ctx *ctx, double *ms);

/* Per-kernel-class device timing: when on, mepb_accumulate serialises its schedule onto one stream and brackets every
 * launch with CUDA events (slower; used for the roofline figures, never for throughput numbers). */
#define MEPB_PROF_INTERP       0   /* fused Fourier(dH) + rotation + static band sums + pair items     */
#define MEPB_PROF_SWEEP_AHC    1   /* velo_ahc / kubo_ohc frequency sweep                             */
#define MEPB_PROF_SWEEP_PHOTO  2   /* photo_2nd_cond frequency sweep                                  */
#define MEPB_PROF_REDUCE       3   /* block-partial reductions                                        */
#define MEPB_PROF_EIG          4   /* thread-per-k Fourier(H) + Jacobi eigensolver (num_wann 4, 8)     */
#define MEPB_PROF_FOURIER      5   /* dense Fourier contraction, general path (num_wann <= 64)         */
#define MEPB_PROF_EIGROT       6   /* CTA-per-k back-transform + rotation + band sums, general path   */
#define MEPB_PROF_TRIDIAG      7   /* Householder tridiagonalisation, general path                     */
#define MEPB_PROF_TQL          8   /* thread-per-k implicit QL on the tridiagonal (rotation log)       */
#define MEPB_PROF_ZROT         9   /* replay of the rotation log -> eigenvectors of the tridiagonal    */
#define MEPB_PROF_POSOP        10  /* position-operator branch: A / Om Fourier sums, gauge transformation, CS / opt / gyro sums */
#define MEPB_PROF_KINDS        11
int mepb_set_profiling(mepb_ctx *ctx, int on);
int mepb_get_profile(const mepb_ctx *ctx, int kind, double *ms_total, int64_t *launches);
/* (k, band pair) items the last packed run handed to the frequency sweep `kind` (MEPB_PROF_SWEEP_AHC / _PHOTO): the
 * unit count behind the sweeps' algorithmic flop figures (valid after mepb_get_partials / mepb_finalize). */
int64_t mepb_work_items(const mepb_ctx *ctx, int kind);

/* Accuracy harness for the device reciprocals that stand in for IEEE division in the sweeps and eigensolvers (a MUFU seed refined
 * by Newton / cubic steps: <= 1 ulp-level error on normal inputs; x = 0 or subnormal x give NaN / flushed results where IEEE
 * division gives +-inf — reached only by an exactly resonant denominator with eta = 0, where the reference's inf is as
 * meaningless).  Evaluates them on n caller-supplied values. */
int mepb_selftest_math(mepb_ctx *ctx, int n, const double *x, double *rcp_newton2, double *rcp_cubic, double *rsqrt_cubic);

/* timing of the last mepb_accumulate / mepb_run_mesh on the device (CUDA events), and kernel launches issued */
double  mepb_last_kernel_ms(const mepb_ctx *ctx);
int64_t mepb_kernel_launches(const mepb_ctx *ctx);

const char *mepb_strerror(int code);
const char *mepb_last_error(void);
const char *mepb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MEPB_H */
