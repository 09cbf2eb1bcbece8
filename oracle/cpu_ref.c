/*
 * cpu_ref.c — CPU ORACLE (plain C99) for the per-k-point Wannier-interpolation path of cMax347/MEPinterp.
 *
 * *** TEST INFRASTRUCTURE ONLY ***  Not linked into libmepb.so, not used by the CLI driver.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs load it, as the checker and as the
 * timed CPU baseline ("port": a C restatement of the reference algorithm, NOT the reference binary — the
 * reference is Fortran+MPI and no Fortran compiler exists in this image).
 *
 * Loop structure and arithmetic follow the Fortran line by line (paths relative to /root/reference):
 *   ft_R_to_k          src/wann_interp.f90:151-281     dense loops over (R, m, n); one cos+sin per (R,n,m) in the
 *                                                      tight-binding convention, per R in the Wannier convention
 *   zheevd             src/matrix_math.f90:192-226     LAPACK zheevd('V','U') with the reference's workspace sizes,
 *                                                      allocated per call like the reference; symbol scipy_zheevd_
 *                                                      from the OpenBLAS 0.3.30 inside the scipy wheel (dlopen)
 *   rotate_gauge       src/wann_interp.f90:502-550     two zgemm per direction (scipy_zgemm_), matrix_math.f90:636-655
 *   get_velo           src/wann_interp.f90:406-444     V = Hbar; with a position operator V(n,m) = Hbar(n,m) - i (E_m - E_n) A(n,m)
 *   ft_pos_R_to_k      src/wann_interp.f90:288-375     A_ka, Om_kab Fourier sums (W gauge), both FT conventions
 *   gauge_pos          src/wann_interp.f90:465-658     rotate_gauge (A, Om), get_gauge_covar_deriv, conn/curv_gaugeTrafo
 *   mep_niu_CS / kubo_mep_CS   src/mep_niu.f90:20-44, src/kubo_mep.f90:16-47
 *   kubo_opt (herm, aherm)     src/kubo.f90:204-405
 *   gyro C / D / Dw    src/gyro.f90:74-268, get_kubo_curv src/wann_interp.f90:113-147, fd_stat_deriv src/statistics.f90:161-197
 *   (position-operator branch: which coded defects are and are not reproduced is stated in oracle/mep_oracle.py's header)
 *   fd_stat            src/statistics.f90:65-90
 *   kubo_ahc_tens      src/kubo.f90:37-78
 *   velo_ahc_tens      src/kubo.f90:86-140
 *   kubo_ohc_tens      src/kubo.f90:145-198
 *   photo_2nd_cond     src/photo.f90:16-86
 *   mep_niu_LC/IC      src/mep_niu.f90:47-159
 *   kubo_mep_LC/IC     src/kubo_mep.f90:51-149
 *   k loop, ownership  src/core.f90:165-270, src/mpi_comm.f90:80-99 (round-robin ranks; here an O(1) test
 *                      instead of the reference's O(N_k/P) scan — favourable to the CPU side)
 *   finalize           src/core.f90:317-415 (fermi_selector), :419-603, src/k_space.f90:81-202,263-293
 *
 * Parity pin: see oracle/mep_oracle.py header (golden AHC tensors of the reference to mesh-convergence level,
 * the reference's KATs, and agreement of the two independent restatements to 1e-12).
 *
 * Packed partial-sum layout (doubles) — identical to libmepb's mepb_get_partials so raw sums can be compared:
 *   [0] n_k  [1] n_degenerate  [2 .. 2+n_eF) Ne
 *   if do_mep : mep_bands_ic (3,3,nval), mep_bands_lc (3,3,nval)
 *   if do_kubo: kubo_mep_ic (3,3,n_eF), kubo_mep_lc (3,3,n_eF)
 *   if do_ahc : kubo_ahc (3,3,n_eF), velo_ahc (3,3,n_hw,n_eF) c16, kubo_ohc (3,3,n_hw,n_eF) c16
 *   if do_opt : photo2 (3,3,3,n_hw,n_eF) c16
 *   with a position operator (model.r != NULL):
 *     if do_mep : mep_cs (nval)  — the scalar cs(i,i,n0) per band         if do_kubo: kubo_mep_cs (n_eF)
 *     if do_opt : opt_herm (3,3,n_hw,n_eF) c16, opt_aherm (3,3,n_hw,n_eF) c16
 *   if do_gyro: gyro_C (3,3,n_eF) c16, gyro_D (3,3,n_eF) c16, gyro_Dw (3,3,n_hw,n_eF) c16
 */
#include <complex.h>
#include <dlfcn.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/mepb.h"

typedef double complex zc;

/* constants.f90:6-23 */
static const double aUtoEv = 27.211385;
#define PI_DP (4.0 * atan(1.0))
#define KBOLTZ_EH_K (8.6173303 * 1e-5 / aUtoEv)
static const double min_temp = 1e-2; /* statistics.f90:32 */

typedef void (*zheevd_fn)(const char *, const char *, const int *, zc *, const int *, double *, zc *, const int *,
                          double *, const int *, int *, const int *, int *);
typedef void (*zgemm_fn)(const char *, const char *, const int *, const int *, const int *, const zc *, const zc *,
                         const int *, const zc *, const int *, const zc *, zc *, const int *);
typedef void (*setthreads_fn)(int);
static zheevd_fn p_zheevd = NULL;
static zgemm_fn p_zgemm = NULL;

int cpu_ref_init(const char *openblas_path) {
    if (p_zheevd && p_zgemm) return 0;
    void *h = dlopen(openblas_path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) { fprintf(stderr, "cpu_ref_init: dlopen(%s): %s\n", openblas_path, dlerror()); return -1; }
    p_zheevd = (zheevd_fn)dlsym(h, "scipy_zheevd_");
    if (!p_zheevd) p_zheevd = (zheevd_fn)dlsym(h, "zheevd_");
    p_zgemm = (zgemm_fn)dlsym(h, "scipy_zgemm_");
    if (!p_zgemm) p_zgemm = (zgemm_fn)dlsym(h, "zgemm_");
    setthreads_fn st = (setthreads_fn)dlsym(h, "scipy_openblas_set_num_threads");
    if (!st) st = (setthreads_fn)dlsym(h, "openblas_set_num_threads");
    if (st) st(1); /* single-threaded BLAS inside each worker */
    return (p_zheevd && p_zgemm) ? 0 : -2;
}

/* ------------------------------------------------------------------ layout of the packed partials */
typedef struct {
    int64_t total, ne, mep_ic, mep_lc, kmep_ic, kmep_lc, ahc, velo, ohc, photo, mep_cs, kmep_cs, opt_h, opt_ah, gyro_c, gyro_d, gyro_dw;
} layout_t;

static layout_t make_layout(const mepb_params *p, int posop) {
    layout_t L; memset(&L, 0, sizeof L);
    int64_t o = 2; L.ne = o; o += p->n_eF;
    L.mep_ic = L.mep_lc = L.kmep_ic = L.kmep_lc = L.ahc = L.velo = L.ohc = L.photo = -1;
    if (p->do_mep)  { L.mep_ic = o; o += 9 * p->valence_bands; L.mep_lc = o; o += 9 * p->valence_bands; }
    if (p->do_kubo) { L.kmep_ic = o; o += 9 * p->n_eF; L.kmep_lc = o; o += 9 * p->n_eF; }
    if (p->do_ahc)  { L.ahc = o; o += 9 * p->n_eF; L.velo = o; o += 2 * 9 * (int64_t)p->n_hw * p->n_eF;
                      L.ohc = o; o += 2 * 9 * (int64_t)p->n_hw * p->n_eF; }
    if (p->do_opt)  { L.photo = o; o += 2 * 27 * (int64_t)p->n_hw * p->n_eF; }
    L.mep_cs = L.kmep_cs = L.opt_h = L.opt_ah = L.gyro_c = L.gyro_d = L.gyro_dw = -1;
    if (posop && p->do_mep)  { L.mep_cs = o; o += p->valence_bands; }
    if (posop && p->do_kubo) { L.kmep_cs = o; o += p->n_eF; }
    if (posop && p->do_opt)  { L.opt_h = o; o += 2 * 9 * (int64_t)p->n_hw * p->n_eF; L.opt_ah = o; o += 2 * 9 * (int64_t)p->n_hw * p->n_eF; }
    if (p->do_gyro) { L.gyro_c = o; o += 2 * 9 * (int64_t)p->n_eF; L.gyro_d = o; o += 2 * 9 * (int64_t)p->n_eF;
                      L.gyro_dw = o; o += 2 * 9 * (int64_t)p->n_hw * p->n_eF; }
    L.total = o;
    return L;
}
int64_t cpu_ref_partials_size(const mepb_model *mo, const mepb_params *p) { return make_layout(p, mo->r != NULL).total; }

static void get_linspace(double lo, double hi, int n, double *out) { /* matrix_math.f90:74-100 */
    if (n > 1) { double d = (hi - lo) / (double)(n - 1); for (int i = 0; i < n; ++i) out[i] = lo + (double)i * d; }
    else out[0] = lo;
}
static int levi(int i, int j, int k) { /* matrix_math.f90:103-122, 0-based */
    if ((i == 0 && j == 1 && k == 2) || (i == 1 && j == 2 && k == 0) || (i == 2 && j == 0 && k == 1)) return 1;
    if ((i == 2 && j == 1 && k == 0) || (i == 0 && j == 2 && k == 1) || (i == 1 && j == 0 && k == 2)) return -1;
    return 0;
}
static void crossP(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
/* k_space.f90:263-293: columns of a_latt are taken as a_i; recip(:,i) = b_i (column-major) */
static double set_recip_latt(const double *a_latt, double *recip) {
    const double *a1 = a_latt, *a2 = a_latt + 3, *a3 = a_latt + 6;
    double c[3]; crossP(a1, a2, c);
    double vol = c[0] * a3[0] + c[1] * a3[1] + c[2] * a3[2];
    double two_pi = 2.0 * PI_DP;
    crossP(a2, a3, c); for (int i = 0; i < 3; ++i) recip[i] = two_pi * c[i] / vol;
    crossP(a3, a1, c); for (int i = 0; i < 3; ++i) recip[3 + i] = two_pi * c[i] / vol;
    crossP(a1, a2, c); for (int i = 0; i < 3; ++i) recip[6 + i] = two_pi * c[i] / vol;
    return vol;
}
static void matvec3(const double *A /*col-major*/, const double *x, double *y) { /* matmul(A, x) */
    for (int i = 0; i < 3; ++i) y[i] = A[i] * x[0] + A[i + 3] * x[1] + A[i + 6] * x[2];
}

/* index helpers, Fortran order */
#define H3(n, m, sc) H[(n) + nw * ((m) + (int64_t)nw * (sc))]
#define VA(a, n, m) V[(a) + 3 * ((n) + nw * (m))]

/* wann_interp.f90:151-281 */
static void ft_R_to_k(const mepb_model *mo, int use_cart_velo, const double *recip, const double *kpt, zc *H_k, zc *H_ka) {
    const int nw = mo->nw, nR = mo->nR;
    const zc *H = (const zc *)mo->H;
    const double two_pi = 2.0 * PI_DP;
    memset(H_k, 0, sizeof(zc) * nw * nw);
    memset(H_ka, 0, sizeof(zc) * 3 * nw * nw);
    double kcart[3]; matvec3(recip, kpt, kcart); /* get_cart_kpt k_space.f90:387 */
    if (mo->wf_centers) {
        const double *tau = mo->wf_centers;
        for (int sc = 0; sc < nR; ++sc)
            for (int m = 0; m < nw; ++m)
                for (int n = 0; n < nw; ++n) {
                    double dR[3], dRc[3], ang;
                    for (int i = 0; i < 3; ++i) dR[i] = mo->R[i + 3 * sc] + (tau[i + 3 * m] - tau[i + 3 * n]); /* :203,:213 */
                    if (use_cart_velo) { matvec3(mo->a_latt, dR, dRc); ang = kcart[0] * dRc[0] + kcart[1] * dRc[1] + kcart[2] * dRc[2]; }
                    else { ang = two_pi * (kpt[0] * dR[0] + kpt[1] * dR[1] + kpt[2] * dR[2]); for (int i = 0; i < 3; ++i) dRc[i] = two_pi * dR[i]; }
                    zc ph = cos(ang) + I * sin(ang);
                    zc h = H3(n, m, sc);
                    H_k[n + nw * m] += ph * h;
                    for (int a = 0; a < 3; ++a) H_ka[a + 3 * (n + nw * m)] += ph * (I * dRc[a]) * h;
                }
    } else {
        for (int sc = 0; sc < nR; ++sc) {
            const double *Rv = mo->R + 3 * sc;
            double ang = two_pi * (kpt[0] * Rv[0] + kpt[1] * Rv[1] + kpt[2] * Rv[2]);
            zc ph = cos(ang) + I * sin(ang);
            double dRc[3];
            if (use_cart_velo) matvec3(mo->a_latt, Rv, dRc); else for (int i = 0; i < 3; ++i) dRc[i] = two_pi * Rv[i];
            for (int m = 0; m < nw; ++m)
                for (int n = 0; n < nw; ++n) {
                    zc h = H3(n, m, sc);
                    H_k[n + nw * m] += ph * h;
                    for (int a = 0; a < 3; ++a) H_ka[a + 3 * (n + nw * m)] += ph * I * dRc[a] * h;
                }
        }
    }
}

/* matrix_math.f90:192-226 — workspace sizes and per-call allocation as in the reference */
static int zheevd_wrapper(int n, zc *A, double *w) {
    int lwork = n * n + 2 * n, lrwork = 2 * n * n + 5 * n + 1, liwork = 5 * n + 3, info = 0;
    zc *work = (zc *)malloc(sizeof(zc) * lwork);
    double *rwork = (double *)malloc(sizeof(double) * lrwork);
    int *iwork = (int *)malloc(sizeof(int) * liwork);
    p_zheevd("V", "U", &n, A, &n, w, work, &lwork, rwork, &lrwork, iwork, &liwork, &info);
    free(work); free(rwork); free(iwork);
    return info;
}
static void blas_matmul(int n, const zc *A, const zc *B, zc *C) { /* matrix_math.f90:636-655 */
    zc one = 1.0, zero = 0.0;
    p_zgemm("N", "N", &n, &n, &n, &one, A, &n, B, &n, &zero, C, &n);
}

/* wann_interp.f90:56-108: e_k(nw), V(3,nw,nw); U_k returned in U (nw,nw) */
static int get_wann_interp(const mepb_model *mo, const mepb_params *p, const double *recip, const double *kpt,
                           double *e_k, zc *V, zc *U, zc *H_ka, zc *tmp, zc *M_in, zc *U_dag) {
    const int nw = mo->nw;
    ft_R_to_k(mo, p->use_cart_velo, recip, kpt, U, H_ka);
    int info = zheevd_wrapper(nw, U, e_k);
    if (info != 0) return info;
    if (p->do_gauge_trafo) { /* rotate_gauge :502-550 */
        for (int n = 0; n < nw; ++n) for (int m = 0; m < nw; ++m) U_dag[n + nw * m] = conj(U[m + nw * n]);
        for (int a = 0; a < 3; ++a) {
            for (int i = 0; i < nw * nw; ++i) M_in[i] = H_ka[a + 3 * i];
            blas_matmul(nw, M_in, U, tmp);
            blas_matmul(nw, U_dag, tmp, M_in);
            for (int i = 0; i < nw * nw; ++i) H_ka[a + 3 * i] = M_in[i];
        }
    }
    memcpy(V, H_ka, sizeof(zc) * 3 * nw * nw); /* get_velo :440 */
    return 0;
}

/* wann_interp.f90:288-375: A (3,nw,nw), Om (3,3,nw,nw) in the W gauge */
#define AA(a, n, m) A[(a) + 3 * ((n) + nw * (m))]
#define OM(a, b, n, m) Om[(a) + 3 * ((b) + 3 * ((n) + nw * (m)))]
static void ft_pos_R_to_k(const mepb_model *mo, int use_cart_velo, const double *kpt, zc *A, zc *Om) {
    const int nw = mo->nw, nR = mo->nR;
    const zc *r = (const zc *)mo->r;
    const double two_pi = 2.0 * PI_DP;
    memset(A, 0, sizeof(zc) * 3 * nw * nw);
    memset(Om, 0, sizeof(zc) * 9 * nw * nw);
    for (int sc = 0; sc < nR; ++sc) {
        const double *Rv = mo->R + 3 * sc;
        for (int m = 0; m < nw; ++m)
            for (int n = 0; n < nw; ++n) {
                double dR[3], dRc[3];
                for (int i = 0; i < 3; ++i) dR[i] = Rv[i] + (mo->wf_centers ? mo->wf_centers[i + 3 * m] - mo->wf_centers[i + 3 * n] : 0.0);
                const double ang = two_pi * (kpt[0] * dR[0] + kpt[1] * dR[1] + kpt[2] * dR[2]);       /* :306, :344 */
                const zc ph = cos(ang) + I * sin(ang);
                if (use_cart_velo) matvec3(mo->a_latt, dR, dRc); else for (int i = 0; i < 3; ++i) dRc[i] = two_pi * dR[i];
                const zc *rr = r + 3 * ((int64_t)n + nw * ((int64_t)m + (int64_t)nw * sc));             /* r(a,n,m,sc) */
                for (int a = 0; a < 3; ++a) {
                    AA(a, n, m) += ph * rr[a];
                    for (int b = 0; b < 3; ++b) {
                        OM(a, b, n, m) += ph * I * dRc[a] * rr[b];
                        OM(a, b, n, m) -= ph * I * dRc[b] * rr[a];
                    }
                }
            }
    }
}

/* X <- U^dagger X U for one nw x nw matrix held with element stride `st` (slices of the (3,..) / (3,3,..) arrays) */
static void rotate_slice(int nw, const zc *U, const zc *U_dag, zc *X, int st, zc *t1, zc *t2) {
    for (int i = 0; i < nw * nw; ++i) t1[i] = X[(int64_t)st * i];
    blas_matmul(nw, t1, U, t2);
    blas_matmul(nw, U_dag, t2, t1);
    for (int i = 0; i < nw * nw; ++i) X[(int64_t)st * i] = t1[i];
}
static void comm_acc(int nw, zc coef, const zc *X, const zc *Y, zc *out, int st, zc *t1, zc *t2) { /* out += coef [X, Y] */
    blas_matmul(nw, X, Y, t1);
    blas_matmul(nw, Y, X, t2);
    for (int i = 0; i < nw * nw; ++i) out[(int64_t)st * i] += coef * (t1[i] - t2[i]);
}

/* wann_interp.f90:56-108 with a position operator.  e_k(nw), V(3,nw,nw), A(3,nw,nw), Om(3,3,nw,nw) in the H gauge */
static int get_wann_interp_pos(const mepb_model *mo, const mepb_params *p, const double *recip, const double *kpt, double *e_k,
                               zc *V, zc *A, zc *Om, zc *U, zc *H_ka, zc *wk /* >= 12 nw^2 */) {
    const int nw = mo->nw, nw2 = nw * nw;
    zc *U_dag = wk, *t1 = wk + nw2, *t2 = wk + 2 * nw2, *D = wk + 3 * nw2 /* 3 nw2, [a][n + nw m] */, *Ax = wk + 6 * nw2 /* 3 nw2 */,
       *Ah = wk + 9 * nw2 /* 3 nw2 */;
    ft_R_to_k(mo, p->use_cart_velo, recip, kpt, U, H_ka);
    ft_pos_R_to_k(mo, p->use_cart_velo, kpt, A, Om);
    int info = zheevd_wrapper(nw, U, e_k);
    if (info != 0) return info;
    const zc *Av = A;                                             /* connection that enters get_velo */
    if (p->do_gauge_trafo) {
        for (int n = 0; n < nw; ++n) for (int m = 0; m < nw; ++m) U_dag[n + nw * m] = conj(U[m + nw * n]);
        for (int a = 0; a < 3; ++a) {
            rotate_slice(nw, U, U_dag, H_ka + a, 3, t1, t2);                       /* :517-528 */
            rotate_slice(nw, U, U_dag, A + a, 3, t1, t2);                          /* eq. (21); the coded line :530-533 is Q1 */
            for (int b = 0; b < 3; ++b) rotate_slice(nw, U, U_dag, Om + a + 3 * b, 9, t1, t2);   /* :536-542 */
        }
        for (int a = 0; a < 3; ++a)                                                 /* get_gauge_covar_deriv :565-617 */
            for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {
                const double dE = e_k[m] - e_k[n];
                D[a * nw2 + n + nw * m] = (n != m && fabs(dE) > p->kubo_tol) ? H_ka[a + 3 * (n + nw * m)] / dE : 0.0;
            }
        for (int a = 0; a < 3; ++a) for (int i = 0; i < nw2; ++i) {
            Ax[a * nw2 + i] = A[a + 3 * i];                                         /* rotated W-gauge connection */
            A[a + 3 * i] += I * D[a * nw2 + i];                                     /* conn_gaugeTrafo :620-633 */
            Ah[a * nw2 + i] = A[a + 3 * i];
        }
        const zc *Ac = (p->posop_mode == 1) ? Ah : Ax;                              /* Q4 coded / eq. (27) */
        for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) {                   /* curv_gaugeTrafo :636-658 */
            comm_acc(nw, -1.0, D + a * nw2, Ac + b * nw2, Om + a + 3 * b, 9, t1, t2);
            comm_acc(nw, 1.0, D + b * nw2, Ac + a * nw2, Om + a + 3 * b, 9, t1, t2);
            comm_acc(nw, -I, D + a * nw2, D + b * nw2, Om + a + 3 * b, 9, t1, t2);
        }
        Av = (p->posop_mode == 1) ? Ah : Ax;                                        /* Q15 coded / eq. (31) */
        for (int a = 0; a < 3; ++a) for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {   /* get_velo :424-437; diagonal = Hbar (Q2) */
            const zc h = H_ka[a + 3 * (n + nw * m)];
            V[a + 3 * (n + nw * m)] = (n == m) ? h : h - I * (e_k[m] - e_k[n]) * Av[a * nw2 + n + nw * m];
        }
    } else {
        for (int a = 0; a < 3; ++a) for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {
            const zc h = H_ka[a + 3 * (n + nw * m)];
            V[a + 3 * (n + nw * m)] = (n == m) ? h : h - I * (e_k[m] - e_k[n]) * Av[a + 3 * (n + nw * m)];
        }
    }
    return 0;
}

/* statistics.f90:41-61 */
static double fd_stat(double e_band, double e_fermi, double T_kelvin) {
    if (T_kelvin > min_temp) { double T_smear = KBOLTZ_EH_K * T_kelvin; return 1.0 / (1.0 + exp((e_band - e_fermi) / T_smear)); }
    return (e_band < e_fermi) ? 1.0 : 0.0;
}

/* ------------------------------------------------------------------ band sums, all accumulate (+=) into out */
#define FD(ef, n) fd[(ef) + n_ef * (n)]

static void kubo_ahc_tens(int nw, int n_ef, const double *en, const zc *V, const double *fd, double *out /*3,3,n_ef*/) {
    for (int n = 0; n < nw; ++n) { /* kubo.f90:58-73 */
        double curv[9]; memset(curv, 0, sizeof curv);
        for (int np = 0; np < nw; ++np) {
            if (np == n) continue;
            for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i)
                curv[i + 3 * j] += -2.0 * cimag(VA(i, n, np) * VA(j, np, n)) / ((en[np] - en[n]) * (en[np] - en[n]));
        }
        for (int ef = 0; ef < n_ef; ++ef) for (int q = 0; q < 9; ++q) out[q + 9 * ef] -= curv[q] * FD(ef, n);
    }
}
static void velo_ahc_tens(int nw, int n_ef, int n_hw, const double *en, const zc *V, const double *hw, const double *fd,
                          zc i_eta, zc *out /*3,3,n_hw,n_ef*/, zc *vv /*9*n_hw scratch*/) {
    for (int n = 0; n < nw; ++n) for (int np = 0; np < nw; ++np) { /* kubo.f90:111-135 */
        if (np == n) continue;
        double dE_sqr = (en[n] - en[np]) * (en[n] - en[np]);
        double v[9];
        for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) v[i + 3 * j] = cimag(VA(i, n, np) * VA(j, np, n));
        for (int h = 0; h < n_hw; ++h) {
            zc den = dE_sqr - (hw[h] + i_eta) * (hw[h] + i_eta);
            for (int q = 0; q < 9; ++q) vv[q + 9 * h] = v[q] / den;
        }
        for (int ef = 0; ef < n_ef; ++ef) {
            double d = FD(ef, n) - FD(ef, np);
            for (int q = 0; q < 9 * n_hw; ++q) out[q + (int64_t)9 * n_hw * ef] += d * vv[q];
        }
    }
}
static void kubo_ohc_tens(int nw, int n_ef, int n_hw, const double *en, const zc *V, const double *hw, const double *fd,
                          zc i_eta, zc *out, zc *vv) {
    for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) { /* kubo.f90:172-194 */
        if (m == n) continue;
        double dE = en[m] - en[n];
        zc v[9];
        for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) v[i + 3 * j] = VA(i, n, m) * VA(j, m, n);
        for (int h = 0; h < n_hw; ++h) {
            zc den = dE * (dE - hw[h] - i_eta);
            for (int q = 0; q < 9; ++q) vv[q + 9 * h] = v[q] / den;
        }
        for (int ef = 0; ef < n_ef; ++ef) {
            double d = FD(ef, m) - FD(ef, n);
            for (int q = 0; q < 9 * n_hw; ++q) out[q + (int64_t)9 * n_hw * ef] += I * vv[q] * d;
        }
    }
}
static void photo_2nd_cond(int nw, int n_ef, int n_hw, const double *en, const zc *V, const double *hw, const double *fd,
                           zc i_eta, zc *out /*3,3,3,n_hw,n_ef*/, zc *tmp /*27*n_hw*/) {
    for (int n = 0; n < nw; ++n) for (int m = 0; m < nw; ++m) { /* photo.f90:44-81 */
        zc dE_nm = (en[n] - en[m]) - i_eta;
        for (int l = 0; l < nw; ++l) {
            if (l == n) continue;
            zc dE_nl = (en[n] - en[l]) - i_eta;
            zc vvv[27];
            for (int c = 0; c < 3; ++c) for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a)
                vvv[a + 3 * (b + 3 * c)] = VA(a, n, l) * (VA(b, l, m) * VA(c, m, n));
            for (int h = 0; h < n_hw; ++h) {
                for (int q = 0; q < 27; ++q) tmp[q + 27 * h] = 0.0;
                for (int om = -1; om <= 1; om += 2) {
                    zc dE_nl_hw = dE_nl + (double)om * hw[h] - i_eta;
                    zc den = dE_nm * dE_nl_hw;
                    for (int q = 0; q < 27; ++q) tmp[q + 27 * h] += vvv[q] / den;
                }
                double w2 = hw[h] * hw[h];
                for (int q = 0; q < 27; ++q) tmp[q + 27 * h] /= w2;
            }
            for (int ef = 0; ef < n_ef; ++ef) {
                double d = FD(ef, l) - FD(ef, n);
                for (int q = 0; q < 27 * n_hw; ++q) out[q + (int64_t)27 * n_hw * ef] += tmp[q] * d;
            }
        }
    }
}
/* mep_niu.f90:47-99 */
static void mep_niu_LC(int nw, int nval, double kubo_tol, const double *en, const zc *V, double *F3 /*3,3,nval*/) {
    for (int n0 = 0; n0 < nval; ++n0) for (int n = 0; n < nw; ++n) {
        if (n == n0) continue;
        double den = pow(en[n0] - en[n], 3);
        if (!(fabs(den) > kubo_tol)) continue;
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) {
            double pre = (double)levi(j, k, l);
            if (fabs(pre) > 1e-1)
                for (int i = 0; i < 3; ++i) F3[i + 3 * j + 9 * n0] += pre * creal(VA(i, n0, n) * VA(k, n, n0) * VA(l, n0, n0)) / den;
        }
    }
}
/* mep_niu.f90:102-159 */
static void mep_niu_IC(int nw, int nval, double kubo_tol, const double *en, const zc *V, double *F2) {
    for (int n0 = 0; n0 < nval; ++n0) for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {
        if (n == n0 || m == n0) continue;
        double den = (en[n0] - en[n]) * (en[n0] - en[n]) * (en[n0] - en[m]);
        if (!(fabs(den) > kubo_tol)) continue;
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) {
            double pre = -(double)levi(j, k, l);
            if (fabs(pre) > 1e-1)
                for (int i = 0; i < 3; ++i) F2[i + 3 * j + 9 * n0] += pre * creal(VA(i, n0, n) * VA(k, n, m) * VA(l, m, n0)) / den;
        }
    }
}
/* kubo_mep.f90:51-97 */
static void kubo_mep_LC(int nw, int n_ef, const double *en, const zc *V, const double *fd, double *F3 /*3,3,n_ef*/) {
    for (int n0 = 0; n0 < nw; ++n0) for (int n = 0; n < nw; ++n) {
        if (n == n0) continue;
        double den = pow(en[n0] - en[n], 3);
        double v[9]; memset(v, 0, sizeof v);
        for (int j = 0; j < 3; ++j) for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l)
            for (int i = 0; i < 3; ++i) v[i + 3 * j] += (double)levi(j, k, l) * creal(VA(i, n0, n) * VA(k, n, n0) * VA(l, n0, n0));
        for (int ef = 0; ef < n_ef; ++ef) for (int q = 0; q < 9; ++q) F3[q + 9 * ef] += v[q] * FD(ef, n0) / den;
    }
}
/* kubo_mep.f90:100-149 */
static void kubo_mep_IC(int nw, int n_ef, const double *en, const zc *V, const double *fd, double *F2) {
    for (int n0 = 0; n0 < nw; ++n0) for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {
        if (n == n0 || m == n0) continue;
        double den = (en[n0] - en[m]) * ((en[n0] - en[n]) * (en[n0] - en[n]));
        double v[9]; memset(v, 0, sizeof v);
        for (int j = 0; j < 3; ++j) for (int l = 0; l < 3; ++l) for (int k = 0; k < 3; ++k)
            for (int i = 0; i < 3; ++i) v[i + 3 * j] -= (double)levi(j, k, l) * creal(VA(i, n0, n) * VA(k, n, m) * VA(l, m, n0));
        for (int ef = 0; ef < n_ef; ++ef) for (int q = 0; q < 9; ++q) F2[q + 9 * ef] += v[q] * FD(ef, n0) / den;
    }
}

/* convert_tens_to_vect, matrix_math.f90:511-529, on Om(:,:,n,n) */
static void om_vect_nn(int nw, const zc *Om, int n, zc *v) {
    for (int c = 0; c < 3; ++c) { v[c] = 0.0; for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) v[c] += (double)levi(a, b, c) * OM(a, b, n, n); }
}
/* mep_niu.f90:20-44 / kubo_mep.f90:16-47: 0.5 Re dot_product(A(:,n0,n0), om_vect)  (dot_product conjugates its first argument) */
static double cs_scalar(int nw, const zc *A, const zc *Om, int n0) {
    zc v[3]; om_vect_nn(nw, Om, n0, v);
    zc d = 0.0;
    for (int c = 0; c < 3; ++c) d += conj(AA(c, n0, n0)) * v[c];
    return 0.5 * creal(d);
}
/* kubo.f90:272-405: herm and aherm sums, (3,3,n_hw,n_ef) each */
static void kubo_opt_sums(int nw, int n_ef, int n_hw, const double *en, const zc *A, const double *hw, const double *fd, zc i_eta,
                          zc *herm, zc *aherm) {
    for (int m = 0; m < nw; ++m) for (int n = 0; n < nw; ++n) {
        if (n == m) continue;
        const double dE = en[m] - en[n];
        zc a_nm_mn[9];
        for (int b = 0; b < 3; ++b) for (int i = 0; i < 3; ++i) a_nm_mn[i + 3 * b] = dE * (AA(i, n, m) * AA(b, m, n));
        const zc dE_smr = PI_DP * (dE - i_eta);                                       /* delta_broad :248-269 */
        for (int h = 0; h < n_hw; ++h) {
            const double d_broad = cimag(1.0 / (dE_smr - PI_DP * hw[h]));
            const double den = creal(dE - hw[h] - i_eta);                            /* :389 real(...) */
            for (int ef = 0; ef < n_ef; ++ef) {
                const double df = FD(ef, m) - FD(ef, n);
                for (int q = 0; q < 9; ++q) {
                    herm[q + 9 * (h + (int64_t)n_hw * ef)] -= PI_DP * (a_nm_mn[q] * d_broad) * df;
                    aherm[q + 9 * (h + (int64_t)n_hw * ef)] += I * (a_nm_mn[q] / den) * df;
                }
            }
        }
    }
}
/* gyro.f90 with fd_deriv = 1 / (2 + 2 cosh(f)) AS CODED (d1_fd_stat_deriv evaluates cosh on the occupations, statistics.f90:186) */
static void gyro_sums(int nw, int n_ef, int n_hw, double kubo_tol, const double *en, const zc *V, const zc *A, const zc *Om,
                      const double *hw, const double *fd, zc *C, zc *Dt, zc *Dw) {
    for (int n = 0; n < nw; ++n) {
        zc om[3];
        if (Om) om_vect_nn(nw, Om, n, om);
        else {                                                                        /* get_kubo_curv, wann_interp.f90:113-147 */
            double curv[9]; memset(curv, 0, sizeof curv);
            for (int np = 0; np < nw; ++np) {
                if (np == n) continue;
                const double dEE = (en[np] - en[n]) * (en[np] - en[n]);
                if (fabs(dEE) < kubo_tol) continue;
                for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) curv[i + 3 * j] += cimag(VA(i, n, np) * VA(j, np, n)) / dEE;
            }
            for (int c = 0; c < 3; ++c) { om[c] = 0.0; for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) om[c] += (double)levi(a, b, c) * (-2.0 * curv[a + 3 * b]); }
        }
        for (int ef = 0; ef < n_ef; ++ef) {
            const double fdd = 1.0 / (2.0 + 2.0 * cosh(FD(ef, n)));
            for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) {
                C[a + 3 * b + 9 * ef] += VA(a, n, n) * VA(b, n, n) * fdd;            /* get_gyro_C :200-232 */
                Dt[a + 3 * b + 9 * ef] -= VA(a, n, n) * om[b] * fdd;                 /* get_gyro_D :74-150 */
            }
        }
        if (A) {                                                                      /* get_gyro_Dw :153-197, get_curvTilde_hw :235-268 */
            for (int h = 0; h < n_hw; ++h) {
                double omt[3] = {0.0, 0.0, 0.0};
                for (int m = 0; m < nw; ++m) {
                    if (m == n) continue;
                    const double ww = (en[m] - en[n]) * (en[m] - en[n]);
                    const zc a1[3] = {AA(0, n, m), AA(1, n, m), AA(2, n, m)}, a2[3] = {AA(0, m, n), AA(1, m, n), AA(2, m, n)};
                    const double cr[3] = {cimag(a1[1] * a2[2] - a1[2] * a2[1]), cimag(a1[2] * a2[0] - a1[0] * a2[2]), cimag(a1[0] * a2[1] - a1[1] * a2[0])};
                    for (int b = 0; b < 3; ++b) omt[b] -= cr[b] * ww / (ww - hw[h]);
                }
                for (int ef = 0; ef < n_ef; ++ef) {
                    const double fdd = 1.0 / (2.0 + 2.0 * cosh(FD(ef, n)));
                    for (int b = 0; b < 3; ++b) for (int a = 0; a < 3; ++a) Dw[a + 3 * b + 9 * (h + (int64_t)n_hw * ef)] -= VA(a, n, n) * omt[b] * fdd;
                }
            }
        }
    }
}

/* k_space.f90:351-375 with the loop order of core.f90:165-167 for a 0-based linear index */
static void mesh_kpt(const int32_t *g, int64_t lin, double *kpt) {
    int64_t kix = lin % g[0] + 1, kiy = (lin / g[0]) % g[1] + 1, kiz = lin / ((int64_t)g[0] * g[1]) + 1;
    kpt[0] = (2.0 * (double)kix - (double)g[0] - 1.0) / (2.0 * (double)g[0]);
    kpt[1] = (2.0 * (double)kiy - (double)g[1] - 1.0) / (2.0 * (double)g[1]);
    kpt[2] = (2.0 * (double)kiz - (double)g[2] - 1.0) / (2.0 * (double)g[2]);
}

/* The per-k response functions of core.f90:187-253 on one set of eigenpairs (en, V): occupations, MEP, Kubo-MEP, AHC,
 * optical sums, all += into the packed partials. */
static void band_sums_k(const mepb_model *mo, const mepb_params *p, const layout_t *Lp, const double *ef_lst, const double *hw_lst,
                        const double *en, const zc *V, const zc *A, const zc *Om, double *fd, zc *scr, double *partials) {
    const int nw = mo->nw, n_ef = p->n_eF, n_hw = p->n_hw;
    const layout_t L = *Lp;
    zc i_eta = I * p->eta;
    for (int n = 0; n < nw; ++n) for (int ef = 0; ef < n_ef; ++ef) { /* core.f90:187-193 */
        double f = fd_stat(en[n], ef_lst[ef], p->T_kelvin);
        FD(ef, n) = f; partials[L.ne + ef] += f;
    }
    for (int n = 0; n < nw; ++n) for (int m = n + 1; m < nw; ++m) if (en[n] == en[m]) partials[1] += 1.0;
    if (p->do_mep) {
        mep_niu_IC(nw, p->valence_bands, p->kubo_tol, en, V, partials + L.mep_ic);
        mep_niu_LC(nw, p->valence_bands, p->kubo_tol, en, V, partials + L.mep_lc);
    }
    if (p->do_kubo) {
        kubo_mep_IC(nw, n_ef, en, V, fd, partials + L.kmep_ic);
        kubo_mep_LC(nw, n_ef, en, V, fd, partials + L.kmep_lc);
    }
    if (p->do_ahc) {
        kubo_ahc_tens(nw, n_ef, en, V, fd, partials + L.ahc);
        velo_ahc_tens(nw, n_ef, n_hw, en, V, hw_lst, fd, i_eta, (zc *)(partials + L.velo), scr);
        kubo_ohc_tens(nw, n_ef, n_hw, en, V, hw_lst, fd, i_eta, (zc *)(partials + L.ohc), scr);
    }
    if (p->do_opt) photo_2nd_cond(nw, n_ef, n_hw, en, V, hw_lst, fd, i_eta, (zc *)(partials + L.photo), scr);
    if (A && Om) {                                                   /* position operator: Chern-Simons terms, kubo_opt_tens */
        if (p->do_mep) for (int n0 = 0; n0 < p->valence_bands; ++n0) partials[L.mep_cs + n0] += cs_scalar(nw, A, Om, n0);
        if (p->do_kubo) for (int n0 = 0; n0 < nw; ++n0) { const double c = cs_scalar(nw, A, Om, n0); for (int ef = 0; ef < n_ef; ++ef) partials[L.kmep_cs + ef] += c * FD(ef, n0); }
        if (p->do_opt) kubo_opt_sums(nw, n_ef, n_hw, en, A, hw_lst, fd, i_eta, (zc *)(partials + L.opt_h), (zc *)(partials + L.opt_ah));
    }
    if (p->do_gyro) gyro_sums(nw, n_ef, n_hw, p->kubo_tol, en, V, A, Om, hw_lst, fd, (zc *)(partials + L.gyro_c), (zc *)(partials + L.gyro_d),
                              (zc *)(partials + L.gyro_dw));
}

/* One "rank": k-points lin_begin + rank, + rank + nranks, ... < lin_end  (mpi_comm.f90:80-99 round-robin).  partials += */
int cpu_ref_accumulate(const mepb_model *mo, const mepb_params *p, int64_t lin_begin, int64_t lin_end, int rank, int nranks,
                       double *partials) {
    if (!p_zheevd) return -1;
    const int nw = mo->nw, n_ef = p->n_eF, n_hw = p->n_hw;
    zc *Ak = NULL, *Omk = NULL, *wk = NULL;
    if (mo->r) { Ak = (zc *)malloc(sizeof(zc) * 3 * nw * nw); Omk = (zc *)malloc(sizeof(zc) * 9 * nw * nw); wk = (zc *)malloc(sizeof(zc) * 12 * nw * nw); }
    layout_t L = make_layout(p, mo->r != NULL);
    double recip[9]; set_recip_latt(mo->a_latt, recip);
    double *ef_lst = (double *)malloc(sizeof(double) * n_ef), *hw_lst = (double *)malloc(sizeof(double) * n_hw);
    get_linspace(p->eF_min, p->eF_max, n_ef, ef_lst); /* core.f90:124-125 */
    get_linspace(p->hw_min, p->hw_max, n_hw, hw_lst);
    double *en = (double *)malloc(sizeof(double) * nw), *fd = (double *)malloc(sizeof(double) * n_ef * nw);
    zc *V = (zc *)malloc(sizeof(zc) * 3 * nw * nw), *U = (zc *)malloc(sizeof(zc) * nw * nw), *H_ka = (zc *)malloc(sizeof(zc) * 3 * nw * nw);
    zc *tmp = (zc *)malloc(sizeof(zc) * nw * nw), *M_in = (zc *)malloc(sizeof(zc) * nw * nw), *U_dag = (zc *)malloc(sizeof(zc) * nw * nw);
    zc *scr = (zc *)malloc(sizeof(zc) * 27 * n_hw);
    int rc = 0;
    for (int64_t lin = lin_begin + rank; lin < lin_end; lin += nranks) {
        double kpt[3]; mesh_kpt(p->mp_grid, lin, kpt);
        int info = mo->r ? get_wann_interp_pos(mo, p, recip, kpt, en, V, Ak, Omk, U, H_ka, wk)
                         : get_wann_interp(mo, p, recip, kpt, en, V, U, H_ka, tmp, M_in, U_dag);
        if (info != 0) { rc = -5; break; }
        band_sums_k(mo, p, &L, ef_lst, hw_lst, en, V, Ak, Omk, fd, scr, partials);
        partials[0] += 1.0;
    }
    free(ef_lst); free(hw_lst); free(en); free(fd); free(V); free(U); free(H_ka); free(tmp); free(M_in); free(U_dag); free(scr);
    free(Ak); free(Omk); free(wk);
    return rc;
}

/* Band sums on supplied eigenpairs: e_k (nw,nk), V_ka (3,nw,nw,nk).  partials += (test helper mirroring mepb_band_sums) */
int cpu_ref_band_sums(const mepb_model *mo, const mepb_params *p, int64_t nk, const double *e_k, const double *V_ka, const double *A_ka,
                      const double *Om_kab, double *partials) {
    const int nw = mo->nw, n_ef = p->n_eF, n_hw = p->n_hw;
    layout_t L = make_layout(p, mo->r != NULL);
    double *ef_lst = (double *)malloc(sizeof(double) * n_ef), *hw_lst = (double *)malloc(sizeof(double) * n_hw);
    get_linspace(p->eF_min, p->eF_max, n_ef, ef_lst);
    get_linspace(p->hw_min, p->hw_max, n_hw, hw_lst);
    double *fd = (double *)malloc(sizeof(double) * n_ef * nw);
    zc *scr = (zc *)malloc(sizeof(zc) * 27 * n_hw);
    for (int64_t ik = 0; ik < nk; ++ik) {
        band_sums_k(mo, p, &L, ef_lst, hw_lst, e_k + (int64_t)nw * ik, (const zc *)V_ka + (int64_t)3 * nw * nw * ik,
                    (mo->r && A_ka) ? (const zc *)A_ka + (int64_t)3 * nw * nw * ik : NULL,
                    (mo->r && Om_kab) ? (const zc *)Om_kab + (int64_t)9 * nw * nw * ik : NULL, fd, scr, partials);
        partials[0] += 1.0;
    }
    free(ef_lst); free(hw_lst); free(fd); free(scr);
    return 0;
}

/* nthreads workers standing in for MPI ranks; private accumulators summed at the end like MPI_REDUCE (core.f90:467-526) */
int cpu_ref_accumulate_mt(const mepb_model *mo, const mepb_params *p, int64_t lin_begin, int64_t lin_end, int nthreads,
                          double *partials) {
    int64_t n = make_layout(p, mo->r != NULL).total;
    if (nthreads < 1) nthreads = 1;
    double *priv = (double *)calloc((size_t)n * nthreads, sizeof(double));
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
    for (int r = 0; r < nthreads; ++r) {
        int e = cpu_ref_accumulate(mo, p, lin_begin, lin_end, r, nthreads, priv + (size_t)n * r);
        if (e != 0) rc = e;
    }
    for (int r = 0; r < nthreads; ++r) for (int64_t i = 0; i < n; ++i) partials[i] += priv[(size_t)n * r + i];
    free(priv);
    return rc;
}

/* Test helper (not a reference function): partials_abs[i] += sum_k |X_i(k)|, the un-cancelled magnitude of every
 * k-sum, so that parity tests can measure errors relative to sum |term| instead of |sum| (entries 0,1 and Ne are
 * plain sums, so cpu_ref_finalize of this buffer gives the per-tensor magnitudes with the same eF_idx). */
int cpu_ref_accumulate_abs(const mepb_model *mo, const mepb_params *p, int64_t lin_begin, int64_t lin_end, int nthreads,
                           double *partials_abs) {
    int64_t n = make_layout(p, mo->r != NULL).total;
    if (nthreads < 1) nthreads = 1;
    double *priv = (double *)calloc((size_t)n * nthreads * 2, sizeof(double));
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static, 1)
#endif
    for (int r = 0; r < nthreads; ++r) {
        double *acc = priv + (size_t)n * (2 * r), *one = priv + (size_t)n * (2 * r + 1);
        for (int64_t lin = lin_begin + r; lin < lin_end; lin += nthreads) {
            memset(one, 0, sizeof(double) * n);
            int e = cpu_ref_accumulate(mo, p, lin, lin + 1, 0, 1, one);
            if (e != 0) rc = e;
            for (int64_t i = 0; i < n; ++i) acc[i] += fabs(one[i]);
        }
    }
    for (int r = 0; r < nthreads; ++r) for (int64_t i = 0; i < n; ++i) partials_abs[i] += priv[(size_t)n * (2 * r) + i];
    free(priv);
    return rc;
}

/* core.f90:317-415 (first minimum, strict <), :419-603, k_space.f90:81-202 */
int cpu_ref_finalize(const mepb_model *mo, const mepb_params *p, const double *partials, mepb_result *out) {
    layout_t L = make_layout(p, mo->r != NULL);
    const int n_ef = p->n_eF, n_hw = p->n_hw, nval = p->valence_bands;
    double recip[9]; double vol = set_recip_latt(mo->a_latt, recip);
    double nk_mesh = (double)p->mp_grid[0] * (double)p->mp_grid[1] * (double)p->mp_grid[2];
    double norm = fabs(vol * nk_mesh);
    double n_k = partials[0];
    int opt = 0; double opt_err = fabs(partials[L.ne] / n_k - (double)nval);
    for (int ef = 0; ef < n_ef; ++ef) { double e = fabs(partials[L.ne + ef] / n_k - (double)nval); if (e < opt_err) { opt = ef; opt_err = e; } }
    out->n_k = (int64_t)n_k; out->eF_idx = opt; out->n_degenerate = (int64_t)partials[1];
    if (out->Ne) for (int ef = 0; ef < n_ef; ++ef) out->Ne[ef] = partials[L.ne + ef] / n_k;
    if (p->do_mep) {
        for (int q = 0; q < 9; ++q) {
            double ic = 0, lc = 0;
            for (int n0 = 0; n0 < nval; ++n0) { ic += partials[L.mep_ic + q + 9 * n0]; lc += partials[L.mep_lc + q + 9 * n0]; }
            if (out->mep_ic) out->mep_ic[q] = ic / norm;
            if (out->mep_lc) out->mep_lc[q] = lc / norm;
            if (out->mep_cs) {                                       /* cs(i,i,n0) summed over n0 <= nval; 0 without A, Om (mep_niu.f90:29) */
                double cs = 0.0;
                if (L.mep_cs >= 0 && (q == 0 || q == 4 || q == 8)) for (int n0 = 0; n0 < nval; ++n0) cs += partials[L.mep_cs + n0];
                out->mep_cs[q] = cs / norm;
            }
        }
        if (out->mep_bands && p->do_write_mep_bands)                 /* core.f90:591-599: ic + lc + cs per band */
            for (int q = 0; q < 9 * nval; ++q) {
                const int e = q % 9;
                const double cs = (L.mep_cs >= 0 && (e == 0 || e == 4 || e == 8)) ? partials[L.mep_cs + q / 9] : 0.0;
                out->mep_bands[q] = (partials[L.mep_ic + q] + partials[L.mep_lc + q] + cs) / norm;
            }
    }
    if (p->do_kubo) for (int q = 0; q < 9; ++q) {
        if (out->kubo_mep_ic) out->kubo_mep_ic[q] = partials[L.kmep_ic + q + 9 * opt] / norm;
        if (out->kubo_mep_lc) out->kubo_mep_lc[q] = partials[L.kmep_lc + q + 9 * opt] / norm;
        if (out->kubo_mep_cs) out->kubo_mep_cs[q] = (L.kmep_cs >= 0 && (q == 0 || q == 4 || q == 8)) ? partials[L.kmep_cs + opt] / norm : 0.0;
    }
    if (p->do_ahc) {
        if (out->ahc) for (int q = 0; q < 9; ++q) out->ahc[q] = partials[L.ahc + q + 9 * opt] / norm;
        int64_t sl = 2 * 9 * (int64_t)n_hw;
        if (out->ahc_velo) for (int64_t q = 0; q < sl; ++q) out->ahc_velo[q] = partials[L.velo + q + sl * opt] / norm;
        if (out->ohc_velo) for (int64_t q = 0; q < sl; ++q) out->ohc_velo[q] = partials[L.ohc + q + sl * opt] / norm;
    }
    if (p->do_opt) {
        int64_t sl = 2 * 9 * (int64_t)n_hw;
        /* kubo.f90:225-233: opt_symm = Re(herm) + i Im(aherm), opt_asymm = Re(aherm) + i Im(herm); sliced at eF_idx (core.f90:510-511) */
        for (int64_t q = 0; q < sl; q += 2) {
            double hr = 0, hi = 0, ar = 0, ai = 0;
            if (L.opt_h >= 0) { hr = partials[L.opt_h + q + sl * opt]; hi = partials[L.opt_h + q + 1 + sl * opt];
                                ar = partials[L.opt_ah + q + sl * opt]; ai = partials[L.opt_ah + q + 1 + sl * opt]; }
            if (out->opt_s) { out->opt_s[q] = hr / norm; out->opt_s[q + 1] = ai / norm; }
            if (out->opt_a) { out->opt_a[q] = ar / norm; out->opt_a[q + 1] = hi / norm; }
        }
        if (out->photo2) for (int64_t q = 0; q < 2 * 27 * (int64_t)n_hw * n_ef; ++q) out->photo2[q] = partials[L.photo + q] / norm;
    }
    if (p->do_gyro) {                                                /* every trial E_F is written (core.f90:523-530) */
        if (out->gyro_c) for (int64_t q = 0; q < 18 * (int64_t)n_ef; ++q) out->gyro_c[q] = partials[L.gyro_c + q] / norm;
        if (out->gyro_d) for (int64_t q = 0; q < 18 * (int64_t)n_ef; ++q) out->gyro_d[q] = partials[L.gyro_d + q] / norm;
        if (out->gyro_dw) for (int64_t q = 0; q < 18 * (int64_t)n_hw * n_ef; ++q) out->gyro_dw[q] = partials[L.gyro_dw + q] / norm;
    }
    return 0;
}

/* explicit k list: e_k (nw,nk), V_ka (3,nw,nw,nk) or NULL, H_k (nw,nw,nk) or NULL, H_ka_W (3,nw,nw,nk) or NULL (W gauge);
 * k-points are independent: OpenMP over k with private scratch (test helper speed only) */
int cpu_ref_interp(const mepb_model *mo, const mepb_params *p, int64_t nk, const double *kpt, double *e_k, double *V_ka,
                   double *H_k_out, double *H_ka_out) {
    if (!p_zheevd) return -1;
    const int nw = mo->nw;
    double recip[9]; set_recip_latt(mo->a_latt, recip);
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        zc *V = (zc *)malloc(sizeof(zc) * 3 * nw * nw), *U = (zc *)malloc(sizeof(zc) * nw * nw), *H_ka = (zc *)malloc(sizeof(zc) * 3 * nw * nw);
        zc *tmp = (zc *)malloc(sizeof(zc) * nw * nw), *M_in = (zc *)malloc(sizeof(zc) * nw * nw), *U_dag = (zc *)malloc(sizeof(zc) * nw * nw);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t ik = 0; ik < nk; ++ik) {
            if (H_k_out || H_ka_out) {
                ft_R_to_k(mo, p->use_cart_velo, recip, kpt + 3 * ik, U, H_ka);
                if (H_k_out) memcpy(H_k_out + 2 * (int64_t)nw * nw * ik, U, sizeof(zc) * nw * nw);
                if (H_ka_out) memcpy(H_ka_out + 2 * (int64_t)3 * nw * nw * ik, H_ka, sizeof(zc) * 3 * nw * nw);
            }
            if (e_k) {
                int info = get_wann_interp(mo, p, recip, kpt + 3 * ik, e_k + (int64_t)nw * ik, V, U, H_ka, tmp, M_in, U_dag);
                if (info) rc = -5;
                if (V_ka) memcpy(V_ka + 2 * (int64_t)3 * nw * nw * ik, V, sizeof(zc) * 3 * nw * nw);
            }
        }
        free(V); free(U); free(H_ka); free(tmp); free(M_in); free(U_dag);
    }
    return rc;
}

/* explicit k list with a position operator: e_k (nw,nk), V_ka (3,nw,nw,nk), A_ka (3,nw,nw,nk), Om_kab (3,3,nw,nw,nk) in the H gauge
 * (or, with do_gauge_trafo = 0, the W-gauge A / Om and the velocity get_velo builds from them) */
int cpu_ref_interp_pos(const mepb_model *mo, const mepb_params *p, int64_t nk, const double *kpt, double *e_k, double *V_ka,
                       double *A_ka, double *Om_kab) {
    if (!p_zheevd) return -1;
    if (!mo->r) return -4;
    const int nw = mo->nw;
    double recip[9]; set_recip_latt(mo->a_latt, recip);
    int rc = 0;
#ifdef _OPENMP
#pragma omp parallel
#endif
    {
        zc *U = (zc *)malloc(sizeof(zc) * nw * nw), *H_ka = (zc *)malloc(sizeof(zc) * 3 * nw * nw), *wk = (zc *)malloc(sizeof(zc) * 12 * nw * nw);
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
        for (int64_t ik = 0; ik < nk; ++ik) {
            int info = get_wann_interp_pos(mo, p, recip, kpt + 3 * ik, e_k + (int64_t)nw * ik, (zc *)V_ka + (int64_t)3 * nw * nw * ik,
                                           (zc *)A_ka + (int64_t)3 * nw * nw * ik, (zc *)Om_kab + (int64_t)9 * nw * nw * ik, U, H_ka, wk);
            if (info) rc = -5;
        }
        free(U); free(H_ka); free(wk);
    }
    return rc;
}
