// mepb_internal.cuh — shared between the kernels and the C-ABI host code of libmepb.so.
// sm_100a only.  All arithmetic FP64.
#pragma once
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mepb.h"

namespace mepb {

// ---------------------------------------------------------------------------------------------------------
// constants (reference: src/constants.f90:6-23, src/statistics.f90:32) — must be the same literals
// ---------------------------------------------------------------------------------------------------------
constexpr double kAuToEv = 27.211385;
constexpr double kBoltzEhK = 8.6173303 * 1e-5 / kAuToEv;
constexpr double kMinTemp = 1e-2;

// ---------------------------------------------------------------------------------------------------------
// "pair items": what the interpolation kernels hand to the frequency-sweep kernels through HBM.
//   For a band pair n<m (dE = E_m - E_n, fdm = f_m - f_n) let P(i,j) = V(i,n,m) V(j,m,n),
//   Asym(i,j) = P(i,j) - P(j,i), Sym(i,j) = P(i,j) + P(j,i).  Combining the (n,m) and (m,n) terms of
//   kubo.f90:111-135 and :172-194 (no hermiticity of V assumed; g(w) = 1/(dE^2 - z^2), z = w + i eta):
//       velo_ahc(i,j,w) = - sum fdm Im Asym(i,j) g(w)
//       kubo_ohc(i,j,w) = i sum fdm g(w) [ Asym(i,j) + (z/dE) Sym(i,j) ]
//   GENERAL item (20 doubles): [0] c = dE^2 + eta^2, [1..3] fdm Re Asym (01,02,12), [4..6] fdm Im Asym,
//       [7..12] (fdm/dE) Re Sym (00,11,22,01,02,12), [13..18] (fdm/dE) Im Sym, [19] pad
//   HERMITIAN item (10 doubles), used when H(R) is Hermitian-closed so that V_a is Hermitian and
//       Re Asym = Im Sym = 0:  [0] c, [1..3] fdm Im Asym, [4..9] (fdm/dE) Re Sym
//   Sweep accumulators per frequency (36 doubles): X1 = sum fdm ReAsym g (3 complex), X2 = sum fdm ImAsym g (3 complex),
//       Y1 = sum (fdm/dE) ReSym z g (6 complex), Y2 = sum (fdm/dE) ImSym z g (6 complex); the Hermitian kernel fills X2, Y1.
//   PHOTO item (58 doubles) for an ordered pair (n,l), l != n, df = f_l - f_n:
//     [0..53] T(a,b,c) = df V(a,n,l) sum_m V(b,l,m) V(c,m,n) / (E_n - E_m - i eta)   complex, index a+3b+9c
//     [54] xr = E_n - E_l     [55..57] zero pad — 29 double2 per item.  The tensor-core sweep reads four contraction units (pairs
//     of items) at once, one per lane quad; at 28 double2 their 896-byte stride put all four on the same shared-memory banks (ncu
//     r02k: 75 % of the wavefronts were conflicts), at 29 the stride is 32 mod 128 bytes and the fragment loads are conflict free
//     (sweep 65.1 -> 57.4 ms on C4, r03q).  The producers WRITE the pad: items start on 16-byte, not 32-byte, boundaries now and
//     a sector left partly unwritten costs a read-modify-write in L2 (r03q: producer 22.9 -> 28.0 ms without the pad store).
//     (Padding only the shared-memory copy — one 896-byte bulk copy per pair — was slower than either: 70.7 ms, r03r.)
//   photo2(a,b,c,w) = sum T(abc) (1/w^2) sum_{s=+-1} 1/(xr + s w - 2 i eta)          (photo.f90:44-81)
// ---------------------------------------------------------------------------------------------------------
constexpr int kAhcItemH = 10;   // Hermitian item (padding it to 12 for 256-bit stores gained nothing in the producer and cost the sweep 2 %: r03u)
constexpr int kAhcItemG = 20;   // general item
constexpr int kPhotoItem = 58;
constexpr int kSweepAcc = 36;   // per frequency: X1(3c) X2(3c) Y1(6c) Y2(6c)
constexpr int kSweepAccH = 18;  // Hermitian kernel: X2, Y1 (contiguous at offset 6 of the 36)
constexpr int kPhotoAcc = 54;   // per frequency: 27 complex

// static (frequency independent) per-block partial sums written by the interpolation kernels
//   [0] n_k  [1] n_degenerate  [2] Ne  [3] # AHC pair items emitted  [4] # photo items emitted (work counters for the
//   roofline)  [5..13] ahc(3,3)  [14..22] kubo_mep_ic  [23..31] kubo_mep_lc
//   [32 .. 32+9*nw) mep_niu_ic(3,3,n0)   [32+9*nw .. 32+18*nw) mep_niu_lc(3,3,n0)
constexpr int S_NK = 0, S_NDEG = 1, S_NE = 2, S_NIT = 3, S_NPH = 4, S_AHC = 5, S_KIC = 14, S_KLC = 23, S_NIU = 32;
__host__ __device__ inline int static_size(int nw) { return S_NIU + 18 * nw; }

struct DevModel {
    int nw, nR;
    const double2 *H;    // [nR][nw*nw], element n + nw*m (Fortran H_tb(n,m,R))
    const double *Rf;    // [nR][3] fractional R
    const double *cR;    // [nR][3] M.R  with M = 2 pi (use_cart_velo=F) or a_latt (T)       wann_interp.f90:268-271
    const double *tau;   // [nw][3] fractional centres or nullptr (Wannier convention)        wann_interp.f90:192
    const double *tM;    // [nw][3] M.tau (zeros when tau == nullptr)
    // sparse form of H(R) (CSR by matrix element n + nw*m), used by the warp-per-k Fourier sums when most entries are zero
    // (FeMn: 110 of 1728; the reference adds the zeros, file_io.f90:732-741 keeps them in H_tb) — nullptr -> dense loops
    const int *sp_ptr;   // [nw*nw + 1]
    const int *sp_R;     // [nnz] R index
    const double2 *sp_val;   // [nnz]
    int sp_nnz;
    int sp_up;           // 1: behind sp_ptr[nw*nw] follow {1, nw + 1 lane offsets, nw (nw + 1) / 2 elements n | m << 8}: upper-triangle schedule
    // padded (ELL) form of the UPPER triangle for the thread-per-k eigensolver (mepb_eig.cu): every element p <= q, in the
    // order q-major / p-minor, has ell_n slots (R index into the compact list of R points that carry any entry, value);
    // empty slots are (0, 0+0i).  ell_x = coordinates whose phases a thread needs: the ell_nR used R points, then
    // (tight-binding convention) the nw centres tau_n.  nullptr -> dense loops.
    int ell_n, ell_nR;
    const int *ell_R;        // [ell_n][nw(nw+1)/2]
    const double2 *ell_val;  // [ell_n][nw(nw+1)/2]
    const double *ell_x;     // [ell_nR (+ nw)][3]
};

struct KSource {          // where the k-points of a launch come from
    const double *list;   // [nk][3] explicit fractional k (band mode / parity) or nullptr -> MP mesh
    long long lin_begin;  // first 0-based linear mesh index of this launch
    long long nk;         // k-points in this launch
    int N1, N2, N3;       // mp_grid
};

struct Jobs {
    int do_gauge, do_mep, do_kubo, do_ahc, do_opt;
    int nval;
    int herm;             // H(R) Hermitian-closed -> Hermitian items (10 doubles) else general items (20)
    int zero_T;           // T_kelvin <= min_temp -> step function                            statistics.f90:74-87
    int hka_upper;        // general path: d_aH(k) in HBM holds the upper triangle only (Hermitian; mirrored on load)
    double eF, inv_kT, eta, kubo_tol;
};

struct InterpOut {
    // list-mode outputs (any may be nullptr)
    double *e_out;        // (nw, nk)
    double2 *V_out;       // (3, nw, nw, nk)
    double2 *U_out;       // (nw, nw, nk) eigenvectors in columns (general path only; consumed by the position-operator kernels)
    double2 *Hk_out;      // (nw, nw, nk)      W gauge
    double2 *Hka_out;     // (3, nw, nw, nk)   W gauge
    // response-mode outputs
    double *items_ahc;    // [nk][max_pairs][10 or 20]
    int *cnt_ahc;         // [nk]
    double *items_photo;  // [nk][nw*(nw-1)][56]
    int *cnt_photo;       // [nk]
    double *block_static; // [gridDim.x][static_size(nw)]
    int *err_flag;        // set to 1 when the eigensolver fails to converge
    // inputs from the thread-per-k eigensolver (mepb_eig.cu); nullptr -> the kernel runs its own warp-level Jacobi
    const double *e_in;   // (nw, nk) ascending
    const double2 *U_in;  // (nw, nw, nk) eigenvectors in columns, same order
    // band sums on supplied eigenpairs (mepb_band_sums, position-operator branch): H-gauge velocities V_ka(a,n,m) in the
    // user layout a + 3 (n + nw m) per k-point, with e_in the energies; the kernels then skip Fourier sums, eigensolve
    // and rotation and run only their band-sum sections
    const double2 *V_in;  // (3, nw, nw, nk) or nullptr
    // bands_sub_kernel -> mep_sums_kernel: H-gauge velocities in the padded shared-memory layout, 3 nw (nw + 1) double2 per k-point
    double2 *V_pad;       // or nullptr (no MEP job)
};

// ---------------------------------------------------------------------------------------------------------
// dense Fourier contraction (mepb_large.cu): host-built plan
// ---------------------------------------------------------------------------------------------------------
constexpr int kFourierTK = 64;    // k-points per CTA
constexpr int kFourierTN = 32;    // matrix elements per CTA
constexpr int kFourierRC = 16;    // R points per pipeline stage
constexpr int kFourierMaxU = 96;  // distinct lattice coordinates over the three axes (else: one sincospi per (k,R))

struct FourierPlan {
    int nel, ntiles, nRp;         // computed elements (upper triangle if herm, else nw^2), tiles of kFourierTN, nR padded to kFourierRC
    int herm;                     // mirror (m,n) = conj(n,m)
    int upper_only;               // herm: store the computed upper triangle only (consumers that mirror on load); 0 = full matrices
    int direct;                   // 1: phases by sincospi per (k,R); 0: product of three per-axis table entries
    int u1, u2, u3, utot;         // distinct coordinates per axis
    double uval[kFourierMaxU];    // the distinct coordinates, axis 1 first
    const int *ridx;              // [nRp] packed table rows i1 | i2 << 8 | i3 << 16
    const int *elem;              // [ntiles*kFourierTN] packed n | m << 16
    const double2 *Hp;            // [ntiles][nRp][kFourierTN]  H(R) elements of each tile, zero padded
    const double *Rf;             // [nRp][3] fractional R (zero padded)
    const double *cR;             // [nRp][3] M.R (zero padded)
};

// ---------------------------------------------------------------------------------------------------------
// mesh-separable Fourier sums (mepb_large.cu, mesh mode): on a Monkhorst-Pack mesh the phase factorises over the axes,
// e^{2 pi i k.R} = p1(q1, R1) p2(q2, R2) p3(q3, R3), so the nR-term sum per k-point is contracted axis by axis:
//   stage A  per mesh plane q3:        G1(a,b; t)   = sum_c p3(q3,c) {1, u3(c)} H(a,b,c)            t = 0, 1
//   stage B  per mesh line (q2,q3):    G2(a; t)     = sum_b p2(q2,b) {G1(a,b;0), u2(b) G1(a,b;0), G1(a,b;1)}   t = 0, 1, 2
//   stage C  per k-point of the line:  H, d_aH      = sum_{a,t} p1(q1,a) w(a,t) G2(a;t)      (3 n1 terms instead of nR)
// (a, b, c index the distinct lattice coordinates u1, u2, u3 of the R set; the weights w carry c_a(R) = (M.R)_a, which is
// linear in R.)  Stage C is the dense contraction kernel run on the per-line table.
// ---------------------------------------------------------------------------------------------------------
struct SepPlan {
    int on;                  // 0 -> not available (explicit k lists, or too many distinct coordinates)
    int n1, n2, n3;          // distinct lattice coordinates per axis
    int nrp;                 // rows of a line table: 3 n1 padded to kFourierRC
    int nelp;                // ntiles * kFourierTN
    double u2[32], u3[32];   // distinct coordinates of axes 2, 3 (axis 1 enters through cw4)
    const int *map3;         // [n1*n2*n3] index into the R list or -1
    const double2 *P1, *P2, *P3;   // [n_a][N_a] e^{2 pi i k_a(q) u_a}
    const double *cw4;       // [nrp][4] weights (H, d_1, d_2, d_3) of a line-table row
    const int *ra;           // [nrp] axis-1 coordinate index of a line-table row
    double2 *G1;             // [planes][n1*n2][2][nelp]      (per stream)
    double2 *G2;             // [lines][ntiles][nrp][kFourierTN] (per stream)
    long long line0;         // first mesh line (q2 + N2 q3) covered by G2
    int nlines, q3_0, nq3;
};

// ---------------------------------------------------------------------------------------------------------
// Householder + implicit-QL eigensolver of the general path (mepb_tri.cu): per-k intermediates in HBM
//   tridiag_kernel   H(k) -> reflectors v_s, tau_s, real tridiagonal (d, e)           (LAPACK zhetd2 'L' arithmetic)
//   tql_kernel       thread per k: implicit QL on (d, e); eigenvalues, sort ranks, and the LOG of plane rotations
//   zrot_kernel      replays the log on Z = 1 (one thread per row of Z, rows in registers) -> real eigenvectors of T
//   the consumer (eig_rot_large_kernel) applies the reflectors to Z: U = H_0 H_1 ... H_{n-2} Z
// ---------------------------------------------------------------------------------------------------------
struct TriBuf {
    int ne;                 // padded matrix size (32 or 64); 0 -> the Jacobi solver inside eig_rot_large_kernel is used
    int hdr_cap, log_cap;   // per-k capacity of the chase headers / rotation log
    double2 *Vr;            // [k][ne*ne]   Vr[s*ne + i] = v_s(i)  (0 for i <= s, 1 at i = s+1)
    double2 *tau;           // [k][ne]
    double *d, *e;          // [k][ne]      T(i,i), T(i+1,i)
    double *E;              // [k][ne]      ascending eigenvalues
    unsigned char *rank;    // [k][ne]      position of tql's eigenvalue j in the ascending order
    unsigned int *hdr;      // [k][hdr_cap] one word per QL chase: l | m << 8  (rotations in planes m-1 ... l)
    int *nhdr;              // [k][2]  number of chases, number of logged rotations
    double2 *log;           // [k][log_cap] (c, s) per rotation, in execution order
    double *Z;              // [k][ne*ne]   Z[col*ne + row], eigenvectors of T in tql's (unsorted) column order
};
constexpr int kTriHdrCap = 512;

// ---------------------------------------------------------------------------------------------------------
// complex helpers on double2
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 cmulc(double2 a, double2 b) {  // a * conj(b)
    return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.y, b.x, -a.x * b.y));
}
__device__ __forceinline__ double2 cconj(double2 a) { return make_double2(a.x, -a.y); }
__device__ __forceinline__ void cfma(double2 &acc, double2 a, double2 b) {  // acc += a*b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
__device__ __forceinline__ void cfmac(double2 &acc, double2 a, double2 b) {  // acc += conj(a)*b
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ double2 cscale(double s, double2 a) { return make_double2(s * a.x, s * a.y); }
__device__ __forceinline__ double2 crcp(double2 a) {  // 1/a  (IEEE division: inf/NaN propagate like the reference)
    double d = 1.0 / fma(a.x, a.x, a.y * a.y);
    return make_double2(a.x * d, -a.y * d);
}
// reciprocal for the sweep inner loops: MUFU.RCP64H seed + 2 Newton steps (7.4 DFMA-equivalents measured vs 11 for '/')
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double fermi(double e, const Jobs &j) {  // statistics.f90:41-61
    if (j.zero_T) return (e < j.eF) ? 1.0 : 0.0;
    return 1.0 / (1.0 + exp((e - j.eF) * j.inv_kT));
}

// ---------------------------------------------------------------------------------------------------------
// PTX helpers (TMA bulk copy)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 16-byte asynchronous copy global -> shared (LDGSTS, no register stage), its group commit and the wait for every group of the thread
__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// one 256-bit global store (sm_100: STG.E.256) of two double2 to a 32-byte aligned address
__device__ __forceinline__ void st_global_256(double2 *p, const double2 a, const double2 b) {
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
// one 256-bit read-only global load (LDG.E.256) of two double2 from a 32-byte aligned address
__device__ __forceinline__ void ld_global_nc_256(const double2 *p, double2 &a, double2 &b) {
    asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(a.x), "=d"(a.y), "=d"(b.x), "=d"(b.y) : "l"(p));
}
// AHC item -> its slot in HBM: 5 128-bit stores (Hermitian, 80 bytes) or 5 256-bit stores (general, 160 bytes = whole sectors)
__device__ __forceinline__ void store_ahc_item(double *dst, const double *item, const bool herm) {
    double2 *d2 = reinterpret_cast<double2 *>(dst);
    if (herm) {
#pragma unroll
        for (int j = 0; j < kAhcItemH / 2; ++j) d2[j] = make_double2(item[2 * j], item[2 * j + 1]);
    } else {
#pragma unroll
        for (int j = 0; j < kAhcItemG / 4; ++j)
            st_global_256(d2 + 2 * j, make_double2(item[4 * j], item[4 * j + 1]), make_double2(item[4 * j + 2], item[4 * j + 3]));
    }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1/sqrt(x) for normal x > 0: MUFU.RSQ64H seed + one cubic step  y' = y + y e (1/2 + 3/8 e), e = 1 - x y^2
__device__ __forceinline__ double fast_rsqrt3(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = x * y;
    const double e = fma(-t, y, 1.0);
    const double u = fma(0.375, e, 0.5);
    const double v = y * e;
    return fma(v, u, y);
}
// reciprocal with a cubic (Halley) correction of the MUFU.RCP64H seed: r' = r + r (e + e^2), e = 1 - x r
__device__ __forceinline__ double fast_rcp3(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
}

// ---------------------------------------------------------------------------------------------------------
// launchers (defined in the .cu files)
// ---------------------------------------------------------------------------------------------------------
bool small_supported(int nw, int nR);
cudaError_t launch_interp_small(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs,
                                int grid, cudaStream_t st);
int interp_small_grid(int nw, int nR, int sm_count, long long nk);
bool eig_thread_supported(int nw);
// NW-lanes-per-k band-sum kernel (mepb_sub.cu): response mode, num_wann 4 / 8, eigenpairs from launch_eig_small
bool sub_supported(int nw, int nR);
int bands_sub_grid(int nw, int nR, int sm_count, long long nk, int nnz);   // nnz: staged sparse records (0: dense)
// MEP (Niu + Kubo) band sums of the num_wann 4 / 8 path from the velocities bands_sub_kernel left in HBM (mepb_mep.cu); `rows`:
// grid rows of static_size(nw) doubles in the block_static layout (only the MEP slots are non-zero)
size_t mep_vpad_elems(int nw);
int mep_sums_grid(int nw, int sm_count, long long nk);
cudaError_t launch_mep_sums(int nw, long long nk, const Jobs &jobs, const double *e_in, const double2 *V_pad, double *rows, int grid,
                            cudaStream_t st);
cudaError_t launch_bands_sub(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs, int grid,
                             cudaStream_t st);
cudaError_t launch_eig_small(const DevModel &m, const KSource &ks, double *E_out, double2 *U_out, int *err_flag, int sm_count,
                             cudaStream_t st);

cudaError_t launch_sweep_ahc(bool herm, const double *items, const int *cnt, int max_pairs, long long nk, int n_hw,
                             const double *hw, double eta, double *block_out, int grid, cudaStream_t st);
int sweep_ahc_grid(int sm_count, long long nk, int n_hw, int *slots_out, bool herm, int max_pairs);
// paired: items of the band pairs (n, l), (l, n) sit in consecutive slots (written by bands_sub_kernel)
cudaError_t launch_sweep_photo(const double *items, const int *cnt, int max_items, long long nk, int n_hw, const double *hw,
                               double eta, double *block_out, int grid, bool paired, cudaStream_t st);
int sweep_photo_grid(int sm_count, long long nk, int n_hw, int *slots_out, int max_items);
// tensor-core (DMMA) formulation of both sweeps (mepb_sweep_mma.cu), used for n_hw >= 16 unless MEPB_SWEEP_FMA is set
struct MmaShape { int NT, G, wtiles, warps, passes, shared; };   // tiles per warp, k-groups, warps per group, warps per CTA, grid.y, shared tile
constexpr int kMmaAhcWarps = 8, kMmaPhotoWarps = 13;      // consumer warps per CTA (+ one TMA producer warp)
MmaShape sweep_mma_shape(int n_hw, int wmax, int nt_max, bool shared_ok);
bool sweep_mma_enabled(int n_hw);
bool sweep_photo_mma_fits(int max_items);
int sweep_ahc_mma_bps(bool herm, int NT);
cudaError_t launch_sweep_ahc_mma(bool herm, const double *items, const int *cnt, int max_pairs, long long nk, int n_hw,
                                 const double *hw, double eta, double *block_out, int grid, int KC, cudaStream_t st);
cudaError_t launch_sweep_photo_mma(const double *items, const int *cnt, int max_items, long long nk, int n_hw, const double *hw,
                                   double eta, double *block_out, int grid, bool paired, cudaStream_t st);
// out[(c / group) * out_stride + out_off + c % group] += sum_r in[r*cols + c]
cudaError_t launch_reduce_rows(const double *in, int rows, int cols, double *out_accum, int group, int out_stride, int out_off,
                               cudaStream_t st);

// general path, num_wann <= 64 (mepb_large.cu)
bool large_supported(int nw);
size_t fourier_dense_smem(const FourierPlan &fp, int nw, bool has_tau);
cudaError_t launch_fourier_dense(const FourierPlan &fp, const DevModel &m, const KSource &ks, double2 *Hk, double2 *Hka,
                                 cudaStream_t st);
// mesh mode: stages A, B, C of the separable transform for the k range of ks (sp.G1/G2 sized by the caller); 3 launches
cudaError_t launch_fourier_mesh(const FourierPlan &fp, SepPlan sp, const DevModel &m, const KSource &ks, double2 *Hk,
                                double2 *Hka, cudaStream_t st);
cudaError_t launch_phase_table(double2 *P, const double *u, int n, int N, cudaStream_t st);
// packed element tiles Hp of a real-space operator (fp.elem must be on the device already)
cudaError_t launch_pack_hp(const FourierPlan &fp, const double2 *src, int stride, int off, int nw, int nR, double2 *Hp, cudaStream_t st);
cudaError_t launch_hka_to_user(const double2 *in, double2 *out, int nw2, long long nk, cudaStream_t st);
int eig_large_grid(int sm_count, long long nk);
cudaError_t launch_eig_rot_large(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs,
                                 const double2 *Hk, const double2 *Hka, double2 *Vscr, const TriBuf &tb, int grid, cudaStream_t st);
// Householder + QL eigensolver stages (mepb_tri.cu); tb.ne must be tri_padded_size(nw)
int tri_padded_size(int nw);
size_t tri_log_cap(int ne);
cudaError_t launch_tridiag(int nw, long long nk, const double2 *Hk, const TriBuf &tb, int sm_count, cudaStream_t st);
cudaError_t launch_tql(int nw, long long nk, const TriBuf &tb, int *err_flag, cudaStream_t st);
cudaError_t launch_zrot(int nw, long long nk, const TriBuf &tb, int sm_count, cudaStream_t st);

// position-operator branch (mepb_posop.cu)
size_t posop_gauge_smem(int nw);
int posop_gauge_grid(int sm_count, long long nk);
cudaError_t launch_posop_gauge(int nw, long long nk, int do_gauge, int mode, double kubo_tol, bool has_pos, const double *E,
                               const double2 *U, const double2 *Hka, const double2 *const Aw[3], const double2 *const dA[3], double2 *V,
                               double2 *AH, double2 *OmV, double2 *scratch, int grid, cudaStream_t st);
int posop_row_len(int nval, int n_hw);
int posop_hw_block(int n_hw);
int posop_sums_grid(int sm_count, long long nk);
cudaError_t launch_posop_sums(int nw, long long nk, const Jobs &jb, bool has_pos, bool do_gyro, int n_hw, const double *hw, const double *E,
                              const double2 *V, const double2 *AH, const double2 *OmV, double *rows, int grid, cudaStream_t st);

}  // namespace mepb
