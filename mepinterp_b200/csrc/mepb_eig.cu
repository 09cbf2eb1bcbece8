// mepb_eig.cu — batched Hermitian eigensolver for small num_wann: ONE THREAD PER k-POINT.
//
// Replaces the per-k LAPACK zheevd('V','U') of the reference (src/matrix_math.f90:192-226, called from
// src/wann_interp.f90:95) for nw in {4, 8}.  Each thread
//   1. builds the upper triangle of H(k) = sum_R e^{2 pi i k.R} H(R) in registers           (wann_interp.f90:206-279)
//   2. diagonalises it by cyclic two-sided Jacobi: the matrix lives in registers in "tournament order" — the pairs of a
//      round are always the positions (0,1),(2,3),.. and the Brent-Luk permutation is applied to rows and columns after
//      every round, so every register index is a compile-time constant (no shuffles, no barriers, no divergence inside
//      a round); the eigenvector matrix U lives in shared memory, one 16-byte column of lanes per matrix element
//      (conflict-free LDS.128/STS.128)
//   3. writes the ascending eigenvalues and the matching eigenvector columns to HBM for the band-sum kernel.
// The rotation formulas are those of the warp-per-k solver in mepb_small.cu (validated against zheevd to 1e-15).
//
// Default solver since r01h: HOUSEHOLDER + IMPLICIT QL, still one thread per k and every index a compile-time constant:
//   a. zhetd2('L') arithmetic on the register-resident lower triangle (reflector + p = tau A v + rank-2 update per column);
//      the reflectors (v_s, tau_s) are parked in shared memory ([slot][thread], 448 B per thread at nw = 8)
//   b. implicit-shift QL (EISPACK tql2 recurrence) on (d, e) with the REAL eigenvector matrix Z of the tridiagonal in
//      registers: the chase over planes i = m-1 .. l is unrolled over all planes and each plane guarded by (l <= i < m),
//      so a warp skips the planes none of its threads needs; dynamic d[l], e[l], d[m] accesses are select chains
//   c. back-transformation u_j = H_0 H_1 ... H_{n-2} z_j one column at a time, written at its ascending rank.
// ~4x fewer FP64 instructions than the cyclic Jacobi (6 sweeps x 7 rounds x ~1000 DFMA); Jacobi stays selectable with
// MEPB_SMALL_EIG=jacobi (it is the independent cross-check of the parity tests).
#include <cstdlib>
#include <string>
#include "mepb_internal.cuh"

namespace mepb {

namespace {

template <int NW> struct EigCfg;
template <> struct EigCfg<4> { static constexpr int TPB = 256; };
template <> struct EigCfg<8> { static constexpr int TPB = 224; };   // 224 threads x 1 KB of U = 224 KB of shared memory

// Brent-Luk tournament: new position i takes the row/column that sat at position sigma(i)
template <int NW> __host__ __device__ constexpr int sigma(int i) {
    constexpr int NP = NW / 2;
    if (NP == 1) return i;
    if (i == 0) return 0;
    if (i == 2) return 1;
    if (i == 2 * NP - 1) return 2 * (NP - 1);
    if (i % 2 == 0) return i - 2;      // top row shifts right
    return i + 2;                       // bottom row shifts left
}

// One Jacobi round on the register-resident matrix: rotation parameters of the pairs (2i, 2i+1), two-sided update of the
// off-diagonal 2x2 blocks (rotation formulas of jacobi_warp in mepb_small.cu).
template <int NW>
__device__ __forceinline__ void jacobi_round_A(double (&ar)[NW][NW], double (&ai)[NW][NW], double (&rc)[NW / 2],
                                               double (&rsx)[NW / 2], double (&rsy)[NW / 2], double &off2) {
    constexpr int NP = NW / 2;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const int p = 2 * i, q = 2 * i + 1;
        const double a = ar[p][p], d = ar[q][q], bx = ar[p][q], by = ai[p][q];
        const double bb = fma(bx, bx, by * by);
        const bool rot = bb > 1e-300;
        const double zeta = 0.5 * (d - a), az = fabs(zeta);
        const double hh = rot ? fma(zeta, zeta, bb) : 1.0;
        const double rh = fast_rsqrt3(hh);
        double h = hh * rh;                                    // sqrt(zeta^2 + |b|^2) ...
        h = fma(fma(-h, h, hh), 0.5 * rh, h);                  // ... plus one Heron step (no division)
        const double y = 2.0 * h * (az + h);
        const double ry = rot ? fast_rsqrt3(y) : 0.0;
        const double sg = zeta < 0.0 ? -1.0 : 1.0;
        rc[i] = rot ? (az + h) * ry : 1.0;
        const double sr = sg * ry;
        rsx[i] = sr * bx; rsy[i] = sr * by;
        const double tb = sg * bb * (2.0 * h) * ry * ry;
        ar[p][p] = a - tb; ar[q][q] = d + tb; ar[p][q] = 0.0; ai[p][q] = 0.0;
        off2 += rot ? bb : 0.0;
    }
    // off-diagonal 2x2 blocks (pair i rows, pair k columns), i < k:  B <- J_i^dagger B J_k
#pragma unroll
    for (int i = 0; i < NP; ++i)
#pragma unroll
        for (int k = i + 1; k < NP; ++k) {
            const int p = 2 * i, q = 2 * i + 1, pp = 2 * k, qq = 2 * k + 1;
            const double ci = rc[i], six = rsx[i], siy = rsy[i], cj = rc[k], sjx = rsx[k], sjy = rsy[k];
            const double bppx = ar[p][pp], bppy = ai[p][pp], bpqx = ar[p][qq], bpqy = ai[p][qq];
            const double bqpx = ar[q][pp], bqpy = ai[q][pp], bqqx = ar[q][qq], bqqy = ai[q][qq];
            // columns: t(.,pp) = c x - conj(s) y ; t(.,qq) = s x + c y
            const double tppx = cj * bppx - (sjx * bpqx + sjy * bpqy), tppy = cj * bppy - (sjx * bpqy - sjy * bpqx);
            const double tqpx = cj * bqpx - (sjx * bqqx + sjy * bqqy), tqpy = cj * bqpy - (sjx * bqqy - sjy * bqqx);
            const double tpqx = sjx * bppx - sjy * bppy + cj * bpqx, tpqy = sjx * bppy + sjy * bppx + cj * bpqy;
            const double tqqx = sjx * bqpx - sjy * bqpy + cj * bqqx, tqqy = sjx * bqpy + sjy * bqpx + cj * bqqy;
            // rows: out_p = c t_p - s t_q ; out_q = conj(s) t_p + c t_q
            ar[p][pp] = ci * tppx - (six * tqpx - siy * tqpy); ai[p][pp] = ci * tppy - (six * tqpy + siy * tqpx);
            ar[q][pp] = six * tppx + siy * tppy + ci * tqpx;   ai[q][pp] = six * tppy - siy * tppx + ci * tqpy;
            ar[p][qq] = ci * tpqx - (six * tqqx - siy * tqqy); ai[p][qq] = ci * tpqy - (six * tqqy + siy * tqqx);
            ar[q][qq] = six * tpqx + siy * tpqy + ci * tqqx;   ai[q][qq] = six * tpqy - siy * tpqx + ci * tqqy;
        }
}

// Tournament permutation of rows and columns (compile-time indices: register moves only)
template <int NW>
__device__ __forceinline__ void tournament_permute(double (&ar)[NW][NW], double (&ai)[NW][NW], int (&perm)[NW]) {
    double br[NW][NW], bi[NW][NW];
#pragma unroll
    for (int j = 0; j < NW; ++j)
#pragma unroll
        for (int i = 0; i <= j; ++i) {
            const int si = sigma<NW>(i), sj = sigma<NW>(j);
            if (si <= sj) { br[i][j] = ar[si][sj]; bi[i][j] = ai[si][sj]; }
            else { br[i][j] = ar[sj][si]; bi[i][j] = -ai[sj][si]; }
        }
    int pn[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) pn[i] = perm[sigma<NW>(i)];
#pragma unroll
    for (int j = 0; j < NW; ++j) {
        perm[j] = pn[j];
#pragma unroll
        for (int i = 0; i <= j; ++i) { ar[i][j] = br[i][j]; ai[i][j] = (i == j) ? 0.0 : bi[i][j]; }
    }
}

// k-point of the launch (k_space.f90:351-375) and the upper triangle of H(k) (wann_interp.f90:206-279), in registers
template <int NW>
__device__ __forceinline__ void fourier_upper(const DevModel &m, const KSource &ks, const long long ik, const int nR,
                                              const bool has_tau, double (&ar)[NW][NW], double (&ai)[NW][NW],
                                              double2 *phase = nullptr, const int pstride = 0) {
    constexpr int NW2 = NW * NW;
    // ------------------------------------------------------------------ k-point (k_space.f90:351-375)
    double k1, k2, k3;
    if (ks.list) { k1 = ks.list[3 * ik]; k2 = ks.list[3 * ik + 1]; k3 = ks.list[3 * ik + 2]; }
    else {
        const long long lin = ks.lin_begin + ik;
        long long q1, q2, q3;
        if (lin < 0x7fffffffLL) {     // 32-bit divisions (a 64-bit one is ~60 instructions)
            const unsigned l32 = (unsigned)lin, n1 = (unsigned)ks.N1, n2 = (unsigned)ks.N2;
            const unsigned t = l32 / n1;
            q1 = l32 - t * n1 + 1; const unsigned t2 = t / n2; q2 = t - t2 * n2 + 1; q3 = t2 + 1;
        } else { q1 = lin % ks.N1 + 1; q2 = (lin / ks.N1) % ks.N2 + 1; q3 = lin / ((long long)ks.N1 * ks.N2) + 1; }
        k1 = (2.0 * (double)q1 - (double)ks.N1 - 1.0) / (2.0 * (double)ks.N1);
        k2 = (2.0 * (double)q2 - (double)ks.N2 - 1.0) / (2.0 * (double)ks.N2);
        k3 = (2.0 * (double)q3 - (double)ks.N3 - 1.0) / (2.0 * (double)ks.N3);
    }
    // ------------------------------------------------------------------ Fourier sum, upper triangle (zheevd 'U')
#pragma unroll
    for (int p = 0; p < NW; ++p)
#pragma unroll
        for (int q = 0; q < NW; ++q) { ar[p][q] = 0.0; ai[p][q] = 0.0; }
    if (phase) {
        // sparse H(R) (FeMn: 110 of 1728 entries, 19 of 27 R points used): the phases of the used R points and of the
        // centres tau_n go to this thread's shared-memory column in ONE rolled loop (a single inlined sincospi), then the
        // padded element lists (ELL, warp-uniform) are summed slot by slot: per slot 36 independent (index, value, phase)
        // loads and 4 DFMA each — no dependent pointer chasing, ~10x fewer DFMA than the dense loops
        const int nx = m.ell_nR + (has_tau ? NW : 0);
#pragma unroll 1
        for (int R = 0; R < nx; ++R) {
            const double x = 2.0 * (k1 * __ldg(&m.ell_x[3 * R]) + k2 * __ldg(&m.ell_x[3 * R + 1]) + k3 * __ldg(&m.ell_x[3 * R + 2]));
            double s, c; sincospi(x, &s, &c);
            phase[R * pstride] = make_double2(c, s);
        }
        constexpr int NEL = NW * (NW + 1) / 2;
#pragma unroll 1
        for (int j = 0; j < m.ell_n; ++j) {
            const int *Rj = m.ell_R + j * NEL;
            const double2 *Vj = m.ell_val + j * NEL;
#pragma unroll
            for (int q = 0; q < NW; ++q)
#pragma unroll
                for (int p = 0; p <= q; ++p) {
                    const int el = q * (q + 1) / 2 + p;
                    const double2 ph = phase[__ldg(Rj + el) * pstride];
                    const double2 h = __ldg(Vj + el);
                    ar[p][q] = fma(ph.x, h.x, ar[p][q]); ar[p][q] = fma(-ph.y, h.y, ar[p][q]);
                    if (p != q) { ai[p][q] = fma(ph.x, h.y, ai[p][q]); ai[p][q] = fma(ph.y, h.x, ai[p][q]); }
                }
        }
        if (has_tau) {   // H(n,m) <- conj(d_n) H(n,m) d_m, d_n = e^{2 pi i k.tau_n}
            double dr[NW], di[NW];
#pragma unroll
            for (int n = 0; n < NW; ++n) { const double2 dn = phase[(m.ell_nR + n) * pstride]; dr[n] = dn.x; di[n] = dn.y; }
#pragma unroll
            for (int q = 0; q < NW; ++q)
#pragma unroll
                for (int p = 0; p < q; ++p) {
                    const double wr = dr[q] * dr[p] + di[q] * di[p], wi = di[q] * dr[p] - dr[q] * di[p];   // d_q conj(d_p)
                    const double hr = ar[p][q], hi = ai[p][q];
                    ar[p][q] = hr * wr - hi * wi; ai[p][q] = hr * wi + hi * wr;
                }
        }
        return;
    }
    for (int R = 0; R < nR; ++R) {
        const double x = 2.0 * (k1 * __ldg(&m.Rf[3 * R]) + k2 * __ldg(&m.Rf[3 * R + 1]) + k3 * __ldg(&m.Rf[3 * R + 2]));
        double s, c; sincospi(x, &s, &c);
        const double2 *HR = m.H + (size_t)R * NW2;
#pragma unroll
        for (int q = 0; q < NW; ++q)
#pragma unroll
            for (int p = 0; p <= q; ++p) {
                const double2 h = __ldg(&HR[p + NW * q]);
                ar[p][q] = fma(c, h.x, ar[p][q]); ar[p][q] = fma(-s, h.y, ar[p][q]);
                if (p != q) { ai[p][q] = fma(c, h.y, ai[p][q]); ai[p][q] = fma(s, h.x, ai[p][q]); }
            }
    }
    if (has_tau) {   // tight-binding convention: H(n,m) <- conj(d_n) H(n,m) d_m, d_n = e^{2 pi i k.tau_n}
        double dr[NW], di[NW];
#pragma unroll
        for (int n = 0; n < NW; ++n) {
            const double x = 2.0 * (k1 * __ldg(&m.tau[3 * n]) + k2 * __ldg(&m.tau[3 * n + 1]) + k3 * __ldg(&m.tau[3 * n + 2]));
            sincospi(x, &di[n], &dr[n]);
        }
#pragma unroll
        for (int q = 0; q < NW; ++q)
#pragma unroll
            for (int p = 0; p < q; ++p) {
                const double wr = dr[q] * dr[p] + di[q] * di[p], wi = di[q] * dr[p] - dr[q] * di[p];   // d_q conj(d_p)
                const double hr = ar[p][q], hi = ai[p][q];
                ar[p][q] = hr * wr - hi * wi; ai[p][q] = hr * wi + hi * wr;
            }
    }
}

template <int NW>
__global__ void __launch_bounds__(EigCfg<NW>::TPB, 1)
eig_small_kernel(const DevModel m, const KSource ks, double *__restrict__ E_out, double2 *__restrict__ U_out,
                 int *__restrict__ err_flag) {
    constexpr int TPB = EigCfg<NW>::TPB, NP = NW / 2, NW2 = NW * NW;
    extern __shared__ __align__(16) double2 Us[];            // [NW2][TPB], element (row + NW*col)
    const int tid = threadIdx.x;
    const int nR = m.nR;
    const bool has_tau = (m.tau != nullptr);

    for (long long ik = (long long)blockIdx.x * TPB + tid; ik < ks.nk; ik += (long long)gridDim.x * TPB) {
        double ar[NW][NW], ai[NW][NW];                         // [p][q], p <= q used
        fourier_upper<NW>(m, ks, ik, nR, has_tau, ar, ai);
        // ------------------------------------------------------------------ Jacobi
        double nrm = 0.0;
#pragma unroll
        for (int q = 0; q < NW; ++q)
#pragma unroll
            for (int p = 0; p <= q; ++p)
                nrm += (p == q) ? ar[p][q] * ar[p][q] : 2.0 * (ar[p][q] * ar[p][q] + ai[p][q] * ai[p][q]);
#pragma unroll
        for (int e = 0; e < NW2; ++e) Us[e * TPB + tid] = make_double2((e % NW) == (e / NW) ? 1.0 : 0.0, 0.0);
        int perm[NW];                                          // position -> storage column of U
#pragma unroll
        for (int i = 0; i < NW; ++i) perm[i] = i;
        bool ok = (nrm >= 0.0);                                // NaN -> fail
        if (nrm > 0.0) {
            const double tol = 1e-26 * nrm;
            ok = false;
            for (int sweep = 0; sweep < 30 && !ok; ++sweep) {
                double off2 = 0.0;
                // rounds are taken two at a time: the matrix work of both happens in registers, then every row of U is read
                // once, rotated by both rounds' rotations and written once (halves the shared-memory traffic of U <- U J)
                for (int r = 0; r < NW - 1; r += 2) {
                    const bool two = (r + 2 <= NW - 1);
                    double c0[NP], x0[NP], y0[NP], c1[NP], x1[NP], y1[NP];
                    unsigned colofs[NW];                                   // U column (in double2 units) of position pos at round r
#pragma unroll
                    for (int i = 0; i < NW; ++i) colofs[i] = (unsigned)(NW * perm[i]) * TPB + tid;
                    jacobi_round_A<NW>(ar, ai, c0, x0, y0, off2);
                    tournament_permute<NW>(ar, ai, perm);
                    if (two) {
                        jacobi_round_A<NW>(ar, ai, c1, x1, y1, off2);
                        tournament_permute<NW>(ar, ai, perm);
                    }
#pragma unroll 2
                    for (int row = 0; row < NW; ++row) {
                        double2 x[NW];
#pragma unroll
                        for (int i = 0; i < NW; ++i) x[i] = Us[colofs[i] + row * TPB];
#pragma unroll
                        for (int i = 0; i < NP; ++i) {                      // round r: positions (2i, 2i+1)
                            const double2 a = x[2 * i], b = x[2 * i + 1];
                            x[2 * i] = make_double2(c0[i] * a.x - (x0[i] * b.x + y0[i] * b.y), c0[i] * a.y - (x0[i] * b.y - y0[i] * b.x));
                            x[2 * i + 1] = make_double2(x0[i] * a.x - y0[i] * a.y + c0[i] * b.x, x0[i] * a.y + y0[i] * a.x + c0[i] * b.y);
                        }
                        if (two) {
#pragma unroll
                            for (int i = 0; i < NP; ++i) {                  // round r+1: new positions (2i, 2i+1) = old (sigma(2i), sigma(2i+1))
                                constexpr int dummy = 0; (void)dummy;
                                const int pa = sigma<NW>(2 * i), pb = sigma<NW>(2 * i + 1);
                                const double2 a = x[pa], b = x[pb];
                                x[pa] = make_double2(c1[i] * a.x - (x1[i] * b.x + y1[i] * b.y), c1[i] * a.y - (x1[i] * b.y - y1[i] * b.x));
                                x[pb] = make_double2(x1[i] * a.x - y1[i] * a.y + c1[i] * b.x, x1[i] * a.y + y1[i] * a.x + c1[i] * b.y);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < NW; ++i) Us[colofs[i] + row * TPB] = x[i];
                    }
                }
                // off-diagonal norm of what is left AFTER this sweep (saves the verification sweep of the classic test)
                double left = 0.0;
#pragma unroll
                for (int q = 0; q < NW; ++q)
#pragma unroll
                    for (int p = 0; p < q; ++p) left += 2.0 * (ar[p][q] * ar[p][q] + ai[p][q] * ai[p][q]);
                ok = (left <= tol) || (off2 <= tol);
            }
        }
        if (!ok) atomicExch(err_flag, 1);
        // ------------------------------------------------------------------ ascending order (zheevd), ties by original index
        double *Eo = E_out + (size_t)ik * NW;
        double2 *Uo = U_out + (size_t)ik * NW2;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
            const double ej = ar[j][j];
            int rank = 0;
#pragma unroll
            for (int i = 0; i < NW; ++i) {
                const double ei = ar[i][i];
                rank += (ei < ej || (ei == ej && perm[i] < perm[j])) ? 1 : 0;
            }
            Eo[rank] = ej;
            const double2 *uc = Us + (size_t)(NW * perm[j]) * TPB + tid;
#pragma unroll
            for (int row = 0; row < NW; ++row) Uo[row + NW * rank] = uc[row * TPB];
        }
    }
}

template <int NW>
cudaError_t launch_eig_t(const DevModel &m, const KSource &ks, double *E_out, double2 *U_out, int *err_flag, int sm_count,
                         cudaStream_t st) {
    constexpr int TPB = EigCfg<NW>::TPB;
    const size_t smem = sizeof(double2) * NW * NW * TPB;
    cudaError_t e = cudaFuncSetAttribute(eig_small_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_sm = (int)((227 * 1024) / (smem + 1024)) < 1 ? 1 : (int)((227 * 1024) / (smem + 1024));
    long long grid = (ks.nk + TPB - 1) / TPB;
    const long long cap = (long long)sm_count * per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    eig_small_kernel<NW><<<(int)grid, TPB, smem, st>>>(m, ks, E_out, U_out, err_flag);
    return cudaGetLastError();
}

// ================================================================================================ Householder + QL
constexpr int kEigSparseMaxR = 48;       // 48 x 16 B x 256 threads = 192 KB of shared memory
template <int NW> struct HqCfg {
    static constexpr int TPB = 256;                        // one CTA per SM (255 registers per thread; 2 x 128 measured 7 % slower)
    static constexpr int NV = (NW - 1) * (NW - 2) / 2;     // parked reflector entries v_s(r), r >= s+2
    static constexpr int SLOTS = NV + (NW - 1);            // + tau_s
    __host__ __device__ static constexpr int voff(int s) { return s * (NW - 2) - s * (s - 1) / 2; }   // first slot of v_s
};

// implicit-shift QL on the real symmetric tridiagonal (d, e) (e[i] couples i and i+1, e[NW-1] is scratch), Z <- Z * rotations.
// EISPACK tql2 recurrence; the same arithmetic as tql_kernel in mepb_tri.cu.
template <int NW>
__device__ __forceinline__ bool tql_regs(double (&d)[NW], double (&e)[NW], double (&z)[NW][NW]) {
    const double eps = 1.1102230246251565e-16;
    int l = 0, iter = 0;
    while (l < NW - 1) {
        int m = NW - 1;                                       // smallest m >= l with a negligible e[m]
#pragma unroll
        for (int i = NW - 2; i >= 0; --i) {
            const bool small = fabs(e[i]) <= eps * (fabs(d[i]) + fabs(d[i + 1]));
            if (i >= l && small) m = i;
        }
        if (m == l) { ++l; iter = 0; continue; }
        if (++iter > 60) return false;
        double dl = d[0], dl1 = d[1], el = e[0], dm = d[NW - 1];
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            if (i == l) { dl = d[i]; el = e[i]; }
            if (i == l + 1) dl1 = d[i];
            if (i == m) dm = d[i];
        }
        double g = (dl1 - dl) / (2.0 * el);
        double r = sqrt(fma(g, g, 1.0));
        g = dm - dl + el / (g + copysign(r, g));
        double s = 1.0, c = 1.0, p = 0.0;
        bool brk = false;
#pragma unroll
        for (int i = NW - 2; i >= 0; --i) {
            if (i < m && i >= l && !brk) {
                const double f = s * e[i], b = c * e[i];
                const double hh = fma(f, f, g * g);
                if (!(hh > 1e-290)) {                         // r == 0: deflate here
                    e[i + 1] = 0.0; d[i + 1] -= p; brk = true;
                } else {
                    const double rr = fast_rsqrt3(hh);
                    e[i + 1] = hh * rr;
                    s = f * rr; c = g * rr;
                    g = d[i + 1] - p;
                    r = fma(d[i] - g, s, 2.0 * c * b);
                    p = s * r;
                    d[i + 1] = g + p;
                    g = fma(c, r, -b);
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        const double zf = z[k][i + 1], zi = z[k][i];
                        z[k][i + 1] = fma(s, zi, c * zf);
                        z[k][i] = fma(c, zi, -(s * zf));
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            if (i == m) e[i] = 0.0;
            if (i == l && !brk) { d[i] -= p; e[i] = g; }
        }
    }
    return true;
}

template <int NW, bool SPARSE>
__global__ void __launch_bounds__(HqCfg<NW>::TPB, 1)
eig_small_hq_kernel(const DevModel m, const KSource ks, double *__restrict__ E_out, double2 *__restrict__ U_out,
                    int *__restrict__ err_flag) {
    using Cfg = HqCfg<NW>;
    constexpr int TPB = Cfg::TPB, NW2 = NW * NW;
    extern __shared__ __align__(16) double2 park[];          // [SLOTS][TPB]
    const int tid = threadIdx.x;
    const int nR = m.nR;
    const bool has_tau = (m.tau != nullptr);
    double2 *pk = park + tid;

    for (long long ik = (long long)blockIdx.x * TPB + tid; ik < ks.nk; ik += (long long)gridDim.x * TPB) {
        double d[NW], e[NW];
        bool ok = true;
        {
            // lower triangle L(r,c), r >= c, from the upper one the Fourier sum fills (zheevd 'U': only it is referenced)
            double lr[NW][NW], li[NW][NW];
            {
                double ar[NW][NW], ai[NW][NW];
                fourier_upper<NW>(m, ks, ik, nR, has_tau, ar, ai, SPARSE ? pk : nullptr, TPB);   // phases share the park
#pragma unroll
                for (int r = 0; r < NW; ++r)
#pragma unroll
                    for (int c = 0; c <= r; ++c) { lr[r][c] = ar[c][r]; li[r][c] = (r == c) ? 0.0 : -ai[c][r]; }
            }
            // -------------------------------------------------------------- zhetd2('L'): T = Q^H A Q, Q = H_0 ... H_{n-2}
#pragma unroll
            for (int s = 0; s < NW - 1; ++s) {
                const double alr = lr[s + 1][s], ali = li[s + 1][s];
                double xn = 0.0;
#pragma unroll
                for (int r = s + 2; r < NW; ++r) xn = fma(lr[r][s], lr[r][s], fma(li[r][s], li[r][s], xn));
                const bool ident = (xn == 0.0 && ali == 0.0);                    // zlarfg: H = 1
                const double nrm = sqrt(fma(alr, alr, fma(ali, ali, xn)));
                const double beta = ident ? alr : -copysign(nrm, alr);
                const double rb = ident ? 0.0 : 1.0 / beta;
                const double tr = (beta - alr) * rb, ti = -ali * rb;             // tau (0 when ident)
                const double sr0 = alr - beta, den = fma(sr0, sr0, ali * ali);
                const double rden = ident ? 0.0 : 1.0 / den;
                const double scr = sr0 * rden, sci = -ali * rden;                // 1 / (alpha - beta)
                double vr[NW], vi[NW];
                vr[s + 1] = 1.0; vi[s + 1] = 0.0;
#pragma unroll
                for (int r = s + 2; r < NW; ++r) {
                    vr[r] = fma(lr[r][s], scr, -(li[r][s] * sci));
                    vi[r] = fma(lr[r][s], sci, li[r][s] * scr);
                }
                d[s] = lr[s][s]; e[s] = beta;
                // p = tau A v
                double pr[NW], pi[NW];
#pragma unroll
                for (int r = s + 1; r < NW; ++r) {
                    double xr = 0.0, xi = 0.0;
#pragma unroll
                    for (int c = s + 1; c < NW; ++c) {
                        const double a_r = (r >= c) ? lr[r][c] : lr[c][r];
                        const double a_i = (r > c) ? li[r][c] : ((r == c) ? 0.0 : -li[c][r]);
                        xr = fma(a_r, vr[c], xr); xi = fma(a_r, vi[c], xi);
                        if (r != c) { xr = fma(-a_i, vi[c], xr); xi = fma(a_i, vr[c], xi); }
                    }
                    pr[r] = fma(tr, xr, -(ti * xi)); pi[r] = fma(tr, xi, ti * xr);
                }
                // alc = -1/2 tau (p^H v);  w = p + alc v
                double dr_ = 0.0, di_ = 0.0;
#pragma unroll
                for (int r = s + 1; r < NW; ++r) {
                    dr_ = fma(pr[r], vr[r], fma(pi[r], vi[r], dr_));
                    di_ = fma(pr[r], vi[r], fma(-pi[r], vr[r], di_));
                }
                const double acr = -0.5 * fma(tr, dr_, -(ti * di_)), aci = -0.5 * fma(tr, di_, ti * dr_);
#pragma unroll
                for (int r = s + 1; r < NW; ++r) {
                    const double wr = fma(acr, vr[r], fma(-aci, vi[r], pr[r]));
                    const double wi = fma(acr, vi[r], fma(aci, vr[r], pi[r]));
                    pr[r] = wr; pi[r] = wi;
                }
                // A -= v w^H + w v^H   (lower triangle; the diagonal stays real)
#pragma unroll
                for (int r = s + 1; r < NW; ++r)
#pragma unroll
                    for (int c = s + 1; c <= r; ++c) {
                        lr[r][c] -= fma(vr[r], pr[c], vi[r] * pi[c]) + fma(pr[r], vr[c], pi[r] * vi[c]);
                        if (r != c) li[r][c] -= fma(vi[r], pr[c], -(vr[r] * pi[c])) + fma(pi[r], vr[c], -(pr[r] * vi[c]));
                    }
                // park the reflector
#pragma unroll
                for (int r = s + 2; r < NW; ++r) pk[(Cfg::voff(s) + (r - s - 2)) * TPB] = make_double2(vr[r], vi[r]);
                pk[(Cfg::NV + s) * TPB] = make_double2(tr, ti);
            }
            d[NW - 1] = lr[NW - 1][NW - 1]; e[NW - 1] = 0.0;
        }
        // ------------------------------------------------------------------ implicit QL, Z in registers
        double z[NW][NW];
#pragma unroll
        for (int r = 0; r < NW; ++r)
#pragma unroll
            for (int c = 0; c < NW; ++c) z[r][c] = (r == c) ? 1.0 : 0.0;
        ok = tql_regs<NW>(d, e, z);
        if (!ok) atomicExch(err_flag, 1);
        // ------------------------------------------------------------------ ascending order (zheevd), ties by index
        double *Eo = E_out + (size_t)ik * NW;
        double2 *Uo = U_out + (size_t)ik * NW2;
        // u_j = H_0 H_1 ... H_{n-2} z_j, two columns per pass (each reflector is read once per pair; two independent chains)
#pragma unroll
        for (int j = 0; j < NW; j += 2) {
            int rank[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const double ej = d[j + t];
                int rk = 0;
#pragma unroll
                for (int i = 0; i < NW; ++i) rk += (d[i] < ej || (d[i] == ej && i < j + t)) ? 1 : 0;
                Eo[rk] = ej; rank[t] = rk;
            }
            double xr[2][NW], xi[2][NW];
#pragma unroll
            for (int r = 0; r < NW; ++r) { xr[0][r] = z[r][j]; xi[0][r] = 0.0; xr[1][r] = z[r][j + 1]; xi[1][r] = 0.0; }
#pragma unroll
            for (int s = NW - 2; s >= 0; --s) {
                const double2 tau = pk[(Cfg::NV + s) * TPB];
                double2 v[NW];
#pragma unroll
                for (int r = s + 2; r < NW; ++r) v[r] = pk[(Cfg::voff(s) + (r - s - 2)) * TPB];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    double dtr = xr[t][s + 1], dti = xi[t][s + 1];                   // v^H x
#pragma unroll
                    for (int r = s + 2; r < NW; ++r) {
                        dtr = fma(v[r].x, xr[t][r], fma(v[r].y, xi[t][r], dtr));
                        dti = fma(v[r].x, xi[t][r], fma(-v[r].y, xr[t][r], dti));
                    }
                    const double tr = fma(tau.x, dtr, -(tau.y * dti)), ti = fma(tau.x, dti, tau.y * dtr);
                    xr[t][s + 1] -= tr; xi[t][s + 1] -= ti;
#pragma unroll
                    for (int r = s + 2; r < NW; ++r) {
                        xr[t][r] -= fma(tr, v[r].x, -(ti * v[r].y));
                        xi[t][r] -= fma(tr, v[r].y, ti * v[r].x);
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
                for (int r = 0; r < NW; r += 2)                                      // 256-bit stores: half the LSU time of 128-bit ones
                    st_global_256(Uo + r + NW * rank[t], make_double2(xr[t][r], xi[t][r]), make_double2(xr[t][r + 1], xi[t][r + 1]));
        }
    }
}

template <int NW, bool SPARSE>
cudaError_t launch_eig_hq_ts(const DevModel &m, const KSource &ks, double *E_out, double2 *U_out, int *err_flag, int sm_count,
                             cudaStream_t st) {
    using Cfg = HqCfg<NW>;
    const int nx = m.ell_nR + (m.tau ? NW : 0);                 // phase slots of the sparse Fourier sum (they share the park)
    const int slots = (SPARSE && nx > Cfg::SLOTS) ? nx : Cfg::SLOTS;
    const size_t smem = sizeof(double2) * slots * Cfg::TPB;
    cudaError_t e = cudaFuncSetAttribute(eig_small_hq_kernel<NW, SPARSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eig_small_hq_kernel<NW, SPARSE>, Cfg::TPB, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    long long grid = (ks.nk + Cfg::TPB - 1) / Cfg::TPB;
    const long long cap = (long long)sm_count * per_sm;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
    eig_small_hq_kernel<NW, SPARSE><<<(int)grid, Cfg::TPB, smem, st>>>(m, ks, E_out, U_out, err_flag);
    return cudaGetLastError();
}
template <int NW>
cudaError_t launch_eig_hq_t(const DevModel &m, const KSource &ks, double *E_out, double2 *U_out, int *err_flag, int sm_count,
                            cudaStream_t st) {
    // sparse Fourier sum: the per-thread phase table must fit beside nothing else (it shares the reflector park)
    const bool sparse = m.ell_R != nullptr && m.ell_nR + NW <= kEigSparseMaxR;
    return sparse ? launch_eig_hq_ts<NW, true>(m, ks, E_out, U_out, err_flag, sm_count, st)
                  : launch_eig_hq_ts<NW, false>(m, ks, E_out, U_out, err_flag, sm_count, st);
}

}  // namespace

bool eig_thread_supported(int nw) { return nw == 4 || nw == 8; }

cudaError_t launch_eig_small(const DevModel &m, const KSource &ks, double *E_out, double2 *U_out, int *err_flag, int sm_count,
                             cudaStream_t st) {
    static const bool jacobi = [] { const char *e = getenv("MEPB_SMALL_EIG"); return e && std::string(e) == "jacobi"; }();
    if (!jacobi) {
        switch (m.nw) {
            case 4: return launch_eig_hq_t<4>(m, ks, E_out, U_out, err_flag, sm_count, st);
            case 8: return launch_eig_hq_t<8>(m, ks, E_out, U_out, err_flag, sm_count, st);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (m.nw) {
        case 4: return launch_eig_t<4>(m, ks, E_out, U_out, err_flag, sm_count, st);
        case 8: return launch_eig_t<8>(m, ks, E_out, U_out, err_flag, sm_count, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mepb
