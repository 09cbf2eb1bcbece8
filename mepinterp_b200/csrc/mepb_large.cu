// mepb_large.cu — the general path for num_wann up to 64 (BASELINE config 5: nw = 64, nR ~ 2000).
//
// Kernels per wave of k-points, intermediates (H(k), dH/dk_a(k): upper triangles, 4 x 64 KB per k at nw = 64) staged through HBM/L2:
//
//  fourier_dense_kernel   FT_R_to_k (wann_interp.f90:151-281) as a dense contraction
//        [k-points x rows] (phases, generated on chip)  x  [rows x 32 matrix elements] (streamed by TMA bulk copies)
//     * LINE = false (explicit k lists): rows = the nR lattice vectors; the phase tile e^{2 pi i k.R} is built in shared memory
//       from three per-axis tables (one sincospi per k and per distinct lattice coordinate, two complex multiplies per (k,R));
//     * LINE = true (Monkhorst-Pack meshes): stage C of the mesh-separable transform — rows = the 3 n1 entries of the mesh line's
//       table G2 built by sep_stage_a/b_kernel (15 instead of 2007 terms per k-point and output at nw = 64, nR = 2007);
//     * the four sums H, d_x H, d_y H, d_z H share the product t = phase * X: 12 FP64 instructions per (k, element, row) on the FMA
//       pipe instead of the 16 a four-output complex GEMM on DMMA would need (DMMA peak == DFMA peak on B200); accumulators live
//       in registers (KJ k-points x 2 elements x 4 sums per thread);
//     * for Hermitian-closed H(R) only the upper triangle n <= m is computed; it is mirrored for callers that want full matrices
//       and left as it is (fp.upper_only) for the mesh pipeline, whose consumers mirror on load — the mirrored 16-byte stores with
//       stride nw were the kernel's bound (half-written sectors: 718 KB of DRAM traffic per k-point for 256 KB of output);
//     * the tight-binding convention (wf_centers) is the diagonal-phase epilogue of SURVEY.md A.1;
//     * run on the operands r_x, r_y, r_z it yields the W-gauge Berry connection A_a and dA_a(b) (wann_interp.f90:288-375).
//  eig_rot_large_kernel   one CTA per k-point: eigenvectors U = Q Z from the Householder + QL pipeline of mepb_tri.cu (reflectors
//     applied to the eigenvectors of the tridiagonal matrix; replaces zheevd, matrix_math.f90:220) — or, with MEPB_LARGE_EIG=jacobi,
//     a cyclic two-sided Jacobi in shared memory kept as the independent cross-check —, V_a = U^dagger (d_a H) U as DMMA zgemms
//     (wann_interp.f90:517-528), occupations, static AHC, MEP (Niu + Kubo) sums, and the pair / photo items for the frequency-sweep
//     kernels: same outputs as the small-num_wann kernels, so everything downstream (sweeps, packing, finalize) is shared.  The same
//     kernel runs its band-sum sections on supplied eigenpairs (out.V_in: mepb_band_sums, position-operator pipeline) and can hand
//     the eigenvectors to HBM (out.U_out).
#include <cstdlib>

#include "mepb_internal.cuh"

namespace mepb {

namespace {

__device__ __forceinline__ void mesh_k(const KSource &ks, long long ik, double &k1, double &k2, double &k3) {
    if (ks.list) { k1 = ks.list[3 * ik]; k2 = ks.list[3 * ik + 1]; k3 = ks.list[3 * ik + 2]; return; }
    const long long lin = ks.lin_begin + ik;                                          // k_space.f90:351-375
    const long long q1 = lin % ks.N1 + 1, q2 = (lin / ks.N1) % ks.N2 + 1, q3 = lin / ((long long)ks.N1 * ks.N2) + 1;
    k1 = (2.0 * (double)q1 - (double)ks.N1 - 1.0) / (2.0 * (double)ks.N1);
    k2 = (2.0 * (double)q2 - (double)ks.N2 - 1.0) / (2.0 * (double)ks.N2);
    k3 = (2.0 * (double)q3 - (double)ks.N3 - 1.0) / (2.0 * (double)ks.N3);
}

// ============================================================================================== Fourier contraction
constexpr int F_TK = kFourierTK;     // k-points per CTA
constexpr int F_TN = kFourierTN;     // matrix elements per CTA
constexpr int F_RC = kFourierRC;     // R points per pipeline stage
constexpr int F_THREADS = 256;

template <int TK>
struct FourierSmem {                 // fixed part; the axis tables (Utot x TK complex) follow
    double2 ph[2][F_RC][TK];         // 2 x 16 KB at TK = 64
    double2 h[2][F_RC][F_TN];        // 2 x  8 KB   (TMA destination)
    double cw[2][F_RC][4];           // c_a(R) = (M.R)_a ; [3] = weight of the row in H itself (line mode)
    uint64_t bar[2];
};

// LINE = false: explicit k list or mesh points, all nR terms per k-point (phases from per-axis tables or sincospi).
// LINE = true : stage C of the mesh-separable transform — one CTA works on 64 k-points of ONE mesh line (q2, q3) and
//               contracts that line's table G2 (sp.nrp rows) with the axis-1 phases.
// KJ = k-points per thread: 4 (CTA tile of 64 k-points, 128 accumulator registers, one CTA per SM: the long nR-term contraction) or
//      2 (32 k-points, two CTAs per SM).  Stage C contracts only 3 n1 rows (three pipeline chunks at n1 = 15): a CTA is mostly
//      prologue (phase lookups) and epilogue (131 KB of stores per 64 k-points), which a second resident CTA overlaps with the
//      first one's multiply-adds (ncu r03c: one CTA per SM left the FP64 pipe at 39 %, 11 % of the warp slots in use); the smaller
//      tile also wastes less of the last tile of a mesh line (N1 = 100: 3 % instead of 22 %).
template <bool LINE, int KJ>
__global__ void __launch_bounds__(F_THREADS, KJ == 2 ? 2 : 1)
fourier_dense_kernel(const FourierPlan fp, const SepPlan sp, const DevModel m, const KSource ks, double2 *__restrict__ Hk_out,
                     double2 *__restrict__ Hka_out) {
    constexpr int F_TK = 16 * KJ;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FourierSmem<F_TK> &S = *reinterpret_cast<FourierSmem<F_TK> *>(smem_raw);
    double2 *tab = reinterpret_cast<double2 *>(smem_raw + sizeof(FourierSmem<F_TK>));      // [Utot][F_TK]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int lk = lane & 7, ln = lane >> 3, wk = warp & 1, wn = warp >> 1;
    const int tile = blockIdx.y;
    const int nw = m.nw, nw2 = nw * nw;
    // k-points of this CTA: ik = kbase + kk for kk < kcnt (line mode: one mesh line; may start before / end after the range)
    long long kbase;
    int q1_0 = 0;
    const double2 *Hsrc;
    int nchunks;
    if (LINE) {
        const int ktl = (ks.N1 + F_TK - 1) / F_TK;
        const int li = blockIdx.x / ktl;
        q1_0 = (blockIdx.x % ktl) * F_TK;
        kbase = (sp.line0 + li) * (long long)ks.N1 + q1_0 - ks.lin_begin;
        Hsrc = sp.G2 + ((size_t)li * fp.ntiles + tile) * sp.nrp * F_TN;
        nchunks = sp.nrp / F_RC;
    } else {
        kbase = (long long)blockIdx.x * F_TK;
        Hsrc = fp.Hp + (size_t)tile * fp.nRp * F_TN;
        nchunks = fp.nRp / F_RC;
    }
    auto kvalid = [&](int kk) -> bool {
        const long long ik = kbase + kk;
        if (LINE && q1_0 + kk >= ks.N1) return false;
        return ik >= 0 && ik < ks.nk;
    };
    auto kclamp = [&](int kk) -> long long {
        long long ik = kbase + kk;
        return ik < 0 ? 0 : (ik >= ks.nk ? ks.nk - 1 : ik);
    };

    if (tid == 0) { mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1); mbar_fence_init(); }
    // ---- per-axis phase tables: tab[u][k] = e^{2 pi i k_a(u) * uval[u]}   (or nothing in direct mode)
    if (!LINE && !fp.direct) {
        for (int idx = tid; idx < fp.utot * F_TK; idx += F_THREADS) {
            const int u = idx / F_TK, kk = idx % F_TK;
            double k1, k2, k3; mesh_k(ks, kclamp(kk), k1, k2, k3);
            const int ax = u < fp.u1 ? 0 : (u < fp.u1 + fp.u2 ? 1 : 2);
            const double kc = ax == 0 ? k1 : (ax == 1 ? k2 : k3);
            double s, c; sincospi(2.0 * kc * fp.uval[u], &s, &c);
            tab[idx] = make_double2(c, s);
        }
    }
    __syncthreads();

    auto issue_chunk = [&](int c) {
        const int b = c & 1;
        if (tid == 0) {
            mbar_expect_tx(&S.bar[b], (uint32_t)(sizeof(double2) * F_RC * F_TN));
            tma_bulk_g2s(&S.h[b][0][0], Hsrc + (size_t)c * F_RC * F_TN, (uint32_t)(sizeof(double2) * F_RC * F_TN), &S.bar[b]);
        }
        for (int idx = tid; idx < F_RC * F_TK; idx += F_THREADS) {
            const int r = idx / F_TK, kk = idx % F_TK;
            const int R = c * F_RC + r;
            double2 p;
            if (LINE) {
                const int q1 = min(q1_0 + kk, ks.N1 - 1);
                p = sp.P1[(size_t)sp.ra[R] * ks.N1 + q1];
            } else if (fp.direct) {
                double k1, k2, k3; mesh_k(ks, kclamp(kk), k1, k2, k3);
                const double x = 2.0 * (k1 * fp.Rf[3 * R] + k2 * fp.Rf[3 * R + 1] + k3 * fp.Rf[3 * R + 2]);
                double s, cc; sincospi(x, &s, &cc);
                p = make_double2(cc, s);
            } else {
                const int packed = fp.ridx[R];
                const double2 a = tab[(packed & 0xff) * F_TK + kk], b2 = tab[((packed >> 8) & 0xff) * F_TK + kk],
                              c2 = tab[((packed >> 16) & 0xff) * F_TK + kk];
                p = cmul(cmul(a, b2), c2);
            }
            S.ph[b][r][kk] = p;
        }
        if (LINE) { if (tid < F_RC * 4) { const int r = tid / 4, a = tid % 4; S.cw[b][r][a] = sp.cw4[4 * (c * F_RC + r) + a]; } }
        else if (tid < F_RC * 3) { const int r = tid / 3, a = tid % 3; S.cw[b][r][a] = fp.cR[3 * (c * F_RC + r) + a]; }
    };

    double2 hk[KJ][2], g[3][KJ][2];
#pragma unroll
    for (int j = 0; j < KJ; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            hk[j][e] = make_double2(0.0, 0.0);
#pragma unroll
            for (int a = 0; a < 3; ++a) g[a][j][e] = make_double2(0.0, 0.0);
        }

    issue_chunk(0);
    for (int c = 0; c < nchunks; ++c) {
        __syncthreads();                       // phases of chunk c visible; everyone is done with buffer (c+1)&1
        if (c + 1 < nchunks) issue_chunk(c + 1);
        const int b = c & 1;
        mbar_wait(&S.bar[b], (uint32_t)((c >> 1) & 1));
#pragma unroll 2
        for (int r = 0; r < F_RC; ++r) {
            const double c0 = S.cw[b][r][0], c1 = S.cw[b][r][1], c2 = S.cw[b][r][2];
            const double ch = LINE ? S.cw[b][r][3] : 1.0;
            double2 p[KJ], h[2];
#pragma unroll
            for (int j = 0; j < KJ; ++j) p[j] = S.ph[b][r][wk * (8 * KJ) + lk + 8 * j];
#pragma unroll
            for (int e = 0; e < 2; ++e) h[e] = S.h[b][r][wn * 8 + ln + 4 * e];
#pragma unroll
            for (int j = 0; j < KJ; ++j)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double tx = fma(p[j].x, h[e].x, -(p[j].y * h[e].y));
                    const double ty = fma(p[j].x, h[e].y, p[j].y * h[e].x);
                    if (LINE) { hk[j][e].x = fma(ch, tx, hk[j][e].x); hk[j][e].y = fma(ch, ty, hk[j][e].y); }
                    else { hk[j][e].x += tx; hk[j][e].y += ty; }
                    g[0][j][e].x = fma(c0, tx, g[0][j][e].x); g[0][j][e].y = fma(c0, ty, g[0][j][e].y);
                    g[1][j][e].x = fma(c1, tx, g[1][j][e].x); g[1][j][e].y = fma(c1, ty, g[1][j][e].y);
                    g[2][j][e].x = fma(c2, tx, g[2][j][e].x); g[2][j][e].y = fma(c2, ty, g[2][j][e].y);
                }
        }
    }
    __syncthreads();
    // ---- epilogue: tight-binding diagonal phases d_n(k) (reuse the phase buffers), * i, mirror, store
    const bool has_tau = (m.tau != nullptr);
    double2 *dn = reinterpret_cast<double2 *>(smem_raw);       // [F_TK][nw]  (<= 64 KB, fits in ph+h)
    if (has_tau) {
        for (int idx = tid; idx < F_TK * nw; idx += F_THREADS) {
            const int kk = idx / nw, n = idx % nw;
            double k1, k2, k3; mesh_k(ks, kclamp(kk), k1, k2, k3);
            double s, c; sincospi(2.0 * (k1 * m.tau[3 * n] + k2 * m.tau[3 * n + 1] + k3 * m.tau[3 * n + 2]), &s, &c);
            dn[idx] = make_double2(c, s);
        }
        __syncthreads();
    }
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const int el = tile * F_TN + wn * 8 + ln + 4 * e;
        if (el >= fp.nel) continue;
        const int packed = fp.elem[el];
        const int n = packed & 0xffff, mm = packed >> 16;
#pragma unroll
        for (int j = 0; j < KJ; ++j) {
            const int kk = wk * (8 * KJ) + lk + 8 * j;
            if (!kvalid(kk)) continue;
            const long long ik = kbase + kk;
            double2 h0 = hk[j][e];
            double2 ga[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) ga[a] = make_double2(-g[a][j][e].y, g[a][j][e].x);          // * i
            if (has_tau) {
                const double2 w = cmulc(dn[kk * nw + mm], dn[kk * nw + n]);                          // d_m conj(d_n)
                h0 = cmul(h0, w);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    ga[a] = cmul(ga[a], w);
                    const double dt = m.tM[3 * mm + a] - m.tM[3 * n + a];
                    ga[a].x = fma(-dt, h0.y, ga[a].x); ga[a].y = fma(dt, h0.x, ga[a].y);             // + i dt H
                }
            }
            Hk_out[(size_t)ik * nw2 + n + nw * mm] = h0;
#pragma unroll
            for (int a = 0; a < 3; ++a) Hka_out[((size_t)ik * 3 + a) * nw2 + n + nw * mm] = ga[a];
            // mirrored element: a 16-byte store with stride nw -> half-written 32-byte sectors, read-modify-write in L2/DRAM
            // (ncu r03c: 718 KB of DRAM traffic per k-point for 256 KB of output); skipped when the consumers mirror on load
            if (fp.herm && n != mm && !fp.upper_only) {
                Hk_out[(size_t)ik * nw2 + mm + nw * n] = cconj(h0);
#pragma unroll
                for (int a = 0; a < 3; ++a) Hka_out[((size_t)ik * 3 + a) * nw2 + mm + nw * n] = cconj(ga[a]);
            }
        }
    }
}

// ---- mesh-separable transform, stages A and B (cheap: O(nR nel) per mesh plane, O(n1 n2 nel) per mesh line) ----
__global__ void phase_table_kernel(double2 *__restrict__ P, const double *__restrict__ u, int n, int N) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n * N) return;
    const int a = idx / N, q = idx % N + 1;
    const double k = (2.0 * (double)q - (double)N - 1.0) / (2.0 * (double)N);                   // k_space.f90:351-375
    double s, c; sincospi(2.0 * k * u[a], &s, &c);
    P[idx] = make_double2(c, s);
}

__global__ void __launch_bounds__(256)
sep_stage_a_kernel(const FourierPlan fp, const SepPlan sp, const int N3) {
    const int n12 = sp.n1 * sp.n2;
    const long long total = (long long)sp.nq3 * n12 * sp.nelp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int el = (int)(idx % sp.nelp);
        const int ab = (int)((idx / sp.nelp) % n12);
        const int qi = (int)(idx / ((long long)sp.nelp * n12));
        const int q3 = min(sp.q3_0 + qi, N3 - 1);
        const double2 *Hel = fp.Hp + (size_t)(el / F_TN) * fp.nRp * F_TN + (el % F_TN);
        double2 g0 = make_double2(0.0, 0.0), g1 = g0;
        for (int c = 0; c < sp.n3; ++c) {
            const int R = sp.map3[ab * sp.n3 + c];
            if (R < 0) continue;
            const double2 t = cmul(sp.P3[(size_t)c * N3 + q3], Hel[(size_t)R * F_TN]);
            g0 = cadd(g0, t);
            g1.x = fma(sp.u3[c], t.x, g1.x); g1.y = fma(sp.u3[c], t.y, g1.y);
        }
        double2 *o = sp.G1 + ((size_t)(qi * n12 + ab) * 2) * sp.nelp + el;
        o[0] = g0; o[sp.nelp] = g1;
    }
}

__global__ void __launch_bounds__(256)
sep_stage_b_kernel(const FourierPlan fp, const SepPlan sp, const int N2) {
    const long long total = (long long)sp.nlines * sp.nrp * sp.nelp;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int el = (int)(idx % sp.nelp);
        const int r = (int)((idx / sp.nelp) % sp.nrp);
        const int li = (int)(idx / ((long long)sp.nelp * sp.nrp));
        double2 acc = make_double2(0.0, 0.0);
        if (r < 3 * sp.n1) {
            const int t = r / sp.n1, a = r % sp.n1;
            const long long line = sp.line0 + li;
            const int q2 = (int)(line % N2);
            const int qi = (int)(line / N2) - sp.q3_0;
            const double2 *src = sp.G1 + ((size_t)(qi * sp.n1 * sp.n2 + a * sp.n2) * 2 + (t == 2 ? 1 : 0)) * sp.nelp + el;
            for (int b = 0; b < sp.n2; ++b) {
                const double2 tt = cmul(sp.P2[(size_t)b * N2 + q2], src[(size_t)b * 2 * sp.nelp]);
                const double w = (t == 1) ? sp.u2[b] : 1.0;
                acc.x = fma(w, tt.x, acc.x); acc.y = fma(w, tt.y, acc.y);
            }
        }
        sp.G2[(((size_t)li * fp.ntiles + el / F_TN) * sp.nrp + r) * F_TN + (el % F_TN)] = acc;
    }
}

// Packed element tiles of a real-space operator for the dense contraction: Hp[t][R][j] = X_R(n, m) for the element (n, m) =
// elem[t TN + j] of tile t (zero padded in R and j), so that one CTA's (tile, RC R points) chunk is one contiguous TMA bulk copy.
// src element (n, m, R) at off + stride (n + nw m + nw^2 R): H(R) has stride 1, the component a of r(R) (3, nw, nw, nR) stride 3, off a.
__global__ void pack_hp_kernel(const double2 *__restrict__ src, int stride, int off, int nw, int nR, int nel, int nRp,
                               const int *__restrict__ elem, double2 *__restrict__ Hp, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % F_TN);
        const long long tr = i / F_TN;
        const int R = (int)(tr % nRp), t = (int)(tr / nRp);
        const int el = t * F_TN + j;
        double2 v = make_double2(0.0, 0.0);
        if (el < nel && R < nR) {
            const int e = elem[el], n = e & 0xffff, mm = e >> 16;
            v = src[(size_t)off + (size_t)stride * ((size_t)n + (size_t)nw * mm + (size_t)nw * nw * R)];
        }
        Hp[i] = v;
    }
}

// (3, nw, nw, nk) user layout from the internal [k][a][idx] layout
__global__ void hka_to_user_kernel(const double2 *__restrict__ in, double2 *__restrict__ out, int nw2, long long nk) {
    const long long total = nk * 3 * nw2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long k = i / (3 * nw2);
        const int rem = (int)(i % (3 * nw2)), idx = rem / 3, a = rem % 3;
        out[i] = in[(k * 3 + a) * nw2 + idx];
    }
}

// ============================================================================================== eigensolve + rotation + sums
constexpr int E_THREADS = 512;
constexpr int E_WARPS = E_THREADS / 32;

__device__ __forceinline__ void rr_pair_rt(int ne, int r, int i, int &p, int &q) {
    const int M = ne - 1;
    int a, b;
    if (i == 0) { a = r; b = M; }
    else { a = r + i; if (a >= M) a -= M; b = r - i; if (b < 0) b += M; }
    p = min(a, b); q = max(a, b);
}

struct LargeSmemLayout {
    int ne, ld, lda;
    size_t oA, oU, oT, oRot, oEs, oFn, oPerm, oMep, oKub, oRed, oCnt, total;
    __host__ __device__ LargeSmemLayout(int nw) {
        ne = nw + (nw & 1); ld = ne + 1; lda = ld;
        // tensor-core rotation (nw % 8 == 0): leading dimensions that make the DMMA fragment loads of zgemm_dmma bank-conflict free.
        // Rows are 16 ld bytes apart; a quarter-warp of an LDS.128 reads 2 rows x 64 B (A operand, row stride must be 64 mod 128:
        // ld = 4 mod 8) or 4 rows x 32 B (B / conjugate-transposed A operand, row stride 32 mod 128: ld = 2 mod 8).  ne + 1 (odd)
        // made every such load a 2-way conflict (profiles/r02b: 30 % of the wavefronts).  Buffer A carries the A operand (d_aH),
        // buffers U and T the other two.
        if ((nw & 7) == 0) { ld = ne + 2; lda = ne + 4; }
        size_t o = 0;
        oA = o; o += sizeof(double2) * ne * lda;
        oU = o; o += sizeof(double2) * ne * ld;
        oT = o; o += sizeof(double2) * ne * ld;
        oRot = o; o += sizeof(double) * 8 * (ne / 2);          // c, sx, sy, new a, new d, bb, p, q
        oEs = o; o += sizeof(double) * ne;
        oFn = o; o += sizeof(double) * ne;
        oMep = o; o += sizeof(double) * (18 * nw);             // niu_ic (9 nw), niu_lc (9 nw)
        oKub = o; o += sizeof(double) * 18 * E_WARPS;          // per-warp kubo ic, lc
        oRed = o; o += sizeof(double) * 16 * E_WARPS;          // block reductions
        oPerm = o; o += sizeof(int) * ne;
        oCnt = o; o += sizeof(int) * (E_WARPS + 2);
        total = (o + 15) / 16 * 16;
    }
};

// C = op(Am) * Bm for nw x nw complex matrices held row-major in shared memory (leading dimension LD), nw % 8 == 0, on the
// FP64 tensor cores: 8 x 8 output tiles, mma.sync.m8n8k4.f64, four real MMAs per complex k-step (re re, -im im, re im, im re).
// A warp owns a tile row and up to four tile columns: its A fragment (one LDS.128 per lane and k-step) is reused for the
// four B fragments.  The 2 x 2 register-blocked FMA loops this replaces issue one LDS.128 per four DFMA and are bound by the
// shared-memory pipe (20 TF); here it is one LDS.128 per 2.5 DMMA = 640 FMA.
// CONJT: op(Am) = Am^dagger (element (n, p) = conj(Am[p LD + n])).  Output: Cs row-major in shared memory, or Cg column-major
// (n + nw j) in global memory.
// LOWCONJ (with CONJT = false): Am holds a Hermitian matrix as its upper triangle copied raw to (p, q) AND (q, p): the strictly
// lower elements are conjugated as they are loaded.
template <bool CONJT, bool LOWCONJ = false>
__device__ __forceinline__ void zgemm_dmma(const int nw, const int LDA, const int LD, const double2 *__restrict__ Am,
                                           const double2 *__restrict__ Bm, double2 *__restrict__ Cs, double2 *__restrict__ Cg) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int kq = lane & 3, nq = lane >> 2;
    const int nt = nw >> 3, ngrp = (nt + 3) >> 2;
    for (int item = warp; item < nt * ngrp; item += E_WARPS) {
        const int tr = item / ngrp, tc0 = (item % ngrp) * 4;
        double cr[4][2], ci[4][2];
#pragma unroll
        for (int t = 0; t < 4; ++t) { cr[t][0] = cr[t][1] = ci[t][0] = ci[t][1] = 0.0; }
        for (int ks = 0; ks < (nw >> 2); ++ks) {
            double2 a = CONJT ? Am[(4 * ks + kq) * LDA + 8 * tr + nq] : Am[(8 * tr + nq) * LDA + 4 * ks + kq];
            if (CONJT) a.y = -a.y;
            if (LOWCONJ && (8 * tr + nq) > (4 * ks + kq)) a.y = -a.y;
            const double nay = -a.y;
            double2 b[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) b[t] = (tc0 + t < nt) ? Bm[(4 * ks + kq) * LD + 8 * (tc0 + t) + nq] : make_double2(0.0, 0.0);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(cr[t][0]), "+d"(cr[t][1]) : "d"(a.x), "d"(b[t].x));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(ci[t][0]), "+d"(ci[t][1]) : "d"(a.x), "d"(b[t].y));
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(cr[t][0]), "+d"(cr[t][1]) : "d"(nay), "d"(b[t].y));
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(ci[t][0]), "+d"(ci[t][1]) : "d"(a.y), "d"(b[t].x));
            }
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (tc0 + t < nt) {
                const int row = 8 * tr + nq, col = 8 * (tc0 + t) + 2 * kq;
                if (Cs) { Cs[row * LD + col] = make_double2(cr[t][0], ci[t][0]); Cs[row * LD + col + 1] = make_double2(cr[t][1], ci[t][1]); }
                else { Cg[row + nw * col] = make_double2(cr[t][0], ci[t][0]); Cg[row + nw * (col + 1)] = make_double2(cr[t][1], ci[t][1]); }
            }
        }
    }
}

// NWT: num_wann as a compile-time constant (0: run-time m.nw).  With it the leading dimensions, buffer offsets and trip counts
// fold into immediates: the run-time form spends a fifth of its instructions on address arithmetic and spill reloads (ncu r03v).
template <int NWT>
__global__ void __launch_bounds__(E_THREADS, 1)
eig_rot_large_kernel(const DevModel m, const KSource ks, const Jobs jb, const InterpOut out, const int max_pairs,
                     const double2 *__restrict__ Hk_in, const double2 *__restrict__ Hka_in, double2 *__restrict__ Vscr,
                     const TriBuf tb) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int nw = NWT > 0 ? NWT : m.nw, nw2 = nw * nw;
    const LargeSmemLayout L(nw);
    const int NE = L.ne, LD = L.ld, LDA = L.lda, NP = NE / 2;     // LDA: buffer A; LD: buffers U, T
    double2 *A = reinterpret_cast<double2 *>(smem_raw + L.oA);
    double2 *U = reinterpret_cast<double2 *>(smem_raw + L.oU);
    double2 *T = reinterpret_cast<double2 *>(smem_raw + L.oT);
    double *rot = reinterpret_cast<double *>(smem_raw + L.oRot);
    double *es = reinterpret_cast<double *>(smem_raw + L.oEs);
    double *fn = reinterpret_cast<double *>(smem_raw + L.oFn);
    double *mep = reinterpret_cast<double *>(smem_raw + L.oMep);
    double *kub = reinterpret_cast<double *>(smem_raw + L.oKub);
    double *red = reinterpret_cast<double *>(smem_raw + L.oRed);
    int *perm = reinterpret_cast<int *>(smem_raw + L.oPerm);
    int *cnt = reinterpret_cast<int *>(smem_raw + L.oCnt);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double2 *Vg = Vscr + (size_t)blockIdx.x * 3 * nw2;          // V_ka(a, n, m) at a*nw2 + n + nw*m
    const bool response = (out.block_static != nullptr);
    const bool need_eig = (out.e_out != nullptr) || (out.U_out != nullptr) || response;

    for (int i = tid; i < 18 * nw; i += E_THREADS) mep[i] = 0.0;
    for (int i = tid; i < 18 * E_WARPS; i += E_THREADS) kub[i] = 0.0;
    double acc_nk = 0.0, acc_ndeg = 0.0, acc_ne = 0.0, acc_nit = 0.0, acc_nph = 0.0, acc_ahc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc_ahc[q] = 0.0;
    __syncthreads();

    for (long long ik = blockIdx.x; ik < ks.nk && need_eig; ik += gridDim.x) {
      if (out.V_in) {
        // ------------------------------------------------------------------ supplied eigenpairs (mepb_band_sums, position-operator
        // branch): energies from e_in, velocities copied below; no eigensolve, no rotation
        if (tid < nw) es[tid] = out.e_in[(size_t)ik * nw + tid];
        __syncthreads();
      } else if (tb.ne) {
        // ------------------------------------------------------------------ U = H_0 H_1 ... H_{n-2} Z  (zunmtr 'L','N')
        // eigenpairs of the tridiagonal matrix come from mepb_tri.cu; thread (col, part) owns rows part + 8 rr of one
        // column of Z in registers, reflectors are broadcast from shared memory, no barrier inside the loop
        const int NEt = tb.ne;
        const double2 *Vr = tb.Vr + (size_t)ik * NEt * NEt;
        double2 *taus = reinterpret_cast<double2 *>(rot);
        for (int idx = tid; idx < (nw - 1) * nw; idx += E_THREADS) {
            const int s = idx / nw, i = idx % nw;
            A[s * LDA + i] = Vr[(size_t)s * NEt + i];
        }
        if (tid < nw) {
            es[tid] = tb.E[(size_t)ik * NEt + tid];
            perm[tid] = (int)tb.rank[(size_t)ik * NEt + tid];
            taus[tid] = tb.tau[(size_t)ik * NEt + tid];
        }
        __syncthreads();
        {
            const int col = min(tid >> 3, nw - 1), part = tid & 7;
            const double *Zc = tb.Z + (size_t)ik * NEt * NEt + (size_t)col * NEt;
            double2 z[8];
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) { const int r = part + 8 * rr; z[rr] = make_double2(r < nw ? Zc[r] : 0.0, 0.0); }
            for (int j = nw - 2; j >= 0; --j) {
                const double2 tj = taus[j];
                if (tj.x == 0.0 && tj.y == 0.0) continue;
                double2 v[8];
                double2 w = make_double2(0.0, 0.0);
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int r = part + 8 * rr;
                    v[rr] = make_double2(0.0, 0.0);
                    if (8 * rr + 7 > j && r < nw) { v[rr] = A[j * LDA + r]; cfmac(w, v[rr], z[rr]); }
                }
                w.x += __shfl_xor_sync(0xffffffffu, w.x, 1); w.y += __shfl_xor_sync(0xffffffffu, w.y, 1);
                w.x += __shfl_xor_sync(0xffffffffu, w.x, 2); w.y += __shfl_xor_sync(0xffffffffu, w.y, 2);
                w.x += __shfl_xor_sync(0xffffffffu, w.x, 4); w.y += __shfl_xor_sync(0xffffffffu, w.y, 4);
                const double2 t = cmul(tj, w);
#pragma unroll
                for (int rr = 0; rr < 8; ++rr)
                    if (8 * rr + 7 > j) { z[rr].x -= t.x * v[rr].x - t.y * v[rr].y; z[rr].y -= t.x * v[rr].y + t.y * v[rr].x; }
            }
            if ((tid >> 3) < nw) {
                const int dst = perm[col];
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) { const int r = part + 8 * rr; if (r < nw) T[r * LD + dst] = z[rr]; }
            }
        }
        __syncthreads();
      } else {
        // ------------------------------------------------------------------ load H(k), U = 1
        const double2 *Hk = Hk_in + (size_t)ik * nw2;
        double nrm_l = 0.0;
        for (int idx = tid; idx < NE * NE; idx += E_THREADS) {
            const int r = idx % NE, c = idx / NE;
            double2 v = make_double2(0.0, 0.0);
            if (r < nw && c < nw) {
                // zheevd 'U': only the upper triangle is referenced, the imaginary part of the diagonal is ignored
                if (r < c) v = Hk[r + nw * c];
                else if (r > c) v = cconj(Hk[c + nw * r]);
                else v = make_double2(Hk[r + nw * c].x, 0.0);
                nrm_l += v.x * v.x + v.y * v.y;
            }
            A[r * LDA + c] = v;
            U[r * LD + c] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        }
        nrm_l = warp_sum(nrm_l);
        if (lane == 0) red[warp] = nrm_l;
        __syncthreads();
        double nrm = 0.0;
        for (int w = 0; w < E_WARPS; ++w) nrm += red[w];
        __syncthreads();
        // ------------------------------------------------------------------ cyclic Jacobi
        bool ok = (nrm >= 0.0);
        if (nrm > 0.0) {
            const double tol = 1e-26 * nrm;
            ok = false;
            for (int sweep = 0; sweep < 40 && !ok; ++sweep) {
                double off2 = 0.0;      // meaningful in threads < NP
                for (int r = 0; r < NE - 1; ++r) {
                    if (tid < NP) {
                        int p, q; rr_pair_rt(NE, r, tid, p, q);
                        const double a = A[p * LDA + p].x, d = A[q * LDA + q].x;
                        const double2 b = A[p * LDA + q];
                        const double bb = fma(b.x, b.x, b.y * b.y);
                        double c = 1.0, sx = 0.0, sy = 0.0, tb = 0.0;
                        if (bb > 1e-300) {
                            const double zeta = 0.5 * (d - a), az = fabs(zeta);
                            const double h = sqrt(fma(zeta, zeta, bb));
                            const double y = 2.0 * h * (az + h);
                            const double ry = rsqrt(y);
                            const double sg = zeta < 0.0 ? -1.0 : 1.0;
                            c = (az + h) * ry;
                            const double sr = sg * ry;
                            sx = sr * b.x; sy = sr * b.y;
                            tb = sg * bb * (2.0 * h) * ry * ry;
                            off2 += bb;
                        }
                        double *ro = rot + 8 * tid;
                        ro[0] = c; ro[1] = sx; ro[2] = sy; ro[3] = a - tb; ro[4] = d + tb;
                        reinterpret_cast<int *>(ro + 6)[0] = p; reinterpret_cast<int *>(ro + 6)[1] = q;
                    }
                    __syncthreads();
                    // ---- columns: M <- M J for M = A (all rows) and M = U; one warp per row, lanes over the pairs
                    for (int i = lane; i < NP; i += 32) {
                        const double *ro = rot + 8 * i;
                        const double c = ro[0], sx = ro[1], sy = ro[2];
                        const int p = reinterpret_cast<const int *>(ro + 6)[0], q = reinterpret_cast<const int *>(ro + 6)[1];
                        for (int row = warp; row < NE; row += E_WARPS) {
                            const double2 x = A[row * LDA + p], y = A[row * LDA + q], ux = U[row * LD + p], uy = U[row * LD + q];
                            A[row * LDA + p] = make_double2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
                            A[row * LDA + q] = make_double2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
                            U[row * LD + p] = make_double2(c * ux.x - (sx * uy.x + sy * uy.y), c * ux.y - (sx * uy.y - sy * uy.x));
                            U[row * LD + q] = make_double2(sx * ux.x - sy * ux.y + c * uy.x, sx * ux.y + sy * ux.x + c * uy.y);
                        }
                    }
                    __syncthreads();
                    // ---- rows: A <- J^dagger A ; exact values on the rotated 2x2 block; one warp per pair, lanes over columns
                    for (int i = warp; i < NP; i += E_WARPS) {
                        const double *ro = rot + 8 * i;
                        const double c = ro[0], sx = ro[1], sy = ro[2];
                        const int p = reinterpret_cast<const int *>(ro + 6)[0], q = reinterpret_cast<const int *>(ro + 6)[1];
                        for (int col = lane; col < NE; col += 32) {
                            const double2 x = A[p * LDA + col], y = A[q * LDA + col];
                            double2 op = make_double2(c * x.x - (sx * y.x - sy * y.y), c * x.y - (sx * y.y + sy * y.x));
                            double2 oq = make_double2(sx * x.x + sy * x.y + c * y.x, sx * x.y - sy * x.x + c * y.y);
                            if (col == p) { op = make_double2(ro[3], 0.0); oq = make_double2(0.0, 0.0); }
                            else if (col == q) { op = make_double2(0.0, 0.0); oq = make_double2(ro[4], 0.0); }
                            A[p * LDA + col] = op; A[q * LDA + col] = oq;
                        }
                    }
                    __syncthreads();
                }
                off2 = warp_sum(off2);
                if (lane == 0) red[warp] = off2;
                __syncthreads();
                double tot = 0.0;
                for (int w = 0; w < E_WARPS; ++w) tot += red[w];
                ok = (tot <= tol);
                __syncthreads();
            }
        }
        if (!ok && tid == 0) atomicExch(out.err_flag, 1);
        // ------------------------------------------------------------------ ascending order, occupations
        if (tid < nw) {
            const double ej = A[tid * LDA + tid].x;
            int rank = 0;
            for (int i = 0; i < nw; ++i) {
                const double ei = A[i * LDA + i].x;
                rank += (ei < ej || (ei == ej && i < tid)) ? 1 : 0;
            }
            es[rank] = ej; perm[rank] = tid;
        }
        __syncthreads();
        // sorted eigenvectors: T <- U[:, perm]
        for (int idx = tid; idx < nw * nw; idx += E_THREADS) { const int r = idx / nw, j = idx % nw; T[r * LD + j] = U[r * LD + perm[j]]; }
      }
        double f_loc = 0.0;
        if (tid < nw) {
            const double f = fermi(es[tid], jb);
            fn[tid] = f; f_loc = f;
            if (out.e_out) out.e_out[(size_t)ik * nw + tid] = es[tid];
        }
        if (warp < 2) { f_loc = warp_sum(f_loc); if (lane == 0) acc_ne += f_loc; }     // nw <= 64: warps 0, 1
        __syncthreads();
        if (out.U_out && !out.V_in)                               // eigenvectors in columns, ascending order: U(r, j) at r + nw j
            for (int idx = tid; idx < nw2; idx += E_THREADS) { const int r = idx % nw, j = idx / nw; out.U_out[(size_t)ik * nw2 + idx] = T[r * LD + j]; }
        const bool want_V = response || (out.V_out != nullptr);
        if (want_V && out.V_in) {
            const double2 *Vk = out.V_in + (size_t)ik * 3 * nw2;     // user layout a + 3 (n + nw m)
            for (int idx = tid; idx < 3 * nw2; idx += E_THREADS) { const int a = idx % 3, e = idx / 3; Vg[a * nw2 + e] = Vk[idx]; }
            __threadfence_block();
            __syncthreads();
        } else if (want_V) {
            // -------------------------------------------------------------- V_a = Us^dagger (d_a H) Us   (Us = T)
            // Fast path (Hermitian d_aH(k) stored as its upper triangle, tensor-core rotation): the operand buffer A is filled by
            // asynchronous copies — element (p <= q) raw to A(p, q) and A(q, p), the first product conjugates the lower half on load —
            // and the copy of direction a + 1 runs under the second product of direction a, which does not touch A
            // (ncu r04c: 15 % of this kernel's stall samples were these loads, nothing overlapping them).
            const bool async_A = jb.do_gauge && jb.herm && jb.hka_upper && (nw & 7) == 0;
            auto fill_A_async = [&](int a) {
                const double2 *Ha = Hka_in + ((size_t)ik * 3 + a) * nw2;
                for (int idx = tid; idx < nw2; idx += E_THREADS) {
                    const int p = idx % nw, q = idx / nw;                 // column-major source: row p, column q
                    if (p <= q) {
                        cp_async16(&A[p * LDA + q], Ha + idx);
                        if (p < q) cp_async16(&A[q * LDA + p], Ha + idx);
                    }
                }
                cp_async_commit();
            };
            if (async_A) {
                fill_A_async(0);
                for (int a = 0; a < 3; ++a) {
                    cp_async_wait_all();
                    __syncthreads();
                    zgemm_dmma<false, true>(nw, LDA, LD, A, T, U, nullptr);          // U(buffer) <- d_aH Us
                    __syncthreads();
                    if (a < 2) fill_A_async(a + 1);
                    zgemm_dmma<true>(nw, LD, LD, T, U, nullptr, Vg + a * nw2);       // V_a <- Us^dagger U
                    __syncthreads();
                }
            }
            for (int a = 0; a < 3 && !async_A; ++a) {
                const double2 *Ha = Hka_in + ((size_t)ik * 3 + a) * nw2;
                if (jb.do_gauge) {
                    if (jb.herm && jb.hka_upper) {      // only the upper triangle (row <= column) of the Hermitian d_aH(k) is in HBM
                        for (int idx = tid; idx < nw2; idx += E_THREADS) {             // lower triangle + diagonal of A: coalesced, conflict-free
                            const int mm = idx % nw, n = idx / nw;
                            if (mm <= n) A[n * LDA + mm] = cconj(Ha[idx]);
                        }
                        for (int idx = tid; idx < nw2; idx += E_THREADS) {             // strict upper triangle: coalesced reads, strided stores
                            const int n = idx % nw, mm = idx / nw;
                            if (n < mm) A[n * LDA + mm] = Ha[idx];
                        }
                    } else if (jb.herm) {      // d_aH(k) Hermitian (mirrored exactly by the Fourier kernel): A(n, m) = conj Ha(m, n) makes the
                        //                 transposing copy coalesced in global AND conflict-free in shared memory
                        for (int idx = tid; idx < nw2; idx += E_THREADS) { const int mm = idx % nw, n = idx / nw; A[n * LDA + mm] = cconj(Ha[idx]); }
                    } else {
                        for (int idx = tid; idx < nw2; idx += E_THREADS) { const int n = idx % nw, mm = idx / nw; A[n * LDA + mm] = Ha[idx]; }
                    }
                    __syncthreads();
                    if ((nw & 7) == 0) {                            // tensor-core path: U(buffer) <- Ha Us ; V <- Us^dagger U
                        zgemm_dmma<false>(nw, LDA, LD, A, T, U, nullptr);
                        __syncthreads();
                        zgemm_dmma<true>(nw, LD, LD, T, U, nullptr, Vg + a * nw2);
                        __syncthreads();
                        continue;
                    }
                    // U(buffer) <- Ha * Us     two rows x two columns per thread
                    const int njb = (nw + 1) / 2;
                    for (int t = tid; t < njb * njb; t += E_THREADS) {
                        const int p0 = 2 * (t / njb), j0 = t % njb;
                        const int p1 = min(p0 + 1, nw - 1), j1 = min(j0 + njb, nw - 1);
                        double2 s00 = make_double2(0, 0), s01 = s00, s10 = s00, s11 = s00;
                        for (int q = 0; q < nw; ++q) {
                            const double2 h0 = A[p0 * LDA + q], h1 = A[p1 * LDA + q], u0 = T[q * LD + j0], u1 = T[q * LD + j1];
                            cfma(s00, h0, u0); cfma(s01, h0, u1); cfma(s10, h1, u0); cfma(s11, h1, u1);
                        }
                        U[p0 * LD + j0] = s00; U[p0 * LD + j1] = s01; U[p1 * LD + j0] = s10; U[p1 * LD + j1] = s11;
                    }
                    __syncthreads();
                    // V(n, j) = sum_p conj(Us(p, n)) (Ha Us)(p, j)
                    for (int t = tid; t < njb * njb; t += E_THREADS) {
                        const int n0 = 2 * (t / njb), j0 = t % njb;
                        const int n1 = min(n0 + 1, nw - 1), j1 = min(j0 + njb, nw - 1);
                        double2 s00 = make_double2(0, 0), s01 = s00, s10 = s00, s11 = s00;
                        for (int p = 0; p < nw; ++p) {
                            const double2 c0 = T[p * LD + n0], c1 = T[p * LD + n1], w0 = U[p * LD + j0], w1 = U[p * LD + j1];
                            cfmac(s00, c0, w0); cfmac(s01, c0, w1); cfmac(s10, c1, w0); cfmac(s11, c1, w1);
                        }
                        Vg[a * nw2 + n0 + nw * j0] = s00; Vg[a * nw2 + n0 + nw * j1] = s01;
                        Vg[a * nw2 + n1 + nw * j0] = s10; Vg[a * nw2 + n1 + nw * j1] = s11;
                    }
                    __syncthreads();
                } else {
                    for (int idx = tid; idx < nw2; idx += E_THREADS) Vg[a * nw2 + idx] = Ha[idx];       // W gauge
                }
            }
            __threadfence_block();
            __syncthreads();
            if (out.V_out)
                for (int idx = tid; idx < 3 * nw2; idx += E_THREADS) {
                    const int a = idx % 3, e = idx / 3;
                    out.V_out[(size_t)ik * 3 * nw2 + idx] = Vg[a * nw2 + e];
                }
        }
        if (!response) { __syncthreads(); continue; }
        if (tid == 0) acc_nk += 1.0;
        auto Vel = [&](int a, int n, int mm) -> double2 { return Vg[a * nw2 + n + nw * mm]; };
        // ------------------------------------------------------------------ pair sums, static AHC, AHC items
        {
            const int NPAIR = nw * (nw - 1) / 2;
            const int istride = jb.herm ? kAhcItemH : kAhcItemG;
            int n_items_before = 0;
            for (int p0 = 0; p0 < NPAIR; p0 += E_THREADS) {
                const int pidx = p0 + tid;
                bool emit = false;
                double item[kAhcItemG];
                if (pidx < NPAIR) {
                    int mm = (int)((1.0 + sqrt(1.0 + 8.0 * (double)pidx)) * 0.5);
                    while (mm * (mm - 1) / 2 > pidx) --mm;
                    while ((mm + 1) * mm / 2 <= pidx) ++mm;
                    const int n = pidx - mm * (mm - 1) / 2;
                    const double dE = es[mm] - es[n];
                    const double f_n = fn[n], f_m = fn[mm], fdm = f_m - f_n;
                    if (dE == 0.0) acc_ndeg += 1.0;
                    if (jb.do_ahc) {
                        double2 vnm[3], vmn[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) { vnm[a] = Vel(a, n, mm); vmn[a] = Vel(a, mm, n); }
                        double2 P[9];
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int i = 0; i < 3; ++i) P[i + 3 * j] = cmul(vnm[i], vmn[j]);
                        if (f_n != 0.0 || f_m != 0.0 || dE == 0.0) {                    // kubo.f90:58-73 (unguarded)
                            const double inv2 = 2.0 / (dE * dE);
#pragma unroll
                            for (int j = 0; j < 3; ++j)
#pragma unroll
                                for (int i = 0; i < 3; ++i) acc_ahc[i + 3 * j] += inv2 * (f_n * P[i + 3 * j].y + f_m * P[j + 3 * i].y);
                        }
                        if (fdm != 0.0 || dE == 0.0) {
                            const double wa = fdm / dE;
                            item[0] = fma(dE, dE, jb.eta * jb.eta);
                            const int ai[3] = {0, 0, 1}, aj[3] = {1, 2, 2};
                            const int si[6] = {0, 1, 2, 0, 0, 1}, sj[6] = {0, 1, 2, 1, 2, 2};
                            if (jb.herm) {
#pragma unroll
                                for (int q = 0; q < 3; ++q) item[1 + q] = fdm * (P[ai[q] + 3 * aj[q]].y - P[aj[q] + 3 * ai[q]].y);
#pragma unroll
                                for (int q = 0; q < 6; ++q) item[4 + q] = wa * (P[si[q] + 3 * sj[q]].x + P[sj[q] + 3 * si[q]].x);
                            } else {
#pragma unroll
                                for (int q = 0; q < 3; ++q) {
                                    item[1 + q] = fdm * (P[ai[q] + 3 * aj[q]].x - P[aj[q] + 3 * ai[q]].x);
                                    item[4 + q] = fdm * (P[ai[q] + 3 * aj[q]].y - P[aj[q] + 3 * ai[q]].y);
                                }
#pragma unroll
                                for (int q = 0; q < 6; ++q) {
                                    item[7 + q] = wa * (P[si[q] + 3 * sj[q]].x + P[sj[q] + 3 * si[q]].x);
                                    item[13 + q] = wa * (P[si[q] + 3 * sj[q]].y + P[sj[q] + 3 * si[q]].y);
                                }
                                item[19] = 0.0;
                            }
                            emit = (out.items_ahc != nullptr);
                        }
                    }
                }
                // deterministic compaction: warp ballots + prefix over the warps of the block
                const unsigned bal = __ballot_sync(0xffffffffu, emit);
                if (lane == 0) cnt[warp] = __popc(bal);
                __syncthreads();
                int before = n_items_before, total = 0;
                for (int w = 0; w < E_WARPS; ++w) { const int cw = cnt[w]; if (w < warp) before += cw; total += cw; }
                if (emit) {
                    const int slot = before + __popc(bal & ((1u << lane) - 1u));
                    double2 *dst = reinterpret_cast<double2 *>(out.items_ahc + ((size_t)ik * max_pairs + slot) * istride);
                    const int nd2 = istride / 2;              // (a rolled loop on purpose: unrolled stores cost this kernel 224 more bytes
                    for (int j = 0; j < nd2; ++j) dst[j] = make_double2(item[2 * j], item[2 * j + 1]);   // of spills and 10 % of its time, r04a)
                }
                n_items_before += total;
                __syncthreads();
            }
            if (out.cnt_ahc && tid == 0) { out.cnt_ahc[ik] = n_items_before; acc_nit += (double)n_items_before; }
        }
        // ------------------------------------------------------------------ MEP sums (mep_niu.f90, kubo_mep.f90)
        if (jb.do_mep || jb.do_kubo) {
            const int n0_end = jb.do_kubo ? nw : min(jb.nval, nw);
            for (int n0 = warp; n0 < n0_end; n0 += E_WARPS) {
                double lc_g[9], lc_u[9], ic_g[9], ic_u[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) { lc_g[q] = 0; lc_u[q] = 0; ic_g[q] = 0; ic_u[q] = 0; }
                double2 v00[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) v00[a] = Vel(a, n0, n0);
                for (int n = lane; n < nw; n += 32) {
                    if (n == n0) continue;
                    const double d0n = es[n0] - es[n];
                    double2 v0n[3], vn0[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) { v0n[a] = Vel(a, n0, n); vn0[a] = Vel(a, n, n0); }
                    {
                        const double den = d0n * d0n * d0n;
                        const bool pass = fabs(den) > jb.kubo_tol;
                        const double inv = 1.0 / den;
                        auto re = [](double2 x, double2 y) { return x.x * y.x - x.y * y.y; };
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            double2 Q[3];
#pragma unroll
                            for (int k = 0; k < 3; ++k) Q[k] = cmul(v0n[i], vn0[k]);
                            const double c0 = re(Q[1], v00[2]) - re(Q[2], v00[1]);
                            const double c1 = re(Q[2], v00[0]) - re(Q[0], v00[2]);
                            const double c2 = re(Q[0], v00[1]) - re(Q[1], v00[0]);
                            lc_u[i] += c0 * inv; lc_u[i + 3] += c1 * inv; lc_u[i + 6] += c2 * inv;
                            if (pass) { lc_g[i] += c0 * inv; lc_g[i + 3] += c1 * inv; lc_g[i + 6] += c2 * inv; }
                        }
                    }
                    {
                        double2 Zg[3], Zu[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) { Zg[j] = make_double2(0, 0); Zu[j] = make_double2(0, 0); }
                        const double d2 = d0n * d0n;
                        for (int mm = 0; mm < nw; ++mm) {
                            if (mm == n0) continue;
                            const double den = d2 * (es[n0] - es[mm]);
                            const double inv = 1.0 / den;
                            const bool pass = fabs(den) > jb.kubo_tol;
                            double2 a_[3], b_[3];
#pragma unroll
                            for (int a = 0; a < 3; ++a) { a_[a] = Vel(a, n, mm); b_[a] = Vel(a, mm, n0); }
                            double2 X[3];
                            X[0] = csub(cmul(a_[1], b_[2]), cmul(a_[2], b_[1]));
                            X[1] = csub(cmul(a_[2], b_[0]), cmul(a_[0], b_[2]));
                            X[2] = csub(cmul(a_[0], b_[1]), cmul(a_[1], b_[0]));
#pragma unroll
                            for (int j = 0; j < 3; ++j) {
                                Zu[j].x = fma(X[j].x, inv, Zu[j].x); Zu[j].y = fma(X[j].y, inv, Zu[j].y);
                                if (pass) { Zg[j].x = fma(X[j].x, inv, Zg[j].x); Zg[j].y = fma(X[j].y, inv, Zg[j].y); }
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                ic_u[i + 3 * j] -= (v0n[i].x * Zu[j].x - v0n[i].y * Zu[j].y);
                                ic_g[i + 3 * j] -= (v0n[i].x * Zg[j].x - v0n[i].y * Zg[j].y);
                            }
                    }
                }
#pragma unroll
                for (int q = 0; q < 9; ++q) { lc_g[q] = warp_sum(lc_g[q]); lc_u[q] = warp_sum(lc_u[q]); ic_g[q] = warp_sum(ic_g[q]); ic_u[q] = warp_sum(ic_u[q]); }
                if (lane == 0) {
                    if (jb.do_mep && n0 < jb.nval)
#pragma unroll
                        for (int q = 0; q < 9; ++q) { mep[9 * n0 + q] += ic_g[q]; mep[9 * nw + 9 * n0 + q] += lc_g[q]; }
                    if (jb.do_kubo) {
                        const double f0 = fn[n0];
#pragma unroll
                        for (int q = 0; q < 9; ++q) { kub[18 * warp + q] += f0 * ic_u[q]; kub[18 * warp + 9 + q] += f0 * lc_u[q]; }
                    }
                }
            }
        }
        // ------------------------------------------------------------------ photo items (photo.f90:44-81)
        if (jb.do_opt && out.items_photo) {
            const int NORD = nw * (nw - 1);
            int n_before = 0;
            for (int p0 = 0; p0 < NORD; p0 += E_THREADS) {
                const int pidx = p0 + tid;
                bool emit = false;
                int n = 0, l = 0;
                double df = 0.0;
                if (pidx < NORD) {
                    n = pidx / (nw - 1); l = pidx % (nw - 1); if (l >= n) ++l;
                    df = fn[l] - fn[n];
                    emit = (df != 0.0);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, emit);
                if (lane == 0) cnt[warp] = __popc(bal);
                __syncthreads();
                int before = n_before, total = 0;
                for (int w = 0; w < E_WARPS; ++w) { const int cw = cnt[w]; if (w < warp) before += cw; total += cw; }
                if (emit) {
                    double2 W[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) W[q] = make_double2(0, 0);
                    for (int mm = 0; mm < nw; ++mm) {
                        const double2 rec = crcp(make_double2(es[n] - es[mm], -jb.eta));
                        double2 qc[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) qc[c] = cmul(Vel(c, mm, n), rec);
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            const double2 vb = Vel(b, l, mm);
#pragma unroll
                            for (int c = 0; c < 3; ++c) cfma(W[b + 3 * c], vb, qc[c]);
                        }
                    }
                    const int slot = before + __popc(bal & ((1u << lane) - 1u));
                    double2 *dst = reinterpret_cast<double2 *>(out.items_photo + ((size_t)ik * NORD + slot) * kPhotoItem);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double2 va = cscale(df, Vel(a, n, l));
#pragma unroll
                        for (int q = 0; q < 9; ++q) dst[a + 3 * q] = cmul(va, W[q]);
                    }
                    dst[27] = make_double2(es[n] - es[l], 0.0);
                    dst[28] = make_double2(0.0, 0.0);                           // pad, written: keeps every 32-byte sector whole
                }
                n_before += total;
                __syncthreads();
            }
            if (tid == 0) { out.cnt_photo[ik] = n_before; acc_nph += (double)n_before; }
        }
        __syncthreads();
    }
    // ---------------------------------------------------------------------- block reduction of the static sums
    if (!response) return;
    __syncthreads();
    const int NS = static_size(nw);
    double *o = out.block_static + (size_t)blockIdx.x * NS;
    double vals[14] = {acc_nk, acc_ndeg, acc_ne, acc_nit, acc_nph, acc_ahc[0], acc_ahc[1], acc_ahc[2], acc_ahc[3], acc_ahc[4],
                       acc_ahc[5], acc_ahc[6], acc_ahc[7], acc_ahc[8]};
#pragma unroll
    for (int q = 0; q < 14; ++q) { const double s = warp_sum(vals[q]); if (lane == 0) red[16 * warp + q] = s; }
    __syncthreads();
    if (tid < 14) {
        double s = 0.0;
        for (int w = 0; w < E_WARPS; ++w) s += red[16 * w + tid];
        o[tid] = s;                                    // S_NK .. S_AHC+8 are the first 14 slots
    }
    for (int i = tid; i < 18; i += E_THREADS) {        // kubo ic (9), lc (9)
        double s = 0.0;
        for (int w = 0; w < E_WARPS; ++w) s += kub[18 * w + i];
        o[S_KIC + i] = s;
    }
    for (int i = tid; i < 18 * nw; i += E_THREADS) o[S_NIU + i] = mep[i];
}

}  // namespace

// ================================================================================================= launchers
template <int TK>
static size_t fourier_smem_t(const FourierPlan &fp, int nw, bool has_tau) {
    size_t s = sizeof(FourierSmem<TK>) + sizeof(double2) * (size_t)(fp.direct ? 0 : fp.utot) * TK;
    const size_t epi = has_tau ? sizeof(double2) * (size_t)TK * nw : 0;        // d_n(k) table of the epilogue reuses the buffers
    return s > epi ? s : epi;
}
size_t fourier_dense_smem(const FourierPlan &fp, int nw, bool has_tau) { return fourier_smem_t<F_TK>(fp, nw, has_tau); }

cudaError_t launch_fourier_dense(const FourierPlan &fp, const DevModel &m, const KSource &ks, double2 *Hk, double2 *Hka,
                                 cudaStream_t st) {
    const size_t smem = fourier_dense_smem(fp, m.nw, m.tau != nullptr);
    cudaError_t e = cudaFuncSetAttribute(fourier_dense_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((ks.nk + F_TK - 1) / F_TK), (unsigned)fp.ntiles);
    SepPlan sp; sp.on = 0;
    fourier_dense_kernel<false, 4><<<grid, F_THREADS, smem, st>>>(fp, sp, m, ks, Hk, Hka);
    return cudaGetLastError();
}

cudaError_t launch_pack_hp(const FourierPlan &fp, const double2 *src, int stride, int off, int nw, int nR, double2 *Hp, cudaStream_t st) {
    const long long total = (long long)fp.ntiles * fp.nRp * F_TN;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    pack_hp_kernel<<<grid, 256, 0, st>>>(src, stride, off, nw, nR, fp.nel, fp.nRp, fp.elem, Hp, total);
    return cudaGetLastError();
}

cudaError_t launch_phase_table(double2 *P, const double *u, int n, int N, cudaStream_t st) {
    phase_table_kernel<<<(n * N + 255) / 256, 256, 0, st>>>(P, u, n, N);
    return cudaGetLastError();
}

cudaError_t launch_fourier_mesh(const FourierPlan &fp, SepPlan sp, const DevModel &m, const KSource &ks, double2 *Hk,
                                double2 *Hka, cudaStream_t st) {
    if (ks.nk <= 0) return cudaSuccess;
    const long long lin_end = ks.lin_begin + ks.nk - 1;
    sp.line0 = ks.lin_begin / ks.N1;
    sp.nlines = (int)(lin_end / ks.N1 - sp.line0 + 1);
    sp.q3_0 = (int)(sp.line0 / ks.N2);
    sp.nq3 = (int)((sp.line0 + sp.nlines - 1) / ks.N2) - sp.q3_0 + 1;
    {
        const long long total = (long long)sp.nq3 * sp.n1 * sp.n2 * sp.nelp;
        const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
        sep_stage_a_kernel<<<grid, 256, 0, st>>>(fp, sp, ks.N3);
    }
    {
        const long long total = (long long)sp.nlines * sp.nrp * sp.nelp;
        const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 32);
        sep_stage_b_kernel<<<grid, 256, 0, st>>>(fp, sp, ks.N2);
    }
    FourierPlan f0 = fp; f0.direct = 1; f0.utot = 0;        // no per-axis tables in shared memory
    static const bool wide = getenv("MEPB_STAGEC_TK64") != nullptr;      // testing knob: the 64-k-point tile of round 1
    if (wide) {
        const size_t smem = fourier_smem_t<64>(f0, m.nw, m.tau != nullptr);
        cudaError_t e = cudaFuncSetAttribute(fourier_dense_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int ktl = (ks.N1 + 63) / 64;
        dim3 grid((unsigned)(ktl * sp.nlines), (unsigned)fp.ntiles);
        fourier_dense_kernel<true, 4><<<grid, F_THREADS, smem, st>>>(fp, sp, m, ks, Hk, Hka);
        return cudaGetLastError();
    }
    const size_t smem = fourier_smem_t<32>(f0, m.nw, m.tau != nullptr);
    cudaError_t e = cudaFuncSetAttribute(fourier_dense_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int ktl = (ks.N1 + 31) / 32;
    dim3 grid((unsigned)(ktl * sp.nlines), (unsigned)fp.ntiles);
    fourier_dense_kernel<true, 2><<<grid, F_THREADS, smem, st>>>(fp, sp, m, ks, Hk, Hka);
    return cudaGetLastError();
}

cudaError_t launch_hka_to_user(const double2 *in, double2 *out, int nw2, long long nk, cudaStream_t st) {
    hka_to_user_kernel<<<592, 256, 0, st>>>(in, out, nw2, nk);
    return cudaGetLastError();
}

bool large_supported(int nw) {
    if (nw < 2 || nw > 64) return false;
    return LargeSmemLayout(nw).total <= 227 * 1024;
}
int eig_large_grid(int sm_count, long long nk) { return (int)(nk < sm_count ? (nk < 1 ? 1 : nk) : sm_count); }

cudaError_t launch_eig_rot_large(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs,
                                 const double2 *Hk, const double2 *Hka, double2 *Vscr, const TriBuf &tb, int grid, cudaStream_t st) {
    const size_t smem = LargeSmemLayout(m.nw).total;
    static const bool generic_only = getenv("MEPB_EIGROT_GENERIC") != nullptr;       // A/B knob: run-time num_wann build for every size
    cudaError_t e;
    if (m.nw == 64 && !generic_only) {
        e = cudaFuncSetAttribute(eig_rot_large_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        eig_rot_large_kernel<64><<<grid, E_THREADS, smem, st>>>(m, ks, jobs, out, max_pairs, Hk, Hka, Vscr, tb);
    } else {
        e = cudaFuncSetAttribute(eig_rot_large_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        eig_rot_large_kernel<0><<<grid, E_THREADS, smem, st>>>(m, ks, jobs, out, max_pairs, Hk, Hka, Vscr, tb);
    }
    return cudaGetLastError();
}

}  // namespace mepb
