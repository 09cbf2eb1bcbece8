// mepb_small.cu — fused per-k-point kernel for small num_wann (one warp per k-point, everything on chip).
//
// For one k-point a warp does, without touching HBM in between:
//   1. phases  e^{2 pi i k.R} via sincospi into shared memory                   (wann_interp.f90:259-260, 215-223)
//   2. Fourier sums H(k), dH/dk_a(k)                                             (wann_interp.f90:206-279)
//   3. Hermitian eigensolve by cyclic two-sided Jacobi in shared memory          (replaces zheevd, matrix_math.f90:220)
//   4. gauge rotation V_a = U^dagger (d_a H) U                                    (wann_interp.f90:517-528, 438-440)
//   5. occupations, static AHC, MEP (Niu + Kubo) band sums                        (statistics.f90, kubo.f90:58-73,
//                                                                                  mep_niu.f90, kubo_mep.f90)
//   6. "pair items" for the frequency sweeps (velo_ahc/kubo_ohc, photo_2nd_cond)  -> HBM, consumed by mepb_sweep.cu
// or, in list mode, writes e_k / V_ka / H_k / H_ka for an explicit k list (band mode, parity harness).
//
// Tight-binding FT convention is evaluated in the factorised form (algebraically identical to
// wann_interp.f90:210-239):  H(n,m) = conj(d_n) [sum_R ph_R H_R(n,m)] d_m,  d_n = e^{2 pi i k.tau_n}
//   H_a(n,m) = conj(d_n) [sum_R ph_R i C_a(R) H_R(n,m)] d_m + i (t_a(m) - t_a(n)) H(n,m),   C = M.R, t = M.tau.
#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int kWarpsPerBlock = 8;

template <int NW>
struct SL {  // shared-memory layout per warp, in bytes
    static constexpr int LD = NW + 1;                       // padded row length (complex elements)
    static constexpr int NP = NW / 2;                       // Jacobi pairs per round
    static constexpr int NPAIR = NW * (NW - 1) / 2;
    static constexpr size_t oA = 0;
    static constexpr size_t oU = oA + sizeof(double2) * NW * LD;
    static constexpr size_t oHa = oU + sizeof(double2) * NW * LD;
    static constexpr size_t oDn = oHa + sizeof(double2) * 3 * NW * LD;
    static constexpr size_t oRot = oDn + sizeof(double2) * NW;
    static constexpr size_t oEs = oRot + sizeof(double) * 4 * NP;
    static constexpr size_t oFn = oEs + sizeof(double) * NW;
    static constexpr size_t oMep = oFn + sizeof(double) * NW;            // 18 + 18*NW doubles
    static constexpr size_t oPerm = oMep + sizeof(double) * (18 + 18 * NW);
    static constexpr size_t oPh = (oPerm + sizeof(int) * NW + 15) / 16 * 16;   // nR complex follow
    static size_t bytes(int nR) { return (oPh + sizeof(double2) * nR + 15) / 16 * 16; }
};

__device__ __forceinline__ void rr_pair(int nw, int r, int i, int &p, int &q) {
    // round-robin tournament: round r in [0, nw-1), pair i in [0, nw/2)
    const int M = nw - 1;
    int a, b;
    if (i == 0) { a = r; b = M; }
    else { a = r + i; if (a >= M) a -= M; b = r - i; if (b < 0) b += M; }
    p = min(a, b); q = max(a, b);
}

template <int NW>
__device__ __forceinline__ double2 ldA(const double2 *As, int r, int c) {
    constexpr int LD = SL<NW>::LD;
    if (r <= c) return As[r * LD + c];
    double2 v = As[c * LD + r];
    return make_double2(v.x, -v.y);
}
template <int NW>
__device__ __forceinline__ void stA(double2 *As, int r, int c, double2 v) {
    constexpr int LD = SL<NW>::LD;
    if (r <= c) As[r * LD + c] = v;
    else As[c * LD + r] = make_double2(v.x, -v.y);
}

// Cyclic Jacobi on the canonical upper triangle of As (NW x NW Hermitian), eigenvectors accumulated in Us (columns).
// Returns false if not converged.
template <int NW>
__device__ bool jacobi_warp(double2 *As, double2 *Us, double *rot, int lane) {
    constexpr int LD = SL<NW>::LD, NP = SL<NW>::NP;
    constexpr int NA = NP * (NP - 1);          // off-diagonal block halves
    constexpr int NB = NP;                     // diagonal blocks
    constexpr int NC2 = (NW * NP + 1) / 2;     // U tasks, two per slot
    constexpr int NSLOT = NA + NB + NC2;
    // U = identity, ||A||_F^2
    double nrm = 0.0;
    for (int idx = lane; idx < NW * NW; idx += 32) {
        int r = idx / NW, c = idx % NW;
        Us[r * LD + c] = make_double2(r == c ? 1.0 : 0.0, 0.0);
        if (r <= c) {
            double2 v = As[r * LD + c];
            if (r == c) { v.y = 0.0; As[r * LD + c] = v; nrm += v.x * v.x; }   // zheevd ignores Im of the diagonal
            else nrm += 2.0 * (v.x * v.x + v.y * v.y);
        }
    }
    nrm = warp_sum(nrm);
    __syncwarp();
    if (!(nrm > 0.0)) return nrm == 0.0;   // zero matrix is already diagonal; NaN -> fail
    const double tol = 1e-26 * nrm;
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off2 = 0.0;
        for (int r = 0; r < NW - 1; ++r) {
            // ---- step 1: rotation parameters, one pair per lane
            if (lane < NP) {
                int p, q; rr_pair(NW, r, lane, p, q);
                double a = As[p * LD + p].x, d = As[q * LD + q].x;
                double2 b = As[p * LD + q];
                double bb = fma(b.x, b.x, b.y * b.y);
                double c = 1.0, sx = 0.0, sy = 0.0, tb = 0.0;
                if (bb > 1e-300) {
                    double zeta = 0.5 * (d - a), az = fabs(zeta);
                    double h = sqrt(fma(zeta, zeta, bb));
                    double y = 2.0 * h * (az + h);
                    double ry = rsqrt(y);
                    double sg = zeta < 0.0 ? -1.0 : 1.0;
                    c = (az + h) * ry;
                    double sr = sg * ry;
                    sx = sr * b.x; sy = sr * b.y;
                    tb = sg * bb * (2.0 * h) * ry * ry;
                    off2 += bb;
                }
                rot[4 * lane + 0] = c; rot[4 * lane + 1] = sx; rot[4 * lane + 2] = sy; rot[4 * lane + 3] = tb;
            }
            __syncwarp();
            // ---- step 2: apply J^dagger A J on the upper blocks and U <- U J
            for (int base = 0; base < NSLOT; base += 32) {
                const int slot = base + lane;
                double2 o0 = make_double2(0, 0), o1 = o0, o2 = o0, o3 = o0;
                int kind = -1, i0 = 0, i1 = 0, i2 = 0, i3 = 0;   // destinations
                if (slot < NA) {
                    // off-diagonal block (pi < pj), half h
                    const int blk = slot >> 1, h = slot & 1;
                    // decode blk -> (pi, pj), pi < pj
                    int pi = 0, rem = blk;
                    while (rem >= NP - 1 - pi) { rem -= NP - 1 - pi; ++pi; }
                    const int pj = pi + 1 + rem;
                    int p, q, pp, qq; rr_pair(NW, r, pi, p, q); rr_pair(NW, r, pj, pp, qq);
                    const double ci = rot[4 * pi], six = rot[4 * pi + 1], siy = rot[4 * pi + 2];
                    const double cj = rot[4 * pj], sjx = rot[4 * pj + 1], sjy = rot[4 * pj + 2];
                    const double2 bpp = ldA<NW>(As, p, pp), bpq = ldA<NW>(As, p, qq);
                    const double2 bqp = ldA<NW>(As, q, pp), bqq = ldA<NW>(As, q, qq);
                    double2 tp, tq;
                    if (h == 0) {   // column pp of B J_j : c x - conj(s) y
                        tp = make_double2(cj * bpp.x - (sjx * bpq.x + sjy * bpq.y), cj * bpp.y - (sjx * bpq.y - sjy * bpq.x));
                        tq = make_double2(cj * bqp.x - (sjx * bqq.x + sjy * bqq.y), cj * bqp.y - (sjx * bqq.y - sjy * bqq.x));
                    } else {        // column qq : s x + c y
                        tp = make_double2(sjx * bpp.x - sjy * bpp.y + cj * bpq.x, sjx * bpp.y + sjy * bpp.x + cj * bpq.y);
                        tq = make_double2(sjx * bqp.x - sjy * bqp.y + cj * bqq.x, sjx * bqp.y + sjy * bqp.x + cj * bqq.y);
                    }
                    // rows: out_p = c t_p - s t_q ; out_q = conj(s) t_p + c t_q
                    o0 = make_double2(ci * tp.x - (six * tq.x - siy * tq.y), ci * tp.y - (six * tq.y + siy * tq.x));
                    o1 = make_double2(six * tp.x + siy * tp.y + ci * tq.x, six * tp.y - siy * tp.x + ci * tq.y);
                    kind = 0; i0 = p; i1 = q; i2 = (h == 0) ? pp : qq;
                } else if (slot < NA + NB) {
                    const int pi = slot - NA;
                    int p, q; rr_pair(NW, r, pi, p, q);
                    const double tb = rot[4 * pi + 3];
                    o0.x = As[p * LD + p].x - tb; o1.x = As[q * LD + q].x + tb;
                    kind = 1; i0 = p; i1 = q;
                } else if (slot < NSLOT) {
                    kind = 2;
                    const int t0 = 2 * (slot - NA - NB);
                    {
                        const int row = t0 / NP, pi = t0 % NP;
                        int p, q; rr_pair(NW, r, pi, p, q);
                        const double c = rot[4 * pi], sx = rot[4 * pi + 1], sy = rot[4 * pi + 2];
                        const double2 x = Us[row * LD + p], y = Us[row * LD + q];
                        o0 = make_double2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
                        o1 = make_double2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
                        i0 = row * LD + p; i1 = row * LD + q;
                    }
                    if (t0 + 1 < NW * NP) {
                        const int row = (t0 + 1) / NP, pi = (t0 + 1) % NP;
                        int p, q; rr_pair(NW, r, pi, p, q);
                        const double c = rot[4 * pi], sx = rot[4 * pi + 1], sy = rot[4 * pi + 2];
                        const double2 x = Us[row * LD + p], y = Us[row * LD + q];
                        o2 = make_double2(c * x.x - (sx * y.x + sy * y.y), c * x.y - (sx * y.y - sy * y.x));
                        o3 = make_double2(sx * x.x - sy * x.y + c * y.x, sx * x.y + sy * x.x + c * y.y);
                        i2 = row * LD + p; i3 = row * LD + q;
                    } else { i2 = -1; }
                }
                __syncwarp();
                if (kind == 0) { stA<NW>(As, i0, i2, o0); stA<NW>(As, i1, i2, o1); }
                else if (kind == 1) {
                    As[i0 * LD + i0] = make_double2(o0.x, 0.0); As[i1 * LD + i1] = make_double2(o1.x, 0.0);
                    As[i0 * LD + i1] = make_double2(0.0, 0.0);
                } else if (kind == 2) {
                    Us[i0] = o0; Us[i1] = o1;
                    if (i2 >= 0) { Us[i2] = o2; Us[i3] = o3; }
                }
                __syncwarp();
            }
        }
        off2 = warp_sum(off2);
        if (off2 <= tol) return true;
    }
    return false;
}

template <int NW> constexpr int kMinBlocks = (NW <= 8) ? 3 : 2;   // resident CTAs per SM the register budget is set for

template <int NW>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, kMinBlocks<NW>)
interp_small_kernel(const DevModel m, const KSource ks, const Jobs jb, const InterpOut out, const int max_pairs,
                    const size_t warp_bytes) {
    using L = SL<NW>;
    constexpr int LD = L::LD, NW2 = NW * NW, NPAIR = L::NPAIR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *base = smem_raw + warp * warp_bytes;
    double2 *As = reinterpret_cast<double2 *>(base + L::oA);
    double2 *Us = reinterpret_cast<double2 *>(base + L::oU);
    double2 *Ha = reinterpret_cast<double2 *>(base + L::oHa);      // [a][n*LD + m]
    double2 *dn = reinterpret_cast<double2 *>(base + L::oDn);
    double *rot = reinterpret_cast<double *>(base + L::oRot);
    double *es = reinterpret_cast<double *>(base + L::oEs);
    double *fn = reinterpret_cast<double *>(base + L::oFn);
    double *mep = reinterpret_cast<double *>(base + L::oMep);       // [0..8] kubo_ic [9..17] kubo_lc [18..) niu_ic, niu_lc
    int *perm = reinterpret_cast<int *>(base + L::oPerm);
    double2 *ph = reinterpret_cast<double2 *>(base + L::oPh);
    const int nR = m.nR;
    const bool has_tau = (m.tau != nullptr);
    const bool need_eig = (out.e_out != nullptr) || (out.block_static != nullptr);

    for (int i = lane; i < 18 + 18 * NW; i += 32) mep[i] = 0.0;
    double acc_nk = 0.0, acc_ndeg = 0.0, acc_ne = 0.0, acc_nit = 0.0, acc_nph = 0.0, acc_ahc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc_ahc[q] = 0.0;
    __syncwarp();

    const long long wstride = (long long)gridDim.x * kWarpsPerBlock;
    for (long long ik = (long long)blockIdx.x * kWarpsPerBlock + warp; ik < ks.nk; ik += wstride) {
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
        const bool vin = (out.V_in != nullptr);                   // band sums on supplied (e, V): sections 1-4 are skipped
        // ------------------------------------------------------------------ 1. phases
        for (int R = lane; R < nR && !vin; R += 32) {
            const double x = 2.0 * (k1 * m.Rf[3 * R] + k2 * m.Rf[3 * R + 1] + k3 * m.Rf[3 * R + 2]);
            double s, c; sincospi(x, &s, &c);
            ph[R] = make_double2(c, s);
        }
        if (has_tau && lane < NW && !vin) {
            const double x = 2.0 * (k1 * m.tau[3 * lane] + k2 * m.tau[3 * lane + 1] + k3 * m.tau[3 * lane + 2]);
            double s, c; sincospi(x, &s, &c);
            dn[lane] = make_double2(c, s);
        }
        __syncwarp();
        // ------------------------------------------------------------------ 2. Fourier sums
        for (int e0 = 0; e0 < NW2 && !vin; e0 += 64) {
            const int idx0 = e0 + lane, idx1 = e0 + 32 + lane;       // two matrix elements per lane
            const bool v0 = idx0 < NW2, v1 = idx1 < NW2;
            double2 hk0 = make_double2(0, 0), hk1 = hk0;
            double2 g0[3], g1[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) { g0[a] = make_double2(0, 0); g1[a] = make_double2(0, 0); }
            if (m.sp_ptr) {
                // sparse H(R): each lane walks the non-zero (R, value) list of its two matrix elements
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int idx = e ? idx1 : idx0;
                    if (!(e ? v1 : v0)) continue;
                    double2 hk = make_double2(0, 0), g[3] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
                    const int t1 = __ldg(&m.sp_ptr[idx + 1]);
                    for (int t = __ldg(&m.sp_ptr[idx]); t < t1; ++t) {
                        const int R = __ldg(&m.sp_R[t]);
                        const double2 tt = cmul(ph[R], __ldg(&m.sp_val[t]));
                        const double c0 = __ldg(&m.cR[3 * R]), c1 = __ldg(&m.cR[3 * R + 1]), c2 = __ldg(&m.cR[3 * R + 2]);
                        hk.x += tt.x; hk.y += tt.y;
                        g[0].x = fma(c0, tt.x, g[0].x); g[0].y = fma(c0, tt.y, g[0].y);
                        g[1].x = fma(c1, tt.x, g[1].x); g[1].y = fma(c1, tt.y, g[1].y);
                        g[2].x = fma(c2, tt.x, g[2].x); g[2].y = fma(c2, tt.y, g[2].y);
                    }
                    if (e) { hk1 = hk; g1[0] = g[0]; g1[1] = g[1]; g1[2] = g[2]; }
                    else { hk0 = hk; g0[0] = g[0]; g0[1] = g[1]; g0[2] = g[2]; }
                }
            } else {
            const double2 *Hp = m.H;
#pragma unroll 3
            for (int R = 0; R < nR; ++R) {
                const double2 p = ph[R];
                const double c0 = __ldg(&m.cR[3 * R]), c1 = __ldg(&m.cR[3 * R + 1]), c2 = __ldg(&m.cR[3 * R + 2]);
                if (v0) {
                    const double2 h = __ldg(&Hp[(size_t)R * NW2 + idx0]);
                    const double2 t = cmul(p, h);
                    hk0.x += t.x; hk0.y += t.y;
                    g0[0].x = fma(c0, t.x, g0[0].x); g0[0].y = fma(c0, t.y, g0[0].y);
                    g0[1].x = fma(c1, t.x, g0[1].x); g0[1].y = fma(c1, t.y, g0[1].y);
                    g0[2].x = fma(c2, t.x, g0[2].x); g0[2].y = fma(c2, t.y, g0[2].y);
                }
                if (v1) {
                    const double2 h = __ldg(&Hp[(size_t)R * NW2 + idx1]);
                    const double2 t = cmul(p, h);
                    hk1.x += t.x; hk1.y += t.y;
                    g1[0].x = fma(c0, t.x, g1[0].x); g1[0].y = fma(c0, t.y, g1[0].y);
                    g1[1].x = fma(c1, t.x, g1[1].x); g1[1].y = fma(c1, t.y, g1[1].y);
                    g1[2].x = fma(c2, t.x, g1[2].x); g1[2].y = fma(c2, t.y, g1[2].y);
                }
            }
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int idx = e ? idx1 : idx0;
                if (!(e ? v1 : v0)) continue;
                double2 hk = e ? hk1 : hk0;
                double2 ga[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) { const double2 g = e ? g1[a] : g0[a]; ga[a] = make_double2(-g.y, g.x); }  // * i
                const int n = idx % NW, mm = idx / NW;
                if (has_tau) {
                    const double2 w = cmulc(dn[mm], dn[n]);      // d_m conj(d_n)
                    hk = cmul(hk, w);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        ga[a] = cmul(ga[a], w);
                        const double dt = m.tM[3 * mm + a] - m.tM[3 * n + a];
                        ga[a].x = fma(-dt, hk.y, ga[a].x); ga[a].y = fma(dt, hk.x, ga[a].y);   // + i dt hk
                    }
                }
                As[n * LD + mm] = hk;
#pragma unroll
                for (int a = 0; a < 3; ++a) Ha[a * NW * LD + n * LD + mm] = ga[a];
                if (out.Hk_out) out.Hk_out[(size_t)ik * NW2 + idx] = hk;
                if (out.Hka_out) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) out.Hka_out[((size_t)ik * NW2 + idx) * 3 + a] = ga[a];
                }
            }
        }
        __syncwarp();
        if (!need_eig) continue;
        // ------------------------------------------------------------------ 3. eigensolve
        if (vin) {        // supplied energies and velocities (user layout a + 3 (n + NW m))
            if (lane < NW) { es[lane] = out.e_in[(size_t)ik * NW + lane]; perm[lane] = lane; }
            for (int idx = lane; idx < 3 * NW2; idx += 32) {
                const int a = idx % 3, n = (idx / 3) % NW, mm = idx / (3 * NW);
                Ha[a * NW * LD + n * LD + mm] = out.V_in[(size_t)ik * 3 * NW2 + idx];
            }
        } else if (out.U_in) {   // eigenpairs come from the thread-per-k solver (mepb_eig.cu), already in ascending order
            for (int idx = lane; idx < NW2; idx += 32) Us[(idx % NW) * LD + idx / NW] = out.U_in[(size_t)ik * NW2 + idx];
            if (lane < NW) { es[lane] = out.e_in[(size_t)ik * NW + lane]; perm[lane] = lane; }
        } else {
            const bool ok = jacobi_warp<NW>(As, Us, rot, lane);
            if (!ok && lane == 0) atomicExch(out.err_flag, 1);
            __syncwarp();
            // sort ascending (zheevd order); rank by (value, index)
            if (lane < NW) {
                const double ej = As[lane * LD + lane].x;
                int rank = 0;
                for (int i = 0; i < NW; ++i) {
                    const double ei = As[i * LD + i].x;
                    rank += (ei < ej || (ei == ej && i < lane)) ? 1 : 0;
                }
                es[rank] = ej; perm[rank] = lane;
            }
        }
        __syncwarp();
        if (lane < NW) {
            const double f = fermi(es[lane], jb);
            fn[lane] = f; acc_ne += f;
            if (out.e_out) out.e_out[(size_t)ik * NW + lane] = es[lane];
        }
        // ------------------------------------------------------------------ 4. gauge rotation (in place in Ha)
        if (jb.do_gauge && !vin) {
            for (int t0 = 0; t0 < 3 * NW; t0 += 32) {
                const int t = t0 + lane;
                const bool act = t < 3 * NW;
                const int a = act ? t / NW : 0, mc = act ? t % NW : 0;
                double2 vout[NW];
                if (act) {
                    const int pm = perm[mc];
                    double2 tmp[NW];
#pragma unroll
                    for (int p = 0; p < NW; ++p) tmp[p] = make_double2(0, 0);
                    const double2 *Hrow = Ha + a * NW * LD;
#pragma unroll
                    for (int q = 0; q < NW; ++q) {
                        const double2 u = Us[q * LD + pm];
#pragma unroll
                        for (int p = 0; p < NW; ++p) cfma(tmp[p], Hrow[p * LD + q], u);
                    }
#pragma unroll
                    for (int n = 0; n < NW; ++n) {
                        const int pn = perm[n];
                        double2 acc = make_double2(0, 0);
#pragma unroll
                        for (int p = 0; p < NW; ++p) cfmac(acc, Us[p * LD + pn], tmp[p]);
                        vout[n] = acc;
                    }
                }
                __syncwarp();
                if (act) {
#pragma unroll
                    for (int n = 0; n < NW; ++n) Ha[a * NW * LD + n * LD + mc] = vout[n];
                }
            }
            __syncwarp();
        }
        double2 *Vs = Ha;   // Vs[a*NW*LD + n*LD + m] = V_ka(a,n,m)
        if (out.V_out) {
            for (int idx = lane; idx < 3 * NW2; idx += 32) {
                const int a = idx % 3, n = (idx / 3) % NW, mm = idx / (3 * NW);
                out.V_out[(size_t)ik * 3 * NW2 + idx] = Vs[a * NW * LD + n * LD + mm];
            }
        }
        if (out.block_static == nullptr) { __syncwarp(); continue; }
        if (lane == 0) acc_nk += 1.0;
        // ------------------------------------------------------------------ 5./6. pair sums + AHC items
        {
            const int istride = jb.herm ? kAhcItemH : kAhcItemG;
            int n_items_before = 0;
            for (int p0 = 0; p0 < NPAIR; p0 += 32) {
                const int pidx = p0 + lane;
                const bool act = pidx < NPAIR;
                bool emit = false;
                double item[kAhcItemG];
                if (act) {
                    // decode pidx -> (n, mm), n < mm
                    int mm = 1, rem = pidx;
                    while (rem >= mm) { rem -= mm; ++mm; }
                    const int n = rem;
                    const double dE = es[mm] - es[n];
                    const double f_n = fn[n], f_m = fn[mm], fdm = f_m - f_n;
                    if (dE == 0.0) acc_ndeg += 1.0;
                    if (jb.do_ahc) {
                        double2 vnm[3], vmn[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) { vnm[a] = Vs[a * NW * LD + n * LD + mm]; vmn[a] = Vs[a * NW * LD + mm * LD + n]; }
                        double2 P[9];   // P(i,j) = V(i,n,m) V(j,m,n) at i + 3j
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int i = 0; i < 3; ++i) P[i + 3 * j] = cmul(vnm[i], vmn[j]);
                        // static AHC (kubo.f90:58-73): pair (n,m) contributes (2/dE^2) [f_n Im P(i,j) + f_m Im P(j,i)], unguarded
                        if (f_n != 0.0 || f_m != 0.0 || dE == 0.0) {
                            const double inv2 = 2.0 / (dE * dE);
#pragma unroll
                            for (int j = 0; j < 3; ++j)
#pragma unroll
                                for (int i = 0; i < 3; ++i)
                                    acc_ahc[i + 3 * j] += inv2 * (f_n * P[i + 3 * j].y + f_m * P[j + 3 * i].y);
                        }
                        if (fdm != 0.0 || dE == 0.0) {
                            const double wa = fdm / dE;                     // unguarded like kubo.f90:186
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
                const unsigned bal = __ballot_sync(0xffffffffu, emit);
                if (emit) {
                    const int slot = n_items_before + __popc(bal & ((1u << lane) - 1u));
                    store_ahc_item(out.items_ahc + ((size_t)ik * max_pairs + slot) * istride, item, jb.herm);
                }
                n_items_before += __popc(bal);
            }
            if (out.cnt_ahc && lane == 0) { out.cnt_ahc[ik] = n_items_before; acc_nit += (double)n_items_before; }
        }
        // ------------------------------------------------------------------ MEP sums (mep_niu.f90, kubo_mep.f90)
        if (jb.do_mep || jb.do_kubo) {
            constexpr int GPW = 32 / NW;                 // n0 groups per warp pass
            const int grp = lane / NW, n = lane % NW;
            for (int nb = 0; nb < NW; nb += GPW) {
                const int n0 = nb + grp;
                const bool act = (grp < GPW) && (n0 < NW) && (n != n0);
                double lc_g[9], lc_u[9], ic_g[9], ic_u[9];
#pragma unroll
                for (int q = 0; q < 9; ++q) { lc_g[q] = 0; lc_u[q] = 0; ic_g[q] = 0; ic_u[q] = 0; }
                if (act) {
                    const double d0n = es[n0] - es[n];
                    double2 v0n[3], vn0[3], v00[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        v0n[a] = Vs[a * NW * LD + n0 * LD + n]; vn0[a] = Vs[a * NW * LD + n * LD + n0];
                        v00[a] = Vs[a * NW * LD + n0 * LD + n0];
                    }
                    // ---- local term: sum_kl eps_jkl Re[V(i,n0,n) V(k,n,n0) V(l,n0,n0)] / (E_n0-E_n)^3
                    {
                        const double den = d0n * d0n * d0n;
                        const bool pass = fabs(den) > jb.kubo_tol;
                        const double inv = 1.0 / den;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            double2 Q[3];    // V(i,n0,n) V(k,n,n0)
#pragma unroll
                            for (int k = 0; k < 3; ++k) Q[k] = cmul(v0n[i], vn0[k]);
                            // Re[Q(k) V(l,n0,n0)] antisymmetrised over (k,l)
                            auto re = [](double2 a, double2 b) { return a.x * b.x - a.y * b.y; };
                            const double c0 = re(Q[1], v00[2]) - re(Q[2], v00[1]);
                            const double c1 = re(Q[2], v00[0]) - re(Q[0], v00[2]);
                            const double c2 = re(Q[0], v00[1]) - re(Q[1], v00[0]);
                            lc_u[i] = c0 * inv; lc_u[i + 3] = c1 * inv; lc_u[i + 6] = c2 * inv;
                            if (pass) { lc_g[i] = lc_u[i]; lc_g[i + 3] = lc_u[i + 3]; lc_g[i + 6] = lc_u[i + 6]; }
                        }
                    }
                    // ---- itinerant term: - sum_{m != n0} sum_kl eps_jkl Re[V(i,n0,n) V(k,n,m) V(l,m,n0)] / ((E_n0-E_n)^2 (E_n0-E_m))
                    {
                        double2 Zg[3], Zu[3];
#pragma unroll
                        for (int j = 0; j < 3; ++j) { Zg[j] = make_double2(0, 0); Zu[j] = make_double2(0, 0); }
                        const double d2 = d0n * d0n;
                        for (int mm = 0; mm < NW; ++mm) {
                            if (mm == n0) continue;
                            const double d0m = es[n0] - es[mm];
                            const double den = d2 * d0m;
                            const double inv = 1.0 / den;
                            const bool pass = fabs(den) > jb.kubo_tol;
                            double2 a_[3], b_[3];
#pragma unroll
                            for (int a = 0; a < 3; ++a) { a_[a] = Vs[a * NW * LD + n * LD + mm]; b_[a] = Vs[a * NW * LD + mm * LD + n0]; }
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
                                ic_u[i + 3 * j] = -(v0n[i].x * Zu[j].x - v0n[i].y * Zu[j].y);
                                ic_g[i + 3 * j] = -(v0n[i].x * Zg[j].x - v0n[i].y * Zg[j].y);
                            }
                    }
                }
                // reduce over the NW lanes of the group (NW is a power of two), leader adds into the warp accumulators
#pragma unroll
                for (int q = 0; q < 9; ++q) {
#pragma unroll
                    for (int o = NW / 2; o > 0; o >>= 1) {
                        lc_g[q] += __shfl_xor_sync(0xffffffffu, lc_g[q], o);
                        lc_u[q] += __shfl_xor_sync(0xffffffffu, lc_u[q], o);
                        ic_g[q] += __shfl_xor_sync(0xffffffffu, ic_g[q], o);
                        ic_u[q] += __shfl_xor_sync(0xffffffffu, ic_u[q], o);
                    }
                }
                // kubo: weight f_n0, sum over n0 groups happens via sequential leaders (different smem slots per group pass)
                if (grp < GPW && n == 0 && n0 < NW) {
                    if (jb.do_mep && n0 < jb.nval) {
#pragma unroll
                        for (int q = 0; q < 9; ++q) { mep[18 + 9 * n0 + q] += ic_g[q]; mep[18 + 9 * NW + 9 * n0 + q] += lc_g[q]; }
                    }
                }
                if (jb.do_kubo) {
                    // serialise the groups' additions into the shared 9+9 kubo accumulators
                    for (int g = 0; g < GPW; ++g) {
                        if (grp == g && n == 0 && n0 < NW) {
                            const double f0 = fn[n0];
#pragma unroll
                            for (int q = 0; q < 9; ++q) { mep[q] += f0 * ic_u[q]; mep[9 + q] += f0 * lc_u[q]; }
                        }
                        __syncwarp();
                    }
                }
            }
        }
        // ------------------------------------------------------------------ photo items (photo.f90:44-81)
        if (jb.do_opt && out.items_photo) {
            constexpr int NORD = NW * (NW - 1);
            int n_before = 0;
            for (int p0 = 0; p0 < NORD; p0 += 32) {
                const int pidx = p0 + lane;
                bool emit = false;
                int n = 0, l = 0;
                double df = 0.0;
                if (pidx < NORD) {
                    n = pidx / (NW - 1); l = pidx % (NW - 1); if (l >= n) ++l;
                    df = fn[l] - fn[n];
                    emit = (df != 0.0);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, emit);
                if (emit) {
                    double2 W[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) W[q] = make_double2(0, 0);
                    for (int mm = 0; mm < NW; ++mm) {
                        const double2 rec = crcp(make_double2(es[n] - es[mm], -jb.eta));
                        double2 qc[3];
#pragma unroll
                        for (int c = 0; c < 3; ++c) qc[c] = cmul(Vs[c * NW * LD + mm * LD + n], rec);
#pragma unroll
                        for (int b = 0; b < 3; ++b) {
                            const double2 vb = Vs[b * NW * LD + l * LD + mm];
#pragma unroll
                            for (int c = 0; c < 3; ++c) cfma(W[b + 3 * c], vb, qc[c]);
                        }
                    }
                    const int slot = n_before + __popc(bal & ((1u << lane) - 1u));
                    double2 *dst = reinterpret_cast<double2 *>(out.items_photo + ((size_t)ik * NORD + slot) * kPhotoItem);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        const double2 va = cscale(df, Vs[a * NW * LD + n * LD + l]);
#pragma unroll
                        for (int q = 0; q < 9; ++q) dst[a + 3 * q] = cmul(va, W[q]);
                    }
                    dst[27] = make_double2(es[n] - es[l], 0.0);
                    dst[28] = make_double2(0.0, 0.0);                           // pad, written: keeps every 32-byte sector whole
                }
                n_before += __popc(bal);
            }
            if (lane == 0) { out.cnt_photo[ik] = n_before; acc_nph += (double)n_before; }
        }
        __syncwarp();
    }
    // ---------------------------------------------------------------------- block reduction of the static sums
    if (out.block_static == nullptr) return;
    __syncthreads();
    constexpr int NS = S_NIU + 18 * NW;
    // each warp reduces its lane registers and stages its NS-vector at the start of its own region (A, U, Ha are dead
    // now and contiguous: 10*NW*LD doubles >= NS = 32+18*NW for NW >= 4)
    double *wv = reinterpret_cast<double *>(As);
    {
        const double nk = warp_sum(acc_nk), nd = warp_sum(acc_ndeg), ne = warp_sum(acc_ne);
        const double nit = warp_sum(acc_nit), nph = warp_sum(acc_nph);
#pragma unroll
        for (int q = 0; q < 9; ++q) acc_ahc[q] = warp_sum(acc_ahc[q]);
        __syncwarp();
        if (lane == 0) {
            wv[S_NK] = nk; wv[S_NDEG] = nd; wv[S_NE] = ne; wv[S_NIT] = nit; wv[S_NPH] = nph;
#pragma unroll
            for (int q = 0; q < 9; ++q) wv[S_AHC + q] = acc_ahc[q];
        }
        for (int i = lane; i < 18 + 18 * NW; i += 32) wv[S_KIC + i] = mep[i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < kWarpsPerBlock; ++w) s += reinterpret_cast<const double *>(smem_raw + w * warp_bytes + L::oA)[i];
        out.block_static[(size_t)blockIdx.x * NS + i] = s;
    }
}

template <int NW>
cudaError_t launch_t(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs, int grid,
                     cudaStream_t st) {
    const size_t wb = SL<NW>::bytes(m.nR);
    const size_t smem = wb * kWarpsPerBlock;
    cudaError_t e = cudaFuncSetAttribute(interp_small_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    interp_small_kernel<NW><<<grid, kWarpsPerBlock * 32, smem, st>>>(m, ks, jobs, out, max_pairs, wb);
    return cudaGetLastError();
}

}  // namespace

bool small_supported(int nw, int nR) {
    if (!(nw == 4 || nw == 8 || nw == 16)) return false;
    size_t wb = (nw == 4) ? SL<4>::bytes(nR) : (nw == 8) ? SL<8>::bytes(nR) : SL<16>::bytes(nR);
    return wb * kWarpsPerBlock <= 200 * 1024;
}

int interp_small_grid(int nw, int nR, int sm_count, long long nk) {
    size_t wb = (nw == 4) ? SL<4>::bytes(nR) : (nw == 8) ? SL<8>::bytes(nR) : SL<16>::bytes(nR);
    size_t smem = wb * kWarpsPerBlock + 1024;
    int per_sm = (int)((227 * 1024) / smem);
    const int cap = (nw <= 8) ? 3 : 2;   // kMinBlocks<NW>
    if (per_sm > cap) per_sm = cap;
    if (per_sm < 1) per_sm = 1;
    long long need = (nk + kWarpsPerBlock - 1) / kWarpsPerBlock;
    long long g = (long long)sm_count * per_sm;
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_interp_small(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs,
                                int grid, cudaStream_t st) {
    switch (m.nw) {
        case 4: return launch_t<4>(m, ks, jobs, out, max_pairs, grid, st);
        case 8: return launch_t<8>(m, ks, jobs, out, max_pairs, grid, st);
        case 16: return launch_t<16>(m, ks, jobs, out, max_pairs, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mepb
