// mepb_posop.cu — the Berry-connection / curvature branch of get_wann_interp and the band sums that need it.
//
// Reference (paths relative to /root/reference/src): with a position operator r_R(a,n,m) (`_tb.dat` / `_r.dat`) the hot path also
//   * Fourier-sums A_ka = sum_R ph r_R(a) and Om_kab = sum_R ph i [c_a r_R(b) - c_b r_R(a)]          wann_interp.f90:288-375
//     -> done by the dense Fourier contraction kernel of mepb_large.cu run on the operands r_x, r_y, r_z: its "H(k)" output is
//        A_a and its three derivative outputs are dA_a(b) = sum_R ph i c_b r_R(a), so Om_ab = dA_b(a) - dA_a(b);
//   * rotates them into the Hamiltonian gauge and adds the covariant-derivative terms               wann_interp.f90:465-658
//        Hbar_a = U^+ H_a U,  Abar_a = U^+ A_a U,  D_a(n,m) = Hbar_a(n,m) / (E_m - E_n)  (|dE| > kubo_tol, else 0),
//        A_a <- Abar_a + i D_a,   Om_ab <- U^+ Om_ab U - [D_a, Ax_b] + [D_b, Ax_a] - i [D_a, D_b]
//   * corrects the velocities  V_a(n,m) = Hbar_a(n,m) - i (E_m - E_n) Av_a(n,m),  n /= m            wann_interp.f90:406-444
//     (Ax = Av = Abar: PRB 74 195118 eqs. 27, 31, posop_mode 0;  = Abar + i D: as coded, posop_mode 1 — see include/mepb.h)
//   * and feeds mep_niu_CS / kubo_mep_CS (mep_niu.f90:20-44, kubo_mep.f90:16-47), kubo_opt_tens (kubo.f90:204-405) and the
//     gyrotropic tensors (gyro.f90:74-268).
// posop_gauge_kernel does the second and third step for one k-point per CTA (eigenpairs from the general path's eigensolver) and
// writes V, A and the curvature pseudo-vector of every band to HBM; the ordinary band sums then run on (E, V) through the
// supplied-eigenpair mode of eig_rot_large_kernel; posop_sums_kernel adds the sums that need A and Om.  Only the diagonal of
// Om_ab is ever consumed (CS terms, gyro D), so only that diagonal is formed.
// This branch is outside the five BASELINE configurations; it is written for correctness on the FP64 FMA pipe, not tuned.
#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int G_THREADS = 512;
constexpr int S_THREADS = 256;
constexpr int S_TILE = 128;        // ordered band pairs per shared-memory tile of posop_sums_kernel
constexpr int S_ITEM = 20;         // [0] dE (or dE^2), [1..18] nine complex weights, [19] pad

struct GaugeSmem {
    int ld;
    size_t oU, oX, oT, oE, oOm, total;
    __host__ __device__ GaugeSmem(int nw) {
        ld = nw + 1;
        size_t o = 0;
        oU = o; o += sizeof(double2) * nw * ld;
        oX = o; o += sizeof(double2) * nw * ld;
        oT = o; o += sizeof(double2) * nw * ld;
        oE = o; o += sizeof(double) * (nw + (nw & 1));       // keeps the double2 array behind it 16-byte aligned
        oOm = o; o += sizeof(double2) * 3 * nw;
        total = (o + 15) / 16 * 16;
    }
};

// Xs (row-major, ld) <- U^+ Xs U ; Ts scratch.  All threads of the CTA.
__device__ void rotate_in_smem(int nw, int LD, const double2 *Us, double2 *Xs, double2 *Ts) {
    const int nw2 = nw * nw;
    for (int idx = threadIdx.x; idx < nw2; idx += blockDim.x) {
        const int p = idx / nw, j = idx % nw;
        double2 s = make_double2(0.0, 0.0);
        for (int q = 0; q < nw; ++q) cfma(s, Xs[p * LD + q], Us[q * LD + j]);
        Ts[p * LD + j] = s;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nw2; idx += blockDim.x) {
        const int n = idx / nw, j = idx % nw;
        double2 s = make_double2(0.0, 0.0);
        for (int p = 0; p < nw; ++p) cfmac(s, Us[p * LD + n], Ts[p * LD + j]);
        Xs[n * LD + j] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(G_THREADS, 1)
posop_gauge_kernel(const int nw, const long long nk, const int do_gauge, const int mode, const double kubo_tol, const int has_pos,
                   const double *__restrict__ E, const double2 *__restrict__ Uin, const double2 *__restrict__ Hka,
                   const double2 *__restrict__ Aw0, const double2 *__restrict__ Aw1, const double2 *__restrict__ Aw2,
                   const double2 *__restrict__ dA0, const double2 *__restrict__ dA1, const double2 *__restrict__ dA2,
                   double2 *__restrict__ Vout, double2 *__restrict__ AHout, double2 *__restrict__ OmVout, double2 *__restrict__ scratch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const GaugeSmem L(nw);
    const int LD = L.ld, nw2 = nw * nw, tid = threadIdx.x;
    double2 *Us = reinterpret_cast<double2 *>(smem_raw + L.oU);
    double2 *Xs = reinterpret_cast<double2 *>(smem_raw + L.oX);
    double2 *Ts = reinterpret_cast<double2 *>(smem_raw + L.oT);
    double *es = reinterpret_cast<double *>(smem_raw + L.oE);
    double2 *omb = reinterpret_cast<double2 *>(smem_raw + L.oOm);        // rotated curvature diagonal, pairs (0,1) (0,2) (1,2)
    double2 *G = scratch + (size_t)blockIdx.x * 6 * nw2;                 // [0..2] Hbar_a, [3..5] Abar_a ; element n + nw m
    const double2 *Aw[3] = {Aw0, Aw1, Aw2};
    const double2 *dA[3] = {dA0, dA1, dA2};

    for (long long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        for (int idx = tid; idx < nw2; idx += G_THREADS) { const int r = idx % nw, j = idx / nw; Us[r * LD + j] = Uin[(size_t)ik * nw2 + idx]; }
        if (tid < nw) es[tid] = E[(size_t)ik * nw + tid];
        __syncthreads();
        // ---- Hbar_a, Abar_a (rotate_gauge, wann_interp.f90:517-533 with the eq. 21 rotation of A)
        for (int t = 0; t < 6; ++t) {
            const int a = t % 3;
            const bool isA = t >= 3;
            if (isA && !has_pos) { for (int idx = tid; idx < nw2; idx += G_THREADS) G[(size_t)t * nw2 + idx] = make_double2(0.0, 0.0); continue; }
            const double2 *src = isA ? Aw[a] + (size_t)ik * nw2 : Hka + ((size_t)ik * 3 + a) * nw2;
            if (!do_gauge) { for (int idx = tid; idx < nw2; idx += G_THREADS) G[(size_t)t * nw2 + idx] = src[idx]; continue; }
            for (int idx = tid; idx < nw2; idx += G_THREADS) { const int n = idx % nw, mm = idx / nw; Xs[n * LD + mm] = src[idx]; }
            __syncthreads();
            rotate_in_smem(nw, LD, Us, Xs, Ts);
            for (int idx = tid; idx < nw2; idx += G_THREADS) { const int n = idx % nw, mm = idx / nw; G[(size_t)t * nw2 + idx] = Xs[n * LD + mm]; }
            __syncthreads();
        }
        // ---- diagonal of U^+ Om_ab U for (a,b) = (0,1), (0,2), (1,2);  Om_ab = dA_b(a) - dA_a(b)   (wann_interp.f90:347-370, 536-542)
        for (int pr = 0; pr < 3; ++pr) {
            const int a = pr == 2 ? 1 : 0, b = pr == 0 ? 1 : 2;
            if (!has_pos) { if (tid < nw) omb[pr * nw + tid] = make_double2(0.0, 0.0); continue; }
            const double2 *xb = dA[b] + ((size_t)ik * 3 + a) * nw2, *xa = dA[a] + ((size_t)ik * 3 + b) * nw2;
            for (int idx = tid; idx < nw2; idx += G_THREADS) { const int n = idx % nw, mm = idx / nw; Xs[n * LD + mm] = csub(xb[idx], xa[idx]); }
            __syncthreads();
            if (!do_gauge) { if (tid < nw) omb[pr * nw + tid] = Xs[tid * LD + tid]; __syncthreads(); continue; }
            for (int idx = tid; idx < nw2; idx += G_THREADS) {
                const int p = idx / nw, j = idx % nw;
                double2 s = make_double2(0.0, 0.0);
                for (int q = 0; q < nw; ++q) cfma(s, Xs[p * LD + q], Us[q * LD + j]);
                Ts[p * LD + j] = s;
            }
            __syncthreads();
            if (tid < nw) {
                double2 s = make_double2(0.0, 0.0);
                for (int p = 0; p < nw; ++p) cfmac(s, Us[p * LD + tid], Ts[p * LD + tid]);
                omb[pr * nw + tid] = s;
            }
            __syncthreads();
        }
        __threadfence_block();
        __syncthreads();
        // D_a(n,m) = Hbar_a(n,m) / (E_m - E_n), guarded (get_gauge_covar_deriv, wann_interp.f90:565-617); zero without gauge trafo
        auto Dv = [&](int a, int n, int mm) -> double2 {
            if (!do_gauge || n == mm) return make_double2(0.0, 0.0);
            const double dE = es[mm] - es[n];
            if (!(fabs(dE) > kubo_tol)) return make_double2(0.0, 0.0);
            const double2 h = G[(size_t)a * nw2 + n + nw * mm];
            return make_double2(h.x / dE, h.y / dE);
        };
        auto Ab = [&](int a, int n, int mm) -> double2 { return G[(size_t)(3 + a) * nw2 + n + nw * mm]; };
        auto Ah = [&](int a, int n, int mm) -> double2 { const double2 d = Dv(a, n, mm), x = Ab(a, n, mm); return make_double2(x.x - d.y, x.y + d.x); };
        // ---- curvature diagonal in the H gauge (curv_gaugeTrafo, wann_interp.f90:636-658) -> pseudo-vector per band
        for (int item = tid; item < 3 * nw; item += G_THREADS) {
            const int pr = item / nw, n = item % nw;
            const int a = pr == 2 ? 1 : 0, b = pr == 0 ? 1 : 2;
            double2 c1 = make_double2(0.0, 0.0), c2 = c1, c3 = c1;
            if (do_gauge)
                for (int mm = 0; mm < nw; ++mm) {
                    if (mm == n) continue;
                    const double2 da_nm = Dv(a, n, mm), da_mn = Dv(a, mm, n), db_nm = Dv(b, n, mm), db_mn = Dv(b, mm, n);
                    const double2 xa_nm = mode ? Ah(a, n, mm) : Ab(a, n, mm), xa_mn = mode ? Ah(a, mm, n) : Ab(a, mm, n);
                    const double2 xb_nm = mode ? Ah(b, n, mm) : Ab(b, n, mm), xb_mn = mode ? Ah(b, mm, n) : Ab(b, mm, n);
                    c1 = cadd(c1, csub(cmul(da_nm, xb_mn), cmul(xb_nm, da_mn)));       // [D_a, Ax_b]_nn
                    c2 = cadd(c2, csub(cmul(db_nm, xa_mn), cmul(xa_nm, db_mn)));       // [D_b, Ax_a]_nn
                    c3 = cadd(c3, csub(cmul(da_nm, db_mn), cmul(db_nm, da_mn)));       // [D_a, D_b]_nn
                }
            const double2 o0 = omb[pr * nw + n];
            // Om_ab - c1 + c2 - i c3
            const double2 om = make_double2(o0.x - c1.x + c2.x + c3.y, o0.y - c1.y + c2.y - c3.x);
            // convert_tens_to_vect (matrix_math.f90:511-529): v_c = sum_ab eps_abc Om_ab = 2 Om_ab for (a,b,c) cyclic, Om antisymmetric
            const int c = pr == 0 ? 2 : (pr == 1 ? 1 : 0);
            const double sg = pr == 1 ? -2.0 : 2.0;
            OmVout[((size_t)ik * 3 + c) * nw + n] = make_double2(sg * om.x, sg * om.y);
        }
        // ---- connection in the H gauge (conn_gaugeTrafo :620-633) and velocities (get_velo :406-444; diagonal = Hbar, Q2)
        for (int idx = tid; idx < 3 * nw2; idx += G_THREADS) {
            const int a = idx / nw2, e = idx % nw2, n = e % nw, mm = e / nw;
            const double2 ah = Ah(a, n, mm);
            AHout[((size_t)ik * 3 + a) * nw2 + e] = ah;
            double2 v = G[(size_t)a * nw2 + e];
            if (n != mm && has_pos) {
                const double2 av = mode ? ah : Ab(a, n, mm);
                const double dE = es[mm] - es[n];
                v.x += dE * av.y; v.y -= dE * av.x;                                 // - i dE Av
            }
            Vout[(size_t)ik * 3 * nw2 + a + 3 * e] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Row written per block (doubles): [0, nval) mep CS per band | [nval] kubo CS (this E_F) | opt: n_hw x 36 (herm 9c, aherm 9c) |
// gyro C 18 | gyro D 18 | gyro Dw n_hw x 18.  blockIdx.y selects a tile of HB frequencies.
__global__ void __launch_bounds__(S_THREADS)
posop_sums_kernel(const int nw, const long long nk, const Jobs jb, const int has_pos, const int do_gyro, const int n_hw, const int HB,
                  const double *__restrict__ hw, const double *__restrict__ E, const double2 *__restrict__ V,
                  const double2 *__restrict__ AH, const double2 *__restrict__ OmV, double *__restrict__ rows, const int row_len) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *es = reinterpret_cast<double *>(smem_raw);                    // [nw]
    double *fn = es + nw;                                                 // [nw]
    double *fdd = fn + nw;                                                // [nw]
    double *tile = fdd + nw;                                              // [S_TILE][S_ITEM]  opt items
    double *tileg = tile + S_TILE * S_ITEM;                               // [S_TILE][S_ITEM]  gyro Dw items
    double *red = tileg + S_TILE * S_ITEM;                                // [S_THREADS] reduction scratch
    const int tid = threadIdx.x, nw2 = nw * nw;
    const int NS = S_THREADS / HB, h_loc = tid % HB, slice = tid / HB;
    const int h = blockIdx.y * HB + h_loc;
    const bool hv = h < n_hw;
    const double w = hv ? hw[h] : 0.0;
    const int nval = jb.nval;
    const bool first_tile = (blockIdx.y == 0);
    const bool want_opt = has_pos && jb.do_opt, want_dw = has_pos && do_gyro;

    double acc_opt[36], acc_dw[18];
#pragma unroll
    for (int q = 0; q < 36; ++q) acc_opt[q] = 0.0;
#pragma unroll
    for (int q = 0; q < 18; ++q) acc_dw[q] = 0.0;
    double acc_cs_mep = 0.0, acc_cs_kub = 0.0, acc_c[18], acc_d[18];      // thread n0 = tid (< nw)
#pragma unroll
    for (int q = 0; q < 18; ++q) { acc_c[q] = 0.0; acc_d[q] = 0.0; }

    for (long long ik = blockIdx.x; ik < nk; ik += gridDim.x) {
        const double2 *Vk = V + (size_t)ik * 3 * nw2;                     // V_ka(a,n,m) at a + 3 (n + nw m)
        const double2 *Ak = AH + (size_t)ik * 3 * nw2;                    // A(a,n,m) at a nw2 + n + nw m
        __syncthreads();
        if (tid < nw) {
            const double e = E[(size_t)ik * nw + tid];
            const double f = fermi(e, jb);
            es[tid] = e; fn[tid] = f; fdd[tid] = 1.0 / (2.0 + 2.0 * cosh(f));      // d1_fd_stat_deriv as coded, statistics.f90:186
        }
        __syncthreads();
        if (first_tile && tid < nw) {
            const int n = tid;
            if (has_pos && (jb.do_mep || jb.do_kubo)) {                   // mep_niu_CS / kubo_mep_CS: 0.5 Re sum_c conj(A_c(nn)) om_c(n)
                double s = 0.0;
                for (int c = 0; c < 3; ++c) {
                    const double2 a = Ak[(size_t)c * nw2 + n + nw * n], o = OmV[((size_t)ik * 3 + c) * nw + n];
                    s += a.x * o.x + a.y * o.y;
                }
                s *= 0.5;
                if (jb.do_mep && n < nval) acc_cs_mep += s;
                if (jb.do_kubo) acc_cs_kub += s * fn[n];
            }
            if (do_gyro) {                                                // get_gyro_C / get_gyro_D, gyro.f90:74-150, 200-232
                double2 vd[3], om[3];
                for (int a = 0; a < 3; ++a) vd[a] = Vk[a + 3 * (n + nw * n)];
                if (has_pos) { for (int c = 0; c < 3; ++c) om[c] = OmV[((size_t)ik * 3 + c) * nw + n]; }
                else {                                                    // get_kubo_curv, wann_interp.f90:113-147
                    double curv[9];
                    for (int q = 0; q < 9; ++q) curv[q] = 0.0;
                    for (int np = 0; np < nw; ++np) {
                        if (np == n) continue;
                        const double dEE = (es[np] - es[n]) * (es[np] - es[n]);
                        if (fabs(dEE) < jb.kubo_tol) continue;
                        for (int j = 0; j < 3; ++j)
                            for (int i = 0; i < 3; ++i) {
                                const double2 x = Vk[i + 3 * (n + nw * np)], y = Vk[j + 3 * (np + nw * n)];
                                curv[i + 3 * j] += (x.x * y.y + x.y * y.x) / dEE;
                            }
                    }
                    om[0] = make_double2(-2.0 * (curv[1 + 3 * 2] - curv[2 + 3 * 1]), 0.0);
                    om[1] = make_double2(-2.0 * (curv[2 + 3 * 0] - curv[0 + 3 * 2]), 0.0);
                    om[2] = make_double2(-2.0 * (curv[0 + 3 * 1] - curv[1 + 3 * 0]), 0.0);
                }
                const double fd_ = fdd[n];
                for (int b = 0; b < 3; ++b)
                    for (int a = 0; a < 3; ++a) {
                        const double2 cc = cmul(vd[a], vd[b]), dd = cmul(vd[a], om[b]);
                        acc_c[2 * (a + 3 * b)] += cc.x * fd_; acc_c[2 * (a + 3 * b) + 1] += cc.y * fd_;
                        acc_d[2 * (a + 3 * b)] -= dd.x * fd_; acc_d[2 * (a + 3 * b) + 1] -= dd.y * fd_;
                    }
            }
        }
        if (!want_opt && !want_dw) continue;
        // ---- frequency-dependent pair sums: kubo_opt herm / aherm (kubo.f90:272-405), gyro Dw (gyro.f90:153-197, 235-268)
        const int NORD = nw * (nw - 1);
        for (int p0 = 0; p0 < NORD; p0 += S_TILE) {
            __syncthreads();
            for (int t = tid; t < S_TILE; t += S_THREADS) {
                const int pidx = p0 + t;
                double *it = tile + t * S_ITEM, *ig = tileg + t * S_ITEM;
                if (pidx >= NORD) { for (int q = 0; q < S_ITEM; ++q) { it[q] = 0.0; ig[q] = 0.0; } it[0] = 1e300; ig[0] = 1e300; continue; }
                const int mm = pidx / (nw - 1);
                int n = pidx % (nw - 1); if (n >= mm) ++n;
                const double dE = es[mm] - es[n], df = fn[mm] - fn[n];
                double2 a_nm[3], a_mn[3];
                for (int a = 0; a < 3; ++a) { a_nm[a] = Ak[(size_t)a * nw2 + n + nw * mm]; a_mn[a] = Ak[(size_t)a * nw2 + mm + nw * n]; }
                it[0] = dE;
                for (int b = 0; b < 3; ++b)
                    for (int i = 0; i < 3; ++i) {                         // dE A(i,n,m) A(b,m,n) (f_m - f_n)
                        const double2 pq = cmul(a_nm[i], a_mn[b]);
                        it[1 + 2 * (i + 3 * b)] = dE * df * pq.x; it[2 + 2 * (i + 3 * b)] = dE * df * pq.y;
                    }
                it[19] = 0.0;
                if (want_dw) {
                    // band index of the sum is the FIRST index of the pair here: (n_band = mm, partner = n):  ww = (E_n - E_mm)^2,
                    // cr = Im[A(:,mm,n) x A(:,n,mm)], weight V(a,mm,mm) f'_mm
                    const double ww = dE * dE;
                    double cr[3];
                    cr[0] = (a_mn[1].x * a_nm[2].y + a_mn[1].y * a_nm[2].x) - (a_mn[2].x * a_nm[1].y + a_mn[2].y * a_nm[1].x);
                    cr[1] = (a_mn[2].x * a_nm[0].y + a_mn[2].y * a_nm[0].x) - (a_mn[0].x * a_nm[2].y + a_mn[0].y * a_nm[2].x);
                    cr[2] = (a_mn[0].x * a_nm[1].y + a_mn[0].y * a_nm[1].x) - (a_mn[1].x * a_nm[0].y + a_mn[1].y * a_nm[0].x);
                    ig[0] = ww;
                    const double fd_ = fdd[mm];
                    for (int b = 0; b < 3; ++b)
                        for (int a = 0; a < 3; ++a) {
                            const double2 vd = Vk[a + 3 * (mm + nw * mm)];
                            ig[1 + 2 * (a + 3 * b)] = vd.x * fd_ * cr[b] * ww; ig[2 + 2 * (a + 3 * b)] = vd.y * fd_ * cr[b] * ww;
                        }
                    ig[19] = 0.0;
                }
            }
            __syncthreads();
            if (!hv) continue;
            const int cnt = min(S_TILE, NORD - p0);
            for (int t = slice; t < cnt; t += NS) {
                if (want_opt) {
                    const double *it = tile + t * S_ITEM;
                    const double x = it[0] - w;
                    const double lor = jb.eta / fma(x, x, jb.eta * jb.eta);   // pi * Im[1 / (pi (dE - i eta) - pi hw)]   (delta_broad :248-269)
                    const double inv = 1.0 / x;                             // 1 / real(dE - hw - i eta)                 (:389)
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const double pr_ = it[1 + 2 * q], pi_ = it[2 + 2 * q];
                        acc_opt[2 * q] = fma(-pr_, lor, acc_opt[2 * q]); acc_opt[2 * q + 1] = fma(-pi_, lor, acc_opt[2 * q + 1]);        // herm
                        acc_opt[18 + 2 * q] = fma(-pi_, inv, acc_opt[18 + 2 * q]); acc_opt[19 + 2 * q] = fma(pr_, inv, acc_opt[19 + 2 * q]);   // i P / x
                    }
                }
                if (want_dw) {
                    const double *ig = tileg + t * S_ITEM;
                    const double g = 1.0 / (ig[0] - w);                     // 1 / (w_mn^2 - hw)   (hw not squared, as coded :262)
#pragma unroll
                    for (int q = 0; q < 18; ++q) acc_dw[q] = fma(ig[1 + q], g, acc_dw[q]);     // Dw -= V f' om~ ,  om~ = - sum cr ww / (ww - hw)
                }
            }
        }
    }
    // ---- block reduction -> row
    double *row = rows + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * row_len;
    __syncthreads();
    if (first_tile) {
        if (jb.do_mep && has_pos && tid < nw && tid < nval) row[tid] = acc_cs_mep;
        // kubo CS, gyro C / D: sum over the bands (threads n < nw)
        double vals[37];
        vals[0] = (tid < nw) ? acc_cs_kub : 0.0;
        for (int q = 0; q < 18; ++q) { vals[1 + q] = (tid < nw) ? acc_c[q] : 0.0; vals[19 + q] = (tid < nw) ? acc_d[q] : 0.0; }
        for (int q = 0; q < 37; ++q) {
            red[tid] = vals[q];
            __syncthreads();
            if (tid == 0) {
                double s = 0.0;
                for (int t = 0; t < nw && t < S_THREADS; ++t) s += red[t];
                if (q == 0) row[nval] = s;
                else if (q <= 18) row[nval + 1 + 36 * n_hw + (q - 1)] = s;
                else row[nval + 1 + 36 * n_hw + 18 + (q - 19)] = s;
            }
            __syncthreads();
        }
    }
    // frequency-dependent accumulators: sum over the pair slices of each frequency
    for (int q = 0; q < 36 + 18; ++q) {
        red[tid] = q < 36 ? acc_opt[q] : acc_dw[q - 36];
        __syncthreads();
        if (slice == 0 && hv) {
            double s = 0.0;
            for (int sl = 0; sl < NS; ++sl) s += red[sl * HB + h_loc];
            if (q < 36) row[nval + 1 + 36 * h + q] = s;
            else row[nval + 1 + 36 * n_hw + 36 + 18 * h + (q - 36)] = s;
        }
        __syncthreads();
    }
}

}  // namespace

size_t posop_gauge_smem(int nw) { return GaugeSmem(nw).total; }
int posop_gauge_grid(int sm_count, long long nk) { return (int)(nk < sm_count ? (nk < 1 ? 1 : nk) : sm_count); }

cudaError_t launch_posop_gauge(int nw, long long nk, int do_gauge, int mode, double kubo_tol, bool has_pos, const double *E,
                               const double2 *U, const double2 *Hka, const double2 *const Aw[3], const double2 *const dA[3], double2 *V,
                               double2 *AH, double2 *OmV, double2 *scratch, int grid, cudaStream_t st) {
    const size_t smem = posop_gauge_smem(nw);
    cudaError_t e = cudaFuncSetAttribute(posop_gauge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    posop_gauge_kernel<<<grid, G_THREADS, smem, st>>>(nw, nk, do_gauge, mode, kubo_tol, has_pos ? 1 : 0, E, U, Hka, Aw[0], Aw[1], Aw[2], dA[0],
                                                      dA[1], dA[2], V, AH, OmV, scratch);
    return cudaGetLastError();
}

int posop_row_len(int nval, int n_hw) { return nval + 1 + 36 * n_hw + 36 + 18 * n_hw; }
int posop_hw_block(int n_hw) { int hb = 1; while (hb < n_hw && hb < S_THREADS) hb <<= 1; return hb; }
int posop_sums_grid(int sm_count, long long nk) { return (int)(nk < 2LL * sm_count ? (nk < 1 ? 1 : nk) : 2LL * sm_count); }

// rows: [tiles][grid][row_len] doubles, zero-initialised by the caller (blocks write only the entries they own)
cudaError_t launch_posop_sums(int nw, long long nk, const Jobs &jb, bool has_pos, bool do_gyro, int n_hw, const double *hw, const double *E,
                              const double2 *V, const double2 *AH, const double2 *OmV, double *rows, int grid, cudaStream_t st) {
    const int HB = posop_hw_block(n_hw), tiles = (n_hw + HB - 1) / HB;
    const size_t smem = sizeof(double) * (3 * (size_t)nw + 2 * S_TILE * S_ITEM + S_THREADS);
    dim3 g((unsigned)grid, (unsigned)tiles);
    posop_sums_kernel<<<g, S_THREADS, smem, st>>>(nw, nk, jb, has_pos ? 1 : 0, do_gyro ? 1 : 0, n_hw, HB, hw, E, V, AH, OmV, rows,
                                                  posop_row_len(jb.nval, n_hw));
    return cudaGetLastError();
}

}  // namespace mepb
