// mepb_sub.cu — band-sum kernel for num_wann 4 and 8 in the response (mesh) mode: NW LANES PER k-POINT.
//
// Successor of the warp-per-k kernel of mepb_small.cu for the hot configurations (that one stays for num_wann = 16, for
// explicit k lists and as the fallback).  A warp carries 32/NW k-points at once; lane g of a group owns row g in the
// Fourier sums, column g in the gauge rotation, and a strided share of the band pairs.
// Everything the groups do is the same instruction stream, so the per-k instruction count drops ~3x against one warp
// per k (profiles/r01c: 2.1 k warp instructions per k, 20 % of the FP64 roofline).  Per k-point:
//   1. phases e^{2 pi i k.R} (sincospi) and tight-binding diagonal phases                (wann_interp.f90:215-223,259-260)
//   2. Fourier sums of d_a H(k) from the sparse or dense H(R)                             (wann_interp.f90:206-279)
//   3. eigenpairs from the thread-per-k solver (mepb_eig.cu)                              (matrix_math.f90:192-226)
//   4. V_a = U^dagger (d_a H) U                                                          (wann_interp.f90:517-528)
//   5. occupations, static AHC, pair items, photo items                                  (kubo.f90, photo.f90:44-81)
//   6. with the MEP jobs: the H-gauge velocities of every k-point go to HBM in the padded layout of the shared-memory buffer
//      (out.V_pad) for mep_sums_kernel (mepb_mep.cu) — fused in here those sums cost this kernel a third of its occupancy
// Output formats (block_static rows, item buffers) are those of mepb_small.cu, so the sweeps and the packing are shared.
#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int kSubWarps = 4;                      // warps per CTA
constexpr int kSubMaxNnz = 256;                   // sparse H(R) entries staged in shared memory (else read through L1)

struct SpRec { double2 val; double c0, c1, c2; int R; int pad; };   // 48 bytes: one non-zero H(R)(n,m) with its weights

template <int NW>
struct SubL {  // shared-memory layout per GROUP (one k-point), bytes
    static constexpr int LD = NW + 1;             // padded rows: row-strided stores of a group (U load, Fourier) hit distinct banks
    static constexpr size_t oHa = 0;                // U stays in registers (column g per lane): 1.1 KB less per k-point
    static constexpr size_t oEs = oHa + sizeof(double2) * 3 * NW * LD;
    static constexpr size_t oFn = oEs + sizeof(double) * NW;
    static constexpr size_t oDn = oFn + sizeof(double) * NW;
    static constexpr size_t oPh = oDn + sizeof(double2) * NW;
    // group stride = 16 mod 128 bytes, so that the same element of the 32/NW groups of a warp falls into different banks
    static size_t bytes(int nR) {
        size_t b = oPh + sizeof(double2) * nR;
        return (b + 127) / 128 * 128 + 16;
    }
};

// PH: build with the photo items;  VIN: band sums on supplied eigenpairs (out.e_in, out.V_in) — Fourier sums and rotation compiled out
template <int NW, bool PH, bool VIN>
__global__ void __launch_bounds__(kSubWarps * 32, 3)                // 168 registers
bands_sub_kernel(const DevModel m, const KSource ks, const Jobs jb, const InterpOut out, const int max_pairs, const size_t grp_bytes,
                 const int rec_cap) {
    using L = SubL<NW>;
    constexpr int LD = L::LD, NW2 = NW * NW, NPAIR = NW * (NW - 1) / 2, NGRP = 32 / NW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane % NW, grp = lane / NW;
    const unsigned gshift = (unsigned)(grp * NW);
    const unsigned gbits = (NW == 32) ? 0xffffffffu : ((1u << NW) - 1u);
    unsigned char *base = smem_raw + ((size_t)warp * NGRP + grp) * grp_bytes;
    double2 *Ha = reinterpret_cast<double2 *>(base + L::oHa);      // [a][n*LD + m]; V_a after the rotation
    double *es = reinterpret_cast<double *>(base + L::oEs);
    double *fn = reinterpret_cast<double *>(base + L::oFn);
    double2 *dn = reinterpret_cast<double2 *>(base + L::oDn);
    double2 *ph = reinterpret_cast<double2 *>(base + L::oPh);
    const int nR = m.nR;
    const bool has_tau = (m.tau != nullptr);
    constexpr bool vin = VIN;                                     // band sums on supplied (e, V): no Fourier sums, no rotation
    // CTA-wide copy of the sparse H(R) list (after the group buffers)
    SpRec *rec = reinterpret_cast<SpRec *>(smem_raw + (size_t)kSubWarps * NGRP * grp_bytes);
    int *sptr = reinterpret_cast<int *>(rec + rec_cap);
    const bool sp_smem = (m.sp_ptr != nullptr) && (m.sp_nnz <= kSubMaxNnz);
    if (sp_smem) {
        for (int t = threadIdx.x; t < m.sp_nnz; t += blockDim.x) {
            const int R = m.sp_R[t];
            SpRec r; r.val = m.sp_val[t]; r.c0 = m.cR[3 * R]; r.c1 = m.cR[3 * R + 1]; r.c2 = m.cR[3 * R + 2]; r.R = R; r.pad = 0;
            rec[t] = r;
        }
        const int n_ptr = NW2 + 1 + (m.sp_up ? 1 + (NW + 1) + NW * (NW + 1) / 2 : 0);     // + the upper-triangle lane schedule
        for (int i = threadIdx.x; i < n_ptr; i += blockDim.x) sptr[i] = m.sp_ptr[i];
    }
    const bool up_sched = sp_smem && m.sp_up;                     // Hermitian-closed model: upper triangle + mirror (section 2)
    const int *up_off = sptr + NW2 + 2, *up_el = sptr + NW2 + 2 + (NW + 1);
    __syncthreads();

    // per-lane accumulators over all k-points of this thread's groups
    double acc_nk = 0.0, acc_ndeg = 0.0, acc_ne = 0.0, acc_nit = 0.0, acc_nph = 0.0, acc_ahc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc_ahc[q] = 0.0;

    const long long gstride = (long long)gridDim.x * kSubWarps * NGRP;
    const long long first = ((long long)blockIdx.x * kSubWarps + warp) * NGRP + grp;
    const long long n_iter = (ks.nk + gstride - 1) / gstride;     // same trip count for every warp: __syncwarp stays uniform
    for (long long it = 0; it < n_iter; ++it) {
        const long long ik_raw = first + it * gstride;
        const bool valid = ik_raw < ks.nk;
        const long long ik = valid ? ik_raw : ks.nk - 1;          // idle groups recompute the last k-point and discard it
        // ------------------------------------------------------------------ k-point (k_space.f90:351-375)
        double k1, k2, k3;
        {
            const long long lin = ks.lin_begin + ik;
            long long q1, q2, q3;
            if (lin < 0x7fffffffLL) {
                const unsigned l32 = (unsigned)lin, n1 = (unsigned)ks.N1, n2 = (unsigned)ks.N2;
                const unsigned t = l32 / n1;
                q1 = l32 - t * n1 + 1; const unsigned t2 = t / n2; q2 = t - t2 * n2 + 1; q3 = t2 + 1;
            } else { q1 = lin % ks.N1 + 1; q2 = (lin / ks.N1) % ks.N2 + 1; q3 = lin / ((long long)ks.N1 * ks.N2) + 1; }
            k1 = (2.0 * (double)q1 - (double)ks.N1 - 1.0) / (2.0 * (double)ks.N1);
            k2 = (2.0 * (double)q2 - (double)ks.N2 - 1.0) / (2.0 * (double)ks.N2);
            k3 = (2.0 * (double)q3 - (double)ks.N3 - 1.0) / (2.0 * (double)ks.N3);
        }
        // ------------------------------------------------------------------ 1. phases; 3. eigenpairs (loads issued early)
        for (int R = g; R < nR && !vin; R += NW) {
            const double x = 2.0 * (k1 * __ldg(&m.Rf[3 * R]) + k2 * __ldg(&m.Rf[3 * R + 1]) + k3 * __ldg(&m.Rf[3 * R + 2]));
            double s, c; sincospi(x, &s, &c);
            ph[R] = make_double2(c, s);
        }
        if (has_tau && !vin) {
            const double x = 2.0 * (k1 * __ldg(&m.tau[3 * g]) + k2 * __ldg(&m.tau[3 * g + 1]) + k3 * __ldg(&m.tau[3 * g + 2]));
            double s, c; sincospi(x, &s, &c);
            dn[g] = make_double2(c, s);
        }
        double2 u[NW];                                             // column g of U (eigenvector of band g)
        constexpr bool kDmmaRot = (NW == 8) && !VIN;                   // tensor-core rotation (section 4)
        // its U fragments, U(2j + ks, lane >> 2) of one k-point at a time, are fetched one k-point ahead: the first here, behind the
        // Fourier sums, the next ones behind the multiply-adds of their predecessor (ncu r03v: 14 % of the kernel's stall samples sat
        // on the first DMMA of a k-point, waiting for this load; fetching all four up front costs 32 registers and was slower).
        // Not in the photo build: there it cost 16 % (r04n), the fragments are fetched where they are used.
        constexpr bool kPipeU = kDmmaRot && !PH;
        double2 un0 = make_double2(0.0, 0.0), un1 = un0;
        auto fetch_u = [&](int gq) {
            long long ikq = first - grp + gq + it * gstride;
            if (ikq >= ks.nk) ikq = ks.nk - 1;
            ld_global_nc_256(out.U_in + (size_t)ikq * NW2 + (size_t)NW * (lane >> 2) + 2 * (lane & 3), un0, un1);   // 32 contiguous bytes
        };
        if (kPipeU && jb.do_gauge) fetch_u(0);
        {
            if (!vin && !kDmmaRot) {                                 // (the tensor-core rotation has its own fragments)
                const double2 *Uk = out.U_in + (size_t)ik * NW2 + (size_t)NW * g;   // element r + NW*c: 128 contiguous bytes per lane
#pragma unroll
                for (int q = 0; q < NW; q += 2) ld_global_nc_256(Uk + q, u[q], u[q + 1]);
            }
            const double e = out.e_in[(size_t)ik * NW + g];
            const double f = fermi(e, jb);
            es[g] = e; fn[g] = f;
            if (valid) acc_ne += f;
        }
        __syncwarp();
        // ------------------------------------------------------------------ 2. Fourier sums of d_a H, row n = g
        if (vin) {                                                  // supplied velocities: user layout a + 3 (n + NW m)
            const double2 *Vk = out.V_in + (size_t)ik * 3 * NW2;
            for (int idx = g; idx < 3 * NW2; idx += NW) {
                const int a = idx % 3, n = (idx / 3) % NW, mm = idx / (3 * NW);
                Ha[a * NW * LD + n * LD + mm] = __ldg(Vk + idx);
            }
        } else {
            // rows: lane g sums row n = g (NW elements).  Hermitian-closed sparse models: the lanes share the upper triangle by
            // non-zero count (host schedule) and mirror each element — half the records, evenly spread
            const int e0 = up_sched ? up_off[g] : 0, e1 = up_sched ? up_off[g + 1] : NW;
#pragma unroll 1
            for (int e = e0; e < e1; ++e) {
                const int code = up_sched ? up_el[e] : 0;
                const int n = up_sched ? (code & 0xff) : g, mm = up_sched ? (code >> 8) : e;
                const int idx = n + NW * mm;
                double2 hk = make_double2(0, 0), gs[3] = {make_double2(0, 0), make_double2(0, 0), make_double2(0, 0)};
                if (sp_smem) {
                    const int t1 = sptr[idx + 1];
                    for (int t = sptr[idx]; t < t1; ++t) {
                        const SpRec r = rec[t];
                        const double2 tt = cmul(ph[r.R], r.val);
                        hk.x += tt.x; hk.y += tt.y;
                        gs[0].x = fma(r.c0, tt.x, gs[0].x); gs[0].y = fma(r.c0, tt.y, gs[0].y);
                        gs[1].x = fma(r.c1, tt.x, gs[1].x); gs[1].y = fma(r.c1, tt.y, gs[1].y);
                        gs[2].x = fma(r.c2, tt.x, gs[2].x); gs[2].y = fma(r.c2, tt.y, gs[2].y);
                    }
                } else if (m.sp_ptr) {
                    const int t1 = __ldg(&m.sp_ptr[idx + 1]);
                    for (int t = __ldg(&m.sp_ptr[idx]); t < t1; ++t) {
                        const int R = __ldg(&m.sp_R[t]);
                        const double2 tt = cmul(ph[R], __ldg(&m.sp_val[t]));
                        const double c0 = __ldg(&m.cR[3 * R]), c1 = __ldg(&m.cR[3 * R + 1]), c2 = __ldg(&m.cR[3 * R + 2]);
                        hk.x += tt.x; hk.y += tt.y;
                        gs[0].x = fma(c0, tt.x, gs[0].x); gs[0].y = fma(c0, tt.y, gs[0].y);
                        gs[1].x = fma(c1, tt.x, gs[1].x); gs[1].y = fma(c1, tt.y, gs[1].y);
                        gs[2].x = fma(c2, tt.x, gs[2].x); gs[2].y = fma(c2, tt.y, gs[2].y);
                    }
                } else {
#pragma unroll 3
                    for (int R = 0; R < nR; ++R) {
                        const double2 tt = cmul(ph[R], __ldg(&m.H[(size_t)R * NW2 + idx]));
                        const double c0 = __ldg(&m.cR[3 * R]), c1 = __ldg(&m.cR[3 * R + 1]), c2 = __ldg(&m.cR[3 * R + 2]);
                        hk.x += tt.x; hk.y += tt.y;
                        gs[0].x = fma(c0, tt.x, gs[0].x); gs[0].y = fma(c0, tt.y, gs[0].y);
                        gs[1].x = fma(c1, tt.x, gs[1].x); gs[1].y = fma(c1, tt.y, gs[1].y);
                        gs[2].x = fma(c2, tt.x, gs[2].x); gs[2].y = fma(c2, tt.y, gs[2].y);
                    }
                }
                double2 ga[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) ga[a] = make_double2(-gs[a].y, gs[a].x);                  // * i
                if (has_tau) {
                    const double2 w = cmulc(dn[mm], dn[n]);                                          // d_m conj(d_n)
                    hk = cmul(hk, w);
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        ga[a] = cmul(ga[a], w);
                        const double dt = __ldg(&m.tM[3 * mm + a]) - __ldg(&m.tM[3 * n + a]);
                        ga[a].x = fma(-dt, hk.y, ga[a].x); ga[a].y = fma(dt, hk.x, ga[a].y);         // + i dt H
                    }
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) Ha[a * NW * LD + n * LD + mm] = ga[a];
                if (up_sched && n != mm) {
#pragma unroll
                    for (int a = 0; a < 3; ++a) Ha[a * NW * LD + mm * LD + n] = cconj(ga[a]);
                }
            }
        }
        __syncwarp();
        // ------------------------------------------------------------------ 4. gauge rotation V_a = U^dagger (d_a H) U
        // lane g: column g of T = H_a U from its own u (registers), written over H_a; then ROW g of V_a = U^dagger T,
        // again from its own u — U never goes through shared memory.
        if (kDmmaRot && jb.do_gauge) {      // (every NW = 8 build; the FMA form below serves NW = 4)
            // Tensor-core rotation (mma.sync.m8n8k4.f64 -> SASS DMMA): the warp takes its four k-points one after the other, all 32
            // lanes on one 8 x 8 complex problem.  With the contraction index ordered p = 2 j + ks (j = lane & 3, ks = k-step) the
            // SAME two U elements per lane, U(2j + ks, lane >> 2), are the A fragment of both products
            //     T^T = U^T (d_aH)^T          (B fragment: two consecutive elements of a row of d_aH in shared memory)
            //     V   = conj(U)^T T           (B fragment: the accumulator registers of the first product, no shuffle)
            // so per k-point and direction the shared-memory traffic is 2 LDS.128 + 2 STS.128 per lane instead of 128 + 16 for the
            // FMA formulation below, whose LDS stream co-limited the FP64 pipe (ncu r02b: pipe 43 %, stalls short_scoreboard 21 %).
            const int j = lane & 3, rq = lane >> 2;
#pragma unroll 1
            for (int gq = 0; gq < NGRP; ++gq) {
                if (!kPipeU) fetch_u(gq);
                const double2 u0 = un0, u1 = un1;
                if (kPipeU && gq + 1 < NGRP) fetch_u(gq + 1);
                double2 *Hq = reinterpret_cast<double2 *>(smem_raw + ((size_t)warp * NGRP + gq) * grp_bytes + L::oHa);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    double2 *Hrow = Hq + a * NW * LD;
                    const double2 h0 = Hrow[rq * LD + 2 * j], h1 = Hrow[rq * LD + 2 * j + 1];      // d_aH(p = rq, q = 2j + ks)
                    double tr[2] = {0.0, 0.0}, ti[2] = {0.0, 0.0};
#define MEPB_DMMA(c0, c1, av, bv) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(av), "d"(bv))
                    MEPB_DMMA(tr[0], tr[1], u0.x, h0.x); MEPB_DMMA(ti[0], ti[1], u0.x, h0.y);
                    MEPB_DMMA(tr[0], tr[1], -u0.y, h0.y); MEPB_DMMA(ti[0], ti[1], u0.y, h0.x);
                    MEPB_DMMA(tr[0], tr[1], u1.x, h1.x); MEPB_DMMA(ti[0], ti[1], u1.x, h1.y);
                    MEPB_DMMA(tr[0], tr[1], -u1.y, h1.y); MEPB_DMMA(ti[0], ti[1], u1.y, h1.x);
                    // lane holds T(p = 2j + e, m = rq), e = 0, 1: the B fragment of the second product for k-step e
                    double vr[2] = {0.0, 0.0}, vi[2] = {0.0, 0.0};
                    MEPB_DMMA(vr[0], vr[1], u0.x, tr[0]); MEPB_DMMA(vi[0], vi[1], u0.x, ti[0]);
                    MEPB_DMMA(vr[0], vr[1], u0.y, ti[0]); MEPB_DMMA(vi[0], vi[1], -u0.y, tr[0]);
                    MEPB_DMMA(vr[0], vr[1], u1.x, tr[1]); MEPB_DMMA(vi[0], vi[1], u1.x, ti[1]);
                    MEPB_DMMA(vr[0], vr[1], u1.y, ti[1]); MEPB_DMMA(vi[0], vi[1], -u1.y, tr[1]);
#undef MEPB_DMMA
                    __syncwarp();                                    // every lane has its d_aH operands: V_a may overwrite them
                    Hrow[rq * LD + 2 * j] = make_double2(vr[0], vi[0]);                          // V_a(n = rq, m = 2j), (n, 2j + 1)
                    Hrow[rq * LD + 2 * j + 1] = make_double2(vr[1], vi[1]);
                }
            }
            __syncwarp();
        } else if (jb.do_gauge && !vin) {
#pragma unroll 1
            for (int a = 0; a < 3; ++a) {
                double2 *Hrow = Ha + a * NW * LD;
                double2 tmp[NW];
#pragma unroll
                for (int p = 0; p < NW; ++p) {
                    double2 s = make_double2(0, 0);
#pragma unroll
                    for (int q = 0; q < NW; ++q) cfma(s, Hrow[p * LD + q], u[q]);
                    tmp[p] = s;
                }
                __syncwarp();                                       // every lane has read H_a: column g can be overwritten
#pragma unroll
                for (int p = 0; p < NW; ++p) Hrow[p * LD + g] = tmp[p];
                __syncwarp();                                       // T complete
#pragma unroll
                for (int mm = 0; mm < NW; ++mm) {
                    double2 s = make_double2(0, 0);
#pragma unroll
                    for (int p = 0; p < NW; ++p) cfmac(s, u[p], Hrow[p * LD + mm]);
                    tmp[mm] = s;
                }
                __syncwarp();                                       // every lane has read T: row g can be overwritten
#pragma unroll
                for (int mm = 0; mm < NW; ++mm) Hrow[g * LD + mm] = tmp[mm];
            }
            __syncwarp();
        }
        const double2 *Vs = Ha;                                     // Vs[a*NW*LD + n*LD + m] = V_ka(a,n,m)
        if (valid && g == 0) acc_nk += 1.0;
        // ------------------------------------------------------------------ 6. velocities for mep_sums_kernel (MEP jobs only)
        // The padded buffers of the warp's k-points leave as they are: 3 NW (NW + 1) double2 each, coalesced 128-bit stores
        // (32 lanes x 16 contiguous bytes per instruction).  (The rotation ended with a __syncwarp: every buffer is complete.)
        if (out.V_pad) {
            constexpr int VSZ = 3 * NW * LD;
#pragma unroll 1
            for (int gq = 0; gq < NGRP; ++gq) {
                const long long ikq = first - grp + gq + it * gstride;
                if (ikq < ks.nk) {
                    const double2 *src = reinterpret_cast<const double2 *>(smem_raw + ((size_t)warp * NGRP + gq) * grp_bytes + L::oHa);
                    double2 *dst = out.V_pad + (size_t)ikq * VSZ;
#pragma unroll
                    for (int idx = lane; idx < VSZ; idx += 32) dst[idx] = src[idx];
                }
            }
        }
        // ------------------------------------------------------------------ 5. pair sums + AHC items
        {
            const int istride = jb.herm ? kAhcItemH : kAhcItemG;
            int n_items_before = 0;
#pragma unroll 1
            for (int p0 = 0; p0 < NPAIR; p0 += NW) {
                const int pidx = p0 + g;
                bool emit = false;
                double item[kAhcItemG];
                if (pidx < NPAIR) {
                    // pair index -> (n < mm), pidx = mm (mm - 1) / 2 + n: closed form (single precision is exact far beyond NPAIR)
                    const int mm = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pidx)) * 0.5f);
                    const int n = pidx - ((mm * (mm - 1)) >> 1);
                    const double dE = es[mm] - es[n];
                    const double f_n = fn[n], f_m = fn[mm], fdm = f_m - f_n;
                    if (dE == 0.0 && valid) acc_ndeg += 1.0;
                    if (jb.do_ahc) {
                        double2 vnm[3], vmn[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) { vnm[a] = Vs[a * NW * LD + n * LD + mm]; vmn[a] = Vs[a * NW * LD + mm * LD + n]; }
                        double2 P[9];   // P(i,j) = V(i,n,m) V(j,m,n) at i + 3j
#pragma unroll
                        for (int j = 0; j < 3; ++j)
#pragma unroll
                            for (int i = 0; i < 3; ++i) P[i + 3 * j] = cmul(vnm[i], vmn[j]);
                        const double rdE = 1.0 / dE;                                        // unguarded like kubo.f90:65,186
                        if (valid && (f_n != 0.0 || f_m != 0.0 || dE == 0.0)) {          // kubo.f90:58-73, unguarded
                            const double inv2 = 2.0 * rdE * rdE;
#pragma unroll
                            for (int j = 0; j < 3; ++j)
#pragma unroll
                                for (int i = 0; i < 3; ++i) acc_ahc[i + 3 * j] += inv2 * (f_n * P[i + 3 * j].y + f_m * P[j + 3 * i].y);
                        }
                        if (fdm != 0.0 || dE == 0.0) {
                            const double wa = fdm * rdE;
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
                            emit = valid && (out.items_ahc != nullptr);
                        }
                    }
                }
                const unsigned bal = (__ballot_sync(0xffffffffu, emit) >> gshift) & gbits;      // this group's lanes
                if (emit) {
                    const int slot = n_items_before + __popc(bal & ((1u << g) - 1u));
                    store_ahc_item(out.items_ahc + ((size_t)ik * max_pairs + slot) * istride, item, jb.herm);
                }
                n_items_before += __popc(bal);
            }
            if (valid && out.cnt_ahc && g == 0) { out.cnt_ahc[ik] = n_items_before; acc_nit += (double)n_items_before; }
        }
        // ------------------------------------------------------------------ photo items (photo.f90:44-81)
        // PAIRED ORDER: the items of the ordered band pairs (n, l) and (l, n) go into two consecutive slots (2p, 2p + 1).
        // Their frequency factors obey h(l,n; w) = -conj h(n,l; w) (x -> -conj x in 2x/(x^2 - w^2), x = E_n - E_l - 2 i eta),
        // which lets the tensor-core sweep (mepb_sweep_mma.cu, PAIRED) contract both with ONE evaluation of h and half the
        // multiply-adds; the FMA sweeps simply see ordinary items.  The emitting pairs (f_l != f_n) are first compacted into a
        // list so that the lanes of a group share them evenly (16 pairs on 8 lanes at half filling: 2 each).
        if (PH && jb.do_opt && out.items_photo) {
            constexpr int NORD = NW * (NW - 1);
            unsigned char *plist = reinterpret_cast<unsigned char *>(dn);          // the centre phases are dead by now: NW x 16 B >= NPAIR bytes
            int n_pairs = 0;
#pragma unroll 1
            for (int p0 = 0; p0 < NPAIR; p0 += NW) {
                const int pidx = p0 + g;
                bool emit = false;
                int n0 = 0, l0 = 0;
                if (pidx < NPAIR) {
                    l0 = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)pidx)) * 0.5f);
                    n0 = pidx - ((l0 * (l0 - 1)) >> 1);                         // n0 < l0
                    emit = valid && (fn[l0] - fn[n0] != 0.0);
                }
                const unsigned bal = (__ballot_sync(0xffffffffu, emit) >> gshift) & gbits;
                if (emit) plist[n_pairs + __popc(bal & ((1u << g) - 1u))] = (unsigned char)(n0 | (l0 << 4));
                n_pairs += __popc(bal);
            }
            __syncwarp();
#pragma unroll 1
            for (int idx = g; idx < n_pairs; idx += NW) {
                const int code = plist[idx], n0 = code & 15, l0 = code >> 4;
#pragma unroll 1
                for (int dir = 0; dir < 2; ++dir) {
                    const int n = dir ? l0 : n0, l = dir ? n0 : l0;
                    const double df = fn[l] - fn[n];
                    double2 W[9];
#pragma unroll
                    for (int q = 0; q < 9; ++q) W[q] = make_double2(0, 0);
#pragma unroll 1
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
                    // The two items of a pair are one 928-byte run that starts on a 32-byte boundary: 29 256-bit stores (STG.256),
                    // each a whole sector.  Lanes write to 32 different lines per instruction whatever the width, so half the
                    // instructions is half the LSU time of this store-bound section (r03s: 58 128-bit stores per pair).
                    double2 *run = reinterpret_cast<double2 *>(out.items_photo + ((size_t)ik * NORD + 2 * idx) * kPhotoItem);
                    double2 va[3];
#pragma unroll
                    for (int a = 0; a < 3; ++a) va[a] = cscale(df, Vs[a * NW * LD + n * LD + l]);
                    const double2 xr = make_double2(es[n] - es[l], 0.0), zero = make_double2(0.0, 0.0);
                    auto elem = [&](int e) { return e < 27 ? cmul(va[e % 3], W[e / 3]) : (e == 27 ? xr : zero); };   // e: compile time
                    if (dir == 0) {
#pragma unroll
                        for (int k = 0; k < 14; ++k) st_global_256(run + 2 * k, elem(2 * k), elem(2 * k + 1));
                    } else {
                        st_global_256(run + 28, zero, elem(0));                 // pad of the first item | head of the second
#pragma unroll
                        for (int k = 0; k < 14; ++k) st_global_256(run + 30 + 2 * k, elem(2 * k + 1), elem(2 * k + 2));
                    }
                }
            }
            if (valid && g == 0) { out.cnt_photo[ik] = 2 * n_pairs; acc_nph += (double)(2 * n_pairs); }
        }
        __syncwarp();                                                   // the group's buffers are reused by the next k
    }
    // ---------------------------------------------------------------------- block reduction of the static sums
    __syncthreads();
    constexpr int NS = S_NIU + 18 * NW;
    double *stage = reinterpret_cast<double *>(smem_raw);               // [kSubWarps][NS]  (group buffers are dead now)
    {
        const double nk = warp_sum(acc_nk), nd = warp_sum(acc_ndeg), ne = warp_sum(acc_ne);
        const double nit = warp_sum(acc_nit), nph = warp_sum(acc_nph);
#pragma unroll
        for (int q = 0; q < 9; ++q) acc_ahc[q] = warp_sum(acc_ahc[q]);
        double *wv = stage + (size_t)warp * NS;
        if (lane == 0) {
            wv[S_NK] = nk; wv[S_NDEG] = nd; wv[S_NE] = ne; wv[S_NIT] = nit; wv[S_NPH] = nph;
#pragma unroll
            for (int q = 0; q < 9; ++q) { wv[S_AHC + q] = acc_ahc[q]; wv[S_KIC + q] = 0.0; wv[S_KLC + q] = 0.0; }   // MEP slots:
        }
        for (int i = lane; i < 18 * NW; i += 32) wv[S_NIU + i] = 0.0;                                                  // mep_sums_kernel
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < kSubWarps; ++w) s += stage[(size_t)w * NS + i];
        out.block_static[(size_t)blockIdx.x * NS + i] = s;
    }
}

inline int sub_rec_cap(int nnz) { return (nnz > 0 && nnz <= kSubMaxNnz) ? (nnz + 7) / 8 * 8 : 8; }   // staged sparse records

template <int NW>
size_t sub_smem(int nR, int nnz = kSubMaxNnz) {
    const size_t groups = (size_t)kSubWarps * (32 / NW) * SubL<NW>::bytes(nR) + sizeof(SpRec) * sub_rec_cap(nnz) + sizeof(int) * (2 * NW * NW + 8);
    const size_t stage = sizeof(double) * kSubWarps * (S_NIU + 18 * NW);
    return groups > stage ? groups : stage;
}

template <int NW>
cudaError_t launch_sub_t(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs, int grid,
                         cudaStream_t st) {
    const int rec_cap = sub_rec_cap(m.sp_ptr ? m.sp_nnz : 0);
    const bool ph = jobs.do_opt != 0;
    const size_t smem = sub_smem<NW>(m.nR, m.sp_ptr ? m.sp_nnz : 0);
    cudaError_t e;
#define MEPB_LAUNCH_SUB(P, V)                                                                                                  \
    do {                                                                                                                        \
        e = cudaFuncSetAttribute(bands_sub_kernel<NW, P, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (e != cudaSuccess) return e;                                                                                         \
        bands_sub_kernel<NW, P, V><<<grid, kSubWarps * 32, smem, st>>>(m, ks, jobs, out, max_pairs, SubL<NW>::bytes(m.nR), rec_cap); \
    } while (0)
    if (out.V_in) MEPB_LAUNCH_SUB(true, true);             // supplied eigenpairs: one build with every job compiled in
    else if (ph) MEPB_LAUNCH_SUB(true, false);
    else MEPB_LAUNCH_SUB(false, false);
#undef MEPB_LAUNCH_SUB
    return cudaGetLastError();
}

}  // namespace

bool sub_supported(int nw, int nR) {
    if (nw == 4) return sub_smem<4>(nR) <= 100 * 1024;
    if (nw == 8) return sub_smem<8>(nR) <= 100 * 1024;
    return false;
}

int bands_sub_grid(int nw, int nR, int sm_count, long long nk, int nnz) {
    const size_t smem = (nw == 4 ? sub_smem<4>(nR, nnz) : sub_smem<8>(nR, nnz)) + 1024;
    int per_sm = (int)((227 * 1024) / smem);
    if (per_sm > 3) per_sm = 3;                        // __launch_bounds__(128, 3)
    if (per_sm < 1) per_sm = 1;
    const int kpb = kSubWarps * (32 / nw);
    long long need = (nk + kpb - 1) / kpb;
    long long g = (long long)sm_count * per_sm;
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t launch_bands_sub(const DevModel &m, const KSource &ks, const Jobs &jobs, const InterpOut &out, int max_pairs, int grid,
                             cudaStream_t st) {
    switch (m.nw) {
        case 4: return launch_sub_t<4>(m, ks, jobs, out, max_pairs, grid, st);
        case 8: return launch_sub_t<8>(m, ks, jobs, out, max_pairs, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mepb
