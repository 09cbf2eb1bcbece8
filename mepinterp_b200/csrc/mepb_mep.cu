// mepb_mep.cu — the orbital-magnetoelectric band sums of the num_wann 4 / 8 path as a kernel of their own.
//
//   mep_niu_IC / mep_niu_LC     src/mep_niu.f90:47-159     band-resolved, n0 <= valence_bands, terms with |den| <= kubo_tol skipped
//   kubo_mep_IC / kubo_mep_LC   src/kubo_mep.f90:51-149    weighted with f(E_n0), no guard
//
// Why a separate kernel.  Fused into bands_sub_kernel (round 1 and most of round 2) the 36 running sums and the triple-band loops
// pushed that kernel to 254 registers and 82 KB of shared memory per CTA: two CTAs of four warps per SM, two warps per SM
// sub-partition for a kernel whose stalls are fixed-latency dependencies (ncu r03w: 37 % `wait`, FP64 pipe 36 %) — and the Fourier
// sums, the rotation and the pair sums in front of the loops ran at that occupancy too.  Here the producer (bands_sub_kernel, 168
// registers, three CTAs per SM) leaves the H-gauge velocities V_a(n, m) of every k-point in HBM in exactly the padded layout its
// shared-memory buffer has (3 x NW rows of NW + 1 double2: 3456 bytes for NW = 8), and this kernel streams them back: three CTAs
// of four warps per SM at <= 168 registers.  The round trip is 6.9 KB per k-point (C3, 256^3: 116 GB per pass, ~ 18 ms of HBM time
// next to a 108 ms kernel) and buys both halves their occupancy.
//
// Work distribution as in bands_sub_kernel: NW lanes per k-point (lane = band n0), 32 / NW k-points per warp.
#include <cstdint>
#include <cuda_runtime.h>

#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int kMepWarps = 4;

// a b - c d for complex operands as two chains of four multiply-adds (cmul / cmul / csub: ten operations)
__device__ __forceinline__ double2 cdet(const double2 a, const double2 b, const double2 c, const double2 d) {
    double x = a.x * b.x, y = a.x * b.y;
    x = fma(-a.y, b.y, x); y = fma(a.y, b.x, y);
    x = fma(-c.x, d.x, x); y = fma(-c.x, d.y, y);
    x = fma(c.y, d.y, x); y = fma(-c.y, d.x, y);
    return make_double2(x, y);
}

template <int NW>
struct MepL {    // shared memory per group (one k-point), bytes
    static constexpr int LD = NW + 1;
    static constexpr int VSZ = 3 * NW * LD;                        // double2 per k-point in HBM and in shared memory
    static constexpr size_t oV = 0;
    static constexpr size_t oEs = oV + sizeof(double2) * VSZ;
    static constexpr size_t oRt = oEs + sizeof(double) * NW;
    static constexpr size_t raw = oRt + sizeof(double) * NW * (NW + 1);
    static constexpr size_t bytes = (raw + 127) / 128 * 128 + 16;  // group stride 16 mod 128 bytes: the groups of a warp on different banks
};

template <int NW>
__global__ void __launch_bounds__(kMepWarps * 32, 3)
mep_sums_kernel(const long long nk, const Jobs jb, const double *__restrict__ e_in, const double2 *__restrict__ V_pad,
                double *__restrict__ block_rows) {
    using L = MepL<NW>;
    constexpr int NGRP = 32 / NW, LD = L::LD, VSZ = L::VSZ;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane % NW, grp = lane / NW;
    unsigned char *base = smem_raw + ((size_t)warp * NGRP + grp) * L::bytes;
    const double2 *Vs = reinterpret_cast<const double2 *>(base + L::oV);      // Vs[a*NW*LD + n*LD + m] = V_ka(a,n,m)
    double *es = reinterpret_cast<double *>(base + L::oEs);
    double *rt = reinterpret_cast<double *>(base + L::oRt) + g * (NW + 1);   // row n0 = g of the table 1/(E_n0 - E_n)

    double niu_ic[9], niu_lc[9], kub_ic[9], kub_lc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) { niu_ic[q] = 0.0; niu_lc[q] = 0.0; kub_ic[q] = 0.0; kub_lc[q] = 0.0; }

    const long long gstride = (long long)gridDim.x * kMepWarps * NGRP;
    const long long first = ((long long)blockIdx.x * kMepWarps + warp) * NGRP;
    const long long n_iter = (nk + gstride - 1) / gstride;        // same trip count for every warp: __syncwarp stays uniform
    for (long long it = 0; it < n_iter; ++it) {
        // ------------------------------------------------------------------ the warp's 32 / NW k-points: V into shared memory
        // (coalesced 16-byte copies: 32 lanes x 16 bytes of one k-point's contiguous 16 VSZ bytes per instruction)
#pragma unroll
        for (int gq = 0; gq < NGRP; ++gq) {
            const long long ikq = first + gq + it * gstride;
            if (ikq < nk) {
                const double2 *src = V_pad + (size_t)ikq * VSZ;
                double2 *dst = reinterpret_cast<double2 *>(smem_raw + ((size_t)warp * NGRP + gq) * L::bytes + L::oV);
                // cp.async: global -> shared without a register stage, so all 4 x 7 copies of a lane are in flight together (through
                // registers the compiler issued them in batches of one k-point: four L2 round trips per iteration instead of one)
#pragma unroll
                for (int idx = lane; idx < VSZ; idx += 32)
                    cp_async16(dst + idx, src + idx);
            }
        }
        cp_async_commit();
        // next iteration's records towards L2 while this one computes (the warp's 32 / NW records are contiguous: one bulk prefetch;
        // ncu r04p: 23 % of the stall samples were these loads, with no room in shared memory for a second buffer)
        {
            const long long ikn = first + (it + 1) * gstride;
            if (lane == 0 && ikn < nk) {
                const unsigned cnt = (unsigned)((nk - ikn < NGRP) ? (nk - ikn) : NGRP);
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(V_pad + (size_t)ikn * VSZ), "r"(cnt * (unsigned)(sizeof(double2) * VSZ)) : "memory");
            }
        }
        const long long ik = first + grp + it * gstride;
        const bool valid = ik < nk;                                // idle groups (past the end of the launch) contribute nothing
        const double e_own = valid ? e_in[(size_t)ik * NW + g] : 0.0;
        es[g] = e_own;
        // Roles.  A band has work when one of its weights is non-zero: 1 for n0 <= valence_bands of mep_niu, f(E_n0) for kubo_mep.
        // At T = 0 that is half of the bands or fewer (the reference's MEP runs: insulators, FeMn at E_F = 0 with valence_bands = 4 —
        // bands 5..8 never contribute), so lane g is paired with lane g ^ NW/2: when exactly one of the two bands has work, both
        // lanes take half of its n loop.  The mep_niu sums stay band-resolved: a lane that helps is one whose own band has no
        // mep_niu weight ever (its band index is above valence_bands), so its registers only ever hold its partner's band.
        constexpr int P = NW / 2;
        const bool own = valid && ((jb.do_mep && g < jb.nval) || (jb.do_kubo && fermi(e_own, jb) != 0.0));
        const unsigned owners = __ballot_sync(0xffffffffu, own);
        const bool partner_own = (owners >> (lane ^ P)) & 1u;
        const int n0 = own ? g : (g ^ P);
        const int n_lo = own ? 0 : P, n_hi = (own && !partner_own) ? P : NW;
        cp_async_wait_all();
        __syncwarp();
        if (own || partner_own) {
            const double e0 = es[n0];
            // reciprocals 1/(E_n0 - E_n): one division per band and lane (row g of the table; a helper fills its own idle row with its
            // partner's values) instead of one per (n, m) term
#pragma unroll
            for (int n = 0; n < NW; ++n) {                                    // entry n0 is inf and never used
                // MUFU-seeded reciprocal (7 instructions against ~25 of the IEEE division, eight times per lane and k-point); an exact
                // degeneracy keeps the reference's 1/0 = +-inf (kubo_mep.f90:73,126 has no guard)
                const double d = e0 - es[n];
                rt[n] = (d == 0.0) ? copysign(__longlong_as_double(0x7ff0000000000000LL), d) : fast_rcp(d);
            }
            const double w_niu = (jb.do_mep && n0 < jb.nval) ? 1.0 : 0.0;
            const double w_kub = jb.do_kubo ? fermi(e0, jb) : 0.0;
            double2 v00[3];
#pragma unroll
            for (int a = 0; a < 3; ++a) v00[a] = Vs[a * NW * LD + n0 * LD + n0];
#pragma unroll 1
            for (int n = n_lo; n < n_hi; ++n) {
                if (n == n0) continue;
                const double d0n = e0 - es[n];
                double2 v0n[3], vn0[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) { v0n[a] = Vs[a * NW * LD + n0 * LD + n]; vn0[a] = Vs[a * NW * LD + n * LD + n0]; }
                const double rn = rt[n], rn2 = rn * rn;
                {   // local term: sum_kl eps_jkl Re[V(i,n0,n) V(k,n,n0) V(l,n0,n0)] / (E_n0-E_n)^3
                    //           = Re[V(i,n0,n) C_j] / (E_n0-E_n)^3,  C = V(:,n,n0) x V(:,n0,n0)  (one cross product per n, not one per i)
                    const double den = d0n * d0n * d0n;
                    const bool pass = fabs(den) > jb.kubo_tol;
                    const double inv = rn2 * rn;
                    const double iu = inv * w_kub, ig = pass ? inv * w_niu : 0.0;
                    double2 C[3];
                    C[0] = cdet(vn0[1], v00[2], vn0[2], v00[1]);
                    C[1] = cdet(vn0[2], v00[0], vn0[0], v00[2]);
                    C[2] = cdet(vn0[0], v00[1], vn0[1], v00[0]);
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const double c = fma(v0n[i].x, C[j].x, -(v0n[i].y * C[j].y));
                            kub_lc[i + 3 * j] = fma(c, iu, kub_lc[i + 3 * j]);
                            niu_lc[i + 3 * j] = fma(c, ig, niu_lc[i + 3 * j]);
                        }
                }
                {   // itinerant term: - sum_{m != n0} eps_jkl Re[V(i,n0,n) V(k,n,m) V(l,m,n0)] / ((E_n0-E_n)^2 (E_n0-E_m))
                    double2 Zg[3], Zu[3];
#pragma unroll
                    for (int j = 0; j < 3; ++j) { Zg[j] = make_double2(0, 0); Zu[j] = make_double2(0, 0); }
                    const double d2 = d0n * d0n;
#pragma unroll 2
                    for (int mm = 0; mm < NW; ++mm) {
                        if (mm == n0) continue;
                        const double den = d2 * (e0 - es[mm]);
                        const double inv = rn2 * rt[mm];
                        const bool pass = fabs(den) > jb.kubo_tol;
                        double2 a_[3], b_[3];
#pragma unroll
                        for (int a = 0; a < 3; ++a) { a_[a] = Vs[a * NW * LD + n * LD + mm]; b_[a] = Vs[a * NW * LD + mm * LD + n0]; }
                        double2 X[3];
                        X[0] = cdet(a_[1], b_[2], a_[2], b_[1]);
                        X[1] = cdet(a_[2], b_[0], a_[0], b_[2]);
                        X[2] = cdet(a_[0], b_[1], a_[1], b_[0]);
                        const double wg = pass ? inv : 0.0;                    // guarded weight: one select instead of six predicated FMAs
#pragma unroll
                        for (int j = 0; j < 3; ++j) {
                            Zu[j].x = fma(X[j].x, inv, Zu[j].x); Zu[j].y = fma(X[j].y, inv, Zu[j].y);
                            Zg[j].x = fma(X[j].x, wg, Zg[j].x); Zg[j].y = fma(X[j].y, wg, Zg[j].y);
                        }
                    }
                    double2 vk[3], vg[3];                                      // - weight x V(i,n0,n)
#pragma unroll
                    for (int i = 0; i < 3; ++i) { vk[i] = cscale(-w_kub, v0n[i]); vg[i] = cscale(-w_niu, v0n[i]); }
#pragma unroll
                    for (int j = 0; j < 3; ++j)
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            kub_ic[i + 3 * j] = fma(vk[i].x, Zu[j].x, fma(-vk[i].y, Zu[j].y, kub_ic[i + 3 * j]));
                            niu_ic[i + 3 * j] = fma(vg[i].x, Zg[j].x, fma(-vg[i].y, Zg[j].y, niu_ic[i + 3 * j]));
                        }
                }
            }
        }
        __syncwarp();                                              // the group's buffers are refilled by the next iteration
    }
    // ---------------------------------------------------------------------- block reduction into one row of static sums
    // (the row layout of bands_sub_kernel; every slot this kernel does not own is written as zero and the rows of both kernels
    // are added by reduce_rows_kernel)
    __syncthreads();
    constexpr int NS = S_NIU + 18 * NW;
    double *stage = reinterpret_cast<double *>(smem_raw);          // [kMepWarps][NS]  (group buffers are dead now)
    {
        // mep_niu: what the partner lane gathered as a helper belongs to this band (a partner with an index above valence_bands never
        // has a mep_niu weight of its own); lanes above valence_bands own nothing
        {
            const bool take = (g < jb.nval) && ((g ^ (NW / 2)) >= jb.nval);
#pragma unroll
            for (int q = 0; q < 9; ++q) {
                const double ti = __shfl_xor_sync(0xffffffffu, niu_ic[q], NW / 2), tl = __shfl_xor_sync(0xffffffffu, niu_lc[q], NW / 2);
                niu_ic[q] = (g < jb.nval) ? niu_ic[q] + (take ? ti : 0.0) : 0.0;
                niu_lc[q] = (g < jb.nval) ? niu_lc[q] + (take ? tl : 0.0) : 0.0;
            }
        }
        // band-resolved sums: add the lanes with the same band across the groups of the warp
#pragma unroll
        for (int q = 0; q < 9; ++q) {
#pragma unroll
            for (int o = NW; o < 32; o <<= 1) {
                niu_ic[q] += __shfl_xor_sync(0xffffffffu, niu_ic[q], o);
                niu_lc[q] += __shfl_xor_sync(0xffffffffu, niu_lc[q], o);
                kub_ic[q] += __shfl_xor_sync(0xffffffffu, kub_ic[q], o);
                kub_lc[q] += __shfl_xor_sync(0xffffffffu, kub_lc[q], o);
            }
        }
        // kubo sums run over all bands: reduce over the band lanes as well
#pragma unroll
        for (int q = 0; q < 9; ++q) {
#pragma unroll
            for (int o = NW / 2; o > 0; o >>= 1) {
                kub_ic[q] += __shfl_xor_sync(0xffffffffu, kub_ic[q], o);
                kub_lc[q] += __shfl_xor_sync(0xffffffffu, kub_lc[q], o);
            }
        }
        double *wv = stage + (size_t)warp * NS;
        for (int i = lane; i < S_NIU; i += 32) wv[i] = 0.0;
        __syncwarp();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < 9; ++q) { wv[S_KIC + q] = kub_ic[q]; wv[S_KLC + q] = kub_lc[q]; }
        }
        if (lane < NW) {
#pragma unroll
            for (int q = 0; q < 9; ++q) { wv[S_NIU + 9 * lane + q] = niu_ic[q]; wv[S_NIU + 9 * NW + 9 * lane + q] = niu_lc[q]; }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NS; i += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < kMepWarps; ++w) s += stage[(size_t)w * NS + i];
        block_rows[(size_t)blockIdx.x * NS + i] = s;
    }
}

template <int NW>
size_t mep_smem() {
    const size_t groups = (size_t)kMepWarps * (32 / NW) * MepL<NW>::bytes;
    const size_t stage = sizeof(double) * kMepWarps * (S_NIU + 18 * NW);
    return groups > stage ? groups : stage;
}

template <int NW>
cudaError_t launch_mep_t(long long nk, const Jobs &jobs, const double *e_in, const double2 *V_pad, double *rows, int grid, cudaStream_t st) {
    const size_t smem = mep_smem<NW>();
    cudaError_t e = cudaFuncSetAttribute(mep_sums_kernel<NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mep_sums_kernel<NW><<<grid, kMepWarps * 32, smem, st>>>(nk, jobs, e_in, V_pad, rows);
    return cudaGetLastError();
}

}  // namespace

size_t mep_vpad_elems(int nw) { return (size_t)3 * nw * (nw + 1); }      // double2 per k-point of the padded velocity record

int mep_sums_grid(int nw, int sm_count, long long nk) {
    const int kpb = kMepWarps * (32 / nw);
    long long need = (nk + kpb - 1) / kpb, g = (long long)sm_count * 3;
    if (g > need) g = need;
    return (int)(g < 1 ? 1 : g);
}

cudaError_t launch_mep_sums(int nw, long long nk, const Jobs &jobs, const double *e_in, const double2 *V_pad, double *rows, int grid,
                            cudaStream_t st) {
    switch (nw) {
        case 4: return launch_mep_t<4>(nk, jobs, e_in, V_pad, rows, grid, st);
        case 8: return launch_mep_t<8>(nk, jobs, e_in, V_pad, rows, grid, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace mepb
