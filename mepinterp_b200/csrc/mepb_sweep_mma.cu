// mepb_sweep_mma.cu — the frequency sweeps as FP64 tensor-core contractions (mma.sync.m8n8k4.f64 = SASS DMMA).
//
// Same sums as mepb_sweep.cu (velo_ahc_tens / kubo_ohc_tens, kubo.f90:86-198; photo_2nd_cond, photo.f90:16-86):
//       out[comp, w] += sum_items W[comp, item] * g[item](w)
// but issued as D(8 comps x 8 frequencies) += A(8 comps x 4 items) . B(4 items x 8 frequencies):
//   * lane (k = lane & 3, n = lane >> 2) GENERATES the B element of item k at frequency n (one complex reciprocal per
//     (item, w), exactly the arithmetic of the FMA kernels) — nothing of B ever touches memory;
//   * the A fragment (weights of 8 components x 4 items) is ONE shared-memory load per lane and is reused for the NT
//     frequency tiles a warp owns.  The thread-per-frequency FMA kernels re-read every weight in every thread (broadcast
//     LDS.128): their shared-memory pipe is as loaded as the FP64 pipe (ncu r01i: 2 wavefronts per LDS.128, LSU at 83 % /
//     91 % of the FP64 pipe's cycles for the AHC / photo sweep), which is what holds them at 65 % / 56 % of the FP64 peak.
//     Here the LSU load is < 25 % of the pipe's.
//   * components beyond a multiple of 8 (the 9th AHC weight, T(24..26) of the photo tensor) are accumulated per lane with
//     plain DFMAs on the lane's own (item, frequency) and reduced over the 4 item lanes at the end.
// DFMA and DMMA share one pipe on B200 (tools/fp64_mixed.cu: 37.1 TF each, 36 TF interleaved), so the gain is operand
// bandwidth and issue slots, not peak.  Items stream through shared memory by TMA bulk copies into a ring of stages with
// full/empty mbarriers: a dedicated producer warp refills a stage as soon as every consumer warp has released it, so the
// consumer warps (which own different numbers of frequency tiles) never meet at a CTA-wide barrier.
#include <algorithm>
#include <cstdlib>

#include "mepb_internal.cuh"

namespace mepb {

namespace {

constexpr int kMmaStages = 3;          // AHC sweep (KC k-points per stage)
constexpr int kMmaPhotoStages = 6;     // photo sweep (one k-point, 25 KB at nw = 8, per stage; one CTA per SM)

__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------------ AHC / OHC sweep
// MT full 8-row tiles + REM remainder weights: Hermitian items (9 weights) MT = 1, REM = 1; general items (18) MT = 2, REM = 2.
// (CTA-wide barrier per stage and one elected producer thread: the full/empty ring of the photo kernel with a producer warp
// measured 10 % slower here — 288-thread CTAs, 96-register cap — profiles/r01m.)
// Warp w of the CTA: k-group g = w / wtiles, frequency tiles [(w % wtiles) NT, +NT) (+ blockIdx.y * wtiles_per_pass * NT).
template <bool HERM, int NT>
__global__ void __launch_bounds__(256, (HERM && NT <= 2) ? 3 : 2)
sweep_ahc_mma_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_pairs, const long long nk,
                     const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                     const int G, const int wtiles, const int KC) {
    constexpr int NACC = HERM ? kSweepAccH : kSweepAcc;
    constexpr int NWT = NACC / 2;                                // weights per item: 9 or 18
    constexpr int MT = NWT / 8, REM = NWT % 8;
    constexpr int ISTRD = HERM ? kAhcItemH : kAhcItemG;          // item stride in doubles
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const size_t item_bytes = sizeof(double) * (size_t)KC * max_pairs * ISTRD;
    const size_t stage_bytes = (item_bytes + 16 * ((KC * sizeof(int) + 15) / 16) + 127) / 128 * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm_raw);
    unsigned char *stage0 = sm_raw + 128;
    long long per = (nk + gridDim.x - 1) / gridDim.x;
    per = (per + KC - 1) / KC * KC;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    const int nchunks = kb0 < kb1 ? (int)((kb1 - kb0 + KC - 1) / KC) : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMmaStages; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        unsigned char *dst = stage0 + (size_t)(c % kMmaStages) * stage_bytes;
        const long long k0 = kb0 + (long long)c * KC;
        const uint32_t cb = (uint32_t)(16 * ((KC * sizeof(int) + 15) / 16));
        mbar_expect_tx(&bar[c % kMmaStages], (uint32_t)item_bytes + cb);
        tma_bulk_g2s(dst, reinterpret_cast<const double *>(items) + (size_t)k0 * max_pairs * ISTRD, (uint32_t)item_bytes, &bar[c % kMmaStages]);
        tma_bulk_g2s(dst + item_bytes, cnt + k0, cb, &bar[c % kMmaStages]);
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < kMmaStages && c < nchunks; ++c) issue(c);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = lane & 3, nq = lane >> 2;                     // A: row nq, col kq;  B: row kq, col nq;  C: row nq, cols 2 kq, 2 kq + 1
    const int g = warp / wtiles;
    const int tile0 = ((warp % wtiles) + blockIdx.y * wtiles) * NT;
    const bool active = g < G;
    double om2[NT], b2[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int f = 8 * (tile0 + t) + nq;
        const double om = (f < n_hw) ? hw[f] : 0.0;
        const double b = -2.0 * om * eta;
        om2[t] = om * om; b2[t] = b * b;
    }
    double c1[NT][MT > 0 ? MT : 1][2], c2[NT][MT > 0 ? MT : 1][2], r1[NT][REM > 0 ? REM : 1], r2[NT][REM > 0 ? REM : 1];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int q = 0; q < MT; ++q) { c1[t][q][0] = c1[t][q][1] = c2[t][q][0] = c2[t][q][1] = 0.0; }
#pragma unroll
        for (int q = 0; q < REM; ++q) { r1[t][q] = 0.0; r2[t][q] = 0.0; }
    }
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % kMmaStages;
        mbar_wait(&bar[st], (uint32_t)((c / kMmaStages) & 1));
        if (active) {
            const unsigned char *base = stage0 + (size_t)st * stage_bytes;
            const double *its = reinterpret_cast<const double *>(base);
            const int *cn = reinterpret_cast<const int *>(base + item_bytes);
            const long long k0 = kb0 + (long long)c * KC;
            const int kcv = (int)((kb1 - k0 < KC) ? (kb1 - k0) : KC);
            for (int kk = g; kk < kcv; kk += G) {
                const int n_it = cn[kk];
                const double *itk = its + (size_t)kk * max_pairs * ISTRD;
                for (int j0 = 0; j0 < n_it; j0 += 4) {
                    const bool valid = (j0 + kq) < n_it;
                    const double *it = itk + (size_t)(valid ? j0 + kq : 0) * ISTRD;
                    const double cit = valid ? it[0] : 1e6;                       // padding items: weights 0, finite g
                    double a[MT > 0 ? MT : 1], r[REM > 0 ? REM : 1];
#pragma unroll
                    for (int q = 0; q < MT; ++q) a[q] = valid ? it[1 + 8 * q + nq] : 0.0;
#pragma unroll
                    for (int q = 0; q < REM; ++q) r[q] = valid ? it[1 + 8 * MT + q] : 0.0;
                    // g = 1/(dE^2 - (w + i eta)^2) = (x - i b)/(x^2 + b^2),  x = dE^2 + eta^2 - w^2,  b = -2 w eta
                    // (the NT reciprocal chains first, interleaved by the scheduler; then the MMAs, none waiting on a chain)
                    double inv[NT], xi[NT];
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        const double x = cit - om2[t];
                        inv[t] = fast_rcp3(fma(x, x, b2[t]));
                        xi[t] = x * inv[t];
                    }
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
#pragma unroll
                        for (int q = 0; q < MT; ++q) {
                            dmma884(c1[t][q][0], c1[t][q][1], a[q], xi[t]);
                            dmma884(c2[t][q][0], c2[t][q][1], a[q], inv[t]);
                        }
#pragma unroll
                        for (int q = 0; q < REM; ++q) { r1[t][q] = fma(r[q], xi[t], r1[t][q]); r2[t][q] = fma(r[q], inv[t], r2[t][q]); }
                    }
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && c + kMmaStages < nchunks) issue(c + kMmaStages);
    }
    if (!active) return;
    // epilogue: S_q = sum w_q g = (pr, -b pi); antisymmetric components as they are, symmetric ones times z = om + i eta
    constexpr int NASYM = HERM ? 3 : 6;
    auto store = [&](int f, int q, double pr, double pi) {
        const double om = hw[f];
        const double b = -2.0 * om * eta;
        double *o = block_out + ((size_t)blockIdx.x * ((size_t)G * n_hw) + (size_t)g * n_hw + f) * NACC;
        const double sr = pr, si = -b * pi;
        if (q < NASYM) { o[2 * q] = sr; o[2 * q + 1] = si; }
        else { o[2 * q] = om * sr - eta * si; o[2 * q + 1] = fma(om, si, eta * sr); }
    };
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int q = 0; q < MT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int f = 8 * (tile0 + t) + 2 * kq + h;
                if (f < n_hw) store(f, 8 * q + nq, c1[t][q][h], c2[t][q][h]);
            }
#pragma unroll
        for (int q = 0; q < REM; ++q) {
            double pr = r1[t][q], pi = r2[t][q];
            pr += __shfl_xor_sync(0xffffffffu, pr, 1); pi += __shfl_xor_sync(0xffffffffu, pi, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2); pi += __shfl_xor_sync(0xffffffffu, pi, 2);
            const int f = 8 * (tile0 + t) + nq;
            if (kq == 0 && f < n_hw) store(f, 8 * MT + q, pr, pi);
        }
    }
}

// ------------------------------------------------------------------------------------------------ photo sweep
// 27 complex components = 3 tiles of 8 + 3 remainder; one k-point per TMA stage.  The G k-groups of a CTA take the items of a
// stage in turn in blocks of 4 (item block jb belongs to group jb % G).
// SH: one more frequency tile (index wtiles * NT) is SHARED by all consumer warps — warp w takes it for the item blocks whose
// running number is w modulo the number of warps — and reduced in fixed order at the end.  With 25 tiles (n_hw = 200) twelve
// warps then carry 2 + 1/12 tiles each, 6.25 per SM sub-partition, where 13 warps x 2 tiles leave one sub-partition with 7.
// PAIRED: the producer (bands_sub_kernel) wrote the items of the band pairs (n, l) and (l, n) into consecutive slots.  Their
// factors obey h(l,n; w) = -conj h(n,l; w)  (x -> -conj x in 2x/(x^2 - w^2), x = E_n - E_l - 2 i eta), so with T1, T2 the two items
//   T1 h + T2 h' = [(T1re - T2re) hr - (T1im + T2im) hi] + i [(T1re + T2re) hi + (T1im - T2im) hr]
// — the four MMAs per tile of ONE item serve two, on operands formed with four additions at fragment-load time: half the
// reciprocals and half the multiply-adds of the unpaired sweep.
template <int NT, bool SH, bool PAIRED>
__global__ void __launch_bounds__(SH ? (NT == 3 ? 256 : 384) : 448, 1)   // SH: 12 warps x 2 tiles (168 registers) or 8 warps x 3 tiles (255), no producer warp
sweep_photo_mma_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_items, const long long nk,
                       const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                       const int G, const int wtiles) {
    constexpr int ISTR = kPhotoItem / 2;                          // double2 per item
    constexpr int MT = 3, REM = 3;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const size_t item_bytes = sizeof(double2) * (size_t)max_items * ISTR;
    const size_t stage_bytes = (item_bytes + 16 + 127) / 128 * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm_raw);
    unsigned char *stage0 = sm_raw + 128;
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    const int nchunks = kb0 < kb1 ? (int)(kb1 - kb0) : 0;
    // consumer warps; without SH the last warp is a dedicated TMA producer, with SH thread 0 refills a stage one k-point after
    // it was consumed (by then the other warps have normally released it)
    const int n_cons = SH ? (int)(blockDim.x >> 5) : (int)(blockDim.x >> 5) - 1;
    uint64_t *full = bar, *empty = bar + kMmaPhotoStages;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kMmaPhotoStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], n_cons); }
        mbar_fence_init();
    }
    __syncthreads();
    auto produce = [&](int c) {
        const int st = c % kMmaPhotoStages;
        if (c >= kMmaPhotoStages) mbar_wait(&empty[st], (uint32_t)(((c / kMmaPhotoStages) + 1) & 1));
        unsigned char *dst = stage0 + (size_t)st * stage_bytes;
        const long long k = kb0 + c;
        mbar_expect_tx(&full[st], (uint32_t)item_bytes + 16u);
        tma_bulk_g2s(dst, items + (size_t)k * max_items * ISTR, (uint32_t)item_bytes, &full[st]);
        tma_bulk_g2s(dst + item_bytes, cnt + (k & ~3LL), 16u, &full[st]);
    };
    if (!SH && (int)(threadIdx.x >> 5) == n_cons) {             // producer warp
        if ((threadIdx.x & 31) == 0)
            for (int c = 0; c < nchunks; ++c) produce(c);
        return;
    }
    if (SH && threadIdx.x == 0)
        for (int c = 0; c < kMmaPhotoStages && c < nchunks; ++c) produce(c);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int kq = lane & 3, nq = lane >> 2;
    const int g = warp / wtiles;
    const int tile0 = ((warp % wtiles) + blockIdx.y * wtiles) * NT;
    const bool active = g < G;
    const double e2 = 2.0 * eta, e22 = e2 * e2;
    double om2[NT], iw2[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        const int f = 8 * (tile0 + t) + nq;
        const double om = (f < n_hw) ? hw[f] : 1.0;               // padding frequencies: any finite value
        om2[t] = om * om; iw2[t] = 1.0 / om2[t];                  // hw = 0 -> inf/NaN like photo.f90:70 (Q13)
    }
    double cr[NT][MT][2], ci[NT][MT][2], rr[NT][REM], ri[NT][REM];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int q = 0; q < MT; ++q) { cr[t][q][0] = cr[t][q][1] = ci[t][q][0] = ci[t][q][1] = 0.0; }
#pragma unroll
        for (int q = 0; q < REM; ++q) { rr[t][q] = 0.0; ri[t][q] = 0.0; }
    }
    // shared tile (SH): its own frequency constants and accumulators
    const int sh_tile = wtiles * NT;
    double som2 = 1.0, siw2 = 1.0, scr[MT][2], sci[MT][2], srr[REM], sri[REM];
    if (SH) {
        const int f = 8 * sh_tile + nq;
        const double om = (f < n_hw) ? hw[f] : 1.0;
        som2 = om * om; siw2 = 1.0 / som2;
    }
#pragma unroll
    for (int q = 0; q < MT; ++q) { scr[q][0] = scr[q][1] = sci[q][0] = sci[q][1] = 0.0; }
#pragma unroll
    for (int q = 0; q < REM; ++q) { srr[q] = 0.0; sri[q] = 0.0; }
    int sh_turn = warp;                                           // item blocks until this warp's next turn at the shared tile
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % kMmaPhotoStages;
        if (SH && threadIdx.x == 0 && c >= 1 && c - 1 + kMmaPhotoStages < nchunks) produce(c - 1 + kMmaPhotoStages);
        if (SH) __syncwarp();
        mbar_wait(&full[st], (uint32_t)((c / kMmaPhotoStages) & 1));
        if (active) {
            const unsigned char *base = stage0 + (size_t)st * stage_bytes;
            const double2 *its = reinterpret_cast<const double2 *>(base);
            const int n_it = reinterpret_cast<const int *>(base + item_bytes)[(kb0 + c) & 3];
            const int n_units = PAIRED ? (n_it >> 1) : n_it;                   // contraction units: merged pairs or single items
            for (int j0 = 4 * g; j0 < n_units; j0 += 4 * G) {
                const bool valid = (j0 + kq) < n_units;
                const int ju = valid ? j0 + kq : 0;
                const double2 *it = its + (size_t)(PAIRED ? 2 * ju : ju) * ISTR;
                const double xr = valid ? it[27].x : 1e3;                      // padding units: T = 0, finite h
                // operands: A1 pairs with hr (real part), A2 with -hi (real part), A3 with hi (imaginary), A4 with hr (imaginary)
                double A1[MT], A2[MT], A3[MT], A4[MT], R1[REM], R2[REM], R3[REM], R4[REM];
#pragma unroll
                for (int q = 0; q < MT + REM; ++q) {
                    const int comp = (q < MT) ? 8 * q + nq : 8 * MT + (q - MT);
                    const double2 t1 = valid ? it[comp] : make_double2(0.0, 0.0);
                    double o1 = t1.x, o2 = t1.y, o3 = t1.x, o4 = t1.y;
                    if (PAIRED) {
                        const double2 t2 = valid ? it[ISTR + comp] : make_double2(0.0, 0.0);
                        o1 = t1.x - t2.x; o2 = t1.y + t2.y; o3 = t1.x + t2.x; o4 = t1.y - t2.y;
                    }
                    if (q < MT) { A1[q] = o1; A2[q] = o2; A3[q] = o3; A4[q] = o4; }
                    else { R1[q - MT] = o1; R2[q - MT] = o2; R3[q - MT] = o3; R4[q - MT] = o4; }
                }
                const double xx = xr * xr - e22, bb = -2.0 * xr * e2, b2 = bb * bb;
                // h = (1/w^2) * 2x/(x^2 - w^2),  x = xr - 2 i eta   (all NT chains first, then the MMAs)
                double hr[NT], hi[NT], nhi[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const double aa = xx - om2[t];
                    const double inv = fast_rcp(fma(aa, aa, b2)) * iw2[t];
                    hr[t] = 2.0 * (xr * aa - e2 * bb) * inv; hi[t] = -2.0 * (xr * bb + e2 * aa) * inv;
                    nhi[t] = -hi[t];
                }
#pragma unroll
                for (int t = 0; t < NT; ++t) {
#pragma unroll
                    for (int q = 0; q < MT; ++q) {                             // dependent pairs (same accumulator) kept 6 MMAs apart
                        dmma884(cr[t][q][0], cr[t][q][1], A1[q], hr[t]);
                        dmma884(ci[t][q][0], ci[t][q][1], A3[q], hi[t]);
                    }
#pragma unroll
                    for (int q = 0; q < MT; ++q) {
                        dmma884(cr[t][q][0], cr[t][q][1], A2[q], nhi[t]);
                        dmma884(ci[t][q][0], ci[t][q][1], A4[q], hr[t]);
                    }
#pragma unroll
                    for (int q = 0; q < REM; ++q) {
                        rr[t][q] = fma(R1[q], hr[t], rr[t][q]); rr[t][q] = fma(R2[q], nhi[t], rr[t][q]);
                        ri[t][q] = fma(R3[q], hi[t], ri[t][q]); ri[t][q] = fma(R4[q], hr[t], ri[t][q]);
                    }
                }
                if (SH) {
                    if (sh_turn == 0) {                                        // this unit block's shared-tile contribution is ours
                        sh_turn = n_cons;
                        const double aa = xx - som2;
                        const double inv = fast_rcp(fma(aa, aa, b2)) * siw2;
                        const double shr = 2.0 * (xr * aa - e2 * bb) * inv, shi = -2.0 * (xr * bb + e2 * aa) * inv, snhi = -shi;
#pragma unroll
                        for (int q = 0; q < MT; ++q) {
                            dmma884(scr[q][0], scr[q][1], A1[q], shr);
                            dmma884(sci[q][0], sci[q][1], A3[q], shi);
                        }
#pragma unroll
                        for (int q = 0; q < MT; ++q) {
                            dmma884(scr[q][0], scr[q][1], A2[q], snhi);
                            dmma884(sci[q][0], sci[q][1], A4[q], shr);
                        }
#pragma unroll
                        for (int q = 0; q < REM; ++q) {
                            srr[q] = fma(R1[q], shr, srr[q]); srr[q] = fma(R2[q], snhi, srr[q]);
                            sri[q] = fma(R3[q], shi, sri[q]); sri[q] = fma(R4[q], shr, sri[q]);
                        }
                    }
                    --sh_turn;
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);                    // this warp is done with stage st
    }
    if (!active) return;
    double2 *out = reinterpret_cast<double2 *>(block_out) + ((size_t)blockIdx.x * ((size_t)G * n_hw) + (size_t)g * n_hw) * (kPhotoAcc / 2);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
#pragma unroll
        for (int q = 0; q < MT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int f = 8 * (tile0 + t) + 2 * kq + h;
                if (f < n_hw) out[(size_t)f * (kPhotoAcc / 2) + 8 * q + nq] = make_double2(cr[t][q][h], ci[t][q][h]);
            }
#pragma unroll
        for (int q = 0; q < REM; ++q) {
            double pr = rr[t][q], pi = ri[t][q];
            pr += __shfl_xor_sync(0xffffffffu, pr, 1); pi += __shfl_xor_sync(0xffffffffu, pi, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2); pi += __shfl_xor_sync(0xffffffffu, pi, 2);
            const int f = 8 * (tile0 + t) + nq;
            if (kq == 0 && f < n_hw) out[(size_t)f * (kPhotoAcc / 2) + 8 * MT + q] = make_double2(pr, pi);
        }
    }
    if (SH) {
        // partial sums of the shared tile -> shared memory (the stage ring is free once every consumer warp is here),
        // summed over the warps in fixed order (bit-reproducible), written by the first 8 x 27 threads
        constexpr int NC = kPhotoAcc / 2;                          // 27 complex components
        double2 *scratch = reinterpret_cast<double2 *>(stage0);    // [n_cons][8][NC]
        const int nthr = n_cons * 32;
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        double2 *mine = scratch + (size_t)warp * 8 * NC;
#pragma unroll
        for (int q = 0; q < MT; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) mine[(2 * kq + h) * NC + 8 * q + nq] = make_double2(scr[q][h], sci[q][h]);
#pragma unroll
        for (int q = 0; q < REM; ++q) {
            double pr = srr[q], pi = sri[q];
            pr += __shfl_xor_sync(0xffffffffu, pr, 1); pi += __shfl_xor_sync(0xffffffffu, pi, 1);
            pr += __shfl_xor_sync(0xffffffffu, pr, 2); pi += __shfl_xor_sync(0xffffffffu, pi, 2);
            if (kq == 0) mine[nq * NC + 8 * MT + q] = make_double2(pr, pi);
        }
        asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
        for (int e = threadIdx.x; e < 8 * NC; e += nthr) {
            double2 sum = make_double2(0.0, 0.0);
            for (int w = 0; w < n_cons; ++w) { const double2 v = scratch[(size_t)w * 8 * NC + e]; sum.x += v.x; sum.y += v.y; }
            const int f = 8 * sh_tile + e / NC;
            if (f < n_hw) out[(size_t)f * NC + e % NC] = sum;
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------ shapes
// ntile = ceil(n_hw / 8) frequency tiles; a warp owns NT of them -> wt = ceil(ntile / NT) warps cover every frequency once.
// wt <= wmax: G = wmax / wt k-groups share a CTA of G wt warps; else the frequencies are split over grid.y passes.
// shared_ok (photo sweep, one CTA per SM): efficiency = useful tiles / (4 x the busiest SM sub-partition's tiles), warps
// being dealt round-robin to the four sub-partitions, and the shape "W warps x NT tiles + one tile shared by all warps" is
// also tried.
MmaShape sweep_mma_shape(int n_hw, int wmax, int nt_max, bool shared_ok) {
    MmaShape best{}; double best_eff = -1.0;
    const int ntile = (n_hw + 7) / 8;
    if (const char *e = getenv("MEPB_MMA_NT")) { const int v = atoi(e); if (v >= 1 && v < nt_max) nt_max = v; }   // tuning knob
    static const bool no_shared = getenv("MEPB_MMA_NOSHARED") != nullptr;
    for (int NT : {4, 3, 2, 1}) {
        if (NT > nt_max || (NT == 3 && !shared_ok)) continue;
        MmaShape s{};
        s.NT = NT; s.shared = 0;
        const int wt = (ntile + NT - 1) / NT;
        if (wt <= wmax) { s.passes = 1; s.wtiles = wt; s.G = wmax / wt; }
        else { s.passes = (wt + wmax - 1) / wmax; s.wtiles = (wt + s.passes - 1) / s.passes; s.G = 1; }
        s.warps = s.G * s.wtiles;
        double busiest = 0.0;
        for (int y = 0; y < s.passes; ++y) {
            int load[4] = {0, 0, 0, 0};
            for (int w = 0; w < s.warps; ++w) {
                const int first = ((w % s.wtiles) + y * s.wtiles) * NT;
                const int valid = ntile - first < 0 ? 0 : (ntile - first > NT ? NT : ntile - first);
                load[w & 3] += valid;
            }
            busiest += (double)std::max(std::max(load[0], load[1]), std::max(load[2], load[3]));
        }
        // one CTA per SM (photo sweep): the busiest sub-partition sets the pace; several co-resident CTAs and grid.y passes
        // (AHC sweep) even that out, there the plain tile-slot utilisation ranks the shapes
        const double eff = (shared_ok ? (double)ntile * s.G / (4.0 * busiest) : (double)n_hw * s.G / (8.0 * s.passes * s.warps * NT)) *
                           (s.warps >= 4 ? 1.0 : 0.25 * s.warps);
        if (NT != 3 && eff > best_eff + 1e-9) { best_eff = eff; best = s; }       // three tiles per warp only in the shared shape
        if (shared_ok && !no_shared && (ntile - 1) % NT == 0) {
            const int W = (ntile - 1) / NT;
            if (W >= 4 && W <= (NT == 3 ? 8 : wmax - 1) && W % 4 == 0) {
                const double effs = (double)ntile / (W * NT + 1.0);
                if (effs > best_eff + 1e-9) { best_eff = effs; best = MmaShape{NT, 1, W, W, 1, 1}; }
            }
        }
    }
    return best;
}
int sweep_ahc_mma_bps(bool herm, int NT) { return (herm && NT <= 2) ? 3 : 2; }   // resident CTAs per SM (launch bounds above)
bool sweep_photo_mma_fits(int max_items) {      // the ring of kMmaPhotoStages one-k-point stages must fit one CTA's shared memory
    const size_t stage_bytes = (sizeof(double) * kPhotoItem * (size_t)max_items + 16 + 127) / 128 * 128;
    return 128 + kMmaPhotoStages * stage_bytes <= 200 * 1024;
}
bool sweep_mma_enabled(int n_hw) {
    static const bool off = getenv("MEPB_SWEEP_FMA") != nullptr;
    return !off && n_hw >= 16;
}

cudaError_t launch_sweep_ahc_mma(bool herm, const double *items, const int *cnt, int max_pairs, long long nk, int n_hw,
                                 const double *hw, double eta, double *block_out, int grid, int KC, cudaStream_t st) {
    const MmaShape s = sweep_mma_shape(n_hw, kMmaAhcWarps, herm ? 4 : 2, false);   // general items at NT = 4 spill (96-register cap)
    const size_t item_bytes = sizeof(double) * (herm ? kAhcItemH : kAhcItemG) * (size_t)max_pairs * KC;
    const size_t stage_bytes = (item_bytes + 16 * ((KC * sizeof(int) + 15) / 16) + 127) / 128 * 128;
    const size_t smem = 128 + kMmaStages * stage_bytes;
    const dim3 g(grid, s.passes), b(32 * s.warps);
    const double2 *it2 = reinterpret_cast<const double2 *>(items);
#define MEPB_LAUNCH_AHC(H, N)                                                                                                   \
    do {                                                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(sweep_ahc_mma_kernel<H, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                         \
        sweep_ahc_mma_kernel<H, N><<<g, b, smem, st>>>(it2, cnt, max_pairs, nk, n_hw, hw, eta, block_out, s.G, s.wtiles, KC);    \
    } while (0)
    if (herm) { if (s.NT == 4) MEPB_LAUNCH_AHC(true, 4); else if (s.NT == 2) MEPB_LAUNCH_AHC(true, 2); else MEPB_LAUNCH_AHC(true, 1); }
    else { if (s.NT == 4) MEPB_LAUNCH_AHC(false, 4); else if (s.NT == 2) MEPB_LAUNCH_AHC(false, 2); else MEPB_LAUNCH_AHC(false, 1); }
#undef MEPB_LAUNCH_AHC
    return cudaGetLastError();
}

cudaError_t launch_sweep_photo_mma(const double *items, const int *cnt, int max_items, long long nk, int n_hw, const double *hw,
                                   double eta, double *block_out, int grid, bool paired, cudaStream_t st) {
    const MmaShape s = sweep_mma_shape(n_hw, kMmaPhotoWarps, 3, true);
    const size_t item_bytes = sizeof(double) * kPhotoItem * (size_t)max_items;
    const size_t stage_bytes = (item_bytes + 16 + 127) / 128 * 128;
    int stages = kMmaPhotoStages;
    const size_t smem = 128 + stages * stage_bytes;              // <= 6 x 64 KB would not fit: sweep_photo_mma_fits() gates this path
    const dim3 g(grid, s.passes), b(32 * (s.warps + (s.shared ? 0 : 1)));   // + the producer warp (none in the shared-tile shape)
    const double2 *it2 = reinterpret_cast<const double2 *>(items);
#define MEPB_LAUNCH_PHOTO(N, S, P)                                                                                                  \
    do {                                                                                                                             \
        cudaError_t e = cudaFuncSetAttribute(sweep_photo_mma_kernel<N, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                                                              \
        sweep_photo_mma_kernel<N, S, P><<<g, b, smem, st>>>(it2, cnt, max_items, nk, n_hw, hw, eta, block_out, s.G, s.wtiles);        \
    } while (0)
#define MEPB_LAUNCH_PHOTO_P(N, S) do { if (paired) MEPB_LAUNCH_PHOTO(N, S, true); else MEPB_LAUNCH_PHOTO(N, S, false); } while (0)
    if (s.shared) { if (s.NT == 3) MEPB_LAUNCH_PHOTO_P(3, true); else if (s.NT == 2) MEPB_LAUNCH_PHOTO_P(2, true); else MEPB_LAUNCH_PHOTO_P(1, true); }
    else { if (s.NT == 2) MEPB_LAUNCH_PHOTO_P(2, false); else MEPB_LAUNCH_PHOTO_P(1, false); }
#undef MEPB_LAUNCH_PHOTO_P
#undef MEPB_LAUNCH_PHOTO
    return cudaGetLastError();
}

}  // namespace mepb
