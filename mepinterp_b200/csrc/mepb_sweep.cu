// mepb_sweep.cu — frequency-sweep kernels: consume the "pair items" written by the interpolation kernels and
// accumulate velo_ahc_tens / kubo_ohc_tens (kubo.f90:86-198) and photo_2nd_cond (photo.f90:16-86) for every hw.
//
// One thread owns one frequency (and one of G k-groups); its accumulators live in registers for the whole launch:
//   out[comp, w] += sum_items weight[item, comp] * g[item](w)
// i.e. a (components x items) . (items x n_hw) contraction whose right-hand factor is generated on the fly
// (one complex reciprocal per (item, w)).  FP64 FMA pipe bound; items are read once from HBM/L2 (broadcast loads).
#include <cstdlib>

#include "mepb_internal.cuh"

namespace mepb {

namespace {

// ------------------------------------------------------------------------------------------------ AHC / OHC sweep
// HERM: Hermitian items (10 doubles) -> accumulators X2 (3 complex), Y1 (6 complex)            = 18 doubles
// else: general items (20 doubles)   -> accumulators X1, X2 (3 complex each), Y1, Y2 (6 each) = 36 doubles
template <bool HERM>
__global__ void __launch_bounds__(256, 2)
sweep_ahc_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_pairs, const long long nk,
                 const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                 const int G, const int NS) {
    constexpr int NACC = HERM ? kSweepAccH : kSweepAcc;
    constexpr int ISTR = (HERM ? kAhcItemH : kAhcItemG) / 2;     // item stride in double2
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    for (int s = threadIdx.x; s < NS; s += blockDim.x) {
        const int w = s % n_hw, g = s / n_hw;
        const double om = hw[w];
        const double om2 = om * om, b = -2.0 * om * eta, b2 = b * b;
        double acc[NACC];
#pragma unroll
        for (int q = 0; q < NACC; ++q) acc[q] = 0.0;
        for (long long k = kb0 + g; k < kb1; k += G) {
            const int c = cnt[k];
            const double2 *it = items + (size_t)k * max_pairs * ISTR;
#pragma unroll 2
            for (int j = 0; j < c; ++j, it += ISTR) {
                double v[2 * ISTR];
#pragma unroll
                for (int q = 0; q < ISTR; ++q) { const double2 t = __ldg(it + q); v[2 * q] = t.x; v[2 * q + 1] = t.y; }
                // g = 1/(dE^2 - (w + i eta)^2) = 1/(a + i b),  a = dE^2 + eta^2 - w^2,  b = -2 w eta
                const double a = v[0] - om2;
                const double inv = fast_rcp(fma(a, a, b2));
                const double gr = a * inv, gi = -b * inv;
                const double zr = om * gr - eta * gi, zi = fma(om, gi, eta * gr);      // z g
                if (HERM) {
#pragma unroll
                    for (int q = 0; q < 3; ++q) { acc[2 * q] = fma(v[1 + q], gr, acc[2 * q]); acc[2 * q + 1] = fma(v[1 + q], gi, acc[2 * q + 1]); }
#pragma unroll
                    for (int q = 0; q < 6; ++q) { acc[6 + 2 * q] = fma(v[4 + q], zr, acc[6 + 2 * q]); acc[7 + 2 * q] = fma(v[4 + q], zi, acc[7 + 2 * q]); }
                } else {
#pragma unroll
                    for (int q = 0; q < 6; ++q) { acc[2 * q] = fma(v[1 + q], gr, acc[2 * q]); acc[2 * q + 1] = fma(v[1 + q], gi, acc[2 * q + 1]); }
#pragma unroll
                    for (int q = 0; q < 12; ++q) { acc[12 + 2 * q] = fma(v[7 + q], zr, acc[12 + 2 * q]); acc[13 + 2 * q] = fma(v[7 + q], zi, acc[13 + 2 * q]); }
                }
            }
        }
        double *o = block_out + ((size_t)blockIdx.x * NS + s) * NACC;
#pragma unroll
        for (int q = 0; q < NACC; ++q) o[q] = acc[q];
    }
}

// ------------------------------------------------------------------------------------------------ staged AHC / OHC sweep
// Same contraction, but the items of KC consecutive k-points are streamed into shared memory by TMA bulk copies
// (3-stage mbarrier pipeline, one elected producer thread) and read by every frequency-thread as broadcast LDS: the
// direct-load kernel above stalls 57 % of its samples on the item loads (profiles/r01a_*).  With g = (a - i b)/(a^2+b^2):
//   sum_items w g = ( sum w a inv , -b sum w inv ),  inv = 1/(a^2+b^2)   ->  2 FMAs per component, no per-item z g
// and the factor z = w + i eta of the symmetric part is applied once per thread at the end.
constexpr int kSweepStages = 3;

template <bool HERM>
__global__ void __launch_bounds__(256, HERM ? 3 : 2)
sweep_ahc_staged_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_pairs, const long long nk,
                        const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                        const int G, const int NS, const int KC) {
    constexpr int NACC = HERM ? kSweepAccH : kSweepAcc;
    constexpr int NW_ = NACC / 2;                                // weights per item: 9 (Hermitian) or 18
    constexpr int ISTR = (HERM ? kAhcItemH : kAhcItemG) / 2;     // item stride in double2
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const size_t item_bytes = sizeof(double2) * (size_t)KC * max_pairs * ISTR;
    const size_t stage_bytes = (item_bytes + 16 * ((KC * sizeof(int) + 15) / 16) + 127) / 128 * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm_raw);        // [kSweepStages]
    unsigned char *stage0 = sm_raw + 128;
    // block's k range, aligned to KC so that every TMA source is 16-byte aligned
    long long per = (nk + gridDim.x - 1) / gridDim.x;
    per = (per + KC - 1) / KC * KC;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    const int nchunks = kb0 < kb1 ? (int)((kb1 - kb0 + KC - 1) / KC) : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSweepStages; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {   // elected thread: chunk c -> stage c % S
        unsigned char *dst = stage0 + (size_t)(c % kSweepStages) * stage_bytes;
        const long long k0 = kb0 + (long long)c * KC;
        const uint32_t cb = (uint32_t)(16 * ((KC * sizeof(int) + 15) / 16));
        mbar_expect_tx(&bar[c % kSweepStages], (uint32_t)item_bytes + cb);
        tma_bulk_g2s(dst, items + (size_t)k0 * max_pairs * ISTR, (uint32_t)item_bytes, &bar[c % kSweepStages]);
        tma_bulk_g2s(dst + item_bytes, cnt + k0, cb, &bar[c % kSweepStages]);
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < kSweepStages && c < nchunks; ++c) issue(c);

    const int s_ = blockIdx.y * blockDim.x + threadIdx.x;       // grid.y passes cover NS > blockDim.x (n_hw > 256)
    const bool active = s_ < NS;
    const int w = active ? s_ % n_hw : 0, g = active ? s_ / n_hw : 0;
    const double om = hw[w];
    const double om2 = om * om, b = -2.0 * om * eta, b2 = b * b;
    double pr[NW_], pi[NW_];
#pragma unroll
    for (int q = 0; q < NW_; ++q) { pr[q] = 0.0; pi[q] = 0.0; }
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % kSweepStages;
        mbar_wait(&bar[st], (uint32_t)((c / kSweepStages) & 1));
        if (active) {
            const unsigned char *base = stage0 + (size_t)st * stage_bytes;
            const double2 *its = reinterpret_cast<const double2 *>(base);
            const int *cn = reinterpret_cast<const int *>(base + item_bytes);
            const long long k0 = kb0 + (long long)c * KC;
            const int kcv = (int)((kb1 - k0 < KC) ? (kb1 - k0) : KC);
            for (int kk = g; kk < kcv; kk += G) {
                const int n_it = cn[kk];
                const double2 *it = its + (size_t)kk * max_pairs * ISTR;
#pragma unroll 2
                for (int j = 0; j < n_it; ++j, it += ISTR) {
                    double v[2 * ISTR];
#pragma unroll
                    for (int q = 0; q < ISTR; ++q) { const double2 t = it[q]; v[2 * q] = t.x; v[2 * q + 1] = t.y; }
                    const double a = v[0] - om2;
                    const double inv = fast_rcp3(fma(a, a, b2));
                    const double ai = a * inv;
#pragma unroll
                    for (int q = 0; q < NW_; ++q) { pr[q] = fma(v[1 + q], ai, pr[q]); pi[q] = fma(v[1 + q], inv, pi[q]); }
                }
            }
        }
        __syncthreads();                                          // everyone is done with stage st
        if (threadIdx.x == 0 && c + kSweepStages < nchunks) issue(c + kSweepStages);
    }
    if (!active) return;
    // epilogue: S_q = sum w_q g = (pr, -b pi); antisymmetric components as they are, symmetric ones times z = om + i eta
    double *o = block_out + ((size_t)blockIdx.x * NS + s_) * NACC;
    constexpr int NASYM = HERM ? 3 : 6;
#pragma unroll
    for (int q = 0; q < NW_; ++q) {
        const double sr = pr[q], si = -b * pi[q];
        if (q < NASYM) { o[2 * q] = sr; o[2 * q + 1] = si; }
        else { o[2 * q] = om * sr - eta * si; o[2 * q + 1] = fma(om, si, eta * sr); }
    }
}

// ------------------------------------------------------------------------------------------------ photo sweep
__global__ void __launch_bounds__(128, 2)
sweep_photo_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_items, const long long nk,
                   const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                   const int G, const int NS) {
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    for (int s = threadIdx.x; s < NS; s += blockDim.x) {
        const int w = s % n_hw, g = s / n_hw;
        const double om = hw[w];
        const double om2 = om * om;
        const double iw2 = 1.0 / om2;                 // hw = 0 -> inf/NaN like photo.f90:70 (Q13)
        const double e2 = 2.0 * eta;
        double2 acc[27];
#pragma unroll
        for (int q = 0; q < 27; ++q) acc[q] = make_double2(0.0, 0.0);
        for (long long k = kb0 + g; k < kb1; k += G) {
            const int c = cnt[k];
            const double2 *it = items + (size_t)k * max_items * (kPhotoItem / 2);
            for (int j = 0; j < c; ++j, it += kPhotoItem / 2) {
                const double xr = __ldg(it + 27).x;
                // h = (1/w^2) * 2x/(x^2 - w^2),  x = xr - 2 i eta
                const double a = xr * xr - e2 * e2 - om2, b = -2.0 * xr * e2;
                const double inv = fast_rcp(fma(a, a, b * b)) * iw2;
                const double hr = 2.0 * (xr * a - e2 * b) * inv, hi = -2.0 * (xr * b + e2 * a) * inv;
#pragma unroll
                for (int q = 0; q < 27; ++q) {
                    const double2 t = __ldg(it + q);
                    acc[q].x = fma(t.x, hr, acc[q].x); acc[q].x = fma(-t.y, hi, acc[q].x);
                    acc[q].y = fma(t.x, hi, acc[q].y); acc[q].y = fma(t.y, hr, acc[q].y);
                }
            }
        }
        double2 *o = reinterpret_cast<double2 *>(block_out + ((size_t)blockIdx.x * NS + s) * kPhotoAcc);
#pragma unroll
        for (int q = 0; q < 27; ++q) o[q] = acc[q];
    }
}

// ------------------------------------------------------------------------------------------------ staged photo sweep
// One k-point (nw(nw-1) items x 448 bytes) per TMA stage; up to 256 threads per CTA, one CTA per SM (178 registers; two CTAs
// of 224 threads at 128 registers measured 10 % slower, profiles/r01j): thread = (frequency, item group g of G) — the G
// groups share a stage and split its items (grid.y passes for n_hw > 256).
constexpr int kPhotoThreads = 256;
__global__ void __launch_bounds__(kPhotoThreads, 1)
sweep_photo_staged_kernel(const double2 *__restrict__ items, const int *__restrict__ cnt, const int max_items, const long long nk,
                          const int n_hw, const double *__restrict__ hw, const double eta, double *__restrict__ block_out,
                          const int G, const int NS) {
    constexpr int ISTR = kPhotoItem / 2;
    extern __shared__ __align__(128) unsigned char sm_raw[];
    const size_t item_bytes = sizeof(double2) * (size_t)max_items * ISTR;
    const size_t stage_bytes = (item_bytes + 16 + 127) / 128 * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm_raw);
    unsigned char *stage0 = sm_raw + 128;
    const long long per = (nk + gridDim.x - 1) / gridDim.x;
    const long long kb0 = (long long)blockIdx.x * per;
    const long long kb1 = (kb0 + per < nk) ? kb0 + per : nk;
    const int nchunks = kb0 < kb1 ? (int)(kb1 - kb0) : 0;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSweepStages; ++s) mbar_init(&bar[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int c) {
        unsigned char *dst = stage0 + (size_t)(c % kSweepStages) * stage_bytes;
        const long long k = kb0 + c;
        mbar_expect_tx(&bar[c % kSweepStages], (uint32_t)item_bytes + 16u);
        tma_bulk_g2s(dst, items + (size_t)k * max_items * ISTR, (uint32_t)item_bytes, &bar[c % kSweepStages]);
        tma_bulk_g2s(dst + item_bytes, cnt + (k & ~3LL), 16u, &bar[c % kSweepStages]);      // the aligned 4 counts around k
    };
    if (threadIdx.x == 0)
        for (int c = 0; c < kSweepStages && c < nchunks; ++c) issue(c);
    const int s_ = blockIdx.y * blockDim.x + threadIdx.x;
    const bool active = s_ < NS;
    const int w = active ? s_ % n_hw : 0, g = active ? s_ / n_hw : 0;
    const double om = hw[w];
    const double om2 = om * om;
    const double iw2 = 1.0 / om2;                 // hw = 0 -> inf/NaN like photo.f90:70 (Q13)
    const double e2 = 2.0 * eta;
    double2 acc[27];
#pragma unroll
    for (int q = 0; q < 27; ++q) acc[q] = make_double2(0.0, 0.0);
    for (int c = 0; c < nchunks; ++c) {
        const int st = c % kSweepStages;
        mbar_wait(&bar[st], (uint32_t)((c / kSweepStages) & 1));
        if (active) {
            const unsigned char *base = stage0 + (size_t)st * stage_bytes;
            const double2 *it = reinterpret_cast<const double2 *>(base) + (size_t)g * ISTR;
            const int n_it = reinterpret_cast<const int *>(base + item_bytes)[(kb0 + c) & 3];
            for (int j = g; j < n_it; j += G, it += (size_t)G * ISTR) {
                const double xr = it[27].x;
                // h = (1/w^2) * 2x/(x^2 - w^2),  x = xr - 2 i eta
                const double a = xr * xr - e2 * e2 - om2, b = -2.0 * xr * e2;
                const double inv = fast_rcp(fma(a, a, b * b)) * iw2;
                const double hr = 2.0 * (xr * a - e2 * b) * inv, hi = -2.0 * (xr * b + e2 * a) * inv;
#pragma unroll
                for (int q = 0; q < 27; ++q) {
                    const double2 t = it[q];
                    acc[q].x = fma(t.x, hr, acc[q].x); acc[q].x = fma(-t.y, hi, acc[q].x);
                    acc[q].y = fma(t.x, hi, acc[q].y); acc[q].y = fma(t.y, hr, acc[q].y);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0 && c + kSweepStages < nchunks) issue(c + kSweepStages);
    }
    if (!active) return;
    double2 *o = reinterpret_cast<double2 *>(block_out + ((size_t)blockIdx.x * NS + s_) * kPhotoAcc);
#pragma unroll
    for (int q = 0; q < 27; ++q) o[q] = acc[q];
}

// ------------------------------------------------------------------------------------------------ deterministic column sums
// out_accum[(c/group)*out_stride + out_off + c%group] += sum_r in[r*cols + c]; one block per column, fixed tree order.
__global__ void __launch_bounds__(256) reduce_rows_kernel(const double *__restrict__ in, const int rows, const int cols,
                                                          double *__restrict__ out_accum, const int group,
                                                          const int out_stride, const int out_off) {
    __shared__ double sm[256];
    for (int c = blockIdx.x; c < cols; c += gridDim.x) {
        double s = 0.0;
        for (int r = threadIdx.x; r < rows; r += 256) s += in[(size_t)r * cols + c];
        sm[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
            __syncthreads();
        }
        if (threadIdx.x == 0) out_accum[(size_t)(c / group) * out_stride + out_off + c % group] += sm[0];
        __syncthreads();
    }
}

void sweep_shape(int n_hw, int T, int *G, int *NS, int *threads) {
    int g = T / n_hw; if (g < 1) g = 1;
    *G = g; *NS = g * n_hw;
    int t = (*NS + 31) / 32 * 32; if (t > T) t = T;
    *threads = t;
}

}  // namespace

int sweep_ahc_kc(bool herm, int max_pairs);
int sweep_ahc_grid(int sm_count, long long nk, int n_hw, int *slots_out, bool herm, int max_pairs) {
    if (sweep_mma_enabled(n_hw) && !getenv("MEPB_SWEEP_DIRECT") && sweep_ahc_kc(herm, max_pairs) > 0) {
        const MmaShape s = sweep_mma_shape(n_hw, kMmaAhcWarps, herm ? 4 : 2, false);
        if (slots_out) *slots_out = s.G * n_hw;
        long long g = (long long)sm_count * sweep_ahc_mma_bps(herm, s.NT);
        const long long need = (nk + s.G - 1) / s.G;
        if (g > need) g = need;
        return (int)(g < 1 ? 1 : g);
    }
    int G, NS, T; sweep_shape(n_hw, 256, &G, &NS, &T);
    if (slots_out) *slots_out = NS;
    long long g = (long long)sm_count * (herm ? 3 : 2);         // resident CTAs per SM of the staged kernel (80 / 128 registers)
    long long need = (nk + G - 1) / G;      // at least G k-points per block to give every group work
    if (g > need) g = need;
    if (g < 1) g = 1;
    return (int)g;
}
// k-points per TMA stage of the staged sweep (0: items of one k do not fit -> direct-load kernel)
int sweep_ahc_kc(bool herm, int max_pairs) {
    const size_t per_k = sizeof(double) * (herm ? kAhcItemH : kAhcItemG) * (size_t)max_pairs;
    if (per_k == 0 || per_k > 24 * 1024) return 0;
    int kc = (int)((20 * 1024) / per_k);
    kc = kc < 4 ? 4 : (kc > 16 ? 16 : kc) / 4 * 4;               // multiple of 4 ints = 16 bytes for the cnt copy
    while ((size_t)kc * per_k > 32 * 1024) kc -= 4;
    return kc < 4 ? 0 : kc;
}

cudaError_t launch_sweep_ahc(bool herm, const double *items, const int *cnt, int max_pairs, long long nk, int n_hw,
                             const double *hw, double eta, double *block_out, int grid, cudaStream_t st) {
    int G, NS, T; sweep_shape(n_hw, 256, &G, &NS, &T);
    const int KC = getenv("MEPB_SWEEP_DIRECT") ? 0 : sweep_ahc_kc(herm, max_pairs);
    if (KC > 0 && sweep_mma_enabled(n_hw))
        return launch_sweep_ahc_mma(herm, items, cnt, max_pairs, nk, n_hw, hw, eta, block_out, grid, KC, st);
    if (KC > 0) {
        const size_t item_bytes = sizeof(double) * (herm ? kAhcItemH : kAhcItemG) * (size_t)max_pairs * KC;
        const size_t stage_bytes = (item_bytes + 16 * ((KC * sizeof(int) + 15) / 16) + 127) / 128 * 128;
        const size_t smem = 128 + kSweepStages * stage_bytes;
        cudaError_t e;
        if (herm) {
            e = cudaFuncSetAttribute(sweep_ahc_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            sweep_ahc_staged_kernel<true><<<dim3(grid, (NS + 255) / 256), 256, smem, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS, KC);
        } else {
            e = cudaFuncSetAttribute(sweep_ahc_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            sweep_ahc_staged_kernel<false><<<dim3(grid, (NS + 255) / 256), 256, smem, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS, KC);
        }
        return cudaGetLastError();
    }
    if (herm)
        sweep_ahc_kernel<true><<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS);
    else
        sweep_ahc_kernel<false><<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_pairs, nk, n_hw, hw, eta, block_out, G, NS);
    return cudaGetLastError();
}
int sweep_photo_grid(int sm_count, long long nk, int n_hw, int *slots_out, int max_items) {
    if (sweep_mma_enabled(n_hw) && !getenv("MEPB_SWEEP_DIRECT") && sweep_photo_mma_fits(max_items)) {
        const MmaShape s = sweep_mma_shape(n_hw, kMmaPhotoWarps, 3, true);
        if (slots_out) *slots_out = s.G * n_hw;
        long long g = sm_count;                                  // one CTA of up to 448 threads per SM
        if (g > nk) g = nk;
        return (int)(g < 1 ? 1 : g);
    }
    int G, NS, T; sweep_shape(n_hw, kPhotoThreads, &G, &NS, &T);   // staged kernel shape; the direct kernel loops over its slots
    if (slots_out) *slots_out = NS;
    long long g = (long long)sm_count * 2;
    if (g > nk) g = nk;
    if (g < 1) g = 1;
    return (int)g;
}
cudaError_t launch_sweep_photo(const double *items, const int *cnt, int max_items, long long nk, int n_hw, const double *hw,
                               double eta, double *block_out, int grid, bool paired, cudaStream_t st) {
    int G, NS, T; sweep_shape(n_hw, kPhotoThreads, &G, &NS, &T);   // same (G, NS) as sweep_photo_grid
    const size_t item_bytes = sizeof(double) * kPhotoItem * (size_t)max_items;
    if (!getenv("MEPB_SWEEP_DIRECT") && sweep_photo_mma_fits(max_items) && sweep_mma_enabled(n_hw))
        return launch_sweep_photo_mma(items, cnt, max_items, nk, n_hw, hw, eta, block_out, grid, paired, st);
    if (!getenv("MEPB_SWEEP_DIRECT") && item_bytes <= 64 * 1024) {
        const size_t stage_bytes = (item_bytes + 16 + 127) / 128 * 128;
        const size_t smem = 128 + kSweepStages * stage_bytes;
        cudaError_t e = cudaFuncSetAttribute(sweep_photo_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sweep_photo_staged_kernel<<<dim3(grid, (NS + T - 1) / T), T, smem, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_items, nk,
                                                                                 n_hw, hw, eta, block_out, G, NS);   // (items in paired order are ordinary items)
        return cudaGetLastError();
    }
    if (T > 128) T = 128;                                        // direct kernel: 128 threads loop over the NS slots
    sweep_photo_kernel<<<grid, T, 0, st>>>(reinterpret_cast<const double2 *>(items), cnt, max_items, nk, n_hw, hw, eta, block_out, G, NS);   // (items in paired order are ordinary items)
    return cudaGetLastError();
}
cudaError_t launch_reduce_rows(const double *in, int rows, int cols, double *out_accum, int group, int out_stride, int out_off,
                               cudaStream_t st) {
    int grid = cols < 4096 ? cols : 4096;
    reduce_rows_kernel<<<grid, 256, 0, st>>>(in, rows, cols, out_accum, group, out_stride, out_off);
    return cudaGetLastError();
}

}  // namespace mepb
