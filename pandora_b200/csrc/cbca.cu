// cbca.cu -- Cross-Based Cost Aggregation: 3x3 NaN-median pre-filter, cross-support arms, aggregation.
//
// Replaces MedianFilter.median_filter (src/pandora/filter/median.py:134-179), cross_support
// (aggregation/cpp/src/aggregation.cpp:224-321) and the per-disparity cbca_step_1..4 loop
// (aggregation.cpp:28-221 driven by aggregation/cbca.py:127-177).
//
// Aggregation semantics, for every disparity k (d = dmin + k) on the interior
// (H', W') = (H - 2*offset, W - 2*offset) view; a column x is "valid" iff 0 <= x + d < W':
//   l,r (per row y')  = min(left arms at (y',x), right-image arms at (y',x+d))
//   Eh[y',x] = sum_{x'=x-l..x+r} c0[y',x']      (c0 = 0 where the cost is NaN),   Nh = l + r
//   t,b (at y)        = min of the up / bottom arms at (y,x) and (y,x+d)
//   E = sum_{y'=y-t..y+b} Eh[y',x],  N = 1 + t + b + sum Nh[y',x]
//   out = (0*c + E) / N   -> NaN wherever the input cost is NaN; non-valid columns: (0*c + 0) / 1.
//
// Kernel: one launch for ALL disparities (the reference makes D Python-level calls).  A CTA owns a
// strip of TX columns x 32 disparities and marches down the rows once: lanes run over the
// disparity axis (coalesced 128-byte reads/writes of the volume), the current row's costs are
// shared through a small shared-memory row buffer for the horizontal arm sums, and the vertical
// stage keeps a running float32 prefix of Eh (exactly the reference's "step 3" column prefix,
// same order of additions) and an int32 prefix of Nh in a shared-memory ring of 4*MA rows, so
// E and N are two differences each.  Every cost is read from HBM once (+ halo columns from L2)
// and written once.
#include "common.cuh"

namespace pb200 {

// ---- 3x3 NaN-aware median -------------------------------------------------------------------------
__device__ __forceinline__ void cswap(float &a, float &b) {
    const float lo = fminf(a, b), hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

__global__ void __launch_bounds__(256) median3_kernel(const float *__restrict__ in, int H, int W, float *__restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int y = (int)(i / W), x = (int)(i % W);
    const float c = in[i];
    if (y < 1 || y >= H - 1 || x < 1 || x >= W - 1 || c != c) {
        out[i] = c;
        return;
    }
    float v[9];
    int m = 0;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const float t = in[(size_t)(y + dy) * W + x + dx];
            const bool ok = (t == t);
            m += ok ? 1 : 0;
            v[(dy + 1) * 3 + dx + 1] = ok ? t : CUDART_INF_F;   // NaNs sort last
        }
    // full sort (36 compare-exchanges): with NaNs present any order statistic may be needed
#pragma unroll
    for (int pass = 0; pass < 8; ++pass)
#pragma unroll
        for (int q = 0; q < 8 - pass; ++q) cswap(v[q], v[q + 1]);
    // median of the m non-NaN values (m >= 1 because the centre is not NaN)
    const int ia = (m - 1) >> 1, ib = m >> 1;
    float a = v[0], b = v[0];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
        a = (q == ia) ? v[q] : a;
        b = (q == ib) ? v[q] : b;
    }
    out[i] = (ia == ib) ? a : (a + b) * 0.5f;                   // even count: mean of the two middles (np.nanmedian)
}

// ---- cross support -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cross_support_kernel(const float *__restrict__ img, int H, int W, int pitch, int len_arms,
                                                            float intensity, int nan_as_inf, short4 *__restrict__ cross) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int y = (int)(i / W), x = (int)(i % W);
    auto px = [&](int yy, int xx) {
        const float v = img[(size_t)yy * pitch + xx];
        return (nan_as_inf && v != v) ? CUDART_INF_F : v;
    };
    const float c = px(y, x);
    short4 o = make_short4(0, 0, 0, 0);
    if (isfinite(c)) {
        int l = 0, r = 0, u = 0, b = 0;
        for (int q = x - 1; q > x - len_arms && q >= 0; --q) {
            if (fabsf(c - px(y, q)) >= intensity) break;
            ++l;
        }
        for (int q = x + 1; q < x + len_arms && q < W; ++q) {
            if (fabsf(c - px(y, q)) >= intensity) break;
            ++r;
        }
        for (int q = y - 1; q > y - len_arms && q >= 0; --q) {
            if (fabsf(c - px(q, x)) >= intensity) break;
            ++u;
        }
        for (int q = y + 1; q < y + len_arms && q < H; ++q) {
            if (fabsf(c - px(q, x)) >= intensity) break;
            ++b;
        }
        if (l < 1 && x >= 1 && isfinite(px(y, x - 1))) l = 1;
        if (r < 1 && x < W - 1 && isfinite(px(y, x + 1))) r = 1;
        if (u < 1 && y >= 1 && isfinite(px(y - 1, x))) u = 1;
        if (b < 1 && y < H - 1 && isfinite(px(y + 1, x))) b = 1;
        o = make_short4((short)l, (short)r, (short)u, (short)b);
    }
    cross[i] = o;
}

// ---- aggregation ---------------------------------------------------------------------------------------
constexpr int CBCA_TX = 8;

// MA = largest possible arm = len_arms - 1 rounded up to {4, 8, 16}.  RAW: write the un-normalised sums E
// (cv_out) and N - 1 (out_n) like the reference's per-disparity cbca() instead of (0*c + E) / N.
template <int MA, bool RAW>
__global__ void __launch_bounds__(32 * CBCA_TX) cbca_aggregate_kernel(const float *__restrict__ cv_in, float *__restrict__ cv_out,
                                                                      float *__restrict__ out_n, int H, int W, int D, int dmin,
                                                                      int off, const short4 *__restrict__ crossL,
                                                                      const short4 *__restrict__ crossR) {
    constexpr int TX = CBCA_TX;
    constexpr int RING = 4 * MA;                 // >= 2*MA + 2 rows, power of two
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *crow = reinterpret_cast<float *>(smem_raw);                       // [TX + 2*MA][32]
    float *PE = crow + (TX + 2 * MA) * 32;                                   // [RING][TX][32]
    int *PN = reinterpret_cast<int *>(PE + RING * TX * 32);                  // [RING][TX][32]

    const int Hi = H - 2 * off, Wi = W - 2 * off;
    const int lane = threadIdx.x, cx = threadIdx.y;
    const int k = blockIdx.y * 32 + lane;
    const int x0 = blockIdx.x * TX;
    const int x = x0 + cx;
    const int d = dmin + k;
    const int xr = x + d;
    const bool active = (k < D) && (x < Wi);
    const bool valid_col = active && xr >= 0 && xr < Wi;
    const int ring_idx = cx * 32 + lane;
    auto cell = [&](int yy, int xx) -> size_t { return ((size_t)(yy + off) * W + (xx + off)) * D + k; };

    float pe_run = 0.f;
    int pn_run = 0;
    for (int i = 0; i < Hi + MA; ++i) {
        // ---- stage 1: horizontal arm sums of row i -----------------------------------------------
        if (i < Hi) {
            for (int jj = cx; jj < TX + 2 * MA; jj += TX) {
                const int xx = x0 - MA + jj;
                float v = 0.f;
                if (k < D && xx >= 0 && xx < Wi) {
                    v = cv_in[cell(i, xx)];
                    if (v != v) v = 0.f;                                    // step 1 does not propagate NaN
                }
                crow[jj * 32 + lane] = v;
            }
        }
        __syncthreads();
        if (i < Hi) {
            float eh = 0.f;
            int nh = 0;
            if (valid_col) {
                const short4 a = crossL[(size_t)i * Wi + x];
                const short4 b = crossR[(size_t)i * Wi + xr];
                const int l = min(min((int)a.x, (int)b.x), MA), r = min(min((int)a.y, (int)b.y), MA);
#pragma unroll
                for (int dx = -MA; dx <= MA; ++dx) {
                    const float v = crow[(cx + MA + dx) * 32 + lane];
                    if (dx >= -l && dx <= r) eh += v;
                }
                nh = l + r;
            }
            pe_run = pe_run + eh;                                           // the reference's step-3 column prefix
            pn_run += nh;
            PE[(i & (RING - 1)) * TX * 32 + ring_idx] = pe_run;
            PN[(i & (RING - 1)) * TX * 32 + ring_idx] = pn_run;
        }
        // ---- stage 2: vertical arm sums of row yo = i - MA ----------------------------------------
        const int yo = i - MA;
        if (yo >= 0 && active) {
            const float c = cv_in[cell(yo, x)];
            float e = 0.f;
            int n = 1;
            if (valid_col) {
                const short4 a = crossL[(size_t)yo * Wi + x];
                const short4 b = crossR[(size_t)yo * Wi + xr];
                const int t = min(min((int)a.z, (int)b.z), MA), bo = min(min((int)a.w, (int)b.w), MA);
                const int r1 = yo + bo, r0 = yo - t - 1;
                const float e1 = PE[(r1 & (RING - 1)) * TX * 32 + ring_idx];
                const int n1 = PN[(r1 & (RING - 1)) * TX * 32 + ring_idx];
                const float e0 = (r0 >= 0) ? PE[(r0 & (RING - 1)) * TX * 32 + ring_idx] : 0.f;
                const int n0 = (r0 >= 0) ? PN[(r0 & (RING - 1)) * TX * 32 + ring_idx] : 0;
                e = e1 - e0;
                n = n1 - n0 + t + bo + 1;
            }
            if (RAW) {
                cv_out[cell(yo, x)] = e;
                out_n[cell(yo, x)] = (float)(n - 1);
            } else {
                cv_out[cell(yo, x)] = (c * 0.f + e) / (float)n;
            }
        }
        __syncthreads();
    }
}

// copy of the `off`-wide border ring (cells the aggregation leaves untouched, cbca.py:173-177)
__global__ void __launch_bounds__(256) cbca_border_kernel(const float *__restrict__ cv_in, float *__restrict__ cv_out, int H, int W,
                                                          int D, int off) {
    // ring pixels enumerated as: top rows, bottom rows, then left/right columns of the middle rows
    const long ring_px = 2L * off * W + 2L * off * (H - 2 * off);
    const long total = ring_px * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long pix = i / D;
        const int kk = (int)(i % D);
        int y, x;
        if (pix < (long)off * W) { y = (int)(pix / W); x = (int)(pix % W); }
        else if (pix < 2L * off * W) { const long q = pix - (long)off * W; y = H - off + (int)(q / W); x = (int)(q % W); }
        else {
            const long q = pix - 2L * off * W;
            y = off + (int)(q / (2 * off));
            const int c = (int)(q % (2 * off));
            x = (c < off) ? c : (W - 2 * off + c);
        }
        const size_t a = ((size_t)y * W + x) * D + kk;
        cv_out[a] = cv_in[a];
    }
}

template <int MA, bool RAW>
static int launch_cbca(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                       const int16_t *cr, cudaStream_t s) {
    constexpr int RING = 4 * MA;
    const size_t smem = (size_t)(CBCA_TX + 2 * MA) * 32 * 4 + 2 * (size_t)RING * CBCA_TX * 32 * 4;
    PB200_CUDA(cudaFuncSetAttribute(cbca_aggregate_kernel<MA, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, CBCA_TX), grid(ceil_div(W - 2 * off, CBCA_TX), ceil_div(D, 32));
    cbca_aggregate_kernel<MA, RAW><<<grid, block, smem, s>>>(in, out, out_n, H, W, D, dmin, off, (const short4 *)cl,
                                                              (const short4 *)cr);
    PB200_LAUNCH_CHECK("cbca_aggregate_kernel");
    return PB200_OK;
}

// shared by pb200_cbca_aggregate (out_n == NULL) and the reference-style per-slice host entry point
int cbca_dispatch(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                  const int16_t *cr, int len_arms, cudaStream_t s) {
    const int ma = len_arms - 1;
#define PB200_C(MA)                                                                                     \
    return out_n ? launch_cbca<MA, true>(in, out, out_n, H, W, D, dmin, off, cl, cr, s)                 \
                 : launch_cbca<MA, false>(in, out, nullptr, H, W, D, dmin, off, cl, cr, s)
    if (ma <= 4) { PB200_C(4); }
    if (ma <= 8) { PB200_C(8); }
    if (ma <= 16) { PB200_C(16); }
#undef PB200_C
    set_error("cbca: cbca_distance %d above the supported maximum (17)", len_arms);
    return PB200_ERR_UNSUPPORTED;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_median3(const float *d_in, int H, int W, float *d_out, void *stream) {
    if (!d_in || !d_out || H <= 0 || W <= 0) {
        set_error("pb200_median3: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    median3_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_in, H, W, d_out);
    PB200_LAUNCH_CHECK("median3_kernel");
    return PB200_OK;
}

extern "C" int pb200_cross_support(const float *d_img, int H, int W, int pitch, int len_arms, float intensity, int nan_as_inf,
                                   int16_t *d_cross, void *stream) {
    if (!d_img || !d_cross || H <= 0 || W <= 0 || pitch < W || len_arms < 1) {
        set_error("pb200_cross_support: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (reinterpret_cast<uintptr_t>(d_cross) & 7) {
        set_error("pb200_cross_support: d_cross must be 8-byte aligned");
        return PB200_ERR_BAD_ARG;
    }
    cross_support_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_img, H, W, pitch, len_arms, intensity,
                                                                                     nan_as_inf, (short4 *)d_cross);
    PB200_LAUNCH_CHECK("cross_support_kernel");
    return PB200_OK;
}

extern "C" int pb200_cbca_aggregate(const float *d_cv_in, float *d_cv_out, int H, int W, int D, int dmin, int offset,
                                    const int16_t *d_cross_left, const int16_t *d_cross_right, int len_arms, void *stream) {
    if (!d_cv_in || !d_cv_out || !d_cross_left || !d_cross_right || H <= 0 || W <= 0 || D <= 0 || offset < 0 ||
        d_cv_in == d_cv_out || len_arms < 1) {
        set_error("pb200_cbca_aggregate: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if ((reinterpret_cast<uintptr_t>(d_cross_left) & 7) || (reinterpret_cast<uintptr_t>(d_cross_right) & 7)) {
        set_error("pb200_cbca_aggregate: supports must be 8-byte aligned");
        return PB200_ERR_BAD_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (offset > 0) {
        const long ring = (2L * offset * W + 2L * offset * (H - 2 * offset)) * D;
        long blocks = (ring + 255) / 256;
        if (blocks > 4096) blocks = 4096;
        if (blocks > 0) {
            cbca_border_kernel<<<(int)blocks, 256, 0, s>>>(d_cv_in, d_cv_out, H, W, D, offset);
            PB200_LAUNCH_CHECK("cbca_border_kernel");
        }
    }
    if (H - 2 * offset <= 0 || W - 2 * offset <= 0) return PB200_OK;
    return cbca_dispatch(d_cv_in, d_cv_out, nullptr, H, W, D, dmin, offset, d_cross_left, d_cross_right, len_arms, s);
}
