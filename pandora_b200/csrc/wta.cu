// wta.cu -- winner-takes-all disparity selection, validity-mask side effects, right cost volume.
//
// Replaces WinnerTakesAll.to_disp / argmin_split / argmax_split (src/pandora/disparity/disparity.py:
// 400-553), mask_invalid_variable_disparity_range + mask_border (criteria.py:291-353) and
// reverse_cost_volume (matching_cost/cpp/src/matching_cost.cpp:26-57).
//
// WTA is HBM-read bound (4*D bytes per pixel in, 4 out): a warp takes four pixels per trip, lanes read 16-byte
// vectors so a warp instruction covers 512 contiguous bytes of a pixel's disparity vector; the (value, index)
// pair is reduced with REDUX on ordered integer keys, lowest index winning ties like np.argmin.
#include <cstdlib>

#include <climits>

#include "common.cuh"

namespace pb200 {

// One element of a lane's scan.  Within a lane the index only grows, so a tie never moves it: one strict comparison (false
// for NaN: NaN never wins, disparity.py:434-444) and two selects; `seen` = NaN-ignoring minimum of everything met, which stays
// NaN exactly when every value was NaN (4 instructions per element; the first version took ~10, and the kernel was bound
// by its instruction issue: profiles/r2_ncu_wta.txt).
template <bool IS_MAX>
__device__ __forceinline__ void wta_take(float v, int k, float &bv, int &bk, float &seen) {
    const bool better = IS_MAX ? (v > bv) : (v < bv);
    bv = better ? v : bv;
    bk = better ? k : bk;
    seen = fminf(seen, v);
}

// float -> unsigned key with the same order (-0 was folded into +0 by the caller)
__device__ __forceinline__ uint32_t wta_key(float v) {
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// WTA_PPW pixels per warp and trip: the loads of all of them are in flight before the first comparison, and the
// (value, index) pair is reduced with two REDUX instructions on ordered integer keys (minimum / maximum key, then the
// lowest index among the lanes that hold it) instead of five rounds of three shuffles.
constexpr int WTA_PPW = 4;

template <bool IS_MAX, bool VEC4>
__global__ void __launch_bounds__(256) wta_kernel(const float *__restrict__ cv, long n_pix, int D, int dmin,
                                                  float invalid_disparity, float *__restrict__ disp,
                                                  uint8_t *__restrict__ all_nan) {
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const float init = IS_MAX ? -CUDART_INF_F : CUDART_INF_F;
    for (long pix0 = warp0 * WTA_PPW; pix0 < n_pix; pix0 += nwarps * WTA_PPW) {
        float bv[WTA_PPW], any[WTA_PPW];
        int bk[WTA_PPW];
#pragma unroll
        for (int q = 0; q < WTA_PPW; ++q) { bv[q] = init; bk[q] = 0x7fffffff; any[q] = nan_f(); }
        if (VEC4) {
            for (int k = lane * 4; k < D; k += 128) {
                float4 v[WTA_PPW];
#pragma unroll
                for (int q = 0; q < WTA_PPW; ++q)
                    v[q] = (pix0 + q < n_pix) ? ld_cs_f4(cv + (pix0 + q) * D + k) : make_float4(nan_f(), nan_f(), nan_f(), nan_f());
#pragma unroll
                for (int q = 0; q < WTA_PPW; ++q) {
                    wta_take<IS_MAX>(v[q].x, k, bv[q], bk[q], any[q]);
                    wta_take<IS_MAX>(v[q].y, k + 1, bv[q], bk[q], any[q]);
                    wta_take<IS_MAX>(v[q].z, k + 2, bv[q], bk[q], any[q]);
                    wta_take<IS_MAX>(v[q].w, k + 3, bv[q], bk[q], any[q]);
                }
            }
        } else {
            for (int k = lane; k < D; k += 32) {
                float v[WTA_PPW];
#pragma unroll
                for (int q = 0; q < WTA_PPW; ++q) v[q] = (pix0 + q < n_pix) ? __ldcs(cv + (pix0 + q) * D + k) : nan_f();
#pragma unroll
                for (int q = 0; q < WTA_PPW; ++q) wta_take<IS_MAX>(v[q], k, bv[q], bk[q], any[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < WTA_PPW; ++q) {
            const uint32_t key = wta_key(bv[q] + 0.0f);                       // -0 == +0 for np.argmin / np.argmax
            const uint32_t best = IS_MAX ? __reduce_max_sync(0xffffffffu, key) : __reduce_min_sync(0xffffffffu, key);
            int k = __reduce_min_sync(0xffffffffu, key == best ? bk[q] : 0x7fffffff);
            const bool anyw = __any_sync(0xffffffffu, any[q] == any[q]);
            // best value still the initial +-inf: every entry is +-inf or NaN, and since NaNs were replaced by
            // the same inf np.argmin / np.argmax return index 0
            if (best == wta_key(init)) k = 0;
            if (lane == q && pix0 + q < n_pix) {
                disp[pix0 + q] = anyw ? (float)(dmin + k) : invalid_disparity;
                if (all_nan) all_nan[pix0 + q] = anyw ? 0 : 1;
            }
        }
    }
}

__global__ void __launch_bounds__(256) validity_mask_kernel(uint16_t *__restrict__ mask, const uint8_t *__restrict__ all_nan,
                                                            int H, int W, int offset, int wta_invalidate) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int y = (int)(i / W), x = (int)(i % W);
    uint16_t m = mask[i];
    const bool missing = all_nan != nullptr && all_nan[i] != 0;
    if (wta_invalidate) {
        if (missing && (m & 0x3C3) == 0) m = 0x3C3;                        // disparity.py:470-474
    } else {
        if (missing && (m & 2) == 0) m += 2;                               // criteria.py:291-322
        if (offset > 0 && (y < offset || y >= H - offset || x < offset || x >= W - offset)) m = 1;  // :325-353
    }
    mask[i] = m;
}

// criteria.validity_mask, no-mask branch (criteria.py:106-147): per-column flags from the disparity range
__global__ void __launch_bounds__(256) validity_init_kernel(uint16_t *__restrict__ mask, int H, int W, int dmin, int dmax, int off) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int x = (int)(i % W);
    uint16_t m = 0;
    if (dmax < 0) {
        if (x + dmax < off) m = 2;                                         // range missing
        else if (x + dmin < off) m = 4;                                    // range incomplete
    } else if (dmin > 0) {
        if (x + dmin > W - 1 - off) m = 2;
        else if (x + dmax > W - 1 - off) m = 4;
    } else {
        if (x + dmin < off || x + dmax > W - 1 - off) m = 4;
    }
    mask[i] = m;
}

__global__ void __launch_bounds__(256) reverse_cv_kernel(const float *__restrict__ left, int H, int W, int D, int min_disp,
                                                         float *__restrict__ right) {
    const long n = (long)H * W * D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % D);
        const long pix = i / D;
        const int j = (int)(pix % W);
        const long row = pix / W;
        const int c = j + k + min_disp;
        right[i] = (c >= 0 && c < W) ? left[(row * W + c) * D + (D - 1 - k)] : nan_f();
    }
}

// Tiled version: right(row, j, d) = left(row, j + d + min_disp, D - 1 - d) is a bijection between cells, and the cells an
// input pixel x gives to a tile of 32 output pixels [j0, j0 + 32) are ONE run of 32 consecutive disparities
// e = e_lo(x) + lane, e_lo(x) = D - 1 - (x - min_disp - j0), landing on a diagonal of the tile (pixel `lane`, disparity
// D - 1 - e).  A CTA owns (row, tile): its warps walk the D + 31 input pixels of the band, every load is one contiguous
// 128-byte run, the diagonal goes into a shared-memory tile whose pitch is even (lane stride pitch - 1 is odd: no bank
// conflict), and the finished tile -- 32 pixels x D floats, contiguous in the volume -- is written with 16-byte stores.
// Every cell is read once and written once.
__global__ void __launch_bounds__(256) reverse_cv_tiled_kernel(const float *__restrict__ left, int H, int W, int D, int min_disp,
                                                               float *__restrict__ right, int tiles_x) {
    extern __shared__ __align__(16) float rtile[];                      // [32][pitch]
    const int pitch = (D + 1) & ~1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long t = blockIdx.x; t < (long)H * tiles_x; t += gridDim.x) {
        const long row = t / tiles_x;
        const int j0 = (int)(t % tiles_x) * 32;
        const int npix = min(32, W - j0);
        const float *lrow = left + (size_t)row * W * D;
        const int xb = j0 + min_disp;                                   // input pixel of (pixel 0, disparity 0)
        // eight independent runs per warp and step: all loads are issued before the first shared-memory store
        for (int q0 = warp * 8; q0 < D + 31; q0 += 64) {
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int q = q0 + u, x = xb + q;
                const int e = D - 1 - q + lane;                         // e_lo = D - 1 - (x - min_disp - j0) = D - 1 - q
                v[u] = (e >= 0 && e < D && x >= 0 && x < W) ? __ldg(lrow + (size_t)x * D + e) : nan_f();
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int d = q0 + u - lane;                            // D - 1 - e
                if (d >= 0 && d < D) rtile[lane * pitch + d] = v[u];
            }
        }
        __syncthreads();
        float *orow = right + ((size_t)row * W + j0) * D;
        if ((D & 3) == 0 && (reinterpret_cast<uintptr_t>(right) & 15) == 0) {   // pitch == D: the tile is contiguous
            const int n4 = npix * D / 4;
            for (int i = threadIdx.x; i < n4; i += 256) reinterpret_cast<float4 *>(orow)[i] = reinterpret_cast<const float4 *>(rtile)[i];
        } else {
            for (int j = warp; j < npix; j += 8)
                for (int d = lane; d < D; d += 32) orow[(size_t)j * D + d] = rtile[j * pitch + d];
        }
        __syncthreads();
    }
}

// MedianFilter.filter_disparity (filter/median.py:96-132), the two element-wise halves around the 3x3 NaN-median:
// (a) masked = invalid pixel ? NaN : disparity; (b) disparity = isfinite(masked) ? median(masked) : disparity.
__global__ void __launch_bounds__(256) filter_mask_kernel(const float *__restrict__ disp, const uint16_t *__restrict__ mask, long n,
                                                          float *__restrict__ masked) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) masked[i] = (mask[i] & 0x3C3) ? nan_f() : disp[i];
}
__global__ void __launch_bounds__(256) filter_select_kernel(float *__restrict__ disp, const float *__restrict__ masked,
                                                            const float *__restrict__ med, long n) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float m = masked[i];
    if (m == m && fabsf(m) != CUDART_INF_F) disp[i] = med[i];      // np.isfinite(masked_data)
}

// reverse_disp_range (matching_cost/cpp/src/matching_cost.cpp:59-131): the reference scatters, for every left pixel
// (row, col) and every d in [int(min), int(max)], the value -d into running min / max of the right pixel col + d.  As a
// gather: left pixel c covers the right interval [c + dmin(c), c + dmax(c)], so right_min(rc) = (smallest covering c) - rc
// and right_max(rc) = (largest covering c) - rc.  One CTA per row: the row's own bounds of d limit the columns a right
// pixel has to look at, and both scans stop at their first hit (constant ranges: the first candidate).
__global__ void __launch_bounds__(256) reverse_disp_range_kernel(const float *__restrict__ lmin, const float *__restrict__ lmax, int H, int W,
                                                                 float *__restrict__ rmin, float *__restrict__ rmax) {
    __shared__ int s_lo[8], s_hi[8];
    const int row = blockIdx.x, tid = threadIdx.x;
    const float *a = lmin + (size_t)row * W, *b = lmax + (size_t)row * W;
    int lo = INT_MAX, hi = INT_MIN;
    for (int c = tid; c < W; c += blockDim.x) {
        const float u = __ldg(a + c), v = __ldg(b + c);
        if (u == u && v == v) {
            lo = min(lo, (int)u);
            hi = max(hi, (int)v);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if ((tid & 31) == 0) { s_lo[tid >> 5] = lo; s_hi[tid >> 5] = hi; }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { lo = min(lo, s_lo[w]); hi = max(hi, s_hi[w]); }
    for (int rc = tid; rc < W; rc += blockDim.x) {
        float mn = nan_f(), mx = nan_f();
        if (lo <= hi) {
            const int c0 = max(0L, (long)rc - hi), c1 = (int)min((long)W - 1, (long)rc - lo);
            for (int c = c0; c <= c1; ++c) {
                const float u = __ldg(a + c), v = __ldg(b + c);
                if (u == u && v == v && (int)u <= rc - c && rc - c <= (int)v) { mn = (float)(c - rc); break; }
            }
            if (mn == mn)
                for (int c = c1; c >= c0; --c) {
                    const float u = __ldg(a + c), v = __ldg(b + c);
                    if (u == u && v == v && (int)u <= rc - c && rc - c <= (int)v) { mx = (float)(c - rc); break; }
                }
        }
        rmin[(size_t)row * W + rc] = mn;
        rmax[(size_t)row * W + rc] = mx;
    }
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_filter_median3(float *d_disp, const uint16_t *d_mask, int H, int W, float *d_scratch, void *stream) {
    if (!d_disp || !d_mask || !d_scratch || H <= 0 || W <= 0) {
        set_error("pb200_filter_median3: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const long n = (long)H * W;
    float *masked = d_scratch, *med = d_scratch + n;
    cudaStream_t s = (cudaStream_t)stream;
    filter_mask_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_disp, d_mask, n, masked);
    PB200_LAUNCH_CHECK("filter_mask_kernel");
    const int rc = pb200_median3(masked, H, W, med, stream);
    if (rc != PB200_OK) return rc;
    filter_select_kernel<<<ceil_div(n, 256), 256, 0, s>>>(d_disp, masked, med, n);
    PB200_LAUNCH_CHECK("filter_select_kernel");
    return PB200_OK;
}

extern "C" int pb200_wta(const float *d_cv, int H, int W, int D, int dmin, int is_max, float invalid_disparity,
                         float *d_disp, uint8_t *d_all_nan, void *stream) {
    if (!d_cv || !d_disp || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_wta: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const long n_pix = (long)H * W;
    long blocks = (n_pix + 8 * WTA_PPW - 1) / (8 * WTA_PPW);     // 8 warps per block, WTA_PPW pixels per warp per trip
    const long cap = (long)sm_count() * 8 * 8;
    if (blocks > cap) blocks = cap;
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_cv) & 15) == 0);
#define PB200_W(MX, V) wta_kernel<MX, V><<<(int)blocks, 256, 0, s>>>(d_cv, n_pix, D, dmin, invalid_disparity, d_disp, d_all_nan)
    if (is_max) { if (vec) PB200_W(true, true); else PB200_W(true, false); }
    else        { if (vec) PB200_W(false, true); else PB200_W(false, false); }
#undef PB200_W
    PB200_LAUNCH_CHECK("wta_kernel");
    return PB200_OK;
}

extern "C" int pb200_validity_mask(uint16_t *d_mask, const uint8_t *d_all_nan, int H, int W, int offset, int wta_invalidate,
                                   void *stream) {
    if (!d_mask || H <= 0 || W <= 0 || offset < 0) {
        set_error("pb200_validity_mask: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    validity_mask_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_mask, d_all_nan, H, W, offset, wta_invalidate);
    PB200_LAUNCH_CHECK("validity_mask_kernel");
    return PB200_OK;
}

extern "C" int pb200_validity_mask_init(uint16_t *d_mask, int H, int W, int dmin, int dmax, int offset, void *stream) {
    if (!d_mask || H <= 0 || W <= 0 || offset < 0 || dmax < dmin) {
        set_error("pb200_validity_mask_init: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    validity_init_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_mask, H, W, dmin, dmax, offset);
    PB200_LAUNCH_CHECK("validity_init_kernel");
    return PB200_OK;
}

extern "C" int pb200_reverse_cost_volume(const float *d_left_cv, int H, int W, int D, int min_disp, float *d_right_cv,
                                         void *stream) {
    if (!d_left_cv || !d_right_cv || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_reverse_cost_volume: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const size_t tile_bytes = (size_t)32 * ((D + 1) & ~1) * sizeof(float);
    if (D >= 8 && tile_bytes <= 96 * 1024 && option(OPT_REVERSE_GATHER) <= 0) {
        const int tiles_x = ceil_div(W, 32);
        long grid = (long)H * tiles_x;
        const long cap = (long)sm_count() * 32;
        if (grid > cap) grid = cap;
        PB200_CUDA(cudaFuncSetAttribute(reverse_cv_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_bytes));
        reverse_cv_tiled_kernel<<<(int)grid, 256, tile_bytes, (cudaStream_t)stream>>>(d_left_cv, H, W, D, min_disp, d_right_cv, tiles_x);
        PB200_LAUNCH_CHECK("reverse_cv_tiled_kernel");
        note_path(STAGE_REVERSE, PATH_REVERSE_TILED);
        return PB200_OK;
    }
    long blocks = ((long)H * W * D + 255) / 256;
    const long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    reverse_cv_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_left_cv, H, W, D, min_disp, d_right_cv);
    PB200_LAUNCH_CHECK("reverse_cv_kernel");
    note_path(STAGE_REVERSE, PATH_REVERSE_GATHER);
    return PB200_OK;
}

extern "C" int pb200_reverse_disp_range(const float *d_left_min, const float *d_left_max, int H, int W, float *d_right_min,
                                        float *d_right_max, void *stream) {
    if (!d_left_min || !d_left_max || !d_right_min || !d_right_max || H <= 0 || W <= 0) {
        set_error("pb200_reverse_disp_range: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    reverse_disp_range_kernel<<<H, 256, 0, (cudaStream_t)stream>>>(d_left_min, d_left_max, H, W, d_right_min, d_right_max);
    PB200_LAUNCH_CHECK("reverse_disp_range_kernel");
    return PB200_OK;
}
