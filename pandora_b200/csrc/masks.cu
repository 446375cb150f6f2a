// masks.cu -- input masks (msk / no_data) and per-pixel disparity grids on the device.
//
// Replaces, for subpix == 1 and step == 1:
//   binary_dilation_msk                       src/pandora/criteria.py:36-63
//   allocate_left_mask / allocate_right_mask  criteria.py:178-288
//   mask_partially_missing_variable_ranges    criteria.py:161-175 + cpp/src/criteria.cpp:27-110
//   masks_dilatation + the masking loops of cv_masked   matching_cost/matching_cost.py:484-602, 815-856
// The reference walks the volume disparity by disparity in Python (xarray .loc per disparity, one np.where per
// disparity for the grids).  Here: one O(H*W*w^2) pass turns each msk into a byte of flags per pixel, one O(H*W*D)
// byte-gather pass adds the criteria bits, and ONE pass over the volume (4*D bytes read per pixel, written only
// where a cell changes) applies both masks and the per-pixel [disp_min, disp_max] and reports all-NaN pixels.
#include "common.cuh"

namespace pb200 {

namespace {

constexpr uint8_t FLAG_NODATA_DILATED = 1;   // a no_data pixel inside the window (masks_dilatation: NaN)
constexpr uint8_t FLAG_INVALID = 2;          // msk is neither valid_pixels nor no_data (NaN as well)
constexpr uint8_t FLAG_NOT_VALID = 4;        // msk != valid_pixels (criteria.py:171)

__global__ void __launch_bounds__(256) mask_flags_kernel(const int16_t *__restrict__ msk, int H, int W, int valid_pixels, int no_data,
                                                         int window, uint8_t *__restrict__ flags) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int x = (int)(i % W), y = (int)(i / W);
    const int half = (window - 1) / 2;
    const int v = msk[i];
    uint8_t f = 0;
    if (v != valid_pixels && v != no_data) f |= FLAG_INVALID;
    if (v != valid_pixels) f |= FLAG_NOT_VALID;
    bool nd = false;
    for (int yy = max(0, y - half); yy <= min(H - 1, y + half) && !nd; ++yy)
        for (int xx = max(0, x - half); xx <= min(W - 1, x + half); ++xx)
            if (msk[(long)yy * W + xx] == no_data) { nd = true; break; }
    if (nd) f |= FLAG_NODATA_DILATED;
    flags[i] = f;
}

__global__ void __launch_bounds__(256) validity_masks_kernel(uint16_t *__restrict__ mask, int H, int W, int dmin, int dmax, int off,
                                                             const uint8_t *__restrict__ fl, const uint8_t *__restrict__ fr,
                                                             const float *__restrict__ gmin, const float *__restrict__ gmax) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int c = (int)(i % W);
    const long row = i / W;
    uint16_t m = mask[i];
    if (fl != nullptr) {                                  // allocate_left_mask
        const uint8_t f = fl[i];
        if (f & FLAG_NODATA_DILATED) m = (uint16_t)(m + 1);
        if (f & FLAG_INVALID) m = (uint16_t)(m + 64);
    }
    if (fr != nullptr) {                                  // allocate_right_mask
        const uint8_t *r = fr + row * W;
        const bool bit_1 = (dmax < 0) ? (c + dmax < off) : ((dmin > 0) ? (c + dmin > W - 1 - off) : false);
        if (!bit_1) {
            const int nd = dmax - dmin + 1;
            int b_2_7 = 0, no_data_right = 0;
            for (int d = dmin; d <= dmax; ++d) {
                const int cd = c + d;
                if (cd >= off && cd <= W - 1 - off) {
                    const uint8_t f = r[cd];
                    b_2_7 += (f & FLAG_INVALID) ? 1 : 0;
                    no_data_right += (f & FLAG_NODATA_DILATED) ? 1 : 0;
                } else {
                    ++b_2_7;
                    ++no_data_right;
                }
            }
            if (b_2_7 == nd) m = (uint16_t)(m + 128);
            if (no_data_right == nd) m = (uint16_t)(m + 2);
        }
        if (gmin != nullptr && gmax != nullptr) {         // partially_missing_variable_ranges (needs gmin <= gmax)
            // a NaN cell of either grid = no range for this pixel (criteria.cpp:66-72 never sees one: the Python side
            // passes nanmin / nanmax bounds): flagged like a range that leaves the image, and never converted to int
            const float gl = gmin[i], gh = gmax[i];
            const bool finite = (gl == gl) && (gh == gh) && fabsf(gl) < 1e9f && fabsf(gh) < 1e9f;
            const int lo = finite ? (int)gl + c : -1, hi = finite ? (int)gh + c : -1;
            bool inside = finite && lo >= 0 && hi < W && lo <= hi;
            for (int x = lo; inside && x <= hi; ++x) inside = (r[x] & FLAG_NOT_VALID) == 0;
            if (!inside) m |= 4096;
        }
    }
    mask[i] = m;
}

// Row version of the kernel above (one CTA per image row): the per-pixel loops over the disparity range -- D byte gathers
// per pixel, 4.3 G of them at 4096 x 4096 x 256 -- become differences of three per-row prefix counts (INVALID,
// NODATA_DILATED, NOT_VALID of the right image's flags) built once in shared memory: O(H * W) instead of O(H * W * D).
__global__ void __launch_bounds__(256) validity_masks_rows_kernel(uint16_t *__restrict__ mask, int H, int W, int dmin, int dmax, int off,
                                                                  const uint8_t *__restrict__ fl, const uint8_t *__restrict__ fr,
                                                                  const float *__restrict__ gmin, const float *__restrict__ gmax) {
    extern __shared__ uint32_t vm_pre[];                 // [3][W + 1] exclusive prefix counts: INVALID | NODATA_DILATED | NOT_VALID
    __shared__ uint32_t warp_tot[3][8];
    const long row = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint32_t *pI = vm_pre, *pN = vm_pre + (W + 1), *pV = vm_pre + 2 * (W + 1);
    if (fr != nullptr) {
        const uint8_t *r = fr + row * W;
        const int chunk = (W + 255) / 256, c0 = min(W, tid * chunk), c1 = min(W, c0 + chunk);
        uint32_t a = 0, b = 0, v = 0;
        for (int c = c0; c < c1; ++c) {
            const uint8_t f = r[c];
            a += (f & FLAG_INVALID) ? 1u : 0u;
            b += (f & FLAG_NODATA_DILATED) ? 1u : 0u;
            v += (f & FLAG_NOT_VALID) ? 1u : 0u;
        }
        uint32_t sa = a, sb = b, sv = v;                 // inclusive scan over the 256 chunk sums
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t ta = __shfl_up_sync(0xffffffffu, sa, o), tb = __shfl_up_sync(0xffffffffu, sb, o), tv = __shfl_up_sync(0xffffffffu, sv, o);
            if (lane >= o) { sa += ta; sb += tb; sv += tv; }
        }
        if (lane == 31) { warp_tot[0][wid] = sa; warp_tot[1][wid] = sb; warp_tot[2][wid] = sv; }
        __syncthreads();
        uint32_t ba = 0, bb = 0, bv = 0;
        for (int w = 0; w < wid; ++w) { ba += warp_tot[0][w]; bb += warp_tot[1][w]; bv += warp_tot[2][w]; }
        uint32_t ea = ba + sa - a, eb = bb + sb - b, ev = bv + sv - v;      // exclusive prefix at this thread's first column
        for (int c = c0; c < c1; ++c) {
            pI[c] = ea; pN[c] = eb; pV[c] = ev;
            const uint8_t f = r[c];
            ea += (f & FLAG_INVALID) ? 1u : 0u;
            eb += (f & FLAG_NODATA_DILATED) ? 1u : 0u;
            ev += (f & FLAG_NOT_VALID) ? 1u : 0u;
        }
        if (c1 == W && c0 < W) { pI[W] = ea; pN[W] = eb; pV[W] = ev; }
        __syncthreads();
    }
    const int nd = dmax - dmin + 1;
    for (int c = tid; c < W; c += blockDim.x) {
        const long i = row * W + c;
        uint16_t m = mask[i];
        if (fl != nullptr) {                              // allocate_left_mask
            const uint8_t f = fl[i];
            if (f & FLAG_NODATA_DILATED) m = (uint16_t)(m + 1);
            if (f & FLAG_INVALID) m = (uint16_t)(m + 64);
        }
        if (fr != nullptr) {                              // allocate_right_mask
            const bool bit_1 = (dmax < 0) ? (c + dmax < off) : ((dmin > 0) ? (c + dmin > W - 1 - off) : false);
            if (!bit_1) {
                // columns c + d inside [off, W - 1 - off] count their flags, the others count as flagged (criteria.py:216-288)
                const int lo = max(c + dmin, off), hi = min(c + dmax, W - 1 - off);
                const int inside_n = max(0, hi - lo + 1);
                int b_2_7 = nd - inside_n, no_data_right = nd - inside_n;
                if (inside_n > 0) {
                    b_2_7 += (int)(pI[hi + 1] - pI[lo]);
                    no_data_right += (int)(pN[hi + 1] - pN[lo]);
                }
                if (b_2_7 == nd) m = (uint16_t)(m + 128);
                if (no_data_right == nd) m = (uint16_t)(m + 2);
            }
            if (gmin != nullptr && gmax != nullptr) {     // partially_missing_variable_ranges
                const float gl = gmin[i], gh = gmax[i];
                const bool finite = (gl == gl) && (gh == gh) && fabsf(gl) < 1e9f && fabsf(gh) < 1e9f;
                const int lo = finite ? (int)gl + c : -1, hi = finite ? (int)gh + c : -1;
                bool inside = finite && lo >= 0 && hi < W && lo <= hi;
                // the per-pixel loop stops at the first pixel that is not valid: "all valid" == no NOT_VALID flag in [lo, hi]
                if (inside) inside = (pV[hi + 1] - pV[lo]) == 0u;
                if (!inside) m |= 4096;
            }
        }
        mask[i] = m;
    }
}

// one warp per pixel; every lane owns float4 groups of the disparity vector
template <bool VEC4>
__global__ void __launch_bounds__(256) cv_masked_kernel(float *__restrict__ cv, long n_pix, int W, int D, int dmin,
                                                        const uint8_t *__restrict__ fl, const uint8_t *__restrict__ fr,
                                                        const float *__restrict__ gmin, const float *__restrict__ gmax,
                                                        uint8_t *__restrict__ all_nan) {
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    constexpr uint8_t BAD = FLAG_NODATA_DILATED | FLAG_INVALID;
    for (long pix = warp0; pix < n_pix; pix += nwarps) {
        const int c = (int)(pix % W);
        const long row = pix / W;
        const bool left_bad = fl != nullptr && (fl[pix] & BAD) != 0;
        const uint8_t *r = fr ? fr + row * W : nullptr;
        const float lo = gmin ? gmin[pix] : 0.f, hi = gmax ? gmax[pix] : 0.f;
        float *p = cv + pix * D;
        bool any = false;
        auto cell = [&](float v, int k, bool &changed) -> float {
            const int d = dmin + k;
            const int cd = c + d;
            bool bad = false;
            if (cd >= 0 && cd < W) bad = left_bad || (r != nullptr && (r[cd] & BAD) != 0);   // mask_column_interval: inside only
            if (gmin != nullptr) bad = bad || ((float)d < lo) || ((float)d > hi);
            if (bad && v == v) { changed = true; v = nan_f(); }
            any = any || (v == v);
            return v;
        };
        if (VEC4) {
            for (int k = lane * 4; k < D; k += 128) {
                float4 v = *reinterpret_cast<const float4 *>(p + k);
                bool changed = false;
                v.x = cell(v.x, k, changed); v.y = cell(v.y, k + 1, changed); v.z = cell(v.z, k + 2, changed); v.w = cell(v.w, k + 3, changed);
                if (changed) *reinterpret_cast<float4 *>(p + k) = v;
            }
        } else {
            for (int k = lane; k < D; k += 32) {
                bool changed = false;
                const float v = cell(p[k], k, changed);
                if (changed) p[k] = v;
            }
        }
        any = __any_sync(0xffffffffu, any);
        if (lane == 0 && all_nan) all_nan[pix] = any ? 0 : 1;
    }
}

}  // namespace

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_mask_flags(const int16_t *d_msk, int H, int W, int valid_pixels, int no_data, int window, uint8_t *d_flags,
                                void *stream) {
    if (!d_msk || !d_flags || H <= 0 || W <= 0 || window < 1 || (window & 1) == 0) {
        set_error("pb200_mask_flags: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    mask_flags_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_msk, H, W, valid_pixels, no_data, window, d_flags);
    PB200_LAUNCH_CHECK("mask_flags_kernel");
    return PB200_OK;
}

extern "C" int pb200_validity_mask_masks(uint16_t *d_mask, int H, int W, int dmin, int dmax, int offset, const uint8_t *d_flags_left,
                                         const uint8_t *d_flags_right, const float *d_grid_min, const float *d_grid_max, void *stream) {
    if (!d_mask || H <= 0 || W <= 0 || dmax < dmin || offset < 0) {
        set_error("pb200_validity_mask_masks: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const size_t smem = 3 * (size_t)(W + 1) * sizeof(uint32_t);
    if (smem <= 200 * 1024) {                            // row kernel: prefix counts of the right flags in shared memory
        PB200_CUDA(cudaFuncSetAttribute(validity_masks_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        validity_masks_rows_kernel<<<H, 256, smem, (cudaStream_t)stream>>>(d_mask, H, W, dmin, dmax, offset, d_flags_left, d_flags_right,
                                                                           d_grid_min, d_grid_max);
        PB200_LAUNCH_CHECK("validity_masks_rows_kernel");
        return PB200_OK;
    }
    validity_masks_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_mask, H, W, dmin, dmax, offset, d_flags_left,
                                                                                       d_flags_right, d_grid_min, d_grid_max);
    PB200_LAUNCH_CHECK("validity_masks_kernel");
    return PB200_OK;
}

extern "C" int pb200_cv_masked(float *d_cv, int H, int W, int D, int dmin, const uint8_t *d_flags_left, const uint8_t *d_flags_right,
                               const float *d_grid_min, const float *d_grid_max, uint8_t *d_all_nan, void *stream) {
    if (!d_cv || H <= 0 || W <= 0 || D <= 0 || ((d_grid_min == nullptr) != (d_grid_max == nullptr))) {
        set_error("pb200_cv_masked: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const long n_pix = (long)H * W;
    long blocks = (n_pix + 7) / 8;
    const long cap = (long)sm_count() * 8 * 8;
    if (blocks > cap) blocks = cap;
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_cv) & 15) == 0);
    cudaStream_t s = (cudaStream_t)stream;
    if (vec) cv_masked_kernel<true><<<(int)blocks, 256, 0, s>>>(d_cv, n_pix, W, D, dmin, d_flags_left, d_flags_right, d_grid_min, d_grid_max, d_all_nan);
    else cv_masked_kernel<false><<<(int)blocks, 256, 0, s>>>(d_cv, n_pix, W, D, dmin, d_flags_left, d_flags_right, d_grid_min, d_grid_max, d_all_nan);
    PB200_LAUNCH_CHECK("cv_masked_kernel");
    return PB200_OK;
}
