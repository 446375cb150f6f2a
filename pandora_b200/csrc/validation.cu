// validation.cu -- fast cross-checking: right disparity map straight from the LEFT cost volume, and the
// left/right consistency check.
//
// (1) wta_right_kernel replaces the pair reverse_cost_volume (matching_cost/cpp/src/matching_cost.cpp:26-57) +
//     WinnerTakesAll.to_disp on the right volume that the reference runs for "cross_checking_fast"
//     (state_machine.py:436-448): right(i, j, k) = left(i, j + k + min_disp_right, D-1-k).  The reference writes the
//     17 GB right volume and reads it back (8*D bytes per pixel, plus the 4*D read of the left one); here the left
//     volume is read ONCE (4*D bytes per pixel) and the right volume never exists.  One CTA walks one image row:
//     tiles of TX left columns x D cells are staged in shared memory with cp.async (coalesced 16-byte copies, double
//     buffered), and every thread owns one right pixel j for as long as the sliding window of left columns
//     [j + min_disp_right, j + min_disp_right + D) overlaps the staged tiles, keeping (best cost, best index) in
//     registers.  Within a tile a warp's 32 pixels read 32 consecutive cells of the same left column: conflict free.
//     Ties: lowest right index k wins, i.e. the candidate met first when the left column x increases.
// (2) cross_checking_kernel replaces CrossCheckingAccurate.disparity_checking (validation/validation.py:226-371):
//     one thread per left pixel, O(D) only for the pixels the check invalidates (mismatch / occlusion search).
#include "common.cuh"

namespace pb200 {

namespace {

constexpr uint16_t MSK_INVALID = 0x3C3;     // constants.py:28
constexpr uint16_t MSK_OCCLUSION = 1 << 8;  // constants.py:46
constexpr uint16_t MSK_MISMATCH = 1 << 9;   // constants.py:48
constexpr int TX = 32;                      // left columns per staged tile (= one warp of right pixels retired per tile)

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void *gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// blockDim.x = R = roundup32(TX + D - 1) threads; dynamic shared memory = 2 * TX * D floats.
template <bool IS_MAX>
__global__ void wta_right_kernel(const float *__restrict__ left_cv, int H, int W, int D, int min_disp_right, float invalid_disparity,
                                 float *__restrict__ disp, uint8_t *__restrict__ all_nan) {
    extern __shared__ __align__(16) float tile_smem[];
    const int R = blockDim.x, t = threadIdx.x;
    const long row = blockIdx.x;
    const float *src_row = left_cv + row * (long)W * D;
    const int ntiles_x = (W + TX - 1) / TX;
    // right pixel j meets left columns x = j + min_disp_right + k, k in [0, D).  jbase(n) = first right pixel still
    // alive when tile n = [n*TX, n*TX + TX) is processed: its last column x = j + mdr + D - 1 >= n*TX.
    const int jbase0 = -min_disp_right - (D - 1);
    // thread t owns the pixels j = jbase0 + t + m*R; its current one is the unique such j in [jbase(n), jbase(n) + R)
    int j = jbase0 + t;
    const float init = IS_MAX ? -CUDART_INF_F : CUDART_INF_F;
    float bv = init;
    int bk = 0x7fffffff;
    bool any = false;

    auto stage = [&](int n) {                              // enqueue tile n into buffer n & 1
        const int x0 = n * TX;
        const int ncols = min(TX, W - x0);
        const int nvec = ncols * D / 4;                    // D % 4 == 0
        const float *g = src_row + (long)x0 * D;
        const uint32_t s = smem_u32(tile_smem + (size_t)(n & 1) * TX * D);
        for (int v = t; v < nvec; v += R) cp_async16(s + (uint32_t)v * 16u, g + (long)v * 4);
    };
    // tiles past the image (n >= ntiles_x) carry no data: they only retire the remaining pixels
    const int jlast = W - 1;
    const int ntiles = max(ntiles_x, (jlast - jbase0) / TX + 1);
    // right pixels whose whole window lies left of the image (j < jbase0) never meet a tile: all-NaN
    for (int jj = t; jj < min(jbase0, W); jj += R) {
        disp[row * W + jj] = invalid_disparity;
        if (all_nan) all_nan[row * W + jj] = 1;
    }
    stage(0);
    cp_commit();
    for (int n = 0; n < ntiles; ++n) {
        if (n + 1 < ntiles_x) stage(n + 1);
        cp_commit();
        cp_wait<1>();
        __syncthreads();                                   // tile n visible to every thread
        if (n < ntiles_x && j >= 0 && j < W) {
            const int x0 = n * TX;
            const int ncols = min(TX, W - x0);
            const float *tile = tile_smem + (size_t)(n & 1) * TX * D;
            // k = x - j - mdr must lie in [0, D): x in [j + mdr, j + mdr + D)
            const int xa = max(x0, j + min_disp_right), xb = min(x0 + ncols, j + min_disp_right + D);
            for (int x = xa; x < xb; ++x) {
                const int k = x - j - min_disp_right;
                const float v = tile[(x - x0) * D + (D - 1 - k)];
                if (v == v) {
                    any = true;
                    if (IS_MAX ? (v > bv) : (v < bv)) { bv = v; bk = k; }
                }
            }
        }
        // retire the pixels whose window ends inside this tile: j + mdr + D - 1 < (n + 1) * TX
        if (j + min_disp_right + D - 1 < (n + 1) * TX) {
            if (j >= 0 && j < W) {
                if (bv == init) bk = 0;                    // nothing but NaN / +-inf: np.argmin / np.argmax give index 0
                disp[row * W + j] = any ? (float)(min_disp_right + bk) : invalid_disparity;
                if (all_nan) all_nan[row * W + j] = any ? 0 : 1;
            }
            j += R;
            bv = init;
            bk = 0x7fffffff;
            any = false;
        }
        __syncthreads();                                   // everyone is done with buffer n & 1 before tile n + 2 lands in it
    }
}

__global__ void __launch_bounds__(256) cross_checking_kernel(const float *__restrict__ disp_left, uint16_t *__restrict__ mask,
                                                             const float *__restrict__ disp_right, int H, int W, float threshold,
                                                             int dmin, int dmax, int offset, float *__restrict__ conf) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int c = (int)(i % W);
    const long row = i / W;
    uint16_t m = mask[i];
    float dist = nan_f();
    const float dl = disp_left[i];
    if ((m & MSK_INVALID) == 0 && dl == dl) {              // valid pixels only (validation.py:301-308)
        const double cr = rint((double)c + (double)dl);     // np.rint: half to even
        if (cr >= 0.0 && cr < (double)W) {
            const float *rrow = disp_right + row * W;
            float rd = rrow[(int)cr];
            if (rd != rd) rd = CUDART_INF_F;                 // NaN -> inf (validation.py:320-323)
            dist = fabsf(rd + dl);
            if (dist > threshold) {
                // mismatch if Disp_right(i + d) = -d for any d of the range, occlusion otherwise (validation.py:331-353)
                bool found = false;
                for (int d = dmin; d <= dmax && !found; ++d) {
                    const int idx = c + d;
                    if (idx >= 0 && idx < W) found = (rintf(rrow[idx]) == (float)(-d));
                }
                m = (uint16_t)(m + (found ? MSK_MISMATCH : MSK_OCCLUSION));
            }
        }
    }
    if (offset > 0) {                                        // mask_border (criteria.py:325-353)
        const int y = (int)row;
        if (y < offset || y >= H - offset || c < offset || c >= W - offset) m = 1;
    }
    mask[i] = m;
    if (conf) conf[i] = dist;
}

}  // namespace

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_wta_right(const float *d_left_cv, int H, int W, int D, int min_disp_right, int is_max, float invalid_disparity,
                               float *d_disp, uint8_t *d_all_nan, void *stream) {
    if (!d_left_cv || !d_disp || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_wta_right: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const int R = (TX + D - 1 + 31) / 32 * 32;
    const size_t smem = (size_t)2 * TX * D * sizeof(float);
    if (D % 4 != 0 || R > 1024 || smem > 200 * 1024 || (reinterpret_cast<uintptr_t>(d_left_cv) & 15)) {
        set_error("pb200_wta_right: needs D %% 4 == 0, D <= 992 and a 16-byte aligned volume (use pb200_reverse_cost_volume + pb200_wta)");
        return PB200_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    if (is_max) {
        PB200_CUDA(cudaFuncSetAttribute((const void *)wta_right_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wta_right_kernel<true><<<H, R, smem, s>>>(d_left_cv, H, W, D, min_disp_right, invalid_disparity, d_disp, d_all_nan);
    } else {
        PB200_CUDA(cudaFuncSetAttribute((const void *)wta_right_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        wta_right_kernel<false><<<H, R, smem, s>>>(d_left_cv, H, W, D, min_disp_right, invalid_disparity, d_disp, d_all_nan);
    }
    PB200_LAUNCH_CHECK("wta_right_kernel");
    return PB200_OK;
}

extern "C" int pb200_cross_checking(const float *d_disp_left, uint16_t *d_mask_left, const float *d_disp_right, int H, int W,
                                    float threshold, int dmin, int dmax, int offset, float *d_conf, void *stream) {
    if (!d_disp_left || !d_mask_left || !d_disp_right || H <= 0 || W <= 0 || dmax < dmin || offset < 0) {
        set_error("pb200_cross_checking: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    cross_checking_kernel<<<ceil_div((long)H * W, 256), 256, 0, (cudaStream_t)stream>>>(d_disp_left, d_mask_left, d_disp_right, H, W,
                                                                                       threshold, dmin, dmax, offset, d_conf);
    PB200_LAUNCH_CHECK("cross_checking_kernel");
    return PB200_OK;
}
