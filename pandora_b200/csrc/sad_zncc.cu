// sad_zncc.cu -- SAD / SSD and ZNCC matching-cost volumes.
//
// SAD/SSD: src/pandora/matching_cost/sad_ssd.py:75-207 (+ point_interval matching_cost.py:429-482).
//   cv[y,x,k] = sum over the w x w window of |L - R(.., +d)| (or squared), float32, accumulated in the
//   order numpy uses for the reference's strided reduction (sad_ssd.py:340-368): column offset
//   outer, row offset inner, starting from the first term -- so float images are bit-exact too.
//   Finite iff half <= y < H-half, half <= x < W-half, 0 <= x-half+d and x+half+d < W.
// ZNCC: matching_cost/zncc.py:114-241, 244-277 with the float64 window statistics of
//   img_tools.py:834-879 (mean) and 915-952 (std, variance clamp 1e-15*|E[x^2]|): products and
//   squares are formed in float32 like numpy does on float32 images, accumulated in float64.
//
// Layout: lanes run over the disparity axis (fastest in memory) so every store is a coalesced
// 128-byte line and the right-image reads of a warp are 32 consecutive pixels.
#include "common.cuh"

namespace pb200 {

template <bool SQUARED>
__global__ void __launch_bounds__(256) sad_ssd_kernel(const float *__restrict__ L, const float *__restrict__ R, int H, int W,
                                                      int win, int dmin, int D, float *__restrict__ cv) {
    const int half = win / 2;
    const long pix = (long)blockIdx.x * blockDim.y + threadIdx.y;
    if (pix >= (long)H * W) return;
    const int y = (int)(pix / W), x = (int)(pix % W);
    float *dst = cv + pix * D;
    const bool centre_ok = (y >= half && y < H - half && x >= half && x < W - half);
    for (int k = threadIdx.x; k < D; k += 32) {
        const int d = dmin + k;
        float acc = nan_f();
        if (centre_ok && x - half + d >= 0 && x + half + d < W) {
            bool first = true;
            for (int dx = -half; dx <= half; ++dx)
                for (int dy = -half; dy <= half; ++dy) {
                    const float a = __ldg(L + (size_t)(y + dy) * W + x + dx);
                    const float b = __ldg(R + (size_t)(y + dy) * W + x + dx + d);
                    const float df = a - b;
                    const float t = SQUARED ? df * df : fabsf(df);
                    acc = first ? t : acc + t;
                    first = false;
                }
        }
        dst[k] = acc;
    }
}

// per-pixel window mean and std in float64 (valid centres only; others 0)
__global__ void __launch_bounds__(256) zncc_stats_kernel(const float *__restrict__ img, int H, int W, int win,
                                                         double *__restrict__ mean, double *__restrict__ stdv) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int half = win / 2;
    const int y = (int)(i / W), x = (int)(i % W);
    double m = 0.0, sd = 0.0;
    if (y >= half && y < H - half && x >= half && x < W - half) {
        double s1 = 0.0, s2 = 0.0;
        for (int dy = -half; dy <= half; ++dy)
            for (int dx = -half; dx <= half; ++dx) {
                const float v = __ldg(img + (size_t)(y + dy) * W + x + dx);
                s1 += (double)v;
                s2 += (double)(v * v);          // float32 square like `selected_band**2` (img_tools.py:936)
            }
        const double n = (double)(win * win);
        m = s1 / n;
        const double m2 = s2 / n;
        double var = m2 - m * m;
        if (var < 1e-15 * fabs(m2)) var = 0.0;  // img_tools.py:951
        sd = sqrt(var);
    }
    mean[i] = m;
    stdv[i] = sd;
}

__global__ void __launch_bounds__(256) zncc_kernel(const float *__restrict__ L, const float *__restrict__ R,
                                                   const double *__restrict__ meanL, const double *__restrict__ stdL,
                                                   const double *__restrict__ meanR, const double *__restrict__ stdR, int H,
                                                   int W, int win, int dmin, int D, float *__restrict__ cv) {
    const int half = win / 2;
    const long pix = (long)blockIdx.x * blockDim.y + threadIdx.y;
    if (pix >= (long)H * W) return;
    const int y = (int)(pix / W), x = (int)(pix % W);
    float *dst = cv + pix * D;
    const bool centre_ok = (y >= half && y < H - half && x >= half && x < W - half);
    const double n = (double)(win * win);
    for (int k = threadIdx.x; k < D; k += 32) {
        const int d = dmin + k;
        float out = nan_f();
        if (centre_ok && x - half + d >= 0 && x + half + d < W) {
            double s = 0.0;
            for (int dy = -half; dy <= half; ++dy)
                for (int dx = -half; dx <= half; ++dx) {
                    const float a = __ldg(L + (size_t)(y + dy) * W + x + dx);
                    const float b = __ldg(R + (size_t)(y + dy) * W + x + dx + d);
                    s += (double)(a * b);        // float32 product (zncc.py:210-213), float64 mean
                }
            double z = s / n - meanL[pix] * meanR[pix + d];
            const double den = stdL[pix] * stdR[pix + d];
            z = (den > 0.0) ? z / den : 0.0;     // zncc.py:244-277
            out = (float)z;
        }
        dst[k] = out;
    }
}

// ------------------------------------------------------------------------------------------------
// Tiled versions (window <= 13): a CTA owns one image row x TX pixels x all disparities.  The WIN image rows it
// needs are staged once in shared memory (left: TX + 2*half columns, right: the TX + 2*half + D - 1 columns the
// disparity range can reach); a thread owns 4 consecutive disparities of one pixel and keeps, per window row, a
// 4-wide register window of right pixels that slides by one column per window column -- one shared-memory read
// per (window row, window column) for four cells -- exactly the tap order of the plain kernels above
// (column offset outer, row offset inner), so the float results are unchanged bit for bit.
// ------------------------------------------------------------------------------------------------
constexpr int MC_TX = 32;

template <int WIN, int MODE>   // MODE 0: SAD, 1: SSD, 2: ZNCC (float64 accumulation of float32 products)
__global__ void __launch_bounds__(256) window_cost_tiled_kernel(const float *__restrict__ L, const float *__restrict__ R, int H, int W,
                                                                int dmin, int D, float *__restrict__ cv,
                                                                const double *__restrict__ meanL, const double *__restrict__ stdL,
                                                                const double *__restrict__ meanR, const double *__restrict__ stdR) {
    constexpr int HALF = WIN / 2;
    constexpr int LW = MC_TX + 2 * HALF;
    extern __shared__ __align__(16) float mc_smem[];
    const int RW = MC_TX + 2 * HALF + D + 3;          // right columns staged per window row
    float *sL = mc_smem;                              // [WIN][LW]
    float *sR = mc_smem + WIN * LW;                   // [WIN][RW]
    const int y = blockIdx.y, x0 = blockIdx.x * MC_TX;
    const int tid = threadIdx.x;
    const bool row_ok = (y >= HALF && y < H - HALF);
    const int npx = min(MC_TX, W - x0);
    const int G = (D + 3) >> 2;
    float *out_row = cv + ((size_t)y * W + x0) * D;
    if (!row_ok) {
        for (int i = tid; i < npx * D; i += blockDim.x) out_row[i] = nan_f();
        return;
    }
    const int xr0 = x0 - HALF + dmin;                 // image column of sR[.][0]
    for (int i = tid; i < WIN * LW; i += blockDim.x) {
        const int wy = i / LW, j = i % LW;
        const int xx = x0 - HALF + j;
        sL[i] = (xx >= 0 && xx < W) ? __ldg(L + (size_t)(y - HALF + wy) * W + xx) : 0.f;
    }
    for (int i = tid; i < WIN * RW; i += blockDim.x) {
        const int wy = i / RW, j = i % RW;
        const int xx = xr0 + j;
        sR[i] = (xx >= 0 && xx < W) ? __ldg(R + (size_t)(y - HALF + wy) * W + xx) : 0.f;
    }
    __syncthreads();
    for (int item = tid; item < npx * G; item += blockDim.x) {
        const int pp = item / G, g = item - pp * G;
        const int k0 = g * 4;
        const int x = x0 + pp;
        float res[4] = {nan_f(), nan_f(), nan_f(), nan_f()};
        if (x >= HALF && x < W - HALF) {
            // right window: win[wy][q] = R[y - HALF + wy][x - HALF + dxi + dmin + k0 + q] for the current dxi
            float win[WIN][4];
            const float *rbase = sR + pp + k0;        // column (x - HALF + dmin + k0) - xr0 = pp + k0
#pragma unroll
            for (int wy = 0; wy < WIN; ++wy)
#pragma unroll
                for (int q = 0; q < 3; ++q) win[wy][q + 1] = rbase[wy * RW + q];
            float accf[4];
            double accd[4];
#pragma unroll
            for (int dxi = 0; dxi < WIN; ++dxi) {
#pragma unroll
                for (int wy = 0; wy < WIN; ++wy) {
                    win[wy][0] = win[wy][1];
                    win[wy][1] = win[wy][2];
                    win[wy][2] = win[wy][3];
                    win[wy][3] = rbase[wy * RW + dxi + 3];
                    const float a = sL[wy * LW + pp + dxi];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (MODE == 2) {
                            const double t = (double)(a * win[wy][q]);           // float32 product, float64 sum
                            accd[q] = (dxi == 0 && wy == 0) ? t : accd[q] + t;
                        } else {
                            const float df = a - win[wy][q];
                            const float t = (MODE == 1) ? df * df : fabsf(df);
                            accf[q] = (dxi == 0 && wy == 0) ? t : accf[q] + t;
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int d = dmin + k0 + q;
                if (x - HALF + d >= 0 && x + HALF + d < W) {
                    if (MODE == 2) {
                        const size_t pix = (size_t)y * W + x;
                        double z = accd[q] / (double)(WIN * WIN) - meanL[pix] * meanR[pix + d];
                        const double den = stdL[pix] * stdR[pix + d];
                        z = (den > 0.0) ? z / den : 0.0;
                        res[q] = (float)z;
                    } else {
                        res[q] = accf[q];
                    }
                }
            }
        }
        float *dst = out_row + (size_t)pp * D + k0;
        if ((D & 3) == 0) {
            *reinterpret_cast<float4 *>(dst) = make_float4(res[0], res[1], res[2], res[3]);
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (k0 + q < D) dst[q] = res[q];
        }
    }
}

template <int MODE>
static int launch_window_cost(const float *L, const float *R, int H, int W, int win, int dmin, int D, float *cv, const double *mL,
                              const double *sL, const double *mR, const double *sR, cudaStream_t s, bool *done) {
    *done = false;
    if (win > 13 || (reinterpret_cast<uintptr_t>(cv) & 15)) return PB200_OK;
    const int half = win / 2;
    const size_t smem = (size_t)win * ((MC_TX + 2 * half) + (MC_TX + 2 * half + D + 3)) * sizeof(float);
    if (smem > 160 * 1024) return PB200_OK;
    dim3 grid(ceil_div(W, MC_TX), H);
#define PB200_W(WIN)                                                                                                         \
    case WIN:                                                                                                                \
        PB200_CUDA(cudaFuncSetAttribute(window_cost_tiled_kernel<WIN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        window_cost_tiled_kernel<WIN, MODE><<<grid, 256, smem, s>>>(L, R, H, W, dmin, D, cv, mL, sL, mR, sR);               \
        break;
    switch (win) {
        PB200_W(1) PB200_W(3) PB200_W(5) PB200_W(7) PB200_W(9) PB200_W(11) PB200_W(13)
        default: return PB200_OK;
    }
#undef PB200_W
    PB200_LAUNCH_CHECK("window_cost_tiled_kernel");
    *done = true;
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_sad_ssd_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin,
                                         int D, int squared, float *d_cv, void *stream) {
    if (!d_left || !d_right || !d_cv || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_sad_ssd_cost_volume: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (window < 1 || (window & 1) == 0) {
        set_error("pb200_sad_ssd_cost_volume: window_size %d must be odd and >= 1", window);
        return PB200_ERR_UNSUPPORTED;
    }
    bool done = false;
    int rc = squared ? launch_window_cost<1>(d_left, d_right, H, W, window, dmin, D, d_cv, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, &done)
                     : launch_window_cost<0>(d_left, d_right, H, W, window, dmin, D, d_cv, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, &done);
    if (rc != PB200_OK || done) return rc;
    dim3 block(32, 8);
    const int grid = ceil_div((long)H * W, 8);
    if (squared) sad_ssd_kernel<true><<<grid, block, 0, (cudaStream_t)stream>>>(d_left, d_right, H, W, window, dmin, D, d_cv);
    else sad_ssd_kernel<false><<<grid, block, 0, (cudaStream_t)stream>>>(d_left, d_right, H, W, window, dmin, D, d_cv);
    PB200_LAUNCH_CHECK("sad_ssd_kernel");
    return PB200_OK;
}

extern "C" size_t pb200_zncc_workspace_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return 4 * (size_t)H * W * sizeof(double);
}

extern "C" int pb200_zncc_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                                      float *d_cv, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_left || !d_right || !d_cv || !d_workspace || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_zncc_cost_volume: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (window < 1 || (window & 1) == 0) {
        set_error("pb200_zncc_cost_volume: window_size %d must be odd and >= 1", window);
        return PB200_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pb200_zncc_workspace_bytes(H, W)) {
        set_error("pb200_zncc_cost_volume: workspace too small");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n = (size_t)H * W;
    double *mL = (double *)d_workspace, *sL = mL + n, *mR = sL + n, *sR = mR + n;
    const int g1 = ceil_div((long)n, 256);
    zncc_stats_kernel<<<g1, 256, 0, s>>>(d_left, H, W, window, mL, sL);
    PB200_LAUNCH_CHECK("zncc_stats_kernel");
    zncc_stats_kernel<<<g1, 256, 0, s>>>(d_right, H, W, window, mR, sR);
    PB200_LAUNCH_CHECK("zncc_stats_kernel");
    bool done = false;
    int rc = launch_window_cost<2>(d_left, d_right, H, W, window, dmin, D, d_cv, mL, sL, mR, sR, s, &done);
    if (rc != PB200_OK || done) return rc;
    dim3 block(32, 8);
    zncc_kernel<<<ceil_div((long)n, 8), block, 0, s>>>(d_left, d_right, mL, sL, mR, sR, H, W, window, dmin, D, d_cv);
    PB200_LAUNCH_CHECK("zncc_kernel");
    return PB200_OK;
}
