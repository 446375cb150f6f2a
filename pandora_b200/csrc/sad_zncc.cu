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
                                                         double *__restrict__ mean, double *__restrict__ stdv,
                                                         double *__restrict__ coefA, double *__restrict__ coefB, double scaleA,
                                                         int2 *__restrict__ pack) {
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
    // running-sum kernel: zncc = S_LR * (A_L * A_R) - B_L * B_R with A = scale / std (scale = 1 / w^2 on the left image, 1 on
    // the right), B = mean / std, both 0 where std == 0 -- the reference writes 0 there (zncc.py:244-277)
    const double r = sd > 0.0 ? 1.0 / sd : 0.0;
    coefA[i] = r * scaleA;
    coefB[i] = m * r;
    // integer-valued images (the running-sum kernel's fast mode): the window sum S as an exact integer and 1 / (w^2 * std) as
    // float32 -- zncc = (w^2 * S_LR - S_L * S_R) * C_L * C_R has an exact integer numerator, so float32 factors are enough
    pack[i] = make_int2(__double2int_rn(m * (double)(win * win)), __float_as_int((float)(r / (double)(win * win))));
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

// ------------------------------------------------------------------------------------------------
// Running-sum version (window <= 13; `sad.taps` = 1 keeps the tap-ordered kernel above).  On integer-valued images
// (8 / 12 / 16-bit data stored as float32, the usual case: cones, the synthetic pairs) every partial sum of window terms
// is an integer below 2^24, so float32 additions AND subtractions are exact in any order and the w x w sum can be
// computed separably: a thread owns ONE disparity of a 16-column strip and walks down a band of rows keeping, per strip
// column, the vertical window sum V[j] in a register (V += term(row y + half) - term(row y - half - 1): two terms per
// cell and row instead of w^2), and a cell's cost is the horizontal running sum of 2*half + 1 V's (one add, one subtract
// per cell).  Lanes run over disparities: right-image reads are 32 consecutive shared-memory words, the left pixel is a
// broadcast float4, the store is one coalesced 128-byte line per pixel.  Rows are staged RUN_RB at a time into a
// shared-memory ring with cp.async, one block ahead of the block being computed; before a block is used every staged value is checked (integer, |v| <= vmax with (w^2 + w) * term(vmax) < 2^24): a
// CTA that meets anything else -- fractions, NaN, large values -- switches, for the rest of its band and before the value
// is used, to the reference's tap order (column offset outer, row offset inner) on the same ring, so float images stay
// bit-exact as well and no flag or second launch is needed.  ~11 FP32 instructions and 3 LDS per cell against 50 / 12.
// ------------------------------------------------------------------------------------------------
constexpr int RUN_TX = 16;
constexpr int RUN_RB = 8;
constexpr int RUN_RS = 32;      // ring rows: the block being computed (RUN_RB + WIN rows) plus the block in flight (RUN_RB)

template <int MODE>
__device__ __forceinline__ float run_term(float a, float b) {
    if (MODE == 2) return a * b;
    const float df = a - b;
    return MODE == 1 ? df * df : fabsf(df);
}
// 4-byte cp.async that writes zero when `ok` is false (src-size 0)
__device__ __forceinline__ void cp_async_f32_zfill(float *smem_dst, const float *gsrc, bool ok) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(ok ? 4 : 0) : "memory");
}

template <int WIN, int MODE>   // MODE 0: SAD, 1: SSD, 2: ZNCC
__global__ void __launch_bounds__(128) window_cost_running_kernel(const float *__restrict__ L, const float *__restrict__ R, int H, int W,
                                                                  int dmin, int D, int band, float vmax, float *__restrict__ cv,
                                                                  const double *__restrict__ coefAL, const double *__restrict__ coefBL,
                                                                  const double *__restrict__ coefAR, const double *__restrict__ coefBR,
                                                                  const int2 *__restrict__ packL, const int2 *__restrict__ packR) {
    constexpr int HALF = WIN / 2, LW = RUN_TX + 2 * HALF, LWP = (LW + 3) & ~3, RS = RUN_RS;
    extern __shared__ __align__(16) float run_smem[];
    const int DC = blockDim.x, RW = LWP + DC;
    float *sL = run_smem;                                        // [RS][LWP]
    float *sR = run_smem + RS * LWP;                             // [RS][RW]: index j + t for strip column j and thread t
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * RUN_TX, kc0 = blockIdx.z * DC, k = kc0 + t, d = dmin + k;
    const int yb0 = blockIdx.y * band, yb1 = min(H, yb0 + band);
    const int ya = max(yb0, HALF), yz = min(yb1, H - HALF);      // rows of the band whose window fits
    const int npx = min(RUN_TX, W - x0);
    const bool mine = k < D;
    if (mine)
        for (int y = yb0; y < yb1; ++y)
            if (y < ya || y >= yz || ya >= yz) {
                float *dst = cv + ((size_t)y * W + x0) * D + k;
                for (int p = 0; p < npx; ++p) __stcs(dst + (size_t)p * D, nan_f());
            }
    if (ya >= yz) return;
    const int xl0 = x0 - HALF, xr0 = x0 - HALF + dmin + kc0;
    // this thread's computable strip columns: half <= x < W - half, 0 <= x - half + d, x + half + d < W
    const int p_lo = max(max(HALF - x0, HALF - d - x0), 0);
    const int p_hi = mine ? min(min(W - HALF - x0, W - HALF - d - x0), npx) : 0;
    const bool full = __all_sync(0xffffffffu, p_lo == 0 && p_hi == RUN_TX);

    // rows [r0, r1) -> ring, asynchronously (LDGSTS); columns outside the image are written as zeros.  Per row a thread
    // copies right columns t and DC + t (< RW = DC + LWP) and left column t (< LWP): no index arithmetic beyond an add.
    const bool okL = t < LWP && t < LW && xl0 + t >= 0 && xl0 + t < W;
    const bool okR0 = xr0 + t >= 0 && xr0 + t < W, okR1 = t < LWP && xr0 + DC + t >= 0 && xr0 + DC + t < W;
    auto stage = [&](int r0, int r1) {
        for (int r = r0; r < r1; ++r) {
            const int slot = r & (RS - 1);
            const float *lrow = L + (size_t)r * W, *rrow = R + (size_t)r * W;
            if (t < LWP) {
                cp_async_f32_zfill(sL + slot * LWP + t, lrow + (okL ? xl0 + t : 0), okL);
                cp_async_f32_zfill(sR + slot * RW + DC + t, rrow + (okR1 ? xr0 + DC + t : 0), okR1);
            }
            cp_async_f32_zfill(sR + slot * RW + t, rrow + (okR0 ? xr0 + t : 0), okR0);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // the data condition on the elements this thread staged (re-read from the ring once they have landed)
    auto check = [&](int r0, int r1) {
        bool bad = false;
        for (int r = r0; r < r1; ++r) {
            const int slot = r & (RS - 1);
            const float v = sR[slot * RW + t];
            bad = bad || !(fabsf(v) <= vmax && v == rintf(v));
            if (t < LWP) {
                const float u = sL[slot * LWP + t], w = sR[slot * RW + DC + t];
                bad = bad || !(fabsf(u) <= vmax && u == rintf(u)) || !(fabsf(w) <= vmax && w == rintf(w));
            }
        }
        return bad;
    };

    float V[LWP];
    bool slow = false;
    int lo_b = ya - HALF, hi_b = min(ya + RUN_RB + HALF, yz + HALF);   // rows staged for the current block
    stage(lo_b, hi_b);
    for (int yblk = ya; yblk < yz; yblk += RUN_RB) {
        const int hi_n = min(yblk + 2 * RUN_RB + HALF, yz + HALF);     // the next block's new rows travel while this one is computed
        const bool more = hi_n > hi_b;
        if (more) stage(hi_b, hi_n);
        if (more) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        const bool bad = check(lo_b, hi_b);
        slow = __syncthreads_or(bad || slow) != 0;               // sticky and CTA-uniform; also publishes the staged rows
        lo_b = hi_b;
        hi_b = hi_n;
        const int yend = min(yblk + RUN_RB, yz);
        for (int y = yblk; y < yend; ++y) {
            float *dst = cv + ((size_t)y * W + x0) * D + k;
            const size_t pix0 = (size_t)y * W + x0;
            if (!slow) {
                if (y == ya) {
#pragma unroll
                    for (int j = 0; j < LWP; ++j) V[j] = 0.f;
                    for (int r = y - HALF; r <= y + HALF; ++r) {
                        const float *lr = sL + (r & (RS - 1)) * LWP, *rr = sR + (r & (RS - 1)) * RW + t;
#pragma unroll
                        for (int jg = 0; jg < LWP / 4; ++jg) {
                            const float4 a = *reinterpret_cast<const float4 *>(lr + 4 * jg);
                            V[4 * jg + 0] += run_term<MODE>(a.x, rr[4 * jg + 0]);
                            V[4 * jg + 1] += run_term<MODE>(a.y, rr[4 * jg + 1]);
                            V[4 * jg + 2] += run_term<MODE>(a.z, rr[4 * jg + 2]);
                            V[4 * jg + 3] += run_term<MODE>(a.w, rr[4 * jg + 3]);
                        }
                    }
                } else {
                    const int sn = (y + HALF) & (RS - 1), so = (y - HALF - 1) & (RS - 1);
                    const float *ln = sL + sn * LWP, *lo = sL + so * LWP, *rn = sR + sn * RW + t, *ro = sR + so * RW + t;
#pragma unroll
                    for (int jg = 0; jg < LWP / 4; ++jg) {
                        const float4 a = *reinterpret_cast<const float4 *>(ln + 4 * jg), b = *reinterpret_cast<const float4 *>(lo + 4 * jg);
                        V[4 * jg + 0] = (V[4 * jg + 0] + run_term<MODE>(a.x, rn[4 * jg + 0])) - run_term<MODE>(b.x, ro[4 * jg + 0]);
                        V[4 * jg + 1] = (V[4 * jg + 1] + run_term<MODE>(a.y, rn[4 * jg + 1])) - run_term<MODE>(b.y, ro[4 * jg + 1]);
                        V[4 * jg + 2] = (V[4 * jg + 2] + run_term<MODE>(a.z, rn[4 * jg + 2])) - run_term<MODE>(b.z, ro[4 * jg + 2]);
                        V[4 * jg + 3] = (V[4 * jg + 3] + run_term<MODE>(a.w, rn[4 * jg + 3])) - run_term<MODE>(b.w, ro[4 * jg + 3]);
                    }
                }
                float c = V[0];
#pragma unroll
                for (int j = 1; j <= 2 * HALF; ++j) c += V[j];
                if (full) {                                       // every cell of the warp's 16 x 32 tile is computable
#pragma unroll
                    for (int p = 0; p < RUN_TX; ++p) {
                        float res = c;
                        if (MODE == 2) {
                            {   // exact integer numerator w^2 * S_LR - S_L * S_R (|.| < 2^31 by the value bound), float32 factors 1 / (w^2 * std)
                                const int2 pl = __ldg(packL + pix0 + p), pr = __ldg(packR + pix0 + p + d);
                                const int num = (WIN * WIN) * __float2int_rn(c) - pl.x * pr.x;
                                res = ((float)num * __int_as_float(pl.y)) * __int_as_float(pr.y);
                            }
                        }
                        __stcs(dst + (size_t)p * D, res);
                        if (p + 1 < RUN_TX) c = (c + V[p + 2 * HALF + 1]) - V[p];
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < RUN_TX; ++p) {
                        float res = nan_f();
                        if (p >= p_lo && p < p_hi) {
                            if (MODE == 2) {
                                {   // exact integer numerator w^2 * S_LR - S_L * S_R (|.| < 2^31 by the value bound), float32 factors 1 / (w^2 * std)
                                const int2 pl = __ldg(packL + pix0 + p), pr = __ldg(packR + pix0 + p + d);
                                const int num = (WIN * WIN) * __float2int_rn(c) - pl.x * pr.x;
                                res = ((float)num * __int_as_float(pl.y)) * __int_as_float(pr.y);
                            }
                            } else {
                                res = c;
                            }
                        }
                        if (mine && p < npx) __stcs(dst + (size_t)p * D, res);
                        if (p + 1 < RUN_TX) c = (c + V[p + 2 * HALF + 1]) - V[p];
                    }
                }
            } else {
                for (int p = 0; p < npx; ++p) {
                    float res = nan_f();
                    if (p >= p_lo && p < p_hi) {
                        float accf = 0.f;
                        double accd = 0.0;
                        for (int dxi = 0; dxi < WIN; ++dxi)
                            for (int wy = 0; wy < WIN; ++wy) {
                                const int slot = (y - HALF + wy) & (RS - 1);
                                const float tm = run_term<MODE>(sL[slot * LWP + p + dxi], sR[slot * RW + p + dxi + t]);
                                if (MODE == 2) accd = (dxi == 0 && wy == 0) ? (double)tm : accd + (double)tm;
                                else accf = (dxi == 0 && wy == 0) ? tm : accf + tm;
                            }
                        if (MODE == 2) {
                            res = (float)fma(accd, __ldg(coefAL + pix0 + p) * __ldg(coefAR + pix0 + p + d), -(__ldg(coefBL + pix0 + p) * __ldg(coefBR + pix0 + p + d)));
                        } else {
                            res = accf;
                        }
                    }
                    if (mine) __stcs(dst + (size_t)p * D, res);
                }
            }
        }
        // the rows staged at the top of the NEXT iteration must not land on rows a slower warp still reads in this block
        if constexpr (3 * RUN_RB + WIN - 1 >= RUN_RS) __syncthreads();
    }
}

template <int MODE>
static int launch_window_running(const float *L, const float *R, int H, int W, int win, int dmin, int D, float *cv, const double *mL,
                                 const double *sL, const double *mR, const double *sR, cudaStream_t s, bool *done,
                                 const int2 *pkL = nullptr, const int2 *pkR = nullptr) {
    *done = false;
    if (win > 13 || option(OPT_SAD_TAPS) > 0) return PB200_OK;
    const int nw = ceil_div(D, 32), nchunk = ceil_div(nw, 4), DC = 32 * ceil_div(nw, nchunk);   // D = 192: two chunks of three warps
    const int half = win / 2, LWP = (RUN_TX + 2 * half + 3) & ~3, RS = RUN_RS;
    const size_t smem = (size_t)RS * (2 * LWP + DC) * sizeof(float);
    // largest |pixel| for which (w^2 + w) terms stay below 2^24: SAD term <= 2v, SSD <= 4v^2, ZNCC <= v^2
    const double room = 16777216.0 / (double)(win * win + win);
    // ZNCC also needs w^4 * v^2 < 2^31: its numerator w^2 * S_LR - S_L * S_R is formed in 32-bit integers
    const float vmax = MODE == 0 ? (float)floor(room / 2.0 - 1.0) : MODE == 1 ? (float)floor(sqrt(room / 4.0) - 1.0)
                                 : (float)floor(fmin(sqrt(room), 46340.0 / (double)(win * win)) - 1.0);
    const int gx = ceil_div(W, RUN_TX), gz = ceil_div(D, DC);
    int band = 64;                                       // shorter bands until the grid fills the SMs a few times over
    while (band > 16 && (long)gx * gz * ceil_div(H, band) < 4L * sm_count()) band >>= 1;
    dim3 grid(gx, ceil_div(H, band), gz);
#define PB200_W(WIN)                                                                                                         \
    case WIN:                                                                                                                \
        window_cost_running_kernel<WIN, MODE><<<grid, DC, smem, s>>>(L, R, H, W, dmin, D, band, vmax, cv, mL, sL, mR, sR, pkL, pkR);   \
        break;
    switch (win) {
        PB200_W(1) PB200_W(3) PB200_W(5) PB200_W(7) PB200_W(9) PB200_W(11) PB200_W(13)
        default: return PB200_OK;
    }
#undef PB200_W
    PB200_LAUNCH_CHECK("window_cost_running_kernel");
    note_path(STAGE_SAD, PATH_SAD_RUNNING, band);
    *done = true;
    return PB200_OK;
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
    int rc = squared ? launch_window_running<1>(d_left, d_right, H, W, window, dmin, D, d_cv, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, &done)
                     : launch_window_running<0>(d_left, d_right, H, W, window, dmin, D, d_cv, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, &done);
    if (rc != PB200_OK || done) return rc;
    note_path(STAGE_SAD, PATH_SAD_TAPS);
    rc = squared ? launch_window_cost<1>(d_left, d_right, H, W, window, dmin, D, d_cv, nullptr, nullptr, nullptr, nullptr, (cudaStream_t)stream, &done)
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
    return 10 * (size_t)H * W * sizeof(double);
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
    double *mL = (double *)d_workspace, *sL = mL + n, *mR = sL + n, *sR = mR + n, *aL = sR + n, *bL = aL + n, *aR = bL + n, *bR = aR + n;
    int2 *pkL = reinterpret_cast<int2 *>(bR + n), *pkR = pkL + n;
    const int g1 = ceil_div((long)n, 256);
    zncc_stats_kernel<<<g1, 256, 0, s>>>(d_left, H, W, window, mL, sL, aL, bL, 1.0 / (double)(window * window), pkL);
    PB200_LAUNCH_CHECK("zncc_stats_kernel");
    zncc_stats_kernel<<<g1, 256, 0, s>>>(d_right, H, W, window, mR, sR, aR, bR, 1.0, pkR);
    PB200_LAUNCH_CHECK("zncc_stats_kernel");
    bool done = false;
    int rc = launch_window_running<2>(d_left, d_right, H, W, window, dmin, D, d_cv, aL, bL, aR, bR, s, &done, pkL, pkR);
    if (rc != PB200_OK || done) return rc;
    note_path(STAGE_SAD, PATH_SAD_TAPS);
    rc = launch_window_cost<2>(d_left, d_right, H, W, window, dmin, D, d_cv, mL, sL, mR, sR, s, &done);
    if (rc != PB200_OK || done) return rc;
    dim3 block(32, 8);
    zncc_kernel<<<ceil_div((long)n, 8), block, 0, s>>>(d_left, d_right, mL, sL, mR, sR, H, W, window, dmin, D, d_cv);
    PB200_LAUNCH_CHECK("zncc_kernel");
    return PB200_OK;
}
