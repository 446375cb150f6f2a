// confidence.cu -- cost-volume confidence: ambiguity and risk.
//
// Replaces compute_ambiguity_and_sampled_ambiguity (src/pandora/cost_volume_confidence/cpp/src/ambiguity.cpp:28-142),
// compute_risk_and_sampled_risk (cost_volume_confidence/cpp/src/risk.cpp:28-197) and their helpers min_max_cost /
// searchsorted (cost_volume_confidence/cpp/src/cost_volume_confidence_tools.cpp:22-87).
//
// The reference normalises every cost with the GLOBAL extrema of the volume and then, per pixel, loops over the
// n_etas thresholds and over the D costs (n_etas * D comparisons per pixel, twice when risk follows ambiguity).
// Here: pass 1 (cv_extrema_kernel) reads the volume once for the per-pixel minimum and the global extrema; pass 2
// (confidence_kernel) reads it once more, one warp per pixel, and -- because ext + eta is non-decreasing in eta --
// finds for every cost the FIRST threshold it passes (index guessed from the even spacing of np.arange etas, then
// corrected against the true thresholds: one or two comparisons instead of n_etas), which gives the ambiguity integral
// directly and the per-eta samples / disparity extents through small per-warp shared-memory histograms followed by a
// prefix scan.  Ambiguity and risk share that single pass.  The comparison semantics are the reference's (ambiguity
// compares with float32 thresholds; risk with float64 ones, reproduced exactly by float32 thresholds rounded down),
// the normalisation is the same float32 expression, counts are exact integers, and the risk sums are either provably
// order-independent (integer counts, dyadic disparities) or accumulated in eta order like the reference, so all
// outputs are bit-identical.
#include <cmath>
#include <vector>

#include "common.cuh"

namespace pb200 {

namespace {

__device__ __forceinline__ uint32_t order_key(float f) {          // monotone float -> uint32
    const uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

struct Extrema { uint32_t min_key, max_key; };

template <bool VEC4>
__global__ void __launch_bounds__(256) cv_extrema_kernel(const float *__restrict__ cv, long n_pix, int D, int is_max,
                                                         float *__restrict__ min_img, Extrema *__restrict__ ext) {
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    float gmin = CUDART_INF_F, gmax = -CUDART_INF_F;
    for (long pix = warp0; pix < n_pix; pix += nwarps) {
        const float *p = cv + pix * D;
        float mn = CUDART_INF_F, mx = -CUDART_INF_F;
        bool any = false;
        auto take = [&](float v) {
            if (is_max) v = -v;
            if (v == v) { any = true; mn = fminf(mn, v); mx = fmaxf(mx, v); }
        };
        if (VEC4) {
            for (int k = lane * 4; k < D; k += 128) {
                const float4 v = ld_cs_f4(p + k);
                take(v.x); take(v.y); take(v.z); take(v.w);
            }
        } else {
            for (int k = lane; k < D; k += 32) take(__ldcs(p + k));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        any = __any_sync(0xffffffffu, any);
        if (any) { gmin = fminf(gmin, mn); gmax = fmaxf(gmax, mx); }
        if (lane == 0) min_img[pix] = any ? mn : nan_f();
    }
    if (lane == 0 && gmin <= gmax) {
        atomicMin(&ext->min_key, order_key(gmin));
        atomicMax(&ext->max_key, order_key(gmax));
    }
}

__global__ void extrema_init_kernel(Extrema *ext) {
    ext->min_key = order_key(CUDART_INF_F);
    ext->max_key = order_key(-CUDART_INF_F);
}

struct ConfParams {
    const float *cv;
    long n_pix;
    int D, is_max, n_etas;
    const float *min_img;
    const Extrema *ext;
    const float *etas_f;            // device, n_etas
    const double *etas_d;           // device, n_etas
    const int32_t *grids;           // (2, n_pix) or NULL
    const float *disparity_range;   // device, D
    float *amb, *samp_amb;          // outputs (optional)
    const float *samp_amb_in;       // risk input; NULL = the samples computed by this pass
    float *risk_max, *risk_min, *disp_sup, *disp_inf, *samp_risk_max, *samp_risk_min;
    int exact_sums, vec4;                 // the risk sums are exact in any order (integer counts, dyadic disparities): reduce in parallel
};

// searchsorted of cost_volume_confidence_tools.cpp:22-38 (lower bound, clamped to n - 1)
__device__ __forceinline__ int searchsorted_dev(const float *arr, int n, float value) {
    int left = 0, right = n - 1;
    while (left < right) {
        const int mid = left + (right - left) / 2;
        if (arr[mid] < value) left = mid + 1; else right = mid;
    }
    return left;
}

// per warp in shared memory: thr_r[-2 .. n+1] | thr_f[-2 .. n+1] floats (sentinels -inf, -inf | +inf, +inf) | hist | lo | hi ints
__host__ __device__ inline size_t conf_warp_bytes(int n) { return (size_t)2 * ((n + 4 + 3) / 4 * 4) * 4 + (size_t)3 * (n + 2) * 4; }

template <bool RISK>
__global__ void __launch_bounds__(256) confidence_kernel(const ConfParams p) {
    extern __shared__ __align__(16) unsigned char conf_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n = p.n_etas, D = p.D;
    unsigned char *base = conf_smem + (size_t)warp * ((conf_warp_bytes(n) + 15) / 16 * 16);
    // thr_r[e] = the double threshold of risk.cpp rounded DOWN to float: for a float x, x > T  <=>  x > rd(T), so the
    // inner loop never touches float64 (stored in the first half of the 8-byte slots the layout reserves)
    const int npad = (n + 4 + 3) / 4 * 4;                  // floats per threshold array incl. the four sentinels
    float *thr_r = reinterpret_cast<float *>(base) + 2;
    float *thr_f = thr_r + npad;
    int *hist = reinterpret_cast<int *>(thr_f - 2 + npad);
    int *lo = hist + (n + 2), *hi = lo + (n + 2);
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const float gmin = key_float(p.ext->min_key), gmax = key_float(p.ext->max_key);
    const float diff = gmax - gmin;
    const bool want_samples = p.samp_amb != nullptr || (RISK && p.samp_amb_in == nullptr);

    for (long pix = warp0; pix < p.n_pix; pix += nwarps) {
        const float ext = (p.min_img[pix] - gmin) / diff;
        if (ext != ext) {                                  // every cost NaN (or a constant volume): ambiguity.cpp:85-91, risk.cpp:88-107
            if (lane == 0) {
                if (p.amb) p.amb[pix] = (float)(n * D);
                if (RISK) { p.risk_max[pix] = p.risk_min[pix] = p.disp_sup[pix] = p.disp_inf[pix] = nan_f(); }
            }
            for (int e = lane; e < n; e += 32) {
                if (p.samp_amb) p.samp_amb[pix * n + e] = (float)D;
                if (RISK && p.samp_risk_max) { p.samp_risk_max[pix * n + e] = nan_f(); p.samp_risk_min[pix * n + e] = nan_f(); }
            }
            continue;
        }
        int imin = 0, imax = D;
        if (p.grids != nullptr) {
            imin = searchsorted_dev(p.disparity_range, D, (float)p.grids[pix]);
            imax = searchsorted_dev(p.disparity_range, D, (float)p.grids[p.n_pix + pix]) + 1;
        }
        __syncwarp();
        for (int e = lane; e < n; e += 32) {
            thr_f[e] = ext + p.etas_f[e];
            if (RISK) thr_r[e] = __double2float_rd((double)ext + p.etas_d[e]);
        }
        if (lane < 2) {
            thr_f[lane - 2] = thr_r[lane - 2] = -CUDART_INF_F;
            thr_f[n + lane] = thr_r[n + lane] = CUDART_INF_F;
        }
        for (int e = lane; e < n + 1; e += 32) { hist[e] = 0; lo[e] = 0x7fffffff; hi[e] = -1; }
        __syncwarp();
        int cnt = 0;
        const float *pc = p.cv + pix * D;
        // first threshold a cost passes: the thresholds are non-decreasing and (for np.arange etas) evenly spaced, so the
        // index is guessed from the spacing and then corrected against the true thresholds -- exact for any spacing,
        // one or two comparisons instead of a log2(n) search when the guess is good
        const float t0 = thr_f[0], tn = thr_f[n - 1];
        const float inv_step = (tn > t0) ? (float)(n - 1) / (tn - t0) : 0.f;
        // thr[-2], thr[-1] hold -inf and thr[n], thr[n+1] hold +inf (see the layout), so a 4-wide window around the
        // guess can be read without bounds checks.  first index = (thresholds below the window, assumed not passed) +
        // (not passed inside it); the assumption is verified on the window's ends and a plain loop handles the rest.
        auto first_pass = [&](const float *thr, float nv, int guess) -> int {
            const int w0 = guess - 2;                         // window = thr[w0 .. w0 + 3], w0 in [-2, n - 2]
            const bool q0 = nv <= thr[w0], q1 = nv <= thr[w0 + 1], q2 = nv <= thr[w0 + 2], q3 = nv <= thr[w0 + 3];
            int e = max(w0, 0) + (int)(w0 >= 0 && !q0) + (int)(w0 + 1 >= 0 && !q1) + (int)!q2 + (int)(w0 + 3 < n && !q3);
            const bool ok = (w0 <= 0 || !q0) && (w0 + 3 >= n - 1 || q3);
            if (!ok) {                                        // guess off by more than the window: exact search
                e = min(max(guess, 0), n);
                while (e > 0 && nv <= thr[e - 1]) --e;
                while (e < n && !(nv <= thr[e])) ++e;
            }
            return e;
        };
        auto one_cost = [&](float v, int k) {
            if (p.is_max) v = -v;
            const float nv = (v != v) ? ((k >= imin && k < imax) ? -CUDART_INF_F : CUDART_INF_F) : (v - gmin) / diff;
            const int guess = (int)fminf(fmaxf((nv - t0) * inv_step, 0.f), (float)n);
            const int a = first_pass(thr_f, nv, guess);
            cnt += n - a;
            if (want_samples && a < n) atomicAdd(&hist[a], 1);
            if (RISK) {
                const int c = RISK ? first_pass(thr_r, nv, a) : 0;   // risk.cpp:137 (not >): the float64 comparison, see thr_r
                if (c < n) { atomicMin(&lo[c], k); atomicMax(&hi[c], k); }
            }
        };
        if (p.vec4) {
            for (int k = lane * 4; k < D; k += 128) {
                const float4 v = *reinterpret_cast<const float4 *>(pc + k);
                one_cost(v.x, k); one_cost(v.y, k + 1); one_cost(v.z, k + 2); one_cost(v.w, k + 3);
            }
        } else {
            for (int k = lane; k < D; k += 32) one_cost(pc[k], k);
        }
        cnt = __reduce_add_sync(0xffffffffu, cnt);
        if (lane == 0 && p.amb) p.amb[pix] = (float)cnt;
        __syncwarp();
        // prefix over eta in chunks of 32: samples (sum), lowest / highest passing index (min / max)
        if (want_samples || RISK) {
            int carry_s = 0, carry_lo = 0x7fffffff, carry_hi = -1;
            for (int e0 = 0; e0 < n; e0 += 32) {
                const int e = e0 + lane;
                int s = (e < n) ? hist[e] : 0, l = (e < n) ? lo[e] : 0x7fffffff, h = (e < n) ? hi[e] : -1;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int ts = __shfl_up_sync(0xffffffffu, s, o), tl = __shfl_up_sync(0xffffffffu, l, o),
                              th = __shfl_up_sync(0xffffffffu, h, o);
                    if (lane >= o) { s += ts; l = min(l, tl); h = max(h, th); }
                }
                s += carry_s; l = min(l, carry_lo); h = max(h, carry_hi);
                if (e < n) { hist[e] = s; lo[e] = l; hi[e] = h; }
                carry_s = __shfl_sync(0xffffffffu, s, 31);
                carry_lo = __shfl_sync(0xffffffffu, l, 31);
                carry_hi = __shfl_sync(0xffffffffu, h, 31);
            }
            __syncwarp();
            if (p.samp_amb)
                for (int e = lane; e < n; e += 32) p.samp_amb[pix * n + e] = (float)hist[e];
        }
        if (RISK) {
            if (p.samp_risk_max) {
                for (int e = lane; e < n; e += 32) {
                    const float e_max = (float)hi[e] - (float)lo[e];
                    const float sa = p.samp_amb_in ? p.samp_amb_in[pix * n + e] : (float)hist[e];
                    p.samp_risk_max[pix * n + e] = e_max;
                    p.samp_risk_min[pix * n + e] = 1.f + e_max - sa;
                }
            }
            if (p.exact_sums && p.samp_amb_in == nullptr) {
                // every term is a small integer or a dyadic disparity: float32 addition is exact in any order, so the
                // eta-ordered sums of risk.cpp:128-175 can be reduced across the lanes
                float s_min = 0.f, s_max = 0.f, s_inf = 0.f, s_sup = 0.f;
                for (int e = lane; e < n; e += 32) {
                    const float e_max = (float)hi[e] - (float)lo[e];
                    s_sup += p.disparity_range[hi[e]];
                    s_inf += p.disparity_range[lo[e]];
                    s_min += 1.f + e_max - (float)hist[e];
                    s_max += e_max;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s_min += __shfl_xor_sync(0xffffffffu, s_min, o);
                    s_max += __shfl_xor_sync(0xffffffffu, s_max, o);
                    s_inf += __shfl_xor_sync(0xffffffffu, s_inf, o);
                    s_sup += __shfl_xor_sync(0xffffffffu, s_sup, o);
                }
                if (lane == 0) {
                    p.risk_min[pix] = s_min / (float)n;
                    p.risk_max[pix] = s_max / (float)n;
                    p.disp_sup[pix] = s_sup / (float)n;
                    p.disp_inf[pix] = s_inf / (float)n;
                }
            } else if (lane == 0) {                         // float32 sums in eta order, like risk.cpp:128-175
                float s_min = 0.f, s_max = 0.f, s_inf = 0.f, s_sup = 0.f;
                for (int e = 0; e < n; ++e) {
                    const float e_max = (float)hi[e] - (float)lo[e];
                    const float sa = p.samp_amb_in ? p.samp_amb_in[pix * n + e] : (float)hist[e];
                    s_sup += p.disparity_range[hi[e]];
                    s_inf += p.disparity_range[lo[e]];
                    s_min += 1.f + e_max - sa;
                    s_max += e_max;
                }
                p.risk_min[pix] = s_min / (float)n;
                p.risk_max[pix] = s_max / (float)n;
                p.disp_sup[pix] = s_sup / (float)n;
                p.disp_inf[pix] = s_inf / (float)n;
            }
        }
        __syncwarp();
    }
}

}  // namespace

}  // namespace pb200

using namespace pb200;

static size_t conf_align(size_t v) { return (v + 255) / 256 * 256; }

extern "C" size_t pb200_confidence_workspace_bytes(int H, int W, int n_etas) {
    if (H <= 0 || W <= 0 || n_etas <= 0) return 0;
    return conf_align((size_t)H * W * sizeof(float)) + 256 + conf_align((size_t)n_etas * sizeof(double)) + conf_align((size_t)n_etas * sizeof(float));
}

extern "C" int pb200_confidence(const float *d_cv, int H, int W, int D, int is_max, const double *etas, int n_etas,
                                const int32_t *d_grids, const float *d_disparity_range, float *d_ambiguity, float *d_sampled_ambiguity,
                                const float *d_sampled_ambiguity_in, float *d_risk_max, float *d_risk_min, float *d_disp_sup,
                                float *d_disp_inf, float *d_sampled_risk_max, float *d_sampled_risk_min, void *d_workspace,
                                size_t workspace_bytes, void *stream) {
    if (!d_cv || !etas || !d_disparity_range || !d_workspace || H <= 0 || W <= 0 || D <= 0 || n_etas <= 0) {
        set_error("pb200_confidence: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const bool risk = d_risk_max || d_risk_min || d_disp_sup || d_disp_inf;
    if (risk && !(d_risk_max && d_risk_min && d_disp_sup && d_disp_inf)) {
        set_error("pb200_confidence: the four risk outputs go together");
        return PB200_ERR_BAD_ARG;
    }
    if ((d_sampled_risk_max == nullptr) != (d_sampled_risk_min == nullptr) || (d_sampled_risk_max && !risk)) {
        set_error("pb200_confidence: sampled risk outputs need both arrays and the risk outputs");
        return PB200_ERR_BAD_ARG;
    }
    if (n_etas > 1024) {
        set_error("pb200_confidence: more than 1024 etas");
        return PB200_ERR_UNSUPPORTED;
    }
    if (!(etas[0] >= 0.0)) {
        set_error("pb200_confidence: etas must be non-negative");
        return PB200_ERR_UNSUPPORTED;
    }
    for (int e = 1; e < n_etas; ++e)
        if (!(etas[e] >= etas[e - 1])) {
            set_error("pb200_confidence: etas must be non-decreasing (np.arange(eta_min, eta_max, eta_step))");
            return PB200_ERR_UNSUPPORTED;
        }
    if (workspace_bytes < pb200_confidence_workspace_bytes(H, W, n_etas) || (reinterpret_cast<uintptr_t>(d_workspace) & 15)) {
        set_error("pb200_confidence: workspace too small or misaligned");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    char *ws = reinterpret_cast<char *>(d_workspace);
    float *min_img = reinterpret_cast<float *>(ws);
    Extrema *ext = reinterpret_cast<Extrema *>(ws + conf_align((size_t)H * W * sizeof(float)));
    double *etas_d = reinterpret_cast<double *>(reinterpret_cast<char *>(ext) + 256);
    float *etas_f = reinterpret_cast<float *>(reinterpret_cast<char *>(etas_d) + conf_align((size_t)n_etas * sizeof(double)));
    // the float etas are what py::array_t<float> hands to ambiguity.cpp (forcecast of the float64 np.arange)
    float etas_host_f[1024];
    for (int e = 0; e < n_etas; ++e) etas_host_f[e] = (float)etas[e];
    PB200_CUDA(cudaMemcpyAsync(etas_d, etas, (size_t)n_etas * sizeof(double), cudaMemcpyHostToDevice, s));
    PB200_CUDA(cudaMemcpyAsync(etas_f, etas_host_f, (size_t)n_etas * sizeof(float), cudaMemcpyHostToDevice, s));
    PB200_CUDA(cudaStreamSynchronize(s));                 // etas_host_f lives on this stack frame

    // are the risk sums order-independent?  counts are integers <= D <= 2^12; disparities must be multiples of 1/16 below
    // 2^12 (then every partial sum of <= 128 terms is an integer multiple of 1/16 below 2^23 / 16: exact in float32)
    bool exact_sums = risk && n_etas <= 128 && D <= 4096;
    if (exact_sums) {
        std::vector<float> dr((size_t)D);
        PB200_CUDA(cudaMemcpyAsync(dr.data(), d_disparity_range, (size_t)D * sizeof(float), cudaMemcpyDeviceToHost, s));
        PB200_CUDA(cudaStreamSynchronize(s));
        for (int k = 0; k < D && exact_sums; ++k)
            exact_sums = std::fabs(dr[k]) < 4096.f && dr[k] * 16.f == std::floor(dr[k] * 16.f);
    }
    const long n_pix = (long)H * W;
    long blocks = (n_pix + 7) / 8;
    const long cap = (long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_cv) & 15) == 0);
    extrema_init_kernel<<<1, 1, 0, s>>>(ext);
    PB200_LAUNCH_CHECK("extrema_init_kernel");
    if (vec) cv_extrema_kernel<true><<<(int)blocks, 256, 0, s>>>(d_cv, n_pix, D, is_max, min_img, ext);
    else cv_extrema_kernel<false><<<(int)blocks, 256, 0, s>>>(d_cv, n_pix, D, is_max, min_img, ext);
    PB200_LAUNCH_CHECK("cv_extrema_kernel");

    ConfParams p;
    p.cv = d_cv; p.n_pix = n_pix; p.D = D; p.is_max = is_max; p.n_etas = n_etas;
    p.min_img = min_img; p.ext = ext; p.etas_f = etas_f; p.etas_d = etas_d;
    p.grids = d_grids; p.disparity_range = d_disparity_range;
    p.amb = d_ambiguity; p.samp_amb = d_sampled_ambiguity; p.samp_amb_in = d_sampled_ambiguity_in;
    p.risk_max = d_risk_max; p.risk_min = d_risk_min; p.disp_sup = d_disp_sup; p.disp_inf = d_disp_inf;
    p.samp_risk_max = d_sampled_risk_max; p.samp_risk_min = d_sampled_risk_min;
    p.vec4 = vec ? 1 : 0;
    p.exact_sums = exact_sums ? 1 : 0;
    const size_t smem = 8 * ((conf_warp_bytes(n_etas) + 15) / 16 * 16);
    if (smem > 200 * 1024) {
        set_error("pb200_confidence: too many etas for the shared-memory histograms");
        return PB200_ERR_UNSUPPORTED;
    }
    if (risk) {
        PB200_CUDA(cudaFuncSetAttribute((const void *)confidence_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        confidence_kernel<true><<<(int)blocks, 256, smem, s>>>(p);
    } else {
        PB200_CUDA(cudaFuncSetAttribute((const void *)confidence_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        confidence_kernel<false><<<(int)blocks, 256, smem, s>>>(p);
    }
    PB200_LAUNCH_CHECK("confidence_kernel");
    return PB200_OK;
}
