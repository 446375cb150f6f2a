// api.cu -- library plumbing (errors, device query, launch counter) and the HOST-buffer entry points
// of the C-ABI declared in include/pandora_b200.h.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <vector>

#include "common.cuh"

namespace pb200 {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t err, const char *what) {
    if (err == cudaSuccess) return PB200_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(err));
    cudaGetLastError();                 // a reported error must not resurface in the launch check of an unrelated later call
    return PB200_ERR_CUDA;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

static std::atomic<int> g_options[OPT_COUNT];
static const char *const g_option_names[OPT_COUNT] = {"sgm.no_wave", "sgm.no_byte_tier", "sgm.wave_kernel", "census.direct", "census.tile",
                                                      "cbca.pipe", "cbca.bands", "reverse.gather", "fuse_census_sgm", "sad.taps"};
static bool g_options_init = [] {
    for (auto &o : g_options) o.store(-1);
    return true;
}();
int option(Option o) { return g_options[o].load(std::memory_order_relaxed); }

static thread_local int g_path[STAGE_COUNT], g_path_detail[STAGE_COUNT];
static const char *const g_stage_names[STAGE_COUNT] = {"sgm", "cbca", "census", "reverse", "sad"};
void note_path(Stage st, int path, int detail) { g_path[st] = path; g_path_detail[st] = detail; }

int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// RAII device buffer for the host entry points
struct DevBuf {
    void *p = nullptr;
    int alloc(size_t bytes) { return check_cuda(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc"); }
    ~DevBuf() { if (p) cudaFree(p); }
    template <typename T> T *as() { return reinterpret_cast<T *>(p); }
};

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_version(void) { return 100; }
extern "C" const char *pb200_last_error(void) { return g_err; }
extern "C" uint64_t pb200_kernel_launches(void) { return g_launches.load(); }
extern "C" int pb200_set_option(const char *name, int value) {
    (void)g_options_init;
    for (int i = 0; name && i < OPT_COUNT; ++i)
        if (strcmp(name, g_option_names[i]) == 0) {
            g_options[i].store(value);
            return PB200_OK;
        }
    set_error("pb200_set_option: unknown option %s", name ? name : "(null)");
    return PB200_ERR_BAD_ARG;
}
extern "C" int pb200_get_option(const char *name) {
    for (int i = 0; name && i < OPT_COUNT; ++i)
        if (strcmp(name, g_option_names[i]) == 0) return g_options[i].load();
    return -1;
}
extern "C" int pb200_last_path(const char *stage, int *detail) {
    for (int i = 0; stage && i < STAGE_COUNT; ++i)
        if (strcmp(stage, g_stage_names[i]) == 0) {
            if (detail) *detail = g_path_detail[i];
            return g_path[i];
        }
    return PB200_ERR_BAD_ARG;
}
// ---- peer-visible device memory for the tile links of a multi-GPU run (one process per GPU: CUDA IPC) ------------------------
extern "C" int pb200_ipc_alloc(size_t bytes, void **d_ptr, void *handle64) {
    if (!d_ptr || !handle64 || bytes == 0) {
        set_error("pb200_ipc_alloc: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the C-ABI passes IPC handles as 64 bytes");
    PB200_CUDA(cudaMalloc(d_ptr, bytes));
    PB200_CUDA(cudaMemset(*d_ptr, 0, bytes));
    PB200_CUDA(cudaDeviceSynchronize());
    PB200_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), *d_ptr));
    return PB200_OK;
}
extern "C" int pb200_ipc_open(const void *handle64, void **d_ptr) {
    if (!d_ptr || !handle64) {
        set_error("pb200_ipc_open: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    PB200_CUDA(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PB200_OK;
}
extern "C" int pb200_ipc_close(void *d_ptr) {
    if (d_ptr) PB200_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return PB200_OK;
}
extern "C" int pb200_ipc_free(void *d_ptr) {
    if (d_ptr) PB200_CUDA(cudaFree(d_ptr));
    return PB200_OK;
}

extern "C" int pb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

#define PB200_RC(call)                  \
    do {                                \
        int _r = (call);                \
        if (_r != PB200_OK) return _r;  \
    } while (0)

extern "C" int pb200_census_cost_volume_host(const float *left, const float *right, int H, int W, int window, const float *disps,
                                             int D, float *cv) {
    if (!left || !right || !disps || !cv || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_census_cost_volume_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const int dmin = (int)lroundf(disps[0]);                     // census.cpp:109
    const size_t img = (size_t)H * W * sizeof(float), vol = (size_t)H * W * D * sizeof(float);
    DevBuf dl, dr, dcv, ws;
    const size_t wsb = pb200_census_workspace_bytes(H, W, window);
    PB200_RC(dl.alloc(img)); PB200_RC(dr.alloc(img)); PB200_RC(dcv.alloc(vol)); PB200_RC(ws.alloc(wsb));
    PB200_CUDA(cudaMemcpy(dl.p, left, img, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemcpy(dr.p, right, img, cudaMemcpyHostToDevice));
    PB200_RC(pb200_census_cost_volume(dl.as<float>(), dr.as<float>(), H, W, window, dmin, D, dcv.as<float>(), ws.p, wsb, nullptr,
                                      0.f, nullptr, nullptr));
    PB200_CUDA(cudaMemcpy(cv, dcv.p, vol, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

// compute_matching_costs(img_left, imgs_right_shift, cv, disps, w, h) with the whole list of shifted right images
// (census.hpp:44-51, census.cpp:97-180): rights[0] is (H, W), rights[i > 0] are (H, W - 1), all contiguous float32.
extern "C" int pb200_census_cost_volume_multi_host(const float *left, const float *const *rights, int n_right, int H, int W, int window,
                                                   const float *disps, int n_disp, float *cv) {
    if (!left || !rights || n_right < 1 || !disps || !cv || H <= 0 || W <= 0 || n_disp <= 0) {
        set_error("pb200_census_cost_volume_multi_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (n_right == 1) return pb200_census_cost_volume_host(left, rights[0], H, W, window, disps, n_disp, cv);
    if (W < 2) {
        set_error("pb200_census_cost_volume_multi_host: shifted images need at least two columns");
        return PB200_ERR_BAD_ARG;
    }
    const int dmin = (int)lroundf(disps[0]);                     // census.cpp:109
    const size_t vol = (size_t)H * W * n_disp * sizeof(float);
    std::vector<DevBuf> imgs(n_right + 1);
    std::vector<const float *> ptrs(n_right);
    PB200_RC(imgs[0].alloc((size_t)H * W * 4));
    PB200_CUDA(cudaMemcpy(imgs[0].p, left, (size_t)H * W * 4, cudaMemcpyHostToDevice));
    for (int i = 0; i < n_right; ++i) {
        if (!rights[i]) { set_error("pb200_census_cost_volume_multi_host: NULL right image"); return PB200_ERR_BAD_ARG; }
        const size_t bytes = (size_t)H * (i == 0 ? W : W - 1) * 4;
        PB200_RC(imgs[i + 1].alloc(bytes));
        PB200_CUDA(cudaMemcpy(imgs[i + 1].p, rights[i], bytes, cudaMemcpyHostToDevice));
        ptrs[i] = imgs[i + 1].as<float>();
    }
    DevBuf dcv, ws;
    const size_t wsb = pb200_census_subpix_workspace_bytes(H, W, window, n_right);
    PB200_RC(dcv.alloc(vol)); PB200_RC(ws.alloc(wsb));
    PB200_RC(pb200_census_cost_volume_subpix(imgs[0].as<float>(), ptrs.data(), n_right, H, W, window, dmin, n_disp, dcv.as<float>(), ws.p,
                                             wsb, nullptr));
    PB200_CUDA(cudaMemcpy(cv, dcv.p, vol, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

extern "C" int pb200_reverse_cost_volume_host(const float *left_cv, int H, int W, int D, int min_disp, float *right_cv) {
    if (!left_cv || !right_cv || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_reverse_cost_volume_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const size_t vol = (size_t)H * W * D * sizeof(float);
    DevBuf a, b;
    PB200_RC(a.alloc(vol)); PB200_RC(b.alloc(vol));
    PB200_CUDA(cudaMemcpy(a.p, left_cv, vol, cudaMemcpyHostToDevice));
    PB200_RC(pb200_reverse_cost_volume(a.as<float>(), H, W, D, min_disp, b.as<float>(), nullptr));
    PB200_CUDA(cudaMemcpy(right_cv, b.p, vol, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

extern "C" int pb200_reverse_disp_range_host(const float *left_min, const float *left_max, int H, int W, float *right_min, float *right_max) {
    if (!left_min || !left_max || !right_min || !right_max || H <= 0 || W <= 0) {
        set_error("pb200_reverse_disp_range_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const size_t n = (size_t)H * W * sizeof(float);
    DevBuf a, b, c, d;
    PB200_RC(a.alloc(n)); PB200_RC(b.alloc(n)); PB200_RC(c.alloc(n)); PB200_RC(d.alloc(n));
    PB200_CUDA(cudaMemcpy(a.p, left_min, n, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemcpy(b.p, left_max, n, cudaMemcpyHostToDevice));
    PB200_RC(pb200_reverse_disp_range(a.as<float>(), b.as<float>(), H, W, c.as<float>(), d.as<float>(), nullptr));
    PB200_CUDA(cudaMemcpy(right_min, c.p, n, cudaMemcpyDeviceToHost));
    PB200_CUDA(cudaMemcpy(right_max, d.p, n, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

extern "C" int pb200_cross_support_host(const float *image, int H, int W, int len_arms, float intensity, int16_t *cross) {
    if (!image || !cross || H <= 0 || W <= 0) {
        set_error("pb200_cross_support_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    DevBuf a, b;
    PB200_RC(a.alloc((size_t)H * W * 4)); PB200_RC(b.alloc((size_t)H * W * 8));
    PB200_CUDA(cudaMemcpy(a.p, image, (size_t)H * W * 4, cudaMemcpyHostToDevice));
    PB200_RC(pb200_cross_support(a.as<float>(), H, W, W, len_arms, intensity, 0, b.as<int16_t>(), nullptr));
    PB200_CUDA(cudaMemcpy(cross, b.p, (size_t)H * W * 8, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

namespace pb200 {
int cbca_dispatch(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                  const int16_t *cr, int len_arms, cudaStream_t s);
__global__ void max_arm_kernel(const int16_t *__restrict__ a, const int16_t *__restrict__ b, long n, int *__restrict__ out) {
    int m = 0;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        m = max(m, max((int)a[i], (int)b[i]));
    atomicMax(out, m);
}
}  // namespace pb200

// One reference-style cbca() call (aggregation.cpp:323-355): a single (H, W) slice is a volume with
// D == 1 whose disparity is range_col_right[0] - range_col[0] (cbca.py:155-156 derives both index
// lists from one disparity); returns the un-normalised sums step4 and sum4.
extern "C" int pb200_cbca_host(const float *input, int H, int W, const int16_t *cross_left, const int16_t *cross_right,
                               const int64_t *range_col, const int64_t *range_col_right, int n, float *step4, float *sum4) {
    if (!input || !cross_left || !cross_right || !step4 || !sum4 || H <= 0 || W <= 0 || n < 0 ||
        (n > 0 && (!range_col || !range_col_right))) {
        set_error("pb200_cbca_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    int d = W;                                       // n == 0: no valid column at all
    if (n > 0) {
        d = (int)(range_col_right[0] - range_col[0]);
        if (range_col_right[n - 1] - range_col[n - 1] != d) {
            set_error("pb200_cbca_host: range_col_right - range_col must be one constant disparity");
            return PB200_ERR_UNSUPPORTED;
        }
    }
    const size_t px = (size_t)H * W;
    DevBuf din, de, dn, dcl, dcr, dmax;
    PB200_RC(din.alloc(px * 4)); PB200_RC(de.alloc(px * 4)); PB200_RC(dn.alloc(px * 4));
    PB200_RC(dcl.alloc(px * 8)); PB200_RC(dcr.alloc(px * 8)); PB200_RC(dmax.alloc(4));
    PB200_CUDA(cudaMemcpy(din.p, input, px * 4, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemcpy(dcl.p, cross_left, px * 8, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemcpy(dcr.p, cross_right, px * 8, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemset(dmax.p, 0, 4));
    max_arm_kernel<<<64, 256>>>(dcl.as<int16_t>(), dcr.as<int16_t>(), (long)px * 4, dmax.as<int>());
    PB200_LAUNCH_CHECK("max_arm_kernel");
    int max_arm = 0;
    PB200_CUDA(cudaMemcpy(&max_arm, dmax.p, 4, cudaMemcpyDeviceToHost));
    PB200_RC(cbca_dispatch(din.as<float>(), de.as<float>(), dn.as<float>(), H, W, 1, d, 0, dcl.as<int16_t>(), dcr.as<int16_t>(),
                           max_arm + 1, nullptr));
    PB200_CUDA(cudaMemcpy(step4, de.p, px * 4, cudaMemcpyDeviceToHost));
    PB200_CUDA(cudaMemcpy(sum4, dn.p, px * 4, cudaMemcpyDeviceToHost));
    return PB200_OK;
}

#ifndef PB200_FUSE_CENSUS_SGM_DEFAULT
#define PB200_FUSE_CENSUS_SGM_DEFAULT 1
#endif

// Whole pipeline on host images (the end-to-end path bench.py reports as `e2e`).
extern "C" int pb200_disparity_host(const float *left, const float *right, int H, int W, int method, int window, int dmin,
                                    int dmax, int cbca_distance, float cbca_intensity, float sgm_p1, float sgm_p2,
                                    int sgm_overcounting, float invalid_disparity, float *disp_map, uint16_t *validity_mask,
                                    float *cv_out) {
    if (!left || !right || !disp_map || H <= 0 || W <= 0 || dmax < dmin || method < 0 || method > 3) {
        set_error("pb200_disparity_host: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const int D = dmax - dmin + 1;
    const int off = (window - 1) / 2;
    const size_t px = (size_t)H * W, img = px * 4, vol = px * (size_t)D * 4;
    const bool do_cbca = cbca_distance > 0, do_sgm = sgm_p2 > 0.f;
    const bool is_max = (method == 3);
    DevBuf dl, dr, cva, cvb, ws, ddisp, dnan, dmask;
    PB200_RC(dl.alloc(img)); PB200_RC(dr.alloc(img)); PB200_RC(cva.alloc(vol));
    if (do_cbca || do_sgm) PB200_RC(cvb.alloc(vol));
    PB200_RC(ddisp.alloc(img)); PB200_RC(dnan.alloc(px)); PB200_RC(dmask.alloc(px * 2));
    PB200_CUDA(cudaMemcpy(dl.p, left, img, cudaMemcpyHostToDevice));
    PB200_CUDA(cudaMemcpy(dr.p, right, img, cudaMemcpyHostToDevice));
    float *cur = cva.as<float>(), *other = cvb.as<float>();
    bool have_disp = false;
    float cmax = 0.f;
    // ---- matching cost ----
    bool sgm_done = false;
    DevBuf sws;
    if (method == 0) {
        const size_t wsb = pb200_census_sgm_workspace_bytes(H, W, window, dmin, D);     // >= pb200_census_workspace_bytes
        PB200_RC(ws.alloc(wsb));
        // Census directly followed by SGM: one fused stage when eligible (the Census volume is never written)
        const int fopt = option(OPT_FUSE_CENSUS_SGM);
        if (do_sgm && !do_cbca && (fopt >= 0 ? fopt != 0 : PB200_FUSE_CENSUS_SGM_DEFAULT)) {
            const size_t swsb = pb200_sgm_workspace_bytes(H, W, D);
            PB200_RC(sws.alloc(swsb));
            int ran = 0;
            PB200_RC(pb200_census_sgm(dl.as<float>(), dr.as<float>(), H, W, window, dmin, D, sgm_p1, sgm_p2, sgm_overcounting, other, ws.p,
                                      wsb, sws.p, swsb, ddisp.as<float>(), invalid_disparity, dnan.as<uint8_t>(), 0, &ran, nullptr));
            if (ran) {
                PB200_CUDA(cudaDeviceSynchronize());
                float *t = cur; cur = other; other = t;
                sgm_done = have_disp = true;
            }
        }
    }
    if (method == 0) {
        if (!sgm_done) {
            const size_t wsb = pb200_census_workspace_bytes(H, W, window);
            const bool fuse = !do_cbca && !do_sgm;
            PB200_RC(pb200_census_cost_volume(dl.as<float>(), dr.as<float>(), H, W, window, dmin, D, cur, ws.p, wsb,
                                              fuse ? ddisp.as<float>() : nullptr, invalid_disparity,
                                              fuse ? dnan.as<uint8_t>() : nullptr, nullptr));
            have_disp = fuse;
        }
        cmax = (float)(window * window);
    } else if (method == 1 || method == 2) {
        PB200_RC(pb200_sad_ssd_cost_volume(dl.as<float>(), dr.as<float>(), H, W, window, dmin, D, method == 2, cur, nullptr));
        float mnl = left[0], mxl = left[0], mnr = right[0], mxr = right[0];       // cmax: sad_ssd.py:125-137
        for (size_t i = 1; i < px; ++i) {
            mnl = fminf(mnl, left[i]); mxl = fmaxf(mxl, left[i]); mnr = fminf(mnr, right[i]); mxr = fmaxf(mxr, right[i]);
        }
        const float m = fmaxf(fabsf(mxl - mnr), fabsf(mxr - mnl));
        cmax = floorf((method == 2 ? m * m : m) * (float)(window * window));
    } else {
        const size_t wsb = pb200_zncc_workspace_bytes(H, W);
        PB200_RC(ws.alloc(wsb));
        PB200_RC(pb200_zncc_cost_volume(dl.as<float>(), dr.as<float>(), H, W, window, dmin, D, cur, ws.p, wsb, nullptr));
        cmax = 1.f;
    }
    // ---- aggregation ----
    if (do_cbca) {
        const int Hi = H - 2 * off, Wi = W - 2 * off;
        if (Hi > 0 && Wi > 0) {
            DevBuf med, cl, cr;
            PB200_RC(med.alloc(img)); PB200_RC(cl.alloc((size_t)Hi * Wi * 8)); PB200_RC(cr.alloc((size_t)Hi * Wi * 8));
            PB200_RC(pb200_median3(dl.as<float>(), H, W, med.as<float>(), nullptr));
            PB200_RC(pb200_cross_support(med.as<float>() + (size_t)off * W + off, Hi, Wi, W, cbca_distance, cbca_intensity, 1,
                                         cl.as<int16_t>(), nullptr));
            PB200_RC(pb200_median3(dr.as<float>(), H, W, med.as<float>(), nullptr));
            PB200_RC(pb200_cross_support(med.as<float>() + (size_t)off * W + off, Hi, Wi, W, cbca_distance, cbca_intensity, 1,
                                         cr.as<int16_t>(), nullptr));
            PB200_RC(pb200_cbca_aggregate(cur, other, H, W, D, dmin, off, cl.as<int16_t>(), cr.as<int16_t>(), cbca_distance, nullptr));
            PB200_CUDA(cudaDeviceSynchronize());
            float *t = cur; cur = other; other = t;
        }
        cmax *= (float)((2 * cbca_distance - 1) * (2 * cbca_distance - 1));
    }
    // ---- optimisation ----
    if (do_sgm && !sgm_done) {
        if (is_max) {
            set_error("pb200_disparity_host: SGM on a max-type measure is not wired in the host pipeline");
            return PB200_ERR_UNSUPPORTED;
        }
        const size_t swsb = pb200_sgm_workspace_bytes(H, W, D);
        if (!sws.p) PB200_RC(sws.alloc(swsb));
        PB200_RC(pb200_sgm(cur, other, H, W, D, sgm_p1, sgm_p2, cmax + sgm_p2 + 1.f, sgm_overcounting, 0xFF, 3, nullptr, nullptr, nullptr,
                           nullptr, ddisp.as<float>(), dmin, invalid_disparity, dnan.as<uint8_t>(), sws.p, swsb, nullptr));
        PB200_CUDA(cudaDeviceSynchronize());          // the workspace must outlive the sweeps
        float *t = cur; cur = other; other = t;
        have_disp = true;
    }
    // ---- disparity ----
    if (!have_disp)
        PB200_RC(pb200_wta(cur, H, W, D, dmin, is_max ? 1 : 0, invalid_disparity, ddisp.as<float>(), dnan.as<uint8_t>(), nullptr));
    PB200_CUDA(cudaMemcpy(disp_map, ddisp.p, img, cudaMemcpyDeviceToHost));
    if (validity_mask) {
        PB200_RC(pb200_validity_mask_init(dmask.as<uint16_t>(), H, W, dmin, dmax, off, nullptr));
        PB200_RC(pb200_validity_mask(dmask.as<uint16_t>(), dnan.as<uint8_t>(), H, W, off, 0, nullptr));
        PB200_RC(pb200_validity_mask(dmask.as<uint16_t>(), dnan.as<uint8_t>(), H, W, off, 1, nullptr));
        PB200_CUDA(cudaMemcpy(validity_mask, dmask.p, px * 2, cudaMemcpyDeviceToHost));
    }
    if (cv_out) PB200_CUDA(cudaMemcpy(cv_out, cur, vol, cudaMemcpyDeviceToHost));
    PB200_CUDA(cudaDeviceSynchronize());
    return PB200_OK;
}
