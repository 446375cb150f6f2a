// sgm.cu -- 8-path semi-global matching (the step Pandora delegates to the libSGM plugin).
//
// Boundary: AbstractOptimization.optimize_cv (src/pandora/optimization/optimization.py:104-123),
// call site state_machine.py:415-419; behaviour documented in
// docs/source/userguide/plugins/plugin_libsgm.rst:9-146 (libSGM itself is not vendored:
// pyproject.toml:59-61, so parity is pinned against oracle/pandora_oracle.c::pbo_sgm only).
//
// Recurrence (Hirschmueller 2008), float32, evaluated in exactly the oracle's order:
//     m = min_k Lp[k];  t = min(Lp[d], min(Lp[d-1], Lp[d+1]) + P1);  t = min(t, m + P2);
//     L[d] = C[d] + (t - m)            (first pixel of a path: L = C; NaN costs -> invalid_value)
// S = sum of L over the directions in the order E, W, S, SE, SW, N, NE, NW; the last direction
// restores NaN, applies the overcounting correction and (optionally) takes the WTA argmin.
// With integer-valued costs and penalties every value is an exact small integer in float32.
//
// Mapping: ONE WARP PER PATH.  The D-vector of a pixel is one contiguous 4*D-byte segment (disparity
// is the fastest axis of Pandora's volume), lane l holds disparities [l*NPL, (l+1)*NPL) in
// registers, so a step is: one coalesced vector load of C (prefetched one pixel ahead), the
// previous pixel's L_r in registers, d+-1 neighbours across lanes through two warp shuffles,
// min_k through a 5-step shuffle reduction, and one coalesced read-modify-write of S.
// Every pixel belongs to exactly one path per direction, so S needs no atomics as long as the
// directions run one after the other on the stream.
#include "common.cuh"

namespace pb200 {

struct SgmParams {
    const float *cv;      // (H, W, D) input costs
    float *S;             // (H, W, D) accumulated / final costs
    int H, W, D;
    float p1, p2, invalid_value;
    int dy, dx;           // direction
    int mode;             // 0: first direction (S = L), 1: accumulate, 2: accumulate + finalise, 3: first and final at once
    int overcounting;
    float over_scale;     // n_directions - 1 (7 for the 8-path sum)
    const float *halo_in;   // (W, D) path states of the row just outside the tile for this direction, or NULL
    float *halo_out;        // (W, D) receives the states of this tile's last row in travel direction, or NULL
    float *disp;            // fused WTA (mode 2), or NULL
    uint8_t *all_nan;
    int dmin;
    float invalid_disparity;
};

template <int NPL, bool VEC>
__device__ __forceinline__ void load_vec(const float *__restrict__ base, int lane, int D, float (&v)[NPL]) {
    if (VEC) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const int d = lane * NPL + q * 4;
            float4 t = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
            if (d < D) t = *reinterpret_cast<const float4 *>(base + d);      // D % 4 == 0 in this mode
            v[q * 4 + 0] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int d = lane * NPL + j;
            v[j] = (d < D) ? base[d] : CUDART_INF_F;
        }
    }
}

template <int NPL, bool VEC>
__device__ __forceinline__ void store_vec(float *__restrict__ base, int lane, int D, const float (&v)[NPL]) {
    if (VEC) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const int d = lane * NPL + q * 4;
            if (d < D) *reinterpret_cast<float4 *>(base + d) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            const int d = lane * NPL + j;
            if (d < D) base[d] = v[j];
        }
    }
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

template <int NPL, bool VEC>
__global__ void __launch_bounds__(128) sgm_path_kernel(const SgmParams p) {
    const int lane = threadIdx.x & 31;
    const long path = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int H = p.H, W = p.W, D = p.D, dy = p.dy, dx = p.dx;
    // ---- path id -> first pixel -------------------------------------------------------------------
    int y, x;
    if (dy == 0) {                       // horizontal: one path per row
        if (path >= H) return;
        y = (int)path;
        x = dx > 0 ? 0 : W - 1;
    } else {
        const int yb = dy > 0 ? 0 : H - 1;          // entry row
        if (path < W) { y = yb; x = (int)path; }
        else {
            if (dx == 0) return;
            const long q = path - W;                // entries on the side column, rows 1..H-1 away from the entry row
            if (q >= H - 1) return;
            y = dy > 0 ? (int)q + 1 : H - 2 - (int)q;
            x = dx > 0 ? 0 : W - 1;
        }
    }
    const long stride = ((long)dy * W + dx) * (long)D;                  // elements between consecutive pixels of the path
    const float *c_ptr = p.cv + ((size_t)y * W + x) * D;
    float *s_ptr = p.S + ((size_t)y * W + x) * D;

    float Lp[NPL], craw[NPL], cnext[NPL];
    float m = 0.f;
    bool have_prev = false;
    // halo hand-over: the predecessor of an entry-row pixel lives in the neighbouring tile
    if (p.halo_in != nullptr && dy != 0 && y == (dy > 0 ? 0 : H - 1)) {
        const int px = x - dx;
        if (px >= 0 && px < W) {
            load_vec<NPL, VEC>(p.halo_in + (size_t)px * D, lane, D, Lp);
            float lm = Lp[0];
#pragma unroll
            for (int j = 1; j < NPL; ++j) lm = fminf(lm, Lp[j]);
            m = warp_min(lm);
            have_prev = true;
        }
    }
    load_vec<NPL, VEC>(c_ptr, lane, D, cnext);
    while (true) {
#pragma unroll
        for (int j = 0; j < NPL; ++j) craw[j] = cnext[j];
        const int ny = y + dy, nx = x + dx;
        const bool more = (ny >= 0 && ny < H && nx >= 0 && nx < W);
        if (more) load_vec<NPL, VEC>(c_ptr + stride, lane, D, cnext);      // prefetch the next pixel's costs
        float sacc[NPL];
        if (p.mode == 1 || p.mode == 2) load_vec<NPL, VEC>(s_ptr, lane, D, sacc);

        float L[NPL];
        if (!have_prev) {
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const float c = craw[j];
                L[j] = (c != c) ? p.invalid_value : c;
            }
            have_prev = true;
        } else {
            const float up = __shfl_up_sync(0xffffffffu, Lp[NPL - 1], 1);
            const float dn = __shfl_down_sync(0xffffffffu, Lp[0], 1);
            const float left_edge = (lane == 0) ? CUDART_INF_F : up;
            const float right_edge = (lane == 31) ? CUDART_INF_F : dn;
            const float mp2 = m + p.p2;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const float lo = (j == 0) ? left_edge : Lp[j - 1];
                const float hi = (j == NPL - 1) ? right_edge : Lp[j + 1];
                float t = fminf(Lp[j], fminf(lo, hi) + p.p1);
                t = fminf(t, mp2);
                const float c = craw[j];
                const float cc = (c != c) ? p.invalid_value : c;
                L[j] = cc + (t - m);
            }
        }
        float lm = L[0];
#pragma unroll
        for (int j = 1; j < NPL; ++j) lm = fminf(lm, L[j]);
        m = warp_min(lm);
#pragma unroll
        for (int j = 0; j < NPL; ++j) Lp[j] = L[j];

        // ---- accumulate / finalise ----------------------------------------------------------------
        if (p.mode == 0) {
            store_vec<NPL, VEC>(s_ptr, lane, D, L);
        } else if (p.mode == 1) {
#pragma unroll
            for (int j = 0; j < NPL; ++j) sacc[j] = sacc[j] + L[j];
            store_vec<NPL, VEC>(s_ptr, lane, D, sacc);
        } else {
            float bv = CUDART_INF_F;
            int bk = 0x7fffffff;
            bool any = false;
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const float c = craw[j];
                float s = (p.mode == 3) ? L[j] : sacc[j] + L[j];
                if (p.overcounting) s = s - p.over_scale * ((c != c) ? p.invalid_value : c);
                if (c != c) s = nan_f();
                sacc[j] = s;
                const int d = lane * NPL + j;
                if (d < D && s == s) {
                    any = true;
                    if (s < bv) { bv = s; bk = d; }
                }
            }
            store_vec<NPL, VEC>(s_ptr, lane, D, sacc);
            if (p.disp != nullptr) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
                    const bool oany = __shfl_xor_sync(0xffffffffu, (int)any, o) != 0;
                    if (ov < bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
                    any = any || oany;
                }
                if (lane == 0) {
                    const size_t pix = (size_t)y * W + x;
                    if (bv == CUDART_INF_F) bk = 0;
                    p.disp[pix] = any ? (float)(p.dmin + bk) : p.invalid_disparity;
                    if (p.all_nan) p.all_nan[pix] = any ? 0 : 1;
                }
            }
        }
        if (p.halo_out != nullptr && dy != 0 && y == (dy > 0 ? H - 1 : 0))
            store_vec<NPL, VEC>(p.halo_out + (size_t)x * D, lane, D, L);
        if (!more) break;
        y = ny; x = nx;
        c_ptr += stride;
        s_ptr += stride;
    }
}

template <int NPL>
static int launch_dir(const SgmParams &p, cudaStream_t s) {
    long paths = (p.dy == 0) ? p.H : ((p.dx == 0) ? p.W : (long)p.W + p.H - 1);
    const int grid = ceil_div(paths, 4);
    const bool vec = (NPL % 4 == 0) && (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.cv) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.S) & 15) == 0) &&
                     (!p.halo_in || (reinterpret_cast<uintptr_t>(p.halo_in) & 15) == 0) &&
                     (!p.halo_out || (reinterpret_cast<uintptr_t>(p.halo_out) & 15) == 0);
    if (vec) sgm_path_kernel<(NPL % 4 == 0 ? NPL : 4), true><<<grid, 128, 0, s>>>(p);
    else sgm_path_kernel<NPL, false><<<grid, 128, 0, s>>>(p);
    PB200_LAUNCH_CHECK("sgm_path_kernel");
    return PB200_OK;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_sgm_workspace_bytes(int H, int W, int D) {
    (void)H; (void)W; (void)D;
    return 16;   // the path-per-warp kernels keep all state in registers
}

extern "C" int pb200_sgm(const float *d_cv_in, float *d_cv_out, int H, int W, int D, float p1, float p2, float invalid_value,
                         int overcounting, int dir_mask, int init_final, const float *d_halo_in_top, const float *d_halo_in_bottom,
                         float *d_halo_out_bottom, float *d_halo_out_top, float *d_disp, int dmin, float invalid_disparity,
                         uint8_t *d_all_nan, void *d_workspace, size_t workspace_bytes, void *stream) {
    (void)d_workspace; (void)workspace_bytes;
    if (!d_cv_in || !d_cv_out || d_cv_in == d_cv_out || H <= 0 || W <= 0 || D <= 0 || (dir_mask & 0xFF) == 0) {
        set_error("pb200_sgm: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (D > PB200_SGM_MAX_DISP) {
        set_error("pb200_sgm: D=%d above the supported maximum (%d)", D, PB200_SGM_MAX_DISP);
        return PB200_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    // direction table in accumulation order; group 0 horizontal, 1 downward, 2 upward
    static const int dirs[8][3] = {{0, 1, 0}, {0, -1, 0}, {1, 0, 1}, {1, 1, 1}, {1, -1, 1}, {-1, 0, 2}, {-1, 1, 2}, {-1, -1, 2}};
    const size_t plane = (size_t)W * D;
    int first_dir = -1, last_dir = -1;
    for (int r = 0; r < 8; ++r)
        if (dir_mask & (1 << r)) {
            if (first_dir < 0) first_dir = r;
            last_dir = r;
        }
    for (int r = 0; r < 8; ++r) {
        const int group = dirs[r][2];
        if (!(dir_mask & (1 << r))) continue;
        SgmParams p;
        p.cv = d_cv_in; p.S = d_cv_out; p.H = H; p.W = W; p.D = D;
        p.p1 = p1; p.p2 = p2; p.invalid_value = invalid_value;
        p.dy = dirs[r][0]; p.dx = dirs[r][1];
        const bool is_init = (init_final & 1) && r == first_dir, is_final = (init_final & 2) && r == last_dir;
        p.mode = is_init ? (is_final ? 3 : 0) : (is_final ? 2 : 1);
        p.overcounting = overcounting;
        p.over_scale = 7.0f;
        p.halo_in = nullptr; p.halo_out = nullptr;
        if (group == 1) {
            if (d_halo_in_top) p.halo_in = d_halo_in_top + (size_t)(r - 2) * plane;
            if (d_halo_out_bottom) p.halo_out = d_halo_out_bottom + (size_t)(r - 2) * plane;
        } else if (group == 2) {
            if (d_halo_in_bottom) p.halo_in = d_halo_in_bottom + (size_t)(r - 5) * plane;
            if (d_halo_out_top) p.halo_out = d_halo_out_top + (size_t)(r - 5) * plane;
        }
        p.disp = is_final ? d_disp : nullptr;
        p.all_nan = is_final ? d_all_nan : nullptr;
        p.dmin = dmin; p.invalid_disparity = invalid_disparity;
        int rc;
        if (D <= 32) rc = launch_dir<1>(p, s);
        else if (D <= 64) rc = launch_dir<2>(p, s);
        else if (D <= 128) rc = launch_dir<4>(p, s);
        else if (D <= 256) rc = launch_dir<8>(p, s);
        else rc = launch_dir<16>(p, s);
        if (rc != PB200_OK) return rc;
    }
    return PB200_OK;
}
