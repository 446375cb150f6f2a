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
#include "sgm_common.cuh"

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
    const int *gate;        // when not NULL the kernel runs only if (*gate != 0) == (gate_run_if != 0) (narrow/wide path switch)
    int gate_run_if;
    uint16_t *dir_argmin;   // min_cost_paths: (H, W) index of this direction's minimal L_r per pixel (first minimum), or NULL
};

template <int NPL, bool VEC, bool FULL = false>
__device__ __forceinline__ void load_vec(const float *__restrict__ base, int lane, int D, float (&v)[NPL]) {
    if (VEC) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const int d = lane * NPL + q * 4;
            float4 t = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, CUDART_INF_F);
            if (FULL || d < D) t = *reinterpret_cast<const float4 *>(base + d);      // D % 4 == 0 in this mode
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

template <int NPL, bool VEC, bool FULL = false>
__device__ __forceinline__ void store_vec(float *__restrict__ base, int lane, int D, const float (&v)[NPL]) {
    if (VEC) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const int d = lane * NPL + q * 4;
            if (FULL || d < D) *reinterpret_cast<float4 *>(base + d) = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
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
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_run_if != 0)) return;
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
        if (p.dir_argmin != nullptr) {                          // min_cost_paths: where this path's own cost is minimal (first minimum)
            int bk = 0x7fffffff;
#pragma unroll
            for (int j = NPL - 1; j >= 0; --j)
                if (lane * NPL + j < D && L[j] == m) bk = lane * NPL + j;
            bk = __reduce_min_sync(0xffffffffu, bk);
            if (lane == 0) p.dir_argmin[(size_t)y * W + x] = (uint16_t)bk;
        }

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


// ------------------------------------------------------------------------------------------------
// Vertical sweep: the three directions that advance one row per step (S, SE, SW or N, NE, NW) in ONE
// launch.  Traffic per pixel drops from 3 x (read C + read S + write S) to read C + read S + write S.
//
// Mapping: the image is cut into column strips, one CTA per strip, all CTAs co-resident (cooperative
// launch) and walking the rows together.  A warp owns CPW adjacent columns and processes, per row, the
// three recurrences of each of its pixels: the vertical state stays in registers, the two diagonal
// states of the previous row come from the neighbouring column through shared memory (double
// buffered, one __syncthreads per row), and across a strip border through a 2-slot ring in global
// memory (L2 resident) guarded by a release/acquire progress counter.  The outgoing border state of a
// row is published early in the row step and consumed by the neighbour one step later, so the
// exchange latency is off the critical path as long as neighbouring strips stay within one row.
// ------------------------------------------------------------------------------------------------
struct SweepParams {
    const float *cv;
    float *S;
    int H, W, D;
    float p1, p2, invalid_value;
    int dy;                 // +1: S, SE, SW ; -1: N, NE, NW
    int mode;               // as SgmParams::mode, for the group as a whole
    int overcounting;
    float over_scale;
    const float *halo_in;   // (3, W, D) states of the row just outside the tile (order dx = 0, +1, -1) or NULL
    float *halo_out;        // (3, W, D) states of this tile's last row in travel direction or NULL
    float *disp;
    uint8_t *all_nan;
    int dmin;
    float invalid_disparity;
    unsigned long long *ring;   // [nstrips][2 sides][2 slots][32 * NPL] {tag, value} words, zero at launch
    const int *gate;            // see SgmParams::gate
    int gate_run_if;
};

// one recurrence step: L = cc + (min(Lp[d], min(Lp[d-1], Lp[d+1]) + P1, m + P2) - m), m = min_k Lp[k]
template <int NPL>
__device__ __forceinline__ void sgm_step_vec(const float (&cc)[NPL], const float (&Lp)[NPL], float (&L)[NPL], int lane, float p1,
                                             float p2) {
    float lm = Lp[0];
#pragma unroll
    for (int j = 1; j < NPL; ++j) lm = fminf(lm, Lp[j]);
    const float m = warp_min_redux(lm);
    const float up = __shfl_up_sync(0xffffffffu, Lp[NPL - 1], 1);
    const float dn = __shfl_down_sync(0xffffffffu, Lp[0], 1);
    const float left_edge = (lane == 0) ? CUDART_INF_F : up;
    const float right_edge = (lane == 31) ? CUDART_INF_F : dn;
    const float mp2 = m + p2;
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const float lo = (j == 0) ? left_edge : Lp[j - 1];
        const float hi = (j == NPL - 1) ? right_edge : Lp[j + 1];
        const float t = fmin3(Lp[j], fminf(lo, hi) + p1, mp2);
        L[j] = cc[j] + (t - m);
    }
}

// Warp roles: warps 0..nwarp-1 compute, warp nwarp is the exchange warp.  Compute warp w owns the strip
// columns w and K-1-w (mirror pair), so warp 0 owns BOTH border columns and computes their outgoing
// diagonal states first; the exchange warp then copies them to the ring (its release fence has no
// long-latency traffic of its own to wait for), polls the neighbours' counters and drops their border
// states into the halo columns of the shared-memory state buffer before the end-of-row barrier.
// FULL: D == 32 * NPL (no tail checks, implies VEC); MODE / WTA >= 0: compile-time copies of p.mode / (p.disp != NULL)
template <int NPL, bool VEC, bool FULL, int MODE, int WTA>
__global__ void __launch_bounds__(512, 1) sgm_vsweep_kernel(const SweepParams p) {
    if (p.gate != nullptr && (*p.gate != 0) != (p.gate_run_if != 0)) return;
    extern __shared__ __align__(16) float sweep_smem[];
    constexpr int VS = NPL * 32;                       // floats per state vector (padded to the warp)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x >> 5) - 1;
    const int K = nwarp * 2;                           // columns per strip
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int H = p.H, W = p.W, D = p.D, dy = p.dy;
    // shared: st[buf][diag][K + 2][VS]; diag 0: dx = +1, diag 1: dx = -1; column index = strip column + 1,
    // index 0 / K+1 = border states of the left / right neighbour strip
    auto st = [&](int buf, int diag, int col) -> float * { return sweep_smem + ((size_t)(buf * 2 + diag) * (K + 2) + col) * VS; };
    const bool has_left = strip > 0, has_right = strip + 1 < nstrips;
    const int mode = (MODE >= 0) ? MODE : p.mode;
    const bool accumulate = (mode == 1 || mode == 2), final = (mode >= 2);
    const bool do_wta = (WTA >= 0) ? (WTA != 0) : (p.disp != nullptr);
    const size_t plane = (size_t)W * D;

    if (warp == nwarp) {
        // ---------------- exchange warp: receive the neighbours' border states of row i -------------------
        for (int i = 0; i + 1 < H; ++i) {
            const int cur = i & 1;
            const uint32_t tag = (uint32_t)(i + 1);
            if (has_left) {
                float v[NPL];
                ll_recv<NPL>(p.ring + ((size_t)((strip - 1) * 2 + 1) * 2 + cur) * VS, lane, tag, v);
                lm_store<NPL>(st(cur, 0, 0), lane, v);
            }
            if (has_right) {
                float v[NPL];
                ll_recv<NPL>(p.ring + ((size_t)((strip + 1) * 2 + 0) * 2 + cur) * VS, lane, tag, v);
                lm_store<NPL>(st(cur, 1, K + 1), lane, v);
            }
            __syncthreads();
        }
        __syncthreads();                               // last row: nothing to exchange
        return;
    }

    // ---------------- compute warps -------------------------------------------------------------------
    const int col[2] = {warp, K - 1 - warp};           // strip columns of this warp (mirror pair)
    const int xs[2] = {strip * K + col[0], strip * K + col[1]};
    float Lv[2][NPL];                                  // vertical states of the previous row
    float cnext[2][NPL];
    int y = dy > 0 ? 0 : H - 1;
#pragma unroll
    for (int c = 0; c < 2; ++c)
        if (xs[c] < W) load_vec<NPL, VEC, FULL>(p.cv + ((size_t)y * W + xs[c]) * D, lane, D, cnext[c]);

    for (int i = 0; i < H; ++i, y += dy) {
        const int cur = i & 1, prv = cur ^ 1;
        const bool first = (i == 0), last = (i == H - 1);
        float cc[2][NPL], sacc[2][NPL];
        uint32_t nanmask[2] = {0u, 0u};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                const float v = cnext[c][j];
                const bool isn = (v != v);
                cc[c][j] = isn ? p.invalid_value : v;
                nanmask[c] |= isn ? (1u << j) : 0u;
            }
            const int x = xs[c];
            if (x < W) {
                if (!last) load_vec<NPL, VEC, FULL>(p.cv + ((size_t)(y + dy) * W + x) * D, lane, D, cnext[c]);
                if (accumulate) load_vec<NPL, VEC, FULL>(p.S + ((size_t)y * W + x) * D, lane, D, sacc[c]);
            }
        }
        // one recurrence of column slot c, group rank g (dx = 0, +1, -1): result in Lout, state stored for the next row
        auto run_dir = [&](const int c, const int g, float (&Lout)[NPL]) {
            const int x = xs[c];
            const int dx = (g == 0) ? 0 : (g == 1 ? 1 : -1);
            const int px = x - dx;
            float Lp[NPL];
            bool have = false;
            if (first) {
                if (p.halo_in != nullptr && px >= 0 && px < W) {
                    load_vec<NPL, VEC, FULL>(p.halo_in + (size_t)g * plane + (size_t)px * D, lane, D, Lp);
                    have = true;
                }
            } else if (g == 0) {
#pragma unroll
                for (int j = 0; j < NPL; ++j) Lp[j] = Lv[c][j];
                have = true;
            } else if (px >= 0 && px < W) {
                have = true;
                lm_load<NPL>(st(prv, g - 1, col[c] - dx + 1), lane, Lp);
            }
            if (have) {
                sgm_step_vec<NPL>(cc[c], Lp, Lout, lane, p.p1, p.p2);
            } else {
#pragma unroll
                for (int j = 0; j < NPL; ++j) Lout[j] = cc[c][j];
            }
            if (g == 0) {
#pragma unroll
                for (int j = 0; j < NPL; ++j) Lv[c][j] = Lout[j];
            } else {
                lm_store<NPL>(st(cur, g - 1, col[c] + 1), lane, Lout);
            }
            if (last && p.halo_out != nullptr) store_vec<NPL, VEC, FULL>(p.halo_out + (size_t)g * plane + (size_t)x * D, lane, D, Lout);
        };
        // the outgoing border diagonals first (column slot 0: dx = -1, slot 1: dx = +1), so that warp 0 can hand the
        // strip's border states to the exchange warp as early as possible
        float Lb[2][NPL];
        if (xs[0] < W) run_dir(0, 2, Lb[0]);
        if (warp == 0 && has_left && !last) ll_send<NPL>(p.ring + ((size_t)(strip * 2 + 0) * 2 + cur) * VS, lane, (uint32_t)(i + 1), Lb[0]);
        if (xs[1] < W) run_dir(1, 1, Lb[1]);
        if (warp == 0 && has_right && !last) ll_send<NPL>(p.ring + ((size_t)(strip * 2 + 1) * 2 + cur) * VS, lane, (uint32_t)(i + 1), Lb[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int x = xs[c];
            if (x >= W) continue;
            float L0[NPL], Lo[NPL];
            run_dir(c, 0, L0);
            run_dir(c, c == 0 ? 1 : 2, Lo);
            // ---- accumulate in the oracle's order (S/N, then dx = +1, then dx = -1) and finalise -----------
            float out[NPL];
#pragma unroll
            for (int j = 0; j < NPL; ++j) {
                float s = accumulate ? sacc[c][j] + L0[j] : L0[j];
                s = s + (c == 0 ? Lo[j] : Lb[1][j]);
                s = s + (c == 0 ? Lb[0][j] : Lo[j]);
                out[j] = s;
            }
            if (final) {
                float bv = CUDART_INF_F;
                int bk = 0x7fffffff;
                bool any = false;
#pragma unroll
                for (int j = 0; j < NPL; ++j) {
                    float s = out[j];
                    if (p.overcounting) s = s - p.over_scale * cc[c][j];
                    if (nanmask[c] & (1u << j)) s = nan_f();
                    out[j] = s;
                    const int d = lane * NPL + j;
                    if (do_wta && (FULL || d < D) && s == s) {
                        any = true;
                        if (s < bv) { bv = s; bk = d; }
                    }
                }
                if (do_wta) {
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
            store_vec<NPL, VEC, FULL>(p.S + ((size_t)y * W + x) * D, lane, D, out);
        }
        __syncthreads();           // every diagonal state of this row (own and halo) is in st[cur] before the next row
    }
}

template <int NPL>
static int launch_sweep(SweepParams p, void *workspace, size_t workspace_bytes, cudaStream_t s, bool *done) {
    *done = false;
    const int nsm = sm_count();
    int K = ceil_div(p.W, nsm);
    if (K < 4) K = 4;
    K = (K + 1) / 2 * 2;
    const int nwarp = K / 2;
    if (nwarp > 15) return PB200_OK;                               // image too wide for one co-resident wave: per-path kernels
    const int nstrips = ceil_div(p.W, K);
    const size_t VS = (size_t)NPL * 32;
    const size_t smem = (size_t)2 * 2 * (K + 2) * VS * sizeof(float);
    const size_t ring_bytes = (size_t)nstrips * 2 * 2 * VS * sizeof(unsigned long long);
    if (workspace == nullptr || workspace_bytes < ring_bytes || smem > 200 * 1024) return PB200_OK;
    const bool vec = (NPL % 4 == 0) && (p.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.cv) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(p.S) & 15) == 0) &&
                     (!p.halo_in || (reinterpret_cast<uintptr_t>(p.halo_in) & 15) == 0) &&
                     (!p.halo_out || (reinterpret_cast<uintptr_t>(p.halo_out) & 15) == 0);
    void (*kern)(const SweepParams) = vec ? sgm_vsweep_kernel<NPL, (NPL % 4 == 0), false, -1, -1> : sgm_vsweep_kernel<NPL, false, false, -1, -1>;
    if constexpr (NPL >= 4) {
        if (vec && p.D == NPL * 32) {                              // full-width vectors: specialised, branch-free variants
            const bool wta = p.disp != nullptr;
            switch (p.mode) {
                case 0: kern = sgm_vsweep_kernel<NPL, true, true, 0, 0>; break;
                case 1: kern = sgm_vsweep_kernel<NPL, true, true, 1, 0>; break;
                case 2: kern = wta ? sgm_vsweep_kernel<NPL, true, true, 2, 1> : sgm_vsweep_kernel<NPL, true, true, 2, 0>; break;
                default: kern = wta ? sgm_vsweep_kernel<NPL, true, true, 3, 1> : sgm_vsweep_kernel<NPL, true, true, 3, 0>; break;
            }
        }
    }
    const int threads = (nwarp + 1) * 32;
    PB200_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)kern, threads, smem));
    if ((long)per_sm * nsm < nstrips) return PB200_OK;             // cannot be co-resident
    p.ring = reinterpret_cast<unsigned long long *>(workspace);
    PB200_CUDA(cudaMemsetAsync(p.ring, 0, ring_bytes, s));                 // tag 0 = nothing published (row tags start at 1)
    void *args[] = {(void *)&p};
    PB200_CUDA(cudaLaunchCooperativeKernel((const void *)kern, dim3(nstrips), dim3(threads), args, smem, s));
    PB200_LAUNCH_CHECK("sgm_vsweep_kernel");
    *done = true;
    return PB200_OK;
}

int sgm_narrow_try(const float *cv, float *out, int H, int W, int D, float p1, float p2, float invalid_value, int overcounting,
                   float *disp, int dmin, float invalid_disparity, uint8_t *all_nan, void *workspace, size_t workspace_bytes,
                   cudaStream_t s, const int **gate, int phase, int dy, int final, const float *halo_in, float *halo_out);   // sgm_narrow.cu

// use_confidence (plugin_libsgm.rst:38-47): E(D) = sum_p C(p, D_p) * Confidence(p) + the penalty terms -- every cost of a pixel is
// scaled by that pixel's confidence before the recurrence; NaN costs stay NaN.  One warp-wide pass over the volume.
__global__ void __launch_bounds__(256) scale_volume_kernel(const float *__restrict__ cv, const float *__restrict__ conf, long n_pix, int D,
                                                          float *__restrict__ out) {
    const long total = n_pix * (long)D;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x)
        out[i] = cv[i] * __ldg(conf + i / D);
}

// min_cost_paths (plugin_libsgm.rst:411-413): "the number of sgm paths that give the same position for minimal optimized cost
// at each point" = how many of the 8 directions have the minimum of their own L_r at the disparity where the sum S is minimal
// (first minima; 0 for a pixel without any valid cost).
__global__ void __launch_bounds__(256) sgm_nb_directions_kernel(const uint16_t *__restrict__ dir_argmin, const float *__restrict__ disp,
                                                                const uint8_t *__restrict__ all_nan, long n_pix, int dmin,
                                                                float *__restrict__ nb) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pix) return;
    float n = 0.f;
    if (!all_nan[i]) {
        const int k = (int)disp[i] - dmin;
        int c = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) c += ((int)dir_argmin[(size_t)r * n_pix + i] == k) ? 1 : 0;
        n = (float)c;
    }
    nb[i] = n;
}

}  // namespace pb200

using namespace pb200;

extern "C" size_t pb200_sgm_workspace_bytes(int H, int W, int D) {
    (void)H;
    // strip-sweep exchange ring (launch_sweep / sgm_narrow.cu) + the narrow-path flag
    if (W <= 0 || D <= 0) return 16;
    return sgm_ring_max_bytes(W, D) + 512;
}

extern "C" size_t pb200_sgm_flag_offset(int W, int D) { return (W <= 0 || D <= 0) ? 0 : sgm_ring_max_bytes(W, D) + 256; }

extern "C" int pb200_sgm(const float *d_cv_in, float *d_cv_out, int H, int W, int D, float p1, float p2, float invalid_value,
                         int overcounting, int dir_mask, int init_final, const float *d_halo_in_top, const float *d_halo_in_bottom,
                         float *d_halo_out_bottom, float *d_halo_out_top, float *d_disp, int dmin, float invalid_disparity,
                         uint8_t *d_all_nan, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_cv_in || !d_cv_out || d_cv_in == d_cv_out || H <= 0 || W <= 0 || D <= 0 || (dir_mask & 0xFF) == 0) {
        set_error("pb200_sgm: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (D > PB200_SGM_MAX_DISP) {
        set_error("pb200_sgm: D=%d above the supported maximum (%d)", D, PB200_SGM_MAX_DISP);
        return PB200_ERR_UNSUPPORTED;
    }
    cudaStream_t s = (cudaStream_t)stream;
    note_path(STAGE_SGM, PATH_SGM_FLOAT);                      // overwritten below when a packed path is launched
    // Exact packed-integer fast path (sgm_narrow.cu) for a whole 8-direction call on integer-valued costs: it
    // verifies the data while it runs and raises a device flag when a cost is not a small integer; the float
    // kernels below are then enqueued gated on that flag (they return at once when the fast path succeeded).
    const int *gate = nullptr;
    const bool packed_ok = (init_final & 4) != 0, float_only = (init_final & 8) != 0;
    init_final &= 3;
    if (dir_mask == 0xFF && init_final == 3 && !d_halo_in_top && !d_halo_in_bottom && !d_halo_out_bottom && !d_halo_out_top) {
        int rc = sgm_narrow_try(d_cv_in, d_cv_out, H, W, D, p1, p2, invalid_value, overcounting, d_disp, dmin, invalid_disparity,
                                d_all_nan, d_workspace, workspace_bytes, s, &gate, 0, 1, 1, nullptr, nullptr);
        if (rc != PB200_OK) return rc;
    } else if (packed_ok) {
        // split calls of a row-tiled run (pandora_b200/tiling.py): horizontal pair first, then one vertical group per call
        int rc = PB200_OK;
        if (dir_mask == 0x03 && init_final == 1) {
            // first call: packed E + W only (they also verify the data).  The float E + W are NOT enqueued here: ranks
            // must first agree on the flag; the caller then repeats the call with bit 3 set, which enqueues only the
            // float kernels, gated on the (agreed) flag.
            rc = sgm_narrow_try(d_cv_in, d_cv_out, H, W, D, p1, p2, invalid_value, overcounting, nullptr, dmin, invalid_disparity, nullptr,
                                d_workspace, workspace_bytes, s, &gate, float_only ? 3 : 1, 1, 0, nullptr, nullptr);
            if (rc != PB200_OK) return rc;
            if (!float_only && gate != nullptr) return PB200_OK;
            if (float_only && gate == nullptr) return PB200_OK;     // not eligible: the first call already ran the float kernels
        } else if (dir_mask == 0x1C && !(init_final & 1))
            rc = sgm_narrow_try(d_cv_in, d_cv_out, H, W, D, p1, p2, invalid_value, overcounting, (init_final & 2) ? d_disp : nullptr, dmin,
                                invalid_disparity, (init_final & 2) ? d_all_nan : nullptr, d_workspace, workspace_bytes, s, &gate, 2, 1,
                                init_final & 2, d_halo_in_top, d_halo_out_bottom);
        else if (dir_mask == 0xE0 && !(init_final & 1))
            rc = sgm_narrow_try(d_cv_in, d_cv_out, H, W, D, p1, p2, invalid_value, overcounting, (init_final & 2) ? d_disp : nullptr, dmin,
                                invalid_disparity, (init_final & 2) ? d_all_nan : nullptr, d_workspace, workspace_bytes, s, &gate, 2, -1,
                                init_final & 2, d_halo_in_bottom, d_halo_out_top);
        else {
            set_error("pb200_sgm: packed intermediates (init_final bit 2) need dir_mask 0x03 (init), 0x1C or 0xE0");
            return PB200_ERR_BAD_ARG;
        }
        if (rc != PB200_OK) return rc;
    }
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
        // a complete vertical group (S, SE, SW or N, NE, NW) runs as one strip sweep when the image fits one
        // co-resident wave and the caller gave a workspace; otherwise direction by direction below
        if ((r == 2 || r == 5) && ((dir_mask >> r) & 7) == 7 && D <= 256) {
            SweepParams q;
            q.cv = d_cv_in; q.S = d_cv_out; q.H = H; q.W = W; q.D = D;
            q.p1 = p1; q.p2 = p2; q.invalid_value = invalid_value;
            q.dy = dirs[r][0];
            const bool g_init = (init_final & 1) && r == first_dir, g_final = (init_final & 2) && (r + 2) == last_dir;
            q.mode = g_init ? (g_final ? 3 : 0) : (g_final ? 2 : 1);
            q.overcounting = overcounting;
            q.over_scale = 7.0f;
            q.halo_in = (r == 2) ? d_halo_in_top : d_halo_in_bottom;
            q.halo_out = (r == 2) ? d_halo_out_bottom : d_halo_out_top;
            q.disp = g_final ? d_disp : nullptr;
            q.all_nan = g_final ? d_all_nan : nullptr;
            q.dmin = dmin; q.invalid_disparity = invalid_disparity;
            q.ring = nullptr;
            q.gate = gate; q.gate_run_if = 1;
            bool done = false;
            int rc;
            if (D <= 32) rc = launch_sweep<1>(q, d_workspace, workspace_bytes, s, &done);
            else if (D <= 64) rc = launch_sweep<2>(q, d_workspace, workspace_bytes, s, &done);
            else if (D <= 128) rc = launch_sweep<4>(q, d_workspace, workspace_bytes, s, &done);
            else rc = launch_sweep<8>(q, d_workspace, workspace_bytes, s, &done);
            if (rc != PB200_OK) return rc;
            if (done) { r += 2; continue; }
        }
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
        p.gate = gate; p.gate_run_if = 1;
        p.dir_argmin = nullptr;
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

extern "C" int pb200_scale_volume(const float *d_cv, const float *d_confidence, int H, int W, int D, float *d_out, void *stream) {
    if (!d_cv || !d_confidence || !d_out || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_scale_volume: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const long n_pix = (long)H * W;
    long blocks = (n_pix * D + 255) / 256;
    const long cap = (long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    scale_volume_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(d_cv, d_confidence, n_pix, D, d_out);
    PB200_LAUNCH_CHECK("scale_volume_kernel");
    return PB200_OK;
}

extern "C" size_t pb200_sgm_paths_workspace_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return (size_t)H * W * (8 * sizeof(uint16_t) + sizeof(float) + 1) + 64;
}

extern "C" int pb200_sgm_min_cost_paths(const float *d_cv_in, float *d_cv_out, int H, int W, int D, float p1, float p2, float invalid_value,
                                        int overcounting, float *d_nb_of_directions, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_cv_in || !d_cv_out || !d_nb_of_directions || !d_workspace || d_cv_in == d_cv_out || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_sgm_min_cost_paths: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (D > PB200_SGM_MAX_DISP) {
        set_error("pb200_sgm_min_cost_paths: D=%d above the supported maximum (%d)", D, PB200_SGM_MAX_DISP);
        return PB200_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pb200_sgm_paths_workspace_bytes(H, W)) {
        set_error("pb200_sgm_min_cost_paths: workspace too small (pb200_sgm_paths_workspace_bytes)");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const size_t n_pix = (size_t)H * W;
    uint16_t *argmin = reinterpret_cast<uint16_t *>(d_workspace);
    float *disp = reinterpret_cast<float *>(reinterpret_cast<char *>(d_workspace) + ((8 * n_pix * sizeof(uint16_t) + 15) & ~(size_t)15));
    uint8_t *all_nan = reinterpret_cast<uint8_t *>(disp + n_pix);
    static const int dirs[8][2] = {{0, 1}, {0, -1}, {1, 0}, {1, 1}, {1, -1}, {-1, 0}, {-1, 1}, {-1, -1}};
    note_path(STAGE_SGM, PATH_SGM_FLOAT, 8);
    for (int r = 0; r < 8; ++r) {                               // one path kernel per direction: each records its own argmin
        SgmParams p;
        p.cv = d_cv_in; p.S = d_cv_out; p.H = H; p.W = W; p.D = D;
        p.p1 = p1; p.p2 = p2; p.invalid_value = invalid_value;
        p.dy = dirs[r][0]; p.dx = dirs[r][1];
        p.mode = r == 0 ? 0 : (r == 7 ? 2 : 1);
        p.overcounting = overcounting;
        p.over_scale = 7.0f;
        p.halo_in = nullptr; p.halo_out = nullptr;
        p.disp = r == 7 ? disp : nullptr;
        p.all_nan = r == 7 ? all_nan : nullptr;
        p.dmin = 0; p.invalid_disparity = -1.f;
        p.gate = nullptr; p.gate_run_if = 1;
        p.dir_argmin = argmin + (size_t)r * n_pix;
        int rc;
        if (D <= 32) rc = launch_dir<1>(p, s);
        else if (D <= 64) rc = launch_dir<2>(p, s);
        else if (D <= 128) rc = launch_dir<4>(p, s);
        else if (D <= 256) rc = launch_dir<8>(p, s);
        else rc = launch_dir<16>(p, s);
        if (rc != PB200_OK) return rc;
    }
    sgm_nb_directions_kernel<<<ceil_div((long)n_pix, 256), 256, 0, s>>>(argmin, disp, all_nan, (long)n_pix, 0, d_nb_of_directions);
    PB200_LAUNCH_CHECK("sgm_nb_directions_kernel");
    return PB200_OK;
}
