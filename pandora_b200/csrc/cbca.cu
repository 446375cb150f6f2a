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
#include <cstdlib>
#include <type_traits>

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

// ---- aggregation, pipelined version (MA <= 8) -------------------------------------------------------------
// Same arithmetic as cbca_aggregate_kernel, but every global read of the row loop (cost row incl. its halo
// columns, both support rows) is staged PF rows ahead into a shared-memory ring with cp.async (LDGSTS), so the
// march down the rows never waits for HBM, the centre cost is taken from the ring instead of being re-read, and
// one __syncthreads per row is enough.
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gmem_src, bool pred) {
    const int n = pred ? 4 : 0;                                  // src-size 0: zero fill, nothing is read
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src, bool pred) {
    const int n = pred ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int MA, bool RAW>
__global__ void __launch_bounds__(32 * CBCA_TX) cbca_aggregate_pipe_kernel(const float *__restrict__ cv_in, float *__restrict__ cv_out,
                                                                           float *__restrict__ out_n, int H, int W, int D, int dmin,
                                                                           int off, const short4 *__restrict__ crossL,
                                                                           const short4 *__restrict__ crossR) {
    constexpr int TX = CBCA_TX;
    constexpr int RING = 4 * MA;                 // prefix rings: >= 2*MA + 2 rows
    constexpr int NS = 16;                       // staged rows: MA + 2 + PF
    constexpr int PF = NS - MA - 2;              // prefetch distance in rows
    constexpr int CW = TX + 2 * MA;              // staged columns per row
    constexpr int NL = (CW + TX - 1) / TX;       // staged cells per thread and row
    constexpr int XRN = TX + 31;                 // right-image supports a CTA can meet per row
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *crow = reinterpret_cast<float *>(smem_raw);                       // [NS][CW][32]  staged costs, then NaN -> 0
    float *hpre = crow + NS * CW * 32;                                       // [CW + 1][32]   row prefix of the current row
    float *PE = hpre + (CW + 1) * 32;                                        // [RING][TX][32]
    int *PN = reinterpret_cast<int *>(PE + RING * TX * 32);                  // [RING][TX][32]
    short4 *XL = reinterpret_cast<short4 *>(PN + RING * TX * 32);            // [NS][TX]
    short4 *XR = XL + NS * TX;                                               // [NS][XRN]
    unsigned char *isn = reinterpret_cast<unsigned char *>(XR + NS * XRN);   // [NS][TX][32]  centre cost was NaN

    const int Hi = H - 2 * off, Wi = W - 2 * off;
    const int lane = threadIdx.x, cx = threadIdx.y, tid = cx * 32 + lane;
    const int k0 = blockIdx.y * 32, k = k0 + lane;
    const int x0 = blockIdx.x * TX;
    const int x = x0 + cx;
    const int d = dmin + k;
    const int xr = x + d;
    const bool active = (k < D) && (x < Wi);
    const bool valid_col = active && xr >= 0 && xr < Wi;
    const int ring_idx = cx * 32 + lane;
    const int xr0 = x0 + dmin + k0;               // right column of XR entry 0
    const size_t row_elems = (size_t)W * D;

    // per-thread staging sources (row 0), advanced by one row per staged row
    const float *src[NL];
    bool src_ok[NL];
#pragma unroll
    for (int q = 0; q < NL; ++q) {
        const int jj = cx + q * TX, xx = x0 - MA + jj;
        src_ok[q] = (jj < CW) && (k < D) && xx >= 0 && xx < Wi;
        src[q] = src_ok[q] ? cv_in + ((size_t)off * W + (xx + off)) * D + k : cv_in;
    }
    const bool xl_ok = tid < TX && x0 + tid < Wi;
    const short4 *xl_src = xl_ok ? crossL + x0 + tid : crossL;
    const int xre = tid - 32;
    const bool xr_ok = tid >= 32 && xre < XRN && xr0 + xre >= 0 && xr0 + xre < Wi;
    const short4 *xr_src = xr_ok ? crossR + xr0 + xre : crossR;
    float *out_ptr = cv_out + ((size_t)off * W + (x + off)) * D + k;       // row yo = 0
    float *outn_ptr = RAW ? out_n + ((size_t)off * W + (x + off)) * D + k : nullptr;

    auto stage_row = [&](int r) {                 // enqueue the copies of row r (r < Hi) into ring stage r % NS
        const int sg = r & (NS - 1);
#pragma unroll
        for (int q = 0; q < NL; ++q) {
            const int jj = cx + q * TX;
            if (jj < CW) cp_async4(crow + (sg * CW + jj) * 32 + lane, src[q], src_ok[q]);
            if (src_ok[q]) src[q] += row_elems;
        }
        if (tid < TX) cp_async8(XL + sg * TX + tid, xl_src, xl_ok);
        if (tid >= 32 && xre < XRN) cp_async8(XR + sg * XRN + xre, xr_src, xr_ok);
        if (xl_ok) xl_src += Wi;
        if (xr_ok) xr_src += Wi;
    };

    for (int r = 0; r < PF; ++r) {
        if (r < Hi) stage_row(r);
        cp_async_commit();
    }
    float pe_run = 0.f;
    int pn_run = 0;
    for (int i = 0; i < Hi + MA; ++i) {
        if (i + PF < Hi) stage_row(i + PF);
        cp_async_commit();
        cp_async_wait<PF>();                      // this thread's copies of row i have landed
        const int sgi = i & (NS - 1);
        if (i < Hi) {
            // clean the cells this thread staged: NaN -> 0 (step 1 does not propagate NaN), remember the centre NaNs
#pragma unroll
            for (int q = 0; q < NL; ++q) {
                const int jj = cx + q * TX;
                if (jj < CW) {
                    float *cellp = crow + (sgi * CW + jj) * 32 + lane;
                    const float v = *cellp;
                    const bool nn = (v != v);
                    if (nn) *cellp = 0.f;
                    if (jj >= MA && jj < MA + TX) isn[(sgi * TX + jj - MA) * 32 + lane] = nn ? 1 : 0;
                }
            }
        }
        __syncthreads();
        // ---- row prefix over the staged columns (warp 0: lane k scans its CW cells) ------------------------
        if (i < Hi && cx == 0) {
            float v[CW];
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = crow[(sgi * CW + j) * 32 + lane];
            float run = 0.f;
            hpre[lane] = 0.f;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                run = run + v[j];
                hpre[(j + 1) * 32 + lane] = run;
            }
        }
        __syncthreads();
        // ---- stage 1: horizontal arm sums of row i ---------------------------------------------------
        if (i < Hi) {
            float eh = 0.f;
            int nh = 0;
            if (valid_col) {
                const short4 a = XL[sgi * TX + cx];
                const short4 b = XR[sgi * XRN + cx + lane];
                const int l = min(min((int)a.x, (int)b.x), MA), r = min(min((int)a.y, (int)b.y), MA);
                eh = hpre[(cx + MA + r + 1) * 32 + lane] - hpre[(cx + MA - l) * 32 + lane];
                nh = l + r;
            }
            pe_run = pe_run + eh;                                           // the reference's step-3 column prefix
            pn_run += nh;
            PE[(i & (RING - 1)) * TX * 32 + ring_idx] = pe_run;
            PN[(i & (RING - 1)) * TX * 32 + ring_idx] = pn_run;
        }
        // ---- stage 2: vertical arm sums of row yo = i - MA (its staged supports are still in the ring) --------
        const int yo = i - MA;
        if (yo >= 0 && active) {
            const int sg = yo & (NS - 1);
            float e = 0.f;
            int n = 1;
            if (valid_col) {
                const short4 a = XL[sg * TX + cx];
                const short4 b = XR[sg * XRN + cx + lane];
                const int t = min(min((int)a.z, (int)b.z), MA), bo = min(min((int)a.w, (int)b.w), MA);
                const int r1 = yo + bo, r0 = yo - t - 1;
                const float e1 = PE[(r1 & (RING - 1)) * TX * 32 + ring_idx];
                const int n1 = PN[(r1 & (RING - 1)) * TX * 32 + ring_idx];
                const float e0 = (r0 >= 0) ? PE[(r0 & (RING - 1)) * TX * 32 + ring_idx] : 0.f;
                const int n0 = (r0 >= 0) ? PN[(r0 & (RING - 1)) * TX * 32 + ring_idx] : 0;
                e = e1 - e0;
                n = n1 - n0 + t + bo + 1;
            }
            const bool cnan = isn[(sg * TX + cx) * 32 + lane] != 0;
            if (RAW) {
                *out_ptr = e;
                *outn_ptr = (float)(n - 1);
                outn_ptr += row_elems;
            } else {
                *out_ptr = cnan ? nan_f() : e / (float)n;                   // (0*c + E) / N
            }
            out_ptr += row_elems;
        }
    }
}

// ---- aggregation, register version (MA <= 4: cbca_distance <= 5, the default) -------------------------------
// Same arithmetic again, organised so that NOTHING is shared between threads: no block barrier, no staging ring, no
// serial row scan.  Lanes run over 32 disparities (every volume access of a warp is one 128-byte line), a thread owns
// TX = 4 adjacent columns and marches down the rows.  Per row it loads the 4 + 2*MA costs its horizontal arm sums can
// meet straight into registers (one row ahead: the loads of row i + 1 are in flight while row i is computed), adds the
// arm taps with predicates (l, r <= 4: at most eight predicated adds around the centre), keeps the running column
// prefixes of Eh (float32, the reference's step-3 order of additions) and Nh in registers and their last RING = 12
// rows in a THREAD-PRIVATE shared-memory ring ([slot][column][lane]: conflict-free, no synchronisation), from which
// the vertical arm sums of row i - MA are two differences.  The vertical arms and the "centre cost is NaN" bits of the
// last MA rows travel in two small register histories, so every support entry and every cost is loaded exactly once
// per thread (halo columns: twice more from L2 by the neighbouring strips).
// Integer costs: every sum is exact, one correctly rounded division -> bit-identical to the reference; float costs:
// direct <= 9-term sums instead of row-prefix differences (more accurate; tests/test_gpu_parity.py states the tolerance).
constexpr int CBR_TX = 4, CBR_MA = 4, CBR_RING = 12;

__device__ __forceinline__ int2 ldg_support(const short4 *p) { return __ldg(reinterpret_cast<const int2 *>(p)); }

// DT: the number of disparities when it is one of the usual ones (column strides become immediates), 0 = run time.
template <int DT>
__global__ void __launch_bounds__(256, 3) cbca_aggregate_reg_kernel(const float *__restrict__ cv_in, float *__restrict__ cv_out, int H, int W,
                                                                 int D_rt, int dmin, int off, const short4 *__restrict__ crossL,
                                                                 const short4 *__restrict__ crossR, int band_rows) {
    constexpr int TX = CBR_TX, MA = CBR_MA, RING = CBR_RING, CW = TX + 2 * MA;
    const int D = DT > 0 ? DT : D_rt;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x, wy = threadIdx.y, nw = blockDim.y;
    // thread-private rings, addressed by 32-bit shared addresses: PE [RING][TX][32] float, PN [RING][TX][32] uint16
    const uint32_t pe_base = smem_u32(smem_raw) + (uint32_t)((wy * RING * TX * 32 + lane) * 4);
    const uint32_t pn_base = smem_u32(smem_raw) + (uint32_t)(nw * RING * TX * 32 * 4) + (uint32_t)((wy * RING * TX * 32 + lane) * 2);
    auto pe_ld = [&](int slot, int c) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(pe_base + (uint32_t)((slot * TX + c) * 128))); return v; };
    auto pn_ld = [&](int slot, int c) { unsigned short v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(pn_base + (uint32_t)((slot * TX + c) * 64))); return (unsigned)v; };
    auto pe_st = [&](int slot, int c, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(pe_base + (uint32_t)((slot * TX + c) * 128)), "f"(v) : "memory"); };
    auto pn_st = [&](int slot, int c, unsigned v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(pn_base + (uint32_t)((slot * TX + c) * 64)), "h"((unsigned short)v) : "memory"); };
    const int Hi = H - 2 * off, Wi = W - 2 * off;
    const int k = (blockIdx.y * nw + wy) * 32 + lane;
    const int x0 = blockIdx.x * TX;
    const bool kact = k < D;
    // Row bands (blockIdx.z): outputs for rows [R0, R1).  The march starts MA + 1 rows above the band with empty prefixes
    // -- only differences of prefixes are used, and no row above R0 - MA - 1 enters one -- and ends MA rows below it.
    const int R0 = blockIdx.z * band_rows, R1 = min(Hi, R0 + band_rows);
    const int i_start = max(0, R0 - MA - 1), i_hend = min(Hi, R1 + MA);
    const int d = dmin + k;
    const size_t row_elems = (size_t)W * D;

    // column ranges as a few integers (cheaper to keep than one predicate per column): c is active for c < c_act, its
    // right-image column exists for v_lo <= c < v_hi; staged column j lies inside the interior view for j_lo <= j < j_hi
    const int c_act = kact ? min(TX, Wi - x0) : 0;
    const int v_lo = max(0, -(x0 + d)), v_hi = min(c_act, Wi - (x0 + d));
    const int j_lo = kact ? max(0, MA - x0) : CW, j_hi = kact ? min(CW, Wi - x0 + MA) : 0;
    // FAST (warp-uniform): every lane has a disparity, every staged column lies inside the view and every column of every
    // lane has its right-image partner -- true for all strips but those near the left / right image border; the march is
    // instantiated twice so that this common case carries no per-load predicate at all
    const bool fast = __all_sync(0xffffffffu, kact && c_act == TX && v_lo == 0 && v_hi == TX && j_lo == 0 && j_hi == CW);
#define act(c) (FAST || (c) < c_act)
#define vcol(c) (FAST || ((c) >= v_lo && (c) < v_hi))
#define cok(j) (FAST || ((j) >= j_lo && (j) < j_hi))
    const float *crow0 = cv_in + ((size_t)off * W + (size_t)(x0 - MA + off)) * D + k;      // column j of row 0: crow0 + j * D (only dereferenced when cok[j])
    const short4 *xl0 = crossL + x0;
    const short4 *xr0 = crossR + x0 + d;
    float *orow = cv_out + ((size_t)off * W + (size_t)(x0 + off)) * D + k;                 // column c of row yo = 0: orow + c * D


    float pe[TX];
    unsigned pn[TX], tbh[TX];
#pragma unroll
    for (int c = 0; c < TX; ++c) { pe[c] = 0.f; pn[c] = 0u; tbh[c] = 0u; }
    unsigned nanh = 0u;                           // centre-is-NaN bits, 4 per row, newest row in the low nibble
    int rm = 0;                                   // i mod RING: the ring slot row i is written to
    // Iteration i: (a) issue the loads of row i, (b) vertical stage of row yo = i - MA - 1 from the ring and the
    // histories -- it needs rows <= yo + MA = i - 1 only, so it runs while the loads are in flight --, (c) horizontal
    // stage of row i.  The kernel is bound by its instruction count (measured: 24 resident warps per SM at 80
    // registers beat 12 warps at 122 registers with every load prefetched a row ahead), hence no register prefetch.
    auto march = [&](auto fast_tag) {
    constexpr bool FAST = decltype(fast_tag)::value;
    for (int i = i_start; i < R1 + MA + 1; ++i) {
        float cc[CW];
        int2 xl[TX], xr[TX];
        if (i < i_hend) {
            const float *src = crow0 + (size_t)i * row_elems;
#pragma unroll
            for (int j = 0; j < CW; ++j) cc[j] = cok(j) ? __ldg(src + (size_t)j * D) : 0.f;
            const short4 *pl = xl0 + (size_t)i * Wi, *pr = xr0 + (size_t)i * Wi;
#pragma unroll
            for (int c = 0; c < TX; ++c) {
                xl[c] = vcol(c) ? ldg_support(pl + c) : make_int2(0, 0);
                xr[c] = vcol(c) ? ldg_support(pr + c) : make_int2(0, 0);
            }
        }
        // ---- vertical arm sums of row yo = i - MA - 1 -----------------------------------------------------------
        const int yo = i - MA - 1;
        if (yo >= R0) {
            float *o = orow + (size_t)yo * row_elems;
#pragma unroll
            for (int c = 0; c < TX; ++c) {
                // branch-free: a column without a right-image partner has t = b = 0 and empty prefixes pushed for it,
                // so the same differences give e = 0, n = 1
                const unsigned tb = tbh[c] >> 24;                           // pushed MA + 1 rows ago, shifted MA times since
                const int t = (int)(tb >> 3) & 7, b = (int)(tb & 7u);
                int s1 = rm - (MA + 1 - b);                                 // slot of row yo + b
                s1 += (s1 < 0) ? RING : 0;
                int s0 = rm - (MA + 2 + t);                                 // slot of row yo - t - 1
                s0 += (s0 < 0) ? RING : 0;
                const bool has0 = (yo - t - 1 >= i_start);                  // rows above the march never entered a prefix
                const float e1 = pe_ld(s1, c), e0r = pe_ld(s0, c);
                const unsigned n1 = pn_ld(s1, c), n0r = pn_ld(s0, c);
                const float e = e1 - (has0 ? e0r : 0.f);
                const unsigned n = ((n1 - (has0 ? n0r : 0u)) & 0xFFFFu) + (unsigned)(t + b + 1);
                const bool cnan = ((nanh >> (4 * MA + c)) & 1u) != 0u;
                if (act(c)) o[(size_t)c * D] = cnan ? nan_f() : e / (float)n;   // (0*c + E) / N
            }
        }
        // ---- horizontal arm sums of row i, pushed into the ring and the histories ---------------------------------
        unsigned nan_i = 0u;
        if (i < i_hend) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const bool nn = (cc[j] != cc[j]);
                if (j >= MA && j < MA + TX) nan_i |= nn ? (1u << (j - MA)) : 0u;
                cc[j] = nn ? 0.f : cc[j];                                   // step 1 does not propagate NaN
            }
#pragma unroll
            for (int c = 0; c < TX; ++c) {
                float eh = 0.f;
                unsigned nh = 0u, tb = 0u;
                if (vcol(c)) {
                    // short4 (left, right, up, bottom) as two words of two 16-bit arms: one SIMD minimum each, clamped to MA
                    const unsigned lr = __vminu2(__vminu2((unsigned)xl[c].x, (unsigned)xr[c].x), (unsigned)MA * 0x10001u);
                    const unsigned ud = __vminu2(__vminu2((unsigned)xl[c].y, (unsigned)xr[c].y), (unsigned)MA * 0x10001u);
                    const int l = (int)(lr & 0xFFFFu), r = (int)(lr >> 16);
                    eh = cc[MA + c];
#pragma unroll
                    for (int q = 1; q <= MA; ++q) {
                        if (q <= l) eh += cc[MA + c - q];
                        if (q <= r) eh += cc[MA + c + q];
                    }
                    nh = (unsigned)(l + r);
                    tb = ((ud & 0xFFFFu) << 3) | (ud >> 16);
                }
                pe[c] = pe[c] + eh;                                         // the reference's step-3 column prefix
                pn[c] += nh;
                pe_st(rm, c, pe[c]);
                pn_st(rm, c, pn[c]);
                tbh[c] = (tbh[c] << 6) | tb;
            }
        } else {
#pragma unroll
            for (int c = 0; c < TX; ++c) tbh[c] <<= 6;
        }
        nanh = (nanh << 4) | nan_i;
        rm = (rm + 1 == RING) ? 0 : rm + 1;
    }
    };
    if (fast) march(std::true_type{});
    else march(std::false_type{});
#undef act
#undef vcol
#undef cok
}

static int launch_cbca_reg(const float *in, float *out, int H, int W, int D, int dmin, int off, const int16_t *cl, const int16_t *cr,
                           cudaStream_t s) {
    const int kg = ceil_div(D, 32);
    const int nw = kg < 8 ? kg : 8;                                          // warps (disparity groups) per CTA
    const size_t smem = (size_t)nw * CBR_RING * CBR_TX * 32 * (sizeof(float) + sizeof(unsigned short));
    void (*kern)(const float *, float *, int, int, int, int, int, const short4 *, const short4 *, int) =
        D == 64 ? cbca_aggregate_reg_kernel<64> : D == 128 ? cbca_aggregate_reg_kernel<128> : D == 192 ? cbca_aggregate_reg_kernel<192>
        : D == 256 ? cbca_aggregate_reg_kernel<256> : cbca_aggregate_reg_kernel<0>;
    PB200_CUDA(cudaFuncSetAttribute((const void *)kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // row bands: enough CTAs for several waves (a CTA marches its whole band), at least 128 rows each so that the
    // 2 * MA + 1 extra rows of a band stay a few per cent
    const int Hi = H - 2 * off, strips = ceil_div(W - 2 * off, CBR_TX), kb = ceil_div(kg, nw);
    int bands = ceil_div(8L * sm_count(), (long)strips * kb);
    if (bands > Hi / 128) bands = Hi / 128;
    if (bands < 1) bands = 1;
    if (option(OPT_CBCA_BANDS) > 0) bands = option(OPT_CBCA_BANDS);
    if (bands < 1) bands = 1;
    const int band_rows = ceil_div(Hi, bands);
    dim3 block(32, nw), grid(strips, kb, ceil_div(Hi, band_rows));
    kern<<<grid, block, smem, s>>>(in, out, H, W, D, dmin, off, (const short4 *)cl, (const short4 *)cr, band_rows);
    PB200_LAUNCH_CHECK("cbca_aggregate_reg_kernel");
    note_path(STAGE_CBCA, PATH_CBCA_REG, (D == 64 || D == 128 || D == 192 || D == 256) ? D : 0);
    return PB200_OK;
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
static int launch_cbca_pipe(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                            const int16_t *cr, cudaStream_t s) {
    constexpr int RING = 4 * MA, NS = 16;
    const size_t smem = (size_t)NS * (CBCA_TX + 2 * MA) * 32 * 4 + (size_t)(CBCA_TX + 2 * MA + 1) * 32 * 4 + 2 * (size_t)RING * CBCA_TX * 32 * 4 +
                        (size_t)NS * (CBCA_TX + CBCA_TX + 31) * 8 + (size_t)NS * CBCA_TX * 32;
    PB200_CUDA(cudaFuncSetAttribute(cbca_aggregate_pipe_kernel<MA, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, CBCA_TX), grid(ceil_div(W - 2 * off, CBCA_TX), ceil_div(D, 32));
    cbca_aggregate_pipe_kernel<MA, RAW><<<grid, block, smem, s>>>(in, out, out_n, H, W, D, dmin, off, (const short4 *)cl,
                                                                   (const short4 *)cr);
    PB200_LAUNCH_CHECK("cbca_aggregate_pipe_kernel");
    note_path(STAGE_CBCA, PATH_CBCA_PIPE, MA);
    return PB200_OK;
}

template <int MA, bool RAW>
static int launch_cbca(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                       const int16_t *cr, cudaStream_t s) {
    if constexpr (MA <= 8) return launch_cbca_pipe<MA, RAW>(in, out, out_n, H, W, D, dmin, off, cl, cr, s);
    constexpr int RING = 4 * MA;
    const size_t smem = (size_t)(CBCA_TX + 2 * MA) * 32 * 4 + 2 * (size_t)RING * CBCA_TX * 32 * 4;
    PB200_CUDA(cudaFuncSetAttribute(cbca_aggregate_kernel<MA, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 block(32, CBCA_TX), grid(ceil_div(W - 2 * off, CBCA_TX), ceil_div(D, 32));
    cbca_aggregate_kernel<MA, RAW><<<grid, block, smem, s>>>(in, out, out_n, H, W, D, dmin, off, (const short4 *)cl,
                                                              (const short4 *)cr);
    PB200_LAUNCH_CHECK("cbca_aggregate_kernel");
    note_path(STAGE_CBCA, PATH_CBCA_STAGED, MA);
    return PB200_OK;
}

// shared by pb200_cbca_aggregate (out_n == NULL) and the reference-style per-slice host entry point
int cbca_dispatch(const float *in, float *out, float *out_n, int H, int W, int D, int dmin, int off, const int16_t *cl,
                  const int16_t *cr, int len_arms, cudaStream_t s) {
    const int ma = len_arms - 1;
    // the register kernel takes the common case (cbca_distance <= 5, normalised output); the option "cbca.pipe" = 1 keeps the staged one
    if (ma <= CBR_MA && out_n == nullptr && option(OPT_CBCA_PIPE) <= 0) return launch_cbca_reg(in, out, H, W, D, dmin, off, cl, cr, s);
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
