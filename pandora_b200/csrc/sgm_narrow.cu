// sgm_narrow.cu -- exact packed-integer fast path of the 8-path SGM stage.
//
// Same recurrence and accumulation as sgm.cu / oracle/pandora_oracle.c::pbo_sgm, for the common case where
// every cost, P1, P2 and the invalid value are small non-negative integers (Census / SAD on integer images):
// then every L_r and every partial sum is an integer below 2^16, float32 arithmetic on them is exact, and the
// same numbers can be computed with 16-bit integer SIMD (two disparities per 32-bit register; VIMNMX.U16x2 /
// VIADDMNMX.U16x2, the DPX dynamic-programming instructions) and kept in 16-bit storage between the passes.
// The result is bit-identical to the float path; only the traffic and the instruction count change:
//
// Single-call runs (one GPU) take the two 4-direction WAVEFRONT passes further down (sgm_wave_kernel: 14D bytes per
// pixel, two launches).  Row-tiled multi-GPU runs, whose vertical groups must be split around the halo exchange, and
// images wider than 28 columns per SM keep the four-launch schedule:
//   pass E   read C float32 (4D B/pixel), check + pack it to C16, write C16 and P16 = L_E        (2D + 2D)
//   pass W   read C16, read P16, P16 += L_W, write P16                                           (2D + 2D + 2D)
//   sweep S  read C16, read P16, P16 += L_S + L_SE + L_SW, write P16                              (6D)
//   sweep N  read C16, read P16, total = P16 + L_N + L_NE + L_NW -> float32 S (+ NaN, WTA)        (4D + 4D)
//
// = 28D bytes per pixel instead of 44D (float sweeps) or 92D (one launch per direction).  C16 and P16 live INSIDE
// the caller's float32 output buffer (2 + 2 bytes per cell), in a private word order: word lane*NR + j of a
// pixel holds the disparities (NR*lane + j) in its low and (D/2 + NR*lane + j) in its high half, so a lane's
// registers are one aligned vector and both halves of a register have their d-1 / d+1 neighbours in the
// adjacent register.  The last sweep converts in place: every lane overwrites exactly the bytes it loaded.
//
// The data condition (integer costs in [0, 8191 - P2]) is verified by pass E on every cell; a violation raises a
// device flag, the remaining narrow kernels return at once and the float kernels, enqueued behind them and gated
// on the same flag, redo the whole stage.  No host synchronisation is involved.
#include <type_traits>

#include "sgm_packed.cuh"

namespace pb200 {

namespace {

// ------------------------------------------------------------------------------------------------
// horizontal passes: one warp per row
// ------------------------------------------------------------------------------------------------
template <int NR, int CB, bool FIRST>
__global__ void __launch_bounds__(128) sgm_narrow_h_kernel(const NarrowParams p) {
    if (!FIRST && *p.flag != 0) return;
    const int lane = threadIdx.x & 31;
    const long path = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (path >= p.H) return;
    const int W = p.W, D = p.D;
    const int dx = FIRST ? 1 : -1;
    const int x = FIRST ? 0 : W - 1;
    const size_t pix0 = (size_t)path * W + x;
    const long step = (long)dx * D;
    uint32_t *b_ptr = p.buf + pix0 * D;                     // this pixel's D-word region
    const float *c_ptr = p.cv + pix0 * D + lane * NR;       // FIRST: low-half costs; high-half costs D/2 further
    const int poff = p16_off<CB>(D) + lane * NR;            // P16 words of this lane
    const int qoff = (CB == 2) ? poff : p8_off(D) + lane * (NR / 2);   // what E hands to W: P16, or P8 in the byte tier
    uint32_t Lp[NR];
    bool bad = false;
    // prefetch registers for the next pixel
    float fa[NR], fb[NR];
    constexpr int RW = NR * CB / 2;                         // raw words per lane (costs, and the E -> W hand-over)
    uint32_t cn[RW], pn[RW];
    if (FIRST) { ld_floats<NR>(c_ptr, fa); ld_floats<NR>(c_ptr + D / 2, fb); }
    else { ld_cost_raw<NR, CB>(b_ptr, lane, cn); ld_words<RW>(b_ptr + qoff, pn); }
    for (int i = 0; i < W; ++i) {
        uint32_t c[NR], ps[NR], cc[NR];
        if (FIRST) {
#pragma unroll
            for (int j = 0; j < NR; ++j)
                c[j] = encode_cost<CB>(fa[j], p.inv, p.cost_ok_max, bad) | (encode_cost<CB>(fb[j], p.inv, p.cost_ok_max, bad) << 16);
        } else {
            unpack_cost<NR, CB>(cn, c);
            unpack_cost<NR, CB>(pn, ps);                    // the hand-over uses the same packing as the costs
        }
        if (i + 1 < W) {
            if (FIRST) { ld_floats<NR>(c_ptr + step, fa); ld_floats<NR>(c_ptr + step + D / 2, fb); }
            else { ld_cost_raw<NR, CB>(b_ptr + step, lane, cn); ld_words<RW>(b_ptr + step + qoff, pn); }
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) cc[j] = c[j] & Tier<CB>::VALUES;
        uint32_t L[NR];
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < NR; ++j) L[j] = cc[j];
        } else {
            nstep<NR>(cc, Lp, L, lane, p.p1p1, p.p2p2);
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            Lp[j] = L[j];
            ps[j] = FIRST ? L[j] : ps[j] + L[j];
        }
        if (FIRST) {
            st_cost<NR, CB>(b_ptr, lane, c);
            if constexpr (CB == 2) st_words<NR>(b_ptr + qoff, ps);
            else st_cost<NR, 1>(b_ptr + p8_off(D), lane, ps);               // L_E <= 255: one byte each
        } else {
            st_words<NR>(b_ptr + poff, ps);
        }
        b_ptr += step;
        c_ptr += step;
    }
    if (FIRST && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.flag, 1);
}

// ------------------------------------------------------------------------------------------------
// vertical sweeps: same strip / exchange-warp scheme as sgm_vsweep_kernel (sgm.cu), packed states
// ------------------------------------------------------------------------------------------------
// A path start (L = C) is the same as a step from a FLAT previous state (all disparities equal: m = Lp[d], so
// t - m = 0).  The state buffers therefore start as zeros and image-border halo columns simply stay zero: the row
// loop has no "first row" / "no predecessor" cases.  Columns right of the image (last strip) run on zero costs,
// which keeps their leftward diagonal state flat, and have their loads / stores predicated off.
template <int NR, int CB, bool FINAL, bool WTA>
__global__ void __launch_bounds__(512, 1) sgm_narrow_vsweep_kernel(const NarrowParams p) {
    if (*p.flag != 0) return;
    extern __shared__ __align__(16) uint32_t nsweep_smem[];
    constexpr int VS = NR * 32;                          // words per packed state vector
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x >> 5) - 1;
    const int K = nwarp * 2;
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int H = p.H, W = p.W, D = p.D, dy = p.dy;
    // shared: st[buf][diag][K + 2][VS] words; diag 0: dx = +1, diag 1: dx = -1; column index = strip column + 1
    const uint32_t BUFB = (uint32_t)(2 * (K + 2) * VS) * 4u;            // bytes per buffer
    const uint32_t DIAGB = (uint32_t)((K + 2) * VS) * 4u;               // bytes per diagonal plane
    const uint32_t sbase = smem_u32(nsweep_smem) + (uint32_t)(lane * NR) * 4u;
    for (int i = threadIdx.x; i < 2 * 2 * (K + 2) * VS + 4 * nwarp * 2 * 32 * (NR * CB / 2 + NR); i += blockDim.x) nsweep_smem[i] = 0u;
    __syncthreads();
    const bool has_left = strip > 0 && !(p.debug & 1), has_right = strip + 1 < nstrips && !(p.debug & 1);
    const size_t hplane = (size_t)W * VS;                 // words per halo plane
    if (p.halo_in != nullptr) {
        // row-tiled run: the previous row lives in the neighbouring tile; its diagonal states go where row 0 will
        // look for them (buffer 1 is "previous" for row 0), the vertical ones into Lv below
        if (warp < nwarp) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int cl = c == 0 ? warp : K - 1 - warp, x = strip * K + cl;
                if (x < W) {
#pragma unroll
                    for (int dg = 0; dg < 2; ++dg) {
                        uint32_t v[NR];
                        ld_words<NR>(p.halo_in + (size_t)(dg + 1) * hplane + (size_t)x * VS + lane * NR, v);
                        sts_words<NR>(sbase + (uint32_t)dg * DIAGB + (uint32_t)((cl + 1) * VS) * 4u + BUFB, v);
                    }
                }
            }
        } else {
            const int xl = strip * K - 1, xr = strip * K + K;
            if (xl >= 0) {
                uint32_t v[NR];
                ld_words<NR>(p.halo_in + 1 * hplane + (size_t)xl * VS + lane * NR, v);
                sts_words<NR>(sbase + BUFB, v);
            }
            if (xr < W) {
                uint32_t v[NR];
                ld_words<NR>(p.halo_in + 2 * hplane + (size_t)xr * VS + lane * NR, v);
                sts_words<NR>(sbase + DIAGB + (uint32_t)((K + 1) * VS) * 4u + BUFB, v);
            }
        }
        __syncthreads();
    }

    if (warp == nwarp) {                                  // exchange warp: neighbours' border states of row i -> halo columns
        const unsigned long long *ring_l = p.ring + (size_t)((strip - 1) * 2 + 1) * 2 * VS;
        const unsigned long long *ring_r = p.ring + (size_t)((strip + 1) * 2 + 0) * 2 * VS;
        const uint32_t halo_l = sbase, halo_r = sbase + DIAGB + (uint32_t)((K + 1) * VS) * 4u;
        for (int i = 0; i + 1 < H; ++i) {
            const int cur = i & 1;
            const uint32_t tag = (uint32_t)(i + 1);
            if (has_left) {
                uint32_t v[NR];
                ll_recv_u32<NR>(ring_l + cur * VS, lane, tag, v);
                sts_words<NR>(halo_l + cur * BUFB, v);
            }
            if (has_right) {
                uint32_t v[NR];
                ll_recv_u32<NR>(ring_r + cur * VS, lane, tag, v);
                sts_words<NR>(halo_r + cur * BUFB, v);
            }
            __syncthreads();
        }
        __syncthreads();
        return;
    }

    // ---- compute warps: strip columns `warp` and K-1-warp ------------------------------------------------
    const int col[2] = {warp, K - 1 - warp};
    const int xs[2] = {strip * K + col[0], strip * K + col[1]};
    const bool valid[2] = {xs[0] < W, xs[1] < W};
    const bool send_l = (warp == 0) && has_left, send_r = (warp == 0) && has_right;
    unsigned long long *ring_sl = p.ring + (size_t)(strip * 2 + 0) * 2 * VS, *ring_sr = p.ring + (size_t)(strip * 2 + 1) * 2 * VS;
    // byte offsets inside a state buffer: own[c][diag] where this column stores, pred[c][diag] where its predecessor did
    uint32_t own[2][2], pred[2][2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        own[c][0] = sbase + (uint32_t)((col[c] + 1) * VS) * 4u;
        own[c][1] = own[c][0] + DIAGB;
        pred[c][0] = own[c][0] - (uint32_t)VS * 4u;        // dx = +1: column to the left
        pred[c][1] = own[c][1] + (uint32_t)VS * 4u;        // dx = -1: column to the right
    }
    const long row_stride = (long)dy * W * D;              // words
    const int y0 = dy > 0 ? 0 : H - 1;
    uint32_t *gp[2];                                       // pixel regions of the current row
#pragma unroll
    for (int c = 0; c < 2; ++c) gp[c] = p.buf + ((size_t)y0 * W + (valid[c] ? xs[c] : 0)) * D;
    const int poff = p16_off<CB>(D) + lane * NR;

    constexpr int RW = NR * CB / 2;                        // raw cost words per lane
    // Input staging: the C / P16 words of the next PFD rows travel global -> shared with cp.async (LDGSTS), one
    // private slot per (stage, warp, column slot, lane): no register prefetch sets, no barrier (a lane only reads
    // what it copied itself), and deep enough to cover the HBM latency of a row step that takes ~1.3 us.
    constexpr int NSTG = 4, PFD = NSTG - 1;
    uint32_t *stg_c = nsweep_smem + 2 * 2 * (K + 2) * VS;                       // [NSTG][nwarp][2][32][RW]
    uint32_t *stg_p = stg_c + (size_t)NSTG * nwarp * 2 * 32 * RW;               // [NSTG][nwarp][2][32][NR]
    const uint32_t stgc_base = smem_u32(stg_c) + (uint32_t)(((warp * 2) * 32 + lane) * RW) * 4u;
    const uint32_t stgp_base = smem_u32(stg_p) + (uint32_t)(((warp * 2) * 32 + lane) * NR) * 4u;
    const uint32_t stgc_stage = (uint32_t)(nwarp * 2 * 32 * RW) * 4u, stgp_stage = (uint32_t)(nwarp * 2 * 32 * NR) * 4u;
    uint32_t Lv[2][NR];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int j = 0; j < NR; ++j) Lv[c][j] = 0u;
        if (valid[c] && p.halo_in != nullptr) ld_words<NR>(p.halo_in + (size_t)xs[c] * VS + lane * NR, Lv[c]);
    }
    auto stage_in = [&](int r) {                           // enqueue row r (travel order) of both column slots
        const uint32_t sg = (uint32_t)(r & (NSTG - 1));
#pragma unroll
        for (int c = 0; c < 2; ++c)
            if (valid[c]) {
                const uint32_t *src = gp[c] + (long)r * row_stride;
                cp_async_words<RW>(stgc_base + sg * stgc_stage + (uint32_t)(c * 32 * RW) * 4u, src + lane * RW);
                cp_async_words<NR>(stgp_base + sg * stgp_stage + (uint32_t)(c * 32 * NR) * 4u, src + poff);
            }
    };
    for (int r = 0; r < PFD; ++r) {
        if (r < H) stage_in(r);
        cp_async_commit();
    }

    // one row of the sweep
    auto row = [&](const int i) {
        const bool last = (i == H - 1);
        const uint32_t curb = (uint32_t)(i & 1) * BUFB, prvb = BUFB - curb;
        if (i + PFD < H) stage_in(i + PFD);
        cp_async_commit();
        cp_async_wait<PFD>();                              // this lane's copies of row i have landed
        uint32_t c16[2][NR], p16[2][NR], cc[2][NR];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t craw[RW];
            const uint32_t sg = (uint32_t)(i & (NSTG - 1));
            lds_words<RW>(stgc_base + sg * stgc_stage + (uint32_t)(c * 32 * RW) * 4u, craw);
            lds_words<NR>(stgp_base + sg * stgp_stage + (uint32_t)(c * 32 * NR) * 4u, p16[c]);
            unpack_cost<NR, CB>(craw, c16[c]);
#pragma unroll
            for (int j = 0; j < NR; ++j) cc[c][j] = c16[c][j] & Tier<CB>::VALUES;
        }
        uint32_t Lb[2][NR], Lp[NR];
        // outgoing border diagonals first: column slot 0 (strip column `warp`) dx = -1, slot 1 dx = +1
        lds_words<NR>(pred[0][1] + prvb, Lp);
        nstep<NR>(cc[0], Lp, Lb[0], lane, p.p1p1, p.p2p2);
        sts_words<NR>(own[0][1] + curb, Lb[0]);
        if (send_l && !last) ll_send_u32<NR>(ring_sl + (i & 1) * VS, lane, (uint32_t)(i + 1), Lb[0]);
        lds_words<NR>(pred[1][0] + prvb, Lp);
        nstep<NR>(cc[1], Lp, Lb[1], lane, p.p1p1, p.p2p2);
        sts_words<NR>(own[1][0] + curb, Lb[1]);
        if (send_r && !last) ll_send_u32<NR>(ring_sr + (i & 1) * VS, lane, (uint32_t)(i + 1), Lb[1]);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t L0[NR], Lo[NR], tot[NR];
            nstep<NR>(cc[c], Lv[c], L0, lane, p.p1p1, p.p2p2);
#pragma unroll
            for (int j = 0; j < NR; ++j) Lv[c][j] = L0[j];
            const int od = (c == 0) ? 0 : 1;               // the diagonal not done above: slot 0 -> dx = +1, slot 1 -> dx = -1
            lds_words<NR>(pred[c][od] + prvb, Lp);
            nstep<NR>(cc[c], Lp, Lo, lane, p.p1p1, p.p2p2);
            sts_words<NR>(own[c][od] + curb, Lo);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] = p16[c][j] + L0[j] + Lo[j] + Lb[c][j];
            if (last && p.halo_out != nullptr && valid[c]) {
                uint32_t *ho = p.halo_out + (size_t)xs[c] * VS + lane * NR;
                st_words<NR>(ho, L0);
                st_words<NR>(ho + hplane, c == 0 ? Lo : Lb[1]);             // dx = +1
                st_words<NR>(ho + 2 * hplane, c == 0 ? Lb[0] : Lo);         // dx = -1
            }
            if (valid[c]) {
                uint32_t *gpix = gp[c] + (long)i * row_stride;
                if (!FINAL) {
                    st_words<NR>(gpix + poff, tot);
                } else {
                    // total -> float32 (exact), overcounting, NaN restore, in place: this lane overwrites the bytes it loaded
                    float fa[NR], fb[NR];
                    uint32_t best = 0xFFFFFFFFu;
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        uint32_t t = tot[j];
                        if (p.overcounting) t = t - 7u * cc[c][j];   // S >= 8 C in every half: no borrow
                        const uint32_t lo = t & 0xFFFFu, hi = t >> 16;
                        const bool nlo = (c16[c][j] & Tier<CB>::FLAG1) != 0, nhi = (c16[c][j] & (Tier<CB>::FLAG1 << 16)) != 0;
                        fa[j] = nlo ? nan_f() : small_int_to_float(lo);
                        fb[j] = nhi ? nan_f() : small_int_to_float(hi);
                        if (WTA) {
                            const uint32_t ka = (lo << 16) | (uint32_t)(lane * NR + j);
                            const uint32_t kb = (hi << 16) | (uint32_t)(D / 2 + lane * NR + j);
                            best = min(best, nlo ? 0xFFFFFFFFu : ka);
                            best = min(best, nhi ? 0xFFFFFFFFu : kb);
                        }
                    }
                    // in place: the stores depend on registers the whole warp's loads of this pixel have filled
                    float *o = reinterpret_cast<float *>(gpix) + lane * NR;
                    st_floats<NR>(o, fa);
                    st_floats<NR>(o + D / 2, fb);
                    if (WTA) {
                        best = __reduce_min_sync(0xffffffffu, best);
                        if (lane == 0) {
                            const size_t pix = (size_t)(y0 + i * dy) * W + xs[c];
                            const bool none = (best == 0xFFFFFFFFu);
                            p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(best & 0xFFFFu));
                            if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                        }
                    }
                }
            }
        }
        __syncthreads();
    };
#pragma unroll 1
    for (int i = 0; i < H; ++i) row(i);
}

// ------------------------------------------------------------------------------------------------
// wavefront sweeps: FOUR directions per pass, two passes for the whole stage
// ------------------------------------------------------------------------------------------------
// A pixel's E, SE, S and SW predecessors are (y, x-1), (y-1, x-1), (y-1, x), (y-1, x+1): all of them are finished
// when the image is walked in the order t = 2y + x.  sgm_wave_kernel runs that wavefront as DATAFLOW instead of a
// row barrier: a warp owns two adjacent columns and walks them downwards; what it needs from its neighbours travels
// through small shared-memory mailboxes guarded by progress counters (the left warp's E / SE states of its right
// column, the right warp's SW state of its left column), and across strips through the L2-resident flag-in-data
// ring, served by two relay warps per CTA so that the compute warps only ever see mailboxes.  The vertical state, the pair's inner E / SE / SW hand-overs and the partial sums never leave registers.
// Pass 1 (top-down: E, SE, S, SW) reads the float32 costs, verifies + packs them and writes C (8 or 16 bit) and the
// 16-bit partial sum; pass 2 runs the same code on the image flipped in both axes (bottom-up: W, NW, N, NE), adds
// its four directions and emits float32 S (+ NaN restore, overcounting, WTA) in place:
//      pass 1   4D read + D (or 2D) + 2D written        pass 2   D (or 2D) + 2D read + 4D written
// = 14D bytes per pixel (16D with 16-bit costs) instead of 22D / 28D, in two launches instead of four.
//
// Mailbox discipline (i = row in travel order, w = warp, A / B = its left / right column; every mailbox has four slots):
//   phase 1 (no E needed)   wait fs[w+1] >= i: SW_in = sw[(i-1)&3][w+1];  SW_B(i);  SW_A(i+1) from it, ONE ROW AHEAD
//                           -> mailbox sw[(i+1)&3][w], counter fs[w] = i+2;  S_A, S_B;  SE_B(i) from the pair's own
//                           SE_A(i-1) -> mailbox se[i&3][w]
//   phase 2 (the E chain)   wait fe[w-1] >= i+1: E_in = e[i&3][w-1];  E_A, E_B -> mailbox e[i&3][w], counter fe[w] = i+1
//   phase 3                 SE_in = se[(i-1)&3][w-1] (visible since that E flag);  SE_A(i)
// A slot is overwritten four rows later, and a writer that reaches row i+4 has -- through its own waits of rows
// i+3 / i+4 -- proof that its reader finished row i+1, the last row that looks at the slot.  Column 0 / nc+1 of every
// mailbox belong to the relay warps (the neighbouring strips); at an image border their counters start at "infinity"
// and the slots stay zero: a flat state, i.e. a path start.
template <int NR, int CB, bool FINAL, bool WTA, bool CENSUS = false, bool BATCH = false>
__global__ void __launch_bounds__(512, 1) sgm_wave_kernel(const NarrowParams p) {
    static_assert(!(CENSUS && FINAL), "the Census source only exists for the first pass");
    if (FINAL && *p.flag != 0) return;
    extern __shared__ __align__(16) uint32_t wave_smem[];
    constexpr int VS = NR * 32;                          // words per packed state vector
    constexpr int RW = NR * CB / 2;                      // raw cost words per lane
    constexpr int SIN = FINAL ? (RW + NR) : 2 * NR;      // staged input words per lane and pixel
    // CENSUS: a staged row of a warp is D + 1 right descriptors (columns xA + dmin ... xB + dmax), the two left ones and
    // padding: CW words.  Pixel A reads row i + 1 while pixel B still reads row i and PFD + 1 rows are in flight: 8 slots.
    constexpr int CW = 2 * VS + 8;
    constexpr int PFD = 3, NSTG = CENSUS ? 8 : 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nc = (blockDim.x >> 5) - 2;                // compute warps; warps nc / nc + 1 relay to the left / right strip
    const int NV = nc + 2;                               // mailbox columns: 0 = left strip, 1 .. nc = compute warps, nc + 1 = right strip
    const int K = nc * 2;
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int H = p.H, W = p.W, D = p.D;
    const bool relaxed = (p.debug & 2) != 0, nowait = (p.debug & 4) != 0;
    // p.debug (PB200_SGM_DEBUG, timing experiments only, wrong results): bit 0 = no strip exchange, bit 2 = no mailbox waits
    const bool has_left = strip > 0 && !(p.debug & 1), has_right = strip + 1 < nstrips && !(p.debug & 1);
    // shared: e[4][NV][VS] | se[4][NV][VS] | sw[4][NV][VS] | fe[32] fs[32] | staging
    const int state_words = 12 * NV * VS + 64 + 256;    // + the hand-over events: FE[16][4] | FS[16][4] mbarriers
    // mailboxes, counters AND the staging ring start as zeros (columns right of the image are never staged and must
    // read as zero costs, which also pass the data check; a strip without a neighbour reads flat zero states)
    const int stage_words = CENSUS ? NSTG * nc * CW : NSTG * nc * 2 * 32 * SIN;
    for (int i = threadIdx.x; i < state_words + stage_words; i += blockDim.x) wave_smem[i] = 0u;
    __syncthreads();
    // Hand-over events are mbarrier phases, not spin flags: a warp that is early sleeps in hardware (mbarrier.try_wait)
    // instead of looping LDS / ISETP / BRA through the issue slots of the working warps -- those loops were 44 % of all
    // issued instructions of this kernel (profiles/r1_ncu_fused_pass1.txt).  Event r of a stream completes phase r >> 2 of
    // its barrier r & 3; the four-slot discipline below keeps a producer at most three events ahead of a waiter, i.e.
    // never a whole phase ahead on one barrier.  An image border is a neighbour that is never waited for.
    if (threadIdx.x < 128)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(wave_smem + 12 * NV * VS + 64) + threadIdx.x * 8u) : "memory");
    __syncthreads();
    const uint32_t lane_b = (uint32_t)(lane * NR) * 4u;
    const uint32_t e_base = smem_u32(wave_smem) + lane_b;
    const uint32_t se_base = e_base + (uint32_t)(4 * NV * VS) * 4u;
    const uint32_t sw_base = se_base + (uint32_t)(4 * NV * VS) * 4u;
    const uint32_t fe_base = smem_u32(wave_smem) + (uint32_t)(12 * NV * VS) * 4u + 256u, fs_base = fe_base + 512u;   // event streams: + 32 v
    const uint32_t SLOT = (uint32_t)(NV * VS) * 4u, VB = (uint32_t)VS * 4u;      // bytes per mailbox slot / per vector
    // ring, per strip boundary b (between strips b and b + 1): 12 vectors of VS 64-bit words: e[4] | se[4] | sw[4]
    unsigned long long *ring_l = p.ring + (size_t)(strip - 1) * 12 * VS;         // boundary on our left (used when has_left)
    unsigned long long *ring_r = p.ring + (size_t)strip * 12 * VS;               // boundary on our right

    // ---- relay warps: mailbox <-> L2 ring, so that no compute warp ever touches the ring -----------------------
    if (warp == nc) {                                     // left relay
        if (!has_left) return;
        for (int i = 0; i < H; ++i) {
            const uint32_t tag = (uint32_t)(i + 1);
            uint32_t v[NR];
            // outbound: SW_A(i) of the first compute warp (published one row early, see phase 1)
            if (!nowait) ev4_wait(fs_base + 1u * 32u, i);
            lds_words<NR>(sw_base + (uint32_t)(i & 3) * SLOT + 1u * VB, v);
            ll_send_u32<NR>(ring_l + (size_t)(8 + (i & 3)) * VS, lane, tag, v);
            // inbound: E(i), then SE(i), of the left strip's last column
            ll_recv_u32<NR>(ring_l + (size_t)(0 + (i & 3)) * VS, lane, tag, v);
            sts_words<NR>(e_base + (uint32_t)(i & 3) * SLOT, v);
            ev4_signal(fe_base, i, lane);
            ll_recv_u32<NR>(ring_l + (size_t)(4 + (i & 3)) * VS, lane, tag, v);
            sts_words<NR>(se_base + (uint32_t)(i & 3) * SLOT, v);                 // visible with the next E flag
        }
        return;
    }
    if (warp == nc + 1) {                                 // right relay
        if (!has_right) return;
        for (int i = 0; i < H; ++i) {
            const uint32_t tag = (uint32_t)(i + 1);
            uint32_t v[NR];
            // outbound: E_B(i), SE_B(i) of the last compute warp
            if (!nowait) ev4_wait(fe_base + (uint32_t)nc * 32u, i);
            lds_words<NR>(e_base + (uint32_t)(i & 3) * SLOT + (uint32_t)nc * VB, v);
            ll_send_u32<NR>(ring_r + (size_t)(0 + (i & 3)) * VS, lane, tag, v);
            lds_words<NR>(se_base + (uint32_t)(i & 3) * SLOT + (uint32_t)nc * VB, v);
            ll_send_u32<NR>(ring_r + (size_t)(4 + (i & 3)) * VS, lane, tag, v);
            // inbound: SW_A(i) of the right strip's first column
            ll_recv_u32<NR>(ring_r + (size_t)(8 + (i & 3)) * VS, lane, tag, v);
            sts_words<NR>(sw_base + (uint32_t)(i & 3) * SLOT + (uint32_t)(nc + 1) * VB, v);
            ev4_signal(fs_base + (uint32_t)(nc + 1) * 32u, i, lane);
        }
        return;
    }

    // ---- compute warps -------------------------------------------------------------------------------------------
    const uint32_t vme = (uint32_t)(warp + 1);            // this warp's mailbox column
    const bool wait_l = has_left || warp > 0, wait_r = has_right || warp + 1 < nc;
    uint32_t *stg = wave_smem + state_words;              // [NSTG][nc][2 pixels][32 * SIN]
    // a pixel's staging block holds two lane-major parts so that every lane's vectors stay naturally aligned:
    // pass 1: [32][NR] low-half floats | [32][NR] high-half floats;  pass 2: [32][RW] cost words | [32][NR] partial sums
    constexpr int P0 = FINAL ? RW : NR;
    const uint32_t stg_pix = (uint32_t)(32 * SIN) * 4u, stg_stage = (uint32_t)(nc * 2) * stg_pix;
    const uint32_t stg_base = smem_u32(stg) + (uint32_t)(warp * 2) * stg_pix;
    const uint32_t off0 = (uint32_t)(lane * P0) * 4u, off1 = (uint32_t)(32 * P0 + lane * NR) * 4u;

    // logical columns of this warp (travel frame: pass 2 sees the image flipped in both axes)
    const int xl[2] = {strip * K + 2 * warp, strip * K + 2 * warp + 1};
    const bool valid[2] = {xl[0] < W, xl[1] < W};
    const int y0 = FINAL ? H - 1 : 0;
    const long row_stride = (FINAL ? -1L : 1L) * W * D;                          // words (== floats)
    size_t pix0[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) pix0[c] = ((size_t)y0 * W + (valid[c] ? (FINAL ? W - 1 - xl[c] : xl[c]) : 0)) * D;
    const int poff = p16_off<CB>(D) + lane * NR;
    const size_t pixi0[2] = {pix0[0] / D, pix0[1] / D};                           // pixel indices of the first row (WTA outputs)

    // Pixel A (the left column) is staged, unpacked and -- for its SW direction -- computed ONE ROW AHEAD of pixel B:
    // SW_A(i + 1) only needs the pair's own SW_B(i), so it is published a whole row before the left neighbour uses it.
    // That takes the program-order coupling "E wait of row i -> SW publish of row i + 1" out of the border cycle
    // (E hop -> remainder of the row -> SW hop back), which otherwise limits the row rate to ~1.4 ring round trips.
    // CENSUS staging: the warp's private block of a row, [0, D] = right descriptors of columns xA + dmin + s (a column
    // outside the descriptor row reads as "window leaves the image"), [D + 1], [D + 2] = left descriptors of A and B
    const uint32_t cstg_base = smem_u32(stg) + (uint32_t)(warp * CW) * 4u, cstg_stage = (uint32_t)(nc * CW) * 4u;
    auto stage_pix = [&](int c, int r) {
        if (CENSUS) {
            if (c != 0 || !valid[0] || r >= H) return;        // one request per row, issued with pixel A (one row ahead)
            const uint32_t sg = cstg_base + (uint32_t)(r & (NSTG - 1)) * cstg_stage;
            const uint32_t *rowR = p.descR + (size_t)r * p.pitch, *rowL = p.descL + (size_t)r * p.pitch;
            // Positions whose column lies outside the descriptor row are never copied: they are the same for every row and
            // were filled with the "window leaves the image" flag once, before the loop (census_prefill below).  No lane
            // ever reads a substitute address: copies of one instruction that alias one global word serialise in the LSU,
            // and the slowest strip sets the pace of the whole wave.
            const int colb = xl[0] + p.dmin + lane * NR;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const int sidx = h * VS + j;
                    const int col = colb + sidx;
                    if ((unsigned)col < (unsigned)p.pitch) cp_async_words<1>(sg + lane_b + (uint32_t)sidx * 4u, rowR + col);
                }
            }
            if (lane < 3) {
                const int col = (lane == 0) ? xl[0] + p.dmin + 2 * VS : xl[0] + lane - 1;
                if ((unsigned)col < (unsigned)p.pitch) cp_async_words<1>(sg + (uint32_t)(2 * VS + lane) * 4u, (lane == 0 ? rowR : rowL) + col);
            }
            return;
        }
        if (!valid[c] || r >= H) return;
        const uint32_t sg = stg_base + (uint32_t)(r & (NSTG - 1)) * stg_stage + c * stg_pix;
        if (!FINAL) {
            const uint32_t *src = reinterpret_cast<const uint32_t *>(p.cv) + pix0[c] + (long)r * row_stride + lane * NR;
            cp_async_words<NR>(sg + off0, src);
            cp_async_words<NR>(sg + off1, src + D / 2);
        } else {
            const uint32_t *src = p.buf + pix0[c] + (long)r * row_stride;
            cp_async_words<RW>(sg + off0, src + lane * RW);
            cp_async_words<NR>(sg + off1, src + poff);
        }
    };
    const float nan_code = (float)(p.inv | Tier<CB>::FLAG1);
    const uint32_t p1p1 = p.p1p1, p2p2 = p.p2p2;
    bool bad = false;
    uint32_t rawA[2 * NR], rawAn[2 * NR];                  // CENSUS: pixel A's right descriptors of the current / next row
#pragma unroll
    for (int j = 0; j < 2 * NR; ++j) rawA[j] = rawAn[j] = 0u;
    // unpack pixel c of row r from its staging slot (pass 1: verify + pack + write the cost code; zeros outside the image)
    auto load_pix = [&](int c, int r, uint32_t (&c16)[NR], uint32_t (&p16)[NR]) {
        const uint32_t sg = stg_base + (uint32_t)(r & (NSTG - 1)) * stg_stage + c * stg_pix;
        if (CENSUS) {
            // Hamming costs straight from the descriptors (census.cpp:97-180): popcount(left ^ right), NaN code when either
            // window leaves the image.  Pixel A reads its D right descriptors as two aligned vectors and keeps them: pixel
            // B of the same row (one iteration later) needs the same window shifted by one column, i.e. A's registers
            // moved down by one, the last one coming from the next lane (lane 31: the first word of the other half / the
            // extra word D).
            const uint32_t cg = cstg_base + (uint32_t)(r & (NSTG - 1)) * cstg_stage;
            uint32_t lw[1], ra[NR], rb[NR];
            lds_words<1>(cg + (uint32_t)(2 * VS + 1 + c) * 4u, lw);
            if (c == 0) {
                lds_words<NR>(cg + lane_b, ra);
                lds_words<NR>(cg + (uint32_t)VS * 4u + lane_b, rb);
#pragma unroll
                for (int j = 0; j < NR; ++j) { rawAn[j] = ra[j]; rawAn[NR + j] = rb[j]; }
            } else {
                uint32_t ex[1];
                lds_words<1>(cg + (uint32_t)(2 * VS) * 4u, ex);
                const uint32_t t1 = __shfl_sync(0xffffffffu, rawA[0], (lane + 1) & 31);
                const uint32_t t2 = __shfl_sync(0xffffffffu, rawA[NR], (lane + 1) & 31);
#pragma unroll
                for (int j = 0; j + 1 < NR; ++j) { ra[j] = rawA[j + 1]; rb[j] = rawA[NR + j + 1]; }
                ra[NR - 1] = (lane == 31) ? t2 : t1;
                rb[NR - 1] = (lane == 31) ? ex[0] : t2;
            }
            const uint32_t nan2 = (p.inv | Tier<CB>::FLAG1) * 0x10001u;
            if (lw[0] >> 31) {                                  // warp-uniform: the left window leaves the image
#pragma unroll
                for (int j = 0; j < NR; ++j) c16[j] = nan2;
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const uint32_t xlo = lw[0] ^ ra[j], xhi = lw[0] ^ rb[j];          // bit 31 = the right window leaves the image
                    const uint32_t pk = __byte_perm(__popc(xlo), __popc(xhi), 0x5410);
                    uint32_t fl;                                                     // sign-replicated top bytes: 0xFFFF per flagged half
                    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(fl) : "r"(xlo), "r"(xhi));
                    c16[j] = (pk & ~fl) | (nan2 & fl);
                }
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) p16[j] = 0u;
            if (valid[c] && r < H) st_cost<NR, CB>(p.buf + pix0[c] + (long)r * row_stride, lane, c16);
        } else if (!FINAL) {
            uint32_t fa[NR], fb[NR];
            lds_words<NR>(sg + off0, fa);
            lds_words<NR>(sg + off1, fb);
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                c16[j] = encode_pair(__uint_as_float(fa[j]), __uint_as_float(fb[j]), nan_code, p.cost_ok_max, bad);
                p16[j] = 0u;
            }
            if (valid[c] && r < H) st_cost<NR, CB>(p.buf + pix0[c] + (long)r * row_stride, lane, c16);
        } else {
            uint32_t craw[RW];
            lds_words<RW>(sg + off0, craw);
            lds_words<NR>(sg + off1, p16);
            unpack_cost<NR, CB>(craw, c16);
        }
        if (!valid[c] || r >= H) {
#pragma unroll
            for (int j = 0; j < NR; ++j) c16[j] = p16[j] = 0u;              // outside the image: zero costs, flat states
        }
    };
    if (CENSUS && valid[0]) {                              // census_prefill: flag words of the positions no copy ever writes
        const int colb = xl[0] + p.dmin + lane * NR;
        for (int sl = 0; sl < NSTG; ++sl) {
            const uint32_t sg = cstg_base + (uint32_t)sl * cstg_stage;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const int sidx = h * VS + j;
                    if (!((unsigned)(colb + sidx) < (unsigned)p.pitch)) sts_u32(sg + lane_b + (uint32_t)sidx * 4u, 0x80000000u);
                }
            }
            if (lane < 3) {
                const int col = (lane == 0) ? xl[0] + p.dmin + 2 * VS : xl[0] + lane - 1;
                if (!((unsigned)col < (unsigned)p.pitch)) sts_u32(sg + (uint32_t)(2 * VS + lane) * 4u, 0x80000000u);
            }
        }
        __syncwarp();
    }
    stage_pix(0, 0);
    cp_async_commit();
    for (int r = 0; r < PFD; ++r) {
        stage_pix(0, r + 1);
        stage_pix(1, r);
        cp_async_commit();
    }

    uint32_t Sv[2][NR], SEA_prev[NR], c16A[NR], p16A[NR], SWA_cur[NR], zero[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) Sv[0][j] = Sv[1][j] = SEA_prev[j] = zero[j] = 0u;
    cp_async_wait<PFD>();                                  // pixel A of row 0
    if (CENSUS) __syncwarp();                              // a lane reads descriptors its neighbours copied
    load_pix(0, 0, c16A, p16A);
#pragma unroll
    for (int j = 0; j < 2 * NR; ++j) rawA[j] = rawAn[j];
    {
        uint32_t ccA[NR];
#pragma unroll
        for (int j = 0; j < NR; ++j) ccA[j] = c16A[j] & Tier<CB>::VALUES;
        nstep<NR>(ccA, zero, SWA_cur, lane, p1p1, p2p2);     // SW_A(0): a path start
        sts_words<NR>(sw_base + vme * VB, SWA_cur);
        ev4_signal(fs_base + vme * 32u, 0, lane);
    }

    // BATCH: a batch of images stacked into one tall image (p.period rows each; the volume, the descriptors and the disparity
    // map of image k + 1 follow those of image k, so no address changes): the wave runs through all of them and its fill and
    // drain across the strips are paid once per pass and batch.  Every vertical / diagonal path restarts at an image's first
    // row (in travel order): the PRODUCERS of the hand-over states publish zeros -- a flat state, i.e. a path start -- for an
    // image's last row (SW_A, SE_B in the mailboxes, and through the relays in the rings), the in-register states (S,
    // the pair's own SE_A -> SE_B and SW_B -> SW_A) are reset.  Waits and events are untouched.  The row program is a
    // lambda instantiated twice in that case: the hot loop carries none of the boundary conditions (this kernel sits at the
    // register limit: two more live values in the loop cost 10 %), the last two rows of every image run the TAIL copy.
    int iend = H;
    auto row = [&](const int i, auto tail_tag) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const uint32_t tag = (uint32_t)(i + 1);
        stage_pix(0, i + PFD + 1);
        stage_pix(1, i + PFD);
        cp_async_commit();
        cp_async_wait<PFD>();                              // pixel A of row i + 1 and pixel B of row i have landed
        if (CENSUS) __syncwarp();
        uint32_t c16[2][NR], p16[2][NR], cc[2][NR], c16An[NR], p16An[NR], ccAn[NR];
        load_pix(0, i + 1, c16An, p16An);
        load_pix(1, i, c16[1], p16[1]);
#pragma unroll
        for (int j = 0; j < NR; ++j) {
            c16[0][j] = c16A[j];
            p16[0][j] = p16A[j];
            cc[0][j] = c16A[j] & Tier<CB>::VALUES;
            cc[1][j] = c16[1][j] & Tier<CB>::VALUES;
            ccAn[j] = c16An[j] & Tier<CB>::VALUES;
        }

        // ---- phase 1: everything that does not need this row's E chain ----------------------------------
        uint32_t L_SE_B[NR], L0[NR], SW_in[NR], L_SW_B[NR], SWA_next[NR];
        // the right neighbour's SW_A(i - 1): published two of its rows ago
#pragma unroll
        for (int j = 0; j < NR; ++j) SW_in[j] = 0u;
        if (i > 0) {
            if (!nowait && wait_r) ev4_wait(fs_base + (vme + 1u) * 32u, i - 1);
            lds_words<NR>(sw_base + (uint32_t)((i - 1) & 3) * SLOT + (vme + 1u) * VB, SW_in);
        }
        // Independent recurrence steps sit in one basic block (no flag, no warp barrier between them) so that the
        // scheduler can fill the latency bubbles of one chain (min tree -> redux -> per-register ops) with the others:
        // SW_B, SE_B and S_B only need values at hand; SW_A(i + 1) follows SW_B, S_A is independent.
        nstep<NR>(cc[1], SW_in, L_SW_B, lane, p1p1, p2p2);                   // SW_B(i)
        nstep<NR>(cc[1], SEA_prev, L_SE_B, lane, p1p1, p2p2);                // SE_B(i) from the pair's own SE_A(i - 1)
        nstep<NR>(cc[1], Sv[1], L0, lane, p1p1, p2p2);                       // S_B(i)
#pragma unroll
        for (int j = 0; j < NR; ++j) Sv[1][j] = L0[j];
        if (TAIL && i == iend - 1) nstep<NR>(ccAn, zero, SWA_next, lane, p1p1, p2p2);     // row i + 1 starts an image: a path start
        else nstep<NR>(ccAn, L_SW_B, SWA_next, lane, p1p1, p2p2);                       // SW_A(i + 1), one row ahead
        nstep<NR>(cc[0], Sv[0], L0, lane, p1p1, p2p2);                       // S_A(i)
#pragma unroll
        for (int j = 0; j < NR; ++j) Sv[0][j] = L0[j];
        // SW_A(i + 1) is the left neighbour's predecessor in row i + 2, SE_B(i) the right neighbour's in row i + 1: zeros when
        // that row starts an image
        if (TAIL && i == iend - 2) sts_words<NR>(sw_base + (uint32_t)((i + 1) & 3) * SLOT + vme * VB, zero);
        else sts_words<NR>(sw_base + (uint32_t)((i + 1) & 3) * SLOT + vme * VB, SWA_next);
        if (TAIL && i == iend - 1) sts_words<NR>(se_base + (uint32_t)(i & 3) * SLOT + vme * VB, zero);
        else sts_words<NR>(se_base + (uint32_t)(i & 3) * SLOT + vme * VB, L_SE_B);
        ev4_signal(fs_base + vme * 32u, i + 1, lane);

        // ---- phase 2: the E chain -- wait, two steps, publish; SE_A (the left neighbour's SE state of the previous row is
        // visible since its E flag of this row) fills the bubbles of the E_A -> E_B chain ----------------------------------
        uint32_t E_in[NR], SE_in[NR], L_E_A[NR], L_E_B[NR], L_SE_A[NR];
        if (!nowait && wait_l) ev4_wait(fe_base + (vme - 1u) * 32u, i);
        lds_words<NR>(e_base + (uint32_t)(i & 3) * SLOT + (vme - 1u) * VB, E_in);
#pragma unroll
        for (int j = 0; j < NR; ++j) SE_in[j] = 0u;
        if (i > 0) lds_words<NR>(se_base + (uint32_t)((i - 1) & 3) * SLOT + (vme - 1u) * VB, SE_in);
        nstep<NR>(cc[0], E_in, L_E_A, lane, p1p1, p2p2);
        nstep<NR>(cc[0], SE_in, L_SE_A, lane, p1p1, p2p2);
        nstep<NR>(cc[1], L_E_A, L_E_B, lane, p1p1, p2p2);
        sts_words<NR>(e_base + (uint32_t)(i & 3) * SLOT + vme * VB, L_E_B);
        ev4_signal(fe_base + vme * 32u, i, lane);

#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t tot[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j)
                tot[j] = p16[c][j] + Sv[c][j] + (c == 0 ? (L_E_A[j] + L_SE_A[j] + SWA_cur[j]) : (L_E_B[j] + L_SE_B[j] + L_SW_B[j]));
            if (valid[c]) {
                uint32_t *gpix = p.buf + pix0[c] + (long)i * row_stride;
                if (!FINAL) {
                    st_words<NR>(gpix + poff, tot);
                } else {
                    float fa[NR], fb[NR];
                    uint32_t bl = 0xFFFFFFFFu, bh = 0xFFFFFFFFu;
#pragma unroll
                    for (int j = 0; j < NR; ++j) {
                        uint32_t t = tot[j];
                        if (p.overcounting) t = t - 7u * cc[c][j];   // S >= 8 C in every half: no borrow
                        // 16-bit integer -> float32: PRMT builds 0x4B00'nnnn, one FADD removes the 2^23.  A NaN cell gets the
                        // upper half 0x7F80 / 0x7FFF instead of 0x4B00: exponent all ones over a non-zero mantissa (its sum
                        // is >= invalid_value > 0), i.e. a NaN that the same FADD passes through -- no select.
                        const uint32_t fl = c16[c][j] & Tier<CB>::FLAGS;
                        const uint32_t sat = (CB == 1) ? fl * 0x1FFu : (fl >> 15) * 0xFFFFu;     // 0xFF80 / 0xFFFF per NaN half
                        const uint32_t hx = 0x4B004B00u | (sat & 0x34FF34FFu);
                        fa[j] = __uint_as_float(__byte_perm(t, hx, 0x5410)) - 8388608.0f;
                        fb[j] = __uint_as_float(__byte_perm(t, hx, 0x7632)) - 8388608.0f;
                        if (WTA) {
                            // NaN cells saturate their half (no valid sum reaches the sentinel), then one 32-bit key per
                            // half: (sum << 16) | register index; the lane offset of the disparity is added once below
                            const uint32_t tk = t | sat;
                            bl = min(bl, __byte_perm(tk, (uint32_t)j, 0x1054));      // (low sum << 16) | j
                            bh = min(bh, __byte_perm(tk, (uint32_t)j, 0x3254));      // (high sum << 16) | j
                        }
                    }
                    float *o = reinterpret_cast<float *>(gpix) + lane * NR;
                    st_floats<NR>(o, fa);
                    st_floats<NR>(o + D / 2, fb);
                    if (WTA) {
                        uint32_t best = min(bl + (uint32_t)(lane * NR), bh + (uint32_t)(D / 2 + lane * NR));
                        best = __reduce_min_sync(0xffffffffu, best);
                        if (lane == 0) {
                            const size_t pix = pixi0[c] + (long)i * (FINAL ? -(long)W : (long)W);
                            const bool none = (best >> 16) >= ((CB == 1) ? 0xFF80u : 0xFFFFu);
                            p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(best & 0xFFFFu));
                            if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) { SEA_prev[j] = L_SE_A[j]; SWA_cur[j] = SWA_next[j]; c16A[j] = c16An[j]; p16A[j] = p16An[j]; }
        if (CENSUS) {
#pragma unroll
            for (int j = 0; j < 2 * NR; ++j) rawA[j] = rawAn[j];
        }
    };
    if (!BATCH) {
#pragma unroll 2
        for (int i = 0; i < H; ++i) row(i, std::false_type{});
    } else {
        const int Hp = (p.period > 0 && p.period < H) ? p.period : H;
        for (int ibeg = 0; ibeg < H; ibeg += Hp) {
            iend = min(H, ibeg + Hp);
            if (ibeg > 0) {
#pragma unroll
                for (int j = 0; j < NR; ++j) Sv[0][j] = Sv[1][j] = SEA_prev[j] = 0u;
            }
            int i = ibeg;
#pragma unroll 2
            for (; i < iend - 2; ++i) row(i, std::false_type{});
#pragma unroll 1
            for (; i < iend; ++i) row(i, std::true_type{});
        }
    }
    if (!FINAL && !CENSUS && __any_sync(0xffffffffu, bad) && lane == 0) atomicOr(p.flag, 1);
}

// the two wavefront passes; CENSUS: the first one computes its costs from census descriptors (p.descL / p.descR)
template <int NR, int CB, bool CENSUS>
int launch_wave(NarrowParams p, int nstrips, int nwarp, void *workspace, cudaStream_t s, bool *done) {
    *done = false;
    const int nsm = sm_count();
    const bool wta = p.disp != nullptr;
    void (*w1)(const NarrowParams) = sgm_wave_kernel<NR, CB, false, false, CENSUS>;
    void (*w2)(const NarrowParams) = wta ? sgm_wave_kernel<NR, CB, true, true> : sgm_wave_kernel<NR, CB, true, false>;
    if (p.period > 0) {                               // a batch of images in one wave: pb200_census_sgm_batch (fused WTA only)
        if (!CENSUS || !wta) return PB200_OK;
        w1 = sgm_wave_kernel<NR, CB, false, false, CENSUS, true>;
        w2 = sgm_wave_kernel<NR, CB, true, true, false, true>;
    }
    const int wthreads = (nwarp + 2) * 32;                       // + the two relay warps
    const size_t state = ((size_t)12 * (nwarp + 2) * NR * 32 + 64 + 256) * sizeof(uint32_t);
    const size_t smem1 = state + (CENSUS ? (size_t)8 * nwarp * (2 * NR * 32 + 8) : (size_t)4 * nwarp * 2 * 32 * (2 * NR)) * sizeof(uint32_t);
    const size_t smem2 = state + (size_t)4 * nwarp * 2 * 32 * (NR * CB / 2 + NR) * sizeof(uint32_t);
    int occ1 = 0, occ2 = 0;
    if (wthreads <= 512 && smem1 <= 220 * 1024 && smem2 <= 220 * 1024) {
        PB200_CUDA(cudaFuncSetAttribute((const void *)w1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        PB200_CUDA(cudaFuncSetAttribute((const void *)w2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, (const void *)w1, wthreads, smem1));
        PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, (const void *)w2, wthreads, smem2));
    }
    if ((long)occ1 * nsm < nstrips || (long)occ2 * nsm < nstrips) return PB200_OK;
    p.ring = reinterpret_cast<unsigned long long *>(workspace);
    const size_t wring = (size_t)nstrips * 12 * NR * 32 * sizeof(unsigned long long);
    if ((size_t)(reinterpret_cast<char *>(p.flag) - reinterpret_cast<char *>(workspace)) < wring) return PB200_OK;   // ring must end before the flag
    PB200_CUDA(cudaMemsetAsync(p.flag, 0, sizeof(int), s));
    void *args[] = {(void *)&p};
    PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
    PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w1, dim3(nstrips), dim3(wthreads), args, smem1, s));
    PB200_LAUNCH_CHECK(CENSUS ? "sgm_wave_kernel<down, census>" : "sgm_wave_kernel<down>");
    PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
    PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w2, dim3(nstrips), dim3(wthreads), args, smem2, s));
    PB200_LAUNCH_CHECK("sgm_wave_kernel<up>");
    note_path(STAGE_SGM, CENSUS ? PATH_SGM_WAVE2_CENSUS : PATH_SGM_WAVE2, NR * 10 + CB);
    *done = true;
    return PB200_OK;
}

enum { NARROW_ALL = 0, NARROW_H = 1, NARROW_V = 2 };

template <int NR, int CB>
int launch_narrow(NarrowParams p, int phase, int final, int nstrips, int nwarp, void *workspace, size_t ring_bytes, cudaStream_t s,
                  bool *done) {
    *done = false;
    const int nsm = sm_count();
    const int K = nwarp * 2;
    const size_t smem = (size_t)2 * 2 * (K + 2) * NR * 32 * sizeof(uint32_t) +                 // state buffers
                        (size_t)4 * nwarp * 2 * 32 * (NR * CB / 2 + NR) * sizeof(uint32_t);    // input staging ring
    if (smem > 220 * 1024) return PB200_OK;
    const bool wta = p.disp != nullptr;
    // wavefront path: the whole stage in two 4-direction passes (single-call runs without tile halos)
    if (phase == NARROW_ALL && p.halo_in == nullptr && p.halo_out == nullptr && option(OPT_SGM_NO_WAVE) <= 0) {
        const int rc = launch_wave<NR, CB, false>(p, nstrips, nwarp, workspace, s, done);
        if (rc != PB200_OK || *done) return rc;
    }
    const int threads = (nwarp + 1) * 32;
    void (*mid)(const NarrowParams) = sgm_narrow_vsweep_kernel<NR, CB, false, false>;
    void (*fin)(const NarrowParams) = wta ? sgm_narrow_vsweep_kernel<NR, CB, true, true> : sgm_narrow_vsweep_kernel<NR, CB, true, false>;
    int per_sm = 0;
    PB200_CUDA(cudaFuncSetAttribute((const void *)mid, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB200_CUDA(cudaFuncSetAttribute((const void *)fin, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)mid, threads, smem));
    if ((long)per_sm * nsm < nstrips) return PB200_OK;
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void *)fin, threads, smem));
    if ((long)per_sm * nsm < nstrips) return PB200_OK;

    if (phase != NARROW_V) {
        PB200_CUDA(cudaMemsetAsync(p.flag, 0, sizeof(int), s));
        const int hgrid = ceil_div(p.H, 4);
        sgm_narrow_h_kernel<NR, CB, true><<<hgrid, 128, 0, s>>>(p);
        PB200_LAUNCH_CHECK("sgm_narrow_h_kernel<E>");
        sgm_narrow_h_kernel<NR, CB, false><<<hgrid, 128, 0, s>>>(p);
        PB200_LAUNCH_CHECK("sgm_narrow_h_kernel<W>");
    }
    if (phase != NARROW_H) {
        p.ring = reinterpret_cast<unsigned long long *>(workspace);
        const int npass = (phase == NARROW_ALL) ? 2 : 1;
        for (int pass = 0; pass < npass; ++pass) {
            const bool is_final = (phase == NARROW_ALL) ? (pass == 1) : (final != 0);
            if (phase == NARROW_ALL) p.dy = pass == 0 ? 1 : -1;
            PB200_CUDA(cudaMemsetAsync(p.ring, 0, ring_bytes, s));
            void *args[] = {(void *)&p};
            PB200_CUDA(cudaLaunchCooperativeKernel((const void *)(is_final ? fin : mid), dim3(nstrips), dim3(threads), args, smem, s));
            PB200_LAUNCH_CHECK("sgm_narrow_vsweep_kernel");
        }
    }
    note_path(STAGE_SGM, PATH_SGM_PACKED4, NR * 10 + CB);
    *done = true;
    return PB200_OK;
}

bool is_small_int(float v, int lo, int hi) { return v >= (float)lo && v <= (float)hi && v == (float)(int)v; }

}  // namespace

// Try the packed-integer path.  On return *gate is NULL when nothing was launched (the caller runs the float
// kernels unconditionally) or points to the device flag the float kernels must be gated on.
// phase 0: the whole stage in one call; 1: the horizontal pair only (data check, C16 / P16 left in `out`);
// 2: one vertical group (dy, final) with optional packed halos -- the split used by row-tiled multi-GPU runs;
// 3: launch nothing, only report the flag when the shape / parameters are eligible.
int sgm_narrow_try(const float *cv, float *out, int H, int W, int D, float p1, float p2, float invalid_value, int overcounting,
                   float *disp, int dmin, float invalid_disparity, uint8_t *all_nan, void *workspace, size_t workspace_bytes,
                   cudaStream_t s, const int **gate, int phase, int dy, int final, const float *halo_in, float *halo_out) {
    *gate = nullptr;
    if (D != 64 && D != 128 && D != 192 && D != 256) return PB200_OK;
    if (!is_small_int(p1, 1, NARROW_MAX) || !is_small_int(p2, 1, NARROW_MAX) || p1 > p2 || !is_small_int(invalid_value, 0, NARROW_MAX) ||
        (int)invalid_value + (int)p2 > NARROW_MAX)
        return PB200_OK;
    if ((reinterpret_cast<uintptr_t>(cv) & 15) || (reinterpret_cast<uintptr_t>(out) & 15)) return PB200_OK;
    const int nsm = sm_count();
    int K = ceil_div(W, nsm);
    if (K < 4) K = 4;
    K = (K + 1) / 2 * 2;
    if (K / 2 > 15) return PB200_OK;
    const int NR = D / 64;
    const size_t flag_off = sgm_ring_max_bytes(W, D) + 256;
    if (workspace == nullptr || workspace_bytes < flag_off + sizeof(int) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PB200_OK;

    NarrowParams p{};
    p.cv = cv; p.buf = reinterpret_cast<uint32_t *>(out); p.H = H; p.W = W; p.D = D;
    p.p1p1 = (uint32_t)p1 * 0x10001u; p.p2p2 = (uint32_t)p2 * 0x10001u;
    p.inv = (uint32_t)invalid_value;
    // byte tier (C8, and P8 between E and W) when cost + P2 fits 7 bits; else 16-bit storage
    const bool bytes = (NR == 2 || NR == 4) && ((int)invalid_value + (int)p2 <= 127) && option(OPT_SGM_NO_BYTE_TIER) <= 0;
    p.cost_ok_max = (float)((bytes ? 127 : NARROW_MAX) - (int)p2);
    p.flag = reinterpret_cast<int *>(reinterpret_cast<char *>(workspace) + flag_off);
    p.dy = dy; p.overcounting = overcounting;
    p.disp = disp; p.all_nan = all_nan; p.dmin = dmin; p.invalid_disparity = invalid_disparity;
    p.ring = nullptr;
    p.halo_in = reinterpret_cast<const uint32_t *>(halo_in);
    p.halo_out = reinterpret_cast<uint32_t *>(halo_out);
    if ((reinterpret_cast<uintptr_t>(halo_in) & 15) || (reinterpret_cast<uintptr_t>(halo_out) & 15)) return PB200_OK;
    p.debug = debug_switches();
    p.descL = p.descR = nullptr; p.pitch = 0; p.half = 0;
    if (phase == 3) {                 // query only: eligible -> the caller gates its float kernels on the flag
        *gate = p.flag;
        return PB200_OK;
    }
    bool done = false;
    int rc;
    const int nwarp = K / 2;
    const int nstrips = ceil_div(W, K);
    const size_t ring_bytes = (size_t)nstrips * 2 * 2 * NR * 32 * sizeof(unsigned long long);
    if (NR == 4) rc = bytes ? launch_narrow<4, 1>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done)
                            : launch_narrow<4, 2>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done);
    else if (NR == 2) rc = bytes ? launch_narrow<2, 1>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done)
                                 : launch_narrow<2, 2>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done);
    else if (NR == 3) rc = launch_narrow<3, 2>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done);   // D = 192: 16-bit tier only
    else rc = launch_narrow<1, 2>(p, phase, final, nstrips, nwarp, workspace, ring_bytes, s, &done);
    if (rc != PB200_OK) return rc;
    if (done) *gate = p.flag;
    return PB200_OK;
}

int sgm_census_wave1_launch(NarrowParams p, int NR, bool bytes, void *workspace, size_t ring_room, const Wave1Peers *peers, cudaStream_t s,
                            bool *done);   // sgm_wave1.cu

// Which fused Census -> SGM kernels take a configuration: 0 = none (the caller runs the Census fill and pb200_sgm), 1 = the
// skewed one-column wavefront (sgm_wave1.cu; reads the shifted descriptor layout), 2 = the two-column wavefront.
static bool census_wave_common(int window, int D, float p1, float p2) {
    if (D != 64 && D != 128 && D != 192 && D != 256) return false;
    if (window != 3 && window != 5) return false;                          // one-word descriptors
    const float invalid_value = (float)(window * window) + p2 + 1.f;       // cmax + P2 + 1 (census.py:116 gives cmax = w^2)
    return is_small_int(p1, 1, NARROW_MAX) && is_small_int(p2, 1, NARROW_MAX) && p1 <= p2 && is_small_int(invalid_value, 0, NARROW_MAX) &&
           (int)invalid_value + (int)p2 <= NARROW_MAX;
}
#ifndef PB200_PREFER_SKEWED
#define PB200_PREFER_SKEWED 0
#endif
int sgm_wave1_strip_width(int W);                                           // sgm_wave1.cu: 0 when the image is too wide for one wave
int sgm_census_plan(int window, int W, int D, float p1, float p2) {
    if (!census_wave_common(window, D, p1, p2) || option(OPT_SGM_NO_WAVE) > 0) return 0;
    const int pin = option(OPT_SGM_WAVE_KERNEL);
    int K = ceil_div(W, sm_count());
    if (K < 4) K = 4;
    K = (K + 1) / 2 * 2;
    const bool two_ok = K / 2 <= 14, one_ok = sgm_wave1_strip_width(W) > 0 && D != 192;    // D = 192 (three registers per lane): two-column kernels only
    if (pin == 1) return one_ok ? 1 : 0;
    if (pin == 2) return two_ok ? 2 : 0;
    // one GPU: whichever is faster for the shape (measured at C3: profiles/r2_wave_kernels.txt); column tiles pin the skewed one
    if (PB200_PREFER_SKEWED && one_ok) return 1;
    return two_ok ? 2 : (one_ok ? 1 : 0);
}

// Fused Census -> SGM: the two wavefront passes with the first one computing the Hamming costs from the census
// descriptors (no float cost volume is written or read).  Census costs are integers in [0, window^2] by construction,
// so the data condition of the packed path holds statically: no flag, no float fall-back.  *done = false when the
// shape / parameters are not eligible (the caller then runs the Census fill and pb200_sgm separately).
int sgm_census_wave_try(const CensusDesc &desc, int window, float *out, int H, int W, int D, float p1,
                        float p2, int overcounting, float *disp, int dmin, float invalid_disparity, uint8_t *all_nan, void *workspace,
                        size_t workspace_bytes, cudaStream_t s, bool *done, const Wave1Peers *peers, int period) {
    // `peers` != NULL: W is the width of this GPU's column tile of a peers->Wg wide image (skewed wavefront only)
    // `period` > 0: H rows are a batch of H / period images stacked into one tall image (two-column wavefront only)
    *done = false;
    const int plan = sgm_census_plan(window, W, D, p1, p2);
    if (plan == 0 || (plan == 1 && desc.R4 == nullptr) || (peers != nullptr && plan != 1)) return PB200_OK;
    if (period > 0 && (plan != 2 || period < 4 || H % period != 0 || disp == nullptr)) return PB200_OK;
    if (reinterpret_cast<uintptr_t>(out) & 15) return PB200_OK;
    const float invalid_value = (float)(window * window) + p2 + 1.f;
    const int NR = D / 64;
    const size_t flag_off = sgm_ring_max_bytes(W, D) + 256;
    if (workspace == nullptr || workspace_bytes < flag_off + sizeof(int) || (reinterpret_cast<uintptr_t>(workspace) & 15)) return PB200_OK;

    NarrowParams p{};
    p.cv = nullptr; p.buf = reinterpret_cast<uint32_t *>(out); p.H = H; p.W = W; p.D = D;
    p.p1p1 = (uint32_t)p1 * 0x10001u; p.p2p2 = (uint32_t)p2 * 0x10001u;
    p.inv = (uint32_t)invalid_value;
    const bool bytes = (NR == 2 || NR == 4) && ((int)invalid_value + (int)p2 <= 127) && option(OPT_SGM_NO_BYTE_TIER) <= 0;
    p.cost_ok_max = 0.f;
    p.flag = reinterpret_cast<int *>(reinterpret_cast<char *>(workspace) + flag_off);
    p.dy = 1; p.overcounting = overcounting;
    p.disp = disp; p.all_nan = all_nan; p.dmin = dmin; p.invalid_disparity = invalid_disparity;
    p.ring = nullptr; p.halo_in = nullptr; p.halo_out = nullptr;
    p.debug = debug_switches();
    p.descL = desc.L; p.descR = desc.R; p.pitch = desc.pitch; p.half = window / 2;
    p.descR4 = desc.R4; p.pitch4 = desc.pitch4; p.padl = desc.padl;
    p.period = period;
    if (plan == 1) return sgm_census_wave1_launch(p, NR, bytes, workspace, flag_off - 256, peers, s, done);
    int K = ceil_div(W, sm_count());
    if (K < 4) K = 4;
    K = (K + 1) / 2 * 2;
    const int nwarp = K / 2;
    const int nstrips = ceil_div(W, K);
    if (NR == 4) return bytes ? launch_wave<4, 1, true>(p, nstrips, nwarp, workspace, s, done) : launch_wave<4, 2, true>(p, nstrips, nwarp, workspace, s, done);
    if (NR == 2) return bytes ? launch_wave<2, 1, true>(p, nstrips, nwarp, workspace, s, done) : launch_wave<2, 2, true>(p, nstrips, nwarp, workspace, s, done);
    if (NR == 3) return launch_wave<3, 2, true>(p, nstrips, nwarp, workspace, s, done);
    return launch_wave<1, 2, true>(p, nstrips, nwarp, workspace, s, done);
}

}  // namespace pb200
