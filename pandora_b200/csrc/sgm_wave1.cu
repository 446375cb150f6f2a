// sgm_wave1.cu -- the wavefront SGM passes with ONE COLUMN PER WARP.
//
// Same arithmetic, same two passes and the same intermediate format as sgm_wave_kernel (sgm_narrow.cu): pass 1 walks the
// image top-down and runs E, SE, S, SW; pass 2 walks it bottom-up on the frame flipped in both axes (W, NW, N, NE), adds
// its four directions to the 16-bit partial sums of pass 1 and emits float32 S (+ NaN restore, overcounting, WTA) in
// place.  What changes is the shape of the machine:
//
//   * a strip is K <= 28 columns and a CTA has one compute warp per column (28 instead of 14 recurrence chains per SM:
//     the two-column kernel ran at 3.5 warps per scheduler and was bound by the dependent-issue latency of its chains,
//     profiles/r1_ncu_fused_pass1.txt), at <= 64 registers per thread: a warp keeps only its vertical state S, the
//     pixel's cost and the running total; E / SE come from the left neighbour's mailbox, SW from the right one's;
//   * mailboxes are two-deep for E and SW and four-deep for SE (slot discipline below), progress counters per column;
//   * the first pass takes its Census costs from descriptor rows staged ONCE per CTA row by a loader warp (cp.async,
//     four word-shifted copies so that every column reads its D-wide window as aligned vectors) instead of once per warp;
//   * the relay warps poll both directions independently (a small two-counter state machine), so a late E of this strip
//     never delays the SW coming in from the next one;
//   * the strips at the two ends of the image may hand their border states to ANOTHER GPU: the relay then stores into
//     peer-mapped memory (NVLink) and polls a local buffer the neighbour stores into -- a column-tiled multi-GPU run is
//     one wave across all GPUs, with no host-side step in between (pandora_b200/tiling.py).
//
// Row program of column x (mailbox column v = x - strip start + 1), row i in travel order:
//   1  cost c(i)                                         (descriptor ring / cp.async staging)
//   2  wait fs[v+1] >= i;      SW(i) = step(c, sw[(i-1)&1][v+1]);  -> sw[i&1][v],  fs[v] = i+1
//   3  S(i) = step(c, S(i-1))                            (registers)
//   4  wait fe[v-1] >= i+1;    E(i)  = step(c, e[i&1][v-1]);       -> e[i&1][v],   fe[v] = i+1
//   5  SE(i) = step(c, se[(i-1)&3][v-1]);                          -> se[i&3][v]   (visible with fe[v] = i+2)
//   6  total = [P16 +] SW + S + E + SE -> global
// Slot discipline: x overwrites sw[i&1] at row i+2, which it reaches only after E(i+1, x-1), published after x-1 read
// SW(i, x) in its row i+1; x overwrites e[i&1] at row i+2 step 4, i.e. after SW(i+1, x+1), published after x+1 read
// E(i, x) in its row i; se[i&3] is overwritten at row i+4, three SW hand-overs later.  Row 0 needs no special case: the
// counters start at 0, the slots it reads ((i-1)&1 = 1, (i-1)&3 = 3) are still zero = a flat state = a path start, and
// an image border is a mailbox column whose counter is "infinity" and whose slots stay zero.
#include "sgm_packed.cuh"

namespace pb200 {

namespace {

constexpr int W1_MAXK = 28;                 // columns (compute warps) per strip
constexpr int W1_THREADS = (W1_MAXK + 3) * 32;   // + left relay, right relay, loader
constexpr int W1_RING = 16;                 // descriptor rows in the loader's ring
constexpr int W1_PF = 4;                    // descriptor rows in flight
constexpr int W1_FLAG_WORDS = 272;          // 128 mbarriers (FE[32][2] | FS[32][2]) + the loader's two counters

template <int NR> struct W1Cpw { static constexpr int value = W1_MAXK + 2 * NR * 32 + 4; };   // words per shifted copy of a descriptor row

template <int NR>
__device__ __forceinline__ bool ll_try_recv_u32(const unsigned long long *slot, int lane, uint32_t tag, uint32_t (&v)[NR]) {
    unsigned long long w[NR];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NR; ++j) w[j] = ll_load(slot + j * 32 + lane);
#pragma unroll
    for (int j = 0; j < NR; ++j) ok = ok && ((uint32_t)(w[j] >> 32) == tag);
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = (uint32_t)w[j];
    return __all_sync(0xffffffffu, ok);
}
__device__ __forceinline__ uint32_t flag_peek(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
// every lane stores the same counter (one shared-memory transaction, no divergent branch); the warp barrier in front
// orders the other lanes' data stores before it
__device__ __forceinline__ void flag_set(uint32_t addr, uint32_t v) {
    __syncwarp();
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// a counter that is only read now and then (loader ring): wait without occupying the issue slots of the working warps
__device__ __forceinline__ void flag_wait_sleep(uint32_t addr, uint32_t target) {
    while (flag_peek(addr) < target) __nanosleep(100);
}

// Hand-over events between warps are mbarrier PHASES, not spin flags: a consumer that is early is suspended by the
// hardware (mbarrier.try_wait) instead of burning issue slots in an LDS / compare / branch loop -- in the two-column
// kernel those loops were 44 % of all issued instructions (profiles/r1_ncu_fused_pass1.txt).  Every event stream (the E
// hand-over of a column, its SW hand-over) owns TWO barriers used by even and odd rows in turn; row i completes phase
// i >> 1 of barrier i & 1.  The protocol never lets a producer run two rows ahead of its consumer's wait, so a barrier
// is at most one completed phase ahead of the phase a waiter asks for, which is what parity waits can distinguish.
__device__ __forceinline__ void ev_wait(uint32_t bar, int i) {
    const uint32_t addr = bar + (uint32_t)(i & 1) * 8u, parity = (uint32_t)(i >> 1) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra W1_DONE;\n"
        "bra W1_WAIT;\n"
        "W1_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ bool ev_test(uint32_t bar, int i) {
    const uint32_t addr = bar + (uint32_t)(i & 1) * 8u, parity = (uint32_t)(i >> 1) & 1u;
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
    return ok != 0;
}
// one arrival (lane 0, predicated: no divergent branch) with release semantics; the warp barrier in front orders the
// other lanes' data stores before it
__device__ __forceinline__ void ev_signal(uint32_t bar, int i, int lane) {
    const uint32_t addr = bar + (uint32_t)(i & 1) * 8u;
    __syncwarp();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(addr), "r"(lane) : "memory");
}

template <int NR, int CB, bool FINAL, bool WTA, bool CENSUS>
__global__ void __launch_bounds__(W1_THREADS, 1) sgm_wave1_kernel(const NarrowParams p) {
    static_assert(!(CENSUS && FINAL), "the Census source only exists for the first pass");
    static_assert(FINAL || CENSUS, "the float-input first pass stays with sgm_wave_kernel");
    extern __shared__ __align__(16) uint32_t w1_smem[];
    constexpr int VS = NR * 32;                          // words per packed state vector
    constexpr int RW = NR * CB / 2;                      // raw cost words per lane
    constexpr int CPW = W1Cpw<NR>::value;
    constexpr int CSLOT = 4 * CPW + 32;                  // words per descriptor-ring row: four shifted copies + the left descriptors
    constexpr int NSTG = 4, PFD = 3;                     // pass 2 input staging (per warp, cp.async)
    constexpr int SIN = RW + NR;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K, NV = K + 2;
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int H = p.H, W = p.W, D = p.D;
    const int x0 = strip * K;
    const int nce = min(K, W - x0);                      // columns of this strip inside the image
    // shared: e[2][NV][VS] | se[4][NV][VS] | sw[2][NV][VS] | events FE[32][2] FS[32][2] (mbarriers) | ld, done | staging
    const int mbox_words = 8 * NV * VS, flag_words = W1_FLAG_WORDS;
    const int stage_words = CENSUS ? W1_RING * CSLOT : NSTG * K * 32 * SIN;
    for (int i = threadIdx.x; i < mbox_words + flag_words + stage_words; i += blockDim.x) w1_smem[i] = 0u;
    __syncthreads();
    const bool has_left = strip > 0 || p.peer_in_l != nullptr;
    const bool has_right = strip + 1 < nstrips || p.peer_in_r != nullptr;
    if (threadIdx.x < 128)                                                       // FE[32][2] | FS[32][2], one arrival per phase
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(w1_smem + mbox_words) + threadIdx.x * 8u) : "memory");
    __syncthreads();
    const uint32_t lane_b = (uint32_t)(lane * NR) * 4u;
    const uint32_t VB = (uint32_t)VS * 4u, SLOTB = (uint32_t)(NV * VS) * 4u;      // bytes per vector / per mailbox slot
    const uint32_t e_base = smem_u32(w1_smem) + lane_b;
    const uint32_t se_base = e_base + 2u * SLOTB;
    const uint32_t sw_base = se_base + 4u * SLOTB;
    const uint32_t fe_base = smem_u32(w1_smem) + (uint32_t)mbox_words * 4u, fs_base = fe_base + 512u;   // event v: base + 16 v
    const uint32_t ld_flag = fe_base + 1024u, done_flag = fe_base + 1028u;
    const uint32_t stg_base0 = fe_base + (uint32_t)flag_words * 4u;
    const uint32_t tag0 = p.tag_base;

    // ---- relay warps: mailbox <-> ring (the local L2 ring between strips, or a neighbouring GPU's memory at a tile edge).
    // Ring block of a boundary: 12 vectors of VS 64-bit {tag, value} words: e[4] | se[4] | sw[4].
    if (warp == K) {                                      // left relay: E / SE in, SW out
        if (!has_left) return;
        const unsigned long long *rin = strip > 0 ? p.ring + (size_t)(strip - 1) * 12 * VS : p.peer_in_l;
        unsigned long long *rout = strip > 0 ? p.ring + (size_t)(strip - 1) * 12 * VS : p.peer_out_l;
        int oi = 0, ii = 0;
        while (oi < H || ii < H) {
            uint32_t v[NR];
            if (oi < H && ev_test(fs_base + 16u, oi)) {                          // SW(oi) of the first column
                lds_words<NR>(sw_base + (uint32_t)(oi & 1) * SLOTB + VB, v);
                ll_send_u32<NR>(rout + (size_t)(8 + (oi & 3)) * VS, lane, tag0 + (uint32_t)(oi + 1), v);
                ++oi;
            }
            if (ii < H && ll_try_recv_u32<NR>(rin + (size_t)(ii & 3) * VS, lane, tag0 + (uint32_t)(ii + 1), v)) {   // E(ii)
                uint32_t u[NR];
                if (ii > 0) {                             // SE(ii - 1) was sent before E(ii): one more poll at most
                    while (!ll_try_recv_u32<NR>(rin + (size_t)(4 + ((ii - 1) & 3)) * VS, lane, tag0 + (uint32_t)ii, u)) {}
                    sts_words<NR>(se_base + (uint32_t)((ii - 1) & 3) * SLOTB, u);
                }
                sts_words<NR>(e_base + (uint32_t)(ii & 1) * SLOTB, v);
                ev_signal(fe_base, ii, lane);
                ++ii;
            }
        }
        return;
    }
    if (warp == K + 1) {                                  // right relay: E / SE out, SW in
        if (!has_right) return;
        unsigned long long *rout = strip + 1 < nstrips ? p.ring + (size_t)strip * 12 * VS : p.peer_out_r;
        const unsigned long long *rin = strip + 1 < nstrips ? p.ring + (size_t)strip * 12 * VS : p.peer_in_r;
        const uint32_t vc = (uint32_t)nce;                // mailbox column of the last image column of this strip
        int oi = 0, ii = 0;
        while (oi < H || ii < H) {
            uint32_t v[NR];
            if (oi < H && ev_test(fe_base + vc * 16u, oi)) {                     // E(oi) is out, and with it SE(oi - 1)
                if (oi > 0) {
                    lds_words<NR>(se_base + (uint32_t)((oi - 1) & 3) * SLOTB + vc * VB, v);
                    ll_send_u32<NR>(rout + (size_t)(4 + ((oi - 1) & 3)) * VS, lane, tag0 + (uint32_t)oi, v);
                }
                lds_words<NR>(e_base + (uint32_t)(oi & 1) * SLOTB + vc * VB, v);
                ll_send_u32<NR>(rout + (size_t)(oi & 3) * VS, lane, tag0 + (uint32_t)(oi + 1), v);
                ++oi;
            }
            if (ii < H && ll_try_recv_u32<NR>(rin + (size_t)(8 + (ii & 3)) * VS, lane, tag0 + (uint32_t)(ii + 1), v)) {   // SW(ii)
                sts_words<NR>(sw_base + (uint32_t)(ii & 1) * SLOTB + (vc + 1u) * VB, v);
                ev_signal(fs_base + (vc + 1u) * 16u, ii, lane);
                ++ii;
            }
        }
        return;
    }
    // ---- loader warp (first pass): descriptor rows of the strip, once per CTA row ----------------------------------------
    // Row r of the ring holds, for s = 0..3, copy_s[m] = right descriptor of column c0 + s + m (c0 = strip start + dmin
    // rounded down to a multiple of 4), and the K left descriptors.  Column x reads its window from the copy whose shift
    // makes x + dmin - c0 - s a multiple of 4: aligned vectors for every column.  Positions outside the descriptor row are
    // the same for every row: filled once with the "window leaves the image" flag, never copied.
    if (warp == K + 2) {
        if (!CENSUS) return;
        const int c0 = (x0 + p.dmin) & ~3;
        for (int sl = 0; sl < W1_RING; ++sl)
            for (int idx = lane; idx < CSLOT; idx += 32) {
                const int col = idx < 4 * CPW ? c0 + idx / CPW + idx % CPW : x0 + (idx - 4 * CPW);
                if (!((unsigned)col < (unsigned)p.pitch)) sts_u32(stg_base0 + (uint32_t)(sl * CSLOT + idx) * 4u, 0x80000000u);
            }
        __syncwarp();
        for (int r = 0; r < H; ++r) {
            if (r >= W1_RING) flag_wait_sleep(done_flag, (uint32_t)(r - W1_RING + 1));   // every column is past row r - RING
            const uint32_t dst = stg_base0 + (uint32_t)((r & (W1_RING - 1)) * CSLOT) * 4u;
            const uint32_t *rowR = p.descR + (size_t)r * p.pitch, *rowL = p.descL + (size_t)r * p.pitch;
#pragma unroll 4
            for (int idx = lane; idx < 4 * CPW; idx += 32) {
                const int col = c0 + idx / CPW + idx % CPW;
                if ((unsigned)col < (unsigned)p.pitch) cp_async_words<1>(dst + (uint32_t)idx * 4u, rowR + col);
            }
            if (lane < K && (unsigned)(x0 + lane) < (unsigned)p.pitch) cp_async_words<1>(dst + (uint32_t)(4 * CPW + lane) * 4u, rowL + x0 + lane);
            cp_async_commit();
            if (r >= W1_PF) {
                cp_async_wait<W1_PF>();
                flag_set(ld_flag, (uint32_t)(r - W1_PF + 1));
            }
        }
        cp_async_wait<0>();
        flag_set(ld_flag, (uint32_t)H);
        return;
    }

    // ---- compute warps: one column each ---------------------------------------------------------------------------------
    if (warp >= nce) return;
    const uint32_t vme = (uint32_t)(warp + 1);            // this column's mailbox column
    const int xl = x0 + warp;                             // logical column (pass 2: the frame is flipped in both axes)
    const int y0 = FINAL ? H - 1 : 0;
    const long row_stride = (FINAL ? -1L : 1L) * W * D;   // words (== floats)
    uint32_t *gpix = p.buf + ((size_t)y0 * W + (FINAL ? W - 1 - xl : xl)) * D;
    const int poff = p16_off<CB>(D) + lane * NR;
    size_t pixi = (size_t)y0 * W + (FINAL ? W - 1 - xl : xl);
    const uint32_t p1p1 = p.p1p1, p2p2 = p.p2p2;
    // first pass: where this column's window starts inside the descriptor ring rows
    const int cen_off = (xl + p.dmin) - ((x0 + p.dmin) & ~3);
    const uint32_t cen_win = stg_base0 + (uint32_t)((cen_off & 3) * CPW + (cen_off & ~3)) * 4u + lane_b;
    const uint32_t cen_left = stg_base0 + (uint32_t)(4 * CPW + warp) * 4u;
    // second pass: private staging ring [NSTG][K][32 * SIN], a pixel's block = [32][RW] cost words | [32][NR] partial sums
    const uint32_t stg_pix = (uint32_t)(32 * SIN) * 4u, stg_stage = (uint32_t)K * stg_pix;
    const uint32_t stg_me = stg_base0 + (uint32_t)warp * stg_pix;
    const uint32_t off0 = (uint32_t)(lane * RW) * 4u, off1 = (uint32_t)(32 * RW + lane * NR) * 4u;
    auto stage_row = [&](int r) {
        if (!FINAL || r >= H) return;
        const uint32_t sg = stg_me + (uint32_t)(r & (NSTG - 1)) * stg_stage;
        const uint32_t *src = gpix + (long)r * row_stride;
        cp_async_words<RW>(sg + off0, src + lane * RW);
        cp_async_words<NR>(sg + off1, src + poff);
    };
    if (FINAL) {
        for (int r = 0; r < PFD; ++r) {
            stage_row(r);
            cp_async_commit();
        }
    }
    uint32_t Sv[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) Sv[j] = 0u;
    const uint32_t nan2 = (p.inv | Tier<CB>::FLAG1) * 0x10001u;
    // an image border is a neighbour that never has to be waited for (its mailbox slots stay zero: flat states = path starts)
    const bool wait_l = has_left || warp > 0, wait_r = has_right || warp + 1 < nce, last_col = warp + 1 == nce;

#pragma unroll 1
    for (int i = 0; i < H; ++i) {
        const uint32_t par = (uint32_t)(i & 1) * SLOTB, q = (uint32_t)(i & 3) * SLOTB, qm = (uint32_t)((i - 1) & 3) * SLOTB;
        uint32_t *grow = gpix + (long)i * row_stride;
        // ---- 1: the pixel's cost codes (16 bits each, NaN flag in bit 15 / 7) and the partial sums so far ----------------
        uint32_t c16[NR], cc[NR], tot[NR];
        if (CENSUS) {
            flag_wait_sleep(ld_flag, (uint32_t)(i + 1));
            const uint32_t cg = (uint32_t)((i & (W1_RING - 1)) * CSLOT) * 4u;
            uint32_t lw[1], ra[NR], rb[NR];
            lds_words<1>(cen_left + cg, lw);
            lds_words<NR>(cen_win + cg, ra);
            lds_words<NR>(cen_win + cg + VB, rb);
            if (lw[0] >> 31) {                            // warp-uniform: the left window leaves the image
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
            st_cost<NR, CB>(grow, lane, c16);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] = 0u;
        } else {
            stage_row(i + PFD);
            cp_async_commit();
            cp_async_wait<PFD>();                          // this lane's copies of row i have landed
            const uint32_t sg = stg_me + (uint32_t)(i & (NSTG - 1)) * stg_stage;
            uint32_t craw[RW];
            lds_words<RW>(sg + off0, craw);
            lds_words<NR>(sg + off1, tot);
            unpack_cost<NR, CB>(craw, c16);
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) cc[j] = c16[j] & Tier<CB>::VALUES;

        uint32_t Lp[NR], L[NR];
        // ---- 2: SW, from the right neighbour's state of the previous row; published first (it travels against the wave) ----
        if (i > 0 && wait_r) ev_wait(fs_base + (vme + 1u) * 16u, i - 1);
        lds_words<NR>(sw_base + (SLOTB - par) + (vme + 1u) * VB, Lp);
        nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);
        sts_words<NR>(sw_base + par + vme * VB, L);
        ev_signal(fs_base + vme * 16u, i, lane);
#pragma unroll
        for (int j = 0; j < NR; ++j) tot[j] += L[j];
        // ---- 3: S, registers only -------------------------------------------------------------------------------------------
        nstep<NR>(cc, Sv, L, lane, p1p1, p2p2);
#pragma unroll
        for (int j = 0; j < NR; ++j) { Sv[j] = L[j]; tot[j] += L[j]; }
        // ---- 4: the E chain: wait, one step, publish -----------------------------------------------------------------------
        if (wait_l) ev_wait(fe_base + (vme - 1u) * 16u, i);
        lds_words<NR>(e_base + par + (vme - 1u) * VB, Lp);
        nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);
        sts_words<NR>(e_base + par + vme * VB, L);
        ev_signal(fe_base + vme * 16u, i, lane);
        if (CENSUS && last_col) sts_u32(done_flag, (uint32_t)(i + 1));            // the loader may reuse the ring row of i + 1 - RING
#pragma unroll
        for (int j = 0; j < NR; ++j) tot[j] += L[j];
        // ---- 5: SE (the left neighbour's SE of the previous row became visible with its E flag of this row) ----------------
        lds_words<NR>(se_base + qm + (vme - 1u) * VB, Lp);
        nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);
        sts_words<NR>(se_base + q + vme * VB, L);
#pragma unroll
        for (int j = 0; j < NR; ++j) tot[j] += L[j];
        // ---- 6: out --------------------------------------------------------------------------------------------------------
        if (!FINAL) {
            st_words<NR>(grow + poff, tot);
        } else {
            float fa[NR], fb[NR];
            uint32_t bl = 0xFFFFFFFFu, bh = 0xFFFFFFFFu;
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                uint32_t t = tot[j];
                if (p.overcounting) t = t - 7u * cc[j];   // S >= 8 C in every half: no borrow
                // 16-bit integer -> float32: PRMT builds 0x4B00'nnnn, one FADD removes the 2^23.  A NaN cell gets the upper
                // half 0x7F80 / 0x7FFF instead of 0x4B00: exponent all ones over a non-zero mantissa, i.e. a NaN that the
                // same FADD passes through -- no select.
                const uint32_t fl = c16[j] & Tier<CB>::FLAGS;
                const uint32_t sat = (CB == 1) ? fl * 0x1FFu : (fl >> 15) * 0xFFFFu;     // 0xFF80 / 0xFFFF per NaN half
                const uint32_t hx = 0x4B004B00u | (sat & 0x34FF34FFu);
                fa[j] = __uint_as_float(__byte_perm(t, hx, 0x5410)) - 8388608.0f;
                fb[j] = __uint_as_float(__byte_perm(t, hx, 0x7632)) - 8388608.0f;
                if (WTA) {
                    const uint32_t tk = t | sat;
                    bl = min(bl, __byte_perm(tk, (uint32_t)j, 0x1054));      // (low sum << 16) | j
                    bh = min(bh, __byte_perm(tk, (uint32_t)j, 0x3254));      // (high sum << 16) | j
                }
            }
            float *o = reinterpret_cast<float *>(grow) + lane * NR;
            st_floats<NR>(o, fa);
            st_floats<NR>(o + D / 2, fb);
            if (WTA) {
                uint32_t best = min(bl + (uint32_t)(lane * NR), bh + (uint32_t)(D / 2 + lane * NR));
                best = __reduce_min_sync(0xffffffffu, best);
                if (lane == 0) {
                    const size_t pix = pixi - (size_t)i * W;
                    const bool none = (best >> 16) >= ((CB == 1) ? 0xFF80u : 0xFFFFu);
                    p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(best & 0xFFFFu));
                    if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                }
            }
        }
    }
}

template <int NR, int CB>
int launch_wave1(NarrowParams p, int K, int nstrips, void *workspace, cudaStream_t s, bool *done) {
    *done = false;
    const bool wta = p.disp != nullptr;
    void (*w1)(const NarrowParams) = sgm_wave1_kernel<NR, CB, false, false, true>;
    void (*w2)(const NarrowParams) = wta ? sgm_wave1_kernel<NR, CB, true, true, false> : sgm_wave1_kernel<NR, CB, true, false, false>;
    const int threads = (K + 3) * 32;
    const size_t fixed = ((size_t)8 * (K + 2) * NR * 32 + W1_FLAG_WORDS) * sizeof(uint32_t);
    const size_t smem1 = fixed + (size_t)W1_RING * (4 * W1Cpw<NR>::value + 32) * sizeof(uint32_t);
    const size_t smem2 = fixed + (size_t)4 * K * 32 * (NR * CB / 2 + NR) * sizeof(uint32_t);
    if (smem1 > 227 * 1024 || smem2 > 227 * 1024) return PB200_OK;
    int occ1 = 0, occ2 = 0;
    PB200_CUDA(cudaFuncSetAttribute((const void *)w1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    PB200_CUDA(cudaFuncSetAttribute((const void *)w2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, (const void *)w1, threads, smem1));
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, (const void *)w2, threads, smem2));
    const int nsm = sm_count();
    if ((long)occ1 * nsm < nstrips || (long)occ2 * nsm < nstrips) return PB200_OK;       // every strip must be resident
    p.K = K;
    p.ring = reinterpret_cast<unsigned long long *>(workspace);
    const size_t wring = (size_t)nstrips * 12 * NR * 32 * sizeof(unsigned long long);
    if ((size_t)(reinterpret_cast<char *>(p.flag) - reinterpret_cast<char *>(workspace)) < wring) return PB200_OK;   // ring must end before the flag
    void *args[] = {(void *)&p};
    PB200_CUDA(cudaMemsetAsync(p.flag, 0, sizeof(int), s));
    PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
    PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w1, dim3(nstrips), dim3(threads), args, smem1, s));
    PB200_LAUNCH_CHECK("sgm_wave1_kernel<down, census>");
    PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
    // pass 2 runs on the flipped frame: its logical left is the physical right
    NarrowParams p2 = p;
    p2.peer_in_l = p.peer_in_r; p2.peer_out_l = p.peer_out_r; p2.peer_in_r = p.peer_in_l; p2.peer_out_r = p.peer_out_l;
    if (p2.peer_in_l) p2.peer_in_l += 12 * NR * 32;       // second block of every edge buffer = pass 2
    if (p2.peer_out_l) p2.peer_out_l += 12 * NR * 32;
    if (p2.peer_in_r) p2.peer_in_r += 12 * NR * 32;
    if (p2.peer_out_r) p2.peer_out_r += 12 * NR * 32;
    void *args2[] = {(void *)&p2};
    PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w2, dim3(nstrips), dim3(threads), args2, smem2, s));
    PB200_LAUNCH_CHECK("sgm_wave1_kernel<up>");
    note_path(STAGE_SGM, PATH_SGM_WAVE1_CENSUS, NR * 10 + CB);
    *done = true;
    return PB200_OK;
}

}  // namespace

// One-column wavefront for the fused Census -> SGM stage.  `peer` (optional): the four edge buffers of a column-tiled
// multi-GPU run in the order {in_left, out_left, in_right, out_right} (physical sides) and the epoch of this call.
int sgm_census_wave1_launch(NarrowParams p, int NR, bool bytes, void *workspace, cudaStream_t s, bool *done) {
    *done = false;
    const int nsm = sm_count();
    int K = ceil_div(p.W, nsm);
    if (K < 4) K = 4;
    if (K > W1_MAXK) return PB200_OK;
    const int nstrips = ceil_div(p.W, K);
    if (NR == 4) return bytes ? launch_wave1<4, 1>(p, K, nstrips, workspace, s, done) : launch_wave1<4, 2>(p, K, nstrips, workspace, s, done);
    if (NR == 2) return bytes ? launch_wave1<2, 1>(p, K, nstrips, workspace, s, done) : launch_wave1<2, 2>(p, K, nstrips, workspace, s, done);
    return launch_wave1<1, 2>(p, K, nstrips, workspace, s, done);
}

}  // namespace pb200
