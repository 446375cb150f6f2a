// sgm_wave1.cu -- the SKEWED wavefront: the two 4-direction SGM passes with one column per warp and no counter-flow.
//
// Same arithmetic, same two passes and the same intermediate format as sgm_wave_kernel (sgm_narrow.cu): pass 1 walks the
// image top-down and runs E, SE, S, SW; pass 2 walks it bottom-up on the frame flipped in both axes (W, NW, N, NE), adds
// its four directions to the 16-bit partial sums of pass 1 and emits float32 S (+ NaN restore, overcounting, WTA) in
// place.  What changes is the geometry of the wave.
//
// In the straight frame a pixel (y, k) takes E from (y, k-1), SE from (y-1, k-1), S from (y-1, k) -- all from the left
// or from itself -- but SW from (y-1, k+1): against the direction the wave travels.  That one counter-flow edge closes a
// cycle between every pair of neighbouring columns (E to the right, SW back to the left) which pins them within one row
// of each other, so the hand-over latency of the SLOWEST link -- a strip boundary through L2, or a GPU boundary through
// NVLink -- bounds the row rate of the whole image (measured: 3.85 us per row with 1.3 k-cycle strip boundaries against
// 2.3 us of issue time).  Here a warp owns a SHEARED column c = (k + y) mod W instead: it steps one image column to the
// left with every row.  In that frame
//      E  (y, k-1)   -> column c-1, this row          S  (y-1, k)   -> column c-1, previous row
//      SE (y-1, k-1) -> column c-2, previous row      SW (y-1, k+1) -> column c itself (registers)
// every edge points to the right.  The pipeline is one-directional and cyclic (column 0 follows column W-1, the row-y
// chain starts at column y mod W, where k = 0 and the left inputs are path starts): neighbours are coupled only by the
// depth of their mailboxes, a slow link costs latency once (fill) instead of bandwidth, and the same kernel runs
// COLUMN-TILED over several GPUs -- the relay warp of a tile's last strip stores straight into the next GPU's memory
// (NVLink peer stores, {tag, value} words) and the first strip's relay polls local memory: one wave across all GPUs,
// no host-side step, no collective.
//
// CTA = one strip of K <= 28 sheared columns: K compute warps (<= 64 registers: the SW state, the pixel's cost, the
// running total), an in-relay (ring -> mailbox columns 0 / 1) and an out-relay (last two columns -> ring).  Every warp
// stages its own inputs three rows ahead with cp.async into a private ring: in the first pass the D right census
// descriptors of its pixel as two aligned 16-byte copies per lane (the descriptors come in four word-shifted, padded
// copies, census.cu), in the second pass the packed costs and partial sums.  A strip-wide descriptor ring does not work
// here: behind the image seam the columns of a strip legitimately sit up to K rows apart.
// Hand-over events are mbarrier phases (a waiting warp sleeps in hardware instead of spinning through issue slots).
//
// Row program of sheared column v (mailbox column; 0 / 1 = the previous strip's last two columns), row y, pixel k:
//   1  cost c(y);  SW(y) = step(c, SW(y-1))  [registers; flat when k = W-1]
//   2  wait E-event(v-1, y);  load E_in = e[y&1][v-1], S_in = s[(y-1)&1][v-1], SE_in = se[(y-1)&1][v-2]
//   3  E(y) = step(c, E_in) -> e[y&1][v], E-event(v, y), prog[v] = y+1     (slot free once prog[v+2] >= y)
//   4  S(y), SE(y) -> s[y&1][v], se[y&1][v]   (visible to the right with the next E-event)
//   5  total = [P16 +] SW + S + E + SE -> global
// k = 0 (chain start, once per W rows): no wait, E_in = SE_in = flat, and steps 3 / 4 swap so that S and SE are stored
// before the E-event: the right neighbour starts the NEXT row's chain and has no later event of ours to acquire.
#include <type_traits>

#include "sgm_packed.cuh"

namespace pb200 {

namespace {

constexpr int W1_MAXK = 28;                        // sheared columns (compute warps) per strip
constexpr int W1_THREADS = (W1_MAXK + 2) * 32;     // + in-relay, out-relay
constexpr int W1_NRG = 8;                          // rows in flight across a strip / GPU boundary
constexpr int W1_FLAG_WORDS = 256;                 // events E[32][2] (512 B) | RD[2][2] | prog[36] | rdc

template <int NR> struct W1Geo {
    static constexpr int VS = NR * 32;
    static constexpr int CIN = 2 * VS + 4;                     // first pass, staged words per pixel: D right descriptors + the left one
    static constexpr size_t BLK = (size_t)W1_NRG * 4 * VS + 16;   // 64-bit words per boundary block: data | ack
};

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
__device__ __forceinline__ void flag_set(uint32_t addr, uint32_t v) {      // release: the warp barrier orders the other lanes' stores
    __syncwarp();
    asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// counters that are read now and then (ring rows, slot credits): wait without occupying the working warps' issue slots
__device__ __forceinline__ void flag_wait_sleep(uint32_t addr, uint32_t target) {
    while ((int)(flag_peek(addr) - target) < 0) __nanosleep(64);
}

// Hand-over events are mbarrier PHASES: a consumer that is early is suspended by the hardware (mbarrier.try_wait) instead
// of burning issue slots in an LDS / compare / branch loop (44 % of all issued instructions in the two-column kernel,
// profiles/r1_ncu_fused_pass1.txt).  Every event stream owns TWO barriers used by even and odd rows in turn; row i
// completes phase i >> 1 of barrier i & 1.  The slot credits (prog) never let a producer run two rows ahead of its
// consumer's wait, so a barrier is at most one completed phase ahead of the phase a waiter asks for, which is what a
// parity wait can tell apart.
__device__ __forceinline__ void ev_wait(uint32_t bar, int i) {
    const uint32_t addr = bar + (uint32_t)(i & 1) * 8u, parity = (uint32_t)(i >> 1) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1_WAIT:\n"
        // (no suspend-time hint: with one, ptxas wraps the wait in NANOSLEEP.SYNCS, which every arrive on ANY barrier of the
        // SM wakes up -- 50 re-checks per row and waiting warp, profiles/r2_ncu_wave1_hint.txt; the plain form sleeps in hardware)
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra W1_DONE;\n"
        "bra W1_WAIT;\n"
        "W1_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
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

// the same with the barrier address and the phase parity resolved by the caller (compile-time constants in the unrolled row loop)
__device__ __forceinline__ void ev_wait_c(uint32_t addr, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "W1C_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra W1C_DONE;\n"
        "bra W1C_WAIT;\n"
        "W1C_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void ev_signal_c(uint32_t addr, int lane) {
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
    using G = W1Geo<NR>;
    constexpr int VS = G::VS;
    constexpr int RW = NR * CB / 2;                      // raw cost words per lane
    constexpr int NSTG = 4, PFD = 3;                     // input staging (per warp, cp.async): rows in the ring / in flight
    constexpr int SIN = CENSUS ? G::CIN : 32 * (RW + NR);   // staged words per pixel
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int K = p.K;
    constexpr int NVM = W1_MAXK + 2;                     // mailbox columns: 0 / 1 = previous strip's last two columns, 2 .. K+1 = ours
    // (the geometry of the shared arrays is that of the widest strip whatever K is: every mailbox / staging address of a warp
    // is then ONE base register plus an immediate)
    const int strip = blockIdx.x, nstrips = gridDim.x;
    const int H = p.H, D = p.D;
    const int HT = p.nimg * H;                           // rows of the whole batch: the images follow each other in ONE wave
    const int Wt = p.W, Wg = p.Wg;                       // local (tile) width = storage pitch, global width of the sheared ring
    const int x0 = strip * K;
    const int nce = min(K, Wt - x0);                     // sheared columns of this strip
    // shared: e[2][NVM][VS] | s[2][NVM][VS] | se[2][NVM][VS] | events, counters | staging
    constexpr int mbox_words = 6 * NVM * VS;
    constexpr int stage_words = NSTG * W1_MAXK * SIN;
    for (int i = threadIdx.x; i < mbox_words + W1_FLAG_WORDS + stage_words; i += blockDim.x) w1_smem[i] = 0u;
    __syncthreads();
    if (threadIdx.x < 68)                                                        // E[32][2] | RD[2][2], one arrival per phase
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(w1_smem + mbox_words) + threadIdx.x * 8u) : "memory");
    __syncthreads();
    const uint32_t lane_b = (uint32_t)(lane * NR) * 4u;
    constexpr uint32_t VB = (uint32_t)VS * 4u, SLOTB = (uint32_t)(NVM * VS) * 4u; // bytes per vector / per mailbox slot
    const uint32_t e_base = smem_u32(w1_smem) + lane_b;
    const uint32_t s_base = e_base + 2u * SLOTB;
    const uint32_t se_base = s_base + 2u * SLOTB;
    const uint32_t ev_base = smem_u32(w1_smem) + (uint32_t)mbox_words * 4u;       // E-event of mailbox column v: ev_base + 16 v
    const uint32_t rd_base = ev_base + 512u;                                      // RD-event of the last (0) / second-to-last (1) column
    const uint32_t prog_base = ev_base + 544u;                                    // prog[v], v = 0 .. K+3 (K+2 / K+3 = the out-relay)
    const uint32_t rdc_flag = ev_base + 704u;
    const uint32_t stg_base0 = ev_base + (uint32_t)W1_FLAG_WORDS * 4u;
    const uint32_t tag0 = p.tag_base;
    const uint32_t v_last = (uint32_t)(nce + 1), v_prev = (uint32_t)nce;          // mailbox columns of this strip's last two sheared columns

    // ---- in-relay: the left boundary's ring block -> mailbox columns 1 (E, S, SE of the previous column) and 0 (its left
    // neighbour's SE).  Block: W1_NRG rows x {E, S, SE, SE2} vectors of VS 64-bit {tag, value} words, then the ack word.
    if (warp == K) {
        const bool peer = strip == 0 && p.peer_in_l != nullptr;
        const size_t bl = (size_t)(strip == 0 ? nstrips - 1 : strip - 1) * G::BLK;
        const unsigned long long *rin = peer ? p.peer_in_l : p.ring + bl;
        unsigned long long *ack = peer ? p.peer_out_l : p.ring + bl + (size_t)W1_NRG * 4 * VS;
        for (int y = 0; y < HT; ++y) {
            uint32_t v[NR];
            if (y >= 1) flag_wait_sleep(prog_base + 3u * 4u, (uint32_t)(y - 1));     // columns 2 / 3 have loaded their inputs of row y - 2
            if (y > 0) {                                  // B(y-1): S, SE of the previous column and SE of the one before, row y - 1
                const unsigned long long *src = rin + (size_t)((y - 1) & (W1_NRG - 1)) * 4 * VS;
                const uint32_t pb = (uint32_t)((y - 1) & 1) * SLOTB;
                while (!ll_try_recv_u32<NR>(src + 1 * VS, lane, tag0 + (uint32_t)y, v)) __nanosleep(PB200_RELAY_BACKOFF_NS);
                sts_words<NR>(s_base + pb + VB, v);
                while (!ll_try_recv_u32<NR>(src + 2 * VS, lane, tag0 + (uint32_t)y, v)) __nanosleep(PB200_RELAY_BACKOFF_NS);
                sts_words<NR>(se_base + pb + VB, v);
                flag_set(rdc_flag, (uint32_t)y);          // the chain start of row y (column 2 when its k = 0) waits for this
                // SE of the column before the previous one: when that column closed the chain of row y - 1 it arrives a whole
                // traversal later -- nobody needs it before E(y), which it always precedes
                while (!ll_try_recv_u32<NR>(src + 3 * VS, lane, tag0 + (uint32_t)y, v)) __nanosleep(PB200_RELAY_BACKOFF_NS);
                sts_words<NR>(se_base + pb, v);
            }
            const unsigned long long *src = rin + (size_t)(y & (W1_NRG - 1)) * 4 * VS;
            while (!ll_try_recv_u32<NR>(src, lane, tag0 + (uint32_t)(y + 1), v)) __nanosleep(PB200_RELAY_BACKOFF_NS);      // A(y): E of the previous column
            sts_words<NR>(e_base + (uint32_t)(y & 1) * SLOTB + VB, v);
            ev_signal(ev_base + 16u, y, lane);
            if (lane == 0) asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(ack), "l"((unsigned long long)(tag0 + (uint32_t)(y + 1))) : "memory");
        }
        return;
    }
    // ---- out-relay: this strip's last two columns -> the right boundary's ring block (the next strip's, or the next GPU's)
    if (warp == K + 1) {
        const bool peer = strip + 1 == nstrips && p.peer_out_r != nullptr;
        const size_t br = (size_t)strip * G::BLK;
        unsigned long long *rout = peer ? p.peer_out_r : p.ring + br;
        const unsigned long long *ack = peer ? p.peer_in_r : p.ring + br + (size_t)W1_NRG * 4 * VS;
        uint32_t acked = tag0;                            // rows the receiver has consumed (tag space)
        if (peer && p.prev_ack != 0u)                     // the neighbour still reads the previous image's last rows from this link
            while ((uint32_t)ll_load(ack) != p.prev_ack) __nanosleep(500);
        for (int y = 0; y < HT; ++y) {
            uint32_t v[NR];
            if (y >= W1_NRG) {                            // ring credit: row y - NRG has left the block
                const uint32_t need = tag0 + (uint32_t)(y - W1_NRG + 1);
                while ((int)(acked - need) < 0) {
                    const uint32_t a = (uint32_t)ll_load(ack);
                    if ((int)(a - tag0) >= 0 && (int)(a - tag0) <= HT) acked = a;     // words of an older epoch are not credits
                    if ((int)(acked - need) < 0) __nanosleep(200);
                }
            }
            if (y > 0) {
                // S, SE of the last column as soon as IT has finished row y - 1: the next strip's first column may be the
                // start of chain y and then needs nothing else.  The SE of the column before it follows when that column has
                // finished -- a whole traversal later when it closed chain y - 1 -- and always before E(y).
                unsigned long long *dst = rout + (size_t)((y - 1) & (W1_NRG - 1)) * 4 * VS;
                const uint32_t pb = (uint32_t)((y - 1) & 1) * SLOTB;
                ev_wait(rd_base, y - 1);
                lds_words<NR>(s_base + pb + v_last * VB, v);
                ll_send_u32<NR>(dst + 1 * VS, lane, tag0 + (uint32_t)y, v);
                lds_words<NR>(se_base + pb + v_last * VB, v);
                ll_send_u32<NR>(dst + 2 * VS, lane, tag0 + (uint32_t)y, v);
                if (nce >= 2) ev_wait(rd_base + 16u, y - 1);
                lds_words<NR>(se_base + pb + v_prev * VB, v);
                ll_send_u32<NR>(dst + 3 * VS, lane, tag0 + (uint32_t)y, v);
            }
            ev_wait(ev_base + v_last * 16u, y);           // A(y)
            lds_words<NR>(e_base + (uint32_t)(y & 1) * SLOTB + v_last * VB, v);
            ll_send_u32<NR>(rout + (size_t)(y & (W1_NRG - 1)) * 4 * VS, lane, tag0 + (uint32_t)(y + 1), v);
            __syncwarp();
            if (lane < 2) sts_u32(prog_base + (v_last + 1u + (uint32_t)lane) * 4u, (uint32_t)(y + 1));   // slot credits of the last two columns
        }
        return;
    }
    // ---- compute warps: one sheared column each -----------------------------------------------------------------------------
    if (warp >= nce) return;
    const uint32_t vme = (uint32_t)(warp + 2);            // this column's mailbox column
    const uint32_t mb = e_base + vme * VB;                // its e[0] vector (this lane's words); every other mailbox vector = mb + immediate
    const uint32_t evme = ev_base + vme * 16u, progme = prog_base + vme * 4u;
    int k = (p.c_off + x0 + warp) % Wg;                   // image column (travel frame) of this sheared column in row 0
    const int poff = p16_off<CB>(D) + lane * NR;
    const uint32_t p1p1 = p.p1p1, p2p2 = p.p2p2;
    const bool is_last = warp + 1 == nce, is_prev = warp + 2 == nce;
    // Storage: by image coordinates (one GPU: the (H, W, D) volume of the C-ABI) or by sheared column (a tile of a multi-GPU
    // run).  The pixel index moves by a constant from row to row -- also from the last row of an image of the batch to the
    // first row of the next one (the second pass walks rows AND images backwards) -- plus a fix-up where the image column wraps.
    const int col0 = p.sheared_store ? (FINAL ? Wt - 1 - (x0 + warp) : x0 + warp) : (FINAL ? Wg - 1 - k : k);
    uint32_t pix = (uint32_t)((FINAL ? HT - 1 : 0) * Wt + col0);
    const int pstep = (FINAL ? -Wt : Wt) + (p.sheared_store ? 0 : (FINAL ? 1 : -1));
    const int pfix = p.sheared_store ? 0 : (FINAL ? -Wg : Wg);
    // private staging ring [NSTG][K][SIN].  First pass: a pixel's block = its D right descriptors (lane-major) | the left one;
    // second pass: [32][RW] cost words | [32][NR] partial sums.
    constexpr uint32_t stg_pix = (uint32_t)SIN * 4u, stg_stage = (uint32_t)W1_MAXK * stg_pix;
    const uint32_t stg_me = stg_base0 + (uint32_t)warp * stg_pix;
    const uint32_t off0 = (uint32_t)(lane * RW) * 4u, off1 = (uint32_t)(32 * RW + lane * NR) * 4u;
    // prefetch cursors (PFD rows ahead of the row being computed); the descriptor rows of a batch are contiguous:
    // row r of the batch = [r][copy][pitch4] / [r][pitch]
    int kpf = k;
    uint32_t pfpix = pix;
    const int lane_w = lane * NR + p.padl;
    auto stage_row = [&](uint32_t sg, int r) {
        if (r < HT) {
            if (CENSUS) {
                // the window [kpf + dmin, kpf + dmin + D) starts 16-byte aligned in the copy whose shift is (kpf + dmin) & 3
                const int ws = kpf + p.dmin, sh = ws & 3;
                const uint32_t *src = p.descR4 + (size_t)(uint32_t)(4 * r + sh) * (uint32_t)p.pitch4 + (ws - sh + lane_w);
                cp_async_words<NR>(sg + lane_b, src);
                cp_async_words<NR>(sg + VB + lane_b, src + VS);
                if (lane == 0) cp_async_words<1>(sg + 2u * VB, p.descL + (size_t)(uint32_t)r * (uint32_t)p.pitch + kpf);
            } else {
                const uint32_t *src = p.buf + (size_t)pfpix * D;
                cp_async_words<RW>(sg + off0, src + lane * RW);
                cp_async_words<NR>(sg + off1, src + poff);
            }
        }
        if (!CENSUS) pfpix += (uint32_t)pstep;
        if (--kpf < 0) { kpf = Wg - 1; if (!CENSUS) pfpix += (uint32_t)pfix; }
    };
    stage_row(stg_me, 0);
    cp_async_commit();
    stage_row(stg_me + stg_stage, 1);
    cp_async_commit();
    stage_row(stg_me + 2u * stg_stage, 2);
    cp_async_commit();
    static_assert(PFD == 3 && NSTG == 4, "the row loop below hard-wires the rotation of the four staging slots");
    // the row loop is unrolled by two: the mailbox slot and the event barrier of a row (y & 1) are immediates; what changes
    // every second row -- the phase parity of the events and the half of the staging ring in use -- lives in three registers
    uint32_t ph = 0u, sg_cur = stg_me, sg_oth = stg_me + 2u * stg_stage;
    uint32_t SW[NR];
#pragma unroll
    for (int j = 0; j < NR; ++j) SW[j] = 0u;
    const uint32_t nan2 = (p.inv | Tier<CB>::FLAG1) * 0x10001u;
    int yl = 0;                                           // row inside the current image of the batch
    int y = 0;                                            // row of the batch

    auto row = [&](auto uc) {
        constexpr int U = decltype(uc)::value;            // y & 1
        constexpr uint32_t par = (uint32_t)U * SLOTB, prv = SLOTB - par;
        constexpr uint32_t EV = (uint32_t)U * 8u;
        const uint32_t PH = ph;
        constexpr uint32_t S_OFF = 2u * SLOTB, SE_OFF = 4u * SLOTB;       // the s / se streams behind the e stream
        uint32_t *grow = p.buf + (size_t)pix * D;
        // ---- the pixel's cost codes (16 bits each, NaN flag in bit 15 / 7) and the partial sums so far -------------------------
        uint32_t c16[NR], cc[NR], tot[NR];
        // row y sits in slot y & 3 = 2 * half + U; row y + 3 goes to slot (y + 3) & 3: the other half's slot 1 (U = 0) or this half's slot 0
        stage_row(U == 0 ? sg_oth + stg_stage : sg_cur, y + PFD);
        cp_async_commit();
        cp_async_wait<PFD>();                              // this lane's copies of row y have landed
        const uint32_t sg = sg_cur + (uint32_t)U * stg_stage;
        if (CENSUS) {
            __syncwarp();                                  // the left descriptor was copied by lane 0
            uint32_t lw[1], ra[NR], rb[NR];
            lds_words<1>(sg + 2u * VB, lw);
            lds_words<NR>(sg + lane_b, ra);
            lds_words<NR>(sg + VB + lane_b, rb);
            if (lw[0] >> 31) {                            // warp-uniform: the left window leaves the image
#pragma unroll
                for (int j = 0; j < NR; ++j) c16[j] = nan2;
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const uint32_t xlo = lw[0] ^ ra[j], xhi = lw[0] ^ rb[j];          // bit 31 = the right window leaves the image
                    uint32_t pk, fl;
                    asm("mad.lo.u32 %0, %1, 65536, %2;" : "=r"(pk) : "r"(__popc(xhi)), "r"(__popc(xlo)));   // both counts in one word
                    asm("prmt.b32 %0, %1, %2, 0xFFBB;" : "=r"(fl) : "r"(xlo), "r"(xhi));   // sign-replicated top bytes: 0xFFFF per flagged half
                    c16[j] = (pk & ~fl) | (nan2 & fl);
                }
            }
            st_cost<NR, CB>(grow, lane, c16);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] = 0u;
        } else {
            uint32_t craw[RW];
            lds_words<RW>(sg + off0, craw);
            lds_words<NR>(sg + off1, tot);
            unpack_cost<NR, CB>(craw, c16);
        }
#pragma unroll
        for (int j = 0; j < NR; ++j) cc[j] = c16[j] & Tier<CB>::VALUES;

        uint32_t Lp[NR], Lq[NR], L[NR];
        const bool first = yl == 0;                       // first row of an image of the batch: every path starts
        // ---- SW: this column's own state of the previous row (the pixel up-right); a path start at the right image border --------
        if (k == Wg - 1 || first) {
#pragma unroll
            for (int j = 0; j < NR; ++j) SW[j] = 0u;
        }
        nstep<NR>(cc, SW, L, lane, p1p1, p2p2);
#pragma unroll
        for (int j = 0; j < NR; ++j) { SW[j] = L[j]; tot[j] += L[j]; }
        // slot credit: e / s / se[y & 1] of this column still hold row y - 2 until column v + 2 has loaded its inputs of row y - 1
        flag_wait_sleep(progme + 8u, (uint32_t)y);
        if (k != 0) {
            // ---- the E chain: wait for the left neighbour, one step, publish; S and SE behind it -----------------------------------
            ev_wait_c(evme - 16u + EV, PH);
            lds_words<NR>(mb - VB + par, Lp);
            lds_words<NR>(mb - VB + S_OFF + prv, Lq);
            if (first) {
#pragma unroll
                for (int j = 0; j < NR; ++j) Lq[j] = 0u;
            }
            nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);
            sts_words<NR>(mb + par, L);
            ev_signal_c(evme + EV, lane);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] += L[j];
            lds_words<NR>(mb - 2u * VB + SE_OFF + prv, Lp);
            if (first) {
#pragma unroll
                for (int j = 0; j < NR; ++j) Lp[j] = 0u;
            }
            nstep<NR>(cc, Lq, L, lane, p1p1, p2p2);                       // S
            sts_words<NR>(mb + S_OFF + par, L);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] += L[j];
            nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);                       // SE
            sts_words<NR>(mb + SE_OFF + par, L);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] += L[j];
        } else {
            // ---- chain start (left image border): E and SE are path starts, S comes from the left neighbour's previous row, which
            // that neighbour finished as a chain start itself; everything is stored BEFORE the E-event (see the file header) ---------
            if (!first) {
                if (vme == 2u) flag_wait_sleep(rdc_flag, (uint32_t)y);    // across the strip boundary: the in-relay has delivered row y - 1
                lds_words<NR>(mb - VB + S_OFF + prv, Lq);
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j) Lq[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < NR; ++j) Lp[j] = 0u;
            nstep<NR>(cc, Lq, L, lane, p1p1, p2p2);                       // S
            sts_words<NR>(mb + S_OFF + par, L);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] += L[j];
            nstep<NR>(cc, Lp, L, lane, p1p1, p2p2);                       // E = SE = a path start: c itself
            sts_words<NR>(mb + SE_OFF + par, L);
            sts_words<NR>(mb + par, L);
            ev_signal_c(evme + EV, lane);
#pragma unroll
            for (int j = 0; j < NR; ++j) tot[j] += 2u * L[j];
        }
        sts_u32(progme, (uint32_t)(y + 1));                 // inputs of row y are in registers: credit for column v - 2
        if (is_last) ev_signal_c(rd_base + EV, lane);                      // the out-relay forwards S / SE of row y
        if (is_prev) ev_signal_c(rd_base + 16u + EV, lane);
        // ---- out ------------------------------------------------------------------------------------------------------------------
        if (!FINAL) {
            st_words<NR>(grow + poff, tot);
        } else {
            uint32_t anyflag = 0u;
#pragma unroll
            for (int j = 0; j < NR; ++j) anyflag |= c16[j];
            const bool nans = __any_sync(0xffffffffu, (anyflag & Tier<CB>::FLAGS) != 0u);   // rare: pixels near the image border
            if (p.overcounting) {
#pragma unroll
                for (int j = 0; j < NR; ++j) tot[j] -= 7u * cc[j];        // S >= 8 C in every half: no borrow
            }
            float *o = reinterpret_cast<float *>(grow) + lane * NR;
            uint32_t best;
            if (!nans) {
                // 16-bit integer -> float32: PRMT builds 0x4B00'nnnn, one FADD removes the 2^23
                {
                    float f[NR];
#pragma unroll
                    for (int j = 0; j < NR; ++j) f[j] = __uint_as_float(__byte_perm(tot[j], 0x4B004B00u, 0x5410)) - 8388608.0f;
                    st_floats<NR>(o, f);
#pragma unroll
                    for (int j = 0; j < NR; ++j) f[j] = __uint_as_float(__byte_perm(tot[j], 0x4B004B00u, 0x7632)) - 8388608.0f;
                    st_floats<NR>(o + D / 2, f);
                }
                if (WTA) {
                    if (CB == 1 && NR == 4) {
                        // byte tier: sums < 2^10, so (sum << 2 | register index) fits a half: two packed min-ops find, per half,
                        // the smallest sum and the FIRST register that holds it (the keys are built on the FMA pipe)
#pragma unroll
                        for (int j = 0; j < NR; ++j) asm("mad.lo.u32 %0, %1, 4, %2;" : "=r"(tot[j]) : "r"(tot[j]), "r"((uint32_t)j * 0x10001u));
                        const uint32_t m = __vminu2(__vimin3_u16x2(tot[0], tot[1], tot[2]), tot[3]);
                        const uint32_t lo = m & 0xFFFFu, hi = m >> 16;
                        best = min(((lo >> 2) << 16) + (lo & 3u) + (uint32_t)(lane * NR),
                                   ((hi >> 2) << 16) + (hi & 3u) + (uint32_t)(D / 2 + lane * NR));
                    } else {
                        uint32_t bl = 0xFFFFFFFFu, bh = 0xFFFFFFFFu;
#pragma unroll
                        for (int j = 0; j < NR; ++j) {
                            bl = min(bl, __byte_perm(tot[j], (uint32_t)j, 0x1054));      // (low sum << 16) | j
                            bh = min(bh, __byte_perm(tot[j], (uint32_t)j, 0x3254));      // (high sum << 16) | j
                        }
                        best = min(bl + (uint32_t)(lane * NR), bh + (uint32_t)(D / 2 + lane * NR));
                    }
                }
            } else {
                // A NaN cell gets the upper half 0x7F80 / 0x7FFF instead of 0x4B00: exponent all ones over a non-zero mantissa,
                // i.e. a NaN that the same FADD passes through -- no select.  In the WTA it saturates its half (no valid sum
                // reaches the sentinel).
                uint32_t bl = 0xFFFFFFFFu, bh = 0xFFFFFFFFu;
                float fa[NR], fb[NR];
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const uint32_t fl = c16[j] & Tier<CB>::FLAGS;
                    const uint32_t sat = (CB == 1) ? fl * 0x1FFu : (fl >> 15) * 0xFFFFu;     // 0xFF80 / 0xFFFF per NaN half
                    const uint32_t hx = 0x4B004B00u | (sat & 0x34FF34FFu);
                    fa[j] = __uint_as_float(__byte_perm(tot[j], hx, 0x5410)) - 8388608.0f;
                    fb[j] = __uint_as_float(__byte_perm(tot[j], hx, 0x7632)) - 8388608.0f;
                    const uint32_t tk = tot[j] | sat;
                    bl = min(bl, __byte_perm(tk, (uint32_t)j, 0x1054));
                    bh = min(bh, __byte_perm(tk, (uint32_t)j, 0x3254));
                }
                st_floats<NR>(o, fa);
                st_floats<NR>(o + D / 2, fb);
                best = min(bl + (uint32_t)(lane * NR), bh + (uint32_t)(D / 2 + lane * NR));
            }
            if (WTA) {
                best = __reduce_min_sync(0xffffffffu, best);
                if (lane == 0) {
                    const bool none = (best >> 16) >= ((CB == 1) ? 0xFF80u : 0xFFFFu);
                    p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(best & 0xFFFFu));
                    if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                }
            }
        }
        pix += (uint32_t)pstep;
        if (--k < 0) { k = Wg - 1; pix += (uint32_t)pfix; }
        if (++yl == H) yl = 0;
        ++y;
    };
#pragma unroll 1
    while (y + 2 <= HT) {
        row(std::integral_constant<int, 0>{});
        row(std::integral_constant<int, 1>{});
        ph ^= 1u;
        const uint32_t t = sg_cur;
        sg_cur = sg_oth;
        sg_oth = t;
    }
    if (y < HT) row(std::integral_constant<int, 0>{});
}

template <int NR, int CB>
int launch_wave1(NarrowParams p, int K, int nstrips, void *workspace, size_t ring_room, const Wave1Peers *peers, cudaStream_t s, bool *done) {
    *done = false;
    using G = W1Geo<NR>;
    const bool wta = p.disp != nullptr;
    void (*w1)(const NarrowParams) = sgm_wave1_kernel<NR, CB, false, false, true>;
    void (*w2)(const NarrowParams) = wta ? sgm_wave1_kernel<NR, CB, true, true, false> : sgm_wave1_kernel<NR, CB, true, false, false>;
    const int threads = (K + 2) * 32;
    const size_t fixed = ((size_t)6 * (W1_MAXK + 2) * NR * 32 + W1_FLAG_WORDS) * sizeof(uint32_t);
    const size_t smem1 = fixed + (size_t)4 * W1_MAXK * G::CIN * sizeof(uint32_t);
    const size_t smem2 = fixed + (size_t)4 * W1_MAXK * 32 * (NR * CB / 2 + NR) * sizeof(uint32_t);
    if (smem1 > 227 * 1024 || smem2 > 227 * 1024) return PB200_OK;
    const size_t wring = (size_t)nstrips * G::BLK * sizeof(unsigned long long);
    if (wring > ring_room) return PB200_OK;                                      // the boundary blocks must end before the flag
    int occ1 = 0, occ2 = 0;
    PB200_CUDA(cudaFuncSetAttribute((const void *)w1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    PB200_CUDA(cudaFuncSetAttribute((const void *)w2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, (const void *)w1, threads, smem1));
    PB200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ2, (const void *)w2, threads, smem2));
    const int nsm = sm_count();
    if ((long)occ1 * nsm < nstrips || (long)occ2 * nsm < nstrips) return PB200_OK;       // every strip must be resident
    p.K = K;
    p.ring = reinterpret_cast<unsigned long long *>(workspace);
    NarrowParams q1 = p, q2 = p;
    if (peers != nullptr) {
        // edge buffers of a column-tiled multi-GPU run: pass 1 travels left -> right, pass 2 (flipped frame) right -> left
        q1.peer_in_l = peers->in[0]; q1.peer_out_l = peers->ack_out[0]; q1.peer_out_r = peers->out[0]; q1.peer_in_r = peers->ack_in[0];
        q2.peer_in_l = peers->in[1]; q2.peer_out_l = peers->ack_out[1]; q2.peer_out_r = peers->out[1]; q2.peer_in_r = peers->ack_in[1];
        q1.tag_base = q2.tag_base = peers->epoch << 16;
        q1.prev_ack = q2.prev_ack = peers->prev_epoch ? (peers->prev_epoch << 16) + (uint32_t)(peers->prev_rows) : 0u;
        q1.c_off = peers->c_off[0]; q2.c_off = peers->c_off[1];
        q1.Wg = q2.Wg = peers->Wg;
        q1.sheared_store = q2.sheared_store = 1;
        q1.nimg = q2.nimg = peers->nimg;
    }
    void *args1[] = {(void *)&q1}, *args2[] = {(void *)&q2};
    const int passes = peers != nullptr ? peers->passes : 3;
    if (passes & 1) {
        PB200_CUDA(cudaMemsetAsync(p.flag, 0, sizeof(int), s));
        PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
        PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w1, dim3(nstrips), dim3(threads), args1, smem1, s));
        PB200_LAUNCH_CHECK("sgm_wave1_kernel<down, census>");
    }
    if (passes & 2) {
        PB200_CUDA(cudaMemsetAsync(p.ring, 0, wring, s));
        PB200_CUDA(cudaLaunchCooperativeKernel((const void *)w2, dim3(nstrips), dim3(threads), args2, smem2, s));
        PB200_LAUNCH_CHECK("sgm_wave1_kernel<up>");
    }
    note_path(STAGE_SGM, PATH_SGM_WAVE1_CENSUS, NR * 10 + CB);
    *done = true;
    return PB200_OK;
}

}  // namespace

size_t sgm_wave1_ring_bytes(int W, int D) {     // boundary blocks of one GPU's strips (at most one strip per 4 columns)
    const int NR = (D + 63) / 64;
    return ((size_t)(W + 3) / 4 + 1) * ((size_t)W1_NRG * 4 * NR * 32 + 16) * sizeof(unsigned long long);
}
size_t sgm_wave1_edge_bytes(int D) {            // one edge buffer of a column-tiled run: a boundary block (data | ack)
    const int NR = (D + 63) / 64;
    return ((size_t)W1_NRG * 4 * NR * 32 + 16) * sizeof(unsigned long long);
}

int sgm_wave1_strip_width(int W) {              // columns per strip, or 0 when one co-resident wave cannot hold the image
    int K = ceil_div(W, sm_count());
    if (K < 4) K = 4;
    while (K <= W1_MAXK && W > 1 && W % K == 1) ++K;         // a strip's out-relay forwards its last TWO columns
    return K <= W1_MAXK ? K : 0;
}

// Skewed one-column wavefront for the fused Census -> SGM stage.  p.W = columns of this GPU's tile (= the image width on
// one GPU); `peers` (optional) = the edge buffers, ring geometry and epoch of a column-tiled multi-GPU run.
int sgm_census_wave1_launch(NarrowParams p, int NR, bool bytes, void *workspace, size_t ring_room, const Wave1Peers *peers, cudaStream_t s,
                            bool *done) {
    *done = false;
    const int K = sgm_wave1_strip_width(p.W);
    if (K == 0 || p.descR4 == nullptr) return PB200_OK;
    const int nstrips = ceil_div(p.W, K);
    if (peers == nullptr) { p.Wg = p.W; p.c_off = 0; p.sheared_store = 0; p.tag_base = 0; }
    if (p.nimg < 1) p.nimg = 1;
    if (NR == 4) return bytes ? launch_wave1<4, 1>(p, K, nstrips, workspace, ring_room, peers, s, done)
                              : launch_wave1<4, 2>(p, K, nstrips, workspace, ring_room, peers, s, done);
    if (NR == 2) return bytes ? launch_wave1<2, 1>(p, K, nstrips, workspace, ring_room, peers, s, done)
                              : launch_wave1<2, 2>(p, K, nstrips, workspace, ring_room, peers, s, done);
    return launch_wave1<1, 2>(p, K, nstrips, workspace, ring_room, peers, s, done);
}

}  // namespace pb200
