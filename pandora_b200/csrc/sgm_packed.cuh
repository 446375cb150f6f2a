// sgm_packed.cuh -- pieces shared by the packed-integer SGM kernels (sgm_narrow.cu: horizontal passes, strip sweeps and the
// two-column wavefront; sgm_wave1.cu: the one-column wavefront): parameter block, storage tiers, the recurrence step on
// packed 16-bit states, flag-in-data ring words, cp.async / shared-memory access by 32-bit address, mailbox flags.
#pragma once
#include <cstdlib>

#include "sgm_common.cuh"

namespace pb200 {

struct NarrowParams {
    const float *cv;          // float32 (H, W, D) input costs (pass E only)
    uint32_t *buf;            // the output buffer as packed words: per pixel [D/2 words C16][D/2 words P16]
    int H, W, D;
    uint32_t p1p1, p2p2;      // penalties replicated in both halves
    uint32_t inv;             // invalid_value
    float cost_ok_max;        // largest admissible cost
    int *flag;                // raised (1) when the volume does not qualify
    int dy, overcounting;     // sweeps
    float *disp;
    uint8_t *all_nan;
    int dmin;
    float invalid_disparity;
    unsigned long long *ring;
    const uint32_t *halo_in;  // packed (3, W, D/2 words) states of the row just outside the tile (order dx = 0, +1, -1) or NULL
    uint32_t *halo_out;       // packed states of this tile's last row in travel direction or NULL
    int debug;                // PB200_SGM_DEBUG bit 0: no strip exchange (timing experiments only, wrong results)
    // fused Census source of the first wavefront pass (CENSUS = true): planar one-word descriptors (census.cu)
    const uint32_t *descL, *descR;
    int pitch, half;
    // skewed wavefront: the right descriptors as FOUR word-shifted, padded copies [row][s][pitch4]: copy s, index i holds the
    // descriptor of image column i + s - padl (flagged outside the image), so that the D-wide window of ANY pixel starts
    // on a 16-byte boundary in the copy s = (column + dmin) & 3 and needs no bounds check
    const uint32_t *descR4;
    int pitch4, padl;
    // one-column wavefront (sgm_wave1.cu): columns per strip, and the mailbox rings of a COLUMN-tiled multi-GPU run:
    // the strip at either edge of this GPU's tile exchanges its border states with the neighbouring GPU through
    // peer-mapped memory (NVLink) instead of the local L2 ring.  `l` / `r` are the logical sides of this pass's travel
    // frame; *_in are local buffers the neighbour writes, *_out are the neighbour's buffers (peer pointers).
    int K;
    unsigned long long *peer_in_l, *peer_out_l, *peer_in_r, *peer_out_r;
    uint32_t tag_base;        // added to every ring tag: (epoch << 16) for peer rings that are never cleared
    uint32_t prev_ack;        // peer links: the credit word value that says "the previous image has left this link" (0: none)
    int Wg, c_off;            // skewed wavefront: width of the (global) sheared ring, sheared column of this tile's column 0
    int sheared_store;        // 1: the volume / disparity tile is stored by sheared column (multi-GPU tiles), 0: by image column
    int nimg;                 // images of a batch that follow each other in one wave (descriptors, volume, disparity: [nimg][...])
    int period;               // two-column wavefront: rows per image of a batch stacked into one tall image of H rows (0: one image)
};


namespace {


// Timing experiments that produce WRONG results (no strip exchange, no mailbox waits) exist only in builds made with
// -DPB200_DEBUG_SWITCHES (tools/prof_fused.py); the shipped library has no such switch.
static inline int debug_switches() {
#ifdef PB200_DEBUG_SWITCHES
    const char *e = getenv("PB200_SGM_DEBUG");
    return e ? atoi(e) : 0;
#else
    return 0;
#endif
}

#ifndef PB200_RELAY_BACKOFF_NS
#define PB200_RELAY_BACKOFF_NS 40
#endif
constexpr uint32_t INF16 = 0x7FFFu;          // "+inf" for a 16-bit lane: larger than any state, INF + P cannot wrap
constexpr int NARROW_MAX = 8191;             // 8 directions x (cost + P2) must stay below 2^16


// NR = registers per lane of one packed state vector = D / 64: 1, 2, 4 (one aligned vector access per lane) and 3 (D = 192:
// no 12-byte vector access exists and a lane's block is only 4-byte aligned, so three scalar accesses).
template <int NR>
__device__ __forceinline__ void ld_words(const uint32_t *p, uint32_t (&v)[NR]) {
    if constexpr (NR == 4) { const uint4 t = *reinterpret_cast<const uint4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (NR == 2) { const uint2 t = *reinterpret_cast<const uint2 *>(p); v[0] = t.x; v[1] = t.y; }
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) v[j] = p[j];
    }
}
template <int NR>
__device__ __forceinline__ void st_words(uint32_t *p, const uint32_t (&v)[NR]) {
    if constexpr (NR == 4) *reinterpret_cast<uint4 *>(p) = make_uint4(v[0], v[1], v[2], v[3]);
    else if constexpr (NR == 2) *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) p[j] = v[j];
    }
}
template <int NR>
__device__ __forceinline__ void ld_floats(const float *p, float (&v)[NR]) {
    if constexpr (NR == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (NR == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); v[0] = t.x; v[1] = t.y; }
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) v[j] = p[j];
    }
}
template <int NR>
__device__ __forceinline__ void st_floats(float *p, const float (&v)[NR]) {
    if constexpr (NR == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (NR == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) p[j] = v[j];
    }
}

// ---- storage tiers ------------------------------------------------------------------------------------------
// CB = bytes per stored cost.  CB == 2: C16 words [0, D/2), P16 words [D/2, D) of a pixel's D-word region.
// CB == 1 (cost + P2 <= 127): C8 words [0, D/4), P16 words [D/4, 3D/4), and the E -> W hand-over P8 (L_E <= 255)
// in words [3D/4, D).  In registers a cost is always one 16-bit half; the NaN flag sits in bit 15 (CB 2) or 7 (CB 1).
template <int CB> struct Tier;
template <> struct Tier<2> { static constexpr uint32_t FLAGS = 0x80008000u, VALUES = 0x7FFF7FFFu, FLAG1 = 0x8000u; };
template <> struct Tier<1> { static constexpr uint32_t FLAGS = 0x00800080u, VALUES = 0x007F007Fu, FLAG1 = 0x80u; };

// raw words of a lane's costs (NR * CB / 2 words) and their expansion to one 16-bit half per cost; kept apart so a
// prefetch can leave the raw words in flight and unpack them only when the row is consumed
template <int NR, int CB>
__device__ __forceinline__ void ld_cost_raw(const uint32_t *pix, int lane, uint32_t (&w)[NR * CB / 2]) {
    ld_words<NR * CB / 2>(pix + lane * (NR * CB / 2), w);
}
template <int NR, int CB>
__device__ __forceinline__ void unpack_cost(const uint32_t (&w)[NR * CB / 2], uint32_t (&c)[NR]) {
    if constexpr (CB == 2) {
#pragma unroll
        for (int j = 0; j < NR; ++j) c[j] = w[j];
    } else {
#pragma unroll
        for (int q = 0; q < NR / 2; ++q) {
            c[2 * q] = __byte_perm(w[q], 0u, 0x4140);            // bytes (b0, 0, b1, 0)
            c[2 * q + 1] = __byte_perm(w[q], 0u, 0x4342);        // bytes (b2, 0, b3, 0)
        }
    }
}
template <int NR, int CB>
__device__ __forceinline__ void st_cost(uint32_t *pix, int lane, const uint32_t (&c)[NR]) {
    if constexpr (CB == 2) {
        st_words<NR>(pix + lane * NR, c);
    } else {
        uint32_t w[NR / 2];
#pragma unroll
        for (int q = 0; q < NR / 2; ++q) w[q] = __byte_perm(c[2 * q], c[2 * q + 1], 0x6420);
        st_words<NR / 2>(pix + lane * (NR / 2), w);
    }
}
// word offsets of the partial sums inside a pixel region
template <int CB> __device__ __forceinline__ int p16_off(int D) { return CB == 2 ? D / 2 : D / 4; }
__device__ __forceinline__ int p8_off(int D) { return 3 * (D / 4); }

// One recurrence step on packed states: L = cc + (min(Lp[d], min(Lp[d-1], Lp[d+1]) + P1, m + P2) - m).
// Register j of a lane holds disparity NR*lane + j (low half) and D/2 + NR*lane + j (high half).
template <int NR>
__device__ __forceinline__ void nstep(const uint32_t (&cc)[NR], const uint32_t (&Lp)[NR], uint32_t (&L)[NR], int lane, uint32_t p1p1,
                                      uint32_t p2p2) {
    uint32_t mn = Lp[0];
    if constexpr (NR >= 3) {
        mn = __vimin3_u16x2(Lp[0], Lp[1], Lp[2]);
#pragma unroll
        for (int j = 3; j < NR; ++j) mn = __vminu2(mn, Lp[j]);
    } else {
#pragma unroll
        for (int j = 1; j < NR; ++j) mn = __vminu2(mn, Lp[j]);
    }
    // both halves := min(low, high); a 32-bit minimum over words of the form (v << 16) | v is the minimum of the v's
    // replicated in both halves, i.e. the word the step needs -- no mask, no multiply
    mn = __vminu2(mn, __byte_perm(mn, 0u, 0x1032));
    const uint32_t mm = __reduce_min_sync(0xffffffffu, mn);
    const uint32_t mp2 = mm + p2p2;
    const uint32_t up = __shfl_sync(0xffffffffu, Lp[NR - 1], (lane + 31) & 31);
    const uint32_t dn = __shfl_sync(0xffffffffu, Lp[0], (lane + 1) & 31);
    // lane 0: d-1 of its low half does not exist, d-1 of its high half (D/2 - 1) is lane 31's last LOW half:
    // (up << 16) | INF16.  lane 31: d+1 of its low half (D/2) is lane 0's first HIGH half, d+1 of its high half does
    // not exist: (dn >> 16) | (INF16 << 16).  One PRMT each, with a per-lane (loop-invariant) selector.
    const uint32_t lo0 = __byte_perm(up, INF16 * 0x10001u, lane == 0 ? 0x1054u : 0x3210u);
    const uint32_t hiN = __byte_perm(dn, INF16 * 0x10001u, lane == 31 ? 0x5432u : 0x3210u);
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const uint32_t lo = (j == 0) ? lo0 : Lp[j - 1];
        const uint32_t hi = (j == NR - 1) ? hiN : Lp[j + 1];
        const uint32_t t = __vimin3_u16x2(Lp[j], __vminu2(lo, hi) + p1p1, mp2);    // INF16 + P1 stays inside its half
        L[j] = cc[j] + (t - mm);          // t >= m in both halves: no borrow, and cc + P2 < 2^16: no carry
    }
}

// (Measured and dropped, round 2: t = __viaddmin_u16x2(__vimin3_u16x2(lo, hi, m + P2 - P1), P1, Lp[j]) -- VIMNMX3 + VIADDMNMX.U16x2,
// one instruction less per register and direction, 16 per pixel and pass -- ran the two-column stage at 14.61 ms instead of 14.07 ms
// at C3 (one-column 17.42 -> 17.26 ms): the fused add-minimum issues slower than the add and the minimum it replaces.)
// (Measured and dropped: the same step with its additions written as a * 1 + c with the 1 in a register the compiler cannot
// see through, i.e. IMAD on the idle FMA pipe instead of IADD3 on the ALU pipe that the packed min-ops saturate.  Two IMAD
// replace one IADD3 and the dependent chain gets longer: 15.40 -> 16.26 ms for the two-column kernels at C3, 17.42 -> 18.42 ms
// for the skewed ones, profiles/r2_wave_kernels.txt.)

// float32 cost -> 16-bit code (value, or invalid_value | 0x8000 for NaN); `bad` is raised for anything else
template <int CB>
__device__ __forceinline__ uint32_t encode_cost(float v, uint32_t inv, float ok_max, bool &bad) {
    const float t = v + 8388608.0f;                       // exact integer extraction for 0 <= v < 2^23
    const bool isn = (v != v);
    const bool ok = (t - 8388608.0f == v) && (v >= 0.f) && (v <= ok_max);
    bad = bad || !(ok || isn);
    return isn ? (inv | Tier<CB>::FLAG1) : (__float_as_uint(t) & 0xFFFFu);
}

// Two float32 costs -> one packed word of 16-bit codes (6.5 instructions per cost): NaN is first replaced by the float
// whose integer code is invalid_value | flag, t = v + 2^23 carries the integer in its low mantissa bits (one PRMT packs
// both), and the data condition "v is an integer in [0, ok_max]" is the single ordered comparison
// min(|t - 2^23|, ok_max) <> v  (false for NaN, true for fractions, negatives, too large values and infinities).
__device__ __forceinline__ uint32_t encode_pair(float vlo, float vhi, float nan_code, float ok_max, bool &bad) {
    const float alo = (vlo != vlo) ? nan_code : vlo, ahi = (vhi != vhi) ? nan_code : vhi;
    const float tlo = alo + 8388608.0f, thi = ahi + 8388608.0f;
    const float clo = fminf(fabsf(tlo - 8388608.0f), ok_max), chi = fminf(fabsf(thi - 8388608.0f), ok_max);
    bad = bad || (clo < vlo) || (clo > vlo) || (chi < vhi) || (chi > vhi);      // ordered <>: false for NaN
    return __byte_perm(__float_as_uint(tlo), __float_as_uint(thi), 0x5410);
}

template <int NR>
__device__ __forceinline__ void ll_send_u32(unsigned long long *slot, int lane, uint32_t tag, const uint32_t (&v)[NR]) {
#pragma unroll
    for (int j = 0; j < NR; ++j) {
        const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)v[j];
        asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(slot + j * 32 + lane), "l"(w) : "memory");
    }
}
template <int NR>
__device__ __forceinline__ void ll_recv_u32(const unsigned long long *slot, int lane, uint32_t tag, uint32_t (&v)[NR]) {
    unsigned long long w[NR];
    bool ok;
    // a failed poll backs off for a few dozen nanoseconds: polled flat out, the two relay warps of a CTA issue a third of
    // all instructions of the kernel (profiles/r2_ncu_wave2_pass1.txt: 134 M iterations), on the ALU pipe the compute warps need
    for (;;) {
        ok = true;
#pragma unroll
        for (int j = 0; j < NR; ++j) w[j] = ll_load(slot + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < NR; ++j) ok = ok && ((uint32_t)(w[j] >> 32) == tag);
        if (__all_sync(0xffffffffu, ok)) break;
        __nanosleep(PB200_RELAY_BACKOFF_NS);
    }
#pragma unroll
    for (int j = 0; j < NR; ++j) v[j] = (uint32_t)w[j];
}

template <int NWORDS>
__device__ __forceinline__ void cp_async_words(uint32_t smem_addr, const uint32_t *gsrc) {
    if constexpr (NWORDS == 1 || NWORDS == 2 || NWORDS == 4) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr), "l"(gsrc), "n"(NWORDS * 4) : "memory");
    } else {                                            // 3 words: cp.async copies 4, 8 or 16 bytes
#pragma unroll
        for (int j = 0; j < NWORDS; ++j)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr + 4u * j), "l"(gsrc + j) : "memory");
    }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// shared-memory access by 32-bit shared address (no generic-pointer conversion in the row loop)
template <int NR>
__device__ __forceinline__ void lds_words(uint32_t addr, uint32_t (&v)[NR]) {
    if constexpr (NR == 4) asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(addr));
    else if constexpr (NR == 2) asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(addr));
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v[j]) : "r"(addr + 4u * j));
    }
}
template <int NR>
__device__ __forceinline__ void sts_words(uint32_t addr, const uint32_t (&v)[NR]) {
    if constexpr (NR == 4) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
    else if constexpr (NR == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(v[0]), "r"(v[1]) : "memory");
    else {
#pragma unroll
        for (int j = 0; j < NR; ++j) asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr + 4u * j), "r"(v[j]) : "memory");
    }
}

__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

__device__ __forceinline__ void flag_publish(uint32_t addr, uint32_t v, bool relaxed) {
    if (relaxed) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
    else asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void flag_wait(uint32_t addr, uint32_t target) {
    uint32_t v;
    do {
        asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    } while (v < target);
}

// Hand-over events as mbarrier phases, four barriers per stream: event r completes phase r >> 2 of barrier r & 3 (32 bytes per
// stream).  The waiter sleeps in hardware; one predicated arrival (lane 0) with release semantics, the warp barrier in front
// orders the other lanes' data stores before it.
__device__ __forceinline__ void ev4_wait(uint32_t stream, int r) {
    const uint32_t addr = stream + (uint32_t)(r & 3) * 8u, parity = (uint32_t)(r >> 2) & 1u;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "EV4_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra EV4_DONE;\n"
        "bra EV4_WAIT;\n"
        "EV4_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void ev4_signal(uint32_t stream, int r, int lane) {
    const uint32_t addr = stream + (uint32_t)(r & 3) * 8u;
    __syncwarp();
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.eq.u32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(addr), "r"(lane) : "memory");
}

}  // namespace

}  // namespace pb200
