// sgm_common.cuh -- device helpers shared by the float (sgm.cu) and packed-integer (sgm_narrow.cu) SGM kernels.
#pragma once
#include "common.cuh"

namespace pb200 {

// upper bound of the strip-exchange ring of any SGM sweep, at most one strip per 4 columns: the float / packed strip
// sweeps use 2 x 2 vectors of vs 64-bit words per strip, the wavefront passes (sgm_wave_kernel) 12 mailbox vectors of
// D/2 <= vs/2 words = 6 vs words -- the larger one.  The narrow-path flag lives 256 bytes behind it
// (pb200_sgm_workspace_bytes = this + 512).
static inline size_t sgm_ring_max_bytes(int W, int D) {
    const size_t nstrips = (size_t)(W + 3) / 4 + 1;
    const size_t vs = (size_t)((D + 31) / 32) * 32;
    const size_t old_block = 6 * vs * sizeof(unsigned long long);
    // skewed wavefront (sgm_wave1.cu): 8 rows x 4 vectors of D / 2 tagged words + a credit line per strip boundary
    const size_t skew_block = ((size_t)8 * 4 * ((D + 63) / 64) * 32 + 16) * sizeof(unsigned long long);
    return nstrips * (old_block > skew_block ? old_block : skew_block);
}

__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float d;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));   // FMNMX3
    return d;
}
__device__ __forceinline__ float warp_min_redux(float a) {
    float d;
    asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(d) : "f"(a));   // CREDUX.MIN.F32
    return d;
}

// lane-major vector layout ([quad][lane][4]) for shared memory and the ring: every access instruction
// of a warp covers one contiguous span (no bank conflicts, fully coalesced).
template <int NPL>
__device__ __forceinline__ void lm_store(float *base, int lane, const float (&v)[NPL]) {
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q)
            reinterpret_cast<float4 *>(base)[q * 32 + lane] = make_float4(v[q * 4], v[q * 4 + 1], v[q * 4 + 2], v[q * 4 + 3]);
    } else if constexpr (NPL == 2) {
        reinterpret_cast<float2 *>(base)[lane] = make_float2(v[0], v[1]);
    } else {
        base[lane] = v[0];
    }
}
template <int NPL>
__device__ __forceinline__ void lm_load(const float *base, int lane, float (&v)[NPL]) {
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const float4 t = reinterpret_cast<const float4 *>(base)[q * 32 + lane];
            v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
        }
    } else if constexpr (NPL == 2) {
        const float2 t = reinterpret_cast<const float2 *>(base)[lane];
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = base[lane];
    }
}
template <int NPL>
__device__ __forceinline__ void lm_load_cg(const float *base, int lane, float (&v)[NPL]) {   // L2 only: never a stale L1 line
    if constexpr (NPL % 4 == 0) {
#pragma unroll
        for (int q = 0; q < NPL / 4; ++q) {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(base) + q * 32 + lane);
            v[q * 4] = t.x; v[q * 4 + 1] = t.y; v[q * 4 + 2] = t.z; v[q * 4 + 3] = t.w;
        }
    } else if constexpr (NPL == 2) {
        const float2 t = __ldcg(reinterpret_cast<const float2 *>(base) + lane);
        v[0] = t.x; v[1] = t.y;
    } else {
        v[0] = __ldcg(base + lane);
    }
}


// Flag-in-data hand-over between strips (the "LL" scheme of collective libraries): every float travels as
// one naturally aligned 64-bit word {row tag, value}.  A 64-bit scalar access is single-copy atomic, so a
// word whose tag matches is valid by itself: no fence on the sender, one L2 round trip on the receiver.
__device__ __forceinline__ void ll_store(unsigned long long *p, uint32_t tag, float v) {
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long *p) {
    unsigned long long w;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
    return w;
}
template <int NPL>
__device__ __forceinline__ void ll_send(unsigned long long *slot, int lane, uint32_t tag, const float (&v)[NPL]) {
#pragma unroll
    for (int j = 0; j < NPL; ++j) ll_store(slot + j * 32 + lane, tag, v[j]);
}
template <int NPL>
__device__ __forceinline__ void ll_recv(const unsigned long long *slot, int lane, uint32_t tag, float (&v)[NPL]) {
    unsigned long long w[NPL];
    bool ok;
    do {
        ok = true;
#pragma unroll
        for (int j = 0; j < NPL; ++j) w[j] = ll_load(slot + j * 32 + lane);
#pragma unroll
        for (int j = 0; j < NPL; ++j) ok = ok && ((uint32_t)(w[j] >> 32) == tag);
    } while (!__all_sync(0xffffffffu, ok));
#pragma unroll
    for (int j = 0; j < NPL; ++j) v[j] = __uint_as_float((uint32_t)w[j]);
}

}  // namespace pb200
