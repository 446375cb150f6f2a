// common.cuh -- shared helpers for the sm_100a kernels of pandora_b200.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "pandora_b200.h"

namespace pb200 {

// ---- error plumbing (api.cu owns the storage) ---------------------------------------------------
void set_error(const char *fmt, ...);
int check_cuda(cudaError_t err, const char *what);
void count_launch(int n = 1);

#define PB200_CUDA(call)                                                \
    do {                                                                \
        int _rc = ::pb200::check_cuda((call), #call);                   \
        if (_rc != PB200_OK) return _rc;                                \
    } while (0)

#define PB200_LAUNCH_CHECK(name)                                        \
    do {                                                                \
        ::pb200::count_launch();                                        \
        int _rc = ::pb200::check_cuda(cudaGetLastError(), name);        \
        if (_rc != PB200_OK) return _rc;                                \
    } while (0)

static inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }
int sm_count();

// ---- kernel-selection options and the record of which path ran (api.cu owns the storage) -----------
// Options are set through the C-ABI (pb200_set_option), never through the environment: the parity tests use them to
// run the alternative kernels of a stage on the same inputs.  -1 = unset (the library's own choice).
enum Option {
    OPT_SGM_NO_WAVE = 0,      // 1: keep the four-launch packed schedule instead of the wavefront passes
    OPT_SGM_NO_BYTE_TIER,     // 1: 16-bit cost storage even when cost + P2 fits a byte
    OPT_SGM_WAVE_KERNEL,      // 1: one column per warp (sgm_wave1_kernel), 2: two columns per warp (sgm_wave_kernel)
    OPT_CENSUS_DIRECT,        // 0 / 1: TMA-tiled / direct fill kernel
    OPT_CENSUS_TILE,          // tile size (floats) of the TMA-tiled fill
    OPT_CBCA_PIPE,            // 1: staged CBCA kernel instead of the register kernel
    OPT_CBCA_BANDS,           // row bands of the register kernel
    OPT_REVERSE_GATHER,       // 1: plain gather kernel for reverse_cost_volume
    OPT_FUSE_CENSUS_SGM,      // 0: pb200_disparity_host keeps Census and SGM apart
    OPT_SAD_TAPS,             // 1: tap-ordered SAD / SSD / ZNCC kernel instead of the running sums
    OPT_COUNT
};
int option(Option o);
// which kernel family served the last call of a stage (thread-local; read back with pb200_last_path)
enum Stage { STAGE_SGM = 0, STAGE_CBCA, STAGE_CENSUS, STAGE_REVERSE, STAGE_SAD, STAGE_COUNT };
enum Path {
    PATH_NONE = 0,
    PATH_SGM_FLOAT = 1, PATH_SGM_PACKED4 = 2, PATH_SGM_WAVE2 = 3, PATH_SGM_WAVE2_CENSUS = 4, PATH_SGM_WAVE1 = 5, PATH_SGM_WAVE1_CENSUS = 6,
    PATH_CBCA_REG = 10, PATH_CBCA_PIPE = 11, PATH_CBCA_STAGED = 12,
    PATH_CENSUS_TMA = 20, PATH_CENSUS_DIRECT = 21, PATH_CENSUS_SUBPIX = 22,
    PATH_REVERSE_TILED = 30, PATH_REVERSE_GATHER = 31,
    PATH_SAD_TAPS = 40, PATH_SAD_RUNNING = 41
};
void note_path(Stage st, int path, int detail = 0);

struct CensusDesc {                                   // what the fused Census -> SGM kernels read (census.cu -> sgm_narrow.cu)
    const uint32_t *L, *R;                            // planar one-word descriptors, row pitch `pitch`
    int pitch;
    const uint32_t *R4;                               // four shifted, padded copies of the right descriptors (skewed wavefront) or NULL
    int pitch4, padl;
};

// Edge buffers of a column-tiled multi-GPU run of the skewed wavefront, per pass (0 = top-down, 1 = bottom-up): `in` =
// local boundary block the travel-frame LEFT neighbour stores into, `ack_out` = that neighbour's credit word (peer
// pointer), `out` = the travel-frame RIGHT neighbour's boundary block (peer pointer), `ack_in` = local credit word it
// stores into; `c_off` = global sheared column of this tile's column 0 in that pass's frame.
struct Wave1Peers {
    unsigned long long *in[2], *ack_out[2], *out[2], *ack_in[2];
    int c_off[2];
    int Wg;
    uint32_t epoch;
    uint32_t prev_rows;       // rows (images x H) of the call that used these links last
    uint32_t prev_epoch;      // epoch of the image that used these links last (0: none): a pass first waits until the neighbour
                              // has consumed that image's last row, so that back-to-back passes of the SAME direction are safe
    int passes;               // bit 0: pass 1, bit 1: pass 2
    int nimg;                 // images of the batch (they follow each other in one wave)
};

// ---- device helpers -----------------------------------------------------------------------------
__device__ __forceinline__ float nan_f() { return __int_as_float(0x7fc00000); }

// exact float of a small non-negative integer (< 2^23) without I2F: one LOP + one FADD.
__device__ __forceinline__ float small_int_to_float(uint32_t n) {
    return __int_as_float(0x4B000000u | n) - 8388608.0f;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier + TMA (cp.async.bulk) wrappers: 1-D bulk copies, SASS UBLKCP -----------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy (bytes multiple of 16, both addresses 16-byte aligned)
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (bulk_group completion)
__device__ __forceinline__ void tma_store_1d(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// streaming (evict-first) vector store / load for data touched once
__device__ __forceinline__ void st_cs_f4(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }
__device__ __forceinline__ float4 ld_cs_f4(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }

}  // namespace pb200
