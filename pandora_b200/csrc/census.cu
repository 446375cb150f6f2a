// census.cu -- Census transform + Hamming cost-volume fill (+ optional fused winner-takes-all).
//
// Replaces census_transform / compute_matching_costs of the reference
// (src/pandora/matching_cost/cpp/src/census.cpp:45-95, 97-180) and the NaN pre-fill of
// matching_cost/census.py:138.  Semantics:
//   bit b (row-major over the w x w window) = neighbour > centre (strict float compare)
//   cv[y,x,k] = popcount(cL[y,x] ^ cR[y,x+dmin+k])  iff  half <= y < H-half, half <= x < W-half and
//               half <= x+dmin+k < W-half ; NaN otherwise.
//
// Data layout in HBM
//   descriptors: planar uint32  [NW][H][pitch]   (NW = ceil(w*w/32) words, pitch = roundup4(W) + 4 so
//                that every row is 16-byte aligned and can be bulk-copied by TMA).  Bit 31 of the LAST
//                word is never a census bit (w*w mod 32 <= 25 for every legal window); the transform
//                sets it for pixels whose window leaves the image, which makes the fill branch-free:
//                n = popc(a^b) | (int(a^b) >> 31)  ->  all ones -> NaN after the int->float trick.
//   cost volume: float32 (H, W, D), disparity fastest (the reference layout): the cells of TX
//                consecutive pixels of one row are ONE contiguous span of TX*D*4 bytes.
//
// Fill kernel (HBM-store bound: 4*D bytes written per pixel, 8 bytes read):
//   persistent CTAs loop over (row, TX-pixel) tiles; TMA (cp.async.bulk, mbarrier completion) stages
//   the left descriptors of the tile and the TX+D-1 right descriptors it can meet into shared
//   memory; every thread owns 4 consecutive disparities and slides over consecutive pixels keeping
//   its 4 right descriptors in registers (one LDS per new pixel); results go to a dense
//   [TX][D] float tile in shared memory (conflict-free 16-byte STS) which one elected thread
//   hands to the TMA store engine (cp.async.bulk.global.shared::cta) as a single contiguous span,
//   double-buffered so the next tile is computed while the previous one drains to HBM.
#include <cstdlib>

#include "common.cuh"

namespace pb200 {

static inline int census_nwords(int w) { return (w * w + 31) / 32; }
static inline int census_pitch(int W) { return ((W + 3) & ~3) + 4; }

// ------------------------------------------------------------------------------------------------
// transform
// ------------------------------------------------------------------------------------------------
template <int WIN>
__global__ void __launch_bounds__(256) census_transform_kernel(const float *__restrict__ img, int H, int W, int pitch,
                                                               uint32_t *__restrict__ desc, int row0, int row1, int col0, int col1) {
    constexpr int HALF = WIN / 2;
    constexpr int NW = (WIN * WIN + 31) / 32;
    constexpr int TW = 32, TH = 8;
    __shared__ float tile[TH + 2 * HALF][TW + 2 * HALF + 1];
    const int x0 = col0 + blockIdx.x * TW, y0 = row0 + blockIdx.y * TH;
    for (int i = threadIdx.y * TW + threadIdx.x; i < (TH + 2 * HALF) * (TW + 2 * HALF); i += TW * TH) {
        const int ty = i / (TW + 2 * HALF), tx = i % (TW + 2 * HALF);
        const int gy = y0 + ty - HALF, gx = x0 + tx - HALF;
        tile[ty][tx] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? img[(size_t)gy * W + gx] : 0.f;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= col1 || y >= row1) return;
    uint32_t words[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) words[i] = 0u;
    const bool inside = (x >= HALF && x < W - HALF && y >= HALF && y < H - HALF);
    if (inside) {
        const float c = tile[threadIdx.y + HALF][threadIdx.x + HALF];
#pragma unroll
        for (int wy = 0; wy < WIN; ++wy)
#pragma unroll
            for (int wx = 0; wx < WIN; ++wx) {
                constexpr int dummy = 0;
                (void)dummy;
                const int b = wy * WIN + wx;
                if (tile[threadIdx.y + wy][threadIdx.x + wx] > c) words[b >> 5] |= 1u << (b & 31);
            }
    } else {
        words[NW - 1] = 0x80000000u;  // "window leaves the image" flag (also for the pitch padding)
    }
#pragma unroll
    for (int i = 0; i < NW; ++i) desc[((size_t)i * H + y) * pitch + x] = words[i];
}

// ------------------------------------------------------------------------------------------------
// fill
// ------------------------------------------------------------------------------------------------
struct FillParams {
    const uint32_t *descL;
    const uint32_t *descR;
    float *cv;
    float *disp;        // optional fused WTA output
    uint8_t *all_nan;   // optional
    int H, W, D, dmin, half, pitch;
    int row0;           // first row of the range this launch fills
    int TX;             // pixels per tile (multiple of 4)
    int CH;             // consecutive pixels per work item
    int tiles_x;
    long n_tiles;
    int r_len;          // staged right descriptors per word plane (multiple of 4)
    float invalid_disparity;
};

template <int NW>
__device__ __forceinline__ uint32_t hamming_flagged(const uint32_t (&a)[NW], const uint32_t (&b)[NW]) {
    uint32_t n = 0;
#pragma unroll
    for (int i = 0; i < NW - 1; ++i) n += __popc(a[i] ^ b[i]);
    const uint32_t v = a[NW - 1] ^ b[NW - 1];
    return (n + __popc(v)) | (uint32_t)((int32_t)v >> 31);
}

template <int NW, bool VEC4>
__global__ void __launch_bounds__(256, 3) census_fill_kernel(const FillParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [2][TX*D] float out tiles | 2 x ([NW][r_len] right words | [NW][TX] left words) | 2 mbarriers.
    // The descriptors of the NEXT tile are requested (TMA) before the current tile is computed, so their HBM / L2
    // latency hides behind the compute instead of being paid once per tile.
    float *sOut = reinterpret_cast<float *>(smem_raw);
    const int tile_elems = p.TX * p.D;
    uint32_t *sR0 = reinterpret_cast<uint32_t *>(sOut + 2 * (size_t)((tile_elems + 3) & ~3));
    const int desc_words = NW * p.r_len + NW * p.TX;          // one staging buffer
    uint64_t *bars = reinterpret_cast<uint64_t *>(sR0 + 2 * desc_words);

    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        fence_mbar_init();
    }
    __syncthreads();

    // request the descriptors of `tile` into staging buffer `buf` (all threads call it; thread 0 drives the TMA)
    auto request = [&](long tile, int buf) {
        const int y = p.row0 + (int)(tile / p.tiles_x);
        if (!(y >= p.half && y < p.H - p.half)) return;        // border rows are all NaN: nothing to stage
        const int x0 = (int)(tile % p.tiles_x) * p.TX;
        uint32_t *sR = sR0 + buf * desc_words, *sL = sR + NW * p.r_len;
        const int lo = x0 + p.dmin;                 // right position of (pixel 0, k = 0)
        const int base = lo & ~3;                   // smem index j <-> right column base + j (floor to 4, also for lo < 0)
        const int clo = max(base, 0);
        const int chi = min(base + p.r_len, p.pitch);
        if (tid == 0) {
            uint32_t bytes = 0;
            if (chi > clo) bytes += (uint32_t)NW * (uint32_t)(chi - clo) * 4u;
            bytes += (uint32_t)NW * (uint32_t)p.TX * 4u;
            mbar_expect_tx(bars + buf, bytes);
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (chi > clo)
                    tma_load_1d(sR + w * p.r_len + (clo - base), p.descR + ((size_t)w * p.H + y) * p.pitch + clo,
                                (uint32_t)(chi - clo) * 4u, bars + buf);
                // pitch >= roundup4(W) + 4 and x0 + TX <= roundup(W, TX): clamp the tail to the pitch
                tma_load_1d(sL + w * p.TX, p.descL + ((size_t)w * p.H + y) * p.pitch + x0, (uint32_t)p.TX * 4u, bars + buf);
            }
        }
        // right columns outside the stored row: flag as invalid by hand (disjoint from the TMA range)
        for (int j = tid; j < p.r_len; j += blockDim.x) {
            const int c = base + j;
            if (c < clo || c >= chi) {
#pragma unroll
                for (int w = 0; w < NW; ++w) sR[w * p.r_len + j] = (w == NW - 1) ? 0x80000000u : 0u;
            }
        }
    };

    const int G = (p.D + 3) >> 2;             // groups of 4 disparities
    const int n_chunks = p.TX / p.CH;
    uint32_t parity = 0u;                       // bit b = phase parity of staging buffer b
    int it = 0;
    if ((long)blockIdx.x < p.n_tiles) request(blockIdx.x, 0);
    for (long tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        uint32_t *sR = sR0 + buf * desc_words, *sL = sR + NW * p.r_len;
        const int y = p.row0 + (int)(tile / p.tiles_x);
        const int x0 = (int)(tile % p.tiles_x) * p.TX;
        const int npx = min(p.TX, p.W - x0);
        float *out = sOut + (size_t)(it & 1) * ((tile_elems + 3) & ~3);
        // the bulk store issued two tiles ago read from this buffer: make sure it has been drained
        if (tid == 0) tma_store_wait_read<1>();
        __syncthreads();

        // the other staging buffer was last read while computing the previous tile (a barrier ago): refill it now
        if (tile + gridDim.x < p.n_tiles) request(tile + gridDim.x, buf ^ 1);
        const bool row_ok = (y >= p.half && y < p.H - p.half);
        if (row_ok) {
            const int lo = x0 + p.dmin;
            const int base = lo & ~3;
            mbar_wait(bars + buf, (parity >> buf) & 1u);
            parity ^= 1u << buf;
            __syncthreads();

            // ---- compute: work item = (chunk of CH consecutive pixels, group of 4 disparities) -----
            const int shift = lo - base;                // 0..3
            for (int item = tid; item < n_chunks * G; item += blockDim.x) {
                const int g = item % G, ch = item / G;
                const int k0 = g * 4;
                const int p0 = ch * p.CH;
                uint32_t win[4][NW];                    // right descriptors for k0..k0+3 at the current pixel
#pragma unroll
                for (int q = 0; q < 3; ++q)
#pragma unroll
                    for (int w = 0; w < NW; ++w) win[q + 1][w] = sR[w * p.r_len + shift + p0 + k0 + q];
#pragma unroll 4
                for (int pp = p0; pp < p0 + p.CH; ++pp) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        win[0][w] = win[1][w];
                        win[1][w] = win[2][w];
                        win[2][w] = win[3][w];
                        win[3][w] = sR[w * p.r_len + shift + pp + k0 + 3];
                    }
                    uint32_t a[NW];
#pragma unroll
                    for (int w = 0; w < NW; ++w) a[w] = sL[w * p.TX + pp];
                    float4 r;
                    if ((int32_t)a[NW - 1] < 0) {       // left window leaves the image: whole pixel NaN
                        r = make_float4(nan_f(), nan_f(), nan_f(), nan_f());
                    } else {
                        r.x = small_int_to_float(hamming_flagged<NW>(a, win[0]));
                        r.y = small_int_to_float(hamming_flagged<NW>(a, win[1]));
                        r.z = small_int_to_float(hamming_flagged<NW>(a, win[2]));
                        r.w = small_int_to_float(hamming_flagged<NW>(a, win[3]));
                    }
                    float *dst = out + (size_t)pp * p.D + k0;
                    if (VEC4) {
                        *reinterpret_cast<float4 *>(dst) = r;
                    } else {
                        dst[0] = r.x;
                        if (k0 + 1 < p.D) dst[1] = r.y;
                        if (k0 + 2 < p.D) dst[2] = r.z;
                        if (k0 + 3 < p.D) dst[3] = r.w;
                    }
                }
            }
        } else {
            for (int i = tid; i < npx * p.D; i += blockDim.x) out[i] = nan_f();
        }
        fence_proxy_async();
        __syncthreads();

        // ---- hand the tile to the TMA store engine (or store by hand when the span is unaligned) ----
        const size_t goff = ((size_t)y * p.W + x0) * p.D;
        const uint32_t bytes = (uint32_t)npx * (uint32_t)p.D * 4u;
        const bool bulk_ok = ((goff & 3) == 0) && ((bytes & 15u) == 0);
        if (bulk_ok) {
            if (tid == 0) {
                tma_store_1d(p.cv + goff, out, bytes);
                tma_store_commit();
            }
        } else {
            for (int i = tid; i < npx * p.D; i += blockDim.x) p.cv[goff + i] = out[i];
        }

        // ---- optional fused WTA straight from the shared-memory tile -----------------------------
        if (p.disp != nullptr) {
            const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
            for (int pp = warp; pp < npx; pp += nwarps) {
                // costs are integers < 2^16: key = cost << 16 | k keeps the first minimum
                uint32_t best = 0xFFFFFFFFu;
                const float *src = out + (size_t)pp * p.D;
                for (int k = lane; k < p.D; k += 32) {
                    const float v = src[k];
                    if (v == v) best = min(best, ((uint32_t)v << 16) | (uint32_t)k);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
                if (lane == 0) {
                    const size_t pix = (size_t)y * p.W + x0 + pp;
                    const bool none = (best == 0xFFFFFFFFu);
                    p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(best & 0xFFFFu));
                    if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                }
            }
        }
    }
    if (tid == 0) tma_store_wait<0>();
}

// ------------------------------------------------------------------------------------------------
// fill, barrier-free variant (D % 4 == 0): one warp per (row, 4 consecutive pixels)
// ------------------------------------------------------------------------------------------------
// The tiled kernel above pays three CTA barriers and a TMA round trip per 32-pixel tile.  Here a warp owns a span of
// 4 pixels x D disparities (4*D*4 contiguous bytes of the volume): a lane takes groups of 4 consecutive disparities,
// loads the 7 right descriptors those meet over the 4 pixels straight from global memory (L1 / L2 resident: a
// descriptor row is 16 KB and is reused by every warp of the row), and writes its float4 results with streaming
// 16-byte stores -- a warp instruction covers 512 contiguous bytes.  No shared memory, no barrier, every warp
// independent; the store stream is the only thing left to wait for.
template <int NW, bool WTA>
__global__ void __launch_bounds__(256) census_fill_direct_kernel(const FillParams p) {
    const int lane = threadIdx.x & 31;
    const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int G = p.D >> 2;
    const int spans_x = (p.W + 3) >> 2;
    const long n_items = (long)spans_x * (p.tiles_x /* rows of this launch */);
    for (long item = gw; item < n_items; item += nwarps) {
        const int y = p.row0 + (int)(item / spans_x);
        const int x0 = (int)(item % spans_x) << 2;
        const int npx = min(4, p.W - x0);
        const bool row_ok = (y >= p.half && y < p.H - p.half);
        float *dst_row = p.cv + ((size_t)y * p.W + x0) * p.D;
        uint32_t best[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        // Interior spans of one-word descriptors (every right column any lane meets lies inside the descriptor row, four
        // pixels): no per-load bounds test, one base pointer with immediate offsets, and a WTA key that needs no select -- a
        // flagged cost is all ones, so (cost << 16 | k) >= 0xFFFF0000 can never beat a real one (cost <= 25).  The first
        // version of this kernel executed 490 instructions per 16 cells and was bound by their issue (70 %, ncu).
        if (NW == 1 && row_ok && npx == 4 && x0 + p.dmin >= 0 && x0 + p.dmin + p.D + 3 <= p.pitch) {
            uint32_t a[4];
            const uint32_t *lrow = p.descL + (size_t)y * p.pitch + x0;
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) a[pp] = lrow[pp];
            const uint32_t *rrow = p.descR + (size_t)y * p.pitch + x0 + p.dmin;
            for (int g = lane; g < G; g += 32) {
                const uint32_t *rp = rrow + 4 * g;
                uint32_t rw[7];
#pragma unroll
                for (int q = 0; q < 7; ++q) rw[q] = __ldg(rp + q);
                float *dst = dst_row + 4 * g;
                const uint32_t k0 = (uint32_t)(4 * g);
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    float4 r = make_float4(nan_f(), nan_f(), nan_f(), nan_f());
                    if ((int32_t)a[pp] >= 0) {               // (warp-uniform) else: left window leaves the image, whole pixel NaN
                        uint32_t h[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t v = a[pp] ^ rw[pp + q];
                            h[q] = (uint32_t)__popc(v) | (uint32_t)((int32_t)v >> 31);
                        }
                        r = make_float4(small_int_to_float(h[0]), small_int_to_float(h[1]), small_int_to_float(h[2]), small_int_to_float(h[3]));
                        if (WTA) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) best[pp] = min(best[pp], (h[q] << 16) | (k0 + q));
                        }
                    }
                    *reinterpret_cast<float4 *>(dst + (size_t)pp * p.D) = r;
                }
            }
            if (WTA) {
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) best[pp] = best[pp] >= 0xFFFF0000u ? 0xFFFFFFFFu : best[pp];
            }
        } else if (row_ok) {
            uint32_t a[4][NW];                           // left descriptors of the 4 pixels (same for every lane)
#pragma unroll
            for (int pp = 0; pp < 4; ++pp)
#pragma unroll
                for (int w = 0; w < NW; ++w) a[pp][w] = p.descL[((size_t)w * p.H + y) * p.pitch + x0 + pp];   // x0 + 3 < pitch
            for (int g = lane; g < G; g += 32) {
                const int k0 = g << 2;
                const int c0 = x0 + p.dmin + k0;         // right column of (pixel 0, disparity k0)
                uint32_t rw[7][NW];
#pragma unroll
                for (int q = 0; q < 7; ++q) {
                    const int c = c0 + q;
                    const bool in = (c >= 0 && c < p.pitch);
#pragma unroll
                    for (int w = 0; w < NW; ++w)
                        rw[q][w] = in ? __ldg(p.descR + ((size_t)w * p.H + y) * p.pitch + c) : ((w == NW - 1) ? 0x80000000u : 0u);
                }
#pragma unroll
                for (int pp = 0; pp < 4; ++pp) {
                    if (pp < npx) {
                        float4 r;
                        if ((int32_t)a[pp][NW - 1] < 0) {   // left window leaves the image: whole pixel NaN
                            r = make_float4(nan_f(), nan_f(), nan_f(), nan_f());
                        } else {
                            const uint32_t h0 = hamming_flagged<NW>(a[pp], rw[pp]), h1 = hamming_flagged<NW>(a[pp], rw[pp + 1]),
                                           h2 = hamming_flagged<NW>(a[pp], rw[pp + 2]), h3 = hamming_flagged<NW>(a[pp], rw[pp + 3]);
                            r = make_float4(small_int_to_float(h0), small_int_to_float(h1), small_int_to_float(h2), small_int_to_float(h3));
                            if (WTA) {
                                // a flagged cost has all ones above bit 8: it can never win, and a pixel with nothing but
                                // flagged costs keeps the sentinel
                                best[pp] = min(best[pp], (h0 & 0x80000000u) ? 0xFFFFFFFFu : ((h0 << 16) | (uint32_t)k0));
                                best[pp] = min(best[pp], (h1 & 0x80000000u) ? 0xFFFFFFFFu : ((h1 << 16) | (uint32_t)(k0 + 1)));
                                best[pp] = min(best[pp], (h2 & 0x80000000u) ? 0xFFFFFFFFu : ((h2 << 16) | (uint32_t)(k0 + 2)));
                                best[pp] = min(best[pp], (h3 & 0x80000000u) ? 0xFFFFFFFFu : ((h3 << 16) | (uint32_t)(k0 + 3)));
                            }
                        }
                        *reinterpret_cast<float4 *>(dst_row + (size_t)pp * p.D + k0) = r;
                    }
                }
            }
        } else {
            const float4 r = make_float4(nan_f(), nan_f(), nan_f(), nan_f());
            for (int i = lane; i < npx * G; i += 32) st_cs_f4(dst_row + (size_t)i * 4, r);
        }
        if (WTA) {
#pragma unroll
            for (int pp = 0; pp < 4; ++pp) {
                const uint32_t b = __reduce_min_sync(0xffffffffu, best[pp]);
                if (lane == pp && pp < npx) {
                    const size_t pix = (size_t)y * p.W + x0 + pp;
                    const bool none = (b == 0xFFFFFFFFu);
                    p.disp[pix] = none ? p.invalid_disparity : (float)(p.dmin + (int)(b & 0xFFFFu));
                    if (p.all_nan) p.all_nan[pix] = none ? 1 : 0;
                }
            }
        }
    }
}

template <int NW>
static int launch_fill_direct(FillParams p, int rows, cudaStream_t s) {
    p.tiles_x = rows;                                    // reused as "rows of this launch" by the direct kernel
    const long items = (long)((p.W + 3) >> 2) * rows;
    long blocks = (items + 7) / 8;
    const long cap = (long)sm_count() * 8;               // 8 CTAs x 8 warps per SM
    if (blocks > cap) blocks = cap;
    if (p.disp != nullptr) census_fill_direct_kernel<NW, true><<<(int)blocks, 256, 0, s>>>(p);
    else census_fill_direct_kernel<NW, false><<<(int)blocks, 256, 0, s>>>(p);
    PB200_LAUNCH_CHECK("census_fill_direct_kernel");
    return PB200_OK;
}

template <int WIN>
static int launch_transform(const float *img, int H, int W, int pitch, uint32_t *desc, int row0, int row1, cudaStream_t s, int col0 = 0,
                            int col1 = -1) {
    if (col1 < 0) col1 = pitch;                      // default: every column of the pitch (the padding gets the flag)
    if (col1 <= col0) return PB200_OK;
    dim3 block(32, 8), grid(ceil_div(col1 - col0, 32), ceil_div(row1 - row0, 8));
    census_transform_kernel<WIN><<<grid, block, 0, s>>>(img, H, W, pitch, desc, row0, row1, col0, col1);
    PB200_LAUNCH_CHECK("census_transform_kernel");
    return PB200_OK;
}

template <int NW>
static int launch_fill(const FillParams &p, size_t smem, int grid, cudaStream_t s) {
    if ((p.D & 3) == 0) {
        PB200_CUDA(cudaFuncSetAttribute(census_fill_kernel<NW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        census_fill_kernel<NW, true><<<grid, 256, smem, s>>>(p);
    } else {
        PB200_CUDA(cudaFuncSetAttribute(census_fill_kernel<NW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        census_fill_kernel<NW, false><<<grid, 256, smem, s>>>(p);
    }
    PB200_LAUNCH_CHECK("census_fill_kernel");
    return PB200_OK;
}

int sgm_census_plan(int window, int W, int D, float p1, float p2);   // sgm_narrow.cu: 0 = not eligible, 1 = skewed wavefront, 2 = two-column wavefront
int sgm_census_wave_try(const CensusDesc &desc, int window, float *out, int H, int W, int D, float p1,
                        float p2, int overcounting, float *disp, int dmin, float invalid_disparity, uint8_t *all_nan, void *workspace,
                        size_t workspace_bytes, cudaStream_t s, bool *done, const Wave1Peers *peers = nullptr, int period = 0);   // sgm_narrow.cu
size_t sgm_wave1_edge_bytes(int D);                                            // sgm_wave1.cu

// Right descriptors in the layout of the skewed wavefront (sgm_wave1.cu): copy s, index i = descriptor of image column
// i + s - padl, or the "window leaves the image" flag.  One thread per (row, column of the padded range) computes the
// descriptor once and stores it into the four copies.
template <int WIN>
__global__ void __launch_bounds__(256) census_transform_shifted_kernel(const float *__restrict__ img, int H, int W, int pitch4, int padl,
                                                                       uint32_t *__restrict__ desc4, int i0, int i1) {
    constexpr int HALF = WIN / 2;
    const int i = i0 + blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (i >= i1) return;
    const int x = i - padl;
    uint32_t word = 0x80000000u;
    if (x >= HALF && x < W - HALF && y >= HALF && y < H - HALF) {
        const float c = __ldg(img + (size_t)y * W + x);
        word = 0u;
#pragma unroll
        for (int wy = 0; wy < WIN; ++wy)
#pragma unroll
            for (int wx = 0; wx < WIN; ++wx)
                if (__ldg(img + (size_t)(y + wy - HALF) * W + (x + wx - HALF)) > c) word |= 1u << (wy * WIN + wx);
    }
#pragma unroll
    for (int s = 0; s < 4; ++s)
        if (i - s >= 0 && i - s < pitch4) desc4[((size_t)y * 4 + s) * pitch4 + (i - s)] = word;      // [row][copy][column]
}
static inline int census_padl(int dmin) { return (((dmin < 0 ? -dmin : 0) + 3) & ~3) + 4; }
static inline int census_pitch4(int W, int dmin, int D) {
    const int dmax = dmin + D - 1;
    return (census_padl(dmin) + W + (dmax > 0 ? dmax : 0) + 8 + 3) & ~3;
}

static int census_transform_pair(const float *d_left, const float *d_right, int H, int W, int window, uint32_t *descL, uint32_t *descR,
                                 int row_begin, int row_end, cudaStream_t s) {
    const int pitch = census_pitch(W);
    int rc;
#define PB200_T(WIN)                                                       \
    case WIN:                                                              \
        rc = launch_transform<WIN>(d_left, H, W, pitch, descL, row_begin, row_end, s);         \
        if (rc == PB200_OK) rc = launch_transform<WIN>(d_right, H, W, pitch, descR, row_begin, row_end, s); \
        break;
    switch (window) {
        PB200_T(3) PB200_T(5) PB200_T(7) PB200_T(9) PB200_T(11) PB200_T(13)
        default: rc = PB200_ERR_UNSUPPORTED;
    }
#undef PB200_T
    return rc;
}

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_census_descriptors_rows(const float *d_left, const float *d_right, int H, int W, int window, void *d_workspace,
                                             size_t workspace_bytes, int row_begin, int row_end, void *stream) {
    if (!d_left || !d_right || !d_workspace || H <= 0 || W <= 0 || row_begin < 0 || row_end > H || row_begin >= row_end) {
        set_error("pb200_census_descriptors_rows: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (window != 3 && window != 5 && window != 7 && window != 9 && window != 11 && window != 13) {
        set_error("pb200_census_descriptors_rows: window_size %d not in {3,5,7,9,11,13}", window);
        return PB200_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pb200_census_workspace_bytes(H, W, window)) {
        set_error("pb200_census_descriptors_rows: workspace too small");
        return PB200_ERR_WORKSPACE;
    }
    uint32_t *descL = (uint32_t *)d_workspace;
    uint32_t *descR = descL + (size_t)census_nwords(window) * H * census_pitch(W);
    return census_transform_pair(d_left, d_right, H, W, window, descL, descR, row_begin, row_end, (cudaStream_t)stream);
}

// descriptors of the fused stage in the layout its kernels read (plan 1: left + four shifted right copies; plan 2: the
// standard pair); workspace: [standard left | standard right] then, for plan 1, the shifted right copies
// `col0 .. col1` (plan 1 only): the image columns whose PIXELS this call will process (a column tile of a multi-GPU run
// touches Wt + H - 1 of them); the left descriptors of those columns and the right descriptors their windows can meet
static int census_sgm_descriptors(int plan, const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D,
                                  void *ws, cudaStream_t s, int col0 = 0, int col1 = -1) {
    const int pitch = census_pitch(W);
    uint32_t *descL = (uint32_t *)ws, *descR = descL + (size_t)H * pitch;
    if (plan != 1) return census_transform_pair(d_left, d_right, H, W, window, descL, descR, 0, H, s);
    const bool all = col1 < 0;
    int rc = window == 3 ? launch_transform<3>(d_left, H, W, pitch, descL, 0, H, s, all ? 0 : col0, all ? -1 : col1)
                         : launch_transform<5>(d_left, H, W, pitch, descL, 0, H, s, all ? 0 : col0, all ? -1 : col1);
    if (rc != PB200_OK) return rc;
    const int pitch4 = census_pitch4(W, dmin, D), padl = census_padl(dmin);
    uint32_t *desc4 = descR + (size_t)H * pitch;
    // copy index i holds image column i + s - padl: the windows of columns [col0, col1) meet [col0 + dmin, col1 + dmin + D)
    int i0 = 0, i1 = pitch4 + 3;
    if (!all) {
        i0 = col0 + dmin + padl - 4;
        i1 = col1 + dmin + D + padl + 4;
        if (i0 < 0) i0 = 0;
        if (i1 > pitch4 + 3) i1 = pitch4 + 3;
    }
    if (i1 <= i0) return PB200_OK;
    dim3 grid(ceil_div(i1 - i0, 256), H);
    if (window == 3) census_transform_shifted_kernel<3><<<grid, 256, 0, s>>>(d_right, H, W, pitch4, padl, desc4, i0, i1);
    else census_transform_shifted_kernel<5><<<grid, 256, 0, s>>>(d_right, H, W, pitch4, padl, desc4, i0, i1);
    PB200_LAUNCH_CHECK("census_transform_shifted_kernel");
    return PB200_OK;
}

extern "C" size_t pb200_census_sgm_workspace_bytes(int H, int W, int window, int dmin, int D) {
    if (H <= 0 || W <= 0 || D <= 0) return 0;
    size_t bytes = pb200_census_workspace_bytes(H, W, window);
    if (window == 3 || window == 5) bytes += 4 * (size_t)H * census_pitch4(W, dmin, D) * sizeof(uint32_t);
    return bytes;
}

extern "C" int pb200_census_sgm_descriptors(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D, float p1,
                                            float p2, void *d_census_workspace, size_t census_workspace_bytes, int *eligible, void *stream) {
    if (!d_left || !d_right || !d_census_workspace || !eligible || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_census_sgm_descriptors: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    const int plan = sgm_census_plan(window, W, D, p1, p2);
    *eligible = plan != 0;
    if (plan == 0) return PB200_OK;
    if (census_workspace_bytes < pb200_census_sgm_workspace_bytes(H, W, window, dmin, D)) {
        set_error("pb200_census_sgm_descriptors: census workspace too small (pb200_census_sgm_workspace_bytes)");
        return PB200_ERR_WORKSPACE;
    }
    return census_sgm_descriptors(plan, d_left, d_right, H, W, window, dmin, D, d_census_workspace, (cudaStream_t)stream);
}

extern "C" int pb200_census_sgm(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D, float p1, float p2,
                                int overcounting, float *d_cv_out, void *d_census_workspace, size_t census_workspace_bytes,
                                void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp, float invalid_disparity,
                                uint8_t *d_all_nan, int descriptors_ready, int *ran, void *stream) {
    if (!ran) {
        set_error("pb200_census_sgm: ran must not be NULL");
        return PB200_ERR_BAD_ARG;
    }
    *ran = 0;
    if (!d_left || !d_right || !d_cv_out || !d_census_workspace || !d_sgm_workspace || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_census_sgm: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    // eligibility first: nothing is launched for a configuration the fused kernels do not take (the caller runs the two steps)
    int plan = sgm_census_plan(window, W, D, p1, p2);
    if (plan == 0) return PB200_OK;
    if (census_workspace_bytes < pb200_census_sgm_workspace_bytes(H, W, window, dmin, D)) {
        set_error("pb200_census_sgm: census workspace too small (pb200_census_sgm_workspace_bytes)");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int pitch = census_pitch(W);
    uint32_t *descL = (uint32_t *)d_census_workspace;
    uint32_t *descR = descL + (size_t)H * pitch;
    if (!descriptors_ready) {                        // else: pb200_census_sgm_descriptors ran with the same arguments
        const int rc = census_sgm_descriptors(plan, d_left, d_right, H, W, window, dmin, D, d_census_workspace, s);
        if (rc != PB200_OK) return rc;
    }
    CensusDesc desc;
    desc.L = descL; desc.R = descR; desc.pitch = pitch;
    desc.R4 = plan == 1 ? descR + (size_t)H * pitch : nullptr;
    desc.pitch4 = census_pitch4(W, dmin, D); desc.padl = census_padl(dmin);
    bool done = false;
    const int rc = sgm_census_wave_try(desc, window, d_cv_out, H, W, D, p1, p2, overcounting, d_disp, dmin, invalid_disparity,
                                       d_all_nan, d_sgm_workspace, sgm_workspace_bytes, s, &done);
    if (rc != PB200_OK) return rc;
    *ran = done ? 1 : 0;
    return PB200_OK;
}

// A batch of `nimg` pairs (d_left / d_right: (nimg, H, W); d_cv_out: (nimg, H, W, D); d_disp / d_all_nan: (nimg, H, W)) through ONE
// wave per pass of the two-column wavefront kernels: the images are stacked into one tall image whose vertical and diagonal
// paths restart at every image's first row, so the time the wave needs to cross the strips (fill and drain, 1.6 ms of a 14 ms
// stage at 4096 x 4096 x 256) is paid once per batch.  Results are those of `nimg` pb200_census_sgm calls, bit for bit.
// Census workspace: nimg * pb200_census_sgm_workspace_bytes(H, W, ...).  *ran = 0: nothing was computed (not eligible).
extern "C" int pb200_census_sgm_batch(const float *d_left, const float *d_right, int nimg, int H, int W, int window, int dmin, int D, float p1,
                                      float p2, int overcounting, float *d_cv_out, void *d_census_workspace, size_t census_workspace_bytes,
                                      void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp, float invalid_disparity,
                                      uint8_t *d_all_nan, int *ran, void *stream) {
    if (!ran) {
        set_error("pb200_census_sgm_batch: ran must not be NULL");
        return PB200_ERR_BAD_ARG;
    }
    *ran = 0;
    if (!d_left || !d_right || !d_cv_out || !d_census_workspace || !d_sgm_workspace || nimg <= 0 || H <= 0 || W <= 0 || D <= 0) {
        set_error("pb200_census_sgm_batch: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (sgm_census_plan(window, W, D, p1, p2) != 2 || H < 4 || (long)nimg * H > 0x3FFFFFFF) return PB200_OK;
    if (census_workspace_bytes < (size_t)nimg * pb200_census_sgm_workspace_bytes(H, W, window, dmin, D)) {
        set_error("pb200_census_sgm_batch: census workspace too small (nimg * pb200_census_sgm_workspace_bytes)");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int pitch = census_pitch(W);
    uint32_t *descL = (uint32_t *)d_census_workspace;
    uint32_t *descR = descL + (size_t)nimg * H * pitch;
    for (int i = 0; i < nimg; ++i) {                 // every image keeps its own borders: a window never crosses into its neighbour
        const int rc = census_transform_pair(d_left + (size_t)i * H * W, d_right + (size_t)i * H * W, H, W, window,
                                             descL + (size_t)i * H * pitch, descR + (size_t)i * H * pitch, 0, H, s);
        if (rc != PB200_OK) return rc;
    }
    CensusDesc desc;
    desc.L = descL; desc.R = descR; desc.pitch = pitch;
    desc.R4 = nullptr; desc.pitch4 = 0; desc.padl = 0;
    bool done = false;
    const int rc = sgm_census_wave_try(desc, window, d_cv_out, nimg * H, W, D, p1, p2, overcounting, d_disp, dmin, invalid_disparity,
                                       d_all_nan, d_sgm_workspace, sgm_workspace_bytes, s, &done, nullptr, nimg > 1 ? H : 0);
    if (rc != PB200_OK) return rc;
    *ran = done ? 1 : 0;
    return PB200_OK;
}

// ---- sub-pixel Census (census.cpp:128-155: a list of shifted right images) -------------------------------------------------
// Cell k of the volume uses the (k % subpix)-th right image at column x + k / subpix + dmin.  Image 0 is the right image
// itself; image i > 0 is resampled at column offset i / subpix and is one column shorter (img_tools.py:713-752), so its own
// transform flags the centres whose window would leave IT -- exactly the reference's extra bound (census.cpp:144-150).
// One warp per pixel, lanes over consecutive cells: 128-byte coalesced stores, descriptor reads served by L1 / L2.
template <int NW>
__global__ void __launch_bounds__(256) census_fill_subpix_kernel(const uint32_t *__restrict__ descL, const uint32_t *__restrict__ descR,
                                                                 float *__restrict__ cv, int H, int W, int n_disp, int dmin, int subpix,
                                                                 int pitch) {
    const int lane = threadIdx.x & 31;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const size_t plane = (size_t)H * pitch;                 // words per descriptor word-plane
    for (long pix = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; pix < (long)H * W; pix += nwarps) {
        const int row = (int)(pix / W), col = (int)(pix % W);
        uint32_t lw[NW];
#pragma unroll
        for (int i = 0; i < NW; ++i) lw[i] = descL[i * plane + (size_t)row * pitch + col];
        const bool left_ok = (lw[NW - 1] >> 31) == 0;
        float *out = cv + (size_t)pix * n_disp;
        for (int k = lane; k < n_disp; k += 32) {
            const int kd = k / subpix, id = k - kd * subpix, rx = col + kd + dmin;
            float v = nan_f();
            if (left_ok && rx >= 0 && rx < W) {
                const uint32_t *r = descR + (size_t)id * NW * plane + (size_t)row * pitch + rx;
                uint32_t rw[NW];
#pragma unroll
                for (int i = 0; i < NW; ++i) rw[i] = r[i * plane];
                if ((rw[NW - 1] >> 31) == 0) {
                    int c = 0;
#pragma unroll
                    for (int i = 0; i < NW; ++i) c += __popc(lw[i] ^ rw[i]);
                    v = (float)c;
                }
            }
            out[k] = v;
        }
    }
}

extern "C" size_t pb200_census_subpix_workspace_bytes(int H, int W, int window, int n_right) {
    if (H <= 0 || W <= 0 || window < 3 || n_right < 1) return 0;
    return (size_t)(1 + n_right) * census_nwords(window) * H * census_pitch(W) * sizeof(uint32_t);
}

extern "C" int pb200_census_cost_volume_subpix(const float *d_left, const float *const *d_rights, int n_right, int H, int W, int window,
                                               int dmin, int n_disp, float *d_cv, void *d_workspace, size_t workspace_bytes, void *stream) {
    if (!d_left || !d_rights || n_right < 1 || !d_cv || !d_workspace || H <= 0 || W <= 0 || n_disp <= 0) {
        set_error("pb200_census_cost_volume_subpix: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (window != 3 && window != 5 && window != 7 && window != 9 && window != 11 && window != 13) {
        set_error("pb200_census_cost_volume_subpix: window_size %d not in {3,5,7,9,11,13}", window);
        return PB200_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pb200_census_subpix_workspace_bytes(H, W, window, n_right)) {
        set_error("pb200_census_cost_volume_subpix: workspace too small");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nw = census_nwords(window), pitch = census_pitch(W);
    uint32_t *descL = (uint32_t *)d_workspace, *descR = descL + (size_t)nw * H * pitch;
    int rc = PB200_OK;
#define PB200_T(WIN)                                                                                             \
    case WIN:                                                                                                    \
        rc = launch_transform<WIN>(d_left, H, W, pitch, descL, 0, H, s);                                         \
        for (int i = 0; i < n_right && rc == PB200_OK; ++i) {                                                    \
            if (!d_rights[i]) { set_error("pb200_census_cost_volume_subpix: NULL right image"); return PB200_ERR_BAD_ARG; } \
            /* image i > 0 has W - 1 columns (contiguous); same pitch, so the missing column reads as flagged */   \
            rc = launch_transform<WIN>(d_rights[i], H, i == 0 ? W : W - 1, pitch, descR + (size_t)i * nw * H * pitch, 0, H, s); \
        }                                                                                                        \
        break;
    switch (window) {
        PB200_T(3) PB200_T(5) PB200_T(7) PB200_T(9) PB200_T(11) PB200_T(13)
        default: rc = PB200_ERR_UNSUPPORTED;
    }
#undef PB200_T
    if (rc != PB200_OK) return rc;
    long blocks = ((long)H * W + 7) / 8;
    const long cap = (long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    switch (nw) {
#define PB200_F(N) case N: census_fill_subpix_kernel<N><<<(int)blocks, 256, 0, s>>>(descL, descR, d_cv, H, W, n_disp, dmin, n_right, pitch); break;
        PB200_F(1) PB200_F(2) PB200_F(3) PB200_F(4) PB200_F(6)
#undef PB200_F
        default: set_error("census: unexpected descriptor size"); return PB200_ERR_UNSUPPORTED;
    }
    PB200_LAUNCH_CHECK("census_fill_subpix_kernel");
    note_path(STAGE_CENSUS, PATH_CENSUS_SUBPIX, nw);
    return PB200_OK;
}

// ---- column-tiled multi-GPU runs of the fused stage -----------------------------------------------------------------------
extern "C" size_t pb200_tile_link_bytes(int D) {
    if (D <= 0) return 0;
    // [pass-1 boundary block | pass-2 boundary block | pass-1 credit line | pass-2 credit line], each block 256-byte aligned
    const size_t blk = (sgm_wave1_edge_bytes(D) + 255) & ~(size_t)255;
    return 2 * blk + 512;
}

// left descriptors + four shifted right copies of the image columns [col0, col1) (see census_sgm_descriptors), explicit placement
static int census_tile_descriptors(const float *d_left, const float *d_right, int H, int W, int window, int dmin, int D, uint32_t *descL,
                                   uint32_t *desc4, int col0, int col1, cudaStream_t s) {
    const int pitch = census_pitch(W), pitch4 = census_pitch4(W, dmin, D), padl = census_padl(dmin);
    int rc = window == 3 ? launch_transform<3>(d_left, H, W, pitch, descL, 0, H, s, col0, col1)
                         : launch_transform<5>(d_left, H, W, pitch, descL, 0, H, s, col0, col1);
    if (rc != PB200_OK) return rc;
    int i0 = col0 + dmin + padl - 4, i1 = col1 + dmin + D + padl + 4;
    if (i0 < 0) i0 = 0;
    if (i1 > pitch4 + 3) i1 = pitch4 + 3;
    if (i1 <= i0) return PB200_OK;
    dim3 grid(ceil_div(i1 - i0, 256), H);
    if (window == 3) census_transform_shifted_kernel<3><<<grid, 256, 0, s>>>(d_right, H, W, pitch4, padl, desc4, i0, i1);
    else census_transform_shifted_kernel<5><<<grid, 256, 0, s>>>(d_right, H, W, pitch4, padl, desc4, i0, i1);
    PB200_LAUNCH_CHECK("census_transform_shifted_kernel");
    return PB200_OK;
}

extern "C" int pb200_census_sgm_tile(const float *d_left, const float *d_right, int nimg, int H, int Wg, int window, int dmin, int D, float p1,
                                     float p2, int overcounting, int tile, int ntiles, float *d_cv_tile, void *d_census_workspace,
                                     size_t census_workspace_bytes, void *d_sgm_workspace, size_t sgm_workspace_bytes, float *d_disp_tile,
                                     float invalid_disparity, uint8_t *d_all_nan_tile, void *link_local, void *link_left, void *link_right,
                                     unsigned epoch, unsigned prev_epoch, unsigned prev_rows, int passes, void *stream) {
    if (!d_left || !d_right || !d_cv_tile || !d_census_workspace || !d_sgm_workspace || !link_local || !link_left || !link_right || H <= 0 ||
        Wg <= 0 || D <= 0 || nimg < 1 || ntiles < 1 || tile < 0 || tile >= ntiles || Wg % ntiles != 0 || (passes & 3) == 0 ||
        (epoch & 0xFFFFu) == 0) {
        set_error("pb200_census_sgm_tile: bad argument (the image width must be a multiple of the number of tiles)");
        return PB200_ERR_BAD_ARG;
    }
    if ((long)nimg * H >= 65535) {
        set_error("pb200_census_sgm_tile: at most 65534 rows per call (row tags share a word with the epoch)");
        return PB200_ERR_UNSUPPORTED;
    }
    const int Wt = Wg / ntiles;
    struct Pin {                                     // the tiles run the skewed wavefront whatever a one-GPU call would pick
        int old;
        Pin() : old(pb200_get_option("sgm.wave_kernel")) { pb200_set_option("sgm.wave_kernel", 1); }
        ~Pin() { pb200_set_option("sgm.wave_kernel", old); }
    } pin;
    if (sgm_census_plan(window, Wt, D, p1, p2) != 1) {
        set_error("pb200_census_sgm_tile: the skewed wavefront does not take this configuration (window 3 / 5, D in {64, 128, 256}, "
                  "small integer penalties, tile width <= 28 columns per SM)");
        return PB200_ERR_UNSUPPORTED;
    }
    if (census_workspace_bytes < (size_t)nimg * pb200_census_sgm_workspace_bytes(H, Wg, window, dmin, D)) {
        set_error("pb200_census_sgm_tile: census workspace too small (nimg x pb200_census_sgm_workspace_bytes of the WHOLE image)");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int pitch = census_pitch(Wg), pitch4 = census_pitch4(Wg, dmin, D);
    // workspace: [nimg][H][pitch] left descriptors, then [nimg][4][H][pitch4] shifted right copies -- indexed by global image column
    uint32_t *descL = (uint32_t *)d_census_workspace, *desc4 = descL + (size_t)nimg * H * pitch;
    if (passes & 1) {
        // only the columns this tile's pixels visit are computed: the sheared tile [tile * Wt, (tile + 1) * Wt) drifts one image
        // column to the left per row (the drift continues through the images of a batch), cyclically
        const long rows = (long)nimg * H;
        const int span = Wt + rows - 1 < Wg ? (int)(Wt + rows - 1) : Wg;
        const int lo = (int)((((long)tile * Wt - (rows - 1)) % Wg + Wg) % Wg);
        for (int im = 0; im < nimg; ++im) {
            const float *l = d_left + (size_t)im * H * Wg, *r = d_right + (size_t)im * H * Wg;
            uint32_t *dl = descL + (size_t)im * H * pitch, *d4 = desc4 + (size_t)im * 4 * H * pitch4;
            int rc;
            if (lo + span <= Wg) rc = census_tile_descriptors(l, r, H, Wg, window, dmin, D, dl, d4, lo, lo + span, s);
            else {
                rc = census_tile_descriptors(l, r, H, Wg, window, dmin, D, dl, d4, lo, Wg, s);
                if (rc == PB200_OK) rc = census_tile_descriptors(l, r, H, Wg, window, dmin, D, dl, d4, 0, lo + span - Wg, s);
            }
            if (rc != PB200_OK) return rc;
        }
    }
    CensusDesc desc;
    desc.L = descL; desc.R = nullptr; desc.pitch = pitch;
    desc.R4 = desc4; desc.pitch4 = pitch4; desc.padl = census_padl(dmin);
    // link buffers: [in pass 1 | in pass 2 | credit pass 1 | credit pass 2].  Pass 1 travels left -> right (data comes from the
    // left tile, credits come from the right one), pass 2 right -> left.
    const size_t blk = (sgm_wave1_edge_bytes(D) + 255) & ~(size_t)255;
    auto at = [](void *base, size_t off) { return reinterpret_cast<unsigned long long *>(reinterpret_cast<char *>(base) + off); };
    Wave1Peers peers;
    peers.in[0] = at(link_local, 0);            peers.in[1] = at(link_local, blk);
    peers.ack_in[0] = at(link_local, 2 * blk);  peers.ack_in[1] = at(link_local, 2 * blk + 256);
    peers.out[0] = at(link_right, 0);           peers.ack_out[0] = at(link_left, 2 * blk);
    peers.out[1] = at(link_left, blk);          peers.ack_out[1] = at(link_right, 2 * blk + 256);
    peers.Wg = Wg;
    // the two passes walk the batch as ONE image of nimg * H rows: the shear (and the pass-2 tile origin) use that height
    const long rows = (long)nimg * H;
    peers.c_off[0] = tile * Wt;
    peers.c_off[1] = (int)((((rows - 1 - (long)(tile + 1) * Wt) % Wg) + Wg) % Wg);
    peers.epoch = epoch & 0xFFFFu;
    peers.prev_epoch = prev_epoch & 0xFFFFu;
    peers.prev_rows = prev_rows;
    peers.passes = passes & 3;
    peers.nimg = nimg;
    bool done = false;
    int rc = sgm_census_wave_try(desc, window, d_cv_tile, H, Wt, D, p1, p2, overcounting, d_disp_tile, dmin, invalid_disparity, d_all_nan_tile,
                                 d_sgm_workspace, sgm_workspace_bytes, s, &done, &peers);
    if (rc != PB200_OK) return rc;
    if (!done) {
        set_error("pb200_census_sgm_tile: the wavefront kernels could not be made co-resident (workspace or shared memory)");
        return PB200_ERR_UNSUPPORTED;
    }
    return PB200_OK;
}

extern "C" size_t pb200_census_workspace_bytes(int H, int W, int window) {
    if (H <= 0 || W <= 0 || window < 3) return 0;
    return 2 * (size_t)census_nwords(window) * H * census_pitch(W) * sizeof(uint32_t);
}

extern "C" int pb200_census_cost_volume(const float *d_left, const float *d_right, int H, int W, int window, int dmin,
                                        int D, float *d_cv, void *d_workspace, size_t workspace_bytes, float *d_disp,
                                        float invalid_disparity, uint8_t *d_all_nan, void *stream) {
    return pb200_census_cost_volume_rows(d_left, d_right, H, W, window, dmin, D, d_cv, d_workspace, workspace_bytes, d_disp,
                                         invalid_disparity, d_all_nan, 0, H, stream);
}

extern "C" int pb200_census_cost_volume_rows(const float *d_left, const float *d_right, int H, int W, int window, int dmin,
                                             int D, float *d_cv, void *d_workspace, size_t workspace_bytes, float *d_disp,
                                             float invalid_disparity, uint8_t *d_all_nan, int row_begin, int row_end, void *stream) {
    if (!d_left || !d_right || !d_cv || !d_workspace || H <= 0 || W <= 0 || D <= 0 || row_begin < 0 || row_end > H || row_begin >= row_end) {
        set_error("pb200_census_cost_volume: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (window != 3 && window != 5 && window != 7 && window != 9 && window != 11 && window != 13) {
        set_error("pb200_census_cost_volume: window_size %d not in {3,5,7,9,11,13}", window);
        return PB200_ERR_UNSUPPORTED;
    }
    if (D > 2048 || (d_disp && D > 65535)) {
        set_error("pb200_census_cost_volume: D=%d above the supported maximum (2048)", D);
        return PB200_ERR_UNSUPPORTED;
    }
    if (workspace_bytes < pb200_census_workspace_bytes(H, W, window)) {
        set_error("pb200_census_cost_volume: workspace too small");
        return PB200_ERR_WORKSPACE;
    }
    cudaStream_t s = (cudaStream_t)stream;
    const int nw = census_nwords(window), pitch = census_pitch(W);
    uint32_t *descL = (uint32_t *)d_workspace;
    uint32_t *descR = descL + (size_t)nw * H * pitch;
    const int rc = census_transform_pair(d_left, d_right, H, W, window, descL, descR, row_begin, row_end, s);
    if (rc != PB200_OK) return rc;

    FillParams p;
    p.descL = descL; p.descR = descR; p.cv = d_cv; p.disp = d_disp; p.all_nan = d_all_nan;
    p.H = H; p.W = W; p.D = D; p.dmin = dmin; p.half = window / 2; p.pitch = pitch;
    // Two fill kernels, same results: the TMA-tiled one (shared-memory staging, bulk stores) and the barrier-free
    // direct one.  Measured (profiles/r1_stage_bench_final2.json): equal at 4096x4096x256 (3.1 ms), the direct one
    // ahead for smaller volumes and whenever the WTA is fused (C1 pipeline 0.25 -> 0.19 ms).
    const bool direct_ok = (D & 3) == 0 && (reinterpret_cast<uintptr_t>(d_cv) & 15) == 0;
    const bool want_direct = option(OPT_CENSUS_DIRECT) >= 0 ? option(OPT_CENSUS_DIRECT) != 0 : (d_disp != nullptr || D < 256);
    if (direct_ok && want_direct) {
        p.TX = 4; p.CH = 4; p.tiles_x = 0; p.n_tiles = 0; p.r_len = 0;
        p.row0 = row_begin;
        p.invalid_disparity = invalid_disparity;
        note_path(STAGE_CENSUS, PATH_CENSUS_DIRECT, nw);
        switch (nw) {
            case 1: return launch_fill_direct<1>(p, row_end - row_begin, s);
            case 2: return launch_fill_direct<2>(p, row_end - row_begin, s);
            case 3: return launch_fill_direct<3>(p, row_end - row_begin, s);
            case 4: return launch_fill_direct<4>(p, row_end - row_begin, s);
            case 6: return launch_fill_direct<6>(p, row_end - row_begin, s);
            default: set_error("census: unexpected descriptor size"); return PB200_ERR_UNSUPPORTED;
        }
    }
    int TX = (8192 / D) & ~3;                       // <= 32 KB of float per tile
    if (option(OPT_CENSUS_TILE) > 0) TX = (option(OPT_CENSUS_TILE) / D) & ~3;
    if (TX < 4) TX = 4;
    if (TX > 128) TX = 128;
    while (TX > 4 && TX >= 2 * (((W + 3) & ~3))) TX >>= 1;   // do not make tiles much wider than the image
    TX &= ~3;
    p.TX = TX;
    const int G = (D + 3) / 4;
    int CH = TX;                                    // largest chunk that still gives >= 256 work items
    while (CH > 1 && (TX / CH) * G < 256) CH >>= 1;
    while (TX % CH) CH >>= 1;
    p.CH = CH;
    p.tiles_x = ceil_div(W, TX);
    p.row0 = row_begin;
    p.n_tiles = (long)p.tiles_x * (row_end - row_begin);
    p.r_len = (TX + D + 3 + 3 + 3) & ~3;            // shift (<=3) + TX + D - 1 + window slack, rounded to 4
    p.invalid_disparity = invalid_disparity;
    // the left-descriptor bulk copy reads TX words from x0: keep it inside the pitch
    if ((long)(p.tiles_x) * TX > pitch) {
        // shrink TX until the last tile fits (pitch >= roundup4(W)+4, so TX=4 always fits)
        while (p.TX > 4 && (long)ceil_div(W, p.TX) * p.TX > pitch) p.TX -= 4;
        TX = p.TX;
        CH = TX;
        while (CH > 1 && (TX / CH) * G < 256) CH >>= 1;
        while (TX % CH) CH >>= 1;
        p.CH = CH;
        p.tiles_x = ceil_div(W, TX);
        p.n_tiles = (long)p.tiles_x * (row_end - row_begin);
        p.r_len = (TX + D + 9) & ~3;
    }
    const size_t tile_elems = ((size_t)TX * D + 3) & ~(size_t)3;
    const size_t smem = 2 * tile_elems * 4 + 2 * ((size_t)nw * p.r_len * 4 + (size_t)nw * TX * 4) + 32;
    int per_sm = (int)(200 * 1024 / smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 3) per_sm = 3;
    long grid = (long)sm_count() * per_sm;
    if (grid > p.n_tiles) grid = p.n_tiles;
    note_path(STAGE_CENSUS, PATH_CENSUS_TMA, nw);
    switch (nw) {
        case 1: return launch_fill<1>(p, smem, (int)grid, s);
        case 2: return launch_fill<2>(p, smem, (int)grid, s);
        case 3: return launch_fill<3>(p, smem, (int)grid, s);
        case 4: return launch_fill<4>(p, smem, (int)grid, s);
        case 6: return launch_fill<6>(p, smem, (int)grid, s);
        default: set_error("census: unexpected descriptor size"); return PB200_ERR_UNSUPPORTED;
    }
}
