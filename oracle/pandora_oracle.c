/*
 * oracle/pandora_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain single-threaded C restatement of the CPU algorithms on Pandora's dense cost-volume hot
 * path.  It is the checker the CUDA kernels are compared with; only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load it.  The product
 * (pandora_b200/) never imports, links or calls anything in this directory.
 *
 * Each function cites the reference file:line it restates (paths relative to /root/reference).
 * Pinning status (see tests/test_oracle_goldens.py, tests/test_oracle_vs_ref.py):
 *   census / cross_support / cbca : pinned to the reference's golden vectors AND to the compiled
 *                                   reference C++ (oracle/_ref) on randomised inputs, bit-exact.
 *   reverse_cost_volume           : pinned to the reference's C++ doctest vectors and oracle/_ref.
 *   sgm                           : PARITY UNPINNED.  The arithmetic lives in the un-vendored
 *                                   dependency pandora_plugin_libsgm==1.5.7 (pyproject.toml:59-61);
 *                                   this file restates Hirschmueller's published 8-path recurrence
 *                                   with the semantic choices listed above pbo_sgm().
 *
 * Build: make -C oracle   ->  oracle/_build/libpandora_oracle.so
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PBO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------ */
/* Census transform: src/pandora/matching_cost/cpp/src/census.cpp:45-95                         */
/* bit b (row-major over the w x w window, centre included) = neighbour > centre (strict).     */
/* Only Hamming distances are observable so the packing (64-bit words here) is free.           */
/* Border pixels (closer than half window to an edge) keep an all-zero descriptor.             */
/* ------------------------------------------------------------------------------------------ */
static int census_words(int w) { return (w * w + 63) / 64; }

static void census_transform(const float *img, int H, int W, int w, uint64_t *out) {
    const int half = w / 2, nw = census_words(w);
    memset(out, 0, (size_t)H * W * nw * sizeof(uint64_t));
    for (int y = half; y < H - half; ++y) {
        for (int x = half; x < W - half; ++x) {
            const float c = img[(size_t)y * W + x];
            uint64_t *dst = out + ((size_t)y * W + x) * nw;
            int b = 0;
            for (int wy = y - half; wy <= y + half; ++wy)
                for (int wx = x - half; wx <= x + half; ++wx, ++b)
                    if (img[(size_t)wy * W + wx] > c) dst[b >> 6] |= (uint64_t)1 << (b & 63);
        }
    }
}

/* Census matching cost: census.cpp:97-180 (subpix == 1 branch) + the NaN pre-fill of
 * matching_cost/census.py:138.  cv is (H, W, D) float32, disparity fastest.
 * cv[y,x,k] = popcount(cL[y,x] ^ cR[y,x+dmin+k]) iff the left centre and the right centre are both
 * at least `half` away from every image edge; everything else is NaN.                          */
PBO_API int pbo_census_cost(const float *left, const float *right, int H, int W, int w, int dmin, int D,
                            float *cv) {
    if (w < 1 || (w & 1) == 0 || w > 13) return -1;
    const int half = w / 2, nw = census_words(w);
    const size_t n = (size_t)H * W * D;
    for (size_t i = 0; i < n; ++i) cv[i] = NAN;
    uint64_t *cl = (uint64_t *)malloc((size_t)H * W * nw * sizeof(uint64_t));
    uint64_t *cr = (uint64_t *)malloc((size_t)H * W * nw * sizeof(uint64_t));
    if (!cl || !cr) { free(cl); free(cr); return -2; }
    census_transform(left, H, W, w, cl);
    census_transform(right, H, W, w, cr);
    for (int y = half; y < H - half; ++y)
        for (int x = half; x < W - half; ++x) {
            const uint64_t *a = cl + ((size_t)y * W + x) * nw;
            float *dst = cv + ((size_t)y * W + x) * D;
            for (int k = 0; k < D; ++k) {
                const int xr = x + dmin + k;
                if (xr < half || xr >= W - half) continue;
                const uint64_t *b = cr + ((size_t)y * W + xr) * nw;
                int cost = 0;
                for (int i = 0; i < nw; ++i) cost += __builtin_popcountll(a[i] ^ b[i]);
                dst[k] = (float)cost;
            }
        }
    free(cl); free(cr);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Cross support arms: src/pandora/aggregation/cpp/src/aggregation.cpp:224-321                  */
/* out is (H, W, 4) int16 in the order (left, right, up, bottom).                               */
/* ------------------------------------------------------------------------------------------ */
PBO_API void pbo_cross_support(const float *img, int H, int W, int len_arms, float intensity, int16_t *out) {
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            int16_t *o = out + ((size_t)y * W + x) * 4;
            const float c = img[(size_t)y * W + x];
            if (!isfinite(c)) { o[0] = o[1] = o[2] = o[3] = 0; continue; }
            int l = 0, r = 0, u = 0, b = 0;
            for (int q = x - 1; q > x - len_arms && q >= 0; --q) {
                if (fabsf(c - img[(size_t)y * W + q]) >= intensity) break;
                ++l;
            }
            for (int q = x + 1; q < x + len_arms && q < W; ++q) {
                if (fabsf(c - img[(size_t)y * W + q]) >= intensity) break;
                ++r;
            }
            for (int q = y - 1; q > y - len_arms && q >= 0; --q) {
                if (fabsf(c - img[(size_t)q * W + x]) >= intensity) break;
                ++u;
            }
            for (int q = y + 1; q < y + len_arms && q < H; ++q) {
                if (fabsf(c - img[(size_t)q * W + x]) >= intensity) break;
                ++b;
            }
            /* minimum support of one pixel per arm when the direct neighbour is a finite pixel */
            if (l < 1 && x >= 1 && isfinite(img[(size_t)y * W + x - 1])) l = 1;
            if (r < 1 && x < W - 1 && isfinite(img[(size_t)y * W + x + 1])) r = 1;
            if (u < 1 && y >= 1 && isfinite(img[(size_t)(y - 1) * W + x])) u = 1;
            if (b < 1 && y < H - 1 && isfinite(img[(size_t)(y + 1) * W + x])) b = 1;
            o[0] = (int16_t)l; o[1] = (int16_t)r; o[2] = (int16_t)u; o[3] = (int16_t)b;
        }
}

/* ------------------------------------------------------------------------------------------ */
/* CBCA for ONE disparity slice: aggregation.cpp:28-221 (cbca_step_1..4) driven by             */
/* aggregation/cbca.py:152-171.  `cost` is a strided (H, W) view: element (y,x) at               */
/* cost[y*row_stride + x*col_stride].  Valid columns are those with 0 <= x + d < W.             */
/* Returns step4 (sum of costs over the cross) and sum4 (support size WITHOUT the +1 anchor,    */
/* exactly like the reference's C++; the driver adds the 1).  float32 prefix sums like the      */
/* reference, the "col - left - 1 == -1" read is the zero pad column (aggregation.cpp:113-114). */
/* ------------------------------------------------------------------------------------------ */
PBO_API int pbo_cbca_slice(const float *cost, long row_stride, long col_stride, int H, int W, int d,
                           const int16_t *cross_l, const int16_t *cross_r, float *step4, float *sum4) {
    float *s1 = (float *)calloc((size_t)H * (W + 1), sizeof(float));      /* horizontal prefix, col 0 = pad */
    float *s2 = (float *)calloc((size_t)H * W, sizeof(float));
    float *n2 = (float *)calloc((size_t)H * W, sizeof(float));
    float *s3 = (float *)calloc((size_t)(H + 1) * W, sizeof(float));      /* vertical prefix, row 0 = pad */
    if (!s1 || !s2 || !n2 || !s3) { free(s1); free(s2); free(n2); free(s3); return -2; }
    for (int y = 0; y < H; ++y) {
        float acc = 0.f;
        for (int x = 0; x < W; ++x) {
            const float v = cost[y * row_stride + x * col_stride];
            if (!isnan(v)) acc = acc + v;
            s1[(size_t)y * (W + 1) + x + 1] = acc;
        }
    }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int xr = x + d;
            if (xr < 0 || xr >= W) continue;
            const int16_t *a = cross_l + ((size_t)y * W + x) * 4;
            const int16_t *b = cross_r + ((size_t)y * W + xr) * 4;
            const int l = a[0] < b[0] ? a[0] : b[0];
            const int r = a[1] < b[1] ? a[1] : b[1];
            s2[(size_t)y * W + x] = s1[(size_t)y * (W + 1) + x + r + 1] - s1[(size_t)y * (W + 1) + x - l];
            n2[(size_t)y * W + x] = (float)(l + r);
        }
    for (int x = 0; x < W; ++x) s3[(size_t)W + x] = s2[x];
    for (int y = 1; y < H; ++y)
        for (int x = 0; x < W; ++x) s3[(size_t)(y + 1) * W + x] = s3[(size_t)y * W + x] + s2[(size_t)y * W + x];
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            step4[(size_t)y * W + x] = 0.f;
            sum4[(size_t)y * W + x] = n2[(size_t)y * W + x];
        }
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int xr = x + d;
            if (xr < 0 || xr >= W) continue;
            const int16_t *a = cross_l + ((size_t)y * W + x) * 4;
            const int16_t *b = cross_r + ((size_t)y * W + xr) * 4;
            const int t = a[2] < b[2] ? a[2] : b[2];
            const int bo = a[3] < b[3] ? a[3] : b[3];
            step4[(size_t)y * W + x] = s3[(size_t)(y + bo + 1) * W + x] - s3[(size_t)(y - t) * W + x];
            float n = sum4[(size_t)y * W + x] + (float)(t + bo);
            if (t > 0) { float s = 0.f; for (int i = 1; i <= t; ++i) s += n2[(size_t)(y - i) * W + x]; n += s; }
            if (bo > 0) { float s = 0.f; for (int i = 1; i <= bo; ++i) s += n2[(size_t)(y + i) * W + x]; n += s; }
            sum4[(size_t)y * W + x] = n;
        }
    free(s1); free(s2); free(n2); free(s3);
    return 0;
}

/* Whole-volume CBCA driver: aggregation/cbca.py:127-177 for subpix == 1, on the already cropped
 * (H, W, D) view (contiguous).  out[y,x,k] = (0*c + step4) / (sum4 + 1): NaN where the input is NaN. */
PBO_API int pbo_cbca_volume(const float *cv, int H, int W, int D, int dmin, const int16_t *cross_l,
                            const int16_t *cross_r, float *out) {
    float *step4 = (float *)malloc((size_t)H * W * sizeof(float));
    float *sum4 = (float *)malloc((size_t)H * W * sizeof(float));
    if (!step4 || !sum4) { free(step4); free(sum4); return -2; }
    for (int k = 0; k < D; ++k) {
        int rc = pbo_cbca_slice(cv + k, (long)W * D, D, H, W, dmin + k, cross_l, cross_r, step4, sum4);
        if (rc) { free(step4); free(sum4); return rc; }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const size_t p = (size_t)y * W + x;
                float agg = cv[p * D + k] * 0.f;           /* NaN-preserving zero, cbca.py:146-147 */
                agg = agg + step4[p];
                out[p * D + k] = agg / (sum4[p] + 1.f);
            }
    }
    free(step4); free(sum4);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Right cost volume from the left one: matching_cost/cpp/src/matching_cost.cpp:26-57           */
/* ------------------------------------------------------------------------------------------ */
PBO_API void pbo_reverse_cost_volume(const float *left_cv, int H, int W, int D, int min_disp, float *right_cv) {
    for (int i = 0; i < H; ++i)
        for (int j = 0; j < W; ++j)
            for (int k = 0; k < D; ++k) {
                const int c = j + k + min_disp;
                right_cv[((size_t)i * W + j) * D + k] =
                    (c >= 0 && c < W) ? left_cv[((size_t)i * W + c) * D + (D - 1 - k)] : NAN;
            }
}

/* ------------------------------------------------------------------------------------------ */
/* Right disparity grids from the left ones: matching_cost/cpp/src/matching_cost.cpp:59-131       */
/* (scatter of -d over [int(min), int(max)], NaN bounds skipped, unseen right pixels -> NaN)      */
/* ------------------------------------------------------------------------------------------ */
PBO_API void pbo_reverse_disp_range(const float *left_min, const float *left_max, int H, int W, float *right_min, float *right_max) {
    for (size_t i = 0; i < (size_t)H * W; ++i) {
        right_min[i] = INFINITY;
        right_max[i] = -INFINITY;
    }
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const float a = left_min[(size_t)row * W + col], b = left_max[(size_t)row * W + col];
            if (isnan(a) || isnan(b)) continue;
            for (int d = (int)a; d <= (int)b; ++d) {
                const int rc = col + d;
                if (rc < 0) continue;
                if (rc >= W) break;
                float *mn = right_min + (size_t)row * W + rc, *mx = right_max + (size_t)row * W + rc;
                if ((float)(-d) < *mn) *mn = (float)(-d);
                if ((float)(-d) > *mx) *mx = (float)(-d);
            }
        }
    for (size_t i = 0; i < (size_t)H * W; ++i)
        if (isinf(right_min[i])) right_min[i] = right_max[i] = NAN;
}

/* ------------------------------------------------------------------------------------------ */
/* SGM 8-path regularisation.  PARITY UNPINNED against pandora_plugin_libsgm==1.5.7 / libSGM     */
/* (not vendored, pyproject.toml:59-61; behaviour documented in                                  */
/* docs/source/userguide/plugins/plugin_libsgm.rst:9-146, boundary optimization/optimization.py  */
/* :104-123, call site state_machine.py:415-419).  Semantics fixed here:          */
/*  (i)   NaN costs are replaced by `invalid_value` (the plugin wrapper passes cmax + P2 + 1)    */
/*        and take part in the recurrence as ordinary finite costs; NaN is restored at the end.  */
/*  (ii)  first pixel of a path: L_r = C.                                                        */
/*  (iii) 8 directions, accumulated into S in the order E, W, S, SE, SW, N, NE, NW.              */
/*  (iv)  step (Hirschmueller 2008, eq. 13), all float32, evaluated in exactly this order:       */
/*          m  = min_k Lp[k]                                                                     */
/*          t  = min(Lp[d], min(Lp[d-1], Lp[d+1]) + P1)     (missing neighbour skipped)          */
/*          t  = min(t, m + P2)                                                                  */
/*          L[d] = C[d] + (t - m)                                                                */
/*  (v)   overcounting != 0: S -= (n_dirs - 1) * C  (rst:101-107).                               */
/*  type_measure == "max" volumes are negated by the caller before and after.                    */
/* ------------------------------------------------------------------------------------------ */
static void sgm_step(const float *C, const float *Lp, float *L, int D, float P1, float P2) {
    float m = Lp[0];
    for (int k = 1; k < D; ++k) m = fminf(m, Lp[k]);
    const float mp2 = m + P2;
    for (int d = 0; d < D; ++d) {
        float t = Lp[d];
        if (D > 1) {
            float nb;
            if (d == 0) nb = Lp[1];
            else if (d == D - 1) nb = Lp[D - 2];
            else nb = fminf(Lp[d - 1], Lp[d + 1]);
            t = fminf(t, nb + P1);
        }
        t = fminf(t, mp2);
        L[d] = C[d] + (t - m);
    }
}

static const int SGM_DIRS[8][2] = {{0, 1}, {0, -1}, {1, 0}, {1, 1}, {1, -1}, {-1, 0}, {-1, 1}, {-1, -1}};

/* nb_dirs (optional, (H, W)): min_cost_paths of the plugin (plugin_libsgm.rst:411-413), "the number of sgm paths that give the
 * same position for minimal optimized cost at each point": how many directions have the (first) minimum of their own L_r at the
 * disparity where the sum is minimal over the valid cells (first minimum); 0 for a pixel without any valid cost. */
static int sgm_impl(const float *cv_in, int H, int W, int D, float P1, float P2, float invalid_value,
                    int overcounting, int n_dirs, float *cv_out, float *nb_dirs) {
    if (n_dirs < 1 || n_dirs > 8) return -1;
    int *dirmin = nb_dirs ? (int *)malloc((size_t)n_dirs * H * W * sizeof(int)) : NULL;
    if (nb_dirs && !dirmin) return -2;
    const size_t n = (size_t)H * W * D;
    float *C = (float *)malloc(n * sizeof(float));
    float *bufa = (float *)malloc((size_t)D * sizeof(float));
    float *bufb = (float *)malloc((size_t)D * sizeof(float));
    if (!C || !bufa || !bufb) { free(C); free(bufa); free(bufb); return -2; }
    for (size_t i = 0; i < n; ++i) { C[i] = isnan(cv_in[i]) ? invalid_value : cv_in[i]; cv_out[i] = 0.f; }
    for (int r = 0; r < n_dirs; ++r) {
        const int dy = SGM_DIRS[r][0], dx = SGM_DIRS[r][1];
        for (int y0 = 0; y0 < H; ++y0)
            for (int x0 = 0; x0 < W; ++x0) {
                const int py = y0 - dy, px = x0 - dx;
                if (py >= 0 && py < H && px >= 0 && px < W) continue;     /* not a path start */
                float *Lp = bufa, *L = bufb;
                int y = y0, x = x0, first = 1;
                while (y >= 0 && y < H && x >= 0 && x < W) {
                    const float *c = C + ((size_t)y * W + x) * D;
                    float *s = cv_out + ((size_t)y * W + x) * D;
                    if (first) { memcpy(L, c, (size_t)D * sizeof(float)); first = 0; }
                    else sgm_step(c, Lp, L, D, P1, P2);
                    for (int d = 0; d < D; ++d) s[d] = s[d] + L[d];
                    if (dirmin) {
                        int best = 0;
                        for (int d = 1; d < D; ++d)
                            if (L[d] < L[best]) best = d;
                        dirmin[((size_t)r * H + y) * W + x] = best;
                    }
                    float *tmp = Lp; Lp = L; L = tmp;
                    y += dy; x += dx;
                }
            }
    }
    for (size_t i = 0; i < n; ++i) {
        if (overcounting) cv_out[i] = cv_out[i] - (float)(n_dirs - 1) * C[i];
        if (isnan(cv_in[i])) cv_out[i] = NAN;
    }
    if (nb_dirs)
        for (size_t p = 0; p < (size_t)H * W; ++p) {
            const float *s = cv_out + p * D;
            int best = -1;
            for (int d = 0; d < D; ++d)
                if (!isnan(s[d]) && (best < 0 || s[d] < s[best])) best = d;
            int c = 0;
            for (int r = 0; r < n_dirs && best >= 0; ++r) c += dirmin[(size_t)r * H * W + p] == best;
            nb_dirs[p] = (float)c;
        }
    free(C); free(bufa); free(bufb); free(dirmin);
    return 0;
}

PBO_API int pbo_sgm(const float *cv_in, int H, int W, int D, float P1, float P2, float invalid_value,
                    int overcounting, int n_dirs, float *cv_out) {
    return sgm_impl(cv_in, H, W, D, P1, P2, invalid_value, overcounting, n_dirs, cv_out, NULL);
}

PBO_API int pbo_sgm_min_cost_paths(const float *cv_in, int H, int W, int D, float P1, float P2, float invalid_value,
                                   int overcounting, float *cv_out, float *nb_dirs) {
    return sgm_impl(cv_in, H, W, D, P1, P2, invalid_value, overcounting, 8, cv_out, nb_dirs);
}

/* One direction of pbo_sgm on a row tile, with the path-state hand-over used by row-tiled (multi-GPU)
 * runs: S (H, W, D) is accumulated in place (S += L, or S = L when init != 0); C must already hold
 * invalid_value instead of NaN.  halo_in (W, D) holds the L vectors of the row just outside the tile
 * (above for dy > 0, below for dy < 0) or is NULL; halo_out (W, D) receives the tile's last row. */
PBO_API int pbo_sgm_direction(const float *C, float *S, int H, int W, int D, float P1, float P2, int dir, int init,
                              const float *halo_in, float *halo_out) {
    if (dir < 0 || dir > 7) return -1;
    const int dy = SGM_DIRS[dir][0], dx = SGM_DIRS[dir][1];
    float *bufa = (float *)malloc((size_t)D * sizeof(float));
    float *bufb = (float *)malloc((size_t)D * sizeof(float));
    if (!bufa || !bufb) { free(bufa); free(bufb); return -2; }
    for (int y0 = 0; y0 < H; ++y0)
        for (int x0 = 0; x0 < W; ++x0) {
            const int py = y0 - dy, px = x0 - dx;
            if (py >= 0 && py < H && px >= 0 && px < W) continue;
            float *Lp = bufa, *L = bufb;
            int y = y0, x = x0, first = 1;
            if (halo_in && dy != 0 && py == (dy > 0 ? -1 : H) && px >= 0 && px < W) {
                memcpy(Lp, halo_in + (size_t)px * D, (size_t)D * sizeof(float));
                first = 0;
            }
            while (y >= 0 && y < H && x >= 0 && x < W) {
                const float *c = C + ((size_t)y * W + x) * D;
                float *s = S + ((size_t)y * W + x) * D;
                if (first) { memcpy(L, c, (size_t)D * sizeof(float)); first = 0; }
                else sgm_step(c, Lp, L, D, P1, P2);
                for (int d = 0; d < D; ++d) s[d] = init ? L[d] : s[d] + L[d];
                if (halo_out && dy != 0 && y == (dy > 0 ? H - 1 : 0)) memcpy(halo_out + (size_t)x * D, L, (size_t)D * sizeof(float));
                float *tmp = Lp; Lp = L; L = tmp;
                y += dy; x += dx;
            }
        }
    free(bufa); free(bufb);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Winner-takes-all: disparity/disparity.py:434-455, 483-553.  First index wins ties, NaN is     */
/* +inf (min) / -inf (max); an all-NaN pixel gets invalid_disparity.  all_nan (optional) gets 1   */
/* for those pixels.                                                                             */
/* ------------------------------------------------------------------------------------------ */
PBO_API void pbo_wta(const float *cv, int H, int W, int D, const float *disp_coord, int is_max,
                     float invalid_disparity, float *disp_map, uint8_t *all_nan) {
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float *c = cv + p * D;
        int best = 0, any = 0;
        float bv = is_max ? -INFINITY : INFINITY;
        for (int k = 0; k < D; ++k) {
            if (isnan(c[k])) continue;
            any = 1;
            if (is_max ? (c[k] > bv) : (c[k] < bv)) { bv = c[k]; best = k; }
        }
        disp_map[p] = any ? disp_coord[best] : invalid_disparity;
        if (all_nan) all_nan[p] = (uint8_t)!any;
    }
}

/* ========================================================================================== */
/* SURVEY.md 8(f) "next" rows: sub-pixel refinement, cost-volume confidence                     */
/* All pinned to the reference's golden vectors (tests/test_refinement.py, tests/test_confidence) */
/* and to the compiled reference C++ (oracle/_ref: refinement_cpp, cost_volume_confidence_cpp). */
/* ========================================================================================== */

#define PBO_MSK_INVALID 0x3C3               /* src/pandora/constants.py:28 */
#define PBO_MSK_STOPPED_INTERPOLATION 8     /* src/pandora/constants.py:36 */

/* validate_costs_and_get_variables: src/pandora/refinement/cpp/src/refinement_tools.cpp:25-56 */
static int refine_valid(float c0, float c1, float c2, int is_max, float *ic0, float *ic2) {
    if (isnan(c0) || isnan(c2)) return 0;
    const float inverse = is_max ? -1.f : 1.f;
    const float i0 = inverse * c0, i1 = inverse * c1, i2 = inverse * c2;
    if (i1 > i0 || i1 > i2) return 0;
    *ic0 = i0;
    *ic2 = i2;
    return 1;
}

/* vfit_refinement_method: refinement/cpp/src/vfit.cpp:28-55;  quadratic_refinement_method:
 * refinement/cpp/src/quadratic.cpp:28-49.  method 0 = vfit, 1 = quadratic.  Returns the mask increment. */
static int refine_method(int method, float c0, float c1, float c2, int is_max, float *sub_disp, float *sub_cost) {
    float ic0 = 0.f, ic2 = 0.f;
    if (!refine_valid(c0, c1, c2, is_max, &ic0, &ic2)) {
        *sub_disp = 0.f;
        *sub_cost = c1;
        return PBO_MSK_STOPPED_INTERPOLATION;
    }
    if (method == 0) {
        const float a = ic0 > ic2 ? c0 - c1 : c2 - c1;
        if (fabs((double)a) < 1.0e-15) {
            *sub_disp = 0.f;
            *sub_cost = c1;
            return 0;
        }
        const float sd = (c0 - c2) / (2 * a);
        *sub_disp = sd;
        *sub_cost = a * (sd - 1) + c2;
        return 0;
    }
    const float alpha = (c0 - 2.f * c1 + c2) / 2.f;
    const float beta = (c2 - c0) / 2.f;
    const float q = -beta / (2.f * alpha);
    const float lo = (-1.f < q) ? q : -1.f;          /* std::max(-1.f, q) */
    const float sd = (lo < 1.f) ? lo : 1.f;          /* std::min(1.f, lo) */
    *sub_disp = sd;
    *sub_cost = (alpha * sd * sd) + (beta * sd) + c1;
    return 0;
}

/* loop_refinement (approx == 0): refinement/cpp/src/refinement.cpp:29-106;
 * loop_approximate_refinement (approx != 0): refinement.cpp:109-181.
 * disp (H, W) float32 and mask (H, W) uint16 are updated in place, itp (H, W) receives the interpolated costs.
 * Cells the reference would read out of bounds (undefined behaviour there) give itp = NaN and leave the pixel. */
PBO_API void pbo_refinement(const float *cv, int H, int W, int D, double d_min, double d_max, int subpix, int is_max,
                            int method, int approx, float *disp, uint16_t *mask, float *itp) {
    for (int row = 0; row < H; ++row)
        for (int col = 0; col < W; ++col) {
            const size_t i = (size_t)row * W + col;
            if ((mask[i] & PBO_MSK_INVALID) != 0) { itp[i] = NAN; continue; }
            const float raw = disp[i];
            int dsp, diag = col;
            if (!approx) dsp = (int)((raw - d_min) * subpix);
            else { dsp = (int)((-raw - d_min) * subpix); diag = (int)((float)col + raw); }
            if (dsp < 0 || dsp >= D || diag < 0 || diag >= W) { itp[i] = NAN; continue; }
            const float *pc = cv + ((size_t)row * W + diag) * D;
            const float c1 = pc[dsp];
            if (isnan(c1)) { itp[i] = c1; continue; }
            if (raw == d_min || raw == d_max || (approx && (diag == 0 || diag == W - 1))) {
                itp[i] = c1;
                mask[i] = (uint16_t)(mask[i] + PBO_MSK_STOPPED_INTERPOLATION);
                continue;
            }
            float c0, c2;
            if (!approx) {
                if (dsp - 1 < 0 || dsp + 1 >= D) { itp[i] = NAN; continue; }
                c0 = pc[dsp - 1]; c2 = pc[dsp + 1];
            } else {
                if (dsp + subpix >= D || dsp - subpix < 0) { itp[i] = NAN; continue; }
                c0 = pc[-D + dsp + subpix]; c2 = pc[D + dsp - subpix];
            }
            float sd, sc;
            const int flag = refine_method(method, c0, c1, c2, is_max, &sd, &sc);
            disp[i] = raw + sd / (float)subpix;
            itp[i] = sc;
            mask[i] = (uint16_t)(mask[i] + flag);
        }
}

/* searchsorted: cost_volume_confidence/cpp/src/cost_volume_confidence_tools.cpp:22-38 */
static size_t pbo_searchsorted(const float *arr, int n, float value) {
    size_t left = 0, right = (size_t)n - 1;
    while (left < right) {
        const size_t mid = left + (right - left) / 2;
        if (arr[mid] < value) left = mid + 1; else right = mid;
    }
    return left;
}

/* min_max_cost: cost_volume_confidence_tools.cpp:40-87 (per-pixel minimum image + global extrema) */
static void pbo_min_max(const float *cv, int H, int W, int D, float *min_img, float *gmin, float *gmax) {
    float mn = INFINITY, mx = -INFINITY;
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        float pmin = INFINITY, pmax = -INFINITY;
        int all_nan = 1;
        for (int k = 0; k < D; ++k) {
            const float v = cv[p * D + k];
            if (!isnan(v)) { all_nan = 0; if (v < pmin) pmin = v; if (v > pmax) pmax = v; }
        }
        if (all_nan) { min_img[p] = NAN; continue; }
        min_img[p] = pmin;
        if (pmin < mn) mn = pmin;
        if (pmax > mx) mx = pmax;
    }
    *gmin = mn; *gmax = mx;
}

/* normalised costs of one pixel with the +-inf convention of ambiguity.cpp:97-116 / risk.cpp:113-126 */
static void pbo_normalise(const float *pc, int D, float gmin, float diff, size_t imin, size_t imax, float *out) {
    for (int k = 0; k < D; ++k) {
        const float v = pc[k];
        if (isnan(v)) out[k] = ((size_t)k >= imin && (size_t)k < imax) ? -INFINITY : INFINITY;
        else out[k] = (v - gmin) / diff;
    }
}

/* compute_ambiguity_and_sampled_ambiguity: cost_volume_confidence/cpp/src/ambiguity.cpp:28-142.
 * grids (2, H, W) int32 = per-pixel [disp_min, disp_max]; etas float32; samp (H, W, n_etas) may be NULL. */
PBO_API void pbo_ambiguity(const float *cv, int H, int W, int D, const float *etas, int n_etas, const int32_t *grids,
                           const float *disparity_range, float *amb, float *samp) {
    float *min_img = (float *)malloc((size_t)H * W * sizeof(float));
    float *norm = (float *)malloc((size_t)D * sizeof(float));
    float gmin, gmax;
    pbo_min_max(cv, H, W, D, min_img, &gmin, &gmax);
    const float diff = gmax - gmin;
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float ext = (min_img[p] - gmin) / diff;
        if (isnan(ext)) {
            amb[p] = (float)(n_etas * D);
            if (samp) for (int e = 0; e < n_etas; ++e) samp[p * n_etas + e] = (float)D;
            continue;
        }
        const size_t imin = pbo_searchsorted(disparity_range, D, (float)grids[p]);
        const size_t imax = pbo_searchsorted(disparity_range, D, (float)grids[(size_t)H * W + p]) + 1;
        pbo_normalise(cv + p * D, D, gmin, diff, imin, imax, norm);
        float amb_sum = 0;
        for (int e = 0; e < n_etas; ++e) {
            float s = 0;
            for (int k = 0; k < D; ++k) s += (norm[k] <= (ext + etas[e])) ? 1.f : 0.f;
            amb_sum += s;
            if (samp) samp[p * n_etas + e] = s;
        }
        amb[p] = amb_sum;
    }
    free(min_img);
    free(norm);
}

/* compute_risk_and_sampled_risk: cost_volume_confidence/cpp/src/risk.cpp:28-197.  etas are DOUBLE here
 * (risk.hpp takes array_t<double>), so the threshold comparison happens in double precision. */
PBO_API void pbo_risk(const float *cv, const float *samp_amb, int H, int W, int D, const double *etas, int n_etas,
                      const int32_t *grids, const float *disparity_range, float *risk_max, float *risk_min,
                      float *disp_sup, float *disp_inf, float *samp_risk_max, float *samp_risk_min) {
    float *min_img = (float *)malloc((size_t)H * W * sizeof(float));
    float *norm = (float *)malloc((size_t)D * sizeof(float));
    float gmin, gmax;
    pbo_min_max(cv, H, W, D, min_img, &gmin, &gmax);
    const float diff = gmax - gmin;
    for (size_t p = 0; p < (size_t)H * W; ++p) {
        const float ext = (min_img[p] - gmin) / diff;
        if (isnan(ext)) {
            risk_min[p] = risk_max[p] = disp_inf[p] = disp_sup[p] = NAN;
            if (samp_risk_max) for (int e = 0; e < n_etas; ++e) samp_risk_min[p * n_etas + e] = samp_risk_max[p * n_etas + e] = NAN;
            continue;
        }
        const size_t imin = pbo_searchsorted(disparity_range, D, (float)grids[p]);
        const size_t imax = pbo_searchsorted(disparity_range, D, (float)grids[(size_t)H * W + p]) + 1;
        pbo_normalise(cv + p * D, D, gmin, diff, imin, imax, norm);
        float s_min = 0, s_max = 0, s_inf = 0, s_sup = 0;
        for (int e = 0; e < n_etas; ++e) {
            float lo = INFINITY, hi = -INFINITY;
            for (int k = 0; k < D; ++k) {
                if (norm[k] > (ext + etas[e])) continue;
                if ((float)k < lo) lo = (float)k;
                if ((float)k > hi) hi = (float)k;
            }
            const float dlo = disparity_range[(int)lo], dhi = disparity_range[(int)hi];
            const float e_max = hi - lo;
            const float e_min = 1 + e_max - samp_amb[p * n_etas + e];
            s_sup += dhi; s_inf += dlo; s_min += e_min; s_max += e_max;
            if (samp_risk_max) { samp_risk_min[p * n_etas + e] = e_min; samp_risk_max[p * n_etas + e] = e_max; }
        }
        risk_min[p] = s_min / n_etas; risk_max[p] = s_max / n_etas;
        disp_sup[p] = s_sup / n_etas; disp_inf[p] = s_inf / n_etas;
    }
    free(min_img);
    free(norm);
}
