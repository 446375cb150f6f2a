// refinement.cu -- sub-pixel refinement of a disparity map from the cost volume (vfit / quadratic).
//
// Replaces loop_refinement / loop_approximate_refinement (src/pandora/refinement/cpp/src/refinement.cpp:29-181)
// with the per-pixel methods vfit_refinement_method (refinement/cpp/src/vfit.cpp:28-55) and
// quadratic_refinement_method (refinement/cpp/src/quadratic.cpp:28-49); cost validation as in
// refinement/cpp/src/refinement_tools.cpp:25-56.  One thread per pixel: the step reads the disparity, the mask and
// three cells of the volume around the winner (12 B + 3 sectors per pixel -- O(H*W), not O(H*W*D)).  All arithmetic
// is float32 in the reference's operation order (the library is built with -fmad=false), so results are bit-identical.
#include "common.cuh"

namespace pb200 {

namespace {

constexpr uint16_t MSK_INVALID = 0x3C3;                 // constants.py:28
constexpr uint16_t MSK_STOPPED_INTERPOLATION = 1 << 3;  // constants.py:36

// returns the validity-mask increment; sub_disp / sub_cost as the reference's tuple
__device__ __forceinline__ int refine_pixel(int method, float c0, float c1, float c2, bool is_max, float &sub_disp, float &sub_cost) {
    const float inverse = is_max ? -1.f : 1.f;
    const float ic0 = inverse * c0, ic1 = inverse * c1, ic2 = inverse * c2;
    const bool valid = !(c0 != c0 || c2 != c2) && !(ic1 > ic0 || ic1 > ic2);
    if (!valid) {
        sub_disp = 0.f;
        sub_cost = c1;
        return MSK_STOPPED_INTERPOLATION;
    }
    if (method == 0) {                                   // vfit
        const float a = ic0 > ic2 ? c0 - c1 : c2 - c1;
        if (fabs((double)a) < 1.0e-15) {
            sub_disp = 0.f;
            sub_cost = c1;
            return 0;
        }
        const float sd = (c0 - c2) / (2.f * a);
        sub_disp = sd;
        sub_cost = a * (sd - 1.f) + c2;
        return 0;
    }
    const float alpha = (c0 - 2.f * c1 + c2) / 2.f;      // quadratic
    const float beta = (c2 - c0) / 2.f;
    const float q = -beta / (2.f * alpha);
    const float lo = (-1.f < q) ? q : -1.f;               // std::max(-1.f, q): NaN -> -1
    const float sd = (lo < 1.f) ? lo : 1.f;               // std::min(1.f, lo)
    sub_disp = sd;
    sub_cost = (alpha * sd * sd) + (beta * sd) + c1;
    return 0;
}

// MODE 0: loop_refinement on the volume itself.  MODE 1: loop_approximate_refinement (right disparities, LEFT volume,
// the reference's own approximations: bounds compared with the left range, diagonal ends stop the interpolation).
// MODE 2: loop_refinement on the REVERSED volume right(i, j, k) = left(i, j + k + d_min, D-1-k) without materialising it
// (what state_machine.py:488-490 does on right_cv in the cross_checking_fast mode; d_min / d_max = right coordinates,
// subpix 1): same cells as MODE 1, exact stop conditions of MODE 0.
template <int MODE>
__global__ void __launch_bounds__(256) refinement_kernel(const float *__restrict__ cv, int H, int W, int D, double d_min, double d_max,
                                                         int subpix, int is_max, int method, float *__restrict__ disp,
                                                         uint16_t *__restrict__ mask, float *__restrict__ itp) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long)H * W) return;
    const int col = (int)(i % W);
    const long row = i / W;
    const uint16_t m = mask[i];
    if ((m & MSK_INVALID) != 0) {                         // no interpolation on invalid points
        itp[i] = nan_f();
        return;
    }
    const float raw = disp[i];
    int dsp, diag = col;
    if (MODE == 0) {
        dsp = (int)(((double)raw - d_min) * subpix);
    } else if (MODE == 1) {
        dsp = (int)((-(double)raw - d_min) * subpix);
        diag = (int)((float)col + raw);                   // position of the best cost in the LEFT volume: cv[row, diag, dsp]
    } else {
        const int k = (int)(((double)raw - d_min) * subpix);   // index in the (virtual) right volume
        dsp = D - 1 - k;
        diag = col + k + (int)d_min;
    }
    // the reference reads without bounds checks (undefined behaviour when a valid pixel carries a disparity outside the
    // volume); here such a pixel gets a NaN coefficient and is left untouched
    if (dsp < 0 || dsp >= D || (MODE == 1 && (diag < 0 || diag >= W))) {
        itp[i] = nan_f();
        return;
    }
    const float *pc = cv + (row * W + diag) * D;
    // MODE 2: columns outside the image are the NaN cells of the reversed volume (matching_cost.cpp:44-52)
    const float c1 = (MODE == 2 && (diag < 0 || diag >= W)) ? nan_f() : pc[dsp];
    if (c1 != c1) {
        itp[i] = c1;
        return;
    }
    if ((double)raw == d_min || (double)raw == d_max || (MODE == 1 && (diag == 0 || diag == W - 1))) {
        itp[i] = c1;                                      // calculations stopped at the pixel step
        mask[i] = (uint16_t)(m + MSK_STOPPED_INTERPOLATION);
        return;
    }
    float c0, c2;
    if (MODE == 0) {
        if (dsp - 1 < 0 || dsp + 1 >= D) { itp[i] = nan_f(); return; }
        c0 = pc[dsp - 1];
        c2 = pc[dsp + 1];
    } else {
        if (dsp - subpix < 0 || dsp + subpix >= D) { itp[i] = nan_f(); return; }
        c0 = (MODE == 2 && diag - 1 < 0) ? nan_f() : pc[-D + dsp + subpix];      // cv[row, diag - 1, dsp + subpix]
        c2 = (MODE == 2 && diag + 1 >= W) ? nan_f() : pc[D + dsp - subpix];      // cv[row, diag + 1, dsp - subpix]
    }
    float sd, sc;
    const int flag = refine_pixel(method, c0, c1, c2, is_max != 0, sd, sc);
    disp[i] = raw + sd / (float)subpix;
    itp[i] = sc;
    mask[i] = (uint16_t)(m + flag);
}

}  // namespace

}  // namespace pb200

using namespace pb200;

extern "C" int pb200_refinement(const float *d_cv, int H, int W, int D, double d_min, double d_max, int subpix, int is_max, int method,
                                int approximate, float *d_disp, uint16_t *d_mask, float *d_itp_coeff, void *stream) {
    if (!d_cv || !d_disp || !d_mask || !d_itp_coeff || H <= 0 || W <= 0 || D <= 0 || subpix <= 0 || !(d_min <= d_max)) {
        set_error("pb200_refinement: bad argument");
        return PB200_ERR_BAD_ARG;
    }
    if (method != 0 && method != 1) {
        set_error("pb200_refinement: method %d is neither 0 (vfit) nor 1 (quadratic)", method);
        return PB200_ERR_UNSUPPORTED;
    }
    const int grid = ceil_div((long)H * W, 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (approximate == 2 && subpix != 1) {
        set_error("pb200_refinement: the right-from-left mode needs subpix == 1");
        return PB200_ERR_UNSUPPORTED;
    }
    if (approximate == 1)
        refinement_kernel<1><<<grid, 256, 0, s>>>(d_cv, H, W, D, d_min, d_max, subpix, is_max, method, d_disp, d_mask, d_itp_coeff);
    else if (approximate == 2)
        refinement_kernel<2><<<grid, 256, 0, s>>>(d_cv, H, W, D, d_min, d_max, subpix, is_max, method, d_disp, d_mask, d_itp_coeff);
    else
        refinement_kernel<0><<<grid, 256, 0, s>>>(d_cv, H, W, D, d_min, d_max, subpix, is_max, method, d_disp, d_mask, d_itp_coeff);
    PB200_LAUNCH_CHECK("refinement_kernel");
    return PB200_OK;
}
