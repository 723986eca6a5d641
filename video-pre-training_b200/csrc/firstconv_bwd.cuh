// Weight / bias gradient of the fused stack-0 first convolution (csrc/firstconv.cuh: u8 -> conv3x3(3->C0)+bias -> ReLU ->
// max_pool 3/2/1).  The forward never materialises the 128x128xC0 pre-pool map, so the backward recomputes it:
//
//   thread = output channel (block = C0 threads); a block walks segments of 8 pooled pixels; the (5 x 19 x 3) u8 input
//   window of a segment is staged in shared memory as floats; per pooled pixel every thread pulls the 5x5x3 window into
//   registers, evaluates the 9 convolution outputs of the pooling window in fp32, picks the first maximum (ReLU: only if
//   it is > 0) and accumulates  dW[k] += g * patch_argmax[k],  db += g  in registers -- K = 27 is far too small for the
//   tensor cores and there is no reduction across threads.  Per-block partials are summed in a fixed order afterwards.
#pragma once
#include "common.cuh"

namespace vpt {

constexpr int kFbSeg = 8;                       // pooled pixels per segment
constexpr int kFbWinCols = 2 * kFbSeg + 3;      // input columns a segment touches
constexpr int kFbWinFloats = 5 * kFbWinCols * 3;

__global__ void __launch_bounds__(256, 1) firstconv_bwd_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                                                const __nv_bfloat16* __restrict__ dy, float* __restrict__ ws, int F, int H, int W, int C0) {
    __shared__ float win[kFbWinFloats];
    const int c = threadIdx.x;  // blockDim.x == C0
    const int Ho = H >> 1, Wo = W >> 1;
    const int segs_per_row = Wo / kFbSeg;
    const long long items = (long long)F * Ho * segs_per_row;
    float wr[27], dW[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        wr[k] = __ldg(w + c * 27 + k);
        dW[k] = 0.f;
    }
    const float bc = __ldg(bias + c);
    float db = 0.f;
    for (long long it = blockIdx.x; it < items; it += gridDim.x) {
        const int seg = (int)(it % segs_per_row);
        const int oy = (int)((it / segs_per_row) % Ho);
        const long long f = it / ((long long)segs_per_row * Ho);
        const int ox0 = seg * kFbSeg;
        const int y0 = 2 * oy - 2, x0 = 2 * ox0 - 2;  // top-left input pixel of the staged window
        __syncthreads();
        for (int e = threadIdx.x; e < kFbWinFloats; e += blockDim.x) {
            const int ch = e % 3, col = (e / 3) % kFbWinCols, r = e / (3 * kFbWinCols);
            const int y = y0 + r, x = x0 + col;
            float v = 0.f;
            if (y >= 0 && y < H && x >= 0 && x < W) v = (float)__ldg(img + ((f * H + y) * (long long)W + x) * 3 + ch);
            win[e] = v;
        }
        __syncthreads();
        const __nv_bfloat16* gy = dy + ((f * (Ho + 1) + oy) * (long long)(Wo + 1) + ox0) * C0 + c;
#pragma unroll 1
        for (int p = 0; p < kFbSeg; ++p) {
            const float g = __bfloat162float(gy[(long long)p * C0]);
            // 5 x 5 x 3 window of this pooled pixel -> registers
            float v[75];
#pragma unroll
            for (int r = 0; r < 5; ++r)
#pragma unroll
                for (int q = 0; q < 15; ++q) v[r * 15 + q] = win[(r * kFbWinCols + 2 * p) * 3 + q];
            // the 9 convolution outputs of the pooling window; positions outside the image never win
            float best = -INFINITY;
            int arg = 0;
#pragma unroll
            for (int py = 0; py < 3; ++py) {
#pragma unroll
                for (int px = 0; px < 3; ++px) {
                    float a = bc;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) a = fmaf(wr[(ky * 3 + kx) * 3 + ch], v[(py + ky) * 15 + (px + kx) * 3 + ch], a);
                    const int yy = 2 * oy - 1 + py, xx = 2 * (ox0 + p) - 1 + px;
                    const bool inside = yy >= 0 && yy < H && xx >= 0 && xx < W;
                    if (inside && a > best) {
                        best = a;
                        arg = py * 3 + px;
                    }
                }
            }
            const float ge = best > 0.f ? g : 0.f;
            db += ge;
#pragma unroll
            for (int py = 0; py < 3; ++py)
#pragma unroll
                for (int px = 0; px < 3; ++px) {
                    const float coef = (arg == py * 3 + px) ? ge : 0.f;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch)
                                dW[(ky * 3 + kx) * 3 + ch] = fmaf(coef, v[(py + ky) * 15 + (px + kx) * 3 + ch], dW[(ky * 3 + kx) * 3 + ch]);
                }
        }
    }
    float* o = ws + ((long long)blockIdx.x * C0 + c) * 28;
#pragma unroll
    for (int k = 0; k < 27; ++k) o[k] = dW[k];
    o[27] = db;
}

__global__ void firstconv_bwd_finalize_kernel(const float* __restrict__ ws, float* __restrict__ dW, float* __restrict__ db, int S, int C0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;  // over C0 * 28
    if (i >= C0 * 28) return;
    double a = 0.0;
    for (int s = 0; s < S; ++s) a += (double)__ldg(ws + (long long)s * C0 * 28 + i);
    const int c = i / 28, k = i % 28;
    if (k < 27) dW[c * 27 + k] = (float)a;
    else db[c] = (float)a;
}

static inline int firstconv_bwd_blocks(long long F, int H, int W) {
    const long long items = F * (H / 2) * ((W / 2) / kFbSeg);
    long long s = 2LL * num_sms();
    if (s > items) s = items;
    if (s < 1) s = 1;
    return (int)s;
}

}  // namespace vpt

extern "C" int vpt_firstconv_bwd_parts(int64_t F, int32_t H, int32_t W) { return vpt::firstconv_bwd_blocks(F, H, W); }

extern "C" int vpt_firstconv_bwd(const uint8_t* img, const float* w, const float* bias, const void* dy, float* dW, float* db, float* workspace, int64_t F,
                                 int32_t H, int32_t W, int32_t C0, void* stream) {
    using namespace vpt;
    VPT_CHECK(img && w && bias && dy && dW && db && workspace && F > 0, "vpt_firstconv_bwd: null argument");
    VPT_CHECK(H % 2 == 0 && W % (2 * kFbSeg) == 0 && C0 % 32 == 0 && C0 >= 32 && C0 <= 256,
              "vpt_firstconv_bwd: need even H, W %% 16 == 0 and C0 in {32..256} a multiple of 32 (H=%d W=%d C0=%d)", H, W, C0);
    const int S = firstconv_bwd_blocks(F, H, W);
    firstconv_bwd_kernel<<<S, C0, 0, (cudaStream_t)stream>>>(img, w, bias, reinterpret_cast<const __nv_bfloat16*>(dy), workspace, (int)F, H, W, C0);
    VPT_LAUNCH_CHECK();
    firstconv_bwd_finalize_kernel<<<(C0 * 28 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(workspace, dW, db, S, C0);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
