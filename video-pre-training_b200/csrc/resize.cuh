// Frame ingest (SURVEY f-2): bilinear u8 resize that is BIT-EXACT with cv2.resize(..., interpolation=cv2.INTER_LINEAR) on uint8
// images (agent.py:100-103, inverse_dynamics_model.py:54-59, data_loader.py:113-120): OpenCV's 8-bit path is integer
// arithmetic -- 11-bit fixed-point weights, horizontal pass in int32, vertical pass
//     dst = ( ((b0 * (row0 >> 4)) >> 16) + ((b1 * (row1 >> 4)) >> 16) + 2 ) >> 2 .
// The per-column / per-row source indices and weights are computed on the host exactly like OpenCV does (double arithmetic,
// cvFloor / cvRound) and passed in as small tables, so no floating-point rounding difference can occur on the device.
#pragma once
#include "common.cuh"

namespace vpt {

__global__ void __launch_bounds__(256) resize_bilinear_u8_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                                   const int* __restrict__ xidx, const short* __restrict__ xw,
                                                                   const int* __restrict__ yidx, const short* __restrict__ yw, int Hs, int Ws,
                                                                   int Hd, int Wd, int C) {
    const long long f = blockIdx.y;
    const int n = Hd * Wd * C;
    const uint8_t* s = src + f * (long long)Hs * Ws * C;
    uint8_t* d = dst + f * (long long)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int c = i % C, x = (i / C) % Wd, y = i / (C * Wd);
        const int x0 = xidx[x], x1 = min(x0 + 1, Ws - 1), y0 = yidx[y], y1 = min(y0 + 1, Hs - 1);
        const int a0 = xw[2 * x], a1 = xw[2 * x + 1], b0 = yw[2 * y], b1 = yw[2 * y + 1];
        const int r0 = (int)s[((long long)y0 * Ws + x0) * C + c] * a0 + (int)s[((long long)y0 * Ws + x1) * C + c] * a1;
        const int r1 = (int)s[((long long)y1 * Ws + x0) * C + c] * a0 + (int)s[((long long)y1 * Ws + x1) * C + c] * a1;
        int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        d[i] = (uint8_t)v;
    }
}

}  // namespace vpt

extern "C" int vpt_resize_bilinear_u8(const uint8_t* src, uint8_t* dst, const int32_t* xidx, const int16_t* xw, const int32_t* yidx,
                                      const int16_t* yw, int32_t F, int32_t Hs, int32_t Ws, int32_t Hd, int32_t Wd, int32_t C, void* stream) {
    using namespace vpt;
    VPT_CHECK(src && dst && xidx && xw && yidx && yw && F > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0, "vpt_resize_bilinear_u8: bad arguments");
    VPT_CHECK(F <= 65535, "vpt_resize_bilinear_u8: at most 65535 frames per call");
    int bx = (Hd * Wd * C + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid(bx, F);
    resize_bilinear_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, xidx, xw, yidx, yw, Hs, Ws, Hd, Wd, C);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
