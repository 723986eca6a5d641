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
                                                                   int Hd, int Wd, int C, int swap_rb) {
    const long long f = blockIdx.y;
    const int n = Hd * Wd * C;
    const uint8_t* s = src + f * (long long)Hs * Ws * C;
    uint8_t* d = dst + f * (long long)n;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const int co = i % C, x = (i / C) % Wd, y = i / (C * Wd);
        const int c = (swap_rb && co < 3) ? 2 - co : co;  // BGR -> RGB (cv2.cvtColor COLOR_BGR2RGB) commutes with the per-channel resize
        const int x0 = xidx[x], x1 = min(x0 + 1, Ws - 1), y0 = yidx[y], y1 = min(y0 + 1, Hs - 1);
        const int a0 = xw[2 * x], a1 = xw[2 * x + 1], b0 = yw[2 * y], b1 = yw[2 * y + 1];
        const int r0 = (int)s[((long long)y0 * Ws + x0) * C + c] * a0 + (int)s[((long long)y0 * Ws + x1) * C + c] * a1;
        const int r1 = (int)s[((long long)y1 * Ws + x0) * C + c] * a0 + (int)s[((long long)y1 * Ws + x1) * C + c] * a1;
        int v = (((b0 * (r0 >> 4)) >> 16) + ((b1 * (r1 >> 4)) >> 16) + 2) >> 2;
        v = v < 0 ? 0 : (v > 255 ? 255 : v);
        d[i] = (uint8_t)v;
    }
}

// Cursor overlay of the BC data loader (data_loader.py:34-45,113-118): for every frame with cursor_xy[f] = (x, y), x >= 0,
//   frame[y:y+ch, x:x+cw, :] = uint8( frame * (1 - alpha) + cursor * alpha )      in float64, truncated like numpy's astype(uint8)
// with the overlay clipped at the right / bottom frame border.  One thread per (frame, cursor pixel, channel); in place.
__global__ void composite_cursor_kernel(uint8_t* __restrict__ frames, const uint8_t* __restrict__ cursor, const double* __restrict__ alpha,
                                        const int* __restrict__ xy, int H, int W, int ch_full, int cw_full) {
    const long long f = blockIdx.y;
    const int x0 = xy[2 * f], y0 = xy[2 * f + 1];
    if (x0 < 0 || y0 < 0) return;  // no cursor on this frame (GUI closed)
    const int chh = max(0, min(H - y0, ch_full)), cww = max(0, min(W - x0, cw_full));
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ch_full * cw_full * 3) return;
    const int c = i % 3, cx = (i / 3) % cw_full, cy = i / (3 * cw_full);
    if (cy >= chh || cx >= cww) return;
    uint8_t* px = frames + ((f * H + y0 + cy) * (long long)W + x0 + cx) * 3 + c;
    const double a = alpha[cy * cw_full + cx];
    const double v = (double)*px * (1.0 - a) + (double)cursor[(cy * cw_full + cx) * 3 + c] * a;
    *px = (uint8_t)(long long)v;
}

}  // namespace vpt

extern "C" int vpt_composite_cursor_u8(uint8_t* frames, const uint8_t* cursor, const double* alpha, const int32_t* xy, int32_t F, int32_t H,
                                       int32_t W, int32_t ch, int32_t cw, void* stream) {
    using namespace vpt;
    VPT_CHECK(frames && cursor && alpha && xy && F > 0 && F <= 65535 && H > 0 && W > 0 && ch > 0 && cw > 0, "vpt_composite_cursor_u8: bad arguments");
    dim3 grid((ch * cw * 3 + 255) / 256, F);
    composite_cursor_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, cursor, alpha, xy, H, W, ch, cw);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_resize_bilinear_u8(const uint8_t* src, uint8_t* dst, const int32_t* xidx, const int16_t* xw, const int32_t* yidx,
                                      const int16_t* yw, int32_t F, int32_t Hs, int32_t Ws, int32_t Hd, int32_t Wd, int32_t C, int32_t swap_rb, void* stream) {
    using namespace vpt;
    VPT_CHECK(src && dst && xidx && xw && yidx && yw && F > 0 && Hs > 0 && Ws > 0 && Hd > 0 && Wd > 0 && C > 0, "vpt_resize_bilinear_u8: bad arguments");
    VPT_CHECK(F <= 65535, "vpt_resize_bilinear_u8: at most 65535 frames per call");
    int bx = (Hd * Wd * C + 255) / 256;
    if (bx > 64) bx = 64;
    dim3 grid(bx, F);
    resize_bilinear_u8_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, dst, xidx, xw, yidx, yw, Hs, Ws, Hd, Wd, C, swap_rb);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
