// Stack-0 first convolution, fully fused: u8 NHWC frame -> (x/255) conv3x3(3->C0)+bias -> ReLU -> max_pool(3,2,1)
// -> bf16 NHWC + per-tile statistics partials.  K = 27 is far too small for the tensor pipe to matter; this layer is
// bounded by its output write (C0*H*W/4 bf16 per frame) and by fp32 FMA issue.  fp32 FMAs keep the u8 input and the
// fp32 weights exact (the first conv is the largest single contributor to bf16 error).
//
// One CTA = one 8x8 tile of POOLED outputs of one frame = 17x17 conv outputs = a 19x19x3 input patch.
#pragma once
#include "common.cuh"

namespace vpt {

constexpr int kFcTile = 8;               // pooled outputs per tile edge
constexpr int kFcConv = 2 * kFcTile + 1; // 17 conv rows/cols
constexpr int kFcIn = kFcConv + 2;       // 19 input rows/cols
constexpr int kFcThreads = 256;

template <int CPT>  // channels per lane per pass; C0 = 32 * CPT * passes
__global__ void __launch_bounds__(kFcThreads) firstconv_pool_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w,
                                                                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                    float2* __restrict__ stat_part, int H, int W, int C0) {
    extern __shared__ uint8_t fc_smem[];
    float* patch = reinterpret_cast<float*>(fc_smem);                                   // [19][19][3]
    __nv_bfloat16* ctile = reinterpret_cast<__nv_bfloat16*>(fc_smem + 4352);            // [289][C0]
    const int tiles_x = (W / 2) / kFcTile, tiles_y = (H / 2) / kFcTile;
    const int tiles = tiles_x * tiles_y;
    const long long f = blockIdx.x / tiles;
    const int tile = blockIdx.x % tiles;
    const int PY0 = (tile / tiles_x) * kFcTile, PX0 = (tile % tiles_x) * kFcTile;
    const int Yin0 = 2 * PY0 - 2, Xin0 = 2 * PX0 - 2;
    const uint8_t* fimg = img + f * (long long)H * W * 3;

    for (int i = threadIdx.x; i < kFcIn * kFcIn * 3; i += kFcThreads) {
        const int c = i % 3, ix = (i / 3) % kFcIn, iy = i / (3 * kFcIn);
        const int Y = Yin0 + iy, X = Xin0 + ix;
        float v = 0.f;
        if (Y >= 0 && Y < H && X >= 0 && X < W) v = (float)__ldg(fimg + ((long long)Y * W + X) * 3 + c);
        patch[i] = v;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int passes = C0 / (32 * CPT);
    for (int ps = 0; ps < passes; ++ps) {
        const int c0 = (ps * 32 + lane) * CPT;
        float wr[CPT][27];
        float br[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) {
            br[j] = __ldg(bias + c0 + j);
#pragma unroll
            for (int k = 0; k < 27; ++k) wr[j][k] = __ldg(w + (size_t)(c0 + j) * 27 + k);
        }
        for (int pos = warp; pos < kFcConv * kFcConv; pos += kFcThreads / 32) {
            const int cy = pos / kFcConv, cx = pos % kFcConv;
            const int Y = 2 * PY0 - 1 + cy, X = 2 * PX0 - 1 + cx;
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = br[j];
            if (Y >= 0 && Y < H && X >= 0 && X < W) {
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float* pr = patch + ((cy + ky) * kFcIn + cx) * 3;
#pragma unroll
                    for (int q = 0; q < 9; ++q) {
                        const float xv = pr[q];
#pragma unroll
                        for (int j = 0; j < CPT; ++j) acc[j] = fmaf(xv, wr[j][ky * 9 + q], acc[j]);
                    }
                }
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[j] = fmaxf(acc[j], 0.f);
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j) acc[j] = 0.f;  // outside the image: neutral for the max (values are >= 0)
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j) ctile[(size_t)pos * C0 + c0 + j] = __float2bfloat16_rn(acc[j]);
        }
    }
    __syncthreads();

    // 3x3 / stride-2 max over the conv tile, two channels per item
    const int C2 = C0 / 2;
    float s = 0.f, ss = 0.f;
    __nv_bfloat16* fout = out + f * (long long)(H / 2) * (W / 2) * C0;
    for (int i = threadIdx.x; i < kFcTile * kFcTile * C2; i += kFcThreads) {
        const int c2 = i % C2, px = (i / C2) % kFcTile, py = i / (C2 * kFcTile);
        __nv_bfloat162 m = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int pos = (2 * py + dy) * kFcConv + 2 * px + dx;
                m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(ctile + (size_t)pos * C0 + 2 * c2));
            }
        *reinterpret_cast<__nv_bfloat162*>(fout + ((long long)(PY0 + py) * (W / 2) + PX0 + px) * C0 + 2 * c2) = m;
        const float a = __low2float(m), b = __high2float(m);
        s += a + b;
        ss = fmaf(a, a, fmaf(b, b, ss));
    }
    if (stat_part) {
        const float2 r = block_sum2(s, ss);
        if (threadIdx.x == 0) stat_part[f * tiles + tile] = r;
    }
}

}  // namespace vpt

extern "C" int vpt_firstconv_stat_parts(int32_t H, int32_t W) { return (H / 16) * (W / 16); }

extern "C" int vpt_firstconv_pool(const uint8_t* img, const float* w, const float* bias, void* out, float* stat_part, int32_t F,
                                  int32_t H, int32_t W, int32_t C0, void* stream) {
    using namespace vpt;
    VPT_CHECK(img && w && bias && out && F > 0, "vpt_firstconv_pool: null argument");
    VPT_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "vpt_firstconv_pool: H, W must be multiples of 16 (H=%d W=%d)", H, W);
    VPT_CHECK(C0 == 64 || C0 == 128 || C0 == 192 || C0 == 256, "vpt_firstconv_pool: C0=%d not in {64,128,192,256}", C0);
    const long long blocks = (long long)F * (H / 16) * (W / 16);
    VPT_CHECK(blocks < 2147483647LL, "vpt_firstconv_pool: too many tiles");
    const size_t smem = 4352 + (size_t)kFcConv * kFcConv * C0 * 2;
    cudaStream_t s = (cudaStream_t)stream;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    float2* sp = reinterpret_cast<float2*>(stat_part);
#define VPT_FC(CPT)                                                                                                     \
    do {                                                                                                                \
        VPT_CUDA(cudaFuncSetAttribute(firstconv_pool_kernel<CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        firstconv_pool_kernel<CPT><<<(unsigned)blocks, kFcThreads, smem, s>>>(img, w, bias, o, sp, H, W, C0);            \
    } while (0)
    if (C0 == 64) VPT_FC(2);
    else if (C0 == 128) VPT_FC(4);
    else if (C0 == 192) VPT_FC(3);
    else VPT_FC(4);
#undef VPT_FC
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
