// IDM temporal pre-stage (lib/policy.py:394-403): Conv3d(3 -> C, kernel (5,1,1), pad (2,0,0)) + bias + ReLU applied per
// sample over its T frames (zero padded in time at both ends of the chunk, like the reference's per-sample loop), fused
// with the u8 -> /255 preprocessing.  K = 15 MACs per output: the stage is bounded by its bf16 output write
// (C*H*W*2 B per frame), so it is a plain streaming kernel: one thread = one pixel x 8 channels (16-byte store),
// output in the ZP layout + per-frame statistics partials for the GroupNorm of the following conv.
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace vpt {

__global__ void __launch_bounds__(256) conv3d_t5_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w, const float* __restrict__ bias,
                                                          uint4* __restrict__ out, float2* __restrict__ stat_part, int T, int H, int W, int C, int out_f32) {
    extern __shared__ float c3_smem[];  // [15][C] weights (k-major so that 8 consecutive channels are contiguous) + [C] bias
    float* ws = c3_smem;
    float* bs = c3_smem + 15 * C;
    for (int i = threadIdx.x; i < 15 * C; i += blockDim.x) {
        const int k = i / C, c = i % C;
        ws[i] = __ldg(w + c * 15 + k);
    }
    for (int i = threadIdx.x; i < C; i += blockDim.x) bs[i] = __ldg(bias + i);
    __syncthreads();
    const long long f = blockIdx.y;          // frame index b*T + t
    const int t = (int)(f % T);
    const int C8 = C / 8, Wp = W + 1;
    const long long items = (long long)(H + 1) * Wp * C8;
    uint4* fout = out + f * items * (out_f32 ? 2 : 1);  // fp32 output (precision mode): two uint4 per 8 channels
    float s = 0.f, ss = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const int c0 = (int)(i % C8) * 8;
        const int pix = (int)(i / C8);
        const int y = pix / Wp, x = pix - y * Wp;
        if (y >= H || x >= W) {
            if (out_f32) fout[2 * i] = fout[2 * i + 1] = make_uint4(0, 0, 0, 0);
            else fout[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = bs[c0 + j];
#pragma unroll
        for (int dt = 0; dt < 5; ++dt) {
            const int tt = t + dt - 2;
            if (tt < 0 || tt >= T) continue;  // zero padding in time
            const uint8_t* px = img + ((f + dt - 2) * H * W + (long long)y * W + x) * 3;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float v = (float)__ldg(px + c);
                const float* wk = ws + (dt * 3 + c) * C + c0;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[j] = fmaf(v, wk[j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
        if (out_f32) {
            fout[2 * i] = make_uint4(__float_as_uint(acc[0]), __float_as_uint(acc[1]), __float_as_uint(acc[2]), __float_as_uint(acc[3]));
            fout[2 * i + 1] = make_uint4(__float_as_uint(acc[4]), __float_as_uint(acc[5]), __float_as_uint(acc[6]), __float_as_uint(acc[7]));
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s += acc[j];
                ss = fmaf(acc[j], acc[j], ss);
            }
            continue;
        }
        uint4 o;
        o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]); o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
        fout[i] = o;
        const uint32_t w4[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float a = bf16_lo(w4[q]), b = bf16_hi(w4[q]);
            s += a + b;
            ss = fmaf(a, a, fmaf(b, b, ss));
        }
    }
    if (stat_part) {
        const float2 r = block_sum2(s, ss);
        if (threadIdx.x == 0) stat_part[f * gridDim.x + blockIdx.x] = r;
    }
}

}  // namespace vpt

extern "C" int vpt_conv3d_stat_parts(int32_t H, int32_t W, int32_t C) { return vpt_blocks_for((long long)(H + 1) * (W + 1) * (C / 8), 4096, 64); }

extern "C" int vpt_conv3d_t5(const uint8_t* img, const float* w, const float* bias, void* out, float* stat_part, int32_t B, int32_t T,
                             int32_t H, int32_t W, int32_t C, int32_t out_f32, void* stream) {
    using namespace vpt;
    VPT_CHECK(img && w && bias && out && B > 0 && T > 0, "vpt_conv3d_t5: null argument");
    VPT_CHECK(C % 8 == 0 && C <= 512, "vpt_conv3d_t5: C=%d must be a multiple of 8 and <= 512", C);
    const long long F = (long long)B * T;
    const int bpf = vpt_conv3d_stat_parts(H, W, C);
    const size_t smem = (size_t)16 * C * sizeof(float);
    for (long long f0 = 0; f0 < F; f0 += 65535 / T * T) {  // grid.y limit; slabs hold whole sequences
        long long fn = F - f0;
        if (fn > 65535 / T * T) fn = 65535 / T * T;
        dim3 grid(bpf, (unsigned)fn);
        conv3d_t5_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(
            img + f0 * H * W * 3, w, bias, reinterpret_cast<uint4*>(out) + f0 * (long long)(H + 1) * (W + 1) * (C / 8) * (out_f32 ? 2 : 1),
            stat_part ? reinterpret_cast<float2*>(stat_part) + f0 * bpf : nullptr, T, H, W, C, out_f32 ? 1 : 0);
        VPT_LAUNCH_CHECK();
    }
    return VPT_OK;
}
