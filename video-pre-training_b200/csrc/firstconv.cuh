// Stack-0 first convolution, fully fused: u8 NHWC frame -> (x/255) conv3x3(3->C0)+bias -> ReLU -> max_pool(3,2,1)
// -> bf16 NHWC + per-tile statistics partials.
//
// K = 27 is too small for tcgen05 to matter (the layer is 0.75 % of the FLOPs and bounded by its epilogue / output
// write), so the contraction runs on warp-level mma.sync m16n8k16: A = the im2col view of the u8 patch (u8 values
// are exact in bf16; fragments are gathered straight from the patch in shared memory, no im2col copy), B = the fp32
// weights split as bf16 hi + bf16 lo (K = 27 + 27 -> 64), fp32 accumulate: products are exact and the result is
// fp32-accurate (the first conv is the largest single contributor to bf16 error otherwise).
//
// One CTA = one 8x8 tile of POOLED outputs of one frame = 17x17 conv outputs (19 m16 tiles) from a 19x19x3 patch.
// ~99 KB of shared memory at C0 = 128 -> two CTAs per SM overlap one CTA's pooling with the other's MMAs.
#pragma once
#include "attention.cuh"  // ldsm_x4 / mma_bf16_16816
#include "common.cuh"
#include "firstconv_tc.cuh"

namespace vpt {

constexpr int kFcTile = 8;               // pooled outputs per tile edge
constexpr int kFcConv = 2 * kFcTile + 1; // 17 conv rows/cols
constexpr int kFcIn = kFcConv + 2;       // 19 input rows/cols
constexpr int kFcThreads = 256;
constexpr int kFcPos = kFcConv * kFcConv;        // 289 conv positions
constexpr int kFcMTiles = (kFcPos + 15) / 16;    // 19
constexpr int kFcBPitch = 72;                    // bf16 elements per weight row in smem (144 B: conflict-free ldmatrix)
constexpr int kFcPatchElems = kFcIn * kFcIn * 3; // 1083
constexpr int kFcPatchBytes = 2192;              // bf16 patch + one zero element, 16-byte multiple

__global__ void __launch_bounds__(kFcThreads, 2) firstconv_pool_kernel(const uint8_t* __restrict__ img, const float* __restrict__ w,
                                                                       const float* __restrict__ bias, __nv_bfloat16* __restrict__ out,
                                                                       float2* __restrict__ stat_part, int H, int W, int C0, long long total_tiles, int zp) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t fc_smem[];
    __nv_bfloat16* patch = reinterpret_cast<__nv_bfloat16*>(fc_smem);                                  // [19][19][3]
    __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(fc_smem + kFcPatchBytes);                     // [C0][72]
    const int cpitch = C0 + 8;
    __nv_bfloat16* ctile = Bs + (size_t)C0 * kFcBPitch;                                                // [289][C0+8]
    const int tiles_x = (W / 2) / kFcTile, tiles_y = (H / 2) / kFcTile;
    const int tiles = tiles_x * tiles_y;

    // ---- hi/lo-split weights, staged once per (persistent) CTA
    for (int i = threadIdx.x; i < C0 * 64; i += kFcThreads) {
        const int n = i >> 6, k = i & 63;
        float v = 0.f;
        if (k < 27) {
            v = __ldg(w + n * 27 + k);
        } else if (k < 54) {
            const float x = __ldg(w + n * 27 + k - 27);
            v = x - __bfloat162float(__float2bfloat16_rn(x));
        } else if (k == 54) {  // bias rides along as two extra K columns against a constant-one A column
            v = __ldg(bias + n);
        } else if (k == 55) {
            const float x = __ldg(bias + n);
            v = x - __bfloat162float(__float2bfloat16_rn(x));
        }
        Bs[n * kFcBPitch + k] = __float2bfloat16_rn(v);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    // per-thread patch offsets of the k indices this lane feeds: k -> (ky, kx*3+c) -> ky*57 + q ; k in [27,54) repeats
    int koff[4][4];
    bool kval[4][4];
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = 16 * s + 2 * tg + (e & 1) + ((e >> 1) << 3);
            const int kk = k < 27 ? k : k - 27;
            kval[s][e] = k < 54;
            koff[s][e] = kval[s][e] ? (kk / 9) * (kFcIn * 3) + kk % 9 : 0;
        }
    // (a) patch staging without div/mod in the tile loop: each thread owns fixed patch elements; the NEXT tile's bytes are
    // prefetched into registers while the current tile is computed
    constexpr int kPE = (kFcPatchElems + kFcThreads - 1) / kFcThreads;  // 5
    int pe_row[kPE], pe_col[kPE];  // input row / byte column (x*3+c) inside the patch
#pragma unroll
    for (int j = 0; j < kPE; ++j) {
        const int i = threadIdx.x + j * kFcThreads;
        pe_row[j] = i / (kFcIn * 3);
        pe_col[j] = i % (kFcIn * 3);
    }
    uint8_t pre[kPE];
    auto prefetch = [&](long long tid) {
        const long long f = tid / tiles;
        const int tile = (int)(tid % tiles);
        const int Yin0 = 2 * (tile / tiles_x) * kFcTile - 2, Xin0 = 2 * (tile % tiles_x) * kFcTile - 2;
        const uint8_t* fimg = img + f * (long long)H * W * 3;
#pragma unroll
        for (int j = 0; j < kPE; ++j) {
            const int Y = Yin0 + pe_row[j], xb = Xin0 * 3 + pe_col[j];  // xb = X*3 + c
            const bool ok = (threadIdx.x + j * kFcThreads < kFcPatchElems) && Y >= 0 && Y < H && xb >= 0 && xb < W * 3;
            pre[j] = ok ? __ldg(fimg + (long long)Y * W * 3 + xb) : (uint8_t)0;
        }
    };
    if ((long long)blockIdx.x < total_tiles) prefetch(blockIdx.x);

  for (long long tid = blockIdx.x; tid < total_tiles; tid += gridDim.x) {
    const long long f = tid / tiles;
    const int tile = (int)(tid % tiles);
    const int PY0 = (tile / tiles_x) * kFcTile, PX0 = (tile % tiles_x) * kFcTile;
    // ---- stage the input patch (u8 -> bf16, exact) from the prefetched registers, then prefetch the next tile
#pragma unroll
    for (int j = 0; j < kPE; ++j) {
        const int i = threadIdx.x + j * kFcThreads;
        if (i < kFcPatchElems) patch[i] = __float2bfloat16_rn((float)pre[j]);
    }
    if (tid + gridDim.x < total_tiles) prefetch(tid + gridDim.x);
    __syncthreads();  // patch (and, first time, weights) visible; previous tile's pooling finished reading ctile

    for (int mt = warp; mt < kFcMTiles; mt += kFcThreads / 32) {
        // A fragments (rows g and g+8 of this m-tile) for the 4 k-steps, gathered from the patch
        int pos[2], pbase[2];
        bool inimg[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int p = min(mt * 16 + g + rr * 8, kFcPos - 1);
            pos[rr] = mt * 16 + g + rr * 8;
            const int cy = p / kFcConv, cx = p % kFcConv;
            pbase[rr] = (cy * kFcIn + cx) * 3;
            const int Y = 2 * PY0 - 1 + cy, X = 2 * PX0 - 1 + cx;
            inimg[rr] = (Y >= 0 && Y < H && X >= 0 && X < W);
        }
        uint32_t af[4][4];
#pragma unroll
        for (int s = 0; s < 4; ++s) {
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {  // hh = 0: k, k+1 ; hh = 1: k+8, k+9
                    const uint16_t lo = kval[s][2 * hh] ? *reinterpret_cast<const uint16_t*>(patch + pbase[rr] + koff[s][2 * hh]) : (uint16_t)0;
                    const uint16_t hi = kval[s][2 * hh + 1] ? *reinterpret_cast<const uint16_t*>(patch + pbase[rr] + koff[s][2 * hh + 1]) : (uint16_t)0;
                    af[s][rr + 2 * hh] = (uint32_t)lo | ((uint32_t)hi << 16);
                }
            }
        }
        if (tg == 3) af[3][0] = af[3][1] = 0x3F803F80u;  // k = 54, 55: constant 1.0 (x the bias columns of B)
        for (int nh = 0; nh < C0 / 64; ++nh) {  // 64 output channels at a time
            float acc[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    uint32_t b0, b1, b2, b3;
                    const int nrow = nh * 64 + np * 16 + (lane & 7) + ((lane >> 4) << 3);
                    const int kcol = s * 16 + (((lane >> 3) & 1) << 3);
                    ldsm_x4(smem_u32(Bs + nrow * kFcBPitch + kcol), b0, b1, b2, b3);
                    mma_bf16_16816(acc[2 * np], af[s][0], af[s][1], af[s][2], af[s][3], b0, b1);
                    mma_bf16_16816(acc[2 * np + 1], af[s][0], af[s][1], af[s][2], af[s][3], b2, b3);
                }
            }
            // raw pre-activation values; ReLU is applied after the max (monotone), positions outside the image hold 0
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                if (pos[rr] >= kFcPos) continue;
                __nv_bfloat16* crow = ctile + (size_t)pos[rr] * cpitch + nh * 64 + 2 * tg;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    const float v0 = inimg[rr] ? acc[nt][2 * rr] : 0.f;
                    const float v1 = inimg[rr] ? acc[nt][2 * rr + 1] : 0.f;
                    *reinterpret_cast<uint32_t*>(crow + nt * 8) = pack_bf16(v0, v1);
                }
            }
        }
    }
    __syncthreads();

    // ---- 3x3 / stride-2 max over the conv tile, 8 channels (16 B) per item
    const int C8 = C0 / 8;
    float s = 0.f, ss = 0.f;
    const int Ho = H / 2, Wo = W / 2, opitch = Wo + zp;  // ZP layout: one extra zero column / row
    __nv_bfloat16* fout = out + f * (long long)(Ho + zp) * opitch * C0;
    for (int i = threadIdx.x; i < kFcTile * kFcTile * C8; i += kFcThreads) {
        const int c8 = i % C8, px = (i / C8) % kFcTile, py = i / (C8 * kFcTile);
        uint4 m = make_uint4(0, 0, 0, 0);  // starting from 0 == applying the ReLU after the max
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int p = (2 * py + dy) * kFcConv + 2 * px + dx;
                const uint4 v = *reinterpret_cast<const uint4*>(ctile + (size_t)p * cpitch + 8 * c8);
                m.x = bf16x2_max(m.x, v.x); m.y = bf16x2_max(m.y, v.y);
                m.z = bf16x2_max(m.z, v.z); m.w = bf16x2_max(m.w, v.w);
            }
        *reinterpret_cast<uint4*>(fout + ((long long)(PY0 + py) * opitch + PX0 + px) * C0 + 8 * c8) = m;
        const uint32_t w4[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float a = bf16_lo(w4[q]), b = bf16_hi(w4[q]);
            s += a + b;
            ss = fmaf(a, a, fmaf(b, b, ss));
        }
    }
    if (zp) {  // edge tiles write the zero column x = Wo and the zero row y = Ho (plus the corner)
        const bool right = (PX0 + kFcTile == Wo), bottom = (PY0 + kFcTile == Ho);
        if (right)
            for (int i = threadIdx.x; i < kFcTile * C8; i += kFcThreads)
                *reinterpret_cast<uint4*>(fout + ((long long)(PY0 + i / C8) * opitch + Wo) * C0 + 8 * (i % C8)) = make_uint4(0, 0, 0, 0);
        if (bottom)
            for (int i = threadIdx.x; i < (kFcTile + (right ? 1 : 0)) * C8; i += kFcThreads)
                *reinterpret_cast<uint4*>(fout + ((long long)Ho * opitch + PX0 + i / C8) * C0 + 8 * (i % C8)) = make_uint4(0, 0, 0, 0);
    }
    if (stat_part) {
        const float2 r = block_sum2(s, ss);  // contains __syncthreads: also orders ctile / patch reuse
        if (threadIdx.x == 0) stat_part[f * tiles + tile] = r;
    } else {
        __syncthreads();
    }
  }
}

}  // namespace vpt

extern "C" int vpt_firstconv_stat_parts(int32_t F, int32_t H, int32_t W, int32_t C0) {
    if (vpt::firstconv_tc_applies(H, W)) return (H / 2 / vpt::kFtStatRows) * 2 * C0;  // per (8 pooled rows, column half, channel)
    return (H / 16) * (W / 16);
}

extern "C" int vpt_set_firstconv_mode(int32_t mode) {
    vpt::g_fc_mode = mode;  // 1: tcgen05 kernel (firstconv_tc.cuh) where it applies; 0: always the mma.sync kernel
    return VPT_OK;
}

extern "C" int vpt_firstconv_pool(const uint8_t* img, const float* w, const float* bias, void* out, float* stat_part, int32_t F,
                                  int32_t H, int32_t W, int32_t C0, int32_t zp, int32_t out_f32, void* stream) {
    using namespace vpt;
    VPT_CHECK(img && w && bias && out && F > 0, "vpt_firstconv_pool: null argument");
    VPT_CHECK(!out_f32 || firstconv_tc_applies(H, W), "vpt_firstconv_pool: fp32 output needs the tcgen05 kernel (W in {32,64,128}, H*W <= 16384)");
    VPT_CHECK(H % 16 == 0 && W % 16 == 0 && H >= 16 && W >= 16, "vpt_firstconv_pool: H, W must be multiples of 16 (H=%d W=%d)", H, W);
    VPT_CHECK(C0 == 64 || C0 == 128 || C0 == 192 || C0 == 256, "vpt_firstconv_pool: C0=%d not in {64,128,192,256}", C0);
    if (firstconv_tc_applies(H, W)) {
        VPT_CHECK(((uintptr_t)img & 15) == 0, "vpt_firstconv_pool: img must be 16-byte aligned");
        FirstconvTcParams p;
        memset(&p, 0, sizeof(p));
        p.img = img; p.w = w; p.bias = bias;
        p.out = reinterpret_cast<__nv_bfloat16*>(out);
        p.stat_part = reinterpret_cast<float2*>(stat_part);
        p.H = H; p.C0 = C0; p.zp = zp ? 1 : 0;
        p.ncb = (C0 + 127) / 128;
        p.nbands = firstconv_tc_bands(F, H, p.ncb);
        p.band_rows = (H / 2) / p.nbands;
        p.items = (long long)F * p.nbands * p.ncb;
        if (out_f32) {
            if (W == 32) return launch_firstconv_tc<32, true>(p, stream);
            if (W == 64) return launch_firstconv_tc<64, true>(p, stream);
            return launch_firstconv_tc<128, true>(p, stream);
        }
        if (W == 32) return launch_firstconv_tc<32, false>(p, stream);
        if (W == 64) return launch_firstconv_tc<64, false>(p, stream);
        return launch_firstconv_tc<128, false>(p, stream);
    }
    const long long blocks = (long long)F * (H / 16) * (W / 16);
    VPT_CHECK(blocks < 2147483647LL, "vpt_firstconv_pool: too many tiles");
    const size_t smem = kFcPatchBytes + (size_t)C0 * kFcBPitch * 2 + (size_t)kFcPos * (C0 + 8) * 2;
    static size_t attr = 0;
    if (smem > attr) {
        VPT_CUDA(cudaFuncSetAttribute(firstconv_pool_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm > 2) per_sm = 2;
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)num_sms() * per_sm;
    if (grid > blocks) grid = blocks;
    launch_k(firstconv_pool_kernel, dim3((unsigned)grid), dim3(kFcThreads), smem, (cudaStream_t)stream, 
        img, w, bias, reinterpret_cast<__nv_bfloat16*>(out), reinterpret_cast<float2*>(stat_part), H, W, C0, blocks, zp ? 1 : 0);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
