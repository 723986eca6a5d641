// Stack-0 first convolution on tcgen05, fully fused:
//   u8 NHWC frame -> (x/255) conv3x3(3 -> C0) + bias -> ReLU -> max_pool(3, 2, 1) -> bf16 (ZP layout) + per-channel statistics.
// (lib/policy.py:39-45, lib/util.py:79-81 with bias, lib/impala_cnn.py:115-117)
//
// Why a second version (firstconv.cuh is the mma.sync one): ncu showed the mma.sync kernel at 7 % of its HBM roofline, issue
// bound (19 k warp instructions per 8x8 pooled tile: fragment gathers, a bf16 conv tile in shared memory, a 9-load pooling
// pass).  Here the contraction is ONE tcgen05.mma group per pair of conv rows with the operand roles swapped:
//
//   D[channel][position] = Wsplit[channel][k] * Patch[position][k]^T         M = 128 channels, N = 2 W positions, K = 64
//
//   * K = 27 taps x {bf16 hi, bf16 lo} of the fp32 weights (+ the bias against a constant-one column): u8 pixels are exact in
//     bf16, products are exact, accumulation is fp32 in TMEM -> fp32-accurate;
//   * the "im2col" operand (one 128-byte K row per position, SWIZZLE_128B) is BUILT in shared memory by four producer warps
//     from the raw u8 frame (bulk-copied once per frame): 9 aligned word loads, byte permutes and the 2^23 magic-number
//     int->float trick, 8 conflict-free 16-byte stores per position (~100 instructions);
//   * TMEM lane = output channel, TMEM column = position, so an epilogue THREAD owns one channel and sees the two conv rows
//     of the tile as registers: the 3x3 / stride-2 max is 3-input FMNMX3 in registers (horizontal, then vertical against the
//     previous tile's odd row carried in registers) -- no shared-memory conv tile, no halo recompute, no pooling pass;
//   * one CTA walks the conv rows of (a band of) a frame in order, double-buffered TMEM (2 x 256 columns), triple-buffered
//     operand tiles, double-buffered frames; the weights stay resident in shared memory.
//
//   warps 0..7: epilogue (quarter = warp % 4, column half = warp / 4)   warps 8..11: operand builders
//   warp 12: MMA issuer (+ TMEM alloc)   (the first builder thread also issues the frame bulk copies, one item ahead)
#pragma once
#include <type_traits>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace vpt {

constexpr int kFtEpiWarps = 8;
constexpr int kFtProdWarps = 4;
constexpr int kFtThreads = 32 * (kFtEpiWarps + kFtProdWarps + 1);  // 416
constexpr int kFtMaxFrameBytes = 128 * 128 * 3;
constexpr int kFtMaxBStages = 3;

struct FirstconvTcParams {
    const uint8_t* img;
    const float* w;      // [C0][27] (ky, kx, c), already / 255
    const float* bias;   // [C0]
    __nv_bfloat16* out;
    float2* stat_part;   // [F][P] with P = (H/2/8) * 2 * C0: index ((pooled row / 8) * 2 + half) * C0 + channel
    int H, C0, zp;
    int ncb;             // 128-channel blocks
    int nbands, band_rows;  // pooled rows per band
    int b_stages;
    int frame_stride;    // bytes between the two frame buffers
    long long items;     // F * nbands * ncb
    // backward variant (firstconv_tc_kernel<W, false, true>): gradient of the fused conv + ReLU + max-pool wrt weights and bias
    const __nv_bfloat16* dy;  // bf16 ZP [F][H/2+1][W/2+1][C0]
    float* bwd_ws;            // [gridDim.x * 2][C0][28] partial (dW[27], db) per (CTA, column half, channel)
    long long fb_count;       // F * nbands (backward: items are ordered channel block first: item = cb * fb_count + fb)
};

template <int N>
__device__ __forceinline__ void tmem_ld_cols(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void tmem_ld_cols<32>(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    tmem_ld_32x32(taddr, r);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<16>(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
template <>
__device__ __forceinline__ void tmem_ld_cols<8>(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tmem_ld_col1(uint32_t taddr) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr) : "memory");
    return __uint_as_float(r);
}

__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// item -> (frame, band, channel block)
struct FtItem {
    long long f;
    int band, cb, t0, t1;  // tiles t0..t1 (t = pooled row; tile t holds conv rows 2t, 2t+1); t0 = first pooled row - 1 primes the carry
};
__device__ __forceinline__ FtItem ft_item(const FirstconvTcParams& p, long long item) {
    FtItem it;
    long long fb;
    if (p.dy != nullptr) {  // backward: channel block outermost, so that a CTA's accumulators change block at most once
        it.cb = (int)(item / p.fb_count);
        fb = item - (long long)it.cb * p.fb_count;
    } else {
        it.cb = (int)(item % p.ncb);
        fb = item / p.ncb;
    }
    it.band = (int)(fb % p.nbands);
    it.f = fb / p.nbands;
    it.t0 = it.band * p.band_rows - (it.band > 0 ? 1 : 0);
    it.t1 = (it.band + 1) * p.band_rows - 1;
    return it;
}


constexpr int kFtStatRows = 8;  // pooled rows per statistics partial (fixed, so that the partial sums do not depend on the band split)

// One epilogue chunk: CW conv columns of the two conv rows of a tile (a = row 2t, b = row 2t+1) -> CW/2 pooled outputs of pooled row t.
//   la / lb: conv column to the left of the chunk (in: previous chunk's last column; out: this chunk's last column)
//   carry:   horizontally pooled row 2t-1 (in) / 2t+1 (out)
// EMIT = false only primes the carry (first tile of a later band).
template <int CW, bool F32OUT, bool EMIT>
__device__ __forceinline__ void ft_pool_chunk(const float (&a)[CW], const float (&b)[CW], float& la, float& lb, float (&carry)[CW / 2], void* optr,
                                              size_t stride, bool valid, float& st_s, float& st_ss) {
    using OutT = typename std::conditional<F32OUT, float, __nv_bfloat16>::type;
    OutT* op = reinterpret_cast<OutT*>(optr);
#pragma unroll
    for (int k = 0; k < CW / 2; ++k) {
        const float pa = (k == 0) ? la : a[2 * k - 1], pb = (k == 0) ? lb : b[2 * k - 1];
        const float m0 = fmaxf(fmaxf(pa, a[2 * k]), a[2 * k + 1]);        // conv row 2t
        const float m1 = fmaxf(fmaxf(pb, b[2 * k]), b[2 * k + 1]);        // conv row 2t + 1
        if (EMIT) {
            const float o = fmaxf(fmaxf(fmaxf(carry[k], m0), m1), 0.f);   // rows 2t-1, 2t, 2t+1 ; ReLU after the max
            float of = o;
            if (F32OUT) {
                if (valid) *reinterpret_cast<float*>(op) = o;
            } else {
                const __nv_bfloat16 ob = __float2bfloat16_rn(o);
                if (valid) *reinterpret_cast<__nv_bfloat16*>(op) = ob;
                of = __bfloat162float(ob);
            }
            st_s += of;
            st_ss = fmaf(of, of, st_ss);
            op += stride;
        }
        carry[k] = m1;
    }
    la = a[CW - 1];
    lb = b[CW - 1];
}

template <int W, bool F32OUT, bool BWD>
__global__ void __launch_bounds__(kFtThreads, 1) firstconv_tc_kernel(const FirstconvTcParams p) {
    pdl_sync();
    constexpr int NPOS = 2 * W;           // UMMA N: two conv rows
    constexpr int CW = 16;                // columns per epilogue chunk
    constexpr int NCH = W / 32;           // chunks per epilogue warp and conv row (the warp owns W/2 columns)
    constexpr int PITCH = W * 3;          // bytes per frame row
    constexpr uint32_t kBStage = NPOS * 128;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* s_w = smem;                                            // ncb x [128 channels][64 k] bf16, SWIZZLE_128B
    uint8_t* s_b = s_w + (size_t)p.ncb * 16384;                     // b_stages x [NPOS positions][64 k]
    uint8_t* s_frames = s_b + (size_t)p.b_stages * kBStage;         // 2 x (16 B slack + frame + slack)
    uint8_t* s_zero = s_frames + 2 * (size_t)p.frame_stride;        // 16 B slack + one zero row + slack
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_zero + ((PITCH + 48 + 15) / 16) * 16);
    uint64_t* b_full = bars;
    uint64_t* b_empty = bars + kFtMaxBStages;
    uint64_t* fr_full = bars + 2 * kFtMaxBStages;
    uint64_t* fr_empty = fr_full + 2;
    uint64_t* tmem_full_bar = fr_empty + 2;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int H = p.H;

    // ---- weights: [channel][k] bf16 rows of 128 B: k 0..26 hi(w), 27 hi(bias), 28..54 lo(w), 55 lo(bias), 56..63 zero
    for (int i = threadIdx.x; i < p.ncb * 128 * 8; i += kFtThreads) {
        const int row = i >> 3, chunk = i & 7;
        const int ch = row;  // global channel (row of block cb = row / 128)
        uint32_t wd[4];
#pragma unroll
        for (int h2 = 0; h2 < 4; ++h2) {
            float v[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = chunk * 8 + h2 * 2 + e;
                float x = 0.f;
                bool lo = false;
                if (ch < p.C0) {
                    if (k < 27) x = __ldg(p.w + ch * 27 + k);
                    else if (k == 27) x = __ldg(p.bias + ch);
                    else if (k < 55) { x = __ldg(p.w + ch * 27 + k - 28); lo = true; }
                    else if (k == 55) { x = __ldg(p.bias + ch); lo = true; }
                }
                v[e] = lo ? x - round_bf16(x) : x;
            }
            wd[h2] = pack_bf16(v[0], v[1]);
        }
        const int r = row & 127;
        uint8_t* dst = s_w + (size_t)(row >> 7) * 16384 + r * 128 + ((chunk ^ (r & 7)) << 4);
        *reinterpret_cast<uint4*>(dst) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
    }
    for (int i = threadIdx.x; i < (PITCH + 48) / 4; i += kFtThreads) reinterpret_cast<uint32_t*>(s_zero)[i] = 0u;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < p.b_stages; ++i) {
            mbar_init(&b_full[i], kFtProdWarps);
            mbar_init(&b_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&fr_full[i], 1);
            mbar_init(&fr_empty[i], kFtProdWarps + (BWD ? kFtEpiWarps : 0));  // backward: the epilogue warps gather patches from the frame too
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kFtEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 12) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    fence_proxy_async();  // the weight tile is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 12) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16(128, NPOS);
            int stage = 0, local = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (long long item = blockIdx.x; item < p.items && ok; item += gridDim.x) {
                const FtItem it = ft_item(p, item);
                const uint32_t w_addr = smem_u32(s_w + (size_t)it.cb * 16384);
                for (int t = it.t0; t <= it.t1 && ok; ++t, ++local) {
                    const int as = local & 1;
                    if (!(ok = mbar_wait(&tmem_empty_bar[as], ((uint32_t)(local >> 1) & 1u) ^ 1u, 0x920u))) break;
                    if (!(ok = mbar_wait(&b_full[stage], phase, 0x921u))) break;
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(s_b + (size_t)stage * kBStage);
                    const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStageCols);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (uint32_t)(k != 0));
                    umma_commit(&b_empty[stage]);
                    umma_commit(&tmem_full_bar[as]);
                    advance(stage, phase, p.b_stages);
                }
            }
        }
    } else if (warp >= kFtEpiWarps) {
        // ================= operand builders (warps 8..11): one 128-byte K row per conv position =================
        const int tp = threadIdx.x - 32 * kFtEpiWarps;  // 0..127
        int stage = 0, li = 0;
        uint32_t phase = 0;
        bool ok = true;
        // frame loader (builder thread 0): the rows a band needs, one bulk copy per item, requested one item ahead
        auto request_frame = [&](long long item, int lj) {
            const FtItem it = ft_item(p, item);
            const int buf = lj & 1;
            if (!mbar_wait(&fr_empty[buf], (uint32_t)((lj >> 1) & 1) ^ 1u, 0x910u)) return false;
            const int r0 = max(0, 2 * it.t0 - 1), r1 = min(H, 2 * it.t1 + 3);  // input rows [r0, r1)
            const uint32_t bytes = (uint32_t)(r1 - r0) * PITCH;
            mbar_expect_tx(&fr_full[buf], bytes);
            bulk_copy_g2s(s_frames + (size_t)buf * p.frame_stride + 16 + (size_t)r0 * PITCH, p.img + (size_t)it.f * H * PITCH + (size_t)r0 * PITCH, bytes,
                          &fr_full[buf]);
            return true;
        };
        if (tp == 0 && (long long)blockIdx.x < p.items) ok = request_frame(blockIdx.x, 0);
        for (long long item = blockIdx.x; item < p.items && ok; item += gridDim.x, ++li) {
            const FtItem it = ft_item(p, item);
            const int buf = li & 1;
            if (tp == 0 && item + gridDim.x < p.items) request_frame(item + gridDim.x, li + 1);
            __syncwarp();
            if (!(ok = mbar_wait(&fr_full[buf], (uint32_t)(li >> 1) & 1u, 0x930u))) break;
            const uint8_t* fr = s_frames + (size_t)buf * p.frame_stride + 16;
            for (int t = it.t0; t <= it.t1 && ok; ++t) {
                if (!(ok = mbar_wait(&b_empty[stage], phase ^ 1u, 0x931u))) break;
                uint8_t* tile = s_b + (size_t)stage * kBStage;
#pragma unroll
                for (int n = tp; n < NPOS; n += 32 * kFtProdWarps) {
                    const int r = n / W, x = n % W;
                    const int b0 = 3 * x - 3;                 // first byte of the 9-byte window (pixels x-1, x, x+1) in a frame row
                    const int a0 = b0 & ~3;                   // aligned word holding it (x = 0: -4, the slack before the row)
                    const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(b0 & 3);
                    const uint32_t m0 = (x == 0) ? 0xFF000000u : 0xFFFFFFFFu;      // pixel -1 is zero padding
                    const uint32_t m1 = (x == W - 1) ? 0x0000FFFFu : 0xFFFFFFFFu;  // pixel W is zero padding
                    const uint32_t m2 = (x == W - 1) ? 0u : 0xFFFFFFFFu;
                    float fv[27];
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const int yy = 2 * t + r + ky - 1;
                        const uint8_t* src = (yy < 0 || yy >= H) ? (s_zero + 16) : (fr + (size_t)yy * PITCH);
                        const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + a0);
                        const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                        const uint32_t A0 = prmt(w0, w1, sel) & m0, A1 = prmt(w1, w2, sel) & m1, A2 = prmt(w2, w2, sel) & m2;
#pragma unroll
                        for (int j = 0; j < 9; ++j) {
                            const uint32_t srcw = j < 4 ? A0 : (j < 8 ? A1 : A2);
                            // 0x4B0000vv = 2^23 + v as fp32; subtracting 2^23 leaves v exactly
                            fv[ky * 9 + j] = __uint_as_float(prmt(srcw, 0x4B000000u, 0x7440u | (uint32_t)(j & 3))) - 8388608.0f;
                        }
                    }
                    uint32_t wd[14];  // bf16 pairs (exact: upper halves of the fp32 values); k = 27 is the constant 1.0
#pragma unroll
                    for (int i = 0; i < 13; ++i) wd[i] = prmt(__float_as_uint(fv[2 * i]), __float_as_uint(fv[2 * i + 1]), 0x7632u);
                    wd[13] = prmt(__float_as_uint(fv[26]), 0x3F800000u, 0x7632u);
                    uint8_t* row = tile + (size_t)n * 128;
                    const int sw = n & 7;
                    *reinterpret_cast<uint4*>(row + ((0 ^ sw) << 4)) = make_uint4(wd[0], wd[1], wd[2], wd[3]);
                    *reinterpret_cast<uint4*>(row + ((1 ^ sw) << 4)) = make_uint4(wd[4], wd[5], wd[6], wd[7]);
                    *reinterpret_cast<uint4*>(row + ((2 ^ sw) << 4)) = make_uint4(wd[8], wd[9], wd[10], wd[11]);
                    *reinterpret_cast<uint4*>(row + ((3 ^ sw) << 4)) = make_uint4(wd[12], wd[13], wd[0], wd[1]);  // k 28.. = second copy
                    *reinterpret_cast<uint4*>(row + ((4 ^ sw) << 4)) = make_uint4(wd[2], wd[3], wd[4], wd[5]);
                    *reinterpret_cast<uint4*>(row + ((5 ^ sw) << 4)) = make_uint4(wd[6], wd[7], wd[8], wd[9]);
                    *reinterpret_cast<uint4*>(row + ((6 ^ sw) << 4)) = make_uint4(wd[10], wd[11], wd[12], wd[13]);
                    *reinterpret_cast<uint4*>(row + ((7 ^ sw) << 4)) = make_uint4(0u, 0u, 0u, 0u);
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&b_full[stage]);
                advance(stage, phase, p.b_stages);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fr_empty[buf]);  // this warp is done reading the frame
        }
    } else if (BWD) {
        // ================= backward epilogue (warps 0..7): thread = output channel =================
        // The conv map is recomputed on the tensor core exactly like the forward; per pooled output the thread finds the FIRST maximum
        // of its 3x3 window (row-major, positions outside the image excluded: max_pool2d pads with -inf), applies the ReLU mask
        // (only a positive maximum passes a gradient), reads dy and accumulates  dW[k] += g * patch(argmax)[k],  db += g  in registers;
        // the 27 patch values are gathered from the u8 frame in shared memory with the builders' aligned-window / PRMT / 2^23 trick.
        const int quarter = warp & 3, half = warp >> 2;
        const int x0 = half * (W / 2);
        const int Ho = H / 2, Wo = W / 2;
        const int dpitch = Wo + 1;  // dy is ZP
        float dW[27], db = 0.f;
#pragma unroll
        for (int k = 0; k < 27; ++k) dW[k] = 0.f;
        int cur_cb = -1;
        unsigned flushed = 0u;
        auto flush = [&](int cb) {  // partial sums of channel block cb -> this CTA's slot
            const int ch = cb * 128 + quarter * 32 + lane;
            if (ch < p.C0) {
                float* o = p.bwd_ws + (((size_t)blockIdx.x * 2 + half) * p.C0 + ch) * 28;
#pragma unroll
                for (int k = 0; k < 27; ++k) o[k] = dW[k];
                o[27] = db;
            }
#pragma unroll
            for (int k = 0; k < 27; ++k) dW[k] = 0.f;
            db = 0.f;
            flushed |= 1u << cb;
        };
        int local = 0, li = 0;
        bool ok = true;
        for (long long item = blockIdx.x; item < p.items && ok; item += gridDim.x, ++li) {
            const FtItem it = ft_item(p, item);
            if (cur_cb >= 0 && it.cb != cur_cb) flush(cur_cb);
            cur_cb = it.cb;
            const int ch = it.cb * 128 + quarter * 32 + lane;
            const bool valid = ch < p.C0;
            const int buf = li & 1;
            const uint8_t* fr = s_frames + (size_t)buf * p.frame_stride + 16;
            const __nv_bfloat16* fdy = p.dy + (size_t)it.f * (Ho + 1) * dpitch * p.C0 + ch;
            float cval[NCH][CW / 2];   // horizontally reduced conv row 2t-1: value ...
            uint32_t ccol[NCH];        // ... and which of its three columns held it (2 bits per pooled column)
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                ccol[j] = 0u;
#pragma unroll
                for (int k = 0; k < CW / 2; ++k) cval[j][k] = -INFINITY;
            }
            const bool warp_idle = it.cb * 128 + quarter * 32 >= p.C0;  // padding channels only (C0 = 64 / 192): keep the barrier protocol, skip the work
            for (int t = it.t0; t <= it.t1 && ok; ++t, ++local) {
                const int as = local & 1;
                const bool emit = (t >= it.band * p.band_rows) && valid;
                if (!(ok = mbar_wait(&tmem_full_bar[as], (uint32_t)(local >> 1) & 1u, 0x950u))) break;
                tc_fence_after();
                if (warp_idle) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                    continue;
                }
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccStageCols);
                float la = -INFINITY, lb = -INFINITY;  // conv column x0 - 1 (outside the image for the left half)
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    float a[CW], b[CW];
                    tmem_ld_cols<CW>(trow + x0 + j * CW, a);
                    tmem_ld_cols<CW>(trow + W + x0 + j * CW, b);
                    if (j == 0 && half == 1) {
                        la = tmem_ld_col1(trow + x0 - 1);
                        lb = tmem_ld_col1(trow + W + x0 - 1);
                    }
                    tmem_ld_wait();
                    if (j == NCH - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                    }
                    uint32_t ncol = 0u;
#pragma unroll
                    for (int k = 0; k < CW / 2; ++k) {
                        // first maximum of the three columns 2px-1, 2px, 2px+1 in conv rows 2t (v0, i0) and 2t+1 (v1, i1)
                        float v0 = (k == 0) ? la : a[2 * k - 1], v1 = (k == 0) ? lb : b[2 * k - 1];
                        int i0 = 0, i1 = 0;
                        if (a[2 * k] > v0) { v0 = a[2 * k]; i0 = 1; }
                        if (a[2 * k + 1] > v0) { v0 = a[2 * k + 1]; i0 = 2; }
                        if (b[2 * k] > v1) { v1 = b[2 * k]; i1 = 1; }
                        if (b[2 * k + 1] > v1) { v1 = b[2 * k + 1]; i1 = 2; }
                        // window rows 2t-1 (carried), 2t, 2t+1 in that order
                        float best = cval[j][k];
                        int br = 0, bj = (int)((ccol[j] >> (2 * k)) & 3u);
                        if (v0 > best) { best = v0; br = 1; bj = i0; }
                        if (v1 > best) { best = v1; br = 2; bj = i1; }
                        cval[j][k] = v1;
                        ncol |= (uint32_t)i1 << (2 * k);
                        if (emit && best > 0.f) {
                            const int px = (x0 + j * CW) / 2 + k;
                            const float g = __bfloat162float(fdy[((size_t)t * dpitch + px) * p.C0]);
                            const int y = 2 * t - 1 + br, x = 2 * px - 1 + bj;  // conv position that won
                            db += g;
                            const int b0 = 3 * x - 3, a0 = b0 & ~3;
                            const uint32_t sel = 0x3210u + 0x1111u * (uint32_t)(b0 & 3);
                            const uint32_t m0 = (x == 0) ? 0xFF000000u : 0xFFFFFFFFu;
                            const uint32_t m1 = (x == W - 1) ? 0x0000FFFFu : 0xFFFFFFFFu;
                            const uint32_t m2 = (x == W - 1) ? 0u : 0xFFFFFFFFu;
#pragma unroll
                            for (int ky = 0; ky < 3; ++ky) {
                                const int yy = y + ky - 1;
                                const uint8_t* src = (yy < 0 || yy >= H) ? (s_zero + 16) : (fr + (size_t)yy * PITCH);
                                const uint32_t* wp = reinterpret_cast<const uint32_t*>(src + a0);
                                const uint32_t w0 = wp[0], w1 = wp[1], w2 = wp[2];
                                const uint32_t A0 = prmt(w0, w1, sel) & m0, A1 = prmt(w1, w2, sel) & m1, A2 = prmt(w2, w2, sel) & m2;
#pragma unroll
                                for (int q = 0; q < 9; ++q) {
                                    const uint32_t srcw = q < 4 ? A0 : (q < 8 ? A1 : A2);
                                    const float pv = __uint_as_float(prmt(srcw, 0x4B000000u, 0x7440u | (uint32_t)(q & 3))) - 8388608.0f;
                                    dW[ky * 9 + q] = fmaf(g, pv, dW[ky * 9 + q]);
                                }
                            }
                        }
                    }
                    ccol[j] = ncol;
                    la = a[CW - 1];
                    lb = b[CW - 1];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&fr_empty[buf]);  // this warp is done gathering from the frame
        }
        if (cur_cb >= 0) flush(cur_cb);
        for (int cb = 0; cb < p.ncb; ++cb)  // channel blocks this CTA never worked on: their slots must still hold zeros
            if (!(flushed & (1u << cb))) flush(cb);
    } else {
        // ================= epilogue (warps 0..7): thread = output channel; pooling in registers =================
        using OutT = typename std::conditional<F32OUT, float, __nv_bfloat16>::type;
        const int quarter = warp & 3, half = warp >> 2;
        const int x0 = half * (W / 2);          // this warp's conv columns [x0, x0 + W/2)
        const int Ho = H / 2, Wo = W / 2;
        const int opitch = Wo + p.zp;
        const size_t cstride = (size_t)p.C0;    // elements between pooled columns
        OutT* const outp = reinterpret_cast<OutT*>(p.out);
        const int sgroups = Ho / kFtStatRows;   // statistics partials per frame and (half, channel)
        int local = 0;
        bool ok = true;
        for (long long item = blockIdx.x; item < p.items && ok; item += gridDim.x) {
            const FtItem it = ft_item(p, item);
            const int ch = it.cb * 128 + quarter * 32 + lane;
            const bool valid = ch < p.C0;
            OutT* const fout = outp + (size_t)it.f * (Ho + p.zp) * opitch * p.C0 + ch;  // (frame, row 0, col 0, ch)
            float2* const fstat = p.stat_part ? p.stat_part + (size_t)it.f * (sgroups * 2 * p.C0) + (size_t)half * p.C0 + ch : nullptr;
            float carry[NCH][CW / 2];
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int k = 0; k < CW / 2; ++k) carry[j][k] = 0.f;
            float st_s = 0.f, st_ss = 0.f;
            for (int t = it.t0; t <= it.t1 && ok; ++t, ++local) {
                const int as = local & 1;
                const bool emit = t >= it.band * p.band_rows;  // the first tile of a later band only primes the carry
                if (!(ok = mbar_wait(&tmem_full_bar[as], (uint32_t)(local >> 1) & 1u, 0x940u))) break;
                tc_fence_after();
                const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccStageCols);
                float la = 0.f, lb = 0.f;  // conv column x0 - 1 (0 stands for "outside": every result is max'ed with 0 by the ReLU)
                OutT* orow = fout + ((size_t)t * opitch + x0 / 2) * cstride;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    float a[CW], b[CW];
                    tmem_ld_cols<CW>(trow + x0 + j * CW, a);
                    tmem_ld_cols<CW>(trow + W + x0 + j * CW, b);
                    if (j == 0 && half == 1) {
                        la = tmem_ld_col1(trow + x0 - 1);
                        lb = tmem_ld_col1(trow + W + x0 - 1);
                    }
                    tmem_ld_wait();
                    if (j == NCH - 1) {  // the accumulator stage is in registers: hand it back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                    }
                    if (emit) ft_pool_chunk<CW, F32OUT, true>(a, b, la, lb, carry[j], orow + (size_t)j * (CW / 2) * cstride, cstride, valid, st_s, st_ss);
                    else ft_pool_chunk<CW, F32OUT, false>(a, b, la, lb, carry[j], orow, cstride, valid, st_s, st_ss);
                }
                if (emit && valid) {
                    if (p.zp && half == 1) fout[((size_t)t * opitch + Wo) * cstride] = OutT(0.f);  // zero column of the ZP layout
                    if (fstat && (t % kFtStatRows) == kFtStatRows - 1) {                            // one partial per 8 pooled rows
                        fstat[(size_t)(t / kFtStatRows) * 2 * p.C0] = make_float2(st_s, st_ss);
                        st_s = st_ss = 0.f;
                    }
                }
            }
            if (ok && valid && p.zp && it.band == p.nbands - 1) {  // zero row y = Ho (+ the corner)
                OutT* zrow = fout + (size_t)Ho * opitch * cstride;
                for (int px = x0 / 2; px < x0 / 2 + W / 4 + (half == 1 ? 1 : 0); ++px) zrow[(size_t)px * cstride] = OutT(0.f);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 12) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int g_fc_mode = 1;  // 1: tcgen05 kernel where it applies; 0: always the mma.sync kernel (A/B knob)

static int firstconv_tc_bands(long long F, int H, int ncb) {
    const int Ho = H / 2;
    int nb = 1;
    const long long target = 2LL * (num_sms() > 0 ? num_sms() : 148);
    // bands hold whole 8-row statistics groups, so the partial sums (hence the results, bit for bit) do not depend on the split
    while (F * ncb * nb < target && Ho % (nb * 2 * kFtStatRows) == 0) nb *= 2;
    return nb;
}
static bool firstconv_tc_applies(int H, int W) {
    return g_fc_mode == 1 && (W == 32 || W == 64 || W == 128) && (long long)H * W * 3 <= kFtMaxFrameBytes && H % 2 == 0;
}

template <int W, bool F32OUT, bool BWD = false>
static int launch_firstconv_tc(const FirstconvTcParams& p0, void* stream, int fixed_grid = 0) {
    FirstconvTcParams p = p0;
    const int pitch = W * 3;
    p.frame_stride = ((p.H * pitch + 48) + 15) / 16 * 16;
    const size_t fixed = 1024 + (size_t)p.ncb * 16384 + 2 * (size_t)p.frame_stride + ((pitch + 48 + 15) / 16) * 16 + (2 * kFtMaxBStages + 8) * 8 + 64;
    const size_t bstage = (size_t)2 * W * 128;
    int bst = (int)((227 * 1024 - fixed) / bstage);
    if (bst > kFtMaxBStages) bst = kFtMaxBStages;
    VPT_CHECK(bst >= 2, "vpt_firstconv_pool: not enough shared memory (H=%d W=%d)", p.H, W);
    p.b_stages = bst;
    const size_t smem = fixed + bst * bstage;
    static size_t attr = 0;
    if (smem > attr) {
        VPT_CUDA(cudaFuncSetAttribute(firstconv_tc_kernel<W, F32OUT, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    long long grid = num_sms() > 0 ? num_sms() : 148;
    if (grid > p.items) grid = p.items;
    if (fixed_grid > 0) grid = fixed_grid;  // backward: the partial-sum workspace is sized for exactly this many CTAs
    launch_k(firstconv_tc_kernel<W, F32OUT, BWD>, dim3((unsigned)grid), dim3(kFtThreads), smem, (cudaStream_t)stream, p);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

}  // namespace vpt
