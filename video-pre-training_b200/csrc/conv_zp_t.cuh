// Operand-swapped variant of conv3x3_zp for Cout == 128: D^T[channel][pixel] = W[channel][k] * X[pixel][k]^T.
//
// Why: single-CTA SS-mode UMMA time is set by the operand rows it fetches from shared memory, ~(M_rows + N_rows)/2 + 12
// cycles per instruction.  With pixels as M (128) and 128 output channels as N that is 140 cycles for
// 128x128x16 MACs (46 % of the tensor pipe); with the 128 channels as M and 256 PIXELS as N it is 204 cycles for twice the
// work (63 %).  The price is a transposed accumulator: TMEM lane = output channel, TMEM column = pixel, so each epilogue
// thread owns one channel and walks over pixels; global accesses stay coalesced because the 32 lanes of a warp are 32
// consecutive channels of one pixel (64 contiguous bytes), and the per-pixel statistics are produced with a butterfly
// transpose-reduce across the warp.
//
//   warp 0: activation-span TMA producer   warp 1: MMA issuer   warp 2: weight-tile TMA producer   warps 3..10: epilogue
#pragma once
#include "conv_zp.cuh"

namespace vpt {

constexpr int kCtPix = 256;  // pixel rows (ZP linear index) per tile = UMMA N
constexpr int kCtPitch = 132;  // fp32 elements per row of the transposed [pixel][channel] tile (528 B: conflict-free)

struct ConvZpTParams {
    long long Q;
    int H, W, Wp, FS;
    int cin, cin_blocks;
    int a_box_rows, a_boxes, a_stage_bytes, b_stages;
    long long num_tiles;
    const float* mr;
    const float* S1;  // [9][128]
    const float* S2;  // [9][128]
    int relu;
    int dbg_skip_epilogue;
    const __nv_bfloat16* residual;
    __nv_bfloat16* out;
    float* stat_part;  // epi_mode 0: [Q] float2 (complete row sums) or null; epi_mode 1: [num_tiles][8 warps][2 frame slots] float2
    const float* Ef;         // [F][9][128] per-frame fold table (two-norm composition, see vpt_conv_zp_args) or null
    const float* res_scale;  // [F][128] or null: the residual enters as res_scale * r + res_shift
    const float* res_shift;
    int epi_mode;      // 1: fragment epilogue (tcgen05.ld.16x256b -> stmatrix.trans -> TMA store, TMA-prefetched residual); 0: round-1 epilogue
    int stage_off;     // epi_mode 1: byte offset of the 64 KB output / residual staging area (after the weight stages)
};

__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
// 32 lanes x 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void stsm_x4_trans(uint32_t addr, const uint32_t (&r)[4]) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]) : "memory");
}

// lane L ends up with the sum over the warp's 32 lanes of x[L]  (31 shuffles)
__device__ __forceinline__ float transpose_reduce32(float (&x)[32], int lane) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int i = 0; i < s; ++i) {
            const float send = up ? x[i] : x[i + s];
            const float keep = up ? x[i + s] : x[i];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, s);
        }
    }
    return x[0];
}

// epilogue v3 helpers: patch / zero ONE compile-time element of the chunk (used under a jump table on a warp-uniform position)
template <int J>
__device__ __forceinline__ void v3_patch(float (&x)[32], const uint32_t (&acc)[32], float ga, float e) {
    if constexpr (J >= 0 && J < 32) x[J] = fmaf(ga, __uint_as_float(acc[J]), e);
}
template <int J>
__device__ __forceinline__ void v3_zero(float (&x)[32]) {
    if constexpr (J >= 0 && J < 32) x[J] = 0.f;
}
#define VPT_V3_CASES(M) M(0) M(1) M(2) M(3) M(4) M(5) M(6) M(7) M(8) M(9) M(10) M(11) M(12) M(13) M(14) M(15) M(16) M(17) M(18) M(19) M(20) \
    M(21) M(22) M(23) M(24) M(25) M(26) M(27) M(28) M(29) M(30) M(31) M(32) M(33)

// kFold (compile time, so that the plain variant's epilogue -- the co-bottleneck of this kernel -- is untouched by the extra code):
// 0 = plain (S1 / S2 class tables), 1 = per-frame fold table Ef, 2 = plain fold + affine residual (two-norm composition, vpt_norm2_fold)
// kEw: epilogue warps (8 or 16).  The two-phase epilogue is bound by its dependent instruction chains, not by bandwidth (section 4 of DESIGN.md):
// with 16 warps (4 per scheduler, <= 104 registers) each quarter-tile phase has half the work per thread and twice the latency hiding.
template <int kFold, int kEw>
__global__ void __launch_bounds__(96 + 32 * kEw, 1)
conv3x3_zp_t_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmO,
                    const __grid_constant__ CUtensorMap tmR, const ConvZpTParams p) {
    pdl_sync();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    constexpr uint32_t w_stage_bytes = 128 * kBlockK * 2;  // [128 channels][64 k]
    uint8_t* smem_x = smem;                                 // 2 activation-span stages
    uint8_t* smem_w = smem + 2 * (size_t)p.a_stage_bytes;   // weight tiles
    uint8_t* smem_stage = smem + p.stage_off;               // epi_mode 1: 4 boxes [128 pixels][64 channels] bf16, SWIZZLE_128B
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_w + (size_t)p.b_stages * w_stage_bytes + (p.epi_mode != 0 ? 65536 : 0));
    uint64_t* x_full = bars;
    uint64_t* x_empty = bars + 2;
    uint64_t* w_full = bars + 4;
    uint64_t* w_empty = bars + 4 + kCzMaxBStages;
    uint64_t* tmem_full_bar = bars + 4 + 2 * kCzMaxBStages;
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;
    uint64_t* slot_ready = tmem_empty_bar + 2;  // epi_mode 1: [4 boxes] staging box free (and its residual tile landed)
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(slot_ready + 4);
    float* s_tile0 = reinterpret_cast<float*>(tmem_ptr_smem + 4);  // 2 x [64 pixels][132] fp32: transposed quarter tiles (epi_mode 0)
    float4* s_info2 = reinterpret_cast<float4*>(tmem_ptr_smem + 4);  // epi_mode 1: [2 buffers][2 halves][128 pixels] (ga, gb, cls, frame slot)
    float4* s_info0 = reinterpret_cast<float4*>(s_tile0 + 2 * 64 * kCtPitch);  // 2 x [64]: per-row (ga, gb, cls)

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        if (p.epi_mode != 0) {
            tma_prefetch_desc(&tmO);
            if (p.residual && p.epi_mode == 1) tma_prefetch_desc(&tmR);
        }
        for (int i = 0; i < 4; ++i) mbar_init(&slot_ready[i], 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&x_full[i], 1);
            mbar_init(&x_empty[i], 1);
            mbar_init(&tmem_full_bar[i], 1);
            mbar_init(&tmem_empty_bar[i], kEw);
        }
        for (int i = 0; i < p.b_stages; ++i) {
            mbar_init(&w_full[i], 1);
            mbar_init(&w_empty[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr_smem, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const int halo = p.Wp + 1;

    if (warp == 0) {
        if (lane == 0) {
            // ================= activation-span producer =================
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                const long long span0 = tile * kCtPix - halo;
                for (int cb = 0; cb < p.cin_blocks; ++cb) {
                    if (!(ok = mbar_wait(&x_empty[stage], phase ^ 1u, 0x510u))) break;
                    mbar_expect_tx(&x_full[stage], (uint32_t)p.a_stage_bytes);
                    uint8_t* sx = smem_x + (size_t)stage * p.a_stage_bytes;
                    for (int b = 0; b < p.a_boxes; ++b)
                        tma_load_2d(sx + (size_t)b * p.a_box_rows * 128, &tmX, &x_full[stage], cb * kBlockK, (int)(span0 + (long long)b * p.a_box_rows));
                    advance(stage, phase, 2);
                }
            }
        }
    } else if (warp == 2) {
        if (lane == 0) {
            // ================= weight producer =================
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x) {
                for (int cb = 0; cb < p.cin_blocks && ok; ++cb) {
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!(ok = mbar_wait(&w_empty[stage], phase ^ 1u, 0x520u))) break;
                        mbar_expect_tx(&w_full[stage], w_stage_bytes);
                        tma_load_2d(smem_w + (size_t)stage * w_stage_bytes, &tmW, &w_full[stage], tap * p.cin + cb * kBlockK, 0);
                        advance(stage, phase, p.b_stages);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer: D[channel][pixel] += W_tile[channel][k] * X_span[pixel + shift][k] =================
            const uint32_t idesc = umma_idesc_bf16(128, kCtPix);
            int xstage = 0, wstage = 0;
            uint32_t xphase = 0, wphase = 0;
            int local = 0;
            bool ok = true;
            for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++local) {
                const int as = local & 1;
                const uint32_t accphase = (uint32_t)(local >> 1) & 1u;
                if (!(ok = mbar_wait(&tmem_empty_bar[as], accphase ^ 1u, 0x610u))) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * kAccStageCols);
                for (int cb = 0; cb < p.cin_blocks && ok; ++cb) {
                    if (!(ok = mbar_wait(&x_full[xstage], xphase, 0x710u))) break;
                    tc_fence_after();
                    const uint32_t x_base = smem_u32(smem_x + (size_t)xstage * p.a_stage_bytes);
                    for (int tap = 0; tap < 9; ++tap) {
                        if (!(ok = mbar_wait(&w_full[wstage], wphase, 0x720u))) break;
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(smem_w + (size_t)wstage * w_stage_bytes);
                        const uint32_t x_addr = x_base + (uint32_t)((tap / 3) * p.Wp + (tap % 3)) * 128u;
#pragma unroll
                        for (int k = 0; k < kBlockK / 16; ++k)
                            umma_bf16(d_tmem, umma_desc_sw128(w_addr + k * 32), umma_desc_sw128(x_addr + k * 32), idesc, (uint32_t)((cb | tap | k) != 0));
                        umma_commit(&w_empty[wstage]);
                        advance(wstage, wphase, p.b_stages);
                    }
                    if (!ok) break;
                    umma_commit(&x_empty[xstage]);
                    advance(xstage, xphase, 2);
                }
                if (ok) umma_commit(&tmem_full_bar[as]);
            }
        }
    } else if (kEw == 8 && p.epi_mode == 1) {
        // ================= epilogue v2 (warps 3..10): fragments -> stmatrix.trans -> TMA store =================
        // Round-1 ncu on this kernel: tensor pipe 55-64 % active with the L1/shared pipe at 47 % -- the fp32 transposing store + the
        // row-major second pass cost ~3000 shared/L1 wavefronts per tile on the pipe the UMMA operand fetch needs.  Here the
        // accumulator is read with tcgen05.ld.16x256b, whose register layout is the mma.sync accumulator fragment (thread T: lanes
        // T/4 and T/4+8 = channels, columns 2(T%4), +1 = pixels), folded / ReLU'd / rounded in that layout, and four 8x8 bf16 blocks
        // at a time are written TRANSPOSED by stmatrix into a [pixel][channel] SWIZZLE_128B staging box that one thread hands to TMA
        // (UTMASTG).  The residual tile of the next tile is TMA-prefetched into the same box and read with ldmatrix.trans (same
        // fragment layout); the output overwrites it in place.  ~16 stmatrix + 16 ldmatrix per warp and tile instead of ~380 accesses.
        // Statistics: per-thread sums split by the (at most two) frames a 256-pixel tile touches, one float2 pair per warp and tile.
        const int ew = warp - 3;
        const int quarter = warp & 3;        // TMEM lane quarter = channels [32 quarter, +32)
        const int g = ew >> 2;               // pixel half of the tile: columns [128 g, +128)
        const int box = (quarter >> 1) * 2 + g;  // staging box: channels [64 (quarter/2), +64) x this half's 128 pixels
        uint8_t* stg = smem_stage + (size_t)box * 16384;
        const uint32_t stg_u32 = smem_u32(stg);
        uint64_t* slot = &slot_ready[box];
        const bool leader_t = ((quarter & 1) == 0) && lane == 0;
        const int T4 = lane >> 2, tq = lane & 3;
        // fold tables of the interior border class (cls 4) for this thread's four channels c(hh, u) = 32 quarter + 16 hh + 8 u + T4
        float s1c[4], s2c[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = quarter * 32 + (i >> 1) * 16 + (i & 1) * 8 + T4;
            s1c[i] = p.S1 ? __ldg(p.S1 + 4 * 128 + c) : 0.f;
            s2c[i] = p.S2 ? __ldg(p.S2 + 4 * 128 + c) : 0.f;
        }
        auto setup_slot = [&](long long tile) {  // box leader: the box is free -> arm it for `tile`
            if (tile >= p.num_tiles) return;
            if (p.residual) {
                mbar_expect_tx(slot, 16384u);
                tma_load_2d(stg, &tmR, slot, (quarter >> 1) * 64, (int)(tile * kCtPix + 128 * g));
            } else {
                mbar_arrive(slot);
            }
        };
        if (leader_t) setup_slot(blockIdx.x);
        // ldmatrix / stmatrix row address of this lane inside a 16-pixel block: matrix m = lane / 8 <-> (u = m & 1: channels +8, pxo = m / 2:
        // pixels +8), row = lane % 8; 16-byte chunk = (quarter & 1) * 4 + hh * 2 + u of the 128-byte row, XOR-swizzled with the row index
        const int m_u = (lane >> 3) & 1, m_pxo = lane >> 4, m_row = lane & 7;
        int local = 0;
        bool ok = true;
        for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const uint32_t accphase = (uint32_t)(local >> 1) & 1u;
            const long long q0 = tile * kCtPix + 128 * g;
            const unsigned fA = (unsigned)((tile * kCtPix) / p.FS);  // first frame this tile touches (the other one, if any, is fA + 1)
            float4* s_info = s_info2 + ((local & 1) * 2 + g) * 128;
            {   // per-pixel constants of this half: thread (quarter, lane) fills pixel 32 quarter + lane
                const long long q = q0 + quarter * 32 + lane;
                float4 info = make_float4(1.f, 0.f, -2.f, 0.f);
                if (q < p.Q) {
                    const unsigned qq = (unsigned)q, f = qq / (unsigned)p.FS, r = qq - f * (unsigned)p.FS;
                    const int y = (int)(r / (unsigned)p.Wp), x = (int)r - y * p.Wp;
                    info.w = (float)(f - fA);
                    if (y < p.H && x < p.W) {
                        const int cy = (y == 0) ? 0 : ((y == p.H - 1) ? 2 : 1);
                        const int cx = (x == 0) ? 0 : ((x == p.W - 1) ? 2 : 1);
                        float ga = 1.f, gb = 0.f;
                        if (p.mr) {
                            const float mean = __ldg(p.mr + 2 * f), rstd = __ldg(p.mr + 2 * f + 1);
                            ga = rstd;
                            gb = rstd * mean;
                        }
                        info = make_float4(ga, gb, (float)(cy * 3 + cx), info.w);
                    } else {
                        info.z = -1.f;
                    }
                }
                s_info[quarter * 32 + lane] = info;
            }
            named_bar_sync(5 + g, 128);  // the half's info table is complete (double buffered: no second barrier needed)
            float sA = 0.f, ssA = 0.f, sB = 0.f, ssB = 0.f;
            if (!(ok = mbar_wait(&tmem_full_bar[as], accphase, 0x810u))) break;
            tc_fence_after();
            if (!mbar_wait(slot, (uint32_t)local & 1u, 0x820u)) asm volatile("trap;");  // (a break would desynchronise the named barriers)
            for (int pb = 0; pb < 8; ++pb) {  // 16-pixel blocks
                uint32_t acc[2][8];
                const uint32_t tcol = (uint32_t)(as * kAccStageCols + 128 * g + 16 * pb);
                tmem_ld_16x256b_x2(tmem_base + ((uint32_t)(quarter * 32) << 16) + tcol, acc[0]);
                tmem_ld_16x256b_x2(tmem_base + ((uint32_t)(quarter * 32 + 16) << 16) + tcol, acc[1]);
                tmem_ld_wait();
                if (pb == 7) {  // accumulator stage fully read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                }
                // this lane's ldmatrix / stmatrix row: pixel 16 pb + 8 m_pxo + m_row of the half
                const int mpx = 16 * pb + 8 * m_pxo + m_row;
                const uint32_t row_addr = stg_u32 + (uint32_t)mpx * 128u;
                float4 inf[4];  // pixels 16 pb + 8 pxo + 2 tq + e, index pxo * 2 + e
#pragma unroll
                for (int i = 0; i < 4; ++i) inf[i] = s_info[16 * pb + 8 * (i >> 1) + 2 * tq + (i & 1)];
                float ps[4] = {0.f, 0.f, 0.f, 0.f}, pss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t chunk = (uint32_t)((quarter & 1) * 4 + hh * 2 + m_u);
                    const uint32_t maddr = row_addr + ((chunk ^ (uint32_t)(mpx & 7)) << 4);
                    uint32_t res[4] = {0u, 0u, 0u, 0u};
                    if (p.residual) ldsm_x4_trans(maddr, res);
                    uint32_t outp[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {  // matrix m: u = m & 1 (channel +8), pxo = m >> 1 (pixel +8)
                        const int u = m & 1, pxo = m >> 1;
                        float v[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float4 in = inf[pxo * 2 + e];
                            const int cls = (int)in.z;
                            float a1 = s1c[hh * 2 + u], a2 = s2c[hh * 2 + u];
                            if (cls != 4 && cls >= 0) {
                                const int c = quarter * 32 + hh * 16 + u * 8 + T4;
                                a1 = p.S1 ? __ldg(p.S1 + cls * 128 + c) : 0.f;
                                a2 = p.S2 ? __ldg(p.S2 + cls * 128 + c) : 0.f;
                            }
                            float x = fmaf(in.x, __uint_as_float(acc[hh][pxo * 4 + u * 2 + e]), fmaf(-in.y, a1, a2));
                            if (p.relu == 1) x = fmaxf(x, 0.f);
                            if (p.residual) x += e ? bf16_hi(res[m]) : bf16_lo(res[m]);
                            if (p.relu == 2) x = fmaxf(x, 0.f);
                            v[e] = cls >= 0 ? x : 0.f;  // zero row / column of the ZP layout (and rows past the tensor, which TMA clips)
                        }
                        outp[m] = pack_bf16(v[0], v[1]);
                        const float lo = bf16_lo(outp[m]), hi = bf16_hi(outp[m]);
                        ps[pxo * 2] += lo;
                        pss[pxo * 2] = fmaf(lo, lo, pss[pxo * 2]);
                        ps[pxo * 2 + 1] += hi;
                        pss[pxo * 2 + 1] = fmaf(hi, hi, pss[pxo * 2 + 1]);
                    }
                    stsm_x4_trans(maddr, outp);
                }
                if (p.stat_part) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float fb = inf[i].w;  // 0: first frame of the tile, 1: second
                        sB = fmaf(fb, ps[i], sB);
                        ssB = fmaf(fb, pss[i], ssB);
                        sA = fmaf(1.f - fb, ps[i], sA);
                        ssA = fmaf(1.f - fb, pss[i], ssA);
                    }
                }
            }
            fence_proxy_async();               // generic-proxy writes -> visible to the TMA store
            named_bar_sync(1 + box, 64);       // both channel quarters of the box are in shared memory
            if (leader_t) {
                tma_store_2d(&tmO, stg, (quarter >> 1) * 64, (int)q0);  // rows >= Q are clipped
                bulk_commit();
                bulk_wait_read<0>();
                setup_slot(tile + gridDim.x);
            }
            if (p.stat_part) {
                sA = warp_sum(sA); ssA = warp_sum(ssA); sB = warp_sum(sB); ssB = warp_sum(ssB);
                if (lane == 0) {
                    float2* sp = reinterpret_cast<float2*>(p.stat_part) + ((size_t)tile * 8 + ew) * 2;
                    sp[0] = make_float2(sA, ssA);
                    sp[1] = make_float2(sB, ssB);
                }
            }
        }
        if (leader_t) bulk_wait_all<0>();
    } else if (kEw == 8 && p.epi_mode == 2) {
        // ================= epilogue v3 (warps 3..10): channel-major single pass -> bf16 staging -> TMA store =================
        // Thread = output channel (its TMEM lane), warp = 32 channels x one 128-pixel half.  Everything that depends on the PIXEL
        // (frame statistics, border class, zero row / column) is warp-uniform in this layout, so the fold is one FFMA per value with
        // per-channel constants in registers; the rare non-interior pixels are patched under uniform branches.  The 32 values of a
        // chunk are processed as straight-line passes (the first version interleaved the per-pixel branches with the arithmetic
        // and was latency bound at 2 warps per scheduler: 4.7 ms instead of 2.1).  Values are rounded to bf16 in registers;
        // neighbouring lanes (channels 2i, 2i+1) exchange one packed pair per two pixels so that each lane owns a 32-bit (channel pair)
        // word, stored into a [pixel][64 channels] SWIZZLE_128B box handed to TMA (UTMASTG) in two 64-pixel halves: the first half's
        // store overlaps the second half's arithmetic and nobody ever waits for a store issued less than half a tile ago.
        // Lane 2i takes pixel j, lane 2i+1 pixel j+4: the two 64-byte row segments land in disjoint banks (XOR swizzle flips chunk
        // bit 2).  The residual comes straight from global memory in the same (channel pair) word layout, one chunk ahead.
        // Shared-memory traffic per tile: 64 KB written + 64 KB read by TMA instead of the 128 KB fp32 transposed tile written and
        // re-read, and no LSU global stores.
        const int ew = warp - 3;
        const int quarter = warp & 3;        // TMEM lane quarter = channels [32 quarter, +32)
        const int g = ew >> 2;               // pixel half of the tile: columns [128 g, +128)
        const int ch = quarter * 32 + lane;
        const int box = (quarter >> 1) * 2 + g;  // staging box: channels [64 (quarter/2), +64) x this half's 128 pixels
        uint8_t* stg = smem_stage + (size_t)box * 16384;
        const bool leader_t = ((quarter & 1) == 0) && lane == 0;
        const int odd = lane & 1, pi = lane >> 1;
        // byte offset of this lane's word inside row (8 m' + k) [+4 for odd lanes] of the box, k = 0..3
        const uint32_t cl = (uint32_t)((quarter & 1) * 4 + (pi >> 2)) ^ (uint32_t)(odd << 2);
        uint32_t off[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) off[k] = smem_u32(stg) + (uint32_t)odd * 512u + (uint32_t)k * 128u + ((cl ^ (uint32_t)k) << 4) + (uint32_t)(pi & 3) * 4u;
        const uint32_t* res_w = reinterpret_cast<const uint32_t*>(p.residual) + (quarter * 16 + pi);  // this lane's channel pair, [pixel] stride 64 words
        // fold-table entries of the three classes an ordinary row has: 3 (x = 0), 4 (interior), 5 (x = W - 1)
        float s1k[3], s2k[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            s1k[k] = (kFold != 1 && p.S1) ? __ldg(p.S1 + (3 + k) * 128 + ch) : 0.f;
            s2k[k] = (kFold != 1 && p.S2) ? __ldg(p.S2 + (3 + k) * 128 + ch) : 0.f;
        }
        // frame state (warp-uniform) and this channel's constants for it
        unsigned fcur = 0xffffffffu;
        float ga = 1.f, gb = 0.f, eI = 0.f, e3 = 0.f, e5 = 0.f, raI = 1.f, rbI = 0.f;
        auto load_frame = [&](unsigned f) {
            fcur = f;
            if (p.mr) {
                const float mean = __ldg(p.mr + 2 * (size_t)f), rstd = __ldg(p.mr + 2 * (size_t)f + 1);
                ga = rstd;
                gb = rstd * mean;
            }
            if (kFold == 1) {
                const float* e = p.Ef + ((size_t)f * 9 + 3) * 128 + ch;
                e3 = __ldg(e); eI = __ldg(e + 128); e5 = __ldg(e + 256);
            } else {
                e3 = fmaf(-gb, s1k[0], s2k[0]); eI = fmaf(-gb, s1k[1], s2k[1]); e5 = fmaf(-gb, s1k[2], s2k[2]);
            }
            if (kFold == 2) {
                raI = __ldg(p.res_scale + (size_t)f * 128 + ch);
                rbI = __ldg(p.res_shift + (size_t)f * 128 + ch);
            }
        };
        uint32_t rw_next[16];  // residual words of the NEXT chunk (pixel pairs (ja, ja + 4): even lanes hold pixel ja, odd lanes pixel jb)
#pragma unroll
        for (int i = 0; i < 16; ++i) rw_next[i] = 0u;
        auto prefetch_res = [&](long long qc) {  // chunk starting at ZP pixel qc
            if (!p.residual) return;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const long long q = qc + (i >> 2) * 8 + (i & 3) + 4 * odd;
                rw_next[i] = q < p.Q ? __ldg(res_w + (size_t)q * 64) : 0u;
            }
        };
        if ((long long)blockIdx.x < p.num_tiles) prefetch_res((long long)blockIdx.x * kCtPix + 128 * g);
        int local = 0;
        bool ok = true;
        for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const uint32_t accphase = (uint32_t)(local >> 1) & 1u;
            const long long qh = tile * kCtPix + 128 * g;
            const unsigned fA = (unsigned)((tile * kCtPix) / p.FS);  // first frame this tile touches (the other one, if any, is fA + 1)
            const unsigned fh = qh < p.Q ? (unsigned)(((qh + 31 < p.Q) ? qh + 31 : p.Q - 1) / p.FS) : fA;  // frame of the first chunk
            if (fh != fcur) load_frame(fh);  // (before the accumulator wait: the loads overlap it)
            float s0 = 0.f, ss0 = 0.f, s1v = 0.f, ss1v = 0.f;  // statistics of the tile's first / second frame
            if (!(ok = mbar_wait(&tmem_full_bar[as], accphase, 0x810u))) break;
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {  // 32-pixel chunks of the half
                // Lane-parallel pixel classification: code 0..8 border class (4 = interior), -1 zero row / column, -2 past the tensor.
                // Maps are >= 32 wide (host check), so the valid pixels of one chunk all belong to ONE frame: that of its last pixel
                // (bottom row of a frame and top row of the next are W + 3 > 32 pixels apart).
                const long long qc = qh + 32 * c;
                int code_l = -2, x_l = 0;
                bool plain_l = false;  // pixel of an ordinary row (1 <= y <= H - 2): classes 3 / 4 / 5 and the zero column only
                {
                    const long long q = qc + lane;
                    if (q < p.Q) {
                        const unsigned qq = (unsigned)q;
                        const unsigned f_l = qq / (unsigned)p.FS;
                        const unsigned r = qq - f_l * (unsigned)p.FS;
                        const int y = (int)(r / (unsigned)p.Wp), x = (int)r - y * p.Wp;
                        x_l = x;
                        plain_l = y >= 1 && y <= p.H - 2;
                        code_l = -1;
                        if (y < p.H && x < p.W) code_l = ((y == 0) ? 0 : ((y == p.H - 1) ? 2 : 1)) * 3 + ((x == 0) ? 0 : ((x == p.W - 1) ? 2 : 1));
                    }
                }
                unsigned special = __ballot_sync(0xffffffffu, code_l != 4);
                unsigned zmask = __ballot_sync(0xffffffffu, code_l < 0);
                // Ordinary rows (93 % of the chunks): the non-interior pixels are one (possibly cut) run x = W-1, W, 0 = class 5, zero, class 3
                // at chunk positions jt, jt + 1, jt + 2 (W >= 33: the next run is out of the chunk); jsw = jt + 2 in [0, 33], 34 = none.
                const bool plain = __ballot_sync(0xffffffffu, !plain_l) == 0u;
                int jsw = 34;
                if (plain) {
                    const int x0 = __shfl_sync(0xffffffffu, x_l, 0);
                    const int jt = x0 == 0 ? -2 : p.W - 1 - x0;  // (x0 == W: -1)
                    jsw = jt + 2 < 34 ? jt + 2 : 34;
                }
                const long long qe = (qc + 31 < p.Q) ? qc + 31 : p.Q - 1;
                unsigned fe = qc < p.Q ? (unsigned)(qe / p.FS) : fcur;
                if (p.dbg_skip_epilogue & 32) { special = 0u; zmask = 0u; fe = fcur; jsw = 34; }
                if (fe != fcur) load_frame(fe);
                uint32_t rw[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) rw[i] = rw_next[i];
                {   // residual of the next chunk (of this tile, or the first one of this CTA's next tile)
                    const long long tn = tile + gridDim.x;
                    if (c < 3) prefetch_res(qc + 32);
                    else if (tn < p.num_tiles) prefetch_res(tn * kCtPix + 128 * g);
                }
                uint32_t acc[32];
                if (p.dbg_skip_epilogue & 128) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = (uint32_t)(j + c);
                } else {
                    tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccStageCols + 128 * g + 32 * c), acc);
                    tmem_ld_wait();
                }
                if (c == 3) {  // accumulator stage fully read
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                }
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) x[j] = fmaf(ga, __uint_as_float(acc[j]), eI);
                if (plain) {
                    switch (jsw) {
#define VPT_V3_PATCH(k) case k: v3_patch<k - 2>(x, acc, ga, e5); v3_patch<k>(x, acc, ga, e3); break;
                        VPT_V3_CASES(VPT_V3_PATCH)
#undef VPT_V3_PATCH
                        default: break;
                    }
                } else if (special != 0u) {  // border rows / frame ends: per-pixel uniform branches (7 % of the chunks)
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if ((special >> j) & 1u) {
                            const int code = __shfl_sync(0xffffffffu, code_l, j);
                            if (code >= 0) {
                                float e;
                                if (kFold == 1) e = __ldg(p.Ef + ((size_t)fcur * 9 + code) * 128 + ch);
                                else e = fmaf(-gb, p.S1 ? __ldg(p.S1 + code * 128 + ch) : 0.f, p.S2 ? __ldg(p.S2 + code * 128 + ch) : 0.f);
                                x[j] = fmaf(ga, __uint_as_float(acc[j]), e);
                            }
                        }
                    }
                }
                if (p.relu == 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (p.residual) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int ja = (i >> 2) * 8 + (i & 3), jb = ja + 4;
                        const uint32_t w = rw[i];  // (even channel, odd channel) of pixel ja (even lanes) / jb (odd lanes)
                        const uint32_t recv_r = __shfl_xor_sync(0xffffffffu, w, 1);
                        // own channel: low half in even lanes, high half in odd lanes; pixel ja comes from the even lane's word
                        const uint32_t wa = odd ? recv_r : w, wb = odd ? w : recv_r;
                        const float ra = odd ? bf16_hi(wa) : bf16_lo(wa), rb = odd ? bf16_hi(wb) : bf16_lo(wb);
                        x[ja] += (kFold == 2) ? fmaf(raI, ra, rbI) : ra;
                        x[jb] += (kFold == 2) ? fmaf(raI, rb, rbI) : rb;
                    }
                }
                if (p.relu == 2) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
                }
                if (plain) {  // the zero column
                    switch (jsw) {
#define VPT_V3_ZERO(k) case k: v3_zero<k - 1>(x); break;
                        VPT_V3_CASES(VPT_V3_ZERO)
#undef VPT_V3_ZERO
                        default: break;
                    }
                } else if (zmask != 0u) {  // zero row / column of the ZP layout (and rows past the tensor, which TMA clips)
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = ((zmask >> j) & 1u) ? 0.f : x[j];
                }
                float cs[4] = {0.f, 0.f, 0.f, 0.f}, css[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int ja = (i >> 2) * 8 + (i & 3), jb = ja + 4;
                    const uint32_t pk = pack_bf16(x[ja], x[jb]);  // own channel: (pixel ja, pixel jb), rounded
                    const float lo = bf16_lo(pk), hi = bf16_hi(pk);
                    cs[i & 3] += lo + hi;
                    css[i & 3] = fmaf(lo, lo, fmaf(hi, hi, css[i & 3]));
                    const uint32_t recv_o = __shfl_xor_sync(0xffffffffu, pk, 1);
                    // even lane: word of pixel ja = (own ja, partner's ja); odd lane: word of pixel jb = (partner's jb, own jb)
                    const uint32_t word = odd ? __byte_perm(recv_o, pk, 0x7632) : __byte_perm(pk, recv_o, 0x5410);
                    if (!(p.dbg_skip_epilogue & 16))
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(off[i & 3] + (uint32_t)(c * 4096 + (i >> 2) * 1024)), "r"(word) : "memory");
                }
                {
                    const float a = (cs[0] + cs[1]) + (cs[2] + cs[3]), b2 = (css[0] + css[1]) + (css[2] + css[3]);
                    if (fe == fA) { s0 += a; ss0 += b2; } else { s1v += a; ss1v += b2; }
                }
                if ((c & 1) && !(p.dbg_skip_epilogue & 64)) {  // a 64-pixel half-box is complete
                    fence_proxy_async();                   // generic-proxy writes -> visible to the TMA store
                    if (leader_t) bulk_wait_read<0>();     // the previous half-box store (other 64 rows) has left shared memory
                    named_bar_sync(1 + box, 64);           // both channel quarters wrote this half-box; the other one is free again
                    if (leader_t) {
                        tma_store_2d(&tmO, stg + (c >> 1) * 8192, (quarter >> 1) * 64, (int)(qh + 64 * (c >> 1)));  // rows >= Q are clipped
                        bulk_commit();
                    }
                }
            }
            if (p.stat_part) {
                s0 = warp_sum(s0); ss0 = warp_sum(ss0); s1v = warp_sum(s1v); ss1v = warp_sum(ss1v);
                if (lane == 0) {
                    float2* sp = reinterpret_cast<float2*>(p.stat_part) + ((size_t)tile * 8 + ew) * 2;
                    sp[0] = make_float2(s0, ss0);
                    sp[1] = make_float2(s1v, ss1v);
                }
            }
        }
        if (leader_t) bulk_wait_all<0>();
    } else {
        // ================= epilogue (warps 3..10) =================
        // The accumulator is transposed (TMEM lane = output channel, column = pixel).  Measured (tools/conv_bench.py history):
        // any per-element shared-memory LOOKUP in the channel-major phase costs more than the swap gains (the UMMA operand
        // fetch already saturates the shared-memory pipe), while a plain transposing store is nearly free.  So:
        //   phase A (thread = channel): TMEM -> fp32 -> transposed store into a [64 pixels][128 channels] fp32 tile
        //           (32 lanes = 32 consecutive channels of one pixel = one 128-byte wavefront per instruction);
        //   phase B (thread = pixel row x 32 channels): the regular epilogue -- fold with per-row constants and 16-byte table
        //           loads, ReLU, residual, bf16 rounding, statistics, 16-byte global stores.
        // Four 64-pixel quarters per tile keep the tile at 33 KB.
        constexpr int kColsA = 256 / kEw;                  // phase A: pixel columns of a 64-pixel quarter per warp (32 or 16)
        constexpr int kItemsB = 32 / kEw;                  // phase B: (row pair, 16 chunks) passes per warp and quarter (4 or 2)
        const int ew = warp - 3;
        const int quarter = warp & 3;
        const int cgrp = ew >> 2;                          // phase A: which kColsA of the quarter's 64 pixel columns
        const int ch = quarter * 32 + lane;                // phase A: output channel of this thread
        const int et = ew * 32 + lane;                     // 0..32 kEw - 1
        int local = 0;
        bool ok = true;
        for (long long tile = blockIdx.x; tile < p.num_tiles && ok; tile += gridDim.x, ++local) {
            const int as = local & 1;
            const uint32_t accphase = (uint32_t)(local >> 1) & 1u;
            const long long q0 = tile * kCtPix;
            if (!(ok = mbar_wait(&tmem_full_bar[as], accphase, 0x810u))) break;
            tc_fence_after();
            // Phase-B work items: (row, 8-channel chunk).  A warp pass covers 2 rows x 16 chunks, lane = chunk*2 + row bit, so
            // every global access of a pass touches 2 x 256 contiguous bytes (4 L1 wavefronts instead of the 32 of a
            // thread-per-row mapping -- the L1/shared pipe is what the UMMA operand fetch competes for), and the fp32 tile
            // reads are bank-conflict free.  Each warp does 4 passes per quarter; residual rows are prefetched a quarter ahead.
            const int chunk = lane >> 1, c0 = chunk * 8;
            // fold tables of the interior border class (cls 4: ~94 % of the rows) for this thread's 8 channels, kept in registers
            // (two-norm composition: the tables are per FRAME; the registers hold those of the tile's first frame fA, rows of a second
            //  frame -- 6 % of the tiles touch one -- and border rows take the global-load path)
            const size_t fA = kFold ? (size_t)((tile * kCtPix) / p.FS) : 0;
            float s1c[8], s2c[8], rac[kFold == 2 ? 8 : 1], rbc[kFold == 2 ? 8 : 1];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s1c[j] = (kFold != 1 && p.S1) ? __ldg(p.S1 + 4 * 128 + c0 + j) : 0.f;
                s2c[j] = kFold == 1 ? __ldg(p.Ef + (fA * 9 + 4) * 128 + c0 + j) : (p.S2 ? __ldg(p.S2 + 4 * 128 + c0 + j) : 0.f);
                if (kFold == 2) {
                    rac[j] = __ldg(p.res_scale + fA * 128 + c0 + j);
                    rbc[j] = __ldg(p.res_shift + fA * 128 + c0 + j);
                }
            }
            // per-row constants (ga, gb, cls; cls -1 = zero row/column, -2 = beyond the tensor) of quarter hh of tile tl, written by
            // the first 64 epilogue threads into buffer hh & 1 ONE QUARTER AHEAD of its use (the statistics loads overlap phase B)
            auto write_info = [&](long long tl, int hh) {
                if (et >= 64 || tl >= p.num_tiles) return;
                const long long q = tl * kCtPix + hh * 64 + et;
                float4 info = make_float4(1.f, 0.f, -2.f, 0.f);
                if (q < p.Q) {
                    const unsigned qq = (unsigned)q, f = qq / (unsigned)p.FS, r = qq - f * (unsigned)p.FS;
                    const int y = (int)(r / (unsigned)p.Wp), x = (int)r - y * p.Wp;
                    if (y < p.H && x < p.W) {
                        const int cy = (y == 0) ? 0 : ((y == p.H - 1) ? 2 : 1);
                        const int cx = (x == 0) ? 0 : ((x == p.W - 1) ? 2 : 1);
                        float ga = 1.f, gb = 0.f;
                        if (p.mr) {
                            const float mean = __ldg(p.mr + 2 * f), rstd = __ldg(p.mr + 2 * f + 1);
                            ga = rstd;
                            gb = rstd * mean;
                        }
                        info = make_float4(ga, gb, (float)(cy * 3 + cx), kFold ? (float)f : 0.f);  // frame index: exact in fp32 (< 2^24)
                    } else {
                        info.z = -1.f;
                    }
                }
                s_info0[(hh & 1) * 64 + et] = info;
            };
            if (local == 0) write_info(tile, 0);  // later tiles: written during the previous tile's last quarter
            uint4 rres_n[kItemsB];
            auto prefetch_res = [&](int hh) {
                if (!p.residual) return;
#pragma unroll
                for (int i = 0; i < kItemsB; ++i) {
                    const long long q = q0 + hh * 64 + 2 * (ew * kItemsB + i) + (lane & 1);
                    if (q < p.Q) rres_n[i] = __ldg(reinterpret_cast<const uint4*>(p.residual + (size_t)q * 128 + c0));
                }
            };
            prefetch_res(0);
            for (int h = 0; h < 4; ++h) {
                uint4 rres[kItemsB];
#pragma unroll
                for (int i = 0; i < kItemsB; ++i) rres[i] = rres_n[i];
                if (h < 3) prefetch_res(h + 1);
                // the transposed tile and the row-info table are double buffered (quarter h uses buffer h & 1): the single
                // barrier below orders A(h) -> B(h) and, because every warp reaches it only after its own B(h - 1), also
                // B(h - 1) -> A(h + 1) on the same buffer.
                float* s_tile = s_tile0 + (h & 1) * (64 * kCtPitch);
                float4* s_info = s_info0 + (h & 1) * 64;
                if (p.dbg_skip_epilogue != 2) {  // ---- phase A: TMEM -> transposed fp32 tile
                    uint32_t acc[kColsA];
                    const int pl0 = cgrp * kColsA;
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * kAccStageCols + h * 64 + pl0);
                    if constexpr (kColsA == 32) tmem_ld_32x32(taddr, acc);
                    else tmem_ld_32x16(taddr, acc);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < kColsA; ++j) s_tile[(pl0 + j) * kCtPitch + ch] = __uint_as_float(acc[j]);
                }
                if (h == 3) {  // TMEM fully drained: release the accumulator stage to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty_bar[as]);
                }
                named_bar_sync(1, 32 * kEw);
                if (h < 3) write_info(tile, h + 1);
                else write_info(tile + gridDim.x, 0);
                if (p.dbg_skip_epilogue == 2) continue;
                // ---- phase B
#pragma unroll
                for (int i = 0; i < kItemsB; ++i) {
                    const int prow = 2 * (ew * kItemsB + i) + (lane & 1);
                    const long long q = q0 + h * 64 + prow;
                    const float4 info = s_info[prow];
                    const int cls = (int)info.z;
                    float st_s = 0.f, st_ss = 0.f;
                    if (cls >= -1) {
                        uint4 o = make_uint4(0, 0, 0, 0);
                        if (cls >= 0) {
                            const float4 t0 = *reinterpret_cast<const float4*>(s_tile + prow * kCtPitch + c0);
                            const float4 t1 = *reinterpret_cast<const float4*>(s_tile + prow * kCtPitch + c0 + 4);
                            float4 a0 = make_float4(s1c[0], s1c[1], s1c[2], s1c[3]), a1 = make_float4(s1c[4], s1c[5], s1c[6], s1c[7]);
                            float4 b0 = make_float4(s2c[0], s2c[1], s2c[2], s2c[3]), b1 = make_float4(s2c[4], s2c[5], s2c[6], s2c[7]);
                            if (kFold != 1 && cls != 4 && p.S1) {
                                a0 = __ldg(reinterpret_cast<const float4*>(p.S1 + cls * 128 + c0));
                                a1 = __ldg(reinterpret_cast<const float4*>(p.S1 + cls * 128 + c0) + 1);
                            }
                            if (kFold != 1 && cls != 4 && p.S2) {
                                b0 = __ldg(reinterpret_cast<const float4*>(p.S2 + cls * 128 + c0));
                                b1 = __ldg(reinterpret_cast<const float4*>(p.S2 + cls * 128 + c0) + 1);
                            }
                            const size_t fidx = kFold ? (size_t)info.w : 0;
                            if (kFold == 1 && (cls != 4 || fidx != fA)) {  // per-frame fold table: out = ga * acc + Ef[f][cls][c]
                                const float4* e = reinterpret_cast<const float4*>(p.Ef + (fidx * 9 + cls) * 128 + c0);
                                b0 = __ldg(e);
                                b1 = __ldg(e + 1);
                            }
                            float v[8] = {fmaf(info.x, t0.x, fmaf(-info.y, a0.x, b0.x)), fmaf(info.x, t0.y, fmaf(-info.y, a0.y, b0.y)),
                                          fmaf(info.x, t0.z, fmaf(-info.y, a0.z, b0.z)), fmaf(info.x, t0.w, fmaf(-info.y, a0.w, b0.w)),
                                          fmaf(info.x, t1.x, fmaf(-info.y, a1.x, b1.x)), fmaf(info.x, t1.y, fmaf(-info.y, a1.y, b1.y)),
                                          fmaf(info.x, t1.z, fmaf(-info.y, a1.z, b1.z)), fmaf(info.x, t1.w, fmaf(-info.y, a1.w, b1.w))};
                            if (p.relu == 1) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                            }
                            if (p.residual) {
                                const uint4 rr = rres[i];
                                float r8[8] = {bf16_lo(rr.x), bf16_hi(rr.x), bf16_lo(rr.y), bf16_hi(rr.y), bf16_lo(rr.z), bf16_hi(rr.z), bf16_lo(rr.w), bf16_hi(rr.w)};
                                if (kFold == 2) {  // residual stream recomputed from the un-normalised tensor: a[f][c] * r + b[f][c]
                                    if (fidx == fA) {
#pragma unroll
                                        for (int e8 = 0; e8 < 8; ++e8) r8[e8] = fmaf(rac[e8], r8[e8], rbc[e8]);
                                    } else {
                                        const float4* ra = reinterpret_cast<const float4*>(p.res_scale + fidx * 128 + c0);
                                        const float4* rb = reinterpret_cast<const float4*>(p.res_shift + fidx * 128 + c0);
                                        const float4 x0 = __ldg(ra), x1 = __ldg(ra + 1), y0 = __ldg(rb), y1 = __ldg(rb + 1);
                                        r8[0] = fmaf(x0.x, r8[0], y0.x); r8[1] = fmaf(x0.y, r8[1], y0.y); r8[2] = fmaf(x0.z, r8[2], y0.z); r8[3] = fmaf(x0.w, r8[3], y0.w);
                                        r8[4] = fmaf(x1.x, r8[4], y1.x); r8[5] = fmaf(x1.y, r8[5], y1.y); r8[6] = fmaf(x1.z, r8[6], y1.z); r8[7] = fmaf(x1.w, r8[7], y1.w);
                                    }
                                }
#pragma unroll
                                for (int e8 = 0; e8 < 8; ++e8) v[e8] += r8[e8];
                            }
                            if (p.relu == 2) {
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                            }
                            o = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
                            const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float lo = bf16_lo(ow[k]), hi = bf16_hi(ow[k]);
                                st_s += lo + hi;
                                st_ss = fmaf(lo, lo, fmaf(hi, hi, st_ss));
                            }
                        }
                        *reinterpret_cast<uint4*>(p.out + (size_t)q * 128 + c0) = o;
                    }
                    if (p.stat_part) {  // row sums: reduce over the 16 chunk lanes of this row (lane bits 1..4)
#pragma unroll
                        for (int m = 2; m <= 16; m <<= 1) {
                            st_s += __shfl_xor_sync(0xffffffffu, st_s, m);
                            st_ss += __shfl_xor_sync(0xffffffffu, st_ss, m);
                        }
                        if (lane < 2 && cls >= -1) reinterpret_cast<float2*>(p.stat_part)[(size_t)q] = make_float2(st_s, st_ss);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static int g_cz_swap = 1;

// host launcher, called from vpt_conv3x3_zp when Cout == 128
static int launch_conv_zp_t(const vpt_conv_zp_args* a, void* stream) {
    const int H = a->H, W = a->W, C = a->Cin;
    ConvZpTParams p;
    memset(&p, 0, sizeof(p));
    p.H = H; p.W = W; p.Wp = W + 1; p.FS = (H + 1) * (W + 1);
    p.Q = (long long)a->F * p.FS;
    p.cin = C; p.cin_blocks = C / 64;
    p.num_tiles = (p.Q + kCtPix - 1) / kCtPix;
    const int span = kCtPix + 2 * (p.Wp + 1);
    p.a_boxes = (span + 255) / 256;
    p.a_box_rows = ((span + p.a_boxes - 1) / p.a_boxes + 7) / 8 * 8;
    VPT_CHECK(p.a_box_rows <= 256, "vpt_conv3x3_zp: span does not fit the TMA box limit");
    p.a_stage_bytes = p.a_boxes * p.a_box_rows * 128;
    const uint32_t w_stage_bytes = 128 * kBlockK * 2;
    // g_cz_swap: 1 = transposing two-phase epilogue (default: measured faster, see below), 4 = fragment epilogue (tcgen05.ld.16x256b ->
    // stmatrix.trans -> TMA store; needs maps of >= 256 rows per frame so that a tile touches at most two frames), 2 = debug (no epilogue).
    // Round-2 A/B on B200 (tools/conv_bench.py, 128->128 @64x64, 2048 frames, residual): two-phase 2.144 ms (1154 TFLOP/s), fragment
    // 2.207 ms (1121): the fragment path moves the same bytes through shared memory (the residual now arrives there too) and pays ~3x
    // the instructions for the per-value border-class fold, so it stays an experiment.
    const int swap_mode = g_cz_swap & 0xff, bst_cap = (g_cz_swap >> 8) & 0xff, dbg_bits = (g_cz_swap >> 16) & 0xff;  // bits 8+: debug cap on the weight pipeline depth
    const bool tma_epi_ok = p.FS >= kCtPix && (swap_mode != 5 || (W >= 33 && H >= 8)) && ((uintptr_t)a->out & 127) == 0 && (!a->residual || ((uintptr_t)a->residual & 127) == 0);
    p.epi_mode = (swap_mode == 4 && tma_epi_ok) ? 1 : ((swap_mode == 5 && tma_epi_ok) ? 2 : 0);
    const size_t bars_bytes = (4 + 2 * kCzMaxBStages + 4 + 4) * 8 + 16;
    const size_t tail = p.epi_mode == 1 ? 65536 + bars_bytes + 2 * 2 * 128 * 16 + 64
                        : (p.epi_mode == 2 ? 65536 + bars_bytes + 64 : bars_bytes + 2 * 64 * kCtPitch * 4 + 2 * 64 * 16 + 64);
    const long long budget = 225 * 1024 - 1024 - 2 * (long long)p.a_stage_bytes - (long long)tail;
    int bst = (int)(budget / w_stage_bytes);
    if (bst > kCzMaxBStages) bst = kCzMaxBStages;
    if (bst_cap > 0 && bst > bst_cap) bst = bst_cap;
    VPT_CHECK(bst >= 2, "vpt_conv3x3_zp: not enough shared memory for the weight pipeline (W=%d)", W);
    p.b_stages = bst;
    p.stage_off = 2 * p.a_stage_bytes + bst * (int)w_stage_bytes;
    const size_t smem_bytes = 1024 + 2 * (size_t)p.a_stage_bytes + (size_t)bst * w_stage_bytes + tail;
    CUtensorMap tmX, tmW, tmO, tmR;
    memset(&tmO, 0, sizeof(tmO));
    memset(&tmR, 0, sizeof(tmR));
    if (p.epi_mode != 0) {  // output / residual [Q][128] bf16: boxes of 128 pixel rows x 64 channels (128-byte rows, SWIZZLE_128B)
        cuuint64_t dims[2] = {128, (cuuint64_t)p.Q};
        cuuint64_t strides[1] = {256};
        cuuint32_t box[2] = {64, (cuuint32_t)(p.epi_mode == 2 ? 64 : 128)};  // v3 stores half-boxes
        int r = make_tmap_bf16(&tmO, a->out, 2, dims, strides, box);
        if (r) return r;
        if (a->residual && p.epi_mode == 1) {
            r = make_tmap_bf16(&tmR, a->residual, 2, dims, strides, box);
            if (r) return r;
        }
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)p.Q};
        cuuint64_t strides[1] = {(cuuint64_t)C * 2};
        cuuint32_t box[2] = {64, (cuuint32_t)p.a_box_rows};
        int r = make_tmap_bf16(&tmX, a->x, 2, dims, strides, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)9 * C, (cuuint64_t)128};
        cuuint64_t strides[1] = {(cuuint64_t)9 * C * 2};
        cuuint32_t box[2] = {64, 128};
        int r = make_tmap_bf16(&tmW, a->w, 2, dims, strides, box);
        if (r) return r;
    }
    p.mr = a->mr; p.S1 = a->mr ? a->S1 : nullptr; p.S2 = a->S2; p.relu = a->relu;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(a->residual);
    p.out = reinterpret_cast<__nv_bfloat16*>(a->out);
    p.stat_part = a->stat_part;
    p.Ef = a->Ef; p.res_scale = a->res_scale; p.res_shift = a->res_shift;
    VPT_CHECK(!a->Ef || a->mr, "vpt_conv3x3_zp: Ef needs mr = (0, rstd) per frame");
    VPT_CHECK(!a->res_scale == !a->res_shift && (!a->res_scale || a->residual), "vpt_conv3x3_zp: res_scale / res_shift come as a pair, with a residual");
    VPT_CHECK(!(p.epi_mode == 1 && (a->Ef || a->res_scale)), "vpt_conv3x3_zp: the fragment-epilogue experiment (swap mode 4) does not implement Ef / res_scale");
    p.dbg_skip_epilogue = (swap_mode == 2) ? 2 : (dbg_bits & 0xf0);  // v3 experiment bits: 16 no STS, 32 no classification, 64 no TMA store, 128 no LDTM  // 2: no epilogue work (MMA-rate experiment)
    static bool attr_set = false;
    if (!attr_set) {
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        VPT_CUDA(cudaFuncSetAttribute(conv3x3_zp_t_kernel<2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    long long grid = num_sms();
    if (grid <= 0) grid = 148;
    if (grid > p.num_tiles) grid = p.num_tiles;
    VPT_CHECK(!(a->Ef && a->res_scale), "vpt_conv3x3_zp: Ef and res_scale are not combined (Cout == 128 kernel)");
    // 16 epilogue warps for the two-phase epilogue (swap mode 1); swap mode 6 = the same with 8 warps (A-B), the experiments (4, 5) are written for 8
    const bool ew16 = p.epi_mode == 0 && swap_mode != 6 && swap_mode != 2;
    const dim3 g((unsigned)grid), b8(96 + 32 * 8), b16(96 + 32 * 16);
    const cudaStream_t st = (cudaStream_t)stream;
    if (ew16) {
        if (a->Ef) launch_k(conv3x3_zp_t_kernel<1, 16>, g, b16, smem_bytes, st, tmX, tmW, tmO, tmR, p);
        else if (a->res_scale) launch_k(conv3x3_zp_t_kernel<2, 16>, g, b16, smem_bytes, st, tmX, tmW, tmO, tmR, p);
        else launch_k(conv3x3_zp_t_kernel<0, 16>, g, b16, smem_bytes, st, tmX, tmW, tmO, tmR, p);
    } else {
        if (a->Ef) launch_k(conv3x3_zp_t_kernel<1, 8>, g, b8, smem_bytes, st, tmX, tmW, tmO, tmR, p);
        else if (a->res_scale) launch_k(conv3x3_zp_t_kernel<2, 8>, g, b8, smem_bytes, st, tmX, tmW, tmO, tmR, p);
        else launch_k(conv3x3_zp_t_kernel<0, 8>, g, b8, smem_bytes, st, tmX, tmW, tmO, tmR, p);
    }
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

// epi_mode 1 statistics: mr[f] = (mean, rstd) of frame f from the per-(tile, warp, frame slot) partials; one warp per frame, fp64
__global__ void __launch_bounds__(256) conv_zp_t_stats_finalize_kernel(const float2* __restrict__ part, float2* __restrict__ mr, long long F, int FS,
                                                                        long long num_tiles, double inv_count, float eps) {
    pdl_sync();
    const long long f = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= F) return;
    const int lane = threadIdx.x & 31;
    const long long t0 = (f * FS) / kCtPix, t1 = min(num_tiles - 1, ((f + 1) * FS - 1) / kCtPix);
    double s = 0.0, ss = 0.0;
    for (long long i = lane; i < (t1 - t0 + 1) * 8; i += 32) {
        const long long t = t0 + (i >> 3);
        const long long slot = f - (t * kCtPix) / FS;  // which of the tile's (at most two) frames is f
        if (slot == 0 || slot == 1) {
            const float2 v = part[(t * 8 + (i & 7)) * 2 + slot];
            s += (double)v.x;
            ss += (double)v.y;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if (lane == 0) {
        const double mean = s * inv_count;
        double var = ss * inv_count - mean * mean;
        if (var < 0.0) var = 0.0;
        mr[f] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
    }
}

int g_cz_swap_enabled() { return g_cz_swap; }
int launch_conv_zp_t_fwd(const vpt_conv_zp_args* a, void* stream) { return launch_conv_zp_t(a, stream); }

}  // namespace vpt

extern "C" int vpt_set_conv_swap_mode(int32_t on) {
    vpt::g_cz_swap = on;  // 2 = debug: skip the epilogue arithmetic (MMA-rate experiment)
    return VPT_OK;
}

/* Statistics plumbing of the fragment-epilogue kernel (Cout == 128): number of floats of its partial buffer ([tiles][8 warps][2 frame
 * slots] float2), 0 when another kernel / epilogue handles this shape (then vpt_conv_zp_stat_parts + vpt_stats_finalize apply). */
extern "C" int64_t vpt_conv_zp_t_stat_floats(int32_t F, int32_t H, int32_t W, int32_t Cout) {
    const int mode = vpt::g_cz_swap & 0xff;
    if (!vpt::conv_zp_use_swapped((long long)F * (H + 1) * (W + 1))) return 0;
    if (Cout != 128 || (mode != 4 && mode != 5) || (H + 1) * (W + 1) < vpt::kCtPix || (mode == 5 && (W < 33 || H < 8))) return 0;
    const long long Q = (long long)F * (H + 1) * (W + 1);
    return ((Q + vpt::kCtPix - 1) / vpt::kCtPix) * 8 * 2 * 2;
}

extern "C" int vpt_conv_zp_t_stats_finalize(const float* part, float* mr, int32_t F, int32_t H, int32_t W, float eps, void* stream) {
    using namespace vpt;
    VPT_CHECK(part && mr && F > 0, "vpt_conv_zp_t_stats_finalize: null argument");
    const int FS = (H + 1) * (W + 1);
    const long long Q = (long long)F * FS, tiles = (Q + kCtPix - 1) / kCtPix;
    launch_k(conv_zp_t_stats_finalize_kernel, dim3((unsigned)((F + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, 
        reinterpret_cast<const float2*>(part), reinterpret_cast<float2*>(mr), F, FS, tiles, 1.0 / ((double)H * W * 128), eps);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
