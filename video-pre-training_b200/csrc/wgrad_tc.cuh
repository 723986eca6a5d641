// Weight-gradient GEMM with tap pairing:  dW[m][tap*N + n] = sum_k a[k][m] * b[k + shift[tap]][n]   (K = pixels / tokens).
//
// Why a second kernel (round 2): the generic MN-major mode of gemm_tc_kernel runs every (tap, m tile, n tile) as its own GEMM tile,
// so each 64-row K block moves 16 KB of dY + 32 KB of X from L2 for 128 x 256 x 64 MACs, nine times over for the nine taps.  At the
// measured 800 TFLOP/s that is 9.2 TB/s of L2 -> SM traffic -- the chip's L2 throughput (~6300 B/clk), not the tensor pipe, bounds it
// (profiles/bc_step_r2.md).  Here two taps whose row shifts differ by one (dx and dx + 1 of the same kernel row) share ONE activation
// span of 64 + 1 rows in shared memory and ONE dY tile; the second tap's B operand is the same buffer one 128-byte row further
// down (MN-major SWIZZLE_128B: the swizzle is a function of the absolute shared-memory address, so a whole-row offset is a valid
// descriptor start).  Two 128 x block_n accumulators fill the 512 TMEM columns.  Bytes per MAC fall 1.96x for the six paired taps
// (1.48x over all nine); the unpaired taps run as one-tap items of the same kernel with half as many K splits, so that every work
// item has the same number of MMAs.
//
//   warp 0 (one lane): TMA producer   warp 1 (one lane): tcgen05.mma issuer   warps 2..9: epilogue (fp32 partials per K split)
#pragma once
#include "gemm_tc.cuh"

namespace vpt {

constexpr int kWgBoxRows = 65;                 // K rows of a two-tap B box: 64 + 1
constexpr uint32_t kWgBoxBytes = 72 * 128;     // 9216: shared-memory pitch of a {64 columns, <= 72 K rows} box (1024-byte multiple) = LBO of the B descriptor

struct WgradParams {
    int M, N;
    int block_n, num_m_tiles, num_n_tiles;
    int k_iters;
    int num_stages;
    uint32_t stage_bytes_b;
    // work units: npair two-tap units (taps unit_tap[u], unit_tap[u] + 1) then nsingle one-tap units
    int npair, nsingle;
    int unit_tap[9], unit_shift[9];
    int splits_pair, iters_pair;      // K splits of the two-tap items and their length in K blocks
    int splits_single, iters_single;  // one-tap items: half as many, twice as long
    int items_pair, items_total;
    float* out;
    long long ld_out, split_stride;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB72 /* 65-row boxes */, const __grid_constant__ CUtensorMap tmB64,
                const WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.num_stages * kStageBytesA;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.num_stages * p.stage_bytes_b);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kMaxStages;
    uint64_t* tmem_full_bar = bars + 2 * kMaxStages;
    uint64_t* tmem_empty_bar = bars + 2 * kMaxStages + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB72);
        tma_prefetch_desc(&tmB64);
        for (int i = 0; i < p.num_stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, kNumEpiWarps);
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

    // item -> (two taps?, unit, split, m tile, n tile, K range); items of one K split are adjacent so that the CTAs running at the same
    // time read the same rows (L2 locality)
    struct Item {
        int pair, tap, shift, split, m0, n0, it0, it1;
    };
    auto decode = [&](int item) {
        Item w;
        w.pair = item < p.items_pair;
        const int idx = w.pair ? item : item - p.items_pair;
        const int units = w.pair ? p.npair : p.nsingle;
        const int per_split = units * p.num_m_tiles * p.num_n_tiles;
        w.split = idx / per_split;
        int rem = idx - w.split * per_split;
        const int m_tile = rem / (units * p.num_n_tiles);
        rem -= m_tile * units * p.num_n_tiles;
        const int u = rem / p.num_n_tiles + (w.pair ? 0 : p.npair);
        const int n_tile = rem % p.num_n_tiles;
        w.tap = p.unit_tap[u];
        w.shift = p.unit_shift[u];
        w.m0 = m_tile * kBlockM;
        w.n0 = n_tile * p.block_n;
        const int len = w.pair ? p.iters_pair : p.iters_single;
        w.it0 = w.split * len;
        w.it1 = min(p.k_iters, w.it0 + len);
        return w;
    };

    if (warp == 0) {
        if (lane == 0) {
            // ================= TMA producer =================
            int stage = 0;
            uint32_t phase = 0;
            bool ok = true;
            const int nbox = p.block_n / 64;
            for (int item = blockIdx.x; item < p.items_total && ok; item += gridDim.x) {
                const Item w = decode(item);
                const uint32_t b_bytes = (uint32_t)nbox * (w.pair ? (uint32_t)kWgBoxRows * 128u : 8192u);
                for (int it = w.it0; it < w.it1; ++it) {
                    if (!(ok = mbar_wait(&empty_bar[stage], phase ^ 1u, 0x100u))) break;
                    mbar_expect_tx(&full_bar[stage], kStageBytesA + b_bytes);
                    uint8_t* sa = smem_a + (size_t)stage * kStageBytesA;
                    uint8_t* sb = smem_b + (size_t)stage * p.stage_bytes_b;
                    tma_load_2d(sa, &tmA, &full_bar[stage], w.m0, it * kBlockK);
                    tma_load_2d(sa + 8192, &tmA, &full_bar[stage], w.m0 + 64, it * kBlockK);
                    for (int bx = 0; bx < nbox; ++bx)
                        tma_load_2d(sb + (size_t)bx * kWgBoxBytes, w.pair ? &tmB72 : &tmB64, &full_bar[stage], w.n0 + bx * 64, it * kBlockK + w.shift);
                    advance(stage, phase, p.num_stages);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ================= MMA issuer =================
            const uint32_t idesc = umma_idesc_bf16_mn(kBlockM, p.block_n);
            int stage = 0;
            uint32_t phase = 0;
            int local = 0;
            bool ok = true;
            for (int item = blockIdx.x; item < p.items_total && ok; item += gridDim.x, ++local) {
                const Item w = decode(item);
                if (!(ok = mbar_wait(tmem_empty_bar, ((uint32_t)local & 1u) ^ 1u, 0x200u))) break;
                tc_fence_after();
                for (int it = w.it0; it < w.it1; ++it) {
                    if (!(ok = mbar_wait(&full_bar[stage], phase, 0x300u))) break;
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + (size_t)stage * kStageBytesA);
                    const uint32_t b_addr = smem_u32(smem_b + (size_t)stage * p.stage_bytes_b);
                    const uint32_t acc = (uint32_t)(it != w.it0);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k) {  // 16 K rows of 128 B per instruction
                        const uint64_t da = umma_desc_sw128_mn(a_addr + k * 2048, 8192u, 1024u);
                        umma_bf16(tmem_base, da, umma_desc_sw128_mn(b_addr + k * 2048, kWgBoxBytes, 1024u), idesc, acc | (uint32_t)(k != 0));
                        if (w.pair)  // second tap: the same span one K row (128 B) further down
                            umma_bf16(tmem_base + kAccStageCols, da, umma_desc_sw128_mn(b_addr + 128u + k * 2048, kWgBoxBytes, 1024u), idesc,
                                      acc | (uint32_t)(k != 0));
                    }
                    umma_commit(&empty_bar[stage]);
                    advance(stage, phase, p.num_stages);
                }
                if (ok) umma_commit(tmem_full_bar);
            }
        }
    } else {
        // ================= epilogue: fp32 partial sums of this K split =================
        const int ew = warp - 2;
        const int quarter = warp & 3;  // TMEM lane quarter this warp may access
        const int half = ew >> 2;
        const int nchunks = (p.block_n + 31) >> 5;
        int local = 0;
        bool ok = true;
        for (int item = blockIdx.x; item < p.items_total && ok; item += gridDim.x, ++local) {
            const Item w = decode(item);
            if (!(ok = mbar_wait(tmem_full_bar, (uint32_t)local & 1u, 0x400u))) break;
            tc_fence_after();
            const int m = w.m0 + quarter * 32 + lane;
            // two taps: warps 2..5 store tap 0's accumulator, warps 6..9 tap 1's; one tap: the halves split its column chunks
            const int tap = w.tap + (w.pair ? half : 0);
            const uint32_t col0 = (uint32_t)(w.pair ? half * kAccStageCols : 0);
            const int c_begin = w.pair ? 0 : (half == 0 ? 0 : (nchunks + 1) >> 1);
            const int c_end = w.pair ? nchunks : (half == 0 ? (nchunks + 1) >> 1 : nchunks);
            float* orow = p.out + (size_t)w.split * p.split_stride + (size_t)m * p.ld_out + (size_t)tap * p.N;
            for (int c = c_begin; c < c_end; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(tmem_base + ((uint32_t)(quarter * 32) << 16) + col0 + (uint32_t)(c * 32), v);
                tmem_ld_wait();
                const int nb = w.n0 + c * 32;
                if (m < p.M) {
                    if (nb + 32 <= p.N) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            reinterpret_cast<float4*>(orow + nb)[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                                                                  __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (nb + j < p.N) orow[nb + j] = __uint_as_float(v[j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// out[m][tap*N + n] = sum over the K splits of that tap (paired taps have splits_pair partials, the others splits_single); 4 floats per thread
__global__ void __launch_bounds__(256) wgrad_sum_splits_kernel(const float4* __restrict__ part, float4* __restrict__ out, long long n4, int row4, int n4_per_tap,
                                                               int tap_splits_mask_pair, int splits_pair, int splits_single) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const int tap = (int)(i % row4) / n4_per_tap;
        const int splits = ((tap_splits_mask_pair >> tap) & 1) ? splits_pair : splits_single;
        float4 a = __ldg(part + i);
        for (int s = 1; s < splits; ++s) {
            const float4 b = __ldg(part + (long long)s * n4 + i);
            a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        }
        out[i] = a;
    }
}

struct WgradPairPlan {
    int block_n, n_tiles, m_tiles, k_iters;
    int npair, nsingle, unit_tap[9], unit_shift[9], pair_mask;
    int splits_pair, iters_pair, splits_single, iters_single;
    int max_splits;
};

static WgradPairPlan wgrad_pair_plan(int M, int N, int ntaps, const int32_t* shifts, long long R) {
    WgradPairPlan w;
    memset(&w, 0, sizeof(w));
    // whole 64-column TMA boxes; among 256 / 192 / 128 pick the width that pads N the least (N = 384 -> 2 x 192, not 256 + 128)
    if (N <= 256) {
        w.block_n = (N + 63) / 64 * 64;
    } else {
        int best = 256, best_pad = (N + 255) / 256 * 256 - N;
        for (int bn = 192; bn >= 128; bn -= 64) {
            const int pad = (N + bn - 1) / bn * bn - N;
            if (pad < best_pad) { best = bn; best_pad = pad; }
        }
        w.block_n = best;
    }
    w.n_tiles = (N + w.block_n - 1) / w.block_n;
    w.m_tiles = (M + kBlockM - 1) / kBlockM;
    w.k_iters = (int)((R + kBlockK - 1) / kBlockK);
    // pair up taps whose shifts differ by exactly one row (dx, dx + 1 of a kernel row); two-tap units first
    int single_tap[9], single_shift[9];
    for (int t = 0; t < ntaps;) {
        if (t + 1 < ntaps && shifts[t + 1] == shifts[t] + 1) {
            w.unit_tap[w.npair] = t;
            w.unit_shift[w.npair] = shifts[t];
            w.pair_mask |= 3 << t;
            ++w.npair;
            t += 2;
        } else {
            single_tap[w.nsingle] = t;
            single_shift[w.nsingle] = shifts[t];
            ++w.nsingle;
            t += 1;
        }
    }
    for (int i = 0; i < w.nsingle; ++i) {
        w.unit_tap[w.npair + i] = single_tap[i];
        w.unit_shift[w.npair + i] = single_shift[i];
    }
    // K splits: a one-tap item runs twice as many K blocks as a two-tap item, i.e. the same number of MMAs; ~2-3 items per SM
    const int mn = w.m_tiles * w.n_tiles;
    const int weight = mn * (2 * w.npair + w.nsingle);  // items per unit of "single" split count if pairs get twice the splits
    int ss = (3 * num_sms() + weight - 1) / weight;
    const int max_ss = w.k_iters / 16 > 0 ? w.k_iters / 16 : 1;  // at least 8 K blocks per two-tap item
    if (ss > max_ss) ss = max_ss;
    if (ss > 32) ss = 32;
    if (ss < 1) ss = 1;
    w.iters_single = (w.k_iters + ss - 1) / ss;
    w.splits_single = (w.k_iters + w.iters_single - 1) / w.iters_single;
    w.iters_pair = (w.iters_single + 1) / 2;
    w.splits_pair = (w.k_iters + w.iters_pair - 1) / w.iters_pair;
    if (w.npair == 0) { w.splits_pair = 1; w.iters_pair = w.k_iters; }
    if (w.nsingle == 0) { w.splits_single = 1; w.iters_single = w.k_iters; }
    w.max_splits = w.splits_pair > w.splits_single ? w.splits_pair : w.splits_single;
    if (w.npair == 0) w.max_splits = w.splits_single;
    if (w.nsingle == 0) w.max_splits = w.splits_pair;
    return w;
}

static int g_wgrad_mode = 1;  // 1: tap-pairing kernel (this file), 0: generic MN-major mode of gemm_tc_kernel (round 1)
int wgrad_mode() { return g_wgrad_mode; }

// largest number of K-split partials any shift pattern can need (the workspace query does not see the shifts)
long long wgrad_pair_max_splits(int M, int N, int ntaps, long long R) {
    int32_t adj[9], far[9];
    for (int t = 0; t < 9; ++t) { adj[t] = (t / 3) * 1000 + t % 3; far[t] = t * 1000; }
    const WgradPairPlan a = wgrad_pair_plan(M, N, ntaps, adj, R), b = wgrad_pair_plan(M, N, ntaps, far, R);
    int32_t all[9];
    for (int t = 0; t < 9; ++t) all[t] = t;  // every tap adjacent to the next
    const WgradPairPlan c = wgrad_pair_plan(M, N, ntaps, all, R);
    int m = a.max_splits > b.max_splits ? a.max_splits : b.max_splits;
    return m > c.max_splits ? m : c.max_splits;
}

int launch_wgrad_pair(const void* a, int64_t lda, const void* b, int64_t ldb, int32_t M, int32_t N, int64_t R, const int32_t* shifts,
                             int32_t ntaps, float* out, void* workspace, int64_t workspace_bytes, void* stream) {
    const WgradPairPlan w = wgrad_pair_plan(M, N, ntaps, shifts, R);
    const long long out_elems = (long long)M * N * ntaps;
    VPT_CHECK(w.max_splits == 1 || (workspace && workspace_bytes >= (int64_t)w.max_splits * out_elems * 4),
              "vpt_wgrad_bf16: workspace too small (%lld bytes, need %lld)", (long long)workspace_bytes, (long long)w.max_splits * out_elems * 4);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N;
    p.block_n = w.block_n; p.num_m_tiles = w.m_tiles; p.num_n_tiles = w.n_tiles; p.k_iters = w.k_iters;
    p.npair = w.npair; p.nsingle = w.nsingle;
    for (int i = 0; i < 9; ++i) { p.unit_tap[i] = w.unit_tap[i]; p.unit_shift[i] = w.unit_shift[i]; }
    p.splits_pair = w.splits_pair; p.iters_pair = w.iters_pair; p.splits_single = w.splits_single; p.iters_single = w.iters_single;
    const int mn = w.m_tiles * w.n_tiles;
    p.items_pair = w.npair * mn * w.splits_pair;
    p.items_total = p.items_pair + w.nsingle * mn * w.splits_single;
    p.out = w.max_splits > 1 ? reinterpret_cast<float*>(workspace) : out;
    p.ld_out = (long long)N * ntaps;
    p.split_stride = out_elems;
    CUtensorMap tmA, tmB72, tmB64;
    {   // MN-major operands: the tensor map's inner dimension is the operand's M (N) index, its rows are K
        cuuint64_t dims[2] = {(cuuint64_t)M, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)lda * 2};
        cuuint32_t box[2] = {64, 64};
        int r = make_tmap_bf16(&tmA, a, 2, dims, strides, box);
        if (r) return r;
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)R};
        cuuint64_t strides[1] = {(cuuint64_t)ldb * 2};
        cuuint32_t box72[2] = {64, (cuuint32_t)kWgBoxRows}, box64[2] = {64, 64};
        int r = make_tmap_bf16(&tmB72, b, 2, dims, strides, box72);
        if (r) return r;
        r = make_tmap_bf16(&tmB64, b, 2, dims, strides, box64);
        if (r) return r;
    }
    p.stage_bytes_b = (uint32_t)(w.block_n / 64) * kWgBoxBytes;
    const uint32_t stage_bytes = kStageBytesA + p.stage_bytes_b;
    const size_t bars_bytes = (2 * kMaxStages + 2) * 8 + 16;
    int stages = (int)((225 * 1024 - 1024 - bars_bytes) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    VPT_CHECK(stages >= 2, "vpt_wgrad_bf16: not enough shared memory for two pipeline stages");
    p.num_stages = stages;
    const size_t smem_bytes = 1024 + (size_t)stages * stage_bytes + bars_bytes;
    static bool attr_set = false;
    if (!attr_set) {
        VPT_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set = true;
    }
    const int grid = p.items_total < num_sms() ? p.items_total : num_sms();
    wgrad_tc_kernel<<<grid, kGemmThreads, smem_bytes, (cudaStream_t)stream>>>(tmA, tmB72, tmB64, p);
    VPT_LAUNCH_CHECK();
    if (w.max_splits > 1) {
        const long long n4 = out_elems / 4;
        long long blocks = (n4 + 255) / 256;
        if (blocks > 4096) blocks = 4096;
        wgrad_sum_splits_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(workspace), reinterpret_cast<float4*>(out),
                                                                                    n4, N * ntaps / 4, N / 4, w.pair_mask, w.splits_pair, w.splits_single);
        VPT_LAUNCH_CHECK();
    }
    return VPT_OK;
}

}  // namespace vpt

extern "C" int vpt_set_wgrad_mode(int32_t mode) {
    vpt::g_wgrad_mode = mode;
    return VPT_OK;
}
