// Banded masked self-attention over [KV memory | chunk] with the learned relative-position bias, flash style:
// warp-level mma.sync (m16n8k16 bf16, fp32 accumulate), online fp32 softmax, the mask and the rank-`nbasis`
// relative term computed arithmetically (never materialised).  head_dim is 128 at every VPT width.
//
//   CTA = (64-query block, head, batch row), 4 warps x 16 queries.  For query i (chunk-local) the visible keys are
//   j in (i, i + maxlen] in [memory|chunk] coordinates (d = maxlen + i - j in [0, maxlen)), so a 64-query block
//   touches at most maxlen + 63 keys.
#pragma once
#include "common.cuh"

namespace vpt {

constexpr int kAttD = 128;             // head dim
constexpr int kAttBQ = 64;             // queries per CTA
constexpr int kAttBK = 64;             // keys per block
constexpr int kAttPitch = kAttD + 8;   // bf16 elements per smem row (272 B: conflict-free ldmatrix)
constexpr int kAttThreads = 128;

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// rows [row0, row0+64) of a [rows_total][ld] bf16 matrix (128 columns starting at col0) -> smem tile, zero beyond rows_total
__device__ __forceinline__ void load_tile_64x128(__nv_bfloat16* dst, const __nv_bfloat16* src, long long ld, int row0, int rows_total,
                                                 int col0) {
    for (int i = threadIdx.x; i < 64 * 16; i += kAttThreads) {
        const int r = i >> 4, ch = i & 15;
        __nv_bfloat16* d = dst + r * kAttPitch + ch * 8;
        const int row = row0 + r;
        if (row >= 0 && row < rows_total) cp_async16(d, src + (long long)row * ld + col0 + ch * 8);
        else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    }
}

__global__ void __launch_bounds__(kAttThreads) attention_kernel(
    const __nv_bfloat16* __restrict__ Q, const __nv_bfloat16* __restrict__ Kf, const __nv_bfloat16* __restrict__ Vf,
    const float* __restrict__ R, long long ld_r, const float* __restrict__ b_nd, const uint8_t* __restrict__ first,
    long long first_stride, const uint8_t* __restrict__ smask, __nv_bfloat16* __restrict__ out, int t, int maxlen, int heads,
    int nbasis, int causal) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t att_smem[];
    __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(att_smem);
    __nv_bfloat16* Ks = Qs + kAttBQ * kAttPitch;
    __nv_bfloat16* Vs = Ks + kAttBK * kAttPitch;
    float* Es = reinterpret_cast<float*>(Vs + kAttBK * kAttPitch);  // [64][maxlen]
    float* Bs = Es + (size_t)kAttBQ * maxlen;                        // [nbasis][maxlen]
    float* Rs = Bs + (size_t)nbasis * maxlen;                        // [64][nbasis]
    uint8_t* Ms = reinterpret_cast<uint8_t*>(Rs + kAttBQ * nbasis);  // [maxlen] memory key usable?

    const int q0 = blockIdx.x * kAttBQ, head = blockIdx.y, b = blockIdx.z;
    const int h = heads * kAttD;
    const int T = maxlen + t;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, tg = lane & 3;
    const __nv_bfloat16* Qb = Q + (long long)b * t * h;
    const __nv_bfloat16* Kb = Kf + (long long)b * T * h;
    const __nv_bfloat16* Vb = Vf + (long long)b * T * h;

    load_tile_64x128(Qs, Qb, h, q0, t, head * kAttD);
    if (causal && maxlen > 0) {
        const bool mem_ok = (first[(long long)b * first_stride] == 0) && (smask != nullptr);
        for (int j = threadIdx.x; j < maxlen; j += kAttThreads) Ms[j] = mem_ok ? smask[(long long)b * maxlen + j] : 0;
        for (int i = threadIdx.x; i < nbasis * maxlen; i += kAttThreads) Bs[i] = __ldg(b_nd + i);
        for (int i = threadIdx.x; i < kAttBQ * nbasis; i += kAttThreads) {
            const int r = i / nbasis, n = i % nbasis;
            Rs[i] = (q0 + r < t) ? __ldg(R + ((long long)b * t + q0 + r) * ld_r + head * nbasis + n) : 0.f;
        }
    }
    cp_async_wait_all();
    __syncthreads();
    if (causal && maxlen > 0) {
        for (int i = threadIdx.x; i < kAttBQ * maxlen; i += kAttThreads) {
            const int r = i / maxlen, d = i % maxlen;
            float e = 0.f;
            for (int n = 0; n < nbasis; ++n) e = fmaf(Rs[r * nbasis + n], Bs[n * maxlen + d], e);
            Es[i] = e;
        }
    }

    // Q fragments for this warp's 16 rows, all 8 k-steps
    uint32_t qf[8][4];
    {
        const int row = warp * 16 + (lane & 15);
        const int colh = (lane >> 4) * 8;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
            ldsm_x4(smem_u32(Qs + row * kAttPitch + ks * 16 + colh), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }

    float o[16][4];
#pragma unroll
    for (int n = 0; n < 16; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    const float kLog2e = 1.4426950408889634f;
    const float qk_scale = 1.0f / (float)kAttD;  // muP 1/d (lib/xf.py:59)

    const int last_q = min(q0 + kAttBQ, t) - 1;
    int j_lo, j_hi;  // inclusive key range in [memory|chunk] coordinates
    if (causal) {
        j_lo = q0 + 1;
        j_hi = min(last_q + maxlen, T - 1);
    } else {
        j_lo = 0;
        j_hi = T - 1;
    }
    const int iq[2] = {q0 + warp * 16 + g, q0 + warp * 16 + g + 8};

    for (int kb0 = j_lo; kb0 <= j_hi; kb0 += kAttBK) {
        __syncthreads();  // previous block's K/V fully consumed (also orders the Es writes before first use)
        load_tile_64x128(Ks, Kb, h, kb0, T, head * kAttD);
        load_tile_64x128(Vs, Vb, h, kb0, T, head * kAttD);
        cp_async_wait_all();
        __syncthreads();

        // S = Q K^T  (16 x 64 per warp)
        float s[8][4];
#pragma unroll
        for (int n = 0; n < 8; ++n) s[n][0] = s[n][1] = s[n][2] = s[n][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
            for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
                uint32_t b0, b1, b2, b3;
                const int krow = np * 16 + (lane & 7) + ((lane >> 4) << 3);
                const int kcol = ks * 16 + (((lane >> 3) & 1) << 3);
                ldsm_x4(smem_u32(Ks + krow * kAttPitch + kcol), b0, b1, b2, b3);
                mma_bf16_16816(s[2 * np], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b0, b1);
                mma_bf16_16816(s[2 * np + 1], qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3], b2, b3);
            }
        }
        // logits (log2 domain), mask, relative bias
        float bmax[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int n = 0; n < 8; ++n) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int rr = e >> 1;
                const int i = iq[rr];
                const int j = kb0 + n * 8 + 2 * tg + (e & 1);
                bool ok = (i < t) && (j < T);
                float extra = 0.f;
                if (causal) {
                    const int d = maxlen + i - j;
                    ok = ok && (d >= 0) && (d < maxlen);
                    if (ok) {
                        ok = (j >= maxlen) || (Ms[j] != 0);
                        extra = Es[(warp * 16 + g + rr * 8) * maxlen + d];
                    }
                }
                const float v = ok ? (s[n][e] * qk_scale + extra) * kLog2e : -INFINITY;
                s[n][e] = v;
                bmax[rr] = fmaxf(bmax[rr], v);
            }
        }
        float scale[2];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            float bm = bmax[rr];
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 1));
            bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 2));
            const float m_new = fmaxf(mrow[rr], bm);
            const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
            scale[rr] = exp2f(mrow[rr] - m_use);  // exp2(-inf) = 0 on the first block
            mrow[rr] = m_new;
            float rs = 0.f;
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const float p0 = exp2f(s[n][2 * rr] - m_use), p1 = exp2f(s[n][2 * rr + 1] - m_use);
                s[n][2 * rr] = p0;
                s[n][2 * rr + 1] = p1;
                rs += p0 + p1;
            }
            rs += __shfl_xor_sync(0xffffffffu, rs, 1);
            rs += __shfl_xor_sync(0xffffffffu, rs, 2);
            lrow[rr] = lrow[rr] * scale[rr] + rs;
        }
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            o[n][0] *= scale[0]; o[n][1] *= scale[0];
            o[n][2] *= scale[1]; o[n][3] *= scale[1];
        }
        // O += P V   (P: 16 x 64 as A fragments; V: [key][d] read with ldmatrix.trans)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a0 = pack_bf16(s[2 * ks][0], s[2 * ks][1]);
            const uint32_t a1 = pack_bf16(s[2 * ks][2], s[2 * ks][3]);
            const uint32_t a2 = pack_bf16(s[2 * ks + 1][0], s[2 * ks + 1][1]);
            const uint32_t a3 = pack_bf16(s[2 * ks + 1][2], s[2 * ks + 1][3]);
#pragma unroll
            for (int np = 0; np < 8; ++np) {  // pairs of 8-wide d tiles
                uint32_t b0, b1, b2, b3;
                const int vrow = ks * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
                const int vcol = np * 16 + ((lane >> 4) << 3);
                ldsm_x4_t(smem_u32(Vs + vrow * kAttPitch + vcol), b0, b1, b2, b3);
                mma_bf16_16816(o[2 * np], a0, a1, a2, a3, b0, b1);
                mma_bf16_16816(o[2 * np + 1], a0, a1, a2, a3, b2, b3);
            }
        }
    }

    // normalise and store
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        const int i = iq[rr];
        if (i >= t) continue;
        const float inv = 1.0f / lrow[rr];
        __nv_bfloat16* op = out + ((long long)b * t + i) * h + head * kAttD;
#pragma unroll
        for (int n = 0; n < 16; ++n)
            *reinterpret_cast<uint32_t*>(op + n * 8 + 2 * tg) = pack_bf16(o[n][2 * rr] * inv, o[n][2 * rr + 1] * inv);
    }
}

}  // namespace vpt

extern "C" int vpt_attention(const void* Q, const void* Kf, const void* Vf, const float* R, int64_t ld_r, const float* b_nd,
                             const uint8_t* first, int64_t first_stride, const uint8_t* smask, void* out, int32_t B, int32_t t,
                             int32_t maxlen, int32_t heads, int32_t nbasis, int32_t causal, void* stream) {
    using namespace vpt;
    VPT_CHECK(Q && Kf && Vf && out && B > 0 && t > 0 && heads > 0 && maxlen >= 0, "vpt_attention: bad arguments");
    if (causal) VPT_CHECK(maxlen > 0 && R && b_nd && first && nbasis > 0, "vpt_attention: causal mode needs maxlen > 0, R, b_nd, first");
    else VPT_CHECK(maxlen == 0, "vpt_attention: mask 'none' has no KV memory (maxlen must be 0)");
    VPT_CHECK(B <= 65535 && heads <= 65535, "vpt_attention: grid too large");
    const int nb = causal ? nbasis : 0;
    size_t smem = (size_t)(kAttBQ + 2 * kAttBK) * kAttPitch * 2 + ((size_t)kAttBQ * maxlen + (size_t)nb * maxlen + (size_t)kAttBQ * nb) * 4 +
                  (size_t)((maxlen + 15) / 16 * 16) + 16;
    VPT_CHECK(smem <= 227 * 1024, "vpt_attention: maxlen=%d too large for the shared-memory budget", maxlen);
    static size_t attr = 0;
    if (smem > attr) {
        VPT_CUDA(cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = smem;
    }
    dim3 grid((t + kAttBQ - 1) / kAttBQ, heads, B);
    launch_k(attention_kernel, dim3(grid), dim3(kAttThreads), smem, (cudaStream_t)stream, 
        reinterpret_cast<const __nv_bfloat16*>(Q), reinterpret_cast<const __nv_bfloat16*>(Kf), reinterpret_cast<const __nv_bfloat16*>(Vf), R,
        ld_r, b_nd, first, first_stride, smask, reinterpret_cast<__nv_bfloat16*>(out), t, maxlen, heads, nb, causal);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
