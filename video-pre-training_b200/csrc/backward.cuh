// HBM-bound kernels of the BC backward pass (video-pre-training_b200/training.py): ReLU masking, the residual add of the
// training forward, the three passes of a GroupNorm / LayerNorm backward (per-group sums, per-channel sums, apply),
// max-pool backward and the softmax-NLL gradient.
// Same conventions as elementwise.cuh: 16-byte vectors, deterministic reductions (partials + fixed-order finalisation).
#pragma once
#include "common.cuh"
#include "elementwise.cuh"

namespace vpt {

__device__ __forceinline__ void unpack8(const uint4& v, float* x) {
    x[0] = bf16_lo(v.x); x[1] = bf16_hi(v.x); x[2] = bf16_lo(v.y); x[3] = bf16_hi(v.y);
    x[4] = bf16_lo(v.z); x[5] = bf16_hi(v.z); x[6] = bf16_lo(v.w); x[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float* x) {
    uint4 o;
    o.x = pack_bf16(x[0], x[1]); o.y = pack_bf16(x[2], x[3]); o.z = pack_bf16(x[4], x[5]); o.w = pack_bf16(x[6], x[7]);
    return o;
}
__device__ __forceinline__ void load8f(const float* p, float* x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
// 0xffff per bf16 lane of `w` that is > 0
__device__ __forceinline__ uint32_t pos_mask2(uint32_t w) {
    const uint32_t lo = w & 0xffffu, hi = w >> 16;
    return ((lo != 0u && lo < 0x8000u) ? 0xffffu : 0u) | ((hi != 0u && hi < 0x8000u) ? 0xffff0000u : 0u);
}

// ---------------------------------------------------------------------------------------------------------
// dz = dout where out > 0
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) relu_mask_kernel(const uint4* __restrict__ dout, const uint4* __restrict__ out, uint4* __restrict__ dz,
                                                          long long n8) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (long long)gridDim.x * blockDim.x) {
        const uint4 g = __ldg(dout + i), o = __ldg(out + i);
        dz[i] = make_uint4(g.x & pos_mask2(o.x), g.y & pos_mask2(o.y), g.z & pos_mask2(o.z), g.w & pos_mask2(o.w));
    }
}

// ---------------------------------------------------------------------------------------------------------
// out = a + b with per-group (sum, sumsq) partials of the stored values; grid = (P, groups)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) add_stats_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out,
                                                          float2* __restrict__ stat_part, long long items) {
    const long long g = blockIdx.y;
    const uint4* ga = a + g * items;
    const uint4* gb = b + g * items;
    uint4* go = out + g * items;
    float s = 0.f, ss = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        float x[8], y[8];
        unpack8(__ldg(ga + i), x);
        unpack8(__ldg(gb + i), y);
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] += y[j];
        const uint4 o = pack8(x);
        go[i] = o;
        unpack8(o, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s += x[j];
            ss = fmaf(x[j], x[j], ss);
        }
    }
    const float2 r = block_sum2(s, ss);
    if (threadIdx.x == 0) stat_part[g * gridDim.x + blockIdx.x] = r;
}

// ---------------------------------------------------------------------------------------------------------
// norm backward 1/3: per statistics group  (sum gamma*du, sum gamma*du*n),  n = (x - mean) * rstd;  grid = (P, G)
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_sums_kernel(const uint4* __restrict__ du, const uint4* __restrict__ x, const float2* __restrict__ mr,
                                                           const float* __restrict__ gamma, float2* __restrict__ part, long long items, int C8) {
    const long long g = blockIdx.y;
    const float2 st = __ldg(mr + g);
    const uint4* gd = du + g * items;
    const uint4* gx = x + g * items;
    float s = 0.f, ss = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C8) * 8;
        float d[8], xv[8], ga[8];
        unpack8(__ldg(gd + i), d);
        unpack8(__ldg(gx + i), xv);
        load8f(gamma + c, ga);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float dn = ga[j] * d[j];
            s += dn;
            ss = fmaf(dn, (xv[j] - st.x) * st.y, ss);
        }
    }
    const float2 r = block_sum2(s, ss);
    if (threadIdx.x == 0) part[g * gridDim.x + blockIdx.x] = r;
}

__global__ void sums_finalize_kernel(const float2* __restrict__ part, float2* __restrict__ ms, long long G, int P, double inv_count) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    double a = 0.0, b = 0.0;
    for (int i = 0; i < P; ++i) {
        const float2 v = __ldg(part + g * P + i);
        a += (double)v.x;
        b += (double)v.y;
    }
    ms[g] = make_float2((float)(a * inv_count), (float)(b * inv_count));
}

// ---------------------------------------------------------------------------------------------------------
// norm backward 2/3: per channel  (sum_rows du*n, sum_rows du);  block = 32 channel vectors x 8 row lanes,
// grid = (ceil(C8/32), S row slabs); partials [S][2][C] summed in a fixed order by col_sums_finalize_kernel
// ---------------------------------------------------------------------------------------------------------
// kRowStats: every row may belong to another statistics group (LayerNorm rows, or slabs that straddle groups)
template <bool kRowStats>
__global__ void __launch_bounds__(256, 3) col_sums_kernel(const __nv_bfloat16* __restrict__ du, long long ld_du, const __nv_bfloat16* __restrict__ x,
                                                         const float2* __restrict__ mr, const float* __restrict__ gamma, float* __restrict__ ws,
                                                         float2* __restrict__ gpart, long long rows, int C, int rows_per_group, int slabs_per_group,
                                                         long long rows_per_slab) {
    __shared__ float red[8][32][17];
    const int vl = threadIdx.x & 31, rl = threadIdx.x >> 5;
    const int cv = blockIdx.x * 32 + vl;
    const bool active = cv * 8 < C;
    // slabs_per_group > 0: every slab lies inside one statistics group (its (mean, rstd) is loaded once, and the block can also
    // emit the group's  sum gamma*du  /  sum gamma*du*n  partials -- the "group sums" of the norm backward -- from its column sums)
    long long r_begin, r_end;
    float2 st = make_float2(0.f, 1.f);
    if (slabs_per_group > 0) {
        const long long g = blockIdx.y / slabs_per_group;
        const int k = blockIdx.y - (int)g * slabs_per_group;
        r_begin = g * rows_per_group + k * rows_per_slab;
        r_end = min((g + 1) * rows_per_group, r_begin + rows_per_slab);
        if (x != nullptr) st = __ldg(mr + g);
    } else {
        r_begin = (long long)blockIdx.y * rows_per_slab;
        r_end = min(rows, r_begin + rows_per_slab);
    }
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.f;
    if (active) {
        // 4 rows per trip with all 8 loads issued before the first use (ncu: the one-row loop ran at half of HBM bandwidth, latency bound)
        for (long long r = r_begin + rl; r < r_end; r += 32) {
            uint4 dv[4], xr[4];
            float2 sv[kRowStats ? 4 : 1];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long rr = r + 8 * u;
                dv[u] = xr[u] = make_uint4(0, 0, 0, 0);  // bf16 zeros: a row past the slab adds nothing
                if (kRowStats) sv[u] = st;
                if (rr < r_end) {
                    dv[u] = __ldg(reinterpret_cast<const uint4*>(du + rr * ld_du + cv * 8));
                    if (x != nullptr) {
                        xr[u] = __ldg(reinterpret_cast<const uint4*>(x + rr * (long long)C + cv * 8));
                        if (kRowStats) sv[u] = __ldg(mr + (rows_per_group == 1 ? rr : rr / rows_per_group));
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                float d[8], xv[8];
                unpack8(dv[u], d);
                unpack8(xr[u], xv);
                if (x != nullptr) {
                    const float2 su = kRowStats ? sv[u] : st;
#pragma unroll
                    for (int j = 0; j < 8; ++j) a0[j] = fmaf(d[j], (xv[j] - su.x) * su.y, a0[j]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) a1[j] += d[j];
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[rl][vl][j] = a0[j];
        red[rl][vl][8 + j] = a1[j];
    }
    __syncthreads();
    // 256 threads: (vector lane, value 0..15) pairs of this block -> 512 outputs, two per thread
    float g1 = 0.f, g2 = 0.f;
    for (int o = threadIdx.x; o < 32 * 16; o += 256) {
        const int v = o >> 4, k = o & 15;
        const int c = (blockIdx.x * 32 + v) * 8 + (k & 7);
        if (c >= C) continue;
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) s += red[q][v][k];
        ws[((long long)blockIdx.y * 2 + (k >> 3)) * C + c] = s;
        if (gpart != nullptr) {
            const float gs = __ldg(gamma + c) * s;
            if (k >> 3) g1 += gs;  // sum gamma * du
            else g2 += gs;         // sum gamma * du * n
        }
    }
    if (gpart != nullptr) {
        const float2 r = block_sum2(g1, g2);
        if (threadIdx.x == 0) gpart[(long long)blockIdx.y * gridDim.x + blockIdx.x] = r;
    }
}

// out[i] = sum_s ws[s][i], i over 2*C; block = 32 outputs x 8 slab lanes, fixed-order combination
__global__ void __launch_bounds__(256) col_sums_finalize_kernel(const float* __restrict__ ws, float* __restrict__ out, int n, int S) {
    __shared__ double red[8][32];
    const int l = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int i = blockIdx.x * 32 + l;
    double a = 0.0;
    if (i < n)
        for (int s = sl; s < S; s += 8) a += (double)__ldg(ws + (long long)s * n + i);
    red[sl][l] = a;
    __syncthreads();
    if (sl == 0 && i < n) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][l];
        out[i] = (float)t;
    }
}

// ---------------------------------------------------------------------------------------------------------
// norm backward 3/3: dx = rstd * (gamma*du - m1 - n*m2) [+ add]; with zpC > 0 every group is a ZP frame
// [(H+1)(W+1)][zpC] whose pad row / column is written as zero
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(const uint4* __restrict__ du, const uint4* __restrict__ x, const float2* __restrict__ mr,
                                                               const float* __restrict__ gamma, const float2* __restrict__ ms,
                                                               const uint4* __restrict__ add, uint4* __restrict__ dx, int C8, int items_per_group,
                                                               int zpH, int zpW, int zpC8, int relu_x) {
    // grid = (blocks per group, groups): no 64-bit index arithmetic in the loop
    const long long g = blockIdx.y, base = g * items_per_group;
    const float2 st = __ldg(mr + g), m = __ldg(ms + g);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items_per_group; i += gridDim.x * blockDim.x) {
        if (zpC8 > 0) {
            const int pix = i / zpC8;
            const int py = pix / (zpW + 1), px = pix - py * (zpW + 1);
            if (py >= zpH || px >= zpW) {
                dx[base + i] = make_uint4(0, 0, 0, 0);
                continue;
            }
        }
        const int c = (i % C8) * 8;
        float d[8], xv[8], ga[8];
        unpack8(__ldg(du + base + i), d);
        const uint4 xr = __ldg(x + base + i);
        unpack8(xr, xv);
        load8f(gamma + c, ga);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float n = (xv[j] - st.x) * st.y;
            d[j] = st.y * (ga[j] * d[j] - m.x - n * m.y);
        }
        if (add != nullptr) {
            float a[8];
            unpack8(__ldg(add + base + i), a);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] += a[j];
        }
        uint4 o = pack8(d);
        if (relu_x) {  // x is itself a ReLU output: chain its backward (dx = 0 where x == 0) instead of a separate masking pass
            o.x &= pos_mask2(xr.x); o.y &= pos_mask2(xr.y); o.z &= pos_mask2(xr.z); o.w &= pos_mask2(xr.w);
        }
        dx[base + i] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward of ReLU -> max_pool2d(3, 2, 1) on ZP tensors, two streaming passes:
//   1. per pooled pixel: which of the 9 window positions holds the FIRST maximum (torch tie semantics), one byte per channel
//      (15 = no gradient: the maximum is 0, i.e. every input was <= 0 before the ReLU)
//   2. per input pixel: sum dy over the (at most 4) windows whose arg-max byte names this pixel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool3s2_argmax_kernel(const uint4* __restrict__ x, uint2* __restrict__ idx, int H, int W, int C8) {
    const int Ho = H >> 1, Wo = W >> 1, ip = W + 1;
    const long long f = blockIdx.y;
    const int items = Ho * Wo * C8;
    const uint4* fx = x + f * (long long)(H + 1) * ip * C8;
    uint2* fi = idx + f * (long long)items;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
        const int c = i % C8, ox = (i / C8) % Wo, oy = i / (C8 * Wo);
        float best[8];
        uint32_t arg[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            best[j] = 0.f;  // inputs are >= 0: a window that never exceeds 0 passes no gradient
            arg[j] = 15u;
        }
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
            const int y = 2 * oy - 1 + dy;
            if (y < 0 || y >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const int xx = 2 * ox - 1 + dx;
                if (xx < 0 || xx >= W) continue;
                float v[8];
                unpack8(__ldg(fx + ((long long)y * ip + xx) * C8 + c), v);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (v[j] > best[j]) {
                        best[j] = v[j];
                        arg[j] = (uint32_t)(dy * 3 + dx);
                    }
            }
        }
        uint2 o;
        o.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
        o.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
        fi[i] = o;
    }
}

__global__ void __launch_bounds__(256) maxpool3s2_bwd_kernel(const uint4* __restrict__ dy, const uint2* __restrict__ idx, uint4* __restrict__ dx, int H,
                                                               int W, int C8) {
    const int Ho = H >> 1, Wo = W >> 1;
    const int ip = W + 1, op = Wo + 1;
    const long long f = blockIdx.y;
    const int items = (H + 1) * ip * C8;
    uint4* fdx = dx + f * (long long)items;
    const uint2* fi = idx + f * (long long)Ho * Wo * C8;
    const uint4* fdy = dy + f * (long long)(Ho + 1) * op * C8;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < items; i += gridDim.x * blockDim.x) {
        const int c = i % C8, ix = (i / C8) % ip, iy = i / (C8 * ip);
        if (ix >= W || iy >= H) {
            fdx[i] = make_uint4(0, 0, 0, 0);
            continue;
        }
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        const int oy0 = iy >> 1, oy1 = (iy + 1) >> 1, ox0 = ix >> 1, ox1 = (ix + 1) >> 1;
        for (int oy = oy0; oy <= oy1; ++oy) {
            if (oy >= Ho) continue;
            for (int ox = ox0; ox <= ox1; ++ox) {
                if (ox >= Wo) continue;
                const uint32_t code = (uint32_t)((iy - (2 * oy - 1)) * 3 + (ix - (2 * ox - 1)));  // this pixel's position in the window
                const uint2 a = __ldg(fi + ((long long)oy * Wo + ox) * C8 + c);
                const uint32_t eq_lo = a.x ^ (code * 0x01010101u), eq_hi = a.y ^ (code * 0x01010101u);
                if (((eq_lo - 0x01010101u) & ~eq_lo & 0x80808080u) == 0u && ((eq_hi - 0x01010101u) & ~eq_hi & 0x80808080u) == 0u)
                    continue;  // no zero byte: none of the 8 channels selected this pixel
                float g[8];
                unpack8(__ldg(fdy + ((long long)oy * op + ox) * C8 + c), g);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (((eq_lo >> (8 * j)) & 0xffu) == 0u) acc[j] += g[j];
                    if (((eq_hi >> (8 * j)) & 0xffu) == 0u) acc[4 + j] += g[4 + j];
                }
            }
        }
        fdx[i] = pack8(acc);
    }
}

// ---------------------------------------------------------------------------------------------------------
// out[r][col0 + j] = (exp(logp[r][j]) - [j == idx[r]]) * scale
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const float* __restrict__ logp, const long long* __restrict__ idx, float scale,
                                                            __nv_bfloat16* __restrict__ out, long long ld_out, int col0, long long rows, int n) {
    const long long total = rows * n;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / n;
        const int j = (int)(i - r * n);
        float p = __expf(__ldg(logp + i));
        if ((long long)j == __ldg(idx + r)) p -= 1.f;
        out[r * ld_out + col0 + j] = __float2bfloat16_rn(p * scale);
    }
}

static inline unsigned grid_for(long long items, int per_block = 256, long long cap = 148LL * 16) {
    long long b = (items + per_block - 1) / per_block;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

static inline int parts_for(long long items8) {  // blocks per statistics group
    long long p = (items8 + 1023) / 1024;
    if (p > 64) p = 64;
    if (p < 1) p = 1;
    return (int)p;
}

}  // namespace vpt

extern "C" int vpt_relu_mask(const void* dout, const void* out, void* dz, int64_t n, void* stream) {
    using namespace vpt;
    VPT_CHECK(dout && out && dz && n > 0 && n % 8 == 0, "vpt_relu_mask: need non-null pointers and n %% 8 == 0 (n=%lld)", (long long)n);
    relu_mask_kernel<<<grid_for(n / 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(dout), reinterpret_cast<const uint4*>(out),
                                                                       reinterpret_cast<uint4*>(dz), n / 8);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_add_stat_parts(int64_t elems_per_group) { return vpt::parts_for(elems_per_group / 8); }

extern "C" int vpt_add_stats(const void* a, const void* b, void* out, float* stat_part, int64_t groups, int64_t elems_per_group, void* stream) {
    using namespace vpt;
    VPT_CHECK(a && b && out && stat_part && groups > 0 && groups <= 65535 && elems_per_group > 0 && elems_per_group % 8 == 0,
              "vpt_add_stats: bad arguments (groups=%lld elems=%lld)", (long long)groups, (long long)elems_per_group);
    const long long items = elems_per_group / 8;
    dim3 grid(parts_for(items), (unsigned)groups);
    add_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b),
                                                            reinterpret_cast<uint4*>(out), reinterpret_cast<float2*>(stat_part), items);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_group_sums_parts(int32_t rows_per_group, int32_t C) { return vpt::parts_for((long long)rows_per_group * C / 8); }

extern "C" int vpt_group_sums(const void* du, const void* x, const float* mr, const float* gamma, float* part, float* ms, int64_t rows, int32_t C,
                              int32_t rows_per_group, double count, void* stream) {
    using namespace vpt;
    VPT_CHECK(du && x && mr && gamma && part && ms, "vpt_group_sums: null argument");
    VPT_CHECK(rows > 0 && C > 0 && C % 8 == 0 && rows_per_group > 0 && rows % rows_per_group == 0 && count > 0,
              "vpt_group_sums: bad shape rows=%lld C=%d rows_per_group=%d", (long long)rows, C, rows_per_group);
    const long long G = rows / rows_per_group, items = (long long)rows_per_group * C / 8;
    const int P = parts_for(items);
    for (long long g0 = 0; g0 < G; g0 += 65535) {  // gridDim.y limit
        const long long gn = min((long long)65535, G - g0);
        dim3 grid(P, (unsigned)gn);
        group_sums_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(du) + g0 * items,
                                                                 reinterpret_cast<const uint4*>(x) + g0 * items,
                                                                 reinterpret_cast<const float2*>(mr) + g0, gamma,
                                                                 reinterpret_cast<float2*>(part) + g0 * P, items, C / 8);
        VPT_LAUNCH_CHECK();
    }
    sums_finalize_kernel<<<(unsigned)((G + 255) / 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(part),
                                                                                       reinterpret_cast<float2*>(ms), G, P, 1.0 / count);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

namespace vpt {
static inline int col_sums_slabs(long long rows, int C) {
    const int colblocks = (C / 8 + 31) / 32;
    long long S = (4LL * 148 + colblocks - 1) / colblocks;
    const long long max_s = (rows + 63) / 64;  // at least 64 rows per slab
    if (S > max_s) S = max_s;
    if (S < 1) S = 1;
    if (S > 65535) S = 65535;
    return (int)S;
}
// slabs per statistics group of the fused (column + group sums) pass
static inline int norm_sums_spg(long long groups, int rows_per_group, int C) {
    const int colblocks = (C / 8 + 31) / 32;
    long long spg = (32LL * 148 + colblocks * groups - 1) / (colblocks * groups);  // several waves of 8 resident blocks per SM
    const long long max_spg = (rows_per_group + 63) / 64;
    if (spg > max_spg) spg = max_spg;
    if (spg < 1) spg = 1;
    return (int)spg;
}
}  // namespace vpt

extern "C" int vpt_col_sums_parts(int64_t rows, int32_t C) { return vpt::col_sums_slabs(rows, C); }

extern "C" int vpt_col_sums(const void* du, int64_t ld_du, const void* x, const float* mr, int64_t rows, int32_t C, int32_t rows_per_group,
                            float* out, float* workspace, void* stream) {
    using namespace vpt;
    VPT_CHECK(du && out && workspace && rows > 0 && C > 0 && C % 8 == 0 && ld_du % 8 == 0 && ld_du >= C, "vpt_col_sums: bad arguments");
    VPT_CHECK(x == nullptr || (mr != nullptr && rows_per_group > 0), "vpt_col_sums: x given without statistics");
    const int S = col_sums_slabs(rows, C);
    const long long per = (rows + S - 1) / S;
    dim3 grid((C / 8 + 31) / 32, S);
    col_sums_kernel<true><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(du), ld_du, reinterpret_cast<const __nv_bfloat16*>(x),
                                                           reinterpret_cast<const float2*>(mr), nullptr, workspace, nullptr, rows, C,
                                                           rows_per_group > 0 ? rows_per_group : 1, 0, per);
    VPT_LAUNCH_CHECK();
    col_sums_finalize_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(workspace, out, 2 * C, S);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int64_t vpt_norm_sums_workspace(int64_t rows, int32_t C, int32_t rows_per_group) {
    if (rows <= 0 || C <= 0 || rows_per_group <= 0) return 0;
    const long long G = rows / rows_per_group;
    const long long S = G * vpt::norm_sums_spg(G, rows_per_group, C);
    return S * 2 * C + S * ((C / 8 + 31) / 32) * 2;
}

extern "C" int vpt_norm_sums(const void* du, const void* x, const float* mr, const float* gamma, int64_t rows, int32_t C, int32_t rows_per_group,
                             double count, float* out, float* ms, float* workspace, void* stream) {
    using namespace vpt;
    VPT_CHECK(du && x && mr && gamma && out && ms && workspace, "vpt_norm_sums: null argument");
    VPT_CHECK(rows > 0 && C > 0 && C % 8 == 0 && rows_per_group > 0 && rows % rows_per_group == 0 && count > 0,
              "vpt_norm_sums: bad shape rows=%lld C=%d rows_per_group=%d", (long long)rows, C, rows_per_group);
    const long long G = rows / rows_per_group;
    const int spg = norm_sums_spg(G, rows_per_group, C);
    const long long S = G * spg;
    VPT_CHECK(S <= 65535, "vpt_norm_sums: too many groups (%lld) for one launch", (long long)G);
    const int colblocks = (C / 8 + 31) / 32;
    const long long per = (rows_per_group + spg - 1) / spg;
    float* ws_cols = workspace;
    float2* gpart = reinterpret_cast<float2*>(workspace + S * 2 * C);
    dim3 grid(colblocks, (unsigned)S);
    col_sums_kernel<false><<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(du), C, reinterpret_cast<const __nv_bfloat16*>(x),
                                                           reinterpret_cast<const float2*>(mr), gamma, ws_cols, gpart, rows, C, rows_per_group, spg, per);
    VPT_LAUNCH_CHECK();
    col_sums_finalize_kernel<<<(2 * C + 31) / 32, 256, 0, (cudaStream_t)stream>>>(ws_cols, out, 2 * C, (int)S);
    VPT_LAUNCH_CHECK();
    sums_finalize_kernel<<<(unsigned)((G + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gpart, reinterpret_cast<float2*>(ms), G, spg * colblocks, 1.0 / count);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_norm_bwd_apply(const void* du, const void* x, const float* mr, const float* gamma, const float* ms, const void* add, void* dx,
                                  int64_t rows, int32_t C, int32_t rows_per_group, int32_t zpH, int32_t zpW, int32_t zpC, int32_t relu_x, void* stream) {
    using namespace vpt;
    VPT_CHECK(du && x && mr && gamma && ms && dx, "vpt_norm_bwd_apply: null argument");
    VPT_CHECK(rows > 0 && C > 0 && C % 8 == 0 && rows_per_group > 0 && rows % rows_per_group == 0, "vpt_norm_bwd_apply: bad shape");
    VPT_CHECK(zpC == 0 || (zpC % 8 == 0 && (long long)(zpH + 1) * (zpW + 1) * zpC == (long long)rows_per_group * C),
              "vpt_norm_bwd_apply: ZP geometry (%d,%d,%d) does not match the group size", zpH, zpW, zpC);
    const long long G = rows / rows_per_group, ipg = (long long)rows_per_group * C / 8;
    VPT_CHECK(ipg < 2147483647LL, "vpt_norm_bwd_apply: group too large");
    const int P = parts_for(ipg * 4);  // ~256 items per block-iteration
    for (long long g0 = 0; g0 < G; g0 += 65535) {
        const long long gn = min((long long)65535, G - g0);
        dim3 grid(P, (unsigned)gn);
        norm_bwd_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(du) + g0 * ipg, reinterpret_cast<const uint4*>(x) + g0 * ipg, reinterpret_cast<const float2*>(mr) + g0, gamma,
            reinterpret_cast<const float2*>(ms) + g0, add ? reinterpret_cast<const uint4*>(add) + g0 * ipg : nullptr,
            reinterpret_cast<uint4*>(dx) + g0 * ipg, C / 8, (int)ipg, zpH, zpW, zpC / 8, relu_x);
        VPT_LAUNCH_CHECK();
    }
    return VPT_OK;
}

extern "C" int vpt_maxpool3s2_bwd(const void* dy, const void* x, void* dx, void* workspace, int32_t F, int32_t H, int32_t W, int32_t C, void* stream) {
    using namespace vpt;
    VPT_CHECK(dy && x && dx && workspace && F > 0 && F <= 65535, "vpt_maxpool3s2_bwd: bad arguments");
    VPT_CHECK(H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "vpt_maxpool3s2_bwd: need even H, W and C %% 8 == 0");
    VPT_CHECK(((uintptr_t)workspace & 7) == 0, "vpt_maxpool3s2_bwd: workspace must be 8-byte aligned");
    const long long out_items = (long long)(H / 2) * (W / 2) * (C / 8);
    dim3 g1(grid_for(out_items, 256, 64), F);
    maxpool3s2_argmax_kernel<<<g1, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(x), reinterpret_cast<uint2*>(workspace), H, W, C / 8);
    VPT_LAUNCH_CHECK();
    const long long items = (long long)(H + 1) * (W + 1) * (C / 8);
    dim3 g2(grid_for(items, 256, 64), F);
    maxpool3s2_bwd_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(dy), reinterpret_cast<const uint2*>(workspace),
                                                              reinterpret_cast<uint4*>(dx), H, W, C / 8);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}

extern "C" int vpt_softmax_bwd(const float* logp, const int64_t* idx, float scale, void* out, int64_t ld_out, int32_t col0, int64_t rows, int32_t n,
                               void* stream) {
    using namespace vpt;
    VPT_CHECK(logp && idx && out && rows > 0 && n > 0 && col0 >= 0 && ld_out >= col0 + n, "vpt_softmax_bwd: bad arguments");
    softmax_bwd_kernel<<<grid_for(rows * n), 256, 0, (cudaStream_t)stream>>>(logp, reinterpret_cast<const long long*>(idx), scale,
                                                                            reinterpret_cast<__nv_bfloat16*>(out), ld_out, col0, rows, n);
    VPT_LAUNCH_CHECK();
    return VPT_OK;
}
