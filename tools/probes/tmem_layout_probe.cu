// Probe: which TMEM (lane, column) does each register of tcgen05.ld.16x256b hold, and does stmatrix.trans of the packed bf16 pairs give
// a [column][lane] tile?  Build: nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_layout_probe tmem_layout_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void probe(float* out_ld, uint16_t* out_tile) {
    __shared__ uint32_t tmem_ptr;
    __shared__ __align__(1024) uint16_t tile[16 * 64];  // [16 px rows][64 ch] bf16, plain (no swizzle)
    const int lane = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = tmem_ptr;
    // write: lane L, column c  <-  L * 100 + c   (32x32b.x16: thread i <-> lane i, 16 consecutive columns)
    uint32_t w[16];
    for (int c = 0; c < 16; ++c) w[c] = __float_as_uint((float)(lane * 100 + c));
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(t0),
                 "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]), "r"(w[8]), "r"(w[9]), "r"(w[10]), "r"(w[11]),
                 "r"(w[12]), "r"(w[13]), "r"(w[14]), "r"(w[15])
                 : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncwarp();
    for (int half = 0; half < 2; ++half) {  // lanes [0,16) and [16,32)
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(t0 + ((uint32_t)(half * 16) << 16))
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out_ld[(half * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
        // pack (r0,r1), (r2,r3), (r4,r5), (r6,r7) and store the four 8x8 matrices transposed: tile[px][ch]
        uint32_t pk[4];
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
            pk[i] = *reinterpret_cast<uint32_t*>(&v);
        }
        // address of matrix m = lane / 8, row = lane % 8.  Guess: matrices 0/1 = columns 0..7 with lanes +0 / +8, matrices 2/3 = columns 8..15.
        const int m = lane >> 3, rr = lane & 7;
        const int px = (m >> 1) * 8 + rr, ch = half * 16 + (m & 1) * 8;
        asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(&tile[px * 64 + ch])), "r"(pk[0]), "r"(pk[1]),
                     "r"(pk[2]), "r"(pk[3])
                     : "memory");
    }
    __syncthreads();
    for (int i = lane; i < 16 * 64; i += 32) out_tile[i] = tile[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(t0) : "memory");
}

int main() {
    float* d_ld;
    uint16_t* d_tile;
    cudaMalloc(&d_ld, 64 * 8 * 4);
    cudaMalloc(&d_tile, 16 * 64 * 2);
    probe<<<1, 32>>>(d_ld, d_tile);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    float h[64 * 8];
    uint16_t t[16 * 64];
    cudaMemcpy(h, d_ld, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemcpy(t, d_tile, sizeof(t), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int half = 0; half < 2; ++half)
        for (int T = 0; T < 32; ++T)
            for (int i = 0; i < 8; ++i) {
                // expected (mma accumulator style): regs (0,1) lane T/4, (2,3) lane T/4+8, cols 2(T%4)+{0,1}; regs 4..7 the same for cols +8
                const int ln = half * 16 + T / 4 + ((i >> 1) & 1) * 8, col = (i >> 2) * 8 + 2 * (T % 4) + (i & 1);
                const float want = (float)(ln * 100 + col);
                if (h[(half * 32 + T) * 8 + i] != want) {
                    if (bad < 12) printf("ld mismatch half %d thread %d reg %d: got %.0f want %.0f\n", half, T, i, h[(half * 32 + T) * 8 + i], want);
                    ++bad;
                }
            }
    printf("tcgen05.ld.16x256b.x2 layout %s (%d mismatches)\n", bad ? "DIFFERS from the mma-accumulator guess" : "matches the mma-accumulator guess", bad);
    int bad2 = 0;
    for (int px = 0; px < 16; ++px)
        for (int ch = 0; ch < 32; ++ch) {
            uint32_t bits = (uint32_t)t[px * 64 + ch] << 16;
            float got = *reinterpret_cast<float*>(&bits);
            __nv_bfloat16 wb = __float2bfloat16_rn((float)(ch * 100 + px));
            float want = __bfloat162float(wb);
            if (got != want) {
                if (bad2 < 12) printf("tile mismatch px %d ch %d: got %.0f want %.0f\n", px, ch, got, want);
                ++bad2;
            }
        }
    printf("stmatrix.trans tile [px][ch] %s (%d mismatches)\n", bad2 ? "WRONG" : "correct", bad2);
    if (bad) {
        printf("raw dump (half 0):\n");
        for (int T = 0; T < 32; ++T) {
            printf("T%2d:", T);
            for (int i = 0; i < 8; ++i) printf(" %5.0f", h[T * 8 + i]);
            printf("\n");
        }
    }
    return 0;
}
