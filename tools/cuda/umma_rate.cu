// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 -> fp32, K = 16 per instruction) for the operand arrangements the
// attention and GEMM kernels use, one CTA per SM, operands resident (no loads in the timed region):
//   SS  M=128 N in {64,128,256}  : A and B from shared memory (SWIZZLE_128B K-major tiles)
//   TS  M=128 N=128              : A from tensor memory, B from shared memory (the P V product of attention)
// Each case issues `reps` groups of 8 MMAs (K = 128) back to back from one elected thread, commits once and waits.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../mmada_parallel_b200/csrc umma_rate.cu -o umma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"

using namespace mmdp;

template <int N, bool TS, int DISTINCT>
__global__ void __launch_bounds__(128, 1) rate_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    // A: DISTINCT tiles of 128 x 64 bf16 (16 KB each); B: DISTINCT tiles of N x 64 (N * 128 B each)
    uint8_t* sA = smem;
    uint8_t* sB = smem + DISTINCT * 16384;
    for (int i = threadIdx.x; i < (DISTINCT * (16384 + N * 128)) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_slot);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, N);
            t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const int slot = r % DISTINCT;
                const uint32_t aA = smem_u32(sA + slot * 16384), aB = smem_u32(sB + slot * N * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (TS)
                        umma_bf16_ts(tm, tm + 256 + k * 8, umma_desc_kmajor_sw128(aB) + k * 2, idesc, 1);
                    else
                        umma_bf16_ss(tm, umma_desc_kmajor_sw128(aA) + k * 2, umma_desc_kmajor_sw128(aB) + k * 2, idesc, 1);
                }
            }
            umma_commit(&bar);
            mbar_wait(&bar, 0);
            t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tm);
    }
}

template <int N, bool TS, int DISTINCT>
static void run(const char* name) {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = DISTINCT * (16384 + N * 128);
    cudaFuncSetAttribute(rate_kernel<N, TS, DISTINCT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2000;
    for (int it = 0; it < 2; ++it) rate_kernel<N, TS, DISTINCT><<<148, 128, smem>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (reps * 4.0);
    printf("%-28s N=%3d distinct tiles=%d: %7.1f cycles / MMA (floor %d)  operand bytes / MMA: A %d + B %d -> %.1f B/clk  [%s]\n", name, N, DISTINCT, per,
           128 * N / 256, TS ? 0 : 4096, N * 32, ((TS ? 0 : 4096) + N * 32) / per, cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<64, false, 1>("SS M=128");
    run<128, false, 1>("SS M=128");
    run<256, false, 1>("SS M=128");
    run<64, false, 4>("SS M=128");
    run<128, false, 4>("SS M=128");
    run<256, false, 4>("SS M=128");
    run<128, true, 1>("TS M=128 (A from TMEM)");
    run<128, true, 4>("TS M=128 (A from TMEM)");
    run<256, true, 2>("TS M=128 (A from TMEM)");
    return 0;
}
