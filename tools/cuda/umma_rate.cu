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

// cta_group::2: a 2-CTA cluster issues M = MM (256: 128 rows per CTA, or 128: 64 rows per CTA) x N MMAs from the leader CTA; each CTA holds
// its rows of A and half of the B tile. Same timing scheme (leader CTA's clock).
__device__ __forceinline__ uint32_t cl_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }
template <int MM, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate2_kernel(int reps, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    uint8_t* sA = smem;            // 128 x 64 bf16 (only MM / 2 rows are read)
    uint8_t* sB = smem + 16384;    // N / 2 rows x 64
    for (int i = threadIdx.x; i < (16384 + N * 64) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    const bool leader = cl_rank() == 0;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async_smem();
    cl_sync();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cl_sync();
    tcgen05_fence_after();
    const uint32_t tm = tmem_slot;
    if (threadIdx.x < 32 && leader) {
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(MM, N);
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ad = umma_desc_kmajor_sw128(smem_u32(sA)) + k * 2, bd = umma_desc_kmajor_sw128(smem_u32(sB)) + k * 2;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tm), "l"(ad), "l"(bd), "r"(idesc), "r"(1) : "memory");
                }
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(&bar)), "h"((uint16_t)1) : "memory");
            mbar_wait(&bar, 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        }
        __syncwarp();
    }
    tcgen05_fence_before();
    cl_sync();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
    }
}
template <int MM, int N>
static void run2(const char* name) {
    long long* d;
    cudaMalloc(&d, 8);
    const int smem = 16384 + N * 64;
    cudaFuncSetAttribute(rate2_kernel<MM, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int reps = 2000;
    for (int it = 0; it < 2; ++it) rate2_kernel<MM, N><<<148, 128, smem>>>(reps, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)cyc / (reps * 4.0);
    printf("%-28s M=%3d N=%3d: %7.1f cycles / MMA (per-SM work %d x %d x 16: floor %d)  [%s]\n", name, MM, N, per, MM / 2, N, (MM / 2) * N / 256, cudaGetErrorString(e));
    cudaFree(d);
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
    run2<256, 256>("SS cta_group::2");
    run2<256, 128>("SS cta_group::2");
    run2<128, 256>("SS cta_group::2");
    run2<128, 128>("SS cta_group::2");
    return 0;
}
