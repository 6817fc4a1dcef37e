// Attention forward v4 (default): same math / layouts / interface as attention.cu (v3), restructured for occupancy.
//
// ncu on v3 (profiles/r01): SFU 25 %, tensor pipe 32 %, issue slots ~16 % busy - the kernel is latency-bound, it needs
// more independent work per SM. v3 keeps the output accumulator in registers (64-128 per thread), which pins it to one
// CTA per SM. v4 accumulates O in TENSOR MEMORY across KV blocks (tcgen05.mma with the accumulate flag) and rescales it
// lazily, so a softmax thread needs < 100 registers and TWO CTAs fit on an SM (96 KB smem, 256 TMEM columns each):
//   * KV blocks of 64; S = Q K^T (128 x 64) double-buffered in TMEM, O (128 x 128) resident in TMEM;
//   * 4 softmax warps (thread = query row): S -> registers, running max m_used is only raised when the row max exceeds
//     it by more than 2^8 (FlashAttention-4's lazy rescale): then O (TMEM) and l are multiplied by 2^(m_old - m_new)
//     through tcgen05.ld / tcgen05.st before P is published; otherwise P = 2^((s - m_used) * c) <= 256 is used as is;
//   * P (bf16) double-buffered in swizzled smem, single-stage K and V^T tiles (the co-resident CTA fills the bubbles);
//   * every 4th exp2 runs on the FMA pipe (see attention.cu).
// The final O / l is identical in exact arithmetic; in bf16 the P values carry the same relative precision.
#include "mmdp_internal.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace mmdp {

static constexpr int k4Threads = 192;
static constexpr int k4BKV = 64;
static constexpr int k4QBytes = 128 * 128 * 2;       // 32 KB (two 64-column halves)
static constexpr int k4KBytes = k4BKV * 128 * 2;     // 16 KB (two halves of 64 rows x 64 cols)
static constexpr int k4VBytes = 128 * k4BKV * 2;     // 16 KB (128 d rows x 64 kv)
static constexpr int k4PBytes = 128 * k4BKV * 2;     // 16 KB
// smem: Q | K[2] | V[2] | P | barriers = 112.25 KB; no alignment slack (the dynamic smem window of a kernel without static
// smem starts 1024-aligned; checked at run time) so that two CTAs fit into the 227 KB of an SM
static constexpr int k4Smem = k4QBytes + 2 * k4KBytes + 2 * k4VBytes + k4PBytes + 256;

__device__ __forceinline__ float ex2_approx4(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_fma4(float x) {  // see attention.cu::ex2_fma
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(0.0551716573536396f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));
}

__global__ void __launch_bounds__(k4Threads, 2)
attention_v4_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ out, int H, int L, int d_model,
                    float scale_log2) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
        printf("mmdp: attention smem base not 1024-byte aligned\n");
        __trap();
    }
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + k4QBytes;      // 2 stages
    uint8_t* sV = sK + 2 * k4KBytes;  // 2 stages
    uint8_t* sP = sV + 2 * k4VBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + k4PBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2]
    uint64_t* s_empty = bars + 11;  // [2]
    uint64_t* p_full = bars + 13;
    uint64_t* pv_done = bars + 14;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = (L + k4BKV - 1) / k4BKV;

    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        fence_barrier_init();
    }
    if (warp == 4) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmVt);
        }
        __syncwarp();
        tmem_alloc<256>(tmem_ptr);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tS0 = tmem_base, tO = tmem_base + 128;  // S[0] cols 0..63, S[1] cols 64..127, O cols 128..255

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int qrow0 = b * L + qt * 128;
            mbar_expect_tx(q_full, k4QBytes);
            tma_load_2d(sQ, &tmQ, q_full, h * 128, qrow0);
            tma_load_2d(sQ + k4QBytes / 2, &tmQ, q_full, h * 128 + 64, qrow0);
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                const uint32_t u = (j >> 1) & 1;
                const int kv0 = j * k4BKV;
                mbar_wait(&k_empty[st], u ^ 1);
                mbar_expect_tx(&k_full[st], k4KBytes);
                tma_load_2d(sK + st * k4KBytes, &tmK, &k_full[st], h * 128, b * L + kv0);
                tma_load_2d(sK + st * k4KBytes + k4KBytes / 2, &tmK, &k_full[st], h * 128 + 64, b * L + kv0);
                mbar_wait(&v_empty[st], u ^ 1);
                mbar_expect_tx(&v_full[st], k4VBytes);
                tma_load_2d(sV + st * k4VBytes, &tmVt, &v_full[st], kv0, (b * H + h) * 128);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, k4BKV);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128);
            const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
            mbar_wait(q_full, 0);
            for (int j = 0; j <= n_kv; ++j) {
                if (j < n_kv) {
                    const int s = j & 1;
                    const uint32_t u = (j >> 1) & 1;
                    const uint32_t aK = smem_u32(sK + s * k4KBytes);
                    mbar_wait(&k_full[s], u);
                    mbar_wait(&s_empty[s], u ^ 1);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {  // K dimension = head_dim 128: two 64-column halves of Q and K
                        const uint32_t qoff = (k >> 2) * (k4QBytes / 2), koff = (k >> 2) * (k4KBytes / 2);
                        umma_bf16_ss(tS0 + s * k4BKV, umma_desc_kmajor_sw128(aQ + qoff) + (k & 3) * 2,
                                     umma_desc_kmajor_sw128(aK + koff) + (k & 3) * 2, idesc_qk, k != 0);
                    }
                    umma_commit(&k_empty[s]);
                    umma_commit(&s_full[s]);
                }
                if (j >= 1) {
                    const int jj = j - 1, s = jj & 1;
                    const uint32_t aV = smem_u32(sV + s * k4VBytes);
                    mbar_wait(&v_full[s], (jj >> 1) & 1);
                    mbar_wait(p_full, jj & 1);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < k4BKV / 16; ++k)  // K dimension = 64 kv of this block
                        umma_bf16_ss(tO, umma_desc_kmajor_sw128(aP) + k * 2, umma_desc_kmajor_sw128(aV) + k * 2, idesc_pv, (jj | k) != 0);
                    umma_commit(&v_empty[s]);
                    umma_commit(pv_done);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax (warps 0..3, thread = query row) =====================
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        float m_used = -INFINITY, l_run = 0.f;
        constexpr float kLazy = 8.0f;  // raise the running max only when it is exceeded by more than 2^8

        for (int j = 0; j < n_kv; ++j) {
            const int s = j & 1;
            const int nvalid = L - j * k4BKV;
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tcgen05_fence_after();
            uint32_t sv[64];
            tmem_ld_32x32b_x32(tS0 + s * k4BKV + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32b_x32(tS0 + s * k4BKV + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);  // S is in registers
            if (nvalid < 64) {
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= nvalid) sv[i] = 0xff800000u;  // -inf
            }
            float m8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) m8[i] = __uint_as_float(sv[i]);
#pragma unroll
            for (int i = 8; i < 64; ++i) m8[i & 7] = fmaxf(m8[i & 7], __uint_as_float(sv[i]));
            const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));

            // lazy rescale decision: only when this row's max exceeds the max in use by more than 2^kLazy (always on block 0)
            const bool need = (mx - m_used) * scale_log2 > kLazy;  // m_used = -inf on block 0 -> true
            const bool any_need = __any_sync(0xffffffffu, need);
            const float m_new = need ? mx : m_used;
            const float alpha = need ? ex2_approx4((m_used - m_new) * scale_log2) : 1.0f;  // 0 on the first block
            m_used = m_new;
            const float mneg = -m_used * scale_log2;

            // P = 2^((s - m_used) * c) in registers (bf16 pairs); row sum in fp32. Overlaps PV(j-1) on the tensor core.
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
            uint32_t pk[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float p0 = ex2_approx4(fmaf(__uint_as_float(sv[2 * i]), scale_log2, mneg));
                const float a1 = fmaf(__uint_as_float(sv[2 * i + 1]), scale_log2, mneg);
                const float p1 = (i & 1) ? ex2_fma4(a1) : ex2_approx4(a1);
                rs4[i & 3] += p0 + p1;
                pk[i] = pack_bf16x2(p0, p1);
            }
            l_run = fmaf(l_run, alpha, (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]));

            // PV(j-1) must have retired before O is rescaled and before the (single) P buffer is overwritten
            if (j >= 1) {
                mbar_wait(pv_done, (j - 1) & 1);
                tcgen05_fence_after();
                if (any_need) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x32(tO + lane_off + c * 32, v);
                    }
                    tmem_st_wait();
                    tcgen05_fence_before();
                }
            }
            uint8_t* prow = sP + r * 128;
#pragma unroll
            for (int lc = 0; lc < 8; ++lc)
                *reinterpret_cast<uint4*>(prow + ((lc ^ (r & 7)) << 4)) = make_uint4(pk[4 * lc], pk[4 * lc + 1], pk[4 * lc + 2], pk[4 * lc + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l
        mbar_wait(pv_done, (n_kv - 1) & 1);
        tcgen05_fence_after();
        const int qrow = qt * 128 + r;
        const float inv_l = 1.0f / l_run;
        __nv_bfloat16* orow = out + (size_t)(b * L + (qrow < L ? qrow : 0)) * d_model + h * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
            tmem_ld_wait();
            if (qrow < L) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
                uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) {
        tcgen05_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

int attention_fwd_v3(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream);

static int g_attn_version = -1;
void set_attention_version(int v) { g_attn_version = v; }

int attention_fwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                  int Lpad, float scale, cudaStream_t stream) {
    if (g_attn_version < 0) {
        const char* e = getenv("MMDP_ATTN");
        g_attn_version = (e && e[0] == '3') ? 3 : 4;
    }
    if (g_attn_version == 3) return attention_fwd_v3(q, k, vt, out, B, H, L, Lpad, scale, stream);
    if (B <= 0 || H <= 0 || L <= 0) return set_error("attention: empty problem");
    if (Lpad < L || (Lpad % 8)) return set_error("attention: Lpad must be >= L and a multiple of 8");
    const int d_model = H * 128;
    CUtensorMap tmQ, tmK, tmVt;
    if (make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmK, k, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, k4BKV, 64)) return -1;
    if (make_tmap_2d_bf16(&tmVt, vt, (uint64_t)B * H * 128, (uint64_t)Lpad, (uint64_t)Lpad, 128, 64)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        MMDP_CUDA(cudaFuncSetAttribute(attention_v4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k4Smem));
        MMDP_CUDA(cudaFuncSetAttribute(attention_v4_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    dim3 grid((L + 127) / 128, H, B);
    const float scale_log2 = scale * 1.4426950408889634f;
    LaunchScope ls(LK_ATTN, 4.0 * B * H * (double)L * L * 128, stream);
    attention_v4_kernel<<<grid, k4Threads, k4Smem, stream>>>(tmQ, tmK, tmVt, out, H, L, d_model, scale_log2);
    MMDP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mmdp
