// Non-causal, unmasked multi-head attention forward for sm_100a (head_dim = 128), the replacement for
// F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=False) at
// MMaDA-Parallel-A/model/modeling_llada.py:672-679 (M variant :656-663).
//
// One CTA = one (batch, head, 128-query tile). Warp roles (320 threads):
//   warps 0..7 : softmax / output accumulation. TWO threads per query row: warp w works on TMEM lanes 32*(w%4).. and on
//                column half w/4 (64 of the 128 S columns of a KV block, 64 of the 128 output columns), so the
//                exp / FMA instruction stream that bounds this kernel is spread over 8 warps (2 per SM sub-partition).
//   warp 8     : TMA producer (Q once, K and V^T tiles through 2-stage rings) + TMEM allocation
//   warp 9     : MMA issuer: S = Q·K^T (128x128x128) into a double-buffered TMEM accumulator, PV = P·V into a third
// P is written by the softmax threads as bf16 into 128B-swizzled smem (K-major A operand); V is consumed
// from the transposed layout V^T[b][h][d][token] produced by the QKV GEMM epilogue, so both P·V operands are
// K-major like every other MMA in this library. O is accumulated in registers (fp32) with the usual
// online-softmax rescale; sequence tails (L % 128 != 0) are handled by masking S columns >= L to -inf.
#include "mmdp_internal.h"
#include "ptx.cuh"

namespace mmdp {

static constexpr int kAttnThreads = 320;
static constexpr int kHalf = 128 * 64 * 2;  // 16 KB: 128 rows x 64 bf16 (one swizzle atom wide)
static constexpr int kTile = 2 * kHalf;     // 32 KB: 128 x 128 bf16
// smem: Q | K[2] | V[2] | P | barriers (256 B) | row-exchange scratch (3 KB)
static constexpr int kAttnSmem = kTile * 6 + 1024 + 256 + 3072;

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x on the FMA/ALU pipes (x <= 0): the SFU executes only 16 ex2 per clock per SM, which bounds a 128x128 score
// block at 1024 cycles - the same as its two MMAs. Routing every 4th element through this polynomial keeps both pipes
// busy. Round-to-nearest split x = n + f (|f| <= 0.5), degree-3 minimax for 2^f (max rel. error 7.5e-5, far below the
// bf16 rounding of P), exponent inserted with integer adds.
__device__ __forceinline__ float ex2_fma(float x) {
    x = fmaxf(x, -126.0f);
    const float t = x + 12582912.0f;                    // 1.5 * 2^23: integer part of x lands in the low mantissa bits
    const float f = x - (t - 12582912.0f);
    float p = fmaf(0.0551716573536396f, f, 0.2426111251115799f);
    p = fmaf(p, f, 0.6932609677314758f);
    p = fmaf(p, f, 0.9999280571937561f);
    return __uint_as_float(__float_as_uint(p) + (__float_as_uint(t) << 23));  // (t bits - magic) << 23 == n << 23 (mod 2^32)
}
__device__ __forceinline__ void softmax_group_sync() {  // the 256 softmax threads only (named barrier 1)
    asm volatile("bar.sync 1, 256;" ::: "memory");
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attention_v3_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ out, int H, int L, int d_model,
                 float scale_log2) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = smem + kTile;      // 2 stages
    uint8_t* sV = smem + 3 * kTile;  // 2 stages
    uint8_t* sP = smem + 5 * kTile;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTile);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_full = bars + 9;    // [2]
    uint64_t* s_empty = bars + 11;  // [2]
    uint64_t* p_full = bars + 13;
    uint64_t* pv_full = bars + 14;
    uint64_t* pv_empty = bars + 15;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
    float* red_max = reinterpret_cast<float*>(smem + 6 * kTile + 256);  // [2 parity][2 halves][128 rows]
    float* red_sum = red_max + 2 * 2 * 128;                             // [2 halves][128 rows]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = (L + 127) / 128;

    if (warp == 9 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&s_empty[s], 8);
        }
        mbar_init(p_full, 8);
        mbar_init(pv_full, 1);
        mbar_init(pv_empty, 8);
        fence_barrier_init();
    }
    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tmQ);
            tma_prefetch_desc(&tmK);
            tma_prefetch_desc(&tmVt);
        }
        __syncwarp();
        tmem_alloc<512>(tmem_ptr);
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tS0 = tmem_base, tPV = tmem_base + 256;

    if (warp == 8) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            const int qrow0 = b * L + qt * 128;
            mbar_expect_tx(q_full, kTile);
            tma_load_2d(sQ, &tmQ, q_full, h * 128, qrow0);
            tma_load_2d(sQ + kHalf, &tmQ, q_full, h * 128 + 64, qrow0);
            for (int j = 0; j < n_kv; ++j) {
                const int s = j & 1;
                const uint32_t u = (j >> 1) & 1;
                const int kv0 = j * 128;
                mbar_wait(&k_empty[s], u ^ 1);
                mbar_expect_tx(&k_full[s], kTile);
                tma_load_2d(sK + s * kTile, &tmK, &k_full[s], h * 128, b * L + kv0);
                tma_load_2d(sK + s * kTile + kHalf, &tmK, &k_full[s], h * 128 + 64, b * L + kv0);
                mbar_wait(&v_empty[s], u ^ 1);
                mbar_expect_tx(&v_full[s], kTile);
                tma_load_2d(sV + s * kTile, &tmVt, &v_full[s], kv0, (b * H + h) * 128);
                tma_load_2d(sV + s * kTile + kHalf, &tmVt, &v_full[s], kv0 + 64, (b * H + h) * 128);
            }
        }
        __syncwarp();
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            const uint32_t aQ = smem_u32(sQ), aP = smem_u32(sP);
            mbar_wait(q_full, 0);
            for (int j = 0; j <= n_kv; ++j) {
                if (j < n_kv) {
                    const int s = j & 1;
                    const uint32_t u = (j >> 1) & 1;
                    mbar_wait(&k_full[s], u);
                    mbar_wait(&s_empty[s], u ^ 1);
                    tcgen05_fence_after();
                    const uint32_t aK = smem_u32(sK + s * kTile);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t off = (k >> 2) * kHalf;
                        umma_bf16_ss(tS0 + s * 128, umma_desc_kmajor_sw128(aQ + off) + (k & 3) * 2,
                                     umma_desc_kmajor_sw128(aK + off) + (k & 3) * 2, idesc, k != 0);
                    }
                    umma_commit(&k_empty[s]);
                    umma_commit(&s_full[s]);
                }
                if (j >= 1) {
                    const int jj = j - 1;
                    const int s = jj & 1;
                    const uint32_t u = (jj >> 1) & 1;
                    mbar_wait(&v_full[s], u);
                    mbar_wait(p_full, jj & 1);
                    mbar_wait(pv_empty, (jj & 1) ^ 1);
                    tcgen05_fence_after();
                    const uint32_t aV = smem_u32(sV + s * kTile);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t off = (k >> 2) * kHalf;
                        umma_bf16_ss(tPV, umma_desc_kmajor_sw128(aP + off) + (k & 3) * 2,
                                     umma_desc_kmajor_sw128(aV + off) + (k & 3) * 2, idesc, k != 0);
                    }
                    umma_commit(&v_empty[s]);
                    umma_commit(pv_full);
                }
            }
        }
        __syncwarp();
    } else {
        // ===================== softmax + output (warps 0..7) =====================
        const int lq = warp & 3;   // TMEM lane quarter
        const int hc = warp >> 2;  // column half owned by this thread (S columns and O columns [64*hc, 64*hc+64))
        const int r = lq * 32 + lane;  // query row in tile == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(lq * 32) << 16;
        float o_acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;

        for (int j = 0; j < n_kv; ++j) {
            const int s = j & 1;
            const uint32_t u = (j >> 1) & 1;
            const int nvalid = L - j * 128 - hc * 64;  // valid columns inside this thread's half
            mbar_wait(&s_full[s], u);
            tcgen05_fence_after();
            const uint32_t tS = tS0 + s * 128 + hc * 64 + lane_off;

            // pass 1: row max over this thread's 64 columns (both TMEM loads in flight, one wait)
            float mx;
            {
                uint32_t sv[64];
                tmem_ld_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
                tmem_ld_32x32b_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
                tmem_ld_wait();
                if (nvalid < 64) {  // sequence tail: padding columns never win the max
#pragma unroll
                    for (int i = 0; i < 64; ++i)
                        if (i >= nvalid) sv[i] = 0xff800000u;
                }
                // 8 independent max chains (a single chain of 64 dependent FMNMX costs ~250 cycles of latency)
                float m8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) m8[i] = __uint_as_float(sv[i]);
#pragma unroll
                for (int i = 8; i < 64; ++i) m8[i & 7] = fmaxf(m8[i & 7], __uint_as_float(sv[i]));
                mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            }
            float* rm = red_max + (j & 1) * 256;
            rm[hc * 128 + r] = mx;
            softmax_group_sync();
            const float m_new = fmaxf(m_run, fmaxf(mx, rm[(hc ^ 1) * 128 + r]));
            const float mneg = -m_new * scale_log2;

            // consume PV(j-1) (also means the P buffer is free again)
            if (j >= 1) {
                mbar_wait(pv_full, (j - 1) & 1);
                tcgen05_fence_after();
                {
                    uint32_t v[64];
                    tmem_ld_32x32b_x32(tPV + lane_off + hc * 64, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                    tmem_ld_32x32b_x32(tPV + lane_off + hc * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                    tmem_ld_wait();
                    tcgen05_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(pv_empty);  // PV accumulator is in registers: the MMA warp may reuse it
#pragma unroll
                    for (int i = 0; i < 64; ++i) o_acc[i] = fmaf(o_acc[i], alpha_prev, __uint_as_float(v[i]));
                }
            }

            // p = exp2((s - m) * scale*log2e) -> bf16 into swizzled smem (half-tile hc); partial row sum
            // (S is re-read from TMEM instead of being held across the PV update: the register file is allocated per
            //  4-warp group, which caps this 10-warp CTA at 168 registers per thread)
            uint32_t sv[64];
            tmem_ld_32x32b_x32(tS, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32b_x32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
            tcgen05_fence_before();  // S[s] fully read -> the MMA warp may overwrite it with block j+2
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
            if (nvalid < 64) {  // padding columns: exp2(-inf) = 0
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    if (i >= nvalid) sv[i] = 0xff800000u;
            }
            float rs4[4] = {0.f, 0.f, 0.f, 0.f};
            uint8_t* prow = sP + hc * kHalf + r * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float p0 = ex2_approx(fmaf(__uint_as_float(sv[c * 32 + 2 * i]), scale_log2, mneg));
                    const float a1 = fmaf(__uint_as_float(sv[c * 32 + 2 * i + 1]), scale_log2, mneg);
                    const float p1 = (i & 1) ? ex2_fma(a1) : ex2_approx(a1);  // every 4th element on the FMA pipe
                    rs4[i & 3] += p0 + p1;
                    pk[i] = pack_bf16x2(p0, p1);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int lc = c * 4 + q;
                    *reinterpret_cast<uint4*>(prow + ((lc ^ (r & 7)) << 4)) =
                        make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            const float rs = (rs4[0] + rs4[1]) + (rs4[2] + rs4[3]);
            fence_proxy_async_smem();  // P stores -> visible to the UMMA (async proxy) read
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            const float alpha = ex2_approx(fmaf(m_run, scale_log2, mneg));  // exp2((m_old - m_new)*sl2); 0 on the first block
            l_run = fmaf(l_run, alpha, rs);
            m_run = m_new;
            alpha_prev = alpha;
        }
        // total row sum = the two halves' partial sums (same rescale history)
        red_sum[hc * 128 + r] = l_run;
        softmax_group_sync();
        const float inv_l = 1.0f / (l_run + red_sum[(hc ^ 1) * 128 + r]);
        // last PV
        mbar_wait(pv_full, (n_kv - 1) & 1);
        tcgen05_fence_after();
        const int qrow = qt * 128 + r;
        __nv_bfloat16* orow = out + (size_t)(b * L + (qrow < L ? qrow : 0)) * d_model + h * 128 + hc * 64;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32b_x32(tPV + lane_off + hc * 64 + c * 32, v);
            tmem_ld_wait();
            if (qrow < L) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float a0 = fmaf(o_acc[c * 32 + 2 * i], alpha_prev, __uint_as_float(v[2 * i])) * inv_l;
                    const float a1 = fmaf(o_acc[c * 32 + 2 * i + 1], alpha_prev, __uint_as_float(v[2 * i + 1])) * inv_l;
                    pk[i] = pack_bf16x2(a0, a1);
                }
                uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) d4[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 8) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

int attention_fwd_v3(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B,
                  int H, int L, int Lpad, float scale, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || L <= 0) return set_error("attention: empty problem");
    if (Lpad < L || (Lpad % 8)) return set_error("attention: Lpad must be >= L and a multiple of 8");
    const int d_model = H * 128;
    CUtensorMap tmQ, tmK, tmVt;
    if (make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmK, k, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmVt, vt, (uint64_t)B * H * 128, (uint64_t)Lpad, (uint64_t)Lpad, 128, 64)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        MMDP_CUDA(cudaFuncSetAttribute(attention_v3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
        attr_set = true;
    }
    dim3 grid((L + 127) / 128, H, B);
    const float scale_log2 = scale * 1.4426950408889634f;
    LaunchScope ls(LK_ATTN, 4.0 * B * H * (double)L * L * 128, stream);
    attention_v3_kernel<<<grid, kAttnThreads, kAttnSmem, stream>>>(tmQ, tmK, tmVt, out, H, L, d_model, scale_log2);
    MMDP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mmdp
