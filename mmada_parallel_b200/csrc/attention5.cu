// Attention forward v5 (alternate, MMDP_ATTN=5): same math / layouts / interface as v6 (attention6.cu) with 128-wide KV
// blocks - the FlashAttention-4 arrangement, with the two query tiles of an SM held by two co-resident CTAs:
//   * S = Q K^T is one 128 x 128 x 128 MMA group (N = 128 keeps the shared-memory operand traffic of the QK MMA at the
//     128 B/clk an SM has; with N = 64 it needs 192 B/clk), 128 TMEM columns, single-buffered;
//   * P (bf16) is written by the softmax threads with tcgen05.st INTO the TMEM columns S was read from (columns 0..63
//     of the S region, two bf16 per column) and the PV MMA takes its A operand from tensor memory; O (128 x 128 fp32)
//     stays in TMEM columns 128..255;
//   * S / P share their columns, so one CTA is a strict chain QK(j) -> softmax(j) -> PV(j) -> QK(j+1); TWO CTAs per SM
//     (96 KB smem, 256 TMEM columns, <= 168 registers each) interleave so that one runs its MMAs while the other runs
//     its softmax; single-stage K / V tiles are enough because the next tile has a whole softmax to arrive;
//   * softmax is ONE pass in the common case: P = 2^((s - m_used) c) is computed with the running max in use while the
//     row max of this block is tracked on the side; only if that max exceeds m_used by more than 2^8 (or on block 0) the
//     block is recomputed with the new max and O (TMEM) / l are rescaled (lazy rescale);
//   * packed-fp32 softmax arithmetic (attention_math.cuh): 3 instructions per score.
// Measured (B=1, L=2414, H=32): 672 TFLOP/s isolated (v6: 697) - kept because its MMA shapes are the better base once
// the softmax time per block drops below the MMA time.
#include "mmdp_internal.h"
#include "ptx.cuh"
#include "attention_math.cuh"

namespace mmdp {

static constexpr int k5Threads = 192;
static constexpr int k5BKV = 128;
static constexpr int k5TileBytes = 128 * 128 * 2;  // Q, K and V^T tiles: 32 KB each (two 64-column SWIZZLE_128B boxes)
static constexpr int k5Smem = 3 * k5TileBytes + 256;

// One sweep over the 128 S columns of this thread's row, 16 columns per tcgen05.ld, the next load in flight while the
// current chunk is processed. Tracks the row max; with kExp also P = 2^(s c + mneg) (bf16 pairs in pk) and its row sum.
template <bool kExp>
__device__ __forceinline__ void softmax_sweep(uint32_t taddr, int nvalid, float scale_log2, float mneg, float& mx, float& sum,
                                              uint32_t (&pk)[64]) {
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
    uint32_t sa[16], sb[16];
    tmem_ld_32x32b_x16(taddr, sa);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint32_t(&cur)[16] = (c & 1) ? sb : sa;
        uint32_t(&nxt)[16] = (c & 1) ? sa : sb;
        tmem_ld_wait();
        if (c < 7) tmem_ld_32x32b_x16(taddr + (c + 1) * 16, nxt);
        if (nvalid < k5BKV) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c * 16 + i >= nvalid) cur[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(cur[i]));
        if (kExp) softmax_exp_block<16>(cur, scale_log2, mneg, &pk[c * 8], acc);
    }
    mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    sum = f32x2_sum4(acc);
}

__global__ void __launch_bounds__(k5Threads, 2)
attention_v5_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ out, int H, int L, int d_model,
                    float scale_log2) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
        printf("mmdp: attention smem base not 1024-byte aligned\n");
        __trap();
    }
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + k5TileBytes;
    uint8_t* sV = sK + k5TileBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + k5TileBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;
    uint64_t* k_empty = bars + 2;
    uint64_t* v_full = bars + 3;
    uint64_t* v_empty = bars + 4;
    uint64_t* s_full = bars + 5;
    uint64_t* p_full = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int n_kv = (L + k5BKV - 1) / k5BKV;

    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(k_full, 1);
        mbar_init(k_empty, 1);
        mbar_init(v_full, 1);
        mbar_init(v_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(p_full, 4);
        mbar_init(o_full, 1);
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
    const uint32_t tS = tmem_base, tO = tmem_base + 128;  // S cols 0..127 (P = packed bf16 in cols 0..63), O cols 128..255

    if (warp == 4) {
        // ===================== TMA producer =====================
        if (elect_one_sync()) {
            const int qrow0 = b * L + qt * 128;
            mbar_expect_tx(q_full, k5TileBytes);
            tma_load_2d(sQ, &tmQ, q_full, h * 128, qrow0);
            tma_load_2d(sQ + k5TileBytes / 2, &tmQ, q_full, h * 128 + 64, qrow0);
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t u = j & 1;
                const int kv0 = j * k5BKV;
                mbar_wait(k_empty, u ^ 1);
                mbar_expect_tx(k_full, k5TileBytes);
                tma_load_2d(sK, &tmK, k_full, h * 128, b * L + kv0);
                tma_load_2d(sK + k5TileBytes / 2, &tmK, k_full, h * 128 + 64, b * L + kv0);
                mbar_wait(v_empty, u ^ 1);
                mbar_expect_tx(v_full, k5TileBytes);
                tma_load_2d(sV, &tmVt, v_full, kv0, (b * H + h) * 128);
                tma_load_2d(sV + k5TileBytes / 2, &tmVt, v_full, kv0 + 64, (b * H + h) * 128);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aV = smem_u32(sV);
            mbar_wait(q_full, 0);
            for (int j = 0; j < n_kv; ++j) {
                const uint32_t u = j & 1;
                // S(j) = Q K(j)^T. The S columns are free: PV(j-1), which read P(j-1) from them, was issued before (the
                // tensor core executes this thread's MMAs in order) and the softmax warps finished with S(j-1) before
                // they published P(j-1).
                mbar_wait(k_full, u);
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {  // K dimension = head_dim 128: two 64-column halves of Q and K
                    const uint32_t off = (k >> 2) * (k5TileBytes / 2);
                    umma_bf16_ss(tS, umma_desc_kmajor_sw128(aQ + off) + (k & 3) * 2, umma_desc_kmajor_sw128(aK + off) + (k & 3) * 2,
                                 idesc, k != 0);
                }
                umma_commit(k_empty);
                umma_commit(s_full);
                // O += P(j) V(j): A = P from tensor memory (8 columns per K=16 step), B = V^T tile (two 64-kv halves)
                mbar_wait(v_full, u);
                mbar_wait(p_full, u);
                tcgen05_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (k >> 2) * (k5TileBytes / 2);
                    umma_bf16_ts(tO, tS + k * 8, umma_desc_kmajor_sw128(aV + off) + (k & 3) * 2, idesc, (j | k) != 0);
                }
                umma_commit(v_empty);
            }
            umma_commit(o_full);
        }
        __syncwarp();
    } else {
        // ===================== softmax (warps 0..3, thread = query row) =====================
        const int r = warp * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
        float m_used = -INFINITY, l_run = 0.f;
        constexpr float kLazy = 8.0f;  // raise the running max only when it is exceeded by more than 2^8

        for (int j = 0; j < n_kv; ++j) {
            const int nvalid = L - j * k5BKV;
            mbar_wait(s_full, j & 1);
            tcgen05_fence_after();
            uint32_t pk[64];
            float mx, sum = 0.f;
            if (j == 0)
                softmax_sweep<false>(tS + lane_off, nvalid, scale_log2, 0.f, mx, sum, pk);
            else
                softmax_sweep<true>(tS + lane_off, nvalid, scale_log2, -m_used * scale_log2, mx, sum, pk);
            const bool need = (mx - m_used) * scale_log2 > kLazy;  // m_used = -inf on block 0 -> true
            if (__any_sync(0xffffffffu, need)) {
                // slow path (always on block 0, afterwards only when a row max jumps): new max, recompute P, rescale O / l.
                // O is quiescent here: PV(j-1) completed before S(j) was committed and PV(j) waits for this P.
                const float m_new = need ? mx : m_used;
                const float alpha = (j == 0) ? 0.f : (need ? ex2_mufu((m_used - m_new) * scale_log2) : 1.0f);
                m_used = m_new;
                float mx2;
                softmax_sweep<true>(tS + lane_off, nvalid, scale_log2, -m_used * scale_log2, mx2, sum, pk);
                if (j > 0) {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                        tmem_st_32x32b_x32(tO + lane_off + c * 32, v);
                    }
                }
                l_run = fmaf(l_run, alpha, sum);
            } else {
                l_run += sum;
            }
            // publish P(j): packed bf16 pairs into columns 0..63 of the S region (all 128 S columns of this row are in
            // registers / consumed by now)
            tmem_st_32x32b_x32(tS + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&pk[0]));
            tmem_st_32x32b_x32(tS + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&pk[32]));
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        // epilogue: O / l
        mbar_wait(o_full, 0);
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
                uint32_t o[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
                uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) d4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
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

int attention_fwd_v5(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream) {
    if (B <= 0 || H <= 0 || L <= 0) return set_error("attention: empty problem");
    if (Lpad < L || (Lpad % 8)) return set_error("attention: Lpad must be >= L and a multiple of 8");
    const int d_model = H * 128;
    CUtensorMap tmQ, tmK, tmVt;
    if (make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmK, k, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmVt, vt, (uint64_t)B * H * 128, (uint64_t)Lpad, (uint64_t)Lpad, 128, 64)) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        MMDP_CUDA(cudaFuncSetAttribute(attention_v5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k5Smem));
        MMDP_CUDA(cudaFuncSetAttribute(attention_v5_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        attr_set = true;
    }
    dim3 grid((L + 127) / 128, H, B);
    const float scale_log2 = scale * 1.4426950408889634f;
    LaunchScope ls(LK_ATTN, 4.0 * B * H * (double)L * L * 128, stream);
    attention_v5_kernel<<<grid, k5Threads, k5Smem, stream>>>(tmQ, tmK, tmVt, out, H, L, d_model, scale_log2);
    MMDP_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace mmdp
