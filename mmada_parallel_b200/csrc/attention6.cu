// Attention forward v6 (default): softmax(Q K^T / sqrt(128)) V for head_dim 128, no mask, whole-sequence KV.
// The third generation of this kernel (earlier ones kept O in registers and passed P through shared memory); restructured
// around tensor memory after three ncu passes
// (profiles/r01/README.md):
//   * one CTA = one 128-row query tile of one head; KV blocks of 64; 7 warps: 4 softmax (thread = query row), K producer,
//     MMA issuer, V^T producer; TWO CTAs per SM (112 KB smem, 256 TMEM columns, 107 registers each) so that one CTA's
//     MMAs run under the other's softmax;
//   * S = Q K^T (128 x 64 fp32) is double-buffered in TMEM columns 0..127, the output accumulator O (128 x 128 fp32)
//     lives in columns 128..255 for the whole KV loop (tcgen05.mma accumulate) - no O registers in the softmax threads;
//   * P never touches shared memory: the softmax threads write the bf16 pairs with tcgen05.st into the TMEM columns of the
//     S buffer they were read from, and the PV MMA takes its A operand from tensor memory (tcgen05.mma [d], [a], b_desc).
//     That removes the st.shared + fence.proxy.async of the smem route (15 % of the softmax warps' samples before);
//   * lazy rescale (FlashAttention-4): the running max m_used is only raised when a row max exceeds it by more than 2^8;
//     only then O and l are multiplied by 2^(m_old - m_new) through tcgen05.ld / tcgen05.st;
//   * K (3 stages) and V^T (2 stages) tiles have their own producer warps: a single in-order producer held the next K tile
//     back until the previous V^T slot was free, and the softmax warps waited 25 % of their time for S;
//   * the S buffers need no "empty" barrier: PV(j) (which waits for P(j), i.e. for the softmax to be done with S(j)) is
//     issued before QK(j+2) and the tensor core executes one thread's MMAs in order;
//   * packed-fp32 softmax arithmetic (attention_math.cuh): 3 instructions per score.
//   * the tiles of a partial last wave are cut along the keys and merged by attention_combine_kernel (see attention_fwd_v6).
// Measured (B=1, L=2414, H=32, round 2): 1067 TFLOP/s isolated (clocks ~1.6 GHz), 768 TFLOP/s inside the power-capped denoising
// loop. What bounds it - the ~0.5 us hand-offs between the MMA thread and the softmax warps, twice per block - and why the
// one-CTA-per-SM arrangement (attention7.cu) does not beat it: DESIGN.md section 5.
#include "mmdp_internal.h"
#include "ptx.cuh"
#include "attention_math.cuh"

#include <stdlib.h>

#include <map>
#include <mutex>
#include <utility>

namespace mmdp {

static constexpr int k6Threads = 224;  // 4 softmax warps, K producer, MMA issuer, V producer
static constexpr int k6BKV = 64;
static constexpr int k6KStages = 3;
static constexpr int k6VStages = 2;
static constexpr int k6QBytes = 128 * 128 * 2;    // 32 KB (two 64-column halves)
static constexpr int k6KBytes = k6BKV * 128 * 2;  // 16 KB (two halves of 64 rows x 64 cols)
static constexpr int k6VBytes = 128 * k6BKV * 2;  // 16 KB (128 d rows x 64 kv)
// smem: Q | K[3] | V[2] | barriers = 112.25 KB; no alignment slack (the dynamic smem window of a kernel without static
// smem starts 1024-aligned; checked at run time) so that two CTAs fit into the 227 KB of an SM
static constexpr int k6Smem = k6QBytes + k6KStages * k6KBytes + k6VStages * k6VBytes + 256;

template <int POLY>
__global__ void __launch_bounds__(k6Threads, 2)
attention_v6_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, __nv_bfloat16* __restrict__ out, int H, int L, int d_model,
                    float scale_log2, int n_full, int splits, float* __restrict__ part_ws, int Lq) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
        printf("mmdp: attention smem base not 1024-byte aligned\n");
        __trap();
    }
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + k6QBytes;
    uint8_t* sV = sK + k6KStages * k6KBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + k6VStages * k6VBytes);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [3]
    uint64_t* k_empty = bars + 4;   // [3]
    uint64_t* v_full = bars + 7;    // [2]
    uint64_t* v_empty = bars + 9;   // [2]
    uint64_t* s_full = bars + 11;   // [2]
    uint64_t* p_full = bars + 13;   // [2]
    uint64_t* pv_done = bars + 15;  // one phase per PV(j): only the lazy-rescale path waits on it (see there)
    uint64_t* o_full = bars + 16;   // all PV MMAs retired: the epilogue may read O
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // work items: [0, n_full) whole (b, h, query tile) problems; then, for each of the remaining tiles (the partial last wave),
    // `splits` pieces that each cover a slice of the KV blocks and leave an un-normalised partial (O, m, l) in part_ws for
    // attention_combine_kernel. tile index = (b * H + h) * n_qt + qt.
    // L = number of keys per batch row; Lq = number of query rows per batch row (== L except in the token-cache forward)
    const int n_qt = (Lq + 127) / 128;
    const int n_kv_all = (L + k6BKV - 1) / k6BKV;
    int tile = blockIdx.x, piece = -1;
    if ((int)blockIdx.x >= n_full) {
        const int t = (int)blockIdx.x - n_full;
        tile = n_full + t / splits;
        piece = t - (t / splits) * splits;
    }
    const int qt = tile % n_qt, h = (tile / n_qt) % H, b = tile / (n_qt * H);
    int jb = 0, je = n_kv_all;
    if (piece >= 0) {
        const int per = (n_kv_all + splits - 1) / splits;
        jb = piece * per;
        je = (jb + per < n_kv_all) ? jb + per : n_kv_all;
    }
    const int n_kv = je - jb;  // >= 1: the host keeps splits <= number of KV blocks / 2

    if (warp == 5 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < k6KStages; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 4);
        }
        mbar_init(pv_done, 1);
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
    pdl_launch_dependents();  // the prologue above overlaps the tail of the QKV GEMM (programmatic dependent launch)
    pdl_wait();
    // S[0] cols 0..63, S[1] cols 64..127 (P(j) = packed bf16 pairs in the first 32 columns of S[j & 1]), O cols 128..255
    const uint32_t tS0 = tmem_base, tO = tmem_base + 128;

    if (warp == 4) {
        // ===================== TMA producer: Q, then the K tiles =====================
        if (elect_one_sync()) {
            const int qrow0 = b * Lq + qt * 128;
            mbar_expect_tx(q_full, k6QBytes);
            tma_load_2d(sQ, &tmQ, q_full, h * 128, qrow0);
            tma_load_2d(sQ + k6QBytes / 2, &tmQ, q_full, h * 128 + 64, qrow0);
            int st = 0;
            uint32_t ph = 0;
            for (int j = 0; j < n_kv; ++j) {
                const int kv0 = (jb + j) * k6BKV;
                mbar_wait(&k_empty[st], ph ^ 1);
                mbar_expect_tx(&k_full[st], k6KBytes);
                tma_load_2d(sK + st * k6KBytes, &tmK, &k_full[st], h * 128, b * L + kv0);
                tma_load_2d(sK + st * k6KBytes + k6KBytes / 2, &tmK, &k_full[st], h * 128 + 64, b * L + kv0);
                if (++st == k6KStages) { st = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp == 6) {
        // ===================== TMA producer: V^T tiles =====================
        if (elect_one_sync()) {
            for (int j = 0; j < n_kv; ++j) {
                const int st = j & 1;
                mbar_wait(&v_empty[st], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&v_full[st], k6VBytes);
                tma_load_2d(sV + st * k6VBytes, &tmVt, &v_full[st], (jb + j) * k6BKV, (b * H + h) * 128);
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, k6BKV);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128);
            const uint32_t aQ = smem_u32(sQ);
            mbar_wait(q_full, 0);
            int kst = 0;
            uint32_t kph = 0;
            for (int j = 0; j <= n_kv; ++j) {
                if (j < n_kv) {
                    // S[j & 1] = Q K(j)^T. The buffer is free: PV(j-2), which read P(j-2) from it, was issued in the
                    // previous iteration (the tensor core executes this thread's MMAs in order) and waited for p_full.
                    const uint32_t aK = smem_u32(sK + kst * k6KBytes);
                    mbar_wait(&k_full[kst], kph);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {  // K dimension = head_dim 128: two 64-column halves of Q and K
                        const uint32_t qoff = (k >> 2) * (k6QBytes / 2), koff = (k >> 2) * (k6KBytes / 2);
                        umma_bf16_ss(tS0 + (j & 1) * k6BKV, umma_desc_kmajor_sw128(aQ + qoff) + (k & 3) * 2,
                                     umma_desc_kmajor_sw128(aK + koff) + (k & 3) * 2, idesc_qk, k != 0);
                    }
                    umma_commit(&k_empty[kst]);
                    umma_commit(&s_full[j & 1]);
                    if (++kst == k6KStages) { kst = 0; kph ^= 1; }
                }
                if (j >= 1) {
                    // O += P(j-1) V(j-1): A = P from tensor memory (8 columns per K=16 step)
                    const int jj = j - 1, s = jj & 1;
                    const uint32_t u = (jj >> 1) & 1;
                    const uint32_t aV = smem_u32(sV + s * k6VBytes);
                    mbar_wait(&v_full[s], u);
                    mbar_wait(&p_full[s], u);
                    tcgen05_fence_after();
#pragma unroll
                    for (int k = 0; k < k6BKV / 16; ++k)
                        umma_bf16_ts(tO, tS0 + s * k6BKV + k * 8, umma_desc_kmajor_sw128(aV) + k * 2, idesc_pv, (jj | k) != 0);
                    umma_commit(&v_empty[s]);
                    umma_commit(pv_done);
                }
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
            const int s = j & 1;
            const int nvalid = L - (jb + j) * k6BKV;
            mbar_wait(&s_full[s], (j >> 1) & 1);
            tcgen05_fence_after();
            uint32_t sv[64];
            tmem_ld_32x32b_x32(tS0 + s * k6BKV + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
            tmem_ld_32x32b_x32(tS0 + s * k6BKV + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
            tmem_ld_wait();
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
            const float alpha = need ? ex2_mufu((m_used - m_new) * scale_log2) : 1.0f;  // 0 on the first block
            m_used = m_new;
            const float mneg = -m_used * scale_log2;

            // P = 2^((s - m_used) * c) as bf16 pairs; row sum in fp32 (packed-fp32 arithmetic, see attention_math.cuh)
            uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
            uint32_t pk[32];
            softmax_exp_block<64, POLY>(sv, scale_log2, mneg, pk, acc);
            l_run = fmaf(l_run, alpha, f32x2_sum4(acc));

            if (any_need && j >= 1) {
                // rare: rescale O (TMEM). PV(j-1) must have retired first; PV(j) cannot start before this P is published.
                // The parity wait is unambiguous here: S(j) is ready, so QK(j) and everything issued before it (PV(j-2))
                // has completed - pv_done has finished j-1 or j phases, never fewer.
                mbar_wait(pv_done, (j - 1) & 1);
                tcgen05_fence_after();
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
            // publish P(j) in the first 32 columns of S[s] (this row's 64 S values are in registers)
            tmem_st_32x32b_x32(tS0 + s * k6BKV + lane_off, pk);
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[s]);
        }
        // epilogue: O / l. Not pv_done: when the last softmax block is done only PV(n_kv-3) is known to have retired, and
        // a parity wait two phases ahead of the barrier returns at once (it did: rare wrong rows, found by the determinism
        // check of tests/test_gpu_model.py::test_full_size_properties).
        mbar_wait(o_full, 0);
        tcgen05_fence_after();
        const int qrow = qt * 128 + r;
        if (piece >= 0) {
            // partial result of this KV slice: O un-normalised (scaled by 2^(-m_used c)), m_used, l - merged by the combine kernel
            float* slot = part_ws + (size_t)((tile - n_full) * splits + piece) * (128 * 128 + 256);
            if (qrow < Lq) {
                slot[128 * 128 + r] = m_used;
                slot[128 * 128 + 128 + r] = l_run;
            }
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
                tmem_ld_wait();
                if (qrow < Lq) {
                    uint4* d4 = reinterpret_cast<uint4*>(slot + (size_t)r * 128 + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d4[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        } else {
            const float inv_l = 1.0f / l_run;
            __nv_bfloat16* orow = out + (size_t)(b * Lq + (qrow < Lq ? qrow : 0)) * d_model + h * 128;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tO + lane_off + c * 32, v);
                tmem_ld_wait();
                if (qrow < Lq) {
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
                    uint4* d4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                }
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

// Merges the `splits` KV-slice partials of one split tile: out = (sum_i w_i O_i) / (sum_i w_i l_i), w_i = 2^((m_i - m) c).
// One CTA per (split tile, query row), thread = output column.
__global__ void __launch_bounds__(128)
attention_combine_kernel(const float* __restrict__ part_ws, __nv_bfloat16* __restrict__ out, int H, int Lq, int d_model,
                         float scale_log2, int n_full, int splits) {
    const int n_qt = (Lq + 127) / 128;
    const int st = blockIdx.x, r = blockIdx.y, c = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int tile = n_full + st;
    const int qt = tile % n_qt, h = (tile / n_qt) % H, b = tile / (n_qt * H);
    const int qrow = qt * 128 + r;
    if (qrow >= Lq) return;
    const float* base = part_ws + (size_t)st * splits * (128 * 128 + 256);
    float m = -INFINITY;
    for (int i = 0; i < splits; ++i) m = fmaxf(m, base[(size_t)i * (128 * 128 + 256) + 128 * 128 + r]);
    float o = 0.f, l = 0.f;
    for (int i = 0; i < splits; ++i) {
        const float* slot = base + (size_t)i * (128 * 128 + 256);
        const float w = exp2f((slot[128 * 128 + r] - m) * scale_log2);
        o = fmaf(w, slot[(size_t)r * 128 + c], o);
        l = fmaf(w, slot[128 * 128 + 128 + r], l);
    }
    out[(size_t)(b * Lq + qrow) * d_model + h * 128 + c] = __float2bfloat16_rn(o / l);
}

// KV-slice partials of the split tail: one buffer per (device, stream), grown on demand
static std::map<std::pair<int, cudaStream_t>, std::pair<float*, size_t>> g_attn_ws;
static std::mutex g_attn_ws_mu;
static int attn_part_workspace(cudaStream_t stream, size_t need, float** out) {
    int dev = 0;
    MMDP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_attn_ws_mu);
    auto& e = g_attn_ws[std::make_pair(dev, stream)];
    if (need > e.second) {
        if (e.first) MMDP_CUDA(cudaFree(e.first));  // synchronises with the launches that still read it
        e.first = nullptr; e.second = 0;
        MMDP_CUDA(cudaMalloc(&e.first, need));
        e.second = need;
    }
    *out = e.first;
    return 0;
}

int attention_fwd_v6(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream, int Lq) {
    if (B <= 0 || H <= 0 || L <= 0) return set_error("attention: empty problem");
    if (Lq <= 0) Lq = L;
    if (Lpad < L || (Lpad % 8)) return set_error("attention: Lpad must be >= L and a multiple of 8");
    const int d_model = H * 128;
    CUtensorMap tmQ, tmK, tmVt;
    if (make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * Lq, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmK, k, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, k6BKV, 64)) return -1;
    if (make_tmap_2d_bf16(&tmVt, vt, (uint64_t)B * H * 128, (uint64_t)Lpad, (uint64_t)Lpad, 128, 64)) return -1;
    // share of exp2 evaluated on the FMA pipe: every POLY-th pair's second element (0 = none; default 4 = 1/8 of all exp2). MMDP_ATTN_POLY = 0|2|4|8.
    int poly = opt(OPT_ATTN_POLY);  // measured 914 / 916 / 925 / 889 TFLOP/s for 0 / 8 / 4 / 2 (B=1, L=2414, H=32)
    if (poly != 0 && poly != 2 && poly != 8) poly = 4;
    const int split_tail = opt(OPT_ATTN_SPLIT_TAIL);
    // Partial last wave: tiles % (2 CTAs x SMs) leftover tiles would run alone at the end (608 tiles on 296 slots: 30 % of the
    // launch at B=1). Split their KV range over the idle slots and merge the partials (attention_combine_kernel).
    const int n_qt = (Lq + 127) / 128, n_kvb = (L + k6BKV - 1) / k6BKV;
    const int tiles = n_qt * H * B, slots = 2 * num_sms();
    int n_full = tiles, n_split = 0, splits = 1;
    auto fit_splits = [&](int want) {
        int sp = want > 8 ? 8 : want;
        if (sp > n_kvb / 2) sp = n_kvb / 2;
        // no empty piece: piece i covers KV blocks [i * per, (i + 1) * per) with per = ceil(n_kvb / splits)
        while (sp >= 2 && (sp - 1) * ((n_kvb + sp - 1) / sp) >= n_kvb) --sp;
        return sp;
    };
    if (split_tail && tiles > slots && (tiles % slots) > 0 && (tiles % slots) * 4 <= slots) {
        n_split = tiles % slots;
        splits = fit_splits(slots / n_split);
        if (splits >= 2) n_full = tiles - n_split; else { n_split = 0; splits = 1; }
    } else if (split_tail && tiles * 2 <= slots) {
        // fewer tiles than half the CTA slots (a tensor-parallel rank with 4 of the 32 heads: 76 tiles on 296 slots): split EVERY
        // tile's KV range so that the whole machine works on the launch
        splits = fit_splits(slots / tiles);
        if (splits >= 2) { n_split = tiles; n_full = 0; } else splits = 1;
    }
    float* part_ws = nullptr;
    if (n_split > 0 && attn_part_workspace(stream, (size_t)n_split * splits * (128 * 128 + 256) * sizeof(float), &part_ws)) return -1;
    const int grid = n_full + n_split * splits;
    const float scale_log2 = scale * 1.4426950408889634f;
    const bool pdl = pdl_mode() != 0;
    LaunchScope ls(LK_ATTN, 4.0 * B * H * (double)Lq * L * 128, stream);
    auto launch = [&](auto kernel) -> int {
        static unsigned long long attr_set = 0;  // bit per device (one instantiation per POLY value)
        int dev = 0;
        MMDP_CUDA(cudaGetDevice(&dev));
        if (!(attr_set >> (dev & 63) & 1ull)) {
            MMDP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k6Smem));
            MMDP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
            attr_set |= 1ull << (dev & 63);
        }
        MMDP_CUDA(launch_ex(kernel, dim3(grid), dim3(k6Threads), k6Smem, stream, pdl, false, tmQ, tmK, tmVt, out, H, L, d_model,
                            scale_log2, n_full, splits, part_ws, Lq));
        return 0;
    };
    if (poly == 2) { if (launch(attention_v6_kernel<2>)) return -1; }
    else if (poly == 4) { if (launch(attention_v6_kernel<4>)) return -1; }
    else if (poly == 8) { if (launch(attention_v6_kernel<8>)) return -1; }
    else { if (launch(attention_v6_kernel<0>)) return -1; }
    MMDP_CUDA(cudaGetLastError());
    if (n_split > 0)
        MMDP_CUDA(launch_ex(attention_combine_kernel, dim3(n_split, 128), dim3(128), 0, stream, pdl, false, (const float*)part_ws, out, H, Lq,
                            d_model, scale_log2, n_full, splits));
    return 0;
}

}  // namespace mmdp
