// Attention forward v7: softmax(Q K^T / sqrt(128)) V for head_dim 128, no mask, whole-sequence KV - same math, layouts and
// interface as v6 (attention6.cu), different arrangement on the SM:
//   * ONE CTA per SM holds TWO 128-row query tiles of one head ("slots") and walks the keys in blocks of 128. The K and V^T
//     tiles are loaded once for both slots (half the L2 -> shared-memory fill of v6) and S = Q K^T is one N = 128 MMA group, whose
//     shared-memory operand stream is the 128 B/clk an SM has (v6's N = 64 groups need 192 B/clk and run at 2/3 rate);
//   * tensor memory (all 512 columns): S0 | S1 | O0 | O1, 128 columns each. S is single-buffered per slot, so one slot is
//     the strict chain QK(j) -> softmax(j) -> PV(j) -> QK(j+1); the two slots run half a period apart: while the four
//     softmax warps of slot 0 work on S0(j), the tensor core runs PV1(j-1) and QK1(j) - and the softmax warps of a
//     slot have their scheduler's MUFU / issue slots to themselves (in v6 two co-resident CTAs' softmax warps collide:
//     ncu shows 27 % issue efficiency and 32 % of their time waiting for S);
//   * P goes back into the TMEM columns S was read from (bf16 pairs, 64 columns) and is the A operand of the PV MMA;
//     lazy rescale, packed-fp32 exponent arithmetic and the exp2 share on the FMA pipe are v6's (attention_math.cuh);
//   * 18 warps: EIGHT softmax warps per slot - two threads per query row, 64 of the block's 128 scores each (warps w and w + 4
//     of a slot share a TMEM lane quarter); the row max is exchanged through shared memory behind a 64-thread named barrier,
//     the row sums stay per thread until the end. ncu on the one-thread-per-row form: a softmax warp issues 0.28
//     instructions per clock (fixed-latency dependency stalls) and its 2 180 cycles per block were 56 % of the slot's period,
//     so the cure is more warps per scheduler, not fewer instructions. One TMA producer (K 3 stages, V^T 2 stages, loads
//     issued in the order the MMAs consume them), one MMA issuer. 226.25 KB shared memory;
//   * work units: pairs of query tiles (2p, 2p+1) of a head; the odd last tile of a head runs alone in slot 0. Units may be
//     cut along the keys into pieces whose un-normalised partials (O, m, l) are merged by attention_combine7_kernel: the
//     odd tiles always (they are the tail of the launch and a one-slot unit keeps the tensor core idle during its
//     softmax), everything when the launch has fewer units than half the SMs (a tensor-parallel rank).
#include "mmdp_internal.h"
#include "ptx.cuh"
#include "attention_math.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace mmdp {

static constexpr int k7Threads = 576;  // warps 0-7 softmax slot 0 (column half = (warp >> 2) & 1), 8-15 slot 1, 16 TMA producer, 17 MMA issuer
static constexpr int k7WarpTma = 16, k7WarpMma = 17;
static constexpr int k7BKV = 128;
static constexpr int k7KStages = 3;
static constexpr int k7VStages = 2;
static constexpr int k7Tile = 128 * 128 * 2;  // 32 KB: a Q, K or V^T tile = two SWIZZLE_128B boxes of 128 rows x 64 columns
static constexpr int k7Smem = (2 + k7KStages + k7VStages) * k7Tile + 256 + 2 * 2 * 128 * 4;  // + barriers + row-max exchange
static constexpr int k7PartFloats = 128 * 128 + 256;  // one partial: O [128][128], m [128], l [128]

struct Attn7Params {
    __nv_bfloat16* out;
    float* part_ws;
    int H, L, Lq, Lpad, d_model, n_qt;
    float scale_log2;
    int n_pair, c_pair;      // pair units (two tiles each) and the number of key pieces each is cut into (1 = direct output)
    int n_single, c_single;  // odd last tiles (one per (b, h) when n_qt is odd) and their pieces
    int probe;               // timing probes (results are garbage): 1 softmax threads skip their work, 2 no PV MMAs, 3 no QK MMAs
};

template <int POLY>
__global__ void __maxnreg__(96)
attention_v7_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmVt, const Attn7Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw;
    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u)) {
        printf("mmdp: attention smem base not 1024-byte aligned\n");
        __trap();
    }
    uint8_t* sQ = smem;                         // [2] tiles
    uint8_t* sK = sQ + 2 * k7Tile;              // [3]
    uint8_t* sV = sK + k7KStages * k7Tile;      // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + k7VStages * k7Tile);
    uint64_t* q_full = bars + 0;
    uint64_t* k_full = bars + 1;    // [3]
    uint64_t* k_empty = bars + 4;   // [3]
    uint64_t* v_full = bars + 7;    // [2]
    uint64_t* v_empty = bars + 9;   // [2]
    uint64_t* s_full = bars + 11;   // [2] per slot: S(j) complete (and with it every MMA issued before QK(j), PV(j-1) included)
    uint64_t* p_full = bars + 13;   // [2] per slot: the four softmax warps have published P(j)
    uint64_t* o_full = bars + 15;   // [2] per slot: last PV retired
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);
    float* xchg = reinterpret_cast<float*>(bars + 32);  // [slot][column half][row]: row max / row sum exchange between the two threads of a row

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- unit decode: [0, n_pair * c_pair) pair units x pieces, then n_single * c_single single-tile pieces
    const int npp = p.n_qt >> 1;
    const int n_kvb = (p.L + k7BKV - 1) / k7BKV;
    int bh, qt0, ntile, c, piece, part0;  // part0: partial slot of slot 0's tile for piece 0 (slot 1's tile: part0 + c)
    {
        const int u = blockIdx.x;
        const int n_pair_ctas = p.n_pair * p.c_pair;
        if (u < n_pair_ctas) {
            const int unit = u / p.c_pair;
            piece = u - unit * p.c_pair;
            c = p.c_pair;
            bh = unit / npp;
            qt0 = 2 * (unit - bh * npp);
            ntile = 2;
            part0 = unit * 2 * p.c_pair;
        } else {
            const int v = u - n_pair_ctas;
            const int s = v / p.c_single;
            piece = v - s * p.c_single;
            c = p.c_single;
            bh = s;
            qt0 = p.n_qt - 1;
            ntile = 1;
            part0 = (p.c_pair > 1 ? p.n_pair * 2 * p.c_pair : 0) + s * p.c_single;
        }
    }
    const int b = bh / p.H, h = bh - b * p.H;
    const int per = (n_kvb + c - 1) / c;
    const int jb = piece * per;
    const int je = (jb + per < n_kvb) ? jb + per : n_kvb;
    const int n_kv = je - jb;  // >= 1 (the host never creates an empty piece)
    const bool partial = c > 1;

    if (warp == k7WarpMma && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < k7KStages; ++s) {
            mbar_init(&k_full[s], 1);
            mbar_init(&k_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&v_full[s], 1);
            mbar_init(&v_empty[s], 1);
            mbar_init(&s_full[s], 1);
            mbar_init(&p_full[s], 8);
            mbar_init(&o_full[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == k7WarpTma) {
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
    pdl_launch_dependents();  // the prologue above overlaps the tail of the QKV GEMM (programmatic dependent launch)
    pdl_wait();
    // TMEM columns: S(slot) at slot * 128 (P(slot) = packed bf16 pairs in its first 64 columns), O(slot) at 256 + slot * 128

    if (warp == k7WarpTma) {
        // ===================== TMA producer: Q tiles, then K(0), V(0), K(1), V(1), ... =====================
        if (elect_one_sync()) {
            mbar_expect_tx(q_full, ntile * k7Tile);
            for (int t = 0; t < ntile; ++t) {
                const int qrow0 = b * p.Lq + (qt0 + t) * 128;
                tma_load_2d(sQ + t * k7Tile, &tmQ, q_full, h * 128, qrow0);
                tma_load_2d(sQ + t * k7Tile + k7Tile / 2, &tmQ, q_full, h * 128 + 64, qrow0);
            }
            for (int j = 0; j < n_kv; ++j) {
                const int kv0 = (jb + j) * k7BKV;
                const int ks = j % k7KStages, vs = j & 1;
                mbar_wait(&k_empty[ks], ((j / k7KStages) & 1) ^ 1);
                mbar_expect_tx(&k_full[ks], k7Tile);
                tma_load_2d(sK + ks * k7Tile, &tmK, &k_full[ks], h * 128, b * p.L + kv0);
                tma_load_2d(sK + ks * k7Tile + k7Tile / 2, &tmK, &k_full[ks], h * 128 + 64, b * p.L + kv0);
                mbar_wait(&v_empty[vs], ((j >> 1) & 1) ^ 1);
                mbar_expect_tx(&v_full[vs], k7Tile);
                tma_load_2d(sV + vs * k7Tile, &tmVt, &v_full[vs], kv0, bh * 128);
                // second half of the keys; when it lies wholly beyond the padded row (short last block) the first half is loaded
                // again instead of a box entirely out of bounds - those keys are masked (P = 0), the values only have to be finite
                tma_load_2d(sV + vs * k7Tile + k7Tile / 2, &tmVt, &v_full[vs], kv0 + 64 < p.Lpad ? kv0 + 64 : kv0, bh * 128);
            }
        }
        __syncwarp();
    } else if (warp == k7WarpMma) {
        // ===================== MMA issuer =====================
        if (elect_one_sync()) {
            constexpr uint32_t idesc = umma_idesc_bf16(128, 128);
            // S(t) = Q(t) K^T over head_dim 128: 8 K=16 steps, two 64-column halves of Q and K
            auto issue_qk = [&](int t, int ks) {
                const uint32_t aQ = smem_u32(sQ + t * k7Tile), aK = smem_u32(sK + ks * k7Tile);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t off = (k >> 2) * (k7Tile / 2);
                    if (p.probe != 3) umma_bf16_ss(tmem_base + t * 128, umma_desc_kmajor_sw128(aQ + off) + (k & 3) * 2,
                                 umma_desc_kmajor_sw128(aK + off) + (k & 3) * 2, idesc, k != 0);
                }
            };
            mbar_wait(q_full, 0);
            mbar_wait(&k_full[0], 0);
            tcgen05_fence_after();
            for (int t = 0; t < ntile; ++t) {
                issue_qk(t, 0);
                umma_commit(&s_full[t]);
            }
            umma_commit(&k_empty[0]);
            for (int j = 0; j < n_kv; ++j) {
                const int vs = j & 1;
                const bool more = j + 1 < n_kv;
                const int ks1 = (j + 1) % k7KStages;
                mbar_wait(&v_full[vs], (j >> 1) & 1);
                for (int t = 0; t < ntile; ++t) {
                    // O(t) += P(t)(j) V(j): A = P from tensor memory (8 columns per K=16 step), B = V^T (two halves of 64 keys)
                    mbar_wait(&p_full[t], j & 1);
                    tcgen05_fence_after();
                    const uint32_t aV = smem_u32(sV + vs * k7Tile);
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (p.probe != 2) umma_bf16_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8,
                                     umma_desc_kmajor_sw128(aV + (k >> 2) * (k7Tile / 2)) + (k & 3) * 2, idesc, (j | k) != 0);
                    if (t == ntile - 1) umma_commit(&v_empty[vs]);
                    if (more) {
                        // S(t)(j+1): the S / P columns of this slot are free once PV(t)(j) above has read P (in-order execution)
                        if (t == 0) {
                            mbar_wait(&k_full[ks1], ((j + 1) / k7KStages) & 1);
                            tcgen05_fence_after();
                        }
                        issue_qk(t, ks1);
                        umma_commit(&s_full[t]);
                        if (t == ntile - 1) umma_commit(&k_empty[ks1]);
                    } else {
                        umma_commit(&o_full[t]);
                    }
                }
            }
        }
        __syncwarp();
    } else if ((warp >> 3) < ntile) {
        // ===================== softmax (slot = warp / 8; two threads per query row: column half hf = (warp / 4) % 2) =====================
        const int t = warp >> 3, hf = (warp >> 2) & 1, qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t tS = tmem_base + t * 128 + lane_off, tO = tmem_base + 256 + t * 128 + hf * 64 + lane_off;
        float* x_mine = xchg + (t * 2 + hf) * 128 + r;
        const float* x_other = xchg + (t * 2 + (hf ^ 1)) * 128 + r;
        const int bar_id = 1 + t * 4 + qd;  // named barrier of the two warps that share these 32 rows
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
        float m_used = -INFINITY, l_run = 0.f;  // l_run: sum over THIS thread's columns only
        constexpr float kLazy = 8.0f;  // raise the running max only when it is exceeded by more than 2^8

        for (int j = 0; j < n_kv; ++j) {
            const int nvalid = p.L - (jb + j) * k7BKV - hf * 64;  // valid columns among this thread's 64
            mbar_wait(&s_full[t], j & 1);
            tcgen05_fence_after();
            if (p.probe == 1) {
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[t]);
                continue;
            }
            uint32_t sv[64];
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) tmem_ld_32x32b_x16(tS + hf * 64 + q4 * 16, *reinterpret_cast<uint32_t(*)[16]>(&sv[q4 * 16]));
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
            float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
            // row max over both column halves (the partner thread computes the same value: max is exact and commutative)
            *x_mine = mx;
            pair_sync();
            mx = fmaxf(mx, *x_other);
            pair_sync();  // the partner has read before this thread overwrites its slot in the next block

            // lazy rescale decision (v6's): only when this row's max exceeds the max in use by more than 2^kLazy (always on block 0)
            const bool need = (mx - m_used) * p.scale_log2 > kLazy;  // m_used = -inf on block 0 -> true
            const bool any_need = __any_sync(0xffffffffu, need);
            const float m_new = need ? mx : m_used;
            const float alpha = need ? ex2_mufu((m_used - m_new) * p.scale_log2) : 1.0f;  // 0 on the first block
            m_used = m_new;
            const float mneg = -m_used * p.scale_log2;

            if (any_need && j >= 1) {
                // rare: rescale this thread's 64 columns of O (TMEM). S(j) is complete, hence PV(j-1) - issued before QK(j) by
                // the same thread - has retired; PV(j) cannot start before all eight warps publish P(j) below.
#pragma unroll 1
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t v[16];
                    tmem_ld_32x32b_x16(tO + cc * 16, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st_32x32b_x16(tO + cc * 16, v);
                }
            }
            // P = 2^((s - m_used) * c) as bf16 pairs: this thread's 64 scores -> 32 packed columns at hf * 32 of the S region.
            // (the partner may still be reading ITS columns hf' * 64 ... of S: columns [32, 64) written by hf = 1 belong to
            // hf = 0's scores - hence the pair_sync above, after both threads hold their scores in registers)
            uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {  // 16-column stores: short register vectors (the budget is 96 registers with 18 warps)
                uint32_t pk[16];
                softmax_exp_block<32, POLY>(sv + hh * 32, p.scale_log2, mneg, pk, acc);
                tmem_st_32x32b_x16(tS + hf * 32 + hh * 16, pk);
            }
            l_run = fmaf(l_run, alpha, f32x2_sum4(acc));
            tmem_st_wait();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_full[t]);
        }
        // row sum over both halves
        *x_mine = l_run;
        pair_sync();
        const float l_row = (hf == 0) ? l_run + *x_other : *x_other + l_run;  // same operand order in both threads
        // epilogue: O / l (or the un-normalised partial of this key piece); this thread handles columns [hf * 64, hf * 64 + 64)
        mbar_wait(&o_full[t], 0);
        tcgen05_fence_after();
        const int qt = qt0 + t;
        const int qrow = qt * 128 + r;
        if (partial) {
            float* slot = p.part_ws + (size_t)(part0 + t * c + piece) * k7PartFloats;
            if (qrow < p.Lq && hf == 0) {
                slot[128 * 128 + r] = m_used;
                slot[128 * 128 + 128 + r] = l_row;
            }
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tO + cc * 32, v);
                tmem_ld_wait();
                if (qrow < p.Lq) {
                    uint4* d4 = reinterpret_cast<uint4*>(slot + (size_t)r * 128 + hf * 64 + cc * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d4[i] = make_uint4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        } else {
            const float inv_l = 1.0f / l_row;
            __nv_bfloat16* orow = p.out + (size_t)(b * p.Lq + (qrow < p.Lq ? qrow : 0)) * p.d_model + h * 128 + hf * 64;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t v[32];
                tmem_ld_32x32b_x32(tO + cc * 32, v);
                tmem_ld_wait();
                if (qrow < p.Lq) {
                    uint32_t o[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = pack_bf16x2(__uint_as_float(v[2 * i]) * inv_l, __uint_as_float(v[2 * i + 1]) * inv_l);
                    uint4* d4 = reinterpret_cast<uint4*>(orow + cc * 32);
#pragma unroll
                    for (int i = 0; i < 4; ++i) d4[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
                }
            }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == k7WarpTma) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// Merges the key-piece partials of one tile: out = (sum_i w_i O_i) / (sum_i w_i l_i), w_i = 2^((m_i - m) c), pieces in index
// order. Grid = (split tiles, 128 query rows), thread = output column. Split tiles: the 2 * n_pair tiles of the pair units
// first (when c_pair > 1), then the n_single odd tiles (when c_single > 1).
__global__ void __launch_bounds__(128)
attention_combine7_kernel(const Attn7Params p) {
    const int st = blockIdx.x, r = blockIdx.y, col = threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    const int npp = p.n_qt >> 1;
    const int n_a = p.c_pair > 1 ? 2 * p.n_pair : 0;
    int bh, qt, c;
    size_t part0;
    if (st < n_a) {
        const int unit = st >> 1, t = st & 1;
        bh = unit / npp;
        qt = 2 * (unit - bh * npp) + t;
        c = p.c_pair;
        part0 = (size_t)unit * 2 * p.c_pair + (size_t)t * p.c_pair;
    } else {
        const int s = st - n_a;
        bh = s;
        qt = p.n_qt - 1;
        c = p.c_single;
        part0 = (size_t)(p.c_pair > 1 ? p.n_pair * 2 * p.c_pair : 0) + (size_t)s * p.c_single;
    }
    const int qrow = qt * 128 + r;
    if (qrow >= p.Lq) return;
    const int b = bh / p.H, h = bh - b * p.H;
    const float* base = p.part_ws + part0 * k7PartFloats;
    float m = -INFINITY;
    for (int i = 0; i < c; ++i) m = fmaxf(m, base[(size_t)i * k7PartFloats + 128 * 128 + r]);
    float o = 0.f, l = 0.f;
    for (int i = 0; i < c; ++i) {
        const float* slot = base + (size_t)i * k7PartFloats;
        const float w = exp2f((slot[128 * 128 + r] - m) * p.scale_log2);
        o = fmaf(w, slot[(size_t)r * 128 + col], o);
        l = fmaf(w, slot[128 * 128 + 128 + r], l);
    }
    p.out[(size_t)(b * p.Lq + qrow) * p.d_model + h * 128 + col] = __float2bfloat16_rn(o / l);
}

// key-piece partials: one buffer per (device, stream), grown on demand
static std::map<std::pair<int, cudaStream_t>, std::pair<float*, size_t>> g_attn7_ws;
static std::mutex g_attn7_ws_mu;
static int attn7_part_workspace(cudaStream_t stream, size_t need, float** out) {
    int dev = 0;
    MMDP_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_attn7_ws_mu);
    auto& e = g_attn7_ws[std::make_pair(dev, stream)];
    if (need > e.second) {
        if (e.first) MMDP_CUDA(cudaFree(e.first));  // synchronises with the launches that still read it
        e.first = nullptr; e.second = 0;
        MMDP_CUDA(cudaMalloc(&e.first, need));
        e.second = need;
    }
    *out = e.first;
    return 0;
}

int attention_fwd_v7(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream, int Lq) {
    if (B <= 0 || H <= 0 || L <= 0) return set_error("attention: empty problem");
    if (Lq <= 0) Lq = L;
    if (Lpad < L || (Lpad % 8)) return set_error("attention: Lpad must be >= L and a multiple of 8");
    const int d_model = H * 128;
    CUtensorMap tmQ, tmK, tmVt;
    if (make_tmap_2d_bf16(&tmQ, q, (uint64_t)B * Lq, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmK, k, (uint64_t)B * L, (uint64_t)d_model, (uint64_t)d_model, 128, 64)) return -1;
    if (make_tmap_2d_bf16(&tmVt, vt, (uint64_t)B * H * 128, (uint64_t)Lpad, (uint64_t)Lpad, 128, 64)) return -1;
    int poly = opt(OPT_ATTN_POLY);
    if (poly != 0 && poly != 2 && poly != 8) poly = 4;
    Attn7Params p{};
    p.out = out; p.H = H; p.L = L; p.Lq = Lq; p.Lpad = Lpad; p.d_model = d_model;
    p.n_qt = (Lq + 127) / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    p.n_pair = B * H * (p.n_qt >> 1);
    p.n_single = (p.n_qt & 1) ? B * H : 0;
    p.c_pair = p.c_single = 1;
    p.probe = opt(OPT_ATTN_PROBE);
    const int n_kvb = (L + k7BKV - 1) / k7BKV, sms = num_sms();
    // pieces of `want` or fewer key blocks each, none empty: piece i covers blocks [i * per, (i + 1) * per), per = ceil(n_kvb / c)
    auto fit = [&](int want) {
        int c = want > 8 ? 8 : want;
        if (c > n_kvb) c = n_kvb;
        while (c >= 2 && (c - 1) * ((n_kvb + c - 1) / c) >= n_kvb) --c;
        return c < 1 ? 1 : c;
    };
    if (opt(OPT_ATTN_SPLIT_TAIL)) {
        const int units = p.n_pair + p.n_single;
        if (units * 2 <= sms) {
            // fewer units than half the SMs (a tensor-parallel rank with 4 of the 32 heads): cut every unit so that the whole machine works
            p.c_pair = p.c_single = fit(sms / units);
        } else if (p.n_single > 0) {
            // the odd tiles are the tail of the launch; a one-slot unit runs at half the tensor rate - cut each into 4 pieces
            p.c_single = fit(4);
        }
    }
    const size_t n_part = (size_t)(p.c_pair > 1 ? 2 * p.n_pair * p.c_pair : 0) + (size_t)(p.c_single > 1 ? p.n_single * p.c_single : 0);
    if (n_part > 0 && attn7_part_workspace(stream, n_part * k7PartFloats * sizeof(float), &p.part_ws)) return -1;
    if (p.c_pair > 1 && p.c_single == 1 && p.n_single > 0) return set_error("attention v7: inconsistent split plan");
    const int grid = p.n_pair * p.c_pair + p.n_single * p.c_single;
    const bool pdl = pdl_mode() != 0;
    LaunchScope ls(LK_ATTN, 4.0 * B * H * (double)Lq * L * 128, stream);
    auto launch = [&](auto kernel) -> int {
        static unsigned long long attr_set = 0;  // bit per device (one instantiation per POLY value)
        int dev = 0;
        MMDP_CUDA(cudaGetDevice(&dev));
        if (!(attr_set >> (dev & 63) & 1ull)) {
            MMDP_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, k7Smem));
            attr_set |= 1ull << (dev & 63);
        }
        MMDP_CUDA(launch_ex(kernel, dim3(grid), dim3(k7Threads), k7Smem, stream, pdl, false, tmQ, tmK, tmVt, p));
        return 0;
    };
    if (poly == 2) { if (launch(attention_v7_kernel<2>)) return -1; }
    else if (poly == 4) { if (launch(attention_v7_kernel<4>)) return -1; }
    else if (poly == 8) { if (launch(attention_v7_kernel<8>)) return -1; }
    else { if (launch(attention_v7_kernel<0>)) return -1; }
    MMDP_CUDA(cudaGetLastError());
    const int n_split_tiles = (p.c_pair > 1 ? 2 * p.n_pair : 0) + (p.c_single > 1 ? p.n_single : 0);
    if (n_split_tiles > 0) MMDP_CUDA(launch_ex(attention_combine7_kernel, dim3(n_split_tiles, 128), dim3(128), 0, stream, pdl, false, p));
    return 0;
}

}  // namespace mmdp
