// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = A[M,K] · W[N,K]^T  (fp32 accumulate in TMEM)
//
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one thread issues tcgen05.mma 128x256x16, tcgen05.commit frees smem slots)
//   warp 2      : TMEM allocator (512 columns = 2 accumulator stages of 128x256 fp32)
//   warps 4..7  : epilogue       (tcgen05.ld 32x32b, fused epilogue, global stores) - overlaps the next tile's MMAs
//
// Every nn.Linear of the reference block (MMaDA-Parallel-A/model/modeling_llada.py:925-927, :744, :962, :968, :1402)
// maps to one launch of this kernel with a fused epilogue that reproduces the reference's bf16 rounding points:
//   EPI_PLAIN   : C = bf16(acc)                                            (LM head, :1402)
//   EPI_RESID   : C = bf16( bf16(acc) + resid )                            (attn_out + residual :744/:953; ff_out :968/:970)
//   EPI_QKVROPE : q,k = bf16( rope_fp32( bf16(acc) ) ), v^T = bf16(acc)    (q/k/v_proj :925-927 + RotaryEmbedding :402-435)
//   EPI_SWIGLU  : C = bf16( bf16(silu(bf16(g))) * bf16(u) )                (ff_proj/up_proj/act/mul :962-967)
#include "mmdp_internal.h"
#include "ptx.cuh"

namespace mmdp {

static constexpr int BM = 128, BK = 64;
static constexpr int kABytes = BM * BK * 2;  // 16 KB
static constexpr int kGemmThreads = 256;
// The N tile width is a template parameter: 256 (default; required by the QKV/SwiGLU epilogues) or 192. The K loop and
// therefore the fp32 accumulation order of every output element is identical for both, so results do not depend on
// the tile width; the host picks the width that minimises (waves x width) for the problem (wave quantisation on 148 SMs).
template <int BN> struct GemmCfg {
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kStages = (BN == 256) ? 4 : 5;
    static constexpr int kSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
    int M, N, K;
    __nv_bfloat16* C;
    int ldc;
    const __nv_bfloat16* resid;
    int ldr;
    // EPI_QKVROPE
    __nv_bfloat16* q;
    __nv_bfloat16* k;
    __nv_bfloat16* vt;
    const float* cos_tab;  // [L, 64]
    const float* sin_tab;  // [L, 64]
    int L, Lpad, d_model, n_heads;
};

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const uint32_t (&p)[16], int ncols_valid) {
    // dst is 16-B aligned when ldc % 8 == 0 and column offsets are multiples of 8
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (i * 8 < ncols_valid) d4[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
    }
}

template <int EPI, int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int kStages = GemmCfg<BN>::kStages;
    constexpr int kStageBytes = GemmCfg<BN>::kStageBytes;
    static_assert(EPI == EPI_PLAIN || EPI == EPI_RESID || BN == 256, "fused QKV / SwiGLU epilogues need 256-wide tiles");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_ptr);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_m = (p.M + BM - 1) / BM;
    const int num_n = (p.N + BN - 1) / BN;
    const int num_k = (p.K + BK - 1) / BK;
    const int num_tiles = num_m * num_n;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    uint8_t* sa = smem + s * kStageBytes;
                    tma_load_2d(sa, &tmA, &full_bar[s], kb * BK, m_blk * BM);
                    tma_load_2d(sa + kABytes, &tmB, &full_bar[s], kb * BK, n_blk * BN);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kStageBytes);
                    const uint64_t adesc = umma_desc_kmajor_sw128(sa);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(sa + kABytes);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes per 16 bf16 of K inside the 128-B swizzle atom
                        umma_bf16_ss(d_tmem, adesc + (k * 2), bdesc + (k * 2), idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                    if (kb == num_k - 1) umma_commit(&tmem_full[as]);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4;  // == warp % 4 -> TMEM lane quarter
        int as = 0;
        uint32_t aph = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            mbar_wait(&tmem_full[as], aph);
            tcgen05_fence_after();
            const int row = m_blk * BM + ew * 32 + lane;
            const bool row_ok = row < p.M;
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
            const int n0 = n_blk * BN;

            if constexpr (EPI == EPI_PLAIN || EPI == EPI_RESID) {
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32b_x32(tbase + c * 32, v);
                    tmem_ld_wait();
                    const int col0 = n0 + c * 32;
                    const int nvalid = p.N - col0;
                    if (row_ok && nvalid > 0) {
                        uint32_t pk[16];
                        if constexpr (EPI == EPI_RESID) {
                            const uint4* r4 = reinterpret_cast<const uint4*>(p.resid + (size_t)row * p.ldr + col0);
                            uint32_t rr[16];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                uint4 t = (i * 8 < nvalid) ? r4[i] : make_uint4(0, 0, 0, 0);
                                rr[4 * i] = t.x; rr[4 * i + 1] = t.y; rr[4 * i + 2] = t.z; rr[4 * i + 3] = t.w;
                            }
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                // nn.Linear output is rounded to bf16 first, then the residual add rounds again
                                float a0 = bf16_round(__uint_as_float(v[2 * i]));
                                float a1 = bf16_round(__uint_as_float(v[2 * i + 1]));
                                pk[i] = pack_bf16x2(__fadd_rn(bf16_lo(rr[i]), a0), __fadd_rn(bf16_hi(rr[i]), a1));
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                pk[i] = pack_bf16x2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
                        }
                        store_bf16x32(p.C + (size_t)row * p.ldc + col0, pk, nvalid);
                    }
                }
            } else if constexpr (EPI == EPI_SWIGLU) {
                // tile columns [0,128) = gate rows of W1, [128,256) = up rows of W3 (weights packed interleaved)
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(tbase + c * 32, g);
                    tmem_ld_32x32b_x32(tbase + 128 + c * 32, u);
                    tmem_ld_wait();
                    const int col0 = n_blk * 128 + c * 32;
                    const int nvalid = p.N / 2 - col0;
                    if (row_ok && nvalid > 0) {
                        uint32_t pk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            float o[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float gg = bf16_round(__uint_as_float(g[2 * i + h]));
                                float uu = bf16_round(__uint_as_float(u[2 * i + h]));
                                float s = bf16_round(__fdiv_rn(gg, __fadd_rn(1.0f, expf(-gg))));  // silu -> bf16
                                o[h] = __fmul_rn(s, uu);
                            }
                            pk[i] = pack_bf16x2(o[0], o[1]);
                        }
                        store_bf16x32(p.C + (size_t)row * p.ldc + col0, pk, nvalid);
                    }
                }
            } else if constexpr (EPI == EPI_QKVROPE) {
                const int region = n0 / p.d_model;  // 0 = Q, 1 = K, 2 = V (d_model % 256 == 0 is checked on the host)
                const int b = row_ok ? row / p.L : 0;
                const int pos = row_ok ? row - b * p.L : 0;
                if (region < 2) {
                    __nv_bfloat16* dst = (region == 0 ? p.q : p.k) + (size_t)row * p.d_model + (n0 - region * p.d_model);
#pragma unroll 1
                    for (int hc = 0; hc < 4; ++hc) {  // (head in tile) x (32-col chunk of the first half)
                        const int head = hc >> 1, cc = hc & 1;
                        uint32_t x1[32], x2[32];
                        tmem_ld_32x32b_x32(tbase + head * 128 + cc * 32, x1);
                        tmem_ld_32x32b_x32(tbase + head * 128 + 64 + cc * 32, x2);
                        tmem_ld_wait();
                        if (row_ok) {
                            const float4* c4 = reinterpret_cast<const float4*>(p.cos_tab + (size_t)pos * 64 + cc * 32);
                            const float4* s4 = reinterpret_cast<const float4*>(p.sin_tab + (size_t)pos * 64 + cc * 32);
                            uint32_t o1[16], o2[16];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 cv = c4[i], sv = s4[i];
                                const float cs[4] = {cv.x, cv.y, cv.z, cv.w};
                                const float sn[4] = {sv.x, sv.y, sv.z, sv.w};
                                float a[4], bb[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float t1 = bf16_round(__uint_as_float(x1[4 * i + j]));
                                    const float t2 = bf16_round(__uint_as_float(x2[4 * i + j]));
                                    // (t * cos) + (rotate_half(t) * sin), fp32, no FMA contraction
                                    a[j] = __fadd_rn(__fmul_rn(t1, cs[j]), __fmul_rn(-t2, sn[j]));
                                    bb[j] = __fadd_rn(__fmul_rn(t2, cs[j]), __fmul_rn(t1, sn[j]));
                                }
                                o1[2 * i] = pack_bf16x2(a[0], a[1]);
                                o1[2 * i + 1] = pack_bf16x2(a[2], a[3]);
                                o2[2 * i] = pack_bf16x2(bb[0], bb[1]);
                                o2[2 * i + 1] = pack_bf16x2(bb[2], bb[3]);
                            }
                            store_bf16x32(dst + head * 128 + cc * 32, o1, 32);
                            store_bf16x32(dst + head * 128 + 64 + cc * 32, o2, 32);
                        }
                    }
                } else {
                    // V is written transposed: vt[b][head][d][token] so that P·V runs with both operands K-major
#pragma unroll 1
                    for (int c = 0; c < BN / 32; ++c) {
                        uint32_t v[32];
                        tmem_ld_32x32b_x32(tbase + c * 32, v);
                        tmem_ld_wait();
                        if (row_ok) {
                            const int n = n0 - 2 * p.d_model + c * 32;
                            const int head = n >> 7, d0 = n & 127;
                            __nv_bfloat16* dst = p.vt + ((size_t)(b * p.n_heads + head) * 128 + d0) * p.Lpad + pos;
#pragma unroll
                            for (int i = 0; i < 32; ++i) dst[(size_t)i * p.Lpad] = __float2bfloat16_rn(__uint_as_float(v[i]));
                        }
                    }
                }
            }
            // all TMEM reads of this accumulator stage are complete (wait::ld above) -> hand it back to the MMA warp
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[as]);
            if (++as == 2) { as = 0; aph ^= 1; }
        }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
template <int EPI, int BN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        MMDP_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<BN>::kSmem));
        attr_set = true;
    }
    const int num_tiles = ((p.M + BM - 1) / BM) * ((p.N + BN - 1) / BN);
    const int grid = num_tiles < num_sms() ? num_tiles : num_sms();
    LaunchScope ls(LK_GEMM, 2.0 * p.M * (double)p.N * p.K, stream);
    gemm_bf16_kernel<EPI, BN><<<grid, kGemmThreads, GemmCfg<BN>::kSmem, stream>>>(tmA, tmB, p);
    MMDP_CUDA(cudaGetLastError());
    return 0;
}

// waves x tile width (plus a small penalty for the narrower tile's lower operand reuse)
static int pick_tile_n(int M, int N) {
    const int g = num_sms();
    const long long m_tiles = (M + BM - 1) / BM;
    const long long w256 = (m_tiles * ((N + 255) / 256) + g - 1) / g * 256 * 100;
    const long long w192 = (m_tiles * ((N + 191) / 192) + g - 1) / g * 192 * 104;
    return w192 < w256 ? 192 : 256;
}

int gemm_bf16(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
              __nv_bfloat16* C, int ldc, const __nv_bfloat16* resid, int ldr, const QkvRopeArgs* qa,
              cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0) return set_error("gemm: empty problem");
    if ((lda % 8) || (ldw % 8) || (K % 8)) return set_error("gemm: lda/ldw/K must be multiples of 8 (16-byte TMA strides)");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
        return set_error("gemm: A/W must be 16-byte aligned");
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc; p.resid = resid; p.ldr = ldr;
    const int bn = (epi == EPI_PLAIN || epi == EPI_RESID) ? pick_tile_n(M, N) : 256;
    CUtensorMap tmA, tmB;
    if (make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return -1;
    if (make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, bn, BK)) return -1;
    switch (epi) {
        case EPI_PLAIN:
            if (!C || (ldc % 8) || (N % 8)) return set_error("gemm: C null or ldc/N not multiple of 8");
            return bn == 192 ? launch_gemm<EPI_PLAIN, 192>(tmA, tmB, p, stream) : launch_gemm<EPI_PLAIN, 256>(tmA, tmB, p, stream);
        case EPI_RESID:
            if (!C || !resid || (ldc % 8) || (ldr % 8) || (N % 8)) return set_error("gemm: bad residual epilogue args");
            return bn == 192 ? launch_gemm<EPI_RESID, 192>(tmA, tmB, p, stream) : launch_gemm<EPI_RESID, 256>(tmA, tmB, p, stream);
        case EPI_SWIGLU:
            if (!C || (ldc % 8) || (N % 256)) return set_error("gemm: swiglu needs N % 256 == 0 (interleaved gate/up tiles)");
            return launch_gemm<EPI_SWIGLU, 256>(tmA, tmB, p, stream);
        case EPI_QKVROPE:
            if (!qa) return set_error("gemm: qkv epilogue needs QkvRopeArgs");
            if (qa->d_model % 256 || N != 3 * qa->d_model || qa->d_model != qa->n_heads * 128)
                return set_error("gemm: qkv epilogue needs head_dim 128, d_model % 256 == 0, N == 3*d_model");
            if (M % qa->L) return set_error("gemm: qkv epilogue needs M == B*L");
            p.q = qa->q; p.k = qa->k; p.vt = qa->vt; p.cos_tab = qa->cos_tab; p.sin_tab = qa->sin_tab;
            p.L = qa->L; p.Lpad = qa->Lpad; p.d_model = qa->d_model; p.n_heads = qa->n_heads;
            return launch_gemm<EPI_QKVROPE, 256>(tmA, tmB, p, stream);
        default:
            return set_error("gemm: unknown epilogue");
    }
}

}  // namespace mmdp
