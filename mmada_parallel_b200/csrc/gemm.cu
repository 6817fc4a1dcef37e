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
#include "gemm_epilogue.cuh"

#include <stdlib.h>

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

template <int EPI, int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    constexpr int kStages = GemmCfg<BN>::kStages;
    constexpr int kStageBytes = GemmCfg<BN>::kStageBytes;
    static_assert(EPI == EPI_PLAIN || EPI == EPI_RESID || EPI == EPI_F32 || BN == 256, "fused QKV / SwiGLU epilogues need 256-wide tiles");
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
            gemm_epilogue_tile<EPI, BN>(p, tbase, row, row_ok, n_blk);
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

static int g_pair_mode = -1;
int gemm_pair_mode() {
    if (g_pair_mode < 0) {
        const char* e = getenv("MMDP_GEMM_PAIR");
        g_pair_mode = (e && e[0] == '1') ? 1 : 0;
    }
    return g_pair_mode;
}
void set_gemm_pair_mode(int on) { g_pair_mode = on ? 1 : 0; }

int gemm_bf16(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
              __nv_bfloat16* C, int ldc, const __nv_bfloat16* resid, int ldr, const QkvRopeArgs* qa,
              cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0) return set_error("gemm: empty problem");
    if ((lda % 8) || (ldw % 8) || (K % 8)) return set_error("gemm: lda/ldw/K must be multiples of 8 (16-byte TMA strides)");
    if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(W) & 15))
        return set_error("gemm: A/W must be 16-byte aligned");
    // argument validation common to both kernels
    switch (epi) {
        case EPI_PLAIN:
            if (!C || (ldc % 8) || (N % 8)) return set_error("gemm: C null or ldc/N not multiple of 8");
            break;
        case EPI_RESID:
            if (!C || !resid || (ldc % 8) || (ldr % 8) || (N % 8)) return set_error("gemm: bad residual epilogue args");
            break;
        case EPI_F32:
            if (!C || (ldc % 4) || (N % 4)) return set_error("gemm: fp32 output needs ldc/N multiples of 4");
            break;
        case EPI_SWIGLU:
            if (!C || (ldc % 8) || (N % 256)) return set_error("gemm: swiglu needs N % 256 == 0 (interleaved gate/up tiles)");
            break;
        case EPI_QKVROPE:
            if (!qa) return set_error("gemm: qkv epilogue needs QkvRopeArgs");
            if (qa->d_model % 256 || N != 3 * qa->d_model || qa->d_model != qa->n_heads * 128)
                return set_error("gemm: qkv epilogue needs head_dim 128, d_model % 256 == 0, N == 3*d_model");
            if (M % qa->L) return set_error("gemm: qkv epilogue needs M == B*L");
            break;
        default:
            return set_error("gemm: unknown epilogue");
    }
    // kernel selection: MMDP_GEMM_PAIR=1 routes large problems to the CTA-pair (cta_group::2) kernel of gemm2.cu
    if (gemm_pair_mode() && M > 256) return gemm_bf16_pair(epi, A, lda, W, ldw, M, N, K, C, ldc, resid, ldr, qa, stream);
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc; p.resid = resid; p.ldr = ldr;
    const int bn = (epi == EPI_PLAIN || epi == EPI_RESID || epi == EPI_F32) ? pick_tile_n(M, N) : 256;
    CUtensorMap tmA, tmB;
    if (make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM, BK)) return -1;
    if (make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, bn, BK)) return -1;
    switch (epi) {
        case EPI_PLAIN:
            return bn == 192 ? launch_gemm<EPI_PLAIN, 192>(tmA, tmB, p, stream) : launch_gemm<EPI_PLAIN, 256>(tmA, tmB, p, stream);
        case EPI_RESID:
            return bn == 192 ? launch_gemm<EPI_RESID, 192>(tmA, tmB, p, stream) : launch_gemm<EPI_RESID, 256>(tmA, tmB, p, stream);
        case EPI_F32:
            return bn == 192 ? launch_gemm<EPI_F32, 192>(tmA, tmB, p, stream) : launch_gemm<EPI_F32, 256>(tmA, tmB, p, stream);
        case EPI_SWIGLU:
            return launch_gemm<EPI_SWIGLU, 256>(tmA, tmB, p, stream);
        case EPI_QKVROPE:
            p.q = qa->q; p.k = qa->k; p.vt = qa->vt; p.cos_tab = qa->cos_tab; p.sin_tab = qa->sin_tab;
            p.L = qa->L; p.Lpad = qa->Lpad; p.d_model = qa->d_model; p.n_heads = qa->n_heads;
            return launch_gemm<EPI_QKVROPE, 256>(tmA, tmB, p, stream);
        default:
            return set_error("gemm: unknown epilogue");
    }
}

}  // namespace mmdp
