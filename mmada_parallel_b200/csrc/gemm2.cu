// CTA-pair (cta_group::2) variant of the bf16 GEMM: two CTAs of a cluster (same TPC) cooperate on one 256 x BN tile.
// Each CTA stages its own 128 rows of A and HALF of the W tile (BN/2 rows); the leader CTA issues
// tcgen05.mma.cta_group::2 (UMMA M = 256) which reads A from both CTAs' shared memory and the two W halves, writing a
// 128 x BN fp32 accumulator into each CTA's TMEM. Per MMA each SM reads 4 KB (A) + BN*16 B (half of W) instead of
// 4 KB + BN*32 B, which lifts the shared-memory-bandwidth ceiling the 1-CTA kernel sits under (ncu: tensor pipe
// 77 % active in profiles/r01). The K loop order is the same as in gemm.cu, so results are bit-identical to it.
//
// Pipelines (per CTA unless noted):
//   full[s]   (leader only) : 2 arrivals (each CTA's producer) + tx bytes of BOTH CTAs' TMA loads (2-SM TMA signals
//                              the leader's barrier)
//   empty[s]                : released in both CTAs by the leader's multicast tcgen05.commit
//   tmem_full[a]            : multicast commit -> each CTA's epilogue warps
//   tmem_empty[a] (leader)  : 8 arrivals = 4 epilogue warps of each CTA (the peer arrives remotely)
#include "gemm_epilogue.cuh"

namespace mmdp {

static constexpr int P_BM = 128, P_BK = 64;
static constexpr int kPABytes = P_BM * P_BK * 2;  // 16 KB
static constexpr int kPairThreads = 256;
template <int BN> struct PairCfg {
    static constexpr int kBHalfBytes = (BN / 2) * P_BK * 2;
    static constexpr int kStageBytes = kPABytes + kBHalfBytes;  // per CTA
    static constexpr int kStages = (BN == 256) ? 6 : 7;
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256 + 4 * kScatStageFloats * 4;  // + reduce-scatter staging
    // epilogues whose 256-wide tile can be cut into two independent 128-wide halves (the rotary epilogue works per 128-column
    // head; the SwiGLU tile is [128 gate | 128 up] and cannot)
    template <int EPI> static constexpr bool can_split() { return BN == 256 && (EPI == EPI_PLAIN || EPI == EPI_RESID || EPI == EPI_QKVROPE); }
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same smem offset in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_u32(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // plain form (as CUTLASS' ClusterBarrier::arrive): an explicit .release.cluster here compiles to MEMBAR.ALL.GPU per
    // arrival, which throttled the peer's TMA producer to half speed (profiles/r01/README.md, pair-kernel note)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// 2-SM TMA load: data lands in THIS CTA's smem, completion bytes are signalled on the LEADER CTA's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tm, uint32_t leader_bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
// arrive (once all previously issued MMAs of this thread completed) on the barrier at this smem offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}

template <int EPI, int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmBh, const __grid_constant__ CUtensorMap tmAh, const GemmParams p) {
    constexpr int kStages = PairCfg<BN>::kStages;
    constexpr int kStageBytes = PairCfg<BN>::kStageBytes;
    static_assert(EPI == EPI_PLAIN || EPI == EPI_RESID || EPI == EPI_F32 || BN == 256, "fused QKV / SwiGLU epilogues need 256-wide tiles");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full = empty_bar + kStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    float* scat_stage = reinterpret_cast<float*>(smem + kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 2);   // one arrival per CTA of the pair (only the leader's copy is used)
            mbar_init(&empty_bar[s], 1);  // multicast commit
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], 8);  // 4 epilogue warps x 2 CTAs (only the leader's copy is used)
        }
        fence_barrier_init();
    }
    cluster_sync_all();  // barrier inits of both CTAs are visible before any remote arrive / 2-SM TMA
    if (warp == 2) tmem_alloc_pair<512>(tmem_ptr);
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // everything above overlaps the tail of the previous kernel in the stream (programmatic dependent launch)
    pdl_launch_dependents();
    pdl_wait();

    const int num_m = (p.M + 2 * P_BM - 1) / (2 * P_BM);  // pair tiles along M
    const int num_n = (p.N + BN - 1) / BN;
    const int num_k = (p.K + P_BK - 1) / P_BK;
    const int num_tiles = num_m * num_n;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    // Tail N-split (PairCfg::kCanSplit epilogues, p.nsplit_tail > 0): the tiles of the partial last wave are each cut into two
    // 256 x BN/2 units, so that the wave takes about half a tile time instead of a whole one. Units [0, t_full) are whole
    // tiles, unit t_full + 2 i + h is half h of tile t_full + i. A half unit runs the same K loop on its own columns:
    // results are bit-identical to the unsplit kernel, no exchange between units.
    constexpr bool kCanSplit = PairCfg<BN>::template can_split<EPI>();
    const int t_full = kCanSplit ? num_tiles - p.nsplit_tail : num_tiles;
    const int num_units = t_full + 2 * (num_tiles - t_full);
    // Half-M units (SwiGLU epilogue, p.mtail): the units of n-tile g are its num_m - 1 whole 256-row tiles plus ONE 128-row unit for
    // the last m-block (M % 256 <= 128), which the pair computes with M = 128 MMAs (64 rows per CTA) in half the cycles. The half
    // unit sits at position (g / mt_period) % num_m of the group so that unit % num_clusters spreads them over all clusters.
    constexpr bool kMTail = EPI == EPI_SWIGLU && BN == 256;
    auto decode = [&](int unit, int& m_blk, int& n_blk, bool& mhalf) {
        n_blk = unit / num_m;
        const int o = unit - n_blk * num_m;
        mhalf = false;
        m_blk = o;
        if (kMTail && p.mtail) {
            const int sp = (n_blk / p.mt_period) % num_m;
            mhalf = o == sp;
            m_blk = mhalf ? num_m - 1 : (o < sp ? o : o - 1);
        }
    };

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one_sync()) {
            int s = 0;
            uint32_t ph = 0;
            for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
                const bool half = kCanSplit && unit >= t_full;
                const int tile = half ? t_full + ((unit - t_full) >> 1) : unit;
                int m_blk, n_blk;
                bool mhalf;
                decode(tile, m_blk, n_blk, mhalf);
                // this CTA's rows of A: 128 of the 256-row tile, or 64 of the 128-row half-M unit
                const int row0 = mhalf ? m_blk * 2 * P_BM + (int)rank * (P_BM / 2) : m_blk * 2 * P_BM + (int)rank * P_BM;
                // this CTA's half of the W rows of the unit (BN rows, or BN/2 for a half-N unit)
                const int wrow0 = half ? n_blk * BN + ((unit - t_full) & 1) * (BN / 2) + (int)rank * (BN / 4) : n_blk * BN + (int)rank * (BN / 2);
                const uint32_t tx = half ? 2 * (kPABytes + PairCfg<BN>::kBHalfBytes / 2)
                                         : (mhalf ? 2 * (kPABytes / 2 + PairCfg<BN>::kBHalfBytes) : 2 * kStageBytes);  // bytes of both CTAs
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    const uint32_t leader_full = mapa_u32(smem_u32(&full_bar[s]), 0);
                    if (leader) {
                        mbar_expect_tx(&full_bar[s], tx);
                    } else {
                        mbar_arrive_cluster(leader_full);
                    }
                    uint8_t* sa = smem + s * kStageBytes;
                    if (kMTail && mhalf) {
                        // half-M unit: 64 rows of A; W rows so that the 128 accumulator columns one TMEM lane half holds are
                        // [64 gate | 64 up] of the SAME 64 outputs (CTA r: gate rows n_blk*256 + 64r.., up rows n_blk*256 + 128 + 64r..)
                        tma_load_2d_pair(sa, &tmAh, leader_full, kb * P_BK, row0);
                        tma_load_2d_pair(sa + kPABytes, &tmBh, leader_full, kb * P_BK, n_blk * BN + (int)rank * (BN / 4));
                        tma_load_2d_pair(sa + kPABytes + PairCfg<BN>::kBHalfBytes / 2, &tmBh, leader_full, kb * P_BK, n_blk * BN + BN / 2 + (int)rank * (BN / 4));
                    } else {
                        tma_load_2d_pair(sa, &tmA, leader_full, kb * P_BK, row0);
                        tma_load_2d_pair(sa + kPABytes, half ? &tmBh : &tmB, leader_full, kb * P_BK, wrow0);
                    }
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (leader && elect_one_sync()) {
            constexpr uint32_t idesc_whole = umma_idesc_bf16(2 * P_BM, BN);
            constexpr uint32_t idesc_half = umma_idesc_bf16(2 * P_BM, kCanSplit ? BN / 2 : BN);
            int s = 0;
            uint32_t ph = 0;
            int as = 0;
            uint32_t aph = 0;
            constexpr uint32_t idesc_mhalf = umma_idesc_bf16(P_BM, BN);  // cta_group::2, M = 128: 64 rows per CTA
            for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
                uint32_t idesc = (kCanSplit && unit >= t_full) ? idesc_half : idesc_whole;
                if (kMTail && p.mtail) {
                    int m_blk, n_blk;
                    bool mhalf;
                    decode(unit, m_blk, n_blk, mhalf);
                    if (mhalf) idesc = idesc_mhalf;
                }
                mbar_wait(&tmem_empty[as], aph ^ 1);
                tcgen05_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full_bar[s], ph);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_u32(smem + s * kStageBytes);
                    const uint64_t adesc = umma_desc_kmajor_sw128(sa);
                    const uint64_t bdesc = umma_desc_kmajor_sw128(sa + kPABytes);
#pragma unroll
                    for (int k = 0; k < P_BK / 16; ++k)
                        umma_bf16_ss_pair(d_tmem, adesc + (k * 2), bdesc + (k * 2), idesc, (kb | k) != 0);
                    umma_commit_pair(&empty_bar[s]);
                    if (kb == num_k - 1) umma_commit_pair(&tmem_full[as]);
                    if (++s == kStages) { s = 0; ph ^= 1; }
                }
                if (++as == 2) { as = 0; aph ^= 1; }
            }
        }
        __syncwarp();
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, each on its own 128 accumulator rows) =====================
        const int ew = warp - 4;
        int as = 0;
        uint32_t aph = 0;
        for (int unit = cluster_id; unit < num_units; unit += num_clusters) {
            const bool half = kCanSplit && unit >= t_full;
            const int tile = half ? t_full + ((unit - t_full) >> 1) : unit;
            int m_blk, n_blk;
            bool mhalf;
            decode(tile, m_blk, n_blk, mhalf);
            mbar_wait(&tmem_full[as], aph);
            tcgen05_fence_after();
            const uint32_t tbase = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + as * BN;
            if (kMTail && mhalf) {
                // M = 128 pair MMA: this CTA's 64 rows x 256 columns sit in TMEM as lanes 0..63 x columns [0,128) | lanes 64..127 x
                // columns [128,256), 128 TMEM columns. With the W rows ordered as above a lane half holds [64 gate | 64 up] of 64 outputs.
                const int ln = ew * 32 + lane, hq = ln >> 6;
                const int row = m_blk * 2 * P_BM + (int)rank * (P_BM / 2) + (ln & 63);
                const bool row_ok = row < p.M;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t g[32], u[32];
                    tmem_ld_32x32b_x32(tbase + c * 32, g);
                    tmem_ld_32x32b_x32(tbase + 64 + c * 32, u);
                    tmem_ld_wait();
                    swiglu_store32(p, g, u, row, row_ok, n_blk * 128 + hq * 64 + c * 32);
                }
            } else {
            const int row = m_blk * 2 * P_BM + (int)rank * P_BM + ew * 32 + lane;
            const bool row_ok = row < p.M;
            if constexpr (kCanSplit) {
                if (half)
                    gemm_epilogue_tile<EPI, BN / 2>(p, tbase, row, row_ok, 2 * n_blk + ((unit - t_full) & 1), scat_stage + ew * kScatStageFloats);
                else
                    gemm_epilogue_tile<EPI, BN>(p, tbase, row, row_ok, n_blk, scat_stage + ew * kScatStageFloats);
            } else {
                gemm_epilogue_tile<EPI, BN>(p, tbase, row, row_ok, n_blk, scat_stage + ew * kScatStageFloats);
            }
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&tmem_empty[as]), 0));
            if (++as == 2) { as = 0; aph ^= 1; }
        }
        if constexpr (EPI == EPI_F32) {
            if (p.scat_R > 0) __threadfence_system();  // the pushed rows are visible to their owners before this grid completes
        }
    }

    tcgen05_fence_before();
    cluster_sync_all();  // the peer's smem / TMEM stay alive until the leader's last MMA has retired
    if (warp == 2) {
        tcgen05_fence_after();
        tmem_dealloc_pair<512>(tmem_base);
    }
}

template <int EPI, int BN>
static int launch_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBh, const CUtensorMap& tmAh, GemmParams& p, cudaStream_t stream) {
    static unsigned long long attr_set = 0;  // bit per device
    int dev = 0;
    MMDP_CUDA(cudaGetDevice(&dev));
    if (!(attr_set >> (dev & 63) & 1ull)) {
        MMDP_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<EPI, BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, PairCfg<BN>::kSmem));
        attr_set |= 1ull << (dev & 63);
    }
    const int num_tiles = ((p.M + 2 * P_BM - 1) / (2 * P_BM)) * ((p.N + BN - 1) / BN);
    const int pairs = num_sms() / 2;
    const int grid = 2 * (num_tiles < pairs ? num_tiles : pairs);
    // tail N-split: a partial last wave of at most half the clusters, and a tile width that divides N (whole half tiles)
    const int tail = num_tiles % pairs;
    p.nsplit_tail = (PairCfg<BN>::template can_split<EPI>() && opt(OPT_GEMM_NSPLIT_TAIL) && num_tiles > pairs && tail > 0 &&
                     2 * tail <= pairs && p.N % BN == 0) ? tail : 0;
    // half-M units for the last m-block (SwiGLU): at most 128 valid rows in it, more than one m-block, whole 256-wide n-tiles
    const int num_m = (p.M + 2 * P_BM - 1) / (2 * P_BM), m_rem = p.M - (num_m - 1) * 2 * P_BM;
    p.mtail = (EPI == EPI_SWIGLU && BN == 256 && opt(OPT_GEMM_MTAIL) && num_m >= 2 && m_rem <= P_BM && p.N % BN == 0 && num_tiles > pairs) ? 1 : 0;
    if (p.mtail) {
        int a = num_m, b = pairs;  // period of (num_m * g) mod pairs
        while (b) { const int t = a % b; a = b; b = t; }
        p.mt_period = pairs / a;
    }
    LaunchScope ls(LK_GEMM, 2.0 * p.M * (double)p.N * p.K, stream);
    MMDP_CUDA(launch_ex(gemm_pair_kernel<EPI, BN>, dim3(grid), dim3(kPairThreads), PairCfg<BN>::kSmem, stream, pdl_mode() != 0, false, tmA, tmB,
                        (p.nsplit_tail || p.mtail) ? tmBh : tmB, p.mtail ? tmAh : tmA, p));
    return 0;
}

// waves x tile width for the pair kernel (74 clusters)
static int pick_pair_tile_n(int M, int N) {
    const int g = num_sms() / 2;
    const long long m_tiles = (M + 2 * P_BM - 1) / (2 * P_BM);
    const long long w256 = (m_tiles * ((N + 255) / 256) + g - 1) / g * 256 * 100;
    const long long w192 = (m_tiles * ((N + 191) / 192) + g - 1) / g * 192 * 102;
    return w192 < w256 ? 192 : 256;
}

int gemm_bf16_pair(int epi, const __nv_bfloat16* A, int lda, const __nv_bfloat16* W, int ldw, int M, int N, int K,
                   __nv_bfloat16* C, int ldc, const __nv_bfloat16* resid, int ldr, const QkvRopeArgs* qa, cudaStream_t stream,
                   const GemmScatter* sc) {
    GemmParams p{};
    p.M = M; p.N = N; p.K = K;
    p.C = C; p.ldc = ldc; p.resid = resid; p.ldr = ldr;
    if (sc) {
        for (int r = 0; r < 8; ++r) p.scat_dst[r] = sc->dst[r];
        p.scat_R = sc->rows_per_rank; p.scat_slot = sc->slot;
    }
    const int bn = (epi == EPI_PLAIN || epi == EPI_RESID || epi == EPI_F32) ? pick_pair_tile_n(M, N) : 256;
    CUtensorMap tmA, tmB;
    if (make_tmap_2d_bf16(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, P_BM, P_BK)) return -1;
    if (make_tmap_2d_bf16(&tmB, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, bn / 2, P_BK)) return -1;
    CUtensorMap tmBh = tmB;  // box of bn / 4 W rows: one CTA's share of a half unit (tail N-split)
    if (bn == 256 && epi != EPI_F32 && make_tmap_2d_bf16(&tmBh, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, bn / 4, P_BK)) return -1;
    CUtensorMap tmAh = tmA;  // box of 64 A rows: one CTA's share of a half-M unit
    if (epi == EPI_SWIGLU && make_tmap_2d_bf16(&tmAh, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, P_BM / 2, P_BK)) return -1;
    switch (epi) {
        case EPI_PLAIN:
            return bn == 192 ? launch_pair<EPI_PLAIN, 192>(tmA, tmB, tmBh, tmAh, p, stream) : launch_pair<EPI_PLAIN, 256>(tmA, tmB, tmBh, tmAh, p, stream);
        case EPI_RESID:
            return bn == 192 ? launch_pair<EPI_RESID, 192>(tmA, tmB, tmBh, tmAh, p, stream) : launch_pair<EPI_RESID, 256>(tmA, tmB, tmBh, tmAh, p, stream);
        case EPI_F32:
            return bn == 192 ? launch_pair<EPI_F32, 192>(tmA, tmB, tmBh, tmAh, p, stream) : launch_pair<EPI_F32, 256>(tmA, tmB, tmBh, tmAh, p, stream);
        case EPI_SWIGLU:
            return launch_pair<EPI_SWIGLU, 256>(tmA, tmB, tmBh, tmAh, p, stream);
        case EPI_QKVROPE:
            p.q = qa->q; p.k = qa->k; p.vt = qa->vt; p.cos_tab = qa->cos_tab; p.sin_tab = qa->sin_tab;
            p.L = qa->L; p.Lpad = qa->Lpad; p.d_model = qa->d_model; p.n_heads = qa->n_heads;
            p.pos_map = qa->pos_map; p.Tq = qa->Tq; p.row0 = qa->row0;
            return launch_pair<EPI_QKVROPE, 256>(tmA, tmB, tmBh, tmAh, p, stream);
        default:
            return set_error("gemm_pair: unknown epilogue");
    }
}

}  // namespace mmdp
