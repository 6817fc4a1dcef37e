// HBM-bound row kernels of the transformer forward: embedding gather (wte lookup, modeling_llada.py:1265)
// and RMSLayerNorm (modeling_llada.py:315-329) with the reference's two bf16 rounding points:
//   y = bf16( w * bf16( x_fp32 * rsqrt(mean(x_fp32^2) + eps) ) )
#include "mmdp_internal.h"
#include "ptx.cuh"

#include <stdlib.h>

namespace mmdp {

__global__ void embed_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ wte,
                             __nv_bfloat16* __restrict__ x, int d, int64_t vocab, int* __restrict__ err) {
    const int row = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    int64_t id = ids[row];
    if (id < 0 || id >= vocab) {
        // torch raises IndexError here; the kernel reads row 0 instead and raises the context's sticky error flag, which
        // the host turns into the same exception at its next read-back (mmdp_model_error_flags)
        if (err && threadIdx.x == 0) atomicOr(err, 1);
        id = 0;
    }
    const uint4* src = reinterpret_cast<const uint4*>(wte + (size_t)id * d);
    uint4* dst = reinterpret_cast<uint4*>(x + (size_t)row * d);
    for (int i = threadIdx.x; i < d / 8; i += blockDim.x) dst[i] = src[i];
}

int embed_rows(const int64_t* ids, const __nv_bfloat16* wte, __nv_bfloat16* x, int M, int d, int64_t vocab,
               cudaStream_t stream, int* err) {
    if (d % 8) return set_error("embed: d must be a multiple of 8");
    LaunchScope ls(LK_ROW, 2.0 * M * (double)d * 2, stream);  // bytes: read row + write row
    MMDP_CUDA(launch_ex(embed_kernel, dim3(M), dim3(128), 0, stream, pdl_mode() != 0, false, ids, wte, x, d, vocab, err));
    return 0;
}

// One CTA per output row. rows == nullptr: input row = output row; else input row = rows[i] (gather).
__global__ void __launch_bounds__(256)
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const int* __restrict__ rows,
               const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y, int ldy, int d, float eps, int src_rows,
               int* __restrict__ err, int row_lo, int row_hi, int row_mod) {
    const int orow = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    int irow = rows ? rows[orow] : orow;
    if (rows && (irow < 0 || irow >= src_rows)) {  // gather index outside the source matrix: flag it, read row 0
        if (err && threadIdx.x == 0) atomicOr(err, 2);
        irow = 0;
    } else if (rows && ((irow % row_mod) < row_lo || (irow % row_mod) >= row_hi)) {  // a row the last block was not computed for (row window of every batch row)
        if (err && threadIdx.x == 0) atomicOr(err, 4);
    }
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)irow * ldx);
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    uint4* dst = reinterpret_cast<uint4*>(y + (size_t)orow * ldy);
    const int nvec = d / 8;

    // the row stays in registers when it fits (d <= 8192): one HBM/L2 read per element
    constexpr int kKeep = 4;
    const bool keep = nvec <= kKeep * 256;
    uint4 held[kKeep];
    float ss = 0.f;
#pragma unroll
    for (int t = 0; t < kKeep; ++t) {
        const int i = threadIdx.x + t * 256;
        held[t] = (keep && i < nvec) ? src[i] : make_uint4(0, 0, 0, 0);
    }
    if (keep) {
#pragma unroll
        for (int t = 0; t < kKeep; ++t) {
            const uint32_t u[4] = {held[t].x, held[t].y, held[t].z, held[t].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
                ss = fmaf(a, a, ss);
                ss = fmaf(b, b, ss);
            }
        }
    } else {
        for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
            const uint4 v = src[i];
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
                ss = fmaf(a, a, ss);
                ss = fmaf(b, b, ss);
            }
        }
    }
    __shared__ float red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    const float var = tot / (float)d;
    const float rstd = __frcp_rn(__fsqrt_rn(__fadd_rn(var, eps)));  // torch.rsqrt on CPU == 1/sqrt(x)

    for (int i = threadIdx.x, t = 0; i < nvec; i += blockDim.x, ++t) {
        uint4 v;
        if (keep) {
            v = held[0];
#pragma unroll
            for (int q = 1; q < kKeep; ++q)
                if (t == q) v = held[q];
        } else {
            v = src[i];
        }
        const uint4 wv = w4[i];
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = bf16_round(__fmul_rn(bf16_lo(u[j]), rstd));
            const float b = bf16_round(__fmul_rn(bf16_hi(u[j]), rstd));
            o[j] = pack_bf16x2(__fmul_rn(bf16_lo(ww[j]), a), __fmul_rn(bf16_hi(ww[j]), b));
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// One WARP per output row (8 rows per CTA) for d = 256 * NV <= 8192: the row is read once into registers with NV independent
// 16-byte loads per lane in flight, reduced with shuffles only (no block barrier), scaled and written. The activations are
// L2-resident between the producing GEMM and this kernel, so the kernel is latency- not HBM-bound: memory-level parallelism
// per thread is what counts (CTA-per-row version: 13.4 us for 2414 x 4096, ncu 68 % issue-active).
template <int NV>
__global__ void __launch_bounds__(256, 2)
rmsnorm_warp_kernel(const __nv_bfloat16* __restrict__ x, int ldx, const int* __restrict__ rows,
                    const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ y, int ldy, int M, float eps, int src_rows,
                    int* __restrict__ err, int row_lo, int row_hi, int row_mod) {
    const int orow = blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_launch_dependents();
    pdl_wait();
    if (orow >= M) return;
    const int lane = threadIdx.x & 31;
    int irow = rows ? rows[orow] : orow;
    if (rows && (irow < 0 || irow >= src_rows)) {  // gather index outside the source matrix: flag it, read row 0
        if (err && lane == 0) atomicOr(err, 2);
        irow = 0;
    } else if (rows && ((irow % row_mod) < row_lo || (irow % row_mod) >= row_hi)) {  // a row the last block was not computed for (row window of every batch row)
        if (err && lane == 0) atomicOr(err, 4);
    }
    const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)irow * ldx);
    const uint4* w4 = reinterpret_cast<const uint4*>(w);
    uint4* dst = reinterpret_cast<uint4*>(y + (size_t)orow * ldy);
    uint4 v[NV];
#pragma unroll
    for (int t = 0; t < NV; ++t) v[t] = src[lane + 32 * t];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < NV; ++t) {
        const uint32_t u[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = bf16_lo(u[j]), b = bf16_hi(u[j]);
            acc[j] = fmaf(a, a, acc[j]);
            acc[j] = fmaf(b, b, acc[j]);
        }
    }
    float ss = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float var = ss / (float)(NV * 256);
    const float rstd = __frcp_rn(__fsqrt_rn(__fadd_rn(var, eps)));  // torch.rsqrt on CPU == 1/sqrt(x)
    // weights in groups of (up to) four 16-byte loads; the "memory" clobber orders each group after the previous group's
    // stores, which keeps ptxas from hoisting all NV weight loads to the top (229 registers -> one CTA per SM)
    constexpr int G = NV < 4 ? NV : 4;
    // opaque touch of the packed row: without it the fp32 values unpacked for the sum of squares are kept alive for the
    // scaling pass (8 registers per 16-byte vector instead of 4)
#pragma unroll
    for (int t = 0; t < NV; ++t) asm volatile("" : "+r"(v[t].x), "+r"(v[t].y), "+r"(v[t].z), "+r"(v[t].w));
#pragma unroll
    for (int t0 = 0; t0 < NV; t0 += G) {
        uint4 wv[G];
#pragma unroll
        for (int g = 0; g < G; ++g)
            asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(wv[g].x), "=r"(wv[g].y), "=r"(wv[g].z), "=r"(wv[g].w)
                         : "l"(w4 + lane + 32 * (t0 + g))
                         : "memory");
#pragma unroll
        for (int g = 0; g < G; ++g) {
            const int t = t0 + g;
            const uint32_t u[4] = {v[t].x, v[t].y, v[t].z, v[t].w};
            const uint32_t ww[4] = {wv[g].x, wv[g].y, wv[g].z, wv[g].w};
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = bf16_round(__fmul_rn(bf16_lo(u[j]), rstd));
                const float b = bf16_round(__fmul_rn(bf16_hi(u[j]), rstd));
                o[j] = pack_bf16x2(__fmul_rn(bf16_lo(ww[j]), a), __fmul_rn(bf16_hi(ww[j]), b));
            }
            dst[lane + 32 * t] = make_uint4(o[0], o[1], o[2], o[3]);
        }
    }
}

int rmsnorm_rows(const __nv_bfloat16* x, int ldx, const int* rows, const __nv_bfloat16* w, __nv_bfloat16* y, int ldy,
                 int M, int d, float eps, cudaStream_t stream, int src_rows, int* err, int row_lo, int row_hi, int row_mod) {
    if (row_hi <= 0 || row_mod <= 0) { row_lo = 0; row_hi = 0x7fffffff; row_mod = 0x7fffffff; }
    if (M <= 0) return 0;
    if ((d % 8) || (ldx % 8) || (ldy % 8)) return set_error("rmsnorm: d/ldx/ldy must be multiples of 8");
    LaunchScope ls(LK_ROW, 2.0 * M * (double)d * 2, stream);  // bytes: read x + write y
    const int grid8 = (M + 7) / 8;
    const int use_warp = opt(OPT_RMSNORM_WARP);
    const bool pdl = pdl_mode() != 0;
    cudaError_t e;
    switch (use_warp ? d : -1) {  // warp-per-row variants for the model widths in use; anything else takes the CTA-per-row kernel
        case 4096: e = launch_ex(rmsnorm_warp_kernel<16>, dim3(grid8), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, M, eps, src_rows, err, row_lo, row_hi, row_mod); break;
        case 2048: e = launch_ex(rmsnorm_warp_kernel<8>, dim3(grid8), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, M, eps, src_rows, err, row_lo, row_hi, row_mod); break;
        case 1024: e = launch_ex(rmsnorm_warp_kernel<4>, dim3(grid8), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, M, eps, src_rows, err, row_lo, row_hi, row_mod); break;
        case 512: e = launch_ex(rmsnorm_warp_kernel<2>, dim3(grid8), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, M, eps, src_rows, err, row_lo, row_hi, row_mod); break;
        case 256: e = launch_ex(rmsnorm_warp_kernel<1>, dim3(grid8), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, M, eps, src_rows, err, row_lo, row_hi, row_mod); break;
        default: e = launch_ex(rmsnorm_kernel, dim3(M), dim3(256), 0, stream, pdl, false, x, ldx, rows, w, y, ldy, d, eps, src_rows, err, row_lo, row_hi, row_mod);
    }
    MMDP_CUDA(e);
    MMDP_CUDA(cudaGetLastError());
    return 0;
}

// x = bf16( bf16(partial_sum) + x ): the rounding points of `x + attn_out(att)` (modeling_llada.py:744,:953) applied to a
// tensor-parallel partial-sum buffer that has already been all-reduced in fp32.
__global__ void __launch_bounds__(256) resid_add_f32_kernel(__nv_bfloat16* __restrict__ x, int ldx, const float* __restrict__ part,
                                                             int ldp, int d) {
    const int row = blockIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    uint4* xr = reinterpret_cast<uint4*>(x + (size_t)row * ldx);
    const float4* pr = reinterpret_cast<const float4*>(part + (size_t)row * ldp);
    for (int i = threadIdx.x; i < d / 8; i += blockDim.x) {
        uint4 xv = xr[i];
        const float4 a = pr[2 * i], b = pr[2 * i + 1];
        const float pf[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        uint32_t u[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
            u[j] = pack_bf16x2(__fadd_rn(bf16_lo(u[j]), bf16_round(pf[2 * j])), __fadd_rn(bf16_hi(u[j]), bf16_round(pf[2 * j + 1])));
        xr[i] = make_uint4(u[0], u[1], u[2], u[3]);
    }
}

int resid_add_f32(__nv_bfloat16* x, int ldx, const float* partial, int ldp, int M, int d, cudaStream_t stream) {
    if (M <= 0) return 0;
    if ((d % 8) || (ldx % 8) || (ldp % 4)) return set_error("resid_add_f32: d/ldx must be multiples of 8, ldp of 4");
    LaunchScope ls(LK_ROW, (double)M * d * 8, stream);
    MMDP_CUDA(launch_ex(resid_add_f32_kernel, dim3(M), dim3(256), 0, stream, pdl_mode() != 0, false, x, ldx, partial, ldp, d));
    return 0;
}

int rmsnorm(const __nv_bfloat16* x, int ldx, const __nv_bfloat16* w, __nv_bfloat16* y, int ldy, int M, int d, float eps,
            cudaStream_t stream) {
    return rmsnorm_rows(x, ldx, nullptr, w, y, ldy, M, d, eps, stream, M, nullptr);
}

}  // namespace mmdp
