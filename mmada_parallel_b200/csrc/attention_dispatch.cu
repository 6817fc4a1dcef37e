// Attention forward entry point (q, k: [B*L, H*128] bf16 after RoPE; vt: [B*H*128, Lpad] bf16; out: [B*Lq, H*128] bf16).
// One kernel generation ships: attention6.cu (O and P in tensor memory, 64-wide KV blocks, S double-buffered, two CTAs per SM,
// KV-split of partial waves). Its predecessors (O in registers / P through shared memory, 563 TFLOP/s; 128-wide KV blocks,
// 688 TFLOP/s - history in DESIGN.md section 5) were kept selectable through round 1 and removed in round 2: they won on no shape.
#include "mmdp_internal.h"

namespace mmdp {

int attention_fwd_v6(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream, int Lq);

int attention_fwd_v7(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream, int Lq);

int attention_fwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                  int Lpad, float scale, cudaStream_t stream, int Lq) {
    if (opt(OPT_ATTN_VERSION) == 7) return attention_fwd_v7(q, k, vt, out, B, H, L, Lpad, scale, stream, Lq);
    return attention_fwd_v6(q, k, vt, out, B, H, L, Lpad, scale, stream, Lq);
}

}  // namespace mmdp
