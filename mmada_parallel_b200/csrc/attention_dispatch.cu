// Attention forward: kernel selection. Three generations are kept, all with the same math, layouts and interface
// (q, k: [B*L, H*128] bf16 after RoPE; vt: [B*H*128, Lpad] bf16; out: [B*L, H*128] bf16):
//   6 (default)  attention6.cu  O and P in tensor memory, 64-wide KV blocks, S double-buffered, two CTAs per SM
//   5            attention5.cu  O and P in tensor memory, 128-wide KV blocks, strict QK -> softmax -> PV chain per CTA
//   3            attention.cu   O in registers, P through shared memory, one CTA per SM
// Measured on the full-size problem (B=1, L=2414, H=32): 697 / 672 / 563 TFLOP/s (profiles/r01/README.md).
// Select with mmdp_set_attention_version() or MMDP_ATTN=3|5|6.
#include "mmdp_internal.h"

#include <stdlib.h>

namespace mmdp {

int attention_fwd_v3(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream);
int attention_fwd_v6(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream, int Lq);
int attention_fwd_v5(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                     int Lpad, float scale, cudaStream_t stream);

static int g_attn_version = -1;
void set_attention_version(int v) { g_attn_version = v; }

int attention_fwd(const __nv_bfloat16* q, const __nv_bfloat16* k, const __nv_bfloat16* vt, __nv_bfloat16* out, int B, int H, int L,
                  int Lpad, float scale, cudaStream_t stream, int Lq) {
    if (Lq > 0 && Lq != L) return attention_fwd_v6(q, k, vt, out, B, H, L, Lpad, scale, stream, Lq);  // only v6 separates query and key lengths
    if (g_attn_version < 0) {
        const char* e = getenv("MMDP_ATTN");
        g_attn_version = (e && e[0] == '3') ? 3 : (e && e[0] == '5') ? 5 : 6;
    }
    if (g_attn_version == 3) return attention_fwd_v3(q, k, vt, out, B, H, L, Lpad, scale, stream);
    if (g_attn_version == 5) return attention_fwd_v5(q, k, vt, out, B, H, L, Lpad, scale, stream);
    return attention_fwd_v6(q, k, vt, out, B, H, L, Lpad, scale, stream, 0);
}

}  // namespace mmdp
