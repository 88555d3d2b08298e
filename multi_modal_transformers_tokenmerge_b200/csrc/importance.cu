// K13: token importance from the attention weights.   attention_blocks/compressed_attention.py:303-306
//
//   importance_scores = mean over heads ( mean over the LAST axis ( attn_weights [B, H, Tq, Tk] ) )      -> [B, T]
//
// As written that is the mean over KEYS of rows that each sum to one, i.e. 1 / T for every token: mode
// TOME_IMPORTANCE_ROW_MEAN computes exactly that expression.  TOME_IMPORTANCE_RECEIVED is the same double mean with the
// inner one over QUERIES -- the attention a token receives, the quantity top-k pruning (token_compression.py:15-46) can
// rank on.  Both are sums of P[q, k] = exp(s[q, k] + log size[k] - lse[q]) with the row statistics lse the attention
// forward kernel already saved, so no softmax pass is repeated: every (q, k) element is independent.
//
// The contraction s = q . k is the one place here that is a matrix product, and it is small and feeds a transcendental
// epilogue: 16-row tiles per warp on the warp-level bf16 MMA (mma.sync m16n8k16, fp32 accumulate -- the same operand
// rounding as the attention kernels' S), A fragments of the warp's 16 rows held in registers across the sweep over the
// other axis, B fragments read straight from global memory (L1 / L2 resident: every warp of a CTA sweeps the same rows).
// Sums run in a fixed order (thread-sequential over heads and column tiles, then a 4-lane butterfly) => deterministic.
// Weights are the UNDROPPED ones (the ranking must not depend on the dropout stream).
#include <float.h>

#include "common.cuh"
#include "host_util.h"

namespace tome {

constexpr int IMP_WARPS = 4;

struct ImpParams {
  int batch, tokens, heads;
  float scale_log2;
  const __nv_bfloat16 *q, *k;
  long long q_bs, q_ts, k_bs, k_ts;
  const float* lse;     // [B, H, T] natural log (attention forward)
  const float* size;    // [B, T] or null
  const uint8_t* gid; const int32_t* pos; const uint8_t* allow; int num_groups;
  float* out;           // [B, T]
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ROWS_ARE_KEYS: the warp's 16 rows are keys and it sweeps the queries (RECEIVED); otherwise rows are queries, sweep over keys.
template <int D, bool ROWS_ARE_KEYS>
__global__ void __launch_bounds__(IMP_WARPS * 32)
attn_importance_kernel(const ImpParams p) {
  pdl_prologue();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, T = p.tokens;
  const int r0 = (blockIdx.x * IMP_WARPS + warp) * 16;
  if (r0 >= T) return;
  const __nv_bfloat16* X = (ROWS_ARE_KEYS ? p.k + b * p.k_bs : p.q + b * p.q_bs);
  const __nv_bfloat16* Y = (ROWS_ARE_KEYS ? p.q + b * p.q_bs : p.k + b * p.k_bs);
  const long long x_ts = ROWS_ARE_KEYS ? p.k_ts : p.q_ts, y_ts = ROWS_ARE_KEYS ? p.q_ts : p.k_ts;
  const int row[2] = {r0 + g, r0 + g + 8};
  const bool row_ok[2] = {row[0] < T, row[1] < T};
  const long long bt = (long long)b * T;
  int rg[2] = {0, 0}, rp[2] = {0, 0};
  float r_lsz[2] = {0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if (row_ok[i]) {
      if (p.gid) { rg[i] = p.gid[bt + row[i]]; rp[i] = p.pos[bt + row[i]]; }
      if (ROWS_ARE_KEYS && p.size) r_lsz[i] = log2f(p.size[bt + row[i]]);
    }
  float acc[2] = {0.f, 0.f};
  for (int h = 0; h < p.heads; ++h) {
    uint32_t a[D / 16][4];
#pragma unroll
    for (int kk = 0; kk < D / 16; ++kk) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ri = i & 1;
        a[kk][i] = row_ok[ri] ? *reinterpret_cast<const uint32_t*>(X + (long long)row[ri] * x_ts + h * D + kk * 16 + (i >> 1) * 8 + 2 * t) : 0u;
      }
    }
    const float* lse_h = p.lse + ((long long)b * p.heads + h) * T;
    float r_lse2[2] = {0.f, 0.f};
    if (!ROWS_ARE_KEYS) {
#pragma unroll
      for (int i = 0; i < 2; ++i) r_lse2[i] = row_ok[i] ? lse_h[row[i]] * 1.4426950408889634f : 0.f;
    }
    for (int c0 = 0; c0 < T; c0 += 8) {
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const int yrow = c0 + g;
      const __nv_bfloat16* yr = Y + (long long)(yrow < T ? yrow : T - 1) * y_ts + h * D + 2 * t;
#pragma unroll
      for (int kk = 0; kk < D / 16; ++kk) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(yr + kk * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(yr + kk * 16 + 8);
        mma_bf16_16816(c, a[kk], b0, b1);
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int col = c0 + 2 * t + j;
        if (col >= T) continue;
        int cg = 0, cp = 0;
        if (p.gid) { cg = p.gid[bt + col]; cp = p.pos[bt + col]; }
        const float c_term = ROWS_ARE_KEYS ? -lse_h[col] * 1.4426950408889634f : (p.size ? log2f(p.size[bt + col]) : 0.f);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (!row_ok[i]) continue;
          const int qg = ROWS_ARE_KEYS ? cg : rg[i], qp = ROWS_ARE_KEYS ? cp : rp[i];
          const int kg = ROWS_ARE_KEYS ? rg[i] : cg, kp = ROWS_ARE_KEYS ? rp[i] : cp;
          bool vis = true;
          if (p.gid) {
            const int al = p.allow[qg * p.num_groups + kg];
            vis = al == 1 || (al == 2 && kp <= qp);
          }
          const float r_term = ROWS_ARE_KEYS ? r_lsz[i] : -r_lse2[i];
          const float s2 = fmaf(c[2 * i + j], p.scale_log2, c_term + r_term);
          acc[i] += vis ? exp2f(s2) : 0.f;
        }
      }
    }
  }
  const float inv = 1.0f / ((float)p.heads * (float)T);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float v = acc[i];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    if (t == 0 && row_ok[i]) p.out[bt + row[i]] = v * inv;
  }
}

}  // namespace tome

using namespace tome;

extern "C" int tome_attention_importance(const tome_attn_desc_t* d, const void* q, const void* k, const float* lse, int mode,
                                         float* importance, void* stream_) {
  clear_error();
  cudaStream_t stream = (cudaStream_t)stream_;
  TOME_CHECK(d && q && k && lse && importance, TOME_ERR_INVALID, "attention_importance: null argument");
  TOME_CHECK(d->batch > 0 && d->batch <= 65535 && d->tokens > 0 && d->heads > 0, TOME_ERR_INVALID, "attention_importance: bad shape");
  TOME_CHECK(d->head_dim == 64 || d->head_dim == 128 || d->head_dim == 256, TOME_ERR_UNSUPPORTED,
             "attention_importance: head_dim %d not supported (64, 128, 256)", d->head_dim);
  TOME_CHECK(mode == TOME_IMPORTANCE_ROW_MEAN || mode == TOME_IMPORTANCE_RECEIVED, TOME_ERR_INVALID, "attention_importance: unknown mode %d", mode);
  TOME_CHECK((d->q_token_stride | d->k_token_stride | d->q_batch_stride | d->k_batch_stride) % 2 == 0 &&
             (((uintptr_t)q | (uintptr_t)k) & 3) == 0, TOME_ERR_INVALID, "attention_importance: q / k rows must be 4-byte aligned");
  TOME_CHECK(!d->gid || (d->pos && d->allow && d->num_groups > 0), TOME_ERR_INVALID, "attention_importance: gid needs pos, allow, num_groups");
  ImpParams p;
  p.batch = d->batch; p.tokens = d->tokens; p.heads = d->heads;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q); p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.q_bs = d->q_batch_stride; p.q_ts = d->q_token_stride; p.k_bs = d->k_batch_stride; p.k_ts = d->k_token_stride;
  p.lse = lse; p.size = d->size;
  p.gid = d->gid; p.pos = d->pos; p.allow = d->allow; p.num_groups = d->num_groups;
  p.out = importance;
  const dim3 grid(ceil_div(d->tokens, 16 * IMP_WARPS), d->batch);
  const double flops = 2.0 * d->batch * d->heads * (double)d->tokens * d->tokens * d->head_dim;
  ProfScope prof(PROF_IMPORTANCE, flops, 1, stream);
#define IMP_LAUNCH(DD)                                                                                          \
  if (mode == TOME_IMPORTANCE_RECEIVED) launch_k(attn_importance_kernel<DD, true>, grid, IMP_WARPS * 32, 0, stream, p); \
  else launch_k(attn_importance_kernel<DD, false>, grid, IMP_WARPS * 32, 0, stream, p)
  if (d->head_dim == 64) { IMP_LAUNCH(64); }
  else if (d->head_dim == 128) { IMP_LAUNCH(128); }
  else { IMP_LAUNCH(256); }
#undef IMP_LAUNCH
  TOME_CUDA(cudaGetLastError());
  return TOME_OK;
}
