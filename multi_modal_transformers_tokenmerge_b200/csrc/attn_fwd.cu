// K5: attention forward (flash style) on tcgen05 with the block-causal GROUP-TABLE mask and the proportional
// log(size) key bias.  Replaces flax dot_product_attention as reached from tome_attention.py:259-285 with the mask of
// token_sequencer.py:313-321 -- without ever materialising [B,H,T,T] logits, weights or booleans.
//
// One CTA = one (batch, head, 128-query tile); 2 CTAs per SM interleave (TMEM 256 columns each).
//   warp 4      TMA producer: Q once, then K_0 V_0 K_1 V_1 ... (128 keys x 64 each) through a 3-slot ring, 128B swizzle
//   warp 5      UMMA issuer:  S = Q K_j^T (128x128x64) -> TMEM cols [0,128);  O_j = P_j V_j (128x64x128) -> cols [128,192)
//   warps 0..3  softmax: thread = query row.  Pass 1 reads S from TMEM, applies scale + log-size bias + mask, takes
//               the row max and writes the finished logits back to TMEM; pass 2 re-reads them, exponentiates and
//               writes P_j to shared memory as the bf16 K-major A operand of the second MMA.  O is kept in fp32
//               REGISTERS and rescaled there (O = O * alpha + O_j), so TMEM never needs a correction pass.
//               The mask costs one bit test per element: per key tile the 128 threads publish, by warp ballots, one
//               128-bit "visible keys" word set per QUERY GROUP (the mask depends on a query only through its
//               group), and the log-size bias is fetched four keys per shared-memory read.
// Logits are kept in the log2 domain: s2 = (q.k) * scale * log2(e) + log2(size_k); masked -> -FLT_MAX (finite, as
// flax's finfo.min), keys past T -> -inf.
#include <float.h>

#include "common.cuh"
#include "host_util.h"

namespace tome {

constexpr int ATT_BM = 128;   // queries per CTA
constexpr int ATT_BN = 128;   // keys per tile
constexpr int ATT_D = 64;
constexpr int ATT_THREADS = 192;
constexpr int ATT_Q_BYTES = ATT_BM * ATT_D * 2;         // 16 KB
constexpr int ATT_KV_BYTES = ATT_BN * ATT_D * 2;        // 16 KB each for K and V
constexpr int ATT_P_BYTES = ATT_BM * ATT_BN * 2;        // 32 KB (two 64-key K-blocks of 16 KB)
constexpr int ATT_SLOTS = 3;  // ring of 16 KB slots; items are loaded in the order K_0 V_0 K_1 V_1 ...
constexpr int ATT_SMEM_META = 2 * ATT_BN * (4 + 4 + 4) + 2 * 32 * 4 * 4 * 2 + 2 * 32 * 4 + 16;  // bias2/pos/gid x 2 parities, visible-key words (all, causal) per group x 2 parities, column words
constexpr int ATT_SMEM = ATT_Q_BYTES + ATT_SLOTS * ATT_KV_BYTES + ATT_P_BYTES + ATT_SMEM_META + 256 + 1024;
constexpr uint32_t ATT_TMEM_COLS = 256;

struct AttnFwdParams {
  int batch, tokens, heads;
  float scale_log2;  // scale * log2(e)
  const uint8_t* gid;
  const int32_t* pos;
  const uint8_t* allow;
  int num_groups;
  const float* size;
  __nv_bfloat16* out;
  long long o_batch_stride, o_token_stride;
  float* lse;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                const __grid_constant__ CUtensorMap tm_v, const AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_kv = s_q + ATT_Q_BYTES;                        // slot i at i * 16 KB
  uint8_t* s_p = s_kv + ATT_SLOTS * ATT_KV_BYTES;
  float* s_bias = reinterpret_cast<float*>(s_p + ATT_P_BYTES);  // [2][128]
  int* s_pos = reinterpret_cast<int*>(s_bias + 2 * ATT_BN);     // [2][128]
  int* s_gid = s_pos + 2 * ATT_BN;                              // [2][128]
  uint32_t* s_vis = reinterpret_cast<uint32_t*>(s_gid + 2 * ATT_BN);  // [2][32 groups][4] keys visible to a query group
  uint32_t* s_visc = s_vis + 2 * 32 * 4;                        // [2][32][4] keys visible iff pos_k <= pos_q
  uint32_t* s_colw = s_visc + 2 * 32 * 4;                       // [32] query groups that see key group g (code 1)
  uint32_t* s_colc = s_colw + 32;                               // [32] ... (code 2)
  uint32_t* s_anyc = s_colc + 32;                               // [1]  some rule is causal
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_anyc + 4);
  uint64_t* q_full = bars;                // 1
  uint64_t* kv_full = bars + 1;           // [3]
  uint64_t* kv_empty = bars + 4;          // [3]
  uint64_t* s_full = bars + 7;            // 1   S_j ready in TMEM
  uint64_t* p_ready = bars + 8;           // 1   P_j in smem, S_j and O_{j-1} consumed (128 arrivals)
  uint64_t* o_full = bars + 9;            // 1   O_j ready in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.tokens;
  const int n_kv = (T + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < ATT_SLOTS; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_ready, ATT_BM);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 5) tmem_alloc(tmem_slot, ATT_TMEM_COLS);
  if (threadIdx.x < 32) {  // transpose the allow table: which query groups may see keys of group g
    uint32_t cw = 0, cc = 0;
    const int g = threadIdx.x;
    if (p.gid != nullptr && g < p.num_groups)
      for (int qg = 0; qg < p.num_groups; ++qg) {
        const int a = p.allow[qg * p.num_groups + g];
        cw |= (a == 1 ? 1u : 0u) << qg;
        cc |= (a == 2 ? 1u : 0u) << qg;
      }
    s_colw[g] = cw;
    s_colc[g] = cc;
    const uint32_t any = __ballot_sync(0xffffffffu, cc != 0);
    if (g == 0) s_anyc[0] = any;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base;         // 128 columns
  const uint32_t tmem_o = tmem_base + 128;   // 64 columns

  if (warp == 4) {
    // ================================================================= TMA producer
    if (lane == 0) {
      mbar_expect_tx(q_full, ATT_Q_BYTES);
      tma_load_3d(s_q, &tm_q, q_full, h * ATT_D, qt * ATT_BM, b);
      for (int item = 0; item < 2 * n_kv; ++item) {
        const int slot = item % ATT_SLOTS, use = item / ATT_SLOTS;
        mbar_wait(&kv_empty[slot], (use & 1) ^ 1);
        mbar_expect_tx(&kv_full[slot], ATT_KV_BYTES);
        tma_load_3d(s_kv + slot * ATT_KV_BYTES, (item & 1) ? &tm_v : &tm_k, &kv_full[slot], h * ATT_D, (item >> 1) * ATT_BN, b);
      }
    }
  } else if (warp == 5) {
    // ================================================================= UMMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(ATT_BM, ATT_BN, false, false);  // Q K-major, K K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(ATT_BM, ATT_D, false, true);    // P K-major, V MN-major
      const uint32_t aq = smem_u32(s_q), ap = smem_u32(s_p);
      mbar_wait(q_full, 0);
      // issue order: S_0 | (wait P_0) S_1, O_0 | (wait P_1) S_2, O_1 | ...   S and O_tmp are single-buffered:
      // p_ready(j-1) certifies that S_{j-1} and O_{j-2} were consumed and P_{j-1} is in shared memory.
      for (int j = 0; j <= n_kv; ++j) {
        if (j >= 1) {
          mbar_wait(p_ready, (j - 1) & 1);
          tc_fence_after();
        }
        if (j < n_kv) {  // S_j = Q K_j^T
          const int item = 2 * j, slot = item % ATT_SLOTS;
          mbar_wait(&kv_full[slot], (item / ATT_SLOTS) & 1);
          tc_fence_after();
          const uint32_t ak = smem_u32(s_kv + slot * ATT_KV_BYTES);
#pragma unroll
          for (int k = 0; k < ATT_D / 16; ++k)
            umma_bf16(tmem_s, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc_s,
                      k > 0 ? 1u : 0u);
          umma_commit(s_full);
          umma_commit(&kv_empty[slot]);
        }
        if (j >= 1) {  // O_{j-1} = P_{j-1} V_{j-1}
          const int item = 2 * (j - 1) + 1, slot = item % ATT_SLOTS;
          mbar_wait(&kv_full[slot], (item / ATT_SLOTS) & 1);
          tc_fence_after();
          const uint32_t av = smem_u32(s_kv + slot * ATT_KV_BYTES);
#pragma unroll
          for (int k = 0; k < ATT_BN / 16; ++k)
            umma_bf16(tmem_o, make_smem_desc(ap + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024),
                      make_smem_desc(av + k * 2048, 8192, 1024), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(o_full);
          umma_commit(&kv_empty[slot]);
        }
      }
    }
  } else {
    // ================================================================= softmax warps (thread = query row)
    const int row = threadIdx.x;                 // 0..127 == TMEM lane
    const int q = qt * ATT_BM + row;
    const uint32_t lane_sel = ((uint32_t)(warp * 32)) << 16;
    const bool has_mask = p.gid != nullptr;
    const bool any_causal = has_mask && s_anyc[0] != 0;
    int gq = 0, pos_q = 0;
    if (has_mask && q < T) {
      gq = p.gid[(long long)b * T + q];
      pos_q = p.pos[(long long)b * T + q];
    }
    float o_acc[ATT_D];
#pragma unroll
    for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 1.f;

    for (int j = 0; j < n_kv; ++j) {
      const int par = j & 1;
      {  // per-key metadata of this tile: thread `row` owns key j*128 + row
        const int kk = j * ATT_BN + row;
        float bias2 = -INFINITY;           // keys past T: logit -inf (and "visible", so the mask keeps the -inf)
        uint32_t cw = 0xffffffffu, cc = 0u;
        int pk = 0;
        if (kk < T) {
          bias2 = p.size ? log2f(p.size[(long long)b * T + kk]) : 0.f;
          if (has_mask) {
            const int gk = p.gid[(long long)b * T + kk];
            cw = s_colw[gk];
            cc = s_colc[gk];
            pk = p.pos[(long long)b * T + kk];
          }
        }
        s_bias[par * ATT_BN + row] = bias2;
        if (has_mask) {
          for (int g = 0; g < p.num_groups; ++g) {
            const uint32_t w = __ballot_sync(0xffffffffu, (cw >> g) & 1u);
            if (lane == 0) s_vis[(par * 32 + g) * 4 + warp] = w;
          }
          if (any_causal) {
            s_pos[par * ATT_BN + row] = pk;
            for (int g = 0; g < p.num_groups; ++g) {
              const uint32_t w = __ballot_sync(0xffffffffu, (cc >> g) & 1u);
              if (lane == 0) s_visc[(par * 32 + g) * 4 + warp] = w;
            }
          }
        }
      }
      named_bar_sync(1, ATT_BM);
      const float4* bias4 = reinterpret_cast<const float4*>(s_bias + par * ATT_BN);
      uint4 vis = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu), visc = make_uint4(0u, 0u, 0u, 0u);
      if (has_mask) {
        vis = *reinterpret_cast<const uint4*>(s_vis + (par * 32 + gq) * 4);
        if (any_causal) visc = *reinterpret_cast<const uint4*>(s_visc + (par * 32 + gq) * 4);
      }
      const bool causal_row = (visc.x | visc.y | visc.z | visc.w) != 0u;

      mbar_wait(s_full, j & 1);
      tc_fence_after();

      // pass 1: finished logits (log2 domain) back to TMEM + row maximum
      float m_tile = -INFINITY;
#pragma unroll 1
      for (int cq = 0; cq < ATT_BN / 32; ++cq) {
        uint32_t v[32];
        tmem_ld_x32(tmem_s + lane_sel + cq * 32, v);
        uint32_t vw = cq == 0 ? vis.x : cq == 1 ? vis.y : cq == 2 ? vis.z : vis.w;
        if (causal_row) {  // rare (Text sets): fold the causal rule into the visibility word
          const uint32_t cwd = cq == 0 ? visc.x : cq == 1 ? visc.y : cq == 2 ? visc.z : visc.w;
          for (int i = 0; i < 32; ++i)
            if (((cwd >> i) & 1u) && s_pos[par * ATT_BN + cq * 32 + i] <= pos_q) vw |= 1u << i;
        }
        tmem_ld_wait();
#pragma unroll
        for (int i4 = 0; i4 < 8; ++i4) {
          const float4 bb = bias4[cq * 8 + i4];
          const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i4 * 4 + u;
            float s2 = fmaf(__uint_as_float(v[i]), p.scale_log2, bv[u]);
            if (has_mask) s2 = ((vw >> i) & 1u) ? s2 : -FLT_MAX;
            m_tile = fmaxf(m_tile, s2);
            v[i] = __float_as_uint(s2);
          }
        }
        tmem_st_x32(tmem_s + lane_sel + cq * 32, v);
      }
      tmem_st_wait();
      const float m_new = fmaxf(m_run, m_tile);  // finite: tile 0 always holds key 0
      const float alpha = fast_exp2(m_run - m_new);  // first tile: exp2(-inf) = 0

      // fold O_{j-1} (its MMA completed before S_j's commit fired) into the register accumulator
      if (j > 0) {
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_x32(tmem_o + lane_sel + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha_prev, __uint_as_float(v[i]));
        }
      }
      alpha_prev = alpha;

      // pass 2: P = exp2(s2 - m_new) -> bf16, K-major 128B-swizzled A operand in shared memory
      float l_tile = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem_s + lane_sel + c0, v);
        tmem_ld_wait();
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          pv[i] = fast_exp2(__uint_as_float(v[i]) - m_new);
          l_tile += pv[i];
        }
        uint8_t* prow = s_p + (c0 >> 6) * 16384 + row * 128;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {  // 4 chunks of 8 keys = 16 bytes
          const int chunk = ((c0 & 63) >> 3) + ch;
          const uint4 w = make_uint4(pack_bf16(pv[ch * 8 + 0], pv[ch * 8 + 1]), pack_bf16(pv[ch * 8 + 2], pv[ch * 8 + 3]),
                                     pack_bf16(pv[ch * 8 + 4], pv[ch * 8 + 5]), pack_bf16(pv[ch * 8 + 6], pv[ch * 8 + 7]));
          *reinterpret_cast<uint4*>(prow + ((chunk ^ (row & 7)) << 4)) = w;
        }
      }
      l_run = fmaf(l_run, alpha, l_tile);
      m_run = m_new;
      fence_proxy_async_smem();  // P visible to the tensor core (async proxy)
      tc_fence_before();         // our TMEM accesses of S_j / O_{j-1} are ordered before the MMAs that overwrite them
      mbar_arrive(p_ready);
    }
    // last tile's O
    mbar_wait(o_full, (n_kv - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int c0 = 0; c0 < ATT_D; c0 += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_o + lane_sel + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c0 + i] = fmaf(o_acc[c0 + i], alpha_prev, __uint_as_float(v[i]));
    }
    if (q < T) {
      const float inv_l = 1.0f / l_run;
      __nv_bfloat16* orow = p.out + (long long)b * p.o_batch_stride + (long long)q * p.o_token_stride + h * ATT_D;
#pragma unroll
      for (int i = 0; i < ATT_D; i += 8) {
        const uint4 w = make_uint4(pack_bf16(o_acc[i] * inv_l, o_acc[i + 1] * inv_l),
                                   pack_bf16(o_acc[i + 2] * inv_l, o_acc[i + 3] * inv_l),
                                   pack_bf16(o_acc[i + 4] * inv_l, o_acc[i + 5] * inv_l),
                                   pack_bf16(o_acc[i + 6] * inv_l, o_acc[i + 7] * inv_l));
        *reinterpret_cast<uint4*>(orow + i) = w;
      }
      if (p.lse) p.lse[((long long)b * p.heads + h) * T + q] = (m_run + log2f(l_run)) * 0.6931471805599453f;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, ATT_TMEM_COLS);
  }
}

}  // namespace tome

using namespace tome;

namespace tome {
int check_attn_desc(const tome_attn_desc_t* d, const char* who) {
  TOME_CHECK(d != nullptr, TOME_ERR_INVALID, "%s: null descriptor", who);
  TOME_CHECK(d->batch > 0 && d->tokens > 0 && d->heads > 0, TOME_ERR_INVALID, "%s: bad shape", who);
  TOME_CHECK(d->head_dim == ATT_D, TOME_ERR_UNSUPPORTED, "%s: head_dim %d not supported (this build: 64)", who, d->head_dim);
  TOME_CHECK(d->batch <= 65535 && d->heads <= 65535, TOME_ERR_INVALID, "%s: batch/heads exceed grid limits", who);
  TOME_CHECK(!d->gid || (d->pos && d->allow && d->num_groups >= 1 && d->num_groups <= 32), TOME_ERR_INVALID,
             "%s: a group mask needs pos, allow and 1 <= num_groups <= 32 (got %d)", who, d->num_groups);
  const long long strides[8] = {d->q_batch_stride, d->q_token_stride, d->k_batch_stride, d->k_token_stride,
                                d->v_batch_stride, d->v_token_stride, d->o_batch_stride, d->o_token_stride};
  for (int i = 0; i < 8; ++i)
    TOME_CHECK(strides[i] % 8 == 0, TOME_ERR_INVALID, "%s: strides must be multiples of 8 elements (16 bytes)", who);
  return TOME_OK;
}
}  // namespace tome

extern "C" int tome_attention_fwd(const tome_attn_desc_t* d, const void* q, const void* k, const void* v, void* out,
                                  float* lse, void* stream_) {
  clear_error();
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_attn_desc(d, "attention_fwd")) return rc;
  TOME_CHECK(q && k && v && out, TOME_ERR_INVALID, "attention_fwd: null tensor");
  TOME_CHECK(((uintptr_t)out & 15) == 0, TOME_ERR_INVALID, "attention_fwd: out must be 16-byte aligned");
  CUtensorMap tq, tk, tv;
  const uint64_t hd = (uint64_t)d->heads * d->head_dim;
  ProfScope prof(PROF_ATTN_FWD, 4.0 * d->batch * d->heads * (double)d->tokens * d->tokens * d->head_dim, 1, stream);
  if (int rc = make_tmap_3d_bf16(&tq, q, hd, d->tokens, d->batch, d->q_token_stride, d->q_batch_stride, ATT_BM)) return rc;
  if (int rc = make_tmap_3d_bf16(&tk, k, hd, d->tokens, d->batch, d->k_token_stride, d->k_batch_stride, ATT_BN)) return rc;
  if (int rc = make_tmap_3d_bf16(&tv, v, hd, d->tokens, d->batch, d->v_token_stride, d->v_batch_stride, ATT_BN)) return rc;
  AttnFwdParams p;
  p.batch = d->batch; p.tokens = d->tokens; p.heads = d->heads;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.gid = d->gid; p.pos = d->pos; p.allow = d->allow; p.num_groups = d->num_groups; p.size = d->size;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.o_batch_stride = d->o_batch_stride; p.o_token_stride = d->o_token_stride;
  p.lse = lse;
  static bool attr_set = false;
  if (!attr_set) {
    TOME_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM));
    attr_set = true;
  }
  dim3 grid(ceil_div(d->tokens, ATT_BM), d->heads, d->batch);
  attn_fwd_kernel<<<grid, ATT_THREADS, ATT_SMEM, stream>>>(tq, tk, tv, p);
  TOME_CUDA(cudaGetLastError());
  return TOME_OK;
}
