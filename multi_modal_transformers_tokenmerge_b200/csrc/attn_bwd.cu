// K6: attention backward on tcgen05 (autodiff of flax dot_product_attention as used at tome_attention.py:259-285,
// with the group-table mask and log(size) bias of the forward kernel).  No atomics, no fp32 dQ buffer: two
// kernels, each owning its outputs.
//
//   prep      delta[b,h,q] = sum_d dO*O ; per-token mask words (which groups a query may see)
//   dkdv      CTA = (128-key tile, head, batch), thread = key row, loop over 64-query tiles:
//               S^T = K Q^T, dP^T = V dO^T (TMEM) -> P^T, dS^T (bf16, shared) -> dV += P^T dO, dK += dS^T Q (TMEM)
//   dq        CTA = (128-query tile, head, batch), thread = query row, loop over 64-key tiles:
//               S = Q K^T, dP = dO V^T (TMEM) -> dS (bf16, shared) -> dQ += dS K (TMEM)
// P is recomputed from the saved log-sum-exp: P = exp2(s2 - lse2), s2 = q.k*scale*log2e + log2(size_k);
// dS = P * (dP - delta) * scale.  Each kernel uses 256 TMEM columns, so two CTAs share an SM.
#include <float.h>

#include "common.cuh"
#include "host_util.h"

namespace tome {

int check_attn_desc(const tome_attn_desc_t* d, const char* who);  // attn_fwd.cu

constexpr int AB_D = 64;
constexpr int AB_THREADS = 192;
constexpr uint32_t AB_TMEM_COLS = 256;

struct AttnBwdParams {
  int batch, tokens, heads;
  float scale, scale_log2;
  const uint8_t* gid;
  const int32_t* pos;
  const uint8_t* allow;
  int num_groups;
  const uint2* mwords;  // [B,T] (m_all, m_causal) or null
  const float* size;
  const float* lse;     // [B,H,T]
  const float* delta;   // [B,H,T]
  __nv_bfloat16* dq; long long dq_bs, dq_ts;
  __nv_bfloat16* dk; long long dk_bs, dk_ts;
  __nv_bfloat16* dv; long long dv_bs, dv_ts;
};

__device__ __forceinline__ void named_bar_sync_b(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------ prep
__global__ void attn_bwd_prep_kernel(int B, int T, int H, const __nv_bfloat16* __restrict__ o, long long o_bs, long long o_ts,
                                     const __nv_bfloat16* __restrict__ dout, long long do_bs, long long do_ts,
                                     float* __restrict__ delta, const uint8_t* __restrict__ gid,
                                     const uint8_t* __restrict__ allow, int G, uint2* __restrict__ mwords) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (b, t, h, chunk of 8)
  const long long total = (long long)B * T * H * 8;
  if (idx < total) {
    const int ch = (int)(idx & 7);
    const long long r = idx >> 3;
    const int h = (int)(r % H);
    const long long bt = r / H;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const uint4 ov = __ldg(reinterpret_cast<const uint4*>(o + b * o_bs + t * o_ts + h * AB_D + ch * 8));
    const uint4 dv = __ldg(reinterpret_cast<const uint4*>(dout + b * do_bs + t * do_ts + h * AB_D + ch * 8));
    float s = bf16_lo(ov.x) * bf16_lo(dv.x) + bf16_hi(ov.x) * bf16_hi(dv.x) + bf16_lo(ov.y) * bf16_lo(dv.y) +
              bf16_hi(ov.y) * bf16_hi(dv.y) + bf16_lo(ov.z) * bf16_lo(dv.z) + bf16_hi(ov.z) * bf16_hi(dv.z) +
              bf16_lo(ov.w) * bf16_lo(dv.w) + bf16_hi(ov.w) * bf16_hi(dv.w);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ch == 0) delta[((long long)b * H + h) * T + t] = s;
    if (mwords && h == 0 && ch == 0) {
      const int gq = gid[bt];
      uint32_t ma = 0, mc = 0;
      for (int g = 0; g < G; ++g) {
        const int a = allow[gq * G + g];
        ma |= (a == 1 ? 1u : 0u) << g;
        mc |= (a == 2 ? 1u : 0u) << g;
      }
      mwords[bt] = make_uint2(ma, mc);
    }
  }
}

// ------------------------------------------------------------------------------------------------ dK / dV
constexpr int DKV_BK = 128;  // keys per CTA
constexpr int DKV_BQ = 64;   // queries per tile
constexpr int DKV_KV_BYTES = DKV_BK * AB_D * 2;  // 16 KB
constexpr int DKV_Q_BYTES = DKV_BQ * AB_D * 2;   // 8 KB
constexpr int DKV_PT_BYTES = DKV_BK * DKV_BQ * 2;  // 16 KB
constexpr int DKV_META = 2 * DKV_BQ * 12 + 2 * 2 * 32 * 2 * 4;  // lse2, delta, pos x 2 parities; query-visibility words (all, causal) per key group x 2 parities
constexpr int DKV_SMEM = 2 * DKV_KV_BYTES + 2 * 2 * DKV_Q_BYTES + 2 * DKV_PT_BYTES + DKV_META + 256 + 1024;

__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                     const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                     const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_k = smem;
  uint8_t* s_v = s_k + DKV_KV_BYTES;
  uint8_t* s_qdo = s_v + DKV_KV_BYTES;            // stage s: Q at s*16K, dO at s*16K + 8K
  uint8_t* s_pt = s_qdo + 2 * 2 * DKV_Q_BYTES;    // P^T  [128 keys][64 queries] bf16, K-major swizzled
  uint8_t* s_dst = s_pt + DKV_PT_BYTES;           // dS^T
  float* s_lse = reinterpret_cast<float*>(s_dst + DKV_PT_BYTES);  // [2][64]
  float* s_delta = s_lse + 2 * DKV_BQ;
  int* s_posq = reinterpret_cast<int*>(s_delta + 2 * DKV_BQ);
  uint32_t* s_qvis = reinterpret_cast<uint32_t*>(s_posq + 2 * DKV_BQ);  // [2][32 key groups][2] queries that see the group
  uint32_t* s_qvisc = s_qvis + 2 * 32 * 2;                              // [2][32][2] ... iff pos_k <= pos_q
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_qvisc + 2 * 32 * 2);
  uint64_t* kv_full = bars;        // 1
  uint64_t* q_full = bars + 1;     // [2]
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* st_full = bars + 5;    // S^T_i and dP^T_i in TMEM
  uint64_t* ps_ready = bars + 6;   // P^T_i, dS^T_i in smem; TMEM S^T/dP^T consumed (128 arrivals)
  uint64_t* pd_free = bars + 7;    // dV/dK MMAs of tile i done: smem P^T/dS^T reusable, accumulators final at the end
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.tokens;
  const int n_q = (T + DKV_BQ - 1) / DKV_BQ;

  if (threadIdx.x == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(st_full, 1);
    mbar_init(ps_ready, DKV_BK);
    mbar_init(pd_free, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 5) tmem_alloc(tmem_slot, AB_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_st = tmem_base, tm_dpt = tmem_base + 64, tm_dk = tmem_base + 128, tm_dv = tmem_base + 192;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * DKV_KV_BYTES);
      tma_load_3d(s_k, &tm_k, kv_full, h * AB_D, kt * DKV_BK, b);
      tma_load_3d(s_v, &tm_v, kv_full, h * AB_D, kt * DKV_BK, b);
      for (int i = 0; i < n_q; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], 2 * DKV_Q_BYTES);
        tma_load_3d(s_qdo + st * 2 * DKV_Q_BYTES, &tm_q, &q_full[st], h * AB_D, i * DKV_BQ, b);
        tma_load_3d(s_qdo + st * 2 * DKV_Q_BYTES + DKV_Q_BYTES, &tm_do, &q_full[st], h * AB_D, i * DKV_BQ, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(DKV_BK, DKV_BQ, false, false);  // K/V K-major, Q/dO K-major
      constexpr uint32_t idesc_g = make_idesc_bf16(DKV_BK, AB_D, false, true);     // P^T/dS^T K-major, dO/Q MN-major
      const uint32_t ak = smem_u32(s_k), av = smem_u32(s_v), apt = smem_u32(s_pt), adst = smem_u32(s_dst);
      mbar_wait(kv_full, 0);
      for (int i = 0; i <= n_q; ++i) {
        if (i >= 1) {
          mbar_wait(ps_ready, (i - 1) & 1);
          tc_fence_after();
        }
        if (i < n_q) {
          const int st = i & 1;
          mbar_wait(&q_full[st], (i >> 1) & 1);
          tc_fence_after();
          const uint32_t aq = smem_u32(s_qdo + st * 2 * DKV_Q_BYTES), ado = aq + DKV_Q_BYTES;
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16(tm_st, make_smem_desc(ak + k * 32, 16, 1024), make_smem_desc(aq + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16(tm_dpt, make_smem_desc(av + k * 32, 16, 1024), make_smem_desc(ado + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(st_full);
        }
        if (i >= 1) {
          const int st = (i - 1) & 1;
          const uint32_t aq = smem_u32(s_qdo + st * 2 * DKV_Q_BYTES), ado = aq + DKV_Q_BYTES;
#pragma unroll
          for (int k = 0; k < DKV_BQ / 16; ++k)  // dV += P^T dO
            umma_bf16(tm_dv, make_smem_desc(apt + k * 32, 16, 1024), make_smem_desc(ado + k * 2048, 8192, 1024), idesc_g,
                      (i > 1 || k > 0) ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < DKV_BQ / 16; ++k)  // dK += dS^T Q
            umma_bf16(tm_dk, make_smem_desc(adst + k * 32, 16, 1024), make_smem_desc(aq + k * 2048, 8192, 1024), idesc_g,
                      (i > 1 || k > 0) ? 1u : 0u);
          umma_commit(pd_free);
          umma_commit(&q_empty[st]);
        }
      }
    }
  } else {
    const int row = threadIdx.x;  // key row == TMEM lane
    const int kk = kt * DKV_BK + row;
    const uint32_t lane_sel = ((uint32_t)(warp * 32)) << 16;
    const bool has_mask = p.mwords != nullptr;
    const bool k_valid = kk < T;
    float bias2 = 0.f;
    int gk = 0, pk = 0;
    if (k_valid) {
      if (p.size) bias2 = log2f(p.size[(long long)b * T + kk]);
      if (has_mask) {
        gk = p.gid[(long long)b * T + kk];
        pk = p.pos[(long long)b * T + kk];
      }
    }
    for (int i = 0; i < n_q; ++i) {
      const int par = i & 1;
      if (row < DKV_BQ) {  // per-query metadata of this tile (warps 0 and 1: one query per thread)
        const int q = i * DKV_BQ + row;
        float l2 = INFINITY, dl = 0.f;    // queries past T: exp2(s - inf) = 0
        uint32_t ma = 0xffffffffu, mc = 0;
        int pq = 0;
        if (q < T) {
          l2 = p.lse[((long long)b * p.heads + h) * T + q] * 1.4426950408889634f;
          dl = p.delta[((long long)b * p.heads + h) * T + q];
          if (has_mask) {
            const uint2 w = p.mwords[(long long)b * T + q];
            ma = w.x;
            mc = w.y;
            pq = p.pos[(long long)b * T + q];
          }
        }
        s_lse[par * DKV_BQ + row] = l2;
        s_delta[par * DKV_BQ + row] = dl;
        if (has_mask) {
          s_posq[par * DKV_BQ + row] = pq;
          for (int g = 0; g < p.num_groups; ++g) {  // bit q of word (g, warp): query q sees keys of group g
            const uint32_t wa = __ballot_sync(0xffffffffu, (ma >> g) & 1u);
            const uint32_t wc = __ballot_sync(0xffffffffu, (mc >> g) & 1u);
            if (lane == 0) {
              s_qvis[(par * 32 + g) * 2 + warp] = wa;
              s_qvisc[(par * 32 + g) * 2 + warp] = wc;
            }
          }
        }
      }
      named_bar_sync_b(1, DKV_BK);
      uint32_t vw[2] = {0xffffffffu, 0xffffffffu}, vc[2] = {0u, 0u};
      if (has_mask) {
        vw[0] = s_qvis[(par * 32 + gk) * 2];
        vw[1] = s_qvis[(par * 32 + gk) * 2 + 1];
        vc[0] = s_qvisc[(par * 32 + gk) * 2];
        vc[1] = s_qvisc[(par * 32 + gk) * 2 + 1];
      }
      if (!k_valid) vw[0] = vw[1] = vc[0] = vc[1] = 0u;  // rows past T contribute nothing
      const float4* lse4 = reinterpret_cast<const float4*>(s_lse + par * DKV_BQ);
      const float4* del4 = reinterpret_cast<const float4*>(s_delta + par * DKV_BQ);
      mbar_wait(st_full, i & 1);
      tc_fence_after();
      if (i >= 1) mbar_wait(pd_free, (i - 1) & 1);  // previous P^T / dS^T fully consumed by the tensor core
#pragma unroll
      for (int cq = 0; cq < DKV_BQ / 32; ++cq) {
        const int c0 = cq * 32;
        uint32_t sv[32], dv[32];
        tmem_ld_x32(tm_st + lane_sel + c0, sv);
        tmem_ld_x32(tm_dpt + lane_sel + c0, dv);
        uint32_t word = vw[cq];
        if (vc[cq]) {  // rare (Text sets): fold the causal rule into the visibility word
          for (int c = 0; c < 32; ++c)
            if (((vc[cq] >> c) & 1u) && pk <= s_posq[par * DKV_BQ + c0 + c]) word |= 1u << c;
        }
        tmem_ld_wait();
        float pv[32], ds[32];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 l4 = lse4[cq * 8 + c4], d4 = del4[cq * 8 + c4];
          const float lv[4] = {l4.x, l4.y, l4.z, l4.w}, dl4[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c4 * 4 + u;
            const float s2 = fmaf(__uint_as_float(sv[c]), p.scale_log2, bias2);
            const float pe = ((word >> c) & 1u) ? fast_exp2(s2 - lv[u]) : 0.f;
            pv[c] = pe;
            ds[c] = pe * p.scale * (__uint_as_float(dv[c]) - dl4[u]);
          }
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (c0 >> 3) + ch;
          const int off = row * 128 + ((chunk ^ (row & 7)) << 4);
          *reinterpret_cast<uint4*>(s_pt + off) =
              make_uint4(pack_bf16(pv[ch * 8 + 0], pv[ch * 8 + 1]), pack_bf16(pv[ch * 8 + 2], pv[ch * 8 + 3]),
                         pack_bf16(pv[ch * 8 + 4], pv[ch * 8 + 5]), pack_bf16(pv[ch * 8 + 6], pv[ch * 8 + 7]));
          *reinterpret_cast<uint4*>(s_dst + off) =
              make_uint4(pack_bf16(ds[ch * 8 + 0], ds[ch * 8 + 1]), pack_bf16(ds[ch * 8 + 2], ds[ch * 8 + 3]),
                         pack_bf16(ds[ch * 8 + 4], ds[ch * 8 + 5]), pack_bf16(ds[ch * 8 + 6], ds[ch * 8 + 7]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ps_ready);
    }
    mbar_wait(pd_free, (n_q - 1) & 1);  // accumulators final
    tc_fence_after();
    {
      // tcgen05.ld is warp-collective: rows past T take part in the loads and only skip the stores
      __nv_bfloat16* dkr = p.dk + (long long)b * p.dk_bs + (long long)(k_valid ? kk : 0) * p.dk_ts + h * AB_D;
      __nv_bfloat16* dvr = p.dv + (long long)b * p.dv_bs + (long long)(k_valid ? kk : 0) * p.dv_ts + h * AB_D;
#pragma unroll
      for (int c0 = 0; c0 < AB_D; c0 += 32) {
        uint32_t a[32], c[32];
        tmem_ld_x32(tm_dk + lane_sel + c0, a);
        tmem_ld_x32(tm_dv + lane_sel + c0, c);
        tmem_ld_wait();
        if (k_valid) {
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            *reinterpret_cast<uint4*>(dkr + c0 + i) =
                make_uint4(pack_bf16(__uint_as_float(a[i]), __uint_as_float(a[i + 1])), pack_bf16(__uint_as_float(a[i + 2]), __uint_as_float(a[i + 3])),
                           pack_bf16(__uint_as_float(a[i + 4]), __uint_as_float(a[i + 5])), pack_bf16(__uint_as_float(a[i + 6]), __uint_as_float(a[i + 7])));
            *reinterpret_cast<uint4*>(dvr + c0 + i) =
                make_uint4(pack_bf16(__uint_as_float(c[i]), __uint_as_float(c[i + 1])), pack_bf16(__uint_as_float(c[i + 2]), __uint_as_float(c[i + 3])),
                           pack_bf16(__uint_as_float(c[i + 4]), __uint_as_float(c[i + 5])), pack_bf16(__uint_as_float(c[i + 6]), __uint_as_float(c[i + 7])));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ dQ
constexpr int DQ_BQ = 128;  // queries per CTA
constexpr int DQ_BK = 64;   // keys per tile
constexpr int DQ_Q_BYTES = DQ_BQ * AB_D * 2;   // 16 KB
constexpr int DQ_K_BYTES = DQ_BK * AB_D * 2;   // 8 KB
constexpr int DQ_DS_BYTES = DQ_BQ * DQ_BK * 2;  // 16 KB
constexpr int DQ_META = 2 * DQ_BK * 8 + 2 * 2 * 32 * 2 * 4 + 2 * 32 * 4;  // bias2, pos x 2 parities; key-visibility words per query group; column words
constexpr int DQ_SMEM = 2 * DQ_Q_BYTES + 2 * 2 * DQ_K_BYTES + DQ_DS_BYTES + DQ_META + 256 + 1024;

__global__ void __launch_bounds__(AB_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                   const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_do,
                   const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* s_q = smem;
  uint8_t* s_do = s_q + DQ_Q_BYTES;
  uint8_t* s_kv = s_do + DQ_Q_BYTES;             // stage s: K at s*16K, V at s*16K + 8K
  uint8_t* s_ds = s_kv + 2 * 2 * DQ_K_BYTES;     // dS [128 queries][64 keys] bf16 K-major swizzled
  float* s_bias = reinterpret_cast<float*>(s_ds + DQ_DS_BYTES);  // [2][64]
  int* s_pos = reinterpret_cast<int*>(s_bias + 2 * DQ_BK);
  uint32_t* s_vis = reinterpret_cast<uint32_t*>(s_pos + 2 * DQ_BK);  // [2][32 query groups][2] keys visible to the group
  uint32_t* s_visc = s_vis + 2 * 32 * 2;                             // [2][32][2] ... iff pos_k <= pos_q
  uint32_t* s_colw = s_visc + 2 * 32 * 2;                            // [32] query groups that see key group g (code 1)
  uint32_t* s_colc = s_colw + 32;                                    // [32] (code 2)
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_colc + 32);
  uint64_t* q_full = bars;          // Q and dO
  uint64_t* kv_full = bars + 1;     // [2]
  uint64_t* kv_empty = bars + 3;    // [2]
  uint64_t* s_full = bars + 5;      // S_j, dP_j in TMEM
  uint64_t* ds_ready = bars + 6;    // dS_j in smem, S_j/dP_j consumed (128 arrivals)
  uint64_t* ds_free = bars + 7;     // dQ MMA of tile j done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int T = p.tokens;
  const int n_k = (T + DQ_BK - 1) / DQ_BK;

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(ds_ready, DQ_BQ);
    mbar_init(ds_free, 1);
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_do);
  }
  if (warp == 5) tmem_alloc(tmem_slot, AB_TMEM_COLS);
  if (threadIdx.x < 32) {  // transpose the mask words: which query groups may see keys of group g
    uint32_t cw = 0, cc = 0;
    const int g = threadIdx.x;
    if (p.allow != nullptr && g < p.num_groups)
      for (int qg = 0; qg < p.num_groups; ++qg) {
        const int a = p.allow[qg * p.num_groups + g];
        cw |= (a == 1 ? 1u : 0u) << qg;
        cc |= (a == 2 ? 1u : 0u) << qg;
      }
    s_colw[g] = cw;
    s_colc[g] = cc;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tm_s = tmem_base, tm_dp = tmem_base + 64, tm_dq = tmem_base + 128;

  if (warp == 4) {
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * DQ_Q_BYTES);
      tma_load_3d(s_q, &tm_q, q_full, h * AB_D, qt * DQ_BQ, b);
      tma_load_3d(s_do, &tm_do, q_full, h * AB_D, qt * DQ_BQ, b);
      for (int j = 0; j < n_k; ++j) {
        const int st = j & 1;
        mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        mbar_expect_tx(&kv_full[st], 2 * DQ_K_BYTES);
        tma_load_3d(s_kv + st * 2 * DQ_K_BYTES, &tm_k, &kv_full[st], h * AB_D, j * DQ_BK, b);
        tma_load_3d(s_kv + st * 2 * DQ_K_BYTES + DQ_K_BYTES, &tm_v, &kv_full[st], h * AB_D, j * DQ_BK, b);
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(DQ_BQ, DQ_BK, false, false);  // Q/dO K-major, K/V K-major
      constexpr uint32_t idesc_g = make_idesc_bf16(DQ_BQ, AB_D, false, true);    // dS K-major, K MN-major
      const uint32_t aq = smem_u32(s_q), ado = smem_u32(s_do), ads = smem_u32(s_ds);
      mbar_wait(q_full, 0);
      for (int j = 0; j <= n_k; ++j) {
        if (j >= 1) {
          mbar_wait(ds_ready, (j - 1) & 1);
          tc_fence_after();
        }
        if (j < n_k) {
          const int st = j & 1;
          mbar_wait(&kv_full[st], (j >> 1) & 1);
          tc_fence_after();
          const uint32_t ak = smem_u32(s_kv + st * 2 * DQ_K_BYTES), av = ak + DQ_K_BYTES;
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16(tm_s, make_smem_desc(aq + k * 32, 16, 1024), make_smem_desc(ak + k * 32, 16, 1024), idesc_s, k > 0);
#pragma unroll
          for (int k = 0; k < AB_D / 16; ++k)
            umma_bf16(tm_dp, make_smem_desc(ado + k * 32, 16, 1024), make_smem_desc(av + k * 32, 16, 1024), idesc_s, k > 0);
          umma_commit(s_full);
        }
        if (j >= 1) {
          const int st = (j - 1) & 1;
          const uint32_t ak = smem_u32(s_kv + st * 2 * DQ_K_BYTES);
#pragma unroll
          for (int k = 0; k < DQ_BK / 16; ++k)  // dQ += dS K
            umma_bf16(tm_dq, make_smem_desc(ads + k * 32, 16, 1024), make_smem_desc(ak + k * 2048, 8192, 1024), idesc_g,
                      (j > 1 || k > 0) ? 1u : 0u);
          umma_commit(ds_free);
          umma_commit(&kv_empty[st]);
        }
      }
    }
  } else {
    const int row = threadIdx.x;
    const int q = qt * DQ_BQ + row;
    const uint32_t lane_sel = ((uint32_t)(warp * 32)) << 16;
    const bool has_mask = p.mwords != nullptr;
    const bool q_valid = q < T;
    float lse2 = INFINITY, dl = 0.f;
    int gq = 0, pos_q = 0;
    if (q_valid) {
      lse2 = p.lse[((long long)b * p.heads + h) * T + q] * 1.4426950408889634f;
      dl = p.delta[((long long)b * p.heads + h) * T + q];
      if (has_mask) {
        gq = p.gid[(long long)b * T + q];
        pos_q = p.pos[(long long)b * T + q];
      }
    }
    for (int j = 0; j < n_k; ++j) {
      const int par = j & 1;
      if (row < DQ_BK) {  // per-key metadata (warps 0 and 1: one key per thread)
        const int kk = j * DQ_BK + row;
        float bias2 = -INFINITY;  // keys past T contribute nothing (and stay "visible" so the -inf survives)
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
        s_bias[par * DQ_BK + row] = bias2;
        if (has_mask) {
          s_pos[par * DQ_BK + row] = pk;
          for (int g = 0; g < p.num_groups; ++g) {
            const uint32_t wa = __ballot_sync(0xffffffffu, (cw >> g) & 1u);
            const uint32_t wc = __ballot_sync(0xffffffffu, (cc >> g) & 1u);
            if (lane == 0) {
              s_vis[(par * 32 + g) * 2 + warp] = wa;
              s_visc[(par * 32 + g) * 2 + warp] = wc;
            }
          }
        }
      }
      named_bar_sync_b(1, DQ_BQ);
      uint32_t vw[2] = {0xffffffffu, 0xffffffffu}, vc[2] = {0u, 0u};
      if (has_mask) {
        vw[0] = s_vis[(par * 32 + gq) * 2];
        vw[1] = s_vis[(par * 32 + gq) * 2 + 1];
        vc[0] = s_visc[(par * 32 + gq) * 2];
        vc[1] = s_visc[(par * 32 + gq) * 2 + 1];
      }
      const float4* bias4 = reinterpret_cast<const float4*>(s_bias + par * DQ_BK);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      if (j >= 1) mbar_wait(ds_free, (j - 1) & 1);
#pragma unroll
      for (int cq = 0; cq < DQ_BK / 32; ++cq) {
        const int c0 = cq * 32;
        uint32_t sv[32], dv[32];
        tmem_ld_x32(tm_s + lane_sel + c0, sv);
        tmem_ld_x32(tm_dp + lane_sel + c0, dv);
        uint32_t word = vw[cq];
        if (vc[cq]) {
          for (int c = 0; c < 32; ++c)
            if (((vc[cq] >> c) & 1u) && s_pos[par * DQ_BK + c0 + c] <= pos_q) word |= 1u << c;
        }
        tmem_ld_wait();
        float ds[32];
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 b4 = bias4[cq * 8 + c4];
          const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int c = c4 * 4 + u;
            const float s2 = fmaf(__uint_as_float(sv[c]), p.scale_log2, bv[u]);
            const float pe = ((word >> c) & 1u) ? fast_exp2(s2 - lse2) : 0.f;
            ds[c] = pe * p.scale * (__uint_as_float(dv[c]) - dl);
          }
        }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int chunk = (c0 >> 3) + ch;
          *reinterpret_cast<uint4*>(s_ds + row * 128 + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(pack_bf16(ds[ch * 8 + 0], ds[ch * 8 + 1]), pack_bf16(ds[ch * 8 + 2], ds[ch * 8 + 3]),
                         pack_bf16(ds[ch * 8 + 4], ds[ch * 8 + 5]), pack_bf16(ds[ch * 8 + 6], ds[ch * 8 + 7]));
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
    }
    mbar_wait(ds_free, (n_k - 1) & 1);
    tc_fence_after();
    __nv_bfloat16* dqr = p.dq + (long long)b * p.dq_bs + (long long)(q_valid ? q : 0) * p.dq_ts + h * AB_D;
#pragma unroll
    for (int c0 = 0; c0 < AB_D; c0 += 32) {
      uint32_t a[32];
      tmem_ld_x32(tm_dq + lane_sel + c0, a);
      tmem_ld_wait();
      if (q_valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8)
          *reinterpret_cast<uint4*>(dqr + c0 + i) =
              make_uint4(pack_bf16(__uint_as_float(a[i]), __uint_as_float(a[i + 1])), pack_bf16(__uint_as_float(a[i + 2]), __uint_as_float(a[i + 3])),
                         pack_bf16(__uint_as_float(a[i + 4]), __uint_as_float(a[i + 5])), pack_bf16(__uint_as_float(a[i + 6]), __uint_as_float(a[i + 7])));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem_base, AB_TMEM_COLS);
  }
}

}  // namespace tome

using namespace tome;

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

extern "C" size_t tome_attention_bwd_workspace_bytes(const tome_attn_desc_t* d) {
  if (!d || d->batch <= 0 || d->tokens <= 0 || d->heads <= 0) return 0;
  const size_t bht = (size_t)d->batch * d->heads * d->tokens;
  return align256(bht * sizeof(float)) + align256((size_t)d->batch * d->tokens * sizeof(uint2));
}

extern "C" int tome_attention_bwd(const tome_attn_desc_t* d, const tome_attn_grad_strides_t* gs, const void* q, const void* k,
                                  const void* v, const void* out, const float* lse, const void* dout, void* dq, void* dk,
                                  void* dv, void* workspace, size_t workspace_bytes, void* stream_) {
  clear_error();
  cudaStream_t stream = (cudaStream_t)stream_;
  if (int rc = check_attn_desc(d, "attention_bwd")) return rc;
  TOME_CHECK(gs && q && k && v && out && lse && dout && dq && dk && dv && workspace, TOME_ERR_INVALID,
             "attention_bwd: null argument");
  TOME_CHECK(((uintptr_t)workspace & 255) == 0, TOME_ERR_INVALID, "attention_bwd: workspace must be 256-byte aligned");
  TOME_CHECK(workspace_bytes >= tome_attention_bwd_workspace_bytes(d), TOME_ERR_INVALID,
             "attention_bwd: workspace too small (%zu < %zu, see tome_attention_bwd_workspace_bytes)", workspace_bytes,
             tome_attention_bwd_workspace_bytes(d));
  const long long st[8] = {gs->dq_batch_stride, gs->dq_token_stride, gs->dk_batch_stride, gs->dk_token_stride,
                           gs->dv_batch_stride, gs->dv_token_stride, gs->do_batch_stride, gs->do_token_stride};
  for (int i = 0; i < 8; ++i) TOME_CHECK(st[i] % 8 == 0, TOME_ERR_INVALID, "attention_bwd: strides must be multiples of 8");
  const int B = d->batch, T = d->tokens, H = d->heads;
  ProfScope prof(PROF_ATTN_BWD, 10.0 * d->batch * d->heads * (double)d->tokens * d->tokens * d->head_dim, 3, stream);
  float* delta = reinterpret_cast<float*>(workspace);
  uint2* mwords = d->gid ? reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(workspace) + align256((size_t)B * H * T * sizeof(float)))
                         : nullptr;
  {
    const long long total = (long long)B * T * H * 8;
    attn_bwd_prep_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        B, T, H, reinterpret_cast<const __nv_bfloat16*>(out), d->o_batch_stride, d->o_token_stride,
        reinterpret_cast<const __nv_bfloat16*>(dout), gs->do_batch_stride, gs->do_token_stride, delta, d->gid, d->allow,
        d->num_groups, mwords);
    TOME_CUDA(cudaGetLastError());
  }
  const uint64_t hd = (uint64_t)H * d->head_dim;
  AttnBwdParams p;
  p.batch = B; p.tokens = T; p.heads = H;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.gid = d->gid; p.pos = d->pos; p.allow = d->gid ? d->allow : nullptr; p.num_groups = d->num_groups; p.mwords = mwords; p.size = d->size; p.lse = lse; p.delta = delta;
  p.dq = reinterpret_cast<__nv_bfloat16*>(dq); p.dq_bs = gs->dq_batch_stride; p.dq_ts = gs->dq_token_stride;
  p.dk = reinterpret_cast<__nv_bfloat16*>(dk); p.dk_bs = gs->dk_batch_stride; p.dk_ts = gs->dk_token_stride;
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv); p.dv_bs = gs->dv_batch_stride; p.dv_ts = gs->dv_token_stride;
  static bool attr_set = false;
  if (!attr_set) {
    TOME_CUDA(cudaFuncSetAttribute(attn_bwd_dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DKV_SMEM));
    TOME_CUDA(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DQ_SMEM));
    attr_set = true;
  }
  {
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_tmap_3d_bf16(&tq, q, hd, T, B, d->q_token_stride, d->q_batch_stride, DKV_BQ)) return rc;
    if (int rc = make_tmap_3d_bf16(&tk, k, hd, T, B, d->k_token_stride, d->k_batch_stride, DKV_BK)) return rc;
    if (int rc = make_tmap_3d_bf16(&tv, v, hd, T, B, d->v_token_stride, d->v_batch_stride, DKV_BK)) return rc;
    if (int rc = make_tmap_3d_bf16(&tdo, dout, hd, T, B, gs->do_token_stride, gs->do_batch_stride, DKV_BQ)) return rc;
    dim3 grid(ceil_div(T, DKV_BK), H, B);
    attn_bwd_dkdv_kernel<<<grid, AB_THREADS, DKV_SMEM, stream>>>(tq, tk, tv, tdo, p);
    TOME_CUDA(cudaGetLastError());
  }
  {
    CUtensorMap tq, tk, tv, tdo;
    if (int rc = make_tmap_3d_bf16(&tq, q, hd, T, B, d->q_token_stride, d->q_batch_stride, DQ_BQ)) return rc;
    if (int rc = make_tmap_3d_bf16(&tk, k, hd, T, B, d->k_token_stride, d->k_batch_stride, DQ_BK)) return rc;
    if (int rc = make_tmap_3d_bf16(&tv, v, hd, T, B, d->v_token_stride, d->v_batch_stride, DQ_BK)) return rc;
    if (int rc = make_tmap_3d_bf16(&tdo, dout, hd, T, B, gs->do_token_stride, gs->do_batch_stride, DQ_BQ)) return rc;
    dim3 grid(ceil_div(T, DQ_BQ), H, B);
    attn_bwd_dq_kernel<<<grid, AB_THREADS, DQ_SMEM, stream>>>(tq, tk, tv, tdo, p);
    TOME_CUDA(cudaGetLastError());
  }
  return TOME_OK;
}
