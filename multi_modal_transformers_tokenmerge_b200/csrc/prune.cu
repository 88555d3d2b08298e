// K9: per-modality top-k token pruning.   tokenizers/token_compression.py:15-46 (compute_top_k_tokens)
//
// For every token set (a contiguous slice [start, start + n) of the sequence) keep the k tokens with the largest
// importance score, in DESCENDING score order (jax.lax.top_k: equal scores keep the lower index first), concatenate
// the sets in the order given, and gather those rows of the embeddings.  The reference is written for one sequence
// and vmapped by its caller; here the batch is a grid dimension.
//
// One CTA per (token set, batch row): the set's scores go to shared memory, every token counts how many tokens of its
// set beat it (its rank; n <= a few hundred, so the n^2 comparisons are cheaper than a sort and exact by
// construction), ranks < k publish their index, and the CTA then copies the k selected rows with 128-bit accesses.
// Index arithmetic only: the gathered rows are bit-identical to the source rows.
#include "common.cuh"
#include "host_util.h"

namespace tome {

constexpr int PRUNE_THREADS = 256;

// "j ranks before i" in top_k order: larger score first, NaN above every number, equal scores by lower index
__device__ __forceinline__ bool topk_before(float vj, int j, float vi, int i) {
  const bool nj = vj != vj, ni = vi != vi;
  if (nj || ni) return nj && (!ni || j < i);
  return vj > vi || (vj == vi && j < i);
}

__global__ void __launch_bounds__(PRUNE_THREADS)
topk_prune_kernel(const tome_prune_desc_t d, const uint8_t* __restrict__ emb, const float* __restrict__ score,
                  uint8_t* __restrict__ out, int32_t* __restrict__ ids) {
  pdl_prologue();
  extern __shared__ float prune_sm[];
  const int s = blockIdx.x, b = blockIdx.y;
  const int start = d.set_start[s], n = d.set_n[s], k = d.set_k[s];
  int off = 0;
  for (int i = 0; i < s; ++i) off += d.set_k[i];
  int ktot = off;
  for (int i = s; i < d.n_sets; ++i) ktot += d.set_k[i];
  float* vals = prune_sm;                                   // [n]
  int* sel = reinterpret_cast<int*>(prune_sm + n);          // [k] token index of rank r
  // importance may arrive as `score_planes` partial planes [P][B][T] (e.g. one per attention head): summed in plane
  // order, so the ranking does not depend on how the caller reduced them
  for (int i = threadIdx.x; i < n; i += PRUNE_THREADS) {
    float v = 0.f;
    for (int p = 0; p < d.score_planes; ++p) v += score[((long long)p * d.batch + b) * d.tokens + start + i];
    vals[i] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += PRUNE_THREADS) {
    const float vi = vals[i];
    int rk = 0;
    for (int j = 0; j < n; ++j) rk += topk_before(vals[j], j, vi, i) ? 1 : 0;
    if (rk < k) {
      sel[rk] = start + i;
      ids[(long long)b * ktot + off + rk] = start + i;
    }
  }
  __syncthreads();
  const int row_bytes = d.channels * (d.dtype == TOME_BF16 ? 2 : 4);
  const int vpr = row_bytes >> 4;  // 16-byte vectors per row
  const uint8_t* eb = emb + (long long)b * d.tokens * row_bytes;
  uint8_t* ob = out + ((long long)b * ktot + off) * row_bytes;
  for (int i = threadIdx.x; i < k * vpr; i += PRUNE_THREADS) {
    const int r = i / vpr, v = i - r * vpr;
    st_na_v4(ob + (long long)r * row_bytes + v * 16, ld_nc_v4(eb + (long long)sel[r] * row_bytes + v * 16));
  }
}

}  // namespace tome

using namespace tome;

extern "C" int tome_topk_prune(const tome_prune_desc_t* d, const void* embeddings, const float* importance, void* out,
                               int32_t* ids, void* stream_) {
  clear_error();
  cudaStream_t stream = (cudaStream_t)stream_;
  TOME_CHECK(d && embeddings && importance && out && ids, TOME_ERR_INVALID, "topk_prune: null argument");
  TOME_CHECK(d->batch > 0 && d->batch <= 65535 && d->tokens > 0 && d->channels > 0, TOME_ERR_INVALID, "topk_prune: bad shape");
  TOME_CHECK(d->dtype == TOME_BF16 || d->dtype == TOME_F32, TOME_ERR_INVALID, "topk_prune: dtype must be bf16 or f32");
  TOME_CHECK((d->channels * (d->dtype == TOME_BF16 ? 2 : 4)) % 16 == 0, TOME_ERR_INVALID,
             "topk_prune: rows must be a multiple of 16 bytes (channels = %d)", d->channels);
  TOME_CHECK((((uintptr_t)embeddings | (uintptr_t)out) & 15) == 0, TOME_ERR_INVALID, "topk_prune: embeddings / out must be 16-byte aligned");
  TOME_CHECK(d->n_sets >= 1 && d->n_sets <= TOME_MAX_TOKEN_SETS, TOME_ERR_INVALID, "topk_prune: 1 <= n_sets <= %d (got %d)",
             TOME_MAX_TOKEN_SETS, d->n_sets);
  TOME_CHECK(d->score_planes >= 1, TOME_ERR_INVALID, "topk_prune: score_planes must be >= 1");
  int nmax = 0;
  for (int s = 0; s < d->n_sets; ++s) {
    TOME_CHECK(d->set_start[s] >= 0 && d->set_n[s] >= 1 && d->set_start[s] + d->set_n[s] <= d->tokens, TOME_ERR_INVALID,
               "topk_prune: token set %d = [%d, +%d) leaves the sequence of %d tokens", s, d->set_start[s], d->set_n[s], d->tokens);
    // jax.lax.top_k raises for k > n (token_compression.py:31)
    TOME_CHECK(d->set_k[s] >= 0 && d->set_k[s] <= d->set_n[s], TOME_ERR_INVALID,
               "topk_prune: token set %d keeps k = %d of %d tokens (top_k needs 0 <= k <= n)", s, d->set_k[s], d->set_n[s]);
    if (d->set_n[s] > nmax) nmax = d->set_n[s];
  }
  const size_t smem = (size_t)2 * nmax * sizeof(float);
  TOME_CHECK(smem <= 200 * 1024, TOME_ERR_UNSUPPORTED, "topk_prune: token set of %d tokens is too large for the shared-memory ranking", nmax);
  ProfScope prof(PROF_OTHER, 0.0, 1, stream);
  TOME_CUDA(cudaFuncSetAttribute(topk_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(d->n_sets, d->batch);
  launch_k(topk_prune_kernel, grid, PRUNE_THREADS, smem, stream, *d, reinterpret_cast<const uint8_t*>(embeddings), importance,
                                                          reinterpret_cast<uint8_t*>(out), ids);
  TOME_CUDA(cudaGetLastError());
  return TOME_OK;
}
