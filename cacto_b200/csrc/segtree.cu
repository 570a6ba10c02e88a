// segtree.cu -- K4: the prioritized-replay sum/min segment trees (segment_tree.py) and the sampled
// row gather (replay_buffer.py), bit-exact with the reference:
//   * every internal node is exactly left (+|min) right in fp64 -- the reference's fixed reduction order
//     (segment_tree.py:76-86) -- so a level-synchronous recomputation of the touched ancestors, after the
//     leaves are final, gives the same bits as the reference's sequential leaf-by-leaf walks;
//   * duplicate indices in one batch: the LAST occurrence wins (sequential semantics of
//     replay_buffer.py:210-216), resolved with an atomicMax stamp;
//   * all fp64 arithmetic uses explicit round-to-nearest intrinsics (no FMA contraction).
// The trees (2 x 1 MiB at capacity 2^16) are L2-resident; the kernels are latency-bound.
#include "common.cuh"

namespace cacto {

constexpr int ST_THREADS = 1024;

// Level-synchronous update by ONE CTA.  The top levels (<= 32 nodes) are folded by warp 0 with shuffles.
__global__ void __launch_bounds__(ST_THREADS) k_segtree_update(double* __restrict__ sum, double* __restrict__ mn, int cap,
                                                               const int64_t* __restrict__ idx, const double* __restrict__ val,
                                                               int n, int* __restrict__ stamp) {
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += ST_THREADS) atomicMax(&stamp[(int)idx[i]], i);
  __syncthreads();
  for (int i = tid; i < n; i += ST_THREADS) {
    const int j = (int)idx[i];
    if (stamp[j] == i) {
      if (sum) sum[cap + j] = val[i];
      if (mn) mn[cap + j] = val[i];
    }
  }
  __syncthreads();
  // levels whose node count exceeds one warp: recompute every touched ancestor from its (final) children
  // (a level is a chain of L2 round trips: 4 nodes per thread are in flight together -- at B = 4096 each of the 1024 threads
  //  owns exactly 4 leaves -- instead of one after the other; the leaf indices are read once)
  int shift = 1;
  if (n <= 4 * ST_THREADS) {                          // every thread walks the level loop (it holds barriers), with 0 - 4 live leaves
    int leaf[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) leaf[u] = (tid + u * ST_THREADS < n) ? cap + (int)idx[tid + u * ST_THREADS] : -1;
    for (; (cap >> shift) > 32; ++shift) {
      double l[4], r[4], a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int node = leaf[u] >> shift;
        if (leaf[u] >= 0 && sum) { l[u] = sum[2 * node]; r[u] = sum[2 * node + 1]; }
        if (leaf[u] >= 0 && mn) { a[u] = mn[2 * node]; b[u] = mn[2 * node + 1]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int node = leaf[u] >> shift;
        if (leaf[u] >= 0 && sum) sum[node] = __dadd_rn(l[u], r[u]);
        if (leaf[u] >= 0 && mn) mn[node] = (b[u] < a[u]) ? b[u] : a[u];    // Python min(a, b): b only if strictly smaller
      }
      __syncthreads();
    }
  }
  if (n > 4 * ST_THREADS) {
    for (; (cap >> shift) > 32; ++shift) {
      for (int i = tid; i < n; i += ST_THREADS) {
        const int node = (cap + (int)idx[i]) >> shift;
        if (sum) {
          const double l = sum[2 * node], r = sum[2 * node + 1];
          sum[node] = __dadd_rn(l, r);
        }
        if (mn) {
          const double a = mn[2 * node], b = mn[2 * node + 1];
          mn[node] = (b < a) ? b : a;
        }
      }
      __syncthreads();
    }
  }
  // remaining levels: W = cap >> shift (<= 32) nodes [W, 2W) and everything above, by warp 0
  if (tid < 32) {
    int W = cap >> shift;
    double s = 0.0, m = 0.0;
    if (W >= 1) {
      if (tid < W) {
        const int node = W + tid;
        if (sum) {
          const double l = sum[2 * node], r = sum[2 * node + 1];
          s = __dadd_rn(l, r);
          sum[node] = s;
        }
        if (mn) {
          const double a = mn[2 * node], b = mn[2 * node + 1];
          m = (b < a) ? b : a;
          mn[node] = m;
        }
      }
      for (W >>= 1; W >= 1; W >>= 1) {
        const int src = (2 * tid) & 31;
        const double l = __shfl_sync(0xffffffffu, s, src), r = __shfl_sync(0xffffffffu, s, src + 1 > 31 ? 31 : src + 1);
        const double a = __shfl_sync(0xffffffffu, m, src), b = __shfl_sync(0xffffffffu, m, src + 1 > 31 ? 31 : src + 1);
        if (tid < W) {
          s = __dadd_rn(l, r);
          m = (b < a) ? b : a;
          if (sum) sum[W + tid] = s;
          if (mn) mn[W + tid] = m;
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n; i += ST_THREADS) stamp[(int)idx[i]] = -1;
}

// Multi-CTA form of the same update (capacity <= 2^18): because every internal node IS left (+|min) right of its children
// at all times, recomputing ALL of them from the final leaves writes the bits the touched-ancestor walk would have written and
// leaves the others as they were.  Each CTA owns a subtree of SR_LEAVES leaves: it applies the batch entries that fall into its
// range (last occurrence wins: shared-memory stamps), folds the subtree in shared memory and writes its nodes; the last CTA to
// finish (ticket in stamp[0], idle value -1) folds the subtree roots into the top of the tree.  Two block-wide L2 round trips
// instead of 17 barrier-separated ones: 82 -> ~8 us at B = 4096, capacity 2^16.
constexpr int SR_LEAVES = 512, SR_THREADS = 256;

__global__ void __launch_bounds__(SR_THREADS) k_segtree_rebuild(double* __restrict__ sum, double* __restrict__ mn, int cap,
                                                                const int64_t* __restrict__ idx, const double* __restrict__ val,
                                                                int n, int* __restrict__ ticket) {
  __shared__ double s_sum[2 * SR_LEAVES], s_min[2 * SR_LEAVES];     // heap order: node k of the subtree at [k], leaves at [L, 2L)
  __shared__ int s_stamp[SR_LEAVES];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int L = cap < SR_LEAVES ? cap : SR_LEAVES, R = cap / L;     // leaves per CTA, number of subtrees (= gridDim.x)
  const int base = blockIdx.x * L;
  for (int j = tid; j < L; j += SR_THREADS) {
    s_stamp[j] = -1;
    if (sum) s_sum[L + j] = sum[cap + base + j];
    if (mn) s_min[L + j] = mn[cap + base + j];
  }
  __syncthreads();
  for (int i = tid; i < n; i += SR_THREADS) {
    const int64_t j = idx[i] - base;
    if (j >= 0 && j < L) atomicMax(&s_stamp[(int)j], i);
  }
  __syncthreads();
  for (int j = tid; j < L; j += SR_THREADS) {
    const int i = s_stamp[j];
    if (i >= 0) {
      const double v = val[i];
      if (sum) { s_sum[L + j] = v; sum[cap + base + j] = v; }
      if (mn) { s_min[L + j] = v; mn[cap + base + j] = v; }
    }
  }
  __syncthreads();
  for (int W = L >> 1; W >= 1; W >>= 1) {
    for (int k = tid; k < W; k += SR_THREADS) {
      const int node = W + k;
      if (sum) s_sum[node] = __dadd_rn(s_sum[2 * node], s_sum[2 * node + 1]);
      if (mn) { const double a = s_min[2 * node], b = s_min[2 * node + 1]; s_min[node] = (b < a) ? b : a; }   // Python min(a, b)
    }
    __syncthreads();
  }
  // subtree node k = 2^d + o  <->  tree node ((R + blockIdx.x) << d) + o
  for (int k = 1 + tid; k < L; k += SR_THREADS) {
    const int d = 31 - __clz(k), g = ((R + (int)blockIdx.x) << d) + (k - (1 << d));
    if (sum) sum[g] = s_sum[k];
    if (mn) mn[g] = s_min[k];
  }
  if (R == 1) return;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 2);      // -1, 0, ..., R - 2
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int r = tid; r < R; r += SR_THREADS) {                                // the subtree roots, as written by their CTAs (L2)
    if (sum) s_sum[R + r] = __ldcg(sum + R + r);
    if (mn) s_min[R + r] = __ldcg(mn + R + r);
  }
  __syncthreads();
  for (int W = R >> 1; W >= 1; W >>= 1) {
    for (int k = tid; k < W; k += SR_THREADS) {
      const int node = W + k;
      if (sum) { s_sum[node] = __dadd_rn(s_sum[2 * node], s_sum[2 * node + 1]); sum[node] = s_sum[node]; }
      if (mn) { const double a = s_min[2 * node], b = s_min[2 * node + 1]; s_min[node] = (b < a) ? b : a; mn[node] = s_min[node]; }
    }
    __syncthreads();
  }
  if (tid == 0) *ticket = -1;
}

// segment_tree.py:36-49 -- the recursion fixes the association order of the partial sums.
__device__ double reduce_sum(const double* v, int start, int end, int node, int lo, int hi) {
  if (start == lo && end == hi) return v[node];
  const int mid = (lo + hi) / 2;
  if (end <= mid) return reduce_sum(v, start, end, 2 * node, lo, mid);
  if (mid + 1 <= start) return reduce_sum(v, start, end, 2 * node + 1, mid + 1, hi);
  const double a = reduce_sum(v, start, mid, 2 * node, lo, mid);
  const double b = reduce_sum(v, mid + 1, end, 2 * node + 1, mid + 1, hi);
  return __dadd_rn(a, b);
}
__device__ double reduce_min(const double* v, int start, int end, int node, int lo, int hi) {
  if (start == lo && end == hi) return v[node];
  const int mid = (lo + hi) / 2;
  if (end <= mid) return reduce_min(v, start, end, 2 * node, lo, mid);
  if (mid + 1 <= start) return reduce_min(v, start, end, 2 * node + 1, mid + 1, hi);
  const double a = reduce_min(v, start, mid, 2 * node, lo, mid);
  const double b = reduce_min(v, mid + 1, end, 2 * node + 1, mid + 1, hi);
  return (b < a) ? b : a;
}

__global__ void k_segtree_reduce(const double* __restrict__ sum, const double* __restrict__ mn, int cap, int start, int end_incl,
                                 double* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    if (sum) out[0] = reduce_sum(sum, start, end_incl, 1, 0, cap - 1);
    if (mn) out[1] = reduce_min(mn, start, end_incl, 1, 0, cap - 1);
  }
}

// replay_buffer.py:139-157 -- stratified proportional sampling.
__global__ void __launch_bounds__(256) k_segtree_sample(const double* __restrict__ sum, const double* __restrict__ mn, int cap,
                                                        int max_idx, const double* __restrict__ uniforms, int n,
                                                        int64_t* __restrict__ idx_out, double* __restrict__ leaf_out,
                                                        double* __restrict__ totals) {
  __shared__ double s_total;
  if (threadIdx.x == 0) {
    // sum(0, max_idx - 1): python end-exclusive -> inclusive max_idx - 2 (quirk Q4: newest slot excluded)
    const double pt = reduce_sum(sum, 0, max_idx - 2, 1, 0, cap - 1);
    s_total = pt;
    if (blockIdx.x == 0) {
      totals[0] = pt;
      totals[1] = sum[1];
      totals[2] = mn[1];
    }
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double segment = __ddiv_rn(s_total, (double)n);
  double p = __dadd_rn(__dmul_rn(uniforms[i], segment), __dmul_rn((double)i, segment));
  int node = 1;
  while (node < cap) {                   // segment_tree.py:124-130
    const double left = sum[2 * node];
    if (left > p) {
      node = 2 * node;
    } else {
      p = __dsub_rn(p, left);
      node = 2 * node + 1;
    }
  }
  idx_out[i] = (int64_t)(node - cap);
  leaf_out[i] = sum[node];
}

// SumSegmentTree.find_prefixsum_idx for a batch of given prefix sums (segment_tree.py:105-131).
__global__ void __launch_bounds__(256) k_segtree_find(const double* __restrict__ sum, int cap, const double* __restrict__ prefix,
                                                      int n, int64_t* __restrict__ idx_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p = prefix[i];
  int node = 1;
  while (node < cap) {
    const double left = sum[2 * node];
    if (left > p) {
      node = 2 * node;
    } else {
      p = __dsub_rn(p, left);
      node = 2 * node + 1;
    }
  }
  idx_out[i] = (int64_t)(node - cap);
}

// replay_buffer.py:47-61 / :178-188 -- one warp per sampled row.
__global__ void __launch_bounds__(256) k_buffer_gather(const double* __restrict__ storage, int ns, const int64_t* __restrict__ idx,
                                                       int n, float* __restrict__ state, float* __restrict__ prtg,
                                                       float* __restrict__ state_next, float* __restrict__ dVdx,
                                                       float* __restrict__ done, double* __restrict__ term) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const int W = 3 * ns + 3;
  const double* row = storage + (int64_t)idx[warp] * W;
  for (int c = lane; c < W; c += 32) {
    const double v = row[c];
    if (c < ns) state[(int64_t)warp * ns + c] = (float)v;
    else if (c == ns) prtg[warp] = (float)v;
    else if (c < 2 * ns + 1) state_next[(int64_t)warp * ns + (c - ns - 1)] = (float)v;
    else if (c < 3 * ns + 1) dVdx[(int64_t)warp * ns + (c - 2 * ns - 1)] = (float)v;
    else if (c == 3 * ns + 1) done[warp] = (float)v;
    else term[warp] = v;
  }
}

// exp_counter[idx] += 1, a duplicated index counted once (NumPy fancy-index +=, replay_buffer.py:174).
__global__ void __launch_bounds__(ST_THREADS) k_exp_counter(double* __restrict__ exp_counter, const int64_t* __restrict__ idx, int n,
                                                            int* __restrict__ stamp) {
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += ST_THREADS) atomicMax(&stamp[(int)idx[i]], i);
  __syncthreads();
  for (int i = tid; i < n; i += ST_THREADS) {
    const int j = (int)idx[i];
    if (stamp[j] == i) {
      exp_counter[j] = __dadd_rn(exp_counter[j], 1.0);
      stamp[j] = -1;
    }
  }
}

}  // namespace cacto

using namespace cacto;

static bool pow2(int c) { return c > 0 && (c & (c - 1)) == 0; }

extern "C" int cacto_segtree_update(double* sum_tree, double* min_tree, int32_t capacity, const int64_t* idx, const double* value,
                                    int32_t n, int32_t* stamp, void* stream) {
  if ((!sum_tree && !min_tree) || !idx || !value || !stamp) return CACTO_E_ARG;
  if (!pow2(capacity) || capacity < 2 || n < 0) return CACTO_E_SIZE;
  if (n == 0) return 0;
  if (capacity <= (SR_LEAVES * SR_LEAVES)) {           // the subtree roots of the rebuild fit its shared-memory arrays
    const int ctas = capacity < SR_LEAVES ? 1 : capacity / SR_LEAVES;
    k_segtree_rebuild<<<ctas, SR_THREADS, 0, (cudaStream_t)stream>>>(sum_tree, min_tree, capacity, idx, value, n, stamp);
  } else {
    k_segtree_update<<<1, ST_THREADS, 0, (cudaStream_t)stream>>>(sum_tree, min_tree, capacity, idx, value, n, stamp);
  }
  CACTO_LAUNCH_CHECK();
  return 0;
}

extern "C" int cacto_segtree_reduce(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t start, int32_t end,
                                    double* out, void* stream) {
  if ((!sum_tree && !min_tree) || !out) return CACTO_E_ARG;
  if (!pow2(capacity)) return CACTO_E_SIZE;
  if (end < 0) end += capacity;          // segment_tree.py:71-72
  end -= 1;                              // :73
  if (start < 0 || end < start || end >= capacity) return CACTO_E_SIZE;
  k_segtree_reduce<<<1, 32, 0, (cudaStream_t)stream>>>(sum_tree, min_tree, capacity, start, end, out);
  CACTO_LAUNCH_CHECK();
  return 0;
}

extern "C" int cacto_segtree_sample(const double* sum_tree, const double* min_tree, int32_t capacity, int32_t max_idx,
                                    const double* uniforms, int32_t n, int64_t* idx, double* leaf, double* totals, void* stream) {
  if (!sum_tree || !min_tree || !uniforms || !idx || !leaf || !totals) return CACTO_E_ARG;
  if (!pow2(capacity) || n <= 0 || max_idx < 2 || max_idx > capacity) return CACTO_E_SIZE;
  k_segtree_sample<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sum_tree, min_tree, capacity, max_idx, uniforms, n, idx, leaf,
                                                                      totals);
  CACTO_LAUNCH_CHECK();
  return 0;
}

extern "C" int cacto_segtree_find(const double* sum_tree, int32_t capacity, const double* prefix, int32_t n, int64_t* idx,
                                  void* stream) {
  if (!sum_tree || !prefix || !idx) return CACTO_E_ARG;
  if (!pow2(capacity) || n < 0) return CACTO_E_SIZE;
  if (n == 0) return 0;
  k_segtree_find<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sum_tree, capacity, prefix, n, idx);
  CACTO_LAUNCH_CHECK();
  return 0;
}

extern "C" int cacto_buffer_gather(const double* storage, int32_t ns, const int64_t* idx, int32_t n, float* state,
                                   float* partial_rtg, float* state_next, float* dVdx, float* done, double* term,
                                   double* exp_counter, int32_t* stamp, void* stream) {
  if (!storage || !idx || !state || !partial_rtg || !state_next || !dVdx || !done || !term) return CACTO_E_ARG;
  if (ns < 2 || ns > CACTO_MAX_NS || n < 0) return CACTO_E_SIZE;
  if (exp_counter && !stamp) return CACTO_E_ARG;
  if (n == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  k_buffer_gather<<<(n * 32 + 255) / 256, 256, 0, st>>>(storage, ns, idx, n, state, partial_rtg, state_next, dVdx, done, term);
  CACTO_LAUNCH_CHECK();
  if (exp_counter) {
    k_exp_counter<<<1, ST_THREADS, 0, st>>>(exp_counter, idx, n, stamp);
    CACTO_LAUNCH_CHECK();
  }
  return 0;
}
