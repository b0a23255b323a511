// Device half of the history-graph batcher: utils.make_subgraph (reference utils.py:115-131) + dgl.batch (utils.py:238)
// for ALL timestamps of a batch at once, against a graph store that lives in HBM.
//
// The host (renet_host_plan_batch) has already chosen the components (timestamps), marked and numbered the selected
// nodes (newid: batched node id per local row, -1 = not selected).  What is left is the O(edges) part: every edge of
// every touched timestamp graph is a CANDIDATE; it survives when both endpoints are selected.  Candidates are laid out
// component-major and each timestamp's edge list is destination-sorted, and batched node ids ascend with (component,
// local row), so the survivors, in candidate order, are already sorted by batched destination: the CSR is a stream
// compaction plus a boundary fill, no sort.
//
//   induce_count_kernel   candidates -> per-CTA survivor counts          (reads ~12 B/candidate, L2 resident)
//   induce_scan_kernel    exclusive scan of the CTA counts, total E
//   induce_emit_kernel    survivors -> col_src / col_type_s / col_type_o / dst (ordered, CTA-local scan)
//   induce_rowptr_kernel  row_ptr from the sorted destinations
//   induce_norm_kernel    norm = 1 / max(in-degree, 1) on the induced graph (utils.py:126-127)
//
// HBM-bound integer work: ~0.75 M candidates x (8 B edge + 2 random 4-B mark reads + 8 B types) per batch.
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include "common.cuh"

namespace renet {
namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 4;
constexpr int kChunk = kThreads * kPerThread;

// largest c in [0, G) with cand_off[c] <= k   (cand_off[G] > k is guaranteed by the caller)
__device__ __forceinline__ int find_comp(const int32_t* __restrict__ cand_off, int G, int k) {
  int lo = 0, hi = G;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(cand_off + mid) <= k) lo = mid; else hi = mid;
  }
  return lo;
}

struct Cand {
  int32_t src, dst;      // batched node ids, or negative when the edge does not survive
  int64_t ge;            // index into the graph store's edge arrays
};

__device__ __forceinline__ void load_cands(Cand (&cd)[kPerThread], const int64_t* __restrict__ g_edge_off,
                                           const int32_t* __restrict__ g_src, const int32_t* __restrict__ g_dst,
                                           const int32_t* __restrict__ comp_graph, const int32_t* __restrict__ mark_off,
                                           const int32_t* __restrict__ cand_off, const int32_t* __restrict__ newid, int G,
                                           int e_cand, int k0) {
  int c = (k0 < e_cand) ? find_comp(cand_off, G, k0) : 0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    const int k = k0 + j;
    cd[j].src = cd[j].dst = -1;
    cd[j].ge = 0;
    if (k < e_cand) {
      while (k >= __ldg(cand_off + c + 1)) ++c;            // also skips timestamps without edges
      const int64_t ge = __ldg(g_edge_off + __ldg(comp_graph + c)) + (k - __ldg(cand_off + c));
      const int32_t* m = newid + __ldg(mark_off + c);
      const int32_t s = __ldg(m + __ldg(g_src + ge));
      const int32_t d = __ldg(m + __ldg(g_dst + ge));
      cd[j].ge = ge;
      if ((s | d) >= 0) { cd[j].src = s; cd[j].dst = d; }
    }
  }
}

__global__ void __launch_bounds__(kThreads) induce_count_kernel(
    const int64_t* __restrict__ g_edge_off, const int32_t* __restrict__ g_src, const int32_t* __restrict__ g_dst,
    const int32_t* __restrict__ comp_graph, const int32_t* __restrict__ mark_off, const int32_t* __restrict__ cand_off,
    const int32_t* __restrict__ newid, int G, int e_cand, int32_t* __restrict__ block_cnt) {
  Cand cd[kPerThread];
  load_cands(cd, g_edge_off, g_src, g_dst, comp_graph, mark_off, cand_off, newid, G, e_cand,
             blockIdx.x * kChunk + threadIdx.x * kPerThread);
  int n = 0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) n += cd[j].dst >= 0;
  using Reduce = cub::BlockReduce<int, kThreads>;
  __shared__ typename Reduce::TempStorage tmp;
  const int total = Reduce(tmp).Sum(n);
  if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

// one CTA: exclusive scan of block_cnt[0..nblk) -> block_off, total -> e_count[0]
__global__ void __launch_bounds__(1024) induce_scan_kernel(const int32_t* __restrict__ block_cnt, int nblk,
                                                           int32_t* __restrict__ block_off, int32_t* __restrict__ e_count) {
  using Scan = cub::BlockScan<int, 1024>;
  __shared__ typename Scan::TempStorage tmp;
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblk; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < nblk ? block_cnt[i] : 0;
    int excl, total;
    Scan(tmp).ExclusiveSum(v, excl, total);
    if (i < nblk) block_off[i] = carry + excl;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) e_count[0] = carry;
}

__global__ void __launch_bounds__(kThreads) induce_emit_kernel(
    const int64_t* __restrict__ g_edge_off, const int32_t* __restrict__ g_src, const int32_t* __restrict__ g_dst,
    const int32_t* __restrict__ g_type_s, const int32_t* __restrict__ g_type_o, const int32_t* __restrict__ comp_graph,
    const int32_t* __restrict__ mark_off, const int32_t* __restrict__ cand_off, const int32_t* __restrict__ newid, int G,
    int e_cand, const int32_t* __restrict__ block_off, int32_t* __restrict__ col_src, int32_t* __restrict__ col_type_s,
    int32_t* __restrict__ col_type_o, int32_t* __restrict__ col_dst) {
  Cand cd[kPerThread];
  load_cands(cd, g_edge_off, g_src, g_dst, comp_graph, mark_off, cand_off, newid, G, e_cand,
             blockIdx.x * kChunk + threadIdx.x * kPerThread);
  int n = 0;
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) n += cd[j].dst >= 0;
  using Scan = cub::BlockScan<int, kThreads>;
  __shared__ typename Scan::TempStorage tmp;
  int pos;
  Scan(tmp).ExclusiveSum(n, pos);
  pos += block_off[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kPerThread; ++j) {
    if (cd[j].dst >= 0) {
      col_src[pos] = cd[j].src;
      col_dst[pos] = cd[j].dst;
      col_type_s[pos] = __ldg(g_type_s + cd[j].ge);
      col_type_o[pos] = __ldg(g_type_o + cd[j].ge);
      ++pos;
    }
  }
}

// row_ptr[v] = first position whose destination is >= v; thread p owns the nodes in (dst[p-1], dst[p]], thread E the tail
__global__ void induce_rowptr_kernel(const int32_t* __restrict__ col_dst, const int32_t* __restrict__ e_count, int N,
                                     int32_t* __restrict__ row_ptr) {
  const int E = e_count[0];
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > E) return;
  const int prev = p > 0 ? col_dst[p - 1] : -1;
  const int last = p < E ? col_dst[p] : N;
  for (int v = prev + 1; v <= last; ++v) row_ptr[v] = p;
}

__global__ void induce_norm_kernel(const int32_t* __restrict__ row_ptr, int N, float* __restrict__ norm) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= N) return;
  const int d = row_ptr[v + 1] - row_ptr[v];
  norm[v] = 1.0f / (float)(d > 0 ? d : 1);
}

inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }
inline int64_t n_blocks(int64_t e_cand) { return (e_cand + kChunk - 1) / kChunk; }

}  // namespace
}  // namespace renet

using namespace renet;

extern "C" int64_t renet_induce_workspace_bytes(int64_t e_cand) {
  if (e_cand < 0) return 0;
  return 2 * align256(n_blocks(e_cand) * 4 + 4) + align256(e_cand * 4 + 4);
}

extern "C" int renet_induce_edges(const int64_t* g_edge_off, const int32_t* g_src, const int32_t* g_dst,
                                  const int32_t* g_type_s, const int32_t* g_type_o, const int32_t* comp_graph,
                                  const int32_t* mark_off, const int32_t* cand_off, const int32_t* newid, int64_t G,
                                  int64_t N, int64_t e_cand, int32_t* row_ptr, int32_t* col_src, int32_t* col_type_s,
                                  int32_t* col_type_o, float* norm, int32_t* e_count, void* workspace,
                                  int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RENET_CHECK_ARG(G >= 0 && N >= 0 && e_cand >= 0 && N < (int64_t(1) << 31) && e_cand < (int64_t(1) << 31),
                  "renet_induce_edges: bad sizes");
  RENET_CHECK_ARG(row_ptr && e_count && (N == 0 || norm), "renet_induce_edges: null outputs");
  RENET_CHECK_ARG(e_cand == 0 || (g_edge_off && g_src && g_dst && g_type_s && g_type_o && comp_graph && mark_off &&
                                  cand_off && newid && col_src && col_type_s && col_type_o && G > 0),
                  "renet_induce_edges: null inputs");
  RENET_CHECK_ARG(workspace_bytes >= renet_induce_workspace_bytes(e_cand) && (workspace || workspace_bytes == 0),
                  "renet_induce_edges: workspace too small");
  const int64_t nblk = n_blocks(e_cand);
  char* ws = (char*)workspace;
  int32_t* block_cnt = (int32_t*)ws;  ws += align256(nblk * 4 + 4);
  int32_t* block_off = (int32_t*)ws;  ws += align256(nblk * 4 + 4);
  int32_t* col_dst = (int32_t*)ws;
  if (nblk > 0) {
    induce_count_kernel<<<(unsigned)nblk, kThreads, 0, stream>>>(g_edge_off, g_src, g_dst, comp_graph, mark_off, cand_off,
                                                                 newid, (int)G, (int)e_cand, block_cnt);
    RENET_CHECK_LAUNCH("induce_count_kernel");
  }
  induce_scan_kernel<<<1, 1024, 0, stream>>>(block_cnt, (int)nblk, block_off, e_count);
  RENET_CHECK_LAUNCH("induce_scan_kernel");
  if (nblk > 0) {
    induce_emit_kernel<<<(unsigned)nblk, kThreads, 0, stream>>>(g_edge_off, g_src, g_dst, g_type_s, g_type_o, comp_graph,
                                                                mark_off, cand_off, newid, (int)G, (int)e_cand, block_off,
                                                                col_src, col_type_s, col_type_o, col_dst);
    RENET_CHECK_LAUNCH("induce_emit_kernel");
  }
  induce_rowptr_kernel<<<(unsigned)((e_cand + 1 + 255) / 256), 256, 0, stream>>>(col_dst, e_count, (int)N, row_ptr);
  RENET_CHECK_LAUNCH("induce_rowptr_kernel");
  if (N > 0) {
    induce_norm_kernel<<<(unsigned)((N + 255) / 256), 256, 0, stream>>>(row_ptr, (int)N, norm);
    RENET_CHECK_LAUNCH("induce_norm_kernel");
  }
  return RENET_OK;
}
