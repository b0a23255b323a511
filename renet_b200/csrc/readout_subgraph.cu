// Read-out sub-graph of a batched history graph (exact optimisation (i)).
//
// The reference computes layer 2 of the aggregator on every node of the batched graph and then keeps only the read-out
// rows (Aggregator.py:139-140: embeds_mean = g.ndata.pop('h'); embeds_mean[node_ids_graph]) -- 26 % of the rows at
// ICEWS18 scale.  Layer 2's value at a read-out node depends on layer 1 at its in-neighbours only, so running layer 2 on
// the sub-graph {edges whose DESTINATION is a read-out node} gives identical values on every consumed row.  This file
// builds that sub-graph on the device, without a host round trip:
//
//   uniq      [S]    distinct read-out nodes, ascending (compact destination u  <->  node uniq[u]); tail = 0
//   readout_c [S]    read-out row -> compact destination
//   row_ptr2  [S+1]  CSR by compact destination; destinations past U have no edges (row_ptr2[u] = E2)
//   col_src2 / col_type2   in-edges of the compact destinations, in the full CSR's order (sources keep FULL-graph ids)
//   norm2     [S]    norm of the compact destinations (tail = 1)
//   counts    [2]    {U, E2} on the device (launches are sized by the capacities S and E_cap: nothing waits for them)
//
// Integer, HBM/L2-bound work: flags over N nodes, two prefix sums (CUB), one compaction, one edge copy.
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace renet {
namespace {

inline int64_t align256(int64_t x) { return (x + 255) & ~int64_t(255); }

__global__ void rs_mark_kernel(const int32_t* __restrict__ readout, int S, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < S) flag[__ldg(readout + i)] = 1;
}

// pos = exclusive scan of flag over N+1 entries (pos[N] = U)
__global__ void rs_compact_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos,
                                  const int32_t* __restrict__ readout, const int32_t* __restrict__ row_ptr,
                                  const float* __restrict__ norm, int N, int S, int32_t* __restrict__ uniq,
                                  int32_t* __restrict__ readout_c, int32_t* __restrict__ deg2, float* __restrict__ norm2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N && flag[i]) {
    const int u = pos[i];
    uniq[u] = i;
    norm2[u] = __ldg(norm + i);
    deg2[u] = __ldg(row_ptr + i + 1) - __ldg(row_ptr + i);
  }
  if (i < S) {
    readout_c[i] = pos[__ldg(readout + i)];
    if (i >= pos[N]) { uniq[i] = 0; norm2[i] = 1.f; deg2[i] = 0; }      // unused capacity: valid, edge-less rows
  }
  if (i == 0) deg2[S] = 0;
}

// one warp per compact destination: copy its in-edges
__global__ void rs_copy_edges_kernel(const int32_t* __restrict__ uniq, const int32_t* __restrict__ row_ptr,
                                     const int32_t* __restrict__ row_ptr2, const int32_t* __restrict__ col_src,
                                     const int32_t* __restrict__ col_type, const int32_t* __restrict__ pos, int N, int S,
                                     int32_t* __restrict__ col_src2, int32_t* __restrict__ col_type2,
                                     int32_t* __restrict__ counts) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w == 0 && lane == 0) { counts[0] = pos[N]; counts[1] = row_ptr2[S]; }
  if (w >= S || w >= pos[N]) return;
  const int v = uniq[w];
  const int e0 = __ldg(row_ptr + v), n = __ldg(row_ptr + v + 1) - e0, o = row_ptr2[w];
  for (int k = lane; k < n; k += 32) {
    col_src2[o + k] = __ldg(col_src + e0 + k);
    col_type2[o + k] = __ldg(col_type + e0 + k);
  }
}

}  // namespace
}  // namespace renet

using namespace renet;

extern "C" {

int64_t renet_readout_subgraph_workspace_bytes(int64_t N, int64_t S) {
  size_t c1 = 0, c2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, c1, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(N + 1));
  cub::DeviceScan::ExclusiveSum(nullptr, c2, (const int32_t*)nullptr, (int32_t*)nullptr, (int)(S + 1));
  return 2 * align256((N + 1) * 4) + align256((S + 1) * 4) + align256((int64_t)(c1 > c2 ? c1 : c2)) + 256;
}

int renet_readout_subgraph(const int32_t* readout, int64_t S, int64_t N, const int32_t* row_ptr, const int32_t* col_src,
                           const int32_t* col_type, const float* norm, int32_t* uniq, int32_t* readout_c,
                           int32_t* row_ptr2, int32_t* col_src2, int32_t* col_type2, float* norm2, int32_t* counts,
                           void* workspace, int64_t workspace_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  RENET_CHECK_ARG(S >= 0 && N >= 0 && N < (int64_t(1) << 31) - 1 && S < (int64_t(1) << 31) - 1, "renet_readout_subgraph: bad S/N");
  if (S == 0) return RENET_OK;
  RENET_CHECK_ARG(readout && row_ptr && norm && uniq && readout_c && row_ptr2 && col_src2 && col_type2 && norm2 && counts &&
                      workspace, "renet_readout_subgraph: null pointer");
  RENET_CHECK_ARG(workspace_bytes >= renet_readout_subgraph_workspace_bytes(N, S), "renet_readout_subgraph: workspace too small");
  char* ws = (char*)workspace;
  int32_t* flag = (int32_t*)ws;  ws += align256((N + 1) * 4);
  int32_t* pos = (int32_t*)ws;   ws += align256((N + 1) * 4);
  int32_t* deg2 = (int32_t*)ws;  ws += align256((S + 1) * 4);
  size_t cub_bytes = (size_t)(workspace_bytes - (ws - (char*)workspace));
  RENET_CHECK_CUDA(cudaMemsetAsync(flag, 0, (N + 1) * 4, stream));
  rs_mark_kernel<<<(unsigned)((S + 255) / 256), 256, 0, stream>>>(readout, (int)S, flag);
  RENET_CHECK_LAUNCH("rs_mark_kernel");
  RENET_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws, cub_bytes, flag, pos, (int)(N + 1), stream));
  count_launch(2);
  const int64_t m = N > S ? N : S;
  rs_compact_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(flag, pos, readout, row_ptr, norm, (int)N, (int)S, uniq,
                                                                   readout_c, deg2, norm2);
  RENET_CHECK_LAUNCH("rs_compact_kernel");
  RENET_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws, cub_bytes, deg2, row_ptr2, (int)(S + 1), stream));
  count_launch(2);
  rs_copy_edges_kernel<<<(unsigned)((S * 32 + 255) / 256), 256, 0, stream>>>(uniq, row_ptr, row_ptr2, col_src, col_type, pos,
                                                                           (int)N, (int)S, col_src2, col_type2, counts);
  RENET_CHECK_LAUNCH("rs_copy_edges_kernel");
  return RENET_OK;
}

}  // extern "C"
