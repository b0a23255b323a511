// Fused gather with a per-warp shared-memory ring of source rows (forward only).
//
// rgcn_gather_d200_kernel (rgcn_tile.cuh) stages an edge's operands in registers: 24 registers per edge, two edges in
// flight per warp, 24 warps per SM.  ncu shows the SM sub-partitions issuing 39 % of the cycles and otherwise waiting
// on L2 latency with every warp stalled: the kernel needs more bytes in flight per SM, and registers cannot provide
// them.  Here the 800-byte source rows -- the streamed, never re-used half of the traffic -- travel through cp.async
// (LDGSTS, L1-bypassing) into a private ring of kRing stages per warp, so a warp has kRing source rows in flight without
// holding a register for them; the relation rows (1600 B, re-used, L1/L2 resident) stay on the LDG path, double
// buffered in registers one edge ahead and announced to L1 three edges ahead with prefetch.global.L1.
//
// Everything else is the tile kernel: CTA = 16 destinations, edge range split evenly over 8 warps, warp-level segmented
// reduction, deterministic atomic-free hand-over of partial sums (TileHeads), fused norm / self-loop / ReLU epilogue.
// The self-loop rows of the tile are fetched into the (then idle) ring stages by each warp as it finishes its slice.
#pragma once
#include "rgcn_tile.cuh"

namespace renet {

constexpr int kRing = 4;

template <int N>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

template <bool RELU, bool HAS_LOOP, bool INDEXED>
__global__ void __launch_bounds__(kTileWarps * 32, 3)
rgcn_gather_ring_kernel(const float* __restrict__ H, const int32_t* __restrict__ h_index, const float* __restrict__ W,
                        const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_src,
                        const int32_t* __restrict__ col_type, const float* __restrict__ norm,
                        float* __restrict__ Hout, int N) {
  __shared__ __align__(16) float agg[kTileNodes][200];
  __shared__ __align__(16) float head[kTileWarps][200];
  __shared__ __align__(16) float ring[kTileWarps][kRing][200];
  __shared__ int head_mask[kTileNodes];
  __shared__ float normbuf[kTileNodes];
  __shared__ int s_rp[kTileNodes + 1];
  static_assert(kRing >= 2 && (kRing & (kRing - 1)) == 0, "ring depth: power of two, and two stages hold the loop rows");
  static_assert(2 * kTileWarps >= kTileNodes, "each warp parks two self-loop rows in its ring");

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ln = lane < 25 ? lane : lane - 25;   // lanes 25..31 shadow lanes 0..6 (same addresses: no extra sectors,
  const bool active = lane < 25;                 //  no divergence in the loop); only their stores are masked
  const int v0 = blockIdx.x * kTileNodes;
  const int nv = min(kTileNodes, N - v0);
  if (tid < kTileNodes) head_mask[tid] = 0;
  if (tid < nv) normbuf[tid] = __ldg(norm + v0 + tid);
  if (tid <= nv) s_rp[tid] = __ldg(row_ptr + v0 + tid);
  __syncthreads();

  const int ebeg = s_rp[0], eend = s_rp[nv];
  const int chunk = (eend - ebeg + kTileWarps - 1) / kTileWarps;
  const int e0 = ebeg + warp * chunk;
  const int e1 = min(eend, e0 + chunk);
  const int n = max(0, e1 - e0);
  float (*my_ring)[200] = ring[warp];

  if (n > 0) {
    // edge indices, 32 at a time, two chunks resident (the prefetch distance is far below 32)
    int cs, ct, ns, nt, cbase = 0;
    auto load_chunk = [&](int base, int& s, int& t) {
      const int e = base + lane;
      s = 0; t = 0;
      if (e < e1) {
        s = __ldg(col_src + e);
        t = __ldg(col_type + e);
        if (INDEXED) s = __ldg(h_index + s);
      }
    };
    load_chunk(e0, cs, ct);
    load_chunk(e0 + 32, ns, nt);
    auto src_of = [&](int k) { const int i = k - cbase; return i < 32 ? __shfl_sync(0xffffffffu, cs, i) : __shfl_sync(0xffffffffu, ns, i - 32); };
    auto type_of = [&](int k) { const int i = k - cbase; return i < 32 ? __shfl_sync(0xffffffffu, ct, i) : __shfl_sync(0xffffffffu, nt, i - 32); };
    auto issue_src = [&](int k) {     // source row of edge k -> ring stage k % kRing; always one commit group per call
      if (k < n) {
        const int s = src_of(k);
        if (active) {
          const float* g = H + (int64_t)s * 200 + 4 * lane;
          float* d = &my_ring[k & (kRing - 1)][4 * lane];
          cp_async16(d, g);
          cp_async16(d + 100, g + 100);
        }
      }
      cp_async_commit();
    };
    auto load_w = [&](float4 (&w)[4], int t) {
      const float* wp = W + (int64_t)t * 400 + 4 * ln;
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = ldg_f4(wp + 100 * q);
    };
    auto prefetch_w = [&](int t) {
      if (lane < 13) prefetch_l1(W + (int64_t)t * 400 + 32 * lane);
    };

    int node = 0;
    while (s_rp[node + 1] <= e0) ++node;
    int node_end = s_rp[node + 1];
    bool continued = e0 > s_rp[node];
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    auto flush = [&](int nd) {
      float* dst = continued ? head[warp] : agg[nd];
      if (continued && lane == 0) atomicOr(head_mask + nd, 1 << warp);
      continued = false;
      if (active) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float2*>(dst + 2 * (lane + 25 * q)) = make_float2(acc[2 * q], acc[2 * q + 1]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    };

#pragma unroll
    for (int i = 0; i < kRing; ++i) issue_src(i);
    float4 wa[4], wb[4];
    load_w(wa, type_of(0));
    for (int i = 1; i <= 2; ++i)
      if (i < n) prefetch_w(type_of(i));

    auto step = [&](int k, float4 (&wc)[4], float4 (&wn)[4]) {
      if (k - cbase == 32) {             // the consumed index entered the second chunk: rotate
        cs = ns; ct = nt; cbase += 32;
        load_chunk(e0 + cbase + 32, ns, nt);
      }
      if (k + 1 < n) load_w(wn, type_of(k + 1));
      if (k + 3 < n) prefetch_w(type_of(k + 3));
      cp_async_wait_group<kRing - 1>();  // the group of edge k has landed (kRing-1 younger ones may be in flight)
      __syncwarp();
      const float* hp = &my_ring[k & (kRing - 1)][2 * ln];
      float2 h[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) h[q] = *reinterpret_cast<const float2*>(hp + 50 * q);
      __syncwarp();
      issue_src(k + kRing);              // refill the stage that was just read
      const int e = e0 + k;
      if (e >= node_end) {
        flush(node);
        do { ++node; node_end = s_rp[node + 1]; } while (e >= node_end);
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 w = wc[q];          // block (w.x w.y; w.z w.w) = W[b][i][j] row-major: out[j] += in[i] * W[i][j]
        acc[2 * q] = fmaf(h[q].x, w.x, fmaf(h[q].y, w.z, acc[2 * q]));
        acc[2 * q + 1] = fmaf(h[q].x, w.y, fmaf(h[q].y, w.w, acc[2 * q + 1]));
      }
    };
    for (int k = 0; k < n; k += 2) {
      step(k, wa, wb);
      if (k + 1 < n) step(k + 1, wb, wa);
    }
    flush(node);
  }
  // the ring is idle now: park this warp's two self-loop rows in its first two stages
  if (HAS_LOOP) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int r = 2 * warp + j;
      if (r < nv && active) {
        const float* g = Hout + (int64_t)(v0 + r) * 200 + 4 * lane;
        float* d = &my_ring[j][4 * lane];
        cp_async16(d, g);
        cp_async16(d + 100, g + 100);
      }
    }
    cp_async_commit();
  }
  cp_async_wait_all();
  __syncthreads();
  const TileHeads th{head, head_mask};
  for (int i = tid; i < nv * 100; i += kTileWarps * 32) {
    const int r = i / 100, c = (i % 100) * 2;
    const float2 a = tile_row_sum(agg, th, s_rp, r, c);
    const float nvv = normbuf[r];
    float2 o = make_float2(a.x * nvv, a.y * nvv);
    if (HAS_LOOP) {
      const float2 l = *reinterpret_cast<const float2*>(&ring[r >> 1][r & 1][c]);
      o.x += l.x; o.y += l.y;
    }
    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
    *reinterpret_cast<float2*>(Hout + (int64_t)(v0 + r) * 200 + c) = o;
  }
}

}  // namespace renet
