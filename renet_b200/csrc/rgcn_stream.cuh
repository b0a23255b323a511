// Persistent bulk-copy ("stream") gather for the RE-Net shape (d_in = d_out = 200, 100 blocks of 2x2): the kernel behind
// renet_rgcn_gather at batch scale (reference RGCN.py:79-94 + 42-48; DGL fn.sum, RGCN.py:91).
//
// Why: the tile kernel (rgcn_tile.cuh) spends ~108 warp instructions per edge -- 16 per-lane LDGs, their address
// arithmetic, index shuffles, predication -- at 34 % warp occupancy (80 registers hold two edges' loads), and every
// 16-destination tile pays its own dependent index chain and two CTA barriers: it is latency/issue-bound at 65 % of the
// roofline (profiles/r02_gather_ncu.txt).  Here the loads leave the LSU path altogether:
//   * ONE persistent CTA per SM (16 warps).  CTA c owns the destinations [A_c, A_c+1) whose edge range is 1/gridDim of the
//     graph (node-aligned: a 32-ary search in row_ptr), and its 16 warps split that EDGE range evenly at arbitrary cuts;
//   * per edge the warp's elected lane issues two cp.async.bulk copies -- the 800-byte source row and the 1600-byte block
//     table row of the edge's relation -- into a per-warp ring of kStDepth 2400-byte slots in shared memory, completing
//     on the slot's mbarrier; the warp consumes slot i (4 x LDS.64 + 4 x LDS.128 per lane, conflict-free) while the
//     copies of edges i+1 .. i+kStDepth are in flight.  No register holds a load in flight, nothing is predicated per
//     lane on the global path, and the ring keeps streaming across destinations (no per-tile prologue);
//   * the running destination's sum stays in registers (edges are destination-sorted: a segmented reduction); a
//     destination that starts and ends inside the warp's range goes straight from registers through the fused
//     norm / self-loop / activation epilogue to global memory; one cut by a warp boundary is handed over through a
//     per-warp head slot + flag in shared memory and finished by the warp that started it, in edge order -- no atomics,
//     bitwise reproducible.
// The same body is the backward dH kernel (BWD: reversed CSR, transposed blocks, per-edge scale norm[dst], dH += sum).
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace renet {

constexpr int kStWarps = 16;
constexpr int kStThreads = kStWarps * 32;
constexpr int kStDepth = 4;              // edges in flight per warp
constexpr int kStSlot = 2400;            // 800 B source row + 1600 B relation row
constexpr int kStRpCap = 4096;           // row_ptr entries of the CTA's destinations kept in shared memory
constexpr size_t kStSmemBytes = (size_t)kStWarps * kStDepth * kStSlot + (size_t)kStWarps * 800 + (size_t)kStRpCap * 4 +
                                (size_t)kStWarps * kStDepth * 8 + 256;

namespace {

__device__ __forceinline__ void st_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void st_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool st_elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// first probe without bookkeeping (the copy was issued kStDepth edges ago: it has usually landed), bounded spin behind it
__device__ __forceinline__ void st_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  if (!done) mbar_wait(bar, parity);
}
__device__ __forceinline__ int ld_acquire_cta(const int* p) {
  int v;
  asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_cta(int* p, int v) {
  asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

// smallest v in [lo, hi] with rp[v] >= target (rp[hi] >= target is the caller's invariant); whole warp, 32 probes per round
__device__ __forceinline__ int warp_lower_bound(const int32_t* __restrict__ rp, int lo, int hi, int target, int lane) {
  while (hi > lo) {
    const int step = (hi - lo) / 32 + 1;
    const int idx = min(lo + (lane + 1) * step - 1, hi);
    const bool ge = __ldg(rp + idx) >= target;
    const unsigned m = __ballot_sync(0xffffffffu, ge);
    const int f = __ffs(m) - 1;                                  // lane 31 probes hi (or beyond, clipped): m != 0
    const int new_hi = min(lo + (f + 1) * step - 1, hi);
    const int new_lo = f == 0 ? lo : min(lo + f * step - 1, hi) + 1;
    hi = new_hi;
    lo = new_lo;
  }
  return lo;
}

}  // namespace

// BWD = false: Hout[v] = act(norm[v] * sum_e blockdiag(W[type_e]) . X[src_e] + (HAS_LOOP ? Hout[v] : 0))
// BWD = true:  Hout[u] = (HAS_LOOP ? Hout[u] : 0) + sum_e blockdiag(W[type_e])^T . (norm[col_a[e]] * X[col_a[e]])
template <bool RELU, bool HAS_LOOP, bool INDEXED, bool BWD>
__global__ void __launch_bounds__(kStThreads, 1)
rgcn_gather_stream_kernel(const float* __restrict__ X, const int32_t* __restrict__ x_index, const float* __restrict__ W,
                          const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col_a,
                          const int32_t* __restrict__ col_type, const float* __restrict__ norm, float* __restrict__ Hout,
                          int N) {
  extern __shared__ __align__(128) uint8_t st_smem[];
  uint8_t* ring_all = st_smem;
  float* heads = reinterpret_cast<float*>(st_smem + (size_t)kStWarps * kStDepth * kStSlot);        // [kStWarps][200]
  int32_t* s_rp = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(heads) + kStWarps * 800);    // [kStRpCap]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_rp + kStRpCap);                                     // [kStWarps][kStDepth]
  int* flags = reinterpret_cast<int*>(bars + kStWarps * kStDepth);                                   // [kStWarps]
  int* s_part = flags + kStWarps;                                                                   // [2]: A_c, A_c+1

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform for the compiler: ring / barrier addresses in UR
  const bool active = lane < 25;
  const int E = __ldg(row_ptr + N);
  // ---- CTA partition: destinations [A, A_next) own 1/gridDim of the edges (node-aligned) ---------------------------------
  if (warp < 2) {
    const int c = blockIdx.x + warp;
    int a;
    if (c == 0) a = 0;
    else if (c >= (int)gridDim.x) a = N;
    else a = warp_lower_bound(row_ptr, 0, N, (int)(((int64_t)c * E) / gridDim.x), lane);
    if (lane == 0) s_part[warp] = a;
  }
  if (tid < kStWarps) flags[tid] = 0;
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kStDepth; ++k) mbar_init(smem_u32(bars + warp * kStDepth + k), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int A = s_part[0], A_next = s_part[1];
  const int n_rp = A_next - A + 1;
  const bool rp_in_smem = n_rp <= kStRpCap;
  if (rp_in_smem)
    for (int i = tid; i < n_rp; i += kStThreads) s_rp[i] = __ldg(row_ptr + A + i);
  __syncthreads();
  // rp(v): row_ptr[v] for v in [A, A_next]
  auto rp = [&](int v) -> int { return rp_in_smem ? s_rp[v - A] : __ldg(row_ptr + v); };
  const int cb = rp(A), ce = rp(A_next);
  const int chunk = (ce - cb + kStWarps - 1) / kStWarps;
  const int e0 = min(cb + warp * chunk, ce), e1 = min(e0 + chunk, ce);
  const int n = e1 - e0;
  const bool last_warp = warp == kStWarps - 1;

  // ---- edge indices: block b = edges e0 + 32 b .. +31, one per lane; the next block is prefetched --------------------------
  int cur_s = 0, cur_t = 0, nxt_s = 0, nxt_t = 0;
  float cur_sc = 1.f, nxt_sc = 1.f;
  auto load_block = [&](int b, int& s, int& t, float& sc) {
    const int e = e0 + b * 32 + lane;
    s = 0; t = 0; sc = 1.f;
    if (e < e1) {
      s = __ldg(col_a + e);
      t = __ldg(col_type + e);
      if (BWD) sc = __ldg(norm + s);
      if (INDEXED) s = __ldg(x_index + s);
    }
  };
  load_block(0, cur_s, cur_t, cur_sc);
  load_block(1, nxt_s, nxt_t, nxt_sc);
  const uint32_t ring = smem_u32(ring_all + (size_t)warp * kStDepth * kStSlot);
  const uint32_t bar0 = smem_u32(bars + warp * kStDepth);
  int cur_block = 0;
  // copies of local edge k into ring slot `slot` (warp-uniform): the lane that holds its indices broadcasts them, one
  // elected lane issues
  auto issue = [&](int k, int slot) {
    const bool from_next = (k >> 5) != cur_block;
    const int s = __shfl_sync(0xffffffffu, from_next ? nxt_s : cur_s, k & 31);
    const int t = __shfl_sync(0xffffffffu, from_next ? nxt_t : cur_t, k & 31);
    if (st_elect_one()) {
      const uint32_t bar = bar0 + slot * 8, dst = ring + slot * kStSlot;
      st_expect_tx(bar, kStSlot);
      st_bulk_g2s(dst, X + (int64_t)s * 200, 800, bar);
      st_bulk_g2s(dst + 800, W + (int64_t)t * 400, 1600, bar);
    }
  };
#pragma unroll
  for (int k = 0; k < kStDepth; ++k)
    if (k < n) issue(k, k);

  // ---- first destination of the range (binary search in the CTA's row_ptr slice while the first copies fly) -----------------
  int va;
  {
    int lo = A, hi = A_next;           // smallest v in [A, A_next] with rp(v) >= e0
    while (hi > lo) {
      const int mid = (lo + hi) >> 1;
      if (rp(mid) >= e0) hi = mid; else lo = mid + 1;
    }
    va = lo;
  }
  int cur = va;
  bool continued = false;
  if (n > 0 && rp(va) > e0) { cur = va - 1; continued = true; }     // edges e0.. finish a destination an earlier warp started
  int cur_end = cur < A_next ? rp(cur + 1) : ce;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  float2 lp[4];
  float nrm = 1.f;
  auto prefetch_dest = [&](int v) {    // self-loop row (already in Hout) and norm of destination v
    if (v < A_next) {
      if (HAS_LOOP && active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) lp[k] = *reinterpret_cast<const float2*>(Hout + (int64_t)v * 200 + 2 * (lane + 25 * k));
      }
      if (!BWD) nrm = __ldg(norm + v);
    }
  };
  auto epilogue = [&](int v) {         // registers -> global, fused norm / self-loop / activation
    if (active) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float2 o = make_float2(acc[2 * k], acc[2 * k + 1]);
        if (!BWD) { o.x *= nrm; o.y *= nrm; }
        if (HAS_LOOP) { o.x += lp[k].x; o.y += lp[k].y; }
        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
        *reinterpret_cast<float2*>(Hout + (int64_t)v * 200 + 2 * (lane + 25 * k)) = o;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  };
  auto publish_head = [&]() {          // partial sum of a destination an earlier warp started
    if (active) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        *reinterpret_cast<float2*>(heads + warp * 200 + 2 * (lane + 25 * k)) = make_float2(acc[2 * k], acc[2 * k + 1]);
    }
    __syncwarp();
    if (lane == 0) st_release_cta(flags + warp, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  };
  // destination `cur` has no more edges in this warp's range and all of its edges end at or before e1
  auto finish = [&]() {
    if (continued) { publish_head(); continued = false; }
    else epilogue(cur);                // covers destinations without edges too (acc = 0)
  };
  if (!continued) prefetch_dest(cur);

  const uint8_t* my_ring = ring_all + (size_t)warp * kStDepth * kStSlot;
  for (int g = 0; g < n; g += kStDepth) {      // one pass over the ring: slot numbers are compile-time constants
    const uint32_t parity = (uint32_t)(g / kStDepth) & 1u;
    if ((g & 31) == 0 && g > 0) {
      cur_s = nxt_s; cur_t = nxt_t; cur_sc = nxt_sc;
      ++cur_block;
      load_block(cur_block + 1, nxt_s, nxt_t, nxt_sc);
    }
#pragma unroll
    for (int slot = 0; slot < kStDepth; ++slot) {
      const int i = g + slot;
      if (i < n) {
        const int e = e0 + i;
        while (e >= cur_end) {         // warp-uniform: the running destination is complete
          finish();
          ++cur;
          cur_end = rp(cur + 1);
          prefetch_dest(cur);
        }
        st_wait(bar0 + slot * 8, parity);
        const uint8_t* sp = my_ring + slot * kStSlot;
        const float sc = BWD ? __shfl_sync(0xffffffffu, cur_sc, i & 31) : 1.f;
        if (active) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 h = *reinterpret_cast<const float2*>(sp + 8 * (lane + 25 * k));
            const float4 w = *reinterpret_cast<const float4*>(sp + 800 + 16 * (lane + 25 * k));
            const float x = BWD ? h.x * sc : h.x, y = BWD ? h.y * sc : h.y;
            if (!BWD) {                // out[j] += sum_i in[i] * W[i][j]
              acc[2 * k] = fmaf(x, w.x, fmaf(y, w.z, acc[2 * k]));
              acc[2 * k + 1] = fmaf(x, w.y, fmaf(y, w.w, acc[2 * k + 1]));
            } else {                   // din[i] += sum_j W[i][j] * g[j]
              acc[2 * k] = fmaf(x, w.x, fmaf(y, w.y, acc[2 * k]));
              acc[2 * k + 1] = fmaf(x, w.z, fmaf(y, w.w, acc[2 * k + 1]));
            }
          }
        }
        __syncwarp();                  // every lane has consumed the slot (the FMAs depend on the loads)
        if (i + kStDepth < n) issue(i + kStDepth, slot);
      }
    }
  }
  // ---- end of the range ------------------------------------------------------------------------------------------------------
  if (n > 0) {
    if (cur_end <= e1) {
      finish();                        // the running destination ends exactly here
      ++cur;
    } else if (continued) {
      publish_head();                  // the whole range lies inside one destination started earlier and finished later
      ++cur;
    } else {
      // this warp started `cur`; later warps hold the rest of its edges: add their heads in warp (= edge) order
      const int k_last = (cur_end - 1 - cb) / chunk;
      for (int k = warp + 1; k <= k_last; ++k) {
        if (lane == 0) {
          int spins = 0;
          while (ld_acquire_cta(flags + k) == 0) {
            if (++spins > (1 << 24)) __trap();
          }
        }
        __syncwarp();
        if (active) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 hv = *reinterpret_cast<const float2*>(heads + k * 200 + 2 * (lane + 25 * q));
            acc[2 * q] += hv.x; acc[2 * q + 1] += hv.y;
          }
        }
      }
      epilogue(cur);
      ++cur;
    }
  }
  // destinations without edges at the very end of the CTA's range (only the last CTA can have them)
  if (last_warp) {
    if (n == 0) cur = va;
    for (; cur < A_next; ++cur) {
      prefetch_dest(cur);
      epilogue(cur);
    }
  }
}

}  // namespace renet
