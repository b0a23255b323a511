// Feature-sliced persistent gather for the RE-Net shape (d_in = d_out = 200, 100 blocks of 2x2): the kernel behind
// renet_rgcn_gather at batch scale (reference RGCN.py:79-94 + 42-48; DGL fn.sum, RGCN.py:91).
//
// Why: the round-1 tile kernel (rgcn_tile.cuh) re-fetched each edge's 1600-byte relation row through L1/L2
// (35 % L1 hit rate): 320 MB of L2->SM traffic per launch against 212 MB of algorithmic bytes, at ~6.5 TB/s -- the L2
// slices' limit, not HBM's.  The relation table (R2 x 400 floats = 819 KB) does not fit in shared memory, but ONE FIFTH
// of its columns does: 20 of the 100 blocks x R2 = 512 relations = 160 KB.  So the feature dimension is cut into 5 slices
// of 40 floats (5 full 32-byte sectors of every 800-byte row) and each persistent CTA (one per SM) works on ONE slice:
//   * its slice of the whole relation table is staged ONCE into shared memory by TMA (cp.async.bulk.tensor.4d through a
//     tensor map that views W [R2, 400] as [R2][50 block pairs][2][4 floats]: a box of 32 relations x 10 pairs x one
//     parity lands as a dense plane, so the even and the odd blocks of a pair sit in two planes and both per-edge weight
//     reads are conflict-free LDS.128), completing on an mbarrier;
//   * per edge and slice only 160 bytes of the source row cross L2->SM (one LDG.128 per lane of a 10-lane group), so
//     the traffic is E*800 + N*1600 + indices -- what the roofline formula counts;
//   * a warp = 3 groups of 10 lanes; a warp owns a contiguous, node-aligned range of destinations whose weight
//     (edges + 2 per node) is 1/(#warps of the slice) of the graph (found by a 16-ary warp search in row_ptr during the
//     TMA staging), and the three groups split the warp's EDGE range evenly, so heavy destinations never idle lanes;
//   * each group walks its edges with 8 row loads in flight (double-buffered blocks) and keeps the running destination's
//     sum in registers (edges are destination-sorted: a segmented reduction); when the destination changes the sum goes
//     to a small per-group buffer in shared memory (12 instructions, no global access: the next row_ptr entries ride in
//     a register window filled 10 destinations ahead), and at the end of every block the whole warp drains the buffers
//     together -- self-loop row and norm loads batched, norm / self-loop / ReLU, 160-byte coalesced stores.  (The first
//     version finished each destination inside the divergent flush through a cp.async ring: 100 instructions at 10
//     active lanes, 17 M of the kernel's 33 M warp instructions and 137 KB of unrolled code.)
//   * a destination cut by a group boundary is summed in group order from registers + per-warp head slots: no atomics
//     of any kind, the result is bitwise reproducible.
// The same body is the backward dH kernel (BWD: reversed CSR, transposed blocks, per-edge scale norm[dst]).
#pragma once
#include <cuda.h>      // CUtensorMap + enums only; the encoder is fetched through cudaGetDriverEntryPoint (no libcuda link)

#include "common.cuh"
#include "umma.cuh"

namespace renet {

constexpr int kSlFeat = 40;          // features per slice (20 blocks of 2x2): 160 B = 5 sectors of a row
constexpr int kSlices = 5;
constexpr int kSlWarps = 16;
constexpr int kSlThreads = kSlWarps * 32;
constexpr int kSlKB = 8;             // edges per index block (row loads in flight per group: one block)
constexpr int kSlDone = 4;           // finished destinations a group buffers before the warp drains them
constexpr int kSlDoneWords = 44;     // 40 sums + destination id + pad (176 B)
constexpr int kSlWin = 10;           // row_ptr entries per register window (one per lane of a group)
constexpr int kSlNodeCost = 2;       // a destination costs about two edges
constexpr int kSlBoxRows = 32;       // relations per TMA box
constexpr int kSlMaxR2 = 576;

inline int sliced_w_rows(int R2) { return (R2 + kSlBoxRows - 1) / kSlBoxRows * kSlBoxRows; }
inline size_t sliced_smem_bytes(int R2) {
  return (size_t)sliced_w_rows(R2) * 320 + (size_t)kSlWarps * 3 * kSlDone * kSlDoneWords * 4 +
         (size_t)kSlWarps * 2 * kSlFeat * 4 + 16;
}

// Tensor map over the relation table W [R2, 400] fp32 (RGCN.py:75-77 layout: 100 blocks of 4 floats per relation) viewed
// as [R2][50 pairs][2 parities][4 floats], box = 32 relations x 10 pairs x 1 parity x 4 floats: one TMA op moves one parity
// plane of 32 relations' columns of one slice.  Pure host computation (no allocation, nothing retained).
inline int sliced_make_tmap(const float* W, int R2, CUtensorMap* tm) {
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || p == nullptr) {
      set_error("cuTensorMapEncodeTiled is not available from this driver");
      return RENET_ERR_CUDA;
    }
    fn = reinterpret_cast<EncodeFn>(p);
  }
  const cuuint64_t dims[4] = {4, 2, 50, (cuuint64_t)R2};
  const cuuint64_t strides[3] = {16, 32, 1600};
  const cuuint32_t box[4] = {4, 1, 10, (cuuint32_t)kSlBoxRows};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(W), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return RENET_ERR_CUDA;
  }
  return RENET_OK;
}

// smallest p in [0, N] with row_ptr[p] + kSlNodeCost * p >= T, for two targets at once (lanes 0-15: T0, 16-31: T1)
__device__ __forceinline__ void warp_lower_bound2(const int32_t* __restrict__ row_ptr, int N, int64_t T0, int64_t T1,
                                                  int& r0, int& r1) {
  const int lane = threadIdx.x & 31, half = lane >> 4, k = lane & 15;
  const int64_t T = half ? T1 : T0;
  int lo = 0, hi = N;
  while (__any_sync(0xffffffffu, hi > lo)) {
    const int span = hi - lo;
    const int p = lo + (int)(((int64_t)span * (k + 1)) / 17);
    const bool ge = span > 0 && ((int64_t)__ldg(row_ptr + p) + (int64_t)kSlNodeCost * p >= T);
    const unsigned mh = (__ballot_sync(0xffffffffu, ge) >> (16 * half)) & 0xffffu;
    if (span > 0) {
      if (mh == 0) {
        lo = lo + (int)(((int64_t)span * 16) / 17) + 1;
      } else {
        const int kk = __ffs(mh) - 1;
        const int nh = lo + (int)(((int64_t)span * (kk + 1)) / 17);
        if (kk > 0) lo = lo + (int)(((int64_t)span * kk) / 17) + 1;
        hi = nh;
      }
    }
  }
  r0 = __shfl_sync(0xffffffffu, lo, 0);
  r1 = __shfl_sync(0xffffffffu, lo, 16);
}

struct SlIdx {
  int ci, ct;
  float sc;
};

template <bool RELU, bool HAS_LOOP, bool INDEXED, bool BWD>
__global__ void __launch_bounds__(kSlThreads, 1)
rgcn_gather_sliced_kernel(const float* __restrict__ X, const int32_t* __restrict__ x_index,
                          const __grid_constant__ CUtensorMap w_map, const int32_t* __restrict__ row_ptr,
                          const int32_t* __restrict__ col_a, const int32_t* __restrict__ col_type,
                          const float* __restrict__ norm, float* __restrict__ Out, int N, int R2) {
  extern __shared__ __align__(128) uint8_t sl_smem[];
  const int w_rows = (R2 + kSlBoxRows - 1) / kSlBoxRows * kSlBoxRows;
  float* W0 = reinterpret_cast<float*>(sl_smem);            // [w_rows][10 pairs][4]: even blocks of the slice
  float* W1 = W0 + (size_t)w_rows * 40;                     // odd blocks
  float* done_all = W1 + (size_t)w_rows * 40;
  float* head_all = done_all + kSlWarps * 3 * kSlDone * kSlDoneWords;
  uint64_t* bar = reinterpret_cast<uint64_t*>(head_all + kSlWarps * 2 * kSlFeat);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int slice = blockIdx.x % kSlices;
  const int cta_in_slice = blockIdx.x / kSlices;
  const int ncta_slice = ((int)gridDim.x - slice + kSlices - 1) / kSlices;
  const int nworkers = ncta_slice * kSlWarps;
  const int worker = cta_in_slice * kSlWarps + warp;
  const int soff = slice * kSlFeat;

  // ---- 1. this slice of the relation table -> shared memory: TMA boxes of 32 relations x 10 block pairs x one parity ----
  const uint32_t bar_a = smem_u32(bar);
  if (tid == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    const int nbox = w_rows / kSlBoxRows;
    if (lane == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)w_rows * 320u) : "memory");
    __syncwarp();
    for (int i = lane; i < 2 * nbox; i += 32) {             // rows past R2 in the last box are zero-filled (and never read)
      const int par = i & 1, bx = i >> 1;
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
              smem_u32((par ? W1 : W0) + bx * kSlBoxRows * 40)),
          "l"(reinterpret_cast<uint64_t>(&w_map)), "r"(0), "r"(par), "r"(slice * 10), "r"(bx * kSlBoxRows), "r"(bar_a)
          : "memory");
    }
  }

  // ---- 2. this warp's destinations (overlaps the staging) -------------------------------------------------------------------
  const int E = __ldg(row_ptr + N);
  const int64_t Wt = (int64_t)E + (int64_t)kSlNodeCost * N;
  int va, vb;
  warp_lower_bound2(row_ptr, N, (Wt * worker) / nworkers, (Wt * (worker + 1)) / nworkers, va, vb);
  const int ea = __ldg(row_ptr + va), eb = __ldg(row_ptr + vb);
  const int len = eb - ea;
  const int g = lane / 10, j = lane - g * 10, gbase = g * 10;
  const bool lane_on = lane < 30;
  const unsigned gmask = lane_on ? (0x3ffu << gbase) : 0xc0000000u;
  auto cut = [&](int gg) { return ea + (int)(((int64_t)len * gg) / 3); };
  const int ge0 = lane_on ? cut(g) : eb, ge1 = lane_on ? cut(g + 1) : eb;
  const int glen = ge1 - ge0;
  const int first_ne = len >= 3 ? 0 : (len == 2 ? 1 : (len == 1 ? 2 : 0));
  // first destination of every group: the warp's first one for the first non-empty group, else the one holding edge ge0
  int node = va;
  {
    const int x1 = cut(1), x2 = cut(2);
    for (int base = va; base < vb; base += 32) {
      const int v = base + lane;
      const int p0 = v < vb ? __ldg(row_ptr + v) : 0x7fffffff, p1 = v < vb ? __ldg(row_ptr + v + 1) : 0x7fffffff;
      const unsigned m1 = __ballot_sync(0xffffffffu, p0 <= x1 && x1 < p1);
      const unsigned m2 = __ballot_sync(0xffffffffu, p0 <= x2 && x2 < p1);
      if (m1 && g == 1 && first_ne < 1) node = base + __ffs(m1) - 1;
      if (m2 && g == 2 && first_ne < 2) node = base + __ffs(m2) - 1;
    }
  }
  const bool walks = lane_on && (glen > 0 || (len == 0 && g == 0 && vb > va));
  bool continued = lane_on && glen > 0 && g > first_ne && __ldg(row_ptr + node) < ge0;
  // register window over row_ptr: lane j of the group holds the end of destination wbase + j (and of wbase + 10 + j)
  int wbase = node;
  int win = __ldg(row_ptr + min(wbase + 1 + j, N)), win2 = __ldg(row_ptr + min(wbase + 1 + kSlWin + j, N));
  int node_end = __shfl_sync(0xffffffffu, win, gbase);
  int node_beg = walks ? __ldg(row_ptr + node) : 0;         // first edge of `node`
  if (!walks) node_end = 0x7fffffff;

  float* done = done_all + (size_t)(warp * 3 + (lane_on ? g : 0)) * kSlDone * kSlDoneWords;   // this group's finished sums
  float* heads = head_all + (size_t)warp * 2 * kSlFeat;     // [2][40]: partial sums of groups 1 and 2 for a destination they continue
  int cnt = 0;                                              // entries in `done`
  float2 accA = make_float2(0.f, 0.f), accB = make_float2(0.f, 0.f);     // blocks 2j and 2j+1 of the slice

  auto finalize = [&](int n, float4 a, float4 lp, float nv) {
    float4 o;
    if (BWD) {
      o = make_float4(a.x + lp.x, a.y + lp.y, a.z + lp.z, a.w + lp.w);
    } else {
      o = make_float4(fmaf(a.x, nv, lp.x), fmaf(a.y, nv, lp.y), fmaf(a.z, nv, lp.z), fmaf(a.w, nv, lp.w));
      if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    }
    st_f4(Out + (size_t)n * 200 + soff + 4 * j, o);
  };
  // the group's buffered destinations: epilogue inputs loaded as one batch, then norm / self-loop / activation / store
  auto drain = [&]() {
    __syncwarp(gmask);
#pragma unroll
    for (int i0 = 0; i0 < kSlDone; i0 += 2) {               // two at a time: bounded register footprint
      float4 lp[2];
      float nv[2];
      int nn[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        lp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        nv[i] = 1.f;
        nn[i] = 0;
        if (i0 + i < cnt) {
          nn[i] = __float_as_int(done[(i0 + i) * kSlDoneWords + 40]);
          if (HAS_LOOP) lp[i] = *reinterpret_cast<const float4*>(Out + (size_t)nn[i] * 200 + soff + 4 * j);
          if (!BWD) nv[i] = __ldg(norm + nn[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
        if (i0 + i < cnt) finalize(nn[i], *reinterpret_cast<const float4*>(done + (i0 + i) * kSlDoneWords + 4 * j), lp[i], nv[i]);
    }
    __syncwarp(gmask);
    cnt = 0;
  };
  // destination `node` is complete as far as this group is concerned: park the sum (or hand the partial sum over), move on
  auto flush = [&]() {
    const float4 a = make_float4(accA.x, accA.y, accB.x, accB.y);
    if (continued) {
      *reinterpret_cast<float4*>(heads + (g - 1) * kSlFeat + 4 * j) = a;
      continued = false;
    } else {
      if (cnt == kSlDone) {                                 // rare (more than 4 destinations ended inside one block): finish it here
        const float4 lp = HAS_LOOP ? *reinterpret_cast<const float4*>(Out + (size_t)node * 200 + soff + 4 * j)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        finalize(node, a, lp, BWD ? 1.f : __ldg(norm + node));
      } else {
        *reinterpret_cast<float4*>(done + cnt * kSlDoneWords + 4 * j) = a;
        if (j == 0) done[cnt * kSlDoneWords + 40] = __int_as_float(node);
        ++cnt;
      }
    }
    accA = make_float2(0.f, 0.f); accB = make_float2(0.f, 0.f);
    node_beg = node_end;
    ++node;
    int wi = node - wbase;
    if (wi == kSlWin) {
      win = win2;
      wbase += kSlWin;
      win2 = __ldg(row_ptr + min(wbase + 1 + kSlWin + j, N));
      wi = 0;
    }
    node_end = __shfl_sync(gmask, win, gbase + wi);
  };

  // ---- 3. wait for the relation table ---------------------------------------------------------------------------------------------
  mbar_wait(bar_a, 0);

  // ---- 4. edge walk: blocks of kSlKB edges, indices three blocks ahead, source rows one block ahead --------------
  const int nblk = ((len + 2) / 3 + kSlKB - 1) / kSlKB;
  auto load_idx = [&](int b) {
    SlIdx r{0, 0, 1.f};
    const int ee = ge0 + b * kSlKB + j;
    if (j < kSlKB && ee < ge1) { r.ci = __ldg(col_a + ee); r.ct = __ldg(col_type + ee); }
    return r;
  };
  auto xform = [&](SlIdx& r, int b) {
    const int ee = ge0 + b * kSlKB + j;
    if ((INDEXED || BWD) && j < kSlKB && ee < ge1) {
      if (BWD) r.sc = __ldg(norm + r.ci);
      if (INDEXED) r.ci = __ldg(x_index + r.ci);
    }
  };
  auto issue = [&](const SlIdx& r, int b, float4 (&hb)[kSlKB]) {
#pragma unroll
    for (int k = 0; k < kSlKB; ++k) {
      const int s = __shfl_sync(0xffffffffu, r.ci, gbase + k);
      hb[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ge0 + b * kSlKB + k < ge1) hb[k] = ldg_f4_stream(X + (size_t)s * 200 + soff + 4 * j);
    }
  };
  auto compute = [&](const SlIdx& r, int b, const float4 (&hb)[kSlKB]) {
#pragma unroll
    for (int k = 0; k < kSlKB; ++k) {
      const int ee = ge0 + b * kSlKB + k;
      const int tk = __shfl_sync(0xffffffffu, r.ct, gbase + k);
      const float sk = BWD ? __shfl_sync(0xffffffffu, r.sc, gbase + k) : 1.f;
      if (ee < ge1) {
        while (ee == node_end) flush();
        const float4 wa = *reinterpret_cast<const float4*>(W0 + tk * 40 + 4 * j);
        const float4 wb = *reinterpret_cast<const float4*>(W1 + tk * 40 + 4 * j);
        float4 h = hb[k];
        if (BWD) { h.x *= sk; h.y *= sk; h.z *= sk; h.w *= sk; }
        if (!BWD) {      // out[jj] += sum_i in[i] * W[i][jj];  block = (W00 W01; W10 W11) row-major
          accA.x = fmaf(h.x, wa.x, fmaf(h.y, wa.z, accA.x));
          accA.y = fmaf(h.x, wa.y, fmaf(h.y, wa.w, accA.y));
          accB.x = fmaf(h.z, wb.x, fmaf(h.w, wb.z, accB.x));
          accB.y = fmaf(h.z, wb.y, fmaf(h.w, wb.w, accB.y));
        } else {         // din[i] += sum_jj W[i][jj] * g[jj]
          accA.x = fmaf(h.x, wa.x, fmaf(h.y, wa.y, accA.x));
          accA.y = fmaf(h.x, wa.z, fmaf(h.y, wa.w, accA.y));
          accB.x = fmaf(h.z, wb.x, fmaf(h.w, wb.y, accB.x));
          accB.y = fmaf(h.z, wb.z, fmaf(h.w, wb.w, accB.y));
        }
      }
    }
    drain();                                                // converged: the three groups finish their destinations together
  };
  float4 hA[kSlKB], hB[kSlKB];
  SlIdx i0 = load_idx(0), i1 = load_idx(1), i2 = load_idx(2), i3;
  xform(i0, 0);
  xform(i1, 1);
  issue(i0, 0, hA);
  for (int b = 0; b < nblk; b += 2) {
    i3 = load_idx(b + 3);
    xform(i2, b + 2);
    issue(i1, b + 1, hB);
    compute(i0, b, hA);
    i0 = i1; i1 = i2; i2 = i3;
    if (b + 1 < nblk) {
      i3 = load_idx(b + 4);
      xform(i2, b + 3);
      issue(i1, b + 2, hA);
      compute(i0, b + 1, hB);
      i0 = i1; i1 = i2; i2 = i3;
    }
  }

  // ---- 5. close the group's range: complete destinations (and trailing edge-less ones); a destination that goes on into the
  //         next group keeps its partial sum in registers (its starter) or in the group's head slot (a continuation) -----------
  bool starter = false;
  float4 mine = make_float4(0.f, 0.f, 0.f, 0.f);
  if (walks) {
    while (node < vb && node_end <= ge1) flush();
    if (glen > 0 && node < vb && node_beg < ge1) {          // this group holds edges of a destination that goes on into g+1
      mine = make_float4(accA.x, accA.y, accB.x, accB.y);
      if (continued) *reinterpret_cast<float4*>(heads + (g - 1) * kSlFeat + 4 * j) = mine;
      else starter = true;
    }
  }
  __syncwarp();
  drain();
  if (starter) {                                            // sum in group order: own tail + head of g+1 (+ head of g+2)
    const float4 h1 = *reinterpret_cast<const float4*>(heads + g * kSlFeat + 4 * j);
    mine.x += h1.x; mine.y += h1.y; mine.z += h1.z; mine.w += h1.w;
    if (g == 0 && node_end > cut(2)) {
      const float4 h2 = *reinterpret_cast<const float4*>(heads + kSlFeat + 4 * j);
      mine.x += h2.x; mine.y += h2.y; mine.z += h2.z; mine.w += h2.w;
    }
    const float4 lp = HAS_LOOP ? *reinterpret_cast<const float4*>(Out + (size_t)node * 200 + soff + 4 * j)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
    finalize(node, mine, lp, BWD ? 1.f : __ldg(norm + node));
  }
}

}  // namespace renet
