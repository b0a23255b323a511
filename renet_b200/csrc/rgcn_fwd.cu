// Fused RGCN block-diagonal message passing, forward (reference RGCN.py:79-94 + 42-48):
//
//   Hout[v] = act( norm[v] * sum_{e: dst(e)=v} blockdiag(W[type_e]) . Hin[src_e]  +  loop[v] )
//
// replaces the reference's index_select (E x 400 weight materialisation, RGCN.py:81-85), the
// E*100 tiny bmm (RGCN.py:86-87), DGL's atomic copy-reduce (fn.sum, RGCN.py:91), the norm multiply
// (RGCN.py:93-94), the self-loop add and the activation (RGCN.py:45-48) with ONE pass over a
// destination-sorted CSR: no per-edge message ever reaches memory and no atomics are needed.
//
// Fast path (d_in = d_out = 200, num_bases = 100, 2x2 blocks -- the only shape RE-Net uses,
// model.py:36): see rgcn_tile.cuh.  One CTA = 16 destination rows, edges split evenly over 8 warps,
// warp-level segmented reduction into a shared tile, fused norm / self-loop / activation epilogue.
#include <stdlib.h>

#include "common.cuh"
#include "rgcn_tile.cuh"
#include "rgcn_comp.cuh"
#include "rgcn_ring.cuh"
#include "rgcn_hot.cuh"
#include <mutex>
#include <vector>

namespace renet {
namespace {

// VARIANT 0: one tile per CTA, source rows streamed past L1.  1: source rows allocate in L1.  2: persistent
// CTAs, each walking a contiguous range of tiles (same component -> source rows re-used from L1).
template <bool RELU, bool HAS_LOOP, bool INDEXED, int NODES = kTileNodes, int VARIANT = 0>
__global__ void __launch_bounds__(kTileWarps * 32, 768 / (kTileWarps * 32))
rgcn_gather_d200_kernel(const float* __restrict__ H, const int32_t* __restrict__ h_index,
                        const float* __restrict__ W, const int32_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ col_src, const int32_t* __restrict__ col_type,
                        const float* __restrict__ norm, float* __restrict__ Hout, int N, int passthrough,
                        const int32_t* __restrict__ tile_order) {
  __shared__ __align__(16) float agg[NODES][200];
  __shared__ __align__(16) float loopbuf[HAS_LOOP ? NODES : 1][200];
  constexpr bool DET = (VARIANT == 0);     // deterministic, atomic-free hand-over of partial sums (rgcn_tile.cuh)
  __shared__ __align__(16) float head[DET ? kTileWarps : 1][200];
  __shared__ int head_mask[NODES];
  __shared__ float normbuf[NODES];
  __shared__ int s_rp[NODES + 1];
  const int tid = threadIdx.x;
  const int n_tiles = (N + NODES - 1) / NODES;
  // tile_order (optional): heaviest tiles first, so the last wave of CTAs is the lightest (LPT scheduling)
  int tile_lo = tile_order ? __ldg(tile_order + blockIdx.x) : blockIdx.x, tile_hi = tile_lo + 1;
  if (VARIANT == 2) {
    const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
    tile_lo = blockIdx.x * per;
    tile_hi = min(n_tiles, tile_lo + per);
  }
  for (int tile = tile_lo; tile < tile_hi; ++tile) {
  if (tile > tile_lo) __syncthreads();
  const int v0 = tile * NODES;
  const int nv = min(NODES, N - v0);
  tile_prefetch_epilogue(loopbuf, normbuf, Hout + (int64_t)v0 * 200, norm + v0, nv, HAS_LOOP, tid, kTileWarps * 32);
  if (DET) {
    if (tid < NODES) head_mask[tid] = 0;
  } else {
    for (int i = tid; i < NODES * 200; i += kTileWarps * 32) (&agg[0][0])[i] = 0.f;
  }
  if (tid <= nv) s_rp[tid] = __ldg(row_ptr + v0 + tid);
  __syncthreads();
  const TileHeads th{head, head_mask};
  tile_accumulate<false, INDEXED, false, VARIANT == 0, DET>(agg, s_rp, nv, H, h_index, W, col_src, col_type, nullptr, th);
  cp_async_wait_all();
  __syncthreads();
  // epilogue: nv rows x 100 float2, coalesced; self-loop rows and norms were prefetched into shared memory
  for (int i = tid; i < nv * 100; i += kTileWarps * 32) {
    const int r = i / 100, c = (i % 100) * 2;
    const int v = v0 + r;
    float2 a = DET ? tile_row_sum(agg, th, s_rp, r, c) : *reinterpret_cast<const float2*>(&agg[r][c]);
    if (passthrough) {  // graph without edges: DGL 0.4 skips the reduce, h is left as is
      const int64_t hr = INDEXED ? (int64_t)__ldg(h_index + v) : v;
      a = *reinterpret_cast<const float2*>(H + hr * 200 + c);
    }
    const float nvv = normbuf[r];
    float* op = Hout + (int64_t)v * 200 + c;
    float2 o = make_float2(a.x * nvv, a.y * nvv);
    if (HAS_LOOP) {
      const float2 l = *reinterpret_cast<const float2*>(&loopbuf[r][c]);
      o.x += l.x; o.y += l.y;
    }
    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
    *reinterpret_cast<float2*>(op) = o;
  }
  }
}

// Component-resident version (rgcn_comp.cuh): one CTA per component, features and hot relation rows staged
// in shared memory once, two 8-warp groups walking 16-destination tiles.
template <bool RELU, bool HAS_LOOP, bool INDEXED>
__global__ void __launch_bounds__(kCompThreads, 1)
rgcn_gather_comp_kernel(const float* __restrict__ H, const int32_t* __restrict__ h_index,
                        const float* __restrict__ W, const int32_t* __restrict__ row_ptr,
                        const int32_t* __restrict__ col_src, const int32_t* __restrict__ col_type,
                        const float* __restrict__ norm, float* __restrict__ Hout,
                        const int32_t* __restrict__ comp_ptr, const int32_t* __restrict__ comp_order,
                        const int32_t* __restrict__ rel_slot, const int32_t* __restrict__ hot_rel, int n_hot) {
  extern __shared__ __align__(16) float csm[];
  float* win = csm;
  float* wc = win + kWinRows * 200;
  float(*agg)[kTileNodes][200] = reinterpret_cast<float(*)[kTileNodes][200]>(wc + kHotRel * 400);
  float(*loopb)[kTileNodes][200] = agg + kCompGroups;
  float* normb_all = reinterpret_cast<float*>(loopb + kCompGroups);
  int* s_rp_all = reinterpret_cast<int*>(normb_all + kCompGroups * kTileNodes);
  const int tid = threadIdx.x;
  const int comp = comp_order != nullptr ? __ldg(comp_order + blockIdx.x) : blockIdx.x;
  const int v_lo = __ldg(comp_ptr + comp), v_hi = __ldg(comp_ptr + comp + 1);
  const int win_n = min(kWinRows, v_hi - v_lo);
  comp_stage<INDEXED>(win, wc, H, h_index, W, hot_rel, n_hot, v_lo, win_n);
  __syncthreads();
  const int group = tid / (kTileWarps * 32), gtid = tid % (kTileWarps * 32), gwarp = gtid >> 5;
  float(*my_agg)[200] = agg[group];
  int* s_rp = s_rp_all + group * (kTileNodes + 1);
  const int n_tiles = (v_hi - v_lo + kTileNodes - 1) / kTileNodes;
  for (int tile = group; tile < n_tiles; tile += kCompGroups) {
    const int v0 = v_lo + tile * kTileNodes;
    const int nv = min(kTileNodes, v_hi - v0);
    tile_prefetch_epilogue(loopb[group], normb_all + group * kTileNodes, Hout + (int64_t)v0 * 200, norm + v0, nv,
                           HAS_LOOP, gtid, kTileWarps * 32);
    for (int i = gtid; i < kTileNodes * 200; i += kTileWarps * 32) (&my_agg[0][0])[i] = 0.f;
    if (gtid <= nv) s_rp[gtid] = __ldg(row_ptr + v0 + gtid);
    group_barrier(group);
    comp_tile_accumulate<false, INDEXED, false>(my_agg, s_rp, nv, gwarp, H, h_index, W, col_src, col_type, nullptr,
                                                win, v_lo, win_n, wc, n_hot > 0 ? rel_slot : nullptr);
    cp_async_wait_all();
    group_barrier(group);
    for (int i = gtid; i < nv * 100; i += kTileWarps * 32) {
      const int r = i / 100, c = (i % 100) * 2;
      const int v = v0 + r;
      const float2 a = *reinterpret_cast<const float2*>(&my_agg[r][c]);
      const float nvv = normb_all[group * kTileNodes + r];
      float* op = Hout + (int64_t)v * 200 + c;
      float2 o = make_float2(a.x * nvv, a.y * nvv);
      if (HAS_LOOP) {
        const float2 l = *reinterpret_cast<const float2*>(&loopb[group][r][c]);
        o.x += l.x; o.y += l.y;
      }
      if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); }
      *reinterpret_cast<float2*>(op) = o;
    }
    group_barrier(group);
  }
}

// Generic shapes (any d_in, d_out, num_bases): one thread per (node, output feature).  Only the
// small known-answer cases use it; RE-Net itself always runs the d200 path.
__global__ void rgcn_gather_generic_kernel(const float* __restrict__ H, const int32_t* __restrict__ h_index,
                                           const float* __restrict__ W, const int32_t* __restrict__ row_ptr,
                                           const int32_t* __restrict__ col_src,
                                           const int32_t* __restrict__ col_type, const float* __restrict__ norm,
                                           float* __restrict__ Hout, int64_t N, int d_in, int d_out, int nb,
                                           int relu, int has_loop, int passthrough) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * d_out) return;
  const int64_t v = i / d_out;
  const int c = (int)(i % d_out);
  const int si = d_in / nb, so = d_out / nb;
  const int b = c / so, j = c % so;
  float acc = 0.f;
  for (int e = row_ptr[v]; e < row_ptr[v + 1]; ++e) {
    int64_t s = col_src[e];
    if (h_index) s = h_index[s];
    const float* w = W + (int64_t)col_type[e] * nb * si * so + (int64_t)b * si * so + j;
    const float* h = H + s * d_in + b * si;
    for (int k = 0; k < si; ++k) acc = fmaf(h[k], w[k * so], acc);
  }
  if (passthrough) {
    const int64_t r = h_index ? (int64_t)h_index[v] : v;
    acc = H[r * d_in + c];
  }
  float o = acc * norm[v];
  if (has_loop) o += Hout[i];
  if (relu) o = fmaxf(o, 0.f);
  Hout[i] = o;
}

}  // namespace

// experiment knob (RENET_GATHER_VARIANT / renet_set_gather_variant) for the d=200 forward gather:
//   0 default tile kernel (deterministic hand-over, source rows streamed past L1)
//   1 source rows allocate in L1, shared-memory atomics     2 persistent CTAs over contiguous tile ranges
//   6 per-warp cp.async ring of source rows (rgcn_ring.cuh) 7 persistent CTAs with the hot relation rows in shared
//     memory (rgcn_hot.cuh, needs renet_set_hot_relations)
// All measured slower than 0; 0, 6 and 7 are bit-identical.
static int g_gather_variant = -1;
int gather_variant() {
  if (g_gather_variant < 0) {
    const char* e = getenv("RENET_GATHER_VARIANT");
    g_gather_variant = e ? atoi(e) : 0;
  }
  return g_gather_variant;
}
int set_gather_variant(int v) {
  const int prev = gather_variant();
  g_gather_variant = v;
  return prev;
}

// ---- hot relations (renet_set_hot_relations): a performance hint for the persistent gather (rgcn_hot.cuh) ------------
namespace {
struct HotSet { int device; int R2; int n_hot; int32_t* rel_slot; int32_t* hot_rel; };
std::mutex g_hot_mu;
std::vector<int32_t> g_hot_host;     // the hint as given; uploaded lazily per device
int g_hot_R2 = 0;
std::vector<HotSet> g_hot_dev;

// device copy of the hint for the current device, or nullptr
const HotSet* hot_set_for_device(int R2) {
  std::lock_guard<std::mutex> lk(g_hot_mu);
  if (g_hot_host.empty() || R2 != g_hot_R2) return nullptr;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  for (const auto& h : g_hot_dev) if (h.device == dev) return &h;
  HotSet h{dev, g_hot_R2, (int)g_hot_host.size(), nullptr, nullptr};
  std::vector<int32_t> slot(g_hot_R2, -1);
  for (int i = 0; i < h.n_hot; ++i) slot[g_hot_host[i]] = i;
  if (cudaMalloc(&h.rel_slot, sizeof(int32_t) * g_hot_R2) != cudaSuccess ||
      cudaMalloc(&h.hot_rel, sizeof(int32_t) * h.n_hot) != cudaSuccess ||
      cudaMemcpy(h.rel_slot, slot.data(), sizeof(int32_t) * g_hot_R2, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(h.hot_rel, g_hot_host.data(), sizeof(int32_t) * h.n_hot, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  g_hot_dev.push_back(h);
  return &g_hot_dev.back();
}
}  // namespace

int set_hot_relations(const int32_t* hot_rel_host, int n_hot, int R2) {
  std::lock_guard<std::mutex> lk(g_hot_mu);
  for (auto& h : g_hot_dev) { cudaFree(h.rel_slot); cudaFree(h.hot_rel); }
  g_hot_dev.clear();
  g_hot_host.clear();
  g_hot_R2 = 0;
  if (n_hot <= 0 || hot_rel_host == nullptr) return RENET_OK;        // hint cleared
  if (n_hot > kHotMax || R2 <= 0) { set_error("renet_set_hot_relations: n_hot must be in [0,%d], R2 > 0", kHotMax); return RENET_ERR_INVALID_ARG; }
  std::vector<char> seen(R2, 0);
  for (int i = 0; i < n_hot; ++i) {
    const int32_t r = hot_rel_host[i];
    if (r < 0 || r >= R2 || seen[r]) { set_error("renet_set_hot_relations: entries must be distinct ids in [0,%d)", R2); return RENET_ERR_INVALID_ARG; }
    seen[r] = 1;
  }
  g_hot_host.assign(hot_rel_host, hot_rel_host + n_hot);
  g_hot_R2 = R2;
  return RENET_OK;
}

int launch_rgcn_gather(const float* H, const int32_t* h_index, const float* W, const int32_t* row_ptr,
                       const int32_t* col_src, const int32_t* col_type, const float* norm, float* Hout,
                       int64_t N, int64_t E, int d_in, int d_out, int nb, int relu, int has_loop,
                       cudaStream_t stream, int R2) {
  if (N == 0) return RENET_OK;
  const int passthrough = (E == 0) ? 1 : 0;
  const bool fast = d_in == 200 && d_out == 200 && nb == 100 &&
                    ((reinterpret_cast<uintptr_t>(H) | reinterpret_cast<uintptr_t>(W) |
                      reinterpret_cast<uintptr_t>(Hout)) & 15) == 0;
  if (fast) {
    const unsigned block = kTileWarps * 32;
    const int variant = passthrough ? 0 : gather_variant();
    const unsigned n_tiles = (unsigned)((N + kTileNodes - 1) / kTileNodes);
    const unsigned grid = variant == 2 ? min(n_tiles, (unsigned)(kNumSMs * 3)) : n_tiles;
    const int32_t* order = nullptr;   // optional heaviest-first tile order: measured, no gain
    // experimental (variant 7 + renet_set_hot_relations): persistent kernel with the hot relation rows in shared
    // memory; bit-identical, measured slower than the tile kernel (61 vs 52 us), see DESIGN.md section 5
    const HotSet* hot = (!passthrough && variant == 7) ? hot_set_for_device(R2) : nullptr;
    if (hot != nullptr) {
      static bool attr_done[8] = {false, false, false, false, false, false, false, false};
      const int hkey = (relu ? 4 : 0) | (has_loop ? 2 : 0) | (h_index ? 1 : 0);
      const unsigned hgrid = min((n_tiles + kHotGroups - 1) / kHotGroups, (unsigned)kNumSMs);
      const unsigned hblock = kHotGroups * kTileWarps * 32;
#define RENET_LAUNCH_HOT(R, L, I)                                                                               \
  do {                                                                                                          \
    if (!attr_done[hkey]) {                                                                                     \
      RENET_CHECK_CUDA(cudaFuncSetAttribute(rgcn_gather_hot_kernel<R, L, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, kHotSmemBytes)); \
      attr_done[hkey] = true;                                                                                   \
    }                                                                                                           \
    rgcn_gather_hot_kernel<R, L, I><<<hgrid, hblock, kHotSmemBytes, stream>>>(H, h_index, W, row_ptr, col_src, col_type, norm, Hout, (int)N, hot->rel_slot, hot->hot_rel, hot->n_hot); \
  } while (0)
      switch (hkey) {
        case 0: RENET_LAUNCH_HOT(false, false, false); break;
        case 1: RENET_LAUNCH_HOT(false, false, true); break;
        case 2: RENET_LAUNCH_HOT(false, true, false); break;
        case 3: RENET_LAUNCH_HOT(false, true, true); break;
        case 4: RENET_LAUNCH_HOT(true, false, false); break;
        case 5: RENET_LAUNCH_HOT(true, false, true); break;
        case 6: RENET_LAUNCH_HOT(true, true, false); break;
        default: RENET_LAUNCH_HOT(true, true, true); break;
      }
#undef RENET_LAUNCH_HOT
      RENET_CHECK_LAUNCH("rgcn_gather_hot_kernel");
      return RENET_OK;
    }
#define RENET_LAUNCH_GATHER(R, L, I)                                                                            \
  if (variant == 6)                                                                                             \
    rgcn_gather_ring_kernel<R, L, I><<<n_tiles, block, 0, stream>>>(H, h_index, W, row_ptr, col_src, col_type, norm, Hout, (int)N); \
  else if (variant == 1)                                                                                        \
    rgcn_gather_d200_kernel<R, L, I, kTileNodes, 1><<<grid, block, 0, stream>>>(H, h_index, W, row_ptr, col_src, col_type, norm, Hout, (int)N, passthrough, nullptr); \
  else if (variant == 2)                                                                                        \
    rgcn_gather_d200_kernel<R, L, I, kTileNodes, 2><<<grid, block, 0, stream>>>(H, h_index, W, row_ptr, col_src, col_type, norm, Hout, (int)N, passthrough, nullptr); \
  else                                                                                                          \
    rgcn_gather_d200_kernel<R, L, I, kTileNodes, 0><<<grid, block, 0, stream>>>(H, h_index, W, row_ptr, col_src, col_type, norm, Hout, (int)N, passthrough, order)
    const int key = (relu ? 4 : 0) | (has_loop ? 2 : 0) | (h_index ? 1 : 0);
    switch (key) {
      case 0: RENET_LAUNCH_GATHER(false, false, false); break;
      case 1: RENET_LAUNCH_GATHER(false, false, true); break;
      case 2: RENET_LAUNCH_GATHER(false, true, false); break;
      case 3: RENET_LAUNCH_GATHER(false, true, true); break;
      case 4: RENET_LAUNCH_GATHER(true, false, false); break;
      case 5: RENET_LAUNCH_GATHER(true, false, true); break;
      case 6: RENET_LAUNCH_GATHER(true, true, false); break;
      default: RENET_LAUNCH_GATHER(true, true, true); break;
    }
#undef RENET_LAUNCH_GATHER
    RENET_CHECK_LAUNCH("rgcn_gather_d200_kernel");
  } else {
    const int64_t total = N * d_out;
    rgcn_gather_generic_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        H, h_index, W, row_ptr, col_src, col_type, norm, Hout, N, d_in, d_out, nb, relu, has_loop, passthrough);
    RENET_CHECK_LAUNCH("rgcn_gather_generic_kernel");
  }
  return RENET_OK;
}

int launch_rgcn_gather_comp(const float* H, const int32_t* h_index, const float* W, const int32_t* row_ptr,
                            const int32_t* col_src, const int32_t* col_type, const float* norm, float* Hout,
                            const int32_t* comp_ptr, const int32_t* comp_order, const int32_t* rel_slot,
                            const int32_t* hot_rel, int n_hot, int64_t G, int relu, int has_loop,
                            cudaStream_t stream) {
  if (G == 0) return RENET_OK;
  static bool attr_set = false;
#define RENET_FOR_ALL_COMP(X) X(false, false, false) X(false, false, true) X(false, true, false) X(false, true, true) \
    X(true, false, false) X(true, false, true) X(true, true, false) X(true, true, true)
  if (!attr_set) {
#define RENET_SET_ATTR(R, L, I)                                                                                 \
    RENET_CHECK_CUDA(cudaFuncSetAttribute(rgcn_gather_comp_kernel<R, L, I>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                          kCompSmemBytes));
    RENET_FOR_ALL_COMP(RENET_SET_ATTR)
#undef RENET_SET_ATTR
    attr_set = true;
  }
  const int key = (relu ? 4 : 0) | (has_loop ? 2 : 0) | (h_index ? 1 : 0);
#define RENET_LAUNCH_COMP(R, L, I)                                                                              \
  if (key == ((R ? 4 : 0) | (L ? 2 : 0) | (I ? 1 : 0)))                                                          \
    rgcn_gather_comp_kernel<R, L, I><<<(unsigned)G, kCompThreads, kCompSmemBytes, stream>>>(                     \
        H, h_index, W, row_ptr, col_src, col_type, norm, Hout, comp_ptr, comp_order, rel_slot, hot_rel, n_hot);
  RENET_FOR_ALL_COMP(RENET_LAUNCH_COMP)
#undef RENET_LAUNCH_COMP
#undef RENET_FOR_ALL_COMP
  RENET_CHECK_LAUNCH("rgcn_gather_comp_kernel");
  return RENET_OK;
}

}  // namespace renet
