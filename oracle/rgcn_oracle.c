/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the RE-Net RGCN block layer (CPU, fp32).
 *
 * Follows reference RGCN.py:33-51 (RGCNLayer.forward: self-loop mm, add, activation) and
 * RGCN.py:79-94 (RGCNBlockLayer: per-edge block-diagonal transform, fn.sum over in-edges, * norm).
 * Pinned against the reference's own outputs through tests/golden/layer_cases.npz
 * (tests/test_oracle_c.py).  Never linked or loaded by the product package (renet_b200/); only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 *   out[v] = act( norm[v] * sum_{e: dst[e]=v} blockdiag(W[etype[e]]) . H[src[e]]  +  H[v] @ Wloop )
 *
 * Build: make -C oracle   (gcc -O3 -fopenmp -shared -fPIC)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* returns 0 on success, -1 on allocation failure */
int oracle_rgcn_block_layer(const float* H, const float* W, const float* Wloop, const int64_t* src,
                            const int64_t* dst, const int64_t* etype, const float* norm, float* out,
                            int64_t N, int64_t E, int d_in, int d_out, int nb, int relu) {
  const int si = d_in / nb, so = d_out / nb;
  /* in-edge lists per destination (counting sort, stable): what DGL's reduce does implicitly */
  int64_t* ptr = (int64_t*)calloc((size_t)N + 1, sizeof(int64_t));
  int64_t* eid = (int64_t*)malloc((size_t)(E > 0 ? E : 1) * sizeof(int64_t));
  int64_t* fill = (int64_t*)malloc((size_t)(N > 0 ? N : 1) * sizeof(int64_t));
  if (!ptr || !eid || !fill) { free(ptr); free(eid); free(fill); return -1; }
  for (int64_t e = 0; e < E; ++e) ptr[dst[e] + 1]++;
  for (int64_t v = 0; v < N; ++v) ptr[v + 1] += ptr[v];
  memcpy(fill, ptr, (size_t)N * sizeof(int64_t));
  for (int64_t e = 0; e < E; ++e) eid[fill[dst[e]]++] = e;

#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t v = 0; v < N; ++v) {
    float* o = out + v * d_out;
    if (E == 0) {                       /* DGL 0.4: a graph without edges skips the reduce */
      for (int c = 0; c < d_out; ++c) o[c] = H[v * d_in + c];
    } else {
      for (int c = 0; c < d_out; ++c) o[c] = 0.f;
      for (int64_t k = ptr[v]; k < ptr[v + 1]; ++k) {
        const int64_t e = eid[k];
        const float* h = H + src[e] * d_in;
        const float* w = W + etype[e] * (int64_t)nb * si * so;      /* RGCN.py:81-85 */
        for (int b = 0; b < nb; ++b)                                 /* RGCN.py:86-87 (bmm) */
          for (int i = 0; i < si; ++i) {
            const float x = h[b * si + i];
            for (int j = 0; j < so; ++j) o[b * so + j] += x * w[(b * si + i) * so + j];
          }
      }
    }
    const float nv = norm[v];
    for (int c = 0; c < d_out; ++c) o[c] *= nv;                      /* RGCN.py:93-94 */
    if (Wloop) {                                                     /* RGCN.py:35,45-46 */
      const float* h = H + v * d_in;
      for (int k = 0; k < d_in; ++k) {
        const float x = h[k];
        const float* wl = Wloop + (int64_t)k * d_out;
        for (int c = 0; c < d_out; ++c) o[c] += x * wl[c];
      }
    }
    if (relu)                                                        /* RGCN.py:47-48 */
      for (int c = 0; c < d_out; ++c) o[c] = o[c] > 0.f ? o[c] : 0.f;
  }
  free(ptr); free(eid); free(fill);
  return 0;
}
