/* CPU restatement (plain C + OpenMP) of the MindRec embedding-and-interaction hot path.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY — never linked into or called from mindrec_b200/.  It is (i)
 * cross-checked against oracle/ref_numpy.py in tests/test_oracle.py and (ii) timed by bench.py's
 * cpu_baseline and `--impl reference` legs as the "port" of the reference's CPU path (MindSpore itself
 * cannot be installed here: PARITY UNPINNED, see oracle/ref_numpy.py header and DESIGN.md).
 *
 * Each function cites the reference lines it follows (paths relative to /root/reference).
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC oracle/ref_c.c -o oracle/_build/libmrec_ref.so -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int mrec_ref_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Pin the OpenMP team size (bench.py: torchrun exports OMP_NUM_THREADS=1, which would otherwise shrink the CPU arm
 * to one core); returns the team size now in force. */
int mrec_ref_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

/* models/wide_deep/src/wide_and_deep.py:302,308-309: deep_in = table[ids] * mask -> [B, F*D] */
void mrec_ref_gather_masked(const float *table, int64_t vocab, int dim, const int32_t *ids,
                            const float *mask, int64_t n, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const int64_t id = ids[i];
    float *o = out + i * dim;
    if (id < 0 || id >= vocab) {
      memset(o, 0, sizeof(float) * dim);
      continue;
    }
    const float *r = table + id * dim;
    const float m = mask ? mask[i] : 1.0f;
    for (int d = 0; d < dim; ++d) o[d] = r[d] * m;
  }
}

/* wide_and_deep.py:300,305-306: wide_out = sum_f Ww[ids] * mask + Wide_b */
void mrec_ref_gather_reduce(const float *table, int64_t vocab, const int32_t *ids, const float *mask,
                            int64_t batch, int fields, float bias, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    double acc = 0.0;
    for (int f = 0; f < fields; ++f) {
      const int64_t id = ids[b * fields + f];
      if (id >= 0 && id < vocab) acc += (double)table[id] * (double)mask[b * fields + f];
    }
    out[b] = (float)(acc + (double)bias);
  }
}

/* P.Unique, ascending (GPU kernel order), via a stable LSD radix sort of (key, position):
 * mindspore_rec/ops/embedding.py:191-192 and the optimizer-side RowTensor dedup.
 * Keys outside [0, bound) collapse onto `bound`.  Returns U.  perm/seg_start may not be NULL. */
int64_t mrec_ref_unique(const int32_t *ids, int64_t n, int64_t bound, int32_t *uniq, int32_t *inverse,
                        int32_t *perm, int32_t *seg_start) {
  if (n == 0) {
    seg_start[0] = 0;
    return 0;
  }
  uint32_t *ka = (uint32_t *)malloc(sizeof(uint32_t) * n), *kb = (uint32_t *)malloc(sizeof(uint32_t) * n);
  int32_t *va = (int32_t *)malloc(sizeof(int32_t) * n), *vb = (int32_t *)malloc(sizeof(int32_t) * n);
  for (int64_t i = 0; i < n; ++i) {
    const int64_t k = ids[i];
    ka[i] = (k >= 0 && k < bound) ? (uint32_t)k : (uint32_t)bound;
    va[i] = (int32_t)i;
  }
  int bits = 1;
  while (bits < 32 && ((uint64_t)bound >> bits)) ++bits;
  for (int shift = 0; shift < bits; shift += 8) {
    int64_t cnt[257] = {0};
    for (int64_t i = 0; i < n; ++i) cnt[((ka[i] >> shift) & 255) + 1]++;
    for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
    for (int64_t i = 0; i < n; ++i) {
      const int64_t p = cnt[(ka[i] >> shift) & 255]++;
      kb[p] = ka[i];
      vb[p] = va[i];
    }
    uint32_t *tk = ka; ka = kb; kb = tk;
    int32_t *tv = va; va = vb; vb = tv;
  }
  int64_t u = -1;
  for (int64_t i = 0; i < n; ++i) {
    if (i == 0 || ka[i] != ka[i - 1]) {
      ++u;
      uniq[u] = (int32_t)ka[i];
      seg_start[u] = (int32_t)i;
    }
    perm[i] = va[i];
    inverse[va[i]] = (int32_t)u;
  }
  ++u;
  seg_start[u] = (int32_t)n;
  free(ka); free(kb); free(va); free(vb);
  return u;
}

/* UnsortedSegmentSum over sorted segments: gsum[u] = sum_n mask[p] * g[p / div], p = perm[n] */
void mrec_ref_segment_sum(const float *g, int dim, int div, const float *mask, const int32_t *perm,
                          const int32_t *seg_start, int64_t n_seg, float *gsum) {
#pragma omp parallel for schedule(dynamic, 256)
  for (int64_t u = 0; u < n_seg; ++u) {
    double acc[512];
    for (int d = 0; d < dim; ++d) acc[d] = 0.0;
    for (int32_t i = seg_start[u]; i < seg_start[u + 1]; ++i) {
      const int32_t p = perm[i];
      const double m = mask ? (double)mask[p] : 1.0;
      const float *r = g + (int64_t)(p / div) * dim;
      for (int d = 0; d < dim; ++d) acc[d] += m * (double)r[d];
    }
    for (int d = 0; d < dim; ++d) gsum[u * dim + d] = (float)acc[d];
  }
}

/* nn.LazyAdam on deduplicated rows (wide_and_deep.py:420-422). hyper: lr_t, b1, b2, eps, scale */
void mrec_ref_lazy_adam(float *w, float *m, float *v, int64_t vocab, int dim, const int32_t *uniq,
                        int64_t n_seg, const float *gsum, float lr_t, float beta1, float beta2, float eps,
                        float grad_scale) {
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < n_seg; ++u) {
    const int64_t row = uniq[u];
    if (row < 0 || row >= vocab) continue;
    float *W = w + row * dim, *M = m + row * dim, *V = v + row * dim;
    const float *G = gsum + u * dim;
    for (int d = 0; d < dim; ++d) {
      const float g = G[d] * grad_scale;
      M[d] = beta1 * M[d] + omb1 * g;
      V[d] = beta2 * V[d] + omb2 * g * g;
      W[d] = W[d] - lr_t * M[d] / (sqrtf(V[d]) + eps);
    }
  }
}

/* nn.Adam dense kernel */
void mrec_ref_adam_dense(float *w, float *m, float *v, const float *g, int64_t n, float lr_t, float beta1,
                         float beta2, float eps, float grad_scale) {
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const float gg = g[i] * grad_scale;
    m[i] = beta1 * m[i] + omb1 * gg;
    v[i] = beta2 * v[i] + omb2 * gg * gg;
    w[i] = w[i] - lr_t * m[i] / (sqrtf(v[i]) + eps);
  }
}

/* nn.FTRL on deduplicated rows (wide_and_deep.py:423-430), lr_power = -0.5 */
void mrec_ref_ftrl(float *w, float *acc, float *lin, int64_t vocab, int dim, const int32_t *uniq,
                   int64_t n_seg, const float *gsum, float lr, float l1, float l2, float grad_scale) {
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < n_seg; ++u) {
    const int64_t row = uniq[u];
    if (row < 0 || row >= vocab) continue;
    for (int d = 0; d < dim; ++d) {
      const int64_t o = row * dim + d;
      const float g = gsum[u * dim + d] * grad_scale;
      const float a_new = acc[o] + g * g;
      const float sigma = (sqrtf(a_new) - sqrtf(acc[o])) / lr;
      const float L = lin[o] + g - sigma * w[o];
      const float q = sqrtf(a_new) / lr + 2.0f * l2;
      const float sgn = L > 0.f ? 1.f : (L < 0.f ? -1.f : 0.f);
      w[o] = (fabsf(L) > l1) ? (sgn * l1 - L) / q : 0.f;
      lin[o] = L;
      acc[o] = a_new;
    }
  }
}

/* models/deepfm/src/deepfm.py:222-228: fm = 0.5 * sum_d[(sum_f vx)^2 - sum_f vx^2] */
void mrec_ref_fm_fwd(const float *vx, int64_t batch, int fields, int dim, float *out) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    double total = 0.0;
    for (int d = 0; d < dim; ++d) {
      double s = 0.0, s2 = 0.0;
      for (int f = 0; f < fields; ++f) {
        const double x = vx[(b * fields + f) * dim + d];
        s += x;
        s2 += x * x;
      }
      total += s * s - s2;
    }
    out[b] = (float)(0.5 * total);
  }
}

/* d fm / d vx = g * (S - vx) */
void mrec_ref_fm_bwd(const float *vx, const float *gout, int64_t batch, int fields, int dim, float *dvx) {
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < batch; ++b) {
    for (int d = 0; d < dim; ++d) {
      double s = 0.0;
      for (int f = 0; f < fields; ++f) s += vx[(b * fields + f) * dim + d];
      for (int f = 0; f < fields; ++f) {
        const int64_t o = (b * fields + f) * dim + d;
        dvx[o] = (float)((double)gout[b] * (s - (double)vx[o]));
      }
    }
  }
}

/* models/deep_and_cross/src/deep_and_cross.py:139-149 applied L times (:301-306), layer by layer.
 * w, b: [L, D']; s_out (optional): [B, L] saved dot products. */
void mrec_ref_cross_fwd(const float *x0, const float *w, const float *b, int64_t batch, int dp, int layers,
                        float *y, float *s_out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < batch; ++r) {
    const float *x = x0 + r * dp;
    float *xl = y + r * dp;
    memcpy(xl, x, sizeof(float) * dp);
    for (int l = 0; l < layers; ++l) {
      double s = 0.0;
      for (int d = 0; d < dp; ++d) s += (double)xl[d] * (double)w[l * dp + d];
      if (s_out) s_out[r * layers + l] = (float)s;
      for (int d = 0; d < dp; ++d) xl[d] = (float)((double)x[d] * s + (double)b[l * dp + d] + (double)xl[d]);
    }
  }
}
