// a8 — SigmoidCrossEntropyWithLogits + ReduceMean and its gradient seed, fused.
//
// Replaces, in models/wide_deep/src/wide_and_deep.py:315,354-355 and :479-486 (also deepfm.py:254-255,
// deep_and_cross.py:323-325), the chain  out = wide_out + deep_out;  log_loss = max(x,0) - x*z +
// log1p(exp(-|x|));  loss = ReduceMean(log_loss)  and the bprop seed  delta = sens * (sigmoid(x) - z) / B,
// plus sum(delta) (the gradient of the scalar bias Wide_b).  B is 16 k: ONE thread-block cluster of 8 CTAs
// x 1024 threads walks the batch (a single CTA was latency bound at 11-15 us, ncu r1g/r1i); the per-CTA sums are
// exchanged through distributed shared memory and added by CTA 0 in rank order, so the reduction is
// deterministic and needs no global workspace.  Writes loss, delta (fp32 and, for the fp16 DenseLayers, fp16)
// and sum(delta).
#include "common.cuh"
#include <cooperative_groups.h>
#include <cuda_fp16.h>

namespace mrec {

constexpr int kXentCtas = 8;       // portable cluster size

__global__ void __cluster_dims__(kXentCtas, 1, 1) __launch_bounds__(1024)
sigmoid_xent_kernel(const float* __restrict__ a, const float* __restrict__ b2, const float* __restrict__ label,
                    const float* __restrict__ scale /* sens */, int64_t batch, float* __restrict__ logit,
                    float* __restrict__ loss, float* __restrict__ delta, __half* __restrict__ delta16,
                    float* __restrict__ delta_sum) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ float s_l[32], s_d[32];
  __shared__ float s_cta[2];
  const float k = scale[0] / (float)batch;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  float acc_l = 0.f, acc_d = 0.f;
  // 8 elements per thread and round, loads first: the walk is latency bound (one CTA, fixed summation order)
  constexpr int U = 2;
  for (int64_t base = tid; base < batch; base += (int64_t)U * nthr) {
    float xa[U], xb[U], zl[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = min(base + (int64_t)u * nthr, batch - 1);
      xa[u] = a[i];
      xb[u] = b2 ? b2[i] : 0.f;
      zl[u] = label[i];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = base + (int64_t)u * nthr;
      if (i >= batch) break;
      const float x = xa[u] + xb[u];
      const float z = zl[u];
      const float e = __expf(-fabsf(x));
      acc_l += fmaxf(x, 0.f) - x * z + log1pf(e);
      // sigmoid(x) without overflow: x >= 0 -> 1/(1+e), x < 0 -> e/(1+e)
      const float sg = (x >= 0.f ? 1.f : e) / (1.f + e);
      const float d = (sg - z) * k;
      if (logit) logit[i] = x;
      delta[i] = d;
      if (delta16) delta16[i] = __float2half_rn(d);
      acc_d += d;
    }
  }
  acc_l = warp_sum(acc_l);
  acc_d = warp_sum(acc_d);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_l[warp] = acc_l; s_d[warp] = acc_d; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float tl = 0.f, td = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tl += s_l[w]; td += s_d[w]; }
    s_cta[0] = tl;
    s_cta[1] = td;
  }
  cluster.sync();                                  // every CTA's pair is visible cluster-wide
  if (cluster.block_rank() == 0 && threadIdx.x == 0) {
    float tl = 0.f, td = 0.f;
    for (unsigned r = 0; r < cluster.num_blocks(); ++r) {
      const float* peer = cluster.map_shared_rank(s_cta, r);
      tl += peer[0];
      td += peer[1];
    }
    loss[0] = tl / (float)batch;
    if (delta_sum) delta_sum[0] = td;
  }
  cluster.sync();                                  // keep the peers' shared memory alive until it has been read
}

}  // namespace mrec

using namespace mrec;

// in : a[B] f32 (e.g. wide_out), b[B] f32 | numel 0 (e.g. deep_out; logit = a + b), label[B] f32, sens[1] f32
// out: logit[B] f32, loss[1] f32 (mean), delta[B] f32 (= sens*(sigmoid(logit)-label)/B), delta16[B] f16 | numel 0,
//      delta_sum[1] f32
MREC_API int mrec_sigmoid_xent(int nparam, void** params, int* ndims, int64_t** shapes, const char** dtypes,
                               void* stream, void* /*extra*/) {
  Aot a{nparam, params, ndims, shapes, dtypes, (cudaStream_t)stream};
  if (a.nparam != 9) return fail(ERR_NPARAM, "mrec_sigmoid_xent: expected 9 params, got %d", a.nparam);
  const int64_t batch = a.numel(0);
  MREC_REQUIRE(a.is_f32(0) && a.is_f32(1) && a.is_f32(2) && a.is_f32(3) && a.is_f32(4) && a.is_f32(5) &&
                   a.is_f32(6) && a.is_f32(8), ERR_DTYPE, "mrec_sigmoid_xent: float32 expected");
  MREC_REQUIRE(a.numel(1) == 0 || a.numel(1) == batch, ERR_SHAPE, "mrec_sigmoid_xent: b must have 0 or B elements");
  MREC_REQUIRE(a.numel(2) == batch && a.numel(4) == batch && a.numel(6) == batch && a.numel(3) >= 1 &&
                   a.numel(5) >= 1 && a.numel(8) >= 1, ERR_SHAPE, "mrec_sigmoid_xent: shape mismatch");
  MREC_REQUIRE(a.numel(7) == 0 || (a.numel(7) == batch && a.is(7, "float16")), ERR_SHAPE,
               "mrec_sigmoid_xent: delta16 must be float16[B] or empty");
  for (int i : {0, 2, 3, 4, 5, 6, 8})
    if (!a.params[i] && a.numel(i) > 0) return fail(ERR_NULL, "mrec_sigmoid_xent: param %d is null", i);
  if (batch == 0) return OK;
  MREC_LAUNCH(sigmoid_xent_kernel, kXentCtas, 1024, 0, a.stream, a.ptr<float>(0), a.numel(1) ? a.ptr<float>(1) : nullptr,
              a.ptr<float>(2), a.ptr<float>(3), batch, a.ptr<float>(4), a.ptr<float>(5), a.ptr<float>(6),
              a.numel(7) ? a.ptr<__half>(7) : nullptr, a.ptr<float>(8));
  return check_launch("sigmoid_xent");
}
