// floss: distance-weighted binary cross entropy (reference floss.py:5-41).
//   per sample: (cx, cy) = mean row / mean col of ALL pixels equal to the sample's max   (floss.py:23-25)
//               w[i][j]  = float32( 1 / ((sqrt((i-cx)^2 + (j-cy)^2) + 1) / W) )  computed in float64 (floss.py:27-39)
//   loss = mean_{b,i,j} -w * ( t*max(log p,-100) + (1-t)*max(log(1-p),-100) )               (floss.py:12)
// The reference builds w in NumPy on the CPU every step (a D2H + H2D pipeline stall); here it is computed on the fly.
// Reductions: warp shuffles -> shared memory -> one fp64 atomic per block.
#include "common.cuh"

namespace {

// One block per sample.  centroid[b] = (cx, cy) as doubles.
__global__ void __launch_bounds__(1024)
floss_centroid_kernel(const float* __restrict__ target, int HW, int W, double* __restrict__ centroid) {
  __shared__ float s_max[32];
  __shared__ unsigned long long s_cnt[32], s_r[32], s_c[32];
  const float* t = target + (size_t)blockIdx.x * HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < HW; i += blockDim.x) m = fmaxf(m, t[i]);
  m = warp_max(m);
  if (lane == 0) s_max[warp] = m;
  __syncthreads();
  m = -INFINITY;
  for (int i = 0; i < nw; ++i) m = fmaxf(m, s_max[i]);
  unsigned long long cnt = 0, sr = 0, sc = 0;
  for (int i = threadIdx.x; i < HW; i += blockDim.x)
    if (t[i] == m) {
      cnt += 1;
      sr += (unsigned)(i / W);
      sc += (unsigned)(i % W);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    sc += __shfl_xor_sync(0xffffffffu, sc, o);
  }
  if (lane == 0) { s_cnt[warp] = cnt; s_r[warp] = sr; s_c[warp] = sc; }
  __syncthreads();
  if (threadIdx.x == 0) {
    cnt = 0; sr = 0; sc = 0;
    for (int i = 0; i < nw; ++i) { cnt += s_cnt[i]; sr += s_r[i]; sc += s_c[i]; }
    centroid[2 * blockIdx.x + 0] = (double)sr / (double)cnt;
    centroid[2 * blockIdx.x + 1] = (double)sc / (double)cnt;
  }
}

__device__ __forceinline__ float floss_weight(int i, int j, double cx, double cy, int W) {
  const double a = (double)i - cx, b = (double)j - cy;
  const double dist = (sqrt(a * a + b * b) + 1.0) / (double)W;
  return (float)(1.0 / dist);
}

// mode 0: write weights; mode 1: accumulate loss; mode 2: write grad wrt input.
template <int MODE>
__global__ void __launch_bounds__(256)
floss_main_kernel(const float* __restrict__ p, const float* __restrict__ t, const double* __restrict__ centroid, int B,
                  int HW, int W, const float* __restrict__ gscale_dev, float* __restrict__ out, double* __restrict__ loss_acc) {
  const size_t total = (size_t)B * HW;
  const float gscale = (MODE == 2) ? gscale_dev[0] / (float)total : 0.f;
  double acc = 0.0;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / HW);
    const int r = (int)(idx - (size_t)b * HW);
    const float w = floss_weight(r / W, r % W, centroid[2 * b], centroid[2 * b + 1], W);
    if (MODE == 0) {
      out[idx] = w;
    } else {
      const float pp = p[idx], tt = t[idx];
      if (MODE == 1) {
        const float lp = fmaxf(logf(pp), -100.f);
        const float lq = fmaxf(log1pf(-pp), -100.f);
        acc += (double)(w * ((tt - 1.f) * lq - tt * lp));
      } else {
        out[idx] = gscale * w * (pp - tt) / fmaxf((1.f - pp) * pp, 1e-12f);
      }
    }
  }
  if (MODE == 1) {
    __shared__ double s[8];
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double v = 0.0;
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) v += s[i];
      atomicAdd(loss_acc, v);
    }
  }
}

__global__ void floss_finish_kernel(const double* acc, double n, float* loss) { loss[0] = (float)(acc[0] / n); }

}  // namespace

// centroid: [B][2] fp64 workspace (device).
extern "C" int egaze_floss_centroid(const float* target, int B, int H, int W, double* centroid, void* stream) {
  EGAZE_CHECK_ARG(target && centroid && B > 0 && H > 0 && W > 0, "floss_centroid: bad args");
  floss_centroid_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(target, H * W, W, centroid);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// floss.build_weight_from_target (floss.py:15-41): weights [B][H][W] fp32.  NOTE the reference divides by the LAST
// dim (image_width) for both axes and assumes square maps; same here.
extern "C" int egaze_floss_weight(const double* centroid, int B, int H, int W, float* weights, void* stream) {
  EGAZE_CHECK_ARG(centroid && weights && B > 0, "floss_weight: bad args");
  const size_t total = (size_t)B * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  floss_main_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(nullptr, nullptr, centroid, B, H * W, W, nullptr, weights,
                                                                 nullptr);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// loss_acc: one fp64 device scalar workspace; loss: one fp32 device scalar.
extern "C" int egaze_floss_fwd(const float* input, const float* target, const double* centroid, int B, int H, int W,
                               double* loss_acc, float* loss, void* stream) {
  EGAZE_CHECK_ARG(input && target && centroid && loss_acc && loss && B > 0, "floss_fwd: bad args");
  const size_t total = (size_t)B * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  EGAZE_CUDA(cudaMemsetAsync(loss_acc, 0, sizeof(double), (cudaStream_t)stream));
  floss_main_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(input, target, centroid, B, H * W, W, nullptr, nullptr,
                                                                 loss_acc);
  EGAZE_LAUNCH_CHECK();
  floss_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(loss_acc, (double)total, loss);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// grad_loss: DEVICE pointer to the scalar upstream gradient (no host sync).
// grad_input = grad_loss * w * (p - t) / max(p(1-p), 1e-12) / numel   (PyTorch BCE backward)
extern "C" int egaze_floss_bwd(const float* input, const float* target, const double* centroid, int B, int H, int W,
                               const float* grad_loss, float* grad_input, void* stream) {
  EGAZE_CHECK_ARG(input && target && centroid && grad_input && grad_loss && B > 0, "floss_bwd: bad args");
  const size_t total = (size_t)B * H * W;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  floss_main_kernel<2><<<blocks, 256, 0, (cudaStream_t)stream>>>(input, target, centroid, B, H * W, W,
                                                                 grad_loss, grad_input, nullptr);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
