// Fused multi-tensor Adam that also maintains the packed tensor-core copies of the conv weights.
//
// The reference steps torch.optim.Adam over all parameters (SP.py:110-113,137; LF.py:77,99).  On this path every 3x3 conv weight
// additionally exists as packed operand copies (forward: [tap][Cout][Cin_p] fp16/bf16 hi+lo; data gradient: flipped /
// transposed [tap][Cin][Cout_p] bf16 hi+lo) that round 1 rebuilt with a separate multi-weight pack launch after ~30 foreach
// optimiser launches.  Here ONE launch (plus a one-block counter bump) walks a device table of parameters: it applies the Adam
// update to the fp32 OIHW master (same arithmetic as torch.optim.Adam: L2 weight decay folded into the gradient, bias
// correction, eps added after the square root) and, for conv weights, writes both packed copies from the freshly updated values
// through a 32 x 32 x 9 shared-memory tile so that all global accesses stay coalesced.  HBM-bound: 16 B read + 12 B written per
// parameter, + 8 B per packed element.
#include "common.cuh"

namespace {

struct AdamJob {              // 112 bytes, mirrored by egaze/optim.py
  float* w;                   // fp32 master (OIHW for conv weights)
  const float* g;             // gradient, same layout
  float* m;                   // exp_avg
  float* v;                   // exp_avg_sq
  float* step;                // device scalar: number of steps taken (already incremented for this update)
  void* p0_hi;                // forward copy [9][rows0][cols0] (null: none)
  void* p0_lo;
  void* p1_hi;                // data-gradient copy [9][rows1][cols1], taps flipped (null: none)
  void* p1_lo;
  long long n;                // elements
  int Co, Ci;                 // conv weight [Co][Ci][3][3]; Ci == 0: flat tensor of n elements, no packed copies
  int rows0, cols0, fmt0;     // forward copy: rows0 >= Co, cols0 >= Ci, fmt 0 bf16 / 1 fp16 (values pre-scaled, see layout.cu)
  int rows1, cols1;           // data-gradient copy: rows1 >= Ci, cols1 >= Co (bf16)
  int sub;                    // the copies are the 16-plane SUB-PIXEL packs (layout.cu modes 2 / 3; no tap flip in the gradient copy)
};

struct AdamHyper {
  float lr, beta1, beta2, eps, weight_decay, f16_scale;
};

__device__ __forceinline__ float adam_update(float w, float g, float& m, float& v, const AdamHyper& h, float step_size,
                                             float inv_sqrt_bias2) {
  g = fmaf(h.weight_decay, w, g);
  m = fmaf(1.f - h.beta1, g - m, m);                       // exp_avg.lerp_(grad, 1 - beta1)
  v = fmaf(1.f - h.beta2, g * g, h.beta2 * v);            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) * inv_sqrt_bias2 + h.eps;
  return w - step_size * (m / denom);
}

__device__ __forceinline__ void store_split(void* hi, void* lo, size_t idx, float x, int fmt, float f16_scale) {
  if (fmt) {
    __half a, b;
    split_f16(x * f16_scale, a, b);
    reinterpret_cast<__half*>(hi)[idx] = a;
    if (lo) reinterpret_cast<__half*>(lo)[idx] = b;
  } else {
    __nv_bfloat16 a, b;
    split_bf16(x, a, b);
    reinterpret_cast<__nv_bfloat16*>(hi)[idx] = a;
    if (lo) reinterpret_cast<__nv_bfloat16*>(lo)[idx] = b;
  }
}

__global__ void adam_step_inc_kernel(const AdamJob* __restrict__ jobs, int njobs) {
  for (int j = threadIdx.x; j < njobs; j += blockDim.x) {
    // several parameters may share one counter (they do not here, but be safe): only the first job of a counter bumps it
    bool first = true;
    for (int k = 0; k < j; ++k)
      if (jobs[k].step == jobs[j].step) { first = false; break; }
    if (first) *jobs[j].step += 1.f;
  }
}

constexpr int kTile = 32;

__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamJob* __restrict__ jobs, const AdamHyper h) {
  __shared__ float tile[9][kTile][kTile + 1];
  const AdamJob j = jobs[blockIdx.y];
  const float step = *j.step;
  const float bias1 = 1.f - powf(h.beta1, step), bias2 = 1.f - powf(h.beta2, step);
  const float step_size = h.lr / bias1, inv_sqrt_bias2 = rsqrtf(bias2);
  if (j.Ci == 0) {
    if (blockIdx.x * blockDim.x >= j.n) return;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < j.n; i += (long long)gridDim.x * blockDim.x) {
      float m = j.m[i], v = j.v[i];
      j.w[i] = adam_update(j.w[i], j.g[i], m, v, h, step_size, inv_sqrt_bias2);
      j.m[i] = m;
      j.v[i] = v;
    }
    return;
  }
  const int tiles_ci = (j.Ci + kTile - 1) / kTile, tiles_co = (j.Co + kTile - 1) / kTile;
  const size_t plane0 = (size_t)j.rows0 * j.cols0, plane1 = (size_t)j.rows1 * j.cols1;
  for (int t = blockIdx.x; t < tiles_ci * tiles_co; t += gridDim.x) {
    const int co0 = (t / tiles_ci) * kTile, ci0 = (t % tiles_ci) * kTile;
    // phase 1: update.  For one output channel the tile's (ci, tap) elements are one contiguous run of 32*9 floats, so the
    // four arrays are streamed with fully coalesced float4 accesses (rows whose length or start is not a multiple of four
    // floats -- only the 3-channel RGB conv -- take the scalar path); the new weights are parked in the shared tile.
    const int tci = min(kTile, j.Ci - ci0), tco = min(kTile, j.Co - co0);
    const int row_len = tci * 9;
    const bool aligned = ((reinterpret_cast<uintptr_t>(j.w) | reinterpret_cast<uintptr_t>(j.g) | reinterpret_cast<uintptr_t>(j.m) |
                           reinterpret_cast<uintptr_t>(j.v)) & 15) == 0;
    if (aligned && ((j.Ci * 9) & 3) == 0 && (row_len & 3) == 0) {
      const int row4 = row_len >> 2;
      for (int f = threadIdx.x; f < tco * row4; f += blockDim.x) {
        const int co_l = f / row4, c4 = f - co_l * row4;
        const size_t base = ((size_t)(co0 + co_l) * j.Ci + ci0) * 9 + (size_t)c4 * 4;
        float4 w4 = *reinterpret_cast<const float4*>(j.w + base);
        const float4 g4 = *reinterpret_cast<const float4*>(j.g + base);
        float4 m4 = *reinterpret_cast<const float4*>(j.m + base);
        float4 v4 = *reinterpret_cast<const float4*>(j.v + base);
        w4.x = adam_update(w4.x, g4.x, m4.x, v4.x, h, step_size, inv_sqrt_bias2);
        w4.y = adam_update(w4.y, g4.y, m4.y, v4.y, h, step_size, inv_sqrt_bias2);
        w4.z = adam_update(w4.z, g4.z, m4.z, v4.z, h, step_size, inv_sqrt_bias2);
        w4.w = adam_update(w4.w, g4.w, m4.w, v4.w, h, step_size, inv_sqrt_bias2);
        *reinterpret_cast<float4*>(j.w + base) = w4;
        *reinterpret_cast<float4*>(j.m + base) = m4;
        *reinterpret_cast<float4*>(j.v + base) = v4;
        const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int q = c4 * 4 + e, ci_l = q / 9, tap = q - ci_l * 9;
          tile[tap][co_l][ci_l] = wv[e];
        }
      }
    } else {
      for (int p = threadIdx.x; p < kTile * kTile; p += blockDim.x) {
        const int co_l = p / kTile, ci_l = p % kTile;
        const int co = co0 + co_l, ci = ci0 + ci_l;
        if (co < j.Co && ci < j.Ci) {
          const size_t base = ((size_t)co * j.Ci + ci) * 9;
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {
            float m = j.m[base + tap], v = j.v[base + tap];
            const float w = adam_update(j.w[base + tap], j.g[base + tap], m, v, h, step_size, inv_sqrt_bias2);
            j.w[base + tap] = w;
            j.m[base + tap] = m;
            j.v[base + tap] = v;
            tile[tap][co_l][ci_l] = w;
          }
        }
      }
    }
    __syncthreads();
    // phase 2: packed copies from the updated tile (padding rows / columns of the copies are zero and never change)
    if (j.sub) {
      for (int half = 0; half < 2; ++half) {
        void* hi = half == 0 ? j.p0_hi : j.p1_hi;
        void* lo = half == 0 ? j.p0_lo : j.p1_lo;
        if (!hi) continue;
        for (int p = threadIdx.x; p < kTile * kTile; p += blockDim.x) {
          // forward copy: adjacent threads write adjacent input channels; gradient copy: adjacent output channels
          const int co_l = half == 0 ? p / kTile : p % kTile, ci_l = half == 0 ? p % kTile : p / kTile;
          const int co = co0 + co_l, ci = ci0 + ci_l;
          if (co < j.Co && ci < j.Ci) {
            float w9[9], q[16];
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) w9[tap] = tile[tap][co_l][ci_l];
            egaze_subpixel_taps(w9, q);
            const size_t plane = half == 0 ? plane0 : plane1;
            const size_t at = half == 0 ? (size_t)co * j.cols0 + ci : (size_t)ci * j.cols1 + co;
#pragma unroll
            for (int t = 0; t < 16; ++t)
              store_split(hi, lo, (size_t)t * plane + at, q[t], half == 0 ? j.fmt0 : 0, half == 0 ? h.f16_scale : 1.f);
          }
        }
      }
      __syncthreads();
      continue;
    }
    if (j.p0_hi) {
      for (int p = threadIdx.x; p < kTile * kTile; p += blockDim.x) {
        const int co_l = p / kTile, ci_l = p % kTile;
        const int co = co0 + co_l, ci = ci0 + ci_l;
        if (co < j.Co && ci < j.Ci) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            store_split(j.p0_hi, j.p0_lo, (size_t)tap * plane0 + (size_t)co * j.cols0 + ci, tile[tap][co_l][ci_l], j.fmt0,
                        h.f16_scale);
        }
      }
    }
    if (j.p1_hi) {
      for (int p = threadIdx.x; p < kTile * kTile; p += blockDim.x) {
        const int ci_l = p / kTile, co_l = p % kTile;   // transposed mapping: adjacent threads write adjacent output channels
        const int co = co0 + co_l, ci = ci0 + ci_l;
        if (co < j.Co && ci < j.Ci) {
#pragma unroll
          for (int tap = 0; tap < 9; ++tap)
            store_split(j.p1_hi, j.p1_lo, (size_t)(8 - tap) * plane1 + (size_t)ci * j.cols1 + co, tile[tap][co_l][ci_l], 0, 1.f);
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" int egaze_adam_job_bytes(int* out) {
  if (out) *out = (int)sizeof(AdamJob);
  return EGAZE_OK;
}

// See include/egaze.h.
extern "C" int egaze_adam_multi(const void* jobs, int njobs, float lr, float beta1, float beta2, float eps, float weight_decay,
                                float f16_scale, void* stream) {
  EGAZE_CHECK_ARG(jobs && njobs > 0, "adam_multi: bad args");
  EGAZE_CHECK_ARG(lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps >= 0.f, "adam_multi: bad hyper-parameters");
  cudaStream_t st = (cudaStream_t)stream;
  adam_step_inc_kernel<<<1, 256, 0, st>>>((const AdamJob*)jobs, njobs);
  EGAZE_LAUNCH_CHECK();
  AdamHyper h = {lr, beta1, beta2, eps, weight_decay, f16_scale};
  dim3 grid(48, (unsigned)njobs);
  adam_multi_kernel<<<grid, 256, 0, st>>>((const AdamJob*)jobs, h);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
