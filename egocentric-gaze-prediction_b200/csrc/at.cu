// AT (attention transition) glue around the LSTM (reference AT.py:25-39 crop_feature, AT.py:58-66 get_weighted,
// AT.py:236-241 mean over the crop; run_spatialstream.py:85-104,136 demo variants).
// Inputs are the NCHW fp32 conv5_3 maps exactly as the reference's forward hook sees them.
#include "common.cuh"

namespace {

// out[b][c] = mean over the size x size window of feat[b][c] centred at clip(gaze[b]//16, size//2, H-ceil(size/2))
__global__ void crop_mean_kernel(const float* __restrict__ feat, const int* __restrict__ gaze, int B, int C, int H, int W,
                                 int size, int down, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int lo = size / 2, hi_off = (size + 1) / 2;
  int fr = gaze[2 * b] / down, fc = gaze[2 * b + 1] / down;   // non-negative gaze: // == /
  fr = min(max(fr, lo), H - hi_off);
  fc = min(max(fc, lo), H - hi_off);                          // reference clips BOTH coords with H (AT.py:32)
  const float* f = feat + ((size_t)b * C + c) * H * W;
  float s = 0.f;
  for (int i = fr - lo; i < fr + hi_off; ++i)
    for (int j = fc - lo; j < fc + hi_off; ++j) s += f[i * W + j];
  out[(size_t)b * C + c] = s / (float)(size * size);
}

// map[b][hw] = sum_c w[b][c] * feat[b][c][hw];  then (map - min) / max(map - min) per sample.
// One block per sample; threads stride over hw (coalesced along the NCHW inner dim), 4 channel groups.
__global__ void __launch_bounds__(1024)
weighted_map_kernel(const float* __restrict__ feat, const float* __restrict__ w, int C, int HW,
                    float* __restrict__ out) {
  extern __shared__ float sm[];  // [4][HW] partial maps + 64 scratch
  const int b = blockIdx.x;
  const int groups = blockDim.x / 256;      // 4
  const int g = threadIdx.x / 256, t = threadIdx.x % 256;
  const float* f = feat + (size_t)b * C * HW;
  const float* wb = w + (size_t)b * C;
  const int cpg = C / groups;
  for (int p = t; p < HW; p += 256) {
    float acc = 0.f;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) acc = fmaf(__ldg(wb + c), f[(size_t)c * HW + p], acc);
    sm[g * HW + p] = acc;
  }
  __syncthreads();
  float* red = sm + groups * HW;
  float vmin = INFINITY, vmax = -INFINITY;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    float v = 0.f;
    for (int gg = 0; gg < groups; ++gg) v += sm[gg * HW + p];
    sm[p] = v;  // group 0 slot now holds the full sum (each p touched by exactly one thread)
    vmin = fminf(vmin, v);
    vmax = fmaxf(vmax, v);
  }
  vmin = warp_min(vmin);
  vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = vmin; red[32 + (threadIdx.x >> 5)] = vmax; }
  __syncthreads();
  vmin = INFINITY; vmax = -INFINITY;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) { vmin = fminf(vmin, red[i]); vmax = fmaxf(vmax, red[32 + i]); }
  const float denom = vmax - vmin;
  for (int p = threadIdx.x; p < HW; p += blockDim.x) out[(size_t)b * HW + p] = (sm[p] - vmin) / denom;
}

// AT.crop_align_feature (AT.py:41-56) + mean (AT.py:239-241): bilinear x`up` upsample (align_corners=True), crop a
// (size*up)^2 window centred at clip(gaze, size*up/2, H*up - size*up/2), average.  The bilinear weights are separable, so
// the mean over the window is sum_ij F[i][j]*wy[i]*wx[j]: the upsampled (B,512,224,224) map is never materialised.
__global__ void crop_align_mean_kernel(const float* __restrict__ feat, const int* __restrict__ gaze, int C, int H, int W,
                                       int size, int up, float* __restrict__ out) {
  __shared__ float wy[64], wx[64];
  const int b = blockIdx.y;
  const int HS = H * up, WS = W * up, win = size * up;
  if (threadIdx.x < 64) { wy[threadIdx.x] = 0.f; wx[threadIdx.x] = 0.f; }
  __syncthreads();
  if (threadIdx.x < 2) {
    const int n_in = threadIdx.x == 0 ? H : W, n_out = threadIdx.x == 0 ? HS : WS;
    float* wv = threadIdx.x == 0 ? wy : wx;
    int f = gaze[2 * b + threadIdx.x];
    f = min(max(f, win / 2), HS - win / 2);   // reference clips both coordinates with H (=224)
    const float scale = n_out > 1 ? (float)(n_in - 1) / (float)(n_out - 1) : 0.f;
    for (int o = f - win / 2; o < f + win / 2; ++o) {
      const float src = (float)o * scale;
      const int i0 = min((int)src, n_in - 1), i1 = min(i0 + 1, n_in - 1);
      const float l = src - (float)i0;
      wv[i0] += 1.f - l;
      wv[i1] += l;
    }
  }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* f = feat + ((size_t)b * C + c) * H * W;
  float s = 0.f;
  for (int i = 0; i < H; ++i) {
    float r = 0.f;
    for (int j = 0; j < W; ++j) r = fmaf(f[i * W + j], wx[j], r);
    s = fmaf(r, wy[i], s);
  }
  out[(size_t)b * C + c] = s / (float)(win * win);
}

// F.upsample(scale_factor=S, mode='bilinear') == align_corners=False (run_spatialstream.py:136). x: [B][h][w] -> [B][hS][wS]
__global__ void bilinear_up_kernel(const float* __restrict__ x, int B, int h, int w, int S, int align_corners,
                                   float* __restrict__ out) {
  const int H = h * S, W = w * S;
  const size_t total = (size_t)B * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % W);
    const int oy = (int)((i / W) % H);
    const int b = (int)(i / ((size_t)W * H));
    float sy, sx;
    if (align_corners) {
      sy = H > 1 ? (float)oy * (float)(h - 1) / (float)(H - 1) : 0.f;
      sx = W > 1 ? (float)ox * (float)(w - 1) / (float)(W - 1) : 0.f;
    } else {
      const float rs = 1.f / (float)S;
      sy = fmaxf(((float)oy + 0.5f) * rs - 0.5f, 0.f);
      sx = fmaxf(((float)ox + 0.5f) * rs - 0.5f, 0.f);
    }
    const int y0 = min((int)sy, h - 1), x0 = min((int)sx, w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* xb = x + (size_t)b * h * w;
    const float v = (1.f - ly) * ((1.f - lx) * xb[y0 * w + x0] + lx * xb[y0 * w + x1]) +
                    ly * ((1.f - lx) * xb[y1 * w + x0] + lx * xb[y1 * w + x1]);
    out[i] = v;
  }
}

}  // namespace

// gaze: [B][2] int32 device (row, col) pixel coordinates in the 224-space; down = 16 (AT.py:31).
extern "C" int egaze_crop_mean(const float* feat_nchw, const int* gaze, int B, int C, int H, int W, int size, int down,
                               float* out, void* stream) {
  EGAZE_CHECK_ARG(feat_nchw && gaze && out && B > 0 && C > 0, "crop_mean: bad args");
  EGAZE_CHECK_ARG(size >= 1 && size <= H && size <= W && down >= 1, "crop_mean: bad crop size");
  dim3 grid(ceil_div(C, 128), B);
  crop_mean_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(feat_nchw, gaze, B, C, H, W, size, down, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_crop_align_mean(const float* feat_nchw, const int* gaze, int B, int C, int H, int W, int size, int up,
                                     float* out, void* stream) {
  EGAZE_CHECK_ARG(feat_nchw && gaze && out && B > 0 && C > 0, "crop_align_mean: bad args");
  EGAZE_CHECK_ARG(H <= 64 && W <= 64 && size >= 1 && up >= 1 && size * up <= H * up, "crop_align_mean: unsupported shape");
  dim3 grid(ceil_div(C, 128), B);
  crop_align_mean_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(feat_nchw, gaze, C, H, W, size, up, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_weighted_map(const float* feat_nchw, const float* chn_weight, int B, int C, int HW, float* out,
                                  void* stream) {
  EGAZE_CHECK_ARG(feat_nchw && chn_weight && out && B > 0, "weighted_map: bad args");
  EGAZE_CHECK_ARG(C % 4 == 0 && HW <= 8192, "weighted_map: unsupported C=%d HW=%d", C, HW);
  const size_t smem = ((size_t)4 * HW + 64) * sizeof(float);
  static unsigned long long attr = 0;
  if (egaze_first_on_device(&attr)) {
    EGAZE_CUDA(cudaFuncSetAttribute(weighted_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  }
  weighted_map_kernel<<<B, 1024, smem, (cudaStream_t)stream>>>(feat_nchw, chn_weight, C, HW, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_bilinear_up(const float* x, int B, int h, int w, int scale, int align_corners, float* out,
                                 void* stream) {
  EGAZE_CHECK_ARG(x && out && B > 0 && h > 0 && w > 0 && scale >= 1, "bilinear_up: bad args");
  const size_t total = (size_t)B * h * w * scale * scale;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  bilinear_up_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, B, h, w, scale, align_corners, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
