// BatchNorm2d (reference utils.py:72, models/model_SP.py:12, models/late_fusion.py:10-12) split into the
// pieces that fuse around the tcgen05 conv:
//   conv epilogue      -> per-tile (mean, M2) partials            (conv3x3_tc.cu, phase 2)
//   egaze_col_stats    -> the same partials for a plain [rows][C] fp32 matrix (fusion layer after pair-max)
//   egaze_bn_finalize  -> Chan-combine partials in fp64 -> batch mean / biased var, running-stat update
//                         (momentum, unbiased var, exactly nn.BatchNorm2d), folded scale/shift
//   egaze_bn_apply     -> y = relu(x*scale+shift) [-> 2x2 max-pool] -> NHWC split-bf16 (and/or fp32)
//   egaze_pairmax      -> max over the two streams of the fusion layer (MaxPool3d((2,1,1)), model_SP.py:11)
// All HBM-bound: float4 (4 channels) per thread, grid sized in multiples of the SM count.
#include "common.cuh"

namespace {

// partial: [T][2][C] (mean, M2), cnt: [T].  One thread column per channel, 32 slices of tiles per block.
// The kernel sits on the critical path of every train-mode layer (conv -> finalize -> apply) and is pure latency: a thread
// first loads ALL of its partials (its first kFinPer -- all of them for one partial per SM -- independent loads, one round trip to L2) and only then
// runs the two Chan passes on registers; the num_batches_tracked bump rides along.
constexpr int kFinPer = 5;   // partials a thread keeps in registers (32 slices x 5 = 160 >= one partial per SM)
__global__ void __launch_bounds__(1024)
bn_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ cnt, int cnt_stride, int cnt_div,
                   int T, int C, float eps,
                   float momentum, const float* __restrict__ gamma, const float* __restrict__ beta,
                   float* __restrict__ running_mean, float* __restrict__ running_var, float* __restrict__ mean_out,
                   float* __restrict__ invstd_out, float* __restrict__ scale_out, float* __restrict__ shift_out,
                   long long* __restrict__ num_batches_tracked) {
  // nn.BatchNorm2d.num_batches_tracked += 1 rides along (one thread of the launch) instead of costing an add_ kernel per layer
  if (num_batches_tracked && blockIdx.x == 0 && threadIdx.x == 0 && threadIdx.y == 0) *num_batches_tracked += 1;
  __shared__ double s_n[32][33], s_a[32][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int sl = threadIdx.y;
  const bool single = T <= 32 * kFinPer;   // every partial of this thread fits its registers: one load round
  float fn[kFinPer], fm[kFinPer], fq[kFinPer];
#pragma unroll
  for (int i = 0; i < kFinPer; ++i) {
    const int t = sl + 32 * i;
    const bool ok = c < C && t < T;
    fn[i] = ok ? cnt[(size_t)t * cnt_stride + c / cnt_div] : 0.f;
    fm[i] = ok ? partial[((size_t)t * 2 + 0) * C + c] : 0.f;
    fq[i] = ok ? partial[((size_t)t * 2 + 1) * C + c] : 0.f;
  }
  // pass 1: N = sum n_t, S = sum n_t * mean_t  (no divisions in the loop)
  double n = 0.0, sm = 0.0;
#pragma unroll
  for (int i = 0; i < kFinPer; ++i)
    if (fn[i] > 0.f) {   // unused partial slot: its (mean, M2) are uninitialised
      n += (double)fn[i];
      sm = fma((double)fn[i], (double)fm[i], sm);
    }
  if (!single && c < C)
    for (int t = sl + 32 * kFinPer; t < T; t += 32) {
      const double nb = (double)cnt[(size_t)t * cnt_stride + c / cnt_div];
      if (nb <= 0.0) continue;
      n += nb;
      sm = fma(nb, (double)partial[((size_t)t * 2 + 0) * C + c], sm);
    }
  s_n[sl][threadIdx.x] = n;
  s_a[sl][threadIdx.x] = sm;
  __syncthreads();
  n = 0.0; sm = 0.0;
  for (int i = 0; i < 32; ++i) { n += s_n[i][threadIdx.x]; sm += s_a[i][threadIdx.x]; }
  const double mean = n > 0.0 ? sm / n : 0.0;
  __syncthreads();
  // pass 2: M2 = sum ( M2_t + n_t * (mean_t - mean)^2 )
  double m2 = 0.0;
#pragma unroll
  for (int i = 0; i < kFinPer; ++i)
    if (fn[i] > 0.f) {
      const double d = (double)fm[i] - mean;
      m2 += (double)fq[i] + (double)fn[i] * d * d;
    }
  if (!single && c < C)
    for (int t = sl + 32 * kFinPer; t < T; t += 32) {
      const double nb = (double)cnt[(size_t)t * cnt_stride + c / cnt_div];
      if (nb <= 0.0) continue;
      const double d = (double)partial[((size_t)t * 2 + 0) * C + c] - mean;
      m2 += (double)partial[((size_t)t * 2 + 1) * C + c] + nb * d * d;
    }
  s_a[sl][threadIdx.x] = m2;
  __syncthreads();
  if (sl == 0 && c < C) {
    m2 = 0.0;
    for (int i = 0; i < 32; ++i) m2 += s_a[i][threadIdx.x];
    const double var_b = m2 / n;                       // biased: used to normalise
    const double var_u = n > 1.0 ? m2 / (n - 1.0) : var_b;  // unbiased: goes into running_var
    const float invstd = (float)(1.0 / sqrt(var_b + (double)eps));
    if (mean_out) mean_out[c] = (float)mean;
    if (invstd_out) invstd_out[c] = invstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)var_u;
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    const float sc = g * invstd;
    if (scale_out) scale_out[c] = sc;
    if (shift_out) shift_out[c] = b - (float)mean * sc;
  }
}

// eval-mode fold: scale = gamma / sqrt(running_var + eps), shift = beta - running_mean * scale (+ conv bias * scale)
__global__ void bn_fold_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                               const float* conv_bias, float eps, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float sc = (gamma ? gamma[c] : 1.f) / sqrtf(rv[c] + eps);
  float sh = (beta ? beta[c] : 0.f) - rm[c] * sc;
  if (conv_bias) sh += conv_bias[c] * sc;
  scale[c] = sc;
  shift[c] = sh;
}

// x: [rows][C] fp32 -> partial (mean, M2) per block of `rows_per_blk` rows.  blockDim = (32 ch-quads? no: 128 ch)
__global__ void col_stats_kernel(const float* __restrict__ x, int rows, int C, int rows_per_blk,
                                 float* __restrict__ partial, float* __restrict__ cnt) {
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.x * rows_per_blk;
  const int r1 = min(rows, r0 + rows_per_blk);
  if (c >= C) return;
  float sum = 0.f;
  for (int r = r0; r < r1; ++r) sum += x[(size_t)r * C + c];
  const float mean = sum / (float)(r1 - r0);
  float m2 = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float d = x[(size_t)r * C + c] - mean;
    m2 = fmaf(d, d, m2);
  }
  partial[((size_t)blockIdx.x * 2 + 0) * C + c] = mean;
  partial[((size_t)blockIdx.x * 2 + 1) * C + c] = m2;
  if (c == 0) cnt[blockIdx.x] = (float)(r1 - r0);
}

// y = relu?(x*scale+shift), optional 2x2 max pool; x NHWC fp32 [N,H,W,C]; outputs NHWC [N,Ho,Wo,C].
// A thread keeps ONE float4 channel group (its scale / shift live in registers) and strides over pixels; all index
// arithmetic is 32-bit (the entry point checks the element count), with no division on the un-pooled path.
__global__ void __launch_bounds__(256)
bn_apply_kernel(const float* __restrict__ x, int N, int H, int W, int C, const float* __restrict__ scale,
                const float* __restrict__ shift, int relu, int pool, float* __restrict__ out_f32,
                __nv_bfloat16* __restrict__ out_hi, __nv_bfloat16* __restrict__ out_lo, __nv_bfloat16* __restrict__ out_xb,
                int fmt) {
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const int C4 = C / 4;
  const int tpp = C4 < 256 ? C4 : 256;               // threads per pixel
  const int ppb = 256 / tpp;                         // pixels per block sweep
  const int lane_c = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  if (pl >= ppb) return;                             // C/4 does not divide 256: the spare threads idle
  const uint32_t npix = (uint32_t)N * Ho * Wo;
  const float lo_clamp = relu ? 0.f : -INFINITY;
  for (int c4 = lane_c; c4 < C4; c4 += tpp) {
    const float4 sc = *reinterpret_cast<const float4*>(scale + c4 * 4);
    const float4 sh = *reinterpret_cast<const float4*>(shift + c4 * 4);
    for (uint32_t pix = blockIdx.x * ppb + pl; pix < npix; pix += gridDim.x * ppb) {
      float4 v;
      if (!pool) {
        const float4 a = *reinterpret_cast<const float4*>(x + (size_t)pix * C + c4 * 4);
        v.x = fmaf(a.x, sc.x, sh.x); v.y = fmaf(a.y, sc.y, sh.y); v.z = fmaf(a.z, sc.z, sh.z); v.w = fmaf(a.w, sc.w, sh.w);
      } else {
        const uint32_t ow = pix % Wo, t = pix / Wo, oh = t % Ho, n = t / Ho;
        const float* src = x + ((size_t)(n * H + 2 * oh) * W + 2 * ow) * C + c4 * 4;
        v = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const float4 a = *reinterpret_cast<const float4*>(src + ((size_t)dy * W + dx) * C);
            v.x = fmaxf(v.x, fmaf(a.x, sc.x, sh.x));
            v.y = fmaxf(v.y, fmaf(a.y, sc.y, sh.y));
            v.z = fmaxf(v.z, fmaf(a.z, sc.z, sh.z));
            v.w = fmaxf(v.w, fmaf(a.w, sc.w, sh.w));
          }
      }
      v.x = fmaxf(v.x, lo_clamp); v.y = fmaxf(v.y, lo_clamp); v.z = fmaxf(v.z, lo_clamp); v.w = fmaxf(v.w, lo_clamp);
      const size_t o = (size_t)pix * C + c4 * 4;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + o) = v;
      if (out_hi) {
        uint2 hi2, lo2;
        if (fmt) split_f16x4(v, hi2, lo2);
        else split_bf16x4(v, hi2, lo2);
        *reinterpret_cast<uint2*>(out_hi + o) = hi2;
        if (out_lo) *reinterpret_cast<uint2*>(out_lo + o) = lo2;
      }
      if (out_xb) *reinterpret_cast<uint2*>(out_xb + o) = pack_bf16x4(v);
    }
  }
}

// out[b] = max(x[b], x[b+B]) elementwise over [B][M] fp32 (M = H*W*C); optional folded affine + relu + split
__global__ void pairmax_kernel(const float* __restrict__ x, size_t per_stream, float* __restrict__ out) {
  const size_t n4 = per_stream / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 a = reinterpret_cast<const float4*>(x)[i];
    const float4 b = reinterpret_cast<const float4*>(x + per_stream)[i];
    reinterpret_cast<float4*>(out)[i] = make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
// Recompute z = raw*scale+shift for the 1 (no pool) or 4 (pool) source pixels of one output pixel and route the
// incoming gradient g: ReLU mask (z > 0) and, for MaxPool2d, to the FIRST maximum in scan order (PyTorch tie rule,
// SURVEY App. D).  gz[q] is the gradient w.r.t. z at source pixel q, xh[q] the normalised activation.
struct BnSrc {
  float gz[4][4];   // [pixel q][channel lane]
  float xh[4][4];
};

// src: the (top-left) source pixel's channel group; row_pitch: floats between image rows of raw.
template <int POOL>
__device__ __forceinline__ void bn_route(const float* __restrict__ src, size_t row_pitch, int C, int relu,
                                         const float4& g, const float4& sc, const float4& sh, const float4& mean,
                                         const float4& invstd, BnSrc& o) {
  constexpr int pool = POOL;
  constexpr int np = POOL ? 4 : 1;
  float z[4][4];
  const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
  const float mv[4] = {mean.x, mean.y, mean.z, mean.w}, iv[4] = {invstd.x, invstd.y, invstd.z, invstd.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (q < np) {
      const float4 a = *reinterpret_cast<const float4*>(src + (size_t)(q >> 1) * row_pitch + (size_t)(q & 1) * C);
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        z[q][l] = fmaf(av[l], scv[l], shv[l]);
        o.xh[q][l] = (av[l] - mv[l]) * iv[l];
      }
    }
  }
  const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    int best = 0;
    if (pool) {
      float bz = z[0][l];
#pragma unroll
      for (int q = 1; q < 4; ++q)
        if (z[q][l] > bz) { bz = z[q][l]; best = q; }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < np) o.gz[q][l] = (q == best && (!relu || z[q][l] > 0.f)) ? gv[l] : 0.f;
  }
}

// acc[slot][0][c] += sum gz ; acc[slot][1][c] += sum gz*xhat (slot = block % kBwdSlots)     blockDim = (C/4, rows)
constexpr int kBwdSlots = 8;
// (POOL is a template parameter: the generic form kept 4 source pixels' worth of state live -- 79 registers, three blocks per SM --
//  also for the eight un-pooled layers of a trunk, which need a quarter of it: 46-48 registers, five blocks per SM)
template <int POOL>
__global__ void __launch_bounds__(256, POOL ? 3 : 5)
bn_bwd_reduce_kernel(const float* __restrict__ raw, const float* __restrict__ g, int N, int H, int W,
                                     int C, const float* __restrict__ scale, const float* __restrict__ shift,
                                     const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                                     double* __restrict__ acc) {
  constexpr int pool = POOL;
  extern __shared__ float red[];  // [rows][2][C]
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const int c4 = threadIdx.x;
  const float4 sc = *reinterpret_cast<const float4*>(scale + c4 * 4), sh = *reinterpret_cast<const float4*>(shift + c4 * 4);
  const float4 mn = *reinterpret_cast<const float4*>(mean + c4 * 4), iv = *reinterpret_cast<const float4*>(invstd + c4 * 4);
  float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
  const uint32_t npix = (uint32_t)N * Ho * Wo;
  const size_t row_pitch = (size_t)W * C;
  for (uint32_t pix = blockIdx.x * blockDim.y + threadIdx.y; pix < npix; pix += gridDim.x * blockDim.y) {
    const float* src;
    if (pool) {
      const uint32_t ow = pix % Wo, t = pix / Wo, oh = t % Ho, n = t / Ho;
      src = raw + ((size_t)(n * H + 2 * oh) * W + 2 * ow) * C + c4 * 4;
    } else {
      src = raw + (size_t)pix * C + c4 * 4;
    }
    const float4 gg = *reinterpret_cast<const float4*>(g + (size_t)pix * C + c4 * 4);
    BnSrc o;
    bn_route<POOL>(src, row_pitch, C, relu, gg, sc, sh, mn, iv, o);
    constexpr int np = POOL ? 4 : 1;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      if (q < np) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          s1[l] += o.gz[q][l];
          s2[l] = fmaf(o.gz[q][l], o.xh[q][l], s2[l]);
        }
      }
  }
  float* my = red + (size_t)threadIdx.y * 2 * C;
#pragma unroll
  for (int l = 0; l < 4; ++l) { my[c4 * 4 + l] = s1[l]; my[C + c4 * 4 + l] = s2[l]; }
  __syncthreads();
  // block sums -> one of kBwdSlots fp64 accumulators per channel (native double atomics; ~150 adds per address and launch)
  double* slot = acc + (size_t)(blockIdx.x % kBwdSlots) * 2 * C;
  for (int i = threadIdx.y * blockDim.x + threadIdx.x; i < 2 * C; i += blockDim.x * blockDim.y) {
    float v = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) v += red[(size_t)r * 2 * C + i];
    atomicAdd(slot + i, (double)v);
  }
}

// dgamma[c] = sum_slot acc[slot][1][c]; dbeta[c] = sum_slot acc[slot][0][c]; the accumulators are left ZEROED for the next launch
// (the workspace is persistent: no fill kernel per layer).  A few microseconds on the critical path of every layer's backward:
// 8 loads per thread instead of the 1,184 per-block partials of the first version.
__global__ void __launch_bounds__(256)
bn_bwd_finalize_kernel(double* __restrict__ acc, int C, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // over [2][C]
  if (i >= 2 * C) return;
  double v[kBwdSlots];
#pragma unroll
  for (int sl = 0; sl < kBwdSlots; ++sl) v[sl] = acc[(size_t)sl * 2 * C + i];
  double t = 0.0;
#pragma unroll
  for (int sl = 0; sl < kBwdSlots; ++sl) {
    t += v[sl];
    acc[(size_t)sl * 2 * C + i] = 0.0;
  }
  if (i < C) dbeta[i] = (float)t;
  else dgamma[i - C] = (float)t;
}

// draw = gamma*invstd * (gz - dbeta/n - xhat*dgamma/n)  -> NHWC split-bf16 (and/or fp32) at the conv-output resolution.
// Same thread mapping as bn_apply_kernel: one float4 channel group per thread, its six per-channel constants in registers.
template <int POOL>
__global__ void __launch_bounds__(256, POOL ? 3 : 5)
bn_bwd_apply_kernel(const float* __restrict__ raw, const float* __restrict__ g, int N, int H, int W, int C,
                    const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ mean, const float* __restrict__ invstd,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, int relu, int batch_stats,
                    float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_hi,
                    __nv_bfloat16* __restrict__ out_lo) {
  constexpr int pool = POOL;
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const int C4 = C / 4;
  const int tpp = C4 < 256 ? C4 : 256;
  const int ppb = 256 / tpp;
  const int lane_c = threadIdx.x % tpp, pl = threadIdx.x / tpp;
  if (pl >= ppb) return;
  // running-statistics BatchNorm (eval mode): mean / variance do not depend on the batch, their terms vanish
  const float inv_n = batch_stats ? 1.f / (float)((size_t)N * H * W) : 0.f;
  const uint32_t npix = (uint32_t)N * Ho * Wo;
  const size_t row_pitch = (size_t)W * C;
  constexpr int np = POOL ? 4 : 1;
  for (int c4 = lane_c; c4 < C4; c4 += tpp) {
    const float4 sc = *reinterpret_cast<const float4*>(scale + c4 * 4), sh = *reinterpret_cast<const float4*>(shift + c4 * 4);
    const float4 mn = *reinterpret_cast<const float4*>(mean + c4 * 4), iv = *reinterpret_cast<const float4*>(invstd + c4 * 4);
    const float4 dg = *reinterpret_cast<const float4*>(dgamma + c4 * 4), db = *reinterpret_cast<const float4*>(dbeta + c4 * 4);
    const float scv[4] = {sc.x, sc.y, sc.z, sc.w};
    // v = sc * (gz - db/n - xh * dg/n) = sc*gz - kb - xh*kg
    const float kb[4] = {sc.x * db.x * inv_n, sc.y * db.y * inv_n, sc.z * db.z * inv_n, sc.w * db.w * inv_n};
    const float kg[4] = {sc.x * dg.x * inv_n, sc.y * dg.y * inv_n, sc.z * dg.z * inv_n, sc.w * dg.w * inv_n};
    for (uint32_t pix = blockIdx.x * ppb + pl; pix < npix; pix += gridDim.x * ppb) {
      size_t off0;
      if (pool) {
        const uint32_t ow = pix % Wo, t = pix / Wo, oh = t % Ho, n = t / Ho;
        off0 = ((size_t)(n * H + 2 * oh) * W + 2 * ow) * C + c4 * 4;
      } else {
        off0 = (size_t)pix * C + c4 * 4;
      }
      const float4 gg = *reinterpret_cast<const float4*>(g + (size_t)pix * C + c4 * 4);
      BnSrc o;
      bn_route<POOL>(raw + off0, row_pitch, C, relu, gg, sc, sh, mn, iv, o);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (q < np) {
          float4 v;
          v.x = fmaf(scv[0], o.gz[q][0], -kb[0]) - o.xh[q][0] * kg[0];
          v.y = fmaf(scv[1], o.gz[q][1], -kb[1]) - o.xh[q][1] * kg[1];
          v.z = fmaf(scv[2], o.gz[q][2], -kb[2]) - o.xh[q][2] * kg[2];
          v.w = fmaf(scv[3], o.gz[q][3], -kb[3]) - o.xh[q][3] * kg[3];
          const size_t off = off0 + (size_t)(q >> 1) * row_pitch + (size_t)(q & 1) * C;
          if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = v;
          if (out_hi) {
            uint2 hi2, lo2;
            split_bf16x4(v, hi2, lo2);
            *reinterpret_cast<uint2*>(out_hi + off) = hi2;
            if (out_lo) *reinterpret_cast<uint2*>(out_lo + off) = lo2;
          }
        }
    }
  }
}

// Backward of out = max(x[b], x[b+B]): gradient goes to the first stream on ties (MaxPool3d((2,1,1))).
__global__ void pairmax_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, size_t per_stream,
                                   __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < per_stream; i += (size_t)gridDim.x * blockDim.x) {
    const float a = x[i], b = x[per_stream + i], gg = g[i];
    const bool first = !(b > a);
    __nv_bfloat16 h, l;
    split_bf16(first ? gg : 0.f, h, l);
    hi[i] = h;
    if (lo) lo[i] = l;
    split_bf16(first ? 0.f : gg, h, l);
    hi[per_stream + i] = h;
    if (lo) lo[per_stream + i] = l;
  }
}

// out[c] += sum_rows (hi+lo)[row][c]   (conv bias gradients).  blockDim = (C/4 capped at 128, rows)
__global__ void col_sum_split_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                     size_t rows, int C, float* __restrict__ out) {
  extern __shared__ float red[];  // [blockDim.y][C]
  const int C4 = C / 4;
  for (int c4 = threadIdx.x; c4 < C4; c4 += blockDim.x) {
    float s[4] = {0, 0, 0, 0};
    for (size_t r = (size_t)blockIdx.x * blockDim.y + threadIdx.y; r < rows; r += (size_t)gridDim.x * blockDim.y) {
      const uint2 h = *reinterpret_cast<const uint2*>(hi + r * C + c4 * 4);
      uint2 l = make_uint2(0, 0);
      if (lo) l = *reinterpret_cast<const uint2*>(lo + r * C + c4 * 4);
      s[0] += bf16_bits_to_float(h.x & 0xffffu) + bf16_bits_to_float(l.x & 0xffffu);
      s[1] += bf16_bits_to_float(h.x >> 16) + bf16_bits_to_float(l.x >> 16);
      s[2] += bf16_bits_to_float(h.y & 0xffffu) + bf16_bits_to_float(l.y & 0xffffu);
      s[3] += bf16_bits_to_float(h.y >> 16) + bf16_bits_to_float(l.y >> 16);
    }
#pragma unroll
    for (int l2 = 0; l2 < 4; ++l2) red[(size_t)threadIdx.y * C + c4 * 4 + l2] = s[l2];
  }
  __syncthreads();
  for (int c = threadIdx.y * blockDim.x + threadIdx.x; c < C; c += blockDim.x * blockDim.y) {
    float v = 0.f;
    for (int r = 0; r < (int)blockDim.y; ++r) v += red[(size_t)r * C + c];
    atomicAdd(out + c, v);
  }
}

}  // namespace

extern "C" int egaze_bn_finalize(const float* partial, const float* cnt, int cnt_stride, int cnt_div, int T, int C,
                                 float eps, float momentum,
                                 const float* gamma, const float* beta, float* running_mean, float* running_var,
                                 float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                                 long long* num_batches_tracked, void* stream) {
  EGAZE_CHECK_ARG(partial && cnt && T > 0 && C > 0 && cnt_stride > 0 && cnt_div > 0, "bn_finalize: bad args");
  dim3 block(32, 32), grid(ceil_div(C, 32));
  bn_finalize_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(partial, cnt, cnt_stride, cnt_div, T, C, eps, momentum, gamma, beta,
                                                               running_mean, running_var, mean_out, invstd_out,
                                                               scale_out, shift_out, num_batches_tracked);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                             const float* conv_bias, float eps, int C, float* scale, float* shift, void* stream) {
  EGAZE_CHECK_ARG(running_mean && running_var && scale && shift && C > 0, "bn_fold: bad args");
  bn_fold_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var, conv_bias,
                                                                     eps, C, scale, shift);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// rows_per_blk is fixed at 128 so the caller can size partial as [ceil(rows/128)][2][C].
extern "C" int egaze_col_stats(const float* x, long long rows, int C, float* partial, float* cnt, void* stream) {
  EGAZE_CHECK_ARG(x && partial && cnt && rows > 0 && C > 0, "col_stats: bad args");
  dim3 grid((unsigned)((rows + 127) / 128), ceil_div(C, 128)), block(128);
  col_stats_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, (int)rows, C, 128, partial, cnt);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_bn_apply(const float* x, int N, int H, int W, int C, const float* scale, const float* shift,
                              int relu, int pool, float* out_f32, void* out_hi, void* out_lo, void* out_xb, int fmt,
                              void* stream) {
  EGAZE_CHECK_ARG(x && scale && shift && (out_f32 || out_hi), "bn_apply: bad args");
  EGAZE_CHECK_ARG(C % 4 == 0, "bn_apply: C %% 4 != 0");
  EGAZE_CHECK_ARG(!pool || ((H | W) & 1) == 0, "bn_apply: pool needs even H, W");
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const size_t total = (size_t)N * Ho * Wo * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  bn_apply_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, N, H, W, C, scale, shift, relu, pool, out_f32,
                                                           (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo,
                                                           (__nv_bfloat16*)out_xb, fmt);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_pairmax(const float* x, long long per_stream, float* out, void* stream) {
  EGAZE_CHECK_ARG(x && out && per_stream > 0 && per_stream % 4 == 0, "pairmax: bad args");
  int blocks = (int)((per_stream / 4 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pairmax_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (size_t)per_stream, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// ---- backward entry points ---------------------------------------------------------------------------------------
static void bn_bwd_block(int C, dim3* block) {
  int rows = 256 / (C / 4);
  if (rows < 1) rows = 1;
  if (rows > 32) rows = 32;
  *block = dim3(C / 4, rows);
}

// partial: persistent workspace of egaze_bn_bwd_blocks() * 2 * C floats (8-byte aligned), ZERO before the first call that uses
// it; every call leaves it zeroed again.  dgamma/dbeta: [C]
extern "C" int egaze_bn_bwd_blocks(int* nblk) {
  *nblk = 2 * kBwdSlots;   // kBwdSlots fp64 accumulators per channel and sum
  return EGAZE_OK;
}

extern "C" int egaze_bn_bwd_reduce(const float* raw, const float* g, int N, int H, int W, int C, const float* scale,
                                   const float* shift, const float* mean, const float* invstd, int pool, int relu,
                                   float* partial, float* dgamma, float* dbeta, void* stream) {
  EGAZE_CHECK_ARG(raw && g && scale && shift && mean && invstd && partial && dgamma && dbeta, "bn_bwd_reduce: null");
  EGAZE_CHECK_ARG(C % 4 == 0 && C <= 4096, "bn_bwd_reduce: unsupported C=%d", C);
  EGAZE_CHECK_ARG((reinterpret_cast<uintptr_t>(partial) & 7) == 0, "bn_bwd_reduce: workspace must be 8-byte aligned");
  dim3 block;
  bn_bwd_block(C, &block);
  const int nblk = 148 * 8;
  const size_t smem = (size_t)block.y * 2 * C * sizeof(float);
  double* acc = reinterpret_cast<double*>(partial);
  if (pool) bn_bwd_reduce_kernel<1><<<nblk, block, smem, (cudaStream_t)stream>>>(raw, g, N, H, W, C, scale, shift, mean, invstd, relu, acc);
  else bn_bwd_reduce_kernel<0><<<nblk, block, smem, (cudaStream_t)stream>>>(raw, g, N, H, W, C, scale, shift, mean, invstd, relu, acc);
  EGAZE_LAUNCH_CHECK();
  bn_bwd_finalize_kernel<<<ceil_div(2 * C, 256), 256, 0, (cudaStream_t)stream>>>(acc, C, dgamma, dbeta);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_bn_bwd_apply(const float* raw, const float* g, int N, int H, int W, int C, const float* scale,
                                  const float* shift, const float* mean, const float* invstd, const float* dgamma,
                                  const float* dbeta, int pool, int relu, int batch_stats, float* out_f32, void* out_hi,
                                  void* out_lo, void* stream) {
  EGAZE_CHECK_ARG(raw && g && dgamma && dbeta && (out_f32 || out_hi), "bn_bwd_apply: null");
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const size_t total = (size_t)N * Ho * Wo * (C / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (pool)
    bn_bwd_apply_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(raw, g, N, H, W, C, scale, shift, mean, invstd, dgamma, dbeta, relu,
                                                                    batch_stats, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  else
    bn_bwd_apply_kernel<0><<<blocks, 256, 0, (cudaStream_t)stream>>>(raw, g, N, H, W, C, scale, shift, mean, invstd, dgamma, dbeta, relu,
                                                                    batch_stats, out_f32, (__nv_bfloat16*)out_hi, (__nv_bfloat16*)out_lo);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_pairmax_bwd(const float* x, const float* g, long long per_stream, void* hi, void* lo, void* stream) {
  EGAZE_CHECK_ARG(x && g && hi && per_stream > 0, "pairmax_bwd: bad args");
  int blocks = (int)((per_stream + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  pairmax_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, g, (size_t)per_stream, (__nv_bfloat16*)hi,
                                                               (__nv_bfloat16*)lo);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// out ([C] fp32) is ACCUMULATED into.
extern "C" int egaze_col_sum_split(const void* hi, const void* lo, long long rows, int C, float* out, void* stream) {
  EGAZE_CHECK_ARG(hi && out && rows > 0 && C % 4 == 0, "col_sum_split: bad args");
  int bx = C / 4;
  if (bx > 128) bx = 128;
  int by = 256 / bx;
  if (by < 1) by = 1;
  dim3 block(bx, by);
  int blocks = (int)((rows + by - 1) / by);
  if (blocks > 148 * 8) blocks = 148 * 8;   // HBM-bound: 205 MB per launch at the last decoder layer
  const size_t smem = (size_t)by * C * sizeof(float);
  col_sum_split_kernel<<<blocks, block, smem, (cudaStream_t)stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo,
                                                                      (size_t)rows, C, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
