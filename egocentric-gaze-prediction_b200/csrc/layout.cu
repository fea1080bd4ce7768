// Layout kernels: the reference API is NCHW fp32 (SURVEY 8b "Tensor conventions"); internally the
// conv kernels consume NHWC split-bf16 activations and [tap][Cout][Cin_p] split-bf16 weights.
// All of these are HBM-bound element shuffles: 32x32 smem-tile transposes so both sides are coalesced.
#include "common.cuh"

namespace {

// x: [N][C][HW] fp32  ->  hi/lo: [N][HW][Cp] bf16 (channels >= C zero-filled)
// fmt: 0 = bf16 planes, 1 = fp16 planes; xb: optional extra bf16(x) plane (what the weight-gradient GEMM reads when the
// forward planes are fp16)
// (the xb plane has its own channel stride Cp_xb: the weight-gradient GEMM wants 64-channel rows, the forward conv only 16 / 32)
__global__ void nchw_to_nhwc_split_kernel(const float* __restrict__ x, int C, int HW, int Cp,
                                          __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                          __nv_bfloat16* __restrict__ xb, int Cp_xb, int fmt) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* xn = x + (size_t)n * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? xn[(size_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && xb && c < Cp_xb) xb[((size_t)n * HW + p) * Cp_xb + c] = __float2bfloat16_rn(tile[threadIdx.x][i]);
    if (p < HW && c < Cp) {
      const float v = tile[threadIdx.x][i];
      const size_t o = ((size_t)n * HW + p) * Cp + c;
      if (fmt) {
        __half h, l;
        split_f16(v, h, l);
        reinterpret_cast<__half*>(hi)[o] = h;
        if (lo) reinterpret_cast<__half*>(lo)[o] = l;
      } else {
        __nv_bfloat16 h, l;
        split_bf16(v, h, l);
        hi[o] = h;
        if (lo) lo[o] = l;
      }
    }
  }
}

// The same for C <= 32 input channels (the module-boundary case: RGB 3 -> 16, flow 20 -> 32, with a 64-channel bf16 copy for the
// first layer's weight gradient): one block converts 128 pixels x all channels, reads are 512-byte runs along the pixel axis,
// every store is a 16-byte vector (8 channels) and consecutive lanes cover consecutive 16-byte pieces -- the padded planes
// are mostly zeros (xb: 61 of 64 channels for RGB), so the kernel is bound by its writes (868 MB per B = 32 step at 224^2).
__global__ void __launch_bounds__(256)
nchw_to_nhwc_split_small_kernel(const float* __restrict__ x, int C, int HW, int Cp, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo, __nv_bfloat16* __restrict__ xb, int Cp_xb, int fmt) {
  __shared__ float tile[32][129];
  const int n = blockIdx.y, p0 = blockIdx.x * 128;
  const float* xn = x + (size_t)n * C * HW;
  for (int i = threadIdx.x; i < 32 * 128; i += 256) {
    const int c = i >> 7, pp = i & 127;
    tile[c][pp] = (c < C && p0 + pp < HW) ? __ldg(xn + (size_t)c * HW + p0 + pp) : 0.f;
  }
  __syncthreads();
  const int G = Cp >> 3;                 // 16-byte groups per pixel in the hi / lo planes
  for (int i = threadIdx.x; i < 128 * G; i += 256) {
    const int pp = i / G, g = i - pp * G;
    if (p0 + pp >= HW) continue;
    const float4 a = make_float4(tile[g * 8 + 0][pp], tile[g * 8 + 1][pp], tile[g * 8 + 2][pp], tile[g * 8 + 3][pp]);
    const float4 b = make_float4(tile[g * 8 + 4][pp], tile[g * 8 + 5][pp], tile[g * 8 + 6][pp], tile[g * 8 + 7][pp]);
    uint2 ha, la, hb, lb;
    if (fmt) { split_f16x4(a, ha, la); split_f16x4(b, hb, lb); }
    else { split_bf16x4(a, ha, la); split_bf16x4(b, hb, lb); }
    const size_t o = ((size_t)n * HW + p0 + pp) * Cp + g * 8;
    *reinterpret_cast<uint4*>(hi + o) = make_uint4(ha.x, ha.y, hb.x, hb.y);
    if (lo) *reinterpret_cast<uint4*>(lo + o) = make_uint4(la.x, la.y, lb.x, lb.y);
  }
  if (xb) {
    const int Gx = Cp_xb >> 3;
    for (int i = threadIdx.x; i < 128 * Gx; i += 256) {
      const int pp = i / Gx, g = i - pp * Gx;
      if (p0 + pp >= HW) continue;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (g * 8 < C) {
        const uint2 u0 = pack_bf16x4(make_float4(tile[g * 8 + 0][pp], tile[g * 8 + 1][pp], tile[g * 8 + 2][pp], tile[g * 8 + 3][pp]));
        const uint2 u1 = pack_bf16x4(make_float4(tile[g * 8 + 4][pp], tile[g * 8 + 5][pp], tile[g * 8 + 6][pp], tile[g * 8 + 7][pp]));
        v = make_uint4(u0.x, u0.y, u1.x, u1.y);
      }
      *reinterpret_cast<uint4*>(xb + ((size_t)n * HW + p0 + pp) * Cp_xb + g * 8) = v;
    }
  }
}

// hi/lo (or f32): [N][HW][Cs] (channel stride Cs >= C)  ->  out: [N][C][HW] fp32
__global__ void nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                    const float* __restrict__ f32, int C, int HW, int Cs, int fmt, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (p < HW && c < C) {
      const size_t o = ((size_t)n * HW + p) * Cs + c;
      if (f32) v = f32[o];
      else {
        v = dec16(reinterpret_cast<const unsigned short*>(hi)[o], fmt);
        if (lo) v += dec16(reinterpret_cast<const unsigned short*>(lo)[o], fmt);
      }
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) out[((size_t)n * C + c) * HW + p] = tile[threadIdx.x][i];
  }
}

// x: [N][C][HW] fp32 -> out: [N][HW][C] fp32
__global__ void nchw_to_nhwc_f32_kernel(const float* __restrict__ x, int C, int HW, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, p = p0 + threadIdx.x;
    tile[i][threadIdx.x] = (c < C && p < HW) ? x[((size_t)n * C + c) * HW + p] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int p = p0 + i, c = c0 + threadIdx.x;
    if (p < HW && c < C) out[((size_t)n * HW + p) * C + c] = tile[threadIdx.x][i];
  }
}

// fp32 NHWC -> split planes (same shape), elementwise
__global__ void f32_to_split_kernel(const float* __restrict__ x, size_t n, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, int fmt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    if (fmt) {
      __half h, l;
      split_f16(x[i], h, l);
      reinterpret_cast<__half*>(hi)[i] = h;
      if (lo) reinterpret_cast<__half*>(lo)[i] = l;
    } else {
      __nv_bfloat16 h, l;
      split_bf16(x[i], h, l);
      hi[i] = h;
      if (lo) lo[i] = l;
    }
  }
}

// w: OIHW fp32 [Co][Ci][3][3].
//  mode 0 (fprop) : out[(r*3+s)][co][ci]            rows = Co, cols = Ci_p  (ci >= Ci zero)
//  mode 1 (dgrad) : out[((2-r)*3+(2-s))][ci][co]    rows = Ci, cols = Co_p  (co >= Co zero)
// One thread per (row, col): it reads the nine taps of its weight (36 contiguous bytes; in mode 0 adjacent threads read
// adjacent runs) and writes one element of each tap plane (adjacent threads write adjacent bf16).
// fmt 1: fp16 planes of w * kF16WScale (the forward operand format; the conv multiplies its accumulators by 1 / kF16WScale):
// with weights around 1e-2 the unscaled lo plane would sit in fp16's subnormal range and keep only ~19 bits of the weight.
constexpr float kF16WScale = 256.f;
//  mode 2 / 3: sub-pixel copies (16 planes, egaze_subpixel_taps in common.cuh); 2: [plane][co][ci], 3: [plane][ci][co]
__device__ __forceinline__ void pack_w3x3_body(const float* __restrict__ w, int Co, int Ci, int rows, int cols_p, int mode,
                                               int fmt, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                               uint32_t first, uint32_t stride) {
  const uint32_t total = (uint32_t)rows * cols_p;
  const size_t plane = (size_t)rows * cols_p;
  for (uint32_t i = first; i < total; i += stride) {
    const uint32_t col = i % cols_p, row = i / cols_p;
    float v[9];
    const bool fwd = (mode & 1) == 0;
    const bool live = fwd ? col < (uint32_t)Ci : col < (uint32_t)Co;
    const float* src = fwd ? w + ((size_t)row * Ci + col) * 9 : w + ((size_t)col * Ci + row) * 9;
#pragma unroll
    for (int t = 0; t < 9; ++t) v[t] = live ? src[t] : 0.f;
    if (mode >= 2) {
      float q[16];
      egaze_subpixel_taps(v, q);
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        if (fmt) {
          __half h, l;
          split_f16(q[t] * kF16WScale, h, l);
          reinterpret_cast<__half*>(hi)[(size_t)t * plane + i] = h;
          if (lo) reinterpret_cast<__half*>(lo)[(size_t)t * plane + i] = l;
        } else {
          __nv_bfloat16 h, l;
          split_bf16(q[t], h, l);
          hi[(size_t)t * plane + i] = h;
          if (lo) lo[(size_t)t * plane + i] = l;
        }
      }
      continue;
    }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (fmt) {
        __half h, l;
        split_f16(v[mode == 0 ? t : 8 - t] * kF16WScale, h, l);
        reinterpret_cast<__half*>(hi)[(size_t)t * plane + i] = h;
        if (lo) reinterpret_cast<__half*>(lo)[(size_t)t * plane + i] = l;
      } else {
        __nv_bfloat16 h, l;
        split_bf16(v[mode == 0 ? t : 8 - t], h, l);
        hi[(size_t)t * plane + i] = h;
        if (lo) lo[(size_t)t * plane + i] = l;
      }
    }
  }
}

__global__ void pack_w3x3_kernel(const float* __restrict__ w, int Co, int Ci, int rows, int cols_p, int mode, int fmt,
                                 __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  pack_w3x3_body(w, Co, Ci, rows, cols_p, mode, fmt, hi, lo, blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}

// The same for MANY weights in one launch (after an optimiser step every packed copy is stale): blockIdx.y = job.
struct PackJob {
  const float* w;
  __nv_bfloat16* hi;
  __nv_bfloat16* lo;
  int Co, Ci, rows, cols_p, mode, fmt;
};
__global__ void pack_w3x3_multi_kernel(const PackJob* __restrict__ jobs) {
  const PackJob j = jobs[blockIdx.y];
  pack_w3x3_body(j.w, j.Co, j.Ci, j.rows, j.cols_p, j.mode, j.fmt, j.hi, j.lo, blockIdx.x * blockDim.x + threadIdx.x,
                 gridDim.x * blockDim.x);
}

// dwp: [9][Co_p][Ci_p] fp32 (wgrad accumulator)  ->  gw: OIHW [Co][Ci][3][3]  (gw = beta*gw + dwp).
// One thread per (co, ci): nine strided reads, one 36-byte contiguous store (adjacent threads store adjacent 36-byte runs).
// sub: dwp holds the 16 sub-pixel planes; tap (r, s) collects the planes it was pre-summed into.
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwp, int Co, int Ci, int Co_p, int Ci_p, float beta, int sub,
                                    float* __restrict__ gw) {
  const uint32_t total = (uint32_t)Co * Ci;
  const size_t tap_stride = (size_t)Co_p * Ci_p;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t ci = i % Ci, co = i / Ci;
    const float* src = dwp + (size_t)co * Ci_p + ci;
    float* dst = gw + (size_t)i * 9;
    if (sub) {
      float q[16], g9[9];
#pragma unroll
      for (int t = 0; t < 16; ++t) q[t] = src[t * tap_stride];
      egaze_subpixel_taps_transpose(q, g9);
#pragma unroll
      for (int tap = 0; tap < 9; ++tap) dst[tap] = beta == 0.f ? g9[tap] : fmaf(beta, dst[tap], g9[tap]);
      continue;
    }
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const float v = src[tap * tap_stride];
      dst[tap] = beta == 0.f ? v : fmaf(beta, dst[tap], v);
    }
  }
}

}  // namespace

extern "C" int egaze_nchw_to_nhwc_split(const float* x, int N, int C, int H, int W, int Cp, void* hi, void* lo, void* xb,
                                        int Cp_xb, int fmt, void* stream) {
  EGAZE_CHECK_ARG(x && hi && Cp >= C && (!xb || Cp_xb >= C), "nchw_to_nhwc_split: bad args");
  const int HW = H * W;
  if (C <= 32 && Cp <= 32 && Cp % 8 == 0 && (!xb || Cp_xb % 8 == 0) && N <= 65535) {
    dim3 grid(ceil_div(HW, 128), N);
    nchw_to_nhwc_split_small_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, C, HW, Cp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo,
                                                                           (__nv_bfloat16*)xb, Cp_xb, fmt);
    EGAZE_LAUNCH_CHECK();
    return EGAZE_OK;
  }
  const int cmax = (xb && Cp_xb > Cp) ? Cp_xb : Cp;
  dim3 grid(ceil_div(HW, 32), ceil_div(cmax, 32), N), block(32, 8);
  nchw_to_nhwc_split_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, C, HW, Cp, (__nv_bfloat16*)hi,
                                                                      (__nv_bfloat16*)lo, (__nv_bfloat16*)xb, Cp_xb, fmt);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_nhwc_to_nchw(const void* hi, const void* lo, const float* f32, int N, int C, int H, int W, int Cs,
                                  int fmt, float* out, void* stream) {
  EGAZE_CHECK_ARG((hi || f32) && out && Cs >= C, "nhwc_to_nchw: bad args");
  const int HW = H * W;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N), block(32, 8);
  nhwc_to_nchw_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)hi, (const __nv_bfloat16*)lo, f32,
                                                                C, HW, Cs, fmt, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_nchw_to_nhwc_f32(const float* x, int N, int C, int H, int W, float* out, void* stream) {
  EGAZE_CHECK_ARG(x && out, "nchw_to_nhwc_f32: bad args");
  const int HW = H * W;
  dim3 grid(ceil_div(HW, 32), ceil_div(C, 32), N), block(32, 8);
  nchw_to_nhwc_f32_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(x, C, HW, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_f32_to_split(const float* x, long long n, void* hi, void* lo, int fmt, void* stream) {
  EGAZE_CHECK_ARG(x && hi && n >= 0, "f32_to_split: bad args");
  if (n == 0) return EGAZE_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  f32_to_split_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, fmt);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_f16_weight_scale(float* out) {
  if (out) *out = kF16WScale;
  return EGAZE_OK;
}

extern "C" int egaze_pack_w3x3(const float* w_oihw, int Cout, int Cin, int cols_p, int mode, int fmt, void* hi, void* lo,
                               void* stream) {
  EGAZE_CHECK_ARG(w_oihw && hi, "pack_w3x3: bad args");
  EGAZE_CHECK_ARG(mode >= 0 && mode <= 3, "pack_w3x3: mode must be 0..3");
  const int rows = (mode & 1) == 0 ? Cout : Cin;
  EGAZE_CHECK_ARG(cols_p >= ((mode & 1) == 0 ? Cin : Cout), "pack_w3x3: cols_p too small");
  const size_t total = (size_t)rows * cols_p;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  pack_w3x3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w_oihw, Cout, Cin, rows, cols_p, mode, fmt, (__nv_bfloat16*)hi,
                                                             (__nv_bfloat16*)lo);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// jobs: device array of njobs records {const float* w; void* hi; void* lo; int Cout, Cin, rows, cols_p, mode, pad;}
// (48 bytes each, the fields of egaze_pack_w3x3).
extern "C" int egaze_pack_w3x3_multi(const void* jobs, int njobs, void* stream) {
  EGAZE_CHECK_ARG(jobs && njobs > 0, "pack_w3x3_multi: bad args");
  dim3 grid(64, (unsigned)njobs);
  pack_w3x3_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const PackJob*)jobs);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_unpack_wgrad(float* dwp, int Cout, int Cin, int Cout_p, int Cin_p, float beta, int clear, int sub,
                                  float* gw_oihw, void* stream) {
  EGAZE_CHECK_ARG(dwp && gw_oihw && Cout_p >= Cout && Cin_p >= Cin, "unpack_wgrad: bad args");
  const size_t total = (size_t)Cout * Cin;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  unpack_wgrad_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dwp, Cout, Cin, Cout_p, Cin_p, beta, sub, gw_oihw);
  EGAZE_LAUNCH_CHECK();
  // clear != 0: leave the (persistent) accumulator zeroed for the next accumulation -- a stream-ordered fill at full bandwidth
  if (clear) EGAZE_CUDA(cudaMemsetAsync(dwp, 0, (size_t)(sub ? 16 : 9) * Cout_p * Cin_p * sizeof(float), (cudaStream_t)stream));
  return EGAZE_OK;
}
