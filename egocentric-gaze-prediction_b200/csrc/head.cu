// The 1x1 conv to one channel + sigmoid that ends the SP decoder (reference models/model_SP.py:30,32,49)
// and the LF net (models/late_fusion.py:13,15,22).  AI ~ 1 FLOP/B -> purely HBM-bound: one pixel per thread,
// 16-byte loads over the pixel's contiguous NHWC channel run, weights in shared memory.
#include "common.cuh"

namespace {

// x: NHWC split-bf16 [P][Cs]; w: [C] fp32 (C <= Cs); out[p] = sigmoid(sum_c (hi+lo)[p][c]*w[c] + b)
template <int CS>
__global__ void head_fwd_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                const float* __restrict__ w, const float* __restrict__ b, int C, size_t P, int fmt,
                                float* __restrict__ out, float* __restrict__ logit_out) {
  __shared__ float ws[CS];
  for (int i = threadIdx.x; i < CS; i += blockDim.x) ws[i] = i < C ? w[i] : 0.f;
  __syncthreads();
  const float bias = b ? b[0] : 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const uint4* ph = reinterpret_cast<const uint4*>(hi + p * CS);
    const uint4* pl = lo ? reinterpret_cast<const uint4*>(lo + p * CS) : nullptr;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CS / 8; ++i) {
      const uint4 h = __ldg(ph + i);
      uint4 l = make_uint4(0, 0, 0, 0);
      if (pl) l = __ldg(pl + i);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x0 = dec16(hw[q] & 0xffffu, fmt) + dec16(lw[q] & 0xffffu, fmt);
        const float x1 = dec16(hw[q] >> 16, fmt) + dec16(lw[q] >> 16, fmt);
        acc = fmaf(x0, ws[i * 8 + q * 2], acc);
        acc = fmaf(x1, ws[i * 8 + q * 2 + 1], acc);
      }
    }
    acc += bias;
    if (logit_out) logit_out[p] = acc;
    out[p] = 1.f / (1.f + expf(-acc));
  }
}

// Backward of y = sigmoid(x.w + b):  dz = gy*y*(1-y);  dx[p][c] = dz*w[c] (written split-bf16, masked by relu of x);
// dw[c] += sum_p dz*x[p][c];  db += sum_p dz.   Block-level reduction of dw/db, then one atomicAdd per block.
template <int CS>
__global__ void head_bwd_kernel(const __nv_bfloat16* __restrict__ hi, const __nv_bfloat16* __restrict__ lo,
                                const float* __restrict__ w, int C, size_t P, int fmt, const float* __restrict__ y,
                                const float* __restrict__ gy, int relu_mask, __nv_bfloat16* __restrict__ dx_hi,
                                __nv_bfloat16* __restrict__ dx_lo, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float ws[CS];
  __shared__ float red[CS + 1];
  for (int i = threadIdx.x; i < CS; i += blockDim.x) { ws[i] = i < C ? w[i] : 0.f; red[i] = 0.f; }
  if (threadIdx.x == 0) red[CS] = 0.f;
  __syncthreads();
  float dwl[CS];
#pragma unroll
  for (int i = 0; i < CS; ++i) dwl[i] = 0.f;
  float dbl = 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float yy = y[p];
    const float dz = gy[p] * yy * (1.f - yy);
    dbl += dz;
    const uint4* ph = reinterpret_cast<const uint4*>(hi + p * CS);
    const uint4* pl = lo ? reinterpret_cast<const uint4*>(lo + p * CS) : nullptr;
#pragma unroll
    for (int i = 0; i < CS / 8; ++i) {
      const uint4 h = __ldg(ph + i);
      uint4 l = make_uint4(0, 0, 0, 0);
      if (pl) l = __ldg(pl + i);
      const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
      uint32_t oh[4], ol[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x0 = dec16(hw[q] & 0xffffu, fmt) + dec16(lw[q] & 0xffffu, fmt);
        const float x1 = dec16(hw[q] >> 16, fmt) + dec16(lw[q] >> 16, fmt);
        dwl[i * 8 + q * 2] = fmaf(dz, x0, dwl[i * 8 + q * 2]);
        dwl[i * 8 + q * 2 + 1] = fmaf(dz, x1, dwl[i * 8 + q * 2 + 1]);
        float g0 = dz * ws[i * 8 + q * 2], g1 = dz * ws[i * 8 + q * 2 + 1];
        if (relu_mask) {
          if (!(x0 > 0.f)) g0 = 0.f;
          if (!(x1 > 0.f)) g1 = 0.f;
        }
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(g0, h0, l0);
        split_bf16(g1, h1, l1);
        oh[q] = pack_bf16x2(h0, h1);
        ol[q] = pack_bf16x2(l0, l1);
      }
      if (dx_hi) reinterpret_cast<uint4*>(dx_hi + p * CS)[i] = make_uint4(oh[0], oh[1], oh[2], oh[3]);
      if (dx_lo) reinterpret_cast<uint4*>(dx_lo + p * CS)[i] = make_uint4(ol[0], ol[1], ol[2], ol[3]);
    }
  }
  // warp reduce, then smem atomics, then one global atomic per channel per block
#pragma unroll
  for (int i = 0; i < CS; ++i) {
    const float v = warp_sum(dwl[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&red[i], v);
  }
  dbl = warp_sum(dbl);
  if ((threadIdx.x & 31) == 0) atomicAdd(&red[CS], dbl);
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(dw + i, red[i]);
  if (threadIdx.x == 0 && db) atomicAdd(db, red[CS]);
}

}  // namespace

extern "C" int egaze_head_fwd(const void* x_hi, const void* x_lo, int fmt, const float* w, const float* b, int C, int Cs,
                              long long P, float* out, float* logit_out, void* stream) {
  EGAZE_CHECK_ARG(x_hi && w && out && P > 0, "head_fwd: bad args");
  EGAZE_CHECK_ARG((Cs == 64 || Cs == 16) && C <= Cs, "head_fwd: channel stride must be 64 or 16 (got %d)", Cs);
  int blocks = (int)((P + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (Cs == 64)
    head_fwd_kernel<64><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w,
                                                                  b, C, (size_t)P, fmt, out, logit_out);
  else
    head_fwd_kernel<16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w,
                                                                  b, C, (size_t)P, fmt, out, logit_out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// dw ([C]) and db ([1]) are ACCUMULATED into (caller zeroes them).
extern "C" int egaze_head_bwd(const void* x_hi, const void* x_lo, int fmt, const float* w, int C, int Cs, long long P,
                              const float* y, const float* gy, int relu_mask, void* dx_hi, void* dx_lo, float* dw,
                              float* db, void* stream) {
  EGAZE_CHECK_ARG(x_hi && w && y && gy && dw && P > 0, "head_bwd: bad args");
  EGAZE_CHECK_ARG((Cs == 64 || Cs == 16) && C <= Cs, "head_bwd: channel stride must be 64 or 16 (got %d)", Cs);
  int blocks = (int)((P + 127) / 128);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (Cs == 64)
    head_bwd_kernel<64><<<blocks, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w,
                                                                  C, (size_t)P, fmt, y, gy, relu_mask, (__nv_bfloat16*)dx_hi,
                                                                  (__nv_bfloat16*)dx_lo, dw, db);
  else
    head_bwd_kernel<16><<<blocks, 128, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x_hi, (const __nv_bfloat16*)x_lo, w,
                                                                  C, (size_t)P, fmt, y, gy, relu_mask, (__nv_bfloat16*)dx_hi,
                                                                  (__nv_bfloat16*)dx_lo, dw, db);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
