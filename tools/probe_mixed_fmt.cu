// Hardware probe (not product code): does tcgen05.mma kind::f16 accept DIFFERENT element formats for A and B
// (idesc a_format = BF16, b_format = F16 and vice versa)?  D = A x I must reproduce A exactly if it does.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I egocentric-gaze-prediction_b200/csrc \
//        tools/probe_mixed_fmt.cu egocentric-gaze-prediction_b200/csrc/runtime.cu -o tools/probe_mixed_fmt
#include "common.cuh"
#include <cuda_fp16.h>
#include <vector>

__global__ void __launch_bounds__(128, 1)
mixed_kernel(const uint16_t* __restrict__ gA, const uint16_t* __restrict__ gB, uint32_t idesc, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint16_t* a_s = reinterpret_cast<uint16_t*>(smem);            // 128 rows x 64 K (128 B rows), SWIZZLE_128B written by hand
  uint16_t* b_s = reinterpret_cast<uint16_t*>(smem + 16384);    // 64 rows x 64 K
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // K-major SWIZZLE_128B: element (row, k) lives at row*128 + ((k/8) ^ (row & 7))*16 + (k%8)*2 bytes
  for (int i = threadIdx.x; i < 128 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    a_s[r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)] = gA[i];
  }
  for (int i = threadIdx.x; i < 64 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    b_s[r * 64 + (((k >> 3) ^ (r & 7)) << 3) + (k & 7)] = gB[i];
  }
  if (threadIdx.x == 0) { ptx::mbar_init(&done, 1); ptx::fence_barrier_init(); }
  if (warp == 0) { ptx::tmem_alloc(&tmem_base_smem, 64); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) {
    const uint64_t ad = ptx::make_smem_desc(ptx::smem_u32(a_s), 16, 1024, 128);
    const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b_s), 16, 1024, 128);
    for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k > 0);
    ptx::umma_commit(&done);
  }
  ptx::mbar_wait(&done, 0);
  ptx::tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t v[32];
    ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

static uint16_t f2bf(float f) { __nv_bfloat16 b = __float2bfloat16(f); return *reinterpret_cast<uint16_t*>(&b); }
static uint16_t f2h(float f) { __half h = __float2half(f); return *reinterpret_cast<uint16_t*>(&h); }

int main() {
  // formats: 0 = F16, 1 = BF16 (idesc bits 7-9: A, 10-12: B)
  for (int afmt = 0; afmt < 2; ++afmt)
    for (int bfmt = 0; bfmt < 2; ++bfmt) {
      std::vector<uint16_t> hA(128 * 64), hB(64 * 64);
      std::vector<float> ref(128 * 64);
      for (int r = 0; r < 128; ++r)
        for (int k = 0; k < 64; ++k) {
          const float v = (float)((r * 7 + k * 3) % 97) * 0.0625f - 3.f;   // exact in both formats
          ref[r * 64 + k] = v;
          hA[r * 64 + k] = afmt ? f2bf(v) : f2h(v);
        }
      for (int n = 0; n < 64; ++n)
        for (int k = 0; k < 64; ++k) hB[n * 64 + k] = bfmt ? f2bf(n == k ? 1.f : 0.f) : f2h(n == k ? 1.f : 0.f);
      uint16_t *dA, *dB; float* dO;
      cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
      cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
      cudaMemset(dO, 0xff, 128 * 64 * 4);
      const uint32_t idesc = (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      cudaFuncSetAttribute(mixed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
      mixed_kernel<<<1, 128, 32 * 1024>>>(dA, dB, idesc, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("MIXED A=%s B=%s : kernel failed: %s\n", afmt ? "bf16" : "f16", bfmt ? "bf16" : "f16", cudaGetErrorString(e)); return 1; }
      std::vector<float> hO(128 * 64);
      cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
      int bad = 0;
      for (size_t i = 0; i < hO.size(); ++i) bad += hO[i] != ref[i];
      printf("MIXED A=%s B=%s : %s (%d mismatches, out[1]=%f ref[1]=%f)\n", afmt ? "bf16" : "f16", bfmt ? "bf16" : "f16", bad ? "MISMATCH" : "ok", bad,
             hO[1], ref[1]);
      cudaFree(dA); cudaFree(dB); cudaFree(dO);
    }
  return 0;
}
