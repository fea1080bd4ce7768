// Hardware probe (not product code): can a tcgen05 K-major SWIZZLE_128B operand descriptor start at an arbitrary
// 128-byte row of a TMA-written window and step its 8-row groups by a stride that is NOT a multiple of 1024 B?
//
// If yes, the 3x3 conv can fetch ONE (BH+2) x (BW+2) activation window per K chunk and address all nine taps inside
// it (row offset (r*(BW+2) + s) * 128 B, SBO = (BW+2) * 128 B with BW = 8) instead of one window per horizontal tap.
//
// The probe multiplies the gathered A rows by an identity B, so D[m][n] must equal A[row(m)][n] exactly, for every
// tap and for two descriptor variants: base_offset = 0 and base_offset = (start >> 7) & 7.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I egocentric-gaze-prediction_b200/csrc \
//        tools/probe_swizzle_shift.cu egocentric-gaze-prediction_b200/csrc/runtime.cu -o tools/probe_swizzle_shift
#include "common.cuh"
#include <vector>
#include <stdlib.h>

constexpr int WH = 18, WW = 10, C = 64;   // window: 18 rows x 10 px x 64 channels (128 B per pixel)

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int sbo_rows) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_smem = smem;                    // 180 rows x 128 B = 23040 B -> pad to 24576
  uint8_t* b_smem = smem + 24576;            // 64 rows x 128 B
  __shared__ uint64_t full, done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&full, 1);
    ptx::mbar_init(&done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_smem, 64);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  uint32_t full_par = 0, done_par = 0;
  int slot = 0;
  for (int img = 0; img < 2; ++img) {
    if (threadIdx.x == 0) {
      ptx::mbar_arrive_expect_tx(&full, (uint32_t)(WH * WW * 128 + 64 * 128));
      ptx::tma_load_4d(a_smem, &tmA, &full, 0, 0, 0, img);
      ptx::tma_load_2d(b_smem, &tmB, &full, 0, 0);
    }
    ptx::mbar_wait(&full, full_par);
    full_par ^= 1;
    for (int variant = 0; variant < 2; ++variant)
      for (int r = 0; r < 3; ++r)
        for (int s = 0; s < 3; ++s) {
          if (threadIdx.x == 0) {
            ptx::tc_fence_after();
            const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
            const uint32_t a_addr = ptx::smem_u32(a_smem) + (uint32_t)((r * WW + s) * 128);
            uint64_t ad = ptx::make_smem_desc(a_addr, 16, (uint32_t)sbo_rows * 128u, 128);
            if (variant == 1) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
            const uint64_t bd = ptx::make_smem_desc(ptx::smem_u32(b_smem), 16, 1024, 128);
            for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k > 0);
            ptx::umma_commit(&done);
          }
          ptx::mbar_wait(&done, done_par);
          done_par ^= 1;
          ptx::tc_fence_after();
          for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t v[32];
            ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
            float* dst = out + ((size_t)slot * 128 + warp * 32 + lane) * 64 + c0;
            for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(v[j]);
          }
          ptx::tc_fence_before();
          __syncthreads();
          ++slot;
        }
  }
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 64);
}

int main() {
  std::vector<__nv_bfloat16> hA((size_t)2 * WH * WW * C), hB(64 * 64);
  for (int q = 0; q < WH * WW; ++q)
    for (int c = 0; c < C; ++c) {
      hA[(size_t)q * C + c] = __float2bfloat16((float)q);                          // image 0: the row index
      hA[(size_t)(WH * WW + q) * C + c] = __float2bfloat16((float)c);              // image 1: the channel index
    }
  for (int n = 0; n < 64; ++n)
    for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB;
  float* dOut;
  const int slots = 2 * 2 * 9;
  cudaMalloc(&dA, hA.size() * 2);
  cudaMalloc(&dB, hB.size() * 2);
  cudaMalloc(&dOut, (size_t)slots * 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  {
    uint64_t dims[4] = {C, WW, WH, 2};
    uint64_t str[3] = {C * 2, WW * C * 2, (uint64_t)WH * WW * C * 2};
    uint32_t box[4] = {C, WW, WH, 1};
    if (egaze_encode_tmap(&tmA, dA, 4, dims, str, box, 128, 2)) { printf("tmap A failed\n"); return 1; }
  }
  {
    uint64_t dims[2] = {64, 64};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 64};
    if (egaze_encode_tmap(&tmB, dB, 2, dims, str, box, 128, 2)) { printf("tmap B failed\n"); return 1; }
  }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> hOut((size_t)slots * 128 * 64);
  for (int sbo_rows : {WW, 8}) {   // 10-row group stride (the scheme under test) and the classic dense 8-row stride
    cudaMemset(dOut, 0xff, hOut.size() * 4);
    probe_kernel<<<1, 128, 40 * 1024, 0>>>(tmA, tmB, dOut, sbo_rows);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sbo_rows=%d: kernel failed: %s\n", sbo_rows, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(hOut.data(), dOut, hOut.size() * 4, cudaMemcpyDeviceToHost);
    int slot = 0;
    for (int img = 0; img < 2; ++img)
      for (int variant = 0; variant < 2; ++variant)
        for (int r = 0; r < 3; ++r)
          for (int s = 0; s < 3; ++s, ++slot) {
            int bad = 0, first_m = -1;
            float first_got = 0, first_exp = 0;
            for (int m = 0; m < 128; ++m)
              for (int n = 0; n < 64; ++n) {
                const int row = (m / 8) * sbo_rows + (m % 8) + r * WW + s;
                if (row >= WH * WW) continue;   // dense stride walks off the window for large r: ignore
                const float exp = img == 0 ? (float)row : (float)n;
                const float got = hOut[((size_t)slot * 128 + m) * 64 + n];
                if (got != exp) {
                  if (!bad) { first_m = m * 64 + n; first_got = got; first_exp = exp; }
                  ++bad;
                }
              }
            printf("PROBE sbo_rows=%d img=%d base_offset_variant=%d r=%d s=%d : %s (%d mismatches", sbo_rows, img, variant, r, s,
                   bad ? "MISMATCH" : "ok", bad);
            if (bad) printf("; first at m=%d n=%d got %.1f expected %.1f", first_m / 64, first_m % 64, first_got, first_exp);
            printf(")\n");
          }
  }
  return 0;
}
