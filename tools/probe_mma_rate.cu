// Hardware probe (not product code): cycles per tcgen05.mma (kind::f16, M = 128, K = 16, SS operands) as a function
// of N and of the operand layout, issued back to back by one elected lane on every SM at once.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I egocentric-gaze-prediction_b200/csrc \
//        tools/probe_mma_rate.cu egocentric-gaze-prediction_b200/csrc/runtime.cu -o tools/probe_mma_rate
#include "common.cuh"
#include <vector>

struct Cfg {
  int n1, n2;       // N of the two MMAs issued alternately (n2 = 0: only the first)
  int a_mn, b_mn;   // 0 = K-major, 1 = MN-major
  int same_a;       // second MMA reads the same A tile (1) or another one (0)
  int d2_off;       // TMEM column offset of the second MMA's accumulator
};

__global__ void __launch_bounds__(128, 1) rate_kernel(Cfg c, int iters, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(&tmem_base_smem, 512);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 1) {
    const bool leader = ptx::elect_one();
    // A tiles: two planes of 128 rows x 128 B at 0 and 16 KB; B: up to 256 rows x 128 B at 32 KB
    const uint32_t a0 = ptx::smem_u32(smem), a1 = a0 + 16384, b0 = a0 + 32768;
    const uint64_t ad0 = c.a_mn ? ptx::make_smem_desc(a0, 8192, 1024, 128) : ptx::make_smem_desc(a0, 16, 1024, 128);
    const uint64_t ad1 = c.a_mn ? ptx::make_smem_desc(a1, 8192, 1024, 128) : ptx::make_smem_desc(a1, 16, 1024, 128);
    const uint64_t bd = c.b_mn ? ptx::make_smem_desc(b0, 8192, 1024, 128) : ptx::make_smem_desc(b0, 16, 1024, 128);
    const uint32_t id1 = ptx::make_idesc_bf16(128, c.n1, c.a_mn, c.b_mn);
    const uint32_t id2 = ptx::make_idesc_bf16(128, c.n2 ? c.n2 : 64, c.a_mn, c.b_mn);
    const uint64_t kstep = c.a_mn ? 128 : 2;   // 16 K rows of 128 B (MN-major) or 32 B along the row (K-major)
    const uint64_t kstep_b = c.b_mn ? 128 : 2;
    const long long t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ptx::umma_bf16(tmem_base, ad0 + k * kstep, bd + k * kstep_b, id1, 1);
          if (c.n2) ptx::umma_bf16(tmem_base + c.d2_off, (c.same_a ? ad0 : ad1) + k * kstep, bd + k * kstep_b, id2, 1);
        }
      }
      ptx::umma_commit(&done);
    }
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    const long long t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem_base, 512);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const Cfg cfgs[] = {
      {64, 0, 0, 0, 1, 0},  {128, 0, 0, 0, 1, 0},  {256, 0, 0, 0, 1, 0},
      {128, 64, 0, 0, 0, 256}, {256, 128, 0, 0, 0, 256}, {128, 128, 0, 0, 0, 256}, {64, 64, 0, 0, 0, 256},
      {128, 128, 0, 0, 0, 128}, {64, 64, 0, 0, 0, 128}, {64, 64, 0, 0, 0, 64}, {128, 64, 0, 0, 0, 128},
      {128, 128, 0, 0, 0, 0}, {64, 64, 0, 0, 0, 0}, {256, 128, 0, 0, 0, 0},
      {128, 64, 0, 0, 0, 32}, {128, 64, 0, 0, 0, 64}, {256, 128, 0, 0, 0, 64}, {128, 0, 0, 0, 1, 0},
      {192, 192, 1, 1, 0, 256}, {192, 192, 1, 1, 0, 0}, {192, 0, 1, 1, 1, 0},
  };
  const int iters = 512;
  std::vector<long long> h(sms);
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; ++rep) {
      rate_kernel<<<sms, 128, 97 * 1024>>>(c, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long v : h) avg += (double)v;
    avg /= sms;
    const int per_iter = 4 * (c.n2 ? 2 : 1);
    const double clk = avg / ((double)iters * per_iter);
    const double math = c.n2 ? (c.n1 + c.n2) / 4.0 : c.n1 / 2.0;
    printf("MMARATE N=%3d+%3d A:%s B:%s second accumulator at +%3d cols : %.1f clk per MMA (math floor %.1f)\n", c.n1, c.n2,
           c.a_mn ? "MN" : "K ", c.b_mn ? "MN" : "K ", c.d2_off, clk, math);
  }
  return 0;
}
