// egaze-b200: shared device/host helpers for the sm_100a kernels.
// Everything here is hand-written inline PTX for Blackwell (mbarrier, TMA, tcgen05/TMEM);
// no CUTLASS/CuTe, no Triton.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

// ---------------------------------------------------------------------------------------------
// Error convention of the C-ABI (include/egaze.h): 0 ok, <0 invalid argument, >0 cudaError_t.
// ---------------------------------------------------------------------------------------------
#define EGAZE_OK 0
#define EGAZE_EINVAL (-1)
#define EGAZE_EUNSUPPORTED (-2)
#define EGAZE_EDRIVER (-3)

void egaze_set_error(const char* fmt, ...);

#define EGAZE_CHECK_ARG(cond, ...)                                   \
  do {                                                               \
    if (!(cond)) {                                                   \
      egaze_set_error(__VA_ARGS__);                                  \
      return EGAZE_EINVAL;                                           \
    }                                                                \
  } while (0)

#define EGAZE_CUDA(call)                                                                  \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      egaze_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return (int)e__;                                                                    \
    }                                                                                     \
  } while (0)

#define EGAZE_LAUNCH_CHECK()                                                              \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess) {                                                             \
      egaze_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      return (int)e__;                                                                    \
    }                                                                                     \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// One-time per-DEVICE setup (cudaFuncSetAttribute is per device and one process may drive several GPUs, e.g. the reference's
// `--device N` without torch.cuda.set_device): true the first time it is called on the current device for this mask.
static inline bool egaze_first_on_device(unsigned long long* mask) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return true;
  const unsigned long long bit = 1ull << (dev & 63);
  if (*mask & bit) return false;
  *mask |= bit;
  return true;
}

// ---------------------------------------------------------------------------------------------
// split-bf16 helpers: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}
__device__ __forceinline__ float bf16_bits_to_float(uint32_t bits16) { return __uint_as_float(bits16 << 16); }
// Four values at once with the packed converter (cvt.rn.bf16x2.f32 -> F2FP, ALU pipe; the scalar F2F goes through the
// quarter-rate XU pipe).  Bit-identical to four split_bf16 calls; hi2/lo2 are ready for an 8-byte NHWC store.
__device__ __forceinline__ void split_bf16x4(const float4& v, uint2& hi2, uint2& lo2) {
  const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
  const uint32_t u01 = *reinterpret_cast<const uint32_t*>(&h01), u23 = *reinterpret_cast<const uint32_t*>(&h23);
  const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __uint_as_float(u01 << 16), v.y - __uint_as_float(u01 & 0xffff0000u));
  const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __uint_as_float(u23 << 16), v.w - __uint_as_float(u23 & 0xffff0000u));
  hi2 = make_uint2(u01, u23);
  lo2 = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}

// ---------------------------------------------------------------------------------------------
// split-fp16 helpers: x ~= hi + lo with hi = fp16(x), lo = fp16(x - hi)  (22 significand bits while |x| >= 2^-3, an absolute
// error <= 2^-25 below: fp16 subnormals are exact multiples of 2^-24).  |x| is clamped to the fp16 range (65504) first, so a
// wild activation saturates instead of turning into inf / NaN.  Used by the FORWARD convolutions only (train-mode
// BatchNorm needs >= 19 operand bits for the 1e-3 gate); gradients keep bf16, whose exponent range they need.
// ---------------------------------------------------------------------------------------------
constexpr float kF16Max = 65504.f;
__device__ __forceinline__ void split_f16x4(const float4& v, uint2& hi2, uint2& lo2) {
  const float x = fminf(fmaxf(v.x, -kF16Max), kF16Max), y = fminf(fmaxf(v.y, -kF16Max), kF16Max);
  const float z = fminf(fmaxf(v.z, -kF16Max), kF16Max), w = fminf(fmaxf(v.w, -kF16Max), kF16Max);
  const __half2 h01 = __floats2half2_rn(x, y), h23 = __floats2half2_rn(z, w);
  const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
  const __half2 l01 = __floats2half2_rn(x - f01.x, y - f01.y), l23 = __floats2half2_rn(z - f23.x, w - f23.y);
  hi2 = make_uint2(*reinterpret_cast<const uint32_t*>(&h01), *reinterpret_cast<const uint32_t*>(&h23));
  lo2 = make_uint2(*reinterpret_cast<const uint32_t*>(&l01), *reinterpret_cast<const uint32_t*>(&l23));
}
__device__ __forceinline__ void split_f16(float x, __half& hi, __half& lo) {
  x = fminf(fmaxf(x, -kF16Max), kF16Max);
  hi = __float2half_rn(x);
  lo = __float2half_rn(x - __half2float(hi));
}
__device__ __forceinline__ float f16_bits_to_float(uint32_t bits16) {
  return __half2float(__ushort_as_half((unsigned short)bits16));
}
// four fp32 -> four bf16 (round to nearest), packed for an 8-byte store
__device__ __forceinline__ uint2 pack_bf16x4(const float4& v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}
// value > 0 test on the raw 16 bits of a bf16 OR fp16 number (sign clear, not zero; NaN counts as positive in both formats,
// which the ReLU masks never hold)
// Sub-pixel decomposition of nearest-2x upsample -> 3x3 conv (conv3x3_tc.cu, ConvTcParams::sub).  Output phase py sees the
// low-resolution rows through two taps a: py = 0: {w[0]} at row -1, {w[1] + w[2]} at row 0;  py = 1: {w[0] + w[1]} at row 0,
// {w[2]} at row +1 -- the same along x.  w9: the 3x3 taps [r*3+s]; q16: [(py*2+px)*4 + a*2 + b], summed in fp32.
__device__ __forceinline__ void egaze_subpixel_taps(const float* w9, float* q16) {
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          // rows / columns of the 3x3 kernel folded into tap (a, b) of phase (py, px): [lo, hi]
          const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
          const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
              if (r >= r0 && r <= r1 && s >= s0 && s <= s1) acc += w9[r * 3 + s];
          q16[(py * 2 + px) * 4 + a * 2 + b] = acc;
        }
}
// transpose of the map above: the gradient of 3x3 tap (r, s) is the sum of the sub-pixel planes it was folded into
__device__ __forceinline__ void egaze_subpixel_taps_transpose(const float* q16, float* g9) {
#pragma unroll
  for (int t = 0; t < 9; ++t) g9[t] = 0.f;
#pragma unroll
  for (int py = 0; py < 2; ++py)
#pragma unroll
    for (int px = 0; px < 2; ++px)
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int r0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), r1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
          const int s0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), s1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
          const float q = q16[(py * 2 + px) * 4 + a * 2 + b];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s)
              if (r >= r0 && r <= r1 && s >= s0 && s <= s1) g9[r * 3 + s] += q;
        }
}

__device__ __forceinline__ bool pos16(uint32_t bits16) { return bits16 != 0u && bits16 < 0x8000u; }
// decode element `fmt` (0 = bf16, 1 = fp16) from its 16 bits
__device__ __forceinline__ float dec16(uint32_t bits16, int fmt) { return fmt ? f16_bits_to_float(bits16) : bf16_bits_to_float(bits16); }

// ---------------------------------------------------------------------------------------------
// warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Blackwell PTX wrappers
// ---------------------------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space accesses by 32-bit shared address (a pointer carved out of the dynamic smem window by integer
// arithmetic is "generic" to the compiler, which then emits the slower LD/ST instead of LDS/STS)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float lds32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier --------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Non-blocking probe of a phase (no suspend hint): issue it early, consume the predicate later -- the ~90-cycle latency of
// a barrier query then overlaps whatever is issued in between.
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must surface as a trapped launch (cudaErrorLaunchFailure), never as a hung GPU (~2 s of
// try_wait retries at most).  The retry loop with its clock reads and the diagnostic printf is kept OUT of line: inlined at
// every wait site it put ~25 instructions of cold code into loops whose instruction-cache footprint is what bounds them
// (conv3x3_tc.cu, MMA warp).
static __device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) {
      printf("egaze: mbarrier wait timeout (block %d,%d thread %d bar@%u parity %u)\n", blockIdx.x, blockIdx.y, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (!mbar_try_wait(bar, parity)) mbar_wait_slow(smem_u32(bar), parity);
}

// ---- TMA (cp.async.bulk.tensor, tiled mode) --------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast: the box lands at the same smem offset (and signals the same-offset mbarrier) in every CTA of cta_mask
__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                               uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3)
      : "memory");
}

// ---- CTA-pair (cta_group::2) variants ------------------------------------------------------------
// In a 2-CTA MMA both CTAs of the cluster load their operand halves into their OWN shared memory, but the bytes are
// accounted on the LEADER's (cluster rank 0) mbarrier: a shared::cta address with bit 24 cleared names the same offset
// in rank 0 when used as a shared::cluster address.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                                int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
        "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the same-offset mbarrier of cluster rank 0 (from any CTA of the cluster)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// The same with .relaxed semantics: for arrivals that publish no memory -- an epilogue warp telling the leader's MMA thread that
// its tcgen05.ld reads of an accumulator stage have completed (tcgen05.wait::ld + tcgen05.fence::before_thread_sync precede it).
// The .release form compiles to MEMBAR.ALL.GPU + ERRBAR, which waits for every earlier global store of the warp (the previous
// tile's output) to be performed: ~8 % of the epilogue time of the rank-1 CTAs (ncu source page, round 2).
__device__ __forceinline__ void mbar_arrive_leader_relaxed(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[each CTA's own 128 rows] * B[N/2 rows from each CTA]; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on the same-offset mbarrier of every CTA in cta_mask once all prior 2-CTA MMAs of this thread have completed.
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- tcgen05 / TMEM ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem], bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// Same, arriving on the same-offset mbarrier of every CTA in cta_mask (cluster launch).
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), swizzled canonical layouts.
//   swizzle_bytes in {128, 64, 32}; lbo/sbo in bytes.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t swizzle_bytes) {
  uint64_t layout = swizzle_bytes == 128 ? 2ull : (swizzle_bytes == 64 ? 4ull : (swizzle_bytes == 32 ? 6ull : 0ull));
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;  // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}
// Instruction descriptor for kind::f16, BF16 x BF16 -> FP32.
//   a_mn / b_mn: 0 = K-major operand, 1 = MN-major operand.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  return (1u << 4)              // C format F32
         | (1u << 7)            // A format BF16
         | (1u << 10)           // B format BF16
         | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// Same with FP16 operands (A / B format field 0); kind::f16 wants both operands in the same 16-bit format.
__host__ __device__ constexpr uint32_t make_idesc_16(int M, int N, int a_mn, int b_mn, int f16) {
  return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace ptx

// ---------------------------------------------------------------------------------------------
// Host: tensor-map encoding (driver entry point fetched at run time; no -lcuda link dependency)
// ---------------------------------------------------------------------------------------------
// dims/strides innermost-first; strides_bytes[i] is the byte stride of dim i+1 (dim 0 is dense).
int egaze_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int elem_bytes);
