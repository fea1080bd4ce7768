// Weight gradient of the 3x3 / pad 1 conv as a tcgen05 GEMM with the PIXEL axis as K (sm_100a).
//
//   dW[tap=(r,s)][co][ci] = sum_{n,h,w} dY[n,h,w,co] * X[n,h+r-1,w+s-1,ci]
//
// (autograd of nn.Conv2d on the reference hot path: loss.backward() in SP.py:136, spatialstream.py:140.)
//
// GEMM view per CTA: M = 128 output channels, N = 3 vertical taps x 64 input channels = 192, K = pixels of the spatial
// tiles this CTA owns (split-K over tiles).  The three taps read the SAME dY tile, so their X operands (the window at row
// offsets 0, BW, 2*BW) are concatenated along N in one MMA: 64-channel chunks LBO = BW*128 bytes apart.  Both operands are
// "MN-major" for the tensor core: dY^T tile [K px][64 co] and X window [K px][64 ci] are exactly what TMA delivers
// from NHWC (pixel rows of 128 B, SWIZZLE_128B), so no transposes are materialised.  Like the forward kernel, the
// X window for horizontal tap s is fetched once with a (BH+2)-row halo and the three vertical taps read it at a row
// offset of r*BW rows.  precise mode: dYhi*Xhi + dYhi*Xlo + dYlo*Xhi as three N = 192 MMAs per K step, alternating between
// two accumulator blocks D1 / D2 (back-to-back MMAs into the same TMEM columns serialise, tools/probe_mma_rate.cu); the
// epilogue adds D1 + D2.  Cout == 64: the M rows are [dY_hi ; dY_lo] of the same 64 channels and two MMAs do all four
// products.
//
// Grid: x = split-K slice, y = (co-tile, ci-tile, s).  Epilogue: TMEM -> registers -> vector fp32 atomics into the
// packed [9][Cout][Cin_p] accumulator (zeroed by the caller), which egaze_unpack_wgrad turns into the OIHW .grad.
#include "common.cuh"

namespace {

constexpr int kThreads = 192;

struct WgradParams {
  int N, H, W, Cin_p, Cout;
  int BH, BW;
  int tiles_h, tiles_w, total_tiles;
  int co_tiles, ci_tiles;
  int m_chunks;        // 64-channel chunks of dY actually present (1 when Cout == 64, else 2)
  int stacked;         // Cout == 64, precise: the M = 128 rows are [dY_hi ; dY_lo] of the same 64 channels, so ONE MMA against
                       // [X_hi | X_lo] yields all four hi/lo products (the epilogue adds the two lane halves into the same dW rows)
  int nsplit;
  int pairs;           // Cout == 64, one plane: the M = 128 rows are [dY ; dY shifted by one pixel column], so ONE job yields two horizontal
                       // taps (s and s + 1) against the same X window -- two jobs per (co, ci) tile instead of three (sub-pixel: one per phase
                       // instead of two), dY and X are streamed a third (half) less often.  The paired job walks one extra tile column:
                       // the shifted rows of tile column t cover pixels [w0 - 1, w0 + BW - 1), so pixel W - 1 needs the column at w0 = W.
  int ci2;             // one plane, Cin_p >= 128: a job covers TWO 64-channel input tiles (two X windows per stage, two N = 192 MMAs per
                       // K step into accumulator blocks 0 / 256): the dY tile is streamed half as often, the kernel's L2 -> SM traffic per
                       // FLOP drops by a third
  int sub;             // sub-pixel form (conv3x3_tc.cu ConvTcParams::sub): X is the low-resolution input, dY the phase-planar gradient
                       // [4N][H][W][Cout]; job = (co-tile, ci-tile, phase, horizontal tap b), the TWO vertical taps a share the dY tile
                       // (N = 128); dwp has 16 planes [phase*4 + a*2 + b]
  int stage_bytes, dy_plane_bytes, x_plane_bytes;
  float* dwp;          // [9][Cout][Cin_p]
};

template <int NSPLIT>
__global__ void __launch_bounds__(kThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY_hi, const __grid_constant__ CUtensorMap tmY_lo,
                const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  int job = blockIdx.y;
  const int nj = p.pairs ? (p.sub ? 4 : 2) : (p.sub ? 8 : 3);
  int s = job % nj;                // horizontal tap; sub-pixel: phase * 2 + b
  job /= nj;
  bool paired = false;             // this job's rows 64..127 hold dY shifted by one pixel column: tap s + 1 as well
  if (p.pairs) {
    if (p.sub) { s = s * 2; paired = true; }        // one job per phase: b = 0 and 1
    else { paired = s == 0; s = s == 0 ? 0 : 2; }   // taps {0, 1} and {2}
  }
  const int tiles_w = paired ? (p.W + 1 + p.BW - 1) / p.BW : p.tiles_w;
  const int total_tiles = p.N * p.tiles_h * tiles_w;
  const int phase = p.sub ? s >> 1 : 0, py = phase >> 1, px = phase & 1, hb = p.sub ? s & 1 : s;
  const int ntap_v = p.sub ? 2 : 3;
  const int ci_jobs = p.ci2 ? p.ci_tiles / 2 : p.ci_tiles;
  const int ci_t = job % ci_jobs;
  const int co_t = job / ci_jobs;
  const int co0 = co_t * 128, ci0 = ci_t * (p.ci2 ? 128 : 64);

  __shared__ uint64_t full[2], empty[2], acc_full;
  __shared__ uint32_t tmem_base_smem;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmY_hi);
    ptx::prefetch_tmap(&tmX_hi);
  }
  // precise: A_hi x [X_hi | X_lo] is ONE MMA of N = 128 (the X planes are 64-channel chunks LBO = plane stride apart),
  // so each tap owns 128 accumulator columns whose halves are added in the epilogue; fast: 64 columns per tap.
  constexpr uint32_t kD2 = 256;                       // column offset of the second accumulator block (precise)
  constexpr uint32_t kTmemCols = 512;   // (one plane: the second block holds the second input-channel tile, ci2)
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_smem, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int KP = p.BH * p.BW;                       // pixels (K) per tile, multiple of 16
  const uint32_t dy_box_bytes = (uint32_t)KP * 128u;
  const uint32_t x_box_bytes = (uint32_t)(p.BH + 2) * p.BW * 128u;
  const uint32_t tx_bytes = NSPLIT * ((p.m_chunks + (paired ? 1 : 0)) * dy_box_bytes + (p.ci2 ? 2 : 1) * x_box_bytes);

  if (warp == 0) {
    if (lane == 0) {
      int st = 0;
      uint32_t par = 1;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tw_i = t % tiles_w;
        const int th_i = (t / tiles_w) % p.tiles_h;
        const int img = t / (tiles_w * p.tiles_h);
        const int h0 = th_i * p.BH, w0 = tw_i * p.BW;
        ptx::mbar_wait(&empty[st], par);
        ptx::mbar_arrive_expect_tx(&full[st], tx_bytes);
        uint8_t* base = smem + (size_t)st * p.stage_bytes;
        const int img_y = phase * p.N + img;   // sub-pixel: the gradient of this phase's output pixels is one plane of 4N images
        for (int c = 0; c < p.m_chunks; ++c) {
          ptx::tma_load_4d(base + c * dy_box_bytes, &tmY_hi, &full[st], co0 + 64 * c, w0, h0, img_y);
          if (NSPLIT == 2)
            ptx::tma_load_4d(base + (p.stacked ? dy_box_bytes : (uint32_t)p.dy_plane_bytes + c * dy_box_bytes), &tmY_lo, &full[st],
                             co0 + 64 * c, w0, h0, img_y);
        }
        // paired job: the second 64-row chunk of A is the same dY tile one pixel column to the left (out-of-image columns are zeros)
        if (paired) ptx::tma_load_4d(base + dy_box_bytes, &tmY_hi, &full[st], co0, w0 - 1, h0, img_y);
        uint8_t* xb = base + (size_t)NSPLIT * p.dy_plane_bytes;
        // window origin: 3x3 tap (r, s) reads X[h + r - 1, w + s - 1]; sub-pixel tap (a, b) of phase (py, px) reads
        // X[i + py + a - 1, j + px + b - 1]
        const int xw = w0 - 1 + px + hb, xh = h0 - 1 + py;
        ptx::tma_load_4d(xb, &tmX_hi, &full[st], ci0, xw, xh, img);
        if (NSPLIT == 2) ptx::tma_load_4d(xb + p.x_plane_bytes, &tmX_lo, &full[st], ci0, xw, xh, img);
        if (NSPLIT == 1 && p.ci2) ptx::tma_load_4d(xb + p.x_plane_bytes, &tmX_hi, &full[st], ci0 + 64, xw, xh, img);
        if (++st == 2) { st = 0; par ^= 1; }
      }
    }
  } else if (warp == 1) {
    // The whole warp walks the pipeline so that the descriptors stay in uniform registers; one elected lane issues
    // (see conv3x3_tc.cu: a loop nest under `if (lane == 0)` costs ~100 cycles of R2UR traffic per tcgen05.mma).
    {
      const bool leader = ptx::elect_one();
      const uint32_t idesc = ptx::make_idesc_bf16(128, 64 * ntap_v, 1, 1);   // both operands MN-major, N = vertical taps x 64 channels
      // MN-major SWIZZLE_128B canonical layout: 64 channels (128 B) contiguous, 8 pixel rows per 1024 B atom (SBO),
      // next 64-channel chunk LBO bytes away: the other dY chunk for A, the window shifted by one tile row for B.
      const uint64_t a_static = ptx::make_smem_desc(0, dy_box_bytes, 1024, 128);
      const uint64_t b_static = ptx::make_smem_desc(0, (uint32_t)p.BW * 128u, 1024, 128);
      const int ksteps = KP / 16;
      int st = 0;
      uint32_t par = 0;
      uint32_t acc = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        ptx::mbar_wait(&full[st], par);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem + (size_t)st * p.stage_bytes);
        const uint64_t a_hi = a_static + (uint64_t)(base >> 4);
        const uint64_t a_lo = a_hi + (uint64_t)((uint32_t)p.dy_plane_bytes >> 4);
        const uint64_t b_hi = b_static + (uint64_t)((base + (uint32_t)NSPLIT * p.dy_plane_bytes) >> 4);
        const uint64_t b_lo = b_hi + (uint64_t)((uint32_t)p.x_plane_bytes >> 4);
        if (leader) {
#pragma unroll 2
          for (int k = 0; k < ksteps; ++k) {  // 16 pixel rows = 2048 B per K step
            const uint64_t ko = (uint64_t)(k * 128);
            const uint32_t first = acc | (uint32_t)k;   // 0 only for the very first K step of this CTA
            if (NSPLIT == 1) {
              ptx::umma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, first);
              if (p.ci2) ptx::umma_bf16(tmem_base + kD2, a_hi + ko, b_lo + ko, idesc, first);   // second input-channel tile -> block 2
            } else if (p.stacked) {
              ptx::umma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, first);         // [dY_hi ; dY_lo] x X_hi -> D1
              ptx::umma_bf16(tmem_base + kD2, a_hi + ko, b_lo + ko, idesc, first);   // [dY_hi ; dY_lo] x X_lo -> D2
            } else if ((k & 1) == 0) {
              ptx::umma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, first);         // hi*hi -> D1
              ptx::umma_bf16(tmem_base + kD2, a_hi + ko, b_lo + ko, idesc, first);   // hi*lo -> D2
              ptx::umma_bf16(tmem_base, a_lo + ko, b_hi + ko, idesc, 1);             // lo*hi -> D1
            } else {
              ptx::umma_bf16(tmem_base + kD2, a_hi + ko, b_lo + ko, idesc, 1);       // hi*lo -> D2
              ptx::umma_bf16(tmem_base, a_hi + ko, b_hi + ko, idesc, 1);             // hi*hi -> D1
              ptx::umma_bf16(tmem_base + kD2, a_lo + ko, b_hi + ko, idesc, 1);       // lo*hi -> D2
            }
          }
          ptx::umma_commit(&empty[st]);
        }
        acc = 1;
        __syncwarp();
        if (++st == 2) { st = 0; par ^= 1; }
      }
      if (leader) ptx::umma_commit(&acc_full);
    }
  } else {
    // epilogue: thread = output-channel row; 64 consecutive ci per tap -> float4 atomics
    const int ew = warp & 3;
    const int m = ew * 32 + lane;
    ptx::mbar_wait(&acc_full, 0);
    ptx::tc_fence_after();
    const bool any_tile = (int)blockIdx.x < total_tiles;
    for (int hh = 0; hh < ((NSPLIT == 1 && p.ci2) ? 2 : 1); ++hh)
    for (int r = 0; r < ntap_v; ++r) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(hh * kD2) + (uint32_t)(r * 64 + c0), v);
        if (NSPLIT == 2) {
          uint32_t v2[32];
          ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + kD2 + (uint32_t)(r * 64 + c0), v2);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
        } else {
          ptx::tmem_ld_wait();
        }
        // stacked: lanes 64..127 hold the dY_lo products of channels 0..63; pairs: those of the next horizontal tap
        const int co = (p.stacked || p.pairs) ? (m & 63) : co0 + m;
        const int hshift = (p.pairs && m >= 64) ? 1 : 0;
        if (any_tile && co < p.Cout && !(p.pairs && m >= 64 && !paired)) {
          const int plane = p.sub ? phase * 4 + r * 2 + hb + hshift : r * 3 + s + hshift;
          float* dst = p.dwp + ((size_t)plane * p.Cout + co) * p.Cin_p + ci0 + hh * 64 + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 val = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                     __uint_as_float(v[j + 3]));
            atomicAdd(reinterpret_cast<float4*>(dst + j), val);
          }
        }
      }
    }
    ptx::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, kTmemCols);
}

void pick_wgrad_tile(int H, int W, int* BH, int* BW) {
  // K per tile = BH*BW must be a multiple of 16 and <= 128; BW % 8 == 0
  const int cand[][2] = {{8, 16}, {4, 32}, {16, 8}, {14, 8}, {6, 16}, {12, 8}, {2, 64}, {10, 8}, {4, 16}, {8, 8}, {2, 32}, {4, 8}, {2, 8}};
  double best = -1.0;
  for (auto& c : cand) {
    const int bh = c[0], bw = c[1];
    if ((bh * bw) % 16) continue;
    const double eff = (double)H * W / ((double)ceil_div(H, bh) * ceil_div(W, bw) * bh * bw);  // useful K fraction
    const double halo = (double)(bh + 2) / bh;
    const double score = eff / (0.6 + 0.4 * halo) * (0.75 + 0.25 * (bh * bw) / 128.0);
    if (score > best) { best = score; *BH = bh; *BW = bw; }
  }
}

}  // namespace

// dwp ([9][Cout][Cin_p] fp32; sub: [16][Cout][Cin_p]) is ACCUMULATED into: zero it first.  Cin_p % 64 == 0, Cout % 64 == 0.
extern "C" int egaze_wgrad3x3_tc(const void* x_hi, const void* x_lo, const void* dy_hi, const void* dy_lo, int N, int H,
                                 int W, int Cin_p, int Cout, float* dwp, int precise, int sub, void* stream) {
  EGAZE_CHECK_ARG(x_hi && dy_hi && dwp, "wgrad3x3_tc: null operand");
  EGAZE_CHECK_ARG(!precise || (x_lo && dy_lo), "wgrad3x3_tc: precise mode needs lo planes");
  EGAZE_CHECK_ARG(Cin_p % 64 == 0 && Cout % 64 == 0, "wgrad3x3_tc: Cin_p=%d, Cout=%d must be multiples of 64", Cin_p, Cout);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin_p = Cin_p; p.Cout = Cout;
  pick_wgrad_tile(H, W, &p.BH, &p.BW);
  p.tiles_h = ceil_div(H, p.BH); p.tiles_w = ceil_div(W, p.BW);
  p.total_tiles = N * p.tiles_h * p.tiles_w;
  p.co_tiles = ceil_div(Cout, 128); p.ci_tiles = Cin_p / 64;
  p.m_chunks = Cout >= 128 ? 2 : 1;
  p.nsplit = precise ? 2 : 1;
  p.sub = sub ? 1 : 0;
  p.stacked = (precise && Cout == 64) ? 1 : 0;
  {
    static int pairs_env = -1;
    if (pairs_env < 0) {
      const char* e = getenv("EGAZE_WGRAD_PAIRS");
      pairs_env = e ? atoi(e) : 1;
    }
    p.pairs = (pairs_env && !precise && Cout == 64) ? 1 : 0;
  }
  {
    const char* e = getenv("EGAZE_WGRAD_STACKED");
    if (e && atoi(e) == 0) p.stacked = 0;
  }
  {
    static int ci2_env = -1;
    if (ci2_env < 0) {
      const char* e = getenv("EGAZE_WGRAD_CI2");
      ci2_env = e ? atoi(e) : 1;
    }
    p.ci2 = (ci2_env && !precise && p.ci_tiles % 2 == 0) ? 1 : 0;
  }
  const int KP = p.BH * p.BW;
  p.dy_plane_bytes = 2 * KP * 128;                                  // room for both 64-channel chunks
  p.x_plane_bytes = ((p.BH + 2) * p.BW * 128 + 1023) / 1024 * 1024;
  // the MMA for tap r reads K rows [r*BW, r*BW + KP): always inside the (BH+2)*BW window
  p.stage_bytes = p.nsplit * (p.dy_plane_bytes + p.x_plane_bytes) + (p.ci2 ? p.x_plane_bytes : 0);
  p.dwp = dwp;
  const size_t smem = (size_t)2 * p.stage_bytes + 1024;
  EGAZE_CHECK_ARG(smem <= 224 * 1024, "wgrad3x3_tc: tile does not fit shared memory");

  CUtensorMap tmY_hi, tmY_lo, tmX_hi, tmX_lo;
  {
    uint64_t dims[4] = {(uint64_t)Cout, (uint64_t)W, (uint64_t)H, (uint64_t)(sub ? 4 * N : N)};
    uint64_t str[3] = {(uint64_t)Cout * 2, (uint64_t)W * Cout * 2, (uint64_t)H * W * Cout * 2};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)p.BH, 1};
    int rc = egaze_encode_tmap(&tmY_hi, dy_hi, 4, dims, str, box, 128, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmY_lo, precise ? dy_lo : dy_hi, 4, dims, str, box, 128, 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[4] = {(uint64_t)Cin_p, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin_p * 2, (uint64_t)W * Cin_p * 2, (uint64_t)H * W * Cin_p * 2};
    uint32_t box[4] = {64, (uint32_t)p.BW, (uint32_t)(p.BH + 2), 1};
    int rc = egaze_encode_tmap(&tmX_hi, x_hi, 4, dims, str, box, 128, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmX_lo, precise ? x_lo : x_hi, 4, dims, str, box, 128, 2);
    if (rc) return rc;
  }
  // split-K factor: one CTA per SM is resident (208 KB of smem), so pick the factor that fills whole waves of SMs best
  const int jobs = p.co_tiles * (p.ci2 ? p.ci_tiles / 2 : p.ci_tiles) * (p.pairs ? (sub ? 4 : 2) : (sub ? 8 : 3));
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    EGAZE_CUDA(cudaGetDevice(&dev));
    EGAZE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  int ksplit = 1;
  double best_util = -1.0;
  const int kmax = p.total_tiles < 4 * sms ? p.total_tiles : 4 * sms;
  for (int k = 1; k <= kmax; ++k) {
    const long long ctas = (long long)jobs * k;
    if (ctas > 8LL * sms) break;
    const long long waves = (ctas + sms - 1) / sms;
    const int tiles_per_cta = (p.total_tiles + k - 1) / k;
    // time ~ waves * tiles_per_cta (+ a fixed per-CTA prologue/epilogue worth ~6 tiles)
    const double cost = (double)waves * (tiles_per_cta + 6);
    const double util = 1.0 / cost;
    if (util > best_util * 1.0001) { best_util = util; ksplit = k; }
  }
  dim3 grid((unsigned)ksplit, (unsigned)jobs);
  if (precise) {
    static unsigned long long attr = 0;
    if (egaze_first_on_device(&attr)) {
      EGAZE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    }
    wgrad_tc_kernel<2><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmY_hi, tmY_lo, tmX_hi, tmX_lo, p);
  } else {
    static unsigned long long attr = 0;
    if (egaze_first_on_device(&attr)) {
      EGAZE_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    }
    wgrad_tc_kernel<1><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmY_hi, tmY_lo, tmX_hi, tmX_lo, p);
  }
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
