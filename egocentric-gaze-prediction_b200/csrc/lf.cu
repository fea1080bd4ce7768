// Late-fusion network (reference models/late_fusion.py:6-23): cat(f, g) -> [conv3x3 + BN + ReLU] x3 (2->32->32->8)
// -> conv1x1 -> sigmoid, forward and backward, as a dedicated kernel family (sm_100a).
//
// Why not the tcgen05 conv of the SP stack: K = 9*2 / 9*32 / 9*32 and N = 32 / 32 / 8 are far too small for a
// 128 x N x 64 tcgen05 tile pipeline (round 1 measured ~1.5 ms per launch there, 50x off the layer's HBM roofline).  These
// layers are HBM-bound (205 MB per 32-channel fp32 map at B = 32, 224^2), so the design goal is ONE read and ONE write of
// every activation map per pass, with everything else fused around a warp-level tensor-core (mma.sync m16n8k16, split
// bf16 = 3 MMAs per product, fp32 accumulate) implicit GEMM whose operands are built in shared memory by the CTA itself:
//
//   forward layer L   : window loader applies the PREVIOUS layer's BatchNorm affine + ReLU to the raw fp32 map while it
//                       stages the (16+2) x (16+2) pixel window (no normalised copy ever reaches HBM), the epilogue adds
//                       the bias, stores the raw conv output (NHWC fp32) and accumulates BatchNorm batch statistics
//                       (shifted sums, one (mean, M2, n) partial per persistent CTA -> egaze_bn_finalize).  Eval mode:
//                       same kernels with folded running statistics, layer 3's epilogue goes straight through
//                       BN + ReLU + conv1x1 + sigmoid.
//   backward layer L  : the loader rebuilds d(raw_L) = gamma*invstd*(gz - mean(gz) - xhat*mean(gz*xhat)) from raw_L, the
//                       incoming gradient and the two BatchNorm sums on the fly; the data-gradient kernel (flipped weights)
//                       stores g_{L-1} and, in the same epilogue, accumulates layer L-1's BatchNorm-backward sums (it reads
//                       raw_{L-1} for the ReLU mask anyway); the weight-gradient kernel is an mma.sync GEMM with the pixel
//                       axis as K (one warp per filter tap, accumulators live in registers across all tiles of the CTA),
//                       per-CTA partials reduced deterministically straight into the OIHW .grad layout.
//
// Weights are read from the nn.Parameter's OIHW fp32 master and split into bf16 hi/lo inside each kernel's prologue
// (9 x 32 x 32 elements): the LF path needs no packed copies, no pack cache and no channel re-padding.
#include "common.cuh"

extern "C" int egaze_bn_finalize(const float* partial, const float* cnt, int cnt_stride, int cnt_div, int T, int C, float eps,
                                 float momentum, const float* gamma, const float* beta, float* running_mean,
                                 float* running_var, float* mean_out, float* invstd_out, float* scale_out, float* shift_out,
                                 long long* num_batches_tracked, void* stream);
extern "C" int egaze_bn_fold(const float* gamma, const float* beta, const float* running_mean, const float* running_var,
                             const float* conv_bias, float eps, int C, float* scale, float* shift, void* stream);

namespace {

constexpr int TH = 16, TW = 16;          // output pixels per tile
constexpr int WH = TH + 2, WW = TW + 2;  // window with the 1-pixel halo of a 3x3 conv
constexpr int kConvThreads = 256;        // 8 warps x 2 tile rows
constexpr int kWgradThreads = 288;       // 9 warps = 9 filter taps
constexpr int kMaxCtas = 512;            // upper bound of any persistent grid here (sizes the partial buffers)

enum LoaderKind { L_INPUT = 0, L_ACT = 1, L_DRAW = 2, L_DRAW_HEAD = 3 };
enum EpiKind { EPI_RAW = 0, EPI_HEAD = 1, EPI_GRAD = 2, EPI_GX = 3 };

// ---- warp-level tensor-core primitives ---------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t addr, uint32_t& r0, uint32_t& r1) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// D (16x8 fp32) += A (16x16 bf16, row) * B (16x8 bf16, col)
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- window loaders ------------------------------------------------------------------------------------------------------
struct LdArgs {
  const float* raw;     // NHWC fp32 [N,H,W,C]: raw conv output of the layer this operand belongs to (L_ACT, L_DRAW, L_DRAW_HEAD)
  const float* g;       // NHWC fp32 [N,H,W,C]: gradient w.r.t. the layer's post-ReLU activation (L_DRAW)
  const float* f;       // [N,1,H,W] the two network inputs (L_INPUT)
  const float* gin;
  const float* gout;    // [N,H,W] gradient w.r.t. the network output (L_DRAW_HEAD)
  const float* out;     // [N,H,W] network output (sigmoid), (L_DRAW_HEAD)
  const float* wh;      // [8] weights of the 1x1 head (L_DRAW_HEAD)
  const float* scale;   // [C] BatchNorm affine of this layer: z = raw*scale + shift
  const float* shift;
  const float* mean;    // [C] batch statistics (L_DRAW*)
  const float* invstd;
  const float* dgamma;  // [C] sum gz*xhat, sum gz over the batch (L_DRAW*)
  const float* dbeta;
  float inv_n;          // 1 / (N*H*W)
  int batch_stats;      // 0: BatchNorm ran on running statistics (the mean / variance terms of its backward vanish)
};

// per-channel constants of a loader, staged once per CTA: 0 scale, 1 shift, 2 mean, 3 invstd, 4 kb, 5 kg, 6 head weight
template <int KIND, int C>
__device__ __forceinline__ void stage_consts(const LdArgs& a, float (*cst)[32]) {
  for (int c = threadIdx.x; c < 32; c += blockDim.x) {
    float sc = 1.f, sh = 0.f, mn = 0.f, iv = 1.f, kb = 0.f, kg = 0.f, wh = 0.f;
    if (KIND != L_INPUT && c < C) {
      sc = a.scale[c];
      sh = a.shift[c];
      if (KIND == L_DRAW || KIND == L_DRAW_HEAD) {
        mn = a.mean[c];
        iv = a.invstd[c];
        if (a.batch_stats) {
          kb = sc * a.dbeta[c] * a.inv_n;
          kg = sc * a.dgamma[c] * a.inv_n;
        }
        if (KIND == L_DRAW_HEAD) wh = a.wh[c];
      }
    }
    cst[0][c] = sc; cst[1][c] = sh; cst[2][c] = mn; cst[3][c] = iv; cst[4][c] = kb; cst[5][c] = kg; cst[6][c] = wh;
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// Stage a WWIN-wide window of wh_ rows whose top-left pixel is (oy, ox) of image img: per pixel CP bf16 channels (the real C
// first, zero padding behind) in two planes hi / lo, pixel pitch CP*2 + 16 bytes (the 16 spare bytes make the 8-row ldmatrix
// reads conflict-free).  Pixels outside the image are zeros (= the conv padding).
// Work item = (pixel, 8-channel group).  A thread first ISSUES the global loads of a whole batch of its items (BATCH x 2 or 4
// 16-byte loads in flight per thread: one DRAM round trip per batch instead of one per item -- the shared-memory stores below
// are asm volatile with a memory clobber, which the compiler never hoists loads across), then transforms and stores the batch.
template <int KIND, int C, int CP, int WWIN, int BATCH = ((KIND == L_DRAW) ? 3 : 6)>
__device__ __forceinline__ void load_window(const LdArgs& a, const float (*cst)[32], uint32_t hi_s, uint32_t lo_s, int img,
                                            int oy, int ox, int wh_, int H, int W) {
  constexpr int G = CP / 8;            // 16-byte channel groups per pixel
  constexpr int GR = (C + 7) / 8;      // groups that hold real channels
  constexpr int STRIDE = CP * 2 + 16;
  const int items = wh_ * WWIN * G;
  const int step = (int)blockDim.x;
  for (int i0 = threadIdx.x; i0 < items; i0 += BATCH * step) {
    float r[BATCH][8], gr[BATCH][8];
    bool live[BATCH];
    // ---- phase A: issue every load of the batch
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int i = i0 + k * step;
      const int p = i / G, cg = i - p * G;
      const int wy = p / WWIN, wx = p - wy * WWIN;
      const int gy = oy + wy, gx = ox + wx;
      live[k] = i < items && cg < GR && gy >= 0 && gy < H && gx >= 0 && gx < W;
#pragma unroll
      for (int e = 0; e < 8; ++e) { r[k][e] = 0.f; gr[k][e] = 0.f; }
      if (live[k]) {
        const size_t pix = ((size_t)img * H + gy) * W + gx;
        const int c0 = cg * 8;
        if (KIND == L_INPUT) {
          r[k][0] = __ldg(a.f + pix);
          r[k][1] = __ldg(a.gin + pix);
        } else {
          if (C >= 8) {
            const float4 r0 = ldg4(a.raw + pix * C + c0), r1 = ldg4(a.raw + pix * C + c0 + 4);
            r[k][0] = r0.x; r[k][1] = r0.y; r[k][2] = r0.z; r[k][3] = r0.w; r[k][4] = r1.x; r[k][5] = r1.y; r[k][6] = r1.z; r[k][7] = r1.w;
          }
          if (KIND == L_DRAW) {
            const float4 g0 = ldg4(a.g + pix * C + c0), g1 = ldg4(a.g + pix * C + c0 + 4);
            gr[k][0] = g0.x; gr[k][1] = g0.y; gr[k][2] = g0.z; gr[k][3] = g0.w; gr[k][4] = g1.x; gr[k][5] = g1.y; gr[k][6] = g1.z; gr[k][7] = g1.w;
          } else if (KIND == L_DRAW_HEAD) {
            gr[k][0] = __ldg(a.out + pix);     // y
            gr[k][1] = __ldg(a.gout + pix);    // d(loss)/dy
          }
        }
      }
    }
    // ---- phase B: transform + split + store
#pragma unroll
    for (int k = 0; k < BATCH; ++k) {
      const int i = i0 + k * step;
      if (i >= items) break;
      const int p = i / G, cg = i - p * G;
      const int c0 = cg * 8;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
      if (live[k]) {
        if (KIND == L_INPUT) {
          v[0] = r[k][0];
          v[1] = r[k][1];
        } else if (KIND == L_ACT) {
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(r[k][e], cst[0][c0 + e], cst[1][c0 + e]), 0.f);
        } else {
          float dz = 0.f;
          if (KIND == L_DRAW_HEAD) dz = gr[k][1] * gr[k][0] * (1.f - gr[k][0]);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float gin = KIND == L_DRAW ? gr[k][e] : dz * cst[6][c0 + e];
            const float sc = cst[0][c0 + e];
            const float z = fmaf(r[k][e], sc, cst[1][c0 + e]);
            const float gz = z > 0.f ? gin : 0.f;
            const float xh = (r[k][e] - cst[2][c0 + e]) * cst[3][c0 + e];
            v[e] = fmaf(sc, gz, -cst[4][c0 + e]) - xh * cst[5][c0 + e];
          }
        }
      }
      uint2 h0, l0, h1, l1;
      split_bf16x4(make_float4(v[0], v[1], v[2], v[3]), h0, l0);
      split_bf16x4(make_float4(v[4], v[5], v[6], v[7]), h1, l1);
      const uint32_t off = (uint32_t)(p * STRIDE + cg * 16);
      ptx::sts128(hi_s + off, h0.x, h0.y, h1.x, h1.y);
      ptx::sts128(lo_s + off, l0.x, l0.y, l1.x, l1.y);
    }
  }
}

// ---- conv / data-gradient kernel -----------------------------------------------------------------------------------------
struct ConvArgs {
  LdArgs ld;
  int N, H, W, tiles_y, tiles_x, ntiles;
  int precise;
  // weights: OIHW fp32 master.  transposed == 0: B[n][k] = w[n][k][tap]      (w is [NOUT][C][3][3])
  //                             transposed == 1: B[n][k] = w[k][n][8 - tap]  (w is [C][w_cin][3][3]: data gradient)
  const float* w;
  int transposed, w_cin;
  const float* bias;        // [NOUT] or null (EPI_RAW, EPI_HEAD)
  float* out;               // EPI_RAW: raw NHWC [N,H,W,NOUT]; EPI_GRAD: g_prev NHWC; EPI_HEAD: [N,H,W]
  float* stat_partial;      // EPI_RAW (train): [grid][2][NOUT] (mean, M2); EPI_GRAD: [grid][2][NOUT] (sum gz, sum gz*xhat)
  float* stat_cnt;          // EPI_RAW (train): [grid]
  // EPI_GRAD: the layer whose BatchNorm-backward sums ride on this epilogue (raw map + its forward constants)
  const float* e_raw;
  const float* e_scale; const float* e_shift; const float* e_mean; const float* e_invstd;
  // EPI_HEAD: folded BatchNorm 3 + 1x1 head
  const float* h_scale; const float* h_shift; const float* h_w; const float* h_b;
  float* gf; float* gg;     // EPI_GX: [N,1,H,W] gradients of the two inputs (either may be null)
};

template <int KIND, int C, int CP, int NOUT, int EPI>
__global__ void __launch_bounds__(kConvThreads, 2) lf_conv_kernel(const ConvArgs a) {
  constexpr int STRIDE = CP * 2 + 16;
  constexpr int NB = (NOUT + 7) / 8;   // n-blocks of 8 output channels
  constexpr int NROWS = NB * 8;
  constexpr int KS = CP / 16;
  constexpr int ACT_BYTES = WH * WW * STRIDE;
  constexpr int W_BYTES = 9 * NROWS * STRIDE;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float cst[7][32];
  __shared__ float ecst[5][32];        // EPI_RAW: 0 bias, 4 stats reference k | EPI_GRAD: 0 scale 1 shift 2 mean 3 invstd | EPI_HEAD: 0 bias 1 scale 2 shift 3 w
  __shared__ float red[8][2][32];
  const uint32_t act_hi = ptx::smem_u32(smem), act_lo = act_hi + ACT_BYTES;
  const uint32_t w_hi = act_lo + ACT_BYTES, w_lo = w_hi + W_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;

  stage_consts<KIND, C>(a.ld, cst);
  for (int c = threadIdx.x; c < 32; c += blockDim.x) {
    float e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
    if (c < NOUT) {
      if (EPI == EPI_RAW) e0 = a.bias ? a.bias[c] : 0.f;
      if (EPI == EPI_GRAD) { e0 = a.e_scale[c]; e1 = a.e_shift[c]; e2 = a.e_mean[c]; e3 = a.e_invstd[c]; }
      if (EPI == EPI_HEAD) { e0 = a.bias ? a.bias[c] : 0.f; e1 = a.h_scale[c]; e2 = a.h_shift[c]; e3 = a.h_w[c]; }
    }
    ecst[0][c] = e0; ecst[1][c] = e1; ecst[2][c] = e2; ecst[3][c] = e3; ecst[4][c] = 0.f;
  }
  // weights -> split bf16 [tap][n][k] (k contiguous), zero padded
  for (int i = threadIdx.x; i < 9 * NROWS * CP; i += blockDim.x) {
    const int k = i % CP, n = (i / CP) % NROWS, tap = i / (CP * NROWS);
    float v = 0.f;
    if (n < NOUT && k < C) v = a.transposed ? a.w[((size_t)k * a.w_cin + n) * 9 + (8 - tap)] : a.w[((size_t)n * C + k) * 9 + tap];
    __nv_bfloat16 h, l;
    split_bf16(v, h, l);
    const uint32_t off = (uint32_t)((tap * NROWS + n) * STRIDE + k * 2);
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(w_hi + off), "h"(__bfloat16_as_ushort(h)) : "memory");
    asm volatile("st.shared.b16 [%0], %1;" ::"r"(w_lo + off), "h"(__bfloat16_as_ushort(l)) : "memory");
  }
  __syncthreads();

  // lane-dependent parts of the ldmatrix addresses
  const uint32_t a_lane = (uint32_t)((lane & 15) * STRIDE + (lane >> 4) * 16);
  const uint32_t b_lane4 = (uint32_t)((((lane >> 4) * 8) + (lane & 7)) * STRIDE + ((lane >> 3) & 1) * 16);
  const uint32_t b_lane2 = (uint32_t)((lane & 7) * STRIDE + ((lane >> 3) & 1) * 16);
  const bool precise = a.precise != 0;

  float s1[NB][2], s2[NB][2];          // running BatchNorm sums of this thread's channels (EPI_RAW train / EPI_GRAD)
#pragma unroll
  for (int nb = 0; nb < NB; ++nb) { s1[nb][0] = s1[nb][1] = s2[nb][0] = s2[nb][1] = 0.f; }
  float npix = 0.f;
  bool have_k = false;
  const bool want_stats = a.stat_partial != nullptr;

  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    const int txi = t % a.tiles_x, tyi = (t / a.tiles_x) % a.tiles_y, img = t / (a.tiles_x * a.tiles_y);
    const int y0 = tyi * TH, x0 = txi * TW;
    load_window<KIND, C, CP, WW>(a.ld, cst, act_hi, act_lo, img, y0 - 1, x0 - 1, WH, a.H, a.W);
    __syncthreads();

    float acc[2][NB][4];
#pragma unroll
    for (int mb = 0; mb < 2; ++mb)
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[mb][nb][j] = 0.f;

#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap) {
      const int r = tap / 3, s = tap - 3 * r;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh[NB][2], bl[NB][2];
        const uint32_t wrow = (uint32_t)(tap * NROWS * STRIDE + ks * 32);
#pragma unroll
        for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
          uint32_t rr[4];
          ldsm_x4(w_hi + wrow + (uint32_t)(nb2 * 16 * STRIDE) + b_lane4, rr);
          bh[2 * nb2][0] = rr[0]; bh[2 * nb2][1] = rr[1]; bh[2 * nb2 + 1][0] = rr[2]; bh[2 * nb2 + 1][1] = rr[3];
          if (precise) {
            ldsm_x4(w_lo + wrow + (uint32_t)(nb2 * 16 * STRIDE) + b_lane4, rr);
            bl[2 * nb2][0] = rr[0]; bl[2 * nb2][1] = rr[1]; bl[2 * nb2 + 1][0] = rr[2]; bl[2 * nb2 + 1][1] = rr[3];
          }
        }
        if (NB & 1) {
          ldsm_x2(w_hi + wrow + (uint32_t)((NB - 1) * 8 * STRIDE) + b_lane2, bh[NB - 1][0], bh[NB - 1][1]);
          if (precise) ldsm_x2(w_lo + wrow + (uint32_t)((NB - 1) * 8 * STRIDE) + b_lane2, bl[NB - 1][0], bl[NB - 1][1]);
        }
#pragma unroll
        for (int mb = 0; mb < 2; ++mb) {
          const uint32_t arow = (uint32_t)(((2 * warp + mb + r) * WW + s) * STRIDE + ks * 32) + a_lane;
          uint32_t ah[4], al[4];
          ldsm_x4(act_hi + arow, ah);
          if (precise) ldsm_x4(act_lo + arow, al);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            mma_bf16(acc[mb][nb], ah, bh[nb][0], bh[nb][1]);
            if (precise) {
              mma_bf16(acc[mb][nb], ah, bl[nb][0], bl[nb][1]);
              mma_bf16(acc[mb][nb], al, bh[nb][0], bh[nb][1]);
            }
          }
        }
      }
    }

    // ---- epilogue: fragment (mb, nb): rows x = g, g + 8 of tile row y = 2*warp + mb; channels nb*8 + 2q, + 1
    if (EPI == EPI_RAW && want_stats && !have_k) {
      // reference values of the shifted sums: the CTA's first stored pixel (tile pixel (0,0) is always inside the image)
      if (warp == 0 && g == 0) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
          const int c = nb * 8 + 2 * q;
          if (c < NOUT) { ecst[4][c] = acc[0][nb][0] + ecst[0][c]; ecst[4][c + 1] = acc[0][nb][1] + ecst[0][c + 1]; }
        }
      }
      __syncthreads();
      have_k = true;
    }
    npix += (float)(min(TH, a.H - y0) * min(TW, a.W - x0));
#pragma unroll
    for (int mb = 0; mb < 2; ++mb) {
      const int gy = y0 + 2 * warp + mb;
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int gx = x0 + g + 8 * hf;
        const bool valid = gy < a.H && gx < a.W;
        const size_t pix = ((size_t)img * a.H + gy) * a.W + gx;
        if (EPI == EPI_HEAD) {
          float part = 0.f;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int c = nb * 8 + 2 * q;
            if (c < NOUT) {
              const float z0 = fmaf(acc[mb][nb][2 * hf] + ecst[0][c], ecst[1][c], ecst[2][c]);
              const float z1 = fmaf(acc[mb][nb][2 * hf + 1] + ecst[0][c + 1], ecst[1][c + 1], ecst[2][c + 1]);
              part = fmaf(fmaxf(z0, 0.f), ecst[3][c], part);
              part = fmaf(fmaxf(z1, 0.f), ecst[3][c + 1], part);
            }
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          if (q == 0 && valid) a.out[pix] = 1.f / (1.f + expf(-(part + a.h_b[0])));
        } else {
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const int c = nb * 8 + 2 * q;
            if (c >= NOUT || !valid) continue;
            float v0 = acc[mb][nb][2 * hf], v1 = acc[mb][nb][2 * hf + 1];
            if (EPI == EPI_RAW) {
              v0 += ecst[0][c]; v1 += ecst[0][c + 1];
              *reinterpret_cast<float2*>(a.out + pix * NOUT + c) = make_float2(v0, v1);
              if (want_stats) {
                const float d0 = v0 - ecst[4][c], d1 = v1 - ecst[4][c + 1];
                s1[nb][0] += d0; s1[nb][1] += d1;
                s2[nb][0] = fmaf(d0, d0, s2[nb][0]); s2[nb][1] = fmaf(d1, d1, s2[nb][1]);
              }
            } else if (EPI == EPI_GRAD) {
              *reinterpret_cast<float2*>(a.out + pix * NOUT + c) = make_float2(v0, v1);
              const float2 rw = __ldg(reinterpret_cast<const float2*>(a.e_raw + pix * NOUT + c));
              const float z0 = fmaf(rw.x, ecst[0][c], ecst[1][c]), z1 = fmaf(rw.y, ecst[0][c + 1], ecst[1][c + 1]);
              const float gz0 = z0 > 0.f ? v0 : 0.f, gz1 = z1 > 0.f ? v1 : 0.f;
              s1[nb][0] += gz0; s1[nb][1] += gz1;
              s2[nb][0] = fmaf(gz0, (rw.x - ecst[2][c]) * ecst[3][c], s2[nb][0]);
              s2[nb][1] = fmaf(gz1, (rw.y - ecst[2][c + 1]) * ecst[3][c + 1], s2[nb][1]);
            } else {  // EPI_GX: channel 0 -> gf, channel 1 -> gg
              if (c == 0) {
                if (a.gf) a.gf[pix] = v0;
                if (a.gg) a.gg[pix] = v1;
              }
            }
          }
        }
      }
    }
    __syncthreads();   // the window is free for the next tile
  }

  // ---- per-CTA partial of the BatchNorm sums
  if ((EPI == EPI_RAW || EPI == EPI_GRAD) && want_stats) {
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float x1 = s1[nb][j], x2 = s2[nb][j];
#pragma unroll
        for (int o = 4; o < 32; o <<= 1) {
          x1 += __shfl_xor_sync(0xffffffffu, x1, o);
          x2 += __shfl_xor_sync(0xffffffffu, x2, o);
        }
        if (g == 0 && nb * 8 + 2 * q + j < 32) { red[warp][0][nb * 8 + 2 * q + j] = x1; red[warp][1][nb * 8 + 2 * q + j] = x2; }
      }
    __syncthreads();
    if (threadIdx.x < NOUT) {
      const int c = threadIdx.x;
      float x1 = 0.f, x2 = 0.f;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) { x1 += red[wv][0][c]; x2 += red[wv][1][c]; }
      float* dst = a.stat_partial + (size_t)blockIdx.x * 2 * NOUT;
      if (EPI == EPI_RAW) {
        float mean = 0.f, m2 = 0.f;
        if (npix > 0.f) { mean = ecst[4][c] + x1 / npix; m2 = fmaxf(x2 - x1 * x1 / npix, 0.f); }
        dst[c] = mean; dst[NOUT + c] = m2;
        if (c == 0) a.stat_cnt[blockIdx.x] = npix;
      } else {
        dst[c] = x1; dst[NOUT + c] = x2;
      }
    }
  }
}

// ---- weight-gradient kernel -------------------------------------------------------------------------------------------------
// dW[tap][co][ci] = sum_pixels dY[p][co] * X[p + tap][ci]: M = co, N = ci, K = the 256 pixels of a tile (16 k-steps = 16 tile rows).
// Warp t owns filter tap t; its accumulators stay in registers across every tile of the CTA.
struct WgArgs {
  LdArgs ldy, ldx;
  int N, H, W, tiles_y, tiles_x, ntiles;
  int precise;
  float* partial;   // [grid][9][CY][CX]
};

template <int KY, int CY, int CPY, int KX, int CX, int CPX>
__global__ void __launch_bounds__(kWgradThreads, 2) lf_wgrad_kernel(const WgArgs a) {
  constexpr int SY = CPY * 2 + 16, SX = CPX * 2 + 16;
  constexpr int MB = CPY / 16, NB = CPX / 8;
  constexpr int Y_BYTES = TH * TW * SY, X_BYTES = WH * WW * SX;
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ float csty[7][32];
  __shared__ float cstx[7][32];
  const uint32_t y_hi = ptx::smem_u32(smem), y_lo = y_hi + Y_BYTES, x_hi = y_lo + Y_BYTES, x_lo = x_hi + X_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, q = lane & 3;
  stage_consts<KY, CY>(a.ldy, csty);
  stage_consts<KX, CX>(a.ldx, cstx);
  __syncthreads();
  const int r = warp / 3, s = warp - 3 * r;
  const bool precise = a.precise != 0;
  // ldmatrix.trans lane addressing: A = dY^T (matrices (m lo, k lo), (m hi, k lo), (m lo, k hi), (m hi, k hi)); B = X
  // (matrices (k lo, n lo), (k hi, n lo), (k lo, n hi), (k hi, n hi))
  const int mi = lane >> 3, lr = lane & 7;
  const uint32_t ya_lane = (uint32_t)(((mi >> 1) * 8 + lr) * SY + (mi & 1) * 16);
  const uint32_t xb_lane = (uint32_t)(((mi & 1) * 8 + lr) * SX + (mi >> 1) * 16);

  float acc[MB][NB][4];
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[mb][nb][j] = 0.f;

  for (int t = blockIdx.x; t < a.ntiles; t += gridDim.x) {
    const int txi = t % a.tiles_x, tyi = (t / a.tiles_x) % a.tiles_y, img = t / (a.tiles_x * a.tiles_y);
    const int y0 = tyi * TH, x0 = txi * TW;
    // (the tap accumulators stay live across the loaders: smaller batches than in the conv kernel, or they would spill)
    load_window<KY, CY, CPY, TW, (KY == L_DRAW ? 1 : 2)>(a.ldy, csty, y_hi, y_lo, img, y0, x0, TH, a.H, a.W);
    load_window<KX, CX, CPX, WW, 2>(a.ldx, cstx, x_hi, x_lo, img, y0 - 1, x0 - 1, WH, a.H, a.W);
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < TH; ++j) {   // k-step j = tile row j (16 pixels)
      uint32_t ah[MB][4], al[MB][4];
#pragma unroll
      for (int mb = 0; mb < MB; ++mb) {
        const uint32_t ad = (uint32_t)(j * TW * SY + mb * 32) + ya_lane;
        ldsm_x4_t(y_hi + ad, ah[mb]);
        if (precise) ldsm_x4_t(y_lo + ad, al[mb]);
      }
      const uint32_t xrow = (uint32_t)(((j + r) * WW + s) * SX) + xb_lane;
#pragma unroll
      for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
        uint32_t bh[4], bl[4];
        ldsm_x4_t(x_hi + xrow + (uint32_t)(nb2 * 32), bh);
        if (precise) ldsm_x4_t(x_lo + xrow + (uint32_t)(nb2 * 32), bl);
#pragma unroll
        for (int mb = 0; mb < MB; ++mb) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            mma_bf16(acc[mb][2 * nb2 + h], ah[mb], bh[2 * h], bh[2 * h + 1]);
            if (precise) {
              mma_bf16(acc[mb][2 * nb2 + h], ah[mb], bl[2 * h], bl[2 * h + 1]);
              mma_bf16(acc[mb][2 * nb2 + h], al[mb], bh[2 * h], bh[2 * h + 1]);
            }
          }
        }
      }
    }
    __syncthreads();
  }
  // partial[cta][tap][co][ci]
  float* dst = a.partial + ((size_t)blockIdx.x * 9 + warp) * CY * CX;
#pragma unroll
  for (int mb = 0; mb < MB; ++mb)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int co = mb * 16 + g + 8 * (j >> 1), ci = nb * 8 + 2 * q + (j & 1);
        if (co < CY && ci < CX) dst[co * CX + ci] = acc[mb][nb][j];
      }
}

// dw[co][ci][tap] (OIHW) = sum over CTAs of partial[cta][tap][co][ci]   (fp64 accumulation, fixed order: deterministic).
// One warp per output element group: lane l sums CTAs l, l+32, ... (coalesced across the 32 consecutive elements a block row
// covers), then a shuffle tree -- 296 dependent loads per thread became ~10.
__global__ void __launch_bounds__(256) lf_wgrad_reduce_kernel(const float* __restrict__ partial, int nctas, int CY, int CX,
                                                              float* __restrict__ dw) {
  __shared__ double red[8][32];
  const int per = 9 * CY * CX;
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);       // element handled by this thread column
  const int slice = threadIdx.x >> 5;                         // 8 slices of CTAs per block
  double v = 0.0;
  if (i < per)
    for (int b = slice; b < nctas; b += 8) v += (double)partial[(size_t)b * per + i];
  red[slice][threadIdx.x & 31] = v;
  __syncthreads();
  if (slice == 0 && i < per) {
    v = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += red[k][threadIdx.x];
    const int ci = i % CX, co = (i / CX) % CY, tap = i / (CX * CY);
    dw[((size_t)co * CX + ci) * 9 + tap] = (float)v;
  }
}

// out[c] = sum over blocks of partial[b][row][c] for two rows (fp64): BatchNorm-backward sums -> dbeta (row 0), dgamma (row 1)
__global__ void lf_sum2_kernel(const float* __restrict__ partial, int nblk, int C, float* __restrict__ out0,
                               float* __restrict__ out1) {
  const int c = threadIdx.x;
  if (c >= 2 * C) return;
  double v = 0.0;
  for (int b = 0; b < nblk; ++b) v += (double)partial[(size_t)b * 2 * C + c];
  if (c < C) out0[c] = (float)v;
  else out1[c - C] = (float)v;
}

// ---- 1x1 head -----------------------------------------------------------------------------------------------------------
// train forward: out = sigmoid(sum_c relu(raw3*scale + shift)[c] * wh[c] + bh), one pixel per thread (32 bytes of raw3)
__global__ void lf_head_fwd_kernel(const float* __restrict__ raw3, const float* __restrict__ scale, const float* __restrict__ shift,
                                   const float* __restrict__ wh, const float* __restrict__ bh, size_t P, float* __restrict__ out) {
  __shared__ float c[3][8];
  if (threadIdx.x < 8) { c[0][threadIdx.x] = scale[threadIdx.x]; c[1][threadIdx.x] = shift[threadIdx.x]; c[2][threadIdx.x] = wh[threadIdx.x]; }
  __syncthreads();
  const float b = bh[0];
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float4 r0 = ldg4(raw3 + p * 8), r1 = ldg4(raw3 + p * 8 + 4);
    const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    float acc = b;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(fmaxf(fmaf(r[j], c[0][j], c[1][j]), 0.f), c[2][j], acc);
    out[p] = 1.f / (1.f + expf(-acc));
  }
}

// backward of the head + first half of BatchNorm 3's backward: per block 25 partial sums
//   [0..7] dwh[c] = sum dz * a3[c]   [8] dbh = sum dz   [9..16] sum gz[c]   [17..24] sum gz[c] * xhat[c]
__global__ void __launch_bounds__(256) lf_head_bwd_kernel(const float* __restrict__ raw3, const float* __restrict__ out,
                                                          const float* __restrict__ gout, const float* __restrict__ scale,
                                                          const float* __restrict__ shift, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ wh, size_t P,
                                                          float* __restrict__ partial) {
  __shared__ float c[5][8];
  __shared__ float red[8][25];
  if (threadIdx.x < 8) {
    c[0][threadIdx.x] = scale[threadIdx.x]; c[1][threadIdx.x] = shift[threadIdx.x]; c[2][threadIdx.x] = mean[threadIdx.x];
    c[3][threadIdx.x] = invstd[threadIdx.x]; c[4][threadIdx.x] = wh[threadIdx.x];
  }
  __syncthreads();
  float s[25];
#pragma unroll
  for (int j = 0; j < 25; ++j) s[j] = 0.f;
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
    const float y = out[p];
    const float dz = gout[p] * y * (1.f - y);
    const float4 r0 = ldg4(raw3 + p * 8), r1 = ldg4(raw3 + p * 8 + 4);
    const float r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
    s[8] += dz;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float z = fmaf(r[j], c[0][j], c[1][j]);
      const float gz = z > 0.f ? dz * c[4][j] : 0.f;
      s[j] = fmaf(dz, fmaxf(z, 0.f), s[j]);
      s[9 + j] += gz;
      s[17 + j] = fmaf(gz, (r[j] - c[2][j]) * c[3][j], s[17 + j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 25; ++j) {
    const float v = warp_sum(s[j]);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < 25) {
    float v = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) v += red[wv][threadIdx.x];
    partial[(size_t)blockIdx.x * 25 + threadIdx.x] = v;
  }
}

__global__ void lf_head_bwd_finalize_kernel(const float* __restrict__ partial, int nblk, float* __restrict__ dwh,
                                            float* __restrict__ dbh, float* __restrict__ dbeta3, float* __restrict__ dgamma3) {
  const int j = threadIdx.x;
  if (j >= 25) return;
  double v = 0.0;
  for (int b = 0; b < nblk; ++b) v += (double)partial[(size_t)b * 25 + j];
  if (j < 8) { if (dwh) dwh[j] = (float)v; }
  else if (j == 8) { if (dbh) dbh[0] = (float)v; }
  else if (j < 17) dbeta3[j - 9] = (float)v;
  else dgamma3[j - 17] = (float)v;
}

// eval-mode BatchNorm bookkeeping for the backward (mean = running_mean, invstd = 1/sqrt(running_var + eps))
__global__ void lf_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, int C,
                                     float* __restrict__ mean, float* __restrict__ invstd) {
  const int c = threadIdx.x;
  if (c < C) { mean[c] = rm[c]; invstd[c] = rsqrtf(rv[c] + eps); }
}

int g_sms = 0;
int lf_grid(int ntiles, int* sms_out) {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  }
  if (sms_out) *sms_out = g_sms;
  int grid = 2 * g_sms;   // two CTAs per SM are resident: one stages its window while the other issues MMAs
  if (grid > kMaxCtas) grid = kMaxCtas;
  if (grid > ntiles) grid = ntiles;
  return grid;
}

template <int KIND, int C, int CP, int NOUT, int EPI>
int launch_conv(const ConvArgs& a, int grid, cudaStream_t st) {
  constexpr int STRIDE = CP * 2 + 16;
  constexpr int NROWS = ((NOUT + 7) / 8) * 8;
  const size_t smem = 2 * (size_t)(WH * WW * STRIDE) + 2 * (size_t)(9 * NROWS * STRIDE);
  static unsigned long long attr = 0;
  if (egaze_first_on_device(&attr)) {
    EGAZE_CUDA(cudaFuncSetAttribute(lf_conv_kernel<KIND, C, CP, NOUT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  lf_conv_kernel<KIND, C, CP, NOUT, EPI><<<grid, kConvThreads, smem, st>>>(a);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

template <int KY, int CY, int CPY, int KX, int CX, int CPX>
int launch_wgrad(const WgArgs& a, int grid, cudaStream_t st) {
  const size_t smem = 2 * (size_t)(TH * TW * (CPY * 2 + 16)) + 2 * (size_t)(WH * WW * (CPX * 2 + 16));
  static unsigned long long attr = 0;
  if (egaze_first_on_device(&attr)) {
    EGAZE_CUDA(cudaFuncSetAttribute(lf_wgrad_kernel<KY, CY, CPY, KX, CX, CPX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  lf_wgrad_kernel<KY, CY, CPY, KX, CX, CPX><<<grid, kWgradThreads, smem, st>>>(a);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

}  // namespace

// Scratch sizes (floats) of egaze_lf_fwd / egaze_lf_bwd.
extern "C" int egaze_lf_scratch(int* fwd_floats, int* bwd_floats) {
  if (fwd_floats) *fwd_floats = kMaxCtas * (2 * 32 + 1);
  if (bwd_floats) *bwd_floats = kMaxCtas * 2 * 32 + kMaxCtas * 9 * 32 * 32;
  return EGAZE_OK;
}

// See include/egaze.h.
extern "C" int egaze_lf_fwd(const float* f, const float* g, int B, int H, int W, const float* const* w, const float* const* b,
                            const float* const* gamma, const float* const* beta, float* const* run_mean, float* const* run_var,
                            int training, float eps, float momentum, float* raw1, float* raw2, float* raw3, float* bn_ws,
                            float* scratch, float* out, int precise, void* stream) {
  EGAZE_CHECK_ARG(f && g && w && b && gamma && beta && raw1 && raw2 && bn_ws && scratch && out, "lf_fwd: null argument");
  EGAZE_CHECK_ARG(B > 0 && H > 0 && W > 0, "lf_fwd: bad shape %d %d %d", B, H, W);
  EGAZE_CHECK_ARG(!training || raw3, "lf_fwd: training needs raw3");
  EGAZE_CHECK_ARG(training || (run_mean && run_var && run_mean[0] && run_var[0]), "lf_fwd: eval mode needs running statistics");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_y = ceil_div(H, TH), tiles_x = ceil_div(W, TW), ntiles = B * tiles_y * tiles_x;
  const int grid = lf_grid(ntiles, nullptr);
  EGAZE_CHECK_ARG(grid > 0, "lf_fwd: no device");
  const int couts[3] = {32, 32, 8};
  float* partial = scratch;
  float* cnt = scratch + (size_t)kMaxCtas * 64;
  // bn_ws: [3][4][32] = per layer (mean, invstd, scale, shift)
  auto ws = [&](int layer, int which) { return bn_ws + ((size_t)layer * 4 + which) * 32; };
  auto finalize = [&](int layer) -> int {
    const int C = couts[layer];
    if (training)
      return egaze_bn_finalize(partial, cnt, 1, C, grid, C, eps, momentum, gamma[layer], beta[layer],
                               run_mean ? run_mean[layer] : nullptr, run_var ? run_var[layer] : nullptr, ws(layer, 0),
                               ws(layer, 1), ws(layer, 2), ws(layer, 3), nullptr, stream);
    lf_eval_stats_kernel<<<1, 32, 0, st>>>(run_mean[layer], run_var[layer], eps, C, ws(layer, 0), ws(layer, 1));
    return egaze_bn_fold(gamma[layer], beta[layer], run_mean[layer], run_var[layer], nullptr, eps, C, ws(layer, 2), ws(layer, 3),
                         stream);
  };
  ConvArgs a;
  memset(&a, 0, sizeof(a));
  a.N = B; a.H = H; a.W = W; a.tiles_y = tiles_y; a.tiles_x = tiles_x; a.ntiles = ntiles; a.precise = precise;
  a.ld.inv_n = 1.f / ((float)B * H * W);
  if (training) { a.stat_partial = partial; a.stat_cnt = cnt; }
  int rc;
  // layer 1: cat(f, g) -> raw1
  a.ld.f = f; a.ld.gin = g;
  a.w = w[0]; a.bias = b[0]; a.out = raw1;
  if ((rc = launch_conv<L_INPUT, 2, 16, 32, EPI_RAW>(a, grid, st))) return rc;
  if ((rc = finalize(0))) return rc;
  // layer 2: relu(bn1(raw1)) -> raw2
  a.ld.raw = raw1; a.ld.scale = ws(0, 2); a.ld.shift = ws(0, 3);
  a.w = w[1]; a.bias = b[1]; a.out = raw2;
  if ((rc = launch_conv<L_ACT, 32, 32, 32, EPI_RAW>(a, grid, st))) return rc;
  if ((rc = finalize(1))) return rc;
  // layer 3 (+ head)
  a.ld.raw = raw2; a.ld.scale = ws(1, 2); a.ld.shift = ws(1, 3);
  a.w = w[2]; a.bias = b[2];
  if (training || raw3) {   // eval mode with raw3 given: the caller wants to run the backward, keep the raw map
    a.out = raw3;
    if ((rc = launch_conv<L_ACT, 32, 32, 8, EPI_RAW>(a, grid, st))) return rc;
    if ((rc = finalize(2))) return rc;
    const size_t P = (size_t)B * H * W;
    int blocks = (int)((P + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    lf_head_fwd_kernel<<<blocks, 256, 0, st>>>(raw3, ws(2, 2), ws(2, 3), w[3], b[3], P, out);
    EGAZE_LAUNCH_CHECK();
  } else {
    if ((rc = finalize(2))) return rc;
    a.out = out; a.stat_partial = nullptr;
    a.h_scale = ws(2, 2); a.h_shift = ws(2, 3); a.h_w = w[3]; a.h_b = b[3];
    if ((rc = launch_conv<L_ACT, 32, 32, 8, EPI_HEAD>(a, grid, st))) return rc;
  }
  return EGAZE_OK;
}

// See include/egaze.h.
extern "C" int egaze_lf_bwd(const float* f, const float* g, int B, int H, int W, const float* const* w, const float* raw1,
                            const float* raw2, const float* raw3, const float* bn_ws, int batch_stats, const float* out,
                            const float* gout, float* g1, float* g2, float* scratch, float* const* dw, float* dbh,
                            float* const* dgamma, float* const* dbeta, float* gf, float* gg, int precise, void* stream) {
  EGAZE_CHECK_ARG(f && g && w && raw1 && raw2 && raw3 && bn_ws && out && gout && g1 && g2 && scratch && dw && dgamma && dbeta,
                  "lf_bwd: null argument");
  EGAZE_CHECK_ARG(dgamma[0] && dgamma[1] && dgamma[2] && dbeta[0] && dbeta[1] && dbeta[2], "lf_bwd: BatchNorm sums need buffers");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles_y = ceil_div(H, TH), tiles_x = ceil_div(W, TW), ntiles = B * tiles_y * tiles_x;
  const int grid = lf_grid(ntiles, nullptr);
  EGAZE_CHECK_ARG(grid > 0, "lf_bwd: no device");
  float* bn_partial = scratch;                             // [grid][2][32]
  float* wg_partial = scratch + (size_t)kMaxCtas * 64;     // [grid][9][CY][CX]
  auto ws = [&](int layer, int which) { return bn_ws + ((size_t)layer * 4 + which) * 32; };
  const float inv_n = 1.f / ((float)B * H * W);
  const size_t P = (size_t)B * H * W;
  int rc;

  // head backward + the two BatchNorm-backward sums of layer 3
  {
    int blocks = (int)((P + 255) / 256);
    if (blocks > kMaxCtas) blocks = kMaxCtas;
    lf_head_bwd_kernel<<<blocks, 256, 0, st>>>(raw3, out, gout, ws(2, 2), ws(2, 3), ws(2, 0), ws(2, 1), w[3], P, wg_partial);
    EGAZE_LAUNCH_CHECK();
    lf_head_bwd_finalize_kernel<<<1, 32, 0, st>>>(wg_partial, blocks, dw[3], dbh, dbeta[2], dgamma[2]);
    EGAZE_LAUNCH_CHECK();
  }
  auto draw_args = [&](int layer, const float* raw, const float* gr) {
    LdArgs l;
    memset(&l, 0, sizeof(l));
    l.raw = raw; l.g = gr;
    l.scale = ws(layer, 2); l.shift = ws(layer, 3); l.mean = ws(layer, 0); l.invstd = ws(layer, 1);
    l.dgamma = dgamma[layer]; l.dbeta = dbeta[layer];
    l.inv_n = inv_n; l.batch_stats = batch_stats;
    return l;
  };
  auto act_args = [&](int layer, const float* raw) {
    LdArgs l;
    memset(&l, 0, sizeof(l));
    l.raw = raw; l.scale = ws(layer, 2); l.shift = ws(layer, 3); l.inv_n = inv_n;
    return l;
  };
  LdArgs d3 = draw_args(2, raw3, nullptr);
  d3.gout = gout; d3.out = out; d3.wh = w[3];
  const LdArgs d2 = draw_args(1, raw2, g2), d1 = draw_args(0, raw1, g1);
  LdArgs in;
  memset(&in, 0, sizeof(in));
  in.f = f; in.gin = g; in.inv_n = inv_n;

  WgArgs wa;
  memset(&wa, 0, sizeof(wa));
  wa.N = B; wa.H = H; wa.W = W; wa.tiles_y = tiles_y; wa.tiles_x = tiles_x; wa.ntiles = ntiles; wa.precise = precise;
  wa.partial = wg_partial;
  ConvArgs ca;
  memset(&ca, 0, sizeof(ca));
  ca.N = B; ca.H = H; ca.W = W; ca.tiles_y = tiles_y; ca.tiles_x = tiles_x; ca.ntiles = ntiles; ca.precise = precise;
  ca.transposed = 1;
  ca.stat_partial = bn_partial;

  // layer 3: dW3, then g2 (+ layer 2's BatchNorm sums)
  if (dw[2]) {
    wa.ldy = d3; wa.ldx = act_args(1, raw2);
    if ((rc = launch_wgrad<L_DRAW_HEAD, 8, 16, L_ACT, 32, 32>(wa, grid, st))) return rc;
    lf_wgrad_reduce_kernel<<<ceil_div(9 * 8 * 32, 32), 256, 0, st>>>(wg_partial, grid, 8, 32, dw[2]);
    EGAZE_LAUNCH_CHECK();
  }
  ca.ld = d3; ca.w = w[2]; ca.w_cin = 32; ca.out = g2;
  ca.e_raw = raw2; ca.e_scale = ws(1, 2); ca.e_shift = ws(1, 3); ca.e_mean = ws(1, 0); ca.e_invstd = ws(1, 1);
  if ((rc = launch_conv<L_DRAW_HEAD, 8, 16, 32, EPI_GRAD>(ca, grid, st))) return rc;
  lf_sum2_kernel<<<1, 64, 0, st>>>(bn_partial, grid, 32, dbeta[1], dgamma[1]);
  EGAZE_LAUNCH_CHECK();
  // layer 2
  if (dw[1]) {
    wa.ldy = d2; wa.ldx = act_args(0, raw1);
    if ((rc = launch_wgrad<L_DRAW, 32, 32, L_ACT, 32, 32>(wa, grid, st))) return rc;
    lf_wgrad_reduce_kernel<<<ceil_div(9 * 32 * 32, 32), 256, 0, st>>>(wg_partial, grid, 32, 32, dw[1]);
    EGAZE_LAUNCH_CHECK();
  }
  ca.ld = d2; ca.w = w[1]; ca.w_cin = 32; ca.out = g1;
  ca.e_raw = raw1; ca.e_scale = ws(0, 2); ca.e_shift = ws(0, 3); ca.e_mean = ws(0, 0); ca.e_invstd = ws(0, 1);
  if ((rc = launch_conv<L_DRAW, 32, 32, 32, EPI_GRAD>(ca, grid, st))) return rc;
  lf_sum2_kernel<<<1, 64, 0, st>>>(bn_partial, grid, 32, dbeta[0], dgamma[0]);
  EGAZE_LAUNCH_CHECK();
  // layer 1
  if (dw[0]) {
    wa.ldy = d1; wa.ldx = in;
    if ((rc = launch_wgrad<L_DRAW, 32, 32, L_INPUT, 2, 16>(wa, grid, st))) return rc;
    lf_wgrad_reduce_kernel<<<ceil_div(9 * 32 * 2, 32), 256, 0, st>>>(wg_partial, grid, 32, 2, dw[0]);
    EGAZE_LAUNCH_CHECK();
  }
  if (gf || gg) {
    ca.ld = d1; ca.w = w[0]; ca.w_cin = 2; ca.out = nullptr; ca.stat_partial = nullptr;
    ca.gf = gf; ca.gg = gg;
    if ((rc = launch_conv<L_DRAW, 32, 32, 2, EPI_GX>(ca, grid, st))) return rc;
  }
  return EGAZE_OK;
}
