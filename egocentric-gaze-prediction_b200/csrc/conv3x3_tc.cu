// 3x3 / pad 1 / stride 1 convolution as a tcgen05 implicit GEMM (sm_100a).
//
// Replaces what PyTorch->cuDNN/oneDNN does for every nn.Conv2d(k=3,p=1) on the reference hot path
// (reference utils.py:70 trunk convs, models/model_SP.py:10,13-30 fusion + decoder convs) and, with
// flipped/transposed weights, their data gradients.
//
// Data layout (HBM):
//   activations  : NHWC bf16, as one plane ("fast") or two planes hi/lo with x ~= hi + lo ("precise").
//   weights      : packed [tap = r*3+s][Cout][Cin_p] bf16 (K-major rows of Cin_p), hi/lo planes alike.
//   output       : NHWC fp32 and/or NHWC split-bf16, after the fused epilogue below.
//
// Tiling: one CTA = one spatial tile of BH x BW output pixels (BH*BW <= 128 GEMM rows) x BN output
// channels.  For each 64/32/16-channel chunk of Cin and each horizontal tap s, TMA loads ONE
// (BH+2) x BW x KC input window whose origin is shifted by s-1 pixels (OOB rows/cols zero-filled by
// the TMA unit = the conv padding).  The three vertical taps r read that same window at a row offset
// of r*BW rows, which is a whole number of 8-row swizzle atoms because BW % 8 == 0 -- so the A operand
// is fetched 3x (+halo) instead of 9x.  Weights stream through their own ring, one [BN x KC] box per tap.
// precise mode issues hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM -> regs -> smem staging -> coalesced NHWC stores, BN batch statistics,
// 2x2 max/sum reduction, ReLU/mask, nearest-2x replicate).
#include "common.cuh"

namespace {

constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;

struct ConvTcParams {
  int N, H, W;         // conv output == input spatial size
  int Cin_p, Cout;     // Cin padded to a multiple of KC
  int KC;              // channels per K chunk: 64, 32 or 16
  int BH, BW, BN;      // tile
  int tiles_h, tiles_w, tiles_n;
  int nsplit;          // 1 = single bf16 pass, 2 = hi/lo split (3 MMAs per product)
  int SA, SB;          // ring depths
  int a_slot_bytes;    // bytes of one A plane slot (1024-aligned)
  int b_slot_bytes;    // bytes of one B plane slot
  // epilogue
  const float* bias;   // [Cout] or null
  const float* scale;  // [Cout] or null: v = v*scale + shift (folded eval BatchNorm)
  const float* shift;
  int relu;
  int reduce;          // 0 none, 1 = 2x2 max (MaxPool2d), 2 = 2x2 sum (grad of nearest upsample)
  int ups;             // replicate every output pixel 2x2 (nn.Upsample(scale_factor=2))
  const __nv_bfloat16* mask;  // optional NHWC [N,Ho,Wo,Cout]: zero the output where mask <= 0 (ReLU backward)
  int mask_ups;               // mask is stored 2x nearest-upsampled ([N,2Ho,2Wo,Cout]); read its (2oh,2ow) sample
  float* out_f32;             // optional NHWC fp32 [N,Ho*,Wo*,Cout]
  __nv_bfloat16* out_hi;      // optional NHWC bf16
  __nv_bfloat16* out_lo;      // optional (precise) NHWC bf16
  float* stats;               // optional [num_tiles][2][Cout] per-tile (mean, M2) of the pre-activation
  float* stats_cnt;           // [num_tiles] valid-pixel count of each tile
};

__device__ __forceinline__ bool row_valid(const ConvTcParams& p, int m, int h0, int w0) {
  if (m >= p.BH * p.BW) return false;
  int th = m / p.BW, tw = m - th * p.BW;
  return (h0 + th < p.H) && (w0 + tw < p.W);
}

template <int NSPLIT, int KSTEPS>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-align the ring base (SWIZZLE_128B atoms repeat every 1024 B of *absolute* smem address).
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // ---- tile coordinates (n-tile fastest so concurrent CTAs share the activation window in L2) ----
  int bid = blockIdx.x;
  const int nt = bid % p.tiles_n;
  bid /= p.tiles_n;
  const int tw_i = bid % p.tiles_w;
  bid /= p.tiles_w;
  const int th_i = bid % p.tiles_h;
  const int img = bid / p.tiles_h;
  const int h0 = th_i * p.BH, w0 = tw_i * p.BW, n0 = nt * p.BN;
  const int tile_linear = (img * p.tiles_h + th_i) * p.tiles_w + tw_i;

  uint8_t* a_ring = smem;                                           // SA slots x nsplit planes
  uint8_t* b_ring = smem + (size_t)p.SA * p.nsplit * p.a_slot_bytes;  // SB slots x nsplit planes
  __shared__ uint64_t a_full[4], a_empty[4], b_full[8], b_empty[8], acc_full;
  __shared__ uint32_t tmem_base_smem;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], 1); }
    ptx::mbar_init(&acc_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA_hi);
    ptx::prefetch_tmap(&tmB_hi);
    if (p.nsplit == 2) { ptx::prefetch_tmap(&tmA_lo); ptx::prefetch_tmap(&tmB_lo); }
  }
  const uint32_t tmem_cols = p.BN < 32 ? 32u : (uint32_t)p.BN;
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_smem, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int chunks = p.Cin_p / p.KC;
  const int row_bytes = p.KC * 2;
  const uint32_t a_box_bytes = (uint32_t)((p.BH + 2) * p.BW * row_bytes);
  const uint32_t b_box_bytes = (uint32_t)(p.BN * row_bytes);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t a_par = 1, b_par = 1;  // a fresh mbarrier passes a parity-1 wait: the first lap never blocks
      for (int kc = 0; kc < chunks; ++kc) {
        for (int s = 0; s < 3; ++s) {
          ptx::mbar_wait(&a_empty[sa], a_par);
          ptx::mbar_arrive_expect_tx(&a_full[sa], a_box_bytes * NSPLIT);
          uint8_t* a_dst = a_ring + (size_t)sa * NSPLIT * p.a_slot_bytes;
          ptx::tma_load_4d(a_dst, &tmA_hi, &a_full[sa], kc * p.KC, w0 - 1 + s, h0 - 1, img);
          if (NSPLIT == 2)
            ptx::tma_load_4d(a_dst + p.a_slot_bytes, &tmA_lo, &a_full[sa], kc * p.KC, w0 - 1 + s, h0 - 1, img);
          if (++sa == p.SA) { sa = 0; a_par ^= 1; }
          for (int r = 0; r < 3; ++r) {
            ptx::mbar_wait(&b_empty[sb], b_par);
            ptx::mbar_arrive_expect_tx(&b_full[sb], b_box_bytes * NSPLIT);
            uint8_t* b_dst = b_ring + (size_t)sb * NSPLIT * p.b_slot_bytes;
            const int brow = (r * 3 + s) * p.Cout + n0;
            ptx::tma_load_2d(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow);
            if (NSPLIT == 2) ptx::tma_load_2d(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow);
            if (++sb == p.SB) { sb = 0; b_par ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.BN, 0, 0);
      const uint32_t sbo = 8u * (uint32_t)row_bytes;
      // Descriptors differ only in their 14-bit start-address field (smem address >> 4; smem < 256 KB so the
      // field never carries): build the static part once, then a descriptor is one 64-bit add.
      const uint64_t desc_static = ptx::make_smem_desc(0, 16, sbo, (uint32_t)row_bytes);
      const uint64_t a_ring_desc = desc_static + (uint64_t)(ptx::smem_u32(a_ring) >> 4);
      const uint64_t b_ring_desc = desc_static + (uint64_t)(ptx::smem_u32(b_ring) >> 4);
      const uint32_t a_slot16 = (uint32_t)(p.nsplit * p.a_slot_bytes) >> 4, a_plane16 = (uint32_t)p.a_slot_bytes >> 4;
      const uint32_t b_slot16 = (uint32_t)(p.nsplit * p.b_slot_bytes) >> 4, b_plane16 = (uint32_t)p.b_slot_bytes >> 4;
      const uint32_t r_step16 = (uint32_t)(p.BW * row_bytes) >> 4;
      int a_it = 0, b_it = 0, sa = 0, sb = 0;
      uint32_t a_par = 0, b_par = 0;
      uint32_t accumulate = 0;
      for (int kc = 0; kc < chunks; ++kc) {
        for (int s = 0; s < 3; ++s) {
          ptx::mbar_wait(&a_full[sa], a_par);
          ptx::tc_fence_after();
          const uint64_t a_desc0 = a_ring_desc + (uint64_t)((uint32_t)sa * a_slot16);
#pragma unroll 1
          for (int r = 0; r < 3; ++r) {
            ptx::mbar_wait(&b_full[sb], b_par);
            ptx::tc_fence_after();
            const uint64_t ad = a_desc0 + (uint64_t)((uint32_t)r * r_step16);
            const uint64_t bd = b_ring_desc + (uint64_t)((uint32_t)sb * b_slot16);
            // products: (hi,hi) [, (hi,lo), (lo,hi)]
#pragma unroll
            for (int k = 0; k < KSTEPS; ++k) {
              ptx::umma_bf16(tmem_base, ad + 2 * k, bd + 2 * k, idesc, accumulate);
              accumulate = 1;
            }
            if (NSPLIT == 2) {
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) ptx::umma_bf16(tmem_base, ad + 2 * k, bd + b_plane16 + 2 * k, idesc, 1);
#pragma unroll
              for (int k = 0; k < KSTEPS; ++k) ptx::umma_bf16(tmem_base, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);
            }
            ptx::umma_commit(&b_empty[sb]);
            if (++sb == p.SB) { sb = 0; b_par ^= 1; }
            ++b_it;
          }
          ptx::umma_commit(&a_empty[sa]);
          if (++sa == p.SA) { sa = 0; a_par ^= 1; }
          ++a_it;
        }
      }
      ptx::umma_commit(&acc_full);
    }
  } else {
    // ================================ epilogue warps ================================
    const int ew = warp & 3;                 // TMEM lane group this warp may access
    const int m = ew * 32 + lane;            // GEMM row == TMEM lane == pixel index inside the tile
    const int et = threadIdx.x - 64;         // 0..127
    const int ldst = p.BN + 4;               // padded staging row (floats)
    float* stage = reinterpret_cast<float*>(smem);

    ptx::mbar_wait(&acc_full, 0);
    ptx::tc_fence_after();

    // -- phase 1: TMEM -> registers -> (+bias, *scale+shift, relu) -> smem staging [128][BN+4]
    for (int c0 = 0; c0 < p.BN; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
      float* dst = stage + (size_t)m * ldst + c0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (c0 + j >= p.BN) break;  // BN == 16: only half of the 32-column TMEM load is live
        float4 o;
        float* po = reinterpret_cast<float*>(&o);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x = __uint_as_float(v[j + q]);
          const int c = n0 + c0 + j + q;
          if (p.bias) x += __ldg(p.bias + c);
          if (p.scale) x = fmaf(x, __ldg(p.scale + c), __ldg(p.shift + c));
          if (p.relu) x = fmaxf(x, 0.f);
          po[q] = x;
        }
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    }
    ptx::tc_fence_before();
    ptx::named_bar_sync(1, kEpiThreads);

    // -- phase 2: per-tile BatchNorm statistics (mean, M2 over the tile's valid pixels).  Two threads per column
    //    when BN <= 64 (row halves combined with Chan's formula), one otherwise; no per-row index arithmetic.
    if (p.stats) {
      const int vh = min(p.BH, p.H - h0), vw = min(p.BW, p.W - w0);   // valid extent of this tile
      const int cnt = vh * vw;
      const int halves = (p.BN <= 64) ? 2 : 1;
      float* red = stage + (size_t)128 * ldst;                        // scratch behind the staging tile
      for (int cbase = 0; cbase < p.BN; cbase += kEpiThreads / halves) {
        const int c = cbase + (et % (kEpiThreads / halves));
        const int half = et / (kEpiThreads / halves);
        const int r0 = halves == 2 ? (half == 0 ? 0 : vh / 2) : 0;
        const int r1 = halves == 2 ? (half == 0 ? vh / 2 : vh) : vh;
        const int n_loc = (r1 - r0) * vw;
        float sum = 0.f;
        if (c < p.BN)
          for (int th = r0; th < r1; ++th) {
            const float* rowp = stage + (size_t)(th * p.BW) * ldst + c;
            for (int tw = 0; tw < vw; ++tw) sum += rowp[(size_t)tw * ldst];
          }
        const float mean = n_loc > 0 ? sum / (float)n_loc : 0.f;
        float m2 = 0.f;
        if (c < p.BN)
          for (int th = r0; th < r1; ++th) {
            const float* rowp = stage + (size_t)(th * p.BW) * ldst + c;
            for (int tw = 0; tw < vw; ++tw) {
              const float d = rowp[(size_t)tw * ldst] - mean;
              m2 = fmaf(d, d, m2);
            }
          }
        if (halves == 2) {
          if (half == 1 && c < p.BN) { red[c * 3 + 0] = mean; red[c * 3 + 1] = m2; red[c * 3 + 2] = (float)n_loc; }
          ptx::named_bar_sync(2, kEpiThreads);
          if (half == 0 && c < p.BN) {
            const float nb = red[c * 3 + 2], mb = red[c * 3 + 0], m2b = red[c * 3 + 1];
            const float na = (float)n_loc, nn = na + nb;
            const float d = mb - mean;
            const float mean_t = nn > 0.f ? mean + d * nb / nn : 0.f;
            const float m2_t = m2 + m2b + (nn > 0.f ? d * d * na * nb / nn : 0.f);
            p.stats[((size_t)tile_linear * 2 + 0) * p.Cout + n0 + c] = mean_t;
            p.stats[((size_t)tile_linear * 2 + 1) * p.Cout + n0 + c] = m2_t;
          }
          ptx::named_bar_sync(2, kEpiThreads);
        } else if (c < p.BN) {
          p.stats[((size_t)tile_linear * 2 + 0) * p.Cout + n0 + c] = mean;
          p.stats[((size_t)tile_linear * 2 + 1) * p.Cout + n0 + c] = m2;
        }
      }
      if (et == 0 && nt == 0) p.stats_cnt[tile_linear] = (float)cnt;
    }

    // -- phase 3: coalesced NHWC stores (float4 of 4 channels per thread)
    const int g_per_pix = p.BN / 4;
    const int oBH = p.reduce ? p.BH / 2 : p.BH;
    const int oBW = p.reduce ? p.BW / 2 : p.BW;
    const int Ho = p.reduce ? p.H / 2 : p.H;
    const int Wo = p.reduce ? p.W / 2 : p.W;
    const int oh0 = p.reduce ? h0 / 2 : h0;
    const int ow0 = p.reduce ? w0 / 2 : w0;
    const int total = oBH * oBW * g_per_pix;
    for (int idx = et; idx < total; idx += kEpiThreads) {
      const int pix = idx / g_per_pix;
      const int g = idx - pix * g_per_pix;
      const int ph = pix / oBW, pw = pix - ph * oBW;
      const int oh = oh0 + ph, ow = ow0 + pw;
      if (oh >= Ho || ow >= Wo) continue;
      float4 v;
      if (p.reduce == 0) {
        v = *reinterpret_cast<const float4*>(stage + (size_t)(ph * p.BW + pw) * ldst + g * 4);
      } else {
        const float4 a = *reinterpret_cast<const float4*>(stage + (size_t)((2 * ph) * p.BW + 2 * pw) * ldst + g * 4);
        const float4 b = *reinterpret_cast<const float4*>(stage + (size_t)((2 * ph) * p.BW + 2 * pw + 1) * ldst + g * 4);
        const float4 c = *reinterpret_cast<const float4*>(stage + (size_t)((2 * ph + 1) * p.BW + 2 * pw) * ldst + g * 4);
        const float4 d = *reinterpret_cast<const float4*>(stage + (size_t)((2 * ph + 1) * p.BW + 2 * pw + 1) * ldst + g * 4);
        if (p.reduce == 1) {
          v.x = fmaxf(fmaxf(a.x, b.x), fmaxf(c.x, d.x));
          v.y = fmaxf(fmaxf(a.y, b.y), fmaxf(c.y, d.y));
          v.z = fmaxf(fmaxf(a.z, b.z), fmaxf(c.z, d.z));
          v.w = fmaxf(fmaxf(a.w, b.w), fmaxf(c.w, d.w));
        } else {
          v.x = (a.x + b.x) + (c.x + d.x);
          v.y = (a.y + b.y) + (c.y + d.y);
          v.z = (a.z + b.z) + (c.z + d.z);
          v.w = (a.w + b.w) + (c.w + d.w);
        }
      }
      const int ch = n0 + g * 4;
      if (p.mask) {
        const size_t mpix = p.mask_ups ? ((size_t)(img * 2 * Ho + 2 * oh) * (2 * Wo) + 2 * ow) : ((size_t)(img * Ho + oh) * Wo + ow);
        const uint2 mk = __ldg(reinterpret_cast<const uint2*>(p.mask + mpix * p.Cout + ch));
        if (!(bf16_bits_to_float(mk.x & 0xffffu) > 0.f)) v.x = 0.f;
        if (!(bf16_bits_to_float(mk.x >> 16) > 0.f)) v.y = 0.f;
        if (!(bf16_bits_to_float(mk.y & 0xffffu) > 0.f)) v.z = 0.f;
        if (!(bf16_bits_to_float(mk.y >> 16) > 0.f)) v.w = 0.f;
      }
      uint2 hi2 = make_uint2(0, 0), lo2 = make_uint2(0, 0);
      if (p.out_hi) {
        __nv_bfloat16 h[4], l[4];
        split_bf16(v.x, h[0], l[0]);
        split_bf16(v.y, h[1], l[1]);
        split_bf16(v.z, h[2], l[2]);
        split_bf16(v.w, h[3], l[3]);
        hi2 = make_uint2(pack_bf16x2(h[0], h[1]), pack_bf16x2(h[2], h[3]));
        lo2 = make_uint2(pack_bf16x2(l[0], l[1]), pack_bf16x2(l[2], l[3]));
      }
      const int rep = p.ups ? 2 : 1;
      const int Hs = Ho * rep, Ws = Wo * rep;
      for (int dy = 0; dy < rep; ++dy)
        for (int dx = 0; dx < rep; ++dx) {
          const size_t off = ((size_t)(img * Hs + oh * rep + dy) * Ws + (ow * rep + dx)) * p.Cout + ch;
          if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + off) = v;
          if (p.out_hi) *reinterpret_cast<uint2*>(p.out_hi + off) = hi2;
          if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + off) = lo2;
        }
    }
  }

  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc(tmem_base, tmem_cols);
}

// Tile shape for an H x W map: BW % 8 == 0, BH*BW <= 128, maximise useful rows then minimise halo.
void pick_tile(int H, int W, int need_even, int* BH, int* BW) {
  const int cand[][2] = {{8, 16}, {4, 32}, {16, 8}, {14, 8}, {7, 16}, {2, 64}, {12, 8}, {6, 16}, {3, 32}, {10, 8}, {5, 24}, {4, 24}};
  double best = -1.0;
  for (auto& c : cand) {
    const int bh = c[0], bw = c[1];
    if (need_even && (bh & 1)) continue;
    const double eff = (double)H * W / ((double)ceil_div(H, bh) * ceil_div(W, bw) * 128.0);
    const double halo = 3.0 * (bh + 2) / bh;
    const double score = eff - 0.01 * halo;
    if (score > best) { best = score; *BH = bh; *BW = bw; }
  }
}

}  // namespace

// Number of spatial tiles the kernel will use for an (N,H,W) problem: sizes the stats workspace.
extern "C" int egaze_conv3x3_tiles(int N, int H, int W, int need_even, int* num_tiles, int* BH_out, int* BW_out) {
  int BH = 8, BW = 16;
  pick_tile(H, W, need_even, &BH, &BW);
  if (num_tiles) *num_tiles = N * ceil_div(H, BH) * ceil_div(W, BW);
  if (BH_out) *BH_out = BH;
  if (BW_out) *BW_out = BW;
  return EGAZE_OK;
}

// See include/egaze.h for the contract.
extern "C" int egaze_conv3x3_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H, int W,
                                int Cin_p, int Cout, const float* bias, const float* scale, const float* shift, int relu,
                                int reduce, int ups, const void* mask, int mask_ups, float* out_f32, void* out_hi,
                                void* out_lo, float* stats, float* stats_cnt, int precise, void* stream) {
  EGAZE_CHECK_ARG(x_hi && w_hi, "conv3x3_tc: null operand");
  EGAZE_CHECK_ARG(!precise || (x_lo && w_lo), "conv3x3_tc: precise mode needs lo planes");
  EGAZE_CHECK_ARG(N > 0 && H > 0 && W > 0, "conv3x3_tc: bad shape %d %d %d", N, H, W);
  EGAZE_CHECK_ARG(Cin_p % 16 == 0, "conv3x3_tc: Cin_p=%d must be a multiple of 16", Cin_p);
  EGAZE_CHECK_ARG(Cout % 16 == 0 && Cout >= 16, "conv3x3_tc: Cout=%d must be a multiple of 16", Cout);
  EGAZE_CHECK_ARG(!(reduce && ((H | W) & 1)), "conv3x3_tc: 2x2 reduce needs even H, W");
  EGAZE_CHECK_ARG(out_f32 || out_hi, "conv3x3_tc: no output");

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin_p = Cin_p; p.Cout = Cout;
  p.KC = (Cin_p % 64 == 0) ? 64 : ((Cin_p % 32 == 0) ? 32 : 16);
  pick_tile(H, W, reduce != 0, &p.BH, &p.BW);
  p.nsplit = precise ? 2 : 1;
  // N tile: largest of 128/64/32/16 dividing Cout (256 only in fast mode where the rings fit).
  p.BN = 16;
  for (int bn : {32, 64, 128}) if (Cout % bn == 0) p.BN = bn;
  if (!precise && Cout % 256 == 0) p.BN = 256;
  p.tiles_h = ceil_div(H, p.BH); p.tiles_w = ceil_div(W, p.BW); p.tiles_n = Cout / p.BN;
  const int row_bytes = p.KC * 2;
  int a_rows = (p.BH + 2) * p.BW;
  if (a_rows < 2 * p.BW + 128) a_rows = 2 * p.BW + 128;
  p.a_slot_bytes = ((a_rows * row_bytes + 1023) / 1024) * 1024;
  p.b_slot_bytes = ((p.BN * row_bytes + 1023) / 1024) * 1024;
  p.SA = 2;
  const int budget = 200 * 1024;
  int sb = (budget - p.SA * p.nsplit * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes);
  if (sb > 6) sb = 6;
  EGAZE_CHECK_ARG(sb >= 2, "conv3x3_tc: tile does not fit shared memory");
  p.SB = sb;
  size_t smem = (size_t)p.SA * p.nsplit * p.a_slot_bytes + (size_t)p.SB * p.nsplit * p.b_slot_bytes;
  const size_t stage_bytes = (size_t)128 * (p.BN + 4) * 4 + 3 * 64 * 4;
  if (smem < stage_bytes) smem = stage_bytes;
  smem += 1024;  // alignment slack
  p.bias = bias; p.scale = scale; p.shift = shift; p.relu = relu; p.reduce = reduce; p.ups = ups;
  p.mask = (const __nv_bfloat16*)mask;
  p.mask_ups = mask_ups;
  p.out_f32 = out_f32; p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
  p.stats = stats; p.stats_cnt = stats_cnt;

  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  {
    uint64_t dims[4] = {(uint64_t)Cin_p, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t str[3] = {(uint64_t)Cin_p * 2, (uint64_t)W * Cin_p * 2, (uint64_t)H * W * Cin_p * 2};
    uint32_t box[4] = {(uint32_t)p.KC, (uint32_t)p.BW, (uint32_t)(p.BH + 2), 1};
    int rc = egaze_encode_tmap(&tmA_hi, x_hi, 4, dims, str, box, row_bytes, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmA_lo, precise ? x_lo : x_hi, 4, dims, str, box, row_bytes, 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Cin_p, (uint64_t)9 * Cout};
    uint64_t str[1] = {(uint64_t)Cin_p * 2};
    uint32_t box[2] = {(uint32_t)p.KC, (uint32_t)p.BN};
    int rc = egaze_encode_tmap(&tmB_hi, w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmB_lo, precise ? w_lo : w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
  }
  dim3 grid((unsigned)(p.tiles_n * p.tiles_w * p.tiles_h * N));
  const int ksteps = p.KC / 16;
#define EGAZE_CONV_LAUNCH(NS, KS)                                                                                     \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      EGAZE_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<NS, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                      224 * 1024));                                                                   \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    conv3x3_tc_kernel<NS, KS><<<grid, kThreads, smem, (cudaStream_t)stream>>>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, p);     \
  } while (0)
  if (p.nsplit == 2) {
    if (ksteps == 4) EGAZE_CONV_LAUNCH(2, 4);
    else if (ksteps == 2) EGAZE_CONV_LAUNCH(2, 2);
    else EGAZE_CONV_LAUNCH(2, 1);
  } else {
    if (ksteps == 4) EGAZE_CONV_LAUNCH(1, 4);
    else if (ksteps == 2) EGAZE_CONV_LAUNCH(1, 2);
    else EGAZE_CONV_LAUNCH(1, 1);
  }
#undef EGAZE_CONV_LAUNCH
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
