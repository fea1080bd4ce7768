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
// Tiling: one work item = one spatial tile of BH x BW output pixels (BH*BW <= 128 GEMM rows) x BN output
// channels.  For each 64/32/16-channel chunk of Cin and each horizontal tap s, TMA loads ONE
// (BH+2) x BW x KC input window whose origin is shifted by s-1 pixels (OOB rows/cols zero-filled by
// the TMA unit = the conv padding).  The three vertical taps r read that same window at a row offset
// of r*BW rows, which is a whole number of 8-row swizzle atoms because BW % 8 == 0 -- so the A operand
// is fetched 3x (+halo) instead of 9x.  Weights stream through their own ring, one [BN x KC] box per tap.
// precise mode issues hi*hi + hi*lo + lo*hi into the same fp32 TMEM accumulator.
//
// Schedule: PERSISTENT CTAs (one per SM), launched as thread-block clusters of CS (1 or 2).  The CTAs of a cluster
// work on adjacent spatial tiles of the SAME output-channel tile in lockstep, so the weight box of every tap is
// fetched from L2 once per cluster: each CTA loads 1/CS of it and TMA-multicasts it into all CTAs' rings
// (weight traffic is the dominant L2->SM stream of this kernel).  Two TMEM accumulator stages let the epilogue of
// item i overlap the MMAs of item i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (one elected lane),
// warps 2..9 = epilogue (TMEM -> regs -> smem staging in 64-column chunks -> coalesced NHWC stores, BN batch
// statistics, 2x2 max/sum reduction, ReLU/mask, nearest-2x replicate).
#include "common.cuh"

namespace {

constexpr int kThreads = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kMaxSA = 4, kMaxSB = 8;

struct ConvTcParams {
  int N, H, W;         // conv output == input spatial size
  int Cin_p, Cout;     // Cin padded to a multiple of KC
  int KC;              // channels per K chunk: 64, 32 or 16
  int BH, BW, BN;      // tile
  int tiles_h, tiles_w, tiles_n;
  int total_tiles;     // N * tiles_h * tiles_w
  int tile_groups;     // ceil(total_tiles / CS)
  int num_items;       // tile_groups * tiles_n
  int nsplit;          // 1 = single bf16 pass, 2 = hi/lo split (3 MMAs per product)
  int merged;          // precise mode: A_hi x [B_hi | B_lo] as ONE MMA of N = 2*BN (the two halves are summed in the epilogue)
  int SA, SB;          // ring depths
  int a_slot_bytes;    // bytes of one A plane slot (1024-aligned)
  int b_slot_bytes;    // bytes of one B plane slot
  int stage_off;       // byte offset of the epilogue staging buffer inside dynamic smem
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
  float* stats;               // optional [gridDim][2][Cout] per-CTA (mean, M2) of the pre-activation (acc + bias)
  float* stats_cnt;           // [gridDim][tiles_n] pixel count behind each per-CTA partial
};

struct Item {
  int nt, tile, img, h0, w0;
  bool valid;
};

__device__ __forceinline__ Item decode_item(const ConvTcParams& p, int w, int cs, int rank) {
  Item it;
  it.nt = w % p.tiles_n;                       // n-tile fastest: consecutive items reuse the activation window in L2
  const int tg = w / p.tiles_n;
  it.tile = tg * cs + rank;
  it.valid = it.tile < p.total_tiles;
  const int t = it.valid ? it.tile : 0;
  const int tw_i = t % p.tiles_w;
  const int th_i = (t / p.tiles_w) % p.tiles_h;
  it.img = it.valid ? t / (p.tiles_w * p.tiles_h) : p.N;   // img == N: every TMA box is out of bounds -> zeros
  it.h0 = th_i * p.BH;
  it.w0 = tw_i * p.BW;
  return it;
}

template <int NSPLIT, int KSTEPS, int CS>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                  const ConvTcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-align the ring base (SWIZZLE_128B atoms repeat every 1024 B of *absolute* smem address).  The dynamic
  // smem window starts at the same offset in every CTA of the cluster, so ring offsets match across CTAs.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = CS > 1 ? (int)ptx::cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CS;
  const int num_clusters = gridDim.x / CS;
  constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1);

  uint8_t* a_ring = smem;                                                 // SA slots x NSPLIT planes
  uint8_t* b_ring = smem + (size_t)p.SA * NSPLIT * p.a_slot_bytes;         // SB slots x NSPLIT planes
  float* stage = reinterpret_cast<float*>(smem + p.stage_off);             // [128][CW+4] + scratch
  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], CS); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&acc_full[i], 1); ptx::mbar_init(&acc_empty[i], kEpiThreads / 32); }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA_hi);
    ptx::prefetch_tmap(&tmB_hi);
    if (NSPLIT == 2) { ptx::prefetch_tmap(&tmA_lo); ptx::prefetch_tmap(&tmB_lo); }
  }
  const uint32_t acc_cols = p.merged ? 2u * (uint32_t)p.BN : (p.BN < 32 ? 32u : (uint32_t)p.BN);   // columns of one accumulator stage
  const uint32_t tmem_cols = 2 * acc_cols;                       // power of two in [64, 512]
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_base_smem, tmem_cols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int chunks = p.Cin_p / p.KC;
  const int row_bytes = p.KC * 2;
  const uint32_t a_box_bytes = (uint32_t)((p.BH + 2) * p.BW * row_bytes);
  const uint32_t b_box_bytes = (uint32_t)(p.BN * row_bytes);
  const int b_rows_cta = p.BN / CS;                              // weight rows this CTA fetches (and multicasts)

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int sa = 0, sb = 0;
      uint32_t a_par = 1, b_par = 1;  // a fresh mbarrier passes a parity-1 wait: the first lap never blocks
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const Item it = decode_item(p, w, CS, rank);
        const int n0 = it.nt * p.BN;
        for (int kc = 0; kc < chunks; ++kc) {
          for (int s = 0; s < 3; ++s) {
            ptx::mbar_wait(&a_empty[sa], a_par);
            ptx::mbar_arrive_expect_tx(&a_full[sa], a_box_bytes * NSPLIT);
            uint8_t* a_dst = a_ring + (size_t)sa * NSPLIT * p.a_slot_bytes;
            ptx::tma_load_4d(a_dst, &tmA_hi, &a_full[sa], kc * p.KC, it.w0 - 1 + s, it.h0 - 1, it.img);
            if (NSPLIT == 2)
              ptx::tma_load_4d(a_dst + p.a_slot_bytes, &tmA_lo, &a_full[sa], kc * p.KC, it.w0 - 1 + s, it.h0 - 1, it.img);
            if (++sa == p.SA) { sa = 0; a_par ^= 1; }
            for (int r = 0; r < 3; ++r) {
              ptx::mbar_wait(&b_empty[sb], b_par);   // every CTA of the cluster has drained this slot
              ptx::mbar_arrive_expect_tx(&b_full[sb], b_box_bytes * NSPLIT);
              uint8_t* b_dst = b_ring + (size_t)sb * NSPLIT * p.b_slot_bytes + (size_t)rank * b_rows_cta * row_bytes;
              const int brow = (r * 3 + s) * p.Cout + n0 + rank * b_rows_cta;
              if (CS > 1) {
                ptx::tma_load_2d_mc(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow, kMask);
                if (NSPLIT == 2) ptx::tma_load_2d_mc(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow, kMask);
              } else {
                ptx::tma_load_2d(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow);
                if (NSPLIT == 2) ptx::tma_load_2d(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow);
              }
              if (++sb == p.SB) { sb = 0; b_par ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, p.BN, 0, 0);
      const uint32_t idesc2 = ptx::make_idesc_bf16(128, 2 * p.BN, 0, 0);
      const uint32_t sbo = 8u * (uint32_t)row_bytes;
      // Descriptors differ only in their 14-bit start-address field (smem address >> 4; smem < 256 KB so the
      // field never carries): build the static part once, then a descriptor is one 64-bit add.
      const uint64_t desc_static = ptx::make_smem_desc(0, 16, sbo, (uint32_t)row_bytes);
      const uint64_t a_ring_desc = desc_static + (uint64_t)(ptx::smem_u32(a_ring) >> 4);
      const uint64_t b_ring_desc = desc_static + (uint64_t)(ptx::smem_u32(b_ring) >> 4);
      const uint32_t a_slot16 = (uint32_t)(NSPLIT * p.a_slot_bytes) >> 4, a_plane16 = (uint32_t)p.a_slot_bytes >> 4;
      const uint32_t b_slot16 = (uint32_t)(NSPLIT * p.b_slot_bytes) >> 4, b_plane16 = (uint32_t)p.b_slot_bytes >> 4;
      const uint32_t r_step16 = (uint32_t)(p.BW * row_bytes) >> 4;
      int sa = 0, sb = 0, as = 0;
      uint32_t a_par = 0, b_par = 0, acc_par[2] = {1, 1};
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        ptx::mbar_wait(&acc_empty[as], acc_par[as]);   // epilogue has drained this accumulator stage
        acc_par[as] ^= 1;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
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
              // products: (hi,hi) [, (hi,lo), (lo,hi)].  merged: the hi and lo weight planes are contiguous in the slot, so
              // A_hi x [B_hi | B_lo] is ONE MMA of N = 2*BN (A_hi is read from smem once instead of twice -- the MMA is
              // shared-memory-read bound at N <= 128); its two column halves are added in the epilogue.
              if (NSPLIT == 2 && p.merged) {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc2, accumulate);
                  ptx::umma_bf16(d_tmem, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);
                  accumulate = 1;
                }
              } else {
#pragma unroll
                for (int k = 0; k < KSTEPS; ++k) {
                  ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accumulate);
                  accumulate = 1;
                }
                if (NSPLIT == 2) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) ptx::umma_bf16(d_tmem, ad + 2 * k, bd + b_plane16 + 2 * k, idesc, 1);
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) ptx::umma_bf16(d_tmem, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);
                }
              }
              if (CS > 1) ptx::umma_commit_mc(&b_empty[sb], kMask);
              else ptx::umma_commit(&b_empty[sb]);
              if (++sb == p.SB) { sb = 0; b_par ^= 1; }
            }
            ptx::umma_commit(&a_empty[sa]);
            if (++sa == p.SA) { sa = 0; a_par ^= 1; }
          }
        }
        ptx::umma_commit(&acc_full[as]);
        as ^= 1;
      }
    }
  } else {
    // ================================ epilogue warps (8) ================================
    // Two warps per TMEM lane group: warps 2..5 take the even 32-column blocks, warps 6..9 the odd ones.
    const int ew = warp - 2;                 // 0..7
    const int lg = warp & 3;                 // TMEM lane group this warp may access
    const int colsel = ew >> 2;              // which 32-column block of a 64-column chunk this warp drains
    const int m = lg * 32 + lane;            // GEMM row == TMEM lane == pixel index inside the tile
    const int et = threadIdx.x - 64;         // 0..255
    const int CW = p.BN < 64 ? p.BN : 64;    // staged column chunk (power of two)
    const int ldst = CW + 4;                 // padded staging row (floats)
    float* red = stage + (size_t)128 * ldst; // [4][64][3] floats of scratch
    float* run_stats = red + 4 * 64 * 3;     // [Cout][3] running (n, mean, M2) of this CTA (BatchNorm statistics)
    if (p.stats) {
      for (int i = et; i < p.Cout * 3; i += kEpiThreads) run_stats[i] = 0.f;
      ptx::named_bar_sync(1, kEpiThreads);
    }
    const int gpp = CW >> 2;                 // float4 channel groups per pixel (power of two)
    const int gpp_log = 31 - __clz(gpp);
    const int g = et & (gpp - 1);            // this thread's channel group ...
    const int pl = et >> gpp_log;            // ... and pixel lane
    const int PS = kEpiThreads >> gpp_log;   // pixels covered per sweep
    const int oBH = p.reduce ? p.BH / 2 : p.BH;
    const int oBW = p.reduce ? p.BW / 2 : p.BW;
    const int Ho = p.reduce ? p.H / 2 : p.H;
    const int Wo = p.reduce ? p.W / 2 : p.W;
    const int ph_start = pl / oBW, pw_start = pl - ph_start * oBW;
    const int rep = p.ups ? 2 : 1;
    const int Hs = Ho * rep, Ws = Wo * rep;
    int as = 0;
    uint32_t full_par[2] = {0, 0};
    for (int w = cluster_id; w < p.num_items; w += num_clusters) {
      const Item it = decode_item(p, w, CS, rank);
      const int h0 = it.h0, w0 = it.w0, img = it.img;
      ptx::mbar_wait(&acc_full[as], full_par[as]);
      full_par[as] ^= 1;
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)as * acc_cols + ((uint32_t)(lg * 32) << 16);

      for (int cc = 0; cc < p.BN; cc += CW) {
        const int n0 = it.nt * p.BN + cc;   // first output channel of this chunk
        // -- phase 1: TMEM -> registers -> smem staging [128][CW+4] (raw fp32 accumulators)
        for (int c0 = colsel * 32; c0 < CW; c0 += 64) {
          uint32_t v[32];
          ptx::tmem_ld_32x32(t_acc + (uint32_t)(cc + c0), v);
          if (p.merged) {
            uint32_t v2[32];
            ptx::tmem_ld_32x32(t_acc + (uint32_t)(p.BN + cc + c0), v2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          } else {
            ptx::tmem_ld_wait();
          }
          float* dst = stage + (size_t)m * ldst + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (c0 + j >= CW) break;  // BN == 16: only half of the 32-column TMEM load is live
            *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          }
        }
        if (cc + CW >= p.BN) {
          // last chunk read: hand the accumulator stage back to the MMA warp (one arrive per epilogue warp)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&acc_empty[as]);
        }
        ptx::named_bar_sync(1, kEpiThreads);

        if (it.valid) {
          // -- phase 2: per-tile BatchNorm statistics of (acc + bias): mean and M2 over the tile's valid pixels.
          //    4 threads per column (row quarters), merged with Chan's formula.
          if (p.stats) {
            const int vh = min(p.BH, p.H - h0), vw = min(p.BW, p.W - w0);   // valid extent of this tile
            const int c = et & 63, q = et >> 6;
            const int r0 = (q * vh) >> 2, r1 = ((q + 1) * vh) >> 2;
            const int n_loc = (r1 - r0) * vw;
            float sum = 0.f, mean = 0.f, m2 = 0.f;
            if (c < CW && n_loc > 0) {
              for (int th = r0; th < r1; ++th) {
                const float* rowp = stage + (size_t)(th * p.BW) * ldst + c;
                for (int tw = 0; tw < vw; ++tw) sum += rowp[(size_t)tw * ldst];
              }
              mean = sum / (float)n_loc;
              for (int th = r0; th < r1; ++th) {
                const float* rowp = stage + (size_t)(th * p.BW) * ldst + c;
                for (int tw = 0; tw < vw; ++tw) {
                  const float d = rowp[(size_t)tw * ldst] - mean;
                  m2 = fmaf(d, d, m2);
                }
              }
            }
            if (c < CW) {
              float* rp = red + (size_t)(q * 64 + c) * 3;
              rp[0] = mean; rp[1] = m2; rp[2] = (float)n_loc;
            }
            ptx::named_bar_sync(2, kEpiThreads);
            if (q == 0 && c < CW) {
              // merge the four row quarters, then fold the tile into this CTA's running (n, mean, M2) of the column
              float* run = run_stats + (size_t)(it.nt * p.BN + cc + c) * 3;
              float na = run[0], ma = run[1], m2a = run[2];
#pragma unroll
              for (int qq = 0; qq < 4; ++qq) {
                const float* rp = red + (size_t)(qq * 64 + c) * 3;
                const float nb = rp[2];
                if (nb > 0.f) {
                  const float nn = na + nb, d = rp[0] - ma;
                  ma += d * nb / nn;
                  m2a += rp[1] + d * d * na * nb / nn;
                  na = nn;
                }
              }
              run[0] = na; run[1] = ma; run[2] = m2a;
            }
          }

          // -- phase 3: (+bias, *scale+shift, relu) -> 2x2 max/sum -> mask -> coalesced NHWC stores.
          //    A thread keeps ONE float4 channel group and strides over pixels: its affine constants live in registers.
          const int ch = n0 + g * 4;
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f), t4 = b4;
          if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
          if (p.scale) {
            s4 = __ldg(reinterpret_cast<const float4*>(p.scale + ch));
            t4 = __ldg(reinterpret_cast<const float4*>(p.shift + ch));
          }
          // fold: v = (acc + b)*s + t = acc*s + (b*s + t)
          t4 = make_float4(fmaf(b4.x, s4.x, t4.x), fmaf(b4.y, s4.y, t4.y), fmaf(b4.z, s4.z, t4.z), fmaf(b4.w, s4.w, t4.w));
          const bool affine = p.bias || p.scale;
          const int oh0 = p.reduce ? h0 / 2 : h0;
          const int ow0 = p.reduce ? w0 / 2 : w0;
          int ph = ph_start, pw = pw_start;
          for (int pix = pl; pix < oBH * oBW; pix += PS) {
            const int oh = oh0 + ph, ow = ow0 + pw;
            if (oh < Ho && ow < Wo) {
              float4 v;
              auto fetch = [&](int row) {
                float4 a = *reinterpret_cast<const float4*>(stage + (size_t)row * ldst + g * 4);
                if (affine) {
                  a.x = fmaf(a.x, s4.x, t4.x); a.y = fmaf(a.y, s4.y, t4.y); a.z = fmaf(a.z, s4.z, t4.z); a.w = fmaf(a.w, s4.w, t4.w);
                }
                if (p.relu) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
                return a;
              };
              if (p.reduce == 0) {
                v = fetch(ph * p.BW + pw);
              } else {
                const int rb = (2 * ph) * p.BW + 2 * pw;
                const float4 a = fetch(rb), b = fetch(rb + 1), c = fetch(rb + p.BW), d = fetch(rb + p.BW + 1);
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
              if (p.mask) {
                const size_t mpix = p.mask_ups ? ((size_t)(img * 2 * Ho + 2 * oh) * (2 * Wo) + 2 * ow)
                                               : ((size_t)(img * Ho + oh) * Wo + ow);
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
              for (int dy = 0; dy < rep; ++dy)
                for (int dx = 0; dx < rep; ++dx) {
                  const size_t off = ((size_t)(img * Hs + oh * rep + dy) * Ws + (ow * rep + dx)) * p.Cout + ch;
                  if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + off) = v;
                  if (p.out_hi) *reinterpret_cast<uint2*>(p.out_hi + off) = hi2;
                  if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + off) = lo2;
                }
            }
            pw += PS;
            while (pw >= oBW) { pw -= oBW; ++ph; }
          }
        }
        ptx::named_bar_sync(1, kEpiThreads);   // staging is free for the next chunk / item
      }
      as ^= 1;
    }
    if (p.stats) {
      // one (mean, M2, n) partial per CTA and channel; bias shifts the mean only
      for (int c = et; c < p.Cout; c += kEpiThreads) {
        const float* run = run_stats + (size_t)c * 3;
        p.stats[((size_t)blockIdx.x * 2 + 0) * p.Cout + c] = run[1] + (p.bias ? __ldg(p.bias + c) : 0.f);
        p.stats[((size_t)blockIdx.x * 2 + 1) * p.Cout + c] = run[2];
        if (c % p.BN == 0) p.stats_cnt[(size_t)blockIdx.x * p.tiles_n + c / p.BN] = run[0];
      }
    }
  }

  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();   // no CTA exits while a peer may still arrive on its barriers
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

static int conv_pick_bn(int Cout, int precise) {
  int bn = 16;
  for (int b : {32, 64, 128}) if (Cout % b == 0) bn = b;
  if (!precise && Cout % 256 == 0) bn = 256;
  return bn;
}

// Shape of the BatchNorm-statistics workspace of egaze_conv3x3_tc (one partial per persistent CTA):
//   stats [partials][2][Cout], stats_cnt [partials][cnt_stride] (ZERO-initialised by the caller: unused CTAs count 0);
//   the count behind channel c of partial t is stats_cnt[t*cnt_stride + c/cnt_div].
extern "C" int egaze_conv3x3_stats_shape(int Cout, int precise, int* partials, int* cnt_stride, int* cnt_div) {
  int dev = 0, sms = 0;
  EGAZE_CUDA(cudaGetDevice(&dev));
  EGAZE_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int bn = conv_pick_bn(Cout, precise);
  *partials = sms;
  *cnt_stride = Cout / bn;
  *cnt_div = bn;
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

  static int sm_count = 0;
  if (sm_count == 0) {
    int dev = 0;
    EGAZE_CUDA(cudaGetDevice(&dev));
    EGAZE_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  static int cluster_env = -1;
  if (cluster_env < 0) {
    const char* e = getenv("EGAZE_CONV_CLUSTER");
    cluster_env = e ? atoi(e) : 2;
    if (cluster_env != 1 && cluster_env != 2) cluster_env = 2;
  }

  ConvTcParams p;
  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin_p = Cin_p; p.Cout = Cout;
  p.KC = (Cin_p % 64 == 0) ? 64 : ((Cin_p % 32 == 0) ? 32 : 16);
  pick_tile(H, W, reduce != 0, &p.BH, &p.BW);
  p.nsplit = precise ? 2 : 1;
  // N tile: largest of 128/64/32/16 dividing Cout (256 only in fast mode where the rings fit).
  p.BN = conv_pick_bn(Cout, precise);
  p.tiles_h = ceil_div(H, p.BH); p.tiles_w = ceil_div(W, p.BW); p.tiles_n = Cout / p.BN;
  p.total_tiles = N * p.tiles_h * p.tiles_w;
  // clusters of 2 share every weight box through TMA multicast; the split needs 8-row-aligned halves
  int CS = cluster_env;
  if (p.total_tiles < 2 || (p.BN / 2) % 8 != 0) CS = 1;
  p.tile_groups = ceil_div(p.total_tiles, CS);
  p.num_items = p.tile_groups * p.tiles_n;
  const int row_bytes = p.KC * 2;
  int a_rows = (p.BH + 2) * p.BW;
  if (a_rows < 2 * p.BW + 128) a_rows = 2 * p.BW + 128;
  p.a_slot_bytes = ((a_rows * row_bytes + 1023) / 1024) * 1024;
  p.b_slot_bytes = ((p.BN * row_bytes + 1023) / 1024) * 1024;
  const int CW = p.BN < 64 ? p.BN : 64;
  const int stage_bytes = ((128 * (CW + 4) * 4 + 4 * 64 * 3 * 4 + (stats ? Cout * 3 * 4 : 0) + 1023) / 1024) * 1024;
  static int sa_env = -1;
  if (sa_env < 0) {
    const char* e = getenv("EGAZE_CONV_SA");
    sa_env = e ? atoi(e) : 0;
  }
  p.SA = 2;
  const int budget = 222 * 1024 - stage_bytes;
  if (sa_env >= 2 && sa_env <= kMaxSA && (budget - sa_env * p.nsplit * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes) >= 3)
    p.SA = sa_env;
  int sb = (budget - p.SA * p.nsplit * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes);
  if (sb > 6) sb = 6;
  EGAZE_CHECK_ARG(sb >= 2, "conv3x3_tc: tile does not fit shared memory");
  p.SB = sb;
  // merged hi|lo weight MMA needs the two planes back to back (no slot padding) and 2*BN accumulator columns x 2 stages
  p.merged = (precise && p.b_slot_bytes == p.BN * row_bytes && 4 * p.BN <= 512 && 2 * p.BN <= 256) ? 1 : 0;
  {
    const char* e = getenv("EGAZE_CONV_MERGED");
    if (e && atoi(e) == 0) p.merged = 0;
  }
  p.stage_off = p.SA * p.nsplit * p.a_slot_bytes + p.SB * p.nsplit * p.b_slot_bytes;
  const size_t smem = (size_t)p.stage_off + stage_bytes + 1024;  // + alignment slack
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
    uint32_t box[2] = {(uint32_t)p.KC, (uint32_t)(p.BN / CS)};   // each CTA of the cluster fetches its share of the box
    int rc = egaze_encode_tmap(&tmB_hi, w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmB_lo, precise ? w_lo : w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
  }
  int clusters = sm_count / CS;
  if (clusters > p.num_items) clusters = p.num_items;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(clusters * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  const int ksteps = p.KC / 16;
#define EGAZE_CONV_LAUNCH(NS, KS, C)                                                                                  \
  do {                                                                                                                \
    static bool attr_set = false;                                                                                     \
    if (!attr_set) {                                                                                                  \
      EGAZE_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<NS, KS, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                      224 * 1024));                                                                   \
      attr_set = true;                                                                                                \
    }                                                                                                                 \
    EGAZE_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NS, KS, C>, tmA_hi, tmA_lo, tmB_hi, tmB_lo, p));            \
  } while (0)
#define EGAZE_CONV_DISPATCH(NS, C)                                                                                    \
  do {                                                                                                                \
    if (ksteps == 4) EGAZE_CONV_LAUNCH(NS, 4, C);                                                                     \
    else if (ksteps == 2) EGAZE_CONV_LAUNCH(NS, 2, C);                                                                \
    else EGAZE_CONV_LAUNCH(NS, 1, C);                                                                                 \
  } while (0)
  if (p.nsplit == 2) {
    if (CS == 2) EGAZE_CONV_DISPATCH(2, 2);
    else EGAZE_CONV_DISPATCH(2, 1);
  } else {
    if (CS == 2) EGAZE_CONV_DISPATCH(1, 2);
    else EGAZE_CONV_DISPATCH(1, 1);
  }
#undef EGAZE_CONV_DISPATCH
#undef EGAZE_CONV_LAUNCH
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
