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
// Tiling: one work item = one spatial tile of BH x BW output pixels (BH*BW <= 128 GEMM rows) x BN output channels; K runs
// over 64/32/16-channel chunks of Cin and the nine taps.  Two ways of feeding the activation operand:
//   classic : per chunk and HORIZONTAL tap s, TMA loads one (BH+2) x BW x KC window whose origin is shifted by s-1 pixels
//             (out-of-bounds rows / columns zero-filled by the TMA unit = the conv padding); the three vertical taps read
//             that window at a row offset of r*BW rows, a whole number of 8-row swizzle atoms because BW % 8 == 0.
//   window  : (KC == 64, BW == 8) ONE (BH+2) x (BW+2) x 64 window per chunk serves all nine taps: tap (r, s) is the same
//             SWIZZLE_128B tile read through a descriptor that starts (r*(BW+2) + s) * 128 bytes further and steps its 8-row
//             groups by (BW+2) * 128 bytes.  Legal because the tensor core swizzles on absolute shared-memory address bits
//             (tools/probe_swizzle_shift.cu); the descriptor's base_offset field must stay 0.
// Weights stream through their own ring, one box per tap.  precise mode issues, per K step, A_hi x [B_hi | B_lo] as ONE
// MMA of N = 2*BN plus A_lo x B_hi; the accumulator blocks are added in the epilogue.
//
// Schedule: PERSISTENT CTAs (one per SM), launched as thread-block clusters of CS (1 or 2) whose CTAs work on adjacent
// spatial tiles of the SAME output-channel tile.
//   pair mode (default for BN >= 64): tcgen05.mma.cta_group::2 with M = 256.  Each CTA loads its own activation window and
//             only HALF of every weight box into its own shared memory (cp.async.bulk.tensor .cta_group::2, completion
//             accounted on rank 0's mbarrier); rank 0's MMA warp issues one MMA for both tiles and multicasts the commits.
//             Per CTA this cuts the shared-memory operand traffic by a third, which is what bounded the kernel before.
//   multicast mode: every CTA keeps the full weight box; each loads 1/CS of it and TMA-multicasts it to its peers.
// Two TMEM accumulator stages let the epilogue of item i overlap the MMAs of item i+1.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2..9 = epilogue.  The producer
// walks its loop nest with the WHOLE warp (uniform control flow keeps box coordinates in uniform registers) and one elected
// lane issues.  The issuer of a CTA pair is ONE lane running a lean loop (barriers by shared-window address, the taps of a
// window as two descriptor strides, one asm block of MMAs per tap): what bounds that warp is the latency -- and the fetching
// -- of its own scalar instruction stream, see the MMA warp below; the generic issuer (single CTAs, fast mode) keeps the
// warp-uniform loop with the next slot's barrier query fused into the MMA asm block.  The epilogue stages the
// accumulator tile through shared memory (64-column chunks) and drains it with compile-time-specialised store loops:
// bias / folded-BN affine / ReLU, 2x2 max or sum, ReLU-backward mask, nearest-2x replicate, packed bf16 split, coalesced
// NHWC stores -- with BatchNorm batch statistics (shifted sums) and bias-gradient column sums accumulated in the same pass.
#include "common.cuh"
#include <new>

namespace {

constexpr int kThreads = 320;      // TMA warp + MMA warp + 8 epilogue warps
constexpr int kEpiThreads = 256;
constexpr int kMaxSA = 4, kMaxSB = 9;
constexpr int kEpiScratch = 32 + 8 * 3 * 64;   // floats of epilogue scratch behind the staging tile

// Division by a launch constant as multiply-high + shift (exact for dividends < 2^31): the item decode runs once per tile in
// every epilogue warp, and four hardware-less integer divisions there cost ~100 instructions of the ~700 a tile used to take.
struct FastDiv {
  uint32_t mul, shr;   // mul == 0: divisor 1
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f = {0u, 0u};
  if (d > 1) {
    int lg = 0;
    while ((1u << lg) < (uint32_t)d) ++lg;
    const unsigned p = 31u + (unsigned)lg;
    f.mul = (uint32_t)(((1ull << p) + (uint64_t)d - 1ull) / (uint64_t)d);
    f.shr = p - 32u;
  }
  return f;
}
__device__ __forceinline__ int fdiv(int n, const FastDiv f) {
  return f.mul ? (int)(__umulhi((uint32_t)n, f.mul) >> f.shr) : n;
}

struct ConvTcParams {
  int N, H, W;         // conv output == input spatial size
  int Cin_p, Cout;     // Cin padded to a multiple of KC
  int KC;              // channels per K chunk: 64, 32 or 16
  int BH, BW, BN;      // tile
  int tiles_h, tiles_w, tiles_n;
  int total_tiles;     // N * tiles_h * tiles_w
  int tile_groups;     // ceil(total_tiles / CS)
  int num_items;       // tile_groups * tiles_n
  FastDiv fd_tn, fd_tw, fd_th;   // division by tiles_n / tiles_w / tiles_h
  int nsplit;          // planes of the WEIGHT operand: 1 = single pass, 2 = hi/lo split
  int nsplit_a;        // planes of the ACTIVATION operand: 2 = hi/lo (3 MMAs per product with nsplit = 2), 1 = hi only
                       //   (nsplit = 2: A_hi x [B_hi | B_lo], 2 MMAs per product -- the data-gradient mode)
  int in_f16;          // operands are fp16 (1) or bf16 (0) bit patterns (kind::f16 needs A and B in the same format)
  float acc_scale;     // accumulators are multiplied by this before anything else (fp16 weights are packed pre-scaled by 2^k)
  // How the hi/lo products are spread over TMEM accumulator blocks of BN columns (the epilogue adds the `nsum` blocks).
  // Back-to-back tcgen05.mma into the SAME accumulator columns serialise (+~43 cycles each, tools/probe_mma_rate.cu), so
  // consecutive MMAs always target different blocks:
  //   0  one block, every MMA into it (fast mode; small-Cout fallbacks)
  //   1  A_hi x [B_hi | B_lo] as ONE MMA of N = 2*BN into blocks 0-1, A_lo x B_hi into block 0   (previous default)
  //   2  like 1, but A_lo x B_hi goes to its own block 2 (fits TMEM for BN <= 64)
  //   3  three N = BN MMAs per K step alternating between blocks X = 0 and Y = 1:
  //      even k: hi*hi -> X, hi*lo -> Y, lo*hi -> X;  odd k: hi*lo -> Y, hi*hi -> X, lo*hi -> Y
  int acc_mode;
  int nsum;            // accumulator blocks the epilogue adds (1, 2 or 3)
  int acc_cols;        // TMEM columns of one accumulator stage (power of two >= 32)
  int pair;            // 1 = CTA-pair MMA (cta_group::2, M = 256): the two CTAs of the cluster each hold HALF of every weight box
                       //     and rank 0 issues one MMA for both spatial tiles (see the kernel header)
  int win;             // 1 = "window" mode: ONE (BH+2) x (BW+2) activation window per K chunk serves all nine taps
  int win_bo;          // window mode: fill the descriptor's base_offset field with (start >> 7) & 7
  int SA, SB;          // ring depths
  int probe;           // 1 = the round-1 MMA loop everywhere (barrier probes fused into a per-tap asm block); 0 = the lean unrolled
                       //     tap loop for window-mode CTA pairs (see the MMA warp)
  int wstat;           // 1 = weight-stationary: every weight box of the layer (9 taps x chunks <= SB) is loaded ONCE per CTA and
                       //     stays in its ring slot for all of the CTA's tiles.  For the 64 -> 64 channel layers at 224^2 the
                       //     per-tile weight stream (147 KB against a 47 KB activation window and ~1700 cycles of MMAs) is what
                       //     bounds the streaming schedule: ~110 B/clk/SM out of L2, 2.5x what the L2 delivers per SM.
  // Sub-pixel decomposition of "nearest-2x upsample -> 3x3 conv" (the four upsample-fed decoder convs, reference
  // models/model_SP.py:16-27): output pixel (2i+py, 2j+px) only ever sees a 2x2 neighbourhood of the LOW-resolution input, so the
  // layer is four 2x2-tap convolutions (one per output phase) on the low-resolution map with weights pre-summed in fp32
  // (packed [phase*4 + a*2 + b][Cout][Cin_p], layout.cu) -- 16 instead of 36 MACs per low-resolution pixel and weight.
  //   1  forward: H x W is the low-resolution input, the output is 2H x 2W; one work item per (tile, n-tile, PHASE); phase
  //      (py, px) reads window taps (py + a, px + b) and stores its 128 pixels at stride 2 from (py, px).
  //   2  data gradient: the operand is the PHASE-PLANAR output gradient [4N][H][W][C] (image phase*N + n holds dY[n, 2i+py, 2j+px]),
  //      one window per phase and K chunk, taps (2 - py - a, 2 - px - b); the output is the low-resolution H x W gradient
  //      (the 2x2 sum that is the gradient of nn.Upsample happens inside the K loop).
  int sub;
  int out_planar;      // store the (H x W, even) output phase-planar: [4N][H/2][W/2][Cout] -- what a sub == 2 launch (and the
                       //   sub-pixel weight gradient) reads
  long long planar_stride;   // elements between two phase planes of the output
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
  __nv_bfloat16* out_hi;      // optional NHWC 16-bit plane: bf16(v) or fp16(v) (out_f16)
  __nv_bfloat16* out_lo;      // optional NHWC 16-bit plane: the rounding residual v - hi in the same format
  __nv_bfloat16* out_xb;      // optional NHWC bf16(v): the copy the weight-gradient GEMM reads when out_hi / out_lo are fp16
  int out_f16;                // format of out_hi / out_lo
  float* stats;               // optional [gridDim][2][Cout] per-CTA (mean, M2) of the pre-activation (acc + bias)
  float* stats_cnt;           // [gridDim][tiles_n] pixel count behind each per-CTA partial
  float* colsum;              // optional [Cout]: += column sums of the stored values (bias gradient of the upstream conv)
  long long* prof;            // optional [gridDim][16] cycle counters per role (tools/conv_prof.py); null in production
  int ablate;                 // -DEGAZE_CONV_PROF builds only (EGAZE_CONV_ABLATE): 1 no MMAs, 2 no store loop, 4 no TMEM->smem, 8 no activation loads
};

// cycle accounting of the three roles (debug aid, enabled by egaze_conv3x3_set_prof)
// (compiled in with -DEGAZE_CONV_PROF only: the run-time checks alone cost ~40 instructions per tile in every epilogue warp)
#ifdef EGAZE_CONV_PROF
#define PROF_ON(p) ((p).prof != nullptr)
#define ABLATE(p, bit) (((p).ablate & (bit)) != 0)
#else
#define PROF_ON(p) false
#define ABLATE(p, bit) false
#endif
#define PROF_T0(p) const long long prof_t0 = PROF_ON(p) ? clock64() : 0
#define PROF_ADD(p, acc) do { if (PROF_ON(p)) (acc) += clock64() - prof_t0; } while (0)

struct Item {
  int nt, tile, img, h0, w0, phase;
  bool valid;
};

__device__ __forceinline__ Item decode_item(const ConvTcParams& p, int w, int cs, int rank) {
  Item it;
  int tg = fdiv(w, p.fd_tn);
  it.nt = w - tg * p.tiles_n;                  // n-tile fastest: consecutive items reuse the activation window in L2
  it.phase = 0;
  if (p.sub == 1) { it.phase = tg & 3; tg >>= 2; }   // then the four output phases of the same window
  it.tile = tg * cs + rank;
  it.valid = it.tile < p.total_tiles;
  const int t = it.valid ? it.tile : 0;
  const int q = fdiv(t, p.fd_tw);
  const int tw_i = t - q * p.tiles_w;
  const int im = fdiv(q, p.fd_th);
  const int th_i = q - im * p.tiles_h;
  it.img = it.valid ? im : p.N;                // img == N: every TMA box is out of bounds -> zeros
  it.h0 = th_i * p.BH;
  it.w0 = tw_i * p.BW;
  return it;
}

// Per-tile constants of the lean store loop below.
struct EpiTile {
  uint32_t st_addr;     // shared address of this thread's channel group in staging row 0
  int ldst_b;           // staging row pitch (bytes)
  int pl, PS, npix;     // this thread's first output pixel, pixels per sweep, output pixels per tile
  int obw_log;          // log2(output tile width)
  int vh, vw;           // valid output extent of this tile
  int BW;               // conv tile width = staging rows per tile row
  float4 s4, t4;        // v = acc * s4 + t4 (bias and folded BatchNorm)
  float lo_clamp;       // 0 with ReLU, -inf without
  size_t out_base;      // element offset of (img, oh0*rep, ow0*rep, first channel of this thread)
  int out_row, out_px;  // element pitch of one tile row / pixel in the output (already times the replicate factor)
  int ws_c;             // element pitch of one output image row (nearest-2x replicate)
  int planar;           // phase-planar output (ConvTcParams::out_planar)
  size_t planar_stride;
  size_t mask_base;     // same for the ReLU mask tensor
  int mask_row, mask_px;
};

__device__ __forceinline__ float4 epi_affine(float4 a, const float4& s, const float4& t, float lo) {
  a.x = fmaxf(fmaf(a.x, s.x, t.x), lo);
  a.y = fmaxf(fmaf(a.y, s.y, t.y), lo);
  a.z = fmaxf(fmaf(a.z, s.z, t.z), lo);
  a.w = fmaxf(fmaf(a.w, s.w, t.w), lo);
  return a;
}

// Phase 3 of the epilogue with every mode flag resolved at compile time and power-of-two tile widths: the accumulator
// tile is drained at ~40 instructions per float4 instead of ~220 in the flag-driven loop.  Layers with a short K loop
// (K = 9*64 .. 9*128) are bound by exactly this loop, not by the MMAs.
// STATS: also accumulate the BatchNorm sums of the raw accumulators, shifted by the per-channel reference k4:
// s1 += a - k, s2 += (a - k)^2.
// OUTK: which 16-bit planes are written -- 0 none, 1 bf16 hi + lo, 2 fp16 hi + lo, 3 fp16 hi + lo + bf16 copy, 4 bf16 hi only.
template <int RED, bool MASK, bool UPS, bool F32, int OUTK, bool STATS = false>
__device__ __forceinline__ void epi_store(const ConvTcParams& p, const EpiTile& e, float4& cs, float4 k4 = float4(),
                                          float4* s1 = nullptr, float4* s2 = nullptr) {
  const int obw_mask = (1 << e.obw_log) - 1;
  // The ReLU mask of an output pixel is fetched (read-only path) BEFORE the staged accumulators it gates are touched: issued
  // next to the shared-memory loads, the global-load latency (~600 cycles) of all pixels of a trip overlaps instead of
  // serialising inside the per-pixel code -- the masked data-gradient layers at 224^2 / 112^2 were bound by exactly that.
  auto load_mask = [&](int ph, int pw) {
    return __ldg(reinterpret_cast<const uint2*>(p.mask + e.mask_base + (size_t)(ph * e.mask_row + pw * e.mask_px)));
  };
  // everything after the staged value(s) of output pixel `pix` are in registers
  auto finish = [&](int ph, int pw, float4 v, const uint2 mk) {
    if (MASK) {
      if (!pos16(mk.x & 0xffffu)) v.x = 0.f;
      if (!pos16(mk.x >> 16)) v.y = 0.f;
      if (!pos16(mk.y & 0xffffu)) v.z = 0.f;
      if (!pos16(mk.y >> 16)) v.w = 0.f;
    }
    if (MASK) { cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w; }   // column sums = bias gradient (masked dgrad modes only)
    uint2 hi2, lo2, xb2;
    if (OUTK == 1) split_bf16x4(v, hi2, lo2);
    if (OUTK == 2 || OUTK == 3) split_f16x4(v, hi2, lo2);
    if (OUTK == 3) xb2 = pack_bf16x4(v);
    if (OUTK == 4) hi2 = pack_bf16x4(v);
    size_t off;
    if (!UPS && e.planar)
      off = e.out_base + (size_t)((ph & 1) * 2 + (pw & 1)) * e.planar_stride + (size_t)((ph >> 1) * e.out_row + (pw >> 1) * e.out_px);
    else
      off = e.out_base + (size_t)(ph * e.out_row + pw * e.out_px);
#pragma unroll
    for (int dy = 0; dy < (UPS ? 2 : 1); ++dy)
#pragma unroll
      for (int dx = 0; dx < (UPS ? 2 : 1); ++dx) {
        const size_t o = off + (size_t)(dy * e.ws_c + dx * p.Cout);
        if (F32) *reinterpret_cast<float4*>(p.out_f32 + o) = v;
        if (OUTK != 0) *reinterpret_cast<uint2*>(p.out_hi + o) = hi2;
        if (OUTK >= 1 && OUTK <= 3) *reinterpret_cast<uint2*>(p.out_lo + o) = lo2;
        if (OUTK == 3) *reinterpret_cast<uint2*>(p.out_xb + o) = xb2;
      }
  };
  if (RED == 0) {
    // four pixels per trip: the four staged float4 (and masks) are loaded up front so their latency overlaps
    constexpr int U = 4;
    for (int pix0 = e.pl; pix0 < e.npix; pix0 += U * e.PS) {
      float4 a[U];
      uint2 mk[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        const int ph = pix >> e.obw_log, pw = pix & obw_mask;
        ok[u] = pix < e.npix && ph < e.vh && pw < e.vw;
        mk[u] = make_uint2(0u, 0u);
        if (MASK && ok[u]) mk[u] = load_mask(ph, pw);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        a[u] = pix < e.npix ? ptx::lds128(e.st_addr + (uint32_t)(pix * e.ldst_b)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        const int ph = pix >> e.obw_log, pw = pix & obw_mask;
        if (!ok[u]) continue;
        if (STATS) {
          const float dx = a[u].x - k4.x, dy = a[u].y - k4.y, dz = a[u].z - k4.z, dw = a[u].w - k4.w;
          s1->x += dx; s1->y += dy; s1->z += dz; s1->w += dw;
          s2->x = fmaf(dx, dx, s2->x); s2->y = fmaf(dy, dy, s2->y); s2->z = fmaf(dz, dz, s2->z); s2->w = fmaf(dw, dw, s2->w);
        }
        finish(ph, pw, epi_affine(a[u], e.s4, e.t4, e.lo_clamp), mk[u]);
      }
    }
  } else {
    // two output pixels (eight staged float4 + two masks) per trip
    constexpr int U = 2;
    for (int pix0 = e.pl; pix0 < e.npix; pix0 += U * e.PS) {
      float4 q[U][4];
      uint2 mk[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        const int ph = pix >> e.obw_log, pw = pix & obw_mask;
        ok[u] = pix < e.npix && ph < e.vh && pw < e.vw;
        mk[u] = make_uint2(0u, 0u);
        if (MASK && ok[u]) mk[u] = load_mask(ph, pw);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        const int ph = pix >> e.obw_log, pw = pix & obw_mask;
        if (pix < e.npix) {
          const uint32_t rb = e.st_addr + (uint32_t)(((2 * ph) * e.BW + 2 * pw) * e.ldst_b);
          q[u][0] = ptx::lds128(rb);
          q[u][1] = ptx::lds128(rb + (uint32_t)e.ldst_b);
          q[u][2] = ptx::lds128(rb + (uint32_t)(e.BW * e.ldst_b));
          q[u][3] = ptx::lds128(rb + (uint32_t)((e.BW + 1) * e.ldst_b));
        } else {
          q[u][0] = q[u][1] = q[u][2] = q[u][3] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int pix = pix0 + u * e.PS;
        const int ph = pix >> e.obw_log, pw = pix & obw_mask;
        if (!ok[u]) continue;
        const float4 a = epi_affine(q[u][0], e.s4, e.t4, e.lo_clamp);
        const float4 b = epi_affine(q[u][1], e.s4, e.t4, e.lo_clamp);
        const float4 c = epi_affine(q[u][2], e.s4, e.t4, e.lo_clamp);
        const float4 d = epi_affine(q[u][3], e.s4, e.t4, e.lo_clamp);
        float4 v;
        if (RED == 1) {
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
        finish(ph, pw, v, mk[u]);
      }
    }
  }
}

// One tap of the precise-mode "merged" schedule -- per K step A_hi x [B_hi | B_lo] (N = 2*BN) then A_lo x B_hi (N = BN) --
// with two barrier probes fused into the same asm block: mbarrier.test_wait on the NEXT weight slot and the NEXT
// activation slot is issued BEFORE the MMAs and its predicate is read AFTER them.  A barrier query costs ~100 cycles
// even when the phase completed long ago and the tensor pipe's instruction queue is shallow, so a query sitting between
// two taps is tensor-pipe idle time; issued here its latency hides behind the (blocking) MMA issue.
// Returns bit 0: next B slot has landed, bit 1: next A slot has landed.
#define EGAZE_MMA_PAIR(CG, OFF)                                                            \
  "add.s64 ah, %2, " #OFF ";\n\t"                                                          \
  "add.s64 al, %3, " #OFF ";\n\t"                                                          \
  "add.s64 bb, %4, " #OFF ";\n\t"                                                          \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], ah, bb, %5, pt;\n\t"                     \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%12], al, bb, %6, pt;\n\t"
// the same without the A_lo x B_hi product (activation operand given as one plane: 2 MMAs per product)
#define EGAZE_MMA_ONE(CG, OFF)                                                             \
  "add.s64 ah, %2, " #OFF ";\n\t"                                                          \
  "add.s64 bb, %4, " #OFF ";\n\t"                                                          \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], ah, bb, %5, pt;\n\t"
#define EGAZE_TAP_HEAD(CG)                                                                 \
  "{\n\t"                                                                                  \
  ".reg .pred pb, pa, pacc, pt;\n\t"                                                       \
  ".reg .b64 ah, al, bb;\n\t"                                                              \
  ".reg .b32 rb, ra;\n\t"                                                                  \
  "mbarrier.test_wait.parity.shared::cta.b64 pb, [%8], %9;\n\t"                            \
  "mbarrier.test_wait.parity.shared::cta.b64 pa, [%10], %11;\n\t"                          \
  "setp.ne.b32 pacc, %7, 0;\n\t"                                                           \
  "setp.eq.b32 pt, %7, %7;\n\t"                                                            \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], %2, %4, %5, pacc;\n\t"                   \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%12], %3, %4, %6, pt;\n\t"
#define EGAZE_TAP_HEAD_ONE(CG)                                                             \
  "{\n\t"                                                                                  \
  ".reg .pred pb, pa, pacc, pt;\n\t"                                                       \
  ".reg .b64 ah, al, bb;\n\t"                                                              \
  ".reg .b32 rb, ra;\n\t"                                                                  \
  "mbarrier.test_wait.parity.shared::cta.b64 pb, [%8], %9;\n\t"                            \
  "mbarrier.test_wait.parity.shared::cta.b64 pa, [%10], %11;\n\t"                          \
  "setp.ne.b32 pacc, %7, 0;\n\t"                                                           \
  "setp.eq.b32 pt, %7, %7;\n\t"                                                            \
  "tcgen05.mma.cta_group::" #CG ".kind::f16 [%1], %2, %4, %5, pacc;\n\t"
#define EGAZE_TAP_TAIL                                                                     \
  "selp.u32 rb, 1, 0, pb;\n\t"                                                             \
  "selp.u32 ra, 2, 0, pa;\n\t"                                                             \
  "or.b32 %0, rb, ra;\n\t"                                                                 \
  "}\n"
#define EGAZE_TAP_OPERANDS                                                                                            \
  : "=r"(flags)                                                                                                       \
  : "r"(d_tmem), "l"(ad_hi), "l"(ad_lo), "l"(bd), "r"(idesc2), "r"(idesc), "r"(accumulate), "r"(bar_b), "r"(par_b),   \
    "r"(bar_a), "r"(par_a), "r"(d_tmem2)                                                                              \
  : "memory"
#define EGAZE_TAP_ASM(CG)                                                                                             \
  do {                                                                                                                \
    if (!LOHI) {                                                                                                      \
      if (KSTEPS == 4) {                                                                                              \
        asm volatile(EGAZE_TAP_HEAD_ONE(CG) EGAZE_MMA_ONE(CG, 2) EGAZE_MMA_ONE(CG, 4) EGAZE_MMA_ONE(CG, 6)            \
                         EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                                                          \
      } else if (KSTEPS == 2) {                                                                                       \
        asm volatile(EGAZE_TAP_HEAD_ONE(CG) EGAZE_MMA_ONE(CG, 2) EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                  \
      } else {                                                                                                        \
        asm volatile(EGAZE_TAP_HEAD_ONE(CG) EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                                       \
      }                                                                                                               \
    } else if (KSTEPS == 4) {                                                                                         \
      asm volatile(EGAZE_TAP_HEAD(CG) EGAZE_MMA_PAIR(CG, 2) EGAZE_MMA_PAIR(CG, 4) EGAZE_MMA_PAIR(CG, 6)               \
                       EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                                                            \
    } else if (KSTEPS == 2) {                                                                                         \
      asm volatile(EGAZE_TAP_HEAD(CG) EGAZE_MMA_PAIR(CG, 2) EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                       \
    } else {                                                                                                          \
      asm volatile(EGAZE_TAP_HEAD(CG) EGAZE_TAP_TAIL EGAZE_TAP_OPERANDS);                                             \
    }                                                                                                                 \
  } while (0)
// d_tmem: accumulator of the N = 2*BN MMA; d_tmem2: accumulator of the N = BN MMA (the same block, or shifted by BN/2
// columns in CTA-pair mode).  PAIR: cta_group::2 (M = 256 across the two CTAs of the cluster).
template <int KSTEPS, bool PAIR, bool LOHI>
__device__ __forceinline__ uint32_t tap_merged_probe(uint32_t d_tmem, uint32_t d_tmem2, uint64_t ad_hi, uint64_t ad_lo,
                                                     uint64_t bd, uint32_t idesc2, uint32_t idesc, uint32_t accumulate,
                                                     uint32_t bar_b, uint32_t par_b, uint32_t bar_a, uint32_t par_a) {
  uint32_t flags;
  if (PAIR) EGAZE_TAP_ASM(2);
  else EGAZE_TAP_ASM(1);
  return flags;
}
#undef EGAZE_TAP_ASM
#undef EGAZE_TAP_HEAD
#undef EGAZE_TAP_HEAD_ONE
#undef EGAZE_TAP_TAIL
#undef EGAZE_TAP_OPERANDS
#undef EGAZE_MMA_PAIR
#undef EGAZE_MMA_ONE

// ---- single-lane helpers of the lean MMA issuer: barriers by shared-window address, a whole tap in one asm block ----------
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar_addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(bar_addr), "r"(parity)
      : "memory");
  if (!ok) ptx::mbar_wait_slow(bar_addr, parity);
}
__device__ __forceinline__ void commit_2sm_addr(uint32_t bar_addr, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
               "h"(cta_mask)
               : "memory");
}
// One tap on a CTA pair: per K step A_hi x [B_hi | B_lo] into d1 (N = 2*BN) and, with LOHI, A_lo x B_hi into d2 (N = BN).
// Descriptors travel as 32-bit halves: only the low word (the start-address field) differs between taps and K steps.
#define EGAZE_LT_K(OFF)                                                                     \
  "add.s64 ah, a0, " #OFF ";\n\t"                                                           \
  "add.s64 bb, b0, " #OFF ";\n\t"                                                           \
  "tcgen05.mma.cta_group::2.kind::f16 [%0], ah, bb, %7, pt;\n\t"
#define EGAZE_LT_K2(OFF)                                                                    \
  EGAZE_LT_K(OFF)                                                                           \
  "add.s64 al, a1, " #OFF ";\n\t"                                                           \
  "tcgen05.mma.cta_group::2.kind::f16 [%1], al, bb, %8, pt;\n\t"
#define EGAZE_LT_HEAD                                                                       \
  "{\n\t"                                                                                   \
  ".reg .pred pacc, pt;\n\t"                                                                \
  ".reg .b64 a0, a1, b0, ah, al, bb;\n\t"                                                   \
  "setp.ne.b32 pacc, %9, 0;\n\t"                                                            \
  "setp.eq.b32 pt, %9, %9;\n\t"                                                             \
  "mov.b64 a0, {%2, %3};\n\t"                                                               \
  "mov.b64 a1, {%4, %3};\n\t"                                                               \
  "mov.b64 b0, {%5, %6};\n\t"                                                               \
  "tcgen05.mma.cta_group::2.kind::f16 [%0], a0, b0, %7, pacc;\n\t"
#define EGAZE_LT_HEAD2 EGAZE_LT_HEAD "tcgen05.mma.cta_group::2.kind::f16 [%1], a1, b0, %8, pt;\n\t"
#define EGAZE_LT_OPERANDS                                                                                             \
  :: "r"(d1), "r"(d2), "r"(a_lo), "r"(a_hi), "r"(a2_lo), "r"(b_lo), "r"(b_hi), "r"(idesc2), "r"(idesc), "r"(accumulate) \
  : "memory"
template <int KSTEPS, bool LOHI>
__device__ __forceinline__ void lean_tap_2sm(uint32_t d1, uint32_t d2, uint32_t a_lo, uint32_t a_hi, uint32_t a2_lo, uint32_t b_lo,
                                             uint32_t b_hi, uint32_t idesc2, uint32_t idesc, uint32_t accumulate) {
  if (LOHI) {
    if (KSTEPS == 4) asm volatile(EGAZE_LT_HEAD2 EGAZE_LT_K2(2) EGAZE_LT_K2(4) EGAZE_LT_K2(6) "}\n" EGAZE_LT_OPERANDS);
    else if (KSTEPS == 2) asm volatile(EGAZE_LT_HEAD2 EGAZE_LT_K2(2) "}\n" EGAZE_LT_OPERANDS);
    else asm volatile(EGAZE_LT_HEAD2 "}\n" EGAZE_LT_OPERANDS);
  } else {
    if (KSTEPS == 4) asm volatile(EGAZE_LT_HEAD EGAZE_LT_K(2) EGAZE_LT_K(4) EGAZE_LT_K(6) "}\n" EGAZE_LT_OPERANDS);
    else if (KSTEPS == 2) asm volatile(EGAZE_LT_HEAD EGAZE_LT_K(2) "}\n" EGAZE_LT_OPERANDS);
    else asm volatile(EGAZE_LT_HEAD "}\n" EGAZE_LT_OPERANDS);
  }
}
#undef EGAZE_LT_K
#undef EGAZE_LT_K2
#undef EGAZE_LT_HEAD
#undef EGAZE_LT_HEAD2
#undef EGAZE_LT_OPERANDS

// NSA / NSPLIT: planes of the activation / weight operand: (2, 2) 3 MMAs per product, (1, 2) 2 MMAs, (1, 1) one.
template <int NSA, int NSPLIT, int KSTEPS, int CS>
__global__ void __maxnreg__(128)
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

  uint8_t* a_ring = smem;                                                 // SA slots x NSA planes
  uint8_t* b_ring = smem + (size_t)p.SA * NSA * p.a_slot_bytes;            // SB slots x NSPLIT planes
  float* stage = reinterpret_cast<float*>(smem + p.stage_off);             // [128][CW+4] + scratch
  __shared__ uint64_t a_full[kMaxSA], a_empty[kMaxSA], b_full[kMaxSB], b_empty[kMaxSB], acc_full[2], acc_empty[2];
  __shared__ uint32_t tmem_base_smem;

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.SA; ++i) { ptx::mbar_init(&a_full[i], 1); ptx::mbar_init(&a_empty[i], 1); }
    // pair mode: only rank 0 issues MMAs / commits (multicast to both CTAs), and rank 0's accumulator-free barrier collects the
    // epilogue warps of BOTH CTAs
    for (int i = 0; i < p.SB; ++i) { ptx::mbar_init(&b_full[i], 1); ptx::mbar_init(&b_empty[i], p.pair ? 1 : CS); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&acc_full[i], 1);
      ptx::mbar_init(&acc_empty[i], (p.pair ? 2 : 1) * (kEpiThreads / 32));
    }
    ptx::fence_barrier_init();
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA_hi);
    ptx::prefetch_tmap(&tmB_hi);
    if (NSA == 2) ptx::prefetch_tmap(&tmA_lo);
    if (NSPLIT == 2) ptx::prefetch_tmap(&tmB_lo);
  }
  const uint32_t acc_cols = (uint32_t)p.acc_cols;                // columns of one accumulator stage
  const uint32_t tmem_cols = 2 * acc_cols;                       // power of two in [64, 512]
  const bool pair = CS == 2 && p.pair;
  if (warp == 1) {
    if (pair) { ptx::tmem_alloc2(&tmem_base_smem, tmem_cols); ptx::tmem_relinquish2(); }
    else { ptx::tmem_alloc(&tmem_base_smem, tmem_cols); ptx::tmem_relinquish(); }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();   // peers' barriers are initialised before any multicast / remote arrive
  ptx::tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const int chunks = p.Cin_p / p.KC;
  const int row_bytes = p.KC * 2;
  const int AW = p.win ? p.BW + 2 : p.BW;                        // pixel columns of one activation window
  const uint32_t a_box_bytes = (uint32_t)((p.BH + 2) * AW * row_bytes);
  // windows per K chunk: one per horizontal tap, one in all (window mode), or one per phase plane (sub-pixel data gradient)
  const int a_loads = p.sub == 2 ? 4 : (p.win ? 1 : 3);
  const int a_taps = p.sub ? 4 : (p.win ? 9 : 3);                // weight boxes consumed per window
  const uint32_t b_box_bytes = (uint32_t)(p.BN * row_bytes);
  const int b_rows_cta = p.BN / CS;                              // weight rows this CTA fetches (and multicasts / keeps, pair mode)

  if (warp == 0) {
    // ================================ TMA producer ================================
    // Like the MMA warp: the whole warp walks the loop nest (warp-uniform control flow keeps the tensor-map pointers, box
    // coordinates and barrier addresses in uniform registers) and one elected lane waits / issues.  Under `if (lane == 0)`
    // every UTMALDG operand went through R2UR and the producer needed ~500 cycles per tap -- more than a tap's MMAs take
    // in the 64-channel layers.
    {
      const bool leader = ptx::elect_one();
      int sa = 0, sb = 0;
      long long prof_c[2] = {0, 0};
      const long long prof_start = PROF_ON(p) ? clock64() : 0;
      uint32_t a_par = 1, b_par = 1;  // a fresh mbarrier passes a parity-1 wait: the first lap never blocks
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const Item it = decode_item(p, w, CS, rank);
        const int n0 = it.nt * p.BN;
        const bool load_b = !(p.wstat && w != cluster_id);   // weight-stationary: the boxes of the first item stay
        for (int kc = 0; kc < chunks; ++kc) {
          for (int al = 0; al < a_loads; ++al) {
            uint8_t* a_dst = a_ring + (size_t)sa * NSA * p.a_slot_bytes;
            const int wx = p.win ? it.w0 - 1 : it.w0 - 1 + al;
            // sub-pixel data gradient: window `al` comes from phase plane al (images al*N ..); an invalid tile reads past the end
            const int img_c = p.sub == 2 ? (it.valid ? al * p.N + it.img : 4 * p.N) : it.img;
            if (leader) {
              { PROF_T0(p); ptx::mbar_wait(&a_empty[sa], a_par); PROF_ADD(p, prof_c[0]); }
              if (ABLATE(p, 8)) {
                if (!pair || rank == 0) ptx::mbar_arrive(&a_full[sa]);
              } else if (pair) {
                // both CTAs' windows are accounted on rank 0's barrier (its MMA thread consumes both)
                if (rank == 0) ptx::mbar_arrive_expect_tx(&a_full[sa], 2 * a_box_bytes * NSA);
                ptx::tma_load_4d_2sm(a_dst, &tmA_hi, &a_full[sa], kc * p.KC, wx, it.h0 - 1, img_c);
                if (NSA == 2)
                  ptx::tma_load_4d_2sm(a_dst + p.a_slot_bytes, &tmA_lo, &a_full[sa], kc * p.KC, wx, it.h0 - 1, img_c);
              } else {
                ptx::mbar_arrive_expect_tx(&a_full[sa], a_box_bytes * NSA);
                ptx::tma_load_4d(a_dst, &tmA_hi, &a_full[sa], kc * p.KC, wx, it.h0 - 1, img_c);
                if (NSA == 2)
                  ptx::tma_load_4d(a_dst + p.a_slot_bytes, &tmA_lo, &a_full[sa], kc * p.KC, wx, it.h0 - 1, img_c);
              }
            }
            if (++sa == p.SA) { sa = 0; a_par ^= 1; }
            for (int t = 0; t < a_taps; ++t) {
              const int s = p.win ? t / 3 : al, r = p.win ? t - 3 * s : t;
              // weight box of this step: tap (r, s), or sub-pixel box [phase*4 + t] (phase = the item's / the window's)
              const int wbox = p.sub ? (p.sub == 1 ? it.phase : al) * 4 + t : r * 3 + s;
              const int brow = wbox * p.Cout + n0 + rank * b_rows_cta;
              // pair mode: this CTA keeps only ITS half of the box (rows rank*BN/2 ..), hi plane then lo plane back to back;
              // multicast mode: it fetches its half and multicasts it into every CTA's full-size slot
              uint8_t* b_dst = b_ring + (size_t)sb * NSPLIT * p.b_slot_bytes + (pair ? 0 : (size_t)rank * b_rows_cta * row_bytes);
              if (leader && load_b) {
                { PROF_T0(p); ptx::mbar_wait(&b_empty[sb], b_par); PROF_ADD(p, prof_c[1]); }   // the slot has been drained
                if (pair) {
                  if (rank == 0) ptx::mbar_arrive_expect_tx(&b_full[sb], b_box_bytes * NSPLIT);
                  ptx::tma_load_2d_2sm(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow);
                  if (NSPLIT == 2) ptx::tma_load_2d_2sm(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow);
                } else {
                  ptx::mbar_arrive_expect_tx(&b_full[sb], b_box_bytes * NSPLIT);
                  if (CS > 1) {
                    ptx::tma_load_2d_mc(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow, kMask);
                    if (NSPLIT == 2) ptx::tma_load_2d_mc(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow, kMask);
                  } else {
                    ptx::tma_load_2d(b_dst, &tmB_hi, &b_full[sb], kc * p.KC, brow);
                    if (NSPLIT == 2) ptx::tma_load_2d(b_dst + p.b_slot_bytes, &tmB_lo, &b_full[sb], kc * p.KC, brow);
                  }
                }
              }
              if (++sb == p.SB) { sb = 0; b_par ^= 1; }
            }
            __syncwarp();
          }
        }
      }
      if (PROF_ON(p) && leader) {
        long long* o = p.prof + (size_t)blockIdx.x * 16;
        o[0] = prof_c[0]; o[1] = prof_c[1]; o[2] = clock64() - prof_start;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // The WHOLE warp walks the pipeline (barrier waits, descriptor arithmetic) so that control flow stays warp-uniform
    // and ptxas keeps the descriptors in uniform registers; one elected lane issues the tcgen05 instructions.  (With the
    // loop nest under `if (lane == 0)` every operand of every MMA went through an R2UR move and the single thread
    // needed ~100 cycles per tcgen05.mma -- more than an N <= 128 MMA takes to execute.)
    {
      // CTA-pair mode: rank 0 issues for both CTAs (M = 256); rank 1's MMA warp has nothing to do
      const bool leader = ptx::elect_one() && !(pair && rank != 0);
      const int mma_m = pair ? 256 : 128;
      const uint32_t idesc = ptx::make_idesc_16(mma_m, p.BN, 0, 0, p.in_f16);
      const uint32_t idesc2 = ptx::make_idesc_16(mma_m, 2 * p.BN, 0, 0, p.in_f16);
      const uint32_t sbo = 8u * (uint32_t)row_bytes;
      // window mode: the 8-pixel row segment of tile row th starts (BW+2) window rows after the one of th-1
      const uint32_t sbo_a = p.win ? (uint32_t)AW * (uint32_t)row_bytes : sbo;
      // Descriptors differ only in their 14-bit start-address field (smem address >> 4; smem < 256 KB so the
      // field never carries): build the static part once, then a descriptor is one 64-bit add.
      const uint64_t desc_static = ptx::make_smem_desc(0, 16, sbo, (uint32_t)row_bytes);
      const uint64_t a_ring_desc = ptx::make_smem_desc(0, 16, sbo_a, (uint32_t)row_bytes) + (uint64_t)(ptx::smem_u32(a_ring) >> 4);
      const uint64_t b_ring_desc = desc_static + (uint64_t)(ptx::smem_u32(b_ring) >> 4);
      const uint32_t a_slot16 = (uint32_t)(NSA * p.a_slot_bytes) >> 4, a_plane16 = (uint32_t)p.a_slot_bytes >> 4;
      const uint32_t b_slot16 = (uint32_t)(NSPLIT * p.b_slot_bytes) >> 4, b_plane16 = (uint32_t)p.b_slot_bytes >> 4;
      const uint32_t r_step16 = (uint32_t)(AW * row_bytes) >> 4;   // one tile row down inside the window
      const uint32_t s_step16 = (uint32_t)row_bytes >> 4;          // one pixel to the right (window mode)
      const int acc_mode = NSPLIT == 2 ? p.acc_mode : 0;
      const bool lean = pair && NSPLIT == 2 && !p.probe;
      const uint32_t bn = (uint32_t)p.BN;
      if (lean) {
        // Lean issuer (CTA pairs): ONE lane runs the whole pipeline.  What bounds this warp is the latency of
        // its own scalar instruction stream (and fetching it: the SMSP's ~6 KB L0 instruction cache is shared with two epilogue
        // warps; ncu shows `no_instruction` as the top stall): with all work ablated the generic loop below still needed
        // ~560 cycles per tap (-DEGAZE_CONV_PROF, EGAZE_CONV_ABLATE=15) against 256-384 cycles of MMAs per tap in the 64-channel
        // layers, so the tensor pipe ran dry between taps.  Hence: no per-tap divergent regions, barriers addressed by plain
        // shared-window offsets, launch constants in registers, the taps of a window walked as `no` x `ni` steps of two descriptor
        // strides (window: 3 columns x 3 rows; sub-pixel: 2 x 2 from the phase's corner, backwards for the data gradient), one
        // asm block per tap, no barrier probes (a plain wait on the next weight slot is covered by the MMAs already queued).
        if (leader) {
          const uint32_t a_full0 = ptx::smem_u32(&a_full[0]), a_empty0 = ptx::smem_u32(&a_empty[0]);
          const uint32_t b_full0 = ptx::smem_u32(&b_full[0]), b_empty0 = ptx::smem_u32(&b_empty[0]);
          const uint32_t acc_full0 = ptx::smem_u32(&acc_full[0]), acc_empty0 = ptx::smem_u32(&acc_empty[0]);
          const uint32_t a_ring_lo = (uint32_t)a_ring_desc, a_hi = (uint32_t)(a_ring_desc >> 32);
          const uint32_t b_ring_lo = (uint32_t)b_ring_desc, b_hi = (uint32_t)(b_ring_desc >> 32);
          const int SA = p.SA, SB = p.SB, sub = p.sub, wstat = p.wstat, n_items = p.num_items;
          // classic mode (one window per horizontal tap): its three vertical taps are one row of the walk
          const int no = !p.win ? 1 : (sub ? 2 : 3), ni = sub ? 2 : 3;
          // strides in 16-byte units; "negative" steps as two's complement (only the low 14 bits of the field matter)
          const uint32_t so = sub == 0 ? s_step16 : (sub == 1 ? r_step16 : 0u - r_step16);
          const uint32_t si = sub == 0 ? r_step16 : (sub == 1 ? s_step16 : 0u - s_step16);
          int sa = 0, sb = 0, as = 0;
          uint32_t a_par = 0, b_par = 0, acc_par = 3u;   // bit `as`: parity to wait for on accumulator stage `as`
          uint32_t b_lo = b_ring_lo;
          long long prof_c[4] = {0, 0, 0, 0};
          const long long prof_start = PROF_ON(p) ? clock64() : 0;
#pragma unroll 1
          for (int w = cluster_id; w < n_items; w += num_clusters) {
            const bool b_resident = wstat && w != cluster_id;   // weight-stationary: boxes already in their slots, never released
            const int phase = sub == 1 ? fdiv(w, p.fd_tn) & 3 : 0;
            { PROF_T0(p); mbar_wait_addr(acc_empty0 + (uint32_t)as * 8u, (acc_par >> as) & 1u); PROF_ADD(p, prof_c[0]); }
            acc_par ^= 1u << as;
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
            uint32_t accumulate = 0;
#pragma unroll 1
            for (int kc = 0; kc < chunks; ++kc) {
#pragma unroll 1
              for (int al = 0; al < a_loads; ++al) {
                { PROF_T0(p); mbar_wait_addr(a_full0 + (uint32_t)sa * 8u, a_par); PROF_ADD(p, prof_c[1]); }
                ptx::tc_fence_after();
                uint32_t base = 0u;
                if (sub == 1) base = (uint32_t)(phase >> 1) * r_step16 + (uint32_t)(phase & 1) * s_step16;
                else if (sub == 2) base = (uint32_t)(2 - (al >> 1)) * r_step16 + (uint32_t)(2 - (al & 1)) * s_step16;
                uint32_t a_lo = a_ring_lo + (uint32_t)sa * a_slot16 + base;
#pragma unroll 1
                for (int o = 0; o < no; ++o) {
                  uint32_t a_in = a_lo;
#pragma unroll 1
                  for (int i = 0; i < ni; ++i) {
                    if (!b_resident) {
                      PROF_T0(p); mbar_wait_addr(b_full0 + (uint32_t)sb * 8u, b_par); PROF_ADD(p, prof_c[2]);
                      ptx::tc_fence_after();
                    }
                    const long long prof_i0 = PROF_ON(p) ? clock64() : 0;
                    if (!ABLATE(p, 1))
                      lean_tap_2sm<KSTEPS, NSA == 2>(d_tmem, d_tmem + (bn >> 1), a_in, a_hi, a_in + a_plane16, b_lo, b_hi, idesc2, idesc,
                                                     accumulate);
                    if (PROF_ON(p)) prof_c[3] += clock64() - prof_i0;
                    if (!wstat) commit_2sm_addr(b_empty0 + (uint32_t)sb * 8u, kMask);
                    accumulate = 1;
                    a_in += si;
                    b_lo += b_slot16;
                    if (++sb == SB) { sb = 0; b_par ^= 1; b_lo = b_ring_lo; }
                  }
                  a_lo += so;
                }
                commit_2sm_addr(a_empty0 + (uint32_t)sa * 8u, kMask);
                if (++sa == SA) { sa = 0; a_par ^= 1; }
              }
            }
            commit_2sm_addr(acc_full0 + (uint32_t)as * 8u, kMask);
            as ^= 1;
          }
          if (PROF_ON(p)) {
            long long* o = p.prof + (size_t)blockIdx.x * 16;
            o[3] = prof_c[0]; o[4] = prof_c[1]; o[5] = prof_c[2]; o[6] = clock64() - prof_start; o[15] = prof_c[3];
          }
        }
        __syncwarp();
      } else {
      int sa = 0, sb = 0, as = 0;
      long long prof_c[4] = {0, 0, 0, 0};
      const long long prof_start = PROF_ON(p) ? clock64() : 0;
      uint32_t a_par = 0, b_par = 0, acc_par[2] = {1, 1};
      // Only the elected lane waits on barriers and issues; the other lanes just keep the loop nest warp-uniform.
      bool a_ready = false, a_next_ready = false, b_ready = false;
      for (int w = cluster_id; w < p.num_items; w += num_clusters) {
        const bool b_resident = p.wstat && w != cluster_id;   // weight-stationary: boxes already in their slots, never released
        const int phase = p.sub == 1 ? fdiv(w, p.fd_tn) & 3 : 0;
        if (leader) { PROF_T0(p); ptx::mbar_wait(&acc_empty[as], acc_par[as]); PROF_ADD(p, prof_c[0]); }   // epilogue has drained this accumulator stage
        acc_par[as] ^= 1;
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_cols;
        uint32_t accumulate = 0;
        for (int kc = 0; kc < chunks; ++kc) {
          for (int al = 0; al < a_loads; ++al) {
            if (leader && !a_ready) { PROF_T0(p); ptx::mbar_wait(&a_full[sa], a_par); PROF_ADD(p, prof_c[1]); }
            ptx::tc_fence_after();
            const uint64_t a_desc0 = a_ring_desc + (uint64_t)((uint32_t)sa * a_slot16);
            const int nsa = sa + 1 == p.SA ? 0 : sa + 1;
            const uint32_t na_bar = ptx::smem_u32(&a_full[nsa]), na_par = nsa == 0 ? a_par ^ 1 : a_par;
#pragma unroll 1
            for (int t = 0; t < a_taps; ++t) {
              uint64_t ad;
              if (p.sub) {
                // window position of 2x2 tap (a, b) = (t >> 1, t & 1): forward phase (py, px) reads (py + a, px + b); the data
                // gradient reads phase plane `al` at (2 - py - a, 2 - px - b)
                const int a = t >> 1, b = t & 1;
                const int r = p.sub == 1 ? (phase >> 1) + a : 2 - (al >> 1) - a;
                const int s = p.sub == 1 ? (phase & 1) + b : 2 - (al & 1) - b;
                ad = a_desc0 + (uint64_t)((uint32_t)r * r_step16 + (uint32_t)s * s_step16);
              } else if (p.win) {
                const int s = t / 3, r = t - 3 * s;
                ad = a_desc0 + (uint64_t)((uint32_t)r * r_step16 + (uint32_t)s * s_step16);
              } else {
                ad = a_desc0 + (uint64_t)((uint32_t)t * r_step16);
              }
              const uint64_t bd = b_ring_desc + (uint64_t)((uint32_t)sb * b_slot16);
              const int nsb = sb + 1 == p.SB ? 0 : sb + 1;
              const uint32_t nb_bar = ptx::smem_u32(&b_full[nsb]), nb_par = nsb == 0 ? b_par ^ 1 : b_par;
              // products: (hi,hi) [, (hi,lo), (lo,hi)].  merged: the hi and lo weight planes are contiguous in the slot, so
              // A_hi x [B_hi | B_lo] is ONE MMA of N = 2*BN; its two column halves are added in the epilogue.
              if (leader) {
                if (!b_ready && !b_resident) { PROF_T0(p); ptx::mbar_wait(&b_full[sb], b_par); PROF_ADD(p, prof_c[2]); }
                ptx::tc_fence_after();
                b_ready = false;
                const long long prof_i0 = PROF_ON(p) ? clock64() : 0;
                if (pair) {
                  // merged schedule on the CTA pair.  N = 2*BN: each CTA contributes its BN rows [hi half | lo half], so the
                  // accumulator columns are [hh(c < BN/2) | hl(c < BN/2) | hh(c >= BN/2) | hl(c >= BN/2)]; the N = BN MMA
                  // (A_lo x B_hi, BN/2 rows per CTA) lands BN/2 columns in, on top of columns of the SAME channels.
                  if (ABLATE(p, 1)) {
                  } else if (NSPLIT == 2) {
                    const uint32_t fl = tap_merged_probe<KSTEPS, true, NSA == 2>(d_tmem, d_tmem + (bn >> 1), ad, ad + a_plane16, bd, idesc2,
                                                                       idesc, accumulate, nb_bar, nb_par, na_bar, na_par);
                    b_ready = (fl & 1u) != 0;
                    a_next_ready = (fl & 2u) != 0;
                  } else {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) {
                      ptx::umma_bf16_2sm(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accumulate);
                      accumulate = 1;
                    }
                  }
                } else if (acc_mode == 1) {
                  const uint32_t fl = tap_merged_probe<KSTEPS, false, NSA == 2>(d_tmem, d_tmem, ad, ad + a_plane16, bd, idesc2, idesc,
                                                                      accumulate, nb_bar, nb_par, na_bar, na_par);
                  b_ready = (fl & 1u) != 0;
                  a_next_ready = (fl & 2u) != 0;
                } else if (acc_mode == 3 && NSA == 2) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    if ((k & 1) == 0) {
                      ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, accumulate);                    // hi*hi -> X
                      ptx::umma_bf16(d_tmem + bn, ad + 2 * k, bd + b_plane16 + 2 * k, idesc, accumulate);   // hi*lo -> Y
                      ptx::umma_bf16(d_tmem, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);                 // lo*hi -> X
                    } else {
                      ptx::umma_bf16(d_tmem + bn, ad + 2 * k, bd + b_plane16 + 2 * k, idesc, 1);            // hi*lo -> Y
                      ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc, 1);                             // hi*hi -> X
                      ptx::umma_bf16(d_tmem + bn, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);            // lo*hi -> Y
                    }
                    accumulate = 1;
                  }
                } else if (acc_mode == 2 && NSA == 2) {
#pragma unroll
                  for (int k = 0; k < KSTEPS; ++k) {
                    ptx::umma_bf16(d_tmem, ad + 2 * k, bd + 2 * k, idesc2, accumulate);                          // blocks 0-1
                    ptx::umma_bf16(d_tmem + 2 * bn, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, accumulate);      // block 2
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
                  }
                  if (NSA == 2) {
#pragma unroll
                    for (int k = 0; k < KSTEPS; ++k) ptx::umma_bf16(d_tmem, ad + a_plane16 + 2 * k, bd + 2 * k, idesc, 1);
                  }
                }
                if (PROF_ON(p)) prof_c[3] += clock64() - prof_i0;
                if (!p.wstat) {
                  if (pair) ptx::umma_commit_2sm(&b_empty[sb], kMask);
                  else if (CS > 1) ptx::umma_commit_mc(&b_empty[sb], kMask);
                  else ptx::umma_commit(&b_empty[sb]);
                }
              }
              accumulate = 1;
              if (++sb == p.SB) { sb = 0; b_par ^= 1; }
            }
            if (leader) {
              if (pair) ptx::umma_commit_2sm(&a_empty[sa], kMask);
              else ptx::umma_commit(&a_empty[sa]);
            }
            a_ready = a_next_ready;   // the probe fused into this slot's last tap
            a_next_ready = false;
            __syncwarp();
            if (++sa == p.SA) { sa = 0; a_par ^= 1; }
          }
        }
        if (leader) {
          if (pair) ptx::umma_commit_2sm(&acc_full[as], kMask);
          else ptx::umma_commit(&acc_full[as]);
        }
        as ^= 1;
      }
      if (PROF_ON(p) && leader) {
        long long* o = p.prof + (size_t)blockIdx.x * 16;
        o[3] = prof_c[0]; o[4] = prof_c[1]; o[5] = prof_c[2]; o[6] = clock64() - prof_start; o[15] = prof_c[3];
      }
      }   // generic issuer
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
    const uint32_t stage_s = ptx::smem_u32(stage);
    float* red = stage + (size_t)128 * ldst; // scratch: [32] pixel counts behind the BatchNorm sums, then
    float* part = red + 32;                  //          [8 warps][3][64]: per-warp partial (s1, s2, column sum) of a chunk
    // BatchNorm statistics of this CTA, per output channel: reference value k (the first accumulator the CTA saw for the
    // channel), s1 = sum(a - k), s2 = sum((a - k)^2) over every pixel it stored.  Shifting by k keeps both sums small, so
    // mean = k + s1/n and M2 = s2 - s1^2/n lose nothing to cancellation.
    float* run_stats = red + kEpiScratch;    // [3][Cout]: k, s1, s2
    float* st_k = run_stats;
    float* st_s1 = run_stats + p.Cout;
    float* st_s2 = run_stats + 2 * p.Cout;
    unsigned long long st_seen = 0;          // bit per CW-channel block: k has been chosen (uniform across the epilogue threads)
    float* colsum_s = run_stats + (p.stats ? p.Cout * 3 : 0);   // [Cout] column sums of this CTA (bias gradient)
    if (p.colsum) {
      for (int i = et; i < p.Cout; i += kEpiThreads) colsum_s[i] = 0.f;
      ptx::named_bar_sync(1, kEpiThreads);
    }
    if (p.stats) {
      for (int i = et; i < p.Cout * 3; i += kEpiThreads) run_stats[i] = 0.f;
      for (int i = et; i < p.tiles_n; i += kEpiThreads) red[i] = 0.f;
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
    const int obw_log = 31 - __clz(oBW);
    const bool lean_ok = (oBW & (oBW - 1)) == 0;
    // which compile-time specialisation of the store loop serves this launch (0 = the generic flag-driven loop), and which
    // 16-bit planes it writes (OUTK of epi_store)
    int lean_mode = 0, outk = 0;
    {
      const bool f32 = p.out_f32 != nullptr, spl = p.out_hi != nullptr;
      if (spl) {
        if (p.out_f16) outk = !p.out_lo ? -1 : (p.out_xb ? 3 : 2);
        else outk = p.out_xb ? -1 : (p.out_lo ? 1 : 4);
      } else if (p.out_lo || p.out_xb) {
        outk = -1;
      }
      if (outk >= 0) {
        if (!p.mask && !p.ups && p.reduce == 0 && f32 && !spl) lean_mode = 1;
        else if (!p.mask && !p.ups && p.reduce == 0 && !f32 && spl) lean_mode = 2;
        else if (!p.mask && p.ups && p.reduce == 0 && !f32 && spl) lean_mode = 3;
        else if (!p.mask && !p.ups && p.reduce == 1 && !f32 && spl) lean_mode = 4;
        else if (p.mask && !p.ups && p.reduce == 0 && !f32 && spl) lean_mode = 5;
        else if (p.mask && !p.ups && p.reduce == 2 && !f32 && spl) lean_mode = 6;
        else if (!p.mask && !p.ups && p.reduce == 0 && f32 && spl) lean_mode = 7;
        // plane sets each mode is instantiated for: forward modes 1 / 2 / 3, gradient modes 1 / 4
        if ((lean_mode >= 2 && lean_mode <= 4 && outk > 3) || (lean_mode == 7 && outk > 2) ||
            ((lean_mode == 5 || lean_mode == 6) && outk != 1 && outk != 4))
          lean_mode = 0;
      }
    }
    const int rep = (p.ups || p.sub == 1) ? 2 : 1;   // sub-pixel forward: this phase's pixels sit at stride 2 in the 2H x 2W output
    const int Hs = Ho * rep, Ws = Wo * rep;
    int as = 0;
    uint32_t full_par = 0;                   // bit `as`: phase parity of accumulator stage `as` (a register, not a local array)
    long long prof_c[2] = {0, 0};
    long long prof_e[5] = {0, 0, 0, 0, 0};   // phase 1 | barrier | store loop | reductions | barrier
    const long long prof_start = PROF_ON(p) ? clock64() : 0;
    // Everything about a tile that does not depend on the tile is set up ONCE: the per-tile code of the eight epilogue warps
    // (two per scheduler, so nothing hides its latency) is what bounds the 64-output-channel layers at 224^2, and it used to
    // spend ~700 instructions per warp and tile on item decoding, address arithmetic and re-loading the affine constants.
    const float asc = p.acc_scale;           // fp16 weights are packed pre-scaled by a power of two: folded into the affine scale
    const int mf = p.mask_ups ? 2 : 1;
    EpiTile e;
    e.st_addr = stage_s + (uint32_t)(g * 16);
    e.ldst_b = ldst * 4;
    e.pl = pl; e.PS = PS; e.npix = oBH * oBW; e.obw_log = obw_log;
    e.BW = p.BW;
    e.lo_clamp = p.relu ? 0.f : -INFINITY;
    e.out_row = rep * Ws * p.Cout; e.out_px = rep * p.Cout; e.ws_c = Ws * p.Cout;
    e.planar = p.out_planar; e.planar_stride = (size_t)p.planar_stride;
    if (p.out_planar) { e.out_row = (Wo >> 1) * p.Cout; e.out_px = p.Cout; }
    e.mask_row = mf * mf * Wo * p.Cout; e.mask_px = mf * p.Cout;
    e.vh = e.vw = 0; e.out_base = e.mask_base = 0;
    e.s4 = e.t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool affine = p.bias || p.scale || asc != 1.f;
    int n0_have = -1;                        // channel chunk whose affine constants e.s4 / e.t4 hold
    for (int w = cluster_id; w < p.num_items; w += num_clusters) {
      const Item it = decode_item(p, w, CS, rank);
      const int h0 = it.h0, w0 = it.w0, img = it.img;
      { PROF_T0(p); ptx::mbar_wait(&acc_full[as], (full_par >> as) & 1u); PROF_ADD(p, prof_c[0]); }
      full_par ^= 1u << as;
      ptx::tc_fence_after();
      const uint32_t t_acc = tmem_base + (uint32_t)as * acc_cols + ((uint32_t)(lg * 32) << 16);

      for (int cc = 0; cc < p.BN; cc += CW) {
        const int n0 = it.nt * p.BN + cc;   // first output channel of this chunk
        const long long tp0 = PROF_ON(p) ? clock64() : 0;
        // -- phase 1: TMEM -> registers -> smem staging [128][CW+4] (raw fp32 accumulators)
        for (int c0 = colsel * 32; c0 < (ABLATE(p, 4) ? 0 : CW); c0 += 64) {
          uint32_t v[32];
          // accumulator columns of the 32 channels [cc + c0, +32): block 0 and the block(s) the epilogue adds to it
          uint32_t col0 = (uint32_t)(cc + c0), col1 = (uint32_t)(p.BN + cc + c0);
          if (pair && NSPLIT == 2) {
            const uint32_t hb = (uint32_t)p.BN >> 1, cb = (uint32_t)(cc + c0);
            const uint32_t half = cb >= hb ? 1u : 0u;
            col0 = half * (uint32_t)p.BN + (cb - half * hb);
            col1 = col0 + hb;
          }
          ptx::tmem_ld_32x32(t_acc + col0, v);
          if (p.nsum >= 2) {
            uint32_t v2[32];
            ptx::tmem_ld_32x32(t_acc + col1, v2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
            if (p.nsum == 3) {
              ptx::tmem_ld_32x32(t_acc + (uint32_t)(2 * p.BN + cc + c0), v2);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
            }
          } else {
            ptx::tmem_ld_wait();
          }
          // (the staged values are the raw accumulators: p.acc_scale is applied with the affine constants / the statistics)
          const uint32_t dst = stage_s + (uint32_t)((m * ldst + c0) * 4);
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (c0 + j < CW)   // BN == 16: only half of the 32-column TMEM load is live
              ptx::sts128(dst + j * 4, v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        if (cc + CW >= p.BN) {
          // last chunk read: hand the accumulator stage back to the MMA warp (one arrive per epilogue warp)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (pair && rank != 0) ptx::mbar_arrive_leader_relaxed(&acc_empty[as]);   // rank 0's MMA thread waits for both CTAs
            else ptx::mbar_arrive(&acc_empty[as]);
          }
        }
        const long long tp1 = PROF_ON(p) ? clock64() : 0;
        ptx::named_bar_sync(1, kEpiThreads);
        const long long tp2 = PROF_ON(p) ? clock64() : 0;

        float4 cs = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 k4 = cs, s1v = cs, s2v = cs;
        if (it.valid && !ABLATE(p, 2)) {
          // -- phase 2: BatchNorm statistics are accumulated inside the store loop below (shifted sums, see st_k).  The first
          //    time the CTA meets a channel block it fixes the block's reference values: row 0 of the staged tile.
          if (p.stats) {
            const int blk = n0 / CW;
            if (!((st_seen >> blk) & 1ull)) {
              if (et < CW) st_k[n0 + et] = stage[et];
              ptx::named_bar_sync(2, kEpiThreads);
              st_seen |= 1ull << blk;
            }
            k4 = *reinterpret_cast<const float4*>(st_k + n0 + g * 4);
            if (cc == 0 && et == 0) red[it.nt] += (float)(min(p.BH, p.H - h0) * min(p.BW, p.W - w0));
          }

          // -- phase 3: (+bias, *scale+shift, relu) -> 2x2 max/sum -> mask -> coalesced NHWC stores.
          //    A thread keeps ONE float4 channel group and strides over pixels: its affine constants live in registers.
          const int ch = n0 + g * 4;
          if (n0 != n0_have) {   // one n-tile and one chunk per tile (the 64-channel layers): loaded once per launch
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f), s4 = make_float4(1.f, 1.f, 1.f, 1.f), t4 = b4;
            if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + ch));
            if (p.scale) {
              s4 = __ldg(reinterpret_cast<const float4*>(p.scale + ch));
              t4 = __ldg(reinterpret_cast<const float4*>(p.shift + ch));
            }
            // fold: v = (asc*acc + b)*s + t = acc*(asc*s) + (b*s + t)
            e.t4 = make_float4(fmaf(b4.x, s4.x, t4.x), fmaf(b4.y, s4.y, t4.y), fmaf(b4.z, s4.z, t4.z), fmaf(b4.w, s4.w, t4.w));
            e.s4 = make_float4(s4.x * asc, s4.y * asc, s4.z * asc, s4.w * asc);
            n0_have = n0;
          }
          const float4 s4 = e.s4, t4 = e.t4;
          const int oh0 = p.reduce ? h0 / 2 : h0;
          const int ow0 = p.reduce ? w0 / 2 : w0;
          bool done = false;
          if (lean_ok) {
            e.vh = min(oBH, Ho - oh0); e.vw = min(oBW, Wo - ow0);
            if (p.out_planar)   // (oh0, ow0) are even: the tile's phase-(0,0) origin inside one [H/2][W/2] plane
              e.out_base = ((size_t)(img * (Ho >> 1) + (oh0 >> 1)) * (Wo >> 1) + (size_t)(ow0 >> 1)) * p.Cout + ch;
            else
              e.out_base = ((size_t)(img * Hs + oh0 * rep + (it.phase >> 1)) * Ws + (size_t)(ow0 * rep + (it.phase & 1))) * p.Cout + ch;
            if (p.mask) e.mask_base = ((size_t)(img * mf * Ho + mf * oh0) * (mf * Wo) + (size_t)(mf * ow0)) * p.Cout + ch;
            done = true;
            // statistics ride on mode 1 only, column sums on the masked modes only; anything else takes the generic loop
            switch ((p.stats && lean_mode != 1) || (p.colsum && lean_mode != 5 && lean_mode != 6) ? 0 : lean_mode) {
              case 1:                                                              // fp32 out: train-mode trunk / dgrad into BN
                if (p.stats) epi_store<0, false, false, true, 0, true>(p, e, cs, k4, &s1v, &s2v);
                else epi_store<0, false, false, true, 0>(p, e, cs);
                break;
#define EGAZE_EPI_FWD(RED, UPS)                                                                 \
  if (outk == 1) epi_store<RED, false, UPS, false, 1>(p, e, cs);                                  \
  else if (outk == 2) epi_store<RED, false, UPS, false, 2>(p, e, cs);                             \
  else epi_store<RED, false, UPS, false, 3>(p, e, cs)
#define EGAZE_EPI_BWD(RED)                                                                      \
  if (outk == 1) epi_store<RED, true, false, false, 1>(p, e, cs);                                 \
  else epi_store<RED, true, false, false, 4>(p, e, cs)
              case 2: EGAZE_EPI_FWD(0, false); break;   // split out: decoder / eval trunk
              case 3: EGAZE_EPI_FWD(0, true); break;    // ... + nearest-2x replicate
              case 4: EGAZE_EPI_FWD(1, false); break;   // ... + 2x2 max-pool
              case 5: EGAZE_EPI_BWD(0); break;          // decoder dgrad: ReLU mask
              case 6: EGAZE_EPI_BWD(2); break;          // ... + 2x2 sum (grad of Upsample)
              case 7:                                   // both outputs
                if (outk == 1) epi_store<0, false, false, true, 1>(p, e, cs);
                else epi_store<0, false, false, true, 2>(p, e, cs);
                break;
#undef EGAZE_EPI_FWD
#undef EGAZE_EPI_BWD
              default: done = false;
            }
          }
          int ph = ph_start, pw = pw_start;
          for (int pix = pl; pix < (done ? 0 : oBH * oBW); pix += PS) {
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
                if (p.stats) {
                  const float4 a = *reinterpret_cast<const float4*>(stage + (size_t)(ph * p.BW + pw) * ldst + g * 4);
                  const float dx = a.x - k4.x, dy = a.y - k4.y, dz = a.z - k4.z, dw = a.w - k4.w;
                  s1v.x += dx; s1v.y += dy; s1v.z += dz; s1v.w += dw;
                  s2v.x = fmaf(dx, dx, s2v.x); s2v.y = fmaf(dy, dy, s2v.y); s2v.z = fmaf(dz, dz, s2v.z); s2v.w = fmaf(dw, dw, s2v.w);
                }
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
                if (!pos16(mk.x & 0xffffu)) v.x = 0.f;
                if (!pos16(mk.x >> 16)) v.y = 0.f;
                if (!pos16(mk.y & 0xffffu)) v.z = 0.f;
                if (!pos16(mk.y >> 16)) v.w = 0.f;
              }
              cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
              uint2 hi2 = make_uint2(0, 0), lo2 = make_uint2(0, 0), xb2 = make_uint2(0, 0);
              if (p.out_hi) {
                if (p.out_f16) split_f16x4(v, hi2, lo2);
                else split_bf16x4(v, hi2, lo2);
              }
              if (p.out_xb) xb2 = pack_bf16x4(v);
              const int nrep = p.ups ? 2 : 1;
              for (int dy = 0; dy < nrep; ++dy)
                for (int dx = 0; dx < nrep; ++dx) {
                  size_t off = ((size_t)(img * Hs + oh * rep + dy + (it.phase >> 1)) * Ws + (ow * rep + dx + (it.phase & 1))) * p.Cout + ch;
                  if (p.out_planar)
                    off = (size_t)((oh & 1) * 2 + (ow & 1)) * (size_t)p.planar_stride +
                          ((size_t)(img * (Ho >> 1) + (oh >> 1)) * (Wo >> 1) + (size_t)(ow >> 1)) * p.Cout + ch;
                  if (p.out_f32) *reinterpret_cast<float4*>(p.out_f32 + off) = v;
                  if (p.out_hi) *reinterpret_cast<uint2*>(p.out_hi + off) = hi2;
                  if (p.out_lo) *reinterpret_cast<uint2*>(p.out_lo + off) = lo2;
                  if (p.out_xb) *reinterpret_cast<uint2*>(p.out_xb + off) = xb2;
                }
            }
            pw += PS;
            while (pw >= oBW) { pw -= oBW; ++ph; }
          }
        }
        const long long tp3 = PROF_ON(p) ? clock64() : 0;
        // Fold the per-thread sums: lanes of a warp that share a channel group are gpp lanes apart (shuffles), then every
        // warp parks its partials in its own scratch row; after the chunk's closing barrier one thread per channel adds
        // the eight rows into the CTA's running sums.  (Shared-memory float atomics are CAS loops: 2k cycles per chunk.)
        const bool fold = (p.stats || p.colsum) && it.valid;
        if (fold) {
          for (int o = 16; o >= gpp; o >>= 1) {
            if (p.stats) {
              s1v.x += __shfl_xor_sync(0xffffffffu, s1v.x, o); s1v.y += __shfl_xor_sync(0xffffffffu, s1v.y, o);
              s1v.z += __shfl_xor_sync(0xffffffffu, s1v.z, o); s1v.w += __shfl_xor_sync(0xffffffffu, s1v.w, o);
              s2v.x += __shfl_xor_sync(0xffffffffu, s2v.x, o); s2v.y += __shfl_xor_sync(0xffffffffu, s2v.y, o);
              s2v.z += __shfl_xor_sync(0xffffffffu, s2v.z, o); s2v.w += __shfl_xor_sync(0xffffffffu, s2v.w, o);
            }
            if (p.colsum) {
              cs.x += __shfl_xor_sync(0xffffffffu, cs.x, o); cs.y += __shfl_xor_sync(0xffffffffu, cs.y, o);
              cs.z += __shfl_xor_sync(0xffffffffu, cs.z, o); cs.w += __shfl_xor_sync(0xffffffffu, cs.w, o);
            }
          }
          if (lane < gpp) {
            float* row = part + (size_t)ew * 3 * 64 + g * 4;
            if (p.stats) {
              *reinterpret_cast<float4*>(row) = s1v;
              *reinterpret_cast<float4*>(row + 64) = s2v;
            }
            if (p.colsum) *reinterpret_cast<float4*>(row + 128) = cs;
          }
        }
        const long long tp4 = PROF_ON(p) ? clock64() : 0;
        ptx::named_bar_sync(1, kEpiThreads);   // staging is free for the next chunk / item
        if (fold && et < 3 * 64) {
          // et -> (which sum, channel); the next write to `part` is two barriers away
          const int which = et >> 6, c = et & 63;
          if (c < CW && (which < 2 ? p.stats != nullptr : p.colsum != nullptr)) {
            float v = 0.f;
#pragma unroll
            for (int wv = 0; wv < 8; ++wv) v += part[(size_t)wv * 3 * 64 + which * 64 + c];
            float* dst = which == 0 ? st_s1 : (which == 1 ? st_s2 : colsum_s);
            dst[n0 + c] += v;
          }
        }
        if (PROF_ON(p)) {
          const long long tp5 = clock64();
          prof_e[0] += tp1 - tp0; prof_e[1] += tp2 - tp1; prof_e[2] += tp3 - tp2; prof_e[3] += tp4 - tp3; prof_e[4] += tp5 - tp4;
        }
      }
      as ^= 1;
    }
    if (PROF_ON(p) && et == 0) {
      long long* o = p.prof + (size_t)blockIdx.x * 16;
      o[7] = prof_c[0]; o[8] = clock64() - prof_start; o[9] = (p.num_items - cluster_id + num_clusters - 1) / num_clusters;
      for (int i = 0; i < 5; ++i) o[10 + i] = prof_e[i];
    }
    // the last chunk's per-warp partials are folded into the CTA sums AFTER its closing barrier: wait for those adds
    if (p.colsum || p.stats) ptx::named_bar_sync(1, kEpiThreads);
    if (p.colsum) {
      for (int c = et; c < p.Cout; c += kEpiThreads) atomicAdd(p.colsum + c, colsum_s[c]);
    }
    if (p.stats) {
      // one (mean, M2, n) partial per CTA and channel; bias shifts the mean only
      for (int c = et; c < p.Cout; c += kEpiThreads) {
        const float n = red[c / p.BN];
        float mean = 0.f, m2 = 0.f;
        if (n > 0.f) {   // the sums were taken over the raw accumulators: scale them (acc_scale is a power of two: exact)
          const float s1 = st_s1[c];
          mean = (st_k[c] + s1 / n) * p.acc_scale;
          m2 = fmaxf(st_s2[c] - s1 * s1 / n, 0.f) * (p.acc_scale * p.acc_scale);
        }
        p.stats[((size_t)blockIdx.x * 2 + 0) * p.Cout + c] = mean + (p.bias ? __ldg(p.bias + c) : 0.f);
        p.stats[((size_t)blockIdx.x * 2 + 1) * p.Cout + c] = m2;
        if (c % p.BN == 0) p.stats_cnt[(size_t)blockIdx.x * p.tiles_n + c / p.BN] = n;
      }
    }
  }

  __syncthreads();
  if (CS > 1) ptx::cluster_sync_all();   // no CTA exits while a peer may still arrive on its barriers
  if (warp == 1) {
    if (pair) ptx::tmem_dealloc2(tmem_base, tmem_cols);
    else ptx::tmem_dealloc(tmem_base, tmem_cols);
  }
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

// Window mode (one (BH+2) x 10-pixel window per K chunk serves all nine taps) needs BW == 8: every 8-row group of the
// A operand is then one tile row, and the groups are a uniform (BW+2) window rows apart.  Returns the tile efficiency.
double pick_window_tile(int H, int W, int need_even, int* BH) {
  double best = -1.0;
  for (int bh = 16; bh >= 2; bh -= 2) {
    (void)need_even;  // every candidate is even
    const double eff = (double)H * W / ((double)ceil_div(H, bh) * ceil_div(W, 8) * 128.0);
    if (eff > best * 1.0001) { best = eff; *BH = bh; }
  }
  return best;
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

// Bytes of the epilogue's shared memory behind the operand rings (1 KB granules): the [128][CW+4] staging tile, its scratch and
// the per-CTA running sums.
static int conv_stage_bytes(int bn, int Cout, bool stats, bool colsum) {
  const int bytes = 128 * ((bn < 64 ? bn : 64) + 4) * 4 + kEpiScratch * 4 + (stats ? Cout * 3 * 4 : 0) + (colsum ? Cout * 4 : 0);
  return (bytes + 1023) / 1024 * 1024;
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

static long long* g_conv_prof = nullptr;
// Debug aid: per-CTA cycle counters of the following egaze_conv3x3_tc launches are written to buf ([grid][16] int64;
// producer: 0 a_empty wait, 1 b_empty wait, 2 total | MMA: 3 acc_empty, 4 a_full, 5 b_full, 6 total | epilogue: 7 acc_full
// wait, 8 total, 9 items).  Pass null to switch it off.
extern "C" int egaze_conv3x3_set_prof(void* buf) {
#ifndef EGAZE_CONV_PROF
  EGAZE_CHECK_ARG(!buf, "conv3x3_set_prof: libegaze.so was built without -DEGAZE_CONV_PROF (EGAZE_CONV_PROF=1 python csrc/build.py)");
#endif
  g_conv_prof = (long long*)buf;
  return EGAZE_OK;
}

// A fully specified launch: tile / ring / accumulator configuration, the four TMA descriptors (cuTensorMapEncodeTiled) and the
// kernel parameters.  Built once per distinct call (egaze_conv3x3_plan_create) or on the stack (egaze_conv3x3_tc).
struct ConvPlan {
  CUtensorMap tmA_hi, tmA_lo, tmB_hi, tmB_lo;
  ConvTcParams p;
  int CS, ksteps, nsa, clusters, device;
  size_t smem;
};

static int conv_build(ConvPlan* pl, const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H, int W,
                      int Cin_p, int Cout, const float* bias, const float* scale, const float* shift, int relu,
                      int reduce, int ups, const void* mask, int mask_ups, float* out_f32, void* out_hi,
                      void* out_lo, void* out_xb, float* stats, float* stats_cnt, float* colsum, int in_f16,
                      int out_f16, float acc_scale, int sub, int out_planar) {
  ConvTcParams& p = pl->p;
  CUtensorMap& tmA_hi = pl->tmA_hi;
  CUtensorMap& tmA_lo = pl->tmA_lo;
  CUtensorMap& tmB_hi = pl->tmB_hi;
  CUtensorMap& tmB_lo = pl->tmB_lo;
  EGAZE_CHECK_ARG(x_hi && w_hi, "conv3x3_tc: null operand");
  EGAZE_CHECK_ARG(!(x_lo && !w_lo), "conv3x3_tc: an activation lo plane needs a weight lo plane (operand modes: hi+lo x hi+lo, hi x hi+lo, hi x hi)");
  EGAZE_CHECK_ARG(acc_scale > 0.f, "conv3x3_tc: acc_scale must be positive");
  const int nsa = x_lo ? 2 : 1, precise = w_lo ? 1 : 0;
  EGAZE_CHECK_ARG(N > 0 && H > 0 && W > 0, "conv3x3_tc: bad shape %d %d %d", N, H, W);
  EGAZE_CHECK_ARG(Cin_p % 16 == 0, "conv3x3_tc: Cin_p=%d must be a multiple of 16", Cin_p);
  EGAZE_CHECK_ARG(Cout % 16 == 0 && Cout >= 16, "conv3x3_tc: Cout=%d must be a multiple of 16", Cout);
  EGAZE_CHECK_ARG(!(reduce && ((H | W) & 1)), "conv3x3_tc: 2x2 reduce needs even H, W");
  EGAZE_CHECK_ARG(out_f32 || out_hi, "conv3x3_tc: no output");
  EGAZE_CHECK_ARG(sub >= 0 && sub <= 2, "conv3x3_tc: sub must be 0, 1 or 2");
  EGAZE_CHECK_ARG(!(sub && (reduce || ups || stats || Cin_p % 64 != 0)),
                  "conv3x3_tc: the sub-pixel modes need Cin_p %% 64 == 0 and take no reduce / ups / stats");
  EGAZE_CHECK_ARG(!(out_planar && (reduce || ups || sub == 1 || ((H | W) & 1))),
                  "conv3x3_tc: a phase-planar output needs even H, W and no reduce / ups / sub-pixel forward");

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

  memset(&p, 0, sizeof(p));
  p.N = N; p.H = H; p.W = W; p.Cin_p = Cin_p; p.Cout = Cout;
  p.KC = (Cin_p % 64 == 0) ? 64 : ((Cin_p % 32 == 0) ? 32 : 16);
  pick_tile(H, W, reduce != 0 || out_planar, &p.BH, &p.BW);
  static int win_env = -1;
  if (win_env < 0) {
    const char* e = getenv("EGAZE_CONV_WINDOW");
    win_env = e ? atoi(e) : 1;
  }
  static int win_minsb = 2;
  if (win_env > 0 && p.KC == 64) {
    const char* e = getenv("EGAZE_CONV_WINDOW_MINSB");
    if (e) win_minsb = atoi(e);
    int wbh = 16;
    const double eff_w = pick_window_tile(H, W, reduce != 0, &wbh);
    const double eff_c = (double)H * W / ((double)ceil_div(H, p.BH) * ceil_div(W, p.BW) * 128.0);
    // ring depth the weight boxes would get next to two 180-row windows (see the smem budget below)
    const int bn = conv_pick_bn(Cout, precise), ns = precise ? 2 : 1;
    const int stage_b = conv_stage_bytes(bn, Cout, stats != nullptr, colsum != nullptr);
    const int sb_w = (222 * 1024 - stage_b - 2 * nsa * 23552) / (ns * bn * 128);
    if ((eff_w >= eff_c * 0.999 && sb_w >= win_minsb) || sub) {
      p.win = 1;
      p.win_bo = win_env == 2 ? 1 : 0;
      p.BH = wbh;
      p.BW = 8;
    }
  }
  EGAZE_CHECK_ARG(!(sub && !p.win), "conv3x3_tc: the sub-pixel modes run in window mode only (EGAZE_CONV_WINDOW=0?)");
  p.sub = sub;
  p.out_planar = out_planar;
  p.planar_stride = (long long)N * (H / 2) * (W / 2) * Cout;
  p.nsplit = precise ? 2 : 1;
  p.nsplit_a = nsa;
  p.in_f16 = in_f16;
  p.acc_scale = acc_scale;
  // N tile: largest of 128/64/32/16 dividing Cout (256 only in fast mode where the rings fit).
  p.BN = conv_pick_bn(Cout, precise);
  p.tiles_h = ceil_div(H, p.BH); p.tiles_w = ceil_div(W, p.BW); p.tiles_n = Cout / p.BN;
  p.total_tiles = N * p.tiles_h * p.tiles_w;
  // clusters of 2 share every weight box through TMA multicast; the split needs 8-row-aligned halves
  int CS = cluster_env;
  if (p.total_tiles < 2 || (p.BN / 2) % 8 != 0) CS = 1;
  p.tile_groups = ceil_div(p.total_tiles, CS);
  p.num_items = p.tile_groups * p.tiles_n * (sub == 1 ? 4 : 1);
  p.fd_tn = make_fastdiv(p.tiles_n); p.fd_tw = make_fastdiv(p.tiles_w); p.fd_th = make_fastdiv(p.tiles_h);
  const int row_bytes = p.KC * 2;
  int a_rows = (p.BH + 2) * p.BW;
  if (a_rows < 2 * p.BW + 128) a_rows = 2 * p.BW + 128;
  // window mode: the MMA always walks 16 row groups, (BW+2) window rows apart, from up to 2 rows + 2 pixels in
  if (p.win) a_rows = 17 * (p.BW + 2) + 2 + 8;
  p.a_slot_bytes = ((a_rows * row_bytes + 1023) / 1024) * 1024;
  // CTA-pair MMA (cta_group::2): needs the merged two-plane weight slot (or the single-plane fast mode), 32-column epilogue
  // blocks that stay inside one channel half (BN >= 64) and N <= 256
  {
    static int pair_env = -1;
    if (pair_env < 0) {
      const char* e = getenv("EGAZE_CONV_PAIR");
      pair_env = e ? atoi(e) : 1;
    }
    p.pair = (pair_env && CS == 2 && p.BN >= 64 && p.BN <= 128 && (p.BN / 2 * row_bytes) % 1024 == 0) ? 1 : 0;
  }
  const int b_plane_rows = p.pair ? p.BN / 2 : p.BN;   // weight rows of one plane kept in THIS CTA's slot
  p.b_slot_bytes = ((b_plane_rows * row_bytes + 1023) / 1024) * 1024;
  const int CW = p.BN < 64 ? p.BN : 64;
  const int stage_bytes = conv_stage_bytes(p.BN, Cout, stats != nullptr, colsum != nullptr);
  (void)CW;
  static int sa_env = -1;
  if (sa_env < 0) {
    const char* e = getenv("EGAZE_CONV_SA");
    sa_env = e ? atoi(e) : 0;
  }
  p.SA = 2;
  const int budget = 222 * 1024 - stage_bytes;
  if (sa_env >= 2 && sa_env <= kMaxSA && (budget - sa_env * nsa * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes) >= 3)
    p.SA = sa_env;
  else if (nsa == 1 && (budget - 3 * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes) >= 6)
    p.SA = 3;   // one activation plane: the freed shared memory buys a third window slot
  int sb = (budget - p.SA * nsa * p.a_slot_bytes) / (p.nsplit * p.b_slot_bytes);
  {
    static int wstat_env = -1;
    if (wstat_env < 0) {
      const char* e = getenv("EGAZE_CONV_WSTAT");
      wstat_env = e ? atoi(e) : 1;
    }
    const int boxes = 9 * (Cin_p / p.KC);
    if (wstat_env && !sub && p.pair && p.tiles_n == 1 && boxes <= kMaxSB && sb >= boxes) {
      p.wstat = 1;
      sb = boxes;
    }
  }
  if (!p.wstat && sb > 6) sb = 6;
  EGAZE_CHECK_ARG(sb >= 2, "conv3x3_tc: tile does not fit shared memory");
  p.SB = sb;
  {
    static int probe_env = -1;
    if (probe_env < 0) {
      const char* e = getenv("EGAZE_CONV_PROBE");
      probe_env = e ? atoi(e) : 0;
    }
    p.probe = probe_env;
  }
  // accumulator layout (see ConvTcParams::acc_mode).  The N = 2*BN MMA of modes 1/2 needs the two weight planes back to
  // back in the slot (no padding) and 2*BN <= 256.
  {
    static int mode_env = -2;
    if (mode_env == -2) {
      const char* e = getenv("EGAZE_CONV_ACCMODE");
      mode_env = e ? atoi(e) : -1;
    }
    const bool can_merge = p.b_slot_bytes == b_plane_rows * row_bytes && 2 * p.BN <= 256;
    int mode = 0;
    if (precise) {
      // measured (B=32 SP layers): mode 1 beats 2 and 3 -- fewer, larger MMAs win; the extra smem operand reads of mode 3
      // cost more than any accumulator interleaving gains
      mode = can_merge ? 1 : (nsa == 2 ? 3 : 0);
      if (mode_env >= 0 && nsa == 2) mode = mode_env;
      if ((mode == 1 || mode == 2) && !can_merge) mode = 3;
      if (mode == 2 && p.BN > 64) mode = 1;
      if (p.pair) {
        EGAZE_CHECK_ARG(can_merge, "conv3x3_tc: CTA-pair mode needs the merged weight slot");
        mode = 1;
      }
    }
    p.acc_mode = mode;
    p.nsum = mode == 0 ? 1 : (mode == 2 ? 3 : 2);
    int cols = p.nsum * p.BN;
    p.acc_cols = 32;
    while (p.acc_cols < cols) p.acc_cols *= 2;
    EGAZE_CHECK_ARG(2 * p.acc_cols <= 512, "conv3x3_tc: accumulators do not fit TMEM (BN=%d mode=%d)", p.BN, mode);
  }
  p.stage_off = p.SA * nsa * p.a_slot_bytes + p.SB * p.nsplit * p.b_slot_bytes;
  const size_t smem = (size_t)p.stage_off + stage_bytes + 1024;  // + alignment slack
  p.bias = bias; p.scale = scale; p.shift = shift; p.relu = relu; p.reduce = reduce; p.ups = ups;
  p.mask = (const __nv_bfloat16*)mask;
  p.mask_ups = mask_ups;
  p.out_f32 = out_f32; p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
  p.out_xb = (__nv_bfloat16*)out_xb; p.out_f16 = out_f16;
  EGAZE_CHECK_ARG(!(out_lo && !out_hi) && !(out_xb && !out_hi), "conv3x3_tc: out_lo / out_xb need out_hi");
  EGAZE_CHECK_ARG(!(out_f16 && out_hi && !out_lo), "conv3x3_tc: fp16 output needs both planes");
  p.stats = stats; p.stats_cnt = stats_cnt;
  p.colsum = colsum;
  p.prof = g_conv_prof;
#ifdef EGAZE_CONV_PROF
  { const char* e = getenv("EGAZE_CONV_ABLATE"); p.ablate = e ? atoi(e) : 0; }
#endif

  {
    uint64_t dims[4] = {(uint64_t)Cin_p, (uint64_t)W, (uint64_t)H, (uint64_t)(sub == 2 ? 4 * N : N)};
    uint64_t str[3] = {(uint64_t)Cin_p * 2, (uint64_t)W * Cin_p * 2, (uint64_t)H * W * Cin_p * 2};
    uint32_t box[4] = {(uint32_t)p.KC, (uint32_t)(p.win ? p.BW + 2 : p.BW), (uint32_t)(p.BH + 2), 1};
    int rc = egaze_encode_tmap(&tmA_hi, x_hi, 4, dims, str, box, row_bytes, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmA_lo, nsa == 2 ? x_lo : x_hi, 4, dims, str, box, row_bytes, 2);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {(uint64_t)Cin_p, (uint64_t)(sub ? 16 : 9) * Cout};
    uint64_t str[1] = {(uint64_t)Cin_p * 2};
    uint32_t box[2] = {(uint32_t)p.KC, (uint32_t)(p.BN / CS)};   // each CTA of the cluster fetches its share of the box
    int rc = egaze_encode_tmap(&tmB_hi, w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
    rc = egaze_encode_tmap(&tmB_lo, precise ? w_lo : w_hi, 2, dims, str, box, row_bytes, 2);
    if (rc) return rc;
  }
  int clusters = sm_count / CS;
  if (clusters > p.num_items) clusters = p.num_items;
  pl->CS = CS; pl->ksteps = p.KC / 16; pl->nsa = nsa; pl->clusters = clusters; pl->smem = smem;
  EGAZE_CUDA(cudaGetDevice(&pl->device));
  return EGAZE_OK;
}

static int conv_run(const ConvPlan& pl, void* stream) {
  const ConvTcParams& p = pl.p;
  const int CS = pl.CS, nsa = pl.nsa;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(pl.clusters * CS));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = CS > 1 ? 1 : 0;
  const int ksteps = pl.ksteps;
#define EGAZE_CONV_LAUNCH(NA, NS, KS, C)                                                                              \
  do {                                                                                                                \
    static unsigned long long attr_set = 0;                                                                           \
    if (egaze_first_on_device(&attr_set)) {                                                                           \
      EGAZE_CUDA(cudaFuncSetAttribute(conv3x3_tc_kernel<NA, NS, KS, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      224 * 1024));                                                                   \
    }                                                                                                                 \
    EGAZE_CUDA(cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel<NA, NS, KS, C>, pl.tmA_hi, pl.tmA_lo, pl.tmB_hi, pl.tmB_lo, p)); \
  } while (0)
#define EGAZE_CONV_DISPATCH(NA, NS, C)                                                                                \
  do {                                                                                                                \
    if (ksteps == 4) EGAZE_CONV_LAUNCH(NA, NS, 4, C);                                                                 \
    else if (ksteps == 2) EGAZE_CONV_LAUNCH(NA, NS, 2, C);                                                            \
    else EGAZE_CONV_LAUNCH(NA, NS, 1, C);                                                                             \
  } while (0)
  if (p.nsplit == 2 && nsa == 2) {
    if (CS == 2) EGAZE_CONV_DISPATCH(2, 2, 2);
    else EGAZE_CONV_DISPATCH(2, 2, 1);
  } else if (p.nsplit == 2) {
    if (CS == 2) EGAZE_CONV_DISPATCH(1, 2, 2);
    else EGAZE_CONV_DISPATCH(1, 2, 1);
  } else {
    if (CS == 2) EGAZE_CONV_DISPATCH(1, 1, 2);
    else EGAZE_CONV_DISPATCH(1, 1, 1);
  }
#undef EGAZE_CONV_DISPATCH
#undef EGAZE_CONV_LAUNCH
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// See include/egaze.h for the contract.
extern "C" int egaze_conv3x3_tc(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H, int W,
                                int Cin_p, int Cout, const float* bias, const float* scale, const float* shift, int relu,
                                int reduce, int ups, const void* mask, int mask_ups, float* out_f32, void* out_hi,
                                void* out_lo, void* out_xb, float* stats, float* stats_cnt, float* colsum, int in_f16,
                                int out_f16, float acc_scale, int sub, int out_planar, void* stream) {
  ConvPlan pl;
  const int rc = conv_build(&pl, x_hi, x_lo, w_hi, w_lo, N, H, W, Cin_p, Cout, bias, scale, shift, relu, reduce, ups, mask, mask_ups,
                            out_f32, out_hi, out_lo, out_xb, stats, stats_cnt, colsum, in_f16, out_f16, acc_scale, sub, out_planar);
  return rc ? rc : conv_run(pl, stream);
}

// Plans (SURVEY 8b "egaze_plan_{create,destroy}: caches CUtensorMap descriptors per (ptr, shape)").  A plan freezes one
// egaze_conv3x3_tc call -- every pointer, shape and flag -- with its four encoded TMA descriptors and its tile configuration;
// egaze_conv3x3_plan_run launches it with two arguments.  The caller owns the handle and must not run it after any of the
// buffers it names has been freed (the Python binding keys its plan cache on the full argument tuple, egaze/ops.py).
extern "C" int egaze_conv3x3_plan_create(const void* x_hi, const void* x_lo, const void* w_hi, const void* w_lo, int N, int H,
                                         int W, int Cin_p, int Cout, const float* bias, const float* scale, const float* shift,
                                         int relu, int reduce, int ups, const void* mask, int mask_ups, float* out_f32,
                                         void* out_hi, void* out_lo, void* out_xb, float* stats, float* stats_cnt,
                                         float* colsum, int in_f16, int out_f16, float acc_scale, int sub, int out_planar,
                                         long long* plan) {
  EGAZE_CHECK_ARG(plan, "conv3x3_plan_create: null handle pointer");
  ConvPlan* pl = new (std::nothrow) ConvPlan;
  EGAZE_CHECK_ARG(pl, "conv3x3_plan_create: out of host memory");
  const int rc = conv_build(pl, x_hi, x_lo, w_hi, w_lo, N, H, W, Cin_p, Cout, bias, scale, shift, relu, reduce, ups, mask, mask_ups,
                            out_f32, out_hi, out_lo, out_xb, stats, stats_cnt, colsum, in_f16, out_f16, acc_scale, sub, out_planar);
  if (rc) {
    delete pl;
    return rc;
  }
  *plan = (long long)reinterpret_cast<intptr_t>(pl);
  return EGAZE_OK;
}

extern "C" int egaze_conv3x3_plan_run(long long plan, void* stream) {
  const ConvPlan* pl = reinterpret_cast<const ConvPlan*>((intptr_t)plan);
  EGAZE_CHECK_ARG(pl, "conv3x3_plan_run: null plan");
  int dev = -1;
  EGAZE_CUDA(cudaGetDevice(&dev));
  EGAZE_CHECK_ARG(dev == pl->device, "conv3x3_plan_run: plan was created on device %d, current device is %d", pl->device, dev);
  return conv_run(*pl, stream);
}

extern "C" int egaze_plan_destroy(long long plan) {
  delete reinterpret_cast<ConvPlan*>((intptr_t)plan);
  return EGAZE_OK;
}
