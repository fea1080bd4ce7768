// lstmnet (reference models/LSTMnet.py:15-37): tanh -> 2-layer nn.LSTM(512,512) -> Linear(512,512) -> ReLU.
// Gate order i,f,g,o; c' = f*c + i*g; h' = o*tanh(c').  fp32 end to end (CUDA cores): the step is
// weight-bandwidth / latency bound (8.9 MFLOP per sample-step against 17.9 MB of weights that stay L2-resident),
// so one warp owns one hidden unit, streams its 8 weight rows with 16-byte loads and reduces with shuffles.
#include "common.cuh"

namespace {

constexpr int HID = 512;
constexpr int BT = 8;  // batch tile per block pass

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// One LSTM cell step for one layer.
//  x    : [B][512] layer input (if apply_tanh, tanh is applied on load: layer 0 input)
//  h,c  : [B][512] previous state;  h_out,c_out : [B][512]
//  gates_out (optional, training): [B][4][512] post-activation gates i,f,g,o
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const float* __restrict__ x, int apply_tanh, const float* __restrict__ h, const float* __restrict__ c,
                 const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                 const float* __restrict__ b_hh, int B, float* __restrict__ h_out, float* __restrict__ c_out,
                 float* __restrict__ gates_out) {
  __shared__ float xs[BT][HID];
  __shared__ float hs[BT][HID];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;  // hidden unit
  const int b0 = blockIdx.y * BT;
  const int nb = min(BT, B - b0);
  for (int i = threadIdx.x; i < nb * HID; i += blockDim.x) {
    const int bb = i / HID, k = i % HID;
    float xv = x[(size_t)(b0 + bb) * HID + k];
    xs[bb][k] = apply_tanh ? tanhf(xv) : xv;
    hs[bb][k] = h[(size_t)(b0 + bb) * HID + k];
  }
  __syncthreads();
  float acc[BT][4];
#pragma unroll
  for (int bb = 0; bb < BT; ++bb)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[bb][g] = 0.f;
#pragma unroll
  for (int it = 0; it < HID / 128; ++it) {
    const int k = it * 128 + lane * 4;
    float4 wi[4], wh[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      wi[g] = __ldg(reinterpret_cast<const float4*>(w_ih + (size_t)(g * HID + j) * HID + k));
      wh[g] = __ldg(reinterpret_cast<const float4*>(w_hh + (size_t)(g * HID + j) * HID + k));
    }
#pragma unroll
    for (int bb = 0; bb < BT; ++bb) {
      if (bb < nb) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[bb][k]);
        const float4 hv = *reinterpret_cast<const float4*>(&hs[bb][k]);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a = acc[bb][g];
          a = fmaf(wi[g].x, xv.x, a); a = fmaf(wi[g].y, xv.y, a); a = fmaf(wi[g].z, xv.z, a); a = fmaf(wi[g].w, xv.w, a);
          a = fmaf(wh[g].x, hv.x, a); a = fmaf(wh[g].y, hv.y, a); a = fmaf(wh[g].z, hv.z, a); a = fmaf(wh[g].w, hv.w, a);
          acc[bb][g] = a;
        }
      }
    }
  }
#pragma unroll
  for (int bb = 0; bb < BT; ++bb)
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[bb][g] = warp_sum(acc[bb][g]);
  if (lane < nb) {
    // lane bb finalises batch element bb (select without dynamic register indexing)
    float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f;
#pragma unroll
    for (int bb = 0; bb < BT; ++bb)
      if (bb == lane) { gi = acc[bb][0]; gf = acc[bb][1]; gg = acc[bb][2]; go = acc[bb][3]; }
    gi += b_ih[0 * HID + j] + b_hh[0 * HID + j];
    gf += b_ih[1 * HID + j] + b_hh[1 * HID + j];
    gg += b_ih[2 * HID + j] + b_hh[2 * HID + j];
    go += b_ih[3 * HID + j] + b_hh[3 * HID + j];
    const float i_ = sigmoidf_acc(gi), f_ = sigmoidf_acc(gf), g_ = tanhf(gg), o_ = sigmoidf_acc(go);
    const size_t o = (size_t)(b0 + lane) * HID + j;
    const float cn = f_ * c[o] + i_ * g_;
    c_out[o] = cn;
    h_out[o] = o_ * tanhf(cn);
    if (gates_out) {
      float* gp = gates_out + (size_t)(b0 + lane) * 4 * HID + j;
      gp[0] = i_; gp[HID] = f_; gp[2 * HID] = g_; gp[3 * HID] = o_;
    }
  }
}

// y[r][j] = act(sum_k x[r][k] * W[j][k] + b[j]),  K = 512;  act: 0 none, 1 relu
__global__ void __launch_bounds__(256)
linear512_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int rows, int J,
                 int act, float* __restrict__ y) {
  __shared__ float xs[BT][HID];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const int r0 = blockIdx.y * BT;
  const int nr = min(BT, rows - r0);
  for (int i = threadIdx.x; i < nr * HID; i += blockDim.x) xs[i / HID][i % HID] = x[(size_t)(r0 + i / HID) * HID + i % HID];
  __syncthreads();
  if (j >= J) return;
  float acc[BT];
#pragma unroll
  for (int bb = 0; bb < BT; ++bb) acc[bb] = 0.f;
#pragma unroll
  for (int it = 0; it < HID / 128; ++it) {
    const int k = it * 128 + lane * 4;
    const float4 wv = __ldg(reinterpret_cast<const float4*>(w + (size_t)j * HID + k));
#pragma unroll
    for (int bb = 0; bb < BT; ++bb)
      if (bb < nr) {
        const float4 xv = *reinterpret_cast<const float4*>(&xs[bb][k]);
        float a = acc[bb];
        a = fmaf(wv.x, xv.x, a); a = fmaf(wv.y, xv.y, a); a = fmaf(wv.z, xv.z, a); a = fmaf(wv.w, xv.w, a);
        acc[bb] = a;
      }
  }
#pragma unroll
  for (int bb = 0; bb < BT; ++bb) acc[bb] = warp_sum(acc[bb]);
  if (lane < nr) {
    float v = 0.f;
#pragma unroll
    for (int bb = 0; bb < BT; ++bb)
      if (bb == lane) v = acc[bb];
    v += b ? b[j] : 0.f;
    if (act == 1) v = fmaxf(v, 0.f);
    y[(size_t)(r0 + lane) * J + j] = v;
  }
}

}  // namespace

// Sequence forward.  All pointers device fp32.
//  x      : [T][B][512]     (raw; tanh applied inside, LSTMnet.py:28)
//  h0,c0  : [2][B][512]     initial state (zeros when the reference passes hidden=None)
//  w_ih/w_hh/b_ih/b_hh : arrays of 2 device pointers (layer 0, 1): [2048][512] / [2048]
//  lin_w  : [512][512], lin_b : [512]
//  out    : [T][B][512]     relu(lin(h_top))
//  hn,cn  : [2][B][512]     final state
//  ws_h   : workspace [T][2][B][512] hidden states per step and layer (kept for backward)
//  ws_c   : workspace [T][2][B][512] cell states
//  ws_gates (optional): [T][2][B][4][512]
extern "C" int egaze_lstm_seq_fwd(const float* x, const float* h0, const float* c0, const float* const* w_ih,
                                  const float* const* w_hh, const float* const* b_ih, const float* const* b_hh,
                                  const float* lin_w, const float* lin_b, int T, int B, float* out, float* hn, float* cn,
                                  float* ws_h, float* ws_c, float* ws_gates, void* stream) {
  EGAZE_CHECK_ARG(x && h0 && c0 && w_ih && w_hh && b_ih && b_hh && lin_w && out && hn && cn && ws_h && ws_c,
                  "lstm_seq_fwd: null pointer");
  EGAZE_CHECK_ARG(T > 0 && B > 0, "lstm_seq_fwd: bad T=%d B=%d", T, B);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sl = (size_t)B * HID;
  dim3 grid(HID / 8, ceil_div(B, BT));
  for (int t = 0; t < T; ++t) {
    for (int l = 0; l < 2; ++l) {
      const float* xin = l == 0 ? x + (size_t)t * sl : ws_h + ((size_t)t * 2 + 0) * sl;
      const float* hp = t == 0 ? h0 + l * sl : ws_h + ((size_t)(t - 1) * 2 + l) * sl;
      const float* cp = t == 0 ? c0 + l * sl : ws_c + ((size_t)(t - 1) * 2 + l) * sl;
      float* ho = ws_h + ((size_t)t * 2 + l) * sl;
      float* co = ws_c + ((size_t)t * 2 + l) * sl;
      float* go = ws_gates ? ws_gates + ((size_t)t * 2 + l) * sl * 4 : nullptr;
      lstm_cell_kernel<<<grid, 256, 0, st>>>(xin, l == 0, hp, cp, w_ih[l], w_hh[l], b_ih[l], b_hh[l], B, ho, co, go);
      EGAZE_LAUNCH_CHECK();
    }
  }
  // top-layer hidden states of all steps -> Linear + ReLU.  Gather rows (t, layer 1) with a strided view:
  // rows are [T][B], row stride 2*sl between steps -> run per step to keep the kernel simple.
  dim3 lgrid(HID / 8, ceil_div(B, BT));
  for (int t = 0; t < T; ++t) {
    linear512_kernel<<<lgrid, 256, 0, st>>>(ws_h + ((size_t)t * 2 + 1) * sl, lin_w, lin_b, B, HID, 1, out + (size_t)t * sl);
    EGAZE_LAUNCH_CHECK();
  }
  EGAZE_CUDA(cudaMemcpyAsync(hn, ws_h + ((size_t)(T - 1) * 2) * sl, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EGAZE_CUDA(cudaMemcpyAsync(cn, ws_c + ((size_t)(T - 1) * 2) * sl, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return EGAZE_OK;
}
