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
struct CellArgs {
  const float* x; int apply_tanh;
  const float* h; const float* c;
  const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
  float* h_out; float* c_out; float* gates_out;
};
// gridDim.z = 2 runs two independent cells in one launch: layer 0 at step t and layer 1 at step t-1 (the wavefront of the
// two-layer recurrence), which halves the number of dependent launches of a sequence.
__global__ void __launch_bounds__(256)
lstm_cell_kernel(const CellArgs a0, const CellArgs a1, int B) {
  const CellArgs& a = blockIdx.z == 0 ? a0 : a1;
  const float* __restrict__ x = a.x; const int apply_tanh = a.apply_tanh;
  const float* __restrict__ h = a.h; const float* __restrict__ c = a.c;
  const float* __restrict__ w_ih = a.w_ih; const float* __restrict__ w_hh = a.w_hh;
  const float* __restrict__ b_ih = a.b_ih; const float* __restrict__ b_hh = a.b_hh;
  float* __restrict__ h_out = a.h_out; float* __restrict__ c_out = a.c_out; float* __restrict__ gates_out = a.gates_out;
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
// Input rows come in groups of `grp` consecutive rows, `grp_stride` floats apart (grp_stride == grp*512: dense).
__global__ void __launch_bounds__(256)
linear512_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, int rows, int J,
                 int act, float* __restrict__ y, int grp, long long grp_stride) {
  __shared__ float xs[BT][HID];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 8 + warp;
  const int r0 = blockIdx.y * BT;
  const int nr = min(BT, rows - r0);
  for (int i = threadIdx.x; i < nr * HID; i += blockDim.x) {
    const int r = r0 + i / HID;
    xs[i / HID][i % HID] = x[(size_t)(r / grp) * grp_stride + (size_t)(r % grp) * HID + i % HID];
  }
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


// ---- backward (BPTT) ------------------------------------------------------------------------------------------------
// dz = gout * (out > 0)   (ReLU after the Linear, LSTMnet.py:37)
__global__ void relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ out, size_t n, float* __restrict__ dz) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dz[i] = out[i] > 0.f ? g[i] : 0.f;
}

// One cell, one time step: dh = dh_a + dh_b;  gates = post-activation (i,f,g,o);  dc_io holds dc from step t+1 on entry
// and dc for step t-1 on exit.  dgates = gradients w.r.t. the PRE-activation gates.
__global__ void lstm_gate_bwd_kernel(const float* __restrict__ dh_a, const float* __restrict__ dh_b,
                                     const float* __restrict__ gates, const float* __restrict__ c,
                                     const float* __restrict__ c_prev, int B, float* __restrict__ dc_io,
                                     float* __restrict__ dgates) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * HID) return;
  const int b = idx / HID, j = idx % HID;
  const float* gp = gates + (size_t)b * 4 * HID + j;
  const float i_ = gp[0], f_ = gp[HID], g_ = gp[2 * HID], o_ = gp[3 * HID];
  float dh = dh_a[idx];
  if (dh_b) dh += dh_b[idx];
  const float tc = tanhf(c[idx]);
  const float dc = dh * o_ * (1.f - tc * tc) + dc_io[idx];
  float* dg = dgates + (size_t)b * 4 * HID + j;
  dg[0] = dc * g_ * i_ * (1.f - i_);
  dg[HID] = dc * c_prev[idx] * f_ * (1.f - f_);
  dg[2 * HID] = dc * i_ * (1.f - g_ * g_);
  dg[3 * HID] = dh * tc * o_ * (1.f - o_);
  dc_io[idx] = dc * f_;
}

// Y[r][n] (+)= sum_k X[r][k] * W[k][n]      (dx = dgates . W_ih, dh_prev = dgates . W_hh, dh_top = dz . W_lin)
__global__ void matmul_nn_kernel(const float* __restrict__ X, const float* __restrict__ W, int R, int K, int N,
                                 float* __restrict__ Y) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (n >= N || r >= R) return;
  const float* x = X + (size_t)r * K;
  float acc = 0.f;
#pragma unroll 4
  for (int k = 0; k < K; ++k) acc = fmaf(__ldg(x + k), W[(size_t)k * N + n], acc);
  Y[(size_t)r * N + n] = acc;
}

// dW[m][n] (+)= sum_{t<T,b<B} A[t*a_st + b*a_sb + m] * Bm[t*b_st + b*b_sb + n]   (weight gradients over all rows)
__global__ void matmul_tn_kernel(const float* __restrict__ A, size_t a_st, size_t a_sb, const float* __restrict__ Bm,
                                 size_t b_st, size_t b_sb, int T, int B, int M, int N, int accumulate,
                                 float* __restrict__ dW) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = accumulate ? dW[(size_t)m * N + n] : 0.f;
  for (int t = 0; t < T; ++t)
    for (int b = 0; b < B; ++b)
      acc = fmaf(__ldg(A + t * a_st + b * a_sb + m), Bm[t * b_st + b * b_sb + n], acc);
  dW[(size_t)m * N + n] = acc;
}

__global__ void col_sum_f32_kernel(const float* __restrict__ X, size_t ld, int R, int N, float* __restrict__ out) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc += X[(size_t)r * ld + n];
  out[n] = acc;
}

// xt = tanh(x) ; optionally dx = g * (1 - tanh(x)^2)
__global__ void tanh_kernel(const float* __restrict__ x, size_t n, float* __restrict__ xt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) xt[i] = tanhf(x[i]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ xt, const float* __restrict__ g, size_t n, float* __restrict__ dx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dx[i] = g[i] * (1.f - xt[i] * xt[i]);
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
  // Wavefront over the two layers: launch t runs layer 0 at step t and layer 1 at step t-1 (both only depend on launch t-1),
  // so a sequence is T+1 dependent launches instead of 2T.
  auto cell = [&](int l, int t) {
    CellArgs a;
    a.x = l == 0 ? x + (size_t)t * sl : ws_h + ((size_t)t * 2 + 0) * sl;
    a.apply_tanh = l == 0;
    a.h = t == 0 ? h0 + l * sl : ws_h + ((size_t)(t - 1) * 2 + l) * sl;
    a.c = t == 0 ? c0 + l * sl : ws_c + ((size_t)(t - 1) * 2 + l) * sl;
    a.w_ih = w_ih[l]; a.w_hh = w_hh[l]; a.b_ih = b_ih[l]; a.b_hh = b_hh[l];
    a.h_out = ws_h + ((size_t)t * 2 + l) * sl;
    a.c_out = ws_c + ((size_t)t * 2 + l) * sl;
    a.gates_out = ws_gates ? ws_gates + ((size_t)t * 2 + l) * sl * 4 : nullptr;
    return a;
  };
  for (int t = 0; t <= T; ++t) {
    const bool has0 = t < T, has1 = t >= 1;
    const CellArgs a0 = has0 ? cell(0, t) : cell(1, t - 1);
    const CellArgs a1 = has1 ? cell(1, t - 1) : a0;
    dim3 grid(HID / 8, ceil_div(B, BT), (has0 && has1) ? 2 : 1);
    lstm_cell_kernel<<<grid, 256, 0, st>>>(a0, a1, B);
    EGAZE_LAUNCH_CHECK();
  }
  // top-layer hidden states of all steps -> Linear + ReLU in ONE launch: rows (t, b) live B at a time, 2*sl floats apart
  dim3 lgrid(HID / 8, ceil_div(T * B, BT));
  linear512_kernel<<<lgrid, 256, 0, st>>>(ws_h + sl, lin_w, lin_b, T * B, HID, 1, out, B, (long long)(2 * sl));
  EGAZE_LAUNCH_CHECK();
  EGAZE_CUDA(cudaMemcpyAsync(hn, ws_h + ((size_t)(T - 1) * 2) * sl, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  EGAZE_CUDA(cudaMemcpyAsync(cn, ws_c + ((size_t)(T - 1) * 2) * sl, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return EGAZE_OK;
}

// Sequence backward (BPTT) of lstmnet.forward -- what loss.backward() runs in AT.trainLSTM (AT.py:138-142).
//  gout [T][B][512] gradient w.r.t. the ReLU output; ghn/gcn [2][B][512] gradients w.r.t. the returned state (may be NULL)
//  out  [T][B][512] forward output (ReLU mask); ws_h/ws_c [T][2][B][512], ws_gates [T][2][B][4][512]: forward workspaces
//  workspaces: xt, dz, dh_top [T][B][512]; dgates [2][T][B][2048]; tmp_x [B][512]; dh_next, dc_next [2][B][512];
//              dx0 [T][B][512] (only when dinput != NULL)
//  outputs: dw_ih/dw_hh: 2 pointers to [2048][512]; db: 2 pointers to [2048] (d bias_ih == d bias_hh); dlin_w [512][512];
//           dlin_b [512]; dinput [T][B][512] (optional); dh0/dc0 [2][B][512] (optional)
extern "C" int egaze_lstm_seq_bwd(const float* x, const float* h0, const float* c0, const float* const* w_ih,
                                  const float* const* w_hh, const float* lin_w, int T, int B, const float* out,
                                  const float* gout, const float* ghn, const float* gcn, const float* ws_h,
                                  const float* ws_c, const float* ws_gates, float* xt, float* dz, float* dgates,
                                  float* dh_top, float* tmp_x, float* dh_next, float* dc_next, float* dx0,
                                  float* const* dw_ih, float* const* dw_hh, float* const* db, float* dlin_w,
                                  float* dlin_b, float* dinput, float* dh0, float* dc0, void* stream) {
  EGAZE_CHECK_ARG(x && h0 && c0 && w_ih && w_hh && lin_w && out && gout && ws_h && ws_c && ws_gates, "lstm_seq_bwd: null input");
  EGAZE_CHECK_ARG(xt && dz && dgates && dh_top && tmp_x && dh_next && dc_next && dw_ih && dw_hh && db && dlin_w && dlin_b,
                  "lstm_seq_bwd: null workspace/output");
  EGAZE_CHECK_ARG(!dinput || dx0, "lstm_seq_bwd: dinput needs the dx0 workspace");
  EGAZE_CHECK_ARG(T > 0 && B > 0, "lstm_seq_bwd: bad T=%d B=%d", T, B);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t sl = (size_t)B * HID;
  const size_t n_all = (size_t)T * sl;
  int eb = (int)((n_all + 255) / 256);
  if (eb > 148 * 8) eb = 148 * 8;
  tanh_kernel<<<eb, 256, 0, st>>>(x, n_all, xt);
  relu_mask_kernel<<<eb, 256, 0, st>>>(gout, out, n_all, dz);
  EGAZE_LAUNCH_CHECK();
  // Linear (LSTMnet.py:36): dW = dz^T . h_top, db = sum dz, dh_top = dz . W
  matmul_tn_kernel<<<dim3(HID / 128, HID), 128, 0, st>>>(dz, sl, HID, ws_h + sl, 2 * sl, HID, T, B, HID, HID, 0, dlin_w);
  col_sum_f32_kernel<<<HID / 128, 128, 0, st>>>(dz, HID, T * B, HID, dlin_b);
  matmul_nn_kernel<<<dim3(HID / 128, T * B), 128, 0, st>>>(dz, lin_w, T * B, HID, HID, dh_top);
  EGAZE_LAUNCH_CHECK();
  if (ghn) EGAZE_CUDA(cudaMemcpyAsync(dh_next, ghn, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else EGAZE_CUDA(cudaMemsetAsync(dh_next, 0, 2 * sl * sizeof(float), st));
  if (gcn) EGAZE_CUDA(cudaMemcpyAsync(dc_next, gcn, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  else EGAZE_CUDA(cudaMemsetAsync(dc_next, 0, 2 * sl * sizeof(float), st));
  const int gb = (int)((sl + 255) / 256);
  for (int t = T - 1; t >= 0; --t) {
    for (int l = 1; l >= 0; --l) {
      const float* gates = ws_gates + ((size_t)t * 2 + l) * sl * 4;
      const float* cc = ws_c + ((size_t)t * 2 + l) * sl;
      const float* cprev = t == 0 ? c0 + l * sl : ws_c + ((size_t)(t - 1) * 2 + l) * sl;
      float* dg = dgates + ((size_t)l * T + t) * sl * 4;
      const float* dh_a = l == 1 ? dh_top + (size_t)t * sl : tmp_x;   // from above: the Linear (layer 1) or layer 1's dx (layer 0)
      lstm_gate_bwd_kernel<<<gb, 256, 0, st>>>(dh_a, dh_next + l * sl, gates, cc, cprev, B, dc_next + l * sl, dg);
      matmul_nn_kernel<<<dim3(HID / 128, B), 128, 0, st>>>(dg, w_hh[l], B, 4 * HID, HID, dh_next + l * sl);   // -> h_{t-1}
      if (l == 1) matmul_nn_kernel<<<dim3(HID / 128, B), 128, 0, st>>>(dg, w_ih[1], B, 4 * HID, HID, tmp_x);   // -> layer 0's h_t
      else if (dinput) matmul_nn_kernel<<<dim3(HID / 128, B), 128, 0, st>>>(dg, w_ih[0], B, 4 * HID, HID, dx0 + (size_t)t * sl);
      EGAZE_LAUNCH_CHECK();
    }
  }
  if (dinput) {
    tanh_bwd_kernel<<<eb, 256, 0, st>>>(xt, dx0, n_all, dinput);
    EGAZE_LAUNCH_CHECK();
  }
  if (dh0) EGAZE_CUDA(cudaMemcpyAsync(dh0, dh_next, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (dc0) EGAZE_CUDA(cudaMemcpyAsync(dc0, dc_next, 2 * sl * sizeof(float), cudaMemcpyDeviceToDevice, st));
  // weight gradients: one pass over all (t, b) rows per matrix
  for (int l = 0; l < 2; ++l) {
    const float* dg = dgates + (size_t)l * T * sl * 4;   // [T][B][2048]
    const size_t g_st = sl * 4, g_sb = 4 * HID;
    dim3 grid(HID / 128, 4 * HID);
    if (l == 0) matmul_tn_kernel<<<grid, 128, 0, st>>>(dg, g_st, g_sb, xt, sl, HID, T, B, 4 * HID, HID, 0, dw_ih[0]);
    else matmul_tn_kernel<<<grid, 128, 0, st>>>(dg, g_st, g_sb, ws_h, 2 * sl, HID, T, B, 4 * HID, HID, 0, dw_ih[1]);
    // recurrent weights: h_{t-1} is the initial state for t = 0 and ws_h[t-1] afterwards
    matmul_tn_kernel<<<grid, 128, 0, st>>>(dg, g_st, g_sb, h0 + l * sl, 0, HID, 1, B, 4 * HID, HID, 0, dw_hh[l]);
    if (T > 1)
      matmul_tn_kernel<<<grid, 128, 0, st>>>(dg + g_st, g_st, g_sb, ws_h + l * sl, 2 * sl, HID, T - 1, B, 4 * HID, HID, 1, dw_hh[l]);
    col_sum_f32_kernel<<<4 * HID / 128, 128, 0, st>>>(dg, 4 * HID, T * B, 4 * HID, db[l]);
    EGAZE_LAUNCH_CHECK();
  }
  return EGAZE_OK;
}
