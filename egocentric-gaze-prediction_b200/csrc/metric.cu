// computeAAEAUC on the device (reference utils.py:96-140): per sample
//   predicted = centre of mass of the predicted map; (i, j) = first arg-max of the target (row-major);
//   AAE  = angle between the rays through the two points (camera at distance 112 / tan(pi/6), image centre 112,112);
//   AUC  = 1 - #{z > z[i][j]} / (H*W), z = the sigma-14 Gaussian (scipy gaussian_filter: radius 56, mode 'reflect')
//          of a one-hot map at int(predicted) -- separable, so z[a][b] = gx[a] * gy[b] with the two 1-D responses.
// The reference does this with scipy on the host, one D2H copy of both 224x224 maps per sample; here only the four
// numbers per sample leave the GPU.  The 224 / 112 constants are the reference's (it only supports 224x224 maps).
#include "common.cuh"

namespace {

constexpr int kS = 224, kRadius = 56;

__device__ __forceinline__ double block_sum_d(double v, double* sh) {
  v = warp_sum_d(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += sh[i];
  return t;
}

// The (at most two) filter taps of scipy.ndimage.correlate1d(delta at p0, w, mode='reflect') that land on position a: the
// direct hit and, within 56 pixels of a border, the mirror image (reflect(x) = -x-1 below 0, 2n-x-1 above n-1).
__device__ __forceinline__ void taps1d(const double* w, int a, int p0, double& t1, double& t2) {
  t1 = 0.0; t2 = 0.0;
  const int d0 = p0 - a;
  if (d0 >= -kRadius && d0 <= kRadius) t1 = w[d0 + kRadius];
  const int d1 = -p0 - 1 - a;
  if (d1 >= -kRadius && d1 <= kRadius && a + d1 < 0) t2 = w[d1 + kRadius];
  const int d2 = 2 * kS - p0 - 1 - a;
  if (d2 >= -kRadius && d2 <= kRadius && a + d2 >= kS) t2 = w[d2 + kRadius];
}

__global__ void __launch_bounds__(256) aae_auc_kernel(const float* __restrict__ out, const float* __restrict__ tgt,
                                                      const double* __restrict__ weights, double* __restrict__ res) {
  __shared__ double sh[8], gy_red[8];
  __shared__ double w[2 * kRadius + 1], gx[kS], gy1[kS], gy2[kS];
  __shared__ float s_max[8];
  __shared__ int s_idx[8];
  const int b = blockIdx.x;
  const float* o = out + (size_t)b * kS * kS;
  const float* t = tgt + (size_t)b * kS * kS;
  // centre of mass (fp64 sums) and first arg-max of the target
  double sv = 0.0, si = 0.0, sj = 0.0;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  for (int idx = threadIdx.x; idx < kS * kS; idx += blockDim.x) {
    const double v = (double)o[idx];
    sv += v; si += v * (double)(idx / kS); sj += v * (double)(idx % kS);
    const float tv = t[idx];
    if (tv > best) { best = tv; best_i = idx; }   // strictly greater: the smallest index of the maximum survives
  }
  sv = block_sum_d(sv, sh); si = block_sum_d(si, sh); sj = block_sum_d(sj, sh);
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
    if (ov > best || (ov == best && oi < best_i)) { best = ov; best_i = oi; }
  }
  if ((threadIdx.x & 31) == 0) { s_max[threadIdx.x >> 5] = best; s_idx[threadIdx.x >> 5] = best_i; }
  // Gaussian weights: the host passes scipy's own table (exp(-0.5 d^2 / sigma^2) normalised by its NumPy sum).  Pixels on the
  // circle through the gaze point tie with it in exact arithmetic; which side of `>` they fall on is decided by the last bit
  // of these weights, so they must be the very same doubles the reference multiplies.
  for (int d = threadIdx.x; d <= 2 * kRadius; d += blockDim.x) w[d] = weights[d];
  __syncthreads();
  float bm = s_max[0];
  int bi = s_idx[0];
  for (int wv = 1; wv < 8; ++wv)
    if (s_max[wv] > bm || (s_max[wv] == bm && s_idx[wv] < bi)) { bm = s_max[wv]; bi = s_idx[wv]; }
  const int gi = bi / kS, gj = bi % kS;
  const double pi_ = si / sv, pj_ = sj / sv;   // NaN when the map sums to zero, like scipy's centre_of_mass
  __syncthreads();
  int p0 = (int)pi_, p1 = (int)pj_;
  p0 = min(max(p0, 0), kS - 1);
  p1 = min(max(p1, 0), kS - 1);
  // scipy filters axis 0 first: column p1 of the intermediate holds gx[a] = t1 + t2; the axis-1 pass then forms
  // z[a][b] = gx[a]*u1 + gx[a]*u2 (two products, one sum -- reproduced operation for operation, no FMA contraction)
  for (int a = threadIdx.x; a < kS; a += blockDim.x) {
    double t1, t2;
    taps1d(w, a, p0, t1, t2);
    gx[a] = __dadd_rn(t1, t2);
    taps1d(w, a, p1, t1, t2);
    gy1[a] = t1; gy2[a] = t2;
  }
  __syncthreads();
  // the reference compares the NORMALISED maps, z = (z - min z) / max(z - min z): the two roundings can merge neighbours, so
  // they are applied here as well
  auto zval = [&](int a, int bb) { return __dadd_rn(__dmul_rn(gx[a], gy1[bb]), __dmul_rn(gx[a], gy2[bb])); };
  double zmin = 1e300, zmax = -1e300;
  for (int idx = threadIdx.x; idx < kS * kS; idx += blockDim.x) {
    const double z = zval(idx / kS, idx % kS);
    zmin = fmin(zmin, z);
    zmax = fmax(zmax, z);
  }
  for (int off = 16; off > 0; off >>= 1) {
    zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, off));
    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, off));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = zmin; gy_red[threadIdx.x >> 5] = zmax; }
  __syncthreads();
  for (int i = 0; i < 8; ++i) { zmin = fmin(zmin, sh[i]); zmax = fmax(zmax, gy_red[i]); }
  const double zden = __dsub_rn(zmax, zmin);
  const double thr = __ddiv_rn(__dsub_rn(zval(gi, gj), zmin), zden);
  double cnt = 0.0;
  for (int idx = threadIdx.x; idx < kS * kS; idx += blockDim.x)
    if (__ddiv_rn(__dsub_rn(zval(idx / kS, idx % kS), zmin), zden) > thr) cnt += 1.0;
  cnt = block_sum_d(cnt, sh);
  if (threadIdx.x == 0) {
    const double d = 112.0 / tan(3.14159265358979323846 / 6.0);
    const double r1[3] = {pi_ - 112.0, pj_ - 112.0, d}, r2[3] = {(double)gi - 112.0, (double)gj - 112.0, d};
    const double cx = r1[1] * r2[2] - r1[2] * r2[1], cy = r1[2] * r2[0] - r1[0] * r2[2], cz = r1[0] * r2[1] - r1[1] * r2[0];
    const double nrm = sqrt(cx * cx + cy * cy + cz * cz), dot = r1[0] * r2[0] + r1[1] * r2[1] + r1[2] * r2[2];
    res[(size_t)b * 4 + 0] = atan2(nrm, dot) * (180.0 / 3.14159265358979323846);
    res[(size_t)b * 4 + 1] = 1.0 - cnt / (double)(kS * kS);
    res[(size_t)b * 4 + 2] = (double)gi;
    res[(size_t)b * 4 + 3] = (double)gj;
  }
}

}  // namespace

// out / tgt: [B][224][224] fp32 on the device; res: [B][4] fp64 = (AAE in degrees, AUC, gaze row, gaze column).
// weights: [113] fp64 on the device = scipy.ndimage's sigma-14 Gaussian kernel (radius 56), built on the host with NumPy.
extern "C" int egaze_aae_auc(const float* out, const float* tgt, int B, int H, int W, const double* weights, double* res,
                             void* stream) {
  EGAZE_CHECK_ARG(out && tgt && weights && res && B > 0, "aae_auc: bad args");
  EGAZE_CHECK_ARG(H == kS && W == kS, "aae_auc: the reference metric is defined for 224x224 maps only (utils.py:107,113), got %dx%d", H, W);
  aae_auc_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(out, tgt, weights, res);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
