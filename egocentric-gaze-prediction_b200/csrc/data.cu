// Input pipeline on the device: what the reference's dataset does per sample on the CPU
// (data/STdatas.py:50-73: cv2.imread of one RGB frame and TWENTY optical-flow JPEGs -- every flow frame is decoded ten times as
// the 10-frame window slides --, uint8 -> float, normalisation, stacking) as three pieces:
//   egaze_jpeg_decode   : nvJPEG (GPU backend) decode of one JPEG into device uint8 (BGR interleaved like cv2.imread, or gray).
//                         The library is loaded with dlopen at first use: libegaze.so has no link-time dependency on it and
//                         reports EGAZE_EUNSUPPORTED where it is missing.  nvJPEG's IDCT / chroma upsampling differ from
//                         libjpeg's by a few grey levels, so this stage is close to, not bit-identical with, cv2.imread.
//   egaze_image_norm    : BGR uint8 HWC -> the reference's normalised NCHW fp32 image (STdatas.py:51-55, bit-identical
//                         arithmetic: x/255, minus mean, divided by std, in BGR order as the reference does).
//   egaze_flow_push / egaze_flow_stack : a ring of the last 10 decoded (flow_x, flow_y) uint8 frames per video; each frame is
//                         decoded and uploaded ONCE, the 20-channel stack [x_n, y_n, x_{n-1}, y_{n-1}, ...] (STdatas.py:18-20,
//                         59-68) is assembled from the ring: ((u/255) - 0.5) / 0.5, NCHW fp32, bit-identical to the reference.
#include "common.cuh"
#include <dlfcn.h>
#include <nvjpeg.h>
#include <mutex>

namespace {

// x: [N][H][W][3] uint8 (BGR) -> out: [N][3][H][W] fp32, out[c] = ((x[c] / 255) - mean[c]) / std[c]
__global__ void image_norm_kernel(const unsigned char* __restrict__ x, int HW, float* __restrict__ out) {
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  const int n = blockIdx.y;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const unsigned char* px = x + ((size_t)n * HW + p) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // same operation order as the reference: float(u8).div(255).sub_(mean).div_(std)
      const float v = __fdiv_rn(__fsub_rn(__fdiv_rn((float)px[c], 255.f), mean[c]), stdv[c]);
      out[((size_t)n * 3 + c) * HW + p] = v;
    }
  }
}

// ring: [V][T][2][H][W] uint8 (T slots per video, slot = frame index mod T); frame: [V][2][H][W] -> slot `slot`
__global__ void flow_push_kernel(unsigned char* __restrict__ ring, const unsigned char* __restrict__ fx,
                                 const unsigned char* __restrict__ fy, int T, int HW, int slot) {
  const int v = blockIdx.y;
  unsigned char* dst = ring + (((size_t)v * T + slot) * 2) * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    dst[p] = fx[(size_t)v * HW + p];
    dst[HW + p] = fy[(size_t)v * HW + p];
  }
}

// out: [V][2T][H][W] fp32; channel 2k = flow_x of frame (newest - k), 2k + 1 = flow_y of it.  Frames older than the first one
// pushed (video start) repeat the oldest available frame.
__global__ void flow_stack_kernel(const unsigned char* __restrict__ ring, int T, int HW, int newest, int count,
                                  float* __restrict__ out) {
  const int v = blockIdx.y, k = blockIdx.z;        // k-th most recent frame
  const int age = k < count ? k : count - 1;
  const int slot = ((newest - age) % T + T) % T;
  const unsigned char* src = ring + (((size_t)v * T + slot) * 2) * HW;
  float* dst = out + ((size_t)v * 2 * T + 2 * k) * HW;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < 2 * HW; p += gridDim.x * blockDim.x) {
    // reference: float(u8).div_(255).sub_(0.5).div_(0.5)
    dst[p] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)src[p], 255.f), 0.5f), 0.5f);
  }
}

// ---- nvJPEG, loaded at run time -------------------------------------------------------------------------------------
struct NvJpegApi {
  void* lib = nullptr;
  nvjpegStatus_t (*CreateSimple)(nvjpegHandle_t*) = nullptr;
  nvjpegStatus_t (*JpegStateCreate)(nvjpegHandle_t, nvjpegJpegState_t*) = nullptr;
  nvjpegStatus_t (*GetImageInfo)(nvjpegHandle_t, const unsigned char*, size_t, int*, nvjpegChromaSubsampling_t*, int*, int*) = nullptr;
  nvjpegStatus_t (*Decode)(nvjpegHandle_t, nvjpegJpegState_t, const unsigned char*, size_t, nvjpegOutputFormat_t, nvjpegImage_t*,
                           cudaStream_t) = nullptr;
  nvjpegHandle_t handle = nullptr;
  nvjpegJpegState_t state = nullptr;
  bool ok = false;
};
NvJpegApi g_nj;
std::once_flag g_nj_once;
std::mutex g_nj_mutex;

void load_nvjpeg() {
  const char* names[] = {"libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
  for (const char* n : names) {
    g_nj.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (g_nj.lib) break;
  }
  if (!g_nj.lib) return;
  g_nj.CreateSimple = (decltype(g_nj.CreateSimple))dlsym(g_nj.lib, "nvjpegCreateSimple");
  g_nj.JpegStateCreate = (decltype(g_nj.JpegStateCreate))dlsym(g_nj.lib, "nvjpegJpegStateCreate");
  g_nj.GetImageInfo = (decltype(g_nj.GetImageInfo))dlsym(g_nj.lib, "nvjpegGetImageInfo");
  g_nj.Decode = (decltype(g_nj.Decode))dlsym(g_nj.lib, "nvjpegDecode");
  if (!g_nj.CreateSimple || !g_nj.JpegStateCreate || !g_nj.GetImageInfo || !g_nj.Decode) return;
  if (g_nj.CreateSimple(&g_nj.handle) != NVJPEG_STATUS_SUCCESS) return;
  if (g_nj.JpegStateCreate(g_nj.handle, &g_nj.state) != NVJPEG_STATUS_SUCCESS) return;
  g_nj.ok = true;
}

}  // namespace

extern "C" int egaze_image_norm(const void* bgr_u8, int N, int H, int W, float* out, void* stream) {
  EGAZE_CHECK_ARG(bgr_u8 && out && N > 0 && H > 0 && W > 0, "image_norm: bad args");
  dim3 grid(ceil_div(H * W, 256) < 64 ? ceil_div(H * W, 256) : 64, N);
  image_norm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)bgr_u8, H * W, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_flow_push(void* ring, const void* flow_x, const void* flow_y, int V, int T, int H, int W, int slot,
                               void* stream) {
  EGAZE_CHECK_ARG(ring && flow_x && flow_y && V > 0 && T > 0 && slot >= 0 && slot < T, "flow_push: bad args");
  dim3 grid(ceil_div(H * W, 256) < 64 ? ceil_div(H * W, 256) : 64, V);
  flow_push_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((unsigned char*)ring, (const unsigned char*)flow_x,
                                                           (const unsigned char*)flow_y, T, H * W, slot);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

extern "C" int egaze_flow_stack(const void* ring, int V, int T, int H, int W, int newest, int count, float* out, void* stream) {
  EGAZE_CHECK_ARG(ring && out && V > 0 && T > 0 && count > 0 && count <= T, "flow_stack: bad args");
  dim3 grid(ceil_div(2 * H * W, 256) < 64 ? ceil_div(2 * H * W, 256) : 64, V, T);
  flow_stack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)ring, T, H * W, newest, count, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}

// width / height / components of a JPEG held in HOST memory (no decode)
extern "C" int egaze_jpeg_info(const void* jpeg, long long nbytes, int* width, int* height, int* components) {
  EGAZE_CHECK_ARG(jpeg && nbytes > 0, "jpeg_info: bad args");
  std::call_once(g_nj_once, load_nvjpeg);
  if (!g_nj.ok) {
    egaze_set_error("nvJPEG is not available (libnvjpeg.so.12 not found or failed to initialise)");
    return EGAZE_EUNSUPPORTED;
  }
  int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
  std::lock_guard<std::mutex> lock(g_nj_mutex);
  const nvjpegStatus_t st = g_nj.GetImageInfo(g_nj.handle, (const unsigned char*)jpeg, (size_t)nbytes, &nc, &ss, ws, hs);
  if (st != NVJPEG_STATUS_SUCCESS) {
    egaze_set_error("nvjpegGetImageInfo failed: status %d", (int)st);
    return EGAZE_EINVAL;
  }
  if (width) *width = ws[0];
  if (height) *height = hs[0];
  if (components) *components = nc;
  return EGAZE_OK;
}

// jpeg: HOST bytes; out: DEVICE uint8 [H][W][3] (BGR interleaved, gray == 0) or [H][W] (gray != 0), pitch = W * (3 | 1)
extern "C" int egaze_jpeg_decode(const void* jpeg, long long nbytes, int gray, void* out, int H, int W, void* stream) {
  EGAZE_CHECK_ARG(jpeg && nbytes > 0 && out && H > 0 && W > 0, "jpeg_decode: bad args");
  std::call_once(g_nj_once, load_nvjpeg);
  if (!g_nj.ok) {
    egaze_set_error("nvJPEG is not available (libnvjpeg.so.12 not found or failed to initialise)");
    return EGAZE_EUNSUPPORTED;
  }
  nvjpegImage_t img;
  memset(&img, 0, sizeof(img));
  img.channel[0] = (unsigned char*)out;
  img.pitch[0] = (size_t)W * (gray ? 1 : 3);
  std::lock_guard<std::mutex> lock(g_nj_mutex);
  const nvjpegStatus_t st = g_nj.Decode(g_nj.handle, g_nj.state, (const unsigned char*)jpeg, (size_t)nbytes,
                                        gray ? NVJPEG_OUTPUT_Y : NVJPEG_OUTPUT_BGRI, &img, (cudaStream_t)stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    egaze_set_error("nvjpegDecode failed: status %d", (int)st);
    return EGAZE_EINVAL;
  }
  return EGAZE_OK;
}

// np.uint8(255 * x) / 255 on the device: the quantisation the reference applies when it writes the SP / AT maps to image files
// between its stages (AT.py:228-230,249-250; lateDataset.py:21-34 reads them back as uint8 / 255).  x in [0, 1].
namespace {
__global__ void quant_u8_kernel(const float* __restrict__ x, size_t n, float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float q = truncf(__fmul_rn(255.f, x[i]));                   // np.uint8() truncates
    out[i] = __fdiv_rn(fminf(fmaxf(q, 0.f), 255.f), 255.f);
  }
}
}  // namespace

extern "C" int egaze_quant_u8(const float* x, long long n, float* out, void* stream) {
  EGAZE_CHECK_ARG(x && out && n > 0, "quant_u8: bad args");
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  quant_u8_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, (size_t)n, out);
  EGAZE_LAUNCH_CHECK();
  return EGAZE_OK;
}
