// egaze-b200 runtime glue: error string, version, device query, tensor-map encoding.
#include "common.cuh"
#include <stdarg.h>
#include <mutex>

static thread_local char g_err[512] = "";

void egaze_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" int egaze_version(void) { return 100; }

extern "C" int egaze_last_error(char* buf, int n) {
  if (!buf || n <= 0) return EGAZE_EINVAL;
  strncpy(buf, g_err, (size_t)n - 1);
  buf[n - 1] = 0;
  return EGAZE_OK;
}

// Number of SMs of the current device (grid sizing for the HBM-bound kernels).
extern "C" int egaze_sm_count(int* out) {
  int dev = 0;
  EGAZE_CUDA(cudaGetDevice(&dev));
  EGAZE_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
  return EGAZE_OK;
}

// The library targets sm_100a only; anything else is an error, never a fallback.
extern "C" int egaze_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  EGAZE_CUDA(cudaGetDevice(&dev));
  EGAZE_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  EGAZE_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    egaze_set_error("egaze kernels are built for sm_100a; device is sm_%d%d", major, minor);
    return EGAZE_EUNSUPPORTED;
  }
  return EGAZE_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled g_encode = nullptr;
static std::once_flag g_encode_once;

static void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_encodeTiled)fn;
}

int egaze_encode_tmap(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes, int elem_bytes) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) {
    egaze_set_error("cuTensorMapEncodeTiled driver entry point unavailable");
    return EGAZE_EDRIVER;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = g_encode(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    egaze_set_error("cuTensorMapEncodeTiled failed: CUresult %d (rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u)", (int)r,
                    rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                    (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                    rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return EGAZE_EDRIVER;
  }
  return EGAZE_OK;
}
