// Library plumbing: error reporting, device-property cache, TMA tensor-map encoding.
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace clipn {

static thread_local char g_err[512] = "";

int set_error_cuda(cudaError_t e, const char* expr, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d: %s", static_cast<int>(e), cudaGetErrorString(e), file,
           line, expr);
  return CLIPN_ERR_CUDA;
}
int set_error_arg(const char* msg, const char* file, int line) {
  snprintf(g_err, sizeof(g_err), "argument error at %s:%d: %s", file, line, msg);
  return CLIPN_ERR_ARG;
}

struct DevInfo {
  int sms = 0, major = 0, minor = 0;
  bool ok = false;
};
static DevInfo g_dev[64];
static std::mutex g_dev_mu;

static const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!g_dev[dev].ok) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (!g_dev[dev].ok) {
      cudaDeviceProp prop;
      if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return nullptr;
      g_dev[dev].sms = prop.multiProcessorCount;
      g_dev[dev].major = prop.major;
      g_dev[dev].minor = prop.minor;
      g_dev[dev].ok = true;
    }
  }
  return &g_dev[dev];
}

int num_sms() {
  const DevInfo* d = dev_info();
  return d ? d->sms : 1;
}

// cuTensorMapEncodeTiled is fetched through the runtime so the library does not link libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static EncodeTiledFn get_encode() {
  if (g_encode == nullptr) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_encode == nullptr) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
          qres != cudaDriverEntryPointSuccess)
        return nullptr;
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
  }
  return g_encode;
}

int make_tmap_2d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                 uint64_t row_pitch_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return set_error_arg("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_pitch_bytes & 15) != 0)
    return set_error_arg("TMA operand must be 16-byte aligned with a 16-byte multiple row pitch", __FILE__, __LINE__);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_pitch_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled failed (%d): base=%p inner=%llu outer=%llu pitch=%llu box=%ux%u", static_cast<int>(r),
             base, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_pitch_bytes, box_inner,
             box_outer);
    return CLIPN_ERR_CUDA;
  }
  return CLIPN_OK;
}

int make_tmap_3d(CUtensorMap* out, const void* base, int elem_bytes, uint64_t d0, uint64_t d1, uint64_t d2,
                 uint64_t stride1_bytes, uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2,
                 int swizzle_bytes) {
  EncodeTiledFn enc = get_encode();
  if (enc == nullptr) return set_error_arg("cuTensorMapEncodeTiled entry point unavailable", __FILE__, __LINE__);
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (stride1_bytes & 15) != 0 || (stride2_bytes & 15) != 0)
    return set_error_arg("TMA operand must be 16-byte aligned with 16-byte multiple strides", __FILE__, __LINE__);
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = enc(out, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(g_err, sizeof(g_err),
             "cuTensorMapEncodeTiled(3D) failed (%d): base=%p dims=%llux%llux%llu strides=%llu,%llu box=%ux%ux%u",
             static_cast<int>(r), base, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box0, box1, box2);
    return CLIPN_ERR_CUDA;
  }
  return CLIPN_OK;
}

}  // namespace clipn

extern "C" int clipn_version(void) { return 100; }
extern "C" const char* clipn_last_error(void) { return clipn::g_err; }
extern "C" int clipn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  const clipn::DevInfo* d = clipn::dev_info();
  if (d == nullptr) return clipn::set_error_arg("no CUDA device", __FILE__, __LINE__);
  if (sm_count) *sm_count = d->sms;
  if (cc_major) *cc_major = d->major;
  if (cc_minor) *cc_minor = d->minor;
  return CLIPN_OK;
}
