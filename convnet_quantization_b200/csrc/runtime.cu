// Host-side plumbing shared by every b200q entry point: error text, device checks, TMA descriptor encoding.
#include <stdarg.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace b200q {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("CUDA error %s (%s) at %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
  return B200Q_ERR_CUDA;
}

static std::atomic<uint64_t> g_launches{0};

// Every kernel launch site reports here: counts the launch and converts a launch error into a status code.
int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return check_cuda(cudaGetLastError(), what);
}

static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
void pdl_set(bool on) { g_pdl = on; }

// A graph replay launches kernels that no launch site sees: the executor reports them here.
void note_graph_replay(int kernels) { g_launches.fetch_add((uint64_t)kernels, std::memory_order_relaxed); }

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

int ensure_dynamic_smem(const void* kernel, int bytes, uint64_t* mask) {
  int dev = 0;
  int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
  if (rc) return rc;
  const uint64_t bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0;  // devices >= 64: set the attribute on every launch
  if (bit && (*mask & bit)) return 0;
  rc = check_cuda(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes),
                  "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
  if (rc) return rc;
  *mask |= bit;
  return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (%s)", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// dims[0] is the contiguous dimension; strides_bytes[i] is the stride of dims[i+1] (rank-1 entries).
// Element type is always uint8 here; OOB elements read as 0.
int encode_tensor_map(CUtensorMap* out, const void* gptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return B200Q_ERR_DRIVER;
  cuuint64_t gdim[5];
  cuuint64_t gstr[5];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                          : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, (cuuint32_t)rank, const_cast<void*>(gptr), gdim, gstr, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dim0 %llu, box0 %u, swizzle %d)", (int)r, rank,
              (unsigned long long)dims[0], box[0], swizzle_bytes);
    return B200Q_ERR_DRIVER;
  }
  return 0;
}

}  // namespace b200q

extern "C" const char* b200q_last_error(void) { return b200q::g_err; }
// 4: b200q_linear_dynamic re-specified (tensor cores, scratch_bytes), + b200q_aminmax, b200q_histc, b200q_lut_u8,
//    b200q_graph_*; b200q_conv12_fused and b200q_conv3x3_simt moved to the development library (-DB200Q_DEV)
extern "C" int b200q_abi_version(void) { return 4; }
extern "C" uint64_t b200q_launch_count(void) { return b200q::g_launches.load(std::memory_order_relaxed); }
