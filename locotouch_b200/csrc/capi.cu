// Library-level entry points of the C ABI (include/locotouch_b200.h): version, error text, device query cache.
#include <string.h>

#include "lt_common.cuh"

namespace lt {

static thread_local char g_last_error[256] = "";

void set_last_cuda_error(cudaError_t e) {
  strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
}

int sm_count() {
  static int cached = 0;
  if (cached > 0) return cached;
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return 148;  // B200
  }
  cached = n;
  return n;
}

}  // namespace lt

extern "C" int lt_abi_version(void) { return LT_ABI_VERSION; }

extern "C" const char* lt_error_string(int status) {
  switch (status) {
    case LT_OK: return "ok";
    case LT_ERR_INVALID_ARG: return "invalid argument (null pointer, bad size, misaligned buffer or unsupported dimension)";
    case LT_ERR_CUDA: return "CUDA launch failed";
    case LT_ERR_WORKSPACE: return "workspace too small";
    case LT_ERR_UNSUPPORTED: return "unsupported configuration";
    default: return "unknown status";
  }
}

extern "C" const char* lt_last_cuda_error(void) { return lt::g_last_error; }

extern "C" int64_t lt_struct_size(int which) {
  switch (which) {
    case 0: return (int64_t)sizeof(LtGatherArgs);
    case 1: return (int64_t)sizeof(LtPpoLossArgs);
    case 2: return (int64_t)sizeof(LtTaxelArgs);
    case 3: return (int64_t)sizeof(LtMdpArgs);
    case 4: return (int64_t)sizeof(LtGaitState);
    case 5: return (int64_t)sizeof(LtGaitParams);
    case 6: return (int64_t)sizeof(LtTaxelForceArgs);
    case 7: return (int64_t)sizeof(LtCommandRanges);
    case 8: return (int64_t)sizeof(LtCommandArgs);
    case 9: return (int64_t)sizeof(LtVelCurriculumArgs);
    case 10: return (int64_t)sizeof(LtPpoHeadsArgs);
    case 11: return (int64_t)sizeof(LtStudentCnnArgs);
    case 12: return (int64_t)sizeof(LtMlp3Net);
    default: return -1;
  }
}
