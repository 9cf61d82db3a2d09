// Library-level entry points of the C ABI (include/esn.h).
#include "esn_common.cuh"

std::atomic<long long> g_esn_launches{0};

extern "C" int esn_version(void) { return ESN_VERSION; }

extern "C" const char* esn_strerror(int code) {
  switch (code) {
    case ESN_OK: return "ok";
    case ESN_ERR_BAD_ARG: return "bad argument (null pointer or dtype/layout combination)";
    case ESN_ERR_BAD_SHAPE: return "tensor shapes inconsistent with the op";
    case ESN_ERR_UNSUPPORTED: return "configuration not supported by this kernel family";
    case ESN_ERR_CUDA: return "CUDA runtime/driver call failed";
    case ESN_ERR_ALIGN: return "pointer or stride alignment requirement not met";
    default: return "unknown error";
  }
}

extern "C" int64_t esn_launch_count(void) { return (int64_t)g_esn_launches.load(); }
extern "C" void esn_launch_count_reset(void) { g_esn_launches.store(0); }
