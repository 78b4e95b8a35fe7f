// Library-level entry points of libavfe.so (version, error strings, launch counter).
#include "avfe_common.cuh"

namespace avfe {
std::atomic<uint64_t> g_launch_count{0};
}

extern "C" int avfe_version(void) { return 0 * 10000 + 1 * 100 + 0; }

extern "C" const char* avfe_strerror(int status) {
  switch (status) {
    case AVFE_OK: return "ok";
    case AVFE_ERR_INVALID_ARG: return "invalid argument (null pointer, negative size or unsupported shape)";
    case AVFE_ERR_UNSUPPORTED: return "unsupported parameter combination";
    case AVFE_ERR_WORKSPACE: return "workspace missing, too small or misaligned";
    case AVFE_ERR_CUDA: return "CUDA runtime call or kernel launch failed";
    case AVFE_ERR_ALIGNMENT: return "pointer alignment requirement not met";
    default: return "unknown avfe status";
  }
}

extern "C" uint64_t avfe_launch_count(void) {
  return avfe::g_launch_count.load(std::memory_order_relaxed);
}
