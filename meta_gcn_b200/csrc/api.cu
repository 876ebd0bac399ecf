// Library-level entry points of libmgcn.so (version, error strings, launch counter).
#include "common.cuh"

namespace mgcn {
std::atomic<long long> g_launch_count{0};
}

extern "C" int mgcn_version(void) { return MGCN_VERSION; }

extern "C" const char* mgcn_error_string(int code) {
  switch (code) {
    case MGCN_OK: return "ok";
    case MGCN_ERR_NULL: return "mgcn: required pointer is NULL";
    case MGCN_ERR_RANGE: return "mgcn: N or E out of the supported int32 range";
    case MGCN_ERR_SHAPE: return "mgcn: unsupported width, stride or enum value";
    case MGCN_ERR_ALIGN: return "mgcn: pointer not 16-byte aligned";
    case MGCN_ERR_WORKSPACE: return "mgcn: workspace too small";
    default: break;
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "mgcn: unknown error";
}

extern "C" int64_t mgcn_launch_count(void) {
  return mgcn::g_launch_count.load(std::memory_order_relaxed);
}

extern "C" void mgcn_reset_launch_count(void) {
  mgcn::g_launch_count.store(0, std::memory_order_relaxed);
}
