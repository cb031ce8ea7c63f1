// Library-level state of the C ABI (include/ransacflow_b200.h).
#include "common.cuh"

namespace rf {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace rf

#ifndef RF_SOURCE_DIGEST
#define RF_SOURCE_DIGEST "unknown"
#endif
extern "C" int rf_version(void) { return 200; }
// sha256 of csrc/*, the header and the nvcc flags this library was built from (build.py passes it as -DRF_SOURCE_DIGEST):
// _lib.py compares it with the sources next to it and rebuilds / refuses a stale library instead of calling it through
// newer ctypes signatures
extern "C" const char* rf_source_digest(void) { return RF_SOURCE_DIGEST; }
extern "C" const char* rf_last_error_string(void) { return rf::g_err; }
extern "C" uint64_t rf_launch_count(void) { return rf::g_launches.load(); }
