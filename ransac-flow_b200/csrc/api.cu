// Library-level state of the C ABI (include/ransacflow_b200.h).
#include "common.cuh"

namespace rf {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
}  // namespace rf

extern "C" int rf_version(void) { return 100; }
extern "C" const char* rf_last_error_string(void) { return rf::g_err; }
extern "C" uint64_t rf_launch_count(void) { return rf::g_launches.load(); }
