// C-ABI bookkeeping: version, per-thread error text, launch counter.
#include "common.cuh"

namespace rc {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace rc

extern "C" int rc_abi_version(void) { return RC_ABI_VERSION; }
extern "C" const char* rc_last_error(void) { return rc::g_err; }
extern "C" int64_t rc_launch_count(void) { return (int64_t)rc::g_launches.load(); }
