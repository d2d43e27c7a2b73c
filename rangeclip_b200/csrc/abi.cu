// C-ABI bookkeeping: version, per-thread error text, launch counter.
#include "common.cuh"

namespace rc {
thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
}  // namespace rc

extern "C" int rc_abi_version(void) { return RC_ABI_VERSION; }
extern "C" const char* rc_last_error(void) { return rc::g_err; }
extern "C" int64_t rc_launch_count(void) { return (int64_t)rc::g_launches.load(); }

#ifdef RC_BRINGUP
// Bring-up query: how many clusters of `cluster_size` CTAs (threads, dynamic shared memory as given) the device can
// keep resident at once -- 148 SMs need not tile into clusters larger than 2 (GPC sizes), which decides whether a
// 4-CTA cluster with multicast operand loads is worth building.
namespace rc { __global__ void occupancy_probe_kernel(int* p) { if (p != nullptr && threadIdx.x == 0 && blockIdx.x == 0) *p = 0; } }
extern "C" int rc_debug_max_active_clusters(int cluster_size, int threads, int smem_bytes) {
  using namespace rc;
  RC_REQUIRE(cluster_size >= 1 && cluster_size <= 16 && threads >= 32 && threads <= 1024 && smem_bytes >= 0, "rc_debug_max_active_clusters: bad argument");
  cudaError_t e = cudaFuncSetAttribute(occupancy_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_debug_max_active_clusters: smem opt-in: %s", cudaGetErrorString(e));
  if (cluster_size > 8) {
    e = cudaFuncSetAttribute(occupancy_probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_debug_max_active_clusters: non-portable cluster size: %s", cudaGetErrorString(e));
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(num_sms() / cluster_size * cluster_size), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeClusterDimension;
  attr.val.clusterDim.x = (unsigned)cluster_size; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
  cfg.attrs = &attr; cfg.numAttrs = 1;
  int n = 0;
  e = cudaOccupancyMaxActiveClusters(&n, occupancy_probe_kernel, &cfg);
  if (e != cudaSuccess) return fail(RC_ERR_CUDA, "rc_debug_max_active_clusters: %s", cudaGetErrorString(e));
  return n;
}
#endif  // RC_BRINGUP
