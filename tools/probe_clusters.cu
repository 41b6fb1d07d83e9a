#include <cstdio>
#include <cuda_runtime.h>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
__global__ void __launch_bounds__(480, 1) dummy(int* p) { extern __shared__ char sm[]; if (p) p[0] = sm[0]; }
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  size_t smem = 7 * 32768 + 1280;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(480); cfg.gridDim = dim3((sms / cs) * cs); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeClusterDimension; a[0].val.clusterDim.x = cs; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("{\"sms\": %d, \"cluster_size\": %d, \"max_active_clusters\": %d, \"sms_used\": %d, \"err\": \"%s\"}\n", sms, cs, n, n * cs, cudaGetErrorString(e));
  }
  // cudaHostRegister speed on touched pageable memory
  for (size_t mb : {256, 1024}) {
    size_t bytes = mb << 20;
    char* p = (char*)aligned_alloc(4096, bytes); memset(p, 1, bytes);
    double t0 = now(); cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault); double t1 = now();
    cudaHostUnregister(p); double t2 = now();
    printf("{\"host_register_mb\": %zu, \"register_s\": %.4f, \"unregister_s\": %.4f, \"GBps_register\": %.2f, \"err\": \"%s\"}\n", mb, t1 - t0, t2 - t1, bytes / (t1 - t0) / 1e9, cudaGetErrorString(e));
    // 4 threads registering 4 separate buffers concurrently
    std::vector<char*> bufs; for (int i = 0; i < 4; ++i) { char* q = (char*)aligned_alloc(4096, bytes / 4); memset(q, 1, bytes / 4); bufs.push_back(q); }
    t0 = now(); { std::vector<std::thread> ts; for (int i = 0; i < 4; ++i) ts.emplace_back([&, i]() { cudaHostRegister(bufs[i], bytes / 4, cudaHostRegisterDefault); }); for (auto& t : ts) t.join(); } t1 = now();
    for (auto q : bufs) { cudaHostUnregister(q); free(q); }
    printf("{\"host_register_mb\": %zu, \"threads\": 4, \"register_s\": %.4f, \"GBps_register\": %.2f}\n", mb, t1 - t0, bytes / (t1 - t0) / 1e9);
    free(p);
  }
  return 0;
}
