// api.cu — library-wide plumbing: version, thread-local error text, device checks.
#include "common.cuh"

#include <map>
#include <mutex>
#include <utility>

namespace tauv {

char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

int num_sms() {
  // Re-queried per call (cheap, cached by the runtime); no mutable globals so the library stays
  // re-entrant across devices.  Falls back to the B200 count when no device is visible so that
  // workspace queries work on a CPU-only build box.
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) {
    (void)cudaGetLastError();
    return 148;
  }
  return sms;
}

// Opt a kernel in to `bytes` of dynamic shared memory.  The limit is only ever RAISED, under a lock, per (kernel,
// device): two host threads that need different sizes can never lower it under each other between one thread's
// cudaFuncSetAttribute and its launch.  The fast path (limit already high enough) takes the lock but makes no CUDA call.
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> limit;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lock(mu);
  size_t& cur = limit[std::make_pair(func, dev)];
  if (cur >= bytes) return cudaSuccess;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

}  // namespace tauv

extern "C" int tauv_version(void) { return TAUV_B200_VERSION; }

extern "C" const char* tauv_last_error(void) { return tauv::last_error_buf(); }

extern "C" int tauv_check_device(void) {
  int dev = 0, major = 0;
  TAUV_CUDA(cudaGetDevice(&dev));
  TAUV_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  TAUV_REQUIRE(major == 10, TAUV_E_ARCH, "device %d has compute capability %d.x; libtauv_b200 is sm_100a only", dev, major);
  return 0;
}
