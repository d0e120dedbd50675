// api.cu — library-wide plumbing: version, thread-local error text, device checks.
#include "common.cuh"

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
