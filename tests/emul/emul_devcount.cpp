// TEST INFRASTRUCTURE: libeph_atomic_emul.so holds the `fix eph/atomic` engine alone; in the product it shares
// libeph_b200.so with the `fix eph` engine, which defines this entry point (csrc/eph_b200.cu).
#include "cuda_runtime.h"
#include "eph_b200.h"
extern "C" int eph_b200_device_count(int *out) {
  if (!out) return EPH_B200_ERR_ARG;
  return cudaGetDeviceCount(out) == cudaSuccess ? EPH_B200_OK : EPH_B200_ERR_NODEVICE;
}
