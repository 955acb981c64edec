// Library-level entry points of libb200fock.so.
#include "common.cuh"

namespace b200 {
thread_local char g_err[512] = "";
long long g_launches = 0;
}  // namespace b200

extern "C" {
int b200_version(void) { return 100; }
const char* b200_last_error(void) { return b200::g_err; }
int64_t b200_packed_size(int D) { return D >= 1 ? (int64_t)b200::packed_size(D) : 0; }
int64_t b200_launch_count(void) { return b200::g_launches; }
void b200_reset_launch_count(void) { b200::g_launches = 0; }

int b200_enable_peer_access(int peer_device) {
  int dev = 0, can = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess && dev == peer_device) return 0;
  if (e == cudaSuccess) e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return b200::fail((int)e, "enable_peer_access: %s", cudaGetErrorString(e));
  if (!can) return b200::fail(B200_EUNSUPPORTED, "%s", "enable_peer_access: no peer path between the devices");
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();  // clear the sticky-free error state
    return 0;
  }
  if (e != cudaSuccess) return b200::fail((int)e, "enable_peer_access: %s", cudaGetErrorString(e));
  return 0;
}
}
