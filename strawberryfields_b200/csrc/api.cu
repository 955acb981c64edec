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
}
