// Library-level entry points of libgsb200 (version, error string, launch counter).
#include "common.cuh"

namespace gsb200 {
thread_local char g_err[512] = "";
unsigned long long g_launches = 0;
}  // namespace gsb200

extern "C" int gsb200_abi_version(void) { return GSB200_ABI_VERSION; }
extern "C" const char *gsb200_last_error(void) { return gsb200::g_err; }
extern "C" uint64_t gsb200_launch_count(void) { return gsb200::g_launches; }
