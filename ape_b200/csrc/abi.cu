// abi.cu — status plumbing of the C-ABI (include/ape_b200.h).
#include "common.cuh"
#include <atomic>
#include <stdlib.h>

namespace ape {
namespace {
thread_local char g_err[512] = {0};
std::atomic<uint64_t> g_launches{0};
}  // namespace
char *last_error_buf() { return g_err; }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
bool pdl_enabled() {
  static const bool on = [] {
    const char *e = getenv("APE_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}
}  // namespace ape

extern "C" int ape_abi_version(void) { return APE_ABI_VERSION; }
extern "C" const char *ape_last_error(void) { return ape::last_error_buf(); }
extern "C" uint64_t ape_launch_count(void) { return ape::g_launches.load(std::memory_order_relaxed); }
