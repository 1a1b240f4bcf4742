// Library-wide state of the C ABI: last error (thread-local), launch counter.
#include <atomic>
#include "common.cuh"
#include "../../include/ursm_b200.h"

namespace ursm {
static thread_local std::string g_error;
static std::atomic<long long> g_launches{0};
void set_error(const std::string& msg) { g_error = msg; }
void count_launch(int n) { g_launches += n; }
}  // namespace ursm

extern "C" {
const char* ursm_last_error(void) { return ursm::g_error.c_str(); }
int ursm_version(void) { return 100; }
int64_t ursm_launch_count(void) { return ursm::g_launches.load(); }
void ursm_launch_count_reset(void) { ursm::g_launches = 0; }
}
