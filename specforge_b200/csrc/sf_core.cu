// sf_core.cu — error reporting, launch accounting, library identity.
#include "sf_host.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace sf {
static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace sf

extern "C" const char* sf_last_error(void) { return sf::g_err; }
extern "C" long long sf_launch_count(void) { return sf::g_launches.load(); }
extern "C" void sf_launch_count_reset(void) { sf::g_launches.store(0); }
extern "C" const char* sf_version(void) { return "specforge_b200 0.1 (sm_100a)"; }
