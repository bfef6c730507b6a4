// sf_core.cu — error reporting, launch accounting, library identity.
#include "sf_host.h"
#include <atomic>
#include <cstdarg>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace sf {
static thread_local char g_err[512] = {0};
static std::atomic<long long> g_launches{0};

int set_opt(const char* name, int value);
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static const char* const kOptNames[OPT_COUNT] = {"no_pdl", "loss_side", "no_overlap", "no_swiglu_fusion", "gemm_group_m", "gemm_group_m_midk",
                                                 "gemm_group_m_wgrad", "dflash_attn_tc", "gemm_stages", "no_teacher_fusion", "no_loss_stats_fusion", "gemm_wide", "gemm_epi_staged", "no_rope_fusion", "gemm_epi8", "dflash_attn_window"};
static std::atomic<int> g_opt[OPT_COUNT];
static std::atomic<bool> g_opt_init{false};
static void opt_init() {
    if (g_opt_init.load(std::memory_order_acquire)) return;
    for (int i = 0; i < OPT_COUNT; ++i) {
        char name[64] = "SF_";
        size_t n = 3;
        for (const char* c = kOptNames[i]; *c && n + 1 < sizeof(name); ++c) name[n++] = (char)toupper((unsigned char)*c);
        name[n] = 0;
        const char* e = getenv(name);
        g_opt[i].store(e ? atoi(e) : 0);
    }
    g_opt_init.store(true, std::memory_order_release);
}
int opt(Opt o) { opt_init(); return g_opt[o].load(std::memory_order_relaxed); }
int set_opt(const char* name, int value) {
    opt_init();
    for (int i = 0; i < OPT_COUNT; ++i)
        if (!strcmp(name, kOptNames[i])) { g_opt[i].store(value); return 0; }
    return set_error(-22, "unknown option '%s'", name);
}
}  // namespace sf

extern "C" const char* sf_last_error(void) { return sf::g_err; }
extern "C" long long sf_launch_count(void) { return sf::g_launches.load(); }
extern "C" void sf_launch_count_reset(void) { sf::g_launches.store(0); }
extern "C" int sf_debug_option(const char* name, int value) { return sf::set_opt(name, value); }
extern "C" const char* sf_version(void) { return "specforge_b200 0.1 (sm_100a)"; }
