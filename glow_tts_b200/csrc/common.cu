// common.cu -- error text, ABI version, launch counter.
#include "common.cuh"

namespace glow {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

}  // namespace glow

extern "C" {

int glow_abi_version(void) { return 1; }

const char *glow_last_error(void) { return glow::err_buf(); }

uint64_t glow_launch_count(void) { return glow::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
