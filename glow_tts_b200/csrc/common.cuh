// common.cuh -- shared plumbing for libglowcore.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/glowcore.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libglowcore is written for sm_100a (B200) only"
#endif

namespace glow {

// thread-local last-error text (glow_last_error)
char *err_buf();
int   fail(int code, const char *fmt, ...);
void  count_launch(int n = 1);

#define GLOW_CHECK_CUDA(expr)                                                        \
    do {                                                                             \
        cudaError_t _e = (expr);                                                     \
        if (_e != cudaSuccess)                                                       \
            return ::glow::fail(GLOW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,      \
                                cudaGetErrorString(_e), __FILE__, __LINE__);         \
    } while (0)

#define GLOW_CHECK_LAUNCH(name)                                                      \
    do {                                                                             \
        cudaError_t _e = cudaGetLastError();                                         \
        if (_e != cudaSuccess)                                                       \
            return ::glow::fail(GLOW_ERR_CUDA, "launch of %s failed: %s", name,      \
                                cudaGetErrorString(_e));                             \
        ::glow::count_launch();                                                      \
    } while (0)

#define GLOW_REQUIRE(cond, code, ...)                                                \
    do {                                                                             \
        if (!(cond)) return ::glow::fail(code, __VA_ARGS__);                         \
    } while (0)

// Optional per-launch device timing (glow_prof_enable / glow_prof_report): when enabled,
// a ProfScope around a launch records a CUDA event before and after it on the launch
// stream; the report sums the elapsed time per name.  Off by default (zero overhead).
struct ProfScope {
    const char *name;
    cudaStream_t st;
    void *slot;
    ProfScope(const char *name, cudaStream_t st);
    ~ProfScope();
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;   // B200
constexpr int kMaxDevices = 64;  // per-device tables (handles, side streams, function attributes)

}  // namespace glow
