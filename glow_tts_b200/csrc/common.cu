// common.cu -- error text, ABI version, launch counter.
#include "common.cuh"

#include <map>
#include <mutex>
#include <string>
#include <vector>

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

// ---- per-launch device timing ------------------------------------------------
struct ProfRec { const char *name; cudaEvent_t e0, e1; };
static std::mutex g_prof_mu;
static std::vector<ProfRec *> g_prof;
static std::atomic<int> g_prof_on{0};
constexpr size_t kProfMax = 200000;

ProfScope::ProfScope(const char *n, cudaStream_t s) : name(n), st(s), slot(nullptr)
{
    if (!g_prof_on.load(std::memory_order_relaxed)) return;
    ProfRec *r = new ProfRec{n, nullptr, nullptr};
    if (cudaEventCreate(&r->e0) != cudaSuccess || cudaEventCreate(&r->e1) != cudaSuccess) { delete r; return; }
    cudaEventRecord(r->e0, s);
    slot = r;
}
ProfScope::~ProfScope()
{
    if (!slot) return;
    ProfRec *r = (ProfRec *)slot;
    cudaEventRecord(r->e1, st);
    std::lock_guard<std::mutex> lock(g_prof_mu);
    if (g_prof.size() < kProfMax) g_prof.push_back(r);
    else { cudaEventDestroy(r->e0); cudaEventDestroy(r->e1); delete r; }
}

}  // namespace glow

extern "C" {

int glow_abi_version(void) { return 2; }

const char *glow_last_error(void) { return glow::err_buf(); }

uint64_t glow_launch_count(void) { return glow::g_launches.load(std::memory_order_relaxed); }

int glow_prof_enable(int on)
{
    glow::g_prof_on.store(on ? 1 : 0, std::memory_order_relaxed);
    return GLOW_OK;
}

int glow_prof_report(char *buf, size_t buf_bytes)
{
    using namespace glow;
    GLOW_REQUIRE(buf && buf_bytes > 0, GLOW_ERR_INVALID, "prof_report: null buffer");
    std::vector<ProfRec *> recs;
    {
        std::lock_guard<std::mutex> lock(g_prof_mu);
        recs.swap(g_prof);
    }
    std::map<std::string, std::pair<long, double>> agg;
    for (ProfRec *r : recs) {
        float ms = 0.f;
        if (cudaEventSynchronize(r->e1) == cudaSuccess && cudaEventElapsedTime(&ms, r->e0, r->e1) == cudaSuccess) {
            auto &a = agg[r->name];
            a.first += 1;
            a.second += ms;
        }
        cudaEventDestroy(r->e0);
        cudaEventDestroy(r->e1);
        delete r;
    }
    size_t pos = 0;
    buf[0] = 0;
    for (auto &kv : agg) {
        int n = snprintf(buf + pos, buf_bytes - pos, "%s %ld %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (n < 0 || (size_t)n >= buf_bytes - pos) break;
        pos += (size_t)n;
    }
    return GLOW_OK;
}

}  // extern "C"
