// Error plumbing, launch counter and the optional per-kernel event profiler of the C ABI.
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace amid {
static thread_local char g_err[512] = "";
int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof{0};
struct Rec { const char* name; cudaEvent_t a, b; };
static std::mutex g_mu;
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;
static thread_local const char* t_name = nullptr;
static thread_local cudaStream_t t_stream = nullptr;
static thread_local cudaEvent_t t_a = nullptr;

static cudaEvent_t get_event() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
void prof_begin(const char* name, cudaStream_t s) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    if (!g_prof.load(std::memory_order_relaxed)) return;
    t_name = name; t_stream = s; t_a = get_event();
    if (t_a) cudaEventRecord(t_a, s);
}
void prof_end() {
    if (!t_name) return;
    cudaEvent_t b = get_event();
    if (b) cudaEventRecord(b, t_stream);
    { std::lock_guard<std::mutex> lk(g_mu); g_recs.push_back({t_name, t_a, b}); }
    t_name = nullptr;
}
}  // namespace amid

using namespace amid;

extern "C" const char* amid_last_error(void) { return amid::g_err; }
extern "C" int amid_version(void) { return 100; }
extern "C" int64_t amid_launch_count(void) { return (int64_t)g_launches.load(); }

// on != 0: start recording (discarding earlier records); on == 0: stop recording.
extern "C" int amid_profile_enable(int32_t on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (on) {
        for (auto& r : g_recs) { if (r.a) g_pool.push_back(r.a); if (r.b) g_pool.push_back(r.b); }
        g_recs.clear();
    }
    g_prof.store(on ? 1 : 0);
    return 0;
}
// Synchronises the device.  Writes "name count total_ms\n" lines (kernel order of first
// appearance) into buf; returns the number of bytes needed (excluding the terminator).
extern "C" int64_t amid_profile_report_host_sync(char* buf, int64_t cap) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_mu);
    std::vector<std::string> order;
    std::map<std::string, std::pair<long long, double>> agg;
    for (auto& r : g_recs) {
        float ms = 0.f;
        if (!r.a || !r.b || cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) continue;
        auto it = agg.find(r.name);
        if (it == agg.end()) { order.push_back(r.name); it = agg.emplace(r.name, std::make_pair(0ll, 0.0)).first; }
        it->second.first += 1;
        it->second.second += ms;
    }
    std::string out;
    char line[256];
    for (auto& n : order) {
        snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), agg[n].first, agg[n].second);
        out += line;
    }
    if (buf && cap > 0) {
        size_t k = out.size() < (size_t)(cap - 1) ? out.size() : (size_t)(cap - 1);
        memcpy(buf, out.data(), k);
        buf[k] = 0;
    }
    return (int64_t)out.size();
}
