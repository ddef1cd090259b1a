// Optional per-kernel CUDA-event timers (ssg_profile_*): bench.py uses them to time the dominant kernel
// on the stream it is launched on, live inside the timed region (events cost ~1 us each; off by default).
#include <map>
#include <string>
#include <vector>
#include <string.h>

#include "common.cuh"

namespace ssg {

struct ProfRec { const char* name; cudaEvent_t e0, e1; };
struct ProfAgg { double ms = 0.0; long long launches = 0; };

static bool g_on = false;
static int g_suspend = 0;      // > 0 while a stream capture is recording launches (events recorded there cannot be timed)
static std::vector<ProfRec> g_open;
static std::vector<cudaEvent_t> g_pool;
static std::map<std::string, ProfAgg> g_agg;
static std::vector<std::string> g_names;   // stable order for ssg_profile_entry

static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

ProfScope::ProfScope(const char* name, cudaStream_t st) : name_(name), st_(st), e0_(nullptr) {
    if (!g_on || g_suspend) return;
    e0_ = get_event();
    cudaEventRecord((cudaEvent_t)e0_, st);
}
ProfScope::~ProfScope() {
    if (!e0_) return;
    cudaEvent_t e1 = get_event();
    cudaEventRecord(e1, st_);
    g_open.push_back(ProfRec{name_, (cudaEvent_t)e0_, e1});
}

void prof_suspend(bool on) { g_suspend += on ? 1 : -1; }

}  // namespace ssg

using namespace ssg;

extern "C" int ssg_profile_enable(int on) { g_on = on != 0; return SSG_OK; }

extern "C" int ssg_profile_reset(void) {
    for (auto& r : g_open) { g_pool.push_back(r.e0); g_pool.push_back(r.e1); }
    g_open.clear();
    g_agg.clear();
    g_names.clear();
    return SSG_OK;
}

extern "C" int ssg_profile_collect(void) {
    SSG_CUDA_TRY(cudaDeviceSynchronize());
    for (auto& r : g_open) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
            auto it = g_agg.find(r.name);
            if (it == g_agg.end()) { g_names.push_back(r.name); it = g_agg.emplace(r.name, ProfAgg()).first; }
            it->second.ms += ms;
            it->second.launches += 1;
        }
        g_pool.push_back(r.e0);
        g_pool.push_back(r.e1);
    }
    g_open.clear();
    return (int)g_names.size();
}

extern "C" int ssg_profile_entry(int i, char* name, size_t cap, double* ms, long long* launches) {
    if (i < 0 || i >= (int)g_names.size() || !name || cap == 0)
        return ssg_set_error(SSG_ERR_INVALID, "profile_entry: index %d out of range", i);
    const std::string& n = g_names[i];
    strncpy(name, n.c_str(), cap - 1);
    name[cap - 1] = 0;
    if (ms) *ms = g_agg[n].ms;
    if (launches) *launches = g_agg[n].launches;
    return SSG_OK;
}
