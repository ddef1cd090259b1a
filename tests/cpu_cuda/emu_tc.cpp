// Host side and asynchronous-unit model of the tcgen05 / TMA emulation (stub_tc/tc_common.cuh).  TEST INFRASTRUCTURE.
//
// Two completion models for the asynchronous units, selected by SSG_EMU_ASYNC:
//   eager (default): a TMA load, a TMA store and an MMA take effect when they are issued -- the EARLIEST legal moment;
//   late           : they take effect at the LATEST legal moment -- a TMA load lands (and completes its transaction
//                    bytes) only when somebody polls the mbarrier it signals; an MMA executes, and its tcgen05.commit
//                    arrives, only when the committed mbarrier is polled; a TMA store reads shared memory only when
//                    cp.async.bulk.wait_group(.read) stops allowing it to be pending.
//   mixed           : every operation draws (seeded) whether it behaves eagerly or late -- see defer_now().
// A kernel whose protocol is right computes the same bytes under all of them; one that refills an operand stage before the MMA
// consumed it, rewrites a staging buffer under an in-flight store, or reads a tile without waiting for its barrier does
// not.
#include <stdlib.h>
#include <string.h>
#include <deque>
#include <map>
#include <vector>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

namespace emu {
struct MBar { int count = 0, pending = 0; long long tx = 0; unsigned phase = 0; };
std::map<const void*, MBar>& mbars();
void note_event();
extern float tmem[128][512];

static int g_late = -1;                                        // -1: take it from SSG_EMU_ASYNC on first use
static unsigned g_arng = 12345u;
static int late_mode() {
    if (g_late < 0) {
        const char* e = getenv("SSG_EMU_ASYNC");
        g_late = (e && !strcmp(e, "late")) ? 1 : (e && !strncmp(e, "mixed", 5)) ? 2 : 0;
        if (g_late == 2 && e[5] == ':') g_arng = (unsigned)atoi(e + 6) * 2654435761u + 1u;
    }
    return g_late;
}
// mixed (SSG_EMU_ASYNC=mixed[:seed], mode 2): every operation draws whether it completes at issue or as late as allowed
// (an MMA or commit behind a deferred one is deferred too: the tensor pipe is in order) -- interleavings between the
// two extremes, e.g. the A tile of a stage landing at once and its B tile at the last moment.
static bool defer_now() {
    if (late_mode() != 2) return late_mode() == 1;
    g_arng = g_arng * 1664525u + 1013904223u;
    return (g_arng >> 16) & 1u;
}

static inline uint32_t swz(uint32_t off, uint32_t mode) {
    const uint32_t bits = mode == CU_TENSOR_MAP_SWIZZLE_128B ? 3 : mode == CU_TENSOR_MAP_SWIZZLE_64B ? 2 : mode == CU_TENSOR_MAP_SWIZZLE_32B ? 1 : 0;
    return off ^ (((off >> 7) & ((1u << bits) - 1u)) << 4);
}
static void flip_if_complete(MBar& b) {
    note_event();
    if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.count; }
}
static void box_copy(const CUtensorMap* m, uint32_t dst0, const int* c, bool store) {
    uint32_t nb[5] = {1, 1, 1, 1, 1};
    for (uint32_t i = 0; i < m->rank; ++i) nb[i] = (m->box[i] + m->estr[i] - 1) / m->estr[i];
    uint32_t lin = 0;
    for (uint32_t b4 = 0; b4 < nb[4]; ++b4)
    for (uint32_t b3 = 0; b3 < nb[3]; ++b3)
    for (uint32_t b2 = 0; b2 < nb[2]; ++b2)
    for (uint32_t b1 = 0; b1 < nb[1]; ++b1)
    for (uint32_t b0 = 0; b0 < nb[0]; ++b0, ++lin) {
        const uint32_t bi[5] = {b0, b1, b2, b3, b4};
        bool inb = true;
        uint64_t goff = 0;
        for (uint32_t i = 0; i < m->rank; ++i) {
            const long long g = (long long)c[i] + (long long)bi[i] * m->estr[i];
            if (g < 0 || g >= (long long)m->dims[i]) { inb = false; break; }
            goff += i == 0 ? (uint64_t)g * 2 : (uint64_t)g * m->strides[i - 1];
        }
        unsigned char* s = dyn_smem + swz(dst0 + lin * 2, m->swizzle);
        uint16_t* g16 = (uint16_t*)(uintptr_t)(m->base + goff);
        if (store) { if (inb) *g16 = *(uint16_t*)s; }
        else *(uint16_t*)s = inb ? *g16 : (uint16_t)0;
    }
}
static uint32_t box_bytes(const CUtensorMap* m) {
    uint32_t n = 2;
    for (uint32_t i = 0; i < m->rank; ++i) n *= (m->box[i] + m->estr[i] - 1) / m->estr[i];
    return n;
}

// ---- deferred work (late mode)
struct Load { CUtensorMap map; uint32_t dst; int c[5]; const void* bar; };
struct Store { CUtensorMap map; uint32_t src; int c[5]; };
struct Mma { bool is_commit; const void* bar; uint32_t tmem_d, idesc, accumulate; uint64_t da, db; };
static std::deque<Load> g_loads;
static std::vector<Store> g_open_group;
static std::deque<std::vector<Store>> g_groups;
static std::deque<Mma> g_mmas;

static float desc_elem(uint64_t desc, int r, int e) {
    const uint32_t start = (uint32_t)(desc & 0x3fffu) << 4;
    const uint32_t sbo = (uint32_t)((desc >> 32) & 0x3fffu) << 4;
    const uint32_t layout = (uint32_t)(desc >> 61) & 7u;                     // 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
    const uint32_t row_bytes = layout == 2 ? 128 : layout == 4 ? 64 : 32;
    const uint32_t mode = layout == 2 ? CU_TENSOR_MAP_SWIZZLE_128B : layout == 4 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
    const uint32_t off = start + (uint32_t)(r / 8) * sbo + (uint32_t)(r % 8) * row_bytes + (uint32_t)e * 2;
    __nv_bfloat16 h;
    h.x = *(const uint16_t*)(dyn_smem + swz(off, mode));
    return __bfloat162float(h);
}
static void mma_now(const Mma& q) {
    const int N = (int)((q.idesc >> 17) & 0x3fu) << 3, M = (int)((q.idesc >> 24) & 0x1fu) << 4;
    const int col0 = (int)(q.tmem_d & 0xffffu);
    static float a[128][16], b[256][16];
    for (int r = 0; r < M; ++r) for (int e = 0; e < 16; ++e) a[r][e] = desc_elem(q.da, r, e);
    for (int n = 0; n < N; ++n) for (int e = 0; e < 16; ++e) b[n][e] = desc_elem(q.db, n, e);
    for (int r = 0; r < M; ++r)
        for (int n = 0; n < N; ++n) {
            float acc = q.accumulate ? tmem[r][col0 + n] : 0.f;
            for (int e = 0; e < 16; ++e) acc += a[r][e] * b[n][e];
            tmem[r][col0 + n] = acc;
        }
}
static void arrive(const void* bar) { MBar& b = mbars()[bar]; --b.pending; flip_if_complete(b); }

void tc_reset() { g_loads.clear(); g_open_group.clear(); g_groups.clear(); g_mmas.clear(); }

void tc_mbar_init(const void* bar, uint32_t count) { MBar& b = mbars()[bar]; b.count = b.pending = (int)count; b.tx = 0; b.phase = 0; }
void tc_mbar_arrive(const void* bar) { arrive(bar); }
void tc_mbar_arrive_expect_tx(const void* bar, uint32_t bytes) { MBar& b = mbars()[bar]; b.tx += bytes; --b.pending; flip_if_complete(b); }
bool tc_mbar_poll(const void* bar, uint32_t parity) {
    if (late_mode()) {
        // everything that signals THIS barrier and is still in flight lands now
        for (auto it = g_loads.begin(); it != g_loads.end();) {
            if (it->bar == bar) {
                box_copy(&it->map, it->dst, it->c, false);
                MBar& b = mbars()[bar]; b.tx -= box_bytes(&it->map); flip_if_complete(b);
                it = g_loads.erase(it);
            } else ++it;
        }
        int last = -1;
        for (int i = 0; i < (int)g_mmas.size(); ++i) if (g_mmas[i].is_commit && g_mmas[i].bar == bar) last = i;
        for (int i = 0; i <= last; ++i) {                      // in order, up to the last commit onto this barrier
            const Mma q = g_mmas.front(); g_mmas.pop_front();
            if (q.is_commit) arrive(q.bar); else mma_now(q);
        }
    }
    const bool done = mbars()[bar].phase != parity;
    if (done) note_event();                 // a thread leaving a wait is progress (emu.cpp's deadlock detection)
    return done;
}
void tc_tma_load(void* smem_dst, const CUtensorMap* m, const void* bar, const int* c) {
    const uint32_t dst = (uint32_t)((unsigned char*)smem_dst - dyn_smem);
    if (defer_now()) { Load l; l.map = *m; l.dst = dst; memcpy(l.c, c, sizeof(l.c)); l.bar = bar; g_loads.push_back(l); note_event(); return; }
    box_copy(m, dst, c, false);
    MBar& b = mbars()[bar]; b.tx -= box_bytes(m); flip_if_complete(b);
}
void tc_tma_store(const CUtensorMap* m, const void* smem_src, const int* c) {
    const uint32_t src = (uint32_t)((const unsigned char*)smem_src - dyn_smem);
    if (defer_now()) { Store s; s.map = *m; s.src = src; memcpy(s.c, c, sizeof(s.c)); g_open_group.push_back(s); return; }
    box_copy(m, src, c, true);
}
void tc_store_commit() { if (late_mode()) { g_groups.push_back(g_open_group); g_open_group.clear(); } }
void tc_store_wait(int allowed_pending) {
    while ((int)g_groups.size() > allowed_pending) {
        for (const Store& s : g_groups.front()) box_copy(&s.map, s.src, s.c, true);
        g_groups.pop_front();
        note_event();
    }
}
void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    Mma q; q.is_commit = false; q.bar = nullptr; q.tmem_d = tmem_d; q.idesc = idesc; q.accumulate = accumulate; q.da = da; q.db = db;
    if (!g_mmas.empty() || defer_now()) { g_mmas.push_back(q); note_event(); } else mma_now(q);
}
void tc_mma_commit(const void* bar) {
    if (!g_mmas.empty() || defer_now()) { Mma q; memset(&q, 0, sizeof(q)); q.is_commit = true; q.bar = bar; g_mmas.push_back(q); note_event(); }
    else arrive(bar);
}
void set_tc_reset(void (*f)());
static int g_registered = (set_tc_reset(tc_reset), 0);       // emu.cpp clears the per-block state through this hook
}  // namespace emu

// switch the completion model between launches (tests run both models in one process)
extern "C" void ssg_emu_set_async(int mode) { emu::g_late = mode < 0 || mode > 2 ? 0 : mode; }    // 0 eager, 1 late, 2 mixed
extern "C" void ssg_emu_seed_async(unsigned seed) { emu::g_arng = seed * 2654435761u + 1u; }

static CUresult emu_encode_tiled(CUtensorMap* m, CUtensorMapDataType, cuuint32_t rank, void* base, const cuuint64_t* dims,
                                 const cuuint64_t* strides, const cuuint32_t* box, const cuuint32_t* estr,
                                 CUtensorMapInterleave, CUtensorMapSwizzle swz, CUtensorMapL2promotion, CUtensorMapFloatOOBfill) {
    memset(m, 0, sizeof(*m));
    m->base = (uint64_t)(uintptr_t)base;
    m->rank = rank;
    m->swizzle = (uint32_t)swz;
    for (cuuint32_t i = 0; i < rank; ++i) {
        m->dims[i] = (uint32_t)dims[i];
        m->box[i] = box[i];
        m->estr[i] = estr[i];
        if (i + 1 < rank) m->strides[i] = strides[i];
    }
    return CUDA_SUCCESS;
}

cudaError_t cudaGetDriverEntryPoint(const char* name, void** fn, unsigned long long, cudaDriverEntryPointQueryResult* res) {
    *fn = strcmp(name, "cuTensorMapEncodeTiled") == 0 ? (void*)emu_encode_tiled : nullptr;
    if (res) *res = cudaDriverEntryPointSuccess;
    return cudaSuccess;
}
