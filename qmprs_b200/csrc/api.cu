// Library identification, launch counter and the optional per-kernel-class event profiler
// used by bench.py to time the dominant kernel live (CUDA events on the launching stream).
#include <cstdlib>
#include <vector>
#include "common.cuh"
#include "qmprs_b200.h"

namespace {
struct Pair { cudaEvent_t a, b; int cls; };
std::vector<Pair> g_pairs;
std::vector<cudaEvent_t> g_pool;
size_t g_pool_next = 0;
bool g_enabled = false;
long long g_launches = 0;
long long g_cls_launches[QM_NCLS] = {0};
double g_cls_work[QM_NCLS] = {0};

cudaEvent_t get_event() {
    if (g_pool_next == g_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        g_pool.push_back(e);
    }
    return g_pool[g_pool_next++];
}
}  // namespace

void qm_prof_pre(int cls, cudaStream_t st) {
    g_launches++;
    g_cls_launches[cls]++;
    if (!g_enabled) return;
    Pair p;
    p.a = get_event();
    p.b = get_event();
    p.cls = cls;
    cudaEventRecord(p.a, st);
    g_pairs.push_back(p);
}

void qm_prof_post(int cls, cudaStream_t st) {
    (void)cls;
    if (!g_enabled) return;
    cudaEventRecord(g_pairs.back().b, st);
}

void qm_prof_work(int cls, double work) { g_cls_work[cls] += work; }

bool qm_prof_active() { return g_enabled; }

// QM_PDL=0 turns programmatic dependent launch off (A/B runs); never used under the event profiler
// (events between launches serialise them anyway and the classes are timed in isolation there).
static int g_pdl = -1;
bool qm_pdl_enabled() {
    if (g_pdl < 0) g_pdl = !(getenv("QM_PDL") && atoi(getenv("QM_PDL")) == 0);
    return g_pdl && !g_enabled;
}
// run-time switch (tests run the same problem with and without programmatic dependent launch); returns the old value
extern "C" int qm_set_pdl(int on) {
    int old = g_pdl < 0 ? !(getenv("QM_PDL") && atoi(getenv("QM_PDL")) == 0) : g_pdl;
    g_pdl = on ? 1 : 0;
    return old;
}

extern "C" int qm_version(void) { return 100; }

extern "C" long long qm_launch_count(void) { return g_launches; }

extern "C" int qm_prof_num_classes(void) { return QM_NCLS; }

extern "C" const char* qm_prof_class_name(int cls) {
    static const char* names[QM_NCLS] = {"zgemm", "svd_gram", "svd_eig", "svd_apply", "svd_layout", "qr_vec",
                                         "qr_apply", "small", "gate", "env_polar", "svd_round"};
    return (cls >= 0 && cls < QM_NCLS) ? names[cls] : "?";
}

extern "C" int qm_prof_begin(void) {
    g_pairs.clear();
    g_pool_next = 0;
    for (int i = 0; i < QM_NCLS; i++) { g_cls_launches[i] = 0; g_cls_work[i] = 0.0; }
    g_enabled = true;
    return 0;
}

// total_ms[QM_NCLS], count[QM_NCLS] (host arrays).  Synchronises the device.
extern "C" int qm_prof_work_get(double* work) {
    for (int i = 0; i < QM_NCLS; i++) work[i] = g_cls_work[i];
    return 0;
}

extern "C" int qm_prof_end(double* total_ms, long long* count) {
    g_enabled = false;
    QM_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < QM_NCLS; i++) { total_ms[i] = 0.0; count[i] = g_cls_launches[i]; }
    for (const Pair& p : g_pairs) {
        float ms = 0.f;
        QM_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
        total_ms[p.cls] += (double)ms;
    }
    g_pairs.clear();
    g_pool_next = 0;
    return 0;
}
