// core.cu -- error plumbing, launch counter, device queries.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace dpm {

static thread_local char g_err[512] = "";
static thread_local long long g_launches = 0;

char *err_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// ---- per-launch profile (bench.py's kernel breakdown) ----------------------------------
// dpm_prof_begin(stream) records a start event; while open, every launch records one more
// event after itself, so event[i-1] -> event[i] is the device time of launch i (all launches of
// a call are serialised on one stream).  dpm_prof_end() synchronises and reports.
struct ProfRec {
    const char *tag;
    long long a, b;
    cudaEvent_t ev;
};
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_start = nullptr;
static thread_local ProfRec *g_prof = nullptr;
static thread_local int g_prof_n = 0, g_prof_cap = 0, g_prof_made = 0;
static thread_local long long g_note_a = 0, g_note_b = 0;

void prof_note(long long a, long long b) {
    g_note_a = a;
    g_note_b = b;
}

static void prof_record(const char *tag, cudaStream_t st);

void count_launch(const char *tag, cudaStream_t st) {
    ++g_launches;
    if (g_prof_on) prof_record(tag, st);
}

bool prof_active() { return g_prof_on; }

// start of a C-ABI call: the stream time since the previous record is host-side gap, not kernel time
void prof_mark(cudaStream_t st) {
    if (g_prof_on) prof_record("host_gap", st);
}

static void prof_record(const char *tag, cudaStream_t st) {
    if (g_prof_n == g_prof_cap) {
        const int ncap = g_prof_cap ? g_prof_cap * 2 : 1024;
        ProfRec *np = (ProfRec *)realloc(g_prof, sizeof(ProfRec) * ncap);
        if (!np) return;
        g_prof = np;
        g_prof_cap = ncap;
    }
    ProfRec &r = g_prof[g_prof_n];
    if (g_prof_n >= g_prof_made) {
        if (cudaEventCreate(&r.ev) != cudaSuccess) return;
        ++g_prof_made;
    }
    r.tag = tag;
    r.a = g_note_a;
    r.b = g_note_b;
    g_note_a = g_note_b = 0;
    cudaEventRecord(r.ev, st);
    ++g_prof_n;
}

int current_device() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev;
}

int device_sm_count() {
    static thread_local int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;  // B200
        cached = n;
    }
    return cached;
}

}  // namespace dpm

extern "C" int dpm_prof_begin(dpm_stream_t stream) {
    using namespace dpm;
    if (!g_prof_start && cudaEventCreate(&g_prof_start) != cudaSuccess) return fail(DPM_ERR_CUDA, "prof: cudaEventCreate failed");
    g_prof_n = 0;
    if (cudaEventRecord(g_prof_start, (cudaStream_t)stream) != cudaSuccess) return fail(DPM_ERR_CUDA, "prof: cudaEventRecord failed");
    g_prof_on = true;
    return DPM_OK;
}

extern "C" int dpm_prof_end(char *buf, size_t buf_bytes) {
    using namespace dpm;
    g_prof_on = false;
    size_t off = 0;
    if (buf && buf_bytes) buf[0] = 0;
    cudaEvent_t prev = g_prof_start;
    for (int i = 0; i < g_prof_n; ++i) {
        if (cudaEventSynchronize(g_prof[i].ev) != cudaSuccess) return fail(DPM_ERR_CUDA, "prof: cudaEventSynchronize failed");
        float ms = 0.f;
        cudaEventElapsedTime(&ms, prev, g_prof[i].ev);
        prev = g_prof[i].ev;
        if (buf && off + 96 < buf_bytes)
            off += (size_t)snprintf(buf + off, buf_bytes - off, "%s %lld %lld %.6f\n", g_prof[i].tag, g_prof[i].a, g_prof[i].b, ms);
    }
    return g_prof_n;
}

extern "C" int dpm_version(void) { return 100; }
extern "C" const char *dpm_last_error(void) { return dpm::err_buf(); }
extern "C" long long dpm_launch_count(void) { return dpm::g_launches; }
extern "C" void dpm_launch_count_reset(void) { dpm::g_launches = 0; }
