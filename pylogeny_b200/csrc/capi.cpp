// Library-wide C-ABI pieces: error string, version, device selection, launch counter.
#include <cstring>

#include "common.h"

namespace plg {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};
std::mutex g_api_mutex;

static thread_local bool g_out_dev = false;
bool out_on_device() { return g_out_dev; }
OutDeviceScope::OutDeviceScope(bool on) : prev(g_out_dev) { g_out_dev = on; }
OutDeviceScope::~OutDeviceScope() { g_out_dev = prev; }

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}
}  // namespace plg

namespace plg {
// keep freed blocks in the pool of every device this process touches
void pool_setup_once() {
  static bool done[64] = {false};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || done[dev & 63]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    unsigned long long keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  done[dev & 63] = true;
}
}  // namespace plg

extern "C" const char *plg_last_error(void) { return plg::g_err; }
extern "C" int plg_version(void) { return 100; }
extern "C" int64_t plg_launch_count(void) { return plg::g_launches.load(); }

extern "C" int plg_device_count(int *count) {
  PLG_API_BEGIN
  if (!count) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
  *count = n;
  PLG_API_END
}

extern "C" int plg_set_device(int device) {
  PLG_API_BEGIN
  PLG_CUDA(cudaSetDevice(device));
  PLG_API_END
}

// Host worker pool check: n slices, slice `fail_at` (if in range) fails the way planner code fails.
// The failure must come back as a status (not std::terminate), after every other slice ran, and the
// pool must serve the next call.
#include "hostpool.h"
extern "C" int plg_hostpool_selftest(int n, int fail_at, int *ran) {
  PLG_API_BEGIN
  if (!ran || n < 1) PLG_FAIL(PLG_ERR_ARG, "bad argument");
  std::atomic<int> count{0};
  *ran = 0;
  try {
    plg::HostPool::get().run(n, [&](int i, int) {
      if (i == fail_at) PLG_FAIL(PLG_ERR_ARG, "slice %d failed on request", i);
      count++;
    });
  } catch (...) {
    *ran = count.load();
    throw;
  }
  *ran = count.load();
  PLG_API_END
}

// ---- host stage timer ------------------------------------------------------------------------
#include <chrono>
namespace plg {
double StageTimer::now() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
StageTimer::StageTimer() : on(getenv("PLG_TIMING") != nullptr), t0(now()) {}
void StageTimer::lap(const char *what) {
  if (!on) return;
  const double t1 = now();
  fprintf(stderr, "[plg] %-14s %8.3f ms\n", what, t1 - t0);
  t0 = t1;
}
}  // namespace plg

// ---- per-kernel profiling -------------------------------------------------------------------
namespace plg {
namespace {
struct ProfRec { int kind; cudaEvent_t a, b; double bytes, units, requested; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof_recs;
std::vector<cudaEvent_t> g_prof_pool;
cudaEvent_t g_prof_open[PROF_KINDS];
cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) { e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEventCreate(&e);
  return e;
}
}  // namespace
bool prof_enabled() { return g_prof_on; }
void prof_begin(int kind, cudaStream_t st) {
  if (!g_prof_on) return;
  g_prof_open[kind] = prof_event();
  cudaEventRecord(g_prof_open[kind], st);
}
void prof_end(int kind, cudaStream_t st, double bytes, double units, double requested) {
  if (!g_prof_on) return;
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, st);
  g_prof_recs.push_back({kind, g_prof_open[kind], b, bytes, units, requested < 0 ? bytes : requested});
}
// fills in the byte accounting of the most recent record of a kind (device-planned walks learn it
// from the GPU only when the call's results come back)
void prof_patch_last(int kind, double bytes, double units, double requested) {
  for (size_t i = g_prof_recs.size(); i-- > 0;)
    if (g_prof_recs[i].kind == kind) {
      g_prof_recs[i].bytes = bytes;
      g_prof_recs[i].units = units;
      g_prof_recs[i].requested = requested < 0 ? bytes : requested;
      return;
    }
}
}  // namespace plg

extern "C" int plg_profile_enable(int on) {
  plg::g_prof_on = on != 0;
  return PLG_OK;
}

// Sums (and clears) what was recorded for one kernel kind: device milliseconds between the
// events around each launch, algorithmic bytes, node-op x pattern units, launches.
extern "C" int plg_profile_read(int kind, double *ms, double *bytes, double *units, int64_t *launches,
                                double *requested_bytes) {
  PLG_API_BEGIN_LOCKED
  if (kind < 0 || kind >= plg::PROF_KINDS) PLG_FAIL(PLG_ERR_ARG, "bad kernel kind");
  PLG_CUDA(cudaDeviceSynchronize());
  double t = 0, by = 0, un = 0, rq = 0;
  int64_t n = 0;
  std::vector<plg::ProfRec> keep;
  for (auto &r : plg::g_prof_recs) {
    if (r.kind != kind) { keep.push_back(r); continue; }
    float e = 0;
    PLG_CUDA(cudaEventElapsedTime(&e, r.a, r.b));
    t += e; by += r.bytes; un += r.units; rq += r.requested; n++;
    plg::g_prof_pool.push_back(r.a);
    plg::g_prof_pool.push_back(r.b);
  }
  plg::g_prof_recs.swap(keep);
  if (ms) *ms = t;
  if (bytes) *bytes = by;
  if (units) *units = un;
  if (launches) *launches = n;
  if (requested_bytes) *requested_bytes = rq;
  PLG_API_END
}
