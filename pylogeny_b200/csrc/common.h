// Shared host-side plumbing for the C-ABI library: error reporting, CUDA checks.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <mutex>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "../../include/pylogeny_b200.h"

namespace plg {

void set_error(const char *fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::mutex g_api_mutex;

// Thrown inside the library, always caught at the extern "C" boundary.
struct Error {
  int code;
};

#define PLG_CUDA(call)                                                               \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess) {                                                        \
      plg::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, \
                     __LINE__);                                                      \
      throw plg::Error{PLG_ERR_CUDA};                                                \
    }                                                                                \
  } while (0)

#define PLG_FAIL(code, ...)        \
  do {                             \
    plg::set_error(__VA_ARGS__);   \
    throw plg::Error{code};        \
  } while (0)

#define PLG_API_BEGIN try {
// compute entry points: the scratch pools are shared per device, so calls are serialised
#define PLG_API_BEGIN_LOCKED try { std::lock_guard<std::mutex> plg_lock__(plg::g_api_mutex);
#define PLG_API_END                                    \
  }                                                    \
  catch (const plg::Error &e) { return e.code; }       \
  catch (const std::bad_alloc &) {                     \
    plg::set_error("out of host memory");              \
    return PLG_ERR_ARG;                                \
  }                                                    \
  catch (...) {                                        \
    plg::set_error("unexpected C++ exception");        \
    return PLG_ERR_ARG;                                \
  }                                                    \
  return PLG_OK;

// Where the per-tree results of the running entry point go: host memory (default) or device memory of
// the current device (the *_out / *_shard entry points with PLG_OUT_DEVICE: scores land in a caller's
// tensor and go straight into a collective).  Set by an OutDeviceScope in the C-ABI wrapper, read where
// the result vector is copied out.
bool out_on_device();
struct OutDeviceScope {
  bool prev;
  explicit OutDeviceScope(bool on);
  ~OutDeviceScope();
};
inline cudaMemcpyKind out_copy_kind() { return out_on_device() ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost; }

inline void count_launch(int64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// Optional per-kernel timing with CUDA events on the launching stream (bench.py's roofline leg).
enum ProfKind { PROF_FITCH_OPS = 0, PROF_CLV_UPDATE = 1, PROF_ROOT_EVAL = 2, PROF_PMATRIX = 3, PROF_WALK = 4, PROF_KINDS = 5 };
bool prof_enabled();
void prof_begin(int kind, cudaStream_t st);
// algorithmic_bytes: SURVEY.md 8d per-node-op figures (independent of fusion / operand sharing);
// requested_bytes: what this implementation actually loads and stores (< 0: same as algorithmic)
void prof_end(int kind, cudaStream_t st, double algorithmic_bytes, double units, double requested_bytes = -1.0);
void prof_patch_last(int kind, double algorithmic_bytes, double units, double requested_bytes);

// PLG_TIMING=1: host-side stage times on stderr.
struct StageTimer {
  bool on;
  double t0;
  static double now();
  StageTimer();
  void lap(const char *what);
};

// Owning device buffer.  Small buffers (< 64 MB: alignment handles, op programs) come from the
// device's stream-ordered memory pool, which keeps freed blocks, so creating and destroying a handle
// does not pay cudaMalloc / cudaFree each time; the multi-GB scratch pools use cudaMalloc so that
// what they give back is visible to cudaMemGetInfo.  Every use of a buffer is complete when it is
// released: compute entry points synchronise their stream before returning.
void pool_setup_once();
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool pooled = false;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) { if (pooled) cudaFreeAsync(p, 0); else cudaFree(p); }
    p = nullptr;
    n = 0;
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    release();
    if (count == 0) return;
    pooled = count * sizeof(T) < ((size_t)64 << 20);
    if (pooled) {
      pool_setup_once();
      PLG_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), 0));
    } else {
      PLG_CUDA(cudaMalloc((void **)&p, count * sizeof(T)));
    }
    n = count;
  }
};

// Owning pinned host buffer (async H2D/D2H staging).
template <typename T>
struct PinBuf {
  T *p = nullptr;
  size_t n = 0;
  PinBuf() = default;
  PinBuf(const PinBuf &) = delete;
  PinBuf &operator=(const PinBuf &) = delete;
  ~PinBuf() {
    if (p) cudaFreeHost(p);
  }
  void alloc(size_t count) {
    if (count <= n && p) return;
    if (p) cudaFreeHost(p);
    p = nullptr;
    n = 0;
    if (count == 0) return;
    PLG_CUDA(cudaMallocHost((void **)&p, count * sizeof(T)));
    n = count;
  }
};

// Growable host array in pinned memory (async H2D of op programs at full PCIe rate); plain
// malloc when no device is present.  Contents are NOT preserved across resize().
template <typename T>
struct HostArr {
  T *p = nullptr;
  size_t n = 0, cap = 0;
  bool pinned = false;
  HostArr() = default;
  HostArr(const HostArr &) = delete;
  HostArr &operator=(const HostArr &) = delete;
  ~HostArr() { release(); }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else free(p); }
    p = nullptr; n = cap = 0;
  }
  void resize(size_t count) {
    if (count > cap) {
      release();
      size_t want = count + count / 4 + 16;
      if (cudaMallocHost((void **)&p, want * sizeof(T)) == cudaSuccess) pinned = true;
      else { cudaGetLastError(); p = (T *)malloc(want * sizeof(T)); pinned = false; if (!p) throw std::bad_alloc(); }
      cap = want;
    }
    n = count;
  }
  void clear() { n = 0; }
  size_t size() const { return n; }
  bool empty() const { return n == 0; }
  T *data() { return p; }
  const T *data() const { return p; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

}  // namespace plg
