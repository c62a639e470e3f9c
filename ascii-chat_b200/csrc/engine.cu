// engine.cu — host engine of libasciichat_b200: plans, per-thread streams and staging,
// device LUT cache, the batch API (include/asciichat_b200.h Part 2).
//
// There is no CPU rendering path in this file: every byte of every frame string is produced by
// the kernels in render_kernels.cu.  Host work is limited to validation, the float aspect fit
// (kept on the host on purpose), building the 256-entry glyph tables from the
// palette string, moving buffers, and copying finished strings into caller-owned memory.
#include "engine.h"

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <immintrin.h>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include <sched.h>

// forwarded to when the host binary provides it (include/ascii-chat/asciichat_errno.h:429)
extern "C" void asciichat_set_errno_with_message(int code, const char *file, int line, const char *function,
                                                 const char *format, ...) __attribute__((weak));

namespace acb {

// ------------------------------------------------------------------ errors / knobs
static thread_local int t_err = 0;
static thread_local char t_errmsg[256] = "";
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_opt_render_mode{0};
static std::atomic<int> g_default_scale{ACB200_SCALE_NN};
static void *(*g_alloc)(size_t) = malloc;
static void (*g_free)(void *) = free;

int set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_errmsg, sizeof(t_errmsg), fmt, ap);
  va_end(ap);
  t_err = code;
  if (asciichat_set_errno_with_message) asciichat_set_errno_with_message(code, "asciichat_b200", 0, "", "%s", t_errmsg);
  return code;
}
void *user_alloc(size_t n) { return g_alloc(n ? n : 1); }
void user_free(void *p) { g_free(p); }
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
int option_render_mode() { return g_opt_render_mode.load(); }
int default_scale() { return g_default_scale.load(); }

// ------------------------------------------------------------------ devices / thread contexts
// One process may drive several GPUs (acb200_init_devices): the reference server is ONE process with a render thread
// per client (src/server/render.c:340-652), so the device pool lives behind the C ABI — calling threads are leased a
// context on one device of the pool, round-robin, the first time they call in, and keep it for their lifetime.
struct DeviceState {
  int sms = 148;
  bool in_pool = false;
  std::mutex mu;                              // pool + LUT cache of this device
  std::vector<ThreadCtx *> pool;              // warm contexts handed back by exited threads
  std::map<std::string, GlyphLut *> luts;     // key = which + palette bytes
  std::vector<GlyphLut *> retired;            // evicted, possibly still referenced by a launch being prepared
  std::atomic<int> fetch_inflight{0};         // drop-in calls whose frame the GPU is fetching from pinned host memory
};
static DeviceState g_ds[kMaxDevices]; // indexed by CUDA ordinal
static std::mutex g_dev_mu;
static int g_dev_status = -1;         // -1 not initialised, 0 ready, else the error code
static int g_pool_devs[kMaxDevices];  // ordinals in use
static int g_npool = 0;
static int g_requested_device = -1;   // acb200_init(device) before the first use
static bool g_peer[kMaxDevices][kMaxDevices];
static std::atomic<unsigned> g_rr{0};
// 0 spin (cudaStreamSynchronize), 1 block, 2 hybrid, 3 yield (poll a completion word the stream writes into mapped
// memory; sched_yield between polls).  Measured on the 16-core B200 box (profiles/r02a_e2e_sweep): spinning callers
// deliver 35.5 k frames/s at 16 threads, sleeping ones 20 k (28 k at 64 threads) — the wake-up latency of a blocking
// event costs more than the core the spin burns.  But a spinning caller also keeps its core from the callers that have
// a gather or a copy-out to do, which is what stops the call from scaling once several GPUs share the host's cores
// (DESIGN §9): mode 3 spins only while nobody else wants the core.
static std::atomic<int> g_sync_mode{0}, g_spin_us{30};
// Frames in page-locked host memory (acb200_register_host_memory, or cudaHostAlloc'd by the caller) need no host-side
// gather: the GPU fetches the source rows nearest-neighbour sampling reads itself (2.2 MB of a 4K frame) and samples
// the columns.  That frees the caller's core — the gather is 100-250 us of a ~330 us call, profiles/r02l_e2e_inproc2.txt
// — but the link carries 12 x the bytes of the host-gathered plan (0.18 MB) and tops out at 13-15 k frames/s per GPU,
// below what 16 gathering cores deliver (34 k, bounded by the strings going the other way).  So it pays where cores
// are scarcer than GPUs, and the depth rations it: at most this many calls per GPU fetch at a time, the rest gather
// on their cores; -1 = always, 0 = never (default; acb200_set_fetch_depth / ACB200_FETCH_DEPTH).
static std::atomic<int> g_fetch_depth{0};
// cuStreamWriteValue32 through the runtime's driver-entry-point lookup (the library links no libcuda symbol directly)
typedef int (*StreamWriteValue32)(cudaStream_t, unsigned long long, uint32_t, unsigned);
static StreamWriteValue32 g_write_value32 = nullptr;
// where a drop-in call's host time goes (ns, summed over calls): gather/staging, enqueue, wait, copy-out; [4] = calls
static std::atomic<uint64_t> g_phase_ns[5];
static inline uint64_t now_ns() {
  timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (uint64_t)t.tv_sec * 1000000000ull + (uint64_t)t.tv_nsec;
}

static int init_pool_locked(const int *devs, int n) { // g_dev_mu held
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0)
    return set_error(E_INVALID_STATE, "asciichat_b200: no CUDA device (%s); this library has no CPU path",
                     e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
  int cur = 0;
  cudaGetDevice(&cur);
  int list[kMaxDevices], m = 0;
  if (!devs || n <= 0) {
    list[m++] = cur;
  } else {
    for (int i = 0; i < n && m < kMaxDevices; i++) {
      if (devs[i] < 0 || devs[i] >= count || devs[i] >= kMaxDevices)
        return set_error(E_INVALID_PARAM, "asciichat_b200: device %d out of range (%d visible)", devs[i], count);
      bool dup = false;
      for (int j = 0; j < m; j++) dup |= list[j] == devs[i];
      if (!dup) list[m++] = devs[i];
    }
  }
  for (int i = 0; i < m; i++) {
    cudaDeviceProp pr;
    if (cudaGetDeviceProperties(&pr, list[i]) != cudaSuccess)
      return set_error(E_INVALID_STATE, "asciichat_b200: cannot query device %d", list[i]);
    if (pr.major < 10)
      return set_error(E_INVALID_STATE, "asciichat_b200: device %s is sm_%d%d; this build is sm_100a only", pr.name,
                       pr.major, pr.minor);
    g_ds[list[i]].sms = pr.multiProcessorCount > 0 ? pr.multiProcessorCount : 148;
  }
  // peer access between every pair of the pool: a viewer's kernels read the sources of clients resident on other GPUs
  // straight over NVLink (server.cu), and the grid's cell renders store into the composing GPU's arena
  for (int i = 0; i < m; i++)
    for (int j = 0; j < m; j++) {
      const int a = list[i], b = list[j];
      if (a == b) {
        g_peer[a][b] = true;
        continue;
      }
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, a, b) == cudaSuccess && can) {
        cudaSetDevice(a);
        cudaError_t pe = cudaDeviceEnablePeerAccess(b, 0);
        if (pe == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
        g_peer[a][b] = pe == cudaSuccess || pe == cudaErrorPeerAccessAlreadyEnabled;
      }
    }
  cudaSetDevice(cur);
  for (int i = 0; i < m; i++) {
    g_pool_devs[i] = list[i];
    g_ds[list[i]].in_pool = true;
  }
  g_npool = m;
  if (const char *e = getenv("ACB200_SYNC")) { // measurement knob: spin | block | hybrid[:microseconds]
    if (!strncmp(e, "spin", 4)) g_sync_mode.store(0);
    else if (!strncmp(e, "block", 5)) g_sync_mode.store(1);
    else if (!strncmp(e, "hybrid", 6)) {
      g_sync_mode.store(2);
      if (e[6] == ':') g_spin_us.store(atoi(e + 7));
    } else if (!strncmp(e, "yield", 5)) g_sync_mode.store(3);
  }
  if (const char *e = getenv("ACB200_FETCH_DEPTH")) g_fetch_depth.store(atoi(e)); // measurement knob
  {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &fn, cudaEnableDefault, &qr) == cudaSuccess &&
        qr == cudaDriverEntryPointSuccess)
      g_write_value32 = (StreamWriteValue32)fn;
    else
      cudaGetLastError();
  }
  return 0;
}

int ensure_device() {
  if (g_dev_status != 0) {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (g_dev_status == -1) {
      if (g_requested_device >= 0) {
        const int d = g_requested_device;
        g_dev_status = init_pool_locked(&d, 1);
      } else {
        g_dev_status = init_pool_locked(nullptr, 0);
      }
    }
  }
  if (g_dev_status != 0 && t_err == 0) set_error(E_INVALID_STATE, "asciichat_b200: CUDA device unavailable");
  return g_dev_status;
}

int device_sms() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < kMaxDevices ? g_ds[dev].sms : 148;
}
bool peer_ok(int from_dev, int to_dev) {
  return from_dev >= 0 && to_dev >= 0 && from_dev < kMaxDevices && to_dev < kMaxDevices && g_peer[from_dev][to_dev];
}

bool grow_pinned(uint8_t **p, size_t *cap, size_t need) {
  if (need <= *cap) return true;
  size_t n = need + need / 4 + 4096;
  if (*p) cudaFreeHost(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaHostAlloc((void **)p, n, cudaHostAllocPortable | cudaHostAllocMapped) != cudaSuccess) {
    set_error(E_MEMORY, "cudaHostAlloc(%zu) failed", n);
    return false;
  }
  *cap = n;
  return true;
}
bool grow_device(uint8_t **p, size_t *cap, size_t need) {
  if (need <= *cap) return true;
  size_t n = need + need / 4 + 4096;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  if (cudaMalloc((void **)p, n) != cudaSuccess) {
    set_error(E_MEMORY, "cudaMalloc(%zu) failed", n);
    return false;
  }
  *cap = n;
  return true;
}

// Streams and pinned staging are expensive to create (cudaHostAlloc is a multi-millisecond call), while the
// reference's callers are threads that come and go with clients.  A thread leases a context on first use and
// hands it back to its device's pool when it exits; a new thread picks up a warm one.
struct CtxLease {
  ThreadCtx *c = nullptr;
  ~CtxLease() { release(); }
  void release() {
    if (!c) return;
    DeviceState &ds = g_ds[c->device];
    std::lock_guard<std::mutex> lk(ds.mu);
    ds.pool.push_back(c);
    c = nullptr;
  }
};
static thread_local CtxLease t_lease;
static thread_local int t_bind = -1; // pool index requested by acb200_bind_thread, -1 = round-robin

ThreadCtx *thread_ctx() {
  if (t_lease.c) {
    int cur = -1;
    if (g_npool > 1 && (cudaGetDevice(&cur) != cudaSuccess || cur != t_lease.c->device)) cudaSetDevice(t_lease.c->device);
    return t_lease.c;
  }
  if (ensure_device() != 0) return nullptr;
  const int k = t_bind >= 0 ? t_bind % g_npool : (int)(g_rr.fetch_add(1) % (unsigned)g_npool);
  const int dev = g_pool_devs[k];
  if (cudaSetDevice(dev) != cudaSuccess) {
    set_error(E_INVALID_STATE, "cudaSetDevice(%d) failed", dev);
    return nullptr;
  }
  DeviceState &ds = g_ds[dev];
  {
    std::lock_guard<std::mutex> lk(ds.mu);
    if (!ds.pool.empty()) {
      t_lease.c = ds.pool.back();
      ds.pool.pop_back();
      return t_lease.c;
    }
  }
  ThreadCtx *c = new ThreadCtx();
  c->device = dev;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    set_error(E_INVALID_STATE, "cannot create CUDA stream");
    return nullptr;
  }
  for (auto &e : c->ev) cudaEventCreateWithFlags(&e, cudaEventDefault);
  cudaEventCreateWithFlags(&c->done, cudaEventBlockingSync | cudaEventDisableTiming);
  {
    void *f = nullptr;
    if (cudaHostAlloc(&f, 64, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
      memset(f, 0, 64);
      c->h_flag = (volatile uint32_t *)f;
    } else {
      cudaGetLastError();
    }
  }
  t_lease.c = c;
  return c;
}

// Waiting for the frame.  cudaStreamSynchronize spins on a host core for the whole GPU + PCIe time of the call; the
// alternatives (sleep on a blocking-sync event, or poll briefly and then sleep) are selectable for hosts where the
// cores are scarcer than on the boxes measured — see g_sync_mode.
int wait_stream(ThreadCtx *cx) {
  const int mode = g_sync_mode.load(std::memory_order_relaxed);
  if (mode == 3 && cx->h_flag && g_write_value32) {
    const uint32_t seq = ++cx->flag_seq;
    // the write is ordered behind everything queued so far and preceded by a system-wide fence (no NO_MEMORY_BARRIER
    // flag): when the word shows `seq`, the strings the kernels stored into mapped memory are visible
    if (g_write_value32(cx->stream, (unsigned long long)(uintptr_t)cx->h_flag, seq, 0) == 0) {
      volatile uint32_t *f = cx->h_flag;
      for (unsigned rounds = 1;; rounds++) {
        for (int i = 0; i < 48; i++) {
          if (*f == seq) {
            std::atomic_thread_fence(std::memory_order_acquire);
            return E_OK;
          }
          _mm_pause();
        }
        sched_yield();
        if ((rounds & 2047u) == 0) { // a faulted stream never writes the word
          cudaError_t q = cudaStreamQuery(cx->stream);
          if (q == cudaSuccess) return E_OK;
          if (q != cudaErrorNotReady) return set_error(E_INVALID_STATE, "CUDA: %s (cudaStreamQuery)", cudaGetErrorString(q));
        }
      }
    }
  }
  if (mode == 0 || mode == 3 || !cx->done) {
    ACB_CUDA(cudaStreamSynchronize(cx->stream));
    return E_OK;
  }
  ACB_CUDA(cudaEventRecord(cx->done, cx->stream));
  if (mode == 2) {
    const int spin_us = g_spin_us.load(std::memory_order_relaxed);
    timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (;;) {
      cudaError_t q = cudaEventQuery(cx->done);
      if (q == cudaSuccess) return E_OK;
      if (q != cudaErrorNotReady) return set_error(E_INVALID_STATE, "CUDA: %s (cudaEventQuery)", cudaGetErrorString(q));
      clock_gettime(CLOCK_MONOTONIC, &t1);
      if ((t1.tv_sec - t0.tv_sec) * 1000000L + (t1.tv_nsec - t0.tv_nsec) / 1000L >= spin_us) break;
    }
  }
  ACB_CUDA(cudaEventSynchronize(cx->done));
  return E_OK;
}

// Entry points that take a caller-owned stream still keep their tables in this thread's scratch.  Before that scratch is
// touched on another stream, whatever the previous caller-owned stream queued on it must have drained.
int sync_foreign(ThreadCtx *cx, cudaStream_t st) {
  if (cx->foreign && cx->foreign != st) ACB_CUDA(cudaStreamSynchronize(cx->foreign));
  cx->foreign = st != cx->stream ? st : nullptr;
  return E_OK;
}

PeerHelper *peer_helper(ThreadCtx *cx, int device) {
  if (device < 0 || device >= kMaxDevices || !g_ds[device].in_pool) {
    set_error(E_INVALID_PARAM, "device %d is not in the pool", device);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    set_error(E_INVALID_STATE, "cudaSetDevice(%d) failed", device);
    return nullptr;
  }
  PeerHelper *h = cx->helper[device];
  if (h) return h;
  h = new PeerHelper();
  if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev, cudaEventDisableTiming) != cudaSuccess) {
    delete h;
    set_error(E_INVALID_STATE, "cannot create the helper stream on device %d", device);
    return nullptr;
  }
  cx->helper[device] = h;
  return h;
}

static void free_ctx(ThreadCtx *c) {
  for (int d = 0; d < kMaxDevices; d++)
    if (PeerHelper *h = c->helper[d]) {
      cudaSetDevice(d);
      if (h->scratch) cudaFree(h->scratch);
      if (h->out) cudaFree(h->out);
      if (h->ev) cudaEventDestroy(h->ev);
      if (h->stream) cudaStreamDestroy(h->stream);
      delete h;
    }
  cudaSetDevice(c->device);
  if (c->h_in) cudaFreeHost(c->h_in);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_len) cudaFreeHost(c->h_len);
  if (c->h_flag) cudaFreeHost((void *)c->h_flag);
  if (c->d_in) cudaFree(c->d_in);
  if (c->d_out) cudaFree(c->d_out);
  if (c->d_rows) cudaFree(c->d_rows);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->d_len) cudaFree(c->d_len);
  if (c->d_frame) cudaFree(c->d_frame);
  for (auto &e : c->ev)
    if (e) cudaEventDestroy(e);
  if (c->done) cudaEventDestroy(c->done);
  if (c->stream) cudaStreamDestroy(c->stream);
  free(c->nn_off);
  delete c;
}

static void destroy_ctx_pool() { // acb200_shutdown: contexts still leased by live threads stay with them
  int cur = 0;
  cudaGetDevice(&cur);
  for (int i = 0; i < g_npool; i++) {
    DeviceState &ds = g_ds[g_pool_devs[i]];
    std::lock_guard<std::mutex> lk(ds.mu);
    for (ThreadCtx *c : ds.pool) free_ctx(c);
    ds.pool.clear();
  }
  cudaSetDevice(cur);
}

// ------------------------------------------------------------------ glyph LUTs
// Host restatement of build_utf8_luminance_cache / build_utf8_ramp64_cache (common.c:380-490) reduced to
// what the renderers read: per luminance Y the UTF-8 bytes of the glyph (with the mode's index mapping,
// quirks Q1/Q2 of SURVEY.md §8a) and the mono run key char_index_ramp[Y>>2].
static bool build_lut_host(const char *palette, int which, GlyphLut &L) {
  struct G {
    uint8_t len, b[4];
  } chars[256];
  int n = 0;
  const unsigned char *p = reinterpret_cast<const unsigned char *>(palette);
  while (*p && n < 255) { // lead-byte length rule, common.c:397-410
    int len = (*p & 0xE0) == 0xC0 ? 2 : (*p & 0xF0) == 0xE0 ? 3 : (*p & 0xF8) == 0xF0 ? 4 : 1;
    G g{(uint8_t)len, {0, 0, 0, 0}};
    int i = 0;
    for (; i < len && p[i]; i++) g.b[i] = p[i];
    chars[n++] = g;
    if (i < len) break; // truncated trailing sequence
    p += len;
  }
  if (n == 0) return false;
  auto map256 = [n](int i) { // common.c:420
    int c = n > 1 ? (i * (n - 1) + 127) / 255 : 0;
    return c >= n ? n - 1 : c;
  };
  auto map64 = [n](int i) { // common.c:476
    int c = n > 1 ? (i * (n - 1) + 31) / 63 : 0;
    return c >= n ? n - 1 : c;
  };
  memset(&L, 0, sizeof(L));
  for (int y = 0; y < 256; y++) {
    int ramp = map64(y >> 2);
    int gi = which == 0 ? map256(y)                       // cache[Y]            foreground.c:279,487
             : which == 1 ? map64(ramp < 64 ? ramp : 63)  // cache64[char_idx]   foreground.c:97-102 (Q1)
                          : map256(ramp);                 // cache[char_idx]     foreground.c:596-599 (Q2)
    L.glyph[y][0] = chars[gi].len;
    memcpy(&L.glyph[y][1], chars[gi].b, 4);
    L.key[y] = (uint8_t)ramp;
  }
  return true;
}

// cached per (current device, palette, mapping)
const GlyphLut *device_lut(const char *palette, int which) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
    set_error(E_INVALID_STATE, "no current CUDA device");
    return nullptr;
  }
  DeviceState &ds = g_ds[dev];
  std::string key(1, (char)('0' + which));
  key += palette;
  std::lock_guard<std::mutex> lk(ds.mu);
  auto it = ds.luts.find(key);
  if (it != ds.luts.end()) return it->second;
  GlyphLut h;
  if (!build_lut_host(palette, which, h)) {
    set_error(E_INVALID_STATE, "empty palette");
    return nullptr;
  }
  GlyphLut *d = nullptr;
  if (cudaMalloc((void **)&d, sizeof(GlyphLut)) != cudaSuccess ||
      cudaMemcpy(d, &h, sizeof(GlyphLut), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error(E_MEMORY, "cannot upload glyph LUT");
    return nullptr;
  }
  if (ds.luts.size() >= 2048) { // same bound as the reference's palette cache (common.c:132)
    // another thread may hold a pointer it has not launched with yet: retire this generation and free the one retired
    // a whole generation (2048 new palettes) ago
    for (GlyphLut *old : ds.retired) cudaFree(old);
    ds.retired.clear();
    for (auto &kv : ds.luts) ds.retired.push_back(kv.second);
    ds.luts.clear();
  }
  ds.luts[key] = d;
  return d;
}
// simd_caches_destroy_all / acb200_shutdown.  A launch that is being prepared on another thread may still hold a LUT
// pointer, so the tables are RETIRED here (unreachable for new lookups) and freed by the next destroy: the reference's
// callers tear down after their render threads have stopped (lib/common.c:235), this only keeps a racing call safe.
void destroy_lut_cache() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (int dev = 0; dev < kMaxDevices; dev++) {
    DeviceState &ds = g_ds[dev];
    std::lock_guard<std::mutex> lk(ds.mu);
    if (ds.luts.empty() && ds.retired.empty()) continue;
    cudaSetDevice(dev);
    for (GlyphLut *old : ds.retired) cudaFree(old);
    ds.retired.clear();
    for (auto &kv : ds.luts) ds.retired.push_back(kv.second);
    ds.luts.clear();
  }
  cudaSetDevice(cur);
}

// ------------------------------------------------------------------ client display steps (display.c:484-671)
// color_filter_registry, lib/video/rgba/color_filter.c:23-142: {r, g, b, foreground_on_bg}
static const struct {
  uint8_t r, g, b, on_white;
} k_filter_registry[13] = {{0, 0, 0, 0},       {0, 0, 0, 1},     {255, 255, 255, 0}, {0, 255, 65, 0},  {255, 0, 255, 0},
                           {255, 0, 170, 0},   {255, 136, 0, 0}, {0, 221, 221, 0},   {0, 255, 255, 0}, {255, 182, 193, 0},
                           {255, 51, 51, 0},   {255, 235, 153, 0}, {255, 0, 0, 0}};

// apply_color_filter semantics (color_filter.c:274-346): which arithmetic, which colour.  false = no-op / invalid.
bool resolve_pixel_filter(int filter, float time, int *mode, uint32_t *rgb) {
  *mode = FM_NONE;
  *rgb = 0;
  if (filter <= 0 || filter >= 13) return false;
  if (filter == 12) { // COLOR_FILTER_RAINBOW as a PIXEL filter (direct apply_color_filter call, :286-322)
    uint8_t r, g, b;
    color_filter_calculate_rainbow(time, &r, &g, &b);
    *mode = FM_RAINBOW;
    *rgb = ((uint32_t)r << 16) | ((uint32_t)g << 8) | b;
    return true;
  }
  *mode = k_filter_registry[filter].on_white ? FM_ON_WHITE : FM_SCALE;
  *rgb = ((uint32_t)k_filter_registry[filter].r << 16) | ((uint32_t)k_filter_registry[filter].g << 8) |
         k_filter_registry[filter].b;
  return true;
}

// what session_display_convert_to_ascii does with (flip_x, flip_y, color_filter) around the convert
struct DisplayOps {
  int flip_x = 0, flip_y = 0, filt_mode = FM_NONE;
  uint32_t filt_rgb = 0, fg_over = 0;
};
static DisplayOps display_ops(const acb200_render_cfg_t &cfg) {
  DisplayOps d;
  if ((cfg.flip_x || cfg.flip_y) && cfg.src_w > 1 && cfg.src_h > 1) { // display.c:548
    d.flip_x = cfg.flip_x ? 1 : 0;
    d.flip_y = cfg.flip_y ? 1 : 0;
  }
  if (cfg.color_filter == 12) { // rainbow: the pixels stay, the string's truecolor-fg SGRs are replaced (:640-649)
    uint8_t r, g, b;
    color_filter_calculate_rainbow(cfg.filter_time, &r, &g, &b);
    d.fg_over = 0x01000000u | ((uint32_t)r << 16) | ((uint32_t)g << 8) | b;
  } else {
    resolve_pixel_filter(cfg.color_filter, cfg.filter_time, &d.filt_mode, &d.filt_rgb); // :609-624
  }
  return d;
}

// ------------------------------------------------------------------ plans
static inline size_t al256(size_t v) { return (v + 255) & ~(size_t)255; }

bool make_plan(const acb200_render_cfg_t &cfg, Plan &pl, int leaf_mode) {
  if (cfg.src_w <= 0 || cfg.src_w > 10000 || cfg.src_h <= 0 || cfg.src_h > 10000) { // ascii.c:204
    set_error(E_INVALID_PARAM, "invalid source dimensions %dx%d", cfg.src_w, cfg.src_h);
    return false;
  }
  if (cfg.cols <= 0 || cfg.rows_px <= 0 || cfg.cols > 3840 || cfg.rows_px > 2160) { // image.h:166,179 via image_new
    set_error(E_INVALID_PARAM, "invalid resized dimensions %dx%d", cfg.cols, cfg.rows_px);
    return false;
  }
  if (!cfg.palette) {
    set_error(E_INVALID_PARAM, "palette is NULL");
    return false;
  }
  const bool half = cfg.render_mode == RENDER_MODE_HALF_BLOCK;
  if (half) {
    pl.mode = cfg.color_level == TERM_COLOR_TRUECOLOR ? EM_HB_TRUE
              : cfg.color_level == TERM_COLOR_256     ? EM_HB_256
              : cfg.color_level == TERM_COLOR_16      ? EM_HB_16
                                                      : EM_HB_MONO;
  } else { // ascii.c:981-1001 with SIMD_SUPPORT defined
    pl.mode = cfg.color_level == TERM_COLOR_TRUECOLOR
                  ? (cfg.render_mode == RENDER_MODE_BACKGROUND ? EM_DITHER_BG : EM_TRUE_FG)
              : cfg.color_level == TERM_COLOR_256 ? EM_256_FG
              : cfg.color_level == TERM_COLOR_16  ? EM_16_FG
                                                  : EM_MONO_FG;
  }
  if (leaf_mode >= 0) pl.mode = leaf_mode; // a leaf printer the capability dispatch never selects (dropin.cu)
  if (!half && cfg.palette[0] == '\0') { // get_utf8_palette_cache rejects "" (common.c:275)
    set_error(E_INVALID_STATE, "empty palette");
    return false;
  }
  pl.text_rows = half ? (cfg.rows_px + 1) / 2 : cfg.rows_px;
  pl.lut_which = pl.mode == EM_MONO_FG ? 1 : (pl.mode == EM_16_FG || pl.mode == EM_DITHER_FG_RAMP) ? 2 : 0;
  pl.row_pitch = row_capacity_bytes(pl.mode, cfg.cols, cfg.pad_left);

  int sp = SP_NN;
  bool pixel_filter = false;
  if (cfg.scale == ACB200_SCALE_BOX) {
    const int band = cfg.src_h / cfg.rows_px + 2;
    // a pixel filter is not linear (truncating /255 per pixel), so the band cannot be summed as raw bytes: the filtered
    // streaming form (role-split kernel only) owns whole pixels, 48 bytes per thread and row
    pixel_filter = cfg.color_filter >= 1 && cfg.color_filter <= 11;
    sp = ((3 * cfg.src_w) % 16 == 0 && band <= 256 && (!pixel_filter || cfg.src_w % 16 == 0)) ? SP_BOX_STREAM : SP_BOX_GENERIC;
  } else if (cfg.scale != ACB200_SCALE_NN) {
    set_error(E_INVALID_PARAM, "unknown scale mode %d", cfg.scale);
    return false;
  }
  const int kmode = is_dither_mode(pl.mode) ? (int)EM_256_FG : pl.mode;
  pl.use_smem_out = pl.row_pitch <= (uint32_t)kSmemOutMax;
  size_t sm = rows_smem_total(kmode, sp, cfg.cols, cfg.src_w, pl.use_smem_out ? pl.row_pitch : 0);
  if (sm > kMaxDynSmem && sp == SP_BOX_STREAM) {
    sp = SP_BOX_GENERIC;
    sm = rows_smem_total(kmode, sp, cfg.cols, cfg.src_w, pl.use_smem_out ? pl.row_pitch : 0);
  }
  if (sm > kMaxDynSmem && pl.use_smem_out) {
    pl.use_smem_out = 0;
    sm = rows_smem_total(kmode, sp, cfg.cols, cfg.src_w, 0);
  }
  if (sm > kMaxDynSmem) {
    set_error(E_INVALID_PARAM, "row of %d cells does not fit shared memory", cfg.cols);
    return false;
  }
  pl.scale_path = sp;
  pl.ring_depth = 0;
  // Persistent variants of the streaming box kernel (downscaling geometry whose text row fits the shared staging
  // buffer).  Default: the role-split kernel (streamer warps + emitter warp).  ACB200_BOX_KERNEL=ldg|tma|split
  // selects one explicitly (A/B measurements; "ldg" is the one-tile-per-CTA kernel).
  if (sp == SP_BOX_STREAM && !is_dither_mode(pl.mode) && pl.use_smem_out && cfg.src_h >= cfg.rows_px) {
    static const char *which_env = getenv("ACB200_BOX_KERNEL");
    const char *which = which_env ? which_env : "split";
    if ((pixel_filter || !strcmp(which, "split")) &&
        ws2_smem_total(pl.mode, cfg.cols, cfg.src_w, pl.row_pitch) <= kMaxDynSmem) {
      pl.scale_path = SP_BOX_SPLIT;
    } else if (pixel_filter) {
      pl.scale_path = SP_BOX_GENERIC; // only the role-split kernel has the filtered band sums
    } else if (!strcmp(which, "tma") && ((3 * cfg.src_w) >> 4) <= 2048 && !cfg.flip_x && !cfg.flip_y) {
      int d = ws_ring_depth(pl.mode, cfg.cols, cfg.src_w, pl.row_pitch);
      static const int d_env = getenv("ACB200_RING_DEPTH") ? atoi(getenv("ACB200_RING_DEPTH")) : 0; // tuning knob
      if (d_env >= 2 && d_env <= d) d = d_env;
      if (d >= 2) {
        pl.scale_path = SP_BOX_TMA;
        pl.ring_depth = d;
      }
    }
  }
  if (pixel_filter && pl.scale_path == SP_BOX_STREAM) pl.scale_path = SP_BOX_GENERIC; // one-tile kernel: byte-load form
  pl.frame_capacity = (((size_t)cfg.pad_top + (size_t)pl.text_rows * pl.row_pitch + 1) + 15) & ~(size_t)15;
  pl.rows_bytes = (size_t)pl.text_rows * pl.row_pitch;
  pl.meta_bytes = (size_t)pl.text_rows * sizeof(RowMeta);
  pl.cells_bytes = is_dither_mode(pl.mode) ? (size_t)cfg.cols * cfg.rows_px * 3 : 0;
  pl.err_bytes = is_dither_mode(pl.mode) ? (size_t)cfg.cols * cfg.rows_px * 3 * sizeof(int) : 0;
  return true;
}

size_t scratch_bytes(const Plan &pl, int n) {
  return 256 /* header: tile ticket */ + al256(pl.rows_bytes * n) + al256(pl.meta_bytes * n) +
         al256(pl.cells_bytes * n) + al256(pl.err_bytes * n);
}

// ------------------------------------------------------------------ the device pipeline
int render_device(const acb200_render_cfg_t &cfg, const Plan &pl, const uint8_t *d_frames, size_t frame_stride,
                  int pregathered, int n_frames, uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len,
                  uint8_t *d_scratch, cudaStream_t st, cudaEvent_t k0, cudaEvent_t k1, LookbackState *ls) {
  if (n_frames <= 0) return E_OK;
  const GlyphLut *lut = device_lut(cfg.palette[0] ? cfg.palette : " ", pl.lut_which);
  if (!lut) return t_err;
  uint8_t *rows = d_scratch + 256; // the first 256 bytes hold the tile ticket (fixed place for any plan)
  RowMeta *meta = reinterpret_cast<RowMeta *>(rows + al256(pl.rows_bytes * n_frames));
  uint8_t *cells = reinterpret_cast<uint8_t *>(meta) + al256(pl.meta_bytes * n_frames);
  int *err = reinterpret_cast<int *>(cells + al256(pl.cells_bytes * n_frames));

  RenderParams rp{};
  rp.frames = d_frames;
  rp.frame_stride = frame_stride;
  rp.src_w = cfg.src_w;
  rp.src_h = cfg.src_h;
  rp.pregathered = pregathered;
  rp.cols = cfg.cols;
  rp.rows_px = cfg.rows_px;
  rp.text_rows = pl.text_rows;
  rp.pad_left = cfg.pad_left;
  rp.use_smem_out = pl.use_smem_out;
  rp.row_pitch = pl.row_pitch;
  rp.rows = rows;
  rp.meta = meta;
  rp.lut = lut;
  rp.cells_out = nullptr;
  rp.n_frames = n_frames;
  { // geometry of the packed horizontal sums (render_dev.cuh: cells_box_stream); bands are floor or floor+1 rows
    const int bx = cfg.src_w / cfg.cols, n0 = cfg.src_h / cfg.rows_px;
    const bool ok = bx * cfg.cols == cfg.src_w && (bx & 1) == 0 && n0 >= 1 &&
                    (uint64_t)(bx >> 1) * (uint64_t)(n0 + 1) * 255u < 65536u && (uint64_t)bx * (uint64_t)(n0 + 1) < 4096u;
    rp.box_bx = ok ? bx : 0;
    rp.box_nrow0 = n0;
    for (int d = 0; d < 2; d++) rp.box_M[d] = ok ? 0xFFFFFFFFu / ((uint32_t)bx * (uint32_t)(n0 + d)) + 1u : 0u;
  }
  rp.nn_xr = (uint32_t)((((uint64_t)cfg.src_w << 16) / (uint64_t)cfg.cols) + 1);
  rp.nn_yr = (uint32_t)((((uint64_t)cfg.src_h << 16) / (uint64_t)cfg.rows_px) + 1);
  {
    const DisplayOps d = display_ops(cfg);
    rp.flip_x = d.flip_x;
    rp.flip_y = pregathered ? 0 : d.flip_y; // the host gather already picked the mirrored rows
    rp.filt_mode = d.filt_mode;
    rp.filt_rgb = d.filt_rgb;
    rp.fg_over = d.fg_over;
  }
  // measurement knobs (never set in production): ACB200_TUNE_NOALIAS=1, ACB200_PHASE_A_ONLY=1
  static const int tune_noalias = getenv("ACB200_TUNE_NOALIAS") ? atoi(getenv("ACB200_TUNE_NOALIAS")) : 0;
  static const int phase_a_only = getenv("ACB200_PHASE_A_ONLY") ? atoi(getenv("ACB200_PHASE_A_ONLY")) : 0;
  static const int ws2_noemit = getenv("ACB200_WS2_NOEMIT") ? atoi(getenv("ACB200_WS2_NOEMIT")) : 0;
  static const int ws2_dbg = getenv("ACB200_WS2_DBG") ? atoi(getenv("ACB200_WS2_DBG")) : 0;
  static unsigned long long *d_dbg = nullptr;
  if (ws2_dbg && !d_dbg) {
    cudaMalloc((void **)&d_dbg, 4 * sizeof(unsigned long long));
    cudaMemset(d_dbg, 0, 4 * sizeof(unsigned long long));
  }
  rp.dbg = d_dbg;
  static const int slow_reduce = getenv("ACB200_SLOW_REDUCE") ? atoi(getenv("ACB200_SLOW_REDUCE")) : 0; // A/B knob
  rp.tune_flags = (tune_noalias ? 1 : 0) | (ws2_noemit ? 2 : 0) | (ws2_dbg ? 4 : 0) | (slow_reduce ? 8 : 0);
  if (ws2_dbg && getenv("ACB200_WS2_DBG_PRINT")) { // print-and-reset on demand (set by the measurement script)
    unsigned long long h[4];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
    if (h[3]) fprintf(stderr, "[ws2 dbg] per emitter iteration: streamers wait-empty %.0f cyc, emitter wait-full %.0f cyc, emitter busy %.0f cyc (n=%llu)\n",
                      (double)h[0] / h[3], (double)h[1] / h[3], (double)h[2] / h[3], h[3]);
    cudaMemset(d_dbg, 0, sizeof(h));
  }
  if (phase_a_only && !is_dither_mode(pl.mode)) { // downscale only: rows == nullptr makes the kernel return after phase A
    rp.rows = nullptr;
    if (k0) cudaEventRecord(k0, st);
    rp.direct = 0;
    ACB_CUDA(launch_render_rows(rp, pl.mode, pl.scale_path >= SP_BOX_TMA ? (int)SP_BOX_STREAM : pl.scale_path, st, nullptr));
    if (k1) cudaEventRecord(k1, st);
    count_launch();
    return E_OK;
  }

  // Direct output (default wherever the text row is staged in shared memory): rows are placed in the final arena by a
  // look-back over per-row records, so there is no stitch pass.  ACB200_DIRECT=0 forces scratch rows + k_stitch.
  static const int direct_env = getenv("ACB200_DIRECT") ? atoi(getenv("ACB200_DIRECT")) : 1; // measurement knob
  const bool direct = direct_env && !is_dither_mode(pl.mode) && pl.use_smem_out && pl.scale_path != SP_BOX_TMA;
  rp.direct = direct ? 1 : 0;
  rp.ticket = reinterpret_cast<int *>(d_scratch);
  const bool uses_ticket = direct || pl.scale_path == SP_BOX_SPLIT;
  if (ls && uses_ticket) {
    // library-owned scratch (host path): it was zeroed when allocated; instead of clearing per launch, every launch
    // gets a fresh epoch for the look-back records and knows how many tickets its predecessors took
    if (++ls->epoch == 0) ls->epoch = 1;
    rp.epoch = ls->epoch;
    rp.ticket_base = ls->tickets;
  } else {
    rp.epoch = 1;
    rp.ticket_base = 0;
  }
  if (direct) {
    rp.out = d_out;
    rp.out_pitch = out_pitch;
    rp.out_len = d_out_len;
    rp.pad_top = cfg.pad_top;
    rp.agg = reinterpret_cast<uint4 *>(meta); // 16-byte look-back records live where the stitch path keeps RowMeta
    if (!ls) ACB_CUDA(cudaMemsetAsync(meta, 0, (size_t)n_frames * pl.text_rows * sizeof(uint4), st));
  }
  if (uses_ticket && !ls) ACB_CUDA(cudaMemsetAsync(rp.ticket, 0, sizeof(int), st));
  if (k0) cudaEventRecord(k0, st);
  const int kernel_sp = pl.scale_path >= SP_BOX_TMA ? (int)SP_BOX_STREAM : pl.scale_path;
  if (!is_dither_mode(pl.mode) && pl.scale_path == SP_BOX_TMA) {
    rp.ring_depth = pl.ring_depth;
    ACB_CUDA(launch_render_rows_ws(rp, pl.mode, st));
    count_launch();
  } else if (!is_dither_mode(pl.mode) && pl.scale_path == SP_BOX_SPLIT) {
    unsigned grid = 0;
    ACB_CUDA(launch_render_rows_ws2(rp, pl.mode, st, &grid));
    count_launch();
    if (ls) ls->tickets += (uint32_t)n_frames * (uint32_t)pl.text_rows + grid; // every CTA draws one ticket past the end
  } else if (!is_dither_mode(pl.mode)) {
    unsigned grid = 0; // direct: persistent CTAs that draw tiles from the ticket, one ticket past the end each
    ACB_CUDA(launch_render_rows(rp, pl.mode, kernel_sp, st, &grid));
    count_launch();
    if (ls && direct) ls->tickets += (uint32_t)n_frames * (uint32_t)pl.text_rows + grid;
  } else {
    rp.rows = nullptr;
    rp.cells_out = cells;
    rp.use_smem_out = 0;
    ACB_CUDA(launch_render_rows(rp, EM_256_FG, kernel_sp, st, nullptr)); // resize-only pass
    ACB_CUDA(cudaMemsetAsync(err, 0, pl.err_bytes * n_frames, st));
    ACB_CUDA(launch_dither_bg(cells, cfg.cols, cfg.rows_px, n_frames, cfg.pad_left, lut, rows, pl.row_pitch, meta, err,
                              pl.mode != EM_DITHER_BG, st));
    count_launch(2);
  }
  if (k1) cudaEventRecord(k1, st);
  if (direct) return E_OK;

  StitchParams sp{};
  sp.rows = rows;
  sp.meta = meta;
  sp.row_pitch = pl.row_pitch;
  sp.text_rows = pl.text_rows;
  sp.pad_top = cfg.pad_top;
  sp.mode = pl.mode;
  sp.out = d_out;
  sp.out_pitch = out_pitch;
  sp.out_len = d_out_len;
  sp.rows_per_cta = 8;
  ACB_CUDA(launch_stitch(sp, n_frames, st));
  count_launch();
  return E_OK;
}

// ------------------------------------------------------------------ host-buffer path
// Transfer plans for nearest neighbour.  image_resize reads dst_w x dst_h of the src_w x src_h pixels (image.c:293-325:
// sx = (x * x_ratio) >> 16, sy = (y * y_ratio) >> 16, clamped); which ones is index arithmetic the host can do while
// it stages the frame, so only those bytes cross PCIe:
//   NN_PIXELS  the dst_w x dst_h sampled pixels (184 KB of a 24.9 MB 4K frame at 320 x 192); the device then renders a
//              1:1 image (x_ratio = 65537 -> sx = x), flips already applied by the gather
//   NN_ROWS    the dst_h sampled rows (2.2 MB): when dst_w >= src_w (nothing to save per row), or ACB200_NN_PLAN=rows
//   FULL       everything: box filter, or nearest neighbour that upscales in both axes
enum { PLAN_FULL = 0, PLAN_NN_ROWS = 1, PLAN_NN_PIXELS = 2 };

static inline uint32_t nn_ratio(int src, int dst) { return (uint32_t)((((uint64_t)src << 16) / (uint64_t)dst) + 1); } // image.c:293-294
static inline uint32_t nn_src_index(int i, uint32_t ratio, int src) { // image.c:300-302, 315-317
  uint32_t s = ((uint32_t)i * ratio) >> 16;
  return s >= (uint32_t)src ? (uint32_t)src - 1 : s;
}

static bool nn_column_table(ThreadCtx *cx, int src_w, int cols, int flip_x) {
  if (cx->nn_off && cx->nn_src_w == src_w && cx->nn_cols == cols && cx->nn_flip == flip_x) return true;
  if (cols > cx->nn_off_cap) {
    free(cx->nn_off);
    cx->nn_off = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cols);
    cx->nn_off_cap = cx->nn_off ? cols : 0;
    if (!cx->nn_off) {
      set_error(E_MEMORY, "out of memory for the column table");
      return false;
    }
  }
  const uint32_t xr = nn_ratio(src_w, cols);
  for (int x = 0; x < cols; x++) {
    uint32_t sx = nn_src_index(x, xr, src_w);
    if (flip_x) sx = (uint32_t)src_w - 1u - sx; // display.c:563-577: the image is mirrored before it is resized
    cx->nn_off[x] = sx * 3u;
  }
  cx->nn_src_w = src_w;
  cx->nn_cols = cols;
  cx->nn_flip = flip_x;
  return true;
}

// dst receives cols x rows packed RGB24 (+ up to 1 byte of slack: pixels are moved as overlapping 4-byte words).
// A sampled source row is used exactly once and its neighbours not at all (192 of 2160 rows at C3, ~11 rows = 129 KB
// apart).  ACB200_GATHER_AHEAD=k asks for the row k ahead while the current one is copied (non-temporal hint): on an
// AMD EPYC (Zen 5) host that halves the gather (155 -> 73 us per 4K frame), on the Xeon hosts of the B200 boxes it
// DOUBLES it (131-142 -> 242-283 us, profiles/r02m_gather_prefetch_ab.txt) — software prefetches compete with the
// demand loads for the same fill buffers there — so the default is 0 (none).
static void gather_nn_pixels(const uint8_t *src, int src_w, int src_h, int cols, int rows, bool flip_y,
                             const uint32_t *off, uint8_t *dst) {
  const size_t R = (size_t)src_w * 3;
  const uint32_t yr = nn_ratio(src_h, rows);
  static const int kAhead = getenv("ACB200_GATHER_AHEAD") ? atoi(getenv("ACB200_GATHER_AHEAD")) : 0; // 0 = no prefetch
  auto src_row = [&](int y) -> const uint8_t * {
    uint32_t sy = nn_src_index(y, yr, src_h);
    if (flip_y) sy = (uint32_t)src_h - 1u - sy;
    return src + (size_t)sy * R;
  };
  for (int y = 0; y < kAhead && y < rows; y++) {
    const uint8_t *row = src_row(y);
    for (int x = 0; x < cols; x++) _mm_prefetch((const char *)row + off[x], _MM_HINT_NTA);
  }
  if (kAhead <= 0) { // plain form
    for (int y = 0; y < rows; y++) {
      const uint8_t *row = src_row(y);
      uint8_t *d = dst + (size_t)y * cols * 3u;
      if (row + R < src + R * (size_t)src_h) {
        for (int x = 0; x < cols; x++) {
          uint32_t v;
          memcpy(&v, row + off[x], 4);
          memcpy(d + 3 * x, &v, 4);
        }
      } else {
        for (int x = 0; x < cols; x++) {
          const uint8_t *q = row + off[x];
          d[3 * x] = q[0];
          d[3 * x + 1] = q[1];
          d[3 * x + 2] = q[2];
        }
      }
    }
    return;
  }
  for (int y = 0; y < rows; y++) {
    const uint8_t *row = src_row(y);
    const uint8_t *ahead = (kAhead > 0 && y + kAhead < rows) ? src_row(y + kAhead) : row;
    uint8_t *d = dst + (size_t)y * cols * 3u;
    if (row + R < src + R * (size_t)src_h) { // a 4-byte read of the row's last pixel stays inside the image
      for (int x = 0; x < cols; x++) {
        _mm_prefetch((const char *)ahead + off[x], _MM_HINT_NTA);
        uint32_t v;
        memcpy(&v, row + off[x], 4);
        memcpy(d + 3 * x, &v, 4);
      }
    } else {
      for (int x = 0; x < cols; x++) {
        const uint8_t *q = row + off[x];
        d[3 * x] = q[0];
        d[3 * x + 1] = q[1];
        d[3 * x + 2] = q[2];
      }
    }
  }
}

// device-side address of `p` if it lies in page-locked host memory the current device can read, else nullptr
static const uint8_t *pinned_dev_ptr(const void *p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  return a.type == cudaMemoryTypeHost ? (const uint8_t *)a.devicePointer : nullptr;
}
static bool is_pinned(const void *p) { return pinned_dev_ptr(p) != nullptr; }

// The source rows NN sampling reads, in ascending order (with flip_y the list is walked backwards: row j of the list is
// output row rows-1-j), as P interleaved arithmetic progressions: list[j + P] - list[j] == D for every j.  Any integer
// or k/4-type ratio has a small P (2160 -> 192: rows 0,11,22,33,45,...: P = 4, D = 45); geometries without one keep the
// host gather.
struct RowSchedule {
  int src_h = 0, rows = 0, flip = -1, P = 0, D = 0; // P == 0: no progression with P <= kMaxP
  int first[8];
};
static constexpr int kMaxP = 8;
static void row_schedule(RowSchedule &rs, int src_h, int rows, bool flip_y) {
  if (rs.src_h == src_h && rs.rows == rows && rs.flip == (flip_y ? 1 : 0)) return;
  rs.src_h = src_h;
  rs.rows = rows;
  rs.flip = flip_y ? 1 : 0;
  rs.P = 0;
  const uint32_t yr = nn_ratio(src_h, rows);
  auto at = [&](int j) -> int { // ascending
    if (!flip_y) return (int)nn_src_index(j, yr, src_h);
    return src_h - 1 - (int)nn_src_index(rows - 1 - j, yr, src_h);
  };
  for (int P = 1; P <= kMaxP && P <= rows; P++) {
    if (rows <= P) { // fewer rows than progressions: each is a single row
      rs.P = rows;
      rs.D = 1;
      for (int k = 0; k < rows; k++) rs.first[k] = at(k);
      return;
    }
    const int D = at(P) - at(0);
    bool ok = D > 0;
    for (int j = 0; ok && j + P < rows; j++) ok = at(j + P) - at(j) == D;
    if (ok) {
      rs.P = P;
      rs.D = D;
      for (int k = 0; k < P; k++) rs.first[k] = at(k);
      return;
    }
  }
}

// Who fetches: 1 = k_gather_nn_rows reads the rows out of the mapped host memory itself (default: one launch, any
// geometry, 34 GB/s of source rows per link with 16 callers), 2 = the same with 256-byte L2 fetches (no difference
// measured), 0 = the copy engine moves the rows as strided 2-D copies and the kernel samples the columns on the device
// (28 GB/s: four copies per frame serialise on the engine).  profiles/r02o_fetch_variants.txt.  ACB200_FETCH=ce|sm|sm256.
static int fetch_kind() {
  static const int kind = [] {
    const char *e = getenv("ACB200_FETCH");
    return !e ? 1 : !strcmp(e, "ce") ? 0 : !strcmp(e, "sm256") ? 2 : 1;
  }();
  return kind;
}

struct FetchSlot { // one of the per-GPU "the device fetches my frame" slots (g_fetch_depth), released on scope exit
  std::atomic<int> *ctr = nullptr;
  bool held = false;
  bool try_take(std::atomic<int> *c) {
    const int depth = g_fetch_depth.load(std::memory_order_relaxed);
    if (depth == 0) return false;
    if (depth < 0) return held = true;
    if (c->fetch_add(1, std::memory_order_relaxed) >= depth) {
      c->fetch_sub(1, std::memory_order_relaxed);
      return false;
    }
    ctr = c;
    return held = true;
  }
  ~FetchSlot() {
    if (ctr) ctr->fetch_sub(1, std::memory_order_relaxed);
  }
};

static int render_batch_host_impl(const acb200_render_cfg_t &cfg, const uint8_t *const *frames, int n_frames,
                                  char **out, size_t *out_len, int leaf_mode = -1) {
  Plan pl0;
  if (!make_plan(cfg, pl0, leaf_mode)) return t_err; // validates the caller's configuration as given
  if (n_frames <= 0) return E_OK;         // an empty batch is legal
  ThreadCtx *cx = thread_ctx();
  if (!cx) return t_err;
  const size_t R = (size_t)cfg.src_w * 3;
  static const char *plan_env = getenv("ACB200_NN_PLAN"); // measurement knob: "rows" = the row-granular plan only
  const bool rows_only = plan_env && !strcmp(plan_env, "rows");
  const bool flips = (cfg.flip_x || cfg.flip_y) && cfg.src_w > 1 && cfg.src_h > 1; // display.c:548
  const bool flip_x = flips && cfg.flip_x, flip_y = flips && cfg.flip_y;
  int tplan = PLAN_FULL;
  if (cfg.scale == ACB200_SCALE_NN) {
    if (cfg.cols < cfg.src_w && !rows_only) tplan = PLAN_NN_PIXELS;
    else if (cfg.rows_px < cfg.src_h) tplan = PLAN_NN_ROWS;
  }
  // what the device sees: the gathered image at 1:1 (NN_PIXELS), else the caller's geometry
  acb200_render_cfg_t dcfg = cfg;
  Plan pl = pl0;
  if (tplan == PLAN_NN_PIXELS) {
    dcfg.src_w = cfg.cols;
    dcfg.src_h = cfg.rows_px;
    dcfg.flip_x = dcfg.flip_y = 0;
    if (!make_plan(dcfg, pl, leaf_mode)) return t_err;
    if (!nn_column_table(cx, cfg.src_w, cfg.cols, flip_x ? 1 : 0)) return t_err;
  }
  // pinned frames: let the GPU do the sampling reads, if one of the device's fetch slots is free
  FetchSlot fetch;
  static thread_local RowSchedule rsched;
  if (tplan == PLAN_NN_PIXELS && R + 32u <= 48u * 1024u && g_fetch_depth.load(std::memory_order_relaxed) != 0) {
    bool all_pinned = true;
    for (int i = 0; i < n_frames && all_pinned; i++) all_pinned = frames[i] && is_pinned(frames[i]);
    if (all_pinned) {
      row_schedule(rsched, cfg.src_h, cfg.rows_px, flip_y);
      if (rsched.P > 0 || fetch_kind() != 0) fetch.try_take(&g_ds[cx->device].fetch_inflight);
    }
  }
  const bool dev_fetch = fetch.held;
  const size_t in_per_frame = tplan == PLAN_NN_PIXELS ? (size_t)cfg.cols * cfg.rows_px * 3
                              : tplan == PLAN_NN_ROWS ? R * cfg.rows_px
                                                      : R * cfg.src_h;
  const size_t in_pitch = al256(in_per_frame + 4); // + slack for the gather's overlapping word stores
  const size_t cap = pl.frame_capacity;
  int chunk = (int)((size_t)(64u << 20) / (in_pitch + cap));
  if (chunk < 1) chunk = 1;
  if (chunk > n_frames) chunk = n_frames;
  const int nchunks = (n_frames + chunk - 1) / chunk;
  const int nslot = nchunks > 1 ? 2 : 1; // two slots: the host drains chunk k-1 while the GPU works on chunk k
  const size_t in_slot = in_pitch * chunk, out_slot = al256(cap * chunk), len_slot = al256(4u * chunk);
  for (int i = 0; i < n_frames; i++) out[i] = nullptr;
  bool need_stage = tplan != PLAN_FULL && !dev_fetch;
  for (int i = 0; i < n_frames && !need_stage && !dev_fetch; i++) need_stage = frames[i] && !is_pinned(frames[i]);
  const size_t scratch_cap_before = cx->d_scratch_cap; // a regrown buffer may come back at the same address
  // How the strings leave the device.  Default: the emitters store them straight into mapped pinned host memory.
  // ACB200_D2H=ce (measurement knob): the emitters write to HBM, the lengths come back first, then the copy engine moves
  // exactly the bytes of each string (full-size PCIe payloads, one more wait per chunk).
  static const bool d2h_ce = getenv("ACB200_D2H") && !strcmp(getenv("ACB200_D2H"), "ce");
  // How the gathered pixels reach the device.  Default: one H2D copy of the staged chunk.  ACB200_H2D=zc (measurement
  // knob): the sampler reads them where the gather left them, in mapped pinned memory (every pixel is read once).
  static const bool h2d_zc = getenv("ACB200_H2D") && !strcmp(getenv("ACB200_H2D"), "zc");
  const bool zero_copy_in = h2d_zc && tplan == PLAN_NN_PIXELS && !dev_fetch;
  uint64_t ph[4] = {0, 0, 0, 0}; // gather/staging, enqueue, wait, copy-out
  if (sync_foreign(cx, cx->stream) != E_OK) return t_err;
  if (d2h_ce && !grow_device(&cx->d_out, &cx->d_out_cap, out_slot * nslot)) return t_err;
  const size_t rows_per_frame = al256(R * (size_t)cfg.rows_px);
  if (dev_fetch && !grow_device(&cx->d_rows, &cx->d_rows_cap, rows_per_frame * chunk * nslot)) return t_err;
  if (!grow_device(&cx->d_in, &cx->d_in_cap, in_slot * nslot) ||
      !grow_device(&cx->d_scratch, &cx->d_scratch_cap, scratch_bytes(pl, chunk)) ||
      !grow_pinned(&cx->h_out, &cx->h_out_cap, out_slot * nslot) ||
      !grow_pinned((uint8_t **)&cx->h_len, &cx->h_len_cap, len_slot * nslot) ||
      (need_stage && !grow_pinned(&cx->h_in, &cx->h_in_cap, in_slot * nslot)))
    return t_err;
  if (cx->d_scratch_cap != scratch_cap_before) cx->scratch_dirty = true;
  // library-owned scratch keeps its look-back state across calls (epochs + ticket base) as long as nothing else
  // wrote into it; otherwise it is zeroed once here and the state restarts
  {
    const bool will_be_direct = !is_dither_mode(pl.mode) && pl.use_smem_out && pl.scale_path != SP_BOX_TMA;
    if (cx->scratch_dirty || !will_be_direct) {
      ACB_CUDA(cudaMemsetAsync(cx->d_scratch, 0, cx->d_scratch_cap, cx->stream));
      cx->lb = LookbackState();
      cx->scratch_dirty = !will_be_direct; // a stitch-path launch leaves RowMeta words behind
    }
  }

  // The emitters write the finished strings and their lengths straight into mapped pinned host memory
  // (UVA: the host pointer is the device pointer), so a chunk costs one wait, no D2H memcpy calls.
  auto issue = [&](int k) -> int {
    const int slot = k & 1, f0 = k * chunk;
    const int n = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
    uint8_t *d_in = cx->d_in + (size_t)slot * in_slot;
    uint8_t *st_base = need_stage ? cx->h_in + (size_t)slot * in_slot : nullptr;
    // gathered frames are small: the whole chunk goes up as one H2D.  Full frames go up one by one, so that the staging
    // copy of frame i+1 overlaps the DMA of frame i.
    const bool all_staged = tplan != PLAN_FULL && !dev_fetch;
    const uint64_t t_issue = now_ns();
    for (int i = 0; i < n; i++) {
      const uint8_t *src = frames[f0 + i];
      if (!src) return set_error(E_INVALID_PARAM, "frame %d is NULL", f0 + i);
      uint8_t *dst = d_in + (size_t)i * in_pitch;
      uint8_t *st = st_base ? st_base + (size_t)i * in_pitch : nullptr;
      if (dev_fetch && fetch_kind() != 0) {
        // measurement knob: the SMs read the rows out of mapped host memory themselves (one launch, any geometry)
        const uint8_t *src_dev = pinned_dev_ptr(src);
        if (!src_dev) return set_error(E_INVALID_STATE, "frame %d is not page-locked", f0 + i);
        ACB_CUDA(launch_gather_nn_rows(src_dev, cfg.src_w, cfg.src_h, cfg.cols, cfg.rows_px, flip_x ? 1 : 0,
                                       flip_y ? 1 : 0, dst, cx->stream, fetch_kind() == 2));
        count_launch(1);
      } else if (dev_fetch) {
        // copy engine: progression k = rows first[k], first[k] + D, ... of the frame -> rows k, k + P, ... of the list
        uint8_t *rows_dev = cx->d_rows + ((size_t)slot * chunk + i) * rows_per_frame;
        const int P = rsched.P;
        for (int k = 0; k < P; k++) {
          const int cnt = (cfg.rows_px - k + P - 1) / P;
          ACB_CUDA(cudaMemcpy2DAsync(rows_dev + (size_t)k * R, (size_t)P * R, src + (size_t)rsched.first[k] * R,
                                     (size_t)rsched.D * R, R, (size_t)cnt, cudaMemcpyHostToDevice, cx->stream));
        }
        // ... and the columns on the device: the list is 1:1 in y (walked backwards when the image is flipped)
        ACB_CUDA(launch_gather_nn_rows(rows_dev, cfg.src_w, cfg.rows_px, cfg.cols, cfg.rows_px, flip_x ? 1 : 0,
                                       flip_y ? 1 : 0, dst, cx->stream));
        count_launch(1);
      } else if (tplan == PLAN_NN_PIXELS) {
        gather_nn_pixels(src, cfg.src_w, cfg.src_h, cfg.cols, cfg.rows_px, flip_y, cx->nn_off, st);
      } else if (tplan == PLAN_NN_ROWS) {
        const uint32_t yr = nn_ratio(cfg.src_h, cfg.rows_px);
        for (int y = 0; y < cfg.rows_px; y++) {
          uint32_t sy = nn_src_index(y, yr, cfg.src_h);
          if (flip_y) sy = (uint32_t)cfg.src_h - 1u - sy;
          memcpy(st + (size_t)y * R, src + (size_t)sy * R, R);
        }
      } else if (is_pinned(src)) {
        ACB_CUDA(cudaMemcpyAsync(dst, src, in_per_frame, cudaMemcpyHostToDevice, cx->stream));
      } else {
        memcpy(st, src, in_per_frame);
        ACB_CUDA(cudaMemcpyAsync(dst, st, in_per_frame, cudaMemcpyHostToDevice, cx->stream));
      }
    }
    const uint64_t t_staged = now_ns();
    if (zero_copy_in)
      d_in = st_base;
    else if (all_staged)
      ACB_CUDA(cudaMemcpyAsync(d_in, st_base, (size_t)(n - 1) * in_pitch + in_per_frame, cudaMemcpyHostToDevice, cx->stream));
    int rc = render_device(dcfg, pl, d_in, in_pitch, tplan == PLAN_NN_ROWS ? 1 : 0, n,
                           (d2h_ce ? cx->d_out : cx->h_out) + (size_t)slot * out_slot,
                           cap, reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(cx->h_len) + (size_t)slot * len_slot),
                           cx->d_scratch, cx->stream, nullptr, nullptr, &cx->lb);
    if (rc) return rc;
    if (nchunks > 1) ACB_CUDA(cudaEventRecord(cx->ev[2 + slot], cx->stream)); // single chunk: collect() waits on the stream
    ph[0] += t_staged - t_issue;
    ph[1] += now_ns() - t_staged;
    return E_OK;
  };
  auto collect = [&](int k) -> int {
    const int slot = k & 1, f0 = k * chunk;
    const int n = (n_frames - f0 < chunk) ? n_frames - f0 : chunk;
    const uint64_t t_wait = now_ns();
    if (nchunks > 1) ACB_CUDA(cudaEventSynchronize(cx->ev[2 + slot]));
    else if (wait_stream(cx) != E_OK) return t_err;
    const uint32_t *lens =
        reinterpret_cast<const uint32_t *>(reinterpret_cast<const uint8_t *>(cx->h_len) + (size_t)slot * len_slot);
    const uint8_t *arena = cx->h_out + (size_t)slot * out_slot;
    if (d2h_ce) {
      for (int i = 0; i < n; i++)
        ACB_CUDA(cudaMemcpyAsync(cx->h_out + (size_t)slot * out_slot + (size_t)i * cap,
                                 cx->d_out + (size_t)slot * out_slot + (size_t)i * cap, lens[i], cudaMemcpyDeviceToHost,
                                 cx->stream));
      if (wait_stream(cx) != E_OK) return t_err;
    }
    const uint64_t t_copy = now_ns();
    ph[2] += t_copy - t_wait;
    for (int i = 0; i < n; i++) {
      const size_t len = lens[i];
      char *sp = (char *)user_alloc(len + 1);
      if (!sp) return set_error(E_MEMORY, "allocator returned NULL for %zu bytes", len + 1);
      memcpy(sp, arena + (size_t)i * cap, len);
      sp[len] = '\0';
      out[f0 + i] = sp;
      if (out_len) out_len[f0 + i] = len;
    }
    ph[3] += now_ns() - t_copy;
    return E_OK;
  };
  int rc = E_OK;
  for (int k = 0; k < nchunks && rc == E_OK; k++) {
    rc = issue(k);
    if (rc == E_OK && k >= 1) rc = collect(k - 1);
  }
  if (rc == E_OK) rc = collect(nchunks - 1);
  if (rc != E_OK) cudaStreamSynchronize(cx->stream); // leave no work in flight that targets our staging
  for (int k = 0; k < 4; k++) g_phase_ns[k].fetch_add(ph[k], std::memory_order_relaxed);
  g_phase_ns[4].fetch_add((uint64_t)n_frames, std::memory_order_relaxed);
  return rc;
}

// one frame that is already resident on the device -> allocator-owned string (server path with resident sources)
char *render_one_device(const acb200_render_cfg_t &cfg, const uint8_t *d_rgb, size_t *out_len, const OneFrameOpts &opts) {
  Plan pl;
  if (!make_plan(cfg, pl)) return nullptr;
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  const size_t scratch_cap_before = cx->d_scratch_cap; // a regrown buffer may come back at the same address
  const int mc = max_crc_chunks(pl.frame_capacity);
  if (sync_foreign(cx, cx->stream) != E_OK) return nullptr;
  if (!grow_device(&cx->d_scratch, &cx->d_scratch_cap, scratch_bytes(pl, 1)) ||
      !grow_pinned(&cx->h_out, &cx->h_out_cap, pl.frame_capacity + 32) ||
      !grow_pinned((uint8_t **)&cx->h_len, &cx->h_len_cap, 256))
    return nullptr;
  // packet path: device arena, followed by the frame-length word (16 words reserved) and the CRC chunk words
  if (opts.packet && !grow_device(&cx->d_frame, &cx->d_frame_cap, pl.frame_capacity + (size_t)(16 + mc) * sizeof(uint32_t)))
    return nullptr;
  uint32_t *d_words = reinterpret_cast<uint32_t *>(cx->d_frame + pl.frame_capacity); // frame_capacity is a multiple of 16
  if (cx->d_scratch_cap != scratch_cap_before) cx->scratch_dirty = true;
  const bool will_be_direct = !is_dither_mode(pl.mode) && pl.use_smem_out && pl.scale_path != SP_BOX_TMA;
  if (cx->scratch_dirty || !will_be_direct) {
    if (cudaMemsetAsync(cx->d_scratch, 0, cx->d_scratch_cap, cx->stream) != cudaSuccess) return nullptr;
    cx->lb = LookbackState();
    cx->scratch_dirty = !will_be_direct;
  }
  // plain: the kernels write the string straight into mapped pinned host memory.  packet: the string stays in HBM for
  // the fix-up and the CRC scan, which streams it to the host buffer (at +32: 8 spare, 24 header, frame 16-aligned).
  uint8_t *arena = opts.packet ? cx->d_frame : cx->h_out;
  uint32_t *lens = opts.packet ? d_words : cx->h_len;
  if (render_device(cfg, pl, d_rgb, (size_t)cfg.src_w * cfg.src_h * 3, 0, 1, arena, pl.frame_capacity, lens,
                    cx->d_scratch, cx->stream, nullptr, nullptr, &cx->lb) != E_OK)
    return nullptr;
  // stream.c:1085-1127 cuts a frame that does not END in ESC[0m after its LAST ESC[0m.  Frames of the 256 / 16 /
  // truecolor-fg / dithered-bg grammars and of the three coloured half-block grammars always end in ESC[0m (per-row or
  // final reset, SURVEY §8a grammar table), so for them the cut is the identity and the launch is skipped.  The two
  // mono grammars emit no SGR of their own, but a palette may contain the bytes, so they keep the device check.
  const bool cut_is_identity = pl.mode != EM_MONO_FG && pl.mode != EM_HB_MONO;
  if (opts.reset_fixup && !cut_is_identity && launch_reset_fixup(arena, pl.frame_capacity, lens, 1, cx->stream) != E_OK)
    return nullptr;
  if (opts.packet && launch_frame_packets(arena, pl.frame_capacity, lens, 1, mc, opts.pk_w, opts.pk_h, d_words + 16,
                                          cx->h_out + 8, 24, cx->h_out + 32, 0, cx->stream) != E_OK)
    return nullptr;
  if (wait_stream(cx) != E_OK) {
    set_error(E_INVALID_STATE, "CUDA failure while rendering a resident frame");
    return nullptr;
  }
  if (opts.packet) {
    const uint8_t *h = cx->h_out + 8;
    const size_t len = ((size_t)h[8] << 24) | ((size_t)h[9] << 16) | ((size_t)h[10] << 8) | h[11]; // original_size
    char *sp = (char *)user_alloc(24 + len + 1);
    if (!sp) return nullptr;
    memcpy(sp, h, 24 + len); // header and frame are adjacent in the staging buffer
    sp[24 + len] = '\0';
    if (out_len) *out_len = 24 + len;
    return sp;
  }
  const size_t len = cx->h_len[0];
  char *sp = (char *)user_alloc(len + 1);
  if (!sp) return nullptr;
  memcpy(sp, cx->h_out, len);
  sp[len] = '\0';
  if (out_len) *out_len = len;
  return sp;
}

char *render_one_host(const acb200_render_cfg_t &cfg, const uint8_t *rgb, size_t *out_len, int leaf_mode) {
  char *s = nullptr;
  size_t n = 0;
  const uint8_t *fr[1] = {rgb};
  if (render_batch_host_impl(cfg, fr, 1, &s, &n, leaf_mode) != E_OK) {
    if (s) user_free(s);
    return nullptr;
  }
  if (out_len) *out_len = n;
  return s;
}

} // namespace acb

// ====================================================================== C ABI (Part 2)
using namespace acb;

extern "C" {

int acb200_init(int device) {
  {
    std::lock_guard<std::mutex> lk(g_dev_mu);
    if (device >= 0 && g_dev_status == -1) g_requested_device = device;
  }
  if (ensure_device() != 0) return t_err ? t_err : E_INVALID_STATE;
  return thread_ctx() ? E_OK : t_err;
}
// One process, several GPUs: the pool calling threads are spread over.  devices == NULL / n <= 0: every visible device.
int acb200_init_devices(const int *devices, int n) {
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (g_dev_status == 0) return set_error(E_INVALID_STATE, "acb200_init_devices: the library is already initialised");
  int all[kMaxDevices], count = 0;
  if (!devices || n <= 0) {
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
      return g_dev_status = set_error(E_INVALID_STATE, "asciichat_b200: no CUDA device; this library has no CPU path");
    if (count > kMaxDevices) count = kMaxDevices;
    for (int i = 0; i < count; i++) all[i] = i;
    devices = all;
    n = count;
  }
  g_dev_status = init_pool_locked(devices, n);
  return g_dev_status;
}
int acb200_device_count(void) { return g_dev_status == 0 ? g_npool : 0; }
int acb200_device_at(int k) { return (g_dev_status == 0 && k >= 0 && k < g_npool) ? g_pool_devs[k] : -1; }
// Pin the calling thread to the k-th device of the pool (k < 0: back to round-robin).  A context the thread already
// holds on another device goes back to that device's pool.
int acb200_bind_thread(int k) {
  if (ensure_device() != 0) return t_err;
  if (k >= g_npool) return set_error(E_INVALID_PARAM, "acb200_bind_thread: %d >= %d devices", k, g_npool);
  t_bind = k;
  if (t_lease.c && k >= 0 && t_lease.c->device != g_pool_devs[k]) {
    cudaSetDevice(t_lease.c->device);
    cudaStreamSynchronize(t_lease.c->stream);
    t_lease.release();
  }
  return thread_ctx() ? E_OK : t_err;
}
int acb200_thread_device(void) {
  ThreadCtx *cx = thread_ctx();
  return cx ? cx->device : -1;
}
int acb200_register_host_memory(void *p, size_t bytes) {
  if (!p || !bytes) return set_error(E_INVALID_PARAM, "acb200_register_host_memory: empty range");
  if (!thread_ctx()) return t_err; // registration needs a current device
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e == cudaErrorHostMemoryAlreadyRegistered) {
    cudaGetLastError();
    return E_OK;
  }
  if (e != cudaSuccess) return set_error(E_MEMORY, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e));
  return E_OK;
}
int acb200_unregister_host_memory(void *p) {
  if (!p) return E_OK;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) return set_error(E_INVALID_PARAM, "cudaHostUnregister: %s", cudaGetErrorString(e));
  return E_OK;
}
void acb200_set_fetch_depth(int depth) { g_fetch_depth.store(depth); }
int acb200_nn_row_schedule(int src_h, int rows, int flip_y, int *period, int *stride, int *first) {
  if (src_h <= 0 || rows <= 0 || !period || !stride || !first) return set_error(E_INVALID_PARAM, "acb200_nn_row_schedule: bad argument");
  RowSchedule rs;
  row_schedule(rs, src_h, rows, flip_y != 0);
  *period = rs.P;
  *stride = rs.D;
  for (int k = 0; k < rs.P; k++) first[k] = rs.first[k];
  return E_OK;
}
void acb200_host_phase_stats(uint64_t out[5], int reset) {
  for (int k = 0; k < 5; k++) {
    if (out) out[k] = g_phase_ns[k].load(std::memory_order_relaxed);
    if (reset) g_phase_ns[k].store(0, std::memory_order_relaxed);
  }
}
// how a caller waits for its frame: 0 spin (cudaStreamSynchronize), 1 sleep on a blocking-sync event, 2 poll for spin_us
// microseconds, then sleep, 3 poll the stream's completion word and yield the core between polls
void acb200_set_sync_mode(int mode, int spin_us) {
  if (mode >= 0 && mode <= 3) g_sync_mode.store(mode);
  if (spin_us >= 0) g_spin_us.store(spin_us);
}
void acb200_shutdown(void) {
  destroy_sources();
  destroy_lut_cache();
  destroy_ctx_pool();
}
int acb200_last_error(void) {
  int e = t_err;
  t_err = 0;
  return e;
}
const char *acb200_last_error_message(void) { return t_errmsg; }
void acb200_set_allocator(void *(*alloc_fn)(size_t), void (*free_fn)(void *)) {
  g_alloc = alloc_fn ? alloc_fn : malloc;
  g_free = free_fn ? free_fn : free;
}
void acb200_set_option_render_mode(int render_mode) { g_opt_render_mode.store(render_mode); }
void acb200_set_default_scale(int scale) { g_default_scale.store(scale); }

size_t acb200_frame_capacity(const acb200_render_cfg_t *cfg) {
  Plan pl;
  if (!cfg || !make_plan(*cfg, pl)) return 0;
  return pl.frame_capacity;
}
size_t acb200_scratch_bytes(const acb200_render_cfg_t *cfg, int n_frames) {
  Plan pl;
  if (!cfg || n_frames <= 0 || !make_plan(*cfg, pl)) return 0;
  return scratch_bytes(pl, n_frames);
}

int acb200_render_batch_device(const acb200_render_cfg_t *cfg, const uint8_t *d_frames, int n_frames, uint8_t *d_out,
                               size_t out_pitch, uint32_t *d_out_len, void *d_scratch, void *stream) {
  if (!cfg || !d_frames || !d_out || !d_out_len || !d_scratch || n_frames < 0)
    return set_error(E_INVALID_PARAM, "acb200_render_batch_device: NULL argument");
  Plan pl;
  if (!make_plan(*cfg, pl)) return t_err;
  if (out_pitch < pl.frame_capacity) return set_error(E_INVALID_PARAM, "out_pitch %zu < capacity %zu", out_pitch, pl.frame_capacity);
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    ThreadCtx *cx = thread_ctx();
    if (!cx) return t_err;
    st = cx->stream;
  } else if (ensure_device() != 0) {
    return t_err;
  }
  return render_device(*cfg, pl, d_frames, (size_t)cfg->src_w * cfg->src_h * 3, 0, n_frames, d_out, out_pitch, d_out_len,
                       (uint8_t *)d_scratch, st);
}

int acb200_synchronize(void) {
  ThreadCtx *cx = thread_ctx();
  if (!cx) return t_err;
  return wait_stream(cx);
}

int acb200_render_batch_host(const acb200_render_cfg_t *cfg, const uint8_t *const *frames, int n_frames, char **out,
                             size_t *out_len) {
  if (!cfg || !frames || !out || n_frames < 0) return set_error(E_INVALID_PARAM, "acb200_render_batch_host: NULL argument");
  int rc = render_batch_host_impl(*cfg, frames, n_frames, out, out_len);
  if (rc != E_OK)
    for (int i = 0; i < n_frames; i++)
      if (out[i]) {
        user_free(out[i]);
        out[i] = nullptr;
      }
  return rc;
}

int acb200_time_batch_device(const acb200_render_cfg_t *cfg, const uint8_t *d_frames, int n_frames, uint8_t *d_out,
                             size_t out_pitch, uint32_t *d_out_len, void *d_scratch, int iters, float *ms_total,
                             float *ms_kernel) {
  if (!cfg || !d_frames || !d_out || !d_out_len || !d_scratch || iters <= 0)
    return set_error(E_INVALID_PARAM, "acb200_time_batch_device: bad argument");
  Plan pl;
  if (!make_plan(*cfg, pl)) return t_err;
  ThreadCtx *cx = thread_ctx();
  if (!cx) return t_err;
  std::vector<cudaEvent_t> ev((size_t)2 * iters);
  for (auto &e : ev) ACB_CUDA(cudaEventCreate(&e));
  ACB_CUDA(cudaStreamSynchronize(cx->stream));
  ACB_CUDA(cudaEventRecord(cx->ev[0], cx->stream));
  int rc = E_OK;
  for (int i = 0; i < iters && rc == E_OK; i++)
    rc = render_device(*cfg, pl, d_frames, (size_t)cfg->src_w * cfg->src_h * 3, 0, n_frames, d_out, out_pitch, d_out_len,
                       (uint8_t *)d_scratch, cx->stream, ev[2 * i], ev[2 * i + 1]);
  cudaEventRecord(cx->ev[1], cx->stream);
  cudaError_t se = cudaStreamSynchronize(cx->stream);
  float tot = 0.f, ker = 0.f;
  if (rc == E_OK && se == cudaSuccess) {
    cudaEventElapsedTime(&tot, cx->ev[0], cx->ev[1]);
    for (int i = 0; i < iters; i++) {
      float k = 0.f;
      cudaEventElapsedTime(&k, ev[2 * i], ev[2 * i + 1]);
      ker += k;
    }
  }
  for (auto &e : ev) cudaEventDestroy(e);
  if (se != cudaSuccess) return set_error(E_INVALID_STATE, "CUDA: %s", cudaGetErrorString(se));
  if (ms_total) *ms_total = tot;
  if (ms_kernel) *ms_kernel = ker;
  return rc;
}

// parity aid: the device's quantisers over the whole colour space, d_out[r<<16|g<<8|b] (16 MiB); which 0 = 256c, 1 = 16c
int acb200_quantize_table_device(int which, uint8_t *d_out, void *stream) {
  if (!d_out || which < 0 || which > 1) return set_error(E_INVALID_PARAM, "acb200_quantize_table_device: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    ThreadCtx *cx = thread_ctx();
    if (!cx) return t_err;
    st = cx->stream;
  } else if (ensure_device() != 0) {
    return t_err;
  }
  ACB_CUDA(launch_quantize_table(which, d_out, st));
  count_launch();
  return E_OK;
}

uint64_t acb200_launch_count(void) { return g_launches.load(); }
const char *acb200_version(void) { return "asciichat_b200 0.1 (sm_100a)"; }

} // extern "C"
