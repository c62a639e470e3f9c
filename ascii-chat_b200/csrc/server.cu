// server.cu — the server's per-client frame generation with the clients' video kept RESIDENT in HBM
// (SURVEY.md §8f row 2).  Mirrors src/server/stream.c:958-1191 (create_mixed_ascii_frame_for_client):
//
//   reference                                   here
//   ---------                                   ----
//   client->incoming_video_buffer (host,        acb200_source_update(slot, rgb, w, h): one H2D upload per received
//     double-buffered, stream.c:253/282)          frame into the slot's back buffer, then a pointer swap
//   collect_video_sources (stream.c:221-455)    the slot table (valid slots = sources_with_video)
//   create_single_source_composite (:476-500)   the slot's device buffer itself
//   create_multi_source_composite (:664-779)    k_composite_all: one launch writes the whole W x 2H device composite
//   convert_composite_to_ascii (:789-853)       plan_convert_with_caps(width, HALF_BLOCK ? 2*height : height,
//                                                 aspect=true, stretch=false) + render_one_device
//   trailing-reset fix-up (:1085-1127)          k_trailing_reset_fixup (effects.cu) for the mono grammars; the identity
//                                                 (frame ends in ESC[0m) for all others (engine.cu: render_one_device)
//   acip_send_ascii_frame (acip/server.c:188)   acb200_mixed_frame_packet: CRC32-C scan + 24-byte header (effects.cu)
//
// Every client's render thread calls acb200_mixed_frame concurrently (one per client at 60 fps in the
// reference, src/server/render.c); the N source frames cross PCIe once per update instead of once per
// (client x frame).  Readers hold the table's shared lock across their launches + stream sync; a writer
// uploads into the slot's back buffer without the lock and takes it exclusively only to swap, so the
// buffer it overwrites next time has no reader left (every reader of it finished before the previous swap).
#include <cstring>
#include <mutex>
#include <shared_mutex>
#include <vector>

#include "engine.h"

using namespace acb;

// Several GPUs in one process (acb200_init_devices): slot s lives on device s % pool size — client frames shard
// across the GPUs of the box, SURVEY.md §8e — and is uploaded on a stream of that device.  A viewer whose render thread
// is leased to another GPU reads the sources it needs straight out of the owner's HBM over NVLink (peer access is
// enabled between all devices of the pool); without peer access the frame is first copied to the viewer's GPU.
namespace {

struct Slot {
  uint8_t *buf[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  int front = 0;
  int w = 0, h = 0;
  bool valid = false;
  int device = -1;             // CUDA ordinal the buffers live on (fixed at first use)
  cudaStream_t up = nullptr;   // upload stream on that device
  uint8_t *recv = nullptr;     // pinned receive buffer handed to the transport (acb200_source_acquire)
  size_t recv_cap = 0;
  std::mutex writer;           // serialises updates of one slot
};

Slot g_slots[ACB200_MAX_SOURCES];
std::shared_mutex g_table;

bool slot_ok(int slot) { return slot >= 0 && slot < ACB200_MAX_SOURCES; }

// s.writer held.  Makes the slot's device current; the caller restores its own.
bool slot_device(Slot &s, int slot) {
  if (s.device < 0) {
    const int n = acb200_device_count();
    s.device = n > 0 ? acb200_device_at(slot % n) : -1;
    if (s.device < 0) return false;
  }
  if (cudaSetDevice(s.device) != cudaSuccess) return false;
  if (!s.up && cudaStreamCreateWithFlags(&s.up, cudaStreamNonBlocking) != cudaSuccess) return false;
  return true;
}

// s.writer held, slot device current.  `staged` = the frame is already in pinned memory (no staging copy).
int upload_locked(Slot &s, const uint8_t *rgb, bool staged, int w, int h, ThreadCtx *cx) {
  const size_t bytes = (size_t)w * h * 3;
  const int back = s.front ^ 1;
  if (!grow_device(&s.buf[back], &s.cap[back], bytes + 16)) return acb200_last_error();
  if (staged) {
    ACB_CUDA(cudaMemcpyAsync(s.buf[back], rgb, bytes, cudaMemcpyHostToDevice, s.up));
  } else {
    // pageable source: staged through this thread's pinned buffer in slices, so that the host copy of slice k+1 runs
    // while slice k crosses PCIe (the copy, not the link, is the slower of the two)
    if (!grow_pinned(&cx->h_in, &cx->h_in_cap, bytes)) return acb200_last_error();
    const size_t slice = 512u << 10;
    for (size_t o = 0; o < bytes; o += slice) {
      const size_t n = bytes - o < slice ? bytes - o : slice;
      memcpy(cx->h_in + o, rgb + o, n);
      ACB_CUDA(cudaMemcpyAsync(s.buf[back] + o, cx->h_in + o, n, cudaMemcpyHostToDevice, s.up));
    }
  }
  // spin, not sleep: a blocking-sync event costs ~0.1 ms of wake-up latency per frame (profiles/r02b: 0.21 ms per pinned
  // 720p commit with it, of which the copy itself is 0.06 ms)
  ACB_CUDA(cudaStreamSynchronize(s.up));
  std::unique_lock<std::shared_mutex> lk(g_table);
  s.front = back;
  s.w = w;
  s.h = h;
  s.valid = true;
  return E_OK;
}

} // namespace

namespace acb {
// acb200_shutdown.  Lock order as in an update: the slot's writer mutex, then the table — so no upload is in flight
// into a buffer that is being freed, and no reader holds a pointer (readers hold the shared table lock across their
// launches and the stream wait).
void destroy_sources() {
  int cur = 0;
  cudaGetDevice(&cur);
  for (Slot &s : g_slots) {
    std::lock_guard<std::mutex> w(s.writer);
    std::unique_lock<std::shared_mutex> lk(g_table);
    if (s.device >= 0) cudaSetDevice(s.device);
    for (int b = 0; b < 2; b++) {
      if (s.buf[b]) cudaFree(s.buf[b]);
      s.buf[b] = nullptr;
      s.cap[b] = 0;
    }
    if (s.recv) cudaFreeHost(s.recv);
    s.recv = nullptr;
    s.recv_cap = 0;
    if (s.up) cudaStreamDestroy(s.up);
    s.up = nullptr;
    s.device = -1;
    s.valid = false;
    s.w = s.h = 0;
  }
  cudaSetDevice(cur);
}
} // namespace acb

extern "C" {

int acb200_source_clear(int slot) {
  if (!slot_ok(slot)) return set_error(E_INVALID_PARAM, "acb200_source_clear: slot %d out of range", slot);
  std::lock_guard<std::mutex> w(g_slots[slot].writer);
  std::unique_lock<std::shared_mutex> lk(g_table);
  g_slots[slot].valid = false;
  return E_OK;
}

int acb200_source_device(int slot) {
  if (!slot_ok(slot)) return -1;
  std::lock_guard<std::mutex> w(g_slots[slot].writer);
  return g_slots[slot].device;
}

int acb200_source_update(int slot, const uint8_t *rgb, int w, int h) {
  if (!slot_ok(slot)) return set_error(E_INVALID_PARAM, "acb200_source_update: slot %d out of range", slot);
  Slot &s = g_slots[slot];
  // collect_video_sources drops a frame whose header says 0, > 4096 wide or > 2160 tall (stream.c:342): that
  // client then contributes no source to this frame.
  if (!rgb || w <= 0 || h <= 0 || w > 4096 || h > 2160) {
    acb200_source_clear(slot);
    return set_error(E_INVALID_PARAM, "acb200_source_update: rejected dimensions %dx%d", w, h);
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  std::lock_guard<std::mutex> wl(s.writer);
  if (!slot_device(s, slot)) {
    cudaSetDevice(cx->device);
    return set_error(E_INVALID_STATE, "acb200_source_update: cannot set up slot %d on its device", slot);
  }
  const bool staged = s.recv && rgb >= s.recv && rgb + (size_t)w * h * 3 <= s.recv + s.recv_cap;
  const int rc = upload_locked(s, rgb, staged, w, h, cx);
  cudaSetDevice(cx->device);
  return rc;
}

// The transport can receive straight into pinned memory: acquire a buffer of `bytes` for the slot, fill it (the RGB24
// payload of one IMAGE_FRAME), then acb200_source_commit uploads it without a staging copy.  The buffer belongs to the
// slot and stays valid until the next acquire that needs a larger one, or acb200_shutdown.
uint8_t *acb200_source_acquire(int slot, size_t bytes) {
  if (!slot_ok(slot) || bytes == 0) {
    set_error(E_INVALID_PARAM, "acb200_source_acquire: bad argument");
    return nullptr;
  }
  if (ensure_device() != 0) return nullptr;
  Slot &s = g_slots[slot];
  std::lock_guard<std::mutex> wl(s.writer);
  if (!grow_pinned(&s.recv, &s.recv_cap, bytes)) return nullptr;
  return s.recv;
}
int acb200_source_commit(int slot, int w, int h) {
  if (!slot_ok(slot)) return set_error(E_INVALID_PARAM, "acb200_source_commit: slot %d out of range", slot);
  Slot &s = g_slots[slot];
  if (w <= 0 || h <= 0 || w > 4096 || h > 2160) { // stream.c:342
    acb200_source_clear(slot);
    return set_error(E_INVALID_PARAM, "acb200_source_commit: rejected dimensions %dx%d", w, h);
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  std::lock_guard<std::mutex> wl(s.writer);
  if (!s.recv || (size_t)w * h * 3 > s.recv_cap)
    return set_error(E_INVALID_PARAM, "acb200_source_commit: no acquired buffer of %zu bytes", (size_t)w * h * 3);
  if (!slot_device(s, slot)) {
    cudaSetDevice(cx->device);
    return set_error(E_INVALID_STATE, "acb200_source_commit: cannot set up slot %d on its device", slot);
  }
  const int rc = upload_locked(s, s.recv, true, w, h, cx);
  cudaSetDevice(cx->device);
  return rc;
}

// The wire form of a received frame, IMAGE_FRAME payload = [width:be32][height:be32][RGB24] — the checks of
// handle_image_frame_packet (src/server/protocol.c:737-889) before it commits the frame: payload >= 8 bytes (:748),
// image_validate_dimensions (1..3840 x 1..2160, lib/util/image.c + image.h:166,179) (:775), exact length
// 8 + w*h*3 (:799).  (The reference additionally drops packets above its 2 MB ring-buffer slot, :833 — a property of
// that buffer, not of the format; resident slots here grow to the frame.)
int acb200_source_update_wire(int slot, const uint8_t *payload, size_t len) {
  if (!slot_ok(slot)) return set_error(E_INVALID_PARAM, "acb200_source_update_wire: slot %d out of range", slot);
  if (!payload || len < 8) return set_error(E_INVALID_PARAM, "IMAGE_FRAME payload too small: %zu bytes", len);
  const uint32_t w = ((uint32_t)payload[0] << 24) | ((uint32_t)payload[1] << 16) | ((uint32_t)payload[2] << 8) | payload[3];
  const uint32_t h = ((uint32_t)payload[4] << 24) | ((uint32_t)payload[5] << 16) | ((uint32_t)payload[6] << 8) | payload[7];
  if (w == 0 || h == 0 || w > 3840u || h > 2160u) return set_error(E_INVALID_PARAM, "IMAGE_FRAME invalid dimensions %ux%u", w, h);
  const size_t expected = 8 + (size_t)w * h * 3;
  if (len != expected) return set_error(E_INVALID_PARAM, "IMAGE_FRAME size mismatch: expected %zu bytes got %zu", expected, len);
  return acb200_source_update(slot, payload + 8, (int)w, (int)h);
}

// stream.c:708-716: the contain-fitted size of source (w, h) inside a cellw x cellh pixel cell
static void fit_in_cell(int w, int h, int cellw, int cellh, int *tw, int *th) {
  const float src_aspect = (float)w / (float)h;
  const float cell_visual_aspect = (float)cellw / (float)cellh;
  if (src_aspect > cell_visual_aspect) {
    *tw = cellw;
    *th = (int)((cellw / src_aspect) + 0.5f);
  } else {
    *th = cellh;
    *tw = (int)((cellh * src_aspect) + 0.5f);
  }
}

// create_single_source_composite / create_multi_source_composite (stream.c:476-500, 664-779) + convert_composite_to_ascii
// (:789-853) + the trailing-reset fix-up (:1085-1127) on `live` device-resident sources of ORIGINAL sizes ws x hs.
// prefit: src[v] already holds source v nearest-neighbour-resized to its fitted size (the multi-GPU pixel-space gather:
// every rank resizes its own clients, only the few-KB cell images travel) — sampling it 1:1 is the reference's blit.
static char *compose_and_render(ThreadCtx *cx, const uint8_t *const *src, const int *ws, const int *hs, int live,
                                bool prefit, unsigned short width, unsigned short height,
                                const terminal_capabilities_t *caps, const char *palette, size_t *out_size, bool packet) {
  const uint8_t *comp = src[0];
  int comp_w = ws[0], comp_h = hs[0];
  if (live > 1) { // create_multi_source_composite, stream.c:664-779
    int gc, gr;
    acb200_grid_layout(ws, hs, live, width, height, &gc, &gr);
    const int CW = width, CH = (int)height * 2;
    const size_t comp_bytes = (size_t)CW * CH * 3;
    if (!grow_device(&cx->d_out, &cx->d_out_cap, comp_bytes + 16)) return nullptr;
    const int cellw = CW / gc, cellh = CH / gr;
    CompositeParams cp{};
    cp.comp = cx->d_out;
    cp.cw = CW;
    cp.ch = CH;
    cp.cellw = cellw;
    cp.cellh = cellh;
    cp.gcols = gc;
    cp.grows = gr;
    cp.n = live < 9 ? live : 9; // max 9 sources (:687)
    for (int v = 0; v < cp.n && cellw > 0 && cellh > 0; v++) {
      int tw, th;
      fit_in_cell(ws[v], hs[v], cellw, cellh, &tw, &th);
      CompositeCell &c = cp.cell[v];
      if (tw <= 0 || th <= 0) continue; // the reference crashes here (NULL image, :723); we leave the cell black
      c.src = src[v];
      c.sw = prefit ? tw : ws[v];
      c.sh = prefit ? th : hs[v];
      c.tw = tw;
      c.th = th;
      c.xp = (cellw - tw) / 2; // :741-742
      c.yp = (cellh - th) / 2;
      c.xr = (uint32_t)((((uint64_t)c.sw << 16) / (uint64_t)tw) + 1);
      c.yr = (uint32_t)((((uint64_t)c.sh << 16) / (uint64_t)th) + 1);
    }
    // clear (image_clear, :683) + every source's NN resize + clipped blit (:723-773) in one launch
    if (cellw <= 0 || cellh <= 0) { // more grid columns/rows than pixels: nothing fits, the canvas stays black
      if (cudaMemsetAsync(cx->d_out, 0, comp_bytes, cx->stream) != cudaSuccess) {
        set_error(E_INVALID_STATE, "CUDA: clearing the composite failed");
        return nullptr;
      }
    } else {
      if (launch_composite_all(cp, cx->stream) != cudaSuccess) {
        set_error(E_INVALID_STATE, "CUDA: composite launch failed");
        return nullptr;
      }
      count_launch();
    }
    comp = cx->d_out;
    comp_w = CW;
    comp_h = CH;
  }

  // convert_composite_to_ascii, stream.c:829-842
  const ssize_t h = caps->render_mode == RENDER_MODE_HALF_BLOCK ? (ssize_t)height * 2 : (ssize_t)height;
  acb200_render_cfg_t cfg;
  if (!plan_convert_with_caps(comp_w, comp_h, width, h, caps, true, false, palette, &cfg)) return nullptr;
  // stream.c:1085-1127 (a frame must end in ESC[0m; otherwise it is cut after its last ESC[0m, if it has one) runs
  // on the device (k_trailing_reset_fixup); packet = also CRC32-C + ascii_frame_packet_t header in front
  OneFrameOpts opts;
  opts.reset_fixup = true;
  opts.packet = packet;
  opts.pk_w = width;
  opts.pk_h = height;
  size_t len = 0;
  char *frame = render_one_device(cfg, comp, &len, opts);
  if (!frame) return nullptr;
  *out_size = len;
  return frame;
}

static char *mixed_frame_impl(const int *slots, int n, unsigned short width, unsigned short height,
                              const terminal_capabilities_t *caps, const char *palette, size_t *out_size,
                              int *out_sources_count, bool packet) {
  if (out_sources_count) *out_sources_count = 0;
  if (!out_size || width == 0 || height == 0) { // stream.c:980
    set_error(E_INVALID_PARAM, "Invalid parameters for acb200_mixed_frame: width=%u, height=%u, out_size=%p", width,
              height, (void *)out_size);
    return nullptr;
  }
  *out_size = 0;
  if ((n > 0 && !slots) || n < 0 || n > ACB200_MAX_SOURCES) {
    set_error(E_INVALID_PARAM, "acb200_mixed_frame: bad slot list");
    return nullptr;
  }
  for (int i = 0; i < n; i++)
    if (!slot_ok(slots[i])) {
      set_error(E_INVALID_PARAM, "acb200_mixed_frame: slot %d out of range", slots[i]);
      return nullptr;
    }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;

  std::shared_lock<std::shared_mutex> lk(g_table);
  const uint8_t *src[ACB200_MAX_SOURCES];
  int ws[ACB200_MAX_SOURCES], hs[ACB200_MAX_SOURCES], devs[ACB200_MAX_SOURCES];
  int live = 0;
  size_t remote_bytes = 0;
  for (int i = 0; i < n; i++) {
    const Slot &s = g_slots[slots[i]];
    if (!s.valid) continue;
    src[live] = s.buf[s.front];
    ws[live] = s.w;
    hs[live] = s.h;
    devs[live] = s.device;
    if (live < 9 && s.device != cx->device && !peer_ok(cx->device, s.device))
      remote_bytes += ((size_t)s.w * s.h * 3 + 255) & ~(size_t)255;
    live++;
  }
  if (out_sources_count) *out_sources_count = live;
  if (live == 0) return nullptr; // stream.c:1036: no frame, not an error
  if (!caps || !palette) {       // convert_composite_to_ascii: caps/palette of the target not known yet (:816-826)
    set_error(E_INVALID_STATE, "acb200_mixed_frame: terminal capabilities / palette not set");
    return nullptr;
  }
  // Sources on another GPU of the pool are read in place over NVLink (peer access).  Where the topology offers no peer
  // access, the (at most nine) frames this viewer composes are first copied to its own GPU.
  if (remote_bytes) {
    if (!grow_device(&cx->d_in, &cx->d_in_cap, remote_bytes)) return nullptr;
    size_t o = 0;
    for (int v = 0; v < live && v < 9; v++) {
      if (devs[v] == cx->device || peer_ok(cx->device, devs[v])) continue;
      const size_t bytes = (size_t)ws[v] * hs[v] * 3;
      if (cudaMemcpyPeerAsync(cx->d_in + o, cx->device, src[v], devs[v], bytes, cx->stream) != cudaSuccess) {
        set_error(E_INVALID_STATE, "CUDA: peer copy of a source frame failed");
        return nullptr;
      }
      src[v] = cx->d_in + o;
      o += (bytes + 255) & ~(size_t)255;
    }
  }

  return compose_and_render(cx, src, ws, hs, live, false, width, height, caps, palette, out_size, packet);
}

// The discovery host's render tick (src/common/session/host.c:664-717) with resident sources, across the GPUs of the pool:
// every client with video is converted (ascii_convert_with_capabilities) ON THE GPU THAT HOLDS ITS FRAMES, the emitters
// store the finished rows straight into the composing GPU's arena (peer stores over NVLink — the gather is part of the
// render kernel's output phase, there is no separate collective; a staged peer copy where the topology has no peer
// access), and ascii_create_grid composes there, reading the string lengths where the kernels left them.
char *acb200_grid_frame(const int *slots, int n, int cell_width, int cell_height, const terminal_capabilities_t *caps,
                        bool use_aspect_ratio, bool stretch, const char *palette, int grid_width, int grid_height,
                        size_t *out_size) {
  if (!out_size || !slots || n <= 0 || n > ACB200_MAX_SOURCES || !caps || !palette || cell_width <= 0 ||
      cell_height <= 0 || grid_width <= 0 || grid_height <= 0) {
    set_error(E_INVALID_PARAM, "acb200_grid_frame: bad argument");
    return nullptr;
  }
  *out_size = 0;
  for (int i = 0; i < n; i++)
    if (!slot_ok(slots[i])) {
      set_error(E_INVALID_PARAM, "acb200_grid_frame: slot %d out of range", slots[i]);
      return nullptr;
    }
  if ((size_t)grid_width * (size_t)grid_height > (size_t)1 << 30) {
    set_error(E_INVALID_PARAM, "acb200_grid_frame: dimensions would overflow: %dx%d", grid_width, grid_height);
    return nullptr;
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  const int D = cx->device;
  struct Job {
    const uint8_t *src;
    int dev;
    acb200_render_cfg_t cfg;
    Plan pl;
  };
  std::vector<Job> jobs;
  std::shared_lock<std::shared_mutex> lk(g_table); // held until the strings are composed: the sources must not move
  size_t pitch = 0;
  for (int i = 0; i < n; i++) {
    const Slot &s = g_slots[slots[i]];
    if (!s.valid) continue; // host.c:672: only participants with video take a cell
    Job j;
    j.src = s.buf[s.front];
    j.dev = s.device;
    if (!plan_convert_with_caps(s.w, s.h, cell_width, cell_height, caps, use_aspect_ratio, stretch, palette, &j.cfg) ||
        !make_plan(j.cfg, j.pl))
      return nullptr;
    if (j.pl.frame_capacity > pitch) pitch = j.pl.frame_capacity;
    jobs.push_back(j);
  }
  const int live = (int)jobs.size();
  if (live == 0) return nullptr; // nothing to show, not an error
  // composing GPU: arena (live x pitch), then the length words
  const size_t arena_bytes = (size_t)live * pitch, canvas = (size_t)grid_width * grid_height + grid_height + 1;
  const size_t out_cap = (canvas > pitch + 1 ? canvas : pitch + 1) + 16;
  if (sync_foreign(cx, cx->stream) != E_OK) return nullptr;
  if (!grow_device(&cx->d_frame, &cx->d_frame_cap, arena_bytes + 256) || !grow_device(&cx->d_out, &cx->d_out_cap, out_cap) ||
      !grow_pinned(&cx->h_out, &cx->h_out_cap, out_cap))
    return nullptr;
  uint8_t *arena = cx->d_frame;
  uint32_t *lens = reinterpret_cast<uint32_t *>(cx->d_frame + arena_bytes); // arena_bytes is a multiple of 16
  bool used[kMaxDevices] = {};
  for (int i = 0; i < live; i++) {
    Job &j = jobs[i];
    cudaStream_t st;
    uint8_t **scratch;
    size_t *scratch_cap;
    LookbackState *lb;
    bool *dirty;
    PeerHelper *h = nullptr;
    if (j.dev == D) {
      cudaSetDevice(D); // the previous job may have left another device current
      st = cx->stream, scratch = &cx->d_scratch, scratch_cap = &cx->d_scratch_cap, lb = &cx->lb, dirty = &cx->scratch_dirty;
    } else {
      h = peer_helper(cx, j.dev);
      if (!h) {
        cudaSetDevice(D);
        return nullptr;
      }
      st = h->stream, scratch = &h->scratch, scratch_cap = &h->scratch_cap, lb = &h->lb, dirty = &h->dirty;
      used[j.dev] = true;
    }
    const size_t cap_before = *scratch_cap;
    bool ok = grow_device(scratch, scratch_cap, scratch_bytes(j.pl, 1));
    const bool will_be_direct = !is_dither_mode(j.pl.mode) && j.pl.use_smem_out && j.pl.scale_path != SP_BOX_TMA;
    if (ok && (*scratch_cap != cap_before || *dirty || !will_be_direct)) {
      ok = cudaMemsetAsync(*scratch, 0, *scratch_cap, st) == cudaSuccess;
      *lb = LookbackState();
      *dirty = !will_be_direct;
    }
    uint8_t *dst = arena + (size_t)i * pitch;
    uint32_t *dlen = lens + i;
    const bool direct_peer = j.dev == D || peer_ok(j.dev, D);
    if (ok && !direct_peer) { // no peer access: render next to the source, then a staged peer copy
      ok = grow_device(&h->out, &h->out_cap, (size_t)live * (pitch + 16));
      dst = h->out + (size_t)i * pitch;
      dlen = reinterpret_cast<uint32_t *>(h->out + (size_t)live * pitch) + i;
    }
    if (ok)
      ok = render_device(j.cfg, j.pl, j.src, (size_t)j.cfg.src_w * j.cfg.src_h * 3, 0, 1, dst, pitch, dlen, *scratch, st,
                         nullptr, nullptr, lb) == E_OK;
    if (ok && !direct_peer)
      ok = cudaMemcpyPeerAsync(arena + (size_t)i * pitch, D, dst, j.dev, pitch, st) == cudaSuccess &&
           cudaMemcpyPeerAsync(lens + i, D, dlen, j.dev, sizeof(uint32_t), st) == cudaSuccess;
    if (!ok) {
      cudaSetDevice(D);
      if (!acb200_last_error()) set_error(E_INVALID_STATE, "acb200_grid_frame: CUDA failure while rendering a cell");
      return nullptr;
    }
  }
  for (int d = 0; d < kMaxDevices; d++)
    if (used[d]) {
      cudaSetDevice(d);
      cudaEventRecord(cx->helper[d]->ev, cx->helper[d]->stream);
    }
  cudaSetDevice(D);
  for (int d = 0; d < kMaxDevices; d++)
    if (used[d]) cudaStreamWaitEvent(cx->stream, cx->helper[d]->ev, 0);
  const uint8_t *ptrs[ACB200_MAX_SOURCES];
  for (int i = 0; i < live; i++) ptrs[i] = arena + (size_t)i * pitch;
  size_t res = 0;
  bool exact = false;
  // host.c:701-702: frame_size = strlen + 1 (the terminator rides along)
  if (text_grid_device(ptrs, nullptr, live, grid_width, grid_height, cx->d_out, &res, cx->stream, cx, lens, 1u, &exact) != E_OK ||
      cudaMemcpyAsync(cx->h_out, cx->d_out, res + 1, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      wait_stream(cx) != E_OK) {
    if (!acb200_last_error()) set_error(E_INVALID_STATE, "acb200_grid_frame: CUDA failure");
    for (int d = 0; d < kMaxDevices; d++) // leave nothing in flight that reads the sources after the lock is dropped
      if (used[d]) cudaStreamSynchronize(cx->helper[d]->stream);
    return nullptr;
  }
  char *r = (char *)user_alloc(res + 1);
  if (!r) return nullptr;
  memcpy(r, cx->h_out, res + 1);
  r[res] = '\0';
  *out_size = exact ? res : strlen(r); // ascii.c:650,705,785 vs :883
  return r;
}

// ---- pixel-space composition from device pointers (the multi-process gather path) ----------------
// The fitted size of source i inside its grid cell for a width x height terminal — what create_multi_source_composite
// resizes it to (stream.c:690-724).  One source: its own size (create_single_source_composite uses the frame as is).
int acb200_mixed_cell_size(const int *ws, const int *hs, int n, int i, unsigned short width, unsigned short height,
                           int *tw, int *th) {
  if (!ws || !hs || !tw || !th || n <= 0 || i < 0 || i >= n || width == 0 || height == 0)
    return set_error(E_INVALID_PARAM, "acb200_mixed_cell_size: bad argument");
  *tw = *th = 0;
  if (n == 1) {
    *tw = ws[0], *th = hs[0];
    return E_OK;
  }
  if (i >= 9) return E_OK; // only nine are placed (stream.c:687)
  int gc, gr;
  acb200_grid_layout(ws, hs, n, width, height, &gc, &gr);
  const int cellw = width / gc, cellh = (int)height * 2 / gr;
  if (cellw > 0 && cellh > 0) fit_in_cell(ws[i], hs[i], cellw, cellh, tw, th);
  if (*tw <= 0 || *th <= 0) *tw = *th = 0;
  return E_OK;
}
// image_resize (image.c:256-328) between device buffers, asynchronous on `stream` (NULL = the thread's own stream)
int acb200_resize_nn_device(const uint8_t *d_src, int sw, int sh, uint8_t *d_dst, int dw, int dh, void *stream) {
  if (!d_src || !d_dst || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0)
    return set_error(E_INVALID_PARAM, "acb200_resize_nn_device: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    ThreadCtx *cx = thread_ctx();
    if (!cx) return acb200_last_error();
    st = cx->stream;
  } else if (ensure_device() != 0) {
    return acb200_last_error();
  }
  ACB_CUDA(launch_resize_nn_only(d_src, sw, sh, d_dst, dw, dh, 0, st));
  count_launch();
  return E_OK;
}
// create_mixed_ascii_frame_for_client's composite + convert on n sources that are on this thread's device already:
// d_srcs[i] = frame i at its ORIGINAL size ws[i] x hs[i] (prefit = 0), or already resized to acb200_mixed_cell_size()
// (prefit = 1; n == 1: the frame itself).  Runs on the thread's own stream; returns like acb200_mixed_frame.
char *acb200_mixed_frame_device(const uint8_t *const *d_srcs, const int *ws, const int *hs, int n, int prefit,
                                unsigned short width, unsigned short height, const terminal_capabilities_t *caps,
                                const char *palette, size_t *out_size) {
  if (!out_size || width == 0 || height == 0 || !d_srcs || !ws || !hs || n <= 0 || n > ACB200_MAX_SOURCES || !caps ||
      !palette) {
    set_error(E_INVALID_PARAM, "acb200_mixed_frame_device: bad argument");
    return nullptr;
  }
  *out_size = 0;
  for (int i = 0; i < n; i++)
    if (ws[i] <= 0 || hs[i] <= 0 || (!d_srcs[i] && i < 9)) {
      set_error(E_INVALID_PARAM, "acb200_mixed_frame_device: source %d invalid", i);
      return nullptr;
    }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  return compose_and_render(cx, d_srcs, ws, hs, n, prefit != 0, width, height, caps, palette, out_size, false);
}

char *acb200_mixed_frame(const int *slots, int n, unsigned short width, unsigned short height,
                         const terminal_capabilities_t *caps, const char *palette, size_t *out_size,
                         int *out_sources_count) {
  return mixed_frame_impl(slots, n, width, height, caps, palette, out_size, out_sources_count, false);
}

// create_mixed_ascii_frame_for_client + acip_send_ascii_frame's packaging (lib/network/acip/server.c:188-236)
uint8_t *acb200_mixed_frame_packet(const int *slots, int n, unsigned short width, unsigned short height,
                                   const terminal_capabilities_t *caps, const char *palette, size_t *out_size,
                                   int *out_sources_count) {
  return reinterpret_cast<uint8_t *>(
      mixed_frame_impl(slots, n, width, height, caps, palette, out_size, out_sources_count, true));
}

} // extern "C"
