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

#include "engine.h"

using namespace acb;

namespace {

struct Slot {
  uint8_t *buf[2] = {nullptr, nullptr};
  size_t cap[2] = {0, 0};
  int front = 0;
  int w = 0, h = 0;
  bool valid = false;
  std::mutex writer; // serialises updates of one slot
};

Slot g_slots[ACB200_MAX_SOURCES];
std::shared_mutex g_table;

bool slot_ok(int slot) { return slot >= 0 && slot < ACB200_MAX_SOURCES; }

} // namespace

namespace acb {
void destroy_sources() {
  std::unique_lock<std::shared_mutex> lk(g_table);
  for (Slot &s : g_slots) {
    for (int b = 0; b < 2; b++) {
      if (s.buf[b]) cudaFree(s.buf[b]);
      s.buf[b] = nullptr;
      s.cap[b] = 0;
    }
    s.valid = false;
    s.w = s.h = 0;
  }
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
  const size_t bytes = (size_t)w * h * 3;
  const int back = s.front ^ 1;
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, bytes) || !grow_device(&s.buf[back], &s.cap[back], bytes + 16))
    return acb200_last_error();
  memcpy(cx->h_in, rgb, bytes);
  ACB_CUDA(cudaMemcpyAsync(s.buf[back], cx->h_in, bytes, cudaMemcpyHostToDevice, cx->stream));
  ACB_CUDA(cudaStreamSynchronize(cx->stream));
  std::unique_lock<std::shared_mutex> lk(g_table);
  s.front = back;
  s.w = w;
  s.h = h;
  s.valid = true;
  return E_OK;
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
  int ws[ACB200_MAX_SOURCES], hs[ACB200_MAX_SOURCES];
  int live = 0;
  for (int i = 0; i < n; i++) {
    const Slot &s = g_slots[slots[i]];
    if (!s.valid) continue;
    src[live] = s.buf[s.front];
    ws[live] = s.w;
    hs[live] = s.h;
    live++;
  }
  if (out_sources_count) *out_sources_count = live;
  if (live == 0) return nullptr; // stream.c:1036: no frame, not an error
  if (!caps || !palette) {       // convert_composite_to_ascii: caps/palette of the target not known yet (:816-826)
    set_error(E_INVALID_STATE, "acb200_mixed_frame: terminal capabilities / palette not set");
    return nullptr;
  }

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
    for (int v = 0; v < cp.n; v++) {
      const float src_aspect = (float)ws[v] / (float)hs[v];
      const float cell_visual_aspect = (float)cellw / (float)cellh;
      int tw, th;
      if (src_aspect > cell_visual_aspect) { // :708-716
        tw = cellw;
        th = (int)((cellw / src_aspect) + 0.5f);
      } else {
        th = cellh;
        tw = (int)((cellh * src_aspect) + 0.5f);
      }
      CompositeCell &c = cp.cell[v];
      if (tw <= 0 || th <= 0) continue; // the reference crashes here (NULL image, :723); we leave the cell black
      c.src = src[v];
      c.sw = ws[v];
      c.sh = hs[v];
      c.tw = tw;
      c.th = th;
      c.xp = (cellw - tw) / 2; // :741-742
      c.yp = (cellh - th) / 2;
      c.xr = (uint32_t)((((uint64_t)ws[v] << 16) / (uint64_t)tw) + 1);
      c.yr = (uint32_t)((((uint64_t)hs[v] << 16) / (uint64_t)th) + 1);
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
