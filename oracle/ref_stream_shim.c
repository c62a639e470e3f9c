/*
 * oracle/ref_stream_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Pulls the reference's server-side pixel-space grid compositor into the oracle library: the translation
 * unit below IS the reference's src/server/stream.c (included where it lies, never copied), so its static
 * functions calculate_optimal_grid_layout() (stream.c:523-651) and create_multi_source_composite()
 * (stream.c:664-779) can be called through the two thin wrappers at the bottom, and the whole per-client
 * entry create_mixed_ascii_frame_for_client() (stream.c:958-1191) through ref_oracle_mixed_frame(), which
 * fills g_client_manager with synthetic clients whose "incoming video buffer" is served by the
 * video_frame_get_latest() defined here (the wire layout [be32 w][be32 h][rgb24] collect_video_sources()
 * parses at stream.c:314-431).  The other server globals/functions stream.c references are inert stubs.
 */
#include "src/server/stream.c"

client_manager_t g_client_manager;
atomic_t g_should_exit;
bool atomic_load_bool_impl(const atomic_t *a) { return a->impl != 0; }
#define SHIM_MAX_SRC 32
static video_frame_t shim_frames[SHIM_MAX_SRC];
/* a synthetic client's incoming_video_buffer is the tagged integer (slot+1) */
const video_frame_t *video_frame_get_latest(video_frame_buffer_t *vfb) {
  uintptr_t k = (uintptr_t)vfb;
  if (k == 0 || k > SHIM_MAX_SRC || !shim_frames[k - 1].data) return NULL;
  return &shim_frames[k - 1];
}
const char *named_register(uintptr_t key, const char *base_name, const char *type, const char *format_spec,
                           const char *file, int line, const char *func, uintptr_t parent_key) {
  (void)key; (void)type; (void)format_spec; (void)file; (void)line; (void)func; (void)parent_key;
  return base_name;
}
int packet_queue_enqueue(packet_queue_t *queue, packet_type_t type, const void *data, size_t data_len,
                         uint32_t client_id, bool copy_data) {
  (void)queue; (void)type; (void)data; (void)data_len; (void)client_id; (void)copy_data;
  return -1;
}

static int wrap_sources(image_source_t *s, image_t *imgs, const unsigned char *const *srcs, const int *ws, const int *hs,
                        int n) {
  if (n > 32) return -1;
  for (int i = 0; i < n; i++) {
    imgs[i].w = ws[i];
    imgs[i].h = hs[i];
    imgs[i].pixels = (rgb_pixel_t *)srcs[i];
    imgs[i].alloc_method = 0;
    memset(&s[i], 0, sizeof(s[i]));
    s[i].image = &imgs[i];
    s[i].has_video = true;
  }
  return 0;
}

/* out must hold width * 2*height * 3 bytes */
int ref_oracle_composite(const unsigned char *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                         unsigned char *out) {
  image_source_t s[32];
  image_t imgs[32];
  if (wrap_sources(s, imgs, srcs, ws, hs, n)) return -1;
  image_t *c = create_multi_source_composite(s, n, n, "oracle", (unsigned short)width, (unsigned short)height);
  if (!c) return -1;
  memcpy(out, c->pixels, (size_t)c->w * (size_t)c->h * 3);
  image_destroy_to_pool(c);
  return 0;
}

void ref_oracle_grid_layout(const int *ws, const int *hs, int n, int term_w, int term_h, int *cols, int *rows) {
  image_source_t s[32];
  image_t imgs[32];
  static const unsigned char *none[32];
  if (wrap_sources(s, imgs, none, ws, hs, n)) return;
  calculate_optimal_grid_layout(s, n, n, term_w, term_h, cols, rows);
}

/* The reference's full per-client server entry.  srcs[i] == NULL models a connected client with no video.
 * Client 0 is the target (its caps/palette drive the conversion).  Returns the reference's malloc'd frame
 * (free with free()) or NULL; *out_size as the reference sets it; *sources = sources_with_video. */
char *ref_oracle_mixed_frame(const unsigned char *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                             const terminal_capabilities_t *caps, const char *palette, size_t *out_size, int *sources) {
  if (n < 1 || n > SHIM_MAX_SRC || n > MAX_CLIENTS) return NULL;
  memset(&g_client_manager, 0, sizeof(g_client_manager));
  for (int i = 0; i < n; i++) {
    client_info_t *c = &g_client_manager.clients[i];
    snprintf(c->client_id, sizeof(c->client_id), "oracle.%d", i);
    c->active.impl = 1;
    c->is_sending_video.impl = srcs[i] != NULL;
    c->incoming_video_buffer = (video_frame_buffer_t *)(uintptr_t)(i + 1);
    memset(&shim_frames[i], 0, sizeof(shim_frames[i]));
    if (srcs[i]) {
      size_t px = (size_t)ws[i] * (size_t)hs[i] * 3;
      unsigned char *d = malloc(px + 8);
      uint32_t wn = HOST_TO_NET_U32((uint32_t)ws[i]), hn = HOST_TO_NET_U32((uint32_t)hs[i]);
      memcpy(d, &wn, 4);
      memcpy(d + 4, &hn, 4);
      memcpy(d + 8, srcs[i], px);
      shim_frames[i].data = d;
      shim_frames[i].size = px + 8;
    }
  }
  client_info_t *t = &g_client_manager.clients[0];
  t->terminal_caps = *caps;
  t->has_terminal_caps = true;
  t->client_palette_initialized = true;
  snprintf(t->client_palette_chars, sizeof(t->client_palette_chars), "%s", palette);
  bool changed = false;
  char *out = create_mixed_ascii_frame_for_client("oracle.0", (unsigned short)width, (unsigned short)height, false,
                                                  out_size, &changed, sources);
  for (int i = 0; i < n; i++) {
    free(shim_frames[i].data);
    shim_frames[i].data = NULL;
  }
  return out;
}
