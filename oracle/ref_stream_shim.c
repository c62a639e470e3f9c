/*
 * oracle/ref_stream_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Pulls the reference's server-side pixel-space grid compositor into the oracle library: the translation
 * unit below IS the reference's src/server/stream.c (included where it lies, never copied), so its static
 * functions calculate_optimal_grid_layout() (stream.c:523-651) and create_multi_source_composite()
 * (stream.c:664-779) can be called through the two thin wrappers at the bottom.  The handful of server
 * globals/functions stream.c references elsewhere are defined as inert stubs; the wrappers never reach them.
 */
#include "src/server/stream.c"

client_manager_t g_client_manager;
atomic_t g_should_exit;
bool atomic_load_bool_impl(const atomic_t *a) { return a->impl != 0; }
const video_frame_t *video_frame_get_latest(video_frame_buffer_t *vfb) { (void)vfb; return NULL; }
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
