/*
 * oracle/ascii_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference's RGB -> glyph/ANSI render path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library; the product (libasciichat_b200.so) never does.
 *
 * Parity status: PINNED.  Every function is checked in tests/test_oracle.py
 * against oracle/_ref/libasciichat_ref.so (the reference's own sources compiled
 * unmodified) and against the committed golden vectors in tests/golden/.
 */
#ifndef ASCII_ORACLE_H
#define ASCII_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* colour depth / render mode values are the reference's (platform/terminal.h:578-589, 660-667) */
enum { ORC_COLOR_AUTO = -1, ORC_COLOR_NONE = 0, ORC_COLOR_16 = 1, ORC_COLOR_256 = 2, ORC_COLOR_TRUE = 3 };
enum { ORC_MODE_FG = 0, ORC_MODE_BG = 1, ORC_MODE_HALF = 2 };
enum { ORC_SCALE_NN = 0, ORC_SCALE_BOX = 1 };

/* scalar arithmetic */
int orc_luma(int r, int g, int b);                           /* foreground.c:93 */
int orc_rgb_to_256(int r, int g, int b);                     /* ansi.c:360-379 */
int orc_rgb_to_16(int r, int g, int b);                      /* ansi.c:437-477 */
int orc_rep_is_profitable(uint32_t run);                     /* output_buffer.c:148-154 */
int orc_digits_u32(uint32_t v);                              /* util/number.h:62 */
void orc_aspect_ratio(long img_w, long img_h, long width, long height, int stretch, long *out_w,
                      long *out_h);                          /* aspect_ratio.c:70-93 */

/* glyph tables: out[256][5] = {len, b0, b1, b2, b3}, indexed by luminance Y.
 * which: 0 = cache[Y] (256c/truecolor, Q3); 1 = mono double mapping (Q1); 2 = 16-colour mapping (Q2).
 * key_out (may be NULL) receives char_index_ramp[Y>>2] for Y in 0..255 (mono run key).
 * returns glyph count or -1. */
int orc_build_glyph_lut(const char *palette, int which, uint8_t out[256][5], uint8_t key_out[256]);

/* downscale */
void orc_resize_nn(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh);   /* image.c:267-328 */
void orc_resize_box(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh);  /* our spec, DESIGN.md */
void orc_resize_box_fast(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh); /* same result, column sums */

/* print an already-resized image (image_print_with_capabilities, ascii.c:955-1002).
 * Returns malloc'd NUL-terminated string, *out_len = strlen. */
char *orc_print(const uint8_t *rgb, int w, int h, int color_level, int render_mode, const char *palette,
                size_t *out_len);

/* the dithered leaf printers (foreground.c:650-846): variant 0 = ..._with_background(img, true), 1 = ..._with_background(img,
 * false), 2 = image_print_16color_dithered(img) */
char *orc_print_dither(const uint8_t *rgb, int w, int h, const char *palette, int variant, size_t *out_len);

/* full convert (ascii_convert_with_capabilities, ascii.c:194-387), scale = ORC_SCALE_* */
char *orc_convert_caps(const uint8_t *rgb, int w, int h, long width, long height, int color_level, int render_mode,
                       int wants_padding, int use_aspect_ratio, int stretch, const char *palette, int scale,
                       size_t *out_len);
/* ascii_convert (ascii.c:72-191); opt_render_mode stands in for GET_OPTION(render_mode) */
char *orc_convert(const uint8_t *rgb, int w, int h, long width, long height, int color, int aspect, int stretch,
                  const char *palette, int opt_render_mode, size_t *out_len);

char *orc_pad_width(const char *frame, size_t pad_left);     /* ascii.c:457-517 */
char *orc_pad_height(const char *frame, size_t pad_top);     /* ascii.c:902-941 */

/* text-space grid (ascii_create_grid, ascii.c:602-885) */
char *orc_create_grid(const char *const *frames, const size_t *sizes, int n, int width, int height, size_t *out_size);

/* server pixel-space composite (stream.c:523-651, 664-779).  srcs[i] = RGB24 w[i] x h[i].
 * out must hold width * (2*height) * 3 bytes.  Returns 0, fills cols/rows. */
void orc_grid_layout(const int *ws, const int *hs, int n, int term_w, int term_h, int *cols, int *rows);
int orc_composite(const uint8_t *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                  uint8_t *out, int *cols, int *rows);

/* ---- client display path (src/common/session/display.c:484-671): flip -> colour filter -> convert -> rainbow */
enum { /* color_filter_t, include/ascii-chat/platform/terminal.h */
  ORC_FILTER_NONE = 0, ORC_FILTER_BLACK, ORC_FILTER_WHITE, ORC_FILTER_GREEN, ORC_FILTER_MAGENTA, ORC_FILTER_FUCHSIA,
  ORC_FILTER_ORANGE, ORC_FILTER_TEAL, ORC_FILTER_CYAN, ORC_FILTER_PINK, ORC_FILTER_RED, ORC_FILTER_YELLOW,
  ORC_FILTER_RAINBOW, ORC_FILTER_COUNT
};
void orc_calculate_rainbow(float time, uint8_t *r, uint8_t *g, uint8_t *b);          /* color_filter.c:165-236 */
int orc_filter_table(int filter, float time, uint8_t table[256][3]);                 /* color_filter.c:238-346 as a LUT */
int orc_apply_color_filter(uint8_t *pixels, uint32_t width, uint32_t height, uint32_t stride, int filter,
                           float time);                                              /* color_filter.c:274-346 */
char *orc_rainbow_replace(const char *ansi_string, float time_seconds);              /* color_filter.c:348-408 */
char *orc_display_convert(const uint8_t *rgb, int w, int h, long width, long height, int color_level, int render_mode,
                          int wants_padding, int preserve_aspect, int stretch, const char *palette, int flip_x,
                          int flip_y, int filter, float time_seconds, int scale, size_t *out_len);

/* ---- digital rain (lib/video/anim/digital_rain.c): stateful brightness field over a finished frame string */
typedef struct {
  int cols, rows, first, rainbow;
  float *offset, *speed, *prev; /* per-column time offset / speed multiplier (:131-136); per-cell filtered brightness */
  float time, fall_speed, drop_len, decay, anim_speed;
  uint8_t cr, cg, cb;
} orc_rain_t;
orc_rain_t *orc_rain_init(int cols, int rows);
void orc_rain_destroy(orc_rain_t *r);
void orc_rain_set_filter(orc_rain_t *r, int filter);                /* digital_rain_set_color_from_filter */
float orc_rain_target(const orc_rain_t *r, int col, int row, float t); /* get_rain_brightness */
char *orc_rain_apply(orc_rain_t *r, const char *frame, float delta_time, size_t *out_len);

/* ---- wire packaging (lib/network/crc32.c, lib/network/acip/server.c:203-214) */
uint32_t orc_crc32c(const uint8_t *p, size_t n);
void orc_frame_packet_header(const uint8_t *frame, size_t frame_size, uint32_t width, uint32_t height,
                             uint8_t out24[24]);

/* synthetic inputs: 0 noise, 1 gradient, 2 bars, 3 grey, 4 solid(seed&255) */
void orc_gen_pattern(int kind, uint32_t frame, uint8_t *dst, int w, int h);
uint32_t orc_fnv1a32(const uint8_t *p, size_t n);
/* exhaustive quantiser table, index r<<16|g<<8|b; which 0 = 256c, 1 = 16c, 2 = call fn(r,g,b) */
void orc_fill_table(int which, void *fn, uint8_t *out);

/* bench helper: render `frames` frames (frame i uses src + (i % ring) * w*h*3) with `threads`
 * pthreads, frame-parallel; returns seconds, sums output bytes into *out_bytes.
 * fn: 0 = this port, else a function pointer to the compiled reference's
 * ascii_convert_with_capabilities (passed as void* by the harness, with its caps blob). */
double orc_bench_convert(const uint8_t *src, int ring, int w, int h, long width, long height, int color_level,
                         int render_mode, const char *palette, int scale, int frames, int threads,
                         void *ref_fn, const void *ref_caps, uint64_t *out_bytes);

/* box-mode CPU baseline: orc_resize_box_fast + printer (print_fn = compiled reference's image_print_with_capabilities
 * and its caps blob, or NULL = the port's orc_print); frame-parallel, returns seconds */
double orc_bench_box(const uint8_t *src, int ring, int w, int h, int cols, int rows_px, int color_level, int render_mode,
                     const char *palette, int frames, int threads, void *print_fn, const void *caps,
                     uint64_t *out_bytes);

#ifdef __cplusplus
}
#endif
#endif
