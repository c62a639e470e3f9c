/*
 * oracle/ascii_oracle.c — TEST INFRASTRUCTURE ONLY (see ascii_oracle.h).
 *
 * A from-scratch CPU restatement of the reference render path.  It is deliberately
 * written in the "cell-local rule" form the CUDA kernels use (every cell decides
 * from its own pixels, its left neighbour and its run what bytes it owns), not as
 * the reference's sequential state machines, so that agreement with the compiled
 * reference (oracle/_ref) also proves that reformulation.
 *
 * Each function cites the reference file:line it restates (paths relative to the
 * reference checkout).  Parity: pinned by tests/test_oracle.py + tests/golden/.
 */
#define _GNU_SOURCE
#include "ascii_oracle.h"

#include <limits.h>
#include <math.h>
#include <pthread.h>
#include <stdbool.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ byte sink */
typedef struct {
  char *p;
  size_t n, cap;
} bb_t;

static void bb_need(bb_t *b, size_t extra) {
  if (b->n + extra + 1 <= b->cap) return;
  size_t c = b->cap ? b->cap : 4096;
  while (c < b->n + extra + 1) c += c / 2;
  b->p = (char *)realloc(b->p, c);
  if (!b->p) abort();
  b->cap = c;
}
static void bb_put(bb_t *b, const void *s, size_t n) {
  bb_need(b, n);
  memcpy(b->p + b->n, s, n);
  b->n += n;
}
static void bb_c(bb_t *b, char c) { bb_put(b, &c, 1); }
static void bb_dec(bb_t *b, uint32_t v) { /* decimal, no leading zeros (common.c:546-570, output_buffer.c:92-104) */
  char t[10];
  int i = 0;
  do {
    t[i++] = (char)('0' + v % 10u);
    v /= 10u;
  } while (v);
  while (i--) bb_c(b, t[i]);
}
static char *bb_finish(bb_t *b, size_t *out_len) {
  bb_need(b, 0);
  b->p[b->n] = '\0';
  if (out_len) *out_len = b->n;
  return b->p;
}

/* ------------------------------------------------------------- scalar arithmetic */
int orc_luma(int r, int g, int b) { return (77 * r + 150 * g + 29 * b + 128) >> 8; } /* foreground.c:93 */

int orc_rgb_to_256(int r, int g, int b) { /* ansi.c:360-379 */
  int avg = (r + g + b) / 3;
  int d = abs(r - avg) + abs(g - avg) + abs(b - avg);
  if (d < 30) return 232 + (avg * 23) / 255;
  return 16 + 36 * ((r * 5) / 255) + 6 * ((g * 5) / 255) + (b * 5) / 255;
}

static const uint8_t k_ansi16[16][3] = { /* ansi.c:442-459 */
    {0, 0, 0},       {128, 0, 0},   {0, 128, 0},   {128, 128, 0}, {0, 0, 128},   {128, 0, 128},
    {0, 128, 128},   {192, 192, 192}, {128, 128, 128}, {255, 0, 0},   {0, 255, 0},   {255, 255, 0},
    {0, 0, 255},     {255, 0, 255}, {0, 255, 255}, {255, 255, 255}};

int orc_rgb_to_16(int r, int g, int b) { /* ansi.c:437-477: first minimum wins */
  int best = 0, bestd = INT_MAX;
  for (int i = 0; i < 16; i++) {
    int dr = r - k_ansi16[i][0], dg = g - k_ansi16[i][1], db = b - k_ansi16[i][2];
    int d = dr * dr + dg * dg + db * db;
    if (d < bestd) {
      bestd = d;
      best = i;
    }
  }
  return best;
}

int orc_digits_u32(uint32_t v) { /* util/number.h:62 */
  int d = 1;
  while (v >= 10u) {
    v /= 10u;
    d++;
  }
  return d;
}

int orc_rep_is_profitable(uint32_t run) { /* output_buffer.c:148-154 */
  if (run <= 2) return 0;
  uint32_t k = run - 1;
  return k > (uint32_t)(orc_digits_u32(k) + 3);
}

/* aspect_ratio.c:17-93.  Float expressions kept operand-for-operand (FLT_EVAL_METHOD 0). */
static long fit_w_from_h(long height, long iw, long ih) {
  if (ih == 0) return 1;
  float w = (float)height * (float)iw / (float)ih * 2.0f;
  int r = (int)(0.5f + w);
  return r > 0 ? r : 1;
}
static long fit_h_from_w(long width, long iw, long ih) {
  if (iw == 0) return 1;
  float h = ((float)width / 2.0f) * (float)ih / (float)iw;
  int r = (int)(0.5f + h);
  return r > 0 ? r : 1;
}
void orc_aspect_ratio(long img_w, long img_h, long width, long height, int stretch, long *out_w, long *out_h) {
  if (img_w <= 0 || img_h <= 0) {
    *out_w = 1;
    *out_h = 1;
    return;
  }
  if (stretch) {
    *out_w = width;
    *out_h = height;
    return;
  }
  long wfh = fit_w_from_h(height, img_w, img_h), hfw = fit_h_from_w(width, img_w, img_h);
  if (wfh <= width) {
    *out_w = wfh;
    *out_h = height;
  } else {
    *out_w = width;
    *out_h = hfw;
  }
  if (*out_w <= 0) *out_w = 1;
  if (*out_h <= 0) *out_h = 1;
}

/* ------------------------------------------------------------------ glyph tables */
typedef struct {
  uint8_t len, b[4];
} glyph_t;

/* common.c:380-430 (parse rule 397-410) */
static int parse_palette(const char *pal, glyph_t chars[256]) {
  int n = 0;
  const unsigned char *p = (const unsigned char *)pal;
  while (*p && n < 255) {
    int len = 1;
    if ((*p & 0xE0) == 0xC0) len = 2;
    else if ((*p & 0xF0) == 0xE0) len = 3;
    else if ((*p & 0xF8) == 0xF0) len = 4;
    chars[n].len = (uint8_t)len;
    memset(chars[n].b, 0, 4);
    int i = 0;
    for (; i < len && p[i]; i++) chars[n].b[i] = p[i];
    n++;
    if (i < len) break; /* truncated sequence at end of string (reference would read past the NUL) */
    p += len;
  }
  return n;
}
static int idx256(int i, int n) { /* common.c:420 */
  int c = n > 1 ? (i * (n - 1) + 127) / 255 : 0;
  return c >= n ? n - 1 : c;
}
static int idx64(int i, int n) { /* common.c:476 */
  int c = n > 1 ? (i * (n - 1) + 31) / 63 : 0;
  return c >= n ? n - 1 : c;
}

int orc_build_glyph_lut(const char *palette, int which, uint8_t out[256][5], uint8_t key_out[256]) {
  glyph_t chars[256];
  if (!palette || !palette[0]) return -1;
  int n = parse_palette(palette, chars);
  if (n <= 0) return -1;
  for (int y = 0; y < 256; y++) {
    int ramp = idx64(y >> 2, n); /* char_index_ramp[Y>>2] */
    int gi;
    if (which == 0) gi = idx256(y, n);                    /* Q3: cache[Y] (foreground.c:279,487) */
    else if (which == 1) gi = idx64(ramp < 64 ? ramp : 63, n); /* Q1: cache64[char_idx] (foreground.c:97-102);
                                                             char_idx >= 64 is an out-of-bounds read in the
                                                             reference (palettes > 64 glyphs): clamped here */
    else gi = idx256(ramp, n);                            /* Q2: cache[char_idx] (foreground.c:596-599) */
    out[y][0] = chars[gi].len;
    memcpy(&out[y][1], chars[gi].b, 4);
    if (key_out) key_out[y] = (uint8_t)ramp;
  }
  return n;
}

/* ---------------------------------------------------------------------- downscale */
void orc_resize_nn(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) { /* image.c:267-328 */
  uint32_t xr = (uint32_t)((((uint64_t)sw << 16) / (uint64_t)dw) + 1);
  uint32_t yr = (uint32_t)((((uint64_t)sh << 16) / (uint64_t)dh) + 1);
  for (int y = 0; y < dh; y++) {
    uint32_t sy = ((uint32_t)y * yr) >> 16;
    if (sy >= (uint32_t)sh) sy = (uint32_t)sh - 1;
    for (int x = 0; x < dw; x++) {
      uint32_t sx = ((uint32_t)x * xr) >> 16;
      if (sx >= (uint32_t)sw) sx = (uint32_t)sw - 1;
      memcpy(dst + ((size_t)y * dw + x) * 3, src + ((size_t)sy * sw + sx) * 3, 3);
    }
  }
}

/* Box filter — OUR specification (the reference has none):
 * dst (dx,dy) averages src x in [floor(dx*sw/dw), max(x0+1, floor((dx+1)*sw/dw))), y likewise,
 * per channel (sum + n/2) / n in unsigned integers. */
void orc_resize_box(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh) {
  for (int dy = 0; dy < dh; dy++) {
    int y0 = (int)(((int64_t)dy * sh) / dh), y1 = (int)(((int64_t)(dy + 1) * sh) / dh);
    if (y1 <= y0) y1 = y0 + 1;
    if (y1 > sh) y1 = sh;
    if (y0 >= sh) y0 = sh - 1;
    for (int dx = 0; dx < dw; dx++) {
      int x0 = (int)(((int64_t)dx * sw) / dw), x1 = (int)(((int64_t)(dx + 1) * sw) / dw);
      if (x1 <= x0) x1 = x0 + 1;
      if (x1 > sw) x1 = sw;
      if (x0 >= sw) x0 = sw - 1;
      uint32_t s[3] = {0, 0, 0};
      for (int y = y0; y < y1; y++) {
        const uint8_t *row = src + ((size_t)y * sw + x0) * 3;
        for (int x = x0; x < x1; x++, row += 3) {
          s[0] += row[0];
          s[1] += row[1];
          s[2] += row[2];
        }
      }
      uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
      uint8_t *d = dst + ((size_t)dy * dw + dx) * 3;
      d[0] = (uint8_t)((s[0] + n / 2) / n);
      d[1] = (uint8_t)((s[1] + n / 2) / n);
      d[2] = (uint8_t)((s[2] + n / 2) / n);
    }
  }
}

/* The same specification arranged for a CPU that is asked to be fast (bench.py's box-mode CPU baseline): every source
 * row of a band is added into 16-bit column sums (one streaming pass over the band; the compiler vectorises the byte ->
 * u16 accumulate), then each destination pixel adds its x-range of column sums.  Falls back to orc_resize_box when a
 * band is taller than 257 rows (the u16 sums would overflow).  Checked equal to orc_resize_box in tests/test_oracle.py. */
__attribute__((optimize("O3"), target("avx2"))) void orc_resize_box_fast(const uint8_t *src, int sw, int sh, uint8_t *dst,
                                                                         int dw, int dh) {
  if (sh / dh + 2 > 257) {
    orc_resize_box(src, sw, sh, dst, dw, dh);
    return;
  }
  const size_t R = (size_t)sw * 3;
  uint16_t *col = malloc(sizeof(uint16_t) * R);
  for (int dy = 0; dy < dh; dy++) {
    int y0 = (int)(((int64_t)dy * sh) / dh), y1 = (int)(((int64_t)(dy + 1) * sh) / dh);
    if (y1 <= y0) y1 = y0 + 1;
    if (y1 > sh) y1 = sh;
    if (y0 >= sh) y0 = sh - 1;
    memset(col, 0, sizeof(uint16_t) * R);
    for (int y = y0; y < y1; y++) {
      const uint8_t *restrict row = src + (size_t)y * R;
      uint16_t *restrict c = col;
      for (size_t i = 0; i < R; i++) c[i] = (uint16_t)(c[i] + row[i]);
    }
    for (int dx = 0; dx < dw; dx++) {
      int x0 = (int)(((int64_t)dx * sw) / dw), x1 = (int)(((int64_t)(dx + 1) * sw) / dw);
      if (x1 <= x0) x1 = x0 + 1;
      if (x1 > sw) x1 = sw;
      if (x0 >= sw) x0 = sw - 1;
      uint32_t s0 = 0, s1 = 0, s2 = 0;
      for (int x = x0; x < x1; x++) {
        s0 += col[3 * x];
        s1 += col[3 * x + 1];
        s2 += col[3 * x + 2];
      }
      uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0);
      uint8_t *d = dst + ((size_t)dy * dw + dx) * 3;
      d[0] = (uint8_t)((s0 + n / 2) / n);
      d[1] = (uint8_t)((s1 + n / 2) / n);
      d[2] = (uint8_t)((s2 + n / 2) / n);
    }
  }
  free(col);
}

/* ---------------------------------------------------------------------- emitters */
static void put_sgr_rgb(bb_t *o, int layer /*38|48*/, int r, int g, int b) { /* ansi.c:143-195, output_buffer.c:186-214 */
  bb_put(o, "\033[", 2);
  bb_dec(o, (uint32_t)layer);
  bb_put(o, ";2;", 3);
  bb_dec(o, (uint32_t)r);
  bb_c(o, ';');
  bb_dec(o, (uint32_t)g);
  bb_c(o, ';');
  bb_dec(o, (uint32_t)b);
  bb_c(o, 'm');
}
static void put_sgr_256(bb_t *o, int layer, int idx) { /* ansi.c:326-357 */
  bb_put(o, "\033[", 2);
  bb_dec(o, (uint32_t)layer);
  bb_put(o, ";5;", 3);
  bb_dec(o, (uint32_t)idx);
  bb_c(o, 'm');
}
static void put_sgr_16(bb_t *o, int bg, int idx) { /* ansi.c:384-435: 30-37/90-97, 40-47/100-107 */
  int code = (idx < 8 ? 30 + idx : 90 + (idx - 8)) + (bg ? 10 : 0);
  bb_put(o, "\033[", 2);
  bb_dec(o, (uint32_t)code);
  bb_c(o, 'm');
}
static void put_reset(bb_t *o) { bb_put(o, "\033[0m", 4); }
static void put_rep(bb_t *o, uint32_t extra) { /* output_buffer.c:156-164 */
  bb_put(o, "\033[", 2);
  bb_dec(o, extra);
  bb_c(o, 'b');
}
static void put_glyph(bb_t *o, const uint8_t g[5]) { bb_put(o, &g[1], g[0]); }

/* run bookkeeping shared by all run-length modes: head[x] and len-of-run-containing-x */
static void runs_from_keys(const uint64_t *key, int w, int *head_of, int *run_len) {
  int h = 0;
  for (int x = 0; x < w; x++) {
    if (x == 0 || key[x] != key[x - 1]) h = x;
    head_of[x] = h;
  }
  int end = w;
  for (int x = w - 1; x >= 0; x--) {
    run_len[x] = end - head_of[x];
    if (head_of[x] == x) end = x;
  }
}

/* mono foreground — image_print (foreground.c:27-138), quirk Q1 */
static char *print_mono_fg(const uint8_t *rgb, int w, int h, const char *pal, size_t *out_len) {
  uint8_t lut[256][5], keyl[256];
  if (orc_build_glyph_lut(pal, 1, lut, keyl) < 0) return NULL;
  bb_t o = {0};
  uint64_t *key = malloc(sizeof(uint64_t) * (size_t)w);
  int *hd = malloc(sizeof(int) * (size_t)w), *rl = malloc(sizeof(int) * (size_t)w);
  for (int y = 0; y < h; y++) {
    const uint8_t *row = rgb + (size_t)y * w * 3;
    for (int x = 0; x < w; x++) key[x] = keyl[orc_luma(row[3 * x], row[3 * x + 1], row[3 * x + 2])];
    runs_from_keys(key, w, hd, rl);
    for (int x = 0; x < w; x++) {
      const uint8_t *g = lut[orc_luma(row[3 * hd[x]], row[3 * hd[x] + 1], row[3 * hd[x] + 2])];
      bool rep = orc_rep_is_profitable((uint32_t)rl[x]);
      if (hd[x] == x) {
        put_glyph(&o, g);
        if (rep) put_rep(&o, (uint32_t)rl[x] - 1);
      } else if (!rep) {
        put_glyph(&o, g);
      }
    }
    if (y != h - 1) bb_c(&o, '\n');
  }
  free(key);
  free(hd);
  free(rl);
  return bb_finish(&o, out_len);
}

/* 256-colour foreground — image_print_256color (foreground.c:433-509) */
static char *print_256_fg(const uint8_t *rgb, int w, int h, const char *pal, size_t *out_len) {
  uint8_t lut[256][5];
  if (orc_build_glyph_lut(pal, 0, lut, NULL) < 0) return NULL;
  bb_t o = {0};
  for (int y = 0; y < h; y++) {
    const uint8_t *p = rgb + (size_t)y * w * 3;
    for (int x = 0; x < w; x++, p += 3) {
      put_sgr_256(&o, 38, orc_rgb_to_256(p[0], p[1], p[2]));
      put_glyph(&o, lut[orc_luma(p[0], p[1], p[2])]);
    }
    put_reset(&o);
    if (y < h - 1) bb_c(&o, '\n');
  }
  return bb_finish(&o, out_len);
}

/* 16-colour foreground — image_print_16color (foreground.c:535-624), quirk Q2 */
static char *print_16_fg(const uint8_t *rgb, int w, int h, const char *pal, size_t *out_len) {
  uint8_t lut[256][5];
  if (orc_build_glyph_lut(pal, 2, lut, NULL) < 0) return NULL;
  bb_t o = {0};
  for (int y = 0; y < h; y++) {
    const uint8_t *p = rgb + (size_t)y * w * 3;
    for (int x = 0; x < w; x++, p += 3) {
      put_sgr_16(&o, 0, orc_rgb_to_16(p[0], p[1], p[2]));
      put_glyph(&o, lut[orc_luma(p[0], p[1], p[2])]);
    }
    put_reset(&o);
    if (y < h - 1) bb_c(&o, '\n');
  }
  return bb_finish(&o, out_len);
}

/* truecolor foreground — image_print_color (foreground.c:195-308) + ansi_rle_* (ansi.c:248-314).
 * Cell rule: an ASCII-glyph cell owns an SGR iff no earlier ASCII-glyph cell exists in raster
 * order or that cell's colour differs; multi-byte cells always own an SGR and are invisible
 * to the comparison.  One reset at the very end, '\n' between rows, no per-row reset. */
static char *print_true_fg(const uint8_t *rgb, int w, int h, const char *pal, size_t *out_len) {
  uint8_t lut[256][5];
  if (orc_build_glyph_lut(pal, 0, lut, NULL) < 0) return NULL;
  bb_t o = {0};
  long prev_ascii = -1; /* pixel index of the latest ASCII-glyph cell */
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      long i = (long)y * w + x;
      const uint8_t *p = rgb + i * 3;
      const uint8_t *g = lut[orc_luma(p[0], p[1], p[2])];
      bool ascii = g[0] == 1 && g[1] < 128;
      if (ascii) {
        bool own = prev_ascii < 0 || memcmp(rgb + prev_ascii * 3, p, 3) != 0;
        if (own) put_sgr_rgb(&o, 38, p[0], p[1], p[2]);
        bb_c(&o, (char)g[1]);
        prev_ascii = i;
      } else {
        put_sgr_rgb(&o, 38, p[0], p[1], p[2]);
        put_glyph(&o, g);
      }
    }
    if (y != h - 1) bb_c(&o, '\n');
  }
  put_reset(&o);
  return bb_finish(&o, out_len);
}

/* 16-colour Floyd–Steinberg — the three dithered printers of foreground.c, one raster-order recurrence
 * (rgb_to_16color_dithered, ansi.c:511-583):
 *   variant 0  image_print_16color_dithered_with_background(img, true)  (foreground.c:752-846): bg = dithered colour,
 *              fg = 15/0 by the luminance of the QUANTISED colour, glyph = cache[Y].  Reached for
 *              TRUECOLOR+BACKGROUND in SIMD builds (sgr.c:429-430, quirk Q4).
 *   variant 1  image_print_16color_dithered_with_background(img, false): fg = dithered colour, glyph = cache[Y]
 *   variant 2  image_print_16color_dithered(img) (foreground.c:650-749): fg = dithered colour,
 *              glyph = cache[char_index_ramp[Y>>2]] (the Q2 double mapping of image_print_16color)
 * Serial by nature. */
char *orc_print_dither(const uint8_t *rgb, int w, int h, const char *pal, int variant, size_t *out_len) {
  uint8_t lut[256][5];
  if (!rgb || !pal || w <= 0 || h <= 0) return NULL;
  if (orc_build_glyph_lut(pal, variant == 2 ? 2 : 0, lut, NULL) < 0) return NULL;
  int *err = calloc((size_t)w * h * 3, sizeof(int));
  bb_t o = {0};
  for (int y = 0; y < h; y++) {
    for (int x = 0; x < w; x++) {
      size_t i = (size_t)y * w + x;
      const uint8_t *p = rgb + i * 3;
      int v[3], c[3];
      for (int k = 0; k < 3; k++) {
        v[k] = p[k] + err[i * 3 + k];
        c[k] = v[k] < 0 ? 0 : (v[k] > 255 ? 255 : v[k]);
      }
      int q = orc_rgb_to_16(c[0], c[1], c[2]);
      for (int k = 0; k < 3; k++) {
        int e = v[k] - (int)k_ansi16[q][k];
        if (x + 1 < w) err[(i + 1) * 3 + k] += (e * 7) / 16;
        if (y + 1 < h) {
          if (x - 1 >= 0) err[(i + w - 1) * 3 + k] += (e * 3) / 16;
          err[(i + w) * 3 + k] += (e * 5) / 16;
          if (x + 1 < w) err[(i + w + 1) * 3 + k] += (e * 1) / 16;
        }
      }
      if (variant == 0) {
        int bl = (k_ansi16[q][0] * 77 + k_ansi16[q][1] * 150 + k_ansi16[q][2] * 29) / 256;
        put_sgr_16(&o, 1, q);
        put_sgr_16(&o, 0, bl < 127 ? 15 : 0);
      } else {
        put_sgr_16(&o, 0, q);
      }
      put_glyph(&o, lut[orc_luma(p[0], p[1], p[2])]);
    }
    put_reset(&o);
    if (y < h - 1) bb_c(&o, '\n');
  }
  free(err);
  return bb_finish(&o, out_len);
}
static char *print_dither_bg(const uint8_t *rgb, int w, int h, const char *pal, size_t *out_len) {
  return orc_print_dither(rgb, w, h, pal, 0, out_len);
}

/* half-block family — halfblock.c:48-165 (truecolor), 416-524 (256), 297-405 (16), 184-286 (mono).
 * depth: 3 truecolor, 2 256c, 1 16c, 0 mono shades. */
static char *print_halfblock(const uint8_t *rgb, int w, int h, int depth, size_t *out_len) {
  static const char HB[3] = {(char)0xE2, (char)0x96, (char)0x80};
  static const char SH[4][3] = {{(char)0xE2, (char)0x96, (char)0x91}, {(char)0xE2, (char)0x96, (char)0x92},
                                {(char)0xE2, (char)0x96, (char)0x93}, {(char)0xE2, (char)0x96, (char)0x88}};
  bb_t o = {0};
  if (w <= 0 || h <= 0) return bb_finish(&o, out_len);
  uint64_t *key = malloc(sizeof(uint64_t) * (size_t)w);
  int *hd = malloc(sizeof(int) * (size_t)w), *rl = malloc(sizeof(int) * (size_t)w);
  for (int y = 0; y < h; y += 2) {
    const uint8_t *T = rgb + (size_t)y * w * 3;
    const uint8_t *B = (y + 1 < h) ? T + (size_t)w * 3 : T; /* odd tail: bottom := top */
    for (int x = 0; x < w; x++) {
      const uint8_t *t = T + 3 * x, *b = B + 3 * x;
      if (depth == 3 || depth == 0)
        key[x] = ((uint64_t)t[0] << 40) | ((uint64_t)t[1] << 32) | ((uint64_t)t[2] << 24) | ((uint64_t)b[0] << 16) |
                 ((uint64_t)b[1] << 8) | b[2];
      else if (depth == 2)
        key[x] = ((uint64_t)orc_rgb_to_256(t[0], t[1], t[2]) << 8) | (uint64_t)orc_rgb_to_256(b[0], b[1], b[2]);
      else
        key[x] = ((uint64_t)orc_rgb_to_16(t[0], t[1], t[2]) << 8) | (uint64_t)orc_rgb_to_16(b[0], b[1], b[2]);
    }
    runs_from_keys(key, w, hd, rl);
    for (int x = 0; x < w; x++) {
      /* everything about a run is decided by its head pixel pair (halfblock.c:111, 357, 476) */
      int hx = hd[x];
      const uint8_t *t = T + 3 * hx, *b = B + 3 * hx;
      bool is_head = hx == x;
      bool rep = orc_rep_is_profitable((uint32_t)rl[x]);
      if (depth == 0) {
        int lt = (t[0] * 76 + t[1] * 150 + t[2] * 29) >> 8, lb = (b[0] * 76 + b[1] * 150 + b[2] * 29) >> 8;
        if (lt < 16 && lb < 16) {
          bb_c(&o, ' ');
        } else if (is_head) {
          bb_put(&o, SH[lt >> 6], 3);
          if (rep) put_rep(&o, (uint32_t)rl[x] - 1);
        } else if (!rep) {
          bb_put(&o, SH[lt >> 6], 3);
        }
        continue;
      }
      bool transparent = !(t[0] | t[1] | t[2] | b[0] | b[1] | b[2]);
      /* colour state in front of this run = previous run's colours unless that run was transparent
       * (it reset/cleared them) or this is the first run of the row (state cleared per row) */
      bool prev_set = false;
      uint64_t pk = 0;
      if (hx > 0) {
        int phx = hd[hx - 1];
        const uint8_t *pt = T + 3 * phx, *pb = B + 3 * phx;
        prev_set = (pt[0] | pt[1] | pt[2] | pb[0] | pb[1] | pb[2]) != 0;
        pk = key[phx];
      }
      if (transparent) {
        if (is_head && prev_set) put_reset(&o);
        bb_c(&o, ' ');
        continue;
      }
      if (is_head) {
        uint64_t k = key[hx];
        if (depth == 3) {
          if (!prev_set || (pk >> 24) != (k >> 24)) put_sgr_rgb(&o, 38, t[0], t[1], t[2]);
          if (!prev_set || (pk & 0xFFFFFF) != (k & 0xFFFFFF)) put_sgr_rgb(&o, 48, b[0], b[1], b[2]);
        } else if (depth == 2) {
          if (!prev_set || (pk >> 8) != (k >> 8)) put_sgr_256(&o, 38, (int)(k >> 8));
          if (!prev_set || (pk & 0xFF) != (k & 0xFF)) put_sgr_256(&o, 48, (int)(k & 0xFF));
        } else {
          if (!prev_set || (pk >> 8) != (k >> 8)) put_sgr_16(&o, 0, (int)(k >> 8));
          if (!prev_set || (pk & 0xFF) != (k & 0xFF)) put_sgr_16(&o, 1, (int)(k & 0xFF));
        }
        bb_put(&o, HB, 3);
        if (rep) put_rep(&o, (uint32_t)rl[x] - 1);
      } else if (!rep) {
        bb_put(&o, HB, 3);
      }
    }
    if (depth != 0) put_reset(&o);
    if (y + 2 < h) bb_c(&o, '\n');
  }
  free(key);
  free(hd);
  free(rl);
  return bb_finish(&o, out_len);
}

/* image_print_with_capabilities — ascii.c:955-1002 (SIMD_SUPPORT build) */
char *orc_print(const uint8_t *rgb, int w, int h, int color_level, int render_mode, const char *palette,
                size_t *out_len) {
  if (!rgb || !palette) return NULL;
  if (render_mode == ORC_MODE_HALF) {
    int depth = color_level == ORC_COLOR_TRUE ? 3 : color_level == ORC_COLOR_256 ? 2 : color_level == ORC_COLOR_16 ? 1 : 0;
    return print_halfblock(rgb, w, h, depth, out_len);
  }
  if (w <= 0 || h <= 0) return NULL;
  switch (color_level) {
  case ORC_COLOR_TRUE:
    return render_mode == ORC_MODE_BG ? print_dither_bg(rgb, w, h, palette, out_len)
                                      : print_true_fg(rgb, w, h, palette, out_len);
  case ORC_COLOR_256:
    return print_256_fg(rgb, w, h, palette, out_len);
  case ORC_COLOR_16:
    return print_16_fg(rgb, w, h, palette, out_len);
  default:
    return print_mono_fg(rgb, w, h, palette, out_len);
  }
}

/* ------------------------------------------------------------------------- padding */
char *orc_pad_width(const char *frame, size_t pad_left) { /* ascii.c:457-517 */
  if (!frame) return NULL;
  size_t n = strlen(frame), lines = 1;
  for (size_t i = 0; i < n; i++) lines += frame[i] == '\n';
  char *out = malloc(n + (pad_left ? lines * pad_left : 0) + 1), *q = out;
  if (pad_left == 0) {
    memcpy(out, frame, n + 1);
    return out;
  }
  bool bol = true;
  for (size_t i = 0; i < n; i++) {
    if (bol) {
      memset(q, ' ', pad_left);
      q += pad_left;
      bol = false;
    }
    *q++ = frame[i];
    if (frame[i] == '\n') bol = true;
  }
  *q = '\0';
  return out;
}
char *orc_pad_height(const char *frame, size_t pad_top) { /* ascii.c:902-941 */
  if (!frame) return NULL;
  size_t n = strlen(frame);
  char *out = malloc(n + pad_top + 1);
  memset(out, '\n', pad_top);
  memcpy(out + pad_top, frame, n + 1);
  return out;
}

static char *finish_convert(const uint8_t *rgb, int w, int h, long rw, long rh, int color_level, int render_mode,
                            const char *palette, int scale, size_t pad_w, size_t pad_h, size_t *out_len) {
  if (rw <= 0 || rh <= 0 || rw > INT_MAX || rh > INT_MAX) return NULL;
  /* image_new limits (image.c:38-85, image.h:166-193): 3840x2160 max */
  if (rw > 3840 || rh > 2160) return NULL;
  uint8_t *small = malloc((size_t)rw * (size_t)rh * 3);
  if (scale == ORC_SCALE_BOX) orc_resize_box(rgb, w, h, small, (int)rw, (int)rh);
  else orc_resize_nn(rgb, w, h, small, (int)rw, (int)rh);
  size_t n = 0;
  char *s = orc_print(small, (int)rw, (int)rh, color_level, render_mode, palette, &n);
  free(small);
  if (!s) return NULL;
  if (n == 0) { /* ascii.c:355-361 */
    free(s);
    return NULL;
  }
  char *a = orc_pad_width(s, pad_w);
  free(s);
  char *b = orc_pad_height(a, pad_h);
  free(a);
  if (out_len) *out_len = strlen(b);
  return b;
}

/* ascii_convert_with_capabilities — ascii.c:194-387 */
char *orc_convert_caps(const uint8_t *rgb, int w, int h, long width, long height, int color_level, int render_mode,
                       int wants_padding, int use_aspect_ratio, int stretch, const char *palette, int scale,
                       size_t *out_len) {
  if (!rgb) return NULL;
  if (w <= 0 || w > 10000 || h <= 0 || h > 10000) return NULL;
  long rw = width, rh = height;
  if (use_aspect_ratio) orc_aspect_ratio(w, h, rw, rh, stretch, &rw, &rh);
  long ow = rw, oh = rh;
  if (render_mode == ORC_MODE_HALF) rh *= 2;
  size_t pw = 0, ph = 0;
  if (use_aspect_ratio && wants_padding) {
    pw = (size_t)(width > ow ? (width - ow) / 2 : 0);
    ph = (size_t)(height > oh ? (height - oh) / 2 : 0);
  }
  if (!palette) return NULL; /* image_print_with_capabilities rejects NULL palette (ascii.c:956) */
  return finish_convert(rgb, w, h, rw, rh, color_level, render_mode, palette, scale, pw, ph, out_len);
}

/* ascii_convert — ascii.c:72-191; opt_render_mode = GET_OPTION(render_mode) */
char *orc_convert(const uint8_t *rgb, int w, int h, long width, long height, int color, int aspect, int stretch,
                  const char *palette, int opt_render_mode, size_t *out_len) {
  if (!rgb || !palette || !palette[0]) return NULL;
  long rw = width, rh = height;
  if (aspect) orc_aspect_ratio(w, h, rw, rh, stretch, &rw, &rh);
  size_t pw = 0, ph = 0;
  if (aspect) {
    pw = (size_t)(width > rw ? (width - rw) / 2 : 0);
    ph = (size_t)(height > rh ? (height - rh) / 2 : 0);
  }
  int level = color ? ORC_COLOR_TRUE : ORC_COLOR_NONE;
  int mode = ORC_MODE_FG;
  if (color) mode = opt_render_mode == ORC_MODE_HALF ? ORC_MODE_HALF : opt_render_mode == ORC_MODE_BG ? ORC_MODE_BG : ORC_MODE_FG;
  return finish_convert(rgb, w, h, rw, rh, level, mode, palette, ORC_SCALE_NN, pw, ph, out_len);
}

/* ------------------------------------------------------------------ text-space grid */
static int visible_width(const char *d, int n) { /* ascii.c:527-551 */
  int v = 0, i = 0;
  while (i < n) {
    if (d[i] == '\033' && i + 1 < n && d[i + 1] == '[') {
      i += 2;
      while (i < n) {
        char c = d[i++];
        if (c >= '@' && c <= '~') break;
      }
    } else {
      v++;
      i++;
    }
  }
  return v;
}
static int truncate_visible(const char *d, int n, int target) { /* ascii.c:562-586 */
  int v = 0, i = 0;
  while (i < n && v < target) {
    if (d[i] == '\033' && i + 1 < n && d[i + 1] == '[') {
      i += 2;
      while (i < n) {
        char c = d[i++];
        if (c >= '@' && c <= '~') break;
      }
    } else {
      v++;
      i++;
    }
  }
  return i;
}

char *orc_create_grid(const char *const *frames, const size_t *sizes, int n, int width, int height,
                      size_t *out_size) { /* ascii.c:602-885 */
  if (!frames || n <= 0 || width <= 0 || height <= 0 || !out_size) return NULL;
  size_t W = (size_t)width, H = (size_t)height, total = W * H + H + 1;
  if (n == 1) {
    char *r = malloc(total);
    memset(r, ' ', total - 1);
    r[total - 1] = '\0';
    for (int row = 0; row < height; row++) r[(size_t)row * (W + 1) + W] = '\n';
    const char *s = frames[0];
    int sn = (int)sizes[0];
    *out_size = total - 1;
    if (!s || sn <= 0) return r;
    int lines = 0;
    for (int i = 0; i < sn; i++) lines += s[i] == '\n';
    int vpad = (height - lines) / 2;
    if (vpad < 0) vpad = 0;
    int row = vpad, pos = 0;
    while (pos < sn && row < height) {
      int ls = pos;
      while (pos < sn && s[pos] != '\n') pos++;
      int ll = pos - ls;
      int hp = (width - visible_width(s + ls, ll)) / 2;
      if (hp < 0) hp = 0;
      size_t dst = (size_t)row * (W + 1) + (size_t)hp;
      int cl = truncate_visible(s + ls, ll, width - hp);
      if (cl > 0 && dst + (size_t)cl < total) memcpy(r + dst, s + ls, (size_t)cl);
      if (pos < sn && s[pos] == '\n') pos++;
      row++;
    }
    return r;
  }
  float best = -1.0f;
  int bc = 1, br = n;
  for (int c = 1; c <= n; c++) {
    int rws = (int)ceil((double)n / c);
    if (c * rws - n > n / 2) continue;
    int cw = (width - (c - 1)) / c, ch = (height - (rws - 1)) / rws;
    if (cw < 10 || ch < 3) continue;
    float cell_aspect = ((float)cw / (float)ch) / 2.0f;
    float as = 1.0f - fabsf(logf(cell_aspect));
    if (as < 0) as = 0;
    float util = (float)n / (float)(c * rws);
    float score = n == 2 ? as * 0.9f + util * 0.1f : as * 0.7f + util * 0.3f;
    if (c == rws) score += 0.05f;
    if (score > best) {
      best = score;
      bc = c;
      br = rws;
    }
  }
  int cw = (width - (bc - 1)) / bc, ch = (height - (br - 1)) / br;
  if (cw < 10 || ch < 3) {
    char *r = malloc(sizes[0] + 1);
    if (frames[0] && sizes[0] > 0) {
      memcpy(r, frames[0], sizes[0]);
      r[sizes[0]] = '\0';
      *out_size = sizes[0];
    } else {
      r[0] = '\0';
      *out_size = 0;
    }
    return r;
  }
  char *m = malloc(total);
  memset(m, ' ', total - 1);
  m[total - 1] = '\0';
  for (int row = 0; row < height; row++) m[(size_t)row * (W + 1) + W] = '\n';
  for (int s = 0; s < n; s++) {
    int gr = s / bc, gc = s % bc, r0 = gr * (ch + 1), c0 = gc * (cw + 1);
    const char *d = frames[s];
    int sn = (int)sizes[s], pos = 0, srow = 0;
    while (pos < sn && srow < ch && r0 + srow < height) {
      int ls = pos;
      while (pos < sn && d[pos] != '\n') pos++;
      int ll = pos - ls;
      int cl = truncate_visible(d + ls, ll, cw);
      int tv = visible_width(d + ls, cl);
      if (cl > 0 && c0 + tv <= width) {
        /* SAFE_MEMCPY (ascii.c:844) = platform_memcpy (lib/platform/posix/system.c:653-666): silently refuses
         * when count > remaining size.  ANSI bytes make cl exceed the cell, so accepted copies spill over the
         * following cells/newlines exactly as in the reference. */
        size_t pos = (size_t)(r0 + srow) * (W + 1) + (size_t)c0;
        if ((size_t)cl <= total - pos) memcpy(m + pos, d + ls, (size_t)cl);
      }
      if (pos < sn && d[pos] == '\n') pos++;
      srow++;
    }
    if (gc < bc - 1 && c0 + cw < width)
      for (int row = r0; row < r0 + ch && row < height; row++) {
        size_t idx = (size_t)row * (W + 1) + (size_t)(c0 + cw);
        if (idx < total - 1) m[idx] = '|';
      }
    if (gr < br - 1 && r0 + ch < height) {
      for (int col = c0; col < c0 + cw && col < width; col++) {
        size_t idx = (size_t)(r0 + ch) * (W + 1) + (size_t)col;
        if (idx < total - 1) m[idx] = '_';
      }
      if (gc < bc - 1 && c0 + cw < width) {
        size_t idx = (size_t)(r0 + ch) * (W + 1) + (size_t)(c0 + cw);
        if (idx < total - 1) m[idx] = '+';
      }
    }
  }
  m[total - 1] = '\0'; /* a spill may land on the terminator; the reference's strlen would then run off the
                          buffer (UB) — we re-terminate, which only differs in that UB case */
  *out_size = strlen(m);
  return m;
}

/* -------------------------------------------------------- server pixel-space composite */
void orc_grid_layout(const int *ws, const int *hs, int n, int term_w, int term_h, int *cols, int *rows) {
  /* calculate_optimal_grid_layout — stream.c:523-651 */
  if (n == 0) {
    *cols = *rows = 0;
    return;
  }
  if (n == 1) {
    *cols = *rows = 1;
    return;
  }
  float avg = 0.0f;
  for (int i = 0; i < n; i++) avg += (float)ws[i] / (float)hs[i];
  avg /= n;
  int bc = 1, br = n;
  float best = 0.0f;
  for (int c = 1; c <= n; c++) {
    int r = (n + c - 1) / c;
    if (c * r - n > c) continue;
    int cw = term_w / c, ch = term_h / r;
    if (cw < 20 || ch < 10) continue;
    float used = 0.0f;
    int area = cw * ch;
    for (int i = 0; i < n; i++) {
      float cva = (float)cw / ((float)ch * 2.0f);
      int fw, fh;
      if (avg > cva) {
        fw = cw;
        fh = (int)((cw / avg) / 2.0f);
      } else {
        fh = ch;
        fw = (int)(ch * 2.0f * avg);
      }
      if (fw > cw) fw = cw;
      if (fh > ch) fh = ch;
      used += fw * fh;
    }
    float util = used / (float)(area * n);
    if (util > best) {
      best = util;
      bc = c;
      br = r;
    }
  }
  *cols = bc;
  *rows = br;
}

int orc_composite(const uint8_t *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                  uint8_t *out, int *cols_out, int *rows_out) { /* create_multi_source_composite — stream.c:664-779 */
  int gc, gr;
  orc_grid_layout(ws, hs, n, width, height, &gc, &gr);
  if (cols_out) *cols_out = gc;
  if (rows_out) *rows_out = gr;
  int CW = width, CH = height * 2;
  memset(out, 0, (size_t)CW * CH * 3);
  if (gc <= 0 || gr <= 0) return 0;
  for (int i = 0, v = 0; i < n && v < 9; i++, v++) {
    int row = v / gc, col = v % gc;
    int cw = CW / gc, ch = CH / gr;
    float sa = (float)ws[i] / (float)hs[i], ca = (float)cw / (float)ch;
    int tw, th;
    if (sa > ca) {
      tw = cw;
      th = (int)((cw / sa) + 0.5f);
    } else {
      th = ch;
      tw = (int)((ch * sa) + 0.5f);
    }
    if (tw <= 0 || th <= 0) continue; /* image_new_from_pool rejects 0-sized (image.c:129); reference would crash */
    uint8_t *rs = malloc((size_t)tw * th * 3);
    orc_resize_nn(srcs[i], ws[i], hs[i], rs, tw, th);
    int x0 = col * cw, y0 = row * ch, xp = (cw - tw) / 2, yp = (ch - th) / 2;
    for (int y = 0; y < th; y++)
      for (int x = 0; x < tw; x++) {
        int dx = x0 + xp + x, dy = y0 + yp + y;
        if (dx < x0 || dx > x0 + cw - 1 || dy < y0 || dy > y0 + ch - 1) continue;
        if (dx < 0 || dx >= CW || dy < 0 || dy >= CH) continue;
        memcpy(out + ((size_t)dy * CW + dx) * 3, rs + ((size_t)y * tw + x) * 3, 3);
      }
    free(rs);
  }
  return 0;
}

/* ------------------------------------------------------------------ synthetic inputs */
void orc_gen_pattern(int kind, uint32_t frame, uint8_t *dst, int w, int h) { /* SURVEY.md Appendix C */
  size_t npx = (size_t)w * h;
  switch (kind) {
  case 0: {
    uint32_t s = 12345u + frame;
    for (size_t i = 0; i < npx * 3; i++) {
      s = s * 1664525u + 1013904223u;
      dst[i] = (uint8_t)(s >> 24);
    }
    break;
  }
  case 1:
    for (size_t i = 0; i < npx; i++) {
      uint8_t v = (uint8_t)(((uint64_t)i * 255u) / npx + frame);
      dst[3 * i] = v;
      dst[3 * i + 1] = v / 2;
      dst[3 * i + 2] = (uint8_t)(255 - v);
    }
    break;
  case 2: {
    int bw = w / 8 ? w / 8 : 1, bh = h / 8 ? h / 8 : 1;
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        uint8_t *p = dst + ((size_t)y * w + x) * 3;
        int g = ((x + (int)frame) / bw) % 3;
        p[0] = g == 0 ? 255 : 0;
        p[1] = g == 1 ? 255 : 0;
        p[2] = g == 2 ? 255 : 0;
        if ((x + (int)frame) % bw == 0 || y % bh == 0) p[0] = p[1] = p[2] = 0;
      }
    break;
  }
  case 3:
    for (int y = 0; y < h; y++)
      for (int x = 0; x < w; x++) {
        uint8_t v = (uint8_t)(w > 1 ? ((uint64_t)x * 255u) / (uint64_t)(w - 1) : 0);
        v = (uint8_t)(v + frame);
        uint8_t *p = dst + ((size_t)y * w + x) * 3;
        p[0] = p[1] = p[2] = v;
      }
    break;
  default:
    memset(dst, (int)(frame & 255u), npx * 3);
  }
}

uint32_t orc_fnv1a32(const uint8_t *p, size_t n) {
  uint32_t h = 2166136261u;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= 16777619u;
  }
  return h;
}

/* exhaustive 2^24-entry quantiser tables (index = r<<16 | g<<8 | b).  which: 0 = 256-colour, 1 = 16-colour,
 * 2 = call fn (a uint8_t(*)(uint8_t,uint8_t,uint8_t), e.g. the compiled reference's rgb_to_256color). */
void orc_fill_table(int which, void *fn, uint8_t *out) {
  uint8_t (*f)(uint8_t, uint8_t, uint8_t) = (uint8_t(*)(uint8_t, uint8_t, uint8_t))fn;
  for (int r = 0; r < 256; r++)
    for (int g = 0; g < 256; g++)
      for (int b = 0; b < 256; b++) {
        size_t i = ((size_t)r << 16) | ((size_t)g << 8) | (size_t)b;
        out[i] = which == 0   ? (uint8_t)orc_rgb_to_256(r, g, b)
                 : which == 1 ? (uint8_t)orc_rgb_to_16(r, g, b)
                              : f((uint8_t)r, (uint8_t)g, (uint8_t)b);
      }
}

/* ------------------------------------------------------------------ CPU bench helper */
typedef struct { /* layout-compatible with the reference's image_t (image.h:143-148) */
  int w, h;
  void *pixels;
  uint8_t alloc_method;
} ref_image_t;
typedef char *(*ref_convert_fn)(ref_image_t *, long, long, const void *, bool, bool, const char *);

typedef struct {
  const uint8_t *src;
  int ring, w, h;
  long width, height;
  int color_level, render_mode, scale, frames, threads, tid;
  const char *palette;
  ref_convert_fn ref_fn;
  const void *ref_caps;
  uint64_t bytes;
} bench_arg_t;

static void *bench_worker(void *vp) {
  bench_arg_t *a = (bench_arg_t *)vp;
  size_t fsz = (size_t)a->w * a->h * 3;
  for (int i = a->tid; i < a->frames; i += a->threads) {
    const uint8_t *px = a->src + (size_t)(i % a->ring) * fsz;
    char *s;
    size_t n = 0;
    if (a->ref_fn) {
      ref_image_t img = {a->w, a->h, (void *)px, 0};
      s = a->ref_fn(&img, a->width, a->height, a->ref_caps, false, false, a->palette);
      n = s ? strlen(s) : 0;
    } else {
      s = orc_convert_caps(px, a->w, a->h, a->width, a->height, a->color_level, a->render_mode, 0, 0, 0, a->palette,
                           a->scale, &n);
    }
    a->bytes += n;
    free(s);
  }
  return NULL;
}

/* Box-mode CPU baseline (SURVEY.md §8d last bullet): CPU box filter (this repo's specification) + the printer.
 * print_fn = the compiled reference's image_print_with_capabilities(const image_t *, const caps *, const char *palette)
 * with its caps blob, or NULL for the port's printer.  Frame-parallel over `threads` pthreads like orc_bench_convert. */
typedef char *(*ref_print_fn)(const ref_image_t *, const void *, const char *);
typedef struct {
  const uint8_t *src;
  int ring, w, h, cols, rows_px, color_level, render_mode, frames, threads, tid;
  const char *palette;
  ref_print_fn print_fn;
  const void *caps;
  uint64_t bytes;
} box_arg_t;

static void *box_worker(void *vp) {
  box_arg_t *a = (box_arg_t *)vp;
  const size_t fsz = (size_t)a->w * a->h * 3;
  uint8_t *small = malloc((size_t)a->cols * a->rows_px * 3);
  for (int i = a->tid; i < a->frames; i += a->threads) {
    orc_resize_box_fast(a->src + (size_t)(i % a->ring) * fsz, a->w, a->h, small, a->cols, a->rows_px);
    char *s;
    size_t n = 0;
    if (a->print_fn) {
      ref_image_t img = {a->cols, a->rows_px, small, 0};
      s = a->print_fn(&img, a->caps, a->palette);
      n = s ? strlen(s) : 0;
    } else {
      s = orc_print(small, a->cols, a->rows_px, a->color_level, a->render_mode, a->palette, &n);
    }
    a->bytes += n;
    free(s);
  }
  free(small);
  return NULL;
}

double orc_bench_box(const uint8_t *src, int ring, int w, int h, int cols, int rows_px, int color_level, int render_mode,
                     const char *palette, int frames, int threads, void *print_fn, const void *caps,
                     uint64_t *out_bytes) {
  if (threads < 1) threads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
  box_arg_t *args = calloc((size_t)threads, sizeof(box_arg_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    args[t] = (box_arg_t){src, ring, w, h, cols, rows_px, color_level, render_mode, frames, threads, t,
                          palette, (ref_print_fn)print_fn, caps, 0};
    pthread_create(&th[t], NULL, box_worker, &args[t]);
  }
  uint64_t bytes = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    bytes += args[t].bytes;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (out_bytes) *out_bytes = bytes;
  free(th);
  free(args);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double orc_bench_convert(const uint8_t *src, int ring, int w, int h, long width, long height, int color_level,
                         int render_mode, const char *palette, int scale, int frames, int threads, void *ref_fn,
                         const void *ref_caps, uint64_t *out_bytes) {
  if (threads < 1) threads = 1;
  pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
  bench_arg_t *args = calloc((size_t)threads, sizeof(bench_arg_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    args[t] = (bench_arg_t){src, ring, w, h, width, height, color_level, render_mode, scale, frames, threads, t,
                            palette, (ref_convert_fn)ref_fn, ref_caps, 0};
    pthread_create(&th[t], NULL, bench_worker, &args[t]);
  }
  uint64_t bytes = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    bytes += args[t].bytes;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (out_bytes) *out_bytes = bytes;
  free(th);
  free(args);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ client display path
 * session_display_convert_to_ascii, src/common/session/display.c:484-671: flip -> colour filter -> convert ->
 * rainbow replace.  The filter is restated as what it is: a 256-entry table over the BT.601 grey value. */
static const struct {
  uint8_t r, g, b, fg_on_bg;
} k_filters[ORC_FILTER_COUNT] = { /* color_filter.c:23-142 */
    {0, 0, 0, 0},       {0, 0, 0, 1},     {255, 255, 255, 0}, {0, 255, 65, 0},   {255, 0, 255, 0},
    {255, 0, 170, 0},   {255, 136, 0, 0}, {0, 221, 221, 0},   {0, 255, 255, 0},  {255, 182, 193, 0},
    {255, 51, 51, 0},   {255, 235, 153, 0}, {255, 0, 0, 0}};

void orc_calculate_rainbow(float time, uint8_t *r, uint8_t *g, uint8_t *b) { /* color_filter.c:165-236 */
  const float period = 3.5f;
  float phase = fmodf(time, period) / period;
  float hue = phase * 360.0f;
  float h = hue / 60.0f;
  int i = (int)floorf(h);
  float f = h - (float)i;
  uint8_t up = (uint8_t)(f * 255.0f + 0.5f), down = (uint8_t)((1.0f - f) * 255.0f + 0.5f);
  switch (i % 6) {
  case 0: *r = 255, *g = up, *b = 0; break;
  case 1: *r = down, *g = 255, *b = 0; break;
  case 2: *r = 0, *g = 255, *b = up; break;
  case 3: *r = 0, *g = down, *b = 255; break;
  case 4: *r = up, *g = 0, *b = 255; break;
  case 5: *r = 255, *g = 0, *b = down; break;
  default: *r = 255, *g = 0, *b = 0; break;
  }
  float lum = 0.2126f * *r + 0.7152f * *g + 0.0722f * *b;
  if (lum < 120.0f) {
    float boost = (120.0f - lum) / 3.0f;
    *r = (uint8_t)fminf(255.0f, *r + boost);
    *g = (uint8_t)fminf(255.0f, *g + boost);
    *b = (uint8_t)fminf(255.0f, *b + boost);
  }
}

/* table[gray] = filtered (r,g,b); returns 0 when the filter is a no-op / invalid (color_filter.c:238-346) */
int orc_filter_table(int filter, float time, uint8_t table[256][3]) {
  if (filter <= ORC_FILTER_NONE || filter >= ORC_FILTER_COUNT) return 0;
  unsigned fr = k_filters[filter].r, fg = k_filters[filter].g, fb = k_filters[filter].b;
  int on_bg = k_filters[filter].fg_on_bg;
  if (filter == ORC_FILTER_RAINBOW) {
    uint8_t r, g, b;
    orc_calculate_rainbow(time, &r, &g, &b);
    fr = r, fg = g, fb = b;
  }
  for (unsigned gray = 0; gray < 256; gray++) {
    unsigned v = gray;
    if (filter == ORC_FILTER_RAINBOW) v = (179u + (uint8_t)((gray * (255u - 179u)) / 255u)) & 255u; /* :315 */
    if (on_bg) { /* :254-260 */
      table[gray][0] = (uint8_t)((fr * (255u - v) + 255u * v) / 255u);
      table[gray][1] = (uint8_t)((fg * (255u - v) + 255u * v) / 255u);
      table[gray][2] = (uint8_t)((fb * (255u - v) + 255u * v) / 255u);
    } else { /* :262-265 */
      table[gray][0] = (uint8_t)((fr * v) / 255u);
      table[gray][1] = (uint8_t)((fg * v) / 255u);
      table[gray][2] = (uint8_t)((fb * v) / 255u);
    }
  }
  return 1;
}

int orc_apply_color_filter(uint8_t *pixels, uint32_t width, uint32_t height, uint32_t stride, int filter, float time) {
  if (!pixels || width == 0 || height == 0 || stride == 0) return -1; /* :276 */
  if (filter == ORC_FILTER_NONE) return 0;
  uint8_t tab[256][3];
  if (!orc_filter_table(filter, time, tab)) return -1;
  for (uint32_t y = 0; y < height; y++) {
    uint8_t *p = pixels + (size_t)y * stride;
    for (uint32_t x = 0; x < width; x++, p += 3) {
      unsigned gray = (77u * p[0] + 150u * p[1] + 29u * p[2]) >> 8; /* color_filter.h:172 */
      p[0] = tab[gray][0], p[1] = tab[gray][1], p[2] = tab[gray][2];
    }
  }
  return 0;
}

/* rainbow_replace_ansi_colors, color_filter.c:348-408: every ESC[38;2;...m becomes the frame's rainbow SGR.
 * NULL when the string holds no truecolor-foreground SGR at all. */
char *orc_rainbow_replace(const char *s, float time_seconds) {
  static const char key[] = "\x1b[38;2;";
  if (!s || !strstr(s, key)) return NULL;
  uint8_t r, g, b;
  orc_calculate_rainbow(time_seconds, &r, &g, &b);
  char code[32];
  int cl = snprintf(code, sizeof(code), "\x1b[38;2;%d;%d;%dm", r, g, b);
  bb_t o = {0};
  size_t n = strlen(s), i = 0;
  while (i < n) {
    if (n - i >= 7 && memcmp(s + i, key, 7) == 0) {
      const char *m = memchr(s + i + 7, 'm', n - i - 7);
      if (m) {
        bb_put(&o, code, (size_t)cl);
        i = (size_t)(m - s) + 1;
        continue;
      }
    }
    bb_put(&o, s + i, 1);
    i++;
  }
  bb_need(&o, 1);
  o.p[o.n] = '\0';
  return o.p;
}

char *orc_display_convert(const uint8_t *rgb, int w, int h, long width, long height, int color_level, int render_mode,
                          int wants_padding, int preserve_aspect, int stretch, const char *palette, int flip_x,
                          int flip_y, int filter, float time_seconds, int scale, size_t *out_len) {
  if (!rgb || w <= 0 || h <= 0) return NULL;
  size_t px = (size_t)w * (size_t)h;
  uint8_t *img = (uint8_t *)malloc(px * 3);
  const int fx = flip_x && w > 1 && h > 1, fy = flip_y && w > 1 && h > 1; /* display.c:548 */
  for (int y = 0; y < h; y++) {
    const uint8_t *srow = rgb + (size_t)(fy ? h - 1 - y : y) * (size_t)w * 3;
    uint8_t *drow = img + (size_t)y * (size_t)w * 3;
    for (int x = 0; x < w; x++) memcpy(drow + 3 * (size_t)x, srow + 3 * (size_t)(fx ? w - 1 - x : x), 3);
  }
  if (filter != ORC_FILTER_NONE && filter != ORC_FILTER_RAINBOW)
    orc_apply_color_filter(img, (uint32_t)w, (uint32_t)h, (uint32_t)w * 3, filter, time_seconds);
  char *res = orc_convert_caps(img, w, h, width, height, color_level, render_mode, wants_padding, preserve_aspect,
                               stretch, palette, scale, out_len);
  free(img);
  if (res && filter == ORC_FILTER_RAINBOW) {
    char *r2 = orc_rainbow_replace(res, time_seconds);
    if (r2) {
      free(res);
      res = r2;
    }
  }
  if (res && out_len) *out_len = strlen(res);
  return res;
}

/* ------------------------------------------------------------------ wire packaging
 * CRC32-C (Castagnoli, reflected 0x82F63B78; lib/network/crc32.c:168-189 and the SSE4.2/ARMv8 instructions of
 * :95-121) and the 24-byte big-endian ascii_frame_packet_t acip_send_ascii_frame() builds
 * (lib/network/acip/server.c:203-214, packet.h:848-862). */
uint32_t orc_crc32c(const uint8_t *p, size_t n) {
  static uint32_t tab[256];
  static int init = 0;
  if (!init) {
    for (uint32_t i = 0; i < 256; i++) {
      uint32_t c = i;
      for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ 0x82F63B78u : c >> 1;
      tab[i] = c;
    }
    init = 1;
  }
  uint32_t crc = 0xFFFFFFFFu;
  for (size_t i = 0; i < n; i++) crc = tab[(crc ^ p[i]) & 255u] ^ (crc >> 8);
  return ~crc;
}

static void put_be32(uint8_t *d, uint32_t v) { d[0] = (uint8_t)(v >> 24), d[1] = (uint8_t)(v >> 16), d[2] = (uint8_t)(v >> 8), d[3] = (uint8_t)v; }
void orc_frame_packet_header(const uint8_t *frame, size_t frame_size, uint32_t width, uint32_t height, uint8_t out24[24]) {
  put_be32(out24 + 0, width);
  put_be32(out24 + 4, height);
  put_be32(out24 + 8, (uint32_t)frame_size);
  put_be32(out24 + 12, 0);
  put_be32(out24 + 16, orc_crc32c(frame, frame_size));
  put_be32(out24 + 20, 0);
}

/* ------------------------------------------------------------------ digital rain (second half)
 * lib/video/anim/digital_rain.c restated: a Matrix-style brightness field laid over a finished frame string.
 *   brightness(col,row,t) = 1 - fract(wobble((offset[col] + t * fall_speed * speed[col] - row) / raindrop_length)),
 *   wobble(x) = x + 0.3 sin(sqrt2 x) + 0.2 sin(sqrt5 x)                                  (:35-46, 70-91), 0 beyond the last column
 * Walking the string with a (col,row) cursor (:405-502): a truecolor SGR (fg or bg) is re-emitted with its components
 * scaled by the cell's brightness; every other escape sequence is copied; every visible character is preceded by the rain
 * colour scaled the same way; '\n' resets the column.  Each such VISIT of a cell inside the grid also low-pass filters the
 * brightness against the value stored for the cell (:421-430), so a cell visited three times in one frame (fg SGR, bg SGR,
 * glyph) is filtered three times — the order of visits is part of the result.  All float, single precision, in the
 * reference's order of operations; sinf/fmodf/floorf are libm's. */
static float rain_hash(float x, float y) { /* :31-35 */
  float dt = x * 12.9898f + y * 78.233f;
  float sn = fmodf(dt, (float)M_PI);
  return fmodf(sinf(sn) * 43758.5453f, 1.0f);
}
orc_rain_t *orc_rain_init(int cols, int rows) { /* :97-153 */
  if (cols <= 0 || rows <= 0) return NULL;
  orc_rain_t *r = calloc(1, sizeof(*r));
  r->cols = cols, r->rows = rows;
  r->offset = malloc(sizeof(float) * (size_t)cols);
  r->speed = malloc(sizeof(float) * (size_t)cols);
  r->prev = calloc((size_t)cols * rows, sizeof(float));
  for (int c = 0; c < cols; c++) {
    r->offset[c] = rain_hash((float)c, 0.0f) * 1000.0f;
    r->speed[c] = rain_hash((float)c + 0.1f, 0.0f) * 0.5f + 0.5f;
  }
  r->fall_speed = 3.0f, r->drop_len = 12.0f, r->decay = 0.1f, r->anim_speed = 1.0f;
  r->cr = 0, r->cg = 255, r->cb = 80;
  r->first = 1;
  return r;
}
void orc_rain_destroy(orc_rain_t *r) {
  if (!r) return;
  free(r->offset);
  free(r->speed);
  free(r->prev);
  free(r);
}
void orc_rain_set_filter(orc_rain_t *r, int filter) { /* :205-230 */
  if (!r) return;
  r->rainbow = 0;
  if (filter == ORC_FILTER_NONE) r->cr = 0, r->cg = 255, r->cb = 80;
  else if (filter == ORC_FILTER_RAINBOW) r->rainbow = 1, r->cr = 255, r->cg = 0, r->cb = 0;
  else if (filter > 0 && filter < ORC_FILTER_COUNT) r->cr = k_filters[filter].r, r->cg = k_filters[filter].g, r->cb = k_filters[filter].b;
}
float orc_rain_target(const orc_rain_t *r, int col, int row, float t) { /* get_rain_brightness, :70-91 */
  if (col < 0 || col >= r->cols) return 0.0f;
  float column_time = r->offset[col] + t * r->fall_speed * r->speed[col];
  float x = (column_time - (float)row) / r->drop_len;
  x = x + 0.3f * sinf((float)1.4142135623730951 * x) + 0.2f * sinf((float)2.23606797749979 * x);
  return 1.0f - (x - floorf(x));
}
/* one visit of cell (col,row): filtered brightness, whether the cell is a drop's cursor (:413-430) */
static float rain_visit(orc_rain_t *r, int col, int row, float t, int *cursor) {
  float b = orc_rain_target(r, col, row, t);
  *cursor = b > orc_rain_target(r, col, row + 1, t);
  if (row < r->rows && col < r->cols) {
    float *p = &r->prev[(size_t)row * r->cols + col];
    if (!r->first) b = *p + (b - *p) * r->decay;
    *p = b;
  }
  return b;
}
static void rain_code(bb_t *o, int layer, int red, int green, int blue, float b, int cursor) { /* :326-364 */
  if (cursor) b *= 2.0f;
  if (b < 0.0f) b = 0.0f;
  if (b > 1.0f) b = 1.0f;
  int v[3] = {(int)((float)red * b), (int)((float)green * b), (int)((float)blue * b)};
  for (int k = 0; k < 3; k++) v[k] = v[k] < 0 ? 0 : v[k] > 255 ? 255 : v[k];
  put_sgr_rgb(o, layer, v[0], v[1], v[2]);
}
/* ESC [ (38|48) ;2; R ; G ; B m — returns the byte after it, or NULL (:240-301) */
static const char *rain_parse_sgr(const char *s, int *rgb, int *layer) {
  if (s[0] != 0x1b || s[1] != '[') return NULL;
  if (s[2] == '3' && s[3] == '8') *layer = 38;
  else if (s[2] == '4' && s[3] == '8') *layer = 48;
  else return NULL;
  if (s[4] != ';' || s[5] != '2' || s[6] != ';') return NULL;
  const char *p = s + 7;
  for (int k = 0; k < 3; k++) {
    rgb[k] = 0;
    while (*p >= '0' && *p <= '9') rgb[k] = rgb[k] * 10 + (*p++ - '0');
    if (*p != (k < 2 ? ';' : 'm')) return NULL;
    p++;
  }
  return p;
}
static int rain_utf8_len(const unsigned char *s) { /* utf8_decode's length rule, lib/util/utf8.c:18-44; invalid = 1 */
  int n = s[0] < 0x80 ? 1 : (s[0] & 0xE0) == 0xC0 ? 2 : (s[0] & 0xF0) == 0xE0 ? 3 : (s[0] & 0xF8) == 0xF0 ? 4 : 0;
  for (int k = 1; k < n; k++)
    if ((s[k] & 0xC0) != 0x80) return 1;
  return n ? n : 1;
}
char *orc_rain_apply(orc_rain_t *r, const char *frame, float dt, size_t *out_len) { /* :366-520 */
  if (!r || !frame) return NULL;
  r->time += dt * r->anim_speed;
  const float t = r->time;
  if (r->rainbow) orc_calculate_rainbow(t, &r->cr, &r->cg, &r->cb);
  bb_t o = {0};
  int col = 0, row = 0;
  for (const char *s = frame; *s;) {
    int rgb[3], layer, cursor;
    const char *after;
    if (*s == 0x1b) {
      if ((after = rain_parse_sgr(s, rgb, &layer)) != NULL) {
        float b = rain_visit(r, col, row, t, &cursor);
        rain_code(&o, layer, rgb[0], rgb[1], rgb[2], b, cursor);
        s = after;
      } else { /* any other sequence is copied: ESC alone, or ESC [ ... up to and including a byte in @..~ (:306-324) */
        const char *e = s + 1;
        if (*e == '[') {
          e++;
          while (*e && !(*e >= '@' && *e <= '~')) e++;
          if (*e) e++;
        }
        bb_put(&o, s, (size_t)(e - s));
        s = e;
      }
    } else if (*s == '\n') {
      bb_c(&o, '\n');
      s++, row++, col = 0;
    } else {
      float b = rain_visit(r, col, row, t, &cursor);
      rain_code(&o, 38, r->cr, r->cg, r->cb, b, cursor);
      int n = rain_utf8_len((const unsigned char *)s);
      bb_put(&o, s, (size_t)n);
      s += n, col++;
    }
  }
  r->first = 0;
  return bb_finish(&o, out_len);
}

