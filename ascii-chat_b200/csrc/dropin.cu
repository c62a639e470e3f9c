// dropin.cu — libasciichat's own entry points (include/asciichat_b200.h Part 1), implemented
// on the B200 engine.  Validation order, NULL/errno behaviour and ownership follow the reference
// functions cited at each definition; the rendering itself is render_kernels.cu.
#include <climits>
#include <cstring>

#include "engine.h"

using namespace acb;

// ------------------------------------------------------------------ host float aspect fit
// lib/util/aspect_ratio.c:17-93.  Deliberately host-side float, expression for expression, so the
// rounding is the host compiler's (the only float arithmetic on the whole path).
static ssize_t fit_width_from_height(ssize_t height, ssize_t img_w, ssize_t img_h) {
  if (img_h == 0) return 1;
  float width = (float)height * (float)img_w / (float)img_h * 2.0f;
  int result = (int)(0.5f + width);
  return result > 0 ? result : 1;
}
static ssize_t fit_height_from_width(ssize_t width, ssize_t img_w, ssize_t img_h) {
  if (img_w == 0) return 1;
  float height = ((float)width / 2.0f) * (float)img_h / (float)img_w;
  int result = (int)(0.5f + height);
  return result > 0 ? result : 1;
}

extern "C" void acb200_aspect_ratio(ssize_t img_w, ssize_t img_h, ssize_t width, ssize_t height, bool stretch,
                                    ssize_t *out_w, ssize_t *out_h) {
  if (!out_w || !out_h) return;
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
  ssize_t wfh = fit_width_from_height(height, img_w, img_h);
  ssize_t hfw = fit_height_from_width(width, img_w, img_h);
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

static char *convert_common(const uint8_t *rgb, int w, int h, ssize_t rw, ssize_t rh, int level, int mode,
                            size_t pad_w, size_t pad_h, const char *palette) {
  if (rw <= 0 || rh <= 0) { // ascii.c:114-117, 256-259
    set_error(E_INVALID_PARAM, "Invalid dimensions for resize: width=%zd, height=%zd", rw, rh);
    return nullptr;
  }
  if (rw > INT_MAX || rh > INT_MAX) {
    set_error(E_INVALID_PARAM, "Dimensions exceed INT_MAX");
    return nullptr;
  }
  acb200_render_cfg_t cfg{};
  cfg.src_w = w;
  cfg.src_h = h;
  cfg.cols = (int)rw;
  cfg.rows_px = (int)rh;
  cfg.color_level = level;
  cfg.render_mode = mode;
  cfg.scale = default_scale();
  cfg.pad_left = (int)pad_w;
  cfg.pad_top = (int)pad_h;
  cfg.palette = palette;
  return render_one_host(cfg, rgb, nullptr);
}

namespace acb {
bool plan_convert_with_caps(int w, int h, ssize_t width, ssize_t height, const terminal_capabilities_t *caps,
                            bool use_aspect_ratio, bool stretch, const char *palette, acb200_render_cfg_t *cfg) {
  if (w <= 0 || w > 10000 || h <= 0 || h > 10000) { // ascii.c:204
    set_error(E_INVALID_PARAM, "Invalid original image dimensions: w=%d, h=%d", w, h);
    return false;
  }
  ssize_t rw = width, rh = height;
  if (use_aspect_ratio) acb200_aspect_ratio(w, h, rw, rh, stretch, &rw, &rh); // :220
  const ssize_t out_w = rw, out_h = rh;
  if (caps->render_mode == RENDER_MODE_HALF_BLOCK) rh = rh * 2; // :230
  size_t pad_w = 0, pad_h = 0;
  if (use_aspect_ratio && caps->wants_padding) { // :238-243
    pad_w = (size_t)(width > out_w ? (width - out_w) / 2 : 0);
    pad_h = (size_t)(height > out_h ? (height - out_h) / 2 : 0);
  }
  if (rw <= 0 || rh <= 0) { // :256-259
    set_error(E_INVALID_PARAM, "Invalid dimensions for resize: width=%zd, height=%zd", rw, rh);
    return false;
  }
  if (rw > INT_MAX || rh > INT_MAX) {
    set_error(E_INVALID_PARAM, "Dimensions exceed INT_MAX");
    return false;
  }
  if (!palette) { // image_print_with_capabilities rejects it, ascii.c:956
    set_error(E_INVALID_PARAM, "palette is NULL");
    return false;
  }
  *cfg = acb200_render_cfg_t{};
  cfg->src_w = w;
  cfg->src_h = h;
  cfg->cols = (int)rw;
  cfg->rows_px = (int)rh;
  cfg->color_level = (int)caps->color_level;
  cfg->render_mode = (int)caps->render_mode;
  cfg->scale = default_scale();
  cfg->pad_left = (int)pad_w;
  cfg->pad_top = (int)pad_h;
  cfg->palette = palette;
  return true;
}
} // namespace acb

extern "C" {

// lib/video/ascii/ascii.c:194-387
char *ascii_convert_with_capabilities(image_t *original, const ssize_t width, const ssize_t height,
                                      const terminal_capabilities_t *caps, const bool use_aspect_ratio,
                                      const bool stretch, const char *palette_chars) {
  if (original == nullptr || caps == nullptr) { // :198
    set_error(E_INVALID_PARAM, "Invalid parameters for ascii_convert_with_capabilities");
    return nullptr;
  }
  if (original->w > 0 && original->w <= 10000 && original->h > 0 && original->h <= 10000 &&
      original->pixels == nullptr) { // :209 (after the dimension check of :204)
    set_error(E_INVALID_PARAM, "Original image pixels pointer is NULL");
    return nullptr;
  }
  acb200_render_cfg_t cfg;
  if (!plan_convert_with_caps(original->w, original->h, width, height, caps, use_aspect_ratio, stretch, palette_chars,
                              &cfg))
    return nullptr;
  return render_one_host(cfg, reinterpret_cast<const uint8_t *>(original->pixels), nullptr);
}

// lib/video/ascii/ascii.c:72-191
char *ascii_convert(image_t *original, const ssize_t width, const ssize_t height, const bool color,
                    const bool aspect_ratio, const bool stretch, const char *palette_chars,
                    const char luminance_palette[256]) {
  if (original == nullptr || !palette_chars || !luminance_palette) { // :75
    set_error(E_INVALID_PARAM, "ascii_convert: invalid parameters");
    return nullptr;
  }
  if (palette_chars[0] == '\0' || luminance_palette[0] == '\0') { // :81
    set_error(E_INVALID_PARAM, "ascii_convert: empty palette strings");
    return nullptr;
  }
  if (!original->pixels || original->w <= 0 || original->h <= 0) {
    set_error(E_INVALID_PARAM, "ascii_convert: invalid image");
    return nullptr;
  }
  ssize_t rw = width, rh = height;
  if (aspect_ratio) acb200_aspect_ratio(original->w, original->h, rw, rh, stretch, &rw, &rh); // :94-98
  size_t pad_w = 0, pad_h = 0;
  if (aspect_ratio) { // :104-111
    pad_w = (size_t)(width > rw ? (width - rw) / 2 : 0);
    pad_h = (size_t)(height > rh ? (height - rh) / 2 : 0);
  }
  int level = TERM_COLOR_NONE, mode = RENDER_MODE_FOREGROUND;
  if (color) { // :136-161: half-block -> truecolor half blocks; else image_print_color_simd(bg?)
    level = TERM_COLOR_TRUECOLOR;
    const int opt = option_render_mode();
    mode = opt == RENDER_MODE_HALF_BLOCK ? RENDER_MODE_HALF_BLOCK
           : opt == RENDER_MODE_BACKGROUND ? RENDER_MODE_BACKGROUND
                                           : RENDER_MODE_FOREGROUND;
  }
  return convert_common(reinterpret_cast<const uint8_t *>(original->pixels), original->w, original->h, rw, rh, level,
                        mode, pad_w, pad_h, palette_chars);
}

static char *print_image(const uint8_t *rgb, int w, int h, int level, int mode, const char *palette, int leaf_mode = -1) {
  acb200_render_cfg_t cfg{};
  cfg.src_w = w;
  cfg.src_h = h;
  cfg.cols = w;
  cfg.rows_px = h;
  cfg.color_level = level;
  cfg.render_mode = mode;
  cfg.scale = ACB200_SCALE_NN; // 1:1, the identity for both scalers
  cfg.palette = palette;
  return render_one_host(cfg, rgb, nullptr, leaf_mode);
}

static char *dup_empty() {
  char *s = (char *)user_alloc(1);
  if (s) s[0] = '\0';
  return s;
}

// lib/video/ascii/ascii.c:955-1002
char *image_print_with_capabilities(const image_t *image, const terminal_capabilities_t *caps, const char *palette) {
  if (!image || !caps || !palette) return nullptr; // :956
  if (caps->render_mode == RENDER_MODE_HALF_BLOCK && (image->w <= 0 || image->h <= 0))
    return dup_empty(); // halfblock.c:50-51 platform_strdup("")
  if (!image->pixels || image->w <= 0 || image->h <= 0) {
    set_error(E_INVALID_PARAM, "image_print: invalid image");
    return nullptr;
  }
  return print_image(reinterpret_cast<const uint8_t *>(image->pixels), image->w, image->h, (int)caps->color_level,
                     (int)caps->render_mode, palette);
}

// leaf printers — each is one fixed (colour depth, mode) of the same kernels
static char *leaf(const image_t *p, const char *palette, int level, int mode, int leaf_mode = -1) {
  if (!p || !palette || !p->pixels || p->w <= 0 || p->h <= 0) {
    set_error(E_INVALID_PARAM, "image or palette invalid");
    return nullptr;
  }
  return print_image(reinterpret_cast<const uint8_t *>(p->pixels), p->w, p->h, level, mode, palette, leaf_mode);
}
char *image_print(const image_t *p, const char *palette) { return leaf(p, palette, TERM_COLOR_NONE, 0); }
char *image_print_color(const image_t *p, const char *palette) { return leaf(p, palette, TERM_COLOR_TRUECOLOR, 0); }
char *image_print_256color(const image_t *p, const char *palette) { return leaf(p, palette, TERM_COLOR_256, 0); }
char *image_print_16color(const image_t *p, const char *palette) { return leaf(p, palette, TERM_COLOR_16, 0); }
// foreground.c:752-846.  use_background = false has no caller on the capability path (sgr.c:429-435 only asks for the
// background form) but it is the same recurrence with one SGR per cell instead of two.
char *image_print_16color_dithered_with_background(const image_t *image, bool use_background, const char *palette) {
  return leaf(image, palette, TERM_COLOR_TRUECOLOR, RENDER_MODE_BACKGROUND, use_background ? (int)EM_DITHER_BG : (int)EM_DITHER_FG);
}
// foreground.c:650-749: foreground-only, glyph through char_index_ramp (the Q2 mapping of image_print_16color)
char *image_print_16color_dithered(const image_t *image, const char *palette) {
  return leaf(image, palette, TERM_COLOR_TRUECOLOR, RENDER_MODE_BACKGROUND, (int)EM_DITHER_FG_RAMP);
}

static char *leaf_hb(const uint8_t *rgb, int width, int height, int stride_bytes, int level) {
  if (width <= 0 || height <= 0) return dup_empty(); // halfblock.c:50-51
  if (!rgb) {
    set_error(E_INVALID_PARAM, "rgb is NULL");
    return nullptr;
  }
  if (stride_bytes > 0 && stride_bytes != width * 3) {
    set_error(E_INVALID_PARAM, "only packed rows (stride = 3*width) are supported, got %d", stride_bytes);
    return nullptr;
  }
  return print_image(rgb, width, height, level, RENDER_MODE_HALF_BLOCK, " ");
}
char *rgb_to_truecolor_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes) {
  return leaf_hb(rgb, width, height, stride_bytes, TERM_COLOR_TRUECOLOR);
}
char *rgb_to_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes, const char *palette) {
  (void)palette;
  return leaf_hb(rgb, width, height, stride_bytes, TERM_COLOR_NONE);
}
char *rgb_to_16color_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes,
                                       const char *palette) {
  (void)palette;
  return leaf_hb(rgb, width, height, stride_bytes, TERM_COLOR_16);
}
char *rgb_to_256color_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes,
                                        const char *palette) {
  (void)palette;
  return leaf_hb(rgb, width, height, stride_bytes, TERM_COLOR_256);
}

// lib/video/rgba/image.c:256-328 — nearest-neighbour resize, result written into dest->pixels
void image_resize(const image_t *source, image_t *dest) {
  if (!source || !dest) { // :257
    set_error(E_INVALID_PARAM, "image_resize: s or d is NULL");
    return;
  }
  if (!source->pixels || !dest->pixels) { // :268
    set_error(E_INVALID_PARAM, "Invalid parameters to image_resize_interpolation");
    return;
  }
  const int sw = source->w, sh = source->h, dw = dest->w, dh = dest->h;
  if (sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) { // :279
    set_error(E_INVALID_PARAM, "Invalid image dimensions for resize: src=%dx%d dst=%dx%d", sw, sh, dw, dh);
    return;
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return;
  const size_t R = (size_t)sw * 3;
  const bool gather = dh < sh;
  const int rows = gather ? dh : sh;
  const size_t in_bytes = R * rows, out_bytes = (size_t)dw * dh * 3;
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, in_bytes) || !grow_device(&cx->d_in, &cx->d_in_cap, in_bytes) ||
      !grow_device(&cx->d_out, &cx->d_out_cap, out_bytes) || !grow_pinned(&cx->h_out, &cx->h_out_cap, out_bytes))
    return;
  const uint8_t *src = reinterpret_cast<const uint8_t *>(source->pixels);
  if (gather) {
    const uint32_t yr = (uint32_t)((((uint64_t)sh << 16) / (uint64_t)dh) + 1);
    for (int y = 0; y < dh; y++) {
      uint32_t sy = ((uint32_t)y * yr) >> 16;
      if (sy >= (uint32_t)sh) sy = (uint32_t)sh - 1;
      memcpy(cx->h_in + (size_t)y * R, src + (size_t)sy * R, R);
    }
  } else {
    memcpy(cx->h_in, src, in_bytes);
  }
  if (cudaMemcpyAsync(cx->d_in, cx->h_in, in_bytes, cudaMemcpyHostToDevice, cx->stream) != cudaSuccess ||
      launch_resize_nn_only(cx->d_in, sw, sh, cx->d_out, dw, dh, gather ? 1 : 0, cx->stream) != cudaSuccess ||
      cudaMemcpyAsync(cx->h_out, cx->d_out, out_bytes, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      wait_stream(cx) != E_OK) {
    set_error(E_INVALID_STATE, "image_resize: CUDA failure (%s)", cudaGetErrorString(cudaGetLastError()));
    return;
  }
  count_launch();
  memcpy(dest->pixels, cx->h_out, out_bytes);
}

// lib/video/ascii/common.c:601-604 / 497-538
void ascii_simd_init(void) { (void)thread_ctx(); }
void simd_caches_destroy_all(void) { destroy_lut_cache(); }

} // extern "C"
