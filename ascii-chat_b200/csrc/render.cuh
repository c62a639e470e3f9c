// render.cuh — shared declarations between the sm_100a kernels (render_kernels.cu)
// and the host engine (engine.cu).  Not part of the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace acb {

// Emission grammars (SURVEY.md §8a "exact output grammar"); one per reference renderer.
enum EmitMode : int {
  EM_MONO_FG = 0,  // image_print                       foreground.c:27-138
  EM_256_FG = 1,   // image_print_256color              foreground.c:433-509
  EM_16_FG = 2,    // image_print_16color               foreground.c:535-624
  EM_TRUE_FG = 3,  // image_print_color + ansi_rle_*    foreground.c:195-308, ansi.c:248-314
  EM_HB_TRUE = 4,  // rgb_to_truecolor_halfblocks_scalar halfblock.c:48-165
  EM_HB_256 = 5,   // rgb_to_256color_halfblocks_scalar  halfblock.c:416-524
  EM_HB_16 = 6,    // rgb_to_16color_halfblocks_scalar   halfblock.c:297-405
  EM_HB_MONO = 7,  // rgb_to_halfblocks_scalar           halfblock.c:184-286
  EM_DITHER_BG = 8, // image_print_16color_dithered_with_background(img, true)  foreground.c:752-846 (own kernel)
  // leaf printers without a capability route (sgr.c:429-435 only ever asks for the background form):
  EM_DITHER_FG = 9,      // image_print_16color_dithered_with_background(img, false): fg SGR only, glyph = cache[Y]
  EM_DITHER_FG_RAMP = 10 // image_print_16color_dithered  foreground.c:650-749: fg SGR only, glyph per quirk Q2
};
__host__ __device__ inline bool is_dither_mode(int m) { return m >= EM_DITHER_BG && m <= EM_DITHER_FG_RAMP; }

enum ScalePath : int {
  SP_NN = 0,        // nearest neighbour, image.c:267-328
  SP_BOX_GENERIC = 1, // box filter, any geometry (byte loads)
  SP_BOX_STREAM = 2,  // box filter, 16-byte streaming loads (3*src_w % 16 == 0, band <= 256 rows)
  SP_BOX_TMA = 3,     // box filter, persistent warp-specialised kernel: bulk-TMA row ring + consumer warps
  SP_BOX_SPLIT = 4    // box filter, persistent role-split kernel: 8 streamer warps + 1 emitter warp
};

// Brightness -> glyph tables for one (palette, mode) pair, built on the host
// (common.c:380-490 mapping incl. quirks Q1/Q2) and kept in device memory.
struct GlyphLut {
  uint8_t glyph[256][8]; // [Y] = {len, b0, b1, b2, b3, 0, 0, 0}
  uint8_t key[256];      // [Y] = char_index_ramp[Y >> 2]  (mono run key)
};

// Per text row bookkeeping produced by the row kernel, consumed by the stitch kernel.
struct RowMeta {
  uint32_t len;       // bytes written to the scratch row
  uint32_t cond_off;  // EM_TRUE_FG: offset of the SGR owned by the row's first ASCII-glyph cell
  uint32_t cond_len;  //             its length (0 = row has no ASCII-glyph cell)
  uint32_t first_rgb; //             colour of that cell   (0x01RRGGBB, 0 = none)
  uint32_t last_rgb;  //             colour of the row's last ASCII-glyph cell (0x01RRGGBB, 0 = none)
  uint32_t _pad[3];
};

struct RenderParams {
  const uint8_t *frames; // frame f at frames + f*frame_stride, packed RGB24, pitch 3*src_w
  size_t frame_stride;
  int src_w, src_h;
  int pregathered;       // NN only: source holds exactly rows_px rows, row y = the row NN would sample
  int cols, rows_px, text_rows;
  int pad_left;
  int use_smem_out;      // row fits the shared staging buffer
  uint32_t row_pitch;    // bytes between scratch rows (multiple of 16)
  uint8_t *rows;         // scratch: (f*text_rows + t) * row_pitch
  RowMeta *meta;         // (f*text_rows + t)
  const GlyphLut *lut;
  uint8_t *cells_out;    // optional: resized RGB24 image, frame f at f*cols*rows_px*3 (image_resize, tests)
  int n_frames;
  // direct output (role-split kernel): final arena + per-row look-back records (16 B each, + 16 B ticket)
  uint8_t *out;
  size_t out_pitch;
  uint32_t *out_len;
  uint4 *agg;
  int *ticket;            // atomic tile ticket: tile = fetched value - ticket_base
  uint32_t ticket_base;   // tickets handed out by earlier launches on this scratch (0 after a clear)
  uint32_t epoch;         // value a look-back record's ready word takes in this launch (never 0)
  int pad_top;
  int direct;             // 1: emitters place rows in the final arena (look-back); 0: scratch rows + k_stitch
  int ring_depth;         // warp-specialised kernel: source rows kept in flight by the producer warp
  // client display path (src/common/session/display.c:484-671), fused into the sampling and the emission:
  int flip_x, flip_y;     // the source is read mirrored (display.c:548-591); flip_y is already applied when pregathered
  int filt_mode;          // apply_color_filter (color_filter.c:238-346): FM_* below, on every pixel that is read
  uint32_t filt_rgb;      // the filter's colour, 0x00RRGGBB
  uint32_t fg_over;       // rainbow_replace_ansi_colors (color_filter.c:348-408): 0, or 0x01RRGGBB printed in place of
                          // every truecolor-foreground SGR colour (run/dedupe decisions still use the pixel colours)
  // streaming box filter, uniform even box width (src_w = cols * box_bx, box_bx even; 0 = use the generic sums):
  // bands are box_nrow0 or box_nrow0 + 1 rows tall; box_M[d] = ceil(2^32 / (box_bx * (box_nrow0 + d)))
  int box_bx, box_nrow0;
  uint32_t box_M[2];
  uint32_t nn_xr, nn_yr;  // nearest neighbour: ((src_w << 16) / cols) + 1, ((src_h << 16) / rows_px) + 1  (image.c:293-294)
  unsigned long long *dbg; // measurement counters (tune_flags bit 2): streamer wait, emitter wait, emitter busy, tiles
  int tune_flags;         // measurement knobs: bit0 no V/staging alias, bit1 no emission, bit2 counters, bit3 generic horizontal sums
};

// colour filter arithmetic variants (colorize_grayscale_pixel, color_filter.c:238-267)
enum FilterMode : int {
  FM_NONE = 0,
  FM_SCALE = 1,   // white-on-colour: c * gray / 255
  FM_ON_WHITE = 2, // colour-on-white ("black" filter): (c * (255 - gray) + 255 * gray) / 255
  FM_RAINBOW = 3  // FM_SCALE with gray lifted to 179 + gray * 76 / 255 (color_filter.c:305-318)
};

struct StitchParams {
  const uint8_t *rows;
  const RowMeta *meta;
  uint32_t row_pitch;
  int text_rows;
  int pad_top;
  int mode;
  uint8_t *out;      // frame f at f*out_pitch
  size_t out_pitch;
  uint32_t *out_len; // [f]
  int rows_per_cta;
};

__host__ __device__ inline uint32_t al16(uint32_t v) { return (v + 15u) & ~15u; }

size_t rows_smem_total(int mode, int scale_path, int cols, int src_w, uint32_t out_bytes);
uint32_t row_capacity_bytes(int mode, int cols, int pad_left);
static constexpr int kSmemOutMax = 48 * 1024;      // rows up to this many bytes are staged in shared memory
static constexpr uint32_t kMaxDynSmem = 224u * 1024u; // dynamic smem ceiling (227 KB opt-in minus static, incl. the 1 KB digit table)

// *grid_out (optional) = CTAs launched; with p.direct the kernel is persistent and every CTA draws one ticket past the end
cudaError_t launch_render_rows(const RenderParams &p, int mode, int scale_path, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_render_rows_ws(const RenderParams &p, int mode, cudaStream_t st);
cudaError_t launch_render_rows_ws2(const RenderParams &p, int mode, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_quantize_table(int which, uint8_t *d_out, cudaStream_t st);
size_t ws2_smem_total(int mode, int cols, int src_w, uint32_t row_pitch);
int ws_ring_depth(int mode, int cols, int src_w, uint32_t row_pitch);
cudaError_t launch_stitch(const StitchParams &p, int n_frames, cudaStream_t st);
cudaError_t launch_resize_nn_only(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh, int pregathered,
                                  cudaStream_t st);
// NN sampling, one CTA per sampled row (whole rows read with 16-byte loads, any alignment): dst = cols x rows RGB24
cudaError_t launch_gather_nn_rows(const uint8_t *src_dev, int sw, int sh, int cols, int rows, int flip_x, int flip_y,
                                  uint8_t *dst, cudaStream_t st, int wide = 0);
// Floyd–Steinberg 16-colour background renderer (serial wavefront): one CTA per frame, reads the resized image
// fg_only: print the dithered colour as the foreground (the two foreground-only leaf printers) instead of bg + contrast fg
cudaError_t launch_dither_bg(const uint8_t *cells, int w, int h, int n_frames, int pad_left, const GlyphLut *lut,
                             uint8_t *rows, uint32_t row_pitch, RowMeta *meta, int *err_scratch, int fg_only,
                             cudaStream_t st);
// the whole pixel-space composite in one launch (stream.c:664-779): up to 9 sources (stream.c:687)
struct CompositeCell {
  const uint8_t *src; // nullptr: the cell stays black (no video / degenerate fit)
  int sw, sh;         // source size
  int tw, th;         // contain-fitted target size inside the cell (stream.c:708-716)
  int xp, yp;         // centring offsets (cellw - tw) / 2, (cellh - th) / 2
  uint32_t xr, yr;    // 16.16 NN ratios ((sw << 16) / tw) + 1, ((sh << 16) / th) + 1
};
struct CompositeParams {
  uint8_t *comp;
  int cw, ch, cellw, cellh, gcols, grows, n;
  CompositeCell cell[9];
};
cudaError_t launch_composite_all(const CompositeParams &p, cudaStream_t st);
// pixel-space composite blit (stream.c:752-773): NN-resize one source into its clipped cell of the composite
cudaError_t launch_composite_cell(const uint8_t *src, int sw, int sh, uint8_t *comp, int cw, int ch, int tw, int th,
                                  int x0, int y0, int cellw, int cellh, cudaStream_t st);

} // namespace acb
