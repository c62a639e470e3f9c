// render_kernels.cu — sm_100a kernels for the RGB -> glyph/ANSI render path.
//
// k_render_rows<MODE, SP>: one CTA per (frame, text row).  It (A) produces the row's resized
// pixels in shared memory — nearest-neighbour sampling (image.c:267-328) or a box filter that
// streams the whole source band with 16-byte loads — and (B) turns them into the exact byte
// string of the reference's renderer for that mode (foreground.c / halfblock.c grammars) using
// cell-local rules: every cell derives, from its run (head, length), its left run and the LUTs,
// which bytes it owns; a block-wide exclusive scan of the byte counts places them.  The row is
// staged in shared memory and leaves the SM as 16-byte stores.
//
// k_stitch: concatenates the fixed-pitch rows of a frame into the final NUL-terminated string
// (row-length prefix sum; for truecolor-foreground also the cross-row colour carry of
// ansi_rle_add_pixel, ansi.c:261-300).
//
// All arithmetic is integer; see oracle/ascii_oracle.c for the CPU restatement these kernels
// are checked against byte for byte.
#include "render_dev.cuh"

namespace acb {


uint32_t row_capacity_bytes(int mode, int cols, int pad_left) { // SURVEY.md §8a grammar table maxima
  uint32_t per;
  switch (mode) {
  case EM_MONO_FG: per = 4; break;
  case EM_256_FG: per = 11 + 4; break;
  case EM_16_FG: per = 5 + 4; break;
  case EM_TRUE_FG: per = 19 + 4; break;
  case EM_HB_TRUE: per = 19 + 19 + 3; break;
  case EM_HB_256: per = 11 + 11 + 3; break;
  case EM_HB_16: per = 5 + 6 + 3; break;
  case EM_HB_MONO: per = 3; break;
  default: per = 6 + 5 + 4; break; // EM_DITHER_BG (and its foreground-only forms, which need less)
  }
  // a run head may also own a REP tail (<= 2+5+1 bytes) or a reset before a transparent run (4 bytes):
  // both replace glyph bytes of other cells, so `per` per cell + 16 slack bounds the row
  return al16((uint32_t)pad_left + per * (uint32_t)cols + 4u /*reset*/ + 1u /*\n*/ + 4u /*frame reset*/ + 16u);
}


// ------------------------------------------------------------------ stitch rows -> frame string
// Block-cooperative copy of n bytes for any mutual alignment: destination-aligned 4-byte stores assembled from
// two source-aligned words with one byte permute (the scratch rows keep >= 16 bytes of slack, so reading one
// word past the range stays inside the row).
__device__ __forceinline__ void copy_shifted(uint8_t *dst, const uint8_t *src, uint32_t n) {
  const uint32_t tid = threadIdx.x, nt = blockDim.x;
  uint32_t head = (4u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 3u)) & 3u;
  if (head > n) head = n;
  if (tid < head) dst[tid] = src[tid];
  dst += head;
  src += head;
  n -= head;
  const uint32_t nw = n >> 2;
  const uint32_t r = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
  const uint32_t *sw = reinterpret_cast<const uint32_t *>(src - r);
  const uint32_t sel = 0x3210u + 0x1111u * r;
  uint32_t *dw = reinterpret_cast<uint32_t *>(dst);
  for (uint32_t j = tid; j < nw; j += nt) dw[j] = __byte_perm(sw[j], sw[j + 1], sel);
  const uint32_t tail = n & 3u;
  if (tid < tail) dst[4u * nw + tid] = src[4u * nw + tid];
}

__global__ void __launch_bounds__(256) k_stitch(const StitchParams p) {
  constexpr int CH = 256; // rows of metadata staged per pass
  __shared__ uint32_t s_len[CH], s_cl[CH], s_first[CH], s_last[CH];
  __shared__ uint32_t s_off[64], s_drop[64];
  __shared__ uint32_t s_total, s_o, s_carry;
  const int groups = (p.text_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  const int f = (int)(blockIdx.x / (unsigned)groups);
  const int g = (int)(blockIdx.x % (unsigned)groups);
  const int r0 = g * p.rows_per_cta;
  const int r1 = min(r0 + p.rows_per_cta, p.text_rows);
  const RowMeta *meta = p.meta + (size_t)f * p.text_rows;
  const bool tf = p.mode == EM_TRUE_FG;
  const int upto = (g == groups - 1) ? p.text_rows : r1;
  if (threadIdx.x == 0) {
    s_o = (uint32_t)p.pad_top;
    s_carry = 0;
  }
  for (int base = 0; base < upto; base += CH) {
    const int nrows = min(CH, upto - base);
    __syncthreads();
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) { // parallel fetch, 16 bytes of each 32-byte record
      const uint4 m = *reinterpret_cast<const uint4 *>(&meta[base + i]);
      s_len[i] = m.x;
      s_cl[i] = m.z;
      s_first[i] = m.w;
      s_last[i] = tf ? meta[base + i].last_rgb : 0u;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t o = s_o, carry = s_carry;
      for (int i = 0; i < nrows; i++) {
        uint32_t drop = 0;
        if (tf) {
          // ansi_rle_add_pixel state across rows: the first ASCII cell of a row re-emits its SGR only if its
          // colour differs from the last ASCII cell seen anywhere above (ansi.c:263)
          if (s_cl[i] && carry && s_first[i] == carry) drop = s_cl[i];
          if (s_last[i]) carry = s_last[i];
        }
        const int r = base + i;
        if (r >= r0 && r < r1) {
          s_off[r - r0] = o;
          s_drop[r - r0] = drop;
        }
        o += s_len[i] - drop;
      }
      s_o = o;
      s_carry = carry;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) s_total = s_o;
  __syncthreads();
  uint8_t *out = p.out + (size_t)f * p.out_pitch;
  if (g == 0)
    for (int i = threadIdx.x; i < p.pad_top; i += blockDim.x) out[i] = '\n';
  if (g == groups - 1 && threadIdx.x == 0) {
    out[s_total] = 0;
    p.out_len[f] = s_total;
  }
  for (int r = r0; r < r1; r++) {
    const RowMeta m = meta[r];
    const uint8_t *src = p.rows + ((size_t)f * p.text_rows + r) * (size_t)p.row_pitch;
    uint8_t *dst = out + s_off[r - r0];
    const uint32_t drop = s_drop[r - r0];
    const uint32_t a = drop ? m.cond_off : m.len; // [0,a) then [a+drop, len)
    copy_shifted(dst, src, a);
    if (drop) copy_shifted(dst + a, src + a + drop, m.len - a - drop);
  }
}

// ------------------------------------------------------------------ NN resize only (image_resize drop-in)
__global__ void k_resize_nn(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh, int pregathered) {
  const uint32_t xr = (uint32_t)((((uint64_t)sw << 16) / (uint64_t)dw) + 1);
  const uint32_t yr = (uint32_t)((((uint64_t)sh << 16) / (uint64_t)dh) + 1);
  const size_t n = (size_t)dw * dh;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t y = (uint32_t)(i / dw), x = (uint32_t)(i % dw);
    uint32_t sy = pregathered ? y : ((y * yr) >> 16);
    if (!pregathered && sy >= (uint32_t)sh) sy = (uint32_t)sh - 1;
    uint32_t sx = (x * xr) >> 16;
    if (sx >= (uint32_t)sw) sx = (uint32_t)sw - 1;
    const uint8_t *q = src + ((size_t)sy * sw + sx) * 3u;
    uint8_t *d = dst + i * 3u;
    d[0] = q[0];
    d[1] = q[1];
    d[2] = q[2];
  }
}

// ------------------------------------------------------------------ NN sampling, a CTA per sampled row
// image_resize's sampling (image.c:293-325): one CTA per SAMPLED source row reads the whole row with 16-byte loads,
// parks it in shared memory and writes the row's `cols` sampled pixels.  Used behind the copy engine when a frame lives
// in page-locked host memory (engine.cu: the rows NN reads are fetched by strided 2-D copies, this kernel then samples
// the columns at 1:1 in y).  It also runs on rows in mapped host memory directly, but the SMs' PCIe reads are
// sector-sized: 9.5 GB/s for a lone 4K frame (profiles/r02m_e2e_registered1.txt), which is why the copy engine does the
// fetch.  Any alignment of `src`: the 16-byte chunks that lie wholly inside the row are loaded as vectors, the ragged
// ends byte by byte (never a byte outside the row).
template <bool WIDE>
__global__ void __launch_bounds__(256) k_gather_nn_rows(const uint8_t *src, int sw, int sh, int cols, int rows,
                                                        uint32_t xr, uint32_t yr, int flip_x, int flip_y,
                                                        uint8_t *dst) {
  extern __shared__ uint4 s_row4[];
  uint8_t *s_row = reinterpret_cast<uint8_t *>(s_row4);
  const int y = blockIdx.x, tid = threadIdx.x;
  uint32_t sy = ((uint32_t)y * yr) >> 16;
  if (sy >= (uint32_t)sh) sy = (uint32_t)sh - 1u;
  if (flip_y) sy = (uint32_t)sh - 1u - sy;
  const size_t R = (size_t)sw * 3u;
  const uint8_t *row = src + (size_t)sy * R;
  const uintptr_t b = reinterpret_cast<uintptr_t>(row), e = b + R;
  const uintptr_t a0 = b & ~(uintptr_t)15;              // s_row[i] mirrors the byte at a0 + i
  const uintptr_t lo = (b + 15u) & ~(uintptr_t)15, hi = e & ~(uintptr_t)15;
  if (lo < hi) {
    for (uintptr_t a = lo + (uintptr_t)tid * 16u; a < hi; a += 256u * 16u) {
      uint4 v;
      if (WIDE) // rows in mapped host memory: ask the L2 for 256-byte fetches (fewer, larger PCIe reads)
        asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(a));
      else
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "l"(a));
      *reinterpret_cast<uint4 *>(s_row + (a - a0)) = v;
    }
    if ((uintptr_t)tid < lo - b) s_row[(b - a0) + tid] = row[tid];
    if (tid >= 32 && (uintptr_t)(tid - 32) < e - hi) s_row[(hi - a0) + (tid - 32)] = *reinterpret_cast<const uint8_t *>(hi + (tid - 32));
  } else {
    for (size_t i = tid; i < R; i += 256) s_row[(b - a0) + i] = row[i];
  }
  __syncthreads();
  const uint8_t *q0 = s_row + (b - a0);
  for (int x = tid; x < cols; x += 256) {
    uint32_t sx = ((uint32_t)x * xr) >> 16;
    if (sx >= (uint32_t)sw) sx = (uint32_t)sw - 1u;
    if (flip_x) sx = (uint32_t)sw - 1u - sx;
    const uint8_t *q = q0 + sx * 3u;
    uint8_t *d = dst + ((size_t)y * cols + x) * 3u;
    d[0] = q[0];
    d[1] = q[1];
    d[2] = q[2];
  }
}

// ------------------------------------------------------------------ pixel-space composite cell (stream.c:723-773)
__global__ void k_composite_cell(const uint8_t *src, int sw, int sh, uint8_t *comp, int cw, int ch, int tw, int th,
                                 int x0, int y0, int cellw, int cellh) {
  const uint32_t xr = (uint32_t)((((uint64_t)sw << 16) / (uint64_t)tw) + 1);
  const uint32_t yr = (uint32_t)((((uint64_t)sh << 16) / (uint64_t)th) + 1);
  const int xp = (cellw - tw) / 2, yp = (cellh - th) / 2;
  const int n = tw * th;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int y = i / tw, x = i % tw;
    int dx = x0 + xp + x, dy = y0 + yp + y;
    if (dx < x0 || dx > x0 + cellw - 1 || dy < y0 || dy > y0 + cellh - 1) continue;
    if (dx < 0 || dx >= cw || dy < 0 || dy >= ch) continue;
    uint32_t sy = ((uint32_t)y * yr) >> 16, sx = ((uint32_t)x * xr) >> 16;
    if (sy >= (uint32_t)sh) sy = (uint32_t)sh - 1;
    if (sx >= (uint32_t)sw) sx = (uint32_t)sw - 1;
    const uint8_t *q = src + ((size_t)sy * sw + sx) * 3u;
    uint8_t *d = comp + ((size_t)dy * cw + dx) * 3u;
    d[0] = q[0];
    d[1] = q[1];
    d[2] = q[2];
  }
}

// The whole composite in one launch (server path with resident sources): every composite pixel finds its grid cell,
// and inside the cell's centred target rectangle samples its source (same 16.16 NN arithmetic as image_resize,
// image.c:293-325); everything else is the black canvas of image_clear (stream.c:683).  Cells are disjoint and every
// blit is clipped to its own cell (stream.c:752-773), so this gather writes exactly what the reference's
// clear + per-source blits leave behind — with one kernel instead of a memset and up to nine launches.
__global__ void __launch_bounds__(256) k_composite_all(const CompositeParams p) {
  const int n = p.cw * p.ch;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int dy = i / p.cw, dx = i - dy * p.cw;
    const int col = dx / p.cellw, row = dy / p.cellh;
    uint32_t r = 0, g = 0, b = 0;
    const int v = row * p.gcols + col;
    if (col < p.gcols && row < p.grows && v < p.n) {
      const CompositeCell &c = p.cell[v];
      const int x = dx - col * p.cellw - c.xp, y = dy - row * p.cellh - c.yp;
      if (c.src && x >= 0 && x < c.tw && y >= 0 && y < c.th) {
        uint32_t sy = ((uint32_t)y * c.yr) >> 16, sx = ((uint32_t)x * c.xr) >> 16;
        if (sy >= (uint32_t)c.sh) sy = (uint32_t)c.sh - 1;
        if (sx >= (uint32_t)c.sw) sx = (uint32_t)c.sw - 1;
        const uint8_t *q = c.src + ((size_t)sy * c.sw + sx) * 3u;
        r = q[0], g = q[1], b = q[2];
      }
    }
    uint8_t *d = p.comp + (size_t)i * 3u;
    d[0] = (uint8_t)r;
    d[1] = (uint8_t)g;
    d[2] = (uint8_t)b;
  }
}

// ------------------------------------------------------------------ Floyd–Steinberg 16-colour background renderer
// rgb_to_16color_dithered (ansi.c:511-583) is a raster-order recurrence; pixel (x,y) only needs the errors pushed by
// (x-1,y) and (x-1..x+1, y-1), so row y may trail row y-1 by 3 pixels: one thread per row, a skewed wavefront of
// w + 3*(rows-1) steps.  Integer += is order independent, so the sums equal the serial ones.
__global__ void __launch_bounds__(256) k_dither_bg(const uint8_t *cells, int w, int h, int pad_left, const GlyphLut *lut,
                                                   uint8_t *rows, uint32_t row_pitch, RowMeta *meta, int *err_all,
                                                   int fg_only) {
  const int f = blockIdx.x;
  const uint8_t *img = cells + (size_t)f * w * h * 3u;
  int *err = err_all + (size_t)f * w * h * 3u;
  for (int band = 0; band < h; band += blockDim.x) {
    const int j = threadIdx.x;
    const int y = band + j;
    const int nrows = min((int)blockDim.x, h - band);
    const bool active = y < h;
    uint8_t *out = active ? rows + ((size_t)f * h + y) * (size_t)row_pitch : nullptr;
    uint32_t o = 0;
    if (active)
      for (int i = 0; i < pad_left; i++) out[o++] = ' ';
    const int steps = w + 3 * (nrows - 1);
    for (int s = 0; s < steps; s++) {
      const int x = s - 3 * j;
      if (active && x >= 0 && x < w) {
        const size_t i = (size_t)y * w + x;
        const uint8_t *px = img + i * 3u;
        int v[3], c[3];
#pragma unroll
        for (int k = 0; k < 3; k++) {
          v[k] = (int)px[k] + err[i * 3 + k];
          c[k] = v[k] < 0 ? 0 : (v[k] > 255 ? 255 : v[k]);
        }
        const int q = q16_rgb(c[0], c[1], c[2]);
#pragma unroll
        for (int k = 0; k < 3; k++) {
          const int e = v[k] - (int)c_ansi16[q][k];
          if (x + 1 < w) err[(i + 1) * 3 + k] += (e * 7) / 16;
          if (y + 1 < h) {
            if (x - 1 >= 0) err[(i + w - 1) * 3 + k] += (e * 3) / 16;
            err[(i + w) * 3 + k] += (e * 5) / 16;
            if (x + 1 < w) err[(i + w + 1) * 3 + k] += (e * 1) / 16;
          }
        }
        const int bl = ((int)c_ansi16[q][0] * 77 + (int)c_ansi16[q][1] * 150 + (int)c_ansi16[q][2] * 29) / 256;
        WriteSink ws{out + o};
        if (fg_only) {
          put_sgr_16(ws, false, (uint32_t)q);             // foreground.c:713, 811
        } else {
          put_sgr_16(ws, true, (uint32_t)q);              // foreground.c:807
          put_sgr_16(ws, false, bl < 127 ? 15u : 0u);     // foreground.c:804-808
        }
        put_glyph(ws, lut->glyph[luma_of(load_px(px))]);  // cache[Y] (:820) or, through the Q2 table, cache[ramp] (:722)
        o = (uint32_t)(ws.p - out);
      }
      __syncthreads();
    }
    if (active) {
      WriteSink ws{out + o};
      put_reset(ws);
      if (y < h - 1) ws.put('\n');
      RowMeta m{};
      m.len = (uint32_t)(ws.p - out);
      for (int z = 0; z < 4; z++) ws.put(0); // k_stitch reads up to one word past the row (copy_shifted)
      meta[(size_t)f * h + y] = m;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------ exhaustive quantiser tables (parity aid)
// out[r << 16 | g << 8 | b] = rgb_to_256color (which = 0) / rgb_to_16color (which = 1) exactly as the emitters compute
// them: the whole 2^24 colour space in one launch, compared against the reference's table in the tests.
__global__ void __launch_bounds__(256) k_quantize_table(int which, uint8_t *out) {
  const uint32_t i = blockIdx.x * 256u + threadIdx.x;
  out[i] = (uint8_t)(which == 0 ? q256_of(i) : q16_of(i));
}
cudaError_t launch_quantize_table(int which, uint8_t *d_out, cudaStream_t st) {
  k_quantize_table<<<(1u << 24) / 256u, 256, 0, st>>>(which, d_out);
  return cudaGetLastError();
}

// ------------------------------------------------------------------ launchers
cudaError_t launch_rows_m0(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m0(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m1(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m1(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m2(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m2(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m3(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m3(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m4(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m4(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m5(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m5(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m6(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m6(const RenderParams &p, cudaStream_t st);
cudaError_t launch_rows_m7(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws_m7(const RenderParams &p, cudaStream_t st);
cudaError_t launch_ws2_m0(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m1(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m2(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m3(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m4(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m5(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m6(const RenderParams &p, cudaStream_t st, unsigned *grid_out);
cudaError_t launch_ws2_m7(const RenderParams &p, cudaStream_t st, unsigned *grid_out);

cudaError_t launch_render_rows(const RenderParams &p, int mode, int sp, cudaStream_t st, unsigned *grid_out) {
  switch (mode) {
  case EM_MONO_FG: return launch_rows_m0(p, sp, st, grid_out);
  case EM_256_FG: return launch_rows_m1(p, sp, st, grid_out);
  case EM_16_FG: return launch_rows_m2(p, sp, st, grid_out);
  case EM_TRUE_FG: return launch_rows_m3(p, sp, st, grid_out);
  case EM_HB_TRUE: return launch_rows_m4(p, sp, st, grid_out);
  case EM_HB_256: return launch_rows_m5(p, sp, st, grid_out);
  case EM_HB_16: return launch_rows_m6(p, sp, st, grid_out);
  case EM_HB_MONO: return launch_rows_m7(p, sp, st, grid_out);
  default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_render_rows_ws(const RenderParams &p, int mode, cudaStream_t st) {
  switch (mode) {
  case EM_MONO_FG: return launch_ws_m0(p, st);
  case EM_256_FG: return launch_ws_m1(p, st);
  case EM_16_FG: return launch_ws_m2(p, st);
  case EM_TRUE_FG: return launch_ws_m3(p, st);
  case EM_HB_TRUE: return launch_ws_m4(p, st);
  case EM_HB_256: return launch_ws_m5(p, st);
  case EM_HB_16: return launch_ws_m6(p, st);
  case EM_HB_MONO: return launch_ws_m7(p, st);
  default: return cudaErrorInvalidValue;
  }
}
cudaError_t launch_render_rows_ws2(const RenderParams &p, int mode, cudaStream_t st, unsigned *grid_out) {
  switch (mode) {
  case EM_MONO_FG: return launch_ws2_m0(p, st, grid_out);
  case EM_256_FG: return launch_ws2_m1(p, st, grid_out);
  case EM_16_FG: return launch_ws2_m2(p, st, grid_out);
  case EM_TRUE_FG: return launch_ws2_m3(p, st, grid_out);
  case EM_HB_TRUE: return launch_ws2_m4(p, st, grid_out);
  case EM_HB_256: return launch_ws2_m5(p, st, grid_out);
  case EM_HB_16: return launch_ws2_m6(p, st, grid_out);
  case EM_HB_MONO: return launch_ws2_m7(p, st, grid_out);
  default: return cudaErrorInvalidValue;
  }
}
size_t ws2_smem_total(int mode, int cols, int src_w, uint32_t row_pitch) {
  return make_layout2(mode, 0, cols, src_w, row_pitch).total; // worst case (scratch-row variant)
}
// ring depth for the warp-specialised kernel: as many source rows as fit half an SM's shared memory (two CTAs per
// SM), 0 if the geometry does not qualify
int ws_ring_depth(int mode, int cols, int src_w, uint32_t row_pitch) {
  const Layout L = make_layout(mode, SP_BOX_STREAM, cols, src_w, row_pitch, 0);
  const size_t fixed = (L.total + 127u) & ~127u;
  const size_t R = (size_t)src_w * 3u;
  const size_t budget = (227u * 1024u) / 2u - 2048u;
  if (fixed + 3 * R > budget) {
    if (fixed + 3 * R > kMaxDynSmem) return 0;
    size_t d1 = (kMaxDynSmem - fixed) / R; // one CTA per SM
    return (int)(d1 > WS_MAXD ? WS_MAXD : d1);
  }
  size_t d = (budget - fixed) / R;
  return (int)(d > WS_MAXD ? WS_MAXD : d);
}

size_t rows_smem_total(int mode, int sp, int cols, int src_w, uint32_t out_bytes) {
  return make_layout(mode, sp, cols, src_w, out_bytes, 1, 1).total; // worst case (no aliasing, both tile sets)
}

cudaError_t launch_stitch(const StitchParams &p, int n_frames, cudaStream_t st) {
  const int groups = (p.text_rows + p.rows_per_cta - 1) / p.rows_per_cta;
  k_stitch<<<(unsigned)n_frames * (unsigned)groups, 256, 0, st>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_resize_nn_only(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh, int pregathered,
                                  cudaStream_t st) {
  size_t n = (size_t)dw * dh;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148u * 16u) grid = 148u * 16u;
  if (grid == 0) grid = 1;
  k_resize_nn<<<grid, 256, 0, st>>>(src, sw, sh, dst, dw, dh, pregathered);
  return cudaGetLastError();
}

cudaError_t launch_gather_nn_rows(const uint8_t *src_dev, int sw, int sh, int cols, int rows, int flip_x, int flip_y,
                                  uint8_t *dst, cudaStream_t st, int wide) {
  const uint32_t xr = (uint32_t)((((uint64_t)sw << 16) / (uint64_t)cols) + 1);
  const uint32_t yr = (uint32_t)((((uint64_t)sh << 16) / (uint64_t)rows) + 1);
  const size_t smem = (size_t)sw * 3u + 32u;
  if (smem > 48u * 1024u) return cudaErrorInvalidValue; // callers keep such rows on the staged plan
  if (wide)
    k_gather_nn_rows<true><<<(unsigned)rows, 256, smem, st>>>(src_dev, sw, sh, cols, rows, xr, yr, flip_x, flip_y, dst);
  else
    k_gather_nn_rows<false><<<(unsigned)rows, 256, smem, st>>>(src_dev, sw, sh, cols, rows, xr, yr, flip_x, flip_y, dst);
  return cudaGetLastError();
}

cudaError_t launch_dither_bg(const uint8_t *cells, int w, int h, int n_frames, int pad_left, const GlyphLut *lut,
                             uint8_t *rows, uint32_t row_pitch, RowMeta *meta, int *err_scratch, int fg_only,
                             cudaStream_t st) {
  k_dither_bg<<<(unsigned)n_frames, 256, 0, st>>>(cells, w, h, pad_left, lut, rows, row_pitch, meta, err_scratch, fg_only);
  return cudaGetLastError();
}

cudaError_t launch_composite_cell(const uint8_t *src, int sw, int sh, uint8_t *comp, int cw, int ch, int tw, int th,
                                  int x0, int y0, int cellw, int cellh, cudaStream_t st) {
  int n = tw * th;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  if (grid == 0) grid = 1;
  k_composite_cell<<<grid, 256, 0, st>>>(src, sw, sh, comp, cw, ch, tw, th, x0, y0, cellw, cellh);
  return cudaGetLastError();
}

cudaError_t launch_composite_all(const CompositeParams &p, cudaStream_t st) {
  const int n = p.cw * p.ch;
  unsigned grid = (unsigned)((n + 255) / 256);
  if (grid > 148u * 8u) grid = 148u * 8u;
  if (grid == 0) grid = 1;
  k_composite_all<<<grid, 256, 0, st>>>(p);
  return cudaGetLastError();
}

} // namespace acb
