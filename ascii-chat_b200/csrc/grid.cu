// grid.cu — the two grid compositors of the reference, on the device.
//
//  * text-space: ascii_create_grid (lib/video/ascii/ascii.c:602-885) — N rendered frames ->
//    one W x H character canvas with '|', '_', '+' separators.  Layout scoring is host float
//    (ceil/logf/fabsf, ascii.c:712-769); everything that touches frame bytes runs in k_text_grid.
//  * pixel-space: create_multi_source_composite (src/server/stream.c:664-779) with
//    calculate_optimal_grid_layout (stream.c:523-651) — N RGB sources -> one W x 2H composite;
//    the NN resize + clipped blit of every source is k_composite_cell (render_kernels.cu).
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.h"

namespace acb {

// ------------------------------------------------------------------ text grid kernel
struct TextGridParams {
  const uint8_t *const *src; // [n] device pointers
  const uint32_t *size;      // [n]
  int n, W, H;
  int gcols, grows, cw, ch;  // layout (multi-source path)
  int single;                // n == 1: centre the lone frame (ascii.c:610-707)
  int maxlines;              // lines kept per source: ch (multi) or H (single)
  uint8_t *out;
  uint32_t total;            // W*H + H + 1
  uint32_t *lstart, *llen, *clen, *dcol; // [n * maxlines] scratch
  uint32_t *nlines, *nl_total;           // [n]
};

// ansi_truncate_to_visual_width (ascii.c:562-586); *vis = visible characters inside the returned prefix,
// which equals ansi_visual_width(prefix) (ascii.c:527-551) because the prefix ends on a token boundary.
__device__ static int truncate_visible(const uint8_t *d, int n, int target, int *vis) {
  int v = 0, i = 0;
  while (i < n && v < target) {
    if (d[i] == 0x1b && i + 1 < n && d[i + 1] == '[') {
      i += 2;
      while (i < n) {
        uint8_t c = d[i++];
        if (c >= '@' && c <= '~') break;
      }
    } else {
      v++;
      i++;
    }
  }
  *vis = v;
  return i;
}

__global__ void __launch_bounds__(256) k_text_grid(const TextGridParams p) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = blockDim.x >> 5;
  const uint32_t W1 = (uint32_t)p.W + 1u;
  // phase 0: blank canvas, '\n' closing every row, NUL (ascii.c:633-640, 806-813)
  for (uint32_t i = tid; i < p.total - 1; i += blockDim.x) p.out[i] = (i % W1 == (uint32_t)p.W) ? '\n' : ' ';
  if (tid == 0) p.out[p.total - 1] = 0;

  // phase 1: line tables, one warp per source, 32 bytes per step
  for (int s = warp; s < p.n; s += nwarp) {
    const uint8_t *d = p.src[s];
    const uint32_t size = d ? p.size[s] : 0u;
    uint32_t line = 0, start = 0, nl = 0;
    for (uint32_t base = 0; base < size; base += 32) {
      uint32_t i = base + lane;
      uint32_t m = __ballot_sync(0xffffffffu, i < size && d[i] == '\n');
      nl += __popc(m);
      while (m && line < (uint32_t)p.maxlines) {
        uint32_t e = base + (uint32_t)(__ffs(m) - 1);
        if (lane == 0) {
          p.lstart[(size_t)s * p.maxlines + line] = start;
          p.llen[(size_t)s * p.maxlines + line] = e - start;
        }
        line++;
        start = e + 1;
        m &= m - 1;
      }
    }
    if (start < size && line < (uint32_t)p.maxlines) { // last line without a trailing newline
      if (lane == 0) {
        p.lstart[(size_t)s * p.maxlines + line] = start;
        p.llen[(size_t)s * p.maxlines + line] = size - start;
      }
      line++;
    }
    if (lane == 0) {
      p.nlines[s] = line;
      p.nl_total[s] = nl;
    }
  }
  __syncthreads();

  // phase 2: per line, ANSI-aware truncation and the copy decision
  const int items = p.n * p.maxlines;
  for (int it = tid; it < items; it += blockDim.x) {
    const int s = it / p.maxlines, r = it % p.maxlines;
    if ((uint32_t)r >= p.nlines[s]) continue;
    const uint8_t *line = p.src[s] + p.lstart[it];
    const int ll = (int)p.llen[it];
    uint32_t cl = 0, col = 0;
    if (p.single) { // ascii.c:661-703
      int vpad = (p.H - (int)p.nl_total[0]) / 2;
      if (vpad < 0) vpad = 0;
      const int row = vpad + r;
      if (row < p.H) {
        int vis;
        truncate_visible(line, ll, 0x7fffffff, &vis); // ansi_visual_width of the whole line
        int hpad = (p.W - vis) / 2;
        if (hpad < 0) hpad = 0;
        int tv;
        int c = truncate_visible(line, ll, p.W - hpad, &tv);
        const size_t dst = (size_t)row * W1 + (size_t)hpad;
        if (c > 0 && dst + (size_t)c < (size_t)p.total) {
          cl = (uint32_t)c;
          col = (uint32_t)dst;
        }
      }
    } else { // ascii.c:829-852
      const int gr = s / p.gcols, gc = s % p.gcols;
      const int r0 = gr * (p.ch + 1), c0 = gc * (p.cw + 1);
      if (r < p.ch && r0 + r < p.H) {
        int tv;
        int c = truncate_visible(line, ll, p.cw, &tv);
        const size_t pos = (size_t)(r0 + r) * W1 + (size_t)c0;
        // SAFE_MEMCPY refuses when count > remaining size (lib/platform/posix/system.c:653-666)
        if (c > 0 && c0 + tv <= p.W && (size_t)c <= (size_t)p.total - pos) {
          cl = (uint32_t)c;
          col = (uint32_t)pos;
        }
      }
    }
    p.clen[it] = cl;
    p.dcol[it] = col;
  }
  __syncthreads();

  // phase 3: the writes, in the reference's order (later writes win where ANSI bytes spill over)
  for (int s = 0; s < p.n; s++) {
    const uint32_t nl = p.nlines[s];
    for (uint32_t r = 0; r < nl; r++) {
      const size_t it = (size_t)s * p.maxlines + r;
      const uint32_t cl = p.clen[it];
      if (cl) {
        const uint8_t *line = p.src[s] + p.lstart[it];
        uint8_t *dst = p.out + p.dcol[it];
        for (uint32_t i = tid; i < cl; i += blockDim.x) dst[i] = line[i];
      }
      __syncthreads();
    }
    if (p.single) continue;
    const int gr = s / p.gcols, gc = s % p.gcols;
    const int r0 = gr * (p.ch + 1), c0 = gc * (p.cw + 1);
    if (gc < p.gcols - 1 && c0 + p.cw < p.W) { // vertical separator, ascii.c:855-863
      for (int row = r0 + tid; row < r0 + p.ch && row < p.H; row += blockDim.x) {
        size_t idx = (size_t)row * W1 + (size_t)(c0 + p.cw);
        if (idx < (size_t)p.total - 1) p.out[idx] = '|';
      }
    }
    __syncthreads();
    if (gr < p.grows - 1 && r0 + p.ch < p.H) { // horizontal separator + corner, ascii.c:865-880
      for (int c = c0 + tid; c < c0 + p.cw && c < p.W; c += blockDim.x) {
        size_t idx = (size_t)(r0 + p.ch) * W1 + (size_t)c;
        if (idx < (size_t)p.total - 1) p.out[idx] = '_';
      }
      __syncthreads();
      if (tid == 0 && gc < p.gcols - 1 && c0 + p.cw < p.W) {
        size_t idx = (size_t)(r0 + p.ch) * W1 + (size_t)(c0 + p.cw);
        if (idx < (size_t)p.total - 1) p.out[idx] = '+';
      }
    }
    __syncthreads();
  }
  if (tid == 0) p.out[p.total - 1] = 0; // a spill may have landed on the terminator (reference: UB)
}

// ------------------------------------------------------------------ host float layouts
// ascii.c:712-776
static void text_grid_layout(int n, int width, int height, int *cols, int *rows, int *cw, int *ch) {
  float char_aspect = 2.0f;
  float best_score = -1.0f;
  int best_cols = 1, best_rows = n;
  for (int test_cols = 1; test_cols <= n; test_cols++) {
    int test_rows = (int)ceil((double)n / test_cols);
    int empty_cells = (test_cols * test_rows) - n;
    if (empty_cells > n / 2) continue;
    int cell_width = (width - (test_cols - 1)) / test_cols;
    int cell_height = (height - (test_rows - 1)) / test_rows;
    if (cell_width < 10 || cell_height < 3) continue;
    float cell_aspect = ((float)cell_width / (float)cell_height) / char_aspect;
    float aspect_score = 1.0f - fabsf(logf(cell_aspect));
    if (aspect_score < 0) aspect_score = 0;
    float utilization = (float)n / (float)(test_cols * test_rows);
    float total_score = n == 2 ? aspect_score * 0.9f + utilization * 0.1f : aspect_score * 0.7f + utilization * 0.3f;
    if (test_cols == test_rows) total_score += 0.05f;
    if (total_score > best_score) {
      best_score = total_score;
      best_cols = test_cols;
      best_rows = test_rows;
    }
  }
  *cols = best_cols;
  *rows = best_rows;
  *cw = (width - (best_cols - 1)) / best_cols;
  *ch = (height - (best_rows - 1)) / best_rows;
}

static int text_grid_device(const uint8_t *const *d_srcs, const size_t *sizes, int n, int width, int height,
                            uint8_t *d_out, size_t *out_size, cudaStream_t st, ThreadCtx *cx) {
  const uint32_t total = (uint32_t)((size_t)width * height + height + 1);
  TextGridParams p{};
  p.n = n;
  p.W = width;
  p.H = height;
  p.out = d_out;
  p.total = total;
  p.single = n == 1;
  if (n == 1) {
    p.maxlines = height;
    p.gcols = p.grows = 1;
    p.cw = width;
    p.ch = height;
  } else {
    text_grid_layout(n, width, height, &p.gcols, &p.grows, &p.cw, &p.ch);
    if (p.cw < 10 || p.ch < 3) { // too small for a grid: first source as-is (ascii.c:778-792)
      if (d_srcs[0] && sizes[0] > 0) ACB_CUDA(cudaMemcpyAsync(d_out, d_srcs[0], sizes[0], cudaMemcpyDeviceToDevice, st));
      const size_t z = (d_srcs[0] && sizes[0] > 0) ? sizes[0] : 0;
      ACB_CUDA(cudaMemsetAsync(d_out + z, 0, 1, st));
      *out_size = z;
      return E_OK;
    }
    p.maxlines = p.ch;
  }
  // parameter block + scratch in one device allocation of the thread context
  const size_t nl = (size_t)n * p.maxlines;
  const size_t bytes = al16((uint32_t)(n * sizeof(void *))) + al16((uint32_t)(n * 4)) + 4 * al16((uint32_t)(nl * 4)) +
                       2 * al16((uint32_t)(n * 4));
  if (!grow_device(&cx->d_scratch, &cx->d_scratch_cap, bytes)) return acb200_last_error();
  cx->scratch_dirty = true; // the render path must re-zero its look-back area before reusing this buffer
  std::vector<uint8_t> h(al16((uint32_t)(n * sizeof(void *))) + al16((uint32_t)(n * 4)));
  for (int i = 0; i < n; i++) {
    reinterpret_cast<const uint8_t **>(h.data())[i] = d_srcs[i];
    reinterpret_cast<uint32_t *>(h.data() + al16((uint32_t)(n * sizeof(void *))))[i] = d_srcs[i] ? (uint32_t)sizes[i] : 0u;
  }
  uint8_t *b = cx->d_scratch;
  ACB_CUDA(cudaMemcpyAsync(b, h.data(), h.size(), cudaMemcpyHostToDevice, st));
  ACB_CUDA(cudaStreamSynchronize(st)); // h is a stack-lifetime buffer
  p.src = reinterpret_cast<const uint8_t *const *>(b);
  b += al16((uint32_t)(n * sizeof(void *)));
  p.size = reinterpret_cast<const uint32_t *>(b);
  b += al16((uint32_t)(n * 4));
  p.lstart = reinterpret_cast<uint32_t *>(b);
  b += al16((uint32_t)(nl * 4));
  p.llen = reinterpret_cast<uint32_t *>(b);
  b += al16((uint32_t)(nl * 4));
  p.clen = reinterpret_cast<uint32_t *>(b);
  b += al16((uint32_t)(nl * 4));
  p.dcol = reinterpret_cast<uint32_t *>(b);
  b += al16((uint32_t)(nl * 4));
  p.nlines = reinterpret_cast<uint32_t *>(b);
  b += al16((uint32_t)(n * 4));
  p.nl_total = reinterpret_cast<uint32_t *>(b);
  k_text_grid<<<1, 256, 0, st>>>(p);
  ACB_CUDA(cudaGetLastError());
  count_launch();
  *out_size = total - 1;
  return E_OK;
}

} // namespace acb

using namespace acb;

extern "C" {

int acb200_create_grid_device(const uint8_t *const *d_srcs, const size_t *sizes, int n, int width, int height,
                              uint8_t *d_out, size_t *out_size, void *stream) {
  if (!d_srcs || !sizes || n <= 0 || width <= 0 || height <= 0 || !d_out || !out_size)
    return set_error(E_INVALID_PARAM, "acb200_create_grid_device: bad argument");
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  return text_grid_device(d_srcs, sizes, n, width, height, d_out, out_size, stream ? (cudaStream_t)stream : cx->stream,
                          cx);
}

// lib/video/ascii/ascii.c:602-885
char *ascii_create_grid(ascii_frame_source_t *sources, int source_count, int width, int height, size_t *out_size) {
  if (!sources || source_count <= 0 || width <= 0 || height <= 0 || !out_size) return nullptr; // :603
  const size_t w = (size_t)width, h = (size_t)height;
  if (w > SIZE_MAX / h || w * h > (size_t)1 << 30) { // :616-631, 798-801 (and a sane device bound)
    set_error(E_INVALID_PARAM, "ascii_create_grid: dimensions would overflow: %dx%d", width, height);
    return nullptr;
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  const int n = source_count;
  size_t in_total = 0, max_src = 0;
  for (int i = 0; i < n; i++) {
    size_t z = sources[i].frame_data ? sources[i].frame_size : 0;
    in_total += (z + 15) & ~(size_t)15;
    if (i == 0) max_src = z;
  }
  const size_t canvas = w * h + h + 1;
  const size_t out_cap = (canvas > max_src + 1 ? canvas : max_src + 1) + 16;
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, in_total + 16) || !grow_device(&cx->d_in, &cx->d_in_cap, in_total + 16) ||
      !grow_device(&cx->d_out, &cx->d_out_cap, out_cap) || !grow_pinned(&cx->h_out, &cx->h_out_cap, out_cap))
    return nullptr;
  std::vector<const uint8_t *> dptr(n);
  std::vector<size_t> sizes(n);
  size_t o = 0;
  for (int i = 0; i < n; i++) {
    size_t z = sources[i].frame_data ? sources[i].frame_size : 0;
    if (z) memcpy(cx->h_in + o, sources[i].frame_data, z);
    dptr[i] = z ? cx->d_in + o : nullptr;
    sizes[i] = z;
    o += (z + 15) & ~(size_t)15;
  }
  size_t res = 0;
  if ((in_total && cudaMemcpyAsync(cx->d_in, cx->h_in, in_total, cudaMemcpyHostToDevice, cx->stream) != cudaSuccess) ||
      text_grid_device(dptr.data(), sizes.data(), n, width, height, cx->d_out, &res, cx->stream, cx) != E_OK ||
      cudaMemcpyAsync(cx->h_out, cx->d_out, res + 1, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      cudaStreamSynchronize(cx->stream) != cudaSuccess) {
    if (!acb200_last_error()) set_error(E_INVALID_STATE, "ascii_create_grid: CUDA failure");
    return nullptr;
  }
  char *r = (char *)user_alloc(res + 1);
  if (!r) return nullptr;
  memcpy(r, cx->h_out, res + 1);
  r[res] = '\0';
  // single-source path reports the canvas size (ascii.c:650,705); the multi-source path reports strlen (ascii.c:883)
  *out_size = (n == 1) ? res : strlen(r);
  return r;
}

// calculate_optimal_grid_layout — src/server/stream.c:523-651 (host float)
void acb200_grid_layout(const int *ws, const int *hs, int n, int term_w, int term_h, int *out_cols, int *out_rows) {
  if (n == 0) {
    *out_cols = 0;
    *out_rows = 0;
    return;
  }
  if (n == 1) {
    *out_cols = 1;
    *out_rows = 1;
    return;
  }
  const float CHAR_ASPECT = 2.0f;
  float avg_aspect = 0.0f;
  for (int i = 0; i < n; i++) avg_aspect += (float)ws[i] / (float)hs[i];
  avg_aspect /= n;
  int best_cols = 1, best_rows = n;
  float best_utilization = 0.0f;
  for (int cols = 1; cols <= n; cols++) {
    int rows = (n + cols - 1) / cols;
    if (cols * rows - n > cols) continue;
    int cell_width = term_w / cols, cell_height = term_h / rows;
    if (cell_width < 20 || cell_height < 10) continue;
    float total_area_used = 0.0f;
    int cell_area = cell_width * cell_height;
    for (int i = 0; i < n; i++) {
      float cell_visual_aspect = (float)cell_width / ((float)cell_height * CHAR_ASPECT);
      int fitted_width, fitted_height;
      if (avg_aspect > cell_visual_aspect) {
        fitted_width = cell_width;
        fitted_height = (int)((cell_width / avg_aspect) / CHAR_ASPECT);
      } else {
        fitted_height = cell_height;
        fitted_width = (int)(cell_height * CHAR_ASPECT * avg_aspect);
      }
      if (fitted_width > cell_width) fitted_width = cell_width;
      if (fitted_height > cell_height) fitted_height = cell_height;
      total_area_used += fitted_width * fitted_height;
    }
    float utilization = total_area_used / (float)(cell_area * n);
    if (utilization > best_utilization) {
      best_utilization = utilization;
      best_cols = cols;
      best_rows = rows;
    }
  }
  *out_cols = best_cols;
  *out_rows = best_rows;
}

// create_multi_source_composite — src/server/stream.c:664-779
int acb200_composite_host(const uint8_t *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                          uint8_t *out_rgb, int *out_cols, int *out_rows) {
  if (!srcs || !ws || !hs || n <= 0 || width <= 0 || height <= 0 || !out_rgb)
    return set_error(E_INVALID_PARAM, "acb200_composite_host: bad argument");
  for (int i = 0; i < n; i++)
    if (!srcs[i] || ws[i] <= 0 || hs[i] <= 0) return set_error(E_INVALID_PARAM, "source %d invalid", i);
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  int gc, gr;
  acb200_grid_layout(ws, hs, n, width, height, &gc, &gr);
  if (out_cols) *out_cols = gc;
  if (out_rows) *out_rows = gr;
  const int CW = width, CH = height * 2; // 1 char = 1 px wide, 2 px tall (stream.c:677-679)
  const size_t comp_bytes = (size_t)CW * CH * 3;
  const int cellw = CW / gc, cellh = CH / gr;
  // per source: which rows/geometry; NN touches th of hs[i] rows -> move only those
  struct Job {
    int tw, th, x0, y0;
    size_t in_off, in_bytes;
    int gather;
  };
  std::vector<Job> jobs;
  size_t in_total = 0;
  for (int i = 0, v = 0; i < n && v < 9; i++, v++) { // max 9 sources (stream.c:687)
    const int row = v / gc, col = v % gc;
    float src_aspect = (float)ws[i] / (float)hs[i];
    float cell_visual_aspect = (float)cellw / (float)cellh;
    int tw, th;
    if (src_aspect > cell_visual_aspect) { // stream.c:708-716
      tw = cellw;
      th = (int)((cellw / src_aspect) + 0.5f);
    } else {
      th = cellh;
      tw = (int)((cellh * src_aspect) + 0.5f);
    }
    Job j{tw, th, col * cellw, row * cellh, in_total, 0, 0};
    if (tw > 0 && th > 0) {
      j.gather = th < hs[i];
      j.in_bytes = (size_t)ws[i] * 3 * (j.gather ? th : hs[i]);
      in_total += (j.in_bytes + 15) & ~(size_t)15;
    }
    jobs.push_back(j);
  }
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, in_total + 16) || !grow_device(&cx->d_in, &cx->d_in_cap, in_total + 16) ||
      !grow_device(&cx->d_out, &cx->d_out_cap, comp_bytes) || !grow_pinned(&cx->h_out, &cx->h_out_cap, comp_bytes))
    return acb200_last_error();
  for (size_t k = 0; k < jobs.size(); k++) {
    const Job &j = jobs[k];
    if (!j.in_bytes) continue;
    const size_t R = (size_t)ws[k] * 3;
    if (j.gather) {
      const uint32_t yr = (uint32_t)((((uint64_t)hs[k] << 16) / (uint64_t)j.th) + 1);
      for (int y = 0; y < j.th; y++) {
        uint32_t sy = ((uint32_t)y * yr) >> 16;
        if (sy >= (uint32_t)hs[k]) sy = (uint32_t)hs[k] - 1;
        memcpy(cx->h_in + j.in_off + (size_t)y * R, srcs[k] + (size_t)sy * R, R);
      }
    } else {
      memcpy(cx->h_in + j.in_off, srcs[k], j.in_bytes);
    }
  }
  ACB_CUDA(cudaMemcpyAsync(cx->d_in, cx->h_in, in_total, cudaMemcpyHostToDevice, cx->stream));
  ACB_CUDA(cudaMemsetAsync(cx->d_out, 0, comp_bytes, cx->stream)); // image_clear, stream.c:683
  for (size_t k = 0; k < jobs.size(); k++) {
    const Job &j = jobs[k];
    if (!j.in_bytes) continue;
    // a gathered source has exactly th rows: sampling it with src_h = th is the identity in y
    ACB_CUDA(launch_composite_cell(cx->d_in + j.in_off, ws[k], j.gather ? j.th : hs[k], cx->d_out, CW, CH, j.tw, j.th,
                                   j.x0, j.y0, cellw, cellh, cx->stream));
    count_launch();
  }
  ACB_CUDA(cudaMemcpyAsync(cx->h_out, cx->d_out, comp_bytes, cudaMemcpyDeviceToHost, cx->stream));
  ACB_CUDA(cudaStreamSynchronize(cx->stream));
  memcpy(out_rgb, cx->h_out, comp_bytes);
  return E_OK;
}

} // extern "C"
