// grid.cu — the two grid compositors of the reference, on the device.
//
//  * text-space: ascii_create_grid (lib/video/ascii/ascii.c:602-885) — N rendered frames ->
//    one W x H character canvas with '|', '_', '+' separators.  Layout scoring is host float
//    (ceil/logf/fabsf, ascii.c:712-769); everything that touches frame bytes runs in k_grid_lines / k_grid_place.
//  * pixel-space: create_multi_source_composite (src/server/stream.c:664-779) with
//    calculate_optimal_grid_layout (stream.c:523-651) — N RGB sources -> one W x 2H composite;
//    the NN resize + clipped blit is k_composite_cell per host source here, k_composite_all for the server's
//    resident sources (server.cu); both kernels live in render_kernels.cu.
#include <cmath>
#include <cstring>
#include <vector>

#include "engine.h"

namespace acb {

// ------------------------------------------------------------------ text grid kernel
constexpr int TG_INLINE = 32; // sources whose pointer/size tables travel in the kernel parameters (MAX_CLIENTS)
struct TextGridParams {
  const uint8_t *const *src; // [n] device pointers    } device tables, used when n > TG_INLINE
  const uint32_t *size;      // [n]                     }
  const uint8_t *src_in[TG_INLINE]; // the same tables inline (n <= TG_INLINE): no upload, no lifetime to manage
  uint32_t size_in[TG_INLINE];
  const uint32_t *size_dev;  // optional: lengths still on the device (written by the render kernels), + size_bias each
  uint32_t size_bias;
  int n, W, H;
  int gcols, grows, cw, ch;  // layout (multi-source path)
  int single;                // n == 1: centre the lone frame (ascii.c:610-707)
  int maxlines;              // lines kept per source: ch (multi) or H (single)
  uint8_t *out;
  uint32_t total;            // W*H + H + 1
  uint32_t *lstart, *llen, *clen, *dcol; // [n * maxlines] scratch
  uint32_t *nlines, *nl_total;           // [n]
  uint32_t *order;                       // [total] write-order id that owns each canvas byte (0 = blank canvas)
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

// The grid is built by three launches, none of them serial in the number of sources or lines:
//   k_grid_lines  (one CTA per source): newline table by a two-pass chunked scan, then per line the ANSI-aware
//                 truncation and the copy decision; the CTAs also blank the canvas and clear the order map.
//   k_grid_claim  (one CTA per line / separator set): every byte a writer would store takes part in an atomicMax on a
//                 per-canvas-byte ORDER id — the reference writes source-major, line by line, then that source's
//                 separators (ascii.c:829-880), and where ANSI bytes spill past a cell the LATER write wins.
//   k_grid_write  (same enumeration): a writer stores its byte only where it holds the maximum.
constexpr int TG_NT = 256;

__device__ __forceinline__ const uint8_t *tg_src(const TextGridParams &p, int s) {
  return p.n <= TG_INLINE ? p.src_in[s] : p.src[s];
}
__device__ __forceinline__ uint32_t tg_size(const TextGridParams &p, int s) {
  if (p.size_dev) return p.size_dev[s] + p.size_bias;
  return p.n <= TG_INLINE ? p.size_in[s] : p.size[s];
}
__device__ __forceinline__ uint32_t tg_line_id(const TextGridParams &p, int s, int r) { // 1-based, reference order
  return (uint32_t)s * (uint32_t)(p.maxlines + 4) + (uint32_t)r + 1u;
}

__global__ void __launch_bounds__(TG_NT) k_grid_lines(const TextGridParams p) {
  __shared__ uint32_t s_cnt[TG_NT], s_tot;
  const int tid = threadIdx.x, s = blockIdx.x;
  const uint32_t W1 = (uint32_t)p.W + 1u;
  // blank canvas, '\n' closing every row, NUL (ascii.c:633-640, 806-813); order map = 0 ("blank")
  for (uint32_t i = blockIdx.x * TG_NT + tid; i < p.total; i += gridDim.x * TG_NT) {
    p.out[i] = i == p.total - 1 ? 0 : (i % W1 == (uint32_t)p.W) ? '\n' : ' ';
    p.order[i] = 0u;
  }
  const uint8_t *d = tg_src(p, s);
  const uint32_t size = d ? tg_size(p, s) : 0u;
  uint32_t *lstart = p.lstart + (size_t)s * p.maxlines, *llen = p.llen + (size_t)s * p.maxlines;
  const uint32_t ML = (uint32_t)p.maxlines;

  // pass 1: newlines per chunk
  const uint32_t chunk = (size + TG_NT - 1) / TG_NT;
  const uint32_t b0 = min((uint32_t)tid * chunk, size), b1 = min(b0 + chunk, size);
  uint32_t cnt = 0;
  for (uint32_t i = b0; i < b1; i++) cnt += d[i] == '\n';
  s_cnt[tid] = cnt;
  __syncthreads();
  if (tid < 32) { // exclusive scan of 256 counts by one warp (8 per lane)
    uint32_t v[8], sum = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      v[k] = s_cnt[tid * 8 + k];
      sum += v[k];
    }
    uint32_t inc = sum;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      uint32_t o = __shfl_up_sync(0xffffffffu, inc, dd);
      if (tid >= dd) inc += o;
    }
    uint32_t run = inc - sum;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      s_cnt[tid * 8 + k] = run;
      run += v[k];
    }
    if (tid == 31) s_tot = inc;
  }
  __syncthreads();
  const uint32_t nl_total = s_tot;
  // pass 2: line L ends at its newline; line L+1 starts right after it (only the first maxlines lines are kept)
  if (tid == 0 && ML > 0) lstart[0] = 0u;
  uint32_t L = s_cnt[tid];
  for (uint32_t i = b0; i < b1 && L < ML; i++)
    if (d[i] == '\n') {
      llen[L] = i; // end offset for now
      if (L + 1 < ML) lstart[L + 1] = i + 1;
      L++;
    }
  __syncthreads();
  // lines closed by a newline, plus a last line without one (if there is room)
  uint32_t nlines = min(nl_total, ML);
  const bool tail = nl_total < ML && lstart[nl_total] < size; // lstart[nl_total] = byte after the last newline (0 if none)
  for (uint32_t r = tid; r < nlines; r += TG_NT) llen[r] = llen[r] - lstart[r];
  if (tid == 0 && tail) llen[nl_total] = size - lstart[nl_total];
  if (tail) nlines++;
  if (tid == 0) {
    p.nlines[s] = nlines;
    p.nl_total[s] = nl_total;
  }
  __syncthreads();

  // per line: ANSI-aware truncation and the copy decision
  for (uint32_t r = tid; r < nlines; r += TG_NT) {
    const size_t it = (size_t)s * p.maxlines + r;
    const uint8_t *line = d + lstart[r];
    const int ll = (int)llen[r];
    uint32_t cl = 0, col = 0;
    if (p.single) { // ascii.c:661-703
      int vpad = (p.H - (int)nl_total) / 2;
      if (vpad < 0) vpad = 0;
      const int row = vpad + (int)r;
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
      if ((int)r < p.ch && r0 + (int)r < p.H) {
        int tv;
        int c = truncate_visible(line, ll, p.cw, &tv);
        const size_t pos = (size_t)(r0 + (int)r) * W1 + (size_t)c0;
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
}

// One CTA per writer: blockIdx.x < n*maxlines -> line (s, r); the n CTAs after that -> the separators of source s.
// WRITE = false: claim the bytes (atomicMax of the order id); WRITE = true: store where the claim was won.
// The canvas terminator (total - 1) is never claimed: a spill that reaches it is undone by the reference's callers
// reading a C string; here it simply stays NUL.
template <bool WRITE> __global__ void __launch_bounds__(128) k_grid_place(const TextGridParams p) {
  const int tid = threadIdx.x;
  const uint32_t W1 = (uint32_t)p.W + 1u, last = p.total - 1u;
  const int items = p.n * p.maxlines;
  auto put = [&](size_t idx, uint8_t c, uint32_t id) {
    if (idx >= (size_t)last) return;
    if (!WRITE) atomicMax(&p.order[idx], id);
    else if (p.order[idx] == id) p.out[idx] = c;
  };
  if ((int)blockIdx.x < items) {
    const int s = blockIdx.x / p.maxlines, r = blockIdx.x % p.maxlines;
    if ((uint32_t)r >= p.nlines[s]) return;
    const uint32_t cl = p.clen[blockIdx.x];
    if (!cl) return;
    const uint8_t *line = tg_src(p, s) + p.lstart[blockIdx.x];
    const size_t pos = p.dcol[blockIdx.x];
    const uint32_t id = tg_line_id(p, s, r);
    for (uint32_t i = tid; i < cl; i += 128) put(pos + i, WRITE ? line[i] : 0, id);
    return;
  }
  if (p.single) return;
  const int s = (int)blockIdx.x - items;
  const int gr = s / p.gcols, gc = s % p.gcols;
  const int r0 = gr * (p.ch + 1), c0 = gc * (p.cw + 1);
  const uint32_t base = (uint32_t)s * (uint32_t)(p.maxlines + 4) + (uint32_t)p.maxlines;
  if (gc < p.gcols - 1 && c0 + p.cw < p.W) // vertical separator, ascii.c:855-863
    for (int row = r0 + tid; row < r0 + p.ch && row < p.H; row += 128)
      put((size_t)row * W1 + (size_t)(c0 + p.cw), '|', base + 1u);
  if (gr < p.grows - 1 && r0 + p.ch < p.H) { // horizontal separator + corner, ascii.c:865-880
    for (int c = c0 + tid; c < c0 + p.cw && c < p.W; c += 128) put((size_t)(r0 + p.ch) * W1 + (size_t)c, '_', base + 2u);
    if (tid == 0 && gc < p.gcols - 1 && c0 + p.cw < p.W) put((size_t)(r0 + p.ch) * W1 + (size_t)(c0 + p.cw), '+', base + 3u);
  }
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

// sizes: host lengths — or nullptr with d_sizes = device-resident lengths (each + size_bias), which saves the read-back
// between the render and the grid (acb200_grid_frame)
int text_grid_device(const uint8_t *const *d_srcs, const size_t *sizes, int n, int width, int height, uint8_t *d_out,
                     size_t *out_size, cudaStream_t st, ThreadCtx *cx, const uint32_t *d_sizes, uint32_t size_bias,
                     bool *size_is_exact) {
  if (size_is_exact) *size_is_exact = n == 1; // ascii.c:650,705: the single-source path reports the canvas size
  const uint32_t total = (uint32_t)((size_t)width * height + height + 1);
  if (d_sizes && n > TG_INLINE) return set_error(E_INVALID_PARAM, "device-resident sizes: at most %d sources", TG_INLINE);
  TextGridParams p{};
  p.size_dev = d_sizes;
  p.size_bias = size_bias;
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
      size_t z0 = sizes ? sizes[0] : 0;
      if (d_sizes) { // rare path: fetch the one length that decides the copy
        uint32_t v = 0;
        ACB_CUDA(cudaMemcpyAsync(&v, d_sizes, sizeof(v), cudaMemcpyDeviceToHost, st));
        ACB_CUDA(cudaStreamSynchronize(st));
        z0 = (size_t)v + size_bias;
      }
      if (d_srcs[0] && z0 > 0) ACB_CUDA(cudaMemcpyAsync(d_out, d_srcs[0], z0, cudaMemcpyDeviceToDevice, st));
      const size_t z = (d_srcs[0] && z0 > 0) ? z0 : 0;
      ACB_CUDA(cudaMemsetAsync(d_out + z, 0, 1, st));
      *out_size = z;
      if (size_is_exact) *size_is_exact = true; // ascii.c:785: frame_size as given
      return E_OK;
    }
    p.maxlines = p.ch;
  }
  // parameter block + scratch in one device allocation of the thread context
  const size_t nl = (size_t)n * p.maxlines;
  const size_t head = al16((uint32_t)(n * sizeof(void *))) + al16((uint32_t)(n * 4));
  const size_t order_bytes = ((size_t)total * 4 + 15) & ~(size_t)15; // 64-bit: total may be close to 2^30
  const size_t bytes = head + 4 * al16((uint32_t)(nl * 4)) + 2 * al16((uint32_t)(n * 4)) + order_bytes;
  // the tables live in this thread's scratch: drain whatever another (caller-owned) stream still has queued on it, and
  // remember `st` so that the render path does the same before it re-zeroes the buffer on the internal stream
  if (sync_foreign(cx, st) != E_OK) return acb200_last_error();
  if (st != cx->stream) ACB_CUDA(cudaStreamSynchronize(cx->stream));
  if (!grow_device(&cx->d_scratch, &cx->d_scratch_cap, bytes)) return acb200_last_error();
  cx->scratch_dirty = true; // the render path must re-zero its look-back area before reusing this buffer
  uint8_t *b = cx->d_scratch;
  if (n <= TG_INLINE) { // tables ride in the kernel parameters
    for (int i = 0; i < n; i++) {
      p.src_in[i] = d_srcs[i];
      p.size_in[i] = (d_srcs[i] && sizes) ? (uint32_t)sizes[i] : 0u;
    }
  } else {
    std::vector<uint8_t> h(head);
    for (int i = 0; i < n; i++) {
      reinterpret_cast<const uint8_t **>(h.data())[i] = d_srcs[i];
      reinterpret_cast<uint32_t *>(h.data() + al16((uint32_t)(n * sizeof(void *))))[i] = d_srcs[i] ? (uint32_t)sizes[i] : 0u;
    }
    ACB_CUDA(cudaMemcpyAsync(b, h.data(), head, cudaMemcpyHostToDevice, st));
    ACB_CUDA(cudaStreamSynchronize(st)); // h is a stack-lifetime buffer
  }
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
  b += al16((uint32_t)(n * 4));
  p.order = reinterpret_cast<uint32_t *>(b);
  const unsigned writers = (unsigned)(nl + (size_t)n);
  k_grid_lines<<<(unsigned)n, TG_NT, 0, st>>>(p);
  k_grid_place<false><<<writers, 128, 0, st>>>(p);
  k_grid_place<true><<<writers, 128, 0, st>>>(p);
  ACB_CUDA(cudaGetLastError());
  count_launch(3);
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
  if ((size_t)width * (size_t)height > (size_t)1 << 30) // same bound as ascii_create_grid (ascii.c:616-631 + device sanity)
    return set_error(E_INVALID_PARAM, "acb200_create_grid_device: dimensions would overflow: %dx%d", width, height);
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  return text_grid_device(d_srcs, sizes, n, width, height, d_out, out_size, stream ? (cudaStream_t)stream : cx->stream,
                          cx, nullptr, 0, nullptr);
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
  bool exact = false;
  if ((in_total && cudaMemcpyAsync(cx->d_in, cx->h_in, in_total, cudaMemcpyHostToDevice, cx->stream) != cudaSuccess) ||
      text_grid_device(dptr.data(), sizes.data(), n, width, height, cx->d_out, &res, cx->stream, cx, nullptr, 0, &exact) != E_OK ||
      cudaMemcpyAsync(cx->h_out, cx->d_out, res + 1, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      wait_stream(cx) != E_OK) {
    if (!acb200_last_error()) set_error(E_INVALID_STATE, "ascii_create_grid: CUDA failure");
    return nullptr;
  }
  char *r = (char *)user_alloc(res + 1);
  if (!r) return nullptr;
  memcpy(r, cx->h_out, res + 1);
  r[res] = '\0';
  // single-source path reports the canvas size (ascii.c:650,705); the multi-source path reports strlen (ascii.c:883)
  *out_size = exact ? res : strlen(r);
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
  if (wait_stream(cx) != E_OK) return acb200_last_error();
  memcpy(out_rgb, cx->h_out, comp_bytes);
  return E_OK;
}

} // extern "C"
