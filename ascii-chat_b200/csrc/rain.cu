// rain.cu — digital rain on a finished frame string (second half):
// lib/video/anim/digital_rain.c (digital_rain_init / _apply / _reset / setters), called from the client display path
// (src/common/session/display.c:657-671) on the string ascii_convert_with_capabilities returned.
//
// What runs where.  The effect is a brightness field b(col,row,t) = 1 - fract(wobble((offset[col] + t*fall*speed[col] -
// row) / len)) with wobble(x) = x + 0.3 sinf(sqrt2 x) + 0.2 sinf(sqrt5 x) (digital_rain.c:35-46, 70-91).  Its values are
// truncated into 8-bit colour components, so matching the reference byte for byte means matching ITS sinf bit for bit,
// and that is the host's libm (whichever variant the box's glibc selects).  The field is therefore evaluated on the
// host, with libm, in the reference's order of operations — one table of (lines + 2) x columns floats per frame, the
// same standing as the aspect fit and the rainbow hue (the only other float on the path).  Everything that touches the
// string runs on the device: tokenising it (SGR colour sequences, other escape sequences, newlines, UTF-8 characters),
// the per-cell low-pass filter against the state kept in device memory (every visit of a cell filters again,
// digital_rain.c:421-430 — the order of visits is part of the result), scaling the colours (IEEE single-precision
// multiply + truncation, no FMA contraction), and writing the result, which is 1.5-3x the input.
//
// The reference walks the string serially with a (col,row) cursor.  Lines are independent of each other as long as no
// escape sequence swallows a newline, so the device form is one thread per line, each running the reference's loop over
// its own byte range (pass 1 counts the output bytes, a scan places the lines, pass 2 writes); a string in which a CSI
// sequence does run across a '\n' is detected in pass 1 and re-run as a single range by one thread — slow, exact.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "engine.h"

using namespace acb;

namespace {

struct RainImpl {
  digital_rain_t pub; // first member: the pointer handed to the caller
  int device = 0;     // CUDA ordinal previous_brightness lives on
  float *alt = nullptr; // the other state buffer: a frame reads pub.previous_brightness and writes this one, then they swap
};

constexpr int RN_NT = 256, RN_PER = 16, RN_CHUNK = RN_NT * RN_PER;
constexpr int RN_LPB = 4;          // lines per block in the walk kernels
constexpr uint32_t RN_STAGE = 12288; // bytes of a line staged in shared memory; the rest is read from L2
constexpr int RN_SCOLS = 512;        // columns of the line's brightness rows (this row, the row below, the state) staged too
constexpr size_t RN_LINE_SMEM = RN_STAGE + 3u * RN_SCOLS * sizeof(float);

struct RainParams {
  const uint8_t *in;
  uint32_t n;
  uint8_t *out;
  uint32_t *chunk_nl;    // [chunks]
  uint32_t *line_start;  // [nlines]
  uint32_t *line_len;    // [nlines] output bytes of the line
  uint32_t *line_off;    // [nlines] exclusive scan of line_len
  uint32_t *result;      // [0] total output bytes, [1] complex flag
  uint32_t nlines;
  const float *target;   // [rows_tab][cols] un-filtered brightness of the frame
  int cols, rows, rows_tab;
  const float *prev;     // [rows][cols] filtered brightness of the frame before (read only)
  float *prev_new;       // same grid, this frame's values (the two buffers swap roles every frame)
  uint32_t *lane_off;    // [nlines][32] output offset of every lane's segment inside its line
  float decay;
  int first_frame;
  uint32_t rain_rgb;     // 0x00RRGGBB
};

__global__ void __launch_bounds__(RN_NT) k_rain_nl_count(const RainParams p) {
  __shared__ uint32_t s_red[RN_NT / 32];
  const uint32_t b0 = blockIdx.x * RN_CHUNK + threadIdx.x * RN_PER;
  uint32_t c = 0;
  for (uint32_t i = b0; i < b0 + RN_PER && i < p.n; i++) c += p.in[i] == '\n';
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < RN_NT / 32; k++) t += s_red[k];
    p.chunk_nl[blockIdx.x] = t;
  }
}

// line k (k >= 1) starts behind the k-th newline; line 0 starts at byte 0
__global__ void __launch_bounds__(RN_NT) k_rain_line_starts(const RainParams p) {
  __shared__ uint32_t s_w[RN_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint32_t base = 0;
  for (int k = tid; k < (int)blockIdx.x; k += RN_NT) base += p.chunk_nl[k];
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
  if (lane == 0) s_w[warp] = base;
  __syncthreads();
  base = 0;
  for (int k = 0; k < RN_NT / 32; k++) base += s_w[k];
  __syncthreads();
  const uint32_t b0 = blockIdx.x * RN_CHUNK + tid * RN_PER;
  uint32_t c = 0;
  for (uint32_t i = b0; i < b0 + RN_PER && i < p.n; i++) c += p.in[i] == '\n';
  uint32_t inc = c;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_w[warp] = inc;
  __syncthreads();
  uint32_t pre = 0;
  for (int k = 0; k < warp; k++) pre += s_w[k];
  uint32_t rank = base + pre + inc - c; // newlines before this thread's first byte
  for (uint32_t i = b0; i < b0 + RN_PER && i < p.n; i++)
    if (p.in[i] == '\n') p.line_start[++rank] = i + 1;
  if (blockIdx.x == 0 && tid == 0) p.line_start[0] = 0;
}

struct CountOut {
  uint32_t n = 0;
  __device__ __forceinline__ void put(uint8_t) { ++n; }
};
struct ByteOut {
  uint8_t *p;
  __device__ __forceinline__ void put(uint8_t c) { *p++ = c; }
};
template <class O> __device__ __forceinline__ void put_dec(O &o, uint32_t v) { // %d of 0..255
  if (v >= 100u) o.put((uint8_t)('0' + v / 100u));
  if (v >= 10u) o.put((uint8_t)('0' + (v / 10u) % 10u));
  o.put((uint8_t)('0' + v % 10u));
}
// generate_modulated_color, digital_rain.c:326-364
template <class O>
__device__ __forceinline__ void put_scaled(O &o, bool fg, int r, int g, int b, float bright, bool cursor) {
  if (cursor) bright = __fmul_rn(bright, 2.0f);
  if (bright < 0.0f) bright = 0.0f;
  if (bright > 1.0f) bright = 1.0f;
  int v[3] = {__float2int_rz(__fmul_rn((float)r, bright)), __float2int_rz(__fmul_rn((float)g, bright)),
              __float2int_rz(__fmul_rn((float)b, bright))};
  o.put(0x1b);
  o.put('[');
  o.put(fg ? '3' : '4');
  o.put('8');
  o.put(';');
  o.put('2');
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.put(';');
    put_dec(o, (uint32_t)(v[k] < 0 ? 0 : v[k] > 255 ? 255 : v[k]));
  }
  o.put('m');
}

// ---- tokens --------------------------------------------------------------------------------------------------
// The reference's loop (digital_rain.c:405-502) is a tokeniser with three byte-level states — N (normal), E (right after
// an ESC), C (inside ESC [ ... up to a byte in @..~) — and four kinds of token: a truecolor SGR (re-emitted scaled: a
// VISIT of the cursor's cell), any other escape sequence (copied), '\n' (row++, col = 0), a visible UTF-8 character
// (a visit, the rain colour in front, col++).  Whether an ESC [ sequence is a colour sequence is decided where it
// starts; both kinds end at the first byte in @..~ (a colour sequence holds only digits and ';' before its 'm'), so
// the state machine does not depend on the kind — which is what lets a warp cut a line into 32 segments: every lane
// computes how its segment maps each entry state to an exit state, the maps are chained, and every lane then walks the
// tokens that START in its segment from the right state.
enum { RS_N = 0, RS_E = 1, RS_C = 2 };
__device__ __forceinline__ int rain_step(int st, uint32_t c) {
  if (st == RS_C) return (c >= '@' && c <= '~') ? RS_N : RS_C;
  if (c == 0x1b) return RS_E;
  return (st == RS_E && c == '[') ? RS_C : RS_N;
}

// Walks the tokens that start in [from, to) (a token may run past `to`; the owner of its first byte handles all of it).
// entry = state in front of byte `from`.  Sink callbacks: colour(fg, r, g, b), raw(i, j), newline(), glyph(i, len).
template <class At, class Sink>
__device__ __forceinline__ void rain_tokens(const At &at, uint32_t n, uint32_t from, uint32_t to, uint32_t line_end, int entry,
                                            Sink &sk, uint32_t *complex) {
  uint32_t i = from;
  if (entry == RS_E && i < to && at(i) == '[') { // the CSI was opened by the ESC in front of this segment
    entry = RS_C;
    i++;
  }
  if (entry == RS_C) { // inside somebody else's sequence: it ends behind the first byte in @..~
    while (i < n && !(at(i) >= '@' && at(i) <= '~')) i++;
    if (i < n) i++;
  }
  while (i < to) {
    const uint32_t c = at(i);
    if (c == 0x1b) {
      // parse_ansi_color (:240-301): ESC [ (38|48) ;2; R ; G ; B m
      bool colour = false, fg = false;
      int rgb[3] = {0, 0, 0};
      uint32_t j = i;
      if (at(i + 1) == '[') {
        const uint32_t a = at(i + 2), b = at(i + 3);
        if ((a == '3' || a == '4') && b == '8' && at(i + 4) == ';' && at(i + 5) == '2' && at(i + 6) == ';') {
          fg = a == '3';
          j = i + 7;
          colour = true;
          for (int k = 0; k < 3 && colour; k++) {
            uint32_t v = 0;
            while (at(j) >= '0' && at(j) <= '9') v = v * 10u + (at(j++) - '0');
            rgb[k] = (int)v;
            if (at(j) != (k < 2 ? ';' : 'm')) colour = false;
            else j++;
          }
        }
      }
      if (colour) {
        sk.colour(fg, rgb[0], rgb[1], rgb[2]);
        i = j;
      } else { // skip_ansi_sequence (:306-324): copied as it is
        j = i + 1;
        if (at(j) == '[') {
          j++;
          while (j < n && !(at(j) >= '@' && at(j) <= '~')) j++;
          if (j < n) j++;
        }
        if (j > line_end) *complex = 1u; // the sequence swallowed this line's newline
        sk.raw(i, j);
        i = j;
      }
    } else if (c == '\n') {
      sk.newline();
      i++;
    } else { // a visible character; utf8_decode's length rule (lib/util/utf8.c:18-44): an invalid sequence is one byte
      int len = c < 0x80u ? 1 : (c & 0xE0u) == 0xC0u ? 2 : (c & 0xF0u) == 0xE0u ? 3 : (c & 0xF8u) == 0xF0u ? 4 : 1;
      for (int k = 1; k < len; k++)
        if ((at(i + k) & 0xC0u) != 0x80u) {
          len = 1;
          break;
        }
      sk.glyph(i, len);
      i += len;
    }
  }
}

// structure only: how many characters start here, how many colour visits before the first / after the last of them
struct StructSink {
  int nchar = 0, v_head = 0, v_tail = 0;
  __device__ __forceinline__ void colour(bool, int, int, int) {
    if (nchar == 0) v_head++;
    v_tail++;
  }
  __device__ __forceinline__ void raw(uint32_t, uint32_t) {}
  __device__ __forceinline__ void newline() {}
  __device__ __forceinline__ void glyph(uint32_t, int) {
    nchar++;
    v_tail = 0;
  }
};

// bytes (counted or written) + the brightness filter.  State is READ from p.prev (the frame before) and WRITTEN to
// p.prev_new, so the order in which lanes and lines run cannot matter; within a cell the visits chain through run_val.
template <class At, class O> struct EmitSink {
  const RainParams &p;
  const At &at;
  O &o;
  const float *frows; // staged rows (row0: brightness, row0 + 1: brightness below, state of row0), or nullptr
  int row0, col, row;
  bool write_state;   // the write pass
  bool every_visit;   // serial walk: every visit stores; warp walk: only a character's visit does (the last of its cell)
  int run_col = -1, run_row = -1;
  float run_val = 0.0f;
  int tail_visits = 0; // colour visits since the last character (their cell's final value is run_val)

  __device__ __forceinline__ float target(int c, int r) const { // get_rain_brightness: 0 beyond the last column (:71-73)
    if (c >= p.cols || r >= p.rows_tab) return 0.0f;
    if (frows && c < RN_SCOLS && (unsigned)(r - row0) < 2u) return frows[(r - row0) * RN_SCOLS + c];
    return p.target[(size_t)r * p.cols + c];
  }
  __device__ __forceinline__ float old_state(int c, int r) const {
    if (frows && r == row0 && c < RN_SCOLS) return frows[2 * RN_SCOLS + c];
    return p.prev[(size_t)r * p.cols + c];
  }
  // the cell already had `ord` visits in front of this walker's first token: replay the filter that often
  __device__ __forceinline__ void catch_up(int ord) {
    if (ord <= 0 || p.first_frame || !(row < p.rows && col < p.cols)) return;
    const float tg = target(col, row);
    float v = old_state(col, row);
    for (int k = 0; k < ord; k++) v = __fadd_rn(v, __fmul_rn(__fsub_rn(tg, v), p.decay));
    run_col = col, run_row = row, run_val = v;
  }
  __device__ __forceinline__ float visit(bool *cursor, bool is_char) { // :413-430
    float b = target(col, row);
    *cursor = b > target(col, row + 1);
    if (row < p.rows && col < p.cols) {
      if (!p.first_frame) {
        const float prev = (run_col == col && run_row == row) ? run_val : old_state(col, row);
        b = __fadd_rn(prev, __fmul_rn(__fsub_rn(b, prev), p.decay));
      }
      run_col = col, run_row = row, run_val = b;
      if (write_state && (every_visit || is_char)) p.prev_new[(size_t)row * p.cols + col] = b;
    }
    return b;
  }
  __device__ __forceinline__ void colour(bool fg, int r, int g, int b) {
    bool cursor;
    const float br = visit(&cursor, false);
    put_scaled(o, fg, r, g, b, br, cursor);
    tail_visits++;
  }
  __device__ __forceinline__ void raw(uint32_t i, uint32_t j) {
    for (uint32_t k = i; k < j; k++) o.put((uint8_t)at(k));
  }
  __device__ __forceinline__ void newline() {
    o.put('\n');
    if (!every_visit) flush_tail(); // colour visits between the last character and the newline: their cell ends here
    row++, col = 0, tail_visits = 0;
  }
  __device__ __forceinline__ void glyph(uint32_t i, int len) { // the rain colour in front of the character (:466-501)
    bool cursor;
    const float br = visit(&cursor, true);
    put_scaled(o, true, (int)((p.rain_rgb >> 16) & 255u), (int)((p.rain_rgb >> 8) & 255u), (int)(p.rain_rgb & 255u), br, cursor);
    for (int k = 0; k < len; k++) o.put((uint8_t)at(i + k));
    col++, tail_visits = 0;
  }
  // warp walk, end of the line's last visiting lane: colour visits behind the last character belong to a cell that
  // never gets one — their final value is the cell's new state
  __device__ __forceinline__ void flush_tail() {
    if (write_state && tail_visits > 0 && run_col == col && run_row == row && row < p.rows && col < p.cols)
      p.prev_new[(size_t)row * p.cols + col] = run_val;
  }
};

// One thread walks the range [i0, i1) serially (the fallback for strings in which an escape sequence swallows a
// newline: the whole string as one range).
template <bool WRITE> __global__ void __launch_bounds__(32) k_rain_serial(const RainParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const uint8_t *s = p.in;
  const uint32_t n = p.n;
  auto at = [&](uint32_t i) -> uint32_t { return i < n ? s[i] : 0u; };
  uint32_t complex = 0;
  if (WRITE) {
    ByteOut o{p.out};
    EmitSink<decltype(at), ByteOut> sk{p, at, o, nullptr, 0, 0, 0, true, true};
    rain_tokens(at, n, 0u, n, n, RS_N, sk, &complex);
  } else {
    CountOut o;
    EmitSink<decltype(at), CountOut> sk{p, at, o, nullptr, 0, 0, 0, false, true};
    rain_tokens(at, n, 0u, n, n, RS_N, sk, &complex);
    p.line_len[0] = o.n;
  }
}

// One WARP per line.  The line (and the three float rows its visits read) is staged in shared memory; lane l owns the
// tokens that start in its 1/32 of the line.
template <bool WRITE> __global__ void __launch_bounds__(32 * RN_LPB) k_rain_warp(const RainParams p) {
  extern __shared__ __align__(16) uint8_t s_stage[]; // RN_LPB x RN_LINE_SMEM
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t line = blockIdx.x * RN_LPB + wib;
  if (line >= p.nlines) return;
  uint8_t *stage = s_stage + (size_t)wib * RN_LINE_SMEM;
  float *fr = reinterpret_cast<float *>(stage + RN_STAGE);
  const uint32_t i0 = p.line_start[line], i1 = line + 1 < p.nlines ? p.line_start[line + 1] : p.n;
  uint32_t staged = min(i1 - i0 + 32u, RN_STAGE);
  if (i0 + staged > p.n) staged = p.n - i0;
  for (uint32_t k = lane; k < staged; k += 32) stage[k] = p.in[i0 + k];
  const int nc = min(p.cols, RN_SCOLS);
  for (int c = lane; c < nc; c += 32) {
    fr[c] = (int)line < p.rows_tab ? p.target[(size_t)line * p.cols + c] : 0.0f;
    fr[RN_SCOLS + c] = (int)line + 1 < p.rows_tab ? p.target[(size_t)(line + 1) * p.cols + c] : 0.0f;
    fr[2 * RN_SCOLS + c] = (int)line < p.rows ? p.prev[(size_t)line * p.cols + c] : 0.0f;
  }
  __syncwarp();
  const uint8_t *s = p.in;
  const uint32_t n = p.n;
  auto at = [&](uint32_t i) -> uint32_t { return (i - i0) < staged ? stage[i - i0] : (i < n ? s[i] : 0u); };

  // segments: nominal start i0 + l * S, moved forward off UTF-8 continuation bytes (they belong to the lane in front,
  // whether a lead byte consumes them or they stand alone)
  const uint32_t len = i1 - i0, S = (len + 31u) / 32u;
  uint32_t from = min(i0 + (uint32_t)lane * S, i1);
  if (lane > 0)
    while (from < i1 && (at(from) & 0xC0u) == 0x80u) from++;
  uint32_t to = __shfl_down_sync(0xffffffffu, from, 1);
  if (lane == 31) to = i1;
  // state maps of the segments, chained from N at the start of the line
  int ex[3] = {RS_N, RS_E, RS_C};
  for (uint32_t i = from; i < to; i++) {
    const uint32_t c = at(i);
#pragma unroll
    for (int e = 0; e < 3; e++) ex[e] = rain_step(ex[e], c);
  }
  const int map = ex[0] | (ex[1] << 2) | (ex[2] << 4);
  int entry = RS_N, cur = RS_N;
  for (int l = 0; l < 32; l++) {
    const int m = __shfl_sync(0xffffffffu, map, l);
    if (l == lane) entry = cur;
    cur = (m >> (2 * cur)) & 3;
  }
  // structure: characters and visits per segment -> every lane's column and the visits its first cell already had
  uint32_t complex = 0;
  StructSink st;
  rain_tokens(at, n, from, to, i1, entry, st, &complex);
  const int visits = st.nchar ? st.v_tail : st.v_head; // colour visits pending behind this segment's last character
  int col0 = 0, ord0 = 0, ccol = 0, cord = 0;
  bool later = false; // does any later lane visit anything
  for (int l = 0; l < 32; l++) {
    const int nch = __shfl_sync(0xffffffffu, st.nchar, l), vh = __shfl_sync(0xffffffffu, st.v_head, l),
              vt = __shfl_sync(0xffffffffu, visits, l);
    if (l == lane) col0 = ccol, ord0 = cord;
    if (l > lane && (nch > 0 || vh > 0)) later = true;
    ccol += nch;
    cord = nch ? vt : cord + vh;
  }
  if (WRITE) {
    ByteOut o{p.out + p.line_off[line] + p.lane_off[(size_t)line * 32 + lane]};
    EmitSink<decltype(at), ByteOut> sk{p, at, o, fr, (int)line, col0, (int)line, true, false};
    sk.catch_up(ord0);
    sk.tail_visits = ord0; // visits of the first cell made by the lanes in front count as pending, too
    rain_tokens(at, n, from, to, i1, entry, sk, &complex);
    if (!later) sk.flush_tail();
  } else {
    CountOut o;
    EmitSink<decltype(at), CountOut> sk{p, at, o, fr, (int)line, col0, (int)line, false, false};
    sk.catch_up(ord0);
    rain_tokens(at, n, from, to, i1, entry, sk, &complex);
    uint32_t inc = o.n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += v;
    }
    p.lane_off[(size_t)line * 32 + lane] = inc - o.n;
    if (lane == 31) p.line_len[line] = inc;
    if (complex) atomicOr(&p.result[1], 1u);
  }
}

// exclusive scan of the line lengths by one CTA (a frame has a few hundred lines)
__global__ void __launch_bounds__(RN_NT) k_rain_scan(const RainParams p) {
  __shared__ uint32_t s_w[RN_NT / 32], s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < p.nlines; base += RN_NT) {
    const uint32_t k = base + tid;
    const uint32_t v = k < p.nlines ? p.line_len[k] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_w[warp] = inc;
    __syncthreads();
    uint32_t pre = s_carry;
    for (int w = 0; w < warp; w++) pre += s_w[w];
    if (k < p.nlines) p.line_off[k] = pre + inc - v;
    __syncthreads();
    if (tid == RN_NT - 1) s_carry = pre + inc;
    __syncthreads();
  }
  if (tid == 0) p.result[0] = s_carry;
}

inline float rain_hash(float x, float y) { // random_float, digital_rain.c:31-35 (host float, libm)
  float dt = x * 12.9898f + y * 78.233f;
  float sn = fmodf(dt, (float)M_PI);
  return fmodf(sinf(sn) * 43758.5453f, 1.0f);
}
// get_rain_brightness, digital_rain.c:70-91, expression for expression (host float, libm sinf/floorf)
inline float rain_target(const digital_rain_t *r, int col, int row, float sim_time) {
  const digital_rain_column_t *c = &r->columns[col];
  float column_time = c->time_offset + sim_time * r->fall_speed * c->speed_multiplier;
  float rain_time = (column_time - (float)row) / r->raindrop_length;
  rain_time = rain_time + 0.3f * sinf((float)1.4142135623730951 * rain_time) + 0.2f * sinf((float)2.23606797749979 * rain_time);
  return 1.0f - (rain_time - floorf(rain_time));
}

} // namespace

extern "C" {

// lib/video/anim/digital_rain.c:97-153
digital_rain_t *digital_rain_init(int num_columns, int num_rows) {
  if (num_columns <= 0 || num_rows <= 0) {
    set_error(E_INVALID_PARAM, "digital_rain_init: invalid dimensions %dx%d", num_columns, num_rows);
    return nullptr;
  }
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  RainImpl *im = new RainImpl();
  digital_rain_t *r = &im->pub;
  memset(r, 0, sizeof(*r));
  im->device = cx->device;
  r->num_columns = num_columns;
  r->num_rows = num_rows;
  r->columns = (digital_rain_column_t *)calloc((size_t)num_columns, sizeof(digital_rain_column_t));
  const size_t grid = (size_t)num_columns * (size_t)num_rows;
  if (!r->columns || cudaMalloc((void **)&r->previous_brightness, grid * sizeof(float)) != cudaSuccess ||
      cudaMalloc((void **)&im->alt, grid * sizeof(float)) != cudaSuccess ||
      cudaMemset(r->previous_brightness, 0, grid * sizeof(float)) != cudaSuccess) {
    set_error(E_MEMORY, "digital_rain_init: cannot allocate a %dx%d grid", num_columns, num_rows);
    free(r->columns);
    if (r->previous_brightness) cudaFree(r->previous_brightness);
    if (im->alt) cudaFree(im->alt);
    delete im;
    return nullptr;
  }
  for (int col = 0; col < num_columns; col++) { // :131-136
    r->columns[col].time_offset = rain_hash((float)col, 0.0f) * 1000.0f;
    r->columns[col].speed_multiplier = rain_hash((float)col + 0.1f, 0.0f) * 0.5f + 0.5f;
    r->columns[col].phase_offset = rain_hash((float)col + 0.2f, 0.0f) * (float)M_PI * 2.0f;
  }
  r->fall_speed = 3.0f; // DIGITAL_RAIN_DEFAULT_* (digital_rain.h)
  r->raindrop_length = 12.0f;
  r->brightness_decay = 0.1f;
  r->animation_speed = 1.0f;
  r->color_r = 0, r->color_g = 255, r->color_b = 80;
  r->cursor_brightness = 2.0f;
  r->rainbow_mode = false;
  r->first_frame = true;
  r->time = 0.0f;
  return r;
}

void digital_rain_destroy(digital_rain_t *rain) {
  if (!rain) return;
  RainImpl *im = reinterpret_cast<RainImpl *>(rain);
  free(rain->columns);
  if (rain->previous_brightness) cudaFree(rain->previous_brightness);
  if (im->alt) cudaFree(im->alt);
  delete im;
}

void digital_rain_reset(digital_rain_t *rain) { // :164-175
  if (!rain) return;
  rain->time = 0.0f;
  rain->first_frame = true;
  cudaMemset(rain->previous_brightness, 0, (size_t)rain->num_columns * rain->num_rows * sizeof(float));
}
void digital_rain_set_fall_speed(digital_rain_t *rain, float speed) {
  if (rain) rain->fall_speed = speed;
}
void digital_rain_set_raindrop_length(digital_rain_t *rain, float length) {
  if (rain) rain->raindrop_length = length;
}
void digital_rain_set_color(digital_rain_t *rain, uint8_t r, uint8_t g, uint8_t b) {
  if (rain) rain->color_r = r, rain->color_g = g, rain->color_b = b;
}
void digital_rain_set_color_from_filter(digital_rain_t *rain, int filter) { // :205-230
  if (!rain) return;
  if (filter == 0) {
    rain->rainbow_mode = false;
    digital_rain_set_color(rain, 0, 255, 80);
    return;
  }
  if (filter == 12) {
    rain->rainbow_mode = true;
    digital_rain_set_color(rain, 255, 0, 0);
    return;
  }
  rain->rainbow_mode = false;
  int mode;
  uint32_t rgb;
  if (filter > 0 && filter < 12 && resolve_pixel_filter(filter, 0.0f, &mode, &rgb)) // the registry's colour (:226-229)
    digital_rain_set_color(rain, (uint8_t)(rgb >> 16), (uint8_t)(rgb >> 8), (uint8_t)rgb);
}

// lib/video/anim/digital_rain.c:366-520
char *digital_rain_apply(digital_rain_t *rain, const char *frame, float delta_time) {
  if (!rain || !frame) {
    set_error(E_INVALID_PARAM, "digital_rain_apply: NULL parameter");
    return nullptr;
  }
  RainImpl *im = reinterpret_cast<RainImpl *>(rain);
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  if (cx->device != im->device && !peer_ok(cx->device, im->device)) {
    set_error(E_INVALID_STATE, "digital_rain_apply: the rain state lives on GPU %d, which GPU %d cannot address", im->device,
              cx->device);
    return nullptr;
  }
  rain->time += delta_time * rain->animation_speed; // :373-379
  const float sim_time = rain->time;
  if (rain->rainbow_mode) color_filter_calculate_rainbow(sim_time, &rain->color_r, &rain->color_g, &rain->color_b);
  const size_t n = strlen(frame);
  if (n == 0 || n > 0x3fffff00u) {
    if (n) {
      set_error(E_INVALID_PARAM, "digital_rain_apply: frame too large");
      return nullptr;
    }
    rain->first_frame = false;
    char *e = (char *)user_alloc(1);
    if (e) e[0] = '\0';
    return e;
  }
  size_t nl = 0;
  for (const char *q = frame; (q = (const char *)memchr(q, '\n', (size_t)(frame + n - q))) != nullptr; q++) nl++;
  const size_t nlines = nl + 1, rows_tab = nl + 2, cols = (size_t)rain->num_columns;
  if (rows_tab * cols > ((size_t)1 << 26)) {
    set_error(E_INVALID_PARAM, "digital_rain_apply: %zu lines x %zu columns is beyond the brightness table", nl, cols);
    return nullptr;
  }
  const int chunks = (int)((n + RN_CHUNK - 1) / RN_CHUNK);
  const size_t tab_bytes = rows_tab * cols * sizeof(float);
  const size_t in_bytes = ((n + 15) & ~(size_t)15) + tab_bytes;
  // worst case: every byte a visible character -> 19 bytes of code + the byte itself
  const size_t out_cap = n * 20 + 64;
  const size_t words = (size_t)chunks + 3 * nlines + 32 * nlines + 16;
  if (sync_foreign(cx, cx->stream) != E_OK) return nullptr;
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, in_bytes) || !grow_device(&cx->d_in, &cx->d_in_cap, in_bytes) ||
      !grow_device(&cx->d_out, &cx->d_out_cap, out_cap) || !grow_device((uint8_t **)&cx->d_len, &cx->d_len_cap, words * 4) ||
      !grow_pinned((uint8_t **)&cx->h_len, &cx->h_len_cap, 256))
    return nullptr;
  memcpy(cx->h_in, frame, n);
  float *tab = reinterpret_cast<float *>(cx->h_in + ((n + 15) & ~(size_t)15));
  for (size_t r = 0; r < rows_tab; r++)
    for (size_t c = 0; c < cols; c++) tab[r * cols + c] = rain_target(rain, (int)c, (int)r, sim_time);

  RainParams p{};
  p.in = cx->d_in;
  p.n = (uint32_t)n;
  p.out = cx->d_out;
  uint32_t *w = cx->d_len;
  p.result = w;
  p.chunk_nl = w + 16;
  p.line_start = p.chunk_nl + chunks;
  p.line_len = p.line_start + nlines;
  p.line_off = p.line_len + nlines;
  p.lane_off = p.line_off + nlines;
  p.nlines = (uint32_t)nlines;
  p.target = reinterpret_cast<const float *>(cx->d_in + ((n + 15) & ~(size_t)15));
  p.cols = rain->num_columns;
  p.rows = rain->num_rows;
  p.rows_tab = (int)rows_tab;
  p.prev = rain->previous_brightness;
  p.prev_new = im->alt;
  p.decay = rain->brightness_decay;
  p.first_frame = rain->first_frame ? 1 : 0;
  p.rain_rgb = ((uint32_t)rain->color_r << 16) | ((uint32_t)rain->color_g << 8) | rain->color_b;
  cudaStream_t st = cx->stream;
  auto fail = [&](const char *what) -> char * {
    set_error(E_INVALID_STATE, "digital_rain_apply: CUDA failure (%s: %s)", what, cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(st);
    return nullptr;
  };
  if (cudaMemcpyAsync(cx->d_in, cx->h_in, in_bytes, cudaMemcpyHostToDevice, st) != cudaSuccess ||
      cudaMemsetAsync(p.result, 0, 16 * sizeof(uint32_t), st) != cudaSuccess)
    return fail("upload");
  k_rain_nl_count<<<(unsigned)chunks, RN_NT, 0, st>>>(p);
  k_rain_line_starts<<<(unsigned)chunks, RN_NT, 0, st>>>(p);
  const unsigned walk_grid = (unsigned)((nlines + RN_LPB - 1) / RN_LPB);
  constexpr size_t walk_smem = (size_t)RN_LPB * RN_LINE_SMEM; // 72 KB: above the 48 KB default, opted in per device
  {
    static std::atomic<uint64_t> configured{0};
    if (!((configured.load(std::memory_order_acquire) >> cx->device) & 1ull)) {
      if (cudaFuncSetAttribute(k_rain_warp<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_smem) != cudaSuccess ||
          cudaFuncSetAttribute(k_rain_warp<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_smem) != cudaSuccess)
        return fail("shared-memory opt-in");
      configured.fetch_or(1ull << cx->device, std::memory_order_release);
    }
  }
  k_rain_warp<false><<<walk_grid, 32 * RN_LPB, walk_smem, st>>>(p);
  k_rain_scan<<<1, RN_NT, 0, st>>>(p);
  count_launch(4);
  if (cudaGetLastError() != cudaSuccess ||
      cudaMemcpyAsync(cx->h_len, p.result, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || wait_stream(cx) != E_OK)
    return fail("count pass");
  const bool serial = cx->h_len[1] != 0;
  if (serial) { // an escape sequence swallowed a newline: one range, one thread, the reference's loop as it is
    p.nlines = 1;
    if (cudaMemsetAsync(p.line_start, 0, sizeof(uint32_t), st) != cudaSuccess ||
        cudaMemsetAsync(p.result, 0, 16 * sizeof(uint32_t), st) != cudaSuccess)
      return fail("fallback setup");
    k_rain_serial<false><<<1, 32, 0, st>>>(p);
    k_rain_scan<<<1, RN_NT, 0, st>>>(p);
    count_launch(2);
    if (cudaGetLastError() != cudaSuccess ||
        cudaMemcpyAsync(cx->h_len, p.result, 8, cudaMemcpyDeviceToHost, st) != cudaSuccess || wait_stream(cx) != E_OK)
      return fail("fallback count pass");
  }
  const size_t total = cx->h_len[0];
  if (total + 1 > out_cap || !grow_pinned(&cx->h_out, &cx->h_out_cap, total + 16)) {
    if (total + 1 > out_cap) set_error(E_INVALID_STATE, "digital_rain_apply: output larger than its bound");
    return nullptr;
  }
  // cells this frame does not visit keep their value: the new state starts as a copy of the old one
  if (cudaMemcpyAsync(im->alt, rain->previous_brightness, (size_t)rain->num_columns * rain->num_rows * sizeof(float),
                      cudaMemcpyDeviceToDevice, st) != cudaSuccess)
    return fail("state copy");
  if (serial) k_rain_serial<true><<<1, 32, 0, st>>>(p);
  else k_rain_warp<true><<<walk_grid, 32 * RN_LPB, walk_smem, st>>>(p);
  count_launch();
  if (cudaGetLastError() != cudaSuccess ||
      cudaMemcpyAsync(cx->h_out, cx->d_out, total, cudaMemcpyDeviceToHost, st) != cudaSuccess || wait_stream(cx) != E_OK)
    return fail("write pass");
  { // the buffers swap roles
    float *t = rain->previous_brightness;
    rain->previous_brightness = im->alt;
    im->alt = t;
  }
  rain->first_frame = false; // :517
  char *res = (char *)user_alloc(total + 1);
  if (!res) return nullptr;
  memcpy(res, cx->h_out, total);
  res[total] = '\0';
  return res;
}

} // extern "C"
