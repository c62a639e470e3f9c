// render_dev.cuh — device code of the row kernels (templates), shared by the per-mode instantiation units
// (rk_mode*.cu) and render_kernels.cu.  See render_kernels.cu for the overview.
#pragma once
#include <atomic>
#include <cstdlib>

#include "render.cuh"

namespace acb {

constexpr uint16_t NONE16 = 0xFFFF;
constexpr int kBandUnroll = 12; // independent 16-byte loads in flight per thread in the streaming box filter

// ------------------------------------------------------------------ small integer helpers
__device__ __forceinline__ int luma_of(uint32_t c) { // foreground.c:93  (77R+150G+29B+128)>>8
  return (int)((77u * ((c >> 16) & 255u) + 150u * ((c >> 8) & 255u) + 29u * (c & 255u) + 128u) >> 8);
}
__device__ __forceinline__ int luma76_of(uint32_t c) { // halfblock.c:239-240  (76R+150G+29B)>>8
  return (int)((76u * ((c >> 16) & 255u) + 150u * ((c >> 8) & 255u) + 29u * (c & 255u)) >> 8);
}
__device__ __forceinline__ int q256_of(uint32_t c) { // ansi.c:360-379
  int r = (c >> 16) & 255, g = (c >> 8) & 255, b = c & 255;
  int avg = (r + g + b) / 3;
  int d = abs(r - avg) + abs(g - avg) + abs(b - avg);
  if (d < 30) return 232 + (avg * 23) / 255;
  return 16 + 36 * ((r * 5) / 255) + 6 * ((g * 5) / 255) + (b * 5) / 255;
}
static __constant__ uint8_t c_ansi16[16][4] = { // ansi.c:442-459
    {0, 0, 0, 0},       {128, 0, 0, 0},   {0, 128, 0, 0},   {128, 128, 0, 0}, {0, 0, 128, 0},   {128, 0, 128, 0},
    {0, 128, 128, 0},   {192, 192, 192, 0}, {128, 128, 128, 0}, {255, 0, 0, 0},   {0, 255, 0, 0},   {255, 255, 0, 0},
    {0, 0, 255, 0},     {255, 0, 255, 0}, {0, 255, 255, 0}, {255, 255, 255, 0}};
__device__ __forceinline__ int q16_rgb(int r, int g, int b) { // ansi.c:437-477, first minimum wins
  int best = 0, bestd = 0x7fffffff;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    int dr = r - c_ansi16[i][0], dg = g - c_ansi16[i][1], db = b - c_ansi16[i][2];
    int d = dr * dr + dg * dg + db * db;
    if (d < bestd) {
      bestd = d;
      best = i;
    }
  }
  return best;
}
__device__ __forceinline__ int q16_of(uint32_t c) { return q16_rgb((c >> 16) & 255, (c >> 8) & 255, c & 255); }

__device__ __forceinline__ bool rep_profitable(uint32_t run) { // output_buffer.c:148-154
  if (run <= 2) return false;
  uint32_t k = run - 1;
  uint32_t digits = k >= 1000u ? (k >= 10000u ? 5u : 4u) : (k >= 100u ? 3u : (k >= 10u ? 2u : 1u)); // k < 100000 here
  return k > digits + 3u;
}

// Barrier policies: the one-tile-per-CTA kernel synchronises the whole CTA; in the warp-specialised persistent
// kernel only the 8 consumer warps (threads 0..255) take part, on named barrier 1, while the producer warp runs ahead.
struct SyncAll {
  static __device__ __forceinline__ void sync() { __syncthreads(); }
};
struct SyncWarp { // a single warp owns the data: warp-level barrier only
  static __device__ __forceinline__ void sync() { __syncwarp(); }
};
template <int NT> struct SyncConsumers {
  static __device__ __forceinline__ void sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }
};

// ------------------------------------------------------------------ byte sinks
// Decimal strings of 0..255 for the emitters: s_dec3[v] = ASCII digits, most significant first, in bytes 0..2 and
// the digit count in byte 3 (the dec3 table of common.c:546-570).  put_dec3 stores all three digit bytes blindly and
// advances by the count: the one or two surplus bytes are overwritten by whatever the same thread emits next (every
// number on this path is followed by ';' or 'm' plus at least one glyph byte), which removes all digit branching.
static __shared__ uint32_t s_dec3[256];
template <int NT> __device__ __forceinline__ void init_dec3(int tid) {
  for (int v = tid; v < 256; v += NT) {
    const uint32_t h = (uint32_t)v / 100u, r = (uint32_t)v - h * 100u, t = r / 10u, u = r - t * 10u;
    uint32_t e;
    if (h) e = ('0' + h) | (('0' + t) << 8) | (('0' + u) << 16) | (3u << 24);
    else if (t) e = ('0' + t) | (('0' + u) << 8) | (2u << 24);
    else e = ('0' + u) | (1u << 24);
    s_dec3[v] = e;
  }
}

struct CountSink {
  uint32_t n = 0;
  __device__ __forceinline__ void put(uint8_t) { ++n; }
  __device__ __forceinline__ void put_dec3(uint32_t e) { n += e >> 24; }
};
struct WriteSink {
  uint8_t *p;
  __device__ __forceinline__ void put(uint8_t c) { *p++ = c; }
  __device__ __forceinline__ void put_dec3(uint32_t e) {
    p[0] = (uint8_t)e;
    p[1] = (uint8_t)(e >> 8);
    p[2] = (uint8_t)(e >> 16);
    p += e >> 24;
  }
};
struct SmemSink { // same, but the destination is known to be shared memory: STS with a 32-bit address
  uint32_t a;
  __device__ __forceinline__ void put(uint8_t c) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"((uint32_t)c) : "memory");
    ++a;
  }
  __device__ __forceinline__ void put_dec3(uint32_t e) {
    asm volatile("st.shared.u8 [%0], %1;\n\tst.shared.u8 [%0+1], %2;\n\tst.shared.u8 [%0+2], %3;" ::"r"(a), "r"(e & 255u),
                 "r"((e >> 8) & 255u), "r"((e >> 16) & 255u)
                 : "memory");
    a += e >> 24;
  }
};

template <class S> __device__ __forceinline__ void put_u8dec(S &s, uint32_t v) { // dec3 table, common.c:546-570
  if (v >= 100u) {
    uint32_t h = v / 100u, r = v - h * 100u;
    s.put((uint8_t)('0' + h));
    s.put((uint8_t)('0' + r / 10u));
    s.put((uint8_t)('0' + r % 10u));
  } else if (v >= 10u) {
    s.put((uint8_t)('0' + v / 10u));
    s.put((uint8_t)('0' + v % 10u));
  } else {
    s.put((uint8_t)('0' + v));
  }
}
template <class S> __device__ __forceinline__ void put_u32dec(S &s, uint32_t v) { // ob_u32, output_buffer.c:92-104
  uint8_t t[10];
  int i = 0;
  do {
    t[i++] = (uint8_t)('0' + v % 10u);
    v /= 10u;
  } while (v);
  while (i--) s.put(t[i]);
}
template <class S> __device__ __forceinline__ void put_sgr_rgb(S &s, bool bg, uint32_t c) { // ansi.c:143-195
  s.put(0x1b);
  s.put('[');
  s.put(bg ? '4' : '3');
  s.put('8');
  s.put(';');
  s.put('2');
  s.put(';');
  put_u8dec(s, (c >> 16) & 255u);
  s.put(';');
  put_u8dec(s, (c >> 8) & 255u);
  s.put(';');
  put_u8dec(s, c & 255u);
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_sgr_256(S &s, bool bg, uint32_t idx) { // ansi.c:326-357
  s.put(0x1b);
  s.put('[');
  s.put(bg ? '4' : '3');
  s.put('8');
  s.put(';');
  s.put('5');
  s.put(';');
  put_u8dec(s, idx);
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_sgr_16(S &s, bool bg, uint32_t idx) { // ansi.c:384-435
  uint32_t code = (idx < 8u ? 30u + idx : 82u + idx) + (bg ? 10u : 0u);
  s.put(0x1b);
  s.put('[');
  put_u8dec(s, code);
  s.put('m');
}
// table-driven variants used by the row emitters (the arithmetic ones above stay for the serial dither kernel)
template <class S> __device__ __forceinline__ void put_sgr_rgb_l(S &s, bool bg, uint32_t c) {
  s.put(0x1b);
  s.put('[');
  s.put(bg ? '4' : '3');
  s.put('8');
  s.put(';');
  s.put('2');
  s.put(';');
  s.put_dec3(s_dec3[(c >> 16) & 255u]);
  s.put(';');
  s.put_dec3(s_dec3[(c >> 8) & 255u]);
  s.put(';');
  s.put_dec3(s_dec3[c & 255u]);
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_sgr_256_l(S &s, bool bg, uint32_t idx) {
  s.put(0x1b);
  s.put('[');
  s.put(bg ? '4' : '3');
  s.put('8');
  s.put(';');
  s.put('5');
  s.put(';');
  s.put_dec3(s_dec3[idx & 255u]);
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_sgr_16_l(S &s, bool bg, uint32_t idx) {
  const uint32_t code = (idx < 8u ? 30u + idx : 82u + idx) + (bg ? 10u : 0u);
  s.put(0x1b);
  s.put('[');
  s.put_dec3(s_dec3[code]);
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_reset(S &s) {
  s.put(0x1b);
  s.put('[');
  s.put('0');
  s.put('m');
}
template <class S> __device__ __forceinline__ void put_rep(S &s, uint32_t extra) { // output_buffer.c:156-164
  s.put(0x1b);
  s.put('[');
  put_u32dec(s, extra);
  s.put('b');
}
template <class S> __device__ __forceinline__ void put_glyph(S &s, const uint8_t *g) {
  int n = g[0];
  for (int i = 0; i < n; i++) s.put(g[1 + i]);
}
template <class S> __device__ __forceinline__ void put3(S &s, uint8_t a, uint8_t b, uint8_t c) {
  s.put(a);
  s.put(b);
  s.put(c);
}

// ------------------------------------------------------------------ block-wide row scans
struct OpAdd {
  static __device__ __forceinline__ int id() { return 0; }
  static __device__ __forceinline__ int ap(int a, int b) { return a + b; }
};
struct OpMax {
  static __device__ __forceinline__ int id() { return -1; }
  static __device__ __forceinline__ int ap(int a, int b) { return a > b ? a : b; }
};

// Scans value_of(x), x in [0,w), in x order over NT threads; calls store(x, inclusive, exclusive).
// Returns the total (valid in every thread).  s_tmp: 2 * (NT / 32) ints of shared memory.
// One barrier per pass of NT cells plus one at the end: the warp totals of a pass go to one of two alternating
// buffers, every thread folds all of them itself (broadcast reads) to get both its prefix and the running carry, so
// there is no single-thread carry update to wait for.  (The first version took four barriers per pass and two around
// the loop — at 8 CTAs per SM the nearest-neighbour kernels spent most of their time in them, profiles/r02a_nn_*.)
template <class Op, class Sync, int NT, class F, class G>
__device__ __forceinline__ int row_scan(int w, F value_of, G store, int *s_tmp, int tid) {
  constexpr int NW = NT / 32;
  const int lane = tid & 31, warp = tid >> 5;
  int carry = Op::id();
  int pass = 0;
  for (int base = 0; base < w; base += NT, pass ^= 1) {
    int *buf = s_tmp + pass * NW;
    int x = base + tid;
    int v = x < w ? value_of(x) : Op::id();
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc = Op::ap(o, inc);
    }
    int up = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 31) buf[warp] = inc;
    Sync::sync();
    // fold the NW warp totals with one more shuffle scan (lane i holds warp i's total): lane w-1 ends up with the
    // prefix of warp w, the last lane with the pass total
    int tot = lane < NW ? buf[lane] : Op::id();
#pragma unroll
    for (int d = 1; d < NW; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, tot, d);
      if (lane >= d) tot = Op::ap(o, tot);
    }
    const int before = __shfl_sync(0xffffffffu, tot, warp > 0 ? warp - 1 : 0);
    const int pre = warp > 0 ? Op::ap(carry, before) : carry;
    if (x < w) store(x, Op::ap(pre, inc), lane == 0 ? pre : Op::ap(pre, up));
    carry = Op::ap(carry, __shfl_sync(0xffffffffu, tot, NW - 1));
  }
  Sync::sync(); // the next scan (or whoever reads what store() wrote) starts from a quiet s_tmp
  return carry;
}

// ------------------------------------------------------------------ shared-memory layout
struct Layout {
  uint32_t lut, cT, cB, key, hpos, rend, off, V, outb, total;
  uint32_t set2; // byte distance from an array of the first per-tile set (cT .. off) to the same array of the second set
};
// V (column sums, phase A only) and the row staging buffer (phase B4 only) never live at the same time, so they
// share one region unless `no_alias` (tuning knob ACB200_TUNE_NOALIAS) asks for separate ones.
// two_sets: the persistent direct-output kernel prepares tile k+1 before it writes out tile k, so the per-tile arrays
// (cells, run keys, heads, ends, offsets) exist twice.
__host__ __device__ inline Layout make_layout(int mode, int sp, int cols, int src_w, uint32_t out_bytes,
                                              int no_alias = 0, int two_sets = 0) {
  Layout L;
  uint32_t o = 0;
  L.lut = o;
  o += (mode <= EM_TRUE_FG) ? al16((uint32_t)sizeof(GlyphLut)) : 0u; // the half-block grammars print no palette glyph
  const uint32_t set_begin = o;
  L.cT = o;
  o += al16(4u * cols);
  L.cB = o;
  o += (mode >= EM_HB_TRUE && mode <= EM_HB_MONO) ? al16(4u * cols) : 0u;
  L.key = o;
  o += al16(2u * cols);
  L.hpos = o;
  o += al16(2u * cols);
  L.rend = o;
  o += al16(2u * cols);
  L.off = o;
  o += al16(4u * cols);
  L.set2 = o - set_begin;
  if (two_sets) o += L.set2;
  const uint32_t vbytes = sp == SP_BOX_STREAM ? al16(2u * 3u * src_w) : 0u;
  const uint32_t obytes = out_bytes ? al16(out_bytes) + 16u : 0u; // + room to stage with the destination's phase
  L.V = o;
  if (no_alias) {
    o += vbytes;
    L.outb = o;
    o += obytes;
  } else {
    L.outb = o;
    o += vbytes > obytes ? vbytes : obytes;
  }
  L.total = o;
  return L;
}

// ------------------------------------------------------------------ phase A: resized pixels
__device__ __forceinline__ uint32_t load_px(const uint8_t *p) {
  return ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | (uint32_t)p[2];
}
// The same pixel out of aligned 32-bit words: one load when the three bytes sit inside a word, two when they straddle
// it (never a read past the pixel's last byte's word), then a funnel shift and a byte swap — instead of three dependent
// byte loads per sample.  Adjacent samples of a 1:1 (pre-gathered) image share their words, so a warp's loads coalesce.
__device__ __forceinline__ uint32_t load_px_w(const uint8_t *p) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  const uint32_t sh = (uint32_t)(a & 3u);
  const uint32_t *w = reinterpret_cast<const uint32_t *>(a - sh);
  const uint32_t w0 = __ldg(w);
  const uint32_t w1 = sh >= 2u ? __ldg(w + 1) : 0u;
  const uint32_t v = __funnelshift_r(w0, w1, sh * 8u); // bytes r,g,b in memory order: r | g<<8 | b<<16
  return __byte_perm(v, 0u, 0x4012);                   // -> 0x00RRGGBB
}

// apply_color_filter on one pixel (color_filter.c:238-267, 305-318, 338-341; rgb_to_grayscale color_filter.h:172).
// A pointwise map commutes with nearest-neighbour sampling, so the reference's whole-image pre-pass
// (display.c:609-624) costs one evaluation per sampled pixel here.
__device__ __forceinline__ uint32_t filter_px(uint32_t c, int mode, uint32_t frgb) {
  if (mode == FM_NONE) return c;
  uint32_t gray = (77u * ((c >> 16) & 255u) + 150u * ((c >> 8) & 255u) + 29u * (c & 255u)) >> 8;
  if (mode == FM_RAINBOW) gray = 179u + (gray * 76u) / 255u;
  const uint32_t fr = (frgb >> 16) & 255u, fg = (frgb >> 8) & 255u, fb = frgb & 255u;
  if (mode == FM_ON_WHITE) {
    const uint32_t w = 255u * gray, ig = 255u - gray;
    return (((fr * ig + w) / 255u) << 16) | (((fg * ig + w) / 255u) << 8) | ((fb * ig + w) / 255u);
  }
  return (((fr * gray) / 255u) << 16) | (((fg * gray) / 255u) << 8) | ((fb * gray) / 255u);
}

// nearest neighbour — image.c:293-325 (u32 fixed-point, wraps like the reference); the sampled coordinate is
// mirrored when the display path flips the image first (display.c:563-590)
__device__ __forceinline__ const uint8_t *nn_row(const RenderParams &p, const uint8_t *frame, int y, uint32_t yr) {
  uint32_t sy;
  if (p.pregathered) {
    sy = (uint32_t)y;
  } else {
    sy = ((uint32_t)y * yr) >> 16;
    if (sy >= (uint32_t)p.src_h) sy = (uint32_t)p.src_h - 1;
    if (p.flip_y) sy = (uint32_t)p.src_h - 1u - sy;
  }
  return frame + (size_t)sy * (size_t)p.src_w * 3u;
}
// One pass samples the top pixel row of the text row and, for half blocks, the bottom one too: both loads of a cell are
// in flight together (one memory latency per tile instead of two).  outB == nullptr: single pixel row.
template <int NT>
__device__ __forceinline__ void cells_nn(const RenderParams &p, const uint8_t *frame, int y, uint32_t *outT,
                                         uint32_t *outB) {
  const uint32_t xr = p.nn_xr, yr = p.nn_yr; // 64-bit divisions done once on the host
  const uint8_t *rowT = nn_row(p, frame, y, yr);
  const uint8_t *rowB = outB ? nn_row(p, frame, y + 1, yr) : rowT;
  for (int x = threadIdx.x; x < p.cols; x += NT) {
    uint32_t sx = ((uint32_t)x * xr) >> 16;
    if (sx >= (uint32_t)p.src_w) sx = (uint32_t)p.src_w - 1;
    if (p.flip_x) sx = (uint32_t)p.src_w - 1u - sx;
    const uint32_t a = load_px_w(rowT + (size_t)sx * 3u);
    const uint32_t b = outB ? load_px_w(rowB + (size_t)sx * 3u) : 0u;
    outT[x] = filter_px(a, p.filt_mode, p.filt_rgb);
    if (outB) outB[x] = filter_px(b, p.filt_mode, p.filt_rgb);
  }
}

__device__ __forceinline__ void box_range(int d, int src, int dst, int &a, int &b) {
  // d < dst <= 3840 and src <= 10000 (make_plan), so the products fit 32 bits: plain u32 divisions
  a = (int)(((uint32_t)d * (uint32_t)src) / (uint32_t)dst);
  b = (int)(((uint32_t)(d + 1) * (uint32_t)src) / (uint32_t)dst);
  if (b <= a) b = a + 1;
  if (b > src) b = src;
  if (a >= src) a = src - 1;
}

// box filter, any geometry: one thread per destination pixel, byte loads
template <int NT>
__device__ __forceinline__ void cells_box_generic(const RenderParams &p, const uint8_t *frame, int y, uint32_t *out) {
  int y0, y1;
  box_range(y, p.src_h, p.rows_px, y0, y1);
  for (int x = threadIdx.x; x < p.cols; x += NT) {
    int x0, x1;
    box_range(x, p.src_w, p.cols, x0, x1);
    uint32_t sr = 0, sg = 0, sb = 0;
    // the box is taken over the flipped, filtered image (display.c order): mirror the range, filter every pixel
    const int xs = p.flip_x ? p.src_w - x1 : x0;
    for (int yy = y0; yy < y1; yy++) {
      const int ys = p.flip_y ? p.src_h - 1 - yy : yy;
      const uint8_t *q = frame + ((size_t)ys * p.src_w + xs) * 3u;
      if (p.filt_mode == FM_NONE) {
        for (int xx = x0; xx < x1; xx++, q += 3) {
          sr += q[0];
          sg += q[1];
          sb += q[2];
        }
      } else {
        for (int xx = x0; xx < x1; xx++, q += 3) {
          const uint32_t c = filter_px(load_px(q), p.filt_mode, p.filt_rgb);
          sr += c >> 16;
          sg += (c >> 8) & 255u;
          sb += c & 255u;
        }
      }
    }
    uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)(y1 - y0), h = n >> 1;
    out[x] = (((sr + h) / n) << 16) | (((sg + h) / n) << 8) | ((sb + h) / n);
  }
}

// L2 policies for the two streams of the box kernels (measured on the 256 x 4K workload, profiles/r01x / r01y sweeps):
// the source is read exactly once, so it is marked evict-first in L2 (+2.3 % bandwidth: it stops pushing the output
// rows out of L2 before they are complete sectors); the output rows are marked evict-last (+0.2 % on top).
// `.L2::256B` prefetch granularity on the loads and `.cs` stores measured neutral / slightly negative.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ uint4 ldg_stream(const uint4 *p) { // read-once data: no L1 allocation, L2 evict-first
  // The policy word is re-created at every load on purpose: ptxas rematerialises it where it is used (createpolicy is
  // one uniform-datapath instruction) instead of holding a 64-bit value across the band loop.  Passing one value in
  // from the caller cost two registers under the kernel's 56-register cap, spilled, and LOST 5 % (r01zz visit).
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void stg_keep(uint4 *p, const uint4 &v, uint64_t pol) { // output rows: L2 evict-last
  asm volatile("st.global.L2::cache_hint.v4.u32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w),
               "l"(pol)
               : "memory");
}
__device__ __forceinline__ void acc16(uint32_t (&a)[8], const uint4 &v) {
  // 16 byte lanes -> 16 u16 partial sums, two per register (even bytes / odd bytes of each word)
  a[0] += v.x & 0x00FF00FFu;
  a[1] += (v.x >> 8) & 0x00FF00FFu;
  a[2] += v.y & 0x00FF00FFu;
  a[3] += (v.y >> 8) & 0x00FF00FFu;
  a[4] += v.z & 0x00FF00FFu;
  a[5] += (v.z >> 8) & 0x00FF00FFu;
  a[6] += v.w & 0x00FF00FFu;
  a[7] += (v.w >> 8) & 0x00FF00FFu;
}
// two rows per call: (even bytes of u) + (even bytes of v) + acc is one LOP3 + LOP3 + IADD3, the odd bytes come out of
// one PRMT each — 12 integer ops per 16 source bytes instead of 20
__device__ __forceinline__ void acc16x2(uint32_t (&a)[8], const uint4 &u, const uint4 &v) {
  a[0] += (u.x & 0x00FF00FFu) + (v.x & 0x00FF00FFu);
  a[1] += __byte_perm(u.x, 0u, 0x4341) + __byte_perm(v.x, 0u, 0x4341);
  a[2] += (u.y & 0x00FF00FFu) + (v.y & 0x00FF00FFu);
  a[3] += __byte_perm(u.y, 0u, 0x4341) + __byte_perm(v.y, 0u, 0x4341);
  a[4] += (u.z & 0x00FF00FFu) + (v.z & 0x00FF00FFu);
  a[5] += __byte_perm(u.z, 0u, 0x4341) + __byte_perm(v.z, 0u, 0x4341);
  a[6] += (u.w & 0x00FF00FFu) + (v.w & 0x00FF00FFu);
  a[7] += __byte_perm(u.w, 0u, 0x4341) + __byte_perm(v.w, 0u, 0x4341);
}

// NR rows of one 16-byte column: NR independent, UNPREDICATED loads in flight, then the pairwise sums.  (A predicated
// "row r+k exists" form costs an ISETP and four zeroing moves per load: 12% of the kernel's instructions, measured.)
template <int NR> __device__ __forceinline__ void band_trip(uint32_t (&a)[8], const uint8_t *&q, uint32_t R) {
  uint4 v[NR];
#pragma unroll
  for (int k = 0; k < NR; k++) {
    v[k] = ldg_stream(reinterpret_cast<const uint4 *>(q));
    q += R;
  }
#pragma unroll
  for (int k = 0; k + 1 < NR; k += 2) acc16x2(a, v[k], v[k + 1]);
  if (NR & 1) acc16x2(a, v[NR - 1], make_uint4(0u, 0u, 0u, 0u));
}

// ---- colour filter fused into the streaming box filter (display path in box mode) -------------
// apply_color_filter is a function of the pixel's BT.601 grey value only (color_filter.c:238-267, 338-341), so the
// filtered image is T[gray(px)] with a 256-entry table.  It is not linear (truncating /255 per pixel), so the band cannot
// be summed first and filtered afterwards: every source pixel goes through the table, then into the same packed u16
// column sums as the unfiltered path.  A thread owns whole pixels — 48 bytes = 16 pixels = three 16-byte loads per
// row — takes four rows per trip (12 independent loads in flight, like the unfiltered band), and per 4 pixels spends
// 2 PRMT + 1 SHR (cut the pixels out of 3 words), 4 DP4A + 4 SHR (grey), 4 LDS (table), 8 ops (re-pack to 3 words).
// The table is replicated once per lane (entry v of lane l at word v*32 + l, i.e. in bank l): a warp's 32 lookups with
// 32 unrelated grey values then hit 32 different banks.  (One shared 1 KB table cost ~3.5 wavefronts per lookup:
// 41.9 M bank conflicts per 64 frames, profiles/r02c_ncu_filtered_summary.txt.)
static __shared__ uint32_t s_filt[256 * 32]; // [gray][lane] = filtered pixel in memory byte order: r | g << 8 | b << 16
template <int NT> __device__ __forceinline__ void init_filt(const RenderParams &p, int tid) {
  for (int w = tid; w < 256 * 32; w += NT) {
    const uint32_t v = (uint32_t)w >> 5; // gray(v,v,v) = v
    s_filt[w] = __byte_perm(filter_px(v * 0x010101u, p.filt_mode, p.filt_rgb), 0u, 0x4012);
  }
}
__device__ __forceinline__ uint32_t gray_rgbx(uint32_t rgbx) { // bytes r,g,b,(ignored): (77r+150g+29b)>>8, color_filter.h:172
  return __dp4a(rgbx, 0x001D964Du, 0u) >> 8;
}
__device__ __forceinline__ void filt4px(const uint32_t *T, uint32_t &w0, uint32_t &w1, uint32_t &w2) { // 4 px in 3 words
  const uint32_t p0 = T[gray_rgbx(w0) << 5]; // T = s_filt + lane
  const uint32_t p1 = T[gray_rgbx(__byte_perm(w0, w1, 0x6543)) << 5];
  const uint32_t p2 = T[gray_rgbx(__byte_perm(w1, w2, 0x5432)) << 5];
  const uint32_t p3 = T[gray_rgbx(w2 >> 8) << 5];
  w0 = p0 | (p1 << 24);
  w1 = (p1 >> 8) | (p2 << 16);
  w2 = (p2 >> 16) | (p3 << 8);
}
__device__ __forceinline__ void filt16px(const uint32_t *T, uint4 &a, uint4 &b, uint4 &c) { // 48 bytes = 16 pixels, in place
  filt4px(T, a.x, a.y, a.z);
  filt4px(T, a.w, b.x, b.y);
  filt4px(T, b.z, b.w, c.x);
  filt4px(T, c.y, c.z, c.w);
}
template <int NR> __device__ __forceinline__ void band_trip_filt(uint32_t (&a)[3][8], const uint8_t *&q, uint32_t R) {
  uint4 v[NR][3];
#pragma unroll
  for (int k = 0; k < NR; k++) {
#pragma unroll
    for (int j = 0; j < 3; j++) v[k][j] = ldg_stream(reinterpret_cast<const uint4 *>(q) + j);
    q += R;
  }
  const uint32_t *T = s_filt + (threadIdx.x & 31);
#pragma unroll
  for (int k = 0; k < NR; k++) filt16px(T, v[k][0], v[k][1], v[k][2]);
#pragma unroll
  for (int j = 0; j < 3; j++) {
#pragma unroll
    for (int k = 0; k + 1 < NR; k += 2) acc16x2(a[j], v[k][j], v[k + 1][j]);
    if (NR & 1) acc16x2(a[j], v[NR - 1][j], make_uint4(0u, 0u, 0u, 0u));
  }
}
// column sums of the FILTERED band into V (same layout as the unfiltered path); requires (3 * src_w) % 48 == 0
template <int NT>
__device__ __forceinline__ void band_sums_filtered(const uint8_t *band, uint32_t R, int nrow, uint16_t *V) {
  const int ngroup = (int)(R / 48u);
  for (int c = threadIdx.x; c < ngroup; c += NT) {
    uint32_t a[3][8];
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = 0; k < 8; k++) a[j][k] = 0u;
    const uint8_t *q = band + (size_t)c * 48u;
    int r = nrow;
    for (; r >= 4; r -= 4) band_trip_filt<4>(a, q, R);
    switch (r) {
    case 3: band_trip_filt<3>(a, q, R); break;
    case 2: band_trip_filt<2>(a, q, R); break;
    case 1: band_trip_filt<1>(a, q, R); break;
    default: break;
    }
#pragma unroll
    for (int j = 0; j < 3; j++) {
      uint4 lo, hi;
      lo.x = __byte_perm(a[j][0], a[j][1], 0x5410);
      lo.y = __byte_perm(a[j][0], a[j][1], 0x7632);
      lo.z = __byte_perm(a[j][2], a[j][3], 0x5410);
      lo.w = __byte_perm(a[j][2], a[j][3], 0x7632);
      hi.x = __byte_perm(a[j][4], a[j][5], 0x5410);
      hi.y = __byte_perm(a[j][4], a[j][5], 0x7632);
      hi.z = __byte_perm(a[j][6], a[j][7], 0x5410);
      hi.w = __byte_perm(a[j][6], a[j][7], 0x7632);
      uint4 *dst = reinterpret_cast<uint4 *>(V + ((size_t)c * 3 + j) * 16);
      dst[0] = lo;
      dst[1] = hi;
    }
  }
}

// box filter, streaming: the band of source rows [y0,y1) is one contiguous byte range; every thread owns
// 16-byte columns of it, sums them down the band in registers (u16 lanes, band <= 256 rows), parks the
// column sums V[3*src_w] in shared memory, then one thread per destination pixel adds its x-range.
template <int NT, class Sync = SyncAll, bool FILT = false>
__device__ __forceinline__ void cells_box_stream(const RenderParams &p, const uint8_t *frame, int y, uint32_t *out,
                                                 uint16_t *V) {
  int y0, y1;
  box_range(y, p.src_h, p.rows_px, y0, y1);
  const uint32_t R = (uint32_t)p.src_w * 3u;
  const int nchunk = (int)(R >> 4);
  const int nrow = y1 - y0;
  if (p.flip_y) y0 = p.src_h - y1; // the band of the mirrored image is the mirrored band (sums do not care about order)
  const uint8_t *band = frame + (size_t)y0 * (size_t)R;
  if (FILT) band_sums_filtered<NT>(band, R, nrow, V);
  for (int c = threadIdx.x; c < (FILT ? 0 : nchunk); c += NT) {
    uint32_t a[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint8_t *q = band + ((size_t)c << 4);
    // whole batches of kBandUnroll rows, then one batch of exactly the rows that are left (a typical band — 11 or 12
    // rows at 4K -> 192 pixel rows — is a single batch: one memory latency per column)
    int r = nrow;
    for (; r >= kBandUnroll; r -= kBandUnroll) band_trip<kBandUnroll>(a, q, R);
    switch (r) {
    case 11: band_trip<11>(a, q, R); break;
    case 10: band_trip<10>(a, q, R); break;
    case 9: band_trip<9>(a, q, R); break;
    case 8: band_trip<8>(a, q, R); break;
    case 7: band_trip<7>(a, q, R); break;
    case 6: band_trip<6>(a, q, R); break;
    case 5: band_trip<5>(a, q, R); break;
    case 4: band_trip<4>(a, q, R); break;
    case 3: band_trip<3>(a, q, R); break;
    case 2: band_trip<2>(a, q, R); break;
    case 1: band_trip<1>(a, q, R); break;
    default: break;
    }
    // byte j of word k is column 16c + 4k + j:  even reg = {b0 | b2<<16}, odd reg = {b1 | b3<<16}
    uint4 lo, hi;
    lo.x = __byte_perm(a[0], a[1], 0x5410); // b0,b1
    lo.y = __byte_perm(a[0], a[1], 0x7632); // b2,b3
    lo.z = __byte_perm(a[2], a[3], 0x5410);
    lo.w = __byte_perm(a[2], a[3], 0x7632);
    hi.x = __byte_perm(a[4], a[5], 0x5410);
    hi.y = __byte_perm(a[4], a[5], 0x7632);
    hi.z = __byte_perm(a[6], a[7], 0x5410);
    hi.w = __byte_perm(a[6], a[7], 0x7632);
    uint4 *dst = reinterpret_cast<uint4 *>(V + (size_t)c * 16);
    dst[0] = lo;
    dst[1] = hi;
  }
  Sync::sync();
  // Horizontal sums.  Uniform even box width (src_w = cols * bx, bx even — every BASELINE shape): a cell's column
  // sums are bx/2 pixel pairs of three 32-bit words {r0|g0, b0|r1, g1|b1}, added as packed u16 lanes (3 adds per
  // pair instead of 6 loads + 6 adds), and the three divisions by the uniform box area become multiply-highs with
  // the exact reciprocal M = ceil(2^32 / n)  (exact for (s + h) * n < 2^32; here s + h < 256 n and n < 4096).
  // (box_bx / box_nrow0 / box_M[] come from the host: no per-thread divisions here.)
  const int bx = p.box_bx;
  const uint32_t dn = (uint32_t)(nrow - p.box_nrow0);
  if (bx > 0 && dn < 2u && !(p.tune_flags & 8)) {
    const uint32_t n = (uint32_t)bx * (uint32_t)nrow, h = n >> 1;
    const uint32_t M = p.box_M[dn];
    const uint32_t *Vw = reinterpret_cast<const uint32_t *>(V);
    const int half = bx >> 1;
    for (int x = threadIdx.x; x < p.cols; x += NT) {
      const uint32_t *q = Vw + 3 * half * (p.flip_x ? p.cols - 1 - x : x);
      uint32_t a0 = 0, a1 = 0, a2 = 0;
#pragma unroll 6
      for (int k = 0; k < half; k++, q += 3) {
        a0 += q[0];
        a1 += q[1];
        a2 += q[2];
      }
      const uint32_t sr = (a0 & 0xFFFFu) + (a1 >> 16), sg = (a0 >> 16) + (a2 & 0xFFFFu), sb = (a1 & 0xFFFFu) + (a2 >> 16);
      out[x] = (__umulhi(sr + h, M) << 16) | (__umulhi(sg + h, M) << 8) | __umulhi(sb + h, M);
    }
  } else {
    for (int x = threadIdx.x; x < p.cols; x += NT) {
      int x0, x1;
      box_range(x, p.src_w, p.cols, x0, x1);
      uint32_t sr = 0, sg = 0, sb = 0;
      const uint16_t *q = V + 3 * (p.flip_x ? p.src_w - x1 : x0);
#pragma unroll 4
      for (int xx = x0; xx < x1; xx++, q += 3) {
        sr += q[0];
        sg += q[1];
        sb += q[2];
      }
      uint32_t n = (uint32_t)(x1 - x0) * (uint32_t)nrow, h = n >> 1;
      out[x] = (((sr + h) / n) << 16) | (((sg + h) / n) << 8) | ((sb + h) / n);
    }
  }
  Sync::sync(); // V is reused by the next pixel row
}

// ------------------------------------------------------------------ phase B: per-cell emission
struct RowCtx {
  const GlyphLut *lut;
  const uint32_t *cT, *cB;
  const uint16_t *key, *hpos, *rend;
  uint32_t drop_first; // EM_TRUE_FG, direct output: the row's first ASCII cell inherits the colour state of the rows above
  uint32_t fg_over;    // 0, or 0x01RRGGBB: colour printed by every truecolor-foreground SGR (rainbow replace)
};
__device__ __forceinline__ uint32_t fg_print(uint32_t px, uint32_t over) { return over ? over : px; }

template <int MODE, class S> __device__ __forceinline__ void emit_cell(S &s, int x, const RowCtx &c) {
  if (MODE == EM_256_FG) {
    uint32_t px = c.cT[x];
    put_sgr_256_l(s, false, (uint32_t)q256_of(px));
    put_glyph(s, c.lut->glyph[luma_of(px)]);
  } else if (MODE == EM_16_FG) {
    uint32_t px = c.cT[x];
    put_sgr_16_l(s, false, (uint32_t)q16_of(px));
    put_glyph(s, c.lut->glyph[luma_of(px)]);
  } else if (MODE == EM_TRUE_FG) {
    // ansi_rle_add_pixel (ansi.c:261-300) as a cell rule; hpos[x] = previous ASCII-glyph cell of this row
    uint32_t px = c.cT[x];
    const uint8_t *g = c.lut->glyph[luma_of(px)];
    bool ascii = g[0] == 1 && g[1] < 128;
    if (ascii) {
      uint16_t pa = c.hpos[x];
      if (pa == NONE16 ? !c.drop_first : c.cT[pa] != px) put_sgr_rgb_l(s, false, fg_print(px, c.fg_over));
      s.put(g[1]);
    } else {
      put_sgr_rgb_l(s, false, fg_print(px, c.fg_over));
      put_glyph(s, g);
    }
  } else if (MODE == EM_MONO_FG) {
    int h = c.hpos[x];
    uint32_t run = (uint32_t)c.rend[h] - (uint32_t)h;
    bool rep = rep_profitable(run);
    const uint8_t *g = c.lut->glyph[luma_of(c.cT[x])]; // same glyph for every cell of the run (same key)
    if (h == x) {
      put_glyph(s, g);
      if (rep) put_rep(s, run - 1);
    } else if (!rep) {
      put_glyph(s, g);
    }
  } else if (MODE == EM_HB_MONO) {
    int h = c.hpos[x];
    uint32_t run = (uint32_t)c.rend[h] - (uint32_t)h;
    bool rep = rep_profitable(run);
    int lt = luma76_of(c.cT[h]), lb = luma76_of(c.cB[h]);
    if (lt < 16 && lb < 16) {
      s.put(' ');
    } else if (h == x || !rep) {
      const uint8_t sh = (uint8_t)((lt >> 6) == 0 ? 0x91 : (lt >> 6) == 1 ? 0x92 : (lt >> 6) == 2 ? 0x93 : 0x88);
      put3(s, 0xE2, 0x96, sh);
      if (h == x && rep) put_rep(s, run - 1);
    }
  } else { // EM_HB_TRUE / EM_HB_256 / EM_HB_16
    int h = c.hpos[x];
    uint32_t run = (uint32_t)c.rend[h] - (uint32_t)h;
    bool rep = rep_profitable(run);
    uint32_t tH = c.cT[h], bH = c.cB[h];
    bool prev_set = false;
    int ph = 0;
    if (h > 0) {
      ph = c.hpos[h - 1];
      prev_set = (c.cT[ph] | c.cB[ph]) != 0u; // a transparent run leaves the colour state cleared
    }
    if ((tH | bH) == 0u) { // transparent: decided by the run head's raw RGB (halfblock.c:111,357,476)
      if (h == x && prev_set) put_reset(s);
      s.put(' ');
    } else if (h == x) {
      if (MODE == EM_HB_TRUE) {
        if (!prev_set || c.cT[ph] != tH) put_sgr_rgb_l(s, false, fg_print(tH, c.fg_over));
        if (!prev_set || c.cB[ph] != bH) put_sgr_rgb_l(s, true, bH);
      } else {
        uint32_t k = c.key[h], pk = prev_set ? c.key[ph] : 0u;
        if (MODE == EM_HB_256) {
          if (!prev_set || (pk >> 8) != (k >> 8)) put_sgr_256_l(s, false, k >> 8);
          if (!prev_set || (pk & 255u) != (k & 255u)) put_sgr_256_l(s, true, k & 255u);
        } else {
          if (!prev_set || (pk >> 8) != (k >> 8)) put_sgr_16_l(s, false, k >> 8);
          if (!prev_set || (pk & 255u) != (k & 255u)) put_sgr_16_l(s, true, k & 255u);
        }
      }
      put3(s, 0xE2, 0x96, 0x80);
      if (rep) put_rep(s, run - 1);
    } else if (!rep) {
      put3(s, 0xE2, 0x96, 0x80);
    }
  }
}

// Phase B for one text row whose resized pixels are in cT/cB: keys -> runs -> byte counts -> offsets -> bytes,
// staged in shared memory (or written straight to the scratch row when it is too wide) and copied out.
// emit_prepare = B1..B3 (returns the byte count of the cells, offsets in off[]); emit_row adds B4 into the scratch row.
template <int MODE, class Sync, int NT, class OffT = uint32_t>
__device__ __forceinline__ uint32_t emit_prepare(const RenderParams &p, GlyphLut *lut, uint32_t *cT, uint32_t *cB,
                                                 uint16_t *key, uint16_t *hpos, uint16_t *rend, OffT *off,
                                                 int *s_tmp, uint32_t *s_cond, int tid) {
  constexpr bool HB = MODE >= EM_HB_TRUE && MODE <= EM_HB_MONO;
  constexpr bool RUNS = MODE == EM_MONO_FG || HB;
  const int w = p.cols;
  if (tid < 4) s_cond[tid] = 0u;
  Sync::sync();

  // ---- phase B1: run keys
  if (MODE == EM_MONO_FG) {
    for (int x = tid; x < w; x += NT) key[x] = lut->key[luma_of(cT[x])];
  } else if (MODE == EM_HB_256) {
    for (int x = tid; x < w; x += NT) key[x] = (uint16_t)((q256_of(cT[x]) << 8) | q256_of(cB[x]));
  } else if (MODE == EM_HB_16) {
    for (int x = tid; x < w; x += NT) key[x] = (uint16_t)((q16_of(cT[x]) << 8) | q16_of(cB[x]));
  }
  Sync::sync();

  // ---- phase B2: run heads / previous-ASCII links
  if (RUNS) {
    auto is_head = [&](int x) -> bool {
      if (x == 0) return true;
      if (MODE == EM_HB_TRUE || MODE == EM_HB_MONO) return cT[x] != cT[x - 1] || cB[x] != cB[x - 1];
      return key[x] != key[x - 1];
    };
    row_scan<OpMax, Sync, NT>(
        w, [&](int x) { return is_head(x) ? x : -1; }, [&](int x, int incl, int) { hpos[x] = (uint16_t)incl; },
        s_tmp, tid);
    Sync::sync();
    for (int x = tid; x < w; x += NT) {
      if (x > 0 && hpos[x] == x) rend[hpos[x - 1]] = (uint16_t)x; // this head closes the previous run
      if (x == w - 1) rend[hpos[x]] = (uint16_t)w;
    }
    Sync::sync();
  } else if (MODE == EM_TRUE_FG) {
    int last_ascii = row_scan<OpMax, Sync, NT>(
        w,
        [&](int x) {
          const uint8_t *g = lut->glyph[luma_of(cT[x])];
          return (g[0] == 1 && g[1] < 128) ? x : -1;
        },
        [&](int x, int, int excl) { hpos[x] = excl < 0 ? NONE16 : (uint16_t)excl; }, s_tmp, tid);
    if (tid == 0) s_cond[2] = last_ascii >= 0 ? (0x01000000u | cT[last_ascii]) : 0u;
  }

  // ---- phase B3: byte counts -> offsets
  RowCtx ctx{lut, cT, cB, key, hpos, rend, 0u, p.fg_over};
  int cells_bytes = row_scan<OpAdd, Sync, NT>(
      w,
      [&](int x) {
        CountSink cs;
        emit_cell<MODE>(cs, x, ctx);
        return (int)cs.n;
      },
      [&](int x, int, int excl) { off[x] = (OffT)((uint32_t)excl + (uint32_t)p.pad_left); }, s_tmp, tid);
  return (uint32_t)cells_bytes;
}

template <int MODE, class Sync, int NT>
__device__ __forceinline__ void emit_row(const RenderParams &p, int f, int t, GlyphLut *lut, uint32_t *cT, uint32_t *cB,
                                         uint16_t *key, uint16_t *hpos, uint16_t *rend, uint32_t *off, uint8_t *outb,
                                         int *s_tmp, uint32_t *s_cond, int tid) {
  const int w = p.cols;
  const bool last_row = t == p.text_rows - 1;
  const uint32_t cells_bytes = emit_prepare<MODE, Sync, NT>(p, lut, cT, cB, key, hpos, rend, off, s_tmp, s_cond, tid);
  RowCtx ctx{lut, cT, cB, key, hpos, rend, 0u, p.fg_over};
  const uint32_t body_end = (uint32_t)p.pad_left + cells_bytes;

  // ---- phase B4: materialise
  uint8_t *grow = p.rows + ((size_t)f * p.text_rows + t) * (size_t)p.row_pitch;
  uint8_t *base = p.use_smem_out ? outb : grow;
  for (int i = tid; i < p.pad_left; i += NT) base[i] = ' ';
  for (int x = tid; x < w; x += NT) {
    if (p.use_smem_out) {
      SmemSink ss{(uint32_t)__cvta_generic_to_shared(outb) + off[x]};
      emit_cell<MODE>(ss, x, ctx);
    } else {
      WriteSink ws{base + off[x]};
      emit_cell<MODE>(ws, x, ctx);
    }
    if (MODE == EM_TRUE_FG && hpos[x] == NONE16) {
      const uint8_t *g = lut->glyph[luma_of(cT[x])];
      if (g[0] == 1 && g[1] < 128) { // the row's first ASCII-glyph cell: its SGR is conditional on the row above
        s_cond[0] = off[x];
        const uint32_t pc = fg_print(cT[x], p.fg_over);
        s_cond[1] = 10u + (s_dec3[(pc >> 16) & 255u] >> 24) + (s_dec3[(pc >> 8) & 255u] >> 24) + (s_dec3[pc & 255u] >> 24);
        s_cond[3] = 0x01000000u | cT[x];
      }
    }
  }
  uint32_t len = body_end;
  if (tid == 0) {
    WriteSink ws{base + body_end};
    if (MODE == EM_256_FG || MODE == EM_16_FG || MODE == EM_HB_TRUE || MODE == EM_HB_256 || MODE == EM_HB_16)
      put_reset(ws);
    if (MODE == EM_TRUE_FG && last_row) put_reset(ws); // ansi_rle_finish, ansi.c:303-314
    if (!last_row) ws.put('\n');
    len = (uint32_t)(ws.p - base);
  }
  Sync::sync();
  if (tid == 0) {
    RowMeta m;
    m.len = len;
    m.cond_off = s_cond[0];
    m.cond_len = s_cond[1];
    m.first_rgb = s_cond[3];
    m.last_rgb = s_cond[2];
    m._pad[0] = m._pad[1] = m._pad[2] = 0;
    p.meta[(size_t)f * p.text_rows + t] = m;
    s_cond[0] = len;
  }
  Sync::sync();
  if (p.use_smem_out) {
    const uint32_t n16 = (s_cond[0] + 4u + 15u) >> 4; // k_stitch reads up to one word past the row (copy_shifted)
    const uint4 *s4 = reinterpret_cast<const uint4 *>(outb);
    uint4 *d4 = reinterpret_cast<uint4 *>(grow);
    for (uint32_t i = tid; i < n16; i += NT) d4[i] = s4[i];
  }
}

// ------------------------------------------------------------------ direct output: row-length look-back
// One warp finishes a text row AND places it in the final frame string, so no stitch pass and no scratch rows:
// after the byte counts are known the warp publishes {ready, len, first_rgb, last_rgb} for its row, reads the
// records of the rows above (spinning until each is ready — they are being produced concurrently by other CTAs, and
// tiles are handed out by an atomic ticket so every lower tile is already running), and derives its byte offset in
// the frame.  For truecolor-foreground the same look-back carries the colour state of ansi_rle_add_pixel across
// rows (ansi.c:263): each row's first ASCII-cell SGR is dropped iff its colour equals the last ASCII colour above.
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// A look-back record is ONE aligned 16-byte word {ready = epoch, len, first_rgb, last_rgb}, published with a single
// 16-byte store and polled with a single 16-byte load: the memory system moves it as one transaction (the same
// single-word protocol CUB's decoupled look-back uses for 16-byte tile descriptors), so whoever sees the new epoch sees
// the payload that was stored with it — no fences on either side, one load per poll instead of four.
__device__ __forceinline__ uint4 ld_record(const uint4 *p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_record(uint4 *p, const uint4 &v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint32_t sgr_rgb_len(uint32_t c) { // bytes of ESC[38;2;R;G;Bm
  uint32_t r = (c >> 16) & 255u, g = (c >> 8) & 255u, b = c & 255u;
  return 10u + (r >= 100u ? 3u : r >= 10u ? 2u : 1u) + (g >= 100u ? 3u : g >= 10u ? 2u : 1u) +
         (b >= 100u ? 3u : b >= 10u ? 2u : 1u);
}

// prepare: B1..B3 for the row, then publish its record; never waits.  Returns the byte count of the cells.
// NT threads (tid 0..NT-1) own the row: one warp in the role-split kernel, the whole CTA in the one-tile kernels.
template <int MODE, class Sync, int NT>
__device__ __forceinline__ uint32_t emit_direct_prepare(const RenderParams &p, int f, int t, GlyphLut *lut, uint32_t *cT,
                                                        uint32_t *cB, uint16_t *key, uint16_t *hpos, uint16_t *rend,
                                                        uint16_t *off, int *s_tmp, uint32_t *s_cond, int tid) {
  const int w = p.cols;
  const bool last_row = t == p.text_rows - 1;
  const uint32_t cells_bytes =
      emit_prepare<MODE, Sync, NT, uint16_t>(p, lut, cT, cB, key, hpos, rend, off, s_tmp, s_cond, tid);
  if (MODE == EM_TRUE_FG) { // locate the row's first ASCII-glyph cell (s_cond[2] = last one, from the scan)
    for (int x = tid; x < w; x += NT)
      if (hpos[x] == NONE16) {
        const uint8_t *g = lut->glyph[luma_of(cT[x])];
        if (g[0] == 1 && g[1] < 128) {
          s_cond[0] = (uint32_t)x;
          s_cond[3] = 0x01000000u | cT[x];
        }
      }
    Sync::sync();
  }
  constexpr bool row_reset = MODE == EM_256_FG || MODE == EM_16_FG || MODE == EM_HB_TRUE || MODE == EM_HB_256 || MODE == EM_HB_16;
  const uint32_t term_len = (row_reset ? 4u : 0u) + ((MODE == EM_TRUE_FG && last_row) ? 4u : 0u) + (last_row ? 0u : 1u);
  if (tid == 0) {
    // "ready" = this launch's epoch: no clearing between launches
    st_record(p.agg + (size_t)f * p.text_rows + t,
              make_uint4(p.epoch, (uint32_t)p.pad_left + cells_bytes + term_len, MODE == EM_TRUE_FG ? s_cond[3] : 0u,
                         MODE == EM_TRUE_FG ? s_cond[2] : 0u));
  }
  return cells_bytes;
}

// finish: look-back for the row's offset (and, truecolor-fg, the colour state above), materialise, store.
// The look-back is done by ONE warp (lane = 0..31 of it); s_lb (2 words of shared memory) carries its result.
template <int MODE>
__device__ __forceinline__ void row_lookback(const RenderParams &p, int f, int t, int lane, uint32_t *s_lb) {
  const uint4 *agg = p.agg + (size_t)f * p.text_rows;
  uint32_t prefix = 0, carry = 0;
  // the records of up to 96 rows above are requested in one go (three independent loads per lane in flight), then
  // only the ones that were not published yet are polled again: one L2 round trip in the common case
  for (int base0 = 0; base0 < t; base0 += 96) {
    uint4 rec[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int j = base0 + 32 * k + lane;
      rec[k] = j < t ? ld_record(agg + j) : make_uint4(p.epoch, 0u, 0u, 0u);
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int j = base0 + 32 * k + lane;
      while (rec[k].x != p.epoch) {
        __nanosleep(32);
        rec[k] = ld_record(agg + j);
      }
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (base0 + 32 * k >= t) break;
      uint32_t len = rec[k].y, fj = rec[k].z, lj = rec[k].w;
      if (MODE == EM_TRUE_FG) {
        uint32_t inc = lj; // inclusive "last non-zero" scan = colour state after row j
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
          if (lane >= d && inc == 0u) inc = o;
        }
        uint32_t before = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) before = 0u;
        if (before == 0u) before = carry;
        if (fj && before && fj == before) len -= sgr_rgb_len(fg_print(fj, p.fg_over));
        const uint32_t tail = __shfl_sync(0xffffffffu, inc, 31);
        if (tail) carry = tail;
      }
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) len += __shfl_xor_sync(0xffffffffu, len, d);
      prefix += len;
    }
  }
  if (lane == 0) {
    s_lb[0] = prefix;
    s_lb[1] = carry;
  }
}

// B4: the row's bytes, already in their final form, into shared memory at sb; NTB threads (tid 0..NTB-1) share the cells
template <int MODE, int NTB>
__device__ __forceinline__ void row_materialise(const RenderParams &p, const RowCtx &ctx, const uint16_t *off, uint8_t *sb,
                                                uint32_t body_end, uint32_t drop, uint32_t first_x, bool last_row, int tid) {
  constexpr bool row_reset = MODE == EM_256_FG || MODE == EM_16_FG || MODE == EM_HB_TRUE || MODE == EM_HB_256 || MODE == EM_HB_16;
  const int w = p.cols;
  for (int i = tid; i < p.pad_left; i += NTB) sb[i] = ' ';
  const uint32_t sb32 = (uint32_t)__cvta_generic_to_shared(sb);
  for (int x = tid; x < w; x += NTB) {
    const uint32_t o = (uint32_t)off[x] - ((MODE == EM_TRUE_FG && drop && (uint32_t)x > first_x) ? drop : 0u);
    SmemSink ss{sb32 + o};
    emit_cell<MODE>(ss, x, ctx);
  }
  if (tid == 0) {
    WriteSink ws{sb + body_end - drop};
    if (row_reset) put_reset(ws);
    if (MODE == EM_TRUE_FG && last_row) put_reset(ws); // ansi_rle_finish, ansi.c:303-314
    if (!last_row) ws.put('\n');
  }
}

template <int MODE, class Sync, int NT>
__device__ __forceinline__ void emit_direct_finish(const RenderParams &p, int f, int t, GlyphLut *lut, uint32_t *cT,
                                                   uint32_t *cB, uint16_t *key, uint16_t *hpos, uint16_t *rend,
                                                   const uint16_t *off, uint8_t *outb, const uint32_t *s_cond,
                                                   uint32_t *s_lb, uint32_t cells_bytes, int tid) {
  const bool last_row = t == p.text_rows - 1;
  constexpr bool row_reset = MODE == EM_256_FG || MODE == EM_16_FG || MODE == EM_HB_TRUE || MODE == EM_HB_256 || MODE == EM_HB_16;
  const uint32_t term_len = (row_reset ? 4u : 0u) + ((MODE == EM_TRUE_FG && last_row) ? 4u : 0u) + (last_row ? 0u : 1u);
  const uint32_t body_end = (uint32_t)p.pad_left + cells_bytes;
  const uint32_t row_len = body_end + term_len;
  const uint32_t first = MODE == EM_TRUE_FG ? s_cond[3] : 0u;
  const uint32_t first_x = s_cond[0];
  uint8_t *frame_out = p.out + (size_t)f * p.out_pitch;
  // Several warps, and nothing in the row's bytes depends on the rows above (every mode but truecolor-fg, whose first
  // SGR may be dropped): the last warp polls the rows above WHILE the others materialise the row at a fixed phase of
  // the staging buffer; the copy-out then realigns to the destination (two LDS.128 + funnel shifts per 16 bytes).
  // One warp (the role-split kernel's emitter) or truecolor-fg: look-back first, then materialise with the
  // destination's 16-byte phase, so that the copy-out is plain vector moves.
  constexpr bool PAR = NT >= 64 && MODE != EM_TRUE_FG;

  if (PAR) {
    RowCtx ctx{lut, cT, cB, key, hpos, rend, 0u, p.fg_over};
    if (tid >= NT - 32) row_lookback<MODE>(p, f, t, tid - (NT - 32), s_lb);
    else row_materialise<MODE, (NT >= 64 ? NT - 32 : 32)>(p, ctx, off, outb, body_end, 0u, first_x, last_row, tid);
    Sync::sync();
    const uint32_t dst_off = (uint32_t)p.pad_top + s_lb[0];
    uint8_t *dst = frame_out + dst_off;
    uint32_t head = (16u - (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u)) & 15u;
    if (head > row_len) head = row_len;
    if ((uint32_t)tid < head) dst[tid] = outb[tid];
    const uint32_t nvec = (row_len - head) >> 4;
    uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
    const uint4 *s4 = reinterpret_cast<const uint4 *>(outb);
    const uint32_t q = head >> 2, ksh = (head & 3u) * 8u; // vector i starts at word 4i + q, byte phase head & 3
    {
      const uint64_t keep = l2_policy_evict_last();
      for (uint32_t i = tid; i < nvec; i += NT) {
        const uint4 A = s4[i], B = s4[i + 1]; // words 4i .. 4i+7 (the staging buffer keeps 16 spare bytes behind the row)
        uint32_t w0, w1, w2, w3, w4;
        switch (q) { // uniform across the CTA
        case 0: w0 = A.x, w1 = A.y, w2 = A.z, w3 = A.w, w4 = B.x; break;
        case 1: w0 = A.y, w1 = A.z, w2 = A.w, w3 = B.x, w4 = B.y; break;
        case 2: w0 = A.z, w1 = A.w, w2 = B.x, w3 = B.y, w4 = B.z; break;
        default: w0 = A.w, w1 = B.x, w2 = B.y, w3 = B.z, w4 = B.w; break;
        }
        stg_keep(d4 + i,
                 make_uint4(__funnelshift_r(w0, w1, ksh), __funnelshift_r(w1, w2, ksh), __funnelshift_r(w2, w3, ksh),
                            __funnelshift_r(w3, w4, ksh)),
                 keep);
      }
    }
    const uint32_t done = head + (nvec << 4);
    if ((uint32_t)tid < row_len - done) dst[done + tid] = outb[done + tid];
    if (t == 0)
      for (int i = tid; i < p.pad_top; i += NT) frame_out[i] = '\n';
    if (last_row && tid == 0) {
      frame_out[dst_off + row_len] = 0;
      p.out_len[f] = dst_off + row_len;
    }
    Sync::sync(); // the staging buffer and s_lb are reused by the next row
    return;
  }

  if (tid < 32) row_lookback<MODE>(p, f, t, tid, s_lb);
  Sync::sync();
  const uint32_t prefix = s_lb[0], carry = s_lb[1];
  const uint32_t drop =
      (MODE == EM_TRUE_FG && first && carry && first == carry) ? sgr_rgb_len(fg_print(first, p.fg_over)) : 0u;
  const uint32_t dst_off = (uint32_t)p.pad_top + prefix;
  const uint32_t final_len = row_len - drop;
  uint8_t *dst = frame_out + dst_off;
  const uint32_t shift = (uint32_t)(reinterpret_cast<uintptr_t>(dst) & 15u); // stage with the destination's alignment
  uint8_t *sb = outb + shift;

  RowCtx ctx{lut, cT, cB, key, hpos, rend, drop ? 1u : 0u, p.fg_over};
  row_materialise<MODE, NT>(p, ctx, off, sb, body_end, drop, first_x, last_row, tid);
  Sync::sync();

  // ---- copy out: unaligned head bytes, 16-byte body (source and destination share their alignment), tail bytes
  uint32_t head = (16u - shift) & 15u;
  if (head > final_len) head = final_len;
  if ((uint32_t)tid < head) dst[tid] = sb[tid];
  const uint32_t nvec = (final_len - head) >> 4;
  const uint4 *s4 = reinterpret_cast<const uint4 *>(sb + head);
  uint4 *d4 = reinterpret_cast<uint4 *>(dst + head);
  {
    const uint64_t keep = l2_policy_evict_last();
    for (uint32_t i = tid; i < nvec; i += NT) stg_keep(d4 + i, s4[i], keep);
  }
  const uint32_t done = head + (nvec << 4);
  if ((uint32_t)tid < final_len - done) dst[done + tid] = sb[done + tid];
  if (t == 0)
    for (int i = tid; i < p.pad_top; i += NT) frame_out[i] = '\n';
  if (last_row && tid == 0) {
    frame_out[dst_off + final_len] = 0;
    p.out_len[f] = dst_off + final_len;
  }
  Sync::sync(); // the staging buffer and s_lb are reused by the next row
}

// Emission-bound variants (NN sampling) want occupancy: eight 256-thread CTAs per SM = 32 registers per thread.
//
// Scratch-row output (p.direct == 0): one CTA per (frame, text row), tile = blockIdx.x.
// Direct output (p.direct == 1): PERSISTENT CTAs (grid = resident CTAs, launch_rows_t) that draw tiles from the atomic
// ticket until it runs past the end — the glyph LUT copy and the decimal table are built once per CTA lifetime, not once
// per text row, and the next tile's ticket is requested while the current one is being emitted.  Whoever holds tile X
// knows every tile < X is held by a running CTA, so the look-back spin cannot deadlock.
template <int MODE, int SP, int NT>
__global__ void __launch_bounds__(NT, SP != SP_NN ? 1 : NT == 160 ? 9 : NT == 320 ? 6 : NT <= 256 ? 2048 / NT : 1)
    k_render_rows(const RenderParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_tmp[2 * (NT / 32)];
  __shared__ uint32_t s_cond2[2][4]; // per tile set; TRUE_FG: {cond_off, cond_len, last_rgb, first_rgb}

  constexpr bool HB = MODE >= EM_HB_TRUE && MODE <= EM_HB_MONO;
  constexpr bool USES_LUT = MODE <= EM_TRUE_FG;

  const int tid = threadIdx.x;
  __shared__ int s_tile;
  __shared__ uint32_t s_lb[2];
  const int w = p.cols;
  const unsigned total = (unsigned)p.n_frames * (unsigned)p.text_rows;

  const uint32_t cap = p.use_smem_out ? p.row_pitch : 0u;
  const Layout L = make_layout(MODE, SP, w, p.src_w, cap, p.tune_flags & 1, p.direct);
  GlyphLut *lut = reinterpret_cast<GlyphLut *>(smem + L.lut);
  uint16_t *V = reinterpret_cast<uint16_t *>(smem + L.V);
  uint8_t *outb = smem + L.outb;

  if (USES_LUT) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.lut);
    uint32_t *dst = reinterpret_cast<uint32_t *>(lut);
    for (int i = tid; i < (int)(sizeof(GlyphLut) / 4); i += NT) dst[i] = src[i];
  }
  init_dec3<NT>(tid);
  __syncthreads();

  // Direct output is software-pipelined by one tile, like the role-split kernel's emitter: iteration k samples,
  // prepares and PUBLISHES tile k (never waits), then looks back and writes out tile k-1 — whose predecessors have had a
  // whole tile period to publish, so the look-back finds them ready.  The per-tile arrays exist twice (set = k & 1).
  int prev_tile = -1, set = 0;
  uint32_t prev_bytes = 0;
#pragma unroll 1
  for (;;) {
    // The ticket is drawn when the CTA is ready to start the tile, NOT ahead of time: a tile that is held but not yet
    // started keeps every later row of its frame spinning in the look-back for a whole tile period (measured: a
    // prefetched ticket made this kernel 3x slower, profiles/r02a_nn_*: 63 barrier-stall cycles per issue).
    if (p.direct) {
      if (tid == 0) s_tile = (int)((uint32_t)atomicAdd(p.ticket, 1) - p.ticket_base);
      __syncthreads();
    }
    const unsigned tile = p.direct ? (unsigned)s_tile : blockIdx.x;
    const bool have = tile < total;
    uint32_t bytes = 0;
    if (have) {
      const uint32_t so = set ? L.set2 : 0u;
      uint32_t *cT = reinterpret_cast<uint32_t *>(smem + L.cT + so);
      uint32_t *cB = reinterpret_cast<uint32_t *>(smem + L.cB + so);
      uint16_t *key = reinterpret_cast<uint16_t *>(smem + L.key + so);
      uint16_t *hpos = reinterpret_cast<uint16_t *>(smem + L.hpos + so);
      uint16_t *rend = reinterpret_cast<uint16_t *>(smem + L.rend + so);
      uint32_t *off = reinterpret_cast<uint32_t *>(smem + L.off + so);
      const int t = (int)(tile % (unsigned)p.text_rows);
      const int f = (int)(tile / (unsigned)p.text_rows);

      // ---- phase A
      const uint8_t *frame = p.frames + (size_t)f * p.frame_stride;
      const int yT = HB ? 2 * t : t;
      const bool hasB = HB && (2 * t + 1 < p.rows_px);
      if (SP == SP_NN) {
        cells_nn<NT>(p, frame, yT, cT, hasB ? cB : nullptr);
      } else if (SP == SP_BOX_GENERIC) {
        cells_box_generic<NT>(p, frame, yT, cT);
        if (hasB) cells_box_generic<NT>(p, frame, yT + 1, cB);
      } else {
#pragma unroll 1
        for (int hrow = 0; hrow < (hasB ? 2 : 1); hrow++) cells_box_stream<NT>(p, frame, yT + hrow, hrow ? cB : cT, V);
      }
      __syncthreads();
      if (HB && !hasB) { // odd pixel height: bottom := top (halfblock.c:73,82-88)
        for (int x = tid; x < w; x += NT) cB[x] = cT[x];
        __syncthreads();
      }
      if (p.cells_out) {
        uint8_t *co = p.cells_out + ((size_t)f * p.rows_px + yT) * (size_t)w * 3u;
        for (int x = tid; x < w; x += NT) {
          uint32_t c = cT[x];
          co[3 * x] = (uint8_t)(c >> 16);
          co[3 * x + 1] = (uint8_t)(c >> 8);
          co[3 * x + 2] = (uint8_t)c;
          if (hasB) {
            uint32_t d = cB[x];
            uint8_t *cb = co + (size_t)w * 3u;
            cb[3 * x] = (uint8_t)(d >> 16);
            cb[3 * x + 1] = (uint8_t)(d >> 8);
            cb[3 * x + 2] = (uint8_t)d;
          }
        }
      }
      if (p.rows == nullptr) return; // resize-only invocation (never direct)

      if (!p.direct) {
        emit_row<MODE, SyncAll, NT>(p, f, t, lut, cT, cB, key, hpos, rend, off, outb, s_tmp, s_cond2[0], tid);
        return;
      }
      bytes = emit_direct_prepare<MODE, SyncAll, NT>(p, f, t, lut, cT, cB, key, hpos, rend,
                                                     reinterpret_cast<uint16_t *>(off), s_tmp, s_cond2[set], tid);
    } else if (!p.direct) {
      return;
    }
    if (prev_tile >= 0) {
      const uint32_t so = set ? 0u : L.set2; // the other set
      const int t = (int)((unsigned)prev_tile % (unsigned)p.text_rows);
      const int f = (int)((unsigned)prev_tile / (unsigned)p.text_rows);
      emit_direct_finish<MODE, SyncAll, NT>(
          p, f, t, lut, reinterpret_cast<uint32_t *>(smem + L.cT + so), reinterpret_cast<uint32_t *>(smem + L.cB + so),
          reinterpret_cast<uint16_t *>(smem + L.key + so), reinterpret_cast<uint16_t *>(smem + L.hpos + so),
          reinterpret_cast<uint16_t *>(smem + L.rend + so), reinterpret_cast<const uint16_t *>(smem + L.off + so), outb,
          s_cond2[set ^ 1], s_lb, prev_bytes, tid);
      // emit_direct_finish ends on a barrier: that set, the staging buffer and s_tile are free again
    }
    if (!have) return;
    prev_tile = (int)tile;
    prev_bytes = bytes;
    set ^= 1;
  }
}

// ------------------------------------------------------------------ warp-specialised persistent row kernel
// Box-filter streaming with the copy engine instead of the LSU: a producer warp keeps a ring of whole source rows
// in flight with 1-D bulk TMA copies (cp.async.bulk, completion on an mbarrier), eight consumer warps sum the rows
// out of shared memory and then emit the text row.  CTAs are persistent (tile = blockIdx.x + k*gridDim.x over
// (frame, text row)), so while the consumers are in the emission phase of tile k the producer is already filling
// the ring with the first rows of tile k+1: HBM traffic never pauses for the byte-emission work.
//
//   full[s]  : producer arms it with expect_tx(row bytes); the bulk copy completes it        (count 1 + tx)
//   empty[s] : one arrival per consumer warp when the warp has read slot s                   (count 8)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

constexpr int WS_MAXD = 16;            // ring slots (source rows) at most

struct WsBand {
  int a0, a1, b1; // top pixel row sums source rows [a0,a1), bottom [a1,b1)  (contiguous when downscaling)
  bool hasB;
};
template <bool HB> __device__ __forceinline__ WsBand ws_band(const RenderParams &p, int t) {
  WsBand b;
  const int yT = HB ? 2 * t : t;
  box_range(yT, p.src_h, p.rows_px, b.a0, b.a1);
  b.hasB = HB && (2 * t + 1 < p.rows_px);
  b.b1 = b.a1;
  if (b.hasB) {
    int b0;
    box_range(yT + 1, p.src_h, p.rows_px, b0, b.b1);
  }
  return b;
}

template <int MODE, int CPT, int NT>
__global__ void __launch_bounds__(NT + 32) k_render_rows_ws(const RenderParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_tmp[2 * (NT / 32)];
  __shared__ uint32_t s_cond[4];
  __shared__ __align__(8) uint64_t s_full[WS_MAXD], s_empty[WS_MAXD];

  constexpr bool HB = MODE >= EM_HB_TRUE && MODE <= EM_HB_MONO;
  constexpr bool USES_LUT = MODE <= EM_TRUE_FG;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int w = p.cols;
  const uint32_t R = (uint32_t)p.src_w * 3u;
  const int nchunk = (int)(R >> 4);
  const uint32_t D = (uint32_t)p.ring_depth;
  const int total = p.n_frames * p.text_rows;

  const Layout L = make_layout(MODE, SP_BOX_STREAM, w, p.src_w, p.row_pitch, 0);
  GlyphLut *lut = reinterpret_cast<GlyphLut *>(smem + L.lut);
  uint32_t *cT = reinterpret_cast<uint32_t *>(smem + L.cT);
  uint32_t *cB = reinterpret_cast<uint32_t *>(smem + L.cB);
  uint16_t *key = reinterpret_cast<uint16_t *>(smem + L.key);
  uint16_t *hpos = reinterpret_cast<uint16_t *>(smem + L.hpos);
  uint16_t *rend = reinterpret_cast<uint16_t *>(smem + L.rend);
  uint32_t *off = reinterpret_cast<uint32_t *>(smem + L.off);
  uint16_t *V = reinterpret_cast<uint16_t *>(smem + L.V);
  uint8_t *outb = smem + L.outb;
  uint8_t *ring = smem + ((L.total + 127u) & ~127u);

  if (tid == 0) {
    for (uint32_t s = 0; s < D; s++) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], (NT / 32));
    }
    mbar_fence_init();
  }
  if (USES_LUT && tid < NT) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.lut);
    uint32_t *dst = reinterpret_cast<uint32_t *>(lut);
    for (int i = tid; i < (int)(sizeof(GlyphLut) / 4); i += NT) dst[i] = src[i];
  }
  if (tid < NT) init_dec3<NT>(tid);
  __syncthreads(); // the only CTA-wide barrier: after it the producer warp and the consumer warps part ways

  if (warp == (NT / 32)) { // ---------------- producer: one lane walks the same tile/row sequence as the consumers
    if (lane == 0) {
      uint32_t g = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int t = tile % p.text_rows, f = tile / p.text_rows;
        const WsBand b = ws_band<HB>(p, t);
        const uint8_t *src = p.frames + (size_t)f * p.frame_stride + (size_t)b.a0 * R;
        for (int r = b.a0; r < b.b1; r++, g++, src += R) {
          const uint32_t slot = g % D, ph = (g / D) & 1u;
          mbar_wait(&s_empty[slot], ph ^ 1u); // a fresh barrier passes the parity-1 wait: the ring starts empty
          mbar_expect_tx(&s_full[slot], R);
          bulk_g2s(ring + (size_t)slot * R, src, R, &s_full[slot]);
        }
      }
    }
    return;
  }

  // ---------------- consumers (threads 0..255)
  uint32_t g = 0;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    const int t = tile % p.text_rows, f = tile / p.text_rows;
    const WsBand b = ws_band<HB>(p, t);
#pragma unroll 1
    for (int half = 0; half < (HB ? 2 : 1); half++) {
      uint32_t *out = half ? cB : cT;
      if (half && !b.hasB) { // odd pixel height: bottom := top (halfblock.c:73,82-88)
        for (int x = tid; x < w; x += NT) cB[x] = cT[x];
        break;
      }
      const int r0 = half ? b.a1 : b.a0, r1 = half ? b.b1 : b.a1;
      uint32_t a[CPT][8];
#pragma unroll
      for (int j = 0; j < CPT; j++)
#pragma unroll
        for (int k = 0; k < 8; k++) a[j][k] = 0u;
      for (int r = r0; r < r1;) {
        const bool two = r + 1 < r1; // rows are summed in pairs (acc16x2); an odd tail pairs with zeros
        const uint32_t slot0 = g % D, ph0 = (g / D) & 1u;
        const uint32_t slot1 = (g + 1u) % D, ph1 = ((g + 1u) / D) & 1u;
        mbar_wait(&s_full[slot0], ph0);
        if (two) mbar_wait(&s_full[slot1], ph1);
        const uint4 *row0 = reinterpret_cast<const uint4 *>(ring + (size_t)slot0 * R);
        const uint4 *row1 = reinterpret_cast<const uint4 *>(ring + (size_t)slot1 * R);
#pragma unroll
        for (int j = 0; j < CPT; j++) {
          const int c = tid + j * NT;
          if (c < nchunk) {
            const uint4 u = row0[c];
            const uint4 v = two ? row1[c] : make_uint4(0u, 0u, 0u, 0u);
            acc16x2(a[j], u, v);
          }
        }
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[slot0]);
          if (two) mbar_arrive(&s_empty[slot1]);
        }
        r += two ? 2 : 1;
        g += two ? 2u : 1u;
      }
#pragma unroll
      for (int j = 0; j < CPT; j++) {
        const int c = tid + j * NT;
        if (c < nchunk) {
          uint4 lo, hi;
          lo.x = __byte_perm(a[j][0], a[j][1], 0x5410);
          lo.y = __byte_perm(a[j][0], a[j][1], 0x7632);
          lo.z = __byte_perm(a[j][2], a[j][3], 0x5410);
          lo.w = __byte_perm(a[j][2], a[j][3], 0x7632);
          hi.x = __byte_perm(a[j][4], a[j][5], 0x5410);
          hi.y = __byte_perm(a[j][4], a[j][5], 0x7632);
          hi.z = __byte_perm(a[j][6], a[j][7], 0x5410);
          hi.w = __byte_perm(a[j][6], a[j][7], 0x7632);
          uint4 *dst = reinterpret_cast<uint4 *>(V + (size_t)c * 16);
          dst[0] = lo;
          dst[1] = hi;
        }
      }
      SyncConsumers<NT>::sync();
      const uint32_t nrow = (uint32_t)(r1 - r0);
      for (int x = tid; x < w; x += NT) {
        int x0, x1;
        box_range(x, p.src_w, p.cols, x0, x1);
        uint32_t sr = 0, sg = 0, sb = 0;
        const uint16_t *q = V + 3 * x0;
#pragma unroll 4
        for (int xx = x0; xx < x1; xx++, q += 3) {
          sr += q[0];
          sg += q[1];
          sb += q[2];
        }
        const uint32_t n = (uint32_t)(x1 - x0) * nrow, h = n >> 1;
        out[x] = (((sr + h) / n) << 16) | (((sg + h) / n) << 8) | ((sb + h) / n);
      }
      SyncConsumers<NT>::sync(); // V is reused by the other pixel row / aliased by the row staging buffer
    }
    SyncConsumers<NT>::sync();
    emit_row<MODE, SyncConsumers<NT>, NT>(p, f, t, lut, cT, cB, key, hpos, rend, off, outb, s_tmp, s_cond, tid);
    SyncConsumers<NT>::sync(); // the staging buffer aliases V: finish copying out before the next tile's sums land
  }
}

// ------------------------------------------------------------------ role-split persistent row kernel (LDG streamers)
// Measured on B200: the streaming phase alone runs at ~98% of the copy-measured HBM peak, the one-tile-per-CTA fused
// kernel at ~79%, because every CTA stops issuing loads for the ~20% of its life it spends in the latency-bound
// emission phase.  Here the two phases run side by side inside one persistent CTA: eight streamer warps sum band
// after band into double-buffered cell rows and never wait for emission; one emitter warp turns tile k into bytes
// (the same emit_row code, instantiated for a single warp) while the streamers are already summing tile k+1.
//   named barriers: 1 streamers only (256) | 2,3 cells[b] full (256 arrive + 32 sync) | 4,5 cells[b] empty (32 arrive + 256 sync)
constexpr int WS2_ST = 256; // streamer threads
template <int ID, int N> __device__ __forceinline__ void nbar_sync() {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(N) : "memory");
}
template <int ID, int N> __device__ __forceinline__ void nbar_arrive() {
  asm volatile("bar.arrive %0, %1;" ::"n"(ID), "n"(N) : "memory");
}

struct Layout2 {
  uint32_t lut, c[3][2], key[2], hpos[2], rend[2], off[2], V, outb, total;
};
// Shared memory is kept as small as the mode allows: four CTAs must fit under the 196 KB carve-out step, or the
// L1 that backs the streamers' in-flight loads shrinks from 60 KB to 28 KB per SM (measured: -20% bandwidth).
__host__ __device__ inline Layout2 make_layout2(int mode, int direct, int cols, int src_w, uint32_t out_bytes) {
  Layout2 L;
  uint32_t o = 0;
  const bool uses_lut = mode <= EM_TRUE_FG;
  const bool uses_key = mode == EM_MONO_FG || mode == EM_HB_256 || mode == EM_HB_16;
  L.lut = o;
  o += uses_lut ? al16((uint32_t)sizeof(GlyphLut)) : 0u;
  for (int b = 0; b < 3; b++) // cell rows: one being summed, one prepared, one being written out
    for (int h = 0; h < 2; h++) {
      L.c[b][h] = o;
      o += (h == 0 || (mode >= EM_HB_TRUE && mode <= EM_HB_MONO)) ? al16(4u * cols) : 0u;
    }
  for (int b = 0; b < 2; b++) { // run/offset arrays: prepared tile + tile being written out
    L.key[b] = o;
    o += uses_key ? al16(2u * cols) : 0u;
    L.hpos[b] = o;
    o += al16(2u * cols);
    L.rend[b] = o;
    o += al16(2u * cols);
    L.off[b] = o;
    o += al16((direct ? 2u : 4u) * cols); // row offsets fit 16 bits when the row is staged in shared memory
  }
  L.V = o;
  o += al16(2u * 3u * src_w);
  L.outb = o; // not aliased with V: the emitter fills it while the streamers refill V
  o += al16(out_bytes) + 16u; // + room to stage the row with the destination's 16-byte phase
  L.total = o;
  return L;
}

// named barriers: 1 = streamers only; FULL(c) = 2+c, EMPTY(c) = 5+c for cell buffer c in {0,1,2}
template <int N> __device__ __forceinline__ void nbar_sync_id(int id) {
  switch (id) {
  case 2: nbar_sync<2, N>(); break;
  case 3: nbar_sync<3, N>(); break;
  case 4: nbar_sync<4, N>(); break;
  case 5: nbar_sync<5, N>(); break;
  case 6: nbar_sync<6, N>(); break;
  default: nbar_sync<7, N>(); break;
  }
}
template <int N> __device__ __forceinline__ void nbar_arrive_id(int id) {
  switch (id) {
  case 2: nbar_arrive<2, N>(); break;
  case 3: nbar_arrive<3, N>(); break;
  case 4: nbar_arrive<4, N>(); break;
  case 5: nbar_arrive<5, N>(); break;
  case 6: nbar_arrive<6, N>(); break;
  default: nbar_arrive<7, N>(); break;
  }
}

// four CTAs per SM (56 registers): the occupancy every measurement in DESIGN.md §7 was taken at.  FILT (a colour
// filter fused into the band sums) keeps 12 x 16 bytes of loads plus 48 column sums in registers: two CTAs per SM.
template <int MODE, bool FILT>
__global__ void __launch_bounds__(WS2_ST + 32, FILT ? 2 : 4) k_render_rows_ws2(const RenderParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_tmp[2];
  __shared__ uint32_t s_cond[2][4];
  __shared__ int s_tile[3];
  __shared__ uint32_t s_lb[2];
  constexpr bool HB = MODE >= EM_HB_TRUE && MODE <= EM_HB_MONO;
  constexpr bool USES_LUT = MODE <= EM_TRUE_FG;
  constexpr int NB = WS2_ST + 32;
  const int tid = threadIdx.x;
  const int w = p.cols;
  const int total = p.n_frames * p.text_rows;
  const Layout2 L = make_layout2(MODE, p.direct, w, p.src_w, p.row_pitch);
  GlyphLut *lut = reinterpret_cast<GlyphLut *>(smem + L.lut);
  uint16_t *V = reinterpret_cast<uint16_t *>(smem + L.V);
  // tiles are handed out by an atomic ticket (the word after the last look-back record): whoever holds tile X knows
  // every tile < X is already held by a running CTA, which makes the emitter's look-back spin deadlock-free even if
  // the grid is not fully resident
  int *ticket = p.ticket;

  if (USES_LUT) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(p.lut);
    uint32_t *dst = reinterpret_cast<uint32_t *>(lut);
    for (int i = tid; i < (int)(sizeof(GlyphLut) / 4); i += NB) dst[i] = src[i];
  }
  init_dec3<NB>(tid);
  if (FILT) init_filt<NB>(p, tid);
  __syncthreads();

  if (tid < WS2_ST) { // ---------------- streamers
    // The ticket for tile k+1 is requested while tile k is being summed: all CTAs ask in bursts, and a same-address
    // atomic that is waited for on the spot costs a queueing delay per tile (measured: +20% kernel time).
    int pending = 0;
    if (tid == 0) pending = (int)((uint32_t)atomicAdd(ticket, 1) - p.ticket_base);
    for (int k = 0;; k++) {
      const int c = k % 3;
      if (k >= 3) { // the emitter has written out tile k-3: cells[c] / s_tile[c] are free
        const long long t0 = (p.tune_flags & 4) ? clock64() : 0;
        nbar_sync_id<NB>(5 + c);
        if ((p.tune_flags & 4) && tid == 0) atomicAdd(p.dbg + 0, (unsigned long long)(clock64() - t0));
      }
      if (tid == 0) {
        const int tk = pending;
        s_tile[c] = tk < total ? tk : -1;
        if (tk < total) pending = (int)((uint32_t)atomicAdd(ticket, 1) - p.ticket_base);
      }
      SyncConsumers<WS2_ST>::sync();
      const int tile = s_tile[c];
      if (tile < 0) { // no more work: pass the sentinel on, absorb the emitter's last two "empty" arrivals, leave
        nbar_arrive_id<NB>(2 + c);
        if (k >= 2) nbar_sync_id<NB>(5 + (k - 2) % 3);
        if (k >= 1) nbar_sync_id<NB>(5 + (k - 1) % 3);
        break;
      }
      const int t = tile % p.text_rows, f = tile / p.text_rows;
      uint32_t *cT = reinterpret_cast<uint32_t *>(smem + L.c[c][0]);
      uint32_t *cB = reinterpret_cast<uint32_t *>(smem + L.c[c][1]);
      const uint8_t *frame = p.frames + (size_t)f * p.frame_stride;
      const int yT = HB ? 2 * t : t;
      const bool hasB = HB && (2 * t + 1 < p.rows_px);
#pragma unroll 1
      for (int hrow = 0; hrow < (hasB ? 2 : 1); hrow++) // one inlined copy of the band code for both pixel rows
        cells_box_stream<WS2_ST, SyncConsumers<WS2_ST>, FILT>(p, frame, yT + hrow, hrow ? cB : cT, V);
      if (HB && !hasB)
        for (int x = tid; x < w; x += WS2_ST) cB[x] = cT[x];
      __threadfence_block();
      nbar_arrive_id<NB>(2 + c); // cells[c] are complete: hand them to the emitter and move on
    }
  } else { // ---------------- emitter warp, software-pipelined by one tile:
    //   iteration k: prepare + publish tile k (never waits), then look-back + write-out of tile k-1, whose
    //   predecessors have had a whole tile period to publish — so a late row cannot start a convoy
    const int lane = tid - WS2_ST;
    uint8_t *outb = smem + L.outb;
    uint32_t prev_bytes = 0;
    int prev_tile = -1;
    long long t_prev = 0;
    for (int k = 0;; k++) {
      const int c = k % 3, a = k & 1;
      const long long t0 = (p.tune_flags & 4) ? clock64() : 0;
      nbar_sync_id<NB>(2 + c);
      const long long t1 = (p.tune_flags & 4) ? clock64() : 0;
      if ((p.tune_flags & 4) && lane == 0) {
        atomicAdd(p.dbg + 1, (unsigned long long)(t1 - t0));          // emitter waiting for cells
        if (k > 0) atomicAdd(p.dbg + 2, (unsigned long long)(t0 - t_prev)); // emitter busy with the previous iteration
        atomicAdd(p.dbg + 3, 1ull);
      }
      t_prev = t1;
      const int tile = s_tile[c];
      uint32_t bytes = 0;
      if (tile >= 0 && (p.tune_flags & 2)) { // measurement knob ACB200_WS2_NOEMIT: streamers alone (output is garbage)
        __threadfence_block();
        nbar_arrive_id<NB>(5 + c);
        continue;
      }
      if (tile >= 0 && !p.direct) { // measurement knob: scratch rows + k_stitch instead of look-back placement
        const int t = tile % p.text_rows, f = tile / p.text_rows;
        emit_row<MODE, SyncWarp, 32>(p, f, t, lut, reinterpret_cast<uint32_t *>(smem + L.c[c][0]),
                                     reinterpret_cast<uint32_t *>(smem + L.c[c][1]),
                                     reinterpret_cast<uint16_t *>(smem + L.key[a]),
                                     reinterpret_cast<uint16_t *>(smem + L.hpos[a]),
                                     reinterpret_cast<uint16_t *>(smem + L.rend[a]),
                                     reinterpret_cast<uint32_t *>(smem + L.off[a]), outb, s_tmp, s_cond[a], lane);
        __threadfence_block();
        nbar_arrive_id<NB>(5 + c);
        continue;
      }
      if (tile >= 0) {
        const int t = tile % p.text_rows, f = tile / p.text_rows;
        bytes = emit_direct_prepare<MODE, SyncWarp, 32>(p, f, t, lut, reinterpret_cast<uint32_t *>(smem + L.c[c][0]),
                                          reinterpret_cast<uint32_t *>(smem + L.c[c][1]),
                                          reinterpret_cast<uint16_t *>(smem + L.key[a]),
                                          reinterpret_cast<uint16_t *>(smem + L.hpos[a]),
                                          reinterpret_cast<uint16_t *>(smem + L.rend[a]),
                                          reinterpret_cast<uint16_t *>(smem + L.off[a]), s_tmp, s_cond[a], lane);
      }
      if (prev_tile >= 0) {
        const int pc = (k - 1) % 3, pa = (k - 1) & 1;
        const int t = prev_tile % p.text_rows, f = prev_tile / p.text_rows;
        emit_direct_finish<MODE, SyncWarp, 32>(p, f, t, lut, reinterpret_cast<uint32_t *>(smem + L.c[pc][0]),
                                 reinterpret_cast<uint32_t *>(smem + L.c[pc][1]),
                                 reinterpret_cast<uint16_t *>(smem + L.key[pa]),
                                 reinterpret_cast<uint16_t *>(smem + L.hpos[pa]),
                                 reinterpret_cast<uint16_t *>(smem + L.rend[pa]),
                                 reinterpret_cast<const uint16_t *>(smem + L.off[pa]), outb, s_cond[pa], s_lb, prev_bytes, lane);
        __threadfence_block();
        nbar_arrive_id<NB>(5 + pc);
      }
      if (tile < 0) break;
      prev_tile = tile;
      prev_bytes = bytes;
    }
  }
}

// kernel attributes are per device: one bit per CUDA ordinal, set once the opt-in has been made on that device
static inline bool attr_done(std::atomic<uint64_t> &mask, int dev) { return (mask.load(std::memory_order_acquire) >> dev) & 1ull; }
static inline void attr_set(std::atomic<uint64_t> &mask, int dev) { mask.fetch_or(1ull << dev, std::memory_order_release); }
static inline int current_sms(int *dev_out) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
  if (dev_out) *dev_out = dev;
  return sms;
}

template <int MODE, bool FILT>
static cudaError_t launch_ws2_mode_f(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  const Layout2 L = make_layout2(MODE, p.direct, p.cols, p.src_w, p.row_pitch);
  if (L.total > kMaxDynSmem) return cudaErrorInvalidConfiguration;
  // nothing cached across calls except the per-device opt-in: callers with different geometries run concurrently
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  const int sms = current_sms(&dev);
  if (!attr_done(configured, dev)) {
    cudaError_t e = allow_max_dyn_smem(k_render_rows_ws2<MODE, FILT>);
    if (e != cudaSuccess) return e;
    attr_set(configured, dev);
  }
  int ctas_per_sm = 0;
  cudaError_t e =
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_render_rows_ws2<MODE, FILT>, WS2_ST + 32, L.total);
  if (e != cudaSuccess || ctas_per_sm < 1) return e != cudaSuccess ? e : cudaErrorInvalidConfiguration;
  const long long total = (long long)p.n_frames * p.text_rows;
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > total) grid = total;
  if (grid_out) *grid_out = (unsigned)grid;
  k_render_rows_ws2<MODE, FILT><<<(unsigned)grid, WS2_ST + 32, L.total, st>>>(p);
  return cudaGetLastError();
}
template <int MODE> static cudaError_t launch_ws2_mode(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  return p.filt_mode != FM_NONE ? launch_ws2_mode_f<MODE, true>(p, st, grid_out)
                                : launch_ws2_mode_f<MODE, false>(p, st, grid_out);
}

// opt in to the largest dynamic shared memory the kernel can have: 227 KB per CTA minus its static allocation
template <class K> static cudaError_t allow_max_dyn_smem(K kernel) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes);
}

// ------------------------------------------------------------------ per-mode launch templates
// NT (threads that own cells / byte columns) is chosen so that one pass of the CTA covers the text row:
// 128, 256, 384 or 512.
// (measured on B200: 256 threads beat 384/512 for the one-tile-per-CTA kernels even when the row then needs two
// passes, because five CTAs per SM keep more loads in flight than three)
__host__ inline int pick_nt(int cols) { return cols <= 128 ? 128 : 256; }

template <int MODE, int SP, int NT>
static cudaError_t launch_rows_t(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  const uint32_t cap = p.use_smem_out ? p.row_pitch : 0u;
  const Layout L = make_layout(MODE, SP, p.cols, p.src_w, cap, p.tune_flags & 1, p.direct);
  if (L.total > kMaxDynSmem) return cudaErrorInvalidConfiguration;
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  const int sms = current_sms(&dev);
  if (!attr_done(configured, dev)) {
    // opt-in limit is 227 KB per CTA for static + dynamic together; the kernel has < 1 KB static
    cudaError_t e = allow_max_dyn_smem(k_render_rows<MODE, SP, NT>);
    if (e != cudaSuccess) return e;
    attr_set(configured, dev);
  }
  unsigned grid = (unsigned)p.n_frames * (unsigned)p.text_rows;
  if (p.direct) { // persistent: as many CTAs as are resident at once, tiles drawn from the ticket
    int ctas_per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_render_rows<MODE, SP, NT>, NT, L.total);
    if (e != cudaSuccess || ctas_per_sm < 1) return e != cudaSuccess ? e : cudaErrorInvalidConfiguration;
    const unsigned resident = (unsigned)sms * (unsigned)ctas_per_sm;
    if (grid > resident) grid = resident;
  }
  if (grid_out) *grid_out = grid;
  k_render_rows<MODE, SP, NT><<<grid, NT, L.total, st>>>(p);
  return cudaGetLastError();
}
template <int MODE, int SP> static cudaError_t launch_rows_nt(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  // (384 threads for 257..384 columns — one pass over the row instead of two — measured slower: five CTAs per SM keep
  // fewer tiles in flight than eight, 0.248 vs 0.228 ms on flat C3 frames, profiles/r02j_configs.txt)
  // Nearest neighbour is instruction-bound (profiles/r02i_ncu_nn_flat: 76 % issue utilisation), so the CTA width is
  // chosen to leave no lane idle: a 320-column row on 256 threads runs every per-cell loop and scan pass twice, the
  // second time with a quarter of the lanes (16 warp-passes); 160 threads do it in two FULL passes (10 warp-passes) and
  // eight such CTAs are resident per SM, as many tiles in flight as before.  Measured on 256 x C3 frames
  // (profiles/r02w_nn_nt.txt): 0.319 / 0.228 ms (noise / flat) with 256 threads, 0.317 / 0.226 with 320 (one pass, but
  // six CTAs per SM), **0.282 / 0.191** with 160.  Rows of 129..160 columns take 160 threads for the same reason (one
  // full pass instead of 256 threads with three idle warps).  ACB200_NN_NT=256|320 forces the other widths (A/B).
  static const int nn_nt = getenv("ACB200_NN_NT") ? atoi(getenv("ACB200_NN_NT")) : 0;
  if (SP == SP_NN && nn_nt != 256) {
    if (nn_nt == 320 && p.cols > 256 && p.cols <= 320) return launch_rows_t<MODE, SP, SP == SP_NN ? 320 : 256>(p, st, grid_out);
    if ((p.cols > 256 && p.cols <= 320) || (p.cols > 128 && p.cols <= 160))
      return launch_rows_t<MODE, SP, SP == SP_NN ? 160 : 256>(p, st, grid_out);
  }
  switch (pick_nt(p.cols)) {
  case 128: return launch_rows_t<MODE, SP, 128>(p, st, grid_out);
  default: return launch_rows_t<MODE, SP, 256>(p, st, grid_out);
  }
}
template <int MODE>
static cudaError_t launch_rows_mode(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out) {
  switch (sp) {
  case SP_NN: return launch_rows_nt<MODE, SP_NN>(p, st, grid_out);
  case SP_BOX_GENERIC: return launch_rows_nt<MODE, SP_BOX_GENERIC>(p, st, grid_out);
  default: return launch_rows_nt<MODE, SP_BOX_STREAM>(p, st, grid_out);
  }
}

template <int MODE, int CPT, int NT> static cudaError_t launch_ws_t(const RenderParams &p, cudaStream_t st) {
  const Layout L = make_layout(MODE, SP_BOX_STREAM, p.cols, p.src_w, p.row_pitch, 0);
  const size_t smem = ((L.total + 127u) & ~127u) + (size_t)p.ring_depth * p.src_w * 3u;
  if (smem > kMaxDynSmem) return cudaErrorInvalidConfiguration;
  static std::atomic<uint64_t> configured{0};
  int dev = 0;
  const int sms = current_sms(&dev);
  if (!attr_done(configured, dev)) {
    cudaError_t e = allow_max_dyn_smem(k_render_rows_ws<MODE, CPT, NT>);
    if (e != cudaSuccess) return e;
    attr_set(configured, dev);
  }
  int ctas_per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, k_render_rows_ws<MODE, CPT, NT>, NT + 32, smem);
  if (e != cudaSuccess || ctas_per_sm < 1) return e != cudaSuccess ? e : cudaErrorInvalidConfiguration;
  const long long total = (long long)p.n_frames * p.text_rows;
  long long grid = (long long)sms * ctas_per_sm;
  if (grid > total) grid = total;
  k_render_rows_ws<MODE, CPT, NT><<<(unsigned)grid, NT + 32, smem, st>>>(p);
  return cudaGetLastError();
}
template <int MODE, int NT> static cudaError_t launch_ws_cpt(const RenderParams &p, cudaStream_t st) {
  const int nchunk = (p.src_w * 3) >> 4;
  const int cpt = (nchunk + NT - 1) / NT;
  if (cpt <= 1) return launch_ws_t<MODE, 1, NT>(p, st);
  if (cpt <= 2) return launch_ws_t<MODE, 2, NT>(p, st);
  if (cpt <= 3) return launch_ws_t<MODE, 3, NT>(p, st);
  if (cpt <= 4) return launch_ws_t<MODE, 4, NT>(p, st);
  return cudaErrorInvalidConfiguration;
}
// the warp-specialised kernel always runs with >= 256 consumer threads (the sums need the issue slots)
__host__ inline int pick_ws_nt(int cols, int src_w) {
  int nt = cols <= 256 ? 256 : cols <= 384 ? 384 : 512;
  static const int nt_env = getenv("ACB200_WS_NT") ? atoi(getenv("ACB200_WS_NT")) : 0; // tuning knob
  if (nt_env == 256 || nt_env == 384 || nt_env == 512) nt = nt_env;
  while (nt < 512 && (((src_w * 3) >> 4) + nt - 1) / nt > 4) nt += 128;
  return nt;
}
template <int MODE> static cudaError_t launch_ws_mode(const RenderParams &p, cudaStream_t st) {
  switch (pick_ws_nt(p.cols, p.src_w)) {
  case 256: return launch_ws_cpt<MODE, 256>(p, st);
  case 384: return launch_ws_cpt<MODE, 384>(p, st);
  default: return launch_ws_cpt<MODE, 512>(p, st);
  }
}

} // namespace acb
