// effects.cu — the steps either side of the convert:
//
//   * client display path, src/common/session/display.c:484-671: flip / colour filter / rainbow replace are fused
//     into the render kernels (render_dev.cuh: cells_nn, filter_px, fg_print); this file holds the host entry
//     acb200_display_convert, the host float rainbow hue, and the standalone whole-image filter kernel
//     k_color_filter behind apply_color_filter / acb200_color_filter_device (lib/video/rgba/color_filter.c:274-346).
//   * wire packaging, lib/network/acip/server.c:188-236 + lib/network/crc32.c: CRC32-C of every finished frame while
//     it is still in HBM (k_crc32c_chunks + k_crc32c_finish) and the 24-byte big-endian ascii_frame_packet_t;
//     k_trailing_reset_fixup is the device form of the server's "frame must end in ESC[0m" cut (stream.c:1085-1127).
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "engine.h"

using namespace acb;

// ====================================================================== whole-image colour filter
// In-place map over packed RGB24: out = T[gray(r,g,b)], T = 256-entry table of the filter (the filtered pixel depends
// on the pixel only through its BT.601 grey value, color_filter.c:238-267).  HBM-bound: 3 B read + 3 B written per
// pixel.  Each thread owns 16 pixels = three 16-byte words; grey by DP4A, table in shared memory.
namespace {

__device__ __forceinline__ uint32_t filt_entry(uint32_t gray, int mode, uint32_t frgb) { // -> 0x00BBGGRR (memory order)
  if (mode == FM_RAINBOW) gray = 179u + (gray * 76u) / 255u; // color_filter.c:315
  const uint32_t fr = (frgb >> 16) & 255u, fg = (frgb >> 8) & 255u, fb = frgb & 255u;
  uint32_t r, g, b;
  if (mode == FM_ON_WHITE) { // :254-260
    const uint32_t w = 255u * gray, ig = 255u - gray;
    r = (fr * ig + w) / 255u, g = (fg * ig + w) / 255u, b = (fb * ig + w) / 255u;
  } else { // :262-265
    r = (fr * gray) / 255u, g = (fg * gray) / 255u, b = (fb * gray) / 255u;
  }
  return r | (g << 8) | (b << 16);
}

__device__ __forceinline__ uint32_t gray_of_bytes(uint32_t rgbx) { // bytes r,g,b,(ignored): (77r+150g+29b)>>8
  return __dp4a(rgbx, 0x001D964Du, 0u) >> 8;                       // color_filter.h:172 (no rounding term)
}

constexpr int CF_NT = 256;

// touched exactly once: keep it out of L1 on the way in, mark it evict-first on the way out
__device__ __forceinline__ uint4 ld_once(const uint4 *p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_once(uint4 *p, const uint4 &v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(CF_NT) k_color_filter(uint8_t *pixels, size_t n_bytes, int mode, uint32_t frgb) {
  __shared__ uint32_t T[256];
  for (int i = threadIdx.x; i < 256; i += CF_NT) T[i] = filt_entry((uint32_t)i, mode, frgb);
  __syncthreads();
  const size_t ngroups = n_bytes / 48; // 16 pixels per group
  uint4 *base = reinterpret_cast<uint4 *>(pixels);
  const size_t stride = (size_t)gridDim.x * CF_NT;
  // two groups per trip: six 16-byte loads in flight per thread before the first table lookup
  for (size_t gidx = (size_t)blockIdx.x * CF_NT + threadIdx.x; gidx < ngroups; gidx += 2 * stride) {
    const bool two = gidx + stride < ngroups;
    uint4 *q0 = base + gidx * 3, *q1 = base + (two ? gidx + stride : gidx) * 3;
    uint4 in[6];
#pragma unroll
    for (int k = 0; k < 3; k++) in[k] = ld_once(q0 + k);
#pragma unroll
    for (int k = 0; k < 3; k++) in[3 + k] = two ? ld_once(q1 + k) : in[k];
#pragma unroll
    for (int g = 0; g < 2; g++) {
      if (g == 1 && !two) break;
      const uint4 a = in[3 * g], b = in[3 * g + 1], c = in[3 * g + 2];
      const uint32_t w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
      uint32_t o[12];
#pragma unroll
      for (int k = 0; k < 4; k++) { // 4 pixels live in 3 consecutive words
        const uint32_t w0 = w[3 * k], w1 = w[3 * k + 1], w2 = w[3 * k + 2];
        const uint32_t p0 = T[gray_of_bytes(w0)];                           // bytes 0..2
        const uint32_t p1 = T[gray_of_bytes(__byte_perm(w0, w1, 0x6543))];  // bytes 3..5
        const uint32_t p2 = T[gray_of_bytes(__byte_perm(w1, w2, 0x5432))];  // bytes 6..8
        const uint32_t p3 = T[gray_of_bytes(w2 >> 8)];                      // bytes 9..11
        o[3 * k] = p0 | (p1 << 24);
        o[3 * k + 1] = (p1 >> 8) | (p2 << 16);
        o[3 * k + 2] = (p2 >> 16) | (p3 << 8);
      }
      uint4 *q = g ? q1 : q0;
      st_once(q, make_uint4(o[0], o[1], o[2], o[3]));
      st_once(q + 1, make_uint4(o[4], o[5], o[6], o[7]));
      st_once(q + 2, make_uint4(o[8], o[9], o[10], o[11]));
    }
  }
  // tail (< 16 pixels) and nothing else: byte-wise by the first threads of block 0
  if (blockIdx.x == 0) {
    const size_t done = ngroups * 48;
    for (size_t i = done + 3u * threadIdx.x; i + 2 < n_bytes; i += 3u * CF_NT) {
      const uint32_t t = T[(77u * pixels[i] + 150u * pixels[i + 1] + 29u * pixels[i + 2]) >> 8];
      pixels[i] = (uint8_t)t;
      pixels[i + 1] = (uint8_t)(t >> 8);
      pixels[i + 2] = (uint8_t)(t >> 16);
    }
  }
}

// strided rows (stride != 3*width) or a base that is not 16-byte aligned: one thread per pixel
__global__ void __launch_bounds__(CF_NT) k_color_filter_rows(uint8_t *pixels, uint32_t width, uint32_t height,
                                                             uint32_t stride, int mode, uint32_t frgb) {
  __shared__ uint32_t T[256];
  for (int i = threadIdx.x; i < 256; i += CF_NT) T[i] = filt_entry((uint32_t)i, mode, frgb);
  __syncthreads();
  const size_t total = (size_t)width * height;
  for (size_t i = (size_t)blockIdx.x * CF_NT + threadIdx.x; i < total; i += (size_t)gridDim.x * CF_NT) {
    uint8_t *p = pixels + (i / width) * stride + (i % width) * 3u;
    const uint32_t t = T[(77u * p[0] + 150u * p[1] + 29u * p[2]) >> 8];
    p[0] = (uint8_t)t;
    p[1] = (uint8_t)(t >> 8);
    p[2] = (uint8_t)(t >> 16);
  }
}

int sm_count() { return device_sms(); }

int filter_device(uint8_t *d_pixels, uint32_t width, uint32_t height, uint32_t stride, int mode, uint32_t frgb,
                  cudaStream_t st) {
  const size_t total_px = (size_t)width * height;
  if (stride == width * 3u && (reinterpret_cast<uintptr_t>(d_pixels) & 15u) == 0) {
    const size_t groups = total_px * 3 / 48;
    size_t want = (groups + 2 * CF_NT - 1) / (2 * CF_NT); // two groups per thread per trip
    static const int cf_ctas = getenv("ACB200_CF_CTAS_PER_SM") ? atoi(getenv("ACB200_CF_CTAS_PER_SM")) : 0; // tuning knob
    const size_t cap = cf_ctas > 0 ? (size_t)sm_count() * cf_ctas : want; // default: one trip per thread (no grid-stride)
    unsigned grid = (unsigned)(want < 1 ? 1 : want > cap ? cap : want);
    k_color_filter<<<grid, CF_NT, 0, st>>>(d_pixels, total_px * 3, mode, frgb);
  } else {
    size_t want = (total_px + CF_NT - 1) / CF_NT;
    const size_t cap = (size_t)sm_count() * 8;
    unsigned grid = (unsigned)(want < 1 ? 1 : want > cap ? cap : want);
    k_color_filter_rows<<<grid, CF_NT, 0, st>>>(d_pixels, width, height, stride, mode, frgb);
  }
  ACB_CUDA(cudaGetLastError());
  count_launch();
  return E_OK;
}

} // namespace

// ====================================================================== CRC32-C of finished frames
// Standard CRC-32C (init ~0, reflected 0x82F63B78, final ~): crc(A||B) = crc(A) * x^(8|B|) mod P  xor  crc(B), so a
// frame is cut into 256-byte segments (one per thread, a lane-private byte table in shared memory), every segment CRC is
// multiplied by x^(8 * bytes-after-it) and the products are XORed — first inside a 64 KB chunk (k_crc32c_chunks),
// then over the chunks of a frame (k_crc32c_finish, which also writes the packet header).
namespace {

constexpr uint32_t CRC_POLY = 0x82F63B78u;
constexpr int CRC_NT = 256, CRC_SEG = 256, CRC_CHUNK = CRC_NT * CRC_SEG; // 64 KB per CTA, 256 B per thread

constexpr int CRC_ROW = 512;          // row kernel: bytes a warp consumes per step (32 lanes x 16 bytes)
constexpr int CRC_ROWPOW = 4096;      // rows covered by the constant shift table (2 MB frames; longer ones square-and-multiply)
constexpr int CRC_NSETS = 7;          // slice tables: multiply by x^32, x^128, x^256, x^512, x^1024, x^2048, x^4096
constexpr int CRC_ROWS_NT = 1024;     // row kernel: threads per CTA (one CTA per SM: 152 KB of tables + 64 KB of rows in flight)
constexpr size_t CRC_ROWS_SMEM = (size_t)(4 * 256 * 32 + 6 * 4 * 256) * sizeof(uint32_t) + (size_t)32 * 4 * 512; // tables + row ring

struct CrcTables {
  uint32_t x2n[32];   // x^(2^k) mod P, reflected
  uint32_t seg[256];  // x^(8 * CRC_SEG * k) mod P: shift past k whole segments
  uint32_t rowpow[CRC_ROWPOW]; // x^(8 * CRC_ROW * k) mod P: shift past k whole rows
  uint32_t bytepow[CRC_ROW];   // x^(8 * k) mod P
};
__constant__ CrcTables c_crc;
// z[s][j][b] = (b << 8j) * x^k mod P for k = 32, 128, 256, 512, 1024, 2048, 4096: multiplying a 32-bit state by x^k is
// four lookups.  Global memory, one copy per device; the row kernel's CTAs copy them into shared memory.
uint32_t *g_crc_slices[kMaxDevices] = {};

__host__ __device__ inline uint32_t gf_mul(uint32_t a, uint32_t b) { // a * b mod P (reflected representation)
  uint32_t p = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    p ^= (a & (0x80000000u >> i)) ? b : 0u;
    b = (b >> 1) ^ ((b & 1u) ? CRC_POLY : 0u);
  }
  return p;
}
__host__ __device__ inline uint32_t gf_xpow8(const uint32_t *x2n, uint64_t n_bytes) { // x^(8 n) mod P
  uint32_t p = 0x80000000u;                                                         // the polynomial "1"
  uint64_t n = n_bytes;
  for (int k = 3; n; n >>= 1, k = (k + 1) & 31)
    if (n & 1u) p = gf_mul(x2n[k], p);
  return p;
}

cudaError_t crc_rows_opt_in(); // the row kernel's shared-memory opt-in (defined behind the kernel)
// 0 = by arena size (launch_frame_packets), 1 = row form, 2 = segment form
std::atomic<int> g_crc_form{[] {
  const char *e = getenv("ACB200_CRC_KERNEL");
  return !e ? 0 : !strcmp(e, "rows") ? 1 : !strcmp(e, "segments") ? 2 : 0;
}()};
// __constant__ memory is per device: the tables are uploaded once on every device the library runs on
std::mutex g_crc_mu;
uint64_t g_crc_done = 0; // bit per CUDA ordinal
int crc_tables_init() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return E_INVALID_STATE;
  std::lock_guard<std::mutex> lk(g_crc_mu);
  if ((g_crc_done >> dev) & 1ull) return E_OK;
  CrcTables t;
  uint32_t p = 0x40000000u; // x^1
  t.x2n[0] = p;
  for (int k = 1; k < 32; k++) t.x2n[k] = p = gf_mul(p, p);
  for (int k = 0; k < 256; k++) t.seg[k] = gf_xpow8(t.x2n, (uint64_t)CRC_SEG * k);
  const uint32_t xrow = gf_xpow8(t.x2n, CRC_ROW);
  t.rowpow[0] = t.bytepow[0] = 0x80000000u;
  for (int k = 1; k < CRC_ROWPOW; k++) t.rowpow[k] = gf_mul(t.rowpow[k - 1], xrow);
  for (int k = 1; k < CRC_ROW; k++) t.bytepow[k] = gf_mul(t.bytepow[k - 1], t.x2n[3]);
  if (cudaMemcpyToSymbol(c_crc, &t, sizeof(t)) != cudaSuccess) return E_INVALID_STATE;
  std::vector<uint32_t> z((size_t)CRC_NSETS * 4 * 256);
  static const int kBits[CRC_NSETS] = {32, 128, 256, 512, 1024, 2048, 4096};
  for (int si = 0; si < CRC_NSETS; si++) {
    const uint32_t K = gf_xpow8(t.x2n, (uint64_t)kBits[si] / 8);
    for (int j = 0; j < 4; j++)
      for (int b = 0; b < 256; b++) z[((size_t)si * 4 + j) * 256 + b] = gf_mul((uint32_t)b << (8 * j), K);
  }
  if (cudaMalloc((void **)&g_crc_slices[dev], z.size() * sizeof(uint32_t)) != cudaSuccess ||
      cudaMemcpy(g_crc_slices[dev], z.data(), z.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess ||
      crc_rows_opt_in() != cudaSuccess) {
    cudaGetLastError();
    return E_INVALID_STATE;
  }
  g_crc_done |= 1ull << dev;
  return E_OK;
}

__device__ __forceinline__ uint4 ld_stream16(const uint4 *p) { // read once; volatile asm = issued where written
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// The byte table, replicated once per lane: entry i of lane l lives at word i*32 + l, i.e. in bank l.  A warp's 32
// lookups (32 unrelated indices) then hit 32 different banks — one shared-memory wavefront per lookup instead of the
// ~3.5 a shared 1 KB table costs.  (Measured: this removes the conflicts, l1tex bank-conflict counter 2.7 M -> 6 k per
// 40 MB, but not the bound — the dependent lookup chain itself, 0.123 ms per 322 MB with the global loads switched off.)
__device__ __forceinline__ uint32_t crc_word(const uint32_t *Tl, uint32_t crc, uint32_t w) { // Tl = table + lane
  crc ^= w;
#pragma unroll
  for (int k = 0; k < 4; k++) crc = Tl[(crc & 255u) << 5] ^ (crc >> 8);
  return crc;
}

// grid (chunks, frames).  part[f * max_chunks + c] = standard CRC of chunk c of frame f (0 for chunks past the end).
// copy_dst != nullptr: the bytes are also streamed to copy_dst + f*copy_pitch (mapped host memory on the one-frame
// path), so the frame leaves HBM once for both purposes.
__global__ void __launch_bounds__(CRC_NT) k_crc32c_chunks(const uint8_t *out, size_t out_pitch, const uint32_t *out_len,
                                                          uint32_t *part, int max_chunks, uint8_t *copy_dst,
                                                          size_t copy_pitch, int noload) {
  __shared__ uint32_t T[256 * 32]; // 32 KB
  __shared__ uint32_t s_t0[256];
  __shared__ uint32_t s_red[CRC_NT / 32];
  const int tid = threadIdx.x, f = blockIdx.y, c = blockIdx.x;
  const uint32_t L = out_len[f];
  const size_t c0 = (size_t)c * CRC_CHUNK;
  if (c0 >= L) {
    if (tid == 0) part[(size_t)f * max_chunks + c] = 0u;
    return;
  }
  {
    uint32_t v = (uint32_t)tid;
#pragma unroll
    for (int k = 0; k < 8; k++) v = (v >> 1) ^ ((v & 1u) ? CRC_POLY : 0u);
    s_t0[tid] = v;
  }
  __syncthreads();
#pragma unroll 4
  for (int k = 0; k < 32; k++) { // word w = k*256 + tid holds entry w >> 5 (a warp writes 32 consecutive words)
    const int w = k * CRC_NT + tid;
    T[w] = s_t0[w >> 5];
  }
  __syncthreads();
  const uint32_t *Tl = T + (tid & 31);

  const uint32_t n = (uint32_t)((L - c0) < (size_t)CRC_CHUNK ? (L - c0) : (size_t)CRC_CHUNK); // bytes in this chunk
  const uint32_t tl = (n - 1u) / CRC_SEG, rem = n - tl * CRC_SEG;                            // last segment, 1..SEG bytes
  const uint8_t *src = out + (size_t)f * out_pitch + c0 + (size_t)tid * CRC_SEG;
  uint8_t *dst = copy_dst ? copy_dst + (size_t)f * copy_pitch + c0 + (size_t)tid * CRC_SEG : nullptr;
  uint32_t contrib = 0u;
  if ((uint32_t)tid <= tl) {
    const uint32_t mine = (uint32_t)tid < tl ? (uint32_t)CRC_SEG : rem; // bytes of this thread's segment
    const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    uint32_t crc = 0xFFFFFFFFu, done = 0;
    // 64-byte batches, software-pipelined: the four loads of batch i+1 are issued (asm volatile keeps them where they
    // are written) before the 64 dependent table steps of batch i, so DRAM latency hides behind the recurrence
    uint4 cur[4], nxt[4];
    if (mine >= 64u) {
#pragma unroll
      for (int k = 0; k < 4; k++) cur[k] = noload ? make_uint4(k, tid, 0u, 0u) : ld_stream16(s4 + k);
    }
    for (; done + 64u <= mine; done += 64u) {
      const bool more = done + 128u <= mine;
#pragma unroll
      for (int k = 0; k < 4; k++)
        nxt[k] = (more && !noload) ? ld_stream16(s4 + (done >> 4) + 4 + k) : make_uint4(done, k, tid, 0u);
      if (dst) {
#pragma unroll
        for (int k = 0; k < 4; k++) d4[(done >> 4) + k] = cur[k];
      }
#pragma unroll
      for (int k = 0; k < 4; k++) {
        crc = crc_word(Tl, crc, cur[k].x);
        crc = crc_word(Tl, crc, cur[k].y);
        crc = crc_word(Tl, crc, cur[k].z);
        crc = crc_word(Tl, crc, cur[k].w);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) cur[k] = nxt[k];
    }
    for (; done < mine; done++) { // the frame's ragged end: < 64 bytes, once per chunk at most
      const uint8_t b = src[done];
      if (dst) dst[done] = b;
      crc = Tl[((crc ^ b) & 255u) << 5] ^ (crc >> 8);
    }
    contrib = ~crc;
    if ((uint32_t)tid < tl) contrib = gf_mul(c_crc.seg[tl - 1u - (uint32_t)tid], contrib); // past the whole segments after it
  }
  // XOR-reduce the whole segments, shift them past the last one, add it
  uint32_t whole = ((uint32_t)tid < tl) ? contrib : 0u;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) whole ^= __shfl_xor_sync(0xffffffffu, whole, d);
  if ((tid & 31) == 0) s_red[tid >> 5] = whole;
  __syncthreads();
  if ((uint32_t)tid == tl) {
    uint32_t w = 0u;
#pragma unroll
    for (int i = 0; i < CRC_NT / 32; i++) w ^= s_red[i];
    if (tl) w = gf_mul(gf_xpow8(c_crc.x2n, rem), w);
    part[(size_t)f * max_chunks + c] = w ^ contrib;
  }
}

// one warp per frame: combine the chunk CRCs, write the header (server.c:206-214, all fields big-endian)
__global__ void __launch_bounds__(32) k_crc32c_finish(const uint32_t *part, int max_chunks, const uint32_t *out_len,
                                                      uint32_t width, uint32_t height, uint8_t *headers,
                                                      size_t header_pitch) {
  const int f = blockIdx.x, lane = threadIdx.x;
  const uint32_t L = out_len[f];
  const uint32_t nchunks = (L + CRC_CHUNK - 1u) / CRC_CHUNK;
  uint32_t acc = 0u;
  for (uint32_t c = lane; c < nchunks; c += 32) {
    const uint64_t end = (uint64_t)(c + 1u) * CRC_CHUNK;
    const uint64_t after = end < L ? L - end : 0u;
    const uint32_t p = part[(size_t)f * max_chunks + c];
    acc ^= after ? gf_mul(gf_xpow8(c_crc.x2n, after), p) : p;
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) acc ^= __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane < 6) {
    const uint32_t v = lane == 0 ? width : lane == 1 ? height : lane == 2 ? L : lane == 4 ? acc : 0u;
    uint8_t *h = headers + (size_t)f * header_pitch + 4 * lane;
    h[0] = (uint8_t)(v >> 24);
    h[1] = (uint8_t)(v >> 16);
    h[2] = (uint8_t)(v >> 8);
    h[3] = (uint8_t)v;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The row form (default).  The segment form above walks 256 contiguous bytes per thread through a byte table: four
// DEPENDENT lookups per word, and every lane of a warp reads a different 128-byte line.  Here a warp reads whole
// 512-byte rows (lane l the 16 bytes at 16 l: coalesced) and every one of its 128 word positions is its own stream —
// a stream sees one word per row, 4096 bits apart, so its recurrence is  t' = t * x^4096 + w  (mod P): multiplying a
// 32-bit state by a fixed power of x is linear, i.e. FOUR INDEPENDENT lookups (one table per state byte) and three
// XORs, and the four streams of a thread are independent chains.  The x^4096 tables are lane-private in shared memory
// (entry b of lane l in bank l: 4 x 32 KB, conflict-free).  At the end of a run of rows the 128 stream states are
// folded: inside a thread by Horner with x^32, across lanes by a butterfly whose level d multiplies the earlier half by
// x^(128 d) — shared 1 KB tables, five levels.  The work is split by ROWS over all warps of a persistent grid (a
// warp's share may straddle frames); a run's remainder, shifted past the full rows of its frame that follow it
// (constant-memory table of x^(4096 k)), is XORed into the frame's accumulator.  k_crc32c_tail (one warp per frame)
// adds the < 512 bytes behind the last full row, applies the init / final complement and writes the header.
//   words: acc[n_frames] | prefix[n_frames + 1]  (prefix = exclusive scan of full rows per frame)
// The integer pipes are the bound here (first capture: alu pipe 80 % busy, profiles/r02p_ncu_crc_rows_summary.txt):
// LOP3 / SHF / PRMT / LEA all issue on the alu pipe at one warp instruction per two cycles, IMAD on the fma pipe at the
// same rate.  A lookup therefore costs one PRMT (cut the state byte out: alu) and one IMAD (byte * 128 + table address:
// fma; the factor lives in a register so that it stays a multiply), the four results fold with two three-input LOP3.
__device__ __forceinline__ uint32_t lds_at(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
template <int OFF, bool PIN> __device__ __forceinline__ uint32_t lds_off(uint32_t addr) {
  uint32_t v;
  if (PIN) // volatile: keeps its place among the (volatile) row loads, see k_crc32c_rows<false>
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  else
    asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
__device__ __forceinline__ uint32_t mad_lo(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// Al = shared-space BYTE address of (x^4096 table + lane); c128 = 128 in a register
template <bool PIN> __device__ __forceinline__ uint32_t z_private(uint32_t Al, uint32_t c128, uint32_t t) {
  const uint32_t a0 = mad_lo(__byte_perm(t, 0u, 0x4440), c128, Al);
  const uint32_t a1 = mad_lo(__byte_perm(t, 0u, 0x4441), c128, Al);
  const uint32_t a2 = mad_lo(__byte_perm(t, 0u, 0x4442), c128, Al);
  const uint32_t a3 = mad_lo(__byte_perm(t, 0u, 0x4443), c128, Al);
  return lds_off<0, PIN>(a0) ^ lds_off<32768, PIN>(a1) ^ lds_off<65536, PIN>(a2) ^ lds_off<98304, PIN>(a3);
}
__device__ __forceinline__ uint32_t z_shared(const uint32_t *S, uint32_t t) { // S = one 4 x 256 set
  return S[t & 255u] ^ S[256u + ((t >> 8) & 255u)] ^ S[512u + ((t >> 16) & 255u)] ^ S[768u + (t >> 24)];
}

__global__ void __launch_bounds__(1024) k_crc32c_plan(const uint32_t *out_len, int n_frames, uint32_t *words) {
  __shared__ uint32_t s_warp[32];
  __shared__ uint32_t s_carry;
  uint32_t *acc = words, *prefix = words + n_frames;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  if (tid == 0) s_carry = 0u;
  __syncthreads();
  for (int base = 0; base < n_frames; base += 1024) {
    const int f = base + tid;
    const uint32_t r = f < n_frames ? out_len[f] / CRC_ROW : 0u;
    if (f < n_frames) acc[f] = 0u;
    uint32_t inc = r;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    uint32_t before = s_carry;
    for (int w = 0; w < wid; w++) before += s_warp[w];
    if (f < n_frames) prefix[f] = before + inc - r;
    __syncthreads();
    if (tid == 1023) s_carry = before + inc;
    __syncthreads();
  }
  if (tid == 0) prefix[n_frames] = s_carry;
}

template <bool RING>
__global__ void __launch_bounds__(CRC_ROWS_NT, 1) k_crc32c_rows(const uint8_t *out, size_t out_pitch, int n_frames,
                                                               uint32_t *words, const uint32_t *slices,
                                                               uint8_t *copy_dst, size_t copy_pitch) {
  extern __shared__ __align__(16) uint32_t s_crc[];
  uint32_t *A = s_crc;                  // [4][256][32]: x^4096, lane-private
  uint32_t *S = s_crc + 4 * 256 * 32;   // [6][4][256]: x^32, x^128 .. x^2048
  uint4 *R = reinterpret_cast<uint4 *>(s_crc + 4 * 256 * 32 + 6 * 4 * 256); // [32 warps][4 slots][32 lanes]: row ring
  uint32_t *acc = words;
  const uint32_t *prefix = words + n_frames;
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t total = prefix[n_frames];
  const uint32_t nwarps = gridDim.x * (CRC_ROWS_NT / 32);
  uint32_t K = (total + nwarps - 1u) / nwarps;
  if (K < 4u) K = 4u; // short inputs: fewer, longer runs (the fold costs about as much as two rows)
  if ((uint64_t)blockIdx.x * (CRC_ROWS_NT / 32) * K >= total) return; // uniform per CTA: nothing to do, no table fill
  for (int i = tid; i < 4 * 256 * 32; i += CRC_ROWS_NT) A[i] = slices[6 * 1024 + (i >> 5)];
  for (int i = tid; i < 6 * 4 * 256; i += CRC_ROWS_NT) S[i] = slices[i];
  __syncthreads();
  const uint32_t Al = (uint32_t)__cvta_generic_to_shared(A + lane);
  uint32_t c128;
  asm volatile("mov.u32 %0, 128;" : "=r"(c128)); // opaque to the compiler: see z_private
  const uint32_t gw = blockIdx.x * (CRC_ROWS_NT / 32) + (tid >> 5);
  uint64_t s64 = (uint64_t)gw * K;
  if (s64 >= total) return;
  uint32_t s = (uint32_t)s64;
  const uint32_t e = (s64 + K < total) ? s + K : total;
  // frame that holds row s: the largest f with prefix[f] <= s (32-ary search, every lane probes)
  int lo = 0, hi = n_frames; // answer in [lo, hi)
  while (hi - lo > 1) {
    const int step = (hi - lo + 31) / 32;
    const int idx = lo + lane * step;
    const bool le = idx < hi && prefix[idx] <= s;
    const int cnt = __popc(__ballot_sync(0xffffffffu, le)); // lanes 0 .. cnt-1 (prefix is non-decreasing), cnt >= 1
    lo = lo + (cnt - 1) * step;
    hi = (lo + step < hi) ? lo + step : hi;
  }
  int f = lo;
  while (s < e) {
    while (prefix[f + 1] <= s) f++; // frames without a full row
    const uint32_t p0 = prefix[f], p1 = prefix[f + 1];
    const uint32_t r0 = s - p0, r1 = (e < p1 ? e : p1) - p0; // rows [r0, r1) of frame f
    const uint4 *src = reinterpret_cast<const uint4 *>(out + (size_t)f * out_pitch + (size_t)r0 * CRC_ROW) + lane;
    uint4 *dst = copy_dst ? reinterpret_cast<uint4 *>(copy_dst + (size_t)f * copy_pitch + (size_t)r0 * CRC_ROW) + lane : nullptr;
    const int n = (int)(r1 - r0);
    uint32_t t0 = 0u, t1 = 0u, t2 = 0u, t3 = 0u;
    // Four rows in flight per warp.  RING: staged through a private ring in shared memory by cp.async — the copy of
    // row k + 4 is issued when row k is taken out, four steps (~2000 issue cycles with 32 warps) before it is needed,
    // and occupies no registers while it flies; costs a 512-byte shared-memory write and read per row in the l1tex
    // pipe, which the sixteen lookups per row already keep 60 % busy.  !RING (measurement knob): the rows are loaded
    // into a register ring.  A row's registers are read by the last XORs of its step, so its successor cannot be asked
    // for earlier, and ptxas moves all four loads of a group to the group's end even with volatile lookups — one step
    // of distance or less, 2.6 warps per issue waiting on the long scoreboard (profiles/r02o_ncu_crc_rows_summary.txt).
    if (RING) {
      uint4 *slot0 = R + ((tid >> 5) * 4) * 32 + lane; // this lane's 16 bytes of slot 0; slot j is 512 bytes on
      const uint32_t ring = (uint32_t)__cvta_generic_to_shared(slot0);
      auto fetch = [&](int k, int j) { // row k -> slot j = k & 3 (an empty group when the row does not exist: the count stays uniform)
        if (k < n)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring + (uint32_t)j * 512u), "l"(src + (size_t)k * 32)
                       : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      auto step = [&](int k, int j) {
        asm volatile("cp.async.wait_group 3;" ::: "memory"); // all but the three youngest groups: row k has landed
        const uint4 v = slot0[j * 32];
        fetch(k + 4, j);
        if (dst) dst[(size_t)k * 32] = v;
        t0 = z_private<false>(Al, c128, t0) ^ v.x;
        t1 = z_private<false>(Al, c128, t1) ^ v.y;
        t2 = z_private<false>(Al, c128, t2) ^ v.z;
        t3 = z_private<false>(Al, c128, t3) ^ v.w;
      };
      fetch(0, 0);
      fetch(1, 1);
      fetch(2, 2);
      fetch(3, 3);
      int k = 0;
      for (; k + 4 <= n; k += 4) {
        step(k, 0);
        step(k + 1, 1);
        step(k + 2, 2);
        step(k + 3, 3);
      }
      if (k < n) step(k, 0);
      if (k + 1 < n) step(k + 1, 1);
      if (k + 2 < n) step(k + 2, 2);
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
      const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
      auto load = [&](int k) { return k < n ? ld_stream16(src + (size_t)k * 32) : zero; };
      uint4 q0 = load(0), q1 = load(1), q2 = load(2), q3 = load(3);
      auto step = [&](uint4 &q, int k) { // consume row k, then ask for row k + 4 into the same registers
        const uint4 v = q;
        if (dst) dst[(size_t)k * 32] = v;
        t0 = z_private<true>(Al, c128, t0) ^ v.x;
        t1 = z_private<true>(Al, c128, t1) ^ v.y;
        t2 = z_private<true>(Al, c128, t2) ^ v.z;
        t3 = z_private<true>(Al, c128, t3) ^ v.w;
        q = load(k + 4);
      };
      int k = 0;
      for (; k + 4 <= n; k += 4) {
        step(q0, k);
        step(q1, k + 1);
        step(q2, k + 2);
        step(q3, k + 3);
      }
      if (k < n) step(q0, k);
      if (k + 1 < n) step(q1, k + 1);
      if (k + 2 < n) step(q2, k + 2);
    }
    // fold the 128 streams: words of a thread by Horner with x^32, lanes by a butterfly with x^128, x^256, ...
    uint32_t u = z_shared(S, z_shared(S, z_shared(S, t0) ^ t1) ^ t2) ^ t3;
#pragma unroll
    for (int lvl = 0; lvl < 5; lvl++) {
      const int d = 1 << lvl;
      const uint32_t o = __shfl_xor_sync(0xffffffffu, u, d);
      const bool upper = (lane & d) != 0;
      u = z_shared(S + (1 + lvl) * 1024, upper ? o : u) ^ (upper ? u : o);
    }
    u = z_shared(S, u); // remainder of the run's bytes * x^32 (state form of the table recurrence, start value 0)
    const uint32_t after = (p1 - p0) - r1; // full rows of this frame behind the run
    if (lane == 0) {
      const uint32_t shift = after < (uint32_t)CRC_ROWPOW ? c_crc.rowpow[after] : gf_xpow8(c_crc.x2n, (uint64_t)after * CRC_ROW);
      atomicXor(&acc[f], after ? gf_mul(shift, u) : u);
    }
    s = p0 + r1;
  }
}

// one warp per frame: the bytes behind the last full row, init / final complement, header (server.c:206-214)
__global__ void __launch_bounds__(32) k_crc32c_tail(const uint8_t *out, size_t out_pitch, const uint32_t *out_len,
                                                    const uint32_t *words, uint32_t width, uint32_t height,
                                                    uint8_t *headers, size_t header_pitch, uint8_t *copy_dst,
                                                    size_t copy_pitch) {
  __shared__ uint32_t T[256];
  const int f = blockIdx.x, lane = threadIdx.x;
  for (int i = lane; i < 256; i += 32) {
    uint32_t v = (uint32_t)i;
#pragma unroll
    for (int k = 0; k < 8; k++) v = (v >> 1) ^ ((v & 1u) ? CRC_POLY : 0u);
    T[i] = v;
  }
  __syncwarp();
  const uint32_t L = out_len[f];
  const uint32_t full = (L / CRC_ROW) * CRC_ROW, r = L - full; // r < 512: lane l takes tail bytes [16 l, 16 l + 16)
  const uint8_t *tail = out + (size_t)f * out_pitch + full;
  uint8_t *dst = copy_dst ? copy_dst + (size_t)f * copy_pitch + full : nullptr;
  uint32_t c = 0u;
  const uint32_t b0 = 16u * lane, b1 = b0 + 16u < r ? b0 + 16u : r;
  if (b0 < r) {
    // sixteen independent byte loads (none behind the string's end), then the byte recurrence out of registers
    const uint32_t cnt = b1 - b0;
    uint32_t bytes[16];
#pragma unroll
    for (uint32_t i = 0; i < 16u; i++) bytes[i] = i < cnt ? (uint32_t)tail[b0 + i] : 0u;
    uint32_t st = 0u;
#pragma unroll
    for (uint32_t i = 0; i < 16u; i++) {
      if (i < cnt) {
        if (dst) dst[b0 + i] = (uint8_t)bytes[i];
        st = T[(st ^ bytes[i]) & 255u] ^ (st >> 8);
      }
    }
    c = gf_mul(c_crc.bytepow[r - b1], st);
  }
  if (lane == 0) c ^= gf_mul(c_crc.bytepow[r], words[f]); // the full rows, shifted past the tail
  if (lane == 1) { // the initial value, shifted past the whole frame: x^(8 L) = x^(8 * 512 * rows) * x^(8 r)
    const uint32_t nrow = L / CRC_ROW;
    const uint32_t xr = nrow < (uint32_t)CRC_ROWPOW ? c_crc.rowpow[nrow] : gf_xpow8(c_crc.x2n, (uint64_t)nrow * CRC_ROW);
    c ^= gf_mul(gf_mul(xr, c_crc.bytepow[r]), 0xFFFFFFFFu);
  }
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, d);
  const uint32_t crc = ~c;
  if (lane < 6) {
    const uint32_t v = lane == 0 ? width : lane == 1 ? height : lane == 2 ? L : lane == 4 ? crc : 0u;
    uint8_t *h = headers + (size_t)f * header_pitch + 4 * lane;
    h[0] = (uint8_t)(v >> 24);
    h[1] = (uint8_t)(v >> 16);
    h[2] = (uint8_t)(v >> 8);
    h[3] = (uint8_t)v;
  }
}
cudaError_t crc_rows_opt_in() {
  cudaError_t e = cudaFuncSetAttribute(k_crc32c_rows<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CRC_ROWS_SMEM);
  return e != cudaSuccess ? e : cudaFuncSetAttribute(k_crc32c_rows<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CRC_ROWS_SMEM);
}

// stream.c:1085-1127 on the device: a frame that does not end in ESC[0m is cut after its last ESC[0m, if it has one.
// One CTA per frame, scanning backwards; almost always decided by the first four bytes looked at.
__global__ void __launch_bounds__(256) k_trailing_reset_fixup(uint8_t *out, size_t out_pitch, uint32_t *out_len) {
  __shared__ int s_found;
  const int f = blockIdx.x, tid = threadIdx.x;
  uint8_t *s = out + (size_t)f * out_pitch;
  const uint32_t L = out_len[f];
  if (L < 4u) return;
  auto is_reset = [&](uint32_t p) { return s[p] == 0x1b && s[p + 1] == '[' && s[p + 2] == '0' && s[p + 3] == 'm'; };
  if (is_reset(L - 4u)) return; // uniform: every thread reads the same bytes
  if (tid == 0) s_found = -1;
  __syncthreads();
  int found = -1;
  for (long long hi = (long long)L - 4; hi >= 0; hi -= 256) { // candidate starts hi, hi-1, ... in blocks of 256
    const long long p = hi - tid;
    if (p >= 0 && is_reset((uint32_t)p)) atomicMax(&s_found, (int)p);
    __syncthreads();
    found = s_found; // every thread reads the block's verdict before anyone can start the next block's atomics
    __syncthreads();
    if (found >= 0) break;
  }
  if (tid == 0 && found >= 0) {
    s[found + 4] = 0;
    out_len[f] = (uint32_t)found + 4u;
  }
}

} // namespace

namespace acb {

// words of device scratch per frame that launch_frame_packets needs (>= 4: the row form keeps 2 n + 1 words)
int max_crc_chunks(size_t frame_capacity) {
  const int c = (int)((frame_capacity + CRC_CHUNK - 1) / CRC_CHUNK);
  return c < 4 ? 4 : c;
}

int launch_reset_fixup(uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len, int n_frames, cudaStream_t st) {
  k_trailing_reset_fixup<<<(unsigned)n_frames, 256, 0, st>>>(d_out, out_pitch, d_out_len);
  ACB_CUDA(cudaGetLastError());
  count_launch();
  return E_OK;
}

// d_part: n_frames * max_chunks words of device scratch
int launch_frame_packets(const uint8_t *d_out, size_t out_pitch, const uint32_t *d_out_len, int n_frames, int max_chunks,
                         uint32_t width, uint32_t height, uint32_t *d_part, uint8_t *headers, size_t header_pitch,
                         uint8_t *copy_dst, size_t copy_pitch, cudaStream_t st) {
  if (crc_tables_init() != E_OK) return set_error(E_INVALID_STATE, "CUDA: CRC table upload failed");
  // measurement knob (never set in production): run the table recurrence on synthetic words, no global loads
  static const int crc_noload = getenv("ACB200_CRC_NOLOAD") ? atoi(getenv("ACB200_CRC_NOLOAD")) : 0;
  // Which form: the row form wins from ~20 4K-frame strings on (32 frames: 30.7 vs 35.6 us, 256 frames: 87 vs 178 us);
  // below that its three launches and the 128 KB table fill per CTA cost more than they save (one frame: 23.6 vs 18.5 us,
  // profiles/r02u_crc_small.txt), so small arenas — the server's one-frame packet path — keep the segment form.
  // acb200_set_crc_form / ACB200_CRC_KERNEL=rows|segments force one.
  int form = g_crc_form.load(std::memory_order_relaxed);
  if (form == 0) form = (uint64_t)n_frames * out_pitch >= (24u << 20) ? 1 : 2;
  const bool segments = form == 2;
  if (!segments && !crc_noload) {
    int dev = 0;
    cudaGetDevice(&dev);
    k_crc32c_plan<<<1, 1024, 0, st>>>(d_out_len, n_frames, d_part);
    ACB_CUDA(cudaGetLastError());
    // at least four rows per warp; one CTA per SM at most (persistent, 152 KB of tables each)
    const uint64_t rows_max = (uint64_t)n_frames * (out_pitch / CRC_ROW);
    uint64_t ctas = (rows_max + 4u * (CRC_ROWS_NT / 32) - 1u) / (4u * (CRC_ROWS_NT / 32));
    const uint64_t sms = (uint64_t)device_sms();
    if (ctas > sms) ctas = sms;
    if (ctas < 1) ctas = 1;
    // ACB200_CRC_ROWS=regs (measurement knob): rows loaded into a register ring instead of the cp.async ring
    static const bool ring = !(getenv("ACB200_CRC_ROWS") && !strcmp(getenv("ACB200_CRC_ROWS"), "regs"));
    if (ring)
      k_crc32c_rows<true><<<(unsigned)ctas, CRC_ROWS_NT, CRC_ROWS_SMEM, st>>>(d_out, out_pitch, n_frames, d_part,
                                                                             g_crc_slices[dev], copy_dst, copy_pitch);
    else
      k_crc32c_rows<false><<<(unsigned)ctas, CRC_ROWS_NT, CRC_ROWS_SMEM, st>>>(d_out, out_pitch, n_frames, d_part,
                                                                              g_crc_slices[dev], copy_dst, copy_pitch);
    ACB_CUDA(cudaGetLastError());
    k_crc32c_tail<<<(unsigned)n_frames, 32, 0, st>>>(d_out, out_pitch, d_out_len, d_part, width, height, headers,
                                                    header_pitch, copy_dst, copy_pitch);
    ACB_CUDA(cudaGetLastError());
    count_launch(3);
    return E_OK;
  }
  for (int f0 = 0; f0 < n_frames; f0 += 65535) { // gridDim.y limit
    const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
    k_crc32c_chunks<<<dim3((unsigned)max_chunks, (unsigned)nf), CRC_NT, 0, st>>>(
        d_out + (size_t)f0 * out_pitch, out_pitch, d_out_len + f0, d_part + (size_t)f0 * max_chunks, max_chunks,
        copy_dst ? copy_dst + (size_t)f0 * copy_pitch : nullptr, copy_pitch, crc_noload);
    ACB_CUDA(cudaGetLastError());
  }
  k_crc32c_finish<<<(unsigned)n_frames, 32, 0, st>>>(d_part, max_chunks, d_out_len, width, height, headers, header_pitch);
  ACB_CUDA(cudaGetLastError());
  count_launch(2);
  return E_OK;
}

} // namespace acb

// ====================================================================== rainbow replace on a finished string
// rainbow_replace_ansi_colors (lib/video/rgba/color_filter.c:348-408): every "ESC[38;2;...m" of the string is replaced
// by one fixed colour code.  The reference scans serially: find the next "ESC[38;2;", copy what precedes it, skip to
// the first 'm' at or after match + 7, continue behind it.  As cell-local rules over byte positions:
//   * a match at p "owns" the bytes p .. e(p), e(p) = first 'm' at >= p + 7;
//   * a match that starts inside an earlier match's span shares its e() and is swallowed (the pattern holds no 'm', so
//     it cannot straddle the span's end): a match is ACTIVE iff the nearest earlier match has an 'm' in [p' + 7, p);
//   * byte i is dropped iff the nearest match p <= i has no 'm' in [p + 7, i); an active match emits the new code.
// Two running maxima (last match start, last 'm') and one running sum (output offset) — three small launches over
// 4 KB chunks: per-chunk maxima, per-chunk output counts, write.  A match with no 'm' anywhere behind it (malformed
// tail) is copied unchanged; the reference's loop re-copies its prefix once per byte there (color_filter.c:398-400).
namespace {

constexpr int RR_NT = 256, RR_PER = 16, RR_CHUNK = RR_NT * RR_PER;

struct RainbowParams {
  const uint8_t *in;
  uint32_t n;
  uint8_t *out;
  int *chunk_match, *chunk_m; // [chunks] last match start / last 'm' inside the chunk, -1 = none
  uint32_t *chunk_count;      // [chunks] output bytes of the chunk
  int *last_m_total;          // last 'm' of the whole string
  uint32_t *out_len;          // [0] = output length, [1] = 1 if anything was replaced
  uint8_t code[24];
  uint32_t code_len;
};

__device__ __forceinline__ bool rr_match(const uint8_t *s, uint32_t n, uint32_t i) {
  return i + 7u <= n && s[i] == 0x1b && s[i + 1] == '[' && s[i + 2] == '3' && s[i + 3] == '8' && s[i + 4] == ';' &&
         s[i + 5] == '2' && s[i + 6] == ';';
}
__device__ __forceinline__ int rr_block_max(int v, int *s_red) { // all threads get the block maximum
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
  __syncthreads();
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  int r = s_red[0];
#pragma unroll
  for (int k = 1; k < RR_NT / 32; k++) r = max(r, s_red[k]);
  return r;
}

__global__ void __launch_bounds__(RR_NT) k_rr_summary(const RainbowParams p) {
  __shared__ int s_red[RR_NT / 32];
  const uint32_t b0 = blockIdx.x * RR_CHUNK + threadIdx.x * RR_PER;
  int lm = -1, ll = -1;
  for (uint32_t i = b0; i < b0 + RR_PER && i < p.n; i++) {
    if (rr_match(p.in, p.n, i)) lm = (int)i;
    if (p.in[i] == 'm') ll = (int)i;
  }
  lm = rr_block_max(lm, s_red);
  ll = rr_block_max(ll, s_red);
  if (threadIdx.x == 0) {
    p.chunk_match[blockIdx.x] = lm;
    p.chunk_m[blockIdx.x] = ll;
    if (ll >= 0) atomicMax(p.last_m_total, ll);
  }
}

template <bool WRITE> __global__ void __launch_bounds__(RR_NT) k_rr_apply(const RainbowParams p) {
  __shared__ int s_red[RR_NT / 32];
  __shared__ int s_wm[RR_NT / 32], s_wl[RR_NT / 32];
  __shared__ uint32_t s_ws[RR_NT / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x;
  // carry-in from the chunks before this one
  int cm = -1, cl = -1;
  uint32_t base = 0;
  for (int k = tid; k < c; k += RR_NT) {
    cm = max(cm, p.chunk_match[k]);
    cl = max(cl, p.chunk_m[k]);
    if (WRITE) base += p.chunk_count[k];
  }
  cm = rr_block_max(cm, s_red);
  cl = rr_block_max(cl, s_red);
  if (WRITE) { // block sum of the counts before this chunk
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) base += __shfl_xor_sync(0xffffffffu, base, d);
    __syncthreads();
    if (lane == 0) s_ws[warp] = base;
    __syncthreads();
    base = 0;
    for (int k = 0; k < RR_NT / 32; k++) base += s_ws[k];
    __syncthreads();
  }
  const int last_m_total = *p.last_m_total;
  const uint32_t b0 = (uint32_t)c * RR_CHUNK + (uint32_t)tid * RR_PER;
  // this thread's own maxima, then an exclusive max-scan over the threads of the chunk
  int lm = -1, ll = -1;
  for (uint32_t i = b0; i < b0 + RR_PER && i < p.n; i++) {
    if (rr_match(p.in, p.n, i)) lm = (int)i;
    if (p.in[i] == 'm') ll = (int)i;
  }
  int im = lm, il = ll;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int om = __shfl_up_sync(0xffffffffu, im, d), ol = __shfl_up_sync(0xffffffffu, il, d);
    if (lane >= d) {
      im = max(im, om);
      il = max(il, ol);
    }
  }
  if (lane == 31) {
    s_wm[warp] = im;
    s_wl[warp] = il;
  }
  int em = __shfl_up_sync(0xffffffffu, im, 1), el = __shfl_up_sync(0xffffffffu, il, 1);
  if (lane == 0) em = el = -1;
  __syncthreads();
  for (int k = 0; k < warp; k++) {
    em = max(em, s_wm[k]);
    el = max(el, s_wl[k]);
  }
  int curM = max(em, cm), curL = max(el, cl); // last match start <= b0 - 1, last 'm' <= b0 - 1
  auto wf = [&](int q) { return last_m_total >= q + 7; };
  // pass 1: output bytes of this thread's 16 positions
  uint32_t cnt = 0;
  uint32_t kind[RR_PER]; // 0 drop, 1 copy, 2 emit the code
  {
    int M = curM, L = curL;
#pragma unroll
    for (int k = 0; k < RR_PER; k++) {
      const uint32_t i = b0 + k;
      uint32_t kd = 0;
      if (i < p.n) {
        const uint8_t ch = p.in[i];
        bool active = false;
        if (rr_match(p.in, p.n, i)) {
          active = wf((int)i) && !(M >= 0 && wf(M) && L < M + 7);
          M = (int)i;
        }
        const bool region = M >= 0 && wf(M) && L < M + 7;
        kd = active ? 2u : region ? 0u : 1u;
        if (ch == 'm') L = (int)i;
      }
      kind[k] = kd;
      cnt += kd == 2u ? p.code_len : kd;
    }
  }
  // exclusive sum-scan of cnt over the chunk
  uint32_t inc = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += o;
  }
  if (lane == 31) s_ws[warp] = inc;
  __syncthreads();
  uint32_t pre = 0, total = 0;
  for (int k = 0; k < RR_NT / 32; k++) {
    if (k < warp) pre += s_ws[k];
    total += s_ws[k];
  }
  if (!WRITE) {
    if (tid == 0) p.chunk_count[c] = total;
    return;
  }
  uint32_t o = base + pre + inc - cnt;
#pragma unroll
  for (int k = 0; k < RR_PER; k++) {
    if (kind[k] == 1u) p.out[o++] = p.in[b0 + k];
    else if (kind[k] == 2u) {
      for (uint32_t j = 0; j < p.code_len; j++) p.out[o + j] = p.code[j];
      o += p.code_len;
    }
  }
  if (c == (int)gridDim.x - 1 && tid == RR_NT - 1) {
    p.out[base + total] = 0;
    p.out_len[0] = base + total;
    p.out_len[1] = (max(cm, p.chunk_match[c]) >= 0) ? 1u : 0u;
  }
}

} // namespace

// ====================================================================== C ABI
extern "C" {

// lib/video/rgba/color_filter.c:348-408.  NULL when ansi_string is NULL or holds no "ESC[38;2;" (the reference's
// "no replacement needed"), else an allocator-owned string.
char *rainbow_replace_ansi_colors(const char *ansi_string, float time_seconds) {
  if (!ansi_string) return nullptr;
  const size_t n = strlen(ansi_string);
  if (n == 0 || n > 0x7fffff00u) return nullptr;
  ThreadCtx *cx = thread_ctx();
  if (!cx) return nullptr;
  RainbowParams p{};
  uint8_t r, g, b;
  color_filter_calculate_rainbow(time_seconds, &r, &g, &b);
  p.code_len = (uint32_t)snprintf(reinterpret_cast<char *>(p.code), sizeof(p.code), "\x1b[38;2;%d;%d;%dm", r, g, b);
  const int chunks = (int)((n + RR_CHUNK - 1) / RR_CHUNK);
  // every match is at least 8 bytes ("ESC[38;2;" + 'm') and becomes at most 19
  const size_t out_cap = n + (n / 8 + 1) * 11 + 32;
  const size_t words = (size_t)3 * chunks + 8;
  if (sync_foreign(cx, cx->stream) != E_OK) return nullptr;
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, n + 16) || !grow_device(&cx->d_in, &cx->d_in_cap, n + 16) ||
      !grow_device(&cx->d_out, &cx->d_out_cap, out_cap) || !grow_pinned(&cx->h_out, &cx->h_out_cap, out_cap) ||
      !grow_device((uint8_t **)&cx->d_len, &cx->d_len_cap, words * 4) ||
      !grow_pinned((uint8_t **)&cx->h_len, &cx->h_len_cap, 256))
    return nullptr;
  memcpy(cx->h_in, ansi_string, n);
  int *w = reinterpret_cast<int *>(cx->d_len);
  p.in = cx->d_in;
  p.n = (uint32_t)n;
  p.out = cx->d_out;
  p.last_m_total = w;
  p.out_len = reinterpret_cast<uint32_t *>(w + 2);
  p.chunk_match = w + 8;
  p.chunk_m = w + 8 + chunks;
  p.chunk_count = reinterpret_cast<uint32_t *>(w + 8 + 2 * chunks);
  bool ok = cudaMemcpyAsync(cx->d_in, cx->h_in, n, cudaMemcpyHostToDevice, cx->stream) == cudaSuccess &&
            cudaMemsetAsync(w, 0xFF, 4, cx->stream) == cudaSuccess;
  if (ok) {
    k_rr_summary<<<(unsigned)chunks, RR_NT, 0, cx->stream>>>(p);
    k_rr_apply<false><<<(unsigned)chunks, RR_NT, 0, cx->stream>>>(p);
    k_rr_apply<true><<<(unsigned)chunks, RR_NT, 0, cx->stream>>>(p);
    count_launch(3);
    ok = cudaGetLastError() == cudaSuccess &&
         cudaMemcpyAsync(cx->h_len, p.out_len, 8, cudaMemcpyDeviceToHost, cx->stream) == cudaSuccess &&
         wait_stream(cx) == E_OK;
  }
  if (!ok) {
    set_error(E_INVALID_STATE, "rainbow_replace_ansi_colors: CUDA failure (%s)", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  if (!cx->h_len[1]) return nullptr; // no colour code in the string: nothing to replace (color_filter.c:363-365)
  const size_t len = cx->h_len[0];
  if (cudaMemcpyAsync(cx->h_out, cx->d_out, len, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      wait_stream(cx) != E_OK) {
    set_error(E_INVALID_STATE, "rainbow_replace_ansi_colors: CUDA failure while reading the result back");
    return nullptr;
  }
  char *res = (char *)user_alloc(len + 1);
  if (!res) return nullptr;
  memcpy(res, cx->h_out, len);
  res[len] = '\0';
  return res;
}


// lib/video/rgba/color_filter.c:165-236 — host float on purpose, expression for expression (like aspect_ratio):
// fmodf / floorf / fminf are exactly rounded, so the host compiler's result is the reference's.
void color_filter_calculate_rainbow(float time, uint8_t *r, uint8_t *g, uint8_t *b) {
  const float cycle_period = 3.5f;
  float phase = fmodf(time, cycle_period) / cycle_period;
  float hue = phase * 360.0f;
  float h = hue / 60.0f;
  int i = (int)floorf(h);
  float f = h - (float)i;
  float q = 1.0f - f;
  float t = f;
  switch (i % 6) {
  case 0: *r = 255; *g = (uint8_t)(t * 255.0f + 0.5f); *b = 0; break;
  case 1: *r = (uint8_t)(q * 255.0f + 0.5f); *g = 255; *b = 0; break;
  case 2: *r = 0; *g = 255; *b = (uint8_t)(t * 255.0f + 0.5f); break;
  case 3: *r = 0; *g = (uint8_t)(q * 255.0f + 0.5f); *b = 255; break;
  case 4: *r = (uint8_t)(t * 255.0f + 0.5f); *g = 0; *b = 255; break;
  case 5: *r = 255; *g = 0; *b = (uint8_t)(q * 255.0f + 0.5f); break;
  default: *r = 255; *g = 0; *b = 0; break;
  }
  const float min_luminance = 120.0f;
  float luminance = 0.2126f * *r + 0.7152f * *g + 0.0722f * *b;
  if (luminance < min_luminance) {
    float boost = (min_luminance - luminance) / 3.0f;
    *r = (uint8_t)fminf(255.0f, *r + boost);
    *g = (uint8_t)fminf(255.0f, *g + boost);
    *b = (uint8_t)fminf(255.0f, *b + boost);
  }
}

int acb200_color_filter_device(uint8_t *d_pixels, uint32_t width, uint32_t height, uint32_t stride, int filter,
                               float time, void *stream) {
  if (!d_pixels || width == 0 || height == 0 || stride == 0) return -1; // color_filter.c:276
  if (filter == 0) return 0;
  int mode;
  uint32_t rgb;
  if (!resolve_pixel_filter(filter, time, &mode, &rgb)) return -1; // :326-329
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    ThreadCtx *cx = thread_ctx();
    if (!cx) return -1;
    st = cx->stream;
  } else if (ensure_device() != 0) {
    return -1;
  }
  return filter_device(d_pixels, width, height, stride, mode, rgb, st) == E_OK ? 0 : -1;
}

int apply_color_filter(uint8_t *pixels, uint32_t width, uint32_t height, uint32_t stride, int filter, float time) {
  if (!pixels || width == 0 || height == 0 || stride == 0) return -1;
  if (filter == 0) return 0;
  int mode;
  uint32_t rgb;
  if (!resolve_pixel_filter(filter, time, &mode, &rgb)) return -1;
  ThreadCtx *cx = thread_ctx();
  if (!cx) return -1;
  const size_t bytes = (size_t)stride * (height - 1) + (size_t)width * 3; // the last row has no trailing stride
  if (!grow_pinned(&cx->h_in, &cx->h_in_cap, bytes) || !grow_device(&cx->d_in, &cx->d_in_cap, bytes + 48)) return -1;
  memcpy(cx->h_in, pixels, bytes);
  if (cudaMemcpyAsync(cx->d_in, cx->h_in, bytes, cudaMemcpyHostToDevice, cx->stream) != cudaSuccess ||
      filter_device(cx->d_in, width, height, stride, mode, rgb, cx->stream) != E_OK ||
      cudaMemcpyAsync(cx->h_in, cx->d_in, bytes, cudaMemcpyDeviceToHost, cx->stream) != cudaSuccess ||
      wait_stream(cx) != E_OK) {
    set_error(E_INVALID_STATE, "apply_color_filter: CUDA failure (%s)", cudaGetErrorString(cudaGetLastError()));
    return -1;
  }
  if (stride == width * 3u) {
    memcpy(pixels, cx->h_in, bytes);
  } else { // only the pixel bytes of each row were touched on the device; copy just those back
    for (uint32_t y = 0; y < height; y++)
      memcpy(pixels + (size_t)y * stride, cx->h_in + (size_t)y * stride, (size_t)width * 3);
  }
  return 0;
}

// src/common/session/display.c:484-671 (flip -> filter -> convert -> rainbow), one fused pass
char *acb200_display_convert(const image_t *image, ssize_t width, ssize_t height, const terminal_capabilities_t *caps,
                             bool preserve_aspect_ratio, bool stretch, const char *palette_chars, bool flip_x,
                             bool flip_y, int color_filter, float time_seconds) {
  if (image == nullptr || caps == nullptr) { // ascii.c:198
    set_error(E_INVALID_PARAM, "Invalid parameters for acb200_display_convert");
    return nullptr;
  }
  if (image->w > 0 && image->w <= 10000 && image->h > 0 && image->h <= 10000 && image->pixels == nullptr) { // :209
    set_error(E_INVALID_PARAM, "Original image pixels pointer is NULL");
    return nullptr;
  }
  acb200_render_cfg_t cfg;
  if (!plan_convert_with_caps(image->w, image->h, width, height, caps, preserve_aspect_ratio, stretch, palette_chars,
                              &cfg))
    return nullptr;
  cfg.flip_x = flip_x ? 1 : 0;
  cfg.flip_y = flip_y ? 1 : 0;
  cfg.color_filter = color_filter;
  cfg.filter_time = time_seconds;
  return render_one_host(cfg, reinterpret_cast<const uint8_t *>(image->pixels), nullptr);
}

int acb200_trailing_reset_fixup_device(uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len, int n_frames,
                                       void *stream) {
  if (!d_out || !d_out_len || n_frames < 0) return set_error(E_INVALID_PARAM, "acb200_trailing_reset_fixup_device: bad argument");
  if (n_frames == 0) return E_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (!st) {
    ThreadCtx *cx = thread_ctx();
    if (!cx) return acb200_last_error();
    st = cx->stream;
  } else if (ensure_device() != 0) {
    return acb200_last_error();
  }
  return launch_reset_fixup(d_out, out_pitch, d_out_len, n_frames, st);
}

void acb200_set_crc_form(int form) {
  if (form >= 0 && form <= 2) g_crc_form.store(form);
}

int acb200_frame_packets_device(const uint8_t *d_out, size_t out_pitch, const uint32_t *d_out_len, int n_frames,
                                uint32_t width, uint32_t height, uint8_t *d_headers, void *stream) {
  if (!d_out || !d_out_len || !d_headers || n_frames < 0 || out_pitch == 0)
    return set_error(E_INVALID_PARAM, "acb200_frame_packets_device: bad argument");
  if ((reinterpret_cast<uintptr_t>(d_out) & 15u) || (out_pitch & 15u))
    return set_error(E_INVALID_PARAM, "acb200_frame_packets_device: arena base and pitch must be multiples of 16");
  if (n_frames == 0) return E_OK;
  ThreadCtx *cx = thread_ctx();
  if (!cx) return acb200_last_error();
  cudaStream_t st = stream ? (cudaStream_t)stream : cx->stream;
  const int mc = max_crc_chunks(out_pitch);
  // the chunk CRCs live in this thread's context: work queued on a different stream by the previous call must have
  // drained before they are overwritten (same stream = ordered anyway; cudaFree inside grow_device synchronises)
  if (sync_foreign(cx, st) != E_OK) return acb200_last_error();
  if (!grow_device((uint8_t **)&cx->d_len, &cx->d_len_cap, (size_t)(16 + (size_t)n_frames * mc) * sizeof(uint32_t)))
    return acb200_last_error();
  return launch_frame_packets(d_out, out_pitch, d_out_len, n_frames, mc, width, height, cx->d_len + 16, d_headers,
                              ACB200_FRAME_HEADER_BYTES, nullptr, 0, st);
}

} // extern "C"
