// engine.h — internal host-side interface shared by engine.cu, dropin.cu and grid.cu.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/asciichat_b200.h"
#include "render.cuh"

namespace acb {

// reference error codes (include/ascii-chat/common/error_codes.h:51-109)
enum { E_OK = 0, E_MEMORY = 3, E_INVALID_STATE = 85, E_INVALID_PARAM = 86 };

// Everything derived from an acb200_render_cfg_t that the launches need.
struct Plan {
  int mode;        // EmitMode
  int scale_path;  // ScalePath
  int text_rows;
  int ring_depth;  // SP_BOX_TMA: rows in the producer ring
  int lut_which;   // 0 cache[Y], 1 mono double map (Q1), 2 16-colour map (Q2)
  uint32_t row_pitch;
  int use_smem_out;
  size_t frame_capacity; // bytes per frame in the output arena (multiple of 16, incl. NUL)
  size_t rows_bytes, meta_bytes, cells_bytes, err_bytes; // per frame scratch
};

int set_error(int code, const char *fmt, ...);
bool make_plan(const acb200_render_cfg_t &cfg, Plan &pl, int leaf_mode = -1); // false => error set; leaf_mode: EmitMode override
int ensure_device();                                       // 0 or error code

// look-back bookkeeping for a library-owned scratch buffer (zeroed once when allocated)
struct LookbackState {
  uint32_t epoch = 0, tickets = 0;
};

constexpr int kMaxDevices = 64; // CUDA ordinals this library can address

// what a thread context keeps on ANOTHER device of the pool, for work it farms out there (acb200_grid_frame renders
// every client's cell on the GPU that holds the client's frames)
struct PeerHelper {
  cudaStream_t stream = nullptr;
  cudaEvent_t ev = nullptr;
  uint8_t *scratch = nullptr; size_t scratch_cap = 0;
  uint8_t *out = nullptr;     size_t out_cap = 0;   // staging arena when the composing GPU is not peer-accessible
  LookbackState lb;
  bool dirty = true;
};

// per-thread stream + growable staging buffers; bound to ONE device of the pool for its whole life
struct ThreadCtx {
  int device = 0;            // CUDA ordinal this context lives on
  LookbackState lb;
  bool scratch_dirty = true; // d_scratch holds something other than zeros / look-back records
  cudaStream_t stream = nullptr;
  uint8_t *h_in = nullptr;   size_t h_in_cap = 0;   // pinned
  uint8_t *h_out = nullptr;  size_t h_out_cap = 0;  // pinned
  uint8_t *d_in = nullptr;   size_t d_in_cap = 0;
  uint8_t *d_out = nullptr;  size_t d_out_cap = 0;
  uint8_t *d_rows = nullptr; size_t d_rows_cap = 0; // source rows the copy engine fetched from page-locked frames
  uint8_t *d_scratch = nullptr; size_t d_scratch_cap = 0;
  uint32_t *d_len = nullptr; size_t d_len_cap = 0;   // CRC chunk words of acb200_frame_packets_device (effects.cu)
  uint8_t *d_frame = nullptr; size_t d_frame_cap = 0; // one-frame packet path: device arena + length + CRC chunk words
  uint32_t *h_len = nullptr; size_t h_len_cap = 0;  // pinned
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t done = nullptr;   // cudaEventBlockingSync: the waiting caller sleeps instead of spinning (wait_stream)
  // completion word in mapped pinned memory, written by the stream itself (cuStreamWriteValue32) behind the frame's
  // work: a waiting caller polls plain memory and yields its core between polls (wait_stream, sync mode 3)
  volatile uint32_t *h_flag = nullptr; uint32_t flag_seq = 0;
  cudaStream_t foreign = nullptr; // last caller-owned stream that used d_scratch / d_len (ordered by sync_foreign)
  // nearest-neighbour transfer plan (host side of image.c:293-325): byte offset of the source column each output
  // column samples, cached per (src_w, cols, flip_x)
  uint32_t *nn_off = nullptr; int nn_off_cap = 0, nn_src_w = 0, nn_cols = 0, nn_flip = -1;
  PeerHelper *helper[kMaxDevices] = {};
};
PeerHelper *peer_helper(ThreadCtx *cx, int device); // lazily created; makes `device` current.  nullptr => error set
ThreadCtx *thread_ctx(); // nullptr => error set.  Leases a context on this thread's device and makes that device current
// wait until everything queued on cx->stream has finished, in the way acb200_set_sync_mode selected
int wait_stream(ThreadCtx *cx);
// d_scratch / d_len are about to be used on `st`: first drain the caller-owned stream that used them last, if different
int sync_foreign(ThreadCtx *cx, cudaStream_t st);
int device_sms();                  // SM count of the current device
bool peer_ok(int from_dev, int to_dev); // kernels running on from_dev may dereference pointers of to_dev
bool grow_pinned(uint8_t **p, size_t *cap, size_t need);
bool grow_device(uint8_t **p, size_t *cap, size_t need);

const GlyphLut *device_lut(const char *palette, int which); // cached per (palette, which); nullptr => error set
void destroy_lut_cache();
void destroy_sources(); // server.cu: resident client frames

// apply_color_filter's arithmetic for `filter` (color_filter.c:274-346): FilterMode + 0x00RRGGBB; false = no-op/invalid
bool resolve_pixel_filter(int filter, float time, int *mode, uint32_t *rgb);

void *user_alloc(size_t n);
void user_free(void *p);
void count_launch(int n = 1);
int option_render_mode();
int default_scale();

// the core: frames resident on the device -> strings in the device arena (async on st)
int render_device(const acb200_render_cfg_t &cfg, const Plan &pl, const uint8_t *d_frames, size_t frame_stride,
                  int pregathered, int n_frames, uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len,
                  uint8_t *d_scratch, cudaStream_t st, cudaEvent_t k0 = nullptr, cudaEvent_t k1 = nullptr,
                  LookbackState *ls = nullptr);

// one frame from a host RGB24 buffer -> allocator-owned string (used by every drop-in entry point)
char *render_one_host(const acb200_render_cfg_t &cfg, const uint8_t *rgb, size_t *out_len, int leaf_mode = -1);
// reset_fixup: apply stream.c:1085-1127 (cut after the last ESC[0m) on the device.  packet: return header||frame
// (lib/network/acip/server.c:203-236) for a terminal of pk_w x pk_h; *out_len then counts the 24 header bytes too.
struct OneFrameOpts {
  bool reset_fixup = false, packet = false;
  uint32_t pk_w = 0, pk_h = 0;
};
char *render_one_device(const acb200_render_cfg_t &cfg, const uint8_t *d_rgb, size_t *out_len,
                        const OneFrameOpts &opts = OneFrameOpts());

// grid.cu: ascii_create_grid on device-resident sources (sizes on the host, or d_sizes[i] + size_bias on the device)
int text_grid_device(const uint8_t *const *d_srcs, const size_t *sizes, int n, int width, int height, uint8_t *d_out,
                     size_t *out_size, cudaStream_t st, ThreadCtx *cx, const uint32_t *d_sizes, uint32_t size_bias,
                     bool *size_is_exact); // *size_is_exact: report *out_size as is (else the caller reports strlen)
size_t scratch_bytes(const Plan &pl, int n);

// effects.cu
int max_crc_chunks(size_t frame_capacity);
int launch_reset_fixup(uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len, int n_frames, cudaStream_t st);
int launch_frame_packets(const uint8_t *d_out, size_t out_pitch, const uint32_t *d_out_len, int n_frames, int max_chunks,
                         uint32_t width, uint32_t height, uint32_t *d_part, uint8_t *headers, size_t header_pitch,
                         uint8_t *copy_dst, size_t copy_pitch, cudaStream_t st);
// ascii_convert_with_capabilities' front half (ascii.c:194-265): validation, aspect fit, padding -> a render cfg
bool plan_convert_with_caps(int w, int h, ssize_t width, ssize_t height, const terminal_capabilities_t *caps,
                            bool use_aspect_ratio, bool stretch, const char *palette, acb200_render_cfg_t *cfg);

#define ACB_CUDA(call)                                                                                                 \
  do {                                                                                                                 \
    cudaError_t _e = (call);                                                                                           \
    if (_e != cudaSuccess) return acb::set_error(acb::E_INVALID_STATE, "CUDA: %s (%s)", cudaGetErrorString(_e), #call); \
  } while (0)

} // namespace acb
