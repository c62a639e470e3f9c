/*
 * asciichat_b200.h — C-ABI of libasciichat_b200.so
 *
 * A Blackwell (sm_100a) implementation of zfogg/ascii-chat's per-frame
 * RGB -> glyph + ANSI render path behind libasciichat's own entry points.
 * Plain C, plain pointers and sizes; no CUDA or torch types appear here.
 *
 * Part 1 re-declares, with identical names / argument meaning / ownership / error
 * behaviour, the reference entry points this library replaces (the file:line each one
 * replaces is cited; paths are relative to the reference checkout).  A host build that
 * already includes the reference headers defines ASCIICHAT_B200_NO_TYPES and gets only
 * the prototypes it does not have (Part 2).
 *
 * Part 2 is the batch interface the drop-in calls are built on: frames resident in
 * HBM, many frames per launch, device-side timing for the roofline numbers.
 *
 * Every function fails loudly (NULL / negative code + acb200_last_error()) when no
 * CUDA device is usable.  There is no CPU fallback in this library.
 */
#ifndef ASCIICHAT_B200_H
#define ASCIICHAT_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <sys/types.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */

/* ======================================================================== Part 1
 * types crossing the boundary (layout-identical to the reference's)            */
#ifndef ASCIICHAT_B200_NO_TYPES

/* include/ascii-chat/video/rgba/image.h:81-85 — packed RGB24 */
typedef struct {
  uint8_t r, g, b;
} __attribute__((packed)) rgb_pixel_t;

/* include/ascii-chat/video/rgba/image.h:143-148 */
typedef struct {
  int w, h;
  rgb_pixel_t *pixels;  /* w*h pixels, row-major, pitch 3*w */
  uint8_t alloc_method; /* untouched by this library */
} image_t;

/* include/ascii-chat/platform/terminal.h:578-589 */
typedef enum {
  TERM_COLOR_AUTO = -1,
  TERM_COLOR_NONE = 0,
  TERM_COLOR_16 = 1,
  TERM_COLOR_256 = 2,
  TERM_COLOR_TRUECOLOR = 3
} terminal_color_mode_t;

/* include/ascii-chat/platform/terminal.h:660-667 */
typedef enum { RENDER_MODE_FOREGROUND = 0, RENDER_MODE_BACKGROUND = 1, RENDER_MODE_HALF_BLOCK = 2 } render_mode_t;

/* include/ascii-chat/platform/terminal.h:707-738 (sizeof == 240 on LP64).
 * Fields read by the render path: color_level, render_mode, wants_padding. */
typedef struct {
  terminal_color_mode_t color_level;
  uint32_t capabilities;
  uint32_t color_count;
  bool utf8_support;
  bool detection_reliable;
  render_mode_t render_mode;
  char term_type[64];
  char colorterm[64];
  bool wants_background;
  int palette_type;
  char palette_custom[64];
  uint8_t desired_fps;
  int color_filter; /* color_filter_t */
  bool wants_padding;
  size_t pad_height;
} terminal_capabilities_t;

/* include/ascii-chat/video/ascii/ascii.h:358-361 */
typedef struct {
  const char *frame_data;
  size_t frame_size;
} ascii_frame_source_t;

/* include/ascii-chat/video/anim/digital_rain.h — same layout; previous_brightness is DEVICE memory in this library */
typedef struct {
  float time_offset;
  float speed_multiplier;
  float phase_offset;
} digital_rain_column_t;
typedef struct {
  digital_rain_column_t *columns; /* host, num_columns entries */
  int num_columns;
  int num_rows;
  float time;
  float fall_speed;
  float raindrop_length;
  float brightness_decay;
  float animation_speed;
  uint8_t color_r;
  uint8_t color_g;
  uint8_t color_b;
  float cursor_brightness;
  bool rainbow_mode;
  bool first_frame;
  float *previous_brightness; /* num_rows x num_columns filtered brightness, resident on the GPU */
} digital_rain_t;

#endif /* ASCIICHAT_B200_NO_TYPES */

/* ---- drop-in entry points.  Inputs are borrowed and never modified; every returned
 * string is NUL-terminated and allocated with the allocator set by acb200_set_allocator()
 * (default malloc), to be released by the caller's SAFE_FREE.  NULL on error (the caller
 * skips the frame, src/server/render.c:629); the error code (ERROR_INVALID_PARAM = 86,
 * ERROR_MEMORY = 3, ERROR_INVALID_STATE = 85) is readable through acb200_last_error() and is
 * forwarded to asciichat_set_errno_with_message() when the host binary defines it. */

/* lib/video/ascii/ascii.c:72-191 (include/ascii-chat/video/ascii/ascii.h:172).
 * GET_OPTION(render_mode) is supplied through acb200_set_option_render_mode(). */
char *ascii_convert(image_t *original, const ssize_t width, const ssize_t height, const bool color,
                    const bool aspect_ratio, const bool stretch, const char *palette_chars,
                    const char luminance_palette[256]);

/* lib/video/ascii/ascii.c:194-387 (ascii.h:213) */
char *ascii_convert_with_capabilities(image_t *original, const ssize_t width, const ssize_t height,
                                      const terminal_capabilities_t *caps, const bool use_aspect_ratio,
                                      const bool stretch, const char *palette_chars);

/* lib/video/ascii/ascii.c:955-1002 (ascii.h:230) — prints an already-resized image */
char *image_print_with_capabilities(const image_t *image, const terminal_capabilities_t *caps, const char *palette);

/* lib/video/rgba/image.c:256-328 (image.h:574) — nearest-neighbour resize into dest->pixels */
void image_resize(const image_t *source, image_t *dest);

/* leaf printers, lib/video/ascii/scalar/foreground.c:27,195,433,535,650,752 and halfblock.c:48,184,297,416 */
char *image_print(const image_t *p, const char *palette);
char *image_print_color(const image_t *p, const char *palette);
char *image_print_256color(const image_t *image, const char *palette);
char *image_print_16color(const image_t *image, const char *palette);
char *image_print_16color_dithered(const image_t *image, const char *palette);                     /* :650-749 */
char *image_print_16color_dithered_with_background(const image_t *image, bool use_background, const char *palette);
char *rgb_to_truecolor_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes);
char *rgb_to_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes, const char *palette);
char *rgb_to_16color_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes,
                                       const char *palette);
char *rgb_to_256color_halfblocks_scalar(const uint8_t *rgb, int width, int height, int stride_bytes,
                                        const char *palette);

/* ascii_pad_frame_width / ascii_pad_frame_height (ascii.c:457-517, 902-941) have no caller outside
 * ascii_convert*(); their work (left pad per row, leading newlines) is fused into the render kernels
 * (acb200_render_cfg_t.pad_left / pad_top), so they are not re-exported. */

/* lib/video/ascii/ascii.c:602-885 (ascii.h:400) — text-space grid of N rendered frames */
char *ascii_create_grid(ascii_frame_source_t *sources, int source_count, int width, int height, size_t *out_size);

/* lib/video/rgba/color_filter.c:274-346 (color_filter.h) — monochrome tint of an RGB24 image, IN PLACE on the
 * host buffer (H2D, k_color_filter, D2H).  0 on success, -1 on NULL / zero size / unknown filter, like the
 * reference.  filter = color_filter_t (0 none .. 12 rainbow, include/ascii-chat/platform/terminal.h:601-627). */
int apply_color_filter(uint8_t *pixels, uint32_t width, uint32_t height, uint32_t stride, int filter, float time);
/* lib/video/rgba/color_filter.c:348-408 — every truecolor-foreground SGR of a finished string replaced by the rainbow
 * colour at time_seconds (H2D, three scan launches, D2H).  NULL when the string is NULL or holds no "ESC[38;2;".
 * (acb200_display_convert fuses the same replacement into the emitters; this is the entry for a caller that already
 * holds a string, src/common/session/display.c:640-649.) */
char *rainbow_replace_ansi_colors(const char *ansi_string, float time_seconds);
/* lib/video/rgba/color_filter.c:165-236 — hue of the rainbow filter at `time` (host float, like aspect_ratio) */
void color_filter_calculate_rainbow(float time, uint8_t *r, uint8_t *g, uint8_t *b);

/* lib/video/anim/digital_rain.c (include/ascii-chat/video/anim/digital_rain.h) — the Matrix rain laid over a finished
 * frame string (src/common/session/display.c:657-671).  The string work (tokenising, the per-cell brightness filter
 * against the state kept on the GPU, colour scaling, re-emission) runs on the device; the brightness field itself is
 * evaluated on the host with libm's sinf, in the reference's order of operations — its values are truncated into colour
 * components, so byte-exactness means the reference's own sinf (see csrc/rain.cu).  filter = color_filter_t. */
digital_rain_t *digital_rain_init(int num_columns, int num_rows);
void digital_rain_destroy(digital_rain_t *rain);
char *digital_rain_apply(digital_rain_t *rain, const char *frame, float delta_time);
void digital_rain_reset(digital_rain_t *rain);
void digital_rain_set_fall_speed(digital_rain_t *rain, float speed);
void digital_rain_set_raindrop_length(digital_rain_t *rain, float length);
void digital_rain_set_color(digital_rain_t *rain, uint8_t r, uint8_t g, uint8_t b);
void digital_rain_set_color_from_filter(digital_rain_t *rain, int filter);

/* lib/video/ascii/common.c:601-604, 497-538 — table init / teardown (device LUT cache here) */
void ascii_simd_init(void);
void simd_caches_destroy_all(void);

/* ======================================================================== Part 2
 * B200 batch interface                                                         */

enum { ACB200_SCALE_NN = 0,   /* reference-exact nearest neighbour (image.c:267-328) */
       ACB200_SCALE_BOX = 1 }; /* full-coverage box filter: reads every source pixel */

/* One render configuration shared by every frame of a batch. */
typedef struct {
  int src_w, src_h;  /* source frame, packed RGB24, pitch 3*src_w */
  int cols;          /* resized width  = text columns */
  int rows_px;       /* resized height in pixels (text rows, or 2*text rows in half-block mode) */
  int color_level;   /* terminal_color_mode_t */
  int render_mode;   /* render_mode_t */
  int scale;         /* ACB200_SCALE_* */
  int pad_left;      /* spaces in front of every text row (ascii_pad_frame_width) */
  int pad_top;       /* leading newlines (ascii_pad_frame_height) */
  const char *palette; /* NUL-terminated UTF-8, <= 255 glyphs */
  /* client display pre/post steps, fused (src/common/session/display.c:484-671); all zero = plain convert */
  int flip_x, flip_y;  /* mirror the source first (display.c:548-591; ignored unless src_w > 1 && src_h > 1) */
  int color_filter;    /* color_filter_t 0..12: 1..11 = apply_color_filter on the pixels before the convert
                          (display.c:609-624); 12 (rainbow) = rainbow_replace_ansi_colors on the result (:640-649) */
  float filter_time;   /* seconds, drives the rainbow hue (color_filter_calculate_rainbow) */
} acb200_render_cfg_t;

/* 0 on success, else an ERROR_* code.  device < 0 keeps the current device. */
int acb200_init(int device);
/* Several GPUs behind ONE process — the reference server is one process with a render thread per client
 * (src/server/render.c:340-652, 1201-1253).  Call before any other entry point: `devices` lists the CUDA ordinals of
 * the pool (NULL / n <= 0: every visible device).  Calling threads are leased a per-thread context (stream, staging) on
 * one device of the pool, round-robin, the first time they call in; acb200_bind_thread(k) pins the calling thread to the
 * k-th device instead (k < 0: round-robin).  Peer access is enabled between all devices of the pool.  Entry points that
 * take a caller-owned stream run on the calling thread's device: the stream must belong to it. */
int acb200_init_devices(const int *devices, int n);
int acb200_device_count(void);   /* devices in the pool (0 before initialisation) */
int acb200_device_at(int k);     /* CUDA ordinal of the k-th pool device, -1 if out of range */
int acb200_bind_thread(int k);
int acb200_thread_device(void);  /* CUDA ordinal the calling thread is leased to (leases one if it has none), -1 on error */
/* How a caller waits for its frame: 0 = spin (cudaStreamSynchronize), 1 = sleep on a blocking-sync event, 2 = poll for
 * spin_us microseconds, then sleep, 3 = poll a completion word the stream writes into mapped memory and sched_yield()
 * between polls.  Sleeping callers leave their cores to the other render threads at the price of the wake-up latency
 * (profiles/r02a_e2e_sweep_pixels.txt); yielding callers give the core up only while another thread wants it. */
void acb200_set_sync_mode(int mode, int spin_us);
/* Frames in page-locked memory.  The drop-in calls take frames from ordinary (pageable) memory: the calling thread then
 * gathers the pixels nearest-neighbour sampling will read (image.c:293-325) into the library's pinned staging, which is
 * most of the call's host time.  A host that keeps its frame buffers for a while — the server's per-client
 * video_frame_buffer_t double buffers (lib/video/rgba/video_frame.c), a capture ring — can page-lock them ONCE:
 * acb200_register_host_memory(base, bytes) (cudaHostRegister, portable + mapped: a multi-millisecond call, not for the
 * frame loop) makes every frame inside [base, base + bytes) readable by the GPUs.  Unregister before the memory is
 * freed.  acb200_set_fetch_depth(k) then lets up to k calls per GPU at a time have the DEVICE fetch the sampled source
 * rows (whole rows over PCIe: 12 x the gathered bytes at 4K, no core time), the others gather on their cores;
 * -1 = always, 0 = never (default).  Worth it where cores are scarcer than GPUs. */
int acb200_register_host_memory(void *base, size_t bytes);
int acb200_unregister_host_memory(void *base);
void acb200_set_fetch_depth(int depth);
/* Host arithmetic behind the copy-engine form of the fetch (ACB200_FETCH=ce), exported for the CPU tests: the source rows
 * nearest-neighbour sampling reads (image.c:294,315-317; mirrored and listed backwards with flip_y), in ascending order,
 * as *period interleaved arithmetic progressions of common *stride: row j of the list = first[j % period] +
 * (j / period) * stride.  *period = 0: no such decomposition with period <= 8 (first[] must hold 8 ints). */
int acb200_nn_row_schedule(int src_h, int rows, int flip_y, int *period, int *stride, int *first);
/* Where the host time of the drop-in / batch-host calls went since the last reset, summed over all calling threads:
 * out[0..3] = nanoseconds spent gathering/staging the input, enqueueing (copies + launches), waiting for the device,
 * copying the strings into allocator-owned memory; out[4] = frames.  reset != 0 clears the counters. */
void acb200_host_phase_stats(uint64_t out[5], int reset);
void acb200_shutdown(void);
int acb200_last_error(void);            /* thread-local, cleared on read */
const char *acb200_last_error_message(void);
void acb200_set_allocator(void *(*alloc_fn)(size_t), void (*free_fn)(void *));
void acb200_set_option_render_mode(int render_mode); /* stands in for GET_OPTION(render_mode), ascii.c:138,152 */
void acb200_set_default_scale(int scale);            /* what the drop-in calls use; default ACB200_SCALE_NN */

/* bytes one rendered frame can occupy at most (fixed pitch of the output arena), incl. NUL */
size_t acb200_frame_capacity(const acb200_render_cfg_t *cfg);
/* bytes of device scratch acb200_render_batch_device needs for n_frames */
size_t acb200_scratch_bytes(const acb200_render_cfg_t *cfg, int n_frames);

/* Render n_frames frames that are ALREADY RESIDENT in device memory.
 *   d_frames : n_frames * src_w*src_h*3 bytes, frame-major
 *   d_out    : n_frames * out_pitch bytes; frame f's string starts at f*out_pitch, NUL-terminated
 *   d_out_len: n_frames uint32 string lengths
 *   d_scratch: acb200_scratch_bytes() bytes
 *   stream   : a cudaStream_t cast to void*.  NULL = this thread's internal (non-blocking) stream, which does NOT
 *              synchronise with the legacy default stream: wait for it with acb200_synchronize().
 * Asynchronous with respect to the host.  Returns 0 or an ERROR_* code. */
int acb200_render_batch_device(const acb200_render_cfg_t *cfg, const uint8_t *d_frames, int n_frames, uint8_t *d_out,
                               size_t out_pitch, uint32_t *d_out_len, void *d_scratch, void *stream);

/* Block until the calling thread's internal stream has drained.  0 or an ERROR_* code. */
int acb200_synchronize(void);

/* Same work from HOST buffers (H2D and D2H inside the call): frames[i] points at src_w*src_h*3 host bytes;
 * out[i] receives an allocator-owned NUL-terminated string, out_len[i] its length. */
int acb200_render_batch_host(const acb200_render_cfg_t *cfg, const uint8_t *const *frames, int n_frames, char **out,
                             size_t *out_len);

/* Device-timed repetition of acb200_render_batch_device (CUDA events on the launch stream):
 * `iters` passes over the batch, returns total milliseconds in *ms_total and the dominant
 * kernel's summed milliseconds in *ms_kernel (events around that kernel only). */
int acb200_time_batch_device(const acb200_render_cfg_t *cfg, const uint8_t *d_frames, int n_frames, uint8_t *d_out,
                             size_t out_pitch, uint32_t *d_out_len, void *d_scratch, int iters, float *ms_total,
                             float *ms_kernel);

/* Server pixel-space grid compositor (src/server/stream.c:523-651 layout, 664-779 composite):
 * n sources (host RGB24) -> one width x 2*height composite written to out_rgb (host). */
int acb200_composite_host(const uint8_t *const *srcs, const int *ws, const int *hs, int n, int width, int height,
                          uint8_t *out_rgb, int *out_cols, int *out_rows);
/* layout only (host float arithmetic, stream.c:523-651) */
void acb200_grid_layout(const int *ws, const int *hs, int n, int term_w, int term_h, int *cols, int *rows);
/* aspect fit (host float arithmetic, lib/util/aspect_ratio.c:70-93) */
void acb200_aspect_ratio(ssize_t img_w, ssize_t img_h, ssize_t width, ssize_t height, bool stretch, ssize_t *out_w,
                         ssize_t *out_h);

/* parity aid: rgb_to_256color (which = 0) / rgb_to_16color (which = 1) as the kernels compute them, over the whole
 * colour space: d_out[r << 16 | g << 8 | b], 16 MiB of device memory */
int acb200_quantize_table_device(int which, uint8_t *d_out, void *stream);

/* number of kernel launches issued by this library since load (bench "gpu_launches") */
uint64_t acb200_launch_count(void);
const char *acb200_version(void);

/* Text-space grid from frames that are already on the device (the multi-GPU gather path):
 * d_srcs[i] = device pointer to source i (host array of n pointers), sizes[i] its length.
 * d_out must hold width*height + height + 1 bytes.  Same result as ascii_create_grid().
 * Returns 0 and the layout-dependent string length in *out_size. */
int acb200_create_grid_device(const uint8_t *const *d_srcs, const size_t *sizes, int n, int width, int height,
                              uint8_t *d_out, size_t *out_size, void *stream);

/* ---- the client's display conversion, one fused pass -------------------------------------------------------
 * Replaces the body of session_display_convert_to_ascii (src/common/session/display.c:484-671) between its option
 * reads and the digital-rain stage: copy+flip X/Y (:548-591), apply_color_filter on a copy (:609-624),
 * ascii_convert_with_capabilities (:632) and rainbow_replace_ansi_colors (:640-649).  Here the flips are index
 * arithmetic in the sampler, the filter is evaluated on the sampled pixels only, and the rainbow colour is printed
 * by the emitter: no image copies, no second pass over the string.  Same result bytes, same NULL/errno behaviour
 * as ascii_convert_with_capabilities. */
char *acb200_display_convert(const image_t *image, ssize_t width, ssize_t height, const terminal_capabilities_t *caps,
                             bool preserve_aspect_ratio, bool stretch, const char *palette_chars, bool flip_x,
                             bool flip_y, int color_filter, float time_seconds);
/* apply_color_filter on an image that is already resident in device memory (in place, asynchronous on `stream`,
 * NULL = this thread's internal stream).  6 bytes of HBM traffic per pixel. */
int acb200_color_filter_device(uint8_t *d_pixels, uint32_t width, uint32_t height, uint32_t stride, int filter,
                               float time, void *stream);

/* ---- wire packaging of finished frames -------------------------------------------------------------------
 * acip_send_ascii_frame (lib/network/acip/server.c:188-236) computes asciichat_crc32 over the frame (CRC32-C,
 * lib/network/crc32.c) and prepends the 24-byte big-endian ascii_frame_packet_t (packet.h:848-862:
 * width, height, original_size, compressed_size = 0, checksum, flags = 0).  Here the checksum is a scan over the
 * frame while it is still in HBM: d_headers[f] receives frame f's 24 header bytes. */
/* The server's "a frame must end in ESC[0m" cut (src/server/stream.c:1085-1127) on frames in a device arena:
 * a frame that does not end in ESC[0m is truncated after its last ESC[0m, if it has one (d_out_len updated). */
int acb200_trailing_reset_fixup_device(uint8_t *d_out, size_t out_pitch, uint32_t *d_out_len, int n_frames,
                                       void *stream);
#define ACB200_FRAME_HEADER_BYTES 24
/* Which CRC32-C kernels run: 0 = chosen by arena size (default: the row form from 24 MB of arena on, the segment form
 * below — the one-frame packet path), 1 = row form, 2 = segment form.  Same result either way; for tests and A/B. */
void acb200_set_crc_form(int form);
int acb200_frame_packets_device(const uint8_t *d_out, size_t out_pitch, const uint32_t *d_out_len, int n_frames,
                                uint32_t width, uint32_t height, uint8_t *d_headers, void *stream);
/* acb200_mixed_frame + the packet header in front: returns header||frame (allocator-owned, *out_size =
 * 24 + frame bytes, no NUL needed by the transport but one is appended), i.e. the buffer acip_send_ascii_frame
 * hands to packet_send_via_transport.  NULL with *out_size = 0 when no client sends video. */
uint8_t *acb200_mixed_frame_packet(const int *slots, int n, unsigned short width, unsigned short height,
                                   const terminal_capabilities_t *caps, const char *palette, size_t *out_size,
                                   int *out_sources_count);

/* ---- the server's per-client frame generation with RESIDENT sources ------------------------------------
 * Replaces src/server/stream.c:958-1191 create_mixed_ascii_frame_for_client() and what it calls
 * (collect_video_sources :221-455, create_single_source_composite :476-500, create_multi_source_composite
 * :664-779, convert_composite_to_ascii :789-853).  The reference reads every client's newest frame from its
 * host-side incoming_video_buffer for every (receiving client x output frame); here each received frame is
 * uploaded ONCE into a device-resident slot and all the per-client renders read it from HBM.
 *
 * slot = the client's index in g_client_manager.clients[] (0 .. ACB200_MAX_SOURCES-1 = MAX_CLIENTS-1,
 * include/ascii-chat/common/limits.h:26). */
#define ACB200_MAX_SOURCES 32
/* A client's receive thread stores its newest RGB24 frame (src/server/protocol.c image-frame handler ->
 * video_frame_commit).  Dimensions 0, > 4096 wide or > 2160 tall are what collect_video_sources rejects
 * (stream.c:342): the slot is cleared and ERROR_INVALID_PARAM returned.  Thread-safe against
 * acb200_mixed_frame(); returns when the frame is resident. */
int acb200_source_update(int slot, const uint8_t *rgb, int w, int h);
/* Same, from the wire form of the frame: IMAGE_FRAME payload [width:be32][height:be32][RGB24] as
 * handle_image_frame_packet receives it (src/server/protocol.c:737-889): payload >= 8 bytes, dimensions
 * 1..3840 x 1..2160 (image_validate_dimensions), len == 8 + w*h*3; anything else is ERROR_INVALID_PARAM and the slot
 * keeps its previous frame (the reference disconnects the client, which then calls acb200_source_clear). */
int acb200_source_update_wire(int slot, const uint8_t *payload, size_t len);
/* Receive without a staging copy: acb200_source_acquire returns a pinned host buffer of at least `bytes` owned by the
 * slot (valid until a later acquire needs a larger one, or acb200_shutdown); the transport receives the RGB24 payload of
 * one frame into it and acb200_source_commit uploads it (same checks and effects as acb200_source_update).
 * acb200_source_update() on a pointer inside that buffer takes the same path. */
uint8_t *acb200_source_acquire(int slot, size_t bytes);
int acb200_source_commit(int slot, int w, int h);
/* CUDA ordinal the slot's frames live on (slot % pool size; -1 before its first update) */
int acb200_source_device(int slot);
/* client stopped sending video / disconnected (is_sending_video = false) */
int acb200_source_clear(int slot);
/* One output frame for one receiving client.  `slots` lists the active clients in g_client_manager order;
 * slots without a resident frame are the reference's "no video" clients.  0 sources with video: returns
 * NULL with *out_size = 0 and no error (stream.c:1036).  1 source: that frame is converted directly;
 * 2..9+: W x 2H pixel-space grid composite (first 9), then ascii_convert_with_capabilities(composite, width,
 * HALF_BLOCK ? 2*height : height, caps, aspect=true, stretch=false, palette) and the trailing-ESC[0m fix-up of
 * stream.c:1085-1127.  Returns an allocator-owned NUL-terminated string, *out_size its length,
 * *out_sources_count (optional) = sources with video.  Bit-identical to the reference for every input on
 * which the reference itself does not crash (a source whose fitted cell size rounds to 0 px NULL-derefs
 * there, stream.c:723-749; here that cell stays black). */
char *acb200_mixed_frame(const int *slots, int n, unsigned short width, unsigned short height,
                         const terminal_capabilities_t *caps, const char *palette, size_t *out_size,
                         int *out_sources_count);

/* ---- pixel-space composition from device pointers: the multi-process form of the server grid ("gather
 * the NN-resized cell images").  Every rank resizes its own clients' frames to acb200_mixed_cell_size() with
 * acb200_resize_nn_device(), the few-KB cell images are gathered (NCCL), and the composing rank calls
 * acb200_mixed_frame_device(prefit = 1): same bytes as acb200_mixed_frame on the full frames. */
int acb200_mixed_cell_size(const int *ws, const int *hs, int n, int i, unsigned short width, unsigned short height,
                           int *tw, int *th);
int acb200_resize_nn_device(const uint8_t *d_src, int sw, int sh, uint8_t *d_dst, int dw, int dh, void *stream);
char *acb200_mixed_frame_device(const uint8_t *const *d_srcs, const int *ws, const int *hs, int n, int prefit,
                                unsigned short width, unsigned short height, const terminal_capabilities_t *caps,
                                const char *palette, size_t *out_size);

/* ---- the discovery host's render tick with RESIDENT sources, over the GPUs of the pool ------------------------
 * Replaces the body of host_render_thread's video tick (src/common/session/host.c:664-717): for every listed slot that
 * has video, ascii_convert_with_capabilities(frame, cell_width, cell_height, caps, use_aspect_ratio, stretch, palette)
 * — rendered on the GPU that owns the slot (slot % pool size), its rows stored directly into the composing GPU's arena
 * over NVLink — then ascii_create_grid(sources, n_with_video, grid_width, grid_height, out_size) with
 * frame_size = strlen + 1 as host.c:701-702 passes it.  The composing GPU is the calling thread's.  Returns the
 * allocator-owned grid (NULL with *out_size = 0 and no error when no listed slot has video). */
char *acb200_grid_frame(const int *slots, int n, int cell_width, int cell_height, const terminal_capabilities_t *caps,
                        bool use_aspect_ratio, bool stretch, const char *palette, int grid_width, int grid_height,
                        size_t *out_size);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* ASCIICHAT_B200_H */
