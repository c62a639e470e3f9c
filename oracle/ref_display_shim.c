/*
 * oracle/ref_display_shim.c — TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's OWN functions (compiled into oracle/_ref from the sources where they lie) in the order
 * the client's display path calls them, session_display_convert_to_ascii (src/common/session/display.c:484-671):
 *
 *     copy + flip X / flip Y (display.c:548-591)  ->  apply_color_filter on a copy (:609-624)
 *       ->  ascii_convert_with_capabilities (:632)  ->  rainbow_replace_ansi_colors (:640-649)
 *
 * display.c itself cannot be compiled here (options/terminal/session state), so the call ORDER is restated;
 * every pixel and byte is still produced by reference code (apply_color_filter, color_filter_calculate_rainbow,
 * ascii_convert_with_capabilities, rainbow_replace_ansi_colors: lib/video/rgba/color_filter.c, lib/video/ascii).
 * The digital-rain stage (:652-671) is float/transcendental and stateful; it is out of scope.
 *
 * Also here: the wire packaging of a finished frame exactly as acip_send_ascii_frame builds it
 * (lib/network/acip/server.c:188-236): 24-byte big-endian ascii_frame_packet_t + frame bytes, checksum from the
 * reference's asciichat_crc32 (lib/network/crc32.c).
 */
#include <stdlib.h>
#include <string.h>

#include <ascii-chat/common.h>
#include <ascii-chat/network/crc32.h>
#include <ascii-chat/network/packet/packet.h>
#include <ascii-chat/util/endian.h>
#include <ascii-chat/video/ascii/ascii.h>
#include <ascii-chat/video/rgba/color_filter.h>
#include <ascii-chat/video/rgba/image.h>

char *ref_oracle_display_convert(const unsigned char *rgb, int w, int h, long width, long height,
                                 const terminal_capabilities_t *caps, int preserve_aspect, int stretch,
                                 const char *palette, int flip_x, int flip_y, int color_filter, float time_seconds,
                                 size_t *out_size) {
  image_t in = {.w = w, .h = h, .pixels = (rgb_pixel_t *)rgb, .alloc_method = 0};
  const image_t *display_image = &in;
  image_t flipped = in, filtered = in;
  flipped.pixels = NULL;
  filtered.pixels = NULL;
  size_t px = (size_t)w * (size_t)h;

  if ((flip_x || flip_y) && w > 1 && h > 1) { /* display.c:548 */
    flipped.pixels = malloc(px * sizeof(rgb_pixel_t));
    memcpy(flipped.pixels, rgb, px * sizeof(rgb_pixel_t));
    if (flip_x) /* :563-576 (scalar branch) */
      for (int y = 0; y < h; y++) {
        rgb_pixel_t *row = &flipped.pixels[(size_t)y * w];
        for (int x = 0; x < w / 2; x++) {
          rgb_pixel_t t = row[x];
          row[x] = row[w - 1 - x];
          row[w - 1 - x] = t;
        }
      }
    if (flip_y) /* :580-590 */
      for (int y = 0; y < h / 2; y++) {
        rgb_pixel_t *a = &flipped.pixels[(size_t)y * w], *b = &flipped.pixels[(size_t)(h - 1 - y) * w];
        for (int x = 0; x < w; x++) {
          rgb_pixel_t t = a[x];
          a[x] = b[x];
          b[x] = t;
        }
      }
    display_image = &flipped;
  }
  if (color_filter != COLOR_FILTER_NONE && color_filter != COLOR_FILTER_RAINBOW) { /* :609-624 */
    filtered.pixels = malloc(px * sizeof(rgb_pixel_t));
    memcpy(filtered.pixels, display_image->pixels, px * sizeof(rgb_pixel_t));
    apply_color_filter((uint8_t *)filtered.pixels, (uint32_t)w, (uint32_t)h, (uint32_t)w * 3, (color_filter_t)color_filter,
                       time_seconds);
    display_image = &filtered;
  }
  char *result = ascii_convert_with_capabilities((image_t *)display_image, width, height, caps, preserve_aspect != 0,
                                                 stretch != 0, palette); /* :632 */
  if (result && color_filter == COLOR_FILTER_RAINBOW) { /* :640-649 */
    char *r2 = rainbow_replace_ansi_colors(result, time_seconds);
    if (r2) {
      free(result);
      result = r2;
    }
  }
  free(flipped.pixels);
  free(filtered.pixels);
  if (out_size) *out_size = result ? strlen(result) : 0;
  return result;
}

/* lib/network/acip/server.c:203-214: the header acip_send_ascii_frame puts in front of the frame bytes */
void ref_oracle_frame_packet_header(const char *frame, size_t frame_size, uint32_t width, uint32_t height,
                                    unsigned char out24[24]) {
  ascii_frame_packet_t header;
  header.width = HOST_TO_NET_U32(width);
  header.height = HOST_TO_NET_U32(height);
  header.original_size = HOST_TO_NET_U32((uint32_t)frame_size);
  header.compressed_size = 0;
  header.checksum = HOST_TO_NET_U32(asciichat_crc32(frame, frame_size));
  header.flags = 0;
  memcpy(out24, &header, sizeof(header));
}

uint32_t ref_oracle_crc32(const void *data, size_t len) { return asciichat_crc32(data, len); }
uint32_t ref_oracle_crc32_sw(const void *data, size_t len) { return asciichat_crc32_sw(data, len); }
