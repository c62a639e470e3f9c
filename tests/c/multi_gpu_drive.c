/*
 * tests/c/multi_gpu_drive.c — ONE plain-C process drives every GPU of the box through the reference's own entry point.
 *
 * Mirrors the server's structure (src/server/render.c:340-652: one render thread per client inside one process):
 * T pthreads call ascii_convert_with_capabilities() on the same host frame; libasciichat_b200 leases each calling
 * thread a context on one device of the pool (acb200_init_devices), round-robin.  The program checks that every
 * device of the pool served at least one thread and that every thread got byte-identical output, and prints
 *     devices=<n> threads=<T> used=<k> fnv=<hex> len=<bytes>
 * for the Python test to compare against the compiled reference's fingerprint of the same frame.
 * No CUDA headers, no rendering code: it links against the C ABI only.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/asciichat_b200.h"

enum { W = 640, H = 360, COLS = 100, ROWS = 30, MAXT = 64 };

typedef struct {
  const uint8_t *rgb;
  int device;
  uint32_t fnv;
  size_t len;
  int calls, ok;
} job_t;

static uint32_t fnv1a(const char *s, size_t n) {
  uint32_t h = 2166136261u;
  for (size_t i = 0; i < n; i++) {
    h ^= (uint8_t)s[i];
    h *= 16777619u;
  }
  return h;
}

static void *worker(void *vp) {
  job_t *j = (job_t *)vp;
  terminal_capabilities_t caps;
  memset(&caps, 0, sizeof(caps));
  caps.color_level = TERM_COLOR_TRUECOLOR;
  caps.render_mode = RENDER_MODE_HALF_BLOCK;
  caps.utf8_support = true;
  image_t img = {W, H, (rgb_pixel_t *)j->rgb, 0};
  j->ok = 1;
  for (int i = 0; i < j->calls; i++) {
    char *s = ascii_convert_with_capabilities(&img, COLS, ROWS, &caps, false, false, "   ...',;:clodxkO0KXNWM");
    if (!s) {
      j->ok = 0;
      fprintf(stderr, "call failed: %s\n", acb200_last_error_message());
      break;
    }
    const size_t n = strlen(s);
    const uint32_t f = fnv1a(s, n);
    if (i == 0) {
      j->fnv = f;
      j->len = n;
    } else if (f != j->fnv || n != j->len) {
      j->ok = 0;
    }
    free(s);
  }
  j->device = acb200_thread_device();
  return NULL;
}

int main(int argc, char **argv) {
  int threads = argc > 1 ? atoi(argv[1]) : 16;
  if (threads < 1) threads = 1;
  if (threads > MAXT) threads = MAXT;
  int rc = acb200_init_devices(NULL, 0); /* every visible GPU */
  if (rc != 0) {
    fprintf(stderr, "acb200_init_devices failed: %d %s\n", rc, acb200_last_error_message());
    return 2;
  }
  const int ndev = acb200_device_count();
  uint8_t *rgb = (uint8_t *)malloc((size_t)W * H * 3);
  uint32_t s = 12345u; /* the survey's LCG noise */
  for (size_t i = 0; i < (size_t)W * H * 3; i++) {
    s = s * 1664525u + 1013904223u;
    rgb[i] = (uint8_t)(s >> 24);
  }
  pthread_t th[MAXT];
  job_t jobs[MAXT];
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){rgb, -1, 0, 0, 8, 0};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  int used[64] = {0}, nused = 0, ok = 1;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    ok &= jobs[t].ok && jobs[t].fnv == jobs[0].fnv && jobs[t].len == jobs[0].len;
    if (jobs[t].device >= 0 && jobs[t].device < 64 && !used[jobs[t].device]++) nused++;
  }
  printf("devices=%d threads=%d used=%d fnv=%08x len=%zu\n", ndev, threads, nused, jobs[0].fnv, jobs[0].len);
  free(rgb);
  acb200_shutdown();
  if (!ok) return 3;
  if (nused != (threads < ndev ? threads : ndev)) return 4; /* every device of the pool must have served a caller */
  return 0;
}
