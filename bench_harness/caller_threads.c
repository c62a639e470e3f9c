/*
 * bench_harness/caller_threads.c — neutral end-to-end driver used by bench.py for BOTH arms.
 *
 * Mirrors how the reference server calls the path (src/server/render.c:340,526: one render thread per
 * client, each calling ascii_convert_with_capabilities on its client's latest host frame): T pthreads,
 * thread t renders frames t, t+T, ... of a ring of host RGB24 frames through a function pointer with
 * the reference's signature, frees the result with free().  It knows nothing about either
 * implementation: bench.py passes libasciichat_b200's entry point for our arm and the compiled
 * reference's (oracle/_ref) for the CPU arm.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <time.h>

typedef struct { /* layout of the reference's image_t (include/ascii-chat/video/rgba/image.h:143-148) */
  int w, h;
  void *pixels;
  uint8_t alloc_method;
} harness_image_t;

typedef char *(*convert_fn)(harness_image_t *, ssize_t, ssize_t, const void *caps, bool, bool, const char *);

typedef struct {
  convert_fn fn;
  const uint8_t *frames;
  int ring, w, h, calls, threads, tid;
  ssize_t cols, rows;
  const void *caps;
  const char *palette;
  uint64_t bytes, failures;
} job_t;

static void *worker(void *vp) {
  job_t *j = (job_t *)vp;
  const size_t fsz = (size_t)j->w * j->h * 3;
  for (int i = j->tid; i < j->calls; i += j->threads) {
    harness_image_t img = {j->w, j->h, (void *)(j->frames + (size_t)(i % j->ring) * fsz), 0};
    char *s = j->fn(&img, j->cols, j->rows, j->caps, false, false, j->palette);
    if (s) {
      j->bytes += strlen(s);
      free(s);
    } else {
      j->failures++;
    }
  }
  return NULL;
}

/* returns wall seconds for `calls` renders spread over `threads` threads */
double harness_run(void *fn, const uint8_t *frames, int ring, int w, int h, long cols, long rows, const void *caps,
                   const char *palette, int calls, int threads, uint64_t *out_bytes, uint64_t *out_failures) {
  if (threads < 1) threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
  job_t *jobs = (job_t *)calloc((size_t)threads, sizeof(job_t));
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int t = 0; t < threads; t++) {
    jobs[t] = (job_t){(convert_fn)fn, frames, ring, w, h, calls, threads, t, cols, rows, caps, palette, 0, 0};
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  uint64_t bytes = 0, fails = 0;
  for (int t = 0; t < threads; t++) {
    pthread_join(th[t], NULL);
    bytes += jobs[t].bytes;
    fails += jobs[t].failures;
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  if (out_bytes) *out_bytes = bytes;
  if (out_failures) *out_failures = fails;
  free(th);
  free(jobs);
  return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* one untimed pass over the ring: FNV-1a-32 of every frame's output, combined order-independently, so the two
 * arms can be compared byte-for-byte through a single number */
uint64_t harness_ring_fingerprint(void *fn, const uint8_t *frames, int ring, int w, int h, long cols, long rows,
                                  const void *caps, const char *palette) {
  const size_t fsz = (size_t)w * h * 3;
  uint64_t acc = 0;
  for (int i = 0; i < ring; i++) {
    harness_image_t img = {w, h, (void *)(frames + (size_t)i * fsz), 0};
    char *s = ((convert_fn)fn)(&img, cols, rows, caps, false, false, palette);
    if (!s) return 0;
    uint32_t hsh = 2166136261u;
    for (const unsigned char *p = (const unsigned char *)s; *p; p++) {
      hsh ^= *p;
      hsh *= 16777619u;
    }
    acc += (uint64_t)hsh * (uint64_t)(i + 1);
    free(s);
  }
  return acc;
}
