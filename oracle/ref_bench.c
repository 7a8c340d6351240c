/*
 * oracle/ref_bench.c -- TEST/BENCH INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A pthread harness around the UNMODIFIED reference decoder (reference src/wasm/mpeg1.c +
 * buffer.c, compiled in place by oracle/Makefile into oracle/_ref/libjsmpeg_ref.so).  It only
 * uses the reference's public 15-function ABI (reference src/wasm/mpeg1.h:10-25), exactly the
 * way the reference's own glue does (src/mpeg1-wasm.js:52-70 write, :103 decode): one decoder
 * per elementary stream, whole stream written once (EXPAND mode), then decode() until it
 * returns false.  This is the `cpu_baseline` / `--impl reference` arm of bench.py.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "mpeg1.h"

typedef struct {
	const uint8_t *const *es;
	const unsigned *es_len;
	int n_clips, first, stride, loops;
	long frames;
	uint64_t checksum;
} job_t;

static uint64_t fnv1a(uint64_t h, const uint8_t *p, size_t n) {
	for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ULL; }
	return h;
}

static void *worker(void *arg) {
	job_t *j = (job_t *)arg;
	for (int l = 0; l < j->loops; l++) {
		for (int c = j->first; c < j->n_clips; c += j->stride) {
			mpeg1_decoder_t *d = mpeg1_decoder_create(j->es_len[c] + 16, BIT_BUFFER_MODE_EXPAND);
			memcpy(mpeg1_decoder_get_write_ptr(d, j->es_len[c]), j->es[c], j->es_len[c]);
			mpeg1_decoder_did_write(d, j->es_len[c]);
			while (mpeg1_decoder_decode(d)) {
				j->frames++;
			}
			/* touch the last picture so the decode cannot be optimised away */
			if (mpeg1_decoder_has_sequence_header(d)) {
				j->checksum = fnv1a(j->checksum, (const uint8_t *)mpeg1_decoder_get_y_ptr(d), 64);
			}
			mpeg1_decoder_destroy(d);
		}
	}
	return 0;
}

/* Decode n_clips elementary streams `loops` times on `threads` host threads (clip c is handled
 * by thread c % threads).  Returns the number of pictures decode() reported; *seconds gets the
 * wall time of the threaded region (decoder create + write + all decode() calls). */
long ref_bench_run(const uint8_t *const *es, const unsigned *es_len, int n_clips,
                   int threads, int loops, double *seconds) {
	if (threads < 1) threads = 1;
	pthread_t *tid = (pthread_t *)calloc(threads, sizeof(pthread_t));
	job_t *jobs = (job_t *)calloc(threads, sizeof(job_t));
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (int t = 0; t < threads; t++) {
		jobs[t].es = es; jobs[t].es_len = es_len; jobs[t].n_clips = n_clips;
		jobs[t].first = t; jobs[t].stride = threads; jobs[t].loops = loops;
		jobs[t].checksum = 1469598103934665603ULL;
		pthread_create(&tid[t], 0, worker, &jobs[t]);
	}
	long frames = 0;
	for (int t = 0; t < threads; t++) { pthread_join(tid[t], 0); frames += jobs[t].frames; }
	clock_gettime(CLOCK_MONOTONIC, &t1);
	if (seconds) *seconds = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
	free(tid); free(jobs);
	return frames;
}
