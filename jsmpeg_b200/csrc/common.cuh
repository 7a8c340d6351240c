// common.cuh -- task descriptors shared by the kernels and the host engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "records.h"

#define CUDA_CHECK(expr)                                                                         \
	do {                                                                                         \
		cudaError_t _e = (expr);                                                                 \
		if (_e != cudaSuccess) {                                                                 \
			fprintf(stderr, "jsmpeg_b200: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e),   \
			        __FILE__, __LINE__, cudaGetErrorString(_e));                                 \
			abort();                                                                             \
		}                                                                                        \
	} while (0)

// Per-stream sequence parameters as the kernels need them (reference mpeg1.js:78-153).
struct SeqParams {
	int32_t mb_width, mb_height, mb_size;
	int32_t coded_width, coded_height;
	uint8_t intra_q[64];     // de-zigzagged (mpeg1.js:100-116)
	uint8_t non_intra_q[64];
};

// Stage 1: one warp parses one picture.
struct ParseTask {
	const uint8_t *es;     // 4-byte aligned base of the stream's ES mirror in HBM
	uint32_t es_len;       // valid bytes (bytes past it read as zero, like a JS typed array)
	uint32_t start_byte;   // first byte after the 00 00 01 00 picture start code
	const SeqParams *seq;  // device pointer
	mb_record_t *hdr;      // [mb_size], pre-zeroed (no MBF_PRESENT)
	int16_t *coef;         // [mb_size][6][64]
	picture_info_t *info;  // out
};

struct PlaneSet {
	uint8_t *y, *cr, *cb;
};

// Stage 2: one task per (stream, picture); blockIdx.y selects the task.
struct ReconTask {
	const mb_record_t *hdr;
	const int16_t *coef;
	PlaneSet cur;   // written
	PlaneSet fwd;   // read (previous I/P picture)
	int32_t mb_width, mb_size;
	int32_t coded_width, coded_height;
	uint8_t *rgba;  // optional fused epilogue target (display size, RGBA8888) or nullptr
	int32_t width, height;
};

// kernel launchers (defined in scan.cu / parse.cu / recon.cu)
void launch_scan_start_codes(const uint8_t *es, uint32_t from, uint32_t len, uint32_t *positions,
                             uint32_t capacity, uint32_t *count, cudaStream_t stream);
// Helper streams/events with which stage 1 forks the (size-sorted) wave into groups: the expand of
// a group of small pictures runs while the walk of the bigger pictures is still going.
constexpr int PARSE_GROUPS = 8;
struct ParseFork {
	cudaStream_t side[PARSE_GROUPS];
	cudaEvent_t fork, join[PARSE_GROUPS];
};
void launch_parse_pictures(const ParseTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream,
                           cudaEvent_t walk_done = nullptr, const ParseFork *fork = nullptr);
// `tasks_host` is read on the host at launch time: the table travels in the kernel parameters
void launch_reconstruct(const ReconTask *tasks_host, int n_tasks, cudaStream_t stream);
void launch_rgba(const ReconTask *tasks, int n_tasks, int max_width, int max_height, cudaStream_t stream);

// device MPEG-TS demux (tsdemux.cu)
struct TsScratch;
TsScratch *ts_scratch_create();
void ts_scratch_destroy(TsScratch *s);
long ts_demux_measure(TsScratch *s, const uint8_t *ts_host, size_t ts_bytes, int stream_id, uint8_t *bound, cudaStream_t st);
int ts_demux_gather(TsScratch *s, size_t ts_bytes, uint8_t *es, uint32_t es_base, uint64_t *pts_out,
                    uint32_t *offset_out, int n_max, cudaStream_t st);
