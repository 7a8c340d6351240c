// common.cuh -- task descriptors shared by the kernels and the host engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <stdexcept>
#include <string>

#include "records.h"

// A failed CUDA call never takes the host process down (the reference "never fails": a decoder that
// cannot work answers decode() == false).  CUDA_CHECK throws; every C-ABI entry point catches, marks
// its decoder dead (jsmpeg_b200_batch_last_error) and returns its failure value (engine.cu).
struct CudaFailure : std::runtime_error {
	cudaError_t code;
	CudaFailure(cudaError_t c, const std::string &what) : std::runtime_error(what), code(c) {}
};
[[noreturn]] inline void throw_cuda_failure(cudaError_t e, const char *file, int line) {
	char msg[512];
	snprintf(msg, sizeof msg, "jsmpeg_b200: CUDA error %s at %s:%d: %s", cudaGetErrorName(e), file, line, cudaGetErrorString(e));
	throw CudaFailure(e, msg);
}
#define CUDA_CHECK(expr)                                                  \
	do {                                                                  \
		cudaError_t _e = (expr);                                          \
		if (_e != cudaSuccess) throw_cuda_failure(_e, __FILE__, __LINE__); \
	} while (0)

// Per-stream sequence parameters as the kernels need them (reference mpeg1.js:78-153).
struct SeqParams {
	int32_t mb_width, mb_height, mb_size;
	int32_t coded_width, coded_height;
	uint8_t intra_q[64];     // de-zigzagged (mpeg1.js:100-116)
	uint8_t non_intra_q[64];
	// the same two matrices in COEFFICIENT (zig-zag) order, as stage 1b wants them: entry n of table k
	// (0 intra, 1 non-intra) = (raster index of coefficient n) * 2 | Q[raster index] << 8 (seq_fill_xq)
	uint16_t xq[2][64];
};

// host helper: derive SeqParams::xq from the de-zigzagged matrices (ZIG_ZAG, mpeg1.js:996-1005)
inline void seq_fill_xq(SeqParams &sp) {
	static const uint8_t zz[64] = {0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
	                               41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22,
	                               15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
	for (int n = 0; n < 64; n++) {
		sp.xq[0][n] = (uint16_t)(zz[n] * 2 | sp.intra_q[zz[n]] << 8);
		sp.xq[1][n] = (uint16_t)(zz[n] * 2 | sp.non_intra_q[zz[n]] << 8);
	}
}

// Stage 1: one warp parses one picture.
struct ParseTask {
	const uint8_t *es;     // 16-byte aligned base of the stream's ES mirror in HBM, >= 64 readable bytes past es_len
	uint32_t es_len;       // valid bytes (bytes past it read as zero, like a JS typed array)
	uint32_t start_byte;   // first byte after the 00 00 01 00 picture start code
	const SeqParams *seq;  // device pointer
	mb_record_t *hdr;      // [mb_size]
	int16_t *coef;         // [mb_size][6][64]
	picture_info_t *info;  // out
	uint2 *park;           // [mb_size][6]: per coded block {bit offset of its first coefficient code, intra dc * 8},
	                       // the walk's hand-over to stage 1b (a dense side array, written and read as a stream)
	int32_t mb_width, mb_size;  // copies of the sequence parameters (no dependent load in front of the walk)
	// lane-parallel walk: staging area for relative macroblock records (walk.cuh, WALK_STAGE): stage_entries
	// entries of 64 bytes, shared out among the lanes; nullptr = no staging (the walk then makes a second,
	// storing pass over the bits)
	uint4 *stage;
	int32_t stage_entries;
	// the stream's start-code prefixes (byte positions of 00 00 01, sorted, all of them up to es_len) and the
	// index of this picture's own start code in that list: a slice ends at the first prefix at or after its
	// first macroblock, so the walk looks the position up instead of searching the bytes.  nullptr = search.
	const uint32_t *codes;
	uint32_t n_codes, code_hint;
};

// staging entries per picture: three per macroblock (a lane keeps its macroblocks in its own stretch and a
// sub-sequence of small macroblocks holds more than the average; a lane that runs out falls back to the
// second pass) + one per lane for rounding
constexpr int stage_entries_for(int mb_size) { return 3 * mb_size + 64; }

struct PlaneSet {
	uint8_t *y, *cr, *cb;
};

// Stage 2: one task per (stream, picture); blockIdx.y selects the task.
struct ReconTask {
	const mb_record_t *hdr;
	const int16_t *coef;
	PlaneSet cur;   // written
	PlaneSet fwd;   // read (previous I/P picture; B picture: the older of the two most recent I/P pictures)
	int32_t mb_width, mb_size;
	int32_t coded_width, coded_height;
	uint8_t *rgba;  // optional fused epilogue target (display size, RGBA8888) or nullptr
	int32_t width, height;
	PlaneSet bwd;   // B picture (the opt-in extension): the newer of the two most recent I/P pictures
};

// kernel launchers (defined in scan.cu / parse.cu / recon.cu)
// Stage 0, all streams of a batch in one launch: span k = bytes [from, len) of the ES mirror `es`; its hits
// (byte positions of 00 00 01, unsorted) go to positions[k * capacity ...], their number to counts[k] (it keeps
// counting past `capacity`: the host sees an overflow and repeats the scan with room).
struct ScanSpan {
	const uint8_t *es;
	uint32_t from, len;
};
void launch_scan_start_codes(const ScanSpan *spans, int n_spans, uint32_t longest_span, uint32_t *positions,
                             uint32_t capacity, uint32_t *counts, cudaStream_t stream);
// Helper streams/events with which stage 1 forks the (size-sorted) wave into groups: the expand of
// a group of small pictures runs while the walk of the bigger pictures is still going.
constexpr int PARSE_GROUPS = 8;
struct ParseFork {
	cudaStream_t side[PARSE_GROUPS];
	cudaEvent_t fork, join[PARSE_GROUPS];
};
void launch_parse_pictures(const ParseTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream,
                           cudaEvent_t walk_done = nullptr, const ParseFork *fork = nullptr);
// how many size groups (walk + expand launch pairs) launch_parse_pictures uses for n_tasks pictures
int parse_group_count(int n_tasks, bool forked);
// `tasks_host` is read on the host at launch time: the table travels in the kernel parameters
void launch_reconstruct(const ReconTask *tasks_host, int n_tasks, cudaStream_t stream);
// the same with the planar -> RGBA conversion fused in (tasks' rgba / width / height)
void launch_reconstruct_rgba(const ReconTask *tasks_host, int n_tasks, cudaStream_t stream);
// B pictures (every task has `bwd`): prediction from two references; `rgba` = with the fused conversion.
// Returns the number of kernels launched.
int launch_reconstruct_b(const ReconTask *tasks_host, int n_tasks, bool rgba, cudaStream_t stream);
// stage 1 for B pictures: the serial walk with the B-picture macroblock layer (walk_b.cuh) + the same expand
void launch_parse_pictures_b(const ParseTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream);
// stage 1 for I/P pictures of many slices (opt-in, walk_slices.cuh): a lane per slice + the same expand
void launch_parse_pictures_slices(const ParseTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream);
// the host routes an I/P picture to it when it has at least this many slices
constexpr int SLICE_WALK_MIN_SLICES = 4;

// device MPEG-TS demux (tsdemux.cu)
struct TsScratch;
TsScratch *ts_scratch_create();
void ts_scratch_destroy(TsScratch *s);
long ts_demux_measure(TsScratch *s, const uint8_t *ts_host, size_t ts_bytes, int stream_id, int16_t *bound, size_t *consumed, cudaStream_t st);
int ts_demux_gather(TsScratch *s, uint8_t *es, uint32_t es_base, uint64_t *pts_out, uint32_t *offset_out, int n_max, cudaStream_t st);
