// tsdemux.cu -- MPEG-TS demultiplexing on the device.
//
// Mirror of the reference's JSMpeg.Demuxer.TS (src/ts.js:25-210) for a buffer of whole 188-byte
// packets: sync byte 0x47 (:45), payload_unit_start / PID / adaptation_field_control (:52-58),
// adaptation field skip (:73-77), PES header on a payload start that begins with 00 00 01
// (:79-126: stream id, PES header length, 33-bit PTS), PID -> stream id binding from the first PES
// header of that PID on (:81-83), payload bytes appended in packet order (:191-197).
// What "packet complete" means in the reference (:65-70, :143-146, :201) only decides how the
// payload is CHUNKED into destination.write() calls; the decoder concatenates the chunks, so the
// elementary stream is the concatenation of the accepted payloads.
//
// 188-byte packets are independent: one thread classifies a packet, an exclusive scan of the
// payload lengths gives every packet its place in the elementary stream, one warp copies a packet's
// payload.  PES starts (offset in the ES, PTS) are appended to a list for the host's PTS table
// (src/decoder.js:36-47).
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace {

struct PacketUse {
	uint16_t start;  // first payload byte inside the packet
	uint16_t len;    // payload bytes that go to the elementary stream (0 = none)
};

// pass 1: which PIDs carry `stream_id`, and from which packet on (ts.js:81-83)
__global__ void ts_bind_kernel(const uint8_t *__restrict__ ts, uint32_t n_packets, int stream_id,
                               uint32_t *__restrict__ first_pusi /* [8192], 0xffffffff */, int *__restrict__ error) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_packets) return;
	const uint8_t *p = ts + (size_t)i * 188;
	if (p[0] != 0x47) { *error = 1; return; }  // not packet aligned: the host demuxer resyncs (ts.js:155-189)
	const int pusi = (p[1] >> 6) & 1;
	const int pid = ((p[1] & 0x1f) << 8) | p[2];
	const int afc = (p[3] >> 4) & 3;
	if (!pusi || !(afc & 1)) return;
	int at = 4;
	if (afc & 2) at += 1 + p[4];
	if (at + 9 <= 188 && p[at] == 0 && p[at + 1] == 0 && p[at + 2] == 1 && p[at + 3] == stream_id)
		atomicMin(&first_pusi[pid], i);
}

// pass 2: payload span of every packet + PES starts
__global__ void ts_measure_kernel(const uint8_t *__restrict__ ts, uint32_t n_packets, int stream_id,
                                  const uint32_t *__restrict__ first_pusi, PacketUse *__restrict__ use,
                                  uint32_t *__restrict__ lens, uint8_t *__restrict__ is_pes, uint64_t *__restrict__ pts) {
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_packets) return;
	const uint8_t *p = ts + (size_t)i * 188;
	const int pusi = (p[1] >> 6) & 1;
	const int pid = ((p[1] & 0x1f) << 8) | p[2];
	const int afc = (p[3] >> 4) & 3;
	PacketUse u{0, 0};
	uint8_t pes = 0;
	uint64_t t = 0;
	if ((afc & 1) && first_pusi[pid] <= i) {
		int at = 4;
		if (afc & 2) at += 1 + p[4];
		if (pusi && at + 9 <= 188 && p[at] == 0 && p[at + 1] == 0 && p[at + 2] == 1) {
			if (p[at + 3] == stream_id) {
				pes = 1;
				if (p[at + 7] & 0x80) {  // PTS present (ts.js:94-116)
					const uint8_t *q = p + at + 9;
					t = ((uint64_t)((q[0] >> 1) & 7) << 30) | ((uint64_t)((q[1] << 7) | (q[2] >> 1)) << 15) |
					    (uint64_t)((q[3] << 7) | (q[4] >> 1));
				}
				at += 9 + p[at + 8];
			} else {
				at = 188;  // the PID was re-bound to another stream id: not ours any more
			}
		}
		if (at < 188) { u.start = (uint16_t)at; u.len = (uint16_t)(188 - at); }
	}
	use[i] = u;
	lens[i] = u.len;
	is_pes[i] = pes;
	pts[i] = t;
}

// pass 3: one warp copies one packet's payload to its place; PES starts are appended to a list
__global__ void ts_gather_kernel(const uint8_t *__restrict__ ts, uint32_t n_packets, const PacketUse *__restrict__ use,
                                 const uint32_t *__restrict__ offsets, const uint8_t *__restrict__ is_pes,
                                 const uint64_t *__restrict__ pts, uint8_t *__restrict__ es, uint32_t es_base,
                                 uint32_t *__restrict__ pes_count, uint32_t pes_cap, uint32_t *__restrict__ pes_offset,
                                 uint64_t *__restrict__ pes_pts) {
	const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31;
	if (i >= n_packets) return;
	const PacketUse u = use[i];
	const uint32_t off = offsets[i];
	if (lane == 0 && is_pes[i]) {
		const uint32_t k = atomicAdd(pes_count, 1u);
		if (k < pes_cap) { pes_offset[k] = es_base + off; pes_pts[k] = pts[i]; }
	}
	const uint8_t *src = ts + (size_t)i * 188 + u.start;
	uint8_t *dst = es + es_base + off;
	for (int k = lane; k < u.len; k += 32) dst[k] = src[k];
}

}  // namespace

// Scratch (device): first_pusi[8192], use[n], lens[n], offsets[n], is_pes[n], pts[n], total, error, cub temp
struct TsScratch {
	uint8_t *ts = nullptr;
	size_t ts_cap = 0;
	uint32_t *first_pusi = nullptr;
	PacketUse *use = nullptr;
	uint32_t *lens = nullptr, *offsets = nullptr;
	uint8_t *is_pes = nullptr;
	uint64_t *pts = nullptr;
	uint32_t *pes_count = nullptr, *pes_offset = nullptr;
	uint64_t *pes_pts = nullptr;
	int *error = nullptr;
	void *cub_temp = nullptr;
	size_t cub_bytes = 0, n_cap = 0;
};

static void ts_reserve(TsScratch &s, size_t ts_bytes, size_t n) {
	if (ts_bytes > s.ts_cap) {
		if (s.ts) CUDA_CHECK(cudaFree(s.ts));
		s.ts_cap = ts_bytes + ts_bytes / 4;
		CUDA_CHECK(cudaMalloc(&s.ts, s.ts_cap));
	}
	if (!s.first_pusi) {
		CUDA_CHECK(cudaMalloc(&s.first_pusi, 8192 * sizeof(uint32_t)));
		CUDA_CHECK(cudaMalloc(&s.error, sizeof(int)));
		CUDA_CHECK(cudaMalloc(&s.pes_count, sizeof(uint32_t)));
	}
	if (n > s.n_cap) {
		void *old[] = {s.use, s.lens, s.offsets, s.is_pes, s.pts, s.pes_offset, s.pes_pts, s.cub_temp};
		for (void *p : old) if (p) CUDA_CHECK(cudaFree(p));
		s.n_cap = n + n / 4;
		CUDA_CHECK(cudaMalloc(&s.use, s.n_cap * sizeof(PacketUse)));
		CUDA_CHECK(cudaMalloc(&s.lens, (s.n_cap + 1) * sizeof(uint32_t)));
		CUDA_CHECK(cudaMalloc(&s.offsets, (s.n_cap + 1) * sizeof(uint32_t)));
		CUDA_CHECK(cudaMalloc(&s.is_pes, s.n_cap));
		CUDA_CHECK(cudaMalloc(&s.pts, s.n_cap * sizeof(uint64_t)));
		CUDA_CHECK(cudaMalloc(&s.pes_offset, s.n_cap * sizeof(uint32_t)));
		CUDA_CHECK(cudaMalloc(&s.pes_pts, s.n_cap * sizeof(uint64_t)));
		s.cub_bytes = 0;
		cub::DeviceScan::ExclusiveSum(nullptr, s.cub_bytes, s.lens, s.offsets, (int)(s.n_cap + 1));
		CUDA_CHECK(cudaMalloc(&s.cub_temp, s.cub_bytes));
	}
}

TsScratch *ts_scratch_create() { return new TsScratch(); }
void ts_scratch_destroy(TsScratch *s) {
	if (!s) return;
	void *all[] = {s->ts, s->first_pusi, s->use, s->lens, s->offsets, s->is_pes, s->pts, s->pes_count, s->pes_offset, s->pes_pts, s->error, s->cub_temp};
	for (void *p : all) if (p) cudaFree(p);
	delete s;
}

// Phase 1: upload + classify + scan.  Returns the ES byte count (or -1: not packet aligned).
long ts_demux_measure(TsScratch *s, const uint8_t *ts_host, size_t ts_bytes, int stream_id, uint8_t *bound /* [8192] in/out */,
                      cudaStream_t st) {
	const size_t n = ts_bytes / 188;
	if (n == 0) return 0;
	ts_reserve(*s, n * 188, n);
	CUDA_CHECK(cudaMemcpyAsync(s->ts, ts_host, n * 188, cudaMemcpyHostToDevice, st));
	// PIDs bound by earlier buffers of this stream stay bound (pidsToStreamIds lives across write() calls, ts.js:9)
	static thread_local uint32_t host_first[8192];
	for (int i = 0; i < 8192; i++) host_first[i] = bound[i] ? 0u : 0xffffffffu;
	CUDA_CHECK(cudaMemcpyAsync(s->first_pusi, host_first, sizeof(host_first), cudaMemcpyHostToDevice, st));
	CUDA_CHECK(cudaMemsetAsync(s->error, 0, sizeof(int), st));
	CUDA_CHECK(cudaMemsetAsync(s->pes_count, 0, sizeof(uint32_t), st));
	const int block = 256;
	const int grid = (int)((n + block - 1) / block);
	ts_bind_kernel<<<grid, block, 0, st>>>(s->ts, (uint32_t)n, stream_id, s->first_pusi, s->error);
	ts_measure_kernel<<<grid, block, 0, st>>>(s->ts, (uint32_t)n, stream_id, s->first_pusi, s->use, s->lens, s->is_pes, s->pts);
	CUDA_CHECK(cudaMemsetAsync(s->lens + n, 0, sizeof(uint32_t), st));
	size_t bytes = s->cub_bytes;
	cub::DeviceScan::ExclusiveSum(s->cub_temp, bytes, s->lens, s->offsets, (int)(n + 1), st);
	uint32_t total = 0;
	int error = 0;
	CUDA_CHECK(cudaMemcpyAsync(&total, s->offsets + n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CUDA_CHECK(cudaMemcpyAsync(&error, s->error, sizeof(int), cudaMemcpyDeviceToHost, st));
	CUDA_CHECK(cudaMemcpyAsync(host_first, s->first_pusi, sizeof(host_first), cudaMemcpyDeviceToHost, st));
	CUDA_CHECK(cudaStreamSynchronize(st));
	if (error) return -1;
	for (int i = 0; i < 8192; i++) bound[i] = host_first[i] != 0xffffffffu;
	return (long)total;
}

// Phase 2: gather into es[es_base ...]; returns the PES count, lists copied to the host arrays.
int ts_demux_gather(TsScratch *s, size_t ts_bytes, uint8_t *es, uint32_t es_base, uint64_t *pts_out,
                    uint32_t *offset_out, int n_max, cudaStream_t st) {
	const size_t n = ts_bytes / 188;
	if (n == 0) return 0;
	const int block = 256;
	const int grid = (int)((n * 32 + block - 1) / block);
	ts_gather_kernel<<<grid, block, 0, st>>>(s->ts, (uint32_t)n, s->use, s->offsets, s->is_pes, s->pts, es, es_base,
	                                         s->pes_count, (uint32_t)s->n_cap, s->pes_offset, s->pes_pts);
	uint32_t count = 0;
	CUDA_CHECK(cudaMemcpyAsync(&count, s->pes_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
	CUDA_CHECK(cudaStreamSynchronize(st));
	if (pts_out && offset_out && n_max > 0) {
		const uint32_t m = count < (uint32_t)n_max ? count : (uint32_t)n_max;
		CUDA_CHECK(cudaMemcpyAsync(pts_out, s->pes_pts, m * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
		CUDA_CHECK(cudaMemcpyAsync(offset_out, s->pes_offset, m * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
		CUDA_CHECK(cudaStreamSynchronize(st));
	}
	return (int)count;
}
