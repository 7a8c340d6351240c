// scan.cu -- stage 0: byte-parallel search for start-code prefixes (00 00 01) in an elementary
// stream resident in HBM.
//
// The reference finds the next picture with a serial byte scan on every decode()
// (src/buffer.js:115-139 findNextStartCode / findStartCode), and ends a slice where the next bytes
// are a start code (buffer.js:141-150).  Here the scan runs once per written span: each thread tests
// 16 byte positions (one 16-byte load plus a look-ahead word) and appends hits to the stream's
// position list; the host sorts the (short) list, picks the picture start codes (fourth byte 00) out
// of it and hands the whole list back to the device: knowing every picture start up front is what
// lets stage 1 parse all buffered pictures concurrently, and knowing every prefix is what tells the
// lane-parallel walk where a slice ends without reading the slice (round 1 and early round 2: a
// 128-bytes-per-step search over every picture, 6 % of the walk's stall samples and 0.8 GB per step).
#include "common.cuh"

namespace {

__global__ void scan_start_codes_kernel(const uint8_t *__restrict__ es, uint32_t from, uint32_t len,
                                        uint32_t *__restrict__ positions, uint32_t capacity,
                                        uint32_t *__restrict__ count) {
	// `es` is 16-byte aligned; thread handles byte positions [base, base + 16)
	const uint32_t first = from & ~15u;
	const uint32_t base = first + (blockIdx.x * blockDim.x + threadIdx.x) * 16u;
	if (base >= len) return;
	const uint4 v = __ldg(reinterpret_cast<const uint4 *>(es + base));
	uint32_t w[5] = {v.x, v.y, v.z, v.w, 0u};
	if (base + 16u < len) w[4] = __ldg(reinterpret_cast<const uint32_t *>(es + base + 16u));
#pragma unroll
	for (int i = 0; i < 16; i++) {
		// four bytes starting at position base + i, byte 0 in the low bits
		const uint32_t q = __funnelshift_r(w[i >> 2], w[(i >> 2) + 1], (i & 3) * 8);
		const uint32_t pos = base + i;
		// 00 00 01 -> low three bytes 0x010000; all three bytes must be inside the buffer (the fourth, the code,
		// is looked at by the host once it is there)
		if ((q & 0x00ffffffu) == 0x00010000u && pos >= from && pos + 2u < len) {
			const uint32_t slot = atomicAdd(count, 1u);
			if (slot < capacity) positions[slot] = pos;
		}
	}
}

}  // namespace

void launch_scan_start_codes(const uint8_t *es, uint32_t from, uint32_t len, uint32_t *positions,
                             uint32_t capacity, uint32_t *count, cudaStream_t stream) {
	if (from >= len) return;
	const uint32_t first = from & ~15u;
	const uint32_t n_threads = (len - first + 15u) / 16u;
	const int block = 256;
	const uint32_t grid = (n_threads + block - 1) / block;
	scan_start_codes_kernel<<<grid, block, 0, stream>>>(es, from, len, positions, capacity, count);
}
