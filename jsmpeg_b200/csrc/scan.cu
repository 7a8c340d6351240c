// scan.cu -- stage 0: byte-parallel search for picture start codes (00 00 01 00) in an
// elementary stream resident in HBM.
//
// The reference finds the next picture with a serial byte scan on every decode()
// (src/buffer.js:115-139 findNextStartCode / findStartCode).  Here the scan runs once per written
// span: each thread tests 16 byte positions (one 16-byte load plus a 3-byte look-ahead) and
// appends hits to the stream's position list; the host sorts the (short) list.  Knowing every
// picture start up front is what lets stage 1 parse all buffered pictures concurrently.
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
		// 00 00 01 00 -> little-endian word 0x00010000; all four bytes must be inside the buffer
		if (q == 0x00010000u && pos >= from && pos + 3u < len) {
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
