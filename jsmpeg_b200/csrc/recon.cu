// recon.cu -- stage 2: integer IDCT + half-pel motion compensation + add/clamp -> planar Y/Cr/Cb.
//
// Replaces the pixel half of the reference's decodeMacroblock/decodeBlock: copyMacroblock
// (src/mpeg1.js:459-687), IDCT (:916-983) and Copy/Add{Block,Value}ToDestination (:864-914).
// Every macroblock of a picture reads only the PREVIOUS picture's planes and writes only its own
// 16x16 / 8x8 / 8x8 pixels, so all macroblocks of a picture (and of all streams) are independent.
//
// Mapping: a CTA reconstructs MBS_PER_CTA = 4 consecutive macroblocks = 24 blocks of 8x8 with
// 192 threads; thread (blk, k) owns line k of block blk:
//   A  one 16-byte load = row k of the block's 64 int16 coefficients (coalesced 128 B per block),
//      x PREMULTIPLIER (src/mpeg1.js:810, 1026-1035) in int32 -> shared memory
//   B  column k of the block: 8-point pass without final shift (mpeg1.js:925-947)
//   C  row k: 8-point pass with (v+128)>>8 (mpeg1.js:952-981) -> 8 residuals in registers
//   D  row k of the prediction: unaligned 9(+9) byte fetch from the forward plane as aligned
//      32-bit words, packed-byte half-pel averaging, + residual, saturate, one 8-byte store.
// Blocks that take the reference's DC-only shortcut (mpeg1.js:838-841, 850-853) skip B and C.
// The block tile in shared memory is padded to 72 words and the two 4-word halves of rows 4..7
// are swapped, so that A/C (128-bit row accesses) and B (stride-8 column accesses) are all
// bank-conflict free.
//
// HBM roofline accounting: per macroblock 16 B header + 128 B per coded block +
// 384 B written + 384 B of forward-plane samples (P pictures), each counted once.
#include "common.cuh"

#define VLC_TABLE_QUALIFIER static __device__ const
#include "vlc_tables.h"

namespace {

constexpr int MBS_PER_CTA = 4;
constexpr int BLOCKS_PER_CTA = MBS_PER_CTA * 6;
constexpr int THREADS = BLOCKS_PER_CTA * 8;  // 192
constexpr int TILE = 72;                     // padded words per 8x8 block

// One 8-point pass of the reference IDCT.  Column pass: no scaling; row pass: (v + 128) >> 8.
template <bool ROW>
__device__ __forceinline__ void idct8(int (&v)[8]) {
	const int b1 = v[4];
	const int b3 = v[2] + v[6];
	const int b4 = v[5] - v[3];
	const int t1 = v[1] + v[7];
	const int t2 = v[3] + v[5];
	const int b6 = v[1] - v[7];
	const int b7 = t1 + t2;
	const int m0 = v[0];
	const int x4 = ((b6 * 473 - b4 * 196 + 128) >> 8) - b7;
	const int x0 = x4 - (((t1 - t2) * 362 + 128) >> 8);
	const int x1 = m0 - b1;
	const int x2 = (((v[2] - v[6]) * 362 + 128) >> 8) - b3;
	const int x3 = m0 + b1;
	const int y3 = x1 + x2, y4 = x3 + b3, y5 = x1 - x2, y6 = x3 - b3;
	const int y7 = -x0 - ((b4 * 473 + b6 * 196 + 128) >> 8);
	if (ROW) {
		v[0] = (b7 + y4 + 128) >> 8; v[1] = (x4 + y3 + 128) >> 8;
		v[2] = (y5 - x0 + 128) >> 8; v[3] = (y6 - y7 + 128) >> 8;
		v[4] = (y6 + y7 + 128) >> 8; v[5] = (x0 + y5 + 128) >> 8;
		v[6] = (y3 - x4 + 128) >> 8; v[7] = (y4 - b7 + 128) >> 8;
	} else {
		v[0] = b7 + y4; v[1] = x4 + y3; v[2] = y5 - x0; v[3] = y6 - y7;
		v[4] = y6 + y7; v[5] = x0 + y5; v[6] = y3 - x4; v[7] = y4 - b7;
	}
}

// 12 bytes starting at flat index i of a plane, as three packed little-endian words whose byte 0
// is sample i.  `p` is 4-byte aligned at index 0.  Caller guarantees [i, i+12) is readable.
__device__ __forceinline__ void load12(const uint8_t *__restrict__ p, int i, uint32_t &a, uint32_t &b, uint32_t &c) {
	const uint32_t *w = reinterpret_cast<const uint32_t *>(p) + (i >> 2);
	const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3);
	const uint32_t sh = (uint32_t)(i & 3) * 8u;
	a = __funnelshift_r(w0, w1, sh);
	b = __funnelshift_r(w1, w2, sh);
	c = __funnelshift_r(w2, w3, sh);
}

// (a + b + c + d + 2) >> 2 per byte, exact (mpeg1.js:481-500)
__device__ __forceinline__ uint32_t avg4_u8x4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
	const uint32_t M = 0x00ff00ffu;
	uint32_t lo = (a & M) + (b & M) + (c & M) + (d & M) + 0x00020002u;
	uint32_t hi = ((a >> 8) & M) + ((b >> 8) & M) + ((c >> 8) & M) + ((d >> 8) & M) + 0x00020002u;
	return ((lo >> 2) & M) | (((hi >> 2) & M) << 8);
}

__device__ __forceinline__ uint32_t pack_sat_u8x4(int a, int b, int c, int d) {
	// PTX: d[7:0] = sat(b_op), d[15:8] = sat(a_op), d[31:16] = c_op[15:0]
	uint32_t hi, r;
	const uint32_t zero = 0;
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(d), "r"(c), "r"(zero));
	asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(b), "r"(a), "r"(hi));
	return r;
}

__global__ void __launch_bounds__(THREADS)
reconstruct_kernel(const ReconTask *__restrict__ tasks) {
	__shared__ int tile[BLOCKS_PER_CTA * TILE];
	__shared__ int premult[64];

	if (threadIdx.x < 64) premult[threadIdx.x] = TBL_PREMULTIPLIER[threadIdx.x];

	const ReconTask &t = tasks[blockIdx.y];
	const int blk = threadIdx.x >> 3;   // 0..23
	const int k = threadIdx.x & 7;      // line within the block
	const int mb = blockIdx.x * MBS_PER_CTA + blk / 6;
	const int b = blk % 6;              // block within the macroblock
	const bool mb_valid = mb < t.mb_size;

	uint32_t rec_y = 0;
	int mv_h = 0, mv_v = 0;
	if (mb_valid) {
		const uint2 r = __ldg(reinterpret_cast<const uint2 *>(t.hdr + mb));
		mv_h = (int)(int16_t)(r.x & 0xffffu);
		mv_v = (int)(int16_t)(r.x >> 16);
		rec_y = r.y;
	}
	const int flags = rec_y & 0xff;
	const int bit = 0x20 >> b;
	const bool present = flags & MBF_PRESENT;
	const bool intra = flags & MBF_INTRA;
	const bool coded = present && (((rec_y >> 8) & 0xff) & bit);
	const bool dc_only = coded && (((rec_y >> 16) & 0xff) & bit);
	const bool full = coded && !dc_only;

	// ---- A: coefficients -> premultiplied int32 tile
	const int16_t *cblk = t.coef + ((size_t)mb * 6 + b) * 64;
	int res[8];
	int *my = tile + blk * TILE;
	__syncthreads();  // premult visible
	if (full) {
		const uint4 q = __ldg(reinterpret_cast<const uint4 *>(cblk) + k);
		const int *pm = premult + k * 8;
		int4 lo, hi;
		lo.x = (int)(int16_t)(q.x & 0xffffu) * pm[0]; lo.y = ((int)q.x >> 16) * pm[1];
		lo.z = (int)(int16_t)(q.y & 0xffffu) * pm[2]; lo.w = ((int)q.y >> 16) * pm[3];
		hi.x = (int)(int16_t)(q.z & 0xffffu) * pm[4]; hi.y = ((int)q.z >> 16) * pm[5];
		hi.z = (int)(int16_t)(q.w & 0xffffu) * pm[6]; hi.w = ((int)q.w >> 16) * pm[7];
		*reinterpret_cast<int4 *>(my + k * 8 + (k & 4)) = lo;
		*reinterpret_cast<int4 *>(my + k * 8 + ((k & 4) ^ 4)) = hi;
	} else if (dc_only) {
		const int c0 = (int)__ldg(cblk);
		const int v = (c0 * premult[0] + 128) >> 8;  // mpeg1.js:838-841, 850-853
#pragma unroll
		for (int j = 0; j < 8; j++) res[j] = v;
	} else {
#pragma unroll
		for (int j = 0; j < 8; j++) res[j] = 0;
	}
	__syncthreads();

	// ---- B: column pass
	if (full) {
		int v[8];
#pragma unroll
		for (int j = 0; j < 8; j++) v[j] = my[j * 8 + (j < 4 ? k : k ^ 4)];
		idct8<false>(v);
#pragma unroll
		for (int j = 0; j < 8; j++) my[j * 8 + (j < 4 ? k : k ^ 4)] = v[j];
	}
	__syncthreads();

	// ---- C: row pass
	if (full) {
		const int4 lo = *reinterpret_cast<const int4 *>(my + k * 8 + (k & 4));
		const int4 hi = *reinterpret_cast<const int4 *>(my + k * 8 + ((k & 4) ^ 4));
		res[0] = lo.x; res[1] = lo.y; res[2] = lo.z; res[3] = lo.w;
		res[4] = hi.x; res[5] = hi.y; res[6] = hi.z; res[7] = hi.w;
		idct8<true>(res);
	}
	if (!present) return;  // untouched macroblock keeps the two-pictures-old content

	// ---- D: prediction + residual -> 8 output samples of line k
	uint8_t *dplane;
	const uint8_t *splane;
	int stride, plane_size, origin, mh, mv;
	const int mb_row = mb / t.mb_width, mb_col = mb - mb_row * t.mb_width;
	if (b < 4) {
		dplane = t.cur.y; splane = t.fwd.y;
		stride = t.coded_width;
		plane_size = t.coded_width * t.coded_height;
		origin = (mb_row * 16 + (b >> 1) * 8 + k) * stride + mb_col * 16 + (b & 1) * 8;
		mh = mv_h; mv = mv_v;
	} else {
		// block 4 -> Cb plane, block 5 -> Cr plane (mpeg1.js:829-834)
		dplane = b == 4 ? t.cur.cb : t.cur.cr;
		splane = b == 4 ? t.fwd.cb : t.fwd.cr;
		stride = t.coded_width >> 1;
		plane_size = (t.coded_width * t.coded_height) >> 2;
		origin = (mb_row * 8 + k) * stride + mb_col * 8;
		mh = mv_h / 2; mv = mv_v / 2;  // truncation toward zero (mpeg1.js:562-565)
	}

	uint32_t p0 = 0, p1 = 0;  // predicted samples 0..3, 4..7
	if (!intra) {
		const int oh = mh & 1, ov = mv & 1;
		const int src = origin + (mv >> 1) * stride + (mh >> 1);  // flat index (mpeg1.js:479, 567)
		const int last = src + 8 + stride;                        // furthest tap that may be used
		if (src >= 0 && last + 16 < plane_size) {
			uint32_t a0, a1, a2;
			load12(splane, src, a0, a1, a2);
			if (!ov) {
				if (!oh) { p0 = a0; p1 = a1; }
				else {
					p0 = __vavgu4(a0, __funnelshift_r(a0, a1, 8));
					p1 = __vavgu4(a1, __funnelshift_r(a1, a2, 8));
				}
			} else {
				uint32_t c0, c1, c2;
				load12(splane, src + stride, c0, c1, c2);
				if (!oh) { p0 = __vavgu4(a0, c0); p1 = __vavgu4(a1, c1); }
				else {
					p0 = avg4_u8x4(a0, __funnelshift_r(a0, a1, 8), c0, __funnelshift_r(c0, c1, 8));
					p1 = avg4_u8x4(a1, __funnelshift_r(a1, a2, 8), c1, __funnelshift_r(c1, c2, 8));
				}
			}
		} else {
			// vector leaves the plane: per-tap bounds check, any outside tap zeroes the sample
			for (int x = 0; x < 8; x++) {
				const int i = src + x;
				const int taps[4] = {i, i + 1, i + stride, i + stride + 1};
				const bool use[4] = {true, (bool)oh, (bool)ov, oh && ov};
				int sum = 0, n = 0;
				bool inside = true;
				for (int q = 0; q < 4; q++) {
					if (!use[q]) continue;
					if (taps[q] < 0 || taps[q] >= plane_size) { inside = false; continue; }
					sum += splane[taps[q]];
					n++;
				}
				const int v = !inside ? 0 : (n == 4 ? (sum + 2) >> 2 : (n == 2 ? (sum + 1) >> 1 : sum));
				if (x < 4) p0 |= (uint32_t)v << (8 * x);
				else p1 |= (uint32_t)v << (8 * (x - 4));
			}
		}
	}
	uint2 out;
	if (coded) {
		out.x = pack_sat_u8x4((int)(p0 & 255u) + res[0], (int)((p0 >> 8) & 255u) + res[1],
		                      (int)((p0 >> 16) & 255u) + res[2], (int)(p0 >> 24) + res[3]);
		out.y = pack_sat_u8x4((int)(p1 & 255u) + res[4], (int)((p1 >> 8) & 255u) + res[5],
		                      (int)((p1 >> 16) & 255u) + res[6], (int)(p1 >> 24) + res[7]);
	} else {
		out.x = p0; out.y = p1;
	}
	*reinterpret_cast<uint2 *>(dplane + origin) = out;
}

}  // namespace

void launch_reconstruct(const ReconTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream) {
	if (n_tasks <= 0 || max_mb_size <= 0) return;
	dim3 grid((max_mb_size + MBS_PER_CTA - 1) / MBS_PER_CTA, n_tasks);
	reconstruct_kernel<<<grid, THREADS, 0, stream>>>(tasks);
}
