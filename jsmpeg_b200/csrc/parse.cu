// parse.cu -- stage 1: bitstream / VLC parse, one warp per picture (sm_100a).
//
// Replaces, for the bitstream half, the reference's decodePicture / decodeSlice / decodeMacroblock /
// decodeMotionVectors / decodeBlock (src/mpeg1.js:174-457, 698-811) and its bit reader
// (src/buffer.js:115-187).  A picture is an inherently serial VLC walk (DC and motion-vector
// predictors, quantiser scale, run/level positions), so the unit of parallelism is the PICTURE:
// every picture starts at a byte-aligned start code and resets all of that state in its slice
// headers, and nothing in the parse depends on decoded pixels.  One warp owns one picture:
//   * the 32 lanes run the walk in lock-step (warp-uniform control flow, no divergence),
//   * the bitstream is fetched 128 B at a time, one word per lane, double buffered, and handed to
//     the 64-bit bit window with a shuffle,
//   * the 64 coefficients of the block being decoded live in the warp's registers (two int16 per
//     lane) and leave as ONE coalesced 128-byte store per coded block,
//   * VLCs are decoded with clz-indexed look-up tables in shared memory (tools/gen_tables.py)
//     instead of the reference's one-bit-per-step tree walk (src/mpeg1.js:66-72); the tables are
//     pinned to the reference's trees by tests/test_vlc_tables.py.
// Output: mb_record_t per macroblock address + dequantised int16 coefficient blocks (records.h).
#include "common.cuh"

#define VLC_TABLE_QUALIFIER static __device__ const
#include "vlc_tables.h"

namespace {

constexpr int WARPS_PER_CTA = 4;
constexpr unsigned FULL = 0xffffffffu;

struct Luts {
	uint16_t dct[(VLC_DCT_MAX_Z + 1) * 32];
	uint16_t mba[(VLC_MBA_MAX_Z + 1) * 32];
	uint16_t cbp[(VLC_CBP_MAX_Z + 1) * 32];
	uint16_t motion[(VLC_MOTION_MAX_Z + 1) * 32];
	uint16_t dc_luma[128];
	uint16_t dc_chroma[256];
	uint16_t type_i[4];
	uint16_t type_p[64];
	uint8_t zigzag[64];
};

struct WarpShared {
	uint8_t intra_q[64];
	uint8_t non_intra_q[64];
};

// MSB-first bit window over a byte span (src/buffer.js:152-187), warp-uniform.
struct BitReader {
	const uint32_t *words;
	const uint8_t *bytes;
	uint32_t len;         // valid bytes; everything past it reads as zero
	uint32_t wpos;        // next word to shift into the window
	uint32_t chunk_base;  // word index held by lane 0 of `chunk`
	uint32_t chunk, chunk_next;
	uint64_t win;         // left-aligned window
	int nbits;            // valid bits in win (>= 32 between calls)
	int lane;

	__device__ __forceinline__ uint32_t load_word(uint32_t w) const {
		uint32_t byte = w * 4u;
		if (byte >= len) return 0u;
		uint32_t v = __byte_perm(__ldg(words + w), 0, 0x0123);  // first byte -> MSB
		uint32_t left = len - byte;
		if (left < 4u) v &= 0xffffffffu << (8u * (4u - left));
		return v;
	}
	__device__ __forceinline__ uint32_t fetch(uint32_t w) {
		if (w >= chunk_base + 32u) {  // warp-uniform
			chunk = chunk_next;
			chunk_base += 32u;
			chunk_next = load_word(chunk_base + 32u + lane);
		}
		return __shfl_sync(FULL, chunk, (int)(w - chunk_base));
	}
	__device__ __forceinline__ void seek_byte(uint32_t byte_pos) {
		wpos = byte_pos >> 2;
		chunk_base = wpos & ~31u;
		chunk = load_word(chunk_base + lane);
		chunk_next = load_word(chunk_base + 32u + lane);
		uint32_t hi = fetch(wpos++);
		uint32_t lo = fetch(wpos++);
		win = ((uint64_t)hi << 32) | lo;
		nbits = 64;
		int drop = (int)(byte_pos & 3u) * 8;
		if (drop) consume(drop);
	}
	__device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(win >> 32); }
	__device__ __forceinline__ void consume(int n) {
		win <<= n;
		nbits -= n;
		if (nbits < 32) {
			uint32_t w = fetch(wpos++);
			win |= (uint64_t)w << (32 - nbits);
			nbits += 32;
		}
	}
	__device__ __forceinline__ uint32_t read(int n) {  // 1 <= n <= 32
		uint32_t v = peek32() >> (32 - n);
		consume(n);
		return v;
	}
	__device__ __forceinline__ uint32_t bitpos() const { return wpos * 32u - (uint32_t)nbits; }

	// src/buffer.js:141-150 nextBytesAreStartCode
	__device__ __forceinline__ bool next_bytes_are_start_code() const {
		uint32_t bp = bitpos();
		uint32_t i = (bp + 7u) >> 3;
		if (i >= len) return true;
		int skip = (int)((8u - (bp & 7u)) & 7u);
		// bytes past `len` are zero in the window, so a code straddling the end cannot match
		return (uint32_t)((win << skip) >> 40) == 0x000001u && i + 2u < len;
	}
	// src/buffer.js:115-128 findNextStartCode: lanes test 32 byte positions per step.
	// Returns the code (and leaves the reader after it) or -1 (reader parked at the end).
	__device__ int find_next_start_code() {
		uint32_t i = (bitpos() + 7u) >> 3;
		while (i + 3u < len) {
			uint32_t j = i + (uint32_t)lane;
			bool hit = false;
			if (j + 3u < len) hit = bytes[j] == 0 && bytes[j + 1] == 0 && bytes[j + 2] == 1;
			unsigned m = __ballot_sync(FULL, hit);
			if (m) {
				uint32_t at = i + (uint32_t)(__ffs(m) - 1);
				int code = bytes[at + 3];
				seek_byte(at + 4u);
				return code;
			}
			i += 32u;
		}
		seek_byte(len);
		return -1;
	}
};

struct PictureState {
	int picture_type, full_pel, r_size, f;
	int qscale, mb_addr;
	bool slice_begin;
	int mv_h, mv_v, mv_h_prev, mv_v_prev;
	int dc_y, dc_b4, dc_b5;  // block 4 / block 5 predictors (the reference's "Cr"/"Cb", mpeg1.js:717)
	int n_present, n_coded, error;
};

__device__ __forceinline__ int clz_lut(const uint16_t *lut, uint32_t w, int max_z) {
	int z = __clz((int)w);
	if (z > max_z) return 0;
	return lut[(z << 5) | ((w << (z + 1)) >> 27)];
}

// One coded block: src/mpeg1.js:698-811.  Returns false on an invalid code.
__device__ __forceinline__ bool parse_block(BitReader &br, const Luts &L, const uint8_t *quant,
                                            PictureState &ps, bool intra, int block,
                                            uint32_t *coef_out /* 32 words */, bool &dc_only) {
	const int lane = br.lane;
	uint32_t acc = 0;  // coefficients 2*lane (low half) and 2*lane+1 (high half)
	int n = 0;
	if (intra) {
		// DC size VLC + differential + predictor (mpeg1.js:705-751)
		uint32_t w = br.peek32();
		int e = block < 4 ? L.dc_luma[w >> 25] : L.dc_chroma[w >> 24];
		int len = e & 31, size = e >> 5;
		if (len == 0) return false;
		br.consume(len);
		int *pred = block < 4 ? &ps.dc_y : (block == 4 ? &ps.dc_b4 : &ps.dc_b5);
		int dc = *pred;
		if (size > 0) {
			int diff = (int)br.read(size);
			dc += (diff & (1 << (size - 1))) ? diff : (int)((0xffffffffu << size) | (uint32_t)(diff + 1));
		}
		*pred = dc;
		int v = max(-32768, min(32767, dc * 8));
		if (lane == 0) acc = (uint32_t)v & 0xffffu;
		n = 1;
	}
	const int qs = ps.qscale;
	for (;;) {  // mpeg1.js:757-811
		uint32_t w = br.peek32();
		int run, level;
		if (w >> 31) {
			if (n == 0) {               // dct_coeff_first: '1s' = (0, +-1)
				run = 0;
				level = (w & 0x40000000u) ? -1 : 1;
				br.consume(2);
			} else if (!(w & 0x40000000u)) {  // '10' end_of_block (mpeg1.js:763)
				br.consume(2);
				break;
			} else {                    // '11s'
				run = 0;
				level = (w & 0x20000000u) ? -1 : 1;
				br.consume(3);
			}
		} else {
			int e = clz_lut(L.dct, w, VLC_DCT_MAX_Z);
			int len = e & 31;
			if (len == 0) return false;
			int payload = e >> 5;
			if (payload == 0) {         // escape (mpeg1.js:767-780): 6 + 6 + 8 (+ 8) bits
				run = (w >> 20) & 63;
				int l8 = (w >> 12) & 255;
				if (l8 == 0) { level = (w >> 4) & 255; br.consume(28); }
				else if (l8 == 128) { level = (int)((w >> 4) & 255) - 256; br.consume(28); }
				else { level = l8 > 128 ? l8 - 256 : l8; br.consume(20); }
			} else {
				run = payload & 31;
				level = payload >> 5;
				if ((w >> (31 - len)) & 1u) level = -level;
				br.consume(len + 1);
			}
		}
		n += run;
		if (n > 63) {  // JS: ZIG_ZAG[n] undefined -> the store is a no-op
			ps.error = PARSE_ERR_COEF_INDEX;
			n++;
			continue;
		}
		int idx = L.zigzag[n];
		n++;
		// dequantise, oddify toward zero, clip (mpeg1.js:794-807)
		level <<= 1;
		if (!intra) level += level < 0 ? -1 : 1;
		level = (level * qs * (int)quant[idx]) >> 4;
		if ((level & 1) == 0) level -= level > 0 ? 1 : -1;
		level = max(-2048, min(2047, level));
		if (lane == (idx >> 1)) acc |= ((uint32_t)level & 0xffffu) << ((idx & 1) * 16);
	}
	dc_only = (n == 1);  // mpeg1.js:838, 850
	coef_out[lane] = acc;  // one coalesced 128-byte store
	ps.n_coded++;
	return true;
}

// mpeg1.js:395-457, one component
__device__ __forceinline__ bool parse_motion(BitReader &br, const Luts &L, const PictureState &ps,
                                             int &prev, int &mv) {
	int e = clz_lut(L.motion, br.peek32(), VLC_MOTION_MAX_Z);
	int len = e & 31;
	if (len == 0) return false;
	br.consume(len);
	int code = (e >> 5) - 16;
	int d = code;
	if (code != 0 && ps.f != 1) {
		int r = (int)br.read(ps.r_size);
		d = ((abs(code) - 1) << ps.r_size) + r + 1;
		if (code < 0) d = -d;
	}
	prev += d;
	if (prev > (ps.f << 4) - 1) prev -= ps.f << 5;
	else if (prev < -(ps.f << 4)) prev += ps.f << 5;
	mv = ps.full_pel ? prev * 2 : prev;
	return true;
}

__device__ __forceinline__ int read_mba(BitReader &br, const Luts &L) {
	int e = clz_lut(L.mba, br.peek32(), VLC_MBA_MAX_Z);
	int len = e & 31;
	if (len == 0) return -1;
	br.consume(len);
	return e >> 5;
}

__device__ __forceinline__ uint4 pack_record(int mv_h, int mv_v, int flags, int cbp, int dc_only,
                                             int qscale, uint32_t bit_pos) {
	uint4 r;
	r.x = ((uint32_t)mv_h & 0xffffu) | ((uint32_t)mv_v << 16);
	r.y = (uint32_t)flags | ((uint32_t)cbp << 8) | ((uint32_t)dc_only << 16) | ((uint32_t)qscale << 24);
	r.z = bit_pos;
	r.w = 0;
	return r;
}

// mpeg1.js:294-392 decodeMacroblock.  false = stop walking this slice.
__device__ bool parse_macroblock(BitReader &br, const Luts &L, const WarpShared &ws,
                                 PictureState &ps, const ParseTask &t, int mb_size) {
	const int lane = br.lane;
	int increment = 0;
	int v = read_mba(br, L);
	while (v == 34) v = read_mba(br, L);                       // macroblock_stuffing
	while (v == 35) { increment += 33; v = read_mba(br, L); }  // macroblock_escape
	if (v < 0) return false;
	increment += v;

	if (ps.slice_begin) {  // mpeg1.js:312-317
		ps.slice_begin = false;
		ps.mb_addr += increment;
	} else {
		if (ps.mb_addr + increment >= mb_size) return true;  // mpeg1.js:319-322
		if (increment > 1) {  // mpeg1.js:323-334
			ps.dc_y = ps.dc_b4 = ps.dc_b5 = 128;
			if (ps.picture_type == 2) ps.mv_h = ps.mv_v = ps.mv_h_prev = ps.mv_v_prev = 0;
			// skipped macroblocks: predicted copy with the current vector (mpeg1.js:336-346)
			int n_skip = increment - 1;
			uint4 rec = pack_record(ps.mv_h, ps.mv_v, MBF_PRESENT | MBF_SKIPPED, 0, 0, ps.qscale, br.bitpos());
			for (int k = lane; k < n_skip; k += 32)
				reinterpret_cast<uint4 *>(t.hdr)[ps.mb_addr + 1 + k] = rec;
			ps.n_present += n_skip;
			ps.mb_addr += n_skip;
		}
		ps.mb_addr++;
	}
	const int mb = ps.mb_addr;
	if (mb < 0 || mb >= mb_size) return false;  // outside the picture: never write there

	uint32_t w = br.peek32();
	int e = ps.picture_type == 1 ? L.type_i[w >> 30] : L.type_p[w >> 26];
	if ((e & 31) == 0) return false;
	br.consume(e & 31);
	const int type = e >> 5;
	const bool intra = type & 0x01;
	if (type & 0x10) ps.qscale = (int)br.read(5);
	const uint32_t mb_bit_pos = br.bitpos();

	if (intra) {
		ps.mv_h = ps.mv_v = ps.mv_h_prev = ps.mv_v_prev = 0;  // mpeg1.js:363-367
	} else {
		ps.dc_y = ps.dc_b4 = ps.dc_b5 = 128;                  // mpeg1.js:370-372
		if (type & 0x08) {
			if (!parse_motion(br, L, ps, ps.mv_h_prev, ps.mv_h)) return false;
			if (!parse_motion(br, L, ps, ps.mv_v_prev, ps.mv_v)) return false;
		} else if (ps.picture_type == 2) {
			ps.mv_h = ps.mv_v = ps.mv_h_prev = ps.mv_v_prev = 0;  // mpeg1.js:452-456
		}
	}

	int cbp = intra ? 0x3f : 0;
	if (type & 0x02) {
		int c = clz_lut(L.cbp, br.peek32(), VLC_CBP_MAX_Z);
		if ((c & 31) == 0) return false;
		br.consume(c & 31);
		cbp = c >> 5;
	}

	const int mv_h = ps.mv_h, mv_v = ps.mv_v, qscale = ps.qscale;
	const uint8_t *quant = intra ? ws.intra_q : ws.non_intra_q;
	uint32_t *coef_mb = reinterpret_cast<uint32_t *>(t.coef) + (size_t)mb * (MB_COEF_INT16 / 2);
	int done = 0, dc_mask = 0;
	bool ok = true;
#pragma unroll 1
	for (int block = 0; block < 6; block++) {
		if (cbp & (0x20 >> block)) {
			bool dc_only;
			ok = parse_block(br, L, quant, ps, intra, block, coef_mb + block * 32, dc_only);
			if (!ok) break;
			done |= 0x20 >> block;
			if (dc_only) dc_mask |= 0x20 >> block;
		}
	}
	if (lane == 0)
		reinterpret_cast<uint4 *>(t.hdr)[mb] =
		    pack_record(mv_h, mv_v, MBF_PRESENT | (intra ? MBF_INTRA : 0), done, dc_mask, qscale, mb_bit_pos);
	ps.n_present++;
	return ok;
}

__global__ void __launch_bounds__(WARPS_PER_CTA * 32)
parse_pictures_kernel(const ParseTask *__restrict__ tasks, int n_tasks) {
	__shared__ Luts L;
	__shared__ WarpShared wsh[WARPS_PER_CTA];

	for (int i = threadIdx.x; i < (int)(sizeof(L.dct) / 2); i += blockDim.x) L.dct[i] = VLC_DCT_COEFF[i];
	for (int i = threadIdx.x; i < (int)(sizeof(L.mba) / 2); i += blockDim.x) L.mba[i] = VLC_MBA[i];
	for (int i = threadIdx.x; i < (int)(sizeof(L.cbp) / 2); i += blockDim.x) L.cbp[i] = VLC_CBP[i];
	for (int i = threadIdx.x; i < (int)(sizeof(L.motion) / 2); i += blockDim.x) L.motion[i] = VLC_MOTION[i];
	for (int i = threadIdx.x; i < 128; i += blockDim.x) L.dc_luma[i] = VLC_DC_SIZE_LUMA[i];
	for (int i = threadIdx.x; i < 256; i += blockDim.x) L.dc_chroma[i] = VLC_DC_SIZE_CHROMA[i];
	for (int i = threadIdx.x; i < 4; i += blockDim.x) L.type_i[i] = VLC_MBTYPE_I[i];
	for (int i = threadIdx.x; i < 64; i += blockDim.x) L.type_p[i] = VLC_MBTYPE_P[i];
	for (int i = threadIdx.x; i < 64; i += blockDim.x) L.zigzag[i] = TBL_ZIG_ZAG[i];
	__syncthreads();

	const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
	const int task_id = blockIdx.x * WARPS_PER_CTA + warp;
	if (task_id >= n_tasks) return;
	const ParseTask t = tasks[task_id];
	WarpShared &ws = wsh[warp];
	for (int i = lane; i < 64; i += 32) {
		ws.intra_q[i] = t.seq->intra_q[i];
		ws.non_intra_q[i] = t.seq->non_intra_q[i];
	}
	__syncwarp();
	const int mb_width = t.seq->mb_width, mb_size = t.seq->mb_size;

	// no macroblock is present until the walk reaches it (an address no slice covers keeps the
	// two-pictures-old samples)
	for (int i = lane; i < mb_size; i += 32) reinterpret_cast<uint4 *>(t.hdr)[i] = make_uint4(0, 0, 0, 0);
	__syncwarp();

	BitReader br;
	br.words = reinterpret_cast<const uint32_t *>(t.es);
	br.bytes = t.es;
	br.len = t.es_len;
	br.lane = lane;
	br.seek_byte(t.start_byte);

	PictureState ps;
	ps.n_present = ps.n_coded = ps.error = 0;
	ps.full_pel = 0; ps.r_size = 0; ps.f = 1;
	int f_code = 0;
	int status = PIC_IGNORED;

	// picture header (mpeg1.js:174-196)
	br.consume(10);
	ps.picture_type = (int)br.read(3);
	br.consume(16);
	bool go = ps.picture_type == 1 || ps.picture_type == 2;
	if (ps.picture_type == 2) {
		ps.full_pel = (int)br.read(1);
		f_code = (int)br.read(3);
		if (f_code == 0) go = false;
		else { ps.r_size = f_code - 1; ps.f = 1 << ps.r_size; }
	}
	uint32_t end_bit;
	if (!go) {
		end_bit = br.bitpos();
	} else {
		status = PIC_DECODED;
		int code;
		do { code = br.find_next_start_code(); } while (code == 0xB5 || code == 0xB2);  // mpeg1.js:198-201
		while (code >= 0x01 && code <= 0xAF) {
			// slice (mpeg1.js:255-276)
			ps.slice_begin = true;
			ps.mb_addr = (code - 1) * mb_width - 1;
			ps.mv_h = ps.mv_v = ps.mv_h_prev = ps.mv_v_prev = 0;
			ps.dc_y = ps.dc_b4 = ps.dc_b5 = 128;
			ps.qscale = (int)br.read(5);
			while (br.read(1)) br.consume(8);
			do {
				if (!parse_macroblock(br, L, ws, ps, t, mb_size)) {
					if (!ps.error) ps.error = PARSE_ERR_INVALID_VLC;
					break;
				}
			} while (!br.next_bytes_are_start_code());
			code = br.find_next_start_code();
		}
		end_bit = br.bitpos();
		if (code != -1) end_bit -= 32;  // mpeg1.js:209-213
	}
	if (lane == 0) {
		picture_info_t info;
		info.start_byte = t.start_byte;
		info.end_bit = end_bit;
		info.status = status;
		info.picture_type = ps.picture_type;
		info.full_pel = ps.full_pel;
		info.f_code = f_code;
		info.n_present = ps.n_present;
		info.n_coded_blocks = ps.n_coded;
		info.error = ps.error;
		info.reserved[0] = info.reserved[1] = info.reserved[2] = 0;
		*t.info = info;
	}
}

}  // namespace

void launch_parse_pictures(const ParseTask *tasks, int n_tasks, cudaStream_t stream) {
	if (n_tasks <= 0) return;
	int grid = (n_tasks + WARPS_PER_CTA - 1) / WARPS_PER_CTA;
	parse_pictures_kernel<<<grid, WARPS_PER_CTA * 32, 0, stream>>>(tasks, n_tasks);
}
