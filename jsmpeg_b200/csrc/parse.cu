// parse.cu -- stage 1: bitstream / VLC parse (sm_100a), in two kernels.
//
// Replaces, for the bitstream half, the reference's decodePicture / decodeSlice / decodeMacroblock /
// decodeMotionVectors / decodeBlock (src/mpeg1.js:174-457, 698-811) and its bit reader
// (src/buffer.js:115-187).
//
// A picture is an inherently serial VLC walk: where symbol k+1 starts is only known once symbol k
// has been decoded.  Everything ELSE the reference does per symbol (run/level bookkeeping,
// zig-zag, dequantisation, oddification, clipping, the store) does not feed that chain.  So:
//
//   1a  walk_pictures_kernel    one warp per PICTURE (every picture starts at a byte-aligned start
//       code and resets all predictor state in its slice headers, and nothing in the parse depends
//       on decoded pixels, so all buffered pictures of all streams are walked concurrently).
//       The walk does the minimum that is serial: macroblock headers (address increment, type,
//       quantiser, motion vectors with their predictors, coded block pattern), intra DC
//       differentials with their predictors, and for the AC coefficients only LUT -> code length
//       -> shift.  It emits the 16-byte macroblock record and, per coded block, the bit offset of
//       its first coefficient code (parked in the block's own 128-byte coefficient slot).
//   1b  expand_blocks_kernel    one thread per coded BLOCK, all blocks of all pictures at once:
//       re-reads the block's codes from its bit offset, does run/level -> zig-zag -> dequantise ->
//       oddify -> clip (src/mpeg1.js:757-811) into a shared-memory tile and writes the 64 x int16
//       block record with 16-byte stores.
//
// VLCs are decoded with clz-indexed look-up tables in shared memory (tools/gen_tables.py; pinned
// to the reference's trees by tests/test_vlc_tables.py) instead of the reference's one-bit-per-step
// tree walk (src/mpeg1.js:66-72).  Bit window: 64-bit, MSB first, refilled one prefetched 32-bit
// word at a time.
// Output: mb_record_t per macroblock address + dequantised int16 coefficient blocks (records.h).
#include <mutex>
#include <vector>

#include "common.cuh"

#define VLC_TABLE_QUALIFIER static __device__ const
#include "vlc_tables.h"

namespace {

constexpr int CTA_THREADS = 128;       // expand kernel
constexpr uint32_t TILE_PITCH = 144;   // bytes between the per-thread 64 x int16 tiles of the expand kernel (128 + 16: bank spread)
#ifndef JSMPEG_MS_BITS
#define JSMPEG_MS_BITS 13
#endif
#ifndef JSMPEG_WALK_THREADS
#define JSMPEG_WALK_THREADS 256
#endif
constexpr int WALK_THREADS = JSMPEG_WALK_THREADS;  // walk kernel: the pictures of a CTA share one multi-symbol table
constexpr int MS_BITS = JSMPEG_MS_BITS;            // multi-symbol table is indexed by the next MS_BITS bits (2 << MS_BITS bytes)
constexpr uint32_t OFF_MS = 4096;      // uint16[1 << MS_BITS], after the per-symbol tables

// shared-memory layout (byte offsets from the dynamic shared base)
constexpr uint32_t OFF_DCT = 0;                                        // uint16[384]
constexpr uint32_t OFF_MBA = OFF_DCT + (VLC_DCT_MAX_Z + 1) * 64;       // uint16[256]
constexpr uint32_t OFF_CBP = OFF_MBA + (VLC_MBA_MAX_Z + 1) * 64;       // uint16[256]
constexpr uint32_t OFF_MOTION = OFF_CBP + (VLC_CBP_MAX_Z + 1) * 64;    // uint16[224]
constexpr uint32_t OFF_DC_LUMA = OFF_MOTION + (VLC_MOTION_MAX_Z + 1) * 64;  // uint16[128]
constexpr uint32_t OFF_DC_CHROMA = OFF_DC_LUMA + 256;                  // uint16[256]
constexpr uint32_t OFF_TYPE_I = OFF_DC_CHROMA + 512;                   // uint16[4]
constexpr uint32_t OFF_TYPE_P = OFF_TYPE_I + 8;                        // uint16[64]
constexpr uint32_t OFF_ZIGZAG = OFF_TYPE_P + 128;                      // uint8[64]
constexpr uint32_t OFF_BLOCKS = (OFF_ZIGZAG + 64 + 127) & ~127u;       // int16[64] per group

__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
	uint16_t v;
	asm("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
	uint32_t v;
	asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
	return v;
}
__device__ __forceinline__ void sts_s16(uint32_t addr, int v) {
	asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
// MSB-first bit window over a byte span (src/buffer.js:152-187); one copy per thread.
struct BitReader {
	const uint32_t *words;
	const uint8_t *bytes;
	uint32_t len;    // valid bytes; everything past it reads as zero (JS typed-array semantics)
	uint32_t wpos;   // index of the word held in `nextw` (the next one to enter the window)
	uint32_t nextw;  // prefetched
	uint64_t win;    // left-aligned window
	int nbits;       // valid bits in win, >= 32 between calls

	// (Measured and rejected on the 3840-picture wave: a branch-free variant relying on the zero pad
	// after the data, 17 % slower; a software prefetch 256 B ahead at every refill, 7 % slower.)
	__device__ __forceinline__ uint32_t load_word(uint32_t w) const {
		const uint32_t byte = w * 4u;
		if (byte >= len) return 0u;
		uint32_t v = __byte_perm(__ldg(words + w), 0, 0x0123);  // first byte -> MSB
		const uint32_t left = len - byte;
		if (left < 4u) v &= 0xffffffffu << (8u * (4u - left));
		return v;
	}
	__device__ __forceinline__ void seek_byte(uint32_t byte_pos) {
		const uint32_t w = byte_pos >> 2;
		win = ((uint64_t)load_word(w) << 32) | load_word(w + 1);
		wpos = w + 2;
		nextw = load_word(wpos);
		nbits = 64;
		const int drop = (int)(byte_pos & 3u) * 8;
		if (drop) consume(drop);
	}
	__device__ __forceinline__ uint32_t peek32() const { return (uint32_t)(win >> 32); }
	__device__ __forceinline__ void consume(int n) {  // 0 <= n <= 32
		win <<= n;
		nbits -= n;
		if (nbits < 32) {
			win |= (uint64_t)nextw << (32 - nbits);
			nbits += 32;
			wpos++;
			nextw = load_word(wpos);
		}
	}
	__device__ __forceinline__ uint32_t read(int n) {  // 1 <= n <= 32
		const uint32_t v = peek32() >> (32 - n);
		consume(n);
		return v;
	}
	__device__ __forceinline__ uint32_t bitpos() const { return wpos * 32u - (uint32_t)nbits; }

	// src/buffer.js:141-150 nextBytesAreStartCode
	__device__ __forceinline__ bool next_bytes_are_start_code() const {
		const uint32_t bp = bitpos();
		const uint32_t i = (bp + 7u) >> 3;
		if (i >= len) return true;
		const int skip = (int)((8u - (bp & 7u)) & 7u);
		// bytes past `len` are zero in the window, so a code straddling the end cannot match
		return (uint32_t)((win << skip) >> 40) == 0x000001u && i + 2u < len;
	}
	// src/buffer.js:115-128 findNextStartCode.  Returns the code (reader left just after it) or -1
	// (reader parked at the end of the data).  A start code needs its 4 bytes inside the buffer.
	__device__ int find_next_start_code() {
		uint32_t i = (bitpos() + 7u) >> 3;
		for (; i + 3u < len; i++) {
			if (bytes[i + 2] > 1) { i += 2; continue; }  // 00 00 01 cannot end at i+2, i+3 or i+4
			if (bytes[i] == 0 && bytes[i + 1] == 0 && bytes[i + 2] == 1) {
				const int code = bytes[i + 3];
				seek_byte(i + 4u);
				return code;
			}
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

// The shared-window address of the dynamic shared memory, made opaque so that the compiler keeps
// it in a register instead of re-deriving it (S2R + LEA) at every table access.
__device__ __forceinline__ uint32_t smem_base(const void *p) {
	uint32_t a = (uint32_t)__cvta_generic_to_shared(p), r;
	asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(a));
	return r;
}

__device__ __forceinline__ uint32_t clz_lut(uint32_t table_addr, uint32_t w, int max_z) {
	const int z = __clz((int)w);
	if (z > max_z) return 0;
	return lds_u16(table_addr + (((uint32_t)z << 5) | ((w << (z + 1)) >> 27)) * 2u);
}

// ==================================================================================================
// 1a: the serial walk

// walk-table entry derived from the DCT table: bits 0..4 = bits to consume (code + sign),
// bits 5..10 = run + 1, bits 11..12 = 1 end_of_block / 2 escape; 0 = invalid code.
__device__ __forceinline__ uint16_t walk_entry(uint16_t e) {
	const int len = e & 31, run = (e >> 5) & 31, level = e >> 10;
	if (len == 0) return 0;
	if (level == 0) return run ? (uint16_t)(2 | (1 << 11)) : (uint16_t)(6 | (2 << 11));
	return (uint16_t)((len + 1) | ((run + 1) << 5));
}

// One coded block (bitstream side of src/mpeg1.js:698-811): intra DC with its predictor, then only
// code lengths.  Leaves {bit offset of the first coefficient code, dc * 8} in the block's slot.
__device__ __forceinline__ bool walk_block(BitReader &br, uint32_t sbase, PictureState &ps, bool intra, int block,
                                           uint32_t *__restrict__ slot, int lane, bool &dc_only) {
	int n = 0;
	int dc8 = 0;
	if (intra) {
		// DC size VLC + differential + predictor (mpeg1.js:705-751)
		const uint32_t w = br.peek32();
		const uint32_t e = block < 4 ? lds_u16(sbase + OFF_DC_LUMA + (w >> 25) * 2u)
		                             : lds_u16(sbase + OFF_DC_CHROMA + (w >> 24) * 2u);
		const int len = e & 31, size = e >> 5;
		if (len == 0) return false;
		br.consume(len);
		int *pred = block < 4 ? &ps.dc_y : (block == 4 ? &ps.dc_b4 : &ps.dc_b5);
		int dc = *pred;
		if (size > 0) {
			const int diff = (int)br.read(size);
			dc += (diff & (1 << (size - 1))) ? diff : (int)((0xffffffffu << size) | (uint32_t)(diff + 1));
		}
		*pred = dc;
		dc8 = max(-32768, min(32767, dc * 8));  // x PREMULTIPLIER[0] = dc << 8 in stage 2 (mpeg1.js:747)
		n = 1;
	}
	if (lane == 0) *reinterpret_cast<uint2 *>(slot) = make_uint2(br.bitpos(), (uint32_t)dc8 & 0xffffu);
	if (!intra && (br.peek32() >> 31)) {  // dct_coeff_first: a leading '1' is (0, +-1), never end_of_block
		br.consume(2);
		n = 1;
	}
	for (;;) {
		const uint32_t w = br.peek32();
		// (Resolving '10' / '11s' arithmetically before the look-up was measured 8 % SLOWER: it defeats the
		// combining of several codes per look-up.)
		// as many complete codes as fit in the next 13 bits, in one look-up
		const uint32_t m = lds_u16(sbase + OFF_MS + (w >> (32 - MS_BITS)) * 2u);
		if (m & 15u) {
			n += (int)((m >> 4) & 63u);
			br.consume((int)(m & 15u));
			if (m & 0x400u) break;  // the last code consumed was end_of_block
			continue;
		}
		// long code or escape: one symbol through the clz-indexed table
		const int z = __clz((int)w);
		if (z > VLC_DCT_MAX_Z) return false;
		const uint32_t e = lds_u16(sbase + OFF_DCT + (((uint32_t)z << 5) | ((w << (z + 1)) >> 27)) * 2u);
		if (e >> 11) {
			// escape (mpeg1.js:767-780): 6-bit code, 6-bit run, 8 (+8) bit level.  (end_of_block is
			// two bits and always resolved by the multi-symbol table.)
			n += (int)((w >> 20) & 63u) + 1;
			br.consume((w & 0x0007f000u) ? 20 : 28);  // level byte 0 or 128 -> a second byte follows
			continue;
		}
		if (e == 0) return false;  // hole in the code space
		n += (int)(e >> 5);
		br.consume((int)(e & 31u));
	}
	if (n > 64) ps.error = PARSE_ERR_COEF_INDEX;  // some run pushed the index past 63 (stores dropped, like JS)
	dc_only = (n == 1);  // mpeg1.js:838, 850
	ps.n_coded++;
	return true;
}

// mpeg1.js:395-457, one component
__device__ __forceinline__ bool parse_motion(BitReader &br, uint32_t sbase, const PictureState &ps, int &prev, int &mv) {
	const uint32_t e = clz_lut(sbase + OFF_MOTION, br.peek32(), VLC_MOTION_MAX_Z);
	const int len = e & 31;
	if (len == 0) return false;
	br.consume(len);
	const int code = (int)(e >> 5) - 16;
	int d = code;
	if (code != 0 && ps.f != 1) {
		const int r = (int)br.read(ps.r_size);
		d = ((abs(code) - 1) << ps.r_size) + r + 1;
		if (code < 0) d = -d;
	}
	prev += d;
	if (prev > (ps.f << 4) - 1) prev -= ps.f << 5;
	else if (prev < -(ps.f << 4)) prev += ps.f << 5;
	mv = ps.full_pel ? prev * 2 : prev;
	return true;
}

__device__ __forceinline__ int read_mba(BitReader &br, uint32_t sbase) {
	const uint32_t e = clz_lut(sbase + OFF_MBA, br.peek32(), VLC_MBA_MAX_Z);
	const int len = e & 31;
	if (len == 0) return -1;
	br.consume(len);
	return (int)(e >> 5);
}

__device__ __forceinline__ uint4 pack_record(int mv_h, int mv_v, int flags, int cbp, int dc_only, int qscale, uint32_t bit_pos) {
	uint4 r;
	r.x = ((uint32_t)mv_h & 0xffffu) | ((uint32_t)mv_v << 16);
	r.y = (uint32_t)flags | ((uint32_t)cbp << 8) | ((uint32_t)dc_only << 16) | ((uint32_t)qscale << 24);
	r.z = bit_pos;
	r.w = 0;
	return r;
}

// mpeg1.js:294-392 decodeMacroblock.  false = stop walking this slice.
__device__ bool walk_macroblock(BitReader &br, uint32_t sbase, PictureState &ps, const ParseTask &t, int mb_size, int lane) {
	int increment = 0;
	int v = read_mba(br, sbase);
	while (v == 34) v = read_mba(br, sbase);                       // macroblock_stuffing
	while (v == 35) { increment += 33; v = read_mba(br, sbase); }  // macroblock_escape
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
			const int n_skip = increment - 1;
			const uint4 rec = pack_record(ps.mv_h, ps.mv_v, MBF_PRESENT | MBF_SKIPPED, 0, 0, ps.qscale, br.bitpos());
			for (int k = lane; k < n_skip; k += 32) reinterpret_cast<uint4 *>(t.hdr)[ps.mb_addr + 1 + k] = rec;
			ps.n_present += n_skip;
			ps.mb_addr += n_skip;
		}
		ps.mb_addr++;
	}
	const int mb = ps.mb_addr;
	if (mb < 0 || mb >= mb_size) return false;  // outside the picture: never write there

	const uint32_t w = br.peek32();
	const uint32_t e = ps.picture_type == 1 ? lds_u16(sbase + OFF_TYPE_I + (w >> 30) * 2u)
	                                        : lds_u16(sbase + OFF_TYPE_P + (w >> 26) * 2u);
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
			if (!parse_motion(br, sbase, ps, ps.mv_h_prev, ps.mv_h)) return false;
			if (!parse_motion(br, sbase, ps, ps.mv_v_prev, ps.mv_v)) return false;
		} else if (ps.picture_type == 2) {
			ps.mv_h = ps.mv_v = ps.mv_h_prev = ps.mv_v_prev = 0;  // mpeg1.js:452-456
		}
	}

	int cbp = intra ? 0x3f : 0;
	if (type & 0x02) {
		const uint32_t ce = clz_lut(sbase + OFF_CBP, br.peek32(), VLC_CBP_MAX_Z);
		if ((ce & 31) == 0) return false;
		br.consume(ce & 31);
		cbp = ce >> 5;
	}

	const int mv_h = ps.mv_h, mv_v = ps.mv_v, qscale = ps.qscale;
	uint32_t *coef_mb = reinterpret_cast<uint32_t *>(t.coef) + (size_t)mb * (MB_COEF_INT16 / 2);
	int done = 0, dc_mask = 0;
	bool ok = true;
#pragma unroll 1
	for (int block = 0; block < 6; block++) {
		if (cbp & (0x20 >> block)) {
			bool dc_only;
			ok = walk_block(br, sbase, ps, intra, block, coef_mb + block * 32, lane, dc_only);
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

__global__ void __launch_bounds__(WALK_THREADS)
walk_pictures_kernel(const ParseTask *__restrict__ tasks, int n_tasks, const uint4 *__restrict__ ms_table) {
	extern __shared__ __align__(128) uint8_t smem[];
	{
		uint16_t *s16 = reinterpret_cast<uint16_t *>(smem);
		for (int i = threadIdx.x; i < (VLC_DCT_MAX_Z + 1) * 32; i += WALK_THREADS) s16[OFF_DCT / 2 + i] = walk_entry(VLC_DCT_COEFF[i]);
		for (int i = threadIdx.x; i < (VLC_MBA_MAX_Z + 1) * 32; i += WALK_THREADS) s16[OFF_MBA / 2 + i] = VLC_MBA[i];
		for (int i = threadIdx.x; i < (VLC_CBP_MAX_Z + 1) * 32; i += WALK_THREADS) s16[OFF_CBP / 2 + i] = VLC_CBP[i];
		for (int i = threadIdx.x; i < (VLC_MOTION_MAX_Z + 1) * 32; i += WALK_THREADS) s16[OFF_MOTION / 2 + i] = VLC_MOTION[i];
		for (int i = threadIdx.x; i < 128; i += WALK_THREADS) s16[OFF_DC_LUMA / 2 + i] = VLC_DC_SIZE_LUMA[i];
		for (int i = threadIdx.x; i < 256; i += WALK_THREADS) s16[OFF_DC_CHROMA / 2 + i] = VLC_DC_SIZE_CHROMA[i];
		for (int i = threadIdx.x; i < 4; i += WALK_THREADS) s16[OFF_TYPE_I / 2 + i] = VLC_MBTYPE_I[i];
		for (int i = threadIdx.x; i < 64; i += WALK_THREADS) s16[OFF_TYPE_P / 2 + i] = VLC_MBTYPE_P[i];
		uint4 *ms = reinterpret_cast<uint4 *>(smem + OFF_MS);
		for (int i = threadIdx.x; i < (2 << MS_BITS) / 16; i += WALK_THREADS) ms[i] = __ldg(ms_table + i);
	}
	__syncthreads();

	const int lane = threadIdx.x & 31;
	const int task_id = blockIdx.x * (WALK_THREADS / 32) + (threadIdx.x >> 5);
	if (task_id >= n_tasks) return;
	const ParseTask t = tasks[task_id];
	const uint32_t sbase = smem_base(smem);
	const int mb_width = t.seq->mb_width, mb_size = t.seq->mb_size;

	// no macroblock is present until the walk reaches it (an address no slice covers keeps the
	// two-pictures-old samples)
	for (int i = lane; i < mb_size; i += 32) reinterpret_cast<uint4 *>(t.hdr)[i] = make_uint4(0, 0, 0, 0);
	__syncwarp();

	BitReader br;
	br.words = reinterpret_cast<const uint32_t *>(t.es);
	br.bytes = t.es;
	br.len = t.es_len;
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
				if (!walk_macroblock(br, sbase, ps, t, mb_size, lane)) {
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

// ==================================================================================================
// 1b: expand every coded block (one thread per block slot)

__global__ void __launch_bounds__(CTA_THREADS)
expand_blocks_kernel(const ParseTask *__restrict__ tasks) {
	extern __shared__ __align__(128) uint8_t smem[];
	{
		uint16_t *s16 = reinterpret_cast<uint16_t *>(smem);
		for (int i = threadIdx.x; i < (VLC_DCT_MAX_Z + 1) * 32; i += CTA_THREADS) s16[OFF_DCT / 2 + i] = VLC_DCT_COEFF[i];
		for (int i = threadIdx.x; i < 64; i += CTA_THREADS) smem[OFF_ZIGZAG + i] = TBL_ZIG_ZAG[i];
		uint32_t *blocks = reinterpret_cast<uint32_t *>(smem + OFF_BLOCKS);
		for (int i = threadIdx.x; i < CTA_THREADS * (int)(TILE_PITCH / 4); i += CTA_THREADS) blocks[i] = 0u;
	}
	__syncthreads();

	const ParseTask &t = tasks[blockIdx.y];
	const int mb_size = t.seq->mb_size;
	const int slot_id = blockIdx.x * CTA_THREADS + threadIdx.x;  // mb * 6 + block
	if (slot_id >= mb_size * 6) return;
	if (t.info->status != PIC_DECODED) return;
	const int mb = slot_id / 6, block = slot_id - mb * 6;
	const uint32_t rec = reinterpret_cast<const uint32_t *>(t.hdr + mb)[1];
	if (!(rec & MBF_PRESENT) || !((rec >> 8) & (0x20u >> block))) return;
	const bool intra = rec & MBF_INTRA;
	const int qs = (int)(rec >> 24);
	const uint8_t *__restrict__ quant = intra ? t.seq->intra_q : t.seq->non_intra_q;

	uint4 *slot = reinterpret_cast<uint4 *>(t.coef) + (size_t)slot_id * 8;
	const uint2 parked = *reinterpret_cast<const uint2 *>(slot);  // left by the walk
	const uint32_t sbase = smem_base(smem);
	// this thread's 64 x int16 tile, linear (it leaves as one bulk copy); tiles are 144 bytes apart so
	// that lanes writing the same coefficient index spread over the banks
	const uint32_t sblock = sbase + OFF_BLOCKS + threadIdx.x * TILE_PITCH;

	BitReader br;
	br.words = reinterpret_cast<const uint32_t *>(t.es);
	br.bytes = t.es;
	br.len = t.es_len;
	br.seek_byte(parked.x >> 3);
	if (parked.x & 7u) br.consume((int)(parked.x & 7u));

	int n = 0;
	if (intra) {
		sts_s16(sblock, (int)(int16_t)(parked.y & 0xffffu));  // coefficient 0
		n = 1;
	}
	bool first = !intra;
	for (;;) {  // mpeg1.js:757-811; the walk has already validated every code of this block
		const uint32_t w = br.peek32();
		const int z = min(__clz((int)w), VLC_DCT_MAX_Z);
		const uint32_t e = lds_u16(sbase + OFF_DCT + (((uint32_t)z << 5) | ((w << (z + 1)) >> 27)) * 2u);
		int len = e & 31;
		int run = (e >> 5) & 31;
		int level = e >> 10;
		if (first && z == 0) {  // '1s'
			len = 1;
			run = 0;
			level = 1;
		}
		first = false;
		if (level == 0) {
			if (run != 0 || len == 0) break;  // end_of_block (or, defensively, an invalid code)
			// escape (mpeg1.js:767-780)
			run = (w >> 20) & 63;
			const int l8 = (w >> 12) & 255;
			if ((l8 & 127) == 0) {
				level = (int)((w >> 4) & 255) - (l8 << 1);  // l8 == 128: second byte - 256
				br.consume(28);
			} else {
				level = l8 > 128 ? l8 - 256 : l8;
				br.consume(20);
			}
		} else {
			if ((w >> (31 - len)) & 1u) level = -level;
			br.consume(len + 1);
		}
		n += run;
		if (n > 63) {  // JS: ZIG_ZAG[n] undefined -> the store is a no-op (the walk flagged the picture)
			if (n > 4096) break;
			n++;
			continue;
		}
		const uint32_t idx = lds_u8(sbase + OFF_ZIGZAG + (uint32_t)n);
		n++;
		// dequantise, oddify toward zero, clip (mpeg1.js:794-807)
		level <<= 1;
		if (!intra) level += level < 0 ? -1 : 1;
		level = (level * qs * (int)__ldg(quant + idx)) >> 4;
		if ((level & 1) == 0) level -= level > 0 ? 1 : -1;
		level = max(-2048, min(2047, level));
		sts_s16(sblock + idx * 2u, level);
	}
	// The finished block leaves as ONE 128-byte TMA bulk store (shared -> global, SASS UBLKCP): whole
	// lines reach L2, whereas eight 16-byte stores per thread half-fill 32-byte sectors and made L2
	// read every sector back before merging (ncu: 23.5 GB read for 22 GB written per step).
	asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the copy engine
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 128;" ::"l"(slot), "r"(sblock) : "memory");
	asm volatile("cp.async.bulk.commit_group;" ::: "memory");
	asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the tile must outlive the read
}

}  // namespace

// Multi-symbol walk table: for every 13-bit prefix, the complete dct_coeff_next codes (with their
// sign bits) that fit, greedily.  Entry: bits 0..3 = bits to consume (0 = first code does not fit
// or is an escape: take the single-symbol path), bits 4..9 = sum of (run + 1), bit 10 = the last
// code consumed was end_of_block.  Built once per device from the same generated DCT table.
static const uint16_t *ms_table_for_current_device() {
	static uint16_t *tables[64] = {};
	static std::mutex lock;  // decoders may be driven from several host threads
	std::lock_guard<std::mutex> guard(lock);
	int dev = 0;
	CUDA_CHECK(cudaGetDevice(&dev));
	if (tables[dev]) return tables[dev];
	CUDA_CHECK(cudaFuncSetAttribute(walk_pictures_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                                (int)(OFF_MS + (2u << MS_BITS))));  // per device, once
	std::vector<uint16_t> dct((VLC_DCT_MAX_Z + 1) * 32);
	CUDA_CHECK(cudaMemcpyFromSymbol(dct.data(), VLC_DCT_COEFF, dct.size() * sizeof(uint16_t)));
	std::vector<uint16_t> ms(1u << MS_BITS);
	for (uint32_t prefix = 0; prefix < (1u << MS_BITS); prefix++) {
		const uint32_t w = prefix << (32 - MS_BITS);
		int pos = 0, n = 0, eob = 0;
		for (;;) {
			const uint32_t v = w << pos;  // bits beyond the prefix read as 0 and are never trusted: lengths are checked
			int z = 0;
			while (z < 32 && !((v << z) & 0x80000000u)) z++;
			if (z > VLC_DCT_MAX_Z) break;
			const uint16_t e = dct[(z << 5) | ((z + 1 < 32 ? (v << (z + 1)) : 0u) >> 27)];
			const int len = e & 31, run = (e >> 5) & 31, level = e >> 10;
			if (len == 0) break;
			if (level == 0) {
				if (run == 1 && pos + 2 <= MS_BITS) { pos += 2; eob = 1; }
				break;  // escape: single-symbol path
			}
			if (pos + len + 1 > MS_BITS) break;
			pos += len + 1;
			n += run + 1;
		}
		ms[prefix] = (uint16_t)(pos | (n << 4) | (eob << 10));
	}
	uint16_t *d = nullptr;
	CUDA_CHECK(cudaMalloc(&d, ms.size() * sizeof(uint16_t)));
	CUDA_CHECK(cudaMemcpy(d, ms.data(), ms.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
	tables[dev] = d;
	return d;
}

void launch_parse_pictures(const ParseTask *tasks, int n_tasks, int max_mb_size, cudaStream_t stream,
                           cudaEvent_t walk_done, const ParseFork *fork) {
	if (n_tasks <= 0) return;
	const uint16_t *ms = ms_table_for_current_device();
	const int per_cta = WALK_THREADS / 32;
	const size_t walk_smem = OFF_MS + (2u << MS_BITS);
	// The wave arrives sorted by picture size, largest first.  Group 0 (largest pictures) stays on
	// `stream`; the other groups go to side streams, each walk followed by its own expand.
	// JSMPEG_B200_PARSE_GROUPS=1 keeps stage 1 on one stream (used for the ncu launch list: ncu
	// serialises concurrent kernels, so only the unforked run has comparable shares)
	static const int max_groups = [] {
		const char *e = getenv("JSMPEG_B200_PARSE_GROUPS");
		const int g = e ? atoi(e) : PARSE_GROUPS;
		return g < 1 ? 1 : (g > PARSE_GROUPS ? PARSE_GROUPS : g);
	}();
	const int groups = (fork && n_tasks >= 64 * max_groups) ? max_groups : 1;
	if (groups > 1) CUDA_CHECK(cudaEventRecord(fork->fork, stream));
	for (int g = 0; g < groups; g++) {
		// equal groups; measured on the 3840-picture wave: unforked 59.8 ms, 4 groups 52.4, 8 groups 50.5,
		// 16 groups 80.8 (too many concurrent kernels), a small first group 55.0
		const int lo = (int)((long)n_tasks * g / groups), hi = (int)((long)n_tasks * (g + 1) / groups);
		const int n = hi - lo;
		if (n <= 0) continue;
		cudaStream_t st = g == 0 ? stream : fork->side[g];
		if (g > 0) CUDA_CHECK(cudaStreamWaitEvent(st, fork->fork, 0));
		walk_pictures_kernel<<<(n + per_cta - 1) / per_cta, WALK_THREADS, walk_smem, st>>>(
		    tasks + lo, n, reinterpret_cast<const uint4 *>(ms));
		if (g == 0 && walk_done) CUDA_CHECK(cudaEventRecord(walk_done, st));
		dim3 grid((max_mb_size * 6 + CTA_THREADS - 1) / CTA_THREADS, n);
		expand_blocks_kernel<<<grid, CTA_THREADS, OFF_BLOCKS + CTA_THREADS * TILE_PITCH, st>>>(tasks + lo);
		if (g > 0) {
			CUDA_CHECK(cudaEventRecord(fork->join[g], st));
			CUDA_CHECK(cudaStreamWaitEvent(stream, fork->join[g], 0));
		}
	}
}
