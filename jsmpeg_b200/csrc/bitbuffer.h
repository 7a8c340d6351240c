// bitbuffer.h -- the host bit buffer's write protocol (plain C++, no CUDA): the two-phase
// get_write_ptr / did_write of the reference (src/wasm/buffer.c:48-71), its EXPAND growth
// (buffer.c:157-164, src/buffer.js:20-28, 84-90) and EVICT compaction (buffer.c:167-190,
// src/buffer.js:30-62).
//
// Header-only and allocator-agnostic so that the same text is compiled into the product (pinned host
// memory, engine.cu) and into a CPU test under AddressSanitizer (tests/emu/bitbuffer_test.cpp).
//
// One deliberate difference from the reference.  The reference sizes an expansion as
// max(2 * capacity, n - available), which is smaller than length + n whenever the buffer is partly
// full and n is large (capacity 1000, length 900, n 1500 -> 2000 where 2400 are needed); and its
// EVICT mode hands out the write position even when the bytes still do not fit after the
// eviction.  The C build then writes past the allocation, the JS build throws a RangeError from
// TypedArray.set (src/buffer.js:111).  Neither is a behaviour a plugin can reproduce, so here the
// returned pointer ALWAYS has room for n bytes: EXPAND grows to max(2 * capacity, length + n), and
// EVICT grows the buffer when the request exceeds its capacity even after everything unread has been
// dropped.  Whenever the reference's own arithmetic leaves enough room, capacity, length and index
// are exactly the reference's.
#pragma once
#include <stdint.h>
#include <string.h>

namespace bitbuffer {

enum { MODE_EVICT = 1, MODE_EXPAND = 2 };  // src/wasm/buffer.h:8-11

struct Buffer {
	uint8_t *bytes = nullptr;
	uint32_t capacity = 0, length = 0, index = 0;  // index in BITS
	int mode = MODE_EXPAND;
};

// alloc(bytes) may throw or return nullptr (then get_write_ptr returns nullptr and changes nothing)
struct Allocator {
	void *(*alloc)(size_t bytes, void *ctx);
	void (*release)(void *p, void *ctx);
	void *ctx;
};

// buffer.c:157-164 resize: a new allocation of `cap` bytes holding the first min(length, cap) bytes
inline bool resize(Buffer &s, uint32_t cap, const Allocator &a) {
	uint8_t *n = static_cast<uint8_t *>(a.alloc(cap ? cap : 1, a.ctx));
	if (!n) return false;
	if (s.bytes) {
		if (s.length > cap) s.length = cap;
		if (s.length) memcpy(n, s.bytes, s.length);
		a.release(s.bytes, a.ctx);
	}
	s.bytes = n;
	s.capacity = cap;
	if (s.index > (s.length << 3)) s.index = s.length << 3;
	return true;
}

// Room for n more bytes at the write position.  `moved` is set when bytes already in the buffer were
// dropped or moved (EVICT), i.e. when byte positions remembered by the caller (a device mirror, a
// start-code index, parsed-ahead pictures) are void.  nullptr: the request cannot be met (more than
// 4 GiB - 1 in one buffer, or the allocator failed); nothing has changed then.
inline uint8_t *get_write_ptr(Buffer &s, uint32_t n, const Allocator &a, bool &moved) {
	moved = false;
	const uint32_t avail = s.capacity - s.length;
	if (n <= avail) return s.bytes + s.length;
	if (s.mode == MODE_EXPAND) {
		const uint64_t need = (uint64_t)s.length + n;
		if (need > 0xffffffffull) return nullptr;
		uint64_t cap = (uint64_t)s.capacity * 2;  // buffer.c:53-57
		if (cap < need) cap = need;               // the reference: cap = n - avail, too small when length > 0
		if (cap > 0xffffffffull) cap = 0xffffffffull;
		if (!resize(s, (uint32_t)cap, a)) return nullptr;
		return s.bytes + s.length;
	}
	// EVICT (buffer.c:167-190)
	const uint32_t pos = s.index >> 3;
	if (pos >= s.length || n > avail + pos) {  // nothing unread, or emergency evacuation: drop everything
		if (s.length) moved = true;
		s.length = 0;
		s.index = 0;
	} else if (pos != 0) {
		memmove(s.bytes, s.bytes + pos, s.length - pos);
		s.length -= pos;
		s.index -= pos << 3;
		moved = true;
	}
	if (n > s.capacity - s.length) {  // still no room (only after an evacuation, or with nothing read yet)
		const uint64_t need = (uint64_t)s.length + n;
		if (need > 0xffffffffull || !resize(s, (uint32_t)need, a)) return nullptr;
	}
	return s.bytes + s.length;
}

}  // namespace bitbuffer
