// bitbuffer_test.cpp -- test shim (tests only): jsmpeg_b200/csrc/bitbuffer.h over plain malloc with
// EXACT-size allocations, so that AddressSanitizer sees any write the protocol does not cover.
#include <stdlib.h>

#include "../../jsmpeg_b200/csrc/bitbuffer.h"

namespace {
void *plain_alloc(size_t n, void *) { return malloc(n); }
void plain_release(void *p, void *) { free(p); }
const bitbuffer::Allocator kAlloc = {plain_alloc, plain_release, nullptr};
}  // namespace

extern "C" {
struct bbt_t {
	bitbuffer::Buffer b;
	int moved_count;
};
bbt_t *bbt_create(unsigned cap, int mode) {
	bbt_t *t = new bbt_t();
	t->b.mode = mode;
	t->moved_count = 0;
	bitbuffer::resize(t->b, cap, kAlloc);
	return t;
}
void bbt_destroy(bbt_t *t) {
	free(t->b.bytes);
	delete t;
}
// the protocol as a caller uses it: ask for room, fill all n bytes, commit.  Returns the write offset or -1.
long bbt_write(bbt_t *t, const unsigned char *src, unsigned n) {
	bool moved = false;
	unsigned char *p = bitbuffer::get_write_ptr(t->b, n, kAlloc, moved);
	if (!p) return -1;
	if (moved) t->moved_count++;
	memcpy(p, src, n);  // ASan: must be inside the allocation
	const long off = (long)(p - t->b.bytes);
	t->b.length += n;
	return off;
}
void bbt_set_index(bbt_t *t, unsigned index) { t->b.index = index; }
void bbt_state(bbt_t *t, unsigned *out) {
	out[0] = t->b.capacity; out[1] = t->b.length; out[2] = t->b.index; out[3] = (unsigned)t->moved_count;
}
const unsigned char *bbt_bytes(bbt_t *t) { return t->b.bytes; }
}
