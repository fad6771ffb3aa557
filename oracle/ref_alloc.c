/* oracle/_ref: the two libobs allocator entry points the reference's
 * callbacks really execute (bzalloc = zeroing malloc, bfree = free).
 * TEST INFRASTRUCTURE. */
#include <stdlib.h>
void *bzalloc(size_t size) { return calloc(1, size ? size : 1); }
void bfree(void *ptr) { free(ptr); }
