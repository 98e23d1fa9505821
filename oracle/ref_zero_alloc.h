/*
 * ref_zero_alloc.h -- TEST INFRASTRUCTURE ONLY.  Force-included (-include) when the reference's
 * descriptor.cpp is compiled for oracle/_ref; the reference source itself is untouched.
 *
 * Descriptor leaves the border pixels of I_desc unwritten (descriptor.cpp:48-112 only
 * fill u in [3,W-4], v in [3,H-4]) and the matcher's stage dump reads them.  With the stock
 * allocator those bytes are zero when glibc serves the block from fresh pages and garbage when it
 * recycles a freed chunk -- observed: the same test passing or failing depending on which tests
 * ran before it.  Here every _mm_malloc of that translation unit is a fresh anonymous mapping:
 * zero-filled like the common case of the stock build, with no extra pass over the memory (the
 * CPU baseline timing is not handicapped).
 */
#ifndef REF_ZERO_ALLOC_H
#define REF_ZERO_ALLOC_H
#include <stddef.h>
#include <stdint.h>
#include <sys/mman.h>
#include <emmintrin.h> /* defines the real _mm_malloc / _mm_free before they are renamed below */

static inline void* ref_zeroed_mm_malloc(size_t n, size_t align) {
  (void)align; /* page alignment covers every alignment the reference asks for (16) */
  const size_t head = 4096;
  uint8_t* p = (uint8_t*)mmap(NULL, n + head, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
  if (p == (uint8_t*)MAP_FAILED) return NULL;
  *(size_t*)p = n + head;
  return p + head;
}
static inline void ref_zeroed_mm_free(void* q) {
  if (!q) return;
  uint8_t* p = (uint8_t*)q - 4096;
  munmap(p, *(size_t*)p);
}
#define _mm_malloc(n, a) ref_zeroed_mm_malloc((n), (a))
#define _mm_free(p) ref_zeroed_mm_free((p))
#endif
