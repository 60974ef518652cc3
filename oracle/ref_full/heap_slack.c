/* TEST INFRASTRUCTURE.  reallocate_memory_for_fragmentation (src/allocations.c:556-645) carves the
 * fragmentation arrays out of a block of memory.frag_allocated bytes, but aligns every array to 32
 * bytes while organize_main_memory (src/allocations.c:173-209) accounts for the six integer arrays
 * without that padding: the last array (frag_map_update) ends up to 2 x 32 bytes past the block
 * (AddressSanitizer: heap-buffer-overflow in create_map, src/fragment.c:713, in the UNMODIFIED
 * reference at 32^3, 64^3 and 128^3).  In the reference's production runs the block is a page-rounded
 * mmap region and the overshoot lands in its slack; when glibc serves the block from the heap it
 * clobbers the next chunk header ("malloc(): invalid size").  The oracle executables are linked
 * with -Wl,--wrap=realloc so that every realloc carries one page of slack. */
#include <stdlib.h>
void* __real_realloc(void* p, size_t n);
void* __wrap_realloc(void* p, size_t n) { return __real_realloc(p, n + 4096); }
