// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for SDSoC <sds_lib.h>: the
// contiguous/non-cacheable allocators become zero-filled host allocations
// (used at src/csr_hw.cpp:180,1446 and freed at :1428,1462).
#ifndef ORACLE_STANDIN_SDS_LIB_H
#define ORACLE_STANDIN_SDS_LIB_H
#include <cstdlib>
static inline void *sds_alloc_non_cacheable(size_t n) { return calloc(n ? n : 1, 1); }
static inline void *sds_alloc(size_t n) { return calloc(n ? n : 1, 1); }
static inline void sds_free(void *p) { free(p); }
#endif
