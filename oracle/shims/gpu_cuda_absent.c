/* The reference tree ships a pre-generated platform switch
 * (src/pangulu_platform_helper.c) that names both built-in platforms, so a CPU-only
 * oracle build still has to resolve the pangulu_platform_0201000_* symbols.
 * Without -DGPU_OPEN they are never reached (PANGULU_DEFAULT_PLATFORM is
 * CPU_NAIVE, src/pangulu_common.h:4-9); these definitions trap if they are.
 * Test infrastructure only. */
#define PANGULU_PLATFORM_ENV
#include "pangulu_common.h"

static void absent(const char *fn)
{
    fprintf(stderr, "[oracle] pangulu_platform_0201000_%s reached in a CPU-only oracle build\n", fn);
    abort();
}
void pangulu_platform_0201000_malloc(void **a, size_t s) { absent("malloc"); }
void pangulu_platform_0201000_malloc_pinned(void **a, size_t s) { absent("malloc_pinned"); }
void pangulu_platform_0201000_synchronize() { absent("synchronize"); }
void pangulu_platform_0201000_memset(void *s, int c, size_t n) { absent("memset"); }
void pangulu_platform_0201000_create_stream(void **s) { absent("create_stream"); }
void pangulu_platform_0201000_memcpy(void *d, const void *s, size_t n, unsigned int k) { absent("memcpy"); }
void pangulu_platform_0201000_memcpy_async(void *d, const void *s, size_t n, unsigned int k, void *st) { absent("memcpy_async"); }
void pangulu_platform_0201000_free(void *p) { absent("free"); }
void pangulu_platform_0201000_get_device_num(int *n) { absent("get_device_num"); }
void pangulu_platform_0201000_set_default_device(int n) { absent("set_default_device"); }
void pangulu_platform_0201000_get_device_name(char *name, int n) { absent("get_device_name"); }
void pangulu_platform_0201000_get_device_memory_usage(size_t *u) { absent("get_device_memory_usage"); }
void pangulu_platform_0201000_getrf(pangulu_inblock_idx nb, pangulu_storage_slot_t *d, int tid) { absent("getrf"); }
void pangulu_platform_0201000_tstrf(pangulu_inblock_idx nb, pangulu_storage_slot_t *d, pangulu_storage_slot_t *g, int tid) { absent("tstrf"); }
void pangulu_platform_0201000_gessm(pangulu_inblock_idx nb, pangulu_storage_slot_t *d, pangulu_storage_slot_t *g, int tid) { absent("gessm"); }
void pangulu_platform_0201000_ssssm(pangulu_inblock_idx nb, pangulu_storage_slot_t *d, pangulu_storage_slot_t *a, pangulu_storage_slot_t *b, int tid) { absent("ssssm"); }
void pangulu_platform_0201000_ssssm_batched(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_task_t *t) { absent("ssssm_batched"); }
void pangulu_platform_0201000_hybrid_batched(pangulu_inblock_idx nb, pangulu_uint64_t n, pangulu_task_t *t) { absent("hybrid_batched"); }
void pangulu_platform_0201000_spmv(pangulu_inblock_idx nb, pangulu_storage_slot_t *a, calculate_type *x, calculate_type *y) { absent("spmv"); }
void pangulu_platform_0201000_vecadd(pangulu_int64_t n, calculate_type *b, calculate_type *x) { absent("vecadd"); }
void pangulu_platform_0201000_sptrsv(pangulu_inblock_idx nb, pangulu_storage_slot_t *s, calculate_type *x, pangulu_int64_t uplo) { absent("sptrsv"); }
