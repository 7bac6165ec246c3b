/* METIS prototypes for linking the PanguLU reference against the METIS build
 * NVIDIA ships inside the CUDA toolkit (targets/x86_64-linux/lib/
 * libmetis_static.a, 64-bit idx_t -- probed with a 20x20 grid NodeND call).
 * Test infrastructure only. */
#ifndef PANGULU_ORACLE_METIS_SHIM_H
#define PANGULU_ORACLE_METIS_SHIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef int64_t idx_t;
typedef float real_t;
#define METIS_OK 1
#define METIS_NOPTIONS 40
int METIS_NodeND(idx_t *nvtxs, idx_t *xadj, idx_t *adjncy, idx_t *vwgt, idx_t *options, idx_t *perm, idx_t *iperm);
int METIS_SetDefaultOptions(idx_t *options);
#ifdef __cplusplus
}
#endif
#endif
