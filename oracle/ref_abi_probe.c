/* Field offsets of the reference's structs as compiled WITH -DGPU_OPEN (device twin pointers present):
 * the layout the sm_100a platform module must agree with.  Compiled alone by oracle/Makefile into
 * oracle/_ref/libpangulu_ref_abi_gpu.so; tests/test_abi.py compares it with include/pangulu_sm100_abi.h.
 * Test infrastructure only. */
#include <stddef.h>
#define PANGULU_PLATFORM_ENV
#include "pangulu_common.h"

int refabi_count(void) { return 30; }
void refabi_offsets(long long *out)
{
    int k = 0;
    out[k++] = offsetof(pangulu_storage_slot_t, brow_pos);
    out[k++] = offsetof(pangulu_storage_slot_t, bcol_pos);
    out[k++] = offsetof(pangulu_storage_slot_t, columnpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, rowindex);
    out[k++] = offsetof(pangulu_storage_slot_t, value);
    out[k++] = offsetof(pangulu_storage_slot_t, rowpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, columnindex);
    out[k++] = offsetof(pangulu_storage_slot_t, idx_of_csc_value_for_csr);
    out[k++] = offsetof(pangulu_storage_slot_t, data_status);
    out[k++] = offsetof(pangulu_storage_slot_t, related_block);
    out[k++] = offsetof(pangulu_storage_slot_t, is_upper);
    out[k++] = offsetof(pangulu_storage_slot_t, bin_id);
    out[k++] = offsetof(pangulu_storage_slot_t, slot_idx);
    out[k++] = offsetof(pangulu_storage_slot_t, task_queue);
    out[k++] = offsetof(pangulu_storage_slot_t, d_columnpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, d_rowindex);
    out[k++] = offsetof(pangulu_storage_slot_t, d_value);
    out[k++] = offsetof(pangulu_storage_slot_t, d_rowpointer);
    out[k++] = offsetof(pangulu_storage_slot_t, d_columnindex);
    out[k++] = offsetof(pangulu_storage_slot_t, d_idx_of_csc_value_for_csr);
    out[k++] = sizeof(pangulu_storage_slot_t);
    out[k++] = offsetof(pangulu_task_t, row);
    out[k++] = offsetof(pangulu_task_t, col);
    out[k++] = offsetof(pangulu_task_t, kernel_id);
    out[k++] = offsetof(pangulu_task_t, task_level);
    out[k++] = offsetof(pangulu_task_t, compare_flag);
    out[k++] = offsetof(pangulu_task_t, opdst);
    out[k++] = offsetof(pangulu_task_t, op1);
    out[k++] = offsetof(pangulu_task_t, op2);
    out[k++] = sizeof(pangulu_task_t);
}
/* task kinds and the owner map, src/pangulu_common.h:124-127,135 */
int refabi_task_ids(int which) { return which == 0 ? PANGULU_TASK_GETRF : which == 1 ? PANGULU_TASK_TSTRF : which == 2 ? PANGULU_TASK_GESSM : PANGULU_TASK_SSSSM; }
int refabi_calc_rank(int row, int col, int p, int q) { return PANGULU_CALC_RANK(row, col, p, q); }
