#!/usr/bin/env python
"""Build-time overlay for PanguLU: route pangulu_gstrf to the sm_100a plan path.

    python apply_overlay.py <reference src/pangulu.c> <output pangulu.c>

Writes a patched COPY of PanguLU's src/pangulu.c (never the original; the copy lands in a git-ignored build
directory): inside pangulu_gstrf the call of pangulu_numeric (src/pangulu.c:236-238) is wrapped so that a build with
-DPANGULU_SM100_PLAN and one rank hands the block storage PanguLU's own init built to
pangulu_sm100_reference_gstrf() (pangulu_b200/csrc/host/reference_adapter.cpp): static schedule, compact tile / compact
panel kernels, persistent executor.  Several ranks keep pangulu_numeric.  This is the whole integration patch --
INTEGRATION.md shows it as a diff a PanguLU maintainer would apply."""
import re
import sys

DECL = '''
#ifdef PANGULU_SM100_PLAN
/* sm_100a platform module, host side (include/pangulu_sm100_driver.h) */
void pangulu_sm100_reference_gstrf(int nb, unsigned int n, unsigned int block_length, pangulu_storage_slot_t *slots,
                                   long long slot_count, long long nondiag_slot_cnt);
#endif
'''

CALL = '''#ifdef PANGULU_SM100_PLAN
    if (size == 1)
        pangulu_sm100_reference_gstrf(block_common->nb, block_common->n, block_common->block_length,
                                      block_smatrix->storage->bins[0].slots, block_smatrix->storage->bins[0].slot_count,
                                      block_smatrix->storage->bins[0].nondiag_slot_cnt);
    else
#endif
'''


def main(src, dst):
    text = open(src).read()
    m = re.search(r'void\s+pangulu_gstrf\s*\(', text)
    if not m:
        raise SystemExit("pangulu_gstrf not found in " + src)
    call = re.compile(r'^([ \t]*)pangulu_numeric\s*\(', re.M)
    c = call.search(text, m.end())
    if not c:
        raise SystemExit("call of pangulu_numeric not found inside pangulu_gstrf")
    text = text[:c.start()] + CALL + text[c.start():]
    text = text[:m.start()] + DECL + "\n" + text[m.start():]
    open(dst, "w").write(text)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
