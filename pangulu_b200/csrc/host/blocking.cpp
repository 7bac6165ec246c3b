// blocking.cpp -- cut the symbolic L+U pattern into nb x nb sparse blocks stored in PanguLU's packed
// block layout (reference src/pangulu_storage.c:199-422, src/pangulu_utils.c:479-575,
// src/pangulu_conversion.c:241-350) and scatter the values of the reordered matrix into them.
//
// One contiguous pinned host buffer and one device buffer with IDENTICAL byte layout hold every block this
// rank keeps (owned blocks first, in the order of the reference's bins[0]; then the remote blocks that the
// rank's tasks read as operands); all value arrays sit in one region, so upload/download of all factors
// is a single copy of that region.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "host_internal.h"

namespace pdrv {

int64_t BlockStore::find_block(eidx brow, eidx bcol) const
{
    const eidx* b = bcol_row.data() + bcol_ptr[bcol];
    const eidx* e = bcol_row.data() + bcol_ptr[bcol + 1];
    const eidx* it = std::lower_bound(b, e, brow);
    if (it == e || *it != brow) return -1;
    return (int64_t)(it - bcol_row.data());
}

namespace {

struct Packed
{
    // sub-array pointers of one block
    T* val;
    iptr* ptr;
    iidx* idx;
    iptr* rmap;
    iptr* rptr;
    iidx* ridx;
};

inline size_t align16(size_t x) { return (x + 15u) & ~(size_t)15u; }

// Every block's value array lives in the VALUE region of the buffer (16-byte aligned so kernels can use
// 128-bit loads) and its index arrays in the INDEX region behind it: uploading / downloading the factors
// moves the value region only.  (The kernels see blocks only through the slot pointers, so the in-slot
// meaning of every array is exactly the reference's; only the placement differs from the interleaved
// wire layout of SURVEY Appendix A, which this driver does not need because every rank derives the
// patterns itself and ships values only.)
// (an empty block -- e.g. the strict-lower half of a diagonal-only diagonal block -- still gets its own 16 bytes: the
// platform identifies a block by its value pointer, which therefore must be unique)
inline size_t value_bytes(size_t nnz) { return align16(sizeof(T) * std::max<size_t>(nnz, 1)); }
inline size_t index_bytes(int nb, size_t nnz, bool mirror)
{
    size_t s = align16(sizeof(iptr) * (size_t)(nb + 1)) + align16(sizeof(iidx) * nnz);
    if (mirror) s += align16(sizeof(iptr) * nnz) + align16(sizeof(iptr) * (size_t)(nb + 1)) + align16(sizeof(iidx) * nnz);
    return s;
}
inline Packed carve(char* vbase, char* ibase, int nb, size_t nnz, bool mirror)
{
    Packed p;
    p.val = (T*)vbase;
    char* c = ibase;
    p.ptr = (iptr*)c;
    c += align16(sizeof(iptr) * (size_t)(nb + 1));
    p.idx = (iidx*)c;
    c += align16(sizeof(iidx) * nnz);
    if (mirror)
    {
        p.rmap = (iptr*)c;
        c += align16(sizeof(iptr) * nnz);
        p.rptr = (iptr*)c;
        c += align16(sizeof(iptr) * (size_t)(nb + 1));
        p.ridx = (iidx*)c;
    }
    else
    {
        p.rmap = nullptr;
        p.rptr = nullptr;
        p.ridx = nullptr;
    }
    return p;
}

// CSC pattern of lower block (I,J), I > J, from the element-level L pattern
void fill_lower_pattern(const Symbolic& S, eidx n, int nb, eidx I, eidx J, iptr* ptr, iidx* idx)
{
    const eidx c0 = J * (eidx)nb, r0 = I * (eidx)nb, r1 = r0 + (eidx)nb;
    iptr k = 0;
    ptr[0] = 0;
    for (int cl = 0; cl < nb; ++cl)
    {
        const eidx c = c0 + cl;
        if (c < n)
        {
            const eidx* b = S.idx.data() + S.ptr[c];
            const eidx* e = S.idx.data() + S.ptr[c + 1];
            const eidx* it = std::lower_bound(b, e, r0);
            for (; it < e && *it < r1; ++it) idx[k++] = (iidx)(*it - r0);
        }
        ptr[cl + 1] = k;
    }
}

// CSR mirror with value index (same result as pangulu_transpose_struct_with_valueidx_inblock)
void mirror_pattern(int nb, const iptr* ptr, const iidx* idx, iptr* rptr, iidx* ridx, iptr* rmap)
{
    const iptr nnz = ptr[nb];
    for (int i = 0; i <= nb; ++i) rptr[i] = 0;
    for (iptr e = 0; e < nnz; ++e) rptr[idx[e] + 1]++;
    for (int i = 0; i < nb; ++i) rptr[i + 1] += rptr[i];
    std::vector<iptr> cur(rptr, rptr + nb);
    for (int c = 0; c < nb; ++c)
        for (iptr e = ptr[c]; e < ptr[c + 1]; ++e)
        {
            const iptr o = cur[idx[e]]++;
            ridx[o] = (iidx)c;
            rmap[o] = e;
        }
}

}  // namespace

void build_blocks(const Csc& A, const Symbolic& S, int nb, int rank, int nproc, BlockStore& B)
{
    const eidx n = A.n;
    B.nb = nb;
    B.n = n;
    B.block_length = (n + nb - 1) / nb;
    B.rank = rank;
    B.nproc = nproc;
    {
        // p = largest divisor of nproc <= sqrt(nproc), q = nproc / p (src/pangulu.c:83-90)
        int p = 1;
        for (int d = 1; d * d <= nproc; ++d)
            if (nproc % d == 0) p = d;
        B.p = p;
        B.q = nproc / p;
    }
    const eidx bl = B.block_length;

    // ---- 1. nnz of every lower block and of the diagonal halves ------------------------------------
    std::vector<std::vector<std::pair<eidx, uint32_t>>> lower(bl);  // per block column: (brow, nnz)
    B.diag_nnz_lower.assign(bl, 0);
    B.diag_nnz_upper.assign(bl, 0);
#pragma omp parallel
    {
        std::vector<uint32_t> cnt(bl, 0);
        std::vector<eidx> touched;
#pragma omp for schedule(dynamic, 4)
        for (int64_t J = 0; J < (int64_t)bl; ++J)
        {
            touched.clear();
            const eidx c0 = (eidx)J * nb, c1 = std::min<eidx>(n, c0 + nb);
            for (eidx c = c0; c < c1; ++c)
            {
                const eidx* it = S.idx.data() + S.ptr[c];
                const eidx* e = S.idx.data() + S.ptr[c + 1];
                while (it < e)
                {
                    const eidx I = *it / nb;
                    const eidx lim = (I + 1) * nb;
                    const eidx* run = it;
                    while (run < e && *run < lim) ++run;
                    if (cnt[I] == 0) touched.push_back(I);
                    cnt[I] += (uint32_t)(run - it);
                    it = run;
                }
            }
            std::sort(touched.begin(), touched.end());
            for (eidx I : touched)
            {
                if (I == (eidx)J)
                {
                    const uint32_t ndiag = c1 - c0;               // every column carries its diagonal
                    B.diag_nnz_upper[J] = cnt[I];                 // diagonal + strict part
                    B.diag_nnz_lower[J] = cnt[I] - ndiag;
                }
                else
                    lower[J].push_back({I, cnt[I]});
                cnt[I] = 0;
            }
        }
    }

    // ---- 2. block-level CSC of all off-diagonal blocks (upper = transposes of lower) ------------------
    B.bcol_ptr.assign((size_t)bl + 1, 0);
    for (eidx J = 0; J < bl; ++J)
    {
        B.bcol_ptr[J + 1] += lower[J].size();
        for (auto& pr : lower[J]) B.bcol_ptr[pr.first + 1] += 1;  // upper block (J, pr.first) lives in column pr.first
    }
    for (eidx J = 0; J < bl; ++J) B.bcol_ptr[J + 1] += B.bcol_ptr[J];
    const size_t nblk = B.bcol_ptr[bl];
    B.bcol_row.resize(nblk);
    B.bnnz.resize(nblk);
    {
        std::vector<eptr> cur(B.bcol_ptr.begin(), B.bcol_ptr.end() - 1);
        // upper blocks first (ascending brow because J ascends), then the lower ones
        for (eidx J = 0; J < bl; ++J)
            for (auto& pr : lower[J])
            {
                const eptr o = cur[pr.first]++;
                B.bcol_row[o] = J;
                B.bnnz[o] = pr.second;
            }
        for (eidx J = 0; J < bl; ++J)
            for (auto& pr : lower[J])
            {
                const eptr o = cur[J]++;
                B.bcol_row[o] = pr.first;
                B.bnnz[o] = pr.second;
            }
    }

    // ---- 3. which blocks this rank keeps, their order and byte offsets ------------------------------------
    const int my_prow = rank / B.q, my_pcol = rank % B.q;
    struct Keep { int64_t gid; eidx brow, bcol; int kind; /*0 offdiag,1 diag lower,2 diag upper*/ size_t voff, ioff; uint32_t nnz; };
    std::vector<Keep> keep;
    B.slot_of_block.assign(nblk, -1);
    B.slot_of_diag.assign(bl, -1);
    auto push = [&](int64_t gid, eidx br, eidx bc, int kind, uint32_t nnz) { keep.push_back({gid, br, bc, kind, 0, 0, nnz}); };
    // owned off-diagonal blocks, block-CSC order
    for (eidx J = 0; J < bl; ++J)
        for (eptr g = B.bcol_ptr[J]; g < B.bcol_ptr[J + 1]; ++g)
            if (B.owner(B.bcol_row[g], J) == rank) push((int64_t)g, B.bcol_row[g], J, 0, B.bnnz[g]);
    B.n_own_offdiag = keep.size();
    for (eidx K = 0; K < bl; ++K)
        if (B.owner(K, K) == rank)
        {
            push(-1, K, K, 1, B.diag_nnz_lower[K]);
            push(-1, K, K, 2, B.diag_nnz_upper[K]);
        }
    B.n_own_diag = (keep.size() - B.n_own_offdiag) / 2;
    const size_t n_owned_slots = keep.size();
    if (nproc > 1)
    {
        // remote operands: L blocks of my process row, U blocks of my process column, and the diagonal
        // halves that my panel blocks are solved against (same sets as the reference's "related" blocks,
        // src/pangulu_communication.c:1079-1090)
        for (eidx J = 0; J < bl; ++J)
            for (eptr g = B.bcol_ptr[J]; g < B.bcol_ptr[J + 1]; ++g)
            {
                const eidx I = B.bcol_row[g];
                if (B.owner(I, J) == rank) continue;
                const bool is_lower = I > J;
                if ((is_lower && (int)(I % B.p) == my_prow) || (!is_lower && (int)(J % B.q) == my_pcol)) push((int64_t)g, I, J, 0, B.bnnz[g]);
            }
        for (eidx K = 0; K < bl; ++K)
        {
            if (B.owner(K, K) == rank) continue;
            // both halves travel together (TSTRF needs the upper half when K%q == my_pcol, GESSM the lower
            // half when K%p == my_prow); keeping the pair keeps related_block valid
            if ((int)(K % B.q) == my_pcol || (int)(K % B.p) == my_prow)
            {
                push(-1, K, K, 1, B.diag_nnz_lower[K]);
                push(-1, K, K, 2, B.diag_nnz_upper[K]);
            }
        }
    }
    size_t voff = 0, ioff = 0;
    B.nnz_total = 0;
    B.own_value_bytes = 0;
    for (size_t s = 0; s < keep.size(); ++s)
    {
        if (s == n_owned_slots) B.own_value_bytes = voff;
        keep[s].voff = voff;
        keep[s].ioff = ioff;
        const bool mirror = keep[s].kind == 0 && keep[s].brow > keep[s].bcol;
        voff += value_bytes(keep[s].nnz);
        ioff += index_bytes(nb, keep[s].nnz, mirror);
        if (s < n_owned_slots) B.nnz_total += keep[s].nnz;
    }
    if (keep.size() == n_owned_slots) B.own_value_bytes = voff;
    B.value_bytes = voff;
    B.buf_bytes = voff + ioff;

    if (B.host_only)
    {
        // structure-only mode for CPU-side tests of the host logic: no device is touched, device twin
        // pointers are offsets from a null base and must never be dereferenced (gstrf refuses to run)
        B.h_buf = (char*)aligned_alloc(64, (std::max<size_t>(B.buf_bytes, 64) + 63) & ~(size_t)63);
        B.d_buf = nullptr;
    }
    else
    {
        PSM(malloc_pinned)((void**)&B.h_buf, std::max<size_t>(B.buf_bytes, 64));
        PSM(malloc)((void**)&B.d_buf, std::max<size_t>(B.buf_bytes, 64));
    }

    // ---- 4. slots ----------------------------------------------------------------------------------------
    B.slots.assign(keep.size(), pangulu_storage_slot_t());
    for (size_t s = 0; s < keep.size(); ++s)
    {
        const Keep& k = keep[s];
        pangulu_storage_slot_t& sl = B.slots[s];
        memset(&sl, 0, sizeof sl);
        const bool mirror = k.kind == 0 && k.brow > k.bcol;
        Packed h = carve(B.h_buf + k.voff, B.h_buf + B.value_bytes + k.ioff, nb, k.nnz, mirror);
        Packed d = carve(B.d_buf + k.voff, B.d_buf + B.value_bytes + k.ioff, nb, k.nnz, mirror);
        sl.brow_pos = k.brow;
        sl.bcol_pos = k.bcol;
        sl.value = h.val;
        sl.columnpointer = h.ptr;
        sl.rowindex = h.idx;
        sl.rowpointer = h.rptr;
        sl.columnindex = h.ridx;
        sl.idx_of_csc_value_for_csr = h.rmap;
        sl.data_status = PANGULU_DATA_PREPARING;
        sl.is_upper = k.kind == 2 ? 1 : (k.kind == 0 && k.brow < k.bcol ? 1 : 0);
        sl.bin_id = s < n_owned_slots ? 0 : 1;
        sl.slot_idx = (pangulu_int32_t)s;
        sl.d_value = d.val;
        if (k.kind == 2)
        {
            // upper half of a diagonal block: the device twins of the CSR arrays are the d_row* fields
            // (src/pangulu_storage.c:346-352); the host keeps them in columnpointer/rowindex
            sl.d_rowpointer = d.ptr;
            sl.d_columnindex = d.idx;
        }
        else
        {
            sl.d_columnpointer = d.ptr;
            sl.d_rowindex = d.idx;
            sl.d_rowpointer = d.rptr;
            sl.d_columnindex = d.ridx;
            sl.d_idx_of_csc_value_for_csr = d.rmap;
        }
        if (k.kind == 0) B.slot_of_block[k.gid] = (int64_t)s;
        else if (k.kind == 1) B.slot_of_diag[k.brow] = (int64_t)s;
    }
    for (size_t s = 0; s < keep.size(); ++s)
        if (keep[s].kind == 1)
        {
            B.slots[s].related_block = &B.slots[s + 1];
            B.slots[s + 1].related_block = &B.slots[s];
        }

    // ---- 5. patterns ---------------------------------------------------------------------------------------
    memset(B.h_buf, 0, B.buf_bytes);
#pragma omp parallel
    {
        std::vector<iptr> tptr((size_t)nb + 1), tmap;
        std::vector<iidx> tidx, tridx;
        std::vector<iptr> trptr((size_t)nb + 1);
#pragma omp for schedule(dynamic, 16)
        for (int64_t s = 0; s < (int64_t)keep.size(); ++s)
        {
            const Keep& k = keep[s];
            pangulu_storage_slot_t& sl = B.slots[s];
            if (k.kind == 0 && k.brow > k.bcol)
            {
                fill_lower_pattern(S, n, nb, k.brow, k.bcol, sl.columnpointer, sl.rowindex);
                mirror_pattern(nb, sl.columnpointer, sl.rowindex, sl.rowpointer, sl.columnindex, sl.idx_of_csc_value_for_csr);
            }
            else if (k.kind == 0)
            {
                // upper block (I,J), I < J: CSC pattern = CSR pattern of lower block (J,I)
                tidx.resize(std::max<size_t>(k.nnz, 1));
                tridx.resize(std::max<size_t>(k.nnz, 1));
                tmap.resize(std::max<size_t>(k.nnz, 1));
                fill_lower_pattern(S, n, nb, k.bcol, k.brow, tptr.data(), tidx.data());
                mirror_pattern(nb, tptr.data(), tidx.data(), sl.columnpointer, sl.rowindex, tmap.data());
            }
            else if (k.kind == 1)
            {
                // strict lower part of the diagonal block, CSC
                const eidx c0 = k.bcol * (eidx)nb, r1 = c0 + (eidx)nb;
                iptr o = 0;
                sl.columnpointer[0] = 0;
                for (int cl = 0; cl < nb; ++cl)
                {
                    const eidx c = c0 + cl;
                    if (c < n)
                    {
                        const eidx* it = S.idx.data() + S.ptr[c] + 1;  // skip the diagonal
                        const eidx* e = S.idx.data() + S.ptr[c + 1];
                        for (; it < e && *it < r1; ++it) sl.rowindex[o++] = (iidx)(*it - c0);
                    }
                    sl.columnpointer[cl + 1] = o;
                }
            }
            else
            {
                // upper part incl. diagonal, CSR, diagonal first: row c = {c} U lower column c (symmetric pattern)
                const eidx c0 = k.bcol * (eidx)nb, r1 = c0 + (eidx)nb;
                iptr o = 0;
                sl.columnpointer[0] = 0;
                for (int cl = 0; cl < nb; ++cl)
                {
                    const eidx c = c0 + cl;
                    if (c < n)
                    {
                        const eidx* it = S.idx.data() + S.ptr[c];
                        const eidx* e = S.idx.data() + S.ptr[c + 1];
                        for (; it < e && *it < r1; ++it) sl.rowindex[o++] = (iidx)(*it - c0);
                    }
                    sl.columnpointer[cl + 1] = o;
                }
            }
        }
    }

    // ---- 6. values of the reordered matrix --------------------------------------------------------------------
    auto find_in = [](const iidx* idx, iptr b, iptr e, iidx key) -> int64_t {
        const iidx* it = std::lower_bound(idx + b, idx + e, key);
        if (it == idx + e || *it != key) return -1;
        return (int64_t)(it - idx);
    };
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : bad)
    for (int64_t c = 0; c < (int64_t)n; ++c)
    {
        const eidx J = (eidx)c / nb;
        const iidx cl = (iidx)((eidx)c - J * nb);
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e)
        {
            const eidx r = A.idx[e];
            const eidx I = r / nb;
            const iidx rl = (iidx)(r - I * nb);
            pangulu_storage_slot_t* sl = nullptr;
            int64_t pos = -1;
            if (I == J)
            {
                const int64_t ds = B.slot_of_diag[J];
                if (ds < 0) continue;
                if (r >= (eidx)c)
                {
                    // strict lower half is CSC; the diagonal entry is the first entry of upper row c
                    if (r == (eidx)c) { sl = &B.slots[ds + 1]; pos = sl->columnpointer[cl]; }
                    else { sl = &B.slots[ds]; pos = find_in(sl->rowindex, sl->columnpointer[cl], sl->columnpointer[cl + 1], rl); }
                }
                else
                {
                    sl = &B.slots[ds + 1];  // upper CSR: row rl, column cl
                    pos = find_in(sl->rowindex, sl->columnpointer[rl], sl->columnpointer[rl + 1], cl);
                }
            }
            else
            {
                const int64_t gid = B.find_block(I, J);
                if (gid < 0) { bad++; continue; }
                const int64_t s = B.slot_of_block[gid];
                if (s < 0) continue;
                sl = &B.slots[s];
                pos = find_in(sl->rowindex, sl->columnpointer[cl], sl->columnpointer[cl + 1], rl);
            }
            if (pos < 0) { bad++; continue; }
            sl->value[pos] = A.val[e];
        }
    }
    if (bad)
    {
        fprintf(stderr, "[PanguLU ERROR] %d matrix entries fall outside the symbolic pattern\n", bad);
        exit(EXIT_FAILURE);
    }

    // ---- 7. device mirror: one copy, identical layout --------------------------------------------------------
    if (B.buf_bytes && !B.host_only) PSM(memcpy)(B.d_buf, B.h_buf, B.buf_bytes, 0);
}

void free_blocks(BlockStore& B)
{
    if (B.host_only) { free(B.h_buf); B.h_buf = nullptr; return; }
    if (B.d_buf) PSM(free)(B.d_buf);
    if (B.h_buf) pangulu_sm100_free_pinned(B.h_buf);
    B.d_buf = nullptr;
    B.h_buf = nullptr;
}

}  // namespace pdrv
