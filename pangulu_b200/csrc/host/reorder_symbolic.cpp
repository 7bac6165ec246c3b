// reorder_symbolic.cpp -- fill-reducing reordering and symmetric-pattern symbolic factorisation.
//
// Behaviour follows the reference so that the nonzero STRUCTURE the numeric kernels work on is
// bit-identical to PanguLU's for the same input and the same METIS library:
//   * missing diagonal entries are inserted as 1e-8       (src/pangulu_reordering.c:715-796)
//   * METIS_NodeND on the graph of A+A^T, adjacency ascending, no self loops, default options;
//     metis_perm = METIS' iperm                           (src/pangulu_reordering.c:934-1088)
//   * B = P A P^T with new index metis_perm[old], columns sorted (…:1090-1128, 1246-1247)
//   * symbolic: L pattern of the symmetrised matrix, parent = first off-diagonal row, children merged
//     into the parent column (src/pangulu_symbolic.c:132-247); symbolic_nnz = 2 nnz(L) - n (:242)
// The implementation is different (sorted-list merges along the elimination tree instead of linked
// prune lists) and returns each column of L with ascending row indices.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "host_internal.h"

#ifdef PANGULU_SM100_HAVE_METIS
extern "C" int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options, int64_t* perm, int64_t* iperm);
#endif

namespace pdrv {

double wall_seconds()
{
    using namespace std::chrono;
    return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

bool metis_available()
{
#ifdef PANGULU_SM100_HAVE_METIS
    return true;
#else
    return false;
#endif
}

// sorted CSC with every diagonal entry present (inserted as 1e-8 where missing)
static void with_full_diagonal(const Csc& A, Csc& out)
{
    const eidx n = A.n;
    out.n = n;
    out.ptr.assign((size_t)n + 1, 0);
    std::vector<char> has(n, 0);
    for (eidx c = 0; c < n; ++c)
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e)
            if (A.idx[e] == c) { has[c] = 1; break; }
    for (eidx c = 0; c < n; ++c) out.ptr[c + 1] = out.ptr[c] + (A.ptr[c + 1] - A.ptr[c]) + (has[c] ? 0 : 1);
    out.idx.resize(out.ptr[n]);
    out.val.resize(out.ptr[n]);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)n; ++c)
    {
        eptr o = out.ptr[c];
        bool placed = has[c];
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e)
        {
            if (!placed && A.idx[e] > (eidx)c)
            {
                out.idx[o] = (eidx)c;
                out.val[o] = t_from_real(1e-8);
                ++o;
                placed = true;
            }
            out.idx[o] = A.idx[e];
            out.val[o] = A.val[e];
            ++o;
        }
        if (!placed)
        {
            out.idx[o] = (eidx)c;
            out.val[o] = t_from_real(1e-8);
        }
    }
}

static void sort_columns(Csc& A)
{
#pragma omp parallel
    {
        std::vector<std::pair<eidx, T>> tmp;
#pragma omp for schedule(dynamic, 256)
        for (int64_t c = 0; c < (int64_t)A.n; ++c)
        {
            const eptr b = A.ptr[c], e = A.ptr[c + 1];
            bool sorted = true;
            for (eptr k = b + 1; k < e; ++k)
                if (A.idx[k - 1] > A.idx[k]) { sorted = false; break; }
            if (sorted) continue;
            tmp.resize(e - b);
            for (eptr k = b; k < e; ++k) tmp[k - b] = {A.idx[k], A.val[k]};
            std::sort(tmp.begin(), tmp.end(), [](const std::pair<eidx, T>& x, const std::pair<eidx, T>& y) { return x.first < y.first; });
            for (eptr k = b; k < e; ++k) { A.idx[k] = tmp[k - b].first; A.val[k] = tmp[k - b].second; }
        }
    }
}

// pattern of A^T (CSC of the transpose), rows ascending
static void transpose_pattern(const Csc& A, std::vector<eptr>& tptr, std::vector<eidx>& tidx)
{
    const eidx n = A.n;
    tptr.assign((size_t)n + 1, 0);
    for (eptr e = 0; e < A.ptr[n]; ++e) tptr[A.idx[e] + 1]++;
    for (eidx i = 0; i < n; ++i) tptr[i + 1] += tptr[i];
    tidx.resize(A.ptr[n]);
    std::vector<eptr> cur(tptr.begin(), tptr.end() - 1);
    for (eidx c = 0; c < n; ++c)
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e) tidx[cur[A.idx[e]]++] = c;
}

void reorder_matrix(const Csc& A_in, int ordering, Reordering& R, Csc& out)
{
    const eidx n = A_in.n;
    Csc A;
    {
        Csc tmp = A_in;
        sort_columns(tmp);
        with_full_diagonal(tmp, A);
    }
    R.row_perm.resize(n);
    R.col_perm.resize(n);
    std::iota(R.row_perm.begin(), R.row_perm.end(), 0);
    std::iota(R.col_perm.begin(), R.col_perm.end(), 0);
    R.row_scale.assign(n, t_from_real(1.0));
    R.col_scale.assign(n, t_from_real(1.0));
    R.metis_perm.resize(n);
    std::iota(R.metis_perm.begin(), R.metis_perm.end(), 0);

    if (ordering == ORDER_METIS)
    {
#ifdef PANGULU_SM100_HAVE_METIS
        std::vector<eptr> tptr;
        std::vector<eidx> tidx;
        transpose_pattern(A, tptr, tidx);
        std::vector<int64_t> xadj((size_t)n + 1), adjncy;
        adjncy.reserve(A.ptr[n] * 2);
        xadj[0] = 0;
        for (eidx v = 0; v < n; ++v)
        {
            eptr a = tptr[v], ae = tptr[v + 1], b = A.ptr[v], be = A.ptr[v + 1];
            while (a < ae || b < be)
            {
                eidx x;
                if (b >= be || (a < ae && tidx[a] < A.idx[b])) x = tidx[a++];
                else if (a >= ae || A.idx[b] < tidx[a]) x = A.idx[b++];
                else { x = tidx[a]; ++a; ++b; }
                if (x != v) adjncy.push_back((int64_t)x);
            }
            xadj[v + 1] = (int64_t)adjncy.size();
        }
        int64_t nv = n;
        std::vector<int64_t> perm(n), iperm(n);
        int rc = METIS_NodeND(&nv, xadj.data(), adjncy.data(), nullptr, nullptr, perm.data(), iperm.data());
        if (rc != 1)
        {
            fprintf(stderr, "[PanguLU ERROR] METIS_NodeND failed (%d)\n", rc);
            exit(EXIT_FAILURE);
        }
        for (eidx i = 0; i < n; ++i) R.metis_perm[i] = (eidx)iperm[i];
#else
        fprintf(stderr, "[PanguLU ERROR] driver built without METIS; use ordering 0 with a pre-permuted matrix\n");
        exit(EXIT_FAILURE);
#endif
    }

    // out(:, metis_perm[c]) = A(:, c) with rows metis_perm[r]
    out.n = n;
    out.ptr.assign((size_t)n + 1, 0);
    for (eidx c = 0; c < n; ++c) out.ptr[R.metis_perm[c] + 1] = A.ptr[c + 1] - A.ptr[c];
    for (eidx c = 0; c < n; ++c) out.ptr[c + 1] += out.ptr[c];
    out.idx.resize(A.ptr[n]);
    out.val.resize(A.ptr[n]);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < (int64_t)n; ++c)
    {
        eptr o = out.ptr[R.metis_perm[c]];
        for (eptr e = A.ptr[c]; e < A.ptr[c + 1]; ++e, ++o)
        {
            out.idx[o] = R.metis_perm[A.idx[e]];
            out.val[o] = A.val[e];
        }
    }
    sort_columns(out);
}

void symbolic_factor(const Csc& A, Symbolic& S)
{
    const eidx n = A.n;
    std::vector<eptr> tptr;
    std::vector<eidx> tidx;
    transpose_pattern(A, tptr, tidx);

    S.ptr.assign((size_t)n + 1, 0);
    S.idx.clear();
    S.idx.reserve((size_t)A.ptr[n] * 4);
    std::vector<eidx> first_child(n, (eidx)~0u), next_sibling(n, (eidx)~0u);
    std::vector<eidx> mark(n, (eidx)~0u);
    std::vector<eidx> acol, work;

    for (eidx i = 0; i < n; ++i)
    {
        // rows >= i of column i of A + A^T (ascending, diagonal first)
        acol.clear();
        {
            eptr a = tptr[i], ae = tptr[i + 1], b = A.ptr[i], be = A.ptr[i + 1];
            while (a < ae && tidx[a] < i) ++a;
            while (b < be && A.idx[b] < i) ++b;
            if ((a >= ae || tidx[a] != i) && (b >= be || A.idx[b] != i)) acol.push_back(i);
            while (a < ae || b < be)
            {
                eidx x;
                if (b >= be || (a < ae && tidx[a] < A.idx[b])) x = tidx[a++];
                else if (a >= ae || A.idx[b] < tidx[a]) x = A.idx[b++];
                else { x = tidx[a]; ++a; ++b; }
                acol.push_back(x);
            }
        }
        const size_t base = S.idx.size();
        const eidx c0 = first_child[i];
        if (c0 == (eidx)~0u)
        {
            S.idx.insert(S.idx.end(), acol.begin(), acol.end());
        }
        else if (next_sibling[c0] == (eidx)~0u)
        {
            // one child: its column is [c0, i, ...]; merge the tail starting at i with acol
            const eidx* cp = S.idx.data() + S.ptr[c0] + 1;
            const eidx* ce = S.idx.data() + S.ptr[c0 + 1];
            const size_t need = acol.size() + (size_t)(ce - cp);
            const size_t cp_off = (size_t)(cp - S.idx.data()), ce_off = (size_t)(ce - S.idx.data());
            S.idx.resize(base + need);
            cp = S.idx.data() + cp_off;
            ce = S.idx.data() + ce_off;
            eidx* o = S.idx.data() + base;
            size_t a = 0;
            while (a < acol.size() || cp < ce)
            {
                if (cp >= ce || (a < acol.size() && acol[a] < *cp)) *o++ = acol[a++];
                else if (a >= acol.size() || *cp < acol[a]) *o++ = *cp++;
                else { *o++ = *cp++; ++a; }
            }
            S.idx.resize((size_t)(o - S.idx.data()));
        }
        else
        {
            work.clear();
            for (eidx x : acol) { mark[x] = i; work.push_back(x); }
            for (eidx c = c0; c != (eidx)~0u; c = next_sibling[c])
                for (eptr e = S.ptr[c] + 1; e < S.ptr[c + 1]; ++e)
                {
                    const eidx x = S.idx[e];
                    if (mark[x] != i) { mark[x] = i; work.push_back(x); }
                }
            std::sort(work.begin(), work.end());
            S.idx.insert(S.idx.end(), work.begin(), work.end());
        }
        S.ptr[i + 1] = S.idx.size();
        if (S.ptr[i + 1] - S.ptr[i] > 1)
        {
            const eidx parent = S.idx[S.ptr[i] + 1];
            next_sibling[i] = first_child[parent];
            first_child[parent] = i;
        }
    }
    S.idx.shrink_to_fit();
    S.symbolic_nnz = (eptr)S.idx.size() * 2 - n;
}

}  // namespace pdrv
