// schedule.cpp -- static task schedule of the block LU and the flop/byte models.
//
// The reference discovers ready tasks at run time with per-block dependency counters, a min-heap and a
// per-destination SSSSM aggregator (src/pangulu_numeric.c, src/pangulu_task.c, counters from
// src/pangulu_preprocessing.c:130-315).  The dependency structure itself is static -- it only depends on
// the block pattern -- so this driver resolves it once, on the host, into WAVES:
//
//   depth(k,k) = 1 + max over updates S(k,k,l) of max(depth(k,l), depth(l,k))          (0 without updates)
//   depth(i,k) = 1 + max(depth(k,k), max over updates S(i,k,l) of max(depth(i,l), depth(l,k)))
//
// Wave w holds, in this order: every SSSSM update whose destination has depth w (all updates of a
// destination are applied together, right before the destination is finalised -- the aggregator's lazy
// flush, src/pangulu_task.c:39-91, made static), then GETRF of the diagonal blocks of depth w, then
// TSTRF/GESSM of the off-diagonal blocks of depth w.  Tasks inside a wave are independent except
// same-destination SSSSMs, which the platform combines per destination.
// Owner computes (src/pangulu_numeric.c:51,70,93,148): a rank lists only tasks whose destination it owns.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "host_internal.h"

namespace pdrv {

namespace {
struct Proto
{
    int32_t wave;
    int8_t cls;  // 0 SSSSM, 1 GETRF, 2 TSTRF/GESSM
    int16_t kernel;
    eidx row, col, level;
    int64_t dst, op1, op2;  // slot indices
    int32_t grp = 0;        // GETRF: group of the diagonal block
};
}  // namespace

// Independent parts of diagonal blocks.  The regular nb grid cuts through the elimination tree: a diagonal block often
// holds the tail of one subtree and the head of the next, which are structurally independent (the block is block
// diagonal over them) -- but as ONE GETRF task it chains the two subtrees, and through them every block row to its
// predecessor (7563 dependent waves for 3907 block rows at the 100^3 benchmark size).  Here every diagonal block is cut
// at the index positions where its filled pattern is block diagonal, adjacent pieces are merged while they agree on
// "receives updates from earlier block columns" (at most DIAG_MAX_GROUPS groups), and GETRF, the dependency counters
// of the updates and the waits of the panel solves work per GROUP.  Storage and kernels' arithmetic are unchanged.
void diag_groups(const BlockStore& B, const Symbolic& Sy, Schedule& S)
{
    // Works on the element-level symbolic pattern (column c of L: rows >= c, ascending), which every rank computes
    // itself: the result does not depend on which blocks a rank holds.
    const eidx bl = B.block_length, n = B.n;
    const int nb = B.nb;
    const size_t nblk = B.bcol_ptr[bl];
    S.group_bounds.assign((size_t)bl * (DIAG_MAX_GROUPS + 1), (uint16_t)nb);
    S.group_count.assign(bl, 1);
    S.colgroups.assign(nblk, 1);
    S.rowgroups.assign(nblk, 1);
    for (eidx k = 0; k < bl; ++k) S.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1)] = 0;
    if (!S.split_diag || Sy.idx.empty()) return;
    std::fill(S.colgroups.begin(), S.colgroups.end(), 0);
    std::fill(S.rowgroups.begin(), S.rowgroups.end(), 0);
    // 1. components: split points p with no entry (r, c), r >= p > c, inside the diagonal block
    std::vector<std::vector<uint16_t>> comp_start(bl);
    std::vector<uint8_t> comp_of((size_t)bl * nb, 0);   // index -> component of its diagonal block (capped at 255)
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t k = 0; k < (int64_t)bl; ++k)
    {
        std::vector<uint16_t>& st = comp_start[k];
        st.push_back(0);
        const eidx c0 = (eidx)k * nb, r1 = c0 + nb;
        int reach = 0;  // largest local row index touched by the columns seen so far
        for (int cl = 0; cl < nb; ++cl)
        {
            if (cl > 0 && reach < cl && st.size() < 255) st.push_back((uint16_t)cl);
            reach = std::max(reach, cl);
            const eidx c = c0 + cl;
            if (c < n)
            {
                const eidx* b = Sy.idx.data() + Sy.ptr[c];
                const eidx* e = Sy.idx.data() + Sy.ptr[c + 1];
                const eidx* it = std::lower_bound(b, e, r1);   // first row beyond the diagonal block
                if (it > b) reach = std::max(reach, (int)(*(it - 1) - c0));
            }
            comp_of[(size_t)k * nb + cl] = (uint8_t)(st.size() - 1);
        }
    }
    // 2. components that receive updates from earlier block columns = components holding a row of a strictly-lower
    //    block of their block row
    std::vector<std::vector<char>> ext(bl);
    for (eidx k = 0; k < bl; ++k) ext[k].assign(comp_start[k].size(), 0);
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t c = 0; c < (int64_t)n; ++c)
    {
        const eidx r1 = ((eidx)c / nb + 1) * nb;
        const eidx* b = Sy.idx.data() + Sy.ptr[c];
        const eidx* e = Sy.idx.data() + Sy.ptr[c + 1];
        for (const eidx* it = std::lower_bound(b, e, r1); it < e; ++it) ext[*it / nb][comp_of[*it]] = 1;  // (benign race: all writers store 1)
    }
    // 3. groups: runs of adjacent components with equal `ext`, at most DIAG_MAX_GROUPS
    std::vector<uint8_t> group_of((size_t)bl * nb, 0);  // index -> group of its diagonal block
    for (eidx k = 0; k < bl; ++k)
    {
        const std::vector<uint16_t>& st = comp_start[k];
        uint16_t* bd = &S.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1)];
        int ng = 1;
        bd[0] = 0;
        for (size_t c = 1; c < st.size(); ++c)
            if (ext[k][c] != ext[k][c - 1] && ng < DIAG_MAX_GROUPS) bd[ng++] = st[c];
        for (int g = ng; g <= DIAG_MAX_GROUPS; ++g) bd[g] = (uint16_t)nb;
        S.group_count[k] = (uint8_t)ng;
        for (int g = 0; g < ng; ++g)
            for (int c = bd[g]; c < bd[g + 1]; ++c) group_of[(size_t)k * nb + c] = (uint8_t)g;
    }
    // 4. which groups a strictly-lower block (I,J) touches: by its columns (groups of J) and by its rows (groups of I);
    //    the upper block (J,I) has the transposed pattern
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t J = 0; J < (int64_t)bl; ++J)
    {
        const eidx c0 = (eidx)J * nb, c1 = std::min<eidx>(n, c0 + nb);
        for (eidx c = c0; c < c1; ++c)
        {
            const uint8_t gc = (uint8_t)(1u << group_of[c]);
            const eidx* b = Sy.idx.data() + Sy.ptr[c];
            const eidx* e = Sy.idx.data() + Sy.ptr[c + 1];
            const eidx* it = std::lower_bound(b, e, c0 + (eidx)nb);
            int64_t gid = -1;
            eidx cur = (eidx)-1;
            for (; it < e; ++it)
            {
                const eidx I = *it / nb;
                if (I != cur) { cur = I; gid = B.find_block(I, (eidx)J); }
                if (gid < 0) continue;
                S.colgroups[gid] |= gc;
                S.rowgroups[gid] |= (uint8_t)(1u << group_of[*it]);
            }
        }
    }
    for (eidx J = 0; J < bl; ++J)
        for (eptr g = B.bcol_ptr[J]; g < B.bcol_ptr[J + 1]; ++g)
        {
            const eidx I = B.bcol_row[g];
            if (I <= J) continue;
            const int64_t gu = B.find_block(J, I);  // upper block (J, I): rows in block J, columns in block I
            if (gu >= 0) { S.rowgroups[gu] = S.colgroups[g]; S.colgroups[gu] = S.rowgroups[g]; }
        }
}

void build_schedule(BlockStore& B, Schedule& S)
{
    const eidx bl = B.block_length;
    const size_t nblk = B.bcol_ptr[bl];
    if (S.group_count.size() != bl)
    {
        // diag_groups() was not run: every diagonal block is one group
        S.group_bounds.assign((size_t)bl * (DIAG_MAX_GROUPS + 1), (uint16_t)B.nb);
        for (eidx k = 0; k < bl; ++k) S.group_bounds[(size_t)k * (DIAG_MAX_GROUPS + 1)] = 0;
        S.group_count.assign(bl, 1);
        S.colgroups.assign(nblk, 1);
        S.rowgroups.assign(nblk, 1);
    }
    constexpr int MG = DIAG_MAX_GROUPS;
    S.depth_offdiag.assign(nblk, 0);
    S.depth_diag.assign((size_t)bl * MG, 0);
    std::vector<int32_t> maxdep_off(nblk, -1), maxdep_diag((size_t)bl * MG, -1);
    std::vector<Proto> protos;
    std::vector<eidx> R;
    std::vector<int64_t> gid_l, gid_u;

    auto slot_off = [&](int64_t gid) { return B.slot_of_block[gid]; };

    for (eidx k = 0; k < bl; ++k)
    {
        // R = block rows below the diagonal in block column k (= block columns right of it in block row k)
        R.clear();
        gid_l.clear();
        gid_u.clear();
        for (eptr g = B.bcol_ptr[k]; g < B.bcol_ptr[k + 1]; ++g)
            if (B.bcol_row[g] > k) { R.push_back(B.bcol_row[g]); gid_l.push_back((int64_t)g); }
        for (eidx j : R) gid_u.push_back(B.find_block(k, j));

        const int ngk = S.group_count[k];
        for (int g = 0; g < ngk; ++g)
        {
            const int32_t dk = maxdep_diag[(size_t)k * MG + g] + 1;
            S.depth_diag[(size_t)k * MG + g] = dk;
            if (B.owner(k, k) == B.rank)
            {
                Proto t{dk, 1, PANGULU_TASK_GETRF, k, k, k, B.slot_of_diag[k], -1, -1};
                t.grp = g;
                protos.push_back(t);
                S.n_getrf++;
            }
        }
        // depth of the diagonal groups a panel block is solved against
        auto diag_depth = [&](uint8_t mask) {
            int32_t d = 0;
            for (int g = 0; g < ngk; ++g)
                if ((mask >> g) & 1u) d = std::max(d, S.depth_diag[(size_t)k * MG + g]);
            return d;
        };
        for (size_t a = 0; a < R.size(); ++a)
        {
            const int32_t dl = 1 + std::max(diag_depth(S.colgroups[gid_l[a]]), maxdep_off[gid_l[a]]);
            const int32_t du = 1 + std::max(diag_depth(S.rowgroups[gid_u[a]]), maxdep_off[gid_u[a]]);
            S.depth_offdiag[gid_l[a]] = dl;
            S.depth_offdiag[gid_u[a]] = du;
            if (B.owner(R[a], k) == B.rank)
            {
                // op1 = the diagonal block (lower-half slot; the platform picks the half it needs)
                Proto t{dl, 2, PANGULU_TASK_TSTRF, R[a], k, k, slot_off(gid_l[a]), B.slot_of_diag[k], -1};
                protos.push_back(t);
                S.n_tstrf++;
            }
            if (B.owner(k, R[a]) == B.rank)
            {
                Proto t{du, 2, PANGULU_TASK_GESSM, k, R[a], k, slot_off(gid_u[a]), B.slot_of_diag[k], -1};
                protos.push_back(t);
                S.n_gessm++;
            }
        }
        // updates S(i,j,k), i,j in R, destination block (i,j) must exist
        for (size_t b = 0; b < R.size(); ++b)
        {
            const eidx j = R[b];
            const int32_t du = S.depth_offdiag[gid_u[b]];
            // walk block column j against R
            eptr g = B.bcol_ptr[j];
            const eptr ge = B.bcol_ptr[j + 1];
            for (size_t a = 0; a < R.size(); ++a)
            {
                const eidx i = R[a];
                const int32_t m = std::max(S.depth_offdiag[gid_l[a]], du);
                if (i == j)
                {
                    // the update touches the groups of block i in which L(i,k) has rows
                    for (int g = 0; g < S.group_count[i]; ++g)
                        if ((S.rowgroups[gid_l[a]] >> g) & 1u) maxdep_diag[(size_t)i * MG + g] = std::max(maxdep_diag[(size_t)i * MG + g], m);
                    if (B.owner(i, i) == B.rank)
                    {
                        Proto t{m + 1, 0, PANGULU_TASK_SSSSM, i, i, k, B.slot_of_diag[i], slot_off(gid_l[a]), slot_off(gid_u[b])};
                        protos.push_back(t);
                        S.n_ssssm++;
                    }
                    continue;
                }
                while (g < ge && B.bcol_row[g] < i) ++g;
                if (g < ge && B.bcol_row[g] == i)
                {
                    maxdep_off[g] = std::max(maxdep_off[g], m);
                    if (B.owner(i, j) == B.rank)
                    {
                        Proto t{m + 1, 0, PANGULU_TASK_SSSSM, i, j, k, slot_off((int64_t)g), slot_off(gid_l[a]), slot_off(gid_u[b])};
                        protos.push_back(t);
                        S.n_ssssm++;
                    }
                }
            }
        }
    }
    // wave of an update:
    //   policy 0 (lazy, left-looking)  = depth of its destination (known only after all levels were seen): a
    //            destination is read and written once, but its updates pile up on the critical path;
    //   policy 1 (eager, right-looking) = 1 + max(depth of its operands), already stored: the update runs as soon
    //            as both panel blocks are final -- wide waves, and only the last update precedes the panel op.
    //   policy 2 (windowed) = max(eager wave, depth of the destination - lookahead): updates with a lot of slack are
    //            gathered `lookahead` waves before their destination is finalised (one multi-pair pass, off the
    //            critical path); the few that become ready inside the window run as soon as they can.  Inside a
    //            wave the updates are ordered by slack, most urgent first (the persistent executor consumes the
    //            list in order).
    int32_t nwave = 0;
    std::vector<int32_t> slack(protos.size(), 0);
    for (size_t i = 0; i < protos.size(); ++i)
    {
        Proto& t = protos[i];
        if (t.cls == 0)
        {
            int32_t dd;
            if (t.row == t.col)
            {
                // earliest finalisation among the groups this update touches
                const uint8_t mask = S.rowgroups[B.find_block(t.row, t.level)];
                dd = INT32_MAX;
                for (int g = 0; g < S.group_count[t.row]; ++g)
                    if ((mask >> g) & 1u) dd = std::min(dd, S.depth_diag[(size_t)t.row * MG + g]);
                if (dd == INT32_MAX) dd = S.depth_diag[(size_t)t.row * MG];
            }
            else
                dd = S.depth_offdiag[B.find_block(t.row, t.col)];
            if (S.policy == 0) t.wave = dd;
            else if (S.policy == 2) t.wave = std::max(t.wave, dd - S.lookahead);
            slack[i] = dd - t.wave;
        }
        nwave = std::max(nwave, t.wave + 1);
    }
    if (S.policy == 2)
    {
        // stable order by (wave, class, slack): done through an index sort, then the bucket pass below keeps it
        std::vector<size_t> ord(protos.size());
        for (size_t i = 0; i < ord.size(); ++i) ord[i] = i;
        std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) {
            if (protos[a].wave != protos[b].wave) return protos[a].wave < protos[b].wave;
            if (protos[a].cls != protos[b].cls) return protos[a].cls < protos[b].cls;
            return slack[a] < slack[b];
        });
        std::vector<Proto> sorted;
        sorted.reserve(protos.size());
        for (size_t i : ord) sorted.push_back(protos[i]);
        protos.swap(sorted);
    }
    // bucket by (wave, class), stable
    std::vector<uint64_t> cnt((size_t)nwave * 3 + 1, 0);
    for (auto& t : protos) cnt[(size_t)t.wave * 3 + t.cls + 1]++;
    for (size_t i = 1; i < cnt.size(); ++i) cnt[i] += cnt[i - 1];
    S.wave_ptr.resize((size_t)nwave + 1);
    for (int32_t w = 0; w <= nwave; ++w) S.wave_ptr[w] = cnt[(size_t)w * 3];
    S.tasks.resize(protos.size());
    std::vector<uint64_t> cur(cnt.begin(), cnt.end() - 1);
    for (auto& t : protos)
    {
        pangulu_task_t& o = S.tasks[cur[(size_t)t.wave * 3 + t.cls]++];
        memset(&o, 0, sizeof o);
        o.row = t.row;
        o.col = t.col;
        o.kernel_id = t.kernel;
        o.task_level = t.level;
        o.compare_flag = (int64_t)t.level | ((int64_t)t.grp << 32);  // GETRF: group of the diagonal block in bits 32..
        o.opdst = &B.slots[t.dst];
        o.op1 = t.op1 >= 0 ? &B.slots[t.op1] : nullptr;
        o.op2 = t.op2 >= 0 ? &B.slots[t.op2] : nullptr;
        if ((t.op1 < 0 && t.kernel != PANGULU_TASK_GETRF) || (t.op2 < 0 && t.kernel == PANGULU_TASK_SSSSM) || t.dst < 0)
        {
            fprintf(stderr, "[PanguLU ERROR] schedule: task (%u,%u,l=%u,k=%d) references a block this rank does not hold\n", t.row, t.col, t.level, (int)t.kernel);
            exit(EXIT_FAILURE);
        }
    }
}

// Reference flop model (src/pangulu_kernel_interface.c:4-176) in closed form.  The merge loops of the
// reference count matching index pairs; on patterns closed under symbolic fill (which build_blocks
// produces) every candidate matches, which gives per task
//   GETRF  sum over levels l with m_l = |L(:,l)|:  m_l + 2 m_l^2
//   TSTRF  sum over entries (r,c) of B:            1 + 2 (|U(c,:)| - 1)
//   GESSM  sum over entries (r,c) of B:            2 |L(:,r)|
//   SSSSM  sum over entries (t,j) of B:            2 |A(:,t)|
// (tests/test_host_structure.py and tests/test_oracle.py check the closed forms against the merge-based oracle and the reference's own
// -DPANGULU_PERF counter).  Bytes per task are the algorithmic figures of SURVEY.md 8(d).
void work_model(const BlockStore& B, const Schedule& S, WorkModel& W)
{
    const int nb = B.nb;
    const double sv = sizeof(T), si = sizeof(iidx), P = sizeof(iptr) * (double)(nb + 1);
    // row counts of upper blocks used as op2 (entries per row t of B), computed once per block
    std::vector<std::vector<uint16_t>> rowcnt(B.slots.size());
    auto get_rowcnt = [&](const pangulu_storage_slot_t* s) -> const std::vector<uint16_t>& {
        const size_t id = (size_t)(s - B.slots.data());
        std::vector<uint16_t>& v = rowcnt[id];
        if (v.empty())
        {
            v.assign((size_t)nb, 0);
            const iptr nnz = s->columnpointer[nb];
            for (iptr e = 0; e < nnz; ++e) v[s->rowindex[e]]++;
        }
        return v;
    };
    // fill the cache serially (cheap, O(nnz) once per op2 block), then sum tasks in parallel
    for (const pangulu_task_t& t : S.tasks)
        if (t.kernel_id == PANGULU_TASK_SSSSM) get_rowcnt(t.op2);
        else if (t.kernel_id == PANGULU_TASK_GESSM) get_rowcnt(t.opdst);

    double fg = 0, ft = 0, fe = 0, fs = 0, bg = 0, bt = 0, be = 0, bs = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : fg, ft, fe, fs, bg, bt, be, bs)
    for (int64_t ti = 0; ti < (int64_t)S.tasks.size(); ++ti)
    {
        const pangulu_task_t& t = S.tasks[ti];
        auto colcnt = [&](const pangulu_storage_slot_t* s, int c) { return (double)(s->columnpointer[c + 1] - s->columnpointer[c]); };
        auto nnz = [&](const pangulu_storage_slot_t* s) { return (double)s->columnpointer[nb]; };
        switch (t.kernel_id)
        {
        case PANGULU_TASK_GETRF:
        {
            const pangulu_storage_slot_t* lo = t.opdst->is_upper ? t.opdst->related_block : t.opdst;
            const pangulu_storage_slot_t* up = lo->related_block;
            double f = 0;
            const int grp = (int)(t.compare_flag >> 32);
            const uint16_t* bd = &S.group_bounds[(size_t)t.row * (DIAG_MAX_GROUPS + 1)];
            const int c0 = bd[grp], c1 = bd[grp + 1];
            for (int l = c0; l < c1; ++l)
            {
                if (up->columnpointer[l] == up->columnpointer[l + 1]) continue;
                const double m = colcnt(lo, l);
                f += m + 2.0 * m * m;
            }
            fg += f;
            bg += ((double)(lo->columnpointer[c1] - lo->columnpointer[c0]) + (double)(up->columnpointer[c1] - up->columnpointer[c0])) * (2 * sv + si) +
                  2 * P * (double)(c1 - c0) / nb;
            break;
        }
        case PANGULU_TASK_TSTRF:
        {
            const pangulu_storage_slot_t* lo = t.op1->is_upper ? t.op1->related_block : t.op1;
            const pangulu_storage_slot_t* up = lo->related_block;
            double f = 0;
            for (int c = 0; c < nb; ++c)
            {
                const double bc = colcnt(t.opdst, c);
                if (bc > 0) f += bc * (1.0 + 2.0 * (colcnt(up, c) - 1.0));
            }
            ft += f;
            bt += nnz(up) * (sv + si) + nnz(t.opdst) * (2 * sv + si) + 2 * P;
            break;
        }
        case PANGULU_TASK_GESSM:
        {
            const pangulu_storage_slot_t* lo = t.op1->is_upper ? t.op1->related_block : t.op1;
            const std::vector<uint16_t>& rc = rowcnt[(size_t)(t.opdst - B.slots.data())];
            double f = 0;
            for (int r = 0; r < nb; ++r)
                if (rc[r]) f += 2.0 * rc[r] * colcnt(lo, r);
            fe += f;
            be += nnz(lo) * (sv + si) + nnz(t.opdst) * (2 * sv + si) + 2 * P;
            break;
        }
        case PANGULU_TASK_SSSSM:
        {
            const std::vector<uint16_t>& rc = rowcnt[(size_t)(t.op2 - B.slots.data())];
            double f = 0;
            for (int k = 0; k < nb; ++k)
                if (rc[k]) f += 2.0 * rc[k] * colcnt(t.op1, k);
            fs += f;
            double nc = nnz(t.opdst);
            if (t.opdst->brow_pos == t.opdst->bcol_pos) nc += nnz(t.opdst->related_block);
            bs += nnz(t.op1) * (sv + si) + nnz(t.op2) * (sv + si) + nc * (2 * sv + si) + 3 * P;
            break;
        }
        }
    }
    W.flop_getrf = fg; W.flop_tstrf = ft; W.flop_gessm = fe; W.flop_ssssm = fs;
    W.bytes_getrf = bg; W.bytes_tstrf = bt; W.bytes_gessm = be; W.bytes_ssssm = bs;
}

}  // namespace pdrv
