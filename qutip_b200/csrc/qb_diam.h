// qb_diam.h -- host-side conversion of CSR / Dia operators into diagonal-masked slices
// (DIAM).  Pure C++ (no CUDA) so the format can be unit-tested without a GPU.
//
// A slice is 32 consecutive rows (one warp).  Its non-zeros are grouped by diagonal
// offset (col - row); each group is one entry (offset, 32-bit lane mask) and its values
// are appended to the value array in lane order.  Duplicate (row, col) pairs -- legal in
// the reference's CSR (core/data/csr.pyx) -- become extra entries with the same offset.
#pragma once
#include <algorithm>
#include <map>
#include <string>
#include <string.h>
#include <thread>
#include <vector>
#include "qb_types.h"

#ifndef __CUDACC__
struct int2 { int x, y; };
#endif

namespace qbdiam {
struct Entry { int off; int lane; qb_c128 v; };
struct SliceOut {
    std::vector<int2> ent;
    std::vector<qb_c128> val;
    std::vector<int> ent_count;          // per slice
    std::vector<long long> val_count;    // per slice
};

// entries of one slice sorted by (off, lane) -> (offset, mask) groups + packed values
inline void emit_slice(std::vector<Entry>& es, SliceOut& o) {
    std::stable_sort(es.begin(), es.end(), [](const Entry& a, const Entry& b) {
        return a.off != b.off ? a.off < b.off : a.lane < b.lane; });
    int ne = 0; long long nv = 0;
    size_t i = 0;
    std::vector<Entry> dup;
    while (i < es.size()) {
        const int off = es[i].off;
        unsigned mask = 0;
        dup.clear();
        size_t j = i;
        for (; j < es.size() && es[j].off == off; j++) {
            if (mask & (1u << es[j].lane)) { dup.push_back(es[j]); continue; }   // duplicate (row, col)
            mask |= 1u << es[j].lane;
            o.val.push_back(es[j].v); nv++;
        }
        int2 d; d.x = off; d.y = (int)mask;
        o.ent.push_back(d); ne++;
        // duplicates of the same (row, col): extra entries with the same offset
        while (!dup.empty()) {
            unsigned m2 = 0;
            std::vector<Entry> rest;
            for (auto& q : dup) {
                if (m2 & (1u << q.lane)) { rest.push_back(q); continue; }
                m2 |= 1u << q.lane; o.val.push_back(q.v); nv++;
            }
            int2 d2; d2.x = off; d2.y = (int)m2;
            o.ent.push_back(d2); ne++;
            dup.swap(rest);
        }
        i = j;
    }
    o.ent_count.push_back(ne);
    o.val_count.push_back(nv);
}

struct DiamHost {
    std::vector<int> slice_ptr;
    std::vector<int2> ent;
    std::vector<long long> slice_vbase;
    std::vector<qb_c128> val;
};

template <class F> inline void build_diam(int64_t nrows, F fill_slice, DiamHost& out) {
    const int64_t nslices = (nrows + 31) / 32;
    int nthreads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (nslices < 4096) nthreads = 1;
    std::vector<SliceOut> parts(nthreads);
    auto work = [&](int t) {
        const int64_t lo = nslices * t / nthreads, hi = nslices * (t + 1) / nthreads;
        std::vector<Entry> es;
        for (int64_t sl = lo; sl < hi; sl++) { es.clear(); fill_slice(sl, es); emit_slice(es, parts[t]); }
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    out.slice_ptr.assign(1, 0);
    long long vb = 0;
    for (auto& p : parts) {
        for (size_t i = 0; i < p.ent_count.size(); i++) {
            out.slice_ptr.push_back(out.slice_ptr.back() + p.ent_count[i]);
            out.slice_vbase.push_back(vb);
            vb += p.val_count[i];
        }
        out.ent.insert(out.ent.end(), p.ent.begin(), p.ent.end());
        out.val.insert(out.val.end(), p.val.begin(), p.val.end());
    }
}

// ---- sliced ELLPACK (32-row slices, slot-major, explicit column indices) ----
// Used for operators that stay L2-resident (many trajectories re-read them): the sweep is
// three coalesced loads and four FMAs per slot with no mask/shuffle arithmetic, i.e. far
// fewer instructions than DIAM, at 20 B instead of ~16.3 B per stored element.
struct SellHost {
    std::vector<int> slice_ptr;          // [nslices + 1], in slots
    std::vector<qb_c128> val;            // [nslots * 32]
    std::vector<int> col;                // [nslots * 32]
    long long nnz = 0;
};

template <class RowFn>   // row_fn(r, std::vector<std::pair<int, qb_c128>>& out) appends (col, val)
inline void build_sell(int64_t nrows, int64_t ncols, RowFn row_fn, SellHost& out) {
    const int64_t nslices = (nrows + 31) / 32;
    out.slice_ptr.assign(1, 0);
    std::vector<std::vector<std::pair<int, qb_c128>>> rows(32);
    for (int64_t sl = 0; sl < nslices; sl++) {
        int width = 0;
        for (int l = 0; l < 32; l++) {
            rows[l].clear();
            const int64_t r = sl * 32 + l;
            if (r < nrows) {
                row_fn(r, rows[l]);
                std::stable_sort(rows[l].begin(), rows[l].end(),
                                 [](const std::pair<int, qb_c128>& a, const std::pair<int, qb_c128>& b) {
                                     return a.first < b.first; });
                out.nnz += (long long)rows[l].size();
            }
            width = std::max(width, (int)rows[l].size());
        }
        const size_t base = out.val.size();
        out.val.resize(base + (size_t)width * 32);
        out.col.resize(base + (size_t)width * 32);
        for (int k = 0; k < width; k++)
            for (int l = 0; l < 32; l++) {
                const int64_t r = sl * 32 + l;
                qb_c128 v = {0.0, 0.0};
                int c = (int)std::min<int64_t>(r, ncols - 1);
                if (c < 0) c = 0;
                if (k < (int)rows[l].size()) { c = rows[l][k].first; v = rows[l][k].second; }
                out.val[base + (size_t)k * 32 + l] = v;
                out.col[base + (size_t)k * 32 + l] = c;
            }
        out.slice_ptr.push_back(out.slice_ptr.back() + width);
    }
}

// ---- RSELL: SELL with per-slot rule descriptors (qb_types.h) ----
// One slot of a slice = one "diagonal" of the slice under the better of two keys:
// col - row (banded operators: ladder operators, hopping terms) or col ^ row (operators of
// tensor products of two-level systems flip bits of the row index).  Slots whose lanes all
// follow the rule need no column block; slots whose 32 values are bitwise equal need no
// value block.  Slices that are not diagonal structured under either key fall back to plain
// ELLPACK slots (explicit columns and values, row order) -- the format is a superset of SELL.
// Descriptor lists are de-duplicated: explicit blocks are addressed relative to the slice's
// own base (sinfo), so every interior slice of a structured operator shares ONE list -- small
// enough to live in the constant bank (kernel parameters) of the fused pass kernel.
struct RsellHost {
    // [nslices][4]: first descriptor, count | fast count << 12 | xor-fast count << 24,
    // value-block base, column-block base.  A list holds the xor-rule constant slots first,
    // then the add-rule constant slots ("fast": no explicit block), then the rest.
    std::vector<int> sinfo;
    bool overflow = false;               // a slice does not fit the packed counts: use another format
    std::vector<QbSlotDesc> desc;        // de-duplicated descriptor lists
    std::vector<qb_c128> val;            // explicit value blocks, 32 per block
    std::vector<int> col;                // explicit column blocks, 32 per block
    long long nnz = 0, slots = 0;
    long long stored() const { return slots; }
    long long bytes() const {
        return (long long)desc.size() * 32 + (long long)val.size() * 16 + (long long)col.size() * 4 +
               (long long)sinfo.size() * 4;
    }
};

// slices [sl_lo, sl_hi) into `out` (its own de-duplication map; lists[] = the unique lists)
template <class RowFn>
inline void build_rsell_range(int64_t nrows, int64_t ncols, int64_t sl_lo, int64_t sl_hi, RowFn& row_fn,
                              RsellHost& out, std::vector<std::pair<int, int>>& lists) {
    std::vector<std::vector<std::pair<int, qb_c128>>> rows(32);
    std::vector<QbSlotDesc> cur;          // descriptor list of the slice being built
    int vbase = 0, cbase = 0;
    std::map<std::string, int> seen;      // descriptor list (bytes) -> first descriptor
    auto emit = [&](int rule, int delta, const int* cols, const qb_c128* vals, const bool* have,
                    int64_t r0) {
        // column rule usable only if every lane's implied column is a valid index
        bool rule_ok = rule != QB_RS_COL_EXPL;
        if (rule_ok)
            for (int l = 0; l < 32; l++) {
                const long long r = r0 + l;
                const long long c = rule == QB_RS_COL_ADD ? r + delta : (r ^ (long long)delta);
                if (c < 0 || c >= ncols) { rule_ok = false; break; }
            }
        bool konst = true;
        for (int l = 0; l < 32; l++)
            if (!have[l] || memcmp(&vals[l], &vals[0], sizeof(qb_c128)) != 0) { konst = false; break; }
        QbSlotDesc d;
        memset(&d, 0, sizeof d);
        d.rule = rule_ok ? rule : QB_RS_COL_EXPL; d.delta = rule_ok ? delta : 0;
        if (!rule_ok) {
            d.cpos = (int)(out.col.size() / 32) - cbase;
            for (int l = 0; l < 32; l++) {
                long long c = have[l] ? cols[l] : std::min<long long>(r0 + l, ncols - 1);
                if (c < 0) c = 0;
                out.col.push_back((int)c);
            }
        }
        if (konst) { d.rule |= QB_RS_VAL_CONST; d.vre = vals[0].re; d.vim = vals[0].im; }
        else {
            d.vpos = (int)(out.val.size() / 32) - vbase;
            for (int l = 0; l < 32; l++) out.val.push_back(have[l] ? vals[l] : qb_c128{0.0, 0.0});
        }
        cur.push_back(d);
    };
    for (int64_t sl = sl_lo; sl < sl_hi; sl++) {
        const int64_t r0 = sl * 32;
        int width = 0; long long cnt = 0;
        for (int l = 0; l < 32; l++) {
            rows[l].clear();
            if (r0 + l < nrows) {
                row_fn(r0 + l, rows[l]);
                std::stable_sort(rows[l].begin(), rows[l].end(),
                                 [](const std::pair<int, qb_c128>& a, const std::pair<int, qb_c128>& b) {
                                     return a.first < b.first; });
            }
            width = std::max(width, (int)rows[l].size());
            cnt += (long long)rows[l].size();
        }
        out.nnz += cnt;
        cur.clear();
        vbase = (int)(out.val.size() / 32); cbase = (int)(out.col.size() / 32);
        // distinct (key, occurrence) pairs under both keys; occurrence > 0 only for duplicate
        // (row, col) pairs, which the reference's CSR allows (core/data/csr.pyx)
        std::vector<std::pair<long long, int>> kadd, kxor;
        for (int l = 0; l < 32; l++) {
            int occ = 0;
            for (size_t k = 0; k < rows[l].size(); k++) {
                occ = (k > 0 && rows[l][k].first == rows[l][k - 1].first) ? occ + 1 : 0;
                const long long r = r0 + l, c = rows[l][k].first;
                kadd.push_back({c - r, occ}); kxor.push_back({c ^ r, occ});
            }
        }
        auto uniq = [](std::vector<std::pair<long long, int>>& v) {
            std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); };
        uniq(kadd); uniq(kxor);
        const bool use_xor = kxor.size() < kadd.size();
        const auto& keys = use_xor ? kxor : kadd;
        const int rule = use_xor ? QB_RS_COL_XOR : QB_RS_COL_ADD;
        int cols[32]; qb_c128 vals[32]; bool have[32];
        if (cnt > 0 && (long long)keys.size() * 32 <= 2 * cnt + 64) {
            for (const auto& kk : keys) {
                for (int l = 0; l < 32; l++) {
                    have[l] = false; vals[l] = qb_c128{0.0, 0.0}; cols[l] = 0;
                    const long long r = r0 + l;
                    int occ = 0;
                    for (size_t k = 0; k < rows[l].size(); k++) {
                        occ = (k > 0 && rows[l][k].first == rows[l][k - 1].first) ? occ + 1 : 0;
                        const long long c = rows[l][k].first;
                        const long long key = use_xor ? (c ^ r) : (c - r);
                        if (key == kk.first && occ == kk.second) {
                            have[l] = true; vals[l] = rows[l][k].second; cols[l] = (int)c; break;
                        }
                    }
                }
                emit(rule, (int)kk.first, cols, vals, have, r0);
            }
        } else {
            for (int k = 0; k < width; k++) {
                for (int l = 0; l < 32; l++) {
                    have[l] = k < (int)rows[l].size();
                    vals[l] = have[l] ? rows[l][k].second : qb_c128{0.0, 0.0};
                    cols[l] = have[l] ? rows[l][k].first : 0;
                }
                emit(QB_RS_COL_EXPL, 0, cols, vals, have, r0);
            }
        }
        // "fast" slots (column rule + constant value: no explicit block at all) first; the
        // sweep runs them in a lean loop of their own
        auto is_fast = [](const QbSlotDesc& d) {
            return (d.rule & QB_RS_COL_MASK) != QB_RS_COL_EXPL && (d.rule & QB_RS_VAL_CONST); };
        auto is_xfast = [&](const QbSlotDesc& d) { return is_fast(d) && (d.rule & QB_RS_COL_MASK) == QB_RS_COL_XOR; };
        std::stable_partition(cur.begin(), cur.end(), is_fast);
        std::stable_partition(cur.begin(), cur.end(), is_xfast);
        int nfast = 0, nxor = 0;
        for (auto& d : cur) { nfast += is_fast(d); nxor += is_xfast(d); }
        if (cur.size() > 4095 || nxor > 255) out.overflow = true;
        // share the descriptor list with an earlier slice that has the same one
        int dstart = (int)out.desc.size();
        if (!cur.empty()) {
            std::string keyb(reinterpret_cast<const char*>(cur.data()), cur.size() * sizeof(QbSlotDesc));
            auto it = seen.find(keyb);
            if (it != seen.end()) dstart = it->second;
            else {
                seen.emplace(std::move(keyb), dstart); lists.push_back({dstart, (int)cur.size()});
                out.desc.insert(out.desc.end(), cur.begin(), cur.end());
            }
        }
        out.slots += (long long)cur.size();
        out.sinfo.push_back(dstart);
        out.sinfo.push_back((int)(cur.size() & 4095) | ((nfast & 4095) << 12) | ((nxor & 255) << 24));
        out.sinfo.push_back(vbase); out.sinfo.push_back(cbase);
    }
}


// all slices; large operators are converted by several threads (contiguous slice ranges,
// merged in order: the result does not depend on the number of threads)
template <class RowFn>
inline void build_rsell(int64_t nrows, int64_t ncols, RowFn row_fn, RsellHost& out) {
    const int64_t nslices = (nrows + 31) / 32;
    int nthreads = (int)std::min<int64_t>(std::max(1u, std::thread::hardware_concurrency()), 16);
    if (nslices < 4096) nthreads = 1;
    std::vector<RsellHost> parts(nthreads);
    std::vector<std::vector<std::pair<int, int>>> lists(nthreads);
    auto work = [&](int t) {
        build_rsell_range(nrows, ncols, nslices * t / nthreads, nslices * (t + 1) / nthreads, row_fn,
                          parts[t], lists[t]);
    };
    if (nthreads == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++) th.emplace_back(work, t);
        for (auto& x : th) x.join();
    }
    if (nthreads == 1) { out = std::move(parts[0]); return; }
    std::map<std::string, int> seen;
    for (int t = 0; t < nthreads; t++) {
        RsellHost& p = parts[t];
        // unique lists of this part -> global position
        std::map<int, int> remap;
        for (auto& l : lists[t]) {
            std::string key(reinterpret_cast<const char*>(p.desc.data() + l.first), (size_t)l.second * sizeof(QbSlotDesc));
            auto it = seen.find(key);
            int g;
            if (it != seen.end()) g = it->second;
            else {
                g = (int)out.desc.size();
                seen.emplace(std::move(key), g);
                out.desc.insert(out.desc.end(), p.desc.begin() + l.first, p.desc.begin() + l.first + l.second);
            }
            remap[l.first] = g;
        }
        const int vb = (int)(out.val.size() / 32), cb = (int)(out.col.size() / 32);
        for (size_t i = 0; i + 3 < p.sinfo.size(); i += 4) {
            const int cnt = p.sinfo[i + 1] & 4095;
            out.sinfo.push_back(cnt ? remap[p.sinfo[i]] : (int)out.desc.size());
            out.sinfo.push_back(p.sinfo[i + 1]);
            out.sinfo.push_back(p.sinfo[i + 2] + vb);
            out.sinfo.push_back(p.sinfo[i + 3] + cb);
        }
        out.val.insert(out.val.end(), p.val.begin(), p.val.end());
        out.col.insert(out.col.end(), p.col.begin(), p.col.end());
        out.nnz += p.nnz; out.slots += p.slots; out.overflow = out.overflow || p.overflow;
        p = RsellHost();
    }
}

// y = A x on the host (format unit tests; never used by the product path)
inline void rsell_matvec_host(const RsellHost& A, int64_t nrows, const qb_c128* x, qb_c128* y) {
    const int64_t nslices = (nrows + 31) / 32;
    for (int64_t sl = 0; sl < nslices; sl++)
        for (int l = 0; l < 32; l++) {
            const int64_t r = sl * 32 + l;
            if (r >= nrows) continue;
            double re = 0.0, im = 0.0;
            const int* si = &A.sinfo[(size_t)sl * 4];
            for (int k = si[0]; k < si[0] + (si[1] & 4095); k++) {
                const QbSlotDesc& d = A.desc[k];
                const int cr = d.rule & QB_RS_COL_MASK;
                const long long c = cr == QB_RS_COL_ADD ? r + d.delta : cr == QB_RS_COL_XOR ? (r ^ (long long)d.delta)
                                                                       : A.col[(size_t)(si[3] + d.cpos) * 32 + l];
                qb_c128 v = {d.vre, d.vim};
                if (!(d.rule & QB_RS_VAL_CONST)) v = A.val[(size_t)(si[2] + d.vpos) * 32 + l];
                re += v.re * x[c].re - v.im * x[c].im;
                im += v.re * x[c].im + v.im * x[c].re;
            }
            y[r].re = re; y[r].im = im;
        }
}
}  // namespace qbdiam
