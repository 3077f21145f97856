// qb_diam.h -- host-side conversion of CSR / Dia operators into diagonal-masked slices
// (DIAM).  Pure C++ (no CUDA) so the format can be unit-tested without a GPU.
//
// A slice is 32 consecutive rows (one warp).  Its non-zeros are grouped by diagonal
// offset (col - row); each group is one entry (offset, 32-bit lane mask) and its values
// are appended to the value array in lane order.  Duplicate (row, col) pairs -- legal in
// the reference's CSR (core/data/csr.pyx) -- become extra entries with the same offset.
#pragma once
#include <algorithm>
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
}  // namespace qbdiam
