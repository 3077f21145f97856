// qb_ops.cu -- data-layer objects and operations behind the C ABI: uploads (with the
// host-side conversion of CSR / Dia into diagonal-masked slices), matmul, and the vector
// kernels the reference's RK loop calls through its dispatchers (axpy, scal, norms, wrms
// error, expectation values).
#include <algorithm>
#include <string.h>
#include <thread>
#include <vector>
#include "qb_host.h"
#include "qb_kernels.cuh"
#include "qb_diam.h"

thread_local std::string g_qb_err;
std::atomic<long long> g_qb_launches{0};

extern "C" int qb_version(void) { return 100; }
extern "C" const char* qb_last_error(void) { return g_qb_err.c_str(); }
extern "C" int64_t qb_launch_count(void) { return g_qb_launches; }
extern "C" int qb_device_count(int* n) {
    if (!n) QB_FAIL(QB_E_ARG, "null");
    cudaError_t e = cudaGetDeviceCount(n);
    if (e != cudaSuccess) { *n = 0; QB_FAIL(QB_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    return QB_OK;
}
extern "C" int qb_set_device(int dev) { QB_CUDA(cudaSetDevice(dev)); return QB_OK; }
extern "C" int qb_device_mem_info(int64_t* free_bytes, int64_t* total_bytes) {
    size_t f = 0, t = 0;
    QB_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (int64_t)f;
    if (total_bytes) *total_bytes = (int64_t)t;
    return QB_OK;
}
extern "C" int qb_synchronize(void) { QB_CUDA(cudaDeviceSynchronize()); return QB_OK; }

// ------------------------------------------------------------------ dense
extern "C" int qb_dense_zeros(int64_t rows, int64_t cols, int fortran, qb_handle* out) {
    if (rows < 0 || cols < 0 || !out) QB_FAIL(QB_E_ARG, "bad dense shape");
    QbDenseH* d = new QbDenseH();
    d->rows = rows; d->cols = cols; d->fortran = fortran ? 1 : 0;
    size_t bytes = std::max<size_t>(16, (size_t)rows * cols * 16);
    cudaError_t e = cudaMalloc((void**)&d->d, bytes);
    if (e != cudaSuccess) { d->d = nullptr; delete d; QB_FAIL(QB_E_ALLOC, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
    e = cudaMemset(d->d, 0, bytes);
    if (e != cudaSuccess) { delete d; QB_FAIL(QB_E_CUDA, "cudaMemset: %s", cudaGetErrorString(e)); }
    *out = d;
    return QB_OK;
}
extern "C" int qb_dense_upload(const void* host, int64_t rows, int64_t cols, int fortran, qb_handle* out) {
    if (!host) QB_FAIL(QB_E_ARG, "null host pointer");
    int rc = qb_dense_zeros(rows, cols, fortran, out);
    if (rc) return rc;
    QbDenseH* d = static_cast<QbDenseH*>(*out);
    cudaError_t e = cudaMemcpy(d->d, host, (size_t)rows * cols * 16, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { delete d; *out = nullptr; QB_FAIL(QB_E_CUDA, "H2D copy: %s", cudaGetErrorString(e)); }
    return QB_OK;
}
extern "C" int qb_dense_download(qb_handle h, void* host) {
    QbDenseH* d = qb_cast<QbDenseH>(h, QB_TAG_DENSE);
    if (!d || !host) QB_FAIL(QB_E_TYPE, "not a dense handle");
    QB_CUDA(cudaMemcpy(host, d->d, (size_t)d->size() * 16, cudaMemcpyDeviceToHost));
    return QB_OK;
}
extern "C" int qb_dense_write(qb_handle h, const void* host) {
    QbDenseH* d = qb_cast<QbDenseH>(h, QB_TAG_DENSE);
    if (!d || !host) QB_FAIL(QB_E_TYPE, "not a dense handle");
    QB_CUDA(cudaMemcpy(d->d, host, (size_t)d->size() * 16, cudaMemcpyHostToDevice));
    return QB_OK;
}
// re-interpret a column-major buffer with another shape of the same size (column stacking /
// unstacking of rho: solver_base.py:134-147 stack_columns / unstack_columns), no copy
extern "C" int qb_dense_reshape(qb_handle h, int64_t rows, int64_t cols) {
    QbDenseH* d = qb_cast<QbDenseH>(h, QB_TAG_DENSE);
    if (!d) QB_FAIL(QB_E_TYPE, "not a dense handle");
    if (rows < 0 || cols < 0 || rows * cols != d->rows * d->cols) QB_FAIL(QB_E_SHAPE, "reshape changes the size");
    if (!d->fortran && d->rows != 1 && d->cols != 1) QB_FAIL(QB_E_TYPE, "reshape needs a column-major buffer");
    d->rows = rows; d->cols = cols; d->fortran = 1;
    return QB_OK;
}
extern "C" int qb_dense_copy(qb_handle h, qb_handle* out) {
    QbDenseH* d = qb_cast<QbDenseH>(h, QB_TAG_DENSE);
    if (!d) QB_FAIL(QB_E_TYPE, "not a dense handle");
    int rc = qb_dense_zeros(d->rows, d->cols, d->fortran, out);
    if (rc) return rc;
    QB_CUDA(cudaMemcpy(static_cast<QbDenseH*>(*out)->d, d->d, (size_t)d->size() * 16, cudaMemcpyDeviceToDevice));
    return QB_OK;
}
extern "C" int qb_dense_info(qb_handle h, int64_t* rows, int64_t* cols, int* fortran, void** devptr) {
    QbDenseH* d = qb_cast<QbDenseH>(h, QB_TAG_DENSE);
    if (!d) QB_FAIL(QB_E_TYPE, "not a dense handle");
    if (rows) *rows = d->rows;
    if (cols) *cols = d->cols;
    if (fortran) *fortran = d->fortran;
    if (devptr) *devptr = d->d;
    return QB_OK;
}
extern "C" int qb_free(qb_handle h) {
    QbObj* o = static_cast<QbObj*>(h);
    if (!o) return QB_OK;
    if (o->tag != QB_TAG_DENSE && o->tag != QB_TAG_OP && o->tag != QB_TAG_SYS && o->tag != QB_TAG_ENG)
        QB_FAIL(QB_E_TYPE, "qb_free: not a live handle");
    delete o;
    return QB_OK;
}

namespace {
using namespace qbdiam;
template <class T> static int to_device(QbOpH* h, const std::vector<T>& v, const T** out) {
    void* p = nullptr;
    size_t bytes = std::max<size_t>(16, v.size() * sizeof(T));
    QB_CUDA(cudaMalloc(&p, bytes));
    h->owned.push_back(p);
    h->device_bytes += (int64_t)bytes;
    if (!v.empty()) QB_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = static_cast<const T*>(p);
    return QB_OK;
}

static int finish_diam(QbOpH* h, DiamHost& dh, int64_t rows, int64_t cols) {
    h->dev.fmt = QB_FMT_DIAM; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols;
    h->dev.nnz = (long long)dh.val.size();
    h->avg_lanes = dh.ent.empty() ? 0.0 : (double)dh.val.size() / (double)dh.ent.size();
    int rc;
    const int2* entp = nullptr;
    if ((rc = to_device(h, dh.slice_ptr, &h->dev.slice_ptr))) return rc;
    if ((rc = to_device(h, dh.ent, &entp))) return rc;
    h->dev.ent_off = reinterpret_cast<const int*>(entp);
    h->dev.ent_mask = nullptr;
    if ((rc = to_device(h, dh.slice_vbase, &h->dev.slice_vbase))) return rc;
    if ((rc = to_device(h, dh.val, &h->dev.val))) return rc;
    // operators much larger than the 126 MB L2 are streamed (read once per pass)
    h->dev.pad_ = h->device_bytes > (int64_t)96 << 20 ? 1 : 0;
    return QB_OK;
}
static int finish_sell(QbOpH* h, SellHost& sh, int64_t rows, int64_t cols) {
    h->dev.fmt = QB_FMT_SELL; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols;
    h->dev.nnz = sh.nnz;
    int rc;
    if ((rc = to_device(h, sh.slice_ptr, &h->dev.slice_ptr))) return rc;
    if ((rc = to_device(h, sh.val, &h->dev.val))) return rc;
    if ((rc = to_device(h, sh.col, &h->dev.col))) return rc;
    return QB_OK;
}
static int finish_rsell(QbOpH* h, RsellHost& rs, int64_t rows, int64_t cols) {
    h->dev.fmt = QB_FMT_RSELL; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols;
    h->dev.nnz = rs.nnz;
    int rc;
    h->dev.ndesc = (int)rs.desc.size();
    if ((rc = to_device(h, rs.sinfo, &h->dev.sinfo))) return rc;
    if ((rc = to_device(h, rs.desc, &h->dev.sdesc))) return rc;
    if ((rc = to_device(h, rs.val, &h->dev.val))) return rc;
    if ((rc = to_device(h, rs.col, &h->dev.col))) return rc;
    return QB_OK;
}
// RSELL replaces SELL when the slices are diagonal structured: at most 1.75 stored lanes per
// non-zero (every slot costs one gather for all 32 lanes) and L2-resident.  It also replaces
// DIAM / CSR for LARGE operators (HBM streams) when rule compression at least halves the bytes
// read per product -- e.g. the C2 Liouvillian: 60 MB instead of 402 MB (DIAM) / 497 MB (CSR),
// because its Hamiltonian part is constant diagonals under the xor key.
static bool want_rsell(const RsellHost& rs, int64_t rows, bool small_sell, long long sell_padded,
                       long long other_bytes) {
    if (getenv("QB_NO_RSELL") || rs.nnz == 0 || rows < 32 || rs.overflow) return false;
    const long long stored = rs.stored() * 32;
    if (small_sell)
        return rs.bytes() <= (48ll << 20) && (double)stored <= 1.75 * (double)rs.nnz &&
               stored <= sell_padded + sell_padded / 4;
    if (getenv("QB_NO_BIG_RSELL")) return false;
    return 2 * rs.bytes() <= other_bytes && (double)stored <= 2.5 * (double)rs.nnz;
}
// operators up to this size stay L2-resident when many trajectories re-read them: prefer
// the instruction-lean SELL sweep; larger ones are HBM streams: prefer the compact DIAM
static const long long QB_SELL_MAX_BYTES = 48ll << 20;
}  // namespace

extern "C" int qb_csr_upload(const void* data, const int32_t* col, const int32_t* rowptr,
                             int64_t rows, int64_t cols, int64_t nnz, int format, qb_handle* out) {
    if (!rowptr || rows < 0 || cols < 0 || nnz < 0 || !out || (nnz > 0 && (!data || !col)))
        QB_FAIL(QB_E_ARG, "bad CSR arguments");
    if (rows > 0x7fffffff || cols > 0x7fffffff || nnz > 0x7fffffff) QB_FAIL(QB_E_ARG, "CSR too large for int32 indices");
    if (rowptr[rows] != nnz) QB_FAIL(QB_E_ARG, "row_index[rows] != nnz");
    const qb_c128* v = static_cast<const qb_c128*>(data);
    for (int64_t p = 0; p < nnz; p++)
        if (col[p] < 0 || col[p] >= cols) QB_FAIL(QB_E_ARG, "column index out of range");
    QbOpH* h = new QbOpH();
    int rc = QB_OK;
    auto row_fn = [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o) {
        for (int p = rowptr[r]; p < rowptr[r + 1]; p++) o.push_back({col[p], v[p]});
    };
    // a large operator (an HBM stream) can neither be SELL nor needs the DIAM analysis when the
    // rule compression already wins against DIAM's lower bound of 16 B per non-zero: the three
    // analysers each walk all non-zeros, so skipping two of them halves the conversion time of
    // e.g. the C2 Liouvillian.  The decisions are the same as with all three built.
    const bool big = (long long)nnz * 20 > QB_SELL_MAX_BYTES;
    RsellHost rs;
    bool use_rsell = (format == 5), rsell_built = false;
    if (format == 0 && big && nnz > 0 && rows >= 32) {
        build_rsell(rows, cols, row_fn, rs);
        rsell_built = true;
        use_rsell = want_rsell(rs, rows, false, 0, (long long)nnz * 16);
    }
    bool use_diam = (format == 2);
    DiamHost dh;
    if (format != 1 && !use_rsell && format != 5 && format != 3) {
        build_diam(rows, [&](int64_t sl, std::vector<Entry>& es) {
            const int64_t r0 = sl * 32, r1 = std::min<int64_t>(rows, r0 + 32);
            for (int64_t r = r0; r < r1; r++)
                for (int p = rowptr[r]; p < rowptr[r + 1]; p++) {
                    Entry e; e.off = (int)(col[p] - r); e.lane = (int)(r - r0); e.v = v[p];
                    es.push_back(e);
                }
        }, dh);
        const double avg = dh.ent.empty() ? 0.0 : (double)dh.val.size() / (double)dh.ent.size();
        // diagonal structured enough that 16 B/nnz + 8 B/entry beats CSR's 20 B/nnz and
        // the warp keeps >= 8 lanes busy per entry
        if (format == 0) use_diam = (nnz == 0) || avg >= 8.0 || rows < 32;
    }
    // format 3: force SELL; auto: SELL for small (L2-resident) operators with little padding
    bool use_sell = (format == 3);
    SellHost sh;
    if (format == 3 || (format == 0 && !big && !use_rsell)) {
        build_sell(rows, cols, row_fn, sh);
        const long long padded = (long long)sh.val.size();
        if (format == 0)
            use_sell = nnz > 0 && rows >= 32 && padded * 20 <= QB_SELL_MAX_BYTES &&
                       (double)padded <= 1.5 * (double)nnz;
    }
    if (format == 5 || (format == 0 && !use_rsell && nnz > 0 && rows >= 32)) {
        if (!rsell_built) build_rsell(rows, cols, row_fn, rs);
        const long long other = use_diam ? (long long)dh.val.size() * 16 + (long long)dh.ent.size() * 8
                                         : (long long)nnz * 20;
        if (format == 0) use_rsell = want_rsell(rs, rows, use_sell, (long long)sh.val.size(), other);
    }
    if (use_rsell) rc = finish_rsell(h, rs, rows, cols);
    else if (use_sell) rc = finish_sell(h, sh, rows, cols);
    else if (use_diam) rc = finish_diam(h, dh, rows, cols);
    else {
        h->dev.fmt = QB_FMT_CSR; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols; h->dev.nnz = nnz;
        std::vector<qb_c128> vv(v, v + nnz);
        std::vector<int> cc(col, col + nnz), rp(rowptr, rowptr + rows + 1);
        if (!(rc = to_device(h, vv, &h->dev.val)) && !(rc = to_device(h, cc, &h->dev.col)))
            rc = to_device(h, rp, &h->dev.rowptr);
    }
    if (rc) { delete h; return rc; }
    *out = h;
    return QB_OK;
}

extern "C" int qb_kron_upload(const void* data, const int32_t* col, const int32_t* rowptr, int64_t n,
                              int64_t nnz, int side, qb_handle* out) {
    if (!rowptr || n <= 0 || nnz < 0 || !out || (nnz > 0 && (!data || !col)) || (side != 0 && side != 1))
        QB_FAIL(QB_E_ARG, "bad Kronecker operator arguments");
    if (n > 46340) QB_FAIL(QB_E_ARG, "n*n exceeds int32 row indices");
    if (rowptr[n] != nnz) QB_FAIL(QB_E_ARG, "row_index[n] != nnz");
    const qb_c128* v = static_cast<const qb_c128*>(data);
    for (int64_t p = 0; p < nnz; p++)
        if (col[p] < 0 || col[p] >= n) QB_FAIL(QB_E_ARG, "column index out of range");
    QbOpH* h = new QbOpH();
    h->dev.fmt = QB_FMT_KRON; h->dev.nrows = (int)(n * n); h->dev.ncols = (int)(n * n);
    h->dev.nnz = nnz * n;                 // non-zeros of the equivalent superoperator
    h->dev.kn = (int)n; h->dev.kside = side;
    std::vector<qb_c128> vv(v, v + nnz);
    std::vector<int> cc(col, col + nnz), rp(rowptr, rowptr + n + 1);
    int rc;
    if (!(rc = to_device(h, vv, &h->dev.val)) && !(rc = to_device(h, cc, &h->dev.col)))
        rc = to_device(h, rp, &h->dev.rowptr);
    if (!rc && side == 0 && n % 32 == 0 && nnz > 0) {
        SellHost sh;
        build_sell(n, n, [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o) {
            for (int p = rowptr[r]; p < rowptr[r + 1]; p++) o.push_back({col[p], v[p]});
        }, sh);
        if ((double)sh.val.size() <= 2.0 * (double)nnz) {
            if (!(rc = to_device(h, sh.slice_ptr, &h->dev.slice_ptr)) &&
                !(rc = to_device(h, sh.val, &h->dev.kval)))
                rc = to_device(h, sh.col, &h->dev.kcol);
        }
    }
    if (rc) { delete h; return rc; }
    *out = h;
    return QB_OK;
}

extern "C" int qb_sandwich_upload(const void* data, const int32_t* col, const int32_t* rowptr,
                                  int64_t n, int64_t nstack, int64_t nnz, qb_handle* out) {
    if (!rowptr || n <= 0 || nstack <= 0 || nnz < 0 || !out || (nnz > 0 && (!data || !col)))
        QB_FAIL(QB_E_ARG, "bad sandwich operator arguments");
    if (n > 46340 || n * nstack > 0x7fffffff) QB_FAIL(QB_E_ARG, "sandwich operator too large for int32 indices");
    if (rowptr[n * nstack] != nnz) QB_FAIL(QB_E_ARG, "row_index[n*nstack] != nnz");
    const qb_c128* v = static_cast<const qb_c128*>(data);
    for (int64_t p = 0; p < nnz; p++)
        if (col[p] < 0 || col[p] >= n) QB_FAIL(QB_E_ARG, "column index out of range");
    QbOpH* h = new QbOpH();
    h->dev.fmt = QB_FMT_KRON; h->dev.nrows = (int)(n * n); h->dev.ncols = (int)(n * n);
    h->dev.kn = (int)n; h->dev.kside = 2; h->dev.kstack = (int)nstack;
    long long sq = 0;                     // non-zeros of the equivalent superoperator
    for (int64_t c = 0; c < nstack; c++) {
        const long long m = rowptr[(c + 1) * n] - rowptr[c * n];
        sq += m * m;
    }
    h->dev.nnz = sq;
    std::vector<qb_c128> vv(v, v + nnz);
    std::vector<int> cc(col, col + nnz), rp(rowptr, rowptr + n * nstack + 1);
    int rc;
    if (!(rc = to_device(h, vv, &h->dev.val)) && !(rc = to_device(h, cc, &h->dev.col)))
        rc = to_device(h, rp, &h->dev.rowptr);
    if (rc) { delete h; return rc; }
    *out = h;
    return QB_OK;
}

extern "C" int qb_dia_upload(const void* data, const int32_t* offsets, int64_t ndiag,
                             int64_t rows, int64_t cols, int format, qb_handle* out) {
    if (rows < 0 || cols < 0 || ndiag < 0 || !out || (ndiag > 0 && (!data || !offsets)))
        QB_FAIL(QB_E_ARG, "bad Dia arguments");
    if (rows > 0x7fffffff || cols > 0x7fffffff) QB_FAIL(QB_E_ARG, "Dia too large");
    const qb_c128* v = static_cast<const qb_c128*>(data);
    // diagonals visited in ascending offset (stable for duplicates): same per-row
    // summation order as a sorted CSR row
    std::vector<int> order(ndiag);
    for (int i = 0; i < ndiag; i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return offsets[a] < offsets[b]; });
    QbOpH* h = new QbOpH();
    DiamHost dh;
    build_diam(rows, [&](int64_t sl, std::vector<Entry>& es) {
        const int64_t r0 = sl * 32, r1 = std::min<int64_t>(rows, r0 + 32);
        for (int k = 0; k < ndiag; k++) {
            const int d = order[k];
            const int64_t off = offsets[d];
            for (int64_t r = r0; r < r1; r++) {
                const int64_t c = r + off;
                if (c < 0 || c >= cols) continue;
                const qb_c128 x = v[(size_t)d * cols + c];
                if (x.re == 0.0 && x.im == 0.0) continue;
                Entry e; e.off = (int)off; e.lane = (int)(r - r0); e.v = x;
                es.push_back(e);
            }
        }
    }, dh);
    int rc = QB_OK;
    const double avg = dh.ent.empty() ? 0.0 : (double)dh.val.size() / (double)dh.ent.size();
    bool use_diam = format == 2 || (format == 0 && (dh.val.empty() || avg >= 8.0 || rows < 32));
    bool use_sell = (format == 3);
    SellHost sh;
    if (format == 3 || format == 0) {
        build_sell(rows, cols, [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o) {
            for (int k = 0; k < ndiag; k++) {
                const int d = order[k];
                const int64_t c = r + offsets[d];
                if (c < 0 || c >= cols) continue;
                const qb_c128 x = v[(size_t)d * cols + c];
                if (x.re == 0.0 && x.im == 0.0) continue;
                o.push_back({(int)c, x});
            }
        }, sh);
        const long long padded = (long long)sh.val.size();
        if (format == 0)
            use_sell = sh.nnz > 0 && rows >= 32 && padded * 20 <= QB_SELL_MAX_BYTES &&
                       (double)padded <= 1.5 * (double)sh.nnz;
    }
    RsellHost rs;
    bool use_rsell = (format == 5);
    if (format == 5 || (format == 0 && !dh.val.empty() && rows >= 32)) {
        build_rsell(rows, cols, [&](int64_t r, std::vector<std::pair<int, qb_c128>>& o) {
            for (int k = 0; k < ndiag; k++) {
                const int d = order[k];
                const int64_t c = r + offsets[d];
                if (c < 0 || c >= cols) continue;
                const qb_c128 x = v[(size_t)d * cols + c];
                if (x.re == 0.0 && x.im == 0.0) continue;
                o.push_back({(int)c, x});
            }
        }, rs);
        const long long other = (long long)dh.val.size() * (use_diam ? 16 : 20) + (long long)dh.ent.size() * 8;
        if (format == 0) use_rsell = want_rsell(rs, rows, use_sell, (long long)sh.val.size(), other);
    }
    if (use_rsell) rc = finish_rsell(h, rs, rows, cols);
    else if (use_sell) rc = finish_sell(h, sh, rows, cols);
    else if (use_diam) rc = finish_diam(h, dh, rows, cols);
    else {
        // sparse diagonals: fall back to CSR built from the slices
        std::vector<std::vector<std::pair<int, qb_c128>>> rowsv(rows);
        for (int64_t sl = 0; sl + 1 < (int64_t)dh.slice_ptr.size(); sl++) {
            long long vb = dh.slice_vbase[sl];
            for (int e = dh.slice_ptr[sl]; e < dh.slice_ptr[sl + 1]; e++) {
                unsigned m = (unsigned)dh.ent[e].y;
                for (int lane = 0; lane < 32; lane++)
                    if (m >> lane & 1u) {
                        int64_t r = sl * 32 + lane;
                        rowsv[r].push_back({(int)(r + dh.ent[e].x), dh.val[vb++]});
                    }
            }
        }
        std::vector<qb_c128> vv; std::vector<int> cc, rp(1, 0);
        for (auto& rr : rowsv) { for (auto& q : rr) { cc.push_back(q.first); vv.push_back(q.second); } rp.push_back((int)cc.size()); }
        h->dev.fmt = QB_FMT_CSR; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols; h->dev.nnz = (long long)vv.size();
        if (!(rc = to_device(h, vv, &h->dev.val)) && !(rc = to_device(h, cc, &h->dev.col)))
            rc = to_device(h, rp, &h->dev.rowptr);
    }
    if (rc) { delete h; return rc; }
    *out = h;
    return QB_OK;
}

extern "C" int qb_op_info(qb_handle hh, int* fmt, int64_t* rows, int64_t* cols, int64_t* nnz,
                          int64_t* device_bytes) {
    QbOpH* h = qb_cast<QbOpH>(hh, QB_TAG_OP);
    if (!h) QB_FAIL(QB_E_TYPE, "not an operator handle");
    if (fmt) *fmt = h->dev.fmt;
    if (rows) *rows = h->dev.nrows;
    if (cols) *cols = h->dev.ncols;
    if (nnz) *nnz = h->dev.nnz;
    if (device_bytes) *device_bytes = h->device_bytes;
    return QB_OK;
}

// ------------------------------------------------------------------ kernels
// out[:, c] += scale * A @ X[:, c] for column-major X/out (ldx/ldo) or row-major
// (element (r,c) at r*sx_r + c*sx_c).
__global__ void __launch_bounds__(256)
qb_matmul_kernel(QbOpDev A, const double2* __restrict__ X, long long xs_r, long long xs_c,
                 double2* __restrict__ O, long long os_r, long long os_c, int ncols, double2 scale)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x * 8 + warp;
    const long long r = (long long)sl * 32 + lane;
    if ((long long)sl * 32 >= A.nrows) return;
    const bool active = r < A.nrows;
    for (int c = blockIdx.y; c < ncols; c += gridDim.y) {
        double2 q;
        if (xs_r == 1) q = qb_rowdot(A, sl, lane, r, active, X + c * xs_c);
        else {
            // strided state (C-ordered multi-column): generic gather
            q = make_double2(0.0, 0.0);
            if (active) {
                if (A.fmt == QB_FMT_CSR) {
                    const double2* val = reinterpret_cast<const double2*>(A.val);
                    for (int p = A.rowptr[r]; p < A.rowptr[r + 1]; p++)
                        qb_fma(q, val[p], X[(long long)A.col[p] * xs_r + c * xs_c]);
                } else if (A.fmt == QB_FMT_DENSE) {
                    const double2* a = reinterpret_cast<const double2*>(A.dense) + r;
                    for (int k = 0; k < A.ncols; k++)
                        qb_fma(q, a[(long long)k * A.nrows], X[(long long)k * xs_r + c * xs_c]);
                } else if (A.fmt == QB_FMT_SELL) {
                    const int s0 = A.slice_ptr[sl], w = A.slice_ptr[sl + 1] - s0;
                    const double2* val = reinterpret_cast<const double2*>(A.val) + ((size_t)s0 * 32 + lane);
                    const int* col = A.col + ((size_t)s0 * 32 + lane);
                    for (int k = 0; k < w; k++)
                        qb_fma(q, val[k * 32], X[(long long)col[k * 32] * xs_r + c * xs_c]);
                } else if (A.fmt == QB_FMT_RSELL) {
                    const double2* val = reinterpret_cast<const double2*>(A.val);
                    const int4 si = reinterpret_cast<const int4*>(A.sinfo)[sl];
                    for (int k = si.x; k < si.x + (si.y & 4095); k++) {
                        const QbSlotDesc d = A.sdesc[k];
                        const int cr = d.rule & QB_RS_COL_MASK;
                        const long long cc = cr == QB_RS_COL_ADD ? r + d.delta
                                           : cr == QB_RS_COL_XOR ? (r ^ (long long)d.delta)
                                                                 : A.col[(size_t)(si.w + d.cpos) * 32 + lane];
                        const double2 vv = (d.rule & QB_RS_VAL_CONST) ? make_double2(d.vre, d.vim)
                                                                      : val[(size_t)(si.z + d.vpos) * 32 + lane];
                        qb_fma(q, vv, X[cc * xs_r + c * xs_c]);
                    }
                }
            }
            if (A.fmt == QB_FMT_DIAM) {
                const int e0 = A.slice_ptr[sl], e1 = A.slice_ptr[sl + 1];
                long long vb = A.slice_vbase[sl];
                const unsigned lt = (1u << lane) - 1u;
                const int2* ent = reinterpret_cast<const int2*>(A.ent_off);
                const double2* val = reinterpret_cast<const double2*>(A.val);
                for (int e = e0; e < e1; e++) {
                    const int2 d = ent[e];
                    const unsigned m = (unsigned)d.y;
                    if ((m >> lane) & 1u)
                        qb_fma(q, val[vb + __popc(m & lt)], X[(r + d.x) * xs_r + c * xs_c]);
                    vb += __popc(m);
                }
            }
        }
        if (active) {
            double2* o = O + r * os_r + c * os_c;
            double2 v = *o;
            v.x += scale.x * q.x - scale.y * q.y;
            v.y += scale.x * q.y + scale.y * q.x;
            *o = v;
        }
    }
}

__global__ void qb_axpy_kernel(long long n, double2 a, const double2* __restrict__ x, double2* __restrict__ y) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 xv = x[i]; double2 yv = y[i];
        yv.x += a.x * xv.x - a.y * xv.y; yv.y += a.x * xv.y + a.y * xv.x;
        y[i] = yv;
    }
}
// strided variant: element (r, c) of x at r*xs_r + c*xs_c (C/F order mismatch, add.pyx:214-217)
__global__ void qb_axpy_strided_kernel(long long rows, long long cols, double2 a, const double2* __restrict__ x,
                                       long long xs_r, long long xs_c, double2* __restrict__ y,
                                       long long ys_r, long long ys_c) {
    const long long n = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i % rows, c = i / rows;
        const double2 xv = x[r * xs_r + c * xs_c]; double2* yp = y + r * ys_r + c * ys_c;
        double2 yv = *yp;
        yv.x += a.x * xv.x - a.y * xv.y; yv.y += a.x * xv.y + a.y * xv.x;
        *yp = yv;
    }
}
__global__ void qb_scal_kernel(long long n, double2 a, double2* __restrict__ x) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double2 v = x[i];
        x[i] = make_double2(a.x * v.x - a.y * v.y, a.x * v.y + a.y * v.x);
    }
}

// generic two-stage reduction: mode selects the per-element term; partials[grid][2]
enum { RED_NRM2 = 0, RED_WRMS = 1, RED_INNER = 2, RED_INNER_NOCONJ = 3, RED_TRACE_KET = 4 };
__global__ void __launch_bounds__(256)
qb_reduce_kernel(int mode, long long n, const double2* __restrict__ a, const double2* __restrict__ b,
                 double p0, double p1, long long aux, double* __restrict__ partials)
{
    double s0 = 0.0, s1 = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        if (mode == RED_NRM2) { const double2 v = a[i]; s0 += v.x * v.x + v.y * v.y; }
        else if (mode == RED_WRMS) {
            const double2 d = a[i], s = b[i];
            const double q = sqrt(d.x * d.x + d.y * d.y) / (p0 + p1 * sqrt(s.x * s.x + s.y * s.y));
            s0 += q * q;
        } else if (mode == RED_INNER) {
            const double2 x = a[i], y = b[i];
            s0 += x.x * y.x + x.y * y.y; s1 += x.x * y.y - x.y * y.x;
        } else if (mode == RED_INNER_NOCONJ) {
            const double2 x = a[i], y = b[i];
            s0 += x.x * y.x - x.y * y.y; s1 += x.x * y.y + x.y * y.x;
        } else {   // trace of a column-stacked n x n operator: elements i*(aux+1), i < aux
            if (i < aux) { const double2 v = a[i * (aux + 1)]; s0 += v.x; s1 += v.y; }
        }
    }
    __shared__ double sh[8][2];
    s0 = qb_warp_sum(s0); s1 = qb_warp_sum(s1);
    if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5][0] = s0; sh[threadIdx.x >> 5][1] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sh[w][threadIdx.x];
        partials[blockIdx.x * 2 + threadIdx.x] = s;
    }
}
__global__ void qb_reduce_final_kernel(int nblocks, const double* __restrict__ partials, double* __restrict__ out) {
    // single warp, fixed order
    double s0 = 0.0, s1 = 0.0;
    for (int i = threadIdx.x; i < nblocks; i += 32) { s0 += partials[2 * i]; s1 += partials[2 * i + 1]; }
    s0 = qb_warp_sum(s0); s1 = qb_warp_sum(s1);
    if (threadIdx.x == 0) { out[0] = s0; out[1] = s1; }
}

// <x|A|x> (ket) / sum_r (A x)[r] * w : fused SpMV + dot, per-block partials
__global__ void __launch_bounds__(256)
qb_expect_kernel(QbOpDev A, const double2* __restrict__ x, const double2* __restrict__ bra,
                 long long bra_stride, int conj_bra, double* __restrict__ partials)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sl = blockIdx.x * 8 + warp;
    const long long r = (long long)sl * 32 + lane;
    double s0 = 0.0, s1 = 0.0;
    if ((long long)sl * 32 < A.nrows) {
        const bool active = r < A.nrows;
        const double2 q = qb_rowdot(A, sl, lane, r, active, x);
        if (active) {
            double2 b = bra ? bra[r * bra_stride] : make_double2(1.0, 0.0);
            if (conj_bra) b.y = -b.y;
            s0 = b.x * q.x - b.y * q.y; s1 = b.x * q.y + b.y * q.x;
        }
    }
    __shared__ double sh[8][2];
    s0 = qb_warp_sum(s0); s1 = qb_warp_sum(s1);
    if (lane == 0) { sh[warp][0] = s0; sh[warp][1] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
        double s = 0.0;
        for (int w = 0; w < 8; w++) s += sh[w][threadIdx.x];
        partials[blockIdx.x * 2 + threadIdx.x] = s;
    }
}

// ------------------------------------------------------------------ host wrappers
namespace {
struct Scratch {
    double* d = nullptr; size_t cap = 0; double* h = nullptr;
    int ensure(size_t n) {
        if (!h && cudaMallocHost((void**)&h, 16) != cudaSuccess) return -1;
        if (n <= cap) return 0;
        if (d) cudaFree(d);
        if (cudaMalloc((void**)&d, n * sizeof(double)) != cudaSuccess) { d = nullptr; cap = 0; return -1; }
        cap = n; return 0;
    }
};
thread_local Scratch g_scr;

static int get_opdev(qb_handle op, QbOpDev* out) {
    if (QbOpH* o = qb_cast<QbOpH>(op, QB_TAG_OP)) { *out = o->dev; return QB_OK; }
    if (QbDenseH* dn = qb_cast<QbDenseH>(op, QB_TAG_DENSE)) {
        memset(out, 0, sizeof *out);
        if (!dn->fortran && dn->rows > 1 && dn->cols > 1)
            QB_FAIL(QB_E_TYPE, "dense left operands must be column-major (fortran)");
        out->fmt = QB_FMT_DENSE; out->nrows = (int)dn->rows; out->ncols = (int)dn->cols;
        out->nnz = dn->rows * dn->cols; out->dense = reinterpret_cast<const qb_c128*>(dn->d);
        return QB_OK;
    }
    QB_FAIL(QB_E_TYPE, "handle is not an operator");
}

static int reduce2(int mode, long long n, const double2* a, const double2* b, double p0, double p1,
                   long long aux, double out[2]) {
    int nb = (int)std::min<long long>(1184, std::max<long long>(1, (n + 255) / 256));
    if (g_scr.ensure((size_t)nb * 2 + 2)) QB_FAIL(QB_E_ALLOC, "scratch allocation failed");
    qb_reduce_kernel<<<nb, 256>>>(mode, n, a, b, p0, p1, aux, g_scr.d + 2);
    QB_LAUNCH_CHECK();
    qb_reduce_final_kernel<<<1, 32>>>(nb, g_scr.d + 2, g_scr.d);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaMemcpy(g_scr.h, g_scr.d, 16, cudaMemcpyDeviceToHost));
    out[0] = g_scr.h[0]; out[1] = g_scr.h[1];
    return QB_OK;
}
}  // namespace

extern "C" int qb_zgemm(qb_handle ah, qb_handle xh, double sre, double sim, qb_handle oh);

extern "C" int qb_matmul(qb_handle op, qb_handle xh, double sre, double sim, qb_handle outh) {
    QbOpDev A; int rc = get_opdev(op, &A); if (rc) return rc;
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* o = qb_cast<QbDenseH>(outh, QB_TAG_DENSE);
    // dense x multi-column column-major state: true ZGEMM on the FP64 tensor cores
    if (A.fmt == QB_FMT_DENSE && x && o && x->cols >= 8 && x->fortran && o->fortran &&
        A.ncols == x->rows && A.nrows == o->rows && x->cols == o->cols)
        return qb_zgemm(op, xh, sre, sim, outh);
    if (!x || !o) QB_FAIL(QB_E_TYPE, "matmul needs dense right operand and output");
    if (A.ncols != x->rows || A.nrows != o->rows || x->cols != o->cols)
        QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes (%d, %d) and (%lld, %lld)", A.nrows, A.ncols,
                (long long)x->rows, (long long)x->cols);
    if (A.nrows == 0 || x->cols == 0) return QB_OK;
    const long long xs_r = x->fortran || x->cols == 1 ? 1 : x->cols, xs_c = x->fortran || x->cols == 1 ? x->rows : 1;
    const long long os_r = o->fortran || o->cols == 1 ? 1 : o->cols, os_c = o->fortran || o->cols == 1 ? o->rows : 1;
    if (A.fmt == QB_FMT_KRON && xs_r != 1)
        QB_FAIL(QB_E_ARG, "Kronecker operators need a column-major (or single-column) right operand");
    dim3 grid((unsigned)((A.nrows + 255) / 256), (unsigned)std::min<int64_t>(x->cols, 65535));
    qb_matmul_kernel<<<grid, 256>>>(A, x->d, xs_r, xs_c, o->d, os_r, os_c, (int)x->cols, make_double2(sre, sim));
    QB_LAUNCH_CHECK();
    return QB_OK;
}

extern "C" int qb_axpy(qb_handle xh, double are, double aim, qb_handle yh) {
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* y = qb_cast<QbDenseH>(yh, QB_TAG_DENSE);
    if (!x || !y) QB_FAIL(QB_E_TYPE, "axpy needs dense handles");
    if (x->rows != y->rows || x->cols != y->cols)
        QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes (%lld, %lld) and (%lld, %lld)", (long long)x->rows,
                (long long)x->cols, (long long)y->rows, (long long)y->cols);
    const long long n = x->size();
    if (n == 0) return QB_OK;
    const int nb = (int)std::min<long long>(148 * 8, (n + 255) / 256);
    const bool same = x->fortran == y->fortran || x->rows == 1 || x->cols == 1;
    if (same) qb_axpy_kernel<<<nb, 256>>>(n, make_double2(are, aim), x->d, y->d);
    else qb_axpy_strided_kernel<<<nb, 256>>>(x->rows, x->cols, make_double2(are, aim), x->d,
            x->fortran ? 1 : x->cols, x->fortran ? x->rows : 1, y->d, y->fortran ? 1 : y->cols, y->fortran ? y->rows : 1);
    QB_LAUNCH_CHECK();
    return QB_OK;
}
extern "C" int qb_scal(qb_handle xh, double are, double aim) {
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    if (!x) QB_FAIL(QB_E_TYPE, "scal needs a dense handle");
    const long long n = x->size();
    if (n == 0) return QB_OK;
    qb_scal_kernel<<<(int)std::min<long long>(148 * 8, (n + 255) / 256), 256>>>(n, make_double2(are, aim), x->d);
    QB_LAUNCH_CHECK();
    return QB_OK;
}
extern "C" int qb_copy(qb_handle sh, qb_handle dh) {
    QbDenseH* s = qb_cast<QbDenseH>(sh, QB_TAG_DENSE);
    QbDenseH* d = qb_cast<QbDenseH>(dh, QB_TAG_DENSE);
    if (!s || !d) QB_FAIL(QB_E_TYPE, "copy needs dense handles");
    if (s->rows != d->rows || s->cols != d->cols) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    if (s->fortran != d->fortran && s->rows > 1 && s->cols > 1) {
        int rc = qb_zero(dh); if (rc) return rc;
        return qb_axpy(sh, 1.0, 0.0, dh);
    }
    QB_CUDA(cudaMemcpy(d->d, s->d, (size_t)s->size() * 16, cudaMemcpyDeviceToDevice));
    return QB_OK;
}
extern "C" int qb_zero(qb_handle xh) {
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    if (!x) QB_FAIL(QB_E_TYPE, "zero needs a dense handle");
    QB_CUDA(cudaMemset(x->d, 0, (size_t)x->size() * 16));
    return QB_OK;
}
extern "C" int qb_nrm2(qb_handle xh, double* out) {
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    if (!x || !out) QB_FAIL(QB_E_TYPE, "nrm2 needs a dense handle");
    if (x->size() == 0) { *out = 0.0; return QB_OK; }
    double r[2]; int rc = reduce2(RED_NRM2, x->size(), x->d, nullptr, 0, 0, 0, r); if (rc) return rc;
    *out = sqrt(r[0]);
    return QB_OK;
}
extern "C" int qb_wrms_error(qb_handle dh, qb_handle sh, double atol, double rtol, double* out) {
    QbDenseH* d = qb_cast<QbDenseH>(dh, QB_TAG_DENSE);
    QbDenseH* s = qb_cast<QbDenseH>(sh, QB_TAG_DENSE);
    if (!d || !s || !out) QB_FAIL(QB_E_TYPE, "wrms_error needs dense handles");
    if (d->rows != s->rows || d->cols != s->cols) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    if (d->fortran != s->fortran && d->rows > 1 && d->cols > 1) QB_FAIL(QB_E_TYPE, "wrms_error: mixed C/F order");
    if (d->size() == 0) { *out = 0.0; return QB_OK; }
    double r[2]; int rc = reduce2(RED_WRMS, d->size(), d->d, s->d, atol, rtol, 0, r); if (rc) return rc;
    *out = sqrt(r[0] / (double)d->size());
    return QB_OK;
}
extern "C" int qb_inner(qb_handle ah, qb_handle bh, int conj_a, double out[2]) {
    QbDenseH* a = qb_cast<QbDenseH>(ah, QB_TAG_DENSE);
    QbDenseH* b = qb_cast<QbDenseH>(bh, QB_TAG_DENSE);
    if (!a || !b || !out) QB_FAIL(QB_E_TYPE, "inner needs dense handles");
    if (a->size() != b->size()) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    if (a->size() == 0) { out[0] = out[1] = 0.0; return QB_OK; }
    return reduce2(conj_a ? RED_INNER : RED_INNER_NOCONJ, a->size(), a->d, b->d, 0, 0, 0, out);
}
extern "C" int qb_trace_oper_ket(qb_handle vh, double out[2]) {
    QbDenseH* v = qb_cast<QbDenseH>(vh, QB_TAG_DENSE);
    if (!v || !out) QB_FAIL(QB_E_TYPE, "trace_oper_ket needs a dense handle");
    const long long n2 = v->size();
    long long n = (long long)llround(sqrt((double)n2));
    if (n * n != n2 || (v->cols != 1 && v->rows != 1)) QB_FAIL(QB_E_SHAPE, "trace_oper_ket: not a stacked square operator");
    return reduce2(RED_TRACE_KET, n, v->d, nullptr, 0, 0, n, out);
}

static int expect_impl(const QbOpDev& A, const double2* x, const double2* bra, long long bra_stride,
                       int conj_bra, double acc[2]) {
    const int nb = (A.nrows + 255) / 256;
    if (g_scr.ensure((size_t)nb * 2 + 2)) QB_FAIL(QB_E_ALLOC, "scratch allocation failed");
    qb_expect_kernel<<<nb, 256>>>(A, x, bra, bra_stride, conj_bra, g_scr.d + 2);
    QB_LAUNCH_CHECK();
    qb_reduce_final_kernel<<<1, 32>>>(nb, g_scr.d + 2, g_scr.d);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaMemcpy(g_scr.h, g_scr.d, 16, cudaMemcpyDeviceToHost));
    acc[0] += g_scr.h[0]; acc[1] += g_scr.h[1];
    return QB_OK;
}
extern "C" int qb_expect_ket(qb_handle op, qb_handle xh, double out[2]) {
    QbOpDev A; int rc = get_opdev(op, &A); if (rc) return rc;
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    if (!x || !out) QB_FAIL(QB_E_TYPE, "expect needs a dense state");
    if (x->cols != 1 || A.ncols != x->rows || A.nrows != x->rows) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    out[0] = out[1] = 0.0;
    if (A.nrows == 0) return QB_OK;
    return expect_impl(A, x->d, x->d, 1, 1, out);
}
extern "C" int qb_expect_dm(qb_handle op, qb_handle rh, double out[2]) {
    // tr(A rho) = sum_c (A rho[:, c])[c]
    QbOpDev A; int rc = get_opdev(op, &A); if (rc) return rc;
    QbDenseH* rho = qb_cast<QbDenseH>(rh, QB_TAG_DENSE);
    if (!rho || !out) QB_FAIL(QB_E_TYPE, "expect needs a dense state");
    if (rho->rows != rho->cols || A.ncols != rho->rows || A.nrows != rho->rows) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    out[0] = out[1] = 0.0;
    if (A.nrows == 0) return QB_OK;
    // Y = A rho (n x n), then trace; uses matmul into a temporary
    qb_handle tmp = nullptr;
    rc = qb_dense_zeros(rho->rows, rho->cols, 1, &tmp); if (rc) return rc;
    rc = qb_matmul(op, rh, 1.0, 0.0, tmp);
    if (!rc) rc = reduce2(RED_TRACE_KET, rho->rows, static_cast<QbDenseH*>(tmp)->d, nullptr, 0, 0, rho->rows, out);
    qb_free(tmp);
    return rc;
}
extern "C" int qb_expect_super(qb_handle op, qb_handle vh, double out[2]) {
    // trace(unstack(A vec)) = sum over rows i*(n+1) of (A vec)[row]
    QbOpDev A; int rc = get_opdev(op, &A); if (rc) return rc;
    QbDenseH* v = qb_cast<QbDenseH>(vh, QB_TAG_DENSE);
    if (!v || !out) QB_FAIL(QB_E_TYPE, "expect_super needs a dense state");
    if (v->cols != 1 || A.ncols != v->rows || A.nrows != v->rows) QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes");
    out[0] = out[1] = 0.0;
    qb_handle tmp = nullptr;
    rc = qb_dense_zeros(v->rows, 1, 1, &tmp); if (rc) return rc;
    rc = qb_matmul(op, vh, 1.0, 0.0, tmp);
    if (!rc) rc = qb_trace_oper_ket(tmp, out);
    qb_free(tmp);
    return rc;
}
