// qb_build.cu -- assemble the Liouvillian superoperator on the device.
//
// Replaces the host assembly qutip.liouvillian does with a chain of scipy-style kron / add
// calls (reference: qutip/core/superoperator.py:116-142; data layer kron at
// qutip/core/data/kron.pyx, add at qutip/core/data/add.pyx):
//
//   L = -i (I (x) H) + i (conj(H) (x) I) + sum_k [ conj(C_k) (x) C_k - 1/2 I (x) C_k^+C_k
//                                                  - 1/2 (C_k^+C_k)^T (x) I ]
//
// on the column-stacked state (row index R = a*n + b for rho[b, a]).  With the n x n matrix
// A = -iH - 1/2 sum_k C_k^+C_k (assembled by the caller: n x n work) this is
//
//   L[(a,b),(a',b')] = d_aa' A[b,b'] + conj(A)[a,a'] d_bb' + sum_k conj(C_k[a,a']) C_k[b,b'],
//
// so every row of L is the merge of three short lists.  One thread per row: count the
// candidates, scan, write them, sort each row's run by column, sum duplicates, scan again and
// compact into canonical CSR (sorted, no duplicates).  The n^2 x n^2 operator never exists on
// the host unless a compressed format is asked for (the slice analysers of qb_diam.h run there).
#include <cub/device/device_scan.cuh>
#include <vector>
#include "qb_host.h"

namespace {
struct LvIn {
    const double2* av; const int* ac; const int* ap;      // A, n x n CSR
    const double2* cv; const int* cc; const int* cp;      // C_k stacked row-wise, (nstack*n) x n CSR
    int n, nstack;
};

__global__ void qb_lv_count_kernel(LvIn I, long long* __restrict__ ub) {
    const long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n2 = (long long)I.n * I.n;
    if (R >= n2) return;
    const int a = (int)(R / I.n), b = (int)(R % I.n);
    long long c = (I.ap[b + 1] - I.ap[b]) + (I.ap[a + 1] - I.ap[a]);
    for (int k = 0; k < I.nstack; k++) {
        const int* p = I.cp + (size_t)k * I.n;
        c += (long long)(p[a + 1] - p[a]) * (p[b + 1] - p[b]);
    }
    ub[R] = c;
}

// writes the candidates of row R at tcol/tval[off[R]..], sorts them by column (stable insertion
// sort: the runs are a few dozen entries), sums duplicates in place, tidies, stores the count
__global__ void qb_lv_fill_kernel(LvIn I, const long long* __restrict__ off, int* __restrict__ tcol,
                                  double2* __restrict__ tval, int* __restrict__ cnt, double tol) {
    const long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n2 = (long long)I.n * I.n;
    if (R >= n2) return;
    const int n = I.n, a = (int)(R / n), b = (int)(R % n);
    int* col = tcol + off[R];
    double2* val = tval + off[R];
    int m = 0;
    for (int p = I.ap[b]; p < I.ap[b + 1]; p++) { col[m] = a * n + I.ac[p]; val[m] = I.av[p]; m++; }
    for (int p = I.ap[a]; p < I.ap[a + 1]; p++) {
        const double2 v = I.av[p];
        col[m] = I.ac[p] * n + b; val[m] = make_double2(v.x, -v.y); m++;
    }
    for (int k = 0; k < I.nstack; k++) {
        const int* rp = I.cp + (size_t)k * n;
        for (int p = rp[a]; p < rp[a + 1]; p++) {
            const double2 ca = I.cv[p];
            const int ja = I.cc[p];
            for (int q = rp[b]; q < rp[b + 1]; q++) {
                const double2 cb = I.cv[q];
                col[m] = ja * n + I.cc[q];
                // conj(ca) * cb
                val[m] = make_double2(ca.x * cb.x + ca.y * cb.y, ca.x * cb.y - ca.y * cb.x);
                m++;
            }
        }
    }
    for (int i = 1; i < m; i++) {
        const int c = col[i]; const double2 v = val[i];
        int j = i - 1;
        while (j >= 0 && col[j] > c) { col[j + 1] = col[j]; val[j + 1] = val[j]; j--; }
        col[j + 1] = c; val[j + 1] = v;
    }
    int u = 0;
    for (int i = 0; i < m; i++) {
        if (u > 0 && col[u - 1] == col[i]) { val[u - 1].x += val[i].x; val[u - 1].y += val[i].y; }
        else { col[u] = col[i]; val[u] = val[i]; u++; }
    }
    // the reference's accumulator (core/data/csr.pxd:88-113, used by add_csr): components below
    // the tidy-up tolerance are zeroed, entries that are exactly zero afterwards are dropped
    int w = 0;
    for (int i = 0; i < u; i++) {
        double2 v = val[i];
        if (fabs(v.x) < tol) v.x = 0.0;
        if (fabs(v.y) < tol) v.y = 0.0;
        if (v.x != 0.0 || v.y != 0.0) { col[w] = col[i]; val[w] = v; w++; }
    }
    cnt[R] = w;
}

__global__ void qb_lv_compact_kernel(long long n2, const long long* __restrict__ off,
                                     const int* __restrict__ rowptr, const int* __restrict__ tcol,
                                     const double2* __restrict__ tval, int* __restrict__ col,
                                     double2* __restrict__ val) {
    // one warp per row: coalesced copy of the row's unique run
    const long long R = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (R >= n2) return;
    const int p0 = rowptr[R], m = rowptr[R + 1] - p0;
    const long long s = off[R];
    for (int i = lane; i < m; i += 32) { col[p0 + i] = tcol[s + i]; val[p0 + i] = tval[s + i]; }
}


// ---- CSR -> DIAM on the device (the host analyser is qbdiam::emit_slice): one warp per 32-row
// slice merges its 32 sorted rows by diagonal offset -- the minimum offset over the lanes is
// the next entry, the lanes that hold it form its mask, their values are packed in lane order.
// Canonical CSR (sorted columns, no duplicates) only.
template <bool FILL>
__global__ void qb_diam_slices_kernel(const int* __restrict__ rowptr, const int* __restrict__ col,
                                      const double2* __restrict__ val, int nrows, int nslices,
                                      int* __restrict__ ent_cnt, long long* __restrict__ val_cnt,
                                      const int* __restrict__ slice_ptr, const long long* __restrict__ slice_vbase,
                                      int2* __restrict__ ent, double2* __restrict__ oval) {
    const int sl = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (sl >= nslices) return;
    const long long r = (long long)sl * 32 + lane;
    int p = 0, pe = 0;
    if (r < nrows) { p = rowptr[r]; pe = rowptr[r + 1]; }
    const unsigned lt = (1u << lane) - 1u;
    int ne = 0;
    long long nv = 0;
    const int e0 = FILL ? slice_ptr[sl] : 0;
    const long long v0 = FILL ? slice_vbase[sl] : 0;
    for (;;) {
        const int off = p < pe ? col[p] - (int)r : 0x7fffffff;
        int m = off;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, d));
        if (m == 0x7fffffff) break;
        const bool hit = off == m;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (FILL) {
            if (lane == 0) ent[e0 + ne] = make_int2(m, (int)mask);
            if (hit) oval[v0 + nv + __popc(mask & lt)] = val[p];
        }
        if (hit) p++;
        ne++; nv += __popc(mask);
    }
    if (!FILL && lane == 0) { ent_cnt[sl] = ne; val_cnt[sl] = nv; }
}


// ---- Kronecker product C = A (x) B in canonical CSR (core/data/kron.pyx: kron_csr): row
// ia * rowsB + ib holds nnzA(ia) * nnzB(ib) entries, column ca * colsB + cb, value a * b --
// sorted and duplicate-free when A and B are.  One warp per output row.
__global__ void qb_kron_rowptr_kernel(const int* __restrict__ ap, const int* __restrict__ bp,
                                      long long rows, int rows_b, long long* __restrict__ cnt) {
    const long long R = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (R >= rows) return;
    const int ia = (int)(R / rows_b), ib = (int)(R % rows_b);
    cnt[R] = (long long)(ap[ia + 1] - ap[ia]) * (bp[ib + 1] - bp[ib]);
}
__global__ void qb_kron_fill_kernel(const double2* __restrict__ av, const int* __restrict__ ac,
                                    const int* __restrict__ ap, const double2* __restrict__ bv,
                                    const int* __restrict__ bc, const int* __restrict__ bp,
                                    long long rows, int rows_b, int cols_b,
                                    const long long* __restrict__ off, int* __restrict__ rowptr,
                                    int* __restrict__ col, double2* __restrict__ val) {
    const long long R = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (R > rows) return;
    if (R == rows) { if (lane == 0) rowptr[rows] = (int)off[rows]; return; }
    const int ia = (int)(R / rows_b), ib = (int)(R % rows_b);
    const int a0 = ap[ia], na = ap[ia + 1] - a0, b0 = bp[ib], nb = bp[ib + 1] - b0;
    const long long o = off[R];
    if (lane == 0) rowptr[R] = (int)o;
    const int m = na * nb;
    for (int i = lane; i < m; i += 32) {
        const int pa = a0 + i / nb, pb = b0 + i % nb;
        const double2 a = av[pa], b = bv[pb];
        col[o + i] = ac[pa] * cols_b + bc[pb];
        // unfused products, as the reference's scalar C multiply (no FMA contraction): bit-equal
        val[o + i] = make_double2(__dsub_rn(__dmul_rn(a.x, b.x), __dmul_rn(a.y, b.y)),
                                  __dadd_rn(__dmul_rn(a.x, b.y), __dmul_rn(a.y, b.x)));
    }
}

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
};
static int check_csr(const int32_t* col, const int32_t* rowptr, int64_t rows, int64_t cols, int64_t nnz,
                     const char* what) {
    if (rowptr[0] != 0 || rowptr[rows] != nnz) QB_FAIL(QB_E_ARG, "%s: row_index does not match nnz", what);
    for (int64_t r = 0; r < rows; r++)
        if (rowptr[r + 1] < rowptr[r]) QB_FAIL(QB_E_ARG, "%s: row_index not monotone", what);
    for (int64_t p = 0; p < nnz; p++)
        if (col[p] < 0 || col[p] >= cols) QB_FAIL(QB_E_ARG, "%s: column index out of range", what);
    return QB_OK;
}
template <class T> static int upload(DevBuf& b, const T* src, size_t count) {
    QB_CUDA(cudaMalloc(&b.p, std::max<size_t>(16, count * sizeof(T))));
    if (count) QB_CUDA(cudaMemcpy(b.p, src, count * sizeof(T), cudaMemcpyHostToDevice));
    return QB_OK;
}
}  // namespace

static int qb_diam_from_device_csr(const double2* val, const int* col, const int* rowptr, int64_t rows,
                                   int64_t cols, int64_t nnz, qb_handle* out);

extern "C" int qb_liouvillian_build(const void* a_data, const int32_t* a_col, const int32_t* a_rowptr,
                                    int64_t a_nnz, const void* c_data, const int32_t* c_col,
                                    const int32_t* c_rowptr, int64_t c_nnz, int64_t n, int64_t nstack,
                                    double tol, int format, qb_handle* out) {
    if (!out || n <= 0 || nstack < 0 || !a_rowptr || a_nnz < 0 || c_nnz < 0 ||
        (a_nnz > 0 && (!a_data || !a_col)) || (nstack > 0 && !c_rowptr) || (c_nnz > 0 && (!c_data || !c_col)))
        QB_FAIL(QB_E_ARG, "bad Liouvillian arguments");
    if (n > 46340) QB_FAIL(QB_E_ARG, "n*n exceeds int32 row indices");
    if (!(tol >= 0.0)) QB_FAIL(QB_E_ARG, "tidy-up tolerance must be >= 0");
    if (format != 0 && format != 1 && format != 2 && format != 3 && format != 5)
        QB_FAIL(QB_E_ARG, "unknown operator format %d", format);
    int rc;
    if ((rc = check_csr(a_col, a_rowptr, n, n, a_nnz, "A"))) return rc;
    if (nstack > 0 && (rc = check_csr(c_col, c_rowptr, n * nstack, n, c_nnz, "C"))) return rc;
    static const int32_t zero_ptr[1] = {0};
    DevBuf av, ac, ap, cv, cc, cp;
    if ((rc = upload(av, static_cast<const double2*>(a_data), (size_t)a_nnz)) ||
        (rc = upload(ac, a_col, (size_t)a_nnz)) || (rc = upload(ap, a_rowptr, (size_t)n + 1)) ||
        (rc = upload(cv, static_cast<const double2*>(c_data), (size_t)c_nnz)) ||
        (rc = upload(cc, c_col, (size_t)c_nnz)) ||
        (rc = upload(cp, nstack > 0 ? c_rowptr : zero_ptr, nstack > 0 ? (size_t)(n * nstack) + 1 : 1)))
        return rc;
    LvIn I{av.as<double2>(), ac.as<int>(), ap.as<int>(), cv.as<double2>(), cc.as<int>(), cp.as<int>(),
           (int)n, (int)nstack};
    const long long n2 = (long long)n * n;
    const int threads = 128;
    const unsigned blocks = (unsigned)((n2 + threads - 1) / threads);

    DevBuf ub, off, cnt, scan_tmp;
    QB_CUDA(cudaMalloc(&ub.p, (size_t)(n2 + 1) * sizeof(long long)));
    QB_CUDA(cudaMalloc(&off.p, (size_t)(n2 + 1) * sizeof(long long)));
    QB_CUDA(cudaMemset(ub.p, 0, (size_t)(n2 + 1) * sizeof(long long)));
    qb_lv_count_kernel<<<blocks, threads>>>(I, ub.as<long long>());
    QB_LAUNCH_CHECK();
    size_t tmp_bytes = 0, tmp2 = 0;
    QB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ub.as<long long>(), off.as<long long>(), n2 + 1));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp2, (int*)nullptr, (int*)nullptr, n2 + 1));
    tmp_bytes = std::max(tmp_bytes, tmp2);
    QB_CUDA(cudaMalloc(&scan_tmp.p, std::max<size_t>(16, tmp_bytes)));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, ub.as<long long>(), off.as<long long>(), n2 + 1));
    g_qb_launches++;
    long long total_ub = 0;
    QB_CUDA(cudaMemcpy(&total_ub, off.as<long long>() + n2, sizeof total_ub, cudaMemcpyDeviceToHost));
    if (total_ub > 0x7fffffffll) QB_FAIL(QB_E_ARG, "Liouvillian too large for int32 indices");

    DevBuf tcol, tval, rowptr;
    QB_CUDA(cudaMalloc(&tcol.p, std::max<size_t>(16, (size_t)total_ub * sizeof(int))));
    QB_CUDA(cudaMalloc(&tval.p, std::max<size_t>(16, (size_t)total_ub * sizeof(double2))));
    QB_CUDA(cudaMalloc(&cnt.p, (size_t)(n2 + 1) * sizeof(int)));
    QB_CUDA(cudaMalloc(&rowptr.p, (size_t)(n2 + 1) * sizeof(int)));
    QB_CUDA(cudaMemset(cnt.p, 0, (size_t)(n2 + 1) * sizeof(int)));
    qb_lv_fill_kernel<<<blocks, threads>>>(I, off.as<long long>(), tcol.as<int>(), tval.as<double2>(),
                                           cnt.as<int>(), tol);
    QB_LAUNCH_CHECK();
    QB_CUDA(cub::DeviceScan::ExclusiveSum(scan_tmp.p, tmp_bytes, cnt.as<int>(), rowptr.as<int>(), n2 + 1));
    g_qb_launches++;
    int nnz = 0;
    QB_CUDA(cudaMemcpy(&nnz, rowptr.as<int>() + n2, sizeof nnz, cudaMemcpyDeviceToHost));

    DevBuf col, val;
    QB_CUDA(cudaMalloc(&col.p, std::max<size_t>(16, (size_t)nnz * sizeof(int))));
    QB_CUDA(cudaMalloc(&val.p, std::max<size_t>(16, (size_t)nnz * sizeof(double2))));
    {
        const unsigned wblocks = (unsigned)((n2 * 32 + 255) / 256);
        qb_lv_compact_kernel<<<wblocks, 256>>>(n2, off.as<long long>(), rowptr.as<int>(), tcol.as<int>(),
                                               tval.as<double2>(), col.as<int>(), val.as<double2>());
        QB_LAUNCH_CHECK();
    }
    QB_CUDA(cudaDeviceSynchronize());

    if (format == 1) {                     // CSR: the device arrays are the operator
        QbOpH* h = new QbOpH();
        h->dev.fmt = QB_FMT_CSR; h->dev.nrows = (int)n2; h->dev.ncols = (int)n2; h->dev.nnz = nnz;
        h->dev.val = val.as<qb_c128>(); h->dev.col = col.as<int>(); h->dev.rowptr = rowptr.as<int>();
        h->owned = {val.p, col.p, rowptr.p};
        h->device_bytes = (int64_t)nnz * 20 + (n2 + 1) * 4;
        val.p = col.p = rowptr.p = nullptr;
        *out = h;
        return QB_OK;
    }
    if (format == 2)                       // diagonal-masked slices: converted on the device too
        return qb_diam_from_device_csr(val.as<double2>(), col.as<int>(), rowptr.as<int>(), n2, n2, nnz, out);
    // the other compressed formats: the padding / column-rule analysers run on the host
    std::vector<qb_c128> hv((size_t)nnz);
    std::vector<int32_t> hc((size_t)nnz), hp((size_t)n2 + 1);
    if (nnz) {
        QB_CUDA(cudaMemcpy(hv.data(), val.p, (size_t)nnz * sizeof(double2), cudaMemcpyDeviceToHost));
        QB_CUDA(cudaMemcpy(hc.data(), col.p, (size_t)nnz * sizeof(int), cudaMemcpyDeviceToHost));
    }
    QB_CUDA(cudaMemcpy(hp.data(), rowptr.p, (size_t)(n2 + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    return qb_csr_upload(hv.data(), hc.data(), hp.data(), n2, n2, nnz, format, out);
}

extern "C" int qb_op_csr_download(qb_handle hh, void* data, int32_t* col, int32_t* rowptr) {
    QbOpH* h = qb_cast<QbOpH>(hh, QB_TAG_OP);
    if (!h) QB_FAIL(QB_E_TYPE, "qb_op_csr_download: not an operator handle");
    if (h->dev.fmt != QB_FMT_CSR) QB_FAIL(QB_E_TYPE, "qb_op_csr_download: operator is not stored as CSR");
    if (!rowptr || (h->dev.nnz > 0 && (!data || !col))) QB_FAIL(QB_E_ARG, "null output buffer");
    QB_CUDA(cudaSetDevice(h->device));
    const size_t nnz = (size_t)h->dev.nnz;
    if (nnz) {
        QB_CUDA(cudaMemcpy(data, h->dev.val, nnz * sizeof(double2), cudaMemcpyDeviceToHost));
        QB_CUDA(cudaMemcpy(col, h->dev.col, nnz * sizeof(int), cudaMemcpyDeviceToHost));
    }
    QB_CUDA(cudaMemcpy(rowptr, h->dev.rowptr, ((size_t)h->dev.nrows + 1) * sizeof(int), cudaMemcpyDeviceToHost));
    return QB_OK;
}

// device CSR arrays -> DIAM operator, everything on the device
static int qb_diam_from_device_csr(const double2* val, const int* col, const int* rowptr, int64_t rows,
                                   int64_t cols, int64_t nnz, qb_handle* out) {
    const int nslices = (int)((rows + 31) / 32);
    DevBuf ecnt, vcnt, sptr, vbase, tmp;
    QB_CUDA(cudaMalloc(&ecnt.p, (size_t)(nslices + 1) * sizeof(int)));
    QB_CUDA(cudaMalloc(&vcnt.p, (size_t)(nslices + 1) * sizeof(long long)));
    QB_CUDA(cudaMalloc(&sptr.p, (size_t)(nslices + 1) * sizeof(int)));
    QB_CUDA(cudaMalloc(&vbase.p, (size_t)(nslices + 1) * sizeof(long long)));
    QB_CUDA(cudaMemset(ecnt.p, 0, (size_t)(nslices + 1) * sizeof(int)));
    QB_CUDA(cudaMemset(vcnt.p, 0, (size_t)(nslices + 1) * sizeof(long long)));
    const unsigned blocks = (unsigned)(((long long)nslices * 32 + 255) / 256);
    if (nslices > 0) {
        qb_diam_slices_kernel<false><<<blocks, 256>>>(rowptr, col, val, (int)rows, nslices, ecnt.as<int>(),
                                                      vcnt.as<long long>(), nullptr, nullptr, nullptr, nullptr);
        QB_LAUNCH_CHECK();
    }
    size_t t1 = 0, t2 = 0;
    QB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t1, ecnt.as<int>(), sptr.as<int>(), nslices + 1));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t2, vcnt.as<long long>(), vbase.as<long long>(), nslices + 1));
    QB_CUDA(cudaMalloc(&tmp.p, std::max<size_t>(16, std::max(t1, t2))));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t1, ecnt.as<int>(), sptr.as<int>(), nslices + 1));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, t2, vcnt.as<long long>(), vbase.as<long long>(), nslices + 1));
    g_qb_launches += 2;
    int nent = 0;
    long long nval = 0;
    QB_CUDA(cudaMemcpy(&nent, sptr.as<int>() + nslices, sizeof nent, cudaMemcpyDeviceToHost));
    QB_CUDA(cudaMemcpy(&nval, vbase.as<long long>() + nslices, sizeof nval, cudaMemcpyDeviceToHost));
    if (nval != nnz) QB_FAIL(QB_E_STATE, "DIAM conversion lost entries (CSR not canonical?)");
    DevBuf ent, oval;
    QB_CUDA(cudaMalloc(&ent.p, std::max<size_t>(16, (size_t)nent * sizeof(int2))));
    QB_CUDA(cudaMalloc(&oval.p, std::max<size_t>(16, (size_t)nval * sizeof(double2))));
    if (nslices > 0) {
        qb_diam_slices_kernel<true><<<blocks, 256>>>(rowptr, col, val, (int)rows, nslices, nullptr, nullptr,
                                                     sptr.as<int>(), vbase.as<long long>(), ent.as<int2>(),
                                                     oval.as<double2>());
        QB_LAUNCH_CHECK();
    }
    QB_CUDA(cudaDeviceSynchronize());
    QbOpH* h = new QbOpH();
    h->dev.fmt = QB_FMT_DIAM; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols; h->dev.nnz = nval;
    h->avg_lanes = nent ? (double)nval / (double)nent : 0.0;
    h->dev.slice_ptr = sptr.as<int>();
    h->dev.ent_off = ent.as<int>();
    h->dev.ent_mask = nullptr;
    h->dev.slice_vbase = vbase.as<long long>();
    h->dev.val = oval.as<qb_c128>();
    h->owned = {sptr.p, ent.p, vbase.p, oval.p};
    h->device_bytes = (int64_t)std::max<size_t>(16, (size_t)(nslices + 1) * sizeof(int)) +
                      (int64_t)std::max<size_t>(16, (size_t)nent * sizeof(int2)) +
                      (int64_t)std::max<size_t>(16, (size_t)(nslices + 1) * sizeof(long long)) +
                      (int64_t)std::max<size_t>(16, (size_t)nval * sizeof(double2));
    h->dev.pad_ = h->device_bytes > (int64_t)96 << 20 ? 1 : 0;     // streamed, as finish_diam decides
    sptr.p = ent.p = vbase.p = oval.p = nullptr;
    *out = h;
    return QB_OK;
}

// Format conversion of a CSR-format operator (the reference's `_data.to(Dia, CSR)` family,
// core/data/convert.pyx:208-329 + csr.pyx:710 / dia.pyx:364): DIAM is produced on the device,
// the rule / padding analysers of the other formats run on the host.
extern "C" int qb_op_convert(qb_handle hh, int format, qb_handle* out) {
    QbOpH* h = qb_cast<QbOpH>(hh, QB_TAG_OP);
    if (!h || !out) QB_FAIL(QB_E_TYPE, "qb_op_convert: not an operator handle");
    if (h->dev.fmt != QB_FMT_CSR) QB_FAIL(QB_E_TYPE, "qb_op_convert: source operator is not stored as CSR");
    QB_CUDA(cudaSetDevice(h->device));
    if (format == 2)
        return qb_diam_from_device_csr(reinterpret_cast<const double2*>(h->dev.val), h->dev.col, h->dev.rowptr,
                                       h->dev.nrows, h->dev.ncols, h->dev.nnz, out);
    if (format != 0 && format != 1 && format != 3 && format != 5) QB_FAIL(QB_E_ARG, "unknown operator format %d", format);
    const size_t nnz = (size_t)h->dev.nnz;
    std::vector<qb_c128> hv(nnz);
    std::vector<int32_t> hc(nnz), hp((size_t)h->dev.nrows + 1);
    if (nnz) {
        QB_CUDA(cudaMemcpy(hv.data(), h->dev.val, nnz * sizeof(double2), cudaMemcpyDeviceToHost));
        QB_CUDA(cudaMemcpy(hc.data(), h->dev.col, nnz * sizeof(int), cudaMemcpyDeviceToHost));
    }
    QB_CUDA(cudaMemcpy(hp.data(), h->dev.rowptr, hp.size() * sizeof(int), cudaMemcpyDeviceToHost));
    return qb_csr_upload(hv.data(), hc.data(), hp.data(), h->dev.nrows, h->dev.ncols, (int64_t)nnz, format, out);
}

// Kronecker product of two CSR matrices, assembled on the device (core/data/kron.pyx kron_csr;
// format as qb_csr_upload: 1 keeps the device CSR, 2 converts on the device, others via the host).
extern "C" int qb_kron_build(const void* a_data, const int32_t* a_col, const int32_t* a_rowptr,
                             int64_t a_rows, int64_t a_cols, const void* b_data, const int32_t* b_col,
                             const int32_t* b_rowptr, int64_t b_rows, int64_t b_cols, int format,
                             qb_handle* out) {
    if (!out || !a_rowptr || !b_rowptr || a_rows < 0 || a_cols < 0 || b_rows < 0 || b_cols < 0)
        QB_FAIL(QB_E_ARG, "bad Kronecker product arguments");
    if (format != 0 && format != 1 && format != 2 && format != 3 && format != 5)
        QB_FAIL(QB_E_ARG, "unknown operator format %d", format);
    const int64_t a_nnz = a_rowptr[a_rows], b_nnz = b_rowptr[b_rows];
    if ((a_nnz > 0 && (!a_data || !a_col)) || (b_nnz > 0 && (!b_data || !b_col)))
        QB_FAIL(QB_E_ARG, "bad Kronecker product arguments");
    const long long rows = (long long)a_rows * b_rows, cols = (long long)a_cols * b_cols;
    const long long nnz = (long long)a_nnz * b_nnz;
    if (rows > 0x7ffffffeLL || cols > 0x7fffffffLL || nnz > 0x7fffffffLL)
        QB_FAIL(QB_E_ARG, "Kronecker product too large for int32 indices");
    int rc;
    if ((rc = check_csr(a_col, a_rowptr, a_rows, a_cols, a_nnz, "A")) ||
        (rc = check_csr(b_col, b_rowptr, b_rows, b_cols, b_nnz, "B"))) return rc;
    DevBuf av, ac, ap, bv, bc, bp, cnt, off, tmp, rowptr, col, val;
    if ((rc = upload(av, static_cast<const double2*>(a_data), (size_t)a_nnz)) ||
        (rc = upload(ac, a_col, (size_t)a_nnz)) || (rc = upload(ap, a_rowptr, (size_t)a_rows + 1)) ||
        (rc = upload(bv, static_cast<const double2*>(b_data), (size_t)b_nnz)) ||
        (rc = upload(bc, b_col, (size_t)b_nnz)) || (rc = upload(bp, b_rowptr, (size_t)b_rows + 1)))
        return rc;
    QB_CUDA(cudaMalloc(&cnt.p, (size_t)(rows + 1) * sizeof(long long)));
    QB_CUDA(cudaMalloc(&off.p, (size_t)(rows + 1) * sizeof(long long)));
    QB_CUDA(cudaMemset(cnt.p, 0, (size_t)(rows + 1) * sizeof(long long)));
    if (rows > 0 && b_rows > 0) {
        qb_kron_rowptr_kernel<<<(unsigned)((rows + 255) / 256), 256>>>(ap.as<int>(), bp.as<int>(), rows,
                                                                      (int)b_rows, cnt.as<long long>());
        QB_LAUNCH_CHECK();
    }
    size_t tb = 0;
    QB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, cnt.as<long long>(), off.as<long long>(), rows + 1));
    QB_CUDA(cudaMalloc(&tmp.p, std::max<size_t>(16, tb)));
    QB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, cnt.as<long long>(), off.as<long long>(), rows + 1));
    g_qb_launches++;
    QB_CUDA(cudaMalloc(&rowptr.p, (size_t)(rows + 1) * sizeof(int)));
    QB_CUDA(cudaMalloc(&col.p, std::max<size_t>(16, (size_t)nnz * sizeof(int))));
    QB_CUDA(cudaMalloc(&val.p, std::max<size_t>(16, (size_t)nnz * sizeof(double2))));
    qb_kron_fill_kernel<<<(unsigned)(((rows + 1) * 32 + 255) / 256), 256>>>(
        av.as<double2>(), ac.as<int>(), ap.as<int>(), bv.as<double2>(), bc.as<int>(), bp.as<int>(), rows,
        (int)std::max<int64_t>(1, b_rows), (int)b_cols, off.as<long long>(), rowptr.as<int>(), col.as<int>(),
        val.as<double2>());
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaDeviceSynchronize());
    if (format == 2)
        return qb_diam_from_device_csr(val.as<double2>(), col.as<int>(), rowptr.as<int>(), rows, cols, nnz, out);
    QbOpH* h = new QbOpH();
    h->dev.fmt = QB_FMT_CSR; h->dev.nrows = (int)rows; h->dev.ncols = (int)cols; h->dev.nnz = nnz;
    h->dev.val = val.as<qb_c128>(); h->dev.col = col.as<int>(); h->dev.rowptr = rowptr.as<int>();
    h->owned = {val.p, col.p, rowptr.p};
    h->device_bytes = (int64_t)nnz * 20 + (rows + 1) * 4;
    val.p = col.p = rowptr.p = nullptr;
    if (format == 1) { *out = h; return QB_OK; }
    rc = qb_op_convert(h, format, out);
    delete h;
    return rc;
}
