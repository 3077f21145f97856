// qb_dense.cu -- dense batched path: Z[:, j] = A @ x_j for a block of state vectors as a
// true complex128 GEMM on the FP64 tensor cores (DMMA, mma.sync m8n8k4 f64).
//
// Replaces the reference's zgemm call for a dense operator times a multi-column state
// (core/data/matmul.pyx:329-346) and, inside the engine, the per-trajectory zgemv of a dense
// H_eff: all trajectory slots are multiplied in one launch.
// tcgen05 has no FP64 kind, so the FP64 tensor path on sm_100a is the warp-level DMMA.
//
// Complex product as four real MMAs on split re/im planes staged in shared memory:
//   Cr += Ar*Xr + (-Ai)*Xi ;  Ci += Ar*Xi + Ai*Xr
// CTA tile 64 (rows) x 64 (columns/trajectories), K step 16, 8 warps as 4 (M) x 2 (N); each
// warp owns 16 x 32 = 2 x 4 DMMA tiles (64 accumulator registers).  The next K step's A and X
// tiles are fetched into registers while the current one is multiplied (software double
// buffering: the global loads overlap 64 DMMA per warp instead of preceding them).  The columns of X and Z
// are addressed through per-column pointers, because inside the engine every trajectory's
// vector lives in its own (relabelled) slot of the state pool.
#include "qb_host.h"

#define ZG_BM 64
#define ZG_BN 64
#define ZG_BK 16
#define ZG_LD 68      // (2*LD) % 32 == 8 -> conflict-free fragment loads

__device__ __forceinline__ void qb_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// A: column-major M x K (ld = lda).  xcols[j] -> K complex, zcols[j] -> M complex.
// If xcols == nullptr the columns are X + j*ldx / Z + j*ldz (plain column-major matrices).
// accumulate != 0: Z += alpha * A X, else Z = alpha * A X.
__global__ void __launch_bounds__(256)
qb_zgemm_dmma_kernel(const double2* __restrict__ A, int M, int K, long long lda,
                     const double2* const* __restrict__ xcols, const double2* __restrict__ X,
                     long long ldx, double2* const* __restrict__ zcols, double2* __restrict__ Z,
                     long long ldz, int ncols, double2 alpha, int accumulate)
{
    __shared__ double As_re[ZG_BK][ZG_LD], As_im[ZG_BK][ZG_LD];
    __shared__ double Xs_re[ZG_BK][ZG_LD], Xs_im[ZG_BK][ZG_LD];
    __shared__ const double2* s_x[ZG_BN];
    __shared__ double2* s_z[ZG_BN];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;
    const int m0 = blockIdx.x * ZG_BM, n0 = blockIdx.y * ZG_BN;
    if (tid < ZG_BN) {
        const int j = n0 + tid;
        const double2* xp = nullptr; double2* zp = nullptr;
        if (j < ncols) {
            xp = xcols ? xcols[j] : X + (long long)j * ldx;
            zp = zcols ? zcols[j] : Z + (long long)j * ldz;
        }
        s_x[tid] = xp; s_z[tid] = zp;
    }
    __syncthreads();

    double acc_re[2][4][2], acc_im[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { acc_re[i][j][0] = acc_re[i][j][1] = 0.0; acc_im[i][j][0] = acc_im[i][j][1] = 0.0; }

    // per thread 4 elements of the A tile (64 rows x 16 k) and 4 of the X tile (16 k x 64 cols)
    double2 pa[4], px[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + i * 256;
            const int row = e & 63, kk = e >> 6;
            pa[i] = make_double2(0.0, 0.0);
            if (m0 + row < M && k0 + kk < K) pa[i] = A[(long long)(m0 + row) + (long long)(k0 + kk) * lda];
            const int kx = e & 15, col = e >> 4;
            px[i] = make_double2(0.0, 0.0);
            const double2* xp = s_x[col];
            if (xp && k0 + kx < K) px[i] = xp[k0 + kx];
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += ZG_BK) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int e = tid + i * 256;
            As_re[e >> 6][e & 63] = pa[i].x; As_im[e >> 6][e & 63] = pa[i].y;
            Xs_re[e & 15][e >> 4] = px[i].x; Xs_im[e & 15][e >> 4] = px[i].y;
        }
        __syncthreads();
        if (k0 + ZG_BK < K) fetch(k0 + ZG_BK);          // in flight during the multiply
#pragma unroll
        for (int ks = 0; ks < ZG_BK; ks += 4) {
            const int kk = ks + (lane & 3);
            double a_re[2], a_im[2], b_re[4], b_im[4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) {
                const int row = wm * 16 + mt * 8 + (lane >> 2);
                a_re[mt] = As_re[kk][row]; a_im[mt] = As_im[kk][row];
            }
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                const int col = wn * 32 + nt * 8 + (lane >> 2);
                b_re[nt] = Xs_re[kk][col]; b_im[nt] = Xs_im[kk][col];
            }
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) {
                    qb_dmma(acc_re[mt][nt][0], acc_re[mt][nt][1], a_re[mt], b_re[nt]);
                    qb_dmma(acc_re[mt][nt][0], acc_re[mt][nt][1], -a_im[mt], b_im[nt]);
                    qb_dmma(acc_im[mt][nt][0], acc_im[mt][nt][1], a_re[mt], b_im[nt]);
                    qb_dmma(acc_im[mt][nt][0], acc_im[mt][nt][1], a_im[mt], b_re[nt]);
                }
        }
        __syncthreads();
    }
    // epilogue: accumulator element i of tile (mt, nt): row = lane/4, col = 2*(lane%4) + i
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int i = 0; i < 2; i++) {
                const int row = m0 + wm * 16 + mt * 8 + (lane >> 2);
                const int cl = wn * 32 + nt * 8 + 2 * (lane & 3) + i;
                double2* zp = s_z[cl];
                if (zp && row < M) {
                    const double re = acc_re[mt][nt][i], im = acc_im[mt][nt][i];
                    double2 o = make_double2(alpha.x * re - alpha.y * im, alpha.x * im + alpha.y * re);
                    if (accumulate) { const double2 c = zp[row]; o.x += c.x; o.y += c.y; }
                    zp[row] = o;
                }
            }
}

// engine hook: Z_j = A x_j for all slots (per-column pointers prepared by the control kernel)
int qb_launch_dense_rhs(cudaStream_t stream, const qb_c128* A, int N, const void* const* xcols,
                        void* const* zcols, int ncols)
{
    dim3 grid((N + ZG_BM - 1) / ZG_BM, (ncols + ZG_BN - 1) / ZG_BN);
    qb_zgemm_dmma_kernel<<<grid, 256, 0, stream>>>(
        reinterpret_cast<const double2*>(A), N, N, N,
        reinterpret_cast<const double2* const*>(xcols), nullptr, 0,
        reinterpret_cast<double2* const*>(zcols), nullptr, 0, ncols, make_double2(1.0, 0.0), 0);
    QB_LAUNCH_CHECK();
    return QB_OK;
}

// out += scale * A @ X for dense column-major handles (reference zgemm path,
// core/data/matmul.pyx:329-346)
extern "C" int qb_zgemm(qb_handle ah, qb_handle xh, double sre, double sim, qb_handle oh) {
    QbDenseH* a = qb_cast<QbDenseH>(ah, QB_TAG_DENSE);
    QbDenseH* x = qb_cast<QbDenseH>(xh, QB_TAG_DENSE);
    QbDenseH* o = qb_cast<QbDenseH>(oh, QB_TAG_DENSE);
    if (!a || !x || !o) QB_FAIL(QB_E_TYPE, "zgemm needs dense handles");
    if (a->cols != x->rows || a->rows != o->rows || x->cols != o->cols)
        QB_FAIL(QB_E_SHAPE, "incompatible matrix shapes (%lld, %lld) and (%lld, %lld)",
                (long long)a->rows, (long long)a->cols, (long long)x->rows, (long long)x->cols);
    if (!(a->fortran || a->rows == 1 || a->cols == 1) || !(x->fortran || x->cols == 1) ||
        !(o->fortran || o->cols == 1))
        QB_FAIL(QB_E_TYPE, "zgemm needs column-major operands");
    if (a->rows == 0 || x->cols == 0) return QB_OK;
    dim3 grid((unsigned)((a->rows + ZG_BM - 1) / ZG_BM), (unsigned)((x->cols + ZG_BN - 1) / ZG_BN));
    qb_zgemm_dmma_kernel<<<grid, 256>>>(a->d, (int)a->rows, (int)a->cols, a->rows, nullptr, x->d,
                                        x->rows, nullptr, o->d, o->rows, (int)x->cols,
                                        make_double2(sre, sim), 1);
    QB_LAUNCH_CHECK();
    return QB_OK;
}

// `iters` back-to-back launches timed with CUDA events (bench.py dense ZGEMM figure)
extern "C" int qb_zgemm_bench(qb_handle ah, qb_handle xh, qb_handle oh, int iters, double* ms_total) {
    cudaEvent_t e0, e1;
    QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
    int rc = qb_zgemm(ah, xh, 1.0, 0.0, oh);
    if (rc) return rc;
    QB_CUDA(cudaEventRecord(e0));
    for (int i = 0; i < iters; i++) { rc = qb_zgemm(ah, xh, 1.0, 0.0, oh); if (rc) return rc; }
    QB_CUDA(cudaEventRecord(e1));
    QB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_total) *ms_total = ms;
    return QB_OK;
}

// FP64 tensor-core (DMMA m8n8k4) peak of this device, measured: every warp of a full grid
// issues independent register-only DMMA chains -- no memory traffic.  The roofline peak the
// dense path is reported against (MEASURED_PEAKS.json holds no FP64 figure).
__global__ void __launch_bounds__(256)
qb_dmma_peak_kernel(int iters, double* sink)
{
    double acc[16][2];
#pragma unroll
    for (int i = 0; i < 16; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
    double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) qb_dmma(acc[i][0], acc[i][1], a, b);
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) t += acc[i][0] + acc[i][1];
    if (t == 12345.678) sink[0] = t;                 // keeps the chains alive
}
extern "C" int qb_dmma_peak_bench(int iters, double* tflops) {
    if (iters < 1 || !tflops) QB_FAIL(QB_E_ARG, "bad arguments");
    int dev = 0, sms = 0;
    QB_CUDA(cudaGetDevice(&dev));
    QB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* sink = nullptr;
    QB_CUDA(cudaMalloc((void**)&sink, 8));
    cudaEvent_t e0, e1;
    QB_CUDA(cudaEventCreate(&e0)); QB_CUDA(cudaEventCreate(&e1));
    const int grid = sms * 8;
    qb_dmma_peak_kernel<<<grid, 256>>>(16, sink);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaEventRecord(e0));
    qb_dmma_peak_kernel<<<grid, 256>>>(iters, sink);
    QB_LAUNCH_CHECK();
    QB_CUDA(cudaEventRecord(e1));
    QB_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    // one m8n8k4 DMMA = 8*8*4 multiply-adds = 512 flop per warp instruction
    const double flop = (double)grid * 8.0 * (double)iters * 16.0 * 512.0;
    *tflops = flop / (ms * 1e-3) / 1e12;
    return QB_OK;
}
