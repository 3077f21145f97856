// qb_host.h -- host-side handle types and error plumbing shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>
#include "qb_types.h"
#include "qb_coeff.h"
#include "qb_control.h"
#include "../../include/qutip_b200.h"

extern thread_local std::string g_qb_err;
#include <atomic>
extern std::atomic<long long> g_qb_launches;

#define QB_FAIL(code, ...) do { char _b[512]; snprintf(_b, sizeof _b, __VA_ARGS__); \
    g_qb_err = _b; return (code); } while (0)
#define QB_CUDA(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { \
    char _b[512]; snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, \
    cudaGetErrorString(_e), __FILE__, __LINE__); g_qb_err = _b; return QB_E_CUDA; } } while (0)
#define QB_LAUNCH_CHECK() do { g_qb_launches++; cudaError_t _e = cudaGetLastError(); \
    if (_e != cudaSuccess) { char _b[512]; snprintf(_b, sizeof _b, "kernel launch failed: %s (%s:%d)", \
    cudaGetErrorString(_e), __FILE__, __LINE__); g_qb_err = _b; return QB_E_CUDA; } } while (0)

enum { QB_TAG_DENSE = 0x51420001, QB_TAG_OP = 0x51420002, QB_TAG_SYS = 0x51420003,
       QB_TAG_ENG = 0x51420004, QB_TAG_COMM = 0x51420005 };

struct QbObj {
    uint32_t tag;
    int device = -1;          // device that was current when the object was created
    explicit QbObj(uint32_t t) : tag(t) { if (cudaGetDevice(&device) != cudaSuccess) { device = -1; cudaGetLastError(); } }
    virtual ~QbObj() { tag = 0; }
};

struct QbDenseH : QbObj {
    int64_t rows = 0, cols = 0;
    int fortran = 1;
    double2* d = nullptr;
    QbDenseH() : QbObj(QB_TAG_DENSE) {}
    ~QbDenseH() override { if (d) cudaFree(d); }
    int64_t size() const { return rows * cols; }
};

struct QbOpH : QbObj {
    QbOpDev dev;
    std::vector<void*> owned;
    int64_t device_bytes = 0;
    double avg_lanes = 0.0;       // DIAM fill statistic
    QbOpH() : QbObj(QB_TAG_OP) { memset(&dev, 0, sizeof dev); }
    ~QbOpH() override { for (void* p : owned) cudaFree(p); }
};

struct QbSysH : QbObj {
    int64_t N = 0;
    int nargs = 0;
    std::vector<QbOpDev> elems, cops, nops, eops;
    std::vector<std::vector<QbInstr>> elem_prog, cop_prog, nop_prog, eop_prog;
    int eop_functional = 0;
    int mc_trace = 0;            // n: mcsolve norm is tr(rho) of the column-stacked n x n state
    std::vector<QbSpline> splines;
    std::vector<double> spool;
    QbSysH() : QbObj(QB_TAG_SYS) {}
};

template <class T> static inline T* qb_cast(qb_handle h, uint32_t tag) {
    QbObj* o = static_cast<QbObj*>(h);
    if (!o || o->tag != tag) return nullptr;
    return static_cast<T*>(o);
}
