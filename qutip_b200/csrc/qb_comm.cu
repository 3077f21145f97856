// qb_comm.cu -- the ONE collective of the sharded workloads: an NCCL all-reduce
// (sum) of the expectation sums [2][n_e_ops][n_times] over NVLink, behind the C ABI.
//
// mcsolve trajectories and sweep members are independent (solver/multitraj.py:250-256), so
// the only data-path exchange is what _TrajectorySum.reduce_expect
// (solver/multitrajresult.py:1116-1124) accumulates.  Two ways to form the group:
//   * one process, several devices (qb_comm_init_all -> ncclCommInitAll): the b200 map of a
//     single python process drives every GPU of the box from one thread per device;
//   * one process per device (qb_comm_unique_id on rank 0, the 128-byte id is handed to the
//     other ranks by the launcher, qb_comm_init_rank -> ncclCommInitRank): torchrun jobs.
// NCCL is bound at run time (dlopen of libnccl.so.2, the image's system library or the copy
// a host framework already mapped); nothing else of a host framework is involved.
#include <dlfcn.h>
#include <mutex>
#include <nccl.h>
#include "qb_host.h"

namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                              cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string err;
};
NcclApi g_nccl;
std::mutex g_nccl_mu;

const NcclApi* nccl_api() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.lib) return &g_nccl;
    const char* names[] = {getenv("QUTIP_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) { g_nccl.err = dlerror() ? dlerror() : "libnccl.so.2 not found"; return nullptr; }
#define QB_SYM(field, name) \
    *(void**)(&g_nccl.field) = dlsym(lib, name); \
    if (!g_nccl.field) { g_nccl.err = std::string("missing NCCL symbol ") + name; dlclose(lib); return nullptr; }
    QB_SYM(GetUniqueId, "ncclGetUniqueId")
    QB_SYM(CommInitAll, "ncclCommInitAll")
    QB_SYM(CommInitRank, "ncclCommInitRank")
    QB_SYM(CommDestroy, "ncclCommDestroy")
    QB_SYM(AllReduce, "ncclAllReduce")
    QB_SYM(GroupStart, "ncclGroupStart")
    QB_SYM(GroupEnd, "ncclGroupEnd")
    QB_SYM(GetErrorString, "ncclGetErrorString")
    QB_SYM(GetVersion, "ncclGetVersion")
#undef QB_SYM
    g_nccl.lib = lib;
    return &g_nccl;
}

#define QB_NCCL(api, call) do { ncclResult_t _r = (call); if (_r != ncclSuccess) { \
    char _b[512]; snprintf(_b, sizeof _b, "%s failed: %s (%s:%d)", #call, (api)->GetErrorString(_r), \
    __FILE__, __LINE__); g_qb_err = _b; return QB_E_CUDA; } } while (0)

struct QbCommH : QbObj {
    int nranks = 0;                       // size of the whole group
    std::vector<int> devs;                // local members (one per device driven by this process)
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;
    std::vector<double*> dbuf;            // per local member: reduction buffer on its device
    size_t cap = 0;                       // doubles
    QbCommH() : QbObj(QB_TAG_COMM) {}
    ~QbCommH() override {
        const NcclApi* api = g_nccl.lib ? &g_nccl : nullptr;
        for (size_t i = 0; i < devs.size(); i++) {
            cudaSetDevice(devs[i]);
            if (i < dbuf.size() && dbuf[i]) cudaFree(dbuf[i]);
            if (i < comms.size() && comms[i] && api) api->CommDestroy(comms[i]);
            if (i < streams.size() && streams[i]) cudaStreamDestroy(streams[i]);
        }
    }
    int reserve(size_t n) {
        if (n <= cap) return QB_OK;
        for (size_t i = 0; i < devs.size(); i++) {
            QB_CUDA(cudaSetDevice(devs[i]));
            if (dbuf[i]) { cudaFree(dbuf[i]); dbuf[i] = nullptr; }
            QB_CUDA(cudaMalloc((void**)&dbuf[i], n * sizeof(double)));
        }
        cap = n;
        return QB_OK;
    }
};

int comm_finish(QbCommH* c) {
    c->streams.assign(c->devs.size(), nullptr);
    c->dbuf.assign(c->devs.size(), nullptr);
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaStreamCreateWithFlags(&c->streams[i], cudaStreamNonBlocking));
    }
    return QB_OK;
}

// in-place sum over the group of dbuf[i][0..count) of every local member: ONE collective
int comm_allreduce(QbCommH* c, size_t count) {
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    QB_NCCL(api, api->GroupStart());
    for (size_t i = 0; i < c->devs.size(); i++)
        QB_NCCL(api, api->AllReduce(c->dbuf[i], c->dbuf[i], count, ncclDouble, ncclSum, c->comms[i],
                                    c->streams[i]));
    QB_NCCL(api, api->GroupEnd());
    g_qb_launches += (long long)c->devs.size();
    return QB_OK;
}
}  // namespace

extern "C" int qb_comm_nccl_version(int* version) {
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    if (!version) QB_FAIL(QB_E_ARG, "null output");
    QB_NCCL(api, api->GetVersion(version));
    return QB_OK;
}

extern "C" int qb_comm_init_all(int ndev, const int* devs, qb_handle* out) {
    if (ndev < 1 || !devs || !out) QB_FAIL(QB_E_ARG, "bad communicator arguments");
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    QbCommH* c = new QbCommH();
    c->nranks = ndev;
    c->devs.assign(devs, devs + ndev);
    c->comms.assign(ndev, nullptr);
    ncclResult_t r = api->CommInitAll(c->comms.data(), ndev, devs);
    if (r != ncclSuccess) { c->comms.assign(ndev, nullptr); delete c; QB_FAIL(QB_E_CUDA, "ncclCommInitAll failed: %s", api->GetErrorString(r)); }
    int rc = comm_finish(c);
    if (rc) { delete c; return rc; }
    *out = c;
    return QB_OK;
}

extern "C" int qb_comm_unique_id(void* id, int nbytes) {
    if (!id || nbytes < (int)sizeof(ncclUniqueId)) QB_FAIL(QB_E_ARG, "id buffer must hold %d bytes", (int)sizeof(ncclUniqueId));
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    ncclUniqueId u;
    QB_NCCL(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return QB_OK;
}

extern "C" int qb_comm_init_rank(int nranks, int rank, const void* id, qb_handle* out) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id || !out) QB_FAIL(QB_E_ARG, "bad communicator arguments");
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    int dev = 0;
    QB_CUDA(cudaGetDevice(&dev));
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    QbCommH* c = new QbCommH();
    c->nranks = nranks;
    c->devs.assign(1, dev);
    c->comms.assign(1, nullptr);
    ncclResult_t r = api->CommInitRank(&c->comms[0], nranks, u, rank);
    if (r != ncclSuccess) { c->comms[0] = nullptr; delete c; QB_FAIL(QB_E_CUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r)); }
    int rc = comm_finish(c);
    if (rc) { delete c; return rc; }
    *out = c;
    return QB_OK;
}

extern "C" int qb_comm_info(qb_handle comm, int* nranks, int* nlocal) {
    QbCommH* c = qb_cast<QbCommH>(comm, QB_TAG_COMM);
    if (!c) QB_FAIL(QB_E_TYPE, "not a communicator handle");
    if (nranks) *nranks = c->nranks;
    if (nlocal) *nlocal = (int)c->devs.size();
    return QB_OK;
}

// bufs[i]: `count` doubles in host memory of local member i; summed over the whole group in
// place (every member ends with the same values)
extern "C" int qb_comm_allreduce_sum(qb_handle comm, double* const* bufs, int64_t count) {
    QbCommH* c = qb_cast<QbCommH>(comm, QB_TAG_COMM);
    if (!c) QB_FAIL(QB_E_TYPE, "not a communicator handle");
    if (!bufs || count < 1) QB_FAIL(QB_E_ARG, "bad all-reduce arguments");
    int rc = c->reserve((size_t)count);
    if (rc) return rc;
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaMemcpyAsync(c->dbuf[i], bufs[i], (size_t)count * 8, cudaMemcpyHostToDevice, c->streams[i]));
    }
    rc = comm_allreduce(c, (size_t)count);
    if (rc) return rc;
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaMemcpyAsync(bufs[i], c->dbuf[i], (size_t)count * 8, cudaMemcpyDeviceToHost, c->streams[i]));
    }
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaStreamSynchronize(c->streams[i]));
    }
    return QB_OK;
}

// same, on device buffers: dbufs[i] = `count` doubles on the device of local member i; the
// collective is enqueued after everything already queued on the legacy default stream of each
// device and the call returns when the sums are in place
extern "C" int qb_comm_allreduce_sum_device(qb_handle comm, void* const* dbufs, int64_t count) {
    QbCommH* c = qb_cast<QbCommH>(comm, QB_TAG_COMM);
    if (!c) QB_FAIL(QB_E_TYPE, "not a communicator handle");
    if (!dbufs || count < 1) QB_FAIL(QB_E_ARG, "bad all-reduce arguments");
    const NcclApi* api = nccl_api();
    if (!api) QB_FAIL(QB_E_STATE, "NCCL unavailable: %s", g_nccl.err.c_str());
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaDeviceSynchronize());          // producers of dbufs[i] have finished
    }
    QB_NCCL(api, api->GroupStart());
    for (size_t i = 0; i < c->devs.size(); i++)
        QB_NCCL(api, api->AllReduce(dbufs[i], dbufs[i], (size_t)count, ncclDouble, ncclSum, c->comms[i], c->streams[i]));
    QB_NCCL(api, api->GroupEnd());
    g_qb_launches += (long long)c->devs.size();
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaStreamSynchronize(c->streams[i]));
    }
    return QB_OK;
}

// internal accessors of qb_engine.cu
int qb_engine_expect_view(qb_handle eng, const void** d_expect, int* neops, int* device, int64_t* ntraj, int* nt);
int qb_reduce_expect_on(cudaStream_t stream, const void* d_expect, int64_t ntraj, int neops, int nt, void* d_sums);

// Expectation sums of a sharded run: engines[i] (on the device of local member i) has just
// finished qb_engine_run; its per-trajectory expectation values are still on its device.
// Each member reduces them to [2][n_e][n_t] complex sums (sum e, sum (re^2, im^2)) with one
// kernel, the members are summed with ONE ncclAllReduce, and member 0's copy is returned in
// sums (host, 4 * n_e * n_t doubles).  Members that ran no trajectory pass a NULL engine.
extern "C" int qb_comm_reduce_expect(qb_handle comm, const qb_handle* engines, int neops, int nt,
                                     void* sums) {
    QbCommH* c = qb_cast<QbCommH>(comm, QB_TAG_COMM);
    if (!c) QB_FAIL(QB_E_TYPE, "not a communicator handle");
    if (!engines || neops < 1 || nt < 1 || !sums) QB_FAIL(QB_E_ARG, "bad reduce arguments");
    const size_t count = (size_t)4 * neops * nt;
    int rc = c->reserve(count);
    if (rc) return rc;
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        if (!engines[i]) { QB_CUDA(cudaMemsetAsync(c->dbuf[i], 0, count * 8, c->streams[i])); continue; }
        const void* d_exp = nullptr; int ne = 0, dev = -1, ent = 0; int64_t ntraj = 0;
        rc = qb_engine_expect_view(engines[i], &d_exp, &ne, &dev, &ntraj, &ent);
        if (rc) return rc;
        if (dev != c->devs[i]) QB_FAIL(QB_E_ARG, "engine %d lives on device %d, communicator member on %d", (int)i, dev, c->devs[i]);
        if (ne != neops || ent != nt) QB_FAIL(QB_E_SHAPE, "engine %d ran %d e_ops x %d times, expected %d x %d", (int)i, ne, ent, neops, nt);
        rc = qb_reduce_expect_on(c->streams[i], d_exp, ntraj, neops, nt, c->dbuf[i]);
        if (rc) return rc;
    }
    rc = comm_allreduce(c, count);
    if (rc) return rc;
    QB_CUDA(cudaSetDevice(c->devs[0]));
    QB_CUDA(cudaMemcpyAsync(sums, c->dbuf[0], count * 8, cudaMemcpyDeviceToHost, c->streams[0]));
    for (size_t i = 0; i < c->devs.size(); i++) {
        QB_CUDA(cudaSetDevice(c->devs[i]));
        QB_CUDA(cudaStreamSynchronize(c->streams[i]));
    }
    return QB_OK;
}
