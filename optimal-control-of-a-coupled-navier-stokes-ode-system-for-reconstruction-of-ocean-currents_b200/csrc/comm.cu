// The one collective of the path: all-reduce(sum, fp64) of the accumulator [b | misfit | n_masked]
// across the ranks that share the buoys, through NCCL over NVLink / NVSwitch.
//
// NCCL is resolved at run time (dlopen / dlsym), not at link time: libocp_b200.so must load on hosts without NCCL
// (the CPU test container, single-GPU users), and inside a Python process it must share the libnccl.so.2 that torch
// has already mapped instead of pulling in a second copy.  Only <nccl.h>'s TYPES are used at compile time.
#include "comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdlib>
#include <mutex>

namespace ocp {

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    std::string why;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        const char *env = getenv("OCP_NCCL_LIB");
        const char *names[] = {env, "libnccl.so.2", "libnccl.so"};
        // a copy already mapped into the process (torch's) wins over a fresh load
        for (const char *nm : names)
            if (nm && !a.handle) a.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        for (const char *nm : names)
            if (nm && !a.handle) a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (!a.handle) {
            a.why = "libnccl.so.2 not found (set OCP_NCCL_LIB)";
            return;
        }
#define OCP_SYM(field, name)                                             \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, name)); \
    if (!a.field) a.why = std::string("NCCL symbol missing: ") + name;
        OCP_SYM(GetUniqueId, "ncclGetUniqueId")
        OCP_SYM(CommInitRank, "ncclCommInitRank")
        OCP_SYM(AllReduce, "ncclAllReduce")
        OCP_SYM(CommDestroy, "ncclCommDestroy")
        OCP_SYM(GetErrorString, "ncclGetErrorString")
        OCP_SYM(GetVersion, "ncclGetVersion")
#undef OCP_SYM
    });
    return a;
}

bool ok(ncclResult_t r, const char *what, std::string &err) {
    if (r == ncclSuccess) return true;
    err = std::string(what) + ": " + (api().GetErrorString ? api().GetErrorString(r) : "NCCL error");
    return false;
}

}  // namespace

bool comm_available(std::string &err) {
    NcclApi &a = api();
    if (!a.handle || !a.why.empty()) {
        err = a.why.empty() ? "NCCL unavailable" : a.why;
        return false;
    }
    return true;
}

int comm_nccl_version() {
    std::string e;
    int v = 0;
    if (!comm_available(e) || api().GetVersion(&v) != ncclSuccess) return 0;
    return v;
}

bool comm_unique_id(void *id128, std::string &err) {
    if (!comm_available(err)) return false;
    static_assert(sizeof(ncclUniqueId) == 128, "OCP_COMM_ID_BYTES must equal NCCL_UNIQUE_ID_BYTES");
    return ok(api().GetUniqueId(reinterpret_cast<ncclUniqueId *>(id128)), "ncclGetUniqueId", err);
}

Communicator::~Communicator() { destroy(); }

void Communicator::destroy() {
    if (comm_ && api().CommDestroy) api().CommDestroy(reinterpret_cast<ncclComm_t>(comm_));
    comm_ = nullptr;
    nranks_ = 1;
    rank_ = 0;
}

bool Communicator::init(int nranks, int rank, const void *id128, std::string &err) {
    if (nranks < 1 || rank < 0 || rank >= nranks || !id128) {
        err = "ocp_comm_init: bad rank / size";
        return false;
    }
    destroy();
    if (nranks == 1) return true;      // nothing to exchange: stays a no-op
    if (!comm_available(err)) return false;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclComm_t c = nullptr;
    if (!ok(api().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank", err)) return false;
    comm_ = c;
    nranks_ = nranks;
    rank_ = rank;
    return true;
}

bool Communicator::allreduce_sum(double *d_buf, size_t n, cudaStream_t s, std::string &err) {
    if (nranks_ <= 1 || n == 0) return true;
    return ok(api().AllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, reinterpret_cast<ncclComm_t>(comm_), s),
              "ncclAllReduce", err);
}

bool Communicator::allreduce_sum_i64(long long *d_buf, size_t n, cudaStream_t s, std::string &err) {
    if (nranks_ <= 1 || n == 0) return true;
    return ok(api().AllReduce(d_buf, d_buf, n, ncclInt64, ncclSum, reinterpret_cast<ncclComm_t>(comm_), s),
              "ncclAllReduce(int64)", err);
}

}  // namespace ocp
