// Statistics all-reduce behind the C-ABI (mimo_comm_init / mimo_comm_allreduce_stats / mimo_comm_destroy).
//
// The only exchange step of a data-sharded sweep is the sum over shards of the packed FP64 statistics (K x F doubles
// + the lower-bound scalar) -- the reference's list-of-arrays semantics, distributions/gaussian.py:503-505 and
// utils/abstraction.py:12-14.  A caller that binds this library without torch gets it here: NCCL is
// loaded at run time with dlopen("libnccl.so.2") (inside a torch process that is the copy torch already mapped), so
// the library itself keeps no link-time dependency beyond the CUDA runtime.  The 128-byte unique id of rank 0 travels
// out of band (any channel the caller has: a file, MPI, torch.distributed's store).
#include <dlfcn.h>
#include <string.h>
#include "common.cuh"
#include "internal.h"

namespace mimo {

namespace {
constexpr size_t NCCL_ID_BYTES = 128;
struct NcclId { char internal[NCCL_ID_BYTES]; };
typedef void* NcclComm;
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef const char* (*fn_errstr)(int);
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;           // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since 2.0)

struct NcclApi {
    void* handle = nullptr;
    fn_get_id get_id = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_destroy destroy = nullptr;
    fn_errstr errstr = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return MIMO_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { set_error("mimo_comm: cannot load libnccl.so.2 (%s)", dlerror()); return MIMO_EUNSUPPORTED; }
    g_nccl.get_id = (fn_get_id)dlsym(h, "ncclGetUniqueId");
    g_nccl.init_rank = (fn_init_rank)dlsym(h, "ncclCommInitRank");
    g_nccl.allreduce = (fn_allreduce)dlsym(h, "ncclAllReduce");
    g_nccl.destroy = (fn_destroy)dlsym(h, "ncclCommDestroy");
    g_nccl.errstr = (fn_errstr)dlsym(h, "ncclGetErrorString");
    if (!g_nccl.get_id || !g_nccl.init_rank || !g_nccl.allreduce || !g_nccl.destroy) {
        set_error("mimo_comm: libnccl lacks an expected symbol");
        dlclose(h);
        return MIMO_EUNSUPPORTED;
    }
    g_nccl.handle = h;
    return MIMO_OK;
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return MIMO_OK;
    set_error("mimo_comm: %s failed: %s", what, g_nccl.errstr ? g_nccl.errstr(rc) : "NCCL error");
    return MIMO_ECUDA;
}
}  // namespace

size_t comm_unique_id_bytes() { return NCCL_ID_BYTES; }

int comm_unique_id(void* out) {
    MIMO_CHECK_ARG(out, "null pointer");
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    rc = nccl_check(g_nccl.get_id(&id), "ncclGetUniqueId");
    if (rc) return rc;
    memcpy(out, id.internal, NCCL_ID_BYTES);
    return MIMO_OK;
}

int comm_init(int world, int rank, const void* unique_id, void** comm_out) {
    MIMO_CHECK_ARG(unique_id && comm_out && world >= 1 && rank >= 0 && rank < world, "arguments");
    int rc = load_nccl();
    if (rc) return rc;
    NcclId id;
    memcpy(id.internal, unique_id, NCCL_ID_BYTES);
    NcclComm c = nullptr;
    rc = nccl_check(g_nccl.init_rank(&c, world, id, rank), "ncclCommInitRank");
    if (rc) return rc;
    *comm_out = c;
    return MIMO_OK;
}

int comm_allreduce_stats(void* comm, double* stat, int64_t count, cudaStream_t st) {
    MIMO_CHECK_ARG(comm && stat && count >= 0, "arguments");
    if (count == 0) return MIMO_OK;
    return nccl_check(g_nccl.allreduce(stat, stat, (size_t)count, NCCL_FLOAT64, NCCL_SUM, (NcclComm)comm, st), "ncclAllReduce");
}

int comm_destroy(void* comm) {
    if (!comm) return MIMO_OK;
    MIMO_CHECK_ARG(g_nccl.handle, "no communicator was created");
    return nccl_check(g_nccl.destroy((NcclComm)comm), "ncclCommDestroy");
}

}  // namespace mimo
