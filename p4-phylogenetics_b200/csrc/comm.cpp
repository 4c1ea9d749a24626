// comm.cpp -- the one collective of the path: an NCCL all-reduce (sum, f64) of
// the per-part partial log-likelihoods across pattern shards.
//
// NCCL is bound at run time with dlopen so that the library loads on hosts
// without it (and shares the copy PyTorch has already loaded when the host
// program uses torch.distributed for its rendezvous).  The reference has no
// counterpart: it is a single process.
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include "engine.h"

namespace p4b {

namespace {
struct UniqueId { char internal[128]; };
typedef void *Comm;
typedef int (*GetUniqueIdFn)(UniqueId *);
typedef int (*CommInitRankFn)(Comm *, int, UniqueId, int);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, Comm, void *);
typedef int (*CommDestroyFn)(Comm);
typedef const char *(*GetErrorStringFn)(int);

struct Nccl {
    void *lib = nullptr;
    GetUniqueIdFn getUniqueId = nullptr;
    CommInitRankFn commInitRank = nullptr;
    AllReduceFn allReduce = nullptr;
    CommDestroyFn commDestroy = nullptr;
    GetErrorStringFn errorString = nullptr;
    Comm comm = nullptr;
    int world = 1;
} N;

const int kNcclDouble = 8;   // ncclFloat64
const int kNcclSum = 0;

int load()
{
    if (N.lib) return 0;
    const char *names[] = {getenv("P4B_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        N.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (N.lib) break;
    }
    if (!N.lib) { setError("cannot load NCCL (%s); set P4B_NCCL_LIB", dlerror()); return 1; }
    N.getUniqueId = (GetUniqueIdFn)dlsym(N.lib, "ncclGetUniqueId");
    N.commInitRank = (CommInitRankFn)dlsym(N.lib, "ncclCommInitRank");
    N.allReduce = (AllReduceFn)dlsym(N.lib, "ncclAllReduce");
    N.commDestroy = (CommDestroyFn)dlsym(N.lib, "ncclCommDestroy");
    N.errorString = (GetErrorStringFn)dlsym(N.lib, "ncclGetErrorString");
    if (!N.getUniqueId || !N.commInitRank || !N.allReduce || !N.commDestroy) {
        setError("the NCCL library lacks a required symbol");
        return 1;
    }
    return 0;
}

int check(int rc, const char *what)
{
    if (rc == 0) return 0;
    setError("NCCL %s failed: %s", what, N.errorString ? N.errorString(rc) : "?");
    return 1;
}
}  // namespace

int commGetUniqueId(char id128[128])
{
    if (load()) return 1;
    UniqueId id;
    if (check(N.getUniqueId(&id), "ncclGetUniqueId")) return 1;
    memcpy(id128, id.internal, 128);
    return 0;
}

int commInitRank(const char id128[128], int rank, int world)
{
    if (load()) return 1;
    if (N.comm) { setError("communicator already initialised"); return 1; }
    if (engineInitPublic()) return 1;
    if (world < 1 || rank < 0 || rank >= world) { setError("p4b_commInitRank: bad rank %d of %d", rank, world); return 1; }
    UniqueId id;
    memcpy(id.internal, id128, 128);
    if (check(N.commInitRank(&N.comm, world, id, rank), "ncclCommInitRank")) { N.comm = nullptr; return 1; }
    N.world = world;
    // the process owns its pattern range only once the communicator exists: a failed init must not
    // leave a shard behind whose partial lnL nobody sums (part mirrors and trees laid out for another
    // range are rebuilt at their next use, csrc/tree.cu partDeviceEnsure / ensureFresh)
    return setShard(rank, world);
}

int commDestroy()
{
    if (N.comm) {
        N.commDestroy(N.comm);
        N.comm = nullptr;
        N.world = 1;
        setShard(0, 1);   // without a communicator every process evaluates the whole alignment again
    }
    return 0;
}

bool commActive() { return N.comm != nullptr && N.world > 1; }

int commAllReduceSum(double *devBuf, int count, void *cudaStream)
{
    if (!N.comm) { setError("no communicator"); return 1; }
    return check(N.allReduce(devBuf, devBuf, (size_t)count, kNcclDouble, kNcclSum, N.comm, cudaStream), "ncclAllReduce");
}

}  // namespace p4b
