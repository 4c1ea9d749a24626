// comm.cpp -- the one collective of the path: an NCCL all-reduce (sum, f64) of
// the per-part partial log-likelihoods across pattern shards.
//
// NCCL is bound at run time with dlopen so that the library loads on hosts
// without it (and shares the copy PyTorch has already loaded when the host
// program uses torch.distributed for its rendezvous).  The reference has no
// counterpart: it is a single process.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"

namespace p4b {

namespace {
struct UniqueId { char internal[128]; };
typedef void *Comm;
typedef int (*GetUniqueIdFn)(UniqueId *);
typedef int (*CommInitRankFn)(Comm *, int, UniqueId, int);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, Comm, void *);
typedef int (*AllGatherFn)(const void *, void *, size_t, int, Comm, void *);
typedef int (*CommDestroyFn)(Comm);
typedef const char *(*GetErrorStringFn)(int);

struct Nccl {
    void *lib = nullptr;
    GetUniqueIdFn getUniqueId = nullptr;
    CommInitRankFn commInitRank = nullptr;
    AllReduceFn allReduce = nullptr;
    AllGatherFn allGather = nullptr;
    CommDestroyFn commDestroy = nullptr;
    GetErrorStringFn errorString = nullptr;
    Comm comm = nullptr;
    int world = 1;
} N;

const int kNcclDouble = 8;   // ncclFloat64
const int kNcclChar = 0;     // ncclInt8
const int kNcclSum = 0;
const int kNcclMin = 3;

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
    N.allGather = (AllGatherFn)dlsym(N.lib, "ncclAllGather");
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
int commWorld() { return N.comm ? N.world : 1; }

// Peer mailboxes for the in-kernel all-reduce of the partial log-likelihoods (csrc/tree.cu): every rank's mailbox is
// opened on every other rank through CUDA IPC, the 64-byte handles travelling by one ncclAllGather.  peers[r] receives
// the address of rank r's mailbox as this process sees it (its own for r == rank).  Returns 0 when EVERY rank succeeded
// (agreed by an all-reduce of the outcome), 1 otherwise -- then nobody uses the mailboxes and NCCL stays in charge.
int commOpenPeerMailboxes(void *mine, int rank, void **peers, void *cudaStream)
{
    if (!N.comm || !N.allGather) return 1;
    const int world = N.world;
    cudaStream_t st = (cudaStream_t)cudaStream;
    int ok = 1;
    cudaIpcMemHandle_t h;
    memset(&h, 0, sizeof(h));
    if (cudaIpcGetMemHandle(&h, mine) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    char *dBuf = nullptr;
    if (cudaMalloc(&dBuf, sizeof(h) * (size_t)(world + 1) + sizeof(double)) != cudaSuccess) { cudaGetLastError(); return 1; }
    std::vector<cudaIpcMemHandle_t> all(world);
    // every rank takes part in both collectives whatever its own outcome so far, so that nobody is left waiting
    cudaMemcpyAsync(dBuf, &h, sizeof(h), cudaMemcpyHostToDevice, st);
    if (N.allGather(dBuf, dBuf + sizeof(h), sizeof(h), kNcclChar, N.comm, st) != 0) ok = 0;
    cudaMemcpyAsync(all.data(), dBuf + sizeof(h), sizeof(h) * world, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); ok = 0; }
    if (ok)
        for (int r = 0; r < world; r++) {
            if (r == rank) { peers[r] = mine; continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = 0; break; }
            peers[r] = p;
        }
    double flag = ok ? 1.0 : 0.0, *dFlag = reinterpret_cast<double *>(dBuf + sizeof(h) * (size_t)(world + 1));
    cudaMemcpyAsync(dFlag, &flag, sizeof(double), cudaMemcpyHostToDevice, st);
    if (N.allReduce(dFlag, dFlag, 1, kNcclDouble, kNcclMin, N.comm, st) != 0) flag = 0.0;
    else cudaMemcpyAsync(&flag, dFlag, sizeof(double), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { cudaGetLastError(); flag = 0.0; }
    cudaFree(dBuf);
    return flag > 0.5 ? 0 : 1;
}
void commClosePeerMailboxes(int rank, void **peers)
{
    for (int r = 0; r < N.world; r++)
        if (r != rank && peers[r]) { cudaIpcCloseMemHandle(peers[r]); peers[r] = nullptr; }
}

int commAllReduceSum(double *devBuf, int count, void *cudaStream)
{
    if (!N.comm) { setError("no communicator"); return 1; }
    return check(N.allReduce(devBuf, devBuf, (size_t)count, kNcclDouble, kNcclSum, N.comm, cudaStream), "ncclAllReduce");
}

}  // namespace p4b
