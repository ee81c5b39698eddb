// comm.cu — the one collective of the hot path: an all-reduce of the 30-double
// ICP system per iteration.  The reference has no collective
// layer at all, so there is nothing upstream to mirror.
//
// NCCL is resolved with dlopen("libnccl.so.2") at first use: inside a torch
// process that is torch's bundled NCCL (already mapped), otherwise the system
// library.  Declarations below restate the stable NCCL 2.x C ABI (nccl.h).
#include <dlfcn.h>

#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

#include "comm.h"
#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;   // NCCL_UNIQUE_ID_BYTES
typedef void* nccl_comm_t;
enum { kNcclSum = 0, kNcclFloat64 = 8 };

struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

NcclApi g_nccl;
std::once_flag g_once;

void load_nccl() {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(h, "ncclAllGather");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    g_nccl.ok = g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.AllReduce && g_nccl.CommDestroy;
}

int nccl_ready() {
    std::call_once(g_once, load_nccl);
    if (!g_nccl.ok) {
        o3db::set_last_error("NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "missing symbols");
        return O3DB_ERR_COMM;
    }
    return O3DB_OK;
}

int nccl_check(int rc, const char* what) {
    if (rc == 0) return O3DB_OK;
    o3db::set_last_error("%s failed: %s", what, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "NCCL error");
    return O3DB_ERR_COMM;
}

}  // namespace

struct o3db_comm {
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
    // in-kernel exchange over peer memory (see icp.cu: peer_all_reduce); absent => NCCL all-reduce
    bool peer = false;
    double* my_box = nullptr;                 // cudaMalloc'ed, exported through CUDA IPC
    void* opened[o3db::kMaxPeers] = {};       // peers' mailboxes as opened here
    unsigned long long* seq = nullptr;
    o3db::PeerView view{};
};

namespace {

// Maps every rank's mailbox into every other rank's address space.  All ranks take the same decision (the
// last all-gather carries each rank's success flag), so either every rank uses the in-kernel exchange or none.
void setup_peer_exchange(o3db_comm* c) {
    using namespace o3db;
    if (c->world < 2 || c->world > kMaxPeers || !g_nccl.AllGather || getenv("O3DB_COMM_NO_PEER")) return;
    const size_t box_bytes = (size_t)2 * c->world * kBoxDoubles * sizeof(double);
    struct Msg {
        cudaIpcMemHandle_t h;
        int ok;
        int pad[15];
    };
    static_assert(sizeof(Msg) == 128, "Msg layout");
    Msg mine{}, *all_dev = nullptr, *mine_dev = nullptr;
    std::vector<Msg> all((size_t)c->world);
    bool ok = cudaMalloc(&c->my_box, box_bytes) == cudaSuccess && cudaMemset(c->my_box, 0, box_bytes) == cudaSuccess &&
              cudaMalloc(&c->seq, sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(c->seq, 0, sizeof(unsigned long long)) == cudaSuccess &&
              cudaIpcGetMemHandle(&mine.h, c->my_box) == cudaSuccess;
    mine.ok = ok ? 1 : 0;
    auto gather = [&]() -> bool {
        return cudaMemcpy(mine_dev, &mine, sizeof(Msg), cudaMemcpyHostToDevice) == cudaSuccess &&
               g_nccl.AllGather(mine_dev, all_dev, sizeof(Msg), /*ncclInt8*/ 0, c->comm, (cudaStream_t)0) == 0 &&
               cudaStreamSynchronize(0) == cudaSuccess &&
               cudaMemcpy(all.data(), all_dev, sizeof(Msg) * c->world, cudaMemcpyDeviceToHost) == cudaSuccess;
    };
    if (cudaMalloc(&all_dev, sizeof(Msg) * c->world) != cudaSuccess || cudaMalloc(&mine_dev, sizeof(Msg)) != cudaSuccess) {
        (void)cudaGetLastError();
        return;   // (cannot even run the agreement round: every rank that fails here fails alike — out of memory)
    }
    bool all_ok = gather();
    for (int p = 0; all_ok && p < c->world; ++p) all_ok = all[p].ok != 0;
    if (all_ok) {
        for (int p = 0; p < c->world && ok; ++p) {
            if (p == c->rank) {
                c->view.box[p] = c->my_box;
            } else if (cudaIpcOpenMemHandle(&c->opened[p], all[p].h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                c->view.box[p] = (double*)c->opened[p];
            } else {
                ok = false;
            }
        }
    }
    // second round: did every rank manage to open every mailbox?
    mine.ok = (all_ok && ok) ? 1 : 0;
    all_ok = gather();
    for (int p = 0; all_ok && p < c->world; ++p) all_ok = all[p].ok != 0;
    cudaFree(all_dev);
    cudaFree(mine_dev);
    (void)cudaGetLastError();
    if (!all_ok) return;
    c->view.seq = c->seq;
    c->view.rank = c->rank;
    c->view.world = c->world;
    c->peer = true;
}

}  // namespace

extern "C" {

int o3db_comm_get_unique_id(uint8_t id_out[O3DB_UNIQUE_ID_BYTES]) {
    O3DB_REQUIRE(id_out != nullptr, "o3db_comm_get_unique_id: null output");
    int rc = nccl_ready();
    if (rc) return rc;
    nccl_unique_id id;
    rc = nccl_check(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
    if (rc) return rc;
    memcpy(id_out, id.internal, O3DB_UNIQUE_ID_BYTES);
    return O3DB_OK;
}

int o3db_comm_create(const uint8_t id_bytes[O3DB_UNIQUE_ID_BYTES], int rank, int world_size, o3db_comm** out) {
    O3DB_REQUIRE(out != nullptr && id_bytes != nullptr, "o3db_comm_create: null argument");
    *out = nullptr;
    O3DB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "o3db_comm_create: bad rank/world_size");
    int rc = nccl_ready();
    if (rc) return rc;
    nccl_unique_id id;
    memcpy(id.internal, id_bytes, O3DB_UNIQUE_ID_BYTES);
    o3db_comm* c = new (std::nothrow) o3db_comm();
    O3DB_REQUIRE(c != nullptr, "out of host memory");
    c->rank = rank;
    c->world = world_size;
    rc = nccl_check(g_nccl.CommInitRank(&c->comm, world_size, id, rank), "ncclCommInitRank");
    if (rc) {
        delete c;
        return rc;
    }
    setup_peer_exchange(c);
    *out = c;
    return O3DB_OK;
}

int o3db_comm_uses_peer_memory(const o3db_comm* comm) { return comm && comm->peer ? 1 : 0; }

int o3db_comm_allreduce_f64(o3db_comm* comm, double* buf_dev, int count, void* stream) {
    O3DB_REQUIRE(comm != nullptr && buf_dev != nullptr && count > 0, "o3db_comm_allreduce_f64: bad arguments");
    return nccl_check(g_nccl.AllReduce(buf_dev, buf_dev, (size_t)count, kNcclFloat64, kNcclSum, comm->comm,
                                       (cudaStream_t)stream),
                      "ncclAllReduce");
}

void o3db_comm_destroy(o3db_comm* comm) {
    if (!comm) return;
    cudaDeviceSynchronize();
    for (void* p : comm->opened)
        if (p) cudaIpcCloseMemHandle(p);
    if (comm->my_box) cudaFree(comm->my_box);
    if (comm->seq) cudaFree(comm->seq);
    if (comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(comm->comm);
    delete comm;
}

}  // extern "C"

const o3db::PeerView* o3db_comm_peer_view(const o3db_comm* comm) { return comm && comm->peer ? &comm->view : nullptr; }
