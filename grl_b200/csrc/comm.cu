// grl_b200 — communicator of the gallery-sharded search: NCCL over NVLink / NVSwitch, one rank per GPU.
// The reference has no counterpart (single process, nn.DataParallel, mars_train.py:80; evaluation on one device,
// attevaluator.py:125-163); BASELINE.json configs[4] shards the gallery over 1/2/4/8 GPUs.
#include <dlfcn.h>

#include <mutex>

#include "comm.h"

namespace grl {

static NcclApi g_api;
static int g_api_state = 0;          // 0 not tried, 1 loaded, -1 failed
static char g_api_err[256];
static std::mutex g_api_mu;

const NcclApi* nccl_api(grl_handle* h) {
    std::lock_guard<std::mutex> lk(g_api_mu);
    if (g_api_state == 0) {
        // a PyTorch process has libnccl.so.2 loaded already: RTLD_NOLOAD hands back that very instance (one NCCL per process)
        void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            snprintf(g_api_err, sizeof(g_api_err), "dlopen(libnccl.so.2) failed: %s", dlerror());
            g_api_state = -1;
        } else {
            bool ok = true;
            auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p) { ok = false; snprintf(g_api_err, sizeof(g_api_err), "NCCL symbol %s missing", name); } return p; };
            g_api.GetUniqueId = (decltype(g_api.GetUniqueId))sym("ncclGetUniqueId");
            g_api.CommInitRank = (decltype(g_api.CommInitRank))sym("ncclCommInitRank");
            g_api.CommDestroy = (decltype(g_api.CommDestroy))sym("ncclCommDestroy");
            g_api.CommCount = (decltype(g_api.CommCount))sym("ncclCommCount");
            g_api.CommUserRank = (decltype(g_api.CommUserRank))sym("ncclCommUserRank");
            g_api.AllGather = (decltype(g_api.AllGather))sym("ncclAllGather");
            g_api.AllReduce = (decltype(g_api.AllReduce))sym("ncclAllReduce");
            g_api.ReduceScatter = (decltype(g_api.ReduceScatter))sym("ncclReduceScatter");
            g_api.Send = (decltype(g_api.Send))sym("ncclSend");
            g_api.Recv = (decltype(g_api.Recv))sym("ncclRecv");
            g_api.GroupStart = (decltype(g_api.GroupStart))sym("ncclGroupStart");
            g_api.GroupEnd = (decltype(g_api.GroupEnd))sym("ncclGroupEnd");
            g_api.GetErrorString = (decltype(g_api.GetErrorString))sym("ncclGetErrorString");
            g_api_state = ok ? 1 : -1;
        }
    }
    if (g_api_state != 1) {
        set_error(h, GRL_ENCCL, "NCCL unavailable: %s", g_api_err);
        return nullptr;
    }
    return &g_api;
}

void comm_release(grl_handle* h) {
    if (h && h->comm && h->comm_owned) {
        const NcclApi* api = nccl_api(h);
        if (api) api->CommDestroy((ncclComm_t)h->comm);
    }
    if (h) { h->comm = nullptr; h->comm_world = 0; h->comm_rank = 0; h->comm_owned = 0; }
}

}  // namespace grl

using namespace grl;

extern "C" int grl_comm_unique_id(grl_handle* h, void* id_host, size_t id_bytes) {
    if (!h || !id_host || id_bytes < sizeof(ncclUniqueId)) return set_error(h, GRL_EINVAL, "grl_comm_unique_id: need a %zu-byte host buffer", sizeof(ncclUniqueId));
    const NcclApi* api = nccl_api(h);
    if (!api) return GRL_ENCCL;
    ncclUniqueId id;
    GRL_NCCL(h, api, api->GetUniqueId(&id));
    memcpy(id_host, &id, sizeof(id));
    return GRL_OK;
}

extern "C" int grl_comm_init(grl_handle* h, const void* id_host, size_t id_bytes, int world, int rank) {
    if (!h || !id_host || id_bytes < sizeof(ncclUniqueId)) return set_error(h, GRL_EINVAL, "grl_comm_init: need the %zu-byte unique id", sizeof(ncclUniqueId));
    if (world < 1 || rank < 0 || rank >= world) return set_error(h, GRL_EINVAL, "grl_comm_init: bad world %d / rank %d", world, rank);
    const NcclApi* api = nccl_api(h);
    if (!api) return GRL_ENCCL;
    comm_release(h);
    GRL_CUDA(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    ncclComm_t c = nullptr;
    GRL_NCCL(h, api, api->CommInitRank(&c, world, id, rank));
    h->comm = c; h->comm_world = world; h->comm_rank = rank; h->comm_owned = 1;
    return GRL_OK;
}

extern "C" int grl_comm_attach(grl_handle* h, void* nccl_comm) {
    if (!h) return GRL_EINVAL;
    comm_release(h);
    if (!nccl_comm) return GRL_OK;                   // detach: back to a single rank
    const NcclApi* api = nccl_api(h);
    if (!api) return GRL_ENCCL;
    int world = 0, rank = 0;
    GRL_NCCL(h, api, api->CommCount((ncclComm_t)nccl_comm, &world));
    GRL_NCCL(h, api, api->CommUserRank((ncclComm_t)nccl_comm, &rank));
    h->comm = nccl_comm; h->comm_world = world; h->comm_rank = rank; h->comm_owned = 0;
    return GRL_OK;
}

extern "C" int grl_comm_destroy(grl_handle* h) {
    if (!h) return GRL_EINVAL;
    comm_release(h);
    return GRL_OK;
}

extern "C" int grl_comm_info(const grl_handle* h, int* world, int* rank) {
    if (!h) return GRL_EINVAL;
    if (world) *world = h->comm ? h->comm_world : 1;
    if (rank) *rank = h->comm ? h->comm_rank : 0;
    return GRL_OK;
}
