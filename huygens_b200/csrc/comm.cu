// comm.cu — filter-partitioned multi-GPU rendering: hg_comm_* and hg_fb_process_sharded.
//
// The reference sums a bank in filter-index order (src/filterbank.h:130).  When the bank is split over the GPUs of
// one box — rank r holds the filters [r N/G, (r+1) N/G) — the same order is kept by extending the chain of
// chain.cuh across devices: the last CTA of rank r stores its running row straight into a ring in rank r+1's memory
// (peer mapping over NVLink; cudaIpc handles between processes, cudaDeviceEnablePeerAccess inside one), rank r+1's
// gateway warp polls that ring locally and publishes its position to a word in rank r's memory.  One kernel per GPU,
// the transfer rides tile by tile under the compute; the last rank holds the result, bit-identical to the
// single-bank sum.  HG_SHARD_REDUCE is the order-free alternative: every rank renders its own partial sum and NCCL
// reduces the rows (tolerance mode, reported next to the chain in bench.py).
//
// NCCL is loaded at run time (dlopen of libnccl.so.2: the copy torch already mapped, else the system one) — it
// bootstraps the peer mapping (allgather of the IPC handles), moves the result to `root` and runs the reduce.
#include <dlfcn.h>
#include <nccl.h>
#include <vector>
#include "hg_common.cuh"
#include "filterbank_handle.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);     // already in the process (torch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
#define HG_SYM(field, name) api.field = reinterpret_cast<decltype(api.field)>(dlsym(h, name)); if (!api.field) return nullptr
    HG_SYM(GetUniqueId, "ncclGetUniqueId");
    HG_SYM(CommInitRank, "ncclCommInitRank");
    HG_SYM(CommInitAll, "ncclCommInitAll");
    HG_SYM(CommDestroy, "ncclCommDestroy");
    HG_SYM(AllGather, "ncclAllGather");
    HG_SYM(Reduce, "ncclReduce");
    HG_SYM(Send, "ncclSend");
    HG_SYM(Recv, "ncclRecv");
    HG_SYM(GroupStart, "ncclGroupStart");
    HG_SYM(GroupEnd, "ncclGroupEnd");
    HG_SYM(GetErrorString, "ncclGetErrorString");
#undef HG_SYM
    api.lib = h;
    return &api;
}

#define HG_NCCL(expr)                                                                                   \
    do {                                                                                                \
        ncclResult_t r__ = (expr);                                                                      \
        if (r__ != ncclSuccess)                                                                         \
            return hg::fail(HG_ERR_NCCL, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, nccl()->GetErrorString(r__)); \
    } while (0)

constexpr int BOX_LOG2 = 11;                                   // inbox ring: 2048 samples
constexpr size_t BOX_RING_WORDS = ((size_t)1 << BOX_LOG2) * 2; // sized for double (two LL words per sample)
constexpr size_t BOX_BYTES = BOX_RING_WORDS * 8 + 64;          // ring + the credit word the next rank publishes to

}  // namespace

// One rank of a filter-partitioned group.  `box` (this device's memory) = the inbox ring rank-1 writes + the
// credit word rank+1 writes; `next_ring` / `prev_credit` are the peer mappings of the neighbours' boxes.
struct hg_comm {
    int rank = 0, size = 1, device = 0;
    ncclComm_t nc = nullptr;
    bool owns_nc = true;
    unsigned long long* box = nullptr;
    unsigned long long* next_ring = nullptr;      // rank+1's inbox ring (null on the last rank)
    unsigned long long* prev_credit = nullptr;    // rank-1's credit word (null on rank 0)
    void* ipc_next = nullptr;                     // base pointers returned by cudaIpcOpenMemHandle (to close)
    void* ipc_prev = nullptr;
    unsigned long long vbase = 4096;              // virtual sample position shared by the group (same on every rank)
    hg::DevBuf scratch;
    unsigned long long* my_ring() const { return box; }
    unsigned long long* my_credit() const { return box + BOX_RING_WORDS; }
};

namespace {

hg_status comm_alloc_box(hg_comm* c) {
    HG_CUDA(cudaSetDevice(c->device));
    HG_CUDA(cudaMalloc((void**)&c->box, BOX_BYTES));
    HG_CUDA(cudaMemset(c->box, 0, BOX_BYTES));
    return HG_OK;
}

}  // namespace

extern "C" {

hg_status hg_comm_get_unique_id(void* id128) {
    if (!id128) return hg::fail(HG_ERR_INVALID, "hg_comm_get_unique_id: null output");
    if (!nccl()) return hg::fail(HG_ERR_NCCL, "hg_comm_get_unique_id: libnccl.so.2 could not be loaded");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    HG_NCCL(nccl()->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return HG_OK;
}

hg_status hg_comm_init_rank(hg_comm** out, int n_ranks, int rank, const void* id128, int device) {
    if (!out || !id128) return hg::fail(HG_ERR_INVALID, "hg_comm_init_rank: null argument");
    *out = nullptr;
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return hg::fail(HG_ERR_INVALID, "hg_comm_init_rank: rank %d of %d", rank, n_ranks);
    if (!nccl()) return hg::fail(HG_ERR_NCCL, "hg_comm_init_rank: libnccl.so.2 could not be loaded");
    if (device < 0 || device >= hg_device_count()) return hg::fail(HG_ERR_CUDA, "hg_comm_init_rank: CUDA device %d not available", device);
    HG_CUDA(cudaSetDevice(device));
    hg_comm* c = new (std::nothrow) hg_comm();
    if (!c) return hg::fail(HG_ERR_NOMEM, "hg_comm_init_rank: host allocation failed");
    c->rank = rank; c->size = n_ranks; c->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    hg_status st = HG_OK;
    do {
        ncclResult_t r = nccl()->CommInitRank(&c->nc, n_ranks, id, rank);
        if (r != ncclSuccess) { st = hg::fail(HG_ERR_NCCL, "ncclCommInitRank -> %s", nccl()->GetErrorString(r)); break; }
        if ((st = comm_alloc_box(c)) != HG_OK) break;
        if (n_ranks == 1) break;
        // exchange the boxes' IPC handles and map the two neighbours
        cudaIpcMemHandle_t mine;
        cudaError_t ce = cudaIpcGetMemHandle(&mine, c->box);
        if (ce != cudaSuccess) { st = hg::fail(HG_ERR_CUDA, "cudaIpcGetMemHandle -> %s", cudaGetErrorString(ce)); break; }
        unsigned char* d_h = nullptr;
        const size_t hs = sizeof(cudaIpcMemHandle_t);
        if ((ce = cudaMalloc((void**)&d_h, hs * (n_ranks + 1))) != cudaSuccess) { st = hg::fail(HG_ERR_NOMEM, "cudaMalloc -> %s", cudaGetErrorString(ce)); break; }
        cudaMemcpy(d_h, &mine, hs, cudaMemcpyHostToDevice);
        r = nccl()->AllGather(d_h, d_h + hs, hs, ncclChar, c->nc, 0);
        cudaError_t se = cudaStreamSynchronize(0);
        std::vector<cudaIpcMemHandle_t> all(n_ranks);
        cudaMemcpy(all.data(), d_h + hs, hs * n_ranks, cudaMemcpyDeviceToHost);
        cudaFree(d_h);
        if (r != ncclSuccess || se != cudaSuccess) { st = hg::fail(HG_ERR_NCCL, "allgather of the IPC handles failed: %s", r != ncclSuccess ? nccl()->GetErrorString(r) : cudaGetErrorString(se)); break; }
        if (rank + 1 < n_ranks) {
            ce = cudaIpcOpenMemHandle(&c->ipc_next, all[rank + 1], cudaIpcMemLazyEnablePeerAccess);
            if (ce != cudaSuccess) { st = hg::fail(HG_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) -> %s", rank + 1, cudaGetErrorString(ce)); break; }
            c->next_ring = (unsigned long long*)c->ipc_next;
        }
        if (rank > 0) {
            ce = cudaIpcOpenMemHandle(&c->ipc_prev, all[rank - 1], cudaIpcMemLazyEnablePeerAccess);
            if (ce != cudaSuccess) { st = hg::fail(HG_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) -> %s", rank - 1, cudaGetErrorString(ce)); break; }
            c->prev_credit = (unsigned long long*)c->ipc_prev + BOX_RING_WORDS;
        }
    } while (0);
    if (st != HG_OK) { hg_comm_destroy(c); return st; }
    *out = c;
    return HG_OK;
}

hg_status hg_comm_init_all(hg_comm** out, int n_devices, const int* devices) {
    if (!out || n_devices < 1) return hg::fail(HG_ERR_INVALID, "hg_comm_init_all: bad arguments");
    for (int i = 0; i < n_devices; i++) out[i] = nullptr;
    if (!nccl()) return hg::fail(HG_ERR_NCCL, "hg_comm_init_all: libnccl.so.2 could not be loaded");
    std::vector<int> devs(n_devices);
    for (int i = 0; i < n_devices; i++) {
        devs[i] = devices ? devices[i] : i;
        if (devs[i] < 0 || devs[i] >= hg_device_count()) return hg::fail(HG_ERR_CUDA, "hg_comm_init_all: CUDA device %d not available", devs[i]);
    }
    std::vector<ncclComm_t> ncs(n_devices);
    HG_NCCL(nccl()->CommInitAll(ncs.data(), n_devices, devs.data()));
    hg_status st = HG_OK;
    for (int i = 0; i < n_devices && st == HG_OK; i++) {
        hg_comm* c = new (std::nothrow) hg_comm();
        if (!c) { st = hg::fail(HG_ERR_NOMEM, "hg_comm_init_all: host allocation failed"); break; }
        c->rank = i; c->size = n_devices; c->device = devs[i]; c->nc = ncs[i];
        out[i] = c;
        st = comm_alloc_box(c);
    }
    // one process: the neighbours' boxes are plain device pointers once peer access is on
    for (int i = 0; i + 1 < n_devices && st == HG_OK; i++) {
        for (int dir = 0; dir < 2 && st == HG_OK; dir++) {
            const int a = devs[i + dir], b = devs[i + 1 - dir];
            int can = 0;
            cudaDeviceCanAccessPeer(&can, a, b);
            if (!can) { st = hg::fail(HG_ERR_UNSUPPORTED, "hg_comm_init_all: device %d cannot map device %d (no peer access)", a, b); break; }
            cudaSetDevice(a);
            cudaError_t ce = cudaDeviceEnablePeerAccess(b, 0);
            if (ce != cudaSuccess && ce != cudaErrorPeerAccessAlreadyEnabled) st = hg::fail(HG_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) -> %s", a, b, cudaGetErrorString(ce));
            cudaGetLastError();
        }
        if (st == HG_OK) {
            out[i]->next_ring = out[i + 1]->my_ring();
            out[i + 1]->prev_credit = out[i]->my_credit();
        }
    }
    if (st != HG_OK) {
        for (int i = 0; i < n_devices; i++) { if (out[i]) hg_comm_destroy(out[i]); else if (ncs[i]) nccl()->CommDestroy(ncs[i]); out[i] = nullptr; }
        return st;
    }
    return HG_OK;
}

hg_status hg_comm_destroy(hg_comm* c) {
    if (!c) return HG_OK;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->ipc_next) cudaIpcCloseMemHandle(c->ipc_next);
    if (c->ipc_prev) cudaIpcCloseMemHandle(c->ipc_prev);
    if (c->box) cudaFree(c->box);
    c->scratch.release();
    if (c->nc && nccl()) nccl()->CommDestroy(c->nc);
    delete c;
    return HG_OK;
}

hg_status hg_comm_group_start(void) {
    if (!nccl()) return hg::fail(HG_ERR_NCCL, "hg_comm_group_start: libnccl.so.2 could not be loaded");
    HG_NCCL(nccl()->GroupStart());
    return HG_OK;
}
hg_status hg_comm_group_end(void) {
    if (!nccl()) return hg::fail(HG_ERR_NCCL, "hg_comm_group_end: libnccl.so.2 could not be loaded");
    HG_NCCL(nccl()->GroupEnd());
    return HG_OK;
}

int hg_comm_rank(const hg_comm* c) { return c ? c->rank : -1; }
int hg_comm_size(const hg_comm* c) { return c ? c->size : 0; }

hg_status hg_fb_process_sharded(hg_fb* fb, hg_comm* c, const void* in, void* out_sum, size_t L, int shard_mode, int root,
                                void* cuda_stream) {
    if (!fb || !c) return hg::fail(HG_ERR_INVALID, "hg_fb_process_sharded: null handle");
    if (shard_mode != HG_SHARD_CHAIN && shard_mode != HG_SHARD_REDUCE) return hg::fail(HG_ERR_INVALID, "hg_fb_process_sharded: bad mode %d", shard_mode);
    if (fb->S != 1) return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process_sharded: filter-partitioned rendering is for one stream per handle (batches of streams are stream-partitioned: no exchange)");
    if (fb->device != c->device) return hg::fail(HG_ERR_INVALID, "hg_fb_process_sharded: handle on device %d, communicator on device %d", fb->device, c->device);
    const int last = c->size - 1;
    if (root < 0) root = shard_mode == HG_SHARD_CHAIN ? last : 0;
    if (root >= c->size) return hg::fail(HG_ERR_INVALID, "hg_fb_process_sharded: root %d of %d ranks", root, c->size);
    if (c->rank == root && !out_sum) return hg::fail(HG_ERR_INVALID, "hg_fb_process_sharded: out_sum is required on the root rank");
    if (L == 0) return HG_OK;
    HG_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t es = hg::dtype_size(fb->dtype);
    const ncclDataType_t nt = fb->dtype == HG_F32 ? ncclFloat : ncclDouble;

    if (shard_mode == HG_SHARD_REDUCE || c->size == 1) {
        // every rank renders its own (ordered) partial sum; NCCL adds the rows on `root`
        void* d_part = out_sum;
        const bool direct = c->size == 1 || (c->rank == root && hg::is_device_ptr(out_sum));
        if (!direct) { HG_TRY(c->scratch.reserve(L * es)); d_part = c->scratch.p; }
        HG_TRY(hg_fb_process(fb, in, d_part, nullptr, L, HG_MODE_EXACT, HG_DIST_NONE, cuda_stream));
        if (c->size == 1) return HG_OK;
        void* d_dst = nullptr;
        if (c->rank == root) d_dst = hg::is_device_ptr(out_sum) ? out_sum : c->scratch.p;
        HG_NCCL(nccl()->Reduce(d_part, d_dst, L, nt, ncclSum, root, c->nc, st));
        if (c->rank == root && !hg::is_device_ptr(out_sum)) {
            HG_CUDA(cudaMemcpyAsync(out_sum, d_dst, L * es, cudaMemcpyDeviceToHost, st));
            HG_CUDA(cudaStreamSynchronize(st));
        }
        return HG_OK;
    }

    // ---- chain mode: the running row crosses the devices in filter order ----
    if (fb->chain_owner != (const void*)c || fb->chain_vbase != c->vbase) {
        // this handle's private rings may hold tags of another position sequence: start them clean
        HG_CUDA(cudaStreamSynchronize(st));
        if (fb->d_chain_ring) HG_CUDA(cudaMemsetAsync(fb->d_chain_ring, 0, fb->chain_ring_words * 8, st));
        if (fb->d_chain_credit) HG_CUDA(cudaMemsetAsync(fb->d_chain_credit, 0, fb->chain_links * (1 + hg::chain::CW) * 8, st));
        fb->chain_owner = c;
    }
    fb->chain_vbase = c->vbase;
    fb->cur_peer = hg::chain::Params();
    if (c->rank > 0) { fb->cur_peer.inbox = c->my_ring(); fb->cur_peer.inbox_credit = c->prev_credit; }
    if (c->rank < last) { fb->cur_peer.outbox = c->next_ring; fb->cur_peer.outbox_credit = c->my_credit(); }
    // the chain's tail holds the result; every other rank renders without a local output row
    void* d_out = nullptr;
    if (c->rank == last) {
        d_out = out_sum;
        if (root != last || !out_sum) { HG_TRY(c->scratch.reserve(L * es)); d_out = c->scratch.p; }
    }
    hg_status rc = hg_fb_process(fb, in, d_out, nullptr, L, HG_MODE_EXACT, HG_DIST_NONE, cuda_stream);
    fb->cur_peer = hg::chain::Params();
    HG_TRY(rc);
    c->vbase = fb->chain_vbase;
    if (root != last) {
        if (c->rank == last) HG_NCCL(nccl()->Send(d_out, L, nt, root, c->nc, st));
        if (c->rank == root) {
            void* d_dst = out_sum;
            if (!hg::is_device_ptr(out_sum)) { HG_TRY(c->scratch.reserve(L * es)); d_dst = c->scratch.p; }
            HG_NCCL(nccl()->Recv(d_dst, L, nt, last, c->nc, st));
            if (d_dst != out_sum) {
                HG_CUDA(cudaMemcpyAsync(out_sum, d_dst, L * es, cudaMemcpyDeviceToHost, st));
                HG_CUDA(cudaStreamSynchronize(st));
            }
        }
    }
    return HG_OK;
}

}  // extern "C"
