// Communicator behind the C-ABI (SURVEY.md 8(b) `smcb_comm_init`, 8(e)): one process per GPU, NCCL over
// NVLink / NVSwitch.  Every collective is enqueued on the caller's stream, so the sampler loop needs no host
// synchronisation around an exchange and no other library on the data path.
//
// The reference's only parallelism is the local ray fan-out of `sim_particle`
// (SMC_example/Micmem_likelihood.py:83-87); it has no exchange step.  The exchanges of the sharded sampler are:
//     tempering      all-gather of per-shard (max, sums)        -> smcb_temper_eval        (temper.cu)
//     MH sweep       all-gather of per-shard (counters, moments) -> smcb_sweep_begin        (sweep.cu)
//     resampling     all-gather of shard totals, then all-to-all(v) of contiguous particle chunks (here)
//
// libnccl.so.2 is opened with dlopen the first time a communicator is made (inside a Python process that has
// imported torch this resolves to the NCCL torch already loaded; SMCB_NCCL_LIB overrides the name), so a
// single-GPU user of libsmcb200.so needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>   // types and enums only: nothing is linked

#include "common.cuh"

namespace {

struct NcclApi {
    void* dso = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl(smcb_handle* h) {
    if (g_nccl.dso != nullptr) return SMCB_OK;
    const char* names[] = {getenv("SMCB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void* dso = nullptr;
    for (const char* nm : names) {
        if (nm == nullptr || nm[0] == 0) continue;
        dso = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (dso != nullptr) break;
    }
    if (dso == nullptr)
        return smcb_fail(h, SMCB_ERR_UNSUPPORTED, "smcb_comm: cannot open libnccl.so.2 (%s); set SMCB_NCCL_LIB", dlerror());
#define SYM(field, name)                                                                                    \
    do {                                                                                                    \
        *reinterpret_cast<void**>(&g_nccl.field) = dlsym(dso, name);                                        \
        if (g_nccl.field == nullptr)                                                                        \
            return smcb_fail(h, SMCB_ERR_UNSUPPORTED, "smcb_comm: libnccl has no symbol %s", name);         \
    } while (0)
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(AllReduce, "ncclAllReduce");
    SYM(Broadcast, "ncclBroadcast");
    SYM(Send, "ncclSend");
    SYM(Recv, "ncclRecv");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.dso = dso;
    return SMCB_OK;
}

}  // namespace

#define NCCL_TRY(h, expr)                                                                                   \
    do {                                                                                                    \
        ncclResult_t _r = (expr);                                                                           \
        if (_r != ncclSuccess)                                                                              \
            return smcb_fail((h), SMCB_ERR_COMM, "%s failed: %s (%s:%d)", #expr, g_nccl.GetErrorString(_r), \
                             __FILE__, __LINE__);                                                           \
    } while (0)

int comm_staging(smcb_handle* h) {
    const int rows = h->world > 1 ? h->world : 1;
    if (h->comm_send == nullptr)
        CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(&h->comm_send), sizeof(double) * SMCB_COMM_ROW));
    if (h->comm_recv == nullptr || h->comm_recv_rows < rows) {
        if (h->comm_recv) cudaFree(h->comm_recv);
        h->comm_recv = nullptr;
        CUDA_TRY(h, cudaMalloc(reinterpret_cast<void**>(&h->comm_recv), sizeof(double) * SMCB_COMM_ROW * (size_t)rows));
        h->comm_recv_rows = rows;
    }
    return SMCB_OK;
}

extern "C" int smcb_comm_unique_id(void* out_host, int nbytes) {
    if (out_host == nullptr || nbytes < (int)sizeof(ncclUniqueId))
        return smcb_fail(nullptr, SMCB_ERR_INVALID, "smcb_comm_unique_id: need a buffer of >= %d bytes",
                         (int)sizeof(ncclUniqueId));
    int rc = load_nccl(nullptr);
    if (rc) return rc;
    ncclUniqueId id;
    NCCL_TRY(nullptr, g_nccl.GetUniqueId(&id));
    memset(out_host, 0, (size_t)nbytes);
    memcpy(out_host, &id, sizeof(id));
    return SMCB_OK;
}

extern "C" int smcb_comm_init(smcb_handle* h, const void* id_host, int nbytes, int rank, int world) {
    REQUIRE(h, h != nullptr && id_host != nullptr, SMCB_ERR_INVALID, "null pointer");
    REQUIRE(h, nbytes >= (int)sizeof(ncclUniqueId), SMCB_ERR_INVALID, "unique id too short");
    REQUIRE(h, world >= 1 && rank >= 0 && rank < world, SMCB_ERR_INVALID, "need 0 <= rank < world");
    REQUIRE(h, h->comm == nullptr, SMCB_ERR_STATE, "this handle already has a communicator");
    int rc = load_nccl(h);
    if (rc) return rc;
    CUDA_TRY(h, cudaSetDevice(h->device));
    ncclUniqueId id;
    memcpy(&id, id_host, sizeof(id));
    ncclComm_t c = nullptr;
    NCCL_TRY(h, g_nccl.CommInitRank(&c, world, id, rank));
    h->comm = c;
    h->rank = rank;
    h->world = world;
    // staging for the small all-gathers of the stage loop: one row per rank, rows of at most SMCB_COMM_ROW doubles
    return comm_staging(h);
}

extern "C" int smcb_comm_destroy(smcb_handle* h) {
    REQUIRE(h, h != nullptr, SMCB_ERR_INVALID, "null handle");
    if (h->comm != nullptr) {
        cudaSetDevice(h->device);
        g_nccl.CommDestroy(static_cast<ncclComm_t>(h->comm));
        h->comm = nullptr;
    }
    h->rank = 0;
    h->world = 1;
    return SMCB_OK;
}

extern "C" int smcb_comm_rank(const smcb_handle* h) { return h ? h->rank : 0; }
extern "C" int smcb_comm_world(const smcb_handle* h) { return h ? h->world : 1; }

// recv_dev[r*count + i] = send_dev[i] of rank r
int comm_all_gather_f64(smcb_handle* h, const double* send_dev, double* recv_dev, int64_t count, cudaStream_t st) {
    if (h->world == 1) {
        if (recv_dev != send_dev)
            CUDA_TRY(h, cudaMemcpyAsync(recv_dev, send_dev, sizeof(double) * (size_t)count, cudaMemcpyDeviceToDevice, st));
        return SMCB_OK;
    }
    NCCL_TRY(h, g_nccl.AllGather(send_dev, recv_dev, (size_t)count, ncclFloat64, static_cast<ncclComm_t>(h->comm), st));
    h->collectives++;
    return SMCB_OK;
}

extern "C" int smcb_comm_all_gather(smcb_handle* h, const void* send_dev, void* recv_dev, int64_t bytes_per_rank,
                                    void* stream) {
    REQUIRE(h, h && send_dev && recv_dev && bytes_per_rank > 0, SMCB_ERR_INVALID, "bad argument");
    cudaStream_t st = as_stream(stream);
    if (h->world == 1) {
        if (recv_dev != send_dev)
            CUDA_TRY(h, cudaMemcpyAsync(recv_dev, send_dev, (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, st));
        return SMCB_OK;
    }
    NCCL_TRY(h, g_nccl.AllGather(send_dev, recv_dev, (size_t)bytes_per_rank, ncclUint8, static_cast<ncclComm_t>(h->comm), st));
    h->collectives++;
    return SMCB_OK;
}

extern "C" int smcb_comm_all_reduce_f64(smcb_handle* h, double* buf_dev, int64_t count, int op, void* stream) {
    REQUIRE(h, h && buf_dev && count > 0, SMCB_ERR_INVALID, "bad argument");
    REQUIRE(h, op == SMCB_OP_SUM || op == SMCB_OP_MAX, SMCB_ERR_INVALID, "op must be SMCB_OP_SUM or SMCB_OP_MAX");
    if (h->world == 1) return SMCB_OK;
    NCCL_TRY(h, g_nccl.AllReduce(buf_dev, buf_dev, (size_t)count, ncclFloat64, op == SMCB_OP_SUM ? ncclSum : ncclMax,
                                 static_cast<ncclComm_t>(h->comm), as_stream(stream)));
    h->collectives++;
    return SMCB_OK;
}

extern "C" int smcb_comm_broadcast(smcb_handle* h, void* buf_dev, int64_t bytes, int root, void* stream) {
    REQUIRE(h, h && buf_dev && bytes > 0 && root >= 0 && root < h->world, SMCB_ERR_INVALID, "bad argument");
    if (h->world == 1) return SMCB_OK;
    NCCL_TRY(h, g_nccl.Broadcast(buf_dev, buf_dev, (size_t)bytes, ncclUint8, root, static_cast<ncclComm_t>(h->comm),
                                 as_stream(stream)));
    h->collectives++;
    return SMCB_OK;
}

// Particle migration of the sharded resampling: rank r sends send_counts[q] doubles to rank q, taken from
// consecutive ranges of send_dev in rank order, and receives recv_counts[q] doubles from rank q into consecutive
// ranges of recv_dev (ancestors are non-decreasing, so every destination gets one contiguous range per source).
extern "C" int smcb_comm_all_to_all_v(smcb_handle* h, const double* send_dev, const int64_t* send_counts_host,
                                      double* recv_dev, const int64_t* recv_counts_host, void* stream) {
    REQUIRE(h, h && send_counts_host && recv_counts_host, SMCB_ERR_INVALID, "null pointer");
    cudaStream_t st = as_stream(stream);
    const int W = h->world, me = h->rank;
    int64_t so = 0, ro = 0;
    if (W == 1) {
        REQUIRE(h, send_counts_host[0] == recv_counts_host[0], SMCB_ERR_INVALID, "counts differ");
        if (send_counts_host[0] > 0)
            CUDA_TRY(h, cudaMemcpyAsync(recv_dev, send_dev, sizeof(double) * (size_t)send_counts_host[0],
                                        cudaMemcpyDeviceToDevice, st));
        return SMCB_OK;
    }
    ncclComm_t c = static_cast<ncclComm_t>(h->comm);
    NCCL_TRY(h, g_nccl.GroupStart());
    for (int q = 0; q < W; ++q) {
        const int64_t sc = send_counts_host[q], rc = recv_counts_host[q];
        if (q == me) {
            // the part that stays on this GPU does not go through NCCL
            if (sc != rc) {
                g_nccl.GroupEnd();
                return smcb_fail(h, SMCB_ERR_INVALID, "smcb_comm_all_to_all_v: self counts differ");
            }
            if (sc > 0)
                CUDA_TRY(h, cudaMemcpyAsync(recv_dev + ro, send_dev + so, sizeof(double) * (size_t)sc,
                                            cudaMemcpyDeviceToDevice, st));
        } else {
            if (sc > 0) NCCL_TRY(h, g_nccl.Send(send_dev + so, (size_t)sc, ncclFloat64, q, c, st));
            if (rc > 0) NCCL_TRY(h, g_nccl.Recv(recv_dev + ro, (size_t)rc, ncclFloat64, q, c, st));
        }
        so += sc;
        ro += rc;
    }
    NCCL_TRY(h, g_nccl.GroupEnd());
    h->collectives++;
    return SMCB_OK;
}

extern "C" int smcb_comm_exchange_rows(smcb_handle* h, const double* send_dev, int64_t ld_send,
                                       const int64_t* send_counts_host, double* recv_dev, int64_t ld_recv,
                                       const int64_t* recv_counts_host, int rows, void* stream) {
    REQUIRE(h, h && send_counts_host && recv_counts_host && rows >= 1, SMCB_ERR_INVALID, "bad argument");
    cudaStream_t st = as_stream(stream);
    const int W = h->world, me = h->rank;
    int64_t so = 0, ro = 0, stot = 0, rtot = 0;
    for (int q = 0; q < W; ++q) {
        REQUIRE(h, send_counts_host[q] >= 0 && recv_counts_host[q] >= 0, SMCB_ERR_INVALID, "negative count");
        stot += send_counts_host[q];
        rtot += recv_counts_host[q];
    }
    REQUIRE(h, (stot == 0 || (send_dev && ld_send >= stot)) && (rtot == 0 || (recv_dev && ld_recv >= rtot)),
            SMCB_ERR_INVALID, "buffer too narrow for the counts");
    REQUIRE(h, send_counts_host[me] == recv_counts_host[me], SMCB_ERR_INVALID, "self counts differ");
    ncclComm_t c = static_cast<ncclComm_t>(h->comm);
    bool group = false;
    for (int q = 0; q < W; ++q) {
        const int64_t sc = send_counts_host[q], rc = recv_counts_host[q];
        if (q == me) {
            if (sc > 0)
                CUDA_TRY(h, cudaMemcpy2DAsync(recv_dev + ro, sizeof(double) * (size_t)ld_recv, send_dev + so,
                                              sizeof(double) * (size_t)ld_send, sizeof(double) * (size_t)sc, (size_t)rows,
                                              cudaMemcpyDeviceToDevice, st));
        } else if (sc > 0 || rc > 0) {
            if (!group) {
                NCCL_TRY(h, g_nccl.GroupStart());
                group = true;
            }
            for (int k = 0; k < rows; ++k) {
                if (sc > 0) NCCL_TRY(h, g_nccl.Send(send_dev + (size_t)k * ld_send + so, (size_t)sc, ncclFloat64, q, c, st));
                if (rc > 0) NCCL_TRY(h, g_nccl.Recv(recv_dev + (size_t)k * ld_recv + ro, (size_t)rc, ncclFloat64, q, c, st));
            }
        }
        so += sc;
        ro += rc;
    }
    if (group) {
        NCCL_TRY(h, g_nccl.GroupEnd());
        h->collectives++;
    }
    return SMCB_OK;
}

extern "C" int64_t smcb_collective_count(const smcb_handle* h) { return h ? h->collectives : 0; }
