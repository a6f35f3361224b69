// mxb_exchange: symmetric peer-memory buffers for the mix-bus all-reduce (see exchange.cuh, bank.cu).
//
// Protocol (per call, sequence number s, slot = s & 1; every rank owns flags[world] + payload[2 slots][world][n], all
// peer-mapped with CUDA IPC). PUSH model -- nothing is ever read over NVLink, nobody polls remote memory:
//   1. every rank stores each locally reduced bus value into payload[slot][its rank] of EVERY rank's buffer (posted
//      peer stores, issued by the reducing warps as they finish), then -- once all of its CTAs are done -- publishes
//      flags[its rank] = s + 1 in every rank's buffer (release, system scope)
//   2. every rank waits on its OWN flags (local polling, acquire, system scope) until all of them reach s + 1 and adds
//      the world buses, all in local HBM, IN RANK ORDER -- every rank computes the same sum, bit for bit, every run
//   3. slot reuse at call s + 2 is safe: a rank finishes call s + 1 only after every peer's flag reached s + 2, i.e.
//      after every peer had launched call s + 1 and therefore finished reading slot(s) (stream order on the peer)
// One kernel does the local reduction and the exchange (bank.cu: mix_reduce_exchange_kernel); nothing here calls NCCL.
#include <new>

#include "exchange.cuh"

using namespace mxb;

namespace mxb {
ExchDev exchange_next(mxb_exchange* ex) {
    ExchDev d;
    memset(&d, 0, sizeof(d));
    const size_t slot_off = ex->flags_bytes + (size_t)(ex->seq & 1) * ex->slot_bytes;
    d.rank = ex->rank; d.world = ex->world; d.stride = ex->max_doubles; d.seq1 = ex->seq + 1; d.ticket = ex->ticket; d.status = ex->status; d.timeout_ns = ex->timeout_ns;
    for (int r = 0; r < ex->world; ++r) {
        d.dst_flag[r] = (unsigned long long*)(ex->peers[r] + (size_t)ex->rank * kExchFlagBytes);
        d.dst_payload[r] = (double*)(ex->peers[r] + slot_off) + (size_t)ex->rank * (size_t)ex->max_doubles;
    }
    d.src_flags = (const unsigned long long*)ex->local;
    d.src_payload = (const double*)(ex->local + slot_off);
    ex->seq += 1;
    return d;
}
}  // namespace mxb

extern "C" {

int32_t mxb_exchange_create(mxb_ctx* ctx, int32_t rank, int32_t world, int32_t max_doubles, mxb_exchange** out) {
    MXB_REQUIRE(ctx && out, MXB_ERR_INVALID, "mxb_exchange_create: NULL argument");
    *out = nullptr;
    MXB_REQUIRE(world >= 1 && world <= kExchMaxWorld && rank >= 0 && rank < world, MXB_ERR_INVALID, "mxb_exchange_create: rank %d of %d", rank, world);
    MXB_REQUIRE(max_doubles > 0, MXB_ERR_INVALID, "mxb_exchange_create: max_doubles %d", max_doubles);
    DeviceGuard g(ctx->device);
    mxb_exchange* ex = new (std::nothrow) mxb_exchange();
    MXB_REQUIRE(ex, MXB_ERR_ALLOC, "mxb_exchange_create: out of host memory");
    memset(ex, 0, sizeof(*ex));
    ex->ctx = ctx; ex->rank = rank; ex->world = world; ex->max_doubles = max_doubles;
    ex->flags_bytes = (size_t)kExchMaxWorld * kExchFlagBytes;
    ex->slot_bytes = (sizeof(double) * (size_t)max_doubles * (size_t)world + 127) & ~(size_t)127;
    const size_t total = ex->flags_bytes + 2 * ex->slot_bytes;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, total);                   // plain cudaMalloc: exportable with cudaIpcGetMemHandle
    if (e != cudaSuccess) { set_error("mxb_exchange_create: cudaMalloc: %s", cudaGetErrorString(e)); delete ex; return MXB_ERR_ALLOC; }
    ex->local = (unsigned char*)p;
    e = cudaMemset(p, 0, total);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ex->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ex->ticket, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&ex->status, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(ex->status, 0, sizeof(unsigned int));
    {   // a dead peer must not hang the others: the flag wait is bounded (the kernel then reports through `status`)
        const char* t = getenv("MXB_EXCHANGE_TIMEOUT_MS");
        const double ms = t ? atof(t) : 5000.0;
        ex->timeout_ns = (unsigned long long)((ms > 0.0 ? ms : 5000.0) * 1e6);
    }
    if (e != cudaSuccess) { set_error("mxb_exchange_create: %s", cudaGetErrorString(e)); cudaFree(p); delete ex; return MXB_ERR_CUDA; }
    ex->peers[rank] = ex->local;
    ex->connected = (world == 1);
    *out = ex;
    return MXB_OK;
}

int32_t mxb_exchange_local_handle(mxb_exchange* ex, void* handle, int32_t handle_bytes) {
    MXB_REQUIRE(ex && handle, MXB_ERR_INVALID, "mxb_exchange_local_handle: NULL argument");
    MXB_REQUIRE(handle_bytes >= (int32_t)sizeof(cudaIpcMemHandle_t), MXB_ERR_INVALID, "mxb_exchange_local_handle: need %zu bytes", sizeof(cudaIpcMemHandle_t));
    DeviceGuard g(ex->ctx->device);
    cudaIpcMemHandle_t h;
    MXB_CUDA(cudaIpcGetMemHandle(&h, ex->local));
    memcpy(handle, &h, sizeof(h));
    return MXB_OK;
}

int32_t mxb_exchange_connect(mxb_exchange* ex, const void* all_handles) {
    MXB_REQUIRE(ex && all_handles, MXB_ERR_INVALID, "mxb_exchange_connect: NULL argument");
    DeviceGuard g(ex->ctx->device);
    const unsigned char* hs = (const unsigned char*)all_handles;
    for (int r = 0; r < ex->world; ++r) {
        if (r == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)r * sizeof(h), sizeof(h));
        void* p = nullptr;
        MXB_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ex->peers[r] = (unsigned char*)p;
    }
    ex->connected = true;
    return MXB_OK;
}

int32_t mxb_exchange_status(mxb_exchange* ex, int32_t* timed_out_ranks) {
    MXB_REQUIRE(ex && timed_out_ranks, MXB_ERR_INVALID, "mxb_exchange_status: NULL argument");
    DeviceGuard g(ex->ctx->device);
    unsigned int m = 0;
    MXB_CUDA(cudaMemcpy(&m, ex->status, sizeof(m), cudaMemcpyDeviceToHost));      // synchronises with the kernels that may set it
    *timed_out_ranks = (int32_t)m;
    return MXB_OK;
}

int32_t mxb_exchange_destroy(mxb_exchange* ex) {
    if (!ex) return MXB_OK;
    DeviceGuard g(ex->ctx->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < ex->world; ++r)
        if (r != ex->rank && ex->peers[r]) cudaIpcCloseMemHandle(ex->peers[r]);
    cudaFree(ex->local); cudaFree(ex->ticket); cudaFree(ex->status);
    delete ex;
    return MXB_OK;
}

}  // extern "C"
