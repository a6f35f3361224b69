// Voice bank host side: handles, parameter upload, coefficient design (the reference's formulas, run once
// per parameter change on the host with the same libm the reference uses), launches.
#include <math.h>
#include <algorithm>

#include <new>
#include <thread>

#include "bank_kernels.cuh"
#include "delay_kernels.cuh"
#include "exchange.cuh"
#include "filter_design.cuh"

using namespace mxb;

struct mxb_bank {
    mxb_ctx* ctx;
    mxb_bank_desc desc;
    int V;
    bool set_mask[MXB_P_COUNT];
    std::vector<double> hp[MXB_P_COUNT];     // host copies of the parameter arrays (coefficient design input)
    double* dp[MXB_P_COUNT];                 // device copies: the buffer the next block reads (== dp_buf[id][dp_cur[id]])
    // Block-rate control data uploaded with mxb_bank_set_param_async from host memory is double-buffered: the copy of block
    // k+1's values runs on a private copy stream into the buffer block k is NOT reading, so it overlaps block k's kernel.
    double* dp_buf[MXB_P_COUNT][2];          // [1] allocated on first asynchronous host upload
    int dp_cur[MXB_P_COUNT];
    bool dp_pending[MXB_P_COUNT];            // an upload into dp_buf[id][1 - cur] is in flight: the next block flips to it
    cudaEvent_t ev_up[MXB_P_COUNT];          // upload finished (copy stream)
    cudaEvent_t ev_read[MXB_P_COUNT][2];     // last kernel that read dp_buf[id][k] finished (process stream)
    bool ev_read_set[MXB_P_COUNT][2];
    cudaStream_t up_stream;
    double* osc_out;
    double *f0, *f1, *f2, *cf[5];
    double *env_amp, *env_output;
    long long *env_holdcount, *env_hold;
    int* env_flags;
    int *trig_on, *trig_off;                 // staging for MXB_MEM_HOST gates
    int *dl_phase, *dl_size, *dl_pos;
    double* ring;                            // [V][delay_taps]
    double* partials; size_t partials_len;
    double* mix_dev;                         // [max_frames][2]
    void* out_stage; size_t out_stage_bytes; // staging for MXB_MEM_HOST out
    double* fm_stage; size_t fm_stage_bytes; // staging for host-resident per-sample frequencies
    double* cm_stage; size_t cm_stage_bytes; // ... and cutoffs
    double* ds_stage; size_t ds_stage_bytes; // ... and delay sizes
    unsigned char* tv_stage; size_t tv_stage_bytes;   // ... and per-sample triggers
    int64_t launches;
    mxb_exchange* ex;                        // peer-memory mix exchange (multi-GPU), or NULL
};

namespace {

// coefficient design, once per parameter change, on the host with the libm the reference calls (filter_design.cuh holds the
// reference's formulas, shared with the per-sample device design of patch stages)
void design_lores(const mxb_bank* b, std::vector<double>* cf, const int v_lo, const int v_hi) {
    const double sr = (double)(size_t)b->ctx->sample_rate;
    for (int v = v_lo; v < v_hi; ++v) design_lores_one(b->hp[MXB_P_CUTOFF][v], b->hp[MXB_P_RESONANCE][v], sr, cf[0][v], cf[1][v]);
}
void design_svf(const mxb_bank* b, std::vector<double>* cf, const int v_lo, const int v_hi) {
    const double sr = (double)(size_t)b->ctx->sample_rate;
    for (int v = v_lo; v < v_hi; ++v) {
        double c[5];
        design_svf_one(b->hp[MXB_P_CUTOFF][v], b->hp[MXB_P_RESONANCE][v], sr, c);
        for (int i = 0; i < 5; ++i) cf[i][v] = c[i];
    }
}
void design_biquad(const mxb_bank* b, std::vector<double>* cf, const int v_lo, const int v_hi) {
    const double sr = (double)(size_t)b->ctx->sample_rate;
    for (int v = v_lo; v < v_hi; ++v) {
        double c[5];
        design_biquad_one(b->desc.biquad_type, b->hp[MXB_P_CUTOFF][v], b->hp[MXB_P_RESONANCE][v], b->hp[MXB_P_GAIN][v], sr, c);
        for (int i = 0; i < 5; ++i) cf[i][v] = c[i];
    }
}

int redesign(mxb_bank* b) {
    const int fk = b->desc.filt_kind;
    if (fk == MXB_FILT_NONE) return MXB_OK;
    if ((fk == MXB_FILT_LORES || fk == MXB_FILT_HIRES) && !(b->set_mask[MXB_P_CUTOFF] && b->set_mask[MXB_P_RESONANCE])) return MXB_OK;
    if (fk == MXB_FILT_BIQUAD && !(b->set_mask[MXB_P_CUTOFF] && b->set_mask[MXB_P_RESONANCE])) return MXB_OK;   // untouched maxiBiquad: all-zero coefficients
    std::vector<double> cf[5];
    for (auto& c : cf) c.assign((size_t)b->V, 0.0);
    // the design runs on the host with the libm the reference calls (bit-identical coefficients); a large bank is split
    // over host threads (voices are independent), so that one cutoff change on 1 Mi voices costs a block, not dozens
    auto part = [&](int lo, int hi) {
        if (fk == MXB_FILT_LORES || fk == MXB_FILT_HIRES) design_lores(b, cf, lo, hi);
        else if (fk == MXB_FILT_SVF) design_svf(b, cf, lo, hi);
        else design_biquad(b, cf, lo, hi);
    };
    const unsigned hw = std::thread::hardware_concurrency();
    const int nt = b->V < 32768 ? 1 : (int)std::min<unsigned>(hw ? hw : 1u, 32u);
    if (nt <= 1) part(0, b->V);
    else {
        std::vector<std::thread> ts;
        for (int i = 0; i < nt; ++i) ts.emplace_back(part, (int)((long long)b->V * i / nt), (int)((long long)b->V * (i + 1) / nt));
        for (auto& t : ts) t.join();
    }
    for (int i = 0; i < 5; ++i) MXB_CUDA(cudaMemcpy(b->cf[i], cf[i].data(), sizeof(double) * (size_t)b->V, cudaMemcpyHostToDevice));
    return MXB_OK;
}

constexpr int kMixReduceThreads = 128;

// Sum of one (frame, channel) row of per-warp partials by one CTA: every thread adds a strided subset in ascending
// order (two interleaved chains), a fixed xor tree per warp, then the four warp sums in a fixed order -- the
// summation order never changes between runs. Returned to every thread of warp 0; other warps get 0.
__device__ __forceinline__ double mix_row_sum(const double* __restrict__ p, const int W, double* sm) {
    double s0 = 0.0, s1 = 0.0;
    int w = threadIdx.x;
    for (; w + kMixReduceThreads < W; w += 2 * kMixReduceThreads) { s0 += p[w]; s1 += p[w + kMixReduceThreads]; }
    if (w < W) s0 += p[w];
    double s = s0 + s1;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    return threadIdx.x < 32 ? (sm[0] + sm[1]) + (sm[2] + sm[3]) : 0.0;
}

__global__ void __launch_bounds__(kMixReduceThreads) mix_reduce_kernel(const double* __restrict__ partials, double* __restrict__ mix, int rows, int W) {
    __shared__ double sm[kMixReduceThreads / 32];
    const int row = blockIdx.x;                       // one CTA per row: rows CTAs keep enough loads in flight for HBM
    const double s = mix_row_sum(partials + (size_t)row * (size_t)W, W, sm);
    if (threadIdx.x == 0) mix[row] = s;
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// K3 + K6 in one kernel: every row sum of the local reduction (as mix_reduce_kernel) is PUSHED into this rank's lane
// of every rank's exchange buffer over NVLink (posted stores); the last CTA to finish publishes the flags, waits on
// its own (local) flags for every peer and adds the world buses in rank order (protocol: exchange.cu).
__global__ void __launch_bounds__(kMixReduceThreads) mix_reduce_exchange_kernel(const double* __restrict__ partials, double* __restrict__ mix, int rows, int W, const ExchDev x) {
    __shared__ double sm[kMixReduceThreads / 32];
    const int row = blockIdx.x;
    const double s = mix_row_sum(partials + (size_t)row * (size_t)W, W, sm);
    if (threadIdx.x < x.world) x.dst_payload[threadIdx.x][row] = s;      // thread r -> rank r's buffer
    __shared__ bool last;
    __threadfence_system();                           // this CTA's peer stores are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(x.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < x.world) {
        __threadfence_system();                       // cumulative: the other CTAs' stores were ordered before their tickets
        st_release_sys(x.dst_flag[threadIdx.x], x.seq1);
        // bounded wait: a peer that died (or never launched this block) costs timeout_ns, not a hung box; the bus is then
        // incomplete and bit `rank` of *status says whose part is missing (mxb_exchange_status / mxb_bank_process report it)
        const unsigned long long t0 = globaltimer_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(x.src_flags + (size_t)threadIdx.x * (kExchFlagBytes / sizeof(unsigned long long))) < x.seq1) {
            if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > x.timeout_ns) { atomicOr(x.status, 1u << threadIdx.x); break; }
        }
    }
    if (threadIdx.x == 0) *x.ticket = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < x.world; ++r)             // rank order on every rank: identical bits everywhere, every run
            s += __ldcg(x.src_payload + (size_t)r * (size_t)x.stride + i);
        mix[i] = s;
    }
}

// ring slots [0, n) of one voice <-> a linear array (set != 0: linear -> ring)
__global__ void ring_linear_kernel(double* __restrict__ ring, size_t V, size_t voice, double* __restrict__ lin, int n, int set) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    if (set) ring[dl_slot(V, voice, r)] = lin[r]; else lin[r] = ring[dl_slot(V, voice, r)];
}

int free_bank(mxb_bank* b) {
    if (!b) return MXB_OK;
    for (int i = 0; i < MXB_P_COUNT; ++i) {
        cudaFree(b->dp_buf[i][0]); cudaFree(b->dp_buf[i][1]);
        if (b->ev_up[i]) cudaEventDestroy(b->ev_up[i]);
        for (int k = 0; k < 2; ++k) if (b->ev_read[i][k]) cudaEventDestroy(b->ev_read[i][k]);
    }
    if (b->up_stream) cudaStreamDestroy(b->up_stream);
    cudaFree(b->osc_out); cudaFree(b->f0); cudaFree(b->f1); cudaFree(b->f2);
    for (int i = 0; i < 5; ++i) cudaFree(b->cf[i]);
    cudaFree(b->env_amp); cudaFree(b->env_output); cudaFree(b->env_holdcount); cudaFree(b->env_hold); cudaFree(b->env_flags);
    cudaFree(b->trig_on); cudaFree(b->trig_off); cudaFree(b->dl_phase); cudaFree(b->dl_size); cudaFree(b->dl_pos); cudaFree(b->ring);
    cudaFree(b->partials); cudaFree(b->mix_dev); cudaFree(b->out_stage); cudaFree(b->fm_stage); cudaFree(b->cm_stage); cudaFree(b->ds_stage); cudaFree(b->tv_stage);
    delete b;
    return MXB_OK;
}

int upload_fill(double* dst, size_t n, double value) {
    std::vector<double> h(n, value);
    MXB_CUDA(cudaMemcpy(dst, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice));
    return MXB_OK;
}

}  // namespace

extern "C" {

int32_t mxb_bank_create(mxb_ctx* ctx, const mxb_bank_desc* d, mxb_bank** out) {
    MXB_REQUIRE(ctx && d && out, MXB_ERR_INVALID, "mxb_bank_create: NULL argument");
    *out = nullptr;
    MXB_REQUIRE(d->voices > 0, MXB_ERR_INVALID, "mxb_bank_create: voices %d", d->voices);
    MXB_REQUIRE(d->osc_kind >= MXB_OSC_SINEWAVE && d->osc_kind <= MXB_OSC_PHASORBETWEEN, MXB_ERR_INVALID, "mxb_bank_create: osc_kind %d", d->osc_kind);
    MXB_REQUIRE(d->filt_kind >= MXB_FILT_NONE && d->filt_kind <= MXB_FILT_BIQUAD, MXB_ERR_INVALID, "mxb_bank_create: filt_kind %d", d->filt_kind);
    MXB_REQUIRE(d->biquad_type >= MXB_BQ_LOWPASS && d->biquad_type <= MXB_BQ_HIGHSHELF, MXB_ERR_INVALID, "mxb_bank_create: biquad_type %d", d->biquad_type);
    MXB_REQUIRE(d->env_kind == MXB_ENV_NONE || d->env_kind == MXB_ENV_ADSR || d->env_kind == MXB_ENV_AR, MXB_ERR_INVALID, "mxb_bank_create: env_kind %d", d->env_kind);
    MXB_REQUIRE(d->delay_mode == MXB_DELAY_DL || d->delay_mode == MXB_DELAY_FROM_POSITION, MXB_ERR_INVALID, "mxb_bank_create: delay_mode %d", d->delay_mode);
    MXB_REQUIRE(d->delay_taps >= 0, MXB_ERR_INVALID, "mxb_bank_create: delay_taps %d", d->delay_taps);
    MXB_REQUIRE(d->max_frames > 0, MXB_ERR_INVALID, "mxb_bank_create: max_frames %d", d->max_frames);
    DeviceGuard g(ctx->device);
    MXB_REQUIRE(g.ok, MXB_ERR_CUDA, "mxb_bank_create: cudaSetDevice failed");
    mxb_bank* b = new (std::nothrow) mxb_bank();
    MXB_REQUIRE(b, MXB_ERR_ALLOC, "mxb_bank_create: out of host memory");
    b->ctx = ctx; b->desc = *d; b->V = d->voices;
    for (int i = 0; i < MXB_P_COUNT; ++i) {
        b->dp[i] = b->dp_buf[i][0] = b->dp_buf[i][1] = nullptr; b->dp_cur[i] = 0; b->dp_pending[i] = false;
        b->ev_up[i] = nullptr; b->ev_read[i][0] = b->ev_read[i][1] = nullptr; b->ev_read_set[i][0] = b->ev_read_set[i][1] = false;
    }
    b->up_stream = nullptr; b->ex = nullptr;
    b->osc_out = b->f0 = b->f1 = b->f2 = nullptr;
    for (auto& c : b->cf) c = nullptr;
    b->env_amp = b->env_output = nullptr; b->env_holdcount = b->env_hold = nullptr; b->env_flags = nullptr;
    b->trig_on = b->trig_off = b->dl_phase = b->dl_size = b->dl_pos = nullptr;
    b->ring = b->partials = b->mix_dev = nullptr; b->partials_len = 0;
    b->out_stage = nullptr; b->out_stage_bytes = 0; b->launches = 0;
    b->fm_stage = nullptr; b->fm_stage_bytes = 0;
    b->cm_stage = nullptr; b->cm_stage_bytes = 0;
    b->ds_stage = nullptr; b->ds_stage_bytes = 0;
    b->tv_stage = nullptr; b->tv_stage_bytes = 0;
    const size_t V = (size_t)d->voices;
    int rc = MXB_OK;
#define TRY(x) do { rc = (x); if (rc != MXB_OK) { free_bank(b); return rc; } } while (0)
    static const double defaults[MXB_P_COUNT] = {0, 0, 0.5, 1000.0, 1.0, 0, 0, 0, 0, 0, 1.0, 1.0, 0, 0.5, 0, 0, 1.0};
    for (int i = 0; i < MXB_P_COUNT; ++i) {
        TRY(dev_alloc(&b->dp_buf[i][0], V));
        b->dp[i] = b->dp_buf[i][0];
        b->hp[i].assign(V, defaults[i]);
        if (defaults[i] != 0.0) TRY(upload_fill(b->dp[i], V, defaults[i]));
    }
    TRY(dev_alloc(&b->osc_out, V));
    TRY(dev_alloc(&b->f0, V)); TRY(dev_alloc(&b->f1, V)); TRY(dev_alloc(&b->f2, V));
    for (int i = 0; i < 5; ++i) TRY(dev_alloc(&b->cf[i], V));
    TRY(dev_alloc(&b->env_amp, V)); TRY(dev_alloc(&b->env_output, V));
    TRY(dev_alloc(&b->env_holdcount, V)); TRY(dev_alloc(&b->env_hold, V)); TRY(dev_alloc(&b->env_flags, V));
    {
        std::vector<long long> one(V, 1);   // maxiEnv::holdtime = 1, src/maximilian.h:913
        cudaError_t e = cudaMemcpy(b->env_hold, one.data(), sizeof(long long) * V, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("cudaMemcpy: %s", cudaGetErrorString(e)); free_bank(b); return MXB_ERR_CUDA; }
    }
    TRY(dev_alloc(&b->trig_on, V)); TRY(dev_alloc(&b->trig_off, V));
    TRY(dev_alloc(&b->mix_dev, (size_t)d->max_frames * 2));
    if (d->delay_taps > 0) {
        TRY(dev_alloc(&b->dl_phase, V));
        TRY(dev_alloc(&b->dl_size, V));
        TRY(dev_alloc(&b->dl_pos, V));
        std::vector<int> one(V, 1);
        cudaError_t e = cudaMemcpy(b->dl_size, one.data(), sizeof(int) * V, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { set_error("cudaMemcpy: %s", cudaGetErrorString(e)); free_bank(b); return MXB_ERR_CUDA; }
        TRY(dev_alloc(&b->ring, dl_ring_doubles(V, d->delay_taps)));   // zeroed: maxiDelayline ctor memset, src/maximilian.cpp:415-417
    }
    if (d->filt_kind == MXB_FILT_SVF) {          // maxiSVF ctor: setParams(1000, 1), src/maximilian.h:1284
        b->set_mask[MXB_P_CUTOFF] = b->set_mask[MXB_P_RESONANCE] = true;
        TRY(redesign(b));
        b->set_mask[MXB_P_CUTOFF] = b->set_mask[MXB_P_RESONANCE] = false;
    }
#undef TRY
    *out = b;
    return MXB_OK;
}

int32_t mxb_bank_destroy(mxb_bank* b) {
    if (!b) return MXB_OK;
    DeviceGuard g(b->ctx->device);
    cudaDeviceSynchronize();
    return free_bank(b);
}

int32_t mxb_bank_voices(const mxb_bank* b) { return b ? b->V : MXB_ERR_INVALID; }

int32_t mxb_bank_set_exchange(mxb_bank* b, mxb_exchange* ex) {
    MXB_REQUIRE(b, MXB_ERR_INVALID, "mxb_bank_set_exchange: NULL bank");
    if (ex) MXB_REQUIRE(ex->ctx->device == b->ctx->device, MXB_ERR_INVALID, "mxb_bank_set_exchange: exchange and bank live on different devices");
    b->ex = ex;
    return MXB_OK;
}
int64_t mxb_bank_launch_count(const mxb_bank* b) { return b ? b->launches : 0; }

int32_t mxb_bank_set_param(mxb_bank* b, int32_t id, const double* values, int32_t mem) {
    MXB_REQUIRE(b && values, MXB_ERR_INVALID, "mxb_bank_set_param: NULL argument");
    MXB_REQUIRE(id >= 0 && id < MXB_P_COUNT, MXB_ERR_INVALID, "mxb_bank_set_param: unknown id %d", id);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_set_param: mem %d", mem);
    DeviceGuard g(b->ctx->device);
    const size_t V = (size_t)b->V, bytes = sizeof(double) * V;
    // Parameters change between blocks. The copies below run on the legacy default stream: they are ordered after every
    // block already enqueued on it or on any other blocking stream; a caller driving the bank from a NON-blocking stream
    // synchronises that stream first. An asynchronous upload still in flight for this id is superseded.
    if (b->dp_pending[id]) { MXB_CUDA(cudaEventSynchronize(b->ev_up[id])); b->dp_pending[id] = false; }
    if (mem == MXB_MEM_HOST) {
        memcpy(b->hp[id].data(), values, bytes);
        MXB_CUDA(cudaMemcpy(b->dp[id], values, bytes, cudaMemcpyHostToDevice));
    } else {
        MXB_CUDA(cudaMemcpy(b->dp[id], values, bytes, cudaMemcpyDeviceToDevice));
        MXB_CUDA(cudaMemcpy(b->hp[id].data(), values, bytes, cudaMemcpyDeviceToHost));
    }
    b->set_mask[id] = true;
    if (id == MXB_P_CUTOFF || id == MXB_P_RESONANCE || id == MXB_P_GAIN) {
        int rc = redesign(b);
        if (rc != MXB_OK) return rc;
    } else if (id == MXB_P_ENV_HOLDTIME) {
        std::vector<long long> h(V);
        for (size_t v = 0; v < V; ++v) {
            // the kernels count holdcount/holdtime in 32-bit registers (holdcount never passes holdtime)
            MXB_REQUIRE(fabs(b->hp[id][v]) < 2147483648.0, MXB_ERR_INVALID, "mxb_bank_set_param: holdtime %.0f of voice %zu does not fit 32 bits", b->hp[id][v], v);
            h[v] = (long long)b->hp[id][v];
        }
        MXB_CUDA(cudaMemcpy(b->env_hold, h.data(), sizeof(long long) * V, cudaMemcpyHostToDevice));
    } else if (id == MXB_P_DELAY_POSITION && b->dl_pos) {
        std::vector<int> h(V);
        for (size_t v = 0; v < V; ++v) {
            const double s = b->hp[id][v];
            MXB_REQUIRE(fabs(s) < 2147483648.0, MXB_ERR_INVALID, "mxb_bank_set_param: delay position %.0f of voice %zu does not fit an int", s, v);
            h[v] = (int)s;
        }
        MXB_CUDA(cudaMemcpy(b->dl_pos, h.data(), sizeof(int) * V, cudaMemcpyHostToDevice));
    } else if (id == MXB_P_DELAY_SIZE && b->dl_size) {
        std::vector<int> h(V);
        for (size_t v = 0; v < V; ++v) {
            const double s = b->hp[id][v];
            // the reference takes any int (size <= 0 pins the index to slot 0); NaN / out-of-int values have no int conversion
            MXB_REQUIRE(s == s && s > -2147483649.0 && s <= (double)b->desc.delay_taps, MXB_ERR_INVALID,
                        "mxb_bank_set_param: delay size %g of voice %zu is not an int <= delay_taps %d", s, v, b->desc.delay_taps);
            h[v] = (int)s;
        }
        MXB_CUDA(cudaMemcpy(b->dl_size, h.data(), sizeof(int) * V, cudaMemcpyHostToDevice));
    }
    return MXB_OK;
}

int32_t mxb_bank_set_param_async(mxb_bank* b, int32_t id, const double* values, int32_t mem, void* stream_) {
    MXB_REQUIRE(b && values, MXB_ERR_INVALID, "mxb_bank_set_param_async: NULL argument");
    MXB_REQUIRE(id >= 0 && id < MXB_P_COUNT, MXB_ERR_INVALID, "mxb_bank_set_param_async: unknown id %d", id);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_set_param_async: mem %d", mem);
    const bool needs_host_pass = id == MXB_P_CUTOFF || id == MXB_P_RESONANCE || id == MXB_P_GAIN ||
                                 id == MXB_P_ENV_HOLDTIME || id == MXB_P_DELAY_SIZE || id == MXB_P_DELAY_POSITION;
    if (needs_host_pass) return mxb_bank_set_param(b, id, values, mem);      // coefficient design / integer conversion on the host
    DeviceGuard g(b->ctx->device);
    const size_t bytes = sizeof(double) * (size_t)b->V;
    if (mem == MXB_MEM_DEVICE) {
        // the source was produced on `stream`: one stream-ordered device copy into the buffer the next block reads
        MXB_CUDA(cudaMemcpyAsync(b->dp[id], values, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
        b->set_mask[id] = true;
        return MXB_OK;
    }
    // Host source: the upload goes to the buffer the running / enqueued blocks do NOT read, on the bank's own copy stream,
    // so it overlaps them; the next mxb_bank_process waits for it (event) and switches over. The buffer being overwritten
    // was last read two blocks ago: the copy waits for that kernel only.
    if (!b->up_stream) MXB_CUDA(cudaStreamCreateWithFlags(&b->up_stream, cudaStreamNonBlocking));
    if (!b->ev_up[id]) MXB_CUDA(cudaEventCreateWithFlags(&b->ev_up[id], cudaEventDisableTiming));
    const int tgt = 1 - b->dp_cur[id];
    if (!b->dp_buf[id][tgt]) { int rc = dev_alloc(&b->dp_buf[id][tgt], (size_t)b->V, false); if (rc != MXB_OK) return rc; }
    if (b->ev_read_set[id][tgt]) MXB_CUDA(cudaStreamWaitEvent(b->up_stream, b->ev_read[id][tgt], 0));
    MXB_CUDA(cudaMemcpyAsync(b->dp_buf[id][tgt], values, bytes, cudaMemcpyHostToDevice, b->up_stream));
    MXB_CUDA(cudaEventRecord(b->ev_up[id], b->up_stream));
    b->dp_pending[id] = true;
    b->set_mask[id] = true;
    return MXB_OK;
}

int32_t mxb_bank_get_state(mxb_bank* b, int32_t id, double* values, int32_t mem) {
    MXB_REQUIRE(b && values, MXB_ERR_INVALID, "mxb_bank_get_state: NULL argument");
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_get_state: mem %d", mem);
    DeviceGuard g(b->ctx->device);
    MXB_CUDA(cudaDeviceSynchronize());
    const size_t V = (size_t)b->V;
    const cudaMemcpyKind kind = mem == MXB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    const double* src = nullptr;
    if (id >= 0 && id < MXB_P_COUNT) src = b->dp[id];
    else if (id == MXB_S_FILT_0) src = b->f0;
    else if (id == MXB_S_FILT_1) src = b->f1;
    else if (id == MXB_S_FILT_2) src = b->f2;
    else if (id == MXB_S_ENV_AMPLITUDE) src = b->env_amp;
    else if (id == MXB_S_ENV_OUTPUT) src = b->env_output;
    else if (id == MXB_S_OSC_OUTPUT) src = b->osc_out;
    if (src) { MXB_CUDA(cudaMemcpy(values, src, sizeof(double) * V, kind)); return MXB_OK; }
    // integer state, returned as doubles
    std::vector<double> h(V, 0.0);
    if (id == MXB_S_ENV_HOLDCOUNT) {
        std::vector<long long> t(V);
        MXB_CUDA(cudaMemcpy(t.data(), b->env_holdcount, sizeof(long long) * V, cudaMemcpyDeviceToHost));
        for (size_t v = 0; v < V; ++v) h[v] = (double)t[v];
    } else if (id == MXB_S_ENV_FLAGS || id == MXB_S_DELAY_PHASE) {
        const int* s = id == MXB_S_ENV_FLAGS ? b->env_flags : b->dl_phase;
        if (s) {
            std::vector<int> t(V);
            MXB_CUDA(cudaMemcpy(t.data(), s, sizeof(int) * V, cudaMemcpyDeviceToHost));
            for (size_t v = 0; v < V; ++v) h[v] = (double)t[v];
        }
    } else {
        set_error("mxb_bank_get_state: unknown id %d", id);
        return MXB_ERR_INVALID;
    }
    MXB_CUDA(cudaMemcpy(values, h.data(), sizeof(double) * V, mem == MXB_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyHostToDevice));
    return MXB_OK;
}

int32_t mxb_bank_get_ring(mxb_bank* b, int32_t voice, double* dst, int32_t n, int32_t mem) {
    MXB_REQUIRE(b && dst, MXB_ERR_INVALID, "mxb_bank_get_ring: NULL argument");
    MXB_REQUIRE(b->ring, MXB_ERR_STATE, "mxb_bank_get_ring: bank has no delay line");
    MXB_REQUIRE(voice >= 0 && voice < b->V && n >= 0 && n <= b->desc.delay_taps, MXB_ERR_INVALID, "mxb_bank_get_ring: voice %d n %d", voice, n);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_get_ring: mem %d", mem);
    if (n == 0) return MXB_OK;
    DeviceGuard g(b->ctx->device);
    // the ring is chunk-interleaved and swizzled (delay_kernels.cuh): one small kernel walks the voice's slots
    double* d_lin = dst;
    if (mem == MXB_MEM_HOST) { int rc = dev_alloc(&d_lin, (size_t)n, false); if (rc != MXB_OK) return rc; }
    ring_linear_kernel<<<(n + 255) / 256, 256>>>(b->ring, (size_t)b->V, (size_t)voice, d_lin, n, 0);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && mem == MXB_MEM_HOST) e = cudaMemcpy(dst, d_lin, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost);
    else if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (mem == MXB_MEM_HOST) cudaFree(d_lin);
    if (e != cudaSuccess) { set_error("mxb_bank_get_ring: %s", cudaGetErrorString(e)); return MXB_ERR_CUDA; }
    return MXB_OK;
}

int32_t mxb_bank_set_ring(mxb_bank* b, int32_t voice, const double* src, int32_t n, int32_t mem) {
    MXB_REQUIRE(b && src, MXB_ERR_INVALID, "mxb_bank_set_ring: NULL argument");
    MXB_REQUIRE(b->ring, MXB_ERR_STATE, "mxb_bank_set_ring: bank has no delay line");
    MXB_REQUIRE(voice >= 0 && voice < b->V && n >= 0 && n <= b->desc.delay_taps, MXB_ERR_INVALID, "mxb_bank_set_ring: voice %d n %d", voice, n);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_set_ring: mem %d", mem);
    if (n == 0) return MXB_OK;
    DeviceGuard g(b->ctx->device);
    double* d_lin = const_cast<double*>(src);
    if (mem == MXB_MEM_HOST) {
        int rc = dev_alloc(&d_lin, (size_t)n, false); if (rc != MXB_OK) return rc;
        cudaError_t e0 = cudaMemcpy(d_lin, src, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice);
        if (e0 != cudaSuccess) { cudaFree(d_lin); set_error("mxb_bank_set_ring: %s", cudaGetErrorString(e0)); return MXB_ERR_CUDA; }
    }
    ring_linear_kernel<<<(n + 255) / 256, 256>>>(b->ring, (size_t)b->V, (size_t)voice, d_lin, n, 1);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (mem == MXB_MEM_HOST) cudaFree(d_lin);
    if (e != cudaSuccess) { set_error("mxb_bank_set_ring: %s", cudaGetErrorString(e)); return MXB_ERR_CUDA; }
    return MXB_OK;
}

int32_t mxb_bank_set_state(mxb_bank* b, int32_t id, const double* values, int32_t mem) {
    MXB_REQUIRE(b && values, MXB_ERR_INVALID, "mxb_bank_set_state: NULL argument");
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_bank_set_state: mem %d", mem);
    DeviceGuard g(b->ctx->device);
    MXB_CUDA(cudaDeviceSynchronize());
    const size_t V = (size_t)b->V;
    double* dst = nullptr;
    if (id == MXB_S_FILT_0) dst = b->f0;
    else if (id == MXB_S_FILT_1) dst = b->f1;
    else if (id == MXB_S_FILT_2) dst = b->f2;
    else if (id == MXB_S_ENV_AMPLITUDE) dst = b->env_amp;
    else if (id == MXB_S_ENV_OUTPUT) dst = b->env_output;
    else if (id == MXB_S_OSC_OUTPUT) dst = b->osc_out;
    if (dst) {
        MXB_CUDA(cudaMemcpy(dst, values, sizeof(double) * V, mem == MXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
        return MXB_OK;
    }
    MXB_REQUIRE(id == MXB_S_ENV_HOLDCOUNT || id == MXB_S_ENV_FLAGS || id == MXB_S_DELAY_PHASE, MXB_ERR_INVALID,
                "mxb_bank_set_state: unknown id %d (parameters go through mxb_bank_set_param)", id);
    std::vector<double> h(V);                     // integer state arrives as doubles, like mxb_bank_get_state returns it
    MXB_CUDA(cudaMemcpy(h.data(), values, sizeof(double) * V, mem == MXB_MEM_HOST ? cudaMemcpyHostToHost : cudaMemcpyDeviceToHost));
    if (id == MXB_S_ENV_HOLDCOUNT) {
        std::vector<long long> t(V);
        for (size_t v = 0; v < V; ++v) {
            MXB_REQUIRE(fabs(h[v]) < 2147483648.0, MXB_ERR_INVALID, "mxb_bank_set_state: holdcount[%zu] = %g", v, h[v]);
            t[v] = (long long)h[v];
        }
        MXB_CUDA(cudaMemcpy(b->env_holdcount, t.data(), sizeof(long long) * V, cudaMemcpyHostToDevice));
    } else {
        int* d = id == MXB_S_ENV_FLAGS ? b->env_flags : b->dl_phase;
        MXB_REQUIRE(d, MXB_ERR_STATE, "mxb_bank_set_state: bank has no delay line");
        std::vector<int> t(V);
        for (size_t v = 0; v < V; ++v) {
            MXB_REQUIRE(fabs(h[v]) < 2147483648.0, MXB_ERR_INVALID, "mxb_bank_set_state: value[%zu] = %g", v, h[v]);
            // a negative ring index is an out-of-bounds access in the reference (memory[phase]); refused here
            if (id == MXB_S_DELAY_PHASE) MXB_REQUIRE(h[v] >= 0.0, MXB_ERR_INVALID, "mxb_bank_set_state: delay phase[%zu] = %g is negative", v, h[v]);
            t[v] = (int)h[v];
            if (id == MXB_S_ENV_FLAGS) t[v] &= 31;
        }
        MXB_CUDA(cudaMemcpy(d, t.data(), sizeof(int) * V, cudaMemcpyHostToDevice));
    }
    return MXB_OK;
}

// Deep copy: parameters, designed coefficients, every state array and the delay rings -- what copying an array of the
// reference's objects (they hold their state by value, src/maximilian.h:266-281) does. The copy lives on the same context.
int32_t mxb_bank_clone(mxb_bank* src, mxb_bank** out) {
    MXB_REQUIRE(src && out, MXB_ERR_INVALID, "mxb_bank_clone: NULL argument");
    *out = nullptr;
    mxb_bank* b = nullptr;
    int rc = mxb_bank_create(src->ctx, &src->desc, &b);
    if (rc != MXB_OK) return rc;
    DeviceGuard g(src->ctx->device);
    const size_t V = (size_t)src->V;
    cudaError_t e = cudaDeviceSynchronize();
    auto cp = [&](void* d, const void* s, size_t bytes) { if (e == cudaSuccess && d && s) e = cudaMemcpy(d, s, bytes, cudaMemcpyDeviceToDevice); };
    for (int i = 0; i < MXB_P_COUNT; ++i) { b->hp[i] = src->hp[i]; b->set_mask[i] = src->set_mask[i]; cp(b->dp[i], src->dp[i], sizeof(double) * V); }   // (an upload still in flight on the source is not part of its state yet)
    cp(b->osc_out, src->osc_out, sizeof(double) * V);
    cp(b->f0, src->f0, sizeof(double) * V); cp(b->f1, src->f1, sizeof(double) * V); cp(b->f2, src->f2, sizeof(double) * V);
    for (int i = 0; i < 5; ++i) cp(b->cf[i], src->cf[i], sizeof(double) * V);
    cp(b->env_amp, src->env_amp, sizeof(double) * V); cp(b->env_output, src->env_output, sizeof(double) * V);
    cp(b->env_holdcount, src->env_holdcount, sizeof(long long) * V); cp(b->env_hold, src->env_hold, sizeof(long long) * V);
    cp(b->env_flags, src->env_flags, sizeof(int) * V);
    if (src->desc.delay_taps > 0) {
        cp(b->dl_phase, src->dl_phase, sizeof(int) * V); cp(b->dl_size, src->dl_size, sizeof(int) * V); cp(b->dl_pos, src->dl_pos, sizeof(int) * V);
        cp(b->ring, src->ring, sizeof(double) * dl_ring_doubles(V, src->desc.delay_taps));
    }
    if (e != cudaSuccess) { set_error("mxb_bank_clone: %s", cudaGetErrorString(e)); mxb_bank_destroy(b); return MXB_ERR_CUDA; }
    *out = b;
    return MXB_OK;
}

int32_t mxb_bank_process(mxb_bank* b, int32_t n_frames, const int32_t* trig_on, const int32_t* trig_off,
                         void* out, int32_t out_dtype, double* mix, int32_t mem, void* stream_) {
    return mxb_bank_process_fm(b, n_frames, nullptr, trig_on, trig_off, out, out_dtype, mix, mem, stream_);
}

int32_t mxb_bank_process_fm(mxb_bank* b, int32_t n_frames, const double* freq_tv, const int32_t* trig_on, const int32_t* trig_off,
                            void* out, int32_t out_dtype, double* mix, int32_t mem, void* stream_) {
    mxb_modulation m; m.freq_tv = freq_tv; m.cutoff_tv = nullptr; m.delay_size_tv = nullptr; m.trig_tv = nullptr;
    return mxb_bank_process_mod(b, n_frames, &m, trig_on, trig_off, out, out_dtype, mix, mem, stream_);
}

int32_t mxb_bank_process_mod(mxb_bank* b, int32_t n_frames, const mxb_modulation* mod, const int32_t* trig_on, const int32_t* trig_off,
                             void* out, int32_t out_dtype, double* mix, int32_t mem, void* stream_) {
    NvtxRange nvtx_("mxb_bank_process_mod");
    MXB_REQUIRE(b, MXB_ERR_INVALID, "mxb_bank_process: NULL bank");
    const double* freq_tv = mod ? mod->freq_tv : nullptr;
    const double* cutoff_tv = mod ? mod->cutoff_tv : nullptr;
    const double* dsize_tv = mod ? mod->delay_size_tv : nullptr;
    const uint8_t* trig_tv = mod ? mod->trig_tv : nullptr;
    if (trig_tv) {
        MXB_REQUIRE(b->desc.env_kind != MXB_ENV_NONE, MXB_ERR_INVALID, "mxb_bank_process_mod: trig_tv on a bank without an envelope stage");
        MXB_REQUIRE(!trig_on && !trig_off, MXB_ERR_INVALID, "mxb_bank_process_mod: trig_tv replaces trig_on / trig_off (pass NULL for both)");
    }
    MXB_REQUIRE(n_frames >= 0 && n_frames <= b->desc.max_frames, MXB_ERR_INVALID, "mxb_bank_process: n_frames %d (max_frames %d)", n_frames, b->desc.max_frames);
    const bool async_host = (mem & MXB_MEM_ASYNC) != 0;   // host buffers, but return as soon as everything is enqueued
    mem &= ~MXB_MEM_ASYNC;
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE || mem == MXB_MEM_SPLIT, MXB_ERR_INVALID, "mxb_bank_process: mem %d", mem);
    const bool host_ctl = mem != MXB_MEM_DEVICE;      // gates and mix in host memory
    const bool host_out = mem == MXB_MEM_HOST;        // out in host memory
    MXB_REQUIRE(out_dtype == MXB_F64 || out_dtype == MXB_F32, MXB_ERR_INVALID, "mxb_bank_process: out_dtype %d", out_dtype);
    MXB_REQUIRE((trig_on == nullptr) == (trig_off == nullptr), MXB_ERR_INVALID, "mxb_bank_process: trig_on/trig_off must both be given or both NULL");
    MXB_REQUIRE(out || mix, MXB_ERR_INVALID, "mxb_bank_process: neither out nor mix requested");
    const int fk = b->desc.filt_kind;
    if (fk == MXB_FILT_LORES || fk == MXB_FILT_HIRES)
        MXB_REQUIRE(b->set_mask[MXB_P_CUTOFF] && b->set_mask[MXB_P_RESONANCE], MXB_ERR_STATE,
                    "mxb_bank_process: lores/hires need MXB_P_CUTOFF and MXB_P_RESONANCE (they are call arguments in the reference)");
    if (mix && b->ex) {     // before any kernel runs: a refused call must leave phase / filter / envelope / ring state untouched
        MXB_REQUIRE(b->ex->connected, MXB_ERR_STATE, "mxb_bank_process: the attached exchange is not connected to its peers");
        MXB_REQUIRE(n_frames * 2 <= b->ex->max_doubles, MXB_ERR_INVALID, "mxb_bank_process: exchange holds %d values, the bus needs %d", b->ex->max_doubles, n_frames * 2);
    }
    if (n_frames == 0) return MXB_OK;
    DeviceGuard g(b->ctx->device);
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t V = (size_t)b->V;
    const size_t esz = out_dtype == MXB_F32 ? 4 : 8;

    const int* d_on = trig_on; const int* d_off = trig_off;
    void* d_out = out; double* d_mix = mix;
    if (cutoff_tv)
        MXB_REQUIRE(fk == MXB_FILT_LORES || fk == MXB_FILT_HIRES || fk == MXB_FILT_SVF, MXB_ERR_UNSUPPORTED,
                    "mxb_bank_process_mod: per-sample cutoff is built for lores / hires / maxiSVF");
    if (dsize_tv)
        MXB_REQUIRE(b->desc.delay_taps > 0, MXB_ERR_UNSUPPORTED, "mxb_bank_process_mod: delay_size_tv on a bank without a delay line");
    // per-sample parameters are control data: they live where the gates live
    auto stage = [&](const double* src, double** buf, size_t* cap, const double** dev) -> int {
        *dev = src;
        if (!src || !host_ctl) return MXB_OK;
        const size_t need = sizeof(double) * (size_t)n_frames * V;
        if (need > *cap) {
            MXB_CUDA(cudaStreamSynchronize(s));
            cudaFree(*buf); *buf = nullptr; *cap = 0;
            cudaError_t e = cudaMalloc((void**)buf, need);
            if (e != cudaSuccess) { set_error("mxb_bank_process_mod: staging cudaMalloc(%zu): %s", need, cudaGetErrorString(e)); return MXB_ERR_ALLOC; }
            *cap = need;
        }
        MXB_CUDA(cudaMemcpyAsync(*buf, src, need, cudaMemcpyHostToDevice, s));
        *dev = *buf;
        return MXB_OK;
    };
    const double *d_fm = nullptr, *d_cm = nullptr;
    { int rc0 = stage(freq_tv, &b->fm_stage, &b->fm_stage_bytes, &d_fm); if (rc0 != MXB_OK) return rc0; }
    { int rc0 = stage(cutoff_tv, &b->cm_stage, &b->cm_stage_bytes, &d_cm); if (rc0 != MXB_OK) return rc0; }
    const double* d_ds = nullptr;
    { int rc0 = stage(dsize_tv, &b->ds_stage, &b->ds_stage_bytes, &d_ds); if (rc0 != MXB_OK) return rc0; }
    const unsigned char* d_tv = trig_tv;
    if (trig_tv && host_ctl) {
        const size_t need = (size_t)n_frames * V;
        if (need > b->tv_stage_bytes) {
            MXB_CUDA(cudaStreamSynchronize(s));
            cudaFree(b->tv_stage); b->tv_stage = nullptr; b->tv_stage_bytes = 0;
            cudaError_t e = cudaMalloc((void**)&b->tv_stage, need);
            if (e != cudaSuccess) { set_error("mxb_bank_process_mod: staging cudaMalloc(%zu): %s", need, cudaGetErrorString(e)); return MXB_ERR_ALLOC; }
            b->tv_stage_bytes = need;
        }
        MXB_CUDA(cudaMemcpyAsync(b->tv_stage, trig_tv, need, cudaMemcpyHostToDevice, s));
        d_tv = b->tv_stage;
    }
    if (host_ctl) {
        if (trig_on) {
            MXB_CUDA(cudaMemcpyAsync(b->trig_on, trig_on, sizeof(int) * V, cudaMemcpyHostToDevice, s));
            MXB_CUDA(cudaMemcpyAsync(b->trig_off, trig_off, sizeof(int) * V, cudaMemcpyHostToDevice, s));
            d_on = b->trig_on; d_off = b->trig_off;
        }
        if (out && host_out) {
            const size_t need = (size_t)n_frames * V * esz;
            if (need > b->out_stage_bytes) {
                MXB_CUDA(cudaStreamSynchronize(s));
                cudaFree(b->out_stage); b->out_stage = nullptr; b->out_stage_bytes = 0;
                cudaError_t e = cudaMalloc(&b->out_stage, need);
                if (e != cudaSuccess) { set_error("mxb_bank_process: staging cudaMalloc(%zu): %s", need, cudaGetErrorString(e)); return MXB_ERR_ALLOC; }
                b->out_stage_bytes = need;
            }
            d_out = b->out_stage;
        }
        if (mix) d_mix = b->mix_dev;
    }

    const int per_cta = kBankBlock * kBankVPT;
    const int grid = (int)((V + per_cta - 1) / per_cta);
    // warps that own at least one voice (warps past the end of the bank exit without writing partials)
    const int W = b->desc.delay_taps > 0 ? delay_bank_warps(b->V) : (int)((V + 32 * kBankVPT - 1) / (32 * kBankVPT));
    if (mix) {
        const size_t need = (size_t)b->desc.max_frames * 2 * (size_t)W;
        if (need > b->partials_len) {
            MXB_CUDA(cudaStreamSynchronize(s));
            cudaFree(b->partials); b->partials = nullptr; b->partials_len = 0;
            int rc = dev_alloc(&b->partials, need, false);
            if (rc != MXB_OK) return rc;
            b->partials_len = need;
        }
    }

    for (int id = 0; id < MXB_P_COUNT; ++id) {
        if (!b->dp_pending[id]) continue;                 // an asynchronous host upload for this block: wait for it on the stream, switch over
        MXB_CUDA(cudaStreamWaitEvent(s, b->ev_up[id], 0));
        b->dp_cur[id] ^= 1; b->dp[id] = b->dp_buf[id][b->dp_cur[id]]; b->dp_pending[id] = false;
    }
    BankArgs a;
    memset(&a, 0, sizeof(a));
    a.V = b->V; a.n_frames = n_frames; a.osc_kind = b->desc.osc_kind;
    a.out_f32 = out_dtype == MXB_F32;
    a.vec_ok = (V % kBankVPT == 0) && d_out && (((uintptr_t)d_out) % (esz * kBankVPT) == 0);
    a.W = W;
    a.sr = (double)(size_t)b->ctx->sample_rate;
    for (int i = 0; i < 4; ++i) a.svf_mix[i] = b->desc.svf_mix[i];
    a.freq_tv = d_fm; a.cutoff_tv = d_cm; a.res = b->dp[MXB_P_RESONANCE];
    a.freq = b->dp[MXB_P_FREQ]; a.duty = b->dp[MXB_P_DUTY]; a.pstart = b->dp[MXB_P_PHASOR_START]; a.pend = b->dp[MXB_P_PHASOR_END]; a.phase = b->dp[MXB_P_PHASE]; a.osc_out = b->osc_out;
    a.f0 = b->f0; a.f1 = b->f1; a.f2 = b->f2;
    for (int i = 0; i < 5; ++i) a.cf[i] = b->cf[i];
    a.env_att = b->dp[MXB_P_ENV_ATTACK]; a.env_dec = b->dp[MXB_P_ENV_DECAY];
    a.env_sus = b->dp[MXB_P_ENV_SUSTAIN]; a.env_rel = b->dp[MXB_P_ENV_RELEASE];
    a.env_hold = b->env_hold; a.env_amp = b->env_amp; a.env_output = b->env_output;
    a.env_holdcount = b->env_holdcount; a.env_flags = b->env_flags;
    a.trig_on = d_on; a.trig_off = d_off; a.trig_tv = d_tv;
    a.out = d_out; a.pan = b->dp[MXB_P_PAN]; a.partials = b->partials;

    int osc_t = OSC_T_GENERIC;
    if (b->desc.osc_kind == MXB_OSC_SINEWAVE) osc_t = OSC_T_SINE;
    else if (b->desc.osc_kind == MXB_OSC_PHASOR) osc_t = OSC_T_PHASOR;
    else if (b->desc.osc_kind == MXB_OSC_SAW) osc_t = OSC_T_SAW;
    const int env = b->desc.env_kind != MXB_ENV_NONE ? 1 : 0;
    a.env_ar = b->desc.env_kind == MXB_ENV_AR ? 1 : 0;
    const bool svf_lp = fk == MXB_FILT_SVF && b->desc.svf_mix[0] == 1.0 && b->desc.svf_mix[1] == 0.0 &&
                        b->desc.svf_mix[2] == 0.0 && b->desc.svf_mix[3] == 0.0;

    int rc;
    if (b->desc.delay_taps > 0) {
        DelayArgs da;
        da.phase = b->dl_phase; da.size = b->dl_size; da.feedback = b->dp[MXB_P_DELAY_FEEDBACK];
        da.ring = b->ring; da.taps = b->desc.delay_taps; da.W_out = 0;
        da.from_position = b->desc.delay_mode == MXB_DELAY_FROM_POSITION ? 1 : 0; da.position = b->dl_pos;
        da.size_tv = d_ds;
        rc = launch_delay_bank(a, da, fk, svf_lp, env, out != nullptr, mix != nullptr, s);
        if (rc != MXB_OK) return rc;
    } else {
        bank_launch_fn fn = launch_bank_none;
        switch (fk) {
            case MXB_FILT_LORES: fn = launch_bank_lores; break;
            case MXB_FILT_HIRES: fn = launch_bank_hires; break;
            case MXB_FILT_SVF: fn = svf_lp ? launch_bank_svf_lp : launch_bank_svf; break;
            case MXB_FILT_BIQUAD: fn = launch_bank_biquad; break;
            default: break;
        }
        const size_t smem = mix ? sizeof(double) * (kBankBlock / 32) * 2 * kMixTT * 33 : 0;
        rc = fn(a, osc_t, env, out != nullptr, mix != nullptr, grid, smem, s);
        if (rc != MXB_OK) return rc;
    }
    b->launches += 1;
    for (int id = 0; id < MXB_P_COUNT; ++id) {            // double-buffered parameters: remember when this block's reads are over
        if (!b->dp_buf[id][1]) continue;
        const int c = b->dp_cur[id];
        if (!b->ev_read[id][c]) MXB_CUDA(cudaEventCreateWithFlags(&b->ev_read[id][c], cudaEventDisableTiming));
        MXB_CUDA(cudaEventRecord(b->ev_read[id][c], s));
        b->ev_read_set[id][c] = true;
    }
    if (mix) {
        const int rows = n_frames * 2;
        const int threads = kMixReduceThreads, blocks = rows;
        if (b->ex) {
            mix_reduce_exchange_kernel<<<blocks, threads, 0, s>>>(b->partials, d_mix, rows, a.W, exchange_next(b->ex));
        } else {
            mix_reduce_kernel<<<blocks, threads, 0, s>>>(b->partials, d_mix, rows, a.W);
        }
        MXB_CUDA(cudaGetLastError());
        b->launches += 1;
    }
    if (host_ctl) {
        if (out && host_out) MXB_CUDA(cudaMemcpyAsync(out, d_out, (size_t)n_frames * V * esz, cudaMemcpyDeviceToHost, s));
        if (mix) MXB_CUDA(cudaMemcpyAsync(mix, d_mix, sizeof(double) * (size_t)n_frames * 2, cudaMemcpyDeviceToHost, s));
        if (async_host) return MXB_OK;            // results are valid once `stream` (or the context) has been synchronised
        MXB_CUDA(cudaStreamSynchronize(s));
        if (mix && b->ex) {       // the call is synchronous in this mode: a timed-out exchange is an error of THIS call
            unsigned int m = 0;
            MXB_CUDA(cudaMemcpy(&m, b->ex->status, sizeof(m), cudaMemcpyDeviceToHost));
            MXB_REQUIRE(m == 0, MXB_ERR_STATE, "mxb_bank_process: mix exchange timed out waiting for rank mask 0x%x (bus incomplete)", m);
        }
    }
    return MXB_OK;
}

// The block-dispatch shim: what routing() does for a reference patch (cpp/commandline/player.cpp:25-44) -- fill the audio
// driver's interleaved RTAUDIO_FLOAT64 buffer [n_frames][channels] with the next block: the bank's stereo bus goes to channels
// 0 and 1 (a mono device gets the left bus), further channels are silent.
int32_t mxb_play_block(mxb_bank* b, double* interleaved_out, int32_t n_frames, int32_t channels) {
    MXB_REQUIRE(b && interleaved_out, MXB_ERR_INVALID, "mxb_play_block: NULL argument");
    MXB_REQUIRE(channels >= 1 && n_frames >= 0, MXB_ERR_INVALID, "mxb_play_block: channels %d n_frames %d", channels, n_frames);
    if (channels == 2) return mxb_bank_process(b, n_frames, nullptr, nullptr, nullptr, MXB_F64, interleaved_out, MXB_MEM_HOST, nullptr);
    std::vector<double> bus((size_t)n_frames * 2);
    int rc = mxb_bank_process(b, n_frames, nullptr, nullptr, nullptr, MXB_F64, bus.data(), MXB_MEM_HOST, nullptr);
    if (rc != MXB_OK) return rc;
    for (int t = 0; t < n_frames; ++t)
        for (int c = 0; c < channels; ++c) interleaved_out[(size_t)t * channels + c] = c < 2 ? bus[(size_t)t * 2 + c] : 0.0;
    return MXB_OK;
}

}  // extern "C"
