// Voice patches: a per-voice signal graph -- the body of any reference play() built from the stages of maxib200.h --
// run sample by sample for V voices by one interpreting kernel (K8).
//
// The fused bank kernels (K1 / K2) are hard-wired to oscillator -> envelope -> filter -> delay -> mix, the chains
// BASELINE.json measures. Real patches are graphs: two oscillators summed into one filter, an LFO added to a frequency,
// an envelope multiplying the filter OUTPUT (cpp/commandline/maximilian_examples/15.polysynth/main.cpp:54-70), a trigger
// that changes on any sample (10.Filters/main.cpp:27-36). A patch is a short list of stages over 16 per-voice registers;
// operands are registers, per-voice parameter arrays, scalar constants or per-sample input streams. Every voice runs the
// SAME program, so the interpreter's dispatch is warp-uniform (no divergence): one thread = one voice, registers /
// parameters / stage state live in shared memory for the block ([slot][thread], conflict-free), loaded from and stored
// to SoA arrays in HBM once per block like the fused kernels do.
//
// Each stage body is the reference method restated (same statements, same order, -fmad=false). Stages whose reference
// method designs coefficients from its arguments on every call (lores/hires/bandpass, maxiSVF::setParams, maxiBiquad::set,
// pow/atan in maxiNonlinearity, pow in maxiEnvGen) call libdevice where the reference calls glibc: 1e-9 relative instead
// of bit-identical; everything else is bit-identical.
#include <math.h>

#include <new>

#include "patch_impl.cuh"
#include "exchange.cuh"

using namespace mxb;

namespace {

// state / register file of one thread in shared memory: slot-major, thread-minor
struct Lane {
    double* base; int stride;
    __device__ __forceinline__ double& operator[](int slot) const { return base[slot * stride]; }
};

__global__ void __launch_bounds__(kPatchThreads) patch_kernel(const PatchArgs a) {
    extern __shared__ double psm[];
    __shared__ mxb_stage s_stage[kMaxStages];
    __shared__ int s_sbase[kMaxStages], s_ring[kMaxStages];
    __shared__ double s_const[kMaxConsts];
    for (int i = threadIdx.x; i < a.n_stages; i += blockDim.x) { s_stage[i] = a.stages[i]; s_sbase[i] = a.state_base[i]; s_ring[i] = a.ring_of[i]; }
    for (int i = threadIdx.x; i < kMaxConsts; i += blockDim.x) s_const[i] = a.consts[i];
    const int tid = threadIdx.x, lane = tid & 31;
    const long long v = (long long)blockIdx.x * blockDim.x + tid;
    const bool live = v < a.V;
    const size_t V = (size_t)a.V, vv = live ? (size_t)v : 0;
    const Lane reg{psm + tid, kPatchThreads};
    const Lane par{psm + (size_t)kMaxRegs * kPatchThreads + tid, kPatchThreads};
    const Lane st{psm + (size_t)(kMaxRegs + a.n_params) * kPatchThreads + tid, kPatchThreads};
    for (int i = 0; i < kMaxRegs; ++i) reg[i] = 0.0;
    for (int i = 0; i < a.n_params; ++i) par[i] = a.params[(size_t)i * V + vv];
    for (int i = 0; i < a.n_state; ++i) st[i] = a.state[(size_t)i * V + vv];
    __syncthreads();
    const int gwarp = (int)(((long long)blockIdx.x * blockDim.x + tid) >> 5);
    const double sr = a.sr;

    for (int t = 0; t < a.n_frames; ++t) {
        double ml = 0.0, mr = 0.0;
        for (int i = 0; i < kMaxRegs; ++i) reg[i] = 0.0;          // registers read 0 until a stage of this sample writes them
        auto fetch = [&](const int src) -> double {
            if (src < 0) return 0.0;
            const int k = src & 0xff;
            switch (src >> 8) {
                case 0: return reg[k];
                case 1: return par[k];
                case 2: return s_const[k];
                default: return live ? patch_input(a, k, (size_t)t, vv, V) : 0.0;
            }
        };
        for (int si = 0; si < a.n_stages; ++si) {
            const mxb_stage& g = s_stage[si];
            const int sb = s_sbase[si];
            double y = 0.0;
            switch (g.op) {
                case MXB_OP_OSC: {
                    double phase = st[sb], oout = st[sb + 1];
                    y = stage_osc(g.kind, phase, oout, fetch(g.src[0]), fetch(g.src[1]), fetch(g.src[2]), sr, a.sine, a.transition, a.sine_before);
                    st[sb] = phase; st[sb + 1] = oout;
                    break;
                }
                case MXB_OP_ENV_ADSR:
                case MXB_OP_ENV_AR: {
                    EnvRegs e;
                    e.amp = st[sb]; e.output = st[sb + 1]; e.holdcount = (int)st[sb + 2]; env_unpack(e, (int)st[sb + 3]);
                    const double in = fetch(g.src[0]);
                    const bool trig = (int)fetch(g.src[1]) == 1;
                    if (g.op == MXB_OP_ENV_ADSR) {
                        e.att = fetch(g.src[2]); e.dec = fetch(g.src[3]); e.sus = fetch(g.src[4]); e.rel = fetch(g.src[5]);
                        e.holdtime = (int)(long long)fetch(g.src[6]);
                        y = env_tick(e, in, trig);
                    } else {
                        e.att = fetch(g.src[2]); e.rel = fetch(g.src[3]); e.holdtime = (int)(long long)fetch(g.src[4]);
                        e.dec = 0.0; e.sus = 0.0;
                        y = env_ar_tick(e, in, trig);
                    }
                    st[sb] = e.amp; st[sb + 1] = e.output; st[sb + 2] = (double)e.holdcount; st[sb + 3] = (double)env_pack(e);
                    break;
                }
                case MXB_OP_ENVGEN: {
                    // slots: 0 envval, 1 phase, 2 state, 3 nxcHappened, 4 counter, 5 currentlevel, 6/7 trigDetector, 8/9 holdDetector,
                    // 10/11 retriggerDetector (previousValue, firstTrigger)
                    EgRegs q;
                    q.envval = st[sb]; q.phase = (int)st[sb + 1]; q.state = (int)st[sb + 2]; q.nxc = st[sb + 3] != 0.0;
                    q.counter = (long long)st[sb + 4]; q.currentlevel = st[sb + 5];
                    q.tp = st[sb + 6]; q.tf = st[sb + 7]; q.hp = st[sb + 8]; q.hf = st[sb + 9]; q.rp = st[sb + 10]; q.rf = st[sb + 11];
                    y = envgen_tick(q, fetch(g.src[0]), a.eg, a.eg_n, a.eg_loop, a.eg_retrigger);
                    st[sb] = q.envval; st[sb + 1] = (double)q.phase; st[sb + 2] = (double)q.state; st[sb + 3] = q.nxc ? 1.0 : 0.0;
                    st[sb + 4] = (double)q.counter; st[sb + 5] = q.currentlevel;
                    st[sb + 6] = q.tp; st[sb + 7] = q.tf; st[sb + 8] = q.hp; st[sb + 9] = q.hf; st[sb + 10] = q.rp; st[sb + 11] = q.rf;
                    break;
                }
                case MXB_OP_FILTER: {
                    const double in = fetch(g.src[0]);
                    if (g.kind == MXB_FILT_LORES || g.kind == MXB_FILT_HIRES) {
                        FiltRegs f; f.s0 = st[sb]; f.s1 = st[sb + 1];
                        filt_design<FILT_T_LORES>(f, fetch(g.src[1]), fetch(g.src[2]), sr);
                        y = g.kind == MXB_FILT_LORES ? filt_tick<FILT_T_LORES>(f, in, nullptr) : filt_tick<FILT_T_HIRES>(f, in, nullptr);
                        st[sb] = f.s0; st[sb + 1] = f.s1;
                    } else if (g.kind == MXB_FILT_LOPASS || g.kind == MXB_FILT_HIPASS) {
                        double z0 = st[sb];
                        y = onepole_tick(g.kind, z0, in, fetch(g.src[1]));
                        st[sb] = z0;
                    } else {
                        double c0, c1, c2, z0 = st[sb], z1 = st[sb + 1];
                        design_bandpass(fetch(g.src[1]), fetch(g.src[2]), sr, c0, c1, c2);
                        y = bandpass_tick(z0, z1, in, c0, c1, c2);
                        st[sb] = z0; st[sb + 1] = z1;
                    }
                    break;
                }
                case MXB_OP_SVF: {
                    FiltRegs f; f.s0 = st[sb]; f.s1 = st[sb + 1]; f.s2 = st[sb + 2];
                    filt_design<FILT_T_SVF>(f, fetch(g.src[1]), fetch(g.src[2]), sr);
                    const double mixw[4] = {fetch(g.src[3]), fetch(g.src[4]), fetch(g.src[5]), fetch(g.src[6])};
                    y = filt_tick<FILT_T_SVF>(f, fetch(g.src[0]), mixw);
                    st[sb] = f.s0; st[sb + 1] = f.s1; st[sb + 2] = f.s2;
                    break;
                }
                case MXB_OP_BIQUAD: {
                    FiltRegs f; f.s0 = st[sb]; f.s1 = st[sb + 1];
                    double cf[5];
                    design_biquad_one(g.kind, fetch(g.src[1]), fetch(g.src[2]), fetch(g.src[3]), sr, cf);
                    f.c0 = cf[0]; f.c1 = cf[1]; f.c2 = cf[2]; f.c3 = cf[3]; f.c4 = cf[4];
                    y = filt_tick<FILT_T_BIQUAD>(f, fetch(g.src[0]), nullptr);
                    st[sb] = f.s0; st[sb + 1] = f.s1;
                    break;
                }
                case MXB_OP_DCBLOCK: {
                    double xm1 = st[sb], ym1 = st[sb + 1];
                    y = dcblock_tick(xm1, ym1, fetch(g.src[0]), fetch(g.src[1]));
                    st[sb] = xm1; st[sb + 1] = ym1;
                    break;
                }
                case MXB_OP_NONLIN: y = nonlin_eval(g.kind, fetch(g.src[0]), fetch(g.src[1]), fetch(g.src[2])); break;
                case MXB_OP_DELAY:
                case MXB_OP_FLANGER: {
                    double* ring = a.rings + (size_t)s_ring[si] * (size_t)a.taps * V + vv;      // slot r of this voice at ring[r * V]
                    const double in = fetch(g.src[0]);
                    int ph = (int)st[sb];
                    if (g.op == MXB_OP_DELAY) {
                        const int size = (int)fetch(g.src[1]);
                        const double fb = fetch(g.src[2]);
                        const bool fp = g.kind == MXB_DELAY_FROM_POSITION;
                        y = delay_tick(fp, ring, V, a.taps, live, ph, in, size, fb, fp ? (int)fetch(g.src[3]) : 0);
                    } else {
                        double lph = st[sb + 1], lout = st[sb + 2];
                        const unsigned int delay = (unsigned int)fetch(g.src[1]);
                        const double fb = fetch(g.src[2]);
                        y = flanger_tick(ring, V, a.taps, live, ph, lph, lout, in, delay, fb, fetch(g.src[3]), fetch(g.src[4]), sr);
                        st[sb + 1] = lph; st[sb + 2] = lout;
                    }
                    st[sb] = (double)ph;
                    break;
                }
                case MXB_OP_CHORUS: {
                    double* ring = a.rings + (size_t)s_ring[si] * (size_t)a.taps * V + vv;      // this voice's two lines: ring index s_ring[si], + 1
                    int ph1 = (int)st[sb], ph2 = (int)st[sb + 1];
                    FiltRegs lp; lp.s0 = st[sb + 2]; lp.s1 = st[sb + 3];
                    filt_design<FILT_T_LORES>(lp, fetch(g.src[3]), 1.0, sr);
                    y = chorus_tick(ring, ring + (size_t)a.taps * V, V, a.taps, live, ph1, ph2, lp, fetch(g.src[0]), (unsigned int)fetch(g.src[1]), fetch(g.src[2]),
                                    fetch(g.src[4]), fetch(g.src[5]));
                    st[sb] = (double)ph1; st[sb + 1] = (double)ph2; st[sb + 2] = lp.s0; st[sb + 3] = lp.s1;
                    break;
                }
                case MXB_OP_ADD: y = fetch(g.src[0]) + fetch(g.src[1]); break;
                case MXB_OP_SUB: y = fetch(g.src[0]) - fetch(g.src[1]); break;
                case MXB_OP_MUL: y = fetch(g.src[0]) * fetch(g.src[1]); break;
                case MXB_OP_DIV: y = fetch(g.src[0]) / fetch(g.src[1]); break;
                case MXB_OP_MIX_STEREO: {        // maxiMix::stereo, src/maximilian.cpp:503-509, accumulated over the stages of this sample
                    const double in = fetch(g.src[0]), x = fetch(g.src[1]);
                    if (live) mix_stereo_acc(ml, mr, in, x);
                    break;
                }
                case MXB_OP_OUT:
                    if (live && a.out) a.out[(size_t)t * V + vv] = fetch(g.src[0]);
                    break;
                default: break;
            }
            if (g.dst >= 0) reg[g.dst] = y;
        }
        if (a.partials) mix_warp_store(ml, mr, lane, a.partials, (size_t)t, (size_t)a.W, (size_t)gwarp);      // the second kernel adds the warps in order
    }
    if (live) for (int i = 0; i < a.n_state; ++i) a.state[(size_t)i * V + (size_t)v] = st[i];
}

__global__ void patch_mix_reduce_kernel(const double* __restrict__ partials, double* __restrict__ mix, int W) {
    // one CTA of 128 threads per (frame, channel) row: strided ascending sums, fixed tree -- the order never changes between runs
    __shared__ double sm[4];
    const double* p = partials + (size_t)blockIdx.x * (size_t)W;
    double s = 0.0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) s += p[w];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) mix[blockIdx.x] = (sm[0] + sm[1]) + (sm[2] + sm[3]);
}

// The same reduction with the cross-GPU exchange of the bus fused in (K6; protocol and buffers: exchange.cuh / exchange.cu, the bank's
// twin of this kernel is mix_reduce_exchange_kernel in bank.cu): every CTA pushes its row sum into this rank's lane of EVERY rank's buffer
// (posted NVLink stores), the last CTA to finish publishes one flag per peer, waits -- bounded -- for the peers' flags in its own buffer
// and adds the world's buses in rank order: the same bits on every rank, every run.
__global__ void patch_mix_reduce_exchange_kernel(const double* __restrict__ partials, double* __restrict__ mix, int rows, int W, const ExchDev x) {
    __shared__ double sm[4];
    __shared__ bool last;
    const double* p = partials + (size_t)blockIdx.x * (size_t)W;
    double s = 0.0;
    for (int w = threadIdx.x; w < W; w += blockDim.x) s += p[w];
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    const double rowsum = (sm[0] + sm[1]) + (sm[2] + sm[3]);
    if (threadIdx.x < x.world) x.dst_payload[threadIdx.x][blockIdx.x] = rowsum;      // thread r -> rank r's buffer
    __threadfence_system();                           // this CTA's peer stores are ordered before its ticket
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(x.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x < x.world) {
        __threadfence_system();                       // cumulative: the other CTAs' stores were ordered before their tickets
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(x.dst_flag[threadIdx.x]), "l"(x.seq1) : "memory");
        const unsigned long long* flag = x.src_flags + (size_t)threadIdx.x * (kExchFlagBytes / sizeof(unsigned long long));
        unsigned long long t0, t1, v;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        unsigned spins = 0;
        for (;;) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
            if (v >= x.seq1) break;
            if ((++spins & 1023u) == 0) {
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                if (t1 - t0 > x.timeout_ns) { atomicOr(x.status, 1u << threadIdx.x); break; }      // a silent peer costs a time-out, not a hung box
            }
        }
    }
    if (threadIdx.x == 0) *x.ticket = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < rows; i += blockDim.x) {
        double t = 0.0;
        for (int r = 0; r < x.world; ++r) t += __ldcg(x.src_payload + (size_t)r * (size_t)x.stride + i);
        mix[i] = t;
    }
}

int state_slots(const mxb_stage& g) { return patch_state_slots(g.op); }

// MXB_PATCH_MODE=interpret | fused overrides the default (fused) for patches created afterwards: an A/B and debugging hook
int patch_default_mode() {
    const char* e = getenv("MXB_PATCH_MODE");
    if (e && strcmp(e, "interpret") == 0) return MXB_PATCH_INTERPRET;
    return MXB_PATCH_FUSED;
}

bool src_ok(int s, const mxb_patch_desc* d) {
    if (s == MXB_NONE) return true;
    if (s < 0) return false;
    const int k = s & 0xff;
    switch (s >> 8) {
        case 0: return k < kMaxRegs;
        case 1: return k < d->n_params;
        case 2: return k < d->n_consts;
        case 3: return k < d->n_inputs;
        default: return false;
    }
}

}  // namespace

extern "C" {

int32_t mxb_ctx_set_tables(mxb_ctx* ctx, const double* sine514, const double* transition1001, double sine_before) {
    MXB_REQUIRE(ctx && sine514 && transition1001, MXB_ERR_INVALID, "mxb_ctx_set_tables: NULL argument");
    DeviceGuard g(ctx->device);
    if (!ctx->d_sine) { int rc = dev_alloc(&ctx->d_sine, 514 + 1001); if (rc != MXB_OK) return rc; }
    MXB_CUDA(cudaMemcpy(ctx->d_sine, sine514, sizeof(double) * 514, cudaMemcpyHostToDevice));
    MXB_CUDA(cudaMemcpy(ctx->d_sine + 514, transition1001, sizeof(double) * 1001, cudaMemcpyHostToDevice));
    ctx->sine_before = sine_before;
    return MXB_OK;
}

int32_t mxb_patch_create(mxb_ctx* ctx, const mxb_patch_desc* d, mxb_patch** out) {
    MXB_REQUIRE(ctx && d && out && d->stages, MXB_ERR_INVALID, "mxb_patch_create: NULL argument");
    *out = nullptr;
    MXB_REQUIRE(d->voices > 0 && d->max_frames > 0, MXB_ERR_INVALID, "mxb_patch_create: voices %d max_frames %d", d->voices, d->max_frames);
    MXB_REQUIRE(d->n_stages > 0 && d->n_stages <= kMaxStages, MXB_ERR_INVALID, "mxb_patch_create: n_stages %d (1..%d)", d->n_stages, kMaxStages);
    MXB_REQUIRE(d->n_params >= 0 && d->n_params <= kMaxParams && d->n_consts >= 0 && d->n_consts <= kMaxConsts && d->n_inputs >= 0 && d->n_inputs <= kMaxInputs,
                MXB_ERR_INVALID, "mxb_patch_create: n_params %d n_consts %d n_inputs %d", d->n_params, d->n_consts, d->n_inputs);
    MXB_REQUIRE(d->n_consts == 0 || d->consts, MXB_ERR_INVALID, "mxb_patch_create: consts is NULL");
    MXB_REQUIRE(d->delay_taps >= 0, MXB_ERR_INVALID, "mxb_patch_create: delay_taps %d", d->delay_taps);
    MXB_REQUIRE(d->eg_stages >= 0 && d->eg_stages <= kMaxEg, MXB_ERR_INVALID, "mxb_patch_create: eg_stages %d (0..%d)", d->eg_stages, kMaxEg);
    for (int i = 0; d->input_types && i < d->n_inputs; ++i)
        MXB_REQUIRE(d->input_types[i] >= MXB_IN_F64 && d->input_types[i] <= MXB_IN_BITS, MXB_ERR_INVALID, "mxb_patch_create: input_types[%d] = %d", i, d->input_types[i]);
    int n_state = 0, n_rings = 0;
    bool tables = false, eg = false;
    for (int i = 0; i < d->n_stages; ++i) {
        const mxb_stage& g = d->stages[i];
        MXB_REQUIRE(g.op >= MXB_OP_OSC && g.op <= MXB_OP_CHORUS, MXB_ERR_INVALID, "mxb_patch_create: stage %d: unknown op %d", i, g.op);
        MXB_REQUIRE(g.dst == MXB_NONE || (g.dst >= 0 && g.dst < kMaxRegs), MXB_ERR_INVALID, "mxb_patch_create: stage %d: dst %d", i, g.dst);
        for (int k = 0; k < MXB_STAGE_SRCS; ++k) MXB_REQUIRE(src_ok(g.src[k], d), MXB_ERR_INVALID, "mxb_patch_create: stage %d: operand %d = 0x%x", i, k, g.src[k]);
        if (g.op == MXB_OP_OSC) {
            MXB_REQUIRE(g.kind >= MXB_OSC_SINEWAVE && g.kind <= MXB_OSC_SAWN, MXB_ERR_INVALID, "mxb_patch_create: stage %d: osc kind %d", i, g.kind);
            tables = tables || g.kind >= MXB_OSC_SINEBUF;
        }
        if (g.op == MXB_OP_FILTER) MXB_REQUIRE(g.kind == MXB_FILT_LORES || g.kind == MXB_FILT_HIRES || g.kind == MXB_FILT_LOPASS || g.kind == MXB_FILT_HIPASS || g.kind == MXB_FILT_BANDPASS,
                                               MXB_ERR_INVALID, "mxb_patch_create: stage %d: filter kind %d", i, g.kind);
        if (g.op == MXB_OP_BIQUAD) MXB_REQUIRE(g.kind >= MXB_BQ_LOWPASS && g.kind <= MXB_BQ_HIGHSHELF, MXB_ERR_INVALID, "mxb_patch_create: stage %d: biquad type %d", i, g.kind);
        if (g.op == MXB_OP_NONLIN) MXB_REQUIRE(g.kind >= MXB_NL_ATANDIST && g.kind <= MXB_NL_FASTATAN, MXB_ERR_INVALID, "mxb_patch_create: stage %d: nonlinearity %d", i, g.kind);
        if (patch_stage_rings(g.op)) { MXB_REQUIRE(d->delay_taps > 0, MXB_ERR_INVALID, "mxb_patch_create: stage %d needs delay_taps > 0", i); n_rings += patch_stage_rings(g.op); }
        if (g.op == MXB_OP_ENVGEN) eg = true;
        n_state += state_slots(g);
    }
    if (tables) MXB_REQUIRE(ctx->d_sine, MXB_ERR_STATE, "mxb_patch_create: sinebuf / sinebuf4 / sawn need the reference's tables: call mxb_ctx_set_tables first");
    if (eg) MXB_REQUIRE(d->eg_stages > 0 && d->eg_levels && d->eg_times && d->eg_curves, MXB_ERR_INVALID, "mxb_patch_create: an ENVGEN stage needs eg_levels / eg_times / eg_curves");
    const size_t per_thread = (size_t)(kMaxRegs + d->n_params + n_state);
    MXB_REQUIRE(per_thread * 8 * kPatchThreads <= 200 * 1024, MXB_ERR_UNSUPPORTED, "mxb_patch_create: %zu doubles of registers, parameters and state per voice do not fit shared memory", per_thread);
    DeviceGuard g(ctx->device);
    mxb_patch* p = new (std::nothrow) mxb_patch();
    MXB_REQUIRE(p, MXB_ERR_ALLOC, "mxb_patch_create: out of host memory");
    p->ctx = ctx; p->V = d->voices; p->n_stages = d->n_stages; p->n_params = d->n_params; p->n_consts = d->n_consts; p->n_inputs = d->n_inputs;
    p->max_frames = d->max_frames; p->taps = d->delay_taps; p->n_state = n_state; p->n_rings = n_rings;
    p->stages.assign(d->stages, d->stages + d->n_stages);
    if (d->n_consts) p->consts.assign(d->consts, d->consts + d->n_consts);
    p->mode = patch_default_mode(); p->fused = nullptr; p->ex = nullptr;
    for (int i = 0; i < d->n_inputs; ++i) p->in_type[i] = d->input_types ? d->input_types[i] : MXB_IN_F64;
    int sb = 0, ri = 0;
    for (int i = 0; i < d->n_stages; ++i) {
        p->state_base.push_back(sb); sb += state_slots(p->stages[i]);
        const int nr = patch_stage_rings(p->stages[i].op);
        p->ring_of.push_back(nr ? ri : -1);
        ri += nr;
    }
    // maxiEnvGen::setup + setupSegmentTime, src/maximilian.h:2371-2402, 2524-2538: segment table from levels / times / curves
    p->eg_n = d->eg_stages; p->eg_loop = d->eg_loop; p->eg_retrigger = d->eg_retrigger;
    {
        double accumulatedTime = 0;
        const double srd = (double)(size_t)ctx->sample_rate;
        for (int i = 0; i < d->eg_stages; ++i) {
            EgStage& s = p->eg[i];
            s.startlevel = d->eg_levels[i]; s.endlevel = d->eg_levels[i + 1]; s.curve = d->eg_curves[i]; s.pad = 0;
            const double stageTime = d->eg_times[i];
            if (stageTime == MXB_ENVGEN_HOLD) { s.length = 0; s.hold = 1; s.gradient = 0; }
            else {
                const double len = ((stageTime / 1000.0) * srd) + accumulatedTime;
                s.length = (long long)(size_t)floor(len);
                accumulatedTime = len - s.length;
                s.gradient = 1.0 / s.length;
                s.hold = 0;
            }
        }
    }
    const size_t V = (size_t)p->V;
    int rc = MXB_OK;
#define TRY(x) do { rc = (x); if (rc != MXB_OK) { mxb_patch_destroy(p); return rc; } } while (0)
    TRY(dev_alloc(&p->d_stages, (size_t)d->n_stages)); TRY(dev_alloc(&p->d_state_base, (size_t)d->n_stages)); TRY(dev_alloc(&p->d_ring_of, (size_t)d->n_stages));
    TRY(dev_alloc(&p->d_consts, (size_t)kMaxConsts));
    TRY(dev_alloc(&p->params, (size_t)(d->n_params ? d->n_params : 1) * V));
    TRY(dev_alloc(&p->state, (size_t)(n_state ? n_state : 1) * V));
    if (n_rings) TRY(dev_alloc(&p->rings, (size_t)n_rings * (size_t)p->taps * V));
    TRY(dev_alloc(&p->mix_dev, (size_t)p->max_frames * 2));
    cudaError_t e = cudaMemcpy(p->d_stages, p->stages.data(), sizeof(mxb_stage) * d->n_stages, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_state_base, p->state_base.data(), sizeof(int) * d->n_stages, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(p->d_ring_of, p->ring_of.data(), sizeof(int) * d->n_stages, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && d->n_consts) e = cudaMemcpy(p->d_consts, d->consts, sizeof(double) * d->n_consts, cudaMemcpyHostToDevice);
    // initial state that is not zero in the reference: maxiTrigger (previousValue = 1, firstTrigger = 1, src/maximilian.h:583-584)
    for (int i = 0; i < d->n_stages && e == cudaSuccess; ++i) {
        if (p->stages[i].op != MXB_OP_ENVGEN) continue;
        std::vector<double> ones(V, 1.0);
        for (int k = 6; k < 12 && e == cudaSuccess; ++k)
            e = cudaMemcpy(p->state + (size_t)(p->state_base[i] + k) * V, ones.data(), sizeof(double) * V, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) { set_error("mxb_patch_create: %s", cudaGetErrorString(e)); mxb_patch_destroy(p); return MXB_ERR_CUDA; }
#undef TRY
    *out = p;
    return MXB_OK;
}

int32_t mxb_patch_destroy(mxb_patch* p) {
    if (!p) return MXB_OK;
    DeviceGuard g(p->ctx->device);
    cudaDeviceSynchronize();
    patch_fused_free(p);
    cudaFree(p->d_stages); cudaFree(p->d_state_base); cudaFree(p->d_ring_of); cudaFree(p->d_consts);
    cudaFree(p->params); cudaFree(p->state); cudaFree(p->rings); cudaFree(p->out_stage); cudaFree(p->partials); cudaFree(p->mix_dev);
    for (auto q : p->in_stage) cudaFree(q);
    delete p;
    return MXB_OK;
}

int32_t mxb_patch_set_param(mxb_patch* p, int32_t j, const double* values, int32_t mem) {
    MXB_REQUIRE(p && values, MXB_ERR_INVALID, "mxb_patch_set_param: NULL argument");
    MXB_REQUIRE(j >= 0 && j < p->n_params, MXB_ERR_INVALID, "mxb_patch_set_param: parameter %d of %d", j, p->n_params);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_patch_set_param: mem %d", mem);
    DeviceGuard g(p->ctx->device);
    MXB_CUDA(cudaMemcpy(p->params + (size_t)j * p->V, values, sizeof(double) * (size_t)p->V, mem == MXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
    return MXB_OK;
}

static int32_t patch_state_ptr(mxb_patch* p, int32_t stage, int32_t slot, double** ptr, const char* who) {
    MXB_REQUIRE(p, MXB_ERR_INVALID, "%s: NULL patch", who);
    MXB_REQUIRE(stage >= 0 && stage < p->n_stages, MXB_ERR_INVALID, "%s: stage %d of %d", who, stage, p->n_stages);
    MXB_REQUIRE(slot >= 0 && slot < state_slots(p->stages[stage]), MXB_ERR_INVALID, "%s: stage %d has %d state slots, not %d", who, stage, state_slots(p->stages[stage]), slot);
    *ptr = p->state + (size_t)(p->state_base[stage] + slot) * p->V;
    return MXB_OK;
}

int32_t mxb_patch_set_state(mxb_patch* p, int32_t stage, int32_t slot, const double* values, int32_t mem) {
    double* dst = nullptr;
    int rc = patch_state_ptr(p, stage, slot, &dst, "mxb_patch_set_state");
    if (rc != MXB_OK) return rc;
    MXB_REQUIRE(values && (mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE), MXB_ERR_INVALID, "mxb_patch_set_state: bad argument");
    DeviceGuard g(p->ctx->device);
    MXB_CUDA(cudaMemcpy(dst, values, sizeof(double) * (size_t)p->V, mem == MXB_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
    return MXB_OK;
}

int32_t mxb_patch_get_state(mxb_patch* p, int32_t stage, int32_t slot, double* values, int32_t mem) {
    double* src = nullptr;
    int rc = patch_state_ptr(p, stage, slot, &src, "mxb_patch_get_state");
    if (rc != MXB_OK) return rc;
    MXB_REQUIRE(values && (mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE), MXB_ERR_INVALID, "mxb_patch_get_state: bad argument");
    DeviceGuard g(p->ctx->device);
    MXB_CUDA(cudaMemcpy(values, src, sizeof(double) * (size_t)p->V, mem == MXB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
    return MXB_OK;
}

int32_t mxb_patch_get_ring(mxb_patch* p, int32_t stage, int32_t voice, double* dst, int32_t n, int32_t mem) {
    MXB_REQUIRE(p && dst, MXB_ERR_INVALID, "mxb_patch_get_ring: NULL argument");
    MXB_REQUIRE(stage >= 0 && stage < p->n_stages && p->ring_of[stage] >= 0, MXB_ERR_INVALID, "mxb_patch_get_ring: stage %d has no delay line", stage);
    MXB_REQUIRE(voice >= 0 && voice < p->V && n >= 0 && n <= p->taps * patch_stage_rings(p->stages[stage].op), MXB_ERR_INVALID, "mxb_patch_get_ring: voice %d n %d", voice, n);
    DeviceGuard g(p->ctx->device);
    const double* src = p->rings + (size_t)p->ring_of[stage] * (size_t)p->taps * (size_t)p->V + (size_t)voice;
    if (n) MXB_CUDA(cudaMemcpy2D(dst, sizeof(double), src, sizeof(double) * (size_t)p->V, sizeof(double), (size_t)n,
                                 mem == MXB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
    return MXB_OK;
}

int64_t mxb_patch_launch_count(const mxb_patch* p) { return p ? p->launches : 0; }

int32_t mxb_patch_set_mode(mxb_patch* p, int32_t mode) {
    MXB_REQUIRE(p, MXB_ERR_INVALID, "mxb_patch_set_mode: NULL patch");
    MXB_REQUIRE(mode == MXB_PATCH_INTERPRET || mode == MXB_PATCH_FUSED, MXB_ERR_INVALID, "mxb_patch_set_mode: mode %d", mode);
    if (mode == MXB_PATCH_FUSED) {          // compile now: a patch that cannot be compiled says so here, not in the audio loop
        DeviceGuard g(p->ctx->device);
        const int rc = patch_fused_load(p);
        if (rc != MXB_OK) return rc;
    }
    p->mode = mode;
    return MXB_OK;
}

int32_t mxb_patch_get_mode(const mxb_patch* p) { return p ? p->mode : MXB_ERR_INVALID; }

int32_t mxb_patch_codegen(const mxb_patch_desc* d, char* buf, int64_t cap, int64_t* needed, int32_t compile) {
    MXB_REQUIRE(d && d->stages && needed, MXB_ERR_INVALID, "mxb_patch_codegen: NULL argument");
    MXB_REQUIRE(d->n_stages > 0 && d->n_stages <= kMaxStages && d->n_params >= 0 && d->n_params <= kMaxParams && d->n_consts >= 0 && d->n_consts <= kMaxConsts &&
                d->n_inputs >= 0 && d->n_inputs <= kMaxInputs && (d->n_consts == 0 || d->consts), MXB_ERR_INVALID, "mxb_patch_codegen: bad descriptor");
    for (int i = 0; i < d->n_stages; ++i) {
        const mxb_stage& g = d->stages[i];
        MXB_REQUIRE(g.op >= MXB_OP_OSC && g.op <= MXB_OP_CHORUS && (g.dst == MXB_NONE || (g.dst >= 0 && g.dst < kMaxRegs)), MXB_ERR_INVALID, "mxb_patch_codegen: stage %d", i);
        for (int k = 0; k < MXB_STAGE_SRCS; ++k) MXB_REQUIRE(src_ok(g.src[k], d), MXB_ERR_INVALID, "mxb_patch_codegen: stage %d: operand %d = 0x%x", i, k, g.src[k]);
    }
    int types[kMaxInputs] = {};
    for (int i = 0; d->input_types && i < d->n_inputs; ++i) {
        MXB_REQUIRE(d->input_types[i] >= MXB_IN_F64 && d->input_types[i] <= MXB_IN_BITS, MXB_ERR_INVALID, "mxb_patch_codegen: input_types[%d] = %d", i, d->input_types[i]);
        types[i] = d->input_types[i];
    }
    const std::string src = patch_generate_source(d->stages, d->n_stages, d->n_params, d->consts, d->n_consts, d->n_inputs, types);
    *needed = (int64_t)src.size() + 1;
    if (buf && cap > 0) { const size_t n = src.size() + 1 <= (size_t)cap ? src.size() : (size_t)cap - 1; memcpy(buf, src.data(), n); buf[n] = 0; }
    if (compile) {
        std::vector<char> cubin;
        const int rc = patch_compile(src, cubin);
        if (rc != MXB_OK) return rc;
        MXB_REQUIRE(!cubin.empty(), MXB_ERR_STATE, "mxb_patch_codegen: the compiler returned no code");
    }
    return MXB_OK;
}

int32_t mxb_patch_process(mxb_patch* p, int32_t n_frames, const void* const* inputs, double* out, double* mix, int32_t mem, void* stream_) {
    NvtxRange nvtx_("mxb_patch_process");
    MXB_REQUIRE(p, MXB_ERR_INVALID, "mxb_patch_process: NULL patch");
    MXB_REQUIRE(n_frames >= 0 && n_frames <= p->max_frames, MXB_ERR_INVALID, "mxb_patch_process: n_frames %d (max_frames %d)", n_frames, p->max_frames);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_patch_process: mem %d", mem);
    MXB_REQUIRE(p->n_inputs == 0 || inputs, MXB_ERR_INVALID, "mxb_patch_process: the patch reads %d input streams, inputs is NULL", p->n_inputs);
    for (int i = 0; i < p->n_inputs; ++i) MXB_REQUIRE(inputs[i], MXB_ERR_INVALID, "mxb_patch_process: input stream %d is NULL", i);
    if (mix && p->ex) {     // before any kernel runs: a refused call leaves every stage's state untouched
        MXB_REQUIRE(p->ex->connected, MXB_ERR_STATE, "mxb_patch_process: the attached exchange is not connected to its peers");
        MXB_REQUIRE(n_frames * 2 <= p->ex->max_doubles, MXB_ERR_INVALID, "mxb_patch_process: exchange holds %d values, the bus needs %d", p->ex->max_doubles, n_frames * 2);
    }
    if (n_frames == 0) return MXB_OK;
    DeviceGuard g(p->ctx->device);
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t V = (size_t)p->V, nb = sizeof(double) * (size_t)n_frames * V;
    PatchArgs a;
    memset(&a, 0, sizeof(a));
    double* d_out = out; double* d_mix = mix;
    for (int i = 0; i < p->n_inputs; ++i) {
        a.inputs[i] = inputs[i];
        a.in_kind[i] = (unsigned char)p->in_type[i];
        if (mem == MXB_MEM_HOST) {
            const size_t nb = patch_input_bytes(p->in_type[i], (size_t)n_frames, V);
            if (nb > p->in_stage_len[i]) {
                MXB_CUDA(cudaStreamSynchronize(s));
                cudaFree(p->in_stage[i]); p->in_stage[i] = nullptr; p->in_stage_len[i] = 0;
                cudaError_t e = cudaMalloc(&p->in_stage[i], nb);
                if (e != cudaSuccess) { set_error("mxb_patch_process: staging cudaMalloc(%zu): %s", nb, cudaGetErrorString(e)); return MXB_ERR_ALLOC; }
                p->in_stage_len[i] = nb;
            }
            MXB_CUDA(cudaMemcpyAsync(p->in_stage[i], inputs[i], nb, cudaMemcpyHostToDevice, s));
            a.inputs[i] = p->in_stage[i];
        }
    }
    if (mem == MXB_MEM_HOST) {
        if (out) {
            if (nb > p->out_stage_len) {
                MXB_CUDA(cudaStreamSynchronize(s));
                cudaFree(p->out_stage); p->out_stage = nullptr; p->out_stage_len = 0;
                cudaError_t e = cudaMalloc((void**)&p->out_stage, nb);
                if (e != cudaSuccess) { set_error("mxb_patch_process: staging cudaMalloc(%zu): %s", nb, cudaGetErrorString(e)); return MXB_ERR_ALLOC; }
                p->out_stage_len = nb;
            }
            d_out = p->out_stage;
        }
        if (mix) d_mix = p->mix_dev;
    }
    const int grid = (int)((V + kPatchThreads - 1) / kPatchThreads);
    const int W = grid * (kPatchThreads / 32);
    if (mix && !p->partials) { int rc = dev_alloc(&p->partials, (size_t)p->max_frames * 2 * (size_t)W, false); if (rc != MXB_OK) return rc; }
    a.V = p->V; a.n_frames = n_frames; a.n_stages = p->n_stages; a.n_params = p->n_params; a.n_state = p->n_state; a.W = W; a.taps = p->taps;
    a.sr = (double)(size_t)p->ctx->sample_rate;
    a.stages = p->d_stages; a.state_base = p->d_state_base; a.ring_of = p->d_ring_of; a.consts = p->d_consts;
    a.params = p->params; a.state = p->state; a.out = d_out; a.partials = mix ? p->partials : nullptr; a.rings = p->rings;
    a.sine = p->ctx->d_sine; a.transition = p->ctx->d_sine ? p->ctx->d_sine + 514 : nullptr; a.sine_before = p->ctx->sine_before;
    a.eg_n = p->eg_n; a.eg_loop = p->eg_loop; a.eg_retrigger = p->eg_retrigger;
    for (int i = 0; i < p->eg_n; ++i) a.eg[i] = p->eg[i];
    if (p->mode == MXB_PATCH_FUSED) {
        // the kernel generated for this stage list (patch_fuse.cu): compiled on first use, then launched like any other
        int rc = patch_fused_load(p);
        if (rc == MXB_OK) rc = patch_fused_launch(p, a, grid, s);
        if (rc != MXB_OK) return rc;
    } else {
        const size_t smem = sizeof(double) * (size_t)(kMaxRegs + p->n_params + p->n_state) * kPatchThreads;
        MXB_CUDA(cudaFuncSetAttribute(patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        patch_kernel<<<grid, kPatchThreads, smem, s>>>(a);
        MXB_CUDA(cudaGetLastError());
    }
    p->launches += 1;
    if (mix) {
        if (p->ex) patch_mix_reduce_exchange_kernel<<<n_frames * 2, 128, 0, s>>>(p->partials, d_mix, n_frames * 2, W, exchange_next(p->ex));
        else patch_mix_reduce_kernel<<<n_frames * 2, 128, 0, s>>>(p->partials, d_mix, W);
        MXB_CUDA(cudaGetLastError());
        p->launches += 1;
    }
    if (mem == MXB_MEM_HOST) {
        if (out) MXB_CUDA(cudaMemcpyAsync(out, d_out, nb, cudaMemcpyDeviceToHost, s));
        if (mix) MXB_CUDA(cudaMemcpyAsync(mix, d_mix, sizeof(double) * (size_t)n_frames * 2, cudaMemcpyDeviceToHost, s));
        MXB_CUDA(cudaStreamSynchronize(s));
        if (mix && p->ex) {       // the call is synchronous in this mode: a timed-out exchange is an error of THIS call
            unsigned int m = 0;
            MXB_CUDA(cudaMemcpy(&m, p->ex->status, sizeof(m), cudaMemcpyDeviceToHost));
            MXB_REQUIRE(m == 0, MXB_ERR_STATE, "mxb_patch_process: mix exchange timed out waiting for rank mask 0x%x (bus incomplete)", m);
        }
    }
    return MXB_OK;
}

int32_t mxb_patch_set_exchange(mxb_patch* p, mxb_exchange* ex) {
    MXB_REQUIRE(p, MXB_ERR_INVALID, "mxb_patch_set_exchange: NULL patch");
    if (ex) MXB_REQUIRE(ex->ctx->device == p->ctx->device, MXB_ERR_INVALID, "mxb_patch_set_exchange: exchange and patch live on different devices");
    p->ex = ex;
    return MXB_OK;
}

}  // extern "C"
