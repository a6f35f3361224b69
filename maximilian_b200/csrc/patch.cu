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

#include "bank_kernels.cuh"
#include "filter_design.cuh"

using namespace mxb;

namespace {

constexpr int kPatchThreads = 128;
constexpr int kMaxStages = 64, kMaxParams = 32, kMaxConsts = 64, kMaxInputs = 8, kMaxRegs = 16, kMaxEg = 16;

struct EgStage { double startlevel, endlevel, gradient, curve; long long length; int hold; int pad; };

struct PatchArgs {
    int V, n_frames, n_stages, n_params, n_state, W, taps;
    double sr;
    const mxb_stage* stages;         // device, n_stages
    const int* state_base;           // device, n_stages: first state slot of each stage
    const int* ring_of;              // device, n_stages: ring index of a delay-like stage, else -1
    const double* consts;            // device, kMaxConsts
    const double* params;            // [n_params][V]
    double* state;                   // [n_state][V]
    const double* inputs[kMaxInputs];// [n_frames][V] each
    double* out;                     // [n_frames][V] or NULL
    double* partials;                // [n_frames][2][W] or NULL
    double* rings;                   // [n_rings][taps][V]
    const double* sine;              // sineBuffer[514], src/maximilian.cpp:63
    const double* transition;        // transition[1001], src/maximilian.cpp:67-200
    double sine_before;              // what sinebuf4 reads at sineBuffer[-1] on its wrap sample (out of bounds in the reference)
    int eg_n, eg_loop, eg_retrigger;
    EgStage eg[kMaxEg];
};

// table oscillators, src/maximilian.cpp:237-274, 342-359
__device__ __forceinline__ double osc_table_tick(const int kind, double& phase, double& output, const double frequency, const double sr,
                                                 const double* __restrict__ sine, const double* __restrict__ transition, const double sine_before) {
    if (kind == MXB_OSC_SINEBUF4) {          // :237-264
        phase += 512. / (sr / (frequency));
        if (phase >= 511) phase -= 512;
        const double remainder = phase - floor(phase);
        double a, b, c, d;
        const long long ip = (long long)phase;
        if (phase == 0) { a = sine[512]; b = sine[ip]; c = sine[ip + 1]; d = sine[ip + 2]; }
        else { a = ip - 1 < 0 ? sine_before : sine[ip - 1]; b = sine[ip]; c = sine[ip + 1]; d = sine[ip + 2]; }
        const double a1 = 0.5 * (c - a);
        const double a2 = a - 2.5 * b + 2.0 * c - 0.5 * d;
        const double a3 = 0.5 * (d - a) + 1.5 * (b - c);
        output = ((a3 * remainder + a2) * remainder + a1) * remainder + b;
    } else if (kind == MXB_OSC_SINEBUF) {    // :266-274 (chandiv == 1)
        phase += 512. / (sr / (frequency * 1.0));
        if (phase >= 511) phase -= 512;
        const double remainder = phase - floor(phase);
        const long long ip = (long long)phase;
        output = (1 - remainder) * sine[1 + ip] + remainder * sine[2 + ip];
    } else {                                 // sawn, :342-359
        if (phase >= 0.5) phase -= 1.0;
        phase += (1. / (sr / (frequency)));
        double temp = (8820.22 / frequency) * phase;
        if (temp < -0.5) temp = -0.5;
        if (temp > 0.5) temp = 0.5;
        temp *= 1000.0;
        temp += 500.0;
        const double remainder = temp - floor(temp);
        const long long it = (long long)temp;
        // transition[1 + it] with it == 1000 is one past the table in the reference, multiplied by remainder == 0
        const double t1 = it + 1 <= 1000 ? transition[it + 1] : 0.0;
        output = ((1.0 - remainder) * transition[it] + remainder * t1) - phase;
    }
    return output;
}

// maxiTrigger::onZX, src/maximilian.h:564-585
__device__ __forceinline__ double on_zx(double& previousValue, double& firstTrigger, const double input) {
    double isZX = 0.0;
    if ((previousValue <= 0.0 || firstTrigger != 0.0) && input > 0) isZX = 1.0;
    previousValue = input;
    firstTrigger = 0.0;
    return isZX;
}

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
                default: return live ? a.inputs[k][(size_t)t * V + vv] : 0.0;
            }
        };
        for (int si = 0; si < a.n_stages; ++si) {
            const mxb_stage& g = s_stage[si];
            const int sb = s_sbase[si];
            double y = 0.0;
            switch (g.op) {
                case MXB_OP_OSC: {
                    double phase = st[sb], oout = st[sb + 1];
                    const double f = fetch(g.src[0]);
                    if (g.kind >= MXB_OSC_SINEBUF) y = osc_table_tick(g.kind, phase, oout, f, sr, a.sine, a.transition, a.sine_before);
                    else {
                        const double d0 = fetch(g.src[1]), d1 = fetch(g.src[2]);
                        const double inc = g.kind == MXB_OSC_PHASORBETWEEN ? ((d1 - d0) / (sr / (f))) : (1. / (sr / (f)));
                        y = osc_tick<OSC_T_GENERIC>(phase, oout, inc, d0, g.kind, d1);
                    }
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
                    // maxiEnvGen::play, src/maximilian.h:2276-2357. Slots: 0 envval, 1 phase, 2 state (0 WAITING, 1 TRIGGERED, 2 HOLDING),
                    // 3 nxcHappened, 4 counter, 5 currentlevel (of the current segment: every other segment's are 0), 6/7 trigDetector,
                    // 8/9 holdDetector, 10/11 retriggerDetector (previousValue, firstTrigger)
                    const double trigger = fetch(g.src[0]);
                    double envval = st[sb]; int phase = (int)st[sb + 1], state = (int)st[sb + 2]; bool nxc = st[sb + 3] != 0.0;
                    long long counter = (long long)st[sb + 4]; double currentlevel = st[sb + 5];
                    double tp = st[sb + 6], tf = st[sb + 7], hp = st[sb + 8], hf = st[sb + 9], rp = st[sb + 10], rf = st[sb + 11];
                    auto reset = [&]() { counter = 0; currentlevel = 0; phase = 0; state = 1; };
                    bool run = true;
                    if (state == 0) {
                        if (on_zx(tp, tf, trigger) != 0.0) { if (a.eg_n > 0) { state = 1; nxc = false; } else run = false; }
                        else run = false;
                    }
                    if (run && state == 1) {
                        const EgStage& cs = a.eg[phase < a.eg_n ? phase : 0];
                        if (on_zx(hp, hf, -trigger) != 0.0) nxc = true;
                        if (cs.hold) state = 2;
                        else {
                            double val = pow(currentlevel, cs.curve);
                            val = fmax(fmin(val, 1.0), 0.0);                                   // maxiMap::linlin, src/maximilian.h:801-805
                            envval = ((val - 0.0) / (1.0 - 0.0) * (cs.endlevel - cs.startlevel)) + cs.startlevel;
                            counter++;
                            if (counter == cs.length) { counter = 0; currentlevel = 0; phase++; }
                            else currentlevel += cs.gradient;
                            if (a.eg_retrigger) { if (on_zx(rp, rf, trigger) != 0.0) { nxc = false; reset(); } }
                            run = false;
                        }
                    }
                    if (run && state == 2) {
                        if (on_zx(hp, hf, -trigger) != 0.0) nxc = true;
                        if (nxc) { state = 1; phase++; }
                        if (a.eg_retrigger) { if (on_zx(rp, rf, trigger) != 0.0) { nxc = false; reset(); } }
                    }
                    if (phase == a.eg_n) { reset(); if (!a.eg_loop) state = 0; }
                    y = envval;
                    st[sb] = envval; st[sb + 1] = (double)phase; st[sb + 2] = (double)state; st[sb + 3] = nxc ? 1.0 : 0.0;
                    st[sb + 4] = (double)counter; st[sb + 5] = currentlevel;
                    st[sb + 6] = tp; st[sb + 7] = tf; st[sb + 8] = hp; st[sb + 9] = hf; st[sb + 10] = rp; st[sb + 11] = rf;
                    break;
                }
                case MXB_OP_FILTER: {
                    const double in = fetch(g.src[0]);
                    if (g.kind == MXB_FILT_LORES || g.kind == MXB_FILT_HIRES) {
                        FiltRegs f; f.s0 = st[sb]; f.s1 = st[sb + 1];
                        filt_design<FILT_T_LORES>(f, fetch(g.src[1]), fetch(g.src[2]), sr);
                        y = g.kind == MXB_FILT_LORES ? filt_tick<FILT_T_LORES>(f, in, nullptr) : filt_tick<FILT_T_HIRES>(f, in, nullptr);
                        st[sb] = f.s0; st[sb + 1] = f.s1;
                    } else if (g.kind == MXB_FILT_LOPASS) {      // src/maximilian.cpp:442-446
                        const double c = fetch(g.src[1]);
                        y = st[sb] + c * (in - st[sb]);
                        st[sb] = y;
                    } else if (g.kind == MXB_FILT_HIPASS) {      // :449-453
                        const double c = fetch(g.src[1]);
                        y = in - (st[sb] + c * (in - st[sb]));
                        st[sb] = y;
                    } else {                                      // bandpass, :487-500
                        double c0, c1, c2;
                        design_bandpass(fetch(g.src[1]), fetch(g.src[2]), sr, c0, c1, c2);
                        y = c0 * in + c1 * st[sb] + c2 * st[sb + 1];
                        st[sb + 1] = st[sb];
                        st[sb] = y;
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
                case MXB_OP_DCBLOCK: {           // maxiDCBlocker::play, src/maximilian.h:1261-1266
                    const double in = fetch(g.src[0]), R = fetch(g.src[1]);
                    const double ym1 = in - st[sb] + R * st[sb + 1];
                    st[sb + 1] = ym1; st[sb] = in;
                    y = ym1;
                    break;
                }
                case MXB_OP_NONLIN: {            // maxiNonlinearity, src/maximilian.h:1076-1137
                    double x = fetch(g.src[0]);
                    const double p1 = fetch(g.src[1]), p2 = fetch(g.src[2]);
                    switch (g.kind) {
                        case MXB_NL_ATANDIST: x = (1.0 / atan(p1)) * atan(x * p1); break;
                        case MXB_NL_FASTATANDIST: x = (1.0 / (p1 / (1.0 + 0.28 * (p1 * p1)))) * ((x * p1) / (1.0 + 0.28 * ((x * p1) * (x * p1)))); break;
                        case MXB_NL_SOFTCLIP: if (x >= 1) x = 1; else if (x <= -1) x = -1; else x = (2 / 3.0) * (x - pow(x, 3.0) / 3.0); break;
                        case MXB_NL_HARDCLIP: x = x >= 1 ? 1 : (x <= -1 ? -1 : x); break;
                        case MXB_NL_ASYMCLIP: if (x >= 1) x = 1; else if (x <= -1) x = -1; else if (x < 0) x = -(pow(-x, p1)); else x = pow(x, p2); break;
                        default: x = (x / (1.0 + 0.28 * (x * x))); break;       // fastatan
                    }
                    y = x;
                    break;
                }
                case MXB_OP_DELAY:
                case MXB_OP_FLANGER: {
                    double* ring = a.rings + (size_t)s_ring[si] * (size_t)a.taps * V + vv;      // slot r of this voice at ring[r * V]
                    const double in = fetch(g.src[0]);
                    int size; double fb;
                    if (g.op == MXB_OP_DELAY) { size = (int)fetch(g.src[1]); fb = fetch(g.src[2]); }
                    else {
                        // maxiFlanger::flange, src/maximilian.h:1167-1175: lfo.triangle(speed), size = delay + lfo*depth*delay + 1 (-> int)
                        const unsigned int delay = (unsigned int)fetch(g.src[1]);
                        fb = fetch(g.src[2]);
                        const double speed = fetch(g.src[3]), depth = fetch(g.src[4]);
                        double lph = st[sb + 1], lout = st[sb + 2];
                        const double lfoVal = osc_tick<OSC_T_GENERIC>(lph, lout, 1. / (sr / (speed)), 0.0, MXB_OSC_TRIANGLE, 0.0);
                        st[sb + 1] = lph; st[sb + 2] = lout;
                        size = (int)(delay + (lfoVal * depth * delay) + 1);
                    }
                    int ph = (int)st[sb];
                    if (ph >= size) ph = 0;                                   // maxiDelayline::dl, src/maximilian.cpp:420-429
                    const int idx = min(max(ph, 0), a.taps - 1);
                    double outv = 0.0;
                    if (live) {
                        const double m = ring[(size_t)idx * V];
                        if (g.op == MXB_OP_DELAY && g.kind == MXB_DELAY_FROM_POSITION) {   // dlFromPosition, :431-439
                            int pos = (int)fetch(g.src[3]);
                            if (pos >= size) pos = 0;
                            outv = ring[(size_t)min(max(pos, 0), a.taps - 1) * V];
                            ring[(size_t)idx * V] = (m * fb) + (in * fb) * 1.0;
                        } else {
                            outv = m;
                            ring[(size_t)idx * V] = (m * fb) + (in * fb) * 0.5;
                        }
                    }
                    ph += 1;
                    st[sb] = (double)ph;
                    if (g.op == MXB_OP_FLANGER) {
                        const double normalise = (1 - fabs(outv));
                        outv *= normalise;
                        y = (outv + in) / 2.0;
                    } else y = outv;
                    break;
                }
                case MXB_OP_ADD: y = fetch(g.src[0]) + fetch(g.src[1]); break;
                case MXB_OP_SUB: y = fetch(g.src[0]) - fetch(g.src[1]); break;
                case MXB_OP_MUL: y = fetch(g.src[0]) * fetch(g.src[1]); break;
                case MXB_OP_DIV: y = fetch(g.src[0]) / fetch(g.src[1]); break;
                case MXB_OP_MIX_STEREO: {        // maxiMix::stereo, src/maximilian.cpp:503-509, accumulated over the stages of this sample
                    const double in = fetch(g.src[0]);
                    double x = fetch(g.src[1]);
                    if (x > 1) x = 1;
                    if (x < 0) x = 0;
                    if (live) { ml += in * sqrt(1.0 - x); mr += in * sqrt(x); }
                    break;
                }
                case MXB_OP_OUT:
                    if (live && a.out) a.out[(size_t)t * V + vv] = fetch(g.src[0]);
                    break;
                default: break;
            }
            if (g.dst >= 0) reg[g.dst] = y;
        }
        if (a.partials) {
            // per-warp sum in a fixed xor tree: deterministic; the second kernel adds the warps in order
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) { ml += __shfl_xor_sync(0xffffffffu, ml, m); mr += __shfl_xor_sync(0xffffffffu, mr, m); }
            if (lane == 0) {
                a.partials[((size_t)t * 2 + 0) * (size_t)a.W + (size_t)gwarp] = ml;
                a.partials[((size_t)t * 2 + 1) * (size_t)a.W + (size_t)gwarp] = mr;
            }
        }
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

int state_slots(const mxb_stage& g) {
    switch (g.op) {
        case MXB_OP_OSC: return 2;
        case MXB_OP_ENV_ADSR: case MXB_OP_ENV_AR: return 4;
        case MXB_OP_ENVGEN: return 12;
        case MXB_OP_FILTER: return 2;
        case MXB_OP_SVF: return 3;
        case MXB_OP_BIQUAD: return 2;
        case MXB_OP_DCBLOCK: return 2;
        case MXB_OP_DELAY: return 1;
        case MXB_OP_FLANGER: return 3;
        default: return 0;
    }
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

struct mxb_patch {
    mxb_ctx* ctx;
    int V, n_stages, n_params, n_consts, n_inputs, max_frames, taps, n_state, n_rings;
    std::vector<mxb_stage> stages;
    std::vector<int> state_base, ring_of;
    mxb_stage* d_stages; int* d_state_base; int* d_ring_of; double* d_consts;
    double* params; double* state; double* rings;
    double* in_stage[kMaxInputs]; size_t in_stage_len[kMaxInputs];
    double* out_stage; size_t out_stage_len;
    double* partials; double* mix_dev;
    int eg_n, eg_loop, eg_retrigger; EgStage eg[kMaxEg];
    int64_t launches;
};

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
    int n_state = 0, n_rings = 0;
    bool tables = false, eg = false;
    for (int i = 0; i < d->n_stages; ++i) {
        const mxb_stage& g = d->stages[i];
        MXB_REQUIRE(g.op >= MXB_OP_OSC && g.op <= MXB_OP_OUT, MXB_ERR_INVALID, "mxb_patch_create: stage %d: unknown op %d", i, g.op);
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
        if (g.op == MXB_OP_DELAY || g.op == MXB_OP_FLANGER) { MXB_REQUIRE(d->delay_taps > 0, MXB_ERR_INVALID, "mxb_patch_create: stage %d needs delay_taps > 0", i); ++n_rings; }
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
    int sb = 0, ri = 0;
    for (int i = 0; i < d->n_stages; ++i) {
        p->state_base.push_back(sb); sb += state_slots(p->stages[i]);
        const bool ring = p->stages[i].op == MXB_OP_DELAY || p->stages[i].op == MXB_OP_FLANGER;
        p->ring_of.push_back(ring ? ri++ : -1);
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
    MXB_REQUIRE(voice >= 0 && voice < p->V && n >= 0 && n <= p->taps, MXB_ERR_INVALID, "mxb_patch_get_ring: voice %d n %d", voice, n);
    DeviceGuard g(p->ctx->device);
    const double* src = p->rings + (size_t)p->ring_of[stage] * (size_t)p->taps * (size_t)p->V + (size_t)voice;
    if (n) MXB_CUDA(cudaMemcpy2D(dst, sizeof(double), src, sizeof(double) * (size_t)p->V, sizeof(double), (size_t)n,
                                 mem == MXB_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice));
    return MXB_OK;
}

int64_t mxb_patch_launch_count(const mxb_patch* p) { return p ? p->launches : 0; }

int32_t mxb_patch_process(mxb_patch* p, int32_t n_frames, const double* const* inputs, double* out, double* mix, int32_t mem, void* stream_) {
    NvtxRange nvtx_("mxb_patch_process");
    MXB_REQUIRE(p, MXB_ERR_INVALID, "mxb_patch_process: NULL patch");
    MXB_REQUIRE(n_frames >= 0 && n_frames <= p->max_frames, MXB_ERR_INVALID, "mxb_patch_process: n_frames %d (max_frames %d)", n_frames, p->max_frames);
    MXB_REQUIRE(mem == MXB_MEM_HOST || mem == MXB_MEM_DEVICE, MXB_ERR_INVALID, "mxb_patch_process: mem %d", mem);
    MXB_REQUIRE(p->n_inputs == 0 || inputs, MXB_ERR_INVALID, "mxb_patch_process: the patch reads %d input streams, inputs is NULL", p->n_inputs);
    for (int i = 0; i < p->n_inputs; ++i) MXB_REQUIRE(inputs[i], MXB_ERR_INVALID, "mxb_patch_process: input stream %d is NULL", i);
    if (n_frames == 0) return MXB_OK;
    DeviceGuard g(p->ctx->device);
    cudaStream_t s = (cudaStream_t)stream_;
    const size_t V = (size_t)p->V, nb = sizeof(double) * (size_t)n_frames * V;
    PatchArgs a;
    memset(&a, 0, sizeof(a));
    double* d_out = out; double* d_mix = mix;
    for (int i = 0; i < p->n_inputs; ++i) {
        a.inputs[i] = inputs[i];
        if (mem == MXB_MEM_HOST) {
            if (nb > p->in_stage_len[i]) {
                MXB_CUDA(cudaStreamSynchronize(s));
                cudaFree(p->in_stage[i]); p->in_stage[i] = nullptr; p->in_stage_len[i] = 0;
                cudaError_t e = cudaMalloc((void**)&p->in_stage[i], nb);
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
    const size_t smem = sizeof(double) * (size_t)(kMaxRegs + p->n_params + p->n_state) * kPatchThreads;
    MXB_CUDA(cudaFuncSetAttribute(patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    patch_kernel<<<grid, kPatchThreads, smem, s>>>(a);
    MXB_CUDA(cudaGetLastError());
    p->launches += 1;
    if (mix) {
        patch_mix_reduce_kernel<<<n_frames * 2, 128, 0, s>>>(p->partials, d_mix, W);
        MXB_CUDA(cudaGetLastError());
        p->launches += 1;
    }
    if (mem == MXB_MEM_HOST) {
        if (out) MXB_CUDA(cudaMemcpyAsync(out, d_out, nb, cudaMemcpyDeviceToHost, s));
        if (mix) MXB_CUDA(cudaMemcpyAsync(mix, d_mix, sizeof(double) * (size_t)n_frames * 2, cudaMemcpyDeviceToHost, s));
        MXB_CUDA(cudaStreamSynchronize(s));
    }
    return MXB_OK;
}

}  // extern "C"
