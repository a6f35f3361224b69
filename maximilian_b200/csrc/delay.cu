// K2: oscillator -> [ADSR] -> [filter] -> maxiDelayline::dl -> out / stereo mix.
//
// maxiDelayline::dl (src/maximilian.cpp:420-429) reads ring[phase] and writes it back every sample:
// 8 B read + 8 B write of ring traffic per voice-sample on top of the 8 B output -- the one genuinely
// HBM-bound stage of the path. A thread walking its own ring straight out of global memory would issue
// one 8-byte load per sample with 32 different lines per warp request; instead each WARP stages the next
// 32 ring slots of each of its 32 voices through shared memory:
//
//   stage k+1:  32 cp.async requests, one per voice, lane = slot: 256 contiguous bytes each (two wraps
//               at most: ring end and chunk end), issued before stage k is computed (double buffer)
//   stage k:    lane = voice; 32 steps of the chain, ring slot j of the window read and updated in smem
//               (row stride 33 doubles: conflict-free for 64-bit accesses)
//   write-back: the window returns to the ring the way it came, lane = slot, 256 B per request
//
// Everything is warp-private (no __syncthreads anywhere). The ring index `phase` is an int and follows
// the reference statement for statement -- `if (phase >= size) phase = 0` BEFORE the access -- so index
// sequences are bit-exact, including rings that shrink between blocks. Voices whose size is below two
// windows (64 slots; a window would meet its own write-back) take a literal per-sample path.
#include "delay_kernels.cuh"

namespace mxb {

namespace {

constexpr int kStageDoubles = 32 * 33;            // one staged window tile per warp
constexpr int kMixDoubles = 2 * kMixTT * 33;
constexpr int kFastMinSize = 2 * kDlChunk;

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;\n" ::: "memory"); }

template <int OSC, int FILT, int ENV>
__global__ void __launch_bounds__(kBankBlock) delay_bank_kernel(const BankArgs a, const DelayArgs d, const int do_out, const int do_mix) {
    const int lane = threadIdx.x & 31;
    const int gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long v0 = (long long)gwarp * 32;
    if (v0 >= a.V) return;
    const long long v = v0 + lane;
    const bool live = v < a.V;
    const long long vv = live ? v : 0;
    const size_t V = (size_t)a.V;

    extern __shared__ double smem[];
    const int per_warp = 2 * kStageDoubles + (do_mix ? kMixDoubles : 0);
    double* wsm = smem + (size_t)(threadIdx.x >> 5) * per_warp;
    double* mixtile = wsm + 2 * kStageDoubles;

    // ---- per-voice state (VPT = 1: lane = voice) ----
    double phase = a.phase[vv], oout = a.osc_out[vv];
    const double duty = (OSC == OSC_T_GENERIC) ? a.duty[vv] : 0.0;
    const double inc = (1. / (a.sr / (a.freq[vv])));
    FiltRegs fr;
    if (FILT != FILT_T_NONE) {
        fr.s0 = a.f0[vv]; fr.s1 = a.f1[vv]; fr.s2 = (FILT == FILT_T_SVF || FILT == FILT_T_SVF_LP) ? a.f2[vv] : 0.0;
        fr.c0 = a.cf[0][vv]; fr.c1 = a.cf[1][vv];
        if (FILT != FILT_T_LORES && FILT != FILT_T_HIRES) { fr.c2 = a.cf[2][vv]; fr.c3 = a.cf[3][vv]; fr.c4 = a.cf[4][vv]; }
    }
    EnvRegs er;
    if (ENV) {
        er.amp = a.env_amp[vv]; er.output = a.env_output[vv];
        er.att = a.env_att[vv]; er.dec = a.env_dec[vv]; er.sus = a.env_sus[vv]; er.rel = a.env_rel[vv];
        er.holdcount = a.env_holdcount[vv]; er.holdtime = a.env_hold[vv]; er.flags = a.env_flags[vv];
        er.on = a.trig_on ? a.trig_on[vv] : 0; er.off = a.trig_off ? a.trig_off[vv] : 0;
    }
    double gl = 0.0, gr = 0.0;
    if (do_mix) {
        double x = a.pan[vv];
        if (x > 1) x = 1;
        if (x < 0) x = 0;
        gl = live ? sqrt(1.0 - x) : 0.0;
        gr = live ? sqrt(x) : 0.0;
    }
    int ph = d.phase[vv];
    const int size = d.size[vv];
    const double fb = d.feedback[vv];
    const bool fast = live && size >= kFastMinSize;
    // ring index of the first access of the next window: the reference tests `phase >= size` before it reads
    int base = (ph >= size) ? 0 : ph;

    const int nstages = (a.n_frames + kDlChunk - 1) / kDlChunk;

    auto issue_loads = [&](double* buf, int wbase) {
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int f_i = __shfl_sync(0xffffffffu, (int)fast, i);
            if (!f_i) continue;
            const int b_i = __shfl_sync(0xffffffffu, wbase, i);
            const int s_i = __shfl_sync(0xffffffffu, size, i);
            int r = b_i + lane;
            if (r >= s_i) r -= s_i;
            cp_async8(buf + i * 33 + lane, d.ring + dl_slot(V, (size_t)(v0 + i), r));
        }
    };

    issue_loads(wsm, base);
    cp_async_commit();

    for (int k = 0; k < nstages; ++k) {
        double* buf = wsm + (k & 1) * kStageDoubles;
        const int t0 = k * kDlChunk;
        const int tn = min(kDlChunk, a.n_frames - t0);
        int next_base = base + kDlChunk;
        if (fast && next_base >= size) next_base -= size;
        if (k + 1 < nstages) issue_loads(wsm + ((k + 1) & 1) * kStageDoubles, next_base);
        cp_async_commit();
        cp_async_wait1();          // everything but the newest group has landed: stage k is in smem
        __syncwarp();

        double* row = buf + lane * 33;
        for (int h0 = 0; h0 < tn; h0 += kMixTT) {
            const int hn = min(kMixTT, tn - h0);
#pragma unroll 4
            for (int jj = 0; jj < hn; ++jj) {
                const int j = h0 + jj;
                const int t = t0 + j;
                double x = osc_tick<OSC>(phase, oout, inc, duty, a.osc_kind);
                if (ENV) x = env_tick(er, x, (t >= er.on && t < er.off) ? 1 : 0);
                x = filt_tick<FILT>(fr, x, a.svf_mix);
                // maxiDelayline::dl, src/maximilian.cpp:420-429
                double y = 0.0;
                if (fast) {
                    const double m = row[j];
                    row[j] = (m * fb) + (x * fb) * 0.5;
                    y = m;
                } else if (live) {
                    if (ph >= size) ph = 0;
                    const int idx = min(max(ph, 0), d.taps - 1);     // size <= taps is enforced when the parameter is set
                    double* slot = d.ring + dl_slot(V, (size_t)v, idx);
                    const double m = *slot;
                    *slot = (m * fb) + (x * fb) * 0.5;
                    ph += 1;
                    y = m;
                }
                if (do_out && live) {
                    if (a.out_f32) __stcs((float*)a.out + (size_t)t * V + (size_t)v, (float)y);
                    else __stcs((double*)a.out + (size_t)t * V + (size_t)v, y);
                }
                if (do_mix) {
                    mixtile[(0 * kMixTT + jj) * 33 + lane] = y * gl;
                    mixtile[(1 * kMixTT + jj) * 33 + lane] = y * gr;
                }
            }
            if (do_mix) {
                __syncwarp();
                const int ch = lane >> 4, rw = lane & 15;
                if (rw < hn) {
                    const double* r = mixtile + (ch * kMixTT + rw) * 33;
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 32; ++q) s += r[q];
                    a.partials[((size_t)(t0 + h0 + rw) * 2 + ch) * (size_t)a.W + (size_t)gwarp] = s;
                }
                __syncwarp();
            }
        }
        __syncwarp();
        // write the window back, lane = slot
#pragma unroll 4
        for (int i = 0; i < 32; ++i) {
            const int f_i = __shfl_sync(0xffffffffu, (int)fast, i);
            if (!f_i) continue;
            const int b_i = __shfl_sync(0xffffffffu, base, i);
            const int s_i = __shfl_sync(0xffffffffu, size, i);
            if (lane < tn) {
                int r = b_i + lane;
                if (r >= s_i) r -= s_i;
                d.ring[dl_slot(V, (size_t)(v0 + i), r)] = buf[i * 33 + lane];
            }
        }
        __syncwarp();
        if (fast) {
            int last = base + tn - 1;
            if (last >= size) last -= size;
            ph = last + 1;                     // phase += 1 after the last access of the window
            base = (tn == kDlChunk) ? next_base : ((ph >= size) ? 0 : ph);
        }
    }

    if (live) {
        a.phase[v] = phase;
        if (OSC == OSC_T_GENERIC) a.osc_out[v] = oout;
        if (FILT != FILT_T_NONE) {
            a.f0[v] = fr.s0; a.f1[v] = fr.s1;
            if (FILT == FILT_T_SVF || FILT == FILT_T_SVF_LP) a.f2[v] = fr.s2;
        }
        if (ENV) { a.env_amp[v] = er.amp; a.env_output[v] = er.output; a.env_holdcount[v] = er.holdcount; a.env_flags[v] = er.flags; }
        d.phase[v] = ph;
    }
}

template <int OSC, int FILT, int ENV>
int launch_one(const BankArgs& a, const DelayArgs& d, bool out, bool mix, int grid, cudaStream_t s) {
    const size_t smem = sizeof(double) * (kBankBlock / 32) * (2 * kStageDoubles + (mix ? kMixDoubles : 0));
    auto kern = delay_bank_kernel<OSC, FILT, ENV>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("delay_bank_kernel smem attribute (%zu B): %s", smem, cudaGetErrorString(e)); return MXB_ERR_CUDA; }
    kern<<<grid, kBankBlock, smem, s>>>(a, d, out ? 1 : 0, mix ? 1 : 0);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("delay_bank_kernel launch: %s", cudaGetErrorString(e)); return MXB_ERR_CUDA; }
    return MXB_OK;
}

template <int FILT>
int launch_filt(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, bool out, bool mix, int grid, cudaStream_t s) {
    if (osc_saw) return env ? launch_one<OSC_T_SAW, FILT, 1>(a, d, out, mix, grid, s) : launch_one<OSC_T_SAW, FILT, 0>(a, d, out, mix, grid, s);
    return env ? launch_one<OSC_T_GENERIC, FILT, 1>(a, d, out, mix, grid, s) : launch_one<OSC_T_GENERIC, FILT, 0>(a, d, out, mix, grid, s);
}

}  // namespace

int delay_bank_warps(int V) { return (V + 31) / 32; }   // warps that own at least one voice

int launch_delay_bank(const BankArgs& a_in, DelayArgs& d, int filt_kind, bool svf_lp, int env, bool out, bool mix, cudaStream_t s) {
    BankArgs a = a_in;
    const int grid = (a.V + kBankBlock - 1) / kBankBlock;
    a.W = delay_bank_warps(a.V);
    d.W_out = a.W;
    const int saw = a.osc_kind == MXB_OSC_SAW;
    switch (filt_kind) {
        case MXB_FILT_NONE:   return launch_filt<FILT_T_NONE>(a, d, saw, env, out, mix, grid, s);
        case MXB_FILT_LORES:  return launch_filt<FILT_T_LORES>(a, d, saw, env, out, mix, grid, s);
        case MXB_FILT_HIRES:  return launch_filt<FILT_T_HIRES>(a, d, saw, env, out, mix, grid, s);
        case MXB_FILT_SVF:    return svf_lp ? launch_filt<FILT_T_SVF_LP>(a, d, saw, env, out, mix, grid, s)
                                            : launch_filt<FILT_T_SVF>(a, d, saw, env, out, mix, grid, s);
        case MXB_FILT_BIQUAD: return launch_filt<FILT_T_BIQUAD>(a, d, saw, env, out, mix, grid, s);
        default: break;
    }
    set_error("launch_delay_bank: filt_kind %d", filt_kind);
    return MXB_ERR_UNSUPPORTED;
}

}  // namespace mxb
