/*
 * maximilian_b200.hpp -- C++ host layer over the C ABI (maxib200.h), keeping the reference's class surface:
 *   maxiSettings, maxiOsc, maxiFilter, maxiSVF, maxiBiquad, maxiEnv, maxiDelayline, maxiMix   (src/maximilian.h)
 *   maxiFFT, maxiIFFT, maxiMFCC                                                               (src/libs/maxiFFT.h, maxiMFCC.h)
 * with the same method names and argument meaning, restructured from "one call = one sample of one object"
 * to "one call = one BLOCK of a whole bank of voices":
 *
 *   reference patch (per sample, per voice)              this header (per block, V voices)
 *   ------------------------------------------------     -------------------------------------------------
 *   maxiOsc osc[V]; maxiSVF svf[V]; maxiMix mix;          maxiVoices voices(V);
 *                                                          maxiOsc osc(voices); maxiSVF svf(voices); maxiMix mix(voices);
 *   void play(double* out) {                              void play(maxiVoices& v) {
 *     for (i < V) {                                         maxiSignal w = svf.play(osc.saw(freq), 1, 0, 0, 0);
 *       w = svf[i].play(osc[i].saw(f[i]), 1,0,0,0);         mix.stereo(w, v.bus(), pan);
 *       mix.stereo(w, two, pan[i]); out[0] += two[0]...   }
 *   } }
 *   routing(): for each frame: play(frame)                maxiRouting(): play(voices) once, then one fused kernel
 *   (cpp/commandline/player.cpp:25-44)                    renders nBufferFrames frames of every voice
 *
 * The calls inside play() do not compute anything on the host: they describe the chain
 * (oscillator -> [maxiEnv::adsr] -> [filter] -> [maxiDelayline::dl] -> maxiMix::stereo / per-voice output) and hand
 * over per-voice parameters; maxiVoices::render() executes the block on the GPU through mxb_bank_process. Chains the
 * kernels do not implement are rejected with an exception, never emulated: there is no CPU path in here.
 *
 * Error behaviour: the reference has none (UB, or exit(1) in fft.cpp:67,131). Here every failure of the C ABI
 * surfaces as a maxiError (std::runtime_error) carrying mxb_last_error().
 */
#ifndef MAXIMILIAN_B200_HPP
#define MAXIMILIAN_B200_HPP

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "maxib200.h"

/* Source compatibility with patches written against src/maximilian.h: the reference header pulls
 * namespace std in (:54) and defines these (:55-67), and user patches rely on both (`vector<double>`, `cout`, PI). */
#ifndef MAXIMILIAN_B200_NO_STD_NAMESPACE
using namespace std;
#endif
#ifndef PI
#define PI 3.1415926535897932384626433832795
#endif
#ifndef TWOPI
#define TWOPI 6.283185307179586476925286766559
#endif
#ifndef CHEERP_EXPORT
#define CHEERP_EXPORT
#endif

struct maxiError : std::runtime_error {
    int code;
    maxiError(int c, const std::string& what) : std::runtime_error(what), code(c) {}
};

namespace maxib200_detail {
inline void check(int rc, const char* what) {
    if (rc != MXB_OK) throw maxiError(rc, std::string(what) + ": " + mxb_last_error());
}
}  // namespace maxib200_detail

/* maxiSettings, src/maximilian.h:117-163: process-global sample rate / channels / buffer size. */
class maxiSettings {
public:
    static inline size_t sampleRate = 44100;   /* defaults of src/maximilian.cpp:57-59 */
    static inline size_t channels = 2;
    static inline size_t bufferSize = 1024;
    static void setup(size_t initSampleRate, size_t initChannels, size_t initBufferSize) {
        sampleRate = initSampleRate; channels = initChannels; bufferSize = initBufferSize;
    }
    static size_t getSampleRate() { return sampleRate; }
};

/* A per-voice parameter: one double per voice, or a scalar broadcast to all voices (what the reference passes
 * by value on every sample). */
class maxiParam {
public:
    maxiParam(double scalar = 0.0) : scalar_(scalar), vec_(nullptr) {}
    maxiParam(const std::vector<double>& perVoice) : scalar_(0.0), vec_(&perVoice) {}
    bool isScalar() const { return vec_ == nullptr; }
    double scalar() const { return scalar_; }
    const std::vector<double>& vec() const { return *vec_; }
private:
    double scalar_;
    const std::vector<double>* vec_;
};

/* trigger_v(t) == 1 for on[v] <= t < off[v], t counting frames inside the current block (a sample-accurate note gate) */
struct maxiGate {
    const std::vector<int32_t>* on = nullptr;
    const std::vector<int32_t>* off = nullptr;
    maxiGate() {}
    maxiGate(const std::vector<int32_t>& on_, const std::vector<int32_t>& off_) : on(&on_), off(&off_) {}
};

class maxiVoices;
/* the signal flowing between stages: a token tying a stage output to its bank */
struct maxiSignal {
    maxiVoices* voices = nullptr;
    int stage = 0;     /* 1 osc, 2 env, 3 filter, 4 delay */
};
struct maxiBus { maxiVoices* voices = nullptr; };

/* The bank of voices behind one play(): owns the mxb_ctx / mxb_bank handles. */
class maxiVoices {
public:
    explicit maxiVoices(int voices, int device = 0) : V_(voices), device_(device) {
        for (auto& d : dirty_) d = false;
        std::memset(&desc_, 0, sizeof(desc_));
        desc_.voices = voices; desc_.osc_kind = -1;
    }
    ~maxiVoices() { if (bank_) mxb_bank_destroy(bank_); if (ctx_) mxb_ctx_destroy(ctx_); }
    maxiVoices(const maxiVoices&) = delete;
    maxiVoices& operator=(const maxiVoices&) = delete;

    int size() const { return V_; }
    maxiBus bus() { return maxiBus{this}; }

    /* Run the chain described since the last render for nFrames frames.
     * out:  per-voice samples [nFrames][V] (host memory) or nullptr;
     * mix:  stereo bus [nFrames][2] (host memory, interleaved like RTAUDIO_FLOAT64) or nullptr. */
    void render(int nFrames, double* out, double* mix) {
        using maxib200_detail::check;
        if (desc_.osc_kind < 0) throw maxiError(MXB_ERR_STATE, "maxiVoices::render: play() described no oscillator");
        if (!bank_) {
            check(mxb_ctx_create(device_, (int32_t)maxiSettings::sampleRate, &ctx_), "mxb_ctx_create");
            desc_.max_frames = (int32_t)(nFrames > (int)maxiSettings::bufferSize ? nFrames : (int)maxiSettings::bufferSize);
            check(mxb_bank_create(ctx_, &desc_, &bank_), "mxb_bank_create");
            built_ = desc_;
        } else if (std::memcmp(&built_, &desc_, sizeof(desc_)) != 0 && !sameChain()) {
            throw maxiError(MXB_ERR_UNSUPPORTED, "maxiVoices: play() changed the chain after the first block");
        }
        for (int id = 0; id < MXB_P_COUNT; ++id) {
            if (!dirty_[id]) continue;
            check(mxb_bank_set_param(bank_, id, params_[id].data(), MXB_MEM_HOST), "mxb_bank_set_param");
            dirty_[id] = false;
        }
        const int32_t* on = gate_.on ? gate_.on->data() : nullptr;
        const int32_t* off = gate_.off ? gate_.off->data() : nullptr;
        check(mxb_bank_process(bank_, nFrames, on, off, out, MXB_F64, (wantMix_ ? mix : nullptr), MXB_MEM_HOST, nullptr), "mxb_bank_process");
        gate_ = maxiGate();
    }

    /* state read-back (checkpointing / tests): MXB_P_PHASE, MXB_S_* */
    std::vector<double> state(int id) {
        std::vector<double> v((size_t)V_);
        maxib200_detail::check(mxb_bank_get_state(bank_, id, v.data(), MXB_MEM_HOST), "mxb_bank_get_state");
        return v;
    }
    mxb_bank* handle() { return bank_; }

private:
    friend class maxiOsc; friend class maxiFilter; friend class maxiSVF; friend class maxiBiquad;
    friend class maxiEnv; friend class maxiDelayline; friend class maxiMix;

    bool sameChain() const {
        return built_.osc_kind == desc_.osc_kind && built_.filt_kind == desc_.filt_kind && built_.env_kind == desc_.env_kind &&
               built_.biquad_type == desc_.biquad_type && built_.delay_taps == desc_.delay_taps && built_.delay_mode == desc_.delay_mode &&
               std::memcmp(built_.svf_mix, desc_.svf_mix, sizeof(desc_.svf_mix)) == 0;
    }
    void setParam(int id, const maxiParam& p) {
        std::vector<double>& dst = params_[id];
        if (p.isScalar()) {
            if (dst.size() == (size_t)V_ && scalarSet_[id] && scalarVal_[id] == p.scalar()) return;   /* unchanged since last block */
            dst.assign((size_t)V_, p.scalar());
            scalarSet_[id] = true; scalarVal_[id] = p.scalar();
        } else {
            if (p.vec().size() != (size_t)V_) throw maxiError(MXB_ERR_INVALID, "maxiParam: per-voice vector has the wrong length");
            if (dst.size() == (size_t)V_ && !scalarSet_[id] && std::memcmp(dst.data(), p.vec().data(), sizeof(double) * (size_t)V_) == 0) return;
            dst = p.vec();
            scalarSet_[id] = false;
        }
        dirty_[id] = true;
    }

    int V_, device_;
    mxb_ctx* ctx_ = nullptr;
    mxb_bank* bank_ = nullptr;
    mxb_bank_desc desc_, built_;
    std::vector<double> params_[MXB_P_COUNT];
    bool dirty_[MXB_P_COUNT];
    bool scalarSet_[MXB_P_COUNT] = {};
    double scalarVal_[MXB_P_COUNT] = {};
    maxiGate gate_;
    bool wantMix_ = false;
};

/* maxiOsc, src/maximilian.h:169-215 / src/maximilian.cpp:209-373 */
class maxiOsc {
public:
    explicit maxiOsc(maxiVoices& v) : v_(&v) {}
    maxiSignal sinewave(const maxiParam& frequency) { return osc(MXB_OSC_SINEWAVE, frequency); }
    maxiSignal coswave(const maxiParam& frequency) { return osc(MXB_OSC_COSWAVE, frequency); }
    maxiSignal phasor(const maxiParam& frequency) { return osc(MXB_OSC_PHASOR, frequency); }
    maxiSignal saw(const maxiParam& frequency) { return osc(MXB_OSC_SAW, frequency); }
    maxiSignal square(const maxiParam& frequency) { return osc(MXB_OSC_SQUARE, frequency); }
    maxiSignal triangle(const maxiParam& frequency) { return osc(MXB_OSC_TRIANGLE, frequency); }
    maxiSignal impulse(const maxiParam& frequency) { return osc(MXB_OSC_IMPULSE, frequency); }
    maxiSignal pulse(const maxiParam& frequency, const maxiParam& duty) { v_->setParam(MXB_P_DUTY, duty); return osc(MXB_OSC_PULSE, frequency); }
    maxiSignal phasorBetween(const maxiParam& frequency, const maxiParam& startphase, const maxiParam& endphase) {
        v_->setParam(MXB_P_PHASOR_START, startphase); v_->setParam(MXB_P_PHASOR_END, endphase);
        return osc(MXB_OSC_PHASORBETWEEN, frequency);
    }
    void phaseReset(const maxiParam& phaseIn) { v_->setParam(MXB_P_PHASE, phaseIn); }
private:
    maxiSignal osc(int kind, const maxiParam& f) { v_->desc_.osc_kind = kind; v_->setParam(MXB_P_FREQ, f); return maxiSignal{v_, 1}; }
    maxiVoices* v_;
};

/* maxiFilter::lores / hires, src/maximilian.cpp:455-484 */
class maxiFilter {
public:
    explicit maxiFilter(maxiVoices& v) : v_(&v) {}
    maxiSignal lores(maxiSignal input, const maxiParam& cutoff1, const maxiParam& resonance) { return f(MXB_FILT_LORES, input, cutoff1, resonance); }
    maxiSignal hires(maxiSignal input, const maxiParam& cutoff1, const maxiParam& resonance) { return f(MXB_FILT_HIRES, input, cutoff1, resonance); }
private:
    maxiSignal f(int kind, maxiSignal in, const maxiParam& c, const maxiParam& r) {
        if (in.voices != v_ || in.stage < 1 || in.stage > 2) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiFilter: input must be an oscillator or envelope of the same maxiVoices");
        v_->desc_.filt_kind = kind; v_->setParam(MXB_P_CUTOFF, c); v_->setParam(MXB_P_RESONANCE, r);
        return maxiSignal{v_, 3};
    }
    maxiVoices* v_;
};

/* maxiSVF, src/maximilian.h:1281-1338 */
class maxiSVF {
public:
    explicit maxiSVF(maxiVoices& v) : v_(&v) {}
    void setCutoff(const maxiParam& cutoff) { v_->setParam(MXB_P_CUTOFF, cutoff); }
    void setResonance(const maxiParam& q) { v_->setParam(MXB_P_RESONANCE, q); }
    maxiSignal play(maxiSignal w, double lpmix, double bpmix, double hpmix, double notchmix) {
        if (w.voices != v_ || w.stage < 1 || w.stage > 2) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiSVF::play: input must be an oscillator or envelope of the same maxiVoices");
        v_->desc_.filt_kind = MXB_FILT_SVF;
        v_->desc_.svf_mix[0] = lpmix; v_->desc_.svf_mix[1] = bpmix; v_->desc_.svf_mix[2] = hpmix; v_->desc_.svf_mix[3] = notchmix;
        return maxiSignal{v_, 3};
    }
private:
    maxiVoices* v_;
};

/* maxiBiquad, src/maximilian.h:1343-1486 */
class maxiBiquad {
public:
    enum filterTypes { LOWPASS, HIGHPASS, BANDPASS, NOTCH, PEAK, LOWSHELF, HIGHSHELF };
    explicit maxiBiquad(maxiVoices& v) : v_(&v) {}
    void set(filterTypes filtType, const maxiParam& cutoff, const maxiParam& Q, const maxiParam& peakGain) {
        v_->desc_.biquad_type = (int)filtType;
        v_->setParam(MXB_P_GAIN, peakGain); v_->setParam(MXB_P_CUTOFF, cutoff); v_->setParam(MXB_P_RESONANCE, Q);
    }
    maxiSignal play(maxiSignal input) {
        if (input.voices != v_ || input.stage < 1 || input.stage > 2) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiBiquad::play: input must be an oscillator or envelope of the same maxiVoices");
        v_->desc_.filt_kind = MXB_FILT_BIQUAD;
        return maxiSignal{v_, 3};
    }
private:
    maxiVoices* v_;
};

/* maxiEnv (ADSR), src/maximilian.h:888-932 / src/maximilian.cpp:1415-1494 */
class maxiEnv {
public:
    explicit maxiEnv(maxiVoices& v) : v_(&v) {}
    void setAttack(const maxiParam& attackMS) { coeff(MXB_P_ENV_ATTACK, 0, attackMS); }
    void setAttackMS(const maxiParam& attackMS) { coeff(MXB_P_ENV_ATTACK, 1, attackMS); }
    void setDecay(const maxiParam& decayMS) { coeff(MXB_P_ENV_DECAY, 2, decayMS); }
    void setRelease(const maxiParam& releaseMS) { coeff(MXB_P_ENV_RELEASE, 2, releaseMS); }
    void setSustain(const maxiParam& sustainL) { v_->setParam(MXB_P_ENV_SUSTAIN, sustainL); }
    void setHoldtime(const maxiParam& holdtime) { v_->setParam(MXB_P_ENV_HOLDTIME, holdtime); }   /* the public member `holdtime` */
    maxiSignal adsr(maxiSignal input, const maxiGate& trigger) {
        if (input.voices != v_ || input.stage != 1) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiEnv::adsr: input must be the oscillator of the same maxiVoices");
        v_->desc_.env_kind = MXB_ENV_ADSR; v_->gate_ = trigger;
        return maxiSignal{v_, 2};
    }
    /* the overload that takes the raw coefficients as arguments (src/maximilian.cpp:1362-1413): the same state machine */
    maxiSignal adsr(maxiSignal input, const maxiParam& attack, const maxiParam& decay, const maxiParam& sustain, const maxiParam& release,
                    const maxiParam& holdtime, const maxiGate& trigger) {
        v_->setParam(MXB_P_ENV_ATTACK, attack); v_->setParam(MXB_P_ENV_DECAY, decay); v_->setParam(MXB_P_ENV_SUSTAIN, sustain);
        v_->setParam(MXB_P_ENV_RELEASE, release); v_->setParam(MXB_P_ENV_HOLDTIME, holdtime);
        return adsr(input, trigger);
    }
    /* maxiEnv::ar(input, attack, release, holdtime, trigger): attack/release are the raw per-sample coefficients the
     * reference takes as arguments (defaults 1, 0.9, holdtime 1) */
    maxiSignal ar(maxiSignal input, const maxiParam& attack, const maxiParam& release, const maxiParam& holdtime, const maxiGate& trigger) {
        if (input.voices != v_ || input.stage != 1) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiEnv::ar: input must be the oscillator of the same maxiVoices");
        v_->desc_.env_kind = MXB_ENV_AR; v_->gate_ = trigger;
        v_->setParam(MXB_P_ENV_ATTACK, attack); v_->setParam(MXB_P_ENV_RELEASE, release); v_->setParam(MXB_P_ENV_HOLDTIME, holdtime);
        return maxiSignal{v_, 2};
    }
private:
    void coeff(int id, int kind, const maxiParam& ms) {
        std::vector<double> in = ms.isScalar() ? std::vector<double>(1, ms.scalar()) : ms.vec();
        std::vector<double> out(in.size());
        maxib200_detail::check(mxb_env_coeffs(kind, in.data(), (int64_t)in.size(), (int32_t)maxiSettings::sampleRate, out.data()), "mxb_env_coeffs");
        if (ms.isScalar()) v_->setParam(id, maxiParam(out[0])); else { tmp_[id - MXB_P_ENV_ATTACK] = out; v_->setParam(id, maxiParam(tmp_[id - MXB_P_ENV_ATTACK])); }
    }
    maxiVoices* v_;
    std::vector<double> tmp_[4];
};

/* maxiDelayline, src/maximilian.h:266-284 / src/maximilian.cpp:415-429 */
class maxiDelayline {
public:
    /* capacity: ring slots per voice (the reference allocates 705600 for every object) */
    maxiDelayline(maxiVoices& v, int capacity) : v_(&v), capacity_(capacity) {}
    maxiSignal dl(maxiSignal input, const maxiParam& size, const maxiParam& feedback) {
        if (input.voices != v_ || input.stage < 1 || input.stage > 3) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiDelayline::dl: input must come from the same maxiVoices");
        v_->desc_.delay_taps = capacity_; v_->desc_.delay_mode = MXB_DELAY_DL;
        v_->setParam(MXB_P_DELAY_SIZE, size); v_->setParam(MXB_P_DELAY_FEEDBACK, feedback);
        return maxiSignal{v_, 4};
    }
    maxiSignal dlFromPosition(maxiSignal input, const maxiParam& size, const maxiParam& feedback, const maxiParam& position) {
        if (input.voices != v_ || input.stage < 1 || input.stage > 3) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiDelayline::dlFromPosition: input must come from the same maxiVoices");
        v_->desc_.delay_taps = capacity_; v_->desc_.delay_mode = MXB_DELAY_FROM_POSITION;
        v_->setParam(MXB_P_DELAY_SIZE, size); v_->setParam(MXB_P_DELAY_FEEDBACK, feedback); v_->setParam(MXB_P_DELAY_POSITION, position);
        return maxiSignal{v_, 4};
    }
private:
    maxiVoices* v_;
    int capacity_;
};

/* maxiMix::stereo, src/maximilian.h:400 / src/maximilian.cpp:503-509; the per-voice results are summed into the bus */
class maxiMix {
public:
    explicit maxiMix(maxiVoices& v) : v_(&v) {}
    void stereo(maxiSignal input, maxiBus two, const maxiParam& x) {
        if (input.voices != v_ || two.voices != v_) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiMix::stereo: signal and bus must belong to the same maxiVoices");
        v_->setParam(MXB_P_PAN, x); v_->wantMix_ = true;
    }
private:
    maxiVoices* v_;
};

/* ---- block-dispatch shim: stands in for routing() of cpp/commandline/player.cpp:25-44 (RtAudio callback signature,
 * cpp/commandline/RtAudio.h:205-209). The user supplies  void play(maxiVoices&)  -- the block-rate twin of the
 * reference's  void play(double*)  -- and passes the maxiVoices as userData; the interleaved RTAUDIO_FLOAT64
 * stereo buffer is exactly the bank's mix bus. ---- */
void play(maxiVoices& voices);
inline int maxiRouting(void* outputBuffer, void* /*inputBuffer*/, unsigned int nBufferFrames, double /*streamTime*/,
                       unsigned int /*status*/, void* userData) {
    maxiVoices* v = static_cast<maxiVoices*>(userData);
    play(*v);
    v->render((int)nBufferFrames, nullptr, static_cast<double*>(outputBuffer));
    return 0;
}

/* ---- spectral classes: C channels at once ---- */

/* maxiMFCC, src/libs/maxiMFCC.h:40-211 */
class maxiMFCC {
public:
    maxiMFCC() {}
    ~maxiMFCC() { if (h_) mxb_mfcc_destroy(h_); if (ctx_) mxb_ctx_destroy(ctx_); }
    maxiMFCC(const maxiMFCC&) = delete;
    void setup(unsigned int numBins, unsigned int numFilters, unsigned int numCoeffs, double minFreq, double maxFreq, int device = 0) {
        using maxib200_detail::check;
        check(mxb_ctx_create(device, (int32_t)maxiSettings::sampleRate, &ctx_), "mxb_ctx_create");
        check(mxb_mfcc_create(ctx_, (int32_t)numBins, (int32_t)numFilters, (int32_t)numCoeffs, minFreq, maxFreq, &h_), "mxb_mfcc_create");
        bins_ = (int)numBins; coeffs_ = (int)numCoeffs;
    }
    /* powerSpectrum: frames x numBins magnitudes (host); returns frames x numCoeffs */
    std::vector<double>& mfcc(const std::vector<float>& powerSpectrum) {
        const int64_t n = (int64_t)(powerSpectrum.size() / (size_t)bins_);
        out_.resize((size_t)n * (size_t)coeffs_);
        maxib200_detail::check(mxb_mfcc_process(h_, powerSpectrum.data(), n, out_.data(), nullptr, MXB_MEM_HOST, nullptr), "mxb_mfcc_process");
        return out_;
    }
    mxb_mfcc* handle() { return h_; }
private:
    mxb_ctx* ctx_ = nullptr; mxb_mfcc* h_ = nullptr; int bins_ = 0, coeffs_ = 0;
    std::vector<double> out_;
};

/* maxiFFT, src/libs/maxiFFT.h:46-120: process() takes one block of samples per channel instead of one sample */
class maxiFFT {
public:
    enum fftModes { NO_POLAR_CONVERSION = 0, WITH_POLAR_CONVERSION = 1 };
    explicit maxiFFT(int channels = 1, int device = 0) : C_(channels), device_(device) {}
    ~maxiFFT() { if (h_) mxb_stft_destroy(h_); if (ctx_) mxb_ctx_destroy(ctx_); }
    maxiFFT(const maxiFFT&) = delete;
    void setup(int fftSize = 1024, int hopSize = 512, int windowSize = 0) {
        using maxib200_detail::check;
        if (windowSize > fftSize) throw maxiError(MXB_ERR_INVALID, "maxiFFT::setup: windowSize > fftSize overflows the reference's buffer (maxiFFT.cpp:48-51); rejected");
        check(mxb_ctx_create(device_, (int32_t)maxiSettings::sampleRate, &ctx_), "mxb_ctx_create");
        check(mxb_stft_create(ctx_, C_, fftSize, hopSize, &h_), "mxb_stft_create");
        fftSize_ = fftSize; hop_ = hopSize; bins_ = fftSize / 2;
    }
    /* values: planar [channels][nSamples] (host). Returns true when at least one new frame fired (the reference
     * returns true on the sample that completes a frame); frames() tells how many per channel. */
    bool process(const float* values, int nSamples, fftModes mode = WITH_POLAR_CONVERSION) {
        const int maxf = nSamples / hop_ + 2;
        const size_t len = (size_t)C_ * (size_t)maxf * (size_t)bins_;
        const bool polar = mode == WITH_POLAR_CONVERSION;
        re_.resize(len); im_.resize(len);
        if (polar) {
            mags_.resize(len); phases_.resize(len); magsdb_.resize(len);
            flatness_.assign((size_t)C_ * (size_t)maxf, 0.f); centroid_.assign((size_t)C_ * (size_t)maxf, 0.f);
        }
        mxb_stft_outputs o;
        o.mags = polar ? mags_.data() : nullptr; o.phases = polar ? phases_.data() : nullptr;
        o.re = re_.data(); o.im = im_.data();
        o.mags_db = polar ? magsdb_.data() : nullptr;
        o.flatness = polar ? flatness_.data() : nullptr; o.centroid = polar ? centroid_.data() : nullptr;
        o.coeffs = nullptr;
        int32_t nf = 0;
        maxib200_detail::check(mxb_stft_process2(h_, values, nSamples, 1, nSamples, maxf, &o, nullptr, &nf, MXB_MEM_HOST, nullptr),
                               "mxb_stft_process2");
        frames_ = nf; maxf_ = maxf;
        return nf > 0;
    }
    int frames() const { return frames_; }
    int frameStride() const { return maxf_; }     /* frame f of channel c starts at ((c*frameStride()) + f)*getNumBins() */
    std::vector<float>& getMagnitudes() { return mags_; }
    std::vector<float>& getPhases() { return phases_; }
    /* the spectral post-processors (src/libs/maxiFFT.cpp:101-132), computed by the same kernel while the magnitudes are on chip */
    std::vector<float>& magsToDB() { return magsdb_; }                 /* laid out like getMagnitudes() */
    std::vector<float>& getMagnitudesDB() { return magsdb_; }
    std::vector<float>& spectralFlatness() { return flatness_; }       /* frame f of channel c at c*frameStride() + f */
    std::vector<float>& spectralCentroid() { return centroid_; }
    float* getReal() { return re_.data(); }
    float* getImag() { return im_.data(); }
    int getNumBins() { return bins_; }
    int getFFTSize() { return fftSize_; }
    int getHopSize() { return hop_; }
    int getWindowSize() { return fftSize_; }
private:
    int C_, device_;
    mxb_ctx* ctx_ = nullptr; mxb_stft* h_ = nullptr;
    int fftSize_ = 0, hop_ = 0, bins_ = 0, frames_ = 0, maxf_ = 0;
    std::vector<float> mags_, phases_, re_, im_, magsdb_, flatness_, centroid_;
};

/* maxiIFFT (SPECTRUM mode), src/libs/maxiFFT.h:125-156; COMPLEX mode yields zeros in the reference on Linux and is not offered */
class maxiIFFT {
public:
    enum fftModes { SPECTRUM = 0 };
    explicit maxiIFFT(int channels = 1, int device = 0) : C_(channels), device_(device) {}
    ~maxiIFFT() { if (h_) mxb_istft_destroy(h_); if (ctx_) mxb_ctx_destroy(ctx_); }
    maxiIFFT(const maxiIFFT&) = delete;
    void setup(int fftSize = 1024, int hopSize = 512, int windowSize = 0) {
        using maxib200_detail::check;
        /* the reference windows with genWindow(3, windowSize ? windowSize : fftSize) (maxiFFT.cpp:141-152): another size is another window */
        if (windowSize != 0 && windowSize != fftSize) throw maxiError(MXB_ERR_UNSUPPORTED, "maxiIFFT::setup: windowSize must be 0 or fftSize");
        check(mxb_ctx_create(device_, (int32_t)maxiSettings::sampleRate, &ctx_), "mxb_ctx_create");
        check(mxb_istft_create(ctx_, C_, fftSize, hopSize, &h_), "mxb_istft_create");
        hop_ = hopSize; bins_ = fftSize / 2;
    }
    /* data1/data2: magnitudes / phases, frame f of channel c at (c*frames + f)*bins; returns planar [channels][frames*hop] */
    std::vector<float>& process(const std::vector<float>& data1, const std::vector<float>& data2, int frames, fftModes = SPECTRUM) {
        out_.resize((size_t)C_ * (size_t)frames * (size_t)hop_);
        maxib200_detail::check(mxb_istft_process(h_, data1.data(), data2.data(), frames, out_.data(), MXB_MEM_HOST, nullptr), "mxb_istft_process");
        return out_;
    }
    int getNumBins() { return bins_; }
private:
    int C_, device_;
    mxb_ctx* ctx_ = nullptr; mxb_istft* h_ = nullptr; int hop_ = 0, bins_ = 0;
    std::vector<float> out_;
};

#endif /* MAXIMILIAN_B200_HPP */
