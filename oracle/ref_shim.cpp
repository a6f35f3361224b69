/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the UNMODIFIED reference classes (compiled from /root/reference/src by
 * oracle/Makefile, never copied into this repository) behind oracle_api.h.
 * Built into oracle/_ref/libmaxiref.so. This file contains no DSP of its own:
 * every sample comes out of a reference method call.
 *
 * Construction rules: objects live in zero-filled storage
 * (calloc + placement new) because maxiDelayline::phase, every maxiEnv field and
 * maxiFilter::outputs[] are otherwise indeterminate; maxiSettings::setup() runs
 * before any object is built (maxiSVF and maxiMFCC read the rate at construction).
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <iostream>
#include <algorithm>
#include <numeric>
#include <functional>
#include <string>
#include <map>
#include <valarray>

/* state read-back for the tests needs private members (maxiOsc::phase, maxiDelayline::memory ...):
 * the Makefile passes -fno-access-control, which switches off access checking without
 * touching layout or code generation. */
#include "maximilian.h"
#include "maxiFFT.h"
#include "maxiMFCC.h"

#include "oracle_api.h"

namespace {

struct RefVoice {
    maxiOsc osc;
    maxiFilter filt;
    maxiSVF svf;
    maxiBiquad bq;
    maxiEnv env;
    maxiMix mixer;
};

struct RefBank {
    mxo_chain chain;
    int V;
    RefVoice* voices;          /* calloc'ed, placement-constructed */
    maxiDelayline** delays;    /* one 5.6 MB object per voice when delay_on */
    std::vector<double> p[MXO_P_COUNT];
};

template <class T> T* zeroed_new() {
    void* mem = calloc(1, sizeof(T));
    if (!mem) return nullptr;
    return new (mem) T();
}

void apply_svf(RefBank* b) {
    if (b->p[MXO_P_CUTOFF].empty() || b->p[MXO_P_RESONANCE].empty()) return;
    for (int v = 0; v < b->V; ++v) {
        b->voices[v].svf.setCutoff(b->p[MXO_P_CUTOFF][v]);
        b->voices[v].svf.setResonance(b->p[MXO_P_RESONANCE][v]);
    }
}
void apply_biquad(RefBank* b) {
    if (b->p[MXO_P_CUTOFF].empty() || b->p[MXO_P_RESONANCE].empty()) return;
    for (int v = 0; v < b->V; ++v) {
        double g = b->p[MXO_P_GAIN].empty() ? 0.0 : b->p[MXO_P_GAIN][v];
        b->voices[v].bq.set((maxiBiquad::filterTypes)b->chain.biquad_type,
                            b->p[MXO_P_CUTOFF][v], b->p[MXO_P_RESONANCE][v], g);
    }
}

inline double run_osc(maxiOsc& o, int kind, double f, double duty, double pstart, double pend) {
    switch (kind) {
        case MXO_OSC_SINEWAVE: return o.sinewave(f);
        case MXO_OSC_COSWAVE:  return o.coswave(f);
        case MXO_OSC_PHASOR:   return o.phasor(f);
        case MXO_OSC_SAW:      return o.saw(f);
        case MXO_OSC_SQUARE:   return o.square(f);
        case MXO_OSC_PULSE:    return o.pulse(f, duty);
        case MXO_OSC_IMPULSE:  return o.impulse(f);
        case MXO_OSC_TRIANGLE: return o.triangle(f);
        case MXO_OSC_PHASORBETWEEN: return o.phasorBetween(f, pstart, pend);
    }
    return 0.0;
}

}  // namespace

extern "C" {

const char* mxo_kind(void) { return "reference"; }

void* mxo_bank_create(const mxo_chain* chain, int32_t voices) {
    if (!chain || voices <= 0) return nullptr;
    maxiSettings::setup((size_t)chain->sample_rate, 2, 1024);
    RefBank* b = new RefBank();
    b->chain = *chain;
    b->V = voices;
    b->voices = (RefVoice*)calloc((size_t)voices, sizeof(RefVoice));
    for (int v = 0; v < voices; ++v) new (&b->voices[v]) RefVoice();
    b->delays = nullptr;
    if (chain->delay_on) {
        b->delays = (maxiDelayline**)calloc((size_t)voices, sizeof(maxiDelayline*));
        for (int v = 0; v < voices; ++v) {
            b->delays[v] = zeroed_new<maxiDelayline>();
            if (!b->delays[v]) { return nullptr; }
        }
    }
    /* defaults so that an unset parameter behaves like an untouched reference object */
    b->p[MXO_P_FREQ].assign(voices, 0.0);
    b->p[MXO_P_DUTY].assign(voices, 0.5);
    b->p[MXO_P_PHASOR_START].assign(voices, 0.0);
    b->p[MXO_P_PHASOR_END].assign(voices, 1.0);
    b->p[MXO_P_DELAY_SIZE].assign(voices, 1.0);
    b->p[MXO_P_DELAY_FEEDBACK].assign(voices, 0.0);
    b->p[MXO_P_PAN].assign(voices, 0.5);
    b->p[MXO_P_DELAY_POSITION].assign(voices, 0.0);
    b->p[MXO_P_ENV_ATTACK].assign(voices, 0.0);
    b->p[MXO_P_ENV_RELEASE].assign(voices, 0.0);
    b->p[MXO_P_ENV_HOLDTIME].assign(voices, 1.0);
    return b;
}

void mxo_bank_destroy(void* h) {
    RefBank* b = (RefBank*)h;
    if (!b) return;
    if (b->delays) {
        for (int v = 0; v < b->V; ++v) { if (b->delays[v]) { b->delays[v]->~maxiDelayline(); free(b->delays[v]); } }
        free(b->delays);
    }
    for (int v = 0; v < b->V; ++v) b->voices[v].~RefVoice();
    free(b->voices);
    delete b;
}

int32_t mxo_bank_set(void* h, int32_t id, const double* x) {
    RefBank* b = (RefBank*)h;
    if (!b || !x || id < 0 || id >= MXO_P_COUNT) return -1;
    maxiSettings::setup((size_t)b->chain.sample_rate, 2, 1024);
    b->p[id].assign(x, x + b->V);
    switch (id) {
        case MXO_P_PHASE: for (int v = 0; v < b->V; ++v) b->voices[v].osc.phaseReset(x[v]); break;
        case MXO_P_CUTOFF: case MXO_P_RESONANCE: case MXO_P_GAIN:
            if (b->chain.filt_kind == MXO_FILT_SVF) apply_svf(b);
            if (b->chain.filt_kind == MXO_FILT_BIQUAD) apply_biquad(b);
            break;
        case MXO_P_ENV_ATTACK:  for (int v = 0; v < b->V; ++v) b->voices[v].env.attack = x[v]; break;
        case MXO_P_ENV_DECAY:   for (int v = 0; v < b->V; ++v) b->voices[v].env.decay = x[v]; break;
        case MXO_P_ENV_SUSTAIN: for (int v = 0; v < b->V; ++v) b->voices[v].env.setSustain(x[v]); break;
        case MXO_P_ENV_RELEASE: for (int v = 0; v < b->V; ++v) b->voices[v].env.release = x[v]; break;
        case MXO_P_ENV_HOLDTIME:for (int v = 0; v < b->V; ++v) b->voices[v].env.holdtime = (long)x[v]; break;
        default: break;
    }
    return 0;
}

int32_t mxo_bank_get(void* h, int32_t id, double* x) {
    RefBank* b = (RefBank*)h;
    if (!b || !x) return -1;
    const int fk = b->chain.filt_kind;
    for (int v = 0; v < b->V; ++v) {
        RefVoice& r = b->voices[v];
        switch (id) {
            case MXO_P_PHASE: x[v] = r.osc.phase; break;
            case MXO_S_FILT_0: x[v] = fk == MXO_FILT_SVF ? r.svf.v0z : fk == MXO_FILT_BIQUAD ? r.bq.v[1] : r.filt.x; break;
            case MXO_S_FILT_1: x[v] = fk == MXO_FILT_SVF ? r.svf.v1  : fk == MXO_FILT_BIQUAD ? r.bq.v[2] : r.filt.y; break;
            case MXO_S_FILT_2: x[v] = fk == MXO_FILT_SVF ? r.svf.v2 : 0.0; break;
            case MXO_S_ENV_AMPLITUDE: x[v] = r.env.amplitude; break;
            case MXO_S_ENV_OUTPUT: x[v] = r.env.output; break;
            case MXO_S_ENV_HOLDCOUNT: x[v] = (double)r.env.holdcount; break;
            case MXO_S_ENV_FLAGS:
                x[v] = (double)((r.env.attackphase & 1) | (r.env.decayphase & 1) << 1 | (r.env.sustainphase & 1) << 2 |
                                (r.env.holdphase & 1) << 3 | (r.env.releasephase & 1) << 4);
                break;
            case MXO_S_DELAY_PHASE: x[v] = b->delays ? (double)b->delays[v]->phase : 0.0; break;
            case MXO_S_OSC_OUTPUT: x[v] = r.osc.output; break;
            default:
                if (id >= 0 && id < MXO_P_COUNT && !b->p[id].empty()) x[v] = b->p[id][v]; else return -1;
        }
    }
    return 0;
}

int32_t mxo_bank_process(void* h, int32_t nframes, const int32_t* trig_on, const int32_t* trig_off,
                         double* out, double* mix, int32_t first, int32_t count) {
    return mxo_bank_process_fm(h, nframes, nullptr, trig_on, trig_off, out, mix, first, count);
}

int32_t mxo_bank_process_fm(void* h, int32_t nframes, const double* freq_tv, const int32_t* trig_on, const int32_t* trig_off,
                            double* out, double* mix, int32_t first, int32_t count) {
    return mxo_bank_process_mod(h, nframes, freq_tv, nullptr, nullptr, trig_on, trig_off, out, mix, first, count);
}

int32_t mxo_bank_process_mod(void* h, int32_t nframes, const double* freq_tv, const double* cutoff_tv, const double* delay_size_tv,
                             const int32_t* trig_on, const int32_t* trig_off, double* out, double* mix, int32_t first, int32_t count) {
    RefBank* b = (RefBank*)h;
    if (!b || nframes < 0 || first < 0 || count < 0 || first + count > b->V) return -1;
    const mxo_chain& c = b->chain;
    const int V = b->V;
    const double* freq = b->p[MXO_P_FREQ].data();
    const double* duty = b->p[MXO_P_DUTY].data();
    const double* fc = b->p[MXO_P_CUTOFF].empty() ? nullptr : b->p[MXO_P_CUTOFF].data();
    const double* q = b->p[MXO_P_RESONANCE].empty() ? nullptr : b->p[MXO_P_RESONANCE].data();
    const double* dsize = b->p[MXO_P_DELAY_SIZE].data();
    const double* dfb = b->p[MXO_P_DELAY_FEEDBACK].data();
    const double* pan = b->p[MXO_P_PAN].data();
    if ((c.filt_kind == MXO_FILT_LORES || c.filt_kind == MXO_FILT_HIRES) && ((!fc && !cutoff_tv) || !q)) return -2;
    if (cutoff_tv && c.filt_kind == MXO_FILT_BIQUAD) return -3;
    if (delay_size_tv && !c.delay_on) return -3;
    std::vector<double> two(2, 0.0);
    for (int t = 0; t < nframes; ++t) {
        double m0 = 0.0, m1 = 0.0;
        for (int v = first; v < first + count; ++v) {
            RefVoice& r = b->voices[v];
            double x = run_osc(r.osc, c.osc_kind, freq_tv ? freq_tv[(size_t)t * V + v] : freq[v], duty[v],
                               b->p[MXO_P_PHASOR_START][v], b->p[MXO_P_PHASOR_END][v]);
            if (c.env_kind == MXO_ENV_ADSR) {
                int trig = (trig_on && trig_off && t >= trig_on[v] && t < trig_off[v]) ? 1 : 0;
                x = r.env.adsr(x, trig);
            } else if (c.env_kind == MXO_ENV_AR) {
                int trig = (trig_on && trig_off && t >= trig_on[v] && t < trig_off[v]) ? 1 : 0;
                x = r.env.ar(x, b->p[MXO_P_ENV_ATTACK][v], b->p[MXO_P_ENV_RELEASE][v], (long)b->p[MXO_P_ENV_HOLDTIME][v], trig);
            }
            switch (c.filt_kind) {
                case MXO_FILT_LORES: x = r.filt.lores(x, cutoff_tv ? cutoff_tv[(size_t)t * V + v] : fc[v], q[v]); break;
                case MXO_FILT_HIRES: x = r.filt.hires(x, cutoff_tv ? cutoff_tv[(size_t)t * V + v] : fc[v], q[v]); break;
                case MXO_FILT_SVF:
                    if (cutoff_tv) r.svf.setCutoff(cutoff_tv[(size_t)t * V + v]);     // the patch's per-sample setter call
                    x = r.svf.play(x, c.svf_mix[0], c.svf_mix[1], c.svf_mix[2], c.svf_mix[3]); break;
                case MXO_FILT_BIQUAD:x = r.bq.play(x); break;
                default: break;
            }
            const int dsz = delay_size_tv ? (int)delay_size_tv[(size_t)t * V + v] : (int)dsize[v];
            if (c.delay_on == 1) x = b->delays[v]->dl(x, dsz, dfb[v]);
            else if (c.delay_on == 2) x = b->delays[v]->dlFromPosition(x, dsz, dfb[v], (int)b->p[MXO_P_DELAY_POSITION][v]);
            if (out) out[(size_t)t * V + v] = x;
            if (mix) { r.mixer.stereo(x, two, pan[v]); m0 += two[0]; m1 += two[1]; }
        }
        if (mix) { mix[2 * t] = m0; mix[2 * t + 1] = m1; }
    }
    if (cutoff_tv && c.filt_kind == MXO_FILT_SVF && fc)      // the modulation lasts for this call
        for (int v = first; v < first + count; ++v) b->voices[v].svf.setCutoff(fc[v]);
    return 0;
}

int32_t mxo_bank_get_ring(void* h, int32_t v, double* dst, int32_t n) {
    RefBank* b = (RefBank*)h;
    if (!b || !b->delays || v < 0 || v >= b->V || n < 0 || n > 88200 * 8) return -1;
    memcpy(dst, b->delays[v]->memory, sizeof(double) * (size_t)n);
    return 0;
}

/* the setters themselves, run on a scratch object */
double mxo_env_attack_coeff(double ms, int32_t sr) {
    maxiSettings::setup((size_t)sr, 2, 1024); maxiEnv* e = zeroed_new<maxiEnv>(); e->setAttack(ms);
    double r = e->attack; free(e); return r;
}
double mxo_env_attack_ms_coeff(double ms, int32_t sr) {
    maxiSettings::setup((size_t)sr, 2, 1024); maxiEnv* e = zeroed_new<maxiEnv>(); e->setAttackMS(ms);
    double r = e->attack; free(e); return r;
}
double mxo_env_decay_coeff(double ms, int32_t sr) {
    maxiSettings::setup((size_t)sr, 2, 1024); maxiEnv* e = zeroed_new<maxiEnv>(); e->setDecay(ms);
    double r = e->decay; free(e); return r;
}

/* ------------------------------------------------------------------ STFT */
struct RefStft { int C, n, hop, bins; std::vector<maxiFFT*> f; };

void* mxo_stft_create(int32_t channels, int32_t fft_size, int32_t hop_size) {
    if (channels <= 0 || fft_size < 4 || (fft_size & (fft_size - 1)) || hop_size <= 0 || hop_size > fft_size) return nullptr;
    RefStft* s = new RefStft();
    s->C = channels; s->n = fft_size; s->hop = hop_size; s->bins = fft_size / 2;
    for (int c = 0; c < channels; ++c) { maxiFFT* f = new maxiFFT(); f->setup(fft_size, hop_size, fft_size); s->f.push_back(f); }
    return s;
}
void mxo_stft_destroy(void* h) { RefStft* s = (RefStft*)h; if (!s) return; for (auto* f : s->f) delete f; delete s; }

int32_t mxo_stft_process(void* h, const float* in, int32_t n, int32_t max_frames,
                         float* mags, float* phases, float* re, float* im) {
    RefStft* s = (RefStft*)h;
    if (!s || !in || n < 0) return -1;
    int frames = 0;
    for (int c = 0; c < s->C; ++c) {
        maxiFFT* f = s->f[c];
        int k = 0;
        for (int t = 0; t < n; ++t) {
            if (f->process(in[(size_t)c * n + t], maxiFFT::WITH_POLAR_CONVERSION)) {
                if (k >= max_frames) return -3;
                size_t o = ((size_t)c * max_frames + k) * s->bins;
                if (mags)   memcpy(mags + o, f->getMagnitudes().data(), sizeof(float) * s->bins);
                if (phases) memcpy(phases + o, f->getPhases().data(), sizeof(float) * s->bins);
                if (re)     memcpy(re + o, f->getReal(), sizeof(float) * s->bins);
                if (im)     memcpy(im + o, f->getImag(), sizeof(float) * s->bins);
                ++k;
            }
        }
        frames = k;
    }
    return frames;
}
int32_t mxo_stft_window(void* h, float* w) {
    RefStft* s = (RefStft*)h; if (!s || !w) return -1;
    memcpy(w, s->f[0]->window.data(), sizeof(float) * s->n); return 0;
}

/* the reference methods themselves, run on a maxiFFT whose (private) magnitudes vector is loaded with each frame */
int32_t mxo_spectral_features(const float* mags, int32_t n_frames, int32_t fft_size, int32_t sample_rate,
                              float* db, float* flatness, float* centroid) {
    if (!mags || n_frames < 0 || fft_size < 4 || (fft_size & (fft_size - 1))) return -1;
    maxiSettings::setup((size_t)sample_rate, 2, 1024);
    maxiFFT f;
    f.setup(fft_size, fft_size / 2, fft_size);
    const int bins = fft_size / 2;
    for (int i = 0; i < n_frames; ++i) {
        memcpy(f.magnitudes.data(), mags + (size_t)i * bins, sizeof(float) * bins);
        f.recalc = true;
        if (db) memcpy(db + (size_t)i * bins, f.getMagnitudesDB().data(), sizeof(float) * bins);
        if (flatness) flatness[i] = f.spectralFlatness();
        if (centroid) centroid[i] = f.spectralCentroid();
    }
    return 0;
}

/* ------------------------------------------------------------------ MFCC */
struct RefMfcc { int bins, filters, coeffs; maxiMFCC m; std::vector<float> spec; };

void* mxo_mfcc_create(int32_t num_bins, int32_t num_filters, int32_t num_coeffs,
                      double min_freq, double max_freq, int32_t sample_rate) {
    if (num_bins <= 0 || num_filters <= 0 || num_coeffs <= 0) return nullptr;
    maxiSettings::setup((size_t)sample_rate, 2, 1024);
    RefMfcc* m = new RefMfcc();
    m->bins = num_bins; m->filters = num_filters; m->coeffs = num_coeffs;
    m->m.setup(num_bins, num_filters, num_coeffs, min_freq, max_freq);
    /* calcMelFilterBank never writes filter 0 (loop starts at 1, src/libs/maxiMFCC.h:149): the column is
     * whatever malloc returned. Define it as zero (what a fresh mmap'ed block holds), SURVEY.md A13. */
    for (int bin = 0; bin < num_bins; ++bin) m->m.melFilters[(size_t)bin * num_filters] = 0.0;
    m->spec.resize(num_bins);
    return m;
}
void mxo_mfcc_destroy(void* h) { delete (RefMfcc*)h; }

int32_t mxo_mfcc_process(void* h, const float* mags, int32_t n, double* coeffs, double* melbands) {
    RefMfcc* m = (RefMfcc*)h;
    if (!m || !mags || !coeffs) return -1;
    for (int i = 0; i < n; ++i) {
        memcpy(m->spec.data(), mags + (size_t)i * m->bins, sizeof(float) * m->bins);
        std::vector<double>& r = m->m.mfcc(m->spec);
        memcpy(coeffs + (size_t)i * m->coeffs, r.data(), sizeof(double) * m->coeffs);
        if (melbands) memcpy(melbands + (size_t)i * m->filters, m->m.melBands, sizeof(double) * m->filters);
    }
    return 0;
}

/* ------------------------------------------------------------------ ISTFT */
struct RefIstft { int C, n, hop, bins; std::vector<maxiIFFT*> f; std::vector<float> mg, ph; };

void* mxo_istft_create(int32_t channels, int32_t fft_size, int32_t hop_size) {
    if (channels <= 0 || fft_size < 4 || (fft_size & (fft_size - 1)) || hop_size <= 0 || hop_size > fft_size) return nullptr;
    RefIstft* s = new RefIstft();
    s->C = channels; s->n = fft_size; s->hop = hop_size; s->bins = fft_size / 2;
    for (int c = 0; c < channels; ++c) { maxiIFFT* f = new maxiIFFT(); f->setup(fft_size, hop_size, fft_size); s->f.push_back(f); }
    s->mg.resize(s->bins); s->ph.resize(s->bins);
    return s;
}
void mxo_istft_destroy(void* h) { RefIstft* s = (RefIstft*)h; if (!s) return; for (auto* f : s->f) delete f; delete s; }

int32_t mxo_istft_process(void* h, const float* mags, const float* phases, int32_t frames, float* out) {
    RefIstft* s = (RefIstft*)h;
    if (!s || !mags || !phases || !out || frames < 0) return -1;
    for (int c = 0; c < s->C; ++c) {
        for (int f = 0; f < frames; ++f) {
            size_t o = ((size_t)c * frames + f) * s->bins;
            memcpy(s->mg.data(), mags + o, sizeof(float) * s->bins);
            memcpy(s->ph.data(), phases + o, sizeof(float) * s->bins);
            for (int t = 0; t < s->hop; ++t)
                out[(size_t)c * frames * s->hop + (size_t)f * s->hop + t] = s->f[c]->process(s->mg, s->ph, maxiIFFT::SPECTRUM);
        }
    }
    return 0;
}

}  // extern "C"
