// A reference-style patch written against include/maximilian_b200.hpp -- the block-rate twin of
// cpp/commandline/maximilian_examples/15.polysynth/main.cpp + tests/svftest/svftest.cpp:
//   every voice: saw -> ADSR -> SVF low-pass -> delay line -> equal-power stereo pan, summed into the output bus,
// driven through maxiRouting() exactly as RtAudio would drive routing() in cpp/commandline/player.cpp.
//
//   patch_poly <params.bin> <out.bin> V B NBLOCKS
// params.bin: 12 arrays of V doubles (freq, phase, cutoff, res, attackMS, decayMS, sustain, releaseMS, size, feedback, pan, unused)
//             followed by NBLOCKS x (V int32 on, V int32 off) gates.
// out.bin:    NBLOCKS x B x 2 doubles (the interleaved stereo buffer the audio callback filled),
//             then the per-voice samples of one extra block [B][V] rendered directly.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "maximilian_b200.hpp"

static int V, B, NB;
static std::vector<double> freq, phase0, cutoff, res, attackMS, decayMS, sustain, releaseMS, dsize, feedback, pan;
static std::vector<std::vector<int32_t>> gate_on, gate_off;
static int block_index = 0;

static maxiVoices* voices;
static maxiOsc* osc;
static maxiEnv* env;
static maxiSVF* svf;
static maxiDelayline* delay;
static maxiMix* mixer;

void setup() {
    maxiSettings::setup(48000, 2, B);
    osc->phaseReset(phase0);
    env->setAttack(attackMS); env->setDecay(decayMS); env->setSustain(sustain); env->setRelease(releaseMS);
    svf->setCutoff(cutoff); svf->setResonance(res);
}

// the block-rate play(): same calls as a reference play(), once per block for all voices
void play(maxiVoices& v) {
    maxiSignal w = osc->saw(freq);
    w = env->adsr(w, maxiGate(gate_on[block_index], gate_off[block_index]));
    w = svf->play(w, 1, 0, 0, 0);
    w = delay->dl(w, dsize, feedback);
    mixer->stereo(w, v.bus(), pan);
}

static void rd(FILE* f, std::vector<double>& a) { a.resize(V); if (fread(a.data(), sizeof(double), V, f) != (size_t)V) exit(3); }

int main(int argc, char** argv) {
    if (argc != 6) { fprintf(stderr, "usage: patch_poly params.bin out.bin V B NBLOCKS\n"); return 2; }
    V = atoi(argv[3]); B = atoi(argv[4]); NB = atoi(argv[5]);
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    std::vector<double> unused;
    rd(f, freq); rd(f, phase0); rd(f, cutoff); rd(f, res); rd(f, attackMS); rd(f, decayMS); rd(f, sustain); rd(f, releaseMS);
    rd(f, dsize); rd(f, feedback); rd(f, pan); rd(f, unused);
    gate_on.resize(NB + 1); gate_off.resize(NB + 1);
    for (int k = 0; k <= NB; ++k) {
        gate_on[k].resize(V); gate_off[k].resize(V);
        if (fread(gate_on[k].data(), 4, V, f) != (size_t)V || fread(gate_off[k].data(), 4, V, f) != (size_t)V) return 3;
    }
    fclose(f);
    try {
        maxiVoices vs(V);
        maxiOsc o(vs); maxiEnv e(vs); maxiSVF s(vs); maxiDelayline d(vs, 512); maxiMix m(vs);
        voices = &vs; osc = &o; env = &e; svf = &s; delay = &d; mixer = &m;
        maxiSettings::setup(48000, 2, B);
        setup();
        std::vector<double> buffer((size_t)B * 2);
        FILE* g = fopen(argv[2], "wb");
        for (block_index = 0; block_index < NB; ++block_index) {
            maxiRouting(buffer.data(), nullptr, (unsigned)B, 0.0, 0, &vs);       // what the audio driver would call
            fwrite(buffer.data(), sizeof(double), buffer.size(), g);
        }
        std::vector<double> per_voice((size_t)B * V);
        play(vs);
        vs.render(B, per_voice.data(), buffer.data());
        fwrite(per_voice.data(), sizeof(double), per_voice.size(), g);
        fclose(g);
    } catch (const maxiError& err) {
        fprintf(stderr, "maxiError %d: %s\n", err.code, err.what());
        return 1;
    }
    return 0;
}
