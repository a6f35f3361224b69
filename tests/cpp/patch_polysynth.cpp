// cpp/commandline/maximilian_examples/15.polysynth/main.cpp, written against include/maximilian_b200.hpp: the same objects, the
// same expressions in the same order -- `maxiOsc VCO1[6]` becomes one `maxiOsc VCO1(voices)` whose calls stand for all six voices,
// the per-voice arrays become per-voice vectors, and the per-SAMPLE control code of the original play() (the metronome that writes
// ADSR[voice].trigger) runs once per block over all of its samples and hands the triggers over as a stream. Two oscillators summed
// into the filter, an LFO on the second oscillator's frequency and on the cutoff, the envelope applied AFTER the filter: a graph the
// fused bank kernels cannot express -- maxiVoices runs it as a voice patch: the kernel the library generates and compiles for this graph
// (MXB_PATCH_MODE=interpret: the interpreting kernel).
//
//   patch_polysynth <tables.bin> <out.bin> NBLOCKS B      tables.bin: sineBuffer[514] ++ transition[1001] ++ sine_before (doubles)
// out.bin: the trigger stream [NBLOCKS][B][6] (doubles), then the interleaved stereo output [NBLOCKS][B][2] the audio callback filled.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "maximilian_b200.hpp"

// six voices: two VCOs and an LFO each, one filter, one envelope
maxiVoices voices(6);
maxiOsc VCO1(voices), VCO2(voices), LFO1(voices), LFO2(voices);
maxiFilter VCF(voices);
maxiEnv ADSR(voices);

double timerPhase = 0;               // phase of the 8 Hz metronome (the original's timer.phasor(8))
int currentCount, lastCount, voice = 0;   // beat detection and the round-robin voice index

vector<double> pitch = {1, 2, 3, 4, 5, 6}, f1(6), f2(6);
vector<double> trigger;              // what the original writes into ADSR[i].trigger, for every sample of the block: [frame][voice]

void setup() {
    ADSR.setAttack(0);
    ADSR.setDecay(200);
    ADSR.setSustain(0.2);
    ADSR.setRelease(2000);
    for (int i = 0; i < 6; i++) { f1[i] = 55 * pitch[i]; f2[i] = 110 * pitch[i]; }
}

// The per-sample CONTROL half of the original play(): on every metronome tick the next voice is triggered for one sample.
void control(int nFrames) {
    trigger.assign((size_t)nFrames * 6, 0.0);
    for (int t = 0; t < nFrames; t++) {
        currentCount = (int)timerPhase;                       // maxiOsc::phasor(8), src/maximilian.cpp:285-291
        if (timerPhase >= 1.0) timerPhase -= 1.0;
        timerPhase += (1. / (maxiSettings::sampleRate / (8.)));
        if (lastCount != currentCount) {                      // a tick
            if (voice == 6) {
                voice = 0;
            }
            trigger[(size_t)t * 6 + voice] = 1;
            voice++;
        }
    }
}

void play(maxiVoices& v) {
    maxiSignal ADSRout = ADSR.adsr(1., maxiStream(trigger));      // envelope of a constant 1
    maxiSignal LFO1out = LFO1.sinebuf(0.2);                        // 0.2 Hz LFO
    maxiSignal VCO1out = VCO1.pulse(f1, 0.6);                      // 55 Hz * pitch
    maxiSignal VCO2out = VCO2.pulse(f2 + LFO1out, 0.2);            // 110 Hz * pitch, LFO on the frequency
    maxiSignal VCFout = VCF.lores((VCO1out + VCO2out) * 0.5, 250 + ((pitch + LFO1out) * 1000), 10);   // both VCOs into the VCF, LFO on the cutoff
    v.sum(VCFout * ADSRout / 6);                               // envelope AFTER the filter; mix += ... over the voices
}

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: patch_polysynth tables.bin out.bin NBLOCKS B\n"); return 2; }
    const int NB = atoi(argv[3]), B = atoi(argv[4]);
    vector<double> tab(514 + 1001 + 1);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(tab.data(), sizeof(double), tab.size(), f) != tab.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    fclose(f);
    try {
        maxiSettings::setup(44100, 2, B);
        voices.setTables(tab.data(), tab.data() + 514, tab[514 + 1001]);
        setup();
        vector<double> triggers, output, bus((size_t)B * 2);
        for (int blk = 0; blk < NB; ++blk) {
            control(B);
            maxiRouting(bus.data(), nullptr, (unsigned)B, 0.0, 0, &voices);      // what the audio driver would call
            for (int t = 0; t < B; ++t) {
                const double mix = bus[(size_t)t * 2];
                output.push_back(mix * 0.5);
                output.push_back(mix * 0.5);
            }
            triggers.insert(triggers.end(), trigger.begin(), trigger.end());
        }
        if (voices.fused()) { fprintf(stderr, "expected a voice patch, got the fused bank\n"); return 1; }
        FILE* g = fopen(argv[2], "wb");
        fwrite(triggers.data(), sizeof(double), triggers.size(), g);
        fwrite(output.data(), sizeof(double), output.size(), g);
        fclose(g);
    } catch (const maxiError& err) {
        fprintf(stderr, "maxiError %d: %s\n", err.code, err.what());
        return 1;
    }
    return 0;
}
