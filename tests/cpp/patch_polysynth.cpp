// cpp/commandline/maximilian_examples/15.polysynth/main.cpp, written against include/maximilian_b200.hpp: the same objects, the
// same expressions in the same order -- `maxiOsc VCO1[6]` becomes one `maxiOsc VCO1(voices)` whose calls stand for all six voices,
// the per-voice arrays become per-voice vectors, and the per-SAMPLE control code of the original play() (the metronome that writes
// ADSR[voice].trigger) runs once per block over all of its samples and hands the triggers over as a stream. Two oscillators summed
// into the filter, an LFO on the second oscillator's frequency and on the cutoff, the envelope applied AFTER the filter: a graph the
// fused bank kernels cannot express -- maxiVoices runs it on the patch interpreter.
//
//   patch_polysynth <tables.bin> <out.bin> NBLOCKS B      tables.bin: sineBuffer[514] ++ transition[1001] ++ sine_before (doubles)
// out.bin: the trigger stream [NBLOCKS][B][6] (doubles), then the interleaved stereo output [NBLOCKS][B][2] the audio callback filled.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "maximilian_b200.hpp"

//This shows how to use maximilian to build a polyphonic synth.

//These are the synthesiser bits
maxiVoices voices(6);
maxiOsc VCO1(voices), VCO2(voices), LFO1(voices), LFO2(voices);
maxiFilter VCF(voices);
maxiEnv ADSR(voices);

//This is a bunch of control signals so that we can hear something

double timerPhase = 0;//this is the metronome (the per-sample control code of the original: timer.phasor(8))
int currentCount, lastCount, voice = 0;//these values are used to check if we have a new beat this sample

//and these are some variables we can use to pass stuff around

vector<double> pitch = {1, 2, 3, 4, 5, 6}, f1(6), f2(6);
vector<double> trigger;//ADSR[i].trigger for every sample of the block: [frame][voice]

void setup() {//some inits
    ADSR.setAttack(0);
    ADSR.setDecay(200);
    ADSR.setSustain(0.2);
    ADSR.setRelease(2000);
    for (int i = 0; i < 6; i++) { f1[i] = 55 * pitch[i]; f2[i] = 110 * pitch[i]; }
}

//the control half of the original play(): a metronome that ticks 8 times a second; every tick triggers the next voice for one sample
void control(int nFrames) {
    trigger.assign((size_t)nFrames * 6, 0.0);
    for (int t = 0; t < nFrames; t++) {
        currentCount = (int)timerPhase;//maxiOsc::phasor(8), src/maximilian.cpp:285-291
        if (timerPhase >= 1.0) timerPhase -= 1.0;
        timerPhase += (1. / (maxiSettings::sampleRate / (8.)));
        if (lastCount != currentCount) {//if we have a new timer int this sample, play the sound
            if (voice == 6) {
                voice = 0;
            }
            trigger[(size_t)t * 6 + voice] = 1;//trigger the envelope from the start
            voice++;
        }
    }
}

void play(maxiVoices& v) {
    //and this is where we build the synth
    maxiSignal ADSRout = ADSR.adsr(1., maxiStream(trigger));//our ADSR env is passed a constant signal of 1 to generate the transient.
    maxiSignal LFO1out = LFO1.sinebuf(0.2);//this lfo is a sinewave at 0.2 hz
    maxiSignal VCO1out = VCO1.pulse(f1, 0.6);//here's VCO1. it's a pulse wave at 55 hz, with a pulse width of 0.6
    maxiSignal VCO2out = VCO2.pulse(f2 + LFO1out, 0.2);//here's VCO2. it's a pulse wave at 110hz with LFO modulation on the frequency, and width of 0.2
    maxiSignal VCFout = VCF.lores((VCO1out + VCO2out) * 0.5, 250 + ((pitch + LFO1out) * 1000), 10);//now we stick the VCO's into the VCF, using the ADSR as the filter cutoff
    v.sum(VCFout * ADSRout / 6);//finally we add the ADSR as an amplitude modulator
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
                output.push_back(mix * 0.5);//left channel
                output.push_back(mix * 0.5);//right channel
            }
            triggers.insert(triggers.end(), trigger.begin(), trigger.end());
        }
        if (voices.fused()) { fprintf(stderr, "expected the interpreter, got the fused bank\n"); return 1; }
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
