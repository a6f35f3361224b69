// A feature extractor written the way the reference's examples write it -- ONE sample per call
// (cpp/openFrameworks/openFrameworksExamples/OSX/OF0.8.4MaximExtractorExample/src/testApp.cpp:audioRequested,
// js/script-processor-node/examples/20-analysis-MFCC.html): `if (fft.process(sample)) { oct.calculate(...); mfcc.mfcc(...); }` and
// `out = ifft.process(fft.getMagnitudes(), fft.getPhases())` on every sample -- against include/maximilian_b200.hpp. The per-sample
// signatures are host-side shims: the samples of a hop are collected and handed to the device together.
//
//   patch_featurex <in.bin> <out.bin> N
// in.bin:  N float32 samples of one channel.
// out.bin: int32 F, int32 nAverages, float32 fired[N] (1 where process() returned true), mags[F][512], octave averages[F][nAverages],
//          peaks[F][nAverages], float64 bark specific[F][24], relative[F][24], total[F], mfcc[F][13], float32 resynth[N].
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "maximilian_b200.hpp"

template <class T> static void put(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
    if (argc != 4) { fprintf(stderr, "usage: patch_featurex in.bin out.bin N\n"); return 2; }
    const int N = atoi(argv[3]);
    const int fftSize = 1024, hopSize = 512, ncoef = 13;
    vector<float> x((size_t)N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(x.data(), sizeof(float), x.size(), fi) != x.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(fi);
    try {
        maxiSettings::setup(48000, 2, 512);
        maxiFFT mfft;
        maxiIFFT ifft;
        maxiMFCC mfcc;
        maxiFFTOctaveAnalyzer oct;
        maxiBark bark;
        mfft.setup(fftSize, hopSize, fftSize);
        ifft.setup(fftSize, hopSize, fftSize);
        mfcc.setup(fftSize / 2, 42, ncoef, 20, 20000);
        oct.setup(48000, fftSize / 2, 3);
        bark.setup(48000, fftSize);
        mfft.attach(oct); mfft.attach(bark);
        oct.peakHoldTime = 2; oct.peakDecayRate = 0.8f; oct.linearEQIntercept = 0.9f; oct.linearEQSlope = 0.01f;     // public members, as a patch sets them

        vector<float> fired((size_t)N, 0.f), mags, avs, pks, y((size_t)N);
        vector<double> spec, rel, tot, co;
        for (int i = 0; i < N; ++i) {
            if (mfft.process(x[(size_t)i])) {
                fired[(size_t)i] = 1.f;
                vector<float>& m = mfft.getMagnitudes();
                mags.insert(mags.end(), m.begin(), m.end());
                oct.calculate(&m[0]);
                avs.insert(avs.end(), oct.averages, oct.averages + oct.nAverages);
                pks.insert(pks.end(), oct.peaks, oct.peaks + oct.nAverages);
                const double* s = bark.specificLoudness(&m[0]); spec.insert(spec.end(), s, s + 24);
                const double* r = bark.relativeLoudness(&m[0]); rel.insert(rel.end(), r, r + 24);
                tot.push_back(bark.totalLoudness(&m[0])[0]);
                vector<double>& c = mfcc.mfcc(m);
                co.insert(co.end(), c.begin(), c.begin() + ncoef);
            }
            y[(size_t)i] = ifft.process(mfft.getMagnitudes(), mfft.getPhases());
        }
        const int32_t F = (int32_t)tot.size(), nA = oct.nAverages;
        FILE* fo = fopen(argv[2], "wb");
        fwrite(&F, sizeof(F), 1, fo); fwrite(&nA, sizeof(nA), 1, fo);
        put(fo, fired); put(fo, mags); put(fo, avs); put(fo, pks); put(fo, spec); put(fo, rel); put(fo, tot); put(fo, co); put(fo, y);
        fclose(fo);
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
