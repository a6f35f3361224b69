// An analysis/resynthesis patch written against include/maximilian_b200.hpp -- the block-rate twin of the reference's
// feature-extractor examples (maxiFFT -> magnitudes / dB / flatness / centroid -> maxiMFCC, and maxiIFFT back to samples;
// cpp/openFrameworks/openFrameworksExamples/OSX/OF0.8.4MaximExtractorExample/src/testApp.cpp, tests/mfcctest).
//
//   patch_spectral <in.bin> <out.bin> C N
// in.bin:  C x N float32 samples, planar. The stream is fed in three ragged chunks, like an audio callback would.
// out.bin: int32 F (frames per channel), then float32 mags[C][F][512], phases[C][F][512], dB[C][F][512], flatness[C][F],
//          centroid[C][F], float64 mfcc[C][F][13], float32 resynth[C][F*512].
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "maximilian_b200.hpp"

template <class T> static void put(FILE* f, const std::vector<T>& v) { fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
    if (argc != 5) { fprintf(stderr, "usage: patch_spectral in.bin out.bin C N\n"); return 2; }
    const int C = atoi(argv[3]), N = atoi(argv[4]);
    const int fftSize = 1024, hop = 512, bins = 512, ncoef = 13;
    std::vector<float> x((size_t)C * N);
    FILE* fi = fopen(argv[1], "rb");
    if (!fi || fread(x.data(), sizeof(float), x.size(), fi) != x.size()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(fi);
    try {
        maxiSettings::setup(48000, 2, 512);
        maxiFFT fft(C); fft.setup(fftSize, hop, fftSize);
        maxiMFCC mfcc; mfcc.setup(bins, 42, ncoef, 20, 20000);
        maxiIFFT ifft(C); ifft.setup(fftSize, hop, fftSize);

        std::vector<float> mags, phases, db, flat, cent;          // [C][F][...] gathered over the chunks
        std::vector<std::vector<float>> cm(C), cp(C), cd(C), cf(C), cc(C);
        const int cuts[4] = {0, N / 3 + 7, N / 3 + 8, N};
        for (int k = 0; k < 3; ++k) {
            const int n = cuts[k + 1] - cuts[k];
            std::vector<float> chunk((size_t)C * n);
            for (int c = 0; c < C; ++c) for (int i = 0; i < n; ++i) chunk[(size_t)c * n + i] = x[(size_t)c * N + cuts[k] + i];
            if (!fft.process(chunk.data(), n)) continue;
            const int F = fft.frames(), S = fft.frameStride();
            for (int c = 0; c < C; ++c) {
                for (int f = 0; f < F; ++f) {
                    const size_t o = ((size_t)c * S + f) * bins;
                    cm[c].insert(cm[c].end(), fft.getMagnitudes().begin() + o, fft.getMagnitudes().begin() + o + bins);
                    cp[c].insert(cp[c].end(), fft.getPhases().begin() + o, fft.getPhases().begin() + o + bins);
                    cd[c].insert(cd[c].end(), fft.magsToDB().begin() + o, fft.magsToDB().begin() + o + bins);
                    cf[c].push_back(fft.spectralFlatness()[(size_t)c * S + f]);
                    cc[c].push_back(fft.spectralCentroid()[(size_t)c * S + f]);
                }
            }
        }
        const int32_t F = (int32_t)cf[0].size();
        for (int c = 0; c < C; ++c) { mags.insert(mags.end(), cm[c].begin(), cm[c].end()); phases.insert(phases.end(), cp[c].begin(), cp[c].end());
                                      db.insert(db.end(), cd[c].begin(), cd[c].end()); flat.insert(flat.end(), cf[c].begin(), cf[c].end());
                                      cent.insert(cent.end(), cc[c].begin(), cc[c].end()); }
        std::vector<double> co = mfcc.mfcc(mags);                  // [C*F][13]
        std::vector<float> y = ifft.process(mags, phases, F);      // [C][F*hop]
        FILE* fo = fopen(argv[2], "wb");
        fwrite(&F, sizeof(F), 1, fo);
        put(fo, mags); put(fo, phases); put(fo, db); put(fo, flat); put(fo, cent); put(fo, co); put(fo, y);
        fclose(fo);
    } catch (const std::exception& e) { fprintf(stderr, "%s\n", e.what()); return 1; }
    return 0;
}
