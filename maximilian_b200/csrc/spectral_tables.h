// Host-side generation of every data-independent sequence of the reference's float FFT and MFCC, with the
// reference's own expressions (types, literals and evaluation order included), so that the device only
// replays per-butterfly arithmetic. Compiled with -ffp-contract=off.
#pragma once

#include <math.h>
#include <stdint.h>

#include <vector>

namespace mxb {

#ifndef MXB_FFT_M_PI
#define MXB_FFT_M_PI 3.14159265358979323846 /* src/libs/fft.h:36-38 */
#endif

struct f2 { float x, y; };

// ReverseBits, src/libs/fft.cpp:75-85
inline int reverse_bits(int index, int bits) {
    int rev = 0;
    for (int i = 0; i < bits; i++) { rev = (rev << 1) | (index & 1); index >>= 1; }
    return rev;
}
inline int ilog2(int n) { int b = 0; while ((1 << b) < n) ++b; return b; }

// Twiddles of FFT(), src/libs/fft.cpp:155-182. The recurrence ar0 = w*ar1 - ar2 restarts at every block, so
// the pair (ar0, ai0) used by butterfly n of the stage with BlockEnd = be depends on (be, n) only. Entry for
// (be, n) is stored at index (be - 1) + n: stages back to back, NumSamples - 1 entries in all.
inline std::vector<f2> fft_twiddles(int NumSamples, bool InverseTransform) {
    std::vector<f2> tw((size_t)NumSamples - 1);
    double angle_numerator = 2.0 * MXB_FFT_M_PI;
    if (InverseTransform) angle_numerator = -angle_numerator;
    int BlockEnd = 1;
    for (int BlockSize = 2; BlockSize <= NumSamples; BlockSize <<= 1) {
        double delta_angle = angle_numerator / (double)BlockSize;
        float sm2 = sin(-2 * delta_angle);
        float sm1 = sin(-delta_angle);
        float cm2 = cos(-2 * delta_angle);
        float cm1 = cos(-delta_angle);
        float w = 2 * cm1;
        float ar0, ar1, ar2, ai0, ai1, ai2;
        ar2 = cm2; ar1 = cm1;
        ai2 = sm2; ai1 = sm1;
        for (int n = 0; n < BlockEnd; n++) {
            ar0 = w * ar1 - ar2; ar2 = ar1; ar1 = ar0;
            ai0 = w * ai1 - ai2; ai2 = ai1; ai1 = ai0;
            tw[(size_t)(BlockEnd - 1) + n].x = ar0;
            tw[(size_t)(BlockEnd - 1) + n].y = ai0;
        }
        BlockEnd = BlockSize;
    }
    return tw;
}

// (wr, wi) of RealFFT()'s untangling loop, src/libs/fft.cpp:243-271, as seen by iteration i (entry i; entry 0 unused)
inline std::vector<f2> realfft_untangle(int NumSamples) {
    int Half = NumSamples / 2;
    std::vector<f2> uw((size_t)(Half / 2 > 0 ? Half / 2 : 1));
    float theta = MXB_FFT_M_PI / Half;
    float wtemp = float(sin(0.5 * theta));
    float wpr = -2.0 * wtemp * wtemp;
    float wpi = float(sin(theta));
    float wr = 1.0 + wpr;
    float wi = wpi;
    uw[0].x = 0; uw[0].y = 0;
    for (int i = 1; i < Half / 2; i++) {
        uw[i].x = wr; uw[i].y = wi;
        wr = (wtemp = wr) * wpr - wi * wpi + wr;
        wi = wi * wpr + wtemp * wpi + wi;
    }
    return uw;
}

// fft::genWindow(3, ...), src/libs/fft.cpp:409-413
inline std::vector<float> hann_window(int NumSamples) {
    std::vector<float> w((size_t)NumSamples);
    for (int i = 0; i < NumSamples; i++) w[i] = 0.50 - 0.50 * cos(2 * MXB_FFT_M_PI * i / (NumSamples - 1));
    return w;
}

// ---- maxiMFCC tables, src/libs/maxiMFCC.h:30-38, 118-203 ----
inline double hzToMel(double hz) { return 2595.0 * (log10(hz / 700.0 + 1.0)); }
inline double melToHz(double mel) { return 700.0 * (pow(10, mel / 2595.0) - 1.0); }

struct MfccTables {
    std::vector<double> melFilters;   // dense, idx = filter + bin*numFilters (filter 0 column: zero, never written by the reference)
    std::vector<double> dct;          // idx = i + j*numCoeffs
    // band structure of the dense matrix: filter f is non-zero on bins [lo[f], lo[f]+cnt[f]), weights at off[f]
    std::vector<int> lo, cnt, off;
    std::vector<double> w;
};

inline MfccTables mfcc_tables(unsigned numBins, unsigned numFilters, unsigned numCoeffs, double minFreq, double maxFreq, unsigned sampleRateU) {
    MfccTables t;
    t.melFilters.assign((size_t)numFilters * numBins, 0.0);
    t.dct.assign((size_t)numCoeffs * numFilters, 0.0);
    {   // calcMelFilterBank(sampleRate, numBins), :118-182
        const double sampleRate = sampleRateU;
        double mel, dMel, maxMel, minMel, nyquist, binFreq, start, thisF, nextF, prevF;
        nyquist = sampleRate / 2;
        if (maxFreq > nyquist) maxFreq = nyquist;
        maxMel = hzToMel(maxFreq);
        minMel = hzToMel(minFreq);
        dMel = (maxMel - minMel) / (numFilters + 2 - 1);
        std::vector<double> filtPos((size_t)numFilters + 2);
        mel = minMel;
        for (unsigned i = 0; i < numFilters + 2; i++) { filtPos[i] = melToHz(mel); mel += dMel; }
        for (int filter = 1; filter < (int)numFilters; filter++) {
            for (int bin = 0; bin < (int)numBins; bin++) {
                binFreq = (double)sampleRate / (double)(int)numBins * (double)bin;   // sr/numBins (not sr/fftSize), as the reference has it
                thisF = filtPos[filter]; nextF = filtPos[filter + 1]; prevF = filtPos[filter - 1];
                size_t idx = (size_t)filter + ((size_t)bin * numFilters);
                if (binFreq > nextF || binFreq < prevF) {
                    t.melFilters[idx] = 0;
                } else {
                    double height = 2.0 / (nextF - prevF);
                    if (binFreq < thisF) {
                        start = prevF;
                        t.melFilters[idx] = (binFreq - start) * (height / (thisF - start));
                    } else {
                        t.melFilters[idx] = height + ((binFreq - thisF) * (-height / (nextF - thisF)));
                    }
                }
            }
        }
    }
    {   // createDCTCoeffs, :183-203
        double k = 3.14159265358979323846 / numFilters;
        double w1 = 1.0 / (sqrt((double)numFilters));
        double w2 = sqrt(2.0 / numFilters);
        for (int i = 0; i < (int)numCoeffs; i++)
            for (int j = 0; j < (int)numFilters; j++) {
                size_t idx = (size_t)i + ((size_t)j * numCoeffs);
                if (i == 0) t.dct[idx] = w1 * cos(k * (i + 1) * (j + 0.5));
                else t.dct[idx] = w2 * cos(k * (i + 1) * (j + 0.5));
            }
    }
    // band structure: melBands[f] = sum over ALL bins in ascending order in the reference; terms with a zero
    // weight add +0.0 to a non-negative partial sum and change nothing, so summing the non-zero run in the
    // same ascending order is bit-identical (magnitudes are finite and >= 0).
    t.lo.assign(numFilters, 0); t.cnt.assign(numFilters, 0); t.off.assign(numFilters, 0);
    for (unsigned f = 0; f < numFilters; ++f) {
        int first = -1, last = -1;
        for (unsigned bin = 0; bin < numBins; ++bin)
            if (t.melFilters[(size_t)f + (size_t)bin * numFilters] != 0.0) { if (first < 0) first = (int)bin; last = (int)bin; }
        t.off[f] = (int)t.w.size();
        if (first >= 0) {
            t.lo[f] = first; t.cnt[f] = last - first + 1;
            for (int bin = first; bin <= last; ++bin) t.w.push_back(t.melFilters[(size_t)f + (size_t)bin * numFilters]);
        }
    }
    return t;
}

// maxiFFTOctaveAnalyzer::setup, src/libs/maxiFFT.cpp:201-256 (float arithmetic; pow on floats is powf there): the bin -> averaging
// band map, turned into the RUNS of bins whose running sum calculate() (:264-287) stores: a run ends ON the first bin whose band
// differs from the previous bin's, and its average goes to every band in [previous band, this band).
struct OctaveMap { int n_avg; std::vector<int> runs; };     // runs: first bin, last bin (inclusive), first band, one past the last band
inline OctaveMap octave_map(float samplingRate, int nBandsInTheFFT, int nAveragesPerOctave) {
    OctaveMap m;
    const int nSpectrum = nBandsInTheFFT;
    const float spectrumFrequencySpan = (samplingRate / 2.0f) / (float)(nSpectrum);
    if (nAveragesPerOctave == 0) nAveragesPerOctave = 1;
    const float averageFrequencyIncrement = powf(2.0f, 1.0f / (float)(nAveragesPerOctave));
    const float firstOctaveFrequency = 55.0f;
    std::vector<int> spe2avg((size_t)nSpectrum);
    int avgidx = 0;
    float averageFreq = firstOctaveFrequency;
    float spectrumFreq = spectrumFrequencySpan;
    for (int speidx = 0; speidx < nSpectrum; speidx++) {
        while (spectrumFreq > averageFreq) { avgidx++; averageFreq *= averageFrequencyIncrement; }
        spe2avg[(size_t)speidx] = avgidx;
        spectrumFreq += spectrumFrequencySpan;
    }
    m.n_avg = avgidx;
    int last_avgidx = 0, first = 0;
    for (int speidx = 0; speidx < nSpectrum; speidx++) {
        const int a = spe2avg[(size_t)speidx];
        if (a != last_avgidx) {
            m.runs.push_back(first); m.runs.push_back(speidx); m.runs.push_back(last_avgidx); m.runs.push_back(a < m.n_avg ? a : m.n_avg);
            first = speidx + 1;
        }
        last_avgidx = a;
    }
    if (first < nSpectrum && last_avgidx < m.n_avg) {       // "the last average was probably not calculated..."
        m.runs.push_back(first); m.runs.push_back(nSpectrum - 1); m.runs.push_back(last_avgidx); m.runs.push_back(last_avgidx + 1);
    }
    return m;
}

// maxiBarkScaleAnalyser::setup, src/libs/maxiBark.h:40-61: bbLimits[0..24]. binToHz is unsigned integer arithmetic returned as a
// double (:30-32), currentBandEnd an int (:52); the reference writes bbLimits[24] into the member behind its int[24].
inline std::vector<int> bark_limits(unsigned int sR, unsigned int bS) {
    const unsigned int specSize = bS / 2;
    const int NUM_BARK_BANDS = 24;
    std::vector<double> barkScale((size_t)specSize);
    std::vector<int> bbLimits(32, 0);
    for (unsigned int i = 0; i < specSize; i++) {
        const double hz = (double)(i * sR / bS);
        barkScale[i] = 13.0 * atan(hz / 1315.8) + 3.5 * atan(pow((hz / 7518.0), 2));
    }
    bbLimits[0] = 0;
    int currentBandEnd = barkScale[specSize - 1] / NUM_BARK_BANDS;
    int currentBand = 1;
    for (unsigned int i = 0; i < specSize; i++) {
        while (barkScale[i] > currentBandEnd) {
            if (currentBand < 32) bbLimits[(size_t)currentBand] = (int)i;
            currentBand++;
            currentBandEnd = currentBand * barkScale[specSize - 1] / NUM_BARK_BANDS;
        }
    }
    bbLimits[NUM_BARK_BANDS] = (int)specSize - 1;
    bbLimits.resize(25);
    return bbLimits;
}

}  // namespace mxb
