// Coefficient design of the reference's filters, one voice at a time: the reference's own formulas, callable from the host
// (glibc, once per parameter change: bit-identical coefficients, bank.cu) and from the device (libdevice, for stages that
// design their coefficients on every sample: patch.cu).
#pragma once
#include <math.h>

#include "common.cuh"

namespace mxb {

// maxiFilter::lores / hires coefficient part, src/maximilian.cpp:456-462 (identical in hires :472-478)
__host__ __device__ inline void design_lores_one(double cutoff, double resonance, const double sr, double& c, double& r) {
    if (cutoff < 10) cutoff = 10;
    if (cutoff > sr) cutoff = sr;
    if (resonance < 1.) resonance = 1.;
    const double z = cos(6.283185307179586476925286766559 * cutoff / sr);
    c = 2 - 2 * z;
    r = (sqrt(2.0) * sqrt(-pow((z - 1.0), 3.0)) + resonance * (z - 1)) / (resonance * (z - 1));
}

// maxiSVF::setParams, src/maximilian.h:1322-1334: cf = g1, g2, g3, g4, k
__host__ __device__ inline void design_svf_one(const double freq, const double res, const double sr, double* cf) {
    const double g = tan(3.1415926535897932384626433832795 * freq / sr);
    const double damping = res == 0 ? 0 : 1.0 / res;
    const double k = damping;
    const double ginv = g / (1.0 + g * (g + k));
    cf[0] = ginv; cf[1] = 2.0 * (g + k) * ginv; cf[2] = g * ginv; cf[3] = 2.0 * ginv; cf[4] = k;
}

// maxiFilter::bandpass coefficient part, src/maximilian.cpp:488-495: inputs[0..2]
__host__ __device__ inline void design_bandpass(double cutoff, double resonance, const double sr, double& c0, double& c1, double& c2) {
    if (cutoff > (sr * 0.5)) cutoff = (sr * 0.5);
    if (resonance >= 1.) resonance = 0.999999;
    const double z = cos(6.283185307179586476925286766559 * cutoff / sr);
    c0 = (1 - resonance) * (sqrt(resonance * (resonance - 4.0 * pow(z, 2.0) + 2.0) + 1));
    c1 = 2 * z * resonance;
    c2 = pow((resonance * -1), 2.0);
}

// maxiBiquad::set, src/maximilian.h:1375-1479: cf = a0, a1, a2, b1, b2
__host__ __device__ inline void design_biquad_one(const int type, const double cutoff, const double Q, const double peakGain, const double sr, double* cf) {
    const double SQRT2 = sqrt(2.0);
    double norm = 0, a0 = 0, a1 = 0, a2 = 0, b1 = 0, b2 = 0;
    const double G = pow(10.0, fabs(peakGain) / 20.0);
    const double K = tan(3.1415926535897932384626433832795 * cutoff / sr);
    switch (type) {
        case MXB_BQ_LOWPASS:
            norm = 1.0 / (1.0 + K / Q + K * K);
            a0 = K * K * norm; a1 = 2.0 * a0; a2 = a0;
            b1 = 2.0 * (K * K - 1.0) * norm; b2 = (1.0 - K / Q + K * K) * norm; break;
        case MXB_BQ_HIGHPASS:
            norm = 1. / (1. + K / Q + K * K);
            a0 = 1 * norm; a1 = -2 * a0; a2 = a0;
            b1 = 2 * (K * K - 1) * norm; b2 = (1 - K / Q + K * K) * norm; break;
        case MXB_BQ_BANDPASS:
            norm = 1. / (1. + K / Q + K * K);
            a0 = K / Q * norm; a1 = 0.; a2 = -a0;
            b1 = 2. * (K * K - 1.) * norm; b2 = (1. - K / Q + K * K) * norm; break;
        case MXB_BQ_NOTCH:
            norm = 1. / (1. + K / Q + K * K);
            a0 = (1. + K * K) * norm; a1 = 2. * (K * K - 1.) * norm; a2 = a0;
            b1 = a1; b2 = (1. - K / Q + K * K) * norm; break;
        case MXB_BQ_PEAK:
            if (peakGain >= 0.0) {
                norm = 1. / (1. + 1. / Q * K + K * K);
                a0 = (1. + G / Q * K + K * K) * norm; a1 = 2. * (K * K - 1.) * norm;
                a2 = (1. - G / Q * K + K * K) * norm; b1 = a1; b2 = (1. - 1. / Q * K + K * K) * norm;
            } else {
                norm = 1. / (1. + G / Q * K + K * K);
                a0 = (1. + 1 / Q * K + K * K) * norm; a1 = 2. * (K * K - 1) * norm;
                a2 = (1. - 1. / Q * K + K * K) * norm; b1 = a1; b2 = (1. - G / Q * K + K * K) * norm;
            }
            break;
        case MXB_BQ_LOWSHELF:
            if (peakGain >= 0.) {
                norm = 1. / (1. + SQRT2 * K + K * K);
                a0 = (1. + sqrt(2. * G) * K + G * K * K) * norm; a1 = 2. * (G * K * K - 1.) * norm;
                a2 = (1. - sqrt(2. * G) * K + G * K * K) * norm;
                b1 = 2. * (K * K - 1.) * norm; b2 = (1. - SQRT2 * K + K * K) * norm;
            } else {
                norm = 1. / (1. + sqrt(2. * G) * K + G * K * K);
                a0 = (1. + SQRT2 * K + K * K) * norm; a1 = 2. * (K * K - 1.) * norm;
                a2 = (1. - SQRT2 * K + K * K) * norm;
                b1 = 2. * (G * K * K - 1.) * norm; b2 = (1. - sqrt(2. * G) * K + G * K * K) * norm;
            }
            break;
        case MXB_BQ_HIGHSHELF:
            if (peakGain >= 0.) {
                norm = 1. / (1. + SQRT2 * K + K * K);
                a0 = (G + sqrt(2. * G) * K + K * K) * norm; a1 = 2. * (K * K - G) * norm;
                a2 = (G - sqrt(2. * G) * K + K * K) * norm;
                b1 = 2. * (K * K - 1) * norm; b2 = (1. - SQRT2 * K + K * K) * norm;
            } else {
                norm = 1. / (G + sqrt(2. * G) * K + K * K);
                a0 = (1. + SQRT2 * K + K * K) * norm; a1 = 2. * (K * K - 1.) * norm;
                a2 = (1. - SQRT2 * K + K * K) * norm;
                b1 = 2. * (K * K - G) * norm; b2 = (G - sqrt(2. * G) * K + K * K) * norm;
            }
            break;
        default: break;
    }
    cf[0] = a0; cf[1] = a1; cf[2] = a2; cf[3] = b1; cf[4] = b2;
}

}  // namespace mxb
