// K2 interface: bank kernel with a maxiDelayline stage (see delay.cu).
#pragma once

#include "bank_kernels.cuh"

namespace mxb {

#ifndef MXB_DL_SHIFT
#define MXB_DL_SHIFT 4
#endif
constexpr int kDlShift = MXB_DL_SHIFT;      // 4, or 3 for an A/B build with 8-slot chunks (half the staging per warp)
constexpr int kDlChunk = 1 << kDlShift;     // ring slots per chunk = time steps per staged window (16)
static_assert(kDlShift == 3 || kDlShift == 4, "chunks of 8 or 16 slots");
// the swizzle of voice v's chunk: v % 16 for 16-slot chunks; (v / 2) % 8 for 8-slot chunks (rows of 64 bytes: two lanes share a
// 128-byte bank row, the 8 even lanes of a half-warp must differ in the low half of it and the 8 odd ones in the high half)
__host__ __device__ inline int dl_swz(size_t v) { return (int)((v >> (4 - kDlShift)) & (size_t)(kDlChunk - 1)); }

// Ring storage is chunk-interleaved and swizzled: slot r of voice v lives at ((r / 16) * V + v) * 16 + ((r % 16) ^ (v % 16)).
// A voice's 16-slot chunk is 128 contiguous bytes (4 full sectors whatever its phase), and voices whose
// ring indices run in step -- the common case: same size, started together -- read and write one
// contiguous run of 32 * 128 B per warp and stage: ONE bulk copy (cp.async.bulk, the TMA engine) in each direction moves a
// warp's window between HBM and shared memory. The XOR with the voice index is what makes the image that lands in shared
// memory usable as it is: lane v reads slot j of its row at position j ^ (v % 16), so the 16 lanes of a half-warp hit 16
// different 8-byte bank pairs -- conflict-free without padding, which a bulk copy could not produce.
__host__ __device__ inline size_t dl_slot(size_t V, size_t v, int r) {
    return (((size_t)(r >> kDlShift)) * V + v) * kDlChunk + (size_t)((r ^ dl_swz(v)) & (kDlChunk - 1));
}
inline size_t dl_ring_doubles(size_t V, int taps) { return (size_t)((taps + kDlChunk - 1) / kDlChunk) * kDlChunk * V; }

struct DelayArgs {
    int* phase;              // maxiDelayline::phase per voice
    const int* size;         // dl() size argument per voice
    const double* feedback;
    double* ring;
    int taps;
    int from_position;       // 1: maxiDelayline::dlFromPosition
    const int* position;     // its position argument per voice
    const double* size_tv;   // optional per-sample size argument [n_frames][V] (integral values), else NULL
    int W_out;               // [out] warps in the launched grid (mix partials stride)
};

// warps the K2 grid will have for V voices (the caller sizes the mix partials with it)
int delay_bank_warps(int V);

int launch_delay_bank(const BankArgs& a, DelayArgs& d, int filt_kind, bool svf_lp, int env, bool out, bool mix,
                      cudaStream_t s);

// one launcher per filter family (delay_k_*.cu); outmode 0 none / 1 fp64 / 2 fp32
int launch_delay_none(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_lores(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_hires(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_svf(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_svf_lp(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_biquad(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
// ... and for blocks with a per-sample oscillator frequency / filter cutoff (delay_km_*.cu)
int launch_delay_none_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_lores_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_hires_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_svf_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_svf_lp_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);
int launch_delay_biquad_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s);

}  // namespace mxb
