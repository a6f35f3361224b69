// Spectral path (STFT / MFCC / ISTFT) -- placeholder entry points, replaced by the kernels next.
#include "common.cuh"

extern "C" {
#define MXB_TODO(name) do { mxb::set_error(name ": not built yet"); return MXB_ERR_UNSUPPORTED; } while (0)
int32_t mxb_stft_create(mxb_ctx*, int32_t, int32_t, int32_t, mxb_stft**) { MXB_TODO("mxb_stft_create"); }
int32_t mxb_stft_destroy(mxb_stft*) { return MXB_OK; }
int32_t mxb_stft_process(mxb_stft*, const float*, int64_t, int64_t, int32_t, int32_t, float*, float*, float*, float*,
                         mxb_mfcc*, double*, int32_t*, int32_t, void*) { MXB_TODO("mxb_stft_process"); }
int64_t mxb_stft_launch_count(const mxb_stft*) { return 0; }
int32_t mxb_mfcc_create(mxb_ctx*, int32_t, int32_t, int32_t, double, double, mxb_mfcc**) { MXB_TODO("mxb_mfcc_create"); }
int32_t mxb_mfcc_destroy(mxb_mfcc*) { return MXB_OK; }
int32_t mxb_mfcc_process(mxb_mfcc*, const float*, int64_t, double*, double*, int32_t, void*) { MXB_TODO("mxb_mfcc_process"); }
int32_t mxb_istft_create(mxb_ctx*, int32_t, int32_t, int32_t, mxb_istft**) { MXB_TODO("mxb_istft_create"); }
int32_t mxb_istft_destroy(mxb_istft*) { return MXB_OK; }
int32_t mxb_istft_process(mxb_istft*, const float*, const float*, int32_t, float*, int32_t, void*) { MXB_TODO("mxb_istft_process"); }
}
