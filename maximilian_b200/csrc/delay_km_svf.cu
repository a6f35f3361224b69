// Explicit instantiations of the K2 delay-line bank kernel for FILT_T_SVF, blocks with a per-sample frequency / cutoff (see delay_impl.cuh).
#include "delay_impl.cuh"

namespace mxb {
int launch_delay_svf_mod(const BankArgs& a, const DelayArgs& d, int osc_saw, int env, int outmode, bool mix, int grid, cudaStream_t s) {
    return launch_delay_filt<FILT_T_SVF, true>(a, d, osc_saw, env, outmode, mix, grid, s);
}
}  // namespace mxb
