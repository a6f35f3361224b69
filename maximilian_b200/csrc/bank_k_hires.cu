// Explicit instantiations of the K1 bank kernel for FILT_T_HIRES (one translation unit per filter family
// so that the families compile in parallel). See bank_kernels.cuh.
#include "bank_kernels.cuh"

namespace mxb {
int launch_bank_hires(const BankArgs& a, int osc_t, int env, bool out, bool mix, int grid, size_t smem, cudaStream_t s) {
    return launch_bank_filt<FILT_T_HIRES>(a, osc_t, env, out, mix, grid, smem, s);
}
}  // namespace mxb
