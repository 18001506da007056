// csrc/rpw_fit_fast.cu — the fit kernels of the default path (K3a, K3b without the reference-order code).
#include "rpw_fit.cuh"

namespace rpw {
cudaError_t fit_configure_fast(int smem_cap, int* blocks_per_sm) { return fit_configure_t<false>(smem_cap, blocks_per_sm); }
cudaError_t launch_fit_roots_fast(cudaStream_t st, const FitArgs& args, int cls, unsigned grid) { return launch_fit_roots_t<false>(st, args, cls, grid); }
cudaError_t launch_fit_levels_fast(cudaStream_t st, const FitArgs& args, int grid_blocks) { return launch_fit_levels_t<false>(st, args, grid_blocks); }
ClassBounds fit_class_bounds(int profile) {
    ClassBounds cb;
    for (int c = 0; c < kNumFitClasses; ++c) cb.hi[c] = kFitClasses[profile ? 1 : 0][c].hi;
    return cb;
}
size_t fit_smem_bytes(int smem_cap, int threads) { return fit_smem_bytes_inl(smem_cap, threads); }
}  // namespace rpw
