// csrc/rpw_fit_replay.cu — the same fit kernels carrying the reference-order arithmetic (RPW_SOLVER_REFERENCE,
// rpw_set_exact_replay): launched instead of the default ones whenever FitParams::exact_replay >= 0.
#include "rpw_fit.cuh"

namespace rpw {
cudaError_t fit_configure_replay(int smem_cap, int* blocks_per_sm) { return fit_configure_t<true>(smem_cap, blocks_per_sm); }
cudaError_t launch_fit_roots_replay(cudaStream_t st, const FitArgs& args, int cls, unsigned grid) { return launch_fit_roots_t<true>(st, args, cls, grid); }
cudaError_t launch_fit_levels_replay(cudaStream_t st, const FitArgs& args, int grid_blocks) { return launch_fit_levels_t<true>(st, args, grid_blocks); }
}  // namespace rpw
