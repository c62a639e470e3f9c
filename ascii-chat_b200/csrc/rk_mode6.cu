// rk_mode6.cu — instantiates the row kernels for EmitMode 6 (one translation unit per mode keeps builds parallel)
#include "render_dev.cuh"
namespace acb {
cudaError_t launch_rows_m6(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out) {
  return launch_rows_mode<6>(p, sp, st, grid_out);
}
cudaError_t launch_ws_m6(const RenderParams &p, cudaStream_t st) { return launch_ws_mode<6>(p, st); }
cudaError_t launch_ws2_m6(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  return launch_ws2_mode<6>(p, st, grid_out);
}
} // namespace acb
