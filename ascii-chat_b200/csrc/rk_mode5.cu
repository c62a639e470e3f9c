// rk_mode5.cu — instantiates the row kernels for EmitMode 5 (one translation unit per mode keeps builds parallel)
#include "render_dev.cuh"
namespace acb {
cudaError_t launch_rows_m5(const RenderParams &p, int sp, cudaStream_t st, unsigned *grid_out) {
  return launch_rows_mode<5>(p, sp, st, grid_out);
}
cudaError_t launch_ws_m5(const RenderParams &p, cudaStream_t st) { return launch_ws_mode<5>(p, st); }
cudaError_t launch_ws2_m5(const RenderParams &p, cudaStream_t st, unsigned *grid_out) {
  return launch_ws2_mode<5>(p, st, grid_out);
}
} // namespace acb
