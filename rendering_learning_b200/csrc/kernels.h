// host-callable launchers of the render / trace kernels (defined in rtc_kernels.cu, ow_kernels.cu)
#pragma once
#include <cuda_runtime.h>

#include "scene.h"

namespace rl {

struct Counters;
struct JobTable;

cudaError_t launch_rtc_render(const DevScene& sc, const rl_rtc_camera* cam, const double inv[12], uint32_t aa,
                              const JobTable& jt, float* d_out, Counters* d_counters, bool instrumented,
                              cudaStream_t stream);
cudaError_t launch_rtc_trace(const DevScene& sc, const rl_ray* d_rays, uint64_t n, rl_hit* d_hits,
                             Counters* d_counters, bool instrumented, cudaStream_t stream);

int ow_num_chunks(int spp);
int ow_image_height(const rl_ow_camera* c);
cudaError_t launch_ow_render(const DevScene& sc, const rl_ow_camera* cam, uint32_t first_sample, const JobTable& jt,
                             float* d_partial, unsigned long long* d_queue, Counters* d_counters, bool instrumented,
                             int sm_count, cudaStream_t stream, bool shared_queue = false);
cudaError_t launch_ow_reduce(const rl_ow_camera* cam, const float* d_partial, float* d_out, cudaStream_t stream);
// 8-bit output encoders (RTC/src/draw/canvas.rs:53-56; OW/src/color.rs:47-57, 130-136), n = W*H*3 channels
cudaError_t launch_encode_rtc_u8(const float* d_rgb, uint8_t* d_out, size_t n, cudaStream_t stream);
cudaError_t launch_encode_ow_u8(const float* d_rgb_sum, uint8_t* d_out, size_t n, int samples, cudaStream_t stream);
cudaError_t launch_ow_trace(const DevScene& sc, const rl_ray* d_rays, uint64_t n, rl_hit* d_hits, Counters* d_counters,
                            bool instrumented, cudaStream_t stream);

}  // namespace rl
