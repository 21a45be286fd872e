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

// Sample partition of an OW render, a function of spp ALONE (so the image is independent of the schedule and the GPU
// count): `body` chunks of about equal size, then OW_TAIL_CHUNKS chunks of OW_TAIL_SIZE samples.  Items are popped
// chunk-major, so the LAST items of a render are the small ones: what the slowest warp still holds when every other
// warp of every GPU has run dry is a 2-sample item, not an 8-sample one (the load-balancing tail of an 8-GPU step).
#ifndef RL_OW_TAIL_CHUNKS  // overridable for tools/build_alt.py experiments only
#define RL_OW_TAIL_CHUNKS 8
#define RL_OW_TAIL_SIZE 2
#endif
constexpr int OW_TAIL_CHUNKS = RL_OW_TAIL_CHUNKS, OW_TAIL_SIZE = RL_OW_TAIL_SIZE;
__host__ __device__ inline int ow_tail_chunks(int spp) { return spp >= 64 ? OW_TAIL_CHUNKS : 0; }
__host__ __device__ inline void ow_chunk_range(int spp, int n_chunks, int chunk, int* s0, int* s1) {
    const int t = ow_tail_chunks(spp), nb = n_chunks - t, body = spp - t * OW_TAIL_SIZE;
    if (chunk < nb) {
        *s0 = (int)(((long long)chunk * body) / nb);
        *s1 = (int)(((long long)(chunk + 1) * body) / nb);
    } else {
        *s0 = body + (chunk - nb) * OW_TAIL_SIZE;
        *s1 = *s0 + OW_TAIL_SIZE;
    }
}
// Scheduling parameters of the OW render kernel.  0 = the measured default of the scene's instantiation.  They change
// WHEN work is done, never what is computed: the image is bit-identical for every setting (tests/test_gpu_ow.py).
struct OwTuning {
    int variant = 5;      // 5: per-lane paths, service rounds inside the warp (production); 6: CTA-pooled paths (the measured
                          // shared-memory wavefront experiment, 1.6-1.9x slower: DESIGN.md §4); 7: the GLOBAL wavefront
                          // (path state in L2 / HBM, one logic + one trace kernel per bounce: the other measured A/B)
    int slots = 0;        // v6: path slots per CTA (256 .. 512)
    int minb = 0;         // resident CTAs per SM the kernel is compiled for (3 or 4)
    int ctas_per_sm = 0;  // launch fewer CTAs per SM than fit
    int svc_lo = 0;       // v6: a warp with at most this many traversing lanes services a partial batch
    int exit_min = 0;     // v6: lanes that must finish before the warp leaves the traversal loop to refill
    int leaf_min = 0;     // parked lanes per leaf round
    int svc_min = 0;      // v5: lanes that must wait before the warp services them
};
cudaError_t launch_tri_planes(const TriVerts* d_verts, int n, OwTriPlane* d_planes, cudaStream_t stream);
int ow_num_chunks(int spp);
int ow_image_height(const rl_ow_camera* c);
cudaError_t launch_ow_render(const DevScene& sc, const rl_ow_camera* cam, uint32_t first_sample, const JobTable& jt,
                             float* d_partial, unsigned long long* d_queue, Counters* d_counters, bool instrumented,
                             int sm_count, cudaStream_t stream, bool shared_queue, const OwTuning& tune);
// global-wavefront variant (ow.variant = 7): state [wf words][slots] f32, ray list [slots] i32, a small counter block
struct WavefrontBuffers {
    float* state;
    int* ray_list;
    void* ctr;
    int slots;
};
constexpr int OW_WF_WORDS = 25;
cudaError_t launch_ow_wavefront(const DevScene& sc, const rl_ow_camera* cam, uint32_t first_sample, const JobTable& jt,
                                float* d_partial, unsigned long long* d_queue, Counters* d_counters, const WavefrontBuffers& wb,
                                int sm_count, cudaStream_t stream, const OwTuning& tune, int* launches);
cudaError_t launch_ow_reduce(const rl_ow_camera* cam, const float* d_partial, float* d_out, cudaStream_t stream);
// 8-bit output encoders (RTC/src/draw/canvas.rs:53-56; OW/src/color.rs:47-57, 130-136), n = W*H*3 channels
cudaError_t launch_encode_rtc_u8(const float* d_rgb, uint8_t* d_out, size_t n, cudaStream_t stream);
cudaError_t launch_encode_ow_u8(const float* d_rgb_sum, uint8_t* d_out, size_t n, int samples, cudaStream_t stream);
cudaError_t launch_ow_trace(const DevScene& sc, const rl_ray* d_rays, const int* d_self_refs, uint64_t n, rl_hit* d_hits,
                            unsigned long long* d_queue, Counters* d_counters, bool instrumented, int sm_count,
                            cudaStream_t stream, const OwTuning& tune);

}  // namespace rl
