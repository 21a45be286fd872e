// Roofline denominators for this path, measured on the device the ctx drives (SURVEY.md §6 / BASELINE.md §2):
// the ray loop is bounded by the FP32 (non-tensor) pipe and by L1/L2 fetch bandwidth, neither of which the
// driver-written MEASURED_PEAKS.json covers (it has HBM copy + bf16 GEMM).
//   fp32 : register-resident FMA chains, 8 independent accumulators per thread, every SM full
//   l2   : repeated 16-byte ld.global.cg sweeps over a 32 MiB buffer (fits L2, bypasses L1)
//   hbm  : one 16-byte-vector read sweep over a 1 GiB buffer (>> L2)
#include <cuda_runtime.h>

#include <algorithm>

namespace rl {
namespace {

__global__ void __launch_bounds__(256) k_fma(float* out, int iters, float b, float c) {
    float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

__global__ void __launch_bounds__(256) k_sweep(const float4* __restrict__ buf, size_t n_vec, int passes, float* out) {
    float acc = 0.0f;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; p++) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
            float4 v = __ldcg(buf + i);
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) out[0] = acc;  // keep the loads alive
}

float best_ms(cudaStream_t s, int reps, void (*launch)(void*), void* arg) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0, s);
        launch(arg);
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (r > 0) best = std::min(best, ms);  // first repetition is the warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return best;
}

struct Args {
    cudaStream_t s;
    int sm;
    float* out;
    const float4* buf;
    size_t n_vec;
    int passes, iters;
};

}  // namespace

cudaError_t measure_peaks(cudaStream_t s, int sm_count, double* fp32_tflops, double* l2_gbs, double* hbm_gbs) {
    Args a{};
    a.s = s;
    a.sm = sm_count;
    cudaError_t e;
    float* out = nullptr;
    if ((e = cudaMalloc(&out, sizeof(float) * (size_t)sm_count * 16 * 256)) != cudaSuccess) return e;
    a.out = out;
    a.iters = 2048;
    float ms = best_ms(s, 5, [](void* p) {
        Args* a = (Args*)p;
        k_fma<<<a->sm * 16, 256, 0, a->s>>>(a->out, a->iters, 1.000001f, 0.5f);
    }, &a);
    double flops = (double)sm_count * 16 * 256 * (double)a.iters * 64.0 * 2.0;
    if (fp32_tflops) *fp32_tflops = flops / (ms * 1e-3) / 1e12;

    float4* buf = nullptr;
    size_t l2_bytes = 32ull << 20;
    if ((e = cudaMalloc(&buf, l2_bytes)) != cudaSuccess) { cudaFree(out); return e; }
    cudaMemsetAsync(buf, 0, l2_bytes, s);
    a.buf = buf;
    a.n_vec = l2_bytes / sizeof(float4);
    a.passes = 100;
    ms = best_ms(s, 5, [](void* p) {
        Args* a = (Args*)p;
        k_sweep<<<a->sm * 8, 256, 0, a->s>>>(a->buf, a->n_vec, a->passes, a->out);
    }, &a);
    if (l2_gbs) *l2_gbs = (double)l2_bytes * a.passes / (ms * 1e-3) / 1e9;
    cudaFree(buf);

    size_t hbm_bytes = 1ull << 30;
    if ((e = cudaMalloc(&buf, hbm_bytes)) != cudaSuccess) { cudaFree(out); return e; }
    cudaMemsetAsync(buf, 0, hbm_bytes, s);
    a.buf = buf;
    a.n_vec = hbm_bytes / sizeof(float4);
    a.passes = 1;
    ms = best_ms(s, 5, [](void* p) {
        Args* a = (Args*)p;
        k_sweep<<<a->sm * 16, 256, 0, a->s>>>(a->buf, a->n_vec, a->passes, a->out);
    }, &a);
    if (hbm_gbs) *hbm_gbs = (double)hbm_bytes / (ms * 1e-3) / 1e9;
    cudaFree(buf);
    cudaFree(out);
    return cudaGetLastError();
}

}  // namespace rl
