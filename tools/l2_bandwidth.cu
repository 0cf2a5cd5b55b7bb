// L2 read bandwidth of this GPU: every SM streams a buffer that fits in L2 (32 MB by default) with 128-bit loads,
// many passes, CUDA events.  SURVEY 8(d) / BASELINE 3.3: the C1/C2/C4 scene (4 MB) is L2-resident, so the memory
// roofline of the traversal kernel on those configs is L2 bandwidth, which MEASURED_PEAKS.json does not hold.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/l2_bandwidth.cu -o tools/l2_bandwidth && tools/l2_bandwidth [MB]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void __launch_bounds__(256) k_read(const uint4 *p, size_t n, int passes, unsigned *sink) {
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < passes; ++r)
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
            const uint4 v = __ldcg(p + i);  // L2 only: L1 must not serve it
            acc ^= v.x ^ v.y ^ v.z ^ v.w;
        }
    if (acc == 0x12345678u) *sink = acc;
}

int main(int argc, char **argv) {
    const size_t mb = argc > 1 ? (size_t)atoi(argv[1]) : 32;
    const size_t bytes = mb << 20, n = bytes / 16;
    uint4 *buf; unsigned *sink;
    cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, bytes);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int grid = prop.multiProcessorCount * 8, passes = 200;
    k_read<<<grid, 256>>>(buf, n, 4, sink);  // warm: pull the buffer into L2
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a);
        k_read<<<grid, 256>>>(buf, n, passes, sink);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double gbs = (double)bytes * passes / (best * 1e-3) * 1e-9;
    printf("{\"l2_read_gbs\": %.1f, \"buffer_mb\": %zu, \"passes\": %d, \"ms\": %.3f, \"gpu\": \"%s\", \"sms\": %d, \"how\": \"ld.global.cg 128-bit, grid 8 x SMs x 256 threads, best of 5\"}\n",
           gbs, mb, passes, best, prop.name, prop.multiProcessorCount);
    return cudaGetLastError() == cudaSuccess ? 0 : 1;
}
