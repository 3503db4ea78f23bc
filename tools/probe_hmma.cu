// probe_hmma.cu -- latency / issue rate of the legacy mma.sync.m16n8k16 bf16 path on sm_100a (one warp, and 8 warps per SM)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_hmma tools/probe_hmma.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

__device__ __forceinline__ void mma(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(long long *out, float *sink, int iters) {
    float c[CHAINS][4];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.f;
    uint32_t a = threadIdx.x * 0x01010101u, b = 0x3f803f80u;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) mma(c[i], a, a + 1, a + 2, a + 3, b, b);
    }
    long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 123.f) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

template <int CHAINS>
int run(int threads, long long *dout, float *sink) {
    const int iters = 1024;
    k<CHAINS><<<1, threads>>>(dout, sink, iters);
    CK(cudaDeviceSynchronize());
    k<CHAINS><<<1, threads>>>(dout, sink, iters);
    CK(cudaDeviceSynchronize());
    long long h;
    CK(cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost));
    printf("threads %4d chains %2d: %7.1f cycles per loop iteration, %6.1f cycles per mma per warp\n", threads, CHAINS,
           (double)h / iters, (double)h / iters / CHAINS);
    return 0;
}

int main() {
    long long *dout; float *sink;
    CK(cudaMalloc(&dout, 64)); CK(cudaMalloc(&sink, 64));
    for (int threads : {32, 128, 256}) {
        run<1>(threads, dout, sink);
        run<2>(threads, dout, sink);
        run<4>(threads, dout, sink);
        run<8>(threads, dout, sink);
    }
    return 0;
}
