// probe_barrier.cu -- latency of grid-barrier designs on 148 co-resident CTAs (cooperative launch), used to pick
// the barrier of the persistent sampler (denoise_mega.cu).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_barrier tools/probe_barrier.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
#define SPIN_LIMIT (1L << 17)

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_release(unsigned int *p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// V0: the current design: one counter, thread 0 fences, adds, spins with acquire loads
__device__ __forceinline__ bool bar_v0(unsigned int *bar, unsigned int &epoch) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        long spins = 0;
        while (ld_acquire(bar) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
    }
    __syncthreads();
    return ok;
}
// V1: all-gather of per-CTA epoch flags: release store of the own flag, one warp polls all flags
__device__ __forceinline__ bool bar_v1(unsigned int *flags, unsigned int &epoch) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < 32) {
        ++epoch;
        if (threadIdx.x == 0) st_release(flags + blockIdx.x, epoch);
        long spins = 0;
        for (;;) {
            bool done = true;
            for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) done &= ld_acquire(flags + i) >= epoch;
            if (__all_sync(0xffffffffu, done)) break;
            if (++spins > SPIN_LIMIT) { ok = false; break; }
        }
    }
    __syncthreads();
    return ok;
}
// V2: like V1, polling with relaxed loads and one fence at the end
__device__ __forceinline__ bool bar_v2(unsigned int *flags, unsigned int &epoch) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x < 32) {
        ++epoch;
        if (threadIdx.x == 0) st_release(flags + blockIdx.x, epoch);
        long spins = 0;
        for (;;) {
            bool done = true;
            for (unsigned int i = threadIdx.x; i < gridDim.x; i += 32) done &= ld_relaxed(flags + i) >= epoch;
            if (__all_sync(0xffffffffu, done)) break;
            if (++spins > SPIN_LIMIT) { ok = false; break; }
        }
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    return ok;
}
// V3: counter + separate release word: the last arriver (sees count == epoch - 1) publishes the epoch; pollers
// spin on the release word only, so polls do not queue behind the atomics
__device__ __forceinline__ bool bar_v3(unsigned int *bar, unsigned int &epoch, unsigned int &gen) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        ++gen;
        __threadfence();
        unsigned int old = atomicAdd(bar, 1u);
        if (old == epoch - 1) st_release(bar + 32, gen);
        long spins = 0;
        while (ld_acquire(bar + 32) < gen) if (++spins > SPIN_LIMIT) { ok = false; break; }
    }
    __syncthreads();
    return ok;
}
// V4: red.release (no separate fence, no return value) + acquire spin
__device__ __forceinline__ bool bar_v4(unsigned int *bar, unsigned int &epoch) {
    __syncthreads();
    bool ok = true;
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        red_release(bar, 1u);
        long spins = 0;
        while (ld_acquire(bar) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
    }
    __syncthreads();
    return ok;
}
// V5: two-level flags: 148 CTAs in groups of 16 (flag lines per group), group leader gathers its group and publishes a
// group flag; everyone polls the <= 10 group flags
__device__ __forceinline__ bool bar_v5(unsigned int *flags, unsigned int &epoch) {
    __syncthreads();
    bool ok = true;
    const unsigned int grp = blockIdx.x >> 4, ngrp = (gridDim.x + 15) >> 4;
    unsigned int *gflags = flags + 1024;
    if (threadIdx.x < 32) {
        ++epoch;
        if (threadIdx.x == 0) st_release(flags + blockIdx.x, epoch);
        long spins = 0;
        if ((blockIdx.x & 15) == 0) {
            const unsigned int n = min(16u, gridDim.x - grp * 16);
            for (;;) {
                bool done = threadIdx.x >= n || ld_acquire(flags + grp * 16 + threadIdx.x) >= epoch;
                if (__all_sync(0xffffffffu, done)) break;
                if (++spins > SPIN_LIMIT) { ok = false; break; }
            }
            if (threadIdx.x == 0) st_release(gflags + grp * 32, epoch);
        }
        for (;;) {
            bool done = threadIdx.x >= ngrp || ld_acquire(gflags + threadIdx.x * 32) >= epoch;
            if (__all_sync(0xffffffffu, done)) break;
            if (++spins > SPIN_LIMIT) { ok = false; break; }
        }
    }
    __syncthreads();
    return ok;
}

// payload: every CTA adds `nred` floats into a shared 16 KB region before the barrier and reads 16 KB after it
// (what a phase boundary of the sampler does)
__global__ void __launch_bounds__(256, 1) probe_kernel(unsigned int *bar, float *acc, int iters, int variant, int nred, int nload,
                                                       long long *cycles, unsigned int *fail) {
    unsigned int epoch = 0, gen = 0;
    float sink = 0.f;
    bar_v0(bar + 2048, epoch);   // line everybody up (separate counter)
    epoch = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int i = threadIdx.x; i < nred; i += 256) atomicAdd(acc + ((i * 8 + (i >> 9)) & 4095), 1.0f);
        bool ok;
        switch (variant) {
            case 0: ok = bar_v0(bar, epoch); break;
            case 1: ok = bar_v1(bar, epoch); break;
            case 2: ok = bar_v2(bar, epoch); break;
            case 3: ok = bar_v3(bar, epoch, gen); break;
            case 4: ok = bar_v4(bar, epoch); break;
            default: ok = bar_v5(bar, epoch); break;
        }
        if (!ok) { atomicExch(fail, 1u); break; }
        if (*reinterpret_cast<volatile unsigned int *>(fail)) break;
        for (int i = threadIdx.x; i < nload; i += 256) sink += __ldcg(acc + i);
    }
    long long t1 = clock64();
    if (sink == -1.f) acc[0] = sink;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    unsigned int *bar, *fail; float *acc; long long *cyc;
    CK(cudaMalloc(&bar, 16384)); CK(cudaMalloc(&acc, 4096 * 4)); CK(cudaMalloc(&cyc, 8)); CK(cudaMalloc(&fail, 4));
    const int iters = 4000;
    const char *names[] = {"v0 counter+spin (current)", "v1 flag all-gather, acquire polls", "v2 flag all-gather, relaxed polls + fence",
                           "v3 counter + release word", "v4 red.release + spin", "v5 two-level flags"};
    for (int nred : {0, 256, 4096})
        for (int nload : {0, 4096}) {
            printf("reds per CTA before the barrier: %d, floats loaded after: %d\n", nred, nload);
            for (int grid : {148, 128})
                for (int v = 0; v < 6; ++v) {
                    CK(cudaMemset(bar, 0, 16384)); CK(cudaMemset(acc, 0, 16384)); CK(cudaMemset(fail, 0, 4));
                    int it = iters;
                    void *args[] = {&bar, &acc, &it, &v, &nred, &nload, &cyc, &fail};
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0); cudaEventCreate(&e1);
                    cudaEventRecord(e0);
                    cudaError_t e = cudaLaunchCooperativeKernel((const void *)probe_kernel, dim3(grid), dim3(256), args, 0, 0);
                    cudaEventRecord(e1);
                    cudaError_t e2 = cudaDeviceSynchronize();
                    float ms = 0.f;
                    cudaEventElapsedTime(&ms, e0, e1);
                    long long c = 0; unsigned int f = 0;
                    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
                    cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost);
                    printf("  grid=%3d %-44s %.3f us / iter (%lld cycles)%s %s\n", grid, names[v], ms * 1e3f / iters, c / iters,
                           f ? "  TIMED OUT" : "", (e != cudaSuccess || e2 != cudaSuccess) ? cudaGetErrorString(e2) : "");
                }
        }
    return 0;
}
