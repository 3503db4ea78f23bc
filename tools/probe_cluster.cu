// probe_cluster.cu -- hardware facts the cluster-fused sampler (denoise_mega3.cu) is sized from:
// how many 16-CTA / 8-CTA clusters are co-resident with ~200 KB of shared memory per CTA, whether a
// cluster launch can also be cooperative, and the latency of: a grid barrier (atomic + acquire spin),
// cluster.sync, a DSMEM reduce-scatter, and a burst of fp32 reductions into one 16 KB region.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/probe_cluster tools/probe_cluster.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)

__device__ __forceinline__ void grid_barrier(unsigned int *bar, unsigned int &target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        long spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
            if (++spins > (1L << 17) || *reinterpret_cast<volatile unsigned int *>(bar + 1)) { atomicExch(bar + 1, 1u); break; }
        } while (v < target);
    }
    __syncthreads();
}

// mode 0: grid barrier only; 1: cluster.sync only; 2: 4096 reds per CTA + grid barrier;
// 3: DSMEM reduce-scatter (each CTA writes 256 B to every peer) + cluster.sync; 4: 256 reds per CTA + barrier
__global__ void __launch_bounds__(256, 1) probe_kernel(unsigned int *bar, float *acc, int iters, int mode, long long *cycles) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *sbuf = reinterpret_cast<float *>(smem);
    cg::cluster_group cluster = cg::this_cluster();
    unsigned int target = 0;
    const unsigned int cs = cluster.num_blocks(), cr = cluster.block_rank();
    grid_barrier(bar, target);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (*reinterpret_cast<volatile unsigned int *>(bar + 1)) break;   // a barrier timed out: everybody leaves
        if (mode == 0) {
            grid_barrier(bar, target);
        } else if (mode == 1) {
            cluster.sync();
        } else if (mode == 2) {
            for (int i = threadIdx.x; i < 4096; i += 256) atomicAdd(acc + i, 1.0f);
            grid_barrier(bar, target);
        } else if (mode == 3) {
            // thread -> (peer = tid / 16, 16 floats... 4 floats each of 64 B) : 16 peers x 64 floats = 1024 floats per CTA
            for (int i = threadIdx.x; i < (int)cs * 64; i += 256) {
                int peer = i / 64, j = i % 64;
                float *dst = cluster.map_shared_rank(sbuf + cr * 64 + j, peer);
                *dst = (float)it;
            }
            cluster.sync();
            float s = 0.f;
            if (threadIdx.x < 64) for (unsigned int r = 0; r < cs; ++r) s += sbuf[r * 64 + threadIdx.x];
            if (s == -1.f) acc[0] = s;
            cluster.sync();
        } else if (mode == 4) {
            atomicAdd(acc + ((blockIdx.x * 64 + threadIdx.x) & 4095), 1.0f);
            grid_barrier(bar, target);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

static float run(int grid, int cs, bool coop, int mode, int iters, size_t smem, unsigned int *bar, float *acc, long long *cyc) {
    CK(cudaMemset(bar, 0, 64));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cs > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
    if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na; }
    cfg.attrs = at; cfg.numAttrs = na;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel, bar, acc, iters, mode, cyc);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e != cudaSuccess || e2 != cudaSuccess) {
        printf("  launch grid=%d cs=%d coop=%d mode=%d: %s / %s\n", grid, cs, (int)coop, mode, cudaGetErrorString(e), cudaGetErrorString(e2));
        cudaGetLastError();
        return -1.f;
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    unsigned int flag[2] = {0, 0};
    cudaMemcpy(flag, bar, 8, cudaMemcpyDeviceToHost);
    if (flag[1]) printf("  (BARRIER TIMED OUT: CTAs not co-resident?) ");
    printf("  grid=%3d cs=%2d coop=%d mode=%d: %.3f us / iter (%lld cycles / iter)\n", grid, cs, (int)coop, mode, ms * 1e3f / iters, c / iters);
    return ms;
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    const size_t smem = 200 * 1024;
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    for (int cs : {2, 4, 8, 16}) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(cs * 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        printf("max active clusters, cluster size %2d, %zu KB smem: %d (%s)\n", cs, smem / 1024, n, cudaGetErrorString(e));
    }
    unsigned int *bar; float *acc; long long *cyc;
    CK(cudaMalloc(&bar, 64)); CK(cudaMalloc(&acc, 4096 * 4)); CK(cudaMalloc(&cyc, 8));
    CK(cudaMemset(acc, 0, 4096 * 4));
    const int iters = 2000;
    printf("grid barrier:\n");
    run(148, 1, true, 0, iters, smem, bar, acc, cyc);
    run(128, 1, true, 0, iters, smem, bar, acc, cyc);
    run(128, 16, true, 0, iters, smem, bar, acc, cyc);
    run(128, 16, false, 0, iters, smem, bar, acc, cyc);
    run(144, 8, true, 0, iters, smem, bar, acc, cyc);
    run(128, 8, true, 0, iters, smem, bar, acc, cyc);
    printf("cluster.sync:\n");
    run(128, 16, false, 1, iters, smem, bar, acc, cyc);
    run(128, 8, false, 1, iters, smem, bar, acc, cyc);
    printf("4096 reds per CTA into one 16 KB region + grid barrier:\n");
    run(128, 16, false, 2, iters, smem, bar, acc, cyc);
    run(148, 1, true, 2, iters, smem, bar, acc, cyc);
    printf("256 reds per CTA + grid barrier:\n");
    run(128, 16, false, 4, iters, smem, bar, acc, cyc);
    printf("DSMEM reduce-scatter (256 B to every peer) + 2 cluster.sync:\n");
    run(128, 16, false, 3, iters, smem, bar, acc, cyc);
    run(128, 8, false, 3, iters, smem, bar, acc, cyc);
    return 0;
}
