// probe_barrier2.cu -- component latencies behind the grid barrier (fence, atomic, acquire load) and more
// barrier variants (sharded counters, relaxed polls, cluster-assisted arrival).
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); } } while (0)
#define SPIN_LIMIT (1L << 17)

__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) { unsigned int v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned int ld_relaxed(const unsigned int *p) { unsigned int v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned int ld_volatile(const unsigned int *p) { unsigned int v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void red_release(unsigned int *p, unsigned int v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_relaxed(unsigned int *p, unsigned int v) { asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// ---- single-CTA component latencies -------------------------------------------------------------------
__global__ void comp_kernel(unsigned int *buf, long long *out) {
    if (threadIdx.x != 0) return;
    unsigned int *p = buf + blockIdx.x * 64;
    const int N = 2000;
    unsigned int s = 0;
    long long t0 = clock64();
    for (int i = 0; i < N; ++i) __threadfence();
    long long t1 = clock64();
    for (int i = 0; i < N; ++i) s += atomicAdd(p, 1u) & 0;     // dependent atomic round trips
    long long t2 = clock64();
    for (int i = 0; i < N; ++i) s += ld_acquire(p + (s & 1));  // dependent acquire loads
    long long t3 = clock64();
    for (int i = 0; i < N; ++i) s += ld_relaxed(p + (s & 1));
    long long t4 = clock64();
    for (int i = 0; i < N; ++i) s += ld_volatile(p + (s & 1));
    long long t5 = clock64();
    for (int i = 0; i < N; ++i) { p[1] = i; __threadfence(); } // fence with one store outstanding
    long long t6 = clock64();
    for (int i = 0; i < N; ++i) { atomicAdd(p + 2, 1u); __threadfence(); } // fence with one red outstanding
    long long t7 = clock64();
    if (blockIdx.x == 0) {
        out[0] = (t1 - t0) / N; out[1] = (t2 - t1) / N; out[2] = (t3 - t2) / N; out[3] = (t4 - t3) / N; out[4] = (t5 - t4) / N;
        out[5] = (t6 - t5) / N; out[6] = (t7 - t6) / N; out[7] = s;
    }
}

// ---- barrier variants ---------------------------------------------------------------------------------------
// 0: current (fence + atomicAdd + acquire spin)             1: relaxed red + fence before, relaxed polls + fence after
// 2: sharded counters (4 lines), one poller sums 4 loads     3: sharded 8
// 4: volatile polls + fence after                            5: cluster-assisted (barrier.cluster, rank 0 adds, cluster release)
// 6: 32 pollers? no: thread 0 polls, but `__nanosleep(20)` between polls
template <int V>
__device__ __forceinline__ bool bar(unsigned int *b, unsigned int &epoch, unsigned int nctas) {
    bool ok = true;
    if (V == 5) {
        cg::cluster_group cl = cg::this_cluster();
        const unsigned int ncl = nctas / cl.num_blocks();
        cl.sync();                         // everybody in the cluster arrived (release/acquire at cluster scope)
        if (cl.block_rank() == 0 && threadIdx.x == 0) {
            epoch += ncl;
            __threadfence();
            atomicAdd(b, 1u);
            long spins = 0;
            while (ld_acquire(b) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
        } else if (threadIdx.x == 0) epoch += ncl;
        cl.sync();
        return ok;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long spins = 0;
        if (V == 0) {
            epoch += nctas;
            __threadfence();
            atomicAdd(b, 1u);
            while (ld_acquire(b) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
        } else if (V == 1) {
            epoch += nctas;
            __threadfence();
            red_relaxed(b, 1u);
            while (ld_relaxed(b) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
            __threadfence();
        } else if (V == 2 || V == 3) {
            constexpr unsigned int NS = V == 2 ? 4 : 8;
            epoch += nctas;
            __threadfence();
            red_relaxed(b + (blockIdx.x % NS) * 64, 1u);
            for (;;) {
                unsigned int s = 0;
#pragma unroll
                for (unsigned int i = 0; i < NS; ++i) s += ld_relaxed(b + i * 64);
                if (s >= epoch) break;
                if (++spins > SPIN_LIMIT) { ok = false; break; }
            }
            __threadfence();
        } else if (V == 4) {
            epoch += nctas;
            __threadfence();
            red_relaxed(b, 1u);
            while (ld_volatile(b) < epoch) if (++spins > SPIN_LIMIT) { ok = false; break; }
            __threadfence();
        } else if (V == 6) {
            epoch += nctas;
            __threadfence();
            red_relaxed(b, 1u);
            while (ld_relaxed(b) < epoch) { __nanosleep(32); if (++spins > SPIN_LIMIT) { ok = false; break; } }
            __threadfence();
        }
    }
    __syncthreads();
    return ok;
}

template <int V>
__global__ void __launch_bounds__(256, 1) bar_kernel(unsigned int *b, float *acc, int iters, int nred, int nload, long long *cycles, unsigned int *fail) {
    unsigned int epoch = 0;
    float sink = 0.f;
    const unsigned int n = gridDim.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        for (int i = threadIdx.x; i < nred; i += 256) atomicAdd(acc + ((i * 8 + (i >> 9)) & 4095), 1.0f);
        bool ok = bar<V>(b, epoch, n);
        if (!ok) { atomicExch(fail, 1u); break; }
        if (*reinterpret_cast<volatile unsigned int *>(fail)) break;
        for (int i = threadIdx.x; i < nload; i += 256) sink += __ldcg(acc + i);
    }
    long long t1 = clock64();
    if (sink == -1.f) acc[0] = sink;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// cluster.sync latency, DSMEM reduce-scatter
__global__ void __launch_bounds__(256, 1) cl_kernel(int iters, int mode, float *acc, long long *cycles) {
    extern __shared__ __align__(16) unsigned char smem[];
    float *sbuf = reinterpret_cast<float *>(smem);
    cg::cluster_group cl = cg::this_cluster();
    const unsigned int cs = cl.num_blocks(), cr = cl.block_rank();
    cl.sync();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) cl.sync();
        else {
            // every CTA sends `mode` floats to every peer, then each sums what it received
            for (int i = threadIdx.x; i < (int)cs * mode; i += 256) {
                int peer = i / mode, j = i % mode;
                *cl.map_shared_rank(sbuf + cr * mode + j, peer) = (float)it;
            }
            cl.sync();
            float s = 0.f;
            for (int j = threadIdx.x; j < mode; j += 256) for (unsigned int r = 0; r < cs; ++r) s += sbuf[r * mode + j];
            if (s == -1.f) acc[0] = s;
            cl.sync();
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int V>
static void run_bar(const char *name, int grid, int cs, int nred, int nload, unsigned int *b, float *acc, long long *cyc, unsigned int *fail) {
    const int iters = 4000;
    CK(cudaMemset(b, 0, 16384)); CK(cudaMemset(fail, 0, 4));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = 0;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (cs > 1) { at[na].id = cudaLaunchAttributeClusterDimension; at[na].val.clusterDim.x = cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1; ++na; }
    at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; ++na;
    cfg.attrs = at; cfg.numAttrs = na;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    cudaError_t e = cudaLaunchKernelEx(&cfg, bar_kernel<V>, b, acc, iters, nred, nload, cyc, fail);
    cudaEventRecord(e1);
    cudaError_t e2 = cudaDeviceSynchronize();
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    long long c = 0; unsigned int f = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&f, fail, 4, cudaMemcpyDeviceToHost);
    printf("  grid=%3d cs=%2d nred=%4d nload=%4d %-40s %.3f us (%lld cycles)%s %s\n", grid, cs, nred, nload, name, ms * 1e3f / iters, c / iters,
           f ? "  TIMED OUT" : "", (e != cudaSuccess || e2 != cudaSuccess) ? cudaGetErrorString(e != cudaSuccess ? e : e2) : "");
    cudaGetLastError();
}

int main() {
    setvbuf(stdout, NULL, _IONBF, 0);
    unsigned int *b, *fail; float *acc; long long *cyc;
    CK(cudaMalloc(&b, 16384)); CK(cudaMalloc(&acc, 4096 * 4)); CK(cudaMalloc(&cyc, 64)); CK(cudaMalloc(&fail, 4));
    CK(cudaMemset(b, 0, 16384)); CK(cudaMemset(acc, 0, 16384));
    for (int g : {1, 148}) {
        comp_kernel<<<g, 32>>>(b, cyc);
        CK(cudaDeviceSynchronize());
        long long o[8];
        cudaMemcpy(o, cyc, 64, cudaMemcpyDeviceToHost);
        printf("components (%d CTAs active, cycles): fence %lld | atomic RTT %lld | ld.acquire %lld | ld.relaxed %lld | ld.volatile %lld | st+fence %lld | red+fence %lld\n",
               g, o[0], o[1], o[2], o[3], o[4], o[5], o[6]);
    }
    for (int nred : {0, 256})
        for (int nload : {0, 4096}) {
            run_bar<0>("v0 fence+atomic+acquire spin", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<1>("v1 relaxed red / relaxed polls + fences", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<2>("v2 4 sharded counters", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<3>("v3 8 sharded counters", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<4>("v4 volatile polls", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<6>("v6 relaxed polls + nanosleep", 148, 1, nred, nload, b, acc, cyc, fail);
            run_bar<5>("v5 cluster-assisted cs=2", 148, 2, nred, nload, b, acc, cyc, fail);
            run_bar<5>("v5 cluster-assisted cs=4", 132, 4, nred, nload, b, acc, cyc, fail);
            run_bar<5>("v5 cluster-assisted cs=8", 120, 8, nred, nload, b, acc, cyc, fail);
            run_bar<0>("v0 at 120 CTAs", 120, 1, nred, nload, b, acc, cyc, fail);
        }
    CK(cudaFuncSetAttribute(cl_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    CK(cudaFuncSetAttribute(cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (int cs : {2, 4, 8, 16})
        for (int mode : {0, 64, 256, 1024}) {
            const int iters = 2000;
            const int grid = cs == 16 ? 112 : cs == 8 ? 120 : cs == 4 ? 132 : 148;
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 100 * 1024; cfg.stream = 0;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaError_t e = cudaLaunchKernelEx(&cfg, cl_kernel, iters, mode, acc, cyc);
            cudaError_t e2 = cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
            printf("  cluster size %2d grid %3d: %s: %lld cycles / iter %s\n", cs, grid,
                   mode == 0 ? "cluster.sync" : mode == 64 ? "all-to-all 256 B/peer + 2 syncs" : mode == 256 ? "all-to-all 1 KB/peer + 2 syncs" : "all-to-all 4 KB/peer + 2 syncs",
                   c / iters, (e != cudaSuccess || e2 != cudaSuccess) ? cudaGetErrorString(e != cudaSuccess ? e : e2) : "");
            cudaGetLastError();
        }
    return 0;
}
