// probe_ll.cu -- sizes the building blocks of the flag-exchange sampler (denoise_mega3.cu) on 148 co-resident CTAs:
//   stream : one producer warp per CTA feeding an mbarrier ring with contiguous cp.async.bulk copies (GB/s)
//   gather : all-gather exchange through 64-bit {payload, sequence} words in L2 (us per exchange)
//   stage  : both together -- every stage consumes `ipc` ring items, publishes its words and gathers everybody's
//   rs     : reduce-scatter exchange (every CTA publishes W words, reducer r collects its slice of all partials)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/probe_ll tools/probe_ll.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
#define DEVINL __device__ __forceinline__

constexpr int NCW = 8, NCT = NCW * 32, NT = NCT + 32;
constexpr long SPIN_LIMIT = 1L << 20;

DEVINL uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEVINL void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
DEVINL void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
DEVINL void mbar_arrive(uint64_t *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory"); }
DEVINL bool mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
DEVINL bool mbar_wait(uint64_t *bar, uint32_t parity) {
    long spins = 0;
    while (!mbar_try(bar, parity)) if (++spins > SPIN_LIMIT) return false;
    return true;
}
DEVINL void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
DEVINL uint64_t policy_evict_first() { uint64_t pol; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol)); return pol; }
DEVINL void ll_store(unsigned long long *dst, uint32_t payload, uint32_t flag) {
    unsigned long long v = ((unsigned long long)flag << 32) | payload;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(dst), "l"(v) : "memory");
}
DEVINL void ll_load2(const unsigned long long *src, unsigned long long &a, unsigned long long &b) {
    asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
}
DEVINL unsigned long long ll_load1(const unsigned long long *src) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(src) : "memory");
    return v;
}
DEVINL void bar_compute() { asm volatile("bar.sync 1, %0;" ::"n"(NCT) : "memory"); }

struct Params {
    const uint8_t *stream; long cta_stride; int items_per_pass; int chunk; int slots;
    unsigned long long *ll[2]; int words;        // all-gather exchange buffers (double buffered), words per exchange
    unsigned long long *rs[2];                   // reduce-scatter buffers [reducer][producer][wpr]
    int n_stages, ipc_num, ipc_den, mode;        // items per stage = ipc_num / ipc_den (fractional: accumulated)
    int sentinel;                                // 1: one thread polls a sentinel word before the gather
    int backoff_ns;                              // sleep between polling rounds
    int max_inflight;                            // stream: bulk copies in flight per CTA (0: as many as there are free slots)
    int replicas;                                // all-gather: every producer writes R copies, consumer c reads copy c % R
    unsigned int *err; float *sink;
    unsigned int *counter; float *cdata[2];      // mode 4 / 5: arrival counter + compact fp32 payload (double buffered)
};

extern __shared__ __align__(1024) uint8_t smem[];

// mode 0: stream only; 1: gather only; 2: stage (stream + gather); 3: reduce-scatter + all-gather per stage (+ stream)
__global__ void __launch_bounds__(NT, 1) probe_kernel(const __grid_constant__ Params p) {
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)p.slots * p.chunk), *empty = full + p.slots;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < p.slots; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // total items this CTA consumes
    long total_items = 0;
    if (p.mode == 0) total_items = (long)p.n_stages;
    else if (p.mode >= 2) total_items = ((long)p.n_stages * p.ipc_num) / p.ipc_den;
    if (warp == NCW) {   // producer warp
        if (lane == 0 && total_items > 0) {
            const uint64_t pol = policy_evict_first();
            const uint8_t *base = p.stream + (long)blockIdx.x * p.cta_stride;
            for (long i = 0; i < total_items; ++i) {
                const int slot = (int)(i % p.slots);
                if (i >= p.slots && !mbar_wait(&empty[slot], (uint32_t)((i / p.slots) - 1) & 1)) { atomicExch(p.err, 2u); break; }
                if (p.max_inflight > 0 && i >= p.max_inflight) {   // copy i - max_inflight must have landed
                    const long k = i - p.max_inflight;
                    if (!mbar_wait(&full[k % p.slots], (uint32_t)(k / p.slots) & 1)) { atomicExch(p.err, 8u); break; }
                }
                mbar_expect_tx(&full[slot], p.chunk);
                bulk_g2s(smem + (size_t)slot * p.chunk, base + (i % p.items_per_pass) * (long)p.chunk, p.chunk, &full[slot], pol);
            }
        }
        return;
    }
    float acc = 0.f;
    long cnt = 0;
    auto consume = [&](int n) {
        for (int j = 0; j < n; ++j, ++cnt) {
            const int slot = (int)(cnt % p.slots);
            if (!mbar_wait(&full[slot], (uint32_t)(cnt / p.slots) & 1)) { atomicExch(p.err, 1u); return; }
            const uint4 *src = reinterpret_cast<const uint4 *>(smem + (size_t)slot * p.chunk);
            for (int i = tid; i < p.chunk / 16; i += NCT) { uint4 v = src[i]; acc += __uint_as_float(v.x ^ v.y ^ v.z ^ v.w); }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
        }
    };
    if (p.mode == 0) { consume((int)total_items); if (acc == 123.456f) p.sink[0] = acc; return; }

    const int G = gridDim.x, c = blockIdx.x;
    const int wpc = (p.words + G - 1) / G;   // words each CTA publishes per all-gather
    long items_done = 0;
    for (int st = 0; st < p.n_stages; ++st) {
        const uint32_t seq = (uint32_t)st + 1;
        unsigned long long *buf = p.ll[st & 1];
        if (p.mode >= 2) {
            long want = ((long)(st + 1) * p.ipc_num) / p.ipc_den;
            consume((int)(want - items_done));
            items_done = want;
        }
        if (p.mode == 3) {
            // reduce-scatter: publish `words` partial values, slice r of them to reducer r: rs[r][c][wpc]
            unsigned long long *rb = p.rs[st & 1];
            for (int i = tid; i < G * wpc; i += NCT) {   // (padded: every reducer gets wpc words from every producer)
                int r = i / wpc, j = i % wpc;
                ll_store(rb + ((long)r * G + c) * wpc + j, __float_as_uint(acc + i), seq);
            }
            // reducer: gather G * wpc words (contiguous), sum over producers
            const unsigned long long *mine = rb + (long)c * G * wpc;
            const int n2 = (G * wpc) / 2;   // double words
            float part = 0.f;
            for (int i0 = 0; i0 < n2; i0 += NCT * 4) {
                unsigned long long v[8];
                long spins = 0;
                for (;;) {
                    uint32_t diff = 0;
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        int i = i0 + u * NCT + tid;
                        if (i < n2) { ll_load2(mine + 2 * i, v[2 * u], v[2 * u + 1]); diff |= ((uint32_t)(v[2 * u] >> 32) ^ seq) | ((uint32_t)(v[2 * u + 1] >> 32) ^ seq); }
                    }
                    if (!diff) break;
                    if (++spins > SPIN_LIMIT) { atomicExch(p.err, 4u); break; }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { int i = i0 + u * NCT + tid; if (i < n2) part += __uint_as_float((uint32_t)v[2 * u]) + __uint_as_float((uint32_t)v[2 * u + 1]); }
            }
            // (a real reducer needs a cross-thread sum here: one shuffle tree + shared-memory pass)
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            float *red = reinterpret_cast<float *>(smem + (size_t)p.slots * p.chunk + 256);
            if (lane == 0) red[warp] = part;
            bar_compute();
            acc += red[0] + red[1] + red[2] + red[3] + red[4] + red[5] + red[6] + red[7];
            bar_compute();
        }
        if (p.mode == 4 || p.mode == 5) {
            // counter exchange: compact payload with plain stores, one release-increment per CTA; one thread polls the
            // counter, then everybody reads the payload once (mode 5: one cp.async.bulk into shared memory instead)
            float *cd = p.cdata[st & 1];
            for (int j = tid; j < wpc; j += NCT) { int w = c * wpc + j; if (w < p.words) cd[w] = acc + j; }
            bar_compute();
            if (tid == 0) {
                __threadfence();
                atomicAdd(p.counter, 1u);
                const unsigned int target = (unsigned int)G * seq;
                long spins = 0;
                unsigned int v;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.counter) : "memory");
                    if (++spins > SPIN_LIMIT) { atomicExch(p.err, 6u); break; }
                } while (v < target);
            }
            bar_compute();
            float part = 0.f;
            if (p.mode == 4) {
                const float4 *src = reinterpret_cast<const float4 *>(cd);
                for (int i = tid; i < p.words / 4; i += NCT) {
                    float4 v;
                    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
                    part += v.x * 1e-9f;
                }
            } else {
                // one thread: bulk copy of the payload into the ring's first bytes beyond the slots (reuse the staging area)
                float *dst = reinterpret_cast<float *>(smem + (size_t)p.slots * p.chunk + 4096);
                uint64_t *xbar = full + 2 * p.slots;
                if (tid == 0) {
                    if (st == 0) { mbar_init(xbar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
                    asm volatile("fence.proxy.async;" ::: "memory");
                    mbar_expect_tx(xbar, (uint32_t)p.words * 4);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(dst)), "l"(cd), "r"((uint32_t)p.words * 4), "r"(smem_u32(xbar)) : "memory");
                    if (!mbar_wait(xbar, (uint32_t)st & 1)) atomicExch(p.err, 7u);
                }
                bar_compute();
                for (int i = tid; i < p.words; i += NCT) part += dst[i] * 1e-9f;
            }
            float *stage = reinterpret_cast<float *>(smem + (size_t)p.slots * p.chunk + 512);
            stage[tid] = part;
            bar_compute();
            acc = acc * 0.5f + stage[(tid + 1) % NCT];
            bar_compute();
            if (*reinterpret_cast<volatile unsigned int *>(p.err)) break;
            continue;
        }
        // all-gather: publish my words, then collect everybody's
        const int R = p.replicas > 0 ? p.replicas : 1;
        for (int j = tid; j < wpc * R; j += NCT) { int w = c * wpc + j % wpc; if (w < p.words) ll_store(buf + (long)(j / wpc) * p.words + w, __float_as_uint(acc + j % wpc), seq); }
        buf += (long)(c % R) * p.words;
        if (p.sentinel) {
            if (tid == 0) {
                int sw = ((c + G / 2) % G) * wpc;
                if (sw >= p.words) sw = p.words - 1;
                const unsigned long long *s = buf + sw;
                long spins = 0;
                while ((uint32_t)(ll_load1(s) >> 32) != seq) {
                    if (++spins > SPIN_LIMIT) { atomicExch(p.err, 3u); break; }
                    if (p.backoff_ns) __nanosleep(p.backoff_ns);
                }
            }
            bar_compute();
        }
        const int n2 = p.words / 2;
        float part = 0.f;
        for (int i0 = 0; i0 < n2; i0 += NCT * 8) {
            unsigned long long v[16];
            long spins = 0;
            for (;;) {
                uint32_t diff = 0;
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    int i = i0 + u * NCT + tid;
                    if (i < n2) { ll_load2(buf + 2 * i, v[2 * u], v[2 * u + 1]); diff |= ((uint32_t)(v[2 * u] >> 32) ^ seq) | ((uint32_t)(v[2 * u + 1] >> 32) ^ seq); }
                }
                if (!diff) break;
                if (++spins > SPIN_LIMIT) { atomicExch(p.err, 5u); break; }
                if (p.backoff_ns) __nanosleep(p.backoff_ns);
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) { int i = i0 + u * NCT + tid; if (i < n2) part += __uint_as_float((uint32_t)v[2 * u]) * 1e-9f; }
        }
        // the consumer stages what it gathered into shared memory and synchronises before using it
        float *stage = reinterpret_cast<float *>(smem + (size_t)p.slots * p.chunk + 512);
        stage[tid] = part;
        bar_compute();
        acc = acc * 0.5f + stage[(tid + 1) % NCT];
        bar_compute();
        if (*reinterpret_cast<volatile unsigned int *>(p.err)) break;
    }
    if (acc == 123.456f) p.sink[0] = acc;
}

static float run(Params p, int grid, size_t smem_bytes, const char *label, double bytes_per_cta, int reps = 3) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaMemset(p.err, 0, 4));
        if (p.counter) CK(cudaMemset(p.counter, 0, 4));
        for (int b = 0; b < 2; ++b) {
            if (p.ll[b]) CK(cudaMemset(p.ll[b], 0, (size_t)p.words * 8 * (p.replicas > 0 ? p.replicas : 1) + 64));
            if (p.rs[b]) CK(cudaMemset(p.rs[b], 0, (size_t)grid * grid * ((p.words + grid - 1) / grid) * 8 + 64));
        }
        void *args[] = {&p};
        CK(cudaEventRecord(e0));
        CK(cudaLaunchCooperativeKernel((const void *)probe_kernel, dim3(grid), dim3(NT), args, smem_bytes, 0));
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    unsigned int err;
    CK(cudaMemcpy(&err, p.err, 4, cudaMemcpyDeviceToHost));
    printf("%-58s %9.3f ms  %8.3f us/stage", label, best, best * 1e3 / p.n_stages);
    if (bytes_per_cta > 0) printf("  %8.1f GB/s", bytes_per_cta * grid / (best * 1e-3) / 1e9);
    if (err) printf("  ERROR %u", err);
    printf("\n");
    return best;
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaSetDevice(dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    printf("SMs %d\n", sms);
    const int grid = sms;
    const size_t stream_bytes = (size_t)768 << 20;
    uint8_t *stream;
    CK(cudaMalloc(&stream, stream_bytes));
    CK(cudaMemset(stream, 1, stream_bytes));
    Params p;
    memset(&p, 0, sizeof(p));
    p.stream = stream;
    CK(cudaMalloc(&p.err, 64));
    CK(cudaMalloc(&p.sink, 64));
    const int maxwords = 16384;
    CK(cudaMalloc(&p.counter, 64));
    for (int b = 0; b < 2; ++b) CK(cudaMalloc(&p.cdata[b], (size_t)maxwords * 4 + 64));
    for (int b = 0; b < 2; ++b) {
        CK(cudaMalloc(&p.ll[b], (size_t)maxwords * 8 * 8 + 64));
        CK(cudaMalloc(&p.rs[b], (size_t)grid * grid * ((maxwords + grid - 1) / grid) * 8 + 64));
    }
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    char label[256];

    // ---- stream only: chunk size x ring depth
    const int cfgs[][2] = {{32768, 5}, {32768, 4}};
    for (auto &cf : cfgs) {
        p.mode = 0; p.chunk = cf[0]; p.slots = cf[1];
        p.cta_stride = (long)(stream_bytes / grid) & ~(long)1023;
        p.items_per_pass = (int)(p.cta_stride / p.chunk);
        p.n_stages = (int)(((size_t)6400 << 20) / grid / p.chunk);   // 6.4 GB in total
        size_t sm = (size_t)p.chunk * p.slots + 4096;
        snprintf(label, sizeof(label), "stream chunk %d ring %d", p.chunk, p.slots);
        run(p, grid, sm, label, (double)p.n_stages * p.chunk);
    }
    // ---- stream alone and stage = stream + all-gather (93 stages per step, 629 MB per step => 148 CTAs x 1.43 items of
    //      32 KB per stage) against the number of bulk copies in flight per CTA
    p.chunk = 32768; p.slots = 5;
    size_t sm = (size_t)p.chunk * p.slots + 4096;
    for (int chunk : {32768, 16384}) {
        for (int mi : {0, 1, 2, 3}) {
            p.chunk = chunk; p.slots = 5 * 32768 / chunk; p.max_inflight = mi;
            p.mode = 0; p.cta_stride = (long)(stream_bytes / grid) & ~(long)1023; p.items_per_pass = (int)(p.cta_stride / p.chunk);
            p.n_stages = (int)(((size_t)3200 << 20) / grid / p.chunk);
            snprintf(label, sizeof(label), "stream chunk %d ring %d in flight <= %d", p.chunk, p.slots, mi);
            run(p, grid, sm, label, (double)p.n_stages * p.chunk);
            for (int words : {1024, 4096}) {
                p.mode = 2; p.words = words; p.sentinel = 1; p.backoff_ns = 0; p.replicas = 1; p.n_stages = 930;
                p.ipc_num = 143 * (32768 / chunk); p.ipc_den = 100;
                snprintf(label, sizeof(label), "  + all-gather %d words (46 KB/stage), in flight <= %d", words, mi);
                run(p, grid, sm, label, (double)((long)p.n_stages * p.ipc_num / p.ipc_den) * p.chunk);
            }
        }
    }
    p.chunk = 32768; p.slots = 5; p.max_inflight = 0;
    p.sentinel = 0; p.backoff_ns = 0; p.replicas = 1;
    return 0;
    // ---- counter exchange (compact payload): without and with the weight stream
    p.slots = 4;
    const size_t smx = (size_t)p.chunk * p.slots + 4096 + 36864;
    for (int mode : {4, 5}) {
        for (int words : {2048, 4096, 8192}) {
            p.mode = mode; p.words = words; p.sentinel = 0; p.n_stages = 2000; p.ipc_num = 0; p.ipc_den = 1;
            snprintf(label, sizeof(label), "counter exchange mode %d, %d fp32 (%d KB)", mode, words, words * 4 / 1024);
            run(p, grid, smx, label, 0);
        }
        p.words = 4096; p.n_stages = 930; p.ipc_num = 143; p.ipc_den = 100;
        snprintf(label, sizeof(label), "stage: 1.43 items + counter exchange mode %d 4096 fp32", mode);
        run(p, grid, smx, label, (double)((long)p.n_stages * p.ipc_num / p.ipc_den) * p.chunk);
    }
    p.slots = 5;
    for (int words : {2048, 4096}) {   // the LL all-gather with the stream, for comparison at equal payload
        p.mode = 2; p.words = words; p.sentinel = 0; p.n_stages = 930; p.ipc_num = 143; p.ipc_den = 100;
        snprintf(label, sizeof(label), "stage: 1.43 items + LL all-gather %d words", words);
        run(p, grid, sm, label, (double)((long)p.n_stages * p.ipc_num / p.ipc_den) * p.chunk);
    }
    // ---- reduce-scatter + all-gather per stage, without and with the weight stream
    p.sentinel = 0; p.words = 4096; p.n_stages = 930;
    p.mode = 3; p.ipc_num = 0; p.ipc_den = 1;
    run(p, grid, sm, "reduce-scatter + all-gather 4096 words, no stream", 0);
    p.ipc_num = 286; p.ipc_den = 100;
    run(p, grid, sm, "reduce-scatter + all-gather 4096 words + 2.86 items", (double)((long)p.n_stages * p.ipc_num / p.ipc_den) * p.chunk);
    return 0;
}
