// api_internal.cuh -- what the orchestration files (api.cu: inference + training forward, train.cu: the backward of the
// training step) share: the handle, error reporting, and the dispatch of one linear layer / one attention call to the
// kernel families.  Host code only.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <type_traits>
#include <vector>

#include "common.cuh"
#include "kernels.h"

extern thread_local std::string g_create_error;

struct pz_handle {
    pz_config cfg;
    pz_weights w;
    bool bound = false;
    std::vector<pz_vit_layer> vit;
    std::vector<pz_mix_layer> vlm, proprio, action;
    std::string err;
    LaunchCounter lc;
    int prefix_chunk = 64;
    int pixel_format = PZ_PIXELS_MODEL_DTYPE;   // pz_set_pixel_format
    // optional CUDA-event timing of one tagged kernel family (bench.py roofline)
    // fork/join side stream: the proprio token's chain of small kernels runs beside the VLM chain
    cudaStream_t side = nullptr;
    std::vector<cudaEvent_t> sync_ev;
    Mega3State mega3[2];                      // packed item streams of the persistent sampler for batch 1 and 2
    int num_sms = 0;
    int sampler = PZ_SAMPLER_AUTO;            // pz_set_sampler
    long long fallbacks = 0;                  // ops that ran on the SIMT kernels under PZ_FLAG_ALLOW_FALLBACK
    // optional caller-side normalisation folded into the path (pz_set_io_normalization): device vectors or null
    const float *prop_scale = nullptr, *prop_shift = nullptr, *act_scale = nullptr, *act_shift = nullptr;
    int prop_clip = 0;
    int timing_tag = 0;                       // 0 = off
    std::vector<cudaEvent_t> ev;              // pairs (start, stop)
    size_t ev_used = 0;
};

enum { TAG_VLM_GATE_UP = 1, TAG_VLM_DOWN = 2, TAG_ACT_GATE_UP = 3, TAG_VIT_FC1 = 4 };

static void tick(pz_handle *h, int tag, cudaStream_t st) {
    if (h->timing_tag != tag) return;
    if (h->ev_used == h->ev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        h->ev.push_back(e);
    }
    cudaEventRecord(h->ev[h->ev_used++], st);
}

static int fail(pz_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}


// bump allocator over one caller-owned workspace
struct Bump {
    size_t off = 0;
    char *base = nullptr;
    template <typename P> P *take(size_t bytes) {
        size_t o = (off + 1023) & ~(size_t)1023;   // 1 KiB alignment (TMA / vector loads)
        off = o + bytes;
        return base ? reinterpret_cast<P *>(base + o) : nullptr;
    }
};

// --------------------------------------------------------------- dispatch --
template <typename T> struct Ops;

template <> struct Ops<float> {
    static int linear(pz_handle *, const LinearArgs &a, cudaStream_t st) {
        launch_linear_simple<float>(a, st);
        return 0;
    }
    static int attention(pz_handle *, const AttnArgs &a, cudaStream_t st) {
        launch_attn_simple<float>(a, st);
        return 0;
    }
};

// bf16: every shape of the shipped configurations has a tensor-core / skinny kernel.  A shape none of them covers is an
// ERROR (a silent drop to the SIMT kernels costs ~50x), unless the caller asked for the SIMT kernels
// (PZ_FLAG_SIMPLE_KERNELS) or explicitly allowed the fallback (PZ_FLAG_ALLOW_FALLBACK; counted, pz_fallback_count).
template <> struct Ops<bf16> {
    static int linear(pz_handle *h, const LinearArgs &a, cudaStream_t st) {
        if (!(h->cfg.flags & PZ_FLAG_SIMPLE_KERNELS)) {
            if (skinny_supported(a)) return launch_linear_skinny(a, st);
            if (gemm_tc_supported(a)) {
                const char *e = nullptr;
                int rc = launch_linear_tc(a, st, &e);
                if (rc) return fail(h, rc, e ? e : "tcgen05 gemm launch failed");
                return 0;
            }
            if (!(h->cfg.flags & PZ_FLAG_ALLOW_FALLBACK)) {
                char msg[256];
                snprintf(msg, sizeof(msg), "no tensor-core kernel for linear M=%d N=%d K=%d lda=%d ldc=%d flags=0x%x "
                         "(set PZ_FLAG_ALLOW_FALLBACK / PZ_ALLOW_FALLBACK=1 to run it on the SIMT kernel)",
                         a.M, a.N, a.K, a.lda, a.ldc, a.flags);
                return fail(h, PZ_ERR_INVALID, msg);
            }
            ++h->fallbacks;
        }
        launch_linear_simple<bf16>(a, st);
        return 0;
    }
    static int attention(pz_handle *h, const AttnArgs &a, cudaStream_t st) {
        if (!(h->cfg.flags & PZ_FLAG_SIMPLE_KERNELS)) {
            if (attn_tc_supported(a)) return launch_attn_tc(a, st);          // prefix vlm rows: tcgen05
            if (attn_tc_vit_supported(a)) return launch_attn_tc_vit(a, st);  // SigLIP encoder: tcgen05
            if (attn_mma_supported(a)) return launch_attn_mma(a, st);
            if (!(h->cfg.flags & PZ_FLAG_ALLOW_FALLBACK))
                return fail(h, PZ_ERR_INVALID, "no tensor-core kernel for this attention shape (set PZ_FLAG_ALLOW_FALLBACK / "
                                               "PZ_ALLOW_FALLBACK=1 to run it on the SIMT kernel)");
            ++h->fallbacks;
        }
        launch_attn_simple<bf16>(a, st);
        return 0;
    }
};

static LinearArgs lin(const void *A, int lda, const void *W, const float *bias, void *C, int ldc,
                      int M, int N, int K, int flags = 0, float alpha = 1.f) {
    LinearArgs a;
    a.A = A; a.W = W; a.bias = bias; a.C = C;
    a.M = M; a.N = N; a.K = K; a.lda = lda; a.ldc = ldc;
    a.alpha = alpha; a.flags = flags; a.norm_w = nullptr; a.ldw = 0;
    a.cmb_splits = a.cmb_q_rows = a.cmb_heads = a.cmb_hd = 0;
    return a;
}

#define PZ_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

// y = linear(rmsnorm(x)): one kernel when the skinny path can normalise while it loads the
// activations (small M), otherwise the norm kernel writes `hbuf` and the GEMM reads it.
template <typename T>
static int norm_linear(pz_handle *h, const float *x, const float *norm_w, void *hbuf, LinearArgs a,
                       int hidden, cudaStream_t st) {
    if (std::is_same<T, bf16>::value && !(h->cfg.flags & PZ_FLAG_SIMPLE_KERNELS)) {
        LinearArgs f = a;
        f.A = x; f.lda = hidden; f.flags |= LIN_NORM_A; f.norm_w = norm_w;
        if (skinny_supported(f)) return launch_linear_skinny(f, st);
    }
    launch_rmsnorm<T>(x, norm_w, (T *)hbuf, a.M, hidden, 1e-6f, st);
    a.A = hbuf; a.lda = hidden;
    return Ops<T>::linear(h, a, st);
}

static void copy_f32(float *dst, const float *src, size_t n, cudaStream_t st) {
    cudaMemcpyAsync(dst, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st);
}

