// LRU complex diagonal linear recurrence  h_t = f_t * h_{t-1} + v_t  (complex, per-step decay,
// supplied initial state), forward and backward, as a single-pass tiled parallel scan.
//
// Replaces the reference's serial Triton kernels fwd_sequential_scan_complex /
// bwd_sequential_scan_complex (ref: offpolicy_rnn/models/lru/scan_triton/complex_rnn.py:43-87,
// 90-170; autograd wrapper :174-242).  Same tiling as scan_real.cu: one CTA per (batch row,
// 32 channels), tiles of 32*S steps staged with cp.async, chunk aggregates (A, H) in C combined by
// warp shuffles + an 8-entry carry chain in shared memory.
//
// Backward (complex_rnn.py:126-170), with gd_t = grad_detach[b,t]:
//   G_t = g_t + (1 - gd_t) * conj(f_{t+1}) * G_{t+1},   dv_t = G_t,   df_t = G_t * conj(h_{t-1}),
// h_{-1} = the supplied initial state; no gradient flows to the initial state (:242).
#include "common.cuh"

namespace rorl {

constexpr int kCThreads = 256;

struct F4c {
    float v[4];
};
__device__ __forceinline__ F4c cld4(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    return F4c{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void cst4(float* p, const F4c& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
struct C4 {
    F4c re, im;
};
__device__ __forceinline__ C4 cshfl_up(const C4& a, int off) {
    C4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.re.v[i] = __shfl_up_sync(0xffffffffu, a.re.v[i], off);
        r.im.v[i] = __shfl_up_sync(0xffffffffu, a.im.v[i], off);
    }
    return r;
}
// later ∘ earlier for the affine maps s -> A s + H
__device__ __forceinline__ void ccombine(C4& A, C4& H, const C4& Ap, const C4& Hp) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float hr = A.re.v[i] * Hp.re.v[i] - A.im.v[i] * Hp.im.v[i] + H.re.v[i];
        float hi = A.re.v[i] * Hp.im.v[i] + A.im.v[i] * Hp.re.v[i] + H.im.v[i];
        float ar = A.re.v[i] * Ap.re.v[i] - A.im.v[i] * Ap.im.v[i];
        float ai = A.re.v[i] * Ap.im.v[i] + A.im.v[i] * Ap.re.v[i];
        H.re.v[i] = hr; H.im.v[i] = hi; A.re.v[i] = ar; A.im.v[i] = ai;
    }
}
__device__ __forceinline__ C4 capply(const C4& A, const C4& H, const C4& s) {
    C4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.re.v[i] = A.re.v[i] * s.re.v[i] - A.im.v[i] * s.im.v[i] + H.re.v[i];
        r.im.v[i] = A.re.v[i] * s.im.v[i] + A.im.v[i] * s.re.v[i] + H.im.v[i];
    }
    return r;
}

// wagg: [2 parities][8 warps][8 quads][16 floats] = A.re, A.im, H.re, H.im
__device__ __forceinline__ C4 ctile_carry_in(C4 A, C4 H, C4& carry, float* wagg, int par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cq = threadIdx.x & 7;
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        C4 Ap = cshfl_up(A, off), Hp = cshfl_up(H, off);
        if (lane >= off) ccombine(A, H, Ap, Hp);
    }
    C4 Ae = cshfl_up(A, 8), He = cshfl_up(H, 8);
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { Ae.re.v[i] = 1.f; Ae.im.v[i] = 0.f; He.re.v[i] = 0.f; He.im.v[i] = 0.f; }
    }
    float* base = wagg + (size_t)par * (8 * 8 * 16);
    if (lane >= 24) {
        float* p = base + (warp * 8 + cq) * 16;
        cst4(p, A.re); cst4(p + 4, A.im); cst4(p + 8, H.re); cst4(p + 12, H.im);
    }
    __syncthreads();
    C4 s = carry, s_in = carry;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const float* p = base + (w * 8 + cq) * 16;
        C4 Aw{cld4(p), cld4(p + 4)}, Hw{cld4(p + 8), cld4(p + 12)};
        if (w == warp) s_in = s;
        s = capply(Aw, Hw, s);
    }
    carry = s;
    return capply(Ae, He, s_in);
}

template <int S, int NST>
__global__ void __launch_bounds__(kCThreads) lru_fwd_kernel(
    const float* __restrict__ gvr, const float* __restrict__ gvi, const float* __restrict__ gfr,
    const float* __restrict__ gfi, const float* __restrict__ gh0r, const float* __restrict__ gh0i,
    float* __restrict__ ghr, float* __restrict__ ghi, int L, int C) {
    constexpr int TL = 32 * S, ARR = TL * 32, STAGE = 4 * ARR;
    extern __shared__ __align__(16) float smem[];
    float* wagg = smem + NST * STAGE;
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int tid = threadIdx.x, cq = tid & 7, ck = tid >> 3;
    const bool cvalid = (c0 + cq * 4) < C;
    const size_t rowbase = (size_t)b * L;
    const int ntiles = (L + TL - 1) / TL;

    auto issue = [&](int tile) {
        if (tile < ntiles) {
            float* st = smem + (tile % NST) * STAGE;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                int q = ck + 32 * i, t = tile * TL + q;
                bool ok = cvalid && t < L;
                size_t g = (rowbase + (ok ? t : 0)) * C + c0 + cq * 4;
                float* d = st + q * 32 + cq * 4;
                cp_async16(d, gvr + g, ok);
                cp_async16(d + ARR, gvi + g, ok);
                cp_async16(d + 2 * ARR, gfr + g, ok);
                cp_async16(d + 3 * ARR, gfi + g, ok);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) issue(i);

    C4 carry;
    {
        size_t g = (size_t)b * C + c0 + cq * 4;
        if (cvalid && gh0r) { carry.re = cld4(gh0r + g); carry.im = cld4(gh0i + g); }
        else {
#pragma unroll
            for (int i = 0; i < 4; ++i) { carry.re.v[i] = 0.f; carry.im.v[i] = 0.f; }
        }
    }
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        issue(tile + NST - 1);
        float* sp = smem + (tile % NST) * STAGE + (ck * S) * 32 + cq * 4;
        const int t0 = tile * TL + ck * S;
        C4 A, H;
#pragma unroll
        for (int i = 0; i < 4; ++i) { A.re.v[i] = 1.f; A.im.v[i] = 0.f; H.re.v[i] = 0.f; H.im.v[i] = 0.f; }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                C4 v{cld4(sp + s * 32), cld4(sp + ARR + s * 32)};
                C4 f{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
                ccombine(f, v, A, H);  // (f, v) ∘ (A, H)
                A = f; H = v;
            }
        }
        C4 h = ctile_carry_in(A, H, carry, wagg, tile & 1);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                C4 v{cld4(sp + s * 32), cld4(sp + ARR + s * 32)};
                C4 f{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
                h = capply(f, v, h);
                if (cvalid) {
                    size_t g = (rowbase + t0 + s) * C + c0 + cq * 4;
                    cst4(ghr + g, h.re);
                    cst4(ghi + g, h.im);
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <int S, int NST>
__global__ void __launch_bounds__(kCThreads) lru_bwd_kernel(
    const float* __restrict__ ggr, const float* __restrict__ ggi, const float* __restrict__ gfr,
    const float* __restrict__ gfi, const float* __restrict__ ghr, const float* __restrict__ ghi,
    const float* __restrict__ gh0r, const float* __restrict__ gh0i, const float* __restrict__ ggd,
    float* __restrict__ gdvr, float* __restrict__ gdvi, float* __restrict__ gdfr, float* __restrict__ gdfi,
    int L, int C) {
    constexpr int TL = 32 * S, ARR = TL * 32, STAGE = 6 * ARR + TL;  // g, f, hprev (re/im), gd
    extern __shared__ __align__(16) float smem[];
    float* wagg = smem + NST * STAGE;
    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int tid = threadIdx.x, cq = tid & 7, ck = tid >> 3;
    const bool cvalid = (c0 + cq * 4) < C;
    const size_t rowbase = (size_t)b * L;
    const int ntiles = (L + TL - 1) / TL;

    auto issue = [&](int tile) {
        if (tile < ntiles) {
            float* st = smem + (tile % NST) * STAGE;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                int q = ck + 32 * i, r = tile * TL + q, t = L - 1 - r;
                bool ok = cvalid && r < L;
                size_t g = (rowbase + (ok ? t : 0)) * C + c0 + cq * 4;
                float* d = st + q * 32 + cq * 4;
                cp_async16(d, ggr + g, ok);
                cp_async16(d + ARR, ggi + g, ok);
                cp_async16(d + 2 * ARR, gfr + g, ok);
                cp_async16(d + 3 * ARR, gfi + g, ok);
                // h_{t-1}: previous output row, or the supplied initial state at t == 0
                bool okp = ok && t > 0;
                bool ok0 = ok && t == 0 && gh0r != nullptr;
                size_t g0 = (size_t)b * C + c0 + cq * 4;
                const float* pr = okp ? ghr + (g - C) : (ok0 ? gh0r + g0 : ghr);
                const float* pi = okp ? ghi + (g - C) : (ok0 ? gh0i + g0 : ghi);
                cp_async16(d + 4 * ARR, pr, okp || ok0);
                cp_async16(d + 5 * ARR, pi, okp || ok0);
            }
            if (tid < TL) {
                int r = tile * TL + tid;
                bool ok = (ggd != nullptr) && r < L;
                cp_async4(st + 6 * ARR + tid, ggd + (ok ? rowbase + (L - 1 - r) : 0), ok);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) issue(i);

    C4 carry;
#pragma unroll
    for (int i = 0; i < 4; ++i) { carry.re.v[i] = 0.f; carry.im.v[i] = 0.f; }
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<NST - 2>();
        __syncthreads();
        issue(tile + NST - 1);
        float* st = smem + (tile % NST) * STAGE;
        float* sp = st + (ck * S) * 32 + cq * 4;
        const int r0 = tile * TL + ck * S;
        C4 A, H;
#pragma unroll
        for (int i = 0; i < 4; ++i) { A.re.v[i] = 1.f; A.im.v[i] = 0.f; H.re.v[i] = 0.f; H.im.v[i] = 0.f; }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (r0 + s < L) {
                C4 g{cld4(sp + s * 32), cld4(sp + ARR + s * 32)};
                C4 f{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
                float keep = 1.0f - st[6 * ARR + ck * S + s];
                // E_r = conj(f) * (g + keep * E_{r-1})  ->  A' = conj(f)*keep, H' = conj(f)*g
                C4 a, h;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a.re.v[i] = f.re.v[i] * keep;
                    a.im.v[i] = -f.im.v[i] * keep;
                    h.re.v[i] = f.re.v[i] * g.re.v[i] + f.im.v[i] * g.im.v[i];
                    h.im.v[i] = f.re.v[i] * g.im.v[i] - f.im.v[i] * g.re.v[i];
                }
                ccombine(a, h, A, H);
                A = a; H = h;
            }
        }
        C4 E = ctile_carry_in(A, H, carry, wagg, tile & 1);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            int r = r0 + s;
            if (r < L) {
                C4 g{cld4(sp + s * 32), cld4(sp + ARR + s * 32)};
                C4 f{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
                C4 hp{cld4(sp + 4 * ARR + s * 32), cld4(sp + 5 * ARR + s * 32)};
                float keep = 1.0f - st[6 * ARR + ck * S + s];
                C4 G, df;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    G.re.v[i] = g.re.v[i] + keep * E.re.v[i];
                    G.im.v[i] = g.im.v[i] + keep * E.im.v[i];
                    df.re.v[i] = G.re.v[i] * hp.re.v[i] + G.im.v[i] * hp.im.v[i];
                    df.im.v[i] = G.im.v[i] * hp.re.v[i] - G.re.v[i] * hp.im.v[i];
                    E.re.v[i] = G.re.v[i] * f.re.v[i] + G.im.v[i] * f.im.v[i];
                    E.im.v[i] = G.im.v[i] * f.re.v[i] - G.re.v[i] * f.im.v[i];
                }
                if (cvalid) {
                    size_t o = (rowbase + (L - 1 - r)) * C + c0 + cq * 4;
                    cst4(gdvr + o, G.re); cst4(gdvi + o, G.im);
                    cst4(gdfr + o, df.re); cst4(gdfi + o, df.im);
                }
            }
        }
    }
    cp_async_wait<0>();
}

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_lru_scan_fwd(const float* v_re, const float* v_im, const float* f_re, const float* f_im,
                      const float* h0_re, const float* h0_im, float* h_re, float* h_im, int64_t B, int64_t L,
                      int64_t C, cudaStream_t stream) {
    if (!v_re || !v_im || !f_re || !f_im || !h_re || !h_im) return RORL_ERR_ARG;
    if ((h0_re == nullptr) != (h0_im == nullptr)) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || C <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (C % 4) return RORL_ERR_ALIGN;
    constexpr int S = 2, NST = 3;
    auto kern = lru_fwd_kernel<S, NST>;
    constexpr size_t smem = sizeof(float) * (NST * (4 * 32 * S * 32) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(v_re, v_im, f_re, f_im, h0_re, h0_im, h_re, h_im, (int)L, (int)C);
    RORL_RETURN_LAUNCH();
}

int rorl_lru_scan_bwd(const float* g_re, const float* g_im, const float* f_re, const float* f_im,
                      const float* h_re, const float* h_im, const float* h0_re, const float* h0_im,
                      const float* grad_detach, float* dv_re, float* dv_im, float* df_re, float* df_im,
                      int64_t B, int64_t L, int64_t C, cudaStream_t stream) {
    if (!g_re || !g_im || !f_re || !f_im || !h_re || !h_im || !dv_re || !dv_im || !df_re || !df_im)
        return RORL_ERR_ARG;
    if ((h0_re == nullptr) != (h0_im == nullptr)) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || C <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (C % 4) return RORL_ERR_ALIGN;
    constexpr int S = 2, NST = 2;
    auto kern = lru_bwd_kernel<S, NST>;
    constexpr size_t smem = sizeof(float) * (NST * (6 * 32 * S * 32 + 32 * S) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(g_re, g_im, f_re, f_im, h_re, h_im, h0_re, h0_im, grad_detach, dv_re,
                                            dv_im, df_re, df_im, (int)L, (int)C);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
