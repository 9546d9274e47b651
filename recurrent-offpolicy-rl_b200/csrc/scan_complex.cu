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

#ifndef RORL_LRU_FWD_S
#define RORL_LRU_FWD_S 4          // steps per thread and tile of the fused forward: 128-step tiles (8 per 1002-step row) halve the
                                  // block barriers and carry chains of 64-step tiles: 94 -> 56 us; 8 would not fit two CTAs per SM
#endif

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

// FUSED: the layer's own parameterisation instead of materialised per-step tensors (ref: lru/lru.py:95-115 builds
// gamma * u and lambda * (1 - start) as four [B, L, C] tensors before its scan): gvr / gvi are the raw in_proj outputs
// u, gfr / gfi are lambda as [C] vectors, ggam is gamma [C], gstart the reset flags [B, L] (may be NULL).  HBM traffic
// is then the algorithmic minimum of SURVEY.md 8(d): 16 B per element forward, 24 B backward.
template <int S, int NST, bool FUSED>
__global__ void __launch_bounds__(kCThreads) lru_fwd_kernel(
    const float* __restrict__ gvr, const float* __restrict__ gvi, const float* __restrict__ gfr,
    const float* __restrict__ gfi, const float* __restrict__ gh0r, const float* __restrict__ gh0i,
    float* __restrict__ ghr, float* __restrict__ ghi, int L, int C, const float* __restrict__ ggam,
    const float* __restrict__ gstart) {
    constexpr int TL = 32 * S, ARR = TL * 32, STAGE = FUSED ? 2 * ARR + TL : 4 * ARR;
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
                if (!FUSED) {
                    cp_async16(d + 2 * ARR, gfr + g, ok);
                    cp_async16(d + 3 * ARR, gfi + g, ok);
                }
            }
            if (FUSED && tid < TL) {
                const int t = tile * TL + tid;
                const bool ok = gstart != nullptr && t < L;
                cp_async4(st + 2 * ARR + tid, gstart + (ok ? rowbase + t : 0), ok);      // zero-filled = no reset
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) issue(i);

    C4 lam;
    F4c gam;
#pragma unroll
    for (int i = 0; i < 4; ++i) { lam.re.v[i] = 0.f; lam.im.v[i] = 0.f; gam.v[i] = 0.f; }
    if (FUSED && cvalid) {
        lam.re = cld4(gfr + c0 + cq * 4); lam.im = cld4(gfi + c0 + cq * 4); gam = cld4(ggam + c0 + cq * 4);
    }
    // (v_t, f_t) of step s of this thread's run, from the staged tile
    auto load_vf = [&](const float* st, const float* sp, int s, C4& v, C4& f) {
        v = C4{cld4(sp + s * 32), cld4(sp + ARR + s * 32)};
        if (FUSED) {
            const float keep = 1.0f - st[2 * ARR + ck * S + s];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v.re.v[i] *= gam.v[i]; v.im.v[i] *= gam.v[i];
                f.re.v[i] = lam.re.v[i] * keep; f.im.v[i] = lam.im.v[i] * keep;
            }
        } else {
            f = C4{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
        }
    };

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
        const float* stg = smem + (tile % NST) * STAGE;
        const float* sp = stg + (ck * S) * 32 + cq * 4;
        const int t0 = tile * TL + ck * S;
        C4 A, H;
#pragma unroll
        for (int i = 0; i < 4; ++i) { A.re.v[i] = 1.f; A.im.v[i] = 0.f; H.re.v[i] = 0.f; H.im.v[i] = 0.f; }
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                C4 v, f;
                load_vf(stg, sp, s, v, f);
                ccombine(f, v, A, H);  // (f, v) ∘ (A, H)
                A = f; H = v;
            }
        }
        C4 h = ctile_carry_in(A, H, carry, wagg, tile & 1);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                C4 v, f;
                load_vf(stg, sp, s, v, f);
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

// FUSED backward: gfr / gfi are lambda [C], ggam gamma [C], gstart [B, L]; outputs du = gamma * G [B, L, C] and the
// per-row partial sums dlam (complex) = sum_t G_t conj(h_{t-1}) (1 - start_t), dgam = sum_t Re(G_t conj(v_t)) / gamma
// with v_t = h_t - f_t h_{t-1} (so u is not re-read), written as [B, C] and summed over B by the caller.
template <int S, int NST, bool FUSED>
__global__ void __launch_bounds__(kCThreads) lru_bwd_kernel(
    const float* __restrict__ ggr, const float* __restrict__ ggi, const float* __restrict__ gfr,
    const float* __restrict__ gfi, const float* __restrict__ ghr, const float* __restrict__ ghi,
    const float* __restrict__ gh0r, const float* __restrict__ gh0i, const float* __restrict__ ggd,
    float* __restrict__ gdvr, float* __restrict__ gdvi, float* __restrict__ gdfr, float* __restrict__ gdfi,
    int L, int C, const float* __restrict__ ggam, float* __restrict__ gdgam) {
    // unfused: g, f, hprev (re/im), gd;  fused: g, h, hprev (re/im), start
    constexpr int TL = 32 * S, ARR = TL * 32, STAGE = 6 * ARR + TL;
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
                if (FUSED) {
                    cp_async16(d + 2 * ARR, ghr + g, ok);
                    cp_async16(d + 3 * ARR, ghi + g, ok);
                } else {
                    cp_async16(d + 2 * ARR, gfr + g, ok);
                    cp_async16(d + 3 * ARR, gfi + g, ok);
                }
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
                bool ok = (ggd != nullptr) && r < L;           // fused: ggd carries the reset flags `start`
                cp_async4(st + 6 * ARR + tid, ggd + (ok ? rowbase + (L - 1 - r) : 0), ok);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int i = 0; i < NST - 1; ++i) issue(i);

    C4 lam, dlam;
    F4c gam, dgam;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        lam.re.v[i] = 0.f; lam.im.v[i] = 0.f; gam.v[i] = 1.f;
        dlam.re.v[i] = 0.f; dlam.im.v[i] = 0.f; dgam.v[i] = 0.f;
    }
    if (FUSED && cvalid) {
        lam.re = cld4(gfr + c0 + cq * 4); lam.im = cld4(gfi + c0 + cq * 4); gam = cld4(ggam + c0 + cq * 4);
    }
    // f_t and the factor that gates the propagation of E (unfused: 1 - grad_detach; fused: resets live inside f)
    auto load_f = [&](const float* st, const float* sp, int s, C4& f, float& gate, float& keep) {
        const float flag = st[6 * ARR + ck * S + s];
        if (FUSED) {
            keep = 1.0f - flag; gate = 1.0f;
#pragma unroll
            for (int i = 0; i < 4; ++i) { f.re.v[i] = lam.re.v[i] * keep; f.im.v[i] = lam.im.v[i] * keep; }
        } else {
            keep = 1.0f; gate = 1.0f - flag;
            f = C4{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
        }
    };

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
                C4 f;
                float gate, keep;
                load_f(st, sp, s, f, gate, keep);
                // E_r = conj(f) * (g + gate * E_{r-1})  ->  A' = conj(f)*gate, H' = conj(f)*g
                C4 a, h;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a.re.v[i] = f.re.v[i] * gate;
                    a.im.v[i] = -f.im.v[i] * gate;
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
                C4 hp{cld4(sp + 4 * ARR + s * 32), cld4(sp + 5 * ARR + s * 32)};
                C4 f;
                float gate, keep;
                load_f(st, sp, s, f, gate, keep);
                C4 G, df;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    G.re.v[i] = g.re.v[i] + gate * E.re.v[i];
                    G.im.v[i] = g.im.v[i] + gate * E.im.v[i];
                    df.re.v[i] = G.re.v[i] * hp.re.v[i] + G.im.v[i] * hp.im.v[i];
                    df.im.v[i] = G.im.v[i] * hp.re.v[i] - G.re.v[i] * hp.im.v[i];
                    E.re.v[i] = G.re.v[i] * f.re.v[i] + G.im.v[i] * f.im.v[i];
                    E.im.v[i] = G.im.v[i] * f.re.v[i] - G.re.v[i] * f.im.v[i];
                }
                if (FUSED) {
                    C4 ht{cld4(sp + 2 * ARR + s * 32), cld4(sp + 3 * ARR + s * 32)};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        dlam.re.v[i] = fmaf(keep, df.re.v[i], dlam.re.v[i]);
                        dlam.im.v[i] = fmaf(keep, df.im.v[i], dlam.im.v[i]);
                        const float vr = ht.re.v[i] - (f.re.v[i] * hp.re.v[i] - f.im.v[i] * hp.im.v[i]);
                        const float vi = ht.im.v[i] - (f.re.v[i] * hp.im.v[i] + f.im.v[i] * hp.re.v[i]);
                        dgam.v[i] += G.re.v[i] * vr + G.im.v[i] * vi;
                        G.re.v[i] *= gam.v[i];                      // du = gamma * G
                        G.im.v[i] *= gam.v[i];
                    }
                }
                if (cvalid) {
                    size_t o = (rowbase + (L - 1 - r)) * C + c0 + cq * 4;
                    cst4(gdvr + o, G.re); cst4(gdvi + o, G.im);
                    if (!FUSED) { cst4(gdfr + o, df.re); cst4(gdfi + o, df.im); }
                }
            }
        }
    }
    cp_async_wait<0>();
    if (FUSED) {
        // sum the 32 time-lanes (ck) of each channel quad: [32][8][12] floats in the (now idle) staging memory
        __syncthreads();
        float* red = smem;
        float* mine = red + (ck * 8 + cq) * 12;
        cst4(mine, dlam.re); cst4(mine + 4, dlam.im); cst4(mine + 8, dgam);
        __syncthreads();
        if (tid < 96) {
            const int q = tid / 12, j = tid % 12;              // quad, value slot
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) acc += red[(k * 8 + q) * 12 + j];
            const int c = c0 + q * 4 + (j & 3);
            if (c < C) {
                const size_t o = (size_t)b * C + c;
                if (j < 4) gdfr[o] = acc;
                else if (j < 8) gdfi[o] = acc;
                else gdgam[o] = acc / __ldg(ggam + c);
            }
        }
    }
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
    auto kern = lru_fwd_kernel<S, NST, false>;
    constexpr size_t smem = sizeof(float) * (NST * (4 * 32 * S * 32) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(v_re, v_im, f_re, f_im, h0_re, h0_im, h_re, h_im, (int)L, (int)C, nullptr, nullptr);
    RORL_RETURN_LAUNCH();
}

int rorl_lru_fused_fwd(const float* u_re, const float* u_im, const float* lam_re, const float* lam_im, const float* gamma,
                       const float* start, const float* h0_re, const float* h0_im, float* h_re, float* h_im, int64_t B,
                       int64_t L, int64_t C, cudaStream_t stream) {
    if (!u_re || !u_im || !lam_re || !lam_im || !gamma || !h_re || !h_im) return RORL_ERR_ARG;
    if ((h0_re == nullptr) != (h0_im == nullptr)) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || C <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (C % 4) return RORL_ERR_ALIGN;
    constexpr int S = RORL_LRU_FWD_S, NST = 3;
    auto kern = lru_fwd_kernel<S, NST, true>;
    constexpr size_t smem = sizeof(float) * (NST * (2 * 32 * S * 32 + 32 * S) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(u_re, u_im, lam_re, lam_im, h0_re, h0_im, h_re, h_im, (int)L, (int)C, gamma, start);
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
    auto kern = lru_bwd_kernel<S, NST, false>;
    constexpr size_t smem = sizeof(float) * (NST * (6 * 32 * S * 32 + 32 * S) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(g_re, g_im, f_re, f_im, h_re, h_im, h0_re, h0_im, grad_detach, dv_re,
                                            dv_im, df_re, df_im, (int)L, (int)C, nullptr, nullptr);
    RORL_RETURN_LAUNCH();
}

int rorl_lru_fused_bwd(const float* g_re, const float* g_im, const float* lam_re, const float* lam_im, const float* gamma,
                       const float* start, const float* h_re, const float* h_im, const float* h0_re, const float* h0_im,
                       float* du_re, float* du_im, float* dlam_re_part, float* dlam_im_part, float* dgamma_part,
                       int64_t B, int64_t L, int64_t C, cudaStream_t stream) {
    if (!g_re || !g_im || !lam_re || !lam_im || !gamma || !h_re || !h_im || !du_re || !du_im || !dlam_re_part ||
        !dlam_im_part || !dgamma_part)
        return RORL_ERR_ARG;
    if ((h0_re == nullptr) != (h0_im == nullptr)) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || C <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (C % 4) return RORL_ERR_ALIGN;
    constexpr int S = 2, NST = 2;
    auto kern = lru_bwd_kernel<S, NST, true>;
    constexpr size_t smem = sizeof(float) * (NST * (6 * 32 * S * 32 + 32 * S) + 2 * 8 * 8 * 16);
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kCThreads, smem, stream>>>(g_re, g_im, lam_re, lam_im, h_re, h_im, h0_re, h0_im, start, du_re, du_im,
                                            dlam_re_part, dlam_im_part, (int)L, (int)C, gamma, dgamma_part);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
