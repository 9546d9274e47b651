// Depthwise causal conv1d over time + SiLU on token-major [B, L, D], valid-step mask applied to the
// conv input.  Replaces `x = mask * x; x = act(conv1d(x)[..., :L])` of the d_conv > 4 path
// (ref: offpolicy_rnn/models/smamba/mamba.py:207-212; the nn.Conv1d is built at :75-83 with
// groups = d_inner, padding = d_conv - 1, so out[t] = bias + sum_k w[k] * xm[t - (K-1) + k]).
//
// One thread per channel (lanes along D: every row access is a coalesced 128-B-per-warp line),
// walking a segment of the sequence with the K-tap window held in registers (static rotation by
// unrolling K steps).  HBM-bound: forward 8 B/element, backward 12 B/element (+ K-1 halo rows).
#include "common.cuh"
#include <cstdlib>

namespace rorl {

constexpr int kConvThreads = 128;
constexpr int kConvSeg = 128;   // steps per CTA segment (multiple of every supported K)

__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_fast(x); }

// All loads of a K-step batch are issued branch-free (time index clamped to the segment, the valid-step mask
// either compiled in or out) one batch ahead of their use, so each thread keeps 2K independent loads in flight;
// a branch per element (bounds / mask tests) had left the kernel latency-bound at ~0.9 TB/s.
template <int K, bool MASK>
__global__ void __launch_bounds__(kConvThreads) conv1d_silu_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ mask, float* __restrict__ y, int L, int D, int ld_x, int ld_y, int act) {
    const int d = blockIdx.x * kConvThreads + threadIdx.x;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kConvSeg, t1 = min(L, t0 + kConvSeg);
    if (d >= D) return;
    float wk[K], win[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = __ldg(w + (size_t)d * K + k);
    const float bs = bias ? __ldg(bias + d) : 0.f;
    const size_t row0 = (size_t)b * L;
    const float* xp = x + row0 * ld_x + d;
    const float* mp = mask + row0;
    // halo: win slot k holds xm[t0 - K + k] for k = 1..K-1 (slot s % K with local time s = t - t0 + K)
    win[0] = 0.f;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const int t = t0 - K + k, tc = max(t, 0);
        float v = __ldg(xp + (size_t)tc * ld_x);
        if (MASK) v *= __ldg(mp + tc);
        win[k] = t >= 0 ? v : 0.f;
    }
    // the mask is only LOADED here and applied when the batch is consumed: multiplying inside fetch() made every
    // batch wait for its own loads right after issuing them (ncu: 56 % of the stalls were long-scoreboard on those
    // multiplies), i.e. no load ever overlapped the arithmetic of the previous batch
    float nx[K], nm[K];
    auto fetch = [&](int tb) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int tc = min(tb + j, t1 - 1);
            nx[j] = __ldg(xp + (size_t)tc * ld_x);
            if (MASK) nm[j] = __ldg(mp + tc);
        }
    };
    fetch(t0);
    for (int tb = t0; tb < t1; tb += K) {
        float cx[K];
#pragma unroll
        for (int j = 0; j < K; ++j) cx[j] = MASK ? nx[j] * nm[j] : nx[j];
        fetch(min(tb + K, t1 - 1));
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int t = tb + j;
            win[j] = cx[j];
            // taps: xm[t - (K-1) + k] lives in slot (j + 1 + k) % K
            float acc = bs;
#pragma unroll
            for (int k = 0; k < K; ++k) acc = fmaf(wk[k], win[(j + 1 + k) % K], acc);
            if (t < t1) y[(row0 + t) * ld_y + d] = act ? siluf_(acc) : acc;      // act: warp-uniform
        }
    }
}

// Backward.  pre_t = bias + conv(xm)_t; dpre_t = dy_t * silu'(pre_t);
//   dxm_s = sum_k w[k] * dpre[s + (K-1) - k];  dw[k] = sum_t dpre_t * xm[t - (K-1) + k];  db = sum_t dpre_t.
// A segment owns dx for s in [t0, t1): it evaluates dpre on [t0, t1 + K - 1) but only counts
// t in [t0, t1) towards dw / dbias.
template <int K, bool MASK>
__global__ void __launch_bounds__(kConvThreads) conv1d_silu_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ mask, const float* __restrict__ dy, float* __restrict__ dx,
    float* __restrict__ dw_part, float* __restrict__ db_part, int L, int D, int ld_x, int ld_dy, int ld_dx, int act) {
    const int d = blockIdx.x * kConvThreads + threadIdx.x;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kConvSeg, t1 = min(L, t0 + kConvSeg);
    if (d >= D) return;
    float wk[K], win[K], acc_dx[K], dwk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        wk[k] = __ldg(w + (size_t)d * K + k);
        acc_dx[k] = 0.f; dwk[k] = 0.f;
    }
    const float bs = bias ? __ldg(bias + d) : 0.f;
    float db = 0.f;
    const size_t row0 = (size_t)b * L;
    const float* xp = x + row0 * ld_x + d;
    const float* gp = dy + row0 * ld_dy + d;
    const float* mp = mask + row0;
    win[0] = 0.f;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const int t = t0 - K + k, tc = max(t, 0);
        float v = __ldg(xp + (size_t)tc * ld_x);
        if (MASK) v *= __ldg(mp + tc);
        win[k] = t >= 0 ? v : 0.f;
    }
    const int tend = min(L, t1 + K - 1);       // dpre is needed on [t0, tend)
    const int tstop = tend + K - 1;            // positions up to t1 - 1 complete by step tend + K - 2
    // acc_dx slot (s % K) accumulates dxm for position t = s - K + t0; position t - (K-1) completes at step t.
    float nx[K], ng[K], nm[K];
    auto fetch = [&](int tb) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int tc = min(tb + j, tend - 1);
            nx[j] = __ldg(xp + (size_t)tc * ld_x);
            ng[j] = __ldg(gp + (size_t)tc * ld_dy);
            if (MASK) nm[j] = __ldg(mp + tc);            // applied at consumption (see the forward kernel)
        }
    };
    fetch(t0);
    for (int tb = t0; tb < tstop; tb += K) {
        float cx[K], cg[K];
#pragma unroll
        for (int j = 0; j < K; ++j) { cx[j] = MASK ? nx[j] * nm[j] : nx[j]; cg[j] = ng[j]; }
        fetch(min(tb + K, tend - 1));
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int t = tb + j;
            const int sp = t - (K - 1);                  // the position this step completes
            // its mask is fetched here, unconditionally (clamped), not inside the guarded store below: a load behind the
            // branch put one exposed L1 round trip at the end of every step
            float msp = 1.f;
            if (MASK) msp = __ldg(mp + min(max(sp, t0), t1 - 1));
            win[j] = cx[j];
            float pre = bs, pre1 = 0.f;
#pragma unroll
            for (int k = 0; k < K; k += 2) {
                pre = fmaf(wk[k], win[(j + 1 + k) % K], pre);
                if (k + 1 < K) pre1 = fmaf(wk[k + 1], win[(j + 2 + k) % K], pre1);
            }
            pre += pre1;
            const float sg = sigmoidf_fast(pre);
            const float dfull = act ? cg[j] * sg * (1.0f + pre * (1.0f - sg)) : cg[j];
            const float dpre = t < tend ? dfull : 0.f;          // steps past the sequence / halo end contribute nothing
            const float down = t < t1 ? dpre : 0.f;             // dw / dbias count this segment's own steps only
            db += down;
#pragma unroll
            for (int k = 0; k < K; ++k) dwk[k] = fmaf(down, win[(j + 1 + k) % K], dwk[k]);
            // scatter dpre_t into positions t-(K-1)+k (slot (j+1+k)%K); slot j is the newest (k = K-1)
            acc_dx[j] = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) acc_dx[(j + 1 + k) % K] = fmaf(wk[k], dpre, acc_dx[(j + 1 + k) % K]);
            // position t-(K-1) (slot (j+1)%K) has now received all its contributions
            if (sp >= t0 && sp < t1) dx[(row0 + sp) * ld_dx + d] = acc_dx[(j + 1) % K] * msp;
        }
    }
    const size_t part = (size_t)b * gridDim.y + blockIdx.y;
#pragma unroll
    for (int k = 0; k < K; ++k) dw_part[(part * D + d) * K + k] = dwk[k];
    db_part[part * D + d] = db;
}


// ---------------------------------------------------------------------------------------------------------------
// Channel-pair forward (D, row strides even, 8-byte aligned rows): thread = two adjacent channels, every tap is one
// packed fma.rn.f32x2, loads / stores are 8 bytes per thread.  ncu on the one-channel kernel above at the update's
// shape ([32, 1019, 512], K = 16): 52 warp instructions per element, issue slots 61 % busy, 1.73 waves of 118-register
// CTAs -- instruction-bound at half the HBM rate.  Here: 28 per element; 64-thread CTAs so that the whole grid (1024
// CTAs at that shape) is resident at once; a rolling register prefetch `PD` steps ahead instead of a K-step batch.
// 41.5 -> 37 us at that shape.  (The same treatment of the backward kernel -- 77 instead of 131 instructions per
// element -- needs 250 registers for its three K-wide pair accumulators, or 168 with a little spilling: 8 to 12 warps
// per SM, latency-bound at 129 us against 124 us for the one-channel kernel, so the backward stays one-channel.)
// ---------------------------------------------------------------------------------------------------------------
constexpr int kConv2Threads = 64;

__device__ __forceinline__ float2 ld2(const float* p) { return __ldg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ float2 silu2(float2 a) { return make_float2(siluf_(a.x), siluf_(a.y)); }

template <int K, bool MASK, bool ACT>
__global__ void __launch_bounds__(kConv2Threads, 8) conv1d_silu_fwd2_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ mask, float* __restrict__ y, int L, int D, int ld_x, int ld_y) {
    constexpr int PD = K < 8 ? K : 8;
    const int d = (blockIdx.x * kConv2Threads + threadIdx.x) * 2;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kConvSeg, t1 = min(L, t0 + kConvSeg);
    if (d >= D) return;
    float2 wk[K], win[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = make_float2(__ldg(w + (size_t)d * K + k), __ldg(w + (size_t)(d + 1) * K + k));
    const float2 bs = bias ? make_float2(__ldg(bias + d), __ldg(bias + d + 1)) : make_float2(0.f, 0.f);
    const size_t row0 = (size_t)b * L;
    const float* xp = x + row0 * ld_x + d;
    const float* mp = mask + row0;
    float* yp = y + row0 * ld_y + d;
    win[0] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 1; k < K; ++k) {
        const int t = t0 - K + k, tc = max(t, 0);
        float2 v = ld2(xp + (size_t)tc * ld_x);
        if (MASK) { const float m = __ldg(mp + tc); v.x *= m; v.y *= m; }
        win[k] = t >= 0 ? v : make_float2(0.f, 0.f);
    }
    float2 px[PD];
    float pm[PD];
#pragma unroll
    for (int j = 0; j < PD; ++j) {
        const int tc = min(t0 + j, t1 - 1);
        px[j] = ld2(xp + (size_t)tc * ld_x);
        if (MASK) pm[j] = __ldg(mp + tc);
    }
    // running 32-bit element offsets (the segment's rows are within 2^31 elements of its first row): the prefetch row
    // min(t + PD, t1 - 1) advances by one row per step until it reaches the last row; recomputing clamped 64-bit
    // addresses per access had cost twice as many integer instructions as there are FMAs
    xp += (size_t)t0 * ld_x; mp += t0; yp += (size_t)t0 * ld_y;
    const int last = t1 - 1 - t0;
    int rl = min(PD, last);                       // prefetch row relative to t0
    int ox = rl * ld_x, oy = 0;
    for (int tb = t0; tb < t1; tb += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int t = tb + j;
            float2 cx = px[j % PD];
            if (MASK) { cx.x *= pm[j % PD]; cx.y *= pm[j % PD]; }
            px[j % PD] = ld2(xp + ox);
            if (MASK) pm[j % PD] = __ldg(mp + rl);
            if (rl < last) { ++rl; ox += ld_x; }
            win[j] = cx;
            // taps: xm[t - (K-1) + k] lives in slot (j + 1 + k) % K; two interleaved chains halve the dependent depth
            float2 a0 = bs, a1 = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < K; k += 2) {
                a0 = __ffma2_rn(wk[k], win[(j + 1 + k) % K], a0);
                if (k + 1 < K) a1 = __ffma2_rn(wk[k + 1], win[(j + 2 + k) % K], a1);
            }
            // the activation is computed unconditionally and only the store is guarded: a guarded block holding the
            // MUFU chain became a branch per step, which kept the scheduler from overlapping one step's tail with the
            // next step's taps (ncu: 34 % issue-slot use at 2.4 warps per scheduler)
            const float2 out = ACT ? silu2(__fadd2_rn(a0, a1)) : __fadd2_rn(a0, a1);
            if (t < t1) *reinterpret_cast<float2*>(yp + oy) = out;
            oy += ld_y;
        }
    }
}

}  // namespace rorl

using namespace rorl;

static bool a8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; }
static const bool g_conv_force_single = getenv("RORL_CONV_SINGLE") != nullptr;     // diagnostic: one-channel kernels

extern "C" {

int rorl_conv1d_nseg(int64_t L) { return (int)((L + kConvSeg - 1) / kConvSeg); }

#define CONV_CASE1(KERN, KK, ...)                                                                                    \
    case KK:                                                                                                          \
        if (mask) KERN##_kernel<KK, true><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__, act != 0);                  \
        else KERN##_kernel<KK, false><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__, act != 0);                      \
        break;
#define CONV_DISPATCH1(KERN, ...)                   \
    switch (K) {                                    \
        CONV_CASE1(KERN, 2, __VA_ARGS__)            \
        CONV_CASE1(KERN, 3, __VA_ARGS__)            \
        CONV_CASE1(KERN, 4, __VA_ARGS__)            \
        CONV_CASE1(KERN, 8, __VA_ARGS__)            \
        CONV_CASE1(KERN, 16, __VA_ARGS__)           \
        default: return RORL_ERR_SHAPE;             \
    }
#define CONV_CASE(KERN, KK, ...)                                                        \
    case KK:                                                                            \
        if (pair && mask && act) KERN##2_kernel<KK, true, true><<<grid2, kConv2Threads, 0, stream>>>(__VA_ARGS__);     \
        else if (pair && mask) KERN##2_kernel<KK, true, false><<<grid2, kConv2Threads, 0, stream>>>(__VA_ARGS__);     \
        else if (pair && act) KERN##2_kernel<KK, false, true><<<grid2, kConv2Threads, 0, stream>>>(__VA_ARGS__);      \
        else if (pair) KERN##2_kernel<KK, false, false><<<grid2, kConv2Threads, 0, stream>>>(__VA_ARGS__);            \
        else if (mask) KERN##_kernel<KK, true><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__, act != 0);             \
        else KERN##_kernel<KK, false><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__, act != 0);                      \
        break;
#define CONV_DISPATCH(KERN, ...)                   \
    switch (K) {                                   \
        CONV_CASE(KERN, 2, __VA_ARGS__)            \
        CONV_CASE(KERN, 3, __VA_ARGS__)            \
        CONV_CASE(KERN, 4, __VA_ARGS__)            \
        CONV_CASE(KERN, 8, __VA_ARGS__)            \
        CONV_CASE(KERN, 16, __VA_ARGS__)           \
        default: return RORL_ERR_SHAPE;            \
    }

int rorl_conv1d_fwd(const float* x, const float* w, const float* bias, const float* mask, float* y, int64_t B, int64_t L,
                    int64_t D, int64_t K, int64_t ld_x, int64_t ld_y, int act, cudaStream_t stream) {
    if (!x || !w || !y) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)((D + kConvThreads - 1) / kConvThreads), (unsigned)rorl_conv1d_nseg(L), (unsigned)B);
    dim3 grid2((unsigned)((D / 2 + kConv2Threads - 1) / kConv2Threads), (unsigned)rorl_conv1d_nseg(L), (unsigned)B);
    const bool pair = !g_conv_force_single && D % 2 == 0 && ld_x % 2 == 0 && ld_y % 2 == 0 && a8(x) && a8(y);
    CONV_DISPATCH(conv1d_silu_fwd, x, w, bias, mask, y, (int)L, (int)D, (int)ld_x, (int)ld_y);
    RORL_RETURN_LAUNCH();
}

int rorl_conv1d_silu_fwd(const float* x, const float* w, const float* bias, const float* mask, float* y,
                         int64_t B, int64_t L, int64_t D, int64_t K, int64_t ld_x, int64_t ld_y,
                         cudaStream_t stream) {
    return rorl_conv1d_fwd(x, w, bias, mask, y, B, L, D, K, ld_x, ld_y, 1, stream);
}

int rorl_conv1d_bwd(const float* x, const float* w, const float* bias, const float* mask, const float* dy, float* dx,
                    float* dw_part, float* dbias_part, int64_t B, int64_t L, int64_t D, int64_t K, int64_t ld_x,
                    int64_t ld_dy, int64_t ld_dx, int act, cudaStream_t stream) {
    if (!x || !w || !dy || !dx || !dw_part || !dbias_part) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)((D + kConvThreads - 1) / kConvThreads), (unsigned)rorl_conv1d_nseg(L), (unsigned)B);
    CONV_DISPATCH1(conv1d_silu_bwd, x, w, bias, mask, dy, dx, dw_part, dbias_part, (int)L, (int)D,
                  (int)ld_x, (int)ld_dy, (int)ld_dx);
    RORL_RETURN_LAUNCH();
}

int rorl_conv1d_silu_bwd(const float* x, const float* w, const float* bias, const float* mask, const float* dy,
                         float* dx, float* dw_part, float* dbias_part, int64_t B, int64_t L, int64_t D, int64_t K,
                         int64_t ld_x, int64_t ld_dy, int64_t ld_dx, cudaStream_t stream) {
    return rorl_conv1d_bwd(x, w, bias, mask, dy, dx, dw_part, dbias_part, B, L, D, K, ld_x, ld_dy, ld_dx, 1, stream);
}

}  // extern "C"
