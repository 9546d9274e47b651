// Depthwise causal conv1d over time + SiLU on token-major [B, L, D], valid-step mask applied to the
// conv input.  Replaces `x = mask * x; x = act(conv1d(x)[..., :L])` of the d_conv > 4 path
// (ref: offpolicy_rnn/models/smamba/mamba.py:207-212; the nn.Conv1d is built at :75-83 with
// groups = d_inner, padding = d_conv - 1, so out[t] = bias + sum_k w[k] * xm[t - (K-1) + k]).
//
// One thread per channel (lanes along D: every row access is a coalesced 128-B-per-warp line),
// walking a segment of the sequence with the K-tap window held in registers (static rotation by
// unrolling K steps).  HBM-bound: forward 8 B/element, backward 12 B/element (+ K-1 halo rows).
#include "common.cuh"

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
            win[j] = cx[j];
            float pre = bs;
#pragma unroll
            for (int k = 0; k < K; ++k) pre = fmaf(wk[k], win[(j + 1 + k) % K], pre);
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
            const int sp = t - (K - 1);
            if (sp >= t0 && sp < t1) {
                float g = acc_dx[(j + 1) % K];
                if (MASK) g *= __ldg(mp + sp);
                dx[(row0 + sp) * ld_dx + d] = g;
            }
        }
    }
    const size_t part = (size_t)b * gridDim.y + blockIdx.y;
#pragma unroll
    for (int k = 0; k < K; ++k) dw_part[(part * D + d) * K + k] = dwk[k];
    db_part[part * D + d] = db;
}

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_conv1d_nseg(int64_t L) { return (int)((L + kConvSeg - 1) / kConvSeg); }

#define CONV_CASE(KERN, KK, ...)                                                        \
    case KK:                                                                            \
        if (mask) KERN<KK, true><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__);       \
        else KERN<KK, false><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__);           \
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
    CONV_DISPATCH(conv1d_silu_fwd_kernel, x, w, bias, mask, y, (int)L, (int)D, (int)ld_x, (int)ld_y, act != 0);
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
    CONV_DISPATCH(conv1d_silu_bwd_kernel, x, w, bias, mask, dy, dx, dw_part, dbias_part, (int)L, (int)D,
                  (int)ld_x, (int)ld_dy, (int)ld_dx, act != 0);
    RORL_RETURN_LAUNCH();
}

int rorl_conv1d_silu_bwd(const float* x, const float* w, const float* bias, const float* mask, const float* dy,
                         float* dx, float* dw_part, float* dbias_part, int64_t B, int64_t L, int64_t D, int64_t K,
                         int64_t ld_x, int64_t ld_dy, int64_t ld_dx, cudaStream_t stream) {
    return rorl_conv1d_bwd(x, w, bias, mask, dy, dx, dw_part, dbias_part, B, L, D, K, ld_x, ld_dy, ld_dx, 1, stream);
}

}  // extern "C"
