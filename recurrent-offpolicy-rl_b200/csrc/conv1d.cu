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
constexpr int kConvSeg = 128;   // steps per CTA segment

__device__ __forceinline__ float siluf_(float x) { return x * sigmoidf_fast(x); }

template <int K>
__global__ void __launch_bounds__(kConvThreads) conv1d_silu_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ mask, float* __restrict__ y, int L, int D, int ld_x, int ld_y) {
    const int d = blockIdx.x * kConvThreads + threadIdx.x;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kConvSeg, t1 = min(L, t0 + kConvSeg);
    if (d >= D) return;
    float wk[K], win[K];
#pragma unroll
    for (int k = 0; k < K; ++k) wk[k] = w[(size_t)d * K + k];
    const float bs = bias ? bias[d] : 0.f;
    const size_t row0 = (size_t)b * L;
    // preload the K-1 halo steps: win slot (t mod K) holds xm[t]
#pragma unroll
    for (int k = 0; k < K; ++k) win[k] = 0.f;
    // Align the main loop so that (t - t0) % K is static: halo fills slots for t0-(K-1) .. t0-1.
    // Use local time s = t - t0 + K (so halo has s in [1, K-1], main loop starts at s = K).
#pragma unroll
    for (int k = 1; k < K; ++k) {
        int t = t0 - K + k;
        float v = 0.f;
        if (t >= 0) {
            v = x[(row0 + t) * ld_x + d];
            if (mask) v *= mask[row0 + t];
        }
        win[k] = v;  // slot s % K with s = k
    }
    for (int tb = t0; tb < t1; tb += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int t = tb + j;
            if (t < t1) {
                float v = x[(row0 + t) * ld_x + d];
                if (mask) v *= mask[row0 + t];
                win[j] = v;  // s = t - t0 + K, s % K == j
                // taps: xm[t - (K-1) + k] lives in slot (j + 1 + k) % K
                float acc = bs;
#pragma unroll
                for (int k = 0; k < K; ++k) acc = fmaf(wk[k], win[(j + 1 + k) % K], acc);
                y[(row0 + t) * ld_y + d] = siluf_(acc);
            }
        }
    }
}

// Backward.  pre_t = bias + conv(xm)_t; dpre_t = dy_t * silu'(pre_t);
//   dxm_s = sum_k w[k] * dpre[s + (K-1) - k];  dw[k] = sum_t dpre_t * xm[t - (K-1) + k];  db = sum_t dpre_t.
// A segment owns dx for s in [t0, t1): it evaluates dpre on [t0, t1 + K - 1) but only counts
// t in [t0, t1) towards dw / dbias.
template <int K>
__global__ void __launch_bounds__(kConvThreads) conv1d_silu_bwd_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ mask, const float* __restrict__ dy, float* __restrict__ dx,
    float* __restrict__ dw_part, float* __restrict__ db_part, int L, int D, int ld_x, int ld_dy, int ld_dx) {
    const int d = blockIdx.x * kConvThreads + threadIdx.x;
    const int b = blockIdx.z;
    const int t0 = blockIdx.y * kConvSeg, t1 = min(L, t0 + kConvSeg);
    if (d >= D) return;
    float wk[K], win[K], acc_dx[K], dwk[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        wk[k] = w[(size_t)d * K + k];
        win[k] = 0.f; acc_dx[k] = 0.f; dwk[k] = 0.f;
    }
    const float bs = bias ? bias[d] : 0.f;
    float db = 0.f;
    const size_t row0 = (size_t)b * L;
#pragma unroll
    for (int k = 1; k < K; ++k) {
        int t = t0 - K + k;
        float v = 0.f;
        if (t >= 0) {
            v = x[(row0 + t) * ld_x + d];
            if (mask) v *= mask[row0 + t];
        }
        win[k] = v;
    }
    const int tend = min(L, t1 + K - 1);
    // acc_dx slot (s % K) accumulates dxm for position t = s - K + t0; position t - (K-1) completes at step t.
    for (int tb = t0; tb < tend + K - 1; tb += K) {
#pragma unroll
        for (int j = 0; j < K; ++j) {
            int t = tb + j;
            if (t < tend + K - 1) {
                float dpre = 0.f;
                if (t < tend) {
                    float v = x[(row0 + t) * ld_x + d];
                    if (mask) v *= mask[row0 + t];
                    win[j] = v;
                    float pre = bs;
#pragma unroll
                    for (int k = 0; k < K; ++k) pre = fmaf(wk[k], win[(j + 1 + k) % K], pre);
                    float sg = sigmoidf_fast(pre);
                    dpre = dy[(row0 + t) * ld_dy + d] * sg * (1.0f + pre * (1.0f - sg));
                    if (t < t1) {
                        db += dpre;
#pragma unroll
                        for (int k = 0; k < K; ++k) dwk[k] = fmaf(dpre, win[(j + 1 + k) % K], dwk[k]);
                    }
                }
                // scatter dpre_t into positions t-(K-1)+k (slot (j+1+k)%K); slot j is the newest (k = K-1)
                acc_dx[j] = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) acc_dx[(j + 1 + k) % K] = fmaf(wk[k], dpre, acc_dx[(j + 1 + k) % K]);
                // position t-(K-1) (slot (j+1)%K) has now received all its contributions
                int s = t - (K - 1);
                if (s >= t0 && s < t1) {
                    float g = acc_dx[(j + 1) % K];
                    if (mask) g *= mask[row0 + s];
                    dx[(row0 + s) * ld_dx + d] = g;
                }
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

#define CONV_DISPATCH(KERN, ...)                                                   \
    switch (K) {                                                                   \
        case 2: KERN<2><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__); break;    \
        case 3: KERN<3><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__); break;    \
        case 4: KERN<4><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__); break;    \
        case 8: KERN<8><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__); break;    \
        case 16: KERN<16><<<grid, kConvThreads, 0, stream>>>(__VA_ARGS__); break;  \
        default: return RORL_ERR_SHAPE;                                            \
    }

int rorl_conv1d_silu_fwd(const float* x, const float* w, const float* bias, const float* mask, float* y,
                         int64_t B, int64_t L, int64_t D, int64_t K, int64_t ld_x, int64_t ld_y,
                         cudaStream_t stream) {
    if (!x || !w || !y) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)((D + kConvThreads - 1) / kConvThreads), (unsigned)rorl_conv1d_nseg(L), (unsigned)B);
    CONV_DISPATCH(conv1d_silu_fwd_kernel, x, w, bias, mask, y, (int)L, (int)D, (int)ld_x, (int)ld_y);
    RORL_RETURN_LAUNCH();
}

int rorl_conv1d_silu_bwd(const float* x, const float* w, const float* bias, const float* mask, const float* dy,
                         float* dx, float* dw_part, float* dbias_part, int64_t B, int64_t L, int64_t D, int64_t K,
                         int64_t ld_x, int64_t ld_dy, int64_t ld_dx, cudaStream_t stream) {
    if (!x || !w || !dy || !dx || !dw_part || !dbias_part) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)((D + kConvThreads - 1) / kConvThreads), (unsigned)rorl_conv1d_nseg(L), (unsigned)B);
    CONV_DISPATCH(conv1d_silu_bwd_kernel, x, w, bias, mask, dy, dx, dw_part, dbias_part, (int)L, (int)D,
                  (int)ld_x, (int)ld_dy, (int)ld_dx);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
