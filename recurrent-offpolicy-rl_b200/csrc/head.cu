// Tanh-Gaussian policy head, forward and backward, one kernel each (sm_100a).
//
// Replaces the ~30 elementwise ATen launches of `process_model_out` and its autograd graph
// (ref: offpolicy_rnn/policy_value_models/contextual_sac_policy_single_head.py:105-123):
//   logstd = clamp(out[:, :A], lo, hi);  mean = out[:, A:];  sample = mean + noise * exp(logstd)
//   log_prob = sum_a [ -noise^2 / 2 - logstd - log(2 pi) / 2 - 2 (log 2 - sample - softplus(-2 sample)) ]
//   returns tanh(mean), tanh(sample), log_prob
// Backward (d log_prob / d sample = 2 tanh(sample), since d/ds [log 2 - s - softplus(-2 s)] = -tanh(s)):
//   d sample = d_as (1 - as^2) + 2 as d_lp;  d mean = d_am (1 - am^2) + d sample;
//   d logstd = (d sample * noise * std - d_lp) * [lo <= raw logstd <= hi]      (torch.clamp's gradient mask)
// One thread per row (A is the action dimension, a handful of columns); precise expf / log1pf / tanhf -- the arithmetic is
// negligible (2 A transcendental calls per row) and the log-density enters the actor and alpha losses directly.
#include "common.cuh"

namespace rorl {

constexpr float kHalfLog2Pi = 0.91893853320467274178f;

__device__ __forceinline__ float softplus_precise(float x) { return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(256) tanh_gaussian_fwd_kernel(const float* __restrict__ out, const float* __restrict__ noise,
                                                                float* __restrict__ am, float* __restrict__ as,
                                                                float* __restrict__ logp, int64_t M, int A, int64_t ld_out,
                                                                float lo, float hi) {
    const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (m >= M) return;
    const float* o = out + m * ld_out;
    float lp = 0.f;
    for (int a = 0; a < A; ++a) {
        const float ls = fminf(fmaxf(o[a], lo), hi), mu = o[A + a], n = noise[m * A + a];
        const float s = fmaf(n, expf(ls), mu);
        lp += -0.5f * n * n - (ls + kHalfLog2Pi);
        lp -= 2.f * (-s - softplus_precise(-2.f * s) + kLn2);
        am[m * A + a] = tanhf(mu);
        as[m * A + a] = tanhf(s);
    }
    logp[m] = lp;
}

__global__ void __launch_bounds__(256) tanh_gaussian_bwd_kernel(const float* __restrict__ out, const float* __restrict__ noise,
                                                                const float* __restrict__ d_am, const float* __restrict__ d_as,
                                                                const float* __restrict__ d_lp, float* __restrict__ d_out,
                                                                int64_t M, int A, int64_t ld_out, float lo, float hi) {
    const int64_t m = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (m >= M) return;
    const float* o = out + m * ld_out;
    const float glp = d_lp ? d_lp[m] : 0.f;
    for (int a = 0; a < A; ++a) {
        const float raw = o[a], mu = o[A + a], n = noise[m * A + a];
        const float ls = fminf(fmaxf(raw, lo), hi), sd = expf(ls);
        const float s = fmaf(n, sd, mu);
        const float ts = tanhf(s), tm = tanhf(mu);
        float ds = 2.f * ts * glp;
        if (d_as) ds = fmaf(d_as[m * A + a], 1.f - ts * ts, ds);
        float dmu = ds;
        if (d_am) dmu = fmaf(d_am[m * A + a], 1.f - tm * tm, dmu);
        const float dls = (raw >= lo && raw <= hi) ? fmaf(ds * n, sd, -glp) : 0.f;
        d_out[m * 2 * A + a] = dls;
        d_out[m * 2 * A + A + a] = dmu;
    }
}

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_tanh_gaussian_fwd(const float* out, const float* noise, float* action_mean, float* action_sample, float* log_prob,
                           int64_t M, int64_t A, int64_t ld_out, float min_logstd, float max_logstd, cudaStream_t stream) {
    if (!out || !noise || !action_mean || !action_sample || !log_prob) return RORL_ERR_ARG;
    if (M <= 0 || A <= 0 || A > 1024 || ld_out < 2 * A) return RORL_ERR_SHAPE;
    tanh_gaussian_fwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(out, noise, action_mean, action_sample, log_prob, M,
                                                                              (int)A, ld_out, min_logstd, max_logstd);
    RORL_RETURN_LAUNCH();
}

int rorl_tanh_gaussian_bwd(const float* out, const float* noise, const float* d_mean, const float* d_sample, const float* d_log_prob,
                           float* d_out, int64_t M, int64_t A, int64_t ld_out, float min_logstd, float max_logstd,
                           cudaStream_t stream) {
    if (!out || !noise || !d_out) return RORL_ERR_ARG;
    if (M <= 0 || A <= 0 || A > 1024 || ld_out < 2 * A) return RORL_ERR_SHAPE;
    tanh_gaussian_bwd_kernel<<<(unsigned)((M + 255) / 256), 256, 0, stream>>>(out, noise, d_mean, d_sample, d_log_prob, d_out, M,
                                                                              (int)A, ld_out, min_logstd, max_logstd);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
