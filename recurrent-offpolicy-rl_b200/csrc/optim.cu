// Multi-tensor AdamW (+ optional Polyak target update) over a flat parameter arena: one launch per
// model instead of ~100 foreach/elementwise launches.
// Replaces torch.optim.AdamW over the RESeL param groups (ref: offpolicy_rnn/algorithm/sac.py:61,
// 81-90; groups built by prepare_param_list, sac_full_length_rnn_redq_sep_optim.py:37-102) and the
// per-parameter soft update target = tau*target + (1-tau)*param
// (ref: offpolicy_rnn/models/rnn_base.py:475-491, called from sac.py:189-197).
// Arithmetic order follows torch's single-tensor AdamW: p *= 1 - lr*wd; m, v EMA; denom =
// sqrt(v)/sqrt(1-b2^t) + eps; p -= (lr/(1-b1^t)) * m/denom.  HBM-bound: 28 B/param (36 with target).
// Gradient clipping (ref: offpolicy_rnn/algorithm/sac_full_length_rnn_ensembleQ.py:239-250,274-287) is folded in:
// clip_grad_norm_ as a scale read from a device scalar (sum of squares from rorl_sumsq: coef = min(1, max_norm /
// (sqrt(ss) + 1e-6)), torch's formula), then clip_grad_value_ per segment (the embedding network's bound, and the
// reference's hard-coded 1e-3 on the smamba A_log tensors); the clipped gradient is written back so that the arena
// holds what torch's in-place clipping would leave in .grad.
#include "common.cuh"

namespace rorl {

constexpr int kOptThreads = 256;
constexpr int kMaxSeg = 64;

__global__ void __launch_bounds__(kOptThreads) adamw_polyak_kernel(
    float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
    float* __restrict__ target, const int64_t* __restrict__ seg_end, const double* __restrict__ seg_lr,
    const double* __restrict__ seg_wd, const double* __restrict__ seg_clip, int nseg, int64_t n, float beta1,
    float beta2, float eps, float tau, const int32_t* __restrict__ step_ptr, const float* __restrict__ gnorm_sq,
    float max_norm) {
    __shared__ int64_t s_end[kMaxSeg];
    __shared__ float s_step[kMaxSeg], s_decay[kMaxSeg], s_clip[kMaxSeg];
    __shared__ float s_bc2s;
    __shared__ double s_bc1;
    if (threadIdx.x == 0) {
        const double t = (double)(step_ptr[0] + 1);
        s_bc1 = 1.0 - pow((double)beta1, t);
        s_bc2s = (float)sqrt(1.0 - pow((double)beta2, t));
    }
    __syncthreads();
    if (threadIdx.x < nseg) {
        // scalars are formed in double and cast once, as torch does with its Python-float scalars
        s_end[threadIdx.x] = seg_end[threadIdx.x];
        s_step[threadIdx.x] = (float)(seg_lr[threadIdx.x] / s_bc1);
        s_decay[threadIdx.x] = (float)(1.0 - seg_lr[threadIdx.x] * seg_wd[threadIdx.x]);
        s_clip[threadIdx.x] = seg_clip ? (float)seg_clip[threadIdx.x] : 0.f;
    }
    __syncthreads();
    const float omt = (float)(1.0 - (double)tau);
    const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
    float coef = 1.0f;                                  // clip_grad_norm_: coef = clamp(max_norm / (norm + 1e-6), max = 1)
    if (gnorm_sq != nullptr) coef = fminf(max_norm / (sqrtf(gnorm_sq[0]) + 1e-6f), 1.0f);
    const bool any_clip = gnorm_sq != nullptr || seg_clip != nullptr;
    for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
        int s = 0;
        while (s < nseg - 1 && i >= s_end[s]) ++s;
        float gi = g[i];
        if (any_clip) {
            if (gnorm_sq != nullptr) gi *= coef;
            const float cv = s_clip[s];
            if (cv > 0.f) gi = fminf(fmaxf(gi, -cv), cv);
            g[i] = gi;
        }
        float pi = p[i];
        pi *= s_decay[s];
        float mi = m[i];
        mi = fmaf(omb1, gi - mi, mi);                   // lerp(m, g, 1 - beta1)
        float vi = v[i] * beta2 + omb2 * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / s_bc2s + eps;
        pi -= s_step[s] * (mi / denom);
        p[i] = pi;
        if (target != nullptr) target[i] = target[i] * tau + omt * pi;
    }
}

__global__ void step_inc_kernel(int32_t* step_ptr) { step_ptr[0] += 1; }

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_adamw_polyak(float* p, float* g, float* m, float* v, float* target, const int64_t* seg_end,
                      const double* seg_lr, const double* seg_wd, const double* seg_clip, int64_t nseg, int64_t n,
                      float beta1, float beta2, float eps, float tau, int32_t* step_ptr, const float* gnorm_sq,
                      float max_norm, cudaStream_t stream) {
    if (!p || !g || !m || !v || !seg_end || !seg_lr || !seg_wd || !step_ptr) return RORL_ERR_ARG;
    if (nseg <= 0 || nseg > kMaxSeg || n <= 0) return RORL_ERR_SHAPE;
    int64_t nb = (n + kOptThreads - 1) / kOptThreads;
    if (nb > 148 * 8) nb = 148 * 8;
    adamw_polyak_kernel<<<(unsigned)nb, kOptThreads, 0, stream>>>(p, g, m, v, target, seg_end, seg_lr, seg_wd, seg_clip,
                                                                 (int)nseg, n, beta1, beta2, eps, tau, step_ptr, gnorm_sq,
                                                                 max_norm);
    step_inc_kernel<<<1, 1, 0, stream>>>(step_ptr);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
