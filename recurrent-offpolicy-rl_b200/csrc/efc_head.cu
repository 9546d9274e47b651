// Output layer of the ensemble-Q head: E independent linear maps hidden -> 1 (the third `efc-E` layer of the value
// network, ref: offpolicy_rnn/models/ensemble_linear_model.py:36-49 with out_dim = 1, called from
// contextual_model.py:97-116), and its backward fused with the ELU backward + bias gradient of the layer below.
//
//   fwd : q[e, m]   = sum_k y[e, m, k] w[e, k] + b[e]
//   bwd : g[e, m, k] = dq[e, m] w[e, k] elu'(y[e, m, k])      (elu' from the layer OUTPUT: y > 0 ? 1 : y + 1)
//         dw[e, k]  = sum_m dq[e, m] y[e, m, k],   db[e] = sum_m dq[e, m],   dbias_below[e, k] = sum_m g[e, m, k]
//
// The reference runs these as bmm [E, M, K] x [E, K, 1] (a GEMV per member) plus, in the backward, an outer-product bmm,
// elu_backward and two reductions: five passes over the [E, M, K] hidden activation (267 MB at the benchmark shape).
// Here the forward is one read of y and the backward one read of y + one write of g: both HBM-bound.
// Deterministic: per-CTA partial sums in global memory, summed by the caller (rorl_colsum); no atomics.
#include "common.cuh"

namespace rorl {

constexpr int kHeadThreads = 256;                       // 8 warps; a warp owns one row at a time

// KQ = K / 128: float4 column quads per lane (lane l holds columns 4 l + 128 j, j < KQ)
template <int KQ>
__global__ void __launch_bounds__(kHeadThreads) efc_dot_fwd_kernel(const float* __restrict__ y, const float* __restrict__ w,
                                                                   const float* __restrict__ b, float* __restrict__ q, int64_t M) {
    constexpr int K = KQ * 128;
    const int e = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 wv[KQ];
#pragma unroll
    for (int j = 0; j < KQ; ++j) wv[j] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)e * K + 4 * lane + 128 * j));
    const float bias = b ? __ldg(b + e) : 0.f;
    const float* ye = y + (int64_t)e * M * K;
    const int64_t stride = (int64_t)gridDim.x * (kHeadThreads / 32);
    int64_t m = (int64_t)blockIdx.x * (kHeadThreads / 32) + warp;
    // two rows per iteration: both rows' loads are in flight before the first shuffle
    for (; m + stride < M; m += 2 * stride) {
        float4 a[KQ], c[KQ];
#pragma unroll
        for (int j = 0; j < KQ; ++j) a[j] = __ldcs(reinterpret_cast<const float4*>(ye + m * K + 4 * lane + 128 * j));
#pragma unroll
        for (int j = 0; j < KQ; ++j) c[j] = __ldcs(reinterpret_cast<const float4*>(ye + (m + stride) * K + 4 * lane + 128 * j));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            s0 += a[j].x * wv[j].x + a[j].y * wv[j].y + a[j].z * wv[j].z + a[j].w * wv[j].w;
            s1 += c[j].x * wv[j].x + c[j].y * wv[j].y + c[j].z * wv[j].z + c[j].w * wv[j].w;
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        if (lane == 0) {
            q[(int64_t)e * M + m] = s0 + bias;
            q[(int64_t)e * M + m + stride] = s1 + bias;
        }
    }
    for (; m < M; m += stride) {
        float s0 = 0.f;
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            const float4 a = __ldcs(reinterpret_cast<const float4*>(ye + m * K + 4 * lane + 128 * j));
            s0 += a.x * wv[j].x + a.y * wv[j].y + a.z * wv[j].z + a.w * wv[j].w;
        }
        s0 = warp_sum(s0);
        if (lane == 0) q[(int64_t)e * M + m] = s0 + bias;
    }
}

// grid (row blocks, E).  part[e][blk][2K + 4]: dw | dbias_below | (db, 0, 0, 0)
template <int KQ, bool ELU>
__global__ void __launch_bounds__(kHeadThreads) efc_head_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ y,
                                                                    const float* __restrict__ w, float* __restrict__ g,
                                                                    float* __restrict__ part, int64_t M) {
    constexpr int K = KQ * 128;
    __shared__ float4 s_acc[kHeadThreads / 32][2 * KQ][32];
    __shared__ float s_db[kHeadThreads / 32];
    const int e = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4 wv[KQ], aw[KQ], ab[KQ];
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        wv[j] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)e * K + 4 * lane + 128 * j));
        aw[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        ab[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float adb = 0.f;
    const float* ye = y + (int64_t)e * M * K;
    float* ge = g + (int64_t)e * M * K;
    const float* dqe = dq + (int64_t)e * M;
    const int64_t stride = (int64_t)gridDim.x * (kHeadThreads / 32);
    auto one = [&](const float4 (&a)[KQ], float d, int64_t m) {
#pragma unroll
        for (int j = 0; j < KQ; ++j) {
            float4 o;
            o.x = d * wv[j].x * (ELU ? (a[j].x > 0.f ? 1.f : a[j].x + 1.f) : 1.f);
            o.y = d * wv[j].y * (ELU ? (a[j].y > 0.f ? 1.f : a[j].y + 1.f) : 1.f);
            o.z = d * wv[j].z * (ELU ? (a[j].z > 0.f ? 1.f : a[j].z + 1.f) : 1.f);
            o.w = d * wv[j].w * (ELU ? (a[j].w > 0.f ? 1.f : a[j].w + 1.f) : 1.f);
            *reinterpret_cast<float4*>(ge + m * K + 4 * lane + 128 * j) = o;
            aw[j].x = fmaf(d, a[j].x, aw[j].x); aw[j].y = fmaf(d, a[j].y, aw[j].y);
            aw[j].z = fmaf(d, a[j].z, aw[j].z); aw[j].w = fmaf(d, a[j].w, aw[j].w);
            ab[j].x += o.x; ab[j].y += o.y; ab[j].z += o.z; ab[j].w += o.w;
        }
        adb += d;
    };
    int64_t m = (int64_t)blockIdx.x * (kHeadThreads / 32) + warp;
    for (; m + stride < M; m += 2 * stride) {
        float4 a[KQ], c[KQ];
#pragma unroll
        for (int j = 0; j < KQ; ++j) a[j] = __ldcs(reinterpret_cast<const float4*>(ye + m * K + 4 * lane + 128 * j));
#pragma unroll
        for (int j = 0; j < KQ; ++j) c[j] = __ldcs(reinterpret_cast<const float4*>(ye + (m + stride) * K + 4 * lane + 128 * j));
        const float d0 = __ldg(dqe + m), d1 = __ldg(dqe + m + stride);
        one(a, d0, m);
        one(c, d1, m + stride);
    }
    for (; m < M; m += stride) {
        float4 a[KQ];
#pragma unroll
        for (int j = 0; j < KQ; ++j) a[j] = __ldcs(reinterpret_cast<const float4*>(ye + m * K + 4 * lane + 128 * j));
        one(a, __ldg(dqe + m), m);
    }
    // combine the CTA's 8 warps in a fixed order
#pragma unroll
    for (int j = 0; j < KQ; ++j) {
        s_acc[warp][j][lane] = aw[j];
        s_acc[warp][KQ + j][lane] = ab[j];
    }
    if (lane == 0) s_db[warp] = adb;            // every lane of a warp saw the same rows
    __syncthreads();
    float* pe = part + ((int64_t)e * gridDim.x + blockIdx.x) * (2 * K + 4);
    for (int i = threadIdx.x; i < 2 * KQ * 32; i += kHeadThreads) {
        const int jj = i / 32, l = i % 32;
        float4 s = s_acc[0][jj][l];
#pragma unroll
        for (int wq = 1; wq < kHeadThreads / 32; ++wq) {
            const float4 v = s_acc[wq][jj][l];
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        const int half = jj / KQ, j = jj % KQ;
        *reinterpret_cast<float4*>(pe + half * K + 4 * l + 128 * j) = s;
    }
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int wq = 0; wq < kHeadThreads / 32; ++wq) s += s_db[wq];
        *reinterpret_cast<float4*>(pe + 2 * K) = make_float4(s, 0.f, 0.f, 0.f);
    }
}

constexpr int kHeadBlocks = 74;                         // row blocks per member: 74 x 8 members = 4 CTAs per SM

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_efc_head_nblk(void) { return kHeadBlocks; }

int rorl_efc_dot_fwd(const float* y, const float* w, const float* b, float* q, int64_t E, int64_t M, int64_t K, cudaStream_t stream) {
    if (!y || !w || !q) return RORL_ERR_ARG;
    if (E <= 0 || M <= 0 || E > 65535) return RORL_ERR_SHAPE;
    if (K != 128 && K != 256 && K != 384 && K != 512) return RORL_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w)) & 15) return RORL_ERR_ALIGN;
    int64_t nb = (M + 15) / 16;
    if (nb > 148 * 2) nb = 148 * 2;
    dim3 grid((unsigned)nb, (unsigned)E);
    switch (K / 128) {
        case 1: efc_dot_fwd_kernel<1><<<grid, kHeadThreads, 0, stream>>>(y, w, b, q, M); break;
        case 2: efc_dot_fwd_kernel<2><<<grid, kHeadThreads, 0, stream>>>(y, w, b, q, M); break;
        case 3: efc_dot_fwd_kernel<3><<<grid, kHeadThreads, 0, stream>>>(y, w, b, q, M); break;
        default: efc_dot_fwd_kernel<4><<<grid, kHeadThreads, 0, stream>>>(y, w, b, q, M); break;
    }
    RORL_RETURN_LAUNCH();
}

int rorl_efc_head_bwd(const float* dq, const float* y, const float* w, float* g, float* part, int64_t E, int64_t M, int64_t K,
                      int elu, cudaStream_t stream) {
    if (!dq || !y || !w || !g || !part) return RORL_ERR_ARG;
    if (E <= 0 || M <= 0 || E > 65535) return RORL_ERR_SHAPE;
    if (K != 128 && K != 256 && K != 384 && K != 512) return RORL_ERR_SHAPE;
    if ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(part)) & 15)
        return RORL_ERR_ALIGN;
    dim3 grid(kHeadBlocks, (unsigned)E);
#define HEAD_LAUNCH(KQ)                                                                                           \
    do {                                                                                                          \
        if (elu) efc_head_bwd_kernel<KQ, true><<<grid, kHeadThreads, 0, stream>>>(dq, y, w, g, part, M);         \
        else efc_head_bwd_kernel<KQ, false><<<grid, kHeadThreads, 0, stream>>>(dq, y, w, g, part, M);            \
    } while (0)
    switch (K / 128) {
        case 1: HEAD_LAUNCH(1); break;
        case 2: HEAD_LAUNCH(2); break;
        case 3: HEAD_LAUNCH(3); break;
        default: HEAD_LAUNCH(4); break;
    }
#undef HEAD_LAUNCH
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
