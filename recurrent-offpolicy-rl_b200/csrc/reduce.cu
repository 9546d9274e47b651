// Deterministic row reductions of the backward pass, sm_100a.  All HBM-bound: every input byte is read once,
// coalesced as float4, and per-CTA partials are summed by a second tiny launch (no atomics, so gradients are
// bit-reproducible -- the reference's CUDA backward is not, ref: results.md:4).
//
//   colsum          out[n] = sum_m x[m, n]                      bias gradients (ref: autograd of F.linear /
//                                                               EnsembleLinear bias, ensemble_linear_model.py:49-58)
//                                                               and sums of per-CTA partial tiles (split-K wgrad,
//                                                               selective-scan dB/dC/dA, conv dW)
//   elu_bwd_colsum  g = dy * (y > 0 ? 1 : y + 1); out[n] = sum_m g[m, n]
//                                                               ELU backward from the layer OUTPUT fused with the
//                                                               bias gradient of the same layer
//   skinny_wgrad    dW[n, k] = sum_m g[m, n] * x[m, k], K <= 16 weight gradient of the narrow-input projections
//                                                               (obs / action encoders K = 6, 9; dt_proj K = dt_rank)
//                                                               that cuBLAS runs as one-CTA SIMT sgemms (110 us each)
#include "common.cuh"

namespace rorl {

constexpr int kRedThreads = 256;
constexpr int kRedMaxQuads = 256;      // column quads (float4) handled by one CTA in x
constexpr int kRedMaxBlocks = 592;     // 4 x 148 row blocks
constexpr int kRedTickets = 4096;      // ticket slots a caller provides for the single-launch form (one per column chunk x group)

// grid: (row blocks, column chunks of 4 * kRedMaxQuads).  thread = (row lane, column quad).
// ELU: 0 = plain column sum; 1 = ELU backward from the layer OUTPUT y; 2 = exact-GELU backward from the PRE-activation y
template <int ELU>
__global__ void __launch_bounds__(kRedThreads) colsum_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                             float* __restrict__ g, float* __restrict__ partial,
                                                             int64_t M, int64_t N, int64_t ldx, int64_t ldy, int64_t ldg,
                                                             int64_t gsx, int64_t gsy, int64_t gsg, int rows_per_block,
                                                             float* __restrict__ out, int* __restrict__ tickets, int chunk_quads) {
    __shared__ float4 s_acc[kRedThreads];
    __shared__ int s_last;
    x += (int64_t)blockIdx.z * gsx;                                        // group (ensemble member)
    if (ELU) { y += (int64_t)blockIdx.z * gsy; g += (int64_t)blockIdx.z * gsg; }
    partial += (int64_t)blockIdx.z * gridDim.x * N;
    const int64_t q0 = (int64_t)blockIdx.y * chunk_quads;
    const int nq = (int)min((int64_t)chunk_quads, N / 4 - q0);           // quads in this chunk
    const int lanes = kRedThreads / nq;                                    // row lanes (>= 1)
    const int q = threadIdx.x % nq, rl = threadIdx.x / nq;
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t m1 = min(M, m0 + rows_per_block);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < lanes) {
        const int64_t col = (q0 + q) * 4;
        auto elu_grad = [&](float4& v, const float4& o) {
            if (ELU == 2) {                        // d/dx [x Phi(x)] = Phi(x) + x phi(x), Phi via erff (torch.nn.GELU default)
                auto dg = [](float x) { return 0.5f * (1.0f + erff(x * 0.70710678118654752f)) + x * 0.39894228040143268f * __expf(-0.5f * x * x); };
                v.x *= dg(o.x); v.y *= dg(o.y); v.z *= dg(o.z); v.w *= dg(o.w);
            } else {
                v.x *= o.x > 0.f ? 1.f : o.x + 1.f;
                v.y *= o.y > 0.f ? 1.f : o.y + 1.f;
                v.z *= o.z > 0.f ? 1.f : o.z + 1.f;
                v.w *= o.w > 0.f ? 1.f : o.w + 1.f;
            }
        };
        int64_t m = m0 + rl;
        // U rows per iteration: all loads are issued before the first use (memory-level parallelism; a one-row loop left
        // each thread with a single 16-byte load in flight and ran at a quarter of the HBM rate)
        constexpr int U = ELU ? 4 : 8;
        for (; m + (U - 1) * (int64_t)lanes < m1; m += U * (int64_t)lanes) {
            float4 v[U], o[ELU ? U : 1];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(x + (m + u * lanes) * ldx + col));
            if (ELU) {
#pragma unroll
                for (int u = 0; u < U; ++u) o[u] = __ldcs(reinterpret_cast<const float4*>(y + (m + u * lanes) * ldy + col));
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    elu_grad(v[u], o[u]);
                    *reinterpret_cast<float4*>(g + (m + u * lanes) * ldg + col) = v[u];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; m < m1; m += lanes) {
            float4 v = __ldcs(reinterpret_cast<const float4*>(x + m * ldx + col));
            if (ELU) {
                const float4 o = __ldcs(reinterpret_cast<const float4*>(y + m * ldy + col));
                elu_grad(v, o);
                *reinterpret_cast<float4*>(g + m * ldg + col) = v;
            }
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 v = s_acc[l * nq + q];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(partial + (int64_t)blockIdx.x * N + (q0 + q) * 4) = acc;
    }
    if (tickets == nullptr || gridDim.x == 1) return;
    // Single-launch form: the LAST CTA of this (column chunk, group) to publish its partial row folds all rows, in a
    // fixed order that does not depend on which CTA that is (deterministic), and re-arms the ticket.  (The separate
    // second-stage launch this replaces cost ~5 us of launch latency for a few kilobytes, ~60 times per update.)
    __threadfence();
    __syncthreads();
    int* ticket = tickets + blockIdx.z * gridDim.y + blockIdx.y;
    if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1) == (int)gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < lanes) {
        const float* pp = partial + (q0 + q) * 4;
        int bI = rl;
        for (; bI + 7 * lanes < (int)gridDim.x; bI += 8 * lanes) {          // 8 partial rows in flight: this is the latency tail
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const float4*>(pp + (int64_t)(bI + u * lanes) * N));
#pragma unroll
            for (int u = 0; u < 8; ++u) { tot.x += v[u].x; tot.y += v[u].y; tot.z += v[u].z; tot.w += v[u].w; }
        }
        for (; bI < (int)gridDim.x; bI += lanes) {
            const float4 v = __ldcg(reinterpret_cast<const float4*>(pp + (int64_t)bI * N));
            tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
    }
    __syncthreads();
    s_acc[threadIdx.x] = tot;
    __syncthreads();
    if (rl == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 v = s_acc[l * nq + q];
            tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        *reinterpret_cast<float4*>(out + (int64_t)blockIdx.z * N + (q0 + q) * 4) = tot;
    }
    if (threadIdx.x == 0) *ticket = 0;
}

// out[n] = sum_b partial[b, n]: a CTA owns 32 column quads, 8 lanes split the nblk partial rows, shared-memory combine
// (a single thread walking hundreds of dependent-latency loads per column made this tiny stage the longer one)
__global__ void __launch_bounds__(kRedThreads) partial_sum_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                                  int nblk, int64_t N) {
    __shared__ float4 s_acc[kRedThreads];
    partial += (int64_t)blockIdx.y * nblk * N;
    out += (int64_t)blockIdx.y * N;
    const int ql = threadIdx.x & 31, rl = threadIdx.x >> 5;               // 32 quads x 8 row lanes
    const int64_t q = (int64_t)blockIdx.x * 32 + ql;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q * 4 < N) {
        const float* pp = partial + q * 4;
        int b = rl;
        for (; b + 56 < nblk; b += 64) {                      // eight partial rows in flight per thread: the chain is latency
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(pp + (int64_t)(b + 8 * u) * N));
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; b < nblk; b += 8) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(pp + (int64_t)b * N));
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
    }
    s_acc[threadIdx.x] = acc;
    __syncthreads();
    if (rl == 0 && q * 4 < N) {
#pragma unroll
        for (int l = 1; l < 8; ++l) {
            const float4 v = s_acc[l * 32 + ql];
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        *reinterpret_cast<float4*>(out + q * 4) = acc;
    }
}

// dW[n, k] = sum_m g[m, n] x[m, k].  grid: (row blocks, column chunks of 256); thread = output column n; x rows are
// staged in shared memory and read back as warp-wide broadcasts.
constexpr int kSkinnyRows = 64;
template <int KP>        // K padded to a multiple of 4, <= 16
__global__ void __launch_bounds__(kRedThreads) skinny_wgrad_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                                   float* __restrict__ partial, int64_t M, int64_t N, int K,
                                                                   int64_t ldg, int64_t ldx, int64_t seg_rows, int64_t seg_stride,
                                                                   int rows_per_block) {
    __shared__ __align__(16) float s_x[kSkinnyRows * KP];
    const int64_t n = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t m1 = min(M, m0 + rows_per_block);
    float acc[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) acc[k] = 0.f;
    for (int64_t mb = m0; mb < m1; mb += kSkinnyRows) {
        const int rows = (int)min((int64_t)kSkinnyRows, m1 - mb);
        __syncthreads();
        for (int i = threadIdx.x; i < kSkinnyRows * KP; i += blockDim.x) {
            const int r = i / KP, k = i % KP;
            const int64_t m = mb + r;
            s_x[i] = (r < rows && k < K) ? __ldg(x + (m / seg_rows) * seg_stride + (m % seg_rows) * ldx + k) : 0.f;
        }
        __syncthreads();
        if (n < N) {
            for (int r0 = 0; r0 < rows; r0 += 8) {
                float gv[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) gv[u] = (r0 + u < rows) ? __ldg(g + (mb + r0 + u) * ldg + n) : 0.f;   // 8 loads in flight
#pragma unroll
                for (int u = 0; u < 8; ++u) {
#pragma unroll
                    for (int k4 = 0; k4 < KP; k4 += 4) {
                        const float4 xv = *reinterpret_cast<const float4*>(s_x + (r0 + u) * KP + k4);     // rows past `rows` are zero
                        acc[k4] = fmaf(gv[u], xv.x, acc[k4]);
                        acc[k4 + 1] = fmaf(gv[u], xv.y, acc[k4 + 1]);
                        acc[k4 + 2] = fmaf(gv[u], xv.z, acc[k4 + 2]);
                        acc[k4 + 3] = fmaf(gv[u], xv.w, acc[k4 + 3]);
                    }
                }
            }
        }
    }
    if (n < N) {
        float* o = partial + ((int64_t)blockIdx.x * N + n) * KP;
#pragma unroll
        for (int k4 = 0; k4 < KP; k4 += 4)
            *reinterpret_cast<float4*>(o + k4) = make_float4(acc[k4], acc[k4 + 1], acc[k4 + 2], acc[k4 + 3]);
    }
}

// y[m, n] = bias[n] + sum_k x[m, k] W[n, k], K <= 16: the forward of the narrow-input projections.  thread = 4 output
// columns with their W rows in registers; x rows are staged in shared memory and read back as broadcasts; the only
// real traffic is the coalesced float4 store of y (HBM-bound on the output).  128-thread CTAs: `ct` threads across the
// columns (N / 4 rounded up to a warp) times 128 / ct row lanes that take alternate rows of the staged block, so that a
// 128-wide output still runs four warps per CTA (one-warp CTAs left the SMs at 3 warps each: 16 us for 16.7 MB).
// Row m of x is at x + (m / seg_rows) * seg_stride + (m % seg_rows) * ldx: a [B, L, K] slice of a wider / longer
// batch tensor is read in place (no contiguous copy).
template <int KP>
__global__ void __launch_bounds__(128) skinny_linear_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                            const float* __restrict__ bias, float* __restrict__ y, int64_t M,
                                                            int64_t N, int K, int64_t ldx, int64_t seg_rows, int64_t seg_stride,
                                                            int64_t ldy, int rows_per_block, int elu, int ct) {
    __shared__ __align__(16) float s_x[kSkinnyRows * KP];
    const int cx = threadIdx.x % ct, rl = threadIdx.x / ct, RL = 128 / ct;
    const int64_t n0 = ((int64_t)blockIdx.y * ct + cx) * 4;
    const bool act = n0 < N && rl < RL;
    float w[4][KP], b4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < KP; ++k) w[j][k] = (act && k < K) ? __ldg(W + (n0 + j) * K + k) : 0.f;
        if (act && bias) b4[j] = __ldg(bias + n0 + j);
    }
    const int64_t m0 = (int64_t)blockIdx.x * rows_per_block;
    const int64_t m1 = min(M, m0 + rows_per_block);
    for (int64_t mb = m0; mb < m1; mb += kSkinnyRows) {
        const int rows = (int)min((int64_t)kSkinnyRows, m1 - mb);
        __syncthreads();
        for (int i = threadIdx.x; i < kSkinnyRows * KP; i += 128) {
            const int r = i / KP, k = i % KP;
            const int64_t m = mb + r;
            s_x[i] = (r < rows && k < K) ? __ldg(x + (m / seg_rows) * seg_stride + (m % seg_rows) * ldx + k) : 0.f;
        }
        __syncthreads();
        if (act) {
#pragma unroll 2
            for (int r = rl; r < rows; r += RL) {
                float o[4] = {b4[0], b4[1], b4[2], b4[3]};
#pragma unroll
                for (int k4 = 0; k4 < KP; k4 += 4) {
                    const float4 xv = *reinterpret_cast<const float4*>(s_x + r * KP + k4);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[j] = fmaf(xv.x, w[j][k4], o[j]);
                        o[j] = fmaf(xv.y, w[j][k4 + 1], o[j]);
                        o[j] = fmaf(xv.z, w[j][k4 + 2], o[j]);
                        o[j] = fmaf(xv.w, w[j][k4 + 3], o[j]);
                    }
                }
                if (elu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = o[j] > 0.f ? o[j] : ex2f(o[j] * kLog2e) - 1.0f;
                }
                *reinterpret_cast<float4*>(y + (mb + r) * ldy + n0) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// dx[m, k] = sum_n g[m, n] W[n, k], K <= 16: the data gradient of the same projections (the actor's path into the
// critic's action encoder).  One warp per row: lane l takes columns l, l + 32, ... of g (coalesced) against W staged
// in shared memory, K-wide partials combined by shuffles, lanes 0..K-1 store.
template <int KP>
__global__ void __launch_bounds__(256) skinny_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ W,
                                                           float* __restrict__ dx, int64_t M, int N, int K, int64_t ldg) {
    extern __shared__ float s_w[];                         // [N][KP]
    for (int i = threadIdx.x; i < N * KP; i += 256) {
        const int n = i / KP, k = i % KP;
        s_w[i] = k < K ? __ldg(W + (int64_t)n * K + k) : 0.f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarp = (int64_t)gridDim.x * 8;
    for (int64_t m = warp; m < M; m += nwarp) {
        float acc[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) acc[k] = 0.f;
        for (int n = lane; n < N; n += 32) {
            const float gv = __ldg(g + m * ldg + n);
#pragma unroll
            for (int k = 0; k < KP; ++k) acc[k] = fmaf(gv, s_w[n * KP + k], acc[k]);
        }
        float mine = 0.f;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            float v = acc[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == k) mine = v;
        }
        if (lane < K) dx[m * K + lane] = mine;
    }
}

// out[r * out_ld + c] = sum_p part[p][r * C + c]: per-tile partial rows summed straight into a column block of a wider
// row-major buffer (the selective scan's dB | dC partials -> their columns of d(x_dbl); no slice-gradient copies)
__global__ void __launch_bounds__(256) sum_leading_rows_kernel(const float* __restrict__ part, float* __restrict__ out, int P,
                                                               int64_t rows, int C, int64_t out_ld) {
    const int cq = C / 4;
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= rows * cq) return;
    const int64_t r = i / cq;
    const int c = (int)(i % cq) * 4;
    const float* src = part + r * C + c;
    const int64_t plane = rows * C;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    int pI = 0;
    for (; pI + 8 <= P; pI += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(src + (pI + u) * plane));
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    for (; pI < P; ++pI) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(src + pI * plane));
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(out + r * out_ld + c) = acc;
}

static inline int red_blocks(int64_t M, int* rows_per_block, int64_t G = 1, int64_t chunks = 1) {
    int nblk = (int)((M + 63) / 64);                     // at least 64 rows per block
    int cap = (int)(kRedMaxBlocks / (G * chunks));       // about 4 CTAs per SM over the whole grid
    if (cap < 1) cap = 1;
    if (nblk > cap) nblk = cap;
    if (nblk < 1) nblk = 1;
    const int rpb = (int)((M + nblk - 1) / nblk);
    *rows_per_block = rpb;
    return (int)((M + rpb - 1) / rpb);
}

// Grid of the column sums.  Single-launch (ticket) form: 128-column chunks and about two CTAs per SM over the whole
// grid -- the last CTA's fold over the row blocks is a latency chain (row blocks / row lanes dependent-free loads), so
// few, long row blocks and many row lanes (256 / 32 quads = 8) keep it to a few microseconds.  Two-launch form: wide
// chunks, about four CTAs per SM.
static inline void colsum_geometry(int64_t G, int64_t M, int64_t N, bool tickets, int* chunk_quads, int* chunks, int* nblk, int* rpb,
                                   int target = 296) {
    const int64_t quads = N / 4;
    *chunk_quads = (tickets && quads > 32) ? 32 : (int)(quads < kRedMaxQuads ? (quads > 0 ? quads : 1) : kRedMaxQuads);
    *chunks = (int)((quads + *chunk_quads - 1) / *chunk_quads);
    if (tickets) {
        int nb = (int)((M + 63) / 64);
        int cap = (int)(target / (G * *chunks));
        if (cap < 1) cap = 1;
        if (nb > cap) nb = cap;
        const int r = (int)((M + nb - 1) / nb);
        *rpb = r;
        *nblk = (int)((M + r - 1) / r);
    } else {
        *nblk = red_blocks(M, rpb, G, *chunks);
    }
}

}  // namespace rorl

using namespace rorl;

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" {

int rorl_colsum_tickets(void) { return kRedTickets; }

int64_t rorl_colsum_work_floats(int64_t G, int64_t M, int64_t N) {
    int rpb;
    return G * (int64_t)red_blocks(M, &rpb) * N;
}

int rorl_colsum(const float* x, float* out, float* work, int64_t G, int64_t M, int64_t N, int64_t ldx, int64_t gsx,
                int32_t* tickets, cudaStream_t stream) {
    if (!x || !out || !work) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || G <= 0 || G > 65535) return RORL_ERR_SHAPE;
    if (N % 4 || ldx % 4 || gsx % 4 || !a16(x) || !a16(out) || !a16(work)) return RORL_ERR_ALIGN;
    int rpb, cq, chunks, nblk;
    colsum_geometry(G, M, N, tickets != nullptr, &cq, &chunks, &nblk, &rpb);
    if (tickets && (int64_t)chunks * G > kRedTickets) {
        tickets = nullptr;
        colsum_geometry(G, M, N, false, &cq, &chunks, &nblk, &rpb);
    }
    if (chunks > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)nblk, (unsigned)chunks, (unsigned)G);
    colsum_kernel<0><<<grid, kRedThreads, 0, stream>>>(x, nullptr, nullptr, nblk == 1 ? out : work, M, N, ldx, 0, 0, gsx, 0, 0, rpb,
                                                           out, tickets, cq);
    if (nblk > 1 && !tickets) {
        dim3 g2((unsigned)((N / 4 + 31) / 32), (unsigned)G);
        partial_sum_kernel<<<g2, kRedThreads, 0, stream>>>(work, out, nblk, N);
    }
    RORL_RETURN_LAUNCH();
}

static int act_bwd_colsum(int g_act_mode, const float* dy, const float* y, float* g, float* out, float* work, int64_t G, int64_t M,
                          int64_t N, int64_t ld_dy, int64_t ld_y, int64_t ld_g, int64_t gs_dy, int64_t gs_y, int64_t gs_g,
                          int32_t* tickets, cudaStream_t stream) {
    if (!dy || !y || !g || !out || !work) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || G <= 0 || G > 65535) return RORL_ERR_SHAPE;
    if (N % 4 || ld_dy % 4 || ld_y % 4 || ld_g % 4 || gs_dy % 4 || gs_y % 4 || gs_g % 4 || !a16(dy) || !a16(y) || !a16(g) ||
        !a16(out) || !a16(work))
        return RORL_ERR_ALIGN;
    int rpb, cq, chunks, nblk;
    colsum_geometry(G, M, N, tickets != nullptr, &cq, &chunks, &nblk, &rpb, 592);   // three streams per element: bandwidth, not the fold
    if (tickets && (int64_t)chunks * G > kRedTickets) {
        tickets = nullptr;
        colsum_geometry(G, M, N, false, &cq, &chunks, &nblk, &rpb);
    }
    if (chunks > 65535) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)nblk, (unsigned)chunks, (unsigned)G);
    if (g_act_mode == 2)
        colsum_kernel<2><<<grid, kRedThreads, 0, stream>>>(dy, y, g, nblk == 1 ? out : work, M, N, ld_dy, ld_y, ld_g, gs_dy, gs_y, gs_g, rpb,
                                                           out, tickets, cq);
    else
        colsum_kernel<1><<<grid, kRedThreads, 0, stream>>>(dy, y, g, nblk == 1 ? out : work, M, N, ld_dy, ld_y, ld_g, gs_dy, gs_y, gs_g, rpb,
                                                           out, tickets, cq);
    if (nblk > 1 && !tickets) {
        dim3 g2((unsigned)((N / 4 + 31) / 32), (unsigned)G);
        partial_sum_kernel<<<g2, kRedThreads, 0, stream>>>(work, out, nblk, N);
    }
    RORL_RETURN_LAUNCH();
}

int rorl_sum_leading_rows(const float* part, float* out, int64_t P, int64_t rows, int64_t C, int64_t out_ld, cudaStream_t stream) {
    if (!part || !out) return RORL_ERR_ARG;
    if (P <= 0 || rows <= 0 || C <= 0 || out_ld < C) return RORL_ERR_SHAPE;
    if (C % 4 || out_ld % 4 || !a16(part) || !a16(out)) return RORL_ERR_ALIGN;
    const int64_t n = rows * (C / 4);
    sum_leading_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(part, out, (int)P, rows, (int)C, out_ld);
    RORL_RETURN_LAUNCH();
}

int rorl_elu_bwd_colsum(const float* dy, const float* y, float* g, float* out, float* work, int64_t G, int64_t M,
                        int64_t N, int64_t ld_dy, int64_t ld_y, int64_t ld_g, int64_t gs_dy, int64_t gs_y, int64_t gs_g,
                        int32_t* tickets, cudaStream_t stream) {
    return act_bwd_colsum(1, dy, y, g, out, work, G, M, N, ld_dy, ld_y, ld_g, gs_dy, gs_y, gs_g, tickets, stream);
}

int rorl_gelu_bwd_colsum(const float* dy, const float* pre, float* g, float* out, float* work, int64_t G, int64_t M,
                         int64_t N, int64_t ld_dy, int64_t ld_pre, int64_t ld_g, int64_t gs_dy, int64_t gs_pre, int64_t gs_g,
                         int32_t* tickets, cudaStream_t stream) {
    return act_bwd_colsum(2, dy, pre, g, out, work, G, M, N, ld_dy, ld_pre, ld_g, gs_dy, gs_pre, gs_g, tickets, stream);
}

int64_t rorl_skinny_wgrad_work_floats(int64_t M, int64_t N, int64_t K) {
    int rpb;
    const int64_t KP = (K + 3) / 4 * 4;
    return (int64_t)red_blocks(M, &rpb) * N * KP;
}

int rorl_skinny_wgrad(const float* g, const float* x, float* dW, float* work, int64_t M, int64_t N, int64_t K,
                      int64_t ldg, int64_t ldx, int64_t seg_rows, int64_t seg_stride, cudaStream_t stream) {
    if (!g || !x || !dW || !work) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || K <= 0 || K > 16 || seg_rows <= 0) return RORL_ERR_SHAPE;
    if (!a16(work) || !a16(dW)) return RORL_ERR_ALIGN;
    int rpb;
    int threads = (int)((N + 31) / 32 * 32);                // one thread per output column
    if (threads > kRedThreads) threads = kRedThreads;
    const int chunks = (int)((N + threads - 1) / threads);
    const int nblk = red_blocks(M, &rpb, 1, chunks);
    const int KP = (int)((K + 3) / 4 * 4);
    dim3 grid((unsigned)nblk, (unsigned)chunks);
    float* dst = nblk == 1 ? dW : work;
    switch (KP) {
        case 4: skinny_wgrad_kernel<4><<<grid, threads, 0, stream>>>(g, x, dst, M, N, (int)K, ldg, ldx, seg_rows, seg_stride, rpb); break;
        case 8: skinny_wgrad_kernel<8><<<grid, threads, 0, stream>>>(g, x, dst, M, N, (int)K, ldg, ldx, seg_rows, seg_stride, rpb); break;
        case 12: skinny_wgrad_kernel<12><<<grid, threads, 0, stream>>>(g, x, dst, M, N, (int)K, ldg, ldx, seg_rows, seg_stride, rpb); break;
        default: skinny_wgrad_kernel<16><<<grid, threads, 0, stream>>>(g, x, dst, M, N, (int)K, ldg, ldx, seg_rows, seg_stride, rpb); break;
    }
    if (nblk > 1) {
        dim3 g2((unsigned)((N * KP / 4 + 31) / 32), 1u);
        partial_sum_kernel<<<g2, kRedThreads, 0, stream>>>(work, dW, nblk, N * KP);
    }
    RORL_RETURN_LAUNCH();
}

int rorl_skinny_linear(const float* x, const float* W, const float* bias, float* y, int64_t M, int64_t N, int64_t K,
                       int64_t ldx, int64_t seg_rows, int64_t seg_stride, int64_t ldy, int elu, cudaStream_t stream) {
    if (!x || !W || !y) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || K <= 0 || K > 16 || seg_rows <= 0) return RORL_ERR_SHAPE;
    if (N % 4 || ldy % 4 || !a16(y)) return RORL_ERR_ALIGN;
    int ct = (int)((N / 4 + 31) / 32 * 32);                 // threads across the columns; 128 / ct row lanes (see the kernel)
    if (ct > 128) ct = 128;
    const int chunks = (int)((N / 4 + ct - 1) / ct);
    int nblk = (int)((M + kSkinnyRows - 1) / kSkinnyRows);
    const int cap = 148 * 8 / (chunks > 0 ? chunks : 1);
    if (nblk > cap) nblk = cap > 0 ? cap : 1;
    int rpb = (int)((M + nblk - 1) / nblk);
    rpb = (rpb + kSkinnyRows - 1) / kSkinnyRows * kSkinnyRows;
    nblk = (int)((M + rpb - 1) / rpb);
    const int KP = (int)((K + 3) / 4 * 4);
    dim3 grid((unsigned)nblk, (unsigned)chunks);
    switch (KP) {
        case 4: skinny_linear_kernel<4><<<grid, 128, 0, stream>>>(x, W, bias, y, M, N, (int)K, ldx, seg_rows, seg_stride, ldy, rpb, elu, ct); break;
        case 8: skinny_linear_kernel<8><<<grid, 128, 0, stream>>>(x, W, bias, y, M, N, (int)K, ldx, seg_rows, seg_stride, ldy, rpb, elu, ct); break;
        case 12: skinny_linear_kernel<12><<<grid, 128, 0, stream>>>(x, W, bias, y, M, N, (int)K, ldx, seg_rows, seg_stride, ldy, rpb, elu, ct); break;
        default: skinny_linear_kernel<16><<<grid, 128, 0, stream>>>(x, W, bias, y, M, N, (int)K, ldx, seg_rows, seg_stride, ldy, rpb, elu, ct); break;
    }
    RORL_RETURN_LAUNCH();
}

int rorl_skinny_dgrad(const float* g, const float* W, float* dx, int64_t M, int64_t N, int64_t K, int64_t ldg, cudaStream_t stream) {
    if (!g || !W || !dx) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || N > 1024 || K <= 0 || K > 16) return RORL_ERR_SHAPE;
    const int KP = (int)((K + 3) / 4 * 4);
    int64_t nb = (M + 7) / 8;
    if (nb > 148 * 8) nb = 148 * 8;
    const size_t smem = (size_t)N * KP * sizeof(float);
    switch (KP) {
        case 4: skinny_dgrad_kernel<4><<<(unsigned)nb, 256, smem, stream>>>(g, W, dx, M, (int)N, (int)K, ldg); break;
        case 8: skinny_dgrad_kernel<8><<<(unsigned)nb, 256, smem, stream>>>(g, W, dx, M, (int)N, (int)K, ldg); break;
        case 12: skinny_dgrad_kernel<12><<<(unsigned)nb, 256, smem, stream>>>(g, W, dx, M, (int)N, (int)K, ldg); break;
        default: skinny_dgrad_kernel<16><<<(unsigned)nb, 256, smem, stream>>>(g, W, dx, M, (int)N, (int)K, ldg); break;
    }
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
