// GILR real gated linear recurrence  h_t = f_t*h_{t-1} + (1-f_t)*v_t  (h_{-1} = 0), forward and
// backward, as a single-pass tiled parallel scan.
//
// Replaces the reference's serial Triton kernels fwd_sequential_scan / bwd_sequential_scan
// (ref: offpolicy_rnn/models/gilr/scan_triton/real_rnn_tie_input_gate.py:9-33,67-116), which walk
// the L dependent steps one by one with grid (B, C/256).  Here one CTA owns (batch row, 32
// channels) and walks the sequence in tiles of 32*S steps: cp.async stages the [steps x 32ch] tile
// of every operand in shared memory (3-deep ring), each thread scans S steps x 4 channels, the 32
// chunk aggregates (A = prod f, H = local state) are combined with two warp-shuffle rounds plus an
// 8-entry shared-memory carry chain, and the second sweep re-reads the tile from shared memory, so
// HBM sees every operand exactly once: 12 B/element forward, 24 B/element backward.
//
// FUSED variants take the raw pre-activations and the reset flag and apply v = tanh(u_v),
// f = sigmoid(u_f) * (1 - start) in the kernel (ref: offpolicy_rnn/models/gilr/gilr.py:52-56).
#include "common.cuh"

namespace rorl {

constexpr int kScanThreads = 256;
constexpr int kScanCq = 8;        // float4 quads per row  -> 32 channels per CTA
constexpr int kScanChunks = 32;   // chunk-threads per quad
constexpr int kScanStages = 3;

struct F4 {
    float v[4];
};
__device__ __forceinline__ F4 ld4(const float* p) {
    float4 t = *reinterpret_cast<const float4*>(p);
    return F4{{t.x, t.y, t.z, t.w}};
}
__device__ __forceinline__ void st4(float* p, const F4& a) {
    *reinterpret_cast<float4*>(p) = make_float4(a.v[0], a.v[1], a.v[2], a.v[3]);
}
__device__ __forceinline__ F4 shfl_up4(const F4& a, int off) {
    F4 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) r.v[i] = __shfl_up_sync(0xffffffffu, a.v[i], off);
    return r;
}

// Combine the 32 per-thread chunk aggregates of one tile into the state entering this thread's
// chunk, and advance the running tile carry.  (A, H) is this thread's aggregate; `carry` is the
// state entering the tile.  wagg: [2 parities][8 warps][8 quads][2] float4.
__device__ __forceinline__ F4 tile_carry_in(F4 A, F4 H, F4& carry, float* wagg, int par) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, cq = threadIdx.x & 7;
    F4 Ai = A, Hi = H;
#pragma unroll
    for (int off = 8; off <= 16; off <<= 1) {
        F4 Ap = shfl_up4(Ai, off), Hp = shfl_up4(Hi, off);
        if (lane >= off) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                Hi.v[i] = fmaf(Ai.v[i], Hp.v[i], Hi.v[i]);
                Ai.v[i] *= Ap.v[i];
            }
        }
    }
    F4 Ae = shfl_up4(Ai, 8), He = shfl_up4(Hi, 8);
    if (lane < 8) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { Ae.v[i] = 1.0f; He.v[i] = 0.0f; }
    }
    float* base = wagg + (size_t)par * (8 * kScanCq * 8);
    if (lane >= 24) {
        st4(base + (warp * kScanCq + cq) * 8, Ai);
        st4(base + (warp * kScanCq + cq) * 8 + 4, Hi);
    }
    __syncthreads();
    F4 s = carry, s_in = carry;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        F4 Aw = ld4(base + (w * kScanCq + cq) * 8), Hw = ld4(base + (w * kScanCq + cq) * 8 + 4);
        if (w == warp) s_in = s;
#pragma unroll
        for (int i = 0; i < 4; ++i) s.v[i] = fmaf(Aw.v[i], s.v[i], Hw.v[i]);
    }
    carry = s;
    F4 h;
#pragma unroll
    for (int i = 0; i < 4; ++i) h.v[i] = fmaf(Ae.v[i], s_in.v[i], He.v[i]);
    return h;
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
template <int S, bool FUSED>
__global__ void __launch_bounds__(kScanThreads) gilr_fwd_kernel(
    const float* __restrict__ gv, const float* __restrict__ gf, const float* __restrict__ gstart,
    float* __restrict__ gh, int L, int C) {
    constexpr int TL = kScanChunks * S;
    constexpr int ARR = TL * 32;                       // floats per staged operand
    constexpr int STAGE = 2 * ARR + TL;                // v, f, start
    extern __shared__ __align__(16) float smem[];
    float* wagg = smem + kScanStages * STAGE;

    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int tid = threadIdx.x, cq = tid & 7, ck = tid >> 3;
    const bool cvalid = (c0 + cq * 4) < C;
    const size_t rowbase = (size_t)b * L;
    const int ntiles = (L + TL - 1) / TL;

    auto issue = [&](int tile) {
        if (tile < ntiles) {
            float* st = smem + (tile % kScanStages) * STAGE;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                int q = ck + 32 * i, t = tile * TL + q;
                bool ok = cvalid && t < L;
                size_t g = (rowbase + (ok ? t : 0)) * C + c0 + cq * 4;
                cp_async16(st + q * 32 + cq * 4, gv + g, ok);
                cp_async16(st + ARR + q * 32 + cq * 4, gf + g, ok);
            }
            if (FUSED && tid < TL) {
                int t = tile * TL + tid;
                bool ok = (gstart != nullptr) && t < L;
                cp_async4(st + 2 * ARR + tid, gstart + (ok ? rowbase + t : 0), ok);
            }
        }
        cp_async_commit();
    };

    issue(0);
    issue(1);
    F4 carry = {{0.f, 0.f, 0.f, 0.f}};
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<kScanStages - 2>();
        __syncthreads();
        issue(tile + kScanStages - 1);
        float* st = smem + (tile % kScanStages) * STAGE;
        float* sv = st + (ck * S) * 32 + cq * 4;
        float* sf = sv + ARR;
        const int t0 = tile * TL + ck * S;
        F4 A = {{1.f, 1.f, 1.f, 1.f}}, H = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                F4 v = ld4(sv + s * 32), f = ld4(sf + s * 32);
                if (FUSED) {
                    float keep = 1.0f - st[2 * ARR + ck * S + s];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v.v[i] = tanhf_fast(v.v[i]);
                        f.v[i] = sigmoidf_fast(f.v[i]) * keep;
                    }
                    st4(sv + s * 32, v);
                    st4(sf + s * 32, f);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    H.v[i] = fmaf(H.v[i] - v.v[i], f.v[i], v.v[i]);
                    A.v[i] *= f.v[i];
                }
            }
        }
        F4 h = tile_carry_in(A, H, carry, wagg, tile & 1);
        float* out = gh + (rowbase + t0) * C + c0 + cq * 4;
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (t0 + s < L) {
                F4 v = ld4(sv + s * 32), f = ld4(sf + s * 32);
#pragma unroll
                for (int i = 0; i < 4; ++i) h.v[i] = fmaf(h.v[i] - v.v[i], f.v[i], v.v[i]);
                if (cvalid) st4(out + (size_t)s * C, h);
            }
        }
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------
// backward:  G_t = dy_t + f_{t+1} G_{t+1};  dv_t = (1-f_t) G_t;  df_t = G_t (h_{t-1} - v_t)
// (ref: real_rnn_tie_input_gate.py:93-116).  Walks time in reverse; the carried quantity is
// E_t = f_t G_t, a linear recurrence of the same (A, H) form as the forward.
// ---------------------------------------------------------------------------------------------
template <int S, bool FUSED>
__global__ void __launch_bounds__(kScanThreads) gilr_bwd_kernel(
    const float* __restrict__ gdy, const float* __restrict__ gv, const float* __restrict__ gf,
    const float* __restrict__ gh, const float* __restrict__ gstart, float* __restrict__ gdv,
    float* __restrict__ gdf, int L, int C) {
    constexpr int TL = kScanChunks * S;
    constexpr int ARR = TL * 32;
    constexpr int STAGE = 4 * ARR + TL;  // dy, v, f, hprev, start
    extern __shared__ __align__(16) float smem[];
    float* wagg = smem + kScanStages * STAGE;

    const int b = blockIdx.y, c0 = blockIdx.x * 32;
    const int tid = threadIdx.x, cq = tid & 7, ck = tid >> 3;
    const bool cvalid = (c0 + cq * 4) < C;
    const size_t rowbase = (size_t)b * L;
    const int ntiles = (L + TL - 1) / TL;

    auto issue = [&](int tile) {
        if (tile < ntiles) {
            float* st = smem + (tile % kScanStages) * STAGE;
#pragma unroll
            for (int i = 0; i < S; ++i) {
                int q = ck + 32 * i, r = tile * TL + q, t = L - 1 - r;
                bool ok = cvalid && r < L;
                size_t g = (rowbase + (ok ? t : 0)) * C + c0 + cq * 4;
                cp_async16(st + q * 32 + cq * 4, gdy + g, ok);
                cp_async16(st + ARR + q * 32 + cq * 4, gv + g, ok);
                cp_async16(st + 2 * ARR + q * 32 + cq * 4, gf + g, ok);
                bool okp = ok && t > 0;
                cp_async16(st + 3 * ARR + q * 32 + cq * 4, gh + (okp ? g - C : 0), okp);
            }
            if (FUSED && tid < TL) {
                int r = tile * TL + tid;
                bool ok = (gstart != nullptr) && r < L;
                cp_async4(st + 4 * ARR + tid, gstart + (ok ? rowbase + (L - 1 - r) : 0), ok);
            }
        }
        cp_async_commit();
    };

    issue(0);
    issue(1);
    F4 carry = {{0.f, 0.f, 0.f, 0.f}};
    for (int tile = 0; tile < ntiles; ++tile) {
        cp_async_wait<kScanStages - 2>();
        __syncthreads();
        issue(tile + kScanStages - 1);
        float* st = smem + (tile % kScanStages) * STAGE;
        float* sdy = st + (ck * S) * 32 + cq * 4;
        float* sv = sdy + ARR;
        float* sf = sdy + 2 * ARR;
        float* shp = sdy + 3 * ARR;
        const int r0 = tile * TL + ck * S;
        F4 A = {{1.f, 1.f, 1.f, 1.f}}, H = {{0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int s = 0; s < S; ++s) {
            if (r0 + s < L) {
                F4 dy = ld4(sdy + s * 32), f = ld4(sf + s * 32);
                if (FUSED) {
                    F4 v = ld4(sv + s * 32);
                    float keep = 1.0f - st[4 * ARR + ck * S + s];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v.v[i] = tanhf_fast(v.v[i]);
                        f.v[i] = sigmoidf_fast(f.v[i]) * keep;
                    }
                    st4(sv + s * 32, v);
                    st4(sf + s * 32, f);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    H.v[i] = f.v[i] * (H.v[i] + dy.v[i]);
                    A.v[i] *= f.v[i];
                }
            }
        }
        F4 E = tile_carry_in(A, H, carry, wagg, tile & 1);
#pragma unroll
        for (int s = 0; s < S; ++s) {
            int r = r0 + s;
            if (r < L) {
                F4 dy = ld4(sdy + s * 32), v = ld4(sv + s * 32), f = ld4(sf + s * 32), hp = ld4(shp + s * 32);
                F4 dv, df;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float G = dy.v[i] + E.v[i];
                    dv.v[i] = (1.0f - f.v[i]) * G;
                    df.v[i] = G * (hp.v[i] - v.v[i]);
                    E.v[i] = f.v[i] * G;
                    if (FUSED) {
                        dv.v[i] *= (1.0f - v.v[i] * v.v[i]);       // d tanh
                        df.v[i] *= f.v[i] * (1.0f - f.v[i]);       // d sigmoid; f == 0 on reset steps
                    }
                }
                if (cvalid) {
                    size_t g = (rowbase + (L - 1 - r)) * C + c0 + cq * 4;
                    st4(gdv + g, dv);
                    st4(gdf + g, df);
                }
            }
        }
    }
    cp_async_wait<0>();
}

template <int S, int NARR>
constexpr size_t scan_smem_bytes() {
    return sizeof(float) * (kScanStages * (NARR * kScanChunks * S * 32 + kScanChunks * S) + 2 * 8 * kScanCq * 8);
}

}  // namespace rorl

using namespace rorl;

static int check_blc(int64_t B, int64_t L, int64_t C) {
    if (B <= 0 || L <= 0 || C <= 0 || B > 65535 || L > (1 << 28)) return RORL_ERR_SHAPE;
    if (C % 4 != 0) return RORL_ERR_ALIGN;
    return RORL_OK;
}

template <bool FUSED>
static int launch_gilr_fwd(const float* v, const float* f, const float* start, float* h, int64_t B, int64_t L,
                           int64_t C, cudaStream_t stream) {
    int rc = check_blc(B, L, C);
    if (rc) return rc;
    constexpr int S = 4;
    auto kern = gilr_fwd_kernel<S, FUSED>;
    constexpr size_t smem = scan_smem_bytes<S, 2>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kScanThreads, smem, stream>>>(v, f, start, h, (int)L, (int)C);
    RORL_RETURN_LAUNCH();
}

template <bool FUSED>
static int launch_gilr_bwd(const float* dy, const float* v, const float* f, const float* h, const float* start,
                           float* dv, float* df, int64_t B, int64_t L, int64_t C, cudaStream_t stream) {
    int rc = check_blc(B, L, C);
    if (rc) return rc;
    constexpr int S = 2;
    auto kern = gilr_bwd_kernel<S, FUSED>;
    constexpr size_t smem = scan_smem_bytes<S, 4>();
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((unsigned)((C + 31) / 32), (unsigned)B);
    kern<<<grid, kScanThreads, smem, stream>>>(dy, v, f, h, start, dv, df, (int)L, (int)C);
    RORL_RETURN_LAUNCH();
}

extern "C" {

int rorl_gilr_scan_fwd(const float* v, const float* f, float* h, int64_t B, int64_t L, int64_t C,
                       cudaStream_t stream) {
    if (!v || !f || !h) return RORL_ERR_ARG;
    return launch_gilr_fwd<false>(v, f, nullptr, h, B, L, C, stream);
}

int rorl_gilr_scan_bwd(const float* dh, const float* v, const float* f, const float* h, float* dv, float* df,
                       int64_t B, int64_t L, int64_t C, cudaStream_t stream) {
    if (!dh || !v || !f || !h || !dv || !df) return RORL_ERR_ARG;
    return launch_gilr_bwd<false>(dh, v, f, h, nullptr, dv, df, B, L, C, stream);
}

int rorl_gilr_fused_fwd(const float* u_v, const float* u_f, const float* start, float* h, int64_t B, int64_t L,
                        int64_t C, cudaStream_t stream) {
    if (!u_v || !u_f || !h) return RORL_ERR_ARG;
    return launch_gilr_fwd<true>(u_v, u_f, start, h, B, L, C, stream);
}

int rorl_gilr_fused_bwd(const float* dh, const float* u_v, const float* u_f, const float* h, const float* start,
                        float* du_v, float* du_f, int64_t B, int64_t L, int64_t C, cudaStream_t stream) {
    if (!dh || !u_v || !u_f || !h || !du_v || !du_f) return RORL_ERR_ARG;
    return launch_gilr_bwd<true>(dh, u_v, u_f, h, start, du_v, du_f, B, L, C, stream);
}

}  // extern "C"
