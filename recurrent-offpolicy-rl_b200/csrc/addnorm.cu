// Fused residual-add + LayerNorm / RMSNorm, forward and backward (rows = tokens, C = features).
// Replaces the Triton _layer_norm_fwd_1pass_kernel / _layer_norm_bwd_kernel behind layer_norm_fn /
// rms_norm_fn (ref: offpolicy_rnn/models/smamba/mamba_ssm/ops/triton/layernorm.py:65-120,196-290,
// 464-478; semantics = layernorm_cpu.py:6-35: the residual is added first and returned in fp32).
//
// One warp per row, the row held in registers (C/32 values per lane, float4 accesses), so the
// forward moves 16 B/element (x, residual in; y, residual_out out) and the backward re-reads only
// r and dy.  dw/db are accumulated per warp in registers over a grid-strided set of rows, summed
// across the CTA in shared memory and written as one partial row per CTA (no atomics).
#include "common.cuh"

namespace rorl {

constexpr int kNormWarps = 8;
constexpr int kNormThreads = kNormWarps * 32;

template <int VPL>  // float4 per lane; C = 128 * VPL
__global__ void __launch_bounds__(kNormThreads) addnorm_fwd_kernel(
    const float* __restrict__ x, const float* __restrict__ res, const float* __restrict__ w,
    const float* __restrict__ b, float* __restrict__ y, float* __restrict__ res_out, float* __restrict__ mean_o,
    float* __restrict__ rstd_o, int64_t rows, int C, float eps, int is_rms) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float invC = 1.0f / (float)C;
    float4 wv[VPL], bv[VPL];
    bool ok[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        ok[i] = (i * 32 + lane) * 4 < C;
        wv[i] = ok[i] ? reinterpret_cast<const float4*>(w)[i * 32 + lane] : zero4;
        bv[i] = (b && ok[i]) ? reinterpret_cast<const float4*>(b)[i * 32 + lane] : zero4;
    }
    for (int64_t row = (int64_t)blockIdx.x * kNormWarps + warp; row < rows; row += (int64_t)gridDim.x * kNormWarps) {
        const float4* xp = reinterpret_cast<const float4*>(x + row * C);
        float4 r[VPL];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            r[i] = ok[i] ? xp[i * 32 + lane] : zero4;
            if (res && ok[i]) {
                float4 q = reinterpret_cast<const float4*>(res + row * C)[i * 32 + lane];
                r[i].x += q.x; r[i].y += q.y; r[i].z += q.z; r[i].w += q.w;
            }
            s += (r[i].x + r[i].y) + (r[i].z + r[i].w);
        }
        if (res_out) {
#pragma unroll
            for (int i = 0; i < VPL; ++i)
                if (ok[i]) reinterpret_cast<float4*>(res_out + row * C)[i * 32 + lane] = r[i];
        }
        float mean = 0.f;
        if (!is_rms) mean = warp_sum(s) * invC;
        float v = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float a = r[i].x - mean, bq = r[i].y - mean, c = r[i].z - mean, dd = r[i].w - mean;
            if (ok[i]) v += (a * a + bq * bq) + (c * c + dd * dd);
        }
        v = warp_sum(v) * invC;
        const float rstd = rsqrtf(v + eps);
        if (lane == 0) {
            if (mean_o) mean_o[row] = mean;
            rstd_o[row] = rstd;
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 o;
            o.x = (r[i].x - mean) * rstd * wv[i].x + bv[i].x;
            o.y = (r[i].y - mean) * rstd * wv[i].y + bv[i].y;
            o.z = (r[i].z - mean) * rstd * wv[i].z + bv[i].z;
            o.w = (r[i].w - mean) * rstd * wv[i].w + bv[i].w;
            if (ok[i]) reinterpret_cast<float4*>(y + row * C)[i * 32 + lane] = o;
        }
    }
}

template <int VPL>
__global__ void __launch_bounds__(kNormThreads) addnorm_bwd_kernel(
    const float* __restrict__ dy, const float* __restrict__ dres, const float* __restrict__ r,
    const float* __restrict__ w, const float* __restrict__ mean_i, const float* __restrict__ rstd_i,
    float* __restrict__ dx, float* __restrict__ dw_part, float* __restrict__ db_part, int64_t rows, int C, int is_rms,
    int has_bias) {
    constexpr int CMAX = 128 * VPL;
    __shared__ float s_acc[kNormWarps][CMAX];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float invC = 1.0f / (float)C;
    float4 wv[VPL], dw[VPL], db[VPL];
    bool ok[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        ok[i] = (i * 32 + lane) * 4 < C;
        wv[i] = ok[i] ? reinterpret_cast<const float4*>(w)[i * 32 + lane] : zero4;
        dw[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int64_t row = (int64_t)blockIdx.x * kNormWarps + warp; row < rows; row += (int64_t)gridDim.x * kNormWarps) {
        const float mean = is_rms ? 0.f : mean_i[row];
        const float rstd = rstd_i[row];
        float4 xh[VPL], wdy[VPL];
        float c1 = 0.f, c2 = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 rv = ok[i] ? reinterpret_cast<const float4*>(r + row * C)[i * 32 + lane] : zero4;
            float4 g = ok[i] ? reinterpret_cast<const float4*>(dy + row * C)[i * 32 + lane] : zero4;
            xh[i] = ok[i] ? make_float4((rv.x - mean) * rstd, (rv.y - mean) * rstd, (rv.z - mean) * rstd, (rv.w - mean) * rstd) : zero4;
            wdy[i] = make_float4(g.x * wv[i].x, g.y * wv[i].y, g.z * wv[i].z, g.w * wv[i].w);
            dw[i].x = fmaf(g.x, xh[i].x, dw[i].x); dw[i].y = fmaf(g.y, xh[i].y, dw[i].y);
            dw[i].z = fmaf(g.z, xh[i].z, dw[i].z); dw[i].w = fmaf(g.w, xh[i].w, dw[i].w);
            db[i].x += g.x; db[i].y += g.y; db[i].z += g.z; db[i].w += g.w;
            c1 += (xh[i].x * wdy[i].x + xh[i].y * wdy[i].y) + (xh[i].z * wdy[i].z + xh[i].w * wdy[i].w);
            c2 += (wdy[i].x + wdy[i].y) + (wdy[i].z + wdy[i].w);
        }
        c1 = warp_sum(c1) * invC;
        c2 = is_rms ? 0.f : warp_sum(c2) * invC;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 o;
            o.x = (wdy[i].x - xh[i].x * c1 - c2) * rstd;
            o.y = (wdy[i].y - xh[i].y * c1 - c2) * rstd;
            o.z = (wdy[i].z - xh[i].z * c1 - c2) * rstd;
            o.w = (wdy[i].w - xh[i].w * c1 - c2) * rstd;
            if (dres && ok[i]) {
                float4 q = reinterpret_cast<const float4*>(dres + row * C)[i * 32 + lane];
                o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            if (ok[i]) reinterpret_cast<float4*>(dx + row * C)[i * 32 + lane] = o;
        }
    }
    // CTA-level sum of the per-warp dw / db accumulators, one partial row per CTA
    for (int pass = 0; pass < (has_bias ? 2 : 1); ++pass) {
        __syncthreads();
#pragma unroll
        for (int i = 0; i < VPL; ++i)
            reinterpret_cast<float4*>(&s_acc[warp][0])[i * 32 + lane] = pass == 0 ? dw[i] : db[i];
        __syncthreads();
        float* out = (pass == 0 ? dw_part : db_part) + (size_t)blockIdx.x * C;
        for (int c = threadIdx.x; c < C; c += kNormThreads) {
            float s = 0.f;
#pragma unroll
            for (int wi = 0; wi < kNormWarps; ++wi) s += s_acc[wi][c];
            out[c] = s;
        }
    }
}

}  // namespace rorl

using namespace rorl;

extern "C" {

int rorl_addnorm_nparts(int64_t rows) {
    int64_t need = (rows + kNormWarps - 1) / kNormWarps;
    int64_t cap = 148 * 4;
    return (int)(need < cap ? (need > 0 ? need : 1) : cap);
}

int rorl_addnorm_fwd(const float* x, const float* residual, const float* w, const float* b, float* y,
                     float* residual_out, float* mean, float* rstd, int64_t rows, int64_t C, float eps, int is_rms,
                     cudaStream_t stream) {
    if (!x || !w || !y || !rstd) return RORL_ERR_ARG;
    if (rows <= 0 || C <= 0 || C % 4 || C > 1024) return RORL_ERR_SHAPE;
    int64_t nb = (rows + kNormWarps - 1) / kNormWarps;
    if (nb > 148 * 16) nb = 148 * 16;
    dim3 grid((unsigned)nb);
    const int vpl = C <= 128 ? 1 : (C <= 256 ? 2 : (C <= 512 ? 4 : 8));
    switch (vpl) {
        case 1: addnorm_fwd_kernel<1><<<grid, kNormThreads, 0, stream>>>(x, residual, w, b, y, residual_out, mean, rstd, rows, (int)C, eps, is_rms); break;
        case 2: addnorm_fwd_kernel<2><<<grid, kNormThreads, 0, stream>>>(x, residual, w, b, y, residual_out, mean, rstd, rows, (int)C, eps, is_rms); break;
        case 4: addnorm_fwd_kernel<4><<<grid, kNormThreads, 0, stream>>>(x, residual, w, b, y, residual_out, mean, rstd, rows, (int)C, eps, is_rms); break;
        case 8: addnorm_fwd_kernel<8><<<grid, kNormThreads, 0, stream>>>(x, residual, w, b, y, residual_out, mean, rstd, rows, (int)C, eps, is_rms); break;
        default: return RORL_ERR_SHAPE;
    }
    RORL_RETURN_LAUNCH();
}

int rorl_addnorm_bwd(const float* dy, const float* dres_out, const float* r, const float* w, const float* mean,
                     const float* rstd, float* dx, float* dw_part, float* db_part, int64_t rows, int64_t C,
                     int is_rms, int has_bias, cudaStream_t stream) {
    if (!dy || !r || !w || !rstd || !dx || !dw_part) return RORL_ERR_ARG;
    if (!is_rms && !mean) return RORL_ERR_ARG;
    if (has_bias && !db_part) return RORL_ERR_ARG;
    if (rows <= 0 || C <= 0 || C % 4 || C > 1024) return RORL_ERR_SHAPE;
    dim3 grid((unsigned)rorl_addnorm_nparts(rows));
    const int vpl = C <= 128 ? 1 : (C <= 256 ? 2 : (C <= 512 ? 4 : 8));
    switch (vpl) {
        case 1: addnorm_bwd_kernel<1><<<grid, kNormThreads, 0, stream>>>(dy, dres_out, r, w, mean, rstd, dx, dw_part, db_part, rows, (int)C, is_rms, has_bias); break;
        case 2: addnorm_bwd_kernel<2><<<grid, kNormThreads, 0, stream>>>(dy, dres_out, r, w, mean, rstd, dx, dw_part, db_part, rows, (int)C, is_rms, has_bias); break;
        case 4: addnorm_bwd_kernel<4><<<grid, kNormThreads, 0, stream>>>(dy, dres_out, r, w, mean, rstd, dx, dw_part, db_part, rows, (int)C, is_rms, has_bias); break;
        case 8: addnorm_bwd_kernel<8><<<grid, kNormThreads, 0, stream>>>(dy, dres_out, r, w, mean, rstd, dx, dw_part, db_part, rows, (int)C, is_rms, has_bias); break;
        default: return RORL_ERR_SHAPE;
    }
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
