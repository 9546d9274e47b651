// Mamba selective scan with reset flag (smamba / s6 encoders), forward and backward, sm_100a.
//
//   delta = softplus(delta_raw + delta_bias)
//   h_t[n] = (1 - start_t) * exp(delta_t * A[d,n]) * h_{t-1}[n] + delta_t * u_t * B_t[n],  h_{-1} = 0
//   y_t    = (sum_n h_t[n] * C_t[n] + D[d] * u_t) * silu(z_t)
//
// Semantics pinned by the reference's selective_scan_ref
// (ref: offpolicy_rnn/models/smamba/mamba_ssm/ops/selective_scan_interface_new.py:96-166, reset at
// :133-135); replaces the binary-only selective_scan_cuda.fwd/.bwd it calls at :47 and :72-75.
//
// Layout is token-major: u/delta/z/y are [B, L, D] with a row stride (so they may be column slices
// of a wider projection output), B/C are [B, L, N].  That is what the projections on either side of
// the scan produce/consume as GEMM operands, and it is also the s6 layer's native layout
// (ref: offpolicy_rnn/models/s6/selective_scan/triton_scan.py:19-72).
//
// Work decomposition (N = d_state, S = 8 states per lane, LPD = N/8 lanes per channel):
//   thread = (channel d, 8 of the N states); a warp holds 32/LPD channels; a CTA of 8 warps holds
//   DT = 8*32/LPD channels of one batch row and walks the sequence serially in time, with h in
//   registers.  The sum over n for y is an LPD-lane shuffle; there is no parallel-scan combine cost,
//   no recomputation in the forward and exactly one exp per (t, d, n).  The kernel is bound by the
//   MUFU (ex2) and FMA pipes, not by HBM; operands are staged per tile of steps
//   in shared memory with cp.async double buffering, so HBM traffic is the algorithmic minimum.
//
// Backward: the forward (when asked) checkpoints h every 16 steps.  The backward walks 16-step
// chunks in reverse: recompute h inside the chunk from the checkpoint into registers, then run the
// adjoint recurrence  lambda_t = g_t C_t + a_{t+1} lambda_{t+1}.  dB/dC need a sum over channels:
// a shuffle reduce-scatter across the warp's channels, a shared-memory sum across the 8 warps and
// one partial tile per CTA in global memory, summed by the caller.  No atomics anywhere, so the
// result is deterministic (the reference's CUDA backward is not, ref: results.md:4).
#include "common.cuh"

namespace rorl {

constexpr int kSelThreads = 256;
constexpr int kCkptEvery = 8;

template <int N>
struct SelCfg {
    static_assert(N == 16 || N == 32 || N == 64, "d_state must be 16, 32 or 64");
    static constexpr int S = 8;
    static constexpr int LPD = N / S;         // lanes per channel
    static constexpr int DPW = 32 / LPD;      // channels per warp
    static constexpr int DT = 8 * DPW;        // channels per CTA
    static constexpr int QPR = DT / 4;        // float4 quads per tile row
};

struct SelFwdParams {
    const float *u, *delta, *z, *Bm, *Cm, *A, *Dskip, *dbias, *start, *h0;
    float *y, *ckpt, *last_state;
    int L, D;
    int ld_u, ld_delta, ld_z, ld_B, ld_C, ld_y;
    int nckpt;
    int a_log;                  // A is handed over as A_log (A = -exp(A_log))
    int dbg;
};

// Forward decomposition (warp specialised, shared-memory-bandwidth aware).
//
// The scan is serial in time per (d, n); the only parallelism is B x D x N and the arithmetic floor is the MUFU
// pipe: one ex2 per (t, d, n), 16 per clock per SM (measured with tools/microbench/pipes.cu: 15.9 / clk / SM;
// fma.f32x2 = 117 fma / clk / SM; a scan-like mix sustains 15.6 ex2 / clk / SM from 16 warps).  What the ncu
// captures of the earlier versions showed, in order:
//   * every warp running load -> transform -> barrier -> scan -> barrier -> write-back phases cost 176 us for the
//     data-movement skeleton alone, not overlapped with the scan (tools/sel_phases.py)  => warp specialisation;
//   * with one channel x 8 states per thread the time loop reads B_t and C_t (64 B per thread per step) through
//     LDS.128, which is 4 shared-memory wavefronts per instruction however much of it is a broadcast: 26 wavefronts
//     per warp-step, l1tex shared pipe 83 % busy, 439 cycles per step against 256 for the MUFU pipe
//     => two channels per thread share one copy of B_t / C_t (22 wavefronts per 16 states instead of 26 per 8).
// Roles per CTA (DT channels of one batch row):
//   * 8 SCAN warps, thread = 2 adjacent channels x 4 states (state, A, B, C as float2 pairs): per step LDS.64 of
//     (delta or +inf) and delta*u for the channel pair, 2 LDS.128 of B/C, 8 ex2 issued one step ahead of the state
//     update they feed and interleaved with it, packed fma/mul.f32x2, one STS.64 of the lane's partial sums of h.C.
//     No shuffles, branches or block barriers in the time loop; tiles are handed over through mbarriers.
//   * 4 HELPER warps: cp.async of the raw tile three tiles ahead (4-stage ring), the per-(t, d) transform
//     (softplus, delta*u, reset folded into the exponent argument as +inf so that ex2(-inf) = 0, silu(z) and
//     the D*u skip pre-multiplied by it), and the write-back of y = partials * silu(z) + skip as coalesced rows.
template <int N>
struct SelFwdCfg {
#ifndef RORL_SEL_S
#define RORL_SEL_S 4
#endif
    static constexpr int S = RORL_SEL_S;                  // states per scan thread (per channel)
    static constexpr int CH = 2;                          // channels per scan thread
    static constexpr int LPD = N / S;                     // lanes per channel pair
    static constexpr int PPW = 32 / LPD;                  // channel pairs per warp
    static constexpr int NSCAN = 32 * 32 / RORL_SEL_S;    // scan threads (8 warps at S = 4)
    static constexpr int DT = (NSCAN / 32) * PPW * CH;    // channels per CTA
    static constexpr int QPR = DT / 4;                    // float4 quads per tile row
    static constexpr int TC = 2 * RORL_SEL_S;             // steps per tile
    static constexpr int NST = 4;                         // raw stages in flight
    static constexpr int STAGE = TC * (4 * DT + 2 * N) + TC * QPR;   // u->skip, delta->dtA, z->silu(z), du, B, C, reset flag per quad-row
    static constexpr int PROW = DT * LPD + DT;            // partial-sum row: [DT/4 quads][2 pairs][LPD][2] + 4 floats of skew per quad
    static constexpr int PART = TC * PROW;                // (the skew makes the helpers' LDS.128 of a quad bank-conflict free)
    static constexpr int NHELP = 128;                     // helper threads
    static constexpr int NTHREADS = NSCAN + NHELP;
    static constexpr size_t SMEM = sizeof(float) * (NST * STAGE + 2 * PART) + 128;
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
// the state matrix as the caller holds it: A itself, or the parameter A_log with A = -exp(A_log) (ref: smamba/mamba.py:215)
__device__ __forceinline__ float sel_A(float a, int a_log) { return a_log ? -expf(a) : a; }

template <int N, bool HAS_Z, bool SOFTPLUS>
__global__ void __launch_bounds__(SelFwdCfg<N>::NTHREADS, 2) selscan_fwd_kernel(const SelFwdParams p) {
    using Cfg = SelFwdCfg<N>;
    constexpr int S = Cfg::S, LPD = Cfg::LPD, PPW = Cfg::PPW, DT = Cfg::DT, QPR = Cfg::QPR, TC = Cfg::TC, STAGE = Cfg::STAGE;
    constexpr int NST = Cfg::NST, PART = Cfg::PART, PROW = Cfg::PROW, NHELP = Cfg::NHELP, NSCAN = Cfg::NSCAN, H2 = S / 2;
    extern __shared__ __align__(16) float smem[];
    float* s_part = smem + NST * STAGE;                     // [2][TC][DT/2][LPD][2]
    const uint32_t bars = smem_u32(s_part + 2 * PART);      // full[NST], done[2], freep[2]
    auto bar_full = [&](int st) { return bars + 8u * st; };
    auto bar_done = [&](int pb) { return bars + 8u * (NST + pb); };
    auto bar_freep = [&](int pb) { return bars + 8u * (NST + 2 + pb); };

    const int tid = threadIdx.x;
    const int b = blockIdx.y, d0 = blockIdx.x * DT;
    const int L = p.L;
    const size_t row0 = (size_t)b * L;
    const int ntiles = (L + TC - 1) / TC;
    if (tid == 0) {
        for (int st = 0; st < NST; ++st) mbar_init(bar_full(st), NHELP / 32);
        for (int pb = 0; pb < 2; ++pb) {
            mbar_init(bar_done(pb), NSCAN / 32);
            mbar_init(bar_freep(pb), NHELP / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NSCAN) {
        // ------------------------------------------------------------------------------------ helper warps
        const int ht = tid - NSCAN;
        auto issue = [&](int tile) {
            if (tile < ntiles) {
                float* st = smem + (tile % NST) * STAGE;
                for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                    const int r = idx / QPR, q = idx % QPR, t = tile * TC + r, col = d0 + q * 4;
                    const bool ok = col < p.D && t < L;
                    const size_t row = row0 + (ok ? t : 0);
                    cp_async16(st + r * DT + q * 4, p.u + row * p.ld_u + (ok ? col : 0), ok);
                    cp_async16(st + TC * DT + r * DT + q * 4, p.delta + row * p.ld_delta + (ok ? col : 0), ok);
                    if (HAS_Z) cp_async16(st + 2 * TC * DT + r * DT + q * 4, p.z + row * p.ld_z + (ok ? col : 0), ok);
                    // the row's reset flag travels with the tile, one private copy per quad-row so that the thread that
                    // transforms it is the thread whose wait_group covers it (ncu: the helpers spent 40 % of their time on
                    // the L2 latency of reading it in the transform, the scan warps a quarter of theirs waiting for them)
                    const bool okf = p.start != nullptr && t < L;
                    cp_async4(st + TC * (4 * DT + 2 * N) + idx, p.start + row0 + (okf ? t : 0), okf);
                }
                for (int idx = ht; idx < TC * N / 4; idx += NHELP) {
                    const int r = idx / (N / 4), q = idx % (N / 4), t = tile * TC + r;
                    const bool ok = t < L;
                    const size_t row = row0 + (ok ? t : 0);
                    cp_async16(st + 4 * TC * DT + r * N + q * 4, p.Bm + row * p.ld_B + q * 4, ok);
                    cp_async16(st + 4 * TC * DT + TC * N + r * N + q * 4, p.Cm + row * p.ld_C + q * 4, ok);
                }
            }
            cp_async_commit();
        };
        // loop-invariant per thread: its quad column (NHELP % QPR == 0), hence its slice of delta_bias and D
        const int hq = ht % QPR, hcol = d0 + hq * 4;
        float4 hbias = make_float4(0.f, 0.f, 0.f, 0.f), hD4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dbias && hcol < p.D) hbias = __ldg(reinterpret_cast<const float4*>(p.dbias + hcol));
        if (p.Dskip && hcol < p.D) hD4 = __ldg(reinterpret_cast<const float4*>(p.Dskip + hcol));
        // per-(t, d) transform of the elements this thread staged itself (visible to it after wait_group)
        auto transform = [&](int tile) {
            float* st = smem + (tile % NST) * STAGE;
            for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                const int r = idx / QPR, q = idx % QPR, t = tile * TC + r;
                const float rs = st[TC * (4 * DT + 2 * N) + idx];
                float4* pu = reinterpret_cast<float4*>(st + r * DT + q * 4);
                float4* pd = reinterpret_cast<float4*>(st + TC * DT + r * DT + q * 4);
                float4* pz = reinterpret_cast<float4*>(st + 2 * TC * DT + r * DT + q * 4);
                float4 x = *pd;
                const float4 uu = *pu;
                x.x += hbias.x; x.y += hbias.y; x.z += hbias.z; x.w += hbias.w;
                if (SOFTPLUS) {
                    x.x = softplusf_fast(x.x); x.y = softplusf_fast(x.y);
                    x.z = softplusf_fast(x.z); x.w = softplusf_fast(x.w);
                }
                // rows past the end of the sequence: delta = 0 (a = 1, du = 0) leaves the state untouched
                if (t >= L) x = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(st + 3 * TC * DT + r * DT + q * 4) =
                    make_float4(x.x * uu.x, x.y * uu.y, x.z * uu.z, x.w * uu.w);
                float4 gz = make_float4(1.f, 1.f, 1.f, 1.f);
                if (HAS_Z) {
                    const float4 zz = *pz;
                    gz = make_float4(zz.x * sigmoidf_fast(zz.x), zz.y * sigmoidf_fast(zz.y), zz.z * sigmoidf_fast(zz.z),
                                     zz.w * sigmoidf_fast(zz.w));
                }
                const float4 D4 = hD4;
                *pz = gz;
                *pu = make_float4(D4.x * uu.x * gz.x, D4.y * uu.y * gz.y, D4.z * uu.z * gz.z, D4.w * uu.w * gz.w);
                // reset: the scan multiplies this slot by A * log2(e) < 0, so +inf gives a = ex2(-inf) = 0
                if (rs != 0.f) x = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
                *pd = x;
            }
        };
        for (int k = 0; k < NST - 1; ++k) issue(k);
        cp_async_wait<NST - 2>();
        transform(0);
        __syncwarp();
        if ((ht & 31) == 0) mbar_arrive(bar_full(0));
        for (int k = 0; k < ntiles; ++k) {
            // stage (k + NST - 1) % NST held tile k - 1, whose write-back finished behind the helper barrier below
            issue(k + NST - 1);
            if (k + 1 < ntiles) {
                cp_async_wait<NST - 2>();
                transform(k + 1);
                __syncwarp();
                if ((ht & 31) == 0) mbar_arrive(bar_full((k + 1) % NST));
            }
            // write-back of tile k once the scan warps have left its partial sums
            mbar_wait(bar_done(k & 1), (k >> 1) & 1);
            const float* st = smem + (k % NST) * STAGE;
            const float* part = s_part + (k & 1) * PART;
            for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                const int r = idx / QPR, q = idx % QPR, t = k * TC + r, col = d0 + q * 4;
                if (col < p.D && t < L) {
                    const float4 skip = *reinterpret_cast<const float4*>(st + r * DT + q * 4);
                    const float4 gz = *reinterpret_cast<const float4*>(st + 2 * TC * DT + r * DT + q * 4);
                    float yv[4];
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp) {         // the quad's two channel pairs
                        const float2* pp = reinterpret_cast<const float2*>(part + r * PROW + q * (4 * LPD + 4) + cp * 2 * LPD);
                        float2 sum = f2(0.f, 0.f);
#pragma unroll
                        for (int kk = 0; kk < LPD; ++kk) {
                            const float2 v = pp[kk];
                            sum.x += v.x; sum.y += v.y;
                        }
                        yv[2 * cp] = sum.x; yv[2 * cp + 1] = sum.y;
                    }
                    *reinterpret_cast<float4*>(p.y + (row0 + t) * p.ld_y + col) =
                        make_float4(fmaf(yv[0], gz.x, skip.x), fmaf(yv[1], gz.y, skip.y), fmaf(yv[2], gz.z, skip.z), fmaf(yv[3], gz.w, skip.w));
                }
            }
            __syncwarp();
            if ((ht & 31) == 0) mbar_arrive(bar_freep(k & 1));
            asm volatile("bar.sync 1, %0;" ::"n"(NHELP) : "memory");   // all helpers are done reading stage k % NST
        }
    } else {
        // ------------------------------------------------------------------------------------ scan warps
        const int lane = tid & 31, warp = tid >> 5;
        const int pl = lane / LPD, ng = lane % LPD;
        const int ploc = warp * PPW + pl;                   // channel pair within the CTA
        const int dloc = ploc * 2;
        const int d = d0 + dloc;
        float2 A2[2][H2], h[2][H2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const bool ok = d + c < p.D;
#pragma unroll
            for (int j = 0; j < H2; ++j) {
                A2[c][j] = ok ? f2(sel_A(p.A[(size_t)(d + c) * N + ng * S + 2 * j], p.a_log) * kLog2e,
                                   sel_A(p.A[(size_t)(d + c) * N + ng * S + 2 * j + 1], p.a_log) * kLog2e)
                              : f2(0.f, 0.f);
                h[c][j] = f2(0.f, 0.f);
                if (p.h0 != nullptr && ok) {                // carried state entering the call: [B, D, N]
                    const float* hp = p.h0 + ((size_t)b * p.D + d + c) * N + ng * S + 2 * j;
                    h[c][j] = f2(__ldg(hp), __ldg(hp + 1));
                }
            }
        }
        for (int k = 0; k < ntiles; ++k) {
            const float* st = smem + (k % NST) * STAGE;
            const float* s_dt = st + TC * DT;
            const float* s_du = st + 3 * TC * DT;
            const float* s_B = st + 4 * TC * DT;
            const float* s_C = s_B + TC * N;
            float* part = s_part + (k & 1) * PART;
            mbar_wait(bar_full(k % NST), (k / NST) & 1);
            if (k >= 2) mbar_wait(bar_freep(k & 1), ((k >> 1) - 1) & 1);
            // software pipeline: a[][] holds exp(delta_i A) for the step about to be applied
            float2 a[2][H2];
            {
                const float2 dt = *reinterpret_cast<const float2*>(s_dt + dloc);
#pragma unroll
                for (int j = 0; j < H2; ++j) {
                    const float2 e0 = __fmul2_rn(f2(dt.x, dt.x), A2[0][j]);
                    const float2 e1 = __fmul2_rn(f2(dt.y, dt.y), A2[1][j]);
                    a[0][j] = f2(ex2f(e0.x), ex2f(e0.y));
                    a[1][j] = f2(ex2f(e1.x), ex2f(e1.y));
                }
            }
#pragma unroll
            for (int i = 0; i < TC; ++i) {
                const float2 du = *reinterpret_cast<const float2*>(s_du + i * DT + dloc);
                const float2 dtn = (i + 1 < TC) ? *reinterpret_cast<const float2*>(s_dt + (i + 1) * DT + dloc) : f2(0.f, 0.f);
                float2 Bv[H2], Cv[H2];
#pragma unroll
                for (int j = 0; j < H2; j += 2) {
                    const float4 bq = *reinterpret_cast<const float4*>(s_B + i * N + ng * S + 2 * j);
                    const float4 cq = *reinterpret_cast<const float4*>(s_C + i * N + ng * S + 2 * j);
                    Bv[j] = f2(bq.x, bq.y); Bv[j + 1] = f2(bq.z, bq.w);
                    Cv[j] = f2(cq.x, cq.y); Cv[j + 1] = f2(cq.z, cq.w);
                }
                float2 acc[2] = {f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
                for (int j = 0; j < H2; ++j) {
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float duc = c ? du.y : du.x;
                        h[c][j] = __ffma2_rn(a[c][j], h[c][j], __fmul2_rn(f2(duc, duc), Bv[j]));
                        acc[c] = __ffma2_rn(h[c][j], Cv[j], acc[c]);
                        if (i + 1 < TC) {          // next step's decay, issued between this step's FMAs
                            const float dtc = c ? dtn.y : dtn.x;
                            const float2 e = __fmul2_rn(f2(dtc, dtc), A2[c][j]);
                            a[c][j] = f2(ex2f(e.x), ex2f(e.y));
                        }
                    }
                }
                *reinterpret_cast<float2*>(part + i * PROW + (ploc >> 1) * (4 * LPD + 4) + (ploc & 1) * 2 * LPD + 2 * ng) =
                    f2(acc[0].x + acc[0].y, acc[1].x + acc[1].y);
                if (((i + 1) % (TC < kCkptEvery ? TC : kCkptEvery)) == 0 && ((k * TC + i) % kCkptEvery) == kCkptEvery - 1) {
                    const int ck = (k * TC + i) / kCkptEvery;
                    if (p.ckpt != nullptr && ck < p.nckpt) {
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            if (d + c < p.D) {
                                float* cpt = p.ckpt + (((size_t)b * p.nckpt + ck) * p.D + d + c) * N + ng * S;
#pragma unroll
                                for (int j = 0; j < H2; j += 2)
                                    *reinterpret_cast<float4*>(cpt + 2 * j) = make_float4(h[c][j].x, h[c][j].y, h[c][j + 1].x, h[c][j + 1].y);
                            }
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_done(k & 1));
        }
        if (p.last_state != nullptr) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                if (d + c < p.D) {
                    float* cpt = p.last_state + ((size_t)b * p.D + d + c) * N + ng * S;
#pragma unroll
                    for (int j = 0; j < H2; j += 2)
                        *reinterpret_cast<float4*>(cpt + 2 * j) = make_float4(h[c][j].x, h[c][j].y, h[c][j + 1].x, h[c][j + 1].y);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
struct SelBwdParams {
    const float *u, *delta, *z, *Bm, *Cm, *A, *Dskip, *dbias, *start, *dy, *ckpt, *h0;
    float *du, *ddelta, *dz, *dBC_part, *dA_part, *dD_part, *dbias_part;
    int L, D, Bsz;
    int ld_u, ld_delta, ld_z, ld_B, ld_C, ld_dy, ld_du, ld_ddelta, ld_dz;
    int nckpt;
    int a_log;                  // A is handed over as A_log; dA_part then receives d A_log partials
};

// Backward decomposition (warp specialised like the forward).
//
// Per element (t, d, n) the adjoint needs h_t (recomputed from a checkpoint: 1 ex2), a_t again for
// lambda_t = g_t C_t + a_{t+1} lambda_{t+1} (1 ex2) and ~14 fp32 operations, plus two reductions: over n for
// d(delta), du, dz (per-lane partial sums left in shared memory, finished by the helper warps) and over d for dB, dC
// (shuffle reduce-scatter across the warp's channels, shared-memory sum across warps, one partial tile per CTA in
// global memory, summed by the caller: deterministic, no atomics).
// The forward checkpoints h every kCkptEvery = 8 steps; a chunk of 8 steps is walked forward (h_t kept in
// registers: 32 per thread and channel) and then backward.  The register file, not the pipes, limits residency
// here: 12 registers of state per (d, n) element, so a CTA holds 64 channels with 8 MAIN warps (thread = 2 channels
// x 4 states, 168 registers) + 4 HELPER warps, one CTA per SM.  (The first version had 16 main warps of one channel
// per thread: ncu showed the LSU shared-memory pipe at 70 %, most of it B_t / C_t LDS.128 broadcasts; with two
// channels per thread sharing them it is 50 % and the kernel went 767 -> 719 us.  What remains is instruction
// issue: ~26 thread instructions per (t, d, n) element at 0.57 IPC per scheduler with 3 warps per scheduler.)  The helpers run the same cp.async / transform
// / write-back pipeline as in the forward: softplus and its derivative, delta*u, g = dy * silu(z), the reset as
// delta = +inf, and afterwards d(delta) = (sum_n t1 A ln2... see below) * softplus', du, dz, the dD / dbias
// accumulators and the cross-warp dB / dC sums.
template <int N>
struct SelBwdCfg {
    static constexpr int S = 4;                           // states per main thread
    static constexpr int LPD = N / S;                     // lanes per channel
    static constexpr int DPW = 32 / LPD;                  // channels per warp
#ifndef RORL_SELBWD_CH
#define RORL_SELBWD_CH 2
#endif
    static constexpr int CH = RORL_SELBWD_CH;             // channels per main thread (they share the B_t / C_t loads)
    static constexpr int NMAINW = 16 / CH;                // main warps
    static constexpr int NMAIN = NMAINW * 32;
    static constexpr int DT = NMAINW * DPW * CH;          // channels per CTA
    static constexpr int QPR = DT / 4;                    // float4 quads per tile row
    static constexpr int TC = kCkptEvery;                 // steps per chunk
    static constexpr int NST = 3;                         // stages in flight
    static constexpr int NARR = 7;                        // u | dtA | du | g | dt | softplus' | dz coefficient
    static constexpr int STAGE = TC * (NARR * DT + 2 * N) + TC * QPR;   // + one reset flag per quad-row (see the forward)
    static constexpr int QS = 4 * LPD + 4;                // partial-sum quad stride (4 channels x LPD lanes + skew)
    static constexpr int PROW = QPR * QS;                 // partial-sum row
    static constexpr int PLANE = TC * PROW;               // one plane (sB | sA | y) of per-lane partial sums
    static constexpr int RED = TC * NMAINW * DPW * 2 * N; // per (step, warp, channel-lane): this lane's dB | dC, summed by the helpers
    static constexpr int NHELP = 128;
    static constexpr int NTHREADS = NMAIN + NHELP;
    static constexpr size_t SMEM = sizeof(float) * (NST * STAGE + 3 * PLANE + RED + 2 * DT * (NHELP / QPR)) + 128;
};

// Reduce-scatter 4 per-lane values across the DPW channel-lanes of a warp (lane bits above the LPD state-group
// bits).  Returns the number of values this lane keeps (sums over all channels of the warp); `first` is the state
// index (within the lane's group of 4) of v[0]; `writer` is false for the duplicate lanes when DPW > 4.
template <int LPD>
__device__ __forceinline__ int channel_reduce_scatter4(float (&v)[4], int dl, int& first, bool& writer) {
    constexpr int DPW = 32 / LPD;
    writer = true;
    if (DPW == 1) { first = 0; return 4; }
    if (DPW == 2) {
        const bool up = (dl & 1) != 0;
        const float s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
        const float k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];
        v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, LPD);
        v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, LPD);
        first = up ? 2 : 0;
        return 2;
    }
    // DPW >= 4: the two highest channel bits select the state, lower bits (DPW = 8) are plain butterflies
    constexpr int HB = DPW / 2, LB = DPW / 4;            // lane-bit values (in units of channels) of the two rounds
    {
        const bool up = (dl & HB) != 0;
        const float s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
        const float k0 = up ? v[2] : v[0], k1 = up ? v[3] : v[1];
        v[0] = k0 + __shfl_xor_sync(0xffffffffu, s0, HB * LPD);
        v[1] = k1 + __shfl_xor_sync(0xffffffffu, s1, HB * LPD);
    }
    {
        const bool up = (dl & LB) != 0;
        const float sd = up ? v[0] : v[1], kp = up ? v[1] : v[0];
        v[0] = kp + __shfl_xor_sync(0xffffffffu, sd, LB * LPD);
    }
    first = ((dl & HB) ? 2 : 0) + ((dl & LB) ? 1 : 0);
#pragma unroll
    for (int m = LB / 2; m >= 1; m >>= 1) {
        v[0] += __shfl_xor_sync(0xffffffffu, v[0], m * LPD);
        writer = writer && ((dl & m) == 0);
    }
    return 1;
}

template <int N, bool HAS_Z, bool SOFTPLUS>
__global__ void __launch_bounds__(SelBwdCfg<N>::NTHREADS, 1) selscan_bwd_kernel(const SelBwdParams p) {
    using Cfg = SelBwdCfg<N>;
    constexpr int S = Cfg::S, LPD = Cfg::LPD, DPW = Cfg::DPW, DT = Cfg::DT, QPR = Cfg::QPR, TC = Cfg::TC, STAGE = Cfg::STAGE;
    constexpr int NST = Cfg::NST, QS = Cfg::QS, PROW = Cfg::PROW, PLANE = Cfg::PLANE, RED = Cfg::RED, NHELP = Cfg::NHELP;
    constexpr int NMAIN = Cfg::NMAIN, NMAINW = Cfg::NMAINW;
    extern __shared__ __align__(16) float smem[];
    float* s_pl = smem + NST * STAGE;                       // [3][TC][PROW]: sB, sA, y partials per lane
    float* s_red = s_pl + 3 * PLANE;                        // [NMAINW][TC][2N]
    float* s_acc = s_red + RED;                             // [2][NHELP / QPR][DT]: dD / dbias partials of the helpers
    const uint32_t bars = smem_u32(s_acc + 2 * DT * (NHELP / QPR));
    auto bar_full = [&](int st) { return bars + 8u * st; };
    const uint32_t bar_done = bars + 8u * NST, bar_freep = bars + 8u * (NST + 1);

    const int tid = threadIdx.x;
    const int b = blockIdx.y, d0 = blockIdx.x * DT;
    const int L = p.L;
    const size_t row0 = (size_t)b * L;
    const int nch = (L + TC - 1) / TC;                      // chunks; processing order c = 0 .. nch-1 is chunk nch-1-c
    if (tid == 0) {
        for (int st = 0; st < NST; ++st) mbar_init(bar_full(st), NHELP / 32);
        mbar_init(bar_done, NMAINW);
        mbar_init(bar_freep, NHELP / 32);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= NMAIN) {
        // ------------------------------------------------------------------------------------ helper warps
        // register re-allocation between the warpgroups (the register file, not the pipes, limits residency here): the
        // helper warpgroup gives up what it does not need, the two main warpgroups take it (8 x 32 x 208 + 4 x 32 x 80
        // <= 64 K registers), which keeps the main loop's ~190 live values out of local memory
        if (NMAIN == 256) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        const int ht = tid - NMAIN;
        const int myq = ht % QPR;                           // this thread's quad column is fixed (NHELP % QPR == 0)
        const int mycol = d0 + myq * 4;
        const bool colok = mycol < p.D;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f), D4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.dbias && colok) bias4 = __ldg(reinterpret_cast<const float4*>(p.dbias + mycol));
        if (p.Dskip && colok) D4 = __ldg(reinterpret_cast<const float4*>(p.Dskip + mycol));
        float4 accD = make_float4(0.f, 0.f, 0.f, 0.f), accB = make_float4(0.f, 0.f, 0.f, 0.f);
        auto issue = [&](int c) {
            if (c < nch) {
                const int k = nch - 1 - c;
                float* st = smem + (c % NST) * STAGE;
                for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                    const int r = idx / QPR, t = k * TC + r;
                    const bool ok = colok && t < L;
                    const size_t row = row0 + (ok ? t : 0);
                    const int col = ok ? mycol : 0;
                    cp_async16(st + 0 * TC * DT + r * DT + myq * 4, p.u + row * p.ld_u + col, ok);
                    cp_async16(st + 4 * TC * DT + r * DT + myq * 4, p.delta + row * p.ld_delta + col, ok);
                    cp_async16(st + 3 * TC * DT + r * DT + myq * 4, p.dy + row * p.ld_dy + col, ok);
                    if (HAS_Z) cp_async16(st + 6 * TC * DT + r * DT + myq * 4, p.z + row * p.ld_z + col, ok);
                    const bool okf = p.start != nullptr && t < L;
                    cp_async4(st + TC * (Cfg::NARR * DT + 2 * N) + idx, p.start + row0 + (okf ? t : 0), okf);
                }
                for (int idx = ht; idx < TC * N / 4; idx += NHELP) {
                    const int r = idx / (N / 4), q = idx % (N / 4), t = k * TC + r;
                    const bool ok = t < L;
                    const size_t row = row0 + (ok ? t : 0);
                    cp_async16(st + Cfg::NARR * TC * DT + r * N + q * 4, p.Bm + row * p.ld_B + q * 4, ok);
                    cp_async16(st + Cfg::NARR * TC * DT + TC * N + r * N + q * 4, p.Cm + row * p.ld_C + q * 4, ok);
                }
            }
            cp_async_commit();
        };
        auto transform = [&](int c) {
            const int k = nch - 1 - c;
            float* st = smem + (c % NST) * STAGE;
            for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                const int r = idx / QPR, t = k * TC + r, o = r * DT + myq * 4;
                const float rs = st[TC * (Cfg::NARR * DT + 2 * N) + idx];
                const float4 uu = *reinterpret_cast<const float4*>(st + o);
                float4 x = *reinterpret_cast<const float4*>(st + 4 * TC * DT + o);
                const float4 dyv = *reinterpret_cast<const float4*>(st + 3 * TC * DT + o);
                x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
                float4 sg = make_float4(1.f, 1.f, 1.f, 1.f);
                if (SOFTPLUS) {
                    sg = make_float4(sigmoidf_fast(x.x), sigmoidf_fast(x.y), sigmoidf_fast(x.z), sigmoidf_fast(x.w));
                    x.x = softplusf_fast(x.x); x.y = softplusf_fast(x.y); x.z = softplusf_fast(x.z); x.w = softplusf_fast(x.w);
                }
                if (t >= L) x = make_float4(0.f, 0.f, 0.f, 0.f);      // a = 1, du = 0 and (zero-filled dy) g = 0: a no-op step
                *reinterpret_cast<float4*>(st + 4 * TC * DT + o) = x;                                     // delta
                *reinterpret_cast<float4*>(st + 5 * TC * DT + o) = sg;                                    // d softplus
                *reinterpret_cast<float4*>(st + 2 * TC * DT + o) = make_float4(x.x * uu.x, x.y * uu.y, x.z * uu.z, x.w * uu.w);
                const float4 xa = rs != 0.f ? make_float4(INFINITY, INFINITY, INFINITY, INFINITY) : x;
                *reinterpret_cast<float4*>(st + 1 * TC * DT + o) = xa;                                    // delta or +inf
                float4 g = dyv;
                if (HAS_Z) {
                    const float4 zz = *reinterpret_cast<const float4*>(st + 6 * TC * DT + o);
                    const float4 sz = make_float4(sigmoidf_fast(zz.x), sigmoidf_fast(zz.y), sigmoidf_fast(zz.z), sigmoidf_fast(zz.w));
                    g = make_float4(dyv.x * zz.x * sz.x, dyv.y * zz.y * sz.y, dyv.z * zz.z * sz.z, dyv.w * zz.w * sz.w);
                    *reinterpret_cast<float4*>(st + 6 * TC * DT + o) =
                        make_float4(dyv.x * sz.x * (1.f + zz.x * (1.f - sz.x)), dyv.y * sz.y * (1.f + zz.y * (1.f - sz.y)),
                                    dyv.z * sz.z * (1.f + zz.z * (1.f - sz.z)), dyv.w * sz.w * (1.f + zz.w * (1.f - sz.w)));
                }
                *reinterpret_cast<float4*>(st + 3 * TC * DT + o) = g;
            }
        };
        for (int c = 0; c < NST - 1; ++c) issue(c);
        cp_async_wait<NST - 2>();
        transform(0);
        __syncwarp();
        if ((ht & 31) == 0) mbar_arrive(bar_full(0));
        for (int c = 0; c < nch; ++c) {
            const int k = nch - 1 - c;
            issue(c + NST - 1);
            if (c + 1 < nch) {
                cp_async_wait<NST - 2>();
                transform(c + 1);
                __syncwarp();
                if ((ht & 31) == 0) mbar_arrive(bar_full((c + 1) % NST));
            }
            mbar_wait(bar_done, c & 1);
            const float* st = smem + (c % NST) * STAGE;
            for (int idx = ht; idx < TC * QPR; idx += NHELP) {
                const int r = idx / QPR, t = k * TC + r, o = r * DT + myq * 4;
                if (colok && t < L) {
                    float sB[4], sA[4], yp[4];
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        const float* pb = s_pl + r * PROW + myq * QS + ch * LPD;
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
#pragma unroll
                        for (int kk = 0; kk < LPD; kk += 4) {
                            const float4 v0 = *reinterpret_cast<const float4*>(pb + kk);
                            const float4 v1 = *reinterpret_cast<const float4*>(pb + PLANE + kk);
                            a0 += (v0.x + v0.y) + (v0.z + v0.w);
                            a1 += (v1.x + v1.y) + (v1.z + v1.w);
                            if (HAS_Z) {
                                const float4 v2 = *reinterpret_cast<const float4*>(pb + 2 * PLANE + kk);
                                a2 += (v2.x + v2.y) + (v2.z + v2.w);
                            }
                        }
                        sB[ch] = a0; sA[ch] = a1; yp[ch] = a2;
                    }
                    const float4 uu = *reinterpret_cast<const float4*>(st + o);
                    const float4 g = *reinterpret_cast<const float4*>(st + 3 * TC * DT + o);
                    const float4 dt = *reinterpret_cast<const float4*>(st + 4 * TC * DT + o);
                    const float4 sg = *reinterpret_cast<const float4*>(st + 5 * TC * DT + o);
                    // sA was accumulated against A * log2(e): * ln2 restores sum_n t1 A
                    const float4 ddt = make_float4(fmaf(sA[0], kLn2, uu.x * sB[0]) * sg.x, fmaf(sA[1], kLn2, uu.y * sB[1]) * sg.y,
                                                   fmaf(sA[2], kLn2, uu.z * sB[2]) * sg.z, fmaf(sA[3], kLn2, uu.w * sB[3]) * sg.w);
                    const float4 duo = make_float4(fmaf(g.x, D4.x, dt.x * sB[0]), fmaf(g.y, D4.y, dt.y * sB[1]),
                                                   fmaf(g.z, D4.z, dt.z * sB[2]), fmaf(g.w, D4.w, dt.w * sB[3]));
                    accB.x += ddt.x; accB.y += ddt.y; accB.z += ddt.z; accB.w += ddt.w;
                    accD.x = fmaf(g.x, uu.x, accD.x); accD.y = fmaf(g.y, uu.y, accD.y);
                    accD.z = fmaf(g.z, uu.z, accD.z); accD.w = fmaf(g.w, uu.w, accD.w);
                    const size_t row = row0 + t;
                    *reinterpret_cast<float4*>(p.du + row * p.ld_du + mycol) = duo;
                    *reinterpret_cast<float4*>(p.ddelta + row * p.ld_ddelta + mycol) = ddt;
                    if (HAS_Z) {
                        const float4 zc = *reinterpret_cast<const float4*>(st + 6 * TC * DT + o);
                        *reinterpret_cast<float4*>(p.dz + row * p.ld_dz + mycol) =
                            make_float4(zc.x * fmaf(D4.x, uu.x, yp[0]), zc.y * fmaf(D4.y, uu.y, yp[1]),
                                        zc.z * fmaf(D4.z, uu.z, yp[2]), zc.w * fmaf(D4.w, uu.w, yp[3]));
                    }
                }
            }
            // dB_t[n] / dC_t[n]: sum over the CTA's channels.  The main warps leave one row per (step, warp, channel-lane);
            // the cross-lane / cross-warp sum is done HERE, by warps that otherwise wait for the main warps (the
            // shuffle reduce-scatter it replaces was ~19 % of the main warps' instructions).
            for (int idx = ht * 4; idx < TC * 2 * N; idx += NHELP * 4) {
                const int r = idx / (2 * N), t = k * TC + r;
                if (t < L) {
                    const float* src = s_red + (size_t)r * (NMAINW * DPW * 2 * N) + (idx % (2 * N));
                    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int w = 0; w < NMAINW * DPW; w += 2) {
                        const float4 v0 = *reinterpret_cast<const float4*>(src + w * 2 * N);
                        const float4 v1 = *reinterpret_cast<const float4*>(src + (w + 1) * 2 * N);
                        s0.x += v0.x; s0.y += v0.y; s0.z += v0.z; s0.w += v0.w;
                        s1.x += v1.x; s1.y += v1.y; s1.z += v1.z; s1.w += v1.w;
                    }
                    *reinterpret_cast<float4*>(p.dBC_part + (((size_t)blockIdx.x * p.Bsz + b) * L + t) * (2 * N) + (idx % (2 * N))) =
                        make_float4(s0.x + s1.x, s0.y + s1.y, s0.z + s1.z, s0.w + s1.w);
                }
            }
            __syncwarp();
            if ((ht & 31) == 0) mbar_arrive(bar_freep);
            asm volatile("bar.sync 1, %0;" ::"n"(NHELP) : "memory");
        }
        // per-channel sums over time: combine the NHELP / QPR helper threads that share a quad column
        float* aD = s_acc + (ht / QPR) * DT + myq * 4;
        float* aB = aD + (NHELP / QPR) * DT;
        *reinterpret_cast<float4*>(aD) = accD;
        *reinterpret_cast<float4*>(aB) = accB;
        asm volatile("bar.sync 1, %0;" ::"n"(NHELP) : "memory");
        for (int ch = ht; ch < DT; ch += NHELP) {
            if (d0 + ch < p.D) {
                float sD = 0.f, sBb = 0.f;
#pragma unroll
                for (int w = 0; w < NHELP / QPR; ++w) {
                    sD += s_acc[w * DT + ch];
                    sBb += s_acc[(NHELP / QPR + w) * DT + ch];
                }
                p.dD_part[(size_t)b * p.D + d0 + ch] = sD;
                p.dbias_part[(size_t)b * p.D + d0 + ch] = sBb;
            }
        }
    } else {
        // ------------------------------------------------------------------------------------ main warps
        // thread = CH channels x 4 states.  The two channels of a thread sit in different quads ((warp * CH + c) * DPW
        // + dl), so that a warp's partial-sum stores stay bank-conflict free; what the channels share is B_t / C_t:
        // one LDS.128 (4 shared-memory wavefronts however much of it is a broadcast) now feeds 8 elements instead
        // of 4 -- ncu had the LSU shared-memory pipe as the busiest unit of this kernel (70 %).
        if (NMAIN == 256) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        constexpr int H2 = S / 2, CH = Cfg::CH;
        const int lane = tid & 31, warp = tid >> 5;
        const int dl = lane / LPD, ng = lane % LPD;
        int dloc[CH], poff[CH];
        bool dvalid[CH];
        float2 A2[CH][H2], dA[CH][H2], lam[CH][H2];
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            dloc[c] = (warp * CH + c) * DPW + dl;
            const int d = d0 + dloc[c];
            dvalid[c] = d < p.D;
            poff[c] = (dloc[c] >> 2) * QS + (dloc[c] & 3) * LPD + ng;      // this lane's slot in a partial-sum row
#pragma unroll
            for (int j = 0; j < H2; ++j) {
                // channels past D: any negative A keeps (+inf) * A = -inf at reset steps (0 would give NaN, and this
                // thread's zero contributions still enter the cross-channel dB / dC sums)
                A2[c][j] = dvalid[c] ? f2(sel_A(p.A[(size_t)d * N + ng * S + 2 * j], p.a_log) * kLog2e,
                                          sel_A(p.A[(size_t)d * N + ng * S + 2 * j + 1], p.a_log) * kLog2e)
                                     : f2(-1.f, -1.f);
                dA[c][j] = f2(0.f, 0.f);
                lam[c][j] = f2(0.f, 0.f);                   // a_{t+1} * lambda_{t+1}
            }
        }
        auto ld_ckpt = [&](int k, int c) {                  // state of channel c entering chunk k
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int d = d0 + dloc[c];
            if (k > 0 && dvalid[c]) v = __ldg(reinterpret_cast<const float4*>(p.ckpt + (((size_t)b * p.nckpt + (k - 1)) * p.D + d) * N + ng * S));
            if (k == 0 && dvalid[c] && p.h0 != nullptr) v = __ldg(reinterpret_cast<const float4*>(p.h0 + ((size_t)b * p.D + d) * N + ng * S));
            return v;
        };
        float4 hin[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) hin[c] = ld_ckpt(nch - 1, c);
        for (int cc = 0; cc < nch; ++cc) {
            const int k = nch - 1 - cc;
            const float* st = smem + (cc % NST) * STAGE;
            const float* s_dtA = st + 1 * TC * DT;
            const float* s_du = st + 2 * TC * DT;
            const float* s_g = st + 3 * TC * DT;
            const float* s_dt = st + 4 * TC * DT;            // delta itself (finite at reset steps, where s_dtA holds +inf)
            const float* s_B = st + Cfg::NARR * TC * DT;
            const float* s_C = s_B + TC * N;
            mbar_wait(bar_full(cc % NST), (cc / NST) & 1);
            // ---- phase F: recompute h_t inside the chunk
            float2 h[CH][H2], hent[CH][H2];                 // hent: the state entering the chunk (h_{t-1} of its first step)
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                h[c][0] = hent[c][0] = f2(hin[c].x, hin[c].y);
                h[c][1] = hent[c][1] = f2(hin[c].z, hin[c].w);
                hin[c] = ld_ckpt(k - 1, c);                 // prefetch the next chunk's entry state
            }
            float2 hb[CH][TC][H2];
#pragma unroll
            for (int i = 0; i < TC; ++i) {
                const float4 bq = *reinterpret_cast<const float4*>(s_B + i * N + ng * S);
                const float2 Bv[H2] = {f2(bq.x, bq.y), f2(bq.z, bq.w)};
#pragma unroll
                for (int c = 0; c < CH; ++c) {
                    const float dtA = s_dtA[i * DT + dloc[c]], du = s_du[i * DT + dloc[c]];
#pragma unroll
                    for (int j = 0; j < H2; ++j) {
                        const float2 e = __fmul2_rn(f2(dtA, dtA), A2[c][j]);
                        const float2 a = f2(ex2f(e.x), ex2f(e.y));
                        h[c][j] = __ffma2_rn(a, h[c][j], __fmul2_rn(f2(du, du), Bv[j]));
                        hb[c][i][j] = h[c][j];
                    }
                }
            }
            // the partial-sum planes and s_red are single-buffered: the helpers must have finished chunk cc-1
            if (cc >= 1) mbar_wait(bar_freep, (cc - 1) & 1);
            // ---- phase R: adjoint recurrence, latest step first
#pragma unroll
            for (int i = TC - 1; i >= 0; --i) {
                const float4 bq = *reinterpret_cast<const float4*>(s_B + i * N + ng * S);
                const float4 cq = *reinterpret_cast<const float4*>(s_C + i * N + ng * S);
                const float2 Bv[H2] = {f2(bq.x, bq.y), f2(bq.z, bq.w)};
                const float2 Cv[H2] = {f2(cq.x, cq.y), f2(cq.z, cq.w)};
                float2 dB2[H2], dC2[H2];                    // the thread's channels are summed here (packed adds),
#pragma unroll                                              // the warp's by the reduce-scatter below
                for (int c = 0; c < CH; ++c) {
                    const float dtA = s_dtA[i * DT + dloc[c]], du = s_du[i * DT + dloc[c]], g = s_g[i * DT + dloc[c]];
                    const float dt = s_dt[i * DT + dloc[c]];
                    const float2 g2 = f2(g, g), du2 = f2(du, du), dt2 = f2(dt, dt);
                    float2 sB2 = f2(0.f, 0.f), sA2 = f2(0.f, 0.f), yp2 = f2(0.f, 0.f);
#pragma unroll
                    for (int j = 0; j < H2; ++j) {
                        const float2 l = __ffma2_rn(g2, Cv[j], lam[c][j]);
                        const float2 dc = __fmul2_rn(g2, hb[c][i][j]);
                        const float2 db = __fmul2_rn(l, du2);
                        dC2[j] = c == 0 ? dc : __fadd2_rn(dC2[j], dc);
                        dB2[j] = c == 0 ? db : __fadd2_rn(dB2[j], db);
                        sB2 = __ffma2_rn(l, Bv[j], sB2);
                        if (HAS_Z) yp2 = __ffma2_rn(hb[c][i][j], Cv[j], yp2);
                        const float2 e = __fmul2_rn(f2(dtA, dtA), A2[c][j]);
                        lam[c][j] = __fmul2_rn(f2(ex2f(e.x), ex2f(e.y)), l);           // a_t * lambda_t (0 at a reset step)
                        // lambda_t * a_t * h_{t-1}: the product with the stored previous state, no h_t - du B_t cancellation
                        const float2 t1 = __fmul2_rn(lam[c][j], i > 0 ? hb[c][i > 0 ? i - 1 : 0][j] : hent[c][j]);
                        dA[c][j] = __ffma2_rn(t1, dt2, dA[c][j]);
                        sA2 = __ffma2_rn(t1, A2[c][j], sA2);
                    }
                    float* pl = s_pl + i * PROW + poff[c];
                    pl[0] = sB2.x + sB2.y;
                    pl[PLANE] = sA2.x + sA2.y;
                    if (HAS_Z) pl[2 * PLANE] = yp2.x + yp2.y;
                }
                // this lane's dB | dC (already summed over the thread's channels): one row per (step, warp, channel-lane)
                float* rr = s_red + ((size_t)(i * NMAINW + warp) * DPW + dl) * (2 * N) + ng * S;
                *reinterpret_cast<float4*>(rr) = make_float4(dB2[0].x, dB2[0].y, dB2[1].x, dB2[1].y);
                *reinterpret_cast<float4*>(rr + N) = make_float4(dC2[0].x, dC2[0].y, dC2[1].x, dC2[1].y);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_done);
        }
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (dvalid[c]) {
                float* a = p.dA_part + ((size_t)b * p.D + d0 + dloc[c]) * N + ng * S;
                if (p.a_log) {                               // A = -exp(A_log): d A_log = dA * A (A2 holds A * log2 e)
#pragma unroll
                    for (int j = 0; j < H2; ++j) dA[c][j] = __fmul2_rn(dA[c][j], __fmul2_rn(A2[c][j], f2(kLn2, kLn2)));
                }
                *reinterpret_cast<float4*>(a) = make_float4(dA[c][0].x, dA[c][0].y, dA[c][1].x, dA[c][1].y);
            }
        }
    }
}

template <int N>
constexpr size_t sel_fwd_smem() {
    return SelFwdCfg<N>::SMEM;
}
template <int N>
constexpr size_t sel_bwd_smem() {
    return SelBwdCfg<N>::SMEM;
}

}  // namespace rorl

using namespace rorl;

static int g_sel_dbg = 0;
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int N, bool HAS_Z, bool SP>
static int launch_fwd(const SelFwdParams& p, int64_t B, cudaStream_t stream) {
    auto kern = selscan_fwd_kernel<N, HAS_Z, SP>;
    constexpr size_t smem = sel_fwd_smem<N>();
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((p.D + SelFwdCfg<N>::DT - 1) / SelFwdCfg<N>::DT), (unsigned)B);
    kern<<<grid, SelFwdCfg<N>::NTHREADS, smem, stream>>>(p);
    RORL_RETURN_LAUNCH();
}
template <int N, bool HAS_Z, bool SP>
static int launch_bwd(const SelBwdParams& p, int64_t B, cudaStream_t stream) {
    auto kern = selscan_bwd_kernel<N, HAS_Z, SP>;
    constexpr size_t smem = sel_bwd_smem<N>();
    static bool once = (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), true);
    (void)once;
    dim3 grid((unsigned)((p.D + SelBwdCfg<N>::DT - 1) / SelBwdCfg<N>::DT), (unsigned)B);
    kern<<<grid, SelBwdCfg<N>::NTHREADS, smem, stream>>>(p);
    RORL_RETURN_LAUNCH();
}

#define SEL_DISPATCH(FN, P, B, STREAM)                                         \
    do {                                                                       \
        if (N == 16) {                                                         \
            if (has_z) return sp ? FN<16, true, true>(P, B, STREAM) : FN<16, true, false>(P, B, STREAM);   \
            return sp ? FN<16, false, true>(P, B, STREAM) : FN<16, false, false>(P, B, STREAM);            \
        } else if (N == 32) {                                                  \
            if (has_z) return sp ? FN<32, true, true>(P, B, STREAM) : FN<32, true, false>(P, B, STREAM);   \
            return sp ? FN<32, false, true>(P, B, STREAM) : FN<32, false, false>(P, B, STREAM);            \
        } else {                                                               \
            if (has_z) return sp ? FN<64, true, true>(P, B, STREAM) : FN<64, true, false>(P, B, STREAM);   \
            return sp ? FN<64, false, true>(P, B, STREAM) : FN<64, false, false>(P, B, STREAM);            \
        }                                                                      \
    } while (0)

extern "C" {

int rorl_selscan_dtile(int64_t N) {
    if (N == 16) return SelBwdCfg<16>::DT;
    if (N == 32) return SelBwdCfg<32>::DT;
    if (N == 64) return SelBwdCfg<64>::DT;
    return RORL_ERR_SHAPE;
}

int rorl_selscan_ckpt_every(void) { return kCkptEvery; }

/* diagnostic only (tools/): bit mask that disables phases of the forward kernel for phase timing */
void rorl_selscan_debug(int v) { g_sel_dbg = v; }

int rorl_selscan_fwd(const float* u, const float* delta, const float* A, const float* Bm, const float* Cm,
                     const float* Dskip, const float* z, const float* delta_bias, const float* start, const float* h0,
                     float* y, float* ckpt, float* last_state, int64_t B, int64_t L, int64_t D, int64_t N, int64_t ld_u,
                     int64_t ld_delta, int64_t ld_z, int64_t ld_B, int64_t ld_C, int64_t ld_y, int delta_softplus,
                     cudaStream_t stream) {
    if (!u || !delta || !A || !Bm || !Cm || !y) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (N != 16 && N != 32 && N != 64) return RORL_ERR_SHAPE;
    if (D % 4 || ld_u % 4 || ld_delta % 4 || ld_B % 4 || ld_C % 4 || ld_y % 4 || (z && ld_z % 4)) return RORL_ERR_ALIGN;
    if (!aligned16(u) || !aligned16(delta) || !aligned16(Bm) || !aligned16(Cm) || !aligned16(y) ||
        (z && !aligned16(z)) || (delta_bias && !aligned16(delta_bias)) || (ckpt && !aligned16(ckpt)) ||
        (last_state && !aligned16(last_state)) || (h0 && !aligned16(h0)))
        return RORL_ERR_ALIGN;
    SelFwdParams p;
    p.u = u; p.delta = delta; p.z = z; p.Bm = Bm; p.Cm = Cm; p.A = A; p.Dskip = Dskip; p.dbias = delta_bias;
    p.start = start; p.h0 = h0; p.y = y; p.ckpt = ckpt; p.last_state = last_state;
    p.L = (int)L; p.D = (int)D;
    p.ld_u = (int)ld_u; p.ld_delta = (int)ld_delta; p.ld_z = (int)ld_z; p.ld_B = (int)ld_B; p.ld_C = (int)ld_C;
    p.ld_y = (int)ld_y;
    p.nckpt = (int)(L / kCkptEvery);
    p.dbg = g_sel_dbg;
    const bool has_z = z != nullptr, sp = (delta_softplus & 1) != 0;
    p.a_log = (delta_softplus & 2) != 0;
    SEL_DISPATCH(launch_fwd, p, B, stream);
}

int rorl_selscan_bwd(const float* u, const float* delta, const float* A, const float* Bm, const float* Cm,
                     const float* Dskip, const float* z, const float* delta_bias, const float* start, const float* h0,
                     const float* dy, const float* ckpt, float* du, float* ddelta, float* dz, float* dBC_part,
                     float* dA_part, float* dD_part, float* dbias_part, int64_t B, int64_t L, int64_t D, int64_t N,
                     int64_t ld_u, int64_t ld_delta, int64_t ld_z, int64_t ld_B, int64_t ld_C, int64_t ld_dy,
                     int64_t ld_du, int64_t ld_ddelta, int64_t ld_dz, int delta_softplus, cudaStream_t stream) {
    if (!u || !delta || !A || !Bm || !Cm || !dy || !du || !ddelta || !dBC_part || !dA_part || !dD_part ||
        !dbias_part)
        return RORL_ERR_ARG;
    if ((z != nullptr) != (dz != nullptr)) return RORL_ERR_ARG;
    if (L > kCkptEvery && !ckpt) return RORL_ERR_ARG;
    if (B <= 0 || L <= 0 || D <= 0 || B > 65535) return RORL_ERR_SHAPE;
    if (N != 16 && N != 32 && N != 64) return RORL_ERR_SHAPE;
    if (D % 4 || ld_u % 4 || ld_delta % 4 || ld_B % 4 || ld_C % 4 || ld_dy % 4 || ld_du % 4 || ld_ddelta % 4 ||
        (z && (ld_z % 4 || ld_dz % 4)))
        return RORL_ERR_ALIGN;
    if (!aligned16(u) || !aligned16(delta) || !aligned16(Bm) || !aligned16(Cm) || !aligned16(dy) || !aligned16(du) ||
        !aligned16(ddelta) || (z && (!aligned16(z) || !aligned16(dz))) || (delta_bias && !aligned16(delta_bias)) ||
        (ckpt && !aligned16(ckpt)) || !aligned16(dA_part) || (h0 && !aligned16(h0)))
        return RORL_ERR_ALIGN;
    SelBwdParams p;
    p.u = u; p.delta = delta; p.z = z; p.Bm = Bm; p.Cm = Cm; p.A = A; p.Dskip = Dskip; p.dbias = delta_bias;
    p.start = start; p.h0 = h0; p.dy = dy; p.ckpt = ckpt;
    p.du = du; p.ddelta = ddelta; p.dz = dz; p.dBC_part = dBC_part; p.dA_part = dA_part; p.dD_part = dD_part;
    p.dbias_part = dbias_part;
    p.L = (int)L; p.D = (int)D; p.Bsz = (int)B;
    p.ld_u = (int)ld_u; p.ld_delta = (int)ld_delta; p.ld_z = (int)ld_z; p.ld_B = (int)ld_B; p.ld_C = (int)ld_C;
    p.ld_dy = (int)ld_dy; p.ld_du = (int)ld_du; p.ld_ddelta = (int)ld_ddelta; p.ld_dz = (int)ld_dz;
    p.nckpt = (int)(L / kCkptEvery);
    const bool has_z = z != nullptr, sp = (delta_softplus & 1) != 0;
    p.a_log = (delta_softplus & 2) != 0;
    SEL_DISPATCH(launch_bwd, p, B, stream);
}

}  // extern "C"
