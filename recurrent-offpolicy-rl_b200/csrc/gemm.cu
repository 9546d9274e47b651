// tcgen05 tensor-core GEMM with fp32 parity (3xTF32 split accumulation), sm_100a.
//
//   D[g][M, N] = act( A[g][M, K] * B[g][N, K]^T + bias[g][N] )        g = 0..G-1 (ensemble members)
//
// Replaces the cuBLAS fp32 SGEMMs behind nn.Linear / EnsembleLinear on the update path
// (ref: offpolicy_rnn/models/ensemble_linear_model.py:36-49 einsum -> bmm; smamba in/x/dt/out_proj,
// ref: offpolicy_rnn/models/smamba/mamba.py:176,231-233,252).  The reference runs these in true fp32
// (TF32 is never enabled there), so a single-pass TF32 MMA is not enough for the 1e-3 parity budget:
// each fp32 operand is split as x = hi + lo (hi = top 19 bits = a TF32 value, lo = x - hi) and the
// product is accumulated as lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator (~2^-21 relative).
//
// Structure (one persistent CTA per SM, 10 warps):
//   warp 0      TMA producer: cp.async.bulk.tensor (SWIZZLE_128B, 32 fp32 = 128 B of K per row) into a
//               3-stage shared-memory ring, mbarrier complete_tx
//   warps 2-5   splitters: read the raw fp32 tiles, write the `lo` tiles in the identical swizzled layout,
//               fence.proxy.async, arrive
//   warp 1      MMA issuer (one elected lane): 4 k-steps x 3 tcgen05.mma.kind::tf32 (M128 x N128 x K8) per
//               stage into one of two TMEM accumulators; tcgen05.commit frees the stage / publishes the tile
//   warps 6-13  epilogue: two warps per TMEM lane quarter, each owning two of the four 32-column chunks:
//               tcgen05.ld (32 lanes x 32 columns per instruction), + bias, branch-free ELU, shared-memory
//               transposition, 128-B-row global stores; overlaps the next tile's main loop (double-buffered
//               TMEM).  One warp per scheduler was latency bound (ncu: the epilogue, not the MMA, set the tile
//               time), hence eight
// Operand layouts: the TN variant takes both operands K-major (reduction dimension contiguous): forward
// (x, W[N,K]) and data-gradient (dY, W^T materialised by the caller; weights are tiny).  The NT variant takes
// both operands MN-major (reduction dimension = rows): weight gradients dW[n,k] = sum_m dY[m,n] X[m,k] straight
// from the row-major activations, with split-K over the (very long) token dimension and per-split partial
// outputs the caller sums (deterministic, no atomics).  tcgen05.mma.kind::tf32 multiplies K-major operands only, so
// in the NT variant the splitter warps transpose each tile in shared memory while they split it.
#include "common.cuh"
#include "tc.cuh"

namespace rorl {

constexpr int kGemmBM = 128, kGemmBN = 128, kGemmBK = 32;     // BK fp32 = one 128-byte swizzle row
constexpr int kGemmThreads = 448;                                // TMA, MMA, 4 splitter and 8 epilogue warps
constexpr int kGemmEpiWarps = 8;
constexpr int kTileBytes = kGemmBM * kGemmBK * 4;             // 16 KiB: the A tile (and the B tile at BN = 128)
constexpr int kStagingBytes = kGemmEpiWarps * 32 * 128;       // per epilogue warp: 32 rows x 32 fp32 columns
// Tile configuration.  The kernel is bound by L2 -> SM operand traffic and shared-memory bandwidth, not by the MMA
// pipe (fp32 operands: a 128 x 128 tile pulls 32 KiB per k-step for 1 MFLOP; measured: the single-pass variant runs
// at the same ~300 TFLOP/s whatever the shape).  BN = 256 (TN variant, N > 128) reads the A tile once for twice the
// columns: 0.75x the operand bytes per FLOP and half the A-splitting work, at the cost of a 2-deep instead of a
// 3-deep ring (227 KiB of shared memory) and both 256-column TMEM accumulators (512 columns).
// BK = 16 (64-byte swizzle rows) halves the stage so that the same 192 KiB ring is twice as deep (4 stages at BN = 256,
// 6 at BN = 128): the TMA round trip (~1 us from L2) then hides behind the other stages' MMAs, which a 2-deep ring of
// 96-KiB stages cannot do.
template <int BN, int BK = kGemmBK>
struct GemmCfg {
    static constexpr int kTileA = kGemmBM * BK * 4;                // A tile bytes
    static constexpr int kTileB = BN * BK * 4;                     // B tile bytes
    static constexpr int kRaw = kTileA + kTileB;                   // A_raw | B_raw
    static constexpr int kStage = 2 * kRaw;                        // A_raw | B_raw | A_lo | B_lo
    static constexpr int kStages = (192 * 1024) / kStage;
    static constexpr int kSmem = kStages * kStage + kStagingBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct GemmParams {
    float* D;
    float* Dpre;      // optional: pre-activation (A B^T + bias) written next to D when act != 0
    const float* bias;
    int M, N, K, G;
    long long ldd, strideD, strideBias;
    int a_batched, b_batched, act, passes, reduce_g;
    int dbg;
    int splits;                 // NT variant: split-K factor; partial s is written at D + s * strideSplit
    long long strideSplit;
};

// ELU(alpha = 1), branch-free: ex2 on min(x, 0) and a select (5 instructions, no divergence).  Absolute error
// ~1e-7 (ex2.approx is 2 ulp on a result in (0, 1]); relative to the layer's activations that is far inside
// the 1e-3 parity budget.
__device__ __forceinline__ float elu1(float x) {
    const float e = ex2f(fminf(x, 0.f) * kLog2e) - 1.0f;
    return x > 0.f ? x : e;
}

template <bool MN, int BN, int BK>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const GemmParams p) {
    static_assert(BN == 128 || (BN == 256 && !MN), "BN = 256 exists for the TN variant only");
    static_assert(BK == 32 || (BK == 16 && !MN), "BK = 16 exists for the TN variant only");
    using Cfg = GemmCfg<BN, BK>;
    constexpr int kGemmStages = Cfg::kStages, kStageBytes = Cfg::kStage, kRawBytes = Cfg::kRaw, kTileB = Cfg::kTileB;
    constexpr int kTileBytes = Cfg::kTileA;                       // these three shadow the namespace defaults inside the kernel
    constexpr int kGemmBN = BN;
    constexpr int kGemmBK = BK;
    (void)kTileB;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bars = base + kGemmStages * kStageBytes + kStagingBytes;
    // barrier map (8 B each): full_raw[S], full_split[S], empty[S], tmem_full[2], tmem_empty[2], then tmem ptr
    auto bar_full_raw = [&](int s) { return bars + 8u * s; };
    auto bar_full_split = [&](int s) { return bars + 8u * (kGemmStages + s); };
    auto bar_empty = [&](int s) { return bars + 8u * (2 * kGemmStages + s); };
    auto bar_tfull = [&](int a) { return bars + 8u * (3 * kGemmStages + a); };
    auto bar_tempty = [&](int a) { return bars + 8u * (3 * kGemmStages + 2 + a); };
    uint8_t* staging = base_ptr + kGemmStages * kStageBytes;                      // 16 KiB, 4 KiB per epilogue warp
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(base_ptr + kGemmStages * kStageBytes + kStagingBytes + 8 * (3 * kGemmStages + 4));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool split = p.passes == 3;
    const int tilesM = (p.M + kGemmBM - 1) / kGemmBM, tilesN = (p.N + kGemmBN - 1) / kGemmBN;
    // reduce_g: the G operand pairs are summed into ONE output (the reduction runs over (g, k)), used for the
    // data-gradient of an ensemble layer whose input is shared by all members.
    const int KTg = (p.K + kGemmBK - 1) / kGemmBK;
    const int nsplit = MN ? p.splits : 1;
    const int ntiles = tilesM * tilesN * (p.reduce_g ? 1 : p.G) * nsplit;
    const int KTs = (KTg + nsplit - 1) / nsplit;                                  // k-tiles per split (NT variant)
    const int KT = MN ? KTs : (p.reduce_g ? KTg * p.G : KTg);

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGemmStages; ++s) {
            mbar_init(bar_full_raw(s), 1);
            mbar_init(bar_full_split(s), 4);
            mbar_init(bar_empty(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(bar_tfull(a), 1);
            mbar_init(bar_tempty(a), kGemmEpiWarps);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32((const void*)tmem_slot)), "r"(2 * kGemmBN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB) : "memory");
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int tn = tile % tilesN, tm = (tile / tilesN) % tilesM;
                const int g = (tile / (tilesN * tilesM)) % p.G, sp = tile / (tilesN * tilesM * p.G);
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kGemmStages;
                    mbar_wait(bar_empty(s), ((it / kGemmStages) & 1) ^ 1);
                    mbar_expect_tx(bar_full_raw(s), kRawBytes);
                    const uint32_t st = base + s * kStageBytes;
                    if (MN) {
                        // operands [rows = reduction][cols = MN]: 4 boxes of 32 columns x 32 rows per operand
                        const int r0 = (sp * KTs + kt) * kGemmBK;                 // rows beyond the tensor are zero-filled
#pragma unroll
                        for (int bI = 0; bI < 4; ++bI) {
                            tma_load_3d(st + bI * (kGemmBK * 128), &mapA, bar_full_raw(s), tm * kGemmBM + 32 * bI, r0, p.a_batched ? g : 0);
                            tma_load_3d(st + kTileBytes + bI * (kGemmBK * 128), &mapB, bar_full_raw(s), tn * kGemmBN + 32 * bI, r0, p.b_batched ? g : 0);
                        }
                    } else {
                        const int gg = p.reduce_g ? kt / KTg : g, kk = p.reduce_g ? kt % KTg : kt;
                        tma_load_3d(st, &mapA, bar_full_raw(s), kk * kGemmBK, tm * kGemmBM, p.a_batched ? gg : 0);
                        tma_load_3d(st + kTileBytes, &mapB, bar_full_raw(s), kk * kGemmBK, tn * kGemmBN, p.b_batched ? gg : 0);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | 
                                   ((uint32_t)(kGemmBN >> 3) << 17) | ((uint32_t)(kGemmBM >> 4) << 24);
            uint32_t it = 0, tcount = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
                const int acc = tcount & 1;
                mbar_wait(bar_tempty(acc), ((tcount >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * kGemmBN;
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kGemmStages;
                    const uint32_t ph = (it / kGemmStages) & 1;
                    mbar_wait(bar_full_raw(s), ph);
                    if (split || MN) mbar_wait(bar_full_split(s), ph);
                    tc_fence_after();
                    const uint32_t st = base + s * kStageBytes;
                    auto mk = [](uint32_t a) { return BK == 16 ? make_kmajor_desc_sw64(a) : make_kmajor_desc(a); };
                    const uint64_t a_hi = mk(st), b_hi = mk(st + kTileBytes);
                    const uint64_t a_lo = mk(st + kRawBytes), b_lo = mk(st + kRawBytes + kTileBytes);
#pragma unroll
                    for (int k = 0; k < kGemmBK / 8; ++k) {
                        const uint64_t adv = (uint64_t)(k * 8 * 4 >> 4);        // 32 B per k-step inside the 128-B swizzle row
                        const uint32_t first = (kt | k) == 0 ? 0u : 1u;
                        if (split) {
                            umma_tf32(tmem_d, a_lo + adv, b_hi + adv, idesc, first);
                            umma_tf32(tmem_d, a_hi + adv, b_lo + adv, idesc, 1u);
                            umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, 1u);
                        } else {
                            umma_tf32(tmem_d, a_hi + adv, b_hi + adv, idesc, first);
                        }
                    }
                    umma_commit(bar_empty(s));                                   // stage free once these MMAs retire
                }
                umma_commit(bar_tfull(acc));                                     // accumulator complete
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------ splitters (hi/lo [+ transpose])
        const int t = threadIdx.x - 64;                                          // 0..127
        if (MN) {
            // TMA delivered each operand as 4 blocks of [32 reduction rows][32 MN columns] (SWIZZLE_128B).  kind::tf32
            // only multiplies K-major operands (with the MN-major bits set the MMA returns zeros on this part), so
            // the split doubles as a transposition: every thread pulls its 64 values into registers, all 128
            // splitter threads meet at a named barrier, then hi / lo are written as K-major SWIZZLE_128B rows
            // (row = MN index, 32 reduction values = 128 B) over the raw tile / into the lo tile.
            const int blk = t >> 5, r = t & 31;                                  // warp <-> MN block, lane <-> reduction row
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kGemmStages;
                    mbar_wait(bar_full_raw(s), (it / kGemmStages) & 1);
                    uint8_t* st = base_ptr + s * kStageBytes;
                    float4 v[2][8];
#pragma unroll
                    for (int op = 0; op < 2; ++op)
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            v[op][c] = *reinterpret_cast<const float4*>(st + op * kTileBytes + blk * 4096 + r * 128 + ((c ^ (r & 7)) << 4));
                    asm volatile("bar.sync 1, 128;" ::: "memory");
#pragma unroll
                    for (int op = 0; op < 2; ++op) {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            const float e[4] = {v[op][c].x, v[op][c].y, v[op][c].z, v[op][c].w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int mn = blk * 32 + c * 4 + i;
                                const int off = mn * 128 + (((r >> 2) ^ (mn & 7)) << 4) + ((r & 3) << 2);
                                const float hi = __uint_as_float((__float_as_uint(e[i]) + 0x1000u) & 0xFFFFE000u);
                                *reinterpret_cast<float*>(st + op * kTileBytes + off) = hi;
                                if (split) *reinterpret_cast<float*>(st + kRawBytes + op * kTileBytes + off) = e[i] - hi;
                            }
                        }
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full_split(s));
                }
            }
        } else if (split) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                for (int kt = 0; kt < KT; ++kt, ++it) {
                    const int s = it % kGemmStages;
                    mbar_wait(bar_full_raw(s), (it / kGemmStages) & 1);
                    uint8_t* st = base_ptr + s * kStageBytes;
#pragma unroll 4
                    for (int i = 0; i < kRawBytes / 16 / 128; ++i) {
                        const int off = (t + 128 * i) * 16;
                        float4 v = *reinterpret_cast<const float4*>(st + off);
                        // The tensor core reads a TF32 operand by IGNORING the low 13 mantissa bits of the fp32 word, so
                        // the raw tile already is `hi = trunc(x)` and only the exact remainder lo = x - trunc(x) has to be
                        // written (|lo| < 2^-10 |x|; the neglected lo*lo term is 2^-20 relative).  This removes a third of
                        // the splitters' shared-memory traffic, which ncu showed to be what the MMA warp waits for.
                        float4 lo;
                        lo.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                        lo.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                        lo.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                        lo.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                        *reinterpret_cast<float4*>(st + kRawBytes + off) = lo;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full_split(s));
                }
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue
        // TMEM -> registers (one accumulator row per lane) -> + bias / ELU -> swizzled per-warp staging tile in
        // shared memory -> row-contiguous 128-B global stores (4 full lines per warp instruction).
        const int q = warp & 3;                                                  // TMEM lane quarter this warp may read
        const int hf = (warp - 6) >> 2;                                          // which pair of 32-column chunks
        uint8_t* stg = staging + (warp - 6) * 4096;
        uint32_t tcount = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
            const int tn = tile % tilesN, tm = (tile / tilesN) % tilesM;
            const int g = (tile / (tilesN * tilesM)) % p.G, sp = tile / (tilesN * tilesM * p.G);
            const int acc = tcount & 1;
            mbar_wait(bar_tfull(acc), (tcount >> 1) & 1);
            tc_fence_after();
            const long long obase = (long long)g * p.strideD + (long long)sp * p.strideSplit;
            const int row0 = tm * kGemmBM + q * 32;
            const float* bias = p.bias ? p.bias + (long long)g * p.strideBias : nullptr;
            auto flush = [&](const float4 (&o)[8], float* out, int col0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = o[j];
                __syncwarp();
#pragma unroll
                for (int rr = 0; rr < 8; ++rr) {
                    const int rloc = rr * 4 + (lane >> 3), ch = lane & 7;
                    const float4 v = *reinterpret_cast<const float4*>(stg + rloc * 128 + ((ch ^ (rloc & 7)) << 4));
                    const int grow = row0 + rloc, col = col0 + ch * 4;
                    if (grow < p.M && col < p.N)
                        *reinterpret_cast<float4*>(out + obase + (long long)grow * p.ldd + col) = v;
                }
                __syncwarp();
            };
            // this warp's chunks are pulled out of TMEM two at a time; after the last pair the accumulator is released
            // to the MMA warp, before the (long) bias / activation / store tail of that pair
            constexpr int kHalves = kGemmBN / 128;
#pragma unroll 1
            for (int half = 0; half < kHalves; ++half) {
                const int chunk0 = (kGemmBN / 64) * hf + 2 * half;               // first of this warp's two 32-column chunks
                uint32_t r[2][32];
#pragma unroll
                for (int cc = 0; cc < 2; ++cc)
                    tmem_ld32(tmem_base + acc * kGemmBN + (chunk0 + cc) * 32 + ((uint32_t)(q * 32) << 16), r[cc]);
                tmem_ld_wait();
                if (half == kHalves - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_tempty(acc));
                }
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int col0 = tn * kGemmBN + (chunk0 + cc) * 32;
                    if (col0 >= p.N) break;                                      // warp-uniform
                    float4 o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        o[j] = make_float4(__uint_as_float(r[cc][4 * j]), __uint_as_float(r[cc][4 * j + 1]), __uint_as_float(r[cc][4 * j + 2]),
                                           __uint_as_float(r[cc][4 * j + 3]));
                        if (bias && col0 + 4 * j < p.N) {
                            const float4 bv = __ldg(reinterpret_cast<const float4*>(bias + col0 + 4 * j));
                            o[j].x += bv.x; o[j].y += bv.y; o[j].z += bv.z; o[j].w += bv.w;
                        }
                    }
                    if (p.Dpre) flush(o, p.Dpre, col0);
                    if (p.act == 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) { o[j].x = elu1(o[j].x); o[j].y = elu1(o[j].y); o[j].z = elu1(o[j].z); o[j].w = elu1(o[j].w); }
                    }
                    flush(o, p.D, col0);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kGemmBN));
    }
}

// gemm_bf16.cu: the two-term bf16 split ("bf16x3", passes == 2)
int gemm_tn_bf16x3(const float* A, const float* B, const float* bias, float* D, float* Dpre, int64_t M, int64_t N, int64_t K, int64_t G,
                   int64_t lda, int64_t ldb, int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD,
                   int64_t strideBias, int act, int reduce_g, int transb, int single, int force_bn, void* bsplit, cudaStream_t stream);

int split_bf16_multi(const void* jobs, int njobs, cudaStream_t stream);
int gemm_nt_bf16x3(const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t R, int64_t G, int64_t lda, int64_t ldb,
                   int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD, int64_t splits, int64_t strideSplit,
                   int single, cudaStream_t stream);

static int g_gemm_dbg = 0;
static int g_gemm_bn = 0;
static int g_gemm_bk = 0;
constexpr int kGemmDefaultBK = 16;      // measured 3-11 % faster than 32 on every update shape (profiles/r01g_kernels.log)
static int gemm_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaFuncSetAttribute(gemm_kernel<false, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, 32>::kSmem);
        cudaFuncSetAttribute(gemm_kernel<false, 256, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, 32>::kSmem);
        cudaFuncSetAttribute(gemm_kernel<false, 128, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, 16>::kSmem);
        cudaFuncSetAttribute(gemm_kernel<false, 256, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<256, 16>::kSmem);
        cudaFuncSetAttribute(gemm_kernel<true, 128, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<128, 32>::kSmem);
    }
    return sms;
}

}  // namespace rorl

using namespace rorl;

extern "C" {

// D[g] = act(A[g] B[g]^T + bias[g]);  A [G?][M, K] (lda), B [G?][N, K] (ldb), D [G][M, N] (ldd); strides in floats,
// strideA / strideB == 0 means the operand is shared by all g.  act: 0 none, 1 ELU (Dpre, if not NULL, receives the
// pre-activation in D's layout so that a backward can form the exact ELU derivative).  passes: 3 = fp32-parity
// 3xTF32, 1 = single TF32.  Requirements: K % 4 == 0, N % 4 == 0, lda/ldb/ldd % 4 == 0, 16-byte aligned bases.
int rorl_gemm_tn(const float* A, const float* B, const float* bias, float* D, float* Dpre, int64_t M, int64_t N, int64_t K, int64_t G,
                 int64_t lda, int64_t ldb, int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD,
                 int64_t strideBias, int act, int passes, int reduce_g, int transb, void* work, cudaStream_t stream) {
    if (!A || !B || !D) return RORL_ERR_ARG;
    const bool bf = passes == 2 || passes == 4;                  // the bf16 kernel: two-term split (2) or its hi * hi term alone (4)
    if ((transb & ~3) || (transb && !bf)) return RORL_ERR_ARG;   // only the pre-splitting form re-lays B out / takes a kept copy
    if ((act & ~7) || (act & 3) == 3 || ((act & 6) && !bf)) return RORL_ERR_ARG;   // GELU, accumulate: bf16 kernel only
    if (M <= 0 || N <= 0 || K <= 0 || G <= 0) return RORL_ERR_SHAPE;
    if (K % 4 || N % 4 || lda % 4 || ldb % 4 || ldd % 4 || strideA % 4 || strideB % 4 || strideD % 4 || strideBias % 4)
        return RORL_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D) |
         reinterpret_cast<uintptr_t>(bias) | reinterpret_cast<uintptr_t>(Dpre)) & 15)
        return RORL_ERR_ALIGN;
    if (passes < 1 || passes > 4) return RORL_ERR_ARG;
    if (reduce_g && (!strideA || !strideB)) return RORL_ERR_ARG;
    if (bf)
        return gemm_tn_bf16x3(A, B, bias, D, Dpre, M, N, K, G, lda, ldb, ldd, strideA, strideB, strideD, strideBias, act, reduce_g,
                              transb, passes == 4, g_gemm_bn, work, stream);
    CUtensorMap mapA, mapB;
    const bool wide = N > kGemmBN && g_gemm_bn != 128;            // N > 128: 128 x 256 tiles (see GemmCfg)
    const int bn = wide ? 256 : kGemmBN;
    const int bk = g_gemm_bk == 32 ? 32 : (g_gemm_bk == 16 ? 16 : kGemmDefaultBK);
    int rc = make_map(&mapA, A, M, K, lda, strideA ? G : 1, strideA, kGemmBM, bk);
    if (rc) return rc;
    rc = make_map(&mapB, B, N, K, ldb, strideB ? G : 1, strideB, bn, bk);
    if (rc) return rc;
    GemmParams p;
    p.D = D; p.Dpre = Dpre; p.bias = bias; p.M = (int)M; p.N = (int)N; p.K = (int)K; p.G = (int)G;
    p.ldd = ldd; p.strideD = strideD; p.strideBias = strideBias;
    p.a_batched = strideA != 0; p.b_batched = strideB != 0; p.act = act; p.passes = passes; p.reduce_g = reduce_g != 0;
    p.splits = 1; p.strideSplit = 0; p.dbg = g_gemm_dbg;
    const int sms = gemm_sms();
    const long long tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + bn - 1) / bn) * (reduce_g ? 1 : G);
    const int grid = (int)(tiles < sms ? tiles : sms);
    if (wide && bk == 16)
        gemm_kernel<false, 256, 16><<<grid, kGemmThreads, GemmCfg<256, 16>::kSmem, stream>>>(mapA, mapB, p);
    else if (wide)
        gemm_kernel<false, 256, 32><<<grid, kGemmThreads, GemmCfg<256, 32>::kSmem, stream>>>(mapA, mapB, p);
    else if (bk == 16)
        gemm_kernel<false, 128, 16><<<grid, kGemmThreads, GemmCfg<128, 16>::kSmem, stream>>>(mapA, mapB, p);
    else
        gemm_kernel<false, 128, 32><<<grid, kGemmThreads, GemmCfg<128, 32>::kSmem, stream>>>(mapA, mapB, p);
    RORL_RETURN_LAUNCH();
}

// B operands whose split copy the CALLER maintains (transb bit 1 of rorl_gemm_tn): (re)build many of them in one launch.
// `jobs`: device array of njobs records {const float* src; void* dst; int32 N, K, G, transposed; int64 ld, gs} (48 bytes
// each, natural alignment); job j writes G groups of [N, K] bf16 hi at dst and lo at dst + 2 * G * N * K bytes.
int rorl_split_bf16_multi(const void* jobs, int64_t njobs, cudaStream_t stream) {
    if (!jobs && njobs > 0) return RORL_ERR_ARG;
    if (njobs < 0 || njobs > 65535) return RORL_ERR_SHAPE;
    return rorl::split_bf16_multi(jobs, (int)njobs, stream);
}

// bytes of `work` rorl_gemm_tn needs for passes == 2 (the B operand's bf16 hi | lo copies); 0 for the TF32 forms
int64_t rorl_gemm_tn_work_bytes(int64_t N, int64_t K, int64_t G, int64_t strideB, int passes) {
    if (passes != 2 && passes != 4) return 0;
    return 2 * (strideB ? G : 1) * N * K * 2;
}

void rorl_gemm_debug(int v) { g_gemm_dbg = v; }
/* diagnostic / A-B benchmarks only: 128 forces the 128 x 128 tile for every shape, 0 restores the default choice */
void rorl_gemm_force_bn(int bn) { g_gemm_bn = bn; }
/* diagnostic / A-B benchmarks only: 16 or 32 forces the k-depth of a stage for the TN variant, 0 restores the default */
void rorl_gemm_force_bk(int bk) { g_gemm_bk = bk; }

// Split-K factor rorl_gemm_nt uses for a [M x N] output reduced over R rows in G batches (the caller sizes D with it).
int rorl_gemm_nt_splits(int64_t M, int64_t N, int64_t R, int64_t G) {
    if (M <= 0 || N <= 0 || R <= 0 || G <= 0) return RORL_ERR_SHAPE;
    const long long tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + kGemmBN - 1) / kGemmBN) * G;
    const long long kt = (R + kGemmBK - 1) / kGemmBK;
    long long s = (2 * 148 + tiles - 1) / tiles;
    const long long smax = kt / 4 > 0 ? kt / 4 : 1;
    if (s > smax) s = smax;
    if (s < 1) s = 1;
    return (int)s;
}

// D[s][g][M, N] = sum over the s-th slice of rows r of A[g][r, M]^T B[g][r, N]   (both operands MN-major: the
// reduction runs over ROWS).  Weight gradients: dW[n, k] = sum_m dY[m, n] X[m, k] with A = dY, B = X.
// splits must equal rorl_gemm_nt_splits(M, N, R, G); the caller sums the partials over s.
int rorl_gemm_nt(const float* A, const float* B, float* D, int64_t M, int64_t N, int64_t R, int64_t G, int64_t lda,
                 int64_t ldb, int64_t ldd, int64_t strideA, int64_t strideB, int64_t strideD, int64_t splits,
                 int64_t strideSplit, int passes, cudaStream_t stream) {
    if (!A || !B || !D) return RORL_ERR_ARG;
    if (M <= 0 || N <= 0 || R <= 0 || G <= 0 || splits <= 0) return RORL_ERR_SHAPE;
    if (M % 4 || N % 4 || lda % 4 || ldb % 4 || ldd % 4 || strideA % 4 || strideB % 4 || strideD % 4 || strideSplit % 4)
        return RORL_ERR_ALIGN;
    if ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(D)) & 15)
        return RORL_ERR_ALIGN;
    if (passes < 1 || passes > 4) return RORL_ERR_ARG;
    if (splits != rorl_gemm_nt_splits(M, N, R, G)) return RORL_ERR_ARG;
    if (passes == 2 || passes == 4)
        return gemm_nt_bf16x3(A, B, D, M, N, R, G, lda, ldb, ldd, strideA, strideB, strideD, splits, strideSplit, passes == 4, stream);
    CUtensorMap mapA, mapB;
    int rc = make_map(&mapA, A, R, M, lda, strideA ? G : 1, strideA, kGemmBK);
    if (rc) return rc;
    rc = make_map(&mapB, B, R, N, ldb, strideB ? G : 1, strideB, kGemmBK);
    if (rc) return rc;
    GemmParams p;
    p.D = D; p.Dpre = nullptr; p.bias = nullptr; p.M = (int)M; p.N = (int)N; p.K = (int)R; p.G = (int)G;
    p.ldd = ldd; p.strideD = strideD; p.strideBias = 0;
    p.a_batched = strideA != 0; p.b_batched = strideB != 0; p.act = 0; p.passes = passes; p.reduce_g = 0;
    p.splits = (int)splits; p.strideSplit = strideSplit; p.dbg = g_gemm_dbg;
    const int sms = gemm_sms();
    const long long tiles = ((M + kGemmBM - 1) / kGemmBM) * ((N + kGemmBN - 1) / kGemmBN) * G * splits;
    const int grid = (int)(tiles < sms ? tiles : sms);
    gemm_kernel<true, 128, 32><<<grid, kGemmThreads, GemmCfg<128, 32>::kSmem, stream>>>(mapA, mapB, p);
    RORL_RETURN_LAUNCH();
}

}  // extern "C"
