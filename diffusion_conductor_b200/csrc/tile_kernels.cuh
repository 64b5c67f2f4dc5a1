// tcgen05 / TMEM tile kernels.
//
// One CTA owns a tile of 128 tokens (= 128 TMEM lanes).  Warp roles (18 warps, 576 threads):
//   warps 0-15 : "row" threads.  Warp w works on TMEM lanes 32*(w%4).. (rows of the tile) and on
//                the column quarter cq = w/4 (32 of the 128 features, i.e. two attention heads).
//                They read accumulators from TMEM, do every row-wise op (LayerNorm, softmax over
//                head-dim, q.A, FiLM, SiLU, GELU) in registers and write the next GEMM's A operand
//                to shared memory (K-major SW128).  The four warps that share a row exchange
//                LayerNorm partial statistics through shared memory + a 128-thread named barrier.
//   warp 16    : producer -- streams packed operand blocks global/L2 -> smem ring with cp.async.bulk
//   warp 17    : MMA issuer (one elected lane) + TMEM allocator
//   warps 18,19: second producer (dependent-GEMM operands) and, in the follower CTA of a pair, the relay that
//                forwards "stage landed" events to the leader
// Synchronisation in the main loop is mbarrier-only between roles: full/empty per ring stage,
// a_ready (row threads -> MMA), d_ready per accumulator (tcgen05.commit -> row threads).
#pragma once
#include "simple_kernels.cuh"

namespace dc {

constexpr int kStages = 3;
constexpr int kStageABytes = kABlockBytes;        // 16 KB: one [128 x 64] A block (streamed operand)
constexpr int kStageWBytes = 32 * 1024;           // up to [256 x 64] weight block, or a whole small weight
constexpr int kStageBytes = kStageABytes + kStageWBytes;
constexpr int kAworkBytes = 2 * kABlockBytes;     // [128 x 128] A operand written by the row threads
constexpr int kRowWarps = 16;
constexpr int kRowThreads = kRowWarps * 32;       // 512
constexpr int kProducerWarp = 16, kMmaWarp = 17, kProducerBWarp = 18, kRelayWarp = 19;
constexpr int kTileThreads = kRowThreads + 128;   // 640

// TMEM column map (512 columns x 128 lanes x fp32)
constexpr uint32_t kColH = 0;      // residual stream h            [128]
constexpr uint32_t kColS = 128;    // FiLM scale|shift accumulator  [256]  (also q|k of the next SA)
constexpr uint32_t kColW = 384;    // work accumulator              [128]

struct TileBarriers {
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t a_ready;
    uint64_t d_ready[3];   // 0: S, 1: H, 2: W
    uint32_t tmem_base;
};

// One GEMM of the static per-launch schedule.
struct TileOp {
    uint32_t w_off;          // byte offset of stage 0's weight bytes in the packed weight buffer
    uint32_t w_stage_bytes;  // weight bytes per ring stage
    uint16_t n_stages;       // ring stages this op consumes
    uint16_t kb_per_stage;   // k-blocks (64) per stage
    uint16_t n;              // UMMA N
    uint16_t d_col;          // accumulator column
    uint8_t a_from_ring;     // 1: A block streamed with the weights (A_emb / Z), 0: A = row-thread operand buffer
    uint8_t accumulate;      // 1: D += (residual add into h)
    uint8_t wait_a;          // 1: wait for a_ready before issuing
    uint8_t commit;          // 0/1/2: arrive d_ready[commit] when done, 255: none
};

constexpr int kMaxOps = 12;

// ---------------------------------------------------------------------------------------------
// shared pieces
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tile_setup(TileBarriers* bars, int warp, int lane) {
    if (warp == kProducerWarp && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(smem_u32(&bars->full[i]), 1);
            mbar_init(smem_u32(&bars->empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->a_ready), kRowThreads);
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars->d_ready[i]), 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

__device__ __forceinline__ void tile_teardown(TileBarriers* bars, int warp) {
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(bars->tmem_base, 512);
    }
}

// producer lane: walk the op list, one ring stage at a time
__device__ __forceinline__ void producer_loop(const TileOp* ops, int n_ops, const uint8_t* wbuf, const uint8_t* a_img_tile,
                                              uint8_t* ring, TileBarriers* bars) {
    uint32_t it = 0;
    for (int o = 0; o < n_ops; ++o) {
        const TileOp op = ops[o];
        for (int s = 0; s < op.n_stages; ++s, ++it) {
            const uint32_t st = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(smem_u32(&bars->empty[st]), ph ^ 1u);
            const uint32_t full = smem_u32(&bars->full[st]);
            uint8_t* stage = ring + st * kStageBytes;
            const uint32_t bytes = op.w_stage_bytes + (op.a_from_ring ? kStageABytes : 0);
            mbar_arrive_expect_tx(full, bytes);
            if (op.a_from_ring) bulk_g2s(smem_u32(stage), a_img_tile + (size_t)s * kStageABytes, kStageABytes, full);
            bulk_g2s(smem_u32(stage + kStageABytes), wbuf + op.w_off + (size_t)s * op.w_stage_bytes, op.w_stage_bytes, full);
        }
    }
}

// MMA lane: same walk; A comes from the ring or from the row threads' operand buffer
template <bool kBf16>
__device__ __forceinline__ void mma_loop(const TileOp* ops, int n_ops, uint8_t* ring, uint8_t* awork, TileBarriers* bars,
                                         uint32_t tmem_base) {
    uint32_t it = 0, a_phase = 0;
    for (int o = 0; o < n_ops; ++o) {
        const TileOp op = ops[o];
        const uint32_t idesc = make_idesc<kBf16>(kTileRows, op.n);
        if (op.wait_a) {
            mbar_wait(smem_u32(&bars->a_ready), a_phase & 1u);
            ++a_phase;
            tc_fence_after();
        }
        for (int s = 0; s < op.n_stages; ++s, ++it) {
            const uint32_t st = it % kStages, ph = (it / kStages) & 1u;
            mbar_wait(smem_u32(&bars->full[st]), ph);
            tc_fence_after();
            const uint32_t stage = smem_u32(ring + st * kStageBytes);
            for (int kb = 0; kb < op.kb_per_stage; ++kb) {
                const uint32_t a_addr = op.a_from_ring ? stage : smem_u32(awork) + kb * kABlockBytes;
                const uint32_t b_addr = stage + kStageABytes + kb * (uint32_t)op.n * 128u;
                umma_kblock(tmem_base + op.d_col, a_addr, b_addr, idesc, op.accumulate || s > 0 || kb > 0);
            }
            umma_commit(smem_u32(&bars->empty[st]));
        }
        if (op.commit != 255) umma_commit(smem_u32(&bars->d_ready[op.commit]));
    }
}

// ---------------------------------------------------------------------------------------------
// Generic row GEMM:  out[M][ldo] (cols [0,N)) = A_img[M][K] . W[N][K]^T + bias,  N <= 256.
// A_img and W are packed 16-bit SW128 images; out is fp32.  Used for the step-invariant
// cross-attention K/V projections and as the tcgen05 self-test.
// ---------------------------------------------------------------------------------------------
struct GemmRowsArgs {
    const uint8_t* a_img;   // [tiles][kblocks][16 KB]
    const uint8_t* w_img;   // [kblocks][N x 128 B]
    const float* bias;      // [N] or null
    float* out;
    int M, N, kblocks, ldo;
    int blocked;            // 1: out is in the blocked activation layout with N columns (ldo ignored)
    // blockIdx.y = layer (the step-invariant cross-attention K | V projections of several layers in one launch): strides between layers
    size_t w_layer_stride;  // bytes
    size_t bias_layer_stride, out_layer_stride;   // floats
};

template <bool kBf16>
__global__ void __launch_bounds__(kTileThreads, 1) gemm_rows_kernel(const __grid_constant__ GemmRowsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = smem;
    TileBarriers* bars = reinterpret_cast<TileBarriers*>(smem + kStages * kStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_trigger();
    tile_setup(bars, warp, lane);
    pdl_wait();
    const uint32_t tmem_base = bars->tmem_base;

    __shared__ TileOp op;
    if (threadIdx.x == 0) {
        op.w_off = 0;
        op.w_stage_bytes = (uint32_t)a.N * 128u;
        op.n_stages = (uint16_t)a.kblocks;
        op.kb_per_stage = 1;
        op.n = (uint16_t)a.N;
        op.d_col = 0;
        op.a_from_ring = 1;
        op.accumulate = 0;
        op.wait_a = 0;
        op.commit = 0;
    }
    __syncthreads();

    if (warp == kProducerWarp) {
        if (lane == 0)
            producer_loop(&op, 1, a.w_img + (size_t)blockIdx.y * a.w_layer_stride, a.a_img + (size_t)blockIdx.x * a.kblocks * kStageABytes, ring, bars);
    } else if (warp == kMmaWarp) {
        if (lane == 0) mma_loop<kBf16>(&op, 1, ring, nullptr, bars, tmem_base);
    } else {
        const int lq = warp & 3, cq = warp >> 2;
        const long g = (long)blockIdx.x * kTileRows + lq * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16);
        float* outp = a.out + (size_t)blockIdx.y * a.out_layer_stride;
        const float* biasp = a.bias ? a.bias + (size_t)blockIdx.y * a.bias_layer_stride : nullptr;
        mbar_wait(smem_u32(&bars->d_ready[0]), 0);
        tc_fence_after();
        // column quarter cq covers columns [64 cq, 64 cq + 64) in 16-column pieces
        for (int c = 64 * cq; c < min(a.N, 64 * cq + 64); c += 16) {
            float v[16];
            tmem_ld16(trow + c, v);
            tmem_wait_ld();
            if (g < a.M) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4* dst = reinterpret_cast<float4*>(a.blocked ? outp + blk_index(g, c + 4 * i, a.N) : outp + (size_t)g * a.ldo + c + 4 * i);
                    float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (biasp) {
                        o.x += biasp[c + 4 * i];
                        o.y += biasp[c + 4 * i + 1];
                        o.z += biasp[c + 4 * i + 2];
                        o.w += biasp[c + 4 * i + 3];
                    }
                    *dst = o;
                }
            }
        }
    }
    tile_teardown(bars, warp);
}

// ---------------------------------------------------------------------------------------------
// Decoder-layer kernel.
//
// Per-layer fp32 parameter block (floats).  LayerNorm affines that feed a Linear directly (sa.norm
// -> q/k/v, ca.norm -> q) are folded into the packed weights / these biases at load time.
// ---------------------------------------------------------------------------------------------
constexpr int kPrmSaBq = 0, kPrmSaBk = 128, kPrmSaBv = 256;
constexpr int kPrmStSa = 384;                      // stylization block: BE[256] G[128] B[128] BO[128]
constexpr int kPrmCaBq = 1024;
constexpr int kPrmStCa = 1152;
constexpr int kPrmFfB1 = 1792, kPrmFfB2 = 1856;
constexpr int kPrmStFf = 1984;
constexpr int kPrmFloats = 2624;
constexpr int kStBe = 0, kStG = 256, kStB = 384, kStBo = 512;

// Two operand streams feed the MMA issuer:
//   ring A : the FiLM projections S = A_emb . We (K = 512, N = 256), 8 stages of (16 KB A_emb k-block +
//            32 KB weight k-block) each.  They depend on no row-thread result, only on the S accumulator
//            being free, so they fill every gap of the tensor pipe.
//   ring B : the dependent GEMMs (<= 32 KB of weights each, A operand written by the row threads) and the
//            per-clip block-diagonal attention matrices.
// The issuer polls both (dependent work first), so a short dependent GEMM never queues behind a long
// FiLM projection.
constexpr int kRingAStages = 3;                 // single-CTA mode: 3 x (16 KB A_emb + 32 KB weights)
constexpr int kRingBStages = 2;
constexpr int kRingBStageBytes = 16 * 1024;     // one k-block (64) of a dependent GEMM's B operand
constexpr int kSopStages = 8;
// pair mode: each CTA holds half of every B operand, so the same shared memory gives deeper rings, which the
// longer leader<->follower signalling round trip needs
constexpr int kPairRingAStages = 4;             // 4 x (16 KB A_emb + 16 KB weights)
constexpr int kPairStageBytes = kStageABytes + kStageWBytes / 2;
constexpr int kPairRingBStages = 4;             // 4 x 8 KB
constexpr int kPairRingBStageBytes = 8 * 1024;

struct DOp {
    uint32_t w_off;        // byte offset in the packed weight buffer (seg == 0)
    uint32_t w_bytes;      // bytes of the whole B operand (kb ring-B stages of w_bytes / kb each)
    uint16_t n;            // UMMA N
    uint16_t d_col;        // accumulator column
    uint8_t kb;            // k-blocks of 64
    uint8_t accumulate;    // D += (residual add into h)
    uint8_t wait;          // 0: none, 1: a_ready (row threads), 2: q_full (bulk copy of the q image)
    uint8_t commit;        // 0/1/2: arrive d_ready[commit] when done, 255: none
    uint8_t seg;           // 0: plain; 1 / 2: one lane-masked GEMM per clip segment of the tile with that clip's
                           //    block-diagonal self- / cross-attention matrix as B operand
    uint8_t releases_s;    // passing this op's wait means the row threads are done with the S accumulator
    uint8_t ring_a;        // 1: weights travel through ring A (idle once the FiLM projections are done) instead of ring B
    uint8_t pad;
};

constexpr int kMaxDOps = 13;
constexpr int kKvPartFloats = 128 + 128 + kH * 256;
constexpr int kMaxSOps = 3;

struct LayerBarriers {
    uint64_t fullA[4], emptyA[4];
    uint64_t fullB[4], emptyB[4];
    uint64_t a_ready;
    uint64_t s_free;       // row threads -> FiLM-projection issuer: the S accumulator has been consumed
    uint64_t q_full;
    uint64_t aemb_ready;   // persistent kernel: this tile's A_emb image has been written (row threads -> producer)
    uint64_t d_ready[3];   // 0: S, 1: H, 2: W
    uint32_t tmem_base;
};

struct LayerArgs {
    uint32_t sop_w_off[kMaxSOps];
    int n_s;
    DOp dops[kMaxDOps];
    int n_d;
    int do_main;            // SA tail (q.A + stylization), cross-attention, FFN of layer l
    int do_sa1;             // LayerNorm + q/k/v projections of layer l+1
    int M, T;
    int mask_invert;        // debug: flip the meaning of the MMA lane mask
    const uint8_t* wbuf;    // packed weights (whole model)
    const uint8_t* aemb;    // A_emb image [tiles][8][16 KB]
    const float* prm;       // parameter block of layer l (do_main)
    const float* prm_next;  // parameter block of layer l+1 (do_sa1; only the SA biases are read)
    float* h;               // [Mpad][128] residual stream (in/out)
    uint8_t* q_img;         // [tiles][32 KB] softmax_hd(Q) of the self-attention as a packed A-operand image (in/out)
    float* kv;              // [Mpad][256] k | v of the self-attention (out)
    const uint8_t* bd_sa;   // [B][32 KB] block-diagonal softmax_T(K)^T V of layer l self-attention
    const uint8_t* bd_ca;   // cross-attention counterpart of layer l; clip stride bd_ca_stride bytes
    size_t bd_ca_stride;
    const long long* length;  // [B] or null (all frames valid)
    int fuse_kv;            // 1 (T >= 128): time-axis softmax + K^T V fused into this kernel's epilogue
    float* kv_part;         // [tiles][2][kKvPartFloats] per (tile, clip segment): max[128] | sum[128] | K^T V [8][16][16]
    int* clip_cnt;          // [B] arrival counters (zero between launches)
    uint8_t* bd_sa_out;     // == bd_sa; written for the NEXT layer by the CTA that completes a clip
    unsigned long long* timeline;   // debug: [0] = event count, then (clock64, id) pairs of CTA 0; null = off
};

// debug timeline of CTA 0 (one writer lane per role)
__device__ __forceinline__ void tl_mark(const LayerArgs& a, unsigned long long id) {
    if (a.timeline != nullptr && blockIdx.x == 0) {
        const unsigned long long slot = atomicAdd(a.timeline, 1ull);
        if (slot < 254) {
            a.timeline[1 + 2 * slot] = (unsigned long long)clock64();
            a.timeline[2 + 2 * slot] = id;
        }
    }
}

template <bool kFast>
__device__ __forceinline__ float silu_f(float v) {
    if constexpr (kFast) {      // x * sigmoid(x) = x * (0.5 + 0.5 tanh(x/2)) : one MUFU
        float th;
        asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * v));
        return v * fmaf(0.5f, th, 0.5f);
    } else {
        return __fdividef(v, 1.f + __expf(-v));
    }
}
__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }

// MUFU.EX2 without the denormal fix-up code exp2f() carries (results below 2^-126 flush to 0)
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// erf-GELU (reference activation="gelu", transformer.py:164/172) for a pair of values in packed f32x2 math:
//   erf(t) = 1 - 2^(t P(t)) on t = min(|x| / sqrt 2, 4), P a degree-6 polynomial fitted to log2(erfc(t)) / t
//   (max abs error of erf 3.5e-7, of GELU 1e-7 -- at the fp32 rounding level);  x (1 + erf) / 2 = x/2 + |x|/2 * erf(|t|).
__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
    const float t0 = fminf(fabsf(x0) * 0.70710678118654752f, 4.f), t1 = fminf(fabsf(x1) * 0.70710678118654752f, 4.f);
    const uint64_t t = pk2(t0, t1);
    uint64_t p = pk2(3.362845746e-05f, 3.362845746e-05f);
    p = ffma2(p, t, pk2(-3.501975880e-05f, -3.501975880e-05f));
    p = ffma2(p, t, pk2(-3.336299676e-03f, -3.336299676e-03f));
    p = ffma2(p, t, pk2(3.064695559e-02f, 3.064695559e-02f));
    p = ffma2(p, t, pk2(-1.496411115e-01f, -1.496411115e-01f));
    p = ffma2(p, t, pk2(-9.181558490e-01f, -9.181558490e-01f));
    p = ffma2(p, t, pk2(-1.627928376e+00f, -1.627928376e+00f));
    float q0, q1;
    upk2(fmul2(p, t), q0, q1);
    const uint64_t e = ffma2(pk2(ex2_ftz(q0), ex2_ftz(q1)), pk2(-1.f, -1.f), pk2(1.f, 1.f));      // erf(|t|)
    const uint64_t hx = fmul2(pk2(x0, x1), pk2(0.5f, 0.5f)), ha = fmul2(pk2(fabsf(x0), fabsf(x1)), pk2(0.5f, 0.5f));
    upk2(ffma2(ha, e, hx), x0, x1);
}

__device__ __forceinline__ void softmax16(float* q) {
    float mx = q[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) mx = fmaxf(mx, q[i]);
    // exp(x - mx) = 2^(x*log2e - mx*log2e): one packed FMA per pair, then MUFU.EX2
    const uint64_t l2 = pk2(1.4426950408889634f, 1.4426950408889634f);
    const uint64_t nm = pk2(-mx * 1.4426950408889634f, -mx * 1.4426950408889634f);
    uint64_t acc = pk2(0.f, 0.f);
    uint64_t e2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float a0, a1;
        upk2(ffma2(pk2(q[2 * i], q[2 * i + 1]), l2, nm), a0, a1);
        e2[i] = pk2(ex2_ftz(a0), ex2_ftz(a1));
        acc = fadd2(acc, e2[i]);
    }
    float s0, s1;
    upk2(acc, s0, s1);
    const float inv = __fdividef(1.f, s0 + s1);
    const uint64_t inv2 = pk2(inv, inv);
#pragma unroll
    for (int i = 0; i < 8; ++i) upk2(fmul2(e2[i], inv2), q[2 * i], q[2 * i + 1]);
}

// Row statistics over 128 features held as 4 x 32 registers by the 4 warps that share a row:
// local (mean, M2) -> shared memory -> 128-thread named barrier -> Chan combine.
struct RowStats {
    float2* xchg;      // [4][128]
    uint32_t bar_id;   // 1 + lq
    uint32_t r;        // row within the tile
    uint32_t cq;
    uint32_t flip;
};

// kGuardReuse: second barrier that keeps a fast warp's NEXT call from overwriting the exchange buffer before everyone has read
// it.  The persistent kernel passes false: between two calls every warp publishes an A operand and waits for the MMA that needed
// all 16 publications, which orders every read of call k before any write of call k + 1.
template <bool kGuardReuse = true>
__device__ __forceinline__ void row_stats32(RowStats& rs, const float* v, float& mean, float& rstd) {
    uint64_t s2a = pk2(0.f, 0.f), s2b = s2a;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s2a = fadd2(s2a, pk2(v[4 * i], v[4 * i + 1]));
        s2b = fadd2(s2b, pk2(v[4 * i + 2], v[4 * i + 3]));
    }
    float sa, sb;
    upk2(fadd2(s2a, s2b), sa, sb);
    const float lm = (sa + sb) * (1.f / 32.f);
    const uint64_t nlm = pk2(-lm, -lm);
    uint64_t qa = pk2(0.f, 0.f), qb = qa;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint64_t d0 = fadd2(pk2(v[4 * i], v[4 * i + 1]), nlm), d1 = fadd2(pk2(v[4 * i + 2], v[4 * i + 3]), nlm);
        qa = ffma2(d0, d0, qa);
        qb = ffma2(d1, d1, qb);
    }
    float m2a, m2b;
    upk2(fadd2(qa, qb), m2a, m2b);
    const float m2 = m2a + m2b;
    float2* buf = rs.xchg;
    buf[rs.cq * 128 + rs.r] = make_float2(lm, m2);
    named_bar_sync(rs.bar_id, 128);
    const float2 p0 = buf[rs.r], p1 = buf[128 + rs.r], p2 = buf[256 + rs.r], p3 = buf[384 + rs.r];
    if constexpr (kGuardReuse) named_bar_sync(rs.bar_id, 128);       // single exchange buffer: everyone has read before the next call writes
    mean = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
    const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
    const float M2 = (p0.y + p1.y) + (p2.y + p3.y) + 32.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    rstd = rsqrtf(M2 * (1.f / kD) + kLnEps);
}

// v <- (v - mean) * rstd for 32 values (LayerNorm without affine)
__device__ __forceinline__ void normalize32(float* v, float mean, float rstd) {
    const uint64_t r2 = pk2(rstd, rstd), nm = pk2(-mean * rstd, -mean * rstd);
#pragma unroll
    for (int i = 0; i < 16; ++i) upk2(ffma2(pk2(v[2 * i], v[2 * i + 1]), r2, nm), v[2 * i], v[2 * i + 1]);
}

// A operand <- SiLU( LN(y) * (1 + scale) + shift ) for this thread's 32 features [c0, c0+32)
// (reference transformer.py:77-80).  The scale|shift accumulator in TMEM columns kColS is laid out
//   [scale 0..63 | shift 0..63 | scale 64..127 | shift 64..127]; st = stylization params in smem.
// s_free_addr != 0: arrive on that mbarrier (one lane per warp) as soon as this warp's LAST scale|shift values have left
// TMEM -- before the second half of the math -- so that the next FiLM projection can start that much earlier.
template <bool kBf16>
__device__ __forceinline__ void film_to_a(uint32_t trow, const float* y, float mean, float rstd, const float* st, uint32_t awork,
                                          uint32_t r, uint32_t c0, uint32_t s_free_addr = 0, int lane = 0) {
    const uint32_t sbase = trow + kColS + (c0 >> 6) * 128 + (c0 & 63);     // scale column of feature c0
    const float* be = st + kStBe + (c0 >> 6) * 128 + (c0 & 63);
    const uint64_t r2 = pk2(rstd, rstd), nm = pk2(-mean * rstd, -mean * rstd), half2 = pk2(0.5f, 0.5f);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        float sc[16], sh[16], o[16];
        tmem_ld16(sbase + 16 * hf, sc);
        tmem_ld16(sbase + 64 + 16 * hf, sh);
        tmem_wait_ld();
        if (hf == 1 && s_free_addr != 0) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free_addr);
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const float4 g4 = *reinterpret_cast<const float4*>(st + kStG + c0 + 16 * hf + 4 * i4);
            const float4 b4 = *reinterpret_cast<const float4*>(st + kStB + c0 + 16 * hf + 4 * i4);
            const float4 es = *reinterpret_cast<const float4*>(be + 16 * hf + 4 * i4);
            const float4 eh = *reinterpret_cast<const float4*>(be + 64 + 16 * hf + 4 * i4);
            const uint64_t gg[2] = {pk2(g4.x, g4.y), pk2(g4.z, g4.w)}, bb[2] = {pk2(b4.x, b4.y), pk2(b4.z, b4.w)};
            const uint64_t ss[2] = {pk2(es.x, es.y), pk2(es.z, es.w)}, hh[2] = {pk2(eh.x, eh.y), pk2(eh.z, eh.w)};
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int i = 4 * i4 + 2 * k;
                const uint64_t z = ffma2(pk2(y[16 * hf + i], y[16 * hf + i + 1]), r2, nm);          // (y - mean) * rstd
                const uint64_t n = ffma2(z, gg[k], bb[k]);                                          // LN affine
                const uint64_t sca = fadd2(pk2(sc[i], sc[i + 1]), ss[k]), shf = fadd2(pk2(sh[i], sh[i + 1]), hh[k]);
                const uint64_t u = ffma2(n, sca, shf);                                              // FiLM
                float u0, u1;
                if constexpr (kBf16) {        // SiLU(u) = u * (0.5 + 0.5 tanh(u / 2)): one MUFU per element
                    upk2(fmul2(u, half2), u0, u1);
                    float t0, t1;
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
                    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
                    upk2(fmul2(u, ffma2(pk2(t0, t1), half2, half2)), o[i], o[i + 1]);
                } else {
                    upk2(u, u0, u1);
                    o[i] = __fdividef(u0, 1.f + __expf(-u0));
                    o[i + 1] = __fdividef(u1, 1.f + __expf(-u1));
                }
            }
        }
        store_a16<kBf16>(awork, r, c0 + 16 * hf, o);
    }
}

// v[32] += bias[c0 .. c0+32) from shared memory
__device__ __forceinline__ void add_bias32(float* v, const float* bias) {
#pragma unroll
    for (int i4 = 0; i4 < 8; ++i4) {
        const float4 b4 = *reinterpret_cast<const float4*>(bias + 4 * i4);
        upk2(fadd2(pk2(v[4 * i4], v[4 * i4 + 1]), pk2(b4.x, b4.y)), v[4 * i4], v[4 * i4 + 1]);
        upk2(fadd2(pk2(v[4 * i4 + 2], v[4 * i4 + 3]), pk2(b4.z, b4.w)), v[4 * i4 + 2], v[4 * i4 + 3]);
    }
}

// row threads signal "A operand (and any TMEM writes) ready": one elected arrive per warp on the leader CTA's barrier
// kRemote: the barrier lives in the pair leader's shared memory (release.cluster arrive: costs a memory barrier).
// Within one CTA the plain arrive (release.cta) is enough: the operand image was fenced for the async proxy above.
template <bool kRemote>
__device__ __forceinline__ void bar_arrive(uint32_t addr) {
    if constexpr (kRemote) mbar_arrive_cluster(addr);
    else mbar_arrive(addr);
}
template <bool kRemote = true>
__device__ __forceinline__ void rows_publish(uint32_t a_ready_addr, int lane) {
    fence_async_smem();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) bar_arrive<kRemote>(a_ready_addr);
}
__device__ __forceinline__ void rows_wait(LayerBarriers* bars, int which, uint32_t& phase) {
    mbar_wait(smem_u32(&bars->d_ready[which]), phase & 1u);
    ++phase;
    tc_fence_after();
}

// Clip segments of the rows one MMA covers (128 rows of a tile, or 256 rows of a CTA pair): rows [lo, hi)
// belong to clip first_clip + s.
struct RowSegs {
    int first_clip, n_seg, row0, T, rows;
    __device__ __forceinline__ void init(int first_tile, int rows_, int M, int T_) {
        T = T_;
        rows = rows_;
        row0 = first_tile * kTileRows;
        first_clip = row0 / T;
        const int last_row = min(row0 + rows - 1, M - 1);
        n_seg = last_row / T - first_clip + 1;
    }
    // bit r set <=> row r is NOT in segment s (the tcgen05 "disable output lane" convention); 8 words = 256 rows
    __device__ __forceinline__ void mask(int s, uint32_t* m, bool invert) const {
        const int lo = max(0, (first_clip + s) * T - row0);
        const int hi = (s == n_seg - 1) ? rows : min(rows, (first_clip + s + 1) * T - row0);
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            uint32_t in = 0;
            const int a = max(lo, 32 * w), b = min(hi, 32 * w + 32);
            if (b > a) in = ((b - a) == 32 ? 0xFFFFFFFFu : ((1u << (b - a)) - 1u)) << (a - 32 * w);
            m[w] = invert ? in : ~in;
        }
    }
};

// kPair: two CTAs of a cluster (same TPC) run one tile each and share every MMA (cta_group::2, M = 256): each CTA
// supplies its own 128 A rows and HALF of the B operand (N/2 rows), which halves the shared-memory operand
// bandwidth and the weight bytes each SM pulls from L2.  The leader CTA (rank 0) issues all MMAs; completion is
// multicast to both CTAs' barriers; the follower relays its "stage landed" events to the leader.
template <bool kBf16, bool kPair>
__global__ void __launch_bounds__(kTileThreads, 1) layer_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    constexpr int kNA = kPair ? kPairRingAStages : kRingAStages, kSA = kPair ? kPairStageBytes : kStageBytes;
    constexpr int kNB = kPair ? kPairRingBStages : kRingBStages, kSB = kPair ? kPairRingBStageBytes : kRingBStageBytes;
    static_assert(kNA * kSA <= kRingAStages * kStageBytes && kNB * kSB <= kRingBStages * kRingBStageBytes,
                  "pair-mode rings must fit the single-CTA ring footprint");
    uint8_t* ringA = smem;
    uint8_t* ringB = smem + kRingAStages * kStageBytes;          // same carve-up in both modes (fused-kv scratch relies on it)
    uint8_t* awork_p = ringB + kRingBStages * kRingBStageBytes;
    float* prm = reinterpret_cast<float*>(awork_p + kAworkBytes);          // [kPrmFloats]
    float* prm_sa = prm + kPrmFloats;                                      // [384] SA biases of layer l+1
    float2* xchg = reinterpret_cast<float2*>(prm_sa + 384);                // [4][128]
    LayerBarriers* bars = reinterpret_cast<LayerBarriers*>(xchg + 512);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = kPair ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    constexpr uint32_t kHalf = kPair ? 2u : 1u;     // each CTA holds 1/kHalf of every B operand
    const uint32_t brank = (a.mask_invert == 3) ? (rank ^ 1u) : rank;   // which half of B this CTA holds (debug: swapped)

    pdl_trigger();
    // prologue that touches only weights / on-chip state overlaps the previous kernel's tail
    if (a.do_main)
        for (int i = threadIdx.x; i < kPrmFloats; i += kTileThreads) prm[i] = a.prm[i];
    if (a.do_sa1)
        for (int i = threadIdx.x; i < 384; i += kTileThreads) prm_sa[i] = a.prm_next[i];
    if (warp == kProducerWarp && lane == 0) {
        // "full" barriers of the leader also collect the follower's relay arrival
        const uint32_t nfull = (kPair && leader) ? 2u : 1u;
        for (int i = 0; i < kNA; ++i) mbar_init(smem_u32(&bars->fullA[i]), nfull), mbar_init(smem_u32(&bars->emptyA[i]), 1);
        for (int i = 0; i < kNB; ++i) mbar_init(smem_u32(&bars->fullB[i]), nfull), mbar_init(smem_u32(&bars->emptyB[i]), 1);
        mbar_init(smem_u32(&bars->a_ready), kRowWarps * kHalf);
        mbar_init(smem_u32(&bars->s_free), kRowWarps * kHalf);
        mbar_init(smem_u32(&bars->q_full), nfull);
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars->d_ready[i]), 1);
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        if constexpr (kPair) {
            tmem_alloc2(smem_u32(&bars->tmem_base), 512);
            tmem_relinquish2();
        } else {
            tmem_alloc(smem_u32(&bars->tmem_base), 512);
            tmem_relinquish();
        }
    }
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();
    if (threadIdx.x == 0) tl_mark(a, 1);

    // segments of the rows covered by one MMA
    RowSegs segs;
    segs.init(kPair ? (int)(blockIdx.x & ~1u) : (int)blockIdx.x, kPair ? 2 * kTileRows : kTileRows, a.M, a.T);

    if (warp == kProducerWarp) {
        // ---------------- ring A: A_emb k-blocks + this CTA's share of the FiLM projection weights
        if (lane == 0) {
            const uint8_t* a_img = a.aemb + (size_t)blockIdx.x * 8 * kStageABytes;
            constexpr uint32_t kWShare = kStageWBytes / kHalf;
            uint32_t it = 0;
            for (int o = 0; o < a.n_s; ++o) {
                const uint8_t* w = a.wbuf + a.sop_w_off[o] + brank * kWShare;
                for (int s = 0; s < kSopStages; ++s, ++it) {
                    const uint32_t st = it % kNA, ph = (it / kNA) & 1u;
                    mbar_wait(smem_u32(&bars->emptyA[st]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bars->fullA[st]);
                    uint8_t* stage = ringA + st * kSA;
                    mbar_arrive_expect_tx(full, kStageABytes + kWShare);
                    bulk_g2s_chunked(smem_u32(stage), a_img + (size_t)s * kStageABytes, kStageABytes, full);
                    bulk_g2s_chunked(smem_u32(stage + kStageABytes), w + (size_t)s * kStageWBytes, kWShare, full);
                }
            }
            for (int o = 0; o < a.n_d; ++o) {
                const DOp op = a.dops[o];
                if (!op.ring_a) continue;
                const uint32_t st = it % kNA, ph = (it / kNA) & 1u;
                ++it;
                mbar_wait(smem_u32(&bars->emptyA[st]), ph ^ 1u);
                const uint32_t full = smem_u32(&bars->fullA[st]);
                const uint32_t kb_bytes = op.w_bytes / op.kb, share = kb_bytes / kHalf;
                mbar_arrive_expect_tx(full, share * op.kb);
                for (int kb = 0; kb < op.kb; ++kb)
                    bulk_g2s_chunked(smem_u32(ringA + st * kSA + kStageABytes + kb * share),
                             a.wbuf + op.w_off + (size_t)kb * kb_bytes + brank * share, share, full);
            }
        }
    } else if (warp == kProducerBWarp) {
        // ---------------- ring B: dependent-GEMM weights, per-clip attention matrices, the q image
        if (lane == 0) {
            if (a.do_main) {
                const uint32_t qf = smem_u32(&bars->q_full);
                mbar_arrive_expect_tx(qf, kAworkBytes);
                bulk_g2s_chunked(smem_u32(awork_p), a.q_img + (size_t)blockIdx.x * kAworkBytes, kAworkBytes, qf);
            }
            uint32_t it = 0;
            for (int o = 0; o < a.n_d; ++o) {
                const DOp op = a.dops[o];
                if (op.ring_a || op.seg == 3) continue;       // seg 3: operands come from the row threads, not from memory
                const int n_st = op.seg ? segs.n_seg : 1;
                const uint32_t kb_bytes = op.w_bytes / op.kb, share = kb_bytes / kHalf;
                for (int s = 0; s < n_st; ++s) {
                    const uint8_t* src = op.seg == 0   ? a.wbuf + op.w_off
                                         : op.seg == 1 ? a.bd_sa + (size_t)(segs.first_clip + s) * kAworkBytes
                                                       : a.bd_ca + (size_t)(segs.first_clip + s) * a.bd_ca_stride;
                    for (int kb = 0; kb < op.kb; ++kb, ++it) {
                        const uint32_t st = it % kNB, ph = (it / kNB) & 1u;
                        mbar_wait(smem_u32(&bars->emptyB[st]), ph ^ 1u);
                        const uint32_t full = smem_u32(&bars->fullB[st]);
                        mbar_arrive_expect_tx(full, share);
                        bulk_g2s_chunked(smem_u32(ringB + st * kSB), src + (size_t)kb * kb_bytes + brank * share, share, full);
                    }
                }
            }
        }
    } else if (warp == kRelayWarp) {
        if (kPair && !leader) {
            // ---------------- follower: forward ring-B / q-image arrivals to the leader's barriers
            if (lane == 0) {
                if (a.do_main) {
                    mbar_wait(smem_u32(&bars->q_full), 0);
                    mbar_arrive_cluster(mapa_u32(smem_u32(&bars->q_full), 0));
                }
                uint32_t total = 0;
                for (int o = 0; o < a.n_d; ++o)
                    if (!a.dops[o].ring_a && a.dops[o].seg != 3) total += (a.dops[o].seg ? segs.n_seg : 1) * a.dops[o].kb;
                for (uint32_t it = 0; it < total; ++it) {
                    const uint32_t st = it % kNB, ph = (it / kNB) & 1u;
                    mbar_wait(smem_u32(&bars->fullB[st]), ph);
                    mbar_arrive_cluster(mapa_u32(smem_u32(&bars->fullB[st]), 0));
                }
            }
        } else if (lane == 0) {
            // ---------------- leader: issuer of the FiLM projections S = A_emb . We (ring A).  They depend only on
            // the S accumulator being free (s_free: the row threads are done reading the previous projection), so
            // this lane runs ahead of the dependent chain and the tensor pipe fills its gaps with these MMAs.
            constexpr int kM = kPair ? 2 * kTileRows : kTileRows;
            const uint32_t idesc_s = make_idesc<kBf16>(kM, 256);
            uint32_t zero_mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint32_t it = 0;
            for (int o = 0; o < a.n_s; ++o) {
                if (o > 0) {
                    mbar_wait(smem_u32(&bars->s_free), (uint32_t)(o - 1) & 1u);
                    tc_fence_after();
                }
                for (int sgi = 0; sgi < kSopStages; ++sgi, ++it) {
                    const uint32_t st = it % kNA, ph = (it / kNA) & 1u;
                    mbar_wait(smem_u32(&bars->fullA[st]), ph);
                    tc_fence_after();
                    const uint32_t stage = smem_u32(ringA + st * kSA);
                    if constexpr (kPair) {
                        umma_kblock_2cta(tmem_base + kColS, stage, stage + kStageABytes, idesc_s, sgi > 0, zero_mask);
                        umma_commit_2cta(smem_u32(&bars->emptyA[st]));
                    } else {
                        umma_kblock(tmem_base + kColS, stage, stage + kStageABytes, idesc_s, sgi > 0);
                        umma_commit(smem_u32(&bars->emptyA[st]));
                    }
                    if (sgi == 0) tl_mark(a, 300 + o);
                }
                tl_mark(a, 310 + o);
                if constexpr (kPair) umma_commit_2cta(smem_u32(&bars->d_ready[0]));
                else umma_commit(smem_u32(&bars->d_ready[0]));
            }
        }
    } else if (warp == kMmaWarp && !leader) {
        // ---------------- follower only: forward ring-A arrivals to the leader
        if (lane == 0) {
            uint32_t total = (uint32_t)a.n_s * kSopStages;
            for (int o = 0; o < a.n_d; ++o) total += a.dops[o].ring_a ? 1u : 0u;
            for (uint32_t it = 0; it < total; ++it) {
                const uint32_t st = it % kNA, ph = (it / kNA) & 1u;
                mbar_wait(smem_u32(&bars->fullA[st]), ph);
                mbar_arrive_cluster(mapa_u32(smem_u32(&bars->fullA[st]), 0));
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- leader: issuer of the dependent GEMMs, strictly in chain order with blocking waits
        if (lane == 0) {
            constexpr int kM = kPair ? 2 * kTileRows : kTileRows;
            const uint32_t awork = smem_u32(awork_p);
            uint32_t zero_mask[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            auto commit = [](uint32_t bar) {
                if constexpr (kPair) umma_commit_2cta(bar);
                else umma_commit(bar);
            };
            auto kblock = [&](uint32_t d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, bool acc, const uint32_t* m, bool masked) {
                if constexpr (kPair) umma_kblock_2cta(d, a_addr, b_addr, idesc, acc, m);
                else if (masked) umma_kblock_masked(d, a_addr, b_addr, idesc, acc, m);
                else umma_kblock(d, a_addr, b_addr, idesc, acc);
            };
            uint32_t itA = (uint32_t)a.n_s * kSopStages, itB = 0, a_phase = 0;     // ring A: behind the FiLM projections
            for (int d_idx = 0; d_idx < a.n_d; ++d_idx) {
                const DOp op = a.dops[d_idx];
                if (op.wait == 1) {
                    mbar_wait(smem_u32(&bars->a_ready), a_phase & 1u);
                    ++a_phase;
                } else if (op.wait == 2) {
                    mbar_wait(smem_u32(&bars->q_full), 0);
                }
                tc_fence_after();
                const uint32_t idesc = make_idesc<kBf16>(kM, op.n);
                const uint32_t kb_stride = (uint32_t)op.n * 128u / kHalf;     // bytes of one k-block of this CTA's B share
                if (op.seg == 3) {
                    // K^T V partial of this tile: E^T . V over the 128 tokens, both operands MN-major images written by
                    // the row threads (E at ring A, V behind it); one pass per clip segment of the tile.
                    if constexpr (!kPair) {
                        const int nvalid = max(0, min(kTileRows, a.M - (int)blockIdx.x * kTileRows));
                        const int e_rows = min(nvalid, ((int)blockIdx.x * kTileRows / a.T + 1) * a.T - (int)blockIdx.x * kTileRows);
                        const int passes = nvalid > e_rows ? 2 : 1;
                        const uint32_t idmn = make_idesc_mn<kBf16>(kTileRows, kTileRows);
                        const uint32_t eimg = smem_u32(ringA), vimg = eimg + kAworkBytes;
                        for (int ps = 0; ps < passes; ++ps) {
                            if (ps > 0) {
                                mbar_wait(smem_u32(&bars->a_ready), a_phase & 1u);
                                ++a_phase;
                                tc_fence_after();
                            }
                            for (int ks = 0; ks < 8; ++ks)
                                umma_f16(tmem_base + kColW, make_desc_mnmajor_sw128(eimg + ks * 2048), make_desc_mnmajor_sw128(vimg + ks * 2048),
                                         idmn, ks > 0);
                            umma_commit(smem_u32(&bars->d_ready[2]));
                        }
                    }
                    tl_mark(a, 200 + d_idx);
                    continue;
                }
                if (op.ring_a) {            // whole operand in one ring-A stage
                    const uint32_t st = itA % kNA, ph = (itA / kNA) & 1u;
                    mbar_wait(smem_u32(&bars->fullA[st]), ph);
                    tc_fence_after();
                    const uint32_t b_base = smem_u32(ringA + st * kSA + kStageABytes);
                    for (int kb = 0; kb < op.kb; ++kb)
                        kblock(tmem_base + op.d_col, awork + kb * kABlockBytes, b_base + kb * kb_stride, idesc, op.accumulate || kb > 0,
                               zero_mask, false);
                    commit(smem_u32(&bars->emptyA[st]));
                    ++itA;
                } else {                    // one k-block per ring-B stage
                    const int n_st = op.seg ? segs.n_seg : 1;
                    for (int sg = 0; sg < n_st; ++sg) {
                        uint32_t m[8];
                        if (op.seg) segs.mask(sg, m, a.mask_invert == 1);
                        for (int kb = 0; kb < op.kb; ++kb, ++itB) {
                            const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                            mbar_wait(smem_u32(&bars->fullB[st]), ph);
                            tc_fence_after();
                            const uint32_t b_base = smem_u32(ringB + st * kSB);
                            if (op.seg) kblock(tmem_base + op.d_col, awork + kb * kABlockBytes, b_base, idesc, kb > 0, m, true);
                            else kblock(tmem_base + op.d_col, awork + kb * kABlockBytes, b_base, idesc, op.accumulate || kb > 0, zero_mask, false);
                            commit(smem_u32(&bars->emptyB[st]));
                        }
                    }
                }
                tl_mark(a, 200 + d_idx);
                if (op.commit != 255) commit(smem_u32(&bars->d_ready[op.commit]));
            }
        }
    } else {
        const uint32_t a_ready_addr = kPair ? mapa_u32(smem_u32(&bars->a_ready), 0) : smem_u32(&bars->a_ready);
        const uint32_t s_free_addr = kPair ? mapa_u32(smem_u32(&bars->s_free), 0) : smem_u32(&bars->s_free);
        const uint32_t lq = warp & 3, cq = warp >> 2;
        const uint32_t r = lq * 32 + lane;            // row of the tile == TMEM lane
        const uint32_t c0 = cq * 32;                  // first of this thread's 32 features (heads 2cq, 2cq+1)
        const uint32_t trow = tmem_base + ((lq * 32) << 16);
        const uint32_t awork = smem_u32(awork_p);
        const long g = (long)blockIdx.x * kTileRows + r;
        const bool valid = g < a.M;
        const int b = valid ? (int)(g / a.T) : 0;
        const int t = valid ? (int)(g - (long)b * a.T) : 0;
        uint32_t ph[3] = {0, 0, 0};
        RowStats rs{xchg, 1 + lq, r, cq, 0};
        float mean, rstd;
        float v[32];

        // ---- residual stream -> TMEM
        {
            const float4* src = reinterpret_cast<const float4*>(a.h + blk_index(g, c0, kD));      // chunk stride 512 floats
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 f = valid ? src[i * 128] : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * i] = f.x, v[4 * i + 1] = f.y, v[4 * i + 2] = f.z, v[4 * i + 3] = f.w;
            }
            tmem_st32(trow + kColH + c0, v);
            tmem_wait_st();
        }

        if (a.do_main) {
            // ================= self-attention tail: y = q . blockdiag(A_sa) (tensor cores) ; h += Styl(y)
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 101);
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 102);                                   // S = A_emb . We_sa
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStSa, awork, r, c0);
            tc_fence_before();                                           // S consumed: the next FiLM projection may start
            __syncwarp();
            if (lane == 0) bar_arrive<kPair>(s_free_addr);
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 151);                                          // -> h += A . Wo_sa

            // ================= cross-attention
            rows_wait(bars, 1, ph[1]); if (threadIdx.x == 0) tl_mark(a, 103);
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmStSa + kStBo + c0);                  // deferred bias of Wo_sa
            tmem_st32(trow + kColH + c0, v);
            row_stats32(rs, v, mean, rstd);
            normalize32(v, mean, rstd);                                  // LN affine folded into Wq_ca
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 152);                                          // -> W = LN(h) . Wq_ca
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 104);
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmCaBq + c0);
            softmax16(v);
            softmax16(v + 16);
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 153);                                          // -> W = softmax(q) . blockdiag(A_ca)
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 105);
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 106);                                   // S = A_emb . We_ca
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStCa, awork, r, c0);
            tc_fence_before();                                           // S consumed: the next FiLM projection may start
            __syncwarp();
            if (lane == 0) bar_arrive<kPair>(s_free_addr);
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 154);                                          // -> h += A . Wo_ca

            // ================= FFN (no pre-norm, reference transformer.py:170-173)
            rows_wait(bars, 1, ph[1]); if (threadIdx.x == 0) tl_mark(a, 107);
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmStCa + kStBo + c0);                  // deferred bias of Wo_ca
            tmem_st32(trow + kColH + c0, v);
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 155);                                          // -> W[0:64] = h . W1
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 108);
            {
                float u[16];                                             // hidden 64 = 4 quarters of 16
                tmem_ld16(trow + kColW + 16 * cq, u);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = gelu_erf_f(u[i] + prm[kPrmFfB1 + 16 * cq + i]);
                store_a16<kBf16>(awork, r, 16 * cq, u);
            }
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 156);                                          // -> W = GELU(.) . W2
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 109);
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmFfB2 + c0);
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 110);                                   // S = A_emb . We_ffn
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStFf, awork, r, c0);
            tc_fence_before();                                           // S consumed: the next FiLM projection may start
            __syncwarp();
            if (lane == 0) bar_arrive<kPair>(s_free_addr);
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 157);                                          // -> h += A . Wo_ffn
            rows_wait(bars, 1, ph[1]); if (threadIdx.x == 0) tl_mark(a, 111);
        }

        // ---- final value of the residual stream for this launch (deferred bias of the last FFN block)
        tmem_ld32(trow + kColH + c0, v);
        tmem_wait_ld();
        if (a.do_main) add_bias32(v, prm + kPrmStFf + kStBo + c0);
        if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.h + blk_index(g, c0, kD));
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i * 128] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }

        if (a.do_sa1) {
            // ================= next layer's self-attention head: LN -> q | k | v
            row_stats32(rs, v, mean, rstd);
            normalize32(v, mean, rstd);                                  // LN affine folded into Wq/Wk/Wv
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            rows_publish<kPair>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 158);
            rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 112);
            const bool keep = valid && (a.length == nullptr || (long long)t < a.length[b]);
            // q: softmax over head-dim, written as this tile's packed A-operand image for the next launch
            tmem_ld32(trow + kColS + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm_sa + kPrmSaBq + c0);
            softmax16(v);
            softmax16(v + 16);
            // staged in the (now idle) operand buffer in exactly the image format, then one bulk store per tile
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            fence_async_smem();
            named_bar_sync(5, kRowThreads);
            if (threadIdx.x == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(a.q_img + (size_t)blockIdx.x * kAworkBytes),
                             "r"(awork), "r"((uint32_t)kAworkBytes)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            // k (masked frames get -1e6 before the time softmax), v (masked frames zeroed): reference :107,:114
            if (!a.fuse_kv) {
                // general path (T < 128: a tile may span many clips): k|v to global, reduced by kv_reduce_kernel
                tmem_ld32(trow + kColS + 128 + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm_sa + kPrmSaBk + c0);
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(a.kv + blk_index(g, c0, 256));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                        if (!keep) o.x += -1000000.f, o.y += -1000000.f, o.z += -1000000.f, o.w += -1000000.f;
                        dst[i * 128] = o;
                    }
                }
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm_sa + kPrmSaBv + c0);
                if (valid) {
                    float4* dst = reinterpret_cast<float4*>(a.kv + blk_index(g, kD + c0, 256));
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        dst[i * 128] = keep ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (!kPair) {
                // Fused time-axis softmax + K^T V (reference :111,:117) on the tensor cores.  With T >= 128 a tile touches
                // at most two clips.  Two 32 KB operand-image buffers X, Y are all the scratch it needs:
                //   k (16-bit) -> X ; per-segment column maxima by a column scan of X ; E = exp(k - max) -> X ;
                //   V -> Y (rows of the other segment zeroed) ; P = E^T V as 8 MN-major MMAs over the tile's tokens ;
                //   per-segment column sums by a column scan of the E image (same rounded values as the MMA sees).
                // The partial (max, sum, diagonal 16x16 blocks of P) goes to global memory; the CTA that completes a clip
                // merges its partials (online-softmax rescaling) into the clip's block-diagonal B-operand image.
                float* pm = reinterpret_cast<float*>(ringB);           // [4 qr][2 seg][128] exchange (max, then sums)
                float* msm = pm + 1024;                                // [2][128] maxima
                float* ssm = msm + 256;                                // [2][128] sums
                int* flags = reinterpret_cast<int*>(ssm + 256);
                uint8_t* Xp = ringA;
                const uint32_t eimg = smem_u32(ringA), vimg = eimg + kAworkBytes;
                const int row0 = blockIdx.x * kTileRows;
                const int first_clip = row0 / a.T;
                const int nvalid = max(0, min(kTileRows, a.M - row0));
                const int e = min(nvalid, (first_clip + 1) * a.T - row0);   // rows [0,e): first clip, [e,nvalid): next clip
                const int n_seg = nvalid == 0 ? 0 : (nvalid > e ? 2 : 1);
                const int tx = threadIdx.x;
                const bool in_tile = (int)r < nvalid;
                const int myseg = (int)r >= e ? 1 : 0;
                const int col = tx & 127, qr = tx >> 7;
                // element (row, col) of a [128 x 128] 16-bit operand image
                auto img_at = [&](int row) -> const uint16_t* {
                    return reinterpret_cast<const uint16_t*>(Xp + (col >> 6) * kABlockBytes + sw128_offset(row, (col & 63) >> 3) + (col & 7) * 2);
                };
                auto to_f = [](uint16_t u) -> float {
                    if constexpr (kBf16) return __uint_as_float((uint32_t)u << 16);
                    else return __half2float(__ushort_as_half(u));
                };
                float kx[32];
                tmem_ld32(trow + kColS + 128 + c0, kx);
                tmem_wait_ld();
                add_bias32(kx, prm_sa + kPrmSaBk + c0);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (!keep) kx[i] += -1000000.f;
                    if (!in_tile) kx[i] = -INFINITY;
                    if constexpr (!kBf16) kx[i] = fmaxf(kx[i], -60000.f);      // fp16 image of k: keep the mask finite
                }
                store_a16<kBf16>(eimg, r, c0, kx);
                store_a16<kBf16>(eimg, r, c0 + 16, kx + 16);
                float vx[32];
                tmem_ld32(trow + kColW + c0, vx);
                tmem_wait_ld();
                add_bias32(vx, prm_sa + kPrmSaBv + c0);
                named_bar_sync(5, kRowThreads);
                {   // column maxima per segment: this thread scans rows [32 qr, 32 qr + 32) of column `col`
                    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = 32 * qr + rr;
                        const float x = to_f(*img_at(row));
                        if (row < e) m0 = fmaxf(m0, x);
                        else m1 = fmaxf(m1, x);
                    }
                    pm[(qr * 2 + 0) * 128 + col] = m0;
                    pm[(qr * 2 + 1) * 128 + col] = m1;
                }
                named_bar_sync(5, kRowThreads);
                if (tx < 256) {
                    const int sg = tx >> 7;
                    msm[tx] = fmaxf(fmaxf(pm[(0 + sg) * 128 + col], pm[(2 + sg) * 128 + col]),
                                    fmaxf(pm[(4 + sg) * 128 + col], pm[(6 + sg) * 128 + col]));
                }
                named_bar_sync(5, kRowThreads);
                {   // E = exp(k - max) (0 for padding rows) -> X ; V of the first segment -> Y
                    const float* mrow = msm + myseg * 128 + c0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) kx[i] = in_tile ? exp2f((kx[i] - mrow[i]) * 1.4426950408889634f) : 0.f;
                    store_a16<kBf16>(eimg, r, c0, kx);
                    store_a16<kBf16>(eimg, r, c0 + 16, kx + 16);
                }
                if (!keep) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) vx[i] = 0.f;
                }
                if (tx == 0) tl_mark(a, 120);
                for (int ps = 0; ps < max(n_seg, 1); ++ps) {
                    {   // V image of this pass: rows of the other segment (and padding rows) are zero
                        float z[32];
                        const bool mine = in_tile && myseg == ps;
#pragma unroll
                        for (int i = 0; i < 32; ++i) z[i] = mine ? vx[i] : 0.f;
                        store_a16<kBf16>(vimg, r, c0, z);
                        store_a16<kBf16>(vimg, r, c0 + 16, z + 16);
                    }
                    rows_publish<kPair>(a_ready_addr, lane);                      // -> W = E^T . V  (8 MMAs over the tokens)
                    if (tx == 0) tl_mark(a, 122);
                    if (ps == 0) {
                        named_bar_sync(5, kRowThreads);                    // E image complete
                        float s0 = 0.f, s1 = 0.f;                          // column sums per segment from the rounded E
#pragma unroll 8
                        for (int rr = 0; rr < 32; ++rr) {
                            const int row = 32 * qr + rr;
                            const float x = to_f(*img_at(row));
                            if (row < e) s0 += x;
                            else s1 += x;
                        }
                        pm[(qr * 2 + 0) * 128 + col] = s0;
                        pm[(qr * 2 + 1) * 128 + col] = s1;
                        named_bar_sync(5, kRowThreads);
                        if (tx < 256) {
                            const int sg = tx >> 7;
                            ssm[tx] = (pm[(0 + sg) * 128 + col] + pm[(2 + sg) * 128 + col]) + (pm[(4 + sg) * 128 + col] + pm[(6 + sg) * 128 + col]);
                        }
                    }
                    rows_wait(bars, 2, ph[2]);
                    if (tx == 0) tl_mark(a, 123);
                    if (ps < n_seg) {
                        float* P = a.kv_part + ((size_t)blockIdx.x * 2 + ps) * kKvPartFloats;
                        if (cq == 0) {      // TMEM lane = key feature r; its head's 16 value columns are the diagonal block
                            float pr[32];
                            tmem_ld32(trow + kColW + 32 * lq, pr);
                            tmem_wait_ld();
                            float4* dst = reinterpret_cast<float4*>(P + 256 + (r >> 4) * 256 + (r & 15) * 16);
                            const int o = (lane & 16);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                dst[i] = o ? make_float4(pr[16 + 4 * i], pr[17 + 4 * i], pr[18 + 4 * i], pr[19 + 4 * i])
                                           : make_float4(pr[4 * i], pr[4 * i + 1], pr[4 * i + 2], pr[4 * i + 3]);
                        }
                        named_bar_sync(5, kRowThreads);                    // ssm written (pass 0) / everyone done reading W
                        if (tx < 128) P[tx] = msm[ps * 128 + tx], P[128 + tx] = ssm[ps * 128 + tx];
                    }
                }
                if (tx == 0) tl_mark(a, 124);
                named_bar_sync(5, kRowThreads);
                if (tx == 0 || tx == 32) {          // one arrival counter per clip; both segments in parallel
                    const int sg = tx >> 5;
                    int f = 0;
                    __threadfence();                // cumulative: orders the whole CTA's partials (bar.sync above)
                    if (sg < n_seg) {
                        const int clip = first_clip + sg;
                        const int ntiles = ((clip + 1) * a.T - 1) / kTileRows - (clip * a.T) / kTileRows + 1;
                        const int old = atomicAdd(a.clip_cnt + clip, 1);
                        if (old + 1 == ntiles) {
                            f = 1;
                            a.clip_cnt[clip] = 0;          // ready for the next launch
                        }
                    }
                    flags[sg] = f;
                }
                named_bar_sync(5, kRowThreads);
                if (tx == 0) tl_mark(a, 125);
                const int hh = tx >> 6, sub = tx & 63, d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
                for (int sg = 0; sg < n_seg; ++sg) {
                    if (!flags[sg]) continue;
                    __threadfence();
                    const int clip = first_clip + sg;
                    const int t_first = (clip * a.T) / kTileRows, t_last = ((clip + 1) * a.T - 1) / kTileRows;
                    // one pass with online rescaling
                    float M0 = -INFINITY, M1 = -INFINITY, a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
                    for (int ti = t_first; ti <= t_last; ++ti) {
                        const float* P = a.kv_part + ((size_t)ti * 2 + (clip - (ti * kTileRows) / a.T)) * kKvPartFloats;
                        const float mi0 = __ldcg(P + 16 * hh + d0), mi1 = __ldcg(P + 16 * hh + d0 + 1);
                        const float si0 = __ldcg(P + 128 + 16 * hh + d0), si1 = __ldcg(P + 128 + 16 * hh + d0 + 1);
                        const float* Pa = P + 256 + hh * 256;
                        const float2 r0 = __ldcg(reinterpret_cast<const float2*>(Pa + d0 * 16 + l0));
                        const float2 r1 = __ldcg(reinterpret_cast<const float2*>(Pa + (d0 + 1) * 16 + l0));
                        const float n0 = fmaxf(M0, mi0), n1 = fmaxf(M1, mi1);
                        const float c0s = __expf(M0 - n0), c1s = __expf(M1 - n1), w0 = __expf(mi0 - n0), w1 = __expf(mi1 - n1);
                        M0 = n0, M1 = n1;
                        s0 = fmaf(s0, c0s, si0 * w0), s1 = fmaf(s1, c1s, si1 * w1);
                        a00 = fmaf(a00, c0s, r0.x * w0), a01 = fmaf(a01, c0s, r0.y * w0);
                        a10 = fmaf(a10, c1s, r1.x * w1), a11 = fmaf(a11, c1s, r1.y * w1);
                    }
                    uint8_t* img = a.bd_sa_out + (size_t)clip * kAworkBytes;
                    const float o[2][2] = {{a00 / s0, a01 / s0}, {a10 / s1, a11 / s1}};
#pragma unroll
                    for (int dd = 0; dd < 2; ++dd) {
                        const int ki = 16 * hh + d0 + dd;
                        uint8_t* base = img + (size_t)(ki >> 6) * kABlockBytes + (ki & 7) * 2;
#pragma unroll
                        for (int ll = 0; ll < 2; ++ll) {
                            const int nj = 16 * hh + l0 + ll;
                            *reinterpret_cast<uint16_t*>(base + sw128_offset(nj, (ki & 63) >> 3)) = pack1<kBf16>(o[dd][ll]);
                        }
                    }
                }
            } else {
                // Fused time-axis softmax + K^T V (reference :111,:117).  With T >= 128 a tile touches at most two
                // clips.  Per (tile, clip segment): column max, exp, column sums and the per-head 16x16 outer products
                // over the segment's tokens, from shared memory (the operand rings are idle now); the partial
                // (max, sum, K^T V) goes to global memory and the CTA that completes a clip merges its partials
                // (online-softmax style) into that clip's block-diagonal B-operand image for the next layer.
                constexpr int LDS = 132;                               // padded row stride (floats)
                float* kS = reinterpret_cast<float*>(ringA);           // [128][132]
                float* vS = kS + kTileRows * LDS;                      // [128][132]
                float* pm = reinterpret_cast<float*>(ringB);           // [4][2][128] partial maxima, then [2][128] maxima
                float* msm = pm + 1024;
                int* flags = reinterpret_cast<int*>(msm + 256);
                const int row0 = blockIdx.x * kTileRows;
                const int first_clip = row0 / a.T;
                const int nvalid = max(0, min(kTileRows, a.M - row0));     // 0 for the padding tile of an odd pair
                const int e = min(nvalid, (first_clip + 1) * a.T - row0);   // rows [0,e): first clip, [e,nvalid): next clip
                const int n_seg = nvalid == 0 ? 0 : (nvalid > e ? 2 : 1);
                tmem_ld32(trow + kColS + 128 + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm_sa + kPrmSaBk + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (!keep) o.x += -1000000.f, o.y += -1000000.f, o.z += -1000000.f, o.w += -1000000.f;
                    if (!valid) o = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    *reinterpret_cast<float4*>(kS + r * LDS + c0 + 4 * i) = o;
                }
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm_sa + kPrmSaBv + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    *reinterpret_cast<float4*>(vS + r * LDS + c0 + 4 * i) =
                        keep ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
                named_bar_sync(5, kRowThreads);
                const int tx = threadIdx.x, col = tx & 127, qr = tx >> 7;
                {   // column maxima per segment
                    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = 32 * qr + rr;
                        const float x = kS[row * LDS + col];
                        if (row < e) m0 = fmaxf(m0, x);
                        else m1 = fmaxf(m1, x);
                    }
                    pm[(qr * 2 + 0) * 128 + col] = m0;
                    pm[(qr * 2 + 1) * 128 + col] = m1;
                }
                named_bar_sync(5, kRowThreads);
                if (tx < 256) {
                    const int sg = tx >> 7;
                    msm[tx] = fmaxf(fmaxf(pm[(0 + sg) * 128 + col], pm[(2 + sg) * 128 + col]),
                                    fmaxf(pm[(4 + sg) * 128 + col], pm[(6 + sg) * 128 + col]));
                }
                named_bar_sync(5, kRowThreads);
                {   // exp in place (padded rows hold -inf -> 0)
                    const float mm0 = msm[col], mm1 = msm[128 + col];
#pragma unroll 8
                    for (int rr = 0; rr < 32; ++rr) {
                        const int row = 32 * qr + rr;
                        const float x = kS[row * LDS + col];
                        kS[row * LDS + col] = row < nvalid ? expf(x - (row < e ? mm0 : mm1)) : 0.f;
                    }
                }
                named_bar_sync(5, kRowThreads);
                const int hh = tx >> 6, sub = tx & 63, d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
                for (int sg = 0; sg < n_seg; ++sg) {
                    const int lo = sg ? e : 0, hi = sg ? nvalid : e;
                    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
#pragma unroll 4
                    for (int tok = lo; tok < hi; ++tok) {
                        const float2 ee = *reinterpret_cast<const float2*>(kS + tok * LDS + 16 * hh + d0);
                        const float2 vv2 = *reinterpret_cast<const float2*>(vS + tok * LDS + 16 * hh + l0);
                        a00 = fmaf(ee.x, vv2.x, a00), a01 = fmaf(ee.x, vv2.y, a01);
                        a10 = fmaf(ee.y, vv2.x, a10), a11 = fmaf(ee.y, vv2.y, a11);
                        s0 += ee.x, s1 += ee.y;
                    }
                    float* P = a.kv_part + ((size_t)blockIdx.x * 2 + sg) * kKvPartFloats;
                    float* Pa = P + 256 + hh * 256;
                    *reinterpret_cast<float2*>(Pa + d0 * 16 + l0) = make_float2(a00, a01);
                    *reinterpret_cast<float2*>(Pa + (d0 + 1) * 16 + l0) = make_float2(a10, a11);
                    if (l0 == 0) P[128 + 16 * hh + d0] = s0, P[128 + 16 * hh + d0 + 1] = s1;
                    if (tx < 128) P[tx] = msm[sg * 128 + tx];
                }
                __threadfence();
                named_bar_sync(5, kRowThreads);
                if (tx == 0) {
                    for (int sg = 0; sg < 2; ++sg) {
                        flags[sg] = 0;
                        if (sg < n_seg) {
                            const int clip = first_clip + sg;
                            const int ntiles = ((clip + 1) * a.T - 1) / kTileRows - (clip * a.T) / kTileRows + 1;
                            const int old = atomicAdd(a.clip_cnt + clip, 1);
                            if (old + 1 == ntiles) {
                                flags[sg] = 1;
                                a.clip_cnt[clip] = 0;          // ready for the next launch
                            }
                        }
                    }
                }
                named_bar_sync(5, kRowThreads);
                for (int sg = 0; sg < n_seg; ++sg) {
                    if (!flags[sg]) continue;
                    __threadfence();
                    const int clip = first_clip + sg;
                    const int t_first = (clip * a.T) / kTileRows, t_last = ((clip + 1) * a.T - 1) / kTileRows;
                    float M0 = -INFINITY, M1 = -INFINITY;
                    for (int ti = t_first; ti <= t_last; ++ti) {
                        const float* P = a.kv_part + ((size_t)ti * 2 + (clip - (ti * kTileRows) / a.T)) * kKvPartFloats;
                        M0 = fmaxf(M0, __ldcg(P + 16 * hh + d0));
                        M1 = fmaxf(M1, __ldcg(P + 16 * hh + d0 + 1));
                    }
                    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
                    for (int ti = t_first; ti <= t_last; ++ti) {
                        const float* P = a.kv_part + ((size_t)ti * 2 + (clip - (ti * kTileRows) / a.T)) * kKvPartFloats;
                        const float w0 = expf(__ldcg(P + 16 * hh + d0) - M0), w1 = expf(__ldcg(P + 16 * hh + d0 + 1) - M1);
                        s0 = fmaf(__ldcg(P + 128 + 16 * hh + d0), w0, s0);
                        s1 = fmaf(__ldcg(P + 128 + 16 * hh + d0 + 1), w1, s1);
                        const float* Pa = P + 256 + hh * 256;
                        const float2 r0 = __ldcg(reinterpret_cast<const float2*>(Pa + d0 * 16 + l0));
                        const float2 r1 = __ldcg(reinterpret_cast<const float2*>(Pa + (d0 + 1) * 16 + l0));
                        a00 = fmaf(r0.x, w0, a00), a01 = fmaf(r0.y, w0, a01);
                        a10 = fmaf(r1.x, w1, a10), a11 = fmaf(r1.y, w1, a11);
                    }
                    uint8_t* img = a.bd_sa_out + (size_t)clip * kAworkBytes;
                    const float o[2][2] = {{a00 / s0, a01 / s0}, {a10 / s1, a11 / s1}};
#pragma unroll
                    for (int dd = 0; dd < 2; ++dd) {
                        const int ki = 16 * hh + d0 + dd;
                        uint8_t* base = img + (size_t)(ki >> 6) * kABlockBytes + (ki & 7) * 2;
#pragma unroll
                        for (int ll = 0; ll < 2; ++ll) {
                            const int nj = 16 * hh + l0 + ll;
                            *reinterpret_cast<uint16_t*>(base + sw128_offset(nj, (ki & 63) >> 3)) = pack1<kBf16>(o[dd][ll]);
                        }
                    }
                }
            }
            // the q image must have left shared memory before the CTA exits
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    }
    if (threadIdx.x == 0) tl_mark(a, 2);
    tc_fence_before();
    if constexpr (kPair) cluster_sync_all();      // the peer may still be reading this CTA's shared memory / barriers
    else __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        if constexpr (kPair) tmem_dealloc2(tmem_base, 512);
        else tmem_dealloc(tmem_base, 512);
    }
}

constexpr int kGemmSmemBytes = kStages * kStageBytes + sizeof(TileBarriers) + 1024;
constexpr int kLayerSmemBytes = kRingAStages * kStageBytes + kRingBStages * kRingBStageBytes + kAworkBytes +
                                (kPrmFloats + 384) * 4 + 512 * 8 + sizeof(LayerBarriers) + 1024;
static_assert(kLayerSmemBytes <= 232448, "layer kernel exceeds the 227 KB shared-memory limit");

}  // namespace dc
