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
constexpr int kProducerWarp = 16, kMmaWarp = 17;
constexpr int kTileThreads = kRowThreads + 64;    // 576

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
};

template <bool kBf16>
__global__ void __launch_bounds__(kTileThreads, 1) gemm_rows_kernel(const __grid_constant__ GemmRowsArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    TileBarriers* bars = reinterpret_cast<TileBarriers*>(smem + kStages * kStageBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tile_setup(bars, warp, lane);
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
            producer_loop(&op, 1, a.w_img, a.a_img + (size_t)blockIdx.x * a.kblocks * kStageABytes, ring, bars);
    } else if (warp == kMmaWarp) {
        if (lane == 0) mma_loop<kBf16>(&op, 1, ring, nullptr, bars, tmem_base);
    } else {
        const int lq = warp & 3, cq = warp >> 2;
        const long g = (long)blockIdx.x * kTileRows + lq * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(lq * 32) << 16);
        mbar_wait(smem_u32(&bars->d_ready[0]), 0);
        tc_fence_after();
        // column quarter cq covers columns [64 cq, 64 cq + 64) in 16-column pieces
        for (int c = 64 * cq; c < min(a.N, 64 * cq + 64); c += 16) {
            float v[16];
            tmem_ld16(trow + c, v);
            tmem_wait_ld();
            if (g < a.M) {
                float4* dst = reinterpret_cast<float4*>(a.out + (size_t)g * a.ldo + c);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (a.bias) {
                        o.x += a.bias[c + 4 * i];
                        o.y += a.bias[c + 4 * i + 1];
                        o.z += a.bias[c + 4 * i + 2];
                        o.w += a.bias[c + 4 * i + 3];
                    }
                    dst[i] = o;
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

struct LayerArgs {
    TileOp ops[kMaxOps];
    int n_ops;
    int do_main;            // SA tail (q.A + stylization), cross-attention, FFN of layer l
    int do_sa1;             // LayerNorm + q/k/v projections of layer l+1
    int M, T;
    const uint8_t* wbuf;    // packed weights (whole model)
    const uint8_t* aemb;    // A_emb image [tiles][8][16 KB]
    const float* prm;       // parameter block of layer l (do_main)
    const float* prm_next;  // parameter block of layer l+1 (do_sa1; only the SA biases are read)
    float* h;               // [Mpad][128] residual stream (in/out)
    uint16_t* q;            // [Mpad][128] softmax_hd(Q) of the self-attention (16-bit)
    float* kv;              // [Mpad][256] k | v of the self-attention
    const float* A_sa;      // [B][8][16][16]   softmax_T(K)^T V of layer l self-attention
    const float* A_ca;      // [B][...]: cross-attention K^T V of layer l, clip stride a_ca_stride
    int a_ca_stride;
    const long long* length;  // [B] or null (all frames valid)
};

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

// y[16] = q[16] . A[16][16]  (A row-major [d][l] in global memory, read through L1)
__device__ __forceinline__ void head_apply(const float* q, const float* __restrict__ Ah, float* y) {
#pragma unroll
    for (int l = 0; l < 16; ++l) y[l] = 0.f;
    const float4* Ap = reinterpret_cast<const float4*>(Ah);
#pragma unroll
    for (int d = 0; d < 16; ++d) {
#pragma unroll
        for (int l4 = 0; l4 < 4; ++l4) {
            const float4 av = __ldg(Ap + d * 4 + l4);
            y[4 * l4 + 0] = fmaf(q[d], av.x, y[4 * l4 + 0]);
            y[4 * l4 + 1] = fmaf(q[d], av.y, y[4 * l4 + 1]);
            y[4 * l4 + 2] = fmaf(q[d], av.z, y[4 * l4 + 2]);
            y[4 * l4 + 3] = fmaf(q[d], av.w, y[4 * l4 + 3]);
        }
    }
}

__device__ __forceinline__ void softmax16(float* q) {
    float mx = q[0];
#pragma unroll
    for (int i = 1; i < 16; ++i) mx = fmaxf(mx, q[i]);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        q[i] = __expf(q[i] - mx);
        s += q[i];
    }
    const float inv = __fdividef(1.f, s);
#pragma unroll
    for (int i = 0; i < 16; ++i) q[i] *= inv;
}

// Row statistics over 128 features held as 4 x 32 registers by the 4 warps that share a row:
// local (mean, M2) -> shared memory -> 128-thread named barrier -> Chan combine.  `buf` alternates
// between two exchange buffers so consecutive calls need no second barrier.
struct RowStats {
    float2* xchg;      // [2][4][128]
    uint32_t bar_id;   // 1 + lq
    uint32_t r;        // row within the tile
    uint32_t cq;
    uint32_t flip;
};

__device__ __forceinline__ void row_stats32(RowStats& rs, const float* v, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += v[i];
    const float lm = s * (1.f / 32.f);
    float m2 = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const float d = v[i] - lm;
        m2 = fmaf(d, d, m2);
    }
    float2* buf = rs.xchg + rs.flip * 512;
    rs.flip ^= 1u;
    buf[rs.cq * 128 + rs.r] = make_float2(lm, m2);
    named_bar_sync(rs.bar_id, 128);
    const float2 p0 = buf[rs.r], p1 = buf[128 + rs.r], p2 = buf[256 + rs.r], p3 = buf[384 + rs.r];
    mean = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
    const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
    const float M2 = (p0.y + p1.y) + (p2.y + p3.y) + 32.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    rstd = rsqrtf(M2 * (1.f / kD) + kLnEps);
}

// A operand <- SiLU( LN(y) * (1 + scale) + shift ) for this thread's 32 features [c0, c0+32)
// (reference transformer.py:77-80).  The scale|shift accumulator in TMEM columns kColS is laid out
//   [scale 0..63 | shift 0..63 | scale 64..127 | shift 64..127]; st = stylization params in smem.
template <bool kBf16>
__device__ __forceinline__ void film_to_a(uint32_t trow, const float* y, float mean, float rstd, const float* st, uint32_t awork,
                                          uint32_t r, uint32_t c0) {
    const uint32_t sbase = trow + kColS + (c0 >> 6) * 128 + (c0 & 63);     // scale column of feature c0
    const float* be = st + kStBe + (c0 >> 6) * 128 + (c0 & 63);
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
        float sc[16], sh[16], o[16];
        tmem_ld16(sbase + 16 * hf, sc);
        tmem_ld16(sbase + 64 + 16 * hf, sh);
        tmem_wait_ld();
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
            const float4 g4 = *reinterpret_cast<const float4*>(st + kStG + c0 + 16 * hf + 4 * i4);
            const float4 b4 = *reinterpret_cast<const float4*>(st + kStB + c0 + 16 * hf + 4 * i4);
            const float4 es = *reinterpret_cast<const float4*>(be + 16 * hf + 4 * i4);
            const float4 eh = *reinterpret_cast<const float4*>(be + 64 + 16 * hf + 4 * i4);
            const float gg[4] = {g4.x, g4.y, g4.z, g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
            const float ss[4] = {es.x, es.y, es.z, es.w}, hh[4] = {eh.x, eh.y, eh.z, eh.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int i = 4 * i4 + k;
                const float n = fmaf((y[16 * hf + i] - mean) * rstd, gg[k], bb[k]);
                const float v = fmaf(n, sc[i] + ss[k], sh[i] + hh[k]);
                o[i] = silu_f<kBf16>(v);
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
        v[4 * i4] += b4.x, v[4 * i4 + 1] += b4.y, v[4 * i4 + 2] += b4.z, v[4 * i4 + 3] += b4.w;
    }
}

// row threads signal "A operand (and any TMEM writes) ready"
__device__ __forceinline__ void rows_publish(TileBarriers* bars) {
    fence_async_smem();
    tc_fence_before();
    mbar_arrive(smem_u32(&bars->a_ready));
}
__device__ __forceinline__ void rows_wait(TileBarriers* bars, int which, uint32_t& phase) {
    mbar_wait(smem_u32(&bars->d_ready[which]), phase & 1u);
    ++phase;
    tc_fence_after();
}

template <bool kBf16>
__global__ void __launch_bounds__(kTileThreads, 1) layer_kernel(const __grid_constant__ LayerArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* ring = smem;
    uint8_t* awork_p = smem + kStages * kStageBytes;
    float* prm = reinterpret_cast<float*>(awork_p + kAworkBytes);          // [kPrmFloats]
    float* prm_sa = prm + kPrmFloats;                                      // [384] SA biases of layer l+1
    float2* xchg = reinterpret_cast<float2*>(prm_sa + 384);                // [2][4][128]
    TileBarriers* bars = reinterpret_cast<TileBarriers*>(xchg + 1024);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (a.do_main)
        for (int i = threadIdx.x; i < kPrmFloats; i += kTileThreads) prm[i] = a.prm[i];
    if (a.do_sa1)
        for (int i = threadIdx.x; i < 384; i += kTileThreads) prm_sa[i] = a.prm_next[i];
    tile_setup(bars, warp, lane);
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == kProducerWarp) {
        if (lane == 0) producer_loop(a.ops, a.n_ops, a.wbuf, a.aemb + (size_t)blockIdx.x * 8 * kStageABytes, ring, bars);
    } else if (warp == kMmaWarp) {
        if (lane == 0) mma_loop<kBf16>(a.ops, a.n_ops, ring, awork_p, bars, tmem_base);
    } else {
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
            const float4* src = reinterpret_cast<const float4*>(a.h + (size_t)g * kD + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 f = valid ? src[i] : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * i] = f.x, v[4 * i + 1] = f.y, v[4 * i + 2] = f.z, v[4 * i + 3] = f.w;
            }
            tmem_st32(trow + kColH + c0, v);
            tmem_wait_st();
        }

        if (a.do_main) {
            // ================= self-attention tail: y = q . A_sa ; h += Styl(y)
            {
                const uint4* qrow = reinterpret_cast<const uint4*>(a.q + (size_t)g * kD + c0);
                const float* Ab = a.A_sa + (size_t)b * (kH * 256) + (2 * cq) * 256;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    float qv[16];
                    uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
                    if (valid) u0 = qrow[2 * hh], u1 = qrow[2 * hh + 1];
                    const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f = unpack2<kBf16>(uu[i]);
                        qv[2 * i] = f.x, qv[2 * i + 1] = f.y;
                    }
                    head_apply(qv, Ab + hh * 256, v + 16 * hh);
                }
            }
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_sa
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStSa, awork, r, c0);
            rows_publish(bars);                                          // -> h += A . Wo_sa

            // ================= cross-attention
            rows_wait(bars, 1, ph[1]);
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmStSa + kStBo + c0);                  // deferred bias of Wo_sa
            tmem_st32(trow + kColH + c0, v);
            row_stats32(rs, v, mean, rstd);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd;    // LN affine folded into Wq_ca
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish(bars);                                          // -> W = LN(h) . Wq_ca
            rows_wait(bars, 2, ph[2]);
            {
                float qv[32];
                tmem_ld32(trow + kColW + c0, qv);
                tmem_wait_ld();
                add_bias32(qv, prm + kPrmCaBq + c0);
                const float* Ab = a.A_ca + (size_t)b * a.a_ca_stride + (2 * cq) * 256;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    softmax16(qv + 16 * hh);
                    head_apply(qv + 16 * hh, Ab + hh * 256, v + 16 * hh);
                }
            }
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_ca
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStCa, awork, r, c0);
            rows_publish(bars);                                          // -> h += A . Wo_ca

            // ================= FFN (no pre-norm, reference transformer.py:170-173)
            rows_wait(bars, 1, ph[1]);
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmStCa + kStBo + c0);                  // deferred bias of Wo_ca
            tmem_st32(trow + kColH + c0, v);
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish(bars);                                          // -> W[0:64] = h . W1
            rows_wait(bars, 2, ph[2]);
            {
                float u[16];                                             // hidden 64 = 4 quarters of 16
                tmem_ld16(trow + kColW + 16 * cq, u);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) u[i] = gelu_erf_f(u[i] + prm[kPrmFfB1 + 16 * cq + i]);
                store_a16<kBf16>(awork, r, 16 * cq, u);
            }
            rows_publish(bars);                                          // -> W = GELU(.) . W2
            rows_wait(bars, 2, ph[2]);
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm + kPrmFfB2 + c0);
            row_stats32(rs, v, mean, rstd);
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_ffn
            film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStFf, awork, r, c0);
            rows_publish(bars);                                          // -> h += A . Wo_ffn
            rows_wait(bars, 1, ph[1]);
        }

        // ---- final value of the residual stream for this launch (deferred bias of the last FFN block)
        tmem_ld32(trow + kColH + c0, v);
        tmem_wait_ld();
        if (a.do_main) add_bias32(v, prm + kPrmStFf + kStBo + c0);
        if (valid) {
            float4* dst = reinterpret_cast<float4*>(a.h + (size_t)g * kD + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        }

        if (a.do_sa1) {
            // ================= next layer's self-attention head: LN -> q | k | v
            row_stats32(rs, v, mean, rstd);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (v[i] - mean) * rstd;    // LN affine folded into Wq/Wk/Wv
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            rows_publish(bars);
            rows_wait(bars, 2, ph[2]);
            const bool keep = valid && (a.length == nullptr || (long long)t < a.length[b]);
            // q: softmax over head-dim, stored 16-bit
            tmem_ld32(trow + kColS + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm_sa + kPrmSaBq + c0);
            softmax16(v);
            softmax16(v + 16);
            if (valid) {
                uint32_t p[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) p[i] = pack2<kBf16>(v[2 * i], v[2 * i + 1]);
                uint4* dst = reinterpret_cast<uint4*>(a.q + (size_t)g * kD + c0);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = make_uint4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
            }
            // k (masked frames get -1e6 before the time softmax), v (masked frames zeroed): reference :107,:114
            tmem_ld32(trow + kColS + 128 + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm_sa + kPrmSaBk + c0);
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.kv + (size_t)g * 256 + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    if (!keep) o.x += -1000000.f, o.y += -1000000.f, o.z += -1000000.f, o.w += -1000000.f;
                    dst[i] = o;
                }
            }
            tmem_ld32(trow + kColW + c0, v);
            tmem_wait_ld();
            add_bias32(v, prm_sa + kPrmSaBv + c0);
            if (valid) {
                float4* dst = reinterpret_cast<float4*>(a.kv + (size_t)g * 256 + kD + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    dst[i] = keep ? make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tile_teardown(bars, warp);
}

constexpr int kGemmSmemBytes = kStages * kStageBytes + sizeof(TileBarriers) + 1024;
constexpr int kLayerSmemBytes =
    kStages * kStageBytes + kAworkBytes + (kPrmFloats + 384) * 4 + 1024 * 8 + sizeof(TileBarriers) + 1024;

}  // namespace dc
