// tcgen05 / TMEM tile kernels.
//
// One CTA owns a tile of 128 tokens (= 128 TMEM lanes).  Warp roles:
//   warps 0-3 : "row" threads -- thread r owns token row r: reads accumulators from its TMEM lane,
//               does every row-wise op (LayerNorm, softmax over head-dim, q.A, FiLM, SiLU, GELU)
//               in registers, and writes the next GEMM's A operand to shared memory (K-major SW128)
//   warp 4    : producer -- streams packed operand blocks global/L2 -> smem ring with cp.async.bulk
//   warp 5    : MMA issuer (one elected lane) + TMEM allocator
// Synchronisation is mbarrier-only inside the main loop: full/empty per ring stage, a_ready
// (row threads -> MMA), d_ready per accumulator (tcgen05.commit -> row threads).
#pragma once
#include "simple_kernels.cuh"

namespace dc {

constexpr int kStages = 3;
constexpr int kStageABytes = kABlockBytes;        // 16 KB: one [128 x 64] A block (streamed operand)
constexpr int kStageWBytes = 32 * 1024;           // up to [256 x 64] weight block, or a whole small weight
constexpr int kStageBytes = kStageABytes + kStageWBytes;
constexpr int kAworkBytes = 2 * kABlockBytes;     // [128 x 128] A operand written by the row threads
constexpr int kTileThreads = 192;

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
struct TileSmem {
    uint8_t* ring;
    uint8_t* awork;
    TileBarriers* bars;
};

__device__ __forceinline__ void tile_setup(TileBarriers* bars, int warp, int lane) {
    if (warp == 4 && lane == 0) {
        for (int i = 0; i < kStages; ++i) {
            mbar_init(smem_u32(&bars->full[i]), 1);
            mbar_init(smem_u32(&bars->empty[i]), 1);
        }
        mbar_init(smem_u32(&bars->a_ready), kTileRows);
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars->d_ready[i]), 1);
        mbar_fence_init();
    }
    if (warp == 5) {
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
    if (warp == 5) {
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

    if (warp == 4) {
        if (lane == 0)
            producer_loop(&op, 1, a.w_img, a.a_img + (size_t)blockIdx.x * a.kblocks * kStageABytes, ring, bars);
    } else if (warp == 5) {
        if (lane == 0) mma_loop<kBf16>(&op, 1, ring, nullptr, bars, tmem_base);
    } else {
        const int r = threadIdx.x;
        const long g = (long)blockIdx.x * kTileRows + r;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        mbar_wait(smem_u32(&bars->d_ready[0]), 0);
        tc_fence_after();
        for (int c = 0; c < a.N; c += 16) {
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

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.f + __expf(-v)); }
__device__ __forceinline__ float gelu_erf_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }

// mean / rstd of the 128 columns at `col` of this thread's TMEM lane; optionally adds a bias vector
// first and writes the sum back (deferred bias of the preceding accumulate-GEMM).
__device__ __forceinline__ void row_stats(uint32_t taddr, const float* bias_smem, float& mean, float& rstd) {
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float v[16];
        tmem_ld16(taddr + 16 * c, v);
        tmem_wait_ld();
        if (bias_smem) {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += bias_smem[16 * c + i];
            tmem_st16(taddr + 16 * c, v);
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) sum += v[i];
    }
    if (bias_smem) tmem_wait_st();
    mean = sum * (1.f / kD);
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float v[16];
        tmem_ld16(taddr + 16 * c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float d = v[i] - mean;
            ss = fmaf(d, d, ss);
        }
    }
    rstd = rsqrtf(ss * (1.f / kD) + kLnEps);
}

// A operand <- (row - mean) * rstd          (LayerNorm without affine; affine folded downstream)
template <bool kBf16>
__device__ __forceinline__ void row_normalize_to_a(uint32_t taddr, float mean, float rstd, uint32_t awork, uint32_t r) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float v[16];
        tmem_ld16(taddr + 16 * c, v);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = (v[i] - mean) * rstd;
        store_a16<kBf16>(awork, r, 16 * c, v);
    }
}

// A operand <- SiLU( LN(y) * (1 + scale) + shift )   (reference transformer.py:77-80)
// y in TMEM columns kColW, scale|shift accumulator in kColS laid out
//   [scale 0..63 | shift 0..63 | scale 64..127 | shift 64..127]; st = stylization params in smem.
template <bool kBf16>
__device__ __forceinline__ void row_film_to_a(uint32_t trow, float mean, float rstd, const float* st, uint32_t awork, uint32_t r) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int kb = c >> 2, j = (c & 3) * 16;
        float y[16], sc[16], sh[16];
        tmem_ld16(trow + kColW + 16 * c, y);
        tmem_ld16(trow + kColS + kb * 128 + j, sc);
        tmem_ld16(trow + kColS + kb * 128 + 64 + j, sh);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const float n = fmaf((y[i] - mean) * rstd, st[kStG + 16 * c + i], st[kStB + 16 * c + i]);
            const float v = fmaf(n, sc[i] + st[kStBe + kb * 128 + j + i], sh[i] + st[kStBe + kb * 128 + 64 + j + i]);
            y[i] = silu_f(v);
        }
        store_a16<kBf16>(awork, r, 16 * c, y);
    }
}

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
    TileBarriers* bars = reinterpret_cast<TileBarriers*>(prm_sa + 384);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (a.do_main)
        for (int i = threadIdx.x; i < kPrmFloats; i += kTileThreads) prm[i] = a.prm[i];
    if (a.do_sa1)
        for (int i = threadIdx.x; i < 384; i += kTileThreads) prm_sa[i] = a.prm_next[i];
    tile_setup(bars, warp, lane);
    const uint32_t tmem_base = bars->tmem_base;

    if (warp == 4) {
        if (lane == 0) producer_loop(a.ops, a.n_ops, a.wbuf, a.aemb + (size_t)blockIdx.x * 8 * kStageABytes, ring, bars);
    } else if (warp == 5) {
        if (lane == 0) mma_loop<kBf16>(a.ops, a.n_ops, ring, awork_p, bars, tmem_base);
    } else {
        const uint32_t r = threadIdx.x;
        const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
        const uint32_t awork = smem_u32(awork_p);
        const long g = (long)blockIdx.x * kTileRows + r;
        const bool valid = g < a.M;
        const int b = valid ? (int)(g / a.T) : 0;
        const int t = valid ? (int)(g - (long)b * a.T) : 0;
        uint32_t ph[3] = {0, 0, 0};
        float mean, rstd;

        // ---- residual stream -> TMEM
        {
            const float4* src = reinterpret_cast<const float4*>(a.h + (size_t)g * kD);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 f = valid ? src[4 * c + i] : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * i] = f.x, v[4 * i + 1] = f.y, v[4 * i + 2] = f.z, v[4 * i + 3] = f.w;
                }
                tmem_st16(trow + kColH + 16 * c, v);
            }
            tmem_wait_st();
        }

        if (a.do_main) {
            // ================= self-attention tail: y = q . A_sa ; h += Styl(y)
            {
                const uint4* qrow = reinterpret_cast<const uint4*>(a.q + (size_t)g * kD);
                const float* Ab = a.A_sa + (size_t)b * (kH * 256);
                float sum = 0.f;
#pragma unroll 1
                for (int hh = 0; hh < kH; ++hh) {
                    float qv[16], y[16];
                    uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
                    if (valid) u0 = qrow[2 * hh], u1 = qrow[2 * hh + 1];
                    const uint32_t uu[8] = {u0.x, u0.y, u0.z, u0.w, u1.x, u1.y, u1.z, u1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float2 f = unpack2<kBf16>(uu[i]);
                        qv[2 * i] = f.x, qv[2 * i + 1] = f.y;
                    }
                    head_apply(qv, Ab + hh * 256, y);
#pragma unroll
                    for (int i = 0; i < 16; ++i) sum += y[i];
                    tmem_st16(trow + kColW + 16 * hh, y);
                }
                tmem_wait_st();
                mean = sum * (1.f / kD);
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float v[16];
                    tmem_ld16(trow + kColW + 16 * c, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float d = v[i] - mean;
                        ss = fmaf(d, d, ss);
                    }
                }
                rstd = rsqrtf(ss * (1.f / kD) + kLnEps);
            }
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_sa
            row_film_to_a<kBf16>(trow, mean, rstd, prm + kPrmStSa, awork, r);
            rows_publish(bars);                                          // -> h += A . Wo_sa

            // ================= cross-attention
            rows_wait(bars, 1, ph[1]);
            row_stats(trow + kColH, prm + kPrmStSa + kStBo, mean, rstd);  // h += bo_sa ; LN stats
            row_normalize_to_a<kBf16>(trow + kColH, mean, rstd, awork, r);
            rows_publish(bars);                                          // -> W = LN(h) . Wq_ca
            rows_wait(bars, 2, ph[2]);
            {
                const float* Ab = a.A_ca + (size_t)b * a.a_ca_stride;
                float sum = 0.f;
#pragma unroll 1
                for (int hh = 0; hh < kH; ++hh) {
                    float qv[16], y[16];
                    tmem_ld16(trow + kColW + 16 * hh, qv);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) qv[i] += prm[kPrmCaBq + 16 * hh + i];
                    softmax16(qv);
                    head_apply(qv, Ab + hh * 256, y);
#pragma unroll
                    for (int i = 0; i < 16; ++i) sum += y[i];
                    tmem_st16(trow + kColW + 16 * hh, y);
                }
                tmem_wait_st();
                mean = sum * (1.f / kD);
                float ss = 0.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float v[16];
                    tmem_ld16(trow + kColW + 16 * c, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float d = v[i] - mean;
                        ss = fmaf(d, d, ss);
                    }
                }
                rstd = rsqrtf(ss * (1.f / kD) + kLnEps);
            }
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_ca
            row_film_to_a<kBf16>(trow, mean, rstd, prm + kPrmStCa, awork, r);
            rows_publish(bars);                                          // -> h += A . Wo_ca

            // ================= FFN (no pre-norm, reference transformer.py:170-173)
            rows_wait(bars, 1, ph[1]);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[16];
                tmem_ld16(trow + kColH + 16 * c, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += prm[kPrmStCa + kStBo + 16 * c + i];
                tmem_st16(trow + kColH + 16 * c, v);
                store_a16<kBf16>(awork, r, 16 * c, v);
            }
            tmem_wait_st();
            rows_publish(bars);                                          // -> W[0:64] = h . W1
            rows_wait(bars, 2, ph[2]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[16];
                tmem_ld16(trow + kColW + 16 * c, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = gelu_erf_f(v[i] + prm[kPrmFfB1 + 16 * c + i]);
                store_a16<kBf16>(awork, r, 16 * c, v);
            }
            rows_publish(bars);                                          // -> W = GELU(.) . W2
            rows_wait(bars, 2, ph[2]);
            row_stats(trow + kColW, prm + kPrmFfB2, mean, rstd);          // y = W + b2 ; LN stats
            rows_wait(bars, 0, ph[0]);                                   // S = A_emb . We_ffn
            row_film_to_a<kBf16>(trow, mean, rstd, prm + kPrmStFf, awork, r);
            rows_publish(bars);                                          // -> h += A . Wo_ffn
            rows_wait(bars, 1, ph[1]);
        }

        if (a.do_sa1) {
            // ================= next layer's self-attention head: LN -> q | k | v
            row_stats(trow + kColH, a.do_main ? prm + kPrmStFf + kStBo : nullptr, mean, rstd);
            row_normalize_to_a<kBf16>(trow + kColH, mean, rstd, awork, r);
            rows_publish(bars);
            rows_wait(bars, 2, ph[2]);
            const bool keep = valid && (a.length == nullptr || (long long)t < a.length[b]);
#pragma unroll 1
            for (int hh = 0; hh < kH; ++hh) {
                float qv[16];
                tmem_ld16(trow + kColS + 16 * hh, qv);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) qv[i] += prm_sa[kPrmSaBq + 16 * hh + i];
                softmax16(qv);
                if (valid) {
                    uint32_t p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) p[i] = pack2<kBf16>(qv[2 * i], qv[2 * i + 1]);
                    uint4* dst = reinterpret_cast<uint4*>(a.q + (size_t)g * kD + 16 * hh);
                    dst[0] = make_uint4(p[0], p[1], p[2], p[3]);
                    dst[1] = make_uint4(p[4], p[5], p[6], p[7]);
                }
            }
#pragma unroll 1
            for (int c = 0; c < 16; ++c) {                               // k: cols 0..127 of kv ; v: cols 128..255
                const bool is_v = c >= 8;
                float v[16];
                tmem_ld16(trow + (is_v ? kColW + 16 * (c - 8) : kColS + 128 + 16 * c), v);
                tmem_wait_ld();
                if (valid) {
                    const float* bias = prm_sa + (is_v ? kPrmSaBv + 16 * (c - 8) : kPrmSaBk + 16 * c);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        v[i] += bias[i];
                        if (!keep) v[i] = is_v ? 0.f : v[i] + -1000000.f;   // reference transformer.py:107,114
                    }
                    float4* dst = reinterpret_cast<float4*>(a.kv + (size_t)g * 256 + 16 * c);
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }

        // ---- residual stream -> global (deferred bias of the last FFN block added here if not yet)
        {
            const bool add_bias = a.do_main && !a.do_sa1;
            float4* dst = reinterpret_cast<float4*>(a.h + (size_t)g * kD);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float v[16];
                tmem_ld16(trow + kColH + 16 * c, v);
                tmem_wait_ld();
                if (add_bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += prm[kPrmStFf + kStBo + 16 * c + i];
                }
                if (valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) dst[4 * c + i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            }
        }
    }
    tile_teardown(bars, warp);
}

constexpr int kGemmSmemBytes = kStages * kStageBytes + sizeof(TileBarriers) + 1024;
constexpr int kLayerSmemBytes = kStages * kStageBytes + kAworkBytes + (kPrmFloats + 384) * 4 + sizeof(TileBarriers) + 1024;

}  // namespace dc
