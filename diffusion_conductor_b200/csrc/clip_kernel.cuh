// Persistent sampling-loop kernel, one thread-block CLUSTER per clip.
//
// A clip of T frames is cut into nt = ceil(T / 128) equal tiles of rows_per = ceil(T / nt) frames; CTA `rank` of the
// cluster owns tile `rank` for the WHOLE sampling loop (all steps, all layers): the residual stream stays in TMEM,
// x / q / the FiLM operands stay in the SM.  Tiles never straddle clips, so every MMA serves one clip (no lane
// masks, one K^T V pass) and clips are completely independent: there is no grid-wide co-residency requirement, any
// batch size runs as B clusters which the hardware schedules as SMs free up.
// The one cross-tile dependency of the model -- the time-axis softmax of the self-attention keys and K^T V, a per-clip
// reduction (reference transformer.py:111,117) -- is exchanged through DISTRIBUTED SHARED MEMORY: each CTA leaves its
// partial (column max, column sum, 8 diagonal 16x16 blocks of E^T V) in its own shared memory, signals the peers'
// mbarriers (release.cluster), and every CTA pulls the nt partials with ld.shared::cluster and merges them (online
// softmax rescaling) straight into its compact head-block B-operand image (bdc_offset).  Clips of more than four tiles exchange
// through L2 instead (kGx), two-tile clips push instead of pull (kPush).
#pragma once
#include "tile_kernels.cuh"

namespace dc {

constexpr int kPRingAStages = 2;        // persistent kernel: 2 x 48 KB (the FiLM projection is shared-memory-bandwidth bound)

struct StepArgs {
    int L, M, T;
    const uint8_t* wbuf;        // packed weights [L][1 MiB]
    const uint8_t* aemb;        // A_emb image [B * nt tiles][8][16 KB]
    const float* prm;           // [L][kPrmFloats] (clip variant: FFN-up bias and deferred output biases folded, see dc_api.cu)
    const uint8_t* wfuse;       // [L][16 KB] W1 . Wo_ca as a [64 x 128] K-major operand image (FFN-up fused across the residual add)
    // step prologue / epilogue fused into the kernel (reference transformer.py:482,488-490,496; gaussian_diffusion.py:812-830)
    // One launch runs n_steps consecutive denoise steps (timestep indices step0, step0 - 1, ...): a tile's x rows are
    // private to its CTA, so the only cross-CTA traffic of the whole sampling loop is the per-clip partial exchange.
    int n_steps, step0;
    const float* x_in;          // [M][26] current sample x_t
    float* x_out;               // [M][26] updated sample (may alias x_in; null when mode == 0; must be set when n_steps > 1)
    float* x0_out;              // [M][26] pred_xstart (model output) of launch step i at x0_out + i * x0_stride
    size_t x0_stride;
    float* x_trace;             // optional: x after launch step i -> x_trace + i * M * 26
    const float* noise;         // [M][26] (+ i * noise_stride for launch step i) or null
    size_t noise_stride;
    const float* xp;            // [M][512] linear(xf_proj)
    const float* te;            // time embedding row(s): te + timestep * te_step_stride + b * te_stride
    int te_stride, te_step_stride;
    const float* coef;          // [S][8] update coefficients, row = timestep index; null when mode == 0
    int mode;                   // 0: model output only, 1: DDIM, 2: DDPM, | 0x10 clamp
    const uint8_t* wj_img;      // joint_embed as a [128 x 128] K-major operand image: k 0..25 = W_hi, 32..57 = W_hi, 64..89 = W_lo (x is fed as hi | lo | hi)
    const float* bj;            // [128]
    const float* pos;           // [num_frames][128] sequence_embedding
    const uint8_t* wout_img;    // output head `out` as a [32 x 384] K-major operand image [W_hi | W_hi | W_lo] (rows >= 26 zero), 24 KB
    const float* bo;            // [32]
    uint8_t* aemb_out;          // == aemb: this CTA writes its own tile's A_emb image first
    size_t aemb_stride;         // bytes between the two A_emb images (step parity): the image of step i + 1 is built during step i
    const uint8_t* bd_ca;       // cross-attention images, compact 4 KB head-block form (bdc_offset): clip stride bd_ca_stride, layer stride kBdcBytes
    size_t bd_ca_stride;
    const long long* length;    // [B] or null
    uint32_t off[12];           // byte offsets of the packed matrices inside a layer slab (see dc_api.cu)
    unsigned long long* timeline;
    int dbg;                    // bit 0: skip the FiLM projection MMAs (timing experiments only; results are wrong)
    int nt, rows_per;           // tiles (= cluster size) per clip, frames per tile
    const float* kshift;        // [L][128] static softmax shift of the self-attention keys (upper bound of |k|, see dc_api.cu)
    uint32_t static_mask;       // bit l set: layer l uses the static shift (no column-max pass)
    // Exchange of the per-clip partials through GLOBAL memory instead of distributed shared memory (gx = 1; launched with
    // cluster size 1).  Clusters of 9..16 CTAs fit only once per GPC (7 x 15 = 105 of the 148 SMs busy on a 1800-frame batch);
    // without the cluster constraint every SM takes a tile.  Reduce-scatter + all-gather as in the large-cluster path; every
    // 8-byte word carries its own flag (value | tag, tag = gx_tag0 + reduction sequence number + 1, unique per reduction and
    // launch), so a reader simply re-reads until the tag is there: no fence, no separate flag, one L2 round trip per phase.
    int gx;
    uint2* gx_part;             // [B][2][nt][kKvPartFloats] (value, tag) partials, double-buffered by reduction parity
    uint2* gx_slice;            // [B][2][128][8] (two 16-bit merged attention values, tag)
    uint32_t gx_tag0;
};
enum { kOWeSa = 0, kOWoSa, kOWeCa, kOWqCa, kOWoCa, kOWeFf, kOW1, kOW2, kOWoFf, kOWq, kOWk, kOWv };

// Debug timeline of CTA 0 (dc_debug_timeline): every marking thread (row thread 0, the two MMA issuers) appends
// (clock64, id) pairs to its OWN lane of the buffer with plain stores -- no atomics, so a mark costs a few cycles.
// Layout: [3 lanes][1 + 2 * kTlEvents] u64, word 0 of a lane = number of events.
constexpr int kTlEvents = 680;
template <bool kOn>
struct Timeline {
    unsigned long long* p;
    uint32_t n;
    __device__ __forceinline__ Timeline(const StepArgs& a, int lane_id, bool writer = true)
        : p(kOn && writer && a.timeline != nullptr && blockIdx.x == 0 ? a.timeline + (size_t)lane_id * (1 + 2 * kTlEvents) : nullptr), n(0) {}
    __device__ __forceinline__ void mark(unsigned long long id) {
        if constexpr (kOn) {
            if (p != nullptr && n < (uint32_t)kTlEvents) {
                p[1 + 2 * n] = (unsigned long long)clock64();
                p[2 + 2 * n] = id;
                ++n;
            }
        }
    }
    __device__ __forceinline__ void finish() {
        if constexpr (kOn) {
            if (p != nullptr) p[0] = n;
        }
    }
};

constexpr int kClipRedFloats = 128 + 128 + 8;   // column maxima, column sums (the 4 KB scan partials alias the LayerNorm exchange buffer)
constexpr uint32_t kRingAItems = 26;    // ring-A items per layer: 3 x 8 FiLM stages + Wk + Wv
constexpr int kMaxClipTiles = 16;       // cluster size limit (non-portable); T <= 16 * 128
constexpr int kDirectMergeTiles = 4;    // up to this cluster size every CTA pulls every partial; beyond: reduce-scatter + all-gather
constexpr int kGxMinTiles = 7;          // from this many tiles per clip on: independent CTAs, exchange through L2 (StepArgs::gx)

struct ClipBarriers : LayerBarriers {
    uint64_t part_ready[2];             // peers -> this CTA: "my partial of reduction seq is in my shared memory" (count nt - 1)
    uint64_t pull_done[2];              // peers -> this CTA: "I have finished reading your partial" (count nt - 1)
    uint64_t slice_ready[2];            // large clusters: "my merged slice of reduction seq is ready" (count nt - 1)
    uint64_t w1c_full;                  // the (W1 Wo_ca) image of the current layer has landed
    uint64_t w_free;                    // row threads -> MMA issuer: y_ca has been read out of W, h16 . W1 may start (16 warp arrivals)
    uint64_t recv_free[2];              // kPush: peer -> this CTA: "my W1 Wo_ca buffer may receive your partial of reduction seq" (count 1)
};

// Merge of the per-clip partials through L2 (StepArgs::gx; only the kGx instantiation of the kernel contains it, so that the
// cluster variant keeps its register allocation).
template <bool kBf16>
__device__ __forceinline__ void gx_merge(const uint2* pbase, uint2* slices, float4* comb /* [G <= 8][per] (M, sum, acc0, acc1) */, uint8_t* xbuf, int nt,
                                         int rank, int tx, uint32_t gtag, bool static_shift) {
    // reduce-scatter + all-gather through L2.  CTA `rank` merges key features [rank fs, rank fs + fs) of all nt partials:
    // thread (g, dl, l2) takes the tiles j = g, g + G, ... (at most 3) of feature rank fs + dl, value columns 2 l2, 2 l2 + 1;
    // the G groups combine through shared memory (ring B is idle here); the merged rows go out as (two 16-bit values, tag)
    // words; then everyone gathers the 128 x 16 merged matrix.
    const int fs = (kD + nt - 1) / nt;
    const int per = fs * 8;                                      // threads per tile group
    const int G = min(8, kRowThreads / per);                     // nt <= 16 -> ceil(nt / G) <= 3: one round trip
    const int g = tx / per, idx = tx - g * per, dl = idx >> 3, l2 = idx & 7, d = rank * fs + dl;
    const bool act = g < G && d < kD;
    float Mx = -INFINITY, ss = 0.f, a0 = 0.f, a1 = 0.f;
    if (act) {
        // three tiles per round trip (registers), online-softmax accumulation across round trips
        for (int k0 = 0; k0 < 3; k0 += 3) {
            if (g + k0 * G >= nt) break;
            uint4 pw[3];
            uint2 sw[3], mw[3];
            bool ok;
            do {
                ok = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const int j = g + (k0 + k) * G;
                    if (j < nt) {
                        const uint2* pp = pbase + (size_t)j * kKvPartFloats;
                        pw[k] = ld_volatile_v4(reinterpret_cast<const uint4*>(pp + 256 + l2 * 256 + d * 2));
                        sw[k] = ld_volatile_v2(pp + 128 + d);
                        mw[k] = static_shift ? make_uint2(0u, gtag) : ld_volatile_v2(pp + d);
                        ok = ok && pw[k].y == gtag && pw[k].w == gtag && sw[k].y == gtag && mw[k].y == gtag;
                    }
                }
            } while (!ok);
#pragma unroll
            for (int k = 0; k < 3; ++k)
                if (g + (k0 + k) * G < nt) {
                    if (static_shift) {
                        Mx = 0.f;
                        ss += __uint_as_float(sw[k].x), a0 += __uint_as_float(pw[k].x), a1 += __uint_as_float(pw[k].z);
                    } else {
                        const float m = __uint_as_float(mw[k].x), nm = fmaxf(Mx, m);
                        const float cs = __expf(Mx - nm), w = __expf(m - nm);          // exp(-inf) = 0 on the first tile
                        ss = fmaf(__uint_as_float(sw[k].x), w, ss * cs);
                        a0 = fmaf(__uint_as_float(pw[k].x), w, a0 * cs), a1 = fmaf(__uint_as_float(pw[k].z), w, a1 * cs);
                        Mx = nm;
                    }
                }
        }
        if (G > 1) comb[g * per + idx] = make_float4(Mx, ss, a0, a1);
    }
    if (G > 1) {
        named_bar_sync(5, kRowThreads);
        if (act && g == 0) {
            for (int k = 1; k < G; ++k) {
                const float4 c = comb[k * per + idx];
                if (static_shift) {
                    ss += c.y, a0 += c.z, a1 += c.w;
                } else {
                    const float nm = fmaxf(Mx, c.x);
                    const float cs = __expf(Mx - nm), w = __expf(c.x - nm);
                    ss = fmaf(c.y, w, ss * cs), a0 = fmaf(c.z, w, a0 * cs), a1 = fmaf(c.w, w, a1 * cs);
                    Mx = nm;
                }
            }
        }
    }
    if (act && g == 0) {
        const float inv = ss > 0.f ? 1.f / ss : 0.f;
        st_global_v2(slices + d * 8 + l2, make_uint2(pack2<kBf16>(a0 * inv, a1 * inv), gtag));
    }
    {
        const int dd = tx >> 2, q4 = tx & 3;
        uint4 w;
        do w = ld_volatile_v4(reinterpret_cast<const uint4*>(slices + dd * 8 + 2 * q4));
        while (w.y != gtag || w.w != gtag);
        const uint16_t vals[4] = {(uint16_t)(w.x & 0xFFFFu), (uint16_t)(w.x >> 16), (uint16_t)(w.z & 0xFFFFu), (uint16_t)(w.z >> 16)};
#pragma unroll
        for (int i = 0; i < 4; ++i)
            *reinterpret_cast<uint16_t*>(xbuf + bdc_offset((uint32_t)(dd >> 4), (uint32_t)(dd & 15), (uint32_t)(4 * q4 + i))) = vals[i];
    }
}

// kTl: instrumented build for dc_debug_timeline (the marks cost ~5 % of the instruction stream, so the production
// instantiation compiles them out)
// kPush (two-tile clips, cluster of 2): each CTA writes its partial straight into the peer's shared memory (the W1 Wo_ca buffer,
// idle between its MMA and the next refill; the peer grants it with a credit) and the merge reads local shared memory only:
// one DSMEM latency on the critical path instead of the arrive + pull round trips.
template <bool kBf16, bool kTl = false, bool kGx = false, bool kPush = false>
__global__ void __launch_bounds__(kTileThreads, 1) clip_kernel(const __grid_constant__ StepArgs a) {
    constexpr int kNA = kPRingAStages, kSA = kStageBytes, kNB = kRingBStages, kSB = kRingBStageBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ringA = smem;
    uint8_t* ringB = ringA + kNA * kSA;                               // also the V image, then this tile's partial (pulled by the peers)
    uint8_t* awork_p = ringB + kNB * kSB;
    uint8_t* xbuf = awork_p + kAworkBytes;                            // k / E image of the reduction, then the merged attention image
    uint8_t* w1c = xbuf + kAworkBytes;                                // [16 KB] (W1 Wo_ca) operand image of the current layer (1024-B aligned)
    float* prm = reinterpret_cast<float*>(w1c + 16384);               // [kPrmFloats] layer `it`
    float* prm_sa = prm + kPrmFloats;                                 // [512] SA biases bq | bk | bv and the static key shift of layer it + 1
    float2* xchg = reinterpret_cast<float2*>(prm_sa + 512);           // [4][128] LayerNorm exchange; column-scan partials during the reduction
    float* red = reinterpret_cast<float*>(xchg + 512);                // kClipRedFloats
    ClipBarriers* bars = reinterpret_cast<ClipBarriers*>(red + kClipRedFloats);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = a.L;
    // clip-aligned tiling: cluster = clip, rank = tile of the clip
    const int nt = a.nt;
    const int rank = (int)blockIdx.x % nt;
    const int clip = (int)blockIdx.x / nt;
    const int t0 = rank * a.rows_per;                                 // first frame of this tile
    const int nrows = max(0, min(a.rows_per, a.T - t0));              // valid rows (TMEM lanes) of this tile
    const long row0g = (long)clip * a.T + t0;                         // first token (global row) of this tile

    pdl_trigger();
    if (warp == kProducerWarp && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars->fullA[i]), 1), mbar_init(smem_u32(&bars->emptyA[i]), 1);
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars->fullB[i]), 1), mbar_init(smem_u32(&bars->emptyB[i]), 1);
        mbar_init(smem_u32(&bars->a_ready), kRowWarps);
        mbar_init(smem_u32(&bars->s_free), kRowWarps);
        mbar_init(smem_u32(&bars->q_full), 1);
        mbar_init(smem_u32(&bars->aemb_ready), kRowWarps);
        for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&bars->d_ready[i]), 1);
        mbar_init(smem_u32(&bars->w1c_full), 1);
        mbar_init(smem_u32(&bars->w_free), kRowWarps);
        for (int i = 0; i < 2; ++i) {
            mbar_init(smem_u32(&bars->part_ready[i]), (uint32_t)max(nt - 1, 1));
            mbar_init(smem_u32(&bars->pull_done[i]), (uint32_t)max(nt - 1, 1));
            mbar_init(smem_u32(&bars->slice_ready[i]), (uint32_t)max(nt - 1, 1));
            mbar_init(smem_u32(&bars->recv_free[i]), 1);
        }
        mbar_fence_init();
    }
    if (warp == kMmaWarp) {
        tmem_alloc(smem_u32(&bars->tmem_base), 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                                               // every peer's barriers exist before anyone arrives remotely
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();

    if (warp == kProducerWarp) {
        // ---------------- ring A: per layer 3 FiLM projections (8 stages each), then Wk, Wv of the next layer
        //                  (kRingAItems = 26 items per layer; item order is what the two consumers below assume)
        if (lane == 0) {
            const uint8_t* a_img0 = a.aemb + (size_t)blockIdx.x * 8 * kStageABytes;
            uint32_t it_ = 0;
            auto stage_in = [&](const uint8_t* a_src, const uint8_t* w_src, uint32_t w_bytes) {
                const uint32_t st = it_ % kNA, ph = (it_ / kNA) & 1u;
                ++it_;
                mbar_wait(smem_u32(&bars->emptyA[st]), ph ^ 1u);
                const uint32_t full = smem_u32(&bars->fullA[st]);
                uint8_t* stage = ringA + st * kSA;
                mbar_arrive_expect_tx(full, w_bytes + (a_src ? kStageABytes : 0));
                if (a_src) bulk_g2s(smem_u32(stage), a_src, kStageABytes, full);
                bulk_g2s(smem_u32(stage + kStageABytes), w_src, w_bytes, full);
            };
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = -1; it < L; ++it) {
                const uint8_t* a_img = a_img0 + (kGx ? (size_t)(si & 1) * a.aemb_stride : (size_t)0);   // (one image is enough when it is rebuilt in place)
                if (it >= 0) {
                    if (it == 0) {
                        mbar_wait(smem_u32(&bars->aemb_ready), (uint32_t)si & 1u);   // the row threads have written this step's A_emb image
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const uint8_t* slab = a.wbuf + ((size_t)it << 20);
                    const uint32_t so[3] = {a.off[kOWeSa], a.off[kOWeCa], a.off[kOWeFf]};
                    for (int o = 0; o < 3; ++o) {
                        for (int s = 0; s < kSopStages; ++s) stage_in(a_img + (size_t)s * kStageABytes, slab + so[o] + (size_t)s * kStageWBytes, kStageWBytes);
                    }
                }
                if (it + 1 < L) {
                    const uint8_t* slab = a.wbuf + ((size_t)(it + 1) << 20);
                    stage_in(nullptr, slab + a.off[kOWk], 32768);
                    stage_in(nullptr, slab + a.off[kOWv], 32768);
                }
            }
        }
    } else if (warp == kProducerBWarp) {
        // ---------------- ring B (one k-block per stage): dependent-GEMM weights and the clip's cross-attention image
        if (lane == 0) {
            uint32_t it_ = 0;
            auto load = [&](const uint8_t* src, int kb, uint32_t kb_bytes) {
                for (int k = 0; k < kb; ++k, ++it_) {
                    const uint32_t st = it_ % kNB, ph = (it_ / kNB) & 1u;
                    mbar_wait(smem_u32(&bars->emptyB[st]), ph ^ 1u);
                    const uint32_t full = smem_u32(&bars->fullB[st]);
                    mbar_arrive_expect_tx(full, kb_bytes);
                    bulk_g2s(smem_u32(ringB + st * kSB), src + (size_t)k * kb_bytes, kb_bytes, full);
                }
            };
            uint32_t qf = 0;
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = -1; it < L; ++it) {
                if (it < 0) load(a.wj_img, 2, 16384);                             // joint_embed [W_hi | W_hi], [W_lo | 0]: first GEMM of the step
                if (it >= 0) {
                    const uint8_t* slab = a.wbuf + ((size_t)it << 20);
                    // ring B held the V image and then this tile's partial of the reduction that opened layer `it`:
                    // refill only when the local q . blockdiag(A_sa) has been issued AND every peer has pulled the partial
                    mbar_wait(smem_u32(&bars->q_full), qf++ & 1u);
                    if (!kGx && !kPush && nt > 1) {
                        const uint32_t seq = (uint32_t)(si * L + it);
                        mbar_wait(smem_u32(&bars->pull_done[seq & 1u]), (seq >> 1) & 1u);
                    }
                    {   // (W1 Wo_ca) of this layer -> its own buffer (the previous layer's MMAs on it completed before q_full fired)
                        const uint32_t full = smem_u32(&bars->w1c_full);
                        mbar_arrive_expect_tx(full, 16384);
                        bulk_g2s(smem_u32(w1c), a.wfuse + (size_t)it * 16384, 16384, full);
                    }
                    load(slab + a.off[kOWoSa], 2, 16384);
                    load(slab + a.off[kOWqCa], 2, 16384);
                    load(a.bd_ca + (size_t)clip * a.bd_ca_stride + (size_t)it * kBdcBytes, 1, kBdcBytes);      // eight 16 x 16 head blocks
                    load(slab + a.off[kOW1], 1, 16384);                            // both 8 KB k-blocks of W1 in one stage
                    load(slab + a.off[kOWoCa], 2, 16384);
                    load(slab + a.off[kOW2], 1, 16384);
                    load(slab + a.off[kOWoFf], 2, 16384);
                    if (it + 1 == L) load(a.wout_img, 1, 16384), load(a.wout_img + 16384, 1, 8192);   // output head [W_hi | W_hi], [W_lo]: 4 + 2 k-blocks of 4 KB
                }
                if (it + 1 < L) load(a.wbuf + ((size_t)(it + 1) << 20) + a.off[kOWq], 2, 16384);
            }
        }
    } else if (warp == kRelayWarp) {
        // ---------------- issuer of the FiLM projections S = A_emb . We (ring A), gated only by s_free
        if (lane == 0) {
            Timeline<kTl> tl(a, 2);
            const uint32_t idesc_s = make_idesc<kBf16>(kTileRows, 256);
            uint32_t sj = 0;
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = 0; it < L; ++it) {
                // ring-A items before this layer: 26 L per earlier step, 2 (Wk,Wv of layer 0), 26 per layer
                const uint32_t base_it = (uint32_t)si * kRingAItems * (uint32_t)L + (uint32_t)it * kRingAItems + 2;
                for (int o = 0; o < 3; ++o, ++sj) {
                    mbar_wait(smem_u32(&bars->s_free), sj & 1u);
                    tc_fence_after();
                    for (int sgi = 0; sgi < kSopStages; ++sgi) {
                        const uint32_t itA = base_it + o * kSopStages + sgi;
                        const uint32_t st = itA % kNA, ph = (itA / kNA) & 1u;
                        mbar_wait(smem_u32(&bars->fullA[st]), ph);
                        tc_fence_after();
                        const uint32_t stage = smem_u32(ringA + st * kSA);
                        if (!(a.dbg & 1)) umma_kblock(tmem_base + kColS, stage, stage + kStageABytes, idesc_s, sgi > 0);
                        umma_commit(smem_u32(&bars->emptyA[st]));
                    }
                    umma_commit(smem_u32(&bars->d_ready[0]));
                    tl.mark(310 + o);
                }
            }
            tl.finish();
        }
    } else if (warp == kMmaWarp) {
        // ---------------- issuer of the dependent GEMMs, strictly in chain order with blocking waits
        if (lane == 0) {
            Timeline<kTl> tl(a, 1);
            const uint32_t awork = smem_u32(awork_p);
            uint32_t itB = 0, a_phase = 0;
            auto wait_a = [&]() {
                mbar_wait(smem_u32(&bars->a_ready), a_phase & 1u);
                ++a_phase;
                tc_fence_after();
                tl.mark(400);
            };
            // B operand through ring B, one k-block per stage
            auto gemm_b = [&](int kb, int n, uint32_t d_col, bool acc, uint32_t a_base) {
                const uint32_t idesc = make_idesc<kBf16>(kTileRows, n);
                for (int k = 0; k < kb; ++k, ++itB) {
                    const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                    mbar_wait(smem_u32(&bars->fullB[st]), ph);
                    tc_fence_after();
                    tl.mark(401);
                    umma_kblock(tmem_base + d_col, a_base + k * kABlockBytes, smem_u32(ringB + st * kSB), idesc, acc || k > 0);
                    umma_commit(smem_u32(&bars->emptyB[st]));
                }
            };
            auto done = [&](int which) { umma_commit(smem_u32(&bars->d_ready[which])); };
            const uint32_t idmn = make_idesc_mn<kBf16>(kTileRows, kTileRows);
            const uint32_t idesc128 = make_idesc<kBf16>(kTileRows, 128);
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = -1; it < L; ++it) {
                if (it < 0) wait_a(), gemm_b(2, 128, kColH, false, awork), done(1), tl.mark(211);   // h0 = [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T
                if (it >= 0) {
                    wait_a();                                                      // merged attention image written by the row threads
                    {   // y = q . blockdiag(A_sa): one N = 16 MMA per head on the compact merged image
                        const uint32_t idesc16 = make_idesc<kBf16>(kTileRows, 16);
#pragma unroll
                        for (int hh = 0; hh < kH; ++hh)
                            umma_f16(tmem_base + kColW + 16 * hh, make_desc_kmajor_sw128(awork + (hh >> 2) * kABlockBytes) + 2 * (hh & 3),
                                     make_desc_kmajor_sw128(smem_u32(xbuf) + (hh >> 2) * 2048) + 2 * (hh & 3), idesc16, 0u);
                    }
                    done(2), tl.mark(200);
                    umma_commit(smem_u32(&bars->q_full));
                    wait_a(), gemm_b(2, 128, kColH, true, awork), done(1), tl.mark(201);   // h += . Wo_sa
                    wait_a(), gemm_b(2, 128, kColW, false, awork), done(2), tl.mark(202);  // q_ca
                    {   // y = softmax(q) . blockdiag(A_ca): one N = 16 MMA per head on the compact image (one ring-B stage)
                        wait_a();
                        const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                        ++itB;
                        mbar_wait(smem_u32(&bars->fullB[st]), ph);
                        tc_fence_after();
                        tl.mark(401);
                        const uint32_t b_base = smem_u32(ringB + st * kSB), idesc16 = make_idesc<kBf16>(kTileRows, 16);
#pragma unroll
                        for (int hh = 0; hh < kH; ++hh)
                            umma_f16(tmem_base + kColW + 16 * hh, make_desc_kmajor_sw128(awork + (hh >> 2) * kABlockBytes) + 2 * (hh & 3),
                                     make_desc_kmajor_sw128(b_base + (hh >> 2) * 2048) + 2 * (hh & 3), idesc16, 0u);
                        umma_commit(smem_u32(&bars->emptyB[st]));
                        done(2), tl.mark(203);
                    }
                    // The FFN has no pre-norm, so its up-projection is linear in the residual add before it:
                    //   u = (h + a . Wo_ca + bo) . W1 = h16 . W1 + a . (W1 Wo_ca) + const.
                    // h16 . W1 is issued as soon as the row threads have read y_ca out of W (while they do the FiLM math);
                    // h += a . Wo_ca and the a . (W1 Wo_ca) half follow when a is published: one round trip less per layer.
                    {
                        const uint32_t idesc64 = make_idesc<kBf16>(kTileRows, 64);
                        mbar_wait(smem_u32(&bars->w_free), (uint32_t)(si * L + it) & 1u);
                        tc_fence_after();
                        const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                        ++itB;
                        mbar_wait(smem_u32(&bars->fullB[st]), ph);
                        tc_fence_after();
                        const uint32_t b_base = smem_u32(ringB + st * kSB);
                        for (int k = 0; k < 2; ++k) umma_kblock(tmem_base + kColW, smem_u32(xbuf) + k * kABlockBytes, b_base + k * 8192, idesc64, k > 0);
                        umma_commit(smem_u32(&bars->emptyB[st]));
                        wait_a(), gemm_b(2, 128, kColH, true, awork), tl.mark(204);
                        mbar_wait(smem_u32(&bars->w1c_full), (uint32_t)(si * L + it) & 1u);
                        tc_fence_after();
                        for (int k = 0; k < 2; ++k) umma_kblock(tmem_base + kColW, awork + k * kABlockBytes, smem_u32(w1c) + k * 8192, idesc64, true);
                    }
                    done(2), tl.mark(205);
                    wait_a(), gemm_b(1, 128, kColW, false, awork), done(2), tl.mark(206);  // FFN down
                    wait_a(), gemm_b(2, 128, kColH, true, awork), done(1), tl.mark(207);   // h += . Wo_ffn
                }
                if (it + 1 == L) {
                    // output head (reference :496): pred_x0 = [h_hi | h_lo | h_hi] . [W_hi | W_hi | W_lo]^T, N = 32, K = 384: both the
                    // final residual stream and the weights enter as (hi, lo) pairs of 16-bit values (fp32-equivalent to 2^-16)
                    const uint32_t idesc32 = make_idesc<kBf16>(kTileRows, 32);
                    wait_a();
                    const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                    ++itB;
                    mbar_wait(smem_u32(&bars->fullB[st]), ph);
                    tc_fence_after();
                    const uint32_t b_base = smem_u32(ringB + st * kSB);
                    for (int k = 0; k < 4; ++k) umma_kblock(tmem_base + kColW, awork + k * kABlockBytes, b_base + k * 4096, idesc32, k > 0);   // awork | xbuf
                    umma_commit(smem_u32(&bars->emptyB[st]));
                    {   // + h_hi . W_lo (the hi image again, against the low half of the weights)
                        const uint32_t st2 = itB % kNB, ph2 = (itB / kNB) & 1u;
                        ++itB;
                        mbar_wait(smem_u32(&bars->fullB[st2]), ph2);
                        tc_fence_after();
                        const uint32_t b2 = smem_u32(ringB + st2 * kSB);
                        for (int k = 0; k < 2; ++k) umma_kblock(tmem_base + kColW, awork + k * kABlockBytes, b2 + k * 4096, idesc32, true);
                        umma_commit(smem_u32(&bars->emptyB[st2]));
                    }
                    done(2), tl.mark(210);
                }
                if (it + 1 < L) {
                    wait_a();
                    gemm_b(2, 128, kColS, false, awork);                           // q -> S[0:128]
                    const uint32_t itA0 = (uint32_t)si * kRingAItems * (uint32_t)L + (uint32_t)(it + 1) * kRingAItems;   // Wk, Wv of layer it+1 in ring A
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t itA = itA0 + j;
                        const uint32_t st = itA % kNA, ph = (itA / kNA) & 1u;
                        mbar_wait(smem_u32(&bars->fullA[st]), ph);
                        tc_fence_after();
                        const uint32_t b_base = smem_u32(ringA + st * kSA + kStageABytes);
                        for (int k = 0; k < 2; ++k)
                            umma_kblock(tmem_base + (j == 0 ? kColS + 128 : kColW), awork + k * kABlockBytes, b_base + k * 16384, idesc128, k > 0);
                        umma_commit(smem_u32(&bars->emptyA[st]));
                    }
                    done(2), tl.mark(208);
                    wait_a();                                                      // K^T V partial: E^T . V, MN-major images
                    const uint32_t eimg = smem_u32(xbuf), vimg = smem_u32(ringB);
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(tmem_base + kColW, make_desc_mnmajor_sw128(eimg + ks * 2048), make_desc_mnmajor_sw128(vimg + ks * 2048), idmn, ks > 0);
                    done(2), tl.mark(209);
                }
            }
            tl.finish();
        }
    } else {
        const uint32_t a_ready_addr = smem_u32(&bars->a_ready);
        const uint32_t s_free_addr = smem_u32(&bars->s_free);
        const uint32_t lq = warp & 3, cq = warp >> 2;
        const uint32_t r = lq * 32 + lane;            // row of the tile == TMEM lane
        const uint32_t c0 = cq * 32;                  // first of this thread's 32 features (heads 2cq, 2cq+1)
        const uint32_t trow = tmem_base + ((lq * 32) << 16);
        const uint32_t awork = smem_u32(awork_p);
        const bool valid = (int)r < nrows;
        const int t = valid ? t0 + (int)r : 0;
        const bool keep = valid && (a.length == nullptr || (long long)t < a.length[clip]);
        uint32_t ph[3] = {0, 0, 0};
        Timeline<kTl> tl(a, 0, threadIdx.x == 0);
        tl.mark(1);
        RowStats rs{xchg, 1 + lq, r, cq, 0};
        float mean, rstd;
        float v[32];
        // One 8-feature chunk (task = chunk * 512 + thread: row = task / 64, features 8 (task % 64) ..) of the A_emb image
        // SiLU(time embedding of `tstep_img` + linear(xf_proj)) of this tile -> image `img`.  The first step of a launch builds all 16
        // chunks per thread in its prologue; every later step finds its image ready, built two chunks per layer inside the MMA wait
        // windows of the step before (the whole build in one place was 13 k cycles of every step's critical path).
        auto aemb_chunk = [&](int chunk, int tstep_img, uint8_t* img) {
            const int task = chunk * kRowThreads + (int)threadIdx.x;
            const int row = task >> 6, ch = task & 63;
            const long gg = row0g + min(row, max(nrows - 1, 0));
            const float4* xr4 = reinterpret_cast<const float4*>(a.xp + gg * kE + ch * 8);
            const float4* tr4 = reinterpret_cast<const float4*>(a.te + (size_t)tstep_img * a.te_step_stride + (size_t)clip * a.te_stride + ch * 8);
            const float4 a0 = __ldg(xr4), a1 = __ldg(xr4 + 1), t0v = __ldg(tr4), t1v = __ldg(tr4 + 1);
            const float e8[8] = {a0.x + t0v.x, a0.y + t0v.y, a0.z + t0v.z, a0.w + t0v.w, a1.x + t1v.x, a1.y + t1v.y, a1.z + t1v.z, a1.w + t1v.w};
            uint32_t p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) p[i] = pack2<kBf16>(silu_f<kBf16>(e8[2 * i]), silu_f<kBf16>(e8[2 * i + 1]));
            const uint4 pk = row < nrows ? make_uint4(p[0], p[1], p[2], p[3]) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(img + (size_t)(ch >> 3) * kABlockBytes + sw128_offset(row, ch & 7)) = pk;
        };
        const int chunks_per_window = (16 + 2 * L - 1) / (2 * L);          // two windows per layer carry the next step's 16 chunks

        for (int si = 0; si < a.n_steps; ++si) {
        // (only where it pays: long clips, +2 %; on two-tile clips the same slices cost 2 - 7 % -- their MMA wait windows are not idle enough)
        uint8_t* img_next = (kGx && si + 1 < a.n_steps) ? a.aemb_out + (size_t)((si + 1) & 1) * a.aemb_stride + (size_t)blockIdx.x * 8 * kStageABytes : nullptr;
        auto next_chunks = [&](int window) {                               // window = 2 it, 2 it + 1
            if (img_next == nullptr) return;
            for (int c = window * chunks_per_window; c < min(16, (window + 1) * chunks_per_window); ++c) aemb_chunk(c, a.step0 - si - 1, img_next);
            // the linear(xf_proj) values of the NEXT window's chunks -> L1 now: an L2 round trip behind the weight stream is 1.5 - 2 k cycles
            for (int c = (window + 1) * chunks_per_window; c < min(16, (window + 2) * chunks_per_window); ++c) {
                const int task = c * kRowThreads + (int)threadIdx.x;
                const long gg = row0g + min(task >> 6, max(nrows - 1, 0));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(a.xp + gg * kE + (task & 63) * 8));
            }
        };
        const int tstep = a.step0 - si;                                  // timestep index of this step
        const float* x_src = si == 0 ? a.x_in : a.x_out;
        // ---- step prologue (reference transformer.py:482,488-490): h0 = joint_embed(x) + sequence_embedding -> TMEM (stays
        //      there for the whole step).  joint_embed runs on the tensor core, but neither x nor W is rounded: both enter as
        //      (hi, lo) pairs of 16-bit values, [x_hi | x_lo | x_hi] . [W_hi | W_hi | W_lo]^T over two 64-wide k-blocks, which is
        //      fp32-equivalent to 2^-16.  This tile's A_emb image follows after the first LayerNorm (below).
        {
            float* ste = reinterpret_cast<float*>(xbuf) + kP * kD + kD + kTileRows * kP;   // [512] time embedding of the clip (read by the A_emb build)
            const int tx = threadIdx.x;
            const int p0 = (int)cq * 8, np = valid ? min(8, kP - p0) : 0;                  // this thread's slice of its x row
            float xv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) xv[i] = i < np ? __ldcg(x_src + (size_t)(row0g + r) * kP + p0 + i) : 0.f;
            const float tt = __ldcg(a.te + (size_t)tstep * a.te_step_stride + (size_t)clip * a.te_stride + tx);
            const float4 psa = tx < 96 ? __ldg(reinterpret_cast<const float4*>(a.prm) + tx)                                      // SA biases of layer 0
                               : tx < 128 ? __ldg(reinterpret_cast<const float4*>(a.kshift) + (tx - 96)) : make_float4(0.f, 0.f, 0.f, 0.f);
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                hi[i] = pack2<kBf16>(xv[2 * i], xv[2 * i + 1]);
                const float2 back = unpack2<kBf16>(hi[i]);
                lo[i] = pack2<kBf16>(xv[2 * i] - back.x, xv[2 * i + 1] - back.y);
            }
            *reinterpret_cast<uint4*>(awork_p + sw128_offset(r, cq)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(awork_p + sw128_offset(r, cq + 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<uint4*>(awork_p + kABlockBytes + sw128_offset(r, cq)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);      // x_hi again, against W_lo
            *reinterpret_cast<uint4*>(awork_p + kABlockBytes + sw128_offset(r, cq + 4)) = make_uint4(0, 0, 0, 0);
            ste[tx] = tt;
            if (tx < 128) reinterpret_cast<float4*>(prm_sa)[tx] = psa;
            rows_publish<false>(a_ready_addr, lane);
            tl.mark(128);
            // sequence embedding + bias of this thread's 32 features, in flight while the GEMM runs
            float pb[32];
            {
                const float4* ps4 = reinterpret_cast<const float4*>(a.pos + (size_t)t * kD + c0);
                const float4* bj4 = reinterpret_cast<const float4*>(a.bj + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 pv = valid ? __ldg(ps4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 bv = __ldg(bj4 + i);
                    pb[4 * i] = pv.x + bv.x, pb[4 * i + 1] = pv.y + bv.y, pb[4 * i + 2] = pv.z + bv.z, pb[4 * i + 3] = pv.w + bv.w;
                }
            }
            rows_wait(bars, 1, ph[1]);
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += pb[i];
            tmem_st32(trow + kColH + c0, v);
            tmem_wait_st();
            tl.mark(129);
        }

        for (int it = -1; it < L; ++it) {
            if (it >= 0) {
            // ================= self-attention tail: y = q . blockdiag(A_sa) (tensor cores) ; h += Styl(y)
                rows_wait(bars, 2, ph[2]); tl.mark(101);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                row_stats32<false>(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); tl.mark(102);                                   // S = A_emb . We_sa
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStSa, awork, r, c0, s_free_addr, lane);   // releases S as soon as it has been read
                rows_publish<false>(a_ready_addr, lane); tl.mark(151);                                          // -> h += A . Wo_sa
                next_chunks(2 * it);

                // ================= cross-attention
                rows_wait(bars, 1, ph[1]); tl.mark(103);
                tmem_ld32(trow + kColH + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmStSa + kStBo + c0);                  // deferred bias of Wo_sa
                tmem_st32(trow + kColH + c0, v);
                store_a16<kBf16>(smem_u32(xbuf), r, c0, v);                  // 16-bit image of h for the fused FFN up-projection
                store_a16<kBf16>(smem_u32(xbuf), r, c0 + 16, v + 16);
                row_stats32<false>(rs, v, mean, rstd);
                normalize32(v, mean, rstd);                                  // LN affine folded into Wq_ca
                store_a16<kBf16>(awork, r, c0, v);
                store_a16<kBf16>(awork, r, c0 + 16, v + 16);
                tmem_wait_st();
                rows_publish<false>(a_ready_addr, lane); tl.mark(152);                                          // -> W = LN(h) . Wq_ca
                next_chunks(2 * it + 1);
                rows_wait(bars, 2, ph[2]); tl.mark(104);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmCaBq + c0);
                softmax16(v);
                softmax16(v + 16);
                store_a16<kBf16>(awork, r, c0, v);
                store_a16<kBf16>(awork, r, c0 + 16, v + 16);
                rows_publish<false>(a_ready_addr, lane); tl.mark(153);                                          // -> W = softmax(q) . blockdiag(A_ca)
                rows_wait(bars, 2, ph[2]); tl.mark(105);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                tc_fence_before();                                           // y_ca is in registers: W may take h16 . W1 now
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->w_free));
                row_stats32<false>(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); tl.mark(106);                                   // S = A_emb . We_ca
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStCa, awork, r, c0, s_free_addr, lane);   // releases S as soon as it has been read
                rows_publish<false>(a_ready_addr, lane); tl.mark(154);                                          // -> h += A . Wo_ca

                // ================= FFN (no pre-norm, reference transformer.py:170-173): W[0:64] = h . W1 was issued together
                //                   with the Wo_ca residual GEMM (see the MMA issuer); the deferred bias of Wo_ca is part
                //                   of the folded FFN-up bias and of the layer's final residual bias
                rows_wait(bars, 2, ph[2]); tl.mark(108);
                if constexpr (kPush) {
                    // the a . (W1 Wo_ca) MMAs have completed: the buffer is idle until the refill that follows the next reduction, so the
                    // peer may deposit its partial of that reduction there
                    const uint32_t nseq = (uint32_t)(si * L + it + 1);
                    if (threadIdx.x == 0 && nseq < (uint32_t)(a.n_steps * L))
                        mbar_arrive_cluster(mapa_u32(smem_u32(&bars->recv_free[nseq & 1u]), (uint32_t)(rank ^ 1)));
                }
                {
                    float u[16];                                             // hidden 64 = 4 quarters of 16
                    tmem_ld16(trow + kColW + 16 * cq, u);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        u[i] += prm[kPrmFfB1 + 16 * cq + i], u[i + 1] += prm[kPrmFfB1 + 16 * cq + i + 1];
                        gelu_erf2(u[i], u[i + 1]);
                    }
                    store_a16<kBf16>(awork, r, 16 * cq, u);
                }
                rows_publish<false>(a_ready_addr, lane); tl.mark(156);                                          // -> W = GELU(.) . W2
                rows_wait(bars, 2, ph[2]); tl.mark(109);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmFfB2 + c0);
                row_stats32<false>(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); tl.mark(110);                                   // S = A_emb . We_ffn
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStFf, awork, r, c0);
                rows_publish<false>(a_ready_addr, lane); tl.mark(157);                                          // -> h += A . Wo_ffn
                rows_wait(bars, 1, ph[1]); tl.mark(111);
            }

            // ---- residual stream after layer `it` (deferred bias of the last FFN block)
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            if (it >= 0) {
                add_bias32(v, prm + kPrmStFf + kStBo + c0);
                if (it + 1 < L) tmem_st32(trow + kColH + c0, v);        // h keeps living in TMEM
            }
            if (it + 1 == L) {
                // ---- step epilogue (reference transformer.py:496, gaussian_diffusion.py:812-830 / 605-665): the output head runs
                //      on the tensor core on the UNROUNDED final h (hi | lo split, 16-bit weights, fp32 accumulate); each of the
                //      four threads of a row then owns 8 of the 32 accumulator columns (26 valid): bias, clamp, sampler update.
                store_a16<kBf16>(awork, r, c0, v);                          // hi image -> awork
                store_a16<kBf16>(awork, r, c0 + 16, v + 16);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {                            // lo = v - float(round16(v)) -> xbuf (contiguous with awork)
                    const float2 back = unpack2<kBf16>(pack2<kBf16>(v[i], v[i + 1]));
                    v[i] -= back.x, v[i + 1] -= back.y;
                }
                store_a16<kBf16>(smem_u32(xbuf), r, c0, v);
                store_a16<kBf16>(smem_u32(xbuf), r, c0 + 16, v + 16);
                rows_publish<false>(a_ready_addr, lane);
                const int smode = a.mode & 0xF;
                const float* nzp = a.noise != nullptr ? a.noise + (size_t)si * a.noise_stride : nullptr;
                const int p0 = (int)cq * 8, np = valid ? min(8, kP - p0) : 0;           // this thread's outputs p0 .. p0 + np - 1
                const size_t e0 = (size_t)(row0g + r) * kP + p0;
                float xo[8], nz[8], bo8[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {                                           // in flight while the head GEMM runs
                    xo[i] = (smode != 0 && i < np) ? __ldcg(x_src + e0 + i) : 0.f;
                    nz[i] = (smode != 0 && nzp != nullptr && i < np) ? __ldcg(nzp + e0 + i) : 0.f;
                    bo8[i] = __ldg(a.bo + p0 + i);
                }
                float cf[8];                                                            // this step's coefficient row, fetched before the wait
                {
                    const float4* cf4 = reinterpret_cast<const float4*>(a.coef + (size_t)(smode != 0 ? tstep : 0) * 8);
                    const float4 ca = smode != 0 ? __ldg(cf4) : make_float4(0.f, 0.f, 0.f, 0.f), cb = smode != 0 ? __ldg(cf4 + 1) : ca;
                    cf[0] = ca.x, cf[1] = ca.y, cf[2] = ca.z, cf[3] = ca.w, cf[4] = cb.x, cf[5] = cb.y, cf[6] = cb.z, cf[7] = cb.w;
                }
                float* x0p = a.x0_out + (size_t)si * a.x0_stride;
                rows_wait(bars, 2, ph[2]);
                float o8[8];
                tmem_ld8(trow + kColW + p0, o8);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < np) {
                        float x0 = o8[i] + bo8[i];
                        if (a.mode & 0x10) x0 = fminf(fmaxf(x0, -1.f), 1.f);
                        x0p[e0 + i] = x0;
                        if (smode != 0) {
                            const float xn = smode == 1 ? ddim_rule(xo[i], x0, cf, nz[i]) : ddpm_rule(xo[i], x0, cf, nz[i]);
                            a.x_out[e0 + i] = xn;
                            if (a.x_trace != nullptr) a.x_trace[(size_t)si * a.M * kP + e0 + i] = xn;
                        }
                    }
                }
                // no CTA barrier: the next step's prologue re-reads exactly the x elements this thread has just written, and the
                // operand buffers it overwrites were last read by the head GEMM every thread has just waited for
                tc_fence_before();
                break;
            }

            // ================= self-attention head of layer it+1: LN -> q | k | v, then the time-axis reduction
            row_stats32<false>(rs, v, mean, rstd);
            normalize32(v, mean, rstd);                                  // LN affine folded into Wq/Wk/Wv
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish<false>(a_ready_addr, lane); tl.mark(158);
            if (it < 0) {
                // ---- rest of the step prologue, overlapped with the first q|k|v MMAs: this tile's A_emb image -> global
                if (!kGx || si == 0) {
                    // the whole image here, four chunks per thread with their loads in flight together
                    uint8_t* img = a.aemb_out + (kGx ? (size_t)(si & 1) * a.aemb_stride : (size_t)0) + (size_t)blockIdx.x * 8 * kStageABytes;
                    const int tx = threadIdx.x;
                    const float* ste = reinterpret_cast<const float*>(xbuf) + kP * kD + kD + kTileRows * kP;
                    for (int k0 = 0; k0 < 16; k0 += 4) {                 // 128 rows x 64 chunks of 8 features; a warp = 1 KB of one row
                        float4 xa[4][2];
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int task = (k0 + k) * kRowThreads + tx;
                            const long gg = row0g + min(task >> 6, max(nrows - 1, 0));
                            const float4* xr4 = reinterpret_cast<const float4*>(a.xp + gg * kE + (task & 63) * 8);
                            xa[k][0] = __ldg(xr4), xa[k][1] = __ldg(xr4 + 1);
                        }
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int task = (k0 + k) * kRowThreads + tx;
                            const int row = task >> 6, ch = task & 63;
                            const float4* tr = reinterpret_cast<const float4*>(ste + ch * 8);
                            const float4 a0 = xa[k][0], a1 = xa[k][1], t0v = tr[0], t1v = tr[1];
                            const float e8[8] = {a0.x + t0v.x, a0.y + t0v.y, a0.z + t0v.z, a0.w + t0v.w, a1.x + t1v.x, a1.y + t1v.y, a1.z + t1v.z, a1.w + t1v.w};
                            uint32_t p[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) p[i] = pack2<kBf16>(silu_f<kBf16>(e8[2 * i]), silu_f<kBf16>(e8[2 * i + 1]));
                            const uint4 pk = row < nrows ? make_uint4(p[0], p[1], p[2], p[3]) : make_uint4(0, 0, 0, 0);
                            *reinterpret_cast<uint4*>(img + (size_t)(ch >> 3) * kABlockBytes + sw128_offset(row, ch & 7)) = pk;
                        }
                    }
                }
                __threadfence();                                       // the image must have reached L2 ...
                asm volatile("fence.proxy.async;" ::: "memory");      // ... and be ordered before the bulk-copy (async proxy) reads
                named_bar_sync(5, kRowThreads);                        // every row thread is done with the staging area in xbuf
                if (lane == 0) mbar_arrive(smem_u32(&bars->aemb_ready));
                tl.mark(127);
            }
            rows_wait(bars, 2, ph[2]);
            tl.mark(112);
            {
                // every row warp has published its LayerNorm output (the QKV MMAs needed all 16 arrivals), so nobody reads the
                // parameters of layer `it` any more: fetch the next layer's block asynchronously, it lands long before it is used
                const float4* pn4 = reinterpret_cast<const float4*>(a.prm + (size_t)(it + 1) * kPrmFloats);
                static_assert(kPrmFloats % 4 == 0 && kPrmFloats / 4 <= 2 * kRowThreads, "parameter block layout");
                cp_async16(reinterpret_cast<float4*>(prm) + threadIdx.x, pn4 + threadIdx.x);
                if (threadIdx.x + kRowThreads < kPrmFloats / 4) cp_async16(reinterpret_cast<float4*>(prm) + threadIdx.x + kRowThreads, pn4 + threadIdx.x + kRowThreads);
            }
            {
                // Time-axis softmax + K^T V (reference :111,:117) on the tensor cores, one clip per cluster.  Two 32 KB
                // operand-image buffers X (xbuf), Y (ring B) are all the scratch it needs:
                //   k (16-bit) -> X ; column maxima by a column scan of X ; E = exp(k - max) -> X ; V -> Y ;
                //   P = E^T V as 8 MN-major MMAs over the tile's tokens ; column sums by a column scan of the E image
                //   (same rounded values as the MMA sees).
                // The partial (max, sum, diagonal 16x16 blocks of P) stays in this CTA's shared memory (over Y); every CTA
                // of the cluster pulls all nt partials and merges them into its own compact head-block B-operand image (X).
                float* pm = reinterpret_cast<float*>(xchg);            // [8 rg][128] exchange (max, then sums); xchg is idle here
                float* msm = red;                                      // [128] maxima
                float* ssm = msm + 128;                                // [128] sums
                uint8_t* Xp = xbuf;
                const uint32_t eimg = smem_u32(xbuf), vimg = smem_u32(ringB);
                const int tx = threadIdx.x;
                const uint32_t seq = (uint32_t)(si * L + it + 1);            // reductions completed so far in this launch
                float* mypart = reinterpret_cast<float*>(ringB);             // [kKvPartFloats] (distributed-shared-memory exchange)
                const int col = tx & 127;
                // column pair (2 cp, 2 cp + 1), rows [16 rg, 16 rg + 16) of a [128 x 128] 16-bit operand image
                const int cp = tx & 63, rg = tx >> 6;
                const uint8_t* pair_base = Xp + (cp >> 5) * kABlockBytes + (cp & 3) * 4;
                const uint32_t pair_chunk = (uint32_t)(cp & 31) >> 2;
                auto pair_at = [&](int row) -> const uint32_t* {
                    return reinterpret_cast<const uint32_t*>(pair_base + row * 128 + ((pair_chunk ^ ((uint32_t)row & 7u)) << 4));
                };
                constexpr uint32_t kNegInf2 = kBf16 ? 0xFF80FF80u : 0xFC00FC00u;
                auto max2 = [](uint32_t x, uint32_t y) -> uint32_t {
                    if constexpr (kBf16) {
                        __nv_bfloat162 r2 = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&x), *reinterpret_cast<__nv_bfloat162*>(&y));
                        return *reinterpret_cast<uint32_t*>(&r2);
                    } else {
                        __half2 r2 = __hmax2(*reinterpret_cast<__half2*>(&x), *reinterpret_cast<__half2*>(&y));
                        return *reinterpret_cast<uint32_t*>(&r2);
                    }
                };
                float kx[32], vx[32];
                tmem_ld32(trow + kColS + 128 + c0, kx);                   // k and v in flight together: one TMEM round trip
                tmem_ld32(trow + kColW + c0, vx);
                tmem_wait_ld();

                add_bias32(kx, prm_sa + kPrmSaBk + c0);
                if (!keep) {                 // masked frame: k - 1e6 (reference :110); padding row of the tile: -inf (E = 0 exactly)
                    const float madd = valid ? -1000000.f : -INFINITY;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        kx[i] += madd;
                        if constexpr (!kBf16) kx[i] = fmaxf(kx[i], -60000.f);  // fp16 image of k: keep the mask finite
                    }
                }
                // Softmax over time is shift-invariant per key column: when the host could bound |k| (Cauchy-Schwarz on the
                // LayerNorm output, ||n|| <= sqrt(128)) by a small constant, that bound replaces the running column max --
                // no 16-bit k image, no column scan, three CTA barriers less.  Otherwise: exact tile-local column max.
                const bool static_shift = (a.static_mask >> (it + 1)) & 1u;
                if (!static_shift) {
                    store_a16<kBf16>(eimg, r, c0, kx);
                    store_a16<kBf16>(eimg, r, c0 + 16, kx + 16);
                    add_bias32(vx, prm_sa + kPrmSaBv + c0);
                    named_bar_sync(5, kRowThreads);
                    {   // column maxima: this thread scans 16 rows of a column PAIR (packed 16-bit max)
                        uint32_t m0 = kNegInf2;
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) m0 = max2(m0, *pair_at(16 * rg + rr));
                        *reinterpret_cast<float2*>(pm + rg * 128 + 2 * cp) = unpack2<kBf16>(m0);
                    }
                    named_bar_sync(5, kRowThreads);
                    if (tx < 128) {
                        float mm = pm[col];
#pragma unroll
                        for (int q8 = 1; q8 < 8; ++q8) mm = fmaxf(mm, pm[q8 * 128 + col]);
                        msm[tx] = mm;
                    }
                    named_bar_sync(5, kRowThreads);
                } else {
                    add_bias32(vx, prm_sa + kPrmSaBv + c0);
                    if (tx < 128) msm[tx] = 0.f;                       // every tile reports the same (virtual) max: the merge just adds
                }
                {   // E = exp(k - shift) (0 for padding rows) -> X ; V -> Y
                    const float* mrow = static_shift ? prm_sa + 384 + c0 : msm + c0;
#pragma unroll
                    for (int i = 0; i < 32; ++i) kx[i] = ex2_ftz((kx[i] - mrow[i]) * 1.4426950408889634f);
                    if constexpr (!kBf16) {
                        if (!valid) {
#pragma unroll
                            for (int i = 0; i < 32; ++i) kx[i] = 0.f;      // fp16: the padding rows were clamped to a finite value above
                        }
                    }
                    store_a16<kBf16>(eimg, r, c0, kx);
                    store_a16<kBf16>(eimg, r, c0 + 16, kx + 16);
                }
                if (!keep) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) vx[i] = 0.f;
                }
                store_a16<kBf16>(vimg, r, c0, vx);
                store_a16<kBf16>(vimg, r, c0 + 16, vx + 16);
                tl.mark(120);
                rows_publish<false>(a_ready_addr, lane);                      // -> W = E^T . V  (8 MMAs over the tokens)
                tl.mark(122);
                {
                    // q: softmax over head-dim -> operand buffer (A operand of the next layer's q . blockdiag(A_sa));
                    // runs while the tensor core does E^T V.  Once q and k have left S the next FiLM projection may start.
                    float qv[32];
                    tmem_ld32(trow + kColS + c0, qv);
                    tmem_wait_ld();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(s_free_addr);
                    add_bias32(qv, prm_sa + kPrmSaBq + c0);
                    softmax16(qv);
                    softmax16(qv + 16);
                    store_a16<kBf16>(awork, r, c0, qv);
                    store_a16<kBf16>(awork, r, c0 + 16, qv + 16);
                    named_bar_sync(5, kRowThreads);                    // E image complete (all rows published); prm_sa is dead
                    if (it + 2 < L && tx < 128)                         // SA biases + static key shift of layer it + 2
                        cp_async16(reinterpret_cast<float4*>(prm_sa) + tx,
                                   tx < 96 ? reinterpret_cast<const float4*>(a.prm + (size_t)(it + 2) * kPrmFloats) + tx
                                           : reinterpret_cast<const float4*>(a.kshift + (size_t)(it + 2) * kD) + (tx - 96));
                    float2 s0 = make_float2(0.f, 0.f);                  // column sums from the rounded E
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        const float2 x = unpack2<kBf16>(*pair_at(16 * rg + rr));
                        s0.x += x.x, s0.y += x.y;
                    }
                    *reinterpret_cast<float2*>(pm + rg * 128 + 2 * cp) = s0;
                    named_bar_sync(5, kRowThreads);
                    if (tx < 128) {
                        float ss = pm[col];
#pragma unroll
                        for (int q8 = 1; q8 < 8; ++q8) ss += pm[q8 * 128 + col];
                        ssm[tx] = ss;
                    }
                }
                rows_wait(bars, 2, ph[2]);                                // E^T V complete: the E and V images are dead
                tl.mark(123);
                const uint32_t gtag = a.gx_tag0 + seq + 1u;
                if constexpr (kGx) {
                    // TMEM lane = key feature r; the diagonal block of its head is 16 value columns: warp cq takes 4 of them, so all 16
                    // row warps share the write.  Words are laid out [column pair][feature]: a warp's store is 512 contiguous bytes.
                    float p0[4], p1[4];
                    tmem_ld4(trow + kColW + 32 * lq + 4 * cq, p0);
                    tmem_ld4(trow + kColW + 32 * lq + 16 + 4 * cq, p1);
                    tmem_wait_ld();
                    const bool hi = (lane & 16) != 0;
                    uint2* gp = a.gx_part + (((size_t)clip * 2 + (seq & 1u)) * nt + rank) * kKvPartFloats;
#pragma unroll
                    for (int i = 0; i < 2; ++i)
                        st_global_v4(reinterpret_cast<uint4*>(gp + 256 + (2 * cq + i) * 256 + r * 2),
                                     make_uint4(__float_as_uint(hi ? p1[2 * i] : p0[2 * i]), gtag, __float_as_uint(hi ? p1[2 * i + 1] : p0[2 * i + 1]), gtag));
                    if (cq == 0) {
                        st_global_v2(gp + 128 + tx, make_uint2(__float_as_uint(ssm[tx]), gtag));
                        if (!static_shift) st_global_v2(gp + tx, make_uint2(__float_as_uint(msm[tx]), gtag));
                    }
                } else if (kPush && cq == 0) {
                    float pr[32];
                    tmem_ld32(trow + kColW + 32 * lq, pr);
                    tmem_wait_ld();
                    const int o = (lane & 16);
                    const uint32_t off = (uint32_t)(256 + (r >> 4) * 256 + (r & 15) * 16);
                    float4* dst = reinterpret_cast<float4*>(mypart + off);
                    if (seq > 0) mbar_wait(smem_u32(&bars->recv_free[seq & 1u]), ((seq - 1u) >> 1) & 1u);
                    const uint32_t rb = mapa_u32(smem_u32(w1c), (uint32_t)(rank ^ 1));
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 pv = o ? make_float4(pr[16 + 4 * i], pr[17 + 4 * i], pr[18 + 4 * i], pr[19 + 4 * i])
                                            : make_float4(pr[4 * i], pr[4 * i + 1], pr[4 * i + 2], pr[4 * i + 3]);
                        dst[i] = pv;
                        st_dsmem_f32x4(rb + (off + 4u * i) * 4u, pv);
                    }
                    mypart[tx] = msm[tx], mypart[128 + tx] = ssm[tx];
                    st_dsmem_f32(rb + (uint32_t)tx * 4u, msm[tx]), st_dsmem_f32(rb + (uint32_t)(128 + tx) * 4u, ssm[tx]);
                } else if (cq == 0) {      // TMEM lane = key feature r; its head's 16 value columns are the diagonal block
                    float pr[32];
                    tmem_ld32(trow + kColW + 32 * lq, pr);
                    tmem_wait_ld();
                    const int o = (lane & 16);
                    float4* dst = reinterpret_cast<float4*>(mypart + 256 + (r >> 4) * 256 + (r & 15) * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        dst[i] = o ? make_float4(pr[16 + 4 * i], pr[17 + 4 * i], pr[18 + 4 * i], pr[19 + 4 * i])
                                   : make_float4(pr[4 * i], pr[4 * i + 1], pr[4 * i + 2], pr[4 * i + 3]);
                    mypart[tx] = msm[tx], mypart[128 + tx] = ssm[tx];     // cq == 0 <=> tx < 128: this thread's own column max / sum
                }
                if (!kGx) named_bar_sync(5, kRowThreads);                  // partial complete (global exchange: every word carries its own flag)
                tl.mark(124);
                // ---- publish to the peers: release at cluster scope, one remote arrive per peer
                // (no separate fence: the arrive is a release at cluster scope, and release is cumulative over the writes of the
                //  other row threads that were ordered before it by the CTA barrier above)
                if (!kGx && nt > 1 && tx < nt && tx != rank) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->part_ready[seq & 1u]), (uint32_t)tx));
                // ---- the parameter blocks fetched with cp.async above must have landed before the barrier below (the merge writes
                //      the compact head-block image over the dead E image: nothing to clear)
                cp_async_wait_all();
                if (!kGx && nt > 1 && tx == 0) mbar_wait_acq_cluster(smem_u32(&bars->part_ready[seq & 1u]), (seq >> 1) & 1u);
                named_bar_sync(5, kRowThreads);                            // peers' partials visible
                tl.mark(125);
                // ---- merge the nt partials (online-softmax rescaling, four tiles per round trip) into the compact head-block
                //      B-operand image.
                if constexpr (kGx) {
                    gx_merge<kBf16>(a.gx_part + ((size_t)clip * 2 + (seq & 1u)) * nt * kKvPartFloats,
                                    a.gx_slice + ((size_t)clip * 2 + (seq & 1u)) * (kD * 8), reinterpret_cast<float4*>(ringB), xbuf, nt, rank, tx,
                                    gtag, static_shift);
                } else if (kPush) {
                    // two tiles, both partials in THIS CTA's shared memory (own: ring B, peer's: W1 Wo_ca buffer): plain loads.
                    // Thread -> head hh, key features d0, d0 + 1, value columns l0, l0 + 1.
                    const int hh = tx >> 6, sub = tx & 63, d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
                    const float* pa = mypart;
                    const float* pb = reinterpret_cast<const float*>(w1c);
                    const int o_m = 16 * hh + d0, o_r = 256 + hh * 256 + d0 * 16 + l0;
                    const float2 ma = *reinterpret_cast<const float2*>(pa + o_m), mb = *reinterpret_cast<const float2*>(pb + o_m);
                    const float2 sa = *reinterpret_cast<const float2*>(pa + 128 + o_m), sb = *reinterpret_cast<const float2*>(pb + 128 + o_m);
                    const float2 ra0 = *reinterpret_cast<const float2*>(pa + o_r), rb0 = *reinterpret_cast<const float2*>(pb + o_r);
                    const float2 ra1 = *reinterpret_cast<const float2*>(pa + o_r + 16), rb1 = *reinterpret_cast<const float2*>(pb + o_r + 16);
                    float wa0 = 1.f, wb0 = 1.f, wa1 = 1.f, wb1 = 1.f;
                    if (!static_shift) {                                  // running column max: rescale to the common maximum
                        const float M0 = fmaxf(ma.x, mb.x), M1 = fmaxf(ma.y, mb.y);
                        wa0 = __expf(ma.x - M0), wb0 = __expf(mb.x - M0), wa1 = __expf(ma.y - M1), wb1 = __expf(mb.y - M1);
                    }
                    const float s0 = fmaf(sa.x, wa0, sb.x * wb0), s1 = fmaf(sa.y, wa1, sb.y * wb1);
                    const float i0 = s0 > 0.f ? 1.f / s0 : 0.f, i1 = s1 > 0.f ? 1.f / s1 : 0.f;      // s = 0: every frame masked (static shift)
                    const float o[2][2] = {{fmaf(ra0.x, wa0, rb0.x * wb0) * i0, fmaf(ra0.y, wa0, rb0.y * wb0) * i0},
                                           {fmaf(ra1.x, wa1, rb1.x * wb1) * i1, fmaf(ra1.y, wa1, rb1.y * wb1) * i1}};
#pragma unroll
                    for (int dd = 0; dd < 2; ++dd) {
#pragma unroll
                        for (int ll = 0; ll < 2; ++ll)
                            *reinterpret_cast<uint16_t*>(xbuf + bdc_offset((uint32_t)hh, (uint32_t)(d0 + dd), (uint32_t)(l0 + ll))) = pack1<kBf16>(o[dd][ll]);
                    }
                } else if (nt <= kDirectMergeTiles) {
                    // small clusters: every CTA pulls every partial.  Thread -> head hh, key features d0, d0 + 1, value columns l0, l0 + 1.
                    const int hh = tx >> 6, sub = tx & 63, d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
                    const uint32_t pbase = smem_u32(mypart);
                    const uint32_t o_m = (uint32_t)(16 * hh + d0) * 4, o_s = o_m + 512;
                    const uint32_t o_r0 = (uint32_t)(256 + hh * 256 + d0 * 16 + l0) * 4, o_r1 = o_r0 + 64;
                    float M0 = -INFINITY, M1 = -INFINITY, a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
                    for (int j0 = 0; j0 < nt; j0 += 4) {
                        float2 mi[4], sj[4], r0[4], r1[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (j0 + j < nt) {                             // block-uniform: no loads for absent tiles
                                const uint32_t pa = mapa_u32(pbase, (uint32_t)(j0 + j));
                                mi[j] = ld_dsmem_f32x2(pa + o_m);
                                sj[j] = ld_dsmem_f32x2(pa + o_s);
                                r0[j] = ld_dsmem_f32x2(pa + o_r0);
                                r1[j] = ld_dsmem_f32x2(pa + o_r1);
                            } else {
                                mi[j] = make_float2(-INFINITY, -INFINITY);
                                sj[j] = r0[j] = r1[j] = make_float2(0.f, 0.f);
                            }
                        }
                        float n0 = M0, n1 = M1;
#pragma unroll
                        for (int j = 0; j < 4; ++j) n0 = fmaxf(n0, mi[j].x), n1 = fmaxf(n1, mi[j].y);
                        const float c0s = __expf(M0 - n0), c1s = __expf(M1 - n1);       // exp(-inf) = 0 on the first round
                        s0 *= c0s, a00 *= c0s, a01 *= c0s, s1 *= c1s, a10 *= c1s, a11 *= c1s;
                        M0 = n0, M1 = n1;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float w0 = __expf(mi[j].x - M0), w1 = __expf(mi[j].y - M1);   // 0 for absent tiles
                            s0 = fmaf(sj[j].x, w0, s0), s1 = fmaf(sj[j].y, w1, s1);
                            a00 = fmaf(r0[j].x, w0, a00), a01 = fmaf(r0[j].y, w0, a01);
                            a10 = fmaf(r1[j].x, w1, a10), a11 = fmaf(r1[j].y, w1, a11);
                        }
                    }
                    const float i0 = s0 > 0.f ? 1.f / s0 : 0.f, i1 = s1 > 0.f ? 1.f / s1 : 0.f;      // s = 0: every frame masked (static shift)
                    const float o[2][2] = {{a00 * i0, a01 * i0}, {a10 * i1, a11 * i1}};
#pragma unroll
                    for (int dd = 0; dd < 2; ++dd) {
#pragma unroll
                        for (int ll = 0; ll < 2; ++ll)
                            *reinterpret_cast<uint16_t*>(xbuf + bdc_offset((uint32_t)hh, (uint32_t)(d0 + dd), (uint32_t)(l0 + ll))) = pack1<kBf16>(o[dd][ll]);
                    }
                } else {
                    // large clusters: reduce-scatter + all-gather (the shared-memory port of an SM serves ~20 B/clk to its
                    // peers, so all-to-all pulls of 9 KB partials do not scale).  CTA `rank` merges key features
                    // [rank fs, rank fs + fs) from all partials into a 16-bit slice; then everyone gathers the nt slices.
                    const int fs = (kD + nt - 1) / nt;
                    uint16_t* myslice = reinterpret_cast<uint16_t*>(ringB + 12288);      // [fs][16], behind the partial
                    {
                        const int dl = tx >> 4, l = tx & 15, d = rank * fs + dl;
                        if (dl < fs && d < kD) {
                            const uint32_t pbase = smem_u32(mypart);
                            const uint32_t o_m = (uint32_t)d * 4, o_s = o_m + 512, o_p = (uint32_t)(256 + (d >> 4) * 256 + (d & 15) * 16 + l) * 4;
                            float Mx = -INFINITY, ss = 0.f, acc = 0.f;
                            for (int j0 = 0; j0 < nt; j0 += 4) {
                                float mi[4], sj[4], pj[4];
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const bool on = j0 + j < nt;
                                    const uint32_t pa = mapa_u32(pbase, (uint32_t)(on ? j0 + j : rank));
                                    mi[j] = ld_dsmem_f32(pa + o_m), sj[j] = ld_dsmem_f32(pa + o_s), pj[j] = ld_dsmem_f32(pa + o_p);
                                    if (!on) mi[j] = -INFINITY;
                                }
                                float nm = Mx;
#pragma unroll
                                for (int j = 0; j < 4; ++j) nm = fmaxf(nm, mi[j]);
                                const float cs = __expf(Mx - nm);
                                ss *= cs, acc *= cs, Mx = nm;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float w = __expf(mi[j] - Mx);
                                    ss = fmaf(sj[j], w, ss), acc = fmaf(pj[j], w, acc);
                                }
                            }
                            myslice[dl * 16 + l] = pack1<kBf16>(ss > 0.f ? acc / ss : 0.f);
                        }
                    }
                    named_bar_sync(5, kRowThreads);
                    if (tx < nt && tx != rank) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->slice_ready[seq & 1u]), (uint32_t)tx));
                    if (tx == 0) mbar_wait_acq_cluster(smem_u32(&bars->slice_ready[seq & 1u]), (seq >> 1) & 1u);
                    named_bar_sync(5, kRowThreads);
                    {
                        const int d = tx >> 2, l4 = (tx & 3) * 4, j = d / fs, dl = d - j * fs;
                        const uint32_t src = mapa_u32(smem_u32(myslice), (uint32_t)j) + (uint32_t)(dl * 16 + l4) * 2;
                        uint32_t w0, w1;
                        asm volatile("ld.shared::cluster.v2.b32 {%0, %1}, [%2];" : "=r"(w0), "=r"(w1) : "r"(src) : "memory");
                        const uint16_t vals[4] = {(uint16_t)(w0 & 0xFFFFu), (uint16_t)(w0 >> 16), (uint16_t)(w1 & 0xFFFFu), (uint16_t)(w1 >> 16)};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            *reinterpret_cast<uint16_t*>(xbuf + bdc_offset((uint32_t)(d >> 4), (uint32_t)(d & 15), (uint32_t)(l4 + i))) = vals[i];
                    }
                }
                rows_publish<false>(a_ready_addr, lane);                          // -> y = q . blockdiag(A_sa) of layer it+1
                tl.mark(126);
                if (!kGx && !kPush && nt > 1) {                            // every pull of this CTA has completed: the peers may reuse ring B
                    named_bar_sync(5, kRowThreads);                        // (off the critical path: the q . A GEMM is already on its way)
                    if (tx < nt && tx != rank) mbar_arrive_cluster(mapa_u32(smem_u32(&bars->pull_done[seq & 1u]), (uint32_t)tx));
                }
            }
        }
        }   // launch step si
        tl.mark(2);
        tl.finish();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                                   // no CTA leaves while a peer may still touch its shared memory
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

constexpr int kClipSmemBytes = kPRingAStages * kStageBytes + kRingBStages * kRingBStageBytes + 2 * kAworkBytes +
                               (kPrmFloats + 512) * 4 + 512 * 8 + kClipRedFloats * 4 + 16384 + sizeof(ClipBarriers) + 1024;
static_assert(kClipSmemBytes <= 232448, "clip kernel exceeds the 227 KB shared-memory limit");

}  // namespace dc
