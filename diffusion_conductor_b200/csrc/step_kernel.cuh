// Persistent whole-step kernel: one CTA per 128-token tile runs ALL decoder layers of a denoise step without
// leaving the SM.  The residual stream stays in TMEM for the whole step, q stays in shared memory, the FiLM
// projection of the next layer streams in while the current layer's tail is still being computed, and the one
// cross-tile dependency of the model -- the time-axis softmax of the self-attention keys, a per-clip reduction --
// is resolved in-kernel: every tile publishes a (max, sum, K^T V) partial per clip segment, the CTA that completes
// a clip merges them into the clip's block-diagonal B-operand image and releases a per-clip counter that the
// consumers' producer lane acquires before loading the image.  Requires every CTA to be co-resident (tiles <= SMs)
// and T >= 128 (a tile touches at most two clips); otherwise the host falls back to one launch per layer.
#pragma once
#include "tile_kernels.cuh"

namespace dc {

constexpr int kPRingAStages = 2;        // persistent kernel: 2 x 48 KB (the FiLM projection is shared-memory-bandwidth bound)
constexpr int kRedFloats = 2048 + 256 + 256 + 8;

struct StepArgs {
    int L, M, T;
    const uint8_t* wbuf;        // packed weights [L][1 MiB]
    const uint8_t* aemb;        // A_emb image [tiles][8][16 KB]
    const float* prm;           // [L][kPrmFloats]
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
    const float* WjT;           // [26][128] joint_embed weight, transposed
    const float* bj;            // [128]
    const float* pos;           // [num_frames][128] sequence_embedding
    const float* WoT;           // [128][32] output head, transposed + padded
    const float* bo;            // [32]
    uint8_t* aemb_out;          // == aemb: this CTA writes its own tile's A_emb image first
    const uint8_t* bd_ca;       // cross-attention images: clip stride bd_ca_stride, layer stride kAworkBytes
    size_t bd_ca_stride;
    uint8_t* bd_sa_out;         // [B][32 KB] self-attention images (written by the merging CTA, read by the clip's tiles)
    float* kv_part;             // [2 parity][tiles][2][kKvPartFloats]
    int* clip_cnt;              // [B] arrival counters: grow monotonically over the launch (zeroed by the host before it)
    int* clip_done;             // [B] number of completed merges in this step (zeroed by step_begin)
    const long long* length;    // [B] or null
    uint32_t off[12];           // byte offsets of the packed matrices inside a layer slab (see dc_api.cu)
    unsigned long long* timeline;
    int dbg;                    // bit 0: skip the FiLM projection MMAs (timing experiments only; results are wrong)
};
enum { kOWeSa = 0, kOWoSa, kOWeCa, kOWqCa, kOWoCa, kOWeFf, kOW1, kOW2, kOWoFf, kOWq, kOWk, kOWv };

__device__ __forceinline__ void tl_mark(const StepArgs& a, unsigned long long id) {
    if (a.timeline != nullptr && blockIdx.x == 0) {
        const unsigned long long slot = atomicAdd(a.timeline, 1ull);
        if (slot < 2040) {
            a.timeline[1 + 2 * slot] = (unsigned long long)clock64();
            a.timeline[2 + 2 * slot] = id;
        }
    }
}

template <bool kBf16>
__global__ void __launch_bounds__(kTileThreads, 1) step_kernel(const __grid_constant__ StepArgs a) {
    constexpr bool kPair = false;
    constexpr int kNA = kPRingAStages, kSA = kStageBytes, kNB = kRingBStages, kSB = kRingBStageBytes;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ringA = smem;
    uint8_t* ringB = ringA + kNA * kSA;                               // also the V image of the fused reduction
    uint8_t* awork_p = ringB + kNB * kSB;
    uint8_t* xbuf = awork_p + kAworkBytes;                            // k / E image of the fused reduction
    float* prm = reinterpret_cast<float*>(xbuf + kAworkBytes);        // [kPrmFloats] layer `it`
    float* prm_sa = prm + kPrmFloats;                                 // [384] SA biases of layer it + 1
    float2* xchg = reinterpret_cast<float2*>(prm_sa + 384);           // [4][128]
    float* red = reinterpret_cast<float*>(xchg + 512);                // kRedFloats
    LayerBarriers* bars = reinterpret_cast<LayerBarriers*>(red + kRedFloats);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int L = a.L;

    pdl_trigger();
    if (warp == kProducerWarp && lane == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars->fullA[i]), 1), mbar_init(smem_u32(&bars->emptyA[i]), 1);
        for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars->fullB[i]), 1), mbar_init(smem_u32(&bars->emptyB[i]), 1);
        mbar_init(smem_u32(&bars->a_ready), kRowWarps);
        mbar_init(smem_u32(&bars->s_free), kRowWarps);
        mbar_init(smem_u32(&bars->q_full), 1);
        mbar_init(smem_u32(&bars->aemb_ready), kRowWarps);
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
    const uint32_t tmem_base = bars->tmem_base;
    pdl_wait();
    if (threadIdx.x == 0) tl_mark(a, 1);

    RowSegs segs;
    segs.init((int)blockIdx.x, kTileRows, a.M, a.T);

    if (warp == kProducerWarp) {
        // ---------------- ring A: per layer 3 FiLM projections (8 stages each), then Wk, Wv of the next layer
        if (lane == 0) {
            const uint8_t* a_img = a.aemb + (size_t)blockIdx.x * 8 * kStageABytes;
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
                if (it >= 0) {
                    if (it == 0) {
                        mbar_wait(smem_u32(&bars->aemb_ready), (uint32_t)si & 1u);   // the row threads have written this step's A_emb image
                        asm volatile("fence.proxy.async;" ::: "memory");
                    }
                    const uint8_t* slab = a.wbuf + ((size_t)it << 20);
                    const uint32_t so[3] = {a.off[kOWeSa], a.off[kOWeCa], a.off[kOWeFf]};
                    for (int o = 0; o < 3; ++o)
                        for (int s = 0; s < kSopStages; ++s) stage_in(a_img + (size_t)s * kStageABytes, slab + so[o] + (size_t)s * kStageWBytes, kStageWBytes);
                }
                if (it + 1 < L) {
                    const uint8_t* slab = a.wbuf + ((size_t)(it + 1) << 20);
                    stage_in(nullptr, slab + a.off[kOWk], 32768);
                    stage_in(nullptr, slab + a.off[kOWv], 32768);
                }
            }
        }
    } else if (warp == kProducerBWarp) {
        // ---------------- ring B (one k-block per stage): dependent-GEMM weights and per-clip attention images
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
                if (it >= 0) {
                    const uint8_t* slab = a.wbuf + ((size_t)it << 20);
                    // ring B doubles as the segment-1 attention image until q . blockdiag(A_sa) of this layer has completed
                    mbar_wait(smem_u32(&bars->q_full), qf++ & 1u);
                    load(slab + a.off[kOWoSa], 2, 16384);
                    load(slab + a.off[kOWqCa], 2, 16384);
                    for (int s = 0; s < segs.n_seg; ++s)
                        load(a.bd_ca + (size_t)(segs.first_clip + s) * a.bd_ca_stride + (size_t)it * kAworkBytes, 2, 16384);
                    load(slab + a.off[kOWoCa], 2, 16384);
                    load(slab + a.off[kOW1], 2, 8192);
                    load(slab + a.off[kOW2], 1, 16384);
                    load(slab + a.off[kOWoFf], 2, 16384);
                }
                if (it + 1 < L) load(a.wbuf + ((size_t)(it + 1) << 20) + a.off[kOWq], 2, 16384);
            }
        }
    } else if (warp == kRelayWarp) {
        // ---------------- issuer of the FiLM projections S = A_emb . We (ring A), gated only by s_free
        if (lane == 0) {
            const uint32_t idesc_s = make_idesc<kBf16>(kTileRows, 256);
            uint32_t sj = 0;
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = 0; it < L; ++it) {
                // ring-A items before this layer: 26 L per earlier step, 2 (Wk,Wv of layer 0), 26 per layer
                const uint32_t base_it = (uint32_t)si * 26u * (uint32_t)L + (uint32_t)it * 26 + 2;
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
                    tl_mark(a, 310 + o);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ---------------- issuer of the dependent GEMMs, strictly in chain order with blocking waits
        if (lane == 0) {
            const uint32_t awork = smem_u32(awork_p);
            uint32_t itB = 0, a_phase = 0;
            auto wait_a = [&]() {
                mbar_wait(smem_u32(&bars->a_ready), a_phase & 1u);
                ++a_phase;
                tc_fence_after();
            };
            // B operand through ring B, one k-block per stage; optional lane mask (clip segment)
            auto gemm_b = [&](int kb, int n, uint32_t d_col, bool acc, const uint32_t* mask) {
                const uint32_t idesc = make_idesc<kBf16>(kTileRows, n);
                for (int k = 0; k < kb; ++k, ++itB) {
                    const uint32_t st = itB % kNB, ph = (itB / kNB) & 1u;
                    mbar_wait(smem_u32(&bars->fullB[st]), ph);
                    tc_fence_after();
                    const uint32_t b_base = smem_u32(ringB + st * kSB);
                    if (mask) umma_kblock_masked(tmem_base + d_col, awork + k * kABlockBytes, b_base, idesc, k > 0, mask);
                    else umma_kblock(tmem_base + d_col, awork + k * kABlockBytes, b_base, idesc, acc || k > 0);
                    umma_commit(smem_u32(&bars->emptyB[st]));
                }
            };
            auto seg_gemm = [&]() {
                for (int s = 0; s < segs.n_seg; ++s) {
                    uint32_t m[8];
                    segs.mask(s, m, false);
                    gemm_b(2, 128, kColW, false, m);
                }
            };
            auto done = [&](int which) { umma_commit(smem_u32(&bars->d_ready[which])); };
            const int nvalid = max(0, min(kTileRows, a.M - (int)blockIdx.x * kTileRows));
            const int e_rows = min(nvalid, ((int)blockIdx.x * kTileRows / a.T + 1) * a.T - (int)blockIdx.x * kTileRows);
            const int passes = nvalid > e_rows ? 2 : 1;
            const uint32_t idmn = make_idesc_mn<kBf16>(kTileRows, kTileRows);
            const uint32_t idesc128 = make_idesc<kBf16>(kTileRows, 128);
            for (int si = 0; si < a.n_steps; ++si)
            for (int it = -1; it < L; ++it) {
                if (it >= 0) {
                    wait_a();                                                      // merged attention images written by the row threads
                    for (int sgi = 0; sgi < segs.n_seg; ++sgi) {                   // y = q . blockdiag(A_sa): one lane-masked GEMM per clip segment
                        uint32_t m[8];
                        segs.mask(sgi, m, false);
                        const uint32_t bimg = sgi == 0 ? smem_u32(xbuf) : smem_u32(ringB);
                        for (int k = 0; k < 2; ++k)
                            umma_kblock_masked(tmem_base + kColW, awork + k * kABlockBytes, bimg + k * kABlockBytes, idesc128, k > 0, m);
                    }
                    done(2), tl_mark(a, 200);
                    umma_commit(smem_u32(&bars->q_full));                          // ring B (segment-1 image) may be refilled now
                    wait_a(), gemm_b(2, 128, kColH, true, nullptr), done(1), tl_mark(a, 201);   // h += . Wo_sa
                    wait_a(), gemm_b(2, 128, kColW, false, nullptr), done(2), tl_mark(a, 202);  // q_ca
                    wait_a(), seg_gemm(), done(2), tl_mark(a, 203);                // y = softmax(q) . blockdiag(A_ca)
                    wait_a(), gemm_b(2, 128, kColH, true, nullptr), done(1), tl_mark(a, 204);   // h += . Wo_ca
                    wait_a(), gemm_b(2, 64, kColW, false, nullptr), done(2), tl_mark(a, 205);   // FFN up
                    wait_a(), gemm_b(1, 128, kColW, false, nullptr), done(2), tl_mark(a, 206);  // FFN down
                    wait_a(), gemm_b(2, 128, kColH, true, nullptr), done(1), tl_mark(a, 207);   // h += . Wo_ffn
                }
                if (it + 1 < L) {
                    wait_a();
                    gemm_b(2, 128, kColS, false, nullptr);                         // q -> S[0:128]
                    const uint32_t itA0 = (uint32_t)si * 26u * (uint32_t)L + (uint32_t)(it + 1) * 26;   // Wk, Wv of layer it+1 in ring A
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
                    done(2), tl_mark(a, 208);
                    for (int ps = 0; ps < passes; ++ps) {                          // K^T V partial(s): E^T . V, MN-major images
                        wait_a();
                        const uint32_t eimg = smem_u32(xbuf), vimg = smem_u32(ringB);
                        for (int ks = 0; ks < 8; ++ks)
                            umma_f16(tmem_base + kColW, make_desc_mnmajor_sw128(eimg + ks * 2048), make_desc_mnmajor_sw128(vimg + ks * 2048), idmn, ks > 0);
                        done(2), tl_mark(a, 209);
                    }
                }
            }
        }
    } else {
        const uint32_t a_ready_addr = smem_u32(&bars->a_ready);
        const uint32_t s_free_addr = smem_u32(&bars->s_free);
        const uint32_t lq = warp & 3, cq = warp >> 2;
        const uint32_t r = lq * 32 + lane;            // row of the tile == TMEM lane
        const uint32_t c0 = cq * 32;                  // first of this thread's 32 features (heads 2cq, 2cq+1)
        const uint32_t trow = tmem_base + ((lq * 32) << 16);
        const uint32_t awork = smem_u32(awork_p);
        const long g = (long)blockIdx.x * kTileRows + r;
        const bool valid = g < a.M;
        const int b = valid ? (int)(g / a.T) : 0;
        const int t = valid ? (int)(g - (long)b * a.T) : 0;
        const bool keep = valid && (a.length == nullptr || (long long)t < a.length[b]);
        uint32_t ph[3] = {0, 0, 0};
        RowStats rs{xchg, 1 + lq, r, cq, 0};
        float mean, rstd;
        float v[32];

        for (int si = 0; si < a.n_steps; ++si) {
        const int tstep = a.step0 - si;                                  // timestep index of this step
        const float* x_src = si == 0 ? a.x_in : a.x_out;
        // ---- step prologue (was step_begin_kernel): this tile's A_emb = SiLU(te + xp) image -> global (streamed back
        //      24 times by ring A), h0 = joint_embed(x) + sequence_embedding -> TMEM (stays there for the whole step).
        //      Small operands are staged through the k-image buffer (idle until the first reduction): with 221 KB of
        //      shared memory the L1 is only a few KB, so repeated global reads of weights would all go to L2.
        {
            float* sWj = reinterpret_cast<float*>(xbuf);                  // [26][128]
            float* sbj = sWj + kP * kD;                                   // [128]
            float* sx = sbj + kD;                                         // [128][26] x rows of this tile
            float* ste = sx + kTileRows * kP;                             // [2][512] time embedding of the tile's (<= 2) clips
            const int tx = threadIdx.x;
            const long row0g = (long)blockIdx.x * kTileRows;
            const int clip0 = (int)(row0g / a.T);
            {   // every global load of the staging phase is issued before the first store (one L2 round trip)
                float tw[7], tv[7], tt[2];
                const long xlim = (long)a.M * kP - row0g * kP;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int i = tx + j * kRowThreads;
                    tw[j] = i < kP * kD ? __ldg(a.WjT + i) : 0.f;
                    tv[j] = (i < kTileRows * kP && i < xlim) ? __ldcg(x_src + row0g * kP + i) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int i = tx + j * kRowThreads, cl = clip0 + (i >> 9);
                    tt[j] = __ldcg(a.te + (size_t)tstep * a.te_step_stride + ((long)cl * a.T < (long)a.M ? (size_t)cl * a.te_stride : 0) + (i & 511));
                }
                const float tb = tx < kD ? __ldg(a.bj + tx) : 0.f;
                const float4 psa = tx < 96 ? __ldg(reinterpret_cast<const float4*>(a.prm) + tx) : make_float4(0.f, 0.f, 0.f, 0.f);   // SA biases of layer 0
                const float4* ps4 = reinterpret_cast<const float4*>(a.pos + (size_t)t * kD + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 pv = valid ? __ldg(ps4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                    v[4 * i] = pv.x, v[4 * i + 1] = pv.y, v[4 * i + 2] = pv.z, v[4 * i + 3] = pv.w;
                }
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                    const int i = tx + j * kRowThreads;
                    if (i < kP * kD) sWj[i] = tw[j], sx[i] = tv[j];
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) ste[tx + j * kRowThreads] = tt[j];
                if (tx < kD) sbj[tx] = tb;
                if (tx < 96) reinterpret_cast<float4*>(prm_sa)[tx] = psa;
            }
            named_bar_sync(5, kRowThreads);
            if (tx == 0) tl_mark(a, 128);
            // h0 for this thread's 32 features (v already holds the sequence embedding)
            {
                uint64_t hv[16];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 bv = *reinterpret_cast<const float4*>(sbj + c0 + 4 * i);
                    hv[2 * i] = pk2(bv.x + v[4 * i], bv.y + v[4 * i + 1]), hv[2 * i + 1] = pk2(bv.z + v[4 * i + 2], bv.w + v[4 * i + 3]);
                }
#pragma unroll 2
                for (int c = 0; c < kP; ++c) {
                    const float xc = sx[r * kP + c];
                    const uint64_t xc2 = pk2(xc, xc);
                    const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(sWj + c * kD + c0);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const ulonglong2 wv = w2[i];
                        hv[2 * i] = ffma2(xc2, wv.x, hv[2 * i]), hv[2 * i + 1] = ffma2(xc2, wv.y, hv[2 * i + 1]);
                    }
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) upk2(hv[i], v[2 * i], v[2 * i + 1]);
            }
            tmem_st32(trow + kColH + c0, v);
            tmem_wait_st();
            if (tx == 0) tl_mark(a, 129);
        }

        for (int it = -1; it < L; ++it) {
            if (it >= 0) {
            // ================= self-attention tail: y = q . blockdiag(A_sa) (tensor cores) ; h += Styl(y)
                rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 101);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                row_stats32(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 102);                                   // S = A_emb . We_sa
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStSa, awork, r, c0);
                tc_fence_before();                                           // S consumed: the next FiLM projection may start
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free_addr);
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 151);                                          // -> h += A . Wo_sa
    
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
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 152);                                          // -> W = LN(h) . Wq_ca
                rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 104);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmCaBq + c0);
                softmax16(v);
                softmax16(v + 16);
                store_a16<kBf16>(awork, r, c0, v);
                store_a16<kBf16>(awork, r, c0 + 16, v + 16);
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 153);                                          // -> W = softmax(q) . blockdiag(A_ca)
                rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 105);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                row_stats32(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 106);                                   // S = A_emb . We_ca
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStCa, awork, r, c0);
                tc_fence_before();                                           // S consumed: the next FiLM projection may start
                __syncwarp();
                if (lane == 0) mbar_arrive(s_free_addr);
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 154);                                          // -> h += A . Wo_ca
    
                // ================= FFN (no pre-norm, reference transformer.py:170-173)
                rows_wait(bars, 1, ph[1]); if (threadIdx.x == 0) tl_mark(a, 107);
                tmem_ld32(trow + kColH + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmStCa + kStBo + c0);                  // deferred bias of Wo_ca
                tmem_st32(trow + kColH + c0, v);
                store_a16<kBf16>(awork, r, c0, v);
                store_a16<kBf16>(awork, r, c0 + 16, v + 16);
                tmem_wait_st();
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 155);                                          // -> W[0:64] = h . W1
                rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 108);
                {
                    float u[16];                                             // hidden 64 = 4 quarters of 16
                    tmem_ld16(trow + kColW + 16 * cq, u);
                    tmem_wait_ld();
    #pragma unroll
                    for (int i = 0; i < 16; ++i) u[i] = gelu_erf_f(u[i] + prm[kPrmFfB1 + 16 * cq + i]);
                    store_a16<kBf16>(awork, r, 16 * cq, u);
                }
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 156);                                          // -> W = GELU(.) . W2
                rows_wait(bars, 2, ph[2]); if (threadIdx.x == 0) tl_mark(a, 109);
                tmem_ld32(trow + kColW + c0, v);
                tmem_wait_ld();
                add_bias32(v, prm + kPrmFfB2 + c0);
                row_stats32(rs, v, mean, rstd);
                rows_wait(bars, 0, ph[0]); if (threadIdx.x == 0) tl_mark(a, 110);                                   // S = A_emb . We_ffn
                film_to_a<kBf16>(trow, v, mean, rstd, prm + kPrmStFf, awork, r, c0);
                rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 157);                                          // -> h += A . Wo_ffn
                rows_wait(bars, 1, ph[1]); if (threadIdx.x == 0) tl_mark(a, 111);
            }

            // ---- residual stream after layer `it` (deferred bias of the last FFN block)
            tmem_ld32(trow + kColH + c0, v);
            tmem_wait_ld();
            if (it >= 0) {
                add_bias32(v, prm + kPrmStFf + kStBo + c0);
                if (it + 1 < L) tmem_st32(trow + kColH + c0, v);        // h keeps living in TMEM
            }
            if (it + 1 == L) {
                // ---- step epilogue (was out_update_kernel): pred_x0 = h . Wout^T + b (fp32), sampler update of x.
                //      Scratch: the operand buffers (partial dot products) and the parameter block (output head); the
                //      rings are left alone -- the producers are already streaming the next step's first weights.
                float* part = reinterpret_cast<float*>(awork_p);          // [4 cq][128 rows][28] (awork | xbuf, 64 KB)
                float* sWo = prm;                                         // [128][32] over prm | prm_sa | xchg | red
                float* sbo = sWo + kD * 32;                               // [32]
                static_assert(4 * kTileRows * 28 * 4 <= 2 * kAworkBytes, "partials fit awork | xbuf");
                static_assert(kD * 32 + 32 <= kPrmFloats + 384 + 1024 + kRedFloats, "output head fits the parameter block");
                const int tx = threadIdx.x;
                const long row0g = (long)blockIdx.x * kTileRows;
                const int nel = (int)min((long)kTileRows * kP, (long)a.M * kP - row0g * kP);
                const int smode = a.mode & 0xF;
                const float* nzp = a.noise != nullptr ? a.noise + (size_t)si * a.noise_stride : nullptr;
                {
                    const float4* wo4 = reinterpret_cast<const float4*>(a.WoT);
                    const float4 w0 = __ldg(wo4 + tx), w1 = __ldg(wo4 + tx + kRowThreads);
                    const float tb = tx < kP ? __ldg(a.bo + tx) : 0.f;
                    named_bar_sync(5, kRowThreads);                        // every thread has read its last bias from prm
                    reinterpret_cast<float4*>(sWo)[tx] = w0, reinterpret_cast<float4*>(sWo)[tx + kRowThreads] = w1;
                    if (tx < 32) sbo[tx] = tb;
                }
                named_bar_sync(5, kRowThreads);
                {
                    uint64_t acc[14];
#pragma unroll
                    for (int p = 0; p < 14; ++p) acc[p] = 0ull;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const ulonglong2* w2 = reinterpret_cast<const ulonglong2*>(sWo + (size_t)(c0 + i) * 32);
                        const uint64_t hv = pk2(v[i], v[i]);
#pragma unroll
                        for (int q4 = 0; q4 < 7; ++q4) {
                            const ulonglong2 wv = w2[q4];
                            acc[2 * q4] = ffma2(hv, wv.x, acc[2 * q4]), acc[2 * q4 + 1] = ffma2(hv, wv.y, acc[2 * q4 + 1]);
                        }
                    }
#pragma unroll
                    for (int p4 = 0; p4 < 7; ++p4)
                        *reinterpret_cast<ulonglong2*>(part + ((size_t)cq * kTileRows + r) * 28 + 4 * p4) = make_ulonglong2(acc[2 * p4], acc[2 * p4 + 1]);
                }
                {   // 128 rows x 26 outputs over 512 threads; x and noise are in flight while the partials settle
                    float to[7], tn[7];
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const int i = tx + j * kRowThreads;
                        to[j] = (smode != 0 && i < nel) ? __ldcg(x_src + row0g * kP + i) : 0.f;
                        tn[j] = (smode != 0 && nzp != nullptr && i < nel) ? __ldcg(nzp + row0g * kP + i) : 0.f;
                    }
                    named_bar_sync(5, kRowThreads);
                    const float* cf = smode != 0 ? a.coef + (size_t)tstep * 8 : nullptr;
                    float* x0p = a.x0_out + (size_t)si * a.x0_stride;
#pragma unroll
                    for (int j = 0; j < 7; ++j) {
                        const int i = tx + j * kRowThreads;
                        if (i < nel) {
                            const int rr = i / kP, p = i - rr * kP;
                            float x0 = sbo[p] + ((part[(0 * kTileRows + rr) * 28 + p] + part[(1 * kTileRows + rr) * 28 + p]) +
                                                 (part[(2 * kTileRows + rr) * 28 + p] + part[(3 * kTileRows + rr) * 28 + p]));
                            if (a.mode & 0x10) x0 = fminf(fmaxf(x0, -1.f), 1.f);
                            x0p[row0g * kP + i] = x0;
                            if (smode != 0) {
                                const float xn = smode == 1 ? ddim_rule(to[j], x0, cf, tn[j]) : ddpm_rule(to[j], x0, cf, tn[j]);
                                a.x_out[row0g * kP + i] = xn;
                                if (a.x_trace != nullptr) a.x_trace[(size_t)si * a.M * kP + row0g * kP + i] = xn;
                            }
                        }
                    }
                }
                named_bar_sync(5, kRowThreads);                            // the scratch is the next step's staging area
                break;
            }

            // ================= self-attention head of layer it+1: LN -> q | k | v, then the time-axis reduction
            row_stats32(rs, v, mean, rstd);
            normalize32(v, mean, rstd);                                  // LN affine folded into Wq/Wk/Wv
            store_a16<kBf16>(awork, r, c0, v);
            store_a16<kBf16>(awork, r, c0 + 16, v + 16);
            tmem_wait_st();
            rows_publish<false>(a_ready_addr, lane); if (threadIdx.x == 0) tl_mark(a, 158);
            if (it < 0) {
                // ---- rest of the step prologue, overlapped with the first q|k|v MMAs: this tile's A_emb image -> global
                const int tx = threadIdx.x;
                const long row0g = (long)blockIdx.x * kTileRows;
                const int clip0 = (int)(row0g / a.T);
                const float* ste = reinterpret_cast<const float*>(xbuf) + kP * kD + kD + kTileRows * kP;
            uint8_t* img = a.aemb_out + (size_t)blockIdx.x * 8 * kStageABytes;
                for (int k0 = 0; k0 < 16; k0 += 4) {                 // 128 rows x 64 chunks of 8 features; a warp = 1 KB of one row
                    float4 xa[4][2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int task = (k0 + k) * kRowThreads + tx;
                        const long gg = min(row0g + (task >> 6), (long)a.M - 1);
                        const float4* xr4 = reinterpret_cast<const float4*>(a.xp + gg * kE + (task & 63) * 8);
                        xa[k][0] = __ldg(xr4), xa[k][1] = __ldg(xr4 + 1);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int task = (k0 + k) * kRowThreads + tx;
                        const int row = task >> 6, ch = task & 63;
                        const long gg = row0g + row;
                        const int bb = (gg < a.M && gg >= (long)(clip0 + 1) * a.T) ? 1 : 0;
                        const float4* tr = reinterpret_cast<const float4*>(ste + bb * kE + ch * 8);
                        const float4 a0 = xa[k][0], a1 = xa[k][1], t0 = tr[0], t1 = tr[1];
                        const float e8[8] = {a0.x + t0.x, a0.y + t0.y, a0.z + t0.z, a0.w + t0.w, a1.x + t1.x, a1.y + t1.y, a1.z + t1.z, a1.w + t1.w};
                        uint32_t p[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            p[i] = pack2<kBf16>(__fdividef(e8[2 * i], 1.f + __expf(-e8[2 * i])), __fdividef(e8[2 * i + 1], 1.f + __expf(-e8[2 * i + 1])));
                        const uint4 pk = gg < a.M ? make_uint4(p[0], p[1], p[2], p[3]) : make_uint4(0, 0, 0, 0);
                        *reinterpret_cast<uint4*>(img + (size_t)(ch >> 3) * kABlockBytes + sw128_offset(row, ch & 7)) = pk;
                    }
                }
                __threadfence();                                       // the image must have reached L2 ...
                asm volatile("fence.proxy.async;" ::: "memory");      // ... and be ordered before the bulk-copy (async proxy) reads
                named_bar_sync(5, kRowThreads);                        // every row thread is done with the staging area in xbuf
                if (lane == 0) mbar_arrive(smem_u32(&bars->aemb_ready));
                if (tx == 0) tl_mark(a, 127);
            }
            rows_wait(bars, 2, ph[2]);
            if (threadIdx.x == 0) tl_mark(a, 112);
            {
                // Fused time-axis softmax + K^T V (reference :111,:117) on the tensor cores.  With T >= 128 a tile touches
                // at most two clips.  Two 32 KB operand-image buffers X, Y are all the scratch it needs:
                //   k (16-bit) -> X ; per-segment column maxima by a column scan of X ; E = exp(k - max) -> X ;
                //   V -> Y (rows of the other segment zeroed) ; P = E^T V as 8 MN-major MMAs over the tile's tokens ;
                //   per-segment column sums by a column scan of the E image (same rounded values as the MMA sees).
                // The partial (max, sum, diagonal 16x16 blocks of P) goes to global memory; the CTA that completes a clip
                // merges its partials (online-softmax rescaling) into the clip's block-diagonal B-operand image.
                float* pm = red;                                       // [8 rg][2 seg][128] exchange (max, then sums)
                float* msm = pm + 2048;                                // [2][128] maxima
                float* ssm = msm + 256;                                // [2][128] sums
                int* flags = reinterpret_cast<int*>(ssm + 256);
                uint8_t* Xp = xbuf;
                const uint32_t eimg = smem_u32(xbuf), vimg = smem_u32(ringB);
                const int row0 = blockIdx.x * kTileRows;
                const int first_clip = row0 / a.T;
                const int nvalid = max(0, min(kTileRows, a.M - row0));
                const int e = min(nvalid, (first_clip + 1) * a.T - row0);   // rows [0,e): first clip, [e,nvalid): next clip
                const int n_seg = nvalid == 0 ? 0 : (nvalid > e ? 2 : 1);
                const int tx = threadIdx.x;
                const int seq = si * L + it + 1;                             // reductions completed so far in this launch
                const bool in_tile = (int)r < nvalid;
                const int myseg = (int)r >= e ? 1 : 0;
                const int col = tx & 127, qr = tx >> 7;
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
                {   // column maxima per segment: this thread scans 16 rows of a column PAIR (packed 16-bit max)
                    uint32_t m0 = kNegInf2, m1 = kNegInf2;
#pragma unroll
                    for (int rr = 0; rr < 16; ++rr) {
                        const int row = 16 * rg + rr;
                        const uint32_t x = *pair_at(row);
                        if (row < e) m0 = max2(m0, x);
                        else m1 = max2(m1, x);
                    }
                    const float2 f0 = unpack2<kBf16>(m0), f1 = unpack2<kBf16>(m1);
                    *reinterpret_cast<float2*>(pm + (rg * 2 + 0) * 128 + 2 * cp) = f0;
                    *reinterpret_cast<float2*>(pm + (rg * 2 + 1) * 128 + 2 * cp) = f1;
                }
                named_bar_sync(5, kRowThreads);
                if (tx < 256) {
                    const int sg = tx >> 7;
                    float mm = pm[sg * 128 + col];
#pragma unroll
                    for (int q8 = 1; q8 < 8; ++q8) mm = fmaxf(mm, pm[(q8 * 2 + sg) * 128 + col]);
                    msm[tx] = mm;
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
                    rows_publish<false>(a_ready_addr, lane);                      // -> W = E^T . V  (8 MMAs over the tokens)
                    if (tx == 0) tl_mark(a, 122);
                    if (ps == 0) {
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
                        named_bar_sync(5, kRowThreads);                    // E image complete (all rows published)
                        float2 s0 = make_float2(0.f, 0.f), s1 = s0;          // column sums per segment from the rounded E
#pragma unroll
                        for (int rr = 0; rr < 16; ++rr) {
                            const int row = 16 * rg + rr;
                            const float2 x = unpack2<kBf16>(*pair_at(row));
                            if (row < e) s0.x += x.x, s0.y += x.y;
                            else s1.x += x.x, s1.y += x.y;
                        }
                        *reinterpret_cast<float2*>(pm + (rg * 2 + 0) * 128 + 2 * cp) = s0;
                        *reinterpret_cast<float2*>(pm + (rg * 2 + 1) * 128 + 2 * cp) = s1;
                        named_bar_sync(5, kRowThreads);
                        if (tx < 256) {
                            const int sg = tx >> 7;
                            float ss = pm[sg * 128 + col];
#pragma unroll
                            for (int q8 = 1; q8 < 8; ++q8) ss += pm[(q8 * 2 + sg) * 128 + col];
                            ssm[tx] = ss;
                        }
                    }
                    rows_wait(bars, 2, ph[2]);
                    if (tx == 0) tl_mark(a, 123);
                    if (ps < n_seg) {
                        float* P = a.kv_part + (((size_t)(seq & 1) * gridDim.x + blockIdx.x) * 2 + ps) * kKvPartFloats;
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
                // ---- publish: one arrival per (tile, clip); the counters only grow within a step (zeroed by the host)
                named_bar_sync(5, kRowThreads);
                if ((tx == 0 || tx == 32) && (tx >> 5) < n_seg) {
                    __threadfence();                                       // cumulative: orders the whole CTA's partials (bar.sync above)
                    atomicAdd(a.clip_cnt + first_clip + (tx >> 5), 1);
                }
                // ---- meanwhile: clear the two image buffers (E / V are dead: the last K^T V pass has completed) and
                //      fetch the next layer's parameters
                {
                    const uint4 z4 = make_uint4(0, 0, 0, 0);
                    for (int i = tx; i < kAworkBytes / 16; i += kRowThreads) {
                        reinterpret_cast<uint4*>(xbuf)[i] = z4;
                        reinterpret_cast<uint4*>(ringB)[i] = z4;
                    }
                    // all loads in flight before the first store (one L2 round trip, not six)
                    const float4* pn4 = reinterpret_cast<const float4*>(a.prm + (size_t)(it + 1) * kPrmFloats);
                    static_assert(kPrmFloats % 4 == 0 && kPrmFloats / 4 <= 2 * kRowThreads, "parameter block layout");
                    const float4 p0 = __ldg(pn4 + tx);
                    const float4 p1 = tx + kRowThreads < kPrmFloats / 4 ? __ldg(pn4 + tx + kRowThreads) : make_float4(0.f, 0.f, 0.f, 0.f);
                    const float4 p2 = (it + 2 < L && tx < 96) ? __ldg(pn4 + kPrmFloats / 4 + tx) : make_float4(0.f, 0.f, 0.f, 0.f);
                    reinterpret_cast<float4*>(prm)[tx] = p0;
                    if (tx + kRowThreads < kPrmFloats / 4) reinterpret_cast<float4*>(prm)[tx + kRowThreads] = p1;
                    if (it + 2 < L && tx < 96) reinterpret_cast<float4*>(prm_sa)[tx] = p2;
                }
                // ---- wait until every tile of this tile's clip(s) has published, then merge the partials (online-softmax
                //      rescaling) straight into block-diagonal B-operand images in shared memory: segment 0 -> xbuf,
                //      segment 1 -> ring B.  Every tile does this for itself: no second global round trip.
                if (tx == 0) tl_mark(a, 121);
                if ((tx == 0 || tx == 32) && (tx >> 5) < n_seg) {
                    const int clip = first_clip + (tx >> 5);
                    const int ntiles = ((clip + 1) * a.T - 1) / kTileRows - (clip * a.T) / kTileRows + 1;
                    int cnt;
                    do {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(cnt) : "l"(a.clip_cnt + clip) : "memory");
                    } while (cnt < ntiles * (seq + 1));
                }
                named_bar_sync(5, kRowThreads);
                if (tx == 0) tl_mark(a, 125);
                const int hh = tx >> 6, sub = tx & 63, d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
                for (int sg = 0; sg < n_seg; ++sg) {
                    const int clip = first_clip + sg;
                    const int t_first = (clip * a.T) / kTileRows, t_last = ((clip + 1) * a.T - 1) / kTileRows;
                    float M0 = -INFINITY, M1 = -INFINITY, a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
                    auto part_of = [&](int ti) {
                        return a.kv_part + (((size_t)(seq & 1) * gridDim.x + ti) * 2 + (clip - (ti * kTileRows) / a.T)) * kKvPartFloats;
                    };
                    if (t_last - t_first < 4) {
                        // short clips (<= 4 tiles): issue every load first, one L2 round trip for the whole merge
                        float mi0[4], mi1[4], si0[4], si1[4];
                        float2 r0[4], r1[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const bool on = t_first + j <= t_last;
                            const float* P = part_of(on ? t_first + j : t_first);
                            const float* Pa = P + 256 + hh * 256;
                            mi0[j] = on ? __ldcg(P + 16 * hh + d0) : -INFINITY;
                            mi1[j] = on ? __ldcg(P + 16 * hh + d0 + 1) : -INFINITY;
                            si0[j] = __ldcg(P + 128 + 16 * hh + d0), si1[j] = __ldcg(P + 128 + 16 * hh + d0 + 1);
                            r0[j] = __ldcg(reinterpret_cast<const float2*>(Pa + d0 * 16 + l0));
                            r1[j] = __ldcg(reinterpret_cast<const float2*>(Pa + (d0 + 1) * 16 + l0));
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) M0 = fmaxf(M0, mi0[j]), M1 = fmaxf(M1, mi1[j]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float w0 = __expf(mi0[j] - M0), w1 = __expf(mi1[j] - M1);      // exp(-inf) = 0 for absent tiles
                            s0 = fmaf(si0[j], w0, s0), s1 = fmaf(si1[j], w1, s1);
                            a00 = fmaf(r0[j].x, w0, a00), a01 = fmaf(r0[j].y, w0, a01);
                            a10 = fmaf(r1[j].x, w1, a10), a11 = fmaf(r1[j].y, w1, a11);
                        }
                    } else {
                        for (int ti = t_first; ti <= t_last; ++ti) {                              // online rescaling
                            const float* P = part_of(ti);
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
                    }
                    uint8_t* img = sg == 0 ? xbuf : ringB;
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
                rows_publish<false>(a_ready_addr, lane);                          // -> y = q . blockdiag(A_sa) of layer it+1
                if (tx == 0) tl_mark(a, 126);
            }
        }
        }   // launch step si
    }
    if (threadIdx.x == 0) tl_mark(a, 2);
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

constexpr int kStepSmemBytes = kPRingAStages * kStageBytes + kRingBStages * kRingBStageBytes + 2 * kAworkBytes +
                               (kPrmFloats + 384) * 4 + 512 * 8 + kRedFloats * 4 + sizeof(LayerBarriers) + 1024;
static_assert(kStepSmemBytes <= 232448, "step kernel exceeds the 227 KB shared-memory limit");

}  // namespace dc
