// M2SNet music encoder on the tensor cores: the six 16/32-channel 3x3 convolutions of
// MusicEncoder.forward (reference Diffusion_Stage/models/transformer.py:289-340) as implicit GEMMs on tcgen05, fp32-equivalent.
//
// Activations live in HBM as NHWC "split" pixels: [C x hi | C x lo] bf16 with value = hi + lo (16 mantissa bits).  For every tap of
// the 3x3 window the 128 pixels of an M tile are gathered (reflect padding by index arithmetic, reads served by L1) straight into a
// K-major SWIZZLE_128B A block; the BatchNorm-folded weights sit in shared memory as B blocks [w_hi | w_hi] and [w_lo], so
//      a . w  ~=  a_hi w_hi + a_lo w_hi + a_hi w_lo        (the dropped a_lo w_lo term is 2^-18 relative)
// accumulates in fp32 in TMEM over 9 taps x (4 + 2) MMAs of M = 128, N = C_out, K = 16.  The epilogue (bias, ReLU, identity or
// 1x1-convolution residual, re-split) runs on the 128 row threads while the MMA warp already works on the next tile (two
// accumulators).  Byte-bound on the gather (1.1 KB of L1 reads + 1.1 KB of shared-memory writes per pixel), not on the tensor pipe.
#pragma once
#include "tc_common.cuh"

namespace dc {

__device__ __forceinline__ int me_reflect(int i, int n) {      // padding_mode='reflect', pad 1
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}
// one 256-bit store (sm_100: STG.256): a full 32-byte sector per thread instead of two half-sector writes; the activations are read
// next by another kernel, so the lines are not allocated in L1 (-2.5 % on the whole encoder)
__device__ __forceinline__ void st_global_v8(void* p, const uint4& a, const uint4& b) {
    asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x),
                 "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void st_global_na_v4(void* p, const uint4& a) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w) : "memory");
}
// 8 fp32 -> hi chunk, lo chunk (8 bf16 each)
__device__ __forceinline__ void me_split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        h[i] = pack2<true>(v[2 * i], v[2 * i + 1]);
        const float2 b = unpack2<true>(h[i]);
        l[i] = pack2<true>(v[2 * i] - b.x, v[2 * i + 1] - b.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]), lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void me_join8(const uint4& hi, const uint4& lo, float* v) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 a = unpack2<true>(h[i]), b = unpack2<true>(l[i]);
        v[2 * i] = a.x + b.x, v[2 * i + 1] = a.y + b.y;
    }
}

// ---------------------------------------------------------------- conv1.0: 1 -> 16 channels from the fp32 mel (CUDA cores)
struct alignas(16) Conv10Weights {
    float w[9 * 16];      // [tap][co], BatchNorm folded
    float b[16];
};
// mel [B][H][W] fp32 -> y [B][H][W][16 hi | 16 lo]; no residual (reference MusicEncoder.conv1[0], residual=False)
__global__ void __launch_bounds__(256) conv10_split_kernel(const float* __restrict__ mel, uint16_t* __restrict__ y, int H, int W,
                                                          const __grid_constant__ Conv10Weights cw) {
    pdl_wait();                                                  // (every encoder kernel is launched as a programmatic dependent of the one before it)
    const long q = (long)blockIdx.x * 256 + threadIdx.x;
    if (q >= (long)H * W) return;
    const int yy = (int)(q / W), xx = (int)(q % W);
    const float* mb = mel + (size_t)blockIdx.y * H * W;
    // two output channels per FFMA2 (fma.rn.f32x2: the same IEEE fma per lane, half the issue slots of this math-bound kernel)
    uint64_t acc2[8];
    const uint64_t* w2 = reinterpret_cast<const uint64_t*>(cw.w);
#pragma unroll
    for (int c = 0; c < 8; ++c) acc2[c] = reinterpret_cast<const uint64_t*>(cw.b)[c];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            const float v = __ldg(mb + (size_t)me_reflect(yy + dy - 1, H) * W + me_reflect(xx + dx - 1, W));
            const uint64_t vv = pk2(v, v);
#pragma unroll
            for (int c = 0; c < 8; ++c) acc2[c] = ffma2(vv, w2[(dy * 3 + dx) * 8 + c], acc2[c]);
        }
    float acc[16];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        upk2(acc2[c], acc[2 * c], acc[2 * c + 1]);
        acc[2 * c] = fmaxf(acc[2 * c], 0.f), acc[2 * c + 1] = fmaxf(acc[2 * c + 1], 0.f);
    }
    uint4 hi0, lo0, hi1, lo1;
    me_split8(acc, hi0, lo0);
    me_split8(acc + 8, hi1, lo1);
    uint4* dst = reinterpret_cast<uint4*>(y + ((size_t)blockIdx.y * H * W + q) * 32);
    st_global_v8(dst, hi0, hi1);
    st_global_v8(dst + 2, lo0, lo1);
}

// ---------------------------------------------------------------- 3x3 convolutions, 16 / 32 channels, tcgen05
// The output candidates of a clip form the raster og = y (W + 2) + x (x >= W: two junk candidates per padded row, skipped by the
// epilogue).  A BAND is NM accumulator tiles = NM x 126 consecutive candidates (not a whole number of image rows: every tile but a
// clip's last is full); the NM x 126 + 2 (W + 2) + 2 input pixels it reads (reflect halo resolved while loading) are staged ONCE in a
// shared-memory strip, one 128-byte row per pixel ([hi | lo], 16-byte chunk c of staged row s stored at c ^ (s & 7)).  The tcgen05
// SWIZZLE_128B pattern is a function of the shared-memory ADDRESS bits (checked on the GPU with tools/experiments/desc_shift.cu:
// a K-major descriptor whose start is 128- but not 1024-byte aligned reads exactly the rows it points at, with base_offset 0), so
// the A operand of window row dy is simply the SAME strip read through a descriptor shifted by dy (W + 2) rows: no im2col copy
// exists anywhere.
// An M = 128 MMA costs the same 64+ cycles for N = 16 as for N = 128, so the three taps of a window row are ONE MMA with
// N = 3 C_out: accumulator row i (staged pixel s_i) gets E_i[dx] = sum_dy pix(s_i + dy (W + 2)) . W[dy][dx] for dx = -1, 0, +1 side
// by side, and the output of pixel s is E(s - 1)[-1] + E(s)[0] + E(s + 1)[+1]: the horizontal shift happens on the OUTPUT side, in
// the epilogue, with two warp shuffles per value (lanes 0 / 127 of a tile are its halo: tiles advance by 126 rows; the lanes at
// warp boundaries go through shared memory).  9 (C_in = 16) or 18 (C_in = 32) MMAs per 126 pixels instead of 27 / 54.
// One CTA per SM, three roles: 4 producer warps stage band j + 1 into the second strip while the MMA warp issues band j (four
// accumulator tiles in TMEM, so it runs ahead) and 16 row warps in 2 or 4 sets finish the tiles (a row thread owns 16 output channels:
// its hi chunks and its lo chunks leave as one 256-bit store each).  Measured alternatives: DESIGN.md section 4.5.
constexpr int kMeNM128 = 4, kMeNM64 = 4, kMeNM64b = 4, kMeNM32 = 4;   // accumulator tiles per band (two strips of that length + the weight blocks fit 227 KB)
constexpr int kMeRowWarps = 16;           // epilogue warps: sets of 4 (C_out = 16) or 8 (C_out = 32: two threads per accumulator row); set s finishes tiles s, s + SETS, ..
constexpr int kMeProducerWarps = 4;       // cp.async staging of the next band
constexpr int kMeThreads = 32 * (kMeRowWarps + kMeProducerWarps + 1);   // + 1 MMA warp
constexpr int kMeAcc = 4;                 // accumulator tiles in TMEM (128 columns each)
constexpr int kMeTile = kTileRows - 2;    // finished pixels per accumulator tile
#ifdef DC_ME_TIMELINE
__device__ unsigned long long me_tl[4 * 1024];       // debug build only: clock64 stamps of CTA 0, thread 0 (tools/experiments)
#define ME_TL(slot) do { if (tl_on && tl_n < 1024) me_tl[tl_base + tl_n++] = ((unsigned long long)(slot) << 56) | (clock64() & 0xFFFFFFFFFFFFFFull); } while (0)
#else
#define ME_TL(slot) do { } while (0)
#endif
struct MeBarriers {
    uint64_t staged[2];                   // producers -> MMA: the strip holds the band (one arrival per producer warp)
    uint64_t strip_free[2];               // rows -> producers: every residual read of the strip is done (one arrival per row warp)
    uint64_t acc_full[kMeAcc], acc_free[kMeAcc];   // accumulators: MMA -> rows (tcgen05.commit), rows -> MMA (the warps of the owning set)
    uint32_t tmem_base;
};
template <int W, int NM>
__host__ __device__ constexpr int me_strip_rows() { return ((NM * kMeTile + 2 * (W + 2) + 4 + 7) / 8) * 8; }
template <int COUT, int RES>
__host__ __device__ constexpr int me_wrows() { return (RES == 2 ? 4 : 3) * COUT; }          // B rows per window row: [dx = -1 | 0 | +1 (| 1x1 residual)]
template <int CIN, int COUT, int RES, int W, int NM>
constexpr int me_smem_bytes() {
    return 2 * me_strip_rows<W, NM>() * 128 + 3 * me_wrows<COUT, RES>() * 128 + 2 * COUT * 4 + kMeRowWarps * 2 * 2 * 16 * 4 + (int)sizeof(MeBarriers) + 1024;
}

// x [B][H][W][CIN hi | CIN lo], y [B][H][W][COUT hi | COUT lo].  wimg: one block per window row dy, me_wrows() rows x 128 B, K-major
// SW128, row dxb * COUT + c = [w_hi (CIN) | w_lo (CIN)] of tap (dy, dxb) (rows 3 COUT + c: the 1x1 residual convolution, centre row
// only); a . w ~= a_hi w_hi + a_lo w_hi + a_hi w_lo is three (CIN = 16) or six (CIN = 32) K = 16 MMAs per window row, each picking
// its own 32-byte K chunk of the A and of the B rows.  bias [COUT] (+ [COUT] of the 1x1 residual when RES == 2).
// RES: 1 = identity residual, 2 = 1x1 convolution + BatchNorm.
template <int CIN, int COUT, int RES, int W, int NM>
__global__ void __launch_bounds__(kMeThreads, 1) conv_tc_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int H, int bands_per_clip,
                                                               int n_jobs, const uint8_t* __restrict__ wimg, const float* __restrict__ bias) {
    static_assert((CIN == 16 || CIN == 32) && (COUT == 16 || COUT == 32) && (RES == 1 || RES == 2), "unsupported shape");
    static_assert(RES == 2 || CIN == COUT, "identity residual needs CIN == COUT");
    constexpr int P = W + 2, SROWS = me_strip_rows<W, NM>();
    constexpr int WROWS = me_wrows<COUT, RES>(), BLK = WROWS * 128;
    constexpr int NCH = CIN / 4;                               // 16-byte chunks of a split pixel: NCH / 2 hi, then NCH / 2 lo
    constexpr int KH = CIN / 16;                               // K = 16 chunks of the hi (and of the lo) half
    constexpr int HC = 16;                                     // output channels per epilogue thread (its hi and its lo chunks: 32 bytes each)
    constexpr int NH = COUT / HC;                              // threads per accumulator row
    constexpr int SETW = 4 * NH, SETS = kMeRowWarps / SETW;    // warps per epilogue set; sets (each finishes every SETS-th tile)
    static_assert(kMeAcc % SETS == 0, "an accumulator must always belong to the same set");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* strip0 = smem;                                    // 2 x [SROWS x 128 B]: band j lives in strip j & 1
    uint8_t* wsm = strip0 + 2 * SROWS * 128;                   // 3 x [WROWS x 128 B]   (SROWS % 8 == 0: 1024-aligned)
    float* bsm = reinterpret_cast<float*>(wsm + 3 * BLK);      // [2 * COUT]
    float* xch = bsm + 2 * COUT;                               // [SETS][2 tile parity][2 side][4 quarter][NH][HC] lanes at the warp boundaries
    MeBarriers* bars = reinterpret_cast<MeBarriers*>(xch + SETS * 2 * 2 * 4 * NH * HC);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 3 * BLK / 16; i += kMeThreads) reinterpret_cast<uint4*>(wsm)[i] = __ldg(reinterpret_cast<const uint4*>(wimg) + i);
    for (int i = tid; i < 2 * SROWS * 8; i += kMeThreads) reinterpret_cast<uint4*>(strip0)[i] = make_uint4(0, 0, 0, 0);   // junk rows must stay finite
    if (tid < (RES == 2 ? 2 : 1) * COUT) bsm[tid] = __ldg(bias + tid);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) mbar_init(smem_u32(&bars->staged[i]), kMeProducerWarps), mbar_init(smem_u32(&bars->strip_free[i]), kMeRowWarps);
        for (int i = 0; i < kMeAcc; ++i) mbar_init(smem_u32(&bars->acc_full[i]), 1), mbar_init(smem_u32(&bars->acc_free[i]), SETW);
        mbar_fence_init();
    }
    if (warp == kMeRowWarps + kMeProducerWarps) {
        tmem_alloc(smem_u32(&bars->tmem_base), kMeAcc * 128);
        tmem_relinquish();
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                                // (launched as a programmatic dependent: everything above overlapped the previous kernel's tail)
    const uint32_t tmem_base = bars->tmem_base;
    const int my_jobs = n_jobs > (int)blockIdx.x ? (n_jobs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // a band is NM * 126 consecutive output candidates og = y P + x of the clip's padded-row raster (x >= W: junk), NOT a whole number
    // of image rows: every tile but a clip's last is full.  All three roles walk the same (band, tile) sequence; tile number mc (counted
    // over the CTA's whole life) uses accumulator mc % kMeAcc and is finished by row-warp set mc % SETS.
    auto band_of = [&](int j, int& clip, int& og0, int& nm) {
        const int job = (int)blockIdx.x + j * (int)gridDim.x;
        clip = job / bands_per_clip, og0 = (job - clip * bands_per_clip) * (NM * kMeTile);
        nm = min(NM, (H * P - og0 + kMeTile - 1) / kMeTile);
    };

    if (warp == kMeRowWarps + kMeProducerWarps) {
        // ================================================================ MMA issuer
        if (lane == 0) {
            const uint32_t idesc3 = make_idesc<true>(kTileRows, 3 * COUT), idesc4 = make_idesc<true>(kTileRows, WROWS);
            uint32_t mc = 0;
            for (int j = 0; j < my_jobs; ++j) {
                int clip, og0, nm;
                band_of(j, clip, og0, nm);
                const uint32_t sbase = smem_u32(strip0 + (j & 1) * SROWS * 128);
                mbar_wait(smem_u32(&bars->staged[j & 1]), (uint32_t)(j >> 1) & 1u);
                tc_fence_after();
                for (int m = 0; m < nm; ++m, ++mc) {
                    const uint32_t acc = mc % kMeAcc, dcol = tmem_base + acc * 128u;
                    mbar_wait(smem_u32(&bars->acc_free[acc]), ((mc / kMeAcc) & 1u) ^ 1u);
                    tc_fence_after();
                    // lane i of this tile <-> output candidate o = kMeTile m + i - 1, centre pixel = staged row o + P + 2.  The centre
                    // window row goes first: with RES == 2 it is the one MMA group that also writes the residual columns.
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        const int dy = q == 0 ? 0 : (q == 1 ? -1 : 1);
                        const uint64_t ad = make_desc_kmajor_sw128(sbase + (uint32_t)(m * kMeTile + 1 + (1 + dy) * P) * 128u);
                        const uint64_t bd = make_desc_kmajor_sw128(smem_u32(wsm + (dy + 1) * BLK));
                        const uint32_t idesc = dy == 0 ? idesc4 : idesc3;
#pragma unroll
                        for (int k = 0; k < KH; ++k) umma_f16(dcol, ad + 2 * k, bd + 2 * k, idesc, (q > 0 || k > 0) ? 1u : 0u);            // a_hi . w_hi
#pragma unroll
                        for (int k = 0; k < KH; ++k) umma_f16(dcol, ad + 2 * (KH + k), bd + 2 * k, idesc, 1u);                              // a_lo . w_hi
#pragma unroll
                        for (int k = 0; k < KH; ++k) umma_f16(dcol, ad + 2 * k, bd + 2 * (KH + k), idesc, 1u);                              // a_hi . w_lo
                    }
                    umma_commit(smem_u32(&bars->acc_full[acc]));
                }
            }
        }
    } else if (warp >= kMeRowWarps) {
        // ================================================================ producers: stage band j into strip j & 1 while band j - 1 computes
        const int pt = tid - kMeRowWarps * 32;
        constexpr int PT = kMeProducerWarps * 32;
        for (int j = 0; j < my_jobs; ++j) {
            int clip, og0, nm;
            band_of(j, clip, og0, nm);
            uint8_t* strip = strip0 + (j & 1) * SROWS * 128;
            const uint16_t* xc = x + (size_t)clip * H * W * (2 * CIN);
            mbar_wait(smem_u32(&bars->strip_free[j & 1]), ((uint32_t)(j >> 1) & 1u) ^ 1u);
            // image row by image row: the source of a row is contiguous, so a thread's chunk q = pt, pt + 128, .. needs one shift and one
            // swizzle; cp.async holds no registers per copy, so every chunk of the band is in flight at once
            const int npix = nm * kMeTile + 2 * P + 2;                        // staged pixel sp <-> padded raster index og0 + sp
            const int ys0 = og0 / P, ys1 = (og0 + npix - 1) / P;
            constexpr int CPR = W * NCH;                                      // 16-byte chunks of an image row
            static_assert(CPR % PT == 0, "row chunks must divide over the staging threads");
            for (int ys = ys0; ys <= ys1; ++ys) {
                const uint4* srow = reinterpret_cast<const uint4*>(xc + (size_t)me_reflect(ys - 1, H) * W * (2 * CIN));
                const int spr = ys * P - og0;                                 // staged index of this row's left halo pixel (may be < 0)
                if (spr + 1 >= 0 && spr + W < npix) {                           // the whole image row is inside the band (all but its first and last)
                    // chunk q = pt + k PT lands PT / NCH strip rows (a multiple of 8) below chunk pt: same swizzle phase, so one
                    // offset per image row and thread, then a constant stride
                    static_assert((PT / NCH) % 8 == 0, "swizzle phase must repeat");
                    uint8_t* d0 = strip + sw128_offset((uint32_t)(2 + spr + pt / NCH), (uint32_t)(pt % NCH));
                    const uint4* s0 = srow + pt;
#pragma unroll
                    for (int k = 0; k < CPR / PT; ++k) cp_async16(d0 + k * (PT / NCH) * 128, s0 + k * PT);
                } else {
#pragma unroll
                    for (int q = pt; q < CPR; q += PT) {
                        const int sp = spr + 1 + q / NCH;
                        if (sp >= 0 && sp < npix) cp_async16(strip + sw128_offset((uint32_t)(1 + sp), (uint32_t)(q % NCH)), srow + q);
                    }
                }
                if (pt < 2 * NCH) {                                           // reflect halo: x = -1 -> 1, x = W -> W - 2
                    const int side = pt / NCH, ch = pt % NCH, sp = spr + (side ? P - 1 : 0);
                    if (sp >= 0 && sp < npix) cp_async16(strip + sw128_offset((uint32_t)(1 + sp), (uint32_t)ch), srow + (side ? W - 2 : 1) * NCH + ch);
                }
            }
            cp_async_wait_all();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->staged[j & 1]));
        }
    } else {
        // ================================================================ epilogue: SETS sets of 4 NH row warps; warp w of a set reads the TMEM
        // lanes 32 (w % 4) .., and with C_out = 32 warps w and w + 4 share them (`half` selects which 16 of the output channels)
        const int set = warp / SETW, qw = warp & 3, half = (warp >> 2) % NH, r = qw * 32 + lane;
        const uint32_t trow = tmem_base + ((uint32_t)(qw * 32) << 16);
        uint32_t mc = 0;
#ifdef DC_ME_TIMELINE
        const bool tl_on = blockIdx.x == 0 && tid == 0;
        const int tl_base = (W == 128 ? 0 : (W == 32 ? 3 : (CIN == 16 ? 1 : 2))) * 1024;
        int tl_n = 0;
#endif
        for (int j = 0; j < my_jobs; ++j) {
            int clip, og0, nm;
            band_of(j, clip, og0, nm);
            const uint8_t* strip = strip0 + (j & 1) * SROWS * 128;
            ME_TL(1);
            for (int m = 0; m < nm; ++m, ++mc) {
                if ((int)(mc % SETS) != set) continue;
                const uint32_t acc = mc % kMeAcc, par = (mc / SETS) & 1u;
                const int o = m * kMeTile + r - 1, og = max(og0 + o, 0), yy = og / P, xx = og - yy * P;
                const bool valid = r >= 1 && r <= kMeTile && xx < W && yy < H;
                ME_TL(4);
                mbar_wait(smem_u32(&bars->acc_full[acc]), (mc / kMeAcc) & 1u);
                ME_TL(5);
                tc_fence_after();
                float el[HC], ec[HC], er[HC], v2[RES == 2 ? HC : 1];      // E[dx = -1], E[0], E[+1] of THIS row
                tmem_ld16(trow + acc * 128u + half * HC, el), tmem_ld16(trow + acc * 128u + COUT + half * HC, ec);
                tmem_ld16(trow + acc * 128u + 2 * COUT + half * HC, er);
                if (RES == 2) tmem_ld16(trow + acc * 128u + 3 * COUT + half * HC, v2);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(smem_u32(&bars->acc_free[acc]));
                // ---- horizontal shift on the output side: out(i) = E(i - 1)[-1] + E(i)[0] + E(i + 1)[+1]
                float* xa = xch + ((((size_t)(set * 2 + par) * 2 + 0) * 4 + qw) * NH + half) * HC;     // lane 31's E[-1] of this quarter (for lane 0 of the next)
                float* xb = xch + ((((size_t)(set * 2 + par) * 2 + 1) * 4 + qw) * NH + half) * HC;     // lane 0's E[+1] of this quarter (for lane 31 of the previous)
                if (lane == 31) {
#pragma unroll
                    for (int c = 0; c < HC; ++c) xa[c] = el[c];
                }
                if (lane == 0) {
#pragma unroll
                    for (int c = 0; c < HC; ++c) xb[c] = er[c];
                }
                named_bar_sync(2 + set, 32 * SETW);
                // (from here on two channels per instruction: add.rn.f32x2 / fma.rn.f32x2 round each lane exactly like FADD / FFMA)
                uint64_t e2[HC / 2];
#pragma unroll
                for (int p2 = 0; p2 < HC / 2; ++p2) {              // interior lanes: two shuffles per value, selects instead of branches
                    const float sl0 = __shfl_up_sync(0xFFFFFFFFu, el[2 * p2], 1), sl1 = __shfl_up_sync(0xFFFFFFFFu, el[2 * p2 + 1], 1);
                    const float sr0 = __shfl_down_sync(0xFFFFFFFFu, er[2 * p2], 1), sr1 = __shfl_down_sync(0xFFFFFFFFu, er[2 * p2 + 1], 1);
                    const uint64_t a2 = pk2(lane == 0 ? 0.f : sl0, lane == 0 ? 0.f : sl1), b2 = pk2(lane == 31 ? 0.f : sr0, lane == 31 ? 0.f : sr1);
                    e2[p2] = fadd2(pk2(ec[2 * p2], ec[2 * p2 + 1]), fadd2(a2, b2));
                }
                if (lane == 0 && qw > 0) {                         // the two lanes at a warp boundary: ONE divergent block per tile each
                    const uint64_t* xp = reinterpret_cast<const uint64_t*>(xa - NH * HC);     // quarter qw - 1, same half: NH * HC floats back
#pragma unroll
                    for (int p2 = 0; p2 < HC / 2; ++p2) e2[p2] = fadd2(e2[p2], xp[p2]);
                }
                if (lane == 31 && qw < 3) {
                    const uint64_t* xp = reinterpret_cast<const uint64_t*>(xb + NH * HC);
#pragma unroll
                    for (int p2 = 0; p2 < HC / 2; ++p2) e2[p2] = fadd2(e2[p2], xp[p2]);
                }
                if (valid) {
                    const uint64_t* b2 = reinterpret_cast<const uint64_t*>(bsm + half * HC);
#pragma unroll
                    for (int p2 = 0; p2 < HC / 2; ++p2) {          // + bias, ReLU
                        float a, b;
                        upk2(fadd2(e2[p2], b2[p2]), a, b);
                        e2[p2] = pk2(fmaxf(a, 0.f), fmaxf(b, 0.f));
                    }
                    if (RES == 1) {                              // + x (identity): the centre pixel is staged row o + 2 + P
                        const uint32_t srow = (uint32_t)(o + 2 + P);
#pragma unroll
                        for (int g8 = 0; g8 < HC / 8; ++g8) {
                            const uint4 xh = *reinterpret_cast<const uint4*>(strip + sw128_offset(srow, (uint32_t)(half * (HC / 8) + g8)));
                            const uint4 xl = *reinterpret_cast<const uint4*>(strip + sw128_offset(srow, (uint32_t)(2 * KH + half * (HC / 8) + g8)));
                            const uint32_t hw[4] = {xh.x, xh.y, xh.z, xh.w}, lw[4] = {xl.x, xl.y, xl.z, xl.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const float2 a = unpack2<true>(hw[i]), b = unpack2<true>(lw[i]);
                                e2[4 * g8 + i] = fadd2(e2[4 * g8 + i], fadd2(pk2(a.x, a.y), pk2(b.x, b.y)));
                            }
                        }
                    } else {
                        const uint64_t* r2 = reinterpret_cast<const uint64_t*>(bsm + COUT + half * HC);
#pragma unroll
                        for (int p2 = 0; p2 < HC / 2; ++p2) e2[p2] = fadd2(e2[p2], fadd2(pk2(v2[2 * p2], v2[2 * p2 + 1]), r2[p2]));
                    }
                    // re-split: hi = bf16(v), lo = bf16(v - hi) (v - hi as fma(hi, -1, v): exact product, one rounding, = the subtraction)
                    uint32_t hw[HC / 2], lw[HC / 2];
                    const uint64_t m1 = pk2(-1.f, -1.f);
#pragma unroll
                    for (int p2 = 0; p2 < HC / 2; ++p2) {
                        float a, b;
                        upk2(e2[p2], a, b);
                        hw[p2] = pack2<true>(a, b);
                        const float2 f = unpack2<true>(hw[p2]);
                        upk2(ffma2(pk2(f.x, f.y), m1, e2[p2]), a, b);
                        lw[p2] = pack2<true>(a, b);
                    }
                    uint4* dst = reinterpret_cast<uint4*>(y + (((size_t)clip * H + yy) * W + xx) * (2 * COUT));
                    // this thread's hi chunks and its lo chunks are 32 contiguous bytes each
                    st_global_v8(dst + half * 2, make_uint4(hw[0], hw[1], hw[2], hw[3]), make_uint4(hw[4], hw[5], hw[6], hw[7]));
                    st_global_v8(dst + COUT / 8 + half * 2, make_uint4(lw[0], lw[1], lw[2], lw[3]), make_uint4(lw[4], lw[5], lw[6], lw[7]));
                }
                ME_TL(6);
            }
            // this warp's MMAs of the band have completed (their acc_full was observed) and its residual reads of the strip are done;
            // the OTHER set's tiles are covered by that set's own arrivals
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->strip_free[j & 1]));
            ME_TL(7);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMeRowWarps + kMeProducerWarps) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kMeAcc * 128);
    }
}

// ---------------------------------------------------------------- max-pool on split pixels (torch MaxPool2d, -inf padding)
// x [B][H][W][C hi | C lo] -> y [B][Ho][Wo][C hi | C lo].  A thread owns 8 channels of one output COLUMN segment and slides down
// the rows: the maximum over the KW input columns of an input row is computed once and kept in a register ring of KH rows, so an
// output costs SH x KW pixel loads instead of KH x KW.  (Used for the small last pool; the two big ones are row-staged, below.)
template <int C, int KH, int KW, int SH, int SW, int PH, int PW>
__global__ void __launch_bounds__(128) maxpool_split_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y, int H, int W, int Ho, int Wo,
                                                           int seg_rows, int nseg, long n_items) {
    pdl_wait();
    const long i = (long)blockIdx.x * 128 + threadIdx.x;
    if (i >= n_items) return;
    constexpr int G = C / 8;
    const int g = (int)(i % G);
    long t = i / G;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int seg = (int)(t % nseg);
    const long clip = t / nseg;
    const uint16_t* xc = x + (size_t)clip * H * W * (2 * C);
    uint16_t* yc = y + (size_t)clip * Ho * Wo * (2 * C);
    const int ho0 = seg * seg_rows, ho1 = min(ho0 + seg_rows, Ho);
    float ring[KH][8];                                           // row maxima of input rows h, indexed h mod KH (fully unrolled below)
    auto row_max = [&](int h, float* m) {
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
        if (h < 0 || h >= H) return;
        // unconditional loads from a clamped column + a select on use: no branch sits between the 2 KW loads, so all of them are in flight
        // together (with `continue` on the padding columns every load waited for the previous one's branch)
        uint4 vh[KW], vl[KW];
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
            const int w_ = min(max(wo * SW - PW + kw, 0), W - 1);
            const uint4* p = reinterpret_cast<const uint4*>(xc + ((size_t)h * W + w_) * (2 * C));
            vh[kw] = __ldg(p + g), vl[kw] = __ldg(p + G + g);
        }
#pragma unroll
        for (int kw = 0; kw < KW; ++kw) {
            const int w_ = wo * SW - PW + kw;
            const bool ok = w_ >= 0 && w_ < W;
            float v[8];
            me_join8(vh[kw], vl[kw], v);
#pragma unroll
            for (int e = 0; e < 8; ++e) m[e] = ok ? fmaxf(m[e], v[e]) : m[e];
        }
    };
    // rows of the first window except its last SH rows
    int hnext = ho0 * SH - PH;                                   // next input row to fetch
#pragma unroll
    for (int k = 0; k < KH - SH; ++k, ++hnext) row_max(hnext, ring[k]);
    int slot = KH - SH;                                          // ring slot of the next fetched row (compile-time pattern: KH, SH small)
    for (int ho = ho0; ho < ho1; ++ho) {
#pragma unroll
        for (int k = 0; k < SH; ++k, ++hnext) {
            float m[8];
            row_max(hnext, m);
#pragma unroll
            for (int s2 = 0; s2 < KH; ++s2)
                if (s2 == slot) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) ring[s2][e] = m[e];
                }
            slot = slot + 1 == KH ? 0 : slot + 1;
        }
        float o8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            float m = ring[0][e];
#pragma unroll
            for (int k = 1; k < KH; ++k) m = fmaxf(m, ring[k][e]);
            o8[e] = m;
        }
        uint4 hi, lo;
        me_split8(o8, hi, lo);
        uint4* dst = reinterpret_cast<uint4*>(yc + ((size_t)ho * Wo + wo) * (2 * C));
        st_global_na_v4(dst + g, hi), st_global_na_v4(dst + G + g, lo);
    }
}

// Row-staged variant for the two big pools (input rows of 8 KB: 128 px x 16 ch or 64 px x 32 ch).  A block of Wo x C / 8 threads owns
// a segment of output rows of ONE clip and walks down the input rows: every input row is copied ONCE, contiguously, into a 4-row
// shared-memory ring with cp.async (one L1 tag per 128-byte line; the per-thread loads above touch every line ~10 times, and the L1
// tag stage was the bound: l1tex 83 %), three rows ahead of the row being reduced.  16-byte chunk c of a row lives in line c / 8 at
// position (c % 8) ^ f(line), f chosen so that the 8 lanes of a quarter-warp (C / 8 channel groups x neighbouring output columns,
// i.e. stride-2 pixels) read 8 different bank groups.  Row maxima go through the same register ring as above.
template <int C, int KH, int KW, int SH, int SW, int PH, int PW, int W>
__global__ void __launch_bounds__(((W + 2 * PW - KW) / SW + 1) * (C / 8)) maxpool_staged_kernel(const uint16_t* __restrict__ x, uint16_t* __restrict__ y,
                                                                                             int H, int Ho, int seg_rows, int nseg) {
    constexpr int G = C / 8, Wo = (W + 2 * PW - KW) / SW + 1, NT = Wo * G;
    constexpr int CH = W * 2 * G, CPT = CH / NT, RING = 4;          // 16-byte chunks of an input row; per thread; rows in flight + 1
    static_assert(CH % NT == 0 && CH * 16 == 8192, "one input row = 8 KB");
    __shared__ __align__(128) uint4 ring_s[RING][CH];
    pdl_wait();
    const int tid = threadIdx.x, g = tid % G, wo = tid / G;
    const int seg = blockIdx.x % nseg;
    const long clip = blockIdx.x / nseg;
    const uint16_t* xc = x + (size_t)clip * H * W * (2 * C);
    uint16_t* yc = y + (size_t)clip * Ho * Wo * (2 * C);
    const int ho0 = seg * seg_rows, ho1 = min(ho0 + seg_rows, Ho);
    auto swz = [](int c) {                                          // chunk index of the row -> swizzled chunk index
        const int line = c >> 3;
        return (c & ~7) | ((c & 7) ^ (C == 16 ? 2 * (line & 3) : 4 * ((line >> 1) & 1)));
    };
    auto stage = [&](int h) {                                       // (rows outside the image are not staged: their maxima are -inf)
        if (h >= 0 && h < H) {
            const uint4* src = reinterpret_cast<const uint4*>(xc + (size_t)h * W * (2 * C));
#pragma unroll
            for (int k = 0; k < CPT; ++k) cp_async16(&ring_s[h & (RING - 1)][swz(tid + k * NT)], src + tid + k * NT);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    const int hfirst = ho0 * SH - PH, hlast = (ho1 - 1) * SH - PH + KH - 1;
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) stage(hfirst + k);
    float rmax[KH][8];                                               // row maxima of the last KH input rows (slot pattern unrolled below)
#pragma unroll
    for (int k = 0; k < KH; ++k)
#pragma unroll
        for (int e = 0; e < 8; ++e) rmax[k][e] = -INFINITY;
    int slot = 0, ho = ho0, hneed = ho0 * SH - PH + KH - 1;          // next output row and the input row that completes its window
    for (int h = hfirst; h <= hlast; ++h) {
        asm volatile("cp.async.wait_group %0;" ::"n"(RING - 2) : "memory");
        __syncthreads();                                            // row h has landed for every thread; row h - 1 is no longer read
        stage(h + RING - 1);                                        // into the slot of row h - 1
        float m[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
        if (h >= 0 && h < H) {
            const uint4* row = ring_s[h & (RING - 1)];
#pragma unroll
            for (int kw = 0; kw < KW; ++kw) {
                const int w0 = wo * SW - PW + kw, w_ = min(max(w0, 0), W - 1);
                const bool ok = w0 >= 0 && w0 < W;
                float v[8];
                me_join8(row[swz(w_ * 2 * G + g)], row[swz(w_ * 2 * G + G + g)], v);
#pragma unroll
                for (int e = 0; e < 8; ++e) m[e] = ok ? fmaxf(m[e], v[e]) : m[e];
            }
        }
#pragma unroll
        for (int s2 = 0; s2 < KH; ++s2)
            if (s2 == slot) {
#pragma unroll
                for (int e = 0; e < 8; ++e) rmax[s2][e] = m[e];
            }
        slot = slot + 1 == KH ? 0 : slot + 1;
        if (h == hneed) {
            float o8[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float t = rmax[0][e];
#pragma unroll
                for (int k = 1; k < KH; ++k) t = fmaxf(t, rmax[k][e]);
                o8[e] = t;
            }
            uint4 hi, lo;
            me_split8(o8, hi, lo);
            uint4* dst = reinterpret_cast<uint4*>(yc + ((size_t)ho * Wo + wo) * (2 * C));
            st_global_na_v4(dst + g, hi), st_global_na_v4(dst + G + g, lo);
            ++ho, hneed += SH;
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// h3 [B][T][16 bins][32 hi | 32 lo] -> flatten (feature = channel * 16 + bin, transformer.py:337) -> conv4 (512 -> 64, folded
// BatchNorm1d) -> xf_out [B][T][64];  xf_proj = proj(xf_out) (transformer.py:458).  w4t [512][64], wpt [64][64] (input-major).
constexpr int kC4Rows = 16;
__global__ void __launch_bounds__(256) conv4_proj_split_kernel(const uint16_t* __restrict__ h3, const float* __restrict__ w4t, const float* __restrict__ b4,
                                                              const float* __restrict__ wpt, const float* __restrict__ bp, float* __restrict__ xf_out,
                                                              float* __restrict__ xf_proj, long M) {
    __shared__ __align__(16) float s_f[kC4Rows][512 + 4];
    __shared__ __align__(16) float s_o[kC4Rows][64];
    pdl_wait();
    const long row0 = (long)blockIdx.x * kC4Rows;
    const int tid = threadIdx.x;
    for (int i = tid; i < kC4Rows * 16 * 4; i += 256) {         // (row, bin, 8-channel group)
        const int r = i >> 6, w = (i >> 2) & 15, g = i & 3;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
        if (row0 + r < M) {
            const uint4* p = reinterpret_cast<const uint4*>(h3 + ((size_t)(row0 + r) * 16 + w) * 64);
            me_join8(__ldg(p + g), __ldg(p + 4 + g), v);
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) s_f[r][(8 * g + e) * 16 + w] = v[e];
    }
    __syncthreads();
    const int o = tid & 63, rq = tid >> 6;                     // output feature, row quarter (4 rows each)
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = __ldg(b4 + o);
#pragma unroll 4
    for (int k = 0; k < 512; k += 4) {                         // four features per 16-byte shared-memory load; k ascending per accumulator
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(w4t + (k + u) * 64 + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 f = *reinterpret_cast<const float4*>(&s_f[4 * rq + j][k]);
            acc[j] = fmaf(f.x, wv[0], acc[j]), acc[j] = fmaf(f.y, wv[1], acc[j]), acc[j] = fmaf(f.z, wv[2], acc[j]), acc[j] = fmaf(f.w, wv[3], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_o[4 * rq + j][o] = acc[j];
        if (row0 + 4 * rq + j < M) xf_out[(row0 + 4 * rq + j) * 64 + o] = acc[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = __ldg(bp + o);
    for (int k = 0; k < 64; k += 4) {
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = __ldg(wpt + (k + u) * 64 + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 f = *reinterpret_cast<const float4*>(&s_o[4 * rq + j][k]);
            acc[j] = fmaf(f.x, wv[0], acc[j]), acc[j] = fmaf(f.y, wv[1], acc[j]), acc[j] = fmaf(f.z, wv[2], acc[j]), acc[j] = fmaf(f.w, wv[3], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (row0 + 4 * rq + j < M) xf_proj[(row0 + 4 * rq + j) * 64 + o] = acc[j];
}

}  // namespace dc
