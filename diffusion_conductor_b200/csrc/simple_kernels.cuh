// CUDA-core kernels of the denoise step: weight packing, timestep embedding, per-step
// conditioning (SiLU(te + xp) as a packed A-operand image), joint embedding, the time-axis
// softmax + K^T V reduction, the output head fused with the x0-parameterised DDIM / DDPM update.
// All of them are HBM/L2-bound elementwise or reduction work (SURVEY.md §8(d) "HBM side").
#pragma once
#include "tc_common.cuh"

namespace dc {

constexpr int kD = 128;     // latent_dim
constexpr int kE = 512;     // time_embed_dim == music latent dim (reference transformer.py:385,405)
constexpr int kH = 8;       // heads
constexpr int kHd = 16;     // head dim
constexpr int kF = 64;      // ff_size
constexpr int kP = 26;      // input_feats (13 joints x 2)
constexpr int kMusic = 64;  // music-encoder feature width
constexpr float kLnEps = 1e-5f;

// "Blocked" activation layout used for h [*,128] and k|v [*,256]: per 128-token tile, 16-byte column chunks are
// the slow index and rows the fast one -- [tile][ncols/4][128 rows][4 floats] -- so that a warp whose lanes are
// 32 consecutive rows reads / writes 512 contiguous bytes per 128-bit instruction.
__device__ __host__ __forceinline__ size_t blk_index(long g, int col, int ncols) {
    return (size_t)(g >> 7) * (size_t)(128 * ncols) + (size_t)(col >> 2) * 512 + (size_t)(g & 127) * 4 + (col & 3);
}

// ---------------------------------------------------------------------------------------------
// Weight packing: fp32 [N_src, K_src] (torch Linear layout) -> 16-bit K-major SW128 image
//   dst[kb][n][...] with kb = k / 64, one block = Nrows x 128 bytes.
// rowmap (optional) permutes / pads output rows; colscale (optional) folds a LayerNorm gamma.
// ---------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void pack_weight_kernel(const float* __restrict__ W, int ldw, int Ksrc, const int* __restrict__ rowmap,
                                   const float* __restrict__ colscale, int Nrows, int Kblocks, uint16_t* __restrict__ dst) {
    const int total = Nrows * Kblocks * 64;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx % (Kblocks * 64);
        const int n = idx / (Kblocks * 64);
        const int src = rowmap ? rowmap[n] : n;
        float v = 0.f;
        if (src >= 0 && k < Ksrc) {
            v = W[(size_t)src * ldw + k];
            if (colscale) v *= colscale[k];
        }
        const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
        const size_t off = (size_t)kb * Nrows * 128 + sw128_offset(n, c) + e * 2;
        dst[off >> 1] = pack1<kBf16>(v);
    }
}

// b'[n] = b[n] + sum_k W[n][k] * beta[k]   (LayerNorm beta folded through the following Linear)
__global__ void fold_bias_kernel(const float* __restrict__ W, int K, const float* __restrict__ beta,
                                 const float* __restrict__ b, int N, float* __restrict__ out) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) acc = fmaf(W[(size_t)n * K + k], beta[k], acc);
    out[n] = b[n] + acc;
}

// out[c][r] = in[r][c]
__global__ void transpose_kernel(const float* __restrict__ in, int R, int C, float* __restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * C) return;
    const int r = idx / C, c = idx % C;
    out[(size_t)c * R + r] = in[idx];
}

// ---------------------------------------------------------------------------------------------
// Timestep embedding + time MLP (reference transformer.py:8-25, 410-414, 482):
//   te[n] = W2 . SiLU(W0 . [cos(t f) | sin(t f)] + b0) + b2 ,  f built on the host exactly as the
//   reference builds it (fp32 exp on CPU).  t == nullptr -> t = first_t + blockIdx.x (schedule table).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) time_embed_kernel(const long long* __restrict__ t, int first_t,
                                                          const float* __restrict__ freqs, const float* __restrict__ W0,
                                                          const float* __restrict__ b0, const float* __restrict__ W2,
                                                          const float* __restrict__ b2, float* __restrict__ out) {
    __shared__ float emb[kD];
    __shared__ float hid[kE];
    const int j = threadIdx.x;
    const float tv = static_cast<float>(t ? t[blockIdx.x] : (long long)(first_t + blockIdx.x));
    if (j < kD / 2) {
        const float arg = __fmul_rn(tv, freqs[j]);
        emb[j] = cosf(arg);
        emb[j + kD / 2] = sinf(arg);
    }
    __syncthreads();
    float acc = b0[j];
    const float* w = W0 + (size_t)j * kD;
#pragma unroll 8
    for (int i = 0; i < kD; ++i) acc = fmaf(emb[i], w[i], acc);
    hid[j] = acc / (1.f + expf(-acc));
    __syncthreads();
    acc = b2[j];
    w = W2 + (size_t)j * kE;
#pragma unroll 8
    for (int i = 0; i < kE; ++i) acc = fmaf(hid[i], w[i], acc);
    out[(size_t)blockIdx.x * kE + j] = acc;
}

// ---------------------------------------------------------------------------------------------
// Conditioning, once per batch of clips (step-invariant; reference recomputes it every step at
// transformer.py:479-480): xp = linear(xf_proj) kept fp32; xf = linear(xf_out) is LayerNorm-
// normalised (text_norm without its affine, which is folded into each layer's K/V weights) and
// written as a packed 16-bit A-operand image Z[tile][kb][128 x 64].
// One block = 4 tokens, thread j owns output features j, j+128, j+256, j+384.
// ---------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(128) cond_prep_kernel(const float* __restrict__ xf_proj, const float* __restrict__ xf_out,
                                                         const float* __restrict__ WlinT /*[64][512]*/,
                                                         const float* __restrict__ blin, int M, float* __restrict__ xp,
                                                         uint8_t* __restrict__ zimg) {
    constexpr int TOK = 4;
    __shared__ float in_p[TOK][kMusic], in_o[TOK][kMusic];
    __shared__ float red[TOK][2][4];
    const int j = threadIdx.x;
    const long g0 = (long)blockIdx.x * TOK;
    for (int i = j; i < TOK * kMusic; i += 128) {
        const int tk = i / kMusic, c = i % kMusic;
        const long g = g0 + tk;
        in_p[tk][c] = g < M ? xf_proj[g * kMusic + c] : 0.f;
        in_o[tk][c] = g < M ? xf_out[g * kMusic + c] : 0.f;
    }
    __syncthreads();
    float ap[TOK][4], ao[TOK][4];
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk)
#pragma unroll
        for (int q = 0; q < 4; ++q) ap[tk][q] = ao[tk][q] = blin[j + 128 * q];
    for (int c = 0; c < kMusic; ++c) {
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = WlinT[c * kE + j + 128 * q];
#pragma unroll
        for (int tk = 0; tk < TOK; ++tk)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                ap[tk][q] = fmaf(in_p[tk][c], w[q], ap[tk][q]);
                ao[tk][q] = fmaf(in_o[tk][c], w[q], ao[tk][q]);
            }
    }
    // LayerNorm statistics of xf rows (two-pass: mean, then centred sum of squares)
    const int lane = j & 31, wid = j >> 5;
    float mean[TOK], rstd[TOK];
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int tk = 0; tk < TOK; ++tk) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float d = pass ? ao[tk][q] - mean[tk] : ao[tk][q];
                s += pass ? d * d : d;
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) red[tk][pass][wid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int tk = 0; tk < TOK; ++tk) {
            const float s = red[tk][pass][0] + red[tk][pass][1] + red[tk][pass][2] + red[tk][pass][3];
            if (pass == 0) mean[tk] = s * (1.f / kE);
            else rstd[tk] = rsqrtf(s * (1.f / kE) + kLnEps);
        }
    }
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk) {
        const long g = g0 + tk;
        if (g >= M) continue;
        const long tile = g >> 7;
        const uint32_t r = (uint32_t)(g & 127);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int e = j + 128 * q;
            xp[g * kE + e] = ap[tk][q];
            const int kb = e >> 6, c = (e & 63) >> 3;
            const size_t off = ((size_t)tile * 8 + kb) * kABlockBytes + sw128_offset(r, c) + (e & 7) * 2;
            reinterpret_cast<uint16_t*>(zimg)[off >> 1] = pack1<kBf16>((ao[tk][q] - mean[tk]) * rstd[tk]);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Start of a denoise step (reference transformer.py:482, 488-490):
//   A_emb = SiLU(te[b] + xp[b,t])  -> packed 16-bit image  (the operand of all 24 FiLM projections;
//           the reference re-evaluates this SiLU 24 times per step)
//   h0    = joint_embed(x) + sequence_embedding[t]
// One block = 8 tokens x 128 threads.  te row: te + te_row0*512 + b*te_stride (stride 0 = one
// timestep shared by the batch, read from the device-side step counter).
// ---------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(128) step_begin_kernel(const float* __restrict__ x, const float* __restrict__ xp,
                                                          const float* __restrict__ te, const int* __restrict__ step_ctr,
                                                          int te_stride, const float* __restrict__ WjT /*[26][128]*/,
                                                          const float* __restrict__ bj, const float* __restrict__ pos,
                                                          int M, int T, uint8_t* __restrict__ aemb, float* __restrict__ h) {
    constexpr int TOK = 8;
    __shared__ float xs[TOK][kP + 2];
    const int j = threadIdx.x;
    const long g0 = (long)blockIdx.x * TOK;
    pdl_trigger();
    pdl_wait();
    const float* te0 = te + (step_ctr ? (size_t)(*step_ctr) * kE : 0);
    for (int i = j; i < TOK * kP; i += 128) {
        const int tk = i / kP, c = i % kP;
        const long g = g0 + tk;
        xs[tk][c] = g < M ? x[g * kP + c] : 0.f;
    }
    // ---- A_emb: 8 tokens x 64 chunks of 8 features
#pragma unroll
    for (int it = 0; it < TOK * 64 / 128; ++it) {
        const int task = it * 128 + j;
        const int tk = task >> 6, ch = task & 63;
        const long g = g0 + tk;
        if (g < M) {
            const int b = (int)(g / T);
            const float4* xr = reinterpret_cast<const float4*>(xp + g * kE + ch * 8);
            const float4* tr = reinterpret_cast<const float4*>(te0 + (size_t)b * te_stride + ch * 8);
            const float4 a0 = __ldg(xr), a1 = __ldg(xr + 1), t0 = __ldg(tr), t1 = __ldg(tr + 1);
            float v[8] = {a0.x + t0.x, a0.y + t0.y, a0.z + t0.z, a0.w + t0.w, a1.x + t1.x, a1.y + t1.y, a1.z + t1.z, a1.w + t1.w};
            uint32_t p[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float s0 = v[2 * i] / (1.f + __expf(-v[2 * i]));
                const float s1 = v[2 * i + 1] / (1.f + __expf(-v[2 * i + 1]));
                p[i] = pack2<kBf16>(s0, s1);
            }
            const long tile = g >> 7;
            const uint32_t r = (uint32_t)(g & 127);
            const int kb = ch >> 3, c = ch & 7;
            uint4* dst = reinterpret_cast<uint4*>(aemb + ((size_t)tile * 8 + kb) * kABlockBytes + sw128_offset(r, c));
            *dst = make_uint4(p[0], p[1], p[2], p[3]);
        }
    }
    __syncthreads();
    // ---- h0
    float acc[TOK];
    const float bjv = bj[j];
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk) acc[tk] = bjv;
    for (int c = 0; c < kP; ++c) {
        const float w = WjT[c * kD + j];
#pragma unroll
        for (int tk = 0; tk < TOK; ++tk) acc[tk] = fmaf(xs[tk][c], w, acc[tk]);
    }
#pragma unroll
    for (int tk = 0; tk < TOK; ++tk) {
        const long g = g0 + tk;
        if (g < M) h[blk_index(g, j, kD)] = acc[tk] + pos[(size_t)(g % T) * kD + j];
    }
}

// ---------------------------------------------------------------------------------------------
// Time-axis softmax and K^T V (reference transformer.py:111,117 / 151,155):
//   A[b,h,d,l] = sum_t softmax_t(k[b,t,h,d]) * v[b,t,h,l]
// kv is fp32 [*,256] in the blocked layout with k in columns [0,128) and v in [128,256).  One block per (clip, head),
// 256 threads = (d,l) pairs.  Two passes over the clip's 16 key columns: max, then exp/sum/outer
// product through shared memory.  All fp32.  The result is written as the 16x16 diagonal block of
// head h in a packed 16-bit B-operand image Bd[n = 16h+l][k = 16h+d] (off-diagonal blocks stay zero),
// so that y = q . blockdiag(A) runs on the tensor cores in the layer kernel.
// ---------------------------------------------------------------------------------------------
template <bool kBf16>
__global__ void __launch_bounds__(256) kv_reduce_kernel(const float* __restrict__ kv /*blocked [*,256]*/, int T,
                                                         uint8_t* __restrict__ bd /*[B] images of 32 KB*/, size_t bd_stride,
                                                         size_t kv_layer_stride = 0 /* floats; blockIdx.y = layer */, size_t bd_layer_stride = 0 /* bytes */,
                                                         int compact = 0 /* 1: 4 KB head-block image (bdc_offset) instead of the [128 x 128] one */) {
    kv += (size_t)blockIdx.y * kv_layer_stride;
    bd += (size_t)blockIdx.y * bd_layer_stride;
    // loads: thread = (column chunk q of the head's 16 key/value columns, token lane) -> float4, 32 consecutive
    //        tokens per warp = 512 contiguous bytes of the blocked layout.
    // accumulation: 4 token groups x 64 threads, each thread a 2x2 block of the 16x16 result.
    constexpr int TT = 128;
    __shared__ __align__(16) float ek[TT][kHd];
    __shared__ __align__(16) float vv[TT][kHd];
    __shared__ float red[8][4];
    __shared__ float cmax[kHd];
    __shared__ float part[4][kHd * kHd + kHd];
    const int b = blockIdx.x / kH, hh = blockIdx.x % kH;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int q = tid >> 6, tt = tid & 63;
    const long g0 = (long)b * T;
    const int kcol = hh * kHd + 4 * q;
    pdl_trigger();
    pdl_wait();
    // pass 1: column max over the clip
    float4 m4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    for (int t0 = 0; t0 < T; t0 += 256) {
        float4 x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + tt + 64 * i;
            x[i] = t < T ? __ldg(reinterpret_cast<const float4*>(kv + blk_index(g0 + t, kcol, 256))) : m4;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            m4.x = fmaxf(m4.x, x[i].x), m4.y = fmaxf(m4.y, x[i].y), m4.z = fmaxf(m4.z, x[i].z), m4.w = fmaxf(m4.w, x[i].w);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        m4.x = fmaxf(m4.x, __shfl_xor_sync(0xffffffffu, m4.x, o));
        m4.y = fmaxf(m4.y, __shfl_xor_sync(0xffffffffu, m4.y, o));
        m4.z = fmaxf(m4.z, __shfl_xor_sync(0xffffffffu, m4.z, o));
        m4.w = fmaxf(m4.w, __shfl_xor_sync(0xffffffffu, m4.w, o));
    }
    if (lane == 0) red[wid][0] = m4.x, red[wid][1] = m4.y, red[wid][2] = m4.z, red[wid][3] = m4.w;
    __syncthreads();
    if (tid < kHd) cmax[tid] = fmaxf(red[2 * (tid >> 2)][tid & 3], red[2 * (tid >> 2) + 1][tid & 3]);
    __syncthreads();
    const float4 mc = *reinterpret_cast<const float4*>(&cmax[4 * q]);
    // pass 2: exp, column sums and the 16x16 outer-product accumulation through shared memory
    const int grp = tid >> 6, sub = tid & 63;
    const int d0 = (sub >> 3) * 2, l0 = (sub & 7) * 2;
    float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f, s0 = 0.f, s1 = 0.f;
    for (int t0 = 0; t0 < T; t0 += TT) {
        float4 xk[2], xv[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int t = t0 + tt + 64 * i;
            const bool ok = t < T;
            const float* p = kv + blk_index(g0 + (ok ? t : 0), kcol, 256);
            xk[i] = ok ? __ldg(reinterpret_cast<const float4*>(p)) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            xv[i] = ok ? __ldg(reinterpret_cast<const float4*>(p + 32 * 512)) : make_float4(0.f, 0.f, 0.f, 0.f);   // v = column + 128
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {                            // exp(-inf) = 0 for the padded tail
            *reinterpret_cast<float4*>(&ek[tt + 64 * i][4 * q]) =
                make_float4(expf(xk[i].x - mc.x), expf(xk[i].y - mc.y), expf(xk[i].z - mc.z), expf(xk[i].w - mc.w));
            *reinterpret_cast<float4*>(&vv[tt + 64 * i][4 * q]) = xv[i];
        }
        __syncthreads();
        const int n = min(TT, T - t0);
#pragma unroll 4
        for (int k = grp; k < n; k += 4) {
            const float2 e = *reinterpret_cast<const float2*>(&ek[k][d0]);
            const float2 v = *reinterpret_cast<const float2*>(&vv[k][l0]);
            a00 = fmaf(e.x, v.x, a00), a01 = fmaf(e.x, v.y, a01);
            a10 = fmaf(e.y, v.x, a10), a11 = fmaf(e.y, v.y, a11);
            s0 += e.x, s1 += e.y;
        }
        __syncthreads();
    }
    float* pp = part[grp];
    pp[d0 * 16 + l0] = a00, pp[d0 * 16 + l0 + 1] = a01, pp[(d0 + 1) * 16 + l0] = a10, pp[(d0 + 1) * 16 + l0 + 1] = a11;
    if (l0 == 0) pp[256 + d0] = s0, pp[256 + d0 + 1] = s1;
    __syncthreads();
    const int d = tid >> 4, l = tid & 15;
    const float acc = (part[0][tid] + part[1][tid]) + (part[2][tid] + part[3][tid]);
    const float se = (part[0][256 + d] + part[1][256 + d]) + (part[2][256 + d] + part[3][256 + d]);
    const int ki = hh * kHd + d, nj = hh * kHd + l;
    const size_t off = (size_t)b * bd_stride + (compact ? (size_t)bdc_offset((uint32_t)hh, (uint32_t)d, (uint32_t)l)
                                                        : (size_t)(ki >> 6) * kABlockBytes + sw128_offset(nj, (ki & 63) >> 3) + (ki & 7) * 2);
    *reinterpret_cast<uint16_t*>(bd + off) = pack1<kBf16>(acc / se);
}

// ---------------------------------------------------------------------------------------------
// Output head + sampler update (reference transformer.py:496; gaussian_diffusion.py:812-830 DDIM,
// :426-429,:656-664 DDPM).  coef row (8 floats, built on the host from the fp64 tables exactly as
// the reference gathers them):
//   [0] sqrt_recip_alphas_cumprod  [1] sqrt_recipm1_alphas_cumprod  [2] sqrt(alpha_bar_prev)
//   [3] sqrt(1 - alpha_bar_prev - sigma^2)  [4] (t!=0)*sigma
//   [5] posterior_mean_coef1  [6] posterior_mean_coef2  [7] (t!=0)*exp(0.5*posterior_log_variance_clipped)
// The update uses explicitly rounded mul/add/div (no FMA contraction) so that, given the same x0, it
// is bit-identical to the reference's chain of separate torch ops.
// mode 0: model output only; 1: DDIM; 2: DDPM; | 0x10: clamp pred_xstart to [-1, 1] first.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ddim_rule(float x, float x0, const float* cf, float nz) {
    const float eps = __fdiv_rn(__fsub_rn(__fmul_rn(cf[0], x), x0), cf[1]);
    const float mean = __fadd_rn(__fmul_rn(x0, cf[2]), __fmul_rn(cf[3], eps));
    return __fadd_rn(mean, __fmul_rn(cf[4], nz));
}
__device__ __forceinline__ float ddpm_rule(float x, float x0, const float* cf, float nz) {
    const float mean = __fadd_rn(__fmul_rn(cf[5], x0), __fmul_rn(cf[6], x));
    return __fadd_rn(mean, __fmul_rn(cf[7], nz));
}

__global__ void __launch_bounds__(256) out_update_kernel(const float* __restrict__ h, const float* __restrict__ WoT /*[128][32]*/,
                                                          const float* __restrict__ bo, int M, int mode,
                                                          const float* __restrict__ coef, const int* __restrict__ step_ctr,
                                                          const float* __restrict__ noise, float* __restrict__ x,
                                                          float* __restrict__ x0_out) {
    constexpr int TOK = 8;
    __shared__ float hs[TOK][kD];
    const long g0 = (long)blockIdx.x * TOK;
    pdl_trigger();
    pdl_wait();
    for (int i = threadIdx.x; i < TOK * kD; i += 256) {
        const long g = g0 + i / kD;
        hs[i / kD][i % kD] = g < M ? h[blk_index(g, i % kD, kD)] : 0.f;
    }
    __syncthreads();
    const int tk = threadIdx.x >> 5, p = threadIdx.x & 31;
    const long g = g0 + tk;
    if (p >= kP || g >= M) return;
    float acc = bo[p];
#pragma unroll 16
    for (int j = 0; j < kD; ++j) acc = fmaf(hs[tk][j], WoT[j * 32 + p], acc);
    const size_t idx = (size_t)g * kP + p;
    if (mode & 0x10) acc = fminf(fmaxf(acc, -1.f), 1.f);          // clip_denoised (gaussian_diffusion.py:506-507)
    x0_out[idx] = acc;
    if ((mode & 0xF) == 0) return;
    const float* cf = coef + (size_t)(*step_ctr) * 8;
    const float nz = noise ? noise[idx] : 0.f;
    x[idx] = (mode & 0xF) == 1 ? ddim_rule(x[idx], acc, cf, nz) : ddpm_rule(x[idx], acc, cf, nz);
}

// Stand-alone sampler update on a caller-supplied x0 (bit-exactness tests; also the ddim_sample /
// p_sample entry when the caller wants the reference's two-output dict).
__global__ void sampler_update_kernel(const float* __restrict__ x0, size_t n, int mode, const float* __restrict__ coef, int step,
                                      const float* __restrict__ noise, float* __restrict__ x) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* cf = coef + (size_t)step * 8;
    const float nz = noise ? noise[i] : 0.f;
    x[i] = (mode & 0xF) == 1 ? ddim_rule(x[i], x0[i], cf, nz) : ddpm_rule(x[i], x0[i], cf, nz);
}

// Post-processing of generated motion (reference Diffusion_Stage/tools/visualization.py:20-26,107-126):
// keypoints * window (pixels), then scipy.signal.savgol_filter(data, kernel, order) along time for each of the
// C coordinates (mode='interp': the first / last kernel/2 frames are evaluated from the polynomial fitted to the
// first / last `kernel` frames).  Both are linear maps with host-computed coefficients: an interior FIR and a
// [kernel/2][kernel] edge matrix (the tail uses it time-reversed).
constexpr int kSavgolMaxWindow = 31;
struct SavgolCoef {
    int window;                                            // odd, <= kSavgolMaxWindow
    float scale;                                           // applied to the input (reference: motions[i] *= 600)
    float fir[kSavgolMaxWindow];
    float edge[kSavgolMaxWindow / 2][kSavgolMaxWindow];
};

__global__ void savgol_kernel(const float* __restrict__ x, float* __restrict__ y, int B, int T, int C, const __grid_constant__ SavgolCoef cf) {
    const long n = (long)B * T * C;
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    const int c = (int)(idx % C);
    const int t = (int)((idx / C) % T);
    const long base = (idx / ((long)C * T)) * (long)T * C + c;      // element (b, 0, c)
    const int half = cf.window / 2;
    float acc = 0.f;
    if (t < half) {
        for (int j = 0; j < cf.window; ++j) acc = fmaf(cf.edge[t][j], x[base + (long)j * C] * cf.scale, acc);
    } else if (t >= T - half) {
        const int i = T - 1 - t;
        for (int j = 0; j < cf.window; ++j) acc = fmaf(cf.edge[i][j], x[base + (long)(T - 1 - j) * C] * cf.scale, acc);
    } else {
        for (int j = 0; j < cf.window; ++j) acc = fmaf(cf.fir[j], x[base + (long)(t - half + j) * C] * cf.scale, acc);
    }
    y[idx] = acc;
}

__global__ void set_step_kernel(int* ctr, int value, int delta) {
    pdl_trigger();
    pdl_wait();
    *ctr = delta ? *ctr + delta : value;
}

}  // namespace dc
