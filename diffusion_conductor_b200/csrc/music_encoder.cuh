// M2SNet music encoder front-end: mel (B, 3T, 128) -> xf_out (B, T, 64), xf_proj = proj(xf_out).
// Reference: Diffusion_Stage/models/transformer.py:289-340 (Conv2dResLayer, MusicEncoder.forward) and :447-459
// (encode_music, eval mode).  Runs once per clip, before the sampling loop.
//
// Seven reflect-padded 3x3 convolutions over (time x mel-bin) images with 1..32 channels, eval-mode BatchNorm folded into
// the weights on the host, ReLU, identity / 1x1-conv residuals, three max-pools, then a 512 -> 64 pointwise convolution
// (+ folded BatchNorm1d) and the 64 -> 64 `proj` Linear.  Channel counts this small do not fill a tensor-core tile and
// the features condition every denoise step, so this stays exact fp32 on the CUDA cores: direct convolution, input tile +
// tile staged in shared memory, weights in the constant bank (kernel parameter), every thread accumulates 4 pixels x 16 output
// channels in registers.  Activations are NCHW fp32; 0.94 GMAC per 6 s clip.
#pragma once
#include <cuda_runtime.h>

namespace dc {

__device__ __forceinline__ int reflect_idx(int i, int n) {      // torch padding_mode='reflect' (pad 1), clamped for off-image tile rows
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

constexpr int kCvTH = 16, kCvTW = 32, kCvChunk = 8;             // output tile (time x bins), input-channel chunk in shared memory
constexpr int kCvCo = 16;                                       // output channels per CTA

// Folded weights of one (layer, group of 16 output channels), passed BY VALUE as a kernel parameter (<= 32 KB): they live in
// the constant bank, every weight is warp-uniform, so the compiler feeds the FFMAs from uniform registers (ULDC) and the
// load/store unit only serves the input pixels -- with the weights in shared memory the kernel was LSU-bound at 28 % of the
// fp32 peak.
template <int CIN, bool kRes>
struct alignas(16) ConvWeights {
    float w[CIN * 9 * kCvCo];                    // [cin][tap][co]; read as float4 (one 128-bit uniform load per 4 weights)
    float b[kCvCo];
    float w1[kRes ? CIN * kCvCo : 1];            // 1x1 residual convolution [cin][co] (RES == 2 only)
    float b1[kRes ? kCvCo : 1];
};

// y = ReLU(conv3x3_reflect(x) * s + b') (+ x | + conv1x1(x) * s1 + b1')     RES: 0 none, 1 identity, 2 1x1 convolution
// x [B][CIN][H][W], y [B][COUT][H][W]; one launch per group `grp` of 16 output channels; blockIdx.z = clip.
template <int CIN, int COUT, int RES>
__global__ void __launch_bounds__(256) conv3x3_bn_relu_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int grp,
                                                              const __grid_constant__ ConvWeights<CIN, RES == 2> cw) {
    constexpr int CC = CIN < kCvChunk ? CIN : kCvChunk;
    __shared__ float s_in[CC][kCvTH + 2][kCvTW + 2];
    const int bi = blockIdx.z, h0 = blockIdx.y * kCvTH, w0 = blockIdx.x * kCvTW;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;     // rows ty, ty + 8, ty + 16, ty + 24 of the tile
    const float* xb = x + (size_t)bi * CIN * H * W;
    constexpr int NP = kCvTH / 8;                                  // pixels per thread: every weight (a uniform-register operand) feeds NP FFMAs
    float acc[NP][kCvCo];
#pragma unroll
    for (int p = 0; p < NP; ++p)
#pragma unroll
        for (int c = 0; c < kCvCo; ++c) acc[p][c] = 0.f;

    {
        for (int c0 = 0; c0 < CIN; c0 += CC) {
            __syncthreads();
            for (int i = tid; i < CC * (kCvTH + 2) * (kCvTW + 2); i += 256) {
                const int c = i / ((kCvTH + 2) * (kCvTW + 2)), r = (i / (kCvTW + 2)) % (kCvTH + 2), q = i % (kCvTW + 2);
                s_in[c][r][q] = __ldg(xb + ((size_t)(c0 + c) * H + reflect_idx(h0 - 1 + r, H)) * W + reflect_idx(w0 - 1 + q, W));
            }
            __syncthreads();
#pragma unroll
            for (int c = 0; c < CC; ++c) {
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        float vin[NP];
#pragma unroll
                        for (int p = 0; p < NP; ++p) vin[p] = s_in[c][ty + 8 * p + dy][tx + dx];
                        const float4* wp = reinterpret_cast<const float4*>(cw.w + ((c0 + c) * 9 + dy * 3 + dx) * kCvCo);
#pragma unroll
                        for (int q4 = 0; q4 < kCvCo / 4; ++q4) {
                            const float4 w4 = wp[q4];
#pragma unroll
                            for (int p = 0; p < NP; ++p) {
                                acc[p][4 * q4] = fmaf(vin[p], w4.x, acc[p][4 * q4]), acc[p][4 * q4 + 1] = fmaf(vin[p], w4.y, acc[p][4 * q4 + 1]);
                                acc[p][4 * q4 + 2] = fmaf(vin[p], w4.z, acc[p][4 * q4 + 2]), acc[p][4 * q4 + 3] = fmaf(vin[p], w4.w, acc[p][4 * q4 + 3]);
                            }
                        }
                    }
                }
            }
        }
        const int gw = w0 + tx;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int gh = h0 + ty + 8 * p;
            if (gh >= H || gw >= W) continue;
            const size_t pix = (size_t)gh * W + gw;
#pragma unroll
            for (int c = 0; c < kCvCo; ++c) acc[p][c] = fmaxf(acc[p][c] + cw.b[c], 0.f);
            if constexpr (RES == 1) {
#pragma unroll
                for (int c = 0; c < kCvCo; ++c) acc[p][c] += __ldg(xb + (size_t)(grp * kCvCo + c) * H * W + pix);
            } else if constexpr (RES == 2) {
#pragma unroll
                for (int c = 0; c < kCvCo; ++c) acc[p][c] += cw.b1[c];
#pragma unroll
                for (int ci = 0; ci < CIN; ++ci) {
                    const float xv = __ldg(xb + (size_t)ci * H * W + pix);
#pragma unroll
                    for (int c = 0; c < kCvCo; ++c) acc[p][c] = fmaf(xv, cw.w1[ci * kCvCo + c], acc[p][c]);
                }
            }
            float* yb = y + ((size_t)bi * COUT + grp * kCvCo) * H * W + pix;
#pragma unroll
            for (int c = 0; c < kCvCo; ++c) yb[(size_t)c * H * W] = acc[p][c];
        }
    }
}

// torch.nn.MaxPool2d (implicit -inf padding): x [N][H][W] -> y [N][Ho][Wo], N = B * C planes.  One CTA per 16 x 32 tile of
// outputs: the input window is staged once in shared memory (coalesced rows), then a separable max (along bins, then along
// time), so every input element is read from HBM/L2 once instead of up to 25 times.
constexpr int kMpTH = 16, kMpTW = 32;
template <int KH, int KW, int SH, int SW, int PH, int PW>
__global__ void __launch_bounds__(256) maxpool2d_kernel(const float* __restrict__ x, float* __restrict__ y, int H, int W, int Ho, int Wo) {
    constexpr int IH = (kMpTH - 1) * SH + KH, IW = (kMpTW - 1) * SW + KW;
    __shared__ float s_in[IH][IW + 1];
    __shared__ float s_row[IH][kMpTW];
    const int ho0 = blockIdx.y * kMpTH, wo0 = blockIdx.x * kMpTW, tid = threadIdx.x;
    const float* xp = x + (size_t)blockIdx.z * H * W;
    const int h_in0 = ho0 * SH - PH, w_in0 = wo0 * SW - PW;
    for (int i = tid; i < IH * IW; i += 256) {
        const int r = i / IW, q = i % IW, h = h_in0 + r, w_ = w_in0 + q;
        s_in[r][q] = (h >= 0 && h < H && w_ >= 0 && w_ < W) ? __ldg(xp + (size_t)h * W + w_) : -INFINITY;
    }
    __syncthreads();
    for (int i = tid; i < IH * kMpTW; i += 256) {
        const int r = i / kMpTW, c = i % kMpTW;
        float m = s_in[r][c * SW];
#pragma unroll
        for (int j = 1; j < KW; ++j) m = fmaxf(m, s_in[r][c * SW + j]);
        s_row[r][c] = m;
    }
    __syncthreads();
    for (int i = tid; i < kMpTH * kMpTW; i += 256) {
        const int oh = i / kMpTW, c = i % kMpTW;
        if (ho0 + oh >= Ho || wo0 + c >= Wo) continue;
        float m = s_row[oh * SH][c];
#pragma unroll
        for (int k = 1; k < KH; ++k) m = fmaxf(m, s_row[oh * SH + k][c]);
        y[((size_t)blockIdx.z * Ho + ho0 + oh) * Wo + wo0 + c] = m;
    }
}

// h3 [B][32][T][16] -> flatten (feature = channel * 16 + bin, transformer.py:337) -> conv4 (512 -> 64, folded BatchNorm1d)
// -> xf_out [B][T][64];  xf_proj = proj(xf_out) (transformer.py:458).  w4t [512][64], wpt [64][64] (input-major).
constexpr int kC4Rows = 16;
__global__ void __launch_bounds__(256) conv4_proj_kernel(const float* __restrict__ h3, const float* __restrict__ w4t, const float* __restrict__ b4,
                                                         const float* __restrict__ wpt, const float* __restrict__ bp, float* __restrict__ xf_out,
                                                         float* __restrict__ xf_proj, int B, int T) {
    __shared__ float s_f[kC4Rows][512 + 4];
    __shared__ float s_o[kC4Rows][64];
    const long row0 = (long)blockIdx.x * kC4Rows, M = (long)B * T;
    const int tid = threadIdx.x;
    for (int i = tid; i < kC4Rows * 512; i += 256) {
        const int r = i / 512, k = i % 512;                    // consecutive threads: consecutive bins of one channel plane
        const long g = row0 + r;
        float v = 0.f;
        if (g < M) {
            const long bi = g / T, t = g % T;
            v = __ldg(h3 + ((bi * 32 + (k >> 4)) * T + t) * 16 + (k & 15));
        }
        s_f[r][k] = v;
    }
    __syncthreads();
    const int o = tid & 63, rq = tid >> 6;                     // output feature, row quarter (4 rows each)
    float acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = __ldg(b4 + o);
    for (int k = 0; k < 512; ++k) {
        const float wv = __ldg(w4t + k * 64 + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(s_f[4 * rq + j][k], wv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        s_o[4 * rq + j][o] = acc[j];
        if (row0 + 4 * rq + j < M) xf_out[(row0 + 4 * rq + j) * 64 + o] = acc[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = __ldg(bp + o);
    for (int k = 0; k < 64; ++k) {
        const float wv = __ldg(wpt + k * 64 + o);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = fmaf(s_o[4 * rq + j][k], wv, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
        if (row0 + 4 * rq + j < M) xf_proj[(row0 + 4 * rq + j) * 64 + o] = acc[j];
}

}  // namespace dc
