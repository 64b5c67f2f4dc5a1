// On-device evaluation features: what the reference computes on generated batches right after the sampling
// path -- the ST-GCN motion encoder whose 64-d latents feed FGD / diversity / latent MAE, the reductions of those metrics, and the
// motion side of the beat-consistency score.  Reference: Diffusion_Stage/tools/eval_new_metrics.py:38-74 (MotionEncoder_STGCN),
// :148-241 (diversity, Frechet statistics, latent MAE), :243-303 (alignment_score, motion_peak_onehot);
// Diffusion_Stage/models/ST_GCN/ST_GCN.py:86-113, 146-228 and st_gcn_utils/tgcn.py:61-73.
//
// 0.55 MMAC per frame in 13 x 32-channel maps: far too little arithmetic for tensor-core tiles and byte-bound on its activations
// (1.6 KB per frame and layer), so these are plain fp32 CUDA-core kernels with coalesced [frame][joint][channel] activations,
// weights staged in shared memory, eval-mode BatchNorms folded into the adjacent linear maps on the host (fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dc {

constexpr int kEvV = 13, kEvC = 32, kEvLat = 64, kEvTT = 16, kEvLayers = 10;

// folded parameters of one st_gcn block, device pointers
struct StgcnLayer {
    const float* wg;      // [cin][32]      s1[c] * gcn.conv.weight[c][ci]
    const float* aeff;    // [13][13]       A * edge_importance
    const float* cg;      // [13][32]       s1[c] * gcn.bias[c] * colsum(aeff)[w] + shift1[c]
    const float* wt;      // [3][32][32]    s2[co] * tcn.conv.weight[co][ci][dt]   as [dt][ci][co]
    const float* ct;      // [32]           s2[co] * tcn.conv.bias[co] + shift2[co]
};

// One st_gcn block (ST_GCN.py:217-228): y = ReLU( BN2(conv_t3( ReLU(BN1( graphmix(conv1x1(x)) )) )) + res(x) ).
// x [N][T][13][CIN], y [N][T][13][32]; grid (ceil(T / 16), N).  The temporal convolution zero-pads its INPUT (the post-ReLU graph
// features) at the ends of the sequence (padding = (1, 0)), so frames outside [0, T) contribute exactly 0.
// dbn != null (first block): data_bn folded to a per-(joint, coordinate) affine applied while loading (ST_GCN.py:97-103).
template <int CIN>
__global__ void __launch_bounds__(256) stgcn_layer_kernel(const float* __restrict__ x, float* __restrict__ y, int T, StgcnLayer p,
                                                         const float* __restrict__ dbn /* [2][13 * CIN] scale | shift */, int residual) {
    extern __shared__ float sm[];
    float* xin = sm;                                          // [(TT + 2)][13][CIN]
    float* g = xin + (kEvTT + 2) * kEvV * CIN;                // [(TT + 2)][13][32]
    float* wg = g + (kEvTT + 2) * kEvV * kEvC;                // [CIN][32]
    float* ae = wg + CIN * kEvC;                              // [13][13]
    float* cg = ae + kEvV * kEvV;                             // [13][32]
    float* wt = cg + kEvV * kEvC;                             // [3][32][32]
    float* ct = wt + 3 * kEvC * kEvC;                         // [32]
    const int tid = threadIdx.x, n = blockIdx.y, t0 = blockIdx.x * kEvTT;
    const float* xn = x + (size_t)n * T * kEvV * CIN;
    for (int i = tid; i < (kEvTT + 2) * kEvV * CIN; i += 256) {
        const int tt = i / (kEvV * CIN), rem = i - tt * (kEvV * CIN), t = t0 - 1 + tt;
        float v = 0.f;
        if (t >= 0 && t < T) {
            v = __ldg(xn + (size_t)t * kEvV * CIN + rem);
            if (dbn != nullptr) v = fmaf(v, __ldg(dbn + rem), __ldg(dbn + kEvV * CIN + rem));
        }
        xin[i] = v;
    }
    for (int i = tid; i < CIN * kEvC; i += 256) wg[i] = __ldg(p.wg + i);
    for (int i = tid; i < kEvV * kEvV; i += 256) ae[i] = __ldg(p.aeff + i);
    for (int i = tid; i < kEvV * kEvC; i += 256) cg[i] = __ldg(p.cg + i);
    for (int i = tid; i < 3 * kEvC * kEvC; i += 256) wt[i] = __ldg(p.wt + i);
    if (tid < kEvC) ct[tid] = __ldg(p.ct + tid);
    __syncthreads();
    // ---- graph convolution: thread -> (frame, channel): 1x1 convolution at the 13 joints, then the 13 x 13 joint mix
    for (int i = tid; i < (kEvTT + 2) * kEvC; i += 256) {
        const int tt = i >> 5, c = i & 31, t = t0 - 1 + tt;
        float* gt = g + tt * kEvV * kEvC + c;
        if (t < 0 || t >= T) {
#pragma unroll
            for (int w = 0; w < kEvV; ++w) gt[w * kEvC] = 0.f;
            continue;
        }
        float z[kEvV];
        const float* xt = xin + tt * kEvV * CIN;
#pragma unroll
        for (int v = 0; v < kEvV; ++v) {
            float acc = 0.f;
#pragma unroll
            for (int ci = 0; ci < CIN; ++ci) acc = fmaf(xt[v * CIN + ci], wg[ci * kEvC + c], acc);
            z[v] = acc;
        }
#pragma unroll
        for (int w = 0; w < kEvV; ++w) {
            float acc = cg[w * kEvC + c];
#pragma unroll
            for (int v = 0; v < kEvV; ++v) acc = fmaf(z[v], ae[v * kEvV + w], acc);
            gt[w * kEvC] = fmaxf(acc, 0.f);
        }
    }
    __syncthreads();
    // ---- temporal convolution (3 x 1) + BN2 + residual + ReLU: a warp -> one (frame, joint), lane = output channel
    const int lane = tid & 31, warp = tid >> 5;
    float* yn = y + (size_t)n * T * kEvV * kEvC;
    for (int i = warp; i < kEvTT * kEvV; i += 8) {
        const int tt = i / kEvV, v = i - tt * kEvV, t = t0 + tt;
        if (t >= T) break;
        float acc = ct[lane];
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            const float* gp = g + ((tt + dt) * kEvV + v) * kEvC;
            const float* wp = wt + dt * kEvC * kEvC + lane;
#pragma unroll
            for (int ci = 0; ci < kEvC; ++ci) acc = fmaf(gp[ci], wp[ci * kEvC], acc);
        }
        if (CIN == kEvC && residual) acc += xin[((tt + 1) * kEvV + v) * CIN + lane];
        yn[((size_t)t * kEvV + v) * kEvC + lane] = fmaxf(acc, 0.f);
    }
}
constexpr int stgcn_smem_bytes(int cin) {
    return ((kEvTT + 2) * kEvV * cin + (kEvTT + 2) * kEvV * kEvC + cin * kEvC + kEvV * kEvV + kEvV * kEvC + 3 * kEvC * kEvC + kEvC) * 4;
}

// fc = Conv1d(416, 64, 1) + BatchNorm1d(64) on the flattened (channel-major) joint features (eval_new_metrics.py:49, 70-72):
// x [rows][13][32] (our order: joint, channel), w [416 = v * 32 + c][64] and b [64] with the BatchNorm folded.  8 frames per CTA.
__global__ void __launch_bounds__(256) stgcn_fc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                                                      float* __restrict__ out, long rows) {
    __shared__ float xs[8][kEvV * kEvC];
    const long r0 = (long)blockIdx.x * 8;
    const int tid = threadIdx.x;
    for (int i = tid; i < 8 * kEvV * kEvC; i += 256) {
        const int f = i / (kEvV * kEvC), k = i - f * (kEvV * kEvC);
        xs[f][k] = r0 + f < rows ? __ldg(x + (r0 + f) * (kEvV * kEvC) + k) : 0.f;
    }
    __syncthreads();
    const int o = tid & 63, fq = tid >> 6;                     // frames fq, fq + 4
    float a0 = __ldg(b + o), a1 = a0;
    for (int k = 0; k < kEvV * kEvC; ++k) {
        const float wv = __ldg(w + k * kEvLat + o);
        a0 = fmaf(xs[fq][k], wv, a0), a1 = fmaf(xs[fq + 4][k], wv, a1);
    }
    if (r0 + fq < rows) out[(r0 + fq) * kEvLat + o] = a0;
    if (r0 + fq + 4 < rows) out[(r0 + fq + 4) * kEvLat + o] = a1;
}

// ---------------------------------------------------------------- statistics of the latents (FGD, eval_new_metrics.py:164-168)
// column sums in fp64: sum[64]
__global__ void __launch_bounds__(256) feat_colsum_kernel(const float* __restrict__ f, long rows, double* __restrict__ sum) {
    const int c = threadIdx.x & 63, q = threadIdx.x >> 6;
    double acc = 0.0;
    for (long r = (long)blockIdx.x * 4 + q; r < rows; r += (long)gridDim.x * 4) acc += (double)__ldg(f + r * kEvLat + c);
    __shared__ double s[256];
    s[threadIdx.x] = acc;
    __syncthreads();
    if (q == 0) atomicAdd(sum + c, s[c] + s[64 + c] + s[128 + c] + s[192 + c]);
}
// centred second moments in fp64: m2[i][j] += sum_r (f[r][i] - mu[i]) (f[r][j] - mu[j]);  mu = sum / rows.  A CTA takes a slab of
// rows through shared memory; thread (i, j-group of 16) accumulates 16 entries.
__global__ void __launch_bounds__(256) feat_cov_kernel(const float* __restrict__ f, long rows, const double* __restrict__ sum, double* __restrict__ m2) {
    __shared__ float xs[32][kEvLat];
    __shared__ double mud[kEvLat];
    const int tid = threadIdx.x;
    if (tid < kEvLat) mud[tid] = sum[tid] / (double)rows;
    __syncthreads();
    const int i = tid >> 2, j0 = (tid & 3) * 16;
    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    for (long r0 = (long)blockIdx.x * 32; r0 < rows; r0 += (long)gridDim.x * 32) {
        __syncthreads();
        for (int k = tid; k < 32 * kEvLat; k += 256) {
            const long r = r0 + (k >> 6);
            xs[k >> 6][k & 63] = r < rows ? __ldg(f + r * kEvLat + (k & 63)) : 0.f;
        }
        __syncthreads();
        const int nr = (int)min((long)32, rows - r0);
        for (int r = 0; r < nr; ++r) {
            const double di = (double)xs[r][i] - mud[i];
#pragma unroll
            for (int k = 0; k < 16; ++k) acc[k] = fma(di, (double)xs[r][j0 + k] - mud[j0 + k], acc[k]);
        }
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) atomicAdd(m2 + i * kEvLat + j0 + k, acc[k]);
}
// sum over rows of sum_c |a - b| in fp64 (diversity and latent MAE, eval_new_metrics.py:154, 181-185)
__global__ void __launch_bounds__(256) feat_l1_kernel(const float* __restrict__ a, const float* __restrict__ b, long n, double* __restrict__ out) {
    double acc = 0.0;
    for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long)gridDim.x * 256) acc += (double)fabsf(__ldg(a + i) - __ldg(b + i));
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    __shared__ double s[8];
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) atomicAdd(out, ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
}

// ---------------------------------------------------------------- motion beats (eval_new_metrics.py:277-303)
// envelope[n][t] = sum over the 13 joints of |x[t] - x[t-1]|_2 (0 at t = 0); motion [N][T][26]
__global__ void __launch_bounds__(256) motion_envelope_kernel(const float* __restrict__ m, float* __restrict__ env, int N, int T) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long)N * T) return;
    const int t = (int)(i % T);
    float e = 0.f;
    if (t > 0) {
        const float* p = m + i * 26;
#pragma unroll
        for (int v = 0; v < kEvV; ++v) {
            const float dx = __fsub_rn(p[2 * v], p[2 * v - 26]), dy = __fsub_rn(p[2 * v + 1], p[2 * v + 1 - 26]);
            e = __fadd_rn(e, __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
        }
    }
    env[i] = e;
}
// beats[n][t] = 1 where the envelope is strictly below its `order` neighbours on both sides (scipy argrelextrema(np.less, order,
// mode='clip'): neighbours beyond the ends repeat the end sample, so the end frames never qualify)
__global__ void __launch_bounds__(256) local_minima_kernel(const float* __restrict__ env, uint8_t* __restrict__ beats, int N, int T, int order) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long)N * T) return;
    const int t = (int)(i % T);
    const float* e = env + (i - t);
    const float c = e[t];
    bool ok = true;
    for (int k = 1; k <= order && ok; ++k) ok = c < e[min(t + k, T - 1)] && c < e[max(t - k, 0)];
    beats[i] = ok ? 1 : 0;
}
// alignment_score (eval_new_metrics.py:243-267): one CTA per clip; for every music beat the distance (in indices, as the reference
// compares them) to the nearest motion beat -> exp(-d^2 / 2 sigma^2); mean over the music beats; 0 without motion beats.
__global__ void __launch_bounds__(256) beat_alignment_kernel(const uint8_t* __restrict__ music, int Tm, const uint8_t* __restrict__ motion, int T,
                                                            float sigma, float* __restrict__ score) {
    const uint8_t* mu = music + (size_t)blockIdx.x * Tm;
    const uint8_t* mo = motion + (size_t)blockIdx.x * T;
    __shared__ float s_sum[256];
    __shared__ int s_cnt[256], s_any[256];
    float sum = 0.f;
    int cnt = 0, any = 0;
    for (int t = threadIdx.x; t < T; t += 256) any |= mo[t];
    for (int i = threadIdx.x; i < Tm; i += 256) {
        if (!mu[i]) continue;
        int best = 0x7FFFFFFF;
        for (int d = 0; d < max(T, Tm) && best == 0x7FFFFFFF; ++d) {           // nearest motion beat: scan outwards
            const int lo = i - d, hi = i + d;
            if ((lo >= 0 && lo < T && mo[lo]) || (hi >= 0 && hi < T && mo[hi])) best = d;
        }
        if (best != 0x7FFFFFFF) {
            const float df = (float)best;
            sum += expf(-(df * df) / 2.f / (sigma * sigma));
        }
        ++cnt;
    }
    s_sum[threadIdx.x] = sum, s_cnt[threadIdx.x] = cnt, s_any[threadIdx.x] = any;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) s_sum[threadIdx.x] += s_sum[threadIdx.x + o], s_cnt[threadIdx.x] += s_cnt[threadIdx.x + o], s_any[threadIdx.x] |= s_any[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) score[blockIdx.x] = (s_any[0] && s_cnt[0] > 0) ? s_sum[0] / (float)s_cnt[0] : 0.f;
}

}  // namespace dc
