// Host side of the C ABI declared in include/dc_b200.h: weight ingestion / packing, workspaces,
// the per-step launch schedule, CUDA-graph capture of the sampling loop.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/dc_b200.h"
#include "clip_kernel.cuh"
#include "music_encoder_tc.cuh"
#include "eval_kernels.cuh"

using namespace dc;

namespace {

thread_local std::string g_last_error;

struct HostTensor {
    std::vector<float> v;
    std::vector<int64_t> shape;
};

// byte offsets of the packed matrices inside one layer's 1 MiB slab (schedule order)
constexpr uint32_t kOffWeSa = 0;
constexpr uint32_t kOffWoSa = kOffWeSa + 8 * 32768;
constexpr uint32_t kOffWeCa = kOffWoSa + 32768;
constexpr uint32_t kOffWqCa = kOffWeCa + 8 * 32768;
constexpr uint32_t kOffWoCa = kOffWqCa + 32768;
constexpr uint32_t kOffWeFf = kOffWoCa + 32768;
constexpr uint32_t kOffW1 = kOffWeFf + 8 * 32768;
constexpr uint32_t kOffW2 = kOffW1 + 16384;
constexpr uint32_t kOffWoFf = kOffW2 + 16384;
constexpr uint32_t kOffWq = kOffWoFf + 32768;
constexpr uint32_t kOffWk = kOffWq + 32768;
constexpr uint32_t kOffWv = kOffWk + 32768;
constexpr uint32_t kLayerSlab = kOffWv + 32768;
static_assert(kLayerSlab == 1u << 20, "layer slab is 1 MiB");

struct GraphKey {
    int sampler = -1;
    const void* noise = nullptr;
    const void* tx0 = nullptr;
    const void* tx = nullptr;
    int steps = 0;
    bool operator==(const GraphKey& o) const {
        return sampler == o.sampler && noise == o.noise && tx0 == o.tx0 && tx == o.tx && steps == o.steps;
    }
};

}  // namespace

struct dc_handle {
    dc_config cfg{};
    bool bf16 = true;
    bool finalized = false;
    std::string err;
    std::map<std::string, HostTensor> w;

    // packed / derived weights (device, owned)
    uint8_t* wbuf = nullptr;      // [L][1 MiB]
    float* prm = nullptr;         // [L][kPrmFloats]
    float* prm_clip = nullptr;    // [L][kPrmFloats] variant read by the persistent kernel (FFN-up fused across the Wo_ca residual add)
    uint8_t* wfuse = nullptr;     // [L][16 KB] (W1 . Wo_ca) [64 x 128] operand image
    float* kshift = nullptr;      // [L][128] static shift of the time-axis softmax (bound of |k| per key column)
    uint32_t static_mask = 0;     // bit l: layer l's bound is small enough to replace the running column max
    uint8_t* wkv = nullptr;       // [L][8][256 x 128 B] folded cross-attention K|V weights
    float* bkv = nullptr;         // [L][256]
    float *WjT = nullptr, *bj = nullptr, *pos = nullptr, *WoT = nullptr, *bo = nullptr;
    uint8_t* wj_img = nullptr;    // joint_embed as a [128 x 128] operand image [W_hi | W_hi | W_lo] (x is fed as hi | lo | hi)
    uint8_t* wout_img = nullptr;  // output head as a [32 x 384] operand image [W_hi | W_hi | W_lo] (h is fed as hi | lo | hi)
    float *WlinT = nullptr, *blin = nullptr;
    float *teW0 = nullptr, *teb0 = nullptr, *teW2 = nullptr, *teb2 = nullptr, *freqs = nullptr;

    // music encoder (BatchNorm folded): per 3x3 layer w [CIN][9][COUT], b [COUT]; conv2.0's 1x1 residual; conv4 + proj
    bool has_music = false;
    Conv10Weights me_c10;          // conv1.0 (1 -> 16 channels): HOST copy, passed by value at launch (CUDA cores)
    uint8_t* me_wimg = nullptr;    // the six tensor-core layers: (hi, lo)-split B-operand blocks, see music_encoder_tc.cuh
    float* me_bias = nullptr;      // [6][64]: folded bias [COUT] (+ [COUT] of conv2.0's 1x1 residual)
    size_t me_woff[6] = {0, 0, 0, 0, 0, 0};
    int me_occ[4] = {1, 1, 1, 1};  // resident CTAs per SM of the four conv_tc_kernel instances (conv1.x, conv2.0, conv2.1, conv3.x)
    float *me_w4t = nullptr, *me_b4 = nullptr, *me_wpt = nullptr, *me_bp = nullptr;
    float *me_buf0 = nullptr, *me_buf1 = nullptr;     // ping-pong activation planes for one chunk of clips
    size_t me_cap = 0;                                // floats per buffer

    // schedule
    int S = 0;
    float* coef = nullptr;        // [S][8]
    float* te_table = nullptr;    // [S][512]
    int* step_ctr = nullptr;

    // workspace for the prepared condition
    int B = 0, T = 0, M = 0, tiles = 0;
    size_t cap_tokens = 0;
    int cap_B = 0;
    size_t cap_clip_tiles = 0;
    float* xp = nullptr;
    uint8_t* zimg = nullptr;
    uint8_t* aemb = nullptr;
    size_t aemb_stride = 0;
    float* hbuf = nullptr;
    uint8_t* q_img = nullptr;     // [tiles][32 KB] packed softmax_hd(Q) image
    float* kv = nullptr;
    uint8_t* bd_sa = nullptr;     // [B][32 KB] block-diagonal self-attention K^T V images
    uint8_t* bd_ca = nullptr;     // [B][L][32 KB] cross-attention counterparts (step-invariant)
    int mask_invert = 0;
    size_t kv_stride = 0;         // floats between the per-layer slices of `kv`
    int kv_layers = 1;            // slices allocated
    float* kv_part = nullptr;     // [tiles][2][kKvPartFloats] partial time-axis reductions (per-layer path)
    int* clip_cnt = nullptr;      // [B]
    bool fuse_kv = false;
    bool persist = false;         // persistent sampling-loop kernel: one thread-block cluster per clip (T <= 16 tiles, any batch)
    int clip_nt = 1;              // tiles (= cluster size) per clip
    int clip_nt_checked = 0;      // last cluster size validated with cudaOccupancyMaxActiveClusters
    bool clip_gx = false;         // per-clip exchange through global memory (cluster size 1) instead of distributed shared memory
    uint2* gx_part = nullptr;     // [B][2][nt][kKvPartFloats] (value, tag)
    uint2* gx_slice = nullptr;    // [B][2][128][8] (two 16-bit values, tag)
    uint32_t gx_tag = 0;          // tags handed out so far (every reduction of every launch gets its own)
    size_t gx_cap = 0, gx_cap_b = 0;   // B * nt and B the three buffers were sized for
    int num_sms = 0;
    unsigned long long* timeline = nullptr;   // debug: [launch][512] u64 (dc_debug_timeline)
    bool timeline_on = false;
    long long* length = nullptr;
    bool has_length = false;
    float* te_b = nullptr;        // [B][512]
    float* xwork = nullptr;       // [M][26]
    float* x0work = nullptr;      // [M][26]
    float* in_proj = nullptr;     // staging for dc_generate_host
    float* in_out = nullptr;
    bool prepared = false;

    // graphs
    bool use_graphs = true;
    bool use_pair = false;         // cta_group::2 CTA pairs in the layer kernel (DC_PAIR=1); measured r01: 7 % slower on
                                   // C2/C3 -- the kernel is bound by the dependent chain, and pairing adds signalling latency
    bool use_pdl = false;          // measured slightly slower on C2 (r01): kernels cannot co-reside with the 213 KB layer CTA
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t gexec = nullptr;
    GraphKey gkey;
    int64_t launches = 0;

    // per-kernel-class event timing (dc_profile_step)
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_cls;
};

namespace {

int fail(dc_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (h) h->err = buf;
    return code;
}

#define DC_CUDA(h, expr)                                                                                          \
    do {                                                                                                          \
        cudaError_t e_ = (expr);                                                                                  \
        if (e_ != cudaSuccess) return fail(h, DC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), \
                                           __FILE__, __LINE__);                                                   \
    } while (0)

uint16_t to16(float f, bool bf16) {
    if (bf16) {
        __nv_bfloat16 b = __float2bfloat16_rn(f);
        uint16_t u;
        memcpy(&u, &b, 2);
        return u;
    }
    __half hh = __float2half_rn(f);
    uint16_t u;
    memcpy(&u, &hh, 2);
    return u;
}

// fp32 [*, ldw] -> K-major SW128 image at dst (bytes), Kblocks blocks of Nrows x 128 B
void pack_image(uint8_t* dst, const float* W, int ldw, int Ksrc, const int* rowmap, const float* colscale, int Nrows,
                int Kblocks, bool bf16) {
    for (int n = 0; n < Nrows; ++n) {
        const int src = rowmap ? rowmap[n] : n;
        for (int k = 0; k < Kblocks * 64; ++k) {
            float v = 0.f;
            if (src >= 0 && k < Ksrc) {
                v = W[(size_t)src * ldw + k];
                if (colscale) v *= colscale[k];
            }
            const int kb = k >> 6, c = (k & 63) >> 3, e = k & 7;
            const size_t off = (size_t)kb * Nrows * 128 + sw128_offset((uint32_t)n, (uint32_t)c) + e * 2;
            const uint16_t u = to16(v, bf16);
            memcpy(dst + off, &u, 2);
        }
    }
}

template <class T>
int upload(dc_handle* h, T** dptr, const void* src, size_t bytes) {
    if (*dptr) cudaFree(*dptr);
    *dptr = nullptr;
    DC_CUDA(h, cudaMalloc((void**)dptr, bytes));
    DC_CUDA(h, cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice));
    return 0;
}

const HostTensor* get(dc_handle* h, const std::string& key, std::initializer_list<int64_t> shape) {
    auto it = h->w.find(key);
    if (it == h->w.end()) {
        fail(h, DC_ERR_INVALID, "missing weight '%s'", key.c_str());
        return nullptr;
    }
    std::vector<int64_t> want(shape);
    if (it->second.shape != want) {
        std::string got;
        for (auto d : it->second.shape) got += std::to_string(d) + ",";
        fail(h, DC_ERR_INVALID, "weight '%s' has shape [%s] (unexpected)", key.c_str(), got.c_str());
        return nullptr;
    }
    return &it->second;
}

DOp make_dop(uint32_t w_off, uint32_t w_bytes, int kb, int n, uint32_t d_col, bool acc, int wait, int commit, int seg = 0,
             bool releases_s = false, bool ring_a = false) {
    DOp o{};
    o.w_off = w_off;
    o.w_bytes = w_bytes;
    o.n = (uint16_t)n;
    o.d_col = (uint16_t)d_col;
    o.kb = (uint8_t)kb;
    o.accumulate = acc;
    o.wait = (uint8_t)wait;
    o.commit = (uint8_t)commit;
    o.seg = (uint8_t)seg;
    o.releases_s = releases_s;
    o.ring_a = ring_a;
    return o;
}

void free_workspace(dc_handle* h) {
    if (h->kv_part) cudaFree(h->kv_part);
    if (h->clip_cnt) cudaFree(h->clip_cnt);
    h->kv_part = nullptr, h->clip_cnt = nullptr;
    if (h->gx_part) cudaFree(h->gx_part);
    if (h->gx_slice) cudaFree(h->gx_slice);
    h->gx_part = nullptr, h->gx_slice = nullptr, h->gx_cap = 0, h->gx_cap_b = 0;
    void* ptrs[] = {h->xp, h->zimg, h->aemb, h->hbuf, h->q_img, h->kv, h->bd_sa, h->bd_ca, h->length,
                    h->te_b, h->xwork, h->x0work, h->in_proj, h->in_out};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    h->xp = nullptr, h->zimg = nullptr, h->aemb = nullptr, h->hbuf = nullptr, h->q_img = nullptr, h->kv = nullptr;
    h->bd_sa = nullptr, h->bd_ca = nullptr, h->length = nullptr, h->te_b = nullptr, h->xwork = nullptr, h->x0work = nullptr;
    h->in_proj = nullptr, h->in_out = nullptr;
    h->cap_tokens = 0, h->cap_B = 0, h->cap_clip_tiles = 0;
}

void drop_graph(dc_handle* h) {
    if (h->gexec) cudaGraphExecDestroy(h->gexec);
    h->gexec = nullptr;
    h->gkey = GraphKey{};
}

int ensure_workspace(dc_handle* h, int B, int T) {
    const size_t M = (size_t)B * T;
    const size_t tiles = (((M + kTileRows - 1) / kTileRows) + 1) & ~size_t(1);     // even: CTA pairs may add a padding tile
    const size_t clip_tiles = (size_t)B * ((T + kTileRows - 1) / kTileRows);       // clip-aligned tiling of the cluster kernel
    if (M <= h->cap_tokens && B <= h->cap_B && clip_tiles <= h->cap_clip_tiles) return 0;
    drop_graph(h);
    free_workspace(h);
    const size_t Mpad = tiles * kTileRows;
    const int L = h->cfg.num_layers;
    DC_CUDA(h, cudaMalloc((void**)&h->xp, Mpad * kE * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->zimg, tiles * 8 * (size_t)kABlockBytes));
    const size_t aemb_tiles = std::max(tiles, clip_tiles);
    h->aemb_stride = aemb_tiles * 8 * (size_t)kABlockBytes;           // two images (step parity) for the persistent kernel
    DC_CUDA(h, cudaMalloc((void**)&h->aemb, 2 * h->aemb_stride));
    DC_CUDA(h, cudaMalloc((void**)&h->hbuf, Mpad * kD * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->q_img, tiles * (size_t)kAworkBytes));
    // K | V workspace: one [Mpad][256] slice per layer of a precompute chunk (as many layers as fit 1 GiB, at least one)
    h->kv_stride = Mpad * 256;
    h->kv_layers = (int)std::max<size_t>(1, std::min<size_t>((size_t)L, ((size_t)1 << 30) / (h->kv_stride * 4)));
    DC_CUDA(h, cudaMalloc((void**)&h->kv, h->kv_stride * 4 * h->kv_layers));
    DC_CUDA(h, cudaMalloc((void**)&h->bd_sa, (size_t)B * kAworkBytes));
    DC_CUDA(h, cudaMalloc((void**)&h->bd_ca, (size_t)B * L * kAworkBytes));
    DC_CUDA(h, cudaMalloc((void**)&h->length, (size_t)B * 8));
    DC_CUDA(h, cudaMalloc((void**)&h->kv_part, tiles * 2 * (size_t)kKvPartFloats * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->clip_cnt, (size_t)B * 4));
    DC_CUDA(h, cudaMemset(h->clip_cnt, 0, (size_t)B * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->te_b, (size_t)B * kE * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->xwork, Mpad * kP * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->x0work, Mpad * kP * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->in_proj, M * kMusic * 4));
    DC_CUDA(h, cudaMalloc((void**)&h->in_out, M * kMusic * 4));
    // padded rows of the operand images must hold finite values
    DC_CUDA(h, cudaMemset(h->zimg, 0, tiles * 8 * (size_t)kABlockBytes));
    DC_CUDA(h, cudaMemset(h->aemb, 0, 2 * h->aemb_stride));
    DC_CUDA(h, cudaMemset(h->q_img, 0, tiles * (size_t)kAworkBytes));
    // off-diagonal head blocks of the attention images are never written: they must be zero
    DC_CUDA(h, cudaMemset(h->bd_sa, 0, (size_t)B * kAworkBytes));
    DC_CUDA(h, cudaMemset(h->bd_ca, 0, (size_t)B * L * kAworkBytes));
    h->cap_tokens = M;
    h->cap_B = B;
    h->cap_clip_tiles = clip_tiles;
    return 0;
}

int init_kernel_attrs(dc_handle* h) {
    DC_CUDA(h, cudaFuncSetAttribute(gemm_rows_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    DC_CUDA(h, cudaFuncSetAttribute(gemm_rows_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmemBytes));
    DC_CUDA(h, cudaFuncSetAttribute(layer_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLayerSmemBytes));
    DC_CUDA(h, cudaFuncSetAttribute(layer_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLayerSmemBytes));
    DC_CUDA(h, cudaFuncSetAttribute(layer_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLayerSmemBytes));
    DC_CUDA(h, cudaFuncSetAttribute(layer_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kLayerSmemBytes));
    {   // music-encoder convolutions: opt in to their shared memory; resident CTAs per SM = min over shared memory (227 KB, 1 KB
        // reserved per CTA), registers and 4 x 128 TMEM columns -- the persistent grids are sized to exactly that
        const void* fn[4] = {(const void*)conv_tc_kernel<16, 16, 1, 128, kMeNM128>, (const void*)conv_tc_kernel<16, 32, 2, 64, kMeNM64>,
                             (const void*)conv_tc_kernel<32, 32, 1, 64, kMeNM64b>, (const void*)conv_tc_kernel<32, 32, 1, 32, kMeNM32>};
        const int sm[4] = {me_smem_bytes<16, 16, 1, 128, kMeNM128>(), me_smem_bytes<16, 32, 2, 64, kMeNM64>(), me_smem_bytes<32, 32, 1, 64, kMeNM64b>(),
                           me_smem_bytes<32, 32, 1, 32, kMeNM32>()};
        for (int i = 0; i < 4; ++i) {
            DC_CUDA(h, cudaFuncSetAttribute(fn[i], cudaFuncAttributeMaxDynamicSharedMemorySize, sm[i]));
            cudaFuncAttributes fa{};
            DC_CUDA(h, cudaFuncGetAttributes(&fa, fn[i]));
            const int by_smem = (227 * 1024) / (sm[i] + (int)fa.sharedSizeBytes + 1024);
            const int by_regs = 65536 / (((fa.numRegs + 7) & ~7) * ((kMeThreads + 31) & ~31));
            if (h) h->me_occ[i] = std::max(1, std::min(std::min(by_smem, by_regs), 4));           // (h is null in dc_selftest_gemm)
            if (h && getenv("DC_VERBOSE")) fprintf(stderr, "[dc_b200] conv_tc variant %d: %d regs, %d B smem -> %d CTAs/SM\n", i, fa.numRegs, sm[i], h->me_occ[i]);
        }
    }
    void (*clip_variants[12])(StepArgs) = {clip_kernel<true, false>, clip_kernel<false, false>, clip_kernel<true, true>, clip_kernel<false, true>,
                                           clip_kernel<true, false, true>, clip_kernel<false, false, true>, clip_kernel<true, true, true>,
                                           clip_kernel<false, true, true>, clip_kernel<true, false, false, true>, clip_kernel<false, false, false, true>,
                                           clip_kernel<true, true, false, true>, clip_kernel<false, true, false, true>};
    for (auto k : clip_variants) {
        DC_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, kClipSmemBytes));
        DC_CUDA(h, cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));     // clusters of 9..16 tiles
    }
    return 0;
}

// Launch with the programmatic-dependent-launch attribute: the kernel may be scheduled while its predecessor
// in the stream drains; every kernel of the step calls griddepcontrol.wait before touching activations.
template <class... KArgs, class... Args>
cudaError_t launch_kc(bool pdl, int cluster, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (pdl) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
template <class... KArgs, class... Args>
cudaError_t launch_k(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    return launch_kc(pdl, 1, kern, grid, block, smem, st, args...);
}

template <bool kBf16>
int launch_gemm_rows(dc_handle* h, const GemmRowsArgs& ga, int tiles, cudaStream_t st, int layers = 1) {
    gemm_rows_kernel<kBf16><<<dim3((unsigned)tiles, (unsigned)layers), kTileThreads, kGemmSmemBytes, st>>>(ga);
    return 0;
}

template <bool kBf16>
int launch_layer(dc_handle* h, const LayerArgs& la, int tiles, cudaStream_t st) {
    if (h->use_pair)      // CTA pairs (cta_group::2): grid padded to an even number of tiles
        DC_CUDA(h, launch_kc(h->use_pdl, 2, layer_kernel<kBf16, true>, dim3((tiles + 1) & ~1), dim3(kTileThreads), kLayerSmemBytes, st, la));
    else
        DC_CUDA(h, launch_k(h->use_pdl, layer_kernel<kBf16, false>, dim3(tiles), dim3(kTileThreads), kLayerSmemBytes, st, la));
    return 0;
}

// Build the launch arguments of the tile kernel that finishes layer `l` (l = -1: only the q/k/v
// head of layer 0) and starts layer l+1.
LayerArgs layer_args(dc_handle* h, int l) {
    const int L = h->cfg.num_layers;
    LayerArgs a{};
    a.do_main = l >= 0;
    a.do_sa1 = l + 1 < L;
    int n = 0;
    if (a.do_main) {
        const uint32_t base = (uint32_t)l * kLayerSlab;
        a.sop_w_off[0] = base + kOffWeSa;
        a.sop_w_off[1] = base + kOffWeCa;
        a.sop_w_off[2] = base + kOffWeFf;
        a.n_s = 3;
        a.dops[n++] = make_dop(0, 32768, 2, 128, kColW, false, 2, 2, 1);                       // y = q . blockdiag(A_sa)
        a.dops[n++] = make_dop(base + kOffWoSa, 32768, 2, 128, kColH, true, 1, 1, 0, true);    // h += . Wo_sa
        a.dops[n++] = make_dop(base + kOffWqCa, 32768, 2, 128, kColW, false, 1, 2);            // q_ca
        a.dops[n++] = make_dop(0, 32768, 2, 128, kColW, false, 1, 2, 2);                       // y = softmax(q) . blockdiag(A_ca)
        a.dops[n++] = make_dop(base + kOffWoCa, 32768, 2, 128, kColH, true, 1, 1, 0, true);    // h += . Wo_ca
        a.dops[n++] = make_dop(base + kOffW1, 16384, 2, 64, kColW, false, 1, 2);               // FFN up
        a.dops[n++] = make_dop(base + kOffW2, 16384, 1, 128, kColW, false, 1, 2);              // FFN down
        a.dops[n++] = make_dop(base + kOffWoFf, 32768, 2, 128, kColH, true, 1, 1);             // h += . Wo_ffn
    }
    if (a.do_sa1) {
        const uint32_t base = (uint32_t)(l + 1) * kLayerSlab;
        a.dops[n++] = make_dop(base + kOffWq, 32768, 2, 128, kColS, false, 1, 255);
        a.dops[n++] = make_dop(base + kOffWk, 32768, 2, 128, kColS + 128, false, 0, 255, 0, false, true);
        a.dops[n++] = make_dop(base + kOffWv, 32768, 2, 128, kColW, false, 0, 2, 0, false, true);
        if (h->fuse_kv && !h->use_pair) a.dops[n++] = make_dop(0, 0, 1, 128, kColW, false, 1, 255, 3);   // K^T V partial (tensor cores)
    }
    a.n_d = n;
    a.M = h->M;
    a.T = h->T;
    a.mask_invert = h->mask_invert;
    a.wbuf = h->wbuf;
    a.aemb = h->aemb;
    a.prm = h->prm + (size_t)(l >= 0 ? l : 0) * kPrmFloats;
    a.prm_next = h->prm + (size_t)(l + 1 < L ? l + 1 : 0) * kPrmFloats;
    a.h = h->hbuf;
    a.q_img = h->q_img;
    a.kv = h->kv;
    a.bd_sa = h->bd_sa;
    a.bd_ca = h->bd_ca + (size_t)(l >= 0 ? l : 0) * kAworkBytes;
    a.bd_ca_stride = (size_t)L * kAworkBytes;
    a.length = h->has_length ? h->length : nullptr;
    a.fuse_kv = h->fuse_kv;
    a.kv_part = h->kv_part;
    a.clip_cnt = h->clip_cnt;
    a.bd_sa_out = h->bd_sa;
    a.timeline = h->timeline_on ? h->timeline + (size_t)(l + 1) * 512 : nullptr;
    return a;
}

// Persistent path: n_steps consecutive denoise steps (timestep indices step0, step0 - 1, ...) in ONE launch.
// te: time-embedding base; the row of a step is te + timestep * te_step_stride (+ clip * te_stride).
int enqueue_persistent(dc_handle* h, const float* x_in, const float* te, int te_stride, int te_step_stride, int mode, float* x_upd,
                       float* x0_out, size_t x0_stride, float* x_trace, const float* noise, size_t noise_stride, int step0, int n_steps,
                       cudaStream_t st) {
    const int L = h->cfg.num_layers;
    StepArgs sa{};
    sa.L = L, sa.M = h->M, sa.T = h->T;
    sa.n_steps = n_steps, sa.step0 = step0;
    sa.wbuf = h->wbuf, sa.aemb = h->aemb, sa.aemb_out = h->aemb, sa.aemb_stride = h->aemb_stride, sa.prm = h->prm_clip, sa.wfuse = h->wfuse;
    sa.kshift = h->kshift, sa.static_mask = h->static_mask;
    sa.bd_ca = h->bd_ca, sa.bd_ca_stride = (size_t)L * kBdcBytes;          // compact head-block images (written so by dc_prepare_cond when persist)
    sa.length = h->has_length ? h->length : nullptr;
    sa.x_in = x_in, sa.x_out = x_upd, sa.x0_out = x0_out, sa.x0_stride = x0_stride, sa.x_trace = x_trace;
    sa.noise = noise, sa.noise_stride = noise_stride, sa.xp = h->xp;
    sa.te = te, sa.te_stride = te_stride, sa.te_step_stride = te_step_stride;
    sa.coef = (mode & 0xF) ? h->coef : nullptr;
    sa.mode = mode;
    sa.wj_img = h->wj_img, sa.bj = h->bj, sa.pos = h->pos, sa.wout_img = h->wout_img, sa.bo = h->bo;
    const uint32_t offs[12] = {kOffWeSa, kOffWoSa, kOffWeCa, kOffWqCa, kOffWoCa, kOffWeFf, kOffW1, kOffW2, kOffWoFf, kOffWq, kOffWk, kOffWv};
    for (int i = 0; i < 12; ++i) sa.off[i] = offs[i];
    sa.timeline = h->timeline_on ? h->timeline : nullptr;
    if (const char* dbg = getenv("DC_DBG")) sa.dbg = atoi(dbg);
    // clip_nt CTAs per clip: one cluster (exchange through distributed shared memory, no global state), or -- long clips --
    // independent CTAs that exchange through L2 (flags zeroed per launch; relies on CTAs being dispatched in blockIdx order, so
    // that the lowest-numbered unfinished clip always has all of its CTAs resident)
    sa.nt = h->clip_nt;
    sa.rows_per = (h->T + h->clip_nt - 1) / h->clip_nt;
    sa.gx = h->clip_gx ? 1 : 0;
    if (h->clip_gx) {
        const uint32_t need = (uint32_t)n_steps * (uint32_t)L + 1u;          // tags gx_tag + 1 .. gx_tag + need - 1 are used by this launch
        if (h->gx_tag > 0xFFFFFFFFu - need - 1u) {                           // wrap-around: start over on cleared buffers
            DC_CUDA(h, cudaMemsetAsync(h->gx_part, 0, h->gx_cap * 2 * kKvPartFloats * sizeof(uint2), st));
            DC_CUDA(h, cudaMemsetAsync(h->gx_slice, 0, h->gx_cap_b * 2 * kD * 8 * sizeof(uint2), st));
            h->gx_tag = 0;
        }
        sa.gx_part = h->gx_part, sa.gx_slice = h->gx_slice, sa.gx_tag0 = h->gx_tag;
        h->gx_tag += need;
    }
    // two-tile clips: the partials are pushed into the peer's shared memory (DC_PUSH=0: pulled, as for clusters of 3 and 4)
    const char* pe = getenv("DC_PUSH");
    const bool push = !h->clip_gx && h->clip_nt == 2 && !(pe && pe[0] == '0');
    void (*kern)(StepArgs) = h->clip_gx ? (h->timeline_on ? (h->bf16 ? clip_kernel<true, true, true> : clip_kernel<false, true, true>)
                                                          : (h->bf16 ? clip_kernel<true, false, true> : clip_kernel<false, false, true>))
                             : push     ? (h->timeline_on ? (h->bf16 ? clip_kernel<true, true, false, true> : clip_kernel<false, true, false, true>)
                                                          : (h->bf16 ? clip_kernel<true, false, false, true> : clip_kernel<false, false, false, true>))
                                        : (h->timeline_on ? (h->bf16 ? clip_kernel<true, true> : clip_kernel<false, true>)
                                                          : (h->bf16 ? clip_kernel<true, false> : clip_kernel<false, false>));
    DC_CUDA(h, launch_kc(h->use_pdl, h->clip_gx ? 1 : h->clip_nt, kern, dim3((unsigned)(h->B * h->clip_nt)), dim3(kTileThreads), kClipSmemBytes, st, sa));
    h->launches++;
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

// One denoise step: A_emb + h0, L+1 tile launches with the time-axis reductions in between, output
// head (+ sampler update).  te/te_stride select the per-sample or per-step time embedding.
// `step` >= 0: the timestep is known on the host (persistent kernel: baked into the launch); -1: read from the device counter.
int enqueue_step(dc_handle* h, const float* x_in, const float* te, int te_stride, bool te_from_ctr, int mode, float* x_upd,
                 float* x0_out, const float* noise, int step, cudaStream_t st) {
    const int L = h->cfg.num_layers;
    const int M = h->M;
    const int blocks8 = (M + 7) / 8;
    auto mark = [&](int cls) {
        if (!h->prof) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        h->prof_ev.push_back(ev);
        h->prof_cls.push_back(cls);
    };
    mark(-1);
    if (h->persist && (step >= 0 || !te_from_ctr)) {
        // whole denoise step in ONE launch: A_emb + h0 prologue, all layers, output head + sampler update
        const int rc = enqueue_persistent(h, x_in, te, te_stride, te_from_ctr ? kE : 0, mode, x_upd, x0_out, 0, nullptr, noise, 0,
                                          te_from_ctr ? step : 0, 1, st);
        mark(1);
        return rc;
    }
    const int* ctr = te_from_ctr ? h->step_ctr : nullptr;
    DC_CUDA(h, launch_k(h->use_pdl, h->bf16 ? step_begin_kernel<true> : step_begin_kernel<false>, dim3(blocks8), dim3(128), 0, st, x_in,
                        (const float*)h->xp, te, ctr, te_stride, (const float*)h->WjT, (const float*)h->bj, (const float*)h->pos, M, h->T,
                        h->aemb, h->hbuf));
    h->launches++;
    mark(0);
    for (int l = -1; l < L; ++l) {
        const LayerArgs la = layer_args(h, l);
        const int rc = h->bf16 ? launch_layer<true>(h, la, h->tiles, st) : launch_layer<false>(h, la, h->tiles, st);
        if (rc) return rc;
        h->launches++;
        mark(1);
        if (l + 1 < L && !h->fuse_kv) {
            DC_CUDA(h, launch_k(h->use_pdl, h->bf16 ? kv_reduce_kernel<true> : kv_reduce_kernel<false>, dim3(h->B * kH), dim3(256), 0, st,
                                (const float*)h->kv, h->T, h->bd_sa, (size_t)kAworkBytes, (size_t)0, (size_t)0, 0));
            h->launches++;
            mark(2);
        }
    }
    DC_CUDA(h, launch_k(h->use_pdl, out_update_kernel, dim3(blocks8), dim3(256), 0, st, (const float*)h->hbuf, (const float*)h->WoT,
                        (const float*)h->bo, M, mode, (const float*)h->coef, (const int*)h->step_ctr, noise, x_upd, x0_out));
    h->launches++;
    mark(3);
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

}  // namespace

// =============================================================================================
extern "C" {

const char* dc_last_error(const dc_handle* h) { return h ? h->err.c_str() : g_last_error.c_str(); }

int dc_create(const dc_config* cfg, dc_handle** out) {
    if (!cfg || !out) return fail(nullptr, DC_ERR_INVALID, "dc_create: null argument");
    *out = nullptr;
    if (cfg->latent_dim != kD || cfg->num_heads != kH || cfg->ff_size != kF || cfg->input_feats != kP)
        return fail(nullptr, DC_ERR_UNSUPPORTED,
                    "dc_create: kernels are specialised for latent_dim=128, num_heads=8, ff_size=64, input_feats=26 "
                    "(got %d, %d, %d, %d)",
                    cfg->latent_dim, cfg->num_heads, cfg->ff_size, cfg->input_feats);
    if (cfg->num_layers < 1 || cfg->num_frames < 1) return fail(nullptr, DC_ERR_INVALID, "dc_create: bad num_layers / num_frames");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, DC_ERR_CUDA, "dc_create: CUDA device %d unavailable (%s); there is no CPU fallback", cfg->device,
                    e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range");
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, cfg->device);
    if (prop.major != 10) return fail(nullptr, DC_ERR_UNSUPPORTED, "dc_create: device is sm_%d%d, this library is sm_100a only", prop.major, prop.minor);
    dc_handle* h = new dc_handle();
    h->cfg = *cfg;
    h->bf16 = cfg->operand != DC_OPERAND_FP16;
    h->num_sms = prop.multiProcessorCount;
    // any failure below: the message is already in g_last_error (dc_last_error(NULL)); release what was created
    auto init = [&]() -> int {
        DC_CUDA(h, cudaSetDevice(cfg->device));
        if (int rc = init_kernel_attrs(h)) return rc;
        DC_CUDA(h, cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        DC_CUDA(h, cudaMalloc((void**)&h->step_ctr, 4));
        DC_CUDA(h, cudaMemset(h->step_ctr, 0, 4));
        return 0;
    };
    if (int rc = init()) {
        dc_destroy(h);
        return rc;
    }
    const char* mi = getenv("DC_MASK_INVERT");
    if (mi && mi[0] == '1') h->mask_invert = 1;
    if (mi && mi[0] == '3') h->mask_invert = 3;
    const char* pr = getenv("DC_PAIR");
    if (pr) h->use_pair = pr[0] == '1';
    const char* np = getenv("DC_PDL");
    if (np && np[0] == '1') h->use_pdl = true;
    const char* ng = getenv("DC_NO_GRAPH");
    if (ng && ng[0] == '1') h->use_graphs = false;
    *out = h;
    return 0;
}

void dc_destroy(dc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    drop_graph(h);
    free_workspace(h);
    void* ptrs[] = {h->wbuf, h->prm, h->prm_clip, h->wfuse, h->kshift, h->wout_img, h->wj_img, h->wkv, h->bkv, h->WjT, h->bj, h->pos, h->WoT, h->bo, h->WlinT, h->blin,
                    h->teW0, h->teb0, h->teW2, h->teb2, h->freqs, h->coef, h->te_table, h->step_ctr, h->timeline};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    void* mptrs[] = {h->me_w4t, h->me_b4, h->me_wpt, h->me_bp, h->me_buf0, h->me_buf1, h->me_wimg, h->me_bias};
    for (void* p : mptrs)
        if (p) cudaFree(p);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    delete h;
}

int dc_set_weight(dc_handle* h, const char* key, const void* data, const int64_t* shape, int ndim) {
    if (!h || !key || !data || ndim < 0 || ndim > 4) return fail(h, DC_ERR_INVALID, "dc_set_weight: bad argument");
    const std::string k(key);
    if (k.find("num_batches_tracked") != std::string::npos) return 0;              // BatchNorm bookkeeping, unused in eval mode
    HostTensor t;
    size_t n = 1;
    for (int i = 0; i < ndim; ++i) {
        t.shape.push_back(shape[i]);
        n *= (size_t)shape[i];
    }
    t.v.resize(n);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    DC_CUDA(h, cudaMemcpy(t.v.data(), data, n * 4, cudaMemcpyDefault));
    h->w[k] = std::move(t);
    h->finalized = false;
    return 0;
}

int dc_finalize_weights(dc_handle* h) {
    if (!h) return fail(h, DC_ERR_INVALID, "null handle");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    // a sampling loop may still be running on a (non-blocking) user stream: nothing below may race with it
    DC_CUDA(h, cudaDeviceSynchronize());
    const int L = h->cfg.num_layers;
    const bool bf = h->bf16;
#define GET(var, key, ...)                          \
    const HostTensor* var = get(h, key, {__VA_ARGS__}); \
    if (!var) return DC_ERR_INVALID;

    std::vector<uint8_t> wbuf((size_t)L * kLayerSlab, 0);
    std::vector<float> prm((size_t)L * kPrmFloats, 0.f);
    std::vector<uint8_t> wkv((size_t)L * 8 * 32768, 0);
    std::vector<uint8_t> wfuse((size_t)L * 16384, 0);
    std::vector<float> kshift((size_t)L * kD, 0.f);
    uint32_t static_mask = 0;
    std::vector<float> bkv((size_t)L * 256, 0.f);

    int film_rows[256];   // accumulator column n <- emb_layers.1 row: [scale 0..63 | shift 0..63 | scale 64..127 | shift 64..127]
    for (int n = 0; n < 256; ++n) film_rows[n] = n < 64 ? n : n < 128 ? n + 64 : n < 192 ? n - 64 : n;

    auto fold_bias = [](const HostTensor* W, const HostTensor* b, const HostTensor* beta, float* out) {
        const int N = (int)W->shape[0], K = (int)W->shape[1];
        for (int n = 0; n < N; ++n) {
            double acc = 0.0;
            for (int k = 0; k < K; ++k) acc += (double)W->v[(size_t)n * K + k] * beta->v[k];
            out[n] = b->v[n] + (float)acc;
        }
    };

    for (int l = 0; l < L; ++l) {
        const std::string p = "temporal_decoder_blocks." + std::to_string(l) + ".";
        uint8_t* slab = wbuf.data() + (size_t)l * kLayerSlab;
        float* pr = prm.data() + (size_t)l * kPrmFloats;

        auto stylization = [&](const std::string& sp, uint32_t off_we, uint32_t off_wo, int prm_off) -> int {
            GET(we, sp + "emb_layers.1.weight", 2 * kD, kE);
            GET(be, sp + "emb_layers.1.bias", 2 * kD);
            GET(g, sp + "norm.weight", kD);
            GET(bt, sp + "norm.bias", kD);
            GET(wo, sp + "out_layers.2.weight", kD, kD);
            GET(bo, sp + "out_layers.2.bias", kD);
            pack_image(slab + off_we, we->v.data(), kE, kE, film_rows, nullptr, 256, 8, bf);
            pack_image(slab + off_wo, wo->v.data(), kD, kD, nullptr, nullptr, kD, 2, bf);
            for (int n = 0; n < 256; ++n) pr[prm_off + kStBe + n] = be->v[film_rows[n]] + (film_rows[n] < kD ? 1.f : 0.f);
            for (int i = 0; i < kD; ++i) {
                pr[prm_off + kStG + i] = g->v[i];
                pr[prm_off + kStB + i] = bt->v[i];
                pr[prm_off + kStBo + i] = bo->v[i];
            }
            return 0;
        };

        // self-attention: LayerNorm affine folded into q/k/v
        {
            GET(g, p + "sa_block.norm.weight", kD);
            GET(bt, p + "sa_block.norm.bias", kD);
            const char* names[3] = {"query", "key", "value"};
            const uint32_t offs[3] = {kOffWq, kOffWk, kOffWv};
            const int poffs[3] = {kPrmSaBq, kPrmSaBk, kPrmSaBv};
            for (int i = 0; i < 3; ++i) {
                GET(W, p + "sa_block." + names[i] + ".weight", kD, kD);
                GET(b, p + "sa_block." + names[i] + ".bias", kD);
                pack_image(slab + offs[i], W->v.data(), kD, kD, nullptr, g->v.data(), kD, 2, bf);
                fold_bias(W, b, bt, pr + poffs[i]);
            }
            if (stylization(p + "sa_block.proj_out.", kOffWeSa, kOffWoSa, kPrmStSa)) return DC_ERR_INVALID;
            // Static shift of softmax_T(k) (transformer.py:111): k_d = w_d . n + b_d with n a LayerNorm output, so
            // ||n||_2 <= sqrt(128) and |k_d| <= ||w_d||_2 sqrt(128) + |b_d| =: c_d (1 % slack for the 16-bit operand
            // rounding).  exp(k - c) then never overflows and stays >= exp(-2c): if every c_d of the layer is small
            // the kernel uses c instead of the per-tile column max (softmax is shift-invariant).  Masked frames
            // (k - 1e6) still underflow to exactly 0, as in the reference.
            GET(Wk, p + "sa_block.key.weight", kD, kD);
            float cmax = 0.f;
            for (int d = 0; d < kD; ++d) {
                double n2 = 0.0;
                for (int k = 0; k < kD; ++k) {
                    const double w = (double)Wk->v[(size_t)d * kD + k] * g->v[k];
                    n2 += w * w;
                }
                const float c = (float)(1.01 * std::sqrt(n2) * std::sqrt((double)kD) + std::fabs((double)pr[kPrmSaBk + d]) + 1e-3);
                kshift[(size_t)l * kD + d] = c;
                cmax = std::max(cmax, c);
            }
            const char* ss = getenv("DC_STATIC_SHIFT");       // "0": always use the running max (tests)
            const float limit = bf ? 30.f : 4.f;              // E must stay a normal number of the 16-bit operand type
            if (cmax <= limit && !(ss && ss[0] == '0')) static_mask |= 1u << l;
        }
        // cross-attention: query side per step, key/value side step-invariant (text_norm folded)
        {
            GET(g, p + "ca_block.norm.weight", kD);
            GET(bt, p + "ca_block.norm.bias", kD);
            GET(Wq, p + "ca_block.query.weight", kD, kD);
            GET(bq, p + "ca_block.query.bias", kD);
            pack_image(slab + kOffWqCa, Wq->v.data(), kD, kD, nullptr, g->v.data(), kD, 2, bf);
            fold_bias(Wq, bq, bt, pr + kPrmCaBq);
            GET(tg, p + "ca_block.text_norm.weight", kE);
            GET(tb, p + "ca_block.text_norm.bias", kE);
            GET(Wk, p + "ca_block.key.weight", kD, kE);
            GET(bk, p + "ca_block.key.bias", kD);
            GET(Wv, p + "ca_block.value.weight", kD, kE);
            GET(bv, p + "ca_block.value.bias", kD);
            std::vector<float> Wcat((size_t)256 * kE);
            memcpy(Wcat.data(), Wk->v.data(), (size_t)kD * kE * 4);
            memcpy(Wcat.data() + (size_t)kD * kE, Wv->v.data(), (size_t)kD * kE * 4);
            pack_image(wkv.data() + (size_t)l * 8 * 32768, Wcat.data(), kE, kE, nullptr, tg->v.data(), 256, 8, bf);
            fold_bias(Wk, bk, tb, bkv.data() + (size_t)l * 256);
            fold_bias(Wv, bv, tb, bkv.data() + (size_t)l * 256 + kD);
            if (stylization(p + "ca_block.proj_out.", kOffWeCa, kOffWoCa, kPrmStCa)) return DC_ERR_INVALID;
        }
        // FFN
        {
            GET(W1, p + "ffn.linear1.weight", kF, kD);
            GET(b1, p + "ffn.linear1.bias", kF);
            GET(W2, p + "ffn.linear2.weight", kD, kF);
            GET(b2, p + "ffn.linear2.bias", kD);
            pack_image(slab + kOffW1, W1->v.data(), kD, kD, nullptr, nullptr, kF, 2, bf);
            pack_image(slab + kOffW2, W2->v.data(), kF, kF, nullptr, nullptr, kD, 1, bf);
            for (int i = 0; i < kF; ++i) pr[kPrmFfB1 + i] = b1->v[i];
            for (int i = 0; i < kD; ++i) pr[kPrmFfB2 + i] = b2->v[i];
            if (stylization(p + "ffn.proj_out.", kOffWeFf, kOffWoFf, kPrmStFf)) return DC_ERR_INVALID;
            // Persistent kernel: the FFN has no pre-norm (transformer.py:170-173), so its up-projection is linear in the
            // residual add that precedes it:  (h + a Wo^T + bo) W1^T = h W1^T + a (W1 Wo)^T + W1 bo.  The kernel issues
            // both GEMMs at once and skips a round trip; here: the product matrix and the folded biases.
            GET(wo, p + "ca_block.proj_out.out_layers.2.weight", kD, kD);
            GET(bo, p + "ca_block.proj_out.out_layers.2.bias", kD);
            std::vector<float> W1c((size_t)kF * kD);
            for (int n = 0; n < kF; ++n)
                for (int k = 0; k < kD; ++k) {
                    double acc = 0.0;
                    for (int j = 0; j < kD; ++j) acc += (double)W1->v[(size_t)n * kD + j] * wo->v[(size_t)j * kD + k];
                    W1c[(size_t)n * kD + k] = (float)acc;
                }
            pack_image(wfuse.data() + (size_t)l * 16384, W1c.data(), kD, kD, nullptr, nullptr, kF, 2, bf);
        }
    }
    // parameter block variant of the persistent kernel: b1 <- b1 + W1 bo_ca ; bo_ffn <- bo_ffn + bo_ca (added once, at the
    // end of the layer, instead of right after the Wo_ca GEMM)
    std::vector<float> prm_clip = prm;
    for (int l = 0; l < L; ++l) {
        const std::string p = "temporal_decoder_blocks." + std::to_string(l) + ".";
        GET(W1, p + "ffn.linear1.weight", kF, kD);
        float* pc = prm_clip.data() + (size_t)l * kPrmFloats;
        const float* bo_ca = prm.data() + (size_t)l * kPrmFloats + kPrmStCa + kStBo;
        for (int n = 0; n < kF; ++n) {
            double acc = 0.0;
            for (int j = 0; j < kD; ++j) acc += (double)W1->v[(size_t)n * kD + j] * bo_ca[j];
            pc[kPrmFfB1 + n] += (float)acc;
        }
        for (int i = 0; i < kD; ++i) pc[kPrmStFf + kStBo + i] += bo_ca[i];
    }
    if (upload(h, &h->wbuf, wbuf.data(), wbuf.size())) return DC_ERR_CUDA;
    if (upload(h, &h->prm, prm.data(), prm.size() * 4)) return DC_ERR_CUDA;
    if (upload(h, &h->prm_clip, prm_clip.data(), prm_clip.size() * 4)) return DC_ERR_CUDA;
    if (upload(h, &h->wfuse, wfuse.data(), wfuse.size())) return DC_ERR_CUDA;
    if (upload(h, &h->kshift, kshift.data(), kshift.size() * 4)) return DC_ERR_CUDA;
    h->static_mask = L <= 32 ? static_mask : 0;
    if (upload(h, &h->wkv, wkv.data(), wkv.size())) return DC_ERR_CUDA;
    if (upload(h, &h->bkv, bkv.data(), bkv.size() * 4)) return DC_ERR_CUDA;

    // embeddings, output head, conditioning projection, time MLP
    {
        GET(Wj, "joint_embed.weight", kD, kP);
        GET(bj, "joint_embed.bias", kD);
        GET(pos, "sequence_embedding", h->cfg.num_frames, kD);
        GET(Wo, "out.weight", kP, kD);
        GET(bo, "out.bias", kP);
        GET(Wl, "linear.weight", kE, kMusic);
        GET(bl, "linear.bias", kE);
        GET(W0, "time_embed.0.weight", kE, kD);
        GET(b0, "time_embed.0.bias", kE);
        GET(W2, "time_embed.2.weight", kE, kE);
        GET(b2, "time_embed.2.bias", kE);
        std::vector<float> WjT((size_t)kP * kD), WoT((size_t)kD * 32, 0.f), WlT((size_t)kMusic * kE), bo32(32, 0.f);
        for (int j = 0; j < kD; ++j)
            for (int c = 0; c < kP; ++c) WjT[(size_t)c * kD + j] = Wj->v[(size_t)j * kP + c];
        for (int pI = 0; pI < kP; ++pI)
            for (int j = 0; j < kD; ++j) WoT[(size_t)j * 32 + pI] = Wo->v[(size_t)pI * kD + j];
        for (int pI = 0; pI < kP; ++pI) bo32[pI] = bo->v[pI];
        auto split16 = [&](float v, float& hi, float& lo) {                 // v = hi + lo, both exactly representable in the operand type
            uint16_t u = to16(v, bf);
            if (bf) {
                __nv_bfloat16 b;
                memcpy(&b, &u, 2);
                hi = __bfloat162float(b);
            } else {
                __half hh;
                memcpy(&hh, &u, 2);
                hi = __half2float(hh);
            }
            lo = v - hi;
        };
        {   // joint_embed [128 x 128]: k 0..25 = W_hi, 32..57 = W_hi (against x_lo), 64..89 = W_lo (against x_hi again)
            std::vector<float> wj2((size_t)kD * 128, 0.f);
            for (int n = 0; n < kD; ++n)
                for (int k = 0; k < kP; ++k) {
                    float hi, lo;
                    split16(Wj->v[(size_t)n * kP + k], hi, lo);
                    wj2[(size_t)n * 128 + k] = wj2[(size_t)n * 128 + 32 + k] = hi;
                    wj2[(size_t)n * 128 + 64 + k] = lo;
                }
            std::vector<uint8_t> img((size_t)2 * kD * 128, 0);
            pack_image(img.data(), wj2.data(), 128, 128, nullptr, nullptr, kD, 2, bf);
            if (upload(h, &h->wj_img, img.data(), img.size())) return DC_ERR_CUDA;
        }
        {   // output head [32 x 384]: k-blocks 0,1 = W_hi (against h_hi), 2,3 = W_hi (against h_lo), 4,5 = W_lo (against h_hi again)
            std::vector<float> whi((size_t)32 * kD, 0.f), wlo((size_t)32 * kD, 0.f);
            for (int pI = 0; pI < kP; ++pI)
                for (int j = 0; j < kD; ++j) split16(Wo->v[(size_t)pI * kD + j], whi[(size_t)pI * kD + j], wlo[(size_t)pI * kD + j]);
            std::vector<uint8_t> img(6 * 32 * 128, 0);
            pack_image(img.data(), whi.data(), kD, kD, nullptr, nullptr, 32, 2, bf);
            memcpy(img.data() + 2 * 32 * 128, img.data(), 2 * 32 * 128);
            pack_image(img.data() + 4 * 32 * 128, wlo.data(), kD, kD, nullptr, nullptr, 32, 2, bf);
            if (upload(h, &h->wout_img, img.data(), img.size())) return DC_ERR_CUDA;
        }
        for (int e = 0; e < kE; ++e)
            for (int c = 0; c < kMusic; ++c) WlT[(size_t)c * kE + e] = Wl->v[(size_t)e * kMusic + c];
        std::vector<float> freqs(kD / 2);
        auto itf = h->w.find("aux.timestep_freqs");
        if (itf != h->w.end() && itf->second.v.size() == (size_t)kD / 2) {
            freqs = itf->second.v;
        } else {
            const float nl = -logf(10000.f);
            for (int i = 0; i < kD / 2; ++i) freqs[i] = expf(nl * (float)i / (float)(kD / 2));
        }
        if (upload(h, &h->WjT, WjT.data(), WjT.size() * 4) || upload(h, &h->bj, bj->v.data(), kD * 4) ||
            upload(h, &h->pos, pos->v.data(), pos->v.size() * 4) || upload(h, &h->WoT, WoT.data(), WoT.size() * 4) ||
            upload(h, &h->bo, bo32.data(), 32 * 4) || upload(h, &h->WlinT, WlT.data(), WlT.size() * 4) ||
            upload(h, &h->blin, bl->v.data(), kE * 4) || upload(h, &h->teW0, W0->v.data(), W0->v.size() * 4) ||
            upload(h, &h->teb0, b0->v.data(), kE * 4) || upload(h, &h->teW2, W2->v.data(), W2->v.size() * 4) ||
            upload(h, &h->teb2, b2->v.data(), kE * 4) || upload(h, &h->freqs, freqs.data(), freqs.size() * 4))
            return DC_ERR_CUDA;
    }
    // music encoder front-end (optional: present whenever a full MotionTransformer state_dict was uploaded)
    h->has_music = false;
    if (h->w.count("music_encoder.conv1.0.conv2d_layer.0.weight")) {
        const float eps = 1e-5f;       // torch BatchNorm default (reference transformer.py:298,304,327)
        struct Spec { const char* name; int cin, cout; };
        const Spec specs[7] = {{"conv1.0", 1, 16}, {"conv1.1", 16, 16}, {"conv1.2", 16, 16}, {"conv2.0", 16, 32},
                               {"conv2.1", 32, 32}, {"conv3.0", 32, 32}, {"conv3.1", 32, 32}};
        auto bn_scale = [&](const std::string& p, int n, std::vector<float>& sc, std::vector<float>& sh) -> int {
            GET(g, p + ".weight", n);
            GET(bt, p + ".bias", n);
            GET(mu, p + ".running_mean", n);
            GET(var, p + ".running_var", n);
            sc.resize(n), sh.resize(n);
            for (int i = 0; i < n; ++i) {
                sc[i] = g->v[i] / std::sqrt(var->v[i] + eps);
                sh[i] = bt->v[i] - mu->v[i] * sc[i];
            }
            return 0;
        };
        // BatchNorm-folded weights of layer li as fp32 [co][ci][9] (+ 1x1 residual [co][ci]) and biases
        auto folded = [&](int li, std::vector<float>& wf, std::vector<float>& bf_, std::vector<float>* w1f, std::vector<float>* b1f) -> int {
            const std::string p = std::string("music_encoder.") + specs[li].name;
            const int ci = specs[li].cin, co = specs[li].cout;
            GET(w, p + ".conv2d_layer.0.weight", co, ci, 3, 3);
            GET(b, p + ".conv2d_layer.0.bias", co);
            std::vector<float> sc, sh;
            if (bn_scale(p + ".conv2d_layer.1", co, sc, sh)) return DC_ERR_INVALID;
            wf.resize((size_t)co * ci * 9), bf_.resize(co);
            for (int o = 0; o < co; ++o) {
                bf_[o] = b->v[o] * sc[o] + sh[o];
                for (int i = 0; i < ci * 9; ++i) wf[(size_t)o * ci * 9 + i] = w->v[(size_t)o * ci * 9 + i] * sc[o];
            }
            if (w1f) {                     // 1x1 residual convolution + BatchNorm (conv2.0)
                GET(rw, p + ".residual.0.weight", co, ci, 1, 1);
                GET(rb, p + ".residual.0.bias", co);
                std::vector<float> rs, rh;
                if (bn_scale(p + ".residual.1", co, rs, rh)) return DC_ERR_INVALID;
                w1f->resize((size_t)co * ci), b1f->resize(co);
                for (int o = 0; o < co; ++o) {
                    (*b1f)[o] = rb->v[o] * rs[o] + rh[o];
                    for (int i = 0; i < ci; ++i) (*w1f)[(size_t)o * ci + i] = rw->v[(size_t)o * ci + i] * rs[o];
                }
            }
            return 0;
        };
        {
            std::vector<float> wf, bf_;
            if (folded(0, wf, bf_, nullptr, nullptr)) return DC_ERR_INVALID;
            for (int t = 0; t < 9; ++t)
                for (int q = 0; q < 16; ++q) h->me_c10.w[t * 16 + q] = wf[(size_t)q * 9 + t];
            for (int q = 0; q < 16; ++q) h->me_c10.b[q] = bf_[q];
        }
        // tensor-core layers: per window row one B block of (hi, lo)-split weights, see music_encoder_tc.cuh
        auto hi_of = [](float v) {
            const uint16_t u = to16(v, true);
            uint32_t w32 = (uint32_t)u << 16;
            float f;
            memcpy(&f, &w32, 4);
            return f;
        };
        std::vector<uint8_t> wimg;
        std::vector<float> biases(6 * 64, 0.f);
        for (int li = 1; li < 7; ++li) {
            const int ci = specs[li].cin, co = specs[li].cout;
            const bool res = li == 3;
            std::vector<float> wf, bf_, w1f, b1f;
            if (folded(li, wf, bf_, res ? &w1f : nullptr, res ? &b1f : nullptr)) return DC_ERR_INVALID;
            h->me_woff[li - 1] = wimg.size();
            // one block per window row dy: rows dxb * co + c = [w_hi (ci) | w_lo (ci)] of tap (dy, dxb); rows 3 co + c: 1x1 residual (centre row)
            const int wrows = (res ? 4 : 3) * co;
            for (int dy = 0; dy < 3; ++dy) {
                std::vector<float> m1((size_t)wrows * 64, 0.f);
                auto put = [&](int row, int i, float wv) {
                    const float whi = hi_of(wv);
                    m1[(size_t)row * 64 + i] = whi, m1[(size_t)row * 64 + ci + i] = wv - whi;
                };
                for (int dxb = 0; dxb < 3; ++dxb)
                    for (int o = 0; o < co; ++o)
                        for (int i = 0; i < ci; ++i) put(dxb * co + o, i, wf[((size_t)o * ci + i) * 9 + dy * 3 + dxb]);
                if (res && dy == 1)
                    for (int o = 0; o < co; ++o)
                        for (int i = 0; i < ci; ++i) put(3 * co + o, i, w1f[(size_t)o * ci + i]);
                const size_t off = wimg.size();
                wimg.resize(off + (size_t)wrows * 128);
                pack_image(wimg.data() + off, m1.data(), 64, 64, nullptr, nullptr, wrows, 1, true);
            }
            for (int o = 0; o < co; ++o) biases[(li - 1) * 64 + o] = bf_[o];
            if (res)
                for (int o = 0; o < co; ++o) biases[(li - 1) * 64 + co + o] = b1f[o];
        }
        if (upload(h, &h->me_wimg, wimg.data(), wimg.size()) || upload(h, &h->me_bias, biases.data(), biases.size() * 4)) return DC_ERR_CUDA;
        GET(w4, "music_encoder.conv4.0.weight", kMusic, 512, 1);
        GET(b4, "music_encoder.conv4.0.bias", kMusic);
        std::vector<float> sc4, sh4;
        if (bn_scale("music_encoder.conv4.1", kMusic, sc4, sh4)) return DC_ERR_INVALID;
        GET(wp, "proj.weight", kMusic, kMusic);
        GET(bp, "proj.bias", kMusic);
        std::vector<float> w4t((size_t)512 * kMusic), b4f(kMusic), wpt((size_t)kMusic * kMusic);
        for (int o = 0; o < kMusic; ++o) {
            b4f[o] = b4->v[o] * sc4[o] + sh4[o];
            for (int k = 0; k < 512; ++k) w4t[(size_t)k * kMusic + o] = w4->v[(size_t)o * 512 + k] * sc4[o];
            for (int k = 0; k < kMusic; ++k) wpt[(size_t)k * kMusic + o] = wp->v[(size_t)o * kMusic + k];
        }
        if (upload(h, &h->me_w4t, w4t.data(), w4t.size() * 4) || upload(h, &h->me_b4, b4f.data(), b4f.size() * 4) ||
            upload(h, &h->me_wpt, wpt.data(), wpt.size() * 4) || upload(h, &h->me_bp, bp->v.data(), kMusic * 4))
            return DC_ERR_CUDA;
        h->has_music = true;
    }
#undef GET
    h->finalized = true;
    h->prepared = false;
    drop_graph(h);
    // a re-finalize after a schedule was set must refresh the time-embedding table
    if (h->S > 0) {
        time_embed_kernel<<<h->S, kE>>>(nullptr, 0, h->freqs, h->teW0, h->teb0, h->teW2, h->teb2, h->te_table);
        DC_CUDA(h, cudaDeviceSynchronize());
    }
    return 0;
}

int dc_set_schedule(dc_handle* h, int num_steps, const float* coef) {
    if (!h || num_steps < 1 || !coef) return fail(h, DC_ERR_INVALID, "dc_set_schedule: bad argument");
    if (!h->finalized) return fail(h, DC_ERR_STATE, "dc_set_schedule: call dc_finalize_weights first");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    DC_CUDA(h, cudaDeviceSynchronize());       // coef / te_table may still be read by a loop on a non-blocking user stream
    drop_graph(h);
    if (upload(h, &h->coef, coef, (size_t)num_steps * 8 * 4)) return DC_ERR_CUDA;
    if (h->te_table) cudaFree(h->te_table);
    h->te_table = nullptr;
    DC_CUDA(h, cudaMalloc((void**)&h->te_table, (size_t)num_steps * kE * 4));
    h->S = num_steps;
    time_embed_kernel<<<num_steps, kE>>>(nullptr, 0, h->freqs, h->teW0, h->teb0, h->teW2, h->teb2, h->te_table);
    h->launches++;
    DC_CUDA(h, cudaGetLastError());
    DC_CUDA(h, cudaDeviceSynchronize());
    return 0;
}

int dc_prepare_cond(dc_handle* h, const float* xf_proj, const float* xf_out, const int64_t* length, int B, int T, void* stream) {
    if (!h || !xf_proj || !xf_out || B < 1 || T < 1) return fail(h, DC_ERR_INVALID, "dc_prepare_cond: bad argument");
    if (!h->finalized) return fail(h, DC_ERR_STATE, "dc_prepare_cond: weights not finalized");
    if (T > h->cfg.num_frames)
        return fail(h, DC_ERR_INVALID, "dc_prepare_cond: T=%d exceeds num_frames=%d (rows of sequence_embedding)", T, h->cfg.num_frames);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_workspace(h, B, T)) return rc;
    if (h->B != B || h->T != T) drop_graph(h);
    h->B = B, h->T = T, h->M = B * T, h->tiles = (h->M + kTileRows - 1) / kTileRows;
    {
        // Fusing the time-axis reduction into the layer kernel pays when a clip spans many tiles (measured r01:
        // T = 1800 -> 1.34x faster loop; T = 180 -> 7 % slower than the stand-alone kv_reduce kernel).
        const char* nf = getenv("DC_FUSE_KV");      // "0" / "1" force the choice
        const bool fuse = T >= kTileRows && (nf ? nf[0] == '1' : T >= 4 * kTileRows);
        // Cluster-per-clip persistent kernel: any batch size, T <= 16 tiles; clusters of more than 8 CTAs need the opt-in
        // and a GPC with that many free SMs (checked with the occupancy query).  DC_PERSIST=0 forces the per-layer path.
        const char* pe = getenv("DC_PERSIST");
        const int nt = (T + kTileRows - 1) / kTileRows;
        bool persist = nt <= kMaxClipTiles && !h->use_pair && !(pe && pe[0] == '0');
        // Large clusters leave SMs idle (measured: 15 resident clusters of 8, 7 of 15 or 16 -- 120 / 105 of 148 SMs): clips of 7 tiles and
        // more run as independent CTAs that exchange their partials through L2.  Measured crossover (tools/exchange_crossover.py, motion-s/s
        // cluster vs L2): 5 tiles 52.7 k vs 51.0 k, 6 tiles 52.1 k vs 50.0 k, 7 tiles 44.5 k vs 57.3 k, 8 tiles 47.3 k vs 54.0 k.
        // DC_GX=0 / 1 forces the choice.
        const char* gxe = getenv("DC_GX");
        const bool gx = persist && nt > 1 && (gxe ? gxe[0] == '1' : nt >= kGxMinTiles);
        if (gx && ((size_t)B * nt > h->gx_cap || (size_t)B > h->gx_cap_b)) {
            if (h->gx_part) cudaFree(h->gx_part);
            if (h->gx_slice) cudaFree(h->gx_slice);
            h->gx_part = nullptr, h->gx_slice = nullptr, h->gx_cap = 0, h->gx_cap_b = 0;
            DC_CUDA(h, cudaMalloc((void**)&h->gx_part, (size_t)B * 2 * nt * kKvPartFloats * sizeof(uint2)));
            DC_CUDA(h, cudaMalloc((void**)&h->gx_slice, (size_t)B * 2 * kD * 8 * sizeof(uint2)));
            // tag 0 is never handed out: cleared buffers hold no valid word (the layout depends on nt, so a resize starts over too)
            DC_CUDA(h, cudaMemsetAsync(h->gx_part, 0, (size_t)B * 2 * nt * kKvPartFloats * sizeof(uint2), st));
            DC_CUDA(h, cudaMemsetAsync(h->gx_slice, 0, (size_t)B * 2 * kD * 8 * sizeof(uint2), st));
            h->gx_tag = 0;
            h->gx_cap = (size_t)B * nt, h->gx_cap_b = (size_t)B;
        }
        h->clip_gx = gx;
        if (persist && nt > 1 && !gx && nt != h->clip_nt_checked) {
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)nt), cfg.blockDim = dim3(kTileThreads), cfg.dynamicSmemBytes = kClipSmemBytes;
            cudaLaunchAttribute at{};
            at.id = cudaLaunchAttributeClusterDimension;
            at.val.clusterDim.x = (unsigned)nt, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
            cfg.attrs = &at, cfg.numAttrs = 1;
            int nclusters = 0;
            const cudaError_t qe = h->bf16 ? cudaOccupancyMaxActiveClusters(&nclusters, clip_kernel<true, false>, &cfg)
                                           : cudaOccupancyMaxActiveClusters(&nclusters, clip_kernel<false, false>, &cfg);
            if (qe != cudaSuccess || nclusters < 1) {
                // never drop to the (1.4x slower) per-layer path silently: the caller asks for it with DC_PERSIST=0
                cudaGetLastError();
                return fail(h, DC_ERR_UNSUPPORTED,
                            "dc_prepare_cond: this device cannot co-schedule a cluster of %d CTAs (one per 128-frame tile of a %d-frame clip): "
                            "cudaOccupancyMaxActiveClusters -> %s, %d clusters.  Set DC_PERSIST=0 to run the per-layer launch path instead.",
                            nt, T, cudaGetErrorString(qe), nclusters);
            } else {
                h->clip_nt_checked = nt;
            }
        }
        if (fuse != h->fuse_kv || persist != h->persist) drop_graph(h);
        // the two paths keep the cross-attention images in different layouts (compact head blocks / block-diagonal with zero
        // off-diagonal blocks that are never rewritten): start from zeros when the path changes
        if (persist != h->persist) DC_CUDA(h, cudaMemsetAsync(h->bd_ca, 0, h->cap_B * (size_t)h->cfg.num_layers * kAworkBytes, st));
        h->fuse_kv = fuse;
        h->persist = persist;
        h->clip_nt = nt;
        DC_CUDA(h, cudaMemsetAsync(h->clip_cnt, 0, (size_t)B * 4, st));
    }
    bool masked = false;
    if (length) {
        for (int i = 0; i < B; ++i) {
            if (length[i] < 0) return fail(h, DC_ERR_INVALID, "dc_prepare_cond: negative length");
            masked |= length[i] < T;
        }
    }
    if (masked != h->has_length) drop_graph(h);
    h->has_length = masked;
    if (masked) DC_CUDA(h, cudaMemcpyAsync(h->length, length, (size_t)B * 8, cudaMemcpyHostToDevice, st));
    const int L = h->cfg.num_layers;
    const int blocks4 = (h->M + 3) / 4;
    if (h->bf16)
        cond_prep_kernel<true><<<blocks4, 128, 0, st>>>(xf_proj, xf_out, h->WlinT, h->blin, h->M, h->xp, h->zimg);
    else
        cond_prep_kernel<false><<<blocks4, 128, 0, st>>>(xf_proj, xf_out, h->WlinT, h->blin, h->M, h->xp, h->zimg);
    h->launches++;
    // cross-attention K | V projections and their time-axis softmax + K^T V (step-invariant): all layers of a chunk in ONE launch
    // each (blockIdx.y = layer) -- 16 launches of ~11 us were 0.18 ms of every host-to-host call
    for (int l0 = 0; l0 < L; l0 += h->kv_layers) {
        const int nl = std::min(h->kv_layers, L - l0);
        GemmRowsArgs ga{};
        ga.a_img = h->zimg;
        ga.w_img = h->wkv + (size_t)l0 * 8 * 32768;
        ga.bias = h->bkv + (size_t)l0 * 256;
        ga.out = h->kv;
        ga.M = h->M, ga.N = 256, ga.kblocks = 8, ga.ldo = 256, ga.blocked = 1;
        ga.w_layer_stride = (size_t)8 * 32768, ga.bias_layer_stride = 256, ga.out_layer_stride = h->kv_stride;
        const int rc = h->bf16 ? launch_gemm_rows<true>(h, ga, h->tiles, st, nl) : launch_gemm_rows<false>(h, ga, h->tiles, st, nl);
        if (rc) return rc;
        // persistent kernel: compact 4 KB head-block images (16.8 -> 2.1 MB on C2: the working set is re-streamed out of L2 every
        // step); per-layer path: the [128 x 128] block-diagonal B operand its GEMM expects
        const size_t lstride = h->persist ? (size_t)kBdcBytes : (size_t)kAworkBytes;
        uint8_t* bd = h->bd_ca + (size_t)l0 * lstride;
        if (h->bf16)
            kv_reduce_kernel<true><<<dim3((unsigned)(B * kH), (unsigned)nl), 256, 0, st>>>(h->kv, T, bd, (size_t)L * lstride, h->kv_stride, lstride, h->persist ? 1 : 0);
        else
            kv_reduce_kernel<false><<<dim3((unsigned)(B * kH), (unsigned)nl), 256, 0, st>>>(h->kv, T, bd, (size_t)L * lstride, h->kv_stride, lstride, h->persist ? 1 : 0);
        h->launches += 2;
    }
    DC_CUDA(h, cudaGetLastError());
    h->prepared = true;
    return 0;
}

int dc_forward(dc_handle* h, const float* x, const int64_t* timesteps, float* out, void* stream) {
    if (!h || !x || !timesteps || !out) return fail(h, DC_ERR_INVALID, "dc_forward: bad argument");
    if (!h->prepared) return fail(h, DC_ERR_STATE, "dc_forward: call dc_prepare_cond first");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    time_embed_kernel<<<h->B, kE, 0, st>>>((const long long*)timesteps, 0, h->freqs, h->teW0, h->teb0, h->teW2, h->teb2, h->te_b);
    h->launches++;
    return enqueue_step(h, x, h->te_b, kE, false, DC_SAMPLER_NONE, nullptr, out, nullptr, -1, st);
}

static int check_sampling(dc_handle* h, int sampler, const char* who) {
    if (!h) return fail(h, DC_ERR_INVALID, "%s: null handle", who);
    if ((sampler & 0xF) != DC_SAMPLER_DDIM && (sampler & 0xF) != DC_SAMPLER_DDPM)
        return fail(h, DC_ERR_INVALID, "%s: unknown sampler %d", who, sampler);
    if (!h->prepared) return fail(h, DC_ERR_STATE, "%s: call dc_prepare_cond first", who);
    if (h->S < 1) return fail(h, DC_ERR_STATE, "%s: call dc_set_schedule first", who);
    return 0;
}

int dc_sample_step(dc_handle* h, int sampler, float* x, float* pred_x0, int step, const float* noise, void* stream) {
    if (int rc = check_sampling(h, sampler, "dc_sample_step")) return rc;
    if (!x || !pred_x0 || step < 0 || step >= h->S) return fail(h, DC_ERR_INVALID, "dc_sample_step: bad argument (step=%d, S=%d)", step, h->S);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    set_step_kernel<<<1, 1, 0, st>>>(h->step_ctr, step, 0);
    h->launches++;
    return enqueue_step(h, x, h->te_table, 0, true, sampler, x, pred_x0, noise, step, st);
}

int dc_sampler_update(dc_handle* h, int sampler, float* x, const float* pred_x0, int step, const float* noise, int64_t n, void* stream) {
    if (!h || !x || !pred_x0 || n < 0) return fail(h, DC_ERR_INVALID, "dc_sampler_update: bad argument");
    if ((sampler & 0xF) != DC_SAMPLER_DDIM && (sampler & 0xF) != DC_SAMPLER_DDPM)
        return fail(h, DC_ERR_INVALID, "dc_sampler_update: unknown sampler");
    if (h->S < 1 || step < 0 || step >= h->S) return fail(h, DC_ERR_STATE, "dc_sampler_update: schedule not set or step out of range");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    if (n == 0) return 0;
    sampler_update_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(pred_x0, (size_t)n, sampler, h->coef, step, noise, x);
    h->launches++;
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

int dc_sample_range(dc_handle* h, int sampler, int step0, int n_steps, float* x, const float* step_noise, float* trace_x0, float* trace_x,
                    void* stream) {
    if (int rc = check_sampling(h, sampler, "dc_sample_range")) return rc;
    if (!x) return fail(h, DC_ERR_INVALID, "dc_sample_range: null x");
    if (step0 < 0 || step0 >= h->S || n_steps < 1 || n_steps > step0 + 1)
        return fail(h, DC_ERR_INVALID, "dc_sample_range: steps %d .. %d are outside the schedule of %d steps", step0, step0 - n_steps + 1, h->S);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)h->M * kP;
    const bool traced = trace_x0 || trace_x || step_noise;
    const bool whole = step0 == h->S - 1 && n_steps == h->S;

    DC_CUDA(h, cudaMemcpyAsync(h->xwork, x, n * 4, cudaMemcpyDeviceToDevice, st));
    set_step_kernel<<<1, 1, 0, st>>>(h->step_ctr, step0, 0);
    h->launches++;
    const int64_t per_step = (h->fuse_kv ? 1 : 2) * (int64_t)h->cfg.num_layers + 4;
    if (h->persist) {
        // the whole block of steps is ONE launch of the persistent kernel (noise / trace slices are strides inside it)
        if (int rc = enqueue_persistent(h, h->xwork, h->te_table, 0, kE, sampler, h->xwork, trace_x0 ? trace_x0 : h->x0work, trace_x0 ? n : 0,
                                        trace_x, step_noise, step_noise ? n : 0, step0, n_steps, st))
            return rc;
    } else if (h->use_graphs && !traced && whole) {
        // per-layer path: the loop body reads its step index from device memory, so one captured 5-step graph is replayed S/5
        // times; per-step noise / trace slices are host pointer arithmetic, which forces the un-captured path below
        GraphKey key;
        key.sampler = sampler;
        key.steps = (n_steps % 5 == 0) ? 5 : 1;
        if (!h->gexec || !(h->gkey == key)) {
            drop_graph(h);
            cudaGraph_t graph = nullptr;
            DC_CUDA(h, cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
            const int64_t before = h->launches;
            int rc = 0;
            for (int i = 0; i < key.steps && !rc; ++i) {
                rc = enqueue_step(h, h->xwork, h->te_table, 0, true, sampler, h->xwork, h->x0work, nullptr, -1, h->cap_stream);
                launch_k(h->use_pdl, set_step_kernel, dim3(1), dim3(1), 0, h->cap_stream, h->step_ctr, 0, -1);
            }
            h->launches = before;
            cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
            if (rc) {
                if (graph) cudaGraphDestroy(graph);
                return rc;
            }
            DC_CUDA(h, ce);
            ce = cudaGraphInstantiate(&h->gexec, graph, 0);
            cudaGraphDestroy(graph);
            DC_CUDA(h, ce);
            h->gkey = key;
        }
        for (int i = 0; i < n_steps / key.steps; ++i) DC_CUDA(h, cudaGraphLaunch(h->gexec, st));
        h->launches += (int64_t)n_steps * per_step;
    } else {
        for (int i = 0; i < n_steps; ++i) {
            const float* nz = step_noise ? step_noise + (size_t)i * n : nullptr;
            float* x0dst = trace_x0 ? trace_x0 + (size_t)i * n : h->x0work;
            if (int rc = enqueue_step(h, h->xwork, h->te_table, 0, true, sampler, h->xwork, x0dst, nz, step0 - i, st)) return rc;
            if (trace_x) DC_CUDA(h, cudaMemcpyAsync(trace_x + (size_t)i * n, h->xwork, n * 4, cudaMemcpyDeviceToDevice, st));
            set_step_kernel<<<1, 1, 0, st>>>(h->step_ctr, 0, -1);
            h->launches++;
        }
    }
    DC_CUDA(h, cudaMemcpyAsync(x, h->xwork, n * 4, cudaMemcpyDeviceToDevice, st));
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

int dc_sample_loop(dc_handle* h, int sampler, int num_steps, float* x, const float* step_noise, float* trace_x0, float* trace_x, void* stream) {
    if (int rc = check_sampling(h, sampler, "dc_sample_loop")) return rc;
    if (num_steps != h->S)
        return fail(h, DC_ERR_STATE, "dc_sample_loop: caller expects %d steps but the schedule set with dc_set_schedule has %d "
                    "(the noise / trace buffers are sized by the caller)", num_steps, h->S);
    return dc_sample_range(h, sampler, h->S - 1, h->S, x, step_noise, trace_x0, trace_x, stream);
}

int dc_generate_host(dc_handle* h, int sampler, const float* xf_proj, const float* xf_out, const int64_t* length, const float* noise,
                     float* motion_out, int B, int T, void* stream) {
    if (!h || !xf_proj || !xf_out || !noise || !motion_out || B < 1 || T < 1) return fail(h, DC_ERR_INVALID, "dc_generate_host: bad argument");
    if (!h->finalized) return fail(h, DC_ERR_STATE, "dc_generate_host: weights not finalized");
    if (h->S < 1) return fail(h, DC_ERR_STATE, "dc_generate_host: call dc_set_schedule first");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    if (int rc = ensure_workspace(h, B, T)) return rc;
    const size_t M = (size_t)B * T;
    DC_CUDA(h, cudaMemcpyAsync(h->in_proj, xf_proj, M * kMusic * 4, cudaMemcpyHostToDevice, st));
    DC_CUDA(h, cudaMemcpyAsync(h->in_out, xf_out, M * kMusic * 4, cudaMemcpyHostToDevice, st));
    if (int rc = dc_prepare_cond(h, h->in_proj, h->in_out, length, B, T, stream)) return rc;
    // x0work doubles as the device-side motion buffer of this call
    float* xdev = h->x0work;
    DC_CUDA(h, cudaMemcpyAsync(xdev, noise, M * kP * 4, cudaMemcpyHostToDevice, st));
    // dc_sample_loop copies x -> xwork first, so aliasing x0work as the in/out buffer is safe: the
    // final copy back happens after the last step has written its pred_xstart.
    if (int rc = dc_sample_loop(h, sampler, h->S, xdev, nullptr, nullptr, nullptr, stream)) return rc;
    DC_CUDA(h, cudaMemcpyAsync(motion_out, xdev, M * kP * 4, cudaMemcpyDeviceToHost, st));
    DC_CUDA(h, cudaStreamSynchronize(st));
    return 0;
}

int dc_profile_step(dc_handle* h, int sampler, float* x, int step, float* ms_out, int* count_out, void* stream) {
    if (int rc = check_sampling(h, sampler, "dc_profile_step")) return rc;
    if (!x || !ms_out || !count_out || step < 0 || step >= h->S) return fail(h, DC_ERR_INVALID, "dc_profile_step: bad argument");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    set_step_kernel<<<1, 1, 0, st>>>(h->step_ctr, step, 0);
    h->launches++;
    h->prof = true;
    h->prof_ev.clear();
    h->prof_cls.clear();
    const int rc = enqueue_step(h, x, h->te_table, 0, true, sampler, x, h->x0work, nullptr, step, st);
    h->prof = false;
    cudaError_t e = cudaStreamSynchronize(st);
    for (int i = 0; i < 4; ++i) ms_out[i] = 0.f, count_out[i] = 0;
    for (size_t i = 1; i < h->prof_ev.size(); ++i) {
        float ms = 0.f;
        if (e == cudaSuccess) cudaEventElapsedTime(&ms, h->prof_ev[i - 1], h->prof_ev[i]);
        const int c = h->prof_cls[i];
        if (c >= 0 && c < 4) ms_out[c] += ms, count_out[c]++;
    }
    for (cudaEvent_t ev : h->prof_ev) cudaEventDestroy(ev);
    h->prof_ev.clear();
    h->prof_cls.clear();
    if (rc) return rc;
    DC_CUDA(h, e);
    return 0;
}

int dc_debug_timeline(dc_handle* h, float* x, int step, unsigned long long* out, int max_launches) {
    if (int rc = check_sampling(h, DC_SAMPLER_DDIM, "dc_debug_timeline")) return rc;
    if (!x || !out || max_launches < 1) return fail(h, DC_ERR_INVALID, "dc_debug_timeline: bad argument");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    const int n = std::min(max_launches, h->cfg.num_layers + 1);
    // per-layer path: [launch][512] u64; persistent kernel: 3 lanes of (1 + 2 * kTlEvents) u64 (see Timeline in clip_kernel.cuh)
    const size_t words = std::max((size_t)(h->cfg.num_layers + 1) * 512, (size_t)3 * (1 + 2 * kTlEvents));
    if (!h->timeline) DC_CUDA(h, cudaMalloc((void**)&h->timeline, words * 8));
    DC_CUDA(h, cudaMemset(h->timeline, 0, words * 8));
    set_step_kernel<<<1, 1>>>(h->step_ctr, step, 0);
    h->timeline_on = true;
    const int rc = enqueue_step(h, x, h->te_table, 0, true, DC_SAMPLER_DDIM, x, h->x0work, nullptr, step, 0);
    h->timeline_on = false;
    if (rc) return rc;
    DC_CUDA(h, cudaDeviceSynchronize());
    DC_CUDA(h, cudaMemcpy(out, h->timeline, std::min((size_t)max_launches * 512, words) * 8, cudaMemcpyDeviceToHost));
    (void)n;
    return 0;
}

int dc_encode_music(dc_handle* h, const float* mel, float* xf_proj, float* xf_out, int B, int Tm, void* stream) {
    if (!h || !mel || !xf_proj || !xf_out || B < 1 || Tm < 1) return fail(h, DC_ERR_INVALID, "dc_encode_music: bad argument");
    if (!h->finalized || !h->has_music) return fail(h, DC_ERR_STATE, "dc_encode_music: music_encoder.* / proj.* weights not loaded");
    // MaxPool2d((5,5), stride (3,2), pad 2) maps Tm mel frames to (Tm - 1) / 3 + 1 motion frames; the 3x3 reflect padding
    // needs at least 2 rows at every stage (torch raises otherwise)
    const int T = (Tm - 1) / 3 + 1;
    if (Tm < 2 || T < 2) return fail(h, DC_ERR_INVALID, "dc_encode_music: %d mel frames are too few for the reflect-padded convolutions", Tm);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t st = (cudaStream_t)stream;
    constexpr int kBins = 128;
    const size_t per_clip = (size_t)16 * Tm * kBins;                  // floats of the largest activation (16 x Tm x 128 == 32 x Tm x 64)
    // clips per pass: <= 1 GiB per activation buffer, and 32 planes per clip must fit gridDim.z of the pooling kernels
    const int chunk = (int)std::max<size_t>(1, std::min<size_t>(std::min<size_t>((size_t)B, 2047), ((size_t)1 << 28) / per_clip));
    if ((size_t)chunk * per_clip > h->me_cap) {
        if (h->me_buf0) cudaFree(h->me_buf0);
        if (h->me_buf1) cudaFree(h->me_buf1);
        h->me_buf0 = h->me_buf1 = nullptr, h->me_cap = 0;
        DC_CUDA(h, cudaMalloc((void**)&h->me_buf0, (size_t)chunk * per_clip * 4));
        DC_CUDA(h, cudaMalloc((void**)&h->me_buf1, (size_t)chunk * per_clip * 4));
        h->me_cap = (size_t)chunk * per_clip;
    }
    uint16_t *p0 = reinterpret_cast<uint16_t*>(h->me_buf0), *p1 = reinterpret_cast<uint16_t*>(h->me_buf1);   // split pixels [C hi | C lo]
    const bool me_pdl = !(getenv("DC_ME_PDL") && atoi(getenv("DC_ME_PDL")) == 0);
    auto conv = [&](auto kern, int smem, int occ, int NM, int W, const uint16_t* src, uint16_t* dst, int H, int nb, int li) {
        const int bands = (H * (W + 2) + NM * kMeTile - 1) / (NM * kMeTile), jobs = bands * nb;   // bands of NM accumulator tiles each
        // programmatic dependent launch: the CTAs of this convolution start on the SMs that the previous kernel's last blocks free, and
        // run their prologue (weight blocks, strip zero-fill, TMEM allocation, barriers) while its tail drains; the kernel calls
        // griddepcontrol.wait before it touches an activation
        launch_k(me_pdl, kern, dim3((unsigned)std::min(jobs, h->num_sms * occ)), dim3(kMeThreads), (size_t)smem, st, src, dst, H, bands, jobs,
                 (const uint8_t*)(h->me_wimg + h->me_woff[li - 1]), (const float*)(h->me_bias + (li - 1) * 64));
    };
    auto blocks256 = [](long n) { return (unsigned)((n + 255) / 256); };
    // output rows per thread of the sliding-window pools.  A segment costs `startup` = KH - SH extra input rows, and the grid runs in
    // waves of num_sms x (resident blocks per SM): pick the segment length that minimises waves x (rows fetched per thread) -- with a
    // fixed length the C2 batch ran 1.11 waves, i.e. a second wave at 11 % occupancy.
    auto pool = [&](auto kern, const uint16_t* src, uint16_t* dst, int H, int W, int Ho, int Wo, int nb, int groups, int startup) {
        int regs = 64;
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, (const void*)kern) == cudaSuccess) regs = fa.numRegs;
        const long per_wave = (long)h->num_sms * std::max(1, std::min(16, 65536 / (((regs + 7) & ~7) * 128)));
        int seg = Ho;
        double best = 1e30;
        for (int c = 4; c <= std::min(Ho, 96); ++c) {
            const long blocks = ((long)nb * ((Ho + c - 1) / c) * Wo * groups + 127) / 128;
            const double cost = (double)((blocks + per_wave - 1) / per_wave) * (c + startup);
            if (cost < best) best = cost, seg = c;
        }
        if (getenv("DC_POOL_SEG")) seg = std::max(1, atoi(getenv("DC_POOL_SEG")));
        const int nseg = (Ho + seg - 1) / seg;
        const long items = (long)nb * nseg * Wo * groups;
        launch_k(me_pdl, kern, dim3((unsigned)((items + 127) / 128)), dim3(128), (size_t)0, st, src, dst, H, W, Ho, Wo, seg, nseg, items);
    };
    // the row-staged pools: one block of 128 threads per (clip, segment of output rows); segment length by the same wave rule
    auto pool_staged = [&](auto kern, const uint16_t* src, uint16_t* dst, int H, int Ho, int nb, int startup, int sh) {
        int regs = 64;
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, (const void*)kern) == cudaSuccess) regs = fa.numRegs;
        const long per_wave = (long)h->num_sms * std::max(1, std::min(6, 65536 / (((regs + 7) & ~7) * 128)));   // (32 KB of shared memory per block: <= 6)
        int seg = Ho;
        double best = 1e30;
        for (int c = 2; c <= std::min(Ho, 96); ++c) {
            const long blocks = (long)nb * ((Ho + c - 1) / c);
            const double cost = (double)((blocks + per_wave - 1) / per_wave) * (c * sh + startup);
            if (cost < best) best = cost, seg = c;
        }
        if (getenv("DC_POOL_SEG")) seg = std::max(1, atoi(getenv("DC_POOL_SEG")));
        const int nseg = (Ho + seg - 1) / seg;
        launch_k(me_pdl, kern, dim3((unsigned)(nb * nseg)), dim3(128), (size_t)0, st, src, dst, H, Ho, seg, nseg);
    };
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        const float* m0 = mel + (size_t)b0 * Tm * kBins;              // (nb, 1, Tm, 128)
        int H = Tm, W = kBins, Ho, Wo;
        launch_k(me_pdl, conv10_split_kernel, dim3(blocks256((long)H * W), (unsigned)nb), dim3(256), (size_t)0, st, m0, p0, H, W, h->me_c10);
        conv(conv_tc_kernel<16, 16, 1, 128, kMeNM128>, me_smem_bytes<16, 16, 1, 128, kMeNM128>(), h->me_occ[0], kMeNM128, W, p0, p1, H, nb, 1);
        conv(conv_tc_kernel<16, 16, 1, 128, kMeNM128>, me_smem_bytes<16, 16, 1, 128, kMeNM128>(), h->me_occ[0], kMeNM128, W, p1, p0, H, nb, 2);
        Ho = (H + 4 - 5) / 1 + 1, Wo = (W + 4 - 5) / 2 + 1;           // MaxPool2d((5,5), stride (1,2), padding 2)
        pool_staged(maxpool_staged_kernel<16, 5, 5, 1, 2, 2, 2, 128>, p0, p1, H, Ho, nb, 4, 1);
        H = Ho, W = Wo;
        conv(conv_tc_kernel<16, 32, 2, 64, kMeNM64>, me_smem_bytes<16, 32, 2, 64, kMeNM64>(), h->me_occ[1], kMeNM64, W, p1, p0, H, nb, 3);
        conv(conv_tc_kernel<32, 32, 1, 64, kMeNM64b>, me_smem_bytes<32, 32, 1, 64, kMeNM64b>(), h->me_occ[2], kMeNM64b, W, p0, p1, H, nb, 4);
        Ho = (H + 4 - 5) / 3 + 1, Wo = (W + 4 - 5) / 2 + 1;           // MaxPool2d((5,5), stride (3,2), padding 2)
        pool_staged(maxpool_staged_kernel<32, 5, 5, 3, 2, 2, 2, 64>, p1, p0, H, Ho, nb, 2, 3);
        H = Ho, W = Wo;
        conv(conv_tc_kernel<32, 32, 1, 32, kMeNM32>, me_smem_bytes<32, 32, 1, 32, kMeNM32>(), h->me_occ[3], kMeNM32, W, p0, p1, H, nb, 5);
        conv(conv_tc_kernel<32, 32, 1, 32, kMeNM32>, me_smem_bytes<32, 32, 1, 32, kMeNM32>(), h->me_occ[3], kMeNM32, W, p1, p0, H, nb, 6);
        Ho = (H + 2 - 3) / 1 + 1, Wo = (W + 2 - 3) / 2 + 1;           // MaxPool2d((3,3), stride (1,2), padding 1)
        pool(maxpool_split_kernel<32, 3, 3, 1, 2, 1, 1>, p0, p1, H, W, Ho, Wo, nb, 4, 2);
        H = Ho, W = Wo;                                               // (nb, T, 16, 32 split)
        if (H != T || W != 16) return fail(h, DC_ERR_INVALID, "dc_encode_music: unexpected feature map %d x %d", H, W);
        const long M = (long)nb * T;
        launch_k(me_pdl, conv4_proj_split_kernel, dim3((unsigned)((M + kC4Rows - 1) / kC4Rows)), dim3(256), (size_t)0, st, (const uint16_t*)p1,
                 (const float*)h->me_w4t, (const float*)h->me_b4, (const float*)h->me_wpt, (const float*)h->me_bp, xf_out + (size_t)b0 * T * kMusic,
                 xf_proj + (size_t)b0 * T * kMusic, M);
        h->launches += 11;
#ifdef DC_ME_TIMELINE
        if (getenv("DC_ME_TIMELINE")) {                               // debug build only (tools/experiments/me_tl_report.py)
            static unsigned long long tl[4 * 1024];
            cudaStreamSynchronize(st);
            cudaMemcpyFromSymbol(tl, me_tl, sizeof(tl));
            for (int k = 0; k < 4; ++k) {
                fprintf(stderr, "[me_tl] kernel %d\n", k);
                unsigned long long t0 = tl[k * 1024] & 0xFFFFFFFFFFFFFFull;
                for (int i = 0; i < 160 && tl[k * 1024 + i]; ++i)
                    fprintf(stderr, "  %d %llu\n", (int)(tl[k * 1024 + i] >> 56), (tl[k * 1024 + i] & 0xFFFFFFFFFFFFFFull) - t0);
            }
        }
#endif
    }
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

int dc_smooth_motion(int device, const float* motion, float* out, int B, int T, int C, int window, const float* fir, const float* edge,
                     float scale, void* stream) {
    if (!motion || !out || !fir || !edge || B < 1 || T < 1 || C < 1) return fail(nullptr, DC_ERR_INVALID, "dc_smooth_motion: bad argument");
    if (window < 3 || !(window & 1) || window > kSavgolMaxWindow)
        return fail(nullptr, DC_ERR_INVALID, "dc_smooth_motion: window must be odd, 3 <= window <= %d", kSavgolMaxWindow);
    if (T < window) return fail(nullptr, DC_ERR_INVALID, "dc_smooth_motion: window=%d exceeds the %d frames (scipy mode='interp' requires window <= size)", window, T);
    DC_CUDA(nullptr, cudaSetDevice(device));
    SavgolCoef cf{};
    cf.window = window;
    cf.scale = scale;
    for (int j = 0; j < window; ++j) cf.fir[j] = fir[j];
    for (int i = 0; i < window / 2; ++i)
        for (int j = 0; j < window; ++j) cf.edge[i][j] = edge[(size_t)i * window + j];
    const long n = (long)B * T * C;
    savgol_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(motion, out, B, T, C, cf);
    DC_CUDA(nullptr, cudaGetLastError());
    return 0;
}

int dc_time_embedding(dc_handle* h, const int64_t* timesteps, int n, float* out, void* stream) {
    if (!h || !timesteps || !out || n < 1) return fail(h, DC_ERR_INVALID, "dc_time_embedding: bad argument");
    if (!h->finalized) return fail(h, DC_ERR_STATE, "dc_time_embedding: weights not finalized");
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    time_embed_kernel<<<n, kE, 0, (cudaStream_t)stream>>>((const long long*)timesteps, 0, h->freqs, h->teW0, h->teb0, h->teW2, h->teb2, out);
    h->launches++;
    DC_CUDA(h, cudaGetLastError());
    return 0;
}

int dc_cluster_occupancy(dc_handle* h, int tiles_per_clip, int* max_clusters) {
    if (!h || !max_clusters || tiles_per_clip < 1 || tiles_per_clip > kMaxClipTiles)
        return fail(h, DC_ERR_INVALID, "dc_cluster_occupancy: need 1 <= tiles_per_clip <= %d", kMaxClipTiles);
    DC_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)tiles_per_clip), cfg.blockDim = dim3(kTileThreads), cfg.dynamicSmemBytes = kClipSmemBytes;
    cudaLaunchAttribute at{};
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = (unsigned)tiles_per_clip, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cfg.attrs = &at, cfg.numAttrs = 1;
    int n = 0;
    DC_CUDA(h, h->bf16 ? cudaOccupancyMaxActiveClusters(&n, clip_kernel<true, false>, &cfg)
                       : cudaOccupancyMaxActiveClusters(&n, clip_kernel<false, false>, &cfg));
    *max_clusters = n;
    return 0;
}

int64_t dc_kernel_launches(const dc_handle* h) { return h ? h->launches : 0; }

int dc_set_graphs(dc_handle* h, int enabled) {
    if (!h) return fail(h, DC_ERR_INVALID, "null handle");
    h->use_graphs = enabled != 0;
    return 0;
}

int dc_selftest_gemm(int device, int operand, int M, int N, int K, const float* A, const float* W, const float* bias, float* out) {
    if (!A || !W || !out || M < 1 || N < 16 || N > 256 || N % 16 || K < 64 || K % 64)
        return fail(nullptr, DC_ERR_INVALID, "dc_selftest_gemm: need K%%64==0, N%%16==0, 16<=N<=256");
    DC_CUDA(nullptr, cudaSetDevice(device));
    if (int rc = init_kernel_attrs(nullptr)) return rc;
    const bool bf = operand != DC_OPERAND_FP16;
    const int tiles = (M + kTileRows - 1) / kTileRows, kb = K / 64;
    std::vector<uint8_t> aimg((size_t)tiles * kb * kABlockBytes, 0), wimg((size_t)kb * N * 128, 0);
    for (int t = 0; t < tiles; ++t) {
        const int rows = std::min(kTileRows, M - t * kTileRows);
        for (int b = 0; b < kb; ++b) {
            std::vector<uint8_t> blk(kABlockBytes, 0);
            // one k-block of one tile: reuse pack_image with a 64-wide window
            std::vector<int> rm(kTileRows);
            for (int r = 0; r < kTileRows; ++r) rm[r] = r < rows ? t * kTileRows + r : -1;
            pack_image(blk.data(), A + (size_t)b * 64, K, 64, rm.data(), nullptr, kTileRows, 1, bf);
            memcpy(aimg.data() + ((size_t)t * kb + b) * kABlockBytes, blk.data(), kABlockBytes);
        }
    }
    pack_image(wimg.data(), W, K, K, nullptr, nullptr, N, kb, bf);
    uint8_t *da = nullptr, *dw = nullptr;
    float *db = nullptr, *dout = nullptr;
    DC_CUDA(nullptr, cudaMalloc((void**)&da, aimg.size()));
    DC_CUDA(nullptr, cudaMalloc((void**)&dw, wimg.size()));
    DC_CUDA(nullptr, cudaMalloc((void**)&dout, (size_t)M * N * 4));
    DC_CUDA(nullptr, cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
    DC_CUDA(nullptr, cudaMemcpy(dw, wimg.data(), wimg.size(), cudaMemcpyHostToDevice));
    if (bias) {
        DC_CUDA(nullptr, cudaMalloc((void**)&db, (size_t)N * 4));
        DC_CUDA(nullptr, cudaMemcpy(db, bias, (size_t)N * 4, cudaMemcpyHostToDevice));
    }
    GemmRowsArgs ga{};
    ga.a_img = da, ga.w_img = dw, ga.bias = db, ga.out = dout, ga.M = M, ga.N = N, ga.kblocks = kb, ga.ldo = N;
    const int rc = bf ? launch_gemm_rows<true>(nullptr, ga, tiles, 0) : launch_gemm_rows<false>(nullptr, ga, tiles, 0);
    if (rc) return rc;
    DC_CUDA(nullptr, cudaGetLastError());
    DC_CUDA(nullptr, cudaDeviceSynchronize());
    DC_CUDA(nullptr, cudaMemcpy(out, dout, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
    cudaFree(da), cudaFree(dw), cudaFree(dout);
    if (db) cudaFree(db);
    return 0;
}

// =============================================================================================
// Evaluation features: ST-GCN motion encoder + metric reductions, see eval_kernels.cuh
// =============================================================================================
struct dc_eval {
    int device = 0;
    bool finalized = false;
    std::map<std::string, std::vector<float>> w;          // host copies by state_dict key
    float* pool = nullptr;                                // one device allocation for all folded parameters
    dc::StgcnLayer layer[dc::kEvLayers];
    const float* dbn = nullptr;                           // [2][26]
    const float* wfc = nullptr;                           // [416][64]
    const float* bfc = nullptr;                           // [64]
    float* act[2] = {nullptr, nullptr};                   // [N * T][13][32] ping-pong
    size_t act_rows = 0;
};

static int efail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}
#define EV_CUDA(expr)                                                                                            \
    do {                                                                                                         \
        cudaError_t e_ = (expr);                                                                                 \
        if (e_ != cudaSuccess) return efail(DC_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_));               \
    } while (0)

int dc_eval_create(int device, dc_eval** out) {
    if (!out) return efail(DC_ERR_INVALID, "dc_eval_create: null argument");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return efail(DC_ERR_CUDA, "dc_eval_create: CUDA device %d unavailable; there is no CPU fallback", device);
    EV_CUDA(cudaSetDevice(device));
    EV_CUDA(cudaFuncSetAttribute(dc::stgcn_layer_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::stgcn_smem_bytes(32)));
    EV_CUDA(cudaFuncSetAttribute(dc::stgcn_layer_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, dc::stgcn_smem_bytes(2)));
    dc_eval* e = new dc_eval();
    e->device = device;
    *out = e;
    return 0;
}

void dc_eval_destroy(dc_eval* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->pool) cudaFree(e->pool);
    if (e->act[0]) cudaFree(e->act[0]);
    if (e->act[1]) cudaFree(e->act[1]);
    delete e;
}

int dc_eval_set_weight(dc_eval* e, const char* key, const float* data, int64_t count) {
    if (!e || !key || !data || count < 0) return efail(DC_ERR_INVALID, "dc_eval_set_weight: bad argument");
    std::vector<float> v((size_t)count);
    EV_CUDA(cudaSetDevice(e->device));
    EV_CUDA(cudaMemcpy(v.data(), data, (size_t)count * 4, cudaMemcpyDefault));
    e->w[key] = std::move(v);
    e->finalized = false;
    return 0;
}

int dc_eval_finalize(dc_eval* e) {
    if (!e) return efail(DC_ERR_INVALID, "dc_eval_finalize: null handle");
    using namespace dc;
    auto get = [&](const std::string& k, size_t n) -> const float* {
        auto it = e->w.find(k);
        if (it == e->w.end()) {
            efail(DC_ERR_INVALID, "dc_eval_finalize: missing weight %s", k.c_str());
            return nullptr;
        }
        if (it->second.size() != n) {
            efail(DC_ERR_INVALID, "dc_eval_finalize: weight %s has %zu elements, expected %zu", k.c_str(), it->second.size(), n);
            return nullptr;
        }
        return it->second.data();
    };
    // eval-mode BatchNorm -> (scale, shift) in fp64 (torch: (x - mean) / sqrt(var + 1e-5) * weight + bias)
    auto bn = [&](const std::string& p, int c, std::vector<double>& sc, std::vector<double>& sh) -> bool {
        const float *g = get(p + ".weight", c), *b = get(p + ".bias", c), *m = get(p + ".running_mean", c), *v = get(p + ".running_var", c);
        if (!g || !b || !m || !v) return false;
        sc.resize(c), sh.resize(c);
        for (int i = 0; i < c; ++i) {
            sc[i] = (double)g[i] / std::sqrt((double)v[i] + 1e-5);
            sh[i] = (double)b[i] - (double)m[i] * sc[i];
        }
        return true;
    };
    std::vector<float> pool;
    auto push = [&](const std::vector<double>& v) -> size_t {
        const size_t off = pool.size();
        for (double x : v) pool.push_back((float)x);
        while (pool.size() % 4) pool.push_back(0.f);
        return off;
    };
    const float* A = get("st_gcn.A", kEvV * kEvV);
    if (!A) return DC_ERR_INVALID;
    size_t off[kEvLayers][5];
    std::vector<double> sc, sh;
    if (!bn("st_gcn.data_bn", 2 * kEvV, sc, sh)) return DC_ERR_INVALID;
    std::vector<double> dbn(4 * kEvV);
    for (int i = 0; i < 2 * kEvV; ++i) dbn[i] = sc[i], dbn[2 * kEvV + i] = sh[i];      // channel index = joint * 2 + coordinate
    const size_t off_dbn = push(dbn);
    for (int l = 0; l < kEvLayers; ++l) {
        const int cin = l == 0 ? 2 : kEvC;
        const std::string p = "st_gcn.st_gcn_networks." + std::to_string(l);
        const float *wg = get(p + ".gcn.conv.weight", (size_t)kEvC * cin), *bg = get(p + ".gcn.conv.bias", kEvC);
        const float *wt = get(p + ".tcn.2.weight", (size_t)kEvC * kEvC * 3), *bt = get(p + ".tcn.2.bias", kEvC);
        const float* imp = get("st_gcn.edge_importance." + std::to_string(l), kEvV * kEvV);
        std::vector<double> s1, h1, s2, h2;
        if (!wg || !bg || !wt || !bt || !imp || !bn(p + ".tcn.0", kEvC, s1, h1) || !bn(p + ".tcn.3", kEvC, s2, h2)) return DC_ERR_INVALID;
        std::vector<double> fwg((size_t)cin * kEvC), ae(kEvV * kEvV), cg(kEvV * kEvC), fwt(3 * kEvC * kEvC), ct(kEvC);
        for (int c = 0; c < kEvC; ++c)
            for (int ci = 0; ci < cin; ++ci) fwg[(size_t)ci * kEvC + c] = s1[c] * (double)wg[c * cin + ci];
        for (int i = 0; i < kEvV * kEvV; ++i) ae[i] = (double)((float)A[i] * (float)imp[i]);          // fp32 product, as self.A * importance
        for (int w = 0; w < kEvV; ++w) {
            double cs = 0.0;
            for (int v = 0; v < kEvV; ++v) cs += ae[v * kEvV + w];
            for (int c = 0; c < kEvC; ++c) cg[w * kEvC + c] = s1[c] * (double)bg[c] * cs + h1[c];
        }
        for (int co = 0; co < kEvC; ++co) {
            for (int ci = 0; ci < kEvC; ++ci)
                for (int dt = 0; dt < 3; ++dt) fwt[((size_t)dt * kEvC + ci) * kEvC + co] = s2[co] * (double)wt[(co * kEvC + ci) * 3 + dt];
            ct[co] = s2[co] * (double)bt[co] + h2[co];
        }
        off[l][0] = push(fwg), off[l][1] = push(ae), off[l][2] = push(cg), off[l][3] = push(fwt), off[l][4] = push(ct);
    }
    const float *wf = get("fc.0.weight", (size_t)kEvLat * kEvV * kEvC), *bf_ = get("fc.0.bias", kEvLat);
    if (!wf || !bf_ || !bn("fc.1", kEvLat, sc, sh)) return DC_ERR_INVALID;
    std::vector<double> fw((size_t)kEvV * kEvC * kEvLat), fb(kEvLat);
    for (int o = 0; o < kEvLat; ++o) {
        for (int c = 0; c < kEvC; ++c)
            for (int v = 0; v < kEvV; ++v) fw[((size_t)v * kEvC + c) * kEvLat + o] = sc[o] * (double)wf[o * (kEvV * kEvC) + c * kEvV + v];   // flatten order: c * 13 + v
        fb[o] = sc[o] * (double)bf_[o] + sh[o];
    }
    const size_t off_fw = push(fw), off_fb = push(fb);
    EV_CUDA(cudaSetDevice(e->device));
    EV_CUDA(cudaDeviceSynchronize());
    if (e->pool) cudaFree(e->pool);
    e->pool = nullptr;
    EV_CUDA(cudaMalloc((void**)&e->pool, pool.size() * 4));
    EV_CUDA(cudaMemcpy(e->pool, pool.data(), pool.size() * 4, cudaMemcpyHostToDevice));
    e->dbn = e->pool + off_dbn, e->wfc = e->pool + off_fw, e->bfc = e->pool + off_fb;
    for (int l = 0; l < kEvLayers; ++l)
        e->layer[l] = StgcnLayer{e->pool + off[l][0], e->pool + off[l][1], e->pool + off[l][2], e->pool + off[l][3], e->pool + off[l][4]};
    e->finalized = true;
    return 0;
}

int dc_eval_motion_features(dc_eval* e, const float* motion, float* feat, int N, int T, void* stream) {
    if (!e || !motion || !feat || N < 1 || T < 1) return efail(DC_ERR_INVALID, "dc_eval_motion_features: bad argument");
    if (!e->finalized) return efail(DC_ERR_STATE, "dc_eval_motion_features: call dc_eval_finalize first");
    using namespace dc;
    EV_CUDA(cudaSetDevice(e->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t rows = (size_t)N * T;
    if (rows > e->act_rows) {
        if (e->act[0]) cudaFree(e->act[0]);
        if (e->act[1]) cudaFree(e->act[1]);
        e->act[0] = e->act[1] = nullptr, e->act_rows = 0;
        EV_CUDA(cudaMalloc((void**)&e->act[0], rows * kEvV * kEvC * 4));
        EV_CUDA(cudaMalloc((void**)&e->act[1], rows * kEvV * kEvC * 4));
        e->act_rows = rows;
    }
    const dim3 grid((unsigned)((T + kEvTT - 1) / kEvTT), (unsigned)N);
    stgcn_layer_kernel<2><<<grid, 256, stgcn_smem_bytes(2), st>>>(motion, e->act[0], T, e->layer[0], e->dbn, 0);
    for (int l = 1; l < kEvLayers; ++l)
        stgcn_layer_kernel<32><<<grid, 256, stgcn_smem_bytes(32), st>>>(e->act[(l - 1) & 1], e->act[l & 1], T, e->layer[l], nullptr, 1);
    stgcn_fc_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(e->act[(kEvLayers - 1) & 1], e->wfc, e->bfc, feat, (long)rows);
    EV_CUDA(cudaGetLastError());
    return 0;
}

int dc_eval_feature_stats(int device, const float* feat, int64_t rows, double* sum, double* m2, void* stream) {
    if (!feat || !sum || !m2 || rows < 1) return efail(DC_ERR_INVALID, "dc_eval_feature_stats: bad argument");
    EV_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    EV_CUDA(cudaMemsetAsync(sum, 0, dc::kEvLat * 8, st));
    EV_CUDA(cudaMemsetAsync(m2, 0, dc::kEvLat * dc::kEvLat * 8, st));
    const unsigned blocks = (unsigned)std::min<int64_t>(296, (rows + 31) / 32);
    dc::feat_colsum_kernel<<<blocks, 256, 0, st>>>(feat, (long)rows, sum);
    dc::feat_cov_kernel<<<blocks, 256, 0, st>>>(feat, (long)rows, sum, m2);
    EV_CUDA(cudaGetLastError());
    return 0;
}

int dc_eval_feature_l1(int device, const float* a, const float* b, int64_t rows, double* out, void* stream) {
    if (!a || !b || !out || rows < 1) return efail(DC_ERR_INVALID, "dc_eval_feature_l1: bad argument");
    EV_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    EV_CUDA(cudaMemsetAsync(out, 0, 8, st));
    const int64_t n = rows * dc::kEvLat;
    dc::feat_l1_kernel<<<(unsigned)std::min<int64_t>(592, (n + 255) / 256), 256, 0, st>>>(a, b, (long)n, out);
    EV_CUDA(cudaGetLastError());
    return 0;
}

int dc_eval_motion_beats(int device, const float* motion, float* envelope, uint8_t* beats, int N, int T, int order, void* stream) {
    if (!motion || !envelope || !beats || N < 1 || T < 1 || order < 1) return efail(DC_ERR_INVALID, "dc_eval_motion_beats: bad argument");
    EV_CUDA(cudaSetDevice(device));
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned blocks = (unsigned)(((long)N * T + 255) / 256);
    dc::motion_envelope_kernel<<<blocks, 256, 0, st>>>(motion, envelope, N, T);
    dc::local_minima_kernel<<<blocks, 256, 0, st>>>(envelope, beats, N, T, order);
    EV_CUDA(cudaGetLastError());
    return 0;
}

int dc_eval_beat_alignment(int device, const uint8_t* music_beats, int Tm, const uint8_t* motion_beats, int T, int N, float sigma, float* scores,
                           void* stream) {
    if (!music_beats || !motion_beats || !scores || N < 1 || T < 1 || Tm < 1 || !(sigma > 0.f))
        return efail(DC_ERR_INVALID, "dc_eval_beat_alignment: bad argument");
    EV_CUDA(cudaSetDevice(device));
    dc::beat_alignment_kernel<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(music_beats, Tm, motion_beats, T, sigma, scores);
    EV_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"
