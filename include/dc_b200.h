/* dc_b200.h -- C ABI of the B200-native Diffusion-Conductor denoising path.
 *
 * The reference (viiika/Diffusion-Conductor) is pure Python and has no FFI: the boundary it
 * exposes for this path is the call surface used by DDPMTrainer.generate_music_motion
 * (Diffusion_Stage/trainers/ddpm_trainer.py:183-201).  Each entry point below names the reference
 * interface it stands in for; the Python shim in diffusion_conductor_b200/ binds them with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - every function returns 0 on success or a negative dc_status; the message is available from
 *     dc_last_error(handle) (or dc_last_error(NULL) when no handle exists yet).  Nothing throws,
 *     nothing calls exit().
 *   - tensors are plain pointers + sizes, fp32 row-major contiguous unless stated; "dev" pointers
 *     live on the handle's CUDA device and are BORROWED for the duration of the call only.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously on it.
 *   - a handle is not thread-safe: one handle per GPU per process.
 *   - there is no CPU fallback: without a CUDA device dc_create fails.
 */
#ifndef DC_B200_H
#define DC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dc_handle dc_handle;

typedef enum {
    DC_OK = 0,
    DC_ERR_INVALID = -1,      /* bad argument / shape / missing weight */
    DC_ERR_UNSUPPORTED = -2,  /* configuration outside the supported fast path */
    DC_ERR_CUDA = -3,         /* CUDA runtime error (message carries cudaGetErrorString) */
    DC_ERR_STATE = -4         /* call order violated (e.g. sampling before dc_prepare_cond) */
} dc_status;

typedef enum { DC_OPERAND_BF16 = 0, DC_OPERAND_FP16 = 1 } dc_operand;
typedef enum { DC_SAMPLER_NONE = 0, DC_SAMPLER_DDIM = 1, DC_SAMPLER_DDPM = 2 } dc_sampler;
/* OR-ed into a `sampler` argument: clamp pred_xstart to [-1,1] (clip_denoised=True,
 * gaussian_diffusion.py:503-507). */
#define DC_FLAG_CLIP 0x10

/* Mirrors the MotionTransformer constructor arguments that reach the kernels
 * (reference Diffusion_Stage/models/transformer.py:361-374). */
typedef struct {
    int input_feats;   /* 26  */
    int num_frames;    /* max sequence length (rows of sequence_embedding), e.g. 1800 */
    int latent_dim;    /* 128 (4*latent_dim must equal 512, reference quirk Q1) */
    int ff_size;       /* 64  */
    int num_layers;    /* >= 1 */
    int num_heads;     /* 8   */
    int device;        /* CUDA ordinal */
    int operand;       /* dc_operand: 16-bit type of the tensor-core operands (fp32 accumulate) */
} dc_config;

const char* dc_last_error(const dc_handle* h);

/* MotionTransformer.__init__ (transformer.py:361-445) */
int dc_create(const dc_config* cfg, dc_handle** out);
void dc_destroy(dc_handle* h);

/* load_state_dict (ddpm_trainer.py:303-319): one call per state_dict key; `data` may be a host or a
 * device pointer to fp32 (copied immediately).  music_encoder.* and proj.* feed dc_encode_music
 * (optional: the denoising entry points work without them); *.num_batches_tracked is ignored.  "aux.timestep_freqs" [latent_dim/2] overrides the built-in frequency
 * table of timestep_embedding (transformer.py:18-20) with the host-computed one. */
int dc_set_weight(dc_handle* h, const char* key, const void* data, const int64_t* shape, int ndim);
/* Folds LayerNorm affines, permutes FiLM rows, converts and packs every matrix into tcgen05
 * operand images.  Fails listing the first missing key. */
int dc_finalize_weights(dc_handle* h);

/* GaussianDiffusion.__init__ (gaussian_diffusion.py:328-379): `coef` is a HOST array [S][8] of the
 * fp32-rounded per-step coefficients the samplers gather (see simple_kernels.cuh for the columns);
 * also tabulates time_embed(timestep_embedding(s)) for s in [0,S) (transformer.py:482). */
int dc_set_schedule(dc_handle* h, int num_steps, const float* coef);

/* Step-invariant half of MotionTransformer.forward (transformer.py:479-480 and the K/V side of
 * LinearTemporalCrossAttention, :149-155): xf_proj, xf_out are DEVICE [B][T][64] (encode_music
 * outputs); length is a HOST int64 [B] or NULL (= all T frames valid, reference quirk Q3). */
int dc_prepare_cond(dc_handle* h, const float* xf_proj, const float* xf_out, const int64_t* length, int B, int T,
                    void* stream);

/* MotionTransformer.forward(x, timesteps, ...) (transformer.py:469-497) on the prepared condition:
 * x DEVICE [B][T][26], timesteps DEVICE int64 [B] (arbitrary per sample), out DEVICE [B][T][26]. */
int dc_forward(dc_handle* h, const float* x, const int64_t* timesteps, float* out, void* stream);

/* GaussianDiffusion.ddim_sample / p_sample for one step with t = step for the whole batch
 * (gaussian_diffusion.py:783-831 / 605-665): x is updated in place to the sample, pred_x0 receives
 * pred_xstart.  noise: DEVICE [B][T][26] or NULL (treated as zeros; exact when sigma == 0). */
int dc_sample_step(dc_handle* h, int sampler, float* x, float* pred_x0, int step, const float* noise, void* stream);

/* The sampler update alone on a caller-supplied pred_xstart (n elements). */
int dc_sampler_update(dc_handle* h, int sampler, float* x, const float* pred_x0, int step, const float* noise, int64_t n,
                      void* stream);

/* GaussianDiffusion.ddim_sample_loop / p_sample_loop (gaussian_diffusion.py:871-965 / 667-781):
 * runs steps S-1 .. 0 from x (initial noise, updated in place to the final sample) as ONE launch of the
 * persistent cluster-per-clip kernel (clips of up to 2048 frames; longer clips: a captured CUDA graph of
 * per-layer launches).  step_noise: DEVICE [S][B][T][26] in loop order or NULL.  trace_x0 / trace_x:
 * DEVICE [S][B][T][26] receiving pred_xstart / sample of every step, or NULL.  num_steps is the S the caller sized those
 * buffers for (GaussianDiffusion.num_timesteps): the call fails with DC_ERR_STATE when it is not the S of dc_set_schedule. */
int dc_sample_loop(dc_handle* h, int sampler, int num_steps, float* x, const float* step_noise, float* trace_x0, float* trace_x,
                   void* stream);

/* A block of the same loop: the n_steps consecutive steps step0, step0 - 1, ... (one launch of the persistent kernel).  Buffers are
 * [n_steps][B][T][26].  Lets a stochastic loop (DDPM, DDIM with eta > 0) bound its noise buffer: 1000 steps of a 32 x 1800-frame batch
 * would otherwise need 6 GB of pre-drawn noise. */
int dc_sample_range(dc_handle* h, int sampler, int step0, int n_steps, float* x, const float* step_noise, float* trace_x0, float* trace_x,
                    void* stream);

/* generate_music_motion (ddpm_trainer.py:183-201) with HOST buffers: uploads the encode_music
 * features and the initial noise, runs dc_prepare_cond + dc_sample_loop, downloads the motion and
 * synchronises the stream.  Host buffers should be pinned for asynchronous copies. */
int dc_generate_host(dc_handle* h, int sampler, const float* xf_proj, const float* xf_out, const int64_t* length,
                     const float* noise, float* motion_out, int B, int T, void* stream);

/* Measurement aid for bench.py: runs ONE denoise step (t = step) with CUDA events between the
 * kernels on `stream`, synchronises, and returns per kernel class the summed device time in ms and
 * the number of launches: [0] step_begin, [1] layer (tcgen05 tile kernel), [2] kv_reduce,
 * [3] out_update.  x is updated in place like dc_sample_step. */
int dc_profile_step(dc_handle* h, int sampler, float* x, int step, float* ms_out, int* count_out, void* stream);

/* Debug aid: runs one denoise step with the layer kernel recording a (clock64, event id) timeline of its
 * CTA 0; out is HOST [max_launches][512] u64: per launch [0] = event count, then (cycle, id) pairs. */
int dc_debug_timeline(dc_handle* h, float* x, int step, unsigned long long* out, int max_launches);

/* MotionTransformer.encode_music in eval mode (transformer.py:447-459) = MusicEncoder.forward (transformer.py:313-340: seven
 * reflect-padded 3x3 Conv2dResLayers with BatchNorm + ReLU, three max-pools, Conv1d 512 -> 64 + BatchNorm1d) followed by `proj`.
 * mel: DEVICE [B][Tm][128] (Tm = 3 T mel frames at 90 Hz); xf_proj, xf_out: DEVICE [B][T][64], T = (Tm - 1) / 3 + 1.  Needs the
 * music_encoder.* and proj.* keys of the state_dict (dc_set_weight); exact fp32 arithmetic. */
int dc_encode_music(dc_handle* h, const float* mel, float* xf_proj, float* xf_out, int B, int Tm, void* stream);

/* smooth_motion + pixel scaling of vis_motion (Diffusion_Stage/tools/visualization.py:20-26, 107-126): for every clip and each of
 * the C coordinates, out = savgol_filter(motion * scale, window, order) along time (scipy mode='interp').  motion, out: DEVICE
 * [B][T][C] (out must not alias motion).  fir: HOST [window] interior coefficients, edge: HOST [window/2][window] rows i = value at
 * frame i of the polynomial fitted to the first `window` frames (the tail uses the same rows time-reversed); both are computed by
 * the host layer from (window, order).  Stateless: no handle. */
int dc_smooth_motion(int device, const float* motion, float* out, int B, int T, int C, int window, const float* fir, const float* edge,
                     float scale, void* stream);

/* timestep_embedding + time_embed (transformer.py:8-25, 410-414, 482): out[i] = time_embed(timestep_embedding(timesteps[i], 128)),
 * timesteps DEVICE int64 [n], out DEVICE [n][512].  The same kernel tabulates the schedule in dc_set_schedule. */
int dc_time_embedding(dc_handle* h, const int64_t* timesteps, int n, float* out, void* stream);

/* How many clusters of `tiles_per_clip` CTAs (one clip of up to 128 * tiles_per_clip frames each) of the persistent sampling
 * kernel this device can run concurrently (cudaOccupancyMaxActiveClusters).  dc_prepare_cond fails with DC_ERR_UNSUPPORTED when
 * this is 0 for the clip length it is given (unless DC_PERSIST=0 selects the per-layer launch path). */
int dc_cluster_occupancy(dc_handle* h, int tiles_per_clip, int* max_clusters);

/* Number of this library's kernels launched so far (graph replays count their kernel nodes). */
int64_t dc_kernel_launches(const dc_handle* h);
/* Per-layer launch path only (clips longer than 2048 frames, DC_PERSIST=0): 1 = replay captured CUDA graphs in
 * dc_sample_loop (default), 0 = plain launches (profiling). */
int dc_set_graphs(dc_handle* h, int enabled);

/* Stand-alone check of the tcgen05 GEMM building block: out[M][N] = A[M][K] . W[N][K]^T + bias with
 * A, W rounded to the 16-bit operand type; all pointers HOST.  K % 64 == 0, N % 16 == 0, N <= 256. */
int dc_selftest_gemm(int device, int operand, int M, int N, int K, const float* A, const float* W, const float* bias,
                     float* out);

/* ---------------------------------------------------------------------------------------------
 * Evaluation features computed right after the sampling path.  Reference: the Evaluator of
 * Diffusion_Stage/tools/eval_new_metrics.py (pure Python / numpy there).  Errors: dc_last_error(NULL).
 * ------------------------------------------------------------------------------------------- */
typedef struct dc_eval dc_eval;

/* MotionEncoder_STGCN (eval_new_metrics.py:38-49: ST_GCN(in 2, out 32, mode 'M2S', edge importance weighting) + fc): create,
 * upload every float tensor of its state_dict by key (HOST or DEVICE pointer, `count` floats; st_gcn.A, st_gcn.data_bn.*,
 * st_gcn.st_gcn_networks.{0..9}.{gcn.conv, tcn.0, tcn.2, tcn.3}.*, st_gcn.edge_importance.{0..9}, fc.0.*, fc.1.*; st_gcn.fcn.* is
 * unused by features()), then finalize (eval-mode BatchNorms are folded on the host in fp64). */
int dc_eval_create(int device, dc_eval** out);
void dc_eval_destroy(dc_eval* e);
int dc_eval_set_weight(dc_eval* e, const char* key, const float* data, int64_t count);
int dc_eval_finalize(dc_eval* e);

/* MotionEncoder_STGCN.features(motion)[-1] (eval_new_metrics.py:62-74; ST_GCN.py:86-113, 217-228; tgcn.py:61-73): the 64-d latent
 * of every frame.  motion: DEVICE [N][T][26] keypoints (13 joints x 2), feat: DEVICE [N][T][64]. */
int dc_eval_motion_features(dc_eval* e, const float* motion, float* feat, int N, int T, void* stream);

/* Sufficient statistics of np.mean / np.cov(rowvar=False) over `rows` latents (get_scores, eval_new_metrics.py:159-168), fp64:
 * sum[64] and the centred second moments m2[64][64] = sum_r (f_r - mean)(f_r - mean)^T  (cov = m2 / (rows - 1)).  feat: DEVICE
 * [rows][64]; sum, m2: DEVICE.  The 64 x 64 matrix square root of the Frechet distance is finished on the host. */
int dc_eval_feature_stats(int device, const float* feat, int64_t rows, double* sum, double* m2, void* stream);

/* sum over rows of sum_c |a - b| in fp64 (np.mean(np.sum(np.absolute(a - b), axis=-1)) * rows): diversity
 * (eval_new_metrics.py:148-156) and the latent-space MAE (:179-185).  a, b: DEVICE [rows][64]; out: DEVICE double[1]. */
int dc_eval_feature_l1(int device, const float* a, const float* b, int64_t rows, double* out, void* stream);

/* motion_peak_onehot (eval_new_metrics.py:277-303): envelope[n][t] = sum over joints of |x_t - x_{t-1}|_2, beats[n][t] = 1 at the
 * strict local minima within +-order frames (scipy argrelextrema(np.less, order, mode='clip'); the reference uses order = 10).
 * motion: DEVICE [N][T][26]; envelope: DEVICE float [N][T]; beats: DEVICE uint8 [N][T]. */
int dc_eval_motion_beats(int device, const float* motion, float* envelope, uint8_t* beats, int N, int T, int order, void* stream);

/* alignment_score, the beat-consistency variant (eval_new_metrics.py:243-267): per clip the mean over the music beats of
 * exp(-d^2 / (2 sigma^2)), d = index distance to the nearest motion beat (0 when the clip has no motion beat).  music_beats:
 * DEVICE uint8 [N][Tm] (librosa's beat tracker in the reference -- third party, stays on the host); motion_beats: DEVICE uint8
 * [N][T]; scores: DEVICE float [N]. */
int dc_eval_beat_alignment(int device, const uint8_t* music_beats, int Tm, const uint8_t* motion_beats, int T, int N, float sigma,
                           float* scores, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DC_B200_H */
