/* libddpm.h -- C ABI of the B200-native DDPM hot path.
 *
 * The reference (paul-vdl/ImageGenerationDiffusionModels.jl) has NO FFI/plugin seam on this path:
 * its drivers call Flux/NNlib/Zygote/Optimisers directly.  This header introduces the seam at
 * exactly those call sites; every entry point names the reference lines it replaces.  The Julia
 * host (the julia/ directory of this repo) binds them with `ccall`; tests and bench bind them with
 * ctypes.  All pointers are HOST pointers unless a name ends in `_dev`.  No C++ types, no
 * exceptions, no callbacks cross this boundary.
 *
 * Conventions
 *   - return 0 on success, nonzero on failure; ddpm_last_error() gives the text (thread-local).
 *   - images cross as Float32, Julia (W,H,1,N) column-major == [N][H][W] row-major.
 *   - timesteps are 1-based Int32 (as drawn by `rand(1:T, B)`, src/train_brain.jl:227).
 *   - weights cross in Flux layouts, in BSON order: 64 arrays (conv W,b; BatchNorm beta,gamma,mu,sigma2;
 *     SURVEY.md Appendix C).  Conv W (k,k,Cin,Cout), ConvTranspose W (k,k,Cout,Cin), column-major.
 *   - the library never keeps a host pointer after a call returns and never frees one.
 *   - one handle == one GPU == one host thread at a time.  Multi-GPU = one process (or task) per GPU.
 *   - there is NO CPU fallback: without a CUDA device every compute call fails with an error.
 */
#ifndef LIBDDPM_H
#define LIBDDPM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ddpm_handle ddpm_handle;

/* arithmetic mode of the convolution contractions */
enum {
    DDPM_PREC_FP32 = 0, /* CUDA-core FP32 everywhere (parity/debug mode)                                   */
    DDPM_PREC_FP16 = 1, /* tcgen05 kind::f16, FP16 activations+weights+gradients (loss-scaled), FP32 accum */
    DDPM_PREC_BF16 = 2, /* tcgen05 kind::f16, BF16 activations+weights+gradients, FP32 accumulate           */
    DDPM_PREC_TF32 = 3  /* tcgen05 kind::tf32: FP32 activations+weights in HBM read as TF32 by the 3x3 convolutions
                           (forward and data gradient), FP32 accumulate; everything else as in FP32 mode: the
                           tensor-core parity mode (eps_hat within 2e-3 of the FP32 reference)                */
};

/* number of Float32 arrays ddpm_set_weights / ddpm_get_weights exchange */
#define DDPM_NUM_ARRAYS 64
#define DDPM_NCCL_ID_BYTES 128

const char* ddpm_last_error(void);
int ddpm_version(void);

/* Number of visible CUDA devices (0 when there is none; never fails). */
int ddpm_device_count(void);

/* Lengths of the 64 weight arrays in ABI order (for the host to validate its buffers). */
int ddpm_array_lengths(int64_t* lens /*[64]*/);

/* Create the per-GPU engine.  T, D, H, W replace the script constants
 * (src/train_brain.jl:17-18; src/generate_images.jl:11-12) and the 32x32 image size. */
int ddpm_create(ddpm_handle** out, int T, int D, int H, int W, int precision, int device);
int ddpm_destroy(ddpm_handle*);

/* Host-computed tables: beta[T], alpha_cum[T] (src/train_brain.jl:20-24) and the timestep-embedding
 * table pe[T][D] (row t-1 == timestep_embedding(t), src/train_brain.jl:54-62).  Passing them from the
 * host keeps them bit-exact with the host language by construction.  Optional: ddpm_create installs
 * the library's own restatement (Float64 libm) as default. */
int ddpm_set_tables(ddpm_handle*, const float* beta, const float* alpha_cum, const float* pe);
int ddpm_get_tables(ddpm_handle*, float* beta, float* alpha_cum, float* pe, float* sampler_scalars /*[T][4] or NULL*/);

/* Model exchange: replaces reading the fields of `SimpleUNet` (src/train_brain.jl:89-96) after
 * `@load` (src/generate_images.jl:250) and before `@save` (src/train_brain.jl:295-300). */
int ddpm_set_weights(ddpm_handle*, const float* const* arrays, const int64_t* lens, int n);
int ddpm_get_weights(ddpm_handle*, float* const* arrays, const int64_t* lens, int n);

/* `Adam(lr)` + `Flux.setup` (src/train_brain.jl:255-257): sets the rule and zeroes the moments. */
int ddpm_set_adam(ddpm_handle*, float eta, float beta1, float beta2, float eps);

/* Optimiser state for a true resume.  The reference saves only the RULE (`@save ... model opt`,
 * src/train_brain.jl:295-300; `state = Flux.setup(opt, model)` at :257 is never written), so a run continued from
 * one of its checkpoints restarts Adam from zero moments.  These two calls move the moment arenas m, v (64 arrays in
 * the weight order; zero for mu/sigma2), the bias-correction powers beta_t = (beta1^t, beta2^t) of the next update
 * and the number of updates applied so far. */
int ddpm_get_adam_state(ddpm_handle*, float* const* m, float* const* v, const int64_t* lens, int n,
                        float* beta_t /*[2]*/, int64_t* steps);
int ddpm_set_adam_state(ddpm_handle*, const float* const* m, const float* const* v, const int64_t* lens, int n,
                        const float* beta_t /*[2]*/, int64_t steps);

/* `x_t = a .* x0 .+ b .* eps` of train_step (src/train_brain.jl:230-233). */
int ddpm_q_sample(ddpm_handle*, const float* x0, const int32_t* ts, const float* eps, int B, float* x_t);

/* `m((x_t, t_emb))` (src/train_brain.jl:238 / src/generate_images.jl:183): train_mode=1 uses batch
 * statistics (and does NOT touch the running statistics), train_mode=0 the running statistics. */
int ddpm_predict_eps(ddpm_handle*, const float* x_t, const int32_t* ts, int B, int train_mode, float* eps_hat);

/* One iteration of the training loop body (src/train_brain.jl:267-274): q_sample, forward in train
 * mode (running statistics updated), mse, backward, (gradient all-reduce,) Adam update.
 * `loss` receives the scalar Float32 loss.  B_global = B * world_size when a communicator is set. */
int ddpm_train_step(ddpm_handle*, const float* x0, const int32_t* ts, const float* eps, int B, float* loss);

/* Same, but gradients only: fills `grads` (64 arrays, zero for mu/sigma2) and leaves weights,
 * moments and running statistics untouched -- `Flux.withgradient` alone (src/train_brain.jl:267). */
int ddpm_loss_and_grad(ddpm_handle*, const float* x0, const int32_t* ts, const float* eps, int B,
                       float* loss, float* const* grads, const int64_t* lens, int n);

/* Throughput form of the train step: x0 taken from a device-resident dataset uploaded once with
 * ddpm_upload_dataset, ts and eps drawn on the device (Philox keyed by seed, step, global image
 * index).  idx = n_idx host indices (0-based) into the dataset, or NULL for the first B images. */
int ddpm_upload_dataset(ddpm_handle*, const float* imgs, int64_t n_imgs);
int ddpm_train_step_device(ddpm_handle*, const int32_t* idx, int B, uint64_t seed, int64_t step, float* loss /*NULL: skip D2H*/);

/* `generate_image` (src/generate_images.jl:231-245) / `denoise_image`: the reverse loop
 * for t = t_start .. 2, then clamp to [-1,1].
 *   x_T : N*H*W start images, or NULL -> device N(0,1) keyed by (seed, first_index+i, step=T... )
 *   z   : (t_start-1)*N*H*W noise, z[k] used at step t = t_start-k, or NULL -> device Philox
 *   first_index : global index of image 0 of this call (sharding across GPUs/calls keeps outputs
 *                 independent of how N is split when the device generator is used)
 *   out : N*H*W */
int ddpm_sample(ddpm_handle*, const float* x_T, const float* z, uint64_t seed, int64_t N,
                int64_t first_index, int t_start, float* out);

/* Device-resident variant for throughput measurement: no host traffic at all; the result stays in
 * an internal buffer readable with ddpm_sample_fetch. */
int ddpm_sample_device(ddpm_handle*, uint64_t seed, int64_t N, int64_t first_index, int t_start);
int ddpm_sample_fetch(ddpm_handle*, int64_t N, float* out);
/* Output step of bulk dumps (`(img .+ 1) ./ 2` -> Gray image, src/generate_images.jl:256-265): the device-resident
 * result of ddpm_sample_device quantised on the device to 8 bits, round(clamp((x+1)/2, 0, 1) * 255) with ties to
 * even -- the byte the host PNG writer would produce --, N*H*W bytes in the image layout of `out` above. */
int ddpm_sample_fetch_u8(ddpm_handle*, int64_t N, uint8_t* out);

/* `apply_noise` (src/ImageGenerationDiffusionModels.jl:60-73): the same-eps recurrence
 * `img = sqrt(1-beta).*img + sqrt(beta).*epsilon` over the host-supplied Float64 `betas`
 * (`collect(beta_min:(beta_max-beta_min)/num_noise_steps:beta_max)`, n_betas = steps+1), evaluated per
 * element in one kernel with separately rounded Float64 multiplies and adds (bit-exact with the
 * reference's broadcast).  Needs a GPU but no handle. */
int ddpm_apply_noise_f64(const double* img, const double* eps, int64_t n, const double* betas, int n_betas,
                         double* out);

/* Data-parallel training: one NCCL communicator over all ranks.  Rank 0 calls ddpm_comm_unique_id,
 * the host distributes the 128 bytes (torch.distributed / MPI / files), every rank calls
 * ddpm_comm_init.  sync_bn=1 all-reduces the BatchNorm batch statistics so the result equals the
 * single-process global-batch semantics of the reference. */
int ddpm_comm_unique_id(void* id_out /*128 bytes*/);
int ddpm_comm_init(ddpm_handle*, const void* id, int rank, int world, int sync_bn);

/* Instrumentation used by bench.py / tests (not needed by the Julia host).  Option keys (value, default):
 *   sample_chunk (images per captured reverse-loop graph, 1300; a batch is cut into equal chunks no larger than this),
 *   sample_streams (1), use_graph (1), conv_impl (0 auto | 1 CUDA-core | 2 tcgen05), sync_bn (1),
 *   fuse_final (1: last conv + final 1x1 conv + reverse update in one epilogue),
 *   tc_pair (bit mask of layer shapes run as CTA pairs / cta_group::2, 31 = all), tc_pdl (1: programmatic dependent
 *   launch of the tcgen05 kernels), tc_reverse (1: consecutive conv layers of the sampler walk their tiles in alternating
 *   directions, so each starts with the part of its input that is still in L2), conv1_tc (2: first conv of the sampler on tensor cores with the timestep constants folded into the contraction, 1: constants added in the epilogue, 0: CUDA cores), tc_tma_store (1),
 *   tc_role_profile (0: in-kernel cycle counters), train_graph (1: replay the training iteration from a CUDA graph),
 *   bn_p2p (1: SyncBN statistics exchanged over peer-memory mailboxes inside the finalize kernels; 0: one NCCL all-reduce
 *   per layer), dp_skip (0; TIMING ONLY, results become wrong: bit 0 skips the gradient all-reduces, bit 1 the SyncBN ones),
 *   fuse_bn (1: train-mode BatchNorm reductions inside the tcgen05 conv / data-gradient epilogues),
 *   bnbwd_blocks (4: resident blocks per SM of the second BatchNorm-backward pass), train_reverse (1: the BatchNorm apply
 *   and backward kernels of a training step walk their tensors from the end, where the producing conv left them in L2),
 *   loss_scale_log2 (0: extra power-of-two factor on the static loss scale of the 16-bit gradient tensors).
 * None of them changes results beyond the documented rounding of the selected kernels.
 * Counter keys: launches, n_params, tc_available, uses_tc, bn_p2p_active, skipped_steps (updates skipped by the overflow guard: a
 * non-finite value in the reduced gradient leaves weights, moments and beta^t untouched), applied_steps. */
int ddpm_set_option(ddpm_handle*, const char* key, int64_t value);
int64_t ddpm_get_counter(ddpm_handle*, const char* key);
/* CUDA-event stopwatch on the engine's launch stream (device time of everything enqueued between
 * the two calls); ddpm_timer_stop synchronises on the stop event. */
int ddpm_timer_start(ddpm_handle*);
int ddpm_timer_stop(ddpm_handle*, float* ms);
/* Time `iters` launches of a named internal kernel on device-resident synthetic data with CUDA
 * events on the launching stream; returns mean milliseconds per launch in *ms and the algorithmic
 * bytes and flops of one launch. */
int ddpm_time_kernel(ddpm_handle*, const char* name, int64_t n_images, int iters, float* ms,
                     double* bytes, double* flops);
/* Copy an internal activation/gradient tensor (by name, e.g. "y3", "a10", "dy2") of the last
 * forward/backward to the host as Float32 [N][C][H][W]; for per-layer parity tests. */
int ddpm_debug_fetch(ddpm_handle*, const char* name, float* out, int64_t capacity, int64_t* written);

#ifdef __cplusplus
}
#endif
#endif /* LIBDDPM_H */
