/*
 * zs_ae.h - C ABI of libzsae.so: the B200 (sm_100a) hot path of the ZeroSpeech
 * "TTS without T" ASR-TTS autoencoder.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; the boundary
 * it exposes for this path is the nn.Module API of
 *   model/model.py:368-489  Encoder.__init__/forward
 *   model/model.py:283-365  Decoder.__init__/forward
 * called from trainer.py:194-228 (test_step / encoder_test_step) and
 * trainer.py:246-254 (encode_step / decode_step).  Each entry point below names
 * the reference lines it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless it says "host";
 *    the library never frees caller memory;
 *  - tensors named like the reference's are fp32, (batch, channel, time)
 *    contiguous, exactly as model/model.py passes them;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), nothing
 *    synchronises the device;
 *  - return 0 = ok; non-zero = error, message via zs_last_error() (thread local);
 *  - there is no CPU fallback: without a CUDA device every compute entry fails.
 */
#ifndef ZS_AE_H
#define ZS_AE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZS_OK 0
#define ZS_ERR_ARG 1      /* bad argument / unsupported shape or mode */
#define ZS_ERR_CUDA 2     /* CUDA runtime / driver error */
#define ZS_ERR_WORKSPACE 3 /* workspace too small */

/* Encoder.enc_mode (model/model.py:457-487).  'binary' projects to enc_size^2 channels (model/model.py:466-472): enc_size <= 128. */
enum { ZS_ENC_CONTINUES = 0, ZS_ENC_ONE_HOT = 1, ZS_ENC_MULTILABEL_BINARY = 2, ZS_ENC_GUMBEL_T = 3, ZS_ENC_BINARY = 4 };
/* tensor-core operand type of activations/weights (accumulation is always fp32) */
enum { ZS_OPERAND_FP16 = 0, ZS_OPERAND_BF16 = 1 };

typedef struct zs_encoder zs_encoder;   /* packed Encoder weights (library owned) */
typedef struct zs_decoder zs_decoder;   /* packed Decoder weights + per-speaker bias tables */

/* Encoder(c_in, c_h1, c_h2, c_h3, ns, dp, enc_size, seg_len, enc_mode): model/model.py:369 */
typedef struct {
    int32_t c_in, c_h1, c_h2, c_h3;
    int32_t enc_size;
    int32_t enc_mode;     /* ZS_ENC_* */
    int32_t seg_len;      /* hyper-parameter; reflect padding iff >= 64 (model/model.py:38). < 64 is rejected */
    int32_t operand;      /* ZS_OPERAND_* */
    float ns;             /* leaky-relu negative slope */
    int32_t train;        /* 1: also pack the data-gradient operands (zs_encoder_forward_train / _backward); fp16, one_hot only */
} zs_encoder_cfg;

/* fp32 parameter pointers in state_dict order (SURVEY.md Appendix B / model/model.py:373-393) */
typedef struct {
    const float* conv1s_w[7];   /* (c_h1, c_in, k) k = 1..7 */
    const float* conv1s_b[7];
    const float* conv_w[7];     /* conv2 (c_h2, 7*c_h1+c_in, 1), conv3..conv8 (c_h2, c_h2, 5) */
    const float* conv_b[7];
    const float* dense_w[4];    /* (c_h2, c_h2) */
    const float* dense_b[4];
    const float* gru_w_ih[2];   /* [0] forward, [1] reverse: (3*c_h3, c_h2) */
    const float* gru_w_hh[2];   /* (3*c_h3, c_h3) */
    const float* gru_b_ih[2];
    const float* gru_b_hh[2];
    const float* linear_w;      /* (n_out, c_h2 + 2*c_h3), n_out = enc_size or 2*enc_size */
    const float* linear_b;
} zs_encoder_weights;

/* Decoder(c_in, c_out, c_h, c_a, ns, seg_len, output_mask): model/model.py:284 */
typedef struct {
    int32_t c_in;         /* enc_size */
    int32_t c_out;        /* 513 */
    int32_t c_h;          /* emb_size */
    int32_t c_a;          /* number of speakers */
    int32_t seg_len;
    int32_t output_mask;  /* 1: tanh (g_mode targeted_residual), 0: sigmoid */
    int32_t operand;
    float ns;
    int32_t train;        /* 1: training handle - speaker embeddings are added to the activations instead of being folded
                           * into bias tables (their gradients need them), data-gradient operands are packed too */
} zs_decoder_cfg;

typedef struct {
    const float* conv_w[6];     /* conv1..conv6: odd (2*c_h, c_h, 3), even (c_h, c_h, 3) */
    const float* conv_b[6];
    const float* dense_w[4];    /* (c_h, c_h) */
    const float* dense_b[4];
    const float* gru_w_ih[2];   /* (3*c_h/2, c_h) */
    const float* gru_w_hh[2];   /* (3*c_h/2, c_h/2) */
    const float* gru_b_ih[2];
    const float* gru_b_hh[2];
    const float* dense5_w;      /* (c_h, 3*c_h) */
    const float* dense5_b;
    const float* linear_w;      /* (c_out, c_h) */
    const float* linear_b;
    const float* input_emb_w;   /* (c_h, c_in) */
    const float* input_emb_b;
    const float* emb[5];        /* emb1..emb5 (c_a, c_h) */
} zs_decoder_weights;

const char* zs_last_error(void);
/* 10000*major + 100*minor + patch of this library */
int zs_version(void);
/* 0 when the current CUDA device is sm_100 class and the driver entry points resolve */
int zs_device_check(void);

/* fp16 operands carry a TF32-class mantissa but only +-65504 of range.  Every GEMM epilogue clamps its fp16 outputs to
 * that range and COUNTS the epilogue threads that had to (nothing on the reference's value ranges comes close: the
 * activations are O(1) after InstanceNorm).  Returns that count for the current device since the last reset after
 * synchronising `stream`; a non-zero count means results are degraded and the checkpoint needs operand = bf16.  The
 * Python front-end raises on it (the reference has no counterpart: it computes in fp32, model/model.py:20-110). */
int zs_saturation_count(void* stream, unsigned long long* count, int reset);

/* Packs fp32 parameters (device pointers) into tensor-core operand layout (fp16/bf16,
 * K-major rows of [taps][c_in padded to 64]) and, for the decoder, folds the five speaker
 * embeddings into per-(speaker, layer) fp32 bias tables (valid because a time-constant
 * survives reflect padding: conv(x+e) = conv(x) + sum_k W_k e).  Replaces what
 * Encoder.__init__/load_state_dict materialise (trainer.py:58-59, 135-140). */
int zs_encoder_pack(const zs_encoder_cfg* cfg, const zs_encoder_weights* w, void* stream, zs_encoder** out);
int zs_decoder_pack(const zs_decoder_cfg* cfg, const zs_decoder_weights* w, void* stream, zs_decoder** out);
void zs_encoder_free(zs_encoder* h);
void zs_decoder_free(zs_decoder* h);

/* bytes of scratch the forward needs for B segments of T frames (T <= 256) */
size_t zs_encoder_workspace_bytes(const zs_encoder* h, int B, int T);
size_t zs_decoder_workspace_bytes(const zs_decoder* h, int B, int T8);

/* Encoder.forward in eval mode (model/model.py:440-489; trainer.py:197, 227).
 *   x            (B, c_in, T) fp32
 *   gumbel_noise one_hot: (B, T8, enc_size); multilabel_binary: (B, T8, enc_size, 2); binary: (B, T8, enc_size, enc_size);
 *                gumbel_t: (B, enc_size, T8) - the value of _sample_gumbel() (model/model.py:95-98);
 *                NULL for `continues`
 *   logits       (B, n_out, T8) fp32  - the reference's second return value `out`
 *   act          (B, enc_size, T8) fp32 - the reference's first return value `out_act`
 *   unit_ids     (B, T8) int32, one_hot only (argmax over units, first index wins ties); may be NULL
 * T8 = ceil(ceil(ceil(T/2)/2)/2). */
int zs_encoder_forward(zs_encoder* h, const float* x, int B, int T, const float* gumbel_noise,
                       float* logits, float* act, int32_t* unit_ids,
                       void* workspace, size_t workspace_bytes, void* stream);

/* zs_encoder_forward for the batched front-end (replaces the per-chunk upload + permute of convert.py:70-83 and
 * trainer.py:196, 226): the same computation with
 *   x_dtype   ZS_X_F32, or ZS_X_F16 - an fp16 upload halves the host-to-device bytes; the path rounds its input to fp16
 *             operands first thing anyway, so the results are bit-identical to the fp32 upload (fp16 operands only);
 *   x_layout  ZS_X_NCT (B, c_in, T) as Encoder.forward receives it, or ZS_X_NTC (B, T, c_in) as Trainer.test_step receives
 *             it BEFORE its permute (the layout of the HDF5 features, dataloader.py:74) - no transpose copy needed;
 *   gumbel_noise NULL in one_hot mode: the noise is generated on the device from `noise_seeds` ((B,) uint64 on the device,
 *             one counter-based stream per segment, so a segment's draw does not depend on the batch it rides in; same
 *             distribution as model/model.py:95-98 but not the reference's CPU generator stream - a throughput mode that
 *             saves 4 KB of upload per unit frame; pass the tensor for reference-exact draws);
 *   logits / act  may be NULL when the caller only wants the unit ids (one_hot). */
enum { ZS_X_F32 = 0, ZS_X_F16 = 1 };
enum { ZS_X_NCT = 0, ZS_X_NTC = 1 };
int zs_encoder_forward_x(zs_encoder* h, const void* x, int x_dtype, int x_layout, int B, int T, const float* gumbel_noise,
                         const uint64_t* noise_seeds, float* logits, float* act, int32_t* unit_ids,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Decoder.forward (model/model.py:344-365; trainer.py:199, 253).
 *   enc_act  (B, c_in, T8) fp32 dense activations, or NULL when unit_ids is given
 *   unit_ids (B, T8) int32 one-hot fast path (input_emb becomes a column gather), or NULL
 *   spk      (B,) int64 speaker ids in [0, c_a)
 *   spec     (B, c_out, 8*T8) fp32
 *   accumulate 0: spec = y; 1: spec += y (trainer.py:207-209 'naive'/'targeted');
 *              2: spec += spec * y (trainer.py:211 'targeted_residual') */
int zs_decoder_forward(zs_decoder* h, const float* enc_act, const int32_t* unit_ids, const int64_t* spk,
                       int B, int T8, float* spec, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* zs_decoder_forward with the output element type chosen: spec_dtype ZS_X_F32 (what Decoder.forward returns) or ZS_X_F16 -
 * the sigmoid output in (0, 1) rounded once to fp16 (absolute error <= 2.5e-4, inside the path's 1e-2 tolerance), which halves
 * the device-to-host bytes of the streaming front-end (the reference downloads fp32 after every chunk, trainer.py:221).
 * The accumulate rules then read and write fp16. */
int zs_decoder_forward_x(zs_decoder* h, const float* enc_act, const int32_t* unit_ids, const int64_t* spk,
                         int B, int T8, void* spec, int spec_dtype, int accumulate,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- Spectrogram_Patcher (model/model.py:503-549): the TTS patcher of g_mode 'spectrogram' (trainer.py:78-79), applied as
 * x_dec += Generator(x_dec, c - shift) (trainer.py:212-213, 280-281).  Per-frame Linear 513 -> c_h, two dense blocks
 * conditioned on emb1, bi-GRU on out + emb2, dense5 on cat([out, rnn, emb2]), Linear -> sigmoid: the recurrent tail of the
 * Decoder on a spectrogram input, run by the same kernels. */
typedef struct zs_patcher zs_patcher;
typedef struct {
    int32_t c_in;         /* 513 */
    int32_t c_out;        /* 513 */
    int32_t c_h;          /* emb_size */
    int32_t c_a;          /* n_target_speakers */
    int32_t operand;
    float ns;
} zs_patcher_cfg;
typedef struct {
    const float* input_w;       /* input_layer (c_h, c_in) */
    const float* input_b;
    const float* dense_w[4];    /* (c_h, c_h) */
    const float* dense_b[4];
    const float* gru_w_ih[2];   /* (3*c_h/2, c_h) */
    const float* gru_w_hh[2];
    const float* gru_b_ih[2];
    const float* gru_b_hh[2];
    const float* dense5_w;      /* (c_h, 3*c_h) */
    const float* dense5_b;
    const float* linear_w;      /* (c_out, c_h) */
    const float* linear_b;
    const float* emb[2];        /* emb1, emb2 (c_a, c_h) */
} zs_patcher_weights;
int zs_patcher_pack(const zs_patcher_cfg* cfg, const zs_patcher_weights* w, void* stream, zs_patcher** out);
void zs_patcher_free(zs_patcher* h);
size_t zs_patcher_workspace_bytes(const zs_patcher* h, int B, int T);
/* x (B, c_in, T) fp32, spk (B,) int64 in [0, c_a), spec (B, c_out, T) fp32; accumulate as in zs_decoder_forward (x and
 * spec may alias: the input is consumed by the first kernel, the output written by the last). T <= 256. */
int zs_patcher_forward(zs_patcher* h, const float* x, const int64_t* spk, int B, int T, float* spec, int accumulate,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- the DSP steps either side of the path: Griffin-Lim vocoder and featurisation --------------------
 * Constants are the reference's (hps/hps.py:22-33): 16 kHz, n_fft 1024, hop 200, Hann window of 800, 513 bins,
 * max_db 100, ref_db 20.  A batch holds U utterances back to back; `meta` is a DEVICE int32 array of 3 x (U + 1) prefix
 * sums: frame_start[u], sample_start[u], tile_start[u] with tiles per utterance = ceil(frames / zs_stft_tile_frames()).
 * librosa (stft / istft: centre = True, reflect padding, window sum-of-squares normalisation) is not part of the
 * reference tree; its published algorithm is restated in oracle/dsp_oracle.py, which the tests hold these kernels to. */
int zs_stft_tile_frames(void);
size_t zs_griffin_lim_workspace_bytes(long long total_frames, long long total_samples);
/* spectrogram2wav (convert.py:55-62) up to (not including) the final librosa.effects.trim:
 *   spec  [total_frames][513] fp32 normalised log-magnitude rows (the Decoder's output, transposed to frames-major);
 *         every utterance needs >= 4 frames
 *   wav   [total_samples] fp32; utterance u has 200 * (frames_u - 1) samples at sample_start[u]
 * = de-normalise, n_iter x { istft, stft, X = mag * est / max(1e-8, |est|) } (convert.py:39-52; one fused kernel per
 * iteration, the waveform stays in shared memory), final istft, de-pre-emphasis y[t] = x[t] + preemphasis * y[t-1]. */
int zs_griffin_lim(const float* spec, const int32_t* meta, int U, long long total_frames, long long total_samples,
                   int total_tiles, int n_iter, float preemphasis, float* wav,
                   void* workspace, size_t workspace_bytes, void* stream);
/* mean-square power of the centred 2048-sample frames (hop 512, reflect padding) librosa.effects.trim thresholds
 * (convert.py:61); pframe_start = device prefix sums of 1 + samples_u / 512; power[pframe_start[u] + f]. */
int zs_frame_power(const float* wav, const int32_t* sample_start, const int32_t* pframe_start, int U, int max_frames,
                   float* power, void* stream);
/* get_spectrograms (preprocess.py:233-256) after file decoding / trimming: pre-emphasis, stft, magnitude,
 * 20 log10(max(1e-5, .)), clip((db - 20 + 100) / 100, 1e-8, 1) -> [total_frames][513] rows, fp32 and / or fp16 (either may
 * be NULL; the fp16 rows are the encoder's ZS_X_F16 / ZS_X_NTC input).  frames_u = 1 + samples_u / 200, samples_u >= 513;
 * here sample_start are prefix sums of the INPUT lengths. */
int zs_spectrogram(const float* wav, const int32_t* meta, int U, int total_tiles, float preemphasis,
                   float* spec32, void* spec16, void* stream);

/* ---- pretrain_AE step (trainer.py:321-332: encode_step, decode_step, L1 loss, backward, clip, Adam) -------------
 *
 * Training handles are packed with cfg.train = 1 and re-packed from the updated fp32 parameters after every
 * optimiser step (zs_*_repack: same allocations, no cudaMalloc).  Activation gradients travel as fp16 buffers
 * multiplied by `loss_scale` (a power of two; 2^15 * B is a good start); weight gradients are written UNSCALED in fp32
 * into caller-owned buffers shaped like the parameters (`zs_*_weights` used as a table of gradient pointers) and are
 * written into ZEROED buffers - zero them before every backward (partial sums are combined with atomic adds, whole
 * reductions are stored; a backward does not accumulate on top of an earlier one).  An fp16 overflow surfaces as a non-finite gradient
 * norm: zs_adam_step then skips the update and raises *skipped so the caller can halve the scale.
 * T must be 64 or 128 (seg_len; the weight-gradient GEMM reduces over 64-row boxes). */
int zs_encoder_repack(zs_encoder* h, const zs_encoder_weights* w, void* stream);
int zs_decoder_repack(zs_decoder* h, const zs_decoder_weights* w, void* stream);
size_t zs_encoder_train_workspace_bytes(const zs_encoder* h, int B, int T);
size_t zs_decoder_train_workspace_bytes(const zs_decoder* h, int B, int T8);

/* Encoder.forward in TRAIN mode (model/model.py:440-489 with the six nn.Dropout active).  Same outputs as
 * zs_encoder_forward; the workspace keeps what the backward needs and must stay untouched until it ran.
 *   dropout_p     Encoder(dp=...); 0 disables dropout
 *   dropout_seed  counter-based masks: keep(seed, layer, b, c, t), reproduced by the backward
 *   dropout_seed_dev  NULL, or a device pointer the kernels read the seed from instead (a CUDA-graph replay of the
 *                 step then changes the masks by updating that word, without re-capture)
 *   keep_masks    NULL, or six device pointers to explicit (B, c_h2, T_l) byte masks in the reference's layout
 *                 (1 = keep) - lets a test replay the reference's own bernoulli draws */
int zs_encoder_forward_train(zs_encoder* h, const float* x, int B, int T, const float* gumbel_noise,
                             float dropout_p, uint64_t dropout_seed, const uint64_t* dropout_seed_dev,
                             const uint8_t* const* keep_masks, float* logits, float* act, int32_t* unit_ids,
                             void* workspace, size_t workspace_bytes, void* stream);
/* backward of the above: d_act (B, enc_size, T8) fp32 = (dLoss/d out_act) * d_act_scale; the same
 * gumbel_noise / logits / dropout arguments as the forward call; gradients accumulate into `grads`. */
int zs_encoder_backward(zs_encoder* h, const float* d_act, float d_act_scale, const float* gumbel_noise,
                        const float* logits, int B, int T, float dropout_p, uint64_t dropout_seed,
                        const uint64_t* dropout_seed_dev, const uint8_t* const* keep_masks, float loss_scale,
                        const zs_encoder_weights* grads,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Decoder.forward in train mode (dense enc_act input so that it can take a gradient). */
int zs_decoder_forward_train(zs_decoder* h, const float* enc_act, const int64_t* spk, int B, int T8, float* spec,
                             void* workspace, size_t workspace_bytes, void* stream);
/* backward: either `target` (B, c_out, T) is given - the L1 loss of trainer.py:327 is fused: *loss += mean|spec - target|
 * (zero it first) - or `d_spec` (B, c_out, T) fp32 = dLoss/dspec (unscaled).  d_act (B, c_in, T8) fp32 receives
 * (dLoss/d enc_act) * loss_scale. */
int zs_decoder_backward(zs_decoder* h, const float* spec, const float* target, const float* d_spec,
                        const int64_t* spk, int B, int T8, float loss_scale, float* loss,
                        const zs_decoder_weights* grads, float* d_act,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Weight / bias gradients off the critical path.  The reference's loss.backward() (trainer.py:329) leaves the order of
 * independent gradient kernels to autograd; here zs_wgrad_async(1) makes every following zs_*_backward call of THIS host
 * thread launch its weight- and bias-gradient kernels on a library-owned side stream of the current device, forked from
 * `stream` right after the kernel that produced their input (event record / wait: capturable into a CUDA graph), so they
 * overlap the data-gradient chain.  The gradient buffers are then complete on `stream` only after zs_wgrad_join(stream);
 * data gradients (d_act) and the loss are unaffected.  zs_wgrad_async(0) restores in-order launches (the default); it
 * fails while gradients are in flight.  Call zs_wgrad_async(1) once outside a stream capture (it creates the stream). */
int zs_wgrad_async(int on);
int zs_wgrad_join(void* stream);

/* sum of squares of a flat fp32 gradient buffer: *out += sum g^2 (zero it first) */
int zs_grad_sqnorm(const float* g, size_t n, float* out, void* stream);
/* nn.utils.clip_grad_norm_(max_norm) over ONE network (utils.py:53-55) fused with torch.optim.Adam's update
 * (trainer.py:64-66: lr, betas (0.5, 0.9), eps 1e-8, no weight decay): `sqnorm` is that network's squared
 * gradient norm (device scalar) BEFORE `grad_mult`, a factor applied to every gradient first (1/world_size after a
 * summing all-reduce); step = 1-based step count for the bias corrections - or bias_corr_dev = device pointer to
 * {1 - beta1^step, sqrt(1 - beta2^step)} (CUDA-graph replays).  When the norm is not
 * finite nothing is updated and *skipped (device int, may be NULL) is set to 1.
 * bias_corr_dev, when given, is words [2..4] of a zs_train_meta block: {float bc1, float bc2_sqrt, int32 apply} - with
 * apply == 0 the call updates nothing (another network of the same optimiser overflowed: trainer.py:64-66 builds ONE
 * Adam over both networks, so they step together or not at all). */
int zs_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n,
                 const float* sqnorm, float grad_mult, float max_norm, float lr, float beta1, float beta2, float eps,
                 int step, const float* bias_corr_dev, int* skipped, void* stream);

/* Device-resident step state of the training iteration: 8 x 32-bit words
 *   [0,1] dropout seed (uint64) | [2] 1 - beta1^n | [3] sqrt(1 - beta2^n) | [4] apply | [5] n = applied optimiser steps |
 *   [6,7] iterations started (uint64).  Zero-initialise once.  Everything advances ON THE STREAM, so a CUDA-graph replay of
 * the iteration needs no host writes (a pinned-host "meta" word could be overwritten while an earlier replay still reads it).
 *   zs_train_meta_begin : iterations += 1; seed = splitmix64(iterations, seed_salt)  (salt = rank-dependent constant, so
 *                         data-parallel ranks draw different dropout masks like the reference step on the concatenated batch)
 *   zs_train_meta_commit: apply = every given squared gradient norm is finite; if so n += 1 and the bias corrections of
 *                         step n are stored, else *skipped = 1 and n stays (only applied steps count, trainer.py:332). */
int zs_train_meta_begin(void* meta, uint64_t seed_salt, void* stream);
int zs_train_meta_commit(void* meta, const float* sqnorm_a, const float* sqnorm_b, float beta1, float beta2, int* skipped,
                         void* stream);

/* ---- measurement hooks (bench.py) --------------------------------------------------
 * Kernel classes: 0 = conv/linear implicit GEMM (tcgen05), 1 = GRU recurrence, 2 = everything else.
 * Between zs_profile_begin() and zs_profile_end() every kernel launch of this library is bracketed
 * by CUDA events on its stream; zs_profile_end() synchronises those events and returns, per class,
 * the summed device milliseconds, the algorithmic FLOPs and the launch count (arrays of 3).
 * zs_launch_counts() returns the launch counters without timing (counted since profile_begin). */
void zs_profile_begin(void);
int zs_profile_end(double* ms, double* flops, long long* launches);
void zs_launch_counts(long long* launches);
/* Per-launch detail of the spans recorded since zs_profile_begin(), in launch order (call BEFORE zs_profile_end;
 * synchronises the events): fills up to `max` entries of milliseconds / algorithmic FLOPs / kernel class and
 * returns the number of recorded launches. */
int zs_profile_detail(double* ms, double* flops, int* cls, int max);
/* kernel label of recorded launch i (valid until zs_profile_end) */
const char* zs_profile_name(int i);

/* ---- building blocks, exported for the unit tests -------------------------------- */

/* gumbel_softmax forward value (model/model.py:93-110) on logits laid out (B, C, T8):
 * ids[b,t] = argmax_c(logits[b,c,t] + noise[b,t,c]); act[b,c,t] = (c == ids[b,t]). */
int zs_bottleneck_one_hot(const float* logits, const float* noise, int B, int C, int T8,
                          float* act, int32_t* unit_ids, void* stream);

/* One fused conv1d / per-frame-linear layer on channels-last operand buffers:
 * implicit GEMM on tcgen05 tensor cores, TMA-fed, fp32 accumulate in TMEM, epilogue =
 * bias (+ per-speaker table) -> leaky-relu -> InstanceNorm -> residual -> activation -> store. */
typedef struct {
    /* A operand: packed weights [m_rows (multiple of 128)][k_total] operand-type, K-major */
    const void* w;
    int32_t m_rows, m_valid;
    int32_t taps;            /* taps per output frame (kernel size) */
    int32_t c_in_pad;        /* channels per tap, multiple of 64; k_total = w_taps * c_in_pad */
    int32_t w_taps;          /* taps stored per weight row (>= taps; 7 for the merged conv bank) */
    int32_t bank;            /* 1: m-tile i uses kernel size i+1 placed at tap 3-(i+1)/2 (conv bank) */
    /* B operand: activations [B][in_rows][in_pitch] operand-type, channels-last, halo rows included */
    const void* in;
    int32_t in_rows, in_pitch, in_row0;  /* in_row0: buffer row read by tap 0 of output frame 0 */
    int32_t c_in_valid;      /* channels that exist in `in` (<= in_pitch); TMA zero-fills up to c_in_pad */
    int32_t stride;          /* 1 or 2 */
    int32_t B, T_out;        /* segments, valid output frames per segment */
    /* epilogue */
    const float* bias;       /* [m_rows] or per-speaker table [n_spk][m_rows] */
    const int64_t* spk;      /* NULL = shared bias */
    int32_t n_spk;           /* rows of the per-speaker table; ids are clamped into [0, n_spk) */
    int32_t lrelu; float ns;
    int32_t inorm;           /* InstanceNorm over the T_out frames of each (segment, channel) */
    int32_t res_mode;        /* 0 none, 1 same frame, 2 avg of frames 2t,2t+1, 3 frame t/2 */
    const void* res; int32_t res_rows, res_pitch, res_halo;
    int32_t act;             /* 0 none, 1 sigmoid, 2 tanh */
    int32_t out_mode;        /* 0 channels-last operand-type, 1 pixel-shuffle channels-last, 2 fp32 (B, m_valid, T_out) */
    void* out; int32_t out_rows, out_pitch, out_halo, out_choff;   /* out_halo: reflected rows written each side, 0..3 (kernel sizes up to 7) */
    int32_t accumulate;      /* out_mode 2 only: 0 store, 1 out += y, 2 out += out*y */
    int32_t operand;         /* ZS_OPERAND_* */
    int32_t nb_hint;         /* segments per N tile, 0 = auto */
    int32_t out_f16;         /* out_mode 2 only: the (B, m_valid, T_out) output is fp16 instead of fp32 */
} zs_conv_desc;
int zs_conv1d_cl(const zs_conv_desc* d, void* stream);
/* Layers with an even number of 128-channel tiles and of segments per tile run as CTA pairs (tcgen05 cta_group::2, M = 256; each CTA
 * stages half of the tile's columns).  mode 0 turns that off process-wide (one CTA per tile everywhere: the tests' A/B reference),
 * mode 1 (default; 2 is accepted as a synonym) pairs every layer that qualifies.
 * Adding 0x100 also turns the four-stage ring of the single-CTA pixel-shuffle layers off (three stages, two output tiles); adding
 * 0x200 uses it for every single-CTA layer without a residual (tests); adding 0x400 turns the tap-reusing main loop off (every
 * tap then re-stages its shifted activation tile), adding 0x800 uses it in CTA pairs too (tests; measured slower there).
 * Results are bit-identical in all modes. */
void zs_set_gemm_pair_mode(int mode);

/* ---- 2-D critic / classifier forward (model/model.py:113-226: PatchDiscriminator, TargetClassifier) --------
 * A k x k stride-s Conv2d over (H, W) with reflect padding (pad_layer(is_2d=True), model/model.py:29-38) runs as ONE zs_conv1d_cl
 * call (taps = k along W, stride s, c_in = KH * C): zs_conv2d_gather writes, for segment (b, ho) and padded frame wp, the KH
 * source rows reflect(stride_h * ho + kh - pad_h) side by side, channel kh * C + ci, frame reflect(wp - pad_w); fp16.
 *   src: fp32 (B, H, W) single-channel network input (src_is_f32 = 1, C = 1, out_pitch = 8) or the fp16 channels-last
 *        output of the previous layer, [(b*H + h)*W + w][src_pitch].
 *   stats / inv_count: [B][C] x 2 int64 = (sum, sum of squares) x 2^20 per (b, c) of the source from zs_instnorm2d_stats (fixed
 *        point, so that the split reduction is order-independent) and 1 / (H*W): the source is
 *        normalised on the way (nn.InstanceNorm2d, biased variance, eps 1e-5: model/model.py:139-144); NULL = as is.
 * zs_instnorm2d_stats: stats[b][c] = 2^20 x (sum, sum of squares) over the P rows of sample b of y = fp16 [B][P][pitch].
 * zs_critic_head: conv7 / conv_classify, whose kernel covers the whole remaining map (model/model.py:123-131):
 *   out[b][j] = bias[j] + sum_(p, c) y[b][p][c] * w[j][p][c]   (w fp32 [J][P][C]). */
int zs_conv2d_gather(const void* src, int src_is_f32, int B, int H, int W, int C, int src_pitch, int KH, int stride_h, int pad_h,
                     int pad_w, int Ho, int Wp, const void* stats, float inv_count, void* out, int out_pitch, void* stream);
int zs_instnorm2d_stats(const void* y, int B, long long P, int C, int pitch, void* stats, void* stream);
int zs_critic_head(const void* y, int B, int P, int C, int pitch, const float* w, const float* bias, int J, float* out, void* stream);

/* (B, C, T) fp32 -> channels-last operand buffer [B][rows][pitch] with `halo` reflected rows each side;
 * optional leaky-relu; channels C..pitch-1 are zero-filled. */
int zs_pack_nct(const float* x, int B, int C, int T, void* out, int rows, int pitch, int halo, int choff,
                int lrelu, float ns, int operand, int zero_pad_channels, void* stream);

/* bidirectional GRU recurrence with zero initial state (model/model.py:59-66) on precomputed
 * input projections gx [B][T][2][3H] fp32 (b_ih folded in; rounded to the operand type like the GEMM's output),
 * w_hh [2][3H][H] fp32, b_hh [2][3H];
 * writes h_t (operand type) to out[b][out_halo+t][out_choff + dir*H + j].
 * impl: 0 = what the forward passes use (tensor-core cluster kernel when H % 64 == 0 and H <= 512, else the
 * CUDA-core kernel), 1 = CUDA-core kernel, 2 = cluster kernel. */
int zs_gru_recurrence(const float* gx, const float* w_hh, const float* b_hh, int B, int T, int H,
                      void* out, int out_rows, int out_pitch, int out_halo, int out_choff,
                      int operand, int impl, void* stream);

/* Weight-gradient GEMM of one conv / linear layer on channels-last fp16 buffers (tcgen05, MN-major operands):
 *   grad[co][ci_off + ci][tap0 + j] += scale * sum_{b,t} dy[b][dy_row0 + t][dy_ch0 + co] * x[b][x_row0 + stride*t + j][x_ch0 + ci]
 * for j < taps; grad is (c_out, c_in_total, k) fp32.  T in {8,16,32,64,128,...}: a power of two, or a multiple of 64.
 * ps_c > 0: dy channel m = r*ps_c + c stands for conv output channel 2c + r.
 * grad_is_zero != 0: the caller guarantees the addressed gradient entries are zero on entry; the library may then
 * keep the whole reduction in one CTA and STORE the result instead of combining partial sums with atomic adds. */
typedef struct {
    const void* dy; int32_t dy_rows, dy_pitch, dy_channels, dy_ch0, dy_row0, c_out;
    const void* x; int32_t x_rows, x_pitch, x_channels, x_ch0, x_row0, c_in, stride;
    int32_t B, T, taps;
    float* grad; int32_t c_in_total, ci_off, k, tap0;
    int32_t ps_c;
    float scale;
    int32_t grad_is_zero;
} zs_wgrad_desc;
int zs_wgrad_cl(const zs_wgrad_desc* d, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ZS_AE_H */
