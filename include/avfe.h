/*
 * avfe.h — C ABI of libavfe.so, the B200 (sm_100a) audio-visual front-end.
 *
 * The reference (hhoangphuoc/AVSL) has no FFI: its hot path is plain Python functions that
 * call torch / OpenCV / scikit-image on the CPU.  Each entry point below replaces the
 * arithmetic of one (or a fused run of) those functions; the reference symbol is cited as
 * file:line into the reference tree.  `avsl_b200/*.py` is the host-side mirror that keeps the
 * reference names and signatures and calls these symbols through ctypes; INTEGRATION.md
 * shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer (cudaMalloc'd on the current device), row-major,
 *    caller-allocated and caller-owned; inputs are const and never written.  One documented
 *    exception: `frames` of avfe_lip_roi_batch / avfe_lip_roi_collate may point to page-locked
 *    HOST memory (cudaHostAlloc / cudaHostRegister: mapped into the device's address space under
 *    unified addressing) when gray_out is NULL -- the kernel then pulls only each frame's ROI
 *    footprint across PCIe ("zero-copy"), about a seventh of the frame bytes;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *    stream); no call synchronises the host, allocates device memory or keeps state, so
 *    calls are re-entrant and may be issued concurrently from several host threads;
 *  - every function returns 0 on success or a negative avfe_status; nothing throws across
 *    the ABI and nothing falls back to the CPU.
 */
#ifndef AVFE_H_
#define AVFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define AVFE_API __declspec(dllexport)
#else
#define AVFE_API __attribute__((visibility("default")))
#endif

typedef void* avfe_stream_t; /* cudaStream_t */

enum avfe_status {
  AVFE_OK = 0,
  AVFE_ERR_INVALID_ARG = -1,   /* NULL pointer, negative size, unsupported shape */
  AVFE_ERR_UNSUPPORTED = -2,   /* parameter combination not implemented (e.g. n_mels > 128) */
  AVFE_ERR_WORKSPACE = -3,     /* workspace NULL / too small / misaligned */
  AVFE_ERR_CUDA = -4,          /* a CUDA runtime call or launch failed (see cudaGetLastError) */
  AVFE_ERR_ALIGNMENT = -5      /* a pointer does not meet the documented alignment */
};

enum avfe_fuse_mode { AVFE_FUSE_CONCAT = 0, AVFE_FUSE_SUM = 1, AVFE_FUSE_WSUM = 2 };
enum avfe_dtype { AVFE_F32 = 0, AVFE_F16 = 1, AVFE_BF16 = 2 };

/* library version: major*10000 + minor*100 + patch */
AVFE_API int avfe_version(void);
/* static string for an avfe_status (never NULL) */
AVFE_API const char* avfe_strerror(int status);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
AVFE_API uint64_t avfe_launch_count(void);

/* ------------------------------------------------------------------ audio (A1..A3) */

/* whisper.pad_or_trim(array, length) on the last axis — call site
 * avsl/whisper_flamingo_ft_ami.py:209-210.  in [B, L_in] -> out [B, L_out], zero padded. */
AVFE_API int avfe_pad_or_trim_f32(const float* in, int64_t B, int64_t L_in, int64_t L_out,
                                  float* out, avfe_stream_t stream);

/* Same, for a ragged batch stored back to back: clip b = in[offsets[b] : offsets[b+1]]
 * (offsets int64 [B+1], device) -> out [B, L_out]. */
AVFE_API int avfe_pad_or_trim_ragged_f32(const float* in, const int64_t* offsets, int64_t B,
                                         int64_t L_out, float* out, avfe_stream_t stream);

/* preprocess_audio_for_whisper's peak normalisation, preprocess/audio_process.py:312-317
 * (same code at :291-293): per clip, if max > 1 or min < -1 divide by max(|max|,|min|).
 * audio [B, L] -> out [B, L] (out may alias audio).  scratch: >= 2*B floats. */
AVFE_API int avfe_peak_normalize_f32(const float* audio, int64_t B, int64_t L, float* out,
                                     float* scratch, avfe_stream_t stream);

/* whisper.log_mel_spectrogram(audio, n_mels, padding) — call site
 * avsl/whisper_flamingo_ft_ami.py:212-213 (HF twin avsl/whisper_ft.py:347-350):
 * zero-pad `padding` samples, reflect-centre, Hann(400) STFT hop 160, drop the last frame,
 * |.|^2, mel_filters @, clamp 1e-10, log10, per-CLIP max-8 floor, (x+4)/4.
 *   audio       [B, L] float32
 *   mel_filters [n_mels, 201] float32 (n_mels <= 128)
 *   out         [B, n_mels, (L+padding)/160] float32
 *   workspace   avfe_logmel_workspace_bytes(...) bytes, 16-byte aligned
 * Requires L + padding >= 201 (reflect padding), like torch.stft. */
AVFE_API size_t avfe_logmel_workspace_bytes(int64_t B, int64_t L, int64_t padding, int n_mels);
AVFE_API int avfe_logmel_f32(const float* audio, int64_t B, int64_t L, int64_t padding,
                             int n_mels, const float* mel_filters, float* out,
                             void* workspace, size_t workspace_bytes, avfe_stream_t stream);

/* The same with the filterbank analysed once: avfe_logmel_prepare turns the dense [n_mels,201]
 * matrix into the sparse form the kernel uses (supports, quad-packed weights, balanced work
 * table) in `pack` (avfe_logmel_pack_bytes() bytes, 16-byte aligned, device memory);
 * avfe_logmel_prepared_f32 then skips that analysis on every call.  `mel_filters` must be the
 * matrix `pack` was prepared from. */
AVFE_API size_t avfe_logmel_pack_bytes(void);
AVFE_API int avfe_logmel_prepare(const float* mel_filters, int n_mels, void* pack,
                                 avfe_stream_t stream);
AVFE_API int avfe_logmel_prepared_f32(const float* audio, int64_t B, int64_t L, int64_t padding,
                                      int n_mels, const float* mel_filters, const void* pack,
                                      float* out, void* workspace, size_t workspace_bytes,
                                      avfe_stream_t stream);

/* pad_or_trim + log_mel_spectrogram fused for a ragged batch stored back to back
 * (avsl/whisper_flamingo_ft_ami.py:209-213 in one call): clip b = audio[offsets[b] : offsets[b+1]],
 * cut or zero-extended to `length` samples, then the log-mel of that.  The zero extension is
 * never materialised and frames that lie entirely in it are not even read.  `pack` may be NULL
 * (filters analysed on the fly).  out [B, n_mels, length/160]; workspace as for
 * avfe_logmel_f32 with L = length, padding = 0. */
AVFE_API int avfe_logmel_ragged_f32(const float* audio, const int64_t* offsets, int64_t B,
                                    int64_t length, int n_mels, const float* mel_filters,
                                    const void* pack, float* out, void* workspace,
                                    size_t workspace_bytes, avfe_stream_t stream);

/* SpecAugment masks on a log-mel batch, in place — the spec_augment call of
 * AmiVideoHFDataset.__getitem__, avsl/whisper_flamingo_ft_ami.py:216-224 (upstream
 * whisper_flamingo.spec_augment, policies "ls-double" / "ls-basic").  The rectangles are drawn on
 * the host; bands [B, n_bands, 4] int32 = (f0, f1, t0, t1): mel[b, f0:f1, t0:t1] = fill (empty or
 * out-of-range parts are clipped / ignored).  mel [B, n_mels, n_frames] float32. */
AVFE_API int avfe_spec_mask_f32(float* mel, int64_t B, int n_mels, int64_t n_frames, const int32_t* bands,
                                int n_bands, float fill, avfe_stream_t stream);

/* SpecAugment's time-warping step (Park et al. 2019) for the same call site.  RNG contract of both
 * SpecAugment entry points: the library draws NOTHING -- mask rectangles and warp points are INPUTS
 * (drawn on the host by avsl_b200.audio.spec_augment_bands / spec_augment_warp_points, whose sampling
 * follows the paper; the upstream sampler is un-vendored, so its random stream is not reproduced).
 *   warp [B, 3] int32 = (tau, center, warped): the first tau frames of clip b are warped so that source
 *   frame `center` lands on frame `warped`, both sides resampled linearly (definition in
 *   avfe_specaug.cu); rows outside 0 < center, warped < tau copy the clip unchanged.
 *   mel -> out, both [B, n_mels, n_frames] float32, out != mel. */
AVFE_API int avfe_spec_time_warp_f32(const float* mel, int64_t B, int n_mels, int64_t n_frames, const int32_t* warp,
                                     float* out, avfe_stream_t stream);

/* AV-HuBERT audio features — extract_logfbank_features + audio_to_tensor,
 * preprocess/audio_process.py:152-197 (and utils/data_loading.py:181-201):
 * python_speech_features.logfbank(audio, samplerate=16000) = pre-emphasis 0.97, 400-sample frames
 * every 160 (rectangular window, zero-padded tail, 1 + ceil((L-400)/160) frames), 512-point power
 * spectrum / 512, `nfilt` (26) triangular mel filters, zeros -> float64 eps, natural log; then
 * `stack` (4) consecutive frames per row (zero rows appended to a multiple of `stack`) and, if
 * `normalize`, (x - mean) / (std + 1e-5) per row with the population std.
 *   audio       packed clips, clip b = audio[offsets[b] : offsets[b+1]] (float32, 16 kHz)
 *   row_offsets [B+1] int64: first output row of clip b (clip b has
 *               ceil(avfe_logfbank_num_frames(L_b) / stack) rows)
 *   max_samples the longest clip (sizes the launch)
 *   fbank       [nfilt, 257] float32 (python_speech_features.get_filterbanks), nfilt <= 40
 *   out         [row_offsets[B], nfilt * stack] float32
 *   workspace   avfe_logfbank_workspace_bytes() bytes (filter supports) */
AVFE_API int64_t avfe_logfbank_num_frames(int64_t n_samples);
AVFE_API size_t avfe_logfbank_workspace_bytes(void);
AVFE_API int avfe_logfbank_f32(const float* audio, const int64_t* offsets, const int64_t* row_offsets,
                               int64_t B, int64_t max_samples, const float* fbank, int nfilt,
                               int stack, int normalize, float* out, void* workspace,
                               size_t workspace_bytes, avfe_stream_t stream);
/* The same in two steps, for callers that keep one filterbank across calls (like avfe_logmel_prepare /
 * avfe_logmel_prepared_f32): avfe_logfbank_prepare turns `fbank` into its sparse form (supports, packed
 * weights, the deal of the filters to the kernel's warps) in `workspace` once;
 * avfe_logfbank_prepared_f32 then runs the feature kernel alone -- `workspace` must hold the result
 * of a prepare call for the same `fbank` / `nfilt`. */
AVFE_API int avfe_logfbank_prepare(const float* fbank, int nfilt, void* workspace, size_t workspace_bytes,
                                   avfe_stream_t stream);
AVFE_API int avfe_logfbank_prepared_f32(const float* audio, const int64_t* offsets, const int64_t* row_offsets,
                                        int64_t B, int64_t max_samples, const float* fbank, int nfilt,
                                        int stack, int normalize, float* out, const void* workspace,
                                        size_t workspace_bytes, avfe_stream_t stream);

/* SNR noise mixing — add_noise(clean_wav, noise_wav, snr), preprocess/audio_process.py:110-150
 * (the optional augmentation of process_audio_for_av_hubert, :222-224, in front of the logfbank
 * features), for a packed batch and BIT-EXACT with the reference's float32 numpy arithmetic
 * (its pairwise sums of squares are re-created in the same order):
 *   noise'[i] = noise_b[i mod Ln], i < Lc;  gain = (rms(clean_b) / snr_ratio[b]) / rms(noise')
 *   mixed = clean_b + noise' * gain;  if max > 32767 or min < -32768: mixed *= 32767 / max when
 *   max >= |min|, else -32768 / min;  int16 by truncation toward zero.
 *   clean / noise   packed float32 waveforms (the reference's .astype(np.float32) is the
 *                   caller's; `clean` 16-byte aligned), clip b = clean[clean_offsets[b] : clean_offsets[b+1]], noise
 *                   likewise; offsets are [B+1] int64 on the device
 *   snr_ratio       [B] float32 = (float)10^(snr_b / 20), evaluated by the caller in double like
 *                   the reference's Python expression
 *   max_len         the longest clean clip (sizes the launch and the workspace), < 2^31; a clip
 *                   longer than this is mixed with a wrong gain (never out of bounds)
 *   out_i16/out_f32 packed like `clean` (8- / 16-byte aligned); either may be NULL; out_f32 holds the same integers as
 *                   float32 (what logfbank takes)
 * A clip whose noise is empty is cast unmixed (the reference raises ZeroDivisionError; the Python
 * shim raises too).  NaN / inf samples and a silent noise clip are outside the contract, as in
 * the reference. */
AVFE_API size_t avfe_add_noise_workspace_bytes(int64_t B, int64_t max_len);
AVFE_API int avfe_add_noise(const float* clean, const int64_t* clean_offsets, const float* noise,
                            const int64_t* noise_offsets, const float* snr_ratio, int64_t B,
                            int64_t max_len, int16_t* out_i16, float* out_f32, void* workspace,
                            size_t workspace_bytes, avfe_stream_t stream);

/* ------------------------------------------------------------------ video (V1..V8) */

/* cv2.cvtColor(frame, COLOR_BGR2GRAY) — preprocess/video_process.py:201-214:
 * Y = (3735*B + 19235*G + 9798*R + 16384) >> 15.  bgr [N,H,W,3] u8 -> gray [N,H,W] u8. */
AVFE_API int avfe_bgr2gray_u8(const uint8_t* bgr, int64_t N, int H, int W, uint8_t* gray,
                              avfe_stream_t stream);

/* skimage.transform.warp(img, inverse_map=tform.inverse, output_shape) followed by
 * (*255).astype(uint8) — utils/lips_cropping.py:104-107 (warp_img) and :122-124
 * (apply_transform) — for a GIVEN transform.  gray [H,W] u8; inv_matrix = 3x3 row-major
 * float64 of tform.inverse.params; out [out_h,out_w] u8.  Bit-exact float64 evaluation. */
AVFE_API int avfe_warp_affine_u8(const uint8_t* gray, int H, int W, const double* inv_matrix,
                                 int out_h, int out_w, uint8_t* out, avfe_stream_t stream);

/* The fused lip-ROI path for a batch of clips: gray conversion (V1), landmark
 * interpolation (V2, utils/lips_cropping.py:41-89), 12-frame forward window smoothing and
 * tail reuse (V3, preprocess/video_process.py:369-370,417-475), similarity fit on the 5
 * stable points to the mean face (V4, utils/lips_cropping.py:104), warp to std_size (V4/V5),
 * landmark transform (V6), cut_patch (V7, utils/lips_cropping.py:127-163), centre crop +
 * normalise (V8, utils/hf_video_utils.py:113-138).
 *
 *   frames       [N, H, W, channels] u8, channels = 3 (BGR) or 1 (already gray); N = total
 *                frames of all clips, clips stored back to back
 *   clip_offsets [n_clips+1] int64 (device): clip k owns frames [off[k], off[k+1])
 *   landmarks    [N, 68, 2] float64 (x, y) in frame pixels
 *   lm_valid     [N] u8 or NULL (= all valid); 0 marks a failed detection (reference: None)
 *   mean_face    [68, 2] float64 (resources/20words_mean_face.npy)
 *   tforms_in    [N, 18] float64 or NULL: per frame the forward 3x3 (tform.params) followed by
 *                the inverse 3x3 (tform.inverse.params), both affine.  When given, the transform
 *                is taken from here instead of being fitted (apply_transform semantics) and the
 *                warp is bit-exact for those matrices.
 *   std_size 300, roi 96 (even), crop 88 (<= roi, same parity), window 12
 *   mean/std     normalisation constants (0.421 / 0.165)
 * outputs (each nullable):
 *   gray_out [N,H,W] u8          bit-exact cv2 gray frames (channels==3 only)
 *   lip_u8   [N,roi,roi] u8      what extract_lip_frames returns
 *   lip_f32  [N,crop,crop] f32   what load_video_feats returns ([T,88,88,1])
 *   crop_rc  [N,2] int32         row/col of the ROI's top-left corner in the std frame
 *   tforms   [N,18] float64      forward 3x3 then inverse 3x3 used for each frame
 * A clip without any valid landmark yields zero ROIs and crop_rc = -1 (the host shim maps
 * that to the reference's empty-array return).
 *   workspace avfe_lip_workspace_bytes(N) bytes, 16-byte aligned. */
AVFE_API size_t avfe_lip_workspace_bytes(int64_t N);
AVFE_API int avfe_lip_roi_batch(const uint8_t* frames, int channels, int64_t N, int H, int W,
                                const int64_t* clip_offsets, int64_t n_clips,
                                const double* landmarks, const uint8_t* lm_valid,
                                const double* mean_face, const double* tforms_in,
                                int std_size, int roi, int crop, int window,
                                float mean, float std,
                                uint8_t* gray_out, uint8_t* lip_u8, float* lip_f32,
                                int32_t* crop_rc, double* tforms,
                                void* workspace, size_t workspace_bytes, avfe_stream_t stream);

/* The same path writing straight into the padded batch the Whisper-Flamingo encoder consumes,
 * i.e. what AmiVideoHFDataset.__getitem__'s trim (avsl/whisper_flamingo_ft_ami.py:299-302:
 * video_feats[:round(len(audio)/16000*25)]) and the upstream WhisperVideoCollatorWithPadding
 * (imported at :126, used at :686; zero-pads every clip to the longest one and returns a boolean
 * padding mask, read at :504) produce on the host:
 *   keep_frames  [n_clips] int64 (device) or NULL: frames kept per clip after the trim
 *   T_pad        frames per clip in the padded batch (>= 1)
 *   video        [n_clips, T_pad, crop, crop] f32  == the collator's [B, 1, T, 88, 88]; frame t of
 *                clip c is written for t < min(T_c, keep_frames[c], T_pad), zero otherwise
 *   padding_mask [n_clips, T_pad] u8 (nullable), 1 = padded frame
 *   gray_out     [N,H,W] u8 (nullable) as above (every frame, trimmed ones included)
 * The ROI of a dropped frame is not computed.  Workspace as for avfe_lip_roi_batch. */
AVFE_API int avfe_lip_roi_collate(const uint8_t* frames, int channels, int64_t N, int H, int W,
                                  const int64_t* clip_offsets, int64_t n_clips,
                                  const double* landmarks, const uint8_t* lm_valid,
                                  const double* mean_face, const double* tforms_in,
                                  int std_size, int roi, int crop, int window,
                                  float mean, float std,
                                  const int64_t* keep_frames, int64_t T_pad,
                                  uint8_t* gray_out, float* video, uint8_t* padding_mask,
                                  void* workspace, size_t workspace_bytes, avfe_stream_t stream);

/* landmarks_interpolate — utils/lips_cropping.py:41-89 — alone: fills frames whose lm_valid is
 * 0 by linear interpolation between the neighbouring detections of the same clip
 * (start + idx/float(n) * delta) and by replication at the clip ends; a clip with no
 * detection at all is filled with NaN (reference: returns None).  out [N,68,2] float64. */
AVFE_API int avfe_landmarks_interpolate(const double* landmarks, const uint8_t* lm_valid,
                                        const int64_t* clip_offsets, int64_t n_clips, int64_t N,
                                        double* out, avfe_stream_t stream);

/* skimage.transform.estimate_transform('similarity', src, dst) — utils/lips_cropping.py:104 —
 * for n 2-D points (closed form of skimage's _umeyama).  out18 = forward 3x3 then inverse 3x3,
 * float64 row-major. */
AVFE_API int avfe_similarity_fit(const double* src, const double* dst, int n, double* out18,
                                 avfe_stream_t stream);

/* cut_patch(img, landmarks, height, width) — utils/lips_cropping.py:127-163 — on a 2-D uint8
 * image: centre = mean of the n landmarks (x, y), clamped so the patch fits, Python round().
 * out [2*half_h, 2*half_w] u8; rc (nullable) receives the chosen top-left (row, col). */
AVFE_API int avfe_cut_patch_u8(const uint8_t* img, int H, int W, const double* landmarks, int n,
                               int half_h, int half_w, uint8_t* out, int32_t* rc,
                               avfe_stream_t stream);

/* load_video_feats arithmetic on an existing ROI stack — utils/hf_video_utils.py:113-138,
 * utils/data_loading.py:46-66,101-118: /255, centre crop, (x-mean)/std in float32.
 * roi_u8 [N, Hin, Win] u8 -> out [N, crop, crop] f32 (crop <= min(Hin,Win)). */
AVFE_API int avfe_video_feats_u8(const uint8_t* roi_u8, int64_t N, int Hin, int Win, int crop,
                                 float mean, float std, float* out, avfe_stream_t stream);

/* load_video_feats_from_decord_reader as its call site runs it — utils/hf_video_utils.py:103-138,
 * called from safe_load_video_feats_from_hf_object (avsl/whisper_flamingo_ft_ami.py:279-286) with what
 * decord returns, uint8 [T,H,W,3] RGB:
 *   channels 3: gray = dot(rgb, [0.2989, 0.5870, 0.1140]) in float64 (:105); if the maximum over the
 *               whole stack is > 1.0, float32(gray) / 255 (:116-117), else the float64 values stay
 *               unscaled (an all-dark video; the reference's own data-dependent branch)
 *   channels 1: float32(u8) / 255 (:114-115)
 *   H, W >= crop: centre crop (:120-125); otherwise cv2.resize(frame, (crop, crop)), INTER_LINEAR
 *               on the float frame (:126-132; OpenCV's generic code path, see avfe_vfeats.cu)
 *   (x - mean) / std in the array's dtype (:135), result float32 (ft_ami:286).
 * frames [N,H,W,channels] u8 -> out [N,crop,crop] f32.  mean / std are doubles because the
 * all-dark branch evaluates them in float64.  workspace: avfe_video_feats_workspace_bytes() bytes,
 * 16-byte aligned (channels 3 only; the stack-wide maximum test runs on the device, no host sync). */
AVFE_API size_t avfe_video_feats_workspace_bytes(void);
AVFE_API int avfe_video_feats(const uint8_t* frames, int channels, int64_t N, int H, int W, int crop,
                              double mean, double std, float* out, void* workspace,
                              size_t workspace_bytes, avfe_stream_t stream);

/* ------------------------------------------------------------------ fusion (F1..F3) */

/* AVHuBERTEncoderWrapper.forward's fusion block — avsl/modules/av_hubert_encoder.py:315-326,
 * with the modality-dropout decision (:292-298) supplied as a per-sample mask.
 *   fa, fv [B, C, T] of `dtype`; mask [B,2] u8 (col 0 audio present, col 1 video present) or
 *   NULL (= all present); a missing modality is zero-filled and NOT read.
 *   mode CONCAT -> out [B, 2C, T]; SUM -> fa+fv [B, C, T]; WSUM -> w_a*fa + w_v*fv [B, C, T]. */
AVFE_API int avfe_fuse(const void* fa, const void* fv, const uint8_t* mask, int mode,
                       float w_a, float w_v, int dtype, int64_t B, int64_t C, int64_t T,
                       void* out, avfe_stream_t stream);

/* The same fusion followed by the rest of the block's tail in one pass —
 * avsl/modules/av_hubert_encoder.py:329-330: features.transpose(1, 2) then self.layer_norm
 * (nn.LayerNorm over the fused channel dimension, evaluated in float32 and cast back to the
 * input dtype, avsl/modules/av_hubert_layers.py:438-440).  The fused [B, C', T] tensor and its
 * transposed copy are never materialised.
 *   gamma, beta [C'] float32 (LayerNorm weight / bias; NULL = 1 / 0), eps (1e-5 in the reference)
 *   out [B, T, C'] of `dtype`, C' = 2C for CONCAT, C otherwise. */
AVFE_API int avfe_fuse_layernorm(const void* fa, const void* fv, const uint8_t* mask, int mode,
                                 float w_a, float w_v, int dtype, int64_t B, int64_t C, int64_t T,
                                 const float* gamma, const float* beta, float eps, void* out,
                                 avfe_stream_t stream);

/* avfe_fuse_layernorm for feature maps whose time rows are `t_pitch` elements apart (t_pitch >= T).
 * When the rows keep 16-byte alignment (t_pitch * sizeof(dtype) % 16 == 0 -- T = 750 in a [B, C, 752]
 * allocation) and C is a multiple of 256, the tile fill is done by tensor-map TMA (avfe_fuse_ln_tma.cu:
 * the kernel is no longer bound by the SM's load/store pipe); avfe_fuse_layernorm_tma_ok tells.  A
 * contiguous layout that does not qualify falls through to avfe_fuse_layernorm; a padded one that
 * does not returns AVFE_ERR_UNSUPPORTED. */
AVFE_API int avfe_fuse_layernorm_tma_ok(int dtype, int64_t C, int64_t T, int64_t t_pitch);
AVFE_API int avfe_fuse_layernorm_pitched(const void* fa, const void* fv, const uint8_t* mask, int mode, float w_a,
                                         float w_v, int dtype, int64_t B, int64_t C, int64_t T, int64_t t_pitch,
                                         const float* gamma, const float* beta, float eps, void* out,
                                         avfe_stream_t stream);

/* post_extract_proj fused behind concat + transpose + LayerNorm -- avsl/modules/av_hubert_encoder.py:315-334:
 *     features = self.post_extract_proj(self.layer_norm(cat([fa, fv], 1).transpose(1, 2)))
 * (nn.Linear(2C -> D) where self.embed != encoder_embed_dim, :160-166; fp16 / bf16 under `precision: 16`).
 * The contraction runs on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands
 * staged by tensor-map TMA); LayerNorm is folded around it (avfe_pep.cu), so the fused, transposed and
 * normalised tensors are never materialised.
 *
 * avfe_proj_fold, once per set of weights: W [D, 2C] (w_dtype: AVFE_F32 master weights, or the GEMM
 * dtype), gamma / beta [2C] float32 (LayerNorm weight / bias, NULL = 1 / 0), bias [D] float32 (NULL = 0)
 * -> `folded` (avfe_proj_fold_bytes(D, 2C) bytes, 16-byte aligned): W' = gamma * W in `dtype`,
 * s = sum_k W', c = W beta + bias.
 *
 * avfe_fuse_ln_proj: fa, fv [B, C, T] of `dtype` (AVFE_F16 / AVFE_BF16) whose rows are `t_pitch`
 * elements apart (t_pitch >= T, a multiple of 8: TMA needs 16-byte aligned rows -- T = 750 lives in a
 * [B, C, 752] allocation); C a multiple of 64, D a multiple of 256; mask [B,2] u8 or NULL as for avfe_fuse
 * (the K blocks of a missing modality are skipped).  out [B, T, D] of `dtype`.
 * workspace: avfe_fuse_ln_proj_workspace_bytes(B, T) bytes (the LayerNorm moments). */
AVFE_API size_t avfe_proj_fold_bytes(int64_t D, int64_t K);
AVFE_API int avfe_proj_fold(const void* W, int w_dtype, const float* gamma, const float* beta, const float* bias,
                            int64_t D, int64_t K, int dtype, void* folded, avfe_stream_t stream);
AVFE_API size_t avfe_fuse_ln_proj_workspace_bytes(int64_t B, int64_t T);
AVFE_API int avfe_fuse_ln_proj(const void* fa, const void* fv, const uint8_t* mask, int dtype, int64_t B, int64_t C,
                               int64_t T, int64_t t_pitch, const void* folded, int64_t D, float eps, void* out,
                               void* workspace, size_t workspace_bytes, avfe_stream_t stream);

/* ------------------------------------------------------------------ fusion, backward
 * The reference's fusion block runs inside the TRAINING forward (modality dropout under
 * self.training, avsl/modules/av_hubert_encoder.py:292-330), so a drop-in has to pass gradients to
 * the feature extractors and to layer_norm.weight / .bias.  The Python shim wraps these in
 * torch.autograd.Function (avsl_b200/fusion.py). */

/* d(avfe_fuse)/d(fa, fv): grad_out [B, 2C, T] (CONCAT) or [B, C, T] -> grad_fa, grad_fv [B, C, T].
 * The gradient of a masked-out (zero-filled) modality is zero; WSUM scales by w_a / w_v. */
AVFE_API int avfe_fuse_backward(const void* grad_out, const uint8_t* mask, int mode, float w_a, float w_v,
                                int dtype, int64_t B, int64_t C, int64_t T, void* grad_fa, void* grad_fv,
                                avfe_stream_t stream);

/* Backward of avfe_fuse_layernorm in one pass: fa, fv, mask, mode, weights, gamma, eps as in the
 * forward call (the moments are recomputed, nothing has to be saved); grad_out [B, T, C'] of `dtype`.
 *   grad_fa, grad_fv [B, C, T] of `dtype`; grad_gamma, grad_beta [C'] float32 (nullable)
 *   workspace avfe_fuse_layernorm_backward_workspace_bytes(...) bytes, 16-byte aligned (per-CTA
 *   partial sums of grad_gamma / grad_beta, added in a fixed order: the result is deterministic). */
AVFE_API size_t avfe_fuse_layernorm_backward_workspace_bytes(int64_t B, int64_t C, int64_t T, int mode);
AVFE_API int avfe_fuse_layernorm_backward(const void* fa, const void* fv, const uint8_t* mask, int mode,
                                          float w_a, float w_v, int dtype, int64_t B, int64_t C, int64_t T,
                                          const float* gamma, float eps, const void* grad_out, void* grad_fa,
                                          void* grad_fv, float* grad_gamma, float* grad_beta, void* workspace,
                                          size_t workspace_bytes, avfe_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AVFE_H_ */
