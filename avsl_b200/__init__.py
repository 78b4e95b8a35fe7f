"""avsl_b200 — B200-native (sm_100a) audio-visual speech front-end.

A drop-in for the data front-end of hhoangphuoc/AVSL: Whisper log-mel spectrograms
(``audio``), landmark-driven lip-ROI extraction, crop and normalisation (``lips``) and
AV-HuBERT modality fusion (``fusion``), as hand-written CUDA kernels behind a C ABI
(``include/avfe.h`` -> ``avsl_b200/lib/libavfe.so``).  There is no CPU fallback: every
function raises if the library has not been built or no CUDA device is present.
"""
from . import _lib  # noqa: F401
from .audio import (HOP_LENGTH, N_FFT, N_FRAMES, N_SAMPLES, SAMPLE_RATE, LogfbankPlan, NoisePlan, add_noise, add_noise_batch,
                    align_audio_video_features, aligned_lengths, extract_logfbank_features, process_audio_dual_encoder,
                    spec_augment_warp_points, spec_time_warp,
                    log_mel_spectrogram, log_mel_spectrogram_ragged, logfbank_batch, logfbank_num_frames,
                    mel_filters, pad_or_trim, peak_normalize, process_audio_for_av_hubert, spec_augment,
                    spec_augment_bands)
from .frontend import (AVFrontEnd, HostPipeline, PackedBatch, algorithmic_bytes, bind_to_gpu_numa_node,
                       pack_utterances, shard, shard_balanced)
from .fusion import (FoldedProjection, FusedProjection, ModalityFusion, alloc_features, fuse_layernorm_project, fuse_modalities,
                     fuse_transpose_layernorm, modality_dropout_flags, modality_dropout_mask)
from .lips import (SimilarityTransform, apply_transform, bgr2gray, cut_patch, extract_lip_frames,
                   landmarks_interpolate, lip_roi_batch, lip_roi_collate, load_video_feats, mean_face_landmarks,
                   trim_video_to_audio, video_frames_for_audio, warp_img)

__version__ = "0.1.0"
