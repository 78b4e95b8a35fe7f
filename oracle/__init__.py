"""CPU oracle for the AV front-end hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``avsl_b200/`` imports this package.  It is used by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` as the checker and the timed CPU baseline, never as a product path.

Every function restates the algorithm of one reference symbol (file:line into
``/root/reference`` given in each docstring).  Pinning status per path:

* log-mel      : pinned against ``transformers.WhisperFeatureExtractor`` (the reference's
                 second call site, ``avsl/whisper_ft.py:347-350``) — golden vectors in
                 ``tests/golden/``.  ``whisper_flamingo/whisper/audio.py`` itself is
                 un-vendored upstream code and absent from the reference snapshot.
* BGR->gray    : pinned bit-exact against ``cv2.cvtColor`` (the reference's literal call,
                 ``preprocess/video_process.py:214``).
* crop/normalise: pinned against the reference function itself
                 (``utils/hf_video_utils.py:73-145`` imported by file path when the
                 golden vectors were generated).
* similarity fit + warp + cut_patch: **parity unpinned** — the arithmetic lives in
                 scikit-image (``requirements.txt:15``, version unpinned) which is not
                 installable here; the restatement follows skimage's published
                 ``_umeyama`` / ``_warp_fast`` algorithm and is cross-checked against
                 ``scipy.ndimage.map_coordinates`` and ``cv2.warpAffine``.
* fusion       : concat / add are the reference's two torch expressions
                 (``avsl/modules/av_hubert_encoder.py:315-326``); weighted-sum and
                 per-sample masks are build-defined (**parity unpinned**: the reference
                 raises ``ValueError``).  fusion + transpose + LayerNorm: the reference's own
                 torch ops (pinned).
* SNR noise mixing: pinned bit-exact against the reference function itself
                 (``preprocess/audio_process.py:110-150``, taken out of its module by name and
                 run unmodified when the golden vectors were generated; ``oracle/noise.py``).
* collation, logfbank, SpecAugment sampling: **parity unpinned** (upstream / third-party code
                 absent from the snapshot and the image; see each file's header).
"""
