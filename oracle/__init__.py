"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain PyTorch-CPU / numpy, fp32 unless noted) of the one
KAN-TTS hot path this repo accelerates: HiFi-GAN Generator / MultiPeriod- /
MultiScale-Discriminator (incl. the neural-source-filter generator variant), the
mel-spectrogram / STFT losses, the LSGAN losses and the GAN train-step schedule;
the SAM-BERT acoustic model (teacher-forced training step and free-running
inference) in oracle/sambert.py.  Every function cites the reference file:line
it restates (paths relative to the KAN-TTS checkout, /root/reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this package -- as the CHECKER, never as
the thing measured or shipped.  The product package (``kan-tts_b200/``,
importable as ``kantts_b200``) never imports it and fails loudly when its CUDA
library is missing.

Parity pin status (see DESIGN.md section "Oracle"):
  * torch arithmetic (conv, weight_norm, spectral_norm, stft): PINNED -- the
    restatement is checked against golden vectors produced by importing the
    unmodified reference in the build container (tests/golden/make_golden.py).
  * pytorch_wavelets DWT1DForward(db3, zero) and librosa.filters.mel (Slaney):
    third-party packages absent from /root/reference and from this image;
    restated from their published algorithms -> "parity unpinned" for those two
    functions (cross-checked against torchaudio's Slaney filterbank and the
    orthonormality / perfect-reconstruction properties of db3).
"""
