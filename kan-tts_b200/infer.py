"""End-to-end inference without the reference's ``.npy`` hand-off (BASELINE configs[4]):
symbols -> SAM-BERT free-running decode -> mel stays on the device -> HiFi-GAN generator -> waveforms.

Reference flow: kantts/bin/infer_sambert.py:205-224 writes ``<utt>_mel.npy`` (the post-net mel, one utterance per
forward because its decoder masks only support batch 1), kantts/bin/infer_hifigan.py:112-124 loads it, transposes to
(1, C, T) and runs ``Generator`` with weight norm removed.  Here a whole batch of utterances goes through both models
in one call; every utterance is cut at its own predicted length (frames x product of the up-sampling scales)."""
import numpy as np
import torch


@torch.no_grad()
def synthesize(sambert, generator, inputs_ling, inputs_emotion, inputs_speaker, input_lengths):
    """sambert: ``KanTtsSAMBERT`` in eval(); generator: ``Generator`` in eval() (``remove_weight_norm()`` optional --
    the prepared weights are cached either way).  Tensors as in ``KanTtsSAMBERT.forward`` (inference branch).
    -> (list of 1-D waveform tensors, dict of the acoustic-model results)."""
    if sambert.training or generator.training:
        raise RuntimeError("synthesize() expects both models in eval() mode")
    res = sambert(inputs_ling, inputs_emotion, inputs_speaker, input_lengths)
    mel = res["postnet_outputs"]                                   # (B, T, num_mels), zero beyond each length
    frames = res["LR_length_rounded"]
    wav = generator(mel.transpose(1, 2).contiguous())              # (B, 1, T * hop)
    hop = int(np.prod(generator.upsample_scales))
    wavs = [wav[b, 0, : int(frames[b]) * hop] for b in range(wav.shape[0])]
    return wavs, res
