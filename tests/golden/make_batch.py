"""Synthetic SAM-BERT batches (reference-free, importable on the GPU box).  Shared by the golden generator
(make_golden_sambert.py), the GPU parity tests and bench/scripts."""
import torch


def make_sambert_batch(cfg, B, L, gen, short=2):
    """Teacher-forcing batch shaped like the collate output of the reference dataset: durations of the padded
    symbols are 0, every row's durations sum to its output length and the padded mel length is a multiple of
    outputs_per_step.  Keys = KanTtsSAMBERT.forward keyword names."""
    r = cfg["outputs_per_step"]
    ling = torch.stack([torch.randint(0, cfg[k], (B, L), generator=gen)
                        for k in ("sy", "tone", "syllable_flag", "word_segment")], -1)
    emo = torch.randint(0, cfg["emotion"], (B, L), generator=gen)
    spk = torch.randint(0, cfg["speaker"], (B, L), generator=gen)
    in_len = torch.tensor([L - short * (i % 2) for i in range(B)])
    dur = torch.randint(1, 5, (B, L), generator=gen)
    dur = dur * (torch.arange(L)[None, :] < in_len[:, None])
    # make row 0 the longest and its length a multiple of r
    dur[0, 0] += (-int(dur[0].sum())) % r + r
    while B > 1 and int(dur[0].sum()) < int(dur[1:].sum(1).max()):
        dur[0, 0] += r
    out_len = dur.sum(1)
    T = int(out_len.max())
    assert T % r == 0 and int(out_len[0]) == T
    return dict(
        inputs_ling=ling, inputs_emotion=emo, inputs_speaker=spk, input_lengths=in_len, output_lengths=out_len,
        mel_targets=torch.randn(B, T, cfg["num_mels"], generator=gen), duration_targets=dur,
        pitch_targets=torch.randn(B, L, generator=gen), energy_targets=torch.randn(B, L, generator=gen))


def make_c4_batch(cfg, gen, B=32, L=256, dur=3):
    """SURVEY.md section 8d config C4 (BASELINE configs[3]): batch 32, 256 symbols of which 255 valid, every
    duration 3 -> 768 mel frames, randn mel / pitch / energy targets.  Keys = the reference collate names read by
    Sambert_Trainer.train_step (train/trainer.py:899-913)."""
    ling = torch.stack([torch.randint(0, cfg[k], (B, L), generator=gen)
                        for k in ("sy", "tone", "syllable_flag", "word_segment")], -1)
    return dict(
        input_lings=ling, input_emotions=torch.randint(0, cfg["emotion"], (B, L), generator=gen),
        input_speakers=torch.randint(0, cfg["speaker"], (B, L), generator=gen),
        valid_input_lengths=torch.full((B,), L - 1, dtype=torch.long),
        valid_output_lengths=torch.full((B,), L * dur, dtype=torch.long),
        mel_targets=torch.randn(B, L * dur, cfg["num_mels"], generator=gen),
        durations=torch.full((B, L), dur, dtype=torch.long),
        pitch_contours=torch.randn(B, L, generator=gen), energy_contours=torch.randn(B, L, generator=gen))
