"""Generate tests/golden/sambert_small.npz by running the UNMODIFIED reference
KanTtsSAMBERT (/root/reference, imported through oracle/ref_shims.py) on CPU.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden_sambert.py

The model runs in ``eval()`` (dropout off -- dropout masks depend on the RNG
stream and are not part of the parity contract) with teacher forcing, exactly
the call Sambert_Trainer.train_step makes (train/trainer.py:919-931), followed
by MelReconLoss + ProsodyReconLoss and one backward of their sum.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle.ref_shims import import_reference  # noqa: E402
from make_batch import make_sambert_batch as make_batch  # noqa: E402

import_reference()
from kantts.models.sambert.kantts_sambert import KanTtsSAMBERT  # noqa: E402
from kantts.train.loss import MelReconLoss, ProsodyReconLoss  # noqa: E402

torch.set_num_threads(8)

SMALL_CFG = dict(
    max_len=40, embedding_dim=48, encoder_num_layers=2, encoder_num_heads=2, encoder_num_units=32,
    encoder_ffn_inner_dim=64, encoder_dropout=0.1, encoder_attention_dropout=0.1, encoder_relu_dropout=0.1,
    encoder_projection_units=8, speaker_units=8, emotion_units=8, predictor_filter_size=5,
    predictor_fsmn_num_layers=2, predictor_num_memory_units=16, predictor_ffn_inner_dim=24, predictor_dropout=0.1,
    predictor_shift=0, predictor_lstm_units=8, dur_pred_prenet_units=[8, 8], dur_pred_lstm_units=8,
    decoder_prenet_units=[16, 16], decoder_num_layers=2, decoder_num_heads=2, decoder_num_units=16,
    decoder_ffn_inner_dim=32, decoder_dropout=0.1, decoder_attention_dropout=0.1, decoder_relu_dropout=0.1,
    outputs_per_step=3, num_mels=8, postnet_filter_size=5, postnet_fsmn_num_layers=2, postnet_num_memory_units=16,
    postnet_ffn_inner_dim=24, postnet_dropout=0.1, postnet_shift=1, postnet_lstm_units=8, MAS=False,
    sy=20, tone=5, syllable_flag=4, word_segment=4, emotion=3, speaker=2,
)


def main():
    torch.manual_seed(1234)
    gen = torch.Generator().manual_seed(1235)
    cfg = SMALL_CFG
    model = KanTtsSAMBERT(cfg).eval()
    with torch.no_grad():
        for n, p in model.named_parameters():
            if p.requires_grad and (n.endswith("bias") or "layer_norm" in n or n.endswith("ln.weight")):
                p.add_(0.1 * torch.randn(p.shape, generator=gen))
    batch = make_batch(cfg, B=3, L=10, gen=gen)
    res = model(batch["inputs_ling"], batch["inputs_emotion"], batch["inputs_speaker"], batch["input_lengths"],
                output_lengths=batch["output_lengths"], mel_targets=batch["mel_targets"],
                duration_targets=batch["duration_targets"], pitch_targets=batch["pitch_targets"],
                energy_targets=batch["energy_targets"])
    l0, l1 = MelReconLoss()(batch["output_lengths"], batch["mel_targets"], res["dec_outputs"], res["postnet_outputs"])
    dl, pl, el = ProsodyReconLoss()(res["valid_inter_lengths"], res["duration_targets"], res["pitch_targets"],
                                    res["energy_targets"], res["log_duration_predictions"], res["pitch_predictions"],
                                    res["energy_predictions"])
    total = l0 + l1 + dl + pl + el
    total.backward()
    arrays = {"sd/" + k: v.detach().numpy().copy() for k, v in model.state_dict().items()}
    arrays.update({"in/" + k: v.numpy() for k, v in batch.items()})
    for k in ("dec_outputs", "postnet_outputs", "log_duration_predictions", "pitch_predictions", "energy_predictions",
              "LR_text_outputs", "LR_emo_outputs", "LR_spk_outputs", "LR_length_rounded"):
        arrays["out/" + k] = res[k].detach().numpy()
    for k in ("enc_slf_attn_lst", "pnca_x_attn_lst", "pnca_h_attn_lst"):
        for i, a in enumerate(res[k]):
            arrays[f"out/{k}.{i}"] = a.detach().numpy()
    arrays["out/band_width"] = np.asarray([res["x_band_width"], res["h_band_width"]])
    arrays["out/losses"] = np.asarray([float(v) for v in (l0, l1, dl, pl, el, total)], dtype=np.float64)
    for n, p in model.named_parameters():
        if p.grad is not None:
            arrays["grad/" + n] = p.grad.numpy().copy()
    path = os.path.join(HERE, "sambert_small.npz")
    np.savez_compressed(path, cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8), **arrays)
    print(f"sambert_small: {os.path.getsize(path) / 1e6:.2f} MB, {len(arrays)} arrays, losses {arrays['out/losses']}")


if __name__ == "__main__":
    main()
