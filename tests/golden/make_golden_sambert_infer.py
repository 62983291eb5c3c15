"""Generate tests/golden/sambert_small_infer.npz: FREE-RUNNING inference (no targets: predicted pitch / energy /
durations, autoregressive duration predictor, step-by-step PNCA decoding with K/V state) of the UNMODIFIED reference
KanTtsSAMBERT (/root/reference via oracle/ref_shims.py), CPU, eval().  Same small model and symbol inputs as
sambert_small.npz; the duration predictor's output bias is raised so that the predicted durations are 2-5 frames
(a random-init model otherwise predicts ~0 frames per symbol).  Build container only:

    python tests/golden/make_golden_sambert_infer.py
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

import_reference()
from kantts.models.sambert.kantts_sambert import KanTtsSAMBERT  # noqa: E402

torch.set_num_threads(8)


def main():
    z = np.load(os.path.join(HERE, "sambert_small.npz"))
    cfg = json.loads(bytes(z["cfg"]).decode())
    sd = {k[3:]: torch.from_numpy(np.array(z[k])) for k in z.files if k.startswith("sd/")}
    sd["variance_adaptor.duration_predictor.fc.bias"] = torch.tensor([1.25])
    model = KanTtsSAMBERT(cfg).eval()
    model.load_state_dict(sd, strict=True)
    # the reference's free-running decoder only works for batch 1 (its band masks are built for one item,
    # kantts_sambert.py:137-168 with mask=None + sambert/__init__.py:275 repeat): utterance 0 of the fixture
    inputs = {k: torch.from_numpy(np.array(z["in/" + k]))[:1] for k in ("inputs_ling", "inputs_emotion", "inputs_speaker",
                                                                         "input_lengths")}
    with torch.no_grad():
        res = model(inputs["inputs_ling"], inputs["inputs_emotion"], inputs["inputs_speaker"], inputs["input_lengths"])
    dur = torch.exp(res["log_duration_predictions"]) - 1
    frac = (dur + 0.5) - torch.floor(dur + 0.5)
    margin = float(torch.minimum(frac, 1 - frac)[dur > 0].min())
    assert margin > 5e-3, f"a predicted duration sits {margin} from a rounding boundary: change the bias"
    arrays = {"sd/" + k: v.numpy().copy() for k, v in model.state_dict().items()}
    arrays.update({"in/" + k: v.numpy() for k, v in inputs.items()})
    for k in ("dec_outputs", "postnet_outputs", "log_duration_predictions", "pitch_predictions", "energy_predictions",
              "LR_text_outputs", "LR_emo_outputs", "LR_spk_outputs", "LR_length_rounded"):
        arrays["out/" + k] = res[k].numpy()
    for k in ("pnca_x_attn_lst", "pnca_h_attn_lst"):
        for i, a in enumerate(res[k]):
            arrays[f"out/{k}.{i}"] = a.numpy()
    arrays["out/band_width"] = np.asarray([res["x_band_width"], res["h_band_width"]])
    path = os.path.join(HERE, "sambert_small_infer.npz")
    np.savez_compressed(path, cfg=np.frombuffer(json.dumps(cfg).encode(), dtype=np.uint8), **arrays)
    print(f"sambert_small_infer: {os.path.getsize(path) / 1e3:.0f} KB; frames {res['LR_length_rounded'].tolist()}, "
          f"decoder steps {res['dec_outputs'].shape[1] // cfg['outputs_per_step']}, band {res['x_band_width']}, "
          f"rounding margin {margin:.3f}, durations[0] {dur[0].tolist()[:6]}")


if __name__ == "__main__":
    main()
