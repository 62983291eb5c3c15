"""(test tooling: it calls the oracle, so it lives under tests/)  Per-tensor gradient error of the full-width SAM-BERT on the short ragged batch, both compute paths."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # repo root (this file lives in tests/)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import kantts_b200
from kantts_b200 import sambert
from oracle import sambert as osb
from golden.make_batch import make_sambert_batch
from test_gpu_sambert import _run_model
from conftest import rel_l2

cfg = kantts_b200.sambert_24k_config()
torch.manual_seed(77)
ref = sambert.KanTtsSAMBERT(cfg)
sd = {k: v.clone() for k, v in ref.state_dict().items()}
batch = make_sambert_batch(cfg, B=3, L=24, gen=torch.Generator().manual_seed(78), short=5)
torch.set_num_threads(16)
sdo = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "position_enc" not in k
           and "inv_timescales" not in k else v) for k, v in sd.items()}
want = osb.sambert_forward(sdo, cfg, batch["inputs_ling"], batch["inputs_emotion"], batch["inputs_speaker"],
                           batch["input_lengths"], batch["output_lengths"], batch["mel_targets"],
                           batch["duration_targets"], batch["pitch_targets"], batch["energy_targets"])
total, parts = osb.total_loss(want, batch)
total.backward()
for ffma in (True, False):
    model, res, losses = _run_model(cfg, sd, batch, ffma)
    errs = []
    for k, p in model.named_parameters():
        if p.requires_grad and sdo[k].grad is not None and float(sdo[k].grad.abs().max()) > 1e-7:
            errs.append((rel_l2(p.grad.cpu(), sdo[k].grad), k, float(sdo[k].grad.norm())))
    v = sorted(e[0] for e in errs)
    print(f"== path {'ffma' if ffma else 'tcgen05'}: median {v[len(v)//2]:.2e} max {v[-1]:.2e}")
    for e, k, n in errs[::6]:
        print(f"   {e:.2e}  |g|={n:.2e}  {k}")
