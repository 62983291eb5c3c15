"""CUDA-graph replay of the train step must reproduce the eager step exactly (same kernels, same order)."""
import copy

import pytest
import torch

import kantts_b200 as K
from test_gpu_parity import _small_config, DEV

import os

pytestmark = [pytest.mark.gpu]


def _build(g, cfg, graph):
    torch.manual_seed(0)
    model, opt, sched = K.hifigan_model_builder(cfg, DEV)
    model["generator"].load_state_dict(g.group("before/g/"))
    model["discriminator"]["MultiScaleDiscriminator"].load_state_dict(g.group("before/msd/"))
    model["discriminator"]["MultiPeriodDiscriminator"].load_state_dict(g.group("before/mpd/"))
    crit = K.criterion_builder(cfg, DEV)
    return K.GanStep(model, opt, sched, crit, cfg, cuda_graph=graph, graph_warmup=2), model


def test_cuda_graph_step_matches_eager(golden):
    g = golden("trainstep_small")
    cfg = _small_config(g)
    y, x = g.t("y").to(DEV), g.t("x").to(DEV)
    batches = [(y, x), (y.flip(0), x.flip(0)), ((y * 0.5).contiguous(), x), (y, (x * 0.9).contiguous()),
               (y.roll(7, -1), x), (y, x)]
    eager, m_e = _build(g, cfg, False)
    traj_e = [K.train.losses_to_float(eager.step(b)) for b in batches]
    torch.cuda.synchronize()
    graph, m_g = _build(g, cfg, True)
    traj_g = [K.train.losses_to_float(graph.step(b)) for b in batches]   # steps 0-1 eager warm-up, 2 capture, 3+ replay
    for i, (le, lg) in enumerate(zip(traj_e, traj_g)):
        for k in le:
            # two independent trajectories: fp32-atomic summation order differs run to run and a GAN step amplifies
            # it (the feature-matching value most of all), so only the first replayed steps are compared tightly;
            # replay bugs (stale weight / input buffers) show up as O(1e-1) differences that keep growing
            tol = (3e-3 if i <= 3 else 3e-2) * (5.0 if k == "feature_matching_loss" else 1.0)
            assert abs(le[k] - lg[k]) <= tol * max(1.0, abs(le[k])), (i, k, le[k], lg[k])
    assert graph._graphs is not None
    for k, v in m_e["generator"].state_dict().items():
        w = m_g["generator"].state_dict()[k]
        assert float((v - w).abs().max()) <= 5e-3 * max(1.0, float(v.abs().max())), k   # fp32 atomics reorder between runs
