"""CUDA-graph replay of the train step must reproduce the eager step exactly (same kernels, same order)."""
import copy

import pytest
import torch

import kantts_b200 as K
from test_gpu_parity import _small_config, DEV

import os

# CUDA-graph replay of the step is EXPERIMENTAL in round 1: it reproduces the eager step on small models but
# cudaGraphLaunch crashes on the full-size C2 graph (profiles/r01_notes.md).  Opt in explicitly.
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("KANTTS_B200_TEST_GRAPH") != "1",
                                 reason="experimental CUDA-graph step: set KANTTS_B200_TEST_GRAPH=1")]


def _build(g, cfg, graph):
    torch.manual_seed(0)
    model, opt, sched = K.hifigan_model_builder(cfg, DEV, capturable=graph)
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
    graph, m_g = _build(g, cfg, True)
    for i, b in enumerate(batches):
        le = K.train.losses_to_float(eager.step(b))
        lg = K.train.losses_to_float(graph.step(b))      # steps 0-1 eager warm-up, 2 capture, 3+ replay
        for k in le:
            # (two independent trajectories: fp32-atomic summation order differs run to run, and a GAN step
            #  amplifies it; replay bugs show up as O(1e-2) differences, see profiles/r01_notes.md)
            assert abs(le[k] - lg[k]) <= 3e-3 * max(1.0, abs(le[k])), (i, k, le[k], lg[k])
    assert graph._graphs is not None
    for k, v in m_e["generator"].state_dict().items():
        w = m_g["generator"].state_dict()[k]
        assert float((v - w).abs().max()) <= 2e-3 * max(1.0, float(v.abs().max())), k   # fp32 atomics reorder between runs
