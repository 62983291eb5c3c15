"""world_size-2 gloo test of the data-parallel gradient exchange (train.FlatGrads): the N>1 path
shards by utterance and all-reduces (mean) one flat buffer per model."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kantts_b200.train import FlatGrads
    torch.manual_seed(0)
    m = torch.nn.Sequential(torch.nn.Linear(8, 4), torch.nn.Linear(4, 2))
    fg = FlatGrads(m)
    x = torch.full((3, 8), float(rank + 1))
    fg.zero()
    m(x).sum().backward()
    local = fg.flat.clone()
    fg.all_reduce_mean()
    # grads are linear in x here: mean over ranks of (rank+1)-scaled weight grads
    torch.save({"local": local, "reduced": fg.flat.clone(), "is_view": m[0].weight.grad.data_ptr() == fg.flat.data_ptr()},
               os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_flat_grad_allreduce_mean_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(world)]
    assert all(x["is_view"] for x in r)
    mean = (r[0]["local"] + r[1]["local"]) / 2
    for x in r:
        assert torch.allclose(x["reduced"], mean, atol=1e-6)
    assert not torch.allclose(r[0]["local"], r[1]["local"])
