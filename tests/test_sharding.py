"""Batch sharding + the single count all-reduce, on CPU with the gloo backend and world_size 2."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from convnet_quantization_b200 import sharding, synth
from convnet_quantization_b200.models.baseline_model import SimpleConvNet


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 64, 1000, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    net = SimpleConvNet()
    net.load_state_dict(synth.make_state_dict(0))
    net.eval()
    images = synth.images_f32(n, seed=21)
    labels = torch.randint(0, 10, (n,), generator=torch.Generator().manual_seed(5))
    out[rank] = sharding.sharded_accuracy(net, images, labels, rank, world, batch=17)
    dist.destroy_process_group()


def test_sharded_accuracy_matches_single_process():
    n, world = 101, 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, out), nprocs=world, join=True)
        got = [out[r] for r in range(world)]
    net = SimpleConvNet()
    net.load_state_dict(synth.make_state_dict(0))
    net.eval()
    images = synth.images_f32(n, seed=21)
    labels = torch.randint(0, 10, (n,), generator=torch.Generator().manual_seed(5))
    want = sharding.sharded_accuracy(net, images, labels, 0, 1, batch=64)
    assert got[0] == got[1] == want and want[2] == n


def test_cpulist_parsing_and_numa_binding_are_safe_without_a_gpu():
    from convnet_quantization_b200 import sharding
    assert sharding.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding.parse_cpulist("") == set()
    info = sharding.bind_to_gpu_numa(0)  # no GPU / no sysfs entry: reports, changes nothing
    assert info["allowed_cpus"] >= 1 and info["bound_cpus"] in (None, info["bound_cpus"])
