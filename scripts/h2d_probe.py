"""Concurrent pinned host->device bandwidth per GPU: every rank copies its own pinned buffer to its
own GPU at the same time, plain cudaMemcpyAsync loops; rank 0 prints one JSON object.
    python scripts/h2d_probe.py                      (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/h2d_probe.py
The ceiling this measures is what the fp32-input e2e number of bench.py can reach: 12 288 B per image over this link."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from convnet_quantization_b200 import sharding

rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
bind = "--no-bind" not in sys.argv
numa = sharding.bind_to_gpu_numa(local) if bind else sharding.gpu_numa_info(local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
res = {}
for name, nbytes in (("fp32_batch_192MiB", 16384 * 12288), ("uint8_batch_48MiB", 16384 * 3072), ("chunk_24MiB", 2048 * 12288)):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.random_(0, 256)
    devbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        devbuf.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    iters = max(10, int(4e9 / nbytes))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        devbuf.copy_(host, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    gbs = nbytes * iters / (e0.elapsed_time(e1) * 1e-3) / 1e9
    res[name] = gbs
    if world > 1:
        dist.barrier()
mine = {"rank": rank, "numa": numa, "h2d_gbs": res}
if world > 1:
    allr = [None] * world
    dist.all_gather_object(allr, mine)
else:
    allr = [mine]
if rank == 0:
    out = {"what": "concurrent pinned H2D cudaMemcpyAsync, one process per GPU, all ranks at once", "n_gpus": world,
           "bound_to_gpu_numa": bind, "host_cpus_allowed": len(os.sched_getaffinity(0)), "ranks": allr,
           "aggregate_gbs": {k: sum(r["h2d_gbs"][k] for r in allr) for k in res},
           "min_per_gpu_gbs": {k: min(r["h2d_gbs"][k] for r in allr) for k in res},
           "fp32_e2e_ceiling_images_per_s": sum(r["h2d_gbs"]["fp32_batch_192MiB"] for r in allr) * 1e9 / 12288,
           "uint8_e2e_ceiling_images_per_s": sum(r["h2d_gbs"]["uint8_batch_48MiB"] for r in allr) * 1e9 / 3072}
    dst = [a.split("=", 1)[1] for a in sys.argv if a.startswith("--out=")]
    if dst:  # NCCL prints its banner on stdout: a file keeps the JSON clean
        with open(dst[0], "w") as f:
            json.dump(out, f, indent=1)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
