"""K2 sharded by (plate, channel) unit: python -m torch.distributed.run --nproc-per-node N tools/bench_illum_dist.py

BASELINE configs[3]-shaped: 4 plates x 4 timepoints = 16 plate-timepoints x 5 channels = 80 units over
the ranks (plate.shard_units), no collective.  One unit = 3456 max-projected 2160^2 fields of one
channel streamed through ips_illum_accumulate (a ring of device-resident fields larger than L2) +
ips_illum_finalize (sigma = 200 / 2.35).  Timed with CUDA events, max over ranks."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from image_processing_suite_b200 import ops, plate

n_fields = int(sys.argv[1]) if len(sys.argv) > 1 else 3456
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H = W = 2160
chans = ["c%d" % i for i in range(5)]
mine = plate.shard_units(16, chans, rank, world)
g = torch.Generator(device="cuda").manual_seed(rank)
ring = torch.randint(0, 4096, (32, 1, H, W), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)   # 299 MB > L2
def unit():
    est = ops.IllumEstimator(1, H, W)
    for b in range(n_fields // 16):
        est.add(ring[(b % 2) * 16:(b % 2) * 16 + 16])
    return est.finalize(200.0 / 2.35, 0.02)
unit(); torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in mine:
    out = unit()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms) * 1e-3
    nbytes = 80 * (n_fields // 16 * 16) * H * W * 2
    print(json.dumps({"what": "illumination estimation, 80 (plate, channel) units sharded over ranks, no collective",
                      "n_gpus": world, "units_per_rank": len(mine), "fields_per_unit": n_fields // 16 * 16, "seconds": t,
                      "units_per_s": 80 / t, "aggregate_gbs": nbytes / t / 1e9, "min_illum": float(out.min())}))
if dist is not None:
    dist.destroy_process_group()
