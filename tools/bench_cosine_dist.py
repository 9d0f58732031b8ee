"""Config 5 sharded over the box: python -m torch.distributed.run --nproc-per-node N tools/bench_cosine_dist.py [rows] [d].

Every rank generates its own block of N(0, 1) rows; timed (CUDA events, max over ranks): normalise +
bf16 split, all-gather of the planes, the rank's share of the tensor-core pass, exchange of the
partial sums.  Checked against the closed form (|sum x^|^2 - sum |x^|^2) / 2 evaluated in float64."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from image_processing_suite_b200.cosine_parallel import ShardedCosine

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n_local = (rows + world - 1) // world
g = torch.Generator(device="cuda").manual_seed(1234 + rank)
x = torch.randn((n_local, d), device="cuda", generator=g)
first = rank * n_local
if first + n_local > rows:                                   # zero rows pad the last block
    x[max(rows - first, 0):] = 0
sc = ShardedCosine(n_local, d)
s = sc.sum_triu(x)                                            # warm-up (communicator, tensor maps)
torch.cuda.synchronize()
if dist is not None:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
s = sc.sum_triu(x)
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
xh = x.double()
nrm = xh.norm(dim=1, keepdim=True)
xh = torch.where(nrm > 0, xh / nrm, torch.zeros_like(xh))
vec = xh.sum(0)
sq = (xh * xh).sum().reshape(1)
if dist is not None:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(vec)
    dist.all_reduce(sq)
ref = 0.5 * (float((vec * vec).sum()) - float(sq))
npairs = rows * (rows - 1) / 2
if rank == 0:
    t = float(ms) * 1e-3
    print(json.dumps({"what": "cosine strict-upper-triangle mean, one group sharded over ranks", "rows": rows, "d": d,
                      "n_gpus": world, "seconds": t, "algorithmic_tflops": npairs * 2 * d / t / 1e12,
                      "mean_cos": s / npairs, "mean_cos_ref": ref / npairs, "abs_err_mean": abs(s - ref) / npairs,
                      "plane_gather_gb_per_rank": 2 * n_local * sc.row_bytes * (world - 1) / 1e9 if world > 1 else 0.0}))
sc.close()
if dist is not None:
    dist.destroy_process_group()
