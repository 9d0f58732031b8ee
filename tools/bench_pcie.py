#!/usr/bin/env python
"""Bare host <-> device copy ceiling of the box, measured the way the pipeline uses the link.

    python tools/bench_pcie.py                      # 1 GPU
    python tools/bench_pcie.py --sweep 1 2 4 8      # one torchrun launch per N, one JSON line each
    torchrun --nproc-per-node N tools/bench_pcie.py # what --sweep runs

Every rank owns one GPU, allocates page-locked buffers of the pipeline's batch sizes
(ips_host_alloc, exactly the allocator bench.py's e2e leg uses) and issues cudaMemcpyAsync
copies on two streams: host -> device of `--h2d-mb` per step and device -> host of `--d2h-mb`
per step, three modes (h2d alone, d2h alone, both concurrently).  Time = max over ranks between
two barriers; aggregate GB/s = bytes of all ranks / that time.  No kernels run.

The e2e leg of bench.py reports `pcie_ceiling_gbs` (this number for its own N, measured in the same
process just before the timed region) and `frac_of_ceiling` = pipeline H2D GB/s / ceiling.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pin_rank_to_cores(local, world_local):
    """Give every rank its own slice of the host cores before it allocates page-locked memory
    (first touch decides the NUMA node of the pages)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(world_local, 1))
        mine = cores[local * per:(local + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        return mine
    except Exception:
        return None


def copy_ceiling(dev, h2d_bytes, d2h_bytes, iters, barrier, flags=0, modes=("h2d", "d2h", "both")):
    """Seconds per mode for `iters` copies of each size on this rank (wall clock, after a barrier)."""
    import ctypes as C
    import torch
    from image_processing_suite_b200 import capi
    from image_processing_suite_b200.pipeline import pinned_empty
    import numpy as np
    torch.cuda.set_device(dev)
    h_in = pinned_empty((h2d_bytes,), np.uint8, flags=flags)
    h_out = pinned_empty((d2h_bytes,), np.uint8)
    h_in[::4096] = 1                                           # touch every page from this rank's cores
    h_out[::4096] = 1
    d_in = torch.empty((h2d_bytes,), dtype=torch.uint8, device=dev)
    d_out = torch.zeros((d2h_bytes,), dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    t_in = torch.from_numpy(h_in)
    t_out = torch.from_numpy(h_out)
    out = {}
    for mode in modes:
        for timed in (False, True):
            n = iters if timed else 2
            barrier()
            t0 = time.perf_counter()
            for _ in range(n):
                if mode in ("h2d", "both"):
                    with torch.cuda.stream(s_in):
                        d_in.copy_(t_in, non_blocking=True)
                if mode in ("d2h", "both"):
                    with torch.cuda.stream(s_out):
                        t_out.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()
            barrier()
            if timed:
                out[mode] = time.perf_counter() - t0
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-mb", type=float, default=597.1968, help="bytes per step host -> device (4 fields of bench.py)")
    ap.add_argument("--d2h-mb", type=float, default=280.992, help="bytes per step device -> host")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--pin-cores", action="store_true", help="sched_setaffinity: each rank its own slice of the cores")
    ap.add_argument("--write-combined", action="store_true", help="cudaHostAllocWriteCombined for the input buffer")
    ap.add_argument("--sweep", type=int, nargs="*", default=None)
    args = ap.parse_args()
    if args.sweep:
        for n in args.sweep:
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                   "--master-addr", "127.0.0.1", "--master-port", str(29600 + n), os.path.abspath(__file__),
                   "--h2d-mb", str(args.h2d_mb), "--d2h-mb", str(args.d2h_mb), "--iters", str(args.iters)]
            if args.pin_cores:
                cmd.append("--pin-cores")
            if args.write_combined:
                cmd.append("--write-combined")
            r = subprocess.run(cmd, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                sys.stdout.write(json.dumps({"n_gpus": n, "error": r.stderr[-400:]}) + "\n")
            sys.stdout.flush()
        return 0
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cores = pin_rank_to_cores(local, world) if args.pin_cores else None
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")

    def barrier():
        if dist is not None:
            dist.barrier()

    h2d = int(args.h2d_mb * 1e6)
    d2h = int(args.d2h_mb * 1e6)
    secs = copy_ceiling(local, h2d, d2h, args.iters, barrier, flags=2 if args.write_combined else 0)
    if rank == 0:
        line = {"what": "bare pinned cudaMemcpyAsync ceiling", "n_gpus": world, "iters": args.iters,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "pin_cores": bool(args.pin_cores),
                "write_combined": bool(args.write_combined), "host_cores": os.cpu_count(),
                "rank0_cores": len(cores) if cores else None,
                "h2d_alone_gbs": world * h2d * args.iters / secs["h2d"] / 1e9,
                "d2h_alone_gbs": world * d2h * args.iters / secs["d2h"] / 1e9,
                "both_h2d_gbs": world * h2d * args.iters / secs["both"] / 1e9,
                "both_d2h_gbs": world * d2h * args.iters / secs["both"] / 1e9,
                "both_steps_per_s": world * args.iters / secs["both"]}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
