#!/usr/bin/env python
"""Tail of the plate step on ONE GPU: the well aggregation over a gathered table of the size an
N-rank run holds (every rank sums the rows of all wells), and the packing of one chunk.

    python tools/bench_wellagg.py [--world 8] [--steps 20] [--chunks 10]

Builds the [chunks * world][block_rows][35] table of header-led blocks bench.py would hold after
the all-gather (2000 objects per field, 16 fields per step, 9 sites per well), times
ips_well_sums_reset + ips_well_sums_add_blocks + ips_well_sums_finalize with CUDA events and checks
the means against a float64 torch reference.  One JSON line.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--chunks", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    from image_processing_suite_b200 import plate
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    Fb, n_max, C_, sites = 16, 2000, 5, 9
    D = 10 + 5 * C_
    n_chunks = a.chunks
    while a.steps % n_chunks:
        n_chunks -= 1
    chunk_fields = a.steps // n_chunks * Fb
    block_rows = chunk_fields * n_max + 1
    n_fields = a.steps * Fb
    n_wells = (n_fields + sites - 1) // sites * a.world
    g = torch.Generator(device=dev).manual_seed(0)
    table = torch.empty((n_chunks, a.world, block_rows, D), dtype=torch.float32, device=dev)
    scale = torch.tensor([1, 1, 2000, 900, 2160, 2160, 2160, 2160, 2160, 2160] + [300.0, 0.3, 0.05, 0.1, 0.9] * C_, device=dev)
    for c in range(n_chunks):
        for r in range(a.world):
            blk = table[c, r]
            blk[1:] = torch.rand((block_rows - 1, D), device=dev, generator=g) * scale
            f = c * chunk_fields + torch.arange(block_rows - 1, device=dev) // n_max
            blk[1:, 0] = ((f // sites) * a.world + r).to(torch.float32)
            blk[1:, 1] = f.to(torch.float32)
            blk[0].zero_()
            blk[0, :2].view(torch.int32)[0] = block_rows - 1
    flat = table.view(n_chunks * a.world, block_rows, D)
    agg = plate.WellAggregator(n_wells, D, device=dev)

    def run():
        agg.reset()
        agg.add_blocks(flat)
        return agg.finalize()

    mean, count = run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    # float64 reference on a sample of wells
    rows = flat[:, 1:].reshape(-1, D)
    ok = True
    for w in (0, 1, n_wells // 2, n_wells - 1):
        sel = rows[:, 0] == float(w)
        ref = rows[sel].to(torch.float64).mean(0)
        ok &= bool(torch.allclose(mean[w], ref, rtol=1e-12, atol=0.0)) and int(count[w]) == int(sel.sum())
    # packing of one chunk from padded per-field outputs
    ints = torch.randint(0, 2000, (chunk_fields, n_max, 6), dtype=torch.int32, device=dev)
    flts = torch.rand((chunk_fields, n_max, 2 + 5 * C_), device=dev)
    n_obj = torch.full((chunk_fields,), n_max, dtype=torch.int32, device=dev)
    fw = (torch.arange(chunk_fields, device=dev, dtype=torch.int32) // sites)
    blk = torch.empty((block_rows, D), dtype=torch.float32, device=dev)
    plate.pack_rows_block(ints, flts, n_obj, fw, blk)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        plate.pack_rows_block(ints, flts, n_obj, fw, blk)
    e1.record()
    torch.cuda.synchronize()
    ms_pack = e0.elapsed_time(e1) / a.iters
    nbytes = flat.numel() * 4
    print(json.dumps({"what": "well aggregation over the gathered table of a %d-rank, %d-step run on one GPU" % (a.world, a.steps),
                      "rows": int(rows.shape[0]), "table_bytes": nbytes, "wells": n_wells,
                      "ms_reset_add_finalize": ms, "gbs": nbytes / ms / 1e6, "means_match_float64_reference": ok,
                      "pack_chunk_rows": chunk_fields * n_max, "ms_pack_chunk": ms_pack,
                      "pack_gbs": (ints.numel() * 4 + flts.numel() * 4 + blk.numel() * 4) / ms_pack / 1e6}))


if __name__ == "__main__":
    main()
