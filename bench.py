#!/usr/bin/env python
"""Benchmark of the per-field hot path: fields/sec on 5-channel 2160^2 uint16 fields.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic fields: K1 (fused z-max ->
illumination divide -> 2x2 sum bin) followed by K3 (per-object statistics over the Cellpose
label mask), through the C-ABI of libips.so.  The workload is BASELINE.json configs[1]
(one 384-well plate, 9 sites/well, 5ch 2160^2 uint16, 3 z-planes, ~2000 cells/site); with
the defaults (216 steps x 16 fields) the timed region is exactly one 3456-field plate.

Printed keys (one JSON line on rank 0): see the task contract.  ``value`` is device-resident
throughput (inputs already in HBM), ``e2e`` is the same metric through the host-buffer
pipeline (pinned host -> device copies and result read-back inside the timed region),
``roofline`` is for the dominant kernel, ``cpu_baseline`` is the oracle (a NumPy/SciPy
port of the reference's CPU path) on this box's host cores on a bounded sample.

``--impl reference`` times that CPU path alone (no GPU work at all).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_RESULT_LINE = []          # the one JSON line, emitted on the real stdout at the very end
METRIC = "fields/sec (5ch 2160^2 uint16)"
UNIT = "fields/s"
C_, Z_, H_, W_, NCELLS, BIN = 5, 3, 2160, 2160, 2000, 2
WORKLOAD = "configs[1]: 384-well plate x 9 sites, 5ch 2160x2160 uint16, 3 z-planes, ~2000 cells/site"
FALLBACK_HBM_GBS = 6650.0


def k1_bytes_per_field(C, Z, H, W, b):
    """raw read + illum read + max-proj write + binned fp32 write."""
    return C * Z * H * W * 2 + C * H * W * 4 + C * H * W * 2 + C * (H // b) * (W // b) * 4


def k3_bytes_per_field(C, H, W, n_cells):
    """labels + max-proj + illum reads + object rows written."""
    return H * W * 4 + C * H * W * 2 + C * H * W * 4 + n_cells * (6 + 2 + 5 * C) * 4


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------
# synthetic plate
# ---------------------------------------------------------------------------------------
def base_label_masks(n_base, h, w, n_cells, seed):
    from image_processing_suite_b200 import synth
    return [synth.make_labels(h, w, n_cells, seed=seed + 17 * k) for k in range(n_base)]


def dihedral(lab, k):
    """k in 0..7: the 8 symmetries of the square (keeps labels non-overlapping, 1..N)."""
    a = np.rot90(lab, k % 4)
    if k >= 4:
        a = a[:, ::-1]
    return np.ascontiguousarray(a)


def ring_masks(n, h, w, n_cells, seed=0):
    n_base = (n + 7) // 8
    bases = base_label_masks(n_base, h, w, n_cells, seed)
    return [dihedral(bases[i // 8], i % 8) for i in range(n)]


# ---------------------------------------------------------------------------------------
# clock sampling during the timed region
# ---------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
        0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting",
    }

    def __init__(self, index, period=0.02):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self.period = period
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons - {"gpu_idle"}), "samples": len(s)}


# ---------------------------------------------------------------------------------------
# CPU path (oracle) -- the reference arm and the cpu_baseline leg
# ---------------------------------------------------------------------------------------
def cpu_field(raw, labels, illum, b, scale):
    """The reference's CPU arithmetic for one field: np.maximum.reduce per channel
    (MaxProjection.py:45), astype(float)/illum (Illumination_QC_mult.py:145-150), b x b sum
    binning, scipy.ndimage labelled statistics (CellProfiler MeasureObject* restatement)."""
    from oracle import object_stats as o_obj
    from oracle import preprocess as o_pre
    mp = np.stack([o_pre.max_projection(list(raw[c])) for c in range(raw.shape[0])])
    corr = np.stack([o_pre.illum_correct(mp[c], illum[c]) for c in range(raw.shape[0])])
    binned = o_pre.sum_bin(corr, b)
    ints, flts = o_obj.object_stats(labels, mp, illum, scale)
    return mp, binned, ints, flts


def cpu_band(raw, labels, illum, rows):
    return raw[:, :, :rows], labels[:rows], illum[:, :rows]


def run_cpu_path(fields, illum, n_threads, b=BIN, scale=1.0 / 65535.0):
    """Process a list of (raw, labels) with a thread pool, as Illumination_QC_mult.py:212 does."""
    import concurrent.futures as cf
    t0 = time.perf_counter()
    with cf.ThreadPoolExecutor(max_workers=n_threads) as ex:
        futs = [ex.submit(cpu_field, r, l, illum, b, scale) for r, l in fields]
        for fu in futs:
            fu.result()
    return time.perf_counter() - t0


def host_fields(n, h, w, seed=0):
    from image_processing_suite_b200 import synth
    masks = ring_masks(n, h, w, NCELLS if h == H_ else max(1, NCELLS * h * w // (H_ * W_)), seed)
    return [(synth.field_numpy(m, c=C_, z=Z_, seed=seed + i), m) for i, m in enumerate(masks)]


def reference_arm(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference is Python scripts around NumPy/SciPy calls and CellProfiler, nothing to
    compile), all host threads, bounded sample per step."""
    from image_processing_suite_b200 import synth
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    illum = synth.make_illum(C_, H_, W_, seed=0).astype(np.float64)
    # a few distinct fields, cycled; generation is outside the timed region
    n_distinct = max(1, min(cores, 8, args.ref_fields))
    fields = host_fields(n_distinct, H_, W_, seed=0)
    # size one step: time one full field on one thread, then pick rows so that
    # (warmup + steps) steps, each `cores` band-fields wide, fit the budget
    t1 = run_cpu_path(fields[:1], illum, 1)
    budget_s = float(args.ref_budget)
    per_step_budget = budget_s / max(args.steps + args.warmup, 1)
    # with `cores` threads a step of `cores` fields takes ~t1 * contention; assume 1.5x
    frac = min(1.0, per_step_budget / (t1 * 1.5))
    rows = max(8, int(H_ * frac) // 4 * 4)
    per_step = cores
    step_fields = [cpu_band(*fields[i % n_distinct], illum, rows) for i in range(per_step)]
    step_in = [(r, l) for r, l, _ in step_fields]
    ill_band = illum[:, :rows]
    for _ in range(args.warmup):
        run_cpu_path(step_in, ill_band, cores)
    t = 0.0
    for _ in range(args.steps):
        t += run_cpu_path(step_in, ill_band, cores)
    field_equiv = per_step * rows / float(H_)
    value = field_equiv * args.steps / t
    sample = "%d threads x %d-row band of a 2160-row field per step (%.3f field-equivalents/step), %d distinct fields" % (
        cores, rows, field_equiv, n_distinct)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "bin": BIN, "cpu_path": "numpy/scipy.ndimage oracle port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _RESULT_LINE.append(json.dumps(line))
    return 0


def cpu_baseline_leg(host_ring, illum_host, budget_s=20.0):
    """Bounded sample of the same workload on this box's cores (rank 0, N=1 only)."""
    cores = os.cpu_count() or 1
    illum64 = illum_host.astype(np.float64)
    n = min(len(host_ring), cores, 16)
    t1 = run_cpu_path(host_ring[:1], illum64, 1)
    if t1 * 1.5 > budget_s:                      # very slow host: shrink to a band
        rows = max(8, int(H_ * budget_s / (t1 * 1.5)) // 4 * 4)
    else:
        rows = H_
    sample = [(r[:, :, :rows], l[:rows]) for r, l in host_ring[:n]]
    t = run_cpu_path(sample, illum64[:, :rows], cores)
    fe = n * rows / float(H_)
    return {"value": fe / t, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d fields x %d rows of %d on %d threads (%.1f s); 1 field on 1 thread: %.2f s" % (
                n, rows, H_, cores, t, t1)}



def secondary_kernels(field_u16):
    """Short device-timed figures for the other kernels of the path (N=1, rank 0, after the timed
    regions): the Pillow-exact LANCZOS re-binning of one 5-plane field and the TIFF-LZW strip codec
    on its result.  Never fatal: a failure is reported as a string."""
    import torch
    from image_processing_suite_b200 import ops
    try:
        def timed(fn, iters):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / iters
        out = {}
        ms = timed(lambda: ops.lanczos_resize_u16(field_u16, (H_ // 2, W_ // 2)), 20)
        out["lanczos_2160_to_1080_ms_per_field"] = ms
        small = ops.lanczos_resize_u16(field_u16, (H_ // 2, W_ // 2)).repeat(16, 1, 1).contiguous()     # 80 planes = 16 re-binned fields
        px = small.numel() * 2
        ms = timed(lambda: ops.tiff_lzw_encode(small), 3)
        out["tiff_lzw_encode_pixel_gbs"] = px / ms / 1e6
        files, nbytes = ops.tiff_lzw_encode(small)
        from image_processing_suite_b200.scripts import tiffio
        blobs = [bytes(files[p, :int(n)].cpu().numpy()) for p, n in enumerate(nbytes)]
        t0 = time.perf_counter()
        back = tiffio.decode_to_device(blobs)
        torch.cuda.synchronize()
        out["tiff_lzw_decode_from_host_bytes_pixel_gbs"] = px / (time.perf_counter() - t0) / 1e9
        out["tiff_round_trip_exact"] = bool(torch.equal(back, small))
        out["planes"] = int(small.shape[0])
        return out
    except Exception as e:                                   # the headline numbers must not depend on this block
        return {"error": "%s: %s" % (type(e).__name__, e)}


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    from image_processing_suite_b200 import capi, ops, synth
    from image_processing_suite_b200.pipeline import FieldPipeline, pinned_empty

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    Fb, R = args.batch, args.ring
    assert R % Fb == 0, "--ring must be a multiple of --batch"
    # ---- synthetic plate shard: a ring of R distinct device-resident fields ---------------
    masks = ring_masks(R, H_, W_, NCELLS, seed=1000 * rank)
    # label masks are uint16, the dtype Cellpose writes them in (Cellpose_GPU_s3fs.py:143)
    labels = torch.empty((R, H_, W_), dtype=torch.uint16, device=dev)
    raw = torch.empty((R, C_, Z_, H_, W_), dtype=torch.uint16, device=dev)
    for i, m in enumerate(masks):
        lab32 = torch.from_numpy(m).to(dev)
        labels[i].copy_(lab32.to(torch.uint16))
        raw[i].copy_(synth.field_torch(lab32, c=C_, z=Z_, seed=1000 * rank + i))
    del lab32
    illum_host = synth.make_illum(C_, H_, W_, seed=0)
    illum = torch.from_numpy(illum_host).to(dev)
    # once per plate: the function's reciprocal (the per-pixel divide becomes a multiply)
    illum_rcp = ops.illum_reciprocal(illum) if args.mode == "fused" else None
    n_max = NCELLS
    scale = 1.0 / 65535.0

    nb = R // Fb
    fused = args.mode == "fused"
    labels32 = None if fused else labels.to(torch.int32)      # the split K3 kernel takes int32 masks
    # object rows of the whole plate shard stay resident (they are what the all-gather moves):
    # step i writes the rows of its 16 fields into its own slice
    n_plate = min(args.steps, max(1, 4096 // Fb)) * Fb
    plate_n = torch.zeros((n_plate,), dtype=torch.int32, device=dev)
    plate_ints = torch.zeros((n_plate, n_max, 6), dtype=torch.int32, device=dev)
    plate_flts = torch.zeros((n_plate, n_max, 2 + 5 * C_), dtype=torch.float32, device=dev)
    n_slots = n_plate // Fb
    # field -> well: 9 sites per well, wells dealt round-robin over ranks (plate.shard_wells)
    sites = 9
    n_wells_local = (n_plate + sites - 1) // sites
    field_well = (torch.arange(n_plate, device=dev, dtype=torch.int32) // sites) * world + rank
    n_wells = n_wells_local * world
    ws_f = None
    k1_out = [None] * nb
    for b in range(nb):                                   # preallocate the image outputs once
        sl = slice(b * Fb, (b + 1) * Fb)
        rows0 = {"n_objects": plate_n[:Fb], "ints": plate_ints[:Fb], "flts": plate_flts[:Fb]}
        if fused:
            k1_out[b] = ops.field_fused(raw[sl], illum, labels[sl], bin=BIN, intensity_scale=scale, n_max=n_max, out=rows0,
                                        illum_rcp=illum_rcp)
            ws_f = k1_out[b]["ws"]
        else:
            k1_out[b] = ops.preprocess_fused(raw[sl], illum, bin=BIN)
            ws_f = ops.object_stats(labels[sl].to(torch.int32), k1_out[b]["maxproj"], illum, scale, n_max=n_max, out=rows0)["ws"]
    torch.cuda.synchronize()
    from image_processing_suite_b200 import plate as plate_mod
    D_row = 10 + 5 * C_
    # The all-gather is issued in N_CHUNKS pieces on a side stream as the plate shard progresses,
    # so NVLink traffic overlaps the remaining fields; the per-well means come once at the end.
    # With one rank there is no gather to overlap: the rows are packed once, after the last field
    # (packing them mid-run on the side stream left every later field launch 5 % slower, measured).
    n_chunks = max(1, min(args.gather_chunks if world > 1 else 1, n_slots))
    while n_slots % n_chunks:
        n_chunks -= 1
    slots_per_chunk = n_slots // n_chunks
    chunk_fields = slots_per_chunk * Fb
    # One header-led block per (chunk, rank): row 0 carries the row count, so neither side of the
    # exchange needs a host round trip.  plate.PlateRowExchange packs a chunk's rows on a side stream
    # and moves the block over NVSwitch peer memory with the copy engines (every block is stored into
    # every peer's table, no SM involved); the one NCCL all-gather that remains (the per-chunk counts,
    # at the end of the plate) is the barrier.  Without CUDA IPC the blocks themselves go through
    # ncclAllGather, chunk by chunk.
    exchange = plate_mod.PlateRowExchange(n_chunks, chunk_fields, n_max, C_, peer_push=not args.no_peer_push)
    table, block_rows, push_note = exchange.table, exchange.block_rows, exchange.transport
    well_agg = plate_mod.WellAggregator(n_wells, D_row, device=dev)

    def gather_chunk(g):
        """Fields of chunk g are done on the compute stream: pack their rows into this rank's block and
        send it on its way.  No host synchronisation."""
        fs = slice(g * chunk_fields, (g + 1) * chunk_fields)
        exchange.submit(g, plate_ints[fs], plate_flts[fs], plate_n[fs], field_well[fs], field_base=g * chunk_fields)

    def finish_plate(ev=None):
        # Per-well sums run after the last field, over the whole gathered table (every rank sums the
        # rows of all wells): the blocks of the peers are only known to have landed after the barrier.
        blocks = exchange.finish()
        if ev is not None:
            ev.record()
        well_agg.reset()
        well_agg.add_blocks(blocks)
        return well_agg.finalize()

    def aggregate_all():
        for g in range(n_chunks):
            gather_chunk(g)
        return finish_plate()

    def step(i, ev=None):
        b = i % nb
        sl = slice(b * Fb, (b + 1) * Fb)
        ps = slice((i % n_slots) * Fb, (i % n_slots + 1) * Fb)
        rows_out = {"n_objects": plate_n[ps], "ints": plate_ints[ps], "flts": plate_flts[ps], "ws": ws_f}
        if ev is not None:
            ev[0].record()
        if fused:
            ops.field_fused(raw[sl], illum, labels[sl], bin=BIN, intensity_scale=scale, n_max=n_max,
                            out={"maxproj": k1_out[b]["maxproj"], "binned": k1_out[b]["binned"], **rows_out},
                            illum_rcp=illum_rcp)
            if ev is not None:
                ev[1].record()
                ev[2].record()
            return
        ops.preprocess_fused(raw[sl], illum, bin=BIN, out=k1_out[b])
        if ev is not None:
            ev[1].record()
        ops.object_stats(labels32[sl], k1_out[b]["maxproj"], illum, scale, n_max=n_max, out=rows_out)
        if ev is not None:
            ev[2].record()

    for i in range(args.warmup):
        step(i)
    aggregate_all()                                         # warm the gather / aggregation path too
    torch.cuda.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    t_begin, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_steps, t_gathered = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # The ranks enter the timed region together ON THE DEVICE: NVML start-up and the barrier's exit differ
    # by milliseconds between eight processes, and a rank that starts early would carry that skew as a
    # wait in front of the plate's final barrier.  NVML is therefore initialised before the barrier, and
    # a stream-ordered all-reduce right in front of t_begin holds every GPU until the last rank has
    # queued its own.
    sampler = ClockSampler(local)                               # NVML start-up: before the barrier
    sync_token = torch.zeros(1, device=dev)
    barrier()
    l0 = capi.launch_count()
    with sampler as clocks:
        pushed0 = exchange.pushed_bytes
        if dist is not None:
            dist.all_reduce(sync_token)
        t_begin.record()
        chunks_done = 0
        for i in range(args.steps):
            step(i, evs[i])
            if args.steps == n_slots and (i + 1) % slots_per_chunk == 0 and chunks_done < n_chunks:
                gather_chunk(chunks_done)                  # rows of these slots are final for the plate
                chunks_done += 1
        t_steps.record()
        for g in range(chunks_done, n_chunks):
            gather_chunk(g)
        well_mean_dev, well_count_dev = finish_plate(t_gathered)
        t_end.record()
        barrier()
    launches = capi.launch_count() - l0
    pushed_last_plate = exchange.pushed_bytes - pushed0
    # ---- content check of the gathered table (outside the timed region) ---------------------
    counts_all = exchange.counts()
    n_rows = int(counts_all[:, rank].sum().item())
    agg_check = "ok"
    local_rows = int(plate_n.clamp(min=0).sum().item())
    if n_rows != local_rows:
        agg_check = "own block counts %d != local object rows %d" % (n_rows, local_rows)
    rows_everywhere = torch.tensor([n_rows], device=dev, dtype=torch.int64)
    if dist is not None:
        dist.all_reduce(rows_everywhere)
    if int(counts_all.sum().item()) != int(rows_everywhere.item()):
        agg_check = "gathered counts %d != sum of the ranks' rows %d" % (int(counts_all.sum().item()), int(rows_everywhere.item()))
    # the per-well means of this rank's own wells, recomputed from its local blocks alone, must equal
    # the ones taken from the gathered table bit for bit (the float32 sums are exact, wellmean.cu)
    own_table = table[:, rank].contiguous()
    well_agg.reset()
    well_agg.add_blocks(own_table)
    own_mean, own_count = well_agg.finalize()
    own_wells = torch.unique(field_well.to(torch.int64))
    if not torch.equal(well_count_dev[own_wells], own_count[own_wells]):
        agg_check = "per-well row counts differ between the gathered table and the local rows"
    elif not torch.equal(well_mean_dev[own_wells].view(torch.int64), own_mean[own_wells].view(torch.int64)):
        agg_check = "per-well means differ between the gathered table and the local rows"
    if world > 1 and agg_check == "ok":
        # every rank must hold the same gathered means
        digest = well_mean_dev.nan_to_num(0.0).sum().reshape(1).clone()
        lo, hi = digest.clone(), digest.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        if float(lo.item()) != float(hi.item()):
            agg_check = "ranks disagree on the gathered per-well means"
    ms_total = t_begin.elapsed_time(t_end)
    ms_aggregate = t_steps.elapsed_time(t_end)
    ms_exchange_tail = t_steps.elapsed_time(t_gathered)
    k1_all = [e[0].elapsed_time(e[1]) for e in evs]
    k3_all = [e[1].elapsed_time(e[2]) for e in evs]
    k1_ms, k3_ms = float(np.mean(k1_all)), float(np.mean(k3_all))
    if os.environ.get("IPS_BENCH_DUMP"):
        print("per-launch ms:", " ".join("%.3f" % v for v in k1_all), file=sys.stderr)
    if dist is not None:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    fields_total = args.steps * Fb * world
    value = fields_total / (ms_total * 1e-3)

    # ---- sustained figure: one whole 3456-field plate (216 launches) back to back --------------
    # (outside the timed region; the board settles under its power limit some 40 ms into a run,
    # so a 20-step region is a burst figure -- profiles/README.md)
    sustained_ms = None
    if True:
        n_sus = args.sustained_steps
        if n_sus > 0:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            sus = []
            torch.cuda.synchronize()
            for i in range(n_sus):
                ev3 = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                step(i, ev3)
                sus.append(ev3)
            torch.cuda.synchronize()
            per = [e[0].elapsed_time(e[1]) if fused else e[0].elapsed_time(e[2]) for e in sus]
            sustained_ms = {"mean": float(np.mean(per)), "last_quarter": float(np.mean(per[-max(1, n_sus // 4):])),
                            "steps": n_sus}

    # ---- end to end through the host-buffer pipeline -------------------------------------
    e2e = None
    Fe = args.e2e_batch
    n_host = args.e2e_ring
    # label masks travel as uint16, the dtype Cellpose writes them in (Cellpose_GPU_s3fs.py:143)
    pipe = FieldPipeline(Fe, C_, Z_, H_, W_, bin=BIN, n_max=n_max, depth=3, illum=illum_host,
                         intensity_scale=scale, label_dtype=np.uint16)
    h_raw = [pinned_empty((Fe, C_, Z_, H_, W_), np.uint16) for _ in range(n_host)]
    h_lab = [pinned_empty((Fe, H_, W_), np.uint16) for _ in range(n_host)]
    for j in range(n_host):
        for k in range(Fe):
            src = (j * Fe + k) % R
            h_raw[j][k] = raw[src].cpu().numpy()
            h_lab[j][k] = labels[src].cpu().numpy()
    h_out = [pipe.output_buffers() for _ in range(3)]
    h_rows = [{k: v for k, v in o.items() if k in ("n_objects", "ints", "flts")} for o in h_out]
    e2e_steps = max(1, args.e2e_fields // Fe)

    def run_pipeline(outs):
        """e2e_steps batches through the host-buffer pipeline, depth 3; seconds (max over ranks)."""
        for i in range(3):
            pipe.wait(pipe.submit(h_raw[i % n_host], h_lab[i % n_host], outs[i % 3]))
        barrier()
        t0 = time.perf_counter()
        tickets = []
        for i in range(e2e_steps):
            if i >= 3:
                pipe.wait(tickets[i - 3])                     # its host output buffers are about to be reused
            tickets.append(pipe.submit(h_raw[i % n_host], h_lab[i % n_host], outs[i % 3]))
        for tk in tickets[-3:]:
            pipe.wait(tk)
        torch.cuda.synchronize()
        sec = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([sec], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    l1 = capi.launch_count()
    e2e_s = run_pipeline(h_out)
    e2e_launches = capi.launch_count() - l1
    rows_s = run_pipeline(h_rows)
    h2d_b, d2h_b = pipe.h2d_bytes(), pipe.d2h_bytes(h_out[0])
    d2h_rows_b = pipe.d2h_bytes(h_rows[0])
    # the box's bare copy ceiling for exactly these transfer sizes, all ranks at once (tools/bench_pcie.py)
    ceiling = None
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("bench_pcie", os.path.join(ROOT, "tools", "bench_pcie.py"))
        bp = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(bp)
        secs = bp.copy_ceiling(local, h2d_b, d2h_b, 10, barrier)
        if dist is not None:
            for k in secs:
                t = torch.tensor([secs[k]], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                secs[k] = float(t.item())
        ceiling = {"h2d_alone_gbs": world * h2d_b * 10 / secs["h2d"] / 1e9,
                   "d2h_alone_gbs": world * d2h_b * 10 / secs["d2h"] / 1e9,
                   "both_h2d_gbs": world * h2d_b * 10 / secs["both"] / 1e9,
                   "both_steps_per_s": world * 10 / secs["both"]}
    except Exception as e:                                    # the headline must not depend on this block
        ceiling = {"error": "%s: %s" % (type(e).__name__, e)}
    e2e_value = e2e_steps * Fe * world / e2e_s
    h2d_gbs = e2e_value / Fe * h2d_b / 1e9
    rows_value = e2e_steps * Fe * world / rows_s
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": h2d_b, "d2h_bytes_per_step": d2h_b,
           "fields_per_step": Fe, "steps": e2e_steps, "timer": "host wall clock around submit..wait (copies + kernels)",
           "gpu_launches": e2e_launches,
           "h2d_gbs": h2d_gbs,
           "pcie_ceiling_gbs": ceiling.get("both_h2d_gbs"),
           "frac_of_ceiling": (h2d_gbs / ceiling["both_h2d_gbs"]) if ceiling.get("both_h2d_gbs") else None,
           "ceiling": ceiling,
           "ceiling_note": "bare pinned cudaMemcpyAsync of the same H2D + D2H sizes on all ranks at once, no kernels "
                           "(tools/bench_pcie.py); aggregate GB/s host -> device",
           "rows_only": {"value": rows_value, "unit": UNIT, "d2h_bytes_per_step": d2h_rows_b,
                         "h2d_gbs": rows_value / Fe * h2d_b / 1e9,
                         "frac_of_h2d_alone_ceiling": (rows_value / Fe * h2d_b / 1e9 / ceiling["h2d_alone_gbs"])
                         if ceiling.get("h2d_alone_gbs") else None,
                         "what": "same pipeline, device -> host of the object rows only (what Feature_extraction consumes)"}}
    n_obj_last = h_out[(e2e_steps - 1) % 3]["n_objects"].copy()
    pipe.close()

    secondary = None
    if rank == 0 and world == 1:
        secondary = secondary_kernels(k1_out[0]["maxproj"][0].contiguous())

    # ---- the drop-in scripts on files (rank 0, N=1 only): TIFF files in -> TIFF / CSV files out -------
    e2e_files = None
    if rank == 0 and world == 1 and not args.no_files:
        try:
            import importlib.util
            spec = importlib.util.spec_from_file_location("bench_files", os.path.join(ROOT, "tools", "bench_files.py"))
            bf = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(bf)
            e2e_files = bf.main(["--sites", str(args.files_sites), "--distinct", "4", "--cpu-sites", "1"])
        except Exception as e:                                   # the headline numbers must not depend on this block
            e2e_files = {"error": "%s: %s" % (type(e).__name__, str(e)[:200])}

    # ---- CPU baseline on a bounded sample (rank 0, N=1 only) -------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        ring = [(h_raw[j][k], h_lab[j][k]) for j in range(n_host) for k in range(Fe)]
        cpu = cpu_baseline_leg(ring, illum_host, budget_s=args.cpu_budget)

    if rank == 0:
        peak, peak_src = hbm_peak()
        # Algorithmic bytes per LAUNCH (DESIGN.md "Roofline accounting"): every per-field
        # stream once per field; the plate-constant illumination function once per launch
        # (it is shared by the launch's fields and served from L2 after the first touch).
        ill_b = C_ * H_ * W_ * 4
        per_field = {
            "K1": k1_bytes_per_field(C_, Z_, H_, W_, BIN) - ill_b,
            "K3": k3_bytes_per_field(C_, H_, W_, NCELLS) - ill_b,
        }
        per_field["fused"] = per_field["K1"] + H_ * W_ * 2 + NCELLS * (6 + 2 + 5 * C_) * 4      # uint16 label masks
        kern = {}
        if fused:
            kern["fused"] = k1_ms
        else:
            kern["K1"], kern["K3"] = k1_ms, k3_ms
        kernels = {}
        for name, ms in kern.items():
            nbytes = per_field[name] * Fb + ill_b
            gbs = nbytes / (ms * 1e-3) / 1e9
            per_launch = k1_all if name in ("fused", "K1") else k3_all
            kernels[name] = {"ms_per_launch": ms, "ms_median": float(np.median(per_launch)), "ms_best": float(np.min(per_launch)),
                             "frac_burst": gbs / peak,
                             "ms_sustained": sustained_ms["mean"] if (sustained_ms and name in ("fused", "K1")) else None,
                             "frac_sustained": (nbytes / (sustained_ms["mean"] * 1e-3) / 1e9 / peak)
                             if (sustained_ms and name in ("fused", "K1")) else None,
                             "frac_sustained_last_quarter": (nbytes / (sustained_ms["last_quarter"] * 1e-3) / 1e9 / peak)
                             if (sustained_ms and name in ("fused", "K1")) else None,
                             "sustained_steps": sustained_ms["steps"] if sustained_ms else None,
                             "gbs": gbs, "frac": gbs / peak,
                             "bytes_per_field": per_field[name], "illum_bytes_per_launch": ill_b,
                             "survey_8d_gbs_illum_per_field": (per_field[name] + ill_b) * Fb / (ms * 1e-3) / 1e9}
        dom = max(kern, key=kern.get)
        names = {"K1": "preprocess_vec_kernel", "K3": "object_stats_scan_kernel (+init +compact)",
                 "fused": "field_fused_kernel (+init +compact)"}
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                with open(tp) as f:
                    traffic = json.load(f).get(dom)
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u16 in / f32 + f64 accumulate",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "fields_per_step": Fb, "bin": BIN, "n_cells": NCELLS,
                       "ring_fields": R, "l2": "inputs larger than L2: ring of %d distinct fields = %.1f GB" % (
                           R, R * (C_ * Z_ * H_ * W_ * 2 + H_ * W_ * 2) / 1e9),
                       "mode": args.mode,
                       "sharding": "fields by well across ranks, no data-path collective"},
            "e2e": e2e,
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": names[dom], "achieved": kernels[dom]["gbs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["frac"], "traffic": traffic,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": per_field[dom] * Fb + ill_b},
            "kernels": kernels,
            "aggregation": {"what": "pack rows -> %s -> per-well mean, inside the timed region" % (
                                ("all-gather of the plate's object rows in %d chunks on a side stream, overlapping the "
                                 "remaining fields (counts in block headers, no host sync)" % n_chunks) if world > 1
                                else "no gather at N=1"),
                            "check": agg_check, "bulk_transport": push_note,
                            "ms_after_last_step": ms_aggregate,
                            "ms_last_chunk_and_barrier": ms_exchange_tail, "ms_well_sums": ms_aggregate - ms_exchange_tail,
                            "rows_per_rank": n_rows, "row_bytes": D_row * 4,
                            "gather_bytes_per_rank": n_rows * D_row * 4 if world > 1 else 0,
                            "bytes_pushed_to_peers_last_plate": pushed_last_plate,
                            "wells": n_wells, "wells_with_rows": int((well_count_dev > 0).sum().item())},
            "cpu_baseline": cpu,
            "e2e_files": e2e_files,
            "secondary": secondary,
            "clocks": clocks.summary(),
            "objects_last_field": int(n_obj_last[-1]),
        }
        _RESULT_LINE.append(json.dumps(line))
    exchange.close()
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=216)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="fused", choices=["fused", "split"],
                    help="fused: one K1+K3 pass per step (ips_field_fused); split: K1 then K3")
    ap.add_argument("--gather-chunks", type=int, default=20, help="pieces the plate's row all-gather is issued in")
    ap.add_argument("--batch", type=int, default=16, help="fields per step")
    ap.add_argument("--ring", type=int, default=32, help="distinct device-resident fields")
    ap.add_argument("--e2e-batch", type=int, default=4)
    ap.add_argument("--e2e-ring", type=int, default=4, help="distinct pinned host batches")
    ap.add_argument("--e2e-fields", type=int, default=256)
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--ref-budget", type=float, default=120.0)
    ap.add_argument("--ref-fields", type=int, default=8, help="distinct synthetic fields the reference arm cycles through")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files", action="store_true", help="skip the file-level run of the drop-in scripts (e2e_files)")
    ap.add_argument("--files-sites", type=int, default=96)
    ap.add_argument("--no-peer-push", action="store_true", help="gather the row blocks with ncclAllGather instead of peer stores")
    ap.add_argument("--sustained-steps", type=int, default=216,
                    help="extra launches after the timed region for the sustained kernel figure (0 = skip)")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: native libraries (NCCL prints its version banner to
    # stdout when NCCL_DEBUG is set) are pointed at stderr for the duration of the run
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        rc = reference_arm(args) if args.impl == "reference" else gpu_arm(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    if _RESULT_LINE:
        os.write(real_stdout, (_RESULT_LINE[0] + "\n").encode())
    os.close(real_stdout)
    return rc


if __name__ == "__main__":
    sys.exit(main())
