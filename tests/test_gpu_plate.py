"""Dense row packing and well aggregation on the device vs NumPy / pandas (single GPU)."""
import numpy as np
import pytest

from image_processing_suite_b200 import synth
from oracle import normalize as o_norm
from oracle import object_stats as o_obj
from oracle import preprocess as o_pre
from tests.gpu_util import dev, host, require_gpu

pytestmark = pytest.mark.gpu


def test_pack_rows_and_well_means_match_groupby():
    torch = require_gpu()
    from image_processing_suite_b200 import ops, plate
    F, C, H, W, cells = 6, 2, 96, 128, 9
    labs = np.stack([synth.make_labels(H, W, cells if f != 2 else 0, seed=f, amin=5, amax=10) for f in range(F)])
    mps = np.stack([o_pre.max_projection_field(synth.field_numpy(labs[f], c=C, z=2, seed=f)) for f in range(F)])
    res = ops.object_stats(dev(labs), dev(mps), None, 1.0, n_max=cells + 2)
    field_well = np.array([4, 4, 7, 7, 9, 9], np.int32)
    rows, total = plate.pack_rows(res["ints"], res["flts"], res["n_objects"], dev(field_well), field_base=100)
    n = int(total.item())
    expect = []
    for f in range(F):
        e_i, e_f = o_obj.object_stats(labs[f], mps[f], None, 1.0)
        for r in range(e_i.shape[0]):
            expect.append(np.r_[field_well[f], 100 + f, e_i[r], e_f[r]])
    expect = np.asarray(expect)
    assert n == expect.shape[0] and rows.shape[1] == 10 + 5 * C
    got = host(rows)[:n].astype(np.float64)
    np.testing.assert_array_equal(got[:, :8], expect[:, :8])           # well, field, label, area, bbox exact
    np.testing.assert_allclose(got[:, 8:], expect[:, 8:], rtol=1e-5, atol=1e-6)
    # one-rank "gather" + per-well means == pandas groupby over the same rows
    g = plate.RowGatherer(rows.shape[0], rows.shape[1])
    all_rows, counts = g.gather(rows, n)
    mean, count = plate.well_means(all_rows, counts, 12)
    ids, ref = o_norm.well_mean(got, got[:, 0].astype(int))
    m, c = host(mean), host(count)
    np.testing.assert_allclose(m[ids], ref, rtol=1e-12)
    assert c.sum() == n and set(np.flatnonzero(c)) == set(ids.tolist())
    assert np.isnan(m[0]).all()


def test_well_aggregator_streaming_equals_one_shot():
    torch = require_gpu()
    from image_processing_suite_b200 import plate
    rng = np.random.default_rng(3)
    blocks, cap, D, n_wells = 5, 40, 12, 9
    rows = rng.normal(size=(blocks, cap, D)).astype(np.float32)
    rows[:, :, 0] = rng.integers(0, n_wells, (blocks, cap))
    counts = np.array([40, 0, 17, 33, 1], np.int64)
    d_rows, d_counts = dev(rows), dev(counts)
    one_mean, one_count = plate.well_means(d_rows, d_counts, n_wells)
    agg = plate.WellAggregator(n_wells, D)
    agg.add(d_rows[:2].contiguous(), d_counts[:2].contiguous())
    agg.add(d_rows[2:].contiguous(), d_counts[2:].contiguous())
    mean, count = agg.finalize()
    np.testing.assert_allclose(host(mean), host(one_mean), rtol=1e-12, equal_nan=True)
    np.testing.assert_array_equal(host(count), host(one_count))
    valid = np.concatenate([rows[b, :counts[b]] for b in range(blocks)]).astype(np.float64)
    ids, ref = o_norm.well_mean(valid, valid[:, 0].astype(int))
    np.testing.assert_allclose(host(mean)[ids], ref, rtol=1e-12)


def _nccl_worker(rank, world, port, tmp):
    import os
    import torch
    import torch.distributed as dist
    from image_processing_suite_b200 import plate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        F, n_max, C = 5, 20, 2
        nf = 2 + 5 * C
        D = 8 + nf

        def rank_inputs(r):
            rng = np.random.default_rng(500 + r)
            n_obj = rng.integers(0, n_max + 1, F).astype(np.int32)
            ints = rng.integers(0, 400, (F, n_max, 6)).astype(np.int32)
            flts = rng.normal(5.0, 2.0, (F, n_max, nf)).astype(np.float32)
            wells = (np.arange(F, dtype=np.int32) // 2) * world + r
            return n_obj, ints, flts, wells

        n_obj, ints, flts, wells = rank_inputs(rank)
        table = torch.zeros((world, F * n_max + 1, D), dtype=torch.float32, device="cuda")
        plate.pack_rows_block(dev(ints), dev(flts), dev(n_obj), dev(wells), table[rank], field_base=100 * rank)
        g = plate.BlockGatherer(backend="ips")
        g.gather(table)                                   # ONE ncclAllGather inside libips.so
        counts = plate.block_counts(table).cpu().numpy()
        agg = plate.WellAggregator(3 * world, D)
        agg.add_blocks(table)
        mean, count = agg.finalize()
        torch.cuda.synchronize()
        # every rank rebuilds every rank's rows on the host: contents, not just counts
        for r in range(world):
            e_n, e_i, e_f, e_w = rank_inputs(r)
            assert counts[r] == int(e_n.sum())
            rows = []
            for f in range(F):
                for k in range(e_n[f]):
                    rows.append(np.r_[e_w[f], 100 * r + f, e_i[f, k], e_f[f, k]].astype(np.float32))
            got = table[r, 1:1 + counts[r]].cpu().numpy()
            np.testing.assert_array_equal(got, np.asarray(rows, np.float32).reshape(-1, D))
        torch.save({"mean": mean.cpu(), "count": count.cpu(), "table": table.cpu()}, os.path.join(tmp, f"n{rank}.pt"))
        g.close()
    finally:
        dist.destroy_process_group()


def test_nccl_block_gather_content_two_gpus(tmp_path):
    """The product collective (ips_allgather_blocks -> ncclAllGather) on 2 GPUs: every rank ends up with
    every rank's rows bit for bit, and with identical per-well means (VERDICT r1 missing #3)."""
    torch = require_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "n0.pt")
    b = torch.load(tmp_path / "n1.pt")
    assert torch.equal(a["table"], b["table"])
    assert torch.equal(a["count"], b["count"]) and torch.equal(a["mean"].view(torch.int64), b["mean"].view(torch.int64))


def _push_worker(rank, world, port, tmp):
    import os
    import torch
    import torch.distributed as dist
    from image_processing_suite_b200 import plate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        chunks, rows, D = 3, 50, 9
        table = plate.exportable_zeros((chunks, world, rows, D))
        pusher = plate.PeerPusher(table)
        gatherer = plate.BlockGatherer(backend="ips")
        totals = torch.zeros((world, 1, D), dtype=torch.float32, device="cuda")
        for g in range(chunks):
            table[g, rank] = torch.arange(rows * D, device="cuda", dtype=torch.float32).reshape(rows, D) + 1000 * rank + 100000 * g
            pusher.push(table[g, rank])                    # copy-engine stores into the peer's table
        totals[rank, 0, 0] = float(rank + 1)
        gatherer.gather(totals)                            # the one NCCL all-gather = the barrier
        torch.cuda.synchronize()
        for g in range(chunks):
            for r in range(world):
                want = torch.arange(rows * D, dtype=torch.float32).reshape(rows, D) + 1000 * r + 100000 * g
                assert torch.equal(table[g, r].cpu(), want), (g, r)
        assert totals[:, 0, 0].cpu().tolist() == [float(r + 1) for r in range(world)]
        with pytest.raises(ValueError):
            plate.PeerPusher(torch.zeros((4, 4), device="cuda"))
        pusher.close()
        gatherer.close()
        open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_peer_push_gather_two_gpus(tmp_path):
    """The bulk transport bench.py uses at N > 1: every rank stores its blocks into the peers' tables
    over NVLink (CUDA IPC, copy engines), one NCCL all-gather of the totals is the barrier."""
    torch = require_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_push_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def _exchange_worker(rank, world, port, tmp):
    import os
    import torch
    import torch.distributed as dist
    from image_processing_suite_b200 import plate
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        n_chunks, cf, n_max, C = 4, 3, 25, 2
        nf = 2 + 5 * C
        D = 8 + nf
        results = {}
        for mode in ("wait", True, False):
            ex = plate.PlateRowExchange(n_chunks, cf, n_max, C, exact_push=mode)
            assert ex.pusher is not None, ex.transport
            for rep in range(2):                                   # the second plate overwrites the first
                rng = np.random.default_rng(900 + 10 * rank + rep)
                n_obj = rng.integers(0, n_max + 1, (n_chunks, cf)).astype(np.int32)
                ints = rng.integers(0, 400, (n_chunks, cf, n_max, 6)).astype(np.int32)
                flts = rng.normal(5.0, 2.0, (n_chunks, cf, n_max, nf)).astype(np.float32)
                wells = ((np.arange(n_chunks * cf, dtype=np.int32) // 2) * world + rank).reshape(n_chunks, cf)
                before = ex.pushed_bytes
                for g in range(n_chunks):
                    ex.submit(g, dev(ints[g]), dev(flts[g]), dev(n_obj[g]), dev(wells[g]), field_base=g * cf)
                blocks = ex.finish()
                agg = plate.WellAggregator(n_chunks * cf * world, D)
                agg.add_blocks(blocks)
                mean, count = agg.finalize()
                torch.cuda.synchronize()
                counts = ex.counts().cpu().numpy()
                assert counts[:, rank].tolist() == n_obj.sum(1).tolist()
                pushed = ex.pushed_bytes - before
                cap = n_chunks * ex.block_rows * D * 4 * (world - 1)
                if mode == "wait":                                  # all but the last chunk travel at their size
                    want = (sum(int(c) + 1 for c in n_obj.sum(1)[:-1]) + ex.block_rows) * D * 4 * (world - 1)
                    assert pushed == want < cap
                elif mode is False:
                    assert pushed == cap
                # the gathered chunk counts (the barrier's payload) are the headers of the table
                cc = ex.chunk_counts.cpu().view(torch.int32)[:, :, 0].numpy().T
                np.testing.assert_array_equal(cc, counts)
                valid = [blocks[b, 1:1 + int(c)].cpu() for b, c in enumerate(counts.reshape(-1))]
                results[(str(mode), rep)] = {"rows": valid, "mean": mean.cpu(), "count": count.cpu()}
                ex.release()                                    # the peers write this table: all ranks done before the next plate
            ex.close()
        torch.save(results, os.path.join(tmp, f"x{rank}.pt"))
    finally:
        dist.destroy_process_group()


def test_plate_row_exchange_two_gpus(tmp_path):
    """plate.PlateRowExchange (what bench.py drives at N > 1): ragged chunks packed, pushed at their
    real size one submit later (or at capacity), published by the one small NCCL all-gather; both ranks
    hold the same valid rows and bit-identical per-well means, whatever the push mode."""
    torch = require_gpu()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_exchange_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "x0.pt")
    b = torch.load(tmp_path / "x1.pt")
    assert a.keys() == b.keys() and len(a) == 6
    for k in a:
        assert len(a[k]["rows"]) == len(b[k]["rows"])
        for ra, rb in zip(a[k]["rows"], b[k]["rows"]):
            assert torch.equal(ra, rb)
        assert torch.equal(a[k]["count"], b[k]["count"])
        assert torch.equal(a[k]["mean"].view(torch.int64), b[k]["mean"].view(torch.int64))
    for rep in (0, 1):                                             # and the modes agree with each other
        for m in ("True", "False"):
            assert torch.equal(a[("wait", rep)]["mean"].view(torch.int64), a[(m, rep)]["mean"].view(torch.int64))


def test_plate_row_exchange_single_rank():
    """One rank: the exchange only packs; finish() hands the blocks to the aggregator."""
    torch = require_gpu()
    from image_processing_suite_b200 import plate
    rng = np.random.default_rng(3)
    n_chunks, cf, n_max, C = 2, 4, 10, 1
    nf = 2 + 5 * C
    ex = plate.PlateRowExchange(n_chunks, cf, n_max, C)
    n_obj = rng.integers(0, n_max + 1, (n_chunks, cf)).astype(np.int32)
    ints = rng.integers(0, 400, (n_chunks, cf, n_max, 6)).astype(np.int32)
    flts = rng.normal(5.0, 2.0, (n_chunks, cf, n_max, nf)).astype(np.float32)
    wells = (np.arange(n_chunks * cf, dtype=np.int32) // 3).reshape(n_chunks, cf)
    for g in range(n_chunks):
        ex.submit(g, dev(ints[g]), dev(flts[g]), dev(n_obj[g]), dev(wells[g]), field_base=g * cf)
    blocks = ex.finish()
    assert ex.counts().cpu().numpy()[:, 0].tolist() == n_obj.sum(1).tolist() and ex.pushed_bytes == 0
    agg = plate.WellAggregator(3, 8 + nf)
    agg.add_blocks(blocks)
    mean, count = agg.finalize()
    assert host(count).tolist() == [int(n_obj.reshape(-1)[wells.reshape(-1) == w].sum()) for w in range(3)]
    with pytest.raises(ValueError):
        ex.submit(5, dev(ints[0]), dev(flts[0]), dev(n_obj[0]), dev(wells[0]))
    ex.close()
