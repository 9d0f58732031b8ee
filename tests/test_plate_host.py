"""Host-side plate logic on the CPU: sharding, row schema, and the padded all-gather protocol
over gloo with world_size 2 (the N > 1 path of bench.py / plate.py without GPUs)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from image_processing_suite_b200 import plate


def test_sharding_covers_every_well_once():
    for world in (1, 2, 4, 8, 5):
        seen = []
        for r in range(world):
            wells = plate.shard_wells(384, r, world)
            assert len(wells) in (384 // world, 384 // world + 1)
            seen += wells
        assert sorted(seen) == list(range(384))
    fields = plate.plate_fields(384, 9, rank=3, world=8)
    assert len(fields) == 48 * 9 and fields[0] == (3, 1) and fields[9] == (11, 1)
    with pytest.raises(ValueError):
        plate.shard_wells(384, 8, 8)


def test_well_names_and_row_schema():
    assert plate.well_name(0) == "A01" and plate.well_name(23) == "A24" and plate.well_name(383) == "P24"
    cols = plate.row_columns(["DNA", "ER"])
    assert cols[:10] == list(plate.ROW_PREFIX) and len(cols) == 10 + 5 * 2 and cols[-1] == "max_ER"


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, cap = 7, 20
        rng = np.random.default_rng(100 + rank)
        wells = plate.shard_wells(6, rank, world)
        n_local = 5 + 4 * rank
        rows = torch.from_numpy(rng.normal(size=(n_local, D)).astype(np.float32))
        rows[:, 0] = torch.tensor([wells[i % len(wells)] for i in range(n_local)], dtype=torch.float32)
        g = plate.RowGatherer(cap, D, backend="torch")
        all_rows, counts = g.gather(rows, n_local)
        assert counts.tolist() == [5 + 4 * r for r in range(world)]
        assert torch.equal(all_rows[rank, :n_local], rows)
        ids = plate.well_ids_of(all_rows, counts)
        assert int((ids >= 0).sum()) == int(counts.sum())
        # every rank ends up with the same table -> same per-well means as a pandas groupby
        valid = ids >= 0
        flat = all_rows.reshape(-1, D)[valid].numpy().astype(np.float64)
        means = {int(w): flat[flat[:, 0] == w].mean(axis=0) for w in np.unique(flat[:, 0])}
        torch.save({"means": means, "counts": counts}, os.path.join(tmp, f"r{rank}.pt"))
        with pytest.raises(ValueError):
            g.gather(torch.zeros((cap + 1, D)), cap + 1)
    finally:
        dist.destroy_process_group()


def test_row_gather_protocol_gloo_world2(tmp_path):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "r0.pt", weights_only=False)
    b = torch.load(tmp_path / "r1.pt", weights_only=False)
    assert a["counts"].tolist() == b["counts"].tolist() == [5, 9]
    assert sorted(a["means"]) == sorted(b["means"]) == [0, 1, 2, 3, 4, 5]
    for w in a["means"]:
        np.testing.assert_array_equal(a["means"][w], b["means"][w])


def _block_worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        D, cap = 7, 20
        rng = np.random.default_rng(200 + rank)
        n_local = 6 + 5 * rank
        table = torch.zeros((world, cap + 1, D), dtype=torch.float32)
        rows = torch.from_numpy(rng.normal(size=(n_local, D)).astype(np.float32))
        table[rank, 1:n_local + 1] = rows
        hdr = torch.tensor([n_local, 0], dtype=torch.int32).view(torch.float32)
        table[rank, 0, :2] = hdr
        g = plate.BlockGatherer(backend="torch")
        g.gather(table)
        counts = plate.block_counts(table)
        assert counts.tolist() == [6 + 5 * r for r in range(world)]
        assert torch.equal(table[rank, 1:n_local + 1], rows)
        torch.save(table, os.path.join(tmp, f"b{rank}.pt"))
        with pytest.raises(ValueError):
            g.gather(table[:, :, :3])
    finally:
        dist.destroy_process_group()


def test_block_gather_protocol_gloo_world2(tmp_path):
    """The header-led block exchange bench.py uses (counts travel in the blocks, no host sync)."""
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_block_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a = torch.load(tmp_path / "b0.pt")
    b = torch.load(tmp_path / "b1.pt")
    assert torch.equal(a, b)


def test_illumination_units_are_dealt_once_and_evenly():
    """K2 shards by (plate, channel), not by field: config 4 = 16 plate-timepoints
    x 5 channels = 80 units over 8 ranks, no collective."""
    chans = ["DNA", "ER", "RNA", "AGP", "Mito"]
    seen = []
    for r in range(8):
        mine = plate.shard_units(16, chans, r, 8)
        assert len(mine) == 10
        seen += mine
    assert sorted(seen) == sorted((p, c) for p in range(16) for c in chans)
    sizes = [len(plate.shard_units(3, chans, r, 4)) for r in range(4)]
    assert sum(sizes) == 15 and max(sizes) - min(sizes) <= 1
    assert plate.shard_units(1, chans, 0, 1) == [(0, c) for c in chans]
    with pytest.raises(ValueError):
        plate.shard_units(2, chans, 4, 4)
