"""Plate-level host logic: sharding fields over ranks, dense object rows, the one all-gather
and the well aggregation that follows it (north_star).

Sharding: fields are independent units; all sites of a well go to one rank
(``rank = well_index mod world``), so the per-field path needs no communication.  After the
last field of a plate-timepoint each rank packs its object rows (``pack_rows``), the ranks
exchange them with ONE all-gather (``RowGatherer``: NCCL through libips.so on GPUs; the same
padded-block protocol over ``torch.distributed`` for the gloo tests), and every rank computes
the per-well means (``well_means``) that Normalize_CP_ami.py:126 computes with pandas.
"""
import ctypes as C

import torch

from . import capi

ROW_PREFIX = ("well", "field", "label", "area", "y0", "x0", "y1", "x1", "cy", "cx")


def row_columns(channels):
    cols = list(ROW_PREFIX)
    for ch in channels:
        cols += [f"sum_{ch}", f"mean_{ch}", f"std_{ch}", f"min_{ch}", f"max_{ch}"]
    return cols


def well_name(index, n_cols=24):
    """0-based well index -> 'A01' ... 'P24' (384-well plate: 16 rows x 24 columns)."""
    r, c = divmod(int(index), n_cols)
    return "%s%02d" % (chr(ord("A") + r), c + 1)


def shard_wells(n_wells, rank, world):
    """Wells owned by ``rank``: every ``world``-th well (balanced to within one well)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return list(range(rank, n_wells, world))


def shard_units(n_plates, channels, rank=0, world=1):
    """Illumination estimation reduces ACROSS the fields of a plate, so its unit of sharding is the
    (plate, channel) pair, not the field: the units, plate-major, are dealt
    round-robin over the ranks -- no collective, balanced to within one unit.  Returns this rank's
    [(plate_index, channel), ...]."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    units = [(p, c) for p in range(int(n_plates)) for c in channels]
    return units[rank::world]


def plate_fields(n_wells, sites_per_well, rank=0, world=1):
    """(well, site) pairs of the rank's shard, wells ascending, sites 1..n inside a well."""
    return [(w, s) for w in shard_wells(n_wells, rank, world) for s in range(1, sites_per_well + 1)]


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def pack_rows(ints, flts, n_objects, field_well, field_base=0, out=None):
    """Padded per-field outputs -> dense float32 rows [total][10 + 5C] (device).

    Returns (rows buffer [F * Nmax][D], total as a 0-d int64 device tensor).  Only the first
    ``total`` rows are valid.  ``field_well`` [F] int32 = well index of each field.
    """
    F, n_max, _ = ints.shape
    nf = flts.shape[2]
    Cn = (nf - 2) // 5
    dev = ints.device
    if not (ints.is_cuda and flts.is_cuda and n_objects.is_cuda and field_well.is_cuda):
        raise ValueError("pack_rows works on device tensors (there is no CPU path)")
    if field_well.dtype != torch.int32 or field_well.shape[0] != F:
        raise ValueError("field_well must be int32 [F]")
    with torch.cuda.device(dev):
        rows = out if out is not None else torch.empty((F * n_max, 8 + nf), dtype=torch.float32, device=dev)
        total = torch.zeros((), dtype=torch.int64, device=dev)
        ws = torch.empty(max(int(capi.call("ips_pack_rows_workspace_bytes", F)), 16), dtype=torch.uint8, device=dev)
        capi.call("ips_pack_rows", _ptr(ints), _ptr(flts), _ptr(n_objects), _ptr(field_well), int(field_base),
                  _ptr(rows), _ptr(total), n_max, Cn, F, _ptr(ws), ws.numel(), _stream(dev))
    return rows, total


def pack_rows_block(ints, flts, n_objects, field_well, block, field_base=0, ws=None):
    """Padded per-field outputs -> one header-led block ``block`` [block_rows][10 + 5C] float32
    (device): row 0 carries the row count (two 32-bit words), rows 1..count the objects.  No host
    synchronisation; ``block_rows - 1 >= F * Nmax``.  This is the unit ``BlockGatherer`` moves."""
    F, n_max, _ = ints.shape
    nf = flts.shape[2]
    Cn = (nf - 2) // 5
    dev = ints.device
    if not (ints.is_cuda and flts.is_cuda and n_objects.is_cuda and field_well.is_cuda and block.is_cuda):
        raise ValueError("pack_rows_block works on device tensors (there is no CPU path)")
    if block.dtype != torch.float32 or block.dim() != 2 or block.shape[1] != 8 + nf or not block.is_contiguous():
        raise ValueError("block must be contiguous float32 [block_rows][%d]" % (8 + nf))
    with torch.cuda.device(dev):
        if ws is None:
            ws = torch.empty(max(int(capi.call("ips_pack_rows_workspace_bytes", F)), 16), dtype=torch.uint8, device=dev)
        capi.call("ips_pack_rows_block", _ptr(ints), _ptr(flts), _ptr(n_objects), _ptr(field_well), int(field_base),
                  _ptr(block), block.shape[0], n_max, Cn, F, _ptr(ws), ws.numel(), _stream(dev))
    return block


def block_counts(table):
    """Header counts of a [n_blocks][block_rows][D] table of header-led blocks (int64).  Device
    tables are read by ``ips_block_counts``; host tables (gloo tests) by reinterpreting the bytes."""
    nb, br, D = table.shape
    if not table.is_cuda:
        return table[:, 0, :2].contiguous().view(torch.int32).to(torch.int64).bitwise_and(0xffffffff).mul(
            torch.tensor([1, 1 << 32])).sum(1)
    out = torch.empty((nb,), dtype=torch.int64, device=table.device)
    with torch.cuda.device(table.device):
        capi.call("ips_block_counts", _ptr(table), _ptr(out), nb, br, D, _stream(table.device))
    return out


class BlockGatherer:
    """The one collective, without a host round trip: ``table`` [world][block_rows][D] float32
    with this rank's header-led block (``pack_rows_block``) already at ``table[rank]``; after
    ``gather`` every rank holds every block.  ONE fixed-size all-gather; the row counts travel
    inside the blocks.

    backend "ips": ``ips_allgather_blocks`` (ncclAllGather inside libips.so on its own
    communicator; the unique id travels over the caller's torch.distributed group).  backend
    "torch": the same exchange with torch.distributed on whatever device the tensors live on
    (the world_size-2 gloo tests; not a product path on GPUs).
    """

    def __init__(self, group=None, backend=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.backend = backend or ("ips" if torch.cuda.is_available() else "torch")
        self._comm = None
        if self.backend == "ips" and self.world > 1:
            self._comm = _make_comm(dist, group, self.rank, self.world)

    def gather(self, table):
        if table.dim() != 3 or table.shape[0] != self.world or table.dtype != torch.float32 or not table.is_contiguous():
            raise ValueError("table must be contiguous float32 [world][block_rows][D]")
        if self.world == 1:
            return table
        if self.backend == "ips":
            dev = table.device
            with torch.cuda.device(dev):
                capi.call("ips_allgather_blocks", self._comm, _ptr(table), table.shape[1], table.shape[2] * 4,
                          _stream(dev))
            return table
        self.dist.all_gather([table[r] for r in range(self.world)], table[self.rank].clone(), group=self.group)
        return table

    def close(self):
        if self._comm is not None:
            capi.call("ips_comm_destroy", self._comm)
            self._comm = None


class _DeviceBlock:
    """One cudaMalloc of its own, viewed as a torch tensor through __cuda_array_interface__."""

    def __init__(self, shape, dtype=torch.float32):
        self.shape = tuple(int(x) for x in shape)
        n = 1
        for x in self.shape:
            n *= x
        itemsize = torch.empty((), dtype=dtype).element_size()
        self.ptr = C.c_void_p()
        capi.call("ips_device_alloc", C.byref(self.ptr), n * itemsize)
        typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int64: "<i8", torch.uint8: "|u1"}[dtype]
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr.value, False),
                                         "version": 2, "strides": None}

    def __del__(self):
        try:
            if self.ptr:
                capi.fn("ips_device_free")(self.ptr)
        except Exception:
            pass


def exportable_zeros(shape, dtype=torch.float32):
    """A zeroed device tensor on an allocation of its own (what ``PeerPusher`` can export)."""
    blk = _DeviceBlock(shape, dtype)
    t = torch.as_tensor(blk, device=torch.device("cuda", torch.cuda.current_device()))
    t._ips_block = blk                            # keeps the allocation alive as long as the tensor
    t.zero_()
    return t


class PeerPusher:
    """The bulk of the row gather over NVSwitch peer memory (``ips_peer_push``): ``table`` is one
    device allocation of the same shape on every rank; ``push(view)`` stores a contiguous view of it
    into the same place of every peer's table with copy-engine transfers on the current stream --
    no kernel, no SM.  ``BlockGatherer.gather`` on a small per-rank table at the end of the plate is
    the barrier.  Raises IpsError when CUDA IPC is not available (the caller then gathers the blocks
    with ``BlockGatherer`` alone)."""

    def __init__(self, table, group=None):
        import torch.distributed as dist
        self.table = table
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        n = int(capi.call("ips_ipc_handle_bytes"))
        buf = (C.c_char * n)()
        if not hasattr(table, "_ips_block"):
            raise ValueError("the table must come from plate.exportable_zeros (an allocation of its own)")
        base = table.data_ptr()
        capi.call("ips_ipc_export", C.c_void_p(base), C.cast(buf, C.c_void_p), n)
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(buf), group=group)
        blob = (C.c_char * (n * self.world)).from_buffer_copy(b"".join(handles))
        h = C.c_void_p()
        capi.call("ips_peer_table_open", C.byref(h), C.cast(blob, C.c_void_p), self.rank, self.world, C.c_void_p(base))
        self._h = h

    def push(self, view):
        if not view.is_contiguous():
            raise ValueError("push takes a contiguous view of the table")
        off = view.data_ptr() - self.table.data_ptr()
        nbytes = view.numel() * view.element_size()
        if off < 0 or off + nbytes > self.table.numel() * self.table.element_size():
            raise ValueError("the view is not inside the table")
        dev = self.table.device
        with torch.cuda.device(dev):
            capi.call("ips_peer_push", self._h, off, nbytes, _stream(dev))

    def close(self):
        if self._h:
            capi.call("ips_peer_table_close", self._h)
            self._h = None


class PlateRowExchange:
    """The plate's one all-gather as the rank's fields are finished: object rows leave chunk by chunk
    while the remaining fields are processed, and ``finish`` returns the gathered table.

    ``table`` [n_chunks][world][block_rows][D] float32 of header-led blocks (``pack_rows_block``),
    block_rows = chunk_fields * n_max + 1.  ``submit(g, ...)`` packs the rows of chunk ``g`` behind
    their count and moves the block on a high-priority side stream, without ever blocking the
    launching thread on work it has just queued:

    * peer push (the default where CUDA IPC works): copy-engine stores into every peer's table
      (``PeerPusher``).  The copy of chunk ``g`` is issued one ``submit`` later; if its header -- read
      back to page-locked memory behind the pack -- has arrived by then, only the rows that exist
      travel (a field holds fewer objects than ``n_max``), otherwise the block goes at capacity
      (``exact_push="wait"`` waits for the header instead; ``False`` pushes at capacity at once).  The
      last chunk goes at capacity from ``finish``, which adds ONE small NCCL all-gather (the per-chunk
      counts of every rank): queued behind this rank's pushes, complete when every rank's has started
      -- the barrier that publishes the pushes.
    * otherwise one fixed-size ``ncclAllGather`` per chunk (``BlockGatherer``).

    All ranks take the same path (agreed with one all-reduce at construction).  One rank: the rows are
    only packed.  Nothing here synchronises with the device except ``counts()`` and ``release()``.
    Between two plates on the same exchange every rank calls ``release()`` (or passes a barrier of its
    own) once it has consumed the table: the tables are written by the peers.
    """

    def __init__(self, n_chunks, chunk_fields, n_max, channels, group=None, peer_push=True, exact_push=True,
                 device=None):
        import torch.distributed as dist
        self.dist, self.group = dist, group
        live = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if live else 0
        self.world = dist.get_world_size(group) if live else 1
        self.dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_chunks, self.chunk_fields, self.n_max = int(n_chunks), int(chunk_fields), int(n_max)
        self.D = 10 + 5 * int(channels)
        self.block_rows = self.chunk_fields * self.n_max + 1
        shape = (self.n_chunks, self.world, self.block_rows, self.D)
        self.pusher, self.transport = None, "n/a (one rank)"
        self.table = None
        with torch.cuda.device(self.dev):
            if self.world > 1 and peer_push:
                try:
                    self.table = exportable_zeros(shape)
                    self.pusher = PeerPusher(self.table, group)
                    self.transport = ("copy-engine stores into the peers' tables (CUDA IPC over NVLink), "
                                      "one NCCL all-gather of the chunk counts as barrier")
                except Exception as e:                          # IPC not available in this container
                    self.pusher, self.table = None, None
                    self.transport = "peer push unavailable (%s): ncclAllGather per chunk" % (str(e)[:80])
            if self.world > 1:
                ok = torch.tensor([1 if self.pusher is not None else 0], device=self.dev)
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
                if int(ok.item()) == 0 and self.pusher is not None:
                    self.pusher.close()
                    self.pusher, self.table = None, None
                    self.transport = "peer push unavailable on another rank: ncclAllGather per chunk"
                if self.pusher is None and not peer_push:
                    self.transport = "ncclAllGather per chunk (peer push disabled)"
            if self.table is None:
                self.table = torch.zeros(shape, dtype=torch.float32, device=self.dev)
            self.gatherer = BlockGatherer(group) if self.world > 1 else None
            self.exact_push = exact_push if self.pusher is not None else False
            self.chunk_counts = torch.zeros((self.world, self.n_chunks, 2), dtype=torch.float32, device=self.dev)
            self.stream = torch.cuda.Stream(device=self.dev, priority=-1)   # its few CTAs must not queue behind a full grid
            self._done = [torch.cuda.Event() for _ in range(self.n_chunks)]
            self._hdr_host = torch.zeros((self.n_chunks, 2), dtype=torch.int32).pin_memory()
            self._hdr_ready = [torch.cuda.Event() for _ in range(self.n_chunks)]
            self._pending = None
            self.pushed_bytes = 0
            n = int(capi.call("ips_pack_rows_workspace_bytes", self.chunk_fields))
            self._ws = torch.empty(max(n, 16), dtype=torch.uint8, device=self.dev)

    def _push(self, g, exact):
        blk = self.table[g, self.rank]
        if exact and (self.exact_push == "wait" or self._hdr_ready[g].query()):
            self._hdr_ready[g].synchronize()
            n = (int(self._hdr_host[g, 0]) & 0xffffffff) | (int(self._hdr_host[g, 1]) << 32)
            blk = blk[:min(max(n, 0), self.block_rows - 1) + 1]
        self.pushed_bytes += blk.numel() * 4 * (self.world - 1)
        self.pusher.push(blk)

    def submit(self, g, ints, flts, n_objects, field_well, field_base=0):
        """The padded rows of chunk ``g``'s fields are final on the current stream: pack and move them."""
        if not (0 <= g < self.n_chunks) or ints.shape[0] != self.chunk_fields or ints.shape[1] != self.n_max:
            raise ValueError("chunk %d with %s rows does not fit the exchange" % (g, tuple(ints.shape[:2])))
        self._done[g].record()
        with torch.cuda.stream(self.stream):
            if self._pending is not None:                       # packed one submit ago: its header is (most likely) here
                self._push(self._pending, True)
                self._pending = None
            self.stream.wait_event(self._done[g])
            pack_rows_block(ints, flts, n_objects, field_well, self.table[g, self.rank], field_base=field_base, ws=self._ws)
            if self.pusher is not None:
                if self.exact_push:
                    self._hdr_host[g].copy_(self.table[g, self.rank, 0, :2].view(torch.int32), non_blocking=True)
                    self._hdr_ready[g].record()
                    self._pending = g
                else:
                    self._push(g, False)
            elif self.gatherer is not None:
                self.gatherer.gather(self.table[g])

    def finish(self):
        """Everything submitted is on every rank when the current stream reaches this point.
        Returns the table as [n_chunks * world][block_rows][D] (``WellAggregator.add_blocks``)."""
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                if self._pending is not None:
                    self._push(self._pending, False)            # at capacity: no wait for its header
                    self._pending = None
                if self.pusher is not None:
                    self.chunk_counts[self.rank].copy_(self.table[:, self.rank, 0, :2])
                    self.gatherer.gather(self.chunk_counts)
        torch.cuda.current_stream(self.dev).wait_stream(self.stream)
        return self.blocks()

    def release(self):
        """This rank is done reading the gathered table.  Returns when every rank is: only then may the
        next plate's ``submit`` calls begin -- a push of the next plate stores into the peers' tables, and a
        peer that is still summing or copying the previous plate would see the new rows.  (One barrier per
        plate, outside the plate's own stream of work; bench.py has its own barrier at that point.)"""
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def blocks(self):
        return self.table.view(self.n_chunks * self.world, self.block_rows, self.D)

    def counts(self):
        """Row counts [n_chunks][world] (int64, device) from the headers of the gathered table."""
        return block_counts(self.blocks()).view(self.n_chunks, self.world)

    def close(self):
        if self.pusher is not None:
            self.pusher.close()
            self.pusher = None
        if self.gatherer is not None:
            self.gatherer.close()
            self.gatherer = None


def _make_comm(dist, group, rank, world):
    n = int(capi.call("ips_comm_unique_id_bytes"))
    buf = (C.c_char * n)()
    if rank == 0:
        capi.call("ips_comm_unique_id", C.cast(buf, C.c_void_p), n)
    box = [bytes(buf)]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = (C.c_char * n).from_buffer_copy(box[0])
    h = C.c_void_p()
    capi.call("ips_comm_create", C.byref(h), C.cast(uid, C.c_void_p), n, rank, world)
    return h


class RowGatherer:
    """The one collective: every rank contributes ``n_local`` rows of ``D`` float32, every rank
    receives all of them as a padded table [world][cap][D] plus the per-rank counts.

    backend "ips": ncclAllGather inside libips.so (its own communicator; the unique id travels
    over the caller's torch.distributed group).  backend "torch": the same protocol with
    torch.distributed collectives on whatever device the tensors live on (used by the
    world_size-2 gloo tests; not a product path on GPUs).
    """

    def __init__(self, cap_per_rank, D, group=None, backend=None):
        import torch.distributed as dist
        self.dist = dist
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cap, self.D = int(cap_per_rank), int(D)
        self.backend = backend or ("ips" if torch.cuda.is_available() else "torch")
        self._comm = None
        if self.backend == "ips" and self.world > 1:
            self._comm = _make_comm(dist, group, self.rank, self.world)

    def gather(self, local_rows, n_local, out=None):
        """local_rows [>= n_local][D] float32 -> (all_rows [world][cap][D], counts [world] int64).

        ``out`` = (all_rows, counts) reuses caller buffers; when ``local_rows`` already is
        ``all_rows[rank]`` (rows packed in place) the staging copy is skipped."""
        n_local = int(n_local)
        if n_local > self.cap:
            raise ValueError("%d rows exceed the per-rank capacity %d" % (n_local, self.cap))
        if local_rows.dtype != torch.float32 or local_rows.dim() != 2 or local_rows.shape[1] != self.D:
            raise ValueError("rows must be float32 [n][%d]" % self.D)
        dev = local_rows.device
        if out is not None:
            all_rows, counts = out
            if tuple(all_rows.shape) != (self.world, self.cap, self.D) or not all_rows.is_contiguous():
                raise ValueError("out rows must be contiguous [world][cap][D]")
        else:
            all_rows = torch.empty((self.world, self.cap, self.D), dtype=torch.float32, device=dev)
            counts = torch.zeros((self.world,), dtype=torch.int64, device=dev)
        if self.world == 1:
            if local_rows.data_ptr() != all_rows.data_ptr():
                all_rows[0, :n_local].copy_(local_rows[:n_local])
            counts.fill_(n_local)
            return all_rows, counts
        if self.backend == "ips":
            with torch.cuda.device(dev):
                capi.call("ips_allgather_rows", self._comm, _ptr(local_rows), n_local, self.D * 4, _ptr(all_rows),
                          _ptr(counts), self.cap, _stream(dev))
            return all_rows, counts
        mine = torch.zeros((self.cap, self.D), dtype=torch.float32, device=dev)
        mine[:n_local].copy_(local_rows[:n_local])
        self.dist.all_gather(list(counts.split(1)), torch.tensor([n_local], dtype=torch.int64, device=dev),
                             group=self.group)
        blocks = [all_rows[r] for r in range(self.world)]
        self.dist.all_gather(blocks, mine, group=self.group)
        return all_rows, counts

    def close(self):
        if self._comm is not None:
            capi.call("ips_comm_destroy", self._comm)
            self._comm = None


def well_ids_of(all_rows, counts):
    """int32 well id per row of a gathered table, -1 for padding."""
    world, cap, D = all_rows.shape
    dev = all_rows.device
    if dev.type != "cuda":
        idx = torch.arange(cap)[None, :] < counts[:, None]
        return torch.where(idx, all_rows[:, :, 0].to(torch.int32), torch.full((), -1, dtype=torch.int32)).reshape(-1)
    well = torch.empty((world * cap,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        capi.call("ips_rows_well_ids", _ptr(all_rows), _ptr(counts), _ptr(well), cap, world, D, _stream(dev))
    return well


def well_means(all_rows, counts, n_wells):
    """Per-well mean of every row column (device): (mean [n_wells][D] float64, count [n_wells])."""
    from . import ops
    world, cap, D = all_rows.shape
    return ops.well_mean(all_rows.reshape(world * cap, D), well_ids_of(all_rows, counts), n_wells)


class WellAggregator:
    """Streaming per-well means: ``add`` gathered row blocks as they arrive (one all-gather
    chunk at a time), ``finalize`` once per plate.  Same arithmetic as ``well_means``."""

    def __init__(self, n_wells, D, device="cuda"):
        self.n_wells, self.D = int(n_wells), int(D)
        self.dev = torch.device(device)
        with torch.cuda.device(self.dev):
            n = int(capi.call("ips_well_mean_workspace_bytes", self.n_wells, self.D))
            self.ws = torch.empty(max(n, 16), dtype=torch.uint8, device=self.dev)
        self.reset()

    def reset(self):
        with torch.cuda.device(self.dev):
            capi.call("ips_well_sums_reset", _ptr(self.ws), self.ws.numel(), self.D, self.n_wells, _stream(self.dev))

    def add(self, all_rows, counts):
        """all_rows [blocks][cap][D] float32 with the first counts[b] rows of block b valid."""
        ids = well_ids_of(all_rows, counts)
        blocks, cap, D = all_rows.shape
        with torch.cuda.device(self.dev):
            capi.call("ips_well_sums_add", _ptr(all_rows), _ptr(ids), blocks * cap, _ptr(self.ws), self.ws.numel(),
                      self.D, self.n_wells, _stream(self.dev))

    def add_blocks(self, table):
        """table [n_blocks][block_rows][D] float32 of header-led blocks (``pack_rows_block`` /
        ``BlockGatherer``): well ids and validity come from the table itself, no id array."""
        nb, br, D = table.shape
        with torch.cuda.device(self.dev):
            capi.call("ips_well_sums_add_blocks", _ptr(table), nb, br, _ptr(self.ws), self.ws.numel(), self.D,
                      self.n_wells, _stream(self.dev))

    def finalize(self):
        with torch.cuda.device(self.dev):
            mean = torch.empty((self.n_wells, self.D), dtype=torch.float64, device=self.dev)
            count = torch.empty((self.n_wells,), dtype=torch.int32, device=self.dev)
            capi.call("ips_well_sums_finalize", _ptr(self.ws), self.ws.numel(), _ptr(mean), _ptr(count), self.D,
                      self.n_wells, _stream(self.dev))
        return mean, count
