"""Plate-scale file I/O for the drop-in scripts: a host thread pool that keeps the GPU fed.

The reference's scripts fetch and decode one image at a time on one thread (MaxProjection.py:36-40,
Image_re-binning.py:48) and spend nearly all of their time there.  Here a pool of reader threads
works a few batches ahead of the device:

  reader threads   read every file of a batch straight into one page-locked staging buffer
                   (``readinto`` on local storage -- no intermediate ``bytes`` --, a copy for clients
                   that only hand out ``bytes``) and parse the TIFF directories;
  device           ONE host->device copy of the staged (still compressed) bytes per batch, LZW strips
                   decoded by the K7 kernel (tiffio.decode_staged), then the arithmetic kernels;
  writer threads   upload / write the results while the next batch is in flight.

Errors keep the reference's convention: a file that cannot be read or decoded fails its own item
(logged by the caller), never the plate.
"""
import concurrent.futures as cf
import threading

import numpy as np

from . import tiffio


def _align16(n):
    return (int(n) + 15) // 16 * 16


# Page-locking memory is slow (hundreds of milliseconds per gigabyte): staging slots are kept for the
# life of the process and handed from one loader to the next.
_SLOT_POOL = []
_SLOT_LOCK = threading.Lock()


def _take_slot(nbytes):
    from ..pipeline import pinned_empty
    with _SLOT_LOCK:
        fit = [k for k, s in enumerate(_SLOT_POOL) if s.shape[0] >= nbytes]
        if fit:
            return _SLOT_POOL.pop(min(fit, key=lambda k: _SLOT_POOL[k].shape[0]))
    return pinned_empty((int(nbytes),), np.uint8)


def _give_slot(slot):
    with _SLOT_LOCK:
        _SLOT_POOL.append(slot)
        _SLOT_POOL.sort(key=lambda s: -s.shape[0])
        del _SLOT_POOL[6:]


class Trace:
    """Wall-clock seconds per stage of a plate loop (IPS_IO_TRACE=1 prints them)."""

    def __init__(self):
        import os
        self.on = os.environ.get("IPS_IO_TRACE", "") not in ("", "0")
        self.t = {}

    def add(self, key, t0):
        import time
        if self.on:
            self.t[key] = self.t.get(key, 0.0) + time.perf_counter() - t0

    def report(self, log, what):
        if self.on:
            log.info("%s stages (s): %s", what, ", ".join("%s %.3f" % kv for kv in sorted(self.t.items())))


class StagedBatch:
    """The files of one batch in page-locked memory: ``buf`` (uint8), ``bases`` / ``sizes`` per file,
    ``infos`` (tiffio.parse) or the exception that file raised, ``tag`` (caller's bookkeeping)."""

    def __init__(self, buf, bases, sizes, infos, tag, slot):
        self.buf, self.bases, self.sizes, self.infos, self.tag, self.slot = buf, bases, sizes, infos, tag, slot

    @property
    def used(self):
        return _align16(self.bases[-1] + self.sizes[-1]) if self.bases else 0

    def ok(self):
        return all(isinstance(i, dict) for i in self.infos)


class BatchLoader:
    """Iterate over ``StagedBatch`` objects, ``depth`` batches ahead of the consumer.

    ``batches``: iterable of (tag, [source, ...]); a source is a local file path (str) or a zero-argument
    callable returning ``bytes``.  ``threads`` reader threads.  The consumer must call
    ``release(batch)`` when the device copy of ``batch.buf`` has been issued AND has completed
    (the staging slot is reused)."""

    def __init__(self, batches, threads=8, depth=3, slot_bytes=1 << 26):
        self._batches = iter(batches)
        self._pool = cf.ThreadPoolExecutor(max_workers=max(1, threads), thread_name_prefix="ips-read")
        self._depth = max(1, depth)
        self._slot_bytes = int(slot_bytes)
        self._free = [None] * self._depth       # slots are taken from the process-wide pool when a batch's size is known
        self._lock = threading.Lock()
        self._pending = []          # futures of staged batches, in order
        self._done = False

    # -- one file into the staging buffer (runs on a reader thread) ---------------------------------
    @staticmethod
    def _read_one(source, view):
        """Returns (size, info-or-exception).  ``view``: writable uint8 memoryview slice, large enough."""
        try:
            if isinstance(source, str):
                with open(source, "rb", buffering=0) as f:
                    n = 0
                    while True:
                        got = f.readinto(view[n:])
                        if not got:
                            break
                        n += got
            else:
                data = source()
                n = len(data)
                view[:n] = data
            try:
                return n, tiffio.parse(view[:n])
            except tiffio.Unsupported as why:
                return n, why
        except Exception as e:      # noqa: BLE001 -- the caller logs it against the item
            return 0, e

    def _sizes(self, sources):
        import os
        sizes = []
        for s in sources:
            if isinstance(s, str):
                try:
                    sizes.append(os.path.getsize(s))
                except OSError:
                    sizes.append(0)
            else:
                sizes.append(None)
        return sizes

    def _stage(self, tag, sources, slot):
        sizes = self._sizes(sources)
        if any(z is None for z in sizes):
            # clients that only return bytes: fetch first (in parallel), then lay the files out
            blobs = list(self._pool.map(lambda s: self._safe_fetch(s), sources))
            sizes = [len(b) if isinstance(b, (bytes, bytearray, memoryview)) else 0 for b in blobs]
            sources = [(lambda b=b: b) if not isinstance(b, Exception) else (lambda b=b: (_ for _ in ()).throw(b)) for b in blobs]
        bases, at = [], 0
        for z in sizes:
            bases.append(at)
            at += _align16(z)
        if slot is None or at > slot.shape[0]:
            if slot is not None:
                _give_slot(slot)
            self._slot_bytes = max(self._slot_bytes, at + at // 8)            # room for the slightly larger batches to come
            slot = _take_slot(self._slot_bytes)
        mv = memoryview(slot)
        futs = [self._pool.submit(self._read_one, s, mv[b:b + max(z, 1)]) for s, b, z in zip(sources, bases, sizes)]
        got = [f.result() for f in futs]
        return StagedBatch(slot, bases, [g[0] for g in got], [g[1] for g in got], tag, slot)

    @staticmethod
    def _safe_fetch(source):
        try:
            return source() if callable(source) else open(source, "rb").read()
        except Exception as e:      # noqa: BLE001
            return e

    def _fill(self):
        while not self._done and len(self._pending) < self._depth and self._free:
            try:
                tag, sources = next(self._batches)
            except StopIteration:
                self._done = True
                break
            slot = self._free.pop()
            # staging runs on its own thread so that several batches are being read at once
            t = cf.Future()
            threading.Thread(target=self._run_stage, args=(t, tag, list(sources), slot), daemon=True).start()
            self._pending.append(t)

    def _run_stage(self, fut, tag, sources, slot):
        try:
            fut.set_result(self._stage(tag, sources, slot))
        except BaseException as e:  # noqa: BLE001
            fut.set_exception(e)

    def __iter__(self):
        return self

    def __next__(self):
        self._fill()
        if not self._pending:
            self.close()
            raise StopIteration
        batch = self._pending.pop(0).result()
        self._fill()
        return batch

    def release(self, batch):
        with self._lock:
            if len(self._free) < self._depth:
                self._free.append(batch.slot)
        self._fill()

    def close(self):
        self._pool.shutdown(wait=False)
        with self._lock:
            for slot in self._free:
                if slot is not None:
                    _give_slot(slot)
            self._free = []


def to_device(batch, device="cuda", stream=None):
    """ONE host->device copy of a staged batch (page-locked, asynchronous): uint8 CUDA tensor."""
    import torch
    n = max(batch.used, 16)
    # sizes in steps of 32 MiB: the batches of a plate differ by a few kilobytes, and a caching allocator asked
    # for a new size each time keeps going back to cudaMalloc
    src = torch.empty(((n + (1 << 25) - 1) >> 25 << 25,), dtype=torch.uint8, device=device)[:n]
    src.copy_(torch.from_numpy(batch.buf[:n]), non_blocking=True)
    return src


class Writers:
    """Background writers (uploads / file writes) with bounded backlog; ``drain`` re-raises nothing:
    every job returns its own error string or None, collected in ``errors``."""

    def __init__(self, threads=4, backlog=64):
        self._pool = cf.ThreadPoolExecutor(max_workers=max(1, threads), thread_name_prefix="ips-write")
        self._sem = threading.Semaphore(backlog)
        self._futs = []
        self.errors = []

    def submit(self, fn, *args):
        self._sem.acquire()

        def job():
            try:
                fn(*args)
            except Exception as e:  # noqa: BLE001
                self.errors.append(str(e))
            finally:
                self._sem.release()
        fut = self._pool.submit(job)
        self._futs.append(fut)
        if len(self._futs) > 4096:
            self._futs = [f for f in self._futs if not f.done()]
        return fut

    def drain(self):
        for f in self._futs:
            f.result()
        self._futs = []

    def close(self):
        self.drain()
        self._pool.shutdown(wait=True)
