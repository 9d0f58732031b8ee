"""Drop-in for MaxProjection.py: same functions, same flags, same output keys; the
elementwise z-max (MaxProjection.py:45) runs in ips_preprocess_fused on the GPU.

``max_projection`` keeps the reference's per-group contract.  ``max_project_chunk`` is what
the CLI loop uses: all channels of one field (C x Z planes) in a single fused launch.
"""
import argparse
import io
import logging
import posixpath
from io import StringIO

import pandas as pd

from . import storage, tiffio

logging.basicConfig(level=logging.INFO)
logger = logging.getLogger(__name__)


def modify_imagepath(filepath):
    """'Images' path component -> 'ImagesStacked' (MaxProjection.py:16-22)."""
    parts = filepath.split('/')
    if 'Images' not in parts:
        return filepath
    parts[parts.index('Images')] = 'ImagesStacked'
    return '/'.join(parts)


def read_csv_from_s3(bucket_name, file_key, s3_client=None):
    """Data-set CSV with ';' or ',' sniffed from the first KiB (MaxProjection.py:24-31)."""
    s3 = s3_client or storage.client()
    content = s3.get_object(Bucket=bucket_name, Key=file_key)['Body'].read().decode('utf-8')
    return pd.read_csv(StringIO(content), sep=storage.sniff_delimiter(content))


def _fetch(image_key, bucket_name, s3_client):
    return s3_client.get_object(Bucket=bucket_name, Key=image_key)['Body'].read()


def _decode_group(blobs, keys):
    """Encoded planes of one channel -> uint16 CUDA tensor [Z][H][W] (TIFF strips decoded on the
    device); ValueError on a shape mismatch like MaxProjection.py:42-43."""
    try:
        return tiffio.load_planes(blobs)
    except ValueError as e:
        if "shape mismatch" in str(e):
            raise ValueError(f"Image shape mismatch in group: {keys}")
        raise


def _load_group(keys, bucket_name, s3_client):
    return _decode_group([_fetch(k, bucket_name, s3_client) for k in keys], keys)


def _project(stack_czhw):
    """[C][Z][H][W] uint16 (device) -> [C][H][W] uint16 (host, through a page-locked buffer) via the CUDA kernel."""
    import torch
    from .. import ops
    dev = ops.preprocess_fused(stack_czhw[None].contiguous(), None, bin=1, want_binned=False)["maxproj"][0]
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
    host.copy_(dev, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return host.numpy()


def max_projection(image_group, bucket_name, s3_client):
    """Max-project the planes of one channel and upload the result (MaxProjection.py:33-52).
    Raises ValueError when the planes differ in shape, like the reference (:42-43)."""
    max_proj = _project(_load_group(image_group, bucket_name, s3_client)[None])[0]
    s3_client.upload_fileobj(io.BytesIO(tiffio.encode(max_proj)), bucket_name, modify_imagepath(image_group[0]))


def max_project_chunk(groups, bucket_name, s3_client):
    """groups: list over channels of lists over planes of keys.  One launch for the field;
    channels whose planes fail to load or mismatch are reported like the reference does
    (logged, the other channels still go through).  Returns the number written."""
    import torch
    blobs = {}
    for j, group in enumerate(groups):
        try:
            blobs[j] = [_fetch(k, bucket_name, s3_client) for k in group]
        except Exception as e:
            logger.error(f"Error processing group {j}: {e}")
    stacks, ok = [], []
    try:                                        # the whole field in one device decode
        flat = tiffio.decode_to_device([b for j in sorted(blobs) for b in blobs[j]])
        at = 0
        for j in sorted(blobs):
            stacks.append(flat[at:at + len(blobs[j])])
            ok.append(j)
            at += len(blobs[j])
    except Exception:                           # mixed shapes, a damaged or foreign file: channel by channel
        stacks, ok = [], []
        for j in sorted(blobs):
            try:
                stacks.append(_decode_group(blobs[j], groups[j]))
                ok.append(j)
            except Exception as e:
                logger.error(f"Error processing group {j}: {e}")
    written = 0
    by_shape = {}
    for j, st in zip(ok, stacks):
        by_shape.setdefault(tuple(st.shape), []).append((j, st))
    for items in by_shape.values():
        # a failure while projecting, encoding or uploading is that channel group's alone, as in the
        # reference's loop (logged, the plate goes on; MaxProjection.py:85-93)
        try:
            proj = _project(torch.stack([st for _, st in items]))
        except Exception as e:
            for j, _ in items:
                logger.error(f"Error processing group {j}: {e}")
            continue
        for (j, _), mp in zip(items, proj):
            try:
                s3_client.upload_fileobj(io.BytesIO(tiffio.encode(mp)), bucket_name, modify_imagepath(groups[j][0]))
                written += 1
            except Exception as e:
                logger.error(f"Error processing group {j}: {e}")
    return written


def _chunks_of(df, num_channels, num_planes):
    """(plate, groups) per complete chunk of C*Z rows, in the reference's order (MaxProjection.py:75-90):
    plane-major, channel-minor -- channel j, plane p is row j + p*C; incomplete tail chunks skipped."""
    group_size = num_channels * num_planes
    for plate in df['PlateID'].unique():
        sub = df[df['PlateID'] == plate]
        paths = [posixpath.join(str(a), str(b)) for a, b in zip(sub['Image_PathName'], sub['Image_FileName'])]
        for i in range(0, len(sub), group_size):
            if len(sub) - i < group_size:
                logger.warning(f"Skipping incomplete chunk in plate {plate} at index {i}")
                continue
            yield plate, [[paths[i + j + p * num_channels] for p in range(num_planes)] for j in range(num_channels)]


class _InFlight:
    """One batch on the device: everything up to the device->host copy of the projections is queued on
    ``stream``; ``finish`` waits for it, checks the decoder's status and hands the files to the writers."""

    def __init__(self, batch, chunks, stream, host, check, shape, keep):
        self.batch, self.chunks, self.stream, self.host, self.check, self.shape, self.keep = \
            batch, chunks, stream, host, check, shape, keep


def _enqueue_staged(batch, chunks, num_channels, num_planes, stream, out_ring, trace=None):
    """Queue one staged batch of fields on ``stream`` without waiting for the device: one host->device copy
    of the (still compressed) bytes, strips decoded on the device, ONE fused z-max launch for all fields and
    channels, one device->host copy of the projections.  Raises when the batch is not uniform (the caller
    then goes field by field, which reports per channel group like the reference)."""
    import torch
    from .. import ops
    from . import batchio
    if not batch.ok():
        raise ValueError("a file of the batch could not be staged")
    import time
    with torch.cuda.stream(stream):
        t0 = time.perf_counter()
        src = batchio.to_device(batch)
        if trace is not None:
            trace.add("  queue: host->device copy", t0)
        t0 = time.perf_counter()
        planes, check = tiffio.decode_staged(src, batch.infos, batch.bases, defer_check=True)      # [B*C*Z][H][W]
        if trace is not None:
            trace.add("  queue: strip tables + decode", t0)
        t0 = time.perf_counter()
        B = len(chunks)
        H, W = planes.shape[1:]
        raw = planes.view(B, num_channels, num_planes, H, W)
        proj = ops.preprocess_fused(raw, None, bin=1, want_binned=False)["maxproj"]                # [B][C][H][W]
        if trace is not None:
            trace.add("  queue: projection", t0)
        t0 = time.perf_counter()
        host = out_ring.take((B, num_channels, H, W))
        if trace is not None:
            trace.add("  queue: wait for a free output buffer", t0)
        t0 = time.perf_counter()
        host.copy_(proj, non_blocking=True)
        if trace is not None:
            trace.add("  queue: device->host copy", t0)
    return _InFlight(batch, chunks, stream, host, check, (B, num_channels, H, W), (src, planes, proj))


def _finish_staged(job, num_channels, bucket_name, s3_client, writers, out_ring):
    job.stream.synchronize()
    job.check()
    B, _, H, W = job.shape
    head, tail = tiffio.plain_u16_parts(H, W)
    pixels = job.host.numpy()
    jobs = []
    for b, (_, groups) in enumerate(job.chunks):
        for j in range(num_channels):
            key = modify_imagepath(groups[j][0])
            jobs.append(writers.submit(storage.upload_parts, s3_client, [head, memoryview(pixels[b, j]).cast("B"), tail],
                                       bucket_name, key))
    out_ring.busy(job.host, jobs)
    job.keep = None
    return B * num_channels


class _PinnedRing:
    """A few page-locked output buffers; one is handed out again only after its writes finished."""

    def __init__(self, n=3):
        self._bufs, self._jobs, self._n, self._at, self._slot_of = {}, {}, n, 0, {}

    def take(self, shape):
        import torch
        k = self._at % self._n
        self._at += 1
        for f in self._jobs.pop(k, []):
            f.result()
        t = self._bufs.get(k)
        count = 1
        for x in shape:
            count *= x
        if t is None or t.numel() < count:
            t = self._bufs[k] = torch.empty((count,), dtype=torch.uint16, pin_memory=True)
        view = t[:count].view(shape)
        self._slot_of[view.data_ptr()] = k
        return view

    def busy(self, tensor, jobs):
        self._jobs[self._slot_of[tensor.data_ptr()]] = jobs


def run(bucket_data_set, data_set, num_channels, num_planes, bucket_images, s3_client=None, batch_fields=4,
        threads=8):
    """The CLI loop of MaxProjection.py:64-95 at plate scale: the chunks (one field each) are staged
    ``batch_fields`` at a time by reader threads (scripts/batchio.py), projected with one launch per
    batch, and written by writer threads while the next batch is being read."""
    from . import batchio
    s3_client = s3_client or storage.client()
    df = read_csv_from_s3(bucket_data_set, data_set, s3_client)
    chunks = list(_chunks_of(df, num_channels, num_planes))
    plates = [c[0] for c in chunks]
    if not chunks:                                 # nothing to project: no device, no threads
        for plate in df['PlateID'].unique():
            logger.info(f"Plate {plate} finished! Check images in bucket.")
        return 0

    def batches():
        for i in range(0, len(chunks), batch_fields):
            part = chunks[i:i + batch_fields]
            yield part, [storage.source_of(s3_client, bucket_images, k) for _, groups in part for g in groups for k in g]

    import time
    import torch
    loader = batchio.BatchLoader(batches(), threads=threads, depth=4)
    writers = batchio.Writers(threads=max(2, threads))          # 47 MB of projections per field leave through these
    ring = _PinnedRing(4)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    trace = batchio.Trace()
    total = 0
    flying = []                                   # batches queued on the device, oldest first (at most two)

    def land(job):
        nonlocal total
        t0 = time.perf_counter()
        try:
            total += _finish_staged(job, num_channels, bucket_images, s3_client, writers, ring)
        except Exception as why:                  # a damaged strip: that batch again, field by field
            logger.info(f"batch of {len(job.chunks)} fields goes field by field ({why})")
            for _, groups in job.chunks:
                total += max_project_chunk(groups, bucket_images, s3_client)
        finally:
            loader.release(job.batch)
        trace.add("wait for the device + hand to writers", t0)

    it = iter(loader)
    k = 0
    while True:
        t0 = time.perf_counter()
        batch = next(it, None)
        trace.add("wait for staged files", t0)
        if batch is None:
            break
        part = batch.tag
        t0 = time.perf_counter()
        try:
            flying.append(_enqueue_staged(batch, part, num_channels, num_planes, streams[k % 2], ring, trace))
            k += 1
        except Exception as why:                      # mixed shapes, foreign formats, a missing file: field by field
            logger.info(f"batch of {len(part)} fields goes field by field ({why})")
            for _, groups in part:
                total += max_project_chunk(groups, bucket_images, s3_client)
            loader.release(batch)
        trace.add("queue copy + decode + project + read back", t0)
        while len(flying) > 1:                        # batch i runs on the device while batch i - 1 is written
            land(flying.pop(0))
    while flying:
        land(flying.pop(0))
    t0 = time.perf_counter()
    writers.close()
    trace.add("wait for writers", t0)
    trace.report(logger, "MaxProjection")
    for e in writers.errors:
        logger.error(f"Error writing a projection: {e}")
    total -= len(writers.errors)
    for plate in dict.fromkeys(plates):
        logger.info(f"Plate {plate} finished! Check images in bucket.")
    return total


def build_parser():
    parser = argparse.ArgumentParser(description="Process image plates using ImageJ and upload results to S3.")
    parser.add_argument("--bucket_data_set", type=str, required=True, help="S3 bucket containing the data set.")
    parser.add_argument("--data_set", type=str, required=True,
                        help="Data set key location containing per PlateID images to process containing 'ChannelName', "
                             "'ChannelID', 'Image_FileName', 'Image_PathName', 'FieldID', 'PlaneID', 'PlateID', 'Row', "
                             "'Col', 'Timestamp'.")
    parser.add_argument("--channels", type=int, required=True, help="Number of channels per group")
    parser.add_argument("--planes", type=int, required=True, help="Number of planes per channel")
    parser.add_argument("--bucket_images", type=str, required=True, help="S3 bucket containing the raw images to max project.")
    return parser


if __name__ == "__main__":
    a = build_parser().parse_args()
    run(a.bucket_data_set, a.data_set, a.channels, a.planes, a.bucket_images)
