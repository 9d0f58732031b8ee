"""The handful of S3 calls the reference scripts make, over boto3 or a local directory.

Reference call sites: get_object / upload_fileobj (MaxProjection.py:24-27, :36-37, :52),
bucket.objects.filter / obj.get / bucket.put_object (Image_re-binning.py:41-55).
"""
import io
import os


class _Body:
    def __init__(self, data):
        self._data = data

    def read(self):
        return self._data


class LocalObject:
    def __init__(self, root, bucket, key):
        self._root, self._bucket, self.key = root, bucket, key

    def get(self):
        with open(os.path.join(self._root, self._bucket, self.key), "rb") as f:
            return {"Body": _Body(f.read())}


class _LocalObjects:
    def __init__(self, root, bucket):
        self._root, self._bucket = root, bucket

    def filter(self, Prefix=""):
        base = os.path.join(self._root, self._bucket)
        out = []
        for d, _, files in os.walk(base):
            for name in files:
                key = os.path.relpath(os.path.join(d, name), base).replace(os.sep, "/")
                if key.startswith(Prefix):
                    out.append(LocalObject(self._root, self._bucket, key))
        return sorted(out, key=lambda o: o.key)


class LocalBucket:
    def __init__(self, root, name):
        self._root, self.name = root, name
        self.objects = _LocalObjects(root, name)

    def put_object(self, Key, Body, ContentType=None):
        path = os.path.join(self._root, self.name, Key)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(Body)


class LocalClient:
    """Subset of boto3's S3 client + resource over ``root/<bucket>/<key>``."""

    def __init__(self, root):
        self.root = root

    def get_object(self, Bucket, Key):
        path = os.path.join(self.root, Bucket, Key)
        if not os.path.exists(path):
            raise FileNotFoundError("no such key: s3://%s/%s" % (Bucket, Key))
        with open(path, "rb") as f:
            return {"Body": _Body(f.read())}

    def upload_fileobj(self, fileobj, Bucket, Key):
        path = os.path.join(self.root, Bucket, Key)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            f.write(fileobj.read())

    def put_object(self, Bucket, Key, Body, ContentType=None):
        self.upload_fileobj(io.BytesIO(Body), Bucket, Key)

    # not part of boto3: what the plate-scale loops use when the store is a directory
    def local_path(self, Bucket, Key):
        """Path of an object (readers ``readinto`` page-locked staging memory straight from it)."""
        return os.path.join(self.root, Bucket, Key)

    def upload_parts(self, parts, Bucket, Key):
        """Write an object given as a sequence of buffers (no concatenation)."""
        path = os.path.join(self.root, Bucket, Key)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "wb") as f:
            for part in parts:
                f.write(part)

    def Bucket(self, name):
        return LocalBucket(self.root, name)


def local_root():
    return os.environ.get("IPS_STORAGE_ROOT")


def client():
    """boto3.client('s3') unless IPS_STORAGE_ROOT selects the directory-backed stand-in."""
    root = local_root()
    if root:
        return LocalClient(root)
    import boto3
    return boto3.client("s3")


def resource():
    root = local_root()
    if root:
        return LocalClient(root)
    import boto3
    return boto3.resource("s3")


def source_of(client, bucket, key):
    """A BatchLoader source for an object: its path on directory-backed storage, else a fetcher."""
    if hasattr(client, "local_path"):
        return client.local_path(bucket, key)
    return lambda: client.get_object(Bucket=bucket, Key=key)['Body'].read()


def upload_parts(client, parts, bucket, key):
    if hasattr(client, "upload_parts"):
        client.upload_parts(parts, bucket, key)
    else:
        client.upload_fileobj(io.BytesIO(b"".join(bytes(p) for p in parts)), bucket, key)


def sniff_delimiter(text):
    """';' or ',' as the reference's readers decide it: csv.Sniffer over the first KiB
    (MaxProjection.py:29-30, Normalize_CP_ami.py:25-26); when the sniffer cannot decide (wide
    tables whose first KiB holds less than two full lines) the header line's own count decides."""
    import csv
    sample = text[:1024]
    try:
        return csv.Sniffer().sniff(sample, delimiters=";,").delimiter
    except csv.Error:
        header = text.split("\n", 1)[0]
        return ";" if header.count(";") > header.count(",") else ","
