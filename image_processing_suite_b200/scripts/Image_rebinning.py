"""Drop-in for Image_re-binning.py: same functions and flags; the LANCZOS resize
(Image_re-binning.py:18) runs in ips_lanczos_resize_u16, bit-exact with Pillow's I;16 path, and
the TIFF strips on either side of it (Image_re-binning.py:17, :19-21) are decoded and LZW-encoded
by the device codec (ips_tiff_lzw_decode / ips_tiff_lzw_encode_u16), byte-identical to Pillow's."""
import argparse
import logging

from . import storage, tiffio

logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(levelname)s - %(message)s')
logger = logging.getLogger(__name__)


def process_images_in_memory(images, target_size=(1080, 1080)):
    """Batch form of process_image_in_memory: list of encoded 16-bit images of equal shape ->
    list of LZW TIFF byte strings.  Strips are decoded, resized and LZW-encoded on the GPU;
    the bytes equal Pillow's for the same pixels."""
    from .. import ops
    out_w, out_h = int(target_size[0]), int(target_size[1])
    planes = tiffio.load_planes(list(images))
    return tiffio.encode_lzw_from_device(ops.lanczos_resize_u16(planes, (out_h, out_w)))


def process_image_in_memory(image_bytes, target_size=(1080, 1080)):
    """bytes of a 16-bit image -> LZW TIFF bytes of the image resized to target_size
    (width, height), as Image_re-binning.py:12-22.  Raises on undecodable input."""
    return process_images_in_memory([image_bytes], target_size)[0]


BATCH = 16      # images per GPU round trip (strips of a batch are coded concurrently)


def process_images_in_s3(bucket_name, image_folder, resolution, s3_resource=None):
    """Image_re-binning.py:24-64: every image under the prefix -> same key with 'Image'
    replaced by 'Image_binned'; per-object failures are logged and skipped."""
    s3 = s3_resource or storage.resource()
    bucket = s3.Bucket(bucket_name)
    valid_extensions = ('.png', '.jpg', '.jpeg', '.tif', '.tiff')
    processed_count = 0
    if not image_folder.endswith('/'):
        image_folder += '/'

    def put(key, processed):
        bucket.put_object(Key=key.replace('Image', 'Image_binned'), Body=processed, ContentType='image/tiff')

    def flush(batch):
        done = 0
        try:
            for (key, _), processed in zip(batch, process_images_in_memory([d for _, d in batch], (resolution, resolution))):
                put(key, processed)
                done += 1
            return done
        except Exception:
            pass                      # mixed shapes or a damaged file: isolate it image by image
        for key, data in batch[done:]:
            try:
                put(key, process_image_in_memory(data, target_size=(resolution, resolution)))
                done += 1
            except Exception:
                logger.error(f"Failed to process '{key}'", exc_info=True)
        return done

    batch = []
    for obj in bucket.objects.filter(Prefix=image_folder):
        if obj.key.endswith('/') or not obj.key.lower().endswith(valid_extensions):
            continue
        logger.info(f"Processing 's3://{bucket_name}/{obj.key}'...")
        try:
            batch.append((obj.key, obj.get()['Body'].read()))
        except Exception:
            logger.error(f"Failed to process '{obj.key}'", exc_info=True)
        if len(batch) >= BATCH:
            processed_count += flush(batch)
            batch = []
    if batch:
        processed_count += flush(batch)
    logger.info(f"Done! Processed {processed_count} images.")
    return processed_count


def build_parser():
    parser = argparse.ArgumentParser(description="Process and re-bin images from an S3 folder.")
    parser.add_argument("--bucket_name", type=str, required=True, help="S3 bucket containing the files.")
    parser.add_argument("--image_folder", type=str, required=True, help="Source folder path in S3 (e.g., 'path/to/experiment/Image/').")
    parser.add_argument("--resolution", type=int, default=1080, required=False, help="Target resolution for the square image (e.g., 1080).")
    return parser


if __name__ == '__main__':
    a = build_parser().parse_args()
    process_images_in_s3(bucket_name=a.bucket_name, image_folder=a.image_folder, resolution=a.resolution)
