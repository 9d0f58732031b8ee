"""Drop-in for Image_re-binning.py: same functions and flags; the LANCZOS resize
(Image_re-binning.py:18) runs in ips_lanczos_resize_u16, bit-exact with Pillow's I;16 path."""
import argparse
import logging

import numpy as np

from . import storage, tiffio

logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(levelname)s - %(message)s')
logger = logging.getLogger(__name__)


def process_image_in_memory(image_bytes, target_size=(1080, 1080)):
    """bytes of a 16-bit image -> LZW TIFF bytes of the image resized to target_size
    (width, height), as Image_re-binning.py:12-22.  Raises on undecodable input."""
    import torch
    from .. import ops
    img = tiffio.decode(image_bytes)
    if img.dtype != np.uint16 or img.ndim != 2:
        raise ValueError("process_image_in_memory handles single-plane 16-bit images (got %s %s)" % (img.dtype, img.shape))
    out_w, out_h = int(target_size[0]), int(target_size[1])
    dev = torch.from_numpy(np.ascontiguousarray(img[None])).cuda()
    out = ops.lanczos_resize_u16(dev, (out_h, out_w))[0].cpu().numpy()
    return tiffio.encode(out, compression='tiff_lzw')


def process_images_in_s3(bucket_name, image_folder, resolution, s3_resource=None):
    """Image_re-binning.py:24-64: every image under the prefix -> same key with 'Image'
    replaced by 'Image_binned'; per-object failures are logged and skipped."""
    s3 = s3_resource or storage.resource()
    bucket = s3.Bucket(bucket_name)
    valid_extensions = ('.png', '.jpg', '.jpeg', '.tif', '.tiff')
    processed_count = 0
    if not image_folder.endswith('/'):
        image_folder += '/'
    for obj in bucket.objects.filter(Prefix=image_folder):
        if obj.key.endswith('/') or not obj.key.lower().endswith(valid_extensions):
            continue
        logger.info(f"Processing 's3://{bucket_name}/{obj.key}'...")
        try:
            image_data = obj.get()['Body'].read()
            processed = process_image_in_memory(image_data, target_size=(resolution, resolution))
            new_key = obj.key.replace('Image', 'Image_binned')
            bucket.put_object(Key=new_key, Body=processed, ContentType='image/tiff')
            processed_count += 1
        except Exception:
            logger.error(f"Failed to process '{obj.key}'", exc_info=True)
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
