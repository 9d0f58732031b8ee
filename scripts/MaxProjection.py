#!/usr/bin/env python
"""MaxProjection.py -- same name, same flags, same outputs as the reference's MaxProjection.py
(Saguaro-Biosciences/image-processing-suite); the arithmetic runs in libips.so on the GPU.

    python scripts/MaxProjection.py --bucket_data_set B --data_set K.csv --channels C --planes Z --bucket_images B2

This file only puts the repository on sys.path and runs
``image_processing_suite_b200.scripts.MaxProjection`` as ``__main__``; S3 is boto3, or the directory
``$IPS_STORAGE_ROOT/<bucket>/<key>`` when that variable is set.
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    runpy.run_module("image_processing_suite_b200.scripts.MaxProjection", run_name="__main__", alter_sys=True)
