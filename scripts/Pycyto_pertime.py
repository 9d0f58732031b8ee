#!/usr/bin/env python
"""Pycyto_pertime.py -- same name, same flags, same outputs as the reference's Pycyto_pertime.py
(Saguaro-Biosciences/image-processing-suite); the arithmetic runs in libips.so on the GPU.

    python scripts/Pycyto_pertime.py --bucket_name B --base_folder F --times T.. --output_bucket B2 --output_prefix X

This file only puts the repository on sys.path and runs
``image_processing_suite_b200.scripts.Pycyto_pertime`` as ``__main__``; S3 is boto3, or the directory
``$IPS_STORAGE_ROOT/<bucket>/<key>`` when that variable is set.
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    runpy.run_module("image_processing_suite_b200.scripts.Pycyto_pertime", run_name="__main__", alter_sys=True)
