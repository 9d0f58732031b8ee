#!/usr/bin/env python
"""Feature_select_cosine_ami.py -- same name, same flags, same outputs as the reference's Feature_select_cosine_ami.py
(Saguaro-Biosciences/image-processing-suite); the arithmetic runs in libips.so on the GPU.

    python scripts/Feature_select_cosine_ami.py --bucket_name B --base_folder F --plates P.. --exp E --output_bucket B2 --output_prefix X [--na_cutoff .5] [--corr_3hold .9] [--per_time]

This file only puts the repository on sys.path and runs
``image_processing_suite_b200.scripts.Feature_select_cosine_ami`` as ``__main__``; S3 is boto3, or the directory
``$IPS_STORAGE_ROOT/<bucket>/<key>`` when that variable is set.
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    runpy.run_module("image_processing_suite_b200.scripts.Feature_select_cosine_ami", run_name="__main__", alter_sys=True)
