#!/usr/bin/env python
"""Illumination_QC_mult.py -- same name, same flags, same outputs as the reference's Illumination_QC_mult.py
(Saguaro-Biosciences/image-processing-suite); the arithmetic runs in libips.so on the GPU.

    python scripts/Illumination_QC_mult.py --load-data L.csv --data-path DIR --channels A B [--illum-path DIR] [--output O.csv] [--threads N]

This file only puts the repository on sys.path and runs
``image_processing_suite_b200.scripts.Illumination_QC_mult`` as ``__main__``; S3 is boto3, or the directory
``$IPS_STORAGE_ROOT/<bucket>/<key>`` when that variable is set.
"""
import os
import runpy
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

if __name__ == "__main__":
    runpy.run_module("image_processing_suite_b200.scripts.Illumination_QC_mult", run_name="__main__", alter_sys=True)
