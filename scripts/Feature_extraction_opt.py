#!/usr/bin/env python
"""Feature_extraction_opt.py -- same name and entry point (``run_batch_processing()``, no flags,
module-level job constants) as the reference's launcher (Feature_extraction_opt.py:45-67, :73-145).

The reference starts EC2 instances and runs, per (plate, time) job,

    docker run cellprofiler:4.2.8 -c -r -p <pipe> -o /mnt/output --data-file load_data_{plate}_{time}_illum.csv

(:166-167), then syncs ``Image / Nuclei / Cells / Cytoplasm.csv`` to ``{S3_BASE_OUTPUT_PATH}/{plate}/{time}``
(:171-178).  Here every job is the same command line on this machine's GPU
(``image_processing_suite_b200.scripts.Feature_extraction``, ips_field_fused), reading
``{INPUT_BASE}/load_data_{plate}_{time}_illum.csv`` and writing ``{OUTPUT_BASE}/{plate}/{time}/``.
Cloud orchestration (EC2, SSM, docker) is out of scope.

The constants keep the reference's names; every one can be overridden from the environment
(``IPS_FOLDER``, ``IPS_PLATES_TO_RUN="P01,P02"``, ``IPS_TIMES_TO_RUN="6,12"``, ``IPS_INPUT_BASE``, ``IPS_OUTPUT_BASE``).
"""
import logging
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def _env_list(name, default):
    v = os.environ.get(name)
    return [x for x in v.split(",") if x] if v else default


FOLDER = os.environ.get("IPS_FOLDER", "Subset1_10uM_Run03")
CCPIPE_NAME = os.environ.get("IPS_CCPIPE_NAME", "Feature_Extraction_CL2.0.cppipe")      # accepted, not interpreted
PLATES_TO_RUN = _env_list("IPS_PLATES_TO_RUN", ['P01', 'P02', 'P03', 'P04', 'P05', 'P06'])
TIMES_TO_RUN = _env_list("IPS_TIMES_TO_RUN", ['6', '12', '24', '48'])
BATCH_SIZE = int(os.environ.get("IPS_BATCH_SIZE", "10"))                                # jobs per log group, as in the reference
_ROOT = os.environ.get("IPS_STORAGE_ROOT", ".")
INPUT_BASE = os.environ.get("IPS_INPUT_BASE", os.path.join(_ROOT, "cellprofiler-resuts", "IRIC", FOLDER))
OUTPUT_BASE = os.environ.get("IPS_OUTPUT_BASE", INPUT_BASE)                             # S3_BASE_OUTPUT_PATH


def job_command(plate, time):
    """The CellProfiler command line of one job (Feature_extraction_opt.py:166-167) as an argv list."""
    return ["-c", "-r", "-p", os.path.join(INPUT_BASE, CCPIPE_NAME), "-o", os.path.join(OUTPUT_BASE, plate, str(time)),
            "--data-file", os.path.join(INPUT_BASE, f"load_data_{plate}_{time}_illum.csv")]


def run_batch_processing():
    from image_processing_suite_b200.scripts import Feature_extraction
    all_jobs = [(plate, time) for plate in PLATES_TO_RUN for time in TIMES_TO_RUN]
    batches = [all_jobs[i:i + BATCH_SIZE] for i in range(0, len(all_jobs), BATCH_SIZE)]
    done = []
    for n, batch in enumerate(batches, 1):
        logging.info("--- Starting Batch %d/%d (%d jobs) ---", n, len(batches), len(batch))
        for plate, time in batch:
            argv = job_command(plate, time)
            if not os.path.exists(argv[-1]):
                logging.error("[JOB: %s_%sh] missing %s", plate, time, argv[-1])
                continue
            Feature_extraction.main(argv)
            done.append((plate, time))
    logging.info("--- All batches have been processed. ---")
    return done


if __name__ == "__main__":
    logging.basicConfig(level=logging.INFO, format='%(asctime)s - %(message)s')
    run_batch_processing()
