"""Drop-in entry points: the reference's script names, functions, CLI flags and file
schemas, with the per-image arithmetic routed through libips.so (CUDA, no CPU fallback).

    python -m image_processing_suite_b200.scripts.MaxProjection --bucket_data_set ... (MaxProjection.py)
    python -m image_processing_suite_b200.scripts.Illumination_QC_mult --load-data ...  (Illumination_QC_mult.py)
    python -m image_processing_suite_b200.scripts.Image_rebinning --bucket_name ...     (Image_re-binning.py)
    python -m image_processing_suite_b200.scripts.Feature_extraction -c -r -p x -o out --data-file load.csv
                                  (the CellProfiler command line of Feature_extraction_opt.py:166-167)
    python -m image_processing_suite_b200.scripts.Normalize_CP_ami --bucket_name ...    (Normalize_CP_ami.py)
    python -m image_processing_suite_b200.scripts.Feature_select_cosine_ami ...         (Feature_select_cosine_ami.py)
    python -m image_processing_suite_b200.scripts.Pycyto_pertime --bucket_name ...      (Pycyto_pertime.py)
    python -m image_processing_suite_b200.scripts.Illumination_estimate ...             (new: writes {ch}_illum.npy)

Storage: the reference talks to S3 through boto3.  ``storage.client()`` returns a boto3
client when boto3 is importable and IPS_STORAGE_ROOT is unset; otherwise an object with the
same few methods over the local directory ``$IPS_STORAGE_ROOT/<bucket>/<key>``.
"""
