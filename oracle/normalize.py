"""Oracle: well-level aggregation and robust-z normalisation (consumer of the
all-gather).  mad_robustize is PARITY UNPINNED.

Test infrastructure only -- see oracle/__init__.py.
"""
import numpy as np
import pandas as pd

MAD_SCALE = 1.4826
MAD_EPS = 1e-18


def well_mean(rows, well_ids):
    """Per-well mean of per-object rows: ``groupby('Metadata_Well').agg('mean')``
    (Normalize_CP_ami.py:126, Pycyto_pertime.py:69-72).  Returns (sorted unique well
    ids, means[n_wells][n_features]) -- pandas sorts group keys."""
    df = pd.DataFrame(np.asarray(rows, dtype=np.float64))
    df["Metadata_Well"] = np.asarray(well_ids)
    g = df.groupby("Metadata_Well", as_index=False).agg("mean")
    return g["Metadata_Well"].to_numpy(), g.drop(columns=["Metadata_Well"]).to_numpy()


def mad_robustize(profiles, is_control):
    """(x - median_ctrl) / (1.4826 * MAD_ctrl + 1e-18), per feature column.

    pycytominer ``normalize(method="mad_robustize")`` (call sites
    Normalize_CP_ami.py:137-142, Pycyto_pertime.py:84-89) fits a RobustMAD transform on
    the rows selected by ``samples`` (the DMSO wells of the timepoint) and applies it to
    every row.  pycytominer is not installed and not pinned by requirements.txt:4, so
    this follows its published RobustMAD definition (median / scaled MAD with epsilon).
    """
    x = np.asarray(profiles, dtype=np.float64)
    ctrl = x[np.asarray(is_control, dtype=bool)]
    med = np.nanmedian(ctrl, axis=0)
    mad = np.nanmedian(np.abs(ctrl - med), axis=0) * MAD_SCALE
    return (x - med) / (mad + MAD_EPS)


def double_sigmoid(x, k=3, alpha=2.3538):
    """Feature_select_cosine_ami.py:22-27 / Pycyto_pertime.py:13-16."""
    x = np.asarray(x, dtype=np.float64)
    r = (x / alpha) ** k
    return r / np.sqrt(1.0 + (x / alpha) ** (2 * k))
