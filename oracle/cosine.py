"""Oracle: replicate-group cosine similarity (A8).

Test infrastructure only -- see oracle/__init__.py.
Feature_select_cosine_ami.py:145-149 and Pycyto_pertime.py:132-140 call
``sklearn.metrics.pairwise.cosine_similarity`` on a (replicates x features) float64
block (NaN already replaced by 0), keep the strict upper triangle and average it.
"""
import numpy as np


def normalise_rows(x):
    """Row L2 normalisation the way sklearn does it: zero rows stay zero."""
    x = np.asarray(x, dtype=np.float64)
    nrm = np.sqrt((x * x).sum(axis=1))
    nrm[nrm == 0.0] = 1.0
    return x / nrm[:, None]


def cosine_matrix(x):
    xh = normalise_rows(x)
    return xh @ xh.T


def triu_values(x):
    """Strict-upper-triangle similarities in row-major order (Pycyto_pertime.py:155)."""
    s = cosine_matrix(x)
    iu = np.triu_indices(s.shape[0], k=1)
    return s[iu]


def triu_mean(x):
    """Mean strict-upper-triangle similarity; NaN when there is no pair (:149)."""
    v = triu_values(x)
    return float(v.mean()) if v.size else float("nan")


def triu_sum_closed_form(x):
    """Independent check: sum_{i<j} x^_i . x^_j = (|sum x^_i|^2 - sum |x^_i|^2) / 2."""
    xh = normalise_rows(x)
    s = xh.sum(axis=0)
    return 0.5 * (float(s @ s) - float((xh * xh).sum()))


def grouped_triu(x, group):
    """Per-group (sum of strict-upper-triangle similarities, number of pairs).

    Mirrors the per-(compound, timepoint, concentration) loop at
    Feature_select_cosine_ami.py:131-156; groups are returned in order of first
    appearance of their id.
    """
    x = np.asarray(x, dtype=np.float64)
    group = np.asarray(group)
    out = {}
    for g in dict.fromkeys(group.tolist()):
        rows = x[group == g]
        v = triu_values(rows)
        out[g] = (float(v.sum()), int(v.size))
    return out
