"""Shared helpers of the parity tests: tree comparison under SURVEY.md H1/F10's equivalence."""
import numpy as np


def trees_identical(a, b):
    if len(a) != len(b):
        return False
    return all(np.array_equal(a[k], b[k]) for k in ("feature_idx", "threshold_idx", "left", "right", "count"))


def same_partition(a, b):
    """True iff the two node-id labelings of the samples define the same set partition."""
    pairs = np.unique(np.stack([np.asarray(a), np.asarray(b)], axis=1), axis=0)
    return len(pairs) == len(np.unique(a)) == len(np.unique(b))


def compare_tree(gn, on, g_node_of_doc, o_node_of_doc):
    """Returns (identical, equivalent).
    identical:  same shape and the same (featureIdx, thresholdIdx) at every split.
    equivalent: the same partition of the training samples into leaves with the same leaf sizes.  This
    is the robust integer criterion of SURVEY.md H1/F10: where two candidate splits induce the same
    (or the mirrored) partition their S values are equal in exact arithmetic, the reference's choice
    among them is decided by double rounding noise of its summation order, and ours by the lowest
    (feature, threshold) because fixed-point sums are exact."""
    identical = trees_identical(gn, on)
    gl = np.sort(gn["count"][gn["feature_idx"] == -1])
    ol = np.sort(on["count"][on["feature_idx"] == -1])
    equivalent = len(gn) == len(on) and np.array_equal(gl, ol) and same_partition(g_node_of_doc, o_node_of_doc)
    return identical, bool(equivalent)


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    den[den == 0] = 1.0
    return np.abs(a - b) / den
