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
    (feature, threshold) because fixed-point sums are exact.  The second half of H1's definition — the
    two S agree — is checked by split_S_error below."""
    identical = trees_identical(gn, on)
    gl = np.sort(gn["count"][gn["feature_idx"] == -1])
    ol = np.sort(on["count"][on["feature_idx"] == -1])
    equivalent = len(gn) == len(on) and np.array_equal(gl, ol) and same_partition(g_node_of_doc, o_node_of_doc)
    return identical, bool(equivalent)


def split_S_error(g_S, o_S):
    """Largest relative difference between the S = sL^2/cL + sR^2/cR values (FeatureHistogram.java:253) of the splits
    of two equivalent trees, in split order.  The device's sums are 2^-39-relative fixed point, the oracle's are
    doubles in sample order: both carry ~1e-11 of rounding, which is what this number measures."""
    g_S, o_S = np.asarray(g_S, np.float64), np.asarray(o_S, np.float64)
    n = min(len(g_S), len(o_S))
    if n == 0:
        return 0.0
    return float(np.max(rel_err(g_S[:n], o_S[:n])))


def first_divergence(gn, on, g_S, o_S):
    """(k, relative S difference) of the first split, in growth order, where two trees disagree, or None.  Split k creates
    nodes 2k+1 / 2k+2, so its parent is the node whose `left` is 2k+1.  A divergence whose S values agree to double
    rounding is a tie between two maximisers of S: exact fixed-point sums resolve it to the lowest
    (feature, threshold), the reference's doubles by the rounding of its summation order; both trees are valid."""
    g_S, o_S = np.asarray(g_S, np.float64), np.asarray(o_S, np.float64)

    def splits(nodes):
        parent = {int(nodes["left"][i]): i for i in range(len(nodes)) if nodes["feature_idx"][i] >= 0}
        out = []
        for k in range((len(nodes) - 1) // 2):
            i = parent[2 * k + 1]
            out.append((i, int(nodes["feature_idx"][i]), int(nodes["threshold_idx"][i])))
        return out

    gs, os_ = splits(gn), splits(on)
    for k in range(min(len(gs), len(os_))):
        if gs[k] != os_[k]:
            return k, float(rel_err(g_S[k:k + 1], o_S[k:k + 1])[0]), gs[k], os_[k]
    return None


class ParityTally:
    """identity / equivalence rates and the largest S disagreement over a run (SURVEY.md H1 asks for both rates)."""

    def __init__(self):
        self.trees = self.identical = self.equivalent = 0
        self.max_S_err = 0.0
        self.ties = None          # the tree at which a tie-broken split ended a stop_at_tie run

    def add(self, identical, equivalent, s_err=0.0):
        self.trees += 1
        self.identical += int(bool(identical))
        self.equivalent += int(bool(equivalent))
        self.max_S_err = max(self.max_S_err, float(s_err))

    def __str__(self):
        return (f"{self.trees} trees: split ids identical {self.identical}/{self.trees}, same partition "
                f"{self.equivalent}/{self.trees}, max |S_gpu - S_oracle| / S = {self.max_S_err:.2e}")


def rel_err(a, b, floor=0.0):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    den = np.maximum(np.maximum(np.abs(a), np.abs(b)), floor)
    den[den == 0] = 1.0
    return np.abs(a - b) / den
