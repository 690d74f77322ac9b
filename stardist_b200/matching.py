"""relabel_sequential (stardist/matching.py:319-406; itself taken from scikit-image) -- host-side
integer bookkeeping used by StarDist3D._instances_from_prediction (model3d.py:645) and
predict_instances_big (base.py:959)."""
import numpy as np


def relabel_sequential(label_field, offset=1):
    """Relabel arbitrary labels to {`offset`, ... `offset` + number_of_labels}.
    Returns (relabeled, forward_map, inverse_map); label 0 is background and never remapped."""
    offset = int(offset)
    if offset <= 0:
        raise ValueError("Offset must be strictly positive.")
    if np.min(label_field) < 0:
        raise ValueError("Cannot relabel array that contains negative values.")
    max_label = int(label_field.max())
    if not np.issubdtype(label_field.dtype, np.integer):
        new_type = np.min_scalar_type(max_label)
        label_field = label_field.astype(new_type)
    labels = np.unique(label_field)
    labels0 = labels[labels != 0]
    new_max_label = offset - 1 + len(labels0)
    new_labels0 = np.arange(offset, new_max_label + 1)
    output_type = label_field.dtype
    required_type = np.min_scalar_type(new_max_label)
    if np.dtype(required_type).itemsize > np.dtype(label_field.dtype).itemsize:
        output_type = required_type
    forward_map = np.zeros(max_label + 1, dtype=output_type)
    forward_map[labels0] = new_labels0
    inverse_map = np.zeros(new_max_label + 1, dtype=output_type)
    inverse_map[offset:] = labels0
    relabeled = forward_map[label_field]
    return relabeled, forward_map, inverse_map


def relabel_sequential_device(label_field, offset=1, max_label=None):
    """relabel_sequential (stardist/matching.py:319-406) on a device int32 label map, IN PLACE (sdb_relabel_sequential).
    Returns (label_field, forward_map[max_label + 1] device int32, n_labels).  `max_label`: an upper bound of the values
    (default: computed on the device).  No CPU fallback: raises without the CUDA library."""
    import ctypes, torch
    from . import _lib as L
    lib = L.require_cuda()
    if not (isinstance(label_field, torch.Tensor) and label_field.is_cuda and label_field.dtype == torch.int32 and label_field.is_contiguous()):
        raise ValueError("relabel_sequential_device expects a contiguous CUDA int32 tensor")
    if int(offset) <= 0:
        raise ValueError("Offset must be strictly positive.")
    if max_label is None:
        max_label = int(label_field.max()) if label_field.numel() else 0
    fwd = torch.empty(int(max_label) + 3, dtype=torch.int32, device=label_field.device)
    cnt = ctypes.c_int(0)
    rc = lib.sdb_relabel_sequential(L.ptr(label_field), label_field.numel(), int(max_label), int(offset), L.ptr(fwd), ctypes.byref(cnt), L.stream_ptr())
    if rc != 0:
        msg = L.load().sdb_last_error().decode('utf-8', 'replace')
        if 'negative' in msg:
            raise ValueError("Cannot relabel array that contains negative values.")
        L.check(rc)
    return label_field, fwd[:int(max_label) + 1], int(cnt.value)


# ---------------------------------------------------------------------------------------------------------
# Detection metrics (stardist/matching.py:109-232).  Host-side evaluation helper: pairs ground-truth and predicted
# objects by an optimal assignment on the IoU (or IoT / IoP) matrix and counts tp / fp / fn at a threshold.
def _overlap_matrix(a, b):
    """counts[i, j] = number of pixels with label i in a and j in b (labels already sequential, 0 = background)"""
    na, nb = int(a.max()) + 1, int(b.max()) + 1
    flat = a.ravel().astype(np.int64) * nb + b.ravel().astype(np.int64)
    return np.bincount(flat, minlength=na * nb).reshape(na, nb)


def _ratio(num, den):
    out = np.zeros(np.broadcast(num, den).shape, np.float32)
    np.divide(num, den, out=out, where=np.abs(den) > 1e-10)
    return out


def _criterion_scores(overlap, criterion):
    if overlap.sum() == 0:
        return overlap.astype(np.float32)
    n_pred = overlap.sum(axis=0, keepdims=True)
    n_true = overlap.sum(axis=1, keepdims=True)
    if criterion == 'iou':
        return _ratio(overlap, n_pred + n_true - overlap)
    if criterion == 'iot':
        return _ratio(overlap, n_true)
    if criterion == 'iop':
        return _ratio(overlap, n_pred)
    raise ValueError("Matching criterion '%s' not supported." % criterion)


def matching(y_true, y_pred, thresh=0.5, criterion='iou'):
    """Matching(criterion, thresh, fp, tp, fn, precision, recall, accuracy, f1, n_true, n_pred, mean_true_score,
    mean_matched_score, panoptic_quality) between two label images (stardist/matching.py:109-232)."""
    from collections import namedtuple
    from scipy.optimize import linear_sum_assignment
    for name, y in (('y_true', y_true), ('y_pred', y_pred)):
        if not (isinstance(y, np.ndarray) and np.issubdtype(y.dtype, np.integer) and (y.size == 0 or y.min() >= 0)):
            raise ValueError("%s must be an array of non-negative integers." % name)
    if y_true.shape != y_pred.shape:
        raise ValueError("y_true and y_pred have different shapes")
    thr = 0.0 if thresh is None else float(thresh)
    yt = relabel_sequential(y_true)[0]
    yp = relabel_sequential(y_pred)[0]
    scores = _criterion_scores(_overlap_matrix(yt, yp), criterion)[1:, 1:]
    n_true, n_pred = scores.shape
    n_matched = min(n_true, n_pred)
    tp, sum_matched = 0, 0.0
    if n_matched > 0:
        # maximise the number of pairs above the threshold, ties broken by the summed score (:185-190)
        costs = -(scores >= thr).astype(float) - scores / (2 * n_matched)
        ti, pi = linear_sum_assignment(costs)
        ok = scores[ti, pi] >= thr
        tp = np.count_nonzero(ok)                      # numpy integer scalar, as in the reference: float32 / np.int64 -> float64 below
        sum_matched = np.sum(scores[ti, pi][ok])       # stays float32 like the reference's (matching.py:190): the means below are float32 quotients
    fp, fn = n_pred - tp, n_true - tp
    div = lambda a, b: (a / b) if np.abs(b) > 1e-10 else 0.0       # _safe_divide on scalars (matching.py:55-58)
    # scalar types follow the reference's expressions (matching.py:180-196) so that the float results are the same numbers:
    # sum (float32) / n_true (int) stays float32, / tp (numpy int) and / (tp + fp/2 + fn/2) are float64 quotients
    fields = dict(criterion=criterion, thresh=thr, fp=fp, tp=tp, fn=fn,
                  precision=(tp / (tp + fp) if tp > 0 else 0), recall=(tp / (tp + fn) if tp > 0 else 0),
                  accuracy=(tp / (tp + fp + fn) if tp > 0 else 0), f1=((2 * tp) / (2 * tp + fp + fn) if tp > 0 else 0),
                  n_true=n_true, n_pred=n_pred, mean_true_score=div(sum_matched, n_true),
                  mean_matched_score=div(sum_matched, tp), panoptic_quality=div(sum_matched, tp + fp / 2 + fn / 2))
    return namedtuple('Matching', fields.keys())(*fields.values())
