"""CPU oracle of the SSD multibox loss, forward pass.  TEST INFRASTRUCTURE ONLY.

numpy restatement of `SSDLoss.compute_loss` of
/root/reference/localisation_part/keras_loss_function/keras_ssd_loss.py:98-211 (log_loss :78-96, smooth_L1_loss :53-76).

PARITY UNPINNED: the reference evaluates this inside TensorFlow (tensorflow-gpu 1.8.0 / 1.14.0 per Pipfile /
Pipfile.lock), which cannot be installed here, and the repository holds no golden values for it.  The restatement
follows the graph op by op in float32 (the dtype of the Keras placeholders) with TensorFlow's documented semantics:
`tf.nn.top_k` returns the lower index first among equal values; `tf.count_nonzero`; `tf.to_int32` truncates.
"""
from __future__ import division

import numpy as np


def smooth_l1_loss(y_true, y_pred):
    """:53-76"""
    absolute_loss = np.abs(y_true - y_pred)
    square_loss = np.float32(0.5) * (y_true - y_pred) ** 2
    l1 = np.where(absolute_loss < np.float32(1.0), square_loss, absolute_loss - np.float32(0.5))
    return np.sum(l1, axis=-1, dtype=np.float32)


def log_loss(y_true, y_pred):
    """:78-96"""
    y_pred = np.maximum(y_pred, np.float32(1e-15))
    return -np.sum(y_true * np.log(y_pred), axis=-1, dtype=np.float32)


def compute_loss(y_true, y_pred, neg_pos_ratio=3, n_neg_min=0, alpha=1.0):
    """:98-211.  Returns (batch_size,) float32."""
    y_true = np.asarray(y_true, dtype=np.float32)
    y_pred = np.asarray(y_pred, dtype=np.float32)
    batch_size, n_boxes = y_pred.shape[0], y_pred.shape[1]
    with np.errstate(all='ignore'):
        classification_loss = log_loss(y_true[:, :, :-12], y_pred[:, :, :-12])                 # :130
        localization_loss = smooth_l1_loss(y_true[:, :, -12:-8], y_pred[:, :, -12:-8])         # :131
    negatives = y_true[:, :, 0]                                                                # :137
    positives = np.max(y_true[:, :, 1:-12], axis=-1)                                           # :138
    n_positive = np.sum(positives, dtype=np.float32)                                           # :141
    pos_class_loss = np.sum(classification_loss * positives, axis=-1, dtype=np.float32)        # :147
    neg_class_loss_all = classification_loss * negatives                                       # :149
    n_neg_losses = int(np.count_nonzero(neg_class_loss_all))                                   # :150
    n_negative_keep = min(max(int(neg_pos_ratio) * int(n_positive), int(n_neg_min)), n_neg_losses)   # :163
    if n_neg_losses == 0:
        neg_class_loss = np.zeros(batch_size, dtype=np.float32)                                # :167-168
    else:
        flat = neg_class_loss_all.reshape(-1)
        # tf.nn.top_k: the k largest, lower index first among equal values = a stable descending sort
        order = np.argsort(-flat.astype(np.float64), kind='stable')[:n_negative_keep]
        keep = np.zeros(flat.shape[0], dtype=np.float32)
        keep[order] = 1.0
        keep = keep.reshape(batch_size, n_boxes)
        neg_class_loss = np.sum(classification_loss * keep, axis=-1, dtype=np.float32)         # :188
    class_loss = pos_class_loss + neg_class_loss                                               # :193
    loc_loss = np.sum(localization_loss * positives, axis=-1, dtype=np.float32)                # :198
    total_loss = (class_loss + np.float32(alpha) * loc_loss) / np.maximum(np.float32(1.0), n_positive)   # :202
    return (total_loss * np.float32(batch_size)).astype(np.float32)                            # :207
