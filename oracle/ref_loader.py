"""Imports the REAL reference codec from /root/reference (read-only, only present in the
build container).  Used by oracle/make_golden.py and tests/test_oracle_vs_reference.py to pin
the oracle; nothing that runs on the GPU box may depend on it.

The reference predates numpy 1.24 and uses the removed aliases `np.float` / `np.int`
(bounding_box_utils.py:60,101; ssd_input_encoder.py:330,349; matching_utils.py:59); they were
exact synonyms of the builtins, so re-creating them does not change semantics.
"""
import importlib
import os
import sys

import numpy as np

REF_ROOT = '/root/reference/localisation_part'


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'ssd_encoder_decoder'))


def load():
    """Returns a namespace with the reference modules."""
    if not available():
        raise RuntimeError('reference not present at ' + REF_ROOT)
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'int'):
        np.int = int
    # the reference's package names collide with the product's; import them under a private
    # sys.path entry and detach afterwards
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils')}
    sys.path.insert(0, REF_ROOT)
    try:
        class NS(object):
            pass
        ns = NS()
        ns.bbox = importlib.import_module('bounding_box_utils.bounding_box_utils')
        ns.matching = importlib.import_module('ssd_encoder_decoder.matching_utils')
        ns.decoder = importlib.import_module('ssd_encoder_decoder.ssd_output_decoder')
        ns.encoder = importlib.import_module('ssd_encoder_decoder.ssd_input_encoder')
        ns.decoder_no_log = importlib.import_module('ssd_encoder_decoder.ssd_output_decoder_no_log')
        ns.encoder_no_log = importlib.import_module('ssd_encoder_decoder.ssd_input_encoder_no_log')
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils'):
                sys.modules.pop(k)
        sys.modules.update(saved)
    return ns


def load_evaluator():
    """The reference's eval_utils.average_precision_evaluator module.  Its imports pull in the data
    generators, which need third-party packages that are absent here (bs4, h5py, keras, jpeg2dct ...);
    none of them is touched by the evaluation core, so they are replaced by empty stub modules."""
    import types
    if not available():
        raise RuntimeError('reference not present at ' + REF_ROOT)
    if not hasattr(np, 'float'):
        np.float = float
    if not hasattr(np, 'int'):
        np.int = int
    stubs = ['bs4', 'h5py', 'jpeg2dct', 'jpeg2dct.numpy', 'jpegdecoder', 'keras', 'keras.preprocessing',
             'keras.preprocessing.image', 'keras.utils', 'tensorflow']
    added = []
    for name in stubs:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
            added.append(name)
    sys.modules['bs4'].BeautifulSoup = getattr(sys.modules['bs4'], 'BeautifulSoup', object)
    sys.modules['keras.utils'].Sequence = getattr(sys.modules['keras.utils'], 'Sequence', object)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils', 'eval_utils', 'data_generator')}
    sys.path.insert(0, REF_ROOT)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            mod = importlib.import_module('eval_utils.average_precision_evaluator')
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils', 'eval_utils', 'data_generator'):
                sys.modules.pop(k)
        sys.modules.update(saved)
        for name in added:
            sys.modules.pop(name, None)
    return mod


def load_data_generator_utils():
    """The reference's geometric / patch-sampling operations, box validation utilities and misc utils.  `cv2` is absent
    here and only touches pixels: a stub with the interpolation constants and a `resize` that returns an array of the
    requested size is enough for the label arithmetic and the inverters.  `np.bool` (removed from numpy) is re-created
    as the builtin it aliased (image_boxes_validation_utils.py:183)."""
    import types
    if not available():
        raise RuntimeError('reference not present at ' + REF_ROOT)
    for name, val in (('float', float), ('int', int), ('bool', bool)):
        if not hasattr(np, name):
            setattr(np, name, val)
    cv2 = types.ModuleType('cv2')
    for i, name in enumerate(['INTER_NEAREST', 'INTER_LINEAR', 'INTER_CUBIC', 'INTER_AREA', 'INTER_LANCZOS4']):
        setattr(cv2, name, i)
    cv2.resize = lambda image, dsize, interpolation=1: np.zeros((dsize[1], dsize[0]) + tuple(image.shape[2:]), dtype=image.dtype)
    had_cv2 = sys.modules.get('cv2')
    sys.modules['cv2'] = cv2
    saved = {k: sys.modules.pop(k) for k in list(sys.modules)
             if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils', 'data_generator')}
    sys.path.insert(0, REF_ROOT)
    try:
        class NS(object):
            pass
        ns = NS()
        ns.misc = importlib.import_module('data_generator.object_detection_2d_misc_utils')
        ns.validation = importlib.import_module('data_generator.object_detection_2d_image_boxes_validation_utils')
        ns.geometric = importlib.import_module('data_generator.object_detection_2d_geometric_ops')
        ns.patch = importlib.import_module('data_generator.object_detection_2d_patch_sampling_ops')
    finally:
        sys.path.remove(REF_ROOT)
        for k in list(sys.modules):
            if k.split('.')[0] in ('ssd_encoder_decoder', 'bounding_box_utils', 'data_generator'):
                sys.modules.pop(k)
        sys.modules.update(saved)
        if had_cv2 is None:
            sys.modules.pop('cv2', None)
        else:
            sys.modules['cv2'] = had_cv2
    return ns


def reference_inverters(ns, specs):
    """Inverter closures of the REAL transformation classes for a list of specifications (oracle/cases.py)."""
    out = []
    for spec in specs:
        if spec is None:
            out.append(None)
        elif spec[0] == 'resize':
            _, H, W, oh, ow = spec
            _, inv = ns.geometric.Resize(height=oh, width=ow)(np.zeros((H, W, 3), np.uint8), return_inverter=True)
            out.append(inv)
        elif spec[0] == 'translate':
            _, dy, dx = spec
            # (CropPad copies `labels` before testing it for None, so it must be given some)
            res = ns.patch.CropPad(patch_ymin=dy, patch_xmin=dx, patch_height=400, patch_width=400)(
                np.zeros((300, 300, 3), np.uint8), np.zeros((1, 5)), return_inverter=True)
            out.append(res[-1])
        else:
            out.append(lambda labels: labels)
    return out
