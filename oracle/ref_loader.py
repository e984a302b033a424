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
