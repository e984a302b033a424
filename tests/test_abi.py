"""CPU tests: the C-ABI library builds for sm_100a, loads, exports every symbol the header
declares, and the product path fails loudly (no CPU fallback) when no device is usable."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'ssdcodec.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(ssdc_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), s
    from jpeg_detection_resnet_ssd_b200 import _lib
    assert sorted(_lib.SIGNATURES) == syms          # the ctypes table binds exactly the header
    assert _lib.load_library().ssdc_version() >= 100


def test_library_is_sm100a_only(built_lib):
    out = subprocess.run(['cuobjdump', '--list-elf', built_lib], capture_output=True, text=True).stdout
    assert 'sm_100a' in out
    assert not re.search(r'sm_(?!100a)\d+', out)


def test_struct_layouts_match_header():
    from jpeg_detection_resnet_ssd_b200 import _lib
    assert ctypes.sizeof(_lib.DecodeParams) == 8 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.EncodeParams) == 8 * 4 + 4 * 8


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must raise, never compute on the host."""
    from jpeg_detection_resnet_ssd_b200 import _lib
    lib = _lib.load_library()
    if lib.ssdc_device_count() > 0:
        pytest.skip('a GPU is visible here')
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections
    from jpeg_detection_resnet_ssd_b200.bounding_box_utils.bounding_box_utils import iou
    y = np.zeros((1, 10, 16), np.float32)
    with pytest.raises(_lib.SSDCodecError) as e:
        decode_detections(y, img_height=10, img_width=10)
    assert e.value.code == _lib.ERR_NODEVICE
    with pytest.raises(_lib.SSDCodecError):
        iou(np.zeros((1, 4)), np.zeros((1, 4)))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'jpeg_detection_resnet_ssd_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'import torch' not in text, f


def test_argument_validation_before_device():
    """Errors the reference raises from pure argument checks keep type and text."""
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    from jpeg_detection_resnet_ssd_b200.keras_layers.keras_layer_DecodeDetections import DecodeDetections
    y = np.zeros((1, 10, 16), np.float32)
    with pytest.raises(ValueError, match='needs the image size'):
        decode_detections(y)
    with pytest.raises(ValueError, match='needs the image size'):
        decode_detections_fast(y, img_height=3)
    with pytest.raises(ValueError, match='Supported input coordinate formats'):
        decode_detections(y, input_coords='xywh', img_height=1, img_width=1)
    with pytest.raises(ValueError, match="only supports the 'centroids'"):
        DecodeDetections(coords='minmax', img_height=1, img_width=1)
    ok = dict(img_height=300, img_width=300, n_classes=20, predictor_sizes=[(4, 4), (2, 2)])
    with pytest.raises(ValueError, match='len\\(scales\\)'):
        SSDInputEncoder(scales=[0.1, 0.2], **ok)
    with pytest.raises(ValueError, match='greater than 0'):
        SSDInputEncoder(scales=[0.1, -0.2, 0.3], **ok)
    with pytest.raises(ValueError, match='min_scale <= max_scale'):
        SSDInputEncoder(min_scale=0.9, max_scale=0.1, **ok)
    with pytest.raises(ValueError, match='aspect_ratios_per_layer'):
        SSDInputEncoder(aspect_ratios_per_layer=[[1.0]], **ok)
    with pytest.raises(ValueError, match='aspect ratios must be greater'):
        SSDInputEncoder(aspect_ratios_global=[1.0, -2.0], **ok)
    with pytest.raises(ValueError, match='4 variance values'):
        SSDInputEncoder(variances=[0.1, 0.2], **ok)
    with pytest.raises(ValueError, match='variances must be >0'):
        SSDInputEncoder(variances=[0.1, 0.1, 0.0, 0.2], **ok)
    with pytest.raises(ValueError, match='Unexpected value for `coords`'):
        SSDInputEncoder(coords='polar', **ok)
    with pytest.raises(ValueError, match='one step value'):
        SSDInputEncoder(steps=[8], **ok)
    with pytest.raises(ValueError, match='one offset value'):
        SSDInputEncoder(offsets=[0.5], **ok)


def test_anchor_generation_matches_oracle():
    """Construction-time host configuration: product anchors == oracle anchors, bit for bit."""
    from oracle import ssd_codec_oracle as orc
    import synth
    from jpeg_detection_resnet_ssd_b200.ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder
    for layout, n in (('ssd300', 8732), ('ssd512', 24564), ('tiny', None)):
        for over in (dict(), dict(coords='corners'), dict(coords='minmax'), dict(clip_boxes=True), dict(normalize_coords=False)):
            a = synth.anchors_of(synth.make_encoder(SSDInputEncoder, layout, **over))
            b = synth.anchors_of(synth.make_encoder(orc.SSDInputEncoder, layout, **over))
            assert np.array_equal(a, b)
            if n:
                assert a.shape == (n, 4)
    e = synth.make_encoder(SSDInputEncoder, 'ssd300')
    a = synth.anchors_of(e)
    assert np.allclose(a[0], [0.013333, 0.013333, 0.1, 0.1], atol=1e-6)
    assert np.allclose(a[-1], [0.5, 0.5, 0.622254, 1.244508], atol=1e-6)
    assert e.n_boxes == [4, 6, 6, 6, 4, 4] and e.n_classes == 21


def test_drop_in_import_paths():
    """with the package directory first on sys.path the reference's own
    import statements resolve to the replacements (training_dct_pascal_j2d_resnet.py:71,
    average_precision_evaluator.py:32, inference.py:21)."""
    import sys
    code = (
        "from ssd_encoder_decoder.ssd_input_encoder import SSDInputEncoder, DegenerateBoxError\n"
        "from ssd_encoder_decoder.ssd_output_decoder import decode_detections, decode_detections_fast, greedy_nms\n"
        "from ssd_encoder_decoder.ssd_output_decoder_no_log import decode_detections as d2\n"
        "from ssd_encoder_decoder.ssd_input_encoder_no_log import SSDInputEncoder as E2\n"
        "from ssd_encoder_decoder.matching_utils import match_bipartite_greedy, match_multi\n"
        "from bounding_box_utils.bounding_box_utils import iou, convert_coordinates, intersection_area\n"
        "from keras_layers.keras_layer_DecodeDetections import DecodeDetections\n"
        "from keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast\n"
        "e = SSDInputEncoder(300, 300, 20, [(38, 38), (19, 19)], scales=[0.1, 0.2, 0.3])\n"
        "assert isinstance(e, SSDInputEncoder) and issubclass(E2, SSDInputEncoder)\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.path.join(ROOT, 'jpeg_detection_resnet_ssd_b200'))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp', env=env)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]


def test_drop_in_keeps_unreplaced_reference_modules(tmp_path):
    """The reference imports modules next to the replaced ones (`keras_layers.keras_layer_AnchorBoxes`,
    `eval_utils.coco_utils`, ...); with the replacement directory first on sys.path they must still resolve to
    the reference's own packages, while the replaced modules resolve to this repo."""
    import sys
    fake = tmp_path / 'localisation_part'
    for pkg, mod in (('keras_layers', 'keras_layer_AnchorBoxes'), ('eval_utils', 'coco_utils'),
                     ('eval_utils', 'average_precision_evaluator'), ('keras_layers', 'keras_layer_DecodeDetections'),
                     ('data_generator', 'object_detection_2d_geometric_ops'), ('data_generator', 'object_detection_2d_misc_utils'),
                     ('data_generator', 'object_detection_2d_image_boxes_validation_utils')):
        (fake / pkg).mkdir(parents=True, exist_ok=True)
        (fake / pkg / '__init__.py').write_text('')
        (fake / pkg / (mod + '.py')).write_text("ORIGIN = 'reference'\n")
    code = (
        "from keras_layers.keras_layer_AnchorBoxes import ORIGIN as a\n"
        "from eval_utils.coco_utils import ORIGIN as b\n"
        "from eval_utils.average_precision_evaluator import Evaluator\n"
        "from keras_layers.keras_layer_DecodeDetections import DecodeDetections\n"
        "import eval_utils.average_precision_evaluator as m\n"
        "from data_generator.object_detection_2d_geometric_ops import ORIGIN as c\n"
        "from data_generator.object_detection_2d_misc_utils import apply_inverse_transforms\n"
        "from data_generator.object_detection_2d_image_boxes_validation_utils import BoxFilter, ImageValidator, BoundGenerator\n"
        "import data_generator.object_detection_2d_misc_utils as m2\n"
        "assert a == b == c == 'reference' and not hasattr(m, 'ORIGIN') and not hasattr(m2, 'ORIGIN')\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'jpeg_detection_resnet_ssd_b200'), str(fake)]))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp', env=env)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]


def test_drop_in_extends_partially_replaced_reference_classes(tmp_path):
    """`eval_utils.average_precision_evaluator` and `keras_loss_function.keras_ssd_loss` replace only a part of the
    reference modules: with the reference importable, `Evaluator` must be the reference's class with the device
    matching core, and `SSDLoss` the reference's own (TensorFlow) loss."""
    import sys
    fake = tmp_path / 'localisation_part'
    (fake / 'eval_utils').mkdir(parents=True)
    (fake / 'eval_utils' / '__init__.py').write_text('')
    (fake / 'eval_utils' / 'average_precision_evaluator.py').write_text(
        "class Evaluator(object):\n"
        "    def __init__(self, *a, **k):\n        self.origin = 'reference'\n"
        "    def predict_on_dataset(self):\n        return 'reference predict'\n"
        "    def match_predictions(self):\n        return 'reference match'\n")
    (fake / 'keras_loss_function').mkdir()
    (fake / 'keras_loss_function' / '__init__.py').write_text('')
    (fake / 'keras_loss_function' / 'keras_ssd_loss.py').write_text("class SSDLoss:\n    origin = 'reference'\n")
    (fake / 'keras_layers').mkdir()
    (fake / 'keras_layers' / '__init__.py').write_text('')
    for name in ('DecodeDetections', 'DecodeDetectionsFast'):
        (fake / 'keras_layers' / ('keras_layer_%s.py' % name)).write_text("class %s:\n    origin = 'reference'\n" % name)
    code = (
        "from eval_utils.average_precision_evaluator import Evaluator, DeviceEvaluator\n"
        "e = Evaluator()\n"
        "assert e.origin == 'reference' and e.predict_on_dataset() == 'reference predict'\n"
        "assert Evaluator.match_predictions is DeviceEvaluator.match_predictions\n"
        "from keras_loss_function.keras_ssd_loss import SSDLoss, DeviceSSDLoss\n"
        "assert SSDLoss.origin == 'reference' and hasattr(DeviceSSDLoss, 'compute_loss')\n"
        # the models build their inference graph with these classes (keras_ssd300_dct_j2d_resnet.py:42-43,885): the
        # reference's Keras layers must come through, the device callables keep their own names
        "from keras_layers.keras_layer_DecodeDetections import DecodeDetections, DeviceDecodeDetections\n"
        "from keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast, DeviceDecodeDetectionsFast\n"
        "assert DecodeDetections.origin == 'reference' and DecodeDetectionsFast.origin == 'reference'\n"
        "assert hasattr(DeviceDecodeDetections, 'call') and issubclass(DeviceDecodeDetectionsFast, DeviceDecodeDetections)\n"
        "print('ok')\n")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, 'jpeg_detection_resnet_ssd_b200'), str(fake)]))
    r = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, cwd='/tmp', env=env)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]
    # stand-alone (no reference on the path): the device classes
    code2 = ("from jpeg_detection_resnet_ssd_b200.eval_utils.average_precision_evaluator import Evaluator, DeviceEvaluator\n"
             "from jpeg_detection_resnet_ssd_b200.keras_loss_function.keras_ssd_loss import SSDLoss, DeviceSSDLoss\n"
             "from jpeg_detection_resnet_ssd_b200.keras_layers.keras_layer_DecodeDetections import DecodeDetections, DeviceDecodeDetections\n"
             "from jpeg_detection_resnet_ssd_b200.keras_layers.keras_layer_DecodeDetectionsFast import DecodeDetectionsFast, DeviceDecodeDetectionsFast\n"
             "assert DecodeDetections is DeviceDecodeDetections and DecodeDetectionsFast is DeviceDecodeDetectionsFast\n"
             "assert Evaluator is DeviceEvaluator and SSDLoss is DeviceSSDLoss\nprint('ok')\n")
    r = subprocess.run([sys.executable, '-c', code2], capture_output=True, text=True, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == 'ok', r.stderr[-1500:]


def test_build_stamp_survives_a_copy_of_the_tree(built_lib, tmp_path):
    """The GPU box runs a snapshot of the repository at another path: the prebuilt library must count as current there
    (the stamp hashes file names and contents, not absolute paths), and a changed source must not."""
    import shutil
    import sys
    src = os.path.join(ROOT, 'jpeg_detection_resnet_ssd_b200')
    dst = tmp_path / 'snap'
    shutil.copytree(src, dst / 'jpeg_detection_resnet_ssd_b200', ignore=shutil.ignore_patterns('__pycache__', '*.o'))
    shutil.copytree(os.path.join(ROOT, 'include'), dst / 'include')
    code = ('import sys; sys.path.insert(0, %r); from jpeg_detection_resnet_ssd_b200 import build; '
            'print(build.LIBPATH.startswith(%r), build.is_current())' % (str(dst), str(dst)))
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout.split()
    assert out == ['True', 'True']
    with open(dst / 'jpeg_detection_resnet_ssd_b200' / 'csrc' / 'thin.cu', 'a') as fh:
        fh.write('\n// changed\n')
    out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, check=True).stdout.split()
    assert out == ['True', 'False']
