"""Execute the reference's own pure-torch files by path (TEST INFRASTRUCTURE).

Used ONLY in the build container, where ``/root/reference`` is mounted, to (a)
generate ``tests/golden/*.npz`` (``tests/golden/make_golden.py``) and (b) let
``tests/test_oracle_vs_reference.py`` assert restatement == verbatim whenever the
mount exists.  Nothing here copies reference source: files are loaded where they
lie, behind ``sys.modules`` stubs for the third-party packages that are not
installed (mmcv, mmengine, prettytable) and for the ``mmseg`` package itself
(whose ``__init__`` asserts on those and whose ``backbones/lednet.py`` is not
Python).  The GPU box has no ``/root/reference``: nothing in the ``-m gpu``
tests, ``smoke()`` or ``bench.py`` touches this module.
"""
import importlib.util
import os
import sys
import types

import torch.nn as nn

REF_ROOT = os.environ.get('LEDNET_REFERENCE_ROOT', '/root/reference')
_PREFIX = '_ledref'          # private module namespace: never shadows a real mmseg


def available():
    return os.path.isfile(os.path.join(REF_ROOT, 'mmseg/models/backbones/ddrnet.py'))


class _Registry:
    """Dict-backed stand-in for mmengine.Registry: register_module + build."""

    def __init__(self):
        self.table = {'ReLU': nn.ReLU}

    def register_module(self, name=None, force=False, module=None):
        def deco(cls):
            self.table[name or cls.__name__] = cls
            return cls
        return deco(module) if module is not None else deco

    def build(self, cfg):
        cfg = dict(cfg)
        return self.table[cfg.pop('type')](**cfg)


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg(name):
    m = _mod(name)
    m.__path__ = []
    return m


def _load(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(REF_ROOT, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[modname] = m
    spec.loader.exec_module(m)
    return m


_cache = {}


def load():
    """Return a namespace with the reference's classes/functions, executed verbatim."""
    if _cache:
        return _cache['ns']
    if not available():
        raise FileNotFoundError(f'reference tree not found at {REF_ROOT}')
    from . import mmcv_shim

    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split('.')[0] in ('mmcv', 'mmengine', 'mmseg', 'prettytable', 'timm')}

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    class BaseMetric:
        def __init__(self, collect_device='cpu', prefix=None, **kw):
            self.results, self.collect_device, self.prefix = [], collect_device, prefix
            self.dataset_meta = None

    class _Logger:
        @staticmethod
        def get_current_instance():
            return None

    models, metrics = _Registry(), _Registry()
    try:
        _pkg('mmcv')
        _mod('mmcv.cnn', ConvModule=mmcv_shim.ConvModule,
             build_norm_layer=mmcv_shim.build_norm_layer,
             build_activation_layer=mmcv_shim.build_activation_layer)
        _pkg('mmengine')
        _mod('mmengine.model', BaseModule=BaseModule, ModuleList=nn.ModuleList,
             Sequential=nn.Sequential)
        _mod('mmengine.dist', is_main_process=lambda: True)
        _mod('mmengine.evaluator', BaseMetric=BaseMetric)
        _mod('mmengine.logging', MMLogger=_Logger, print_log=lambda *a, **k: None)
        _mod('mmengine.utils', mkdir_or_exist=lambda p: os.makedirs(p, exist_ok=True))

        class PrettyTable:
            def add_column(self, *a, **k): pass
            def get_string(self): return ''
        _mod('prettytable', PrettyTable=PrettyTable)

        _pkg('mmseg')
        _mod('mmseg.registry', MODELS=models, METRICS=metrics)
        _mod('mmseg.utils', OptConfigType=object, ConfigType=object, SampleList=list,
             OptSampleList=object, OptMultiConfig=object, MultiConfig=object)
        _mod('mmseg.structures', build_pixel_sampler=lambda cfg, **kw: None)
        _pkg('mmseg.models')
        # mmseg.models.utils : wrappers, basic_block, ppm  (verbatim)
        utils = _pkg('mmseg.models.utils')
        wr = _load('mmseg.models.utils.wrappers', 'mmseg/models/utils/wrappers.py')
        bb = _load('mmseg.models.utils.basic_block', 'mmseg/models/utils/basic_block.py')
        ppm = _load('mmseg.models.utils.ppm', 'mmseg/models/utils/ppm.py')
        utils.__dict__.update(resize=wr.resize, Upsample=wr.Upsample, BasicBlock=bb.BasicBlock,
                              Bottleneck=bb.Bottleneck, DAPPM=ppm.DAPPM, PAPPM=ppm.PAPPM)
        # mmseg.models.losses : accuracy, ohem  (verbatim)
        losses = _pkg('mmseg.models.losses')
        acc = _load('mmseg.models.losses.accuracy', 'mmseg/models/losses/accuracy.py')
        ohem = _load('mmseg.models.losses.ohem_cross_entropy_loss',
                     'mmseg/models/losses/ohem_cross_entropy_loss.py')
        losses.__dict__.update(accuracy=acc.accuracy, Accuracy=acc.Accuracy,
                               OhemCrossEntropy=ohem.OhemCrossEntropy)
        # backbone (DDRNet = R0 body), heads, metric  (verbatim)
        _pkg('mmseg.models.backbones')
        ddr = _load('mmseg.models.backbones.ddrnet', 'mmseg/models/backbones/ddrnet.py')
        _pkg('mmseg.models.decode_heads')
        dh = _load('mmseg.models.decode_heads.decode_head',
                   'mmseg/models/decode_heads/decode_head.py')
        lh = _load('mmseg.models.decode_heads.led_head', 'mmseg/models/decode_heads/led_head.py')
        _pkg('mmseg.evaluation')
        _pkg('mmseg.evaluation.metrics')
        iou = _load('mmseg.evaluation.metrics.iou_metric',
                    'mmseg/evaluation/metrics/iou_metric.py')
        # SESP block (eesp.py needs `..classification.espnetv2_config` only for a constant table)
        nnl = _pkg('mmseg.models.nn_layers')
        _pkg('mmseg.models.classification')
        try:
            _load('mmseg.models.classification.espnetv2_config',
                  'mmseg/models/classification/espnetv2_config.py')
            eu = _load('mmseg.models.nn_layers.espnet_utils',
                       'mmseg/models/nn_layers/espnet_utils.py')
            nnl.espnet_utils = eu
            eesp = _load('mmseg.models.nn_layers.eesp', 'mmseg/models/nn_layers/eesp.py')
        except Exception:           # pragma: no cover - optional block
            eesp = None
        # MFAF gate (Muti_AFF, pure torch) and GETB (needs timm only for DropPath / trunc_normal_)
        mu = _load('mmseg.models.classification.model_utils',
                   'mmseg/models/classification/model_utils.py')
        import torch.nn.init as _init
        _pkg('timm')
        _pkg('timm.models')
        _mod('timm.models.layers', DropPath=nn.Identity, to_2tuple=lambda v: (v, v),
             trunc_normal_=_init.trunc_normal_)
        getb = _load('mmseg.models.backbones.UNetFormer_GETB',
                     'mmseg/models/backbones/UNetFormer_GETB.py')
    finally:
        # drop the stubs again so nothing else in the process sees a fake mmseg/mmcv
        for k in [k for k in sys.modules
                  if k.split('.')[0] in ('mmcv', 'mmengine', 'mmseg', 'prettytable', 'timm')]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v

    ns = types.SimpleNamespace(
        DDRNet=ddr.DDRNet, LEDHead=lh.LEDHead, BaseDecodeHead=dh.BaseDecodeHead,
        IoUMetric=iou.IoUMetric, OhemCrossEntropy=ohem.OhemCrossEntropy,
        accuracy=acc.accuracy, resize=wr.resize, BasicBlock=bb.BasicBlock,
        Bottleneck=bb.Bottleneck, DAPPM=ppm.DAPPM,
        SESP=getattr(eesp, 'SESP', None) if eesp else None, Muti_AFF=mu.Muti_AFF,
        GETBBlock=getb.GETBBlock, MODELS=models, METRICS=metrics)
    _cache['ns'] = ns
    return ns


# ---------------------------------------------------------------------------------------------------------------
# SEAM edge gate: inline code of the authors' speed prototype, tools/speed/ddrnet_speed.py (class DDRNet1).
# The file cannot be imported (it imports `model_utils_speed`, `thop`, mmcv, and its __init__ calls `.cuda()`), and the
# edge path is not a module of its own.  It IS executed verbatim here: the statements are sliced out of the file's AST
# (nothing is copied into this repository) and compiled into two functions:
#   __init__ : the assignments of self.conv_1 / conv_2 / laplacian_kernel / fusion_kernel / boundary_threshold
#              (ddrnet_speed.py:88-113)
#   forward  : `seg_label = self.conv_1(x)` ... `seg_labels = boudary_targets_pyramid.float()` (:282-338) followed by
#              `result = self.conv_2(seg_labels) * x_s` / `x_s = result + x_s` (:388-389)
# plus the module-level `normalize_tensor` (:24-37).  ConvModule is oracle/mmcv_shim's (third-party restated, as for
# every other block); `.cuda()` is a no-op while the sliced __init__ runs.
def load_seam():
    import ast

    import torch
    import torch.nn.functional as F

    from . import mmcv_shim

    if 'seam' in _cache:
        return _cache['seam']
    path = os.path.join(REF_ROOT, 'tools/speed/ddrnet_speed.py')
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    norm_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'normalize_tensor')
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'DDRNet1')
    init = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == '__init__')
    fwd = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == 'forward')

    def self_attr_target(st):
        if isinstance(st, ast.Assign) and len(st.targets) == 1:
            t = st.targets[0]
            if isinstance(t, ast.Attribute) and isinstance(t.value, ast.Name) and t.value.id == 'self':
                return t.attr
        return None

    def name_target(st):
        if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
            return st.targets[0].id
        return None

    wanted = ('conv_1', 'conv_2', 'laplacian_kernel', 'fusion_kernel', 'boundary_threshold')
    init_stmts = [st for st in init.body if self_attr_target(st) in wanted]
    assert sorted(self_attr_target(s) for s in init_stmts) == sorted(wanted), 'prototype __init__ changed'
    body = fwd.body
    i0 = next(i for i, st in enumerate(body) if name_target(st) == 'seg_label')
    i1 = next(i for i, st in enumerate(body) if name_target(st) == 'seg_labels')
    j0 = next(i for i, st in enumerate(body) if name_target(st) == 'result')
    assert i0 < i1 < j0 and name_target(body[j0 + 1]) == 'x_s'
    edge_stmts, gate_stmts = body[i0:i1 + 1], body[j0:j0 + 2]

    def make_fn(name, args, stmts, ret_src):
        fn = ast.parse(f'def {name}({args}):\n    pass\n    return {ret_src}').body[0]
        fn.body = list(stmts) + [fn.body[-1]]
        return fn

    mod = ast.Module(body=[norm_fn,
                           make_fn('seam_init', 'self', init_stmts, 'None'),
                           make_fn('seam_forward', 'self, x, x_s', edge_stmts + gate_stmts, 'x_s, seg_labels, seg_label')],
                     type_ignores=[])
    ast.fix_missing_locations(mod)
    g = {'torch': torch, 'F': F, 'nn': nn, 'ConvModule': mmcv_shim.ConvModule, 'math': __import__('math')}
    exec(compile(mod, path, 'exec'), g)

    class RefSEAM(nn.Module):
        """The prototype's own statements (see load_seam) wrapped as a module: forward(x, x_s) -> gated x_s;
        `edge` returns (0/1 mask, normalised edge response)."""

        def __init__(self, norm_cfg=None):
            super().__init__()
            self.norm_cfg = norm_cfg or dict(type='BN', requires_grad=True)
            self.act_cfg = dict(type='ReLU', inplace=True)
            cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda t, *a, **k: t          # ddrnet_speed.py:88 moves the Laplacian to the GPU
            try:
                g['seam_init'](self)
            finally:
                torch.Tensor.cuda = cuda

        def forward(self, x, x_s):
            return g['seam_forward'](self, x, x_s)[0]

        def edge(self, x):
            _, m, e = g['seam_forward'](self, x, torch.zeros_like(x))
            return m, e

    _cache['seam'] = RefSEAM
    return RefSEAM


# ---- stack_batch (mmseg/utils/misc.py:30-128), executed where it lies: the module's only package-relative import is a
#      typing alias, stubbed here.
def load_stack_batch():
    if 'stack_batch' in _cache:
        return _cache['stack_batch']
    if not os.path.isfile(os.path.join(REF_ROOT, 'mmseg/utils/misc.py')):
        raise FileNotFoundError(f'reference tree not found at {REF_ROOT}')
    pkg = _PREFIX + '_utils'
    saved = {k: sys.modules.get(k) for k in (pkg, pkg + '.typing_utils', pkg + '.misc')}
    try:
        _pkg(pkg)
        _mod(pkg + '.typing_utils', SampleList=list)
        m = _load(pkg + '.misc', 'mmseg/utils/misc.py')
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _cache['stack_batch'] = m.stack_batch
    return m.stack_batch


# ---------------------------------------------------------------------------------------------------------------
# The LED wiring: class DDRNet1 of the authors' speed prototype (tools/speed/ddrnet_speed.py:39-406), the closest public
# statement of figure 3 (the registered backbones/lednet.py is withheld).  The file cannot be imported (model_utils_speed,
# thop, mmcv, `.cuda()` in __init__), so the CLASS DEFINITION is taken from the file's AST and executed where it lies
# (nothing is copied) in a namespace that supplies its names:
#   ConvModule / build_norm_layer ... oracle/mmcv_shim (third-party, restated)
#   DAPPM, BasicBlock, Bottleneck, resize, GETB `Block`, Muti_AFF ... the reference's own files via load()
#   STDCModule ... mmseg/models/backbones/stdc.py executed where it lies
# Muti_AFF: the prototype imports the BN-less copy in tools/speed/model_utils_speed.py; the registered package's
# classification/model_utils.py version (with BatchNorm) is used here, as the package's own LED-Net would.
def load_ddrnet1():
    import ast
    import math

    import torch
    import torch.nn.functional as F

    from . import mmcv_shim

    if 'ddrnet1' in _cache:
        return _cache['ddrnet1']
    ref = load()
    path = os.path.join(REF_ROOT, 'tools/speed/ddrnet_speed.py')
    tree = ast.parse(open(path).read(), filename=path)
    norm_fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'normalize_tensor')
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == 'DDRNet1')
    cls.decorator_list = []                                   # @MODELS.register_module(): no registry here

    class BaseModule(nn.Module):
        def __init__(self, init_cfg=None):
            super().__init__()
            self.init_cfg = init_cfg

    # stdc.py where it lies, behind stubs for its imports
    saved = {k: sys.modules.get(k) for k in list(sys.modules)
             if k.split('.')[0] in ('mmcv', 'mmengine', 'mmseg')}
    try:
        _pkg('mmcv')
        _mod('mmcv.cnn', ConvModule=mmcv_shim.ConvModule, build_norm_layer=mmcv_shim.build_norm_layer)
        _pkg('mmengine')
        _mod('mmengine.model', BaseModule=BaseModule, ModuleList=nn.ModuleList, Sequential=nn.Sequential)
        _pkg('mmseg')
        _mod('mmseg.registry', MODELS=_Registry())
        _pkg('mmseg.models')
        _mod('mmseg.models.utils', resize=ref.resize)
        _pkg('mmseg.models.backbones')
        _mod('mmseg.models.backbones.bisenetv1', AttentionRefinementModule=nn.Identity)
        stdc = _load(_PREFIX + '_stdc', 'mmseg/models/backbones/stdc.py')
    finally:
        for k in [k for k in sys.modules if k.split('.')[0] in ('mmcv', 'mmengine', 'mmseg')]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
        sys.modules.pop(_PREFIX + '_stdc', None)

    mod = ast.Module(body=[norm_fn, cls], type_ignores=[])
    ast.fix_missing_locations(mod)
    g = {'torch': torch, 'nn': nn, 'F': F, 'math': math, 'np': __import__('numpy'),
         'ConvModule': mmcv_shim.ConvModule, 'build_norm_layer': mmcv_shim.build_norm_layer,
         'BaseModule': BaseModule, 'Sequential': nn.Sequential, 'OptConfigType': object,
         'DAPPM': ref.DAPPM, 'BasicBlock': ref.BasicBlock, 'Bottleneck': ref.Bottleneck, 'resize': ref.resize,
         'STDCModule': stdc.STDCModule, 'Block': ref.GETBBlock, 'GlobalLocalAttention': None,
         'Muti_AFF': ref.Muti_AFF}
    exec(compile(mod, path, 'exec'), g)
    proto = g['DDRNet1']

    class RefLEDTrunk(proto):
        """DDRNet1 built on the CPU (its __init__ moves the Laplacian to the GPU: `.cuda()` is a no-op while it runs) plus
        the two stem taps LEDHead consumes (stem[0], stem[1] outputs: the R0 contract)."""

        def __init__(self, **kw):
            cuda = torch.Tensor.cuda
            torch.Tensor.cuda = lambda t, *a, **k: t
            try:
                super().__init__(**kw)
            finally:
                torch.Tensor.cuda = cuda

        def forward_with_taps(self, x):
            taps = {}
            h0 = self.stem[0].register_forward_hook(lambda m, i, o: taps.__setitem__('x1', o.clone()))
            h1 = self.stem[1].register_forward_hook(lambda m, i, o: taps.__setitem__('x2', o.clone()))
            try:
                c5 = self.forward(x)
            finally:
                h0.remove(); h1.remove()
            return c5, taps['x1'], taps['x2']

    _cache['ddrnet1'] = RefLEDTrunk
    return RefLEDTrunk
