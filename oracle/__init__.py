"""CPU oracle for the LED-Net hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-PyTorch (CPU, fp32, NCHW) restatement of the reference
algorithm for the one path this repo accelerates:

    backbone forward -> LEDHead -> 3-level logit fusion -> argmax -> IoU
    confusion matrix, plus OHEM cross-entropy / accuracy for the training step.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline - never from the product package ``led-net_b200/`` (a test
enforces that).  The product path fails loudly when its CUDA library is missing.

Parity pin status: PINNED against the reference's own modules.  The reference
(``/root/reference``, an mmsegmentation-1.2.2 fork) cannot be imported as a
package (mmcv/mmengine absent, ``backbones/lednet.py`` withheld), but its
pure-torch files execute verbatim when loaded by path behind small
``sys.modules`` stubs (``oracle/ref_loader.py``).  ``tests/golden/make_golden.py``
ran those verbatim files in the build container on seeded inputs and committed
the outputs under ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks
this restatement against them on every run (no ``/root/reference`` needed).
What stays restated rather than executed is mmcv's ``ConvModule`` wrapper
(third-party, ``mmcv>=2.0.0rc4,<2.2.0`` per ``mmseg/__init__.py:10-11``) and the
LED-Net trunk wiring itself, whose source the authors withhold
(``mmseg/models/backbones/lednet.py:1-8``); the trunk used is R0 = the
DDRNet-23-slim body (``mmseg/models/backbones/ddrnet.py``) plus two stem taps,
which satisfies the output contract ``LEDHead`` consumes.
"""

from .mmcv_shim import ConvModule, build_norm_layer, build_activation_layer  # noqa: F401
from .r0 import OracleLEDNet, BasicBlock, Bottleneck, DAPPM, resize  # noqa: F401
from .head import OracleLEDHead, fuse_logits  # noqa: F401
from .metrics import (intersect_and_union, total_area_to_metrics,  # noqa: F401
                      confusion_matrix, confusion_to_areas, compute_metrics)
from .losses import ohem_cross_entropy, accuracy  # noqa: F401
from .segmentor import (OracleSegmentor, preprocess, stack_batch,  # noqa: F401
                        postprocess_argmax, slide_inference)
