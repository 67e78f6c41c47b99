"""Subprocess body of tests/test_gpu_reference_driver.py: runs the reference's UNMODIFIED eval driver
(baseline/_ref/trainer_objective.py: condGANTrainer.gen_exampleSRHL, trainer_objective.py:55-165) with
tgsr_b200/compat FIRST on sys.path, so every `from model import ...` / `from miscc.config import cfg` of the driver
resolves to the B200 implementation.  Environment stand-ins only (none touches the driver's code): nltk / skimage import
shims (absent packages), a default PIL font for the hard-coded Windows font path (miscc/utils.py:36), and a loader whose
iterator has the Python-2 `.next()` the driver calls (trainer_objective.py:106).

    python tests/_run_reference_driver.py <workdir> <precision>
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
work, precision = sys.argv[1], sys.argv[2]
sys.path[:0] = [os.path.join(ROOT, 'baseline', 'shims'), os.path.join(ROOT, 'tgsr_b200', 'compat'),
                os.path.join(ROOT, 'baseline', '_ref'), ROOT]

import torch  # noqa: E402
from PIL import ImageFont  # noqa: E402

_default_font = ImageFont.load_default()
ImageFont.truetype = lambda *a, **k: _default_font

import tgsr_b200  # noqa: E402

tgsr_b200.set_precision(precision)
from miscc.config import cfg, cfg_from_file  # noqa: E402  (tgsr_b200/compat/miscc/config.py)

cfg_from_file(os.path.join(work, 'eval.yml'))

import model  # noqa: E402
import trainer_objective  # noqa: E402

assert os.path.join('baseline', '_ref') in trainer_objective.__file__, trainer_objective.__file__
assert os.path.join('tgsr_b200', 'compat') in model.__file__, model.__file__
assert model.G_SR_NET_low.__module__.startswith('tgsr_b200'), model.G_SR_NET_low.__module__


class _Iter(object):
    def __init__(self, it):
        self._it = it

    def __iter__(self):
        return self

    def __next__(self):
        return next(self._it)

    next = __next__                      # the driver's Python-2 spelling (trainer_objective.py:106)


class Loader(object):
    def __init__(self, batches):
        self.batches = batches

    def __len__(self):
        return len(self.batches)

    def __iter__(self):
        return _Iter(iter(self.batches))


batches = torch.load(os.path.join(work, 'batches.pt'), weights_only=False)
ixtoword = {i: 'w%d' % i for i in range(64)}
trainer = trainer_objective.condGANTrainer(work, Loader(batches), 41, ixtoword, cfg)
try:
    trainer.gen_exampleSRHL('sr_out')
except StopIteration:                    # the driver asks for 100 test images (trainer_objective.py:59); the loader holds fewer
    pass
print('DRIVER_DONE')
