"""Install the reference's own hot-path modules, UNMODIFIED, into git-ignored baseline/_ref/ (BASELINE.md section 4).

    python baseline/install_ref.py [--src /root/reference]

The reference is plain Python without packaging (`pip install /root/reference` has nothing to build), so the install is
a byte-for-byte copy of the six modules the hot path lives in, the package marker and the eval driver
(trainer_objective.py with its datasets.py / miscc/utils.py helpers), with a sha256 manifest.
baseline/_ref/ is git-ignored (never in history) but travels to the GPU box with the snapshot, where
`bench.py --impl reference` and the `cpu_baseline` leg import it through baseline/ref_arm.py.  Nothing of the product
(tgsr_b200/) reads it.
"""
import argparse
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
FILES = ['GlobalAttention.py', 'util.py', 'model.py', 'models16.py', 'miscc/__init__.py', 'miscc/config.py', 'miscc/losses.py',
         # the eval driver and its helpers: tests/test_gpu_reference_driver.py runs the UNMODIFIED condGANTrainer.gen_exampleSRHL
         # against tgsr_b200/compat
         'trainer_objective.py', 'datasets.py', 'miscc/utils.py']


def install(src: str = '/root/reference', quiet: bool = False) -> bool:
    if not os.path.isfile(os.path.join(src, 'GlobalAttention.py')):
        return False
    manifest = {}
    for f in FILES:
        d = os.path.join(DST, f)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(os.path.join(src, f), d)
        manifest[f] = hashlib.sha256(open(d, 'rb').read()).hexdigest()
    json.dump({'source': src, 'sha256': manifest}, open(os.path.join(DST, 'MANIFEST.json'), 'w'), indent=1)
    if not quiet:
        print('installed %d reference modules into %s' % (len(FILES), DST))
    return True


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    a = ap.parse_args()
    if not install(a.src):
        raise SystemExit('reference tree not found at %s' % a.src)
