"""Global configuration object with the reference's key set (miscc/config.py:10-67).

Modules read ``cfg`` inside their constructors exactly like the reference does
(util.py:179,185,377-378,408; model.py:37-39), so the same YAML files drive both.
``cfg_from_file`` keeps the reference's strict merge (unknown key -> KeyError, type
mismatch -> ValueError; miscc/config.py:70-100) but uses yaml.safe_load.
"""
from __future__ import annotations


class AttrDict(dict):
    """dict with attribute access; nested dicts are converted on assignment."""

    def __init__(self, mapping=None):
        super().__init__()
        for k, v in (mapping or {}).items():
            self[k] = v

    def __setitem__(self, key, value):
        if isinstance(value, dict) and not isinstance(value, AttrDict):
            value = AttrDict(value)
        super().__setitem__(key, value)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key) from None

    __setattr__ = __setitem__


def _defaults() -> AttrDict:
    return AttrDict({
        'METHOD': 'S16', 'DATASET_NAME': 'birds', 'CONFIG_NAME': '', 'DATA_DIR': '', 'DATA_DIRIM': '',
        'GPU_ID': 0, 'CUDA': True, 'WORKERS': 6, 'RNN_TYPE': 'LSTM', 'B_VALIDATION': False,
        'TREE': {'BRANCH_NUM': 5, 'BASE_SIZE': 64},
        'TRAIN': {'BATCH_SIZE': 64, 'MAX_EPOCH': 600, 'SNAPSHOT_INTERVAL': 2000, 'DISCRIMINATOR_LR': 2e-4,
                  'GENERATOR_LR': 2e-4, 'ENCODER_LR': 2e-4, 'RNN_GRAD_CLIP': 0.25, 'FLAG': False, 'NET_E': '',
                  'NET_G': '', 'B_NET_D': True,
                  'SMOOTH': {'GAMMA1': 5.0, 'GAMMA3': 10.0, 'GAMMA2': 5.0, 'LAMBDA': 1.0}},
        'GAN': {'DF_DIM': 64, 'GF_DIM': 128, 'Z_DIM': 100, 'CONDITION_DIM': 100, 'R_NUM': 2,
                'B_ATTENTION': True, 'B_DCGAN': False},
        'TEXT': {'CAPTIONS_PER_IMAGE': 10, 'EMBEDDING_DIM': 256, 'WORDS_NUM': 18},
    })


cfg = _defaults()


def _merge(src: dict, dst: AttrDict, path: str = '') -> None:
    for k, v in src.items():
        if k not in dst:
            raise KeyError('{} is not a valid config key'.format(path + k))
        old = dst[k]
        if isinstance(old, AttrDict):
            if not isinstance(v, dict):
                raise ValueError('Type mismatch (dict vs. {}) for config key: {}'.format(type(v), path + k))
            _merge(v, old, path + k + '.')
        else:
            if type(old) is not type(v):
                raise ValueError('Type mismatch ({} vs. {}) for config key: {}'.format(type(old), type(v), path + k))
            dst[k] = v


def cfg_from_file(filename: str) -> None:
    """Load a YAML file and merge it into the defaults (miscc/config.py:103-109)."""
    import yaml
    with open(filename, 'r', encoding='UTF-8') as f:
        _merge(yaml.safe_load(f) or {}, cfg)


def cfg_from_dict(d: dict) -> None:
    _merge(d, cfg)


def set_eval_defaults() -> None:
    """The model dimensions of cfg/eval_faceSR_attn2.yml / eval_birdSR_attn2.yml (:5-7, :24-31)."""
    cfg_from_dict({'TREE': {'BRANCH_NUM': 4, 'BASE_SIZE': 32},
                   'GAN': {'DF_DIM': 64, 'GF_DIM': 32, 'Z_DIM': 100, 'R_NUM': 2},
                   'TEXT': {'EMBEDDING_DIM': 256}})
