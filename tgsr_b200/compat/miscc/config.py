from tgsr_b200.config import cfg, cfg_from_file  # noqa: F401
