from tgsr_b200.attention import GlobalAttentionGeneral, conv1x1, func_attention  # noqa: F401
