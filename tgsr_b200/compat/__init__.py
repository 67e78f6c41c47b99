"""Drop-in import aliases: put this directory FIRST on sys.path and the reference driver's
`from model import G_SR_NET_low, NetG_highweight`, `from GlobalAttention import func_attention`,
`from miscc.config import cfg, cfg_from_file` resolve to the B200 implementation.  See INTEGRATION.md."""
