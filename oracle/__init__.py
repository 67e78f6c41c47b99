"""TEST INFRASTRUCTURE ONLY -- never imported by the product package `tgsr_b200`.

`oracle/` holds a CPU restatement (plain fp32 PyTorch functional ops, no
nn.Module state) of the TGSR generator forward path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it, and only as the checker / the reported CPU baseline.

Pinning: the reference repository ships no tests or golden vectors for this
path.  The restatement is therefore pinned by
executing the reference's own modules in the authoring container
(`oracle/make_golden.py`, which imports /root/reference through
`oracle/ref_loader.py`) and committing their inputs/outputs under
`tests/golden/`.  `tests/test_oracle_golden.py` replays every fixture.
"""
