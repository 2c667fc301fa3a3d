"""CPU oracle for the SBGM_DANRA hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a CPU restatement (plain torch fp32 / numpy, functional style over a
state-dict) of the reference's hot path: the conditional score-UNet forward
(`sbgm/score_unet.py`), the VE-SDE samplers (`sbgm/score_sampling.py`) and the DSM loss.
It exists so that the CUDA path can be checked on a GPU box where `/root/reference` is
not mounted.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference`
leg may import anything from here.  The product package (`sbgm_danra_b200/`) never does.

Parity pin: the reference ships no tests or golden vectors of its own.
The oracle is therefore pinned against outputs of the reference itself, imported from
`/root/reference` in the build container by `tests/golden/make_golden.py`; the resulting
fixtures are committed under `tests/golden/` and `tests/test_oracle_golden.py` re-checks the
oracle against them on every run.
"""
