"""Synthetic weights and inputs (TEST INFRASTRUCTURE): re-export of `sbgm_danra_b200.synth`.

The generators are shared by the parity tests, the oracle and the benchmarks, so they live beside the package
(`bench.py` and `tools/` must not import anything under `oracle/`); the oracle and the tests keep this import path."""
from sbgm_danra_b200.synth import (FMAP_CHANNELS, Batch, NetConfig, config_for, decoder_plan, param_schema,  # noqa: F401
                                   synth_batch, synth_state_dict)
