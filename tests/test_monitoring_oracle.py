"""CPU: the oracle's extreme-value sentinel (oracle/monitoring_ref.py) against the outputs of the reference's own
`sbgm/utils.py::report_precip_extremes`, committed by tests/golden/make_monitoring_golden.py."""
import json
import os

from conftest import GOLDEN_DIR


def test_oracle_sentinel_reproduces_reference_outputs():
    from oracle import monitoring_ref
    with open(os.path.join(GOLDEN_DIR, "monitoring_golden.json")) as f:
        gold = json.load(f)
    cases = monitoring_ref.cases()
    assert len(gold) == 2 * len(cases)
    for name, x in cases.items():
        for cap in (500.0, 50.0):
            msgs = []
            got = monitoring_ref.report_precip_extremes(x, name=name, cap_mm_day=cap, logger=msgs.append)
            want = gold[f"{name}/cap{cap:g}"]
            assert got == want["result"], (name, cap)
            assert msgs == want["messages"], (name, cap)
    kinds = {k: v["result"] for k, v in gold.items()}
    assert kinds["spikes/cap500"]["has_extreme"] and kinds["negative_sample/cap500"].get("has_below_zero")
    assert kinds["both/cap500"].get("has_below_zero") and kinds["both/cap500"]["has_extreme"]
