"""CPU-only: the ensemble-statistics oracle (oracle/ensemble_ref.py) against closed forms and the brute-force
CRPS definition E|X - y| - 1/2 E|X - X'| (edge cases: one member, identical members, ties)."""
import numpy as np

from oracle.ensemble_ref import crps_bruteforce, ensemble_statistics


def test_crps_sorted_form_equals_bruteforce():
    rng = np.random.default_rng(0)
    for m in (2, 3, 16, 64):
        x = rng.normal(size=(m, 5, 7))
        x[0] = x[1]                                   # a tie
        y = rng.normal(size=(5, 7))
        s = ensemble_statistics(x, y)
        assert np.allclose(s["crps"], crps_bruteforce(x, y), rtol=1e-12, atol=1e-12)
        assert np.allclose(s["mean"], x.mean(0)) and np.allclose(s["std"], x.std(0, ddof=1))


def test_crps_edge_cases():
    y = np.array([[0.5, -1.0]])
    one = np.array([[[2.0, 3.0]]])
    s = ensemble_statistics(one, y)
    assert np.allclose(s["crps"], np.abs(one[0] - y)) and np.all(s["std"] == 0)      # one member: absolute error
    same = np.repeat(one, 8, axis=0)
    assert np.allclose(ensemble_statistics(same, y)["crps"], np.abs(one[0] - y))     # no spread: absolute error
    perfect = np.stack([y - 1.0, y + 1.0])
    assert np.allclose(ensemble_statistics(perfect, y)["crps"], 1.0 - 0.5 * 1.0)     # E|X-y| = 1, E|X-X'| = 1
