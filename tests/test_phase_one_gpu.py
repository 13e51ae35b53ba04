"""Stand-alone PhaseOne.PhaseOneSolver on the B200 engine vs the reference's functional tests
(AutomatedTestsPhaseOne.py:235-343; fixtures recorded from the real class in tests/golden)."""

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu
PHASE1 = load_golden("phase_one_standalone.json")


@pytest.mark.parametrize("case", [c for c in PHASE1 if "G" in c], ids=[c["name"] for c in PHASE1 if "G" in c])
def test_polytopes(case):
    from ipm_b200.PhaseOne import PhaseOneSolver

    s = PhaseOneSolver(np.array(case["G"]), np.array(case["h"]), case["mu"], x0=np.array(case["x0"]))
    x, sv, warn = s.solve()
    print(case["name"], np.asarray(x), sv, case["x"], case["s"])
    assert warn == case["warn"]
    assert sv == pytest.approx(case["s"], rel=1e-6, abs=1e-9)
    np.testing.assert_allclose(np.asarray(x), case["x"], rtol=1e-6, atol=1e-8)
    if case["name"] == "empty_set":
        assert sv > 0
    elif case["name"] != "inside":
        G, h = np.array(case["G"]), np.array(case["h"])
        assert sv < 0 and np.max(G @ np.asarray(x) - h) <= 0


def test_random_polytope():
    from ipm_b200.PhaseOne import PhaseOneSolver

    case = [c for c in PHASE1 if c["name"] == "random_60x40"][0]
    np.random.seed(case["seed"])
    G = np.random.randn(case["m"], case["n"])
    xf = np.random.randn(case["n"])
    h = G @ xf + np.random.rand(case["m"])
    x, sv, warn = PhaseOneSolver(G, h, case["mu"]).solve()
    assert sv == pytest.approx(case["s"], rel=1e-6)
    assert sv < 0 and np.max(G @ np.asarray(x) - h) <= 0
    np.testing.assert_allclose(np.asarray(x), case["x"], rtol=1e-5, atol=1e-7)
