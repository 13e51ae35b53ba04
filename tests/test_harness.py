"""CSV harness (ipm_b200/harness.py): file layout of the reference's testSolver.py / parseAndPlot.py."""

import numpy as np
import pytest

from conftest import load_golden


def test_csv_layout_round_trip(tmp_path):
    from ipm_b200 import harness

    f = tmp_path / "resLP.csv"
    n_values = np.array([100, 200])
    t = np.array([[0.5, 0.25, 0.0], [1.5, 0.0, 0.0]])
    cols = {"n_values": np.repeat(n_values, 3), "cvxpy_times": np.zeros(6), "cvxpy_values": np.zeros(6),
            "ls_gpu_times": t.ravel(), "ls_gpu_values": -t.ravel(), "ls_cpu_times": np.zeros(6),
            "ls_cpu_values": np.zeros(6), "jax_times": np.zeros(6), "jax_values": np.zeros(6)}
    harness._write(str(f), 2, 3, cols)
    lines = f.read_text().splitlines()
    assert lines[0] == "2,3"                                   # testSolver.py:268-271
    assert lines[1] == ("n_values,cvxpy_times,cvxpy_values,ls_gpu_times,ls_gpu_values,ls_cpu_times,ls_cpu_values,"
                        "jax_times,jax_values")              # the columns parse_csv("LP") reads
    assert len(lines) == 2 + 6
    N, num_tests, n_out, out = harness.parse_csv(str(f), "LP")
    assert (N, num_tests) == (3, 2) and list(n_out) == [100, 200]
    np.testing.assert_array_equal(np.isnan(out["ls_gpu_times"]), t == 0)   # zeros are "not run" (parseAndPlot.py:86)
    assert out["ls_gpu_values"][0, 1] == -0.25 and np.isnan(out["cvxpy_times"]).all()


def test_repetition_rule():
    from ipm_b200 import harness

    assert [harness._repetitions(n, 10) for n in (100, 999, 1000, 2499, 2500, 5000)] == [10, 10, 5, 5, 3, 3]


@pytest.mark.gpu
def test_lp_and_lasso_streams_match_the_reference_goldens(tmp_path):
    """The harness draws the reference's seed-1 streams: its first LP instances and its Lasso batch are exactly the
    golden cases recorded from the real reference."""
    from ipm_b200 import harness

    f = tmp_path / "b200_LP.csv"
    harness.test_LP([100], N=2, filename=str(f))
    _, _, n_values, out = harness.parse_csv(str(f), "LP")
    gold = {c["name"]: c for c in load_golden("barrier_cases.json")}
    assert list(n_values) == [100]
    assert out["ls_gpu_values"][0, 0] == pytest.approx(gold["lp_seed1_n100_0"]["value"], rel=1e-6)
    assert out["ls_gpu_values"][0, 1] == pytest.approx(gold["lp_seed1_n100_1"]["value"], rel=1e-6)
    assert np.all(out["ls_gpu_times"] > 0) and np.isnan(out["ls_cpu_times"]).all()
    g = tmp_path / "b200_LASSO.csv"
    _, v_gpu, _, _ = harness.test_LASSO([100], N=1, filename=str(g))
    assert (tmp_path / "b200_LASSOTimes.csv").read_text().splitlines()[0] == "1,1"
    vals = (tmp_path / "b200_LASSOValues.csv").read_text().splitlines()
    assert vals[0] == "cvxpy_values,lasso_gpu_values,lasso_cpu_values,lasso_jax_values" and len(vals) == 1 + 30
    np.testing.assert_allclose(v_gpu[0, 0, :3], [20.27555468, 22.9057783, 28.76297346], rtol=1e-6)
