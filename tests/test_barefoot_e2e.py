"""End-to-end functional parity of the orchestrator (BASELINE config #1 and three siblings): the
B200 `barefoot` class against the UNMODIFIED reference orchestrator run on the CPU under the oracle
shims (oracle/gen_config1.py -> tests/golden/config1.npz).  Same seeds, same toy models: every
selected point (model index, iteration, coordinates) must be identical."""
import numpy as np
import pytest

import toy_models
from conftest import golden

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", toy_models.CASES, ids=[c[0] for c in toy_models.CASES])
def test_selections_match_reference(case, tmp_path):
    from barefoot_framework_b200.barefoot import barefoot
    name, ctor, init, seed = case
    ev, it = toy_models.run_case(barefoot, name, ctor, init, seed, str(tmp_path))
    g = golden("config1.npz")
    ref_ev, ref_it = g[name + "_evaluatedPoints"], g[name + "_iterationData"]
    assert ev.shape == ref_ev.shape
    np.testing.assert_array_equal(ev[:, :2], ref_ev[:, :2])            # model index, iteration
    np.testing.assert_array_equal(ev[:, 3:], ref_ev[:, 3:])            # selected coordinates: bit-exact
    np.testing.assert_allclose(ev[:, 2], ref_ev[:, 2], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(np.delete(it, 2, axis=1), ref_it, rtol=1e-12, atol=1e-12)
    # outputs on disk like the reference
    assert (tmp_path / "results" / name / "evaluatedPoints.csv").is_file()
    assert (tmp_path / "results" / name / "iterationData.csv").is_file()
