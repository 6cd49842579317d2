"""-m gpu: whole-graph parity of the CUDA train step / inference forward against the oracle and golden fixtures."""
import pytest

import graph_cases as G


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(G.CASES))
def test_graph_case(name):
    r = G.CASES[name]()
    assert r["ok"], f"{name}: {r}"
