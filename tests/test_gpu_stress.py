"""A short run of the randomised API-sequence stress (scripts/stress.py): add / search / reset sequences over mixed sizes,
k, metrics, input kinds and one-/multi-device indexes, every search checked against the oracle."""
import importlib.util
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [3, 4])
def test_random_api_sequences_match_the_oracle(seed):
    spec = importlib.util.spec_from_file_location("agp_stress", Path(__file__).resolve().parent.parent / "scripts" / "stress.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    n_seq, n_checks = mod.run(budget=12.0, seed=seed)
    assert n_seq > 10 and n_checks > 10
