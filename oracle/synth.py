"""Re-export of the deterministic synthetic weights / inputs (mrn_b200/synth.py) for the oracle-side scripts."""
from mrn_b200.synth import *  # noqa: F401,F403
from mrn_b200.synth import _rng  # noqa: F401
