"""Re-export of the synthetic data generators (moved to synthdata/ so that bench.py's B200 arm does not import oracle/)."""
from synthdata.synth import *            # noqa: F401,F403
from synthdata.synth import _gen         # noqa: F401
