"""Re-export of the state-dict shape tables (moved to synthdata/ so that bench.py's B200 arm does not import oracle/)."""
from synthdata.shapes import *           # noqa: F401,F403
