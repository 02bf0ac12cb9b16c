"""Seeded synthetic weights, inputs and state-dict shape tables for the benchmark, the smoke test and the tests.

Data generators only (no reference arithmetic lives here): bench.py's B200 arm may not touch ``oracle/``, so the generators it
shares with the tests are kept outside it; ``oracle/synth.py`` and ``oracle/shapes.py`` re-export them for the test code."""
