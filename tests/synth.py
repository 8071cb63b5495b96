"""Synthetic inputs live in the package (bench.py uses them too); re-exported here for the tests."""
from mods_b200.synth import *  # noqa: F401,F403
