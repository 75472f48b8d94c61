"""Moved: the seeded generators live in deeprecommendation_b200/synth.py (bench.py may not import oracle/ outside its
cpu_baseline leg).  Re-exported here so oracle code and older imports keep working."""
from deeprecommendation_b200.synth import *  # noqa: F401,F403
from deeprecommendation_b200.synth import _counts_lognormal, _linear, _mlp  # noqa: F401
