"""Synthetic checkpoint / input generators live in marconet_b200.testing.synth (they contain no reference arithmetic and
the benchmark's product leg needs them); re-exported here for the oracle tooling and the tests."""
from marconet_b200.testing.synth import *  # noqa: F401,F403
from marconet_b200.testing.synth import _Gen  # noqa: F401
