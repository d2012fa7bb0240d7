"""Synthetic inputs for the oracle-side tests: the generator itself lives in the product package
(myimagecaptioningmodel_b200/synth.py) so that bench.py's own arm never imports oracle/."""
from myimagecaptioningmodel_b200.synth import make_batch, make_captions, make_features  # noqa: F401
