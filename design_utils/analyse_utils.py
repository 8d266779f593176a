"""Hot-path subset of the reference's ``design_utils/analyse_utils.py``: only ``calculate_seq_metrics`` (analyse_utils.py:351-371)
sits on the sampling path (called at sampling_utils.py:132).  Backed by timed_design_b200.seq_metrics -- the residue tables are
a recollection of ampal's (ampal is not vendored): parity unverified, SURVEY.md App. G."""
from timed_design_b200.seq_metrics import calculate_seq_metrics  # noqa: F401
