"""Host-side data plane with the reference's `processing` API (ark / feature_reader /
batchdispenser / target_coder / readfiles), re-implemented for Python 3 and a fast feeder."""
from . import ark, batchdispenser, feature_reader, readfiles, target_coder  # noqa: F401
