"""diarizen_b200: B200-native (sm_100a) implementation of the DiariZen inference hot path."""
from .archs import SegArch, get_arch, init_state_dict  # noqa: F401

__version__ = "0.1.0"
