"""Drop-in for the reference's `datasets/audio_utils.py:get_spec` (main-process use; see selavi_b200/audio_utils.py)."""
from selavi_b200.audio_utils import get_spec, logfbank_batch  # noqa: F401
