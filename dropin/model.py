"""Drop-in for the reference's `model.py`: put `dropin/` ahead of the reference on sys.path."""
from selavi_b200.model import *  # noqa: F401,F403
from selavi_b200.model import AVModel, load_model, get_model  # noqa: F401
