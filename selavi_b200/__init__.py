"""selavi_b200 — B200-native (sm_100a) implementation of SeLaVi's data-parallel training hot path.

Mirrors of the reference's Python interface for this path (same names, arguments and error behaviour):
  selavi_b200.model      <- model.py        (AVModel, load_model / get_model)
  selavi_b200.sk_utils   <- src/sk_utils.py (optimize_L_sk_gpu / optimize_L_sk_multi, ...)
  selavi_b200.utils      <- utils.py        (get_loss)
  selavi_b200.optim      fused SGD with torch.optim.SGD's rule
All compute goes through the C ABI of include/selavi_b200.h (libselavi_b200.so, hand-written CUDA).
"""
__version__ = "0.1.0"
