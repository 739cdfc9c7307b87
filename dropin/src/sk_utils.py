"""Drop-in for the reference's `src/sk_utils.py` (cluster, get_cluster_assignments_gpu, optimize_L_sk_gpu, match_order)."""
from selavi_b200.sk_utils import (cluster, get_cluster_assignments_gpu, match_order, optimize_L_sk_gpu,  # noqa: F401
                                  optimize_L_sk_multi)
