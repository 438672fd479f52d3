"""`scape.apa_core` namespace of the reference (apa_core.py), served by scape_b200."""
from scape_b200.apa_core import (  # noqa: F401
    Parameters, infer_pa, _infer_pa, infer, infer_files, subsample_run, fit_chunks, exp_pa_len,
    cal_exp_pa_len_by_cluster, watch_dog, run_infer_pa,
)
