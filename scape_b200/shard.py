"""Multi-GPU partitioning of the infer_pa path.

UTRs are independent; the only coupling is the per-file RNG stream, so the unit of distribution is
the chunk file (one stream).  Streams are bin-packed over the ranks by estimated cost with the
longest-processing-time rule; every rank fits its own streams on its own GPU and the results are
gathered on the host in input order.  There is no data-path collective (nothing to exchange);
`torch.distributed` is only used by the callers for the barrier / max-over-ranks timing and for
gathering result objects.
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def utr_cost(n_reads: int, utr_len_hint: int = 2000, n_max_apa: int = 5, n_min_apa: int = 1) -> float:
    """Cheap a-priori cost of one UTR, before binning: likelihood phases ~ N*T*(S+307) exp, EM ~
    chains * iterations * window * B * N MACs.  N (bins) grows roughly like
    reads^0.55 on 10x-like data (27,829 reads -> 1,376 bins; 500 -> ~230)."""
    n_bins = min(float(n_reads), 7.0 * float(n_reads) ** 0.55)
    T = max(utr_len_hint, 2000) / 9.0
    chains = 10.0 * (n_max_apa - n_min_apa + 1)
    lik = n_bins * T * 320.0 * 4.0
    em = chains * 12.0 * (T / 2.0) * 13.0 * n_bins
    return lik + em


def stream_costs(reads_per_utr: Sequence[Sequence[int]], utr_len_hint: Sequence[Sequence[int]] = None) -> np.ndarray:
    out = np.zeros(len(reads_per_utr))
    for s, counts in enumerate(reads_per_utr):
        hints = utr_len_hint[s] if utr_len_hint is not None else [2000] * len(counts)
        out[s] = sum(utr_cost(int(c), int(h)) for c, h in zip(counts, hints))
    return out


def lpt_partition(costs: Sequence[float], n_ranks: int) -> List[List[int]]:
    """Longest-processing-time bin packing.  Deterministic: ties broken by stream index, each
    rank's list is returned in ascending stream order."""
    costs = np.asarray(costs, dtype=float)
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0.0] * n_ranks
    parts: List[List[int]] = [[] for _ in range(n_ranks)]
    for i in order:
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        parts[r].append(i)
        load[r] += float(costs[i])
    return [sorted(p) for p in parts]


def imbalance(costs: Sequence[float], parts: List[List[int]]) -> float:
    """max rank load / mean rank load."""
    costs = np.asarray(costs, dtype=float)
    loads = np.array([costs[p].sum() if len(p) else 0.0 for p in parts])
    return float(loads.max() / max(loads.mean(), 1e-300))
