"""Many independent adjustments on one GPU, run concurrently (BASELINE.json configs[4]: the
BatchRun.m sweep).

The reference's ``BatchRun.m:57-65`` calls ``main`` on one data folder after the other.  Small
blocks (u_c ~ 1,200) are launch- and latency-bound on a B200, so here every block gets its own
handle -- and with it its own CUDA stream -- and the Gauss-Newton loops advance in lock step: ONE CUDA graph
holds a step of every still-active block (``feba_batch``: the handles' streams fork from the first one inside the
capture and join it again), then one ``feba_sync`` each.  Kernels of different blocks overlap on the device; there
is no exchange between blocks (replicas only).
Results are bit-identical to running the blocks one at a time (each block's arithmetic is untouched).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np

from .lib import Batch, Handle
from .problem import Buildxhat, Problem


def adjust_batch(problems: Sequence[Problem], xhat0s: Optional[Sequence[np.ndarray]] = None,
                 want_residuals: bool = True, handles: Optional[List[Handle]] = None) -> List[dict]:
    """main.m:386-602 for every problem; returns one result dict per problem (as ``adjust``)."""
    own = handles is None
    if own:
        handles = [Handle(p) for p in problems]
    try:
        n = len(problems)
        if xhat0s is None:
            xhat0s = []
            for p in problems:
                err, x0, _ = Buildxhat(p)
                if err:
                    raise ValueError("Error building xhat")
                xhat0s.append(x0)
        for h, x0 in zip(handles, xhat0s):
            h.set_xhat(x0)
        deltasum = [100.0] * n                                  # main.m:407
        count = [0] * n
        trace: List[List[float]] = [[] for _ in range(n)]
        active = [i for i in range(n) if deltasum[i] > problems[i].settings.threshold]
        batch, batch_of = None, None
        while active:
            # one step of every active block: a single captured CUDA graph from the second step on (feba_batch);
            # the batch is rebuilt when blocks drop out (they converge after different numbers of iterations)
            if batch_of != active:
                if batch is not None:
                    batch.close()
                batch, batch_of = Batch([handles[i] for i in active]), list(active)
            batch.iterate_async()
            nxt = []
            for i in active:                                     # main.m:484-493 per block
                deltasum[i] = handles[i].sync()
                count[i] += 1
                trace[i].append(deltasum[i])
                s = problems[i].settings
                if deltasum[i] > s.threshold and count[i] < s.Iteration_Cap:
                    nxt.append(i)
            active = nxt
        if batch is not None:
            batch.close()
        out = []
        for i, h in enumerate(handles):
            res = h.residuals() if want_residuals else {}
            res.update(xhat=h.get_xhat(), iterations=count[i], deltasum=trace[i])
            out.append(res)
        return out
    finally:
        if own:
            for h in handles:
                h.close()
