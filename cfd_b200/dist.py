"""One process per GPU: build this rank's sub-domain solver and wire its NCCL communicator.

torch.distributed is used for plumbing only (rendezvous, broadcasting the ncclUniqueId, gathering
results in tests); the data path — ghost refresh and all-reduces — is NCCL called from inside
libcfdb200.so on the solver's own stream.
"""
from __future__ import annotations

import numpy as np

from . import partition
from .solver import NSComp2D


def make_rank_solver(lc_or_window, rank, nranks, device, dist=None, use_gcl=0):
    """lc_or_window: global LoadedCase, or the tuple returned by partition.square_window()."""
    if isinstance(lc_or_window, tuple):
        part = partition.build_local(lc_or_window[0], nranks, rank, *lc_or_window[1:])
    else:
        part = partition.build_local(lc_or_window, nranks, rank)
    # the communicator comes before cfdb_init: DERIV's HMIN is a global minimum (subrutinas.f90:124)
    g = NSComp2D(part.lc, device=device, use_gcl=use_gcl, init=False)
    g.attach_partition(part)
    if nranks > 1:
        if dist is None:
            raise ValueError("nranks > 1 needs an initialised torch.distributed module")
        box = [NSComp2D.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        g.comm_init(box[0], rank, nranks)
    g.init()
    return g, part


def gather_owned(g, part, name, width, dist, npoin_global):
    """Assemble a global nodal array from the owned entries of every rank (tests)."""
    a = g.get(name).reshape(-1, width)[: part.n_owned]
    box = [None] * part.nranks
    dist.all_gather_object(box, (part.node_gid[: part.n_owned], a))
    out = np.zeros((npoin_global, width))
    for gid, val in box:
        out[gid] = val
    return out
