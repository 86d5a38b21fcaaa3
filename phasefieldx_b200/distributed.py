"""
One process per GPU: partition a mesh over the ranks of a `torch.distributed` (NCCL) job, build the
rank-local `Engine` (owned nodes + ghosts + one overlap layer of cells) and attach the
communication paths (NCCL communicator, then the NVLink peer path via CUDA IPC).

Replaces what dolfinx does implicitly for the reference under `mpirun` (mesh partitioning at
creation, IndexMap/Scatterer ghost updates, `comm.allreduce` in energy.py / errors_functions.py);
the partition itself is `pfx_partition_*` of the C-ABI (deterministic, bit-exact against
oracle/partition.py).  A dolfinx partition can be replayed through `cell_rank`.
"""
import os

import numpy as np

from . import _lib


class RankMesh:
    """Local view of the global mesh on one rank."""

    def __init__(self, coords, cells, cell_type, world, rank, cell_rank=None, partitioner=None):
        """partitioner: "rcb" (default: recursive coordinate bisection of the cell centroids) or "metis" (k-way graph
        partition of the cell dual graph, deterministic: every rank computes the same one); env PFX_PARTITIONER."""
        self.world, self.rank = world, rank
        partitioner = partitioner or os.environ.get("PFX_PARTITIONER", "rcb")
        if partitioner not in ("rcb", "metis"):
            raise ValueError("partitioner must be 'rcb' or 'metis'")
        if cell_rank is None and partitioner == "metis" and world > 1:
            cell_rank, self.edge_cut = _lib.graph_cell_rank(cell_type, len(coords), cells, world)
        self.partition = _lib.Partition(cell_type, coords, cells, world, cell_rank=cell_rank)
        R = self.partition.rank(rank)
        self.plan = R
        self.l2g = R["l2g"]
        self.n_owned = R["n_owned"]
        self.g2l = np.full(len(coords), -1, dtype=np.int64)
        self.g2l[self.l2g] = np.arange(len(self.l2g))
        self.coords = np.ascontiguousarray(coords[self.l2g])
        self.cells = self.g2l[np.asarray(cells)[R["cell_gid"]]].astype(np.int32)
        self.cell_primary = R["cell_primary"]
        self.n_global = len(coords)

    def piece(self, cell_type):
        """Rank-local mesh view for the parallel VTU writer (io_vtk.VTKFile with rank/world): owned + ghost
        nodes, local cells, global ids, ghost marks (nodes owned elsewhere; cells integrated elsewhere)."""
        import types
        ghost_pt = (np.arange(len(self.l2g)) >= self.n_owned).astype(np.uint8)
        return types.SimpleNamespace(geometry=types.SimpleNamespace(x=self.coords, dim=self.coords.shape[1]), cells=self.cells,
                                     cell_type=cell_type, num_nodes=len(self.coords), point_gid=self.l2g,
                                     cell_gid=self.plan["cell_gid"], point_ghost=ghost_pt,
                                     cell_ghost=(1 - np.asarray(self.cell_primary, dtype=np.uint8)))

    def local_dofs(self, dofs, bs):
        """global unrolled dofs -> (keep mask, local unrolled dofs) for the dofs whose node is held here"""
        dofs = np.asarray(dofs, dtype=np.int64)
        node, comp = dofs // bs, dofs % bs
        keep = self.g2l[node] >= 0
        return keep, (self.g2l[node[keep]] * bs + comp[keep]).astype(np.int32)

    def local_nodes(self, nodes):
        nodes = np.asarray(nodes, dtype=np.int64)
        keep = self.g2l[nodes] >= 0
        return keep, self.g2l[nodes[keep]].astype(np.int32)


def attach(engine, rmesh, dist):
    """NCCL communicator + (unless PFX_NO_P2P=1) the NVLink peer path."""
    uid = [_lib.comm_unique_id() if rmesh.rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    engine.comm_attach(rmesh.rank, rmesh.world, uid[0], rmesh.plan)
    if os.environ.get("PFX_NO_P2P", "0") != "1":
        handles = [None] * rmesh.world
        dist.all_gather_object(handles, engine.p2p_export())
        engine.p2p_import(handles, _lib.peer_ghost_start(rmesh.partition, rmesh.rank))


def gather_field(engine, rmesh, name, dist, device):
    """Global nodal field (caller's global node order) on every rank from the owned parts."""
    import torch
    nc = engine.dim if name in ("u", "u_old") else 1
    loc = engine.get_field(name).reshape(-1, nc)[: rmesh.n_owned]
    full = torch.zeros((rmesh.n_global, nc), dtype=torch.float64, device=device)
    full[torch.from_numpy(rmesh.l2g[: rmesh.n_owned]).to(device)] = torch.from_numpy(loc).to(device)
    dist.all_reduce(full)
    return full.cpu().numpy().reshape(-1) if nc == 1 else full.cpu().numpy()
