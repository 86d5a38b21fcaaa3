"""
ORACLE (test infrastructure, not product code) — Python restatement of the rank partition and
ghost maps built by phasefieldx_b200/csrc/pfx_blocks.cpp::build_partition, for the bit-exact
check "mesh partitioning/ghost maps" of BASELINE.json.

Rule set (dolfinx IndexMap conventions; the partitioner itself replaces SCOTCH/ParMETIS, which
are not available, by recursive coordinate bisection of the cell centroids):
  * cell_rank: RCB, parts split ⌊p/2⌋ : p−⌊p/2⌋, item counts proportional (round half up),
    axis of largest extent, ties broken by cell index;
  * node owner = lowest rank among adjacent cells;
  * rank-local nodes: owned (ascending global id), then ghosts sorted by (owner, global id);
  * rank-local cells: every cell touching an owned node (ascending global id); a cell is
    `primary` on the rank that owns its lowest-numbered node;
  * exchange plan: symmetric neighbour lists; receive ranges = ghost groups per owner; send lists
    in the receiver's ghost order.
"""
import numpy as np


def rcb(xyz, nparts):
    n = len(xyz)
    part = np.zeros(n, dtype=np.int32)

    def rec(ids, p, base):
        if p <= 1:
            part[ids] = base
            return
        pl = p // 2
        nl = (len(ids) * pl + p // 2) // p
        ext = xyz[ids].max(0) - xyz[ids].min(0)
        ax = 0
        for a in range(1, xyz.shape[1]):
            if ext[a] > ext[ax]:
                ax = a
        order = np.lexsort((ids, xyz[ids, ax]))
        srt = ids[order]
        rec(srt[:nl], pl, base)
        rec(srt[nl:], p - pl, base + pl)
    if nparts > 1:
        rec(np.arange(n), nparts, 0)
    return part


def partition(coords, cells, n_ranks, cell_rank=None):
    coords = np.asarray(coords, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    nn, nv = len(coords), cells.shape[1]
    if cell_rank is None:
        cen = np.zeros((len(cells), 3))
        for a in range(nv):                       # same accumulation order as the C++ (a = 0..nv-1)
            cen += coords[cells[:, a]] / nv
        cell_rank = rcb(cen, n_ranks)
    cell_rank = np.asarray(cell_rank, dtype=np.int32)
    owner = np.full(nn, np.iinfo(np.int32).max, dtype=np.int32)
    np.minimum.at(owner, cells.ravel(), np.repeat(cell_rank, nv))
    ranks = []
    for r in range(n_ranks):
        touch = (owner[cells] == r).any(axis=1)
        cell_gid = np.nonzero(touch)[0]
        low = cells[cell_gid].min(axis=1)
        primary = (owner[low] == r).astype(np.uint8)
        owned = np.nonzero(owner == r)[0]
        need = np.zeros(nn, dtype=bool)
        need[cells[cell_gid].ravel()] = True
        gh = np.nonzero(need & (owner != r))[0]
        gh = gh[np.lexsort((gh, owner[gh]))]
        ranks.append(dict(n_owned=len(owned), l2g=np.concatenate([owned, gh]), cell_gid=cell_gid,
                          ghost_owner=owner[gh].astype(np.int32), cell_primary=primary))
    for r in range(n_ranks):
        R = ranks[r]
        want_from = {q: ranks[r]["l2g"][R["n_owned"]:][R["ghost_owner"] == q] for q in range(n_ranks)}
        R["_want"] = want_from
    for r in range(n_ranks):
        R = ranks[r]
        nbrs = [q for q in range(n_ranks) if q != r and (len(R["_want"][q]) or len(ranks[q]["_want"][r]))]
        g2l = {int(g): i for i, g in enumerate(R["l2g"][: R["n_owned"]])}
        send_idx, send_off, recv_off = [], [0], [0]
        for q in nbrs:
            send_idx += [g2l[int(g)] for g in ranks[q]["_want"][r]]
            send_off.append(len(send_idx))
            recv_off.append(recv_off[-1] + len(R["_want"][q]))
        R["neighbors"] = np.array(nbrs, dtype=np.int32)
        R["send_idx"] = np.array(send_idx, dtype=np.int32)
        R["send_off"] = np.array(send_off, dtype=np.int64)
        R["recv_off"] = np.array(recv_off, dtype=np.int64)
    for R in ranks:
        R.pop("_want")
    return cell_rank, owner, ranks
