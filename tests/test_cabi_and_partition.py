"""The C-ABI library loads on a CPU-only box and exports every symbol include/pfx_b200.h
declares; context creation fails loudly without a GPU (no CPU fallback); the host-only
partition / ghost-map API is bit-exact against the Python restatement (oracle/partition.py)."""
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def test_library_exports_every_declared_symbol():
    from phasefieldx_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "pfx_b200.h")).read()
    declared = set(re.findall(r"\b(pfx_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.pfx_version()


def test_no_cpu_fallback():
    import torch
    from phasefieldx_b200 import _lib, mesh as M
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    m = M.create_rectangle(None, [[0, 0], [1, 1]], [4, 4])
    with pytest.raises(_lib.PfxError, match="no CUDA device"):
        _lib.Engine(m.geometry.x, m.cells, "triangle", 1.0, 1.0, 1.0, 1.0)


def test_struct_layout_matches_header():
    """ctypes mirrors of pfx_params / pfx_step_report have the C field order of the header."""
    from phasefieldx_b200 import _lib
    header = open(os.path.join(ROOT, "include", "pfx_b200.h")).read()
    for cname, mirror in (("pfx_params", _lib.PfxParams), ("pfx_ec_opts", _lib.PfxEcOpts), ("pfx_ec_report", _lib.PfxEcReport)):
        body = re.search(r"typedef struct \{([^}]*)\} %s;" % cname, header).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            typ, rest = decl.split(None, 1)
            names += [n.strip() for n in rest.split(",")]
        assert names == [f[0] for f in mirror._fields_], cname


@pytest.mark.parametrize("n_ranks", [2, 3, 8])
def test_partition_and_ghost_maps_bit_exact(n_ranks):
    from oracle.partition import partition
    from phasefieldx_b200 import _lib, mesh as M
    for m in (M.create_rectangle(None, [[0, 0], [1, 1]], [23, 17]), M.create_box(None, [[0, 0, 0], [2, 1, 1]], [7, 5, 4]),
              M.create_rectangle(None, [[0, 0], [2, 1]], [16, 9], M.CellType.quadrilateral)):
        P = _lib.Partition(m.cell_type, m.geometry.x, m.cells, n_ranks)
        cr, ow, ranks = partition(m.geometry.x, m.cells, n_ranks)
        assert np.array_equal(P.cell_rank(), cr) and np.array_equal(P.node_owner(), ow)
        seen_primary = np.zeros(m.num_cells, dtype=int)
        for r in range(n_ranks):
            A, B = P.rank(r), ranks[r]
            assert A["n_owned"] == B["n_owned"]
            for k in ("l2g", "cell_gid", "ghost_owner", "cell_primary", "neighbors", "send_off", "send_idx", "recv_off"):
                assert np.array_equal(A[k], B[k]), (r, k)
            seen_primary[A["cell_gid"][A["cell_primary"] == 1]] += 1
        assert np.all(seen_primary == 1)          # every cell integrated exactly once
        # replay of an external partition
        P2 = _lib.Partition(m.cell_type, m.geometry.x, m.cells, n_ranks, cell_rank=cr[::-1].copy())
        assert np.array_equal(P2.cell_rank(), cr[::-1])


@pytest.mark.parametrize("partitioner", ["rcb", "metis"])
def test_rank_mesh_local_views_cover_the_mesh(partitioner):
    """distributed.RankMesh: local meshes, Dirichlet dof maps and peer ghost offsets (host logic), for the coordinate
    bisection and for the METIS graph partition."""
    from phasefieldx_b200 import _lib, mesh as M
    from phasefieldx_b200.distributed import RankMesh
    m = M.create_rectangle(None, [[0, 0], [1, 1]], [12, 9])
    world = 3
    top = np.nonzero(np.isclose(m.geometry.x[:, 1], 1.0))[0]
    dofs = (top[:, None] * 2 + np.arange(2)).ravel()
    owned_seen = np.zeros(m.num_nodes, dtype=int)
    for r in range(world):
        rm = RankMesh(m.geometry.x, m.cells, m.cell_type, world, r, partitioner=partitioner)
        owned_seen[rm.l2g[: rm.n_owned]] += 1
        assert np.array_equal(rm.coords, m.geometry.x[rm.l2g])
        assert rm.cells.min() >= 0 and rm.cells.max() < len(rm.l2g)
        keep, ld = rm.local_dofs(dofs, 2)
        assert np.array_equal(rm.l2g[ld // 2] * 2 + ld % 2, dofs[keep])
        gs = _lib.peer_ghost_start(rm.partition, r)
        for j, q in enumerate(rm.plan["neighbors"]):
            Q = rm.partition.rank(int(q))
            n_send = int(rm.plan["send_off"][j + 1] - rm.plan["send_off"][j])
            sent_global = rm.l2g[rm.plan["send_idx"][rm.plan["send_off"][j]: rm.plan["send_off"][j + 1]]]
            assert np.array_equal(Q["l2g"][gs[j]: gs[j] + n_send], sent_global)     # lands on the right ghosts
    assert np.all(owned_seen == 1)


def test_unknown_degradation_function_is_rejected_before_any_device_work():
    """Only the degradation functions built on the GPU path are accepted;
    the binding refuses the others without needing a device."""
    from phasefieldx_b200 import _lib, mesh as M
    msh = M.create_rectangle(None, [[0, 0], [1, 1]], [2, 2])
    assert _lib.DEGRADATION == {"quadratic": 0, "borden": 1}
    for name in ("alessi", "sargado", "cubic"):
        with pytest.raises(NotImplementedError):
            _lib.Engine(msh.geometry.x, msh.cells, "triangle", 1.0, 1.0, 1.0, 0.1, degradation_function=name)


def test_degradation_function_check_replicates_the_reference_behaviour():
    """`check_degradation_function` (called first thing by the drop-in solve() modules): quadratic / borden pass;
    'sargado' raises the OverflowError the reference raises while building its forms (exp(100 * 100) in
    g_degradation_functions.py:120-124); 'alessi' is refused explicitly; unknown names give the reference's
    ValueError (:149-150)."""
    from phasefieldx_b200 import _lib
    _lib.check_degradation_function("quadratic")
    _lib.check_degradation_function("borden")
    with pytest.raises(OverflowError):
        _lib.check_degradation_function("sargado")
    with pytest.raises(NotImplementedError, match="alessi"):
        _lib.check_degradation_function("alessi")
    with pytest.raises(ValueError, match="Invalid degradation_type"):
        _lib.check_degradation_function("cubic")


@pytest.mark.parametrize("n_ranks", [2, 4, 8])
def test_metis_graph_partition_feeds_bit_exact_ghost_maps(n_ranks):
    """north_star 'graph-partitioned': pfx_partition_graph_cell_rank (METIS k-way on the cell dual graph) gives a
    balanced, deterministic cell partition with a facet cut no worse than coordinate bisection on a structured mesh;
    replayed through pfx_partition_create its owner / ghost / exchange maps are bit-exact against the oracle."""
    from oracle.partition import partition
    from phasefieldx_b200 import _lib, mesh as M
    for m in (M.create_rectangle(None, [[0, 0], [2, 1]], [40, 20]), M.create_box(None, [[0, 0, 0], [2, 1, 1]], [10, 6, 5]),
              M.create_rectangle(None, [[0, 0], [2, 1]], [24, 12], M.CellType.quadrilateral)):
        cr, cut = _lib.graph_cell_rank(m.cell_type, m.num_nodes, m.cells, n_ranks)
        cr2, cut2 = _lib.graph_cell_rank(m.cell_type, m.num_nodes, m.cells, n_ranks)
        assert np.array_equal(cr, cr2) and cut == cut2                               # deterministic
        counts = np.bincount(cr, minlength=n_ranks)
        assert counts.min() > 0 and counts.max() <= 1.1 * m.num_cells / n_ranks + 2   # METIS default imbalance 1.03
        # facet cut recomputed from the partition == what METIS reports
        nv = m.cells.shape[1]
        if m.cell_type == "quadrilateral":
            facets = [(0, 1), (1, 3), (3, 2), (2, 0)]
        else:
            facets = [tuple(a for a in range(nv) if a != f) for f in range(nv)]
        seen = {}
        mycut = 0
        for c, cell in enumerate(m.cells):
            for f in facets:
                key = tuple(sorted(int(cell[a]) for a in f))
                if key in seen:
                    mycut += int(cr[seen[key]] != cr[c])
                else:
                    seen[key] = c
        assert mycut == cut
        P = _lib.Partition(m.cell_type, m.geometry.x, m.cells, n_ranks, cell_rank=cr)
        cr_o, ow, ranks = partition(m.geometry.x, m.cells, n_ranks, cell_rank=cr)
        assert np.array_equal(P.cell_rank(), cr) and np.array_equal(P.node_owner(), ow)
        for r in range(n_ranks):
            A, B = P.rank(r), ranks[r]
            for k in ("l2g", "cell_gid", "ghost_owner", "cell_primary", "neighbors", "send_off", "send_idx", "recv_off"):
                assert np.array_equal(A[k], B[k]), (r, k)
