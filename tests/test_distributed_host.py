"""N > 1 host logic on CPU: two gloo ranks each hold their partition (owned nodes + ghosts + one
overlap layer of cells); a forward halo exchange of the input vector followed by the local
operator on owned rows reproduces the global operator — the communication pattern pfx_comm_attach
/ exchange() use on the GPUs.  The local operator here is the oracle's assembled
matrix on the rank-local mesh (checker only)."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, ctype, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.solver import Oracle, Params
        from phasefieldx_b200 import _lib, mesh as M
        m = M.create_rectangle(None, [[0, 0], [1, 1]], [14, 11]) if ctype == "triangle" else M.create_box(None, [[0, 0, 0], [1, 1, 1]], [5, 4, 3])
        torch.set_num_threads(1)
        P = (Params(degradation="anisotropic", split_energy="spectral", k=1e-4) if ctype == "triangle"
             else Params(degradation="anisotropic", split_energy="deviatoric", k=1e-4))
        rng = np.random.default_rng(0)
        og = Oracle(m.geometry.x, m.cells, m.cell_type, P)
        d = og.d
        u, phi, v = rng.normal(size=og.nn * d) * 1e-3, rng.uniform(0, 1, og.nn), rng.normal(size=og.nn * d)
        ref = og.J_u(u, phi) @ v
        R = _lib.Partition(m.cell_type, m.geometry.x, m.cells, world).rank(rank)
        l2g, no = R["l2g"], R["n_owned"]
        g2l = np.full(og.nn, -1)
        g2l[l2g] = np.arange(len(l2g))
        ol = Oracle(m.geometry.x[l2g], g2l[m.cells[R["cell_gid"]]], m.cell_type, P)
        ldof = (l2g[:, None] * d + np.arange(d)).ravel()
        vl = v[ldof].reshape(-1, d).copy()
        vl[no:] = np.nan                                        # ghosts must come from the exchange
        reqs, bufs = [], []
        for k, nb in enumerate(R["neighbors"]):
            s = torch.from_numpy(vl[R["send_idx"][R["send_off"][k]:R["send_off"][k + 1]]].copy())
            reqs.append(dist.isend(s, int(nb)))
            rb = torch.empty((int(R["recv_off"][k + 1] - R["recv_off"][k]), d), dtype=torch.float64)
            bufs.append((k, rb))
            reqs.append(dist.irecv(rb, int(nb)))
        for r_ in reqs:
            r_.wait()
        for k, rb in bufs:
            vl[no + R["recv_off"][k]: no + R["recv_off"][k + 1]] = rb.numpy()
        assert not np.isnan(vl).any()
        yl = ol.J_u(u[ldof], phi[l2g]) @ vl.ravel()
        err = np.abs(yl[: no * d] - ref[ldof[: no * d]]).max() / np.abs(ref).max()
        # scalar reductions: primary cells only, all-reduced
        prim = R["cell_primary"].astype(bool)
        vol = torch.tensor([ol.t.W[prim].sum()], dtype=torch.float64)
        dist.all_reduce(vol)
        q.put((rank, err, float(vol), float(og.t.W.sum())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ctype", ["triangle", "tetrahedron"])
def test_two_rank_halo_apply_matches_global(ctype):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 200) + (0 if ctype == "triangle" else 1)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ctype, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    res = [q.get(timeout=10) for _ in range(2)]
    for rank, err, vol, vol_ref in res:
        assert err < 1e-12, (rank, err)
        assert abs(vol - vol_ref) < 1e-12 * vol_ref
