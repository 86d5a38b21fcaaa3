// pfx_blocks.cpp — see pfx_blocks.h.  Pure host C++ (also compiled into the CPU test helper).
#include "pfx_blocks.h"

#include <algorithm>
#include <cmath>
#include <numeric>

namespace pfx {

namespace {

// Recursive coordinate bisection into `nparts` parts of (almost) equal size.  Deterministic:
// ties on the coordinate are broken by the item index.
struct Rcb {
    const double* xyz;   // [n,3]
    int dim;
    std::vector<int32_t>& ids;
    std::vector<int32_t>& part;
    bool even = false;   // round split points to even counts (block starts stay 16-byte aligned for 8-byte fields)
    void run(int64_t lo, int64_t hi, int nparts, int base) {
        if (nparts <= 1) {
            for (int64_t i = lo; i < hi; ++i) part[ids[i]] = base;
            return;
        }
        int pl = nparts / 2;
        int64_t n = hi - lo;
        int64_t nl = (n * pl + nparts / 2) / nparts;
        if (even && (nl & 1) && nl < n) ++nl;
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int64_t i = lo; i < hi; ++i)
            for (int a = 0; a < dim; ++a) {
                double v = xyz[3 * (int64_t)ids[i] + a];
                mn[a] = std::min(mn[a], v);
                mx[a] = std::max(mx[a], v);
            }
        int ax = 0;
        for (int a = 1; a < dim; ++a)
            if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
        const double* c = xyz;
        auto cmp = [c, ax](int32_t p, int32_t q) {
            double vp = c[3 * (int64_t)p + ax], vq = c[3 * (int64_t)q + ax];
            return vp < vq || (vp == vq && p < q);
        };
        std::nth_element(ids.begin() + lo, ids.begin() + lo + nl, ids.begin() + hi, cmp);
        run(lo, lo + nl, pl, base);
        run(lo + nl, hi, nparts - pl, base + pl);
    }
};

}  // namespace

std::string build_blocks(int nv, int dim, int64_t n_nodes, int64_t n_owned, int64_t n_cells, const double* coords,
                         const int32_t* cells, const uint8_t* cell_primary, int max_own, BlockMesh& bm) {
    if (n_owned <= 0 || n_owned > n_nodes) return "bad n_owned";
    if (max_own < 8) return "block size too small";
    bm = BlockMesh();
    bm.nv = nv; bm.dim = dim; bm.n_nodes = n_nodes; bm.n_owned = n_owned; bm.n_cells = n_cells;

    // ---- 1. blocks of owned nodes (retry with more blocks until every block fits) ---------
    std::vector<int32_t> part(n_owned), ids(n_owned);
    int nb = (int)((n_owned + max_own - 1) / max_own);
    for (;;) {
        std::iota(ids.begin(), ids.end(), 0);
        Rcb r{coords, dim, ids, part, true};
        r.run(0, n_owned, nb, 0);
        std::vector<int32_t> cnt(nb, 0);
        for (int64_t i = 0; i < n_owned; ++i) cnt[part[i]]++;
        int mx = *std::max_element(cnt.begin(), cnt.end());
        if (mx <= max_own) break;
        nb += std::max(1, nb / 64);
    }
    bm.nb = nb;

    // ---- 2. renumber: block-contiguous owned nodes (ascending old id inside a block) ------
    bm.node0.assign(nb + 1, 0);
    for (int64_t i = 0; i < n_owned; ++i) bm.node0[part[i] + 1]++;
    for (int b = 0; b < nb; ++b) bm.node0[b + 1] += bm.node0[b];
    bm.perm.resize(n_nodes); bm.iperm.resize(n_nodes);
    {
        std::vector<int32_t> cur(bm.node0.begin(), bm.node0.end() - 1);
        for (int64_t i = 0; i < n_owned; ++i) {
            int32_t nw = cur[part[i]]++;
            bm.perm[nw] = (int32_t)i; bm.iperm[i] = nw;
        }
        for (int64_t i = n_owned; i < n_nodes; ++i) { bm.perm[i] = (int32_t)i; bm.iperm[i] = (int32_t)i; }
    }
    auto block_of_new = [&](int32_t nw) -> int { return nw < n_owned ? part[bm.perm[nw]] : -1; };

    // ---- 3. element lists per block: own cells first (the block holds the cell's lowest owned node; rank-primary
    //         ones lead), then the other cells touching an owned node ---------------------------------------------
    std::vector<int32_t> cnt_p(nb, 0), cnt_l(nb, 0), cnt_s(nb, 0);
    auto visit = [&](auto&& fn) {
        for (int64_t c = 0; c < n_cells; ++c) {
            int blk[8]; int nblk = 0; int32_t low = INT32_MAX; int lowb = -1;
            for (int a = 0; a < nv; ++a) {
                int32_t nw = bm.iperm[cells[c * nv + a]];
                int b = block_of_new(nw);
                if (b < 0) continue;
                if (nw < low) { low = nw; lowb = b; }
                bool seen = false;
                for (int k = 0; k < nblk; ++k) seen |= (blk[k] == b);
                if (!seen) blk[nblk++] = b;
            }
            bool prim = cell_primary ? cell_primary[c] != 0 : true;
            for (int k = 0; k < nblk; ++k) fn(c, blk[k], blk[k] == lowb ? (prim ? 0 : 1) : 2);
        }
    };
    visit([&](int64_t, int b, int grp) { (grp == 0 ? cnt_p : (grp == 1 ? cnt_l : cnt_s))[b]++; });
    bm.elem_off.assign(nb + 1, 0);
    bm.nprim.assign(nb, 0);
    bm.nlow.assign(nb, 0);
    bm.own_c0.assign(nb + 1, 0);
    for (int b = 0; b < nb; ++b) {
        bm.elem_off[b + 1] = bm.elem_off[b] + cnt_p[b] + cnt_l[b] + cnt_s[b];
        bm.nprim[b] = cnt_p[b];
        bm.nlow[b] = cnt_p[b] + cnt_l[b];
        bm.own_c0[b + 1] = bm.own_c0[b] + bm.nlow[b];
    }
    bm.n_own_cells = bm.own_c0[nb];
    bm.total_elems = bm.elem_off[nb];
    bm.elem_gid.resize(bm.total_elems);
    {
        std::vector<int32_t> cur[3] = {std::vector<int32_t>(nb), std::vector<int32_t>(nb), std::vector<int32_t>(nb)};
        for (int b = 0; b < nb; ++b) {
            cur[0][b] = bm.elem_off[b]; cur[1][b] = cur[0][b] + cnt_p[b]; cur[2][b] = cur[1][b] + cnt_l[b];
        }
        visit([&](int64_t c, int b, int grp) { bm.elem_gid[cur[grp][b]++] = (int32_t)c; });
    }

    // ---- 4. halo lists + local connectivity ------------------------------------------------
    // 4 (8 for hexahedra) 16-bit block-local vertex ids per listed cell; incidence codes (cell << vb | vertex)
    const int ls = nv > 4 ? 8 : 4, vb = nv > 4 ? 3 : 2;
    bm.lstride = ls; bm.vbits = vb;
    bm.lconn.assign((size_t)ls * bm.total_elems, 0);
    bm.halo_off.assign(nb + 1, 0);
    std::vector<int32_t> stamp(n_nodes, -1), loc(n_nodes, 0), tmp;
    for (int b = 0; b < nb; ++b) {
        int32_t n0 = bm.node0[b], n1 = bm.node0[b + 1];
        tmp.clear();
        for (int32_t e = bm.elem_off[b]; e < bm.elem_off[b + 1]; ++e)
            for (int a = 0; a < nv; ++a) {
                int32_t nw = bm.iperm[cells[(int64_t)bm.elem_gid[e] * nv + a]];
                if (nw >= n0 && nw < n1) continue;
                if (stamp[nw] != b) { stamp[nw] = b; tmp.push_back(nw); }
            }
        std::sort(tmp.begin(), tmp.end());
        int n_own = n1 - n0;
        for (size_t k = 0; k < tmp.size(); ++k) loc[tmp[k]] = n_own + (int32_t)k;
        bm.halo.insert(bm.halo.end(), tmp.begin(), tmp.end());
        bm.halo_off[b + 1] = (int32_t)bm.halo.size();
        int n_loc = n_own + (int)tmp.size();
        if (n_loc > 65535) return "block has more than 65535 local nodes";
        int ne = bm.elem_off[b + 1] - bm.elem_off[b];
        if (ne > (0xFFFF >> vb) - 1) return "block has too many elements for its 16-bit incidence codes; lower block_nodes";
        bm.max_own = std::max(bm.max_own, n_own);
        bm.max_loc = std::max(bm.max_loc, n_loc);
        bm.max_elems = std::max(bm.max_elems, ne);
        for (int32_t e = bm.elem_off[b]; e < bm.elem_off[b + 1]; ++e)
            for (int a = 0; a < nv; ++a) {
                int32_t nw = bm.iperm[cells[(int64_t)bm.elem_gid[e] * nv + a]];
                bm.lconn[ls * (int64_t)e + a] = (uint16_t)((nw >= n0 && nw < n1) ? nw - n0 : loc[nw]);
            }
    }

    // ---- 5. node -> (element, vertex) incidence of owned nodes ------------------------------
    bm.inc_ptr.assign(n_owned + 1, 0);
    for (int b = 0; b < nb; ++b) {
        int n_own = bm.node0[b + 1] - bm.node0[b];
        for (int32_t e = bm.elem_off[b]; e < bm.elem_off[b + 1]; ++e)
            for (int a = 0; a < nv; ++a) {
                int l = bm.lconn[ls * (int64_t)e + a];
                if (l < n_own) bm.inc_ptr[bm.node0[b] + l + 1]++;
            }
    }
    // pad every block's list to a multiple of 8 entries (16-byte vector staging in the kernels);
    // the padding (sentinel 0xFFFF = element 16383, never valid) hangs off the block's last node
    {
        int64_t run = 0;
        for (int b = 0; b < nb; ++b) {
            for (int32_t i = bm.node0[b]; i < bm.node0[b + 1]; ++i) run += bm.inc_ptr[i + 1];
            int pad = (int)((8 - run % 8) % 8);
            if (bm.node0[b + 1] > bm.node0[b]) bm.inc_ptr[bm.node0[b + 1]] += pad;
            run += pad;
        }
    }
    for (int64_t i = 0; i < n_owned; ++i) bm.inc_ptr[i + 1] += bm.inc_ptr[i];
    bm.inc.assign(bm.inc_ptr[n_owned], (uint16_t)0xFFFF);
    {
        std::vector<int32_t> cur(bm.inc_ptr.begin(), bm.inc_ptr.end() - 1);
        for (int b = 0; b < nb; ++b) {
            int n_own = bm.node0[b + 1] - bm.node0[b];
            for (int32_t e = bm.elem_off[b]; e < bm.elem_off[b + 1]; ++e)
                for (int a = 0; a < nv; ++a) {
                    int l = bm.lconn[ls * (int64_t)e + a];
                    if (l < n_own) bm.inc[cur[bm.node0[b] + l]++] = (uint16_t)(((e - bm.elem_off[b]) << vb) | a);
                }
        }
    }
    // ---- 6. packed per-block metadata and (triangles) fixed-width incidence ------------------
    bm.meta.resize(8 * (size_t)nb);
    for (int b = 0; b < nb; ++b) {
        int32_t* m = &bm.meta[8 * (size_t)b];
        m[0] = bm.node0[b]; m[1] = bm.node0[b + 1] - bm.node0[b];
        m[2] = bm.halo_off[b]; m[3] = bm.halo_off[b + 1] - bm.halo_off[b];
        m[4] = bm.elem_off[b]; m[5] = bm.elem_off[b + 1] - bm.elem_off[b];
        m[6] = bm.inc_ptr[bm.node0[b]]; m[7] = bm.inc_ptr[bm.node0[b + 1]] - m[6];
    }
    if (nv == 3) {
        bm.inc8.assign(8 * (size_t)n_owned + 8, (uint16_t)0xFFFF);
        for (int64_t i = 0; i < n_owned; ++i) {
            int32_t k0 = bm.inc_ptr[i], k1 = bm.inc_ptr[i + 1];
            while (k1 > k0 && bm.inc[k1 - 1] == 0xFFFF) --k1;          // strip block padding
            int cnt = k1 - k0;
            for (int k = 0; k < std::min(cnt, 8); ++k) bm.inc8[8 * i + k] = bm.inc[k0 + k];
            if (cnt > 8) bm.inc8[8 * i + 7] = 0xFFFE;                   // entries 7.. continue in inc[]
        }
    }
    return "";
}

// =============================================================================== graph partition
// METIS k-way partition of the cell dual graph (cells adjacent when they share a facet) — the role of dolfinx's graph
// partitioner at mesh creation (dolfinx.mesh.create_mesh -> ParMETIS / SCOTCH; third-party, not in the reference tree).
// METIS 5 from the CUDA toolkit's libmetis_static.a (idx_t = int64, real_t = float), default options: deterministic.
extern "C" int METIS_PartGraphKway(int64_t* nvtxs, int64_t* ncon, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* vsize,
                                   int64_t* adjwgt, int64_t* nparts, float* tpwgts, float* ubvec, int64_t* options, int64_t* objval,
                                   int64_t* part);
extern "C" int METIS_SetDefaultOptions(int64_t* options);

std::string metis_cell_rank(int nv, int dim, int64_t n_nodes, int64_t n_cells, const int32_t* cells, int n_ranks, int32_t* cell_rank,
                            int64_t* edge_cut) {
    if (n_ranks < 1) return "n_ranks < 1";
    if (edge_cut) *edge_cut = 0;
    if (n_ranks == 1) { std::fill(cell_rank, cell_rank + n_cells, 0); return ""; }
    // facets: simplices — every vertex subset of size nv-1; quadrilaterals (tensor vertex order) — the four edges
    const int nf = nv, fv = (nv == 4 && dim == 2) ? 2 : nv - 1;
    static const int quad_edges[4][2] = {{0, 1}, {1, 3}, {3, 2}, {2, 0}};
    struct Facet { int32_t v[3]; int32_t cell; };
    std::vector<Facet> F((size_t)n_cells * nf);
    for (int64_t c = 0; c < n_cells; ++c)
        for (int f = 0; f < nf; ++f) {
            Facet& q = F[(size_t)c * nf + f];
            q.v[0] = q.v[1] = q.v[2] = -1;
            if (fv == 2 && nv == 4) { q.v[0] = cells[c * nv + quad_edges[f][0]]; q.v[1] = cells[c * nv + quad_edges[f][1]]; }
            else { int k = 0; for (int a = 0; a < nv; ++a) if (a != f) q.v[k++] = cells[c * nv + a]; }
            std::sort(q.v, q.v + fv);
            q.cell = (int32_t)c;
        }
    std::sort(F.begin(), F.end(), [](const Facet& a, const Facet& b) {
        if (a.v[0] != b.v[0]) return a.v[0] < b.v[0];
        if (a.v[1] != b.v[1]) return a.v[1] < b.v[1];
        if (a.v[2] != b.v[2]) return a.v[2] < b.v[2];
        return a.cell < b.cell;
    });
    std::vector<int64_t> xadj(n_cells + 1, 0);
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (size_t i = 0; i + 1 < F.size(); ++i)
        if (F[i].v[0] == F[i + 1].v[0] && F[i].v[1] == F[i + 1].v[1] && F[i].v[2] == F[i + 1].v[2]) {
            pairs.emplace_back(F[i].cell, F[i + 1].cell);
            ++xadj[F[i].cell + 1]; ++xadj[F[i + 1].cell + 1];
        }
    std::vector<Facet>().swap(F);
    for (int64_t c = 0; c < n_cells; ++c) xadj[c + 1] += xadj[c];
    std::vector<int64_t> adj(xadj[n_cells]), fill(xadj.begin(), xadj.end() - 1);
    for (auto& p : pairs) { adj[fill[p.first]++] = p.second; adj[fill[p.second]++] = p.first; }
    for (int64_t c = 0; c < n_cells; ++c) std::sort(adj.begin() + xadj[c], adj.begin() + xadj[c + 1]);
    int64_t nvtx = n_cells, ncon = 1, nparts = n_ranks, objval = 0, options[64];
    METIS_SetDefaultOptions(options);
    std::vector<int64_t> part(n_cells, 0);
    int rc = METIS_PartGraphKway(&nvtx, &ncon, xadj.data(), adj.data(), nullptr, nullptr, nullptr, &nparts, nullptr, nullptr, options,
                                 &objval, part.data());
    if (rc != 1) return "METIS_PartGraphKway failed (status " + std::to_string(rc) + ")";
    for (int64_t c = 0; c < n_cells; ++c) cell_rank[c] = (int32_t)part[c];
    if (edge_cut) *edge_cut = objval;
    (void)n_nodes;
    return "";
}

// =============================================================================== partition
std::string build_partition(int nv, int64_t n_nodes, int64_t n_cells, const double* coords, const int32_t* cells,
                            int n_ranks, const int32_t* cell_rank_in, Partition& P) {
    if (n_ranks < 1) return "n_ranks < 1";
    P = Partition();
    P.n_ranks = n_ranks; P.n_nodes = n_nodes; P.n_cells = n_cells;
    P.cell_rank.assign(n_cells, 0);
    if (cell_rank_in) {
        for (int64_t c = 0; c < n_cells; ++c) {
            if (cell_rank_in[c] < 0 || cell_rank_in[c] >= n_ranks) return "cell_rank out of range";
            P.cell_rank[c] = cell_rank_in[c];
        }
    } else if (n_ranks > 1) {
        std::vector<double> cen(3 * n_cells, 0.0);
        for (int64_t c = 0; c < n_cells; ++c)
            for (int a = 0; a < nv; ++a)
                for (int i = 0; i < 3; ++i) cen[3 * c + i] += coords[3 * (int64_t)cells[c * nv + a] + i] / nv;
        std::vector<int32_t> ids(n_cells);
        std::iota(ids.begin(), ids.end(), 0);
        Rcb r{cen.data(), 3, ids, P.cell_rank};
        r.run(0, n_cells, n_ranks, 0);
    }
    // node owner = lowest adjacent cell rank
    P.node_owner.assign(n_nodes, INT32_MAX);
    for (int64_t c = 0; c < n_cells; ++c)
        for (int a = 0; a < nv; ++a) {
            int32_t& o = P.node_owner[cells[c * nv + a]];
            o = std::min(o, P.cell_rank[c]);
        }
    for (int64_t i = 0; i < n_nodes; ++i)
        if (P.node_owner[i] == INT32_MAX) return "mesh has a node without cells";

    P.ranks.assign(n_ranks, RankPart());
    std::vector<int64_t> g2l(n_nodes);
    for (int r = 0; r < n_ranks; ++r) {
        RankPart& R = P.ranks[r];
        // local cells = every cell touching a node owned by r (ascending global id)
        std::vector<uint8_t> need(n_nodes, 0);
        for (int64_t c = 0; c < n_cells; ++c) {
            bool touch = false;
            for (int a = 0; a < nv; ++a) touch |= (P.node_owner[cells[c * nv + a]] == r);
            if (!touch) continue;
            R.cell_gid.push_back(c);
            for (int a = 0; a < nv; ++a) need[cells[c * nv + a]] = 1;
            // primary: the rank owning the cell's lowest-numbered node integrates it
            int32_t low = cells[c * nv];
            for (int a = 1; a < nv; ++a) low = std::min(low, cells[c * nv + a]);
            R.cell_primary.push_back(P.node_owner[low] == r ? 1 : 0);
        }
        for (int64_t i = 0; i < n_nodes; ++i)
            if (P.node_owner[i] == r) R.l2g.push_back(i);
        R.n_owned = (int64_t)R.l2g.size();
        std::vector<std::pair<int32_t, int64_t>> gh;
        for (int64_t i = 0; i < n_nodes; ++i)
            if (need[i] && P.node_owner[i] != r) gh.push_back({P.node_owner[i], i});
        std::sort(gh.begin(), gh.end());
        for (auto& g : gh) { R.l2g.push_back(g.second); R.ghost_owner.push_back(g.first); }
        // receive plan: ghost ranges per owner
        R.recv_off.push_back(0);
        for (size_t k = 0; k < gh.size(); ++k) {
            if (R.neighbors.empty() || R.neighbors.back() != gh[k].first) {
                if (!R.neighbors.empty()) R.recv_off.push_back((int64_t)k);
                R.neighbors.push_back(gh[k].first);
            }
        }
        if (!R.neighbors.empty()) R.recv_off.push_back((int64_t)gh.size());
    }
    // send plan: what neighbour q receives from r, in q's ghost order.  Neighbour sets are made
    // symmetric (a rank may send to a rank it receives nothing from).
    std::vector<std::vector<std::vector<int64_t>>> want(n_ranks, std::vector<std::vector<int64_t>>(n_ranks));
    for (int q = 0; q < n_ranks; ++q) {
        RankPart& Q = P.ranks[q];
        for (size_t k = 0; k < Q.ghost_owner.size(); ++k) want[Q.ghost_owner[k]][q].push_back(Q.l2g[Q.n_owned + k]);
    }
    for (int r = 0; r < n_ranks; ++r) {
        RankPart& R = P.ranks[r];
        std::vector<int32_t> nb;
        for (int q = 0; q < n_ranks; ++q)
            if (q != r && (!want[r][q].empty() || !want[q][r].empty())) nb.push_back(q);
        // rebuild recv offsets on the symmetric neighbour list
        std::vector<int64_t> roff(1, 0);
        for (int q : nb) roff.push_back(roff.back() + (int64_t)want[q][r].size());
        R.neighbors = nb; R.recv_off = roff;
        for (int64_t l = 0; l < R.n_owned; ++l) g2l[R.l2g[l]] = l;
        R.send_off.assign(1, 0);
        for (int q : nb) {
            for (int64_t g : want[r][q]) R.send_idx.push_back((int32_t)g2l[g]);
            R.send_off.push_back((int64_t)R.send_idx.size());
        }
    }
    return "";
}

// =============================================================================== node graph (sliced ELL)
std::string build_node_graph(const BlockMesh& bm, int64_t n_cells, const int32_t* cells, NodeGraph& g) {
    const int nv = bm.nv;
    if (nv != 4 && nv != 3) return "node graph: cells of three or four nodes only (triangles, quadrilaterals, tetrahedra)";
    g = NodeGraph();
    const int64_t n_rows = bm.n_owned;
    g.n_rows = n_rows;
    // nodes of every local cell in the compact own-cell order
    g.cell_nodes.assign((size_t)bm.n_own_cells * 4, 0);
    for (int b = 0; b < bm.nb; ++b)
        for (int j = 0; j < bm.nlow[b]; ++j) {
            const int64_t c = bm.elem_gid[bm.elem_off[b] + j];
            for (int a = 0; a < nv; ++a) g.cell_nodes[4 * ((size_t)bm.own_c0[b] + j) + a] = bm.iperm[cells[c * nv + a]];
        }
    (void)n_cells;
    // node -> incident cells
    g.nc_ptr.assign(n_rows + 1, 0);
    for (int64_t c = 0; c < bm.n_own_cells; ++c)
        for (int a = 0; a < nv; ++a) {
            const int32_t n = g.cell_nodes[4 * c + a];
            if (n < n_rows) g.nc_ptr[n + 1]++;
        }
    for (int64_t i = 0; i < n_rows; ++i) g.nc_ptr[i + 1] += g.nc_ptr[i];
    if ((int64_t)bm.n_own_cells >= ((int64_t)1 << 29)) return "node graph: too many cells";
    g.nc_code.resize(g.nc_ptr[n_rows]);
    g.nc_kpos.resize(g.nc_ptr[n_rows]);
    {
        std::vector<int32_t> cur(g.nc_ptr.begin(), g.nc_ptr.end() - 1);
        for (int64_t c = 0; c < bm.n_own_cells; ++c)
            for (int a = 0; a < nv; ++a) {
                const int32_t n = g.cell_nodes[4 * c + a];
                if (n < n_rows) g.nc_code[cur[n]++] = (int32_t)((c << 2) | a);
            }
    }
    // rows: sorted distinct nodes of the incident cells (the node itself included)
    g.n_slices = (n_rows + kSlice - 1) / kSlice;
    g.slice_k0.assign(g.n_slices + 1, 0);
    std::vector<int32_t> row_ptr(n_rows + 1, 0), row_col, tmp;
    row_col.reserve((size_t)n_rows * 16);
    for (int64_t i = 0; i < n_rows; ++i) {
        tmp.clear();
        tmp.push_back((int32_t)i);
        for (int32_t q = g.nc_ptr[i]; q < g.nc_ptr[i + 1]; ++q) {
            const int64_t c = g.nc_code[q] >> 2;
            for (int a = 0; a < nv; ++a) tmp.push_back(g.cell_nodes[4 * c + a]);
        }
        std::sort(tmp.begin(), tmp.end());
        tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
        if (tmp.size() > 255) return "node graph: a node has more than 255 neighbours";
        row_col.insert(row_col.end(), tmp.begin(), tmp.end());
        row_ptr[i + 1] = (int32_t)row_col.size();
        g.max_k = std::max(g.max_k, (int)tmp.size());
    }
    for (int64_t s = 0; s < g.n_slices; ++s) {
        int mk = 0;
        for (int64_t i = s * kSlice; i < std::min<int64_t>(n_rows, (s + 1) * kSlice); ++i) mk = std::max(mk, row_ptr[i + 1] - row_ptr[i]);
        g.slice_k0[s + 1] = g.slice_k0[s] + mk;
    }
    g.total_k = g.slice_k0[g.n_slices];
    if (g.total_k * kSlice >= ((int64_t)1 << 31)) return "node graph: too many entries";
    g.col.resize((size_t)g.total_k * kSlice);
    for (int64_t s = 0; s < g.n_slices; ++s)
        for (int lane = 0; lane < kSlice; ++lane) {
            const int64_t i = s * kSlice + lane;
            const int len = g.slice_k0[s + 1] - g.slice_k0[s];
            for (int k = 0; k < len; ++k) {
                int32_t cj = (int32_t)std::min<int64_t>(i, n_rows - 1);          // padding: the row itself (value 0)
                if (i < n_rows && k < row_ptr[i + 1] - row_ptr[i]) cj = row_col[row_ptr[i] + k];
                g.col[((size_t)g.slice_k0[s] + k) * kSlice + lane] = cj;
            }
        }
    // entry positions of the four nodes of every incident cell inside the row
    for (int64_t i = 0; i < n_rows; ++i) {
        const int32_t* rc = &row_col[row_ptr[i]];
        const int len = row_ptr[i + 1] - row_ptr[i];
        for (int32_t q = g.nc_ptr[i]; q < g.nc_ptr[i + 1]; ++q) {
            const int64_t c = g.nc_code[q] >> 2;
            uint32_t kp = 0;
            for (int a = 0; a < nv; ++a) {
                const int32_t* f = std::lower_bound(rc, rc + len, g.cell_nodes[4 * c + a]);
                kp |= (uint32_t)(f - rc) << (8 * a);
            }
            g.nc_kpos[q] = kp;
        }
    }
    return "";
}

std::string build_aggregates(const BlockMesh& bm, int64_t n_cells, const double* coords, const int32_t* cells, int g,
                             int reg_aggs, int chunk_nodes, AggregateSet& out) {
    if (g < 1 || reg_aggs < 1 || chunk_nodes < 2 || bm.nb < 1) return "build_aggregates: bad arguments";
    const int D = bm.dim, nv = bm.nv;
    const int n_agg = (bm.nb + g - 1) / g;
    out = AggregateSet();
    out.n_agg = n_agg;
    out.n_reg = std::max(1, (n_agg + reg_aggs - 1) / reg_aggs);
    out.reg_agg0.resize(out.n_reg + 1);
    out.agg_reg.resize(n_agg);
    for (int r = 0; r <= out.n_reg; ++r) out.reg_agg0[r] = (int32_t)((int64_t)n_agg * r / out.n_reg);
    for (int r = 0; r < out.n_reg; ++r)
        for (int a = out.reg_agg0[r]; a < out.reg_agg0[r + 1]; ++a) out.agg_reg[a] = r;
    out.agg_node0.resize(n_agg + 1);
    int max_nodes = 0;
    for (int a = 0; a <= n_agg; ++a) out.agg_node0[a] = bm.node0[std::min(bm.nb, a * g)];
    for (int a = 0; a < n_agg; ++a) max_nodes = std::max(max_nodes, out.agg_node0[a + 1] - out.agg_node0[a]);
    out.S = std::max(1, (max_nodes + chunk_nodes - 1) / chunk_nodes);
    out.n_chunks = n_agg * out.S;
    out.chunk_node0.resize(out.n_chunks + 1);
    for (int a = 0; a < n_agg; ++a)
        for (int s = 0; s < out.S; ++s)
            out.chunk_node0[a * out.S + s] =
                out.agg_node0[a] + ((int32_t)((int64_t)(out.agg_node0[a + 1] - out.agg_node0[a]) * s / out.S) & ~1);
    out.chunk_node0[out.n_chunks] = out.agg_node0[n_agg];
    // geometry of the modes: coordinates relative to the aggregate centroid, scaled by its half extent
    const int64_t n_owned = bm.n_owned;
    std::vector<int32_t> node_agg(n_owned);
    out.rel.resize((size_t)n_owned * D);
    out.cen.assign((size_t)n_agg * 3, 0.0);
    out.half.assign(n_agg, 1.0);
    for (int a = 0; a < n_agg; ++a) {
        double cen[3] = {0, 0, 0}, lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        const int n0 = out.agg_node0[a], n1 = out.agg_node0[a + 1];
        for (int i = n0; i < n1; ++i)
            for (int k = 0; k < D; ++k) {
                double x = coords[3 * (int64_t)bm.perm[i] + k];
                cen[k] += x; lo[k] = std::min(lo[k], x); hi[k] = std::max(hi[k], x);
            }
        double size = 0;
        for (int k = 0; k < D; ++k) { cen[k] /= std::max(1, n1 - n0); size = std::max(size, 0.5 * (hi[k] - lo[k])); }
        const double sc = size > 0 ? 1.0 / size : 1.0;
        for (int k = 0; k < D; ++k) out.cen[(size_t)a * 3 + k] = cen[k];
        out.half[a] = size > 0 ? size : 1.0;
        for (int i = n0; i < n1; ++i) {
            node_agg[i] = a;
            for (int k = 0; k < D; ++k) out.rel[(size_t)i * D + k] = (float)((coords[3 * (int64_t)bm.perm[i] + k] - cen[k]) * sc);
        }
    }
    // aggregate graph
    std::vector<std::vector<int32_t>> adj(n_agg);
    for (int64_t e = 0; e < n_cells; ++e) {
        int32_t ag[8]; int na = 0;
        for (int k = 0; k < nv; ++k) {
            int32_t nid = bm.iperm[cells[e * nv + k]];
            if (nid >= n_owned) continue;
            int32_t a = node_agg[nid];
            bool seen = false;
            for (int q = 0; q < na; ++q) seen |= (ag[q] == a);
            if (!seen) ag[na++] = a;
        }
        for (int p = 0; p < na; ++p)
            for (int q = 0; q < na; ++q)
                if (p != q) adj[ag[p]].push_back(ag[q]);
    }
    for (auto& v : adj) { std::sort(v.begin(), v.end()); v.erase(std::unique(v.begin(), v.end()), v.end()); }
    // greedy distance-2 colouring
    out.colour.assign(n_agg, -1);
    int n_col = 0;
    std::vector<char> used;
    for (int a = 0; a < n_agg; ++a) {
        used.assign(n_col + 1, 0);
        for (int32_t j : adj[a]) {
            if (out.colour[j] >= 0) used[out.colour[j]] = 1;
            for (int32_t k : adj[j]) if (k != a && out.colour[k] >= 0) used[out.colour[k]] = 1;
        }
        int col = 0;
        while (col < n_col && used[col]) ++col;
        out.colour[a] = col;
        n_col = std::max(n_col, col + 1);
    }
    out.n_col = n_col;
    out.nbrcol.assign((size_t)n_agg * n_col, -1);
    for (int a = 0; a < n_agg; ++a) {
        out.nbrcol[(size_t)a * n_col + out.colour[a]] = a;
        for (int32_t j : adj[a]) {
            int32_t& slot = out.nbrcol[(size_t)a * n_col + out.colour[j]];
            if (slot >= 0 && slot != j) return "coarse space: colouring is not distance-2";
            slot = j;
        }
    }
    return "";
}

}  // namespace pfx
