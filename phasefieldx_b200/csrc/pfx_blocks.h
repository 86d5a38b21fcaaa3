// pfx_blocks.h — host-side mesh preprocessing: CTA tiles ("blocks") and rank partitions.
//
// Replaces what dolfinx does at mesh/function-space creation for the reference (graph
// partitioner, IndexMap, Scatterer, dofmap reordering; third-party, not in the reference
// tree).  Two levels share one idea — owner computes, with one overlap layer of cells:
//   * ranks  (GPUs): owned nodes first, ghosts after; every cell touching an owned node is
//     kept locally, so an operator apply needs only a FORWARD halo of its input vector;
//   * blocks (CTAs): each block owns a contiguous range of (renumbered) nodes and lists every
//     cell touching one of them, so a CTA accumulates its own rows in shared memory and
//     writes them coalesced — no atomics, deterministic summation order.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace pfx {

struct BlockMesh {
    int nv = 0, dim = 0;
    int64_t n_nodes = 0;      // local nodes (owned + ghosts)
    int64_t n_owned = 0;
    int64_t n_cells = 0;
    int nb = 0;               // number of blocks
    std::vector<int32_t> perm;       // new -> old (caller) node id, size n_nodes
    std::vector<int32_t> iperm;      // old -> new
    std::vector<int32_t> node0;      // [nb+1] owned range of each block (new ids)
    std::vector<int32_t> halo_off;   // [nb+1]
    std::vector<int32_t> halo;       // new ids of non-owned nodes referenced by the block, sorted
    std::vector<int32_t> elem_off;   // [nb+1]
    std::vector<int32_t> nprim;      // [nb] leading elements of the block integrated in scalar reductions
    std::vector<uint16_t> lconn;     // [lstride*sum elems] block-local vertex ids (triangles: 4th = 0)
    int lstride = 4, vbits = 2;      // hexahedra: 8 ids per cell, 3-bit vertex codes in the incidence lists
    std::vector<int32_t> elem_gid;   // [sum elems] caller cell id (diagnostics/tests)
    std::vector<int32_t> inc_ptr;    // [n_owned+1] into inc
    std::vector<uint16_t> inc;       // per owned node: (local elem << 2 | vertex), ascending
    std::vector<int32_t> meta;       // [8*nb] {n0, n_own, h0, n_halo, e0, n_elems, inc0, n_inc} per block
    std::vector<uint16_t> inc8;      // triangles: [8*n_owned] first 8 incidence codes per node (0xFFFF = none,
                                     // slot 7 = 0xFFFE when the node has more than 8: rest in inc[])
    int max_own = 0, max_loc = 0, max_elems = 0;
    int64_t total_elems = 0;         // with overlap duplicates
    // ---- own cells: every local cell is "own" for exactly ONE block, the one holding its lowest-numbered owned node.
    // A block's list starts with its own cells (rank-primary ones first); own_c0[b] + j is a compact index over all
    // local cells (per-cell data such as the cached element tangents is stored once, in this order).
    std::vector<int32_t> nlow;       // [nb] own cells of the block (>= nprim)
    std::vector<int32_t> own_c0;     // [nb+1] compact own-cell index of the block's first own cell
    int64_t n_own_cells = 0;
};

// Build blocks of at most `max_own` owned nodes by recursive coordinate bisection of the owned
// nodes.  coords [n_nodes,3]; cells [n_cells,nv] (caller ids; every cell must touch >=1 owned
// node or it is ignored); cell_primary may be null (= all 1).  Returns "" or an error text.
std::string build_blocks(int nv, int dim, int64_t n_nodes, int64_t n_owned, int64_t n_cells, const double* coords,
                         const int32_t* cells, const uint8_t* cell_primary, int max_own, BlockMesh& out);

// Aggregates of the PCG coarse space (pfx_coarse.cuh): runs of `g` consecutive blocks; chunks = even splits
// of an aggregate into <= chunk_nodes nodes; regions = contiguous runs of aggregates whose couplings the
// coarse matrix keeps; distance-2 colouring of the aggregate graph (i ~ j when a cell has owned vertices in
// both) so that one operator apply probes all aggregates of a colour at once.
struct AggregateSet {
    int n_agg = 0, S = 1, n_chunks = 0, n_col = 0, n_reg = 1;
    std::vector<int32_t> agg_node0;      // [n_agg+1] owned-node range (new ids)
    std::vector<int32_t> chunk_node0;    // [n_chunks+1], chunk c belongs to aggregate c / S; even starts
    std::vector<int32_t> reg_agg0;       // [n_reg+1]
    std::vector<int32_t> agg_reg;        // [n_agg]
    std::vector<int32_t> colour;         // [n_agg]
    std::vector<int32_t> nbrcol;         // [n_agg*n_col] the aggregate of that colour within distance one (or -1)
    std::vector<float> rel;              // [n_owned*dim] (x - centroid(aggregate)) / half extent(aggregate)
    std::vector<double> cen;             // [n_agg*3] centroid of every aggregate
    std::vector<double> half;            // [n_agg] the half extent used to scale rel (1 if degenerate)
};

// cells / coords in caller ids as given to build_blocks.  g blocks per aggregate, regions of at most
// reg_aggs aggregates.  Returns "" or an error text.
std::string build_aggregates(const BlockMesh& bm, int64_t n_cells, const double* coords, const int32_t* cells, int g,
                             int reg_aggs, int chunk_nodes, AggregateSet& out);

// Node graph of a simplex mesh in sliced-ELL layout (3D assembled operators, pfx_sparse.cuh).  Rows = owned nodes
// in block order; a slice = kSlice consecutive rows; entry k of every row of a slice is stored side by side
// (position (slice_k0[s] + k) * kSlice + lane), so a warp that handles one slice reads its column indices and
// matrix values fully coalesced.  Rows are padded to the longest row of their slice with the row's own index.
constexpr int kSlice = 32;
struct NodeGraph {
    int64_t n_rows = 0, n_slices = 0, total_k = 0;     // total_k = sum of slice lengths
    int max_k = 0;
    std::vector<int32_t> slice_k0;     // [n_slices+1]
    std::vector<int32_t> col;          // [total_k * kSlice] column node (new numbering; ghosts allowed)
    std::vector<int32_t> cell_nodes;   // [n_own_cells * 4] nodes of every local cell (new numbering), compact own-cell index
    std::vector<int32_t> nc_ptr;       // [n_rows+1] cells incident to every owned node ...
    std::vector<int32_t> nc_code;      // ... as (compact own-cell index << 2 | vertex of the node in the cell), ascending
    std::vector<uint32_t> nc_kpos;     // ... and the row entries k of the cell's four nodes, one byte per vertex
};

// cells / coords in caller ids as given to build_blocks.  Returns "" or an error text.
std::string build_node_graph(const BlockMesh& bm, int64_t n_cells, const int32_t* cells, NodeGraph& out);

struct RankPart {
    std::vector<int64_t> l2g;            // local -> global node (owned first, then ghosts by owner, by gid)
    int64_t n_owned = 0;
    std::vector<int64_t> cell_gid;       // local cells (global ids, ascending)
    std::vector<int32_t> ghost_owner;    // [n_ghost]
    std::vector<uint8_t> cell_primary;   // [n_local_cells]
    std::vector<int32_t> neighbors;      // ranks exchanged with (ascending)
    std::vector<int64_t> send_off, recv_off;   // [n_neighbors+1]
    std::vector<int32_t> send_idx;       // local owned indices to pack, grouped by neighbor
};

struct Partition {
    int n_ranks = 0;
    int64_t n_nodes = 0, n_cells = 0;
    std::vector<int32_t> cell_rank, node_owner;
    std::vector<RankPart> ranks;
};

// cell_rank_in may be null -> recursive coordinate bisection of the cell centroids.
// node owner = lowest rank among the adjacent cells' ranks (deterministic rule).
// cell_rank[n_cells] by METIS k-way on the cell dual graph; edge_cut (optional) = facets shared by cells of two ranks
std::string metis_cell_rank(int nv, int dim, int64_t n_nodes, int64_t n_cells, const int32_t* cells, int n_ranks, int32_t* cell_rank,
                            int64_t* edge_cut);
std::string build_partition(int nv, int64_t n_nodes, int64_t n_cells, const double* coords, const int32_t* cells,
                            int n_ranks, const int32_t* cell_rank_in, Partition& out);

}  // namespace pfx
