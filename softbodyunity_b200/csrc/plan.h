// plan.h -- host-side build of everything the kernels consume: canonical topology,
// vertex tiling, per-tile local colouring, leftover global colouring, and the
// equivalent sequential Gauss-Seidel order.  Pure C++ (no CUDA) so it can be
// exercised without a device.
//
// Reference: NOT IN MOUNT (/root/reference/README.md:1 is the whole reference).
// BASELINE.json:5 asks for "greedy graph colouring done once on the host so
// Gauss-Seidel projection is race-free per colour" and "shared-memory staging of
// each graph-colour's constraint tile"; this file is that host step.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sb {

struct U2 { uint32_t x, y; };
struct U4 { uint32_t x, y, z, w; };
struct I2 { int32_t x, y; };
struct I4 { int32_t x, y, z, w; };

// Where record k of a colour sits in its round: records fill the CTA's warps from the front (so that the
// warps behind the last record hold nothing but padding and skip the round), eight consecutive records
// on eight consecutive lanes (the octets the bank-aware ordering arranges), sub-records interleaved by octet.
inline void round_slot(uint32_t k, uint32_t subs, uint32_t &thr, uint32_t &sub) {
  thr = 8u * (k / (8u * subs)) + (k & 7u);
  sub = (k >> 3) % subs;
}

struct PlanOptions {
  int tile_cap = 0;         // max vertices per tile (0 = auto)
  int later_cap = 0;        // max vertices per tile in passes >= 1 (0 = auto)
  int max_tile_passes = -1; // -1 = auto
  int n_sm = 148;           // grid sizing target
  int threads = 0;          // host build threads (0 = hardware_concurrency)
  int block_threads = 0;    // threads per tile CTA: 32, 64, 128 or 256 (0 = auto per pass, by constraints per tile)
  int round_width = 0;      // 16-byte record words per thread per round: 1 or 2 (0 = 1)
  int compounds = -1;       // attach edges to tets (-1 = auto: on unless the mesh is one rank of a partition, 0 off, 1 on)
  int tilings = 0;          // 0 = auto; 1 = hierarchical passes only; N >= 2 = N balanced shifted tilings
  int dist_ranks = 0;       // >= 2: the mesh will be spread over this many GPUs (sb_dist_setup): the boxes of the unshifted
                            // tiling are cut into that many compact blocks (recursive bisection of the box grid, weighted by
                            // vertex count) and numbered block by block, so that every rank's vertices are one range of
                            // device ids and few tiles straddle two ranks
};

// One shared-memory tile pass: every tile is a vertex-disjoint set of vertices
// whose assigned constraints touch only those vertices.
//
// Device data: a tile's constraints are a stream of ROUNDS.  A round is one colour of
// one kind, capped at what the CTA projects in one go: every thread owns `width`
// 16-byte words of every round,
//   edge round: word = 2 x { a | b << 16, bits(L0) }              -> 2 * width * bt edges per colour
//   tet round : word = { p0 | p1 << 16, p2 | p3 << 16, bits(6 V0), bits(L01) } -> width * bt tets per colour
//               plus one float per word in `aux`: L23.  A tet record is a COMPOUND: the tet, then the edge
//               between its roles (0,1) with rest length L01, then the edge between roles (2,3) with rest
//               length L23 (NaN = no such edge attached).  An attached edge is projected from the registers
//               that already hold the tet's vertices, so it costs no shared-memory traffic and no round.
// (local 16-bit vertex ids; an all-zero record is padding: a == b / p0 == p1 is never projected).
// Thread `tid` reads its words of round r at stream[off + (r * bt + tid) * width ...], a fully
// coalesced access that it prefetches several rounds ahead; the CTA synchronises after every
// round.  Records are packed towards the first warps (round_slot): a warp that holds only padding
// skips the round's work.  The colouring is capacity-limited (first free colour that is not full), so a tile
// needs max(valence, ceil(n / capacity)) rounds and every round but the last is full.
// rounds[t] = { offset / 16 of tile t's first round in `stream`, edge rounds, tet rounds, offset of its first
//               tet round in `aux` (floats) }.
struct TilePass {
  bool contiguous = false;          // tile t == device vertex range [vert_off[t], vert_off[t+1])
  std::vector<uint32_t> vert_off;   // n_tiles + 1
  std::vector<uint32_t> tile_verts; // device vertex ids (empty when contiguous)
  std::vector<uint32_t> run_off;    // n_tiles + 1 offsets into runs (non-contiguous passes)
  std::vector<U2> runs;             // per tile: {first device id, first local id} per run, then {0, n_verts}
  std::vector<U4> rounds;           // n_tiles
  std::vector<uint32_t> launch_order; // CTA j of the pass runs tile launch_order[j]: heaviest tiles first, dealt to
                                      // the SMs in a snake so that every SM's single wave of CTAs weighs the same
  std::vector<uint32_t> stream;     // 32-bit words; every round starts 16-byte aligned
  std::vector<float> aux;           // per tet round: width * bt floats (L23 of the compound, NaN if none)
  std::vector<uint8_t> col_flags;   // per tet colour index: bit 0 some tile attaches a (0,1) edge there, 1 a (2,3) edge, 2 a mate
                                    // (bi-tets), 3 / 4 a (0,1) / (2,3) edge of a mate: the batches the exported schedule has
  uint32_t bt = 64, width = 1;      // CTA threads and 16-byte words per thread per round of this pass
  // schedule bookkeeping (host only): per tile, constraints in processing order
  std::vector<uint64_t> ent_off;    // n_tiles + 1 offsets into ents
  std::vector<int32_t> ents;        // >= 0 edge id, < 0 tet id | 0x80000000
  std::vector<uint32_t> col_off;    // n_tiles + 1 offsets into col_cnt
  std::vector<uint32_t> col_cnt;    // constraints per colour, edge colours first
  std::vector<uint32_t> n_ecol;     // edge colours of tile t
  uint32_t max_ecol = 0, max_tcol = 0, max_tile_verts = 0;
  int group = 0;                    // 0 interior constraints, 1 constraints that touch a ghost vertex
  // tile dependencies for the persistent (DAG) kernel: tiles of the PREVIOUS tiling pass (cyclically) that
  // share a vertex with tile t, by full box membership (empty unless the plan is DAG-capable)
  std::vector<uint32_t> dep_off;    // n_tiles + 1
  std::vector<uint32_t> dep_list;
  uint64_t n_edges = 0, n_tets = 0, rounds_total = 0;
  uint32_t n_tiles() const { return vert_off.empty() ? 0u : (uint32_t)vert_off.size() - 1; }
};

struct GlobalBatch {
  bool tet;
  uint32_t off, cnt; // into g_edges/g_tets
  int group;
};

struct Plan {
  // sizes
  uint32_t V = 0, E = 0, T = 0, F = 0;
  uint32_t n_ghost = 0; // the last n_ghost caller vertices are ghost copies owned by another rank
  // canonical topology, caller's numbering
  std::vector<float> pos;       // 3V rest pose
  std::vector<int32_t> tets;    // 4T
  std::vector<int32_t> edges;   // 2E, a < b, sorted
  std::vector<int32_t> tris;    // 3F
  std::vector<int32_t> tet_roles; // 4T: tets[t] in the role order the projection uses (an even permutation)
  std::vector<int32_t> tet_e01, tet_e23; // T: edge attached to roles (0,1) / (2,3) of the tet, or -1
  std::vector<int32_t> edge_owner;       // E: tet the edge is attached to, or -1 (projected in an edge round)
  std::vector<int32_t> tet_mate;         // T: the tet this one forms a bi-tet with (they share a face), or -1
  std::vector<uint8_t> tet_lead;         // T: 1 for a single tet or the first tet (A) of a bi-tet, 0 for its second (B)
  uint64_t tets_paired = 0;
  uint64_t edges_attached = 0;
  std::vector<float> rest_len;  // E
  std::vector<float> rest_vol6; // T
  std::vector<float> inv_mass;  // V
  // device numbering
  std::vector<uint32_t> perm; // device id -> caller id
  std::vector<uint32_t> inv;  // caller id -> device id
  // surface
  std::vector<int32_t> surf_ids;      // caller ids of vertices on surf_tris, ascending
  std::vector<uint32_t> surf_tri_off; // n_surface + 1 (CSR over incident triangles, ascending tri id)
  std::vector<uint32_t> surf_tri_ids;
  // schedule
  std::vector<TilePass> passes;
  std::vector<I2> g_edges; // device vertex ids
  std::vector<float> g_elen;
  std::vector<int32_t> g_eid;
  std::vector<I4> g_tets;
  std::vector<float> g_trest;
  std::vector<int32_t> g_tid;
  std::vector<GlobalBatch> gbatches;
  // options actually used
  uint32_t tile_cap = 0, round_width = 1, n_tilings = 1;
  uint32_t dist_ranks = 0;            // blocks the boxes of tiling 0 were numbered by (PlanOptions::dist_ranks), or 0
  std::vector<uint32_t> dist_slab_lo; // dist_ranks + 1: first device id of every block
  bool dag_ok = false; // the passes are exactly the balanced tilings (one group, no leftovers): DAG kernel usable
  double build_seconds = 0;

  // The equivalent sequential order of one iteration (see sb_get_schedule); reverse: the tile passes last to first
  // (the order of the odd iterations of a substep, sb_get_schedule_odd).
  void export_schedule(std::vector<int32_t> &order, std::vector<int64_t> &batch_off, bool reverse = false) const;
};

struct MeshInput {
  const float *pos_xyz;
  const int32_t *tets;
  const int32_t *surf_tris;
  const float *inv_mass;
  uint32_t n_verts, n_tets, n_tris;
  float density;
  uint32_t n_ghost = 0;
  const int32_t *edges = nullptr; // optional explicit constraint edges (2 * n_edges ids)
  uint32_t n_edges = 0;
};

void lumped_inv_mass_into(const float *pos, uint32_t V, const int32_t *tets, uint32_t T, float density, float *out);

// Returns an empty string on success, else a message (argument errors).
std::string build_plan(const MeshInput &in, const PlanOptions &opt, Plan &out);

} // namespace sb
