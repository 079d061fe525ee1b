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
struct I2 { int32_t x, y; };
struct I4 { int32_t x, y, z, w; };

struct PlanOptions {
  int tile_cap = 0;         // max vertices per tile (0 = auto)
  int later_cap = 0;        // max vertices per tile in passes >= 1 (0 = auto)
  int max_tile_passes = -1; // -1 = auto
  int n_sm = 148;           // grid sizing target
  int threads = 0;          // host build threads (0 = hardware_concurrency)
  int slot_bytes = 0;       // bytes per shared-memory staging slot (0 = 2016); multiple of 48
  int n_slots = 0;          // staging slots per CTA (0 = 4)
  int tilings = 0;          // 0 = auto; 1 = hierarchical passes only; N >= 2 = N balanced shifted tilings
};

// One shared-memory tile pass: every tile is a vertex-disjoint set of vertices
// whose assigned constraints touch only those vertices.
//
// Device data: a tile's constraints are a byte stream cut into CHUNKS that the
// kernel stages into shared memory with bulk copies.  A chunk holds records of
// one kind and one colour:
//   edge chunk, n records : n2 = roundup(n, 2) x { a | b << 16, bits(L0) }          (8 B each)
//   tet chunk,  n records : n4 = roundup(n, 4) x { p0 | p1 << 16, p2 | p3 << 16 }   (8 B each)
//                           followed by n4 x float (6 * rest volume)
// (local 16-bit vertex ids; padding records are zero and never executed).
// chunks[i] = { byte offset / 16 into `stream`, n | kind << 30 | barrier << 31 }:
// `barrier` asks for a CTA barrier after the chunk (set on the last chunk of every
// colour, and often enough that the staging ring can always be refilled).
struct TilePass {
  bool contiguous = false;          // tile t == device vertex range [vert_off[t], vert_off[t+1])
  std::vector<uint32_t> vert_off;   // n_tiles + 1
  std::vector<uint32_t> tile_verts; // device vertex ids (empty when contiguous)
  std::vector<uint32_t> run_off;    // n_tiles + 1 offsets into runs (non-contiguous passes)
  std::vector<U2> runs;             // per tile: {first device id, first local id} per run, then {0, n_verts}
  std::vector<uint32_t> chunk_off;  // n_tiles + 1 offsets into chunks
  std::vector<U2> chunks;
  std::vector<uint32_t> stream;     // 32-bit words; every chunk starts 16-byte aligned
  // schedule bookkeeping (host only): per tile, constraints in processing order
  std::vector<uint64_t> ent_off;    // n_tiles + 1 offsets into ents
  std::vector<int32_t> ents;        // >= 0 edge id, < 0 tet id | 0x80000000
  std::vector<uint32_t> col_off;    // n_tiles + 1 offsets into col_cnt
  std::vector<uint32_t> col_cnt;    // constraints per colour, edge colours first
  std::vector<uint32_t> n_ecol;     // edge colours of tile t
  uint32_t max_ecol = 0, max_tcol = 0, max_tile_verts = 0, max_chunks = 0;
  int group = 0;                    // 0 interior constraints, 1 constraints that touch a ghost vertex
  uint64_t n_edges = 0, n_tets = 0;
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
  uint32_t tile_cap = 0, slot_bytes = 0, n_slots = 0, n_tilings = 1;
  double build_seconds = 0;

  // The equivalent sequential order of one iteration (see sb_get_schedule).
  void export_schedule(std::vector<int32_t> &order, std::vector<int64_t> &batch_off) const;
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
