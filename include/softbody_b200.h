/*
 * softbody_b200.h -- C ABI of libsoftbody_b200.so, the B200-native replacement for
 * the per-timestep inner loop of a Unity soft-body solver component.
 *
 * WHAT EACH ENTRY POINT REPLACES.  The reference mount contains only
 * /root/reference/README.md:1 ("# SoftbodyUnity"); the C# solver component this
 * library is meant to sit behind is NOT IN MOUNT, so no
 * reference file:line can be cited for any entry point.  The surface below is the
 * one BASELINE.json:5 describes: "the reference's solver component API (its
 * MonoBehaviour-facing Step/parameters: stiffness, damping, substeps,
 * iterations)".  Mapping to the (hypothetical) managed component:
 *
 *   MonoBehaviour.Start()        -> sb_create        (mesh ingest, colouring, tiling, upload)
 *   inspector fields             -> sb_params / sb_set_params
 *   FixedUpdate() { Step(dt); }  -> sb_step          (predict, project, collide, velocity, normals)
 *   LateUpdate() mesh write-back -> sb_read_positions / sb_read_normals / sb_read_surface
 *   OnDestroy()                  -> sb_destroy
 *
 * Conventions: cdecl, plain pointers and sizes, all structs are blittable
 * (4-byte scalars and 8-byte pointers only, natural alignment, no packing
 * pragmas).  Every function returns SB_OK (0) or a negative sb_status; no C++
 * exception crosses this boundary.  The library owns all device memory; the
 * caller owns every host buffer and the library keeps no host pointer after a
 * call returns.  One handle is used by one thread at a time.  There is NO CPU
 * fallback: without a CUDA device sb_create fails with SB_E_CUDA.
 */
#ifndef SOFTBODY_B200_H
#define SOFTBODY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_ABI_VERSION 3u

typedef struct sb_solver *sb_handle;

typedef enum sb_status {
  SB_OK = 0,
  SB_E_ARG = -1,   /* null pointer, bad size, index out of range, degenerate mesh */
  SB_E_CUDA = -2,  /* CUDA runtime error (message via sb_last_error) */
  SB_E_NCCL = -3,  /* reserved for the multi-GPU halo exchange */
  SB_E_NAN = -4,   /* non-finite state detected */
  SB_E_STATE = -5, /* call not valid in the current state (e.g. a device call on a host-only handle, a peer wait that timed out) */
  SB_E_NOMEM = -6
} sb_status;

/* sb_params.flags */
#define SB_FLAG_NO_GROUND 1  /* disable the ground plane */
#define SB_FLAG_FAST_MATH 2  /* rsqrt/rcp approximations in the projection kernels
                                (default is IEEE sqrt/div, bit-identical to the CPU oracle) */
#define SB_FLAG_NO_GRAPH 4   /* launch kernels one by one instead of one CUDA graph per frame */
#define SB_FLAG_NO_NORMALS 8 /* skip the per-frame normal recompute */
#define SB_FLAG_NO_PDL 16    /* launch the tile passes WITHOUT programmatic dependent launch.  By default each pass releases
                                its dependents once its rounds are done, so the next pass's prologue (tile tables, first
                                constraint records) runs under this pass's tail: 8.8 -> 8.5 ms per frame at 1 M vertices */
#define SB_FLAG_DAG 32       /* run all tile passes of a substep as ONE persistent kernel over the tile dependency graph
                                (single-GPU meshes planned as balanced shifted tilings; ignored otherwise) */
#define SB_FLAG_NO_SNAKE 64  /* run the tile passes of every iteration in the same order.  By default odd iterations of a
                                substep run them backwards (a symmetric Gauss-Seidel sweep): the last pass of one iteration
                                and the first of the next then work on the same tiles and share one launch */
#define SB_FLAG_NO_FUSE 128  /* one launch per tile pass and separate predict / finish kernels.  By default consecutive
                                occurrences of a pass are one launch that keeps the positions in shared memory, and
                                predict / collide + velocity update run inside the tile launches at the substep
                                boundaries (same arithmetic, same order: the results do not depend on this flag) */

/*
 * Solver parameters (the inspector fields).  Names per BASELINE.json:5; units,
 * defaults and the stiffness->compliance mapping are [SPEC]:
 *   compliance = 1/stiffness for finite stiffness > 0, 0 for +INFINITY,
 *   and stiffness <= 0 switches that constraint family off.
 */
typedef struct sb_params {
  float dt;                 /* frame step in seconds, used when sb_step is given dt <= 0 */
  int32_t substeps;         /* >= 1 */
  int32_t iterations;       /* constraint sweeps per substep, >= 0 */
  float stiffness_distance; /* edge (spring) constraints */
  float stiffness_volume;   /* tet-volume constraints */
  float damping;            /* 1/s: v *= max(0, 1 - h*damping) after each substep */
  float friction;           /* 0..1: share of tangential motion removed on ground contact */
  float gravity[3];
  float ground_y;
  int32_t flags;
} sb_params; /* 48 bytes */

/* Mesh and build options, read only during sb_create. */
typedef struct sb_mesh_desc {
  const float *pos_xyz;     /* n_verts * 3, rest pose */
  const int32_t *tets;      /* n_tets * 4 vertex ids */
  const int32_t *surf_tris; /* n_tris * 3 vertex ids, may be NULL when n_tris == 0 */
  const float *inv_mass;    /* n_verts, or NULL: lumped from density; 0 pins a vertex */
  void *stream;             /* cudaStream_t to run on, or NULL: the library creates one */
  const int32_t *edges;     /* n_edges * 2 vertex ids, or NULL: the unique edges of the tets.  Given explicitly by
                               a partitioner, whose ranks must each sweep an edge exactly once */
  uint32_t n_verts;
  uint32_t n_tets;
  uint32_t n_tris;
  float density;            /* kg/m^3, used when inv_mass == NULL */
  int32_t device;           /* CUDA device ordinal */
  int32_t tile_cap;         /* max vertices per shared-memory tile; 0 = auto */
  int32_t max_tile_passes;  /* -1 = auto; 0 = global colour batches only */
  int32_t block_threads;    /* threads per tile CTA (32, 64, 128, 160, 192 or 256); 0 = auto per pass */
  int32_t later_tile_cap;   /* max vertices per tile in passes after the first; 0 = auto */
  int32_t host_threads;     /* threads for the host-side build; 0 = auto */
  int32_t round_width;      /* 16-byte constraint-record words per thread per round: a round = one colour of 2 * width *
                               block_threads free edges or of block_threads compounds.  1: a compound is one tet with its
                               attached edges; 2: a bi-tet, two tets that share a face with up to four attached edges
                               (sb_get_tet_mates).  0 = auto: 1 (bi-tets halve the rounds but serialise two projections per
                               thread: measured slower) */
  int32_t attach_edges;     /* 0 = auto, 1 = on, 2 = off: project each edge right after a tet that contains it, from the
                               registers holding the tet's vertices (the tet's vertex ROLES are then an even permutation
                               of the caller's order, see sb_get_tet_roles).  Auto: on */
  int32_t tilings;          /* 0 = auto; 1 = hierarchical tile passes only; N >= 2 = N balanced shifted tilings */
  int32_t n_ghost_verts;    /* partitioned meshes: the LAST n_ghost_verts vertices are ghost copies of vertices
                               another rank owns (never integrated here; constraints among ghosts are dropped) */
  uint32_t n_edges;         /* entries of `edges` (ignored when edges == NULL) */
  int32_t dist_ranks;       /* 0, or the number of GPUs (2..8) the mesh will be spread over with sb_dist_setup: the
                               boxes of the unshifted tiling are then cut into that many compact blocks and numbered
                               block by block (fewer tiles straddle two GPUs than with slabs of the default order) */
} sb_mesh_desc; /* 112 bytes */

/* Sizes and build statistics, for logs, benches and the byte model. */
typedef struct sb_info {
  uint32_t n_verts, n_edges, n_tets, n_tris;
  uint32_t n_surface_verts;
  uint32_t n_tile_passes;        /* shared-memory tile passes per iteration */
  uint32_t n_tilings;            /* balanced shifted tilings among them (1 = hierarchical only) */
  uint32_t n_ghost_verts;
  uint32_t first_cut_pass;       /* first pass of constraint group 1 (those touching ghosts); == n_tile_passes if none */
  uint64_t constraints_cut;      /* constraints in group 1 */
  uint32_t n_global_batches;     /* leftover global colour batches per iteration */
  uint32_t n_batches;            /* independent sets per iteration in the exported schedule */
  uint32_t tiles_in_pass[8];
  uint32_t max_colours_in_pass[8];
  uint64_t constraints_in_pass[8];
  uint64_t edges_in_pass[8];     /* of which distance constraints */
  uint64_t runs_in_pass[8];      /* contiguous vertex runs over all tiles (0: pass is one range per tile) */
  uint64_t constraints_global;
  uint32_t tile_cap;
  uint32_t block_threads;
  uint32_t smem_bytes;           /* dynamic shared memory per tile CTA */
  uint32_t round_width;          /* 16-byte record words per thread per round */
  uint32_t reserved0;
  uint64_t edges_attached;       /* distance constraints that ride with a tet instead of an edge round */
  uint64_t rounds_in_pass[8];    /* rounds (barrier-separated colours) summed over the tiles of the pass */
  uint32_t launches_per_frame;   /* kernel launches inside one sb_step at current params */
  uint64_t device_bytes;         /* device memory held by the handle */
  double build_seconds;          /* host time spent in sb_create */
} sb_info;

/* Layout guard for managed mirrors: writes the ABI version and struct sizes. */
int sb_abi_check(uint32_t *version, uint32_t *sizeof_params, uint32_t *sizeof_desc, uint32_t *sizeof_info);

void sb_default_params(sb_params *out);

int sb_create(const sb_mesh_desc *mesh, const sb_params *params, sb_handle *out);
/*
 * Host-only handle: topology, tiling, colouring and schedule, no CUDA call.  The
 * sb_get_info / sb_get_topology / sb_get_schedule / sb_get_tiles / sb_surface_vertices
 * queries work on it; everything that needs the device returns SB_E_STATE.
 */
int sb_plan(const sb_mesh_desc *mesh, const sb_params *params, sb_handle *out);
/* Releases the handle and everything it owns (synchronises its stream first).  The handle is INVALID afterwards: like
 * free(), a second sb_destroy or any other call on it is undefined.  The managed mirrors null their copy. */
int sb_destroy(sb_handle h);

int sb_set_params(sb_handle h, const sb_params *params);
int sb_get_params(sb_handle h, sb_params *out);

/* Analytic sphere colliders, n * (cx, cy, cz, r); n <= 16; n == 0 clears. */
int sb_set_colliders(sb_handle h, const float *spheres_xyzr, uint32_t n);

/*
 * Analytic colliders of mixed kinds (Unity's SphereCollider / CapsuleCollider / BoxCollider in world space), applied
 * after the ground plane in list order; n <= SB_MAX_COLLIDERS, n == 0 clears; replaces what sb_set_colliders set
 * (which is the same list with kind SPHERE and friction 0).  A vertex inside a collider is moved to its surface
 * along the surface normal there (box: through the nearest face); `friction` in [0, 1] then removes that share of
 * the vertex's tangential motion since the start of the substep.  The operation order is in
 * oracle/xpbd_oracle_impl.h (COLLIDERS); the kernels follow it bit for bit.
 */
#define SB_MAX_COLLIDERS 16
#define SB_COLLIDER_SPHERE 0  /* p = centre xyz, radius */
#define SB_COLLIDER_CAPSULE 1 /* p = end point A xyz, radius, end point B xyz (segment AB swept by the radius) */
#define SB_COLLIDER_BOX 2     /* p = centre xyz, half extents xyz, rotation quaternion (x, y, z, w) box -> world;
                                 a zero quaternion means no rotation */
typedef struct sb_collider {
  int32_t kind;
  float friction;
  float p[10];
} sb_collider; /* 48 bytes */
int sb_set_colliders_ex(sb_handle h, const sb_collider *colliders, uint32_t n);

/*
 * Advance one frame of `dt` seconds (dt <= 0: params.dt): substeps x (predict,
 * iterations x projection sweep, ground/collider response, velocity update), then
 * the surface normals.  Asynchronous: returns after the work is enqueued.
 */
int sb_step(sb_handle h, float dt);
int sb_synchronize(sb_handle h);

/* All vertices, caller's numbering, tightly packed xyz.  Synchronises. */
int sb_read_positions(sb_handle h, float *dst_xyz, uint32_t n_verts);
int sb_read_normals(sb_handle h, float *dst_xyz, uint32_t n_verts);
/* Surface vertices only (ascending vertex id of those on surf_tris). */
int sb_surface_vertices(sb_handle h, int32_t *ids, uint32_t capacity, uint32_t *n_surface);
int sb_read_surface(sb_handle h, float *dst_pos_xyz, float *dst_nrm_xyz, uint32_t n_surface);

/*
 * Render mesh driven by the tets (Unity: the MeshFilter's surface mesh embedded in the simulated tet mesh).
 * sb_skin_bind finds, for every render vertex, the tet that encloses its REST position (or the nearest one, with
 * extrapolating weights) and its four barycentric weights (w0 = 1 - w1 - w2 - w3), on the host, once.
 * sb_read_skinned evaluates sum_k w_k * x[tet vertex k] on the GPU (operation order: oracle's orc_skin) and, when
 * dst_nrm_xyz is given, area-weighted normals over render_tris; either destination may be NULL.  Synchronises.
 * sb_skin_bind also works on an sb_plan handle (binding only).
 */
int sb_skin_bind(sb_handle h, const float *render_pos_xyz, uint32_t n_render_verts, const int32_t *render_tris,
                 uint32_t n_render_tris);
int sb_skin_get_binding(sb_handle h, int32_t *tet_of, float *bary4, uint32_t n_render_verts);
int sb_read_skinned(sb_handle h, float *dst_pos_xyz, float *dst_nrm_xyz, uint32_t n_render_verts);
/* The binding alone, no handle: tets in the caller's vertex order, weights relative to tets[4 * tet_of[i] ..]. */
int sb_skin_compute(const float *tet_pos_xyz, uint32_t n_verts, const int32_t *tets, uint32_t n_tets,
                    const float *points_xyz, uint32_t n_points, int32_t *tet_of, float *bary4);

/*
 * One frame in one buffer, one copy and one synchronisation each way (what a per-frame host loop should use):
 *   sb_read_packed  -> [ x4 of n_verts | v4 of n_verts | xyz of n_surface surface vertices | their normals ]
 *   sb_write_packed <- [ x4 of n_verts | v4 of n_verts ]
 * n_verts / n_surface are all vertices / surface vertices in ascending vertex id -- or, on one rank of a distributed
 * mesh (sb_dist_setup), the ones that rank owns, ascending vertex id (sb_dist_owned marks them).  sb_packed_sizes
 * gives the counts and the two buffer sizes in bytes (32 n_verts, 32 n_verts + 24 n_surface).  Pinned host memory
 * makes the copies asynchronous to the host until the call's final synchronisation.
 */
int sb_packed_sizes(sb_handle h, uint32_t *n_verts, uint32_t *n_surface, uint64_t *bytes_in, uint64_t *bytes_out);
int sb_read_packed(sb_handle h, void *dst, uint64_t bytes);
int sb_write_packed(sb_handle h, const void *src, uint64_t bytes);

/* Full state as float4 arrays in the caller's numbering: x4 = (x,y,z,inv_mass), v4 = (vx,vy,vz,0). */
int sb_get_state(sb_handle h, float *x4, float *v4, uint32_t n_verts);
int sb_set_state(sb_handle h, const float *x4, const float *v4, uint32_t n_verts);

/*
 * State snapshots for replay and regression vectors (.sbs: "SBSTATE1", version, vertex count, frame number, topology
 * hash, sb_params, x4, v4, FNV-1a checksum; little endian).  sb_save_state / sb_load_state move a handle's state
 * through such a file; loading checks the vertex count and the topology hash, restores the frame counter and, when
 * apply_params != 0, the parameters.  A run resumed from a snapshot is bit-identical to the uninterrupted one.
 * sb_state_write / sb_state_read are the file layer alone (no device; x4 == v4 == NULL reads the header only).
 */
int sb_save_state(sb_handle h, const char *path);
int sb_load_state(sb_handle h, const char *path, int32_t apply_params);
int sb_frames_done(sb_handle h, uint64_t *out); /* sb_step calls since creation or the frame of the loaded snapshot */
int sb_state_write(const char *path, const float *x4, const float *v4, uint32_t n_verts, const sb_params *params,
                   uint64_t frame, uint64_t topo_hash);
int sb_state_read(const char *path, float *x4, float *v4, uint32_t capacity_verts, uint32_t *n_verts, sb_params *params,
                  uint64_t *frame, uint64_t *topo_hash);
uint64_t sb_topology_hash(uint32_t n_verts, const int32_t *tets, uint32_t n_tets);

/* Same 16 doubles as the oracle's diagnostics (energy, volume, momenta, strain, NaN count, min y). */
int sb_diagnostics(sb_handle h, double *out16);

int sb_get_info(sb_handle h, sb_info *out);

/*
 * The vertex roles the tet-volume projection uses: tets_4T[4t..4t+3] is an even permutation of the caller's tet t
 * (same orientation, same volume; the floating-point rounding of the projection follows the roles, so a CPU
 * replay must use them), and the edges projected right after tet t from its roles (0,1) and (2,3), or -1.
 * Identity / -1 when attach_edges is off.  Any pointer may be NULL.
 */
int sb_get_tet_roles(sb_handle h, int32_t *tets_4T, int32_t *edge01_T, int32_t *edge23_T);
/*
 * Bi-tets (round_width 2, the default): tets are projected in pairs that share a face, by one thread that keeps the
 * shared vertices in registers.  mate_T[t] = the tet paired with t or -1; lead_T[t] = 1 for a single tet or the first
 * tet of a pair (projected first, with its attached edges; then its mate with its own).  The roles of a pair are
 * A = (a, s0, s1, s2), B = (b, s1, s0, s2).  Either pointer may be NULL.
 */
int sb_get_tet_mates(sb_handle h, int32_t *mate_T, int32_t *lead_T);

/* Derived topology in the caller's numbering (any pointer may be NULL). */
int sb_get_topology(sb_handle h, int32_t *edges_2E, float *rest_len_E, float *rest_vol6_T, float *inv_mass_V);

/*
 * The Gauss-Seidel processing order of one iteration, equivalent to what the
 * kernels do: entry >= 0 is an edge id, entry < 0 a tet id in the low 31 bits;
 * batch_off (n_batches + 1) delimits vertex-disjoint runs.  Call with NULL
 * arrays to get the sizes.  This is what a CPU solver must replay to be compared
 * ("re-run under the same colour ordering", BASELINE.json:5).
 */
int sb_get_schedule(sb_handle h, int64_t *n_order, int32_t *order, int32_t *n_batches, int64_t *batch_off);
/*
 * The same for the iterations 1, 3, 5 ... of a substep (sb_get_schedule: iterations 0, 2, 4 ...): the tile passes in
 * reverse order unless SB_FLAG_NO_SNAKE / SB_FLAG_DAG is set or the plan has constraints outside the tile passes,
 * in which case it equals sb_get_schedule.  A CPU replay alternates the two orders within every substep.
 */
int sb_get_schedule_odd(sb_handle h, int64_t *n_order, int32_t *order, int32_t *n_batches, int64_t *batch_off);
/*
 * The launches of one frame at the current parameters, 6 int32 per launch: kind (0 predict, 1 finish, 2 tile pass,
 * 3 global batches, 4 constraint group, 5 exchange, 6 normals, 7 tile DAG), arg (pass / group / phase), segments,
 * repetitions per segment, predict-before, finish-after.  *n_launches receives the count; ops6 may be NULL.
 * sb_enqueue(h, SB_OP_LAUNCH, i) enqueues launch i alone (drivers that interleave several handles on one stream).
 */
int sb_frame_program(sb_handle h, int32_t *n_launches, int32_t *ops6, uint32_t capacity);

/* For tile pass `pass`: per caller vertex, the tile that stages it, or -1 (tile_of may be NULL). */
int sb_get_tiles(sb_handle h, uint32_t pass, int32_t *tile_of, uint32_t *n_tiles);

/*
 * Device-timed runs for benches: enqueue n_frames frames on the solver stream
 * between two CUDA events and return the elapsed milliseconds.
 */
int sb_time_frames(sb_handle h, int32_t n_frames, float dt, float *elapsed_ms);
/*
 * Time `reps` back-to-back launches of one kernel of the path in isolation
 * (which: 0 predict, 1 finish, 2 normals, 16+p tile pass p, 32 all global batches,
 * 48 the persistent tile-DAG kernel = every pass of every iteration of one substep).
 * The state is saved and restored around the run.
 */
int sb_time_kernel(sb_handle h, int32_t which, int32_t reps, float *avg_ms);

/*
 * Phased stepping, for one rank of a mesh partitioned over several GPUs.  The caller interleaves
 * these with its halo exchange (INTEGRATION.md / DESIGN.md section 7); all work is enqueued on the
 * handle's stream without synchronising.  sb_step is the same sequence with no exchange.
 */
#define SB_OP_PREDICT 0
#define SB_OP_PROJECT 1 /* arg: constraint group (0 interior, 1 touching ghosts), -1 both */
#define SB_OP_FINISH 2
#define SB_OP_NORMALS 3
#define SB_OP_EXCHANGE 4 /* arg: 0 = exchange A (ghosts refreshed), 1 = exchange B (ghost values returned) */
#define SB_OP_HALO_SEND 5 /* arg: list id; the two halves of an exchange, for several ranks driven from one stream */
#define SB_OP_HALO_RECV 6
#define SB_OP_PASS 7 /* arg: tile pass index (one launch; for drivers that interleave several handles on one stream) */
#define SB_OP_LAUNCH 8 /* arg: index into sb_frame_program (one launch of the frame as sb_step would issue it) */
int sb_set_stream(sb_handle h, void *stream);
int sb_prepare(sb_handle h, float dt); /* pushes parameters for this dt; synchronises */
int sb_enqueue(sb_handle h, int32_t op, int32_t arg);
int sb_halo_set(sb_handle h, int32_t list_id, const int32_t *vertex_ids, uint32_t n);
int sb_halo_pack(sb_handle h, int32_t list_id, void *dst_device);        /* n float4 positions -> dst */
int sb_halo_unpack(sb_handle h, int32_t list_id, const void *src_device); /* src -> positions */
/* halo exchange over peer memory (NVLink P2P stores + flag words), see solver.cu */
int sb_halo_alloc(sb_handle h, int32_t list_id, void **base_out, uint64_t *bytes_out);
int sb_halo_connect(sb_handle h, int32_t list_id, void *peer_base);
int sb_halo_error(sb_handle h, int32_t *out);
int sb_ipc_export(void *device_ptr, unsigned char *handle64);
int sb_ipc_open(int32_t device, const unsigned char *handle64, void **ptr_out);
int sb_lumped_inv_mass(const float *pos_xyz, uint32_t n_verts, const int32_t *tets, uint32_t n_tets, float density, float *out);

/*
 * ONE mesh over several GPUs with no ghosts and no exchange kernels: one address space over NVLink peer memory
 * (DESIGN.md section 7).  Every rank creates the SAME handle (whole mesh, same options) on its own GPU, then:
 *   sb_dist_setup    picks this rank's slab of the device numbering and its share of the tiles of every pass, and
 *                    returns the device addresses of its position array and control block (export them with
 *                    sb_ipc_export, or pass them directly between handles of one process);
 *   sb_dist_connect  (once per peer) gives the peer's two addresses; when all peers are connected the handle is live.
 * sb_step then runs this rank's tiles; a vertex's position lives with the rank that touches it next: a tile loads from
 * its own GPU's array and stores every run of vertices into the array of the rank that holds it in the next launch (at the
 * end of a frame: of the rank that owns it), and kernels of neighbouring ranks order themselves through an epoch word.  All ranks must issue the
 * same sequence of steps (and hold the same flags: the launch sequence must be the same everywhere).  Positions /
 * state read back from a rank are valid for the vertices sb_dist_owned marks; sb_read_packed carries exactly those.
 * Surface normals: each rank computes them for the surface vertices it owns (triangles that reach into a
 * neighbour's slab read the neighbour's positions over NVLink).  A frame always ends with a launch that waits for the
 * neighbours' last stores (the normals launch; with SB_FLAG_NO_NORMALS or no surface, a one-CTA handshake), so a read-back
 * after sb_step sees every owned vertex.  Writing state (sb_set_state, sb_write_packed) while a peer may still be inside
 * its frame is the caller's race: synchronise the ranks first.  Refused on such a handle with SB_E_STATE: whole-mesh
 * reads that would mix current and stale vertices (sb_read_surface, sb_read_normals, sb_diagnostics, sb_save_state /
 * sb_load_state, sb_skin_bind) and stray launches the peers do not run (sb_time_kernel, sb_debug_trace_pass).
 */
int sb_dist_setup(sb_handle h, int32_t rank, int32_t n_ranks, void **x_base_out, void **ctl_base_out);
int sb_dist_connect(sb_handle h, int32_t peer, void *peer_x, void *peer_ctl);
int sb_dist_owned(sb_handle h, uint8_t *owned_V, uint32_t *tiles_per_pass8); /* either pointer may be NULL */
int sb_dist_error(sb_handle h, int32_t *out); /* 1 if a wait for a peer ever timed out (results invalid) */
/* Host only (works on an sb_plan handle): the vertices rank would own and, for one pass, which tiles it would run
   (tile_owner_pass[t] = 0: not this rank's; j + 1: CTA j of the rank's launch, an interior tile; -(j + 1): CTA j, a zone
   tile -- one that waits for the neighbours' epoch and counts towards this rank's; sb_get_tiles gives the tile count).
   Either pointer may be NULL. */
int sb_dist_layout(sb_handle h, int32_t rank, int32_t n_ranks, uint8_t *owned_V, int32_t *tile_owner_pass, uint32_t pass);
/* Host only: replays the hand-over protocol of one frame symbolically for n_ranks ranks -- every tile launch of the frame
   program, every tile of every rank -- and counts the vertices a tile would load while their current value sits in another
   rank's array (n_stale), the vertices that are not back with their owner at a per-vertex kernel, the normals or the end
   of the frame (n_not_home), and the vertices touched by a tile whose rank did not classify the tile as a zone tile although
   the previous holder was another rank, or between ranks that do not poll each other, or read by a neighbour's normals launch
   from a tile that is not a zone tile (n_unordered).  All three are 0 for a correct layout.  n_crossings: how many
   vertex values changed rank during the frame (the NVLink stores of one frame, in vertices).  Any pointer may be NULL. */
int sb_dist_verify(sb_handle h, int32_t n_ranks, uint64_t *n_stale, uint64_t *n_not_home, uint64_t *n_unordered,
                   uint64_t *n_crossings);

/* Debug aid: timestamps of one run of tile pass `pass` (see solver.cu); out holds 64 * 80 words. */
int sb_debug_trace_pass(sb_handle h, uint32_t pass, unsigned long long *out, uint32_t n_words);

/* Debug aid, host only: decodes the device constraint streams as the kernel reads them and compares them with
   the exported schedule; *n_bad = records that differ (0 on a sound plan).  Optionally also the modelled
   shared-memory load wavefronts of one sweep (bank conflicts included) and their conflict-free count. */
int sb_debug_verify_streams(sb_handle h, uint64_t *n_bad, uint64_t *wavefronts, uint64_t *wavefronts_ideal);

const char *sb_last_error(sb_handle h); /* h may be NULL: last sb_create failure on this thread */

/*
 * Mesh ingest, host only (no device needed): the step BEFORE the path.  An sb_tetmesh owns positions, positively
 * oriented tets and outward-wound boundary triangles; sb_tetmesh_desc points an sb_mesh_desc at them for sb_create
 * (valid until sb_tetmesh_free).  Errors: negative sb_status, message via sb_ingest_last_error (per thread).
 *   sb_tetmesh_from_surface  closed triangle surface (Unity Mesh.vertices / triangles) -> lattice of `spacing`-sized
 *                            cells whose centre has non-zero winding number, five tets per cell
 *   sb_tetmesh_snap_to_surface  optional second step: the staircase boundary of that lattice pulled onto the surface
 *   sb_tetmesh_from_arrays   caller's arrays; orientation fixed, boundary extracted when n_tris == 0
 *   sb_tetmesh_load / save   TetGen <base>.node + .ele (+ .face) or Gmsh MSH ASCII (.msh: 2.x and 4.1 are read, 2.2 is written), by extension
 */
typedef struct sb_tetmesh *sb_tetmesh_handle;
int sb_tetmesh_from_surface(const float *surf_pos_xyz, uint32_t n_verts, const int32_t *surf_tris, uint32_t n_tris,
                            float spacing, sb_tetmesh_handle *out);
/* Boundary vertices of m moved onto the surface where it is nearer than max_dist and no tet at the vertex shrinks
   below 30 % of its lattice volume (ingest.cpp); *n_moved (may be NULL) counts them. */
int sb_tetmesh_snap_to_surface(sb_tetmesh_handle m, const float *surf_pos_xyz, uint32_t n_verts, const int32_t *surf_tris,
                               uint32_t n_tris, float max_dist, uint32_t *n_moved);
int sb_tetmesh_from_arrays(const float *pos_xyz, uint32_t n_verts, const int32_t *tets, uint32_t n_tets,
                           const int32_t *tris, uint32_t n_tris, sb_tetmesh_handle *out);
int sb_tetmesh_load(const char *path, sb_tetmesh_handle *out);
int sb_tetmesh_save(sb_tetmesh_handle m, const char *path);
int sb_tetmesh_sizes(sb_tetmesh_handle m, uint32_t *n_verts, uint32_t *n_tets, uint32_t *n_tris);
int sb_tetmesh_copy(sb_tetmesh_handle m, float *pos_xyz, int32_t *tets, int32_t *tris); /* any pointer may be NULL */
int sb_tetmesh_desc(sb_tetmesh_handle m, sb_mesh_desc *desc);
int sb_tetmesh_free(sb_tetmesh_handle m);
const char *sb_ingest_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* SOFTBODY_B200_H */
