// solver.cu -- device runtime and C ABI of libsoftbody_b200.so (include/softbody_b200.h).
//
// Reference: NOT IN MOUNT (/root/reference/README.md:1 is the whole reference); the
// surface mirrors the solver component BASELINE.json:5 describes (Step + stiffness,
// damping, substeps, iterations).  No CPU fallback exists: every stage of sb_step is
// a CUDA kernel from kernels.cuh, replayed as one CUDA graph per frame.
#include <cuda_runtime.h>

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/softbody_b200.h"
#include "kernels.cuh"
#include <algorithm>
#include "plan.h"

namespace sb {

struct CudaError {
  cudaError_t code;
  const char *what;
  int line;
};

#define CK(expr)                                          \
  do {                                                    \
    cudaError_t e_ = (expr);                              \
    if (e_ != cudaSuccess) throw CudaError{e_, #expr, __LINE__}; \
  } while (0)

template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count, uint64_t *tally) {
    release();
    n = count;
    if (count) {
      CK(cudaMalloc(&p, count * sizeof(T)));
      if (tally) *tally += count * sizeof(T);
    }
  }
  template <class S>
  void upload(const std::vector<S> &h, uint64_t *tally) {
    static_assert(sizeof(S) == sizeof(T), "layout");
    alloc(h.size(), tally);
    if (!h.empty()) CK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { release(); }
};

struct PassBufs {
  DevBuf<uint32_t> vert_off, tile_verts, stream, run_off, order;
  DevBuf<uint2> runs;
  DevBuf<uint4> rounds, desc;
  DevBuf<float> aux;
  PassDev dev{};
  uint32_t smem = 0;
  uint32_t grid = 0;           // CTAs launched (tiles of the pass; this rank's share of them when distributed)
  uint32_t bt = 64, width = 1; // threads per CTA and record words per thread per round of this pass
  bool empty = false;          // no constraint in the pass: never launched
};

static const int kDiagBlocks = 592;

// Caller's collider -> what k_finish reads.  Same arithmetic as the oracle's orc_prepare_collider (host code is
// built with -ffp-contract=off): capsule axis and 1/|axis|^2 in float, box axes from the quaternion in double
// (normalised by 2/|q|^2, so any non-zero quaternion is a rotation), rounded to float once.
static DevCollider prepare_collider(const sb_collider &c) {
  DevCollider d{};
  const float *p = c.p;
  d.a = make_float4(p[0], p[1], p[2], p[3]);
  if (c.kind == SB_COLLIDER_CAPSULE) {
    const float ax = p[4] - p[0], ay = p[5] - p[1], az = p[6] - p[2];
    const float l2 = fmaf(az, az, fmaf(ay, ay, ax * ax));
    d.b = make_float4(ax, ay, az, l2 > 0.f ? 1.0f / l2 : 0.f);
  } else if (c.kind == SB_COLLIDER_BOX) {
    const double x = p[6], y = p[7], z = p[8], w = p[9];
    const double n2 = ((x * x + y * y) + z * z) + w * w;
    double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    if (n2 > 0) {
      const double s = 2.0 / n2;
      R[0][0] = 1.0 - s * (y * y + z * z); R[0][1] = s * (x * y - z * w); R[0][2] = s * (x * z + y * w);
      R[1][0] = s * (x * y + z * w); R[1][1] = 1.0 - s * (x * x + z * z); R[1][2] = s * (y * z - x * w);
      R[2][0] = s * (x * z - y * w); R[2][1] = s * (y * z + x * w); R[2][2] = 1.0 - s * (x * x + y * y);
    }
    d.a.w = 0.f;
    d.b = make_float4((float)R[0][0], (float)R[1][0], (float)R[2][0], p[3]); // box axis k = column k of R
    d.c = make_float4((float)R[0][1], (float)R[1][1], (float)R[2][1], p[4]);
    d.d = make_float4((float)R[0][2], (float)R[1][2], (float)R[2][2], p[5]);
  }
  return d;
}

} // namespace sb

using namespace sb;

struct sb_solver {
  Plan plan;
  sb_params prm{};
  bool on_device = false;
  int device = 0;
  int n_sm = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t cap_stream = nullptr;
  uint32_t block_threads = 64;
  uint64_t dev_bytes = 0;
  std::string err;

  DevBuf<float4> x, v, xp, nrm, stage_a, stage_b;
  DevBuf<float> stage_f;
  DevBuf<uint32_t> inv, surf_slot, surf_tri_off, surf_tri_ids, ghost_slot;
  std::map<int, DevBuf<uint32_t>> halo; // registered halo index lists (device slots)
  // peer-memory exchange state per list: my receive buffer (+ flag word after it), where to send
  struct HaloLink {
    DevBuf<unsigned char> recv; // n * 16 bytes + 16 (flag)
    DevBuf<uint32_t> ctl_send, ctl_recv;
    float4 *peer_buf = nullptr;
    uint32_t *peer_flag = nullptr;
  };
  std::map<int, HaloLink> links;
  DevBuf<int32_t> tris_dev;
  DevBuf<DevParams> dprm;
  DevParams *hprm = nullptr; // pinned
  std::vector<PassBufs> passes;
  // persistent tile-DAG kernel (plan.dag_ok): dependency lists, completion counters + ticket + error word
  DevBuf<uint32_t> dag_dep_off[SB_DAG_MAX_PASSES], dag_dep_list[SB_DAG_MAX_PASSES], dag_ctr;
  DagDev dag{};
  bool dag_ready = false;
  uint32_t dag_smem = 0, dag_grid = 0, dag_bt = 64, dag_width = 1;
  // one mesh over several GPUs through peer memory (sb_dist_setup / sb_dist_connect)
  DistDev dist{};
  DevBuf<DistDev> dist_dev;
  DevBuf<uint32_t> dist_ctl;
  bool dist_on = false;
  uint32_t dist_connected = 0; // bit p: peer p connected
  // the vertices / surface vertices a packed frame carries: all of them, or (distributed) the ones this rank owns,
  // ascending caller id either way
  DevBuf<uint32_t> own_slot, own_surf;
  uint32_t n_own = 0, n_own_surf = 0;
  DevBuf<float4> pack_buf;
  DevBuf<int2> g_edges;
  DevBuf<float> g_elen;
  DevBuf<int4> g_tets;
  DevBuf<float> g_trest;
  // diagnostics (lazy)
  DevBuf<int2> d_edges;
  DevBuf<float> d_elen;
  DevBuf<int4> d_tets;
  DevBuf<double> d_part;
  // pinned staging for read-backs
  float *pin = nullptr;
  size_t pin_bytes = 0;

  float cur_dt = -1.f;
  bool prm_dirty = true;
  sb_collider colliders[SB_MAX_COLLIDERS];
  int n_col = 0;
  uint64_t frames_done = 0; // sb_step calls since creation / the last sb_load_state (stored in snapshots)
  // render mesh bound to the tets (sb_skin_bind)
  uint32_t skin_n = 0, skin_f = 0;
  std::vector<int32_t> skin_tet;
  std::vector<float> skin_bary;
  DevBuf<uint4> skin_slots;
  DevBuf<float4> skin_w, skin_x, skin_nrm;
  DevBuf<uint32_t> skin_tri_off, skin_tri_ids;
  DevBuf<int32_t> skin_tris;
  DevBuf<float> skin_stage;
  std::map<uint64_t, cudaGraphExec_t> graphs;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;

  ~sb_solver() { release_device(); }

  void release_device() {
    if (!on_device) return;
    cudaSetDevice(device);
    if (stream) cudaStreamSynchronize(stream);
    for (auto &g : graphs) cudaGraphExecDestroy(g.second);
    graphs.clear();
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (hprm) cudaFreeHost(hprm);
    if (pin) cudaFreeHost(pin);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (own_stream && stream) cudaStreamDestroy(stream);
    hprm = nullptr; pin = nullptr; ev0 = ev1 = nullptr; cap_stream = nullptr; stream = nullptr;
    on_device = false;
  }

  bool fast() const { return (prm.flags & SB_FLAG_FAST_MATH) != 0; }

  // ---- upload ---------------------------------------------------------------------
  void upload(const sb_mesh_desc &m) {
    device = m.device;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    n_sm = prop.multiProcessorCount;
    on_device = true;
    if (m.stream) {
      stream = (cudaStream_t)m.stream;
    } else {
      CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      own_stream = true;
    }
    CK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&ev0));
    CK(cudaEventCreate(&ev1));
    CK(cudaMallocHost(&hprm, sizeof(DevParams)));
    const uint32_t V = plan.V;
    x.alloc(V, &dev_bytes);
    v.alloc(V, &dev_bytes);
    xp.alloc(V, &dev_bytes);
    stage_a.alloc(V, &dev_bytes);
    stage_b.alloc(V, &dev_bytes);
    stage_f.alloc(3 * (size_t)V, &dev_bytes);
    dprm.alloc(1, &dev_bytes);
    inv.upload(plan.inv, &dev_bytes);
    {
      std::vector<float4> hx(V), hv(V, make_float4(0, 0, 0, 0));
      for (uint32_t d = 0; d < V; d++) {
        uint32_t o = plan.perm[d];
        hx[d] = make_float4(plan.pos[3 * (size_t)o], plan.pos[3 * (size_t)o + 1], plan.pos[3 * (size_t)o + 2], plan.inv_mass[o]);
      }
      CK(cudaMemcpy(x.p, hx.data(), V * sizeof(float4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(xp.p, hx.data(), V * sizeof(float4), cudaMemcpyHostToDevice));
      CK(cudaMemcpy(v.p, hv.data(), V * sizeof(float4), cudaMemcpyHostToDevice));
    }
    if (plan.n_ghost) {
      std::vector<uint32_t> gs(plan.n_ghost);
      for (uint32_t k = 0; k < plan.n_ghost; k++) gs[k] = plan.inv[V - plan.n_ghost + k];
      ghost_slot.upload(gs, &dev_bytes);
      k_mark_ghosts<<<grid_for(plan.n_ghost, 256), 256>>>(plan.n_ghost, ghost_slot.p, v.p);
      CK(cudaGetLastError());
    }
    // surface
    {
      const size_t ns = plan.surf_ids.size();
      std::vector<uint32_t> slot(ns);
      for (size_t s = 0; s < ns; s++) slot[s] = plan.inv[plan.surf_ids[s]];
      surf_slot.upload(slot, &dev_bytes);
      surf_tri_off.upload(plan.surf_tri_off, &dev_bytes);
      surf_tri_ids.upload(plan.surf_tri_ids, &dev_bytes);
      std::vector<int32_t> td(plan.tris.size());
      for (size_t i = 0; i < td.size(); i++) td[i] = (int32_t)plan.inv[plan.tris[i]];
      tris_dev.upload(td, &dev_bytes);
      nrm.alloc(ns ? ns : 1, &dev_bytes);
      CK(cudaMemset(nrm.p, 0, (ns ? ns : 1) * sizeof(float4)));
    }
    // tile passes
    passes.resize(plan.passes.size());
    for (size_t k = 0; k < plan.passes.size(); k++) {
      const TilePass &tp = plan.passes[k];
      PassBufs &pb = passes[k];
      pb.vert_off.upload(tp.vert_off, &dev_bytes);
      if (!tp.contiguous) pb.tile_verts.upload(tp.tile_verts, &dev_bytes);
      {
        std::vector<uint4> r4(tp.rounds.size());
        for (size_t i = 0; i < r4.size(); i++) r4[i] = make_uint4(tp.rounds[i].x, tp.rounds[i].y, tp.rounds[i].z, tp.rounds[i].w);
        pb.rounds.upload(r4, &dev_bytes);
      }
      pb.stream.upload(tp.stream, &dev_bytes);
      pb.aux.upload(tp.aux, &dev_bytes);
      const uint32_t pos_bytes = (tp.max_tile_verts * 16u + 127u) & ~127u;
      pb.smem = pos_bytes + 16u; // positions + the mbarrier their bulk copies complete on
      if (pb.smem > (uint32_t)prop.sharedMemPerBlockOptin) throw std::string("tile_cap exceeds the shared memory of this device");
      // bulk copies per run pay off when runs are long; else threads gather vertex by vertex
      const bool use_runs = !tp.contiguous && !tp.tile_verts.empty() &&
                            (double)tp.tile_verts.size() / (double)(tp.runs.size() - tp.n_tiles()) >= 8.0;
      if (use_runs) {
        pb.run_off.upload(tp.run_off, &dev_bytes);
        pb.runs.upload(tp.runs, &dev_bytes);
      }
      pb.order.upload(tp.launch_order, &dev_bytes);
      pb.dev = PassDev{pb.order.p, pb.vert_off.p, tp.contiguous ? nullptr : pb.tile_verts.p, use_runs ? pb.run_off.p : nullptr,
                       use_runs ? pb.runs.p : nullptr, pb.rounds.p, nullptr, {nullptr}, 0u, 0u, reinterpret_cast<const uint4 *>(pb.stream.p),
                       pb.aux.p, tp.n_tiles(), pos_bytes, nullptr, nullptr, 1u, 1u, 0u, 0u, nullptr, nullptr, 0u, 0u};
      pb.grid = tp.n_tiles();
      build_desc(k, tp.launch_order);
      pb.bt = tp.bt;
      pb.width = tp.width;
      pb.empty = tp.n_edges + tp.n_tets == 0;
      if (k == 0) block_threads = tp.bt;
    }
    uint32_t max_smem = 0;
    for (auto &pb : passes) max_smem = std::max(max_smem, pb.smem);
    set_smem_attr(max_smem);
    setup_dag(max_smem);
    g_edges.upload(plan.g_edges, &dev_bytes);
    g_elen.upload(plan.g_elen, &dev_bytes);
    g_tets.upload(plan.g_tets, &dev_bytes);
    g_trest.upload(plan.g_trest, &dev_bytes);
    d_part.alloc((size_t)kDiagBlocks * 16, &dev_bytes);
    CK(cudaDeviceSynchronize());
  }

  // One record per CTA of pass k, in launch order (k_tile_rounds reads nothing else about its tile).
  // One record per CTA of pass k, in launch order (k_tile_rounds reads nothing else about its tile).
  // Distributed (`tuple` = runner ranks per device vertex, dist_layout): every tile -- of a contiguous pass too -- gets a
  // run list whose runs are cut where the tuple changes, with the tuple above the local offset (kernels.cuh: a tile
  // stores each run into the array of the rank that touches it next).
  void build_desc(size_t k, const std::vector<uint32_t> &order, const std::vector<uint32_t> *tuple = nullptr) {
    const TilePass &tp = plan.passes[k];
    PassBufs &pb = passes[k];
    std::vector<uint32_t> roff;
    if (tuple) {
      std::vector<uint2> ro;
      roff.assign((size_t)tp.n_tiles() + 1, 0);
      auto emit = [&](uint32_t first, uint32_t len, uint32_t local) { // [first, first + len) at local offset `local`
        uint32_t a = 0;
        while (a < len) {
          uint32_t b = a + 1;
          while (b < len && (*tuple)[first + b] == (*tuple)[first + a]) b++;
          ro.push_back(make_uint2(first + a, (local + a) | ((*tuple)[first + a] << 16)));
          a = b;
        }
      };
      for (uint32_t t = 0; t < tp.n_tiles(); t++) {
        roff[t] = (uint32_t)ro.size();
        const uint32_t nv_t = tp.vert_off[t + 1] - tp.vert_off[t];
        if (nv_t > 0xffffu) throw std::string("tile too large for a distributed mesh");
        if (tp.contiguous) {
          emit(tp.vert_off[t], nv_t, 0);
        } else { // maximal ranges of consecutive device ids in the tile's (ascending) vertex list
          const uint32_t *tv = tp.tile_verts.data() + tp.vert_off[t];
          for (uint32_t a = 0; a < nv_t;) {
            uint32_t b = a + 1;
            while (b < nv_t && tv[b] == tv[b - 1] + 1) b++;
            emit(tv[a], b - a, a);
            a = b;
          }
        }
        ro.push_back(make_uint2(0, nv_t)); // closes the tile's list
      }
      roff[tp.n_tiles()] = (uint32_t)ro.size();
      pb.runs.upload(ro, &dev_bytes);
      pb.run_off.upload(roff, &dev_bytes);
      pb.dev.runs = pb.runs.p;
      pb.dev.run_off = pb.run_off.p;
    }
    const bool runs = pb.dev.run_off != nullptr;
    const std::vector<uint32_t> &ro_off = tuple ? roff : tp.run_off;
    std::vector<uint4> d(2 * order.size() + 2, make_uint4(0, 0, 0, 0));
    for (size_t j = 0; j < order.size(); j++) {
      const uint32_t t = order[j];
      d[2 * j] = make_uint4(tp.vert_off[t], tp.vert_off[t + 1] - tp.vert_off[t], runs ? ro_off[t] : 0u, runs ? ro_off[t + 1] - ro_off[t] - 1u : 0u);
      d[2 * j + 1] = make_uint4(tp.rounds[t].x, tp.rounds[t].y, tp.rounds[t].z, tp.rounds[t].w);
    }
    pb.desc.upload(d, &dev_bytes);
    pb.dev.desc = pb.desc.p;
  }

  template <bool FAST, int BT, int W16>
  static void set_attr_one(uint32_t smem) {
    CK(cudaFuncSetAttribute(k_tile_rounds<FAST, BT, W16, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_tile_rounds<FAST, BT, W16, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if constexpr (!FAST) CK(cudaFuncSetAttribute(k_tile_rounds<FAST, BT, W16, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  template <bool FAST>
  static void set_attr_math(uint32_t smem) {
    set_attr_one<FAST, 32, 1>(smem); set_attr_one<FAST, 64, 1>(smem); set_attr_one<FAST, 128, 1>(smem); set_attr_one<FAST, 256, 1>(smem);
    set_attr_one<FAST, 160, 1>(smem); set_attr_one<FAST, 192, 1>(smem); set_attr_one<FAST, 160, 2>(smem); set_attr_one<FAST, 192, 2>(smem);
    set_attr_one<FAST, 32, 2>(smem); set_attr_one<FAST, 64, 2>(smem); set_attr_one<FAST, 128, 2>(smem); set_attr_one<FAST, 256, 2>(smem);
  }
  void set_smem_attr(uint32_t smem) {
    if (smem <= 48 * 1024) return;
    set_attr_math<false>(smem);
    set_attr_math<true>(smem);
  }
  // ---- parameters -----------------------------------------------------------------
  void refresh_params(float dt) {
    if (!prm_dirty && dt == cur_dt) return;
    DevParams p{};
    // the oracle's make_consts, in float
    p.h = dt / (float)prm.substeps;
    p.inv_h = 1.0f / p.h;
    const float hh = p.h * p.h;
    auto compliance = [](float k) -> float {
      if (std::isinf(k) && k > 0) return 0.0f;
      if (k > 0) return 1.0f / k;
      return -1.0f;
    };
    const float cd = compliance(prm.stiffness_distance), cv = compliance(prm.stiffness_volume);
    p.use_d = cd >= 0;
    p.use_v = cv >= 0;
    p.a_d = p.use_d ? cd / hh : 0.f;
    p.a_v36 = p.use_v ? 36.0f * (cv / hh) : 0.f;
    const float dm = 1.0f - p.h * prm.damping;
    p.damp = dm > 0 ? dm : 0.f;
    p.keep = 1.0f - prm.friction;
    p.gx = prm.gravity[0]; p.gy = prm.gravity[1]; p.gz = prm.gravity[2];
    p.ground_y = prm.ground_y;
    p.use_ground = !(prm.flags & SB_FLAG_NO_GROUND);
    p.n_col = n_col;
    for (int s = 0; s < n_col; s++) {
      p.col_kind[s] = colliders[s].kind;
      p.col_fric[s] = colliders[s].friction;
      p.col[s] = prepare_collider(colliders[s]);
    }
    CK(cudaStreamSynchronize(stream)); // hprm may still be in flight from an earlier change
    *hprm = p;
    CK(cudaMemcpyAsync(dprm.p, hprm, sizeof(DevParams), cudaMemcpyHostToDevice, stream));
    cur_dt = dt;
    prm_dirty = false;
  }

  // ---- launches -------------------------------------------------------------------
  int grid_for(size_t n, int bt) const {
    size_t g = (n + bt - 1) / bt;
    size_t cap = (size_t)n_sm * 8;
    return (int)std::max<size_t>(1, std::min(g, cap));
  }

  template <bool FAST, int BT, int W16>
  void launch_tile_cfg(const PassBufs &pb, const PassDev &dev, cudaStream_t s) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(pb.grid);
    cfg.blockDim = dim3(BT);
    cfg.dynamicSmemBytes = pb.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (prm.flags & SB_FLAG_NO_PDL) ? 0 : 1;
    if constexpr (!FAST) {
      if (dev.trace) {
        CK(cudaLaunchKernelEx(&cfg, k_tile_rounds<FAST, BT, W16, true, false>, dev, x.p, (const DevParams *)dprm.p));
        return;
      }
    }
    if (dev.n_seg > 1 || dev.reps > 1 || dev.pre || dev.post)
      CK(cudaLaunchKernelEx(&cfg, k_tile_rounds<FAST, BT, W16, false, true>, dev, x.p, (const DevParams *)dprm.p));
    else
      CK(cudaLaunchKernelEx(&cfg, k_tile_rounds<FAST, BT, W16, false, false>, dev, x.p, (const DevParams *)dprm.p));
  }
  template <bool FAST, int W16>
  void launch_tile_w(const PassBufs &pb, const PassDev &dev, cudaStream_t s) {
    switch (pb.bt) {
      case 32: launch_tile_cfg<FAST, 32, W16>(pb, dev, s); break;
      case 64: launch_tile_cfg<FAST, 64, W16>(pb, dev, s); break;
      case 128: launch_tile_cfg<FAST, 128, W16>(pb, dev, s); break;
      case 160: launch_tile_cfg<FAST, 160, W16>(pb, dev, s); break;
      case 192: launch_tile_cfg<FAST, 192, W16>(pb, dev, s); break;
      default: launch_tile_cfg<FAST, 256, W16>(pb, dev, s); break;
    }
  }
  // One launch of tile pass `pb`: n_seg segments (substeps) of `reps` repetitions of every tile's rounds, with the
  // substep boundaries between segments -- and predict before / finish after, if asked -- done on the tiles.
  template <bool FAST>
  void launch_tile(const PassBufs &pb, cudaStream_t s, uint32_t n_seg = 1, uint32_t reps = 1, bool pre = false, bool post = false, uint32_t next_pass = 0,
                   uint32_t next_full_pass = 0) {
    if (pb.empty && !(pre || post || n_seg > 1)) return;
    if (!pb.grid) {
      // distributed: this rank has no tile in the pass, but its epoch moves with every launch of the sequence
      if (dist_on && !pb.empty) k_dist_bump<<<1, 32, 0, s>>>(dist_dev.p);
      return;
    }
    PassDev dev = pb.dev;
    dev.n_seg = n_seg; dev.reps = reps; dev.pre = pre; dev.post = post; dev.next_pass = next_pass; dev.next_full_pass = next_full_pass;
    if (dist_on) for (uint32_t r = 0; r < dist.n_ranks; r++) dev.xs[r] = dist.x_of[r];
    else dev.xs[0] = x.p;
    dev.v = v.p; dev.xp = xp.p;
    if (pb.width == 2) launch_tile_w<FAST, 2>(pb, dev, s);
    else launch_tile_w<FAST, 1>(pb, dev, s);
  }
  // ---- persistent tile-DAG kernel -----------------------------------------------------
  template <bool FAST, int BT, int W16>
  void dag_config(bool launch, cudaStream_t s) {
    auto kern = k_tile_dag<FAST, BT, W16>;
    if (!launch) {
      if (dag_smem > 48 * 1024) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dag_smem));
      int per_sm = 0;
      CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BT, dag_smem));
      dag_grid = std::max(dag_grid, (uint32_t)std::max(1, per_sm) * (uint32_t)n_sm);
      return;
    }
    DagDev g = dag;
    g.iterations = (uint32_t)prm.iterations;
    const uint64_t total = (uint64_t)g.tile_base[g.n_pass] * g.iterations;
    if (!total) return;
    const uint32_t grid = (uint32_t)std::min<uint64_t>(dag_grid, total);
    kern<<<grid, BT, dag_smem, s>>>(g, x.p, (const DevParams *)dprm.p);
  }
  template <bool FAST>
  void dag_dispatch(bool launch, cudaStream_t s) {
#define SB_DAG_CASE(BT_, W_) if (dag_bt == BT_ && dag_width == W_) return dag_config<FAST, BT_, W_>(launch, s)
    SB_DAG_CASE(32, 1); SB_DAG_CASE(64, 1); SB_DAG_CASE(128, 1); SB_DAG_CASE(256, 1);
    SB_DAG_CASE(32, 2); SB_DAG_CASE(64, 2); SB_DAG_CASE(128, 2); SB_DAG_CASE(256, 2);
#undef SB_DAG_CASE
  }
  void setup_dag(uint32_t max_smem) {
    dag_ready = false;
    const size_t np = plan.passes.size();
    if (!plan.dag_ok || np < 2 || np > SB_DAG_MAX_PASSES) return;
    for (size_t k = 0; k < np; k++)
      if (plan.passes[k].bt != plan.passes[0].bt || plan.passes[k].width != plan.passes[0].width ||
          plan.passes[k].dep_off.size() != (size_t)plan.passes[k].n_tiles() + 1)
        return;
    dag = DagDev{};
    uint32_t pos_bytes = 0;
    for (size_t k = 0; k < np; k++) pos_bytes = std::max(pos_bytes, passes[k].dev.pos_bytes);
    for (size_t k = 0; k < np; k++) {
      dag_dep_off[k].upload(plan.passes[k].dep_off, &dev_bytes);
      dag_dep_list[k].upload(plan.passes[k].dep_list, &dev_bytes);
      dag.pass[k] = passes[k].dev;
      // tickets walk a pass in tile order (spatial), not in the balanced launch order: a tile's dependencies
      // then finish at about the same point of the previous pass as the tile itself starts in this one
      // (measured: heaviest-first tickets turn the end of every pass into a global barrier, 9.3 -> 10.8 ms)
      dag.pass[k].order = nullptr;
      dag.pass[k].pos_bytes = pos_bytes; // one barrier address for the whole run
      dag.dep_off[k] = dag_dep_off[k].p;
      dag.dep_list[k] = dag_dep_list[k].p;
      dag.tile_base[k + 1] = dag.tile_base[k] + plan.passes[k].n_tiles();
    }
    dag.n_pass = (uint32_t)np;
    dag_ctr.alloc((size_t)dag.tile_base[np] + 2, &dev_bytes);
    CK(cudaMemset(dag_ctr.p, 0, ((size_t)dag.tile_base[np] + 2) * sizeof(uint32_t)));
    dag.done = dag_ctr.p;
    dag.ticket = dag_ctr.p + dag.tile_base[np];
    dag.error = dag_ctr.p + dag.tile_base[np] + 1;
    dag_bt = plan.passes[0].bt;
    dag_width = plan.passes[0].width;
    dag_smem = std::max(max_smem, pos_bytes + 16u);
    dag_grid = 0;
    dag_dispatch<false>(false, nullptr);
    dag_dispatch<true>(false, nullptr);
    dag_ready = dag_grid > 0;
  }
  bool use_dag() const { return dag_ready && (prm.flags & SB_FLAG_DAG) && !halo_active() && !dist_on && prm.iterations > 0; }
  // all passes of all iterations of one substep: counters and ticket (not the error word) are cleared first
  void launch_dag(cudaStream_t s) {
    CK(cudaMemsetAsync(dag_ctr.p, 0, ((size_t)dag.tile_base[dag.n_pass] + 1) * sizeof(uint32_t), s));
    if (fast()) dag_dispatch<true>(true, s);
    else dag_dispatch<false>(true, s);
  }
  // sticky error words of the multi-GPU waits (epoch of a peer, halo sequence number): 1 if any timed out
  int dist_error_word() {
    uint32_t e = 0, bad = 0;
    if (dist.ctl) {
      CK(cudaMemcpyAsync(&e, dist_ctl.p + 2, sizeof e, cudaMemcpyDeviceToHost, stream));
      CK(cudaStreamSynchronize(stream));
      bad |= e;
    }
    for (auto &kv : links)
      if (kv.second.ctl_recv.p) {
        CK(cudaMemcpyAsync(&e, kv.second.ctl_recv.p + 2, sizeof e, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        bad |= e;
      }
    return (int)bad;
  }
  int dag_error() {
    if (!dag_ready) return 0;
    uint32_t e = 0;
    CK(cudaMemcpyAsync(&e, dag.error, sizeof e, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    return (int)e;
  }

  // ---- one mesh over several GPUs ------------------------------------------------------
  // Host-only part: the slab of the device numbering a rank owns, the tiles of every pass it runs (zone tiles first)
  // and how many of them are zone tiles.
  void dist_layout(int rank, int n_ranks, DistDev &D, std::vector<std::vector<uint32_t>> &tiles, std::vector<uint32_t> &n_zone,
                   std::vector<uint32_t> *runner_tuple = nullptr) const {
    // (tiles[k] = this rank's tiles of pass k, the first n_zone[k] of them zone tiles)
    if (n_ranks < 2 || n_ranks > SB_MAX_RANKS || rank < 0 || rank >= n_ranks) throw std::string("rank / n_ranks out of range (2..8 ranks)");
    if (plan.n_tilings < 2 || plan.n_ghost || !plan.gbatches.empty() || plan.passes.empty() || plan.passes.size() > 5)
      throw std::string("a distributed mesh must be planned as balanced shifted tilings (one big component, no ghosts, no global colour batches, at most one leftover pass)");
    const TilePass &t0 = plan.passes[0];
    if (!t0.contiguous || t0.n_tiles() < (uint32_t)n_ranks) throw std::string("fewer tiles than ranks");
    D = DistDev{};
    D.n_ranks = (uint32_t)n_ranks;
    D.rank = (uint32_t)rank;
    if (plan.dist_ranks == (uint32_t)n_ranks) {
      // the planner numbered the boxes of the unshifted tiling rank by rank (blocks of boxes, PlanOptions::dist_ranks)
      for (int r = 0; r <= n_ranks; r++) D.slab_lo[r] = plan.dist_slab_lo[r];
    } else {
      // slabs: consecutive tiles of the unshifted tiling (consecutive device ids), equal vertex counts
      uint32_t prev_tile = 0;
      for (int r = 1; r < n_ranks; r++) {
        const uint64_t target = (uint64_t)plan.V * r / n_ranks;
        uint32_t tile = (uint32_t)(std::lower_bound(t0.vert_off.begin(), t0.vert_off.end() - 1, (uint32_t)target) - t0.vert_off.begin());
        tile = std::max(tile, prev_tile + 1);
        tile = std::min(tile, t0.n_tiles() - (uint32_t)(n_ranks - r));
        D.slab_lo[r] = t0.vert_off[tile];
        prev_tile = tile;
      }
      D.slab_lo[0] = 0;
      D.slab_lo[n_ranks] = plan.V;
    }
    for (int r = n_ranks + 1; r <= SB_MAX_RANKS; r++) D.slab_lo[r] = plan.V;
    auto owner_of = [&](uint32_t dev) {
      uint32_t r = 0;
      while (r + 1 < D.n_ranks && dev >= D.slab_lo[r + 1]) r++;
      return r;
    };
    // Who runs a tile: the rank that owns most of its vertices (ties: the lower rank).  (Dealing a spanning tile
    // to the less loaded of its ranks instead was measured slower: more of its runs become remote.)
    const size_t np = plan.passes.size();
    std::vector<std::vector<uint8_t>> runner(np);
    std::vector<std::vector<uint32_t>> tile_of(np, std::vector<uint32_t>(plan.V, 0xffffffffu));
    for (size_t k = 0; k < np; k++) {
      const TilePass &tp = plan.passes[k];
      runner[k].assign(tp.n_tiles(), 0xff);
      for (uint32_t t = 0; t < tp.n_tiles(); t++) {
        const uint32_t a = tp.vert_off[t], b = tp.vert_off[t + 1];
        if (a == b) continue;
        uint32_t cnt[SB_MAX_RANKS] = {0};
        if (tp.contiguous) {
          cnt[owner_of(a)] = b - a;
          for (uint32_t i = a; i < b; i++) tile_of[k][i] = t;
        } else {
          for (uint32_t i = a; i < b; i++) {
            cnt[owner_of(tp.tile_verts[i])]++;
            tile_of[k][tp.tile_verts[i]] = t;
          }
        }
        uint32_t best = 0;
        for (uint32_t r = 1; r < D.n_ranks; r++)
          if (cnt[r] > cnt[best]) best = r;
        runner[k][t] = (uint8_t)best;
      }
    }
    // Zone: a tile with a vertex that some tile of ANOTHER pass, run by ANOTHER rank, holds as well.  Everything else
    // is interior: all its vertices are touched by this rank's tiles only, in every pass.  (A tile with a run in a
    // peer's memory is a zone tile: the pass-0 tile of that run belongs to the peer.)
    std::vector<std::vector<uint8_t>> in_zone(np);
    for (size_t k = 0; k < np; k++) in_zone[k].assign(plan.passes[k].n_tiles(), 0);
    uint32_t nbr = 0;
    for (uint32_t d = 0; d < plan.V; d++) {
      uint32_t ranks = 0;
      for (size_t k = 0; k < np; k++)
        if (tile_of[k][d] != 0xffffffffu) ranks |= 1u << runner[k][tile_of[k][d]];
      if (ranks & (ranks - 1)) { // more than one rank touches this vertex
        for (size_t k = 0; k < np; k++)
          if (tile_of[k][d] != 0xffffffffu) in_zone[k][tile_of[k][d]] = 1;
        if (ranks >> rank & 1u) nbr |= ranks;
      }
    }
    // Surface normals read across a cut: the normals launch of the rank that owns a surface vertex reads the other
    // vertices of its triangles where they live, in their owners' arrays (k_normals_dist).  The owner's last tile
    // launch of the frame must have published such a vertex before the reader starts, and its first launch of the
    // next frame must wait for the reader: the tiles of all three vertices of a triangle whose vertices have different
    // owners are zone tiles, and those owners neighbours.  (With cuts along box faces the shifted tilings make these
    // vertices two-rank vertices anyway -- nothing changes for a lattice block; this makes it so by construction.)
    // (Neighbours: every rank that touches one of the three vertices in any pass -- the last holder of a vertex, which
    // stores it home, need not be its owner.)
    // TEST ONLY (reopens the race described above): lets the CPU test-suite show that sb_dist_verify reports it
    const bool no_cut_zones = getenv("SB_DEBUG_NO_CUT_TRIANGLE_ZONES") != nullptr;
    for (size_t f = 0; f + 2 < plan.tris.size(); f += 3) {
      uint32_t d[3], own = 0, touch = 0;
      for (int j = 0; j < 3; j++) {
        d[j] = plan.inv[(size_t)plan.tris[f + j]];
        own |= 1u << owner_of(d[j]);
        for (size_t k = 0; k < np; k++)
          if (tile_of[k][d[j]] != 0xffffffffu) touch |= 1u << runner[k][tile_of[k][d[j]]];
      }
      if (!(own & (own - 1))) continue;
      if (no_cut_zones) continue;
      for (int j = 0; j < 3; j++)
        for (size_t k = 0; k < np; k++)
          if (tile_of[k][d[j]] != 0xffffffffu) in_zone[k][tile_of[k][d[j]]] = 1;
      if (touch >> rank & 1u) nbr |= touch;
    }
    D.nbr_mask = nbr & ~(1u << rank);
    if (runner_tuple) { // per device vertex: the rank that runs its tile in pass k, three bits per pass (pass 0 = its slab's rank)
      if (np > 5) throw std::string("a distributed mesh has at most five tile passes");
      // (only the last pass may leave vertices out -- a leftover pass after the tilings; bit 15: this vertex is in it)
      runner_tuple->assign(plan.V, 0u);
      for (uint32_t d = 0; d < plan.V; d++) {
        uint32_t tup = 0;
        for (size_t k = 0; k < np; k++) {
          if (tile_of[k][d] != 0xffffffffu) tup |= (uint32_t)runner[k][tile_of[k][d]] << (3 * k) | (k + 1 == np && np > plan.n_tilings ? 1u << 15 : 0u);
          else if (k + 1 != np || np <= plan.n_tilings) throw std::string("a distributed mesh needs tilings that stage every vertex (SB_WHOLE_BOXES)");
        }
        if ((tup & 7u) != owner_of(d)) throw std::string("the unshifted tiling's tiles must lie within one slab each");
        (*runner_tuple)[d] = tup;
      }
    }
    // this rank's tiles of every pass: zone tiles first, each part heaviest first and dealt to the SMs in a snake
    // like the single-GPU launch order
    tiles.assign(np, {});
    n_zone.assign(np, 0);
    for (size_t k = 0; k < np; k++) {
      const TilePass &tp = plan.passes[k];
      std::vector<uint32_t> part[2];
      for (uint32_t t = 0; t < tp.n_tiles(); t++) {
        if (runner[k][t] != (uint8_t)rank) continue;
        // (a tile without constraints in this pass runs all the same: it hands its vertices on to the ranks that hold
        // them in the next pass)
        if (tp.vert_off[t + 1] == tp.vert_off[t]) continue;
        part[in_zone[k][t] ? 0 : 1].push_back(t);
      }
      auto work = [&](uint32_t t) { return ((uint64_t)(tp.rounds[t].y + tp.rounds[t].z) << 32) | (uint32_t)(tp.ent_off[t + 1] - tp.ent_off[t]); };
      const size_t layer = (size_t)std::max(1, n_sm);
      for (auto &mine : part) {
        std::stable_sort(mine.begin(), mine.end(), [&](uint32_t a, uint32_t b) { return work(a) > work(b); });
        for (size_t lo = layer; lo < mine.size(); lo += 2 * layer) std::reverse(mine.begin() + lo, mine.begin() + std::min(mine.size(), lo + layer));
      }
      n_zone[k] = (uint32_t)part[0].size();
      tiles[k] = part[0];
      tiles[k].insert(tiles[k].end(), part[1].begin(), part[1].end());
    }
  }
  void dist_setup(int rank, int n_ranks) {
    if (halo_active()) throw std::string("halo lists and the peer-memory distribution are alternatives");
    std::vector<std::vector<uint32_t>> tiles;
    std::vector<uint32_t> n_zone;
    std::vector<uint32_t> tuple;
    dist_layout(rank, n_ranks, dist, tiles, n_zone, &tuple);
    for (size_t k = 0; k < plan.passes.size(); k++) {
      passes[k].order.upload(tiles[k], &dev_bytes);
      passes[k].dev.order = passes[k].order.p;
      build_desc(k, tiles[k], &tuple);
      passes[k].dev.n_zone = n_zone[k];
      passes[k].grid = (uint32_t)tiles[k].size();
    }
    {
      std::vector<uint32_t> slots, surf;
      const uint32_t lo = dist.slab_lo[rank], hi = dist.slab_lo[rank + 1];
      for (uint32_t c = 0; c < plan.V; c++)
        if (plan.inv[c] >= lo && plan.inv[c] < hi) slots.push_back(plan.inv[c]);
      for (size_t s = 0; s < plan.surf_ids.size(); s++) {
        const uint32_t d = plan.inv[plan.surf_ids[s]];
        if (d >= lo && d < hi) surf.push_back((uint32_t)s);
      }
      n_own = (uint32_t)slots.size();
      n_own_surf = (uint32_t)surf.size();
      own_slot.upload(slots, &dev_bytes);
      own_surf.upload(surf, &dev_bytes);
      pack_buf.release();
    }
    dist_ctl.alloc(4 + SB_MAX_RANKS, &dev_bytes);
    CK(cudaMemset(dist_ctl.p, 0, (4 + SB_MAX_RANKS) * sizeof(uint32_t)));
    dist.ctl = dist_ctl.p;
    dist.x_of[rank] = x.p;
    dist_connected = 1u << rank;
    dist_on = false; // until every peer is connected
  }
  void dist_connect(int peer, void *peer_x, void *peer_ctl) {
    if (!dist.ctl) throw std::string("call sb_dist_setup first");
    if (peer < 0 || peer >= (int)dist.n_ranks || peer == (int)dist.rank || !peer_x || !peer_ctl) throw std::string("bad peer");
    dist.x_of[peer] = (float4 *)peer_x;
    dist.peer_flag[peer] = (uint32_t *)peer_ctl + 4 + dist.rank;
    dist_connected |= 1u << peer;
    if (dist_connected == (1u << dist.n_ranks) - 1u) {
      dist_dev.alloc(1, &dev_bytes);
      CK(cudaMemcpy(dist_dev.p, &dist, sizeof dist, cudaMemcpyHostToDevice));
      for (auto &pb : passes) pb.dev.dist = dist_dev.p;
      dist_on = true;
      for (auto &g : graphs) cudaGraphExecDestroy(g.second);
      graphs.clear();
    }
  }

  void launch_pass(size_t k, cudaStream_t s, uint32_t n_seg = 1, uint32_t reps = 1, bool pre = false, bool post = false, uint32_t next_pass = 0,
                   uint32_t next_full_pass = 0) {
    if (fast()) launch_tile<true>(passes[k], s, n_seg, reps, pre, post, next_pass, next_full_pass);
    else launch_tile<false>(passes[k], s, n_seg, reps, pre, post, next_pass, next_full_pass);
  }
  void launch_global(cudaStream_t s, int group = -1) {
    for (const GlobalBatch &b : plan.gbatches) {
      if (!b.cnt || (group >= 0 && b.group != group)) continue;
      int g = grid_for(b.cnt, 256);
      if (b.tet) {
        if (fast()) k_global_tets<true><<<g, 256, 0, s>>>(g_tets.p + b.off, g_trest.p + b.off, b.cnt, x.p, dprm.p);
        else k_global_tets<false><<<g, 256, 0, s>>>(g_tets.p + b.off, g_trest.p + b.off, b.cnt, x.p, dprm.p);
      } else {
        if (fast()) k_global_edges<true><<<g, 256, 0, s>>>(g_edges.p + b.off, g_elen.p + b.off, b.cnt, x.p, dprm.p);
        else k_global_edges<false><<<g, 256, 0, s>>>(g_edges.p + b.off, g_elen.p + b.off, b.cnt, x.p, dprm.p);
      }
    }
  }
  void launch_group(int group, cudaStream_t s) {
    for (size_t k = 0; k < passes.size(); k++)
      if (plan.passes[k].group == group) launch_pass(k, s);
    launch_global(s, group);
  }
  bool halo_active() const { return !links.empty(); }
  void halo_send(int list, cudaStream_t s) {
    auto li = links.find(list);
    auto hl = halo.find(list);
    if (li == links.end() || hl == halo.end() || !li->second.peer_buf || !hl->second.n) return;
    const uint32_t n = (uint32_t)hl->second.n;
    k_halo_send<<<std::min(grid_for(n, 256), 32), 256, 0, s>>>(n, hl->second.p, x.p, li->second.peer_buf, li->second.peer_flag,
                                                              li->second.ctl_send.p);
  }
  void halo_recv(int list, cudaStream_t s) {
    auto li = links.find(list);
    auto hl = halo.find(list);
    if (li == links.end() || hl == halo.end() || !li->second.recv.p || !hl->second.n) return;
    const uint32_t n = (uint32_t)hl->second.n;
    k_halo_recv<<<std::min(grid_for(n, 256), 32), 256, 0, s>>>(n, hl->second.p, x.p, (const float4 *)li->second.recv.p,
                                                              (const uint32_t *)(li->second.recv.p + (size_t)n * 16), li->second.ctl_recv.p);
  }
  // exchange A: lower-boundary vertices (list 1) go down to the rank that holds them as ghosts (list 0)
  // exchange B: the ghost values (list 0) go back up and overwrite the owner's copies (list 1)
  void exchange(int phase, cudaStream_t s) {
    if (phase == 0) { halo_send(1, s); halo_recv(0, s); }
    else { halo_send(0, s); halo_recv(1, s); }
  }
  // distributed: only this rank's slab of the device numbering is integrated here
  void launch_predict(cudaStream_t s) {
    const uint32_t lo = dist_on ? dist.slab_lo[dist.rank] : 0u, hi = dist_on ? dist.slab_lo[dist.rank + 1] : plan.V;
    k_predict<<<grid_for(hi - lo, 256), 256, 0, s>>>(lo, hi, x.p, v.p, xp.p, dprm.p, dist_on ? dist_dev.p : nullptr);
  }
  void launch_finish(cudaStream_t s) {
    const uint32_t lo = dist_on ? dist.slab_lo[dist.rank] : 0u, hi = dist_on ? dist.slab_lo[dist.rank + 1] : plan.V;
    k_finish<<<grid_for(hi - lo, 256), 256, 0, s>>>(lo, hi, x.p, v.p, xp.p, dprm.p, dist_on ? dist_dev.p : nullptr);
  }
  void launch_normals(cudaStream_t s) {
    const uint32_t ns = (uint32_t)plan.surf_ids.size();
    if (!ns || (prm.flags & SB_FLAG_NO_NORMALS)) {
      // distributed: the frame still ends with a launch of the epoch sequence -- one that waits for the neighbours'
      // last tile launch, whose stores bring this rank's vertices home (a read-back after the frame needs them)
      if (dist_on) k_dist_sync<<<1, 32, 0, s>>>(dist_dev.p);
      return;
    }
    if (dist_on) { // the surface vertices this rank owns; a launch of the epoch sequence on every rank
      if (n_own_surf) k_normals_dist<<<grid_for(n_own_surf, 256), 256, 0, s>>>(n_own_surf, own_surf.p, surf_tri_off.p, surf_tri_ids.p, tris_dev.p, nrm.p, dist_dev.p);
      else k_dist_bump<<<1, 32, 0, s>>>(dist_dev.p);
      return;
    }
    k_normals<<<grid_for(ns, 256), 256, 0, s>>>(ns, surf_tri_off.p, surf_tri_ids.p, tris_dev.p, x.p, nrm.p);
  }

  // ---- the frame program ---------------------------------------------------------------
  //
  // One frame = substeps x { predict, iterations x sweep, finish } + normals, as a flat list of launches.
  //
  // Snake: the tile passes of a sweep are independent sets of constraints in SOME order; iteration 0, 2, 4 ... of a
  // substep runs them forwards (pass 0, 1, ... n-1), iteration 1, 3, ... backwards (a symmetric Gauss-Seidel sweep),
  // so the last pass of one iteration and the first of the next are the SAME pass, on the same tiles.
  // Fusion: consecutive occurrences of one pass become one launch that keeps the tiles' positions in shared memory:
  // within a substep the rounds are simply repeated (`reps`); across a substep boundary (possible when the pass is a
  // contiguous one, whose tiles partition the vertices) the launch also carries collide + velocity update + predict of
  // the tile's own vertices (`n_seg` segments), so that no separate per-vertex kernel runs there.  The first launch
  // of a frame predicts (`pre`) and the last one finishes (`post`) in the same way.  With 4 tilings and 10 iterations
  // that is 30 launches per substep instead of 42, and a batch of bodies that fit one tile each runs a whole frame
  // (substeps x iterations sweeps) in ONE launch.
  // The order is what sb_get_schedule / sb_get_schedule_odd export, and the oracle replays exactly that.
  struct Launch {
    enum Kind { PREDICT, FINISH, PASS, GLOBAL, GROUP, EXCHANGE, NORMALS, DAG } kind;
    int arg = 0;                  // pass index / constraint group / exchange phase
    uint32_t n_seg = 1, reps = 1; // PASS
    bool pre = false, post = false;
    uint32_t next_pass = 0, next_full_pass = 0; // PASS: the pass of the next launch that touches positions (0 also stands for the vertices' home)
  };
  // every constraint sits in a tile pass of group 0: any pass order is a Gauss-Seidel order, and passes can be fused
  bool pure() const {
    if (!plan.gbatches.empty() || plan.n_ghost || plan.passes.empty()) return false;
    for (const TilePass &tp : plan.passes)
      if (tp.group != 0) return false;
    return true;
  }
  bool snake() const { return pure() && !(prm.flags & (SB_FLAG_NO_SNAKE | SB_FLAG_DAG)) && !halo_active(); }
  bool fuse() const { return pure() && !(prm.flags & (SB_FLAG_NO_FUSE | SB_FLAG_DAG)) && !halo_active(); }

  std::vector<Launch> program() const {
    std::vector<Launch> L;
    const int S = prm.substeps, I = prm.iterations;
    auto simple = [&](Launch::Kind k, int arg = 0) { Launch l; l.kind = k; l.arg = arg; L.push_back(l); };
    // (a distributed handle always ends its frame with the normals launch: without normals it is the closing handshake)
    const bool want_normals = (!plan.surf_ids.empty() && !(prm.flags & SB_FLAG_NO_NORMALS)) || dist.ctl != nullptr;
    if (use_dag() || halo_active() || !pure()) {
      for (int ss = 0; ss < S; ss++) {
        simple(Launch::PREDICT);
        if (use_dag()) simple(Launch::DAG);
        else
          for (int it = 0; it < I; it++) {
            simple(Launch::GROUP, 0);
            if (halo_active()) simple(Launch::EXCHANGE, 0);
            simple(Launch::GROUP, 1);
            if (halo_active()) simple(Launch::EXCHANGE, 1);
          }
        simple(Launch::FINISH);
      }
      if (want_normals) simple(Launch::NORMALS);
      return L;
    }
    std::vector<int> live; // passes that hold constraints
    for (size_t k = 0; k < plan.passes.size(); k++)
      if (plan.passes[k].n_edges + plan.passes[k].n_tets) live.push_back((int)k);
    const bool sn = snake(), fu = fuse();
    auto contiguous = [&](int k) { return plan.passes[k].contiguous; };
    // a pass launch under construction: reps of every segment (substep) it spans
    struct Open { int pass = -1; std::vector<uint32_t> seg; bool pre = false; } open;
    auto flush = [&](bool post) {
      if (open.pass < 0) return;
      // the kernel runs segments of equal length: split where the repetition count changes; a split falls on a
      // substep boundary of a contiguous pass, so the halves carry it as finish (post) and predict (pre)
      size_t lo = 0;
      while (lo < open.seg.size()) {
        size_t hi = lo + 1;
        while (hi < open.seg.size() && open.seg[hi] == open.seg[lo]) hi++;
        Launch l;
        l.kind = Launch::PASS; l.arg = open.pass;
        l.n_seg = (uint32_t)(hi - lo); l.reps = open.seg[lo];
        l.pre = lo == 0 ? open.pre : true;
        l.post = hi == open.seg.size() ? post : true;
        L.push_back(l);
        lo = hi;
      }
      open = Open();
    };
    for (int ss = 0; ss < S; ss++) {
      bool first_of_substep = true;
      auto occurrence = [&](int k) {
        if (first_of_substep) {
          first_of_substep = false;
          // substep boundary: finish of the previous substep (none before the first), predict of this one
          if (fu && open.pass == k && contiguous(k)) { open.seg.push_back(1); return; } // carried inside the launch
          const bool had = open.pass >= 0 || ss > 0;
          if (fu && open.pass >= 0 && contiguous(open.pass)) flush(true);
          else { flush(false); if (had) simple(Launch::FINISH); }
          if (fu && contiguous(k)) { open.pass = k; open.seg = {1}; open.pre = true; return; }
          simple(Launch::PREDICT);
          open.pass = k; open.seg = {1}; open.pre = false;
          return;
        }
        if (fu && open.pass == k) { open.seg.back()++; return; }
        flush(false);
        open.pass = k; open.seg = {1}; open.pre = false;
      };
      if (live.empty() || I == 0) { // nothing to project: the per-vertex kernels alone
        flush(false);
        simple(Launch::PREDICT);
        simple(Launch::FINISH);
        continue;
      }
      for (int it = 0; it < I; it++) {
        if (sn && (it & 1)) for (size_t j = live.size(); j-- > 0;) occurrence(live[j]);
        else for (int k : live) occurrence(k);
      }
    }
    if (open.pass >= 0) {
      if (fu && contiguous(open.pass)) flush(true);
      else { flush(false); simple(Launch::FINISH); }
    }
    if (want_normals) simple(Launch::NORMALS);
    // distributed meshes: a tile launch hands its vertices to the ranks that run the next launch's tiles; a per-vertex
    // kernel, the normals and the next frame (which starts with pass 0 or a per-vertex kernel) find them at home
    const int partial = plan.passes.size() > plan.n_tilings ? (int)plan.passes.size() - 1 : -1; // a leftover pass covers only some vertices
    for (size_t i = 0; i < L.size(); i++) {
      if (L[i].kind != Launch::PASS) continue;
      L[i].next_pass = (i + 1 < L.size() && L[i + 1].kind == Launch::PASS) ? (uint32_t)L[i + 1].arg : 0u;
      L[i].next_full_pass = L[i].next_pass;
      if ((int)L[i].next_pass == partial) { // the vertices it leaves out go on to the first launch after it that is not that pass
        size_t j = i + 1;
        while (j < L.size() && L[j].kind == Launch::PASS && L[j].arg == partial) j++;
        L[i].next_full_pass = (j < L.size() && L[j].kind == Launch::PASS) ? (uint32_t)L[j].arg : 0u;
      }
    }
    return L;
  }

  void run_launch(const Launch &l, cudaStream_t s) {
    switch (l.kind) {
      case Launch::PREDICT: launch_predict(s); break;
      case Launch::FINISH: launch_finish(s); break;
      case Launch::PASS: launch_pass((size_t)l.arg, s, l.n_seg, l.reps, l.pre, l.post, l.next_pass, l.next_full_pass); break;
      case Launch::GLOBAL: launch_global(s, l.arg); break;
      case Launch::GROUP: launch_group(l.arg, s); break;
      case Launch::EXCHANGE: exchange(l.arg, s); break;
      case Launch::NORMALS: launch_normals(s); break;
      case Launch::DAG: launch_dag(s); break;
    }
  }

  uint32_t launches_per_frame() const {
    uint32_t n = 0;
    for (const Launch &l : program()) {
      if (l.kind == Launch::GROUP) {
        for (size_t k = 0; k < passes.size(); k++) n += plan.passes[k].group == l.arg && passes[k].grid && !passes[k].empty;
        for (auto &b : plan.gbatches) n += b.cnt && b.group == l.arg;
      } else if (l.kind == Launch::EXCHANGE) {
        for (auto &kv : links) {
          auto hl = halo.find(kv.first);
          if (hl == halo.end() || !hl->second.n) continue;
          // phase 0: list 1 sends, list 0 receives; phase 1 the other way round
          const bool sends = kv.first == (l.arg == 0 ? 1 : 0);
          n += sends ? (kv.second.peer_buf ? 1 : 0) : (kv.second.recv.p ? 1 : 0);
        }
      } else {
        n++;
      }
    }
    return n;
  }

  void enqueue_frame(cudaStream_t s) {
    for (const Launch &l : program()) run_launch(l, s);
    CK(cudaGetLastError());
  }

  void step(float dt) {
    CK(cudaSetDevice(device));
    if (!(dt > 0)) dt = prm.dt;
    if (!(dt > 0) || !std::isfinite(dt)) throw std::string("dt must be finite and > 0");
    if (dist.ctl && !dist_on) throw std::string("distributed handle: not every peer is connected yet");
    refresh_params(dt);
    if (prm.flags & SB_FLAG_NO_GRAPH) {
      enqueue_frame(stream);
      return;
    }
    const uint64_t key = ((uint64_t)(uint32_t)prm.substeps << 40) | ((uint64_t)(uint32_t)prm.iterations << 16) |
                         (uint64_t)(prm.flags & 0xffff);
    auto it = graphs.find(key);
    if (it == graphs.end()) {
      cudaGraph_t g = nullptr;
      CK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
      try {
        enqueue_frame(cap_stream);
      } catch (...) {
        cudaStreamEndCapture(cap_stream, &g);
        if (g) cudaGraphDestroy(g);
        throw;
      }
      CK(cudaStreamEndCapture(cap_stream, &g));
      cudaGraphExec_t ge = nullptr;
      cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
      cudaGraphDestroy(g);
      CK(e);
      it = graphs.emplace(key, ge).first;
    }
    CK(cudaGraphLaunch(it->second, stream));
  }

  // ---- I/O ------------------------------------------------------------------------
  float *pinned(size_t bytes) {
    if (bytes > pin_bytes) {
      if (pin) cudaFreeHost(pin);
      pin = nullptr;
      pin_bytes = 0;
      CK(cudaMallocHost(&pin, bytes));
      pin_bytes = bytes;
    }
    return pin;
  }
  void d2h(void *dst, const void *src, size_t bytes) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  void read_positions(float *dst) {
    CK(cudaSetDevice(device));
    k_gather_xyz<<<grid_for(plan.V, 256), 256, 0, stream>>>(plan.V, inv.p, x.p, stage_f.p);
    CK(cudaGetLastError());
    d2h(dst, stage_f.p, 3 * (size_t)plan.V * sizeof(float));
  }
  void read_surface(float *dpos, float *dnrm) {
    CK(cudaSetDevice(device));
    const uint32_t ns = (uint32_t)plan.surf_ids.size();
    if (!ns) return;
    if (dpos) {
      k_gather_xyz<<<grid_for(ns, 256), 256, 0, stream>>>(ns, surf_slot.p, x.p, stage_f.p);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(dpos, stage_f.p, 3 * (size_t)ns * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    if (dnrm) {
      float *st = stage_f.p + 3 * (size_t)ns; // ns <= V/2 is not guaranteed: use stage_a when tight
      if (6 * (size_t)ns > 3 * (size_t)plan.V) st = (float *)stage_a.p;
      k_gather_xyz<<<grid_for(ns, 256), 256, 0, stream>>>(ns, nullptr, nrm.p, st);
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(dnrm, st, 3 * (size_t)ns * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaStreamSynchronize(stream));
  }
  // Render mesh: upload the binding (tet per render vertex + weights) and the CSR of its triangles for the normals.
  void skin_upload(const int32_t *tris, uint32_t n, uint32_t nf) {
    CK(cudaSetDevice(device));
    std::vector<uint4> slots(n);
    for (uint32_t i = 0; i < n; i++) {
      const int32_t *q = &plan.tets[4 * (size_t)skin_tet[i]];
      slots[i] = make_uint4(plan.inv[q[0]], plan.inv[q[1]], plan.inv[q[2]], plan.inv[q[3]]);
    }
    skin_slots.upload(slots, &dev_bytes);
    std::vector<float4> w(n);
    for (uint32_t i = 0; i < n; i++)
      w[i] = make_float4(skin_bary[4 * (size_t)i], skin_bary[4 * (size_t)i + 1], skin_bary[4 * (size_t)i + 2], skin_bary[4 * (size_t)i + 3]);
    skin_w.upload(w, &dev_bytes);
    skin_x.alloc(n, &dev_bytes);
    skin_nrm.alloc(n, &dev_bytes);
    skin_stage.alloc(6 * (size_t)n, &dev_bytes);
    std::vector<uint32_t> off((size_t)n + 1, 0), ids(3 * (size_t)nf);
    for (size_t k = 0; k < 3 * (size_t)nf; k++) off[(size_t)tris[k] + 1]++;
    for (uint32_t i = 0; i < n; i++) off[i + 1] += off[i];
    {
      std::vector<uint32_t> cur(off.begin(), off.end() - 1);
      for (uint32_t f = 0; f < nf; f++) // ascending triangle id per vertex, as the oracle sums them
        for (int j = 0; j < 3; j++) ids[cur[tris[3 * (size_t)f + j]]++] = f;
    }
    skin_tri_off.upload(off, &dev_bytes);
    skin_tri_ids.upload(ids, &dev_bytes);
    skin_tris.upload(std::vector<int32_t>(tris, tris + 3 * (size_t)nf), &dev_bytes);
    skin_n = n;
    skin_f = nf;
  }
  void read_skinned(float *dpos, float *dnrm) {
    CK(cudaSetDevice(device));
    const uint32_t n = skin_n;
    k_skin<<<grid_for(n, 256), 256, 0, stream>>>(n, skin_slots.p, skin_w.p, x.p, skin_x.p);
    if (dpos) {
      k_gather_xyz<<<grid_for(n, 256), 256, 0, stream>>>(n, nullptr, skin_x.p, skin_stage.p);
      CK(cudaMemcpyAsync(dpos, skin_stage.p, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    if (dnrm) {
      k_normals<<<grid_for(n, 256), 256, 0, stream>>>(n, skin_tri_off.p, skin_tri_ids.p, skin_tris.p, skin_x.p, skin_nrm.p);
      k_gather_xyz<<<grid_for(n, 256), 256, 0, stream>>>(n, nullptr, skin_nrm.p, skin_stage.p + 3 * (size_t)n);
      CK(cudaMemcpyAsync(dnrm, skin_stage.p + 3 * (size_t)n, 3 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, stream));
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream));
  }
  void get_state(float *x4, float *v4) {
    CK(cudaSetDevice(device));
    const uint32_t V = plan.V;
    if (x4) k_gather4<<<grid_for(V, 256), 256, 0, stream>>>(V, inv.p, x.p, stage_a.p);
    if (v4) k_gather4<<<grid_for(V, 256), 256, 0, stream>>>(V, inv.p, v.p, stage_b.p);
    CK(cudaGetLastError());
    if (x4) CK(cudaMemcpyAsync(x4, stage_a.p, V * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    if (v4) CK(cudaMemcpyAsync(v4, stage_b.p, V * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
  }
  void set_state(const float *x4, const float *v4) {
    CK(cudaSetDevice(device));
    const uint32_t V = plan.V;
    if (x4) {
      CK(cudaMemcpyAsync(stage_a.p, x4, V * sizeof(float4), cudaMemcpyHostToDevice, stream));
      k_scatter4<<<grid_for(V, 256), 256, 0, stream>>>(V, inv.p, stage_a.p, x.p);
    }
    if (v4) {
      CK(cudaMemcpyAsync(stage_b.p, v4, V * sizeof(float4), cudaMemcpyHostToDevice, stream));
      k_scatter4<<<grid_for(V, 256), 256, 0, stream>>>(V, inv.p, stage_b.p, v.p);
      if (plan.n_ghost) k_mark_ghosts<<<grid_for(plan.n_ghost, 256), 256, 0, stream>>>(plan.n_ghost, ghost_slot.p, v.p);
    }
    CK(cudaGetLastError());
    // the host buffers belong to the caller again when this returns
    CK(cudaStreamSynchronize(stream));
  }

  // ---- one frame in one buffer (sb_read_packed / sb_write_packed) --------------------------------------
  uint32_t packed_verts() const { return dist.ctl ? n_own : plan.V; }
  uint32_t packed_surf() const { return dist.ctl ? n_own_surf : (uint32_t)plan.surf_ids.size(); }
  size_t packed_bytes(bool with_surface) const { return 32 * (size_t)packed_verts() + (with_surface ? 24 * (size_t)packed_surf() : 0); }
  float4 *pack_staging() {
    const size_t need = (packed_bytes(true) + 15) / 16;
    if (pack_buf.n < need) pack_buf.alloc(need, &dev_bytes);
    return pack_buf.p;
  }
  void read_packed(void *dst) {
    CK(cudaSetDevice(device));
    const uint32_t n = packed_verts(), ns = packed_surf();
    float4 *st = pack_staging();
    k_pack_frame<<<grid_for((size_t)n + ns, 256), 256, 0, stream>>>(n, dist.ctl ? own_slot.p : inv.p, ns, dist.ctl ? own_surf.p : nullptr,
                                                                 surf_slot.p, x.p, v.p, nrm.p, st);
    CK(cudaGetLastError());
    d2h(dst, st, packed_bytes(true));
  }
  void write_packed(const void *src) {
    CK(cudaSetDevice(device));
    const uint32_t n = packed_verts();
    float4 *st = pack_staging();
    CK(cudaMemcpyAsync(st, src, packed_bytes(false), cudaMemcpyHostToDevice, stream));
    k_unpack_state<<<grid_for(n, 256), 256, 0, stream>>>(n, dist.ctl ? own_slot.p : inv.p, st, x.p, v.p);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(stream)); // the host buffer belongs to the caller again when this returns
  }

  void diagnostics(double *out) {
    CK(cudaSetDevice(device));
    if (!d_edges.p && plan.E) {
      std::vector<int2> e(plan.E);
      for (uint32_t k = 0; k < plan.E; k++)
        e[k] = make_int2((int)plan.inv[plan.edges[2 * (size_t)k]], (int)plan.inv[plan.edges[2 * (size_t)k + 1]]);
      d_edges.upload(e, &dev_bytes);
      d_elen.upload(plan.rest_len, &dev_bytes);
    }
    if (!d_tets.p && plan.T) {
      std::vector<int4> q(plan.T);
      for (uint32_t t = 0; t < plan.T; t++) {
        const int32_t *s = &plan.tets[4 * (size_t)t];
        q[t] = make_int4((int)plan.inv[s[0]], (int)plan.inv[s[1]], (int)plan.inv[s[2]], (int)plan.inv[s[3]]);
      }
      d_tets.upload(q, &dev_bytes);
    }
    if (cur_dt < 0) refresh_params(prm.dt);
    CK(cudaMemsetAsync(d_part.p, 0, (size_t)kDiagBlocks * 16 * sizeof(double), stream));
    k_diag_verts<<<kDiagBlocks, 256, 0, stream>>>(plan.V, x.p, v.p, dprm.p, d_part.p);
    k_diag_tets<<<kDiagBlocks, 256, 0, stream>>>(plan.T, d_tets.p, x.p, d_part.p);
    k_diag_edges<<<kDiagBlocks, 256, 0, stream>>>(plan.E, d_edges.p, d_elen.p, x.p, d_part.p);
    CK(cudaGetLastError());
    std::vector<double> part((size_t)kDiagBlocks * 16);
    d2h(part.data(), d_part.p, part.size() * sizeof(double));
    double acc[16] = {0};
    double maxneg = -INFINITY, smax = 0;
    for (int b = 0; b < kDiagBlocks; b++) {
      const double *p = &part[(size_t)b * 16];
      for (int k = 0; k < 12; k++) acc[k] += p[k];
      maxneg = std::max(maxneg, p[12]);
      acc[13] += p[13];
      smax = std::max(smax, p[14]);
      acc[15] += p[15];
    }
    const double V = plan.V;
    out[0] = acc[0]; out[1] = acc[1]; out[2] = acc[13];
    out[3] = acc[2] / V; out[4] = acc[3] / V; out[5] = acc[4] / V;
    for (int k = 0; k < 6; k++) out[6 + k] = acc[5 + k];
    out[12] = smax;
    out[13] = plan.E ? std::sqrt(acc[15] / plan.E) : 0.0;
    out[14] = acc[11];
    out[15] = -maxneg;
  }

  float time_kernel(int which, int reps) {
    CK(cudaSetDevice(device));
    if (dist.ctl) throw std::string("sb_time_kernel on a distributed handle would run launches the peers do not run (the epochs would part)");
    if (cur_dt < 0) refresh_params(prm.dt);
    const size_t nb = plan.V * sizeof(float4);
    // save state in the staging buffers plus one temporary
    DevBuf<float4> keep;
    keep.alloc(plan.V, nullptr);
    CK(cudaMemcpyAsync(stage_a.p, x.p, nb, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(stage_b.p, v.p, nb, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(keep.p, xp.p, nb, cudaMemcpyDeviceToDevice, stream));
    auto run = [&]() {
      if (which == 0) launch_predict(stream);
      else if (which == 1) launch_finish(stream);
      else if (which == 2) launch_normals(stream);
      else if (which >= 16 && which < 16 + (int)passes.size()) launch_pass((size_t)(which - 16), stream);
      else if (which == 32) launch_global(stream);
      else if (which == 48 && dag_ready) launch_dag(stream);
      else throw std::string("unknown kernel selector");
    };
    run(); // warm-up
    CK(cudaEventRecord(ev0, stream));
    for (int r = 0; r < reps; r++) run();
    CK(cudaEventRecord(ev1, stream));
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(x.p, stage_a.p, nb, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(v.p, stage_b.p, nb, cudaMemcpyDeviceToDevice, stream));
    CK(cudaMemcpyAsync(xp.p, keep.p, nb, cudaMemcpyDeviceToDevice, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));
    return ms / (float)reps;
  }
};

// ---- C ABI -----------------------------------------------------------------------

static thread_local std::string g_create_error;

static std::string validate_params(const sb_params *p) {
  if (!p) return "params is NULL";
  if (p->substeps < 1 || p->substeps > 1024) return "substeps must be in [1, 1024]";
  if (p->iterations < 0 || p->iterations > 4096) return "iterations must be in [0, 4096]";
  if (!(p->dt > 0) || !std::isfinite(p->dt)) return "dt must be finite and > 0";
  if (std::isnan(p->stiffness_distance) || std::isnan(p->stiffness_volume)) return "stiffness is NaN";
  if (!std::isfinite(p->damping) || p->damping < 0) return "damping must be finite and >= 0";
  if (!(p->friction >= 0 && p->friction <= 1)) return "friction must be in [0, 1]";
  for (int k = 0; k < 3; k++)
    if (!std::isfinite(p->gravity[k])) return "gravity must be finite";
  if (!std::isfinite(p->ground_y)) return "ground_y must be finite";
  return "";
}

template <class F>
static int guarded(sb_handle h, F fn) {
  std::string &err = h ? h->err : g_create_error;
  try {
    return fn();
  } catch (const CudaError &e) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at solver.cu:%d: %s", (int)e.code, cudaGetErrorString(e.code), e.line, e.what);
    err = buf;
    return SB_E_CUDA;
  } catch (const std::string &s) {
    err = s;
    return SB_E_ARG;
  } catch (const std::bad_alloc &) {
    err = "out of host memory";
    return SB_E_NOMEM;
  } catch (...) {
    err = "unexpected exception";
    return SB_E_STATE;
  }
}

#define NEED_HANDLE(h) \
  if (!(h)) return SB_E_ARG
#define NEED_DEVICE(h)                                   \
  NEED_HANDLE(h);                                        \
  if (!(h)->on_device) {                                 \
    (h)->err = "handle was created by sb_plan (host only)"; \
    return SB_E_STATE;                                   \
  }
// whole-mesh reads / writes and stray launches make no sense on ONE rank of a distributed mesh (a rank holds current
// values for the vertices it owns only, and a launch its peers do not run parts the epochs)
#define NOT_DISTRIBUTED(h, what)                                                                         \
  do {                                                                                                   \
    if ((h)->dist.ctl) {                                                                                 \
      (h)->err = what " is not available on one rank of a distributed mesh (use sb_read_packed / sb_write_packed per rank)"; \
      return SB_E_STATE;                                                                                 \
    }                                                                                                    \
  } while (0)

// copy of a vector into a caller's buffer; an empty vector's data() may be null, which memcpy must not be given
template <class T>
static inline void copy_out(T *dst, const std::vector<T> &src) {
  if (dst && !src.empty()) std::memcpy(dst, src.data(), src.size() * sizeof(T));
}

extern "C" {

int sb_abi_check(uint32_t *version, uint32_t *sizeof_params, uint32_t *sizeof_desc, uint32_t *sizeof_info) {
  if (version) *version = SB_ABI_VERSION;
  if (sizeof_params) *sizeof_params = (uint32_t)sizeof(sb_params);
  if (sizeof_desc) *sizeof_desc = (uint32_t)sizeof(sb_mesh_desc);
  if (sizeof_info) *sizeof_info = (uint32_t)sizeof(sb_info);
  return SB_OK;
}

void sb_default_params(sb_params *p) {
  if (!p) return;
  std::memset(p, 0, sizeof *p);
  p->dt = 1.0f / 60.0f;
  p->substeps = 10;
  p->iterations = 10;
  p->stiffness_distance = INFINITY;
  p->stiffness_volume = INFINITY;
  p->damping = 0.f;
  p->friction = 0.f;
  p->gravity[1] = -9.81f;
  p->ground_y = 0.f;
  p->flags = 0;
}

static int create_impl(const sb_mesh_desc *mesh, const sb_params *params, sb_handle *out, bool device) {
  if (!out) return SB_E_ARG;
  *out = nullptr;
  g_create_error.clear();
  if (!mesh) { g_create_error = "mesh is NULL"; return SB_E_ARG; }
  if (mesh->dist_ranks != 0 && (mesh->dist_ranks < 2 || mesh->dist_ranks > SB_MAX_RANKS)) { g_create_error = "dist_ranks must be 0 or 2..8"; return SB_E_ARG; }
  if (mesh->attach_edges < 0 || mesh->attach_edges > 2) { g_create_error = "attach_edges must be 0, 1 or 2"; return SB_E_ARG; }
  sb_params dp;
  sb_default_params(&dp);
  if (params) dp = *params;
  std::string perr = validate_params(&dp);
  if (!perr.empty()) { g_create_error = perr; return SB_E_ARG; }
  sb_solver *s = nullptr;
  int rc = guarded(nullptr, [&]() -> int {
    s = new sb_solver();
    s->prm = dp;
    MeshInput in{mesh->pos_xyz, mesh->tets, mesh->surf_tris, mesh->inv_mass, mesh->n_verts, mesh->n_tets, mesh->n_tris, mesh->density,
                 (uint32_t)std::max(0, mesh->n_ghost_verts), mesh->edges, mesh->edges ? mesh->n_edges : 0u};
    PlanOptions opt;
    opt.tile_cap = mesh->tile_cap;
    opt.max_tile_passes = mesh->max_tile_passes;
    opt.later_cap = mesh->later_tile_cap;
    opt.threads = mesh->host_threads;
    opt.block_threads = mesh->block_threads;
    opt.round_width = mesh->round_width;
    opt.compounds = mesh->attach_edges == 0 ? -1 : mesh->attach_edges == 1 ? 1 : 0;
    opt.tilings = mesh->tilings;
    opt.dist_ranks = mesh->dist_ranks;
    if (device) {
      int ndev = 0;
      CK(cudaGetDeviceCount(&ndev));
      if (mesh->device < 0 || mesh->device >= ndev) throw std::string("no such CUDA device");
      cudaDeviceProp prop;
      CK(cudaGetDeviceProperties(&prop, mesh->device));
      opt.n_sm = prop.multiProcessorCount;
    }
    std::string e = build_plan(in, opt, s->plan);
    if (!e.empty()) throw e;
    if (device) {
      auto t0 = std::chrono::steady_clock::now();
      s->upload(*mesh);
      s->plan.build_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    return SB_OK;
  });
  if (rc != SB_OK) {
    delete s;
    return rc;
  }
  *out = s;
  return SB_OK;
}

int sb_create(const sb_mesh_desc *mesh, const sb_params *params, sb_handle *out) { return create_impl(mesh, params, out, true); }

/* Host-only handle: topology, tiling, colouring and schedule without touching CUDA.
   sb_get_info / sb_get_topology / sb_get_schedule / sb_surface_vertices work on it. */
int sb_plan(const sb_mesh_desc *mesh, const sb_params *params, sb_handle *out) { return create_impl(mesh, params, out, false); }

int sb_destroy(sb_handle h) {
  if (!h) return SB_E_ARG;
  delete h;
  return SB_OK;
}

int sb_set_params(sb_handle h, const sb_params *p) {
  NEED_HANDLE(h);
  std::string e = validate_params(p);
  if (!e.empty()) { h->err = e; return SB_E_ARG; }
  h->prm = *p;
  h->prm_dirty = true;
  return SB_OK;
}

int sb_get_params(sb_handle h, sb_params *out) {
  NEED_HANDLE(h);
  if (!out) return SB_E_ARG;
  *out = h->prm;
  return SB_OK;
}

int sb_set_colliders_ex(sb_handle h, const sb_collider *c, uint32_t n) {
  NEED_HANDLE(h);
  if (n > SB_MAX_COLLIDERS || (n && !c)) { h->err = "at most 16 colliders"; return SB_E_ARG; }
  for (uint32_t k = 0; k < n; k++) {
    const int np = c[k].kind == SB_COLLIDER_SPHERE ? 4 : c[k].kind == SB_COLLIDER_CAPSULE ? 7 : 10;
    if (c[k].kind < SB_COLLIDER_SPHERE || c[k].kind > SB_COLLIDER_BOX) { h->err = "unknown collider kind"; return SB_E_ARG; }
    if (!(c[k].friction >= 0.f && c[k].friction <= 1.f)) { h->err = "collider friction outside [0, 1]"; return SB_E_ARG; }
    for (int j = 0; j < np; j++)
      if (!std::isfinite(c[k].p[j])) { h->err = "collider is not finite"; return SB_E_ARG; }
    const bool neg = c[k].kind == SB_COLLIDER_BOX ? (c[k].p[3] < 0.f || c[k].p[4] < 0.f || c[k].p[5] < 0.f) : c[k].p[3] < 0.f;
    if (neg) { h->err = "collider radius / half extent is negative"; return SB_E_ARG; }
  }
  for (uint32_t k = 0; k < n; k++) h->colliders[k] = c[k];
  h->n_col = (int)n;
  h->prm_dirty = true;
  return SB_OK;
}

int sb_set_colliders(sb_handle h, const float *s, uint32_t n) {
  NEED_HANDLE(h);
  if (n > SB_MAX_COLLIDERS || (n && !s)) { h->err = "at most 16 sphere colliders"; return SB_E_ARG; }
  for (uint32_t k = 0; k < 4 * n; k++)
    if (!std::isfinite(s[k])) { h->err = "collider is not finite"; return SB_E_ARG; }
  for (uint32_t k = 0; k < n; k++) {
    sb_collider c{};
    c.kind = SB_COLLIDER_SPHERE;
    for (int j = 0; j < 4; j++) c.p[j] = s[4 * k + j];
    h->colliders[k] = c;
  }
  h->n_col = (int)n;
  h->prm_dirty = true;
  return SB_OK;
}

int sb_step(sb_handle h, float dt) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int { h->step(dt); h->frames_done++; return SB_OK; });
}

int sb_synchronize(sb_handle h) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    if (h->dag_error()) { h->err = "tile DAG: a dependency wait timed out"; return SB_E_STATE; }
    if (h->dist_error_word()) { h->err = "a wait for a peer GPU timed out: the state is invalid"; return SB_E_STATE; }
    return SB_OK;
  });
}

int sb_read_positions(sb_handle h, float *dst, uint32_t n) {
  NEED_DEVICE(h);
  if (!dst || n != h->plan.V) { h->err = "dst is NULL or n_verts mismatch"; return SB_E_ARG; }
  return guarded(h, [&]() -> int {
    h->read_positions(dst);
    if (h->dist_error_word()) { h->err = "a wait for a peer GPU timed out: the state is invalid"; return SB_E_STATE; }
    return SB_OK;
  });
}

int sb_surface_vertices(sb_handle h, int32_t *ids, uint32_t capacity, uint32_t *n_surface) {
  NEED_HANDLE(h);
  const uint32_t ns = (uint32_t)h->plan.surf_ids.size();
  if (n_surface) *n_surface = ns;
  if (ids) {
    if (capacity < ns) { h->err = "capacity too small"; return SB_E_ARG; }
    copy_out(ids, h->plan.surf_ids);
  }
  return SB_OK;
}

int sb_read_surface(sb_handle h, float *dpos, float *dnrm, uint32_t n_surface) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_read_surface");
  if (n_surface != h->plan.surf_ids.size()) { h->err = "n_surface mismatch"; return SB_E_ARG; }
  return guarded(h, [&]() -> int { h->read_surface(dpos, dnrm); return SB_OK; });
}

int sb_read_normals(sb_handle h, float *dst, uint32_t n) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_read_normals");
  if (!dst || n != h->plan.V) { h->err = "dst is NULL or n_verts mismatch"; return SB_E_ARG; }
  return guarded(h, [&]() -> int {
    const size_t ns = h->plan.surf_ids.size();
    std::memset(dst, 0, 3 * (size_t)n * sizeof(float));
    if (!ns) return SB_OK;
    float *tmp = h->pinned(3 * ns * sizeof(float));
    h->read_surface(nullptr, tmp);
    for (size_t s = 0; s < ns; s++) std::memcpy(dst + 3 * (size_t)h->plan.surf_ids[s], tmp + 3 * s, 3 * sizeof(float));
    return SB_OK;
  });
}

// ---- render mesh bound to the tets ---------------------------------------------------------------------
int sb_skin_bind(sb_handle h, const float *pos, uint32_t n, const int32_t *tris, uint32_t nf) {
  NEED_HANDLE(h);
  if (!n || !pos || (nf && !tris)) { h->err = "null or empty render mesh"; return SB_E_ARG; }
  for (size_t k = 0; k < 3 * (size_t)nf; k++)
    if (tris[k] < 0 || (uint32_t)tris[k] >= n) { h->err = "render triangle index out of range"; return SB_E_ARG; }
  if (h->dist.ctl) { h->err = "a render mesh cannot be bound to one rank of a distributed mesh"; return SB_E_STATE; }
  return guarded(h, [&]() -> int {
    h->skin_tet.assign(n, 0);
    h->skin_bary.assign(4 * (size_t)n, 0.f);
    const int rc = sb_skin_compute(h->plan.pos.data(), h->plan.V, h->plan.tets.data(), h->plan.T, pos, n, h->skin_tet.data(),
                                   h->skin_bary.data());
    if (rc != SB_OK) {
      h->skin_tet.clear();
      h->skin_bary.clear();
      h->skin_n = 0;
      h->err = sb_ingest_last_error();
      return rc;
    }
    if (h->on_device) h->skin_upload(tris, n, nf);
    else { h->skin_n = n; h->skin_f = nf; }
    return SB_OK;
  });
}

int sb_skin_get_binding(sb_handle h, int32_t *tet_of, float *bary4, uint32_t n) {
  NEED_HANDLE(h);
  if (!h->skin_n) { h->err = "no render mesh bound (sb_skin_bind)"; return SB_E_STATE; }
  if (n != h->skin_n) { h->err = "render vertex count mismatch"; return SB_E_ARG; }
  if (tet_of) std::memcpy(tet_of, h->skin_tet.data(), (size_t)n * sizeof(int32_t));
  if (bary4) std::memcpy(bary4, h->skin_bary.data(), 4 * (size_t)n * sizeof(float));
  return SB_OK;
}

int sb_read_skinned(sb_handle h, float *dst_pos, float *dst_nrm, uint32_t n) {
  NEED_DEVICE(h);
  if (!h->skin_n) { h->err = "no render mesh bound (sb_skin_bind)"; return SB_E_STATE; }
  if (n != h->skin_n) { h->err = "render vertex count mismatch"; return SB_E_ARG; }
  if (!dst_pos && !dst_nrm) return SB_E_ARG;
  return guarded(h, [&]() -> int { h->read_skinned(dst_pos, dst_nrm); return SB_OK; });
}

// ---- snapshots (file layout: ingest.cpp) ------------------------------------------------------------------
int sb_save_state(sb_handle h, const char *path) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_save_state");
  if (!path) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    const uint32_t V = h->plan.V;
    std::vector<float> x4(4 * (size_t)V), v4(4 * (size_t)V);
    h->get_state(x4.data(), v4.data());
    const int rc = sb_state_write(path, x4.data(), v4.data(), V, &h->prm, h->frames_done,
                                  sb_topology_hash(V, h->plan.tets.data(), h->plan.T));
    if (rc != SB_OK) h->err = sb_ingest_last_error();
    return rc;
  });
}

int sb_load_state(sb_handle h, const char *path, int32_t apply_params) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_load_state");
  if (!path) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    const uint32_t V = h->plan.V;
    uint32_t n = 0;
    uint64_t frame = 0, topo = 0;
    sb_params p;
    int rc = sb_state_read(path, nullptr, nullptr, 0, &n, &p, &frame, &topo);
    if (rc == SB_OK && (n != V || topo != sb_topology_hash(V, h->plan.tets.data(), h->plan.T))) {
      h->err = "snapshot belongs to a different mesh";
      return SB_E_ARG;
    }
    std::vector<float> x4(4 * (size_t)V), v4(4 * (size_t)V);
    if (rc == SB_OK) rc = sb_state_read(path, x4.data(), v4.data(), V, nullptr, nullptr, nullptr, nullptr);
    if (rc != SB_OK) { h->err = sb_ingest_last_error(); return rc; }
    if (apply_params) {
      rc = sb_set_params(h, &p);
      if (rc != SB_OK) return rc;
    }
    h->set_state(x4.data(), v4.data());
    h->frames_done = frame;
    return SB_OK;
  });
}

int sb_frames_done(sb_handle h, uint64_t *out) {
  NEED_HANDLE(h);
  if (!out) return SB_E_ARG;
  *out = h->frames_done;
  return SB_OK;
}

int sb_get_state(sb_handle h, float *x4, float *v4, uint32_t n) {
  NEED_DEVICE(h);
  if (n != h->plan.V) { h->err = "n_verts mismatch"; return SB_E_ARG; }
  return guarded(h, [&]() -> int {
    h->get_state(x4, v4);
    if (h->dist_error_word()) { h->err = "a wait for a peer GPU timed out: the state is invalid"; return SB_E_STATE; }
    return SB_OK;
  });
}

int sb_set_state(sb_handle h, const float *x4, const float *v4, uint32_t n) {
  NEED_DEVICE(h);
  if (n != h->plan.V) { h->err = "n_verts mismatch"; return SB_E_ARG; }
  return guarded(h, [&]() -> int { h->set_state(x4, v4); return SB_OK; });
}

int sb_packed_sizes(sb_handle h, uint32_t *n_verts, uint32_t *n_surface, uint64_t *bytes_in, uint64_t *bytes_out) {
  NEED_DEVICE(h);
  if (n_verts) *n_verts = h->packed_verts();
  if (n_surface) *n_surface = h->packed_surf();
  if (bytes_in) *bytes_in = h->packed_bytes(false);
  if (bytes_out) *bytes_out = h->packed_bytes(true);
  return SB_OK;
}

int sb_read_packed(sb_handle h, void *dst, uint64_t bytes) {
  NEED_DEVICE(h);
  if (!dst || bytes != h->packed_bytes(true)) { h->err = "dst is NULL or the size is not sb_packed_sizes' bytes_out"; return SB_E_ARG; }
  return guarded(h, [&]() -> int {
    h->read_packed(dst);
    if (h->dist_error_word()) { h->err = "a wait for a peer GPU timed out: the state is invalid"; return SB_E_STATE; }
    return SB_OK;
  });
}

int sb_write_packed(sb_handle h, const void *src, uint64_t bytes) {
  NEED_DEVICE(h);
  if (!src || bytes != h->packed_bytes(false)) { h->err = "src is NULL or the size is not sb_packed_sizes' bytes_in"; return SB_E_ARG; }
  return guarded(h, [&]() -> int { h->write_packed(src); return SB_OK; });
}

int sb_diagnostics(sb_handle h, double *out16) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_diagnostics");
  if (!out16) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    h->diagnostics(out16);
    if (out16[14] > 0) { h->err = "non-finite positions"; return SB_E_NAN; }
    return SB_OK;
  });
}

int sb_get_info(sb_handle h, sb_info *o) {
  NEED_HANDLE(h);
  if (!o) return SB_E_ARG;
  std::memset(o, 0, sizeof *o);
  const Plan &P = h->plan;
  o->n_verts = P.V; o->n_edges = P.E; o->n_tets = P.T; o->n_tris = P.F;
  o->n_surface_verts = (uint32_t)P.surf_ids.size();
  o->n_tile_passes = (uint32_t)P.passes.size();
  o->n_ghost_verts = P.n_ghost;
  o->first_cut_pass = (uint32_t)P.passes.size();
  for (size_t k = P.passes.size(); k-- > 0;)
    if (P.passes[k].group == 1) o->first_cut_pass = (uint32_t)k;
  for (size_t k = 0; k < P.passes.size(); k++)
    if (P.passes[k].group == 1) o->constraints_cut += P.passes[k].n_edges + P.passes[k].n_tets;
  for (const GlobalBatch &b : P.gbatches)
    if (b.group == 1) o->constraints_cut += b.cnt;
  o->n_tilings = P.n_tilings;
  uint32_t nb = 0;
  for (size_t k = 0; k < P.passes.size(); k++) {
    nb += P.passes[k].max_ecol + P.passes[k].max_tcol;
    for (uint8_t f : P.passes[k].col_flags) nb += (uint32_t)__builtin_popcount(f);
    if (k < 8) {
      o->tiles_in_pass[k] = P.passes[k].n_tiles();
      o->max_colours_in_pass[k] = P.passes[k].max_ecol + P.passes[k].max_tcol;
      o->constraints_in_pass[k] = P.passes[k].n_edges + P.passes[k].n_tets;
      o->edges_in_pass[k] = P.passes[k].n_edges;
      o->runs_in_pass[k] = P.passes[k].contiguous ? 0 : P.passes[k].runs.size() - P.passes[k].n_tiles();
    }
    o->smem_bytes = std::max<uint32_t>(o->smem_bytes, h->on_device && k < h->passes.size()
                                                          ? h->passes[k].smem
                                                          : P.passes[k].max_tile_verts * 16u + 16u);
    if (k < 8) o->rounds_in_pass[k] = P.passes[k].rounds_total;
  }
  o->round_width = P.round_width;
  o->edges_attached = P.edges_attached;
  o->n_global_batches = (uint32_t)P.gbatches.size();
  o->n_batches = nb + (uint32_t)P.gbatches.size();
  o->constraints_global = P.g_edges.size() + P.g_tets.size();
  o->tile_cap = P.tile_cap;
  o->block_threads = P.passes.empty() ? 0u : P.passes[0].bt;
  o->launches_per_frame = h->on_device ? h->launches_per_frame() : 0;
  o->device_bytes = h->dev_bytes;
  o->build_seconds = P.build_seconds;
  return SB_OK;
}

int sb_get_tet_roles(sb_handle h, int32_t *tets_4T, int32_t *edge01_T, int32_t *edge23_T) {
  NEED_HANDLE(h);
  const Plan &P = h->plan;
  copy_out(tets_4T, P.tet_roles);
  copy_out(edge01_T, P.tet_e01);
  copy_out(edge23_T, P.tet_e23);
  return SB_OK;
}

int sb_get_tet_mates(sb_handle h, int32_t *mate_T, int32_t *lead_T) {
  NEED_HANDLE(h);
  const Plan &P = h->plan;
  for (uint32_t t = 0; t < P.T; t++) {
    if (mate_T) mate_T[t] = P.tet_mate[t];
    if (lead_T) lead_T[t] = P.tet_lead[t];
  }
  return SB_OK;
}

int sb_get_topology(sb_handle h, int32_t *edges, float *rest_len, float *rest_vol6, float *inv_mass) {
  NEED_HANDLE(h);
  const Plan &P = h->plan;
  copy_out(edges, P.edges);
  copy_out(rest_len, P.rest_len);
  copy_out(rest_vol6, P.rest_vol6);
  copy_out(inv_mass, P.inv_mass);
  return SB_OK;
}

static int schedule_impl(sb_handle h, bool odd, int64_t *n_order, int32_t *order, int32_t *n_batches, int64_t *batch_off) {
  NEED_HANDLE(h);
  return guarded(h, [&]() -> int {
    std::vector<int32_t> ord;
    std::vector<int64_t> off;
    h->plan.export_schedule(ord, off, odd && h->snake());
    if (n_order) *n_order = (int64_t)ord.size();
    if (n_batches) *n_batches = (int32_t)off.size() - 1;
    copy_out(order, ord);
    copy_out(batch_off, off);
    return SB_OK;
  });
}

int sb_get_schedule(sb_handle h, int64_t *n_order, int32_t *order, int32_t *n_batches, int64_t *batch_off) {
  return schedule_impl(h, false, n_order, order, n_batches, batch_off);
}
int sb_get_schedule_odd(sb_handle h, int64_t *n_order, int32_t *order, int32_t *n_batches, int64_t *batch_off) {
  return schedule_impl(h, true, n_order, order, n_batches, batch_off);
}

int sb_frame_program(sb_handle h, int32_t *n_launches, int32_t *ops6, uint32_t capacity) {
  NEED_HANDLE(h);
  if (!n_launches) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    const std::vector<sb_solver::Launch> L = h->program();
    *n_launches = (int32_t)L.size();
    if (ops6) {
      if (capacity < L.size()) throw std::string("capacity too small");
      for (size_t i = 0; i < L.size(); i++) {
        int32_t *o = ops6 + 6 * i;
        o[0] = (int32_t)L[i].kind; o[1] = L[i].arg; o[2] = (int32_t)L[i].n_seg; o[3] = (int32_t)L[i].reps;
        o[4] = L[i].pre; o[5] = L[i].post;
      }
    }
    return SB_OK;
  });
}

/* Tile membership, for tests of the host plan: for pass p writes, per caller vertex,
   the tile that stages it (or -1).  tile_of may be NULL to query n_tiles only. */
int sb_get_tiles(sb_handle h, uint32_t pass, int32_t *tile_of, uint32_t *n_tiles) {
  NEED_HANDLE(h);
  const Plan &P = h->plan;
  if (pass >= P.passes.size()) { h->err = "no such pass"; return SB_E_ARG; }
  const TilePass &tp = P.passes[pass];
  if (n_tiles) *n_tiles = tp.n_tiles();
  if (tile_of) {
    for (uint32_t i = 0; i < P.V; i++) tile_of[i] = -1;
    for (uint32_t t = 0; t < tp.n_tiles(); t++)
      for (uint32_t k = tp.vert_off[t]; k < tp.vert_off[t + 1]; k++) {
        uint32_t d = tp.contiguous ? k : tp.tile_verts[k];
        tile_of[P.perm[d]] = (int32_t)t;
      }
  }
  return SB_OK;
}

/* Debug aid (host only): decode the device constraint streams the way the kernel reads them -- tile by
   tile, round by round, thread by thread -- and compare every record with the schedule bookkeeping that
   sb_get_schedule exports (same constraint, same vertex roles, same rest value).  Returns the number of
   mismatching or misplaced records in *n_bad. */
int sb_debug_verify_streams(sb_handle h, uint64_t *n_bad, uint64_t *wavefronts, uint64_t *wavefronts_ideal) {
  NEED_HANDLE(h);
  if (!n_bad) return SB_E_ARG;
  const Plan &P = h->plan;
  uint64_t bad = 0, wf = 0, wf_ideal = 0;
  for (const TilePass &tp : P.passes) {
    const uint32_t bt = tp.bt, W = tp.width;
    for (uint32_t t = 0; t < tp.n_tiles(); t++) {
      const U4 meta = tp.rounds[t];
      const uint32_t ncol = tp.col_off[t + 1] - tp.col_off[t];
      if (meta.y != tp.n_ecol[t] || meta.y + meta.z != ncol) { bad++; continue; }
      const uint32_t v0 = tp.vert_off[t];
      auto dev_id = [&](uint32_t local) { return tp.contiguous ? v0 + local : tp.tile_verts[v0 + local]; };
      uint64_t at = tp.ent_off[t];
      for (uint32_t c = 0; c < ncol; c++) {
        const bool tet = c >= meta.y;
        const uint32_t cnt = tp.col_cnt[tp.col_off[t] + c];
        const uint32_t *rw = tp.stream.data() + ((size_t)meta.x * 4 + (size_t)c * 4 * W * bt);
        // shared-memory model: one LDS.128 per vertex slot and sub-record; a quarter-warp (8 lanes) costs as
        // many wavefronts as its most crowded 16-byte bank group holds distinct addresses
        const bool bitet = W == 2;
        const uint32_t n_sub = tet ? 1u : 2 * W, n_slot = tet ? (bitet ? 5u : 4u) : 2u;
        auto vid = [&](const uint32_t *r, uint32_t slot) -> uint32_t { // local vertex id in a record's vertex slot
          if (!tet) return slot ? r[0] >> 16 : r[0] & 0xffffu;
          if (slot == 4) return r[4] & 0xffffu;
          const uint32_t wd = r[slot / 2];
          return slot & 1 ? wd >> 16 : wd & 0xffffu;
        };
        for (uint32_t sub = 0; sub < n_sub; sub++)
          for (uint32_t slot = 0; slot < n_slot; slot++)
            for (uint32_t o = 0; o < bt; o += 8) {
              uint32_t ids[8], nid = 0, mult[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              for (uint32_t l = 0; l < 8; l++) {
                const uint32_t *r = rw + (size_t)(o + l) * 4 * W + (tet ? 0 : 2 * sub);
                if (slot == 4 && r[5] == 0x7fc00000u) continue; // no mate: no fifth access
                const uint32_t id = vid(r, slot);
                bool dup = false;
                for (uint32_t q = 0; q < nid; q++) dup |= ids[q] == id;
                if (!dup) { ids[nid++] = id; mult[id & 7]++; }
              }
              uint32_t m = 0;
              for (uint32_t q = 0; q < 8; q++) m = std::max(m, mult[q]);
              wf += m;
              wf_ideal += nid ? 1 : 0;
            }
        uint32_t seen = 0;
        for (uint32_t thr = 0; thr < bt; thr++)
          for (uint32_t sub = 0; sub < n_sub; sub++) {
            const uint32_t *r = rw + (size_t)thr * 4 * W + (tet ? 0 : 2 * sub);
            const uint32_t k = 8u * n_sub * (thr / 8u) + 8u * sub + (thr & 7u); // inverse of round_slot
            const bool pad = (r[0] & 0xffffu) == (r[0] >> 16);
            if (k >= cnt) { bad += !pad; continue; }
            if (pad) { bad++; continue; }
            seen++;
            const int32_t ent = tp.ents[at + k];
            if ((ent < 0) != tet) { bad++; continue; }
            if (!tet) {
              const int32_t a = P.edges[2 * (size_t)ent], b = P.edges[2 * (size_t)ent + 1];
              float L0; std::memcpy(&L0, &r[1], 4);
              bad += P.perm[dev_id(r[0] & 0xffffu)] != (uint32_t)a || P.perm[dev_id(r[0] >> 16)] != (uint32_t)b ||
                     std::memcmp(&L0, &P.rest_len[ent], 4) != 0;
            } else {
              const int32_t id = ent & 0x7fffffff;
              const int32_t *q = &P.tet_roles[4 * (size_t)id];
              const uint32_t l[4] = {r[0] & 0xffffu, r[0] >> 16, r[1] & 0xffffu, r[1] >> 16};
              bool ok = std::memcmp(&r[2], &P.rest_vol6[id], 4) == 0 && P.tet_lead[id];
              for (int j = 0; j < 4; j++) ok = ok && P.perm[dev_id(l[j])] == (uint32_t)q[j];
              // attached edges: rest lengths in the record / aux stream, and they do join roles (0,1) / (2,3)
              const float a23 = tp.aux[(size_t)meta.w + (size_t)(c - meta.y) * W * bt + (size_t)thr * W];
              auto joins = [&](int32_t e, int32_t owner, int32_t u, int32_t v) {
                return (P.edges[2 * (size_t)e] == std::min(u, v)) && (P.edges[2 * (size_t)e + 1] == std::max(u, v)) &&
                       P.edge_owner[e] == owner;
              };
              auto slot_ok = [&](int32_t e, int32_t owner, uint32_t bits, int32_t u, int32_t v) {
                if (e >= 0) return std::memcmp(&bits, &P.rest_len[e], 4) == 0 && joins(e, owner, u, v);
                return bits == 0x7fc00000u;
              };
              uint32_t a23b; std::memcpy(&a23b, &a23, 4);
              ok = ok && slot_ok(P.tet_e01[id], id, r[3], q[0], q[1]);
              ok = ok && (P.tet_e23[id] >= 0 ? slot_ok(P.tet_e23[id], id, a23b, q[2], q[3]) : a23 != a23);
              const int32_t mate = P.tet_mate[id];
              if (bitet) {
                if (mate >= 0) {
                  // the mate runs on the registers (4, 2, 1, 3): its roles must be (p4, p2, p1, p3)
                  const int32_t *m = &P.tet_roles[4 * (size_t)mate];
                  const uint32_t lm[4] = {r[4] & 0xffffu, l[2], l[1], l[3]};
                  for (int j = 0; j < 4; j++) ok = ok && P.perm[dev_id(lm[j])] == (uint32_t)m[j];
                  ok = ok && std::memcmp(&r[5], &P.rest_vol6[mate], 4) == 0 && !P.tet_lead[mate] && P.tet_mate[mate] == id;
                  ok = ok && slot_ok(P.tet_e01[mate], mate, r[6], m[0], m[1]) && slot_ok(P.tet_e23[mate], mate, r[7], m[2], m[3]);
                } else {
                  ok = ok && r[5] == 0x7fc00000u && r[6] == 0x7fc00000u && r[7] == 0x7fc00000u;
                }
              } else {
                ok = ok && mate < 0;
              }
              bad += !ok;
            }
          }
        bad += seen != cnt;
        at += cnt;
      }
      bad += at != tp.ent_off[t + 1];
    }
  }
  *n_bad = bad;
  if (wavefronts) *wavefronts = wf;
  if (wavefronts_ideal) *wavefronts_ideal = wf_ideal;
  return SB_OK;
}

int sb_time_frames(sb_handle h, int32_t n_frames, float dt, float *elapsed_ms) {
  NEED_DEVICE(h);
  if (n_frames < 1 || !elapsed_ms) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    h->refresh_params(dt > 0 ? dt : h->prm.dt);
    CK(cudaEventRecord(h->ev0, h->stream));
    for (int f = 0; f < n_frames; f++) h->step(dt);
    CK(cudaEventRecord(h->ev1, h->stream));
    CK(cudaEventSynchronize(h->ev1));
    CK(cudaEventElapsedTime(elapsed_ms, h->ev0, h->ev1));
    return SB_OK;
  });
}

int sb_time_kernel(sb_handle h, int32_t which, int32_t reps, float *avg_ms) {
  NEED_DEVICE(h);
  if (reps < 1 || !avg_ms) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    *avg_ms = h->time_kernel(which, reps);
    return SB_OK;
  });
}

/* ---- phased stepping for partitioned meshes (one rank of a multi-GPU body) -----------------
   The caller interleaves these with its halo exchange; everything is enqueued on the handle's
   stream and nothing synchronises, so the sequence can be captured in a CUDA graph. */
int sb_set_stream(sb_handle h, void *stream) {
  NEED_DEVICE(h);
  if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); h->own_stream = false; }
  h->stream = (cudaStream_t)stream;
  return SB_OK;
}

int sb_prepare(sb_handle h, float dt) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    if (!(dt > 0)) dt = h->prm.dt;
    h->refresh_params(dt);
    CK(cudaStreamSynchronize(h->stream));
    return SB_OK;
  });
}

int sb_enqueue(sb_handle h, int32_t op, int32_t arg) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    if (h->cur_dt < 0 || h->prm_dirty) throw std::string("call sb_prepare after creating the handle or changing parameters");
    switch (op) {
      case SB_OP_PREDICT: h->launch_predict(h->stream); break;
      case SB_OP_PROJECT:
        if (arg < 0) { h->launch_group(0, h->stream); h->launch_group(1, h->stream); }
        else if (arg <= 1) h->launch_group(arg, h->stream);
        else throw std::string("no such constraint group");
        break;
      case SB_OP_FINISH: h->launch_finish(h->stream); break;
      case SB_OP_NORMALS: h->launch_normals(h->stream); break;
      case SB_OP_EXCHANGE: h->exchange(arg, h->stream); break;
      case SB_OP_HALO_SEND: h->halo_send(arg, h->stream); break;
      case SB_OP_HALO_RECV: h->halo_recv(arg, h->stream); break;
      case SB_OP_PASS:
        if (arg < 0 || (size_t)arg >= h->passes.size()) throw std::string("no such tile pass");
        h->launch_pass((size_t)arg, h->stream);
        break;
      case SB_OP_LAUNCH: {
        const std::vector<sb_solver::Launch> L = h->program();
        if (arg < 0 || (size_t)arg >= L.size()) throw std::string("no such launch in the frame program");
        h->run_launch(L[(size_t)arg], h->stream);
        break;
      }
      default: throw std::string("unknown op");
    }
    CK(cudaGetLastError());
    return SB_OK;
  });
}

int sb_halo_set(sb_handle h, int32_t list_id, const int32_t *vertex_ids, uint32_t n) {
  NEED_DEVICE(h);
  if (n && !vertex_ids) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    std::vector<uint32_t> slots(n);
    for (uint32_t k = 0; k < n; k++) {
      if (vertex_ids[k] < 0 || (uint32_t)vertex_ids[k] >= h->plan.V) throw std::string("halo vertex id out of range");
      slots[k] = h->plan.inv[vertex_ids[k]];
    }
    if (h->links.count(list_id)) throw std::string("this halo list is linked to a peer already (sb_halo_alloc / sb_halo_connect): its size is fixed");
    CK(cudaStreamSynchronize(h->stream));
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.second); // a captured frame holds the old index buffer and count
    h->graphs.clear();
    h->halo[list_id].upload(slots, &h->dev_bytes);
    return SB_OK;
  });
}

/* positions (float4) of the list's vertices -> n * 16 bytes of device memory at dst, in list order */
int sb_halo_pack(sb_handle h, int32_t list_id, void *dst_device) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    auto it = h->halo.find(list_id);
    if (it == h->halo.end() || !dst_device) throw std::string("unknown halo list or NULL buffer");
    const uint32_t n = (uint32_t)it->second.n;
    if (n) k_gather4<<<h->grid_for(n, 256), 256, 0, h->stream>>>(n, it->second.p, h->x.p, (float4 *)dst_device);
    CK(cudaGetLastError());
    return SB_OK;
  });
}

int sb_halo_unpack(sb_handle h, int32_t list_id, const void *src_device) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    auto it = h->halo.find(list_id);
    if (it == h->halo.end() || !src_device) throw std::string("unknown halo list or NULL buffer");
    const uint32_t n = (uint32_t)it->second.n;
    if (n) k_scatter4<<<h->grid_for(n, 256), 256, 0, h->stream>>>(n, it->second.p, (const float4 *)src_device, h->x.p);
    CK(cudaGetLastError());
    return SB_OK;
  });
}

/* ---- halo exchange over peer memory: set-up ------------------------------------------------
   sb_halo_alloc: receive buffer for a registered list (n float4 + a flag word); *base_out is its
   device address (export it with sb_ipc_export for another process).  sb_halo_connect: where this
   rank SENDS the list's vertices (the neighbour's receive buffer for the matching list; a pointer
   in this process' address space: the neighbour's own pointer when both ranks live in one process,
   else the result of sb_ipc_open).  Once any link exists, sb_step runs the exchanges inside the
   frame graph: list 0 = ghosts (received in exchange A, sent back in B), list 1 = own vertices that
   are ghosts on the lower rank (sent in A, received in B). */
int sb_halo_alloc(sb_handle h, int32_t list_id, void **base_out, uint64_t *bytes_out) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    auto it = h->halo.find(list_id);
    if (it == h->halo.end()) throw std::string("register the list with sb_halo_set first");
    sb_solver::HaloLink &L = h->links[list_id];
    const size_t bytes = it->second.n * 16 + 16;
    L.recv.alloc(bytes, &h->dev_bytes);
    CK(cudaMemset(L.recv.p, 0, bytes));
    if (!L.ctl_recv.p) { L.ctl_recv.alloc(4, &h->dev_bytes); CK(cudaMemset(L.ctl_recv.p, 0, 16)); }
    if (base_out) *base_out = L.recv.p;
    if (bytes_out) *bytes_out = bytes;
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
    h->graphs.clear();
    return SB_OK;
  });
}

int sb_halo_connect(sb_handle h, int32_t list_id, void *peer_base) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    auto it = h->halo.find(list_id);
    if (it == h->halo.end() || !peer_base) throw std::string("unknown halo list or NULL peer buffer");
    sb_solver::HaloLink &L = h->links[list_id];
    L.peer_buf = (float4 *)peer_base;
    L.peer_flag = (uint32_t *)((unsigned char *)peer_base + it->second.n * 16);
    if (!L.ctl_send.p) { L.ctl_send.alloc(4, &h->dev_bytes); CK(cudaMemset(L.ctl_send.p, 0, 16)); }
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.second);
    h->graphs.clear();
    return SB_OK;
  });
}

int sb_dist_setup(sb_handle h, int32_t rank, int32_t n_ranks, void **x_base_out, void **ctl_base_out) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    h->dist_setup(rank, n_ranks);
    if (x_base_out) *x_base_out = h->x.p;
    if (ctl_base_out) *ctl_base_out = h->dist_ctl.p;
    return SB_OK;
  });
}

int sb_dist_connect(sb_handle h, int32_t peer, void *peer_x, void *peer_ctl) {
  NEED_DEVICE(h);
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    h->dist_connect(peer, peer_x, peer_ctl);
    return SB_OK;
  });
}

int sb_dist_owned(sb_handle h, uint8_t *owned_V, uint32_t *tiles_per_pass8) {
  NEED_DEVICE(h);
  if (!h->dist.ctl) { h->err = "not a distributed handle"; return SB_E_STATE; }
  if (owned_V) {
    std::memset(owned_V, 0, h->plan.V);
    for (uint32_t d = h->dist.slab_lo[h->dist.rank]; d < h->dist.slab_lo[h->dist.rank + 1]; d++) owned_V[h->plan.perm[d]] = 1;
  }
  if (tiles_per_pass8)
    for (size_t k = 0; k < 8; k++) tiles_per_pass8[k] = k < h->passes.size() ? h->passes[k].grid : 0u;
  return SB_OK;
}

/* Host only: symbolic replay of one frame's hand-overs (see the header). */
int sb_dist_verify(sb_handle h, int32_t n_ranks, uint64_t *n_stale, uint64_t *n_not_home, uint64_t *n_unordered, uint64_t *n_crossings) {
  NEED_HANDLE(h);
  return guarded(h, [&]() -> int {
    const Plan &P = h->plan;
    const size_t np = P.passes.size();
    std::vector<std::vector<std::vector<uint32_t>>> tiles(n_ranks); // [rank][pass] -> tiles, zone tiles first
    std::vector<std::vector<uint32_t>> n_zone(n_ranks);
    std::vector<uint32_t> tuple;
    DistDev D{};
    std::vector<uint32_t> nbr((size_t)n_ranks, 0u); // whose epoch a rank waits for, and to whom it publishes its own
    for (int r = 0; r < n_ranks; r++) {
      std::vector<uint32_t> tup;
      h->dist_layout(r, n_ranks, D, tiles[r], n_zone[r], &tup);
      nbr[(size_t)r] = D.nbr_mask;
      if (r == 0) tuple = tup;
      else if (tup != tuple) throw std::string("the ranks disagree on who runs which tile");
    }
    // two ranks between which a value travels (or whose launches read each other's arrays) must poll each other
    auto linked = [&](uint32_t a, uint32_t b) { return a == b || ((nbr[a] >> b & 1u) && (nbr[b] >> a & 1u)); };
    auto owner_of = [&](uint32_t dev) {
      uint32_t r = 0;
      while (r + 1 < D.n_ranks && dev >= D.slab_lo[r + 1]) r++;
      return r;
    };
    const int partial = np > P.n_tilings ? (int)np - 1 : -1;
    std::vector<uint32_t> holder(P.V);   // the rank whose array holds the current value of a vertex
    for (uint32_t d = 0; d < P.V; d++) holder[d] = owner_of(d);
    uint64_t stale = 0, away = 0, unordered = 0, crossings = 0;
    auto check_home = [&]() {
      for (uint32_t d = 0; d < P.V; d++) away += holder[d] != owner_of(d);
    };
    for (const sb_solver::Launch &l : h->program()) {
      if (l.kind != sb_solver::Launch::PASS) { // predict / finish / normals (and anything else) work on the owners' copies
        check_home();
        continue;
      }
      const TilePass &tp = P.passes[(size_t)l.arg];
      std::vector<uint32_t> next(holder);
      for (int r = 0; r < n_ranks; r++) {
        const auto &mine = tiles[r][(size_t)l.arg];
        for (size_t j = 0; j < mine.size(); j++) {
          const uint32_t t = mine[j];
          const bool zone = j < n_zone[r][(size_t)l.arg];
          for (uint32_t i = tp.vert_off[t]; i < tp.vert_off[t + 1]; i++) {
            const uint32_t d = tp.contiguous ? i : tp.tile_verts[i];
            if (holder[d] != (uint32_t)r) stale++;
            const uint32_t tup = tuple[d];
            const bool covered = (int)l.next_pass != partial || (tup >> 15 & 1u);
            const uint32_t dst = (tup >> (3 * (covered ? l.next_pass : l.next_full_pass))) & 7u;
            // a vertex that came from, or goes to, another rank needs the epoch handshake: only zone tiles do it
            if (dst != (uint32_t)r && (!zone || !linked((uint32_t)r, dst))) unordered++;
            crossings += dst != (uint32_t)r;
            next[d] = dst;
          }
        }
      }
      // a vertex that arrived from another rank must be loaded by a zone tile of the NEXT launch: checked there through
      // `stale` (wrong place) and here (right place, but would the loader wait for the sender?)
      holder.swap(next);
    }
    check_home();
    // the loader side of the handshake: replay once more and look at who loads what a foreign rank stored
    {
      std::vector<uint32_t> prev_writer(P.V, 0xffffffffu); // rank that stored the vertex last (0xffffffff: nobody yet this frame)
      for (const sb_solver::Launch &l : h->program()) {
        if (l.kind != sb_solver::Launch::PASS) continue;
        const TilePass &tp = P.passes[(size_t)l.arg];
        std::vector<uint32_t> writer(prev_writer);
        for (int r = 0; r < n_ranks; r++) {
          const auto &mine = tiles[r][(size_t)l.arg];
          for (size_t j = 0; j < mine.size(); j++) {
            const uint32_t t = mine[j];
            const bool zone = j < n_zone[r][(size_t)l.arg];
            for (uint32_t i = tp.vert_off[t]; i < tp.vert_off[t + 1]; i++) {
              const uint32_t d = tp.contiguous ? i : tp.tile_verts[i];
              if (prev_writer[d] != 0xffffffffu && prev_writer[d] != (uint32_t)r && (!zone || !linked(prev_writer[d], (uint32_t)r))) unordered++;
              writer[d] = (uint32_t)r;
            }
          }
        }
        prev_writer.swap(writer);
      }
    }
    // the normals launch: the owner of a surface vertex reads the other vertices of its triangles in THEIR owners' arrays.
    // Every tile that holds a vertex of a triangle cut between owners must be a zone tile of its runner (its last store
    // is then published before the readers start, and its next load waits for them), and the readers must be linked
    // with every rank that touches the vertex (the last holder stores it home, not necessarily the owner).
    {
      std::vector<std::vector<uint8_t>> zone_tile(np);
      std::vector<std::vector<uint8_t>> runner(np);
      for (size_t k = 0; k < np; k++) {
        zone_tile[k].assign(P.passes[k].n_tiles(), 0);
        runner[k].assign(P.passes[k].n_tiles(), 0xff);
        for (int r = 0; r < n_ranks; r++)
          for (size_t j = 0; j < tiles[(size_t)r][k].size(); j++) {
            runner[k][tiles[(size_t)r][k][j]] = (uint8_t)r;
            if (j < n_zone[(size_t)r][k]) zone_tile[k][tiles[(size_t)r][k][j]] = 1;
          }
      }
      std::vector<std::vector<uint32_t>> tile_of(np, std::vector<uint32_t>(P.V, 0xffffffffu));
      for (size_t k = 0; k < np; k++) {
        const TilePass &tp = P.passes[k];
        for (uint32_t t = 0; t < tp.n_tiles(); t++)
          for (uint32_t i = tp.vert_off[t]; i < tp.vert_off[t + 1]; i++) tile_of[k][tp.contiguous ? i : tp.tile_verts[i]] = t;
      }
      for (size_t f = 0; f + 2 < P.tris.size(); f += 3) {
        uint32_t d[3], own = 0;
        for (int j = 0; j < 3; j++) {
          d[j] = P.inv[(size_t)P.tris[f + j]];
          own |= 1u << owner_of(d[j]);
        }
        if (!(own & (own - 1))) continue;
        for (int j = 0; j < 3; j++)
          for (size_t k = 0; k < np; k++) {
            const uint32_t t = tile_of[k][d[j]];
            if (t == 0xffffffffu || runner[k][t] == 0xff) continue;
            if (!zone_tile[k][t]) unordered++;
            for (uint32_t o = 0; o < (uint32_t)n_ranks; o++)
              if ((own >> o & 1u) && !linked(o, runner[k][t])) unordered++;
          }
      }
    }
    if (n_stale) *n_stale = stale;
    if (n_not_home) *n_not_home = away;
    if (n_unordered) *n_unordered = unordered;
    if (n_crossings) *n_crossings = crossings;
    return SB_OK;
  });
}

/* Host only (works on an sb_plan handle): what sb_dist_setup(rank, n_ranks) would decide. */
int sb_dist_layout(sb_handle h, int32_t rank, int32_t n_ranks, uint8_t *owned_V, int32_t *tile_owner_pass, uint32_t pass) {
  NEED_HANDLE(h);
  return guarded(h, [&]() -> int {
    DistDev D;
    std::vector<std::vector<uint32_t>> tiles;
    std::vector<uint32_t> n_zone;
    h->dist_layout(rank, n_ranks, D, tiles, n_zone);
    if (owned_V) {
      std::memset(owned_V, 0, h->plan.V);
      for (uint32_t d = D.slab_lo[rank]; d < D.slab_lo[rank + 1]; d++) owned_V[h->plan.perm[d]] = 1;
    }
    if (tile_owner_pass) { // tile_owner_pass[t] = 0: another rank's tile; j + 1 > 0: CTA j of this rank's launch, an interior tile;
                           // -(j + 1) < 0: CTA j, a zone tile (zone tiles come first)
      if (pass >= tiles.size()) throw std::string("no such pass");
      for (uint32_t t = 0; t < h->plan.passes[pass].n_tiles(); t++) tile_owner_pass[t] = 0;
      for (size_t j = 0; j < tiles[pass].size(); j++) tile_owner_pass[tiles[pass][j]] = j < n_zone[pass] ? -(int32_t)(j + 1) : (int32_t)(j + 1);
    }
    return SB_OK;
  });
}

int sb_dist_error(sb_handle h, int32_t *out) {
  NEED_DEVICE(h);
  if (!out) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    *out = 0;
    if (!h->dist.ctl) return SB_OK;
    uint32_t e = 0;
    CK(cudaMemcpyAsync(&e, h->dist_ctl.p + 2, sizeof e, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *out = (int32_t)e;
    return SB_OK;
  });
}

/* 1 if a receive ever timed out waiting for its neighbour (the step's results are then invalid) */
int sb_halo_error(sb_handle h, int32_t *out) {
  NEED_DEVICE(h);
  if (!out) return SB_E_ARG;
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    CK(cudaStreamSynchronize(h->stream));
    *out = 0;
    for (auto &kv : h->links)
      if (kv.second.ctl_recv.p) {
        uint32_t c[4];
        CK(cudaMemcpy(c, kv.second.ctl_recv.p, 16, cudaMemcpyDeviceToHost));
        if (c[2]) *out = 1;
      }
    return SB_OK;
  });
}

int sb_ipc_export(void *device_ptr, unsigned char *handle64) {
  if (!device_ptr || !handle64) return SB_E_ARG;
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t hd;
  if (cudaIpcGetMemHandle(&hd, device_ptr) != cudaSuccess) { cudaGetLastError(); return SB_E_CUDA; }
  std::memcpy(handle64, &hd, 64);
  return SB_OK;
}

int sb_ipc_open(int32_t device, const unsigned char *handle64, void **ptr_out) {
  if (!handle64 || !ptr_out) return SB_E_ARG;
  cudaIpcMemHandle_t hd;
  std::memcpy(&hd, handle64, 64);
  if (cudaSetDevice(device) != cudaSuccess || cudaIpcOpenMemHandle(ptr_out, hd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    g_create_error = cudaGetErrorString(cudaGetLastError());
    return SB_E_CUDA;
  }
  return SB_OK;
}

/* lumped inverse masses of a whole mesh, exactly as sb_create derives them (host only): a
   partitioner needs the global values because a rank does not see every tet of its vertices */
int sb_lumped_inv_mass(const float *pos_xyz, uint32_t n_verts, const int32_t *tets, uint32_t n_tets, float density, float *out) {
  if (!pos_xyz || !out || (n_tets && !tets) || !(density > 0)) return SB_E_ARG;
  for (size_t i = 0; i < 4 * (size_t)n_tets; i++)
    if (tets[i] < 0 || (uint32_t)tets[i] >= n_verts) return SB_E_ARG;
  return guarded(nullptr, [&]() -> int {
    sb::lumped_inv_mass_into(pos_xyz, n_verts, tets, n_tets, density, out);
    return SB_OK;
  });
}

/* Debug aid (not part of the component surface): run tile pass `pass` once with per-CTA
   timestamps and copy them out: out[cta * 80 + k], k = 0 start, 1 loop start, 2 loop end,
   3 CTA end, 4 + i start of chunk i (globaltimer ns; first 64 CTAs). */
int sb_debug_trace_pass(sb_handle h, uint32_t pass, unsigned long long *out, uint32_t n_words) {
  NEED_DEVICE(h);
  NOT_DISTRIBUTED(h, "sb_debug_trace_pass");
  if (pass >= h->passes.size() || !out || n_words < 64u * 80u + 256u + 3u * 4096u) { h->err = "bad trace arguments"; return SB_E_ARG; }
  return guarded(h, [&]() -> int {
    CK(cudaSetDevice(h->device));
    if (h->cur_dt < 0) h->refresh_params(h->prm.dt);
    DevBuf<unsigned long long> buf;
    buf.alloc(64 * 80 + 256 + 3 * 4096, nullptr);
    CK(cudaMemsetAsync(buf.p, 0, (64 * 80 + 256 + 3 * 4096) * 8, h->stream));
    PassBufs &pb = h->passes[pass];
    h->launch_pass(pass, h->stream); // warm
    pb.dev.trace = buf.p;
    h->launch_pass(pass, h->stream);
    pb.dev.trace = nullptr;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, buf.p, (64 * 80 + 256 + 3 * 4096) * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return SB_OK;
  });
}

const char *sb_last_error(sb_handle h) {
  if (!h) return g_create_error.c_str();
  return h->err.c_str();
}

} // extern "C"
