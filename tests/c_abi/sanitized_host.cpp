// sanitized_host.cpp -- the host-only C++ of the library (planner: attachment, colouring + recolouring, rim merge,
// streams; ingest: surface -> tets, snapping, skin binding, files) built with -fsanitize=address,undefined and run on a
// few lattice meshes (tests/test_sanitizers.py).  compute-sanitizer is closed on the GPU pool, so this is the
// memory-safety evidence for the code that runs inside sb_create / sb_plan / sb_tetmesh_*.
// usage: sanitized_host nx ny nz tile_cap tmpdir
#include <cstdio>
#include <string>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "plan.h"
#include "softbody_b200.h"
using namespace sb;
static const int A[5][4] = {{0,3,5,6},{1,0,3,5},{2,0,6,3},{4,0,5,6},{7,3,6,5}};
int main(int argc, char **argv) {
  int nx = argc > 1 ? atoi(argv[1]) : 20, ny = argc > 2 ? atoi(argv[2]) : nx, nz = argc > 3 ? atoi(argv[3]) : nx;
  int cap = argc > 4 ? atoi(argv[4]) : 0;
  const std::string tmp = argc > 5 ? argv[5] : "/tmp", msh = tmp + "/sanitized.msh", node = tmp + "/sanitized.node";
  std::vector<float> pos; std::vector<int32_t> tets;
  unsigned rng = 12345;
  for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < nx; i++) {
    float jit[3]; for (auto &q : jit) { rng = rng * 1664525u + 1013904223u; q = ((rng >> 8) / 16777216.0f - 0.5f) * 0.002f; }
    pos.push_back(i * 0.01f + jit[0]); pos.push_back(j * 0.01f + jit[1] + 0.01f); pos.push_back(k * 0.01f + jit[2]);
  }
  for (int k = 0; k + 1 < nz; k++) for (int j = 0; j + 1 < ny; j++) for (int i = 0; i + 1 < nx; i++) {
    int c[8]; for (int b = 0; b < 8; b++) c[b] = (i + (b & 1)) + nx * ((j + ((b >> 1) & 1)) + ny * (k + ((b >> 2) & 1)));
    int flip = (i + j + k) & 1;
    for (int t = 0; t < 5; t++) {
      int q[4]; for (int m = 0; m < 4; m++) q[m] = c[A[t][m] ^ flip];
      // orient
      const float *p0=&pos[3*q[0]],*p1=&pos[3*q[1]],*p2=&pos[3*q[2]],*p3=&pos[3*q[3]];
      double e1[3],e2[3],e3[3]; for(int d=0;d<3;d++){e1[d]=p1[d]-p0[d];e2[d]=p2[d]-p0[d];e3[d]=p3[d]-p0[d];}
      double det=e1[0]*(e2[1]*e3[2]-e2[2]*e3[1])+e1[1]*(e2[2]*e3[0]-e2[0]*e3[2])+e1[2]*(e2[0]*e3[1]-e2[1]*e3[0]);
      if (det < 0) std::swap(q[2], q[3]);
      for (int m = 0; m < 4; m++) tets.push_back(q[m]);
    }
  }
  MeshInput in{pos.data(), tets.data(), nullptr, nullptr, (uint32_t)(pos.size()/3), (uint32_t)(tets.size()/4), 0, 1000.0f};
  PlanOptions opt; opt.tile_cap = cap; opt.threads = 4;
  Plan P;
  std::string err = build_plan(in, opt, P);
  if (!err.empty()) { printf("plan error: %s\n", err.c_str()); return 1; }
  std::vector<int32_t> order; std::vector<int64_t> off;
  P.export_schedule(order, off);
  size_t rounds = 0; for (auto &tp : P.passes) rounds += tp.rounds_total;
  printf("V %u T %u E %u passes %zu tilings %u rounds %zu attached %llu order %zu batches %zu\n", P.V, P.T, P.E, P.passes.size(), P.n_tilings, rounds,
         (unsigned long long)P.edges_attached, order.size(), off.size() - 1);
  // ingest under the sanitizer as well: cube surface -> tets -> snap -> skin -> files
  static const float CP[24] = {0,0,0,1,0,0,0,1,0,1,1,0,0,0,1,1,0,1,0,1,1,1,1,1};
  static const int32_t CT[36] = {0,2,1,1,2,3,4,5,6,5,7,6,0,1,4,1,5,4,2,6,3,3,6,7,0,4,2,2,4,6,1,3,5,3,7,5};
  sb_tetmesh_handle tm = nullptr;
  if (sb_tetmesh_from_surface(CP, 8, CT, 12, 0.13f, &tm) != 0) { printf("ingest: %s\n", sb_ingest_last_error()); return 1; }
  uint32_t moved = 0, V = 0, T = 0, F = 0;
  sb_tetmesh_snap_to_surface(tm, CP, 8, CT, 12, 0.1f, &moved);
  sb_tetmesh_sizes(tm, &V, &T, &F);
  std::vector<float> p(3 * V); std::vector<int32_t> t(4 * T), f(3 * F);
  sb_tetmesh_copy(tm, p.data(), t.data(), f.data());
  std::vector<int32_t> tet_of(8); std::vector<float> w(32);
  sb_skin_compute(p.data(), V, t.data(), T, CP, 8, tet_of.data(), w.data());
  sb_tetmesh_save(tm, msh.c_str()); sb_tetmesh_save(tm, node.c_str());
  sb_tetmesh_handle b1 = nullptr, b2 = nullptr;
  int r1 = sb_tetmesh_load(msh.c_str(), &b1), r2 = sb_tetmesh_load(node.c_str(), &b2);
  printf("ingest V %u T %u F %u moved %u load %d %d\n", V, T, F, moved, r1, r2);
  sb_tetmesh_free(tm); sb_tetmesh_free(b1); sb_tetmesh_free(b2);
  return 0;
}
