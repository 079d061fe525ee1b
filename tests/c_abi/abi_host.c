/*
 * abi_host.c -- the C ABI used from plain C, the way a foreign host (C# P/Invoke, cgo, JNI) binds it: compiled
 * with gcc -std=c11 against include/softbody_b200.h alone and linked to libsoftbody_b200.so.  Exercises every
 * entry point that works without a device (layout guard, planner, ingest, files, snapshots, skin binding) and
 * checks that the device entry points refuse a host-only handle instead of falling back to the CPU.
 * Prints "ok" and returns 0, or the failed check and 1.   (tests/test_c_abi.py builds and runs it.)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "softbody_b200.h"

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);        \
      return 1;                                                       \
    }                                                                 \
  } while (0)

/* unit cube surface, 8 vertices, 12 outward triangles */
static const float CUBE_POS[24] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 1};
static const int32_t CUBE_TRI[36] = {0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4,
                                     2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5};

int main(int argc, char **argv) {
  const char *tmp = argc > 1 ? argv[1] : "/tmp";
  char path[1024];

  /* layout guard: what a managed mirror checks before its first call */
  uint32_t ver = 0, sp = 0, sd = 0, si = 0;
  CHECK(sb_abi_check(&ver, &sp, &sd, &si) == SB_OK);
  CHECK(ver == SB_ABI_VERSION && sp == sizeof(sb_params) && sd == sizeof(sb_mesh_desc) && si == sizeof(sb_info));
  CHECK(sizeof(sb_params) == 48 && sizeof(sb_mesh_desc) == 112 && sizeof(sb_collider) == 48);

  /* surface -> tets */
  sb_tetmesh_handle tm = NULL;
  CHECK(sb_tetmesh_from_surface(CUBE_POS, 8, CUBE_TRI, 12, 0.25f, &tm) == SB_OK);
  uint32_t V = 0, T = 0, F = 0;
  CHECK(sb_tetmesh_sizes(tm, &V, &T, &F) == SB_OK);
  CHECK(V == 125 && T == 5 * 64 && F == 12 * 16);
  CHECK(sb_tetmesh_from_surface(CUBE_POS, 8, CUBE_TRI, 12, -1.0f, &tm) == SB_E_ARG);
  CHECK(strstr(sb_ingest_last_error(), "spacing") != NULL);

  /* files: both formats, and back */
  snprintf(path, sizeof path, "%s/abi_cube.msh", tmp);
  CHECK(sb_tetmesh_save(tm, path) == SB_OK);
  sb_tetmesh_handle back = NULL;
  CHECK(sb_tetmesh_load(path, &back) == SB_OK);
  float *p0 = malloc(sizeof(float) * 3 * V), *p1 = malloc(sizeof(float) * 3 * V);
  int32_t *t0 = malloc(sizeof(int32_t) * 4 * T), *t1 = malloc(sizeof(int32_t) * 4 * T);
  CHECK(sb_tetmesh_copy(tm, p0, t0, NULL) == SB_OK && sb_tetmesh_copy(back, p1, t1, NULL) == SB_OK);
  CHECK(memcmp(p0, p1, sizeof(float) * 3 * V) == 0 && memcmp(t0, t1, sizeof(int32_t) * 4 * T) == 0);
  sb_tetmesh_free(back);
  snprintf(path, sizeof path, "%s/abi_cube.node", tmp);
  CHECK(sb_tetmesh_save(tm, path) == SB_OK && sb_tetmesh_load(path, &back) == SB_OK);
  CHECK(sb_tetmesh_copy(back, p1, t1, NULL) == SB_OK && memcmp(t0, t1, sizeof(int32_t) * 4 * T) == 0);
  sb_tetmesh_free(back);
  CHECK(sb_tetmesh_load("/nonexistent/mesh.msh", &back) == SB_E_ARG);

  /* planner on the ingested mesh: a host-only handle */
  sb_mesh_desc d;
  sb_params prm;
  sb_default_params(&prm);
  CHECK(sb_tetmesh_desc(tm, &d) == SB_OK && d.n_verts == V && d.n_tets == T && d.n_tris == F);
  d.tile_cap = 64;
  sb_handle h = NULL;
  CHECK(sb_plan(&d, &prm, &h) == SB_OK && h != NULL);
  sb_info info;
  CHECK(sb_get_info(h, &info) == SB_OK && info.n_verts == V && info.n_tets == T && info.n_surface_verts == 98);
  int64_t n_order = 0;
  int32_t n_batches = 0;
  CHECK(sb_get_schedule(h, &n_order, NULL, &n_batches, NULL) == SB_OK);
  CHECK(n_order == (int64_t)info.n_edges + (int64_t)info.n_tets && n_batches > 0);
  uint64_t bad = 1;
  CHECK(sb_debug_verify_streams(h, &bad, NULL, NULL) == SB_OK && bad == 0);

  /* one mesh over several GPUs, host-only views: a mesh this small does not plan as balanced tilings and is refused cleanly */
  uint64_t n_stale = 7, n_away = 7, n_unordered = 7;
  CHECK(sb_dist_layout(h, 0, 2, NULL, NULL, 0) == SB_E_ARG && strstr(sb_last_error(h), "distributed mesh") != NULL);
  CHECK(sb_dist_verify(h, 2, &n_stale, &n_away, &n_unordered, NULL) == SB_E_ARG);
  CHECK(sb_dist_verify(h, 9, NULL, NULL, NULL, NULL) == SB_E_ARG && strstr(sb_last_error(h), "ranks") != NULL);

  /* colliders are validated on the host */
  sb_collider col[2];
  memset(col, 0, sizeof col);
  col[0].kind = SB_COLLIDER_CAPSULE; col[0].friction = 0.5f; col[0].p[3] = 0.2f; col[0].p[4] = 1.0f;
  col[1].kind = SB_COLLIDER_BOX; col[1].p[3] = col[1].p[4] = col[1].p[5] = 0.5f; col[1].p[9] = 1.0f;
  CHECK(sb_set_colliders_ex(h, col, 2) == SB_OK);
  col[1].kind = 7;
  CHECK(sb_set_colliders_ex(h, col, 2) == SB_E_ARG);
  col[1].kind = SB_COLLIDER_BOX; col[0].friction = 1.5f;
  CHECK(sb_set_colliders_ex(h, col, 2) == SB_E_ARG);
  CHECK(sb_set_colliders_ex(h, col, 17) == SB_E_ARG);

  /* render mesh binding through the handle == the free function */
  int32_t tet_a[8], tet_b[8];
  float w_a[32], w_b[32];
  CHECK(sb_skin_bind(h, CUBE_POS, 8, CUBE_TRI, 12) == SB_OK);
  CHECK(sb_skin_get_binding(h, tet_a, w_a, 8) == SB_OK);
  CHECK(sb_skin_compute(p0, V, t0, T, CUBE_POS, 8, tet_b, w_b) == SB_OK);
  CHECK(memcmp(tet_a, tet_b, sizeof tet_a) == 0 && memcmp(w_a, w_b, sizeof w_a) == 0);
  for (int i = 0; i < 8; i++) {
    float s = w_a[4 * i] + w_a[4 * i + 1] + w_a[4 * i + 2] + w_a[4 * i + 3];
    CHECK(fabsf(s - 1.0f) < 1e-6f);
  }

  /* no CPU fallback: everything that needs the device refuses a host-only handle */
  float buf[3 * 125];
  CHECK(sb_step(h, 0.0f) == SB_E_STATE);
  CHECK(sb_read_positions(h, buf, V) == SB_E_STATE);
  CHECK(sb_read_skinned(h, buf, NULL, 8) == SB_E_STATE);
  CHECK(sb_save_state(h, "/tmp/never.sbs") == SB_E_STATE);
  CHECK(strstr(sb_last_error(h), "host only") != NULL);

  /* snapshots: file layer alone */
  float *x4 = calloc(4 * (size_t)V, sizeof(float)), *v4 = calloc(4 * (size_t)V, sizeof(float));
  for (uint32_t i = 0; i < V; i++) { memcpy(x4 + 4 * i, p0 + 3 * i, 12); x4[4 * i + 3] = 1.0f; v4[4 * i + 1] = -0.5f * (float)i; }
  const uint64_t topo = sb_topology_hash(V, t0, T);
  snprintf(path, sizeof path, "%s/abi_cube.sbs", tmp);
  CHECK(sb_state_write(path, x4, v4, V, &prm, 42, topo) == SB_OK);
  uint32_t n = 0;
  uint64_t frame = 0, topo2 = 0;
  sb_params prm2;
  CHECK(sb_state_read(path, NULL, NULL, 0, &n, &prm2, &frame, &topo2) == SB_OK);
  CHECK(n == V && frame == 42 && topo2 == topo && memcmp(&prm, &prm2, sizeof prm) == 0);
  float *x4b = malloc(16 * (size_t)V), *v4b = malloc(16 * (size_t)V);
  CHECK(sb_state_read(path, x4b, v4b, V - 1, NULL, NULL, NULL, NULL) == SB_E_ARG); /* buffers too small */
  CHECK(sb_state_read(path, x4b, v4b, V, NULL, NULL, NULL, NULL) == SB_OK);
  CHECK(memcmp(x4, x4b, 16 * (size_t)V) == 0 && memcmp(v4, v4b, 16 * (size_t)V) == 0);

  CHECK(sb_destroy(h) == SB_OK);
  CHECK(sb_tetmesh_free(tm) == SB_OK);
  free(p0); free(p1); free(t0); free(t1); free(x4); free(v4); free(x4b); free(v4b);
  printf("ok\n");
  return 0;
}
