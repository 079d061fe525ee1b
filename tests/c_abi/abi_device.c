/*
 * abi_device.c -- the DEVICE entry points of the C ABI driven from plain C (gcc -std=c11 against
 * include/softbody_b200.h alone, linked to libsoftbody_b200.so): surface -> tets -> sb_create on cuda:0 -> sb_step x N ->
 * sb_read_surface / sb_get_state -> sb_destroy, the call sequence a P/Invoke, cgo or JNI host makes once per frame.
 * Prints an FNV-1a checksum of the final state and of the surface read-back; tests/test_gpu_c_abi.py runs the same
 * sequence through the Python mirror (ctypes) and compares the two lines.   usage: abi_device <frames> <spacing>
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "softbody_b200.h"

#define CHECK(cond)                                                                                    \
  do {                                                                                                 \
    if (!(cond)) {                                                                                     \
      printf("FAILED %s:%d: %s (%s)\n", __FILE__, __LINE__, #cond, h ? sb_last_error(h) : "no handle"); \
      return 1;                                                                                        \
    }                                                                                                  \
  } while (0)

static unsigned long long fnv1a(const void *p, size_t n, unsigned long long h) {
  const unsigned char *b = (const unsigned char *)p;
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 1099511628211ull;
  return h;
}

/* unit cube surface lifted 5 cm off the ground, 8 vertices, 12 outward triangles */
static const float CUBE_POS[24] = {0, 0.05f, 0, 1, 0.05f, 0, 0, 1.05f, 0, 1, 1.05f, 0, 0, 0.05f, 1, 1, 0.05f, 1, 0, 1.05f, 1, 1, 1.05f, 1};
static const int32_t CUBE_TRI[36] = {0, 2, 1, 1, 2, 3, 4, 5, 6, 5, 7, 6, 0, 1, 4, 1, 5, 4,
                                     2, 6, 3, 3, 6, 7, 0, 4, 2, 2, 4, 6, 1, 3, 5, 3, 7, 5};

int main(int argc, char **argv) {
  const int frames = argc > 1 ? atoi(argv[1]) : 12;
  const float spacing = argc > 2 ? (float)atof(argv[2]) : 0.125f;
  sb_handle h = NULL;
  sb_tetmesh_handle tm = NULL;
  CHECK(sb_tetmesh_from_surface(CUBE_POS, 8, CUBE_TRI, 12, spacing, &tm) == SB_OK);
  sb_mesh_desc desc;
  memset(&desc, 0, sizeof desc);
  CHECK(sb_tetmesh_desc(tm, &desc) == SB_OK);
  desc.device = 0;
  desc.tile_cap = 256; /* several tiles and all four tilings on this small mesh */
  sb_params prm;
  sb_default_params(&prm);
  prm.stiffness_distance = 2.0e5f;
  CHECK(sb_create(&desc, &prm, &h) == SB_OK);
  const uint32_t V = desc.n_verts;
  for (int f = 0; f < frames; f++) CHECK(sb_step(h, 0.0f) == SB_OK);
  CHECK(sb_synchronize(h) == SB_OK);
  uint32_t ns = 0;
  CHECK(sb_surface_vertices(h, NULL, 0, &ns) == SB_OK && ns > 0);
  float *x4 = malloc(sizeof(float) * 4 * V), *v4 = malloc(sizeof(float) * 4 * V);
  float *sp = malloc(sizeof(float) * 3 * ns), *sn = malloc(sizeof(float) * 3 * ns);
  CHECK(x4 && v4 && sp && sn);
  CHECK(sb_get_state(h, x4, v4, V) == SB_OK);
  CHECK(sb_read_surface(h, sp, sn, ns) == SB_OK);
  float ymin = 1e30f;
  for (uint32_t i = 0; i < V; i++) ymin = x4[4 * i + 1] < ymin ? x4[4 * i + 1] : ymin;
  unsigned long long a = fnv1a(x4, sizeof(float) * 4 * V, 1469598103934665603ull);
  a = fnv1a(v4, sizeof(float) * 4 * V, a);
  unsigned long long b = fnv1a(sp, sizeof(float) * 3 * ns, 1469598103934665603ull);
  b = fnv1a(sn, sizeof(float) * 3 * ns, b);
  printf("V=%u ns=%u frames=%d state=%016llx surface=%016llx min_y=%.6f\n", V, ns, frames, a, b, (double)ymin);
  CHECK(sb_destroy(h) == SB_OK);
  h = NULL;
  sb_tetmesh_free(tm);
  free(x4); free(v4); free(sp); free(sn);
  printf("ok\n");
  return 0;
}
