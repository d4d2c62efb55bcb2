/* Plain-C client of include/djets.h: proves the header is C (not C++), and -- on a GPU box -- that the
 * C ABI works without Python: builds a 2x2 dense Float64 operator on a [2,1] grid of two ranks sharing
 * GPU 0, applies it forward and adjoint and checks the result against a scalar loop.
 *   gcc -std=c99 -Wall -Iinclude tests/c_abi_smoke.c -o /tmp/c_abi_smoke -ldl   (library loaded with dlopen) */
#include <dlfcn.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "djets.h"

#define N 37 /* block size: odd on purpose (misaligned columns) */

#define LOAD(name) \
  *(void**)(&p_##name) = dlsym(lib, #name); \
  if (!p_##name) { fprintf(stderr, "missing symbol %s\n", #name); return 2; }

static int32_t (*p_djets_ctx_create)(int32_t, int32_t, const int32_t*, const int32_t*, const uint8_t*, djets_ctx*);
static int32_t (*p_djets_ctx_destroy)(djets_ctx);
static int32_t (*p_djets_op_create)(djets_ctx, int32_t, int64_t, int64_t, const int64_t*, const int64_t*, int32_t, int32_t,
                                    const int32_t*, const int64_t*, const int64_t*, int32_t, djets_op*);
static int32_t (*p_djets_op_set_dense)(djets_op, int64_t, int64_t, const void*, int64_t);
static int32_t (*p_djets_op_finalize)(djets_op);
static int32_t (*p_djets_op_new_vec)(djets_op, int32_t, djets_vec*);
static int32_t (*p_djets_op_mul)(djets_op, djets_vec, djets_vec);
static int32_t (*p_djets_op_mul_adjoint)(djets_op, djets_vec, djets_vec);
static int32_t (*p_djets_op_destroy)(djets_op);
static int32_t (*p_djets_vec_scatter)(djets_vec, const void*);
static int32_t (*p_djets_vec_collect)(djets_vec, void*);
static int32_t (*p_djets_vec_dot)(djets_vec, djets_vec, double*);
static int32_t (*p_djets_vec_destroy)(djets_vec);
static int32_t (*p_djets_device_count)(int32_t*);
static const char* (*p_djets_last_error)(void);
static int32_t (*p_djets_version)(void);

#define CHECK(call) \
  do { int32_t s__ = (call); if (s__ != DJETS_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, s__, p_djets_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const char* path = argc > 1 ? argv[1] : "distributedjets.jl_b200/libdjets_b200.so";
  void* lib = dlopen(path, RTLD_NOW);
  if (!lib) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; }
  LOAD(djets_ctx_create) LOAD(djets_ctx_destroy) LOAD(djets_op_create) LOAD(djets_op_set_dense) LOAD(djets_op_finalize)
  LOAD(djets_op_new_vec) LOAD(djets_op_mul) LOAD(djets_op_mul_adjoint) LOAD(djets_op_destroy) LOAD(djets_vec_scatter)
  LOAD(djets_vec_collect) LOAD(djets_vec_dot) LOAD(djets_vec_destroy) LOAD(djets_device_count) LOAD(djets_last_error)
  LOAD(djets_version)
  printf("libdjets_b200 version %d\n", p_djets_version());
  int32_t ndev = 0;
  if (p_djets_device_count(&ndev) != DJETS_OK || ndev == 0) {
    djets_ctx none = 0;
    int32_t s = p_djets_ctx_create(1, 1, 0, 0, 0, &none);       /* must refuse: no CPU path */
    printf("no GPU: ctx_create -> %d (%s)\n", s, p_djets_last_error());
    return s == DJETS_ERR_CUDA ? 0 : 1;
  }
  const int32_t ranks[2] = {0, 1}, devs[2] = {0, 0};
  djets_ctx ctx;
  CHECK(p_djets_ctx_create(2, 2, ranks, devs, 0, &ctx));
  const int64_t len[2] = {N, N};
  djets_op A;
  CHECK(p_djets_op_create(ctx, DJETS_F64, 2, 2, len, len, 2, 1, 0, 0, 0, 0, &A));
  static double blk[2][2][N * N], x[2 * N], y[2 * N], want[2 * N], got[2 * N];
  for (int i = 0; i < 2; ++i)
    for (int j = 0; j < 2; ++j) {
      for (int k = 0; k < N * N; ++k) blk[i][j][k] = (double)((k * 7 + i * 3 + j) % 11) / 11.0;   /* column-major */
      CHECK(p_djets_op_set_dense(A, i, j, blk[i][j], N));
    }
  CHECK(p_djets_op_finalize(A));
  djets_vec vx, vy, vx2;
  CHECK(p_djets_op_new_vec(A, 0, &vx));
  CHECK(p_djets_op_new_vec(A, 1, &vy));
  CHECK(p_djets_op_new_vec(A, 0, &vx2));
  for (int k = 0; k < 2 * N; ++k) x[k] = 1.0 / (1.0 + k);
  CHECK(p_djets_vec_scatter(vx, x));
  CHECK(p_djets_op_mul(A, vy, vx));
  CHECK(p_djets_vec_collect(vy, got));
  double err = 0, nrm = 0;
  for (int i = 0; i < 2; ++i)
    for (int r = 0; r < N; ++r) {
      double s = 0;
      for (int j = 0; j < 2; ++j)
        for (int c = 0; c < N; ++c) s += blk[i][j][r + c * N] * x[j * N + c];
      want[i * N + r] = s;
      err += (got[i * N + r] - s) * (got[i * N + r] - s);
      nrm += s * s;
    }
  printf("forward relerr %.3e\n", sqrt(err / nrm));
  if (sqrt(err / nrm) > 1e-12) return 1;
  memcpy(y, want, sizeof(y));
  CHECK(p_djets_vec_scatter(vy, y));
  CHECK(p_djets_op_mul_adjoint(A, vx2, vy));
  double lhs = 0, rhs = 0;                                       /* <A x, y> == <x, A' y> */
  CHECK(p_djets_vec_dot(vy, vy, &lhs));
  CHECK(p_djets_vec_dot(vx, vx2, &rhs));
  printf("dot test %.3e\n", fabs(lhs - rhs) / fabs(lhs));
  if (fabs(lhs - rhs) > 1e-12 * fabs(lhs)) return 1;
  CHECK(p_djets_vec_destroy(vx)); CHECK(p_djets_vec_destroy(vy)); CHECK(p_djets_vec_destroy(vx2));
  CHECK(p_djets_op_destroy(A));
  CHECK(p_djets_ctx_destroy(ctx));
  printf("c abi ok\n");
  return 0;
}
