/* Plain-C driver for the libpsagpu C ABI (include/psa_gpu.h): what a non-Python host (e.g. Julia's ccall) does.
 * Usage: abi_driver <n> <lx> <ly> [full]  -- builds a deterministic upper-triangular complex matrix, runs one grid and
 * prints "Z j k value iters" lines.  With `full` it then walks the rest of the boundary the way the Julia shim does:
 * psa_gpu_set_thresholds, psa_gpu_grid_batched (two row batches, result kept on the device), psa_gpu_postprocess
 * (mirror of 2 rows + clamp + statistics) and psa_gpu_pseudomode (n >= 200), printing "B", "P", "Y", "S", "M", "V"
 * lines.  Exit code: 0 ok, 10 no device (PSA_ERR_NODEVICE), 1 other failure. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "psa_gpu.h"

int main(int argc, char** argv) {
  int n = argc > 1 ? atoi(argv[1]) : 64, lx = argc > 2 ? atoi(argv[2]) : 3, ly = argc > 3 ? atoi(argv[3]) : 2;
  void* ctx = NULL;
  int st = psa_gpu_init(1, NULL, &ctx);
  if (st == PSA_ERR_NODEVICE) {
    printf("no device: %s\n", psa_gpu_strerror(st));
    return 10;
  }
  if (st != PSA_OK) return 1;
  double* T = (double*)calloc((size_t)2 * n * n, sizeof(double));
  for (int j = 0; j < n; ++j)
    for (int i = 0; i <= j; ++i) { /* column-major, interleaved complex */
      T[2 * ((size_t)j * n + i)] = (i == j) ? 0.05 * i - 1.0 : 0.3 * sin(1.0 + i + 2.0 * j);
      T[2 * ((size_t)j * n + i) + 1] = (i == j) ? 0.02 * i : 0.3 * cos(2.0 + 3.0 * i + j);
    }
  double* x = (double*)malloc(sizeof(double) * lx);
  double* y = (double*)malloc(sizeof(double) * ly);
  for (int k = 0; k < lx; ++k) x[k] = -2.0 + 4.0 * k / (lx > 1 ? lx - 1 : 1);
  for (int j = 0; j < ly; ++j) y[j] = -1.0 + 3.0 * j / (ly > 1 ? ly - 1 : 1);
  double* q0 = (double*)malloc(sizeof(double) * 2 * n);
  double nrm = 0;
  for (int i = 0; i < n; ++i) { q0[2 * i] = cos(0.7 * i); q0[2 * i + 1] = sin(1.3 * i + 0.2); nrm += q0[2 * i] * q0[2 * i] + q0[2 * i + 1] * q0[2 * i + 1]; }
  nrm = sqrt(nrm);
  for (int i = 0; i < 2 * n; ++i) q0[i] /= nrm;
  double* Z = (double*)malloc(sizeof(double) * lx * ly);
  int32_t* it = (int32_t*)malloc(sizeof(int32_t) * lx * ly);
  int64_t counts[PSA_NCOUNTS];
  int algo = 0;
  st = psa_gpu_set_matrix(ctx, T, n, n, n);
  if (st != PSA_OK) { printf("set_matrix: %d %s | %s\n", st, psa_gpu_strerror(st), psa_gpu_last_error(ctx)); return 1; }
  st = psa_gpu_grid(ctx, x, lx, y, ly, q0, 1e-5, 99, 0.1 * 1.7976931348623157e308, Z, it, counts, &algo, NULL);
  if (st != PSA_OK) { printf("grid: %d %s | %s\n", st, psa_gpu_strerror(st), psa_gpu_last_error(ctx)); return 1; }
  printf("algo %d points %lld iters %lld\n", algo, (long long)counts[PSA_COUNT_POINTS], (long long)counts[PSA_COUNT_TOTAL_ITERS]);
  for (int k = 0; k < lx; ++k)
    for (int j = 0; j < ly; ++j) printf("Z %d %d %.17g %d\n", j, k, Z[(size_t)k * ly + j], it[(size_t)k * ly + j]);
  if (argc > 4) {
    /* thresholds as the reference's defaults (compute.jl:27), set explicitly; re-upload so that they take effect */
    st = psa_gpu_set_thresholds(ctx, 55, 200);
    if (st == PSA_OK) st = psa_gpu_set_matrix_ex(ctx, T, n, n, n, PSA_MATRIX_DEFAULT);
    if (st != PSA_OK) { printf("thresholds/set_matrix_ex: %d\n", st); return 1; }
    /* two start vectors: batch 0 = q0, batch 1 = conj(q0) reversed; rows alternate between them */
    double* qs = (double*)malloc(sizeof(double) * 4 * n);
    for (int i = 0; i < n; ++i) {
      qs[2 * i] = q0[2 * i]; qs[2 * i + 1] = q0[2 * i + 1];
      qs[2 * (n + i)] = q0[2 * (n - 1 - i)]; qs[2 * (n + i) + 1] = -q0[2 * (n - 1 - i) + 1];
    }
    int32_t* rb = (int32_t*)malloc(sizeof(int32_t) * ly);
    for (int j = 0; j < ly; ++j) rb[j] = j & 1;
    st = psa_gpu_grid_batched(ctx, x, lx, y, ly, qs, 2, rb, 1e-5, 150, 0.1 * 1.7976931348623157e308, NULL, it, counts,
                              &algo, NULL);
    if (st != PSA_OK) { printf("grid_batched: %d %s | %s\n", st, psa_gpu_strerror(st), psa_gpu_last_error(ctx)); return 1; }
    int nm = ly >= 2 ? 2 : 0, lyf = ly + nm;
    double* Zf = (double*)malloc(sizeof(double) * lx * lyf);
    double* yf = (double*)malloc(sizeof(double) * lyf);
    double zs[3];
    st = psa_gpu_postprocess(ctx, y, ly, lx, nm, 1.0e-2 /* visible clamp */, 1e-150, Zf, lyf, yf, zs);
    if (st != PSA_OK) { printf("postprocess: %d %s | %s\n", st, psa_gpu_strerror(st), psa_gpu_last_error(ctx)); return 1; }
    for (int k = 0; k < lx; ++k)
      for (int j = 0; j < ly; ++j) printf("B %d %d %d\n", j, k, it[(size_t)k * ly + j]);
    for (int k = 0; k < lx; ++k)
      for (int r = 0; r < lyf; ++r) printf("P %d %d %.17g\n", r, k, Zf[(size_t)k * lyf + r]);
    for (int r = 0; r < lyf; ++r) printf("Y %d %.17g\n", r, yf[r]);
    printf("S %.17g %.17g %.17g\n", zs[0], zs[1], zs[2]);
    if (n >= 200) {
      double sig = 0;
      double* mode = (double*)malloc(sizeof(double) * 2 * n);
      int its = 0, info = -1;
      st = psa_gpu_pseudomode(ctx, 0.3, 0.4, q0, 1e-10, 60, &sig, mode, &its, &info);
      if (st != PSA_OK) { printf("pseudomode: %d %s | %s\n", st, psa_gpu_strerror(st), psa_gpu_last_error(ctx)); return 1; }
      printf("M %.17g %d %d\n", sig, its, info);
      for (int i = 0; i < n; ++i) printf("V %d %.17g %.17g\n", i, mode[2 * i], mode[2 * i + 1]);
    }
  }
  psa_gpu_free(ctx);
  return 0;
}
