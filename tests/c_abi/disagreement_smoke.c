/* Plain-C consumer of libec_b200.so: proves the boundary is a C ABI (include/ec_b200.h compiles as C, no torch, no
 * C++ types).  Reads N, D and N*D floats + N int64 keys from stdin, prints "key count score" per instance.
 * Built and run by tests/test_c_abi.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <cuda_runtime_api.h>
#include "ec_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "cuda: %s\n", cudaGetErrorString(e_)); return 2; } } while (0)

int main(void) {
  int n, d, mode;
  if (scanf("%d %d %d", &n, &d, &mode) != 3) return 1;
  float* emb = (float*)malloc(sizeof(float) * (size_t)n * d);
  int64_t* keys = (int64_t*)malloc(sizeof(int64_t) * (size_t)n);
  for (long i = 0; i < (long)n * d; ++i) if (scanf("%f", &emb[i]) != 1) return 1;
  for (int i = 0; i < n; ++i) { long long k; if (scanf("%lld", &k) != 1) return 1; keys[i] = (int64_t)k; }
  if (ec_abi_version() != EC_ABI_VERSION) { fprintf(stderr, "ABI mismatch\n"); return 3; }
  float *d_emb, *d_score; int64_t *d_keys, *d_okeys; int32_t *d_count, *d_m; void* d_ws;
  size_t ws_bytes = ec_disagreement_workspace_bytes(n);
  CK(cudaMalloc((void**)&d_emb, sizeof(float) * (size_t)n * d)); CK(cudaMalloc((void**)&d_keys, sizeof(int64_t) * n));
  CK(cudaMalloc((void**)&d_okeys, sizeof(int64_t) * n)); CK(cudaMalloc((void**)&d_score, sizeof(float) * n));
  CK(cudaMalloc((void**)&d_count, sizeof(int32_t) * n)); CK(cudaMalloc((void**)&d_m, sizeof(int32_t)));
  CK(cudaMalloc(&d_ws, ws_bytes));
  CK(cudaMemcpy(d_emb, emb, sizeof(float) * (size_t)n * d, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_keys, keys, sizeof(int64_t) * n, cudaMemcpyHostToDevice));
  int rc = ec_instance_disagreement(d_emb, d, d_keys, n, d, mode, d_okeys, d_score, d_count, d_m, d_ws, ws_bytes, NULL);
  if (rc != EC_OK) { fprintf(stderr, "ec_instance_disagreement: %d %s\n", rc, ec_last_error()); return 4; }
  CK(cudaDeviceSynchronize());
  int32_t m = 0;
  CK(cudaMemcpy(&m, d_m, sizeof(int32_t), cudaMemcpyDeviceToHost));
  int64_t* okeys = (int64_t*)malloc(sizeof(int64_t) * n); float* score = (float*)malloc(sizeof(float) * n);
  int32_t* count = (int32_t*)malloc(sizeof(int32_t) * n);
  CK(cudaMemcpy(okeys, d_okeys, sizeof(int64_t) * m, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(score, d_score, sizeof(float) * m, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(count, d_count, sizeof(int32_t) * m, cudaMemcpyDeviceToHost));
  for (int i = 0; i < m; ++i) printf("%lld %d %.9g\n", (long long)okeys[i], (int)count[i], (double)score[i]);
  /* error convention: negative status + message, no exception, no abort */
  rc = ec_instance_disagreement(d_emb, d, d_keys, n, 3 /* not a multiple of 4 */, mode, d_okeys, d_score, d_count, d_m,
                                d_ws, ws_bytes, NULL);
  if (rc >= 0 || ec_last_error()[0] == 0) { fprintf(stderr, "expected an error status\n"); return 5; }
  return 0;
}
