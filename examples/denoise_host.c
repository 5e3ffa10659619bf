/* A complete C host for D(x; sigma) through libdiffsci_b200.so -- no Python in the process.
 *
 *   gcc -O2 examples/denoise_host.c -Iinclude -I/usr/local/cuda/include -Ldiffsci_b200 -ldiffsci_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/diffsci_b200 -o denoise_host
 *   ./denoise_host net.tape x.f32 sigma.f32 D.f32
 *
 * net.tape: written by diffsci_b200.tape.export_denoiser(module, B, shape, path) (the launch list of the network's plan, its
 * buffer sizes and its packed weights).  x.f32: fp32 [B, C, *S] as the reference lays it out (NC(D)HW), sigma.f32: fp32 [B];
 * D.f32 receives D(x; sigma) = c_skip x + c_out F(c_in x, c_noise) (reference karras/karrasmodule.py:673-719). */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "diffsci_b200.h"

#define CK(call)                                                                          \
  do {                                                                                    \
    if ((call) != 0) { fprintf(stderr, "%s failed: %s\n", #call, dsk_last_error()); return 1; } \
  } while (0)
#define CU(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 1; } \
  } while (0)

static int read_file(const char* path, void* dst, size_t n) {
  FILE* f = fopen(path, "rb");
  if (!f) return 0;
  const size_t got = fread(dst, 1, n, f);
  fclose(f);
  return got == n;
}

int main(int argc, char** argv) {
  if (argc != 5) { fprintf(stderr, "usage: %s net.tape x.f32 sigma.f32 out.f32\n", argv[0]); return 2; }
  CK(dsk_check_device(0));
  dsk_plan* plan = NULL;
  CK(dsk_plan_load(argv[1], &plan));
  const int64_t B = dsk_plan_info(plan, DSK_PLAN_BATCH), n = B * dsk_plan_info(plan, DSK_PLAN_SAMPLE_ELEMS);
  const int64_t ws_bytes = dsk_plan_info(plan, DSK_PLAN_WORKSPACE_BYTES);
  float *hx = (float*)malloc(n * 4), *hs = (float*)malloc(B * 4), *ho = (float*)malloc(n * 4);
  if (!read_file(argv[2], hx, n * 4) || !read_file(argv[3], hs, B * 4)) { fprintf(stderr, "cannot read the inputs\n"); return 1; }
  void* ws = NULL;
  float *dx = NULL, *ds = NULL, *dout = NULL;
  cudaStream_t st;
  CU(cudaStreamCreate(&st));
  CU(cudaMalloc(&ws, ws_bytes));                       /* cudaMalloc returns 256-byte aligned memory */
  CU(cudaMalloc((void**)&dx, n * 4));
  CU(cudaMalloc((void**)&ds, B * 4));
  CU(cudaMalloc((void**)&dout, n * 4));
  CK(dsk_plan_bind(plan, ws, st));
  CU(cudaMemcpyAsync(dx, hx, n * 4, cudaMemcpyHostToDevice, st));
  CU(cudaMemcpyAsync(ds, hs, B * 4, cudaMemcpyHostToDevice, st));
  CK(dsk_denoiser_fwd(plan, dx, ds, dout, st));
  CU(cudaMemcpyAsync(ho, dout, n * 4, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  FILE* f = fopen(argv[4], "wb");
  if (!f || fwrite(ho, 4, n, f) != (size_t)n) { fprintf(stderr, "cannot write %s\n", argv[4]); return 1; }
  fclose(f);
  printf("denoise_host: %lld launches of libdiffsci_b200 kernels, batch %lld, %lld values, workspace %.1f MB\n",
         (long long)dsk_plan_info(plan, DSK_PLAN_LAUNCHES), (long long)B, (long long)n, ws_bytes / 1e6);
  CK(dsk_plan_destroy(plan));
  return 0;
}
