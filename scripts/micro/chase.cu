// Dependent-load latency on the box: one warp (one active lane) chasing pointers with a 12 KB-ish stride through buffers of several sizes,
// with and without a store to the line just before it is read (the rollout chain reads records another phase has just written).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chase(unsigned* buf, int steps, int mode, long long* out) {
  unsigned idx = 0;
  long long t0 = clock64();
  for (int i = 0; i < steps; i++) {
    if (mode == 1) buf[idx + 8] = i;                 // store to the same line (different word) before the dependent load
    idx = buf[idx];
  }
  long long t1 = clock64();
  out[0] = t1 - t0; out[1] = idx;
}
int main() {
  for (size_t mb : {4, 64, 400}) {
    size_t n = mb * 1024 * 1024 / 4;
    unsigned* h = (unsigned*)malloc(n * 4);
    size_t stride = 3072 + 16;                        // words: ~12 KB apart, like consecutive games' records
    size_t cnt = n / stride;
    for (size_t i = 0; i < cnt; i++) h[i * stride] = (unsigned)(((i * 7919 + 1) % cnt) * stride);
    unsigned* d; long long* o; long long ho[2];
    cudaMalloc(&d, n * 4); cudaMalloc(&o, 16);
    cudaMemcpy(d, h, n * 4, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 2; mode++) {
      int steps = 4000;
      chase<<<1, 32>>>(d, steps, mode, o);            // warm
      chase<<<1, 32>>>(d, steps, mode, o);
      cudaMemcpy(ho, o, 16, cudaMemcpyDeviceToHost);
      printf("buffer %4zu MB  mode %d (%s): %.0f cycles per dependent load\n", mb, mode, mode ? "store to the line first" : "read only", (double)ho[0] / steps);
    }
    cudaFree(d); cudaFree(o); free(h);
  }
  return 0;
}
