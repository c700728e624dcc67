// tools/micro/h2d.cu -- host->device copy rate from pinned memory: one stream, two streams side by side, and with a D2H copy running
// nvcc -O2 -o /dev/shm/h2d tools/micro/h2d.cu && /dev/shm/h2d
#include <cuda_runtime.h>
#include <stdio.h>
#include <chrono>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
   const size_t N = (size_t)4 << 30, C = (size_t)36 << 20;
   char *h, *d, *h2, *d2; cudaHostAlloc(&h, N, 0); cudaMalloc(&d, N); cudaHostAlloc(&h2, N / 4, 0); cudaMalloc(&d2, N / 4);
   cudaStream_t s[4]; for (auto &x : s) cudaStreamCreateWithFlags(&x, cudaStreamNonBlocking);
   for (int rep = 0; rep < 2; ++rep) {
      for (int ns = 1; ns <= 4; ns *= 2) {
         cudaDeviceSynchronize(); double t0 = now();
         size_t i = 0; for (size_t at = 0; at < N; at += C, ++i) cudaMemcpyAsync(d + at, h + at, C < N - at ? C : N - at, cudaMemcpyHostToDevice, s[i % ns]);
         cudaDeviceSynchronize(); double t1 = now();
         printf("H2D %d stream(s), 36 MB chunks: %.1f GB/s\n", ns, N / (t1 - t0) / 1e9); }
      cudaDeviceSynchronize(); double t0 = now();
      for (size_t at = 0; at < N; at += C) cudaMemcpyAsync(d + at, h + at, C < N - at ? C : N - at, cudaMemcpyHostToDevice, s[0]);
      for (size_t at = 0; at < N / 4; at += C) cudaMemcpyAsync(h2 + at, d2 + at, C < N / 4 - at ? C : N / 4 - at, cudaMemcpyDeviceToHost, s[1]);
      cudaDeviceSynchronize(); double t1 = now();
      printf("H2D 4 GB with a 1 GB D2H beside it: %.1f GB/s (H2D bytes / total time)\n", N / (t1 - t0) / 1e9);
      cudaDeviceSynchronize(); t0 = now();
      cudaMemcpyAsync(d, h, N, cudaMemcpyHostToDevice, s[0]); cudaDeviceSynchronize(); t1 = now();
      printf("H2D one 4 GB copy: %.1f GB/s\n", N / (t1 - t0) / 1e9); }
   return 0; }
