// Micro-probe: which host-side CUDA calls of a second stream block while a persistent kernel spins on another stream?
// nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o concurrency_probe concurrency_probe.cu && ./concurrency_probe
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

__global__ void spin_kernel(volatile int *flag, long long limit) {
  extern __shared__ char smem[];
  long long n = 0;
  if (threadIdx.x == 0) while (*flag == 0 && ++n < limit) {}
  __syncthreads();
  if (threadIdx.x == 1) smem[0] = 1;
}
__global__ void set_kernel(int *flag) { *flag = 1; }
__global__ void small_kernel(double *p) { p[threadIdx.x] += 1.0; }

static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main() {
  int *flag; double *buf;
  cudaMalloc(&flag, 4); cudaMemset(flag, 0, 4); cudaMalloc(&buf, 1 << 24);
  cudaStream_t a, b;
  cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&b, cudaStreamNonBlocking);
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 64);
  small_kernel<<<1, 32, 0, b>>>(buf); set_kernel<<<1, 1, 0, b>>>(flag); cudaDeviceSynchronize(); cudaMemset(flag, 0, 4);   // load everything
  std::vector<double> host(1 << 20, 1.0);
  double t0 = now();
  spin_kernel<<<63, 256, 60 * 1024, a>>>(flag, 1LL << 26);      // gives up after ~20 s on its own
  auto step = [&](const char *what, cudaError_t e) { printf("%-58s %8.3f s  %s\n", what, now() - t0, cudaGetErrorString(e)); fflush(stdout); };
  std::thread th([&] {
    step("thread B: start", cudaSuccess);
    step("cudaMemcpyAsync pageable H2D (8 MB) on stream b", cudaMemcpyAsync(buf, host.data(), 8 << 20, cudaMemcpyHostToDevice, b));
    step("cudaStreamSynchronize(b)", cudaStreamSynchronize(b));
    small_kernel<<<64, 32, 0, b>>>(buf);
    step("small kernel launched on b", cudaGetLastError());
    step("cudaStreamSynchronize(b)", cudaStreamSynchronize(b));
    step("cudaMemcpyAsync pageable D2H on b", cudaMemcpyAsync(host.data(), buf, 8 << 20, cudaMemcpyDeviceToHost, b));
    step("cudaStreamSynchronize(b)", cudaStreamSynchronize(b));
    step("cudaMemsetAsync on b", cudaMemsetAsync(buf, 0, 1024, b));
    step("cudaStreamSynchronize(b)", cudaStreamSynchronize(b));
    int nb = 0;
    step("cudaOccupancyMaxActiveBlocksPerMultiprocessor(spin)", cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spin_kernel, 256, 60 * 1024));
    step("cudaFuncSetAttribute(spin, carveout) while it runs", cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 64));
    spin_kernel<<<56, 256, 60 * 1024, b>>>(flag, 1LL << 26);
    step("second spin kernel launched on b", cudaGetLastError());
    set_kernel<<<1, 1, 0, b>>>(flag);      // behind the second spinner in stream b: only runs if BOTH spinners are resident?  no -- it is
    step("set kernel queued behind it", cudaGetLastError());
  });
  th.join();
  // release from a third stream: if the two spinners run concurrently this ends everything at once
  cudaStream_t c; cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking);
  set_kernel<<<1, 1, 0, c>>>(flag);
  step("set kernel on third stream launched", cudaGetLastError());
  step("cudaStreamSynchronize(c)", cudaStreamSynchronize(c));
  step("cudaStreamSynchronize(a)", cudaStreamSynchronize(a));
  step("cudaStreamSynchronize(b)", cudaStreamSynchronize(b));
  return 0;
}
