// Host-side cost of CUDA API calls on this platform: kernel launches, event records, and the same
// work submitted as one CUDA graph.  nvcc -O2 -gencode arch=compute_100a,code=sm_100a launch_rate.cu -o launch_rate
#include <chrono>
#include <cstdio>
#include <cuda_runtime.h>
__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 9999) *p = 1; }
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  cudaEvent_t ev[64]; for (auto& e : ev) cudaEventCreate(&e);
  const int N = 2000;
  for (int i = 0; i < 100; i++) empty_kernel<<<1, 32, 0, st>>>(nullptr);
  cudaStreamSynchronize(st);
  double t0 = now_us();
  for (int i = 0; i < N; i++) empty_kernel<<<1, 32, 0, st>>>(nullptr);
  double t1 = now_us(); cudaStreamSynchronize(st); double t2 = now_us();
  printf("launch only: %.2f us per launch to enqueue, %.2f us per launch incl. drain\n", (t1 - t0) / N, (t2 - t0) / N);
  t0 = now_us();
  for (int i = 0; i < N; i++) { empty_kernel<<<1, 32, 0, st>>>(nullptr); cudaEventRecord(ev[i & 63], st); }
  t1 = now_us(); cudaStreamSynchronize(st); t2 = now_us();
  printf("launch + event record: %.2f us per pair to enqueue, %.2f incl. drain\n", (t1 - t0) / N, (t2 - t0) / N);
  t0 = now_us();
  for (int i = 0; i < N; i++) empty_kernel<<<5000, 256, 0, st>>>(nullptr);
  t1 = now_us(); cudaStreamSynchronize(st); t2 = now_us();
  printf("5000-CTA empty kernels: %.2f us per launch to enqueue, %.2f incl. drain (GPU-side cost of scheduling the CTAs)\n", (t1 - t0) / N, (t2 - t0) / N);
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < 200; i++) empty_kernel<<<1, 32, 0, st>>>(nullptr);
  cudaStreamEndCapture(st, &g); cudaGraphInstantiate(&ge, g, 0);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  t0 = now_us();
  for (int i = 0; i < 10; i++) cudaGraphLaunch(ge, st);
  t1 = now_us(); cudaStreamSynchronize(st); t2 = now_us();
  printf("graph of 200 kernels: %.2f us per kernel to enqueue, %.2f incl. drain\n", (t1 - t0) / 2000, (t2 - t0) / 2000);
  return 0;
}
