// FP64 peak calibration for B200 (sm_100a): DMMA.8x8x4 issue rate vs DFMA issue rate.
// MEASURED_PEAKS.json holds only HBM and bf16 numbers; the FP64 roofline denominator used by
// bench.py comes from this microbenchmark (register-resident, no memory traffic).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void __launch_bounds__(1024) dmma_loop(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[CHAINS][2];
#pragma unroll
  for (int j = 0; j < CHAINS; j++) { c[j][0] = 0.0; c[j][1] = 0.0; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < CHAINS; j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; j++) s += c[j][0] + c[j][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CHAINS>
__global__ void __launch_bounds__(1024) dfma_loop(double* out, const double* in, int iters) {
  double a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
  double c[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; j++) c[j] = j;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int j = 0; j < CHAINS; j++) c[j] = fma(a, c[j], b);
  }
  double s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; j++) s += c[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_kernel(void (*launch)(int, int, double*, const double*, int), int blocks, int threads,
                         double* out, const double* in, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(blocks, threads, out, in, iters / 8);  // warm
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < 5; r++) {
    cudaEventRecord(e0); launch(blocks, threads, out, in, iters); cudaEventRecord(e1);
    cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}
static void l_dmma8(int b, int t, double* o, const double* i, int it) { dmma_loop<8><<<b, t>>>(o, i, it); }
static void l_dmma16(int b, int t, double* o, const double* i, int it) { dmma_loop<16><<<b, t>>>(o, i, it); }
static void l_dfma8(int b, int t, double* o, const double* i, int it) { dfma_loop<8><<<b, t>>>(o, i, it); }
static void l_dfma16(int b, int t, double* o, const double* i, int it) { dfma_loop<16><<<b, t>>>(o, i, it); }

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  double *in, *out; cudaMalloc(&in, 64 * 8); cudaMalloc(&out, (size_t)sms * 4 * 1024 * 8);
  double h[64]; for (int i = 0; i < 64; i++) h[i] = 1e-3 * (i + 1);
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"results\": [\n", p.name, sms, p.clockRate);
  int iters = 20000; bool first = true;
  int warps_list[] = {4, 8, 16, 32};
  for (int wi = 0; wi < 4; wi++) {
    int threads = warps_list[wi] * 32;
    struct { const char* name; void (*f)(int, int, double*, const double*, int); int chains; int flop_per_thread_per_op; } ks[] = {
        {"dmma_8x8x4_c8", l_dmma8, 8, 16}, {"dmma_8x8x4_c16", l_dmma16, 16, 16},
        {"dfma_c8", l_dfma8, 8, 2}, {"dfma_c16", l_dfma16, 16, 2}};
    for (auto& k : ks) {
      float ms = time_kernel(k.f, sms, threads, out, in, iters);
      double flop = (double)sms * threads * (double)iters * k.chains * k.flop_per_thread_per_op;
      printf("%s  {\"kernel\": \"%s\", \"warps_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.3f}", first ? "" : ",\n", k.name,
             warps_list[wi], ms, flop / ms * 1e-9);
      first = false;
    }
  }
  printf("\n]}\n");
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
