// Phase timing of potf2_kernel / trsm_panel_kernel with clock64 marks (debug harness, not part of the library).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DIPM_PHASE_TIMING -o tools/potf2_phases tools/potf2_phases.cu
#include <cstdio>
#include <vector>
#include <cmath>
#include "../interiorpoint-gpu_b200/csrc/chol.cu"

extern "C" int ipm_set_cuda_error(cudaError_t) { return -2; }
extern "C" void ipm_count_launch(void) {}
extern "C" int ipm_gemm_tn_f64(const double*, int, const double*, int, const double*, double, double, double*, int, int,
                               int, int, int, void*) { return 0; }

int main() {
  const int n = 128, ld = 128;
  std::vector<double> h(n * ld);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) h[i * ld + j] = (i == j ? n : 0.0) + 1.0 / (1.0 + abs(i - j));
  double* d; int* info;
  cudaMalloc(&d, sizeof(double) * n * ld); cudaMalloc(&info, 4); cudaMemset(info, 0, 4);
  for (int rep = 0; rep < 3; ++rep) {
    cudaMemcpy(d, h.data(), sizeof(double) * n * ld, cudaMemcpyHostToDevice);
    launch_potf2(d, ld, n, 0, info, 0);
    cudaDeviceSynchronize();
  }
  long long t[64];
  cudaMemcpyFromSymbol(t, g_phase_t, sizeof(t));
  printf("load %lld\n", t[1] - t[0]);
  for (int kb = 0; kb < 4; ++kb)
    printf("kb%d diag %lld  trsm %lld  update %lld\n", kb, t[3 + 4 * kb] - t[2 + 4 * kb], kb < 3 ? t[4 + 4 * kb] - t[3 + 4 * kb] : 0,
           kb < 3 ? t[5 + 4 * kb] - t[4 + 4 * kb] : 0);
  printf("store %lld  total %lld cycles\n", t[21] - t[20], t[21] - t[0]);
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(e));
  return 0;
}
