// Standalone device check of the pack kernel + tcgen05 pair engine against an
// fp64 host computation.  Development tool (run under gpurun), not shipped API.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>
#include <cuda_runtime.h>
#include "../../../include/amb200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)
#define AK(x) do { int r_ = (x); if (r_ != 0) { printf("amb error %d: %s at %s:%d\n", r_, amb_last_error(), __FILE__, __LINE__); return 1; } } while (0)

static int run_case(long long na, long long nb, int d, unsigned lbo, unsigned sbo, int positive, unsigned seed) {
  std::mt19937 rng(seed);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::vector<float> A(na * d), B(nb * d);
  auto fill = [&](std::vector<float>& M, long long n) {
    for (long long i = 0; i < n; ++i) {
      double nrm = 0;
      for (int k = 0; k < d; ++k) { float v = nd(rng); if (positive) v = std::fabs(v) + 0.1f; M[i * d + k] = v; nrm += double(v) * v; }
      float inv = float(1.0 / std::sqrt(nrm));
      for (int k = 0; k < d; ++k) M[i * d + k] *= inv;
    }
  };
  fill(A, na); fill(B, nb);
  float *dA, *dB, *dC; void *pA, *pB;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dC, na * nb * 4));
  CK(cudaMalloc(&pA, amb_packed_bytes(na, d))); CK(cudaMalloc(&pB, amb_packed_bytes(nb, d)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, na * nb * 4));
  AK(amb_pack(0, nullptr, dA, AMB_F32, na, d, d, pA));
  AK(amb_pack(0, nullptr, dB, AMB_F32, nb, d, d, pB));
  AK(amb_debug_dot_matrix(0, nullptr, pA, na, pB, nb, d, dC, nb, lbo, sbo));
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case na=%lld nb=%lld d=%d lbo=%u sbo=%u: KERNEL FAILED: %s\n", na, nb, d, lbo, sbo, cudaGetErrorString(e)); return 2; }
  std::vector<float> C(na * nb);
  CK(cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost));
  double max_abs = 0, sum_signed = 0, sum_sq = 0; long long cnt = 0, nan_cnt = 0;
  for (long long i = 0; i < na; ++i)
    for (long long j = 0; j < nb; ++j) {
      double ref = 0;
      for (int k = 0; k < d; ++k) ref += double(A[i * d + k]) * double(B[j * d + k]);
      double got = C[i * nb + j];
      if (!(got == got)) { nan_cnt++; continue; }
      double err = got - ref;
      max_abs = std::fmax(max_abs, std::fabs(err));
      sum_signed += err; sum_sq += err * err; cnt++;
    }
  printf("case na=%lld nb=%lld d=%d lbo=%u sbo=%u pos=%d: max_abs_err=%.3e mean_signed=%.3e rms=%.3e nan=%lld  %s\n",
         na, nb, d, lbo, sbo, positive, max_abs, cnt ? sum_signed / cnt : 0.0, cnt ? std::sqrt(sum_sq / cnt) : 0.0, nan_cnt,
         (max_abs < 5e-6 && nan_cnt == 0) ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(pA); cudaFree(pB);
  return 0;
}

static int time_case(long long n, int d) {
  std::vector<float> A(n * d);
  std::mt19937 rng(1);
  std::normal_distribution<float> nd(0.f, 1.f);
  for (auto& v : A) v = nd(rng) * 0.044f;
  float *dA, *dC; void* pA;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dC, n * 4)); CK(cudaMalloc(&pA, amb_packed_bytes(n, d)));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1, e2; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2);
  AK(amb_pack(0, nullptr, dA, AMB_F32, n, d, d, pA));
  AK(amb_debug_dot_matrix(0, nullptr, pA, n, pA, n, d, dC, 0, 0, 0));
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  AK(amb_pack(0, nullptr, dA, AMB_F32, n, d, d, pA));
  cudaEventRecord(e1);
  AK(amb_debug_dot_matrix(0, nullptr, pA, n, pA, n, d, dC, 0, 0, 0));
  cudaEventRecord(e2);
  CK(cudaDeviceSynchronize());
  float t_pack, t_eng; cudaEventElapsedTime(&t_pack, e0, e1); cudaEventElapsedTime(&t_eng, e1, e2);
  double pairs = double(n) * n;
  printf("time n=%lld d=%d: pack %.3f ms (%.1f GB/s in+out), engine %.3f ms -> %.3e pairs/s, %.1f TFLOP/s algorithmic (x3 executed = %.1f)\n",
         n, d, t_pack, (double(n) * d * 8) / t_pack * 1e-6, t_eng, pairs / (t_eng * 1e-3), pairs * 2 * d / (t_eng * 1e-3) * 1e-12,
         pairs * 6 * d / (t_eng * 1e-3) * 1e-12);
  cudaFree(dA); cudaFree(dC); cudaFree(pA);
  return 0;
}

int main(int argc, char** argv) {
  int mode = argc > 1 ? atoi(argv[1]) : 0;
  printf("amb version %d\n", amb_version());
  if (mode == 0) {
    // descriptor semantics: library default vs swapped fields
    if (run_case(128, 256, 32, 128, 512, 0, 1) == 2) return 3;
    if (run_case(128, 256, 32, 512, 128, 0, 1) == 2) return 3;
    if (run_case(128, 256, 512, 128, 512, 0, 2) == 2) return 3;
    if (run_case(300, 700, 512, 128, 512, 0, 3) == 2) return 3;
    if (run_case(300, 700, 128, 128, 512, 0, 4) == 2) return 3;
    if (run_case(77, 1000, 10, 128, 512, 0, 5) == 2) return 3;
    if (run_case(1000, 1000, 512, 128, 512, 1, 6) == 2) return 3;   // positive data: accumulate rounding bias
  } else if (mode == 1) {
    time_case(16384, 512);
    time_case(32768, 512);
    time_case(32768, 128);
  }
  return 0;
}
