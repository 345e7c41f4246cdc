// Host-buffer entry points (amb_host_*): host arrays in, host results out.  They
// stage through device memory this file allocates and frees, call the same
// device-pointer entry points as everyone else on a private stream, and
// synchronise before returning.  This is the surface a numpy / ctypes binding of
// the reference's metric functions calls directly.
#include <vector>

#include "internal.cuh"

namespace amb {

struct HostCall {
  int dev;
  cudaStream_t st = nullptr;
  std::vector<void*> bufs;
  int rc = AMB_OK;
  DeviceGuard guard;
  explicit HostCall(int d) : dev(d), guard(d) {
    if (!guard.ok) { rc = AMB_ERR_CUDA; return; }
    rc = check_cuda(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), "cudaStreamCreate");
  }
  ~HostCall() {
    for (void* p : bufs) cudaFree(p);
    if (st) cudaStreamDestroy(st);
  }
  template <typename T>
  T* alloc(size_t count) {
    if (rc) return nullptr;
    void* p = nullptr;
    rc = check_cuda(cudaMalloc(&p, (count ? count : 1) * sizeof(T)), "cudaMalloc");
    if (rc) return nullptr;
    bufs.push_back(p);
    return static_cast<T*>(p);
  }
  template <typename T>
  T* upload(const T* host, size_t count) {
    T* p = alloc<T>(count);
    if (!rc && count) rc = check_cuda(cudaMemcpyAsync(p, host, count * sizeof(T), cudaMemcpyHostToDevice, st), "H2D");
    return p;
  }
  void* upload_raw(const void* host, size_t bytes) {
    uint8_t* p = alloc<uint8_t>(bytes);
    if (!rc && bytes) rc = check_cuda(cudaMemcpyAsync(p, host, bytes, cudaMemcpyHostToDevice, st), "H2D");
    return p;
  }
  template <typename T>
  void download(T* host, const T* dev_ptr, size_t count) {
    if (!rc && count) rc = check_cuda(cudaMemcpyAsync(host, dev_ptr, count * sizeof(T), cudaMemcpyDeviceToHost, st), "D2H");
  }
  void zero(void* p, size_t bytes) {
    if (!rc) rc = check_cuda(cudaMemsetAsync(p, 0, bytes, st), "memset");
  }
  int finish() {
    if (!rc) rc = check_cuda(cudaStreamSynchronize(st), "cudaStreamSynchronize");
    return rc;
  }
  void run(int r) { if (!rc) rc = r; }
};

static size_t esize(int dtype) { return dtype == AMB_F64 ? 8 : 4; }

}  // namespace amb

using namespace amb;

extern "C" {

int amb_host_stats(int dev, const void* X, int dtype, long long n, int d, double* mean, double* cov) {
  if (!X || !mean || !cov || n <= 0 || d <= 0 || (dtype != AMB_F32 && dtype != AMB_F64))
    return set_error(AMB_ERR_ARG, "amb_host_stats: bad argument");
  HostCall h(dev);
  const size_t dd = static_cast<size_t>(d) * d;
  void* dX = h.upload_raw(X, static_cast<size_t>(n) * d * esize(dtype));
  double* sum = h.alloc<double>(d);
  double* gram = h.alloc<double>(dd);
  double* dmean = h.alloc<double>(d);
  double* dcov = h.alloc<double>(dd);
  const size_t wsb = amb_cov_ws_bytes(n, d);
  void* ws = h.alloc<uint8_t>(wsb);
  if (h.rc) return h.rc;
  h.zero(sum, d * 8);
  h.zero(gram, dd * 8);
  h.run(amb_cov_accumulate(dev, h.st, dX, dtype, n, d, d, sum, gram, ws, wsb));
  h.run(amb_cov_finalize(dev, h.st, n, d, sum, gram, dmean, dcov));
  h.download(mean, dmean, d);
  h.download(cov, dcov, dd);
  return h.finish();
}

int amb_host_frechet(int dev, int d, const double* mu_x, const double* cov_x, const double* mu_y,
                     const double* cov_y, double* out) {
  if (!mu_x || !cov_x || !mu_y || !cov_y || !out || d <= 0) return set_error(AMB_ERR_ARG, "amb_host_frechet: bad argument");
  HostCall h(dev);
  const size_t dd = static_cast<size_t>(d) * d;
  double* mx = h.upload(mu_x, d);
  double* cx = h.upload(cov_x, dd);
  double* my = h.upload(mu_y, d);
  double* cy = h.upload(cov_y, dd);
  double* o = h.alloc<double>(1);
  const size_t wsb = amb_frechet_ws_bytes(1, d);
  void* ws = h.alloc<uint8_t>(wsb);
  if (h.rc) return h.rc;
  h.run(amb_frechet(dev, h.st, 1, d, mx, cx, my, cy, o, ws, wsb));
  h.download(out, o, 1);
  return h.finish();
}

int amb_host_kd(int dev, const void* F1, long long n1, const void* F2, long long n2, int d,
                int dtype, const int32_t* idx, int S, int m, double gamma, double coef0, int degree,
                double* mmd2_out, double* stats_out) {
  if (!F1 || !F2 || !idx || !stats_out || n1 <= 0 || n2 <= 0 || d <= 0 || S <= 0 || m <= 0 ||
      (dtype != AMB_F32 && dtype != AMB_F64))
    return set_error(AMB_ERR_ARG, "amb_host_kd: bad argument");
  HostCall h(dev);
  void* d1 = h.upload_raw(F1, static_cast<size_t>(n1) * d * esize(dtype));
  void* d2 = h.upload_raw(F2, static_cast<size_t>(n2) * d * esize(dtype));
  int32_t* didx = h.upload(idx, static_cast<size_t>(S) * 2 * m);
  double* mm = h.alloc<double>(S);
  double* stt = h.alloc<double>(2);
  const size_t wsb = amb_kd_ws_bytes(S, m, d);
  void* ws = h.alloc<uint8_t>(wsb);
  if (h.rc) return h.rc;
  h.run(amb_kd_subsets(dev, h.st, d1, n1, d, d2, n2, d, d, dtype, didx, S, m, AMB_KERNEL_POLY, gamma, coef0,
                       degree, 1.0, AMB_MMD_UNBIASED, mm, stt, ws, wsb));
  if (mmd2_out) h.download(mmd2_out, mm, S);
  h.download(stats_out, stt, 2);
  return h.finish();
}

int amb_host_knn_radii(int dev, const void* X, int dtype, long long n, int d, int k, float* radii) {
  if (!X || !radii || n <= 0 || d <= 0 || (dtype != AMB_F32 && dtype != AMB_F64))
    return set_error(AMB_ERR_ARG, "amb_host_knn_radii: bad argument");
  HostCall h(dev);
  void* dX = h.upload_raw(X, static_cast<size_t>(n) * d * esize(dtype));
  void* packed = h.alloc<uint8_t>(amb_packed_bytes(n, d));
  float* r = h.alloc<float>(n);
  const size_t wsb = amb_knn_ws_bytes(n, n, d, k);
  void* ws = h.alloc<uint8_t>(wsb);
  if (h.rc) return h.rc;
  h.run(amb_pack(dev, h.st, dX, dtype, n, d, d, packed));
  h.run(amb_knn_radii(dev, h.st, dX, dtype, d, packed, n, d, 0, n, k, r, nullptr, ws, wsb));
  h.download(radii, r, n);
  return h.finish();
}

int amb_host_prdc(int dev, const void* ref, long long n, const void* cand, long long m, int d,
                  int dtype, int k, double* out) {
  if (!ref || !cand || !out || n <= 0 || m <= 0 || d <= 0 || (dtype != AMB_F32 && dtype != AMB_F64))
    return set_error(AMB_ERR_ARG, "amb_host_prdc: bad argument");
  HostCall h(dev);
  void* dR = h.upload_raw(ref, static_cast<size_t>(n) * d * esize(dtype));
  void* dC = h.upload_raw(cand, static_cast<size_t>(m) * d * esize(dtype));
  void* pR = h.alloc<uint8_t>(amb_packed_bytes(n, d));
  void* pC = h.alloc<uint8_t>(amb_packed_bytes(m, d));
  float* rR = h.alloc<float>(n);
  float* rC = h.alloc<float>(m);
  int32_t* col = h.alloc<int32_t>(m);
  uint8_t* rec = h.alloc<uint8_t>(n);
  uint8_t* cov = h.alloc<uint8_t>(n);
  long long* totals = h.alloc<long long>(8);
  size_t wsb = amb_knn_ws_bytes(n, n, d, k);
  const size_t w2 = amb_knn_ws_bytes(m, m, d, k), w3 = amb_prdc_ws_bytes(n, m);
  wsb = wsb > w2 ? wsb : w2;
  wsb = wsb > w3 ? wsb : w3;
  void* ws = h.alloc<uint8_t>(wsb);
  if (h.rc) return h.rc;
  h.run(amb_pack(dev, h.st, dR, dtype, n, d, d, pR));
  h.run(amb_pack(dev, h.st, dC, dtype, m, d, d, pC));
  h.run(amb_knn_radii(dev, h.st, dR, dtype, d, pR, n, d, 0, n, k, rR, nullptr, ws, wsb));
  h.run(amb_knn_radii(dev, h.st, dC, dtype, d, pC, m, d, 0, m, k, rC, nullptr, ws, wsb));
  // Counts, with the overflow ladder of amb200.h: default refine list -> a list sized to the number
  // of near-tie pairs the first attempt reported -> exhaustive exact counts.
  long long t[8] = {0};
  void* big = nullptr;
  for (int attempt = 0; attempt < 3; ++attempt) {
    h.zero(col, static_cast<size_t>(m) * 4);
    h.zero(totals, 64);
    if (attempt == 0) {
      h.run(amb_prdc_counts(dev, h.st, dR, d, pR, n, rR, dC, d, pC, m, rC, d, dtype, 0, n, col, rec, cov, totals + 4, ws, wsb));
    } else if (attempt == 1 && big) {
      const size_t bb = amb_prdc_ws_bytes_cap(n, m, t[4]);
      h.run(amb_prdc_counts(dev, h.st, dR, d, pR, n, rR, dC, d, pC, m, rC, d, dtype, 0, n, col, rec, cov, totals + 4, big, bb));
    } else {
      h.run(amb_prdc_counts_exact(dev, h.st, dR, d, n, rR, dC, d, m, rC, d, dtype, 0, n, col, rec, cov));
    }
    h.run(amb_prdc_reduce(dev, h.st, col, m, rec, cov, n, totals));
    h.download(t, totals, 8);
    int rc = h.finish();
    if (rc) return rc;
    const long long cap = attempt == 0 ? amb_prdc_ws_list_cap(n, m, wsb) : t[4];
    if (attempt == 2 || t[4] <= cap) break;
    if (attempt == 0) {   // a list for exactly the reported number of pairs, if the device can hold it
      size_t free_b = 0, total_b = 0;
      const size_t bb = amb_prdc_ws_bytes_cap(n, m, t[4]);
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && bb < free_b / 2 && cudaMalloc(&big, bb) == cudaSuccess)
        h.bufs.push_back(big);
      else
        big = nullptr;
      (void)cudaGetLastError();
    }
  }
  out[0] = static_cast<double>(t[0]) / static_cast<double>(m);                         // precision  prdc.py:36-38
  out[1] = static_cast<double>(t[2]) / static_cast<double>(n);                         // recall     prdc.py:40-42
  out[2] = (1.0 / static_cast<double>(k)) * (static_cast<double>(t[1]) / static_cast<double>(m));  // density prdc.py:44-46
  out[3] = static_cast<double>(t[3]) / static_cast<double>(n);                         // coverage   prdc.py:48
  return AMB_OK;
}

}  // extern "C"
