// Host-buffer entry points (amb_host_*): host arrays in, host results out.  They
// stage through device memory this file allocates and frees, call the same
// device-pointer entry points as everyone else on a private stream, and
// synchronise before returning.  This is the surface a numpy / ctypes binding of
// the reference's metric functions calls directly.
#include <condition_variable>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
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

// Reusable barrier for the worker threads of amb_host_evaluate (one thread per device).
class ThreadBarrier {
 public:
  explicit ThreadBarrier(int n) : n_(n) {}
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    const int gen = gen_;
    if (++count_ == n_) {
      count_ = 0;
      ++gen_;
      cv_.notify_all();
    } else {
      cv_.wait(lk, [&] { return gen != gen_; });
    }
  }

 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int n_, count_ = 0, gen_ = 0;
};

// 256-aligned even split of n rows over `world` workers (the CTA-pair engine takes row tiles in pairs)
static void even_rows(long long n, int world, int rank, long long* row0, long long* nrows) {
  long long chunk = (n + world - 1) / world;
  chunk = (chunk + 255) / 256 * 256;
  long long a = chunk * rank, b = a + chunk;
  if (a > n) a = n;
  if (b > n) b = n;
  *row0 = a;
  *nrows = b - a;
}

struct EvalShared {
  // inputs
  const void* ref; const void* cand; long long n, m; int d, dtype, k;
  const int32_t* kd_idx; int S, msub; int want_fad; int n_dev;
  amb_comm_t* comm = nullptr;         // n_dev > 1: the exchanges between the devices (amb_comm_*)
  std::vector<long long> totals;      // [n_dev][4]: hits, sum of counts (device 0's are used), rows recalled, rows covered
  double fad = 0, kd[2] = {0, 0};
  std::vector<int> rc;
  std::vector<std::string> err;
};

// One device's share of amb_host_evaluate.
static void eval_worker(EvalShared* sh, ThreadBarrier* bar, int rank, int dev) {
  const long long n = sh->n, m = sh->m;
  const int d = sh->d, dtype = sh->dtype, k = sh->k;
  int& rc_out = sh->rc[rank];
  auto fail = [&](int rc) {
    rc_out = rc;
    sh->err[rank] = amb_last_error();
  };
  bool alive = true;
  {
    HostCall h(dev);
    const bool want_prdc = k > 0;
    const bool first = rank == 0;
    void* dR = nullptr; void* dC = nullptr; void* pR = nullptr; void* pC = nullptr;
    float* rR = nullptr; float* rC = nullptr;
    long long r0 = 0, rn = 0, c0 = 0, cn = 0;
    size_t wsb = 0;
    void* ws = nullptr;
    if (want_prdc || first) {
      dR = h.upload_raw(sh->ref, static_cast<size_t>(n) * d * esize(dtype));
      dC = h.upload_raw(sh->cand, static_cast<size_t>(m) * d * esize(dtype));
    }
    // ---- N-independent metrics on the first device, queued behind its uploads
    double* fad_dev = nullptr; double* kd_dev = nullptr;
    if (first && sh->want_fad && !h.rc) {
      const size_t dd = static_cast<size_t>(d) * d;
      double* mom = h.alloc<double>(2 * (d + dd));
      double* st = h.alloc<double>(2 * (d + dd));
      fad_dev = h.alloc<double>(1);
      const size_t cw = amb_cov_ws_bytes(n > m ? n : m, d), fw = amb_frechet_ws_bytes(1, d);
      void* w = h.alloc<uint8_t>(cw > fw ? cw : fw);
      if (!h.rc) {
        h.zero(mom, 2 * (d + dd) * 8);
        double* mr = mom; double* mc = mom + d + dd;
        h.run(amb_cov_accumulate(dev, h.st, dR, dtype, n, d, d, mr, mr + d, w, cw));
        h.run(amb_cov_accumulate(dev, h.st, dC, dtype, m, d, d, mc, mc + d, w, cw));
        double* sr = st; double* sc = st + d + dd;
        h.run(amb_cov_finalize(dev, h.st, n, d, mr, mr + d, sr, sr + d));
        h.run(amb_cov_finalize(dev, h.st, m, d, mc, mc + d, sc, sc + d));
        h.run(amb_frechet(dev, h.st, 1, d, sc, sc + d, sr, sr + d, fad_dev, w, fw));   // (cand, ref): audio_metrics.py:257
      }
    }
    if (first && sh->kd_idx && !h.rc) {
      int32_t* didx = h.upload(sh->kd_idx, static_cast<size_t>(sh->S) * 2 * sh->msub);
      kd_dev = h.alloc<double>(2);
      const size_t kw = amb_kd_ws_bytes(sh->S, sh->msub, d);
      void* w = h.alloc<uint8_t>(kw);
      if (!h.rc)
        h.run(amb_kd_subsets(dev, h.st, dC, m, d, dR, n, d, d, dtype, didx, sh->S, sh->msub, AMB_KERNEL_POLY, 1.0 / d, 1.0, 3,
                             1.0, AMB_MMD_UNBIASED, nullptr, kd_dev, w, kw));                // features_1 = candidate
    }
    // ---- radii of this device's row shards
    if (want_prdc && !h.rc) {
      even_rows(n, sh->n_dev, rank, &r0, &rn);
      even_rows(m, sh->n_dev, rank, &c0, &cn);
      pR = h.alloc<uint8_t>(amb_packed_bytes(n, d));
      pC = h.alloc<uint8_t>(amb_packed_bytes(m, d));
      // radii buffers padded to whole shards: the allgather moves equal counts per device
      long long chunk_r = 0, chunk_c = 0, tmp0 = 0;
      even_rows(n, sh->n_dev, 0, &tmp0, &chunk_r);
      even_rows(m, sh->n_dev, 0, &tmp0, &chunk_c);
      if (sh->n_dev == 1) { chunk_r = n; chunk_c = m; }
      rR = h.alloc<float>(static_cast<size_t>(chunk_r) * sh->n_dev);
      rC = h.alloc<float>(static_cast<size_t>(chunk_c) * sh->n_dev);
      wsb = amb_knn_ws_bytes(rn, n, d, k);
      const size_t w2 = amb_knn_ws_bytes(cn, m, d, k), w3 = amb_prdc_ws_bytes(n, m);
      wsb = wsb > w2 ? wsb : w2;
      wsb = wsb > w3 ? wsb : w3;
      ws = h.alloc<uint8_t>(wsb);
      if (!h.rc) {
        h.run(amb_pack(dev, h.st, dR, dtype, n, d, d, pR));
        h.run(amb_pack(dev, h.st, dC, dtype, m, d, d, pC));
        h.run(amb_knn_radii(dev, h.st, dR, dtype, d, pR, n, d, r0, rn, k, rR + r0, nullptr, ws, wsb));
        h.run(amb_knn_radii(dev, h.st, dC, dtype, d, pC, m, d, c0, cn, k, rC + c0, nullptr, ws, wsb));
      }
    }
    if (fad_dev) h.download(&sh->fad, fad_dev, 1);
    if (kd_dev) h.download(sh->kd, kd_dev, 2);
    h.finish();
    if (h.rc) { fail(h.rc); alive = false; }
    bar->wait();                                    // every device got this far (or reported why not)
    for (int q = 0; q < sh->n_dev; ++q) alive = alive && sh->rc[q] == AMB_OK;
    // ---- every device's radii slices to every device: in-place allgather over NVLink
    if (alive && want_prdc && sh->n_dev > 1) {
      long long chunk_r = 0, chunk_c = 0, tmp0 = 0;
      even_rows(n, sh->n_dev, 0, &tmp0, &chunk_r);
      even_rows(m, sh->n_dev, 0, &tmp0, &chunk_c);
      h.run(amb_comm_allgather(sh->comm, rank, rR + chunk_r * rank, rR, chunk_r, AMB_F32, h.st));
      h.run(amb_comm_allgather(sh->comm, rank, rC + chunk_c * rank, rC, chunk_c, AMB_F32, h.st));
    }
    // ---- counts of this device's reference rows against all candidates
    if (alive && want_prdc) {
      int32_t* col = h.alloc<int32_t>(m);
      uint8_t* rec = h.alloc<uint8_t>(rn ? rn : 1);
      uint8_t* cov = h.alloc<uint8_t>(rn ? rn : 1);
      long long* totals = h.alloc<long long>(8);
      long long t[8] = {0};
      void* big = nullptr;
      // the overflow ladder of amb200.h: default list -> list of the reported size -> exhaustive kernel
      for (int attempt = 0; attempt < 3 && !h.rc; ++attempt) {
        h.zero(col, static_cast<size_t>(m) * 4);
        h.zero(totals, 64);
        if (attempt == 0)
          h.run(amb_prdc_counts(dev, h.st, dR, d, pR, n, rR, dC, d, pC, m, rC, d, dtype, r0, rn, col, rec, cov, totals + 4, ws, wsb));
        else if (attempt == 1 && big)
          h.run(amb_prdc_counts(dev, h.st, dR, d, pR, n, rR, dC, d, pC, m, rC, d, dtype, r0, rn, col, rec, cov, totals + 4, big,
                                amb_prdc_ws_bytes_cap(n, m, t[4])));
        else
          h.run(amb_prdc_counts_exact(dev, h.st, dR, d, n, rR, dC, d, m, rC, d, dtype, r0, rn, col, rec, cov));
        h.download(t, totals, 8);                   // t[4]: near-tie pairs this device met
        h.finish();
        if (h.rc) break;
        const long long cap = attempt == 0 ? amb_prdc_ws_list_cap(n, m, wsb) : t[4];
        if (attempt == 2 || t[4] <= cap) break;
        if (attempt == 0) {
          size_t free_b = 0, total_b = 0;
          const size_t bb = amb_prdc_ws_bytes_cap(n, m, t[4]);
          if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && bb < free_b / 2 && cudaMalloc(&big, bb) == cudaSuccess)
            h.bufs.push_back(big);
          else
            big = nullptr;
          (void)cudaGetLastError();
        }
      }
      // ---- per-candidate counts summed over the devices (allreduce), then the four numerators.
      // A device that failed must not leave the others waiting inside the collective: agree first.
      if (h.rc) fail(h.rc);
      if (sh->n_dev > 1) {
        bar->wait();
        bool all_ok = true;
        for (int q = 0; q < sh->n_dev; ++q) all_ok = all_ok && sh->rc[q] == AMB_OK;
        if (!all_ok) return;
        h.run(amb_comm_allreduce(sh->comm, rank, col, col, m, AMB_I32, AMB_SUM, h.st));
      }
      if (!h.rc) {
        h.zero(totals, 32);
        h.run(amb_prdc_reduce(dev, h.st, col, m, rec, cov, rn, totals));
        h.download(t, totals, 4);
        h.finish();
        if (!h.rc)
          for (int q = 0; q < 4; ++q) sh->totals[4 * rank + q] = t[q];
      }
      if (h.rc) fail(h.rc);
    }
  }
}

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

int amb_host_evaluate(const int* devs, int n_dev, const void* ref, long long n, const void* cand, long long m, int d,
                      int dtype, int k, const int32_t* kd_idx, int S, int msub, int want_fad, double* out) {
  if (!devs || n_dev < 1 || n_dev > 64 || !ref || !cand || !out || n <= 0 || m <= 0 || d <= 0 ||
      (dtype != AMB_F32 && dtype != AMB_F64) || k < 0 || (kd_idx && (S <= 0 || msub <= 0)))
    return set_error(AMB_ERR_ARG, "amb_host_evaluate: bad argument");
  if (k > 0 && (k + 1 > n || k + 1 > m))
    return set_error(AMB_ERR_ARG, "amb_host_evaluate: k=%d needs at least k+1 rows in both sets", k);
  EvalShared sh;
  sh.ref = ref; sh.cand = cand; sh.n = n; sh.m = m; sh.d = d; sh.dtype = dtype; sh.k = k;
  sh.kd_idx = kd_idx; sh.S = S; sh.msub = msub; sh.want_fad = want_fad; sh.n_dev = n_dev;
  sh.totals.assign(static_cast<size_t>(4 * n_dev), 0);
  sh.rc.assign(n_dev, AMB_OK);
  sh.err.assign(n_dev, std::string());
  const int workers = k > 0 ? n_dev : 1;          // without PRDC everything is N-independent: one device
  sh.n_dev = workers;
  if (workers > 1) {
    // one set of communicators per device list, created on first use and kept (ncclCommInitAll
    // costs a few hundred milliseconds)
    static std::mutex mu;
    static std::map<std::vector<int>, amb_comm_t*> comms;
    std::lock_guard<std::mutex> lk(mu);
    const std::vector<int> key(devs, devs + workers);
    auto it = comms.find(key);
    if (it == comms.end()) {
      amb_comm_t* c = nullptr;
      const int rc = amb_comm_init(devs, workers, &c);
      if (rc) return rc;
      it = comms.emplace(key, c).first;
    }
    sh.comm = it->second;
  }
  ThreadBarrier bar(workers);
  std::vector<std::thread> th;
  for (int r = 0; r < workers; ++r) th.emplace_back(eval_worker, &sh, &bar, r, devs[r]);
  for (auto& t : th) t.join();
  for (int r = 0; r < workers; ++r)
    if (sh.rc[r] != AMB_OK) return set_error(sh.rc[r], "amb_host_evaluate (device %d): %s", devs[r], sh.err[r].c_str());
  const double nan = __builtin_nan("");
  for (int i = 0; i < 7; ++i) out[i] = nan;
  if (want_fad) out[0] = sh.fad;
  if (kd_idx) { out[1] = sh.kd[0]; out[2] = sh.kd[1]; }
  if (k > 0) {
    // hits and the sum of counts come from the reduced column vector (the same on every device);
    // the row flags belong to each device's own reference rows
    const long long hits = sh.totals[0], total = sh.totals[1];
    long long recalled = 0, covered = 0;
    for (int r = 0; r < workers; ++r) { recalled += sh.totals[4 * r + 2]; covered += sh.totals[4 * r + 3]; }
    out[3] = static_cast<double>(hits) / static_cast<double>(m);                                        // prdc.py:36-38
    out[4] = static_cast<double>(recalled) / static_cast<double>(n);                                    // prdc.py:40-42
    out[5] = (1.0 / static_cast<double>(k)) * (static_cast<double>(total) / static_cast<double>(m));    // prdc.py:44-46
    out[6] = static_cast<double>(covered) / static_cast<double>(n);                                     // prdc.py:48
  }
  return AMB_OK;
}

}  // extern "C"
