// PROTOTYPE for the next round, not part of the product library: symmetric eigen-decomposition of the reduced camera
// system by Householder tridiagonalisation + implicit-shift QL in ONE CTA, as a replacement for the one-sided Jacobi iteration of
// marginal_eig_kernel (5.1 ms at n = 106, bound by 17 sweeps x 105 barrier-separated steps of division / square-root chains).
// Expected cost here: n - 2 reflections of O(n^2 / 1024) work each, then ~1.7 n QL iterations whose rotation chains are computed by
// one thread and applied by a thread per row of the eigenvector matrix.
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/sym_eig_tridiag tools/sym_eig_tridiag.cu
//   tools/sym_eig_tridiag                 self-test on synthetic spectra (serial reference always; the kernel when a GPU is present)
//   tools/sym_eig_tridiag <file> <n>      eigenvalues of a row-major n x n matrix of doubles (serial reference), one per line
//
// The serial routine below is the specification of the kernel: same reflections, same rotations, in the same order.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

// ---- serial reference: A (row-major, symmetric) -> d (eigenvalues, unsorted), Z (row-major, columns = eigenvectors) ---------------
static bool sym_eig_serial(std::vector<double>& A, int n, std::vector<double>& d, std::vector<double>& Z) {
  std::vector<double> e(n, 0.0), v(n), p(n), w(n), q(n);
  Z.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; ++i) Z[(size_t)i * n + i] = 1.0;
  for (int k = 0; k + 2 < n; ++k) {  // Householder: zero A[k+2.., k]
    double norm2 = 0.0;
    for (int i = k + 1; i < n; ++i) norm2 += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    const double norm = std::sqrt(norm2);
    if (norm == 0.0) continue;
    const double x0 = A[(size_t)(k + 1) * n + k];
    const double alpha = x0 >= 0.0 ? -norm : norm;
    double vn2 = 0.0;
    for (int i = k + 1; i < n; ++i) {
      v[i] = A[(size_t)i * n + k] - (i == k + 1 ? alpha : 0.0);
      vn2 += v[i] * v[i];
    }
    if (vn2 == 0.0) continue;
    const double ivn = 1.0 / std::sqrt(vn2);
    for (int i = k + 1; i < n; ++i) v[i] *= ivn;
    double K = 0.0;
    for (int j = k + 1; j < n; ++j) {
      double s = 0.0;
      for (int i = k + 1; i < n; ++i) s += A[(size_t)j * n + i] * v[i];
      p[j] = s;
      K += v[j] * s;
    }
    for (int j = k + 1; j < n; ++j) w[j] = 2.0 * (p[j] - K * v[j]);
    for (int j = k + 1; j < n; ++j)
      for (int i = k + 1; i < n; ++i) A[(size_t)j * n + i] -= v[j] * w[i] + w[j] * v[i];
    A[(size_t)(k + 1) * n + k] = A[(size_t)k * n + k + 1] = alpha;
    for (int i = k + 2; i < n; ++i) A[(size_t)i * n + k] = A[(size_t)k * n + i] = 0.0;
    for (int r = 0; r < n; ++r) {  // Z <- Z H
      double s = 0.0;
      for (int i = k + 1; i < n; ++i) s += Z[(size_t)r * n + i] * v[i];
      q[r] = 2.0 * s;
    }
    for (int r = 0; r < n; ++r)
      for (int i = k + 1; i < n; ++i) Z[(size_t)r * n + i] -= q[r] * v[i];
  }
  d.resize(n);
  for (int i = 0; i < n; ++i) d[i] = A[(size_t)i * n + i];
  for (int i = 0; i + 1 < n; ++i) e[i] = A[(size_t)(i + 1) * n + i];
  // implicit-shift QL on (d, e), e[i] couples d[i] and d[i + 1]
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m + 1 < n; ++m) {
        const double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
        if (std::fabs(e[m]) <= 2.220446049250313e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return false;
        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, pp = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          r = std::hypot(f, g);
          e[i + 1] = r;
          if (r == 0.0) {
            d[i + 1] -= pp;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = d[i + 1] - pp;
          r = (d[i] - g) * s + 2.0 * c * b;
          pp = s * r;
          d[i + 1] = g + pp;
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = Z[(size_t)k * n + i + 1];
            Z[(size_t)k * n + i + 1] = s * Z[(size_t)k * n + i] + c * f;
            Z[(size_t)k * n + i] = c * Z[(size_t)k * n + i] - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        d[l] -= pp;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  return true;
}

// ---- one-CTA kernel: the same algorithm, reflections and rotation chains applied in parallel -----------------------------------------
constexpr int T = 1024;
__device__ __forceinline__ double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int i = 0; i < T / 32; ++i) s += red[i];  // fixed order, identical in every thread
  return s;
}
__global__ void __launch_bounds__(T, 1) sym_eig_kernel(double* __restrict__ A, int n, double* __restrict__ d_out, double* __restrict__ Z, int* __restrict__ status) {
  __shared__ double red[T / 32];
  __shared__ double v[256], w[256], d[256], e[256], cs[256], sn[256];
  __shared__ int s_m, s_lo, s_fail;
  const int tid = threadIdx.x;
  for (int i = tid; i < n * n; i += T) Z[i] = (i / n == i % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int k = 0; k + 2 < n; ++k) {
    double part = 0.0;
    for (int i = k + 1 + tid; i < n; i += T) part += A[(size_t)i * n + k] * A[(size_t)i * n + k];
    const double norm = sqrt(block_sum(part, red));
    if (norm == 0.0) continue;  // uniform
    const double x0 = A[(size_t)(k + 1) * n + k];
    const double alpha = x0 >= 0.0 ? -norm : norm;
    part = 0.0;
    for (int i = k + 1 + tid; i < n; i += T) {
      const double vi = A[(size_t)i * n + k] - (i == k + 1 ? alpha : 0.0);
      v[i] = vi;
      part += vi * vi;
    }
    const double vn2 = block_sum(part, red);
    if (vn2 == 0.0) continue;
    const double ivn = 1.0 / sqrt(vn2);
    for (int i = k + 1 + tid; i < n; i += T) v[i] *= ivn;
    __syncthreads();
    // p = A_sub v: a warp per row j (coalesced along i), K = v^T p
    part = 0.0;
    for (int j = k + 1 + (tid >> 5); j < n; j += T / 32) {
      double s = 0.0;
      for (int i = k + 1 + (tid & 31); i < n; i += 32) s += A[(size_t)j * n + i] * v[i];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) {
        w[j] = s;  // p for now
        part += v[j] * s;
      }
    }
    const double K = block_sum(part, red);
    for (int j = k + 1 + tid; j < n; j += T) w[j] = 2.0 * (w[j] - K * v[j]);
    __syncthreads();
    const int m = n - k - 1;
    for (int idx = tid; idx < m * m; idx += T) {
      const int j = k + 1 + idx / m, i = k + 1 + idx % m;
      A[(size_t)j * n + i] -= v[j] * w[i] + w[j] * v[i];
    }
    for (int i = k + 1 + tid; i < n; i += T) {
      const double val = i == k + 1 ? alpha : 0.0;
      A[(size_t)i * n + k] = val;
      A[(size_t)k * n + i] = val;
    }
    // Z <- Z H: a warp per row r
    for (int r = tid >> 5; r < n; r += T / 32) {
      double s = 0.0;
      for (int i = k + 1 + (tid & 31); i < n; i += 32) s += Z[(size_t)r * n + i] * v[i];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      s *= 2.0;
      for (int i = k + 1 + (tid & 31); i < n; i += 32) Z[(size_t)r * n + i] -= s * v[i];
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += T) {
    d[i] = A[(size_t)i * n + i];
    e[i] = i + 1 < n ? A[(size_t)(i + 1) * n + i] : 0.0;
  }
  if (tid == 0) s_fail = 0;
  __syncthreads();
  for (int l = 0; l < n; ++l) {
    int iter = 0;
    for (;;) {
      if (tid == 0) {  // one thread walks the chain: split point, shift, rotations (c_i, s_i) for i = m - 1 .. lo
        int m;
        for (m = l; m + 1 < n; ++m)
          if (fabs(e[m]) <= 2.220446049250313e-16 * (fabs(d[m]) + fabs(d[m + 1]))) break;
        s_m = m;
        s_lo = m;  // no rotations unless set below
        if (m != l) {
          if (iter == 60) s_fail = 1;
          double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
          double r = hypot(g, 1.0);
          g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
          double s = 1.0, c = 1.0, pp = 0.0;
          int i;
          for (i = m - 1; i >= l; --i) {
            const double f = s * e[i], b = c * e[i];
            r = hypot(f, g);
            e[i + 1] = r;
            if (r == 0.0) {
              d[i + 1] -= pp;
              e[m] = 0.0;
              break;
            }
            s = f / r;
            c = g / r;
            g = d[i + 1] - pp;
            r = (d[i] - g) * s + 2.0 * c * b;
            pp = s * r;
            d[i + 1] = g + pp;
            g = c * r - b;
            cs[i] = c;
            sn[i] = s;
          }
          s_lo = i + 1;  // rotations exist for indices m - 1 .. s_lo
          if (!(r == 0.0 && i >= l)) {  // otherwise: a zero rotation split the chain, iterate again without the shift update
            d[l] -= pp;
            e[l] = g;
            e[m] = 0.0;
          }
        }
      }
      __syncthreads();
      const int m = s_m, lo = s_lo;
      if (s_fail) break;
      if (m == l) break;
      for (int k = tid; k < n; k += T) {  // a thread per row of Z applies the chain
        double* z = Z + (size_t)k * n;
        for (int i = m - 1; i >= lo; --i) {
          const double f = z[i + 1];
          z[i + 1] = sn[i] * z[i] + cs[i] * f;
          z[i] = cs[i] * z[i] - sn[i] * f;
        }
      }
      ++iter;
      __syncthreads();
    }
    if (s_fail) break;
  }
  for (int i = tid; i < n; i += T) d_out[i] = d[i];
  if (tid == 0) status[0] = s_fail;
}

// ---- checks ---------------------------------------------------------------------------------------------------------------------------
static double residual(const std::vector<double>& A0, int n, const std::vector<double>& d, const std::vector<double>& Z, double* ortho) {
  double res = 0.0, nrm = 0.0, orth = 0.0;
  for (int i = 0; i < n * n; ++i) nrm = std::max(nrm, std::fabs(A0[i]));
  for (int c = 0; c < n; ++c)
    for (int r = 0; r < n; ++r) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += A0[(size_t)r * n + k] * Z[(size_t)k * n + c];
      res = std::max(res, std::fabs(s - d[c] * Z[(size_t)r * n + c]));
    }
  for (int a = 0; a < n; ++a)
    for (int b = 0; b < n; ++b) {
      double s = 0.0;
      for (int k = 0; k < n; ++k) s += Z[(size_t)k * n + a] * Z[(size_t)k * n + b];
      orth = std::max(orth, std::fabs(s - (a == b ? 1.0 : 0.0)));
    }
  *ortho = orth;
  return res / nrm;
}

static std::vector<double> test_matrix(int n, int rank, double span, unsigned seed) {
  // Q diag(lambda) Q^T with a random orthogonal Q (product of reflections) and a log-uniform spectrum over `span` decades
  std::vector<double> A((size_t)n * n, 0.0);
  srand(seed);
  for (int i = 0; i < n; ++i) A[(size_t)i * n + i] = i < rank ? std::pow(10.0, 8.0 - span * i / std::max(1, rank - 1)) : 0.0;
  std::vector<double> u(n), t(n);
  for (int rep = 0; rep < 6; ++rep) {
    double nu = 0.0;
    for (int i = 0; i < n; ++i) { u[i] = rand() / (double)RAND_MAX - 0.5; nu += u[i] * u[i]; }
    for (int i = 0; i < n; ++i) u[i] /= std::sqrt(nu);
    for (int side = 0; side < 2; ++side)  // A <- H A, then A <- A H
      for (int c = 0; c < n; ++c) {
        double s = 0.0;
        for (int k = 0; k < n; ++k) s += u[k] * (side == 0 ? A[(size_t)k * n + c] : A[(size_t)c * n + k]);
        for (int k = 0; k < n; ++k) (side == 0 ? A[(size_t)k * n + c] : A[(size_t)c * n + k]) -= 2.0 * s * u[k];
      }
  }
  for (int r = 0; r < n; ++r)
    for (int c = r + 1; c < n; ++c) A[(size_t)c * n + r] = A[(size_t)r * n + c] = 0.5 * (A[(size_t)r * n + c] + A[(size_t)c * n + r]);
  return A;
}

int main(int argc, char** argv) {
  if (argc == 3) {  // eigenvalues of a matrix file (serial reference): used by the Python cross-check against numpy
    const int n = atoi(argv[2]);
    std::vector<double> A((size_t)n * n), d, Z;
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(A.data(), sizeof(double), A.size(), f) != A.size()) return 2;
    fclose(f);
    std::vector<double> A0 = A;
    if (!sym_eig_serial(A, n, d, Z)) return 3;
    double orth;
    const double res = residual(A0, n, d, Z, &orth);
    fprintf(stderr, "residual %.3e orthogonality %.3e\n", res, orth);
    for (double x : d) printf("%.17g\n", x);
    return 0;
  }
  int n_dev = 0;
  const bool gpu = cudaGetDeviceCount(&n_dev) == cudaSuccess && n_dev > 0;
  bool all_ok = true;
  for (int n : {8, 22, 47, 106, 218}) {
    for (int deficient = 0; deficient < 2; ++deficient) {
      const int rank = deficient ? n - std::max(1, n / 8) : n;
      std::vector<double> A0 = test_matrix(n, rank, 14.0, 1234u + n), A = A0, d, Z;
      const bool ok = sym_eig_serial(A, n, d, Z);
      double orth;
      const double res = residual(A0, n, d, Z, &orth);
      printf("n=%3d rank=%3d serial: converged=%d residual/|A|=%.2e orthogonality=%.2e", n, rank, ok, res, orth);
      all_ok = all_ok && ok && res < 1e-13 * n && orth < 1e-13 * n;
      if (gpu) {
        double *dA, *dd, *dZ;
        int* ds;
        cudaMalloc(&dA, sizeof(double) * n * n);
        cudaMalloc(&dZ, sizeof(double) * n * n);
        cudaMalloc(&dd, sizeof(double) * n);
        cudaMalloc(&ds, sizeof(int));
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
          cudaMemcpy(dA, A0.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
          cudaEventRecord(e0);
          sym_eig_kernel<<<1, T>>>(dA, n, dd, dZ, ds);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          best = std::min(best, ms);
        }
        std::vector<double> gd(n), gZ((size_t)n * n);
        int st = 0;
        cudaMemcpy(gd.data(), dd, sizeof(double) * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(gZ.data(), dZ, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, ds, sizeof(int), cudaMemcpyDeviceToHost);
        double gorth;
        const double gres = residual(A0, n, gd, gZ, &gorth);
        std::vector<double> a = d, b = gd;
        std::sort(a.begin(), a.end());
        std::sort(b.begin(), b.end());
        double dev = 0.0;
        for (int i = 0; i < n; ++i) dev = std::max(dev, std::fabs(a[i] - b[i]));
        printf(" | kernel: %.3f ms status=%d residual=%.2e orthogonality=%.2e max|lambda - serial|/|A|=%.2e", best, st, gres, gorth, dev / 1e8);
        all_ok = all_ok && st == 0 && gres < 1e-13 * n && gorth < 1e-13 * n;
        cudaFree(dA); cudaFree(dZ); cudaFree(dd); cudaFree(ds);
      }
      printf("\n");
    }
  }
  printf(all_ok ? "SYM_EIG PASS\n" : "SYM_EIG FAIL\n");
  return all_ok ? 0 : 1;
}
