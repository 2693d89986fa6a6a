// Stand-alone check + timing of csrc/gemm3x_tma.cuh (TMA-fed 3xTF32 GEMM on tcgen05), all four operand-major cases.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -I../unsupervised-asr_b200/csrc gemm_tma.cu -o gemm_tma
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "gemm3x_tma.cuh"

using namespace eodm_tma;

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(1);                                                                        \
    }                                                                                 \
  } while (0)

__global__ void fill_kernel(float* x, size_t n, uint32_t seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u ^ seed;
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    const float u = (h >> 8) * (1.f / 16777216.f);
    x[i] = (h & 7) == 0 ? u * 1e-6f : u;   // some tiny entries, like a peaked posterior
  }
}
__global__ void split_kernel(const float* x, float* lo, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    const float r = v - __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    lo[i] = __uint_as_float((__float_as_uint(r) + 0x1000u) & 0xffffe000u);
  }
}

struct Mat {
  float *x = nullptr, *lo = nullptr;
  long long rows, cols;
  void init(long long r, long long c, uint32_t seed) {
    rows = r; cols = c;
    CK(cudaMalloc(&x, (size_t)r * c * 4));
    CK(cudaMalloc(&lo, (size_t)r * c * 4));
    fill_kernel<<<1024, 256>>>(x, (size_t)r * c, seed);
    split_kernel<<<1024, 256>>>(x, lo, (size_t)r * c);
    CK(cudaDeviceSynchronize());
  }
  void free_() { cudaFree(x); cudaFree(lo); }
};

// operand stored [K][MN] if mn-major else [MN][K]
template <bool A_MN, bool B_MN>
double run(int M, int N, int K, bool check, int reps, const char* name, bool two = false, int split_from = -1) {
  Mat A, B;
  if (A_MN) A.init(K, M, 1u); else A.init(M, K, 1u);
  if (B_MN) B.init(K, N, 2u); else B.init(N, K, 2u);
  float *C, *so;
  CK(cudaMalloc(&C, (size_t)(M + 1) * N * 4));
  CK(cudaMalloc(&so, (size_t)M * 4));
  fill_kernel<<<64, 256>>>(so, M, 3u);
  CK(cudaMemset(C, 0, (size_t)(M + 1) * N * 4));
  CUtensorMap ta, tal, tb, tbl;
  bool ok = make_operand_map(&ta, A.x, A.rows, A.cols, A.cols, A_MN, kTM) && make_operand_map(&tal, A.lo, A.rows, A.cols, A.cols, A_MN, kTM) &&
            make_operand_map(&tb, B.x, B.rows, B.cols, B.cols, B_MN, two ? 128 : kTN) &&
            make_operand_map(&tbl, B.lo, B.rows, B.cols, B.cols, B_MN, two ? 128 : kTN);
  if (!ok) { printf("tensor map encode failed\n"); exit(1); }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  Args a;
  a.M = M; a.N = N; a.K = K; a.scale_out = check ? so : nullptr; a.C = C; a.ldc = N; a.c_row_shift = check ? 1 : 0;
  a.accumulate = 0; a.m_tiles = two ? (M + 255) / 256 : (M + kTM - 1) / kTM; a.n_tiles = (N + kTN - 1) / kTN;
  float* C2 = nullptr;
  a.split_from = a.m_tiles * a.n_tiles; a.C2 = nullptr;
  if (two && split_from >= 0) {
    CK(cudaMalloc(&C2, (size_t)(M + 1) * N * 4));
    CK(cudaMemset(C2, 0, (size_t)(M + 1) * N * 4));
    a.split_from = split_from; a.C2 = C2;
  }
  auto go = [&]() { return two ? launch2<A_MN, B_MN>(ta, tal, tb, tbl, a, sms, 0) : launch<A_MN, B_MN>(ta, tal, tb, tbl, a, sms, 0); };
  CK(go());
  CK(cudaDeviceSynchronize());
  double result = 0;
  if (check) {
    a.accumulate = 1;   // second pass doubles the result
    CK(go());
    CK(cudaDeviceSynchronize());
    std::vector<float> ha((size_t)M * K), hb((size_t)N * K), hc((size_t)(M + 1) * N), hs(M);
    CK(cudaMemcpy(ha.data(), A.x, ha.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hb.data(), B.x, hb.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.data(), C, hc.size() * 4, cudaMemcpyDeviceToHost));
    if (C2) {   // the second halves of the split tiles
      std::vector<float> h2(hc.size());
      CK(cudaMemcpy(h2.data(), C2, h2.size() * 4, cudaMemcpyDeviceToHost));
      for (size_t q = 0; q < hc.size(); ++q) hc[q] += h2[q];
    }
    CK(cudaMemcpy(hs.data(), so, hs.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0;
    int bad_i = -1, bad_j = -1;
    for (int i = 0; i < M; ++i)
      for (int j = 0; j < N; ++j) {
        double s = 0;
        for (int k = 0; k < K; ++k) {
          const double x = A_MN ? ha[(size_t)k * M + i] : ha[(size_t)i * K + k];
          const double y = B_MN ? hb[(size_t)k * N + j] : hb[(size_t)j * K + k];
          s += x * y;
        }
        s *= 2.0 * hs[i];
        const double err = fabs(hc[(size_t)(i + 1) * N + j] - s) / (fabs(s) + 1e-30);
        if (err > maxerr) { maxerr = err; bad_i = i; bad_j = j; }
      }
    double row0 = 0;
    for (int j = 0; j < N; ++j) row0 += fabs(hc[j]);
    printf("%-28s M=%d N=%d K=%d  max rel err %.3e at (%d,%d)  |row 0| = %g (0 expected)\n", name, M, N, K, maxerr, bad_i, bad_j, row0);
    result = maxerr;
  } else {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) CK(go());
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double tf = 2.0 * M * N * K / ms * 1e-9;
    printf("%-28s M=%d N=%d K=%d  %.3f ms  %.1f TFLOP/s algorithmic  (%.1f %% of 1003/3 tf32 peak)\n", name, M, N, K, ms, tf, tf / (1003.0 / 3) * 100);
    result = ms;
  }
  A.free_(); B.free_(); cudaFree(C); cudaFree(so);
  if (C2) cudaFree(C2);
  return result;
}

int main(int argc, char** argv) {
  const bool timing = argc > 1 && atoi(argv[1]) > 0;
  double worst = 0;
  worst = fmax(worst, run<false, false>(300, 640, 200, true, 0, "A K-major,  B K-major"));
  worst = fmax(worst, run<true, false>(288, 640, 200, true, 0, "A MN-major, B K-major"));
  worst = fmax(worst, run<false, true>(300, 640, 200, true, 0, "A K-major,  B MN-major"));
  worst = fmax(worst, run<true, true>(288, 640, 200, true, 0, "A MN-major, B MN-major"));
  worst = fmax(worst, run<true, true>(128, 256, 16, true, 0, "MN/MN one chunk"));
  worst = fmax(worst, run<false, false>(128, 256, 16, true, 0, "K/K one chunk"));
  worst = fmax(worst, run<true, true>(20 * 128, 8 * 256, 1000, true, 0, "MN/MN 160 tiles (persistent)"));
  printf("worst %.3e -> %s\n", worst, worst < 5e-6 ? "OK" : "FAIL");
  if (argc > 2 && atoi(argv[2]) > 0) {   // cta_group::2 variants
    double w2 = 0;
    w2 = fmax(w2, run<false, false>(128, 256, 16, true, 0, "2-CTA K/K one chunk", true));
    w2 = fmax(w2, run<false, false>(300, 640, 200, true, 0, "2-CTA K/K", true));
    w2 = fmax(w2, run<true, false>(288, 640, 200, true, 0, "2-CTA MN/K", true));
    w2 = fmax(w2, run<false, true>(300, 640, 200, true, 0, "2-CTA K/MN", true));
    w2 = fmax(w2, run<true, true>(288, 640, 200, true, 0, "2-CTA MN/MN", true));
    w2 = fmax(w2, run<true, true>(20 * 128, 8 * 256, 1000, true, 0, "2-CTA MN/MN 80 tiles", true));
    w2 = fmax(w2, run<true, true>(20 * 128, 8 * 256, 1000, true, 0, "2-CTA MN/MN, tiles 74.. split", true, 74));
    w2 = fmax(w2, run<false, true>(300, 640, 200, true, 0, "2-CTA K/MN, every tile split", true, 0));
    printf("2-CTA worst %.3e -> %s\n", w2, w2 < 5e-6 ? "OK" : "FAIL");
    worst = fmax(worst, w2);
    if (timing && w2 < 5e-6) {
      run<true, true>(5120, 5120, 32768, false, 3, "2-CTA fwd  (MN/MN)", true);
      run<true, true>(5120, 5120, 32768, false, 3, "2-CTA fwd, last wave split", true, 370);
      run<false, false>(32768, 5120, 5120, false, 3, "2-CTA bwd1 (K/K)", true);
      run<false, true>(32768, 5120, 5120, false, 3, "2-CTA bwd2 (K/MN)", true);
    }
  }
  if (timing) {
    run<true, true>(5120, 5120, 32768, false, 3, "fwd  (MN/MN)");
    run<false, false>(32768, 5120, 5120, false, 3, "bwd1 (K/K)");
    run<false, true>(32768, 5120, 5120, false, 3, "bwd2 (K/MN)");
  }
  return worst < 5e-6 ? 0 : 1;
}
