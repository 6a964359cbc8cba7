// imma_probe.cu -- bring-up / measurement of the INT8 tensor-core form of the modular matrix product (tcgen05.mma kind::i8).
//
//   O[row][d] = sum_j M[row][j] * V[d][j]  (mod q),   M, V < 2^62
// A 64-bit operand is its 8 little-endian bytes, so  M*V = sum_u 2^(8u) sum_{s+t=u} m_s v_t: byte-plane GEMMs accumulated on
// overlapping windows of the TMEM accumulator (imma.cu); the epilogue recombines the 15 diagonal sums of every output.
//
// usage: imma_probe [rows] [D] [k] [planes] [reps] [ell] [mode] [dt] [epilogue warps]     (rows % 256 == 0, D % 16 == 0, k % 16 == 0)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pvw-rs_b200/csrc/imma.cuh"

using namespace pvw;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void fill_kernel(u64* p, size_t n, u64 q, u64 seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    u64 x = seed + i * 0x9E3779B97F4A7C15ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull; x ^= x >> 27; x *= 0x94D049BB133111EBull; x ^= x >> 31;
    p[i] = x % q;
  }
}
// reference: plain modular dot products
__global__ void ref_kernel(const u64* M, const u64* V, u64* O, uint32_t rows, uint32_t D, uint32_t k, LimbConst lc, uint32_t row_lim, uint32_t d_lim) {
  const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x, d = blockIdx.y, plane = blockIdx.z;
  if (row >= row_lim || d >= d_lim) return;
  const u64* m = M + ((size_t)plane * rows + row) * k;
  const u64* v = V + ((size_t)plane * D + d) * k;
  u64 acc = 0;
  for (uint32_t j = 0; j < k; j++) acc = addmod(acc, mulmod(m[j], v[j], lc), lc.q);
  O[((size_t)plane * D + d) * rows + row] = acc;
}

// co-residency probe: a CUDA-core kernel (Shoup multiplies in registers, 128 threads, few registers, no shared memory) to run on a
// second stream beside the persistent tensor-core kernel
__global__ void __launch_bounds__(128) busy_kernel(u64* out, uint32_t iters, u64 q) {
  u64 a = threadIdx.x + blockIdx.x * 131ull + 3, w = 0x123456789abcdefull % q, w_sh = (u64)(((unsigned __int128)w << 64) / q);
  for (uint32_t i = 0; i < iters; i++) a = mulmod_shoup(a, w, w_sh, q) + i;
  if (a == 0x5555) out[0] = a;
}

static LimbConst make_lc(u64 q) {
  LimbConst c{};
  c.q = q;
  unsigned __int128 one = 1;
  // floor(2^128 / q): long division
  unsigned __int128 hi = (~(unsigned __int128)0) / q;   // floor((2^128 - 1) / q) == floor(2^128 / q) unless q | 2^128
  c.mu_hi = (u64)(hi >> 64); c.mu_lo = (u64)hi;
  c.mu64 = (u64)(((one << 64)) / q);
  c.r128 = (u64)((((~(unsigned __int128)0) % q) + 1) % q);
  c.c124 = (u64)((one << 124) % q);
  return c;
}

int main(int argc, char** argv) {
  const uint32_t rows = argc > 1 ? atoi(argv[1]) : 4096, D = argc > 2 ? atoi(argv[2]) : 256, k = argc > 3 ? atoi(argv[3]) : 256;
  const uint32_t planes = argc > 4 ? atoi(argv[4]) : 8, reps = argc > 5 ? atoi(argv[5]) : 5;
  const uint32_t ell = argc > 6 ? atoi(argv[6]) : 1;     // > 1: the library's output layout O[d][limb][row][c] (timing only)
  const int mode = argc > 7 ? atoi(argv[7]) : 2;
  const int dt = argc > 8 ? atoi(argv[8]) : 0;           // 16: half-width tile (N = 128 per MMA)
  const int ew = argc > 9 ? atoi(argv[9]) : 0;           // 16: four epilogue warps per TMEM lane group; default 8
  const int fr = argc > 10 ? atoi(argv[10]) : 1;         // 0: the general reduction of the 160-bit sums
  const u64 q = 0x3ffffffffffffdc1ull;
  const LimbConst lc = make_lc(q);
  u64 *M, *V, *O, *Oref;
  uint8_t *Vb, *Mb;
  LimbConst* dlc;
  CK(cudaMalloc(&M, (size_t)planes * rows * k * 8));
  CK(cudaMalloc(&V, (size_t)planes * D * k * 8));
  const uint32_t kp = imma_kp(k);
  CK(cudaMalloc(&Vb, (size_t)planes * D * 8 * kp));
  CK(cudaMalloc(&Mb, (size_t)planes * rows * 8 * kp));
  CK(cudaMemset(Vb, 0, (size_t)planes * D * 8 * kp));
  CK(cudaMemset(Mb, 0, (size_t)planes * rows * 8 * kp));
  CK(cudaMalloc(&O, (size_t)planes * D * rows * 8));
  CK(cudaMalloc(&Oref, (size_t)planes * D * rows * 8));
  CK(cudaMalloc(&dlc, planes * sizeof(LimbConst)));
  {
    std::vector<LimbConst> lcs(planes, lc);
    CK(cudaMemcpy(dlc, lcs.data(), planes * sizeof(LimbConst), cudaMemcpyHostToDevice));
  }
  fill_kernel<<<1024, 256>>>(M, (size_t)planes * rows * k, q, 1);
  fill_kernel<<<1024, 256>>>(V, (size_t)planes * D * k, q, 2);
  CK(cudaMemset(O, 0xff, (size_t)planes * D * rows * 8));
  // worst case rows: all residues q - 1 in row 0 / dealer 0 of plane 0
  {
    std::vector<u64> ones(k, q - 1);
    CK(cudaMemcpy(M, ones.data(), k * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(V, ones.data(), k * 8, cudaMemcpyHostToDevice));
  }
  ImmaArgs a{};
  a.Mb = Mb; a.Mb_plane = (size_t)rows * 8 * kp; a.rows = rows; a.k = k; a.L = planes; a.ell = 1;   // plane = limb (ell = 1 view)
  a.Vb = Vb; a.Vb_plane = (size_t)D * 8 * kp; a.Vb_D = D; a.d_first = 0; a.D = D;
  a.O = O; a.O_ds = rows; a.O_ls = (size_t)D * rows; a.O_rs = 1; a.O_cs = 0;
  a.lc = dlc; a.mode = 2; a.epi_warps = ew; a.fast_reduce = fr; a.dt = dt == 2 ? 0 : dt; a.pair = dt == 2 ? 1 : 0;   // dt == 2: the two-SM kernel
  ImmaArgs b = a;                                                        // timing variant
  if (ell > 1) { b.L = planes / ell; b.ell = ell; b.O_ds = (size_t)b.L * rows * ell; b.O_ls = (size_t)rows * ell; b.O_rs = ell; b.O_cs = 1; }
  b.mode = mode;
  launch_imma_planes_v(V, k, (size_t)D * k, 1, D, k, planes, Vb, a.Vb_plane, false, nullptr, 0);
  launch_imma_planes_m(M, (size_t)rows * k, k, rows, k, planes, 1, Mb, a.Mb_plane, false, 0);
  CK(cudaGetLastError());
  if (!launch_imma_gemm(a, 0)) { printf("launch_imma_gemm: tensor map creation failed\n"); return 1; }
  CK(cudaDeviceSynchronize());
  // check a window against the reference
  const uint32_t row_lim = rows < 512 ? rows : 512, d_lim = D < 48 ? D : 48;
  ref_kernel<<<dim3((row_lim + 127) / 128, d_lim, planes), 128>>>(M, V, Oref, rows, D, k, lc, row_lim, d_lim);
  CK(cudaDeviceSynchronize());
  std::vector<u64> h((size_t)planes * D * rows), hr((size_t)planes * D * rows);
  CK(cudaMemcpy(h.data(), O, h.size() * 8, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hr.data(), Oref, hr.size() * 8, cudaMemcpyDeviceToHost));
  size_t bad = 0, checked = 0;
  for (uint32_t p = 0; p < planes; p++)
    for (uint32_t d = 0; d < d_lim; d++)
      for (uint32_t r = 0; r < row_lim; r++) {
        const size_t i = ((size_t)p * D + d) * rows + r;
        checked++;
        if (h[i] != hr[i]) { if (bad < 5) printf("mismatch plane %u d %u row %u: got %llx want %llx\n", p, d, r, h[i], hr[i]); bad++; }
      }
  printf("checked %zu outputs, %zu mismatches\n", checked, bad);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int pass = 0; pass < 2; pass++) {
    CK(cudaEventRecord(e0));
    for (uint32_t i = 0; i < reps; i++) {
      if (pass == 1) launch_imma_planes_v(V, k, (size_t)D * k, 1, D, k, planes, Vb, a.Vb_plane, false, nullptr, 0);
      else launch_imma_gemm(b, 0);
    }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double macs = (double)rows * D * k * planes * reps;
    if (pass == 0) printf("imma_gemm: %.3f ms per launch, %.3e 62-bit MAC/s, %.1f int8 TOPS\n", ms / reps, macs / (ms * 1e-3), macs * 64 * 2 / (ms * 1e-3) / 1e12);
    else printf("imma_planes_v: %.3f ms per launch (%.1f GB/s written)\n", ms / reps, (double)planes * D * 8 * kp * reps / (ms * 1e-3) / 1e9);
  }
  // co-residency: imma_probe ... [stages] [co-resident CTAs' carve-out: 0 default, 1 max shared]
  if (argc > 11) {
    const int stages = atoi(argv[11]), carve = argc > 12 ? atoi(argv[12]) : 0;
    b.stages = stages;
    cudaStream_t s1, s2;
    CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
    if (carve) CK(cudaFuncSetAttribute(busy_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const uint32_t iters = 20000, grid = 148 * 16;
    auto run = [&](bool gemm, bool busy) {
      float best = 1e9f;
      for (int rep = 0; rep < 3; rep++) {
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        CK(cudaStreamWaitEvent(s2, e0, 0));
        if (gemm) launch_imma_gemm(b, s1);
        if (busy) busy_kernel<<<grid, 128, 0, s2>>>(O, iters, q);
        cudaEvent_t j; CK(cudaEventCreate(&j)); CK(cudaEventRecord(j, s2)); CK(cudaStreamWaitEvent(s1, j, 0));
        CK(cudaEventRecord(e1, s1));
        CK(cudaDeviceSynchronize());
        float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = ms < best ? ms : best;
        CK(cudaEventDestroy(j));
      }
      return best;
    };
    const float tg = run(true, false), tb = run(false, true), tgb = run(true, true);
    printf("co-residency (ring stages %d, carve-out %d): imma %.3f ms, busy %.3f ms, both %.3f ms (sum %.3f)\n", stages, carve, tg, tb, tgb, tg + tb);
  }
  return bad ? 2 : 0;
}
