// Evidence for the kernel choices of the coefficient update (north star: "no tensor cores unless an ncu-measured
// b1 AND.popc MMA path actually beats the popcount kernels"):
//
//   1. popc32 issue peak of an SM (the "XU pipe" ceiling the coefficient kernel is reported against);
//   2. one greedy pass of update_coefficients (for every row: argmin over all atoms of |E_i xor D_k|, lowest k on ties,
//      src/bsvd.cpp:1067-1082) done two ways on the same data:
//        popc : lane per row, XOR + POPC per word (what csrc/coef.cu does),
//        mma  : mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc -- 16 rows x 8 atoms x 256 bits per instruction;
//               |E_i xor D_k| = |E_i| + |D_k| - 2 |E_i and D_k|. sm_100a has no native 1-bit MMA: ptxas expands it into
//               8 x IMMA.16832.U8 on bit-sliced operands (check with cuobjdump -sass), so this measures whether the
//               int8 tensor pipe beats the popcount pipe at this job.
//      Both produce the same keys (checked); times are CUDA-event means over REPS launches.
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o profiles/_bin/microbench profiles/microbench.cu
// Run  : profiles/_bin/microbench > gpurun_out/microbench.json
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ------------------------------------------------------------------ 1. popc issue peak
template <int CHAINS>
__global__ void __launch_bounds__(1024) k_popc_peak(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[CHAINS];
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) a[j] = seed * (threadIdx.x + 1) + j * 0x9E3779B9u;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) a[j] = __popc(a[j] ^ seed) + a[j];  // LOP3 + POPC + IADD per step, CHAINS independent chains
  }
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < CHAINS; ++j) s ^= a[j];
  if (s == 0xDEADBEEFu) out[0] = s;  // keeps the loop alive
}

// ------------------------------------------------------------------ 2. one greedy pass, two ways
template <int WORDS>
__global__ void __launch_bounds__(256) k_pass_popc(const uint32_t* __restrict__ E, const uint32_t* __restrict__ D, uint32_t* __restrict__ keys,
                                                   uint32_t n, uint32_t p) {
  extern __shared__ uint32_t Ds[];
  for (uint32_t i = threadIdx.x; i < p * WORDS; i += blockDim.x) Ds[i] = D[i];
  __syncthreads();
  for (uint32_t r = blockIdx.x * blockDim.x + threadIdx.x; r < n; r += gridDim.x * blockDim.x) {
    uint32_t e[WORDS];
#pragma unroll
    for (int w = 0; w < WORDS; ++w) e[w] = E[(size_t)r * WORDS + w];
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t k = 0; k < p; ++k) {
      const uint32_t* dk = Ds + k * WORDS;
      uint32_t d = 0;
#pragma unroll
      for (int w = 0; w < WORDS; ++w) d += __popc(e[w] ^ dk[w]);
      best = min(best, (d << 16) | k);
    }
    keys[r] = best;
  }
}

__device__ __forceinline__ void bmma_16x8x256(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k256.row.col.s32.b1.b1.s32.and.popc {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// a warp takes 16 rows; atoms in groups of 8; KB = WORDS / 8 MMAs of 256 bits per group. D sits in shared memory with a
// padded row stride (WORDS + 4) so the B fragments load without bank conflicts; |D_k| in shared memory too.
template <int WORDS>
__global__ void __launch_bounds__(256) k_pass_mma(const uint32_t* __restrict__ E, const uint32_t* __restrict__ D, uint32_t* __restrict__ keys,
                                                  uint32_t n, uint32_t p) {
  constexpr int STR = WORDS + 4, KB = WORDS / 8;
  extern __shared__ uint32_t sm[];
  uint32_t* Ds = sm;                 // p * STR
  uint32_t* Dw = Ds + (size_t)p * STR;  // p weights
  for (uint32_t i = threadIdx.x; i < p * WORDS; i += blockDim.x) Ds[(i / WORDS) * STR + (i % WORDS)] = D[i];
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < p; k += blockDim.x) {
    uint32_t w = 0;
    for (int j = 0; j < WORDS; ++j) w += __popc(Ds[k * STR + j]);
    Dw[k] = w;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t r0 = warp * 16; r0 < n; r0 += nwarps * 16) {
    uint32_t a[KB][4];
    uint32_t w0 = 0, w1 = 0;         // partial weights of rows r0+g and r0+g+8 (this thread's words)
    const uint32_t* e0 = E + (size_t)(r0 + g) * WORDS;
    const uint32_t* e1 = E + (size_t)(r0 + g + 8) * WORDS;
#pragma unroll
    for (int kb = 0; kb < KB; ++kb) {
      a[kb][0] = e0[kb * 8 + t]; a[kb][1] = e1[kb * 8 + t]; a[kb][2] = e0[kb * 8 + 4 + t]; a[kb][3] = e1[kb * 8 + 4 + t];
      w0 += __popc(a[kb][0]) + __popc(a[kb][2]);
      w1 += __popc(a[kb][1]) + __popc(a[kb][3]);
    }
    w0 += __shfl_xor_sync(0xffffffffu, w0, 1); w0 += __shfl_xor_sync(0xffffffffu, w0, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1); w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
    uint32_t best0 = 0xFFFFFFFFu, best1 = 0xFFFFFFFFu;
    for (uint32_t k0 = 0; k0 < p; k0 += 8) {
      int c[4] = {0, 0, 0, 0};
      const uint32_t* dk = Ds + (size_t)(k0 + g) * STR;
#pragma unroll
      for (int kb = 0; kb < KB; ++kb) {
        uint32_t b[2] = {dk[kb * 8 + t], dk[kb * 8 + 4 + t]};
        bmma_16x8x256(c, a[kb], b);
      }
      const uint32_t ka = k0 + 2 * t, kbi = ka + 1;
      const uint32_t wa = Dw[ka], wb = Dw[kbi];
      best0 = min(best0, ((w0 + wa - 2u * (uint32_t)c[0]) << 16) | ka);
      best0 = min(best0, ((w0 + wb - 2u * (uint32_t)c[1]) << 16) | kbi);
      best1 = min(best1, ((w1 + wa - 2u * (uint32_t)c[2]) << 16) | ka);
      best1 = min(best1, ((w1 + wb - 2u * (uint32_t)c[3]) << 16) | kbi);
    }
    best0 = min(best0, __shfl_xor_sync(0xffffffffu, best0, 1)); best0 = min(best0, __shfl_xor_sync(0xffffffffu, best0, 2));
    best1 = min(best1, __shfl_xor_sync(0xffffffffu, best1, 1)); best1 = min(best1, __shfl_xor_sync(0xffffffffu, best1, 2));
    if (t == 0) { keys[r0 + g] = best0; keys[r0 + g + 8] = best1; }
  }
}

static uint32_t rnd(uint64_t& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(s >> 32); }

template <int WORDS>
static void run_pass(uint32_t n, uint32_t p, int sms, int reps) {
  std::vector<uint32_t> hE((size_t)n * WORDS), hD((size_t)p * WORDS);
  uint64_t s = 12345 + WORDS;
  // sparse-ish rows and atoms (text-like 15 % ink): AND of two random words ~ 25 %, of three ~ 12.5 %
  for (auto& v : hD) v = rnd(s) & rnd(s) & rnd(s);
  for (size_t i = 0; i < hE.size(); ++i) hE[i] = rnd(s) & rnd(s) & rnd(s);
  for (uint32_t r = 0; r < n; r += 3)  // every third row is an atom plus a little noise: realistic near matches and ties
    for (int w = 0; w < WORDS; ++w) hE[(size_t)r * WORDS + w] = hD[(size_t)(r % p) * WORDS + w] ^ (rnd(s) & rnd(s) & rnd(s) & rnd(s) & rnd(s));
  uint32_t *dE, *dD, *k1, *k2;
  CK(cudaMalloc(&dE, hE.size() * 4)); CK(cudaMalloc(&dD, hD.size() * 4)); CK(cudaMalloc(&k1, n * 4)); CK(cudaMalloc(&k2, n * 4));
  CK(cudaMemcpy(dE, hE.data(), hE.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, hD.data(), hD.size() * 4, cudaMemcpyHostToDevice));
  const size_t sm1 = (size_t)p * WORDS * 4, sm2 = (size_t)p * (WORDS + 4) * 4 + p * 4;
  CK(cudaFuncSetAttribute(k_pass_popc<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  CK(cudaFuncSetAttribute(k_pass_mma<WORDS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  int per1 = (int)((200 * 1024) / sm1); per1 = per1 < 1 ? 1 : (per1 > 8 ? 8 : per1);
  int per2 = (int)((200 * 1024) / sm2); per2 = per2 < 1 ? 1 : (per2 > 8 ? 8 : per2);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms1 = 0, ms2 = 0;
  for (int which = 0; which < 2; ++which) {
    for (int it = 0; it < reps + 2; ++it) {
      if (it == 2) CK(cudaEventRecord(e0));
      if (which == 0) k_pass_popc<WORDS><<<sms * per1, 256, sm1>>>(dE, dD, k1, n, p);
      else k_pass_mma<WORDS><<<sms * per2, 256, sm2>>>(dE, dD, k2, n, p);
    }
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(which ? &ms2 : &ms1, e0, e1));
  }
  ms1 /= reps; ms2 /= reps;
  std::vector<uint32_t> h1(n), h2(n);
  CK(cudaMemcpy(h1.data(), k1, n * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(h2.data(), k2, n * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (uint32_t r = 0; r < n; ++r) bad += h1[r] != h2[r];
  const double dist = (double)n * p;  // (row, atom) distances per pass
  printf("{\"bench\": \"greedy_pass\", \"m\": %d, \"p\": %u, \"n\": %u, \"popc_ms\": %.4f, \"mma_ms\": %.4f, \"mma_speedup\": %.3f, "
         "\"popc_Gdist_s\": %.2f, \"mma_Gdist_s\": %.2f, \"popc32_per_clk_per_sm_in_popc_kernel\": %.2f, \"keys_differ\": %zu, "
         "\"ctas_per_sm\": [%d, %d]}\n",
         WORDS * 32, p, n, ms1, ms2, ms1 / ms2, dist / ms1 / 1e6, dist / ms2 / 1e6,
         dist * WORDS / (ms1 * 1e-3) / (sms * 1.965e9), bad, per1, per2);
  cudaFree(dE); cudaFree(dD); cudaFree(k1); cudaFree(k2);
}

int main() {
  int dev = 0, sms = 0, khz = 0;
  CK(cudaSetDevice(dev));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
  uint32_t* out;
  CK(cudaMalloc(&out, 64));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 4096;
  for (int chains = 4; chains <= 8; chains += 4) {
    for (int ctas = 1; ctas <= 2; ++ctas) {
      float best = 1e9f;
      for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0));
        if (chains == 4) k_popc_peak<4><<<sms * ctas, 1024>>>(out, 0x12345u + rep, iters);
        else k_popc_peak<8><<<sms * ctas, 1024>>>(out, 0x12345u + rep, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep && ms < best) best = ms;
      }
      const double popcs = (double)sms * ctas * 1024 * chains * iters;
      printf("{\"bench\": \"popc_peak\", \"chains\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"Gpopc32_per_s\": %.1f, "
             "\"popc32_per_clk_per_sm_at_max_clock\": %.2f, \"sm_count\": %d, \"max_clock_mhz\": %.0f}\n",
             chains, ctas, best, popcs / best / 1e6, popcs / (best * 1e-3) / (sms * (khz * 1e3)), sms, khz / 1e3);
    }
  }
  run_pass<8>(1u << 18, 256, sms, 10);    // 16x16 patches, 256 atoms (configs[2])
  run_pass<32>(1u << 16, 1024, sms, 5);   // 32x32 patches, 1024 atoms (configs[3])
  run_pass<8>(1u << 18, 64, sms, 10);
  return 0;
}
