// Issue-rate microbenchmarks for packed FP32 (FFMA2 / FMUL2 / FADD2) and the mixes of the round-2 rate consumers (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_packed.bin scripts/ubench_packed.cu
// Prints cycles per warp-instruction per SM sub-partition (1.0 = one instruction issued per cycle).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

template <int KIND>
__global__ void k(float* out, const float* in, long long* cycles) {
  float a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)], c = in[64 + (threadIdx.x & 31)];
  u64 r[8];
  float s8[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { r[i] = pk(in[i] + threadIdx.x, in[i + 8] - threadIdx.x); s8[i] = in[i + 16] + threadIdx.x; }
  const u64 ab = pk(a, b), bb = pk(b, b), cc = pk(c, a);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) r[i] = ffma2(r[i], ab, cc);                       // FFMA2, all packed operands
      if (KIND == 1) r[i] = ffma2(r[i], bb, cc);                       // FFMA2, broadcast scalar operand
      if (KIND == 2) r[i] = fmul2(r[i], ab);                           // FMUL2
      if (KIND == 3) r[i] = fadd2(r[i], ab);                           // FADD2
      if (KIND == 4) { r[i] = ffma2(r[i], bb, cc); s8[i] = fminf(fminf(s8[i], a), b); }                 // FFMA2 + FMNMX3 (2 ops)
      if (KIND == 5) { r[i] = ffma2(r[i], bb, cc); s8[i] = fmaf(s8[i], a, b); }                          // FFMA2 + FFMA (2 ops)
      if (KIND == 6) { r[i] = ffma2(r[i], bb, cc); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(s8[i])); }   // FFMA2 + MUFU (2 ops)
      if (KIND == 7) { r[i] = ffma2(r[i], bb, cc); r[i] = fmul2(r[i], ab); s8[i] = fminf(fminf(s8[i], a), b); }  // 2 packed + FMNMX3 (3 ops)
      if (KIND == 9) {   // packed LoS mix per cell pair and wall: 2 FFMA2 + FMUL2 + 2 FMNMX3 (5 ops)
        const u64 X = ffma2(r[i], bb, cc), Y = ffma2(r[i], ab, cc), Q = fmul2(r[i], bb);
        float x0, x1, y0, y1, q0, q1; upk(X, x0, x1); upk(Y, y0, y1); upk(Q, q0, q1);
        s8[i] = fmaxf(s8[i], fminf(fminf(x0, y0), q0)); s8[i] = fmaxf(s8[i], fminf(fminf(x1, y1), q1));
        r[i] = fadd2(r[i], ab);
      }
      if (KIND == 10) {  // the same in scalar form: 4 FFMA + 2 FMUL + 2 FMNMX3 (+ the chain FADDs)
        float r0, r1; upk(r[i], r0, r1);
        const float x0 = fmaf(r0, b, c), x1 = fmaf(r1, b, a), y0 = fmaf(r0, a, c), y1 = fmaf(r1, b, a), q0 = r0 * b, q1 = r1 * b;
        s8[i] = fmaxf(s8[i], fminf(fminf(x0, y0), q0)); s8[i] = fmaxf(s8[i], fminf(fminf(x1, y1), q1));
        r[i] = pk(r0 + a, r1 + b);
      }
      if (KIND == 8) { float x, y; upk(r[i], x, y); s8[i] = fminf(fminf(x, y), s8[i]); r[i] = ffma2(r[i], bb, cc); } // dependent unpack
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float x, y; upk(r[i], x, y); s += x + y + s8[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

// global atomics (RED.OR) sparse: one lane in 32 issues a red per iteration
__global__ void k_red(unsigned* words, long long* cycles, int n_words, int iters) {
  const long long t0 = clock64();
  unsigned h = blockIdx.x * 9781u + threadIdx.x * 6271u;
  for (int it = 0; it < iters; ++it) {
    h = h * 1664525u + 1013904223u;
    if ((h >> 27) == 0u) atomicOr(words + (h % (unsigned)n_words), 1u << (threadIdx.x & 31));
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int KIND>
void run(const char* name, int warps_per_smsp, int ops_per_iter) {
  float *out, *in;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&in, 1024 * 4);
  cudaMalloc(&cyc, 8);
  float h[1024];
  for (int i = 0; i < 1024; ++i) h[i] = 0.5f + 0.001f * i;
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  const int threads = warps_per_smsp * 4 * 32;
  k<KIND><<<148, threads>>>(out, in, cyc);
  k<KIND><<<148, threads>>>(out, in, cyc);
  cudaDeviceSynchronize();
  long long c;
  cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / ((double)ITER * 8 * ops_per_iter * warps_per_smsp);
  printf("%-36s warps/smsp=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, per);
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4}) {
    run<0>("FFMA2 packed", w, 1);
    run<1>("FFMA2 broadcast", w, 1);
    run<2>("FMUL2", w, 1);
    run<3>("FADD2", w, 1);
    run<4>("FFMA2 + FMNMX3 (2 ops)", w, 2);
    run<5>("FFMA2 + FFMA (2 ops)", w, 2);
    run<6>("FFMA2 + MUFU.EX2 (2 ops)", w, 2);
    run<7>("FFMA2 + FMUL2 + FMNMX3 (3 ops)", w, 3);
    run<8>("FMNMX3(unpack) + FFMA2 (2 ops)", w, 2);
    run<9>("LoS mix packed (per cell pair+wall)", w, 1);
    run<10>("LoS mix scalar (per cell pair+wall)", w, 1);
  }
  unsigned* words; long long* cyc;
  const int n_words = 2 * 1024 * 1024;    // 8 MB of spike words
  cudaMalloc(&words, n_words * 4); cudaMalloc(&cyc, 8);
  cudaMemset(words, 0, n_words * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_red<<<148 * 4, 512>>>(words, cyc, n_words, 64);
  cudaEventRecord(e0);
  k_red<<<148 * 4, 512>>>(words, cyc, n_words, 64);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double reds = 148.0 * 4 * 512 * 64 / 32;
  printf("sparse RED.OR: %.0f reds in %.1f us  (%.2f G reds/s)\n", reds, ms * 1e3, reds / ms * 1e-6);
  return 0;
}
