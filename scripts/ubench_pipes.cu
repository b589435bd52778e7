// Issue-rate microbenchmarks for the instruction mix of the rate consumers (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_pipes.bin scripts/ubench_pipes.cu
// Each test: WARPS warps per SM sub-partition x ITER iterations of 8 independent chains of one
// instruction kind; prints cycles per warp-instruction per sub-partition (1.0 = full issue rate).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 4096;

template <int KIND>
__global__ void k(float* out, const float* in, long long* cycles) {
  float r[8], a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)], c = in[64 + (threadIdx.x & 31)];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = in[i] + threadIdx.x;
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = __float_as_uint(r[i]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) r[i] = fmaf(r[i], a, b);                       // FFMA, 3 register sources
      if (KIND == 1) r[i] = fmaf(r[i], 1.0009765625f, b);           // FFMA, immediate
      if (KIND == 2) r[i] = r[i] * a;                               // FMUL 2 regs
      if (KIND == 3) r[i] = r[i] + a;                               // FADD
      if (KIND == 4) r[i] = fminf(fminf(r[i], a), b);               // FMNMX3
      if (KIND == 5) r[i] = fminf(r[i], a);                         // FMNMX
      if (KIND == 6) r[i] = __saturatef(r[i] * a);                  // FMUL.SAT
      if (KIND == 7) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r[i])); }      // MUFU
      if (KIND == 8) { r[i] = (float)(u[i] & 0xffffu); u[i] += 3; } // I2F.U16 (+IADD)
      if (KIND == 9) { u[i] = __ballot_sync(0xffffffffu, r[i] < a); r[i] += 1.f; }     // FSETP + VOTE + FADD
      if (KIND == 10) { unsigned long long p; asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(u[i]), "r"(0xD2511F53u)); u[i] = (unsigned)(p >> 32) ^ (unsigned)p; }  // IMAD.WIDE + LOP3
      if (KIND == 11) r[i] = (r[i] < a) ? b : r[i];                 // FSETP + FSEL
      if (KIND == 12) r[i] = fmaf(fabsf(a), b, r[i]);               // FFMA acc in place (2 uniform-ish regs)
      if (KIND == 13) r[i] = fmaf(r[i], a, c);                      // FFMA distinct c
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += r[i] + __uint_as_float(u[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
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
  printf("%-28s warps/smsp=%d  cycles per warp-instr per SMSP = %.3f\n", name, warps_per_smsp, per);
  cudaFree(out); cudaFree(in); cudaFree(cyc);
}

int main() {
  for (int w : {1, 4}) {
    run<0>("FFMA 3-reg", w, 1);
    run<13>("FFMA 3-reg distinct", w, 1);
    run<1>("FFMA imm", w, 1);
    run<12>("FFMA acc (|a|*b+r)", w, 1);
    run<2>("FMUL", w, 1);
    run<3>("FADD", w, 1);
    run<4>("FMNMX3", w, 1);
    run<5>("FMNMX", w, 1);
    run<6>("FMUL.SAT", w, 1);
    run<7>("MUFU.EX2", w, 1);
    run<8>("I2F.U16 + IADD (2 ops)", w, 2);
    run<9>("FSETP+VOTE+FADD (3 ops)", w, 3);
    run<10>("IMAD.WIDE+LOP3 (2 ops)", w, 2);
    run<11>("FSETP+FSEL (2 ops)", w, 2);
  }
  return 0;
}
