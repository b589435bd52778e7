// Shared device helpers: non-contracting float64 wrapper, Philox4x32-10, TMA bulk
// copy + mbarrier primitives (sm_100a), cache-hinted stores.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/riab_b200.h"

#define RIAB_DEV __device__ __forceinline__
#define RIAB_HD __host__ __device__ __forceinline__

namespace riab {

// ---------------------------------------------------------------------------
// D: a float64 whose + - * / are the IEEE round-to-nearest operations and are
// never contracted into FMAs, so that expressions written in the order NumPy
// evaluates them give bit-identical results (the geometry predicates that
// decide wall collisions depend on this).
struct D {
  double v;
  RIAB_DEV D() {}
  RIAB_DEV D(double x) : v(x) {}
};
RIAB_DEV D operator+(D a, D b) { return D(__dadd_rn(a.v, b.v)); }
RIAB_DEV D operator-(D a, D b) { return D(__dsub_rn(a.v, b.v)); }
RIAB_DEV D operator*(D a, D b) { return D(__dmul_rn(a.v, b.v)); }
RIAB_DEV D operator/(D a, D b) { return D(__ddiv_rn(a.v, b.v)); }
RIAB_DEV D operator-(D a) { return D(-a.v); }
RIAB_DEV bool operator<(D a, D b) { return a.v < b.v; }
RIAB_DEV bool operator>(D a, D b) { return a.v > b.v; }
RIAB_DEV bool operator<=(D a, D b) { return a.v <= b.v; }
RIAB_DEV bool operator>=(D a, D b) { return a.v >= b.v; }
RIAB_DEV bool operator==(D a, D b) { return a.v == b.v; }
RIAB_DEV D dsqrt(D a) { return D(__dsqrt_rn(a.v)); }

// ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11).  Known-answer vectors are checked in
// tests/test_philox.py against a NumPy implementation of the same rounds.
RIAB_HD void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
RIAB_HD void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int i = 0; i < 10; ++i) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

// Philox4x32-R with the round keys precomputed (warp-uniform): 2 wide multiplies + 2 three-input
// XORs per round.
template <int R>
RIAB_DEV void philox_keyed(uint32_t (&c)[4], const uint32_t (&rk)[2 * R]) {
#pragma unroll
  for (int i = 0; i < R; ++i) {
    unsigned long long p0, p1;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p0) : "r"(c[0]), "r"(0xD2511F53u));
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(p1) : "r"(c[2]), "r"(0xCD9E8D57u));
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ rk[2 * i], n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ rk[2 * i + 1];
    c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
  }
}
template <int R>
RIAB_DEV void philox_round_keys(uint32_t (&rk)[2 * R], unsigned long long seed) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int i = 0; i < R; ++i) { rk[2 * i] = k0; rk[2 * i + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
}

// Stream ids (third counter word, top byte)
enum : uint32_t { RIAB_STREAM_AGENT_OU = 0, RIAB_STREAM_CELL_NOISE = 1, RIAB_STREAM_SPIKES = 2, RIAB_STREAM_MEASURE = 3,
                  RIAB_STREAM_THIN = 4 };          // thinned spikes: calls n = 0, 1, ... of a (row, 128-cell block), n in bits 16.. of the sub-index

// counter = (agent id lo32, sub-index, step lo32, (step hi & 0xffff) | stream<<24 | population<<16)
RIAB_HD void philox_ctr(uint32_t (&c)[4], uint64_t agent, uint32_t sub, uint64_t step, uint32_t stream, uint32_t pop) {
  c[0] = (uint32_t)agent;
  c[1] = sub ^ ((uint32_t)(agent >> 32) << 24);
  c[2] = (uint32_t)step;
  c[3] = ((uint32_t)(step >> 32) & 0xffffu) | ((pop & 0xffu) << 16) | (stream << 24);
}

// two uint32 -> uniform double in (0,1) with 53 random bits
RIAB_HD double u01_53(uint32_t hi, uint32_t lo) {
  const uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}
RIAB_HD float u01_24(uint32_t x) { return ((float)(x >> 8) + 0.5f) * (1.0f / 16777216.0f); }

// Two standard normals for the agent OU draws: Philox4x32-10 -> two 32-bit uniforms ->
// Box-Muller in float32 (logf / sqrtf / sincospif are the accurate single-precision routines),
// widened to double.  The draws are noise: their float32 resolution is irrelevant to the
// dynamics, and float32 keeps ~150 dependent float64 operations off the motion chain.
RIAB_DEV void agent_normals(uint64_t seed, uint64_t step, uint64_t agent, double& n1, double& n2) {
  uint32_t c[4];
  philox_ctr(c, agent, 0u, step, RIAB_STREAM_AGENT_OU, 0u);
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float u1 = fmaf(__uint2float_rn(c[0]), 2.3283064365386963e-10f, 1.1641532182693481e-10f);   // (0,1]
  const float u2 = __uint2float_rn(c[2]) * 2.3283064365386963e-10f;                                  // [0,1]
  const float r = sqrtf(-2.0f * logf(u1));
  float s, co;
  sincospif(2.0f * u2, &s, &co);
  n1 = (double)(r * co); n2 = (double)(r * s);
}

// ---------------------------------------------------------------------------
// mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS: UBLKCP) helpers.
RIAB_DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
RIAB_DEV void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
RIAB_DEV void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
RIAB_DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
RIAB_DEV void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
RIAB_DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// bytes must be a multiple of 16; src/dst 16-byte aligned.
RIAB_DEV void tma_bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// CPT consecutive floats of a packed array (16- / 8-byte vector load)
template <int CPT>
RIAB_DEV void ldv(float (&d)[CPT], const float* __restrict__ p) {
  if constexpr (CPT == 4) { const float4 v = *reinterpret_cast<const float4*>(p); d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
  else { const float2 v = *reinterpret_cast<const float2*>(p); d[0] = v.x; d[1] = v.y; }
}

// Streaming (evict-first) vector stores for the write-once rate rows.
RIAB_DEV void st_cs_f4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
RIAB_DEV void st_cs_f2(float* p, float a, float b) { asm volatile("st.global.cs.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory"); }
template <int CPT>
RIAB_DEV void st_cs_fv(float* p, const float (&o)[CPT]) {
  if constexpr (CPT == 4) st_cs_f4(p, o[0], o[1], o[2], o[3]);
  else st_cs_f2(p, o[0], o[1]);
}
RIAB_DEV void st_cs_f1(float* p, float a) { asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory"); }

RIAB_DEV float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Packed float32 pairs (PTX f32x2, SASS FFMA2 / FMUL2 / FADD2, sm_100): one issue slot for two FMA-pipe operations on an
// even-aligned register pair.  `bc2(x)` is the broadcast pair (x, x): ptxas folds it into the instruction's scalar
// `.F32` operand form, so per-agent shared-memory broadcasts need no duplicated record fields and no MOVs.
// scripts/ubench_packed.cu: 2.0 cycles per packed instruction per sub-partition = the FMA-pipe time of two scalar
// operations in one issue slot (the rate consumers are issue-bound, not FMA-pipe bound).
typedef unsigned long long f32x2;
RIAB_DEV f32x2 pk2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
RIAB_DEV f32x2 bc2(float x) { return pk2(x, x); }
RIAB_DEV void upk2(f32x2 r, float& lo, float& hi) { asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r)); }
#ifndef RIAB_SCALAR_PAIRS
RIAB_DEV f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
RIAB_DEV f32x2 fmul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
RIAB_DEV f32x2 fadd2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#else   // experiment: the same arithmetic as two scalar FMA-pipe instructions per pair
RIAB_DEV f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { float a0, a1, b0, b1, c0, c1; upk2(a, a0, a1); upk2(b, b0, b1); upk2(c, c0, c1); return pk2(fmaf(a0, b0, c0), fmaf(a1, b1, c1)); }
RIAB_DEV f32x2 fmul2(f32x2 a, f32x2 b) { float a0, a1, b0, b1; upk2(a, a0, a1); upk2(b, b0, b1); return pk2(a0 * b0, a1 * b1); }
RIAB_DEV f32x2 fadd2(f32x2 a, f32x2 b) { float a0, a1, b0, b1; upk2(a, a0, a1); upk2(b, b0, b1); return pk2(a0 + b0, a1 + b1); }
#endif

// 2^x for x <= 0 on the FMA / ALU pipes (no MUFU): round-to-nearest split x = n + f through the 1.5*2^23 trick,
// degree-5 polynomial for 2^f on [-0.5, 0.5] (max relative error 2.5e-7, ex2.approx's own is ~2e-7), exponent
// added in the integer domain.  11 instructions; used next to MUFU.EX2 where the MUFU pipe is the limit.
RIAB_DEV float ex2_fma(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float f = x - (t - 12582912.0f);
  float p = fmaf(1.3400433e-3f, f, 9.6760374e-3f);
  p = fmaf(p, f, 5.5503272e-2f);
  p = fmaf(p, f, 2.4022107e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0000001f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// NumPy's floating remainder (result takes the sign of the divisor) -- np.mod
RIAB_DEV double np_mod(double x, double m) {
  double r = fmod(x, m);
  if (r != 0.0 && ((r < 0.0) != (m < 0.0))) r += m;
  return r;
}

}  // namespace riab
