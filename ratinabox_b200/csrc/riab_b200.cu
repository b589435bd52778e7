// libriab_b200.so -- kernels + C ABI (include/riab_b200.h).  sm_100a only.
//
// Kernel inventory
//   k_agent_update      Agent.update, one agent per thread (float64)
//   k_step<P,MODE,..>   persistent warp-specialised step kernel for PlaceCells / GridCells:
//                       producer warps run Agent.update (float64) and publish per-agent float32
//                       records through an mbarrier ring; consumer warps keep 4 cells per thread
//                       in registers and stream float4 rate rows (+ OU noise, + bit-packed spikes)
//   k_bvc_rays<FUSED>   BVC phase A (float64 rays) [+ Agent.update]
//   k_bvc_integrate     BVC phase B (float32 angular integral, TMA-staged tables)
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "riab_bvc.cuh"
#include "riab_ovc.cuh"
#include "riab_grid.cuh"
#include "riab_motion.cuh"
#include "riab_place.cuh"

using namespace riab;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define RIAB_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) return fail(RIAB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

constexpr int TA = 32;      // agents per tile (one warp runs their motion)
constexpr int NT = 256;     // threads per CTA in the tile kernels
constexpr int MAXW = 64;    // walls staged in shared memory
constexpr int CELL_PAD = 128;  // packed per-cell arrays are padded to 4 cells x 32 lanes
#ifndef RIAB_BVC_MUFU_TERMS
#define RIAB_BVC_MUFU_TERMS 8
#endif
constexpr int BVC_MUFU_TERMS = RIAB_BVC_MUFU_TERMS;   // of 8 agents per thread: exponentials on the MUFU pipe (rest: ex2_fma)

struct EnvK {
  const double* walls;
  int W, nb, aligned, periodic;
  int polygon, nh, h0;      // RIAB_BOUNDARY_POLYGON: even-odd in-environment test over walls [0,nb) and [h0,h0+nh)
  double ext[4];
  double cxm, cym, scale;
};

struct OutK {
  float* rates;
  long long ld;
  uint32_t* spikes;
  long long spike_ld;       // words per row
  float* noise;
  float dt, noise_decay, noise_sig;   // n <- n + (-n*dt/tau) + sig*xi ; decay = dt/tau
  unsigned long long seed, step;
  long long id_offset;
  int pop, vec_ok;
  uint32_t rk7[14];         // Philox4x32-7 round keys of the spike stream (host-computed, constant bank)
  // thinned spikes (thin_block): candidates at rate p' = dt * (an upper bound of the rate), accepted with rate / bound
  int tile_agents;          // agents per ring slot of k_step (<= TA, even): launch_tile shrinks the tiles of small batches so
                            // that every (CTA, consumer group) gets an equal share (strong scaling: 8 192 agents per GPU
                            // are 256 tiles of 32 on 148 x 2 groups -- 1.7 waves -- but 293 tiles of 28)
  int thin;                 // 1: the population's rates are bounded and p' = dt * bound <= 1/16 (thin_block)
  uint32_t thin_cdf[32];    // cdf[k] = floor(2^32 P(Binomial(128, p') <= k)): candidates of a (row, 128-cell block) = #{k: word >= cdf[k]}
  float thin_c1, thin_c0;   // accept <=> fma(float(20-bit uniform), c1, c0) < rate;  c1 = 2^-20 bound, c0 = 2^-21 bound
};

// MODE 3 of k_step: the whole riab_run loop of a single Place / Grid population in ONE launch.  Agents are independent and
// the tile -> CTA assignment is static, so a CTA can run all n_steps for its own agents with no grid-wide synchronisation:
// the producers keep advancing their tiles into the next step while the consumers still write the rates of this one, and
// the per-launch ramp (wall staging, cell registers, first records) and tail are paid once per run instead of once per step.
struct RunK {
  long long n_steps;
  float* rates_ring;            // (ring_rows, A, ld): step s writes row (ring_next + s) % ring_rows
  uint32_t* spikes_ring;        // (ring_rows, A, spike_ld) or NULL
  long long ring_rows, ring_next;
  float* hist_ring;             // (hist_rows, A, 8) agent history rows or NULL
  long long hist_rows, hist_next;
};

// ---------------------------------------------------------------------------
// walls -> shared memory through a 1-D TMA bulk copy when 16-byte aligned.
__device__ __forceinline__ void stage_walls(double* s_walls, uint64_t* bar, const EnvK& env) {
  const uint32_t bytes = (uint32_t)env.W * 32u;
  if (env.aligned) {
    if (threadIdx.x == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      mbar_expect_tx(bar, bytes);
      tma_bulk_g2s(s_walls, env.walls, bytes, bar);
    }
    __syncthreads();
    mbar_wait(bar, 0);
  } else {
    for (int i = threadIdx.x; i < env.W * 4; i += blockDim.x) s_walls[i] = env.walls[i];
    __syncthreads();
  }
}

// One agent's Agent.update inside a kernel (loads/stores its state).
template <bool REC>
__device__ __forceinline__ void agent_update_one(const riab_agents& ag, const riab_motion_params& mp,
                                                 const MotionDerived& md, const riab_step_io& io, const EnvK& env,
                                                 const double* s_walls, long long i, AgentState& s) {
  load_agent(ag, i, s);
  double n1, n2;
  const unsigned long long gid = (unsigned long long)(ag.id_offset + i);
  if (io.xi != nullptr) { n1 = io.xi[2 * i]; n2 = io.xi[2 * i + 1]; }
  else agent_normals(io.seed, io.step, gid, n1, n2);
  const bool has_drift = io.drift_velocity != nullptr;
  double drx = 0.0, dry = 0.0;
  if (has_drift) { drx = io.drift_velocity[2 * i]; dry = io.drift_velocity[2 * i + 1]; }
  // exactly-zero displacement (Agent.py:460-461) draws from a separate Philox stream, lazily
  const double f1 = __longlong_as_double((long long)io.seed), f2 = __longlong_as_double((long long)(io.step ^ (gid << 20)));
  uint8_t* mask = (REC && io.collision_mask) ? io.collision_mask + (size_t)i * RIAB_MAX_REC_ITERS * env.W : nullptr;
  int32_t* fh = (REC && io.first_hit) ? io.first_hit + (size_t)i * RIAB_MAX_REC_ITERS : nullptr;
  int32_t* ni = (REC && io.n_iters) ? io.n_iters + i : nullptr;
  motion_step<REC>(s, s_walls, env.W, mp, md, env.ext, env.periodic != 0, env.scale, env.polygon != 0, env.nb, env.h0, env.nh,
                   n1, n2, has_drift, drx, dry, f1, f2,
                   mask, fh, ni);
  store_agent(ag, i, s);
  if (io.pos_mirror != nullptr) *reinterpret_cast<double2*>(io.pos_mirror + 2 * (size_t)i) = make_double2(s.px, s.py);
  if (io.history_row != nullptr) store_history_row(io.history_row + 8 * (size_t)i, s);
}

template <bool REC>
__global__ void __launch_bounds__(128) k_agent_update(const riab_agents ag, const riab_motion_params mp,
                                                      const MotionDerived md, const riab_step_io io, const EnvK env) {
  __shared__ __align__(16) double s_walls[MAXW * 4];
  __shared__ uint64_t s_bar;
  stage_walls(s_walls, &s_bar, env);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ag.n_agents) return;
  AgentState s;
  agent_update_one<REC>(ag, mp, md, io, env, s_walls, i, s);
}

// ---------------------------------------------------------------------------
// Neurons.update tail for 4 consecutive cells of one agent: OU noise
// (Neurons.py:153-160,168), rate store, spikes (Neurons.py:681-684).
// `full4`: all 4 cells exist and the row is 16-byte aligned (vector store).
// Per-thread constants of the rate tail (4 consecutive cells).
struct TailCtx {
  int cell0, n_cells;
  unsigned vmask;          // which of the 4 cells exist
  bool full4;              // all 4 exist and rows are 16-byte aligned -> one vector store
  uint32_t sub;            // Philox counter word 1: cell group index
  uint32_t c2, c3_spk;     // Philox counter words 2,3 (step, stream | population)
};

// Row cursor of one thread: pointers into the current agent's rows, advanced per agent.
struct RowCursor {
  float* dst;              // rates row + cell0
  float* nz;               // noise-state row + cell0 (or NULL)
  uint32_t* spk;           // spikes row + 4 * (cell0 / 128): the 4 ballot words of this warp's 128 cells
  unsigned long long gid;  // global agent id
};

template <int CPT = 4>
__device__ __forceinline__ void tail_init(TailCtx& t, const OutK& out, int cell0, int n_cells, unsigned long long step);
template <int CPT = 4>
__device__ __forceinline__ void tail_init(TailCtx& t, const OutK& out, int cell0, int n_cells) {
  tail_init<CPT>(t, out, cell0, n_cells, out.step);
}
template <int CPT>
__device__ __forceinline__ void tail_init(TailCtx& t, const OutK& out, int cell0, int n_cells, unsigned long long step) {
  t.cell0 = cell0; t.n_cells = n_cells;
  t.vmask = 0u;
#pragma unroll
  for (int i = 0; i < CPT; ++i) t.vmask |= (cell0 + i < n_cells) ? (1u << i) : 0u;
  t.full4 = out.vec_ok && (t.vmask == ((1u << CPT) - 1u));
  t.sub = (uint32_t)(cell0 >> 2);
  t.c2 = (uint32_t)step;
  const uint32_t hi = ((uint32_t)(step >> 32) & 0xffffu) | (((uint32_t)out.pop & 0xffu) << 16);
  t.c3_spk = hi | (RIAB_STREAM_SPIKES << 24);
}

__device__ __forceinline__ void cursor_init(RowCursor& rc, const OutK& out, const TailCtx& t, long long row) {
  rc.dst = out.rates + row * out.ld + t.cell0;
  rc.nz = out.noise ? out.noise + row * out.ld + t.cell0 : nullptr;
  rc.spk = out.spikes ? out.spikes + row * out.spike_ld + ((t.cell0 >> 7) << 2) : nullptr;
  rc.gid = (unsigned long long)(out.id_offset + row);
}
struct RowStride { long long rate, spk; int rows; };      // element strides for `rows` agents (warp-uniform)
__device__ __forceinline__ RowStride make_stride(const OutK& out, int rows) {
  RowStride st;
  st.rate = (long long)rows * out.ld; st.spk = (long long)rows * out.spike_ld; st.rows = rows;
  return st;
}
__device__ __forceinline__ void cursor_advance(RowCursor& rc, const RowStride& st) {
  rc.dst += st.rate;
  if (rc.nz) rc.nz += st.rate;
  if (rc.spk) rc.spk += st.spk;
  rc.gid += (unsigned long long)st.rows;
}

// Neurons.update tail for 4 consecutive cells of one agent, part 1: OU noise
// (Neurons.py:153-160,168) and the rate store.
template <bool NOISE>
__device__ __forceinline__ void store4(float (&o)[4], const OutK& out, const TailCtx& t, const RowCursor& rc,
                                       long long row_off /* extra rows, in elements of ld */) {
  if (NOISE && rc.nz != nullptr) {
    uint32_t c[4];
    philox_ctr(c, rc.gid + (row_off != 0 ? 1ull : 0ull), t.sub, out.step, RIAB_STREAM_CELL_NOISE, (uint32_t)out.pop);
    philox4x32_10(c, (uint32_t)out.seed, (uint32_t)(out.seed >> 32));
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float u1 = u01_24(c[2 * h]), u2 = u01_24(c[2 * h + 1]);
      const float r = sqrtf(-2.0f * __logf(u1));
      float sn, cs;
      __sincosf(6.2831853071795865f * u2, &sn, &cs);
      z[2 * h] = r * cs; z[2 * h + 1] = r * sn;
    }
    float* nzp = rc.nz + row_off;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if ((t.vmask >> i) & 1u) {
        float n = nzp[i];
        n = n + (-n * out.noise_decay) + out.noise_sig * z[i];
        nzp[i] = n;
        o[i] += n;
      }
    }
  }
  float* dst = rc.dst + row_off;
  if (t.full4) {
    st_cs_f4(dst, o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if ((t.vmask >> i) & 1u) st_cs_f1(dst + i, o[i]);
  }
}

// Part 2: spikes (Neurons.py:681-684), spike <=> uniform < dt * rate.
// One Philox4x32-7 call serves the 4 cells of TWO agents (global ids 2k, 2k+1):
//   r = Philox7(ctr = (gid >> 1, cell group, step, stream|population), key = seed)
//   agent half h = gid & 1 takes words r[2h], r[2h+1]: four 16-bit integers m_i (cell i = half-word i);
//   all eight share the dither v = ((r0^r2) >> 8) * 2^-24; uniform_i = (m_i + v) * 2^-16, and
//   spike_i <=> m_i < fma(rate_i, dt*65536, -v)      (one FFMA, one I2F, one FSETP per rate).
// (m_i + v) is uniform on [0, 65536) with 40 random bits and v is independent of every single m_i;
// sharing v only correlates the sub-2^-16 fractions of the eight uniforms (covariance < 6e-11).
// Layout of a spike row: a warp owns 128 consecutive cells (lane L holds cells 128B+4L .. +3) and
// writes the four ballots of its cell slots as one 16-byte store:
//   bit L of word 4B+i  =  spike of cell 128B + 4L + i.
__device__ __forceinline__ void spike_words(uint32_t (&c)[4], const OutK& out, const TailCtx& t, unsigned long long gid) {
  const unsigned long long pair = gid >> 1;
  c[0] = (uint32_t)pair; c[1] = t.sub ^ ((uint32_t)(pair >> 32) << 24); c[2] = t.c2; c[3] = t.c3_spk;
  philox_keyed<7>(c, out.rk7);
}
__device__ __forceinline__ float spike_neg_dither(const uint32_t (&c)[4]) {
  return (float)((c[0] ^ c[2]) >> 8) * -5.9604644775390625e-08f;
}
// ballots of one agent's 4 cell slots; `ok` = this thread's cells exist (all 4 or none) when !MASKED.
// XU_BOUND: for a cell type that saturates the XU pipe the four integer -> float conversions can be done as
// as_float(0x4B000000 | m) - 2^23  (PRMT + FADD, exact, same bits) instead of I2F; no current policy needs it.
template <bool MASKED, bool XU_BOUND = false>
__device__ __forceinline__ void spike_ballots(uint32_t (&b)[4], uint32_t w0, uint32_t w1, float nv, const float (&o)[4],
                                              float q /* dt * 65536 */, unsigned vmask, bool ok) {
  float m0, m1, m2, m3;           // the four 16-bit integers as floats
  if (XU_BOUND) {
    m0 = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7610)) - 8388608.0f;
    m1 = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7632)) - 8388608.0f;
    m2 = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7610)) - 8388608.0f;
    m3 = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7632)) - 8388608.0f;
  } else {                        // I2F.U16 reads either half-word directly
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f32.u16 %0, l;\n\tcvt.rn.f32.u16 %1, h;\n\t}" : "=f"(m0), "=f"(m1) : "r"(w0));
    asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f32.u16 %0, l;\n\tcvt.rn.f32.u16 %1, h;\n\t}" : "=f"(m2), "=f"(m3) : "r"(w1));
  }
  bool s0 = m0 < fmaf(o[0], q, nv);
  bool s1 = m1 < fmaf(o[1], q, nv);
  bool s2 = m2 < fmaf(o[2], q, nv);
  bool s3 = m3 < fmaf(o[3], q, nv);
  if (MASKED) { s0 = s0 && (vmask & 1u); s1 = s1 && (vmask & 2u); s2 = s2 && (vmask & 4u); s3 = s3 && (vmask & 8u); }
  else { s0 = s0 && ok; s1 = s1 && ok; s2 = s2 && ok; s3 = s3 && ok; }
  b[0] = __ballot_sync(0xffffffffu, s0);
  b[1] = __ballot_sync(0xffffffffu, s1);
  b[2] = __ballot_sync(0xffffffffu, s2);
  b[3] = __ballot_sync(0xffffffffu, s3);
}
__device__ __forceinline__ void spike_store(const uint32_t (&b)[4], uint32_t* spk) {
  // every lane holds the same four words: one elected lane stores them (whole warp must call)
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "@p st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};\n\t}" ::"l"(spk), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3])
      : "memory");
}
// one agent (whole warp must call)
__device__ __forceinline__ void spikes1(const float (&o)[4], const OutK& out, const TailCtx& t, const RowCursor& rc) {
  uint32_t c[4], b[4];
  spike_words(c, out, t, rc.gid);
  const float nv = spike_neg_dither(c);
  const unsigned h = (unsigned)(rc.gid & 1ull);
  spike_ballots<true>(b, h ? c[2] : c[0], h ? c[3] : c[1], nv, o, out.dt * 65536.0f, t.vmask, true);
  spike_store(b, rc.spk);
}

template <bool SPIKES, bool NOISE>
__device__ __forceinline__ void finish4(float (&o)[4], const OutK& out, const TailCtx& t, const RowCursor& rc) {
  store4<NOISE>(o, out, t, rc, 0);
  if (SPIKES && (!NOISE || rc.spk != nullptr)) spikes1(o, out, t, rc);
}

// ---------------------------------------------------------------------------
// Cell-type policies for the step kernel.
template <int WI, int DESC, int CPT_ = 4>
struct PlacePolicy {
  using Const = PlaceConst;
  using Regs = PlaceCellRegs<WI, CPT_>;
  static constexpr int CPT = CPT_;                          // cells per consumer thread
  static constexpr int REC = place_rec(WI);
  static constexpr bool LIGHT = (WI == 0) && (DESC >= 0);   // few instructions per rate: HBM-bound consumers
  static constexpr bool XU_BOUND = false;
  // rates lie in [min_fr, max_fr], so the thinned spike stream applies, but it measured slower for place cells: the
  // Euclidean Gaussian loop is HBM-bound and hides the dense stream's instructions under its stores (c2e 56-60 us dense,
  // 75 us thinned), the line-of-sight loop with the post-pass needs 8 producer warps and loses next to them (c2 105 us
  // dense, 109-120 us thinned; gpurun_out/r02[b-f]_*).  The dense stream stays.
  static constexpr bool THIN = false;
  static __device__ __forceinline__ const double* head_dir(const Const&) { return nullptr; }
  static __device__ __forceinline__ void prepare(double* aux, const double* s_walls, const Const& c) {
    place_wall_invariants(aux, s_walls + 4 * c.wall0, WI > 0 ? c.n_inner : 0);
  }
  static __device__ __forceinline__ void record(float* rec, double px, double py, double, double, const double* s_walls,
                                                const double* aux, const Const& c, const EnvK& env) {
    place_agent_record<WI>(rec, px, py, s_walls + 4 * c.wall0, aux, WI > 0 ? c.n_inner : 0, c.geometry, env.cxm, env.cym, c.band, c.expanded, c.kx, c.fold ? c.lspan : 0.f);
  }
  static __device__ __forceinline__ void load(Regs& r, const Const& c, int cell0) { place_load_cells<WI, CPT_>(r, c, cell0); }
  template <bool DEFER, int EXP = -1>
  static __device__ __forceinline__ void rates4(float (&o)[CPT_], const Regs& r, const Const& c, int cell0,
                                                const float* rec, uint32_t inner_s, bool& unsure) {
    place_rates4<WI, DESC, DEFER, EXP, CPT_>(o, r, c, cell0, rec, inner_s, unsure);
  }
  // 0: direct form, 1: expanded exponent, 2: expanded with the [0, max_fr] scale folded into the exponent
  static __device__ __forceinline__ int expanded(const Const& c) { return (DESC == RIAB_PC_GAUSSIAN && c.expanded) ? 1 + c.fold : 0; }
  static __device__ __forceinline__ int wall0(const Const& c) { return c.wall0; }
};

template <int CPT_ = 4>
struct GridPolicy {
  using Const = GridConst;
  using Regs = GridCellRegs<CPT_>;
  static constexpr int CPT = CPT_;
  static constexpr int REC = 4;
  static constexpr bool LIGHT = false;    // 36 cell registers per thread do not fit StepCfg<8>'s 56-register consumers
  static constexpr bool THIN = true;      // bounded rates, consumer-bound loop: thinned spikes (c3 93.5 -> 80 us)
  static constexpr bool XU_BOUND = false; // 3 MUFU.COS per rate, yet issue-bound: PRMT+FADD instead of I2F measured slower (95.6 vs 91.3 us, c3)
  static __device__ __forceinline__ const double* head_dir(const Const&) { return nullptr; }
  static __device__ __forceinline__ void prepare(double*, const double*, const Const&) {}
  static __device__ __forceinline__ void record(float* rec, double px, double py, double, double, const double*, const double*,
                                                const Const&, const EnvK& env) {
    rec[0] = (float)(px - env.cxm);
    rec[1] = (float)(py - env.cym);
  }
  static __device__ __forceinline__ void load(Regs& r, const Const& c, int cell0) { grid_load_cells<CPT_>(r, c, cell0); }
  template <bool DEFER, int EXP = -1>
  static __device__ __forceinline__ void rates4(float (&o)[CPT_], const Regs& r, const Const& c, int, const float* rec,
                                                uint32_t, bool&) {
    grid_rates4<CPT_>(o, r, c, rec);
  }
  static __device__ __forceinline__ int expanded(const Const&) { return 0; }
  static __device__ __forceinline__ int wall0(const Const&) { return 0; }
};

struct OvcPolicy {
  using Const = OvcConst;
  using Regs = OvcCellRegs;
  static constexpr int CPT = 4;
  static constexpr int REC = OVC_REC;
  static constexpr bool LIGHT = false;
  static constexpr bool THIN = false;     // sums over objects: no a-priori rate bound
  static constexpr bool XU_BOUND = false;
  static __device__ __forceinline__ const double* head_dir(const Const& c) { return c.head_dir; }
  static __device__ __forceinline__ void prepare(double*, const double*, const Const&) {}
  static __device__ __forceinline__ void record(float* rec, double px, double py, double hdx, double hdy,
                                                const double* s_walls, const double*, const Const& c, const EnvK&) {
    ovc_agent_record(rec, px, py, hdx, hdy, s_walls, c);
  }
  static __device__ __forceinline__ void load(Regs& r, const Const& c, int cell0) { ovc_load_cells(r, c, cell0); }
  template <bool DEFER, int EXP = -1>
  static __device__ __forceinline__ void rates4(float (&o)[4], const Regs& r, const Const& c, int, const float* rec,
                                                uint32_t, bool&) {
    ovc_rates4(o, r, c, rec);
  }
  static __device__ __forceinline__ int expanded(const Const&) { return 0; }
  static __device__ __forceinline__ int wall0(const Const&) { return 0; }
};

// ---------------------------------------------------------------------------
// k_step: persistent, warp-specialised step kernel (one CTA per SM).
//   warps [0, MW)        producers: each takes a tile of 32 agents, runs Agent.update for its
//                        lane's agent in float64 (MOTION) or just reads the position, writes the
//                        agent state back, and publishes a float32 "rate record" per agent into a
//                        shared-memory ring slot (mbarrier full[slot]).
//   warps [MW, MW+RW)    consumers: every thread keeps 4 consecutive cells in registers, waits for
//                        a slot, streams the slot's agents through the rate evaluator and writes
//                        float4 rate rows (+ noise + bit-packed spikes), then frees the slot
//                        (mbarrier empty[slot]).
// The float64 motion latency (a ~2.5k-instruction dependent chain) is thereby hidden behind the
// HBM-bound rate writes of earlier tiles instead of idling the CTA.
// Warp-role configuration.  The register file is re-balanced between the roles with setmaxnreg (the float64 motion code wants
// ~130 registers, the consumers 56..104).  Three splits, all 16 consumer warps (measurements: DESIGN.md section 4):
//   StepCfg<4>:  4 producers x 64 registers (spilling), consumers x 104 (640 threads x 96 at launch).  Pair loops with
//                spikes or OU noise: the consumers set the pace (c2 105 us, c3 with thinned spikes 80 us) and want the
//                registers.  (8 producers next to them measured worse for the thinned grid cells, 91 us, and within the
//                build-to-build spread for the dense stream: 105-111 us.)
//   StepCfg<8>:  8 producers x 128, consumers x 56 (768 threads x 80): light consumers without spikes (Euclidean Gaussian
//                place cells) run at the HBM write rate, so the float64 motion chain (~14 us per 32-agent tile) needs the
//                producer warps and the registers to keep up.
//   StepCfg<12>: 8 producers x 48 (spilling), consumers x 96 (768 threads x 80): heavier loops WITHOUT spikes (line of sight,
//                grid cells) and the light loop WITH the dense spike stream (c2e 59 -> 56 us).  With 4 producers these sat on the edge -- 3.5 tiles x 21-25 us per producer and step: the
//                no-spike c2 whole run measured 73-86 us from one build to the next (the placement of the consumers' loop
//                relative to the 64 KB of motion code the producers stream through the instruction cache seems to decide),
//                72-76 us with 8 producers; consumers at 96 instead of 104 registers lose < 2 %.
// A split that exceeds a sub-partition's launch allocation hangs (step_cfg_fits below).
constexpr int RW = 16;    // consumer warps
template <int ID>
struct StepCfg {
  static_assert(ID == 4 || ID == 8 || ID == 12, "unknown warp-role configuration");
  static constexpr int MW = (ID == 4) ? 4 : 8;                     // producer warps
  static constexpr int CTAS = 1;                                  // CTAs per SM
  static constexpr int NS = (ID == 8) ? MW : 2 * MW;              // ring slots (multiple of MW; static smem <= 48 KB)
  static constexpr int THREADS = (MW + RW) * 32;
  // what __launch_bounds__(THREADS, 1) allocates: the register file is per SM sub-partition (16384 registers, warp w on
  // sub-partition w % 4), so ptxas sizes for the fullest one -- 704 threads (22 warps, 6 on one sub-partition) get 80, not 88
  static constexpr int WARPS_SP = (MW + RW + 3) / 4;
  static constexpr int REGS_LAUNCH = (16384 / (WARPS_SP * 32)) / 8 * 8;
  static constexpr int REGS_PRODUCER = (ID == 4) ? 64 : (ID == 8) ? 128 : 48;
  static constexpr int REGS_CONSUMER = (ID == 4) ? 104 : (ID == 8) ? 56 : 96;
};
// (Round 2 also tried TWO CTAs per SM of 2 producer + 16 consumer warps with 2 cells per consumer thread -- 56 registers,
// twice the resident consumer warps.  Same step time within 3 %: the consumers are bound by the math dispatch port -- ALU-pipe
// and packed FP32 instructions hold it two cycles each -- not by latency, so more warps bought nothing.  The cell-count
// template parameter CPT of the policies is what remains of it.)
// setmaxnreg only moves registers WITHIN the launch allocation of a sub-partition (its warps x REGS_LAUNCH): an .inc blocks
// until enough warps have released theirs with .dec.  A split whose total exceeds the allocation therefore never
// completes -- the hang of round 1's 112 / 56 experiment (4 x 32 x 112 + 32 x 56 = 16128 > 5 x 32 x 96 = 15360) and of a
// 6-producer variant with 96-register consumers in round 2 (4 x 32 x 96 + 2 x 32 x 64 = 16384 > 6 x 32 x 80).
template <class C>
constexpr bool step_cfg_fits() {
  return (RW / 4) * 32 * C::REGS_CONSUMER + ((C::MW + 3) / 4) * 32 * C::REGS_PRODUCER <= C::WARPS_SP * 32 * C::REGS_LAUNCH;
}
static_assert(step_cfg_fits<StepCfg<4>>() && step_cfg_fits<StepCfg<8>>() && step_cfg_fits<StepCfg<12>>(),
              "setmaxnreg split exceeds the CTA's register allocation: the kernel would hang");
// setmaxnreg towards N registers from the launch allocation L (inc when N > L, dec when N < L)
template <int N, int L> __device__ __forceinline__ void reg_set() {
  if (N > L) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
  else if (N < L) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// ring slots of a (policy, configuration): the configuration's count, halved for the fat records (8 inner walls, object
// vector cells: 160 B per agent) while the ring would not fit in 40 KB of the 48 KB static shared memory
template <class P, class C>
constexpr int ring_slots() {
  int ns = C::NS;
  while (ns >= 2 * C::MW && ns * (TA * P::REC * 4 + 16) > 40 * 1024) ns /= 2;
  return ns;
}

// Consumer groups of the lean slot loop: each ring slot is consumed by ONE group, slot q by group q % G.  A group must see
// every phase of the slots it waits on (an mbarrier parity wait cannot skip a phase), so G has to divide the ring size.
__host__ __device__ constexpr int lean_groups(int cell_threads, int ring) {
  int g = (RW * 32) / (cell_threads > 0 ? cell_threads : 1);
  if (g < 1) g = 1;
  if (g > ring) g = ring;
  while (ring % g != 0) --g;
  return g;
}

template <int REC>
struct __align__(16) StepSlot {
  float rec[TA][REC];
  int na;
  unsigned nanmask;         // bit a: agent a's position is NaN -> its rates are zero (Neurons.py:163-164)
  int pad[2];
};

// The consumers' hot loop: consecutive agent pairs (2p, 2p+1) of one ring slot, 4 cells per thread, no OU noise,
// 16-byte aligned rows, even first global id.  The line-of-sight band test is deferred: a pair whose float32 decision fell
// inside the band only sets its bit in `redo`; the caller redoes those pairs through the general path
// (per-agent exact float64 fall-back) after the loop -- no call and no branch in here.
// DENSE: the dense spike stream (one Philox4x32-7 call per pair, a threshold test per rate) runs in the loop;
// thinned spikes are a post-pass over the slot (thin_block) and leave this loop spike-free.
template <class P, bool DENSE, int EXP>
__device__ __forceinline__ void consume_pairs(int& a, const int a_end, const typename P::Regs& regs,
                                              const typename P::Const& pc, const OutK& out, const TailCtx& tc,
                                              const int cell0, const float*& recp, const uint32_t inner_s, RowCursor& rc,
                                              const bool act, const float q16, uint32_t& redo) {
  float* dst = rc.dst;
  uint32_t bit = 1u;
  uint32_t* spk = rc.spk;
  unsigned long long pair = rc.gid >> 1;
  const long long pair_rate = 2ll * out.ld, pair_spk = 2ll * out.spike_ld;     // elements per pair step
  for (; a + 1 < a_end; a += 2) {
    float o[P::CPT];
    uint32_t c[4], bl[4];
    bool unsure = false;
    P::template rates4<true, EXP>(o, regs, pc, cell0, recp, inner_s, unsure);
    if (act) st_cs_fv<P::CPT>(dst, o);
    float nv = 0.f;
    if constexpr (DENSE && P::CPT == 4) {
      c[0] = (uint32_t)pair; c[1] = tc.sub ^ ((uint32_t)(pair >> 32) << 24); c[2] = tc.c2; c[3] = tc.c3_spk;
      philox_keyed<7>(c, out.rk7);
      nv = spike_neg_dither(c);
      spike_ballots<false, P::XU_BOUND>(bl, c[0], c[1], nv, o, q16, 0u, act);
      spike_store(bl, spk);
    }
    P::template rates4<true, EXP>(o, regs, pc, cell0, recp + P::REC, inner_s, unsure);
    if (act) st_cs_fv<P::CPT>(dst + out.ld, o);
    if constexpr (DENSE && P::CPT == 4) {
      spike_ballots<false, P::XU_BOUND>(bl, c[2], c[3], nv, o, q16, 0u, act);
      spike_store(bl, spk + out.spike_ld);
    }
    redo |= unsure ? bit : 0u;
    bit <<= 1;
    dst += pair_rate;
    spk += pair_spk;
    pair += 1ull;
    recp += 2 * P::REC;
  }
  rc.dst = dst; rc.spk = spk; rc.gid = pair << 1;
}

// ---------------------------------------------------------------------------
// Thinned spikes (Neurons.py:681-684: spike <=> uniform < dt * rate) for populations whose rates are bounded by `bound`
// with p' = dt * bound <= 1/16 (the usual case: dt = 10 ms, max_fr = 1 Hz gives p' = 0.01).  Exact Bernoulli(dt * rate) by
// thinning: every (agent, cell) is a CANDIDATE with probability p', a candidate spikes with probability rate / bound.
// Candidates are drawn per (agent row gid, 128-cell block B) -- the 32 x 128 rates one consumer warp has just stored for a
// ring slot, lane = row -- so that the result does not depend on tiles, shards or the launch path:
//   call n = 0, 1, ...:  R = Philox7(ctr = (gid, B | n << 16, step, THIN | population))
//   K       = #{k < 32 : R_0[0] >= cdf[k]},  cdf[k] = floor(2^32 P(Binomial(128, p') <= k))        (call 0 only)
//   draws   d = 4 n + j, j = 0..3:  position pos_d = (R_n[1] >> 7 j) & 127,
//                                   20-bit uniform x_d = (half-word j of (R_n[2], R_n[3])) << 4 | R_n[1] >> 28
//   the candidates are the first K DISTINCT positions of the draw sequence (a draw that repeats an earlier position is
//   skipped: sampling without replacement, i.e. a uniform K-subset, i.e. 128 independent Bernoulli(p') cells);
//   candidate at pos_d spikes  <=>  fma(float(x_d), c1, c0) < rate[gid, 128 B + pos_d]      (c1 = 2^-20 bound, c0 = 2^-21 bound).
// The pair loop does nothing for spikes.  After a slot's rates are stored the warp runs this once: ~1.3 candidates per
// lane at p' = 0.01, one Philox call per lane serves four of them; the rates are read back from L2 (this warp stored them),
// accepted bits go into the spike rows the producer cleared with RED.OR.  NumPy mirror: tests/philox_np.py (expected_spikes_thin).
// An out-of-line call (like slot_fixups): inlined, its ~40 live registers made ptxas park cell registers of the pair loop in
// local memory (8 LDL per pair iteration: the pass then cost more than the dense stream it replaces).  `out` is the kernel's
// __grid_constant__ parameter, so its address can be passed without a local copy.
__device__ __noinline__ void thin_block(const OutK* __restrict__ outp, const int cell0, const int n_cells, const uint32_t c2,
                                        const uint32_t c3_spk, const float* __restrict__ rates, uint32_t* __restrict__ spikes,
                                        const long long row_lo, const int rows) {
  const OutK& out = *outp;
  const int lane = threadIdx.x & 31;
  const int blk0 = cell0 - 4 * lane;                      // first cell of the warp's block (warp-uniform)
  const int cells_left = n_cells - blk0;
  if (lane >= rows || cells_left <= 0 || spikes == nullptr) return;   // (lanes leave independently: no warp-wide operation in here)
  const uint32_t B = (uint32_t)(blk0 >> 7);
  const long long row = row_lo + lane;
  const unsigned long long gid = (unsigned long long)(out.id_offset + row);
  const uint32_t c1w = B ^ ((uint32_t)(gid >> 32) << 24);
  const uint32_t c3w = (c3_spk & 0x00ffffffu) | (RIAB_STREAM_THIN << 24);
  uint32_t rk[14];
#pragma unroll
  for (int i = 0; i < 14; ++i) rk[i] = out.rk7[i];
  uint32_t R[4] = {(uint32_t)gid, c1w, c2, c3w};
  philox_keyed<7>(R, rk);
  int K = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) K += (R[0] >= out.thin_cdf[k]) ? 1 : 0;        // independent compares
  if (K == 8) while (K < 32 && R[0] >= out.thin_cdf[K]) ++K;
  if (K == 0) return;
  const float c1 = out.thin_c1, c0 = out.thin_c0;
  const float* rrow = rates + row * out.ld + blk0;
  uint32_t* srow = spikes + row * out.spike_ld + 4u * B;
  unsigned long long occ_lo = 0ull, occ_hi = 0ull;
  int cnt = 0;
  for (uint32_t n = 0u;;) {
    uint32_t take = 0u;
    int pos[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      pos[j] = (int)((R[1] >> (7 * j)) & 127u);
      const unsigned long long oh_lo = (pos[j] < 64) ? (1ull << pos[j]) : 0ull;
      const unsigned long long oh_hi = (pos[j] < 64) ? 0ull : (1ull << (pos[j] - 64));
      const bool fresh = (cnt < K) && (((occ_lo & oh_lo) | (occ_hi & oh_hi)) == 0ull);
      if (fresh) { occ_lo |= oh_lo; occ_hi |= oh_hi; ++cnt; }
      if (fresh && pos[j] < cells_left) take |= 1u << j;   // candidates on padding cells count, but have no rate
    }
    float rate[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      rate[j] = 0.f;
      if ((take >> j) & 1u) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(rate[j]) : "l"(rrow + pos[j]));
    }
    const uint32_t dith = R[1] >> 28;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t m = ((j < 2 ? R[2] : R[3]) >> (16 * (j & 1))) & 0xffffu;
      const float x = (float)((m << 4) | dith);
      if (((take >> j) & 1u) && fmaf(x, c1, c0) < rate[j]) atomicOr(srow + (pos[j] & 3), 1u << (pos[j] >> 2));
    }
    if (cnt >= K || ++n >= 256u) break;
    R[0] = (uint32_t)gid; R[1] = c1w ^ (n << 16); R[2] = c2; R[3] = c3w;
    philox_keyed<7>(R, rk);
  }
}

// The consumers' slot loop (see k_step).  EXP: exponent form of the fast pair loop (PlacePolicy::expanded).
template <class P, int SPK, bool NOISE, class C, int EXP>
__device__ __forceinline__ void consumer_slots(const typename P::Const& pc, const OutK& out, StepSlot<P::REC>* s_slot,
                                               uint64_t* s_full, uint64_t* s_empty, const double* s_walls,
                                               const long long nq, const int ctid, const int lane,
                                               const long long n_rows) {
  constexpr int NS = ring_slots<P, C>();
  constexpr int NC = RW * 32;
  constexpr bool DENSE = (SPK == 1);
  constexpr int CPT = P::CPT;
  static_assert(CPT == 4 || (SPK != 1 && !NOISE), "2 cells per thread: no dense spikes / OU noise (ballot layout)");
  const int CT = pc.n_pad / CPT;                      // cell-threads needed (multiple of 32)
  const int chunks = (CT + NC - 1) / NC;
  const int G = (chunks == 1) ? (NC / CT) : 1;        // agent groups when the cells need fewer threads
  const int grp = (chunks == 1) ? (ctid / CT) : 0;
  const bool idle = (chunks == 1) && (grp >= G);
  // group `grp` takes the consecutive agents [2 grp PPG, 2 (grp+1) PPG) of a slot (PPG pairs)
  const int PPG = (TA / 2 + G - 1) / G;
  typename P::Regs regs;
  int cell0 = (chunks == 1) ? (ctid % CT) * CPT : 0;
  TailCtx tc;
  if (chunks == 1 && !idle) {
    P::load(regs, pc, cell0);
    tail_init<CPT>(tc, out, cell0, pc.n_cells);
  }
  const uint32_t inner_s = smem_u32(s_walls) + 32u * (uint32_t)P::wall0(pc);   // float64 inner walls (exact fall-back)
  // Fast pair loop: rows are 16-byte aligned and every thread owns 4 existing cells or none, the
  // tile starts on an even global id (one Philox call per agent pair) and there is no OU noise.
  const bool fast = !NOISE && out.vec_ok && ((pc.n_cells & 3) == 0) && ((out.id_offset & 1ll) == 0 || SPK != 1);
  const float q16 = out.dt * 65536.0f;
  for (long long q = 0; q < nq; ++q) {
    const int s = (int)(q % NS);
    mbar_wait(&s_full[s], (uint32_t)((q / NS) & 1));
    const long long a0 = ((long long)blockIdx.x + q * gridDim.x) * out.tile_agents;
    const int na = s_slot[s].na;
    const int a_lo = 2 * grp * PPG;
    const int a_hi = (a_lo + 2 * PPG < na) ? a_lo + 2 * PPG : na;       // this group's agents of the slot: [a_lo, a_hi)
    // chunks == 1: the cell registers loaded above serve every slot; more than 2048 cells: the 16 warps walk
    // the cells in chunks of 2048 and reload their registers per chunk (G = 1, all warps on the same agents)
    for (int ch = 0; ch < chunks; ++ch) {
      if (chunks > 1) {
        cell0 = (ch * NC + ctid) * CPT;
        if (cell0 >= pc.n_pad) continue;              // warp-uniform (n_pad is a multiple of 128)
        P::load(regs, pc, cell0);
        tail_init<CPT>(tc, out, cell0, pc.n_cells);
      } else if (idle) {
        continue;
      }
      if (a_lo >= a_hi) continue;                     // warp-uniform
      const bool act = cell0 < pc.n_cells;
      {
        // agents are taken in pairs (2p, 2p+1) so that one Philox call feeds the (dense) spikes of both
        RowCursor rc;
        cursor_init(rc, out, tc, a0 + a_lo);
        const float* recp = s_slot[s].rec[a_lo];
        int a = a_lo;
        uint32_t only = 0xffffffffu;       // pairs (by iteration index) the general loop below evaluates
        if (fast) {
          uint32_t redo = 0u;
          consume_pairs<P, DENSE, EXP>(a, a_hi, regs, pc, out, tc, cell0, recp, inner_s, rc, act, q16, redo);
          redo = __reduce_or_sync(0xffffffffu, redo);
          if (const unsigned nm = s_slot[s].nanmask; nm != 0u)          // pairs with a NaN position: zero rates below
            for (int it = 0; a_lo + 2 * it < a_hi; ++it)
              if ((nm >> (a_lo + 2 * it)) & 3u) redo |= 1u << it;
          if (redo != 0u) {
            // some float32 line-of-sight decision was inside the band: redo those pairs through the
            // general path (the stores are idempotent); complete pairs not in `redo` are skipped
            only = redo;
            a = a_lo;
            cursor_init(rc, out, tc, a0 + a_lo);
            recp = s_slot[s].rec[a_lo];
          }
        }
        // general path: the last agent of an odd tile, odd shard offsets, OU noise, ragged cell counts
        const RowStride stride = make_stride(out, 2);
        const bool even = ((rc.gid & 1ull) == 0ull);      // uniform: a0 and a_lo are even
        for (uint32_t it = (uint32_t)((a - a_lo) >> 1); a < a_hi; a += 2, ++it) {
          float oa[CPT], ob[CPT];
          const bool has_b = (a + 1 < a_hi);
          if (has_b && !((only >> (it & 31u)) & 1u)) {      // warp-uniform
            cursor_advance(rc, stride);
            recp += 2 * P::REC;
            continue;
          }
          bool dummy = false;
          P::template rates4<false>(oa, regs, pc, cell0, recp, inner_s, dummy);
          if (has_b) P::template rates4<false>(ob, regs, pc, cell0, recp + P::REC, inner_s, dummy);
          if (const unsigned nm = s_slot[s].nanmask; nm != 0u) {        // NaN position -> zero rates (Neurons.py:163-164)
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
              if ((nm >> a) & 1u) oa[i] = 0.f;
              if (has_b && ((nm >> (a + 1)) & 1u)) ob[i] = 0.f;
            }
          }
          if constexpr (CPT == 4) {
            store4<NOISE>(oa, out, tc, rc, 0);
            if (has_b) store4<NOISE>(ob, out, tc, rc, out.ld);
            if (DENSE && (!NOISE || rc.spk != nullptr)) {
              if (has_b && even) {
                uint32_t c[4], bl[4];
                spike_words(c, out, tc, rc.gid);
                const float nv = spike_neg_dither(c);
                spike_ballots<true>(bl, c[0], c[1], nv, oa, q16, tc.vmask, true);
                spike_store(bl, rc.spk);
                spike_ballots<true>(bl, c[2], c[3], nv, ob, q16, tc.vmask, true);
                spike_store(bl, rc.spk + out.spike_ld);
              } else {
                spikes1(oa, out, tc, rc);
                if (has_b) {
                  RowCursor rb = rc;
                  rb.gid += 1; rb.spk += out.spike_ld;
                  spikes1(ob, out, tc, rb);
                }
              }
            }
          } else {
            // 2 cells per thread (launched only for 8-byte aligned rows and even cell counts, no noise, no dense spikes)
            if (tc.full4) {
              st_cs_fv<CPT>(rc.dst, oa);
              if (has_b) st_cs_fv<CPT>(rc.dst + out.ld, ob);
            } else {
#pragma unroll
              for (int i = 0; i < CPT; ++i)
                if ((tc.vmask >> i) & 1u) { st_cs_f1(rc.dst + i, oa[i]); if (has_b) st_cs_f1(rc.dst + out.ld + i, ob[i]); }
            }
          }
          cursor_advance(rc, stride);
          recp += 2 * P::REC;
        }
        if (SPK == 2) {
          __syncwarp();      // orders this warp's rate stores before the read-back
          if constexpr (CPT == 4) thin_block(&out, tc.cell0, tc.n_cells, tc.c2, tc.c3_spk, out.rates, out.spikes, a0 + a_lo, a_hi - a_lo);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&s_empty[s]);
  }
}

// Out-of-line repairs of one ring slot for consumer_fast (rare): pairs whose float32 line-of-sight decision fell inside the
// band (bit `it` of redo) are re-evaluated with the exact float64 fall-back, and an odd last agent gets its row.  A real
// call: its register needs must not shape the allocation of the hot loop (the arguments travel through the stack).
template <class P, bool DENSE>
__device__ __forceinline__ void slot_fixups(const typename P::Regs& regs, const typename P::Const& pc, const OutK& out,
                                            const TailCtx& tc, const float* rec, const uint32_t inner_s, float* d,
                                            uint32_t* spikes, const long long a0, const int n_agents, const uint32_t redo,
                                            const bool act) {
  constexpr int CPT = P::CPT;
  for (int a = 0; a < n_agents; a += 2, d += 2 * out.ld, rec += 2 * P::REC) {
    const bool has_b = a + 1 < n_agents;
    if (has_b && !((redo >> (a >> 1)) & 1u)) continue;              // warp-uniform
    float oa[CPT], ob[CPT];
    bool dummy = false;
    P::template rates4<false>(oa, regs, pc, tc.cell0, rec, inner_s, dummy);
    if (has_b) P::template rates4<false>(ob, regs, pc, tc.cell0, rec + P::REC, inner_s, dummy);
    if (act) {
      st_cs_fv<CPT>(d, oa);
      if (has_b) st_cs_fv<CPT>(d + out.ld, ob);
    }
    if constexpr (DENSE && CPT == 4) {                              // whole warp: ballots
      RowCursor rc;
      rc.gid = (unsigned long long)(out.id_offset + a0 + a);
      rc.spk = spikes + (a0 + a) * out.spike_ld + ((tc.cell0 >> 7) << 2);
      const float q16 = out.dt * 65536.0f;
      if (has_b) {
        uint32_t c[4], bl[4];
        spike_words(c, out, tc, rc.gid);
        const float nv = spike_neg_dither(c);
        spike_ballots<true>(bl, c[0], c[1], nv, oa, q16, tc.vmask, true);
        spike_store(bl, rc.spk);
        spike_ballots<true>(bl, c[2], c[3], nv, ob, q16, tc.vmask, true);
        spike_store(bl, rc.spk + out.spike_ld);
      } else {
        spikes1(oa, out, tc, rc);
      }
    }
  }
}

// The consumers' slot loop for the common case: no OU noise, no dense spike stream, vector-aligned rows, whole 4-cell
// groups, all cells resident in one set of registers (n_pad <= 512 * CPT).  Pointers advance incrementally, the rare
// repairs are a real call (slot_fixups), thinned spikes are a post-pass per slot (thin_block).
template <class P, int SPK, class C, int EXP, bool MULTI>
__device__ __forceinline__ void consumer_fast(const typename P::Const& pc, const OutK& out, const RunK& run, StepSlot<P::REC>* s_slot,
                                              uint64_t* s_full, uint64_t* s_empty, const double* s_walls, const long long nq,
                                              const int ctid, const int lane, const long long n_rows) {
  constexpr int NS = ring_slots<P, C>(), MW = C::MW, NSP = NS / MW, CPT = P::CPT;
  const int CT = pc.n_pad / CPT;                      // cell-threads needed (multiple of 32, <= RW * 32)
  const int G = lean_groups(CT, NS), grp = ctid / CT; // groups of CT threads; group g consumes the tiles q = g, g + G, ...
  if (grp >= G) return;                               // spare warps (the slots' release count is one group's warps)
  constexpr int a_lo = 0;
  typename P::Regs regs;
  const int cell0 = (ctid % CT) * CPT;
  TailCtx tc;
  P::load(regs, pc, cell0);
  tail_init<CPT>(tc, out, cell0, pc.n_cells);
  const bool act = cell0 < pc.n_cells;                // all CPT cells exist or none (n_cells % 4 == 0)
  const uint32_t inner_s = smem_u32(s_walls) + 32u * (uint32_t)P::wall0(pc);
  const long long ld = out.ld;
  [[maybe_unused]] const float q16 = out.dt * 65536.0f;
  const long long a_first = ((long long)blockIdx.x + (long long)grp * gridDim.x) * out.tile_agents;   // first row of the group's first tile
  const long long a_step = (long long)G * gridDim.x * out.tile_agents, slot_step = a_step * ld;
  const long long n_steps = MULTI ? run.n_steps : 1;
  for (long long st = 0; st < n_steps; ++st) {
    // this step's rows: the caller's (single step) or the rings' row (ring_next + st) % ring_rows
    float* rates = out.rates;
    uint32_t* spikes = out.spikes;
    if (MULTI) {
      const long long slot = (run.ring_next + st) % run.ring_rows;
      rates = run.rates_ring + slot * n_rows * ld;
      spikes = run.spikes_ring ? run.spikes_ring + slot * n_rows * out.spike_ld : nullptr;
      tail_init<CPT>(tc, out, cell0, pc.n_cells, out.step + (unsigned long long)st);
    }
    long long a0 = a_first;
    float* dst0 = rates + a0 * ld + cell0;
    for (long long q = grp; q < nq; q += G, dst0 += slot_step, a0 += a_step) {
      // tile q is produced by warp q % MW as its n-th tile overall: slot and phase of that producer's private ring
      const int pw = (int)(q % MW);
      const long long npw = (nq - pw + MW - 1) / MW;    // tiles of that producer per step
      const long long n = st * npw + q / MW;
      const int s = pw + MW * (int)(n % NSP);
      mbar_wait(&s_full[s], (uint32_t)((n / NSP) & 1));
      const int a_hi = s_slot[s].na;
      if (a_lo < a_hi) {
        const float* recp = s_slot[s].rec[a_lo];
        float* d = dst0;
        uint32_t redo = 0u;
        const int n2 = (a_hi - a_lo) >> 1;
        [[maybe_unused]] uint32_t* spk = nullptr;
        [[maybe_unused]] unsigned long long pair = 0ull;
        if constexpr (SPK == 1) {
          spk = spikes + a0 * out.spike_ld + ((cell0 >> 7) << 2);
          pair = (unsigned long long)(out.id_offset + a0) >> 1;          // even first global id: rows (2p, 2p+1) are one pair
        }
        for (int it = 0; it < n2; ++it) {
          float o[CPT];
          bool unsure = false;
          P::template rates4<true, EXP>(o, regs, pc, cell0, recp, inner_s, unsure);
          if (act) st_cs_fv<CPT>(d, o);
          [[maybe_unused]] uint32_t c[4], bl[4];
          [[maybe_unused]] float nv = 0.f;
          if constexpr (SPK == 1) {
            // dense spike stream: one Philox4x32-7 call per (agent pair, 4-cell group), a threshold test per rate
            c[0] = (uint32_t)pair; c[1] = tc.sub ^ ((uint32_t)(pair >> 32) << 24); c[2] = tc.c2; c[3] = tc.c3_spk;
            philox_keyed<7>(c, out.rk7);
            nv = spike_neg_dither(c);
            spike_ballots<false, P::XU_BOUND>(bl, c[0], c[1], nv, o, q16, 0u, act);
            spike_store(bl, spk);
          }
          P::template rates4<true, EXP>(o, regs, pc, cell0, recp + P::REC, inner_s, unsure);
          if (act) st_cs_fv<CPT>(d + ld, o);
          if constexpr (SPK == 1) {
            spike_ballots<false, P::XU_BOUND>(bl, c[2], c[3], nv, o, q16, 0u, act);
            spike_store(bl, spk + out.spike_ld);
            spk += 2 * out.spike_ld;
            pair += 1ull;
          }
          redo |= (unsure ? 1u : 0u) << it;
          d += 2 * ld;
          recp += 2 * P::REC;
        }
        redo = __reduce_or_sync(0xffffffffu, redo);
        if (redo != 0u || ((a_hi - a_lo) & 1)) {
          const typename P::Regs rcopy = regs;            // stack copies, made on this path only
          const typename P::Const pcopy = pc;
          slot_fixups<P, SPK == 1>(rcopy, pcopy, out, tc, s_slot[s].rec[a_lo], inner_s, dst0, spikes, a0, a_hi - a_lo, redo, act);
        }
        if (const unsigned nm = s_slot[s].nanmask; nm != 0u && act) {     // NaN position -> zero rates (Neurons.py:163-164)
          float z[CPT];
#pragma unroll
          for (int i = 0; i < CPT; ++i) z[i] = 0.f;
          for (int a = a_lo; a < a_hi; ++a)
            if ((nm >> a) & 1u) st_cs_fv<CPT>(dst0 + (long long)(a - a_lo) * ld, z);
        }
        if constexpr (SPK == 2) {
          __syncwarp();                                  // this warp's rate stores before the read-back
          thin_block(&out, tc.cell0, tc.n_cells, tc.c2, tc.c3_spk, rates, spikes, a0 + a_lo, a_hi - a_lo);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[s]);
    }
  }
}

// SPK: 0 no spikes, 1 dense spike stream (in the loops), 2 thinned spikes (thin_block per ring slot)
template <class P, int MODE, int SPK, bool NOISE, class C>
__global__ void __launch_bounds__(C::THREADS, C::CTAS) k_step(const EnvK env, const riab_agents ag,
                                                          const riab_motion_params mp, const MotionDerived md,
                                                          const riab_step_io io, const typename P::Const pc, const __grid_constant__ OutK out,
                                                          const double* __restrict__ pos_in, const long long n_rows,
                                                          const RunK run) {
  __shared__ __align__(16) double s_walls[MAXW * 4];
  constexpr int MW = C::MW, NS = ring_slots<P, C>();
  __shared__ StepSlot<P::REC> s_slot[NS];
  __shared__ uint64_t s_bar, s_full[NS], s_empty[NS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // lean consumers (consumer_fast): every ring slot is consumed by ONE group of warps (n_pad / CPT threads), the general
  // loop (consumer_slots) by all RW consumer warps
  bool lean = false;
  if constexpr (!NOISE && (SPK != 1 || P::CPT == 4))
    lean = out.vec_ok && ((pc.n_cells & 3) == 0) && (pc.n_pad <= RW * 32 * P::CPT) && (out.spikes == nullptr || ((out.id_offset & 1ll) == 0));
  if (threadIdx.x == 0) {
    const uint32_t n_release = lean ? (uint32_t)((pc.n_pad / P::CPT) >> 5) : (uint32_t)RW;
    for (int i = 0; i < NS; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], n_release); }
    mbar_fence_init();
  }
  stage_walls(s_walls, &s_bar, env);     // includes __syncthreads()
  __shared__ double s_aux[2 * PLACE_MAX_WI];             // per-wall invariants of the agent records (policy-specific)
  P::prepare(s_aux, s_walls, pc);
  __syncthreads();

  const int ta = out.tile_agents;
  const long long n_tiles = (n_rows + ta - 1) / ta;
  const long long nq = (n_tiles > (long long)blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

#ifdef RIAB_PRODUCERS_FIRST
  const bool producer = warp < MW;
  const int pw = warp, ctid0 = MW * 32;
#else
  // producers take the HIGHEST warp ids: the issue arbiter prefers higher warp ids among eligible warps, and
  // the float64 motion chain (one instruction every ~20 cycles) must not wait behind 4 busy consumers
  const bool producer = warp >= RW;
  const int pw = warp - RW, ctid0 = 0;
#endif
  if (producer) {
    // ------------------------------------------------------------- producers
    reg_set<C::REGS_PRODUCER, C::REGS_LAUNCH>();
    if constexpr (MODE == 3) {
      // whole run: this warp advances ITS tiles (q = pw, pw + MW, ...) step after step -- the same warp for every step of a
      // tile, so a tile's steps are ordered -- and publishes each (step, tile) record into its private ring (slots
      // pw + MW * i): the n-th record it produces goes to slot i = n % NSP in phase n / NSP
      constexpr int NSP = NS / MW;
      const long long npw = (nq > pw) ? (nq - pw + MW - 1) / MW : 0;
      long long n = 0;
      for (long long st = 0; st < run.n_steps; ++st) {
        riab_step_io io_st = io;
        io_st.step = io.step + (unsigned long long)st;
        io_st.history_row = run.hist_ring ? run.hist_ring + ((run.hist_next + st) % run.hist_rows) * n_rows * 8 : nullptr;
        uint32_t* const spikes_st = run.spikes_ring ? run.spikes_ring + ((run.ring_next + st) % run.ring_rows) * n_rows * out.spike_ld : nullptr;
        for (long long q = pw; q < nq; q += MW, ++n) {
          const int s = pw + MW * (int)(n % NSP);
          mbar_wait(&s_empty[s], (uint32_t)(((n / NSP) & 1) ^ 1));
          const long long a0 = ((long long)blockIdx.x + q * gridDim.x) * ta;
          const int na = (int)((n_rows - a0) < ta ? (n_rows - a0) : ta);
          if (SPK == 2 && lane < na && spikes_st != nullptr) {
            uint32_t* z = spikes_st + (a0 + lane) * out.spike_ld;
            for (long long w = 0; w < out.spike_ld; w += 4)
              asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(z + w), "r"(0u) : "memory");
          }
          bool nanpos = false;
          if (lane < na) {
            AgentState as;
            agent_update_one<false>(ag, mp, md, io_st, env, s_walls, a0 + lane, as);
            nanpos = (as.px != as.px);
            P::record(s_slot[s].rec[lane], as.px, as.py, as.hdx, as.hdy, s_walls, s_aux, pc, env);
          }
          const unsigned nanmask = __ballot_sync(0xffffffffu, nanpos);
          if (lane == 0) { s_slot[s].na = na; s_slot[s].nanmask = nanmask; }
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_full[s]);
        }
      }
      return;
    }
    for (long long q = pw; q < nq; q += MW) {
      const int s = (int)(q % NS);
      const uint32_t k = (uint32_t)(q / NS);
      mbar_wait(&s_empty[s], (k & 1u) ^ 1u);
      const long long tile = (long long)blockIdx.x + q * gridDim.x;
      const long long a0 = tile * ta;
      const int na = (int)((n_rows - a0) < ta ? (n_rows - a0) : ta);
      if (SPK == 2 && lane < na) {
        // thinned spikes: clear the tile's spike rows; the consumers OR accepted bits in after the slot is published
        // (mbarrier release / acquire orders these stores before their RED.ORs)
        uint32_t* z = out.spikes + (a0 + lane) * out.spike_ld;
        for (long long w = 0; w < out.spike_ld; w += 4)
          asm volatile("st.global.cs.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(z + w), "r"(0u) : "memory");
      }
      if (MODE == 2) {
        // skewed: publish the records of the CURRENT positions first, then advance the agents
        // (the next launch's rates) -- consumers never wait for the float64 motion chain.
        bool nanpos = false;
        if (lane < na) {
          const long long i = a0 + lane;
          const double px = ag.pos[2 * i], py = ag.pos[2 * i + 1];
          nanpos = (px != px);
          P::record(s_slot[s].rec[lane], px, py, ag.head_direction[2 * i], ag.head_direction[2 * i + 1], s_walls, s_aux, pc, env);
        }
        const unsigned nanmask = __ballot_sync(0xffffffffu, nanpos);
        if (lane == 0) { s_slot[s].na = na; s_slot[s].nanmask = nanmask; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_full[s]);
        if (lane < na) {
          AgentState st;
          agent_update_one<false>(ag, mp, md, io, env, s_walls, a0 + lane, st);
        }
        continue;
      }
      bool nanpos = false;
      if (lane < na) {
        const long long i = a0 + lane;
        double px, py, hdx = 1.0, hdy = 0.0;
        if (MODE == 1) {
          AgentState st;
          agent_update_one<false>(ag, mp, md, io, env, s_walls, i, st);
          px = st.px; py = st.py; hdx = st.hdx; hdy = st.hdy;
        } else {
          px = pos_in[2 * i]; py = pos_in[2 * i + 1];
          const double* hd = P::head_dir(pc);                 // egocentric cells evaluated at given positions
          if (hd != nullptr) { hdx = hd[2 * i]; hdy = hd[2 * i + 1]; }
        }
        nanpos = (px != px);
        P::record(s_slot[s].rec[lane], px, py, hdx, hdy, s_walls, s_aux, pc, env);
      }
      const unsigned nanmask = __ballot_sync(0xffffffffu, nanpos);
      if (lane == 0) { s_slot[s].na = na; s_slot[s].nanmask = nanmask; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_full[s]);
    }
  } else {
    // ------------------------------------------------------------- consumers
    reg_set<C::REGS_CONSUMER, C::REGS_LAUNCH>();
    const int ctid = threadIdx.x - ctid0;
    // one copy of the slot loop per exponent form (0: direct, 1: expanded, 2: expanded + folded scale), chosen once:
    // the cell registers then stay in registers across slots (a run-time switch inside the loop made ptxas park them
    // in local memory around every slot)
    const int ex = P::expanded(pc);
    if constexpr (!NOISE && (SPK != 1 || P::CPT == 4)) {
      if (lean) {
        if (ex == 2) consumer_fast<P, SPK, C, 2, MODE == 3>(pc, out, run, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
        else if (ex == 1) consumer_fast<P, SPK, C, 1, MODE == 3>(pc, out, run, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
        else consumer_fast<P, SPK, C, 0, MODE == 3>(pc, out, run, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
        return;
      }
    }
    if constexpr (MODE == 3) return;                    // (the host launches whole runs only where the lean loop applies)
    if (ex == 2) consumer_slots<P, SPK, NOISE, C, 2>(pc, out, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
    else if (ex == 1) consumer_slots<P, SPK, NOISE, C, 1>(pc, out, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
    else consumer_slots<P, SPK, NOISE, C, 0>(pc, out, s_slot, s_full, s_empty, s_walls, nq, ctid, lane, n_rows);
  }
}

// ---------------------------------------------------------------------------
// PlaceCells description "one_hot" (Neurons.py:972-974): rate = 1 for the cell with the smallest
// distance (np.argmin: first index of the minimum), 0 elsewhere.  One warp per position.  Pass 1 finds the
// minimum float32 squared distance (line-of-sight flags from the same float32 predicate as the rate
// kernels; pairs inside the predicate's uncertainty band count with their unblocked distance); pass 2
// re-evaluates every cell within a relative 1e-5 of that minimum in float64 exactly like the reference
// (np.linalg.norm, utils.vector_intercepts) and keeps the smallest (distance, index).
__device__ __forceinline__ double onehot_exact_dist(const PlaceConst& c, int cell, double px, double py,
                                                    const double* __restrict__ inner64) {
  const double cx = c.centres64[2 * cell], cy = c.centres64[2 * cell + 1];
  bool blocked = false;
  for (int j = 0; j < c.n_inner; ++j) blocked = blocked || los_blocked_exact(cx, cy, px, py, inner64 + 4 * j);
  D ex = D(cx) - D(px), ey = D(cy) - D(py);
  if (c.periodic) {
    if (fabs(ex.v) > c.scale / 2) ex = D(-copysign(1.0, ex.v)) * (D(c.scale) - D(fabs(ex.v)));
    if (fabs(ey.v) > c.scale / 2) ey = D(-copysign(1.0, ey.v)) * (D(c.scale) - D(fabs(ey.v)));
  }
  const double d = dsqrt(ex * ex + ey * ey).v;
  if (!blocked) return d;
  if (c.geometry == RIAB_GEOM_GEODESIC) {
    double via = INFINITY;
    for (int e = 0; e < 2; ++e) {
      if (!((c.ep_valid >> e) & 1)) continue;
      const D wx(inner64[2 * e]), wy(inner64[2 * e + 1]);
      const D ax = D(cx) - wx, ay = D(cy) - wy, bx = wx - D(px), by = wy - D(py);
      const double v = (dsqrt(ax * ax + ay * ay) + dsqrt(bx * bx + by * by)).v;
      via = fmin(via, v);
    }
    return via;
  }
  return 1000.0;
}

__global__ void __launch_bounds__(NT) k_place_onehot(const EnvK env, const PlaceConst pc, const double* __restrict__ pos,
                                                     const long long n_rows, const OutK out) {
  __shared__ __align__(16) double s_walls[MAXW * 4];
  __shared__ uint64_t s_bar;
  stage_walls(s_walls, &s_bar, env);
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const double px = pos[2 * row], py = pos[2 * row + 1];
  const double* inner = s_walls + 4 * pc.wall0;
  const float pxf = (float)(px - env.cxm), pyf = (float)(py - env.cym);
  float fp[PLACE_MAX_WI], tp[PLACE_MAX_WI];
  for (int j = 0; j < PLACE_MAX_WI; ++j) {
    fp[j] = 1.f; tp[j] = 0.f;
    if (j < pc.n_inner) {
      double f, t;
      wall_coords(px, py, inner[4 * j], inner[4 * j + 1], inner[4 * j + 2], inner[4 * j + 3], f, t);
      fp[j] = (float)f; tp[j] = (fabs(f) < 1e-9) ? nanf("") : (float)t;
    }
  }
  const int np = pc.n_pad;
  auto d2_of = [&](int cell, bool& unsure) -> float {          // optimistic float32 squared distance of one cell
    float dx = fabsf(pxf - pc.packed[cell]), dy = fabsf(pyf - pc.packed[np + cell]);
    if (pc.periodic) {
      dx = (dx > pc.half_f) ? pc.scale_f - dx : dx;
      dy = (dy > pc.half_f) ? pc.scale_f - dy : dy;
    }
    float d2 = fmaf(dy, dy, dx * dx);
    bool hit = false;
    for (int j = 0; j < pc.n_inner; ++j) {
      const float fc = pc.packed[(size_t)(4 + 2 * j) * np + cell], tcv = pc.packed[(size_t)(5 + 2 * j) * np + cell];
      const float afp = fabsf(fp[j]);
      const float Mp = fmaf(afp, tcv, fabsf(fc) * tp[j]);
      const float mn = fminf(Mp, (fabsf(fc) + afp) - Mp);
      const bool u = !(fabsf(mn) >= pc.eps[j]);
      unsure = unsure || u;
      hit = hit || (((fc * fp[j]) < 0.f) && (mn > 0.f) && !u);
    }
    return hit ? 1.0e6f : d2;
  };
  float best = INFINITY;
  for (int cell = lane; cell < pc.n_cells; cell += 32) {
    bool u = false;
    best = fminf(best, d2_of(cell, u));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = fminf(best, __shfl_xor_sync(0xffffffffu, best, o));
  // geodesic detours are shorter than 1000: every blocked cell is a candidate there
  const float thr = (pc.geometry == RIAB_GEOM_GEODESIC && pc.n_inner > 0) ? INFINITY : best * (1.0f + 1e-5f) + 1e-12f;
  double bd = INFINITY;
  int bi = 0x7fffffff;
  for (int cell = lane; cell < pc.n_cells; cell += 32) {
    bool u = false;
    const float d2 = d2_of(cell, u);
    if (d2 <= thr || u) {
      const double d = onehot_exact_dist(pc, cell, px, py, inner);
      if (d < bd || (d == bd && cell < bi)) { bd = d; bi = cell; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, bd, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
  }
  const float lo = pc.min_fr, hi = pc.min_fr + pc.span;
  float* dst = out.rates + row * out.ld;
  for (int cell = lane; cell < pc.n_cells; cell += 32) st_cs_f1(dst + cell, cell == bi ? hi : lo);
}

// Noise + spikes post-pass over rate rows that a kernel without finish4 produced (BVC).
__global__ void __launch_bounds__(NT) k_finish_rows(const OutK out, const int n_cells, const int n_pad128,
                                                    const long long n_rows) {
  const long long row = blockIdx.x;                        // rows on x: gridDim.y stops at 65535
  const int cell0 = (blockIdx.y * NT + threadIdx.x) * 4;
  if (row >= n_rows || cell0 >= n_pad128) return;          // warp-uniform: a warp covers 128 consecutive cells
  float o[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) o[i] = (cell0 + i < n_cells) ? out.rates[row * out.ld + cell0 + i] : 0.f;
  TailCtx tc;
  tail_init(tc, out, cell0, n_cells);
  tc.full4 = false;
  RowCursor rc;
  cursor_init(rc, out, tc, row);
  finish4<true, true>(o, out, tc, rc);
}

// ---------------------------------------------------------------------------
// BVC phase A (+ optional fused Agent.update): one CTA per tile of 32 agents.
template <bool FUSED, bool REC, bool TABLE>
__global__ void __launch_bounds__(NT, (TABLE && !FUSED) ? 4 : 1) k_bvc_rays(const EnvK env, const riab_agents ag, const riab_motion_params mp,
                                                 const MotionDerived md, const riab_step_io io, const BvcConst bc,
                                                 const double* __restrict__ pos_in, const long long n_rows,
                                                 float* __restrict__ scratch, int32_t* __restrict__ first_wall,
                                                 uint32_t* __restrict__ spikes_zero, const long long spike_ld) {
  extern __shared__ __align__(128) unsigned char dyn[];
  double* s_dirs = reinterpret_cast<double*>(dyn);                 // T*2
  __shared__ __align__(16) double s_walls[MAXW * 4];
  __shared__ __align__(16) float4 s_wf[MAXW];
  __shared__ __align__(16) double s_pos[TA][2];
  __shared__ uint64_t s_bar;
  stage_walls(s_walls, &s_bar, env);
  for (int i = threadIdx.x; i < 2 * bc.T; i += blockDim.x) s_dirs[i] = bc.test_dirs[i];
  for (int w = threadIdx.x; w < env.W; w += blockDim.x)
    s_wf[w] = make_float4((float)s_walls[4 * w], (float)s_walls[4 * w + 1], (float)(s_walls[4 * w + 2] - s_walls[4 * w]),
                          (float)(s_walls[4 * w + 3] - s_walls[4 * w + 1]));

  const long long a0 = (long long)blockIdx.x * TA;
  const int na = (int)((n_rows - a0) < TA ? (n_rows - a0) : TA);
  if (threadIdx.x < TA) {
    double px = 0.5 * (env.ext[0] + env.ext[1]), py = 0.5 * (env.ext[2] + env.ext[3]);   // padding rows: box centre
    if ((int)threadIdx.x < na) {
      const long long i = a0 + threadIdx.x;
      if (FUSED) {
        AgentState s;
        agent_update_one<REC>(ag, mp, md, io, env, s_walls, i, s);
        px = s.px; py = s.py;
      } else {
        px = pos_in[2 * i]; py = pos_in[2 * i + 1];
      }
    }
    s_pos[threadIdx.x][0] = px;
    s_pos[threadIdx.x][1] = py;
  }
  __syncthreads();
  if (spikes_zero != nullptr) {
    // the integration kernel ORs this step's spikes into the tile's (contiguous) spike rows: clear them here
    uint4* z = reinterpret_cast<uint4*>(spikes_zero + a0 * spike_ld);
    for (long long w = threadIdx.x; w < (long long)na * spike_ld / 4; w += blockDim.x) z[w] = make_uint4(0u, 0u, 0u, 0u);
  }
  float* tile = scratch + (size_t)blockIdx.x * bc.T * BVC_AT;
  float cmax = 1.f, lmax = 0.f;                         // error floors of the float32 screens (riab_bvc.cuh)
  for (int w = 0; w < env.W; ++w) {
    const float4 wl = s_wf[w];
    cmax = fmaxf(cmax, fmaxf(fmaxf(fabsf(wl.x), fabsf(wl.y)), fmaxf(fabsf(wl.x + wl.z), fabsf(wl.y + wl.w))));
    lmax = fmaxf(lmax, fabsf(wl.z) + fabsf(wl.w));
  }
  if (TABLE) {
    // (angle, wall) table + float32 directions, then one agent per thread for all its angles (idx & 31 is constant)
    float2* s_dirf = reinterpret_cast<float2*>(s_dirs + 2 * bc.T);
    BvcTab* s_tab = reinterpret_cast<BvcTab*>(s_dirf + bc.T);
    const int W = env.W;
    for (int e = threadIdx.x; e < bc.T * W; e += blockDim.x) {
      const int th = e / W, w = e - th * W;
      s_tab[e] = bvc_table_entry(s_dirs[2 * th], s_dirs[2 * th + 1], s_walls + 4 * w, cmax);
    }
    for (int th = threadIdx.x; th < bc.T; th += blockDim.x) s_dirf[th] = make_float2((float)s_dirs[2 * th], (float)s_dirs[2 * th + 1]);
    __syncthreads();
    const int a = threadIdx.x & 31;
    const double px = s_pos[a][0], py = s_pos[a][1];
    const float pxf = (float)px, pyf = (float)py;
    const float ka = 1e-9f * cmax * lmax;
    float numA[BVC_NW];
#pragma unroll
    for (int w = 0; w < BVC_NW; ++w) {
      numA[w] = 0.f;
      if (w < W) {
        const double ax = s_walls[4 * w], ay = s_walls[4 * w + 1];
        numA[w] = (float)((ax - px) * (s_walls[4 * w + 3] - ay) - (ay - py) * (s_walls[4 * w + 2] - ax));   // (a - p) x sb
      }
    }
    for (int th = threadIdx.x >> 5; th < bc.T; th += NT / 32) {
      const float2 u = s_dirf[th];
      uint32_t mask = bvc_table_mask<BVC_NW>(s_tab + th * W, numA, W, pxf * u.y - pyf * u.x, ka);
      if (!(fabsf(pxf) + fabsf(pyf) <= 4.f * cmax)) mask = 0xffffffffu >> (32 - W);   // far outside (or NaN): no screen
      double d;
      int wid;
      bvc_walk<uint32_t>(mask, px, py, s_dirs[2 * th], s_dirs[2 * th + 1], s_walls, d, wid);
      tile[th * BVC_AT + a] = (float)d;
      if (first_wall != nullptr && a < na) first_wall[(a0 + a) * bc.T + th] = wid;
    }
    return;
  }
  const float flo_env = 1e-6f * cmax * lmax;
  for (int idx = threadIdx.x; idx < bc.T * BVC_AT; idx += blockDim.x) {
    const int th = idx >> 5, a = idx & 31;
    double d;
    int wid;
    bvc_first_wall(s_pos[a][0], s_pos[a][1], s_dirs[2 * th], s_dirs[2 * th + 1], s_walls, s_wf, env.W, flo_env, d, wid);
    tile[idx] = (float)d;
    if (first_wall != nullptr && a < na) first_wall[(a0 + a) * bc.T + th] = wid;
  }
}

// BVC phase B: grid.x = cell tiles, grid.y = agent-tile lanes; 256 threads:
// cell = tid & 63, agent group g = tid >> 6 handles agents 8g..8g+7 of the tile.
// Neurons.save_to_history spikes (Neurons.py:681-684) of one cell for 8 consecutive agents (rows r0 .. r0+7, r0 even, even
// shard offset), the dense stream of spike_words / spike_ballots: one Philox4x32-7 call per (agent pair, 4-cell group) --
// every thread of a group repeats it for its own cell, 4 calls per 180 x 8 integrand terms -- and RED.OR into the rows
// k_bvc_rays cleared.
__device__ __forceinline__ void bvc_spikes8(const OutK& out, const float (&v)[8], const long long r0, const long long n_rows,
                                            const int cell) {
  const int i = cell & 3;
  const uint32_t sub = (uint32_t)(cell >> 2);
  const uint32_t hi = ((uint32_t)(out.step >> 32) & 0xffffu) | (((uint32_t)out.pop & 0xffu) << 16) | (RIAB_STREAM_SPIKES << 24);
  const float q16 = out.dt * 65536.0f;
  const uint32_t bit = 1u << ((cell >> 2) & 31);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const long long row = r0 + 2 * k;
    if (row >= n_rows) break;
    const unsigned long long pair = (unsigned long long)(out.id_offset + row) >> 1;
    uint32_t c[4];
    c[0] = (uint32_t)pair; c[1] = sub ^ ((uint32_t)(pair >> 32) << 24); c[2] = (uint32_t)out.step; c[3] = hi;
    philox_keyed<7>(c, out.rk7);
    const float nv = spike_neg_dither(c);
    const uint32_t wa = (i < 2) ? c[0] : c[1], wb = (i < 2) ? c[2] : c[3];
    const float ma = (float)((wa >> (16 * (i & 1))) & 0xffffu), mb = (float)((wb >> (16 * (i & 1))) & 0xffffu);
    uint32_t* w = out.spikes + row * out.spike_ld + ((cell >> 7) << 2) + i;
    if (ma < fmaf(v[2 * k], q16, nv)) atomicOr(w, bit);
    if (row + 1 < n_rows && mb < fmaf(v[2 * k + 1], q16, nv)) atomicOr(w + out.spike_ld, bit);
  }
}

__global__ void __launch_bounds__(NT) k_bvc_integrate(const BvcConst bc, const float* __restrict__ scratch,
                                                      const long long n_rows, const long long n_tiles,
                                                      const OutK out, const int fold_spikes) {
  extern __shared__ __align__(128) unsigned char dyn[];
  const int T = bc.T;
  float* s_vm = reinterpret_cast<float*>(dyn);                     // [T][64]
  float* s_d0 = s_vm + (size_t)T * BVC_CT;                         // [T][32] x 2 buffers
  float* s_d1 = s_d0 + (size_t)T * BVC_AT;
  __shared__ uint64_t bar_vm, bar_d[2];
  const int ct = blockIdx.x;
  const int tid = threadIdx.x, cl = tid & 63, g = tid >> 6;
  const int slot = ct * BVC_CT + cl;
  // slot -> cell and this warp's angular window (riab_bvc_pack): outside [th0, th0 + tlen) mod T all 32 von Mises weights
  // are < 2^-30 of their peak
  const int32_t* perm = reinterpret_cast<const int32_t*>(bc.packed + 6 * (size_t)bc.n_pad + (size_t)bc.n_pad * T + 2 * (size_t)T);
  const int cell = perm[slot];
  const int th0 = perm[bc.n_pad + 2 * (slot >> 5)], tlen = perm[bc.n_pad + 2 * (slot >> 5) + 1];
  const uint32_t vm_bytes = (uint32_t)T * BVC_CT * 4u, d_bytes = (uint32_t)T * BVC_AT * 4u;
  const float* vm_src = bc.packed + 3 * (size_t)bc.n_pad + (size_t)ct * T * BVC_CT;
  if (tid == 0) {
    mbar_init(&bar_vm, 1); mbar_init(&bar_d[0], 1); mbar_init(&bar_d[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&bar_vm, vm_bytes);
    tma_bulk_g2s(s_vm, vm_src, vm_bytes, &bar_vm);
    long long t0 = blockIdx.y;
    if (t0 < n_tiles) { mbar_expect_tx(&bar_d[0], d_bytes); tma_bulk_g2s(s_d0, scratch + (size_t)t0 * T * BVC_AT, d_bytes, &bar_d[0]); }
  }
  __syncthreads();
  const float sc = bc.packed[cell], mc = bc.packed[bc.n_pad + cell], scale = bc.packed[2 * bc.n_pad + cell];
  mbar_wait(&bar_vm, 0);
  uint32_t phase[2] = {0u, 0u};
  int buf = 0;
  for (long long t = blockIdx.y; t < n_tiles; t += gridDim.y, buf ^= 1) {
    const long long tn = t + gridDim.y;
    if (tid == 0 && tn < n_tiles) {          // prefetch the next agent tile into the other buffer
      mbar_expect_tx(&bar_d[buf ^ 1], d_bytes);
      tma_bulk_g2s(buf ? s_d0 : s_d1, scratch + (size_t)tn * T * BVC_AT, d_bytes, &bar_d[buf ^ 1]);
    }
    mbar_wait(&bar_d[buf], phase[buf]);
    phase[buf] ^= 1u;
    const float* sd = (buf ? s_d1 : s_d0) + 8 * g;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    // the window [th0, th0 + tlen) mod T as two plain segments (no wrap test in the loop); 4 angles in flight per thread:
    // the loop is bound by MUFU.EX2 (one warp instruction per 8 cycles), whose queue has to be kept fed across the
    // LDS -> FFMA -> FMUL head of every angle
    const int n0 = min(tlen, T - th0);
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int tb = seg ? 0 : th0, tn = seg ? tlen - n0 : n0;
      const float* pv = s_vm + tb * BVC_CT + cl;
      const float* pd = sd + tb * BVC_AT;
#pragma unroll 4
      for (int j = 0; j < tn; ++j, pv += BVC_CT, pd += BVC_AT) {
        const float vm = *pv;
        const float4 da = *reinterpret_cast<const float4*>(pd);
        const float4 db = *reinterpret_cast<const float4*>(pd + 4);
        const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float u = fmaf(dv[i], sc, -mc);                    // (d - mu_d) * s
          // gaussian * von Mises.  Moving some of the eight exponentials to the 11-instruction FMA-pipe form (ex2_fma,
          // -DRIAB_BVC_MUFU_TERMS=7) measured 545 vs 558 us/step on c4 in a --split-compile build but 572 vs 559 in the
          // default build (register allocation of the unrolled loop decides): all eight stay on MUFU.
          const float e = (i < BVC_MUFU_TERMS) ? ex2f(-u * u) : ex2_fma(-u * u);
          acc[i] = fmaf(e, vm, acc[i]);
        }
      }
    }
    if (cell < bc.n_cells) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long row = t * BVC_AT + 8 * g + i;
        // a NaN position makes every distance NaN: the reference returns zero rates for it (Neurons.py:163-164)
        v[i] = (acc[i] != acc[i]) ? 0.f : fmaf(acc[i] * scale, bc.span, bc.min_fr);
        if (row < n_rows) st_cs_f1(out.rates + row * out.ld + cell, v[i]);
      }
      if (fold_spikes) bvc_spikes8(out, v, t * BVC_AT + 8 * g, n_rows, cell);
    }
    __syncthreads();   // everyone done with this buffer before it is refilled two iterations later
  }
}

// BVC phase B, egocentric frame: same tiling as k_bvc_integrate, no von Mises table.
__global__ void __launch_bounds__(NT) k_bvc_integrate_ego(const BvcConst bc, const float* __restrict__ scratch,
                                                          const long long n_rows, const long long n_tiles,
                                                          const OutK out, const int fold_spikes) {
  extern __shared__ __align__(128) unsigned char dyn[];
  const int T = bc.T;
  float2* s_th = reinterpret_cast<float2*>(dyn);                    // [T] (cos, sin) of the test angles
  float* s_d0 = reinterpret_cast<float*>(dyn) + 2 * (size_t)((T + 1) / 2 * 2);   // [T][32] x 2 buffers
  float* s_d1 = s_d0 + (size_t)T * BVC_AT;
  __shared__ uint64_t bar_d[2];
  const int ct = blockIdx.x;
  const int tid = threadIdx.x, cl = tid & 63, g = tid >> 6;
  const int cell = ct * BVC_CT + cl;
  const uint32_t d_bytes = (uint32_t)T * BVC_AT * 4u;
  const float* base = bc.packed;
  const size_t np = (size_t)bc.n_pad;
  const float* ext = base + 3 * np + np * (size_t)T;               // kap | cmu | smu | cth | sth
  for (int i = tid; i < T; i += blockDim.x) s_th[i] = make_float2(ext[3 * np + i], ext[3 * np + T + i]);
  if (tid == 0) {
    mbar_init(&bar_d[0], 1); mbar_init(&bar_d[1], 1);
    mbar_fence_init();
    long long t0 = blockIdx.y;
    if (t0 < n_tiles) { mbar_expect_tx(&bar_d[0], d_bytes); tma_bulk_g2s(s_d0, scratch + (size_t)t0 * T * BVC_AT, d_bytes, &bar_d[0]); }
  }
  __syncthreads();
  const float sc = base[cell], mc = base[np + cell], scale = base[2 * np + cell];
  const float kap = ext[cell], cmu = ext[np + cell], smu = ext[2 * np + cell];
  uint32_t phase[2] = {0u, 0u};
  int buf = 0;
  for (long long t = blockIdx.y; t < n_tiles; t += gridDim.y, buf ^= 1) {
    const long long tn = t + gridDim.y;
    if (tid == 0 && tn < n_tiles) {
      mbar_expect_tx(&bar_d[buf ^ 1], d_bytes);
      tma_bulk_g2s(buf ? s_d0 : s_d1, scratch + (size_t)tn * T * BVC_AT, d_bytes, &bar_d[buf ^ 1]);
    }
    // (cos, sin)(head bearing + mu_theta) for this thread's 8 agents
    float cph[8], sph[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const long long row = t * BVC_AT + 8 * g + i;
      float hx = 1.0f + 1e-6f, hy = 0.f;                            // default head direction [1,0] (Neurons.py:1703)
      if (bc.head_dir != nullptr && row < n_rows) {
        hx = (float)(bc.head_dir[2 * row] + 1e-6);                  // utils.get_angle: atan2(y, x + eps)
        hy = (float)bc.head_dir[2 * row + 1];
      }
      const float rn = rsqrtf(fmaf(hx, hx, hy * hy));
      const float ch = hx * rn, sh = hy * rn;
      cph[i] = fmaf(-sh, smu, ch * cmu);
      sph[i] = fmaf(ch, smu, sh * cmu);
    }
    mbar_wait(&bar_d[buf], phase[buf]);
    phase[buf] ^= 1u;
    const float* sd = (buf ? s_d1 : s_d0) + 8 * g;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int th = 0; th < T; ++th) {
      const float2 cs = s_th[th];
      const float4 da = *reinterpret_cast<const float4*>(sd + th * BVC_AT);
      const float4 db = *reinterpret_cast<const float4*>(sd + th * BVC_AT + 4);
      const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float u = fmaf(dv[i], sc, -mc);                       // (d - mu_d) * s
        const float c = fmaf(cs.y, sph[i], cs.x * cph[i]);          // cos(theta - bearing - mu_theta)
        acc[i] += ex2f(fmaf(kap, c - 1.0f, -u * u));                // gaussian * von Mises in one ex2
      }
    }
    if (cell < bc.n_cells) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const long long row = t * BVC_AT + 8 * g + i;
        // a NaN position makes every distance NaN: the reference returns zero rates for it (Neurons.py:163-164)
        v[i] = (acc[i] != acc[i]) ? 0.f : fmaf(acc[i] * scale, bc.span, bc.min_fr);
        if (row < n_rows) st_cs_f1(out.rates + row * out.ld + cell, v[i]);
      }
      if (fold_spikes) bvc_spikes8(out, v, t * BVC_AT + 8 * g, n_rows, cell);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------
// History analytics: occupancy and rate-weighted histograms of the agents' positions (utils.py:544-589).
// np.histogram2d with explicit edges: bin = searchsorted(edges, x, 'right') - 1, the right-most edge belongs to
// the last bin, samples outside are dropped.
__device__ __forceinline__ int hist_bin(const double* __restrict__ e, int n_edges, double x) {
  if (!(x >= e[0]) || !(x <= e[n_edges - 1])) return -1;
  if (x == e[n_edges - 1]) return n_edges - 2;
  int lo = 0, hi = n_edges;                       // first edge > x
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= x) lo = mid + 1; else hi = mid;
  }
  return lo - 1;
}
__global__ void __launch_bounds__(NT) k_history_maps(const riab_history_view h, const double* __restrict__ ex, int nex,
                                                     const double* __restrict__ ey, int ney, float* __restrict__ sum,
                                                     float* __restrict__ count) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (NT / 32);
  const long long n_samples = h.n_steps * h.n_agents;
  const int ny = ney - 1;
  for (long long w = (long long)blockIdx.x * (NT / 32) + (threadIdx.x >> 5); w < n_samples; w += warps) {
    const long long step = w / h.n_agents, agent = w - step * h.n_agents;
    const long long arow = (h.agent_row0 + step) % h.agent_ring_rows;
    const float* p = h.agent_ring + (arow * h.n_agents + agent) * 8;
    const int ix = hist_bin(ex, nex, (double)p[0]), iy = hist_bin(ey, ney, (double)p[1]);
    if (ix < 0 || iy < 0) continue;                              // warp-uniform
    const long long bin = (long long)ix * ny + iy;
    if (lane == 0) atomicAdd(count + bin, 1.0f);
    if (h.rates_ring != nullptr) {
      const long long rrow = (h.rates_row0 + step) % h.rates_ring_rows;
      const float* r = h.rates_ring + (rrow * h.n_agents + agent) * h.ld;
      float* dst = sum + bin * h.ld;
      for (int c = lane; c < h.n_cells; c += 32) atomicAdd(dst + c, r[c]);
    }
  }
}

// ---------------------------------------------------------------------------
// host helpers
int make_env(const riab_env* env, EnvK& k) {
  if (env == nullptr || (env->walls_dev == nullptr && env->n_walls > 0)) return fail(RIAB_ERR_INVALID, "env / walls_dev is NULL");
  if (env->n_walls < 0 || env->n_walls > MAXW) return fail(RIAB_ERR_UNSUPPORTED, "n_walls=%d exceeds %d", env->n_walls, MAXW);
  if (env->n_boundary_walls < 0 || env->n_boundary_walls > env->n_walls)
    return fail(RIAB_ERR_INVALID, "n_boundary_walls=%d out of range", env->n_boundary_walls);
  k.walls = env->walls_dev; k.W = env->n_walls; k.nb = env->n_boundary_walls;
  k.aligned = (((uintptr_t)env->walls_dev) % 16 == 0) && env->n_walls > 0;
  for (int i = 0; i < 4; ++i) k.ext[i] = env->extent[i];
  k.cxm = 0.5 * (env->extent[0] + env->extent[1]);
  k.cym = 0.5 * (env->extent[2] + env->extent[3]);
  if (env->boundary_mode < 0 || env->boundary_mode > RIAB_BOUNDARY_POLYGON) return fail(RIAB_ERR_INVALID, "bad boundary_mode %d", env->boundary_mode);
  k.periodic = env->boundary_mode == RIAB_BOUNDARY_PERIODIC_BOX ? 1 : 0;
  k.polygon = env->boundary_mode == RIAB_BOUNDARY_POLYGON ? 1 : 0;
  k.nh = k.polygon ? env->n_hole_walls : 0;
  k.h0 = k.polygon ? env->hole_wall0 : 0;
  if (k.nh < 0 || k.h0 < 0 || (k.nh > 0 && (k.h0 < k.nb || k.h0 + k.nh > k.W)))
    return fail(RIAB_ERR_INVALID, "hole walls [%d, %d) out of range", env->hole_wall0, env->hole_wall0 + env->n_hole_walls);
  if (k.polygon && k.nb < 3) return fail(RIAB_ERR_INVALID, "a polygon boundary needs at least 3 boundary walls");
  k.scale = env->scale;
  if (k.periodic && !(env->scale > 0.0)) return fail(RIAB_ERR_INVALID, "periodic environment needs scale > 0");
  return 0;
}

int check_agents(const riab_agents* a) {
  if (a == nullptr) return fail(RIAB_ERR_INVALID, "agents is NULL");
  if (a->n_agents < 0) return fail(RIAB_ERR_INVALID, "n_agents < 0");
  if (a->n_agents > 0 && (!a->pos || !a->velocity || !a->rotational_velocity || !a->measured_velocity ||
                          !a->measured_rotational_velocity || !a->head_direction || !a->distance_travelled ||
                          !a->distance_to_closest_wall))
    return fail(RIAB_ERR_INVALID, "agents: NULL state array");
  return 0;
}

int check_motion(const riab_motion_params* p) {
  if (p == nullptr) return fail(RIAB_ERR_INVALID, "motion params NULL");
  if (!(p->dt > 0.0)) return fail(RIAB_ERR_INVALID, "dt must be > 0");
  return 0;
}

// fr_bound: an upper bound of the population's rates (thinned spikes), negative when there is none
int make_out(const riab_rates_out* o, const riab_neuron_noise* nz, int n_cells, double dt, long long id_offset, OutK& k,
             double fr_bound = -1.0) {
  if (o == nullptr || o->rates_row == nullptr) return fail(RIAB_ERR_INVALID, "rates_row is NULL");
  if (o->ld < n_cells) return fail(RIAB_ERR_INVALID, "ld (%lld) < n_cells (%d)", (long long)o->ld, n_cells);
  memset(&k, 0, sizeof(k));
  k.rates = o->rates_row; k.ld = o->ld;
  k.spikes = o->spikes_row; k.spike_ld = 4 * ((n_cells + 127) / 128);     // 4 ballot words per 128 cells
  if (k.spikes != nullptr && (((uintptr_t)k.spikes) % 16 != 0)) return fail(RIAB_ERR_INVALID, "spikes_row must be 16-byte aligned");
  k.noise = nullptr;
  k.dt = (float)dt;
  k.id_offset = id_offset;
  if (nz != nullptr) {
    k.seed = nz->seed; k.step = nz->step; k.pop = nz->population_id;
    if (nz->noise_std != 0.f) {
      if (o->noise_state == nullptr) return fail(RIAB_ERR_INVALID, "noise_std != 0 needs noise_state");
      k.noise = o->noise_state;
      const double tau = nz->noise_coherence_time;
      k.noise_decay = (float)(dt / tau);
      k.noise_sig = (float)(sqrt(2.0 * (double)nz->noise_std * nz->noise_std / (tau * dt)) * dt);
    }
  } else if (o->spikes_row != nullptr) {
    return fail(RIAB_ERR_INVALID, "spikes need a riab_neuron_noise (seed/step)");
  }
  k.vec_ok = (o->ld % 4 == 0) && (((uintptr_t)o->rates_row) % 16 == 0);
  {
    uint32_t k0 = (uint32_t)k.seed, k1 = (uint32_t)(k.seed >> 32);
    for (int i = 0; i < 7; ++i) { k.rk7[2 * i] = k0; k.rk7[2 * i + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
  }
  // thinned spikes (thin_block): p' = dt * bound * (1 + 2^-10) -- the margin covers the rates' float32 rounding above `bound`.
  // The default for bounded populations without OU noise when p' <= 1/16 (beyond that the candidates are no rarer than the
  // dense stream's work); RIAB_DENSE_SPIKES=1 in the environment keeps the dense stream everywhere.
  k.thin = 0;
  if (k.spikes != nullptr && k.noise == nullptr && fr_bound >= 0.0 && getenv("RIAB_DENSE_SPIKES") == nullptr) {
    const double bound = fr_bound * (1.0 + 1.0 / 1024.0), p = dt * bound;
    if (p > 0.0 && p <= 0.0625) {
      k.thin = 1;
      // Binomial(128, p) by the pmf recurrence, IEEE operations in this order (tests/philox_np.py: thin_tables repeats them)
      const double q = 1.0 - p, r = p / q;
      double pmf = q;
      for (int i = 0; i < 7; ++i) pmf = pmf * pmf;                       // q^128
      double cdf = 0.0;
      for (int i = 0; i < 32; ++i) {
        cdf = cdf + pmf;
        const double t = floor(4294967296.0 * cdf);
        k.thin_cdf[i] = (t >= 4294967295.0) ? 4294967295u : (uint32_t)t;
        pmf = pmf * ((double)(128 - i) * r) / (double)(i + 1);
      }
      k.thin_c1 = (float)(bound * (1.0 / 1048576.0));
      k.thin_c0 = (float)(bound * (1.0 / 2097152.0));
    }
  }
  return 0;
}

int make_place(const riab_place_cells* pc, const EnvK& env, PlaceConst& c) {
  if (pc == nullptr || pc->packed_dev == nullptr) return fail(RIAB_ERR_INVALID, "place cells / packed_dev NULL");
  if (pc->description < 0 || pc->description > RIAB_PC_ONE_HOT) return fail(RIAB_ERR_INVALID, "bad description %d", pc->description);
  if (pc->wall_geometry < 0 || pc->wall_geometry > RIAB_GEOM_GEODESIC) return fail(RIAB_ERR_INVALID, "bad wall_geometry");
  // Environment.py:715-717 hard-codes `walls[4:]`: with a polygon boundary the walls after the first FOUR count as
  // "inner" walls whatever the polygon's vertex count; the box has exactly its 4 boundary walls first
  const int skip = env.polygon ? (env.W < 4 ? env.W : 4) : env.nb;
  const int n_inner = env.W - skip;
  c.wall0 = skip;
  if (pc->wall_geometry != RIAB_GEOM_EUCLIDEAN) {
    if (pc->n_inner_walls != n_inner) return fail(RIAB_ERR_INVALID, "packed for %d inner walls, env has %d (re-pack after add_wall)", pc->n_inner_walls, n_inner);
    if (n_inner > PLACE_MAX_WI) return fail(RIAB_ERR_UNSUPPORTED, "line_of_sight supports at most %d inner walls (got %d)", PLACE_MAX_WI, n_inner);
    if (pc->centres_dev == nullptr) return fail(RIAB_ERR_INVALID, "centres_dev NULL");
    if (pc->wall_geometry == RIAB_GEOM_GEODESIC && n_inner > 1)
      return fail(RIAB_ERR_INVALID, "geodesic geometry is only defined with one additional wall (Environment.py:736-739)");
  }
  if ((pc->description == RIAB_PC_TOP_HAT || pc->description == RIAB_PC_ONE_HOT) && pc->centres_dev == nullptr)
    return fail(RIAB_ERR_INVALID, "centres_dev NULL");
  c.desc = pc->description; c.geometry = pc->wall_geometry; c.n_cells = pc->n_cells; c.n_pad = pc->n_pad;
  c.n_inner = (pc->wall_geometry == RIAB_GEOM_EUCLIDEAN) ? 0 : n_inner;
  c.ep_valid = pc->ep_valid;
  c.min_fr = pc->min_fr; c.span = pc->max_fr - pc->min_fr;
  c.top_hat_w = pc->top_hat_width; c.top_hat_w2 = (float)(pc->top_hat_width * pc->top_hat_width);
  c.band = 0.f;
  for (int j = 0; j < PLACE_MAX_WI; ++j) { c.eps[j] = pc->eps[j]; c.band = fmaxf(c.band, pc->eps[j]); }
  c.packed = pc->packed_dev; c.centres64 = pc->centres_dev;
  c.cxm = env.cxm; c.cym = env.cym;
  // expanded Gaussian (place_rates4): one common width, plain Gaussian, no wrap-around, and small enough
  // exponents at the far corner that the float32 cancellation stays below 4e-6 relative
  c.expanded = (pc->description == RIAB_PC_GAUSSIAN && pc->wall_geometry != RIAB_GEOM_GEODESIC && !env.periodic &&
                pc->k_uniform > 0.f && pc->k_uniform * pc->r2_max <= 10.0f) ? 1 : 0;
  c.kx = -pc->k_uniform;
  c.fold = (c.expanded && pc->min_fr == 0.f && c.span > 0.f) ? 1 : 0;
  c.lspan = c.fold ? log2f(c.span) : 0.f;
  c.periodic = env.periodic; c.scale = env.scale; c.scale_f = (float)env.scale; c.half_f = (float)(env.scale / 2);
  if (env.periodic && pc->wall_geometry != RIAB_GEOM_EUCLIDEAN)
    return fail(RIAB_ERR_INVALID, "line_of_sight / geodesic wall geometry only possible when the boundary conditions are solid (Neurons.py:907-921)");
  return 0;
}

int make_grid(const riab_grid_cells* gc, const EnvK& env, GridConst& c) {
  if (gc == nullptr || gc->packed_dev == nullptr) return fail(RIAB_ERR_INVALID, "grid cells / packed_dev NULL");
  c.n_cells = gc->n_cells; c.n_pad = gc->n_pad;
  if (gc->description == RIAB_GC_RECTIFIED_COSINES) {
    if (!(gc->width_ratio > 0.0 && gc->width_ratio <= 1.0)) return fail(RIAB_ERR_INVALID, "width_ratio must be between 0 and 1");
    const double full = (1.0 / 3.0) * (2.0 * cos(sqrt(3.0) * M_PI * gc->width_ratio / 2.0) + 1.0);   // Neurons.py:1211
    c.A = (float)((1.0 / 3.0) / (1.0 - full));
    c.B = (float)(-full / (1.0 - full));
    c.rectify = 1;
  } else if (gc->description == RIAB_GC_SHIFTED_COSINES) {
    c.A = (float)(2.0 / 9.0); c.B = (float)(1.0 / 3.0); c.rectify = 0;                                 // Neurons.py:1216-1218
  } else return fail(RIAB_ERR_INVALID, "bad grid description %d", gc->description);
  c.min_fr = gc->min_fr; c.span = gc->max_fr - gc->min_fr;
  c.As = c.A * c.span; c.Bs = fmaf(c.B, c.span, c.min_fr);
  c.clamp = c.rectify ? (c.span >= 0.f ? 1 : 2) : 0;
  c.packed = gc->packed_dev; c.cxm = env.cxm; c.cym = env.cym;
  return 0;
}

int make_ovc(const riab_ovc_cells* oc, const EnvK& env, const double* head_dir, OvcConst& c) {
  if (oc == nullptr || oc->packed_dev == nullptr) return fail(RIAB_ERR_INVALID, "object vector cells / packed_dev NULL");
  if (oc->n_objects < 0 || oc->n_objects > RIAB_MAX_OBJECTS)
    return fail(RIAB_ERR_UNSUPPORTED, "n_objects=%d exceeds %d", oc->n_objects, RIAB_MAX_OBJECTS);
  if (env.periodic) return fail(RIAB_ERR_UNSUPPORTED, "object vector cells need solid boundary conditions here");
  memset(&c, 0, sizeof(c));
  c.n_cells = oc->n_cells; c.n_pad = oc->n_pad; c.n_obj = oc->n_objects;
  c.ego = oc->egocentric ? 1 : 0; c.occlude = oc->walls_occlude ? 1 : 0;
  c.wall0 = env.W < 4 ? env.W : 4;                       // Environment.py:715-717: walls[4:]
  c.n_inner = env.W - c.wall0;
  c.min_fr = oc->min_fr; c.span = oc->max_fr - oc->min_fr;
  c.packed = oc->packed_dev; c.head_dir = head_dir;
  for (int o = 0; o < oc->n_objects; ++o) {
    c.obj[2 * o] = oc->objects[2 * o]; c.obj[2 * o + 1] = oc->objects[2 * o + 1];
    c.type[o] = (float)oc->object_types[o];
  }
  return 0;
}

int g_num_sms = 0;

// MODE 0: rates for given positions; 1: motion -> rates (one step); 2: skewed (rates of the current
// positions, then motion for the NEXT step -- used inside riab_run).
template <class P, int MODE>
int launch_tile(const EnvK& env, const riab_agents& ag, const riab_motion_params& mp, const riab_step_io& io,
                const typename P::Const& pc, const OutK& out_in, const double* pos_in, long long n_rows, cudaStream_t s,
                const RunK* run_in = nullptr) {
  if (n_rows == 0) return 0;
  {
    static int sms_of[64] = {0};                      // SM count per device ordinal (a process may drive several GPUs)
    int dev = 0;
    RIAB_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(RIAB_ERR_UNSUPPORTED, "device ordinal %d", dev);
    if (sms_of[dev] == 0) RIAB_CUDA_OK(cudaDeviceGetAttribute(&sms_of[dev], cudaDevAttrMultiProcessorCount, dev));
    g_num_sms = sms_of[dev];
  }
  const bool spikes = out_in.spikes != nullptr, noise = out_in.noise != nullptr;
  // thinned stream: bounded rates AND a policy for which it measured faster (P::THIN); see the policies
  const bool thin = spikes && !noise && out_in.thin && P::THIN;
  // warp-role configuration (see StepCfg)
  const int cfg = (noise || thin || (spikes && !P::LIGHT)) ? 4 : (spikes ? 12 : ((P::LIGHT && MODE != 0) ? 8 : 12));
  // agents per ring slot: 32 for large batches; small ones get equal shares per (CTA, consumer group)
  OutK outk = out_in;
  {
    const int ct = pc.n_pad / P::CPT;                                 // cell-threads of one consumer group
    const int ring = (cfg == 4) ? ring_slots<P, StepCfg<4>>() : (cfg == 8) ? ring_slots<P, StepCfg<8>>() : ring_slots<P, StepCfg<12>>();
    const long long groups = (ct > 0 && ct <= RW * 32) ? (long long)g_num_sms * lean_groups(ct, ring) : (long long)g_num_sms;
    int ta = TA;
    if (n_rows < 4ll * TA * groups) {
      const long long per = (n_rows + groups - 1) / groups;          // agents per group if every group gets one slot
      const long long rounds = (per + TA - 1) / TA;                   // slots per group
      ta = (int)((n_rows + groups * rounds - 1) / (groups * rounds));
      ta = (ta + 1) & ~1;
      if (ta < 2) ta = 2;
      if (ta > TA) ta = TA;
    }
    outk.tile_agents = ta;
  }
  const OutK& out = outk;
  RunK run;
  memset(&run, 0, sizeof(run));
  if (run_in != nullptr) run = *run_in;
  const long long n_tiles = (n_rows + out.tile_agents - 1) / out.tile_agents;
  const unsigned grid = (unsigned)(n_tiles < g_num_sms ? n_tiles : g_num_sms);
  MotionDerived md;
  memset(&md, 0, sizeof(md));
  if (MODE != 0) derive_motion(mp, md);
  if (noise) k_step<P, MODE, 1, true, StepCfg<4>><<<grid, StepCfg<4>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
  else if (thin) {
    if constexpr (P::THIN) k_step<P, MODE, 2, false, StepCfg<4>><<<grid, StepCfg<4>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
  }
  else if (spikes) {
    // dense stream: light consumers (Euclidean Gaussian place cells) are producer-bound next to 4 producer warps (c2e 59 us),
    // with 8 they run at 56 us; the heavier loops keep the 104-register consumers
    if constexpr (P::LIGHT) k_step<P, MODE, 1, false, StepCfg<12>><<<grid, StepCfg<12>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
    else k_step<P, MODE, 1, false, StepCfg<4>><<<grid, StepCfg<4>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
  }
  else {
    // light consumers without spikes run at the HBM write rate: fat producers (only instantiated for them)
    if constexpr (P::LIGHT && MODE != 0) k_step<P, MODE, 0, false, StepCfg<8>><<<grid, StepCfg<8>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
    else k_step<P, MODE, 0, false, StepCfg<12>><<<grid, StepCfg<12>::THREADS, 0, s>>>(env, ag, mp, md, io, pc, out, pos_in, n_rows, run);
  }
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int MODE, int DESC>
int launch_place_d(const EnvK& env, const riab_agents& ag, const riab_motion_params& mp, const riab_step_io& io,
                   const PlaceConst& pc, const OutK& out, const double* pos_in, long long n_rows, cudaStream_t s,
                   const RunK* run = nullptr) {
  const int wi = pc.n_inner;
  if (wi == 0) return launch_tile<PlacePolicy<0, DESC>, MODE>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
  if (wi == 1) return launch_tile<PlacePolicy<1, DESC>, MODE>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
  if (wi == 2) return launch_tile<PlacePolicy<2, DESC>, MODE>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
  if (wi <= 4) return launch_tile<PlacePolicy<4, DESC>, MODE>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
  return launch_tile<PlacePolicy<8, DESC>, MODE>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
}

int launch_onehot(const EnvK& env, const PlaceConst& pc, const OutK& out, const double* pos, long long n_rows,
                  cudaStream_t s) {
  if (n_rows == 0) return 0;
  k_place_onehot<<<(unsigned)((n_rows + NT / 32 - 1) / (NT / 32)), NT, 0, s>>>(env, pc, pos, n_rows, out);
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  if (out.noise != nullptr || out.spikes != nullptr) {
    const int np128 = (pc.n_cells + CELL_PAD - 1) / CELL_PAD * CELL_PAD;
    k_finish_rows<<<dim3((unsigned)n_rows, (unsigned)((np128 / 4 + NT - 1) / NT)), NT, 0, s>>>(out, pc.n_cells, np128, n_rows);
    g_launches++;
    RIAB_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

template <int MODE>
int launch_place(const EnvK& env, const riab_agents& ag, const riab_motion_params& mp, const riab_step_io& io,
                 const PlaceConst& pc, const OutK& out, const double* pos_in, long long n_rows, cudaStream_t s,
                 const RunK* run = nullptr) {
  if (pc.desc == RIAB_PC_ONE_HOT) {
    if (MODE != 0) return fail(RIAB_ERR_INVALID, "one_hot is launched unfused");
    return launch_onehot(env, pc, out, pos_in, n_rows, s);
  }
  // the common Gaussian profile without geodesic detours gets a compile-time specialisation
  if (pc.desc == RIAB_PC_GAUSSIAN && pc.geometry != RIAB_GEOM_GEODESIC)
    return launch_place_d<MODE, RIAB_PC_GAUSSIAN>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
  return launch_place_d<MODE, -1>(env, ag, mp, io, pc, out, pos_in, n_rows, s, run);
}

// riab_run pipelines BoundaryVectorCells across steps: the float64 ray kernel of step s+1 (latency-bound, FP64 pipe) runs
// on the caller's stream while the angular integral of step s (MUFU-bound) still runs on a side stream -- different pipes,
// so the two overlap almost completely.  Needs a second ray-distance buffer (steps alternate) and, per buffer, an event
// that the integral which read it has finished.  Allocentric cells only (egocentric ones read the head directions the next
// motion step overwrites).
constexpr int PIPE_POPS = 8;
struct BvcPipe {
  cudaStream_t side = nullptr;
  cudaEvent_t rays_done = nullptr, int_done[2][PIPE_POPS] = {};
  bool used[2][PIPE_POPS] = {};
  float* scratch2[PIPE_POPS] = {};
  long long step = 0;
};
thread_local BvcPipe* g_pipe = nullptr;

template <bool FUSED>
int launch_bvc(const EnvK& env, const riab_agents& ag, const riab_motion_params& mp, const riab_step_io& io,
               const riab_bvc_cells* bvc, const OutK& out, const double* pos_in, long long n_rows, float* scratch,
               int32_t* first_wall, const double* head_dir, cudaStream_t s) {
  if (bvc == nullptr || bvc->packed_dev == nullptr || bvc->test_dirs_dev == nullptr)
    return fail(RIAB_ERR_INVALID, "bvc cells / packed_dev / test_dirs_dev NULL");
  if (scratch == nullptr) return fail(RIAB_ERR_INVALID, "bvc scratch NULL");
  if (env.periodic) return fail(RIAB_ERR_INVALID, "boundary cells only possible with solid boundary conditions (Neurons.py:1580-1582)");
  if (n_rows == 0) return 0;
  BvcConst bc;
  bc.n_cells = bvc->n_cells; bc.n_pad = bvc->n_pad; bc.T = bvc->n_test_angles;
  bc.min_fr = bvc->min_fr; bc.span = bvc->max_fr - bvc->min_fr;
  bc.packed = bvc->packed_dev; bc.test_dirs = bvc->test_dirs_dev;
  bc.ego = bvc->egocentric; bc.head_dir = head_dir;
  const long long n_tiles = (n_rows + BVC_AT - 1) / BVC_AT;
  const size_t smemA = (size_t)bc.T * 2 * sizeof(double);
  const size_t smemB = (size_t)bc.T * (BVC_CT + 2 * BVC_AT) * sizeof(float);
  if (smemB > 220 * 1024) return fail(RIAB_ERR_UNSUPPORTED, "n_test_angles=%d too large for shared memory", bc.T);
  if ((bc.T * BVC_AT * 4) % 16 != 0 || ((uintptr_t)scratch) % 16 != 0 || ((uintptr_t)bvc->packed_dev) % 16 != 0)
    return fail(RIAB_ERR_INVALID, "bvc buffers must be 16-byte aligned");
  const bool rec = FUSED && (io.collision_mask || io.first_hit || io.n_iters);
  MotionDerived md;
  memset(&md, 0, sizeof(md));
  if (FUSED) derive_motion(mp, md);
  // spikes without OU noise are drawn in the integration kernel's epilogue (the ray kernel clears the rows first);
  // OU noise (a read-modify-write of the noise state per rate) and odd shard offsets keep the k_finish_rows post-pass
  const int fold = (out.spikes != nullptr && out.noise == nullptr && (out.id_offset & 1ll) == 0) ? 1 : 0;
  uint32_t* const zsp = fold ? out.spikes : nullptr;
  BvcPipe* const pipe = (g_pipe != nullptr && !bc.ego && out.pop >= 0 && out.pop < PIPE_POPS && g_pipe->scratch2[out.pop] != nullptr) ? g_pipe : nullptr;
  const int pb = pipe ? (int)(pipe->step & 1) : 0;
  if (pipe) {
    if (pb) scratch = pipe->scratch2[out.pop];
    // the integral of two steps ago read this buffer (and wrote the ring slot a short ring re-uses now)
    if (pipe->used[pb][out.pop]) RIAB_CUDA_OK(cudaStreamWaitEvent(s, pipe->int_done[pb][out.pop], 0));
  }
  // (angle, wall) table of the float32 screen in shared memory: up to BVC_NW walls and 40 KB (one ray CTA still fits next to
  // two integration CTAs of the previous step, 2 x 92 KB at T = 180)
  const size_t smemT = smemA + (size_t)bc.T * (sizeof(float2) + (size_t)env.W * sizeof(BvcTab));
  const bool table = env.W <= BVC_NW && smemT <= 40 * 1024;
  if (table) {
    if (rec) k_bvc_rays<FUSED, FUSED, true><<<(unsigned)n_tiles, NT, smemT, s>>>(env, ag, mp, md, io, bc, pos_in, n_rows, scratch, first_wall, zsp, out.spike_ld);
    else k_bvc_rays<FUSED, false, true><<<(unsigned)n_tiles, NT, smemT, s>>>(env, ag, mp, md, io, bc, pos_in, n_rows, scratch, first_wall, zsp, out.spike_ld);
  } else {
    if (rec) k_bvc_rays<FUSED, FUSED, false><<<(unsigned)n_tiles, NT, smemA, s>>>(env, ag, mp, md, io, bc, pos_in, n_rows, scratch, first_wall, zsp, out.spike_ld);
    else k_bvc_rays<FUSED, false, false><<<(unsigned)n_tiles, NT, smemA, s>>>(env, ag, mp, md, io, bc, pos_in, n_rows, scratch, first_wall, zsp, out.spike_ld);
  }
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  if (pipe) {                                         // the integral (and its post-pass) go to the side stream
    RIAB_CUDA_OK(cudaEventRecord(pipe->rays_done, s));
    RIAB_CUDA_OK(cudaStreamWaitEvent(pipe->side, pipe->rays_done, 0));
    s = pipe->side;
  }
  // (the attribute belongs to the current device's function image: set per call, a single process may drive several GPUs)
  RIAB_CUDA_OK(cudaFuncSetAttribute(k_bvc_integrate, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  int dev = 0, sms = 148;
  RIAB_CUDA_OK(cudaGetDevice(&dev));
  RIAB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned cts = (unsigned)(bc.n_pad / BVC_CT);
  unsigned gy = (unsigned)((2 * sms + cts - 1) / cts);           // ~2 CTAs per SM in total
  if (gy > n_tiles) gy = (unsigned)n_tiles;
  if (gy < 1) gy = 1;
  if (bc.ego) {
    RIAB_CUDA_OK(cudaFuncSetAttribute(k_bvc_integrate_ego, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    const size_t smemE = ((size_t)((bc.T + 1) / 2 * 2) * 2 + (size_t)bc.T * 2 * BVC_AT) * sizeof(float);
    k_bvc_integrate_ego<<<dim3(cts, gy), NT, smemE, s>>>(bc, scratch, n_rows, n_tiles, out, fold);
  } else {
    k_bvc_integrate<<<dim3(cts, gy), NT, smemB, s>>>(bc, scratch, n_rows, n_tiles, out, fold);
  }
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  if (out.noise != nullptr || (out.spikes != nullptr && !fold)) {
    const int np128 = (bc.n_cells + CELL_PAD - 1) / CELL_PAD * CELL_PAD;
    k_finish_rows<<<dim3((unsigned)n_rows, (unsigned)((np128 / 4 + NT - 1) / NT)), NT, 0, s>>>(out, bc.n_cells, np128, n_rows);
    g_launches++;
    RIAB_CUDA_OK(cudaGetLastError());
  }
  if (pipe) {
    RIAB_CUDA_OK(cudaEventRecord(pipe->int_done[pb][out.pop], s));
    pipe->used[pb][out.pop] = true;
  }
  return 0;
}

}  // namespace

// ===========================================================================
extern "C" {

int riab_abi_version(void) { return RIAB_ABI_VERSION; }
const char* riab_last_error(void) { return g_err; }
int riab_history_rate_maps(const riab_history_view* h, const double* edges_x_dev, int32_t n_edges_x,
                           const double* edges_y_dev, int32_t n_edges_y, float* sum_dev, float* count_dev, void* stream) {
  if (h == nullptr || h->agent_ring == nullptr || edges_x_dev == nullptr || edges_y_dev == nullptr || count_dev == nullptr)
    return fail(RIAB_ERR_INVALID, "riab_history_rate_maps: NULL argument");
  if (n_edges_x < 2 || n_edges_y < 2 || h->agent_ring_rows <= 0 || h->n_steps < 0 || h->n_agents <= 0)
    return fail(RIAB_ERR_INVALID, "riab_history_rate_maps: bad sizes");
  if (h->rates_ring != nullptr && (sum_dev == nullptr || h->rates_ring_rows <= 0 || h->ld < h->n_cells))
    return fail(RIAB_ERR_INVALID, "riab_history_rate_maps: rates ring without sum_dev / ld < n_cells");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bins = (size_t)(n_edges_x - 1) * (size_t)(n_edges_y - 1);
  RIAB_CUDA_OK(cudaMemsetAsync(count_dev, 0, bins * sizeof(float), s));
  if (h->rates_ring != nullptr) RIAB_CUDA_OK(cudaMemsetAsync(sum_dev, 0, bins * (size_t)h->ld * sizeof(float), s));
  const long long n_samples = h->n_steps * h->n_agents;
  if (n_samples == 0) return 0;
  long long blocks = (n_samples + NT / 32 - 1) / (NT / 32);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_history_maps<<<(unsigned)blocks, NT, 0, s>>>(*h, edges_x_dev, n_edges_x, edges_y_dev, n_edges_y, sum_dev, count_dev);
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  return 0;
}

int64_t riab_launch_count(void) { return (int64_t)g_launches.load(); }
int riab_stream_synchronize(void* stream) {
  RIAB_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return 0;
}

int riab_agent_update(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                      const riab_step_io* io, void* stream) {
  EnvK ek;
  int rc;
  if ((rc = check_agents(agents)) || (rc = make_env(env, ek)) || (rc = check_motion(prm))) return rc;
  if (io == nullptr) return fail(RIAB_ERR_INVALID, "io is NULL");
  if (agents->n_agents == 0) return 0;
  const unsigned grid = (unsigned)((agents->n_agents + 127) / 128);
  const bool rec = io->collision_mask || io->first_hit || io->n_iters;
  MotionDerived md;
  derive_motion(*prm, md);
  if (rec) k_agent_update<true><<<grid, 128, 0, (cudaStream_t)stream>>>(*agents, *prm, md, *io, ek);
  else k_agent_update<false><<<grid, 128, 0, (cudaStream_t)stream>>>(*agents, *prm, md, *io, ek);
  g_launches++;
  RIAB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ----------------------------------------------------------------- PlaceCells
static int place_n_pad(int n) { return (n + CELL_PAD - 1) / CELL_PAD * CELL_PAD; }

int64_t riab_place_pack_floats(int32_t n_cells, int32_t n_inner_walls) {
  return (int64_t)place_n_pad(n_cells) * (4 + 2 * (n_inner_walls > 0 ? n_inner_walls : 0) + 2);
}

int riab_place_pack(const double* centres, const double* widths, int32_t n, const double* walls, int32_t n_walls,
                    int32_t n_boundary, const double* extent, int32_t geometry, riab_place_cells* meta, float* out) {
  if (!centres || !widths || !extent || !meta || !out || n <= 0) return fail(RIAB_ERR_INVALID, "riab_place_pack: bad argument");
  const int np = place_n_pad(n);
  const int n_inner = (geometry == RIAB_GEOM_EUCLIDEAN) ? 0 : (n_walls - n_boundary);
  if (n_inner < 0) return fail(RIAB_ERR_INVALID, "n_walls < n_boundary_walls");
  if (n_inner > 0 && !walls) return fail(RIAB_ERR_INVALID, "walls NULL");
  const double cxm = 0.5 * (extent[0] + extent[1]), cym = 0.5 * (extent[2] + extent[3]);
  const int64_t total = riab_place_pack_floats(n, n_inner);
  for (int64_t i = 0; i < total; ++i) out[i] = 0.f;
  float *cx = out, *cy = out + np, *kk = out + 2 * np;
  for (int i = 0; i < np; ++i) {
    if (i < n) {
      cx[i] = (float)(centres[2 * i] - cxm);
      cy[i] = (float)(centres[2 * i + 1] - cym);
      kk[i] = (float)(1.4426950408889634 / (2.0 * widths[i] * widths[i]));   // log2(e) / (2 w^2)
    } else { cx[i] = 1.0e3f; cy[i] = 1.0e3f; kk[i] = 0.f; }
  }
  {
    bool uniform = true;
    for (int i = 1; i < n; ++i) uniform = uniform && (widths[i] == widths[0]);
    const double hx = 0.5 * (extent[1] - extent[0]), hy = 0.5 * (extent[3] - extent[2]);
    double r2 = hx * hx + hy * hy;
    for (int i = 0; i < n; ++i) {
      const double c2 = (double)cx[i] * cx[i] + (double)cy[i] * cy[i];
      if (c2 > r2) r2 = c2;
    }
    meta->k_uniform = uniform ? kk[0] : 0.f;
    meta->r2_max = (float)r2;
    float* aa = out + 3 * (size_t)np;                     // -k |c|^2 of the float32-rounded centre
    for (int i = 0; i < np; ++i) aa[i] = (i < n && uniform) ? (float)(-(double)kk[0] * ((double)cx[i] * cx[i] + (double)cy[i] * cy[i])) : -1.0e5f;
  }
  meta->n_pad = np;
  meta->n_inner_walls = n_inner;
  meta->ep_valid = 0;
  for (int j = 0; j < 8; ++j) meta->eps[j] = 0.f;
  for (int j = 0; j < n_inner; ++j) {
    const double* w = walls + 4 * (n_boundary + j);
    float* fc = out + (size_t)(4 + 2 * j) * np;
    float* tc = out + (size_t)(5 + 2 * j) * np;
    double tmax = 1.0, dmax = 0.0;
    for (int cxi = 0; cxi < 2; ++cxi)
      for (int cyi = 0; cyi < 2; ++cyi) {            // agents live inside the box: bound |f|, |t| over its corners
        double f, t;
        wall_coords(extent[cxi], extent[2 + cyi], w[0], w[1], w[2], w[3], f, t);
        if (fabs(t) + 1.0 > tmax) tmax = fabs(t) + 1.0;
        if (fabs(f) > dmax) dmax = fabs(f);
      }
    double fcmax = 0.0;
    for (int i = 0; i < n; ++i) {
      double f, t;
      wall_coords(centres[2 * i], centres[2 * i + 1], w[0], w[1], w[2], w[3], f, t);
      if (fabs(t) + 1.0 > tmax) tmax = fabs(t) + 1.0;
      if (fabs(f) > fcmax) fcmax = fabs(f);
    }
    const double band = 2.0e-6 * (dmax + fcmax) * tmax;   // ~10x the float32 rounding error of M' and |D|
    for (int i = 0; i < np; ++i) {
      if (i < n) {
        double f, t;
        wall_coords(centres[2 * i], centres[2 * i + 1], w[0], w[1], w[2], w[3], f, t);
        fc[i] = (float)f; tc[i] = (float)t;
        // a centre (numerically) on the wall's line: (0,0) makes M' = 0, inside the band -> exact float64 path
        if (fabs(f) < 1.0e-6 * (dmax + fcmax)) { fc[i] = 0.f; tc[i] = 0.f; }
      } else { fc[i] = 1.f; tc[i] = -1.f; }
    }
    if (j < 8) meta->eps[j] = (float)band;
  }
  if (geometry == RIAB_GEOM_GEODESIC && n_inner >= 1) {
    const double* w = walls + 4 * n_boundary;
    float* ce0 = out + (size_t)(4 + 2 * n_inner) * np;
    float* ce1 = out + (size_t)(5 + 2 * n_inner) * np;
    for (int e = 0; e < 2; ++e) {
      const double ex = w[2 * e], ey = w[2 * e + 1];
      if (ex > extent[0] && ex < extent[1] && ey > extent[2] && ey < extent[3]) meta->ep_valid |= (1 << e);
    }
    for (int i = 0; i < n; ++i) {
      const double a = centres[2 * i] - w[0], b = centres[2 * i + 1] - w[1];
      const double c = centres[2 * i] - w[2], d = centres[2 * i + 1] - w[3];
      ce0[i] = (float)sqrt(a * a + b * b);
      ce1[i] = (float)sqrt(c * c + d * d);
    }
  }
  return 0;
}

int riab_place_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_place_cells* pc,
                     float* out_dev, int64_t ld_out, void* stream) {
  if (n_pos == 0) return 0;
  EnvK ek;
  PlaceConst c;
  OutK ok;
  int rc;
  if ((rc = make_env(env, ek)) || (rc = make_place(pc, ek, c))) return rc;
  if (n_pos > 0 && pos_dev == nullptr) return fail(RIAB_ERR_INVALID, "pos_dev NULL");
  riab_rates_out ro;
  memset(&ro, 0, sizeof(ro));
  ro.rates_row = out_dev; ro.ld = ld_out;
  if ((rc = make_out(&ro, nullptr, pc->n_cells, 1.0, 0, ok))) return rc;
  riab_agents ag; memset(&ag, 0, sizeof(ag));
  riab_motion_params mp; memset(&mp, 0, sizeof(mp));
  riab_step_io io; memset(&io, 0, sizeof(io));
  return launch_place<0>(ek, ag, mp, io, c, ok, pos_dev, n_pos, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ GridCells
int64_t riab_grid_pack_floats(int32_t n_cells) { return (int64_t)place_n_pad(n_cells) * 9; }

int riab_grid_pack(const double* gridscales, const double* phase_offsets, const double* w, int32_t n,
                   const double* extent, riab_grid_cells* meta, float* out) {
  if (!gridscales || !phase_offsets || !w || !extent || !meta || !out || n <= 0)
    return fail(RIAB_ERR_INVALID, "riab_grid_pack: bad argument");
  const int np = place_n_pad(n);
  const double cxm = 0.5 * (extent[0] + extent[1]), cym = 0.5 * (extent[2] + extent[3]);
  for (int64_t i = 0; i < (int64_t)np * 9; ++i) out[i] = 0.f;
  for (int i = 0; i < n; ++i) {
    const double kappa = (2.0 * M_PI) / gridscales[i];
    // origin = gridscale * phase_offset / (2 pi)   (Neurons.py:1191)
    const double ox = gridscales[i] * phase_offsets[2 * i] / (2.0 * M_PI) - cxm;
    const double oy = gridscales[i] * phase_offsets[2 * i + 1] / (2.0 * M_PI) - cym;
    for (int k = 0; k < 3; ++k) {
      const double wx = w[6 * i + 2 * k], wy = w[6 * i + 2 * k + 1];
      double ph = kappa * (ox * wx + oy * wy);
      ph = remainder(ph, 2.0 * M_PI);
      out[(size_t)(3 * k + 0) * np + i] = (float)(kappa * wx);
      out[(size_t)(3 * k + 1) * np + i] = (float)(kappa * wy);
      out[(size_t)(3 * k + 2) * np + i] = (float)ph;
    }
  }
  meta->n_pad = np;
  return 0;
}

int riab_grid_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_grid_cells* gc,
                    float* out_dev, int64_t ld_out, void* stream) {
  if (n_pos == 0) return 0;
  EnvK ek;
  GridConst c;
  OutK ok;
  int rc;
  if ((rc = make_env(env, ek)) || (rc = make_grid(gc, ek, c))) return rc;
  if (n_pos > 0 && pos_dev == nullptr) return fail(RIAB_ERR_INVALID, "pos_dev NULL");
  riab_rates_out ro;
  memset(&ro, 0, sizeof(ro));
  ro.rates_row = out_dev; ro.ld = ld_out;
  if ((rc = make_out(&ro, nullptr, gc->n_cells, 1.0, 0, ok))) return rc;
  riab_agents ag; memset(&ag, 0, sizeof(ag));
  riab_motion_params mp; memset(&mp, 0, sizeof(mp));
  riab_step_io io; memset(&io, 0, sizeof(io));
  return launch_tile<GridPolicy<4>, 0>(ek, ag, mp, io, c, ok, pos_dev, n_pos, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------ BVC
static int bvc_n_pad(int n) { return (n + BVC_CT - 1) / BVC_CT * BVC_CT; }

int64_t riab_bvc_pack_floats(int32_t n_cells, int32_t T) {
  const int64_t np = bvc_n_pad(n_cells);
  return 3 * np + np * (int64_t)T + 3 * np + 2 * (int64_t)T + np + 2 * (np / 32);
}
// Cut-off of the angular windows: a von Mises weight (peak 1) below 2^-30 is dropped.  The dropped part of a rate is at
// most T 2^-30 / norm of the population's peak rate (norm = sum of the peak-1 weights >= 1), below float32 resolution of the sum.
static const double BVC_VM_CUT = 9.313225746154785e-10;
int64_t riab_bvc_scratch_floats(int64_t n_pos, int32_t T) {
  return ((n_pos + BVC_AT - 1) / BVC_AT) * (int64_t)T * BVC_AT;
}

int riab_bvc_pack(const double* mu_d, const double* mu_t, const double* sg_d, const double* sg_t, int32_t n,
                  const double* test_angles, int32_t T, riab_bvc_cells* meta, float* out) {
  if (!mu_d || !mu_t || !sg_d || !sg_t || !test_angles || !meta || !out || n <= 0 || T <= 0)
    return fail(RIAB_ERR_INVALID, "riab_bvc_pack: bad argument");
  const int np = bvc_n_pad(n);
  const int64_t total = riab_bvc_pack_floats(n, T);
  for (int64_t i = 0; i < total; ++i) out[i] = 0.f;
  float* s = out; float* m = out + np; float* sc = out + 2 * np; float* vm = out + 3 * (size_t)np;
  // Slots: the cells sorted by preferred angle, so that the 32 cells of a warp of k_bvc_integrate share a narrow angular
  // window outside which every von Mises weight is < BVC_VM_CUT and the terms are skipped.  s | m | sc stay in cell order;
  // the von Mises table is in slot order; perm[slot] = cell (padding slots map to themselves).
  std::vector<int> order(n);
  std::vector<double> key(n);
  const double two_pi = 6.283185307179586;
  for (int i = 0; i < n; ++i) { order[i] = i; double a = fmod(mu_t[i], two_pi); key[i] = a < 0 ? a + two_pi : a; }
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key[a] < key[b]; });
  for (int i = 0; i < n; ++i) {
    const double sv = sqrt(1.4426950408889634 / 2.0) / sg_d[i];       // exp(-(d-mu)^2/(2 sg^2)) = 2^-((d-mu) s)^2
    s[i] = (float)sv; m[i] = (float)(mu_d[i] * sv);
    const double kappa = 1.0 / (sg_t[i] * sg_t[i]);                   // utils.von_mises (utils.py:441-457), norm=1
    double norm = 0.0;
    for (int t = 0; t < T; ++t) norm += exp(kappa * cos(test_angles[t] - 0.0)) * (1.0 / exp(kappa));   // Neurons.py:1599-1604
    sc[i] = (float)(1.0 / norm);
  }
  int32_t* perm = reinterpret_cast<int32_t*>(out + 3 * (size_t)np + (size_t)np * T + 3 * (size_t)np + 2 * (size_t)T);
  int32_t* win = perm + np;
  std::vector<char> keep(T);
  for (int slot = 0; slot < np; ++slot) {
    const int i = slot < n ? order[slot] : slot;
    perm[slot] = i;
    if (slot < n) {
      const double kappa = 1.0 / (sg_t[i] * sg_t[i]);
      const int tile = slot / BVC_CT, cl = slot % BVC_CT;
      for (int t = 0; t < T; ++t)
        vm[((size_t)tile * T + t) * BVC_CT + cl] = (float)(exp(kappa * cos(test_angles[t] - mu_t[i])) * (1.0 / exp(kappa)));
    }
    if (slot % 32 == 31) {                                            // window of this warp's 32 slots: [th0, th0 + len) mod T
      const int s0 = slot - 31, tile = s0 / BVC_CT;
      for (int t = 0; t < T; ++t) {
        keep[t] = 0;
        for (int q = s0; q <= slot; ++q)
          if (!(vm[((size_t)tile * T + t) * BVC_CT + q % BVC_CT] < (float)BVC_VM_CUT)) { keep[t] = 1; break; }   // NaN keeps
      }
      int best_len = 0, best_end = 0;                                 // longest circular run of skippable angles
      for (int t0 = 0; t0 < T; ++t0) {
        if (keep[t0] || !keep[(t0 + T - 1) % T]) continue;            // runs start after a kept angle
        int len = 0;
        while (len < T && !keep[(t0 + len) % T]) ++len;
        if (len > best_len) { best_len = len; best_end = (t0 + len) % T; }
      }
      bool any = false;
      for (int t = 0; t < T; ++t) any = any || keep[t];
      win[2 * (s0 / 32)] = any ? (best_len ? best_end : 0) : 0;
      win[2 * (s0 / 32) + 1] = any ? T - best_len : 0;
    }
  }
  float* ext = vm + (size_t)np * T;                                  // egocentric: kap | cmu | smu | cth | sth
  for (int i = 0; i < n; ++i) {
    ext[i] = (float)(1.4426950408889634 / (sg_t[i] * sg_t[i]));
    ext[np + i] = (float)cos(mu_t[i]);
    ext[2 * np + i] = (float)sin(mu_t[i]);
  }
  for (int t = 0; t < T; ++t) {
    ext[3 * (size_t)np + t] = (float)cos(test_angles[t]);
    ext[3 * (size_t)np + T + t] = (float)sin(test_angles[t]);
  }
  meta->n_pad = np;
  return 0;
}

int riab_bvc_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_bvc_cells* bvc,
                   float* scratch_dev, int32_t* first_wall_dev, const double* head_direction_dev, float* out_dev,
                   int64_t ld_out, void* stream) {
  if (n_pos == 0) return 0;
  EnvK ek;
  OutK ok;
  int rc;
  if ((rc = make_env(env, ek))) return rc;
  if (bvc == nullptr) return fail(RIAB_ERR_INVALID, "bvc NULL");
  if (n_pos > 0 && pos_dev == nullptr) return fail(RIAB_ERR_INVALID, "pos_dev NULL");
  riab_rates_out ro;
  memset(&ro, 0, sizeof(ro));
  ro.rates_row = out_dev; ro.ld = ld_out;
  if ((rc = make_out(&ro, nullptr, bvc->n_cells, 1.0, 0, ok))) return rc;
  riab_agents ag; memset(&ag, 0, sizeof(ag));
  riab_motion_params mp; memset(&mp, 0, sizeof(mp));
  riab_step_io io; memset(&io, 0, sizeof(io));
  return launch_bvc<false>(ek, ag, mp, io, bvc, ok, pos_dev, n_pos, scratch_dev, first_wall_dev, head_direction_dev,
                           (cudaStream_t)stream);
}

// ------------------------------------------------------------ ObjectVectorCells
int64_t riab_ovc_pack_floats(int32_t n_cells) { return (int64_t)place_n_pad(n_cells) * 6; }

int riab_ovc_pack(const double* tuning_distances, const double* tuning_angles, const double* sigma_distances,
                  const double* sigma_angles, const int32_t* tuning_types, int32_t n, riab_ovc_cells* meta, float* out) {
  if (!tuning_distances || !tuning_angles || !sigma_distances || !sigma_angles || !tuning_types || !meta || !out || n <= 0)
    return fail(RIAB_ERR_INVALID, "riab_ovc_pack: bad argument");
  const int np = place_n_pad(n);
  const double log2e = 1.4426950408889634;
  for (int i = 0; i < np; ++i) {
    const bool in = i < n;
    const double kappa = in ? 1.0 / (sigma_angles[i] * sigma_angles[i]) : 0.0;       // utils.von_mises (utils.py:441-457)
    out[i] = in ? (float)tuning_distances[i] : 0.f;
    out[(size_t)np + i] = in ? (float)(sqrt(0.5 * log2e) / sigma_distances[i]) : 0.f;  // utils.gaussian (utils.py:424-438)
    out[(size_t)2 * np + i] = in ? (float)cos(0.5 * tuning_angles[i]) : 1.f;
    out[(size_t)3 * np + i] = in ? (float)sin(0.5 * tuning_angles[i]) : 0.f;
    out[(size_t)4 * np + i] = (float)sqrt(2.0 * kappa * log2e);
    out[(size_t)5 * np + i] = in ? (float)tuning_types[i] : -2.f;                     // padding cells match no object
  }
  meta->n_pad = np;
  return 0;
}

int riab_ovc_rates(const double* pos_dev, int64_t n_pos, const riab_env* env, const riab_ovc_cells* ovc,
                   const double* head_direction_dev, float* out_dev, int64_t ld_out, void* stream) {
  if (n_pos == 0) return 0;
  EnvK ek;
  OvcConst c;
  OutK ok;
  int rc;
  if ((rc = make_env(env, ek)) || (rc = make_ovc(ovc, ek, head_direction_dev, c))) return rc;
  if (n_pos > 0 && pos_dev == nullptr) return fail(RIAB_ERR_INVALID, "pos_dev NULL");
  riab_rates_out ro;
  memset(&ro, 0, sizeof(ro));
  ro.rates_row = out_dev; ro.ld = ld_out;
  if ((rc = make_out(&ro, nullptr, ovc->n_cells, 1.0, 0, ok))) return rc;
  riab_agents ag; memset(&ag, 0, sizeof(ag));
  riab_motion_params mp; memset(&mp, 0, sizeof(mp));
  riab_step_io io; memset(&io, 0, sizeof(io));
  return launch_tile<OvcPolicy, 0>(ek, ag, mp, io, c, ok, pos_dev, n_pos, (cudaStream_t)stream);
}

// ----------------------------------------------------------------- fused step
// MODE 1: motion + rates of one population; 0: rates for agents->pos as it is; 2: skewed (riab_run).
}  // extern "C"
template <int MODE>
static int neurons_update_impl(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                               const riab_step_io* io, int32_t cells_kind, const void* cells,
                               const riab_neuron_noise* noise, const riab_rates_out* out, void* stream) {
  EnvK ek;
  OutK ok;
  int rc;
  if ((rc = check_agents(agents)) || (rc = make_env(env, ek))) return rc;
  if (MODE != 0 && (rc = check_motion(prm))) return rc;
  if (cells == nullptr || (MODE != 0 && io == nullptr)) return fail(RIAB_ERR_INVALID, "io / cells NULL");
  if (MODE == 1 && (io->collision_mask || io->first_hit || io->n_iters) && cells_kind != RIAB_CELLS_BVC) {
    // parity taps are only implemented in the stand-alone motion kernel
    if ((rc = riab_agent_update(agents, env, prm, io, stream))) return rc;
    return neurons_update_impl<0>(agents, env, prm, io, cells_kind, cells, noise, out, stream);
  }
  riab_motion_params mp0; memset(&mp0, 0, sizeof(mp0));
  riab_step_io io0; memset(&io0, 0, sizeof(io0));
  const riab_motion_params& mp = (MODE != 0) ? *prm : mp0;
  const riab_step_io& sio = (MODE != 0) ? *io : io0;
  const double dt = (prm != nullptr) ? prm->dt : (noise ? (double)noise->dt : 1.0);
  const double* pos_in = (MODE == 1) ? nullptr : agents->pos;
  cudaStream_t s = (cudaStream_t)stream;
  if (cells_kind == RIAB_CELLS_PLACE) {
    const riab_place_cells* pc = (const riab_place_cells*)cells;
    PlaceConst c;
    const bool onehot = pc != nullptr && pc->description == RIAB_PC_ONE_HOT;     // post-pass spikes (k_finish_rows): dense
    if ((rc = make_place(pc, ek, c)) ||
        (rc = make_out(out, noise, pc->n_cells, dt, agents->id_offset, ok, onehot ? -1.0 : (double)fmaxf(pc->min_fr, pc->max_fr)))) return rc;
    if (c.desc == RIAB_PC_ONE_HOT) {             // arg-min across cells: its own kernel after the motion kernel
      if (MODE == 2) return fail(RIAB_ERR_UNSUPPORTED, "one_hot populations are stepped unskewed");
      if (MODE == 1 && (rc = riab_agent_update(agents, env, prm, io, stream))) return rc;
      return launch_place<0>(ek, *agents, mp0, io0, c, ok, agents->pos, agents->n_agents, s);
    }
    return launch_place<MODE>(ek, *agents, mp, sio, c, ok, pos_in, agents->n_agents, s);
  }
  if (cells_kind == RIAB_CELLS_GRID) {
    const riab_grid_cells* gc = (const riab_grid_cells*)cells;
    GridConst c;
    if ((rc = make_grid(gc, ek, c)) ||
        (rc = make_out(out, noise, gc->n_cells, dt, agents->id_offset, ok, (double)fmaxf(gc->min_fr, gc->max_fr)))) return rc;
    return launch_tile<GridPolicy<4>, MODE>(ek, *agents, mp, sio, c, ok, pos_in, agents->n_agents, s);
  }
  if (cells_kind == RIAB_CELLS_BVC) {
    if (MODE == 2) return fail(RIAB_ERR_UNSUPPORTED, "BVC populations are stepped unskewed");
    const riab_bvc_cells* bvc = (const riab_bvc_cells*)cells;
    if ((rc = make_out(out, noise, bvc->n_cells, dt, agents->id_offset, ok))) return rc;
    // The ray kernel wants all its CTAs resident in ONE wave (the float64 ray chains are latency-bound);
    // fusing the 128-register motion code into it halves its occupancy and adds a tail wave, so the
    // motion runs as its own (14 us) kernel first.
    if (MODE == 1 && (rc = riab_agent_update(agents, env, prm, io, stream))) return rc;
    return launch_bvc<false>(ek, *agents, mp0, io0, bvc, ok, agents->pos, agents->n_agents, out->bvc_scratch, nullptr,
                             agents->head_direction, s);
  }
  if (cells_kind == RIAB_CELLS_OVC) {
    const riab_ovc_cells* oc = (const riab_ovc_cells*)cells;
    OvcConst c;
    if ((rc = make_ovc(oc, ek, agents->head_direction, c)) || (rc = make_out(out, noise, oc->n_cells, dt, agents->id_offset, ok))) return rc;
    return launch_tile<OvcPolicy, MODE>(ek, *agents, mp, sio, c, ok, pos_in, agents->n_agents, s);
  }
  return fail(RIAB_ERR_INVALID, "bad cells_kind %d", cells_kind);
}
extern "C" {

int riab_step_fused(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                    const riab_step_io* io, int32_t cells_kind, const void* cells, const riab_neuron_noise* noise,
                    const riab_rates_out* out, void* stream) {
  return neurons_update_impl<1>(agents, env, prm, io, cells_kind, cells, noise, out, stream);
}

int riab_neurons_update(const riab_agents* agents, const riab_env* env, int32_t cells_kind, const void* cells,
                        const riab_neuron_noise* noise, const riab_rates_out* out, void* stream) {
  return neurons_update_impl<0>(agents, env, nullptr, nullptr, cells_kind, cells, noise, out, stream);
}

int riab_run(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm, const riab_step_io* io,
             const riab_population* pops, int32_t n_pops, const riab_agent_history* hist, int64_t n_steps,
             void* stream) {
  if (agents == nullptr || io == nullptr || n_pops < 0 || (n_pops > 0 && pops == nullptr))
    return fail(RIAB_ERR_INVALID, "riab_run: bad argument");
  const int64_t A = agents->n_agents;
  // Skewed schedule (population 0 is a Place/Grid population): motion(0) alone, then per step one
  // kernel that evaluates rates(s) of the current positions while its producer warps already run
  // motion(s+1); the last step is rates only.  Same results as the plain sequence, but the
  // float64 motion chain never gates the rate warps.
  const bool onehot0 = n_pops >= 1 && pops[0].kind == RIAB_CELLS_PLACE && pops[0].cells != nullptr &&
                       ((const riab_place_cells*)pops[0].cells)->description == RIAB_PC_ONE_HOT;
  const bool skew = n_pops >= 1 && pops[0].kind != RIAB_CELLS_BVC && !onehot0 && n_steps >= 1 && io->xi == nullptr &&
                    !io->collision_mask && !io->first_hit && !io->n_iters;
  auto step_io = [&](int64_t st) {
    riab_step_io sio = *io;
    sio.step = io->step + (uint64_t)st;
    sio.history_row = nullptr;
    if (hist != nullptr && hist->ring != nullptr && hist->ring_rows > 0)
      sio.history_row = hist->ring + (size_t)((hist->ring_next + st) % hist->ring_rows) * A * 8;
    return sio;
  };
  int rc;
  // ---- a single Place / Grid population: the whole run as ONE launch (k_step MODE 3, see RunK) where the lean consumer
  // loop applies: vector-aligned rows, whole 4-cell groups, all cells in one register set, consumer groups that divide the
  // producers, no OU noise, no parity taps, and rings that hold the run without wrapping onto rows still being written
  if (n_pops == 1 && skew && getenv("RIAB_NO_WHOLE_RUN") == nullptr && io->drift_velocity == nullptr && io->pos_mirror == nullptr) {
    const riab_population& pp = pops[0];
    EnvK ek;
    OutK ok;
    if ((rc = check_agents(agents)) || (rc = make_env(env, ek)) || (rc = check_motion(prm))) return rc;
    const bool place = pp.kind == RIAB_CELLS_PLACE, grid = pp.kind == RIAB_CELLS_GRID;
    if ((place || grid) && pp.rates_ring != nullptr && pp.ring_rows > 0 && pp.noise.noise_std == 0.f) {
      PlaceConst pcst;
      GridConst gcst;
      int n_cells = 0, n_pad = 0;
      double bound = -1.0;
      bool ok_cells = true;
      if (place) {
        const riab_place_cells* pc = (const riab_place_cells*)pp.cells;
        if ((rc = make_place(pc, ek, pcst))) return rc;
        n_cells = pc->n_cells; n_pad = pc->n_pad; bound = (double)fmaxf(pc->min_fr, pc->max_fr);
        ok_cells = pc->description != RIAB_PC_ONE_HOT;
      } else {
        const riab_grid_cells* gc = (const riab_grid_cells*)pp.cells;
        if ((rc = make_grid(gc, ek, gcst))) return rc;
        n_cells = gc->n_cells; n_pad = gc->n_pad; bound = (double)fmaxf(gc->min_fr, gc->max_fr);
      }
      riab_rates_out ro = pp.out;
      ro.rates_row = pp.rates_ring;                                    // (alignment / vector checks of make_out)
      ro.spikes_row = pp.spikes_ring;
      riab_neuron_noise nz = pp.noise;
      nz.dt = (float)prm->dt;
      if ((rc = make_out(&ro, &nz, n_cells, prm->dt, agents->id_offset, ok, bound))) return rc;
      const int ct = n_pad / 4;
      const bool rows_ok = ((A * ok.ld) % 4 == 0) && ((A * ok.spike_ld) % 4 == 0);        // every ring row stays 16-byte aligned
      const bool lean = ok_cells && ok.vec_ok && rows_ok && (n_cells % 4 == 0) && ct <= RW * 32 &&
                        (ok.spikes == nullptr || (agents->id_offset & 1) == 0) &&
                        (4 % lean_groups(ct, 8) == 0) && (4 % lean_groups(ct, 4) == 0);      // groups divide the 4 producers
      const bool hist_ok = hist == nullptr || hist->ring == nullptr || hist->ring_rows > 0;
      if (lean && hist_ok) {
        RunK run;
        memset(&run, 0, sizeof(run));
        run.n_steps = n_steps;
        run.rates_ring = pp.rates_ring; run.spikes_ring = pp.spikes_ring;
        run.ring_rows = pp.ring_rows; run.ring_next = pp.ring_next;
        if (hist != nullptr && hist->ring != nullptr && hist->ring_rows > 0) {
          run.hist_ring = hist->ring; run.hist_rows = hist->ring_rows; run.hist_next = hist->ring_next;
        }
        riab_step_io io0 = *io;
        io0.history_row = nullptr;
        cudaStream_t s = (cudaStream_t)stream;
        if (place) return launch_place<3>(ek, *agents, *prm, io0, pcst, ok, nullptr, A, s, &run);
        return launch_tile<GridPolicy<4>, 3>(ek, *agents, *prm, io0, gcst, ok, nullptr, A, s, &run);
      }
    }
  }
  if (skew) {
    const riab_step_io s0 = step_io(0);
    if ((rc = riab_agent_update(agents, env, prm, &s0, stream))) return rc;
  }
  // ---- BoundaryVectorCells: pipeline rays(s+1) against the integral of step s (see BvcPipe)
  static thread_local BvcPipe pipes[16];
  BvcPipe* pipe = nullptr;
  {
    int dev = 0;
    // only when every population is a BoundaryVectorCells one: next to Place / Grid rate kernels (HBM- and dispatch-bound) the
    // overlapped integral just competes for the same SMs (measured: configs[4] 2.22 ms without, 2.37 ms with the pipeline)
    bool want = n_pops >= 1;
    for (int p = 0; p < n_pops; ++p)
      want = want && p < PIPE_POPS && (pops[p].kind == RIAB_CELLS_BVC && pops[p].cells != nullptr && !((const riab_bvc_cells*)pops[p].cells)->egocentric &&
                                       pops[p].ring_rows >= 2 && pops[p].out.bvc_scratch != nullptr);
    if (want && n_steps >= 2 && getenv("RIAB_NO_BVC_PIPELINE") == nullptr && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 16) {
      pipe = &pipes[dev];
      if (pipe->side == nullptr) {
        // keep the stream-ordered pool's memory across runs (by default it goes back to the driver at every synchronisation
        // and each run would pay a ~1 ms cudaMalloc for its second ray buffer again)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
          unsigned long long keep = ~0ull;
          cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        RIAB_CUDA_OK(cudaStreamCreateWithFlags(&pipe->side, cudaStreamNonBlocking));
        RIAB_CUDA_OK(cudaEventCreateWithFlags(&pipe->rays_done, cudaEventDisableTiming));
        for (int b = 0; b < 2; ++b)
          for (int p = 0; p < PIPE_POPS; ++p) RIAB_CUDA_OK(cudaEventCreateWithFlags(&pipe->int_done[b][p], cudaEventDisableTiming));
      }
      for (int b = 0; b < 2; ++b)
        for (int p = 0; p < PIPE_POPS; ++p) pipe->used[b][p] = false;
      for (int p = 0; p < PIPE_POPS; ++p) pipe->scratch2[p] = nullptr;
      for (int p = 0; p < n_pops && p < PIPE_POPS; ++p) {
        if (pops[p].kind != RIAB_CELLS_BVC || pops[p].cells == nullptr || pops[p].ring_rows < 2 || pops[p].out.bvc_scratch == nullptr) continue;
        const riab_bvc_cells* bc = (const riab_bvc_cells*)pops[p].cells;
        if (bc->egocentric) continue;
        const int pid = pops[p].noise.population_id;
        if (pid < 0 || pid >= PIPE_POPS) continue;
        RIAB_CUDA_OK(cudaMallocAsync((void**)&pipe->scratch2[pid], (size_t)riab_bvc_scratch_floats(A, bc->n_test_angles) * sizeof(float),
                                     (cudaStream_t)stream));
      }
    }
  }
  struct PipeGuard {                                      // leaves riab_run: join the side stream, free the second buffers
    BvcPipe* p; cudaStream_t s;
    ~PipeGuard() {
      g_pipe = nullptr;
      if (p == nullptr) return;
      for (int b = 0; b < 2; ++b)
        for (int q = 0; q < PIPE_POPS; ++q)
          if (p->used[b][q]) cudaStreamWaitEvent(s, p->int_done[b][q], 0);
      for (int q = 0; q < PIPE_POPS; ++q)
        if (p->scratch2[q] != nullptr) { cudaFreeAsync(p->scratch2[q], s); p->scratch2[q] = nullptr; }
    }
  } guard{pipe, (cudaStream_t)stream};
  g_pipe = pipe;
  for (int64_t st = 0; st < n_steps; ++st) {
    if (pipe) pipe->step = st;
    const riab_step_io sio = step_io(st);
    if (n_pops == 0) {
      if ((rc = riab_agent_update(agents, env, prm, &sio, stream))) return rc;
      continue;
    }
    // populations 1.. first (they read the positions of step st), population 0 last (it may advance them)
    for (int pi = 0; pi < n_pops; ++pi) {
      const int p = skew ? ((pi + 1) % n_pops) : pi;
      const riab_population& pp = pops[p];
      if (pp.rates_ring == nullptr || pp.ring_rows <= 0) return fail(RIAB_ERR_INVALID, "population %d: no rates ring", p);
      const size_t slot = (size_t)((pp.ring_next + st) % pp.ring_rows);
      riab_rates_out ro = pp.out;
      ro.rates_row = pp.rates_ring + slot * A * pp.out.ld;
      int n_cells = 0;
      if (pp.kind == RIAB_CELLS_PLACE) n_cells = ((const riab_place_cells*)pp.cells)->n_cells;
      else if (pp.kind == RIAB_CELLS_GRID) n_cells = ((const riab_grid_cells*)pp.cells)->n_cells;
      else if (pp.kind == RIAB_CELLS_BVC) n_cells = ((const riab_bvc_cells*)pp.cells)->n_cells;
      else if (pp.kind == RIAB_CELLS_OVC) n_cells = ((const riab_ovc_cells*)pp.cells)->n_cells;
      ro.spikes_row = pp.spikes_ring ? pp.spikes_ring + slot * A * (size_t)(4 * ((n_cells + 127) / 128)) : nullptr;
      riab_neuron_noise nz = pp.noise;
      nz.step = pp.noise.step + (uint64_t)st;
      nz.dt = (float)prm->dt;
      if (p == 0 && !skew) rc = riab_step_fused(agents, env, prm, &sio, pp.kind, pp.cells, &nz, &ro, stream);
      else if (p == 0 && st + 1 < n_steps) {
        const riab_step_io nxt = step_io(st + 1);          // the motion it runs belongs to step st+1
        rc = neurons_update_impl<2>(agents, env, prm, &nxt, pp.kind, pp.cells, &nz, &ro, stream);
      } else rc = riab_neurons_update(agents, env, pp.kind, pp.cells, &nz, &ro, stream);
      if (rc) return rc;
    }
  }
  return 0;
}

int riab_step_fused_host(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                         riab_step_io* io, int32_t cells_kind, const void* cells, const riab_neuron_noise* noise,
                         const riab_rates_out* out, const double* drift_host, double* drift_staging_dev,
                         double* pos_out_host, void* stream) {
  if (agents == nullptr || io == nullptr) return fail(RIAB_ERR_INVALID, "agents / io NULL");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)agents->n_agents * 2 * sizeof(double);
  if (drift_host != nullptr) {
    if (drift_staging_dev == nullptr) return fail(RIAB_ERR_INVALID, "drift_staging_dev NULL");
    RIAB_CUDA_OK(cudaMemcpyAsync(drift_staging_dev, drift_host, bytes, cudaMemcpyHostToDevice, s));
    io->drift_velocity = drift_staging_dev;
  }
  const int rc = riab_step_fused(agents, env, prm, io, cells_kind, cells, noise, out, stream);
  if (rc) return rc;
  if (pos_out_host != nullptr) RIAB_CUDA_OK(cudaMemcpyAsync(pos_out_host, agents->pos, bytes, cudaMemcpyDeviceToHost, s));
  return 0;
}


// ---- host-buffer motion step on two streams (see include/riab_b200.h)
namespace {
struct HostIo {
  cudaStream_t side = nullptr;
  cudaEvent_t up_done = nullptr, motion_done = nullptr, pos_done = nullptr;
  bool pos_inflight = false;
};
thread_local HostIo g_hostio[16];
int hostio(HostIo*& h) {
  int dev = 0;
  RIAB_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16) return fail(RIAB_ERR_UNSUPPORTED, "device ordinal %d", dev);
  h = &g_hostio[dev];
  if (h->side == nullptr) {
    RIAB_CUDA_OK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    RIAB_CUDA_OK(cudaEventCreateWithFlags(&h->up_done, cudaEventDisableTiming));
    RIAB_CUDA_OK(cudaEventCreateWithFlags(&h->motion_done, cudaEventDisableTiming));
    RIAB_CUDA_OK(cudaEventCreateWithFlags(&h->pos_done, cudaEventDisableTiming));
  }
  return 0;
}
}  // namespace

int riab_positions_fence(void* stream) {
  HostIo* h = nullptr;
  int rc;
  if ((rc = hostio(h))) return rc;
  if (h->pos_inflight) RIAB_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)stream, h->pos_done, 0));
  return 0;
}

int riab_positions_wait(void) {
  HostIo* h = nullptr;
  int rc;
  if ((rc = hostio(h))) return rc;
  if (h->pos_inflight) RIAB_CUDA_OK(cudaEventSynchronize(h->pos_done));
  return 0;
}

int riab_agent_update_host(const riab_agents* agents, const riab_env* env, const riab_motion_params* prm,
                           riab_step_io* io, const double* drift_host, double* drift_staging_dev,
                           double* pos_out_host, void* stream) {
  if (agents == nullptr || io == nullptr) return fail(RIAB_ERR_INVALID, "agents / io NULL");
  HostIo* h = nullptr;
  int rc;
  if ((rc = hostio(h))) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t bytes = (size_t)agents->n_agents * 2 * sizeof(double);
  if (drift_host != nullptr) {
    if (drift_staging_dev == nullptr) return fail(RIAB_ERR_INVALID, "drift_staging_dev NULL");
    // the previous motion kernel (the last reader of the staging buffer) finished before the side stream's last copy
    // started (that copy waited for motion_done), so the upload may start at once
    RIAB_CUDA_OK(cudaMemcpyAsync(drift_staging_dev, drift_host, bytes, cudaMemcpyHostToDevice, h->side));
    RIAB_CUDA_OK(cudaEventRecord(h->up_done, h->side));
    RIAB_CUDA_OK(cudaStreamWaitEvent(s, h->up_done, 0));
    io->drift_velocity = drift_staging_dev;
  }
  if (h->pos_inflight) RIAB_CUDA_OK(cudaStreamWaitEvent(s, h->pos_done, 0));   // the copy that still reads agents->pos
  if ((rc = riab_agent_update(agents, env, prm, io, stream))) return rc;
  RIAB_CUDA_OK(cudaEventRecord(h->motion_done, s));
  RIAB_CUDA_OK(cudaStreamWaitEvent(h->side, h->motion_done, 0));
  if (pos_out_host != nullptr) {
    RIAB_CUDA_OK(cudaMemcpyAsync(pos_out_host, agents->pos, bytes, cudaMemcpyDeviceToHost, h->side));
    RIAB_CUDA_OK(cudaEventRecord(h->pos_done, h->side));
    h->pos_inflight = true;
  }
  return 0;
}

}  // extern "C"
