// ObjectVectorCells.get_state (ratinabox/Neurons.py:1989-2113), float32 rates from float64 agent-object geometry.
//
//   fr_i = sum over objects o with type_o == tuning_type_i of
//            gaussian(d_o; mu_d_i, sigma_d_i, norm=1) * von_mises(bearing_o; mu_theta_i, sigma_theta_i, norm=1)
//   d_o      = |pos - object_o|, or 1000 when walls_occlude and an inner wall crosses the segment
//              (Environment.get_distances_between___accounting_for_environment, `line_of_sight`, Environment.py:710-730)
//   bearing  = utils.get_angle(object_o - pos) [- utils.get_angle(head_direction) when egocentric]
//
// The producer warp evaluates the per-(agent, object) geometry exactly like the reference, in float64
// (few objects: the exact utils.vector_intercepts test is affordable here), and publishes per object
//   (d, cos(b/2), sin(b/2), type)
// in the agent record.  Consumers hold per cell (mu_d, s_d, cos(mu/2), sin(mu/2), k_q, type) and evaluate
//   u = (d - mu_d) s_d,  h = sin((b - mu)/2) = sin(b/2) cos(mu/2) - cos(b/2) sin(mu/2),  g = h k_q
//   term = 2^-(u^2 + g^2)         with s_d^2 = log2(e)/(2 sigma_d^2),  k_q^2 = 2 kappa log2(e)
// (von Mises with norm=1 is exp(kappa (cos(b-mu) - 1)) = exp(-2 kappa sin^2((b-mu)/2)): the half-angle form has no
// cancellation for the narrow field-of-view cells, kappa ~ 10^2..10^3).
#pragma once
#include "riab_common.cuh"
#include "riab_motion.cuh"
#include "riab_place.cuh"

namespace riab {

constexpr int OVC_MAX_OBJ = 9;                    // objects per environment held in one agent record
constexpr int OVC_REC = 4 + 4 * OVC_MAX_OBJ;      // 40 floats = 160 B, like the PlaceCells record

struct OvcConst {                                 // uniform per launch
  int n_cells, n_pad, n_obj, ego, occlude, wall0, n_inner;
  float min_fr, span;
  const float* packed;                            // mu_d | s_d | cos(mu/2) | sin(mu/2) | k_q | type   (Np each)
  const double* head_dir;                         // positions-only launches: (n_pos,2) head directions or NULL = [1,0]
  double obj[2 * OVC_MAX_OBJ];
  float type[OVC_MAX_OBJ];
};

struct OvcCellRegs {
  float mu[4], sd[4], cm[4], sm[4], kq[4], ty[4];
};

RIAB_DEV void ovc_load_cells(OvcCellRegs& r, const OvcConst& c, int cell0) {
  const int np = c.n_pad;
  const float* b = c.packed + cell0;
  const float4 a0 = *reinterpret_cast<const float4*>(b), a1 = *reinterpret_cast<const float4*>(b + np),
               a2 = *reinterpret_cast<const float4*>(b + 2 * np), a3 = *reinterpret_cast<const float4*>(b + 3 * np),
               a4 = *reinterpret_cast<const float4*>(b + 4 * np), a5 = *reinterpret_cast<const float4*>(b + 5 * np);
  r.mu[0] = a0.x; r.mu[1] = a0.y; r.mu[2] = a0.z; r.mu[3] = a0.w;
  r.sd[0] = a1.x; r.sd[1] = a1.y; r.sd[2] = a1.z; r.sd[3] = a1.w;
  r.cm[0] = a2.x; r.cm[1] = a2.y; r.cm[2] = a2.z; r.cm[3] = a2.w;
  r.sm[0] = a3.x; r.sm[1] = a3.y; r.sm[2] = a3.z; r.sm[3] = a3.w;
  r.kq[0] = a4.x; r.kq[1] = a4.y; r.kq[2] = a4.z; r.kq[3] = a4.w;
  r.ty[0] = a5.x; r.ty[1] = a5.y; r.ty[2] = a5.z; r.ty[3] = a5.w;
}

// Per-agent record from the float64 position / head direction.  walls = all walls (W*4 doubles).
RIAB_DEV void ovc_agent_record(float* __restrict__ rec, double px, double py, double hdx, double hdy,
                               const double* __restrict__ walls, const OvcConst& c) {
  const double hb = c.ego ? get_angle(hdx, hdy) : 0.0;                       // Neurons.py:2046-2047
  *reinterpret_cast<float4*>(rec) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int o = 0; o < OVC_MAX_OBJ; ++o) {
    float4 q = make_float4(1000.f, 1.f, 0.f, -1.f);                           // no such object: type -1 matches no cell
    if (o < c.n_obj) {
      const double ox = c.obj[2 * o], oy = c.obj[2 * o + 1];
      const D vx = D(px) - D(ox), vy = D(py) - D(oy);                         // pos1 - pos2 (utils.py:213)
      double d = dsqrt(vx * vx + vy * vy).v;                                  // np.linalg.norm
      if (c.occlude) {
        const double* inner = walls + 4 * c.wall0;                            // walls[4:] (Environment.py:715-717)
        bool blocked = false;
        for (int j = 0; j < c.n_inner; ++j) blocked = blocked || los_blocked_exact(px, py, ox, oy, inner + 4 * j);
        if (blocked) d = 1000.0;                                              // Environment.py:730
      }
      const double b = get_angle(-vx.v, -vy.v) - hb;                          // bearing of object - pos (Neurons.py:2026-2047)
      double sh, ch;
      sincos(0.5 * b, &sh, &ch);
      q = make_float4((float)d, (float)ch, (float)sh, c.type[o]);
    }
    *reinterpret_cast<float4*>(rec + 4 + 4 * o) = q;
  }
}

RIAB_DEV void ovc_rates4(float (&out)[4], const OvcCellRegs& r, const OvcConst& c, const float* __restrict__ rec) {
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int o = 0; o < c.n_obj; ++o) {                                         // warp-uniform trip count
    const float4 q = *reinterpret_cast<const float4*>(rec + 4 + 4 * o);       // d, cos(b/2), sin(b/2), type
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float u = (q.x - r.mu[i]) * r.sd[i];
      const float g = fmaf(q.z, r.cm[i], -q.y * r.sm[i]) * r.kq[i];
      const float t = ex2f(fmaf(-g, g, -u * u));
      acc[i] += (q.w == r.ty[i]) ? t : 0.f;                                   // tuning_mask (Neurons.py:2098-2101)
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) out[i] = fmaf(acc[i], c.span, c.min_fr);        // Neurons.py:2106-2108
}

}  // namespace riab
