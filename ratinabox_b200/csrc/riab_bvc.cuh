// BoundaryVectorCells.get_state, allocentric (ratinabox/Neurons.py:1617-1778).
//
// Phase A (float64, per agent x test angle): the reference casts T rays
// pos -> pos + test_direction, intersects each with every wall
// (utils.vector_intercepts, utils.py:30-118), scores the walls with
// boundary_vector_preference_function (Neurons.py:1746-1778) and keeps
// l_a of the arg-max wall = distance to the first wall along the ray.  We
// evaluate exactly those float64 expressions (D = non-contracting double) so
// the chosen wall and distance equal the oracle's; the only short-cut is that
// the `l_b < 0` / `l_b > 1` rejections are decided from the signs / magnitudes of
// numerator and denominator (exact for IEEE division) so l_a's division is only
// done for walls that survive.
//
// Phase B (float32, per agent x cell): fr = sum_theta gauss(d_theta; mu_d, sigma_d) *
// vonmises(theta; mu_theta, sigma_theta) / cell_fr_norm.  The von Mises factor does
// not depend on the agent, so it is a precomputed (cell tile x T) table staged into
// shared memory by TMA bulk copies; the Gaussian is one FFMA + FMUL + MUFU.EX2.
//
// Packed block (float32, riab_bvc_pack), Np = n_cells rounded up to BVC_CT (64):
//   s[Np] | m[Np] | scale[Np] | VM tiles: [Np/64][T][64] | kap[Np] | cmu[Np] | smu[Np] | cth[T] | sth[T]
//   s = sqrt(log2(e)/2)/sigma_d, m = mu_d*s, scale = 1/cell_fr_norm,
//   kap = log2(e)/sigma_theta^2, (cmu,smu) = (cos,sin)(mu_theta), (cth,sth) = (cos,sin)(test angle).
// Egocentric cells (FieldOfViewBVCs): the von Mises argument is theta - head_bearing - mu_theta, which
// depends on the agent, so no table: cos(theta - phi) = cth*cos(phi) + sth*sin(phi) with
// (cos,sin)(phi = bearing + mu_theta) formed algebraically per (agent, cell) from the head direction
// (cos(bearing), sin(bearing)) = (x+1e-6, y)/|(x+1e-6, y)|  -- utils.get_angle's eps quirk included --
// and both exponentials share ONE ex2:  2^(kap*(cos-1) - u^2).
// Scratch (phase A -> B): dist_to_first_wall as [agent tile of 32][T][32] float32.
#pragma once
#include "riab_common.cuh"

namespace riab {

constexpr int BVC_CT = 64;   // cells per tile
constexpr int BVC_AT = 32;   // agents per tile

// One ray against all walls -> (distance to first wall, wall id).  Neurons.py:1655-1684.
// wf: float32 copy of the walls as (ax, ay, sbx, sby) per wall (pre-filter only).
//
// Two phases per ray so that the lanes of a warp (different agents / angles) do not serialise on each
// other's walls: (1) a float32 pre-filter over all walls builds a bit mask of the walls whose l_b is
// within [0,1] up to a margin 1e-3 (>> float32 rounding; anything it drops is rejected by the exact
// test too, pref = -1); (2) each lane walks ITS OWN set bits in increasing wall order -- same code path
// for every lane, different wall index as data -- and evaluates the reference's float64 expressions.
RIAB_DEV void bvc_first_wall(double px, double py, double ux, double uy, const double* __restrict__ walls,
                             const float4* __restrict__ wf, int W, double& dist, int& wall_id) {
  const D a0x(px), a0y(py);
  const D a1x = a0x + D(ux), a1y = a0y + D(uy);      // pos_line_segments[:, :, 1, :] += test_directions
  const D sax = a1x - a0x, say = a1y - a0y;
  const D sapx = -say, sapy = sax;
  const float pxf = (float)px, pyf = (float)py, sapxf = -(float)uy, sapyf = (float)ux;
  unsigned long long mask = 0ull;
  for (int w = 0; w < W; ++w) {
    const float4 wl = wf[w];
    const float d0xf = wl.x - pxf, d0yf = wl.y - pyf;
    const float t1 = d0xf * sapxf, t2 = d0yf * sapyf, t3 = wl.z * sapxf, t4 = wl.w * sapyf;
    const float nb = -(t1 + t2), db = t3 + t4;
    const float tol = fmaf(1e-3f, fabsf(db), 1e-6f * (fabsf(t1) + fabsf(t2) + fabsf(t3) + fabsf(t4)));
    const float sgn = (db >= 0.f) ? 1.f : -1.f;
    const float nbs = nb * sgn, dbs = db * sgn;          // dbs >= 0 ;  l_b = nbs / dbs
    const bool rejected = (nbs < -tol) || (nbs > dbs + tol);
    mask |= rejected ? 0ull : (1ull << w);
  }
  double best = -1.0;                                    // rejected walls all score -1; np.argmax keeps the first max
  int besti = -1;
  double best_la = 0.0;
  while (mask) {
    const int w = __ffsll((long long)mask) - 1;
    mask &= mask - 1;
    const D bx0(walls[4 * w]), by0(walls[4 * w + 1]), bx1(walls[4 * w + 2]), by1(walls[4 * w + 3]);
    const D d0x = bx0 - a0x, d0y = by0 - a0y;
    const D sbx = bx1 - bx0, sby = by1 - by0;
    const D sbpx = -sby, sbpy = sbx;
    const D numB = (-d0x) * sapx + (-d0y) * sapy;
    const D denB = sbx * sapx + sby * sapy;
    // l_b = numB/denB : decide (l_b < 0) || (l_b > 1) without dividing (exact for IEEE division)
    bool rej;
    if (denB.v != 0.0 && numB.v == numB.v && fabs(denB.v) != INFINITY && fabs(numB.v) != INFINITY) {
      const bool same = (numB.v > 0.0) == (denB.v > 0.0);
      rej = (numB.v != 0.0) && (!same || fabs(numB.v) > fabs(denB.v));
    } else {
      const double lb = (numB / denB).v;
      rej = (lb < 0.0) || (lb > 1.0);
    }
    if (rej) continue;                                   // pref = -1: never beats the running maximum
    const D numA = d0x * sbpx + d0y * sbpy;
    const D denA = sax * sbpx + say * sbpy;
    const double la = (numA / denA).v;
    const double pref = (la > 0.0) ? __ddiv_rn(1.0, la) : ((la < 0.0) ? -1.0 : 0.0);
    if (pref > best) { best = pref; besti = w; best_la = la; }
  }
  if (besti < 0) {
    // every wall scored -1: np.argmax returns wall 0, whose l_a is reported (Neurons.py:1677-1684)
    besti = 0;
    const D bx0(walls[0]), by0(walls[1]), bx1(walls[2]), by1(walls[3]);
    const D d0x = bx0 - a0x, d0y = by0 - a0y;
    const D sbx = bx1 - bx0, sby = by1 - by0;
    const D sbpx = -sby, sbpy = sbx;
    best_la = ((d0x * sbpx + d0y * sbpy) / (sax * sbpx + say * sbpy)).v;
  }
  dist = best_la;
  wall_id = besti;
}

struct BvcConst {
  int n_cells, n_pad, T, ego;
  const double* head_dir;    // device (n_rows,2) or NULL (egocentric only)
  float min_fr, span;
  const float* packed;
  const double* test_dirs;   // device (T,2)
};

}  // namespace riab
