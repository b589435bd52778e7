// GridCells.get_state, 2D (ratinabox/Neurons.py:1172-1236): rectified / shifted sum
// of three cosines, float32.
//
// Packed block (float32, riab_grid_pack), Np = n_cells rounded up to 4:
//   for k = 0..2:  kx_k[Np] | ky_k[Np] | ph_k[Np]
// with (kx,ky) = (2 pi / gridscale) * w_k  and  ph_k = (2 pi / gridscale) * ((origin - box centre) . w_k)
// reduced to [-pi, pi] in float64, so that   phi_k = ph_k - (p' . k_k),  p' = pos - box centre
// (Neurons.py:1191-1201: vecs = origin - pos, phi = (2 pi / gridscale) (vecs . w)).
#pragma once
#include "riab_common.cuh"

namespace riab {

struct GridConst {
  int n_cells, n_pad, rectify;
  float A, B;          // f = A * (cos1+cos2+cos3) + B   (then max(0,.) when rectify)
  float min_fr, span;
  // scale folded into the affine map (make_grid): rate = clamp(As * sum + Bs) with As = A span, Bs = B span + min_fr;
  // max(f, 0) span + min_fr = max(As sum + Bs, min_fr) for span >= 0 (min(.) for span < 0)   (Neurons.py:1214,1232-1234)
  float As, Bs;
  int clamp;           // 0: none (shifted cosines), 1: max(., min_fr), 2: min(., min_fr)
  const float* packed;
  double cxm, cym;
};

template <int CPT = 4>
struct GridCellRegs {
  float kx[3][CPT], ky[3][CPT], ph[3][CPT];
};

template <int CPT>
RIAB_DEV void grid_load_cells(GridCellRegs<CPT>& r, const GridConst& c, int cell0) {
  const int np = c.n_pad;
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    ldv<CPT>(r.kx[k], c.packed + (3 * k + 0) * np + cell0);
    ldv<CPT>(r.ky[k], c.packed + (3 * k + 1) * np + cell0);
    ldv<CPT>(r.ph[k], c.packed + (3 * k + 2) * np + cell0);
  }
}

template <int CPT>
RIAB_DEV void grid_rates4(float (&out)[CPT], const GridCellRegs<CPT>& r, const GridConst& c, const float* __restrict__ rec) {
  const float2 p = *reinterpret_cast<const float2*>(rec);
  const f32x2 npx = bc2(-p.x), npy = bc2(-p.y);
#pragma unroll
  for (int h = 0; h < CPT / 2; ++h) {                 // cell pairs: the three phases are 2 FFMA2 each per two rates
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      f32x2 phi = ffma2(pk2(r.kx[k][2 * h], r.kx[k][2 * h + 1]), npx, pk2(r.ph[k][2 * h], r.ph[k][2 * h + 1]));
      phi = ffma2(pk2(r.ky[k][2 * h], r.ky[k][2 * h + 1]), npy, phi);
      float a, b;
      upk2(phi, a, b);
      if (k == 0) { s0 = __cosf(a); s1 = __cosf(b); }
      else { s0 += __cosf(a); s1 += __cosf(b); }
    }
    float v0 = fmaf(s0, c.As, c.Bs), v1 = fmaf(s1, c.As, c.Bs);
    if (c.clamp == 1) { v0 = fmaxf(v0, c.min_fr); v1 = fmaxf(v1, c.min_fr); }
    else if (c.clamp == 2) { v0 = fminf(v0, c.min_fr); v1 = fminf(v1, c.min_fr); }
    out[2 * h] = v0; out[2 * h + 1] = v1;
  }
}

}  // namespace riab
