// rf_geometry.cuh — geometry stage: vertex shader + outcodes, primitive assembly, view-frustum
// clipping, perspective divide / viewport, face culling, triangle setup and the edge walk that
// turns each surviving triangle into per-scanline span records.
//
// Reference path: render.rs:158-196 -> clip.rs:229-400 -> prim.rs:62-88 -> raster.rs:185-302,80-114.
#pragma once
#include "rf_device.cuh"

// =============================================================================================
// K1: vertex shader (catalogue) + ClipVert::new outcode.  One thread per vertex of the pass.
// render.rs:158-165; shader.rs:31-41; clip.rs:303-309
// =============================================================================================
// K0 k_objects: scene-style per-object culling (render/scene.rs:59-87; the loop of demos/src/bin/crates.rs:100-122).
// One thread per draw that carries a bounding box: the 8 corners go through the draw's model-to-projection matrix
// (ProjMat3::apply, mat.rs:968-972), get their outcodes (ClipVert::new, clip.rs:180-190) and the box is Hidden when all
// corners are outside one plane (view_frustum::status, clip.rs:245-267). A hidden draw is skipped by k_vertex/k_assemble.
__global__ void __launch_bounds__(128) k_objects(PassParams P) {
  if (rf_poisoned(P)) return;
  for (uint32_t d = blockIdx.x * blockDim.x + threadIdx.x; d < P.n_draws; d += gridDim.x * blockDim.x) {
    const DrawDesc& D = P.draws[d];
    if (!(D.flags & RF_F_BBOX)) continue;
    uint32_t all = 0x3Fu;
#pragma unroll
    for (int k = 0; k < 8; k++) {  // BBox::verts order (scene.rs:62-69); the order does not matter for the AND
      const float x = D.bbox[(k & 4) ? 3 : 0], y = D.bbox[(k & 2) ? 4 : 1], z = D.bbox[(k & 1) ? 5 : 2];
      float pos[4];
#pragma unroll
      for (int r = 0; r < 4; r++) pos[r] = dot4p(D.vs_u + 4 * r, x, y, z, 1.0f);
      all &= outcode(pos[0], pos[1], pos[2], pos[3]);
    }
    if (all != 0) P.dstats[d].hidden = 1ull;
  }
}

// =============================================================================================
// K2: per primitive — assembly, clip, to_screen, cull, setup, edge walk -> span records.
// One thread per input primitive; span/half storage is claimed with one warp-aggregated atomic
// (warp prefix sum over the lanes' row counts), so the output buffer is compact.
// Submission order is carried by the key (global prim index * 8 + fan index), not by position.
// =============================================================================================
template <int LT> struct CVert {
  float p[4];
  float a[LT];
  uint32_t oc;
};

template <int LT>
__device__ __forceinline__ void load_cv(const float* __restrict__ cv, uint32_t gv, CVert<LT>& v) {
  constexpr int CVS = Rec<LT>::CVS;
  const float4* q = reinterpret_cast<const float4*>(cv + (size_t)gv * CVS);
  float buf[CVS];
#pragma unroll
  for (int i = 0; i < CVS / 4; i++) {
    float4 t = __ldg(q + i);
    buf[4 * i] = t.x; buf[4 * i + 1] = t.y; buf[4 * i + 2] = t.z; buf[4 * i + 3] = t.w;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) v.p[i] = buf[i];
  v.oc = __float_as_uint(buf[4]);
#pragma unroll
  for (int i = 0; i < LT; i++) v.a[i] = buf[5 + i];
}

// clip.rs:121-145 ClipPlane::intersect
template <int LT>
__device__ __forceinline__ bool clip_intersect(int plane, const CVert<LT>& v0, const CVert<LT>& v1, CVert<LT>& out) {
  const float d0 = plane_dist(plane, v0.p[0], v0.p[1], v0.p[2], v0.p[3]);
  const float d1 = plane_dist(plane, v1.p[0], v1.p[1], v1.p[2], v1.p[3]);
  if (!(d0 * d1 < 0.0f)) return false;
  const float t = -d0 / (d1 - d0);
#pragma unroll
  for (int i = 0; i < 4; i++) out.p[i] = lerpf(v0.p[i], v1.p[i], t);
#pragma unroll
  for (int i = 0; i < LT; i++) out.a[i] = lerpf(v0.a[i], v1.a[i], t);
  out.oc = outcode(out.p[0], out.p[1], out.p[2], out.p[3]);
  return true;
}

// clip.rs:283-301 + 167-193: Sutherland–Hodgman over the six planes in PLANES order.
// poly holds the result (n <= 9 vertices). Runs only for the <= 2 % of triangles with Status::Clipped.
template <int LT>
__device__ __noinline__ int clip_polygon(CVert<LT>* poly /*[10], in: 3 verts*/, CVert<LT>* tmp /*[10]*/) {
  int n = 3;
  CVert<LT>* in = poly;
  CVert<LT>* out = tmp;
  for (int p = 0; p < 6; p++) {
    int m = 0;
    const uint32_t bit = 1u << p;
    for (int k = 0; k < n; k++) {
      const CVert<LT>& v0 = in[k];
      const CVert<LT>& v1 = in[(k + 1 == n) ? 0 : k + 1];
      if ((v0.oc & bit) == 0 && m < 10) out[m++] = v0;
      CVert<LT> x;
      if (clip_intersect<LT>(p, v0, v1, x) && m < 10) out[m++] = x;
    }
    n = m;
    CVert<LT>* t = in; in = out; out = t;
    if (n == 0) break;
  }
  if (in != poly)
    for (int k = 0; k < n; k++) poly[k] = in[k];
  return n;
}

// Screen-space vertex lanes used by setup: x, y, z, attr
template <int LT> struct SVert {
  float x, y, z;
  float a[LT];
};

// prim.rs:62-88
template <int LT>
__device__ __forceinline__ void to_screen(const CVert<LT>& c, const float* __restrict__ vp, uint32_t persp_mask, SVert<LT>& s) {
  const float w = c.p[3];
  const float px = c.p[0] / w, py = c.p[1] / w, pz = 1.0f / w;
  s.x = dot4p(vp + 0, px, py, pz, 1.0f);
  s.y = dot4p(vp + 4, px, py, pz, 1.0f);
  s.z = dot4p(vp + 8, px, py, pz, 1.0f);
#pragma unroll
  for (int i = 0; i < LT; i++) s.a[i] = ((persp_mask >> i) & 1u) ? zdiv(c.a[i], w) : c.a[i];
}

// K1 k_vertex: catalogue vertex shader, outcode, and the screen-space form of inside vertices.
template <int LT>
__global__ void __launch_bounds__(256) k_vertex(PassParams P) {
  constexpr int CVS = Rec<LT>::CVS;
  if (rf_poisoned(P)) return;
  for (uint32_t gv = blockIdx.x * blockDim.x + threadIdx.x; gv < P.NV; gv += gridDim.x * blockDim.x) {
    const uint32_t d = find_draw(P.vbase, P.n_draws, gv, P.verts_per_draw);
    if (P.any_bbox && P.dstats[d].hidden) continue;  // object culled: its clip vertices are never read
    const DrawDesc& D = P.draws[d];
    const float* __restrict__ in = D.verts + (size_t)(gv - __ldg(P.vbase + d)) * D.vstride;
    const uint32_t L = D.L;
    const float x = __ldg(in), y = __ldg(in + 1), z = __ldg(in + 2);
    float a[LT];
#pragma unroll
    for (int i = 0; i < LT; i++) a[i] = (i < (int)L) ? __ldg(in + 3 + i) : 0.0f;

    float pos[4], out[LT];
#pragma unroll
    for (int i = 0; i < LT; i++) out[i] = 0.0f;
    const float* u = D.vs_u;
    switch (D.vs) {
      case RF_VS_MVP:  // mat.rs:968-972 (ProjMat3::apply on a point: [p,1])
#pragma unroll
        for (int r = 0; r < 4; r++) pos[r] = dot4p(u + 4 * r, x, y, z, 1.0f);
#pragma unroll
        for (int i = 0; i < LT; i++) out[i] = a[i];
        break;
      case RF_VS_MVP_LINEARIZE:  // hello_tri.rs:13-17; color.rs:277-285
#pragma unroll
        for (int r = 0; r < 4; r++) pos[r] = dot4p(u + 4 * r, x, y, z, 1.0f);
#pragma unroll
        for (int i = 0; i < LT; i++) out[i] = (i < (int)L) ? powf(a[i], 2.2f) : 0.0f;
        break;
      case RF_VS_SOLIDS: {  // solids.rs:70-79
        if (LT >= 3) {
          const float nz = dot4p(u + 16 + 8, a[0], a[1], a[2], 0.0f);  // spin.apply(normal): w = 0
          const float diffuse = fmaxf(nz + 0.2f, 0.2f) * 0.8f;
#pragma unroll
          for (int i = 0; i < 3; i++) out[i] = ((a[i] + 1.1f) * 0.45f) * diffuse;
        }
#pragma unroll
        for (int r = 0; r < 4; r++) pos[r] = dot4p(u + 4 * r, x, y, z, 1.0f);
        break;
      }
      default: {  // RF_VS_SPRITE, sprites.rs:40-45
        float view[3];
        const float vp[3] = {a[0] * 0.008f, (LT >= 2 ? a[1] : 0.0f) * 0.008f, 0.0f * 0.008f};
#pragma unroll
        for (int r = 0; r < 3; r++) view[r] = dot4p(u + 4 * r, x, y, z, 1.0f) + vp[r];
#pragma unroll
        for (int r = 0; r < 4; r++) pos[r] = dot4p(u + 16 + 4 * r, view[0], view[1], view[2], 1.0f);
        out[0] = a[0];
        if (LT >= 2) out[1] = a[1];
        break;
      }
    }
    const uint32_t oc = outcode(pos[0], pos[1], pos[2], pos[3]);
    float* o = P.cv + (size_t)gv * CVS;
    *reinterpret_cast<float4*>(o) = make_float4(pos[0], pos[1], pos[2], pos[3]);
    float rest[CVS - 4];
    rest[0] = __uint_as_float(oc);
#pragma unroll
    for (int i = 0; i < CVS - 5; i++) rest[1 + i] = (i < LT) ? out[i] : 0.0f;
#pragma unroll
    for (int q = 0; q < (CVS - 4) / 4; q++)
      *reinterpret_cast<float4*>(o + 4 + 4 * q) = make_float4(rest[4 * q], rest[4 * q + 1], rest[4 * q + 2], rest[4 * q + 3]);
    // Screen-space form of the vertex (prim.rs:62-88), once per vertex instead of once per use: to_screen is a pure function of
    // the clip vertex and the draw, so a primitive whose three vertices are inside the frustum (no clipping) reads these.
    constexpr int SVS = Rec<LT>::SVS;
    float sw[SVS];
#pragma unroll
    for (int i = 0; i < SVS; i++) sw[i] = 0.0f;
    if (!P.use_sv) continue;  // few uses per vertex somewhere in the pass: k_assemble transforms per primitive, nothing is stored here
    if (oc == 0u) {
      CVert<LT> c;
#pragma unroll
      for (int i = 0; i < 4; i++) c.p[i] = pos[i];
#pragma unroll
      for (int i = 0; i < LT; i++) c.a[i] = out[i];
      c.oc = 0u;
      SVert<LT> sc;
      to_screen<LT>(c, D.vp, D.persp_mask, sc);
      sw[0] = sc.x; sw[1] = sc.y; sw[2] = sc.z;
#pragma unroll
      for (int i = 0; i < LT; i++) sw[3 + i] = sc.a[i];
    }
    sw[3 + LT] = __uint_as_float(oc);
    float* so = P.sv + (size_t)gv * SVS;
#pragma unroll
    for (int q = 0; q < SVS / 4; q++) *reinterpret_cast<float4*>(so + 4 * q) = make_float4(sw[4 * q], sw[4 * q + 1], sw[4 * q + 2], sw[4 * q + 3]);
  }
}

template <int LT>
__device__ __forceinline__ uint32_t load_sv(const float* __restrict__ sv, uint32_t gv, SVert<LT>& s) {
  constexpr int SVS = Rec<LT>::SVS;
  const float4* q = reinterpret_cast<const float4*>(sv + (size_t)gv * SVS);
  float buf[SVS];
#pragma unroll
  for (int i = 0; i < SVS / 4; i++) {
    float4 t = __ldg(q + i);
    buf[4 * i] = t.x; buf[4 * i + 1] = t.y; buf[4 * i + 2] = t.z; buf[4 * i + 3] = t.w;
  }
  s.x = buf[0]; s.y = buf[1]; s.z = buf[2];
#pragma unroll
  for (int i = 0; i < LT; i++) s.a[i] = buf[3 + i];
  return __float_as_uint(buf[3 + LT]);
}


// One trapezoid half: everything raster.rs:248-302 precomputes, minus the unused y lane.
// lane order: 0 = x, 1 = z, 2.. = attr
template <int LT> struct HalfSetup {
  float L[2 + LT];   // running left-edge lanes (aligned to the first pixel-centre row)
  float dl[2 + LT];  // per-row step of the left edge
  float dv[2 + LT];  // dv/dx  (lane 0 unused downstream)
  float R, dr;       // running right-edge x and its step
  float y;           // first row centre (y0_rounded)
  uint32_t n;        // rows
};

template <int LT>
__device__ __forceinline__ void half_setup(float y0, float y1, const float* l0, const float* l1, const float* r0, const float* r1,
                                           HalfSetup<LT>& H) {
  constexpr int NL = 2 + LT;
  const float rdy = 1.0f / (y1 - y0);
  float dr[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) H.dl[i] = (l1[i] - l0[i]) * rdy;  // space.rs:205-207 dv_dt
#pragma unroll
  for (int i = 0; i < NL; i++) dr[i] = (r1[i] - r0[i]) * rdy;
  {
    float ls[NL], rs[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) ls[i] = l0[i] + H.dl[i];
#pragma unroll
    for (int i = 0; i < NL; i++) rs[i] = r0[i] + dr[i];
    const float dx = rs[0] - ls[0];
    const float rdx = 1.0f / dx;
#pragma unroll
    for (int i = 0; i < NL; i++) H.dv[i] = (rs[i] - ls[i]) * rdx;
  }
  const float y0r = round_up_to_half(y0), y1r = round_up_to_half(y1);
  const float tw = y0r - y0;
#pragma unroll
  for (int i = 0; i < NL; i++) H.L[i] = l0[i] + ((l0[i] + H.dl[i]) - l0[i]) * tw;  // l0.lerp(&l0.step(&dl), tw)
  H.R = r0[0] + dr[0] * tw;
  H.dr = dr[0];
  H.y = y0r;
  H.n = sat_u32(y1r - y0r);
}

// tri_fill's vertex order (raster.rs:191): stable sort of the three vertices by y with f32::total_cmp -> top, middle, bottom
__device__ __forceinline__ void tri_order(float y0, float y1, float y2, int& o0, int& o1, int& o2) {
  o0 = 0; o1 = 1; o2 = 2;
  int32_t ka = total_key(y0), kb = total_key(y1), kc = total_key(y2);
  if (kb < ka) { int t_ = o0; o0 = o1; o1 = t_; int32_t tk = ka; ka = kb; kb = tk; }
  if (kc < kb) { int t_ = o1; o1 = o2; o2 = t_; int32_t tk = kb; kb = kc; kc = tk; }
  if (kb < ka) { int t_ = o0; o0 = o1; o1 = t_; }
}

// Scanline counts of the two trapezoid halves and the first row centre, from the three y alone — the same expressions as
// half_setup's y0r / y1r / n (raster.rs:262-267), so k_assemble can classify and bin a triangle without setting it up.
__device__ __forceinline__ void tri_rows(float y0, float y1, float y2, float& yfirst, uint32_t& n0, uint32_t& n1) {
  int o0, o1, o2;
  tri_order(y0, y1, y2, o0, o1, o2);
  const float ty = o0 == 0 ? y0 : (o0 == 1 ? y1 : y2), my = o1 == 0 ? y0 : (o1 == 1 ? y1 : y2), by = o2 == 0 ? y0 : (o2 == 1 ? y1 : y2);
  const float t = round_up_to_half(ty), m = round_up_to_half(my), b = round_up_to_half(by);
  n0 = sat_u32(m - t);
  n1 = sat_u32(b - m);
  yfirst = n0 ? t : m;
}

// tri_fill (raster.rs:185-224) on a screen-triangle record (Rec<LT>::QW words: key, draw, 3 x (x, y, z, attr[LT])): vertex
// order, mid1, left/right, and the `scan` setup of both halves. Used by k_setup and — for SMALL triangles, which have no
// triangle record — by k_raster: the same instructions in the same order, so both produce the same bits.
template <int LT>
__device__ __forceinline__ void tri_setup(const uint32_t* w, HalfSetup<LT>& H0, HalfSetup<LT>& H1, float& xabs) {
  constexpr int NL = 2 + LT;
  int o0, o1, o2;
  tri_order(__uint_as_float(w[3]), __uint_as_float(w[3 + (3 + LT)]), __uint_as_float(w[3 + 2 * (3 + LT)]), o0, o1, o2);
  float top[NL], mid0[NL], bot[NL], mid1[NL];
  float ty, my, by;
#define RF_PICK(dst, yy, idx)                                                                      \
  {                                                                                                \
    _Pragma("unroll") for (int k = 0; k < 3; k++) if (k == idx) {                                  \
      dst[0] = __uint_as_float(w[2 + k * (3 + LT)]); yy = __uint_as_float(w[3 + k * (3 + LT)]);    \
      dst[1] = __uint_as_float(w[4 + k * (3 + LT)]);                                               \
      _Pragma("unroll") for (int i = 0; i < LT; i++) dst[2 + i] = __uint_as_float(w[5 + k * (3 + LT) + i]); \
    }                                                                                              \
  }
  RF_PICK(top, ty, o0)
  RF_PICK(mid0, my, o1)
  RF_PICK(bot, by, o2)
#undef RF_PICK
  const float tt = (my - ty) / (by - ty);
#pragma unroll
  for (int i = 0; i < NL; i++) mid1[i] = lerpf(top[i], bot[i], tt);
  const bool m0left = mid0[0] < mid1[0];
  const float* left = m0left ? mid0 : mid1;
  const float* right = m0left ? mid1 : mid0;
  half_setup<LT>(ty, my, top, left, top, right, H0);
  half_setup<LT>(my, by, left, bot, right, bot, H1);
  xabs = fmaxf(fmaxf(fabsf(top[0]), fabsf(mid0[0])), fabsf(bot[0]));
}

// ---------------------------------------------------------------------------------------------
// Triangle record (Rec<LT>::TW words), written by k_setup, read by k_edge_ckpt, k_walk, k_ckpt, k_raster:
//   [0] key  [1] draw  [2] sbase  [3] Yf0 (int32: first scanline's row, negative above the target)  [4] nU  [5] nL | target << 16  [6] chunk position of half 0  [7] of half 1
//   [8 + h*HS ...] half h: dv[1+LT] (dz/dx, dattr/dx), L[2+LT], dl[2+LT], R, dr, y
// ---------------------------------------------------------------------------------------------
template <int LT> struct TriRec {
  static constexpr int NL = 2 + LT, NV = 1 + LT;
  static constexpr int HS = Rec<LT>::HS;
  static constexpr int O_DV = 0, O_L = NV, O_DL = NV + NL, O_R = NV + 2 * NL, O_DR = O_R + 1, O_Y = O_R + 2;
};

// Conservative tile-column range of a triangle in one tile row, from the closed-form edge lines
// widened by a margin that bounds the drift of the reference's running sums (|sum_j - (x0 + j*dx)|
// <= j * 2^-24 * max|x|) plus the half-pixel rounding of round_up_to_half.
// Scanline j of a triangle whose first row centre is Yf0 + 0.5 lands on framebuffer row max(0, Yf0 + j): `self.y as usize`
// saturates (raster.rs:106), so every scanline above the target is drawn at row 0. The scanlines of framebuffer rows [ra, rb]:
__device__ __forceinline__ void rows_to_scanlines(int32_t Yf0, uint32_t nrows, uint32_t ra, uint32_t rb, uint32_t& ja, uint32_t& jb_excl) {
  const int32_t a = (int32_t)ra - Yf0;
  ja = (ra == 0u || a < 0) ? 0u : (uint32_t)a;  // row 0 also takes every scanline above it
  const int32_t e = (int32_t)rb + 1 - Yf0;
  jb_excl = e <= 0 ? 0u : min(nrows, (uint32_t)e);
  if (ja > jb_excl) ja = jb_excl;
}
template <int LT>
__device__ __forceinline__ void tile_row_cols(const HalfSetup<LT>& H0, const HalfSetup<LT>& H1, int32_t Yf0, uint32_t nrows, uint32_t trow,
                                              float margin, uint32_t tiles_x, uint32_t& ca, uint32_t& cb) {
  uint32_t j0, j1;  // scanlines [j0, j1) of this tile row
  rows_to_scanlines(Yf0, nrows, trow << RF_TILE_SHIFT, (trow << RF_TILE_SHIFT) + RF_TILE - 1, j0, j1);
  float lo = 3.0e38f, hi = -3.0e38f;
  const uint32_t nU = H0.n;
  if (j0 < j1 && j0 < nU) {  // rows of the upper half
    const float ja = (float)j0, jb = (float)(min(j1, nU) - 1u);
    lo = fminf(lo, fminf(H0.L[0] + H0.dl[0] * ja, H0.L[0] + H0.dl[0] * jb));
    hi = fmaxf(hi, fmaxf(H0.R + H0.dr * ja, H0.R + H0.dr * jb));
  }
  if (j0 < j1 && j1 > nU) {  // rows of the lower half
    const float ja = (float)(max(j0, nU) - nU), jb = (float)(j1 - 1u - nU);
    lo = fminf(lo, fminf(H1.L[0] + H1.dl[0] * ja, H1.L[0] + H1.dl[0] * jb));
    hi = fmaxf(hi, fmaxf(H1.R + H1.dr * ja, H1.R + H1.dr * jb));
  }
  const float fa = floorf(lo - 0.5f - margin), fb = floorf(hi + 0.5f + margin);
  if (!(fa <= fb) || !(fabsf(fa) < 1.0e9f) || !(fabsf(fb) < 1.0e9f)) { ca = 0; cb = tiles_x - 1; return; }  // NaN/inf: whole row
  ca = min(sat_u32(fa) >> RF_TILE_SHIFT, tiles_x - 1);
  cb = min(sat_u32(fb) >> RF_TILE_SHIFT, tiles_x - 1);
}

// Serial walk of one trapezoid half by the thread that set it up (ScanlineIter::next, raster.rs:80-114).
// Used for triangles with few rows, where a separate row-parallel kernel costs more than it saves.
template <int LT>
__device__ __forceinline__ void walk_half_inline(uint32_t* span_base, uint32_t th, uint32_t tw, uint32_t by0, uint32_t by1, HalfSetup<LT>& H,
                                                 uint32_t& sidx, uint32_t& row, uint32_t& long_rows, unsigned long long& frags_i, bool& oob) {
  constexpr int NL = 2 + LT;
  constexpr int SW = Rec<LT>::SW;
  float y = H.y;
  for (uint32_t j = 0; j < H.n; j++) {
    float v0[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) { v0[i] = H.L[i]; H.L[i] = H.L[i] + H.dl[i]; }
    const float x1 = H.R;
    H.R = H.R + H.dr;
    const float x0r = round_up_to_half(v0[0]), x1r = round_up_to_half(x1);
    const float tx = x0r - v0[0];
    uint32_t w[SW];
#pragma unroll
    for (int i = 1; i < NL; i++) w[2 + (i - 1)] = __float_as_uint(v0[i] + ((v0[i] + H.dv[i]) - v0[i]) * tx);
#pragma unroll
    for (int i = 2 + NL - 1; i < SW; i++) w[i] = 0u;
    const uint32_t cnt = sat_u32(x1r - x0r);
    const uint32_t Yf = sat_u32(y), X0 = sat_u32(x0r), X1 = max(sat_u32(x1r), X0);
    uint32_t nn = min(cnt, X1 - X0);
    if (Yf >= th || X1 > tw) {  // target.rs:148,173-174 (slice index panics)
      oob = true;
      nn = 0;
    } else if (Yf < by0 || Yf >= by1) {
      nn = 0;  // not this GPU's row band
    } else {
      frags_i += X1 - X0;
    }
    // crosses a tile column: remembered in a bit mask (row <= RF_INLINE_ROWS), appended to the long list after the walk
    if (nn && ((X0 + nn - 1) >> RF_TILE_SHIFT) != (X0 >> RF_TILE_SHIFT)) long_rows |= 1u << row;
    w[0] = X0 | nn << 16;
    w[1] = RF_NO_CKPT;
    uint32_t* sr = span_base + (size_t)sidx * SW;  // span record `sidx` of span_base (the global array, or the warp's staging buffer)
#pragma unroll
    for (int q = 0; q < SW / 2; q++) *reinterpret_cast<uint2*>(sr + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
    sidx++;
    row++;
    y = y + 1.0f;
  }
}

// raster::line (raster.rs:122-177): a one-pixel-thick line, every pixel carrying v0's position and varyings. The
// pixels of one framebuffer row are consecutive, so the line is stored like a triangle: one span record per row
// with zero dv/dx. PASS 0 measures (row range, bin entries), PASS 1 writes span records and bin entries.
struct LineGeom {
  bool wide;
  float start, step;   // wide: y at the first pixel centre and dy/dx; tall: x at the first row centre and dx/dy
  uint32_t ia, ib;     // wide: pixel columns [ia, ib); tall: rows [ia, ib)
};
struct LineMeasure {
  uint32_t ymin, ymax, nent;
  bool oob;
};

template <int LT, int PASS>
__device__ __forceinline__ void line_walk(const PassParams& P, const TargetDesc& T, const LineGeom& G, LineMeasure& M, uint32_t lane,
                                          uint32_t sbase, uint32_t ebase, uint32_t key, uint32_t tri_idx, const float* lanes /*z, attr*/,
                                          unsigned long long& frags_i) {
  constexpr int SW = Rec<LT>::SW;
  float acc = G.start;
  uint32_t run_y = 0xFFFFFFFFu, run_x0 = 0, run_n = 0;  // current row run
  uint32_t trow = 0xFFFFFFFFu, cmin = 0xFFFFFFFFu, cmax = 0;
  uint32_t eidx = ebase;
  auto flush_tile_row = [&]() {
    if (trow != 0xFFFFFFFFu && cmin <= cmax) {
      if (PASS == 0) M.nent += cmax - cmin + 1;
      else {
        const uint32_t tbase = T.tile_base + trow * T.tiles_x;
        for (uint32_t c = cmin; c <= cmax; c++) { P.entries[eidx++] = make_uint4(tbase + c, key, tri_idx, 0u); atomicAdd(P.tile_cnt + tbase + c, 1u); }
      }
    }
    cmin = 0xFFFFFFFFu; cmax = 0;
  };
  auto flush_run = [&]() {
    if (run_n == 0) return;
    const bool in_band = run_y >= T.band_y0 && run_y < T.band_y1;
    if (PASS == 0) { M.ymin = min(M.ymin, run_y); M.ymax = max(M.ymax, run_y); }
    if (in_band) {
      if ((run_y >> RF_TILE_SHIFT) != trow) { flush_tile_row(); trow = run_y >> RF_TILE_SHIFT; }
      cmin = min(cmin, run_x0 >> RF_TILE_SHIFT);
      cmax = max(cmax, (run_x0 + run_n - 1) >> RF_TILE_SHIFT);
    }
    if (PASS == 1) {
      const uint32_t nn = in_band ? run_n : 0u;
      if (in_band) frags_i += run_n;
      const uint32_t sidx = sbase + (run_y - M.ymin);
      uint32_t w[SW];
      w[0] = run_x0 | nn << 16; w[1] = RF_NO_CKPT;
#pragma unroll
      for (int i = 0; i < 1 + LT; i++) w[2 + i] = __float_as_uint(lanes[i]);
#pragma unroll
      for (int i = 3 + LT; i < SW; i++) w[i] = 0u;
      uint32_t* sr = P.spans + (size_t)sidx * SW;
#pragma unroll
      for (int q = 0; q < SW / 2; q++) *reinterpret_cast<uint2*>(sr + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
      if (nn && ((run_x0 + nn - 1) >> RF_TILE_SHIFT) != (run_x0 >> RF_TILE_SHIFT)) {
        const unsigned long long slot = agg_atomic_inc(&P.status->long_needed, lane);
        if (slot < P.cap_long) P.longlist[slot] = make_uint2(sidx, tri_idx * 2u);
        else rf_overflow(P);
      }
    }
    run_n = 0;
  };
  for (uint32_t i = G.ia; i < G.ib; i++) {
    const uint32_t py = G.wide ? sat_u32(acc) : i;
    const uint32_t px = G.wide ? i : sat_u32(acc);
    if (py >= T.h || px + 1 > T.w) { M.oob = true; break; }  // target.rs:148,173-174
    if (py != run_y || px != run_x0 + run_n) { flush_run(); run_y = py; run_x0 = px; }
    run_n++;
    acc = acc + G.step;
  }
  flush_run();
  flush_tile_row();
}

// Triangles with at most this many scanlines are walked inline by k_setup; taller ones go to k_walk in chunks.
#ifndef RF_INLINE_ROWS
#define RF_INLINE_ROWS 12u
#endif

// =============================================================================================
// K2a k_assemble: one thread per input primitive — assembly (render.rs:168-172), status / Sutherland–
// Hodgman clip (clip.rs:350-400), to_screen (prim.rs:62-88), winding + face cull (ctx.rs:95-101).
// The surviving screen-space triangles are appended to a COMPACT buffer with one atomic per warp
// (warp ballot / prefix sum), so the setup kernel that follows runs with every lane busy.
// Screen triangle record (Rec<LT>::QW words): key, draw, then 3 x (x, y, z, attr[LT]).
// =============================================================================================
#define RF_CHUNK 32u
#define RF_LONG_BLOCK 256u

#ifndef RF_ASSEMBLE_MIN_BLOCKS
#define RF_ASSEMBLE_MIN_BLOCKS 4   // <= 128 registers: measured -2 % on the bunny step against the unconstrained 147
#endif
// Optional (off): stage the screen-triangle records a warp appends in shared memory and write them out by the whole warp,
// every sector once, as k_setup does with its records. Measured on the bunny batch it costs more than it saves here
// (k_assemble 0.41 -> 0.49 ms, profiles/r01_ab2_assemble_prefetch.txt): the kernel is bound by its IEEE divides and gather
// latency, not by L2 write sectors, and the two extra warp barriers per fan index sit on its critical path.
#ifndef RF_ASSEMBLE_STAGE
#define RF_ASSEMBLE_STAGE 0
#endif
template <int LT, bool SV>  // SV: every draw of the pass carries RF_F_SV (k_vertex stored screen-space vertices)
__global__ void __launch_bounds__(128, RF_ASSEMBLE_MIN_BLOCKS) k_assemble(PassParams P) {
  constexpr int QW = Rec<LT>::QW;
  // record stride QW = 20 / 28 / 36 words: the 128-bit accesses of 8 consecutive lanes fall into 8 different bank groups
  __shared__ uint4 s_q[RF_ASSEMBLE_STAGE ? 4 : 1][RF_ASSEMBLE_STAGE ? 32 * QW / 4 : 1];
  if (rf_poisoned(P)) return;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u;
  const uint32_t n_iter = (P.NP + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t gp = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const bool have = gp < P.NP;
    uint32_t d = 0, ntri = 0;
    CVert<LT> c0, c1;             // edge endpoints while they are clipped
    SVert<LT> sv0, sv1, sv2;      // screen vertices of an unclipped primitive (from k_vertex)
    uint32_t gv0 = 0, gv1 = 0, gv2 = 0;
    CVert<LT> poly[10], tmp[10];
    bool clipped = false;   // the primitive's vertices are in poly[] in clip space and go through to_screen at emission
    bool is_edge = false;
    constexpr bool use_sv = SV;
    if (have) {
      d = find_draw(P.pbase, P.n_draws, gp, P.prims_per_draw);
      const DrawDesc& D = P.draws[d];
      is_edge = D.prim_kind == RF_PRIM_EDGES;
    }
    const bool culled_obj = have && P.any_bbox && P.dstats[d].hidden != 0ull;  // k_objects: render() is not called at all
    if (culled_obj) {
    } else if (have && is_edge) {
      // Render for Edge<usize> (prim.rs:41-60): inline two vertices, Clip for [Edge] (clip.rs:311-348)
      const DrawDesc& D = P.draws[d];
      const uint32_t* ip = D.indices + 2 * (size_t)(gp - __ldg(P.pbase + d));
      const uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1);
      if (i0 >= D.n_verts || i1 >= D.n_verts) {
        atomicOr(&P.status->error, RF_ERRBIT_INDEX_OOB);
      } else {
        const uint32_t vb = __ldg(P.vbase + d);
        gv0 = vb + i0; gv1 = vb + i1; gv2 = gv0;
        uint32_t oc0, oc1;
        if (use_sv) {
          oc0 = load_sv<LT>(P.sv, gv0, sv0); oc1 = load_sv<LT>(P.sv, gv1, sv1);
          sv2 = sv0;  // unused third vertex
        } else {
          load_cv<LT>(P.cv, gv0, c0); load_cv<LT>(P.cv, gv1, c1);
          oc0 = c0.oc; oc1 = c1.oc;
        }
        if ((oc0 & oc1) != 0) ntri = 0;       // both outside one plane
        else if ((oc0 | oc1) == 0) {          // neither outside
          ntri = 1;
          clipped = !use_sv;                  // with RF_F_SV the screen vertices k_vertex stored are used as they are
        } else {
          ntri = 1;
          clipped = true;
          if (use_sv) { load_cv<LT>(P.cv, gv0, c0); load_cv<LT>(P.cv, gv1, c1); }
          for (int pl = 0; pl < 6; pl++) {
            const uint32_t bit = 1u << pl;
            const bool a_in = (c0.oc & bit) == 0, b_in = (c1.oc & bit) == 0;
            if (!a_in && !b_in) { ntri = 0; break; }
            CVert<LT> x;
            if (clip_intersect<LT>(pl, c0, c1, x)) {
              if (a_in) c1 = x;
              else if (b_in) c0 = x;
            }
          }
        }
        if (clipped) { poly[0] = c0; poly[1] = c1; poly[2] = c0; }  // emission reads poly[0], poly[t + 1], poly[t + 2] with t = 0
      }
    } else if (have) {
      const DrawDesc& D = P.draws[d];
      const uint32_t* ip = D.indices + 3 * (size_t)(gp - __ldg(P.pbase + d));
      const uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
      if (i0 >= D.n_verts || i1 >= D.n_verts || i2 >= D.n_verts) {
        atomicOr(&P.status->error, RF_ERRBIT_INDEX_OOB);  // prim.rs:17-19 panics
      } else {
        const uint32_t vb = __ldg(P.vbase + d);
        gv0 = vb + i0; gv1 = vb + i1; gv2 = vb + i2;
        uint32_t oc0, oc1, oc2;
        if (use_sv) {
          oc0 = load_sv<LT>(P.sv, gv0, sv0); oc1 = load_sv<LT>(P.sv, gv1, sv1); oc2 = load_sv<LT>(P.sv, gv2, sv2);
        } else {
          load_cv<LT>(P.cv, gv0, poly[0]); load_cv<LT>(P.cv, gv1, poly[1]); load_cv<LT>(P.cv, gv2, poly[2]);
          oc0 = poly[0].oc; oc1 = poly[1].oc; oc2 = poly[2].oc;
        }
        const uint32_t all = oc0 & oc1 & oc2, any = oc0 | oc1 | oc2;
        if (all != 0) ntri = 0;          // Status::Hidden, clip.rs:245-267
        else if (any == 0) {             // Status::Visible
          ntri = 1;
          clipped = !use_sv;             // with RF_F_SV the screen vertices k_vertex stored are used as they are; else poly[0..2]
        } else {
          clipped = true;
          if (use_sv) { load_cv<LT>(P.cv, gv0, poly[0]); load_cv<LT>(P.cv, gv1, poly[1]); load_cv<LT>(P.cv, gv2, poly[2]); }
          const int n = clip_polygon<LT>(poly, tmp);
          ntri = n >= 3 ? (uint32_t)(n - 2) : 0u;  // fan (p0, pk, pk+1), clip.rs:375-394
        }
      }
    }
    uint32_t max_tri = ntri;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_tri = max(max_tri, __shfl_xor_sync(0xFFFFFFFFu, max_tri, o));
    uint32_t my_prims_o = 0;
    for (uint32_t t = 0; t < max_tri; t++) {
      bool emit = false;
      float depth = 0.0f;
      SVert<LT> s[3];
      if (t < ntri) {
        const DrawDesc& D = P.draws[d];
        if (clipped) {
          to_screen<LT>(poly[0], D.vp, D.persp_mask, s[0]);
          to_screen<LT>(poly[t + 1], D.vp, D.persp_mask, s[1]);
          to_screen<LT>(poly[t + 2], D.vp, D.persp_mask, s[2]);
        } else {
          s[0] = sv0; s[1] = sv1; s[2] = sv2;
        }
        // Tri::winding geom/prim.rs:288-294 ; Context::face_cull ctx.rs:95-101
        const float abx = s[1].x - s[0].x, aby = s[1].y - s[0].y;
        const float acx = s[2].x - s[0].x, acy = s[2].y - s[0].y;
        float wz = 0.0f;
        wz = wz + (-aby) * acx;
        wz = wz + abx * acy;
        const bool back = is_edge ? false : wz < 0.0f;  // Render::is_backface defaults to false for edges (render.rs:72-74)
        const uint32_t cull = D.flags & RF_F_CULL_MASK;
        emit = !((cull == RF_CULL_BACK && back) || (cull == RF_CULL_FRONT && !back));
        if (emit) my_prims_o++;  // render.rs:195-196: counted before rasterisation, whatever it covers
        if (emit && P.sdepth != nullptr) {
          // Render::depth: prim.rs:21-23 for triangles (clip-space z, left to right, then / 3.0); f32::INFINITY for edges
          if (is_edge) depth = __int_as_float(0x7F800000);
          else if (clipped) depth = ((poly[0].p[2] + poly[t + 1].p[2]) + poly[t + 2].p[2]) / 3.0f;  // also the unclipped triangle without RF_F_SV
          else {  // unclipped: clip-space z of the three vertices
            constexpr int CVS = Rec<LT>::CVS;
            depth = ((__ldg(P.cv + (size_t)gv0 * CVS + 2) + __ldg(P.cv + (size_t)gv1 * CVS + 2)) + __ldg(P.cv + (size_t)gv2 * CVS + 2)) / 3.0f;
          }
        }
      }
      if (__ballot_sync(0xFFFFFFFFu, emit) == 0u) continue;
      // ---- SMALL or LARGE (see RF_BIN_SMALL). A SMALL triangle has few scanlines, a narrow bounding box and lies inside the
      // target, so no scanline can leave the target (target.rs:148,173-174 cannot panic) and its pixels lie in the tiles of
      // its bounding box: it is set up (tri_fill, raster.rs:185-302) and binned right here, and k_raster walks its scanlines
      // from that record. Everything else — also anything with a non-finite coordinate — goes to k_setup as a screen triangle.
      bool small = false;
      uint32_t s_Y0 = 0, s_n0 = 0, s_n1 = 0, s_tr0 = 1, s_tr1 = 0, s_ca = 0, s_cb = 0, s_tbase = 0, s_tx = 1, nent = 0;
      if (emit && !is_edge && RF_SMALL_ROWS != 0u && P.sdepth == nullptr) {
        float yfirst;
        tri_rows(s[0].y, s[1].y, s[2].y, yfirst, s_n0, s_n1);
        const uint32_t nrows = s_n0 + s_n1;
        if (nrows == 0u) emit = false;  // no pixel-centre row between its vertices: counted above (render.rs:195-196), nothing to draw
        const float xmin = fminf(fminf(s[0].x, s[1].x), s[2].x), xmax = fmaxf(fmaxf(s[0].x, s[1].x), s[2].x);
        const float ylo = fminf(fminf(s[0].y, s[1].y), s[2].y), yhi = fmaxf(fmaxf(s[0].y, s[1].y), s[2].y);
        bool finite = true;
#pragma unroll
        for (int k = 0; k < 3; k++) finite = finite && fabsf(s[k].x) < 1.0e9f && fabsf(s[k].y) < 1.0e9f;  // false for NaN
        const TargetDesc& T = P.targets[P.draws[d].target];
        // the running sums of an edge stay within the bounding box up to rounding drift (|sum_j - (x0 + j * dx)| <= j * 2^-24 * max|x|)
        const float margin = 1.0f + (float)nrows * fmaxf(fabsf(xmin), fabsf(xmax)) * 1.2e-7f;
        if (emit && finite && nrows <= RF_SMALL_ROWS && xmax - xmin <= RF_SMALL_WIDTH && ylo >= 0.0f && yhi + 1.0f < (float)T.h &&
            xmin - margin > 0.0f && xmax + margin < (float)T.w) {
          small = true;
          s_Y0 = sat_u32(yfirst);
          const uint32_t Ya = max(s_Y0, T.band_y0), Yb = min(s_Y0 + nrows, T.band_y1);  // only tile rows of this GPU's row band
          if (Ya < Yb) {
            s_tr0 = Ya >> RF_TILE_SHIFT; s_tr1 = (Yb - 1) >> RF_TILE_SHIFT;
            s_ca = min(sat_u32(floorf(xmin - 0.5f - margin)) >> RF_TILE_SHIFT, T.tiles_x - 1);
            s_cb = min(sat_u32(floorf(xmax + 0.5f + margin)) >> RF_TILE_SHIFT, T.tiles_x - 1);
            s_tbase = T.tile_base; s_tx = T.tiles_x;
            nent = (s_tr1 - s_tr0 + 1) * (s_cb - s_ca + 1);
          } else {
            emit = false; small = false;  // sort-first sharding: no scanline in this GPU's row band
          }
        }
      }
      const uint32_t lmask = __ballot_sync(0xFFFFFFFFu, emit && !small), smask = __ballot_sync(0xFFFFFFFFu, small);
      if ((lmask | smask) == 0u) continue;
      const uint32_t incl_e = warp_scan_incl(nent, lane), tot_e = __shfl_sync(0xFFFFFFFFu, incl_e, 31);
      unsigned long long base = 0, sbase = 0, ebase = 0;
      if (lane == 0) {
        if (lmask) base = atomicAdd(&P.status->stris_needed, (unsigned long long)__popc(lmask));
        if (smask) sbase = atomicAdd(&P.status->small_needed, (unsigned long long)__popc(smask));
        if (tot_e) ebase = atomicAdd(&P.status->entries_needed, (unsigned long long)tot_e);
      }
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      sbase = __shfl_sync(0xFFFFFFFFu, sbase, 0);
      ebase = __shfl_sync(0xFFFFFFFFu, ebase, 0);
      if (base + __popc(lmask) > P.cap_stris || sbase + __popc(smask) > P.cap_smalls || ebase + tot_e > P.cap_entries) {
        if (lane == 0) rf_overflow(P);
        continue;
      }
      const uint32_t emask = lmask;  // screen-triangle records: the LARGE ones only
      uint32_t* const qstg = reinterpret_cast<uint32_t*>(s_q[RF_ASSEMBLE_STAGE ? threadIdx.x >> 5 : 0]);
      if (emit) {
        uint32_t w[QW];
        w[0] = gp * 8u + t; w[1] = is_edge ? (d | RF_STRI_LINE) : d;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          w[2 + k * (3 + LT)] = __float_as_uint(s[k].x); w[3 + k * (3 + LT)] = __float_as_uint(s[k].y); w[4 + k * (3 + LT)] = __float_as_uint(s[k].z);
#pragma unroll
          for (int i = 0; i < LT; i++) w[5 + k * (3 + LT) + i] = __float_as_uint(s[k].a[i]);
        }
#pragma unroll
        for (int i = 2 + 3 * (3 + LT); i < QW; i++) w[i] = 0u;
        if (small) {
          using SR = SmallRec<LT>;
          HalfSetup<LT> H0, H1;
          float xabs;
          tri_setup<LT>(w, H0, H1, xabs);
          const uint32_t sidx = (uint32_t)sbase + __popc(smask & lt);
          uint32_t* r = P.smalls + (size_t)sidx * SR::W;
          *reinterpret_cast<uint4*>(r) = make_uint4(w[0], d, s_Y0, s_n0 | s_n1 << 16);
#pragma unroll
          for (int hh = 0; hh < 2; hh++) {
            const HalfSetup<LT>& H = hh ? H1 : H0;
            uint32_t hw[SR::HW];
#pragma unroll
            for (int i = 0; i < SR::HW; i++) hw[i] = 0u;
#pragma unroll
            for (int i = 0; i < SR::NV; i++) hw[SR::O_DV + i] = __float_as_uint(H.dv[1 + i]);
#pragma unroll
            for (int i = 0; i < SR::NL; i++) { hw[SR::O_L + i] = __float_as_uint(H.L[i]); hw[SR::O_DL + i] = __float_as_uint(H.dl[i]); }
            hw[SR::O_R] = __float_as_uint(H.R); hw[SR::O_DR] = __float_as_uint(H.dr);
#pragma unroll
            for (int qd = 0; qd < SR::HW / 4; qd++) *reinterpret_cast<uint4*>(r + 4 + hh * SR::HW + 4 * qd) = make_uint4(hw[4 * qd], hw[4 * qd + 1], hw[4 * qd + 2], hw[4 * qd + 3]);
          }
          // bin entries: every tile of the bounding box (inside this GPU's row band)
          uint32_t eidx = (uint32_t)ebase + (incl_e - nent);
          for (uint32_t tr = s_tr0; tr <= s_tr1; tr++)
            for (uint32_t c = s_ca; c <= s_cb; c++) {
              const uint32_t tile = s_tbase + tr * s_tx + c;
              P.entries[eidx++] = make_uint4(tile, w[0], sidx | RF_BIN_SMALL, 0u);
              atomicAdd(P.tile_cnt + tile, 1u);
            }
        } else {
          uint32_t* q = RF_ASSEMBLE_STAGE ? qstg + __popc(emask & lt) * QW : P.stris + (size_t)((uint32_t)base + __popc(emask & lt)) * QW;
          if (P.sdepth != nullptr) P.sdepth[(uint32_t)base + __popc(emask & lt)] = (uint32_t)total_key(depth) ^ 0x80000000u;  // unsigned order == total_cmp
#pragma unroll
          for (int qd = 0; qd < QW / 4; qd++) *reinterpret_cast<uint4*>(q + 4 * qd) = make_uint4(w[4 * qd], w[4 * qd + 1], w[4 * qd + 2], w[4 * qd + 3]);
        }
      }
      if (RF_ASSEMBLE_STAGE) {
        __syncwarp();
        const uint32_t nq = (uint32_t)__popc(emask) * (QW / 4);
        uint4* dst = reinterpret_cast<uint4*>(P.stris + (size_t)(uint32_t)base * QW);
        for (uint32_t qi = lane; qi < nq; qi += 32) dst[qi] = reinterpret_cast<const uint4*>(qstg)[qi];
        __syncwarp();  // the buffer is reused by the next fan index
      }
    }
    // ---- per-draw prims.o: aggregate over the warp when every lane has the same draw
    {
      const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, d, 0);
      const bool uniform = __all_sync(0xFFFFFFFFu, !have || d == d0);
      if (uniform) {
        uint32_t po = my_prims_o;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) po += __shfl_xor_sync(0xFFFFFFFFu, po, o);
        if (lane == 0 && po) atomicAdd(&P.dstats[d0].prims_o, (unsigned long long)po);
      } else if (have && my_prims_o) {
        atomicAdd(&P.dstats[d].prims_o, (unsigned long long)my_prims_o);
      }
    }
  }
}

// =============================================================================================
// K2a' k_setup: one thread per surviving screen triangle — tri_fill's y-sort and the `scan` setup of
// both trapezoid halves (raster.rs:185-302). Emits, with warp-aggregated allocation (warp prefix
// sums): a triangle record, its span range, its (triangle x tile) bin entries, and either walks
// its scanlines inline (few rows) or cuts them into <= 32-row chunks for k_walk.
// =============================================================================================
#ifndef RF_SETUP_MIN_BLOCKS
#define RF_SETUP_MIN_BLOCKS 4
#endif
// Triangle records and the span records of inline-walked triangles are staged in shared memory and written out by the
// whole warp as contiguous runs, every 32-byte sector once. Written by each lane directly (192-byte records at a 192-byte
// lane stride, 24-byte spans as three 8-byte stores at a stride of rows x 24 bytes) the same data cost 2-3 L2 sector writes
// per sector of payload: 111 M of the kernel's 135 M L2 write sectors on the bunny batch, 71 M of them excess
// (profiles/r01_hot_lines.txt), with L2 the busiest unit of the kernel. Only at 3 varying lanes (static shared memory).
#ifndef RF_SETUP_STAGE
#define RF_SETUP_STAGE 1
#endif
template <int LT> struct SetupStage {
  static constexpr bool ON = RF_SETUP_STAGE && LT == 3;
  static constexpr int TWP = Rec<LT>::TW + 4;  // padded record stride in the buffer: 52 words = 20 mod 32 -> 128-bit accesses of 8 lanes hit 32 different banks
  static constexpr int SPAN_WORDS = 32 * (int)RF_INLINE_ROWS * Rec<LT>::SW, TRI_WORDS = 32 * TWP;
  static constexpr int WORDS = ON ? (SPAN_WORDS > TRI_WORDS ? SPAN_WORDS : TRI_WORDS) : 4;
};
template <int LT>
__global__ void __launch_bounds__(128, LT == 3 ? RF_SETUP_MIN_BLOCKS : 3) k_setup(PassParams P) {
  constexpr int NL = 2 + LT, NV = 1 + LT;
  constexpr int TW = Rec<LT>::TW, HS = Rec<LT>::HS, QW = Rec<LT>::QW;
  using TR = TriRec<LT>;
  __shared__ uint32_t s_tot[4][4];
  __shared__ unsigned long long s_base[4];
  __shared__ uint32_t s_fit[4];
  using SS = SetupStage<LT>;
  __shared__ uint4 s_stage[4][SS::WORDS / 4];  // uint4: 16-byte aligned for the 128-bit accesses
  // The poison test must be ONE decision per block: other blocks of this very kernel set the poison when they overflow, and a block
  // whose warps read it on either side of that moment lost some warps while the rest went on to the block-wide allocation below with
  // the missing warps' s_tot entries uninitialised — garbage counters (arena requests of hundreds of GB) and triangle records
  // written far outside the arena (compute-sanitizer: profiles/r02_memcheck.txt). Seen when a 32-frame crates batch followed the
  // bunny batch in one context.
  __shared__ uint32_t s_poisoned;
  if (threadIdx.x == 0) s_poisoned = rf_poisoned(P) ? 1u : 0u;
  __syncthreads();
  if (s_poisoned) return;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u;
  // the triangles k_assemble did not set up and bin itself (see RF_BIN_SMALL): tall or wide ones, those near the target's edges, lines
  const uint32_t NT = (uint32_t)min(P.status->stris_needed, (unsigned long long)P.cap_stris);
  const uint32_t n_iter = (NT + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t ti = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const bool have = ti < NT;
    bool emit = false;
    HalfSetup<LT> H0, H1;
    H0.n = H1.n = 0;
    int32_t Yf0 = 0;  // first scanline's framebuffer row, signed (see rows_to_scanlines)
    uint32_t key = 0, d = 0, tgt = 0, tr0 = 0, tr1 = 0, tiles_x = 1, nent = 0, nchunk = 0;
    float margin = 0.0f;
    unsigned long long my_frags_i = 0;
    uint32_t long_rows = 0, long_sbase = 0, long_tri = 0, long_nU = 0;  // inline-walked rows that cross a tile column
    uint32_t t_h = 0, t_w = 0, t_by0 = 0, t_by1 = 0;
    bool is_line = false;
    LineGeom LG{};
    LineMeasure LM{};
    float line_lanes[1 + LT];
    if (have) {
      const uint32_t* q = P.stris + (size_t)ti * QW;
      uint32_t w[QW];
#pragma unroll
      for (int qd = 0; qd < QW / 4; qd++) {
        const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(q) + qd);
        w[4 * qd] = t4.x; w[4 * qd + 1] = t4.y; w[4 * qd + 2] = t4.z; w[4 * qd + 3] = t4.w;
      }
      key = w[0]; d = w[1] & ~RF_STRI_LINE;
      is_line = (w[1] & RF_STRI_LINE) != 0;
      const DrawDesc& D = P.draws[d];
      if (is_line) {
        // raster::line (raster.rs:122-177)
        int a = 0, b = 1;
        if (__uint_as_float(w[3]) > __uint_as_float(w[3 + (3 + LT)])) { a = 1; b = 0; }
        const float ax = __uint_as_float(w[2 + a * (3 + LT)]), ay = __uint_as_float(w[3 + a * (3 + LT)]);
        const float bx = __uint_as_float(w[2 + b * (3 + LT)]), by_ = __uint_as_float(w[3 + b * (3 + LT)]);
        const float dx = bx - ax, dy = by_ - ay;
        LG.wide = fabsf(dx) > dy;
        int v0i = a;
        if (LG.wide) {
          float v0x = ax, v0y = ay, v1x = bx;
          if (dx < 0.0f) { v0i = b; v0x = bx; v0y = by_; v1x = ax; }  // draw left to right; dx, dy keep their old values
          const float x0 = round_up_to_half(v0x), x1 = round_up_to_half(v1x);
          LG.step = dy / dx;
          LG.start = v0y + LG.step * (x0 - v0x);
          LG.ia = sat_u32(x0); LG.ib = sat_u32(x1);
        } else {
          const float y0 = round_up_to_half(ay), y1 = round_up_to_half(by_);
          LG.step = dx / dy;
          LG.start = ax + LG.step * (y0 - ay);
          LG.ia = sat_u32(y0); LG.ib = sat_u32(y1);
        }
        line_lanes[0] = __uint_as_float(w[4 + v0i * (3 + LT)]);
#pragma unroll
        for (int i = 0; i < LT; i++) line_lanes[1 + i] = __uint_as_float(w[5 + v0i * (3 + LT) + i]);
        const TargetDesc& T = P.targets[D.target];
        tgt = D.target; tiles_x = T.tiles_x;
        if (LG.ib > LG.ia && LG.ib - LG.ia > RF_MAX_ROWS) { atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB); LG.ib = LG.ia; }
        LM.ymin = 0xFFFFFFFFu; LM.ymax = 0; LM.nent = 0; LM.oob = false;
        line_walk<LT, 0>(P, T, LG, LM, lane, 0u, 0u, key, 0u, line_lanes, my_frags_i);
        if (LM.oob) { atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB); }
        else if (LM.ymin <= LM.ymax) {
          emit = true;
          Yf0 = (int32_t)LM.ymin;
          H0.n = LM.ymax - LM.ymin + 1; H1.n = 0;
          nent = LM.nent;
#pragma unroll
          for (int i = 0; i < NL; i++) { H0.dv[i] = 0.0f; H0.L[i] = 0.0f; H0.dl[i] = 0.0f; H1.dv[i] = 0.0f; H1.L[i] = 0.0f; H1.dl[i] = 0.0f; }
          H0.R = H0.dr = H0.y = 0.0f; H1.R = H1.dr = H1.y = 0.0f;
        }
      } else {
      float xabs;
      tri_setup<LT>(w, H0, H1, xabs);  // tri_fill raster.rs:185-224
      const TargetDesc& T = P.targets[D.target];
      tgt = D.target; tiles_x = T.tiles_x;
      t_h = T.h; t_w = T.w; t_by0 = T.band_y0; t_by1 = T.band_y1;
      // Row-range sanity. A scanline at y >= h panics in the reference (target.rs:148,173). Scanlines ABOVE the target (a
      // viewport partly outside it) are drawn at row 0, one after the other, as `self.y as usize` does (raster.rs:106); a
      // trapezoid half is bounded at 65,535 scanlines (record field width) — more than 32,767 rows above the largest target.
      uint32_t nrows = H0.n + H1.n;
      if (nrows != 0) {
        const float yfirst = H0.n ? H0.y : H1.y;
        const float ylast = yfirst + (float)(nrows - 1);
        if (H0.n > 65535u || H1.n > 65535u || sat_u32(ylast) >= T.h) {
          atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB);
          nrows = 0;
        }
        if (nrows == 0) H0.n = H1.n = 0;
        else {
          emit = true;
          Yf0 = (int32_t)floorf(yfirst);  // row centres are k + 0.5; |yfirst| < 2^17 here
          // framebuffer rows [max(Yf0, 0), max(Yf0 + nrows, 1)); only tile rows inside this GPU's row band get bin entries
          const uint32_t Ya = max((uint32_t)max(Yf0, 0), T.band_y0), Yb = min((uint32_t)max(Yf0 + (int32_t)nrows, 1), T.band_y1);
          margin = 1.0f + (float)nrows * xabs * 1.2e-7f;
          if (Ya < Yb) {
            tr0 = Ya >> RF_TILE_SHIFT; tr1 = (Yb - 1) >> RF_TILE_SHIFT;
            for (uint32_t tr = tr0; tr <= tr1; tr++) {
              uint32_t ca, cb;
              tile_row_cols<LT>(H0, H1, Yf0, nrows, tr, margin, tiles_x, ca, cb);
              nent += cb - ca + 1;
            }
          } else {
            // sort-first sharding: no scanline of this triangle is in this GPU's row band -> nothing to walk or rasterise
            // (its x-extent can no longer raise RF_E_TARGET_OOB here; the rank that owns the rows reports it)
            emit = false; H0.n = H1.n = 0; tr0 = 1; tr1 = 0;
          }
          if (emit && nrows > RF_INLINE_ROWS) nchunk = (H0.n + RF_CHUNK - 1) / RF_CHUNK + (H1.n + RF_CHUNK - 1) / RF_CHUNK;
        }
      }
      }  // triangle
    }
    // ---- block-aggregated allocation: warp prefix sums, then ONE atomic per counter per block
    const uint32_t nsp = emit ? (H0.n + H1.n) : 0u;
    const uint32_t incl_s = warp_scan_incl(nsp, lane), incl_e = warp_scan_incl(nent, lane), incl_c = warp_scan_incl(nchunk, lane);
    const uint32_t emask = __ballot_sync(0xFFFFFFFFu, emit);
    const uint32_t wid = threadIdx.x >> 5;
    if (lane == 31) { s_tot[wid][0] = incl_s; s_tot[wid][1] = __popc(emask); s_tot[wid][2] = incl_e; s_tot[wid][3] = incl_c; }
    __syncthreads();
    if (threadIdx.x < 4) {  // thread k allocates counter k for the whole block
      uint32_t tot = 0;
#pragma unroll
      for (int w = 0; w < 4; w++) tot += s_tot[w][threadIdx.x];
      PaddedCounter* ctr = threadIdx.x == 0 ? &P.status->spans_needed : threadIdx.x == 1 ? &P.status->tris_needed
                         : threadIdx.x == 2 ? &P.status->entries_needed : &P.status->chunks_needed;
      const uint32_t cap = threadIdx.x == 0 ? P.cap_spans : threadIdx.x == 1 ? P.cap_tris : threadIdx.x == 2 ? P.cap_entries : P.cap_chunks;
      unsigned long long base = 0;
      if (tot) base = atomicAdd(ctr, (unsigned long long)tot);
      s_base[threadIdx.x] = base;
      s_fit[threadIdx.x] = base + tot <= cap ? 1u : 0u;
    }
    __syncthreads();
    const bool fits = s_fit[0] && s_fit[1] && s_fit[2] && s_fit[3];
    unsigned long long sb = s_base[0], tb = s_base[1], eb = s_base[2], cb_ = s_base[3];
    for (uint32_t w = 0; w < wid; w++) { sb += s_tot[w][0]; tb += s_tot[w][1]; eb += s_tot[w][2]; cb_ += s_tot[w][3]; }
    __syncthreads();  // s_tot / s_base are reused by the next iteration
    if (!fits) {
      if (threadIdx.x == 0) rf_overflow(P);
      continue;  // keep counting what is needed, write nothing
    }
    const uint32_t sbase = (uint32_t)sb + (incl_s - nsp);  // defined on every lane (nsp = 0 where nothing is emitted)
    const uint32_t tri_idx = (uint32_t)tb + __popc(emask & lt);
    uint32_t eidx = (uint32_t)eb + (incl_e - nent);
    uint32_t cidx = (uint32_t)cb_ + (incl_c - nchunk);
    const bool inline_walk = emit && !is_line && H0.n + H1.n <= RF_INLINE_ROWS;
    const bool chunked = emit && !inline_walk && !is_line;
    const uint32_t ch0 = chunked ? (H0.n + RF_CHUNK - 1) / RF_CHUNK : 0u, ch1 = chunked ? (H1.n + RF_CHUNK - 1) / RF_CHUNK : 0u;
    const uint32_t eck0 = cidx, eck1 = cidx + ch0;  // edge checkpoints are indexed by chunk position
    uint32_t* const stg = reinterpret_cast<uint32_t*>(s_stage[wid]);
    {  // triangle record: into the warp's staging buffer (record = rank among the emitting lanes), or straight to global memory
      uint32_t* tr = SS::ON ? stg + __popc(emask & lt) * SS::TWP : P.tris + (size_t)tri_idx * TW;
      if (emit) {
        *reinterpret_cast<uint4*>(tr) = make_uint4(key, d, sbase, (uint32_t)Yf0);
        *reinterpret_cast<uint4*>(tr + 4) = make_uint4(H0.n, H1.n | (tgt << 16), eck0, eck1);
#pragma unroll
        for (int hh = 0; hh < 2; hh++) {
          const HalfSetup<LT>& H = hh ? H1 : H0;
          uint32_t w[HS];
#pragma unroll
          for (int i = 0; i < HS; i++) w[i] = 0u;
#pragma unroll
          for (int i = 0; i < NV; i++) w[TR::O_DV + i] = __float_as_uint(H.dv[1 + i]);
#pragma unroll
          for (int i = 0; i < NL; i++) { w[TR::O_L + i] = __float_as_uint(H.L[i]); w[TR::O_DL + i] = __float_as_uint(H.dl[i]); }
          w[TR::O_R] = __float_as_uint(H.R); w[TR::O_DR] = __float_as_uint(H.dr); w[TR::O_Y] = __float_as_uint(H.y);
#pragma unroll
          for (int q = 0; q < HS / 4; q++) *reinterpret_cast<uint4*>(tr + 8 + hh * HS + 4 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        }
      }
      if (SS::ON) {  // the warp's records are consecutive in P.tris: 128-bit stores, consecutive lanes -> consecutive 16 bytes
        __syncwarp();
        const uint32_t nq = (uint32_t)__popc(emask) * (TW / 4);
        uint4* dst = reinterpret_cast<uint4*>(P.tris + (size_t)(uint32_t)tb * TW);
        for (uint32_t qi = lane; qi < nq; qi += 32) {
          const uint32_t rec = qi / (TW / 4), part = qi - rec * (TW / 4);
          dst[qi] = *reinterpret_cast<const uint4*>(stg + rec * SS::TWP + part * 4);
        }
        __syncwarp();  // the buffer is reused for the span records below
      }
    }
    if (emit) {
      if (is_line) {  // spans (one run per row, rows without pixels in band keep n = 0) and exact bin entries
        const TargetDesc& T = P.targets[tgt];
        constexpr int SW_ = Rec<LT>::SW;
        for (uint32_t r = 0; r < H0.n; r++) *reinterpret_cast<uint2*>(P.spans + (size_t)(sbase + r) * SW_) = make_uint2(0u, RF_NO_CKPT);
        line_walk<LT, 1>(P, T, LG, LM, lane, sbase, eidx, key, tri_idx, line_lanes, my_frags_i);
      }
      // bin entries: every tile of the conservative per-tile-row column range
      if (nent && !is_line) {
        const TargetDesc& T = P.targets[tgt];
        for (uint32_t tr = tr0; tr <= tr1; tr++) {
          uint32_t ca, cb;
          tile_row_cols<LT>(H0, H1, Yf0, H0.n + H1.n, tr, margin, tiles_x, ca, cb);
          const uint32_t tbase = T.tile_base + tr * tiles_x;
          for (uint32_t c = ca; c <= cb; c++) {
            P.entries[eidx++] = make_uint4(tbase + c, key, tri_idx, 0u);
            atomicAdd(P.tile_cnt + tbase + c, 1u);
          }
        }
      }
      // chunk record: {tri*2+half, chunk | rows << 16, first span index, draw | target << 16}
      for (uint32_t c = 0; c < ch0; c++)
        P.chunks[cidx++] = make_uint4(tri_idx * 2u, c | min(RF_CHUNK, H0.n - c * RF_CHUNK) << 16, sbase + c * RF_CHUNK, d | tgt << 16);
      for (uint32_t c = 0; c < ch1; c++)
        P.chunks[cidx++] = make_uint4(tri_idx * 2u + 1u, c | min(RF_CHUNK, H1.n - c * RF_CHUNK) << 16, sbase + H0.n + c * RF_CHUNK, d | tgt << 16);
      if (ch0 > 1) { const unsigned long long sl = agg_atomic_inc(&P.status->tall_needed, lane); if (sl < P.cap_tall) P.talllist[sl] = tri_idx * 2u; else rf_overflow(P); }
      if (ch1 > 1) { const unsigned long long sl = agg_atomic_inc(&P.status->tall_needed, lane); if (sl < P.cap_tall) P.talllist[sl] = tri_idx * 2u + 1u; else rf_overflow(P); }
    }
    // few rows: walk them here, serially (sequential adds down both edges). The span records go to the staging buffer at
    // the lane's offset among the warp's inline-walked rows (or straight to P.spans).
    const uint32_t nin = inline_walk ? nsp : 0u;
    const uint32_t incl_in = SS::ON ? warp_scan_incl(nin, lane) : 0u;
    if (inline_walk) {
      uint32_t sidx = SS::ON ? incl_in - nin : sbase, row = 0;
      uint32_t* const span_base = SS::ON ? stg : P.spans;
      bool oob = false;
      const uint32_t nU = H0.n;
      walk_half_inline<LT>(span_base, t_h, t_w, t_by0, t_by1, H0, sidx, row, long_rows, my_frags_i, oob);
      walk_half_inline<LT>(span_base, t_h, t_w, t_by0, t_by1, H1, sidx, row, long_rows, my_frags_i, oob);
      if (oob) atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB);
      long_sbase = sbase; long_tri = tri_idx; long_nU = nU;
    }
    if (SS::ON) {
      // Copy-out. The staged spans of consecutive lanes are consecutive in P.spans too, except across a lane whose spans are
      // written elsewhere (a chunked triangle: k_walk; a line: above): such lanes cut the warp into runs, each one contiguous
      // copy of 8-byte units (span records are 24 bytes, 8-byte aligned).
      constexpr int SW_ = Rec<LT>::SW;
      __syncwarp();
      const uint32_t total_in = __shfl_sync(0xFFFFFFFFu, incl_in, 31);
      uint32_t breakers = __ballot_sync(0xFFFFFFFFu, emit && !inline_walk && nsp != 0);
      uint32_t a = 0;
      while (total_in != 0 && a < 32) {
        const uint32_t b = breakers ? (uint32_t)__ffs(breakers) - 1u : 32u;  // the run is lanes [a, b)
        const uint32_t u0 = __shfl_sync(0xFFFFFFFFu, incl_in - nin, a);
        const uint32_t u1 = b < 32 ? __shfl_sync(0xFFFFFFFFu, incl_in - nin, b & 31u) : total_in;
        const uint32_t g0 = __shfl_sync(0xFFFFFFFFu, sbase, a);
        const uint2* src = reinterpret_cast<const uint2*>(stg + (size_t)u0 * SW_);
        uint2* dst = reinterpret_cast<uint2*>(P.spans + (size_t)g0 * SW_);
        const uint32_t nu = (u1 - u0) * (SW_ / 2);
        for (uint32_t i = lane; i < nu; i += 32) dst[i] = src[i];
        if (b >= 32) break;
        breakers &= breakers - 1;
        a = b + 1;
      }
      __syncwarp();  // the buffer is reused by the next iteration
    }
    // ---- long list of the inline walks: one warp-aggregated allocation (warp prefix sum) for all lanes
    {
      const uint32_t nl = __popc(long_rows);
      const uint32_t incl_l = warp_scan_incl(nl, lane);
      const uint32_t tot_l = __shfl_sync(0xFFFFFFFFu, incl_l, 31);
      if (tot_l) {
        unsigned long long lb = 0;
        if (lane == 0) lb = atomicAdd(&P.status->long_needed, (unsigned long long)tot_l);
        lb = __shfl_sync(0xFFFFFFFFu, lb, 0);
        if (lb + tot_l <= P.cap_long) {
          unsigned long long slot = lb + (incl_l - nl);
          uint32_t m = long_rows;
          while (m) {
            const uint32_t r = __ffs(m) - 1;
            m &= m - 1;
            P.longlist[slot++] = make_uint2(long_sbase + r, long_tri * 2u + (r >= long_nU ? 1u : 0u));
          }
        } else if (lane == 0) rf_overflow(P);
      }
    }
    // ---- per-draw frags.i of the inline walks
    {
      const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, d, 0);
      const bool uniform = __all_sync(0xFFFFFFFFu, !have || d == d0);
      if (uniform) {
        unsigned long long fi = my_frags_i;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) fi += __shfl_xor_sync(0xFFFFFFFFu, fi, o);
        if (lane == 0 && fi) atomicAdd(&P.dstats[d0].frags_i, fi);
      } else if (have && my_frags_i) {
        atomicAdd(&P.dstats[d].frags_i, my_frags_i);
      }
    }
  }
}

// =============================================================================================
// K2b k_edge_ckpt: for halves taller than one chunk, walk ONLY the running sums of both edges
// (the reference's sequential adds, raster.rs:88-104) and store the state at every chunk start,
// so that k_walk can process all chunks of a tall triangle in parallel. One thread per tall half.
// =============================================================================================
template <int LT>
__global__ void __launch_bounds__(128) k_edge_ckpt(PassParams P) {
  constexpr int NL = 2 + LT;
  constexpr int TW = Rec<LT>::TW, HS = Rec<LT>::HS, EW = Rec<LT>::EW;
  using TR = TriRec<LT>;
  if (rf_poisoned(P)) return;
  const uint32_t nt = (uint32_t)min(P.status->tall_needed, (unsigned long long)P.cap_tall);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nt; i += gridDim.x * blockDim.x) {
    const uint32_t own = P.talllist[i];
    const uint32_t* tr = P.tris + (size_t)(own >> 1) * TW;
    const uint32_t hh = own & 1u;
    const uint32_t n = hh ? (tr[5] & 0xFFFFu) : tr[4];
    const uint32_t ebase = tr[6 + hh];
    const uint32_t* hs = tr + 8 + hh * HS;
    float L[NL], dl[NL];
#pragma unroll
    for (int k = 0; k < NL; k++) { L[k] = __uint_as_float(hs[TR::O_L + k]); dl[k] = __uint_as_float(hs[TR::O_DL + k]); }
    float R = __uint_as_float(hs[TR::O_R]);
    const float dr = __uint_as_float(hs[TR::O_DR]);
    uint32_t c = 0;
    for (uint32_t j = 0; j + RF_CHUNK < n; j += RF_CHUNK) {
      for (uint32_t r = 0; r < RF_CHUNK; r++) {
#pragma unroll
        for (int k = 0; k < NL; k++) L[k] = L[k] + dl[k];
        R = R + dr;
      }
      uint32_t* e = P.ecks + (size_t)(ebase + c + 1) * EW;  // state at the start of chunk c+1
      uint32_t w[EW];
#pragma unroll
      for (int k = 0; k < EW; k++) w[k] = k < NL ? __float_as_uint(L[k]) : (k == NL ? __float_as_uint(R) : 0u);
#pragma unroll
      for (int q = 0; q < EW / 2; q++) *reinterpret_cast<uint2*>(e + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
      c++;
    }
  }
}

// =============================================================================================
// K2c k_walk: ScanlineIter::next (raster.rs:80-114) for every row of every drawn triangle, one ROW
// per lane. A warp takes 32 chunks (<= 32 rows each), expands them to rows with a warp prefix sum,
// and every lane brings its row's edge state up to date with r sequential adds from the chunk
// start (r < 32) — the same additions, in the same order, as the reference's running sums.
// Emits one span record per row; spans crossing a tile column go to the long list for k_ckpt.
// =============================================================================================
template <int LT>
__global__ void __launch_bounds__(128) k_walk(PassParams P) {
  constexpr int NL = 2 + LT;
  constexpr int TW = Rec<LT>::TW, HS = Rec<LT>::HS, EW = Rec<LT>::EW, SW = Rec<LT>::SW;
  using TR = TriRec<LT>;
  if (rf_poisoned(P)) return;
  const uint32_t lane = lane_id(), lt = (1u << lane) - 1u;
  const uint32_t nch = (uint32_t)min(P.status->chunks_needed, (unsigned long long)P.cap_chunks);
  const uint32_t wpb = blockDim.x >> 5;
  const uint32_t stride = gridDim.x * wpb * 32u;
  uint32_t cb = (blockIdx.x * wpb + (threadIdx.x >> 5)) * 32u;
  // software pipeline over chunk batches: the records of batch n+1 are requested while batch n is walked
  uint4 ch_next = make_uint4(0u, 0u, 0u, 0u);
  if (cb + lane < nch) ch_next = __ldg(P.chunks + cb + lane);
  // Long-list slots are reserved RF_LONG_BLOCK at a time per warp: one same-address atomic per block
  // instead of one per row batch. Unused slots of a block are left as {~0, ~0} for k_ckpt to skip.
  unsigned long long ll_next = 0, ll_end = 0;  // warp-uniform
  for (; cb < nch; cb += stride) {
    const uint4 ch = ch_next;
    const bool c_have = cb + lane < nch;
    ch_next = make_uint4(0u, 0u, 0u, 0u);
    if (cb + stride + lane < nch) ch_next = __ldg(P.chunks + cb + stride + lane);
    const uint32_t c_rows = c_have ? (ch.y >> 16) : 0u;
    const uint32_t incl = warp_scan_incl(c_rows, lane);
    const uint32_t n_items = __shfl_sync(0xFFFFFFFFu, incl, 31);
    unsigned long long my_frags_i = 0;
    uint32_t my_draw = 0xFFFFFFFFu;

    struct Pre {
      bool valid;
      uint32_t r, own, cidx, sidx, dt, cpos;
      uint32_t buf[HS];
      uint32_t eck[EW];
    };
    auto fetch = [&](uint32_t ib, Pre& p) {
      const uint32_t item = ib + lane;
      p.valid = item < n_items;
      uint32_t oc = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const uint32_t cand = oc + step;
        const uint32_t e = __shfl_sync(0xFFFFFFFFu, incl, (cand - 1) & 31);
        if (cand <= 32 && e <= item) oc = cand;
      }
      oc &= 31u;
      const uint32_t o_incl = __shfl_sync(0xFFFFFFFFu, incl, oc), o_rows = __shfl_sync(0xFFFFFFFFu, c_rows, oc);
      p.own = __shfl_sync(0xFFFFFFFFu, ch.x, oc);
      p.cidx = __shfl_sync(0xFFFFFFFFu, ch.y, oc) & 0xFFFFu;
      p.sidx = __shfl_sync(0xFFFFFFFFu, ch.z, oc);
      p.dt = __shfl_sync(0xFFFFFFFFu, ch.w, oc);
      p.cpos = cb + oc;
      p.r = p.valid ? item - (o_incl - o_rows) : 0u;  // row inside the chunk = number of adds
      if (p.valid) {
        const uint32_t* hs = P.tris + (size_t)(p.own >> 1) * TW + 8 + (p.own & 1u) * HS;
#pragma unroll
        for (int q = 0; q < HS / 4; q++) {
          const uint4 t = __ldg(reinterpret_cast<const uint4*>(hs) + q);
          p.buf[4 * q] = t.x; p.buf[4 * q + 1] = t.y; p.buf[4 * q + 2] = t.z; p.buf[4 * q + 3] = t.w;
        }
        if (p.cidx > 0) {  // chunk start state from the edge checkpoints
          const uint32_t* e = P.ecks + (size_t)p.cpos * EW;
#pragma unroll
          for (int q = 0; q < EW / 2; q++) {
            const uint2 t = __ldg(reinterpret_cast<const uint2*>(e) + q);
            p.eck[2 * q] = t.x; p.eck[2 * q + 1] = t.y;
          }
        }
      }
    };
    Pre cur, nxt;
    fetch(0, cur);
    for (uint32_t ib = 0; ib < n_items; ib += 32, cur = nxt) {
      nxt.valid = false;
      if (ib + 32 < n_items) fetch(ib + 32, nxt);
      const bool valid = cur.valid;
      const uint32_t r = cur.r;
      float L[NL], dl[NL], dv[NL], R = 0.0f, dr = 0.0f, y = 0.0f;
#pragma unroll
      for (int k = 0; k < NL; k++) { L[k] = 0.0f; dl[k] = 0.0f; dv[k] = 0.0f; }
      if (valid) {
#pragma unroll
        for (int k = 0; k < NL; k++) { L[k] = __uint_as_float(cur.buf[TR::O_L + k]); dl[k] = __uint_as_float(cur.buf[TR::O_DL + k]); }
#pragma unroll
        for (int k = 1; k < NL; k++) dv[k] = __uint_as_float(cur.buf[TR::O_DV + k - 1]);
        R = __uint_as_float(cur.buf[TR::O_R]); dr = __uint_as_float(cur.buf[TR::O_DR]);
        y = __uint_as_float(cur.buf[TR::O_Y]) + (float)(cur.cidx * RF_CHUNK + r);  // exact: row centres are k + 0.5 below 2^24
        if (cur.cidx > 0) {
#pragma unroll
          for (int k = 0; k < NL; k++) L[k] = __uint_as_float(cur.eck[k]);
          R = __uint_as_float(cur.eck[NL]);
        }
      }
      uint32_t maxr = r;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) maxr = max(maxr, __shfl_xor_sync(0xFFFFFFFFu, maxr, o));
      for (uint32_t j = 0; j < maxr; j++) {  // r sequential adds down both edges
        if (j < r) {
#pragma unroll
          for (int k = 0; k < NL; k++) L[k] = L[k] + dl[k];
          R = R + dr;
        }
      }
      bool is_long = false;
      uint32_t long_sidx = 0;
      if (valid) {
        const uint32_t o_draw = cur.dt & 0xFFFFu;
        const TargetDesc& T = P.targets[cur.dt >> 16];
        const float x0r = round_up_to_half(L[0]), x1r = round_up_to_half(R);
        const float tx = x0r - L[0];
        uint32_t w[SW];
#pragma unroll
        for (int k = 1; k < NL; k++) w[2 + (k - 1)] = __float_as_uint(L[k] + ((L[k] + dv[k]) - L[k]) * tx);
#pragma unroll
        for (int k = 2 + NL - 1; k < SW; k++) w[k] = 0u;
        const uint32_t cnt = sat_u32(x1r - x0r);
        const uint32_t Yf = sat_u32(y), X0 = sat_u32(x0r), X1 = max(sat_u32(x1r), X0);
        uint32_t nn = min(cnt, X1 - X0);
        if (Yf >= T.h || X1 > T.w) {  // target.rs:148,173-174 (slice index panics)
          atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB);
          nn = 0;
        } else if (Yf < T.band_y0 || Yf >= T.band_y1) {
          nn = 0;  // not this GPU's row band
        } else {
          if (my_draw != o_draw) {
            if (my_frags_i) atomicAdd(&P.dstats[my_draw].frags_i, my_frags_i);
            my_frags_i = 0; my_draw = o_draw;
          }
          my_frags_i += X1 - X0;
        }
        const uint32_t sidx = cur.sidx + r;
        is_long = nn && ((X0 + nn - 1) >> RF_TILE_SHIFT) != (X0 >> RF_TILE_SHIFT);  // crosses a tile column: k_ckpt adds checkpoints
        long_sidx = sidx;
        w[0] = X0 | nn << 16;
        w[1] = RF_NO_CKPT;
        uint32_t* sr = P.spans + (size_t)sidx * SW;
#pragma unroll
        for (int q = 0; q < SW / 2; q++) *reinterpret_cast<uint2*>(sr + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
      }
      {  // long-list append from warp-private reserved blocks
        const uint32_t lm = __ballot_sync(0xFFFFFFFFu, is_long);
        if (lm) {
          const uint32_t nl = __popc(lm);
          if (ll_next + nl > ll_end) {  // reserve a fresh block (the tail of the old one stays marked unused)
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(&P.status->long_needed, (unsigned long long)RF_LONG_BLOCK);
            base = __shfl_sync(0xFFFFFFFFu, base, 0);
            if (base + RF_LONG_BLOCK <= P.cap_long) {
              for (uint32_t q = lane; q < RF_LONG_BLOCK; q += 32) P.longlist[base + q] = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
              __syncwarp();
              ll_next = base; ll_end = base + RF_LONG_BLOCK;
            } else {
              if (lane == 0) rf_overflow(P);
              ll_next = ll_end = 0;
            }
          }
          if (ll_next + nl <= ll_end) {
            if (is_long) P.longlist[ll_next + __popc(lm & lt)] = make_uint2(long_sidx, cur.own);
            ll_next += nl;
          }
        }
      }
    }
    // ---- frags.i: one atomic per warp when the whole warp worked on one draw
    {
      const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, my_draw, 0);
      const bool uniform = __all_sync(0xFFFFFFFFu, my_draw == d0 || my_frags_i == 0);
      if (uniform) {
        unsigned long long fi = my_frags_i;
        uint32_t dd = my_frags_i ? my_draw : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { fi += __shfl_xor_sync(0xFFFFFFFFu, fi, o); dd = max(dd, __shfl_xor_sync(0xFFFFFFFFu, dd, o)); }
        if (lane == 0 && fi) atomicAdd(&P.dstats[dd].frags_i, fi);
      } else if (my_frags_i) {
        atomicAdd(&P.dstats[my_draw].frags_i, my_frags_i);
      }
    }
  }
}
