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
template <int LT>
__global__ void __launch_bounds__(256) k_vertex(PassParams P) {
  constexpr int CVS = Rec<LT>::CVS;
  if (P.cstatus->poison) return;
  for (uint32_t gv = blockIdx.x * blockDim.x + threadIdx.x; gv < P.NV; gv += gridDim.x * blockDim.x) {
    const uint32_t d = find_draw(P.vbase, P.n_draws, gv);
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
  for (int i = 0; i < LT; i++) s.a[i] = ((persp_mask >> i) & 1u) ? c.a[i] / w : c.a[i];
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

// Walk one trapezoid half row by row (ScanlineIter::next, raster.rs:80-114): sequential adds down
// both edges, one span record per row. Tracks the tile-column range touched in the current tile
// row and flushes exact bin entries whenever the walk leaves a tile row.
template <int LT> struct WalkState {
  uint32_t sidx;        // next span record
  uint32_t Y;           // current row
  uint32_t cmin, cmax;  // tile columns touched in the current tile row (cmin > cmax: none)
  unsigned long long frags_i;
};

template <int LT>
__device__ __forceinline__ void flush_tile_row(const PassParams& P, const TargetDesc& T, WalkState<LT>& W, uint32_t trow, uint32_t tr0,
                                               uint32_t c_lo, uint32_t ncols, uint32_t ebase, uint32_t key, uint32_t tri_idx) {
  if (W.cmin <= W.cmax && (W.cmin < c_lo || W.cmax >= c_lo + ncols)) atomicOr(&P.status->error, RF_ERRBIT_INTERNAL);
  uint4* e = P.entries + ebase + (size_t)(trow - tr0) * ncols;
  const uint32_t tbase = T.tile_base + trow * T.tiles_x;
  for (uint32_t c = 0; c < ncols; c++) {
    const uint32_t col = c_lo + c;
    const bool valid = col >= W.cmin && col <= W.cmax;
    e[c] = make_uint4(valid ? tbase + col : RF_NO_TILE, key, tri_idx, 0u);
    if (valid) atomicAdd(P.tile_cnt + tbase + col, 1u);
  }
  W.cmin = 0xFFFFFFFFu;
  W.cmax = 0u;
}

template <int LT>
__device__ __forceinline__ void walk_half(const PassParams& P, const TargetDesc& T, HalfSetup<LT>& H, WalkState<LT>& W, uint32_t rows_left_after,
                                          uint32_t own, uint32_t tr0, uint32_t c_lo, uint32_t ncols, uint32_t ebase, uint32_t key, uint32_t tri_idx) {
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
    if (Yf >= T.h || X1 > T.w) {  // target.rs:148,173-174 (slice index panics)
      atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB);
      nn = 0;
    } else if (Yf < T.band_y0 || Yf >= T.band_y1) {
      nn = 0;  // not this GPU's row band
    } else {
      W.frags_i += X1 - X0;
    }
    if (nn) {
      const uint32_t ca = X0 >> RF_TILE_SHIFT, cb = (X0 + nn - 1) >> RF_TILE_SHIFT;
      W.cmin = min(W.cmin, ca);
      W.cmax = max(W.cmax, cb);
      if (ca != cb) {  // crosses a tile-column boundary: k_ckpt will add checkpoints
        const unsigned long long slot = agg_atomic_inc(&P.status->long_needed);
        if (slot < P.cap_long) P.longlist[slot] = make_uint2(W.sidx, own);
        else { P.status->overflow = 1; P.cstatus->poison = 1; }
      }
    }
    w[0] = X0 | nn << 16;
    w[1] = RF_NO_CKPT;
    uint32_t* sr = P.spans + (size_t)W.sidx * SW;
#pragma unroll
    for (int q = 0; q < SW / 2; q++) *reinterpret_cast<uint2*>(sr + 2 * q) = make_uint2(w[2 * q], w[2 * q + 1]);
    W.sidx++;
    const bool last = (j + 1 == H.n) && rows_left_after == 0;
    if (last || ((W.Y + 1) >> RF_TILE_SHIFT) != (W.Y >> RF_TILE_SHIFT))
      flush_tile_row<LT>(P, T, W, W.Y >> RF_TILE_SHIFT, tr0, c_lo, ncols, ebase, key, tri_idx);
    W.Y++;
    y = y + 1.0f;
  }
}

template <int LT>
__global__ void __launch_bounds__(128) k_prim(PassParams P) {
  constexpr int NL = 2 + LT;
  constexpr int TW = Rec<LT>::TW;
  if (P.cstatus->poison) return;
  const uint32_t lane = lane_id();
  const uint32_t n_iter = (P.NP + blockDim.x * gridDim.x - 1) / (blockDim.x * gridDim.x);
  for (uint32_t it = 0; it < n_iter; it++) {
    const uint32_t gp = (it * gridDim.x + blockIdx.x) * blockDim.x + threadIdx.x;
    const bool have = gp < P.NP;
    uint32_t d = 0, ntri = 0;
    CVert<LT> c0, c1, c2;
    CVert<LT> poly[10], tmp[10];
    bool clipped = false;
    if (have) {
      d = find_draw(P.pbase, P.n_draws, gp);
      const DrawDesc& D = P.draws[d];
      const uint32_t* ip = D.indices + 3 * (size_t)(gp - __ldg(P.pbase + d));
      const uint32_t i0 = __ldg(ip), i1 = __ldg(ip + 1), i2 = __ldg(ip + 2);
      if (i0 >= D.n_verts || i1 >= D.n_verts || i2 >= D.n_verts) {
        atomicOr(&P.status->error, RF_ERRBIT_INDEX_OOB);  // prim.rs:17-19 panics
      } else {
        const uint32_t vb = __ldg(P.vbase + d);
        load_cv<LT>(P.cv, vb + i0, c0);
        load_cv<LT>(P.cv, vb + i1, c1);
        load_cv<LT>(P.cv, vb + i2, c2);
        const uint32_t all = c0.oc & c1.oc & c2.oc, any = c0.oc | c1.oc | c2.oc;
        if (all != 0) ntri = 0;          // Status::Hidden, clip.rs:245-267
        else if (any == 0) ntri = 1;     // Status::Visible
        else {
          clipped = true;
          poly[0] = c0; poly[1] = c1; poly[2] = c2;
          const int n = clip_polygon<LT>(poly, tmp);
          ntri = n >= 3 ? (uint32_t)(n - 2) : 0u;  // fan (p0, pk, pk+1), clip.rs:375-394
        }
      }
    }
    uint32_t max_tri = ntri;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_tri = max(max_tri, __shfl_xor_sync(0xFFFFFFFFu, max_tri, o));

    unsigned long long my_frags_i = 0;
    uint32_t my_prims_o = 0;

    for (uint32_t t = 0; t < max_tri; t++) {
      bool emit = false;
      HalfSetup<LT> H0, H1;
      H0.n = H1.n = 0;
      uint32_t tgt = 0, Y0 = 0, tr0 = 0, c_lo = 0, ncols = 0, nent = 0;
      if (t < ntri) {
        const DrawDesc& D = P.draws[d];
        SVert<LT> s[3];
        if (clipped) {
          to_screen<LT>(poly[0], D.vp, D.persp_mask, s[0]);
          to_screen<LT>(poly[t + 1], D.vp, D.persp_mask, s[1]);
          to_screen<LT>(poly[t + 2], D.vp, D.persp_mask, s[2]);
        } else {
          to_screen<LT>(c0, D.vp, D.persp_mask, s[0]);
          to_screen<LT>(c1, D.vp, D.persp_mask, s[1]);
          to_screen<LT>(c2, D.vp, D.persp_mask, s[2]);
        }
        // Tri::winding geom/prim.rs:288-294 ; Context::face_cull ctx.rs:95-101
        const float abx = s[1].x - s[0].x, aby = s[1].y - s[0].y;
        const float acx = s[2].x - s[0].x, acy = s[2].y - s[0].y;
        float wz = 0.0f;
        wz = wz + (-aby) * acx;
        wz = wz + abx * acy;
        const bool back = wz < 0.0f;
        const uint32_t cull = D.flags & RF_F_CULL_MASK;
        if (!((cull == RF_CULL_BACK && back) || (cull == RF_CULL_FRONT && !back))) {
          my_prims_o++;  // render.rs:195-196: counted before rasterisation, whatever it covers
          // tri_fill raster.rs:185-224: stable sort by y (total_cmp)
          int o0 = 0, o1 = 1, o2 = 2;
          {
            const int32_t k0 = total_key(s[0].y), k1 = total_key(s[1].y), k2 = total_key(s[2].y);
            int32_t ka = k0, kb = k1, kc = k2;
            if (kb < ka) { int ti = o0; o0 = o1; o1 = ti; int32_t tk = ka; ka = kb; kb = tk; }
            if (kc < kb) { int ti = o1; o1 = o2; o2 = ti; int32_t tk = kb; kb = kc; kc = tk; }
            if (kb < ka) { int ti = o0; o0 = o1; o1 = ti; }
          }
          float top[NL], mid0[NL], bot[NL], mid1[NL];
          float ty, my, by;
#define RF_PICK(dst, yy, idx)                                            \
  {                                                                      \
    const SVert<LT>& q = (idx == 0) ? s[0] : ((idx == 1) ? s[1] : s[2]); \
    dst[0] = q.x; dst[1] = q.z; yy = q.y;                                \
    _Pragma("unroll") for (int i = 0; i < LT; i++) dst[2 + i] = q.a[i];  \
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
          const TargetDesc& T = P.targets[D.target];
          tgt = D.target;
          // Row-range sanity. A scanline at y >= h panics in the reference (target.rs:148,173);
          // RF_MAX_ROWS bounds the loop against absurd coordinates; a negative first row can only
          // come from a viewport outside the target and is rejected (the reference would draw it at row 0).
          uint32_t nrows = H0.n + H1.n;
          if (nrows != 0) {
            const float yfirst = H0.n ? H0.y : H1.y;
            const float ylast = yfirst + (float)(nrows - 1);
            if (H0.n > RF_MAX_ROWS || H1.n > RF_MAX_ROWS || sat_u32(ylast) >= T.h) {
              atomicOr(&P.status->error, RF_ERRBIT_TARGET_OOB);
              nrows = 0;
            } else if (yfirst < 0.0f) {
              atomicOr(&P.status->error, RF_ERRBIT_NEG_ROW);
              nrows = 0;
            }
            if (nrows == 0) H0.n = H1.n = 0;
            else {
              emit = true;
              Y0 = sat_u32(yfirst);
              tr0 = Y0 >> RF_TILE_SHIFT;
              const uint32_t tr1 = (Y0 + nrows - 1) >> RF_TILE_SHIFT;
              const float xmin = fminf(s[0].x, fminf(s[1].x, s[2].x)), xmax = fmaxf(s[0].x, fmaxf(s[1].x, s[2].x));
              c_lo = min(sat_u32(floorf(xmin) - 1.0f) >> RF_TILE_SHIFT, T.tiles_x - 1);
              const uint32_t c_hi = min(sat_u32(floorf(xmax) + 1.0f) >> RF_TILE_SHIFT, T.tiles_x - 1);
              ncols = (c_hi >= c_lo ? c_hi - c_lo : 0u) + 1u;
              nent = (tr1 - tr0 + 1) * ncols;
            }
          }
        }
      }
      // ---- warp-aggregated allocation (warp prefix sums): span records, triangle records, bin slots
      const uint32_t nsp = emit ? (H0.n + H1.n) : 0u;
      const uint32_t incl_s = warp_scan_incl(nsp);
      const uint32_t incl_e = warp_scan_incl(nent);
      const uint32_t tot_s = __shfl_sync(0xFFFFFFFFu, incl_s, 31), tot_e = __shfl_sync(0xFFFFFFFFu, incl_e, 31);
      const uint32_t emask = __ballot_sync(0xFFFFFFFFu, emit);
      unsigned long long sb = 0, tb = 0, eb = 0;
      if (lane == 0 && emask) {
        sb = atomicAdd(&P.status->spans_needed, (unsigned long long)tot_s);
        tb = atomicAdd(&P.status->tris_needed, (unsigned long long)__popc(emask));
        eb = atomicAdd(&P.status->entries_needed, (unsigned long long)tot_e);
      }
      sb = __shfl_sync(0xFFFFFFFFu, sb, 0);
      tb = __shfl_sync(0xFFFFFFFFu, tb, 0);
      eb = __shfl_sync(0xFFFFFFFFu, eb, 0);
      const bool fits = sb + tot_s <= P.cap_spans && tb + __popc(emask) <= P.cap_tris && eb + tot_e <= P.cap_entries;
      if (!fits) {
        if (lane == 0 && emask) { P.status->overflow = 1; P.cstatus->poison = 1; }
        continue;  // keep counting what is needed, write nothing
      }
      if (!emit) continue;
      const uint32_t sbase = (uint32_t)sb + (incl_s - nsp);
      const uint32_t tri_idx = (uint32_t)tb + __popc(emask & lanemask_lt());
      const uint32_t ebase = (uint32_t)eb + (incl_e - nent);
      const uint32_t key = gp * 8u + t;
      {  // triangle record
        uint32_t w[TW];
        w[0] = key; w[1] = d; w[2] = sbase; w[3] = Y0; w[4] = H0.n; w[5] = H1.n | (tgt << 16);
#pragma unroll
        for (int i = 0; i < 1 + LT; i++) { w[6 + i] = __float_as_uint(H0.dv[1 + i]); w[6 + (1 + LT) + i] = __float_as_uint(H1.dv[1 + i]); }
#pragma unroll
        for (int i = 6 + 2 * (1 + LT); i < TW; i++) w[i] = 0u;
        uint32_t* tr = P.tris + (size_t)tri_idx * TW;
#pragma unroll
        for (int q = 0; q < TW / 4; q++) *reinterpret_cast<uint4*>(tr + 4 * q) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      }
      const TargetDesc& T = P.targets[tgt];
      WalkState<LT> W;
      W.sidx = sbase; W.Y = Y0; W.cmin = 0xFFFFFFFFu; W.cmax = 0u; W.frags_i = 0;
      walk_half<LT>(P, T, H0, W, H1.n, tri_idx * 2u, tr0, c_lo, ncols, ebase, key, tri_idx);
      walk_half<LT>(P, T, H1, W, 0u, tri_idx * 2u + 1u, tr0, c_lo, ncols, ebase, key, tri_idx);
      my_frags_i += W.frags_i;
    }
    // ---- per-draw stats: aggregate over the warp when every lane has the same draw
    {
      const uint32_t d0 = __shfl_sync(0xFFFFFFFFu, d, 0);
      const bool uniform = __all_sync(0xFFFFFFFFu, !have || d == d0);
      if (uniform) {
        unsigned long long fi = my_frags_i;
        uint32_t po = my_prims_o;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          fi += __shfl_xor_sync(0xFFFFFFFFu, fi, o);
          po += __shfl_xor_sync(0xFFFFFFFFu, po, o);
        }
        if (lane == 0) {
          if (po) atomicAdd(&P.dstats[d0].prims_o, (unsigned long long)po);
          if (fi) atomicAdd(&P.dstats[d0].frags_i, fi);
        }
      } else if (have) {
        if (my_prims_o) atomicAdd(&P.dstats[d].prims_o, (unsigned long long)my_prims_o);
        if (my_frags_i) atomicAdd(&P.dstats[d].frags_i, my_frags_i);
      }
    }
  }
}
