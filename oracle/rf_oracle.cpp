// rf_oracle.cpp — CPU ORACLE for retrofire's render() hot path.  TEST INFRASTRUCTURE ONLY.
//
// This file is a single-threaded C++17 restatement of the reference algorithm
// (jdahlstrom/retrofire v0.4.0, pure Rust). It exists to CHECK the CUDA path; nothing in
// retrofire_b200/ may link, import or call it. Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs use it.
//
// Why a restatement and not the reference itself: the reference is Rust and this image has
// no cargo/rustc (probed), so oracle/_ref cannot be built. Parity is pinned instead by the
// reference's own golden artefacts, which tests/test_oracle_golden.py checks bit-exactly:
//   core/tests/textured_quad.ppm (core/tests/rendering.rs:18-60), core/triangle.ppm
//   (core/examples/hello_tri.rs:47-53), render/raster.rs:326-436 KATs, the 5^9 clip histogram
//   (render/clip.rs:667-719) and the sampler KATs (render/tex.rs:381-418).
// Depth test, cull, discard and Stats have no reference test: "pinned by restatement only".
//
// Every + - * / below is one IEEE binary32 operation; build with -ffp-contract=off, no
// -ffast-math (Rust never fuses a*b+c). All citations are core/src/... in the reference.

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/retrofire_b200.h"  // struct layouts and enums only (no product code)

namespace {

constexpr int MAXL = RF_MAX_ATTR_LANES;

// ---- Rust `as` casts: saturating, NaN -> 0 ---------------------------------------------
inline uint32_t sat_u32(float f) {
  if (!(f > 0.0f)) return 0;  // NaN, negatives, zero
  if (f >= 4294967296.0f) return 0xFFFFFFFFu;
  return (uint32_t)f;
}
inline uint64_t sat_usize(float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 18446744073709551616.0f) return ~0ull;
  return (uint64_t)f;
}
inline int32_t sat_i32(float f) {
  if (f != f) return 0;
  if (f >= 2147483648.0f) return INT32_MAX;
  if (f <= -2147483648.0f) return INT32_MIN;
  return (int32_t)f;
}
inline uint8_t sat_u8(float f) {
  if (!(f > 0.0f)) return 0;
  if (f >= 255.0f) return 255;
  return (uint8_t)f;
}

// f32::max (used by the demo shaders, solids.rs:75, crates.rs:44): "if one of the arguments is NaN, the other is returned" —
// std::max(NaN, x) would return the NaN. NaN reaches the shaders through dv_dx = 0 * (1 / 0) on zero-width first rows.
inline float rust_max(float a, float b) { return a != a ? b : (b != b ? a : (a < b ? b : a)); }

// math/vec.rs:231-238 — dot() folds from Sc::zero(): (((0 + a0*b0) + a1*b1) + ...)
inline float dot4(const float* a, const float* b) {
  float acc = 0.0f;
  for (int i = 0; i < 4; i++) acc = acc + a[i] * b[i];
  return acc;
}
inline float dot3(const float* a, const float* b) {
  float acc = 0.0f;
  for (int i = 0; i < 3; i++) acc = acc + a[i] * b[i];
  return acc;
}
inline float dot2(const float* a, const float* b) {
  float acc = 0.0f;
  for (int i = 0; i < 2; i++) acc = acc + a[i] * b[i];
  return acc;
}

// ---- clip space -----------------------------------------------------------------------
struct ClipVert {  // render/clip.rs:56-61
  float pos[4];
  float attr[MAXL];
  uint8_t oc;
};

// render/clip.rs:215-222 — ClipPlane::new(x,y,z,off,bit) stores [x,y,z,-off]
const float PLANES[6][4] = {
    {0.0f, 0.0f, -1.0f, -1.0f},  // Near   0x01
    {0.0f, 0.0f, 1.0f, -1.0f},   // Far    0x02
    {-1.0f, 0.0f, 0.0f, -1.0f},  // Left   0x04
    {1.0f, 0.0f, 0.0f, -1.0f},   // Right  0x08
    {0.0f, -1.0f, 0.0f, -1.0f},  // Bottom 0x10
    {0.0f, 1.0f, 0.0f, -1.0f},   // Top    0x20
};

// render/clip.rs:240-242,108-111
inline uint8_t outcode(const float* pos) {
  uint8_t oc = 0;
  for (int p = 0; p < 6; p++)
    if (dot4(PLANES[p], pos) > 0.0f) oc |= (uint8_t)(1u << p);
  return oc;
}

// math.rs:196-198  self.add(&other.sub(self).mul(t))
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }

// render/clip.rs:121-145
inline bool intersect(int plane, const ClipVert& v0, const ClipVert& v1, int L, ClipVert& out) {
  float d0 = dot4(PLANES[plane], v0.pos);
  float d1 = dot4(PLANES[plane], v1.pos);
  if (!(d0 * d1 < 0.0f)) return false;
  float t = -d0 / (d1 - d0);
  for (int i = 0; i < 4; i++) out.pos[i] = lerp(v0.pos[i], v1.pos[i], t);
  for (int i = 0; i < L; i++) out.attr[i] = lerp(v0.attr[i], v1.attr[i], t);
  for (int i = L; i < MAXL; i++) out.attr[i] = 0.0f;
  out.oc = outcode(out.pos);  // ClipVert::new recomputes, clip.rs:303-309
  return true;
}

// render/clip.rs:167-193
void clip_plane(int plane, const std::vector<ClipVert>& in, std::vector<ClipVert>& out, int L) {
  const size_t n = in.size();
  const uint8_t bit = (uint8_t)(1u << plane);
  for (size_t k = 0; k < n; k++) {
    const ClipVert& v0 = in[k];
    const ClipVert& v1 = in[(k + 1) % n];
    if ((v0.oc & bit) == 0) out.push_back(v0);
    ClipVert x;
    if (intersect(plane, v0, v1, L, x)) out.push_back(x);
  }
}

// render/clip.rs:283-301 + 350-400: returns polygon (possibly empty) in `res`
void clip_tri(const ClipVert t[3], int L, std::vector<ClipVert>& vin, std::vector<ClipVert>& vout) {
  vin.assign(t, t + 3);
  vout.clear();
  for (int p = 0; p < 6; p++) {
    clip_plane(p, vin, vout, L);
    vin.clear();
    if (vout.empty()) break;
    if (p < 5) std::swap(vin, vout);
  }
}

// ---- vertex shaders (catalogue) -------------------------------------------
// mat.rs:968-972  ProjMat3::apply(Point3): 4 dots with [p,1]
inline void apply_proj(const float* m, const float* p3, float* out4) {
  float h[4] = {p3[0], p3[1], p3[2], 1.0f};
  for (int r = 0; r < 4; r++) out4[r] = dot4(m + 4 * r, h);
}

void shade_vertex(const rf_draw& d, const float* vin, ClipVert& cv) {
  const int L = (int)d.n_attr_lanes;
  const float* a = vin + 3;
  for (int i = 0; i < MAXL; i++) cv.attr[i] = 0.0f;
  switch (d.vs) {
    case RF_VS_MVP:  // crates.rs:39-41, tests/rendering.rs:27-29
      apply_proj(d.vs_uniform, vin, cv.pos);
      for (int i = 0; i < L; i++) cv.attr[i] = a[i];
      break;
    case RF_VS_MVP_LINEARIZE:  // hello_tri.rs:13-17; color.rs:277-285 (GAMMA = 2.2)
      apply_proj(d.vs_uniform, vin, cv.pos);
      for (int i = 0; i < L; i++) cv.attr[i] = powf(a[i], 2.2f);
      break;
    case RF_VS_SOLIDS: {  // solids.rs:70-79
      const float* mvp = d.vs_uniform;
      const float* spin = d.vs_uniform + 16;
      // Mat4::apply(Vec3): homogeneous w = 0, rows 0..2   (mat.rs:922-926)
      float nh[4] = {a[0], a[1], a[2], 0.0f};
      float nz = dot4(spin + 8, nh);
      float diffuse = rust_max(nz + 0.2f, 0.2f) * 0.8f;
      // 0.45 * (n + splat(1.1)) -> (n + 1.1) * 0.45 ; diffuse * rgb -> c * diffuse
      for (int i = 0; i < 3; i++) cv.attr[i] = ((a[i] + 1.1f) * 0.45f) * diffuse;
      apply_proj(mvp, vin, cv.pos);
      break;
    }
    case RF_VS_SPRITE: {  // sprites.rs:40-45
      const float* mv = d.vs_uniform;
      const float* proj = d.vs_uniform + 16;
      float vp[3] = {a[0] * 0.008f, a[1] * 0.008f, 0.0f * 0.008f};
      float h[4] = {vin[0], vin[1], vin[2], 1.0f};
      float view[3];
      for (int r = 0; r < 3; r++) view[r] = dot4(mv + 4 * r, h) + vp[r];  // mat.rs:945-949
      apply_proj(proj, view, cv.pos);
      cv.attr[0] = a[0];
      cv.attr[1] = a[1];
      break;
    }
  }
  cv.oc = outcode(cv.pos);
}

// ---- textures (render/tex.rs) ---------------------------------------------------------------
struct Tex {
  uint32_t w, h, fmt;
  const uint8_t* data;
  size_t stride;  // elements
};

inline void texel(const Tex& t, uint32_t u, uint32_t v, uint8_t rgba[4]) {
  const size_t bpp = t.fmt == RF_TEXEL_RGB888 ? 3 : 4;
  const uint8_t* p = t.data + ((size_t)v * t.stride + u) * bpp;
  rgba[0] = p[0];
  rgba[1] = p[1];
  rgba[2] = p[2];
  rgba[3] = bpp == 4 ? p[3] : 0xFF;  // Color3::to_rgba, color.rs:206-213
}
inline float rust_clamp(float x, float lo, float hi) {  // f32::clamp: NaN stays NaN
  if (x < lo) return lo;
  if (x > hi) return hi;
  return x;
}
// tex.rs:272-304
inline void sample_clamp(const Tex& t, float tu, float tv, uint8_t rgba[4]) {
  float w = (float)t.w, h = (float)t.h;
  float su = tu * w, sv = tv * h;
  uint32_t u = sat_u32(floorf(rust_clamp(su, 0.0f, w - 1.0f)));
  uint32_t v = sat_u32(floorf(rust_clamp(sv, 0.0f, h - 1.0f)));
  texel(t, u, v, rgba);
}
// tex.rs:313-357 SamplerOnce: `tc.u() as u32` of the scaled coordinate, no wrapping or clamping; `d[[u, v]]` panics when the
// texel is outside the texture (util/buf.rs Index) -> false
inline bool sample_once(const Tex& t, float tu, float tv, uint8_t rgba[4]) {
  float w = (float)t.w, h = (float)t.h;
  uint32_t u = sat_u32(w * tu), v = sat_u32(h * tv);
  if (u >= t.w || v >= t.h) return false;
  texel(t, u, v, rgba);
  return true;
}
// tex.rs:218-267
inline void sample_repeat_pot(const Tex& t, float tu, float tv, uint8_t rgba[4]) {
  float w = (float)t.w, h = (float)t.h;
  float su = w * tu, sv = h * tv;
  uint32_t u = (uint32_t)sat_i32(floorf(su)) & (t.w - 1);
  uint32_t v = (uint32_t)sat_i32(floorf(sv)) & (t.h - 1);
  texel(t, u, v, rgba);
}

// ---- fragment shaders (catalogue) ----------------------------------------------------------
// returns false = discard (shader.rs:48-55). var[] is already perspective-corrected.
bool shade_fragment(const rf_draw& d, const Tex* tex, const float* var, uint8_t rgba[4], bool* panicked) {
  switch (d.fs) {
    case RF_FS_TEX_ONCE:
      if (!sample_once(*tex, var[0], var[1], rgba)) { *panicked = true; return false; }
      return true;
    case RF_FS_COLOR3F:  // color.rs:246-263  to_color4: (256*c) as u8, a = 0xFF
      for (int i = 0; i < 3; i++) rgba[i] = sat_u8(256.0f * var[i]);
      rgba[3] = 0xFF;
      return true;
    case RF_FS_COLOR3F_SRGB:  // hello_tri.rs:18; color.rs:383-391 INV_GAMMA = 1/2.2
      for (int i = 0; i < 3; i++) rgba[i] = sat_u8(256.0f * powf(var[i], 1.0f / 2.2f));
      rgba[3] = 0xFF;
      return true;
    case RF_FS_COLOR4F:  // color.rs:347-360
      for (int i = 0; i < 4; i++) rgba[i] = sat_u8(256.0f * var[i]);
      return true;
    case RF_FS_CHECKER: {  // crates.rs:33-36
      bool eo = (var[0] > 0.5f) ^ (var[1] > 0.5f);
      uint8_t g = sat_u8(256.0f * (eo ? 0.8f : 0.1f));
      rgba[0] = rgba[1] = rgba[2] = g;
      rgba[3] = 0xFF;
      return true;
    }
    case RF_FS_TEX_CLAMP_LIT: {  // crates.rs:42-47
      float ndl = rust_max(dot3(var, d.fs_uniform), 0.0f);
      float kd = 0.4f + (1.0f - 0.4f) * ndl;  // lerp(t, 0.4, 1.0), math.rs:30-32
      uint8_t c[4];
      sample_clamp(*tex, var[3], var[4], c);
      // Color3::to_color3f: c as f32 / 256.0 ; * kd ; to_color4
      for (int i = 0; i < 3; i++) rgba[i] = sat_u8(256.0f * (((float)c[i] / 256.0f) * kd));
      rgba[3] = 0xFF;
      return true;
    }
    case RF_FS_TEX_CLAMP:  // tests/rendering.rs:30
      sample_clamp(*tex, var[0], var[1], rgba);
      return true;
    case RF_FS_TEX_REPEAT_POT:  // benches/fill.rs:74-91
      sample_repeat_pot(*tex, var[0], var[1], rgba);
      return true;
    case RF_FS_SPRITE_DISC: {  // sprites.rs:46-52
      float d2 = dot2(var, var);
      if (!(d2 < 1.0f)) return false;
      const float k[3] = {0.25f, 0.5f, 1.0f};
      // gray(1.0) - d2*rgb(..): Sub = self + (zero - rhs)   color.rs:701-709
      for (int i = 0; i < 3; i++) rgba[i] = sat_u8(256.0f * (1.0f + (0.0f - k[i] * d2)));
      rgba[3] = 0xFF;
      return true;
    }
    case RF_FS_NORMAL_VIS:  // curses.rs:53-56:  var / 2.0 + splat(0.5) ; Div = mul by recip
      for (int i = 0; i < 3; i++) rgba[i] = sat_u8(256.0f * (var[i] * 0.5f + 0.5f));
      rgba[3] = 0xFF;
      return true;
  }
  return false;
}

// util/pixfmt.rs:45-142 — Color4 -> uint32 device container (see include/retrofire_b200.h)
inline uint32_t pack_pixel(uint32_t fmt, const uint8_t c[4]) {
  const uint32_t r = c[0], g = c[1], b = c[2], a = c[3];
  switch (fmt) {
    case RF_FMT_RGBA8888: return r | g << 8 | b << 16 | a << 24;
    case RF_FMT_XRGB8888: return r << 16 | g << 8 | b;
    case RF_FMT_ARGB8888: return a | r << 8 | g << 16 | b << 24;
    case RF_FMT_BGRA8888: return b | g << 8 | r << 16 | a << 24;
    case RF_FMT_RGB888: return r | g << 8 | b << 16;
    case RF_FMT_RGB565: return ((r >> 3) & 0x1F) << 11 | ((g >> 2) & 0x3F) << 5 | ((b >> 3) & 0x1F);
    case RF_FMT_RGBA4444: return (r >> 4) << 12 | (g >> 4) << 8 | (b >> 4) << 4 | (a >> 4);
  }
  return 0;
}

// ---- rasterisation -----------------------------------------------------------------------
struct Lanes {  // Varyings<V> = (ScreenPt, V) flattened: [x, y, z, a0..]
  float v[3 + MAXL];
};

inline float round_up_to_half(float x) { return floorf(x + 0.5f) + 0.5f; }  // raster.rs:304-307

struct Target {
  uint32_t w, h, fmt;
  uint32_t* color;  // w*h containers
  float* depth;     // or nullptr (Buf2<Color4> etc: rasterize() without depth, target.rs:138-161)
  uint32_t band_y0, band_y1;
};

struct Raster {
  const rf_draw& d;
  const Tex* tex;
  Target& tg;
  rf_stats& st;
  int NL;  // 3 + L
  bool oob = false;
  bool tex_oob = false;  // SamplerOnce indexed outside the texture (the reference panics)

  // render/target.rs:138-198 on one Scanline (raster.rs:80-114 produced it)
  void scanline(uint64_t Y, uint64_t X0, uint64_t X1e, uint32_t cnt, Lanes v, const Lanes& dvdx) {
    const int L = NL - 3;
    uint64_t X1 = std::max(X1e, X0);
    if (Y >= tg.h || X1 > tg.w) {  // slice index panics: target.rs:148,173-174
      oob = true;
      return;
    }
    if (Y < tg.band_y0 || Y >= tg.band_y1) return;  // sort-first row band (not in the reference)
    st.frags_i += X1 - X0;
    uint64_t n = std::min<uint64_t>(cnt, X1 - X0);  // zip: shortest wins
    uint32_t* crow = tg.color + (size_t)Y * tg.w;
    float* zrow = tg.depth ? tg.depth + (size_t)Y * tg.w : nullptr;
    for (uint64_t k = 0; k < n; k++) {
      // Scanline::fragments raster.rs:60-69: var.z_div(pos.z)
      float z = v.v[2];
      float var[MAXL];
      for (int i = 0; i < L; i++) var[i] = ((d.persp_mask >> i) & 1) ? v.v[3 + i] / z : v.v[3 + i];
      bool pass = true;
      if (zrow && d.depth_test != RF_DEPTH_NONE) {  // ctx.rs:86-89: curr.partial_cmp(&new)
        float curr = zrow[X0 + k];
        if (d.depth_test == RF_DEPTH_LESS) pass = curr < z;
        else if (d.depth_test == RF_DEPTH_EQUAL) pass = curr == z;
        else pass = curr > z;
      }
      if (pass) {
        uint8_t c[4];
        if (shade_fragment(d, tex, var, c, &tex_oob)) {
          if (d.color_write) {
            st.frags_o += 1;
            crow[X0 + k] = pack_pixel(tg.fmt, c);
          }
          if (zrow && d.depth_write) zrow[X0 + k] = z;
        }
      }
      for (int i = 0; i < NL; i++) v.v[i] = v.v[i] + dvdx.v[i];  // vary.rs:146-154
    }
  }

  // raster.rs:248-302 + ScanlineIter::next 80-114
  void scan(float y0, float y1, const Lanes& l0, const Lanes& l1, const Lanes& r0, const Lanes& r1) {
    float recip_dy = 1.0f / (y1 - y0);
    Lanes dl, dr, dvdx;
    for (int i = 0; i < NL; i++) dl.v[i] = (l1.v[i] - l0.v[i]) * recip_dy;  // space.rs:205-207
    for (int i = 0; i < NL; i++) dr.v[i] = (r1.v[i] - r0.v[i]) * recip_dy;
    {
      Lanes ls, rs;
      for (int i = 0; i < NL; i++) ls.v[i] = l0.v[i] + dl.v[i];
      for (int i = 0; i < NL; i++) rs.v[i] = r0.v[i] + dr.v[i];
      float dx = rs.v[0] - ls.v[0];
      float rdx = 1.0f / dx;
      for (int i = 0; i < NL; i++) dvdx.v[i] = (rs.v[i] - ls.v[i]) * rdx;
    }
    float y0r = round_up_to_half(y0);
    float y1r = round_up_to_half(y1);
    float tw = y0r - y0;
    Lanes left;
    for (int i = 0; i < NL; i++) left.v[i] = l0.v[i] + ((l0.v[i] + dl.v[i]) - l0.v[i]) * tw;  // lerp(l0, l0.step(dl), tw)
    float right = r0.v[0] + dr.v[0] * tw;
    float y = y0r;
    uint32_t n = sat_u32(y1r - y0r);
    while (n != 0) {
      Lanes v0 = left;
      for (int i = 0; i < NL; i++) left.v[i] = left.v[i] + dl.v[i];
      float x1 = right;
      right = right + dr.v[0];
      float x0r = round_up_to_half(v0.v[0]);
      float x1r = round_up_to_half(x1);
      float t = x0r - v0.v[0];
      Lanes v;
      for (int i = 0; i < NL; i++) v.v[i] = v0.v[i] + ((v0.v[i] + dvdx.v[i]) - v0.v[i]) * t;
      uint32_t cnt = sat_u32(x1r - x0r);
      scanline(sat_usize(y), sat_usize(x0r), sat_usize(x1r), cnt, v, dvdx);
      if (oob) return;
      y = y + 1.0f;
      n--;
    }
  }

  // raster.rs:122-177 — one-pixel-thick line: every pixel carries v0's position and attributes
  void line(Lanes v0, Lanes v1) {
    if (v0.v[1] > v1.v[1]) std::swap(v0, v1);
    const float dx = v1.v[0] - v0.v[0], dy = v1.v[1] - v0.v[1];
    Lanes zero;
    for (int i = 0; i < 3 + MAXL; i++) zero.v[i] = 0.0f;  // vary_to(vs, vs, 1): step = (vs - vs) * 1
    if (fabsf(dx) > dy) {  // more wide than tall
      if (dx < 0.0f) std::swap(v0, v1);  // always draw from left to right (dx, dy keep their old values)
      const float x0 = round_up_to_half(v0.v[0]), x1 = round_up_to_half(v1.v[0]);
      const float dy_dx = dy / dx;
      float y = v0.v[1] + dy_dx * (x0 - v0.v[0]);
      const uint64_t xa = sat_usize(x0), xb = sat_usize(x1);
      for (uint64_t x = xa; x < xb; x++) {
        scanline(sat_usize(y), x, x + 1, 1, v0, zero);
        if (oob) return;
        y = y + dy_dx;
      }
    } else {  // more tall than wide
      const float y0 = round_up_to_half(v0.v[1]), y1 = round_up_to_half(v1.v[1]);
      const float dx_dy = dx / dy;
      float x = v0.v[0] + dx_dy * (y0 - v0.v[1]);
      const uint64_t ya = sat_usize(y0), yb = sat_usize(y1);
      for (uint64_t yy = ya; yy < yb; yy++) {
        const uint64_t xi = sat_usize(x);
        scanline(yy, xi, xi + 1, 1, v0, zero);
        if (oob) return;
        x = x + dx_dy;
      }
    }
  }

  // raster.rs:185-224
  void tri_fill(const Lanes in[3]) {
    Lanes s[3] = {in[0], in[1], in[2]};
    // stable sort by y with total_cmp
    auto key = [](float f) {
      int32_t b;
      std::memcpy(&b, &f, 4);
      return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1);
    };
    std::stable_sort(s, s + 3, [&](const Lanes& a, const Lanes& b) { return key(a.v[1]) < key(b.v[1]); });
    const Lanes &top = s[0], &mid0 = s[1], &bot = s[2];
    float top_y = top.v[1], mid_y = mid0.v[1], bot_y = bot.v[1];
    float t = (mid_y - top_y) / (bot_y - top_y);
    Lanes mid1;
    for (int i = 0; i < NL; i++) mid1.v[i] = lerp(top.v[i], bot.v[i], t);
    const Lanes *left, *right;
    if (mid0.v[0] < mid1.v[0]) { left = &mid0; right = &mid1; }
    else { left = &mid1; right = &mid0; }
    scan(top_y, mid_y, top, *left, top, *right);
    if (oob) return;
    scan(mid_y, bot_y, *left, bot, *right, bot);
  }
};

}  // namespace

// =========================================================================================
// C entry points (ctypes from tests/ and bench.py)
// =========================================================================================
extern "C" {

struct rfo_target {
  uint32_t w, h, fmt;
  uint32_t* color;
  float* depth;
  uint32_t band_y0, band_y1;  // [0,h) for all rows
};
struct rfo_texture {
  uint32_t w, h, fmt;
  const uint8_t* data;
  uint64_t stride;
};

// render() — render.rs:134-207. Returns rf_status. Stats are ADDED to *st.
int rfo_render(const rf_draw* dp, const rfo_texture* texp, rfo_target* tp, rf_stats* st) {
  const rf_draw& d = *dp;
  const int L = (int)d.n_attr_lanes;
  if (L > MAXL || d.vert_stride_f32 < 3u + (uint32_t)L) return RF_E_INVALID;
  if (d.depth_sort > RF_SORT_BACK_TO_FRONT) return RF_E_INVALID;
  if (d.bbox_cull) {
    // The scene loop around render(): BBox::visibility (scene.rs:59-87) with the draw's model-to-projection matrix; a Hidden
    // object is not rendered at all (crates.rs:100-122): only objs.i is counted (crates.rs:101,131).
    if (d.bbox_cull > 1 || d.vs == RF_VS_SPRITE) return RF_E_INVALID;
    st->objs_i += 1;
    uint8_t all = 0x3F;
    for (int k = 0; k < 8; k++) {  // BBox::verts scene.rs:62-69
      const float p[3] = {d.bbox[(k & 4) ? 3 : 0], d.bbox[(k & 2) ? 4 : 1], d.bbox[(k & 1) ? 5 : 2]};
      float pos[4];
      apply_proj(d.vs_uniform, p, pos);
      all &= outcode(pos);          // view_frustum::status clip.rs:245-267: Hidden <=> all corners outside one plane
    }
    if (all != 0) return RF_OK;
    st->objs_o += 1;
  }
  Tex tex{};
  if (texp) tex = Tex{texp->w, texp->h, texp->fmt, texp->data, (size_t)texp->stride};
  const bool needs_tex = d.fs == RF_FS_TEX_CLAMP_LIT || d.fs == RF_FS_TEX_CLAMP || d.fs == RF_FS_TEX_REPEAT_POT || d.fs == RF_FS_TEX_ONCE;
  if (needs_tex && !texp) return RF_E_INVALID;
  if (d.fs == RF_FS_TEX_REPEAT_POT && ((tex.w & (tex.w - 1)) || (tex.h & (tex.h - 1)) || !tex.w || !tex.h))
    return RF_E_BAD_TEXTURE;  // tex.rs:230-231

  auto t0 = std::chrono::steady_clock::now();
  rf_stats s{};
  s.calls = 1;
  s.prims_i = d.n_prims;
  s.verts_i = d.n_verts;

  // 1. vertex shader + outcodes   render.rs:158-165
  std::vector<ClipVert> cvs(d.n_verts);
  for (uint32_t i = 0; i < d.n_verts; i++) shade_vertex(d, d.verts + (size_t)i * d.vert_stride_f32, cvs[i]);

  Target tg{tp->w, tp->h, tp->fmt, tp->color, tp->depth, tp->band_y0, tp->band_y1};
  Raster R{d, texp ? &tex : nullptr, tg, s, 3 + L};
  const float* VP = d.viewport;
  int status = RF_OK;
  auto screen = [&](const ClipVert& cv, Lanes& out) {  // prim.rs:62-88
    float w = cv.pos[3];
    float p[4] = {cv.pos[0] / w, cv.pos[1] / w, 1.0f / w, 1.0f};  // pt3(x,y,1).z_div(w)
    for (int r = 0; r < 3; r++) out.v[r] = dot4(VP + 4 * r, p);   // mat.rs:945-949
    for (int i = 0; i < L; i++) out.v[3 + i] = ((d.persp_mask >> i) & 1) ? cv.attr[i] / w : cv.attr[i];
    for (int i = L; i < MAXL; i++) out.v[3 + i] = 0.0f;
  };

  if (d.prim_kind == RF_PRIM_EDGES) {
    // Render for Edge<usize> (prim.rs:41-60) + Clip for [Edge] (clip.rs:311-348) + raster::line
    struct CEdge { ClipVert a, b; };
    std::vector<CEdge> clipped;
    for (uint32_t p = 0; p < d.n_prims; p++) {
      const uint32_t* idx = d.indices + 2 * (size_t)p;
      if (idx[0] >= d.n_verts || idx[1] >= d.n_verts) return RF_E_INDEX_OOB;
      ClipVert a = cvs[idx[0]], b = cvs[idx[1]];
      if ((a.oc & b.oc) != 0) continue;                                    // both outside one plane
      if ((a.oc | b.oc) == 0) { clipped.push_back(CEdge{a, b}); continue; }  // neither outside
      bool keep = true;
      for (int pl = 0; pl < 6 && keep; pl++) {
        const uint8_t bit = (uint8_t)(1u << pl);
        const bool a_in = (a.oc & bit) == 0, b_in = (b.oc & bit) == 0;
        if (!a_in && !b_in) { keep = false; break; }
        ClipVert x;
        if (intersect(pl, a, b, L, x)) {
          if (a_in) b = x;
          else if (b_in) a = x;
        }
      }
      if (keep) clipped.push_back(CEdge{a, b});
    }
    for (const CEdge& e : clipped) {
      Lanes sa, sb;
      screen(e.a, sa);
      screen(e.b, sb);
      // Render::is_backface defaults to false (render.rs:72-74): only FaceCull::Front culls an edge (ctx.rs:95-101)
      if (d.face_cull == RF_CULL_FRONT) continue;
      s.prims_o += 1;
      s.verts_o += 3;  // render.rs:196 adds 3 whatever the primitive
      R.line(sa, sb);
      if (R.oob) { status = RF_E_TARGET_OOB; break; }
      if (R.tex_oob) { status = RF_E_BAD_TEXTURE; break; }
    }
    auto t1e = std::chrono::steady_clock::now();
    s.time_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1e - t0).count();
    st->calls += s.calls;
    st->prims_i += s.prims_i; st->prims_o += s.prims_o;
    st->verts_i += s.verts_i; st->verts_o += s.verts_o;
    st->frags_i += s.frags_i; st->frags_o += s.frags_o;
    st->time_ns += s.time_ns;
    return status;
  }

  // 2+3. assembly and clipping     render.rs:168-177; clip.rs:350-400
  struct CTri { ClipVert v[3]; };
  std::vector<CTri> clipped;
  clipped.reserve(d.n_prims / 2);
  std::vector<ClipVert> vin, vout;
  for (uint32_t p = 0; p < d.n_prims; p++) {
    const uint32_t* idx = d.indices + 3 * (size_t)p;
    if (idx[0] >= d.n_verts || idx[1] >= d.n_verts || idx[2] >= d.n_verts) return RF_E_INDEX_OOB;
    CTri t{{cvs[idx[0]], cvs[idx[1]], cvs[idx[2]]}};
    uint8_t all = t.v[0].oc & t.v[1].oc & t.v[2].oc;
    uint8_t any = t.v[0].oc | t.v[1].oc | t.v[2].oc;
    if (all != 0) continue;                          // Hidden
    if (any == 0) { clipped.push_back(t); continue; }  // Visible
    clip_tri(t.v, L, vin, vout);
    for (size_t k = 1; k + 1 < vout.size(); k++) clipped.push_back(CTri{{vout[0], vout[k], vout[k + 1]}});
  }

  // Optional depth sort   render.rs:180-182, 209-219; Render::depth prim.rs:21-23.
  // The reference's sort_unstable_by leaves the order of equal depths unspecified; this restatement
  // (and the CUDA path) keep equal depths in primitive order. Edges all have depth +inf
  // (render.rs:68-70), i.e. they keep their order, so the edge branch above has nothing to sort.
  if (d.depth_sort) {
    auto total_key = [](float f) { int32_t b; std::memcpy(&b, &f, 4); return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1); };  // f32::total_cmp
    auto depth = [](const CTri& t) { return ((t.v[0].pos[2] + t.v[1].pos[2]) + t.v[2].pos[2]) / 3.0f; };
    const bool f2b = d.depth_sort == RF_SORT_FRONT_TO_BACK;
    std::stable_sort(clipped.begin(), clipped.end(), [&](const CTri& t, const CTri& u) {
      const int32_t z = total_key(depth(t)), w = total_key(depth(u));
      return f2b ? z < w : w < z;
    });
  }

  // 4. per primitive: to_screen, cull, rasterise     render.rs:185-205
  for (const CTri& t : clipped) {
    Lanes scr[3];
    for (int k = 0; k < 3; k++) screen(t.v[k], scr[k]);
    // geom/prim.rs:150-156,288-294 ; vec.rs:443-445,480-482
    float abx = scr[1].v[0] - scr[0].v[0], aby = scr[1].v[1] - scr[0].v[1];
    float acx = scr[2].v[0] - scr[0].v[0], acy = scr[2].v[1] - scr[0].v[1];
    float perp[2] = {-aby, abx}, ac[2] = {acx, acy};
    bool back = dot2(perp, ac) < 0.0f;
    if ((d.face_cull == RF_CULL_BACK && back) || (d.face_cull == RF_CULL_FRONT && !back)) continue;  // ctx.rs:95-101
    s.prims_o += 1;
    s.verts_o += 3;
    R.tri_fill(scr);
    if (R.oob) { status = RF_E_TARGET_OOB; break; }
    if (R.tex_oob) { status = RF_E_BAD_TEXTURE; break; }
  }
  auto t1 = std::chrono::steady_clock::now();
  s.time_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
  st->calls += s.calls;
  st->prims_i += s.prims_i; st->prims_o += s.prims_o;
  st->verts_i += s.verts_i; st->verts_o += s.verts_o;
  st->frags_i += s.frags_i; st->frags_o += s.frags_o;
  st->time_ns += s.time_ns;
  return status;
}

// ---- small probes used by the golden tests (KATs of the reference's unit tests) ------------

// view_frustum::clip on one triangle given by clip-space positions; returns number of output
// triangles and writes up to 7*3*4 floats of positions. render/clip.rs:667-719 uses this shape.
int rfo_clip_tri(const float pos[12], float out_pos[84]) {
  ClipVert t[3];
  for (int k = 0; k < 3; k++) {
    std::memcpy(t[k].pos, pos + 4 * k, 16);
    std::memset(t[k].attr, 0, sizeof t[k].attr);
    t[k].oc = outcode(t[k].pos);
  }
  uint8_t all = t[0].oc & t[1].oc & t[2].oc, any = t[0].oc | t[1].oc | t[2].oc;
  if (all) return 0;
  if (!any) { std::memcpy(out_pos, pos, 48); return 1; }
  std::vector<ClipVert> vin, vout;
  clip_tri(t, 0, vin, vout);
  int n = 0;
  for (size_t k = 1; k + 1 < vout.size(); k++, n++) {
    std::memcpy(out_pos + 12 * n + 0, vout[0].pos, 16);
    std::memcpy(out_pos + 12 * n + 4, vout[k].pos, 16);
    std::memcpy(out_pos + 12 * n + 8, vout[k + 1].pos, 16);
  }
  return n;
}

// The exhaustive lattice test of clip.rs:667-719 run natively (1.95 M triangles): fills
// hist[8] with the output-triangle-count histogram and returns the number of out-of-bounds
// output vertices (the reference asserts 0).
int64_t rfo_clip_lattice_histogram(int64_t hist[8]) {
  for (int i = 0; i < 8; i++) hist[i] = 0;
  int64_t bad = 0;
  float out[84];
  const float w = 1.0f;
  float c[5] = {-2.0f, -1.0f, 0.0f, 1.0f, 2.0f};
  float pos[12];
  for (int i = 0; i < 1953125; i++) {
    int r = i;
    for (int k = 0; k < 3; k++) {
      for (int j = 0; j < 3; j++) { pos[4 * k + j] = c[r % 5]; r /= 5; }
      pos[4 * k + 3] = w;
    }
    int n = rfo_clip_tri(pos, out);
    hist[n]++;
    for (int q = 0; q < n * 3; q++)  // in_bounds(), clip.rs:724-728
      for (int j = 0; j < 4; j++)
        if (!(fabsf(out[4 * q + j] / out[4 * q + 3]) <= 1.00001f)) bad++;
  }
  return bad;
}

uint8_t rfo_outcode(const float pos[4]) { return outcode(pos); }

// ClipPlane::intersect (clip.rs:121-145) on one edge, for the edge_clip_* KATs (clip.rs:455-486): returns 1 and the
// intersection position if the edge crosses plane `plane` (index into PLANES), else 0.
int rfo_edge_plane(int plane, const float a[4], const float b[4], float out[4]) {
  ClipVert va{}, vb{}, vx{};
  std::memcpy(va.pos, a, 16); std::memcpy(vb.pos, b, 16);
  va.oc = outcode(va.pos); vb.oc = outcode(vb.pos);
  if (!intersect(plane, va, vb, 0, vx)) return 0;
  std::memcpy(out, vx.pos, 16);
  return 1;
}

// tri_fill (raster.rs:185-224) with a recording callback: for KATs raster.rs:326-401.
// lanes_in: 3 x (3+L) floats. For each scanline writes (y, x0, x1) to spans and the
// z-divided lane-0 varying per fragment to frag_vals (if non-null, up to max_frags).
struct SpanRec { uint64_t y, x0, x1; };
int rfo_tri_fill_spans(const float* lanes_in, int L, uint32_t persp_mask, SpanRec* spans, int max_spans,
                       float* frag_vals, int max_frags, int* n_frags_out) {
  // Drive Raster with a 1-lane recording pseudo-target: emulate by a tiny custom walk.
  struct Rec {
    SpanRec* spans; int max_spans; int n = 0;
    float* fv; int max_f; int nf = 0;
  } rec{spans, max_spans, 0, frag_vals, max_frags, 0};
  const int NL = 3 + L;
  Lanes s[3];
  for (int k = 0; k < 3; k++) {
    for (int i = 0; i < 3 + MAXL; i++) s[k].v[i] = 0.0f;
    for (int i = 0; i < NL; i++) s[k].v[i] = lanes_in[k * NL + i];
  }
  auto key = [](float f) { int32_t b; std::memcpy(&b, &f, 4); return b ^ (int32_t)(((uint32_t)(b >> 31)) >> 1); };
  std::stable_sort(s, s + 3, [&](const Lanes& a, const Lanes& b) { return key(a.v[1]) < key(b.v[1]); });
  const Lanes &top = s[0], &mid0 = s[1], &bot = s[2];
  float t = (mid0.v[1] - top.v[1]) / (bot.v[1] - top.v[1]);
  Lanes mid1;
  for (int i = 0; i < NL; i++) mid1.v[i] = lerp(top.v[i], bot.v[i], t);
  const Lanes* left = (mid0.v[0] < mid1.v[0]) ? &mid0 : &mid1;
  const Lanes* right = (mid0.v[0] < mid1.v[0]) ? &mid1 : &mid0;
  auto scan = [&](float y0, float y1, const Lanes& l0, const Lanes& l1, const Lanes& r0, const Lanes& r1) {
    float rdy = 1.0f / (y1 - y0);
    Lanes dl, dr, dvdx, ls, rs;
    for (int i = 0; i < NL; i++) dl.v[i] = (l1.v[i] - l0.v[i]) * rdy;
    for (int i = 0; i < NL; i++) dr.v[i] = (r1.v[i] - r0.v[i]) * rdy;
    for (int i = 0; i < NL; i++) ls.v[i] = l0.v[i] + dl.v[i];
    for (int i = 0; i < NL; i++) rs.v[i] = r0.v[i] + dr.v[i];
    float rdx = 1.0f / (rs.v[0] - ls.v[0]);
    for (int i = 0; i < NL; i++) dvdx.v[i] = (rs.v[i] - ls.v[i]) * rdx;
    float y0r = round_up_to_half(y0), y1r = round_up_to_half(y1), tw = y0r - y0;
    Lanes lft;
    for (int i = 0; i < NL; i++) lft.v[i] = l0.v[i] + ((l0.v[i] + dl.v[i]) - l0.v[i]) * tw;
    float rgt = r0.v[0] + dr.v[0] * tw;
    float y = y0r;
    uint32_t n = sat_u32(y1r - y0r);
    while (n--) {
      Lanes v0 = lft;
      for (int i = 0; i < NL; i++) lft.v[i] = lft.v[i] + dl.v[i];
      float x1 = rgt;
      rgt = rgt + dr.v[0];
      float x0r = round_up_to_half(v0.v[0]), x1r = round_up_to_half(x1);
      Lanes v;
      for (int i = 0; i < NL; i++) v.v[i] = v0.v[i] + ((v0.v[i] + dvdx.v[i]) - v0.v[i]) * (x0r - v0.v[0]);
      uint32_t cnt = sat_u32(x1r - x0r);
      if (rec.n < rec.max_spans) rec.spans[rec.n] = SpanRec{sat_usize(y), sat_usize(x0r), sat_usize(x1r)};
      rec.n++;
      for (uint32_t k = 0; k < cnt; k++) {
        if (L > 0 && rec.fv && rec.nf < rec.max_f)
          rec.fv[rec.nf] = (persp_mask & 1) ? v.v[3] / v.v[2] : v.v[3];
        rec.nf++;
        for (int i = 0; i < NL; i++) v.v[i] = v.v[i] + dvdx.v[i];
      }
      y = y + 1.0f;
    }
  };
  scan(top.v[1], mid0.v[1], top, *left, top, *right);
  scan(mid0.v[1], bot.v[1], *left, bot, *right, bot);
  if (n_frags_out) *n_frags_out = rec.nf;
  return rec.n;
}

// Sampler KATs (tex.rs:381-418): kind 0 clamp, 1 repeat_pot, 2 once (a texel outside the texture, where the reference panics, gives 4 x 0xEE)
void rfo_sample(const rfo_texture* t, int kind, float u, float v, uint8_t rgba[4]) {
  Tex tex{t->w, t->h, t->fmt, t->data, (size_t)t->stride};
  if (kind == 0) sample_clamp(tex, u, v, rgba);
  else if (kind == 2) { if (!sample_once(tex, u, v, rgba)) rgba[0] = rgba[1] = rgba[2] = rgba[3] = 0xEE; }
  else sample_repeat_pot(tex, u, v, rgba);
}

uint32_t rfo_pack_pixel(uint32_t fmt, const uint8_t rgba[4]) { return pack_pixel(fmt, rgba); }

}  // extern "C"
