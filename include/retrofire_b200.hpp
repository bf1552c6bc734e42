// retrofire_b200.hpp — header-only C++17 host-side mirror of retrofire-core's render API over the
// C ABI (retrofire_b200.h): Context / Stats / Shader / Target / render() / Batch with the reference's
// names and argument meaning. The reference is compiled Rust and no Rust toolchain is present in
// the build image, so this is the compiled-language host layer; rust/ holds the Rust shim sources.
//
//   re::Gpu gpu(0);
//   re::Target fb(gpu, 640, 480, RF_FMT_RGBA8888, /*depth*/false);
//   re::Context ctx;                                  // render/ctx.rs defaults
//   auto sh = re::shader::make(RF_VS_MVP, RF_FS_COLOR3F, /*lanes*/3, /*persp*/0);
//   re::render(tris, 1, verts, 3, 6, sh, mvp, viewport, fb, ctx);   // render.rs:134-147
//   fb.download(pixels.data(), 640);
#pragma once
#include <array>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "retrofire_b200.h"

namespace re {

struct Error : std::runtime_error {  // the reference panics where this is thrown
  rf_status status;
  Error(rf_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

struct Throughput { size_t i = 0, o = 0; };
struct Stats {  // render/stats.rs:16-40
  double time = 0; float calls = 0, frames = 0;
  Throughput objs, prims, verts, frags;
  Stats& operator+=(const Stats& s) {  // stats.rs:199-208
    time += s.time; calls += s.calls; frames += s.frames;
    objs.i += s.objs.i; objs.o += s.objs.o; prims.i += s.prims.i; prims.o += s.prims.o;
    verts.i += s.verts.i; verts.o += s.verts.o; frags.i += s.frags.i; frags.o += s.frags.o;
    return *this;
  }
};

struct BBox { std::array<float, 3> low, upp; };  // render/scene.rs:22, model space

struct Context {  // render/ctx.rs:11-127 (defaults :104-127); depth_sort: RF_SORT_* (ctx.rs:39)
  bool has_color_clear = true; std::array<uint8_t, 4> color_clear{0, 0, 0, 0xFF};
  bool has_depth_clear = true; float depth_clear = __builtin_inff();
  uint8_t face_cull = RF_CULL_BACK, depth_test = RF_DEPTH_LESS, depth_sort = 0;
  bool color_write = true, depth_write = true;
  mutable Stats stats;
};

class Gpu {
 public:
  explicit Gpu(int device = 0, void* stream = nullptr) { if (rf_status s = rf_ctx_create(device, stream, &c_)) throw Error(s, "rf_ctx_create: no sm_100 GPU"); }
  ~Gpu() { rf_ctx_destroy(c_); }
  Gpu(const Gpu&) = delete; Gpu& operator=(const Gpu&) = delete;
  rf_ctx* raw() const { return c_; }
  void check(rf_status s) const { if (s) throw Error(s, rf_last_error(c_)); }
  void flush() { check(rf_flush(c_)); }
  void sync() { check(rf_sync(c_)); }
 private:
  rf_ctx* c_ = nullptr;
};

struct Texture {  // render/tex.rs:33-37
  Texture(Gpu& g, uint32_t w, uint32_t h, rf_texel_fmt fmt, const void* data, size_t stride) : g_(g) { g.check(rf_texture_create(g.raw(), w, h, fmt, data, stride, &t_)); }
  ~Texture() { rf_texture_destroy(t_); }
  rf_texture* raw() const { return t_; }
 private:
  Gpu& g_; rf_texture* t_ = nullptr;
};

struct Shader {  // render/shader.rs:88-126, catalogue pair
  uint32_t vs, fs, lanes, persp_mask; std::array<float, RF_FS_UNIFORM_F32> fs_uniform{}; const Texture* texture = nullptr;
};
namespace shader {
inline Shader make(uint32_t vs, uint32_t fs, uint32_t lanes, uint32_t persp_mask, const Texture* tex = nullptr) { return Shader{vs, fs, lanes, persp_mask, {}, tex}; }
}

class Target {  // Framebuf / Colorbuf / Buf2<Color4> resident on the device (render/target.rs:35-136)
 public:
  Target(Gpu& g, uint32_t w, uint32_t h, rf_color_fmt fmt, bool depth) : g_(g), w(w), h(h) { g.check(rf_target_create(g.raw(), w, h, fmt, depth, &t_)); }
  ~Target() { rf_target_destroy(t_); }
  void clear(const Context& ctx) {  // Frame::clear, front/src/lib.rs:103-120
    const float z = 1.0f / ctx.depth_clear;
    g_.check(rf_target_clear(g_.raw(), t_, ctx.has_color_clear ? ctx.color_clear.data() : nullptr, ctx.has_depth_clear ? &z : nullptr));
  }
  void download(void* host, size_t stride_elems) { g_.check(rf_target_download_color(g_.raw(), t_, host, stride_elems)); }
  void download_depth(float* host, size_t stride_elems) { g_.check(rf_target_download_depth(g_.raw(), t_, host, stride_elems)); }
  rf_target* raw() const { return t_; }
  Gpu& gpu() const { return g_; }
  const uint32_t w, h;
 private:
  Gpu& g_; rf_target* t_ = nullptr;
};

// render() — render.rs:134-207. verts: n_verts records of `stride` floats [x,y,z,a0..]; uniform: up to 32 floats
// (row-major matrices, layout per rf_vs_id); to_screen: row-major 4x4. Synchronous like the reference:
// when it returns the draw has executed and ctx.stats has been updated (render.rs:206).
// `edges`: prims are Edge<usize> pairs (render/prim.rs:41-60) instead of Tri<usize> triples. `bbox`: the scene loop's
// `if obj.bbox.visibility(&model_to_project) == Hidden { continue }` (scene.rs:81-87, crates.rs:100-122) evaluated on the device.
inline void render(const uint32_t* prims, uint32_t n_prims, const float* verts, uint32_t n_verts, uint32_t stride, const Shader& sh,
                   const float* uniform, size_t n_uniform, const float to_screen[16], Target& target, const Context& ctx,
                   bool edges = false, const BBox* bbox = nullptr) {
  rf_draw d{};
  d.prim_kind = edges ? RF_PRIM_EDGES : RF_PRIM_TRIS;
  if (bbox) { d.bbox_cull = 1; std::memcpy(d.bbox, bbox->low.data(), 12); std::memcpy(d.bbox + 3, bbox->upp.data(), 12); }
  d.indices = prims; d.n_prims = n_prims; d.verts = verts; d.n_verts = n_verts; d.vert_stride_f32 = stride;
  d.n_attr_lanes = sh.lanes; d.persp_mask = sh.persp_mask; d.vs = sh.vs; d.fs = sh.fs;
  std::memcpy(d.vs_uniform, uniform, sizeof(float) * (n_uniform < RF_VS_UNIFORM_F32 ? n_uniform : RF_VS_UNIFORM_F32));
  std::memcpy(d.fs_uniform, sh.fs_uniform.data(), sizeof d.fs_uniform);
  d.texture = sh.texture ? sh.texture->raw() : nullptr;
  std::memcpy(d.viewport, to_screen, sizeof d.viewport);
  d.face_cull = ctx.face_cull; d.depth_test = ctx.depth_test; d.color_write = ctx.color_write; d.depth_write = ctx.depth_write; d.depth_sort = ctx.depth_sort;
  rf_stats st{};
  target.gpu().check(rf_render(target.gpu().raw(), target.raw(), &d, &st));
  Stats s; s.time = st.time_ns * 1e-9; s.calls = (float)st.calls;
  s.prims = {st.prims_i, st.prims_o}; s.verts = {st.verts_i, st.verts_o}; s.frags = {st.frags_i, st.frags_o};
  s.objs = {st.objs_i, st.objs_o};
  ctx.stats += s;
}

// Batch — render/batch.rs:31-147: owns clones of prims/verts, `.render()` calls render().
struct Batch {
  std::vector<uint32_t> prims; std::vector<float> verts; uint32_t stride = 3;
  std::vector<float> uniform_; Shader shader_{}; std::array<float, 16> viewport_{}; Target* target_ = nullptr; const Context* ctx_ = nullptr;
  Batch& primitives(const uint32_t* p, size_t n) { prims.assign(p, p + 3 * n); return *this; }
  Batch& vertices(const float* v, size_t n, uint32_t s) { verts.assign(v, v + n * s); stride = s; return *this; }
  Batch& uniform(const float* u, size_t n) { uniform_.assign(u, u + n); return *this; }
  Batch& shader(const Shader& s) { shader_ = s; return *this; }
  Batch& viewport(const float m[16]) { std::memcpy(viewport_.data(), m, 64); return *this; }
  Batch& target(Target& t) { target_ = &t; return *this; }
  Batch& context(const Context& c) { ctx_ = &c; return *this; }
  Batch clone() const { return *this; }  // #[derive(Clone)], batch.rs:31: crates.rs clones one configured batch per object
  void render() { re::render(prims.data(), (uint32_t)(prims.size() / 3), verts.data(), (uint32_t)(verts.size() / stride), stride, shader_, uniform_.data(), uniform_.size(), viewport_.data(), *target_, *ctx_); }
};

}  // namespace re
