// Compile-and-run check of include/retrofire_b200.hpp: hello_tri (core/examples/hello_tri.rs) through the C++ mirror.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../include/retrofire_b200.hpp"

int main() {
  try {
    re::Gpu gpu(0);
    const uint32_t w = 640, h = 480;
    re::Target fb(gpu, w, h, RF_FMT_RGBA8888, false);
    re::Context ctx;
    const float verts[] = {-1, 1, 0, 1, 0, 0, 1, 1, 0, 0, 0.8f, 0, 0, -1, 0, 0.4f, 0.4f, 1};
    const uint32_t tri[] = {0, 1, 2};
    // translate3(0,0,2).then(perspective(1, 640/480, 0.1..1000))  (values as mathx.py computes them)
    const float a = (float)w / (float)h, e22 = (1000.0f + 0.1f) / (1000.0f - 0.1f), e23 = 2.0f * 1000.0f * 0.1f / (0.1f - 1000.0f);
    const float mvp[16] = {1, 0, 0, 0, 0, a, 0, 0, 0, 0, e22, e22 * 2.0f + e23, 0, 0, 1, 2};
    const float vp[16] = {320, 0, 0, 320, 0, -240, 0, 240, 0, 0, 1, 0, 0, 0, 0, 1};
    auto sh = re::shader::make(RF_VS_MVP, RF_FS_COLOR3F, 3, 0);
    re::render(tri, 1, verts, 3, 6, sh, mvp, 16, vp, fb, ctx);
    std::vector<uint8_t> px((size_t)w * h * 4);
    fb.download(px.data(), w);
    const uint8_t* c = &px[((size_t)240 * w + 320) * 4];
    std::printf("center %u %u %u %u frags %zu/%zu\n", c[0], c[1], c[2], c[3], ctx.stats.frags.i, ctx.stats.frags.o);
    if (!(c[0] == 114 && c[1] == 102 && c[2] == 128 && c[3] == 255 && ctx.stats.frags.i == 51200)) return 1;
    // The same draw through re::Batch (batch.rs:31-147): configure once, clone per object, set uniform and target, render.
    re::Target fb2(gpu, w, h, RF_FMT_RGBA8888, false);
    re::Context ctx2;
    re::Batch base;
    base.primitives(tri, 1).vertices(verts, 3, 6).shader(sh).viewport(vp).context(ctx2);
    re::Batch b = base.clone();
    b.uniform(mvp, 16).target(fb2);
    b.render();
    b.render();  // a batch can be reused (batch.rs:27-30): Stats accumulate like render.rs:206
    std::vector<uint8_t> px2((size_t)w * h * 4);
    fb2.download(px2.data(), w);
    std::printf("batch frags %zu/%zu calls %.0f same %d\n", ctx2.stats.frags.i, ctx2.stats.frags.o, ctx2.stats.calls, (int)(px == px2));
    return (px == px2 && ctx2.stats.frags.i == 2 * 51200 && ctx2.stats.calls == 2.0f && base.target_ == nullptr) ? 0 : 3;
  } catch (const re::Error& e) {
    std::fprintf(stderr, "error %d: %s\n", (int)e.status, e.what());
    return 2;
  }
}
