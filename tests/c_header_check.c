/* A plain C99 translation unit: include/retrofire_b200.h must be usable from C (the Rust -sys crate, cgo or a C host
 * bind exactly these declarations). Compiled (not run) by tests/test_build.py::test_header_is_plain_c. */
#include "retrofire_b200.h"

static rf_status draw_one(rf_ctx* ctx, rf_target* t, const float* verts, uint32_t n_verts, const uint32_t* tris, uint32_t n_tris) {
  rf_draw d;
  rf_stats st;
  unsigned i;
  for (i = 0; i < sizeof d; i++) ((unsigned char*)&d)[i] = 0;
  d.indices = tris; d.n_prims = n_tris;
  d.verts = verts; d.n_verts = n_verts; d.vert_stride_f32 = 6;
  d.n_attr_lanes = 3; d.persp_mask = 0;
  d.vs = RF_VS_MVP; d.fs = RF_FS_COLOR3F;
  d.face_cull = RF_CULL_BACK; d.depth_test = RF_DEPTH_LESS; d.color_write = 1; d.depth_write = 1;
  d.prim_kind = RF_PRIM_TRIS;
  return rf_render(ctx, t, &d, &st);
}

int rf_c_header_check(void) {
  rf_status (*f)(rf_ctx*, rf_target*, const float*, uint32_t, const uint32_t*, uint32_t) = draw_one;
  return f != 0 && rf_abi_version() == RF_ABI_VERSION && RF_N_KERNELS == 12 && sizeof(rf_draw) > 0 && RF_FMT_RGBA4444 > RF_FMT_RGBA8888 ? 0 : 1;
}
