// Exercises serstacker_b200/host/ssk_adapter.h the way a reference call site would (register_frame -> remap ->
// accumulate, c_image_stacking_pipeline.cc:1358-1862), on an analytic scene with known sub-pixel shifts.
//   adapter_smoke --no-gpu : option defaults / handle-free calls only (CPU test)
//   adapter_smoke          : full path on cuda:0; exit code 0 iff the recovered shifts are within 0.1 px
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "ssk_adapter.h"

static void render(ssk::Mat &m, int w, int h, float dx, float dy) {
  m.create(h, w, SSK_32FC1);
  const float bx[5] = {0.30f, 0.62f, 0.45f, 0.75f, 0.22f}, by[5] = {0.35f, 0.30f, 0.66f, 0.70f, 0.72f};
  const float bs[5] = {9.f, 6.f, 12.f, 5.f, 7.f}, ba[5] = {0.8f, 0.6f, 0.5f, 0.9f, 0.7f};
  for (int y = 0; y < h; ++y) {
    float *p = m.ptr<float>(y);
    for (int x = 0; x < w; ++x) {
      float v = 0.05f;
      for (int k = 0; k < 5; ++k) {
        const float ux = x - dx - bx[k] * w, uy = y - dy - by[k] * h;
        v += ba[k] * std::exp(-(ux * ux + uy * uy) / (2 * bs[k] * bs[k]));
      }
      p[x] = v * 0.5f;
    }
  }
}

int main(int argc, char **argv) {
  ssk_registration_options ro;
  ssk_registration_options_default(&ro);
  if (ssk_version() < 100 || ro.ecc.min_rho <= 0 || ro.ecc.scale <= 0) return 2;
  if (argc > 1 && !std::strcmp(argv[1], "--no-gpu")) {
    ssk::c_image_transform t(SSK_MOTION_AFFINE);
    if (t.parameters().size() != 6) return 3;
    t.set_translation(1.5f, -2.f);
    float tx = 0, ty = 0;
    t.translation(&tx, &ty);
    if (tx != 1.5f || ty != -2.f || !t.invertible()) return 3;
    t.reset();
    t.translation(&tx, &ty);
    if (tx != 0.f || ty != 0.f || t.parameters()[0] != 1.f) return 3;
    // eps(dp, size) / invert_and_compose(p, dp) (host arithmetic of libssk): a translation step of (3, 4) is 5 px long, and
    // composing the identity with the inverse of a step (dx, dy) on the affine translation terms moves by (-dx, -dy)
    ssk::c_image_transform tt(SSK_MOTION_TRANSLATION);
    if (tt.eps({3.f, 4.f}, 640, 480) != 5.0) return 3;
    tt.set_translation(10.f, -1.f);
    const std::vector<float> tn = tt.invert_and_compose({0.5f, 0.25f});
    if (tn.size() != 2 || tn[0] != 9.5f || tn[1] != -1.25f) return 3;
    const std::vector<float> an = t.invert_and_compose({0.f, 0.f, 2.f, 0.f, 0.f, -3.f});
    if (an.size() != 6 || an[0] != 1.f || an[4] != 1.f || an[2] != -2.f || an[5] != 3.f) return 3;
    if (t.eps({0.f, 0.f, 3.f, 0.f, 0.f, 4.f}, 640, 480) != 5.0 || !t.invert_and_compose({1.f}).empty()) return 3;
    // the reference's option structs: defaults of c_frame_registration.h:47-64, 119-136 survive the conversion
    ssk::c_image_registration_options io;
    const ssk_registration_options so = ssk::to_ssk_options(io);
    if (so.motion_type != SSK_MOTION_AFFINE || so.ecc.ecc_method != SSK_ECC_LM || so.ecc.scale != 0.5 || so.enable_ecc_registration) return 3;
    // c_eccflow_registration_options (c_frame_registration.h:88-100) and c_eccflow_options (ecc2.h:515-527)
    if (so.enable_eccflow_registration || so.eccflow.scale_factor != 0.75 || so.eccflow.max_iterations != 3 || so.eccflow.support_scale != 4 ||
        so.eccflow.min_image_size != -1 || so.eccflow.update_multiplier != 1.5) return 3;
    ssk_eccflow_options fo;
    ssk_eccflow_options_default(&fo);
    ssk::c_eccflow_options fa;
    if (fo.scale_factor != fa.scale_factor || fo.support_scale != fa.support_scale || fo.max_iterations != fa.max_iterations ||
        fo.min_image_size != fa.min_image_size || fo.noise_level != fa.noise_level) return 3;
    ssk::c_frame_registration feature_only(io);            // feature registration is the default stage: not in this library
    ssk::Mat dummy(8, 8, SSK_32FC1);
    if (feature_only.setup_reference_frame(dummy)) return 3;
    int lo = 0, hi = 0;
    ssk::c_image_stacking_pipeline::master_frame_range(100, 50, 30, &lo, &hi);
    if (lo != 35 || hi != 65) return 3;
    int nc = 0, nr = 0;
    ssk::c_ecch::compute_next_pyramid_layer_size(1920, 1080, &nc, &nr);
    if (nc != 960 || nr != 540) return 3;
    std::printf("adapter_smoke: no-gpu checks ok (version %d)\n", ssk_version());
    return 0;
  }
  const int W = 192, H = 144, N = 5;
  const float sx[N] = {0.f, 1.37f, -2.21f, 0.52f, 3.08f}, sy[N] = {0.f, -0.83f, 1.46f, 2.65f, -1.91f};
  ro.motion_type = SSK_MOTION_TRANSLATION;
  ro.enable_ecc_registration = 1;   // c_image_registration_options: the ECC stage is opt-in, as in the reference
  ro.ecc.ecch_max_level = -1;
  ro.ecc.ecc_method = SSK_ECC_INVERSE_COMPOSITIONAL_LM;   // what the GUI presets select
  ro.ecc.eps = 0.01;
  ro.ecc.update_step_scale = 1.0;
  ssk::c_frame_registration reg(ro);
  ssk::c_weigthed_average acc;
  ssk::Mat ref, frame, warped, mask;
  render(ref, W, H, 0, 0);
  if (!reg.setup_reference_frame(ref)) { std::fprintf(stderr, "setup_reference_frame: %s\n", ssk_last_error()); return 4; }
  double worst = 0;
  for (int i = 0; i < N; ++i) {
    render(frame, W, H, sx[i], sy[i]);
    if (!reg.register_frame(frame, ssk::Mat(), &warped, &mask)) { std::fprintf(stderr, "register_frame: %s\n", ssk_last_error()); return 5; }
    const std::vector<float> p = reg.image_transform()->parameters();
    std::printf("frame %d: estimated (%.4f, %.4f) true (%.2f, %.2f) rho %.4f iterations %d\n", i, p[0], p[1], sx[i], sy[i], reg.status().ecc.rho, reg.status().ecc.num_iterations);
    worst = std::fmax(worst, std::fmax(std::fabs(p[0] - sx[i]), std::fabs(p[1] - sy[i])));
    if (!acc.add(warped, mask)) { std::fprintf(stderr, "add: %s\n", ssk_last_error()); return 6; }
  }
  ssk::Mat avg, amask;
  if (!acc.compute(avg, &amask) || acc.accumulated_frames() != N) return 7;
  double err = 0; int cnt = 0;
  for (int y = 8; y < H - 8; ++y)
    for (int x = 8; x < W - 8; ++x)
      if (amask.ptr<uint8_t>(y)[x]) { err = std::fmax(err, std::fabs(avg.ptr<float>(y)[x] - ref.ptr<float>(y)[x])); ++cnt; }
  std::printf("adapter_smoke: worst |shift error| = %.4f px, stack max |avg - ref| = %.4g over %d px\n", worst, err, cnt);
  // the steps either side of the loop: the inpainted read-out keeps every valid pixel and fills the mask; a sharpened
  // copy of the reference keeps its size and type
  ssk::Mat filled, fmask, sharp;
  if (!acc.compute_inpainted(filled, &fmask)) { std::fprintf(stderr, "compute_inpainted: %s\n", ssk_last_error()); return 8; }
  int holes = 0, kept_bad = 0, unfilled = 0;
  for (int y = 0; y < H; ++y)
    for (int x = 0; x < W; ++x) {
      if (!amask.ptr<uint8_t>(y)[x]) ++holes;
      else if (filled.ptr<float>(y)[x] != avg.ptr<float>(y)[x]) ++kept_bad;
      if (!fmask.ptr<uint8_t>(y)[x]) ++unfilled;
    }
  std::printf("adapter_smoke: inpaint filled %d holes, %d valid pixels changed, %d left empty\n", holes, kept_bad, unfilled);
  if (kept_bad || unfilled) return 9;
  if (!ssk::unsharp_mask(ref, sharp, 1.0, 0.8) || sharp.rows != H || sharp.cols != W) { std::fprintf(stderr, "unsharp_mask: %s\n", ssk_last_error()); return 10; }
  // accumulator surface: counters, accumulator(), compute(..., ddepth)
  ssk::Mat cntr, accm, avg16;
  if (!acc.get_acc_counters(cntr) || !acc.counter(cntr) || cntr.type != SSK_32FC1 || !acc.accumulator(accm) || accm.rows != H) return 11;
  if (cntr.ptr<float>(H / 2)[W / 2] != (float)N) return 11;
  if (!acc.compute(avg16, nullptr, 65535.0, SSK_16U) || avg16.type != SSK_16UC1) return 11;
  if (std::abs((int)avg16.ptr<uint16_t>(H / 2)[W / 2] - (int)std::lround(avg.ptr<float>(H / 2)[W / 2] * 65535.0)) > 1) return 11;
  // c_ecch surface: two-step align, create_remap, pyramid images
  ssk::c_image_transform tr(SSK_MOTION_TRANSLATION);
  ssk::c_ecch ecch(&tr);
  ecch.set_maxlevel(-1);
  ecch.set_epsx(0.01);
  render(frame, W, H, 1.25f, -0.75f);
  ssk::Mat rm, rimg, cimg;
  if (!ecch.set_reference_image(ref) || !ecch.set_current_image(frame) || !ecch.align()) { std::fprintf(stderr, "c_ecch: %s\n", ssk_last_error()); return 12; }
  if (!ecch.create_remap(rm) || rm.type != SSK_32FC2 || rm.rows != H || !ecch.reference_image(rimg) || !ecch.current_image(cimg) || cimg.cols != W) return 12;
  std::printf("adapter_smoke: c_ecch two-step align -> (%.3f, %.3f), %d iterations\n", tr.parameters()[0], tr.parameters()[1], ecch.num_iterations());
  if (std::fabs(tr.parameters()[0] - 1.25f) > 0.1 || std::fabs(tr.parameters()[1] + 0.75f) > 0.1) return 12;
  // masked current frame through the adapter
  ssk::Mat cmask(H, W, SSK_8UC1);
  std::memset(cmask.buf.data(), 255, cmask.buf.size());
  for (int y = 20; y < 40; ++y) std::memset(cmask.ptr<uint8_t>(y) + 30, 0, 40);
  tr.reset();
  if (!ecch.align(frame, cmask) || std::fabs(tr.parameters()[0] - 1.25f) > 0.1) return 12;
  // the pipeline class: generated master frame + stacking pass + inpainted read-out
  std::vector<ssk::Mat> seq(N);
  for (int i = 0; i < N; ++i) render(seq[i], W, H, sx[i], sy[i]);
  ssk::c_image_stacking_pipeline pipe;
  for (ssk::c_image_registration_options *o : {&pipe.registration_options, &pipe.master_options.registration}) {
    o->motion_type = SSK_MOTION_TRANSLATION; o->enable_feature_registration = false; o->enable_ecc_registration = true;
    o->ecc.ecc_method = SSK_ECC_INVERSE_COMPOSITIONAL_LM; o->ecc.ecch_max_level = -1; o->ecc.eps = 0.01; o->ecc.update_step_scale = 1.0;
  }
  pipe.master_options.max_frames_to_generate_master_frame = 3;
  pipe.max_batch = 4;
  ssk::Mat stacked, smask;
  if (!pipe.run(seq, 0, stacked, smask) || pipe.accumulated_frames() != N || stacked.rows != H) { std::fprintf(stderr, "pipeline: %s\n", ssk_last_error()); return 13; }
  double perr = 0;
  for (int y = 12; y < H - 12; ++y)
    for (int x = 12; x < W - 12; ++x) perr = std::fmax(perr, std::fabs(stacked.ptr<float>(y)[x] - ref.ptr<float>(y)[x]));
  std::printf("adapter_smoke: c_image_stacking_pipeline::run stacked %d frames, max |stack - scene| = %.4g (master sharpened)\n", pipe.accumulated_frames(), perr);
  if (perr > 0.05) return 13;
  // c_eccflow: the flow between the reference and a shifted copy is the shift (the disk's textured interior)
  {
    ssk::c_eccflow flow;
    flow.set_support_scale(3);
    flow.set_max_iterations(3);
    flow.set_scale_factor(0.75);
    render(frame, W, H, 1.25f, -0.75f);
    ssk::Mat fmap, uv;
    if (!flow.compute(frame, ref, fmap, ssk::Mat()) || fmap.type != SSK_32FC2 || fmap.rows != H || !flow.current_uv(uv) || flow.num_levels() < 3) {
      std::fprintf(stderr, "c_eccflow: %s\n", ssk_last_error());
      return 15;
    }
    const float *c = uv.ptr<float>(H / 2) + 2 * (W / 2);
    std::printf("adapter_smoke: c_eccflow flow at the centre (%.3f, %.3f), true shift (1.25, -0.75), %d levels\n", c[0], c[1], flow.num_levels());
    if (std::fabs(c[0] - 1.25f) > 0.25 || std::fabs(c[1] + 0.75f) > 0.25) return 15;
    // and as the second stage of c_frame_registration
    ssk::c_image_registration_options fo;
    fo.motion_type = SSK_MOTION_TRANSLATION; fo.enable_feature_registration = false; fo.enable_ecc_registration = true;
    fo.enable_eccflow_registration = true;
    fo.ecc.ecc_method = SSK_ECC_INVERSE_COMPOSITIONAL_LM; fo.ecc.ecch_max_level = -1;
    ssk::c_frame_registration freg(fo);
    ssk::Mat fw, fm, cr;
    if (!freg.setup_reference_frame(ref) || !freg.register_frame(frame, ssk::Mat(), &fw, &fm) || !freg.current_remap(cr, W, H)) {
      std::fprintf(stderr, "c_frame_registration + eccflow: %s\n", ssk_last_error());
      return 15;
    }
    double ferr = 0;
    for (int y = 24; y < H - 24; ++y)
      for (int x = 24; x < W - 24; ++x) ferr = std::fmax(ferr, std::fabs(fw.ptr<float>(y)[x] - ref.ptr<float>(y)[x]));
    std::printf("adapter_smoke: register_frame with eccflow: max |warped - reference| = %.4g\n", ferr);
    if (ferr > 0.05) return 15;
  }
  // c_canvas_average and the up-scaling helpers
  {
    ssk::c_canvas_average canvas;
    canvas.setCanvasSize(2 * W, 2 * H);
    ssk::Mat ident(H, W, SSK_32FC2), cavg, cmsk;
    for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) { ident.ptr<float>(y)[2 * x] = (float)x; ident.ptr<float>(y)[2 * x + 1] = (float)y; }
    int bb[4];
    if (!canvas.add(ref)) { std::fprintf(stderr, "c_canvas_average: %s\n", ssk_last_error()); return 16; }
    canvas.last_bbox(bb);
    const int moved[4] = {bb[0] + 40, bb[1] + 10, W, H};
    if (!canvas.add(ref, ssk::Mat(), ident, moved) || canvas.accumulated_frames() != 2 || !canvas.compute(cavg, &cmsk)) return 16;
    int cc = 0, cr = 0;
    canvas.accumulator_size(&cc, &cr);
    if (cc != 2 * W || cr != 2 * H || cavg.cols != cc || cmsk.ptr<uint8_t>(bb[1] + 20)[bb[0] + W + 20] != 255 || cmsk.ptr<uint8_t>(2)[2] != 0) return 16;
    ssk::Mat up, upm, upmap;
    if (!ssk::upscale_image(ssk::frame_upscale_x15, ref, amask, up, &upm) || up.cols != W * 3 / 2 || upm.rows != H * 3 / 2) return 16;
    if (!ssk::upscale_remap(ssk::frame_upscale_pyrUp, ident, upmap) || upmap.cols != 2 * W) return 16;
    if (std::fabs(upmap.ptr<float>(50)[2 * 100] - 49.75f) > 0.26f) return 16;   // pyrUp of the identity map: x / 2 up to the half-pixel phase
    ssk::c_image_stacking_pipeline up_pipe;
    up_pipe.registration_options.motion_type = SSK_MOTION_TRANSLATION; up_pipe.registration_options.enable_feature_registration = false;
    up_pipe.registration_options.enable_ecc_registration = true; up_pipe.registration_options.ecc.ecc_method = SSK_ECC_INVERSE_COMPOSITIONAL_LM;
    up_pipe.registration_options.ecc.ecch_max_level = -1;
    up_pipe.master_options.generate_master_frame = false; up_pipe.master_options.unsharp_alpha = 0;
    up_pipe.upscale_options.upscale_option = ssk::frame_upscale_x15;
    up_pipe.max_batch = 4;
    ssk::Mat ustack, umask;
    if (!up_pipe.run(seq, 0, ustack, umask) || ustack.cols != W * 3 / 2 || ustack.rows != H * 3 / 2) { std::fprintf(stderr, "pipeline x1.5: %s\n", ssk_last_error()); return 16; }
    std::printf("adapter_smoke: c_canvas_average %dx%d canvas, upscale helpers, x1.5 stack %dx%d ok\n", cc, cr, ustack.cols, ustack.rows);
  }
  // input side
  ssk::Mat hmask(H, W, SSK_8UC1), lin, raw(H, W, SSK_16UC1), planes;
  std::memset(hmask.buf.data(), 255, hmask.buf.size());
  for (int y = 50; y < 60; ++y) std::memset(hmask.ptr<uint8_t>(y) + 70, 0, 25);
  if (!ssk::linear_interpolation_inpaint(ref, hmask, lin) || lin.ptr<float>(10)[10] != ref.ptr<float>(10)[10]) return 14;
  for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) raw.ptr<uint16_t>(y)[x] = (uint16_t)(100 * ((y & 1) * 2 + (x & 1)));
  if (!ssk::average_bayer_planes(raw, planes) || planes.rows != H / 2 || planes.ptr<uint16_t>(3)[5] != 150) return 14;
  // filter_bad_pixels (bad_pixels.cc:58-70, debayer.cc:1599-1611): a planted hot pixel goes, its neighbours stay
  raw.ptr<uint16_t>(40)[60] = 60000;
  const uint16_t keep = raw.ptr<uint16_t>(40)[62];
  if (!ssk::median_filter_bad_pixels(raw, 5.0, true) || raw.ptr<uint16_t>(40)[60] != 0 || raw.ptr<uint16_t>(40)[62] != keep) return 14;
  ssk::Mat hot(H, W, SSK_32FC1);
  std::memcpy(hot.buf.data(), ref.buf.data(), hot.buf.size());
  const float before = hot.ptr<float>(70)[90];
  hot.ptr<float>(70)[90] = before + 0.7f;
  if (!ssk::median_filter_bad_pixels(hot, 5.0) || std::fabs(hot.ptr<float>(70)[90] - before) > 0.2f) return 14;
  // c_saturn_derotation_remap (c_saturn_derotation_remap.cc:21-55) and one frame of c_sdr_pipeline's loop
  {
    ssk::c_saturn_derotation_remap sat;
    const double center[2] = {W / 2.0, H / 2.0}, axes[3] = {50, 45, 50}, pose[3] = {0.3, 0.05, -0.1};
    if (sat.rotation_period_sec() != 10 * 3600. + 33 * 60. + 38 || !sat.set_reference_pose(W, H, center, axes, pose)) return 17;
    if (!sat.compute_derotation_for_time(300.0, 0.8)) { std::fprintf(stderr, "saturn remap: %s\n", ssk_last_error()); return 17; }
    const int cy = H / 2, cx = W / 2;
    if (sat.rmask().ptr<uint8_t>(cy)[cx] != 255 || sat.rmask().ptr<uint8_t>(2)[2] != 0 || sat.rmap().ptr<float>(2)[2 * 2] != 2.f) return 17;
    if (sat.wmap().ptr<float>(cy)[cx] < 0.7f || sat.crop_box()[2] < 100 || std::fabs(sat.ebox()[2] - 100.f) > 1.f) return 17;
    ssk::c_weigthed_average dacc;
    ssk::Mat nomask;
    if (!sat.derotate_and_add(dacc, ref, nomask, 300.0, 0.8, false) || !sat.derotate_and_add(dacc, ref, nomask, 0.0, 1.0, true) ||
        dacc.accumulated_frames() != 2) { std::fprintf(stderr, "saturn derotate_and_add: %s\n", ssk_last_error()); return 17; }
    if (ssk::set_stream_ordered(true) || !ssk::set_stream_ordered(false) || !ssk::device_synchronize()) return 17;   // mode toggles, host frames: blocking as before
    std::printf("adapter_smoke: c_saturn_derotation_remap ebox %.1f x %.1f @ %.2f deg, crop %d x %d ok\n", sat.ebox()[2], sat.ebox()[3], sat.ebox()[4],
                sat.crop_box()[2], sat.crop_box()[3]);
  }
  return (worst <= 0.1 && err <= 5e-3 && cnt > W * H / 2) ? 0 : 1;
}
