// Exercises serstacker_b200/host/ssk_adapter.h the way a reference call site would (register_frame -> remap ->
// accumulate, c_image_stacking_pipeline.cc:1358-1862), on an analytic scene with known sub-pixel shifts.
//   adapter_smoke --no-gpu : option defaults / handle-free calls only (CPU test)
//   adapter_smoke          : full path on cuda:0; exit code 0 iff the recovered shifts are within 0.1 px
#include <cmath>
#include <cstdio>
#include <cstring>

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
  return (worst <= 0.1 && err <= 5e-3 && cnt > W * H / 2) ? 0 : 1;
}
